"""Boundary types of the RoI stage: ``Boxes``, ``Instances``, ``ShapeSpec``, ``pairwise_iou``.

Same names, fields and behaviour as the Detectron2 containers the reference passes around
(modeling/roi_heads/roi_heads.py:17,356,417,430,548-550) so that the heads read like
the reference's.  Containers are host-side bookkeeping over device tensors; ``pairwise_iou`` runs the CUDA kernel.
"""
from __future__ import annotations

import itertools
from typing import Any, Dict, List, NamedTuple, Optional, Tuple, Union

import torch

from . import ops


class ShapeSpec(NamedTuple):
    channels: Optional[int] = None
    height: Optional[int] = None
    width: Optional[int] = None
    stride: Optional[int] = None


class Boxes:
    """N x 4 float32 boxes, (x1, y1, x2, y2) absolute coordinates."""

    def __init__(self, tensor: torch.Tensor):
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32)
        tensor = tensor.to(torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape((-1, 4))
        if tensor.dim() != 2 or tensor.size(-1) != 4:
            raise ValueError(f"Boxes expects [N,4], got {tuple(tensor.shape)}")
        self.tensor = tensor

    def __len__(self) -> int:
        return self.tensor.shape[0]

    @property
    def device(self) -> torch.device:
        return self.tensor.device

    def clone(self) -> "Boxes":
        return Boxes(self.tensor.clone())

    def to(self, *args, **kwargs) -> "Boxes":
        return Boxes(self.tensor.to(*args, **kwargs))

    def area(self) -> torch.Tensor:
        t = self.tensor
        return (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])

    def clip(self, box_size: Tuple[int, int]) -> None:
        h, w = box_size
        t = self.tensor
        self.tensor = torch.stack((t[:, 0].clamp(min=0, max=w), t[:, 1].clamp(min=0, max=h),
                                   t[:, 2].clamp(min=0, max=w), t[:, 3].clamp(min=0, max=h)), dim=-1)

    def nonempty(self, threshold: float = 0.0) -> torch.Tensor:
        t = self.tensor
        return ((t[:, 2] - t[:, 0]) > threshold) & ((t[:, 3] - t[:, 1]) > threshold)

    def scale(self, scale_x: float, scale_y: float) -> None:
        self.tensor[:, 0::2] *= scale_x
        self.tensor[:, 1::2] *= scale_y

    def __getitem__(self, item) -> "Boxes":
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        b = self.tensor[item]
        if b.dim() != 2:
            raise IndexError(f"Indexing on Boxes with {item} failed to return a matrix")
        return Boxes(b)

    @classmethod
    def cat(cls, boxes_list: List["Boxes"]) -> "Boxes":
        if len(boxes_list) == 0:
            return cls(torch.empty(0))
        return cls(torch.cat([b.tensor for b in boxes_list], dim=0))

    def __iter__(self):
        yield from self.tensor

    def __repr__(self) -> str:
        return f"Boxes({self.tensor})"


def pairwise_iou(boxes1: Boxes, boxes2: Boxes) -> torch.Tensor:
    """[D2] pairwise_iou: IoU matrix [len(boxes1), len(boxes2)] computed by unit_pairwise_iou."""
    return ops.pairwise_iou(boxes1.tensor, boxes2.tensor)


class Instances:
    """Fields of equal length attached to one image of size (height, width)."""

    def __init__(self, image_size: Tuple[int, int], **kwargs: Any):
        object.__setattr__(self, "_image_size", image_size)
        object.__setattr__(self, "_fields", {})
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self) -> Tuple[int, int]:
        return self._image_size

    def __setattr__(self, name: str, val: Any) -> None:
        if name.startswith("_"):
            object.__setattr__(self, name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name: str) -> Any:
        fields = object.__getattribute__(self, "_fields")
        if name not in fields:
            raise AttributeError(f"Cannot find field '{name}' in the given Instances!")
        return fields[name]

    def set(self, name: str, value: Any) -> None:
        if len(self._fields) and len(value) != len(self):
            raise ValueError(f"Adding a field of length {len(value)} to a Instances of length {len(self)}")
        self._fields[name] = value

    def has(self, name: str) -> bool:
        return name in self._fields

    def get(self, name: str) -> Any:
        return self._fields[name]

    def remove(self, name: str) -> None:
        del self._fields[name]

    def get_fields(self) -> Dict[str, Any]:
        return self._fields

    def to(self, *args: Any, **kwargs: Any) -> "Instances":
        out = Instances(self._image_size)
        for k, v in self._fields.items():
            out.set(k, v.to(*args, **kwargs) if hasattr(v, "to") else v)
        return out

    def __getitem__(self, item: Union[int, slice, torch.Tensor]) -> "Instances":
        if type(item) is int:
            if item >= len(self) or item < -len(self):
                raise IndexError("Instances index out of range!")
            item = slice(item, None, len(self))
        out = Instances(self._image_size)
        for k, v in self._fields.items():
            out.set(k, v[item])
        return out

    def __len__(self) -> int:
        for v in self._fields.values():
            return len(v)
        raise NotImplementedError("Empty Instances does not support __len__!")

    @staticmethod
    def cat(instance_lists: List["Instances"]) -> "Instances":
        assert len(instance_lists) > 0
        if len(instance_lists) == 1:
            return instance_lists[0]
        size = instance_lists[0].image_size
        out = Instances(size)
        for k in instance_lists[0]._fields.keys():
            values = [i.get(k) for i in instance_lists]
            v0 = values[0]
            if isinstance(v0, torch.Tensor):
                values = torch.cat(values, dim=0)
            elif isinstance(v0, list):
                values = list(itertools.chain(*values))
            elif hasattr(type(v0), "cat"):
                values = type(v0).cat(values)
            else:
                raise ValueError(f"Unsupported type {type(v0)} for concatenation")
            out.set(k, values)
        return out

    def __repr__(self) -> str:
        n = len(self) if len(self._fields) else 0
        return (f"Instances(num_instances={n}, image_height={self._image_size[0]}, "
                f"image_width={self._image_size[1]}, fields=[{', '.join(self._fields)}])")
