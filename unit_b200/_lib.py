"""ctypes binding of libunit_b200.so (the C ABI declared in include/unit_b200.h).

There is no CPU fallback: if the library is missing or a call fails this module raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_size_t, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UNIT_B200_LIB") or os.path.join(_HERE, "libunit_b200.so")  # override: kernel-variant experiments

UNIT_F32, UNIT_BF16 = 0, 1


class TransferParams(Structure):
    _fields_ = [
        ("R", c_int), ("K", c_int), ("B", c_int), ("Nn", c_int),
        ("vis_threshold", c_float),
        ("wv_cls", c_float), ("wv_bbox", c_float), ("wv_seg", c_float),
        ("norm_cls", c_int), ("norm_bbox", c_int), ("norm_seg", c_int),
        ("do_transfer", c_int), ("novel_neg_inf", c_int), ("static_per_roi", c_int),
        ("ld_delta_scores", c_int), ("ld_proposal_deltas", c_int), ("ld_ft_scores", c_int), ("ld_ft_deltas", c_int),
        ("ld_vis_logits", c_int), ("ld_weak_scores", c_int),
    ]


P = c_void_p
_SIGNATURES = {
    "unit_version": (c_int, []),
    "unit_last_error": (c_char_p, []),
    "unit_source_digest": (c_char_p, []),
    "unit_launch_count": (c_ulonglong, []),
    "unit_roi_align_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "unit_roi_align_fwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int,
                                   c_int, c_int, P, c_size_t, P]),
    "unit_roi_align_bwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int,
                                   c_int, c_int, P, c_size_t, P]),
    "unit_pairwise_iou": (c_int, [P, P, P, c_int, c_int, P]),
    "unit_matcher": (c_int, [P, c_int, c_int, POINTER(c_float), POINTER(c_int), c_int, c_int, P, P, P, P, c_size_t,
                             P]),
    "unit_iou_match": (c_int, [P, P, P, P, c_int, c_int, POINTER(c_float), POINTER(c_int), c_int, P, P, P, P]),
    "unit_label_proposals": (c_int, [P, P, P, P, P, c_int, c_int, c_int, P, P, P, P, P]),
    "unit_sample_gather": (c_int, [P, P, P, P, P, P, P, P, P, P, c_int, c_int, P, P, P, P, P, P, P, P, P, P, P, P]),
    "unit_append_gt": (c_int, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int), POINTER(c_int),
                               c_int, c_float, P, P, P]),
    "unit_softmax_decode": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_float, c_float, c_float, c_float, c_float,
                                    P]),
    "unit_box_get_deltas": (c_int, [P, P, P, c_int, c_float, c_float, c_float, c_float, P]),
    "unit_fastrcnn_loss": (c_int, [P, P, P, P, P, c_int, c_int, c_float, c_float, c_float, c_float, c_float, P, P, P, P,
                                   c_size_t, P]),
    "unit_detect_filter": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_float, P, P, P, P, P, P, c_size_t, P]),
    "unit_nms_workspace_bytes": (c_size_t, [c_int, c_int]),
    "unit_detect_nms": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_float, c_int, c_int, P, P, P, P, P, P,
                                c_size_t, P]),
    "unit_batched_nms": (c_int, [P, P, P, c_int, c_float, c_int, c_int, P, P, P, c_size_t, P]),
    "unit_lingual_similarity": (c_int, [P, P, P, P, c_int, c_int, c_int, P, P, P]),
    "unit_similarity_transfer": (c_int, [POINTER(TransferParams)] + [P] * 17 + [P]),
    "unit_similarity_transfer_bwd": (c_int, [POINTER(TransferParams), P, P, P, P, P, P, P, c_int, P, P, P]),
    "unit_similarity_transfer_bwd_vis": (c_int, [POINTER(TransferParams)] + [P] * 10 + [P]),
    "unit_mask_transfer": (c_int, [P, P, c_int, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "unit_mask_paste": (c_int, [P, P, c_int, c_int, c_int, c_int, c_float, P, P]),
    "unit_predictor_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "unit_predictor_gemm": (c_int, [P, P, P, P, c_int, c_int, c_int, P, c_size_t, P]),
    "unit_predictor_gemm2_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "unit_predictor_gemm2": (c_int, [P, P, P, P, c_int, c_int, P, P, P, P, c_int, c_int, c_int, c_int, P, c_size_t, P]),
    "unit_predictor_wgrad_workspace_bytes": (c_size_t, [c_int, c_int]),
    "unit_predictor_wgrad": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_void_p),
                                     POINTER(c_void_p), POINTER(c_void_p), c_int, P, c_size_t, P]),
    "unit_fastrcnn_loss_packed": (c_int, [P, P, P, P, P, c_int, c_int, c_float, c_float, c_float, c_float, c_float, P, P,
                                          c_int, P, c_size_t, P]),
    "unit_fastrcnn_row_losses": (c_int, [P, P, P, P, P, c_int, c_int, c_float, c_float, c_float, c_float, c_float, c_int,
                                         P, P, P, P, P]),
    "unit_boxes_to_rois": (c_int, [P, P, c_int, c_int, P, P]),
    "unit_mil_loss_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "unit_mil_loss": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_float, P, P, P, P, P, P, c_size_t, P]),
    "unit_oicr_targets": (c_int, [P, c_int, P, P, P, c_int, c_int, c_int, P, P, c_int, c_float, P, P, P, P, P]),
    "unit_weighted_ce_loss": (c_int, [P, P, P, c_int, c_int, P, P, P, c_size_t, P]),
}
_OPTIONAL = {}

_lib = None


class UnitLibraryError(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGNATURES)


def lib() -> ctypes.CDLL:
    """Load libunit_b200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if "UNIT_B200_LIB" not in os.environ:
        # The library normally travels with the tree.  It carries the digest of the sources it was built from: a
        # missing or STALE library (sources changed after a pull) is rebuilt when nvcc is present, and a stale one
        # that cannot be rebuilt is refused instead of silently running old kernels behind a new ABI.
        from . import build as _build
        want, have = _build.source_digest(), _build.embedded_digest(LIB_PATH)
        if have != want:
            try:
                _build.build()
            except Exception as e:  # no nvcc / compile error
                if os.path.exists(LIB_PATH):
                    raise UnitLibraryError(
                        f"{LIB_PATH} was built from other sources (digest {have}, csrc/ is {want}) and could not be "
                        f"rebuilt: {e}") from e
    if not os.path.exists(LIB_PATH):
        raise UnitLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m unit_b200.build` (needs nvcc). "
            "unit_b200 has no CPU or PyTorch fallback."
        )
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in _OPTIONAL.items():
        if hasattr(handle, name):
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
    _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().unit_last_error()
        raise UnitLibraryError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(lib().unit_launch_count())
