"""The scoped hot path as one callable: UniT's per-image RoI stage around an out-of-scope box head.

``RoIStage`` drives a registered head (``WSROIHeadFineTune`` / ``WSROIHeadNoMeta`` ...) exactly along the reference's
call stacks (SURVEY.md section 3.1 inference, 3.3 fine-tune step) but with the res5 box head factored out: the
caller supplies ``box_head_fn(pooled) -> (box_features, weak_box_features)`` (stock PyTorch in production; a synthetic
stand-in in bench.py and the tests, because res5 is outside the hand-written path).  Everything else -- fused
IoU+match, label+sample, ROIAlign forward/backward, similarity + transfer, losses' box targets, softmax/decode,
filter, class-wise NMS -- runs in the kernels of libunit_b200.so.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import layers, ops
from .distributed import FlatGradBucket
from .structures import Boxes, Instances


class RoIStage:
    def __init__(self, head, box_head_fn: Callable[[torch.Tensor], Tuple[torch.Tensor, Optional[torch.Tensor]]],
                 bucket: Optional[FlatGradBucket] = None):
        self.head = head
        self.box_head_fn = box_head_fn
        self.bucket = bucket

    # ------------------------------------------------------------------------------------------- inference
    @torch.no_grad()
    def infer(self, features: torch.Tensor, proposals: List[Instances]):
        """pool -> box head -> similarity -> transfer -> softmax/decode -> filter -> NMS -> top-k."""
        head = self.head
        head.move_mappings_to_gpu()
        pooled = head.box_pooler([features], [p.proposal_boxes for p in proposals])
        x, xw = self.box_head_fn(pooled)
        sim = head.get_similarity_matrices(x)
        predictions, _ = head.box_predictor(x, supervised_branch_x_weak=xw, novel_classes=head._novel_classes_tensor,
                                            base_classes=head._base_classes_tensor, similarity=sim)
        return head.box_predictor.inference(predictions, proposals)

    # ------------------------------------------------------------------------------------------- fine-tune step
    def train_step(self, features: torch.Tensor, proposals: List[Instances], targets: List[Instances],
                   grad_pooled_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None):
        """label+sample -> ROIAlign fwd -> box head -> transfer (training too) -> CE + smooth-L1 ->
        backward into the flat bucket of cls_score_ft / bbox_pred_ft -> ROIAlign backward (dL/dfeatures, as in base
        training where BACKBONE.FREEZE_AT=2 leaves res3/res4 trainable) -> one all-reduce of the bucket.

        ``grad_pooled_fn(pooled)`` supplies dL/dpooled when the box head is a stand-in that does not backpropagate.
        Returns (loss tensor, dL/dfeatures)."""
        head = self.head
        head.move_mappings_to_gpu()
        if self.bucket is not None:
            self.bucket.zero_()
        sampled = head.label_and_sample_proposals(proposals, targets)
        boxes = [p.proposal_boxes for p in sampled]
        rois = layers.cat([torch.cat((b.tensor.new_full((len(b), 1), float(i)), b.tensor), dim=1)
                           for i, b in enumerate(boxes)])
        pool = head.box_pooler
        pooled = ops.roi_align_forward(features, rois, pool.output_size, pool.scales[0], pool.sampling_ratio,
                                       pool.aligned, True)
        x, xw = self.box_head_fn(pooled)
        sim = head.get_similarity_matrices(x)
        predictions, _ = head.box_predictor(x, supervised_branch_x_weak=xw, novel_classes=head._novel_classes_tensor,
                                            base_classes=head._base_classes_tensor, similarity=sim)
        losses = head.box_predictor.losses(predictions, sampled)
        loss = losses["loss_cls"] + losses["loss_box_reg"]
        loss.backward()
        grad_feat = None
        if grad_pooled_fn is not None:
            grad_feat = ops.roi_align_backward(grad_pooled_fn(pooled), rois, features.shape, pool.scales[0],
                                               pool.sampling_ratio, pool.aligned, True)
        if self.bucket is not None:
            self.bucket.all_reduce_mean()
        return loss, grad_feat
