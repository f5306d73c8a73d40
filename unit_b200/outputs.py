"""[D2 <= 0.4] ``FastRCNNOutputs`` and UniT's three variants of it:

  FastRCNNOutputs            losses over sampled proposals, fast_rcnn.py:438-445 (the fused CE + smooth-L1 kernel)
  FastRCNNOutputsReduction   the same losses UNREDUCED (per RoI / per foreground coordinate), fast_rcnn.py:24-101;
                             used for the pseudo-labelled weak images at modeling/meta_arch/rcnn.py:615
  FastRCNNOutputsNLL         nll_loss on log-probabilities instead of cross_entropy on logits, fast_rcnn.py:103-115
  FastRCNNOutputsRegression  per-RoI weighted cross entropy + the box loss, named loss_regression_{cls,bbox},
                             fast_rcnn.py:117-130 / weak_detector_fast_rcnn.py:23-37 (the OICR regression branches)

Same constructor arguments, attribute names and return values as the reference; the arithmetic runs in the kernels of
libunit_b200.so (csrc/detect.cu: fastrcnn_loss_kernel, row_losses_kernel; csrc/weak.cu: weighted_ce_kernel).
"""
from __future__ import annotations

from typing import List

import torch

from . import layers, ops
from .structures import Boxes, Instances


class FastRCNNOutputs:
    def __init__(self, box2box_transform, pred_class_logits, pred_proposal_deltas, proposals: List[Instances],
                 smooth_l1_beta=0.0, box_reg_loss_type="smooth_l1"):
        self.box2box_transform = box2box_transform
        self.num_preds_per_image = [len(p) for p in proposals]
        self.pred_class_logits = pred_class_logits
        self.pred_proposal_deltas = pred_proposal_deltas
        self.smooth_l1_beta = smooth_l1_beta
        self.box_reg_loss_type = box_reg_loss_type
        self.image_shapes = [x.image_size for x in proposals]
        if box_reg_loss_type != "smooth_l1":
            raise NotImplementedError("only the smooth_l1 box loss is used by the reference YAMLs")
        if len(proposals):
            self.proposals = Boxes(layers.cat([p.proposal_boxes.tensor for p in proposals]))
            assert not self.proposals.tensor.requires_grad, "Proposals should not require gradients!"
            if proposals[0].has("gt_boxes"):
                self.gt_boxes = Boxes(layers.cat([p.gt_boxes.tensor for p in proposals]))
                assert proposals[0].has("gt_classes")
                self.gt_classes = layers.cat([p.gt_classes for p in proposals])
        else:
            self.proposals = Boxes(torch.zeros(0, 4, device=self.pred_proposal_deltas.device))
        self._no_instances = len(proposals) == 0

    # -- reduced losses: one fused launch for both ---------------------------------------------------------------
    def _fused(self):
        if "_fused_losses" not in self.__dict__:
            self._fused_losses = ops.fastrcnn_loss(self.pred_class_logits, self.pred_proposal_deltas,
                                                   self.proposals.tensor, self.gt_boxes.tensor, self.gt_classes,
                                                   self.box2box_transform.weights, self.smooth_l1_beta)
        return self._fused_losses

    def softmax_cross_entropy_loss(self):
        if self._no_instances:
            return 0.0 * self.pred_class_logits.sum()
        return self._fused()[0]

    def box_reg_loss(self):
        if self._no_instances:
            return 0.0 * self.pred_proposal_deltas.sum()
        return self._fused()[1]

    def _predict_boxes(self):
        return self.box2box_transform.apply_deltas(self.pred_proposal_deltas, self.proposals.tensor)

    def losses(self):
        return {"loss_cls": self.softmax_cross_entropy_loss(), "loss_box_reg": self.box_reg_loss()}

    def predict_boxes(self):
        return self._predict_boxes().split(self.num_preds_per_image, dim=0)

    def predict_probs(self):
        probs, _ = ops.softmax_decode(self.pred_class_logits, None, None, want_boxes=False)
        return probs.split(self.num_preds_per_image, dim=0)

    # -- unreduced pieces shared by the variants ------------------------------------------------------------------
    def _rows(self, nll=False):
        key = "_rows_nll" if nll else "_rows_ce"
        if key not in self.__dict__:
            self.__dict__[key] = ops.fastrcnn_row_losses(self.pred_class_logits, self.pred_proposal_deltas,
                                                         self.proposals.tensor, self.gt_boxes.tensor, self.gt_classes,
                                                         self.box2box_transform.weights, self.smooth_l1_beta, nll=nll)
        return self.__dict__[key]


class FastRCNNOutputsReduction(FastRCNNOutputs):
    """fast_rcnn.py:24-101: ``reduction="none"`` -- loss_cls is [R]; loss_box_reg is [n_fg, 4] / R."""

    def softmax_cross_entropy_loss(self):
        if self._no_instances:
            return 0.0 * self.pred_class_logits.sum()
        return self._rows()[0]

    def box_reg_loss(self):
        if self._no_instances:
            return 0.0 * self.pred_proposal_deltas.sum()
        bg = self.pred_class_logits.shape[1] - 1
        fg = ((self.gt_classes >= 0) & (self.gt_classes < bg)).nonzero().squeeze(1)  # the reference's nonzero_tuple
        return self._rows()[1][fg] / self.gt_classes.numel()


class FastRCNNOutputsNLL(FastRCNNOutputs):
    """fast_rcnn.py:103-115: the class input already holds log-probabilities."""

    def softmax_cross_entropy_loss(self):
        if self._no_instances:
            return 0.0 * self.pred_class_logits.sum()
        return self._rows(nll=True)[0].mean()

    def box_reg_loss(self):
        if self._no_instances:
            return 0.0 * self.pred_proposal_deltas.sum()
        return self._rows(nll=True)[1].sum() / self.gt_classes.numel()


class FastRCNNOutputsRegression(FastRCNNOutputs):
    """fast_rcnn.py:117-130 / weak_detector_fast_rcnn.py:23-37: mean(CE * weights), and the usual box loss."""

    def __init__(self, box2box_transform, pred_class_logits, pred_proposal_deltas, proposals, weights,
                 smooth_l1_beta=0.0, box_reg_loss_type="smooth_l1"):
        super().__init__(box2box_transform, pred_class_logits, pred_proposal_deltas, proposals, smooth_l1_beta,
                         box_reg_loss_type)
        self.weights = weights

    def softmax_cross_entropy_loss(self):
        if self._no_instances:
            return 0.0 * self.pred_class_logits.sum()
        return ops.weighted_ce_loss(self.pred_class_logits, self.gt_classes, self.weights)

    def box_reg_loss(self):
        if self._no_instances:
            return 0.0 * self.pred_proposal_deltas.sum()
        return self._rows()[1].sum() / self.gt_classes.numel()

    def losses(self):
        return {"loss_regression_cls": self.softmax_cross_entropy_loss(), "loss_regression_bbox": self.box_reg_loss()}


def predict_boxes_for_gt_classes(box2box_transform, predictions, proposals: List[Instances]):
    """[D2] FastRCNNOutputLayers.predict_boxes_for_gt_classes (called at roi_heads.py:97,426,535,628,753,865 when
    TRAIN_ON_PRED_BOXES): decode every class's deltas, keep the box of each proposal's ground-truth class."""
    if not len(proposals):
        return []
    _, proposal_deltas = predictions
    boxes = layers.cat([p.proposal_boxes.tensor for p in proposals])
    n = boxes.shape[0]
    pred = box2box_transform.apply_deltas(proposal_deltas, boxes)
    k = pred.shape[1] // 4
    if k > 1:
        gt = layers.cat([p.gt_classes for p in proposals]).clamp(0, k - 1)
        pred = pred.view(n, k, 4)[torch.arange(n, device=pred.device), gt]
    return pred.split([len(p) for p in proposals])
