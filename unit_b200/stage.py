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
        self._graphs: Dict[tuple, "_StepGraphs"] = {}
        self.max_graphs = 8  # captured (input buffers, sample counts) keys kept alive; each holds its pooled tensors
        self.graph_launches = 0  # library kernels executed through graph replays (not visible to unit_launch_count)

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

    @torch.no_grad()
    def infer_graphed(self, features: torch.Tensor, proposals: List[Instances], padded: bool = False):
        """``infer`` with everything up to the padded detections replayed from one CUDA graph (keyed by the input
        buffers, which the caller must reuse); the only host work per call is the read of the detection counts.  The
        returned Instances are views of the graph's static outputs: consume them before the next call with that key.
        ``padded=True`` skips that host read and returns the device tuple (det_boxes [n,topk,4], det_scores,
        det_classes, det_roi, det_counts) -- what ``distributed.gather_detections`` takes."""
        head = self.head
        head.move_mappings_to_gpu()
        key = ("infer", features.data_ptr(), tuple(features.shape),
               tuple(p.proposal_boxes.tensor.data_ptr() for p in proposals), tuple(len(p) for p in proposals))
        st = self._graphs.get(key)
        if st is None:
            while len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))

            def device_part():
                pooled = head.box_pooler([features], [p.proposal_boxes for p in proposals])
                x, xw = self.box_head_fn(pooled)
                sim = head.get_similarity_matrices(x)
                predictions, _ = head.box_predictor(x, supervised_branch_x_weak=xw,
                                                    novel_classes=head._novel_classes_tensor,
                                                    base_classes=head._base_classes_tensor, similarity=sim)
                return head.box_predictor.inference_device(predictions, proposals)

            dev = features.device
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):  # warm-up outside capture: lazy handles, workspaces, cached constants
                device_part()
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                dets = device_part()
            st = (graph, dets)
            self._graphs[key] = st
        st[0].replay()
        if padded:
            return st[1]
        return layers.instances_from_detections(st[1], [p.image_size for p in proposals])

    @torch.no_grad()
    def infer_tta(self, features: Sequence[torch.Tensor], aug_proposals: Sequence[List[Instances]],
                  proposals: List[Instances]):
        """rcnn.py:495-527: one pass per augmentation (its own feature map and transformed proposals, same RoI order)
        with ``tta=True``, then ``inference_tta`` on the un-augmented ``proposals``."""
        head = self.head
        head.move_mappings_to_gpu()
        outs = []
        for feats, props in zip(features, aug_proposals):
            pooled = head.box_pooler([feats], [p.proposal_boxes for p in props])
            x, xw = self.box_head_fn(pooled)
            sim = head.get_similarity_matrices(x)
            predictions, _ = head.box_predictor(x, supervised_branch_x_weak=xw,
                                                novel_classes=head._novel_classes_tensor,
                                                base_classes=head._base_classes_tensor, similarity=sim)
            res, _ = head.box_predictor.inference(predictions, props, tta=True)
            outs.append(res)
        return head.box_predictor.inference_tta(outs, proposals)

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
        self._seed(features.device)
        sampled = head.label_and_sample_proposals(proposals, targets)
        loss, rois, pooled = self._forward_backward(features, sampled)
        work = self.bucket.all_reduce_mean(async_op=True) if self.bucket is not None else None
        grad_feat = self._roi_backward(features, rois, pooled, grad_pooled_fn)  # overlaps the all-reduce
        if self.bucket is not None:
            self.bucket.finish(work)
        return loss, grad_feat

    def _forward_backward(self, features, sampled: List[Instances]):
        """Sampled RoIs -> ROIAlign fwd -> box head -> transfer -> losses -> backward into the bucket.  Shapes are
        fixed by the sample counts and nothing synchronises (capturable)."""
        head = self.head
        boxes = [p.proposal_boxes for p in sampled]
        rois = ops.boxes_to_rois(layers.cat([b.tensor for b in boxes]),
                                 ops.offsets_from_counts([len(b) for b in boxes], features.device))
        pool = head.box_pooler
        pooled = ops.roi_align_forward(features, rois, pool.output_size, pool.scales[0], pool.sampling_ratio,
                                       pool.aligned, True)
        x, xw = self.box_head_fn(pooled)
        losses, _ = head.box_losses(x, xw, sampled)  # ONE fused node in the shipped fine-tune setting
        total = getattr(losses, "total", None)
        if total is not None and self._bucket_is_exactly(head.box_predictor):
            # fused node + a bucket that holds exactly its four gradients: no loss add, no seed fill (persistent ones),
            # no zero fill of the bucket (the weight-gradient kernel writes instead of accumulating)
            self.bucket.bind()
            one = self._seed(total.device)
            with ops.overwrite_bound_grads():
                torch.autograd.backward([losses["loss_cls"], losses["loss_box_reg"]], [one, one])
            return total, rois, pooled
        if self.bucket is not None:
            self.bucket.zero_()  # gradients are not touched before this point
        loss = losses["loss_cls"] + losses["loss_box_reg"]
        loss.backward()
        return loss.detach(), rois, pooled

    def _seed(self, device) -> torch.Tensor:
        one = self.__dict__.get("_one")
        if one is None or one.device != device:
            one = self._one = torch.ones((), dtype=torch.float32, device=device)
        return one

    def _bucket_is_exactly(self, predictor) -> bool:
        if self.bucket is None:
            return False
        want = [getattr(predictor, n, None) for n in ("cls_score_ft", "bbox_pred_ft")]
        if any(m is None for m in want):
            return False
        ids = {id(m.weight) for m in want} | {id(m.bias) for m in want}
        # every one of them must still be trainable: a tensor frozen after the bucket was built gets no gradient
        # written, and without the zero fill its slice would be all-reduced as it was left
        return {id(p) for p in self.bucket.params} == ids and all(p.requires_grad for p in self.bucket.params)

    def _roi_backward(self, features, rois, pooled, grad_pooled_fn):
        if grad_pooled_fn is None:
            return None
        pool = self.head.box_pooler
        return ops.roi_align_backward(grad_pooled_fn(pooled), rois, features.shape, pool.scales[0],
                                      pool.sampling_ratio, pool.aligned, True)

    # ------------------------------------------------------------------------------------------- CUDA-graph replay
    def train_step_graphed(self, features: torch.Tensor, proposals: List[Instances], targets: List[Instances],
                           grad_pooled_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None):
        """``train_step`` with the launch-bound parts replayed from CUDA graphs.

        The step is cut at its one unavoidable host round trip (the reference's ``subsample_labels`` needs the fg/bg
        counts on the host to draw ``randperm``): graph A = append GT + fused IoU/match + label; then the count read,
        the host draw and ONE host->device copy of the permutations into a fixed-layout buffer; graph B = gather ->
        ROIAlign fwd -> box head -> similarity/transfer -> losses -> backward; graph C = ROIAlign bwd, replayed while
        the gradient bucket's all-reduce runs on NCCL's stream.  Graphs are keyed by the
        input buffers (the caller must reuse them: same pointers, same shapes) and by the per-image sample counts; a
        step whose counts differ from the captured ones runs graph A + the eager remainder.  The returned tensors are
        the graph's static outputs: they are overwritten by the next replay with the same key."""
        head = self.head
        head.move_mappings_to_gpu()
        self._seed(features.device)  # created outside any capture
        key = self._step_key(features, proposals, targets)
        st = self._graphs.get(key)
        if st is None:
            while len(self._graphs) >= self.max_graphs:  # callers that do not reuse buffers must not leak graphs
                self._graphs.pop(next(iter(self._graphs)))
            st = _StepGraphs(self, features, proposals, targets, grad_pooled_fn)
            self._graphs[key] = st
        return st.run()

    @staticmethod
    def _step_key(features, proposals, targets):
        return (features.data_ptr(), tuple(features.shape),
                tuple(p.proposal_boxes.tensor.data_ptr() for p in proposals), tuple(len(p) for p in proposals),
                tuple(t.gt_boxes.tensor.data_ptr() for t in targets), tuple(len(t) for t in targets),
                tuple(t.gt_classes.data_ptr() for t in targets))

    def prefetch_labels(self, features: torch.Tensor, proposals: List[Instances], targets: List[Instances],
                        after: Optional[torch.cuda.Event] = None) -> bool:
        """Start the labelling half of the NEXT ``train_step_graphed`` call on these buffers -- graph A and the read
        of its fg/bg counts into pinned memory -- on a side stream, so that it overlaps the step that is running now.
        Labelling depends only on proposals and ground truth (not on the weights), which is what makes this legal;
        the ``randperm`` draw still happens inside ``train_step_graphed``, in call order, so sampled indices are the
        ones the unpipelined sequence would draw.  The buffers must already hold the next batch, or ``after`` must be
        an event that marks the end of their host->device copy.  They must not be the buffers of the step in flight.
        Returns False (and does nothing) when this key has not been captured yet."""
        st = self._graphs.get(self._step_key(features, proposals, targets))
        if st is None or st._pending is not None:
            return False
        st.prefetch(after)
        return True


class _StepGraphs:
    """The two captured graphs of one (input buffers, sample counts) key, with their static tensors."""

    def __init__(self, stage: RoIStage, features, proposals, targets, grad_pooled_fn):
        from . import _lib
        self.stage, self.features, self.proposals, self.targets = stage, features, proposals, targets
        self.grad_pooled_fn = grad_pooled_fn
        head = stage.head
        self.kw = dict(num_classes=head.num_classes, thresholds=head.proposal_matcher.user_thresholds,
                       labels=head.proposal_matcher.labels)
        dev = features.device
        self.launches = 0  # kernels of libunit_b200.so inside the two graphs (replays do not pass through the ABI)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside capture: lazy handles, workspaces, autograd engine
            lm = self._label()
            counts_h = lm.counts.cpu().tolist()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.capacity = int(sum(lm.prop_counts))
        self.host = torch.empty(2 * self.capacity + 4 * (len(proposals) + 1), dtype=torch.int64).pin_memory()
        self.devbuf = torch.zeros(self.host.numel(), dtype=torch.int64, device=dev)
        # ---- graph A
        self.graph_a = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph_a):
            self.lm = self._label()
        self.launches += _lib.launch_count() - n0
        self.graph_a.replay()
        draw = self._draw(self.lm.counts.cpu().tolist())
        self.sizes = draw.sizes
        self.devbuf.copy_(draw.host, non_blocking=True)
        # ---- graphs B and C (one eager pass first, on the side stream, with this step's draw)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            _, rois, pooled = self._after_draw(draw)
            stage._roi_backward(features, rois, pooled, grad_pooled_fn)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph_b = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph_b):
            self.loss, self.rois, self.pooled = self._after_draw(draw)
        self.graph_c = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_c, pool=self.graph_b.pool()):
            self.grad_feat = stage._roi_backward(features, self.rois, self.pooled, grad_pooled_fn)
        self.launches += _lib.launch_count() - n0
        self._pending = draw  # the capture call is itself a step: its draw is replayed once by run()
        # label prefetch (RoIStage.prefetch_labels): graph A on a side stream, counts into pinned memory
        self._counts_pinned = torch.empty(tuple(self.lm.counts.shape), dtype=self.lm.counts.dtype).pin_memory()
        self._counts_ready = torch.cuda.Event()
        self._done = torch.cuda.Event()      # last step on these buffers has finished reading graph A's outputs
        self._h2d_done = torch.cuda.Event()  # the draw's host buffer has been copied out
        self._done.record(torch.cuda.current_stream(dev))
        self._h2d_done.record(torch.cuda.current_stream(dev))
        self._prefetched = False

    def _label(self):
        head = self.stage.head
        props = self.proposals
        if head.proposal_append_gt:
            props = layers.add_ground_truth_to_proposals([t.gt_boxes for t in self.targets], props)
        self.props_with_gt = props
        return layers.label_match(props, self.targets, **self.kw)

    def _draw(self, counts_h):
        head = self.stage.head
        return layers.draw_permutations(counts_h, head.batch_size_per_image, head.positive_fraction,
                                        head.sampling_generator, capacity=self.capacity, out=self.host)

    def _after_draw(self, draw):
        with torch.no_grad():
            sampled, _, _ = layers.sample_from_draw(self.lm, draw, self.devbuf, self.props_with_gt, self.targets)
        return self.stage._forward_backward(self.features, sampled)

    def run(self):
        stage = self.stage
        main = torch.cuda.current_stream(self.features.device)
        if self._pending is not None:
            draw, self._pending = self._pending, None
        else:
            if self._prefetched:  # graph A already ran on the label stream
                self._prefetched = False
                self._counts_ready.synchronize()
                main.wait_event(self._counts_ready)
                counts_h = self._counts_pinned.tolist()
            else:
                self.graph_a.replay()
                counts_h = self.lm.counts.cpu().tolist()
            self._h2d_done.synchronize()  # the previous draw has left the pinned buffer (normally long ago)
            draw = self._draw(counts_h)
            self.devbuf.copy_(draw.host, non_blocking=True)
            self._h2d_done.record(main)
        graphed = draw.sizes == self.sizes  # rare otherwise: fewer candidates than the batch size -> other shapes
        if graphed:
            self.graph_b.replay()
            loss, rois, pooled = self.loss, self.rois, self.pooled
        else:
            loss, rois, pooled = self._after_draw(draw)
        work = stage.bucket.all_reduce_mean(async_op=True) if stage.bucket is not None else None
        if graphed:
            self.graph_c.replay()
            grad_feat = self.grad_feat
            stage.graph_launches += self.launches
        else:
            grad_feat = stage._roi_backward(self.features, rois, pooled, self.grad_pooled_fn)
        if stage.bucket is not None:
            stage.bucket.finish(work)
        self._done.record(main)
        return loss, grad_feat

    def prefetch(self, after: Optional[torch.cuda.Event] = None):
        stage = self.stage
        side = stage.__dict__.get("_label_stream")
        if side is None:
            side = stage._label_stream = torch.cuda.Stream(device=self.features.device)
        if after is not None:
            side.wait_event(after)
        side.wait_event(self._done)
        with torch.cuda.stream(side):
            self.graph_a.replay()
            self._counts_pinned.copy_(self.lm.counts, non_blocking=True)
            self._counts_ready.record(side)
        self._prefetched = True
