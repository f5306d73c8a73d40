"""CPU: host-side logic (config, registries, containers) and the C-ABI surface (symbols only, no compute)."""
import ctypes
import glob
import os
import re

import pytest
import torch

from conftest import ROOT

REF_CFG = "/root/reference/configs"


def test_config_yacs_semantics(tmp_path):
    from unit_b200.config import CfgNode, get_cfg, load_cfg

    cfg = load_cfg(os.path.join(ROOT, "configs", "voc_split1_ft.yaml"), ["MODEL.ROI_HEADS.SCORE_THRESH_TEST", "0.01",
                                                                         "INPUT.MIN_SIZE_TRAIN", "(640, 800)"])
    assert cfg.MODEL.ROI_HEADS.NAME == "WSROIHeadFineTune"            # own value
    assert cfg.MODEL.ROI_BOX_HEAD.POOLER_TYPE == "ROIAlignV2"          # inherited through _BASE_
    assert cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST == 0.01 and cfg.INPUT.MIN_SIZE_TRAIN == (640, 800)
    assert cfg.MODEL.ROI_HEADS.FINETUNE_TERMS.CLASSIFIER == ["lingual", "visual"]
    with pytest.raises(KeyError):
        cfg.merge_from_list(["MODEL.NOT_A_KEY", 1])
    with pytest.raises(ValueError):
        cfg.merge_from_list(["MODEL.ROI_HEADS.NUM_CLASSES", "twenty"])
    bad = tmp_path / "bad.yaml"
    bad.write_text("MODEL:\n  UNKNOWN_SECTION: {A: 1}\n")
    with pytest.raises(KeyError):
        load_cfg(str(bad))
    c2 = cfg.clone()
    c2.MODEL.ROI_HEADS.NUM_CLASSES = 3
    assert cfg.MODEL.ROI_HEADS.NUM_CLASSES == 20
    cfg.freeze()
    with pytest.raises(AttributeError):
        cfg.MODEL.MASK_ON = True
    assert isinstance(get_cfg().MODEL, CfgNode)


@pytest.mark.skipif(not os.path.isdir(REF_CFG), reason="reference checkout not present")
def test_reference_yamls_load_unchanged():
    from unit_b200.config import load_cfg

    files = [f for f in glob.glob(REF_CFG + "/**/*.yaml", recursive=True) if not f.endswith("Base-RCNN-C4.yaml")]
    assert len(files) == 24
    for f in files:
        cfg = load_cfg(f)
        assert cfg.MODEL.ROI_HEADS.NAME.startswith("WSROIHead")
        assert cfg.MODEL.ROI_HEADS.FAST_RCNN.NAME.startswith("SupervisedDetectorOutputs")


def test_registries_hold_reference_names():
    from unit_b200 import d2compat  # noqa: F401
    from unit_b200.registry import (FAST_RCNN_REGISTRY, ROI_BOX_HEAD_REGISTRY, ROI_HEADS_REGISTRY,
                                    ROI_MASK_HEAD_REGISTRY, WEAK_DETECTOR_FAST_RCNN_REGISTRY)

    for n in ("WSROIHeadNoMeta", "WSROIHeadFineTune", "WSROIHeadNoMetaWithMask", "WSROIHeadWithMaskFineTune",
              "WeakDetectorHead"):
        assert n in ROI_HEADS_REGISTRY
    for n in ("SupervisedDetectorOutputsBase", "SupervisedDetectorOutputsFineTune",
              "SupervisedDetectorOutputsWeakFineTune", "WeakDetectorOutputsBaseWrapper"):
        assert n in FAST_RCNN_REGISTRY
    assert "WeakDetectorOutputsBase" in WEAK_DETECTOR_FAST_RCNN_REGISTRY
    assert "Res5BoxHead" in ROI_BOX_HEAD_REGISTRY and "Res5BoxHeadWithMask" in ROI_BOX_HEAD_REGISTRY
    assert "MaskRCNNConvUpsampleHeadWithFineTune" in ROI_MASK_HEAD_REGISTRY


def test_heads_build_with_reference_state_dict_keys():
    from unit_b200.config import load_cfg
    from unit_b200.roi_heads import build_roi_heads
    from unit_b200.structures import ShapeSpec

    cfg = load_cfg(os.path.join(ROOT, "configs", "voc_split1_ft.yaml"),
                   ["MODEL.ROI_HEADS.EMBEDDING_PATH", os.path.join(ROOT, "tests", "golden", "glove_mean.pt")])
    head = build_roi_heads(cfg, {"res4": ShapeSpec(channels=1024, stride=16)})
    keys = set(head.state_dict().keys())
    for k in ("box_predictor.cls_score_delta.weight", "box_predictor.bbox_pred_delta.bias",
              "box_predictor.cls_score_ft.weight", "box_predictor.bbox_pred_ft.bias",
              "box_predictor.weak_detector_head.oicr_predictors.2.weight",
              "box_predictor.weak_detector_head.classifier_stream.weight", "box_predictor.embeddings.weight",
              "box_head.res5.0.conv1.weight", "box_head.res5.0.shortcut.norm.running_var",
              "weak_box_head.res5.2.conv3.weight"):
        assert k in keys, k
    trainable = sorted(n for n, p in head.named_parameters() if p.requires_grad)
    assert trainable == ["box_predictor.bbox_pred_ft.bias", "box_predictor.bbox_pred_ft.weight",
                         "box_predictor.cls_score_ft.bias", "box_predictor.cls_score_ft.weight"]
    assert head._coco_indexer_tensor.tolist() == [4, 1, 14, 8, 39, 5, 2, 15, 56, 19, 60, 16, 17, 3, 0, 58, 18, 57, 6, 62]
    assert head.box_pooler.aligned and head.box_pooler.output_size == (14, 14) and head.box_pooler.scales == (1 / 16,)


def test_stage_skips_the_bucket_zero_only_when_it_owns_exactly_the_ft_gradients():
    """RoIStage lets the weight-gradient kernel OVERWRITE the flat bucket (no zero fill, no loss add) only when the
    bucket holds exactly cls_score_ft / bbox_pred_ft -- the four tensors the fused node writes; host logic only."""
    import torch

    from unit_b200 import ops
    from unit_b200.config import load_cfg
    from unit_b200.distributed import FlatGradBucket
    from unit_b200.predictors import LossDict
    from unit_b200.roi_heads import build_roi_heads
    from unit_b200.stage import RoIStage
    from unit_b200.structures import ShapeSpec

    cfg = load_cfg(os.path.join(ROOT, "configs", "voc_split1_ft.yaml"),
                   ["MODEL.ROI_HEADS.EMBEDDING_PATH", os.path.join(ROOT, "tests", "golden", "glove_mean.pt")])
    head = build_roi_heads(cfg, {"res4": ShapeSpec(channels=1024, stride=16)})
    ft = [p for p in head.parameters() if p.requires_grad]
    assert len(ft) == 4
    stage = RoIStage(head, lambda pooled: (pooled, pooled), FlatGradBucket(ft))
    assert stage._bucket_is_exactly(head.box_predictor)
    ft[0].requires_grad_(False)  # frozen after the bucket was built: its slice would never be written
    assert not stage._bucket_is_exactly(head.box_predictor)
    ft[0].requires_grad_(True)
    extra = torch.nn.Parameter(torch.zeros(3))
    stage.bucket = FlatGradBucket(ft + [extra])
    assert not stage._bucket_is_exactly(head.box_predictor)  # something else lives in the bucket: zero + accumulate
    stage.bucket = None
    assert not stage._bucket_is_exactly(head.box_predictor)
    # the context only affects the backward that runs inside it
    assert ops._OVERWRITE_BOUND[0] is False
    with ops.overwrite_bound_grads():
        assert ops._OVERWRITE_BOUND[0] is True
    assert ops._OVERWRITE_BOUND[0] is False
    # the losses dict the trainer sums is unchanged by the extra attribute
    d = LossDict(loss_cls=torch.tensor(1.0), loss_box_reg=torch.tensor(2.0))
    d.total = torch.tensor(3.0)
    assert sorted(d) == ["loss_box_reg", "loss_cls"] and float(sum(d.values())) == 3.0
    assert getattr({"loss_cls": 1}, "total", None) is None


def test_containers():
    from unit_b200.structures import Boxes, Instances

    b = Boxes(torch.tensor([[0.0, 0.0, 10.0, 20.0], [5.0, 5.0, 4.0, 30.0]]))
    assert b.area().tolist() == [200.0, -25.0] and b.nonempty().tolist() == [True, False]
    b.clip((15, 8))
    assert b.tensor.tolist() == [[0.0, 0.0, 8.0, 15.0], [5.0, 5.0, 4.0, 15.0]]
    i = Instances((15, 8), proposal_boxes=b, objectness_logits=torch.tensor([1.0, 2.0]))
    assert len(i) == 2 and len(i[torch.tensor([True, False])]) == 1
    with pytest.raises(ValueError):
        i.scores = torch.zeros(3)
    c = Instances.cat([i, i])
    assert len(c) == 4 and isinstance(c.proposal_boxes, Boxes)
    assert Boxes(torch.empty(0)).tensor.shape == (0, 4)


def test_cat_returns_a_view_only_for_adjacent_row_blocks():
    """layers.cat: per-image slices of one flat buffer (what the sampling kernel hands out) come back as ONE view --
    no copy kernel in the step graph; anything else is an ordinary torch.cat."""
    from unit_b200.layers import cat

    flat = torch.arange(40, dtype=torch.float32).reshape(10, 4)
    parts = [flat[0:3], flat[3:3], flat[3:7], flat[7:10]]
    v = cat(parts)
    assert torch.equal(v, flat) and v.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
    assert cat([flat[2:5], flat[5:9]]).storage_offset() == 8                      # starts inside the buffer
    for other in ([flat[0:3], flat[4:7]],                                           # gap
                  [flat[3:7], flat[0:3]],                                           # out of order
                  [flat[0:3], flat[3:7].clone()],                                   # another buffer
                  [flat[0:3, :2], flat[3:7, :2]],                                   # not contiguous
                  [flat[0:3], flat[3:7].double()]):                                 # (dtype promotion -> torch.cat)
        out = cat(list(other))
        assert out.untyped_storage().data_ptr() != flat.untyped_storage().data_ptr()
        assert torch.equal(out, torch.cat(list(other)))
    w = torch.zeros(6, 2, requires_grad=True)
    g = cat([w[0:2], w[2:6]])
    assert g.grad_fn is not None and g.grad_fn.name().startswith("Cat")            # autograd keeps the real cat
    assert cat([flat[0:3]]) is not None and cat([flat[1:2]]).shape == (1, 4)


def test_draw_permutations_layout_and_generator_order():
    """layers.draw_permutations = the host half of [D2] subsample_labels: randperm(#pos) then randperm(#neg), image by
    image, from ONE generator; the int32 offset arrays ride in the tail of the int64 buffer (viewed as int32)."""
    from unit_b200.layers import draw_permutations

    counts = [(40, 900), (0, 1000), (300, 5)]
    gen = torch.Generator().manual_seed(9)
    d = draw_permutations(counts, 512, 0.25, gen)
    ref = torch.Generator().manual_seed(9)
    want = [torch.randperm(n, generator=ref) for pair in counts for n in pair]
    pos = torch.cat(want[0::2]); neg = torch.cat(want[1::2])
    assert d.pos_len == 340 and d.neg_len == 1905
    assert torch.equal(d.host[:340], pos) and torch.equal(d.host[340:340 + 1905], neg)
    offs = d.host[340 + 1905:].view(torch.int32)
    k = len(counts) + 1
    assert offs[:k].tolist() == [0, 40, 40, 340]                      # full positive permutation lengths
    assert offs[k:2 * k].tolist() == [0, 900, 1900, 1905]             # full negative permutation lengths
    assert offs[2 * k:3 * k].tolist() == [0, 40, 40, 168] == d.pso    # taken positives: min(#pos, 128)
    assert offs[3 * k:4 * k].tolist() == [0, 472, 984, 989] == d.nso  # taken negatives: min(#neg, 512 - pos)
    # fixed-capacity form (CUDA-graph replay): same draws into a caller-owned buffer
    buf = torch.full((2 * 2048 + 4 * k,), -1, dtype=torch.int64)
    d2 = draw_permutations(counts, 512, 0.25, torch.Generator().manual_seed(9), capacity=2048, out=buf)
    assert d2.host is buf and d2.pos_len == d2.neg_len == 2048
    assert torch.equal(buf[:340], pos) and torch.equal(buf[2048:2048 + 1905], neg)
    assert buf[4096:].view(torch.int32)[:4 * k].tolist() == offs[:4 * k].tolist()


def test_no_cpu_fallback():
    from unit_b200 import ops

    with pytest.raises(RuntimeError, match="CUDA"):
        ops.pairwise_iou(torch.zeros(1, 4), torch.zeros(1, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.roi_align(torch.zeros(1, 8, 4, 4), torch.zeros(1, 5), 14)


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads and exports exactly the entry points include/unit_b200.h declares."""
    from unit_b200 import _lib

    header = open(os.path.join(ROOT, "include", "unit_b200.h")).read()
    declared = set(re.findall(r"\b(unit_[a-z0-9_]+)\s*\(", header))
    declared.discard("unit_stream_t")
    assert len(declared) >= 20
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge

        ge.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in include/unit_b200.h but not exported"
    assert set(_lib.exported_symbols()) <= declared
    lib = _lib.lib()
    assert lib.unit_version() >= 100
    assert lib.unit_roi_align_workspace_bytes(4, 8, 4, 4, 16, 0) >= 16
    assert lib.unit_nms_workspace_bytes(2, 1000) > 36 * 2000


def test_transfer_params_struct_matches_the_header():
    """The one struct that crosses the ABI: same fields, same order, same C types in include/unit_b200.h and in the
    ctypes mirror (a silent drift would shift every later field)."""
    from unit_b200._lib import TransferParams

    hdr = open(os.path.join(ROOT, "include", "unit_b200.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} unit_transfer_params;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ctype, names = decl.split(None, 1)
        fields += [(n.strip(), ctype) for n in names.split(",")]
    want = [(n, {"c_int": "int", "c_float": "float"}[t.__name__]) for n, t in TransferParams._fields_]
    assert fields == want
    assert ctypes.sizeof(TransferParams) == 4 * len(fields)


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/unit_b200.h against unit_b200._lib._SIGNATURES: same argument count and the same
    class of C type (pointer / int / float / 64-bit unsigned) in every position, same return type."""
    from unit_b200 import _lib

    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "unit_b200.h")).read(), flags=re.S)
    protos = re.findall(r"^\s*((?:const\s+)?[\w ]+?\**)\s*(unit_\w+)\s*\(([^;{]*?)\)\s*;", hdr, re.M)
    assert len(protos) >= 30

    def c_kind(decl):
        decl = decl.strip()
        if "*" in decl or decl.startswith("unit_stream_t"):
            return "ptr"
        base = decl.replace("const", "").split()
        base = " ".join(base[:-1]) if len(base) > 1 else base[0]
        return {"int": "int", "float": "float", "size_t": "u64", "unsigned long long": "u64"}[base]

    def ct_kind(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or getattr(t, "_type_", None) is None:
            return "ptr"
        # c_size_t and c_ulonglong are the same 8-byte class on this ABI (ctypes aliases them)
        return {ctypes.c_int: "int", ctypes.c_float: "float", ctypes.c_size_t: "u64",
                ctypes.c_ulonglong: "u64"}.get(t, "ptr")

    seen = set()
    for ret, name, params in protos:
        assert name in _lib._SIGNATURES, name
        seen.add(name)
        restype, argtypes = _lib._SIGNATURES[name]
        plist = [] if params.strip() in ("", "void") else [p for p in params.split(",")]
        assert len(plist) == len(argtypes), (name, len(plist), len(argtypes))
        for i, (decl, t) in enumerate(zip(plist, argtypes)):
            assert c_kind(decl) == ct_kind(t), (name, i, decl.strip(), t)
        want_ret = "ptr" if "*" in ret else {"int": "int", "size_t": "u64", "unsigned long long": "u64"}[ret.strip()]
        assert ct_kind(restype) == want_ret, (name, ret)
    assert seen == set(_lib._SIGNATURES), set(_lib._SIGNATURES) ^ seen


def test_weak_head_image_label_vector():
    """The multi-hot image labels that drive both the MIL loss and the OICR class order: non-zero columns in ascending
    order == torch.unique(gt_classes) of the reference (weak_detector_fast_rcnn.py:203,213-216)."""
    from unit_b200.config import load_cfg
    from unit_b200.predictors import WeakDetectorOutputsBase
    from unit_b200.structures import ShapeSpec

    cfg = load_cfg(os.path.join(ROOT, "configs", "voc_split1_base.yaml"),
                   ["MODEL.ROI_HEADS.EMBEDDING_PATH", os.path.join(ROOT, "tests", "golden", "glove_mean.pt")])
    wd = WeakDetectorOutputsBase(cfg, ShapeSpec(channels=8))
    assert wd.weak_detector_type == "OICR" and wd.oicr_iter == 3 and wd.bg_threshold == 0.1
    targets = [torch.tensor([3, 7, 3, 11]), torch.tensor([], dtype=torch.int64), torch.tensor([19, 0, 8, 8])]
    v = wd.image_label_vector(targets, torch.device("cpu"))
    assert v.shape == (3, 20) and v.dtype == torch.float32
    for row, t in zip(v, targets):
        assert torch.equal(row.nonzero().flatten(), torch.unique(t))
    assert set(v.unique().tolist()) <= {0.0, 1.0}


def test_product_never_imports_oracle():
    for f in glob.glob(os.path.join(ROOT, "unit_b200", "*.py")):
        src = open(f).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_custom_ops_registered_with_fake_and_autograd():
    """north_star: 'a thin C-ABI torch custom-op layer'.  Every op of unit_b200.torch_ops is a torch.library custom op
    with a fake (meta) implementation; the differentiable ones have an autograd formula registered."""
    import torch
    from unit_b200 import ops, torch_ops  # noqa: F401

    for name in torch_ops.OP_NAMES:
        packet = getattr(torch.ops.unit_b200, name)
        assert packet.default._schema.name == f"unit_b200::{name}"
    from torch._subclasses.fake_tensor import FakeTensorMode

    with FakeTensorMode():
        feat = torch.empty(2, 64, 50, 84, device="cuda", requires_grad=True)
        rois = torch.empty(96, 5, device="cuda")
        out = ops.roi_align(feat, rois, 14, 1 / 16, 0, True, rois_sorted=True)
        assert out.shape == (96, 64, 14, 14) and out.device.type == "cuda"
        # an autograd formula is registered: the output carries a grad_fn (running the engine on a CUDA device needs a
        # GPU, so the backward itself is exercised by opcheck on the GPU box: tests/test_ops_gpu.py::test_opcheck)
        assert out.requires_grad and out.grad_fn is not None
        g = ops.roi_align_backward(torch.empty_like(out), rois, feat.shape, 1 / 16, 0, True, True)
        assert g.shape == feat.shape
        x = torch.empty(128, 256, device="cuda")
        w = torch.empty(40, 256, device="cuda", requires_grad=True)
        b = torch.empty(40, device="cuda", requires_grad=True)
        y = ops.linear_tf32(x, w, b)
        assert y.shape == (128, 40) and y.grad_fn is not None
        gw, gb = torch.ops.unit_b200.predictor_wgrad(torch.empty(128, 40, device="cuda"), x)
        assert gw.shape == w.shape and gb.shape == b.shape


def test_heads_trace_under_fake_tensor():
    """The device half of inference (ROIPooler -> packed predictor GEMM -> fused similarity + transfer -> softmax +
    decode -> filter + NMS + top-k) runs under FakeTensorMode on a box without a GPU: only shapes flow, every kernel
    call is a torch.ops.unit_b200 op with a registered fake implementation."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode

    from unit_b200 import d2compat  # noqa: F401
    from unit_b200.config import load_cfg
    from unit_b200.registry import ROI_BOX_HEAD_REGISTRY
    from unit_b200.roi_heads import build_roi_heads
    from unit_b200.structures import Boxes, Instances, ShapeSpec

    class _Feat(torch.nn.Module):
        def __init__(self, cfg, input_shape):
            super().__init__()

        @property
        def output_shape(self):
            return ShapeSpec(channels=256, height=1, width=1)

    if "FakeTraceHead" not in ROI_BOX_HEAD_REGISTRY:
        ROI_BOX_HEAD_REGISTRY._do_register("FakeTraceHead", _Feat)
    cfg = load_cfg(os.path.join(ROOT, "configs", "voc_split1_ft.yaml"),
                   ["MODEL.ROI_BOX_HEAD.NAME", "FakeTraceHead", "MODEL.ROI_HEADS.EMBEDDING_PATH",
                    os.path.join(ROOT, "tests", "golden", "glove_mean.pt")])
    with FakeTensorMode(allow_non_fake_inputs=True), torch.device("cuda"):
        head = build_roi_heads(cfg, {"res4": ShapeSpec(channels=64, stride=16)}).eval()  # parameters: fake CUDA tensors
        emb = head.box_predictor.embeddings  # loaded from disk (a real CPU tensor): stand in a fake CUDA one
        emb.weight = torch.nn.Parameter(torch.empty(tuple(emb.weight.shape), device="cuda"), requires_grad=False)
        feats = torch.empty(2, 64, 50, 84, device="cuda")
        props = [Instances((800, 1333), proposal_boxes=Boxes(torch.empty(300, 4, device="cuda")),
                           objectness_logits=torch.empty(300, device="cuda")) for _ in range(2)]
        with torch.no_grad():
            head.move_mappings_to_gpu()
            pooled = head.box_pooler([feats], [p.proposal_boxes for p in props])
            assert pooled.shape == (600, 64, 14, 14)
            x = torch.empty(600, 256, device="cuda")
            sim = head.get_similarity_matrices(x)
            predictions, _ = head.box_predictor(x, supervised_branch_x_weak=x, novel_classes=head._novel_classes_tensor,
                                                base_classes=head._base_classes_tensor, similarity=sim)
            assert predictions[0].shape == (600, 21) and predictions[1].shape == (600, 80)
            db, ds, dc, dr, cnt = head.box_predictor.inference_device(predictions, props)
            assert db.shape == (2, 100, 4) and ds.shape == (2, 100) and cnt.shape == (2,)
