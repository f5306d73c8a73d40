"""Run only the labelling kernels (fused IoU+Matcher, label, sample gather) for ncu, at configs[2] size:
64 images x (1000 proposals + 4 GT)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from unit_b200 import layers
from unit_b200.structures import Boxes, Instances
dev = torch.device("cuda")
g = torch.Generator().manual_seed(64)
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lp, lt = [], []
for _ in range(n_img):
    gtb = bench._boxes(4, 800, 1333, g, 32.0)
    pb = torch.cat([bench._boxes(1000, 800, 1333, g), gtb])
    lp.append(Instances((800, 1333), proposal_boxes=Boxes(pb.to(dev)), objectness_logits=torch.zeros(len(pb), device=dev)))
    lt.append(Instances((800, 1333), gt_boxes=Boxes(gtb.to(dev)), gt_classes=torch.randint(0, 20, (4,), generator=g).to(dev)))
gen = torch.Generator().manual_seed(1)
for _ in range(3):
    layers.label_and_sample(lp, lt, num_classes=20, batch_size_per_image=512, positive_fraction=0.25, thresholds=[0.5],
                            labels=[0, 1], generator=gen)
torch.cuda.synchronize()
