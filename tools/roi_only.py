"""Run only the ROIAlign forward kernel a few times (for ncu)."""
import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import torch
from unit_b200 import ops
from conftest import random_boxes, seeded
g=seeded(0); n,per=2,512
feat=torch.randn(n,1024,50,84,generator=g).cuda()
rois=torch.cat([torch.cat([torch.full((per,1),float(i)), random_boxes(per,800,1333,g,16.0)],1) for i in range(n)]).cuda()
mode=sys.argv[1] if len(sys.argv)>1 else 'fwd'
for _ in range(3):
    if mode=='fwd':
        o=ops.roi_align_forward(feat,rois,(14,14),1/16,0,True,True)
    else:
        gout=torch.randn(n*per,1024,14,14,device='cuda')
        o=ops.roi_align_backward(gout,rois,feat.shape,1/16,0,True,True)
torch.cuda.synchronize()
