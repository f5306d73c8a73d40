"""Micro-benchmark: ROIAlign fwd/bwd (ours vs torchvision CUDA) at the BASELINE shapes.  CUDA-event timed."""
import sys, os, json
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'tests'))
import torch, torchvision
from unit_b200 import ops
from conftest import random_boxes, seeded

def timeit(fn, iters=20, warm=5, flush=None):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(iters):
        if flush is not None: flush.zero_()
        s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts)//2], ts[0]

def main():
    dt = torch.bfloat16 if "--bf16" in sys.argv else torch.float32
    g=seeded(0)
    n, per = 2, 512
    feat=torch.randn(n,1024,50,84,generator=g).cuda().to(dt)
    rois=torch.cat([torch.cat([torch.full((per,1),float(i)), random_boxes(per,800,1333,g,16.0)],1) for i in range(n)]).cuda()
    flush=torch.empty(256*1024*1024, dtype=torch.uint8, device='cuda')
    es = 4 if dt==torch.float32 else 2
    bytes_alg = n*1024*50*84*es + n*per*20 + n*per*1024*196*es
    out={}
    med,mn=timeit(lambda: ops.roi_align_forward(feat,rois,(14,14),1/16,0,True,True), flush=flush)
    out['ours_fwd_ms']=med; out['ours_fwd_GBs']=bytes_alg/med/1e6
    gout=torch.randn(n*per,1024,14,14,device='cuda').to(dt)
    med,mn=timeit(lambda: ops.roi_align_backward(gout,rois,feat.shape,1/16,0,True,True), flush=flush)
    out['ours_bwd_ms']=med; out['ours_bwd_GBs']=bytes_alg/med/1e6
    if dt==torch.float32:
        med,mn=timeit(lambda: torchvision.ops.roi_align(feat,rois,14,1/16,0,True), flush=flush)
        out['tv_fwd_ms']=med; out['tv_fwd_GBs']=bytes_alg/med/1e6
        med,mn=timeit(lambda: torch.ops.torchvision._roi_align_backward(gout,rois,1/16,14,14,n,1024,50,84,0,True), iters=5, warm=2, flush=flush)
        out['tv_bwd_ms']=med
        a=ops.roi_align_forward(feat,rois,(14,14),1/16,0,True,True); b=torchvision.ops.roi_align(feat,rois,14,1/16,0,True)
        out['max_abs_diff_vs_tv_cuda']=(a-b).abs().max().item()
    print(json.dumps(out))
if __name__=="__main__": main()
