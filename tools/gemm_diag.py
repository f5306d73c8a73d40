import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from unit_b200 import ops
g=torch.Generator().manual_seed(0)
for (M,N,K) in [(1024,256,2048),(1024,240,2048),(1024,482,2048),(1000,202,2048),(1024,512,2048),(1024,21,2048),(7,21,256),(1024,208,2048),(128,256,32),(128,256,128),(128,128,32),(128,192,32)]:
    x=torch.relu(torch.randn(M,K,generator=g)); w=torch.randn(N,K,generator=g)*0.05; b=torch.randn(N,generator=g)
    ref=(x.double()@w.double().t()+b.double())
    y=ops.predictor_gemm_forward(x.cuda(),w.cuda(),b.cuda()).double().cpu()
    scale=(x.double().abs()@w.double().abs().t()).clamp(min=1e-6)
    e=((y-ref).abs()/scale)
    bad=(e>2e-3)
    cols=bad.any(0).nonzero().flatten(); rows=bad.any(1).nonzero().flatten()
    print((M,N,K),'max',e.max().item(),'bad cols',cols[:6].tolist(),'..',cols[-3:].tolist(),len(cols),'bad rows',rows[:4].tolist(),len(rows))
