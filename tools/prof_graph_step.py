"""Where the graph-replayed step spends host time: graph A, count read, host draw, H2D, graphs B + C."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
wl = bench.Workload(dev, torch.float32, 0)
for i in range(12):
    wl.step(i)
torch.cuda.synchronize()
st = list(wl.stage._graphs.values())[0]
T = {k: 0.0 for k in ("graph_a", "counts", "draw", "h2d", "graph_bc", "sync")}
n = 200
for i in range(n):
    t0 = time.perf_counter(); st.graph_a.replay()
    t1 = time.perf_counter(); c = st.lm.counts.cpu().tolist()
    t2 = time.perf_counter(); draw = st._draw(c)
    t3 = time.perf_counter(); st.devbuf.copy_(draw.host, non_blocking=True)
    t4 = time.perf_counter(); st.graph_b.replay(); st.graph_c.replay()
    t5 = time.perf_counter(); torch.cuda.synchronize()
    t6 = time.perf_counter()
    for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)):
        T[k] += v
print({k: round(1e6 * v / n, 1) for k, v in T.items()}, "us per step")
