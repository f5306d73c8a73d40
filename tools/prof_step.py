"""Host-side profile of the bench train step (cProfile): where the CPU time between kernel launches goes."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
wl = bench.Workload(dev, torch.float32, 0)
for i in range(5):
    wl.step(i)
torch.cuda.synchronize()
n = 30
t0 = time.perf_counter()
for i in range(n):
    wl.step(i)
torch.cuda.synchronize()
print("wall ms/step", 1000 * (time.perf_counter() - t0) / n)
pr = cProfile.Profile()
pr.enable()
for i in range(n):
    wl.step(i)
torch.cuda.synchronize()
pr.disable()
for key in ("tottime", "cumtime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue()[:9000])
