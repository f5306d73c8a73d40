"""HBM write / read / copy rates on this GPU (CUDA events), to put the write-dominated ROIAlign forward (822 MB out,
34 MB in) and the read-dominated backward next to what a pure stream of that kind reaches."""
import json, torch


def ev(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    n = 822_083_584 // 4  # the forward's output bytes
    a = torch.empty(n, dtype=torch.float32, device='cuda')
    b = torch.empty(n, dtype=torch.float32, device='cuda')
    big = torch.empty(4 * n, dtype=torch.float32, device='cuda')
    out = {}
    t = ev(lambda: a.fill_(1.0)); out['fill_822MB_GBs'] = n * 4 / t / 1e6
    t = ev(lambda: a.zero_()); out['memset_822MB_GBs'] = n * 4 / t / 1e6
    t = ev(lambda: big.fill_(2.0)); out['fill_3.3GB_GBs'] = 4 * n * 4 / t / 1e6
    t = ev(lambda: a.sum()); out['read_sum_822MB_GBs'] = n * 4 / t / 1e6
    t = ev(lambda: big.sum()); out['read_sum_3.3GB_GBs'] = 4 * n * 4 / t / 1e6
    t = ev(lambda: b.copy_(a)); out['copy_822MB_rw_GBs'] = 2 * n * 4 / t / 1e6
    print(json.dumps(out))


if __name__ == '__main__':
    main()
