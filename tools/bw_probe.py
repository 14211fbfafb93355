"""HBM bandwidth probes on the GPU box: copy (read+write), write-only and read-only streams.
Used to put the D-TSDF kernel's write-dominated traffic (86 % writes) next to what the memory
system sustains for that mix.  Prints one JSON line."""
import json
import torch

def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3

dev = torch.device("cuda:0")
n = 1 << 30                      # 1 Gi floats = 4 GiB
a = torch.empty(n, dtype=torch.float32, device=dev).normal_()
b = torch.empty_like(a)
res = {}
t = timeit(lambda: b.copy_(a)); res["copy_GBps(read+write)"] = 2 * a.numel() * 4 / t / 1e9
t = timeit(lambda: b.fill_(1.0)); res["fill_GBps(write only)"] = a.numel() * 4 / t / 1e9
t = timeit(lambda: b.zero_()); res["memset_GBps(write only)"] = a.numel() * 4 / t / 1e9
t = timeit(lambda: a.sum()); res["sum_GBps(read only)"] = a.numel() * 4 / t / 1e9
# 14 % read / 86 % write mix like the voxeliser: out[i] = f(in[i // 6.3])  ~ read 1/7 of the bytes written
small = a[: n // 8]
t = timeit(lambda: torch.repeat_interleave(small, 7, output_size=7 * small.numel())); res["expand7_GBps(1 read : 7 write)"] = 8 * small.numel() * 4 / t / 1e9
print(json.dumps(res))
