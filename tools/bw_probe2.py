"""Write-bandwidth probes with compressible (constant) and incompressible (random) data."""
import json, torch
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3
dev = torch.device("cuda:0")
n = 1 << 30
res = {}
b = torch.empty(n, dtype=torch.float32, device=dev)
small = torch.randn(n // 64, dtype=torch.float32, device=dev)          # 64 MiB source, stays in L2
ones = torch.ones(n // 64, dtype=torch.float32, device=dev)
bv = b.view(64, n // 64)
t = timeit(lambda: b.fill_(1.0)); res["fill_const"] = n * 4 / t / 1e9
t = timeit(lambda: bv.copy_(ones.view(1, -1).expand(64, -1))); res["bcast_copy_const(write 64x, read 1x)"] = n * 4 / t / 1e9
t = timeit(lambda: bv.copy_(small.view(1, -1).expand(64, -1))); res["bcast_copy_random(write 64x, read 1x)"] = n * 4 / t / 1e9
a = torch.randn(n, dtype=torch.float32, device=dev)
t = timeit(lambda: b.copy_(a)); res["copy_random(read+write bytes)"] = 2 * n * 4 / t / 1e9
a.fill_(1.0)
t = timeit(lambda: b.copy_(a)); res["copy_const(read+write bytes)"] = 2 * n * 4 / t / 1e9
a.normal_()
t = timeit(lambda: a.sum()); res["sum_random(read)"] = n * 4 / t / 1e9
a.fill_(1.0)
t = timeit(lambda: a.sum()); res["sum_const(read)"] = n * 4 / t / 1e9
print(json.dumps(res, indent=1))
