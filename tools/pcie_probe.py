"""Dev probe: PCIe copy rates on the GPU box (pinned host memory): H2D, D2H, both at once, 10 MB and 64 MB copies."""
import torch
dev = torch.device("cuda:0")
def rate(nbytes, reps, direction):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def go():
        if direction in ("h2d", "both"):
            with torch.cuda.stream(s1):
                for _ in range(reps): d.copy_(h, non_blocking=True)
        if direction in ("d2h", "both"):
            with torch.cuda.stream(s2):
                for _ in range(reps): h2.copy_(d2, non_blocking=True)
    go(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s1.wait_event(e0); s2.wait_event(e0)
    go()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
for nbytes, reps in ((10 << 20, 50), (64 << 20, 20)):
    for d in ("h2d", "d2h", "both"):
        print("%3d MB x %d  %-5s %.1f GB/s per direction" % (nbytes >> 20, reps, d, rate(nbytes, reps, d)), flush=True)
