"""PCIe ceiling of the end-to-end path: pinned H2D alone, D2H alone, both at once (two streams), sizes of one bench step."""
import torch
F = 100_000_000
h_in = torch.empty((8, F, 2), dtype=torch.float32, pin_memory=True)
h_out = torch.empty((F, 3), dtype=torch.float32, pin_memory=True)
d_in = torch.empty_like(h_in, device="cuda:0")
d_out = torch.empty((F, 3), dtype=torch.float32, device="cuda:0")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    s1.synchronize(); s2.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
for name, fn, gb in (("H2D 6.4 GB", h2d, 6.4), ("D2H 1.2 GB", d2h, 1.2), ("both", lambda: (h2d(), d2h()), 7.6)):
    ms = min(timed(fn) for _ in range(3))
    print("%-12s %.1f ms  %.1f GB/s" % (name, ms, gb / ms * 1e3))
