"""H2D rate of pinned buffers by allocation time and allocator (torch pin_memory vs the library's b2s_host_alloc)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from b2slam import _lib

MB = 1 << 20
dst = torch.empty(64 * MB, dtype=torch.uint8, device="cuda")


def lib_pinned(nbytes):
    p = ctypes.c_void_p()
    _lib.check(_lib.lib().b2s_host_alloc(ctypes.byref(p), nbytes))
    arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(nbytes,))
    arr[:] = 1
    return torch.from_numpy(arr)


def rate(t, label):
    n = t.numel()
    best = 0.0
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dst[:n].copy_(t, non_blocking=True); b.record(); torch.cuda.synchronize()
        best = max(best, n / a.elapsed_time(b) / 1e6)
    print("%-46s %6.1f GB/s" % (label, best), flush=True)


early_t = torch.ones(29 * MB, dtype=torch.uint8).pin_memory()
early_l = lib_pinned(29 * MB)
rate(early_t, "torch pin_memory, allocated first")
rate(early_l, "b2s_host_alloc, allocated first")
# the kind of activity bench.py has before its e2e legs: big pinned inputs, device buffers, kernels
big = [torch.ones(70 * MB, dtype=torch.uint8).pin_memory() for _ in range(4)]
junk = [torch.empty(256 * MB, dtype=torch.uint8, device="cuda").zero_() for _ in range(8)]
torch.cuda.synchronize()
for k, t in enumerate(big):
    rate(t[:29 * MB], "torch pin_memory 70 MB buffer #%d" % k)
late_t = torch.ones(29 * MB, dtype=torch.uint8).pin_memory()
late_l = lib_pinned(29 * MB)
rate(late_t, "torch pin_memory, allocated late")
rate(late_l, "b2s_host_alloc, allocated late")
rate(early_t, "torch pin_memory, allocated first (again)")
odd = torch.ones(28800000, dtype=torch.uint8).pin_memory()
rate(odd, "torch pin_memory 28.8 MB (bench size), late")
view = torch.from_numpy(np.ascontiguousarray(np.ones((10000, 2, 360), dtype=np.float32))).pin_memory()
rate(view.view(torch.uint8).reshape(-1), "torch pin_memory of a float32 (10000,2,360), late")
