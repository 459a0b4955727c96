"""Development aid: phase timestamps (globaltimer, ns) of CTA (0,0) of the tensor-core projection kernel."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aimnet_x2d_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
lib.ax2d_debug_timing.argtypes = [C.c_void_p]
lib.ax2d_debug_timing.restype = None
dev = "cuda"
buf = torch.zeros(16, dtype=torch.int64, device=dev)
names = ["start", "setup", "tile0 landed", "tile0 split", "mma issued", "acc ready", "epilogue done", "cta done",
         "ep:begin", "ep:tmem_ld", "ep:sts", "ep:col", "ep:rows01", "ep:chunk0 done"]
for (m, k, n) in [(2048, 16, 160), (2048, 160, 160), (37376, 160, 160), (37376, 320, 320), (2048, 512, 512)]:
    a = torch.randn(m, k, device=dev)
    w = torch.randn(n, k, device=dev)
    out = torch.empty(m, n, device=dev)
    for it in range(3):
        lib.ax2d_debug_timing(C.c_void_p(buf.data_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hi, lo = ops.split_tf32(w)
        torch.cuda.synchronize()
        e0.record()
        old = ops.TC_MIN_ROWS
        ops.TC_MIN_ROWS = 1
        ops.gemm([(a, k)], [(w, k)], [(out, n)], m, n, k)
        ops.TC_MIN_ROWS = old
        e1.record()
        torch.cuda.synchronize()
        lib.ax2d_debug_timing(None)
    t = buf.cpu().tolist()
    print(f"M={m} K={k} N={n}: events {e0.elapsed_time(e1) * 1e3:.1f} us; " +
          ", ".join(f"{nm} +{(t[i] - t[0]) / 1e3:.2f}" for i, nm in enumerate(names)))
