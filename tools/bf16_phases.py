"""Development tool: %globaltimer stamps of CTA 0 of ax2d_gemm_bf16 on the C2 MLP shape."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aimnet_x2d_b200 import _lib, ops
lib = _lib.load()
f = lib._lib.ax2d_debug_bf16_timing; f.argtypes = [C.c_void_p]; f.restype = None
M, D = 37632, 160
BF = torch.bfloat16
x = torch.randn(M, D, device="cuda").to(BF); W = (torch.randn(D, D, device="cuda") / 12).to(BF); b = torch.randn(D, device="cuda")
out = torch.empty_like(x); pre = torch.empty_like(x); res = torch.randn(M, D, device="cuda").to(BF)
names = ["start", "setup done", "t0 kb0 landed", "t1 kb0 landed", "t0 mma issued", "t1 mma issued", "t0 acc ready", "t0 c0 inputs", "t0 c0 tmem ld",
         "t0 c0 stored", "t1 acc ready", "t1 c0 inputs", "t1 c0 tmem ld", "t1 c0 stored", "last chunk issued", "end"]
for label, kw in (("act+pre+drop", dict(bias=b, pre_segs=[(pre, D)], act="silu", drop_p=0.05, drop_seed=1)), ("bias+resid", dict(bias=b, resid=[(res, D)])),
                  ("plain", dict())):
    buf = torch.zeros(16, dtype=torch.int64, device="cuda")
    for _ in range(3):
        ops.gemm_bf16([(x, D)], W, [(out, D)], M, D, D, **kw)
    f(C.c_void_p(buf.data_ptr()))
    ops.gemm_bf16([(x, D)], W, [(out, D)], M, D, D, **kw)
    torch.cuda.synchronize()
    f(None)
    t = buf.cpu().tolist()
    print(label, " | ".join(f"{n} {(v - t[0]) / 1e3:.2f}" for n, v in zip(names, t) if v))
