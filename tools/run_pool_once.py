"""Development aid for ncu: a few launches of the attention-pooling forward on drug-like molecules (mode, group from argv)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from aimnet_x2d_b200 import _lib  # noqa: E402

lib = _lib.load()
mode, group = int(sys.argv[1]), int(sys.argv[2])
B, lo, hi = (int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (4160, 20, 80)
rng = np.random.Generator(np.random.PCG64(0))
F, heads = 512, 4
P = lambda t: C.c_void_p(t.data_ptr())
sizes = rng.integers(lo, hi, size=B)
N = int(sizes.sum())
seg = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32), device="cuda")
xs = [torch.randn(N, F, device="cuda") for _ in range(2)]
w = torch.randn(heads, F, device="cuda") / 16
b = torch.zeros(heads, device="cuda")
T = torch.tensor(1.0, device="cuda")
pooled = torch.empty(B, F, device="cuda")
attn = torch.empty(heads, N, device="cuda")
z = torch.empty(heads, N, device="cuda")
_lib.check(lib.ax2d_attn_pool_fwd_config(mode, group), "config")
for i in range(3):
    _lib.check(lib.ax2d_attn_pool_fwd(P(xs[i % 2]), F, P(seg), B, N, F, heads, P(w), P(b), P(T), P(pooled), P(attn), P(z), int(sizes.max()),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pool")
torch.cuda.synchronize()
print("ok", N)
