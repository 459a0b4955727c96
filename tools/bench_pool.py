"""Development aid: attention-pooling forward, direct (hint <= 256) vs shared-memory-staged kernel, graph-timed."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from aimnet_x2d_b200 import _lib  # noqa: E402

lib = _lib.load()
rng = np.random.Generator(np.random.PCG64(0))
B, F, heads = 2112, 512, 4
sizes = rng.integers(9, 30, size=B)
N = int(sizes.sum())
seg = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32), device="cuda")
xs = [torch.randn(N, F, device="cuda") for _ in range(4)]
w = torch.randn(heads, F, device="cuda") / 16
b = torch.zeros(heads, device="cuda")
T = torch.tensor(1.0, device="cuda")
pooled = torch.empty(B, F, device="cuda")
attn = torch.empty(heads, N, device="cuda")
z = torch.empty(heads, N, device="cuda")
P = lambda t: C.c_void_p(t.data_ptr())
for hint in (29, 1000):
    def launch(i):
        _lib.check(lib.ax2d_attn_pool_fwd(P(xs[i % 4]), F, P(seg), B, N, F, heads, P(w), P(b), P(T), P(pooled), P(attn), P(z), hint,
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pool")
    launch(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(16):
            launch(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 16
    print(f"hint {hint:5d}: {us:7.1f} us per launch, {4.0 * N * F / us / 1e3:7.1f} GB/s of x ({N} atoms, {B} molecules)")
