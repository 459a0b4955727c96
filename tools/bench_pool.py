"""Development aid: attention-pooling forward, graph-timed: the two-phase kernels (direct for hint <= 256, shared-memory
staged otherwise) against the single-pass streaming kernel with 1 / 2 / 4 warps per molecule, on QM9-sized and
drug-like-sized molecules (C2 / c3 / c4 shapes)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from aimnet_x2d_b200 import _lib  # noqa: E402

lib = _lib.load()
rng = np.random.Generator(np.random.PCG64(0))
F, heads = 512, 4
P = lambda t: C.c_void_p(t.data_ptr())
for label, B, lo, hi in (("qm9 x 2112", 2112, 9, 30), ("drug x 1088", 1088, 20, 80), ("drug x 4160", 4160, 20, 80)):
    sizes = rng.integers(lo, hi, size=B)
    N = int(sizes.sum())
    seg = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32), device="cuda")
    nx = max(2, int(300e6 // (4 * N * F)) + 1)               # rotating inputs larger than L2
    xs = [torch.randn(N, F, device="cuda") for _ in range(nx)]
    w = torch.randn(heads, F, device="cuda") / 16
    b = torch.zeros(heads, device="cuda")
    T = torch.tensor(1.0, device="cuda")
    pooled = torch.empty(B, F, device="cuda")
    attn = torch.empty(heads, N, device="cuda")
    z = torch.empty(heads, N, device="cuda")
    ref = None
    for name, mode, group, hint in (("two-phase direct", 1, 0, int(sizes.max())), ("two-phase staged", 1, 0, 1000), ("stream auto", 0, 0, 0),
                                    ("stream G=1", 0, 1, 0), ("stream G=2", 0, 2, 0), ("stream G=4", 0, 4, 0),
                                    ("stream reg-w G=1", 2, 1, 0), ("stream reg-w G=4", 2, 4, 0)):
        _lib.check(lib.ax2d_attn_pool_fwd_config(mode, group), "config")

        def launch(i):
            _lib.check(lib.ax2d_attn_pool_fwd(P(xs[i % nx]), F, P(seg), B, N, F, heads, P(w), P(b), P(T), P(pooled), P(attn), P(z), hint,
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pool")
        for t in (pooled, attn, z):
            t.fill_(float("nan"))                    # nothing may survive from the previous configuration
        launch(0)
        torch.cuda.synchronize()
        if ref is None:
            ref = (pooled.clone(), attn.clone())
        err = (float((pooled - ref[0]).abs().max() / ref[0].abs().max()), float((attn - ref[1]).abs().max()))
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
        print(f"{label:12s} {name:18s}: {us:7.1f} us per launch, {4.0 * N * F / us / 1e3:7.1f} GB/s of x ({N} atoms); "
              f"vs two-phase: pooled {err[0]:.1e} attn {err[1]:.1e}", flush=True)
    lib.ax2d_attn_pool_fwd_config(0, 0)
