"""Development aid: per-CTA phase timestamps (globaltimer) and cycle accounting of the tensor-core projection kernel.
The cycle counters need a profile build: AX2D_NVCC_EXTRA=-DAX2D_TC_PROFILE python aimnet_x2d_b200/build.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from aimnet_x2d_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
lib.ax2d_debug_timing.argtypes = [C.c_void_p]
lib.ax2d_debug_timing.restype = None
dev = "cuda"
buf = torch.zeros(16 + 32 * 4096, dtype=torch.int64, device=dev)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_gemm import SHAPES, make_case  # noqa: E402

for name, m, widths, n, kind in SHAPES:
    k = sum(widths)
    a, W, c, kw = make_case(m, widths, n, kind, dev)
    for it in range(3):
        buf.zero_()
        lib.ax2d_debug_timing(C.c_void_p(buf.data_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        ops.gemm(list(zip(a, widths)), [(W, k)], c, m, n, k, **kw)
        e1.record()
        torch.cuda.synchronize()
        lib.ax2d_debug_timing(None)
    knob = name
    t = buf.cpu().numpy()
    per = t[16:16 + 32 * 4096].reshape(4096, 32)
    per = per[per[:, 0] > 0]
    t0 = per[:, 0].min()
    st, en, mm, ar, ep = [(per[:, i] - t0) / 1e3 for i in (0, 1, 3, 4, 5)]
    pct = lambda v: "/".join(f"{np.percentile(v, q):.1f}" for q in (0, 50, 100))
    nkb = (k // 16) * (((m + 127) // 128) * max(1, -(-n // 192)) / len(per))
    mhz = np.median((per[:, 12] - per[:, 11]) / ((per[:, 1] - per[:, 0]) / 1e3))
    print(f"   SM clock during the kernel: {mhz:.0f} MHz (clock64 / globaltimer)")
    med = lambda i: np.median(per[:, i]) / mhz
    print(f"   [{m} {k} {n}] per CTA (us): MMA warp issue {med(6):.1f} + wait stages {med(7):.1f} + wait acc_empty {med(8):.1f} (loop total {med(13):.1f}); epilogue warp wait {med(9):.1f} + work {med(10):.1f}  ({nkb:.0f} k-blocks per CTA)")
    ep = per[:, 16:22].astype(np.float64)
    ep = (ep - ep[:, :1]) / 1e3
    print("   first chunk of the first tile, epilogue warp 0 (us from chunk start): tmem_ld %.2f, staged %.2f, col setup %.2f, rows01 %.2f, chunk done %.2f" % tuple(np.median(ep[:, i]) for i in (1, 2, 3, 4, 5)))
    print(f"knob={knob} ev={e0.elapsed_time(e1)*1e3:.1f} M={m} K={k} N={n}: {len(per)} CTAs on {len(set(per[:, 2].tolist()))} SMs: start {pct(st)}  end {pct(en)}  "
          f"all MMAs issued {pct(mm)}  first accumulator ready {pct(ar)}  last epilogue done {pct(ep)} us (min/50/max)")
