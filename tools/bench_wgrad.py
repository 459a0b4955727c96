"""Development tool: weight-gradient kernel timings on the shapes of the C2 step (graph of 30 launches, rotating buffers).
    python tools/bench_wgrad.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aimnet_x2d_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
M = int(os.environ.get("ROWS", "37632"))
NBUF = 6


def timed(fn, reps=30):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(reps):
            fn(i)
    graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    graph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


for dtype in ((torch.float32,) if os.environ.get('F32_ONLY') else (torch.float32, torch.bfloat16)):
    for (no, ki) in [(160, 160), (160, 128), (160, 96), (152, 160), (128, 160), (160, 640), (160, 320), (512, 512)]:
        rows = 2048 if no == 512 else M
        gs = [torch.randn(rows, no, device="cuda").to(dtype) for _ in range(NBUF)]
        xs = [torch.randn(rows, ki, device="cuda").to(dtype) for _ in range(NBUF)]
        outs = [torch.empty(no, ki, device="cuda") for _ in range(NBUF)]
        obs = [torch.empty(no, device="cuda") for _ in range(NBUF)]
        fn = lambda i: ops._weight_grad([(gs[i % NBUF], no)], [(xs[i % NBUF], ki)], rows, no, ki, "cuda", bias=True,
                                        out=outs[i % NBUF], out_bias=obs[i % NBUF])
        ref = gs[0].double().t() @ xs[0].double()
        dW, db = fn(0)
        err = float((dW.double() - ref).abs().max() / ref.abs().max())
        rb = gs[0].double().sum(0)
        errb = float((db.double() - rb).abs().max() / rb.abs().max())
        print(f"{str(dtype):15s} rows={rows} dW[{no} x {ki}]  {timed(fn):6.1f} us (err {err:.1e}, bias-gradient err {errb:.1e})", flush=True)
