"""Micro-benchmark of the dense projection kernels on the shapes of the C2 training step (not the bench contract:
a development tool; `python tools/bench_gemm.py [--simt] [--iters N] [--shape i]`)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aimnet_x2d_b200 import ops  # noqa: E402

M = 37376
SHAPES = [  # (name, M, K segments, N, epilogue kwargs builder)
    ("mlp 160->160 bias+act+pre+drop", M, [160], 160, "act"),
    ("mlp 160->160 bias+resid", M, [160], 160, "resid"),
    ("in/skip [x|agg] 320->320", M, [160, 160], 320, "io"),
    ("embed-proj 256->544", M, [256], 544, "act"),
    ("concat 544->512", M, [384, 160], 512, "bias"),
    ("head 512->512", 2048, [512], 512, "act"),
    ("dgrad 320->320", M, [160, 160], 320, "plain"),
]


def make_case(m, widths, n, kind, dev="cuda"):
    """Operands and epilogue arguments of one benchmark shape: (a_list, W, c_segs, kwargs)."""
    k = sum(widths)
    a = [torch.randn(m, w, device=dev) for w in widths]
    W = torch.randn(n, k, device=dev) / k ** 0.5
    bias = torch.randn(n, device=dev)
    out = torch.empty(m, n, device=dev)
    pre = torch.empty(m, n, device=dev)
    res = torch.randn(m, n, device=dev)
    kw = {}
    c = [(out, n)]
    if kind == "act":
        kw = dict(bias=bias, pre_segs=[(pre, n)], act="silu", drop_p=0.05, drop_seed=1)
    elif kind == "resid":
        kw = dict(bias=bias, resid=[(res, n)])
    elif kind == "io":
        h = torch.empty(m, n // 2, device=dev); g = torch.empty(m, n // 2, device=dev); z = torch.empty(m, n // 2, device=dev)
        c = [(h, n // 2), (g, n // 2)]
        kw = dict(bias=bias, pre_segs=[(z, n // 2), (None, n // 2)], act="silu", act_cols=n // 2)
    elif kind == "bias":
        kw = dict(bias=bias)
    return a, W, c, kw


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--simt", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--shape", type=int, default=-1)
    ap.add_argument("--plain", action="store_true", help="two eager launches per shape and nothing else (for ncu captures)")
    ap.add_argument("--sweep", action="store_true", help="time vs K (slope = per-k-block cost, intercept = prologue + epilogue)")
    args = ap.parse_args()
    global SHAPES
    if args.sweep:
        SHAPES = [(f"sweep N={n} K={k}", m, [k], n, "plain") for m in (37376, 2048) for n in (160, 256) for k in (16, 160, 320, 640, 1280)]
    ops.USE_TENSOR_CORES = not args.simt
    dev = "cuda"
    torch.manual_seed(0)
    for idx, (name, m, widths, n, kind) in enumerate(SHAPES):
        if args.shape >= 0 and idx != args.shape:
            continue
        k = sum(widths)
        a, W, c, kw = make_case(m, widths, n, kind, dev)
        for _ in range(2 if args.plain else 3):
            ops.gemm(list(zip(a, widths)), [(W, k)], c, m, n, k, **kw)
        torch.cuda.synchronize()
        if args.plain:
            continue
        # pure device time: `reps` launches captured in a CUDA graph, replayed between two events.  Inputs rotate over
        # several buffers so that they do not stay in L2 between launches.
        nbuf = 4 if m > 10000 else 1
        As = [[torch.randn(m, w, device=dev) for w in widths] for _ in range(nbuf)]
        reps = 16
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(reps):
                ops.gemm(list(zip(As[i % nbuf], widths)), [(W, k)], c, m, n, k, **kw)
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps      # includes the tiny split_tf32 launch of the weight (~3 us)
        flops = 2.0 * m * n * k
        bytes_ = 4.0 * (m * k + n * k + m * n)
        print(f"{name:36s} M={m:6d} K={k:4d} N={n:4d}  {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  "
              f"{bytes_ / us / 1e3:7.1f} GB/s (A+W+C only)")


if __name__ == "__main__":
    main()
