"""Micro-benchmark of the bf16 tensor-core kernels on the shapes of the training step (development tool):
    python tools/bench_gemm_bf16.py [--plain] [--shape i]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aimnet_x2d_b200 import ops  # noqa: E402

M = 37632
BF = torch.bfloat16
SHAPES = [  # (name, M, K segments, N, kind)
    ("mlp 160->160 bias+act+pre+drop", M, [160], 160, "act"),
    ("mlp 160->160 bias+resid", M, [160], 160, "resid"),
    ("mlp 160->160 bias+3 resid", M, [160], 160, "resid3"),
    ("dgrad 160->160 act'*drop", M, [160], 160, "dact"),
    ("in/skip [x|agg] 320->320", M, [160, 160], 320, "io"),
    ("embed-proj 256->544", M, [256], 544, "act"),
    ("concat 544->512", M, [384, 160], 512, "bias"),
    ("head 512->512", 2048, [512], 512, "act"),
    ("wgrad 160x160", M, [160], 160, "wgrad"),
    ("wgrad 320x320", M, [160, 160], 320, "wgrad"),
    ("wgrad 544x256", M, [384, 160], 256, "wgrad"),
    ("wgrad head 512x512", 2048, [512], 512, "wgrad"),
]


def run(name, m, widths, n, kind, plain, dev="cuda"):
    k = sum(widths)
    nbuf = 4 if m > 10000 else 1
    As = [[torch.randn(m, w, device=dev).to(BF) for w in widths] for _ in range(nbuf)]
    W = (torch.randn(n, k, device=dev) / k ** 0.5).to(BF)
    bias = torch.randn(n, device=dev)
    out = torch.empty(m, n, device=dev, dtype=BF)
    pre = torch.empty(m, n, device=dev, dtype=BF)
    res = [torch.randn(m, n, device=dev).to(BF) for _ in range(3)]
    c = [(out, n)]
    kw = {}
    if kind == "act":
        kw = dict(bias=bias, pre_segs=[(pre, n)], act="silu", drop_p=0.05, drop_seed=1)
    elif kind == "resid":
        kw = dict(bias=bias, resid=[(res[0], n)])
    elif kind == "resid3":
        kw = dict(bias=bias, resid=[(r, n) for r in res])
    elif kind == "dact":
        kw = dict(dact_pre=pre, dact="silu", drop_p=0.05, drop_seed=1)
    elif kind == "io":
        h, g, z = (torch.empty(m, n // 2, device=dev, dtype=BF) for _ in range(3))
        c = [(h, n // 2), (g, n // 2)]
        kw = dict(bias=bias, pre_segs=[(z, n // 2), (None, n // 2)], act="silu", act_cols=n // 2)
    elif kind == "bias":
        kw = dict(bias=bias)
    if kind == "wgrad":
        G = [torch.randn(m, n, device=dev).to(BF) for _ in range(nbuf)]
        lib = ops._lib.load()
        ws = torch.empty(lib.ax2d_gemm_bf16_wgrad_workspace(n, k, m) // 4, dtype=torch.float32, device=dev)
        fn = lambda i: ops.weight_grad_bf16([(G[i % nbuf], n)], list(zip(As[i % nbuf], widths)), m, n, k, ws=ws, leave_partials=True)
        nbytes = 2.0 * m * (n + k) + 4.0 * lib.ax2d_gemm_bf16_wgrad_splits(n, k, m) * n * k
    else:
        fn = lambda i: ops.gemm_bf16(list(zip(As[i % nbuf], widths)), W, c, m, n, k, **kw)
        nbytes = 2.0 * (m * k + n * k + m * n * (1 + ("pre_segs" in kw) * (0.5 if kind == "io" else 1) + len(kw.get("resid", ()))
                                                  + ("dact_pre" in kw)))
    for i in range(2):
        fn(i)
    torch.cuda.synchronize()
    if plain:
        return
    reps = 16
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(reps):
            fn(i)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{name:36s} M={m:6d} K={k:4d} N={n:4d}  {us:8.1f} us  {2.0 * m * n * k / us / 1e6:7.1f} TFLOP/s  "
          f"{nbytes / us / 1e3:7.1f} GB/s (all operands)", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, default=-1)
    ap.add_argument("--plain", action="store_true", help="two eager launches per shape and nothing else (for ncu captures)")
    args = ap.parse_args()
    torch.manual_seed(0)
    for idx, (name, m, widths, n, kind) in enumerate(SHAPES):
        if args.shape >= 0 and idx != args.shape:
            continue
        run(name, m, widths, n, kind, args.plain)


if __name__ == "__main__":
    main()
