"""Micro-benchmark of the aggregation kernels at BASELINE configs[1] size (development tool).
    python tools/bench_agg.py            # fp32 + bf16, both kernel generations, CTAs/SM sweep"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aimnet_x2d_b200 import _lib, ops, synthetic as S  # noqa: E402

dev = "cuda"
TILE_ROWS = int(os.environ.get("TILE_ROWS", "32"))
KIND = os.environ.get("KIND", "qm9")
NMOL = int(os.environ.get("NMOL", "2048"))
HOPS = int(os.environ.get("HOPS", "3"))
batch = S.make_batch(1234 + 2000, NMOL, HOPS, KIND, tile_rows=TILE_ROWS)
gi = batch.graph_index.to(dev)
N, E = gi.num_atoms, gi.num_edges
lib = _lib.load()
cfg = lib._lib.ax2d_debug_agg_config
cfg.argtypes = [C.c_int, C.c_int, C.c_int]
cfg.restype = None
NBUF = 8      # rotate over 8 input / output buffer pairs (> 126 MB L2 in fp32) inside one event pair: no host gaps


def run(dtype, label):
    es = 2 if dtype == torch.bfloat16 else 4
    xs = [torch.randn(N, 160, device=dev).to(dtype) for _ in range(NBUF)]
    gs = [torch.randn(N, 160, device=dev).to(dtype) for _ in range(NBUF)]
    res = []
    for name, fn, nb in (("fwd", lambda i: ops.agg(xs[i % NBUF], gi), ops.agg_bytes(N, N, E, 153, False, es)),
                         ("bwd+addend", lambda i: ops.agg(gs[i % NBUF], gi, transpose=True, addend=xs[i % NBUF]),
                          ops.agg_bytes(N, N, E, 153, True, es))):
        for i in range(8):
            fn(i)
        torch.cuda.synchronize()
        reps = 64
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
        us = a.elapsed_time(b) * 1e3 / reps
        res.append(f"{name} {us:6.1f} us {100 * nb / us / 1e3 / 6547.2:5.1f}%")
    print(f"{label:44s} " + " | ".join(res), flush=True)


print(f"N={N} E={E} tiles={gi.n_tiles} tile_rows={TILE_ROWS} kind={KIND}; % = algorithmic bytes (unpadded D=153) / time / 6547 GB/s")
for dtype in (torch.float32, torch.bfloat16):
    cfg(0, 0, 0)
    run(dtype, f"{dtype} gen-1 (8 / 4 lanes per row)")
    for ctas in (2, 3, 4):
        for stages in (0, 2):
            cfg(ctas, stages, 1)
            run(dtype, f"{dtype} warp-per-row ctas/SM={ctas} stages={stages or 'auto'}")
cfg(0, 0, 1)

# phase stamps of CTA 0, thread 0 (development aid): [start, init, (tile ready, tile done) x 6, end]
lib._lib.ax2d_debug_agg_timing.argtypes = [C.c_void_p]
lib._lib.ax2d_debug_agg_timing.restype = None
buf = torch.zeros(16, dtype=torch.int64, device=dev)
x32 = torch.randn(N, 160, device=dev)
for ctas in (2, 3):
    cfg(ctas, 0, 1)
    lib._lib.ax2d_debug_agg_timing(C.c_void_p(buf.data_ptr()))
    for _ in range(3):
        ops.agg(x32, gi)
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    print(f"warp-per-row ctas/SM={ctas}: CTA0 stamps (us from start):", [round((v - t[0]) / 1e3, 2) if v else None for v in t])
    lib._lib.ax2d_debug_agg_timing(None)
    buf.zero_()
cfg(0, 0, 1)
