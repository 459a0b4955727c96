"""Micro-benchmark of the aggregation kernel at BASELINE configs[1] size (development tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from aimnet_x2d_b200 import ops, synthetic as S  # noqa: E402

dev = "cuda"
TILE_ROWS = int(os.environ.get("TILE_ROWS", "32"))
batch = S.make_batch(1234 + 2000, 2048, 3, "qm9", tile_rows=TILE_ROWS)
gi = batch.graph_index.to(dev)
N, E = gi.num_atoms, gi.num_edges
x = torch.randn(N, 160, device=dev)
x[:, 153:] = 0
g = torch.randn(N, 160, device=dev)
NBUF = 8      # rotate over 8 input / output buffer pairs (8 x 48 MB > 126 MB L2) inside one event pair: no host gaps
xs = [torch.randn(N, 160, device=dev) for _ in range(NBUF)]
gs = [torch.randn(N, 160, device=dev) for _ in range(NBUF)]
for name, fn, nb in (("fwd", lambda i: ops.agg(xs[i % NBUF], gi), ops.agg_bytes(N, N, E, 160)),
                     ("bwd(+addend)", lambda i: ops.agg(gs[i % NBUF], gi, transpose=True, addend=xs[i % NBUF]),
                      ops.agg_bytes(N, N, E, 160, True)),
                     ("fwd same buffer (L2 resident)", lambda i: ops.agg(xs[0], gi), ops.agg_bytes(N, N, E, 160))):
    for i in range(8):
        fn(i)
    torch.cuda.synchronize()
    reps = 64
    graph = torch.cuda.CUDAGraph()          # replayed as a graph: pure GPU time, no Python between launches
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
    print(f"agg {name:30s}: {us:7.1f} us/launch  {nb / us / 1e3:7.1f} GB/s algorithmic ({nb / 1e6:.1f} MB)  "
          f"{100 * nb / us / 1e3 / 6547.2:5.1f}% of measured HBM peak; tiles={gi.n_tiles} N={N} E={E}")

# phase stamps of CTA 0 (development aid)
import ctypes as C
from aimnet_x2d_b200 import _lib
lib = _lib.load()
lib.ax2d_debug_agg_timing.argtypes = [C.c_void_p]
lib.ax2d_debug_agg_timing.restype = None
buf = torch.zeros(16, dtype=torch.int64, device=dev)
lib.ax2d_debug_agg_timing(C.c_void_p(buf.data_ptr()))
for _ in range(3):
    ops.agg(xs[0], gi)
torch.cuda.synchronize()
t = buf.cpu().tolist()
print("CTA0 stamps (us from start):", [round((v - t[0]) / 1e3, 2) if v else None for v in t])
lib.ax2d_debug_agg_timing(None)
