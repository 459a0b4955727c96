import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aimnet_x2d_b200 import _lib, ops, synthetic as S
dev = "cuda"
batch = S.make_batch(1234 + 2000, 2048, 3, "qm9"); gi = batch.graph_index.to(dev); N = gi.num_atoms
lib = _lib.load()
tm = lib._lib.ax2d_debug_agg_timing; tm.argtypes = [C.c_void_p]; tm.restype = None
def bench(fn, label):
    for i in range(4): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(64): fn(i)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    print(f"{label:40s} {a.elapsed_time(b) * 1e3 / 64:6.1f} us", flush=True)
for dt in (torch.float32, torch.bfloat16):
    xs = [torch.randn(N, 160, device=dev).to(dt) for _ in range(8)]
    for tc in (False, True):
        ops.AGG_TENSOR_CORES = tc
        bench(lambda i: ops.agg(xs[i % 8], gi), f"{dt} tc={tc} fwd")
        bench(lambda i: ops.agg(xs[i % 8], gi, transpose=True, addend=xs[(i + 1) % 8]), f"{dt} tc={tc} bwd+addend")
    ops.AGG_TENSOR_CORES = True
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    tm(C.c_void_p(buf.data_ptr()))
    for _ in range(3): ops.agg(xs[0], gi)
    torch.cuda.synchronize(); tm(None)
    t = buf.cpu().tolist()
    print(dt, "stamps [start, (ready, packed, scattered, products done) x3, end]:", [round((v - t[0]) / 1e3, 2) if v else None for v in t])
