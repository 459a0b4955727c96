"""Development tool: the packed-add (third generation) aggregation kernel against the earlier per-edge kernels.
Checks bit-identity first, then times fwd / bwd+addend launches (64 per CUDA graph, rotating buffers), the elimination
variants (no edge loop / no stores) and prints the %globaltimer stamps of CTA 0.
    python tools/agg_gen3.py                       # QM9-shaped (BASELINE configs[1])
    KIND=drug HOPS=4 NMOL=1024 python tools/agg_gen3.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aimnet_x2d_b200 import _lib, ops, synthetic as S  # noqa: E402

dev = "cuda"
KIND = os.environ.get("KIND", "qm9")
NMOL = int(os.environ.get("NMOL", "2048"))
HOPS = int(os.environ.get("HOPS", "3"))
WIDTH = int(os.environ.get("WIDTH", "160"))
batch = S.make_batch(1234 + 2000, NMOL, HOPS, KIND)
gi = batch.graph_index.to(dev)
N, E = gi.num_atoms, gi.num_edges
lib = _lib.load()
cfg = lib._lib.ax2d_debug_agg_config
cfg.argtypes = [C.c_int, C.c_int, C.c_int]
cfg.restype = None
tm = lib._lib.ax2d_debug_agg_timing
tm.argtypes = [C.c_void_p]
tm.restype = None
NBUF = 8
print(f"kind={KIND} hops={HOPS} N={N} E={E} tiles={gi.n_tiles} max_tile_rows={gi.max_tile_rows} max_tile_edges={gi.max_tile_edges} width={WIDTH}", flush=True)


def timed(fn, reps=64):
    for i in range(4):
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


for dtype in (torch.float32, torch.bfloat16):
    es = 2 if dtype == torch.bfloat16 else 4
    xs = [torch.randn(N, WIDTH, device=dev).to(dtype) for _ in range(NBUF)]
    gs = [torch.randn(N, WIDTH, device=dev).to(dtype) for _ in range(NBUF)]
    fwd = lambda i: ops.agg(xs[i % NBUF], gi)
    bwd = lambda i: ops.agg(gs[i % NBUF], gi, transpose=True, addend=xs[i % NBUF])
    D = int(0.3 * 512) if WIDTH == 160 else WIDTH
    nb_f, nb_b = ops.agg_bytes(N, N, E, D, False, es), ops.agg_bytes(N, N, E, D, True, es)
    # ---- bit identity against the first-generation kernels (fp32) / warp-per-row kernel (bf16)
    cfg(0, 0, 1)
    ref_f, ref_b = fwd(0).clone(), bwd(0).clone()
    cfg(0, 0, 3)
    new_f, new_b = fwd(0).clone(), bwd(0).clone()
    torch.cuda.synchronize()
    same = torch.equal(ref_f.view(torch.int16 if es == 2 else torch.int32), new_f.view(torch.int16 if es == 2 else torch.int32))
    same_b = torch.equal(ref_b.view(torch.int16 if es == 2 else torch.int32), new_b.view(torch.int16 if es == 2 else torch.int32))
    print(f"{dtype}: gen-3 bit-identical to the default kernels: fwd {same}, bwd+addend {same_b}", flush=True)
    for use, label in ((1, "default (fp32 gen-1 / bf16 warp-per-row)"), (2, "warp-per-row"), (3, "gen-3 packed adds")):
        for ctas in ((0,) if use == 1 else (0, 2)):
            cfg(ctas, 0, use)
            tf, tb = timed(fwd), timed(bwd)
            print(f"  {label:42s} ctas={ctas or 'auto'}: fwd {tf:7.1f} us {100 * nb_f / tf / 1e3 / 6547.2:5.1f}% | bwd+addend {tb:7.1f} us "
                  f"{100 * nb_b / tb / 1e3 / 6547.2:5.1f}% | mean frac {100 * (nb_f + nb_b) / (tf + tb) / 1e3 / 6547.2:5.1f}%", flush=True)
    for flags, label in ((1, "no edge loop"), (2, "no stores"), (3, "neither")):
        cfg(0, flags << 8, 3)
        print(f"  gen-3 {label:14s}: fwd {timed(fwd):7.1f} us | bwd+addend {timed(bwd):7.1f} us", flush=True)
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    cfg(0, 0, 3)
    tm(C.c_void_p(buf.data_ptr()))
    for _ in range(3):
        fwd(0)
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    print("  gen-3 CTA0 stamps [start, consumers start, (tile ready, tile done) x 6, end] us:",
          [round((v - t[0]) / 1e3, 2) if v else None for v in t], flush=True)
    tm(None)
cfg(0, 0, 1)


# per-CTA start / end stamps of one launch in the middle of a back-to-back sequence (development flag 4)
x32 = [torch.randn(N, WIDTH, device=dev) for _ in range(4)]
big = torch.zeros(16 + 2 * 148 * 3 + 16, dtype=torch.int64, device=dev)
cfg(0, 4 << 8, 3)
tm(C.c_void_p(big.data_ptr()))
for i in range(6):
    ops.agg(x32[i % 4], gi)
torch.cuda.synchronize()
t = big.cpu()[16:].view(-1, 2)
t = t[(t[:, 0] > 0) & (t[:, 1] > 0)]
t0 = int(t[:, 0].min())
st, en = (t[:, 0] - t0).float() / 1e3, (t[:, 1] - t0).float() / 1e3
q = lambda v: [round(float(v.quantile(p)), 2) for p in (0.0, 0.1, 0.5, 0.9, 1.0)]
print(f"per-CTA stamps of the last of 6 back-to-back fp32 launches ({len(t)} CTAs): start quantiles {q(st)} us, end quantiles {q(en)} us, "
      f"CTA lifetime quantiles {q(en - st)} us", flush=True)
tm(None)
cfg(0, 0, 1)
if os.environ.get("DUMP"):
    import numpy as np
    np.savez(os.environ["DUMP"], stamps=big.cpu().numpy(), tile_info=gi.tile_info.cpu().numpy())
