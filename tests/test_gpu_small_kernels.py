"""GPU: the small C-ABI kernels around the projections, each against a float64 torch restatement of what it replaces:
one-pass embedding backward (nn.Embedding autograd, gnn.py:262-274), weight packing / gradient collection
(packed.PackedWeights) and the weighted loss (losses.py:14-87)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _lib():
    from aimnet_x2d_b200 import _lib
    return _lib


@pytest.mark.parametrize("N,E,vocabs", [(5000, 64, (119, 9, 7, 7)), (37, 16, (5, 3)), (1, 4, (2,)), (70001, 32, (300, 2, 11))])
def test_embed_bwd_all_matches_index_add(N, E, vocabs):
    L = _lib()
    lib = L.load()
    rng = np.random.Generator(np.random.PCG64(N + E))
    nt = len(vocabs)
    # skewed indices (a few hot rows, like atom types), every table row possible
    idx = [torch.from_numpy(np.minimum(rng.geometric(0.3, size=N) - 1, v - 1).astype(np.int64)).to(DEV) for v in vocabs]
    g = torch.from_numpy(rng.normal(size=(N, nt * E)).astype(np.float32)).to(DEV)
    outs = [torch.full((v, E), float("nan"), device=DEV) for v in vocabs]
    rows = sum(vocabs)
    ws = torch.empty(max(lib.ax2d_embed_bwd_all_workspace(N, rows, E) // 4, 1), dtype=torch.float32, device=DEV)
    ip = (C.c_void_p * nt)(*[i.data_ptr() for i in idx])
    gp = (C.c_void_p * nt)(*[o.data_ptr() for o in outs])
    vp = (C.c_int64 * nt)(*vocabs)
    L.check(lib.ax2d_embed_bwd_all(C.c_void_p(g.data_ptr()), g.stride(0), nt, E, N, ip, vp, gp, C.c_void_p(ws.data_ptr()),
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)), "ax2d_embed_bwd_all")
    torch.cuda.synchronize()
    for t, v in enumerate(vocabs):
        ref = torch.zeros((v, E), dtype=torch.float64, device=DEV).index_add_(0, idx[t], g[:, t * E:(t + 1) * E].double())
        scale = float(ref.abs().max()) + 1e-30
        assert float((outs[t].double() - ref).abs().max()) <= 1e-5 * scale, f"table {t}"
        assert torch.all(outs[t][ref.abs().sum(1) == 0] == 0)          # untouched rows are exact zeros


def test_embed_bwd_all_is_deterministic():
    L = _lib()
    lib = L.load()
    N, E, vocabs = 20000, 64, (119, 9)
    rng = np.random.Generator(np.random.PCG64(3))
    idx = [torch.from_numpy(rng.integers(0, v, size=N).astype(np.int64)).to(DEV) for v in vocabs]
    g = torch.from_numpy(rng.normal(size=(N, 2 * E)).astype(np.float32)).to(DEV)
    res = []
    for _ in range(2):
        outs = [torch.empty((v, E), device=DEV) for v in vocabs]
        ws = torch.empty(lib.ax2d_embed_bwd_all_workspace(N, sum(vocabs), E) // 4, dtype=torch.float32, device=DEV)
        L.check(lib.ax2d_embed_bwd_all(C.c_void_p(g.data_ptr()), g.stride(0), 2, E, N,
                                       (C.c_void_p * 2)(*[i.data_ptr() for i in idx]), (C.c_int64 * 2)(*vocabs),
                                       (C.c_void_p * 2)(*[o.data_ptr() for o in outs]), C.c_void_p(ws.data_ptr()),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "ax2d_embed_bwd_all")
        res.append([o.cpu().numpy() for o in outs])
    assert all(np.array_equal(a, b) for a, b in zip(*res))


def test_pack_weights_and_unpack_grads():
    """Blocks of two parameters into one padded matrix: packed fp32 copy, TF32 terms (hi + lo == w up to 2^-22, hi has
    13 zero mantissa bits), transposed terms, zero padding untouched; gradients scattered back additively."""
    from aimnet_x2d_b200.packed import PackedWeights
    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.randn(37, 50, device=DEV))
    b = torch.nn.Parameter(torch.randn(21, device=DEV))
    pk = PackedWeights(DEV)
    pk.add("W", 64, 96, [(a, 0, 37, 0, 20, 0, 0), (a, 5, 37, 20, 50, 0, 32)])     # two column chunks, padded apart
    pk.add("b", 1, 32, [(b, 0, 1, 0, 21, 0, 0)], vector=True)
    pk.finalize()
    pk.refresh()
    W, info = pk["W"], pk["W"]._ax2d
    want = torch.zeros(64, 96, device=DEV)
    want[:37, :20] = a.data[:, :20]
    want[:32, 32:62] = a.data[5:37, 20:50]
    assert torch.equal(W, want)
    assert torch.equal(pk["b"][:21], b.data) and torch.all(pk["b"][21:] == 0)
    assert torch.all((info.hi.view(torch.int32) & 0x1FFF) == 0) and torch.all((info.lo.view(torch.int32) & 0x1FFF) == 0)
    assert float((info.hi.double() + info.lo.double() - W.double()).abs().max()) <= 2.0 ** -22 * float(W.abs().max())
    assert torch.equal(info.hiT, info.hi.t()) and torch.equal(info.loT, info.lo.t())
    # gradients: written into the packed buffers by the kernels, collected additively into .grad
    a.grad = torch.ones_like(a)
    info.grad.copy_(torch.arange(64 * 96, device=DEV, dtype=torch.float32).view(64, 96))
    pk["b"]._ax2d.grad.fill_(2.0)
    pk.dirty = True
    pk.unpack_grads()
    ga = torch.ones(37, 50, device=DEV)
    full = torch.arange(64 * 96, device=DEV, dtype=torch.float32).view(64, 96)
    ga[:, :20] += full[:37, :20]
    ga[5:37, 20:50] += full[:32, 32:62]
    assert torch.equal(a.grad, ga)
    assert torch.equal(b.grad, torch.full((21,), 2.0, device=DEV))
    assert not pk.dirty and torch.all(info.grad == 0)              # cleared for the next step
    # a parameter update shows up after the next refresh
    with torch.no_grad():
        a.mul_(0.5)
    pk.refresh()
    assert torch.equal(pk["W"][:37, :20], a.data[:, :20])


@pytest.mark.parametrize("kind", ["l1", "mse"])
@pytest.mark.parametrize("B,T", [(2112, 12), (7, 1), (1, 3)])
def test_weighted_loss_kernel(kind, B, T):
    import aimnet_x2d_b200 as ax
    rng = np.random.Generator(np.random.PCG64(B * 31 + T))
    pred = torch.from_numpy(rng.normal(size=(B, T)).astype(np.float32)).to(DEV).requires_grad_(True)
    tgt = torch.from_numpy(rng.normal(size=(B, T)).astype(np.float32)).to(DEV)
    w = torch.from_numpy(rng.uniform(0.5, 2.0, size=T).astype(np.float32))
    crit = (ax.WeightedL1Loss if kind == "l1" else ax.WeightedMSELoss)(w).to(DEV)
    loss = crit(pred, tgt)
    loss.backward()
    p64 = pred.detach().double().requires_grad_(True)
    d = p64 - tgt.double()
    ref = ((d.abs() if kind == "l1" else d * d) * w.double().to(DEV)).sum(1).mean()      # losses.py:36-43, 71-78
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= 1e-5 * abs(float(ref.detach()))
    assert float((pred.grad.double() - p64.grad).abs().max()) <= 1e-5 * float(p64.grad.abs().max())


def _attn_pool_reference(x, w, b, T, sizes):
    heads, F, N = w.shape[0], x.shape[1], x.shape[0]
    zz = (x.double() @ w.double().t() + b.double()) / T                                       # [N, heads]
    ref_pooled = torch.zeros((len(sizes), F), dtype=torch.float64, device=DEV)
    ref_attn = torch.zeros((heads, N), dtype=torch.float64, device=DEV)
    o = 0
    for g, n in enumerate(sizes):
        if n:
            a = torch.softmax(zz[o:o + n], dim=0)                                             # per head over the molecule
            ref_attn[:, o:o + n] = a.t()
            ref_pooled[g] = (a.t().unsqueeze(2) * x[o:o + n].double().unsqueeze(0)).sum(1).mean(0)
        o += n
    return zz, ref_attn, ref_pooled


@pytest.mark.parametrize("mode,group", [(1, 0), (0, 0), (0, 1), (0, 2), (0, 4), (2, 0), (2, 1), (2, 4)])
@pytest.mark.parametrize("hint", [8, 300, 1000])
def test_attn_pool_fwd_direct_and_staged_kernels(hint, mode, group):
    """mode 1 (two-phase kernels): hint <= 256 selects the kernel that reads x straight from global memory; a molecule
    larger than the hint (300 atoms against hint 8) must still be handled (scores through global memory); hint 1000
    selects the staged kernel.  mode 0 / 2: the single-pass streaming kernel (online softmax; head weights in shared memory /
    in registers) with 1 / 2 / 4 warps per molecule (empty molecules, a single atom, fewer rows than warps, 300 rows).  Reference: pooling.py:134-161 in float64."""
    L = _lib()
    lib = L.load()
    rng = np.random.Generator(np.random.PCG64(hint))
    sizes = [300, 5, 0, 40, 1, 29, 2, 3, 0]
    N, F, heads = sum(sizes), 64, 4
    seg = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32), device=DEV)
    x = torch.from_numpy(rng.normal(size=(N, F)).astype(np.float32)).to(DEV)
    w = torch.from_numpy(rng.normal(size=(heads, F)).astype(np.float32) / 8).to(DEV)
    b = torch.from_numpy(rng.normal(size=heads).astype(np.float32)).to(DEV)
    T = torch.tensor(0.7, device=DEV)
    pooled = torch.full((len(sizes), F), float("nan"), device=DEV)
    attn = torch.full((heads, N), float("nan"), device=DEV)
    z = torch.full((heads, N), float("nan"), device=DEV)
    P = lambda t: C.c_void_p(t.data_ptr())
    L.check(lib.ax2d_attn_pool_fwd_config(mode, group), "ax2d_attn_pool_fwd_config")
    try:
        L.check(lib.ax2d_attn_pool_fwd(P(x), F, P(seg), len(sizes), N, F, heads, P(w), P(b), P(T), P(pooled), P(attn), P(z), hint,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "ax2d_attn_pool_fwd")
        torch.cuda.synchronize()
    finally:
        lib.ax2d_attn_pool_fwd_config(0, 0)
    zz, ref_attn, ref_pooled = _attn_pool_reference(x, w, b, 0.7, sizes)
    assert float((attn.double() - ref_attn).abs().max()) <= 1e-6
    assert float((pooled.double() - ref_pooled).abs().max()) <= 1e-5 * float(ref_pooled.abs().max())
    assert float((z.double() - zz.t()).abs().max()) <= 1e-5 * float(zz.abs().max())


@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("F,heads,group", [(512, 4, 0), (512, 4, 1), (512, 4, 4), (160, 4, 2), (256, 8, 4), (512, 8, 0), (32, 1, 2),
                                           (100, 2, 4), (640, 4, 0)])
def test_attn_pool_fwd_streaming_shapes(F, heads, group, mode):
    """The streaming forward over the row widths and head counts it is compiled for (1 / 2 / 4 float4 per lane; per-head
    accumulators in registers), and the fall-back to the two-phase kernels beyond them (8 heads x 512, F = 640);
    scores with a large spread (sharp softmax, maxima that keep moving), against float64."""
    L = _lib()
    lib = L.load()
    rng = np.random.Generator(np.random.PCG64(F + heads))
    sizes = [int(v) for v in rng.integers(0, 70, size=150)] + [1, 0, 257]
    N = sum(sizes)
    seg = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32), device=DEV)
    x = torch.from_numpy(rng.normal(size=(N, F)).astype(np.float32)).to(DEV)
    x[::7] *= 3.0                                                                             # scores spread over +-20
    w = torch.from_numpy((rng.normal(size=(heads, F)) / np.sqrt(F)).astype(np.float32)).to(DEV)
    b = torch.from_numpy(rng.normal(size=heads).astype(np.float32)).to(DEV)
    T = torch.tensor(0.3, device=DEV)
    pooled = torch.full((len(sizes), F), float("nan"), device=DEV)
    attn = torch.full((heads, N), float("nan"), device=DEV)
    z = torch.full((heads, N), float("nan"), device=DEV)
    P = lambda t: C.c_void_p(t.data_ptr())
    L.check(lib.ax2d_attn_pool_fwd_config(mode, group), "ax2d_attn_pool_fwd_config")
    try:
        L.check(lib.ax2d_attn_pool_fwd(P(x), F, P(seg), len(sizes), N, F, heads, P(w), P(b), P(T), P(pooled), P(attn), P(z), 257,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "ax2d_attn_pool_fwd")
        torch.cuda.synchronize()
    finally:
        lib.ax2d_attn_pool_fwd_config(0, 0)
    zz, ref_attn, ref_pooled = _attn_pool_reference(x, w, b, 0.3, sizes)
    assert float((attn.double() - ref_attn).abs().max()) <= 2e-5      # scores of +-30: their own fp32 rounding
    assert float((pooled.double() - ref_pooled).abs().max()) <= 1e-5 * float(ref_pooled.abs().max())
    assert float((z.double() - zz.t()).abs().max()) <= 1e-5 * float(zz.abs().max())


@pytest.mark.parametrize("kind,hops,count", [("qm9", 3, 300), ("drug", 4, 64), ("qm9", 1, 5)])
def test_shell_csr_device_equals_host(kind, hops, count):
    """f-1 on the device: the batched bitset BFS (one warp per molecule) writes the same rowptr / col / col_t as
    ax2d_host_shell_csr, bit for bit -- including single atoms, molecules without bonds, duplicate and self bonds."""
    from aimnet_x2d_b200 import collate as CL, molgen
    rng = np.random.Generator(np.random.PCG64(91 + hops))
    mols = [molgen.make_molecule(rng, kind, 4, False) for _ in range(count)]
    mols.insert(3, dict(num_atoms=1, bonds=np.zeros((0, 2), np.int32)))
    mols.insert(7, dict(num_atoms=6, bonds=np.zeros((0, 2), np.int32)))
    mols.append(dict(num_atoms=4, bonds=np.array([[0, 1], [1, 0], [2, 2], [1, 3]], np.int32)))
    counts, bonds = [m["num_atoms"] for m in mols], [m["bonds"] for m in mols]
    h_rowptr, h_col, h_col_t = CL.shell_csr(counts, bonds, hops)
    d_rowptr, d_col, d_col_t = CL.shell_csr(counts, bonds, hops, device=DEV)
    assert d_rowptr.is_cuda and d_col.is_cuda
    N = sum(counts)
    E = int(h_rowptr[N])
    assert int(d_rowptr[N]) == E and E > 0
    assert torch.equal(d_rowptr.cpu()[: N + 1], h_rowptr[: N + 1])
    assert torch.equal(d_col.cpu()[:E], h_col[:E])
    assert torch.equal(d_col_t.cpu()[:E], h_col_t[:E])
    a = CL.GraphIndex.from_bonds(counts, bonds, hops)
    b = CL.GraphIndex.from_bonds(counts, bonds, hops, device=DEV)
    for k in CL.GraphIndex._TENSORS:
        assert torch.equal(getattr(a, k), getattr(b, k)), k


def test_shell_csr_device_full_batch_feeds_the_aggregation():
    """A whole C2-sized batch: CSR built on the device from the bond lists == the collated batch's GraphIndex, and the
    aggregation over it equals the aggregation over the collated index."""
    from aimnet_x2d_b200 import collate as CL, ops, synthetic as S
    mols = S.make_molecules(77, 2048, 3, "qm9", 4)
    batch = CL.MolBatch.from_data_list([S.to_data(m) for m in mols], S.FEATURE_SIZES)
    gi = CL.GraphIndex.from_bonds([m["num_atoms"] for m in mols], [m["bonds"] for m in mols], 3, device=DEV)
    ref = batch.graph_index
    for k in CL.GraphIndex._TENSORS:
        assert torch.equal(getattr(gi, k), getattr(ref, k)), k
    x = torch.randn(gi.num_atoms, 160, device=DEV)
    assert torch.equal(ops.agg(x, gi.to(DEV)), ops.agg(x, ref.to(DEV)))


# ------------------------------------------------------------------------------------------------ a5: stereo kernels
def _stereo_index(N, tetra, cis, trans):
    from aimnet_x2d_b200.collate import GraphIndex
    gi = GraphIndex()
    gi.num_atoms = N
    gi.set_stereo(tetra, cis, trans)
    if gi.tetra is not None:
        i, sp, si, M = gi.tetra
        gi.tetra = (i.to(DEV), sp.to(DEV), si.to(DEV), M)
    if gi.cistrans is not None:
        s, t, sg, n = gi.cistrans
        gi.cistrans = (s.to(DEV), t.to(DEV), sg.to(DEV), n)
    return gi


@pytest.mark.parametrize("N,D,Dp,M,seed", [(64, 12, 12, 9, 0), (200, 30, 32, 70, 1), (5, 4, 4, 1, 2), (900, 153, 160, 400, 3)])
def test_tetra_kernels_match_oracle(N, D, Dp, M, seed):
    """Stand-alone ax2d_tetra_fwd / ax2d_tetra_bwd against the oracle's restatement of gnn.py:387-462 in float64:
    atoms shared by several centres (index_add_ order), repeated atoms inside one centre, atoms in no centre (zeroed,
    quirk Q4), padded feature columns; forward values and the gradient with respect to x at 1e-5 of the largest entry."""
    from aimnet_x2d_b200 import ops
    from oracle import model_port
    rng = np.random.Generator(np.random.PCG64(seed))
    tetra = rng.integers(0, max(N // 2, 4), size=(M, 4)).astype(np.int64)           # upper half of the atoms: never used
    if M > 2:
        tetra[1, 1] = tetra[1, 0]                                                       # a repeated neighbour
        tetra[2] = tetra[0]                                                             # two centres on the same atoms
    x = np.zeros((N, Dp), dtype=np.float32)
    x[:, :D] = rng.normal(size=(N, D)).astype(np.float32)
    x[3 % N, :D] *= 0.05                                                                # a short row (its gradient is amplified by 1 / norm)
    g = rng.normal(size=(N, Dp)).astype(np.float32)
    gi = _stereo_index(N, tetra, None, None)
    xd = torch.from_numpy(x).to(DEV).requires_grad_(True)
    out = ops.TetraFn.apply(xd, D, gi)
    out.backward(torch.from_numpy(g).to(DEV))
    xr = torch.from_numpy(x[:, :D]).double().requires_grad_(True)
    ref = model_port.tetrahedral(xr, torch.from_numpy(tetra))
    ref.backward(torch.from_numpy(g[:, :D]).double())
    o, gx = out.detach().cpu().double(), xd.grad.cpu().double()
    assert float((o[:, :D] - ref.detach()).abs().max()) <= 1e-5 * float(ref.detach().abs().max())
    assert float((gx[:, :D] - xr.grad).abs().max()) <= 1e-5 * float(xr.grad.abs().max())
    untouched = np.setdiff1d(np.arange(N), np.unique(tetra))
    assert torch.all(o[untouched, :D] == 0)                                             # quirk Q4: exact zeros
    # fp32 oracle (the reference's own arithmetic): same bar
    ref32 = model_port.tetrahedral(torch.from_numpy(x[:, :D]), torch.from_numpy(tetra))
    assert float((o[:, :D] - ref32.double()).abs().max()) <= 1e-5 * float(ref32.abs().max())


@pytest.mark.parametrize("N,D,Dp,Kc,Kt,seed", [(50, 12, 12, 3, 2, 0), (300, 153, 160, 40, 25, 1), (10, 8, 8, 0, 2, 2), (10, 8, 8, 1, 0, 3)])
def test_cistrans_kernel_matches_oracle(N, D, Dp, Kc, Kt, seed):
    """Stand-alone ax2d_cistrans (forward and its transposed backward) against the oracle's gnn.py:465-509 restatement:
    only rows 0 and 1 of the [2K, 2] index tensors are used (quirk Q3), cis subtracts, trans adds, repeated targets
    accumulate."""
    from aimnet_x2d_b200 import ops
    from oracle import model_port
    rng = np.random.Generator(np.random.PCG64(100 + seed))
    cis = rng.integers(0, N, size=(2 * Kc, 2)).astype(np.int64) if Kc else np.zeros((0, 2), dtype=np.int64)
    trans = rng.integers(0, N, size=(2 * Kt, 2)).astype(np.int64) if Kt else np.zeros((0, 2), dtype=np.int64)
    if Kc and Kt:
        trans[1] = cis[1]                                                               # the same targets from both lists
    x = np.zeros((N, Dp), dtype=np.float32)
    x[:, :D] = rng.normal(size=(N, D)).astype(np.float32)
    g = rng.normal(size=(N, Dp)).astype(np.float32)
    gi = _stereo_index(N, None, cis, trans)
    xd = torch.from_numpy(x).to(DEV).requires_grad_(True)
    out = ops.CisTransFn.apply(xd, gi)
    out.backward(torch.from_numpy(g).to(DEV))
    xr = torch.from_numpy(x[:, :D]).double().requires_grad_(True)
    ref = model_port.cis_trans(xr, torch.from_numpy(cis), torch.from_numpy(trans))
    ref.backward(torch.from_numpy(g[:, :D]).double())
    o, gx = out.detach().cpu().double(), xd.grad.cpu().double()
    assert float((o[:, :D] - ref.detach()).abs().max()) <= 1e-5 * float(ref.detach().abs().max())
    assert float((gx[:, :D] - xr.grad).abs().max()) <= 1e-5 * float(xr.grad.abs().max())
