"""GPU: the tensor-core projection kernel (ax2d_gemm_tc: tcgen05 + TMEM + TMA, 3xTF32 operand split) against a
float64 torch reference of the same contraction, and against the exact-fp32 SIMT kernel for the fused epilogue
(same counter-based dropout decisions in both).  Bar: 1e-5 relative to the tensor scale (fp32 configuration)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from aimnet_x2d_b200 import ops
    return ops


def _segs(rng, M, widths, scale=1.0):
    return [torch.from_numpy((rng.normal(0, scale, size=(M, w))).astype(np.float32)).to(DEV) for w in widths]


def _rel(a, b):
    b = b.double()
    return float((a.double() - b).abs().max() / b.abs().max())


SHAPES = [
    (1000, [160, 160], 320),     # ShellConv input/skip projection (stacked), ragged M
    (300, [384, 160], 512),      # concat_self_other
    (257, [256], 544),           # embedding projection: N > 512 -> three column tiles, last one ragged
    (4096, [160], 160),          # MLP block
    (2048, [512], 512),          # head
    (513, [64, 32], 36),         # N not a multiple of 32
    (128, [32], 32),             # single k-block, single tile
    (2112, [512, 512], 12),      # output layer: N far below the 32-column tile (TMA box larger than the weight)
]


@pytest.mark.parametrize("M,widths,N", SHAPES)
def test_gemm_tc_matches_float64(M, widths, N):
    ops = _ops()
    rng = np.random.Generator(np.random.PCG64(M + N))
    K = sum(widths)
    a = _segs(rng, M, widths)
    W = torch.from_numpy(rng.normal(0, 1.0 / np.sqrt(K), size=(N, K)).astype(np.float32)).to(DEV)
    bias = torch.from_numpy(rng.normal(0, 0.1, size=N).astype(np.float32)).to(DEV)
    out = torch.full((M, N), float("nan"), device=DEV)
    assert ops.USE_TENSOR_CORES and M >= 128
    old = ops.TC_MIN_ROWS
    ops.TC_MIN_ROWS = 1
    try:
        t = ops.KernelTimer()
        with t:
            ops.gemm(list(zip(a, widths)), [(W, K)], [(out, N)], M, N, K, bias=bias)
        assert "gemm_tc" in t.events, "the tensor-core kernel was not selected"
    finally:
        ops.TC_MIN_ROWS = old
    ref = torch.cat(a, 1).double() @ W.double().t() + bias.double()
    err = _rel(out, ref)
    assert err <= 1e-5, f"relative error {err:.3e}"
    # and the data-gradient form dX = dY W (weight transposed by the split kernel)
    dy = _segs(rng, M, [N])[0] if N % 32 == 0 else None
    if dy is not None and K % 4 == 0:
        dx = torch.full((M, K), float("nan"), device=DEV)
        ops.TC_MIN_ROWS = 1
        try:
            ops.gemm([(dy, N)], [(W, K)], [(dx, K)], M, K, N, trans_b=False)
        finally:
            ops.TC_MIN_ROWS = old
        err = _rel(dx, dy.double() @ W.double())
        assert err <= 1e-5, f"dgrad relative error {err:.3e}"


def test_gemm_tc_large_dynamic_range():
    """Operands with a large common offset (|x| ~ 100, like the atom embeddings of the deterministic fixtures)."""
    ops = _ops()
    rng = np.random.Generator(np.random.PCG64(5))
    M, K, N = 2000, 512, 256
    a = torch.from_numpy((100.0 + rng.normal(0, 1, size=(M, K))).astype(np.float32)).to(DEV)
    W = torch.from_numpy(rng.normal(0, 0.05, size=(N, K)).astype(np.float32)).to(DEV)
    out = torch.empty((M, N), device=DEV)
    ops.gemm([(a, K)], [(W, K)], [(out, N)], M, N, K)
    ref = a.double() @ W.double().t()
    # error relative to sum |a||w| (the conditioning of the sum): the split drops only the rounding of the two lo
    # terms (2 x 2^-23); what remains is the tensor core's truncating fp32 accumulation over K / 8 = 64 steps
    cond = (a.double().abs() @ W.double().abs().t()).max()
    err = float((out.double() - ref).abs().max() / cond)
    assert err <= 1e-6, err


@pytest.mark.parametrize("act", ["silu", "gelu", "relu"])
def test_gemm_tc_fused_epilogue_equals_simt(act):
    """Same fused epilogue in both kernels: bias, pre-activation copy, activation on a column prefix, hash dropout,
    residuals, segmented outputs; then the backward-style epilogue (act' * dropout)."""
    ops = _ops()
    rng = np.random.Generator(np.random.PCG64(11))
    M, widths, N = 3000, [160, 160], 320
    K = sum(widths)
    a = _segs(rng, M, widths)
    W = torch.from_numpy(rng.normal(0, 1.0 / np.sqrt(K), size=(N, K)).astype(np.float32)).to(DEV)
    bias = torch.from_numpy(rng.normal(0, 0.1, size=N).astype(np.float32)).to(DEV)
    r1 = _segs(rng, M, [N])[0]
    r2 = _segs(rng, M, [160])[0]

    def run(tc):
        ops.USE_TENSOR_CORES = tc
        h = torch.empty((M, 160), device=DEV); gsk = torch.empty((M, 160), device=DEV)
        z0 = torch.empty((M, 160), device=DEV)
        ops.gemm(list(zip(a, widths)), [(W, K)], [(h, 160), (gsk, 160)], M, N, K, bias=bias,
                 pre_segs=[(z0, 160), (None, 160)], act=act, act_cols=160, drop_p=0.3, drop_seed=1234,
                 resid=[(r1, N), (r2, 160)])
        d = torch.empty((M, N), device=DEV)
        ops.gemm(list(zip(a, widths)), [(W, K)], [(d, N)], M, N, K, dact_pre=r1, dact=act, drop_p=0.3, drop_seed=99,
                 resid=[(r1, N)])
        return h, gsk, z0, d

    try:
        tc = run(True)
        simt = run(False)
    finally:
        ops.USE_TENSOR_CORES = True
    for x, y, name in zip(tc, simt, ("h", "gskip", "pre", "dact")):
        assert _rel(x, y) <= 1e-5, name
    # dropout really happened and the decisions agree: zeros of the non-residual part coincide
    assert float((tc[0] == simt[0]).float().mean()) > 0.0


def test_split_tf32_is_exact():
    ops = _ops()
    rng = np.random.Generator(np.random.PCG64(3))
    w = torch.from_numpy(rng.normal(0, 1, size=(70, 45)).astype(np.float32)).to(DEV)
    hi, lo = ops.split_tf32(w)
    assert float(((hi + lo) - w).abs().max() / w.abs().max()) <= 2.0 ** -22
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0          # hi and lo are TF32-representable
    assert int((lo.view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert float((lo.abs() / w.abs().clamp_min(1e-30)).max()) <= 2.0 ** -11 + 1e-7
    hit, lot = ops.split_tf32(w, transpose=True)
    assert torch.equal(hit, hi.t()) and torch.equal(lot, lo.t())


WG_SHAPES = [
    (37376, [160, 160], [160, 160]),     # ShellConv [dz0 | g]^T [x | agg]
    (5000, [160], [160]),                # MLP block
    (3001, [384, 160], [256]),           # embedding projection, ragged row count
    (2048, [512], [512]),                # head
    (300, [32], [64, 32]),               # small, fewer k-blocks than SMs
]


@pytest.mark.parametrize("rows,gw,xw", WG_SHAPES)
def test_gemm_tc_wgrad_matches_float64(rows, gw, xw):
    ops = _ops()
    rng = np.random.Generator(np.random.PCG64(rows))
    G = _segs(rng, rows, gw)
    X = _segs(rng, rows, xw)
    No, Ki = sum(gw), sum(xw)
    old = ops.TC_MIN_ROWS
    ops.TC_MIN_ROWS = 1
    try:
        t = ops.KernelTimer()
        with t:
            dW, db = ops._weight_grad(list(zip(G, gw)), list(zip(X, xw)), rows, No, Ki, DEV, bias=True)
        assert "gemm_tc_wgrad" in t.events
    finally:
        ops.TC_MIN_ROWS = old
    Gd, Xd = torch.cat(G, 1).double(), torch.cat(X, 1).double()
    ref = Gd.t() @ Xd
    # zero-mean random operands: the sum over `rows` terms is ~sqrt(rows) while its conditioning sum |g||x| is ~rows,
    # so the fp32 bar is stated against the conditioning (and 1e-5 of the result scale for the short sums)
    cond = float((Gd.abs().t() @ Xd.abs()).max())
    err = float((dW.double() - ref).abs().max())
    assert err <= max(4e-7 * cond, 1e-5 * float(ref.abs().max())) , f"abs error {err:.3e}, cond {cond:.3e}"
    # fused bias gradient = column sums of G
    db_ref = Gd.sum(0)
    assert float((db.double() - db_ref).abs().max()) <= max(4e-7 * float(Gd.abs().sum(0).max()), 1e-5 * float(db_ref.abs().max()))
    # deterministic: fixed-order split-K reduction
    dW2 = ops._weight_grad(list(zip(G, gw)), list(zip(X, xw)), rows, No, Ki, DEV)
    assert torch.equal(dW, dW2)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("N", [32, 64, 96, 128, 160, 192, 224, 256, 320, 384, 512, 544])
def test_gemm_tc_persistent_tiles_every_tile_width(N):
    """Several tiles per CTA (the persistent loop, both TMEM accumulator buffers, ring wrap-around) for every tile
    width / TMEM layout the host can choose, with residual + bias epilogue; also guards against a hang."""
    ops = _ops()
    M = 148 * 128 * 2 + 77            # > 2 row tiles per SM, ragged tail
    K = 48
    rng = np.random.Generator(np.random.PCG64(N))
    a = _segs(rng, M, [K])
    W = torch.from_numpy(rng.normal(0, 1.0 / np.sqrt(K), size=(N, K)).astype(np.float32)).to(DEV)
    bias = torch.from_numpy(rng.normal(0, 0.1, size=N).astype(np.float32)).to(DEV)
    res = _segs(rng, M, [N])[0]
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm([(a[0], K)], [(W, K)], [(out, N)], M, N, K, bias=bias, resid=[(res, N)])
    torch.cuda.synchronize()
    ref = a[0].double() @ W.double().t() + bias.double() + res.double()
    assert _rel(out, ref) <= 1e-5
