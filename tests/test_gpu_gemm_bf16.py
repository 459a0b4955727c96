"""GPU: the bf16 tensor-core kernels (csrc/gemm_bf16.cu) through the C ABI against a float64 evaluation of the same formulas on
the SAME bf16-rounded inputs.  Bar: 2e-2 relative (BASELINE.json bf16 tolerance) is the model-level bar; at kernel level the
only error sources are the fp32 accumulation order and the final bf16 rounding of the output (2^-9 relative), so the kernels
are held to 1e-2 of the tensor scale for bf16 outputs and 1e-4 for fp32 outputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bf(t):
    return t.to(torch.bfloat16)


def _act(name, v):
    import torch.nn.functional as F
    return {"silu": F.silu, "relu": F.relu, "gelu": F.gelu, "elu": F.elu, "leakyrelu": lambda x: F.leaky_relu(x, 0.01), None: lambda x: x}[name](v)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("M,widths,N", [(300, [160], 160), (37632, [160, 160], 320), (2048, [512], 512), (129, [64, 96], 96),
                                        (5000, [384, 160], 512), (1000, [256], 544)])
def test_gemm_bf16_plain_matches_float64(M, widths, N):
    from aimnet_x2d_b200 import ops
    torch.manual_seed(M + N)
    K = sum(widths)
    a = [_bf(torch.randn(M, w, device=DEV)) for w in widths]
    W = _bf(torch.randn(N, K, device=DEV) / K ** 0.5)
    bias = torch.randn(N, device=DEV)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.gemm_bf16(list(zip(a, widths)), W, [(out, N)], M, N, K, bias=bias)
    torch.cuda.synchronize()
    ref = torch.cat([t.double() for t in a], 1) @ W.double().t() + bias.double()
    err = float((out.double() - ref).abs().max())
    assert err <= 1e-2 * float(ref.abs().max()), (err, float(ref.abs().max()))
    out32 = torch.full((M, N), float("nan"), dtype=torch.float32, device=DEV)
    ops.gemm_bf16(list(zip(a, widths)), W, [(out32, N)], M, N, K, bias=bias, out_dtype=torch.float32)
    err = float((out32.double() - ref).abs().max())
    assert err <= 1e-4 * float(ref.abs().max()), (err, float(ref.abs().max()))


@pytest.mark.timeout(120)
@pytest.mark.parametrize("act", ["silu", "gelu", "relu"])
def test_gemm_bf16_fused_epilogue(act):
    """Segmented outputs [h | gskip], pre-activation copy on the first segment, activation on a column prefix, hash dropout,
    residuals (1 and 3), then the backward-style epilogue (act' * dropout) with identical keep decisions."""
    from aimnet_x2d_b200 import ops
    torch.manual_seed(3)
    M, D = 4000, 160
    x, ag = _bf(torch.randn(M, D, device=DEV)), _bf(torch.randn(M, D, device=DEV))
    W = _bf(torch.randn(2 * D, 2 * D, device=DEV) / (2 * D) ** 0.5)
    b = torch.randn(2 * D, device=DEV) * 0.1
    h = torch.empty((M, D), dtype=torch.bfloat16, device=DEV)
    gs = torch.empty_like(h)
    z0 = torch.empty_like(h)
    ops.gemm_bf16([(x, D), (ag, D)], W, [(h, D), (gs, D)], M, 2 * D, 2 * D, bias=b, pre_segs=[(z0, D), (None, D)], act=act, act_cols=D)
    torch.cuda.synchronize()
    ref = torch.cat([x, ag], 1).double() @ W.double().t() + b.double()
    scale = float(ref.abs().max())
    assert float((z0.double() - ref[:, :D]).abs().max()) <= 1e-2 * scale
    assert float((h.double() - _act(act, ref[:, :D])).abs().max()) <= 1e-2 * scale
    assert float((gs.double() - ref[:, D:]).abs().max()) <= 1e-2 * scale
    # MLP block: t = drop(act(h W1^T + b1)) with u = pre-activation; out = t W2^T + b2 + h + gs + x
    W1 = _bf(torch.randn(D, D, device=DEV) / D ** 0.5)
    W2 = _bf(torch.randn(D, D, device=DEV) / D ** 0.5)
    b1 = torch.randn(D, device=DEV) * 0.1
    u = torch.empty_like(h)
    t = torch.empty_like(h)
    tick = torch.tensor([5], dtype=torch.int64, device=DEV)
    ops.gemm_bf16([(h, D)], W1, [(t, D)], M, D, D, bias=b1, pre_segs=[(u, D)], act=act, drop_p=0.25, drop_seed=77, drop_tick=tick)
    uref = h.double() @ W1.double().t() + b1.double()
    assert float((u.double() - uref).abs().max()) <= 1e-2 * float(uref.abs().max())
    tact = _act(act, u.double())
    kept = t.double().abs() > 0
    frac = float(kept.double().mean())
    nz = float((tact.abs() > 1e-3).double().mean())
    assert abs(frac - 0.75 * nz) < 0.03 + (1 - nz), (frac, nz)             # ~ 25 % dropped
    mask = torch.where(tact.abs() > 1e-3, t.double() / tact, torch.zeros_like(tact))
    on = mask[(tact.abs() > 1e-3) & kept]
    assert float((on - 1 / 0.75).abs().max()) < 0.03                       # kept entries are scaled by 1 / (1 - p)
    out = torch.empty_like(h)
    ops.gemm_bf16([(t, D)], W2, [(out, D)], M, D, D, bias=b1, resid=[(h, D), (gs, D), (x, D)])
    oref = t.double() @ W2.double().t() + b1.double() + h.double() + gs.double() + x.double()
    assert float((out.double() - oref).abs().max()) <= 1e-2 * float(oref.abs().max())
    # backward epilogue: du = (g W2) * act'(u) * drop -- zero exactly where the forward dropped
    g = _bf(torch.randn(M, D, device=DEV))
    W2T = W2.t().contiguous()
    du = torch.empty_like(h)
    ops.gemm_bf16([(g, D)], W2T, [(du, D)], M, D, D, dact_pre=u, dact=act, drop_p=0.25, drop_seed=77, drop_tick=tick)
    ud = u.double().requires_grad_(True)
    _act(act, ud).sum().backward()
    duref = (g.double() @ W2.double()) * ud.grad * torch.where(kept | (tact.abs() <= 1e-3), torch.full_like(tact, 1 / 0.75), torch.zeros_like(tact))
    sure = tact.abs() > 1e-3
    err = float(((du.double() - duref) * sure).abs().max())
    assert err <= 1.5e-2 * float(duref.abs().max()), err
    assert float((du.double()[sure & ~kept]).abs().max()) == 0.0


@pytest.mark.timeout(120)
@pytest.mark.parametrize("rows,gw,xw", [(37632, [160], [160]), (37632, [160, 160], [160, 160]), (2048, [512], [512, 512]),
                                        (1000, [32], [256]), (70, [96], [64])])
def test_gemm_bf16_wgrad_matches_float64(rows, gw, xw):
    from aimnet_x2d_b200 import ops
    torch.manual_seed(rows + sum(gw))
    G = [_bf(torch.randn(rows, w, device=DEV)) for w in gw]
    X = [_bf(torch.randn(rows, w, device=DEV) + 0.5) for w in xw]
    No, Ki = sum(gw), sum(xw)
    dW, db = ops.weight_grad_bf16(list(zip(G, gw)), list(zip(X, xw)), rows, No, Ki, bias=True)
    torch.cuda.synchronize()
    Gd, Xd = torch.cat([t.double() for t in G], 1), torch.cat([t.double() for t in X], 1)
    ref = Gd.t() @ Xd
    cond = (Gd.abs().t() @ Xd.abs())
    assert float(((dW.double() - ref).abs() / cond).max()) <= 2e-6          # fp32 accumulation of exact bf16 products
    refb = Gd.sum(0)
    assert float((db.double() - refb).abs().max()) <= 2e-6 * float(Gd.abs().sum(0).max())
    # deferred mode: the partials left in the workspace sum to the same matrix
    lib = ops._lib.load()
    ws = torch.zeros(lib.ax2d_gemm_bf16_wgrad_workspace(No, Ki, rows) // 4, dtype=torch.float32, device=DEV)
    splits = ops.weight_grad_bf16(list(zip(G, gw)), list(zip(X, xw)), rows, No, Ki, ws=ws, leave_partials=True)
    parts = ws[: splits * No * Ki].view(splits, No, Ki).sum(0)
    assert float((parts - dW).abs().max()) <= 1e-5 * float(dW.abs().max())
    pb = ws[splits * No * Ki: splits * No * Ki + splits * No].view(splits, No).sum(0)
    assert float((pb - db).abs().max()) <= 1e-5 * float(db.abs().max()) + 1e-6


def test_convert_and_act_bwd_bf16():
    from aimnet_x2d_b200 import ops
    x = torch.randn(1000, 160, device=DEV)
    xb = ops.convert(x, torch.bfloat16)
    assert torch.equal(xb, x.to(torch.bfloat16))
    assert torch.equal(ops.convert(xb, torch.float32), xb.float())
