"""Python host side of the C ABI: thin launch wrappers + the autograd Functions of the hot path.

torch is used here for device memory (caching allocator), the current CUDA stream and autograd
bookkeeping only -- every computation on atom / molecule tensors is a libax2d kernel.  There is no
CPU path: calling any op with a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .packed import packed_info
from ._lib import ACT_CODES, CMat, Epilogue, MAX_SEG, SEG_MODES


class KernelTimer:
    """CUDA-event timing of individual C-ABI launches on the launching (= torch current) stream.

    Installed with ``with timer:`` (sets ``_lib.TIMER_HOOK``); every kernel-enqueueing entry point of libax2d is then
    bracketed by an event pair.  While a CUDA graph is being captured the events are *external* event-record nodes of
    the graph: after a replay ``summary()`` gives the duration of every launch INSIDE the replayed step, which is what
    ``bench.py`` computes its roofline fractions and per-kernel shares from (an event pair around a single ~20 us launch
    issued eagerly from Python mostly measures the launch path)."""

    def __init__(self):
        self.events = {}          # name -> list of (start, stop, algorithmic_bytes, flops)
        self._meta = (0, 0)
        self._alias = None

    def __enter__(self):
        global TIMER
        self._prev = (_lib.TIMER_HOOK[0], TIMER)
        # event records between the launches: programmatic dependent launch off while the brackets are in place
        self._pdl = _lib.load().ax2d_set_pdl(0)
        _lib.TIMER_HOOK[0] = self
        TIMER = self
        return self

    def __exit__(self, *exc):
        global TIMER
        _lib.TIMER_HOOK[0], TIMER = self._prev
        _lib.load().ax2d_set_pdl(self._pdl)
        return False

    def bracket(self, name, fn, args):
        ext = torch.cuda.is_current_stream_capturing()
        a = torch.cuda.Event(enable_timing=True, external=ext)
        b = torch.cuda.Event(enable_timing=True, external=ext)
        a.record()
        rc = fn(*args)
        b.record()
        nbytes, flops = self._meta
        self._meta = (0, 0)
        self.events.setdefault(self._alias or name, []).append((a, b, nbytes, flops))
        return rc

    def launch(self, name, fn, nbytes=0, flops=0):
        """Attach the algorithmic bytes / flops of the next launch (recorded under ``name``)."""
        self._meta, self._alias = (nbytes, flops), name
        try:
            fn()
        finally:
            self._meta, self._alias = (0, 0), None

    def summary(self, replays: int = 1):
        """name -> dict(launches, ms_total, ms_avg, bytes_avg, flops_avg); call after a device synchronize.  For events
        recorded inside a captured graph the values are those of the LAST replay."""
        out = {}
        for name, evs in self.events.items():
            ms = [a.elapsed_time(b) for a, b, _, _ in evs]
            out[name] = dict(launches=len(evs), ms_total=sum(ms), ms_avg=sum(ms) / len(ms),
                             bytes_avg=sum(e[2] for e in evs) / len(evs), flops_avg=sum(e[3] for e in evs) / len(evs),
                             bytes_total=sum(e[2] for e in evs), flops_total=sum(e[3] for e in evs))
        return out


TIMER: Optional[KernelTimer] = None


def launch_count() -> int:
    """Kernels enqueued by libax2d.so in this process so far."""
    return int(_lib.load().ax2d_launch_count())


def pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("aimnet_x2d_b200 ops run on CUDA tensors only (there is no CPU fallback); got a "
                               f"{t.device} tensor")


def _mat(segs: Sequence, allow_none: bool = False, dtype=torch.float32) -> CMat:
    """segs: sequence of (tensor | None, width) -- 2-D tensors of ``dtype`` with unit column stride."""
    m = CMat()
    if len(segs) > MAX_SEG:
        raise RuntimeError(f"at most {MAX_SEG} segments are supported, got {len(segs)}")
    m.n_seg = len(segs)
    for i, (t, w) in enumerate(segs):
        if t is None:
            if not allow_none:
                raise RuntimeError("segment tensor is None")
            m.ptr[i], m.ld[i] = None, 4
        else:
            _need_cuda(t)
            if t.dtype != dtype or t.dim() != 2 or t.stride(1) != 1:
                raise RuntimeError(f"segment must be a 2-D {dtype} tensor with unit column stride, got {t.dtype} "
                                   f"{tuple(t.shape)} strides {t.stride()}")
            m.ptr[i], m.ld[i] = t.data_ptr(), t.stride(0)
        m.width[i] = int(w)
    return m


# ----------------------------------------------------------------------------------------------- raw launches
USE_TENSOR_CORES = True     # dense contractions with M >= TC_MIN_ROWS run on tcgen05 (3xTF32 split); SIMT fp32 otherwise
TC_MIN_ROWS = 256


def split_tf32(w: torch.Tensor, transpose: bool = False):
    """(hi, lo) fp32 pair with hi + lo == w (or w^T), hi representable in TF32 -- the weight operand of ax2d_gemm_tc."""
    lib = _lib.load()
    rows, cols = (w.shape[1], w.shape[0]) if transpose else (w.shape[0], w.shape[1])
    hi = torch.empty((rows, cols), dtype=torch.float32, device=w.device)
    lo = torch.empty_like(hi)
    _lib.check(lib.ax2d_split_tf32(_p(w), w.stride(0), rows, cols, int(transpose), _p(hi), _p(lo), cols, _stream()),
               "ax2d_split_tf32")
    return hi, lo


def gemm(a_segs, b_segs, c_segs, M: int, N: int, K: int, trans_a: bool = False, trans_b: bool = True, *,
         bias=None, pre_segs=None, act=None, act_cols=None, mask=None, drop_p: float = 0.0, drop_seed: int = 0,
         drop_tick=None, resid=(), dact_pre=None, dact=None, dact_cols=None, accumulate: bool = False,
         split_k: int = 1) -> None:
    lib = _lib.load()
    if a_segs[0][0].dtype == BF16:          # bf16 configuration: activations bf16, weights from the bf16 packed copies
        if trans_a or split_k != 1 or len(b_segs) != 1 or mask is not None or accumulate:
            raise RuntimeError("bf16 projections support C = epilogue(A W^T) / (A W) with one weight operand only")
        w = b_segs[0][0]
        info = packed_info(w)
        if info is not None and info.wb is not None:
            wb = info.wb if trans_b else info.wbT
            if tuple(wb.shape) != (N, K):
                raise RuntimeError(f"packed bf16 weight {tuple(wb.shape)} does not match the contraction ({N}, {K})")
        else:                               # stand-alone module: derive the operand per call
            wb = (w[:N, :K] if trans_b else w[:K, :N].t()).to(BF16).contiguous()
        return gemm_bf16(a_segs, wb, c_segs, M, N, K, bias=bias, pre_segs=pre_segs, act=act, act_cols=act_cols, drop_p=drop_p,
                         drop_seed=drop_seed, drop_tick=drop_tick, resid=resid, dact_pre=dact_pre, dact=dact,
                         dact_cols=dact_cols, out_dtype=c_segs[0][0].dtype)
    a, b, c = _mat(a_segs), _mat(b_segs), _mat(c_segs)
    ep = Epilogue()
    ep.bias = None if bias is None else bias.data_ptr()
    if pre_segs:
        ep.pre = _mat(pre_segs, allow_none=True)
    ep.act = ACT_CODES[act]
    ep.act_cols = int(N if act_cols is None else act_cols)
    if mask is not None:
        ep.mask, ep.ld_mask = mask.data_ptr(), mask.stride(0)
    ep.drop_p, ep.drop_seed = float(drop_p), int(drop_seed) & 0xFFFFFFFFFFFFFFFF
    ep.drop_tick = None if drop_tick is None else drop_tick.data_ptr()
    if resid:
        ep.resid = _mat([(r, w) for r, w in resid])
    if dact_pre is not None:
        ep.dact_pre, ep.ld_dact, ep.dact = dact_pre.data_ptr(), dact_pre.stride(0), ACT_CODES[dact]
        ep.dact_cols = int(N if dact_cols is None else dact_cols)
    ep.accumulate = int(accumulate)
    if (USE_TENSOR_CORES and not trans_a and split_k == 1 and M >= TC_MIN_ROWS and len(b_segs) == 1
            and lib.ax2d_gemm_tc_supported(C.byref(a), M, N, K)):
        # y = x W^T (trans_b) takes W [N, K] as it is; dx = dy W (not trans_b) needs W^T, produced by the split kernel
        w = b_segs[0][0]
        info = packed_info(w)
        if info is not None and info.hi is not None:      # operands kept ready by packed.PackedWeights
            hi, lo = (info.hi, info.lo) if trans_b else (info.hiT, info.loT)
            if tuple(hi.shape) != (N, K):
                raise RuntimeError(f"packed weight {tuple(hi.shape)} does not match the contraction ({N}, {K})")
        else:
            hi, lo = split_tf32(w[:N, :K] if trans_b else w[:K, :N], transpose=not trans_b)
        call = lambda: _lib.check(lib.ax2d_gemm_tc(C.byref(a), _p(hi), _p(lo), hi.stride(0), C.byref(c), M, N, K,
                                                   C.byref(ep), _stream()), "ax2d_gemm_tc")
        if TIMER is None:
            call()
        else:
            TIMER.launch("gemm_tc", call, nbytes=4 * (M * K + K * N + M * N), flops=2 * M * N * K)
        return
    ws = None
    if split_k > 1:
        nbytes = lib.ax2d_gemm_workspace(M, N, K, int(trans_a), split_k)
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=c_segs[0][0].device)
    call = lambda: _lib.check(lib.ax2d_gemm(C.byref(a), int(trans_a), C.byref(b), int(trans_b), C.byref(c), M, N, K,
                                            C.byref(ep), split_k, _p(ws), _stream()), "ax2d_gemm")
    if TIMER is None:
        call()
    else:
        TIMER.launch("gemm", call, nbytes=4 * (M * K + K * N + M * N), flops=2 * M * N * K)


BF16 = torch.bfloat16
DT_CODES = {torch.float32: 0, torch.bfloat16: 1}


def gemm_bf16(a_segs, w: torch.Tensor, c_segs, M: int, N: int, K: int, *, bias=None, pre_segs=None, act=None, act_cols=None,
              drop_p: float = 0.0, drop_seed: int = 0, drop_tick=None, resid=(), dact_pre=None, dact=None, dact_cols=None,
              out_dtype=torch.bfloat16) -> None:
    """C = epilogue(A W^T) on the tensor cores in bf16 (fp32 accumulation): ``a_segs`` bf16 column segments, ``w`` bf16
    [N, K] (K-major), ``c_segs`` segments of ``out_dtype`` (bf16 or fp32), epilogue operands bf16, ``bias`` fp32."""
    lib = _lib.load()
    a = _mat(a_segs, dtype=BF16)
    c = _mat(c_segs, dtype=out_dtype)
    if w.dtype != BF16 or w.dim() != 2 or w.stride(1) != 1 or w.shape[0] < N or w.shape[1] < K:
        raise RuntimeError(f"bf16 weight operand must be a [{N}, {K}] K-major bf16 matrix, got {w.dtype} {tuple(w.shape)}")
    ep = Epilogue()
    ep.bias = None if bias is None else bias.data_ptr()
    if bias is not None and bias.dtype != torch.float32:
        raise RuntimeError("bias stays fp32 in the bf16 configuration")
    if pre_segs:
        ep.pre = _mat(pre_segs, allow_none=True, dtype=BF16)
    ep.act = ACT_CODES[act]
    ep.act_cols = int(N if act_cols is None else act_cols)
    ep.drop_p, ep.drop_seed = float(drop_p), int(drop_seed) & 0xFFFFFFFFFFFFFFFF
    ep.drop_tick = None if drop_tick is None else drop_tick.data_ptr()
    if resid:
        ep.resid = _mat([(r, wd) for r, wd in resid], dtype=BF16)
    if dact_pre is not None:
        if dact_pre.dtype != BF16:
            raise RuntimeError("dact_pre must be bf16")
        ep.dact_pre, ep.ld_dact, ep.dact = dact_pre.data_ptr(), dact_pre.stride(0), ACT_CODES[dact]
        ep.dact_cols = int(N if dact_cols is None else dact_cols)
    call = lambda: _lib.check(lib.ax2d_gemm_bf16(C.byref(a), _p(w), w.stride(0), C.byref(c), DT_CODES[out_dtype], M, N, K,
                                                 C.byref(ep), _stream()), "ax2d_gemm_bf16")
    if TIMER is None:
        call()
    else:
        TIMER.launch("gemm_bf16", call, nbytes=2 * (M * K + K * N + M * N), flops=2 * M * N * K)


def weight_grad_bf16(g_segs, x_segs, M_rows: int, Nout: int, Kin: int, bias: bool = False, out=None, out_bias=None, ws=None,
                     leave_partials: bool = False):
    """dW[o, i] = sum_r G[r, o] X[r, i] (fp32) from bf16 activations; with ``bias`` also db[o] = sum_r G[r, o].
    ``leave_partials``: the fp32 split partials stay in ``ws`` (summed later by ``ax2d_unpack_grads``); returns the split count."""
    lib = _lib.load()
    a, b = _mat(g_segs, dtype=BF16), _mat(x_segs, dtype=BF16)
    if not lib.ax2d_gemm_bf16_wgrad_supported(C.byref(a), C.byref(b), Nout, Kin, M_rows):
        raise RuntimeError(f"bf16 weight gradient [{Nout}, {Kin}] over {M_rows} rows: segment widths must be multiples of 32")
    dev = g_segs[0][0].device
    if ws is None:
        ws = torch.empty(max(lib.ax2d_gemm_bf16_wgrad_workspace(Nout, Kin, M_rows) // 4, 1), dtype=torch.float32, device=dev)
    if leave_partials:
        call = lambda: _lib.check(lib.ax2d_gemm_bf16_wgrad(C.byref(a), C.byref(b), None, Nout, Kin, M_rows, 2, None, _p(ws),
                                                           _stream()), "ax2d_gemm_bf16_wgrad")
    else:
        dW = torch.empty((Nout, Kin), dtype=torch.float32, device=dev) if out is None else out
        db = (torch.empty(Nout, dtype=torch.float32, device=dev) if out_bias is None else out_bias) if bias else None
        c = _mat([(dW, Kin)])
        call = lambda: _lib.check(lib.ax2d_gemm_bf16_wgrad(C.byref(a), C.byref(b), C.byref(c), Nout, Kin, M_rows, 0, _p(db),
                                                           _p(ws), _stream()), "ax2d_gemm_bf16_wgrad")
    if TIMER is None:
        call()
    else:
        TIMER.launch("gemm_bf16_wgrad", call, nbytes=2 * M_rows * (Nout + Kin), flops=2 * M_rows * Nout * Kin)
    if leave_partials:
        return lib.ax2d_gemm_bf16_wgrad_splits(Nout, Kin, M_rows)
    return (dW, db) if bias else dW


def convert(x: torch.Tensor, dtype) -> torch.Tensor:
    """fp32 <-> bf16 copy of a 2-D matrix (width % 8 == 0) through ax2d_convert."""
    if x.dtype == dtype:
        return x
    x2 = x if x.dim() == 2 else x.reshape(-1, x.shape[-1])
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    out = torch.empty(x2.shape, dtype=dtype, device=x.device)
    _lib.check(_lib.load().ax2d_convert(_p(x2), x2.stride(0), DT_CODES[x.dtype], _p(out), out.stride(0), DT_CODES[dtype],
                                        x2.shape[0], x2.shape[1], _stream()), "ax2d_convert")
    return out.reshape(x.shape)


def colsum(segs, M: int, N: int, out: torch.Tensor, accumulate: bool = False) -> None:
    lib = _lib.load()
    ws = torch.empty(max(lib.ax2d_colsum_workspace(M, N) // 4, 1), dtype=torch.float32, device=out.device)
    a = _mat(segs)
    _lib.check(lib.ax2d_colsum(C.byref(a), M, N, _p(out), int(accumulate), _p(ws), _stream()), "ax2d_colsum")


def act_bwd(g: torch.Tensor, pre: torch.Tensor, act: str) -> torch.Tensor:
    out = torch.empty_like(pre)
    if pre.dtype == BF16:
        g = convert(g, BF16)
        _lib.check(_lib.load().ax2d_act_bwd_bf16(_p(g), g.stride(0), _p(pre), pre.stride(0), _p(out), out.stride(0),
                                                 pre.shape[0], pre.shape[1], ACT_CODES[act], _stream()), "ax2d_act_bwd_bf16")
        return out
    _lib.check(_lib.load().ax2d_act_bwd(_p(g), g.stride(0), _p(pre), pre.stride(0), _p(out), out.stride(0),
                                        pre.shape[0], pre.shape[1], ACT_CODES[act], _stream()), "ax2d_act_bwd")
    return out


def split_k_for(M: int, N: int, K: int) -> int:
    tiles = ((M + 127) // 128) * ((N + 63) // 64)
    s = max(1, min((148 * 3 + tiles - 1) // tiles, K // 256))
    return int(s)


# Whole-molecule tiles as dense products on the tensor cores (ax2d_agg_tiles_mma).  Correct (tests/test_gpu_parity.py) but as
# measured on B200 not faster than the per-edge kernels yet (C2 shapes, fwd / bwd: fp32 27.9 / 53.5 us against 18.1 / 19.4 us;
# bf16 18.2 / 23.9 us against 19.1 / 20.6 us, profiles/r2e_agg_variants.txt), so it is OFF by default: the per-edge kernels
# are also the ones that are bit-identical to the reference's sequential scatter_add.
AGG_TENSOR_CORES = False


def agg(x: torch.Tensor, gi, transpose: bool = False, addend: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Forward (transpose=False): out[r] = sum x[col]; rows = gi.num_rows.  Backward: rows = N."""
    lib = _lib.load()
    _need_cuda(x)
    N, width = gi.num_atoms, x.shape[1]
    rows = N if transpose else gi.num_rows
    out = torch.empty((rows, width), dtype=x.dtype, device=x.device)
    rowptr, col, info = (gi.rowptr_t, gi.col_t, gi.tile_info_t) if transpose else (gi.rowptr, gi.col, gi.tile_info)
    es = x.element_size()
    if x.dtype not in DT_CODES or (addend is not None and addend.dtype != x.dtype):
        raise RuntimeError(f"aggregation runs on fp32 or bf16 features, got {x.dtype}")
    tiled = (gi.tile_local and gi.collapsed and x.stride(0) == width and width % 32 == 0 and info is not None
             and gi.max_tile_rows * width * es + gi.max_tile_edges * 4 <= 200 * 1024 and gi.n_tiles > 0)
    if (tiled and AGG_TENSOR_CORES and getattr(gi, "unique_edges", False)
            and lib.ax2d_agg_tiles_mma_supported(width, gi.max_tile_rows, gi.max_tile_edges, DT_CODES[x.dtype])):
        # whole-molecule tiles as dense products on the tensor cores (fp32: same sums, other order of the fp32 additions)
        call = lambda: _lib.check(lib.ax2d_agg_tiles_mma(_p(x), x.stride(0), rows, _p(out), out.stride(0), _p(rowptr), _p(col),
                                                         _p(addend), 0 if addend is None else addend.stride(0), width, _p(info),
                                                         gi.n_tiles, gi.max_tile_rows, gi.max_tile_edges, DT_CODES[x.dtype],
                                                         _stream()), "ax2d_agg_tiles_mma")
        if TIMER is None:
            call()
        else:
            TIMER.launch("agg", call, nbytes=agg_bytes(x.shape[0], rows, gi.num_edges, width, addend is not None, es))
        return out
    call = lambda: _lib.check(lib.ax2d_agg(_p(x), x.stride(0), x.shape[0], _p(out), out.stride(0), rows, _p(rowptr),
                                           _p(col), _p(addend), 0 if addend is None else addend.stride(0), width,
                                           _p(info) if tiled else None, gi.n_tiles if tiled else 0,
                                           gi.max_tile_rows if tiled else 0, gi.max_tile_edges if tiled else 0,
                                           DT_CODES[x.dtype], _stream()), "ax2d_agg")
    if TIMER is None:
        call()
    else:
        TIMER.launch("agg", call, nbytes=agg_bytes(x.shape[0], rows, gi.num_edges, width, addend is not None, es))
    return out


def agg_bytes(n_src: int, n_rows: int, n_edges: int, width: int, with_addend: bool = False, elem: int = 4) -> int:
    """ALGORITHMIC HBM bytes of one aggregation launch:
    read x once + col indices + rowptr + write the result (+ read the addend in the fused backward)."""
    return elem * n_src * width + 4 * n_edges + 4 * (n_rows + 1) + elem * n_rows * width * (2 if with_addend else 1)


# ----------------------------------------------------------------------------------------------- dense helpers
def _bias_grad(segs, M, N, device):
    out = torch.empty(N, dtype=torch.float32, device=device)
    colsum(segs, M, N, out)
    return out


def _weight_grad(g_segs, x_segs, M_rows: int, Nout: int, Kin: int, device, bias: bool = False, out=None, out_bias=None):
    """dW[o, i] = sum_r G[r, o] X[r, i]  (G, X column-segmented) and, with ``bias``, db[o] = sum_r G[r, o].
    Returns dW or (dW, db).  Tensor-core kernel (split over the rows, fixed-order reduce, bias gradient fused)
    when the operands qualify; exact-fp32 SIMT GEMM + column-sum kernels otherwise.  ``out`` / ``out_bias``:
    write into these (contiguous) buffers instead of fresh tensors."""
    dW = torch.empty((Nout, Kin), dtype=torch.float32, device=device) if out is None else out
    if tuple(dW.shape) != (Nout, Kin) or not dW.is_contiguous():
        raise RuntimeError(f"weight-gradient buffer {tuple(dW.shape)} vs ({Nout}, {Kin})")
    lib = _lib.load()
    if g_segs[0][0].dtype == BF16:
        return weight_grad_bf16(g_segs, x_segs, M_rows, Nout, Kin, bias=bias, out=dW, out_bias=out_bias)
    if USE_TENSOR_CORES and M_rows >= TC_MIN_ROWS:
        a, b, c = _mat(g_segs), _mat(x_segs), _mat([(dW, Kin)])
        if lib.ax2d_gemm_tc_wgrad_supported(C.byref(a), C.byref(b), Nout, Kin, M_rows):
            nbytes = lib.ax2d_gemm_tc_wgrad_workspace(Nout, Kin, M_rows)
            ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=device)
            db = (torch.empty(Nout, dtype=torch.float32, device=device) if out_bias is None else out_bias) if bias else None
            call = lambda: _lib.check(lib.ax2d_gemm_tc_wgrad(C.byref(a), C.byref(b), C.byref(c), Nout, Kin, M_rows, 0, _p(db),
                                                             _p(ws), _stream()), "ax2d_gemm_tc_wgrad")
            if TIMER is None:
                call()
            else:
                TIMER.launch("gemm_tc_wgrad", call, nbytes=4 * M_rows * (Nout + Kin), flops=2 * M_rows * Nout * Kin)
            return (dW, db) if bias else dW
    gemm(g_segs, x_segs, [(dW, Kin)], Nout, Kin, M_rows, trans_a=True, trans_b=False,
         split_k=split_k_for(Nout, Kin, M_rows))
    if not bias:
        return dW
    if out_bias is None:
        return dW, _bias_grad(g_segs, M_rows, Nout, device)
    colsum(g_segs, M_rows, Nout, out_bias)
    return dW, out_bias


DEFER_SPLITK_REDUCE = True      # packed weights: the gradient-collect kernel sums the split-K partials (no reduce launches)


def _deferred_weight_grad(iw, ib, g_segs, x_segs, M_rows: int, Nout: int, Kin: int) -> bool:
    """Tensor-core weight gradient that leaves its split-K partial tiles (and partial bias vectors) in a persistent
    workspace of the packed matrix; ``ax2d_unpack_grads`` sums them in split order while it scatters the gradients,
    which saves one reduce launch per matrix and step.  False: not applicable (caller takes the ordinary path)."""
    lib = _lib.load()
    if g_segs[0][0].dtype == BF16:
        if not DEFER_SPLITK_REDUCE or tuple(iw.grad.shape) != (Nout, Kin):
            return False
        ws = iw.owner.workspace(iw.name, lib.ax2d_gemm_bf16_wgrad_workspace(Nout, Kin, M_rows))
        splits = weight_grad_bf16(g_segs, x_segs, M_rows, Nout, Kin, ws=ws, leave_partials=True)
        iw.partials = (ws.data_ptr(), splits, Nout, Kin)
        if ib is not None:
            ib.partials = (ws.data_ptr() + 4 * splits * Nout * Kin, splits, 1, Nout)
        return True
    if not (DEFER_SPLITK_REDUCE and USE_TENSOR_CORES and M_rows >= TC_MIN_ROWS):
        return False
    a, b = _mat(g_segs), _mat(x_segs)
    if not lib.ax2d_gemm_tc_wgrad_supported(C.byref(a), C.byref(b), Nout, Kin, M_rows):
        return False
    splits = lib.ax2d_gemm_tc_wgrad_splits(Nout, Kin, M_rows)
    if splits <= 1 or tuple(iw.grad.shape) != (Nout, Kin):
        return False
    ws = iw.owner.workspace(iw.name, lib.ax2d_gemm_tc_wgrad_workspace(Nout, Kin, M_rows))
    c = _mat([(iw.grad, Kin)])              # not written in this mode
    want_bias = ib is not None
    call = lambda: _lib.check(lib.ax2d_gemm_tc_wgrad(C.byref(a), C.byref(b), C.byref(c), Nout, Kin, M_rows, 2,
                                                     _p(ib.grad) if want_bias else None, _p(ws), _stream()), "ax2d_gemm_tc_wgrad")
    if TIMER is None:
        call()
    else:
        TIMER.launch("gemm_tc_wgrad", call, nbytes=4 * M_rows * (Nout + Kin), flops=2 * M_rows * Nout * Kin)
    iw.partials = (ws.data_ptr(), splits, Nout, Kin)
    if want_bias:
        ib.partials = (ws.data_ptr() + 4 * splits * Nout * Kin, splits, 1, Nout)
    return True


def _param_grads(W, b, g_segs, x_segs, M_rows: int, Nout: int, Kin: int, device):
    """(dW, db) for autograd.  For packed weights (``packed.PackedWeights``) the gradients go straight into the packed
    gradient buffers -- collected into ``.grad`` by ONE kernel at the end of backward -- and autograd gets (None, None)."""
    iw = packed_info(W)
    if iw is None:
        if b is None:
            return _weight_grad(g_segs, x_segs, M_rows, Nout, Kin, device), None
        return _weight_grad(g_segs, x_segs, M_rows, Nout, Kin, device, bias=True)
    ib = packed_info(b)
    if b is not None and ib is None:
        raise RuntimeError("a packed weight needs a packed bias")
    if not iw.written and _deferred_weight_grad(iw, ib, g_segs, x_segs, M_rows, Nout, Kin):
        pass
    elif not iw.written:
        _weight_grad(g_segs, x_segs, M_rows, Nout, Kin, device, bias=b is not None, out=iw.grad,
                     out_bias=None if b is None else ib.grad)
    else:           # a weight shared by several calls (stereochemical_embedding_2 serves every layer): sum them
        r = _weight_grad(g_segs, x_segs, M_rows, Nout, Kin, device, bias=b is not None)
        iw.grad.add_(r[0] if b is not None else r)
        if b is not None:
            ib.grad.add_(r[1])
    iw.written = True
    iw.owner.dirty = True
    return None, None


class _Opts:
    """Non-tensor options handed to the autograd Functions."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def _keep_packed(ctx, *ws) -> None:
    """Remember the packed weight tensors themselves: the objects autograd hands back from ``saved_tensors`` need not
    carry the ``_ax2d`` attribute."""
    ctx.packed_ws = [w if packed_info(w) is not None else None for w in ws]


def _packed_or(ctx, *saved):
    return [s if p is None else p for p, s in zip(ctx.packed_ws, saved)]


class LinearFn(torch.autograd.Function):
    """y = [a_0 | a_1 | ...] W^T + b without materialising the concatenation (gnn.py:245-246, 252, 256-258, 327).

    args: opts(widths=[w_i], n_out=Np), W [Np, sum w_i], b [Np] | None, *a_segs"""

    @staticmethod
    def forward(ctx, opts, W, b, *a):
        _need_cuda(W, *a)
        M = a[0].shape[0]
        K = sum(opts.widths)
        out = torch.empty((M, opts.n_out), dtype=getattr(opts, "out_dtype", None) or a[0].dtype, device=W.device)
        gemm(list(zip(a, opts.widths)), [(W, K)], [(out, opts.n_out)], M, opts.n_out, K, bias=b)
        ctx.opts = opts
        ctx.bias = b
        _keep_packed(ctx, W)
        ctx.save_for_backward(W, *a)
        return out

    @staticmethod
    def backward(ctx, g):
        W, *a = ctx.saved_tensors
        (W,) = _packed_or(ctx, W)
        opts = ctx.opts
        g = g.contiguous()
        if g.dtype != a[0].dtype:            # fp32 outputs of the bf16 configuration (the last layer feeds the loss)
            if g.shape[1] % 8 != 0:
                g = torch.nn.functional.pad(g, (0, 8 - g.shape[1] % 8))
            g = convert(g, a[0].dtype)
        M, Np, K = g.shape[0], opts.n_out, sum(opts.widths)
        gs = [(g, Np)]
        d_a = [torch.empty((M, w), dtype=a[0].dtype, device=g.device) for w in opts.widths]
        gemm(gs, [(W, K)], list(zip(d_a, opts.widths)), M, K, Np, trans_a=False, trans_b=False)
        dW, db = _param_grads(W, ctx.bias, gs, list(zip(a, opts.widths)), M, Np, K, g.device)
        return (None, dW, db, *d_a)


class MLPBlockFn(torch.autograd.Function):
    """out = W2 drop(act(W1 h + b1)) + b2 (+ h)   (LinearBlock, layers.py:203-219).

    args: opts(act, p, seed, tick, skip, w_in, w_out), h [M, w_in], W1 [w_out, w_in], b1, W2 [w_out, w_out], b2"""

    @staticmethod
    def forward(ctx, opts, h, W1, b1, W2, b2):
        _need_cuda(h, W1, W2)
        M, Wi, Wo = h.shape[0], opts.w_in, opts.w_out
        u = torch.empty((M, Wo), dtype=h.dtype, device=h.device)
        t = torch.empty_like(u)
        out = torch.empty_like(u)
        gemm([(h, Wi)], [(W1, Wi)], [(t, Wo)], M, Wo, Wi, bias=b1, pre_segs=[(u, Wo)], act=opts.act,
             drop_p=opts.p, drop_seed=opts.seed, drop_tick=opts.tick)
        gemm([(t, Wo)], [(W2, Wo)], [(out, Wo)], M, Wo, Wo, bias=b2, resid=[(h, Wo)] if opts.skip else [])
        ctx.opts = opts
        ctx.biases = (b1, b2)
        _keep_packed(ctx, W1, W2)
        ctx.save_for_backward(h, W1, W2, u, t)
        return out

    @staticmethod
    def backward(ctx, g):
        h, W1, W2, u, t = ctx.saved_tensors
        W1, W2 = _packed_or(ctx, W1, W2)
        b1, b2 = ctx.biases
        opts = ctx.opts
        g = g.contiguous()
        M, Wi, Wo = g.shape[0], opts.w_in, opts.w_out
        dev = g.device
        gs = [(g, Wo)]
        dW2, db2 = _param_grads(W2, b2, gs, [(t, Wo)], M, Wo, Wo, dev)
        du = torch.empty_like(u)
        gemm(gs, [(W2, Wo)], [(du, Wo)], M, Wo, Wo, trans_b=False, dact_pre=u, dact=opts.act,
             drop_p=opts.p, drop_seed=opts.seed, drop_tick=opts.tick)
        dus = [(du, Wo)]
        dW1, db1 = _param_grads(W1, b1, dus, [(h, Wi)], M, Wo, Wi, dev)
        dh = torch.empty_like(h)
        gemm(dus, [(W1, Wi)], [(dh, Wi)], M, Wi, Wo, trans_b=False, resid=[(g, Wi)] if opts.skip else [])
        return None, dh, dW1, db1, dW2, db2


class ShellConvFn(torch.autograd.Function):
    """One ShellConvolutionLayer (layers.py:63-108) on padded features, optionally with the caller's
    ``+ x`` residual (gnn.py:302-306) folded into the last epilogue.

    args: opts(act, ps, seeds, tick, w_in, w_out, n_mlp, add_input, gi), x, W_io, b_io, (W1, b1, W2, b2) * n_mlp
    W_io = [W_in ; W_skip] stacked over rows, columns = the non-zero hop chunks of the padded input."""

    @staticmethod
    def _segments(x, ag, Di):
        N = x.shape[0]
        return [(x, Di)] + [(ag[h * N:(h + 1) * N], Di) for h in range(ag.shape[0] // max(N, 1))]

    @staticmethod
    def forward(ctx, opts, x, W_io, b_io, *mlp):
        _need_cuda(x, W_io)
        gi, Di, Do, N = opts.gi, opts.w_in, opts.w_out, x.shape[0]
        dev = x.device
        ag = agg(x, gi)                                                     # layers.py:133-167
        a_segs = ShellConvFn._segments(x, ag, Di)
        K = Di * len(a_segs)
        z0 = torch.empty((N, Do), dtype=x.dtype, device=dev)
        h = torch.empty_like(z0)
        gskip = torch.empty_like(z0)
        gemm(a_segs, [(W_io, K)], [(h, Do), (gskip, Do)], N, 2 * Do, K, bias=b_io,
             pre_segs=[(z0, Do), (None, Do)], act=opts.act, act_cols=Do)    # layers.py:82-87
        saved = [x, ag, z0, W_io]
        for k in range(opts.n_mlp):
            W1, b1, W2, b2 = mlp[4 * k:4 * k + 4]
            u = torch.empty_like(z0)
            t = torch.empty_like(z0)
            hn = torch.empty_like(z0)
            gemm([(h, Do)], [(W1, Do)], [(t, Do)], N, Do, Do, bias=b1, pre_segs=[(u, Do)], act=opts.act,
                 drop_p=opts.ps[k], drop_seed=opts.seeds[k], drop_tick=opts.tick)
            resid = [(h, Do)]
            if k == opts.n_mlp - 1:                                         # layers.py:106 (+ gnn.py:306)
                resid.append((gskip, Do))
                if opts.add_input:
                    resid.append((x, Do))
            gemm([(t, Do)], [(W2, Do)], [(hn, Do)], N, Do, Do, bias=b2, resid=resid)
            saved += [h, u, t, W1, W2]
            h = hn
        if opts.n_mlp == 0:
            h = h + gskip + (x if opts.add_input else 0)
        ctx.opts = opts
        ctx.biases = (b_io, *[mlp[4 * k + j] for k in range(opts.n_mlp) for j in (1, 3)])
        _keep_packed(ctx, *saved)
        ctx.save_for_backward(*saved)
        return h

    @staticmethod
    def backward(ctx, g):
        opts = ctx.opts
        gi, Di, Do = opts.gi, opts.w_in, opts.w_out
        saved = _packed_or(ctx, *ctx.saved_tensors)
        x, ag, z0, W_io = saved[:4]
        b_io, *b_mlp = ctx.biases
        g = g.contiguous()
        N, dev = g.shape[0], g.device
        dh = g
        grads_mlp: List[torch.Tensor] = []
        if opts.n_mlp == 0:
            dz0 = act_bwd(g, z0, opts.act)
        for k in reversed(range(opts.n_mlp)):
            h, u, t, W1, W2 = saved[4 + 5 * k:9 + 5 * k]
            dhs = [(dh, Do)]
            dW2, db2 = _param_grads(W2, b_mlp[2 * k + 1], dhs, [(t, Do)], N, Do, Do, dev)
            du = torch.empty_like(z0)
            gemm(dhs, [(W2, Do)], [(du, Do)], N, Do, Do, trans_b=False, dact_pre=u, dact=opts.act,
                 drop_p=opts.ps[k], drop_seed=opts.seeds[k], drop_tick=opts.tick)
            dus = [(du, Do)]
            dW1, db1 = _param_grads(W1, b_mlp[2 * k], dus, [(h, Do)], N, Do, Do, dev)
            dprev = torch.empty_like(z0)
            if k == 0:      # d z0 = (dh + du W1) * act'(z0) in one epilogue
                gemm(dus, [(W1, Do)], [(dprev, Do)], N, Do, Do, trans_b=False, resid=dhs, dact_pre=z0, dact=opts.act)
                dz0 = dprev
            else:
                gemm(dus, [(W1, Do)], [(dprev, Do)], N, Do, Do, trans_b=False, resid=dhs)
                dh = dprev
            grads_mlp = [dW1, db1, dW2, db2] + grads_mlp
        # input / skip projections: [dz0 | g] against [x | agg_1 .. agg_H]
        a_segs = ShellConvFn._segments(x, ag, Di)
        K = Di * len(a_segs)
        gz = [(dz0, Do), (g, Do)]
        dW_io, db_io = _param_grads(W_io, b_io, gz, a_segs, N, 2 * Do, K, dev)
        dx1 = torch.empty((N, Di), dtype=ag.dtype, device=dev)
        dag = torch.empty_like(ag)
        c_segs = ShellConvFn._segments(dx1, dag, Di)
        gemm(gz, [(W_io, K)], c_segs, N, K, 2 * Do, trans_b=False,
             resid=[(g, Di)] if opts.add_input else [])                   # + g only on the x columns
        dx = agg(dag, gi, transpose=True, addend=dx1)                       # gather backward, CSR of src
        return (None, dx, dW_io, db_io, *grads_mlp)


class EmbedProjFn(torch.autograd.Function):
    """Embedding lookups + cat + Linear + activation + split into (x_self, x_other)  (gnn.py:220-231, 262-274).

    args: opts(act, names, emb_dim, s_pad, d_pad, gi, vocabs), W [s_pad+d_pad, T*E], b, *tables, then *indices"""

    @staticmethod
    def forward(ctx, opts, W, b, *rest):
        nt = len(opts.names)
        tables, indices = rest[:nt], rest[nt:]
        _need_cuda(W, *tables, *indices)
        lib = _lib.load()
        N, E = indices[0].shape[0], opts.emb_dim
        dev = W.device
        dt = getattr(opts, "dtype", torch.float32)
        e0 = torch.empty((N, nt * E), dtype=dt, device=dev)
        tp = (C.c_void_p * nt)(*[t.data_ptr() for t in tables])
        ip = (C.c_void_p * nt)(*[i.data_ptr() for i in indices])
        for i in indices:
            if i.dtype != torch.int64 or not i.is_contiguous():
                raise RuntimeError("atom feature indices must be contiguous int64 tensors")
        if dt == BF16:
            _lib.check(lib.ax2d_embed_fwd_bf16(tp, ip, nt, E, N, _p(e0), e0.stride(0), _stream()), "ax2d_embed_fwd_bf16")
        else:
            _lib.check(lib.ax2d_embed_fwd(tp, ip, nt, E, N, _p(e0), e0.stride(0), _stream()), "ax2d_embed_fwd")
        Sp, Dp = opts.s_pad, opts.d_pad
        xs = torch.empty((N, Sp), dtype=dt, device=dev)
        xo = torch.empty((N, Dp), dtype=dt, device=dev)
        zs, zo = torch.empty_like(xs), torch.empty_like(xo)
        gemm([(e0, nt * E)], [(W, nt * E)], [(xs, Sp), (xo, Dp)], N, Sp + Dp, nt * E, bias=b,
             pre_segs=[(zs, Sp), (zo, Dp)], act=opts.act)
        ctx.opts = opts
        ctx.bias = b
        ctx.indices = indices          # integer inputs: no gradient, kept for the one-pass table gradient
        _keep_packed(ctx, W)
        ctx.save_for_backward(W, e0, zs, zo, *tables)
        return xs, xo

    @staticmethod
    def backward(ctx, gxs, gxo):
        opts = ctx.opts
        W, e0, zs, zo, *tables = ctx.saved_tensors
        (W,) = _packed_or(ctx, W)
        lib = _lib.load()
        nt, E, Sp, Dp = len(opts.names), opts.emb_dim, opts.s_pad, opts.d_pad
        N, dev = e0.shape[0], e0.device
        dzs = act_bwd(gxs.contiguous(), zs, opts.act)
        dzo = act_bwd(gxo.contiguous(), zo, opts.act)
        gz = [(dzs, Sp), (dzo, Dp)]
        dW, db = _param_grads(W, ctx.bias, gz, [(e0, nt * E)], N, Sp + Dp, nt * E, dev)
        de0 = torch.empty_like(e0)
        gemm(gz, [(W, nt * E)], [(de0, nt * E)], N, nt * E, Sp + Dp, trans_b=False)
        g_tables = [torch.empty_like(t) for t in tables]
        rows = sum(t.shape[0] for t in tables)
        if rows * E * 4 <= 200 * 1024 and all(t.is_contiguous() and t.shape[1] == E for t in tables):
            # all tables in one pass through shared memory (fixed summation order)
            for ti, name in enumerate(opts.names):
                if name in opts.gi.embed and opts.gi.embed[name][2] != tables[ti].shape[0]:
                    raise RuntimeError(f"embedding index for '{name}' was built for vocab {opts.gi.embed[name][2]}, table "
                                       f"has {tables[ti].shape[0]} rows")
            ws = torch.empty(max(lib.ax2d_embed_bwd_all_workspace(N, rows, E) // 4, 1), dtype=torch.float32, device=dev)
            ip = (C.c_void_p * nt)(*[i.data_ptr() for i in ctx.indices])
            gp = (C.c_void_p * nt)(*[t.data_ptr() for t in g_tables])
            vp = (C.c_int64 * nt)(*[t.shape[0] for t in tables])
            fn = lib.ax2d_embed_bwd_all_bf16 if de0.dtype == BF16 else lib.ax2d_embed_bwd_all
            _lib.check(fn(_p(de0), de0.stride(0), nt, E, N, ip, vp, gp, _p(ws), _stream()), "ax2d_embed_bwd_all")
        else:
            if de0.dtype == BF16:
                de0 = convert(de0, torch.float32)
            for ti, name in enumerate(opts.names):
                order, ptr, vocab = opts.gi.embed[name]
                if vocab != tables[ti].shape[0]:
                    raise RuntimeError(f"embedding index for '{name}' was built for vocab {vocab}, table has "
                                       f"{tables[ti].shape[0]} rows")
                gt = g_tables[ti]
                ws = torch.empty(lib.ax2d_embed_bwd_workspace(vocab, E) // 4, dtype=torch.float32, device=dev)
                _lib.check(lib.ax2d_embed_bwd(_p(de0), de0.stride(0), ti, E, vocab, _p(order), _p(ptr), _p(gt), _p(ws),
                                              _stream()), "ax2d_embed_bwd")
        return (None, dW, db, *g_tables, *([None] * nt))


class CastFn(torch.autograd.Function):
    """fp32 <-> bf16 through ``ax2d_convert`` (the bf16 configuration runs segmented softmax, charge equilibration and
    the stereo terms in fp32 on up-cast values)."""

    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src = x.dtype
        return convert(x, dtype)

    @staticmethod
    def backward(ctx, g):
        return convert(g.contiguous(), ctx.src), None


def cast(x: torch.Tensor, dtype) -> torch.Tensor:
    return x if x.dtype == dtype else CastFn.apply(x, dtype)


# ----------------------------------------------------------------------------------------------- segment ops
class AttnPoolFn(torch.autograd.Function):
    """MultiHeadAttentionPoolingLayer.forward (pooling.py:122-172) as one fused kernel per direction."""

    @staticmethod
    def forward(ctx, x, w, b, temperature, gi):
        _need_cuda(x, w, b, temperature)
        lib = _lib.load()
        x = x.contiguous()
        N, F = x.shape
        heads, B = w.shape[0], gi.num_graphs
        pooled = torch.empty((B, F), dtype=torch.float32, device=x.device)
        attn = torch.empty((heads, N), dtype=torch.float32, device=x.device)
        z = torch.empty_like(attn)
        _lib.check(lib.ax2d_attn_pool_fwd(_p(x), F, _p(gi.seg_ptr), B, N, F, heads, _p(w), _p(b), _p(temperature),
                                          _p(pooled), _p(attn), _p(z), gi.max_seg, _stream()), "ax2d_attn_pool_fwd")
        ctx.gi = gi
        ctx.packed_wb = (w, b) if packed_info(w) is not None else None
        ctx.save_for_backward(x, w, temperature, attn, z)
        return pooled, attn

    @staticmethod
    def backward(ctx, g_pooled, g_attn):
        x, w, temperature, attn, z = ctx.saved_tensors
        gi = ctx.gi
        lib = _lib.load()
        N, F = x.shape
        heads, B = w.shape[0], gi.num_graphs
        dev = x.device
        gx = torch.empty_like(x)
        iw = ib = None
        if ctx.packed_wb is not None:      # gradients go to the packed buffers (collected into .grad by one kernel)
            iw, ib = packed_info(ctx.packed_wb[0]), packed_info(ctx.packed_wb[1])
            if iw.written or ib is None:
                iw = ib = None             # (a pooling layer applied twice in one backward: ordinary autograd path)
        gw = torch.empty_like(w) if iw is None else iw.grad
        gb = torch.empty(heads, dtype=torch.float32, device=dev) if ib is None else ib.grad
        gT = torch.empty((), dtype=torch.float32, device=dev)
        ws = torch.empty(lib.ax2d_attn_pool_bwd_workspace(B, F, heads) // 4, dtype=torch.float32, device=dev)
        if g_pooled is None:
            g_pooled = torch.zeros((B, F), dtype=torch.float32, device=dev)
        g_attn = None if g_attn is None else g_attn.contiguous()
        _lib.check(lib.ax2d_attn_pool_bwd(_p(x), F, _p(gi.seg_ptr), B, N, F, heads, _p(w), _p(temperature), _p(attn),
                                          _p(z), _p(g_pooled.contiguous()), _p(g_attn), _p(gx), F, _p(gw), _p(gb),
                                          _p(gT), _p(ws), gi.max_seg, _stream()), "ax2d_attn_pool_bwd")
        if iw is not None:
            iw.written = True
            iw.owner.dirty = True
            return gx, None, None, gT, None
        return gx, gw, gb, gT, None


class SegPoolFn(torch.autograd.Function):
    """Mean / Max / Sum pooling (pooling.py:15-80) with torch_scatter semantics."""

    @staticmethod
    def forward(ctx, x, mode, gi):
        _need_cuda(x)
        lib = _lib.load()
        N, F = x.shape
        B = gi.num_graphs
        out = torch.empty((B, F), dtype=torch.float32, device=x.device)
        arg = torch.empty((B, F), dtype=torch.int32, device=x.device) if mode == "max" else None
        _lib.check(lib.ax2d_seg_reduce_fwd(_p(x), x.stride(0), _p(gi.seg_ptr), B, F, SEG_MODES[mode], _p(out), _p(arg),
                                           _stream()), "ax2d_seg_reduce_fwd")
        ctx.gi, ctx.mode, ctx.shape = gi, mode, (N, F)
        ctx.save_for_backward(arg if arg is not None else out.new_empty(0))
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        gi, mode = ctx.gi, ctx.mode
        N, F = ctx.shape
        gx = torch.empty((N, F), dtype=torch.float32, device=g.device)
        _lib.check(_lib.load().ax2d_seg_reduce_bwd(_p(g.contiguous()), _p(gi.seg_ptr), gi.num_graphs, F, SEG_MODES[mode],
                                                   _p(arg) if mode == "max" else None, _p(gx), F, _stream()),
                   "ax2d_seg_reduce_bwd")
        return gx, None, None


class ChargeEqFn(torch.autograd.Function):
    """GNN._partial_charge_calculation (gnn.py:622-658) on padded features (columns 0 and 1 change)."""

    @staticmethod
    def forward(ctx, x, total_charges, gi):
        _need_cuda(x, total_charges)
        out = torch.empty_like(x)
        stats = torch.empty((gi.num_graphs, 2), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().ax2d_charge_eq_fwd(_p(x), x.stride(0), _p(gi.seg_ptr), gi.num_graphs, x.shape[1],
                                                  _p(total_charges.contiguous()), _p(out), out.stride(0), _p(stats),
                                                  _stream()), "ax2d_charge_eq_fwd")
        ctx.gi = gi
        ctx.save_for_backward(x, out, stats)
        return out

    @staticmethod
    def backward(ctx, g):
        x, out, stats = ctx.saved_tensors
        gi = ctx.gi
        g = g.contiguous()
        gx = torch.empty_like(x)
        _lib.check(_lib.load().ax2d_charge_eq_bwd(_p(x), x.stride(0), _p(out), out.stride(0), _p(gi.seg_ptr),
                                                  gi.num_graphs, x.shape[1], _p(stats), _p(g), g.stride(0), _p(gx),
                                                  gx.stride(0), _stream()), "ax2d_charge_eq_bwd")
        return gx, None, None


class TetraFn(torch.autograd.Function):
    """GNN._tetrahedral_feature_calculation_physics_inspired (gnn.py:387-462), quirk Q4."""

    @staticmethod
    def forward(ctx, x, true_width, gi):
        idx, slot_ptr, slot_idx, M = gi.tetra
        out = torch.empty_like(x)
        contrib = torch.empty((4 * M, x.shape[1]), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().ax2d_tetra_fwd(_p(x), x.stride(0), x.shape[0], x.shape[1], true_width, _p(idx), M,
                                              _p(slot_ptr), _p(slot_idx), _p(contrib), _p(out), out.stride(0),
                                              _stream()), "ax2d_tetra_fwd")
        ctx.gi, ctx.true_width = gi, true_width
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        idx, slot_ptr, slot_idx, M = ctx.gi.tetra
        g = g.contiguous()
        gx = torch.empty_like(x)
        rows = torch.empty((4 * M, x.shape[1]), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().ax2d_tetra_bwd(_p(x), x.stride(0), x.shape[0], x.shape[1], ctx.true_width, _p(idx), M,
                                              _p(slot_ptr), _p(slot_idx), _p(g), g.stride(0), _p(rows), _p(gx),
                                              gx.stride(0), _stream()), "ax2d_tetra_bwd")
        return gx, None, None


class CisTransFn(torch.autograd.Function):
    """GNN._cis_trans_calculation (gnn.py:465-509), quirk Q3."""

    @staticmethod
    def forward(ctx, x, gi):
        ctx.gi = gi
        return CisTransFn._run(x, gi, 0)

    @staticmethod
    def _run(x, gi, transpose):
        src, tgt, sign, n = gi.cistrans
        out = torch.empty_like(x)
        _lib.check(_lib.load().ax2d_cistrans(_p(x), x.stride(0), x.shape[0], x.shape[1], _p(src), _p(tgt), _p(sign), n,
                                             transpose, _p(out), out.stride(0), _stream()), "ax2d_cistrans")
        return out

    @staticmethod
    def backward(ctx, g):
        return CisTransFn._run(g.contiguous(), ctx.gi, 1), None


class WeightedLossFn(torch.autograd.Function):
    """WeightedL1Loss / WeightedMSELoss (losses.py:14-87) forward + gradient in one launch."""

    @staticmethod
    def forward(ctx, pred, target, weights, kind):
        _need_cuda(pred, target, weights)
        pred, target = pred.contiguous(), target.contiguous()
        B, T = pred.shape
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        gp = torch.empty_like(pred)
        _lib.check(_lib.load().ax2d_weighted_loss(_p(pred), _p(target), _p(weights.contiguous()), B, T, kind, _p(loss),
                                                  _p(gp), _stream()), "ax2d_weighted_loss")
        ctx.save_for_backward(gp)
        return loss

    @staticmethod
    def backward(ctx, g):
        (gp,) = ctx.saved_tensors
        return gp * g, None, None, None
