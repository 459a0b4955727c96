"""Development tool: systematic (signed) error of the 3xTF32 tensor-core projections against float64 -- the tensor core
accumulates in fp32 with truncation, which shrinks every accumulator by ~half an ulp per MMA."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aimnet_x2d_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
for (M, widths, N, pos) in [(37632, [160], 160, False), (37632, [160, 160], 320, False), (37632, [256], 544, False),
                            (37632, [384, 160], 512, False), (2048, [512], 512, False), (2048, [512, 512], 32, False),
                            (37632, [160], 160, True), (2048, [512], 512, True)]:
    K = sum(widths)
    a = [torch.randn(M, w, device=dev) for w in widths]
    if pos:
        a = [t.abs() for t in a]
    W = torch.randn(N, K, device=dev) / K ** 0.5
    if pos:
        W = W.abs()
    out = torch.empty(M, N, device=dev)
    ops.gemm(list(zip(a, widths)), [(W, K)], [(out, N)], M, N, K)
    ref = torch.cat([t.double() for t in a], 1) @ W.double().t()
    big = ref.abs() > 0.5 * ref.abs().mean()
    rel = ((out.double() - ref) / ref)[big]
    simt = torch.empty(M, N, device=dev)
    ops.USE_TENSOR_CORES = False
    ops.gemm(list(zip(a, widths)), [(W, K)], [(simt, N)], M, N, K)
    ops.USE_TENSOR_CORES = True
    rel_s = ((simt.double() - ref) / ref)[big]
    print(f"M={M:6d} K={K:4d} N={N:4d} pos={int(pos)}: TC mean signed rel err {float(rel.mean()):+.3e} (ulp24 units {float(rel.mean()) / 2 ** -24:+.2f}), "
          f"rms {float(rel.pow(2).mean().sqrt()):.2e} | SIMT mean {float(rel_s.mean()):+.3e} rms {float(rel_s.pow(2).mean().sqrt()):.2e}", flush=True)
