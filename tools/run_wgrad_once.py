"""Development aid: a few weight-gradient launches on the shapes of the C2 step (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aimnet_x2d_b200 import ops  # noqa: E402

M = 37632
for (no, ki) in [(160, 160), (320, 320)]:
    g = torch.randn(M, no, device="cuda")
    x = torch.randn(M, ki, device="cuda")
    for _ in range(2):
        dW, db = ops._weight_grad([(g, no)], [(x, ki)], M, no, ki, "cuda", bias=True)
    torch.cuda.synchronize()
    ref = g.double().t() @ x.double()
    print(no, ki, "rel err", float((dW.double() - ref).abs().max() / ref.abs().max()))
