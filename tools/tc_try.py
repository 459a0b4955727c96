"""Development aid: run one tensor-core projection shape and check it against torch (fp64)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aimnet_x2d_b200 import ops  # noqa: E402

m, k, n = (int(v) for v in sys.argv[1:4])
torch.manual_seed(0)
a = torch.randn(m, k, device="cuda")
w = torch.randn(n, k, device="cuda") / k ** 0.5
out = torch.empty(m, n, device="cuda")
ops.TC_MIN_ROWS = 1
ops.gemm([(a, k)], [(w, k)], [(out, n)], m, n, k)
torch.cuda.synchronize()
ref = (a.double() @ w.double().t())
print(m, k, n, "max err", float((out.double() - ref).abs().max()), flush=True)
