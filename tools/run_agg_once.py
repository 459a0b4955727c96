"""Run the aggregation kernel a few times eagerly at C2 size (target of `ncu -k regex:agg_tiles`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from aimnet_x2d_b200 import ops, synthetic as S  # noqa: E402

batch = S.make_batch(1234 + 2000, 2048, 3, "qm9")
gi = batch.graph_index.to("cuda")
dtype = torch.bfloat16 if os.environ.get("DTYPE") == "bf16" else torch.float32
x = torch.randn(gi.num_atoms, 160, device="cuda").to(dtype)
g = torch.randn(gi.num_atoms, 160, device="cuda").to(dtype)
for _ in range(4):
    ops.agg(x, gi)
    ops.agg(g, gi, transpose=True, addend=x)
torch.cuda.synchronize()
print("ok")
