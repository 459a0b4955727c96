import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import load_golden
import test_gpu_bf16 as TB
from aimnet_x2d_b200 import ops
name = sys.argv[1] if len(sys.argv) > 1 else "gnn_h4_l3"
g = load_golden(name)
def grads(bf, defer=True):
    ops.DEFER_SPLITK_REDUCE = defer
    model, _, _ = TB._model(g)
    model.compute_dtype = torch.bfloat16 if bf else torch.float32
    o, a, q, l = TB._run(model, g)
    l.backward()
    torch.cuda.synchronize()
    return {k: (p.grad.double().cpu() if p.grad is not None else torch.zeros(p.shape, dtype=torch.double)) for k, p in model.named_parameters()}, o.detach().double().cpu()
g32, o32 = grads(False)
for defer in (True, False):
    g16, o16 = grads(True, defer)
    print("defer", defer, "out err", float((o16 - o32).abs().max() / o32.abs().max()))
    for k in reversed(list(g32)):
        s = float(g32[k].abs().max())
        if s == 0: continue
        e = float((g16[k] - g32[k]).abs().max()) / s
        if e > 0.05: print(f"   {k:60s} {e:.3f}  shape {tuple(g32[k].shape)}")
