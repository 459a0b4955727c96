"""Development tool: how far from the 1e-5 bar the full-size training-step comparison sits, per parameter gradient
(max |cuda - oracle| / max |oracle| and the norm difference), for the C2 and C3 configurations of tests/test_gpu_parity.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import test_gpu_parity as T  # noqa: E402
from aimnet_x2d_b200 import synthetic as S  # noqa: E402

CASES = [("c2", dict(hidden_dim=512, num_shells=3, num_message_passing_layers=3), 11,
          lambda: S.make_batch(1234 + 2000 + 7, 2048, 3, "qm9", 12), torch.linspace(0.5, 1.5, 12)),
         ("c3", dict(hidden_dim=512, num_shells=3, num_message_passing_layers=3, use_partial_charges=True,
                     use_stereochemistry=True), 13,
          lambda: S.make_batch(1234 + 3000 + 7, 256, 3, "drug", 12, stereo=True), torch.ones(12))]
for name, cfg, seed, mk, w in CASES:
    cuda, oracle = T._train_step_pair(cfg, 12, seed, mk(), w)
    rows = []
    for k, ref in oracle["grads"].items():
        got = cuda["grads"][k]
        scale = float(np.max(np.abs(ref))) or 1e-30
        rel = float(np.max(np.abs(got - ref))) / scale
        rn, gn = float(np.linalg.norm(ref.astype(np.float64))), float(np.linalg.norm(got.astype(np.float64)))
        rows.append((rel, abs(gn - rn) / max(rn, 1e-30), k, T._f64_criterion_allowed(k) or T.is_softmax_bias(k)))
    rows.sort(reverse=True)
    print(f"== {name}: loss rel {abs(cuda['loss'] - oracle['loss']) / abs(oracle['loss']):.2e}, total norm rel "
          f"{abs(cuda['norm'] - oracle['norm']) / oracle['norm']:.2e}, output rel "
          f"{float(np.max(np.abs(cuda['out'] - oracle['out']))) / float(np.max(np.abs(oracle['out']))):.2e}")
    print("   worst entries NOT on the allow-list (max|diff|/max|ref|, |norm diff|/norm):")
    for rel, nr, k, allowed in [r for r in rows if not r[3]][:8]:
        print(f"     {rel:.2e}  {nr:.2e}  {k}")
    print("   worst allow-listed entries:")
    for rel, nr, k, allowed in [r for r in rows if r[3]][:5]:
        print(f"     {rel:.2e}  {nr:.2e}  {k}")
