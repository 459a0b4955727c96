"""Development tool: featurisation-time graph construction -- reference route (BFS edge lists + collation + stable sort,
host) against the direct CSR emission on the host and on the device.
    python tools/bench_shell_csr.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from aimnet_x2d_b200 import collate as CL, synthetic as S  # noqa: E402

for kind, nmol, hops in (("qm9", 2048, 3), ("drug", 8192, 4)):
    mols = S.make_molecules(5, nmol, hops, kind, 4)
    counts, bonds = [m["num_atoms"] for m in mols], [m["bonds"] for m in mols]
    t0 = time.perf_counter()
    S.shell_edges_batch(mols, hops)
    batch = CL.MolBatch.from_data_list([S.to_data(m) for m in mols], S.FEATURE_SIZES)
    t_ref = time.perf_counter() - t0
    t0 = time.perf_counter()
    h = CL.shell_csr(counts, bonds, hops)
    t_host = time.perf_counter() - t0
    CL.shell_csr(counts, bonds, hops, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d = CL.shell_csr(counts, bonds, hops, device="cuda")
    torch.cuda.synchronize()
    t_dev = time.perf_counter() - t0
    # kernels only (inputs resident)
    N = sum(counts)
    E = int(h[0][N])
    same = torch.equal(d[1].cpu()[:E], h[1][:E]) and torch.equal(d[2].cpu()[:E], h[2][:E])
    print(f"{kind:5s} {nmol} molecules x {hops} hops: N={N} E={E}  edge lists + collation (host): {t_ref * 1e3:8.1f} ms | "
          f"direct CSR host: {t_host * 1e3:8.1f} ms | direct CSR device (incl. H2D of the bonds, sync): {t_dev * 1e3:6.2f} ms | "
          f"identical {same}", flush=True)
