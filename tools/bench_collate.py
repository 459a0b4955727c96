"""Host-side row a9: collation of a training batch, timed on the CPU beside the reference's own
``MyBatch.from_data_list`` (datasets/molecular.py:339-458, imported UNMODIFIED from /root/reference with the same stubs
for the absent rdkit / h5py / torch_geometric as tests/golden/make_golden.py -- so this tool runs only where
/root/reference exists; its output is kept in profiles/).

ours = ``MolBatch.from_data_list``: the same Batch fields (bit-identical, tests/test_host_abi.py) PLUS everything the
CUDA path needs so that nothing is rebuilt per step: forward / transposed CSR (stable counting sort, C++), the tile
plan of the aggregation kernel, the embedding orders and the stereo slot CSR.
"""
import importlib.util
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from aimnet_x2d_b200 import synthetic as S  # noqa: E402
from aimnet_x2d_b200.collate import MolBatch  # noqa: E402
from oracle.fixtures import FEATURE_SIZES  # noqa: E402

REF = "/root/reference/src"


def load_ref_molecular():
    for name in ("rdkit", "rdkit.Chem", "rdkit.Chem.rdBase", "rdkit.Chem.rdchem", "h5py", "torch_geometric", "torch_geometric.data"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["rdkit"].Chem = sys.modules["rdkit.Chem"]
    sys.modules["rdkit.Chem"].rdBase = sys.modules["rdkit.Chem.rdBase"]
    sys.modules["rdkit.Chem"].rdchem = sys.modules["rdkit.Chem.rdchem"]
    sys.modules["rdkit.Chem.rdchem"].HybridizationType = types.SimpleNamespace(S=0, SP=1, SP2=2, SP3=3, SP3D=4, SP3D2=5)
    tgd = sys.modules["torch_geometric.data"]

    class Bag:
        def __init__(self, *a, **k):
            pass
    tgd.Data, tgd.Batch, tgd.InMemoryDataset = type("Data", (Bag,), {}), type("Batch", (Bag,), {}), type("IMD", (Bag,), {})
    pkg = types.ModuleType("refds")
    pkg.__path__ = [os.path.join(REF, "datasets")]
    sys.modules["refds"] = pkg
    mods = {}
    for n in ("constants", "features", "molecular"):
        spec = importlib.util.spec_from_file_location(f"refds.{n}", os.path.join(REF, "datasets", f"{n}.py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"refds.{n}"] = m
        spec.loader.exec_module(m)
        mods[n] = m
    return mods["molecular"], tgd


def best_of(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, float(np.median(ts)) * 1e3, out


def main():
    torch.set_num_threads(1)                      # a DataLoader worker
    have_ref = os.path.isdir(REF)
    mol_ref, tgd = load_ref_molecular() if have_ref else (None, None)
    print(f"# collation of one training batch on ONE host core (a DataLoader worker); ms = best / median of 5")
    for label, B, hops, kind, stereo in (("C2: 2048 QM9-shaped molecules, 3 hops", 2048, 3, "qm9", False),
                                         ("c3: 1024 drug-like molecules, 3 hops, stereo", 1024, 3, "druglike", True),
                                         ("c4 per-GPU share at 8 GPUs: 1024 drug-like molecules, 4 hops", 1024, 4, "druglike", False)):
        mols = S.make_molecules(5, B, hops, kind, 12, stereo)
        ours_list = [S.to_data(m) for m in mols]
        n_atoms = sum(m["num_atoms"] for m in mols)
        n_edges = sum(h.shape[1] for m in mols for h in m["hops"])
        t_ours = best_of(lambda: MolBatch.from_data_list(ours_list, FEATURE_SIZES), 5)
        line = f"{label}: {n_atoms} atoms, {n_edges} shell edges\n    ours (Batch fields + CSR pair + tile plan + embedding orders): {t_ours[0]:8.1f} / {t_ours[1]:8.1f} ms"
        if have_ref:
            ref_list = []
            for d in ours_list:
                r = tgd.Data()
                r.__dict__.update(d.__dict__)
                ref_list.append(r)
            t_ref = best_of(lambda: mol_ref.MyBatch.from_data_list(ref_list), 5)
            rb, ob = t_ref[2], t_ours[2]
            same = (torch.equal(rb.multi_hop_edge_indices, ob.multi_hop_edge_indices) and torch.equal(rb.batch_indices, ob.batch_indices)
                    and torch.equal(rb.final_tetrahedral_chiral_tensor.reshape(-1, 4), ob.final_tetrahedral_chiral_tensor.reshape(-1, 4))
                    and torch.equal(rb.targets, ob.targets))
            line += (f"\n    reference MyBatch.from_data_list (Batch fields only):          {t_ref[0]:8.1f} / {t_ref[1]:8.1f} ms"
                     f"   -> x {t_ref[0] / t_ours[0]:.2f}; edges / batch index / tetra / targets identical: {same}")
        print(line, flush=True)


def shard_read():
    """f2 beside a9: the same C2 batches pre-collated once into a shard file and read back record by record."""
    import tempfile
    import aimnet_x2d_b200 as ax
    from aimnet_x2d_b200.collate import pad_batch
    sys.path.insert(0, ROOT)
    import bench
    raw = [S.make_batch(s, 2048, 3, "qm9") for s in range(6)]
    t0 = time.perf_counter()
    padded = bench.pad_ring(raw)
    t_pad = (time.perf_counter() - t0) * 1e3 / len(raw) / 2          # pad_ring pads every batch twice
    path = os.path.join(tempfile.mkdtemp(), "c2.ax2d")
    t0 = time.perf_counter()
    ax.write_shard(path, padded)
    t_write = (time.perf_counter() - t0) * 1e3 / len(padded)
    ds = ax.ShardDataset(path, 0, 1, shuffle=False)
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        n = sum(1 for _ in ds)
        best = min(best, (time.perf_counter() - t0) * 1e3 / n)
    mb = os.path.getsize(path) / 1e6 / len(padded)
    print(f"C2 batch as a pre-collated shard record ({mb:.1f} MB): pad_batch {t_pad:.1f} ms, write {t_write:.1f} ms (once), "
          f"read back into a HostBatch {best:.2f} ms = {mb / best:.1f} GB/s (mmap + one memcpy; no per-molecule work)")


if __name__ == "__main__":
    main()
    shard_read()
