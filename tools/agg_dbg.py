import ctypes as C, os, sys
sys.path.insert(0,'/root/repo')
import torch
from aimnet_x2d_b200 import _lib, ops, synthetic as S
dev="cuda"
batch=S.make_batch(1234+2000,2048,3,"qm9"); gi=batch.graph_index.to(dev); N=gi.num_atoms
lib=_lib.load(); cfg=lib._lib.ax2d_debug_agg_config; cfg.argtypes=[C.c_int,C.c_int,C.c_int]; cfg.restype=None
xs=[torch.randn(N,160,device=dev) for _ in range(8)]
def t(label):
    fn=lambda i: ops.agg(xs[i%8],gi)
    for i in range(4): fn(i)
    torch.cuda.synchronize()
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(64): fn(i)
    g.replay(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    print(f"{label:40s} {a.elapsed_time(b)*1e3/64:6.1f} us")
for flags,label in ((0,"full"),(1,"no edge loop"),(2,"no stores"),(3,"neither")):
    cfg(3, flags<<8, 1); t(label)
cfg(0,0,1)
