"""Stand-alone timing of the fused coordinate head (ds_coord_head) at the configs[1] size with CUDA events."""
import ctypes, sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffspectra_b200 import _lib as L
from diffspectra_b200.engine import Plan
from diffspectra_b200.synthetic import sample_n_atoms

h = ctypes.c_void_p()
L.check(L.lib().ds_create(ctypes.byref(h), 0, L.MODE_BF16, 3), 'ds_create')
class E: h, device = h, torch.device('cuda')
n = sample_n_atoms(1024, seed=1234).numpy().astype(np.int32)
plan = Plan(E, n, 29)
Mn, Mp, B = plan.Mn, plan.Mp, len(n)
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(Mp, 128, device='cuda', generator=g).bfloat16()
ab = torch.randn(Mn, 512, device='cuda', generator=g).bfloat16()
we = (torch.randn(256, 128, device='cuda', generator=g) / 11).bfloat16()
wc1 = (torch.randn(256, 256, device='cuda', generator=g) / 32).bfloat16()
bc1 = torch.randn(256, device='cuda', generator=g) * 0.05
wc2 = torch.randn(3, 256, device='cuda', generator=g) / 16
ada = torch.randn(B, 19584, device='cuda', generator=g) * 0.3
pf = torch.randint(0, 4, (Mp,), device='cuda', generator=g, dtype=torch.uint8)
wdir = torch.zeros(2 * Mp, device='cuda')
scratch = torch.empty(B * 1024, dtype=torch.uint8, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def run():
    L.check(L.lib().ds_coord_head(h, *plan.args(), L.ptr(X), L.ptr(ab), L.ptr(ada), L.ptr(pf), L.ptr(we), L.ptr(wc1), L.ptr(bc1),
                                  L.ptr(wc2), L.ptr(wdir), L.ptr(scratch), L.stream_ptr()), 'ds_coord_head')
for _ in range(3): run()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1000)
dbg = torch.zeros(8 * 16, dtype=torch.int64, device='cuda')
os.environ['DS_COORD_DBG'] = str(dbg.data_ptr())
run(); torch.cuda.synchronize()
d = dbg.cpu().view(16, 8)
t0 = int(d[0, 0])
names = ['start', 'g_full', 'win_full', 'passA', 'z_free', 'passB', '-']
for it in range(11):
    print('tile %2d ' % it + ' '.join('%s %6d' % (names[k], int(d[it, k]) - t0 if int(d[it, k]) else -1) for k in range(7)))
del os.environ['DS_COORD_DBG']
print('coord_head_kernel Mn=%d Mp=%d: min %.1f us median %.1f us (L2 flushed between launches)' % (Mn, Mp, min(ts), sorted(ts)[5]))
