"""Stand-alone timing of the lin_edge0|1 GEMM ([Mp, 64] x [512, 64]^T -> bf16) with each epilogue activation."""
import ctypes, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffspectra_b200 import _lib as L
h = ctypes.c_void_p()
L.check(L.lib().ds_create(ctypes.byref(h), 0, L.MODE_BF16, 3), 'ds_create')
M, N, K = 162305, 512, 64
g = torch.Generator(device='cuda').manual_seed(0)
A = torch.randn(M, K, device='cuda', generator=g).bfloat16()
W = (torch.randn(N, K, device='cuda', generator=g) / 8).bfloat16()
out = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def run(act):
    L.check(L.lib().ds_gemm(h, 1, L.ptr(A), K, L.ptr(W), K, None, None, 0, L.ptr(out), N, M, N, K, L.DT_BF16, L.DT_BF16, act, L.stream_ptr()), 'ds_gemm')
for name, act in (('none', 0), ('silu', 1), ('tanh', 2), ('tanh_mix', 5), ('gelu', 3)):
    for _ in range(3): run(act)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(act); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    print('act %-8s min %.1f us median %.1f us' % (name, min(ts), sorted(ts)[5]))
