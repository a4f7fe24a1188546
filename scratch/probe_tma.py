import ctypes, torch, sys
sys.path.insert(0,'/root/repo')
from diffspectra_b200 import _lib as L
h = ctypes.c_void_p(); L.check(L.lib().ds_create(ctypes.byref(h), 0, 1, 3), 'create')
for N in (500, 504, 508, 510, 72, 100):
    M, K = 300, 64
    A = torch.randn(M, K, device='cuda').bfloat16(); W = torch.randn(N, K, device='cuda').bfloat16()
    ld = ((N + 63)//64)*64
    buf = torch.full((M+40, ld), -7.0, device='cuda', dtype=torch.bfloat16)
    out = buf[:M]
    L.check(L.lib().ds_gemm(h, 1, L.ptr(A), K, L.ptr(W), K, None, None, 0, L.ptr(out), ld, M, N, K, 1, 1, 0, L.stream_ptr()), 'gemm')
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    bad_cols = (buf[:M, N:] != -7).any(0).nonzero().flatten().tolist()
    bad_rows = (buf[M:] != -7).any(1).nonzero().flatten().tolist()
    print(N, 'err', (out[:, :N].float()-ref).abs().max().item(), 'pad cols touched', [N+c for c in bad_cols], 'rows beyond M touched', bad_rows[:5])
