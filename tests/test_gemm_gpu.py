"""Parity of the two GEMM kernels (tcgen05/TMA bf16 and CUDA-core fp32) against torch matmul."""
import ctypes

import pytest
import torch

from diffspectra_b200 import _lib as L

pytestmark = pytest.mark.gpu

ACTS = {L.ACT_NONE: lambda x: x, L.ACT_SILU: torch.nn.functional.silu, L.ACT_TANH: torch.tanh,
        L.ACT_GELU: torch.nn.functional.gelu, L.ACT_TANH_MIX: torch.tanh}


@pytest.fixture(scope='module')
def ctx():
    h = ctypes.c_void_p()
    L.check(L.lib().ds_create(ctypes.byref(h), 0, L.MODE_BF16, 3), 'ds_create')
    yield h
    L.lib().ds_destroy(h)


def run_gemm(ctx, tc, A, W, bias, addmat, out, act, M, N, K):
    in_dt = L.DT_BF16 if A.dtype == torch.bfloat16 else L.DT_F32
    out_dt = L.DT_BF16 if out.dtype == torch.bfloat16 else L.DT_F32
    L.check(L.lib().ds_gemm(ctx, int(tc), L.ptr(A), A.stride(0), L.ptr(W), W.stride(0), L.ptr(bias), L.ptr(addmat),
                            addmat.stride(0) if addmat is not None else 0, L.ptr(out), out.stride(0), M, N, K,
                            in_dt, out_dt, act, L.stream_ptr()), 'ds_gemm')
    torch.cuda.synchronize()


SHAPES = [  # (M, N, K) — the shapes the denoiser uses plus ragged tails
    (128, 64, 64), (300, 64, 128), (1000, 512, 64), (257, 768, 256), (513, 256, 512), (129, 16, 64),
    (4, 19584, 1024), (77, 128, 192), (2000, 256, 256), (64, 256, 768), (5, 256, 44416), (1, 1024, 1024),
    (40000, 256, 256), (333, 500, 64), (150, 32, 64),
]


@pytest.mark.parametrize('M,N,K', SHAPES)
def test_gemm_tc_matches_torch(ctx, M, N, K):
    g = torch.Generator(device='cuda').manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, device='cuda', generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device='cuda', generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device='cuda', generator=g) * 0.1
    ref = A.float() @ W.float().t() + bias
    out = torch.full((M, N), float('nan'), device='cuda')
    run_gemm(ctx, True, A, W, bias, None, out, L.ACT_NONE, M, N, K)
    err = (out - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), (M, N, K, err)


@pytest.mark.parametrize('act', [L.ACT_SILU, L.ACT_TANH, L.ACT_GELU])
@pytest.mark.parametrize('out_dtype', [torch.float32, torch.bfloat16])
def test_gemm_tc_epilogue(ctx, act, out_dtype):
    M, N, K = 777, 508, 128
    g = torch.Generator(device='cuda').manual_seed(act)
    A = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    W = (torch.randn(N, K, device='cuda', generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device='cuda', generator=g) * 0.1
    add = torch.randn(M, N, device='cuda', generator=g)
    ref = ACTS[act](A.float() @ W.float().t() + bias + add)
    buf = torch.full((M, 512), -7.0, device='cuda', dtype=out_dtype)      # ldo > N: padding must stay untouched
    run_gemm(ctx, True, A, W, bias, add, buf, act, M, N, K)
    tol = 2e-2 if out_dtype == torch.bfloat16 else 5e-3
    assert (buf[:, :N].float() - ref).abs().max().item() < tol
    assert (buf[:, N:] == -7.0).all()


@pytest.mark.parametrize('scale', [0.3, 1.0, 4.0, 30.0])
def test_gemm_tc_tanh_mix_epilogue(ctx, scale):
    """ACT_TANH_MIX (half of the column pairs through the FMA-pipe polynomial, common.cuh tanh_poly2) against torch.tanh at
    fp32 output: the polynomial's bound is 6.0e-4 (tests/test_host_logic.py pins the coefficients), MUFU.TANH's ~5e-4; `scale`
    moves the pre-activations from the linear range into deep saturation (|x| >> 3.75, the clamp)."""
    M, N, K = 1500, 512, 64
    g = torch.Generator(device='cuda').manual_seed(11)
    A = (torch.randn(M, K, device='cuda', generator=g) * scale).bfloat16()
    W = (torch.randn(N, K, device='cuda', generator=g) / K ** 0.5).bfloat16()
    pre = A.float() @ W.float().t()
    for out_dtype, tol in ((torch.float32, 1.2e-3), (torch.bfloat16, 6e-3)):
        out = torch.empty(M, N, device='cuda', dtype=out_dtype)
        run_gemm(ctx, True, A, W, None, None, out, L.ACT_TANH_MIX, M, N, K)
        err = (out.float() - torch.tanh(pre)).abs().max().item()
        assert err < tol, (scale, out_dtype, err)
        assert out.float().abs().max().item() <= 1.0
        # both halves of the mix are exercised and agree with the MUFU-only epilogue to the sum of the two bounds
        ref = torch.empty(M, N, device='cuda', dtype=out_dtype)
        run_gemm(ctx, True, A, W, None, None, ref, L.ACT_TANH, M, N, K)
        d = (out.float() - ref.float()).abs()
        assert d.max().item() < (1.5e-3 if out_dtype == torch.float32 else 8e-3)
        if out_dtype == torch.float32:
            cols = torch.arange(N, device='cuda')
            assert (d[:, (cols % 4) >= 2] == 0).all()          # the MUFU half is bit-identical
            assert (d[:, (cols % 4) < 2] > 0).any()            # the polynomial half is really a different evaluation


def test_gemm_tc_strided_views(ctx):
    """A and out as column slices of wider buffers (how the denoiser writes atom_hids / edge_hids)."""
    M, N, K = 1000, 16, 64
    g = torch.Generator(device='cuda').manual_seed(5)
    Xbuf = torch.randn(M, 128, device='cuda', generator=g).bfloat16()
    A = Xbuf[:, 64:]
    W = (torch.randn(N, K, device='cuda', generator=g) / 8).bfloat16()
    hid = torch.zeros(M, 192, device='cuda', dtype=torch.bfloat16)
    out = hid[:, 80:96]
    run_gemm(ctx, True, A, W, None, None, out, L.ACT_NONE, M, N, K)
    ref = A.float() @ W.float().t()
    assert (out.float() - ref).abs().max().item() < 2e-2
    assert (hid[:, :80] == 0).all() and (hid[:, 96:] == 0).all()


@pytest.mark.parametrize('M,N,K', [(100, 64, 68), (33, 1024, 17), (257, 6, 128), (500, 256, 640), (64, 128, 20)])
def test_gemm_simt_fp32(ctx, M, N, K):
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    A = torch.randn(M, K, device='cuda', generator=g)
    W = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    bias = torch.randn(N, device='cuda', generator=g)
    ref = (A.double() @ W.double().t() + bias.double()).float()
    out = torch.empty(M, N, device='cuda')
    run_gemm(ctx, False, A, W, bias, None, out, L.ACT_NONE, M, N, K)
    assert (out - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())


ADA_LD = 19584


def _fused_inputs(M, N, K, n_mol, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    A = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    W = (torch.randn(N, K, device='cuda', generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device='cuda', generator=g) * 0.1
    mol = torch.sort(torch.randint(0, n_mol, (M,), device='cuda', generator=g)).values
    info = (mol << 12).to(torch.int32)
    ada = torch.randn(n_mol, ADA_LD, device='cuda', generator=g) * 0.5
    return A, W, bias, mol, info, ada


@pytest.mark.parametrize('M', [300, 128, 5000, 40000])
def test_gemm_fused_lnmod(ctx, M):
    N, K = 64, 128
    A, W, bias, mol, info, ada = _fused_inputs(M, N, K, 7, M)
    off_a, off_b = 1536, 1600
    out = torch.full((M + 8, N), -7.0, device='cuda', dtype=torch.bfloat16)
    L.check(L.lib().ds_gemm_fused(ctx, 1, L.ptr(A), K, L.ptr(W), K, L.ptr(bias), M, N, K, L.ptr(info), 12, L.ptr(ada), off_a,
                                  off_b, None, 0, L.ptr(out), N, None, 0, None, None, None, L.stream_ptr()), 'ds_gemm_fused')
    torch.cuda.synchronize()
    y = torch.nn.functional.layer_norm(A.float() @ W.float().t() + bias, (N,), eps=1e-6)
    ref = y * (1 + ada[mol, off_b:off_b + N]) + ada[mol, off_a:off_a + N]
    assert (out[:M].float() - ref).abs().max().item() < 5e-2
    assert (out[M:] == -7).all()


@pytest.mark.parametrize('M,N,K', [(300, 64, 128), (1000, 256, 512), (40000, 64, 128)])
def test_gemm_fused_resgate(ctx, M, N, K):
    A, W, bias, mol, info, ada = _fused_inputs(M, N, K, 5, M + N)
    off = 320
    resid = torch.randn(M, N, device='cuda')
    out = torch.empty(M, N, device='cuda')
    buf2 = torch.zeros(M, 2 * N, device='cuda', dtype=torch.bfloat16)
    out2 = buf2[:, N:]
    L.check(L.lib().ds_gemm_fused(ctx, 2, L.ptr(A), K, L.ptr(W), K, L.ptr(bias), M, N, K, L.ptr(info), 12, L.ptr(ada), off, 0,
                                  L.ptr(resid), N, L.ptr(out), N, L.ptr(out2), 2 * N, None, None, None, L.stream_ptr()),
            'ds_gemm_fused')
    torch.cuda.synchronize()
    ref = resid + ada[mol, off:off + N] * (A.float() @ W.float().t() + bias)
    assert (out - ref).abs().max().item() < 5e-3
    assert (out2.float() - ref).abs().max().item() < 5e-2
    assert (buf2[:, :N] == 0).all()


@pytest.mark.parametrize('M', [3001, 40001])
def test_gemm_fused_coord(ctx, M):
    N, K = 256, 256
    A, W, bias, mol, info, ada = _fused_inputs(M, N, K, 5, 3)
    g = torch.Generator(device='cuda').manual_seed(4)
    wc2 = torch.randn(3, 256, device='cuda', generator=g) / 16
    dflags = torch.randint(0, 4, (M,), device='cuda', generator=g, dtype=torch.uint8)
    wdir = torch.full((M + 4,), -7.0, device='cuda')
    Wh, bh = (W.float() * 0.5).bfloat16(), bias * 0.5          # COORD takes the first layer pre-halved: SiLU(2 (A Wh^T + bh))
    L.check(L.lib().ds_gemm_fused(ctx, 3, L.ptr(A), K, L.ptr(Wh), K, L.ptr(bh), M, N, K, None, 0, None, 0, 0, None, 0, None, 0,
                                  None, 0, L.ptr(wc2), L.ptr(dflags), L.ptr(wdir), L.stream_ptr()), 'ds_gemm_fused')
    torch.cuda.synchronize()
    u = torch.tanh(torch.nn.functional.silu(A.float() @ W.float().t() + bias) @ wc2.t())
    adj = torch.stack([torch.ones(M, device='cuda'), (dflags & 1).float(), ((dflags >> 1) & 1).float()], dim=1)
    ref = (u * adj).mean(-1)
    assert (wdir[:M] - ref).abs().max().item() < 5e-3
    assert (wdir[M:] == -7).all()


@pytest.mark.parametrize('M,N,K,act', [(40000, 512, 64, L.ACT_TANH), (40000, 256, 128, L.ACT_NONE), (50001, 128, 64, L.ACT_SILU),
                                       (40000, 16, 64, L.ACT_NONE), (40000, 64, 256, L.ACT_NONE)])
def test_gemm_tc_many_tiles_per_cta_bf16_out(ctx, M, N, K, act):
    """Many m-tiles per SM: every persistent CTA walks several output tiles (ring + TMEM stage phases wrap)."""
    g = torch.Generator(device='cuda').manual_seed(M + N)
    A = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    W = (torch.randn(N, K, device='cuda', generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device='cuda', generator=g) * 0.1
    ref = ACTS[act](A.float() @ W.float().t() + bias)
    out = torch.full((M + 3, N), -7.0, device='cuda', dtype=torch.bfloat16)
    run_gemm(ctx, True, A, W, bias, None, out[:M], act, M, N, K)
    assert (out[:M].float() - ref).abs().max().item() < 3e-2
    assert (out[M:] == -7).all()
