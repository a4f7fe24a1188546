"""Parity of the CUDA denoiser (through the reference-facing DMT_B200.forward -> C-ABI ds_denoise) against
(a) golden outputs of the UNMODIFIED reference and (b) the oracle restatement run on the same inputs.

Tolerances (BASELINE.json north_star / SURVEY.md §8(d)): teacher-forced per-step outputs, rel-L2 per tensor
<= 1e-5 in fp32 validation mode and <= 2e-2 in bf16 mode."""
import pytest
import torch

from oracle import dense_oracle as O
from oracle import weights as W
from tests.helpers import build_model, load_golden, max_abs, rel_l2

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 2e-2


def _run_case(model, g, case, dev='cuda'):
    n = g['n_atoms']
    nm, em = W.make_masks(n, g['N'])
    ctx = W.synthetic_spectra(len(n), g['version'], seed=g['ctx_seed'])
    c = g['cases'][case]
    kw = dict(edge_x=c['edge_x'].to(dev), noise_level=c['noise_level'].to(dev),
              cond_x=None if c['cond_x'] is None else c['cond_x'].to(dev),
              cond_edge_x=None if c['cond_edge_x'] is None else c['cond_edge_x'].to(dev))
    ctx = [t for t in ctx] if isinstance(ctx, list) else ctx          # CPU tensors, like sampling.py:423-427
    with torch.no_grad():
        pred, epred = model(c['noise_level'].to(dev), c['x'].to(dev), nm.to(dev), em.to(dev), context=ctx, **kw)
    return pred, epred, c


@pytest.mark.parametrize('version', ['allspectra', 'ir'])
@pytest.mark.parametrize('case', ['step0', 'selfcond', 'advcond'])
def test_fp32_matches_reference_golden(version, case):
    g = load_golden('denoiser_%s.pt' % version)
    model = build_model(version, g['salt'], g['coord_scale'], 'fp32')
    pred, epred, c = _run_case(model, g, case)
    e_pos = rel_l2(pred[..., :3], c['pred'][..., :3])
    e_atom = rel_l2(pred[..., 3:], c['pred'][..., 3:])
    e_edge = rel_l2(epred, c['edge_pred'])
    print(version, case, 'pos %.2e atom %.2e edge %.2e | max %.2e %.2e' % (
        e_pos, e_atom, e_edge, max_abs(pred, c['pred']), max_abs(epred, c['edge_pred'])))
    assert torch.isfinite(pred).all() and torch.isfinite(epred).all()
    assert e_pos <= FP32_TOL and e_atom <= FP32_TOL and e_edge <= FP32_TOL
    # padded atoms / edges / diagonal are exactly zero, like the reference
    nm, em = W.make_masks(g['n_atoms'], g['N'])
    assert (pred.cpu() * (1 - nm)).abs().max() == 0
    assert (epred.cpu() * (1 - em.reshape(len(g['n_atoms']), g['N'], g['N'], 1))).abs().max() == 0
    assert (epred - epred.transpose(1, 2)).abs().max() == 0


@pytest.mark.parametrize('version', ['allspectra', 'ir'])
def test_context_embedding_matches_reference_golden(version):
    g = load_golden('denoiser_%s.pt' % version)
    ctx = W.synthetic_spectra(len(g['n_atoms']), version, seed=g['ctx_seed'])
    for prec, tol in (('fp32', FP32_TOL), ('bf16', BF16_TOL)):
        model = build_model(version, g['salt'], g['coord_scale'], prec)
        emb = model.context_embedding(ctx)
        err = rel_l2(emb, g['ctx_emb'])
        print(version, prec, 'ctx_emb rel-L2 %.2e' % err)
        assert err <= tol


@pytest.mark.parametrize('version', ['allspectra', 'ir'])
@pytest.mark.parametrize('case', ['step0', 'selfcond', 'advcond'])
def test_bf16_matches_reference_golden(version, case):
    g = load_golden('denoiser_%s.pt' % version)
    model = build_model(version, g['salt'], g['coord_scale'], 'bf16')
    pred, epred, c = _run_case(model, g, case)
    e_pos = rel_l2(pred[..., :3], c['pred'][..., :3])
    e_atom = rel_l2(pred[..., 3:], c['pred'][..., 3:])
    e_edge = rel_l2(epred, c['edge_pred'])
    print(version, case, 'bf16 pos %.2e atom %.2e edge %.2e | max %.2e %.2e' % (
        e_pos, e_atom, e_edge, max_abs(pred, c['pred']), max_abs(epred, c['edge_pred'])))
    assert torch.isfinite(pred).all() and torch.isfinite(epred).all()
    assert e_pos <= BF16_TOL and e_atom <= BF16_TOL and e_edge <= BF16_TOL


def test_fp32_matches_oracle_random_batch():
    """Fresh seeded inputs (not in the golden set), ragged batch incl. n=1 and n=2, vs the oracle on the GPU in fp64."""
    version = 'allspectra'
    model = build_model(version, salt=3, coord_scale=0.03, precision='fp32')
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in model.state_dict().items()}
    n = torch.tensor([1, 2, 29, 5, 18, 18, 7, 26])
    B, N = len(n), 29
    nm, em = [t.cuda() for t in W.make_masks(n, N)]
    g = torch.Generator().manual_seed(11)
    x = O.node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g), nm.cpu()).cuda()
    ex = O.edge_noise_from_raw(torch.randn(B, 2, N, N, generator=g), em.cpu()).cuda()
    cx = O.node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g) * 0.2, nm.cpu()).cuda()
    cex = O.edge_noise_from_raw(torch.randn(B, 2, N, N, generator=g) * 0.4, em.cpu()).cuda()
    nl = torch.linspace(-8, 8, B).cuda()
    ctx = [t.cuda() for t in W.synthetic_spectra(B, version, seed=9)]
    with torch.no_grad():
        pred, epred = model(nl, x, nm, em, context=ctx, edge_x=ex, noise_level=nl, cond_x=cx, cond_edge_x=cex)
        cemb = O.context_embedding(sd, [c.double() for c in ctx], version)
        ref, eref = O.dmt_forward(sd, x.double(), nm.double(), em.double(), ex.double(), nl.double(), cx.double(),
                                  cex.double(), cemb)
    print('vs fp64 oracle: pos %.2e atom %.2e edge %.2e' % (rel_l2(pred[..., :3], ref[..., :3]),
                                                            rel_l2(pred[..., 3:], ref[..., 3:]), rel_l2(epred, eref)))
    assert rel_l2(pred[..., :3], ref[..., :3]) <= FP32_TOL
    assert rel_l2(pred[..., 3:], ref[..., 3:]) <= FP32_TOL
    assert rel_l2(epred, eref) <= FP32_TOL


def test_inputs_not_mutated_and_weights_repack():
    g = load_golden('denoiser_ir.pt')
    model = build_model('ir', g['salt'], g['coord_scale'], 'fp32')
    c = g['cases']['selfcond']
    x = c['x'].cuda()
    x0 = x.clone()
    pred, _, _ = _run_case(model, g, 'selfcond')
    assert torch.equal(x, x0)
    # parameters change (e.g. ema.copy_to) -> packed weights are rebuilt and the output changes
    with torch.no_grad():
        model.node_pred_mlp[4].bias.add_(1.0)
    pred2, _, _ = _run_case(model, g, 'selfcond')
    nm, _ = W.make_masks(g['n_atoms'], g['N'])
    d = (pred2 - pred)[..., 3:].cpu()
    assert torch.allclose(d, torch.ones_like(d) * nm, atol=1e-5)


def test_weights_changed_through_data_copy_are_repacked():
    """run_lib.py:359-362 order is restore_checkpoint -> ema.copy_to -> sampling_fn; ema.copy_to / ema.restore write
    through `param.data.copy_` (models/ema.py:52-55), invisible to version counters.  A model that has already run
    (packed weights cached) must pick the new weights up at the next sampling round (= next new context)."""
    g = load_golden('denoiser_ir.pt')
    model = build_model('ir', g['salt'] + 11, g['coord_scale'], 'fp32')       # wrong weights first
    _run_case(model, g, 'step0')
    right = build_model('ir', g['salt'], g['coord_scale'], 'fp32', device='cpu').state_dict()
    key = model._params_key()
    with torch.no_grad():
        for (n, p) in list(model.named_parameters()) + list(model.named_buffers()):   # what ema.copy_to does
            p.data.copy_(right[n].data)
    assert model._params_key() == key                                        # the cheap key is blind to it
    pred, epred, c = _run_case(model, g, 'step0')                            # fresh context tensors = new round
    assert rel_l2(pred, c['pred']) <= FP32_TOL and rel_l2(epred, c['edge_pred']) <= FP32_TOL
