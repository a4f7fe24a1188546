"""Parity of the CUDA DMT_WO_EQ path (non-equivariant ablation, BASELINE configs[3]) against golden outputs of the
UNMODIFIED reference models/dmt_wo_eq.py and against the oracle restatement, through DMT_WO_EQ_B200.forward -> C-ABI.
Tolerances as for DMT: teacher-forced rel-L2 <= 1e-5 (fp32 validation mode) / <= 2e-2 (bf16 mode)."""
import pytest
import torch

from oracle import dense_oracle as O
from oracle import weights as W
from tests.helpers import keyed_state_dict, load_golden, rel_l2

pytestmark = pytest.mark.gpu


def _model(version, salt, precision):
    from diffspectra_b200.config import get_config
    from diffspectra_b200.model import DMT_WO_EQ_B200
    m = DMT_WO_EQ_B200(get_config(version, device='cuda', precision=precision)).eval()
    keyed_state_dict(m, salt)
    return m.to('cuda')


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-5), ('bf16', 2e-2)])
@pytest.mark.parametrize('case', ['step0', 'selfcond'])
def test_wo_eq_matches_reference_golden(case, precision, tol):
    g = load_golden('denoiser_wo_eq_allspectra.pt')
    model = _model(g['version'], g['salt'], precision)
    n = g['n_atoms']
    nm, em = [t.cuda() for t in W.make_masks(n, g['N'])]
    ctx = W.synthetic_spectra(len(n), g['version'], seed=g['ctx_seed'])
    c = g['cases'][case]
    cu = lambda t: None if t is None else t.cuda()
    with torch.no_grad():
        pred, epred = model(c['noise_level'].cuda(), c['x'].cuda(), nm, em, context=ctx, edge_x=c['edge_x'].cuda(),
                            noise_level=c['noise_level'].cuda(), cond_x=cu(c['cond_x']), cond_edge_x=cu(c['cond_edge_x']))
    e_pos, e_atom, e_edge = rel_l2(pred[..., :3], c['pred'][..., :3]), rel_l2(pred[..., 3:], c['pred'][..., 3:]), rel_l2(epred, c['edge_pred'])
    print('wo_eq', case, precision, 'pos %.2e atom %.2e edge %.2e' % (e_pos, e_atom, e_edge))
    assert torch.isfinite(pred).all() and torch.isfinite(epred).all()
    assert max(e_pos, e_atom, e_edge) <= tol
    nmc, emc = W.make_masks(n, g['N'])
    assert (pred.cpu() * (1 - nmc)).abs().max() == 0
    assert (epred.cpu() * (1 - emc.reshape(len(n), g['N'], g['N'], 1))).abs().max() == 0
    assert (epred - epred.transpose(1, 2)).abs().max() == 0


def test_wo_eq_sampling_trajectory_matches_reference():
    """10 free-running ancestral steps with the reference's torch noise stream, fp32 mode, vs the reference sampler."""
    from diffspectra_b200.noise_schedule import NoiseScheduleVP
    from diffspectra_b200.sampling import AncestralSampler
    g = load_golden('denoiser_wo_eq_allspectra.pt')
    t = g['traj']
    model = _model(g['version'], g['salt'], 'fp32')
    n, N = g['n_atoms'], g['N']
    B = len(n)
    nm, em = W.make_masks(n, N)
    ctx = W.synthetic_spectra(B, g['version'], seed=g['ctx_seed'])
    ns = NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
    sampler = AncestralSampler(ns, torch.linspace(ns.T, 1e-3, t['steps']), True, True, True, None, 1.0, noise='torch')
    torch.manual_seed(t['seed'])
    z = O.node_noise_from_raw(torch.randn(B, N, 3), torch.randn(B, N, 6), nm)
    ez = O.edge_noise_from_raw(torch.randn(B, 2, N, N), em)
    raw = [O.draw_step_noise(B, N, nm, em) for _ in range(t['steps'])]
    eng = model.engine('cuda')
    with torch.no_grad():
        x_mean, e_mean = eng.sample_loop(model.plan_for(nm.cuda()), model.context_embedding(ctx), sampler.coefficients(), z, ez,
                                         tuple(torch.stack([r[i] for r in raw]) for i in range(3)))
    ex, ee = rel_l2(x_mean, t['x_mean']), rel_l2(e_mean, t['edge_x_mean'])
    print('wo_eq trajectory fp32: x %.2e edge %.2e' % (ex, ee))
    assert ex < 1e-4 and ee < 1e-4
    # the sampler entry point itself (sampling.py:565) accepts the ablation model, bare or DataParallel-wrapped
    ps = AncestralSampler(ns, torch.linspace(ns.T, 1e-3, 5), True, True, True, None, 1.0, noise='philox', seed=3)
    with torch.no_grad():
        a = ps.sampling(model, None, nm.cuda(), em.cuda(), None, ctx)
        b = ps.sampling(torch.nn.DataParallel(model), None, nm.cuda(), em.cuda(), None, ctx)
    assert torch.isfinite(a[0]).all() and torch.isfinite(a[1]).all()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_wo_eq_n64_and_large_batch():
    """N=64 padded molecules + a 300-molecule batch (many tiles per CTA) against the oracle in fp64 / batch invariance."""
    version = 'ir'
    for precision, tol in (('fp32', 1e-5), ('bf16', 2e-2)):
        model = _model(version, 9, precision)
        sd = {k: (v.double() if v.is_floating_point() else v) for k, v in model.state_dict().items()}
        n = torch.tensor([64, 3, 40, 1])
        B, N = len(n), 64
        nm, em = [t.cuda() for t in W.make_masks(n, N)]
        g = torch.Generator().manual_seed(5)
        x = O.node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g), nm.cpu()).cuda()
        ex = O.edge_noise_from_raw(torch.randn(B, 2, N, N, generator=g), em.cpu()).cuda()
        cx = O.node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g) * 0.2, nm.cpu()).cuda()
        cex = O.edge_noise_from_raw(torch.randn(B, 2, N, N, generator=g) * 0.4, em.cpu()).cuda()
        nl = torch.linspace(-5, 5, B).cuda()
        ctx = W.synthetic_spectra(B, version, seed=4).cuda()
        with torch.no_grad():
            pred, epred = model(nl, x, nm, em, context=ctx, edge_x=ex, noise_level=nl, cond_x=cx, cond_edge_x=cex)
            cemb = O.context_embedding(sd, ctx.double(), version)
            ref, eref = O.dmt_wo_eq_forward(sd, x.double(), nm.double(), em.double(), ex.double(), nl.double(), cx.double(),
                                            cex.double(), cemb)
        errs = (rel_l2(pred[..., :3], ref[..., :3]), rel_l2(pred[..., 3:], ref[..., 3:]), rel_l2(epred, eref))
        print('wo_eq N=64', precision, 'pos %.2e atom %.2e edge %.2e' % errs)
        assert torch.isfinite(pred).all() and torch.isfinite(epred).all()
        assert max(errs) <= tol
