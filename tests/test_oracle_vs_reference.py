"""When the reference tree is present (build container only) cross-check the oracle against the live, unmodified
reference modules on fresh inputs — a second pin besides the committed goldens.  Skipped on the GPU box."""
import pytest
import torch

from oracle import dense_oracle as O
from oracle import weights as W
from oracle.ref_harness import load_reference, reference_available
from tests.helpers import rel_l2

pytestmark = pytest.mark.skipif(not reference_available(), reason='/root/reference not present')


def test_oracle_matches_live_reference_ir():
    R = load_reference()
    R.config.data.spectra_version = 'ir'
    m = R.DMT(R.config).eval()
    sd = m.state_dict()
    W.keyed_fill_(sd, salt=5, coord_scale=0.1)
    m.load_state_dict(sd)
    n = torch.tensor([2, 29, 8])
    B, N = 3, 29
    nm, em = W.make_masks(n, N)
    ctx = W.synthetic_spectra(B, 'ir', seed=3)
    g = torch.Generator().manual_seed(0)
    x = O.node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g), nm)
    ex = O.edge_noise_from_raw(torch.randn(B, 2, N, N, generator=g), em)
    cx = O.node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g), nm)
    cex = O.edge_noise_from_raw(torch.randn(B, 2, N, N, generator=g), em)
    nl = torch.tensor([-4., 0.5, 7.])
    with torch.no_grad():
        p, e = m(nl, x, nm, em, edge_x=ex, noise_level=nl, cond_x=cx, cond_edge_x=cex, context=ctx)
        q, f = O.dmt_forward(sd, x, nm, em, ex, nl, cx, cex, O.context_embedding(sd, ctx, 'ir'))
    assert rel_l2(q, p) < 3e-6 and rel_l2(f, e) < 3e-6


@pytest.mark.parametrize('with_cond', [True, False])
def test_oracle_matches_live_reference_dmt_wo_eq(with_cond):
    R = load_reference()
    R.config.data.spectra_version = 'ir'
    m = R.DMT_WO_EQ(R.config).eval()
    sd = m.state_dict()
    W.keyed_fill_(sd, salt=7)
    m.load_state_dict(sd)
    n = torch.tensor([2, 19, 8, 1])
    B, N = 4, 19
    nm, em = W.make_masks(n, N)
    ctx = W.synthetic_spectra(B, 'ir', seed=3)
    g = torch.Generator().manual_seed(1)
    x = O.node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g), nm)
    ex = O.edge_noise_from_raw(torch.randn(B, 2, N, N, generator=g), em)
    cx = O.node_noise_from_raw(torch.randn(B, N, 3, generator=g), torch.randn(B, N, 6, generator=g), nm) if with_cond else None
    cex = O.edge_noise_from_raw(torch.randn(B, 2, N, N, generator=g), em) if with_cond else None
    nl = torch.tensor([-4., 0.5, 7., 2.])
    with torch.no_grad():
        p, e = m(nl, x, nm, em, edge_x=ex, noise_level=nl, cond_x=cx, cond_edge_x=cex, context=ctx)
        q, f = O.dmt_wo_eq_forward(sd, x, nm, em, ex, nl, cx, cex, O.context_embedding(sd, ctx, 'ir'))
    assert rel_l2(q, p) < 3e-6 and rel_l2(f, e) < 3e-6


def test_our_schedule_class_matches_reference_class():
    R = load_reference()
    from diffspectra_b200.noise_schedule import NoiseScheduleVP, ancestral_coefficients
    ours, ref = NoiseScheduleVP('cosine', continuous_beta_0=0.1, continuous_beta_1=20.), R.NoiseScheduleVP(
        'cosine', continuous_beta_0=0.1, continuous_beta_1=20.)
    assert ours.T == ref.T
    t = torch.linspace(ours.T, 1e-3, 200)
    for a, b in zip(ours.marginal_prob(t), ref.marginal_prob(t)):
        assert torch.equal(a, b)
    assert torch.equal(ancestral_coefficients(ours, t), ancestral_coefficients(ref, t))
    assert torch.equal(ancestral_coefficients(ours, t), O.schedule_table(200))
