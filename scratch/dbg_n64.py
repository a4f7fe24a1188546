import sys, torch
sys.path.insert(0, '.')
from oracle import dense_oracle as O
from oracle import weights as W
from tests.helpers import build_model, rel_l2
from tests.test_scale_gpu import _inputs
version = 'allspectra'
for nlist, N in (([64, 40, 33, 64, 1], 64), ([29, 20, 33, 30, 1], 33), ([40, 40], 40), ([64, 64], 64)):
    for precision in ('fp32', 'bf16'):
        model = build_model(version, salt=6, coord_scale=0.01, precision=precision)
        sd = {k: (v.double() if v.is_floating_point() else v) for k, v in model.state_dict().items()}
        n = torch.tensor(nlist)
        nm, em, x, ex, cx, cex, nl = _inputs(n, N, seed=31)
        ctx = [t.cuda() for t in W.synthetic_spectra(len(n), version, seed=5)]
        with torch.no_grad():
            pred, epred = model(nl, x, nm, em, context=ctx, edge_x=ex, noise_level=nl, cond_x=cx, cond_edge_x=cex)
            cemb = O.context_embedding(sd, [c.double() for c in ctx], version)
            ref, eref = O.dmt_forward(sd, x.double(), nm.double(), em.double(), ex.double(), nl.double(), cx.double(), cex.double(), cemb)
        print(nlist, precision, 'pos %.2e atom %.2e edge %.2e' % (rel_l2(pred[..., :3], ref[..., :3]), rel_l2(pred[..., 3:], ref[..., 3:]), rel_l2(epred, eref)),
              'per-mol edge', ['%.1e' % rel_l2(epred[b], eref[b]) if n[b] > 1 else '-' for b in range(len(n))],
              'ch', ['%.1e' % rel_l2(epred[..., c], eref[..., c]) for c in range(2)])
        if precision == 'bf16':
            d = (epred.double() - eref).abs()
            b = 0
            idx = d[b].flatten().topk(5).indices
            print('   worst entries mol0:', [(int(i) // (N * 2), (int(i) // 2) % N, int(i) % 2, float(epred[b].flatten()[i]), float(eref[b].flatten()[i])) for i in idx])
