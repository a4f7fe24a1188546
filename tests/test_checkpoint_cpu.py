"""SURVEY.md §8(f).3 — checkpoint / EMA / pretrained-SpecFormer ingestion of the drop-in modules, checked on CPU against
the reference's own loader code (utils.restore_checkpoint, models/ema.py, DMT.load_pretrained_specformer) when
/root/reference is present, and against their documented behaviour otherwise."""
import os

import pytest
import torch

from oracle import ref_harness
from oracle import weights as W

needs_ref = pytest.mark.skipif(not ref_harness.reference_available(), reason='reference tree not present')


def _ours(kind, version, **model_kw):
    from diffspectra_b200.config import get_config
    from diffspectra_b200.model import DMT_B200, DMT_WO_EQ_B200
    cfg = get_config(version, device='cpu')
    for k, v in model_kw.items():
        setattr(cfg.model, k, v)
    return (DMT_B200 if kind == 'DMT' else DMT_WO_EQ_B200)(cfg)


def _pretrain_ckpt(path, enc_state, prefix, salt, drop=(), reshape=()):
    """A SpecFormer pre-training checkpoint as the reference expects it (models/dmt.py:276-296): a 'state_dict' whose
    encoder tensors sit under `prefix`, out_norm under 'model.representation_model', plus unrelated keys."""
    sd = {}
    for i, (k, v) in enumerate(enc_state.items()):
        g = torch.Generator().manual_seed(salt + i)
        val = torch.randn(v.shape, generator=g).to(v.dtype) if v.dtype.is_floating_point else torch.full_like(v, 7)
        if k in drop:
            continue
        if k in reshape:
            val = torch.zeros(tuple(v.shape) + (2,))
        if k.startswith('out_norm.'):
            sd['model.representation_model.' + k] = val
            sd[prefix + '.' + k] = torch.full_like(val, -5.0)        # must NOT be the one that is taken
        else:
            sd[prefix + '.' + k] = val
    sd['model.mol_encoder.weight'] = torch.ones(3)
    torch.save({'state_dict': sd, 'epoch': 3}, path)
    return sd


@needs_ref
@pytest.mark.parametrize('kind', ['DMT', 'DMT_WO_EQ'])
@pytest.mark.parametrize('prefix', ['model.representation_spec_model', 'model.representation_model'])
def test_pretrained_specformer_loader_equals_reference(tmp_path, kind, prefix):
    ref = ref_harness.load_reference()
    ours = _ours(kind, 'allspectra')
    enc = ours.cond_encoder.state_dict()
    keys = list(enc.keys())
    path = str(tmp_path / 'specformer.ckpt')
    _pretrain_ckpt(path, enc, prefix, salt=100, drop=(keys[3],), reshape=(keys[5],))

    cfg = ref.config
    cfg.data.spectra_version = 'allspectra'
    cfg.model.pretrained_specformer_path = ''
    torch.manual_seed(0)
    rm = (ref.DMT if kind == 'DMT' else ref.DMT_WO_EQ)(cfg)
    ours.load_state_dict(rm.state_dict(), strict=True)               # same starting point on both sides
    rm.load_pretrained_specformer(path)
    n = ours.load_pretrained_specformer(path)
    assert n == len(keys) - 2
    a, b = rm.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # the constructor hook takes the same path (models/dmt.py:263-267)
    ours2 = _ours(kind, 'allspectra', pretrained_specformer_path=path)
    for k, v in ours2.cond_encoder.state_dict().items():
        if k not in (keys[3], keys[5]):                              # those two keep their (random) initial values
            assert torch.equal(v, b['cond_encoder.' + k]), k


def test_pretrained_specformer_loader_ignores_foreign_checkpoints(tmp_path, capsys):
    ours = _ours('DMT', 'ir')
    before = {k: v.clone() for k, v in ours.state_dict().items()}
    p1 = str(tmp_path / 'a.ckpt')
    torch.save({'model': {}}, p1)                                    # no 'state_dict' (models/dmt.py:273-275)
    assert ours.load_pretrained_specformer(p1) == 0
    p2 = str(tmp_path / 'b.ckpt')
    torch.save({'state_dict': {'encoder.foo': torch.zeros(2)}}, p2)  # no known prefix (models/dmt.py:286-288)
    assert ours.load_pretrained_specformer(p2) == 0
    for k, v in ours.state_dict().items():
        assert torch.equal(v, before[k])
    assert 'Warning' in capsys.readouterr().out


@needs_ref
@pytest.mark.parametrize('kind', ['DMT', 'DMT_WO_EQ'])
def test_restore_checkpoint_and_ema_copy_to_through_reference_code(tmp_path, kind):
    """run_lib.diffspectra_evaluate (run_lib.py:310-313,359-362): create_model -> DataParallel, EMA over
    model.parameters(), restore_checkpoint(strict=True), ema.copy_to.  The checkpoint is written by the REFERENCE
    model through the reference's save_checkpoint and read into OUR module through the reference's restore_checkpoint;
    the drop-in must end up with the reference's EMA weights, parameter by parameter, in registration order."""
    import importlib
    ref = ref_harness.load_reference()
    ref_utils = importlib.import_module('utils')
    ema_mod = importlib.import_module('models.ema')
    cfg = ref.config
    cfg.data.spectra_version = 'allspectra'
    cfg.model.pretrained_specformer_path = ''
    torch.manual_seed(1)
    rm = torch.nn.DataParallel((ref.DMT if kind == 'DMT' else ref.DMT_WO_EQ)(cfg))
    rema = ema_mod.ExponentialMovingAverage(rm.parameters(), decay=0.999)
    with torch.no_grad():                                             # one fake optimiser step + EMA update
        for i, p in enumerate(rm.parameters()):
            p.add_(0.01 * ((i % 7) - 3))
    rema.update(rm.parameters())
    ropt = torch.optim.Adam(rm.parameters(), lr=1e-4)
    path = str(tmp_path / 'checkpoints' / 'checkpoint_40.pth')
    os.makedirs(os.path.dirname(path))
    ref_utils.save_checkpoint(path, dict(optimizer=ropt, model=rm, ema=rema, step=40))

    ours = torch.nn.DataParallel(_ours(kind, 'allspectra'))
    ema = ema_mod.ExponentialMovingAverage(ours.parameters(), decay=0.999)
    opt = torch.optim.Adam(ours.parameters(), lr=1e-4)
    state = ref_utils.restore_checkpoint(path, dict(optimizer=opt, model=ours, ema=ema, step=0), device='cpu')
    assert state['step'] == 40
    key_before, fp_before = ours.module._params_key(), ours.module._fingerprint()
    for (n1, p1), (n2, p2) in zip(rm.named_parameters(), ours.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2), n1
    # EMA is positional over the parameters that require grad (models/ema.py:20,52-55): the drop-in must freeze exactly
    # the parameters the reference freezes (SpecFormer's attention scale, specformer.py:381) or every later tensor shifts
    assert [p.requires_grad for p in rm.parameters()] == [p.requires_grad for p in ours.parameters()]
    ema.copy_to(ours.parameters())
    trainable = [(n, p) for n, p in ours.named_parameters() if p.requires_grad]
    assert len(trainable) == len(rema.shadow_params)
    for s, (n, p) in zip(rema.shadow_params, trainable):
        assert torch.equal(s, p), n
    # ema.copy_to writes through .data.copy_, which bumps no version counter: the cheap (data_ptr, version) key cannot
    # see it, the content fingerprint checked at the start of every sampling round does
    assert ours.module._params_key() == key_before
    assert not torch.equal(ours.module._fingerprint(), fp_before)


@needs_ref
def test_create_model_resolves_the_b200_names_through_the_reference_registry():
    """models/utils.py:5-28: `create_model(config)` = `_MODELS[config.model.name](config).to(device)` in a DataParallel;
    `diffspectra_b200.model.register()` is the only hook needed for `--config.model.name DMT_B200 | DMT_WO_EQ_B200`."""
    from diffspectra_b200 import model as M
    ref = ref_harness.load_reference()
    M.register(ref.mutils)
    M.register(ref.mutils)                                   # idempotent
    cfg = ref.config
    cfg.data.spectra_version = 'allspectra'
    cfg.model.pretrained_specformer_path = ''
    for name, cls, ref_cls in (('DMT_B200', M.DMT_B200, ref.DMT), ('DMT_WO_EQ_B200', M.DMT_WO_EQ_B200, ref.DMT_WO_EQ)):
        cfg.model.name = name
        m = ref.mutils.create_model(cfg)
        assert isinstance(m, torch.nn.DataParallel) and isinstance(m.module, cls)
        # same parameter inventory as the reference class built from the same config object (EMA order)
        cfg.model.name = ref_cls.__name__
        r = ref.mutils.create_model(cfg)
        assert [(n, tuple(p.shape)) for n, p in m.named_parameters()] == [(n, tuple(p.shape)) for n, p in r.named_parameters()]
