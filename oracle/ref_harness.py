"""Import the UNMODIFIED reference (read-only /root/reference) behind third-party stand-ins.

TEST INFRASTRUCTURE ONLY.  Works only in the build container (the GPU box has no
/root/reference); used by oracle/gen_golden.py to generate the committed fixtures under
tests/golden/ and by the `not gpu` tests that cross-check the oracle restatement against
the real reference when it is present.  Never imported by the product package.

Recipe = SURVEY.md Appendix E.
"""
import importlib.util
import os
import sys

REFERENCE_ROOT = os.environ.get('DIFFSPECTRA_REFERENCE', '/root/reference')
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_shims')


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'models', 'dmt.py'))


def load_reference():
    """Returns a namespace of the reference symbols on the sampling hot path."""
    import torch
    if not reference_available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    for p in (REFERENCE_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [_SHIMS, REFERENCE_ROOT]
    # configs/diffspectra_qm9s.py:110 divides by torch.cuda.device_count()
    real_count = torch.cuda.device_count
    if real_count() == 0:
        torch.cuda.device_count = lambda: 1
    try:
        spec = importlib.util.spec_from_file_location(
            '_ref_cfg_qm9s', os.path.join(REFERENCE_ROOT, 'configs', 'diffspectra_qm9s.py'))
        cfg = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(cfg)
        config = cfg.get_config()
    finally:
        torch.cuda.device_count = real_count
    config.device = torch.device('cpu')

    import models.dmt as ref_dmt                      # noqa: E402  (reference module)
    import models.dmt_wo_eq as ref_dmt_wo_eq          # noqa: E402
    import models.utils as ref_mutils                 # noqa: E402
    import models.specformer as ref_specformer        # noqa: E402
    import diffusion.noise_schedule as ref_ns         # noqa: E402
    import sampling as ref_sampling                   # noqa: E402
    import utils as ref_utils                         # noqa: E402

    class NS:
        pass
    ns = NS()
    ns.config = config
    ns.DMT = ref_dmt.DMT
    ns.DMT_WO_EQ = ref_dmt_wo_eq.DMT_WO_EQ
    ns.SpecFormer = ref_specformer.SpecFormer
    ns.NoiseScheduleVP = ref_ns.NoiseScheduleVP
    ns.AncestralSampler = ref_sampling.AncestralSampler
    ns.post_process = ref_sampling.post_process
    ns.mol_process = ref_sampling.mol_process
    ns.get_self_cond_fn = ref_utils.get_self_cond_fn
    ns.get_data_inverse_scaler = ref_utils.get_data_inverse_scaler
    ns.mutils = ref_mutils
    ns.sampling = ref_sampling
    return ns
