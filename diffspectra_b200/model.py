"""``DMT_B200`` — drop-in replacement for the reference denoiser ``models.dmt.DMT`` (models/dmt.py:178-413).

Same constructor (``DMT_B200(config)``), same call signature (sampling.py:588-589), same parameter / buffer
names, shapes and REGISTRATION ORDER (the reference's EMA is a positional list, models/ema.py:20,52-55, and
checkpoints are loaded with strict=True, utils.py:15-19) — but ``forward`` runs the hand-written sm_100a kernels
of libdiffspectra_b200.so instead of ~1200 ATen/PyG launches.  The nn.Module tree below only HOLDS parameters;
its sub-modules are never called.
"""
import math

import numpy as np
import torch
from torch import nn

from . import _lib as L
from .engine import Engine

_SPEC_LEN = (701, 3501, 3501)


def _cfg(obj, name, default):
    try:
        return getattr(obj, name)
    except (AttributeError, KeyError):
        return default


# ----------------------------------------------------------------------------- parameter containers
class _CondGaussianParams(nn.Module):          # models/layers.py:314-326
    def __init__(self, K, time_dim):
        super().__init__()
        self.means = nn.Embedding(1, K - 1)
        self.stds = nn.Embedding(1, K - 1)
        self.time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_dim, 2))
        nn.init.uniform_(self.means.weight, 0, 3)
        nn.init.uniform_(self.stds.weight, 0, 3)


class _CoorsNormParams(nn.Module):             # models/layers.py:337-342
    def __init__(self, scale_init):
        super().__init__()
        self.scale = nn.Parameter(torch.zeros(1).fill_(scale_init))


class _TransMixParams(nn.Module):              # models/layers.py:98-120
    def __init__(self, x_channels, out_channels, extra_heads, heads, edge_dim):
        super().__init__()
        sub_heads = heads - extra_heads
        sub_channels = (heads * out_channels) // sub_heads
        self.lin_key = nn.Linear(x_channels, sub_heads * sub_channels)
        self.lin_query = nn.Linear(x_channels, sub_heads * sub_channels)
        self.lin_value = nn.Linear(x_channels, heads * out_channels)
        self.lin_edge0 = nn.Linear(edge_dim, sub_heads * sub_channels, bias=False)
        self.lin_edge1 = nn.Linear(edge_dim, heads * out_channels, bias=False)


class _EquiUpdateParams(nn.Module):            # models/dmt.py:20-35
    def __init__(self, hidden_dim, edge_dim, dist_dim, time_dim, extra_heads):
        super().__init__()
        self.coord_norm = _CoorsNormParams(1e-2)
        self.time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_dim, hidden_dim * 2))
        self.input_lin = nn.Linear(hidden_dim * 2 + edge_dim + dist_dim, hidden_dim)
        self.coord_mlp = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.SiLU(),
                                       nn.Linear(hidden_dim, 1 + extra_heads, bias=False))


class _BlockParams(nn.Module):                 # models/dmt.py:66-116
    def __init__(self, node_dim, edge_dim, time_dim, extra_heads, heads, mlp_ratio):
        super().__init__()
        self.edge_emb = nn.Linear(edge_dim * 2, edge_dim)
        self.node2edge_lin = nn.Linear(node_dim, edge_dim)
        self.attn_mpnn = _TransMixParams(node_dim, node_dim // heads, extra_heads, heads, edge_dim)
        self.ff_linear1 = nn.Linear(node_dim, node_dim * mlp_ratio)
        self.ff_linear2 = nn.Linear(node_dim * mlp_ratio, node_dim)
        self.ff_linear3 = nn.Linear(edge_dim, edge_dim * mlp_ratio)
        self.ff_linear4 = nn.Linear(edge_dim * mlp_ratio, edge_dim)
        self.equi_update = _EquiUpdateParams(node_dim, edge_dim, edge_dim, time_dim, extra_heads)
        self.node_time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_dim, node_dim * 6))
        self.edge_time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_dim, edge_dim * 6))
        self.dist_layer = _CondGaussianParams(edge_dim, time_dim)


class _SinEmbParams(nn.Module):                # models/layers.py:277-281
    def __init__(self, dim):
        super().__init__()
        self.weights = nn.Parameter(torch.randn(dim // 2))


class _SdpParams(nn.Module):                   # models/specformer.py:377-382
    def __init__(self, d_model, n_heads):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor((d_model // n_heads) ** -0.5), requires_grad=False)


class _MhaParams(nn.Module):                   # models/specformer.py:313-333
    def __init__(self, d_model, n_heads):
        super().__init__()
        self.W_Q = nn.Linear(d_model, d_model)
        self.W_K = nn.Linear(d_model, d_model)
        self.W_V = nn.Linear(d_model, d_model)
        self.sdp_attn = _SdpParams(d_model, n_heads)
        self.to_out = nn.Sequential(nn.Linear(d_model, d_model), nn.Dropout(0.))
        for w in (self.W_Q, self.W_K, self.W_V):
            nn.init.xavier_uniform_(w.weight)
            w.bias.data.fill_(0)


class _EncLayerParams(nn.Module):              # models/specformer.py:232-263
    def __init__(self, d_model, n_heads, d_ff):
        super().__init__()
        self.self_attn = _MhaParams(d_model, n_heads)
        self.norm_attn = nn.Sequential(nn.Identity(), nn.BatchNorm1d(d_model), nn.Identity())
        self.ff = nn.Sequential(nn.Linear(d_model, d_ff), nn.GELU(), nn.Dropout(0.), nn.Linear(d_ff, d_model))
        self.norm_ffn = nn.Sequential(nn.Identity(), nn.BatchNorm1d(d_model), nn.Identity())
        for name, p in self.ff.named_parameters():
            if 'weight' in name:
                nn.init.xavier_uniform_(p)
            else:
                nn.init.constant_(p, 0.)


class _EncoderParams(nn.Module):
    def __init__(self, d_model, n_heads, d_ff, n_layers):
        super().__init__()
        self.layers = nn.ModuleList([_EncLayerParams(d_model, n_heads, d_ff) for _ in range(n_layers)])


class _BackboneParams(nn.Module):              # models/specformer.py:123-157
    def __init__(self, patch_nums, patch_len, spectra_version, used, d_model, n_heads, d_ff, n_layers):
        super().__init__()
        self.W_P = nn.ModuleList([nn.Linear(patch_len[i], d_model) for i in used])

        def pos(q):
            w = torch.empty(q, d_model)
            nn.init.uniform_(w, -0.02, 0.02)
            return nn.Parameter(w)

        if spectra_version == 'allspectra':
            self.W_pos_uv = pos(patch_nums[0])
            self.W_pos_ir = pos(patch_nums[1])
            self.W_pos_raman = pos(patch_nums[2])
        else:
            self.W_pos = pos(patch_nums[0])
        self.encoder = _EncoderParams(d_model, n_heads, d_ff, n_layers)
        for w in self.W_P:
            nn.init.xavier_uniform_(w.weight)
            w.bias.data.fill_(0)


class _HeadParams(nn.Module):
    def __init__(self, nf, out):
        super().__init__()
        self.linear = nn.Linear(nf, out)
        nn.init.xavier_uniform_(self.linear.weight)
        self.linear.bias.data.fill_(0)


class _SpecFormerParams(nn.Module):            # models/specformer.py:14-69
    def __init__(self, patch_len, stride, output_dim, spectra_version, n_layers=3, d_model=128, n_heads=16, d_ff=256):
        super().__init__()
        used = {'uv': [0], 'ir': [1], 'raman': [2], 'allspectra': [0, 1, 2]}.get(spectra_version)
        if used is None:
            raise ValueError('spectra_version should be uv, ir, raman or allspectra')
        patch_nums = [int((_SPEC_LEN[i] - patch_len[i]) / stride[i] + 1) for i in used]
        self.patch_nums = patch_nums
        self.backbone = _BackboneParams(patch_nums, patch_len, spectra_version, used, d_model, n_heads, d_ff, n_layers)
        self.head = _HeadParams(d_model * sum(patch_nums), output_dim)
        self.out_norm = nn.LayerNorm(output_dim)


def _mlp3(i, h1, h2, o):
    return nn.Sequential(nn.Linear(i, h1), nn.SiLU(), nn.Linear(h1, h2), nn.SiLU(), nn.Linear(h2, o))


# ----------------------------------------------------------------------------- the modules
class _B200Denoiser(nn.Module):
    """Engine plumbing + the reference call signature shared by DMT_B200 and DMT_WO_EQ_B200.  Sub-classes only
    register parameters (same names / shapes / order as the reference class) and set MODEL_KIND."""
    MODEL_KIND = 'DMT'

    def _init_plumbing(self, config):
        self.spectra_version = config.data.spectra_version
        self.precision = _cfg(config.model, 'b200_precision', 'bf16')          # 'bf16' (tcgen05) | 'fp32' (validation)
        self._pretrained_specformer_path = _cfg(config.model, 'pretrained_specformer_path', '')
        self._engine = None
        self._weights_key = None
        self._weights_fp = None
        self._plan_cache = {}          # identity key -> (mask tensor kept alive, plan)
        self._plan_content = {}        # (n_atoms bytes, N, device) -> plan
        self._ctx_cache = (None, None)

    # ------------------------------------------------------------------ checkpoint ingestion
    _SPECFORMER_PREFIXES = ('model.representation_spec_model', 'model.representation_model')

    def _maybe_load_pretrained_specformer(self):
        """Called at the end of the sub-class constructors, where the reference does it (models/dmt.py:263-267)."""
        if self._pretrained_specformer_path:
            self.load_pretrained_specformer(self._pretrained_specformer_path)

    def load_pretrained_specformer(self, ckpt_path):
        """Same rule as DMT.load_pretrained_specformer (models/dmt.py:268-303): the checkpoint's 'state_dict' holds the
        encoder under the first prefix of _SPECFORMER_PREFIXES that occurs; every cond_encoder key whose prefixed
        source exists with the same shape is taken (out_norm.* always from 'model.representation_model.out_norm.*'),
        anything else keeps its initial value; a checkpoint without 'state_dict' or without a known prefix is ignored
        with a warning, like in the reference.  Returns the number of tensors taken.  The packed kernel weights are
        rebuilt lazily because load_state_dict bumps the parameter versions (see engine())."""
        ckpt = torch.load(ckpt_path, map_location='cpu', weights_only=False)
        if not isinstance(ckpt, dict) or 'state_dict' not in ckpt:
            print("Warning: pretrained model does not contain 'state_dict' key. Loading the entire checkpoint.")
            return 0
        src = ckpt['state_dict']
        prefix = next((p for p in self._SPECFORMER_PREFIXES if any(k.startswith(p) for k in src)), None)
        if prefix is None:
            print('Warning: No matching prefix found in the state_dict.')
            return 0
        own = self.cond_encoder.state_dict()
        taken = {}
        for key, cur in own.items():
            if key in ('out_norm.weight', 'out_norm.bias'):
                skey = 'model.representation_model.out_norm.' + key.rsplit('.', 1)[-1]
            else:
                skey = prefix + '.' + key
            val = src.get(skey)
            if val is not None and tuple(val.shape) == tuple(cur.shape):
                taken[key] = val
        if taken:
            self.cond_encoder.load_state_dict(taken, strict=False)
            print('Loaded %d keys from the pretrained SpecFormer model.' % len(taken))
        else:
            print('0 Warning: No matching keys found in the pretrained SpecFormer model.')
        return len(taken)

    # ------------------------------------------------------------------ engine plumbing
    def _params_key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _fingerprint(self):
        """Per-tensor L2 norms of every floating-point parameter / buffer, on their device.  `ema.copy_to` and
        `ema.restore` (models/ema.py:52-55,74-76) write through `param.data.copy_`, which changes neither data_ptr nor
        the autograd version counter, so the cheap key above cannot see them; this content check (a few multi-tensor
        kernels over 130 MB + one 2.6 KB compare) runs once per sampling round, not per denoiser call."""
        ts = [t.detach() for t in list(self.parameters()) + list(self.buffers()) if t.is_floating_point()]
        return torch.stack(torch._foreach_norm(ts))

    def engine(self, device=None, verify=False):
        """verify=True additionally compares the content fingerprint of the weights with the one taken when they were
        packed (used at the start of every sampling round)."""
        device = torch.device(device) if device is not None else next(self.parameters()).device
        if self._engine is None or self._engine.device != device or self._engine.mode_name != self.precision:
            self._engine = Engine(device, mode=self.precision, spectra_version=self.spectra_version, model_kind=self.MODEL_KIND)
            self._weights_key = None
            self._plan_cache = {}
            self._plan_content = {}
            self._ctx_cache = (None, None)
        key = self._params_key()
        stale = key != self._weights_key                 # load_state_dict / optimizer step / .to() happened
        fp = None
        if verify and not stale:
            fp = self._fingerprint()
            stale = self._weights_fp is None or fp.shape != self._weights_fp.shape or not torch.equal(fp, self._weights_fp)
        if stale:
            self._engine.pack_weights(self.state_dict())
            self._weights_key = key
            self._weights_fp = fp if fp is not None else self._fingerprint()
            self._ctx_cache = (None, None)
        return self._engine

    def invalidate_packed_weights(self):
        """Force a re-pack at the next call (for callers that modify parameters through `.data` between calls of the
        same sampling round)."""
        self._weights_key = None
        return self

    def set_precision(self, precision):
        assert precision in ('bf16', 'fp32')
        self.precision = precision
        return self

    def plan_for(self, node_mask):
        """Molecule plan of a node mask.  Two-level cache:
        * identity fast path (a denoiser call per step passes the SAME tensor 1000 times): keyed by
          (data_ptr, _version, shape, device) and the entry HOLDS the mask tensor, so its address cannot be handed to a
          different mask by the caching allocator while the entry lives (round r's mask used to be freed when round r+1
          rebound it, and round r+2 could get the same address back -> a stale plan for different molecules);
        * content key (tuple of atom counts, N): a new mask object with the same atom counts re-uses the plan, any other
          content builds a new one.  Costs one small D2H copy per NEW mask object (once per sampling round)."""
        key = (node_mask.data_ptr(), node_mask._version, tuple(node_mask.shape), str(node_mask.device))
        hit = self._plan_cache.get(key)
        if hit is not None and hit[0] is node_mask:
            return hit[1]
        nm = node_mask.reshape(node_mask.shape[0], -1)
        n_atoms = nm.sum(dim=1).round().to(torch.int32).cpu().numpy()
        # valid atoms must be a prefix (sampling.py:432-434)
        if not bool((nm[:, :1] > 0).all()) or not bool(((nm[:, 1:] - nm[:, :-1]) <= 0).all()):
            raise ValueError('%s expects prefix node masks (first n atoms valid) as built by sampling.py:432-434'
                             % type(self).__name__)
        ckey = (n_atoms.tobytes(), int(nm.shape[1]), str(node_mask.device))
        plan = self._plan_content.get(ckey)
        if plan is None:
            plan = self._engine.plan(n_atoms, nm.shape[1])
            if len(self._plan_content) >= 8:
                self._plan_content.clear()
            self._plan_content[ckey] = plan
        if len(self._plan_cache) >= 8:
            self._plan_cache.clear()
        self._plan_cache[key] = (node_mask, plan)
        return plan

    def context_embedding(self, context):
        """cond_lin(SpecFormer(context)), cached while the same spectra tensors are passed (they are constant
        across the 1000 steps of a sampling round; the reference recomputes them every call).  The cache entry holds
        the tensors themselves (identity compare + version), so a recycled address can never alias an old entry."""
        ts = list(context) if isinstance(context, (list, tuple)) else [context]
        key = tuple((t.data_ptr(), t._version, tuple(t.shape), str(t.device)) for t in ts)
        held = self._ctx_cache[0]
        same = held is not None and held[0] == key and len(held[1]) == len(ts) and all(a is b for a, b in zip(held[1], ts))
        if not same:                                     # a new sampling round: also the point where weights are re-verified
            self._ctx_cache = (None, None)
            eng = self.engine(verify=True)
            self._ctx_cache = ((key, ts), eng.context_embedding(context))
        return self._ctx_cache[1]

    # ------------------------------------------------------------------ reference interface
    def forward(self, t, xh, node_mask, edge_mask, context=None, *args, **kwargs):
        """Same contract as DMT.forward / DMT_WO_EQ.forward (models/dmt.py:306-321,413; models/dmt_wo_eq.py:811-829):
        returns ([B,N,9], [B,N,N,2]) fp32 on xh.device; `t`, `alpha_t`, `sigma_t` are accepted and unused like in the
        reference."""
        edge_x = kwargs['edge_x']
        cond_x, cond_edge_x = kwargs.get('cond_x'), kwargs.get('cond_edge_x')
        noise_level = kwargs['noise_level']
        if context is None:
            raise ValueError('%s is the spectra-conditioned denoiser: context must be given' % type(self).__name__)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError('%s implements the inference (sampling) path only; wrap calls in torch.no_grad() / '
                                      'model.eval() (training is out of scope, SURVEY.md §8(f).4)' % type(self).__name__)
        eng = self.engine(xh.device)
        plan = self.plan_for(node_mask)
        ctx_emb = self.context_embedding(context)
        return eng.denoise(plan, xh, edge_x, noise_level, ctx_emb, cond_x, cond_edge_x)


def _check_supported(config, who):
    m, d = config.model, config.data
    in_node_dim = d.atom_types + int(m.include_fc_charge)
    supported = (in_node_dim == 6 and m.nf == 256 and m.n_layers == 8 and m.n_heads == 16 and
                 m.n_extra_heads == 2 and m.mlp_ratio == 2 and m.edge_ch == 2 and m.dist_gbf and m.cond_time and
                 m.pred_data and m.CoM and m.softmax_inf and m.gbf_name == 'CondGaussianLayer' and
                 float(m.spatial_cut_off) == 2.0 and float(m.edge_quan_th) == 0.0 and
                 list(m.patch_len) == [20, 50, 50] and list(m.stride) == [10, 25, 25])
    if not supported:
        raise ValueError('%s kernels are specialised for the QM9S DiffSpectra configuration '
                         '(configs/diffspectra_qm9s.py:44-76); got a different model config' % who)


class DMT_B200(_B200Denoiser):
    """Conditional Diffusion Molecule Transformer with self-conditioning, B200-native forward."""
    MODEL_KIND = 'DMT'

    def __init__(self, config):
        super().__init__()
        m, d = config.model, config.data
        in_node_dim = d.atom_types + int(m.include_fc_charge)
        hidden_dim = m.nf
        edge_hidden_dim = m.nf // 4
        n_layers = m.n_layers
        time_dim = hidden_dim * 4
        _check_supported(config, 'DMT_B200')
        self.n_layers = n_layers

        self.node_emb = nn.Linear(in_node_dim * 2, hidden_dim)
        self.edge_emb = nn.Linear(m.edge_ch * 2 + edge_hidden_dim, edge_hidden_dim)
        self.dist_layer = _CondGaussianParams(edge_hidden_dim, time_dim)
        cat_node_dim = (hidden_dim * 2) // n_layers
        cat_edge_dim = (edge_hidden_dim * 2) // n_layers
        for i in range(n_layers):
            self.add_module('e_block_%d' % i, _BlockParams(hidden_dim, edge_hidden_dim, time_dim, m.n_extra_heads,
                                                           m.n_heads, m.mlp_ratio))
            self.add_module('node_%d' % i, nn.Linear(hidden_dim, cat_node_dim))
            self.add_module('edge_%d' % i, nn.Linear(edge_hidden_dim, cat_edge_dim))
        self.node_pred_mlp = _mlp3(cat_node_dim * n_layers + hidden_dim, hidden_dim, hidden_dim // 2, in_node_dim)
        self.edge_type_mlp = _mlp3(cat_edge_dim * n_layers + edge_hidden_dim, edge_hidden_dim, edge_hidden_dim // 2,
                                   m.edge_ch - 1)
        self.edge_exist_mlp = _mlp3(cat_edge_dim * n_layers + edge_hidden_dim, edge_hidden_dim, edge_hidden_dim // 2, 1)
        self.time_mlp = nn.Sequential(_SinEmbParams(16), nn.Linear(17, time_dim), nn.GELU(), nn.Linear(time_dim, time_dim))
        self.cond_encoder = _SpecFormerParams(m.patch_len, m.stride, hidden_dim, d.spectra_version)
        self.cond_lin = nn.Linear(hidden_dim, time_dim)
        self._init_plumbing(config)
        self._maybe_load_pretrained_specformer()


class _NodeEmbedParams(nn.Module):             # models/dmt_wo_eq.py:629-637
    def __init__(self, in_node_features, hidden_size):
        super().__init__()
        self.x_linear = nn.Linear(in_node_features, hidden_size * 2)
        self.pos_linear = nn.Linear(3, hidden_size * 2)
        self.mlp = nn.Sequential(nn.GELU(), nn.Linear(hidden_size * 2, hidden_size))


class _TransOptimV2Params(nn.Module):          # models/dmt_wo_eq.py:175-206
    def __init__(self, x_channels, out_channels, heads, edge_dim):
        super().__init__()
        self.lin_qkv = nn.Linear(x_channels, heads * out_channels * 3)
        self.lin_kv_e = nn.Linear(edge_dim, heads * out_channels * 2, bias=False)
        self.proj = nn.Linear(heads * out_channels, heads * out_channels)


class _WoBlockParams(nn.Module):               # models/dmt_wo_eq.py:389-472 (pair_update=True, cond_time=True, trans_ver='v2')
    def __init__(self, node_dim, edge_dim, time_dim, heads, mlp_ratio):
        super().__init__()
        self.attn_mpnn = _TransOptimV2Params(node_dim, node_dim // heads, heads, edge_dim)
        self.ff_linear1 = nn.Linear(node_dim, node_dim * mlp_ratio)
        self.ff_linear2 = nn.Linear(node_dim * mlp_ratio, node_dim)
        self.node2edge_lin = nn.Linear(node_dim * 2, edge_dim)
        self.ff_linear3 = nn.Linear(edge_dim, edge_dim * mlp_ratio)
        self.ff_linear4 = nn.Linear(edge_dim * mlp_ratio, edge_dim)
        self.node_time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_dim, node_dim * 6))
        self.edge_time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_dim, edge_dim * 6))


class DMT_WO_EQ_B200(_B200Denoiser):
    """Non-equivariant ablation (reference: models/dmt_wo_eq.py:646-937, registry name 'DMT_WO_EQ'), B200-native
    forward.  Same parameter names / shapes / registration order as the reference class."""
    MODEL_KIND = 'DMT_WO_EQ'

    def __init__(self, config):
        super().__init__()
        m, d = config.model, config.data
        _check_supported(config, 'DMT_WO_EQ_B200')
        if _cfg(m, 'trans_ver', 'v2') == 'v1':
            raise ValueError("DMT_WO_EQ_B200 implements trans_ver='v2' (TransLayerOptimV2), the reference default")
        in_node_dim = d.atom_types + int(m.include_fc_charge)
        hidden_dim, edge_hidden_dim, n_layers = m.nf, m.nf // 4, m.n_layers
        time_dim = hidden_dim * 4
        self.n_layers = n_layers
        self.node_emb = _NodeEmbedParams(in_node_dim * 2, hidden_dim)
        self.edge_emb = nn.Linear(m.edge_ch * 2 + edge_hidden_dim, edge_hidden_dim)
        self.dist_layer = _CondGaussianParams(edge_hidden_dim, time_dim)
        cat_node_dim = (hidden_dim * 2) // n_layers
        cat_edge_dim = (edge_hidden_dim * 2) // n_layers
        for i in range(n_layers):
            self.add_module('dmt_block_%d' % i, _WoBlockParams(hidden_dim, edge_hidden_dim, time_dim, m.n_heads, m.mlp_ratio))
            self.add_module('node_%d' % i, nn.Linear(hidden_dim, cat_node_dim))
            self.add_module('edge_%d' % i, nn.Linear(edge_hidden_dim, cat_edge_dim))
        self.node_pred_mlp = _mlp3(cat_node_dim * n_layers + hidden_dim, hidden_dim, hidden_dim // 2, in_node_dim)
        self.pos_pred_mlp = nn.Sequential(nn.Linear(cat_node_dim * n_layers + hidden_dim, hidden_dim, bias=False), nn.Tanh(),
                                          nn.Linear(hidden_dim, 3, bias=False))
        self.edge_type_mlp = _mlp3(cat_edge_dim * n_layers + edge_hidden_dim, edge_hidden_dim, edge_hidden_dim // 2,
                                   m.edge_ch - 1)
        self.edge_exist_mlp = _mlp3(cat_edge_dim * n_layers + edge_hidden_dim, edge_hidden_dim, edge_hidden_dim // 2, 1)
        self.time_mlp = nn.Sequential(_SinEmbParams(16), nn.Linear(17, time_dim), nn.GELU(), nn.Linear(time_dim, time_dim))
        self.cond_encoder = _SpecFormerParams(m.patch_len, m.stride, hidden_dim, d.spectra_version)
        self.cond_lin = nn.Linear(hidden_dim, time_dim)
        self._init_plumbing(config)
        self._maybe_load_pretrained_specformer()


def register(models_utils=None):
    """Register DMT_B200 / DMT_WO_EQ_B200 in the reference's model registry (models/utils.py:5-21) so that
    `main.py --mode eval --config.model.name DMT_B200` builds them through create_model()."""
    if models_utils is None:
        import models.utils as models_utils          # the reference package, when it is on sys.path
    for cls in (DMT_B200, DMT_WO_EQ_B200):
        if cls.__name__ not in models_utils._MODELS:
            models_utils.register_model(cls, name=cls.__name__)
    return DMT_B200
