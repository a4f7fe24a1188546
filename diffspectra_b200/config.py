"""QM9S DiffSpectra hyper-parameters needed by the sampling hot path (values of
configs/diffspectra_qm9s.py:9-151), as a plain attribute dict so no ml_collections install is needed.
A reference ``ml_collections.ConfigDict`` works interchangeably wherever a config is accepted."""


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def get_config(spectra_version='allspectra', device='cuda:0', precision='bf16'):
    config = ConfigDict()
    config.exp_type = 'diffspectra'
    config.pred_edge = True
    config.only_2D = False
    config.data = data = ConfigDict()
    data.name = 'QM9S'
    data.info_name = 'qm9_second_half'
    data.compress_edge = True
    data.centered = True
    data.include_aromatic = False
    data.atom_types = 5
    data.bond_types = 4
    data.fc_scale = [-1., 1.]
    data.max_node = 29
    data.spectra_version = spectra_version
    config.sde = sde = ConfigDict()
    sde.schedule = 'cosine'
    sde.continuous_beta_0 = 0.1
    sde.continuous_beta_1 = 20.
    config.model = model = ConfigDict()
    model.name = 'DMT_B200'
    model.b200_precision = precision
    model.pred_data = True
    model.include_fc_charge = True
    model.normalize_factors = '1, 4, 4, 1'
    model.edge_ch = 2
    model.nf = 256
    model.n_layers = 8
    model.n_heads = 16
    model.dropout = 0.1
    model.cond_time = True
    model.dist_gbf = True
    model.gbf_name = 'CondGaussianLayer'
    model.self_cond = True
    model.self_cond_type = 'ori'
    model.edge_quan_th = 0.
    model.n_extra_heads = 2
    model.CoM = True
    model.mlp_ratio = 2
    model.spatial_cut_off = 2.
    model.softmax_inf = True
    model.trans_name = 'TransMixLayer'
    model.cond_ch = 1
    model.pretrained_specformer_path = ''
    model.patch_len = [20, 50, 50]
    model.stride = [10, 25, 25]
    config.sampling = sampling = ConfigDict()
    sampling.method = 'ancestral'
    sampling.steps = 1000
    config.eval = evaluate = ConfigDict()
    evaluate.batch_size = 128
    evaluate.num_samples = 10000
    evaluate.sampling_temperature = 1.0
    config.seed = 42
    config.device = device
    return config
