"""Stand-ins for torch_geometric.utils.{dense_to_sparse, softmax} (PyG 2.4.0 semantics)."""
import torch
from torch_scatter import scatter


def dense_to_sparse(adj, mask=None):
    assert adj.dim() == 3 and mask is None
    idx = adj.nonzero(as_tuple=True)
    edge_attr = adj[idx]
    row = idx[1] + adj.size(-2) * idx[0]
    col = idx[2] + adj.size(-1) * idx[0]
    return torch.stack([row, col], dim=0), edge_attr


def softmax(src, index=None, ptr=None, num_nodes=None, dim=0):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    src_max = scatter(src.detach(), index, dim, dim_size=n, reduce='max')
    out = (src - src_max.index_select(dim, index)).exp()
    out_sum = scatter(out, index, dim, dim_size=n, reduce='sum') + 1e-16
    return out / out_sum.index_select(dim, index)


def to_dense_adj(*a, **k):
    raise NotImplementedError


def to_dense_batch(*a, **k):
    raise NotImplementedError
