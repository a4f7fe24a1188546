"""Stand-in for torch_geometric.nn.conv.MessagePassing (aggr='add', node_dim=0, source_to_target)."""
import inspect
import torch
from torch_scatter import scatter


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr='add', flow='source_to_target', node_dim=-2, **kw):
        super().__init__()
        assert aggr == 'add' and flow == 'source_to_target' and node_dim == 0
        self._msg_params = [p for p in inspect.signature(self.message).parameters]

    def propagate(self, edge_index, size=None, **kwargs):
        j, i = edge_index[0], edge_index[1]
        n = next(v for v in kwargs.values() if torch.is_tensor(v)).size(0)
        args = {}
        for p in self._msg_params:
            if p == 'size_i':
                args[p] = n
            elif p == 'index':
                args[p] = i
            elif p == 'ptr':
                args[p] = None
            elif p.endswith('_i'):
                args[p] = kwargs[p[:-2]].index_select(0, i)
            elif p.endswith('_j'):
                args[p] = kwargs[p[:-2]].index_select(0, j)
            else:
                args[p] = kwargs[p]
        return scatter(self.message(**args), i, 0, dim_size=n, reduce='sum')
