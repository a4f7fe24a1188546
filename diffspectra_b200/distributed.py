"""Multi-GPU plumbing of the sampling path: one process per GPU, independent contiguous shards of
the sample list, NO collective inside the step loop, one all-gather of fixed-size molecule records at the end of a
round.  Replaces nn.DataParallel's per-forward parameter broadcast + scatter/gather (models/utils.py:27)."""
import math

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous slice [lo, hi) of rank `rank` over n_items (ceil split, last ranks may be short or empty)."""
    per = int(math.ceil(n_items / world_size))
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def record_bytes(N):
    return N * 12 + N + N + N * N + 1


def pack_records(pos, atom_type, fc, bond, n_atoms):
    """[B, record_bytes(N)] uint8: pos f32[N*3] | atom u8[N] | fc i8[N] | bond u8[N*N] | n u8.
    Host / eager restatement of the record layout, kept for the CPU tests of the gather logic and as the checker of the
    device kernel (ds_molecule_records, Engine.molecule_records), which is what the product path uses."""
    B, N = atom_type.shape
    return torch.cat([pos.contiguous().view(torch.uint8).reshape(B, N * 12),
                      atom_type.to(torch.uint8),
                      fc.to(torch.int8).view(torch.uint8),
                      bond.to(torch.uint8).reshape(B, N * N),
                      n_atoms.to(torch.uint8).reshape(B, 1)], dim=1).contiguous()


def unpack_records(rec, N):
    """Inverse of pack_records on the host: list of (pos [n,3] f32, atom_type [n] i64, bond [n,n] f32, fc [n] i64) —
    the tuple layout of sampling.mol_process (sampling.py:12-32).  Molecules are sliced in groups of equal atom count
    (one gather + unbind per distinct n instead of four tensor slices per molecule: 10 000 records in ~30 ms)."""
    rec = rec.cpu()
    B = rec.shape[0]
    o = 0
    pos = rec[:, o:o + N * 12].contiguous().view(torch.float32).reshape(B, N, 3); o += N * 12
    atom = rec[:, o:o + N].long(); o += N
    fc = rec[:, o:o + N].contiguous().view(torch.int8).long(); o += N
    bond = rec[:, o:o + N * N].float().reshape(B, N, N); o += N * N
    n = rec[:, o].long()
    out = [None] * B
    for k in torch.unique(n).tolist():
        idx = (n == k).nonzero().squeeze(1)
        p, a = pos[idx, :k].unbind(0), atom[idx, :k].unbind(0)
        b, f = bond[idx, :k, :k].unbind(0), fc[idx, :k].unbind(0)
        for j, i in enumerate(idx.tolist()):
            out[i] = (p[j], a[j], b[j], f[j])
    return out


def gather_records(rec):
    """All ranks end up with the records of every rank, rank-major (equal shard sizes required; pad short shards)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rec
    out = torch.empty((dist.get_world_size() * rec.shape[0],) + tuple(rec.shape[1:]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous())
    return out
