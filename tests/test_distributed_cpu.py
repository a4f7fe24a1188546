"""world_size-2 gloo test of the multi-GPU host logic: sharding, record packing, the single all-gather."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffspectra_b200.distributed import gather_records, pack_records, record_bytes, shard_range, unpack_records


def _fake_shard(rank, B, N):
    g = torch.Generator().manual_seed(100 + rank)
    n = torch.randint(1, N + 1, (B,), generator=g)
    pos = torch.randn(B, N, 3, generator=g)
    atom = torch.randint(0, 5, (B, N), generator=g, dtype=torch.int32)
    fc = torch.randint(-2, 3, (B, N), generator=g, dtype=torch.int32)
    bond = torch.randint(0, 4, (B, N, N), generator=g).float()
    return pos, atom, fc, bond, n


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    B, N = 5, 29
    rec = pack_records(*_fake_shard(rank, B, N))
    assert rec.shape == (B, record_bytes(N)) and rec.dtype == torch.uint8
    allrec = gather_records(rec)
    q.put((rank, allrec.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n, w in ((10000, 8), (10, 4), (3, 8), (1024, 1)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_records_roundtrip_and_allgather_gloo_world2():
    world, port = 2, 29731
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(got[0], got[1])                   # every rank holds the same, rank-major gathered records
    mols = unpack_records(got[0], 29)
    assert len(mols) == 10
    for rank in range(world):
        pos, atom, fc, bond, n = _fake_shard(rank, 5, 29)
        for i in range(5):
            p, a, e, f = mols[rank * 5 + i]
            k = int(n[i])
            assert torch.equal(p, pos[i, :k]) and torch.equal(a, atom[i, :k].long())
            assert torch.equal(e, bond[i, :k, :k]) and torch.equal(f, fc[i, :k].long())
