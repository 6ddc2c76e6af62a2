"""world_size-2 (gloo, CPU) test of the host-side logic of the multi-GPU path: the Exchange
routing function (mzgpu_route, the one the partition kernel uses), the counts-then-payload
all-to-all protocol of mzgpu_exchange, and the claim the sharding rests on — keyed operators
(consolidate, accumulable reduce) over key-hash shards union to the single-worker result.
The per-shard arithmetic is the CPU oracle here; on GPUs it is libmzgpu (tools/q3_multi_gpu_check.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, seed, ret):
    import torch
    import torch.distributed as dist

    from materialize_b200 import _ffi as F
    from oracle import binding as oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed + rank)
    n = 20000
    rows = np.zeros(n, dtype=oracle.R32)
    rows["key"] = rng.integers(0, 3000, size=n, dtype=np.uint64)
    rows["val"] = rng.integers(0, 1000, size=n, dtype=np.uint64)
    rows["time"] = rng.integers(0, 3, size=n, dtype=np.uint64)
    rows["diff"] = rng.integers(-2, 3, size=n, dtype=np.int64)

    # ---- Exchange: partition by mzgpu_route, counts all-to-all, payload all-to-all
    dest = np.array([F.lib.mzgpu_route(int(k), world) for k in rows["key"]], dtype=np.int64)
    parts = [np.ascontiguousarray(rows[dest == p]) for p in range(world)]
    send_counts = torch.tensor([len(p) for p in parts], dtype=torch.int64)
    recv_counts = torch.zeros(world, dtype=torch.int64)
    dist.all_to_all_single(recv_counts, send_counts)
    send = [torch.from_numpy(p.view(np.uint8).reshape(-1).copy()) for p in parts]
    recv = [torch.zeros(int(c) * 32, dtype=torch.uint8) for c in recv_counts]
    # grouped point-to-point sends/receives, as mzgpu_exchange does with ncclSend/ncclRecv
    reqs = []
    for p in range(world):
        if p == rank:
            recv[p].copy_(send[p])
            continue
        if len(send[p]):
            reqs.append(dist.isend(send[p], p))
        if len(recv[p]):
            reqs.append(dist.irecv(recv[p], p))
    for r in reqs:
        r.wait()
    mine = np.concatenate([r.numpy().view(oracle.R32) for r in recv]) if world else rows
    # every row I hold routes to me
    assert all(F.lib.mzgpu_route(int(k), world) == rank for k in mine["key"][:2000])

    # ---- keyed operators on my shard
    cons = oracle.consolidate(mine)
    red = oracle.Reduce(0).step(mine, 3)

    # ---- gather everything on rank 0 and compare with one worker doing it all
    def gather(a, dt):
        n_ = torch.tensor([a.nbytes], dtype=torch.int64)
        ns = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(ns, n_)
        mx = max(int(x) for x in ns)
        buf = torch.zeros(mx, dtype=torch.uint8)
        buf[: a.nbytes] = torch.from_numpy(a.view(np.uint8).reshape(-1).copy())
        bufs = [torch.zeros(mx, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(bufs, buf)
        return np.concatenate([b[: int(k)].numpy().view(dt) for b, k in zip(bufs, ns)])

    all_rows = gather(rows, oracle.R32)
    all_cons = gather(cons, oracle.R32)
    all_red = gather(red, oracle.ROUT)
    ok = True
    if rank == 0:
        want_cons = oracle.consolidate(all_rows)
        want_red = oracle.Reduce(0).step(all_rows, 3)
        ok = oracle.consolidate(all_cons).tobytes() == want_cons.tobytes() and len(all_cons) == len(want_cons)
        ok = ok and oracle.consolidate(all_red).tobytes() == oracle.consolidate(want_red).tobytes()
    dist.barrier()
    dist.destroy_process_group()
    ret[rank] = ok


def test_two_workers_exchange_and_sharded_operators():
    import torch.multiprocessing as mp

    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, 11, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)


def test_route_is_fnv1a_mod_peers():
    """Hashable::hashed for u64 in DD 0.23 = FNV-1a 64 over the 8 LE bytes;
    timely routes by hash % peers."""
    from materialize_b200 import _ffi as F

    def fnv(k):
        h = 0xCBF29CE484222325
        for i in range(8):
            h ^= (k >> (8 * i)) & 0xFF
            h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
        return h

    rng = np.random.default_rng(5)
    for k in [0, 1, 2**63, 2**64 - 1] + [int(x) for x in rng.integers(0, 2**63, size=200)]:
        for peers in (1, 2, 3, 8):
            assert F.lib.mzgpu_route(k, peers) == fnv(k) % peers
