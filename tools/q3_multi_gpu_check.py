#!/usr/bin/env python
"""Multi-GPU parity: the key-sharded Q3 dataflow over N GPUs (NCCL exchange) vs the CPU
oracle dataflow with N workers, on the same seeded inputs.  Run under torchrun:
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/q3_multi_gpu_check.py
Rank 0 prints PARITY OK / FAIL."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import materialize_b200 as mz  # noqa: E402
from materialize_b200 import _ffi as F  # noqa: E402
from materialize_b200 import harness  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = mz.Context(local, rank, world)
    idbuf = (C.c_uint8 * F.COMM_ID_BYTES)()
    if rank == 0:
        ctx.check(F.lib.mzgpu_comm_unique_id(idbuf))
    t = torch.tensor(list(idbuf), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    idbuf = (C.c_uint8 * F.COMM_ID_BYTES)(*t.cpu().tolist())
    ctx.check(F.lib.mzgpu_comm_init(ctx.h, idbuf))

    args = dict(seed=7, n_customer=3000, n_orders=30000, n_part=4000, per_batch=500)
    g = harness.Q3Dataflow(ctx, worker=rank, peers=world, **args)
    if os.environ.get("MZGPU_P2P", "1") != "0":
        # update-batch exchange rounds over peer memory (CUDA IPC handles all-gathered on the host side)
        hnd = mz.p2p_export(ctx, 1 << 16, 32)
        ht = torch.tensor(list(hnd), dtype=torch.uint8, device="cuda")
        hs = [torch.zeros_like(ht) for _ in range(world)]
        dist.all_gather(hs, ht)
        mz.p2p_import(ctx, [bytes(x.cpu().tolist()) for x in hs])
        dist.barrier()
        g.use_p2p(True)
        if rank == 0:
            print("exchange rounds over peer memory", flush=True)

    def gather_rows(rows):
        """all ranks' ROUT rows on rank 0"""
        n = torch.tensor([len(rows)], dtype=torch.int64, device="cuda")
        ns = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(ns, n)
        mx = max(int(x.item()) for x in ns)
        buf = torch.zeros(max(mx, 1) * 64, dtype=torch.uint8, device="cuda")
        if len(rows):
            buf[: len(rows) * 64] = torch.from_numpy(rows.view(np.uint8).copy()).cuda()
        bufs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(bufs, buf)
        out = [b[: int(k.item()) * 64].cpu().numpy().view(mz.ROUT) for b, k in zip(bufs, ns)]
        return np.concatenate(out) if out else np.zeros(0, dtype=mz.ROUT)

    ok = True
    if rank == 0:
        from oracle import binding as oracle

        o = oracle.Q3(workers=world, **args)
        o.hydrate()
    g.hydrate()
    got = gather_rows(g.out_rows())
    g.clear_out()
    if rank == 0:
        want = o.drain()
        ok = ok and oracle.consolidate(got).tobytes() == want.tobytes()
        print("hydrate", len(got), len(want), ok, flush=True)
    for b in range(8):
        g.stage_batch(b, g.time())
        g.step()
        got = gather_rows(g.out_rows())
        g.clear_out()
        if rank == 0:
            o.step(b)
            want = o.drain()
            same = oracle.consolidate(got).tobytes() == want.tobytes()
            ok = ok and same
            print("batch", b, len(got), len(want), same, flush=True)
    if rank == 0:
        print("PARITY OK" if ok else "PARITY FAIL", ctx.stats(), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
