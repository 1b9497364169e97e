#!/usr/bin/env python
"""Diagnostics (run under torchrun, one rank per GPU): time the stand-alone peer-memory bitmap merge."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from intel_sat_solver_b200 import BatchedTopor, sharding  # noqa: E402


class DevArr:
    def __init__(self, ptr, n, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    nv = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    rng = np.random.default_rng(1)
    cl = rng.integers(1, nv + 1, size=(nv * 2, 3)).astype(np.int32) * rng.choice([-1, 1], size=(nv * 2, 3)).astype(np.int32)
    words = np.concatenate([cl, np.zeros((nv * 2, 1), np.int32)], axis=1).ravel()
    t = BatchedTopor(device=lr)
    t.AddClauses(words)
    assert t.Finalize() == 0
    s = sharding.shard_for(rank, world, t.max_var)
    sharding.peer_attach(t, rank, world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{lr}")
    res = []
    for it in range(12):
        flush.fill_(it)
        dist.barrier()
        torch.cuda.synchronize()
        r = t.ProbeRange(s.first, s.last, fetch=False, merge=s.words_per_rank, rendezvous=True, kernel_times=True)
        r0 = t.ProbeRange(s.first, s.last, fetch=False, kernel_times=True)
        res.append((round(r.info["total_ms"] * 1e3, 1), round(r.info["compact_ms"] * 1e3, 1), round(r0.info["total_ms"] * 1e3, 1), round(r0.info["compact_ms"] * 1e3, 1)))
    print(rank, "in-batch merge: total us, compact us; without merge: total us, compact us:", res[2:], flush=True)
    out = []
    for it in range(12):
        dist.barrier()
        torch.cuda.synchronize()
        sharding.device_barrier(t)
        ms = sharding.merge_bitmaps(t, s)
        out.append(round(ms * 1e3, 1))
    print(rank, "stand-alone merge, us (events around the launch, incl. waiting for the peers):", out[2:], flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
