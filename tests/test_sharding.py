"""CPU tests of the N>1 host logic: shard ranges tile the probe universe on bitmap-word boundaries, and a
world_size-2 (and 3) gloo all-gather of per-rank bitmap slices reproduces the full failed-literal bitmap."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from intel_sat_solver_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("max_var", [1, 7, 15, 16, 17, 100, 1000, 12345])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shards_tile_the_universe(max_var, world):
    shards = sharding.all_shards(world, max_var)
    assert shards[0].first == 0 and shards[-1].last == 2 * max_var
    for a, b in zip(shards[:-1], shards[1:]):
        assert a.last == b.first and a.word_hi == b.word_lo
    for s in shards:
        assert s.first <= s.last
        if s.last > s.first:
            # every probe of the shard sets bits only inside the shard's words
            assert (s.first + 2) // 32 >= s.word_lo and (s.last - 1 + 2) // 32 < s.word_hi


def failed_bitmap(flags, lo, hi, words):
    bits = np.zeros(words * 32, dtype=np.uint8)
    idx = np.nonzero(flags == 1)[0]
    idx = idx[(idx >= lo) & (idx < hi)]
    bits[idx + 2] = 1
    return np.packbits(bits, bitorder="little").view(np.int32)


def _worker(rank, world, port, flags, max_var, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = sharding.shard_for(rank, world, max_var)
    words = sharding.bitmap_words(max_var)
    local = torch.from_numpy(failed_bitmap(flags, s.first, s.last, words).copy())
    merged = sharding.allgather_bitmap(local, s)
    q.put((rank, merged.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_allgather_merges_failed_bitmaps(golden, world):
    flags = golden["fuzz_10/probe_flag"]
    max_var = flags.size // 2
    words = sharding.bitmap_words(max_var)
    expect = failed_bitmap(flags, 0, flags.size, words)
    assert expect.any()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, flags, max_var, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, merged in got:
        assert np.array_equal(merged, expect), rank
