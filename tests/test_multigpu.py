"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): one process per GPU, probe universe sharded on
bitmap-word boundaries, failed/unit bitmaps merged by the engine's NVLink peer-memory kernel
(bbcp_merge_bitmaps). Checked against the golden failed-literal set, against an NCCL all-gather of the same
slices, and the per-rank implied sets against the golden ones."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, names, q):
    import torch.distributed as dist
    from intel_sat_solver_b200 import BatchedTopor, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    golden = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
    ok = True
    msgs = []
    for name in names:
        words = golden[f"{name}/words"]
        t = BatchedTopor(device=rank)
        t.AddClauses(words)
        assert t.Finalize() == 0
        s = sharding.shard_for(rank, world, t.max_var)
        sharding.peer_attach(t, rank, world)
        for rep in range(3):   # repeated merges: epochs advance, bitmaps accumulate
            res = t.ProbeRange(s.first, s.last)
            ms = sharding.merge_bitmaps(t, s)
        # the same with the merge as the last kernel of the batch (and a device-side rendezvous before it), with
        # state budgets that send some probes through the second launch sequence on some ranks only.
        # NO host barrier between the merge above and the clear: a rank must not wipe its slice while a slower peer
        # is still pulling it (the pull-done acknowledgement; the ranks are skewed on purpose)
        if rank % 2:
            import time
            time.sleep(0.05 * rank)
        t.ClearBitmaps()
        for rep in range(2):   # merge -> clear -> merge without any host synchronisation between the ranks
            res = t.ProbeRange(s.first, s.last, merge=s.words_per_rank)
            f1, _ = t.Bitmaps()
            nw = sharding.bitmap_words(t.max_var)
            bits = np.zeros(nw * 32, dtype=np.uint8)
            bits[np.nonzero(golden[f"{name}/probe_flag"] == 1)[0] + 2] = 1
            ok &= bool(np.array_equal(f1, np.packbits(bits, bitorder="little").view(np.uint32)))
            t.ClearBitmaps()
        dist.barrier()
        for kw in (dict(), dict(small_trail=32, mid_trail=32), dict(out_capacity=16), dict(implied=False)):
            res = t.ProbeRange(s.first, s.last, merge=s.words_per_rank, rendezvous=True, **kw)
        res = t.ProbeRange(s.first, s.last, merge=s.words_per_rank)
        flag, off, lits = golden[f"{name}/probe_flag"], golden[f"{name}/probe_off"], golden[f"{name}/probe_lits"]
        ok &= bool(np.array_equal(res.flag, flag[s.first:s.last]))
        ok &= bool(np.array_equal(res.impl_lits, lits[int(off[s.first]):int(off[s.last])]))
        failed, unit = t.Bitmaps()
        nwords = sharding.bitmap_words(t.max_var)
        bits = np.zeros(nwords * 32, dtype=np.uint8)
        bits[np.nonzero(flag == 1)[0] + 2] = 1
        expect = np.packbits(bits, bitorder="little").view(np.uint32)
        ok &= bool(np.array_equal(failed, expect))
        ub = np.unpackbits(unit.view(np.uint8), bitorder="little")
        ok &= bool(np.array_equal(np.nonzero(ub)[0], np.sort(np.nonzero(bits)[0] ^ 1)))
        # the NCCL all-gather of the same slices gives the same bitmap
        local = torch.from_numpy(failed.view(np.int32).copy()).cuda(rank)
        ok &= bool(np.array_equal(sharding.allgather_bitmap(local, s).cpu().numpy().view(np.uint32), expect))
        msgs.append((name, ok, ms))
        t.lib.bbcp_peer_detach(t.handle)
        dist.barrier()
    q.put((rank, ok, msgs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_p2p_merge_matches_golden_and_nccl(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000) + world
    names = ["fuzz_10", "regr_47", "fuzz_5"]
    procs = [ctx.Process(target=_worker, args=(r, world, port, names, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, msgs in got:
        assert ok, (rank, msgs)
