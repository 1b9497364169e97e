"""Probe sharding across ranks and the merge of the failed-literal / unit bitmaps.

The probe universe (+1,-1,+2,-2,...) is cut into one CONTIGUOUS range per rank. Ranges are
aligned to 32-bit words of the failed-literal bitmap (bit index = internal literal = universe index + 2), so
that each rank owns a whole slice of bitmap words and ONE all-gather of the slices (NCCL over NVLink on the
GPU box, gloo in the CPU tests) merges them -- no reduction, no second pass. The clause database is
replicated; nothing else is exchanged on the data path.

Two implementations of that all-gather:

* ``peer_attach`` + ``merge_bitmaps``: the engine's own kernel over NVLink peer memory (``bbcp_merge_bitmaps``), a
  PULL protocol: every rank raises an epoch flag in every peer once its own slice is final, waits for the peers'
  flags and then copies the peers' slices of the FAILED-literal bitmap out of their (CUDA-IPC-mapped) memory with
  16-byte remote loads; the unit bitmap of those slices is derived on the spot (negating a literal swaps the two
  bits of its variable). A rank OVERWRITES its copy of the peers' slices (it does not OR into it): bits it held
  there from another sharding are discarded. When a rank has finished pulling it tells its peers so (a second
  epoch word per rank); ``bbcp_clear_bitmaps``, a re-finalize and ``bbcp_peer_detach`` wait for that before they
  touch the bitmaps, so no host barrier is needed between a merge and a clear. One launch, no staging copies;
  ``torch.distributed`` is used once, at set-up, to exchange the CUDA IPC handles.
* ``allgather_bitmap``: ``torch.distributed.all_gather_into_tensor`` (NCCL / gloo) -- the baseline the kernel
  is checked against, and the path the CPU tests cover.
"""
from __future__ import annotations

import dataclasses

import torch
import torch.distributed as dist


@dataclasses.dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    words_per_rank: int    # bitmap words owned by each rank (the last one may own fewer real words)
    word_lo: int           # this rank's words are [word_lo, word_hi)
    word_hi: int
    first: int             # probe-universe indices [first, last) belong to this rank
    last: int


def bitmap_words(max_var: int) -> int:
    return (2 * (max_var + 1) + 31) // 32


def shard_for(rank: int, world: int, max_var: int) -> Shard:
    universe = 2 * max_var
    words = bitmap_words(max_var)
    wp = ((words + world - 1) // world + 3) & ~3    # multiple of 4 words: the merge kernel moves 16-byte vectors
    word_lo, word_hi = min(rank * wp, words), min((rank + 1) * wp, words)
    first = min(max(32 * word_lo - 2, 0), universe)
    last = min(max(32 * word_hi - 2, 0), universe)
    return Shard(rank, world, wp, word_lo, word_hi, first, last)


def all_shards(world: int, max_var: int) -> list[Shard]:
    return [shard_for(r, world, max_var) for r in range(world)]


def allgather_bitmap(local: torch.Tensor, shard: Shard, group=None) -> torch.Tensor:
    """local: int32[bitmap_words] holding (at least) this rank's slice. Returns the merged int32[bitmap_words]."""
    words = local.numel()
    if shard.world == 1:
        return local.clone()
    send = torch.zeros(shard.words_per_rank, dtype=local.dtype, device=local.device)
    n = shard.word_hi - shard.word_lo
    if n > 0:
        send[:n].copy_(local[shard.word_lo:shard.word_hi])
    out = torch.empty(shard.words_per_rank * shard.world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, send, group=group)
    return out[:words]


def peer_attach(engine, rank: int, world: int, group=None) -> None:
    """Exchange the CUDA IPC handles of every rank's bitmap block and map the peers (once per finalize)."""
    import ctypes as C

    from ._lib import PEER_HANDLE_BYTES
    L, h = engine.lib, engine.handle
    mine = (C.c_ubyte * PEER_HANDLE_BYTES)()
    rc = L.bbcp_peer_export(h, mine)
    if rc:
        raise RuntimeError(f"bbcp_peer_export: {rc}: {engine.GetStatusExplanation()}")
    blobs = [None] * world
    if world > 1:
        dist.all_gather_object(blobs, bytes(mine), group=group)
    else:
        blobs[0] = bytes(mine)
    allb = (C.c_ubyte * (PEER_HANDLE_BYTES * world)).from_buffer_copy(b"".join(blobs))
    rc = L.bbcp_peer_attach(h, rank, world, allb)
    if rc:
        raise RuntimeError(f"bbcp_peer_attach: {rc}: {engine.GetStatusExplanation()}")


def device_barrier(engine) -> None:
    """The handshake of the merge alone: returns when every rank's stream has reached this point."""
    rc = engine.lib.bbcp_merge_bitmaps(engine.handle, 0, None)
    if rc:
        raise RuntimeError(f"bbcp_merge_bitmaps: {rc}: {engine.GetStatusExplanation()}")


def merge_bitmaps(engine, shard: Shard) -> float:
    """All-gather of the failed-literal and unit bitmap slices over NVLink peer memory; blocks until every
    peer's slice has arrived. Returns the device time in milliseconds (incl. waiting for the slowest rank)."""
    import ctypes as C
    ms = C.c_float(0.0)
    rc = engine.lib.bbcp_merge_bitmaps(engine.handle, shard.words_per_rank, C.byref(ms))
    if rc:
        raise RuntimeError(f"bbcp_merge_bitmaps: {rc}: {engine.GetStatusExplanation()}")
    return ms.value
