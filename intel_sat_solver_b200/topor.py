"""Host-side mirror of the reference's public facade for the batched-BCP path.

``BatchedTopor`` keeps the method names, argument meaning and value conventions of the reference's ``CTopor``
(``Topor.hpp:26-121``) for the subset that sits on this path -- ``AddClause`` / ``Solve(assumps)`` /
``GetLitValue`` / ``GetLitDecLevel`` / ``Backtrack`` / ``IsError`` / ``GetStatusExplanation`` /
``GetPropagations`` -- and adds the batch entry points the GPU engine exists for (``ProbeAll``,
``ProbeRange``, ``PropagateBatch``). Every call goes through the C ABI of ``include/bbcp.h`` (ctypes); no
propagation is done in Python or on the CPU.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import enum

import numpy as np

from . import _lib
from ._lib import BatchInfo, DbInfo, Opts

OPT_IMPLIED = 1
OPT_NO_LEN_PRUNE = 2
OPT_OFF32 = 4
OPT_KERNEL_TIMES = 8
OPT_MERGE = 16
OPT_RENDEZVOUS = 32
OPT_NO_REUSE = 64
OPT_COMPACT_SELF = 128
OPT_BINARY_ONLY = 256
FLAG_SELF = 0x80


class TToporReturnVal(enum.IntEnum):
    """ToporExternalTypes.hpp:52-78 (the values this path can produce)."""
    RET_SAT = 0
    RET_UNSAT = 1
    RET_MEM_OUT = 4
    RET_USER_INTERRUPT = 5   # fixpoint reached, search not run (what SetCbStopNow->VAL_STOP yields in the reference)
    RET_INDEX_TOO_NARROW = 6
    RET_EXOTIC_ERROR = 11


class TToporLitVal(enum.IntEnum):
    """ToporExternalTypes.hpp:81-87."""
    VAL_SATISFIED = 0
    VAL_UNSATISFIED = 1
    VAL_UNASSIGNED = 2
    VAL_DONT_CARE = 3


class BbcpError(RuntimeError):
    pass


@dataclasses.dataclass
class BatchResult:
    info: dict
    flag: np.ndarray | None = None       # u8[n]
    impl_off: np.ndarray | None = None   # u64[n+1]
    impl_lits: np.ndarray | None = None  # i32[impl_off[n]] external literals sorted by |var| per set

    def implied(self, i: int) -> np.ndarray:
        return self.impl_lits[int(self.impl_off[i]): int(self.impl_off[i + 1])]


class BatchedTopor:
    def __init__(self, varsNumHint: int = 0, device: int | None = None):
        self._L = _lib.load()
        self._h = self._L.bbcp_create(int(varsNumHint))
        if not self._h:
            raise BbcpError("bbcp_create failed")
        if device is not None:
            self._L.bbcp_set_device(self._h, int(device))
        self._assumps: list[int] = []

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.bbcp_release(h)

    # ---- CTopor-shaped surface -------------------------------------------------------------------------
    def AddClause(self, lits) -> None:
        """Topor.hpp:33-38: the span may be 0-terminated; reading stops at the first 0."""
        a = np.ascontiguousarray(lits, dtype=np.int32)
        self._L.bbcp_add_clause(self._h, a.ctypes.data, a.size)

    def AddClauses(self, words: np.ndarray) -> None:
        """Bulk form: one int32 stream of 0-terminated clauses."""
        a = np.ascontiguousarray(words, dtype=np.int32)
        self._check(self._L.bbcp_add_clauses(self._h, a.ctypes.data, a.size))

    def Finalize(self) -> int:
        rc = self._L.bbcp_finalize(self._h)
        if rc < 0:
            raise BbcpError(self.GetStatusExplanation())
        return rc

    def Solve(self, assumps=()) -> TToporReturnVal:
        """Topor.hpp:48 restricted to propagation: RET_UNSAT iff unit propagation of the assumptions derives a
        conflict, RET_SAT iff every variable got assigned, else RET_USER_INTERRUPT (fixpoint, no search)."""
        for l in assumps:
            if l == 0:
                break
            self._L.bbcp_assume(self._h, int(l))
        r = self._L.bbcp_solve(self._h)
        if self._L.bbcp_status(self._h) < 0:
            return TToporReturnVal.RET_EXOTIC_ERROR
        return {20: TToporReturnVal.RET_UNSAT, 10: TToporReturnVal.RET_SAT}.get(r, TToporReturnVal.RET_USER_INTERRUPT)

    def GetLitValue(self, l: int) -> TToporLitVal:
        v = self._L.bbcp_val(self._h, int(l))
        return {1: TToporLitVal.VAL_SATISFIED, -1: TToporLitVal.VAL_UNSATISFIED, 0: TToporLitVal.VAL_UNASSIGNED}.get(v, TToporLitVal.VAL_DONT_CARE)

    def GetLitDecLevel(self, l: int) -> int:
        return self._L.bbcp_declevel(self._h, int(l))

    def Backtrack(self, decLevel: int) -> None:
        self._L.bbcp_backtrack(self._h, int(decLevel))

    def IsError(self) -> bool:
        return self._L.bbcp_status(self._h) < 0

    def GetStatusExplanation(self) -> str:
        return self._L.bbcp_error_string(self._h).decode()

    def GetPropagations(self) -> int:
        return self._L.bbcp_propagations(self._h)

    # ---- batch surface ------------------------------------------------------------------------------------
    def _check(self, rc: int) -> int:
        if rc < 0:
            raise BbcpError(f"bbcp error {rc}: {self.GetStatusExplanation()}")
        return rc

    @staticmethod
    def _opts(implied=True, small_trail=0, out_capacity=0, no_len_prune=False, mid_trail=0, off32=False, kernel_times=False,
              merge=0, rendezvous=False, no_reuse=False, compact_self=False, binary_only=False) -> Opts:
        """merge: 0, or the bitmap words owned by each rank (sharding.Shard.words_per_rank) to all-gather the slices in the batch"""
        return Opts((OPT_IMPLIED if implied else 0) | (OPT_NO_LEN_PRUNE if no_len_prune else 0) | (OPT_OFF32 if off32 else 0)
                    | (OPT_KERNEL_TIMES if kernel_times else 0) | (OPT_MERGE if merge else 0) | (OPT_RENDEZVOUS if rendezvous else 0)
                    | (OPT_NO_REUSE if no_reuse else 0) | (OPT_COMPACT_SELF if compact_self else 0) | (OPT_BINARY_ONLY if binary_only else 0),
                    small_trail, out_capacity, mid_trail, int(merge))

    def _fetch(self, info: BatchInfo, implied: bool, fetch: bool, off32: bool = False) -> BatchResult:
        res = BatchResult(info.as_dict())
        if fetch:
            res.flag = np.empty(info.n, dtype=np.uint8)
            self._check(self._L.bbcp_fetch_flags(self._h, res.flag.ctypes.data))
            if implied:
                res.impl_off = np.empty(info.n + 1, dtype=np.uint32 if off32 else np.uint64)
                res.impl_lits = np.empty(max(info.n_implied, 1), dtype=np.int32)
                f = self._L.bbcp_fetch_implied32 if off32 else self._L.bbcp_fetch_implied
                self._check(f(self._h, res.impl_off.ctypes.data, res.impl_lits.ctypes.data))
                res.impl_lits = res.impl_lits[: info.n_implied]
        return res

    def ProbeRange(self, first: int, last: int, implied=True, fetch=True, **kw) -> BatchResult:
        info = BatchInfo()
        o = self._opts(implied=implied, **kw)
        self._check(self._L.bbcp_probe_range(self._h, first, last, C.byref(o), C.byref(info)))
        return self._fetch(info, implied, fetch, kw.get("off32", False))

    def ProbeLits(self, lits, implied=True, fetch=True, **kw) -> BatchResult:
        """Explicit list of depth-1 probes (4 bytes per probe on the way in)."""
        a = np.ascontiguousarray(lits, dtype=np.int32)
        info = BatchInfo()
        o = self._opts(implied=implied, **kw)
        self._check(self._L.bbcp_probe_lits(self._h, a.ctypes.data if a.size else None, a.size, C.byref(o), C.byref(info)))
        return self._fetch(info, implied, fetch, kw.get("off32", False))

    def ProbeAll(self, implied=True, fetch=True, **kw) -> BatchResult:
        return self.ProbeRange(0, 2 * self.max_var, implied=implied, fetch=fetch, **kw)

    def ProbeLitsToHost(self, lits, impl_cap=None, flag_out=None, off_out=None, impl_out=None, **kw) -> BatchResult:
        """ProbeLits + fetch in one call (bbcp_probe_lits_to_host). Arrays in PINNED memory (views of torch pin_memory()
        tensors) are used in place: the first kernel reads the list and writes the flags over PCIe, no staging copies;
        otherwise upload, kernels and read-back of successive chunks overlap on copy streams. Output arrays may be
        passed in; impl_cap = literals they hold."""
        a = np.ascontiguousarray(lits, dtype=np.int32)
        off32 = kw.get("off32", False)
        n = a.size
        flag_out = np.empty(n, dtype=np.uint8) if flag_out is None else flag_out
        off_out = np.empty(n + 1, dtype=np.uint32 if off32 else np.uint64) if off_out is None else off_out
        if impl_out is None:
            impl_out = np.empty(max(int(impl_cap if impl_cap is not None else 4 * n + 1024), 1), dtype=np.int32)
        info = BatchInfo()
        o = self._opts(implied=True, **kw)
        self._check(self._L.bbcp_probe_lits_to_host(self._h, a.ctypes.data if n else None, n, C.byref(o), flag_out.ctypes.data, off_out.ctypes.data,
                                                    impl_out.ctypes.data, impl_out.size, C.byref(info)))
        return BatchResult(info.as_dict(), flag_out[:n], off_out[: n + 1], impl_out[: info.n_implied])

    @staticmethod
    def ExpandCompact(lits, res: BatchResult) -> BatchResult:
        """Host-side expansion of a compact read-back (compact_self=True: BBCP_FLAG_SELF instead of a stored one-literal
        set; offsets not written at all when nothing was stored) into the standard flag / CSR form."""
        lits = np.asarray(lits, dtype=np.int32)
        n = lits.size
        self_ = (res.flag & FLAG_SELF) != 0
        flag = (res.flag & ~np.uint8(FLAG_SELF)).astype(np.uint8)
        stored = np.diff(res.impl_off.astype(np.int64)) if res.info["n_implied"] else np.zeros(n, dtype=np.int64)
        sizes = stored + self_
        off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(sizes, out=off[1:])
        out = np.empty(int(off[n]), dtype=np.int32)
        out[off[:-1][self_].astype(np.int64)] = lits[self_]
        if res.info["n_implied"]:
            nz = np.nonzero(stored)[0]
            src0, dst0 = res.impl_off.astype(np.int64)[nz], off[:-1].astype(np.int64)[nz]
            idx = np.repeat(dst0 - src0, stored[nz]) + np.arange(int(res.info["n_implied"]))
            out[idx] = res.impl_lits
        return BatchResult(res.info, flag, off, out)

    def PropagateBatch(self, cube_lits, cube_off, implied=True, fetch=True, **kw) -> BatchResult:
        lits = np.ascontiguousarray(cube_lits, dtype=np.int32)
        off = np.ascontiguousarray(cube_off, dtype=np.uint64)
        if lits.size == 0:
            lits = np.zeros(1, dtype=np.int32)
        info = BatchInfo()
        o = self._opts(implied=implied, **kw)
        self._check(self._L.bbcp_propagate_batch(self._h, lits.ctypes.data, off.ctypes.data, off.size - 1, C.byref(o), C.byref(info)))
        return self._fetch(info, implied, fetch, kw.get("off32", False))

    def ProbeHbr(self, pair_cap: int = 1 << 20) -> np.ndarray:
        """Hyper-binary resolvents of an all-literal pass (bbcp_probe_hbr): an (m, 2) array of (v, x), sorted."""
        n = C.c_uint64(0)
        pairs = np.zeros((max(pair_cap, 1), 2), dtype=np.int32)
        self._check(self._L.bbcp_probe_hbr(self._h, pairs.ctypes.data, pair_cap, C.byref(n)))
        if n.value > pair_cap:
            return self.ProbeHbr(int(n.value))
        p = pairs[: n.value]
        return p[np.lexsort((p[:, 1], p[:, 0]))] if p.size else p

    def CheckFixpoints(self, first: int = 0, count: int = 1 << 62, stride: int = 1) -> dict:
        """No-missed-implication check of the last batch's fixpoints on the device (bbcp_check_fixpoints)."""
        v, c, i = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._check(self._L.bbcp_check_fixpoints(self._h, int(first), int(count), int(stride), C.byref(v), C.byref(c), C.byref(i)))
        return {"violations": v.value, "clauses_checked": c.value, "items_checked": i.value}

    def ProbeToFixpoint(self, max_rounds: int = 0) -> dict:
        """Failed-literal probing with unit feedback until a pass finds no new failed literal (bbcp_probe_to_fixpoint).
        Returns {"status": 0 | 1 (contradictory), "rounds": passes run, "units": failed literals turned into units}."""
        rounds, units = C.c_uint32(0), C.c_uint64(0)
        rc = self._check(self._L.bbcp_probe_to_fixpoint(self._h, int(max_rounds), C.byref(rounds), C.byref(units)))
        return {"status": rc, "rounds": rounds.value, "units": units.value}

    def ProbeProducts(self, equiv_cap: int = 1 << 20) -> tuple[np.ndarray, np.ndarray]:
        """Products of the last ProbeAll / even-bounded ProbeRange with implied sets: (necessary literals, ascending
        |var|; equivalence pairs (v, x) as an (m, 2) array, sorted). Computed on the device (bbcp_probe_products)."""
        words = self._L.bbcp_bitmap_words(self._h)
        nec = np.zeros(words, dtype=np.uint32)
        n_nec, n_eq = C.c_uint64(0), C.c_uint64(0)
        while True:
            eq = np.zeros((max(equiv_cap, 1), 2), dtype=np.int32)
            self._check(self._L.bbcp_probe_products(self._h, nec.ctypes.data, C.byref(n_nec), eq.ctypes.data, equiv_cap, C.byref(n_eq)))
            if n_eq.value <= equiv_cap:
                break
            equiv_cap = int(n_eq.value)
        il = np.nonzero(np.unpackbits(nec.view(np.uint8), bitorder="little"))[0]
        lits = np.where(il & 1, -(il >> 1), il >> 1).astype(np.int32)
        eq = eq[: n_eq.value]
        eq = eq[np.lexsort((eq[:, 1], eq[:, 0]))] if eq.size else eq
        return lits, eq

    def Bitmaps(self) -> tuple[np.ndarray, np.ndarray]:
        n = self._L.bbcp_bitmap_words(self._h)
        failed, unit = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
        self._check(self._L.bbcp_fetch_bitmaps(self._h, failed.ctypes.data, unit.ctypes.data))
        return failed, unit

    def ClearBitmaps(self) -> None:
        self._check(self._L.bbcp_clear_bitmaps(self._h))

    def FailedLiterals(self) -> np.ndarray:
        """External literals whose depth-1 probe hit a conflict, ascending |var| (positive before negative)."""
        failed, _ = self.Bitmaps()
        bits = np.unpackbits(failed.view(np.uint8), bitorder="little")
        il = np.nonzero(bits)[0]
        return np.where(il & 1, -(il >> 1), il >> 1).astype(np.int32)

    def Level0(self) -> np.ndarray:
        n = self._L.bbcp_level0_lits(self._h, None, 0)
        out = np.zeros(max(n, 1), dtype=np.int32)
        self._L.bbcp_level0_lits(self._h, out.ctypes.data, n)
        return out[:n]

    def DbInfo(self) -> dict:
        d = DbInfo()
        self._L.bbcp_get_db_info(self._h, C.byref(d))
        return d.as_dict()

    @property
    def max_var(self) -> int:
        return self.DbInfo()["max_var"]

    @property
    def status(self) -> int:
        return self._L.bbcp_status(self._h)

    @property
    def handle(self):
        return self._h

    @property
    def lib(self):
        return self._L

    def HostPrepare(self) -> int:
        return self._L.bbcp_host_prepare(self._h)

    def HostRow(self, ilit: int) -> np.ndarray:
        n = self._L.bbcp_host_row(self._h, ilit, None, 0)
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self._L.bbcp_host_row(self._h, ilit, out.ctypes.data, n)
        return out[:n]

    def HostOrder(self) -> np.ndarray:
        """The probe schedule of the prepared database (children of the binary implication graph first); empty
        when the database has no binary clause."""
        n = self._L.bbcp_host_order(self._h, None, 0)
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self._L.bbcp_host_order(self._h, out.ctypes.data, n)
        return out[:n]
