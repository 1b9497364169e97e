"""Binary file formats shared by the engine's tools, the bench and the test-side checkers.

Formats (little-endian; the test-side C tools read and write the same ones):

* instance ``.bcnf``: u64 magic "BCNF0001", u64 nvars_hint, u64 nclauses, u64 nwords, i32 words[nwords]
  (clauses, each 0-terminated)
* cubes ``.cubes``: u64 magic "CUBE0001", u64 ncubes, u64 nlits, u64 off[ncubes+1], i32 lits[nlits]
* results ``.bres``: u64 magic "BRES0001", u64 n, u64 nlits, u64 status, u64 nlevel0, u64 implications,
  u64 assignments, f64 seconds, u8 flag[n] (pad 8), u32 props[n] (pad 8), u64 off[n+1], i32 lits[nlits] (pad 8),
  i32 level0[nlevel0]

Text instances follow the reference CLI's incremental format (``Main.cc:207-221``): clause lines are
kept in file order; lines starting with s / r / o / l / b / n / c / p are ignored.
"""
from __future__ import annotations

import dataclasses

import numpy as np

BCNF_MAGIC = int.from_bytes(b"BCNF0001", "little")
CUBE_MAGIC = int.from_bytes(b"CUBE0001", "little")
BRES_MAGIC = int.from_bytes(b"BRES0001", "little")

FLAG_OK, FLAG_CONFLICT, FLAG_TRIVIAL_L0, FLAG_FALSE_L0, FLAG_UNMAPPED, FLAG_NOT_RUN = 0, 1, 2, 3, 4, 255


def _pad8(n: int) -> int:
    return (n + 7) & ~7


def read_text_cnf(path: str) -> np.ndarray:
    """Clause lines of a (possibly incremental) CNF file as one int32 stream of 0-terminated clauses."""
    words: list[int] = []
    with open(path, "r") as f:
        for line in f:
            s = line.strip()
            if not s or not (s[0] == "-" or s[0].isdigit()):
                continue
            toks = [int(t) for t in s.split()]
            if 0 in toks:
                toks = toks[: toks.index(0) + 1]
            else:
                toks.append(0)
            words.extend(toks)
    return np.asarray(words, dtype=np.int32)


def read_text_queries(path: str) -> list[list[int]]:
    """The ``s <assumptions> 0`` query lines of an incremental CNF file (each is a cube)."""
    out = []
    with open(path, "r") as f:
        for line in f:
            s = line.split()
            if s and s[0] == "s":
                lits = [int(t) for t in s[1:]]
                if 0 in lits:
                    lits = lits[: lits.index(0)]
                out.append(lits)
    return out


def write_bcnf(path: str, words: np.ndarray, nvars_hint: int | None = None) -> None:
    words = np.ascontiguousarray(words, dtype=np.int32)
    if nvars_hint is None:
        nvars_hint = int(np.abs(words).max()) if words.size else 0
    ncls = int((words == 0).sum())
    with open(path, "wb") as f:
        np.asarray([BCNF_MAGIC, nvars_hint, ncls, words.size], dtype=np.uint64).tofile(f)
        words.tofile(f)


def read_bcnf(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        hdr = np.fromfile(f, dtype=np.uint64, count=4)
        if hdr.size != 4 or int(hdr[0]) != BCNF_MAGIC:
            raise ValueError(f"{path}: not a .bcnf file")
        return np.fromfile(f, dtype=np.int32, count=int(hdr[3]))


def read_instance(path: str) -> np.ndarray:
    return read_bcnf(path) if path.endswith(".bcnf") else read_text_cnf(path)


def cubes_to_csr(cubes) -> tuple[np.ndarray, np.ndarray]:
    off = np.zeros(len(cubes) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(c) for c in cubes], dtype=np.uint64)
    lits = np.asarray([l for c in cubes for l in c], dtype=np.int32)
    return lits, off


def write_cubes(path: str, lits: np.ndarray, off: np.ndarray) -> None:
    lits = np.ascontiguousarray(lits, dtype=np.int32)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    with open(path, "wb") as f:
        np.asarray([CUBE_MAGIC, off.size - 1, lits.size], dtype=np.uint64).tofile(f)
        off.tofile(f)
        lits.tofile(f)


@dataclasses.dataclass
class Results:
    status: int                # 0 ok, 1 contradictory at level 0
    flag: np.ndarray           # u8[n]
    props: np.ndarray          # u32[n]
    off: np.ndarray            # u64[n+1]
    lits: np.ndarray           # i32[off[n]] external literals, sorted by |var| inside each set
    level0: np.ndarray         # i32[], ascending |var|
    implications: int = 0
    assignments: int = 0
    seconds: float = 0.0

    def implied(self, i: int) -> np.ndarray:
        return self.lits[int(self.off[i]): int(self.off[i + 1])]


def read_bres(path: str) -> Results:
    buf = np.fromfile(path, dtype=np.uint8)
    hdr = buf[:56].view(np.uint64)
    if int(hdr[0]) != BRES_MAGIC:
        raise ValueError(f"{path}: not a .bres file")
    n, nlits, status, nl0, impl, asg = (int(x) for x in hdr[1:7])
    seconds = float(buf[56:64].view(np.float64)[0])
    p = 64
    flag = buf[p:p + n].copy(); p += _pad8(n)
    props = buf[p:p + 4 * n].view(np.uint32).copy(); p += _pad8(4 * n)
    off = buf[p:p + 8 * (n + 1)].view(np.uint64).copy(); p += 8 * (n + 1)
    lits = buf[p:p + 4 * nlits].view(np.int32).copy(); p += _pad8(4 * nlits)
    level0 = buf[p:p + 4 * nl0].view(np.int32).copy()
    return Results(status, flag, props, off, lits, level0, impl, asg, seconds)
