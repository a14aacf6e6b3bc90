"""Flat multigrid hierarchies (SURVEY.md 8a') on the host.

Holds, per level, what `gpuls` PreProcess extracts from UG's VECTOR/MATRIX lists
(ug_b200/host/gpuls_flatten.cc): BSR in VSTART->MNEXT order, per-row flags and the standard
P/R stencils.  `load_ugh` reads the dump format written by the flattening code's host tools
(records `[u32 namelen][name][u8 dtype][u64 count][raw]` after the magic ``UGH1\\n``).

This module is plumbing for tests and the bench; it does no arithmetic of the hot path.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

_DTYPES = {0: np.int32, 1: np.float64, 2: np.uint8, 3: np.uint32}


def load_ugh(path: str) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:5] != b"UGH1\n":
        raise ValueError(f"{path}: not a UGH1 file")
    pos = 5
    while pos < len(buf):
        (nl,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        name = buf[pos:pos + nl].decode()
        pos += nl
        dt = buf[pos]
        pos += 1
        (cnt,) = struct.unpack_from("<Q", buf, pos)
        pos += 8
        dtype = np.dtype(_DTYPES[dt])
        out[name] = np.frombuffer(buf, dtype=dtype, count=cnt, offset=pos).copy()
        pos += cnt * dtype.itemsize
    return out


@dataclass
class Level:
    n: int
    bs: int
    rowptr: np.ndarray
    col: np.ndarray
    val: np.ndarray
    vclass: np.ndarray
    vnclass: np.ndarray
    ctl: np.ndarray
    skip: np.ndarray
    p_rowptr: Optional[np.ndarray] = None
    p_col: Optional[np.ndarray] = None
    p_w: Optional[np.ndarray] = None
    r_rowptr: Optional[np.ndarray] = None
    r_col: Optional[np.ndarray] = None
    r_w: Optional[np.ndarray] = None
    rhs: Optional[np.ndarray] = None
    xyz: Optional[np.ndarray] = None

    @property
    def nnz(self) -> int:
        return int(self.rowptr[-1])


@dataclass
class Hierarchy:
    dim: int
    bs: int
    fullrefinelevel: int
    levels: List[Level]
    meta: Dict[str, float] = field(default_factory=dict)
    raw: Dict[str, np.ndarray] = field(default_factory=dict)

    @property
    def top(self) -> int:
        return len(self.levels) - 1

    @classmethod
    def from_ugh(cls, path: str) -> "Hierarchy":
        d = load_ugh(path)
        top = int(d["toplevel"][0])
        bs = int(d["bs"][0])
        levels = []
        for l in range(top + 1):
            g = lambda k: d.get(f"L{l}/{k}")
            levels.append(Level(
                n=int(g("n")[0]), bs=bs, rowptr=g("rowptr"), col=g("col"), val=g("val"),
                vclass=g("vclass"), vnclass=g("vnclass"), ctl=g("ctl"), skip=g("skip"),
                p_rowptr=g("p_rowptr"), p_col=g("p_col"), p_w=g("p_w"),
                r_rowptr=g("r_rowptr"), r_col=g("r_col"), r_w=g("r_w"),
                rhs=g("rhs"), xyz=g("xyz")))
        meta = {k: float(d[k][0]) for k in ("damp", "nu1", "nu2", "gamma") if k in d}
        return cls(dim=int(d["dim"][0]), bs=bs, fullrefinelevel=int(d["fullrefinelevel"][0]),
                   levels=levels, meta=meta, raw=d)
