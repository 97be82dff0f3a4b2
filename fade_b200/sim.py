"""Synthetic inputs: random reference + simulated paired reads with injected
inverted-repeat soft clips.  Thin ctypes wrapper over libfadesim.so (csrc/sim/fadesim.cpp).
Test / bench infrastructure -- produces INPUTS only."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfadesim.so")
_lib = None


class SimCfg(C.Structure):
    _fields_ = [("read_seed", C.c_uint64), ("read_len", C.c_int32), ("window", C.c_int32),
                ("clip_min_art", C.c_int32), ("clip_max", C.c_int32), ("frag_mean", C.c_double),
                ("frag_sd", C.c_double), ("frag_max", C.c_int32), ("p_artifact", C.c_double),
                ("p_random_clip", C.c_double), ("p_both", C.c_double), ("p_indel", C.c_double),
                ("p_unmapped", C.c_double), ("p_sa", C.c_double), ("p_outside", C.c_double),
                ("sub_rate", C.c_double), ("short_clip_law", C.c_int32)]


def _l():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(f"{_SO} missing: run __graft_entry__.build()")
        L = C.CDLL(_SO)
        L.fadesim_contig.argtypes = [C.c_uint64, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_void_p]
        L.fadesim_contig.restype = None
        L.fadesim_default_cfg.argtypes = [C.POINTER(SimCfg)]
        L.fadesim_default_cfg.restype = None
        L.fadesim_reads.argtypes = [C.POINTER(SimCfg), C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_void_p),
                                    C.c_void_p] + [C.c_void_p] * 14
        L.fadesim_reads.restype = None
        L.fadesim_write_bam.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_char_p), C.c_void_p, C.c_int64, C.c_int32, C.c_int64] + \
            [C.c_void_p] * 9 + [C.c_int32]
        L.fadesim_write_bam.restype = C.c_int
        L.fadesim_write_fasta.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_char_p), C.c_void_p, C.POINTER(C.c_void_p), C.c_int32]
        L.fadesim_write_fasta.restype = C.c_int
        L.fadesim_set_threads.argtypes = [C.c_int]
        L.fadesim_set_threads.restype = None
        _lib = L
    return _lib


def set_threads(n: int):
    _l().fadesim_set_threads(n)


def default_cfg(**kw) -> SimCfg:
    c = SimCfg()
    _l().fadesim_default_cfg(C.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def make_contig(ref_seed: int, index: int, length: int, n_run_every: int = 0, n_run_len: int = 0,
                lower_frac: float = 0.0) -> np.ndarray:
    out = np.empty(length, dtype=np.uint8)
    _l().fadesim_contig(ref_seed, index, length, n_run_every, n_run_len, lower_frac, out.ctypes.data)
    return out


@dataclass
class Reads:
    n: int
    read_len: int
    seq4: np.ndarray
    seq_off: np.ndarray
    l_qseq: np.ndarray
    qual: np.ndarray | None
    cigar: np.ndarray | None
    n_cigar: np.ndarray | None
    flag: np.ndarray
    tid: np.ndarray
    pos: np.ndarray
    aligned_len: np.ndarray
    clip_left: np.ndarray
    clip_right: np.ndarray
    has_sa: np.ndarray
    truth: np.ndarray


def make_reads(cfg: SimCfg, first: int, n: int, contigs: list[np.ndarray], with_records: bool = True) -> Reads:
    """contigs: list of uint8 ASCII arrays (kept alive by the caller)."""
    L = cfg.read_len
    stride = (L + 1) // 2
    ptrs = (C.c_void_p * len(contigs))(*[c.ctypes.data for c in contigs])
    lens = np.array([len(c) for c in contigs], dtype=np.int64)
    seq4 = np.zeros(n * stride, dtype=np.uint8)
    seq_off = np.zeros(n + 1, dtype=np.int64)
    l_qseq = np.zeros(n, dtype=np.int32)
    qual = np.zeros(n * L, dtype=np.uint8) if with_records else None
    cigar = np.zeros((n, 6), dtype=np.uint32) if with_records else None
    n_cigar = np.zeros(n, dtype=np.int32) if with_records else None
    flag = np.zeros(n, dtype=np.int32)
    tid = np.zeros(n, dtype=np.int32)
    pos = np.zeros(n, dtype=np.int64)
    aligned_len = np.zeros(n, dtype=np.int32)
    clip_left = np.zeros(n, dtype=np.int32)
    clip_right = np.zeros(n, dtype=np.int32)
    has_sa = np.zeros(n, dtype=np.uint8)
    truth = np.zeros(n, dtype=np.uint8)
    p = lambda a: a.ctypes.data if a is not None else None  # noqa: E731
    _l().fadesim_reads(C.byref(cfg), first, n, len(contigs), ptrs, lens.ctypes.data, p(seq4), p(seq_off), p(l_qseq),
                       p(qual), p(cigar), p(n_cigar), p(flag), p(tid), p(pos), p(aligned_len), p(clip_left),
                       p(clip_right), p(has_sa), p(truth))
    return Reads(n, L, seq4, seq_off, l_qseq, qual, cigar, n_cigar, flag, tid, pos, aligned_len, clip_left,
                 clip_right, has_sa, truth)


def write_bam(path: str, names: list[str], contigs: list[np.ndarray], rd: Reads, name_base: int = 0, level: int = 1):
    """The simulated records (make_reads(..., with_records=True)) as a BGZF-compressed BAM file; the same records
    tests/samio.py:write_sam writes as text (QNAME r<k>, MAPQ 60, NM:i:0, SA:Z for has_sa reads)."""
    assert rd.qual is not None and rd.cigar is not None, "make_reads(..., with_records=True) needed"
    cn = (C.c_char_p * len(names))(*[s.encode() for s in names])
    lens = np.array([len(c) for c in contigs], dtype=np.int64)
    cig = np.ascontiguousarray(rd.cigar, dtype=np.uint32)
    rc = _l().fadesim_write_bam(str(path).encode(), len(names), cn, lens.ctypes.data, rd.n, rd.read_len, name_base,
                                rd.seq4.ctypes.data, rd.qual.ctypes.data, cig.ctypes.data, rd.n_cigar.ctypes.data,
                                rd.flag.ctypes.data, rd.tid.ctypes.data, rd.pos.ctypes.data, rd.aligned_len.ctypes.data,
                                rd.has_sa.ctypes.data, level)
    if rc:
        raise OSError(f"cannot write {path}")


def write_fasta(path: str, names: list[str], contigs: list[np.ndarray], width: int = 60):
    cn = (C.c_char_p * len(names))(*[s.encode() for s in names])
    lens = np.array([len(c) for c in contigs], dtype=np.int64)
    ptrs = (C.c_void_p * len(contigs))(*[c.ctypes.data for c in contigs])
    if _l().fadesim_write_fasta(str(path).encode(), len(names), cn, lens.ctypes.data, ptrs, width):
        raise OSError(f"cannot write {path}")


def config_c1():
    """C1 (correctness): 1 Mbp contig seed 1001 with one 500-bp N run at 400,000 and 1 % lower-case
    tiles; 10,000 reads 2x150 seed 2001."""
    ref = make_contig(1001, 0, 1_000_000, 0, 0, 0.01)
    ref[400_000:400_500] = ord("N")
    cfg = default_cfg(read_seed=2001)
    return ["chrS1"], [ref], cfg, 10_000
