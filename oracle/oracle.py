"""ctypes binding of oracle/libfadeoracle.so (the C restatement, fade_oracle.c).

TEST INFRASTRUCTURE ONLY.  Build with `make -C oracle` (done by __graft_entry__.build()).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfadeoracle.so")

OPS = "MIDNSHP=XB"
FO_S, FO_EQ, FO_X, FO_I, FO_D = 4, 7, 8, 1, 2
# fo_params.switches (fade_oracle.h)
FO_SW_NO_SOFTCLIP_PAD, FO_SW_END_LAST_COL, FO_SW_E_BEFORE_F, FO_SW_GAP_TIE_OPEN, FO_SW_EQ_BY_MATRIX = 1, 2, 4, 8, 16
FO_SW_SWAP_ID, FO_SW_WILD_MISMATCH = 32, 64


class Params(C.Structure):
    _fields_ = [("gap_open", C.c_int32), ("gap_extend", C.c_int32), ("match", C.c_int32),
                ("mismatch", C.c_int32), ("window_size", C.c_int32), ("min_length", C.c_int32),
                ("switches", C.c_uint32)]


class SwResult(C.Structure):
    _fields_ = [("score", C.c_int32), ("end_query", C.c_int32), ("end_ref", C.c_int32),
                ("beg_query", C.c_int32), ("beg_ref", C.c_int32), ("n_ops", C.c_int32),
                ("ref_span", C.c_int32)]


class ReadResult(C.Structure):
    _fields_ = [("aligned", C.c_int32), ("art_left", C.c_int32), ("art_right", C.c_int32),
                ("win_start", C.c_int64), ("tlen", C.c_int32), ("sw", SwResult)]


class Record(C.Structure):
    _fields_ = [("is_mapped", C.c_int32), ("has_sa", C.c_int32), ("cigar", C.POINTER(C.c_uint32)),
                ("n_cigar", C.c_int32), ("seq4", C.POINTER(C.c_uint8)), ("qual", C.POINTER(C.c_uint8)),
                ("l_qseq", C.c_int32), ("pos", C.c_int64), ("contig_name", C.c_char_p),
                ("ref_seq", C.c_char_p), ("ref_len", C.c_int64)]


class Tags(C.Structure):
    _fields_ = [("rs", C.c_uint8), ("has_tags", C.c_int32), ("am", C.c_void_p), ("as_", C.c_void_p),
                ("ar", C.c_void_p), ("ab", C.c_void_p)]


class FuzzReport(C.Structure):
    _fields_ = [("n_pairs", C.c_int64), ("n_diverged", C.c_int64), ("n_gapped", C.c_int64),
                ("n_multi_max", C.c_int64), ("n_zero_ef", C.c_int64), ("first_div", C.c_int64),
                ("first_qlen", C.c_int32), ("first_tlen", C.c_int32), ("first_q", C.c_char * 512),
                ("first_t", C.c_char * 2048), ("explained_by", C.c_uint32)]


_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "fade_oracle.c")) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "fade_oracle_simd.c")) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "parasail_striped.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.fo_default_params.argtypes = [C.POINTER(Params)]
        L.fo_sw_trace.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(Params),
                                  C.POINTER(SwResult), C.POINTER(C.c_uint32), C.c_int]
        L.fo_sw_trace.restype = C.c_int
        L.fo_revcomp_nt16.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.fo_decode_nt16.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.fo_parse_clips.argtypes = [C.POINTER(C.c_uint32), C.c_int, C.POINTER(C.c_uint32)]
        L.fo_cigar_ref_span.argtypes = [C.POINTER(C.c_uint32), C.c_int]
        L.fo_cigar_ref_span.restype = C.c_int64
        L.fo_align_read.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_uint32, C.c_uint32,
                                    C.c_char_p, C.c_int64, C.POINTER(Params), C.POINTER(ReadResult),
                                    C.POINTER(C.c_uint32), C.c_int]
        L.fo_align_read.restype = C.c_int
        L.fo_annotate_record.argtypes = [C.POINTER(Record), C.POINTER(Params), C.POINTER(Tags)]
        L.fo_annotate_record.restype = C.c_int
        L.fo_free_tags.argtypes = [C.POINTER(Tags)]
        L.fo_align_batch.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_char_p),
                                     C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.fo_align_batch.restype = C.c_int
        L.fo_align_batch_simd.argtypes = L.fo_align_batch.argtypes
        L.fo_align_batch_simd.restype = C.c_int
        L.fo_simd_lanes.argtypes = []
        L.fo_simd_lanes.restype = C.c_int
        L.ps_sw_trace.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(Params), C.c_int,
                                  C.POINTER(SwResult), C.POINTER(C.c_uint32), C.c_int]
        L.ps_sw_trace.restype = C.c_int
        L.ps_fuzz.argtypes = [C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(Params), C.c_int,
                              C.POINTER(FuzzReport)]
        L.ps_fuzz.restype = C.c_int
        _lib = L
    return _lib


def simd_lanes() -> int:
    """int16 lanes of the SIMD port on this host: 32 (AVX-512BW), 16 (AVX2; also when FADE_ORACLE_SIMD=avx2) or 0."""
    return int(lib().fo_simd_lanes())


def default_params(**kw) -> Params:
    p = Params()
    lib().fo_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def cigar_string(ops) -> str:
    return "".join(f"{int(o) >> 4}{OPS[int(o) & 0xf]}" for o in ops)


def cigar_from_string(s: str) -> np.ndarray:
    out, num = [], ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append((int(num) << 4) | OPS.index(ch))
            num = ""
    return np.array(out, dtype=np.uint32)


@dataclass
class Sw:
    score: int
    end_query: int
    end_ref: int
    beg_query: int
    beg_ref: int
    n_ops: int
    ref_span: int
    ops: list = field(default_factory=list)

    @property
    def cigar(self) -> str:
        return cigar_string(self.ops)


def sw_trace(q: bytes | str, t: bytes | str, params: Params | None = None, ops_cap: int = 4096) -> Sw:
    """P1-P5 on ASCII query (already reverse-complemented) and target."""
    if isinstance(q, str):
        q = q.encode()
    if isinstance(t, str):
        t = t.encode()
    p = params or default_params()
    r = SwResult()
    ops = (C.c_uint32 * ops_cap)()
    rc = lib().fo_sw_trace(q, len(q), t, len(t), C.byref(p), C.byref(r), ops, ops_cap)
    if rc:
        raise ValueError("fo_sw_trace failed")
    return Sw(r.score, r.end_query, r.end_ref, r.beg_query, r.beg_ref, r.n_ops, r.ref_span,
              [int(ops[k]) for k in range(min(r.n_ops, ops_cap))])


def sw_trace_striped(q: bytes | str, t: bytes | str, lanes: int = 16, params: Params | None = None,
                     ops_cap: int = 4096) -> Sw:
    """The same contract through the structural restatement of parasail's striped kernel
    (oracle/parasail_striped.c): lanes = 8 for the 128-bit builds, 16 for AVX2."""
    if isinstance(q, str):
        q = q.encode()
    if isinstance(t, str):
        t = t.encode()
    p = params or default_params()
    r = SwResult()
    ops = (C.c_uint32 * ops_cap)()
    rc = lib().ps_sw_trace(q, len(q), t, len(t), C.byref(p), lanes, C.byref(r), ops, ops_cap)
    if rc:
        raise ValueError("ps_sw_trace failed")
    return Sw(r.score, r.end_query, r.end_ref, r.beg_query, r.beg_ref, r.n_ops, r.ref_span,
              [int(ops[k]) for k in range(min(r.n_ops, ops_cap))])


SWITCH_NAMES = {1: "U1 FO_SW_NO_SOFTCLIP_PAD", 2: "U4 FO_SW_END_LAST_COL", 4: "U5 FO_SW_E_BEFORE_F",
                8: "U5 FO_SW_GAP_TIE_OPEN", 16: "U7 FO_SW_EQ_BY_MATRIX", 32: "U3 FO_SW_SWAP_ID",
                64: "P1 FO_SW_WILD_MISMATCH"}


def fuzz_striped(seed: int, n_pairs: int, lanes: int, qmax: int = 72, tmax: int = 160,
                 params: Params | None = None, n_threads: int = 0) -> dict:
    """ps_fuzz: n_pairs generated pairs (random / planted / gapped / low-complexity / wildcard letters)
    through fo_sw_trace and the striped restatement; returns the census and the first divergence."""
    p = params or default_params()
    rep = FuzzReport()
    rc = lib().ps_fuzz(seed, n_pairs, lanes, qmax, tmax, C.byref(p), n_threads, C.byref(rep))
    if rc:
        raise ValueError("ps_fuzz failed")
    out = {k: getattr(rep, k) for k in ("n_pairs", "n_diverged", "n_gapped", "n_multi_max", "n_zero_ef", "first_div")}
    if rep.first_div >= 0:
        out["first_q"] = rep.first_q.decode()
        out["first_t"] = rep.first_t.decode()
        out["explained_by"] = SWITCH_NAMES.get(rep.explained_by, "no single switch")
    return out


NT16 = "=ACMGRSVTWYHKDBN"
_NT16_CODE = {c: i for i, c in enumerate(NT16)}


def pack_nt16(seq: str) -> np.ndarray:
    """ASCII -> BAM 4-bit packed (first base in the high nibble), like bam1_t."""
    codes = [_NT16_CODE.get(c.upper(), 15) for c in seq]
    if len(codes) & 1:
        codes.append(0)
    a = np.array(codes, dtype=np.uint8)
    return ((a[0::2] << 4) | a[1::2]).astype(np.uint8)


def revcomp_nt16(seq4: np.ndarray, l_qseq: int) -> str:
    out = C.create_string_buffer(l_qseq)
    seq4 = np.ascontiguousarray(seq4, dtype=np.uint8)
    lib().fo_revcomp_nt16(seq4.ctypes.data, l_qseq, out)
    return out.raw.decode()


def decode_nt16(seq4: np.ndarray, l_qseq: int) -> str:
    out = C.create_string_buffer(l_qseq)
    seq4 = np.ascontiguousarray(seq4, dtype=np.uint8)
    lib().fo_decode_nt16(seq4.ctypes.data, l_qseq, out)
    return out.raw.decode()


def parse_clips(cigar) -> tuple[int, int]:
    cg = np.ascontiguousarray(cigar, dtype=np.uint32)
    clips = (C.c_uint32 * 2)()
    lib().fo_parse_clips(cg.ctypes.data_as(C.POINTER(C.c_uint32)), len(cg), clips)
    return int(clips[0]) >> 4, int(clips[1]) >> 4


def ref_span(cigar) -> int:
    cg = np.ascontiguousarray(cigar, dtype=np.uint32)
    return int(lib().fo_cigar_ref_span(cg.ctypes.data_as(C.POINTER(C.c_uint32)), len(cg)))


def align_read(seq4, l_qseq, pos, aligned_len, clip_left, clip_right, ref_seq: bytes,
               params: Params | None = None, ops_cap: int = 64):
    p = params or default_params()
    seq4 = np.ascontiguousarray(seq4, dtype=np.uint8)
    r = ReadResult()
    ops = (C.c_uint32 * ops_cap)()
    rc = lib().fo_align_read(seq4.ctypes.data, l_qseq, pos, aligned_len, clip_left, clip_right,
                             ref_seq, len(ref_seq), C.byref(p), C.byref(r), ops, ops_cap)
    if rc:
        raise ValueError("fo_align_read failed")
    return r, [int(ops[k]) for k in range(min(r.sw.n_ops, ops_cap))]


def annotate_record(*, is_mapped: bool, has_sa: bool, cigar, seq4, qual, l_qseq: int, pos: int,
                    contig_name: str, ref_seq: bytes, params: Params | None = None) -> dict:
    """anno.d:55-110 for one record -> {'rs': int, 'am':..., 'as':..., 'ar':..., 'ab':...}."""
    p = params or default_params()
    cg = np.ascontiguousarray(cigar, dtype=np.uint32)
    s4 = np.ascontiguousarray(seq4, dtype=np.uint8)
    ql = np.ascontiguousarray(qual, dtype=np.uint8)
    rec = Record(int(is_mapped), int(has_sa), cg.ctypes.data_as(C.POINTER(C.c_uint32)), len(cg),
                 s4.ctypes.data_as(C.POINTER(C.c_uint8)), ql.ctypes.data_as(C.POINTER(C.c_uint8)),
                 l_qseq, pos, contig_name.encode(), ref_seq, len(ref_seq))
    t = Tags()
    rc = lib().fo_annotate_record(C.byref(rec), C.byref(p), C.byref(t))
    if rc:
        raise ValueError("fo_annotate_record failed")
    out = {"rs": int(t.rs)}
    if t.has_tags:
        out["am"] = C.string_at(t.am).decode()
        out["as"] = C.string_at(t.as_).decode()
        out["ar"] = C.string_at(t.ar).decode()
        out["ab"] = C.string_at(t.ab).decode()
    lib().fo_free_tags(C.byref(t))
    return out


def align_batch(seq4, seq_off, l_qseq, tid, pos, aligned_len, clip_left, clip_right, contigs: list[bytes],
                params: Params | None = None, ops_cap: int = 32, n_threads: int = 0, simd: bool = False):
    """fo_align_read over struct-of-arrays inputs; returns (structured results array, ops[n, ops_cap])."""
    p = params or default_params()
    n = len(l_qseq)
    seq4 = np.ascontiguousarray(seq4, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
    l_qseq = np.ascontiguousarray(l_qseq, dtype=np.int32)
    tid = np.ascontiguousarray(tid, dtype=np.int32)
    pos = np.ascontiguousarray(pos, dtype=np.int64)
    aligned_len = np.ascontiguousarray(aligned_len, dtype=np.int32)
    clip_left = np.ascontiguousarray(clip_left, dtype=np.int32)
    clip_right = np.ascontiguousarray(clip_right, dtype=np.int32)
    names = (C.c_char_p * len(contigs))()
    for k_, c_ in enumerate(contigs):      # bytes, or uint8 numpy arrays passed without a copy
        if isinstance(c_, np.ndarray):
            assert c_.dtype == np.uint8 and c_.flags["C_CONTIGUOUS"]
            C.cast(names, C.POINTER(C.c_void_p))[k_] = c_.ctypes.data
        else:
            names[k_] = c_
    clen = np.array([len(c) for c in contigs], dtype=np.int64)
    dt = np.dtype([("aligned", "<i4"), ("art_left", "<i4"), ("art_right", "<i4"), ("_pad0", "<i4"),
                   ("win_start", "<i8"), ("tlen", "<i4"), ("score", "<i4"), ("end_query", "<i4"),
                   ("end_ref", "<i4"), ("beg_query", "<i4"), ("beg_ref", "<i4"), ("n_ops", "<i4"),
                   ("ref_span", "<i4")])
    assert dt.itemsize == C.sizeof(ReadResult), (dt.itemsize, C.sizeof(ReadResult))
    arr = np.zeros(n, dtype=dt)
    ops = np.zeros((n, ops_cap), dtype=np.uint32)
    fn = lib().fo_align_batch_simd if simd else lib().fo_align_batch
    rc = fn(n, seq4.ctypes.data, seq_off.ctypes.data, l_qseq.ctypes.data, tid.ctypes.data,
                              pos.ctypes.data, aligned_len.ctypes.data, clip_left.ctypes.data,
                              clip_right.ctypes.data, len(contigs), names, clen.ctypes.data, C.byref(p),
                              arr.ctypes.data, ops.ctypes.data, ops_cap, n_threads)
    if rc:
        raise ValueError("fo_align_batch failed")
    return arr, ops
