"""Python host layer over the C ABI: Context / Batch wrappers and an `annotate_records` loop that
mirrors source/anno.d:44-50 + annotateTask (source/anno.d:55-110) in batched form."""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass, field

import numpy as np

from ._lib import BatchView, FadeGpuError, HostRecord, Inputs, Params, Result, ResultsView, Stats, lib

MAX_OPS = 10
R_ALIGNED, R_ART_LEFT, R_ART_RIGHT, R_OPS_TRUNC, R_GENERIC, R_OVERSIZE, R_SCORE_ONLY = 1, 2, 4, 8, 16, 32, 64
F_FORCE_GENERIC = 1
F_NO_SCATTER = 2
F_TAGS_ONLY = 4
F_NO_SHORTCUT = 8
F_HOST_BINNING = 16
F_SYNC_SUBMIT = 32
RESULT_DTYPE = np.dtype([("score", "<i4"), ("end_query", "<i4"), ("end_ref", "<i4"), ("beg_query", "<i4"),
                         ("beg_ref", "<i4"), ("n_ops", "<i4"), ("flags", "<u4"), ("read", "<i4"),
                         ("ops", "<u4", (MAX_OPS,))])
# fadegpu_read_meta: one read of the compact input layout (fadegpu_submit_compact)
META_DTYPE = np.dtype([("pos", "<i8"), ("seq_off", "<u4"), ("l_qseq", "<i4"), ("tid", "<i4"), ("aligned_len", "<i4"),
                       ("clip_left", "<u4"), ("clip_right", "<u4")])
assert META_DTYPE.itemsize == 32
OPCHARS = "MIDNSHP=XB"


def default_params(**kw) -> Params:
    p = Params()
    rc = lib().fadegpu_default_params(C.byref(p))
    if rc:
        raise FadeGpuError(rc, "fadegpu_default_params")
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().fadegpu_device_count(C.byref(n))
    if rc:
        raise FadeGpuError(rc, lib().fadegpu_last_error(None).decode())
    return n.value


def _np_view(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    addr = C.cast(ptr, C.c_void_p).value
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype, count=n)


class Context:
    """One per (host thread, GPU): scoring profile + device-resident packed reference.
    Replaces Parasail("ACTGN",10,2,2,-3) (anno.d:36) and IndexedFastaFile (anno.d:23)."""

    def __init__(self, device: int = 0, params: Params | None = None):
        self._h = C.c_void_p()
        self.params = params or default_params()
        rc = lib().fadegpu_create(device, C.byref(self.params), C.byref(self._h))
        if rc:
            raise FadeGpuError(rc, lib().fadegpu_last_error(None).decode())
        self.device = device
        self.contig_names: list[str] = []
        self._batches = weakref.WeakSet()      # freed before the ctx: a batch must not outlive it

    def _check(self, rc: int):
        if rc:
            raise FadeGpuError(rc, lib().fadegpu_last_error(self._h).decode())

    def load_reference(self, names: list[str], seqs: list):
        """seqs: ASCII contigs as bytes or as C-contiguous uint8 numpy arrays (passed without a copy)."""
        n = len(seqs)
        cn = (C.c_char_p * n)(*[s.encode() for s in names])
        cs = (C.c_char_p * n)()
        for k, s in enumerate(seqs):
            if isinstance(s, np.ndarray):
                assert s.dtype == np.uint8 and s.flags["C_CONTIGUOUS"]
                C.cast(cs, C.POINTER(C.c_void_p))[k] = s.ctypes.data
            else:
                cs[k] = s
        ln = (C.c_int64 * n)(*[len(s) for s in seqs])
        self._check(lib().fadegpu_load_reference(self._h, n, cn, ln, cs))
        self.contig_names = list(names)

    def share_reference_from(self, other: "Context"):
        self._check(lib().fadegpu_share_reference(self._h, other._h))
        self.contig_names = list(other.contig_names)

    def reference_info(self):
        a, b, c = C.c_int32(), C.c_int64(), C.c_int64()
        self._check(lib().fadegpu_reference_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def measure_alu_peak(self):
        ops, mhz = C.c_double(), C.c_double()
        self._check(lib().fadegpu_measure_alu_peak(self._h, C.byref(ops), C.byref(mhz)))
        return ops.value, mhz.value

    def replay_batches(self, batches, iters: int = 1) -> float:
        """fadegpu_replay_batches: device ms of `iters` passes of only the kernels over resident batches,
        queued back to back as consecutive submits queue them."""
        arr = (C.c_void_p * len(batches))(*[b._h.value for b in batches])
        ms = C.c_float()
        self._check(lib().fadegpu_replay_batches(self._h, arr, len(batches), iters, C.byref(ms)))
        return ms.value

    def alloc_batch(self, max_reads: int, max_seq_bytes: int) -> "Batch":
        return Batch(self, max_reads, max_seq_bytes)

    def close(self):
        if self._h:
            for b in list(self._batches):
                b.close()
            lib().fadegpu_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """Pinned struct-of-arrays buffers (numpy views) + device mirrors, owned by the library."""

    def __init__(self, ctx: Context, max_reads: int, max_seq_bytes: int):
        self.ctx = ctx
        self._h = C.c_void_p()
        ctx._check(lib().fadegpu_alloc_batch(ctx._h, max_reads, max_seq_bytes, C.byref(self._h)))
        ctx._batches.add(self)
        v = BatchView()
        ctx._check(lib().fadegpu_get_batch_view(self._h, C.byref(v)))
        self.max_reads, self.max_seq_bytes = max_reads, max_seq_bytes
        n = max_reads
        self.seq4 = _np_view(v.seq4, max_seq_bytes, np.uint8)
        self.seq_off = _np_view(v.seq_off, n + 1, np.int64)
        self.l_qseq = _np_view(v.l_qseq, n, np.int32)
        self.tid = _np_view(v.tid, n, np.int32)
        self.pos = _np_view(v.pos, n, np.int64)
        self.aligned_len = _np_view(v.aligned_len, n, np.int32)
        self.clip_left = _np_view(v.clip_left, n, np.int32)
        self.clip_right = _np_view(v.clip_right, n, np.int32)
        self.flags = _np_view(v.flags, n, np.uint8)
        self.gate = _np_view(v.gate, n, np.uint8)
        self.meta = _np_view(C.cast(v.meta, C.POINTER(C.c_uint8)), n * META_DTYPE.itemsize, np.uint8).view(META_DTYPE)
        self.seq_bytes = 0
        if v.score:     # per-read output arrays exist unless the ctx has F_NO_SCATTER
            self.score = _np_view(v.score, n, np.int32)
            self.beg_query = _np_view(v.beg_query, n, np.int32)
            self.end_query = _np_view(v.end_query, n, np.int32)
            self.beg_ref = _np_view(v.beg_ref, n, np.int32)
            self.end_ref = _np_view(v.end_ref, n, np.int32)
            self.win_start = _np_view(v.win_start, n, np.int64)
            self.n_ops = _np_view(v.n_ops, n, np.int32)
            self.ops = _np_view(v.ops, n * MAX_OPS, np.uint32).reshape(n, MAX_OPS)
        else:
            self.score = self.beg_query = self.end_query = self.beg_ref = self.end_ref = None
            self.win_start = self.n_ops = self.ops = None
        self.n = 0

    def fill(self, seq4, seq_off, l_qseq, tid, pos, aligned_len, clip_left, clip_right):
        n = len(l_qseq)
        if n > self.max_reads or int(seq_off[n]) > self.max_seq_bytes:
            raise ValueError("batch too small")
        self.seq4[:int(seq_off[n])] = seq4[:int(seq_off[n])]
        self.seq_off[:n + 1] = seq_off[:n + 1]
        self.l_qseq[:n] = l_qseq
        self.tid[:n] = tid
        self.pos[:n] = pos
        self.aligned_len[:n] = aligned_len
        self.clip_left[:n] = clip_left
        self.clip_right[:n] = clip_right
        self.n = n
        return self

    def fill_compact(self, seq4, seq_off, l_qseq, tid, pos, aligned_len, clip_left, clip_right):
        """the same reads in the compact layout: gate[] (one byte per read) + meta[] (32-byte records) + seq4"""
        n = len(l_qseq)
        nb = int(seq_off[n])
        if n > self.max_reads or nb > self.max_seq_bytes:
            raise ValueError("batch too small")
        self.seq4[:nb] = seq4[:nb]
        m = self.meta[:n]
        m["pos"] = pos
        m["seq_off"] = seq_off[:n]
        m["l_qseq"] = l_qseq
        m["tid"] = tid
        m["aligned_len"] = aligned_len
        m["clip_left"] = np.asarray(clip_left).astype(np.uint32)
        m["clip_right"] = np.asarray(clip_right).astype(np.uint32)
        np.minimum(np.maximum(m["clip_left"], m["clip_right"]), 255, out=self.gate[:n], casting="unsafe")
        self.n = n
        self.seq_bytes = nb
        return self

    def submit_compact(self, n: int | None = None, seq_bytes: int | None = None):
        if n is not None:
            self.n = n
        if seq_bytes is not None:
            self.seq_bytes = seq_bytes
        self.ctx._check(lib().fadegpu_submit_compact(self.ctx._h, self._h, self.n, self.seq_bytes))

    def submit(self, n: int | None = None):
        if n is not None:
            self.n = n
        self.ctx._check(lib().fadegpu_submit(self.ctx._h, self._h, self.n))

    def submit_arrays(self, n, seq4, seq_off, l_qseq, tid, pos, aligned_len, clip_left, clip_right):
        """fadegpu_submit_inputs: read the inputs straight from caller-owned (pageable) numpy arrays.
        seq_off may hold absolute offsets into seq4; arrays must be C-contiguous with the ABI dtypes."""
        def ptr(a, dt):
            assert a.dtype == dt and a.flags["C_CONTIGUOUS"], (a.dtype, dt)
            return a.ctypes.data
        inp = Inputs(ptr(seq4, np.uint8), ptr(seq_off, np.int64), ptr(l_qseq, np.int32), ptr(tid, np.int32),
                     ptr(pos, np.int64), ptr(aligned_len, np.int32), ptr(clip_left, np.int32),
                     ptr(clip_right, np.int32))
        self.n = n
        self.ctx._check(lib().fadegpu_submit_inputs(self.ctx._h, self._h, n, C.byref(inp)))

    def wait(self):
        self.ctx._check(lib().fadegpu_wait(self.ctx._h, self._h))

    def run(self, n: int | None = None):
        self.submit(n)
        self.wait()
        return self

    def results(self):
        """fadegpu_get_results: (records[n_results] structured array, win_start[n_results], result_index[n])
        as zero-copy numpy views valid until the next submit."""
        rv = ResultsView()
        self.ctx._check(lib().fadegpu_get_results(self._h, C.byref(rv)))
        k = int(rv.n_results)
        assert RESULT_DTYPE.itemsize == C.sizeof(Result)
        if k == 0:
            rec = np.zeros(0, dtype=RESULT_DTYPE)
            ws = np.zeros(0, dtype=np.int64)
        else:
            addr = C.cast(rv.results, C.c_void_p).value
            rec = np.frombuffer((C.c_char * (k * RESULT_DTYPE.itemsize)).from_address(addr), dtype=RESULT_DTYPE, count=k)
            ws = _np_view(rv.win_start, k, np.int64)
        return rec, ws, _np_view(rv.result_index, self.n, np.int32)

    def stats(self) -> Stats:
        s = Stats()
        self.ctx._check(lib().fadegpu_get_stats(self._h, C.byref(s)))
        return s

    def timeline(self, origin: "Batch"):
        """device ms of (uploads start, first fill starts, kernels done, results home, inputs ready, last fill done) relative to origin's start"""
        ms = (C.c_float * 6)()
        self.ctx._check(lib().fadegpu_get_timeline(self._h, origin._h, ms))
        return [float(x) for x in ms]

    def replay_kernels(self, iters: int = 1) -> float:
        ms = C.c_float()
        self.ctx._check(lib().fadegpu_replay_kernels(self.ctx._h, self._h, iters, C.byref(ms)))
        return ms.value

    def cigar(self, r: int) -> str:
        k = min(int(self.n_ops[r]), MAX_OPS)
        return "".join(f"{int(o) >> 4}{OPCHARS[int(o) & 0xf]}" for o in self.ops[r, :k])

    def close(self):
        if self._h:
            lib().fadegpu_free_batch(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class Record:
    """The fields of a SAM/BAM record that annotateTask reads (dhtslib SAMRecord subset)."""
    qname: str
    flag: int
    tid: int
    pos: int                 # 0-based
    cigar: np.ndarray        # uint32 BAM-encoded
    seq4: np.ndarray         # uint8 BAM packed bases
    qual: np.ndarray         # uint8 raw phred
    l_qseq: int
    has_sa: bool = False
    tags: dict = field(default_factory=dict)   # filled by annotate_records: rs, am, as, ar, ab


def _host_record(rec: Record):
    cg = np.ascontiguousarray(rec.cigar, dtype=np.uint32)
    s4 = np.ascontiguousarray(rec.seq4, dtype=np.uint8)
    ql = np.ascontiguousarray(rec.qual, dtype=np.uint8)
    hr = HostRecord(rec.flag, int(rec.has_sa), cg.ctypes.data_as(C.POINTER(C.c_uint32)), len(cg),
                    s4.ctypes.data_as(C.POINTER(C.c_uint8)), ql.ctypes.data_as(C.POINTER(C.c_uint8)),
                    rec.l_qseq, rec.tid, rec.pos)
    return hr, (cg, s4, ql)


def annotate_records(ctx: Context, records: list[Record], batch: Batch | None = None) -> list[Record]:
    """Batched equivalent of `foreach(rec; parallel(bam.allRecords)) annotateTask(...)`
    (source/anno.d:44-50): every record gets tags['rs'], artifact records also am/as/ar/ab."""
    L = lib()
    n = len(records)
    total_seq = sum((r.l_qseq + 1) // 2 for r in records)
    own = batch is None
    if own:
        batch = ctx.alloc_batch(max(n, 1), max(total_seq, 16))
    hrs, keep, prep = [], [], []
    off = 0
    for k, rec in enumerate(records):
        hr, refs = _host_record(rec)
        hrs.append(hr)
        keep.append(refs)
        al, cl, cr, rs = C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint8()
        L.fadehost_prepare(C.byref(hr), C.byref(al), C.byref(cl), C.byref(cr), C.byref(rs))
        prep.append((al.value, cl.value, cr.value, rs.value))
        nb = (rec.l_qseq + 1) // 2
        batch.seq_off[k] = off
        batch.seq4[off:off + nb] = refs[1][:nb]
        off += nb
        batch.l_qseq[k] = rec.l_qseq
        batch.tid[k] = rec.tid
        batch.pos[k] = rec.pos
        batch.aligned_len[k] = al.value
        batch.clip_left[k] = cl.value
        batch.clip_right[k] = cr.value
    batch.seq_off[n] = off
    batch.run(n)
    for k, rec in enumerate(records):
        al, cl, cr, rs = prep[k]
        cap = 4 * rec.l_qseq + 256 + (len(ctx.contig_names[rec.tid]) if 0 <= rec.tid < len(ctx.contig_names) else 0)
        am, as_, ar, ab = (C.create_string_buffer(cap) for _ in range(4))
        rs_out = C.c_uint8()
        name = ctx.contig_names[rec.tid].encode() if 0 <= rec.tid < len(ctx.contig_names) else b""
        ops = np.ascontiguousarray(batch.ops[k])
        rc = L.fadehost_finish(C.byref(hrs[k]), name, rs, cl, cr, al, int(batch.flags[k]), int(batch.win_start[k]),
                               int(batch.beg_ref[k]), int(batch.n_ops[k]), ops.ctypes.data_as(C.POINTER(C.c_uint32)),
                               C.byref(rs_out), am, as_, ar, ab, cap)
        if rc < 0:
            raise FadeGpuError(rc, "fadehost_finish: buffer too small")
        rec.tags = {"rs": rs_out.value}
        if rc == 1:
            rec.tags.update(am=am.value.decode(), ar=ar.value.decode(), ab=ab.value.decode())
            rec.tags["as"] = as_.value.decode()
    if own:
        batch.close()
    return records
