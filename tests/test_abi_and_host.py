"""CPU-side checks: the C-ABI library loads and exports every symbol the headers declare, fails
loudly without a GPU (no fallback), and the host-side mirror (fadehost_*) reproduces the oracle's
record-level tags when fed the oracle's alignment results."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from fade_b200 import _lib, api, sim
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fade(?:gpu|host)_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_functions("fadegpu.h") + declared_functions("fadehost.h")
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(names) == sorted(_lib.ABI_SYMBOLS)
    assert L.fadegpu_abi_version() == 2


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Params) == 48
    assert C.sizeof(_lib.Inputs) == 8 * 8
    assert C.sizeof(_lib.BatchView) == 8 * 21
    assert api.META_DTYPE.itemsize == 32                              # fadegpu_read_meta
    assert C.sizeof(_lib.Stats) == 120
    assert C.sizeof(_lib.Result) == 32 + 4 * 10                      # fadegpu_result, FADEGPU_MAX_OPS == 10
    p = api.default_params()
    assert (p.window_size, p.min_length, p.gap_open, p.gap_extend, p.match, p.mismatch) == (300, 5, 10, 2, 2, -3)


def test_no_gpu_means_loud_failure_not_fallback():
    """On a box without a CUDA device every compute entry point must fail; with one it must work."""
    try:
        n = api.device_count()
    except _lib.FadeGpuError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present: covered by the -m gpu tests")
    with pytest.raises(_lib.FadeGpuError) as e:
        api.Context(0)
    assert e.value.code in (-5, -2)
    L = _lib.lib()
    assert L.fadegpu_submit(None, None, 0) < 0 and L.fadegpu_wait(None, None) < 0
    assert L.fadegpu_load_reference(None, 0, None, None, None) < 0
    assert b"" != L.fadegpu_last_error(None)


def test_host_clip_parsing_matches_oracle():
    L = _lib.lib()
    for cg in ("12S40M", "5H3S10M2I4M7S2H", "10M", "3S10M", "10M4S", "2H10M", "1S1M1S", "3S10M2I4M3D5N7S"):
        ops = orc.cigar_from_string(cg)
        clips = (C.c_uint32 * 2)()
        L.fadehost_parse_clips(ops.ctypes.data_as(C.POINTER(C.c_uint32)), len(ops), clips)
        assert (clips[0] >> 4, clips[1] >> 4) == orc.parse_clips(ops), cg
        assert L.fadehost_aligned_length(ops.ctypes.data_as(C.POINTER(C.c_uint32)), len(ops)) == orc.ref_span(ops)


def _records(n):
    names, contigs, cfg, _ = sim.config_c1()
    rd = sim.make_reads(cfg, 0, n, contigs)
    return names, contigs, rd


def test_host_tag_assembly_matches_oracle_and_golden():
    """fadehost_prepare/finish fed with the ORACLE's alignment results must give the oracle's tags
    (this isolates the host logic of anno.d:94-107 / analysis.d:82-118 from the device)."""
    L = _lib.lib()
    names, contigs, rd = _records(1500)
    refb = contigs[0].tobytes()
    res, ops = orc.align_batch(rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left,
                               rd.clip_right, [refb])
    with open(os.path.join(ROOT, "tests", "golden", "c1_head.json")) as f:
        golden = json.load(f)["tags"]
    Lr = rd.read_len
    stride = (Lr + 1) // 2
    n_art = 0
    for k in range(rd.n):
        cg = np.ascontiguousarray(rd.cigar[k, : rd.n_cigar[k]])
        s4 = np.ascontiguousarray(rd.seq4[k * stride:(k + 1) * stride])
        ql = np.ascontiguousarray(rd.qual[k * Lr:(k + 1) * Lr])
        hr = _lib.HostRecord(int(rd.flag[k]), int(rd.has_sa[k]), cg.ctypes.data_as(C.POINTER(C.c_uint32)), len(cg),
                             s4.ctypes.data_as(C.POINTER(C.c_uint8)), ql.ctypes.data_as(C.POINTER(C.c_uint8)), Lr,
                             int(rd.tid[k]), int(rd.pos[k]))
        al, cl, cr, rs = C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint8()
        go = L.fadehost_prepare(C.byref(hr), C.byref(al), C.byref(cl), C.byref(cr), C.byref(rs))
        if go:
            assert (al.value, cl.value, cr.value) == (rd.aligned_len[k], rd.clip_left[k], rd.clip_right[k])
        flags = int(res["aligned"][k]) | (int(res["art_left"][k]) << 1) | (int(res["art_right"][k]) << 2)
        bufs = [C.create_string_buffer(1024) for _ in range(4)]
        rs_out = C.c_uint8()
        o = np.ascontiguousarray(ops[k])
        rc = L.fadehost_finish(C.byref(hr), names[0].encode(), rs, cl, cr, al, flags, int(res["win_start"][k]),
                               int(res["beg_ref"][k]), int(res["n_ops"][k]), o.ctypes.data_as(C.POINTER(C.c_uint32)),
                               C.byref(rs_out), *bufs, 1024)
        got = {"rs": rs_out.value}
        if rc == 1:
            got.update(am=bufs[0].value.decode(), ar=bufs[2].value.decode(), ab=bufs[3].value.decode())
            got["as"] = bufs[1].value.decode()
            n_art += 1
        exp = orc.annotate_record(is_mapped=not (rd.flag[k] & 4), has_sa=bool(rd.has_sa[k]), cigar=cg, seq4=s4, qual=ql,
                                  l_qseq=Lr, pos=int(rd.pos[k]), contig_name=names[0], ref_seq=refb)
        assert got == exp, (k, got, exp)
        if exp["rs"] != 0:
            assert golden[str(k)] == exp
        else:
            assert str(k) not in golden
    assert n_art > 100


def test_finish_reports_small_buffers():
    L = _lib.lib()
    ops = orc.cigar_from_string("126S24=")
    s4 = orc.pack_nt16("A" * 150)
    ql = np.full(150, 30, np.uint8)
    cg = orc.cigar_from_string("30S120M")
    hr = _lib.HostRecord(0, 0, cg.ctypes.data_as(C.POINTER(C.c_uint32)), 2, s4.ctypes.data_as(C.POINTER(C.c_uint8)),
                         ql.ctypes.data_as(C.POINTER(C.c_uint8)), 150, 0, 1000)
    bufs = [C.create_string_buffer(8) for _ in range(4)]
    rs_out = C.c_uint8()
    rc = L.fadehost_finish(C.byref(hr), b"chr1", 1, 30, 0, 120, 1 | 2, 700, 10, 2,
                           ops.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(rs_out), *bufs, 8)
    assert rc == -1 and rs_out.value == 3


def test_d_binding_matches_the_c_abi():
    """integration/fadegpu.d (the extern(C) binding the D host links with) cannot be compiled here; at least every
    function it declares must be exported by the library, and the struct sizes it pins must be the C ones."""
    import re
    txt = open(os.path.join(ROOT, "integration", "fadegpu.d")).read()
    L = _lib.lib()
    fns = set(re.findall(r"\b(fade(?:gpu|host)_\w+)\s*\(", txt))
    assert len(fns) >= 20
    missing = [f for f in fns if not hasattr(L, f)]
    assert not missing, missing
    sizes = dict(re.findall(r"static assert\((\w+)\.sizeof == (\d+)\);", txt))
    c_sizes = {"fadegpu_params": C.sizeof(_lib.Params), "fadegpu_read_meta": api.META_DTYPE.itemsize,
               "fadegpu_batch_view": C.sizeof(_lib.BatchView), "fadegpu_result": C.sizeof(_lib.Result),
               "fadegpu_stats": C.sizeof(_lib.Stats)}
    assert {k: int(v) for k, v in sizes.items()} == c_sizes
    assert f"enum FADEGPU_ABI_VERSION = {L.fadegpu_abi_version()};" in txt and f"enum FADEGPU_MAX_OPS = {api.MAX_OPS};" in txt
    # the loop calls only what the binding declares
    loop = open(os.path.join(ROOT, "integration", "anno.d")).read()
    used = set(re.findall(r"\b(fade(?:gpu|host)_[a-z_]+)\s*\(", loop))
    assert used and used <= fns, used - fns
