"""The oracle against the known-answer vectors of SURVEY.md 8(c) (hand-derived from the published
parasail rules; the reference ships no tests, so parity is UNPINNED by reference vectors), against
the second independent Python restatement, and against the committed golden fixture."""
import json
import os
import random

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle as orc
from oracle.py_sw import sw_trace_py

HERE = os.path.dirname(os.path.abspath(__file__))


def test_k1_exact_suffix():
    r = orc.sw_trace("AAAAAAAAAACATTAGCCAT", "GGGGGCATTAGCCATGGGGG")
    assert (r.score, r.end_query, r.end_ref, r.beg_query, r.beg_ref, r.cigar) == (20, 19, 14, 10, 5, "10S10=")


def test_k2_one_mismatch():
    r = orc.sw_trace("AAAAAAAAAACATTAGCCAT", "GGGGGCATTTGCCATGGGGG")
    assert (r.score, r.end_query, r.end_ref, r.beg_ref, r.cigar) == (15, 19, 14, 5, "10S4=1X5=")


def test_k3_column_tie_takes_first_column():
    r = orc.sw_trace("ACGTACGT", "ACGTACGTCCACGTACGT")
    assert (r.score, r.end_ref, r.beg_ref, r.cigar) == (16, 7, 0, "8=")


def test_k4_row_tie_takes_first_row():
    r = orc.sw_trace("ACGCCACG", "TTACGTT")
    assert (r.score, r.end_ref, r.end_query, r.beg_ref, r.cigar) == (6, 4, 2, 2, "3=5S")


def test_k5_n_is_a_symbol_and_wildcards_score_zero():
    assert orc.sw_trace("ACGNNACG", "ACGNNACG").score == 16
    # any other letter is the wildcard row/column: score 0, but '=' by byte equality (U7)
    r = orc.sw_trace("ACGTRACGT", "ACGTRACGT")
    assert r.score == 16 and r.cigar == "9="
    r = orc.sw_trace("ACGTRACGT", "ACGTYACGT")
    assert r.score == 16 and r.cigar == "4=1X4="


def test_k6_cutoff_boundary_and_integer_equivalence():
    # score 18 with clip_len 10 is rejected (18 > 18.0 false), score 19 accepted (analysis.d:43,76)
    for clip in range(0, 3000):
        cutoff = np.float32(clip * 0.9 * 2)
        for score in (int(1.8 * clip) - 1, int(1.8 * clip), int(1.8 * clip) + 1, int(1.8 * clip) + 2):
            assert (np.float32(score) > cutoff) == (5 * score > 9 * clip), (clip, score)
    assert not np.float32(18) > np.float32(10 * 0.9 * 2)
    assert np.float32(19) > np.float32(10 * 0.9 * 2)


def test_k7_reverse_complement_table():
    assert orc.revcomp_nt16(orc.pack_nt16("ACGTN"), 5) == "NACGT"
    # IUPAC complements of util.d:18-21 (R<->Y, K<->M, B<->V, D<->H, S, W fixed)
    assert orc.revcomp_nt16(orc.pack_nt16("RYKMBVDHSW="), 11) == "=WSDHBVKMRY"


def test_k8_rs_values_and_clip_parsing():
    ref = ("G" * 400 + "CATTAGCCATAC" + "T" * 400).encode()
    # read = RC("CATTAGCCATAC") as a 12-base left clip followed by 40 matching bases at pos 500
    clip = "GTATGGCTAATG"
    read = clip + "T" * 40
    cigar = orc.cigar_from_string("12S40M")
    t = orc.annotate_record(is_mapped=True, has_sa=True, cigar=cigar, seq4=orc.pack_nt16(read),
                            qual=np.full(52, 30, np.uint8), l_qseq=52, pos=500, contig_name="chrT", ref_seq=ref)
    assert t["rs"] == 1 + 2 + 32                      # sc + art_left + sup
    assert t["am"].startswith("chrT,400,") and t["am"].endswith(";")
    assert orc.parse_clips(orc.cigar_from_string("5H3S10M2I4M7S2H")) == (3, 7)
    assert orc.parse_clips(orc.cigar_from_string("10M")) == (0, 0)
    assert orc.ref_span(orc.cigar_from_string("3S10M2I4M3D5N7S")) == 22
    # unmapped / unclipped records: rs = 0 and nothing else (anno.d:61-65)
    t = orc.annotate_record(is_mapped=False, has_sa=True, cigar=cigar, seq4=orc.pack_nt16(read),
                            qual=np.full(52, 30, np.uint8), l_qseq=52, pos=500, contig_name="chrT", ref_seq=ref)
    assert t == {"rs": 0}


def test_gap_tie_breaks():
    # a 1-base deletion in the query: D consumes target; gap open 10, so it needs long flanks
    left, right = "ACGTTGCAAGGCTTAACCGGTTAAGGAG", "TTGACCAGTACCGGATATTCCGGAACCA"
    r = orc.sw_trace(left + right, "GG" + left + "C" + right + "GG")
    assert r.cigar == f"{len(left)}=1D{len(right)}=" and r.score == 2 * (len(left) + len(right)) - 10
    r = orc.sw_trace(left + "C" + right, "GG" + left + right + "GG")
    assert r.cigar == f"{len(left)}=1I{len(right)}="
    # ambiguous gap position: DIAG has priority during the traceback (walking backwards), so the
    # gap ends up at the LEFT end of the homopolymer (P4)
    r = orc.sw_trace(left[:-2] + "CC" + right, "GG" + left[:-2] + "CCC" + right + "GG")
    assert r.cigar == f"{len(left) - 2}=1D{len(right) + 2}="


@settings(max_examples=300, deadline=None)
@given(st.text(alphabet="ACGTN", min_size=1, max_size=40), st.text(alphabet="ACGTN", min_size=1, max_size=70))
def test_c_oracle_equals_python_restatement(q, t):
    r, p = orc.sw_trace(q, t), sw_trace_py(q, t)
    assert (r.score, r.end_query, r.end_ref, r.beg_query, r.beg_ref, r.cigar if r.score else "") == \
        (p["score"], p["end_query"], p["end_ref"], p["beg_query"], p["beg_ref"], p["cigar"])


def test_c_oracle_equals_python_restatement_planted():
    rng = random.Random(3)
    for _ in range(300):
        t = "".join(rng.choice("ACGT") for _ in range(rng.randint(20, 120)))
        a = rng.randint(0, len(t) - 10)
        frag = list(t[a:a + rng.randint(8, 60)])
        for _ in range(rng.randint(0, 3)):
            k = rng.randrange(len(frag))
            frag[k:k + 1] = rng.choice([[], [rng.choice("ACGT")], [frag[k], rng.choice("ACGT")]])
        q = "".join(rng.choice("ACGT") for _ in range(rng.randint(0, 8))) + "".join(frag)
        if not q:
            continue
        r, p = orc.sw_trace(q, t), sw_trace_py(q, t)
        assert (r.score, r.end_query, r.end_ref, r.beg_query, r.beg_ref, r.cigar) == \
            (p["score"], p["end_query"], p["end_ref"], p["beg_query"], p["beg_ref"], p["cigar"])


def test_switches_change_what_they_name():
    p = orc.default_params(switches=1)         # U1 off: no S padding
    assert orc.sw_trace("AAAAAAAAAACATTAGCCAT", "GGGGGCATTAGCCATGGGGG", p).cigar == "10="
    p = orc.default_params(switches=2)         # U4 flipped: last column among maxima
    assert orc.sw_trace("ACGTACGT", "ACGTACGTCCACGTACGT", p).end_ref == 17


def test_golden_fixture_c1_head():
    """tests/golden/c1_head.json (made by tests/golden/make_golden.py from the oracle + simulator)
    pins the oracle and the simulator against silent drift."""
    from fade_b200 import sim
    with open(os.path.join(HERE, "golden", "c1_head.json")) as f:
        g = json.load(f)
    names, contigs, cfg, _ = sim.config_c1()
    rd = sim.make_reads(cfg, 0, g["n_reads"], contigs)
    res, ops = orc.align_batch(rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left,
                               rd.clip_right, [c.tobytes() for c in contigs])
    al = np.where(res["aligned"] == 1)[0]
    assert al.tolist() == g["aligned_reads"]
    for k, e in zip(al, g["alignments"]):
        got = [int(res[f][k]) for f in ("score", "beg_query", "end_query", "beg_ref", "end_ref", "art_left", "art_right")]
        got.append(orc.cigar_string(ops[k, : res["n_ops"][k]]))
        got.append(int(res["win_start"][k]))
        assert got == e, (int(k), got, e)


@pytest.mark.parametrize("width", ["widest", "avx2"])
def test_simd_cpu_baseline_equals_scalar_oracle(width, monkeypatch):
    """oracle/fade_oracle_simd.c (the AVX2 / AVX-512BW CPU baseline timed by bench.py) must reproduce the scalar
    oracle bit for bit at both lane counts: simulated reads, ragged lengths / contig ends, wildcard letters, other flags."""
    import random as _random

    if width == "avx2":
        monkeypatch.setenv("FADE_ORACLE_SIMD", "avx2")
        assert orc.simd_lanes() in (0, 16)
    else:
        monkeypatch.delenv("FADE_ORACLE_SIMD", raising=False)
        assert orc.simd_lanes() in (0, 16, 32)

    import readsets
    from fade_b200 import sim

    def same(rd, contigs, prm):
        a, oa = orc.align_batch(rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left,
                                rd.clip_right, contigs, params=prm)
        b, ob = orc.align_batch(rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left,
                                rd.clip_right, contigs, params=prm, simd=True, n_threads=3)
        for f in a.dtype.names:
            assert np.array_equal(a[f], b[f]), f
        k = np.minimum(a["n_ops"], 32)
        m = np.arange(32)[None, :] < k[:, None]
        assert np.array_equal(np.where(m, oa, 0), np.where(m, ob, 0))
        return int(a["aligned"].sum())

    names, contigs, cfg, _ = sim.config_c1()
    rd = sim.make_reads(cfg, 0, 3000, contigs)
    assert same(rd, [c.tobytes() for c in contigs], orc.default_params()) > 300
    assert same(rd, [c.tobytes() for c in contigs], orc.default_params(min_length=20, window_size=40)) > 50
    rng = _random.Random(21)
    base = bytearray(readsets.random_ref(rng, 5000))
    for _ in range(30):
        base[rng.randrange(len(base))] = ord(rng.choice("RYKMNn"))
    cs = [bytes(base), readsets.random_ref(rng, 400)]
    rs = readsets.build(readsets.ragged_reads(rng, cs, 1200, wild_read_rate=0.002, alpha="ACGTN"))
    assert same(rs, cs, orc.default_params()) > 200
    assert same(rs, cs, orc.default_params(gap_open=12, gap_extend=3, match=9, mismatch=-9)) > 200
