"""The scalar oracle (oracle/fade_oracle.c: clean Gotoh recurrence + rules P3/P4) against the STRUCTURAL
restatement of parasail's striped kernel (oracle/parasail_striped.c: striped layout, query profile, lazy-F
loop with its trace-table rewrites, column-max end-cell logic, striped cigar walk) at 8 lanes (SSE builds)
and 16 lanes (AVX2 builds), plus 32 lanes (the 8-bit AVX2 kernel a saturating dispatcher would try first).  Replaces the "lane-width independence" ARGUMENT of SURVEY.md 8a by a check:
>= 10^6 generated pairs (random, planted, gapped, low-complexity = tie- and zero-heavy, wildcard letters)
must agree in score, end cell, begin cell, op count and every CIGAR op; on a divergence the assertion
names the first pair and the oracle switch (U1-U7) that would explain it.
Reference call sites: /root/reference/source/analysis.d:67,69 (p.sw_striped, res.cigar)."""
import pytest

from oracle import oracle as orc

KATS = [  # (query, target, score, end_query, end_ref, position, cigar)
    ("AAAAAAAAAACATTAGCCAT", "GGGGGCATTAGCCATGGGGG", 20, 19, 14, 5, "10S10="),
    ("AAAAAAAAAACATTAGCCAT", "GGGGGCATTTGCCATGGGGG", 15, 19, 14, 5, "10S4=1X5="),
    ("ACGTACGT", "ACGTACGTCCACGTACGT", 16, 7, 7, 0, "8="),
    ("ACGCCACG", "TTACGTT", 6, 2, 4, 2, "3=5S"),
    ("ACGNNACG", "ACGNNACG", 16, 7, 7, 0, "8="),
]


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_kats_through_the_striped_restatement(lanes):
    for q, t, score, eq, er, pos, cigar in KATS:
        r = orc.sw_trace_striped(q, t, lanes)
        assert (r.score, r.end_query, r.end_ref, r.beg_ref, r.cigar) == (score, eq, er, pos, cigar), (q, t, lanes)


def _report(r, lanes):
    if r["n_diverged"] == 0:
        return ""
    a = orc.sw_trace(r["first_q"], r["first_t"])
    b = orc.sw_trace_striped(r["first_q"], r["first_t"], lanes)
    return (f"{r['n_diverged']} of {r['n_pairs']} pairs diverge at {lanes} lanes; first: pair {r['first_div']}\n"
            f"  q = {r['first_q']}\n  t = {r['first_t']}\n"
            f"  scalar : score {a.score} end ({a.end_query},{a.end_ref}) beg ({a.beg_query},{a.beg_ref}) {a.cigar}\n"
            f"  striped: score {b.score} end ({b.end_query},{b.end_ref}) beg ({b.beg_query},{b.beg_ref}) {b.cigar}\n"
            f"  oracle switch that makes them agree: {r['explained_by']}")


@pytest.mark.parametrize("lanes,seed", [(8, 11), (16, 12), (32, 13)])
def test_one_million_pairs_scalar_equals_striped(lanes, seed):
    """3 x 340,000 pairs with fade's scoring (10, 2, +2, -3); the census shows what the sample exercised.
    8 lanes = the 16-bit kernel of the 128-bit builds, 16 lanes = 16-bit AVX2 (and the 8-bit kernel of the 128-bit
    builds, should dparasail's sw_striped be the saturating dispatcher that tries 8 bits first), 32 lanes = 8-bit AVX2."""
    r = orc.fuzz_striped(seed, 340_000, lanes)
    assert r["n_diverged"] == 0, _report(r, lanes)
    # gapped CIGARs, several cells holding the maximum (P3 tie-breaks), cells with H == 0 and E or F == 0 (U8)
    assert r["n_gapped"] > 8_000 and r["n_multi_max"] > 40_000 and r["n_zero_ef"] > 100_000, r


@pytest.mark.parametrize("scoring", [(3, 1, 2, -3), (4, 2, 2, -3), (2, 2, 1, -1), (6, 1, 5, -4)])
def test_cheap_gaps_agree_as_well(scoring):
    """gap-heavy regimes (where lazy-F has the most to repair); not fade's parameters, a stress of the structure"""
    o, e, m, x = scoring
    for lanes in (8, 16, 32):
        r = orc.fuzz_striped(21, 20_000, lanes, params=orc.default_params(gap_open=o, gap_extend=e, match=m, mismatch=x))
        assert r["n_diverged"] == 0, _report(r, lanes)
        assert r["n_gapped"] > 2_000


@pytest.mark.parametrize("switch,name", [(orc.FO_SW_END_LAST_COL, "U4"), (orc.FO_SW_E_BEFORE_F, "U5"),
                                         (orc.FO_SW_GAP_TIE_OPEN, "U5"), (orc.FO_SW_EQ_BY_MATRIX, "U7")])
def test_the_fuzz_has_teeth(switch, name):
    """flipping any tie-break rule of the scalar oracle is caught, and attributed to that switch"""
    scoring = dict(gap_open=4, gap_extend=2) if switch in (orc.FO_SW_E_BEFORE_F, orc.FO_SW_GAP_TIE_OPEN) else {}
    r = orc.fuzz_striped(31, 60_000, 16, params=orc.default_params(switches=switch, **scoring))
    assert r["n_diverged"] > 0 and r["explained_by"].startswith(name), r


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_read_sized_pairs_agree(lanes):
    """the shapes fade actually aligns (2x150 / 2x250 reads against windows of up to a thousand bases): segLen
    10..33 instead of the 1..9 of the short pairs above, scores well into the hundreds, long runs of one state"""
    r = orc.fuzz_striped(41 + lanes, 8_000, lanes, qmax=260, tmax=1000)
    assert r["n_diverged"] == 0, _report(r, lanes)
    assert r["n_gapped"] > 1_000 and r["n_multi_max"] > 1_000, r
