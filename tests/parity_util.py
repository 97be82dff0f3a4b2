"""Shared helpers for the GPU parity tests: run a read set through libfadegpu (C ABI) and through
the oracle and compare every per-read output bit for bit."""
from __future__ import annotations

import numpy as np

from fade_b200.api import MAX_OPS
from oracle import oracle as orc

FIELDS = ("score", "beg_query", "end_query", "beg_ref", "end_ref", "n_ops")


def oracle_params(ctx_params):
    return orc.default_params(gap_open=ctx_params.gap_open, gap_extend=ctx_params.gap_extend,
                              match=ctx_params.match, mismatch=ctx_params.mismatch,
                              window_size=ctx_params.window_size, min_length=ctx_params.min_length)


def run_gpu(ctx, rd, batch=None):
    own = batch is None
    if own:
        batch = ctx.alloc_batch(max(rd.n, 1), max(int(rd.seq_off[rd.n]), 16))
    batch.fill(rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left, rd.clip_right)
    batch.run()
    return batch


def compare(batch, rd, contigs, params, max_report=5):
    """Returns the number of aligned reads; raises AssertionError on the first mismatches."""
    n = rd.n
    res, ops = orc.align_batch(rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left,
                               rd.clip_right, [c.tobytes() if hasattr(c, "tobytes") else c for c in contigs],
                               params=params, ops_cap=MAX_OPS)
    flags = batch.flags[:n]
    errs = []
    g_al = (flags & 1).astype(np.int32)
    if not np.array_equal(g_al, res["aligned"]):
        bad = np.where(g_al != res["aligned"])[0]
        errs.append(f"aligned flag differs at reads {bad[:max_report].tolist()} ({len(bad)} total)")
    al = res["aligned"] == 1
    # the oracle reports coordinates for score-0 alignments as zeros as well
    for f, of in (("score", "score"), ("beg_query", "beg_query"), ("end_query", "end_query"),
                  ("beg_ref", "beg_ref"), ("end_ref", "end_ref"), ("n_ops", "n_ops")):
        g = getattr(batch, f)[:n]
        bad = np.where(al & (g != res[of]))[0]
        if len(bad):
            errs.append(f"{f} differs at reads {bad[:max_report].tolist()} ({len(bad)} total): "
                        f"gpu {g[bad[:max_report]].tolist()} oracle {res[of][bad[:max_report]].tolist()}")
    bad = np.where(al & (batch.win_start[:n] != res["win_start"]))[0]
    if len(bad):
        errs.append(f"win_start differs at {bad[:max_report].tolist()}")
    gl, gr = ((flags >> 1) & 1).astype(np.int32), ((flags >> 2) & 1).astype(np.int32)
    for name, g, o in (("art_left", gl, res["art_left"]), ("art_right", gr, res["art_right"])):
        bad = np.where(g != o)[0]
        if len(bad):
            errs.append(f"{name} differs at reads {bad[:max_report].tolist()} ({len(bad)} total)")
    k = np.minimum(res["n_ops"], MAX_OPS)
    mask = np.arange(MAX_OPS)[None, :] < k[:, None]
    gops = np.where(mask, batch.ops[:n], 0)
    oops = np.where(mask, ops, 0)
    bad = np.where(al & (gops != oops).any(axis=1))[0]
    if len(bad):
        b = int(bad[0])
        errs.append(f"CIGAR differs at reads {bad[:max_report].tolist()} ({len(bad)} total): gpu "
                    f"{orc.cigar_string(gops[b][:k[b]])} oracle {orc.cigar_string(oops[b][:k[b]])}")
    assert not errs, "\n".join(errs)
    return int(al.sum())


def compare_compact(batch, n, res, ops, max_report=5):
    """Same comparison on the compact results (fadegpu_get_results, F_NO_SCATTER contexts): every
    per-read output of `batch` against oracle results `res` / `ops` for reads [0, n).  Vectorised,
    so it can run over millions of reads.  Returns the number of aligned reads."""
    rec, ws, ridx = batch.results()
    flags = batch.flags[:n]
    errs = []
    al = res["aligned"] == 1
    g_al = (flags & 1) == 1
    if not np.array_equal(g_al, al):
        bad = np.where(g_al != al)[0]
        errs.append(f"aligned flag differs at reads {bad[:max_report].tolist()} ({len(bad)} total)")
        assert not errs, "\n".join(errs)
    idx = np.where(al)[0]
    k = ridx[idx]
    assert (k >= 0).all() and len(rec) == len(idx) and len(np.unique(k)) == len(k), "result index is not a bijection"
    r = rec[k]
    o = res[idx]
    assert np.array_equal(r["read"], idx.astype(np.int32)), "result record points at another read"
    for f in FIELDS:
        bad = np.where(r[f] != o[f])[0]
        if len(bad):
            errs.append(f"{f} differs at reads {idx[bad[:max_report]].tolist()} ({len(bad)} total): "
                        f"gpu {r[f][bad[:max_report]].tolist()} oracle {o[f][bad[:max_report]].tolist()}")
    bad = np.where(ws[k] != o["win_start"])[0]
    if len(bad):
        errs.append(f"win_start differs at {idx[bad[:max_report]].tolist()}")
    gl, gr = ((flags >> 1) & 1).astype(np.int32), ((flags >> 2) & 1).astype(np.int32)
    for name, g, oo in (("art_left", gl, res["art_left"]), ("art_right", gr, res["art_right"])):
        bad = np.where(g != oo)[0]
        if len(bad):
            errs.append(f"{name} differs at reads {bad[:max_report].tolist()} ({len(bad)} total)")
    if not np.array_equal((r["flags"] >> 1) & 3, (flags[idx] >> 1) & 3):
        errs.append("per-read flag byte and result record disagree")
    kk = np.minimum(o["n_ops"], MAX_OPS)
    mask = np.arange(MAX_OPS)[None, :] < kk[:, None]
    gops = np.where(mask, r["ops"], 0)
    oops = np.where(mask, ops[idx][:, :MAX_OPS], 0)
    bad = np.where((gops != oops).any(axis=1))[0]
    if len(bad):
        b = int(bad[0])
        errs.append(f"CIGAR differs at reads {idx[bad[:max_report]].tolist()} ({len(bad)} total): gpu "
                    f"{orc.cigar_string(gops[b][:kk[b]])} oracle {orc.cigar_string(oops[b][:kk[b]])}")
    assert not errs, "\n".join(errs)
    return int(al.sum())
