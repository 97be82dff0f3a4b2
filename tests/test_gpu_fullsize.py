"""Full-size checks (one 1 M-read chunk of BASELINE configs[1] / a chunk of the stress config)
through size-independent properties, plus bit-exact comparison on random subsamples."""
import numpy as np
import pytest

from fade_b200 import Context, api, default_params, sim
from oracle import oracle as orc
from parity_util import compare, oracle_params

MAX_OPS = api.MAX_OPS

pytestmark = pytest.mark.gpu

N16 = "=ACMGRSVTWYHKDBN"


def _subsample(rd, idx, read_len):
    from types import SimpleNamespace
    stride = (read_len + 1) // 2
    seq4 = np.concatenate([rd.seq4[k * stride:(k + 1) * stride] for k in idx])
    off = np.arange(len(idx) + 1, dtype=np.int64) * stride
    take = lambda a: np.ascontiguousarray(a[idx])  # noqa: E731
    return SimpleNamespace(n=len(idx), seq4=seq4, seq_off=off, l_qseq=take(rd.l_qseq), tid=take(rd.tid),
                           pos=take(rd.pos), aligned_len=take(rd.aligned_len), clip_left=take(rd.clip_left),
                           clip_right=take(rd.clip_right))


def _check_properties(b, rd, ref, params, rng, n_score_checks=3000):
    n = rd.n
    rec, ws, ridx = b.results()
    al = np.where(b.flags[:n] & 1)[0]
    assert len(rec) == len(al)
    ops = rec["ops"]
    nops = rec["n_ops"]
    assert (nops <= MAX_OPS).mean() > 0.95      # fade rejects CIGARs of more than 10 ops (analysis.d:69)
    ok = nops <= MAX_OPS
    ln = (ops >> 4).astype(np.int64)
    op = ops & 0xf
    valid = np.arange(MAX_OPS)[None, :] < nops[:, None]
    qcons = np.where(valid & np.isin(op, (1, 4, 7, 8)), ln, 0).sum(1)
    rcons = np.where(valid & np.isin(op, (2, 7, 8)), ln, 0).sum(1)
    acons = np.where(valid & np.isin(op, (1, 7, 8)), ln, 0).sum(1)
    L = rd.l_qseq[rec["read"]]
    pos = rec["score"] > 0
    m = ok & pos
    assert np.array_equal(qcons[m], L[m]), "query-consuming ops must cover the whole read (S padding, P5)"
    assert np.array_equal(rcons[m], (rec["end_ref"] - rec["beg_ref"] + 1)[m])
    assert np.array_equal(acons[m], (rec["end_query"] - rec["beg_query"] + 1)[m])
    assert (rec["score"][m] <= 2 * L[m]).all() and (rec["beg_query"][m] >= 0).all()
    tl = np.minimum(rd.pos[rec["read"]] + rd.aligned_len[rec["read"]] + params.window_size, len(ref)) - ws
    assert (rec["end_ref"][m] < tl[m]).all() and (ws >= 0).all()
    # accept flags imply the predicates of analysis.d:69-80 / :98-104
    fl = rec["flags"]
    first = op[:, 0]
    last = op[np.arange(len(rec)), np.maximum(nops - 1, 0) % MAX_OPS]
    cl, cr = rd.clip_left[rec["read"]], rd.clip_right[rec["read"]]
    left = (fl & 2) != 0
    right = (fl & 4) != 0
    assert (left <= ((nops >= 1) & (nops <= 10) & (last == 7) & (first == 4) & (5 * rec["score"] > 9 * cl) & (cl > params.min_length))).all()
    assert (right <= ((nops >= 1) & (nops <= 10) & (first == 7) & (last == 4) & (5 * rec["score"] > 9 * cr) & (cr > params.min_length))).all()
    # score recomputed from CIGAR + sequences on a sample
    stride = (rd.read_len + 1) // 2
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    for k in rng.choice(np.where(m)[0], size=min(n_score_checks, int(m.sum())), replace=False):
        r = int(rec["read"][k])
        s4 = rd.seq4[r * stride:(r + 1) * stride]
        read = "".join(N16[(s4[i >> 1] >> ((~i & 1) << 2)) & 0xf] for i in range(int(L[k])))
        q = "".join(comp[c] for c in reversed(read))
        t = ref[int(ws[k]): int(ws[k]) + int(tl[k])].tobytes().decode().upper()
        i, j, sc = 0, int(rec["beg_ref"][k]), 0
        for o in ops[k, : nops[k]]:
            ln_, op_ = int(o) >> 4, int(o) & 0xf
            if op_ == 4:
                i += ln_
            elif op_ in (7, 8):
                for _ in range(ln_):
                    assert (q[i] == t[j]) == (op_ == 7)
                    sc += 2 if q[i] == t[j] else -3
                    i += 1
                    j += 1
            elif op_ == 1:
                sc -= 10 + 2 * (ln_ - 1)
                i += ln_
            elif op_ == 2:
                sc -= 10 + 2 * (ln_ - 1)
                j += ln_
        assert sc == int(rec["score"][k]), (r, sc, int(rec["score"][k]))
    return rec, al


def test_one_million_reads_properties_and_subsample():
    rng = np.random.default_rng(5)
    ref = sim.make_contig(1002, 0, 20_000_000, 0, 0, 0.0)
    cfg = sim.default_cfg(read_seed=2002)
    rd = sim.make_reads(cfg, 0, 1_000_000, [ref], with_records=False)
    with Context(0, default_params(flags=api.F_NO_SCATTER)) as ctx:
        ctx.load_reference(["chrS"], [ref.tobytes()])
        b = ctx.alloc_batch(rd.n, int(rd.seq_off[rd.n]))
        b.submit_arrays(rd.n, rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left, rd.clip_right)
        b.wait()
        rec, al = _check_properties(b, rd, ref, ctx.params, rng)
        assert 150_000 < len(al) < 190_000
        # recall of the planted in-window artifacts
        tl_, tr_ = (rd.truth & 1) == 1, (rd.truth & 2) == 2
        fl = b.flags[: rd.n]
        assert ((fl[tl_] & 2) != 0).mean() > 0.85 and ((fl[tr_] & 4) != 0).mean() > 0.85
        # idempotence: a second pass over the same inputs gives identical records
        first = rec[np.argsort(rec["read"])].copy()
        b.submit_arrays(rd.n, rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left, rd.clip_right)
        b.wait()
        rec2, _, _ = b.results()
        assert np.array_equal(first, rec2[np.argsort(rec2["read"])])
        b.close()
    # bit-exact against the oracle on a random subsample of the reads
    idx = np.sort(rng.choice(rd.n, size=20_000, replace=False))
    sub = _subsample(rd, idx, cfg.read_len)
    with Context(0) as ctx:
        ctx.load_reference(["chrS"], [ref.tobytes()])
        b = ctx.alloc_batch(sub.n, int(sub.seq_off[sub.n]))
        b.fill(sub.seq4, sub.seq_off, sub.l_qseq, sub.tid, sub.pos, sub.aligned_len, sub.clip_left, sub.clip_right).run()
        compare(b, sub, [ref], oracle_params(ctx.params))
        b.close()


def test_stress_config_properties():
    """BASELINE configs[3] at scale: 2x250 reads, --window-size 1000, clip law U{1..40}."""
    rng = np.random.default_rng(6)
    ref = sim.make_contig(1002, 0, 20_000_000, 0, 0, 0.0)
    cfg = sim.default_cfg(read_seed=2004, read_len=250, window=1000, frag_mean=600, frag_sd=80, short_clip_law=1)
    rd = sim.make_reads(cfg, 0, 200_000, [ref], with_records=False)
    with Context(0, default_params(window_size=1000, min_length=5, flags=api.F_NO_SCATTER)) as ctx:
        ctx.load_reference(["chrS"], [ref.tobytes()])
        b = ctx.alloc_batch(rd.n, int(rd.seq_off[rd.n]))
        b.submit_arrays(rd.n, rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left, rd.clip_right)
        b.wait()
        rec, al = _check_properties(b, rd, ref, ctx.params, rng, n_score_checks=800)
        st = b.stats()
        assert len(al) > 20_000 and st.n_generic == 0
        b.close()
