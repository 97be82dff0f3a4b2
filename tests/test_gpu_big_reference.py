"""BASELINE configs[2] geometry: a 3.1 Gbp, 24-contig reference resident in HBM (global base offsets
beyond 2^31).  Heavy (several GB of host RAM), so it only runs when FADE_BIG=1; it was run by hand
on the B200 box."""
import os

import numpy as np
import pytest

from fade_b200 import Context, api, default_params, sim
from parity_util import compare, oracle_params

pytestmark = pytest.mark.gpu

# hg38 chr1-22,X,Y lengths scaled to sum 3.1 Gbp
HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
        135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167,
        46709983, 50818468, 156040895, 57227415]


@pytest.mark.skipif(os.environ.get("FADE_BIG") != "1", reason="set FADE_BIG=1 (needs ~10 GB host RAM)")
def test_hg38_sized_reference_and_late_contigs():
    scale = 3.1e9 / sum(HG38)
    lens = [int(x * scale) for x in HG38]
    names = [f"chr{i + 1}" for i in range(22)] + ["chrX", "chrY"]
    contigs = [sim.make_contig(1003, i, n, 1_000_000, 10_000, 0.0) for i, n in enumerate(lens)]
    cfg = sim.default_cfg(read_seed=2003)
    rd = sim.make_reads(cfg, 0, 400_000, contigs, with_records=False)
    assert len(np.unique(rd.tid)) == 24
    with Context(0, default_params(flags=api.F_NO_SCATTER)) as ctx:
        ctx.load_reference(names, [c.tobytes() for c in contigs])
        n_contigs, total, dev_bytes = ctx.reference_info()
        assert n_contigs == 24 and total == sum(lens) and dev_bytes < 0.52 * total
        b = ctx.alloc_batch(rd.n, int(rd.seq_off[rd.n]))
        b.submit_arrays(rd.n, rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left, rd.clip_right)
        b.wait()
        rec, ws, ridx = b.results()
        fl = b.flags[: rd.n]
        tl_, tr_ = (rd.truth & 1) == 1, (rd.truth & 2) == 2
        assert ((fl[tl_] & 2) != 0).mean() > 0.8 and ((fl[tr_] & 4) != 0).mean() > 0.8
        # recall must hold on the LAST contigs too (global offsets > 2^31)
        late = rd.tid >= 20
        assert late.sum() > 1000 and ((fl[tl_ & late] & 2) != 0).mean() > 0.8
        b.close()
    # bit-exact on a subsample restricted to the last four contigs (keeps the oracle's copy small)
    from types import SimpleNamespace
    idx = np.where(rd.tid >= 20)[0][:6000]
    stride = (cfg.read_len + 1) // 2
    sub = SimpleNamespace(n=len(idx), seq4=np.concatenate([rd.seq4[k * stride:(k + 1) * stride] for k in idx]),
                          seq_off=np.arange(len(idx) + 1, dtype=np.int64) * stride, l_qseq=rd.l_qseq[idx].copy(),
                          tid=rd.tid[idx].copy(), pos=rd.pos[idx].copy(), aligned_len=rd.aligned_len[idx].copy(),
                          clip_left=rd.clip_left[idx].copy(), clip_right=rd.clip_right[idx].copy())
    with Context(0) as ctx:
        ctx.load_reference(names, [c.tobytes() for c in contigs])
        b = ctx.alloc_batch(sub.n, int(sub.seq_off[sub.n]))
        b.fill(sub.seq4, sub.seq_off, sub.l_qseq, sub.tid, sub.pos, sub.aligned_len, sub.clip_left, sub.clip_right).run()
        small = [c if i >= 20 else c[:1] for i, c in enumerate(contigs)]   # oracle only needs the late contigs
        compare(b, sub, small, oracle_params(ctx.params))
        b.close()
