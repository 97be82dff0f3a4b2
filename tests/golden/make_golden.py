"""Regenerates tests/golden/*.json.  The reference (blachlylab/fade) has no golden vectors and
cannot be run here, so these fixtures come from the ORACLE (oracle/fade_oracle.c) over the
deterministic simulator: they pin oracle + simulator against drift and give the GPU tests a
committed expectation that does not depend on rebuilding the oracle.  PARITY UNPINNED vs real fade.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from fade_b200 import sim  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    names, contigs, cfg, _ = sim.config_c1()
    n = 1500
    rd = sim.make_reads(cfg, 0, n, contigs)
    res, ops = orc.align_batch(rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left,
                               rd.clip_right, [c.tobytes() for c in contigs])
    al = np.where(res["aligned"] == 1)[0]
    out = {"config": "C1 head: ref seed 1001 (1 Mbp, N run at 400000, 1% lower-case), read seed 2001, W=300, m=5",
           "n_reads": n, "aligned_reads": al.tolist(), "alignments": []}
    for k in al:
        e = [int(res[f][k]) for f in ("score", "beg_query", "end_query", "beg_ref", "end_ref", "art_left", "art_right")]
        e.append(orc.cigar_string(ops[k, : res["n_ops"][k]]))
        e.append(int(res["win_start"][k]))
        out["alignments"].append(e)
    # record-level tags for the artifact reads among the first 1500
    L = rd.read_len
    stride = (L + 1) // 2
    tags = {}
    for k in range(n):
        t = orc.annotate_record(is_mapped=not (rd.flag[k] & 4), has_sa=bool(rd.has_sa[k]),
                                cigar=rd.cigar[k, : rd.n_cigar[k]], seq4=rd.seq4[k * stride:(k + 1) * stride],
                                qual=rd.qual[k * L:(k + 1) * L], l_qseq=L, pos=int(rd.pos[k]), contig_name=names[0],
                                ref_seq=contigs[0].tobytes())
        if t["rs"] != 0:
            tags[str(k)] = t
    out["tags"] = tags
    with open(os.path.join(HERE, "c1_head.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote c1_head.json:", len(al), "alignments,", sum("am" in t for t in tags.values()), "artifact records")


if __name__ == "__main__":
    main()
