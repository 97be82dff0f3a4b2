#!/usr/bin/env python
"""BASELINE.json configs[4] at scale: annotate -> name sort -> out -c / out / extract on 10 M simulated reads through
the repo's own driver (fade-b200, fade_b200/csrc/host), with the downstream outputs verified:

  * for ALL reads: the statistics `fade out` prints (/root/reference/source/stats.d:56-72) and the record counts of
    every output against what the oracle's rs / artifact flags of all reads imply (vectorised, oracle/fade_oracle_simd.c);
  * record for record on the first --check reads (they are the first records of the name-sorted file, and `out -c`,
    `out` and `extract` keep the order): the driver's outputs against oracle/consumers.py (the Python restatement of
    filter.d:15-91,167-269 and remap.d:11-87) applied to ORACLE-annotated records (oracle/fade_oracle.c annotateTask).

    python tools/c5_chain.py [--reads 10000000] [--check 100000] [--gpus 1] [--json profiles/r02_c5_chain.json]"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from fade_b200 import sim  # noqa: E402
from oracle import consumers as cons  # noqa: E402
from oracle import oracle as orc  # noqa: E402

BIN = os.path.join(ROOT, "fade_b200", "bin", "fade-b200")


def run(args, out_path, env=None):
    t0 = time.perf_counter()
    with open(out_path, "wb") as f:
        p = subprocess.run([BIN, *args], stdout=f, stderr=subprocess.PIPE, text=True, env=env)
    if p.returncode != 0:
        raise SystemExit(f"{' '.join(args)} failed: {p.stderr[-500:]}")
    return time.perf_counter() - t0, p.stderr


def count(path):
    return int(subprocess.run([BIN, "view", "--count", path], capture_output=True, text=True, check=True).stdout)


def head_records(path, limit_index, max_lines=None):
    """SAM lines of the records whose qname r<k> has k < limit_index, read from the front of the file"""
    p = subprocess.Popen([BIN, "view", path], stdout=subprocess.PIPE, text=True)
    out = []
    for ln in p.stdout:
        if ln.startswith("@"):
            continue
        if int(ln[1:ln.index("\t")]) >= limit_index or (max_lines and len(out) >= max_lines):
            break
        out.append(ln.rstrip("\n"))
    p.kill()
    p.wait()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--ref-len", type=int, default=100_000_000)
    ap.add_argument("--check", type=int, default=100_000)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--dir", default="/tmp/fade_c5")
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    os.makedirs(a.dir, exist_ok=True)
    sim.set_threads(a.threads)
    T = {}
    t0 = time.perf_counter()
    ref = sim.make_contig(1002, 0, a.ref_len, 0, 0, 0.0)
    cfg = sim.default_cfg(read_seed=2002)
    rd = sim.make_reads(cfg, 0, a.reads, [ref])
    names = ["chrS"]
    fa, bam = os.path.join(a.dir, "ref.fa"), os.path.join(a.dir, "in.bam")
    sim.write_fasta(fa, names, [ref])
    sim.write_bam(bam, names, [ref], rd)
    T["generate_and_write_inputs_s"] = time.perf_counter() - t0
    f = lambda x: os.path.join(a.dir, x)  # noqa: E731
    env = dict(os.environ, FADE_TIMING="1")

    T["annotate_s"], err = run(["annotate", "-b", "-t", str(a.threads), "--gpus", str(a.gpus), bam, fa], f("anno.bam"), env)
    timing = [ln for ln in err.splitlines() if "record loop" in ln or "records," in ln]
    T["sort_n_s"], _ = run(["sort", "-n", "-b", f("anno.bam")], f("sorted.bam"))
    T["out_clip_s"], st_clip = run(["out", "-c", "-b", f("sorted.bam")], f("clipped.bam"))
    T["out_s"], st_out = run(["out", "-b", f("sorted.bam")], f("filtered.bam"))
    T["extract_s"], _ = run(["extract", "-b", f("sorted.bam")], f("extracted.bam"))
    sizes = {k: os.path.getsize(f(k)) for k in ("in.bam", "anno.bam", "sorted.bam", "clipped.bam", "filtered.bam", "extracted.bam")}
    t0 = time.perf_counter()
    counts = {k: count(f(k)) for k in ("anno.bam", "sorted.bam", "clipped.bam", "filtered.bam", "extracted.bam")}

    # ---- all reads: what the oracle's flags imply ----
    res, _ = orc.align_batch(rd.seq4, rd.seq_off, rd.l_qseq, rd.tid, rd.pos, rd.aligned_len, rd.clip_left, rd.clip_right, [ref],
                             ops_cap=10, simd=True, n_threads=a.threads)
    mapped = (rd.flag & 4) == 0
    has_s = (rd.clip_left > 0) | (rd.clip_right > 0)            # the simulator's S ops are leading / trailing clips
    live = mapped & has_s                                        # anno.d:61-65
    rs = np.where(live, 1 | (res["art_left"] << 1) | (res["art_right"] << 2) | (rd.has_sa.astype(np.int32) << 5), 0)
    st = cons.Stats()
    st.read_count = int(rd.n)
    st.clipped = int((rs & 1).sum()); st.art = int(((rs & 6) != 0).sum()); st.sup = int(((rs >> 5) & 1).sum())
    st.art_sup = int((((rs & 6) != 0) & (((rs >> 5) & 1) == 1)).sum())
    st.aln_l = int(((rs >> 1) & 1).sum()); st.aln_r = int(((rs >> 2) & 1).sum())
    exp_counts = {"anno.bam": rd.n, "sorted.bam": rd.n, "clipped.bam": rd.n, "filtered.bam": rd.n - st.art,
                  "extracted.bam": st.aln_l + st.aln_r}
    checks = {"counts_equal": counts == exp_counts, "counts": counts, "expected_counts": exp_counts,
              "stats_out_clip_equal": st_clip.strip().splitlines()[-7:] == st.lines(),
              "stats_out_equal": st_out.strip().splitlines()[-7:] == st.lines(), "stats": st.lines()}

    # ---- the first --check reads, record for record, against consumers.py on oracle-annotated records ----
    n_chk = min(a.check, rd.n)
    L = rd.read_len
    stride = (L + 1) // 2
    refb = ref.tobytes()
    in_lines = head_records(bam, n_chk)
    assert len(in_lines) == n_chk
    recs = []
    for k, ln in enumerate(in_lines):
        t = orc.annotate_record(is_mapped=not (rd.flag[k] & 4), has_sa=bool(rd.has_sa[k]), cigar=rd.cigar[k, : rd.n_cigar[k]],
                                seq4=rd.seq4[k * stride:(k + 1) * stride], qual=rd.qual[k * L:(k + 1) * L], l_qseq=L,
                                pos=int(rd.pos[k]), contig_name=names[0], ref_seq=refb)
        assert t["rs"] == int(rs[k]), (k, t["rs"], int(rs[k]))
        r = cons.parse_sam_line(ln)
        r["tags"]["rs"] = ("i", str(t["rs"]))
        for tag in ("am", "as", "ar", "ab"):
            if tag in t:
                r["tags"][tag] = ("Z", t[tag])
        recs.append(r)
    fmt = lambda rs_: [cons.format_sam_line(r) for r in rs_]  # noqa: E731
    exp_clip = fmt(cons.fade_out(recs, True, names)[0])
    exp_out = fmt(cons.fade_out(recs, False, names)[0])
    exp_ext = fmt(cons.fade_extract(recs, names))
    got_anno = head_records(f("sorted.bam"), n_chk)
    got_clip = head_records(f("clipped.bam"), 1 << 62, max_lines=n_chk)     # blank records lose their name order: by position
    got_out = head_records(f("filtered.bam"), n_chk)
    got_ext = head_records(f("extracted.bam"), n_chk)
    checks.update(prefix_reads=n_chk, annotated_prefix_equal=got_anno == fmt(recs), out_clip_prefix_equal=got_clip == exp_clip,
                  out_prefix_equal=got_out == exp_out, extract_prefix_equal=got_ext == exp_ext,
                  prefix_artifact_records=sum("am" in r["tags"] for r in recs), prefix_extract_records=len(exp_ext))
    T["verify_s"] = time.perf_counter() - t0
    ok = all(v for k, v in checks.items() if k.endswith("_equal"))
    summary = {"config": f"{rd.n} simulated 2x150 reads (seed 2002) vs {a.ref_len} bp reference (seed 1002); annotate -b -> sort -n -b -> "
                         f"out -c -b / out -b / extract -b; {a.threads} host threads, {a.gpus} GPU(s)",
               "ok": ok, "seconds": {k: round(v, 2) for k, v in T.items()},
               "records_per_s": {k[:-2]: round(rd.n / v) for k, v in T.items() if k in ("annotate_s", "sort_n_s", "out_clip_s", "out_s", "extract_s")},
               "annotate_timing": timing, "bytes": sizes, "checks": checks}
    print(json.dumps(summary, indent=1))
    if a.json:
        with open(a.json, "w") as fo:
            json.dump(summary, fo, indent=1)
    shutil.rmtree(a.dir, ignore_errors=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
