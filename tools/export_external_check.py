#!/usr/bin/env python
"""Package ONE external run of the real `fade annotate` that would pin this repo's parity.

The reference cannot be built in the development container (D + dparasail/parasail + dhtslib/htslib
are absent), so every parity claim rests on oracle/ (a restatement).  This tool writes everything a
person with a working `fade` needs to check the restatement against the real thing:

    python tools/export_external_check.py --out external_check [--reads 10000]

    external_check/ref.fa              BASELINE configs[0] reference (1 Mbp, seed 1001, N run, lower-case tiles)
    external_check/reads.sam           10,000 simulated 2x150 reads (seed 2001), already aligned, with soft clips
    external_check/expected_tags.tsv   per record (qname, flag): rs, am, as, ar, ab as the ORACLE predicts them
    external_check/expected_tags.<U>.tsv   the same under each single uncertainty switch U1..U7 flipped
    external_check/compare_external.py     dependency-free comparer
    external_check/README.txt          the two-line recipe

Recipe (on a machine with fade 0.x and samtools):
    fade annotate reads.sam ref.fa | samtools sort -n -O sam - > fade_out.sam
    python compare_external.py fade_out.sam            # prints PARITY OK or the first mismatches

(`fade annotate` follows /root/reference/source/anno.d:16-110; its output order is unspecified, hence the name sort;
the comparer sorts by (qname, flag) itself, so `samtools sort` is optional.)"""
import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SWITCHES = {"U1_no_softclip_pad": 1, "U4_end_last_col": 2, "U5_e_before_f": 4, "U5_gap_tie_open": 8, "U7_eq_by_matrix": 16,
            "U3_swap_id": 32, "P1_wildcard_scores_mismatch": 64}

COMPARER = r'''#!/usr/bin/env python
"""compare_external.py FADE_OUT.sam [expected_tags.tsv] -- compares the rs/am/as/ar/ab tags of a real
`fade annotate` run with the oracle's prediction (same directory).  No dependencies."""
import glob, os, sys

def load_expected(path):
    exp = {}
    with open(path) as f:
        next(f)
        for line in f:
            q, flag, rs, am, as_, ar, ab = line.rstrip("\n").split("\t")
            exp[(q, int(flag))] = (int(rs), am, as_, ar, ab)
    return exp

def load_sam(path):
    got = {}
    with open(path) as f:
        for line in f:
            if line.startswith("@"):
                continue
            fl = line.rstrip("\n").split("\t")
            tags = {}
            for t in fl[11:]:
                k, ty, v = t.split(":", 2)
                tags[k] = v
            rs = tags.get("rs")
            got[(fl[0], int(fl[1]))] = (int(rs) if rs is not None else None, tags.get("am", ""), tags.get("as", ""),
                                         tags.get("ar", ""), tags.get("ab", ""))
    return got

def diff(got, exp):
    bad = []
    for key in sorted(exp):
        if key not in got:
            bad.append((key, "missing record", exp[key], None))
        elif got[key] != exp[key]:
            bad.append((key, "tags differ", exp[key], got[key]))
    extra = [k for k in got if k not in exp]
    return bad, extra

def main():
    here = os.path.dirname(os.path.abspath(__file__))
    sam = sys.argv[1]
    exp_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(here, "expected_tags.tsv")
    got = load_sam(sam)
    bad, extra = diff(got, load_expected(exp_path))
    n_art = sum(1 for v in got.values() if v[0] is not None and v[0] & 6)
    print(f"{len(got)} records read, {n_art} carry an artifact bit")
    if not bad and not extra:
        print("PARITY OK: every rs/am/as/ar/ab tag equals the oracle's prediction")
        return 0
    print(f"MISMATCH: {len(bad)} records differ, {len(extra)} unexpected records")
    for key, what, e, g in bad[:10]:
        print(f"  {key}: {what}\n    expected {e}\n    got      {g}")
    # does one of the documented uncertainty switches explain the run?
    for alt in sorted(glob.glob(os.path.join(here, "expected_tags.U*.tsv"))):
        b2, x2 = diff(got, load_expected(alt))
        print(f"  against {os.path.basename(alt)}: {len(b2)} records differ" + ("  <-- this switch explains the run" if not b2 and not x2 else ""))
    return 1

if __name__ == "__main__":
    sys.exit(main())
'''

README = """External parity check for fade-b200 (see tools/export_external_check.py in the repository).

  fade annotate reads.sam ref.fa | samtools sort -n -O sam - > fade_out.sam
  python compare_external.py fade_out.sam

"PARITY OK" pins the oracle (and with it every bit-exact GPU test) to real fade {version unknown here}/parasail 2.4.3.
A mismatch lists the first differing records and tells whether one of the documented uncertainty
switches (U1, U4, U5, U7 of SURVEY.md 8c / oracle/fade_oracle.h) explains the whole run.
Inputs: {n} reads, reference {ref_len} bp; expected: {n_art} artifact records, rs histogram {hist}.
"""


def expected_tags(names, contigs, rd, params):
    from oracle import oracle as orc
    L = rd.read_len
    stride = (L + 1) // 2
    refs = [c.tobytes() for c in contigs]
    rows = []
    for k in range(rd.n):
        t = orc.annotate_record(is_mapped=not (rd.flag[k] & 4), has_sa=bool(rd.has_sa[k]),
                                cigar=rd.cigar[k, : rd.n_cigar[k]], seq4=rd.seq4[k * stride:(k + 1) * stride],
                                qual=rd.qual[k * L:(k + 1) * L], l_qseq=L, pos=int(rd.pos[k]),
                                contig_name=names[int(rd.tid[k])], ref_seq=refs[int(rd.tid[k])], params=params)
        rows.append((f"r{k}", int(rd.flag[k]), t["rs"], t.get("am", ""), t.get("as", ""), t.get("ar", ""), t.get("ab", "")))
    return rows


def write_tsv(path, rows):
    with open(path, "w") as f:
        f.write("qname\tflag\trs\tam\tas\tar\tab\n")
        for r in rows:
            f.write("\t".join(str(x) for x in r) + "\n")


def export(out_dir: str, n_reads: int = 10_000, variants: bool = True) -> dict:
    import samio
    from fade_b200 import sim
    from oracle import oracle as orc
    os.makedirs(out_dir, exist_ok=True)
    names, contigs, cfg, n_default = sim.config_c1()
    n = n_reads or n_default
    rd = sim.make_reads(cfg, 0, n, contigs)
    samio.write_fasta(os.path.join(out_dir, "ref.fa"), names, contigs)
    samio.write_sam(os.path.join(out_dir, "reads.sam"), names, contigs, rd)
    rows = expected_tags(names, contigs, rd, orc.default_params())
    write_tsv(os.path.join(out_dir, "expected_tags.tsv"), rows)
    if variants:
        for name, bit in SWITCHES.items():
            write_tsv(os.path.join(out_dir, f"expected_tags.{name}.tsv"),
                      expected_tags(names, contigs, rd, orc.default_params(switches=bit)))
    with open(os.path.join(out_dir, "compare_external.py"), "w") as f:
        f.write(COMPARER)
    hist = {}
    for r in rows:
        hist[r[2]] = hist.get(r[2], 0) + 1
    n_art = sum(1 for r in rows if r[2] & 6)
    with open(os.path.join(out_dir, "README.txt"), "w") as f:
        f.write(README.replace("{n}", str(n)).replace("{ref_len}", str(sum(len(c) for c in contigs)))
                .replace("{n_art}", str(n_art)).replace("{hist}", str(dict(sorted(hist.items()))))
                .replace(" {version unknown here}", ""))
    return {"reads": n, "artifact_records": n_art, "rs_hist": hist}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="external_check")
    ap.add_argument("--reads", type=int, default=10_000)
    ap.add_argument("--no-variants", action="store_true")
    a = ap.parse_args()
    info = export(a.out, a.reads, not a.no_variants)
    print(f"wrote {a.out}/: {info}")
    print(open(os.path.join(a.out, "README.txt")).read())


if __name__ == "__main__":
    main()
