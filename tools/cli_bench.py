#!/usr/bin/env python
"""File-level throughput of `fade-b200 annotate` (C++ driver, fade_b200/csrc/host): simulated reads ->
SAM / BAM files -> annotate -> SAM / uBAM / BAM, records per second of wall clock per combination, and
BAM -> BAM over 1, 2, 4, 8 GPUs when the box has them (--gpus 1,2,4,8).
    python tools/cli_bench.py [--reads N] [--ref-len L] [--gpus 1,2]"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fade_b200 import sim  # noqa: E402

BIN = os.path.join(ROOT, "fade_b200", "bin", "fade-b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--ref-len", type=int, default=20_000_000)
    ap.add_argument("--dir", default="/tmp/fade_cli_bench")
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--combos", default="all", help="all = every input / output container; bam = BAM -> BAM only")
    a = ap.parse_args()
    os.makedirs(a.dir, exist_ok=True)
    cfg = sim.default_cfg(read_seed=2002)
    ref = sim.make_contig(1002, 0, a.ref_len, 0, 0, 0.0)
    rd = sim.make_reads(cfg, 0, a.reads, [ref])
    fa, sam, bam = (os.path.join(a.dir, x) for x in ("ref.fa", "in.sam", "in.bam"))
    t = time.time()
    sim.write_fasta(fa, ["chrS"], [ref])
    sim.write_bam(bam, ["chrS"], [ref], rd)
    with open(sam, "wb") as f:
        subprocess.run([BIN, "view", bam], stdout=f, check=True)
    print(f"inputs: {a.reads} reads, SAM {os.path.getsize(sam) >> 20} MiB, BAM {os.path.getsize(bam) >> 20} MiB ({time.time() - t:.0f} s to write)",
          flush=True)
    env = dict(os.environ, FADE_TIMING="1")
    combos = [(bam, ["-b"]), (bam, ["-b"]), (bam, ["-b", "--level", "6"]), (bam, ["-b", "--level", "1"]), (bam, ["-u"]), (bam, []), (sam, []), (sam, ["-b"])]
    if a.combos == "bam":
        combos = combos[:2]
    for g in [int(x) for x in a.gpus.split(",")]:
        for src, flags in combos:
            t = time.time()
            with open(os.path.join(a.dir, "out.bin"), "wb") as f:
                p = subprocess.run([BIN, "annotate", "--gpus", str(g), *flags, src, fa], stdout=f, stderr=subprocess.PIPE, text=True, env=env)
            dt = time.time() - t
            assert p.returncode == 0, p.stderr
            print(f"{g} GPU(s) {os.path.basename(src)} -> {flags or ['sam']}: {dt:.2f} s wall, {a.reads / dt / 1e6:.2f} M records/s (incl. FASTA load + upload)")
            for ln in p.stderr.splitlines():
                if "threads" in ln or "records," in ln:
                    print("   ", ln)


if __name__ == "__main__":
    main()
