#!/usr/bin/env python
"""File-level throughput of `fade-b200 annotate` (C++ driver, fade_b200/csrc/host): simulated reads ->
SAM / BAM files -> annotate -> SAM / uBAM / BAM, records per second of wall clock per combination.
    python tools/cli_bench.py [--reads N] [--ref-len L]"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import samio  # noqa: E402
from fade_b200 import sim  # noqa: E402

BIN = os.path.join(ROOT, "fade_b200", "bin", "fade-b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--ref-len", type=int, default=20_000_000)
    ap.add_argument("--dir", default="/tmp/fade_cli_bench")
    a = ap.parse_args()
    os.makedirs(a.dir, exist_ok=True)
    cfg = sim.default_cfg(read_seed=2002)
    ref = sim.make_contig(1002, 0, a.ref_len, 0, 0, 0.0)
    rd = sim.make_reads(cfg, 0, a.reads, [ref])
    fa, sam, bam = (os.path.join(a.dir, x) for x in ("ref.fa", "in.sam", "in.bam"))
    t = time.time()
    samio.write_fasta(fa, ["chrS"], [ref])
    samio.write_sam(sam, ["chrS"], [ref], rd)
    with open(bam, "wb") as f:
        subprocess.run([BIN, "view", "-b", sam], stdout=f, check=True)
    print(f"inputs: {a.reads} reads, SAM {os.path.getsize(sam) >> 20} MiB, BAM {os.path.getsize(bam) >> 20} MiB ({time.time() - t:.0f} s to write)",
          flush=True)
    env = dict(os.environ, FADE_TIMING="1")
    for src, flags in ((bam, ["-b"]), (bam, ["-b"]), (bam, ["-u"]), (bam, []), (sam, []), (sam, ["-b"])):
        t = time.time()
        with open(os.path.join(a.dir, "out.bin"), "wb") as f:
            p = subprocess.run([BIN, "annotate", *flags, src, fa], stdout=f, stderr=subprocess.PIPE, text=True, env=env)
        dt = time.time() - t
        assert p.returncode == 0, p.stderr
        print(f"{os.path.basename(src)} -> {flags or ['sam']}: {dt:.2f} s wall, {a.reads / dt / 1e6:.2f} M records/s (incl. FASTA load + upload)")
        for ln in p.stderr.splitlines():
            if "threads" in ln or "records," in ln:
                print("   ", ln)


if __name__ == "__main__":
    main()
