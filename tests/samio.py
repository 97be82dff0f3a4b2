"""Minimal SAM / FASTA text writers and a tag parser for the CLI tests (no htslib here)."""
from __future__ import annotations

import numpy as np

from oracle import oracle as orc


def write_fasta(path, names, contigs, width=60):
    with open(path, "w") as f:
        for name, c in zip(names, contigs):
            s = c.tobytes().decode() if hasattr(c, "tobytes") else c.decode()
            f.write(f">{name} synthetic\n")
            for a in range(0, len(s), width):
                f.write(s[a:a + width] + "\n")


def write_sam(path, names, contigs, rd, extra_header=("@PG\tID:simulator\tPN:fadesim",)):
    L = rd.read_len
    stride = (L + 1) // 2
    with open(path, "w") as f:
        f.write("@HD\tVN:1.6\tSO:unsorted\n")
        for name, c in zip(names, contigs):
            f.write(f"@SQ\tSN:{name}\tLN:{len(c)}\n")
        for h in extra_header:
            f.write(h + "\n")
        for k in range(rd.n):
            seq = orc.decode_nt16(rd.seq4[k * stride:(k + 1) * stride], L)
            qual = "".join(chr(int(q) + 33) for q in rd.qual[k * L:(k + 1) * L])
            cig = orc.cigar_string(rd.cigar[k, : rd.n_cigar[k]]) or "*"
            fields = [f"r{k}", str(int(rd.flag[k])), names[int(rd.tid[k])], str(int(rd.pos[k]) + 1), "60", cig, "*", "0",
                      "0", seq, qual, "NM:i:0"]
            if rd.has_sa[k]:
                fields.append(f"SA:Z:{names[0]},1,+,50M100S,60,0;")
            f.write("\t".join(fields) + "\n")


def read_sam_tags(path):
    """-> (header lines, {qname: (mandatory fields, tag dict)}); integer tags become int."""
    header, recs = [], {}
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith("@"):
                header.append(line)
                continue
            fl = line.split("\t")
            tags = {}
            for t in fl[11:]:
                k, ty, v = t.split(":", 2)
                tags[k] = int(v) if ty == "i" else v
            recs[fl[0]] = (fl[:11], tags)
    return header, recs
