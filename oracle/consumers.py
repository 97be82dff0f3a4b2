"""CPU restatement (plain Python) of the CONSUMERS of the rs/am tags: `fade out [-c]` and
`fade extract` -- SURVEY.md 8(f) next rows 1 and 2.  TEST INFRASTRUCTURE ONLY, PARITY UNPINNED.

Follows, line by line:
    source/filter.d:15-91    clipRead
    source/filter.d:127-165  numericallyAwareStringComparison
    source/filter.d:167-269  filter!(clip)
    source/remap.d:11-87     remapArtifacts
    source/stats.d:45-72     Stats.parse / Stats.print
Records are dicts of SAM fields: qname flag rname pos(1-based) mapq cigar rnext pnext tlen seq qual
tags (ordered dict name -> (type, value)).

Uncertain point U9 (dhtslib/htslib are absent): a "blank" record `SAMRecord(header)` is htslib's
bam_init1(), i.e. calloc: every core field 0 -> tid 0 / pos 0 / mtid 0 / mpos 0, which SAM text shows
as RNAME = first contig, POS 1, RNEXT "=" (or the first contig when tid differs), PNEXT 1.
"""
from __future__ import annotations

import re
from collections import OrderedDict

CIG_RE = re.compile(r"(\d+)([MIDNSHP=XB])")
QUERY_OPS = set("MIS=X")
REF_OPS = set("MDN=X")
COMP = {"=": "=", "A": "T", "C": "G", "M": "K", "G": "C", "R": "Y", "S": "S", "V": "B", "T": "A", "W": "W",
        "Y": "R", "H": "D", "K": "M", "D": "H", "B": "V", "N": "N"}


def parse_sam_line(line: str) -> dict:
    f = line.rstrip("\n").split("\t")
    tags = OrderedDict()
    for t in f[11:]:
        k, ty, v = t.split(":", 2)
        tags[k] = (ty, v)
    return dict(qname=f[0], flag=int(f[1]), rname=f[2], pos=int(f[3]), mapq=int(f[4]), cigar=f[5], rnext=f[6],
                pnext=int(f[7]), tlen=int(f[8]), seq=f[9], qual=f[10], tags=tags)


def format_sam_line(r: dict) -> str:
    out = [r["qname"], str(r["flag"]), r["rname"], str(r["pos"]), str(r["mapq"]), r["cigar"], r["rnext"],
           str(r["pnext"]), str(r["tlen"]), r["seq"], r["qual"]]
    out += [f"{k}:{ty}:{v}" for k, (ty, v) in r["tags"].items()]
    return "\t".join(out)


def cigar_ops(s: str):
    return [] if s == "*" else [[int(n), op] for n, op in CIG_RE.findall(s)]


def cigar_str(ops) -> str:
    return "".join(f"{n}{op}" for n, op in ops) or "*"


def aligned_length(ops) -> int:
    return sum(n for n, op in ops if op in REF_OPS)


def blank_record(qname: str, seq: str, qual: str, contigs: list[str]) -> dict:
    """U9: SAMRecord(rec.h) with only queryName / sequence / qscores set (filter.d:47-50,79-82)."""
    first = contigs[0] if contigs else "*"
    return dict(qname=qname, flag=0, rname=first, pos=1 if contigs else 0, mapq=0, cigar="*",
                rnext="=" if contigs else "*", pnext=1 if contigs else 0, tlen=0, seq=seq, qual=qual, tags=OrderedDict())


def clip_read(rec: dict, rs: int, contigs: list[str]) -> dict:
    """filter.d:15-91.  Returns the (possibly blank) record."""
    new_cigar = cigar_ops(rec["cigar"])
    pos, seq, qual = rec["pos"], rec["seq"], rec["qual"]
    am = rec["tags"]["am"][1]
    if rs & 2:   # art_left, filter.d:22-54
        art = cigar_ops(am.split(";")[0].split(",")[2])
        to_trim = aligned_length(art)
        hard = 0
        if to_trim < aligned_length(cigar_ops(rec["cigar"])):       # rec.cigar: the ORIGINAL cigar (filter.d:29)
            while to_trim:
                if new_cigar[0][1] in QUERY_OPS:
                    seq, qual = seq[1:], qual[1:]
                    hard += 1
                if new_cigar[0][1] in REF_OPS:
                    pos += 1
                    to_trim -= 1
                new_cigar[0][0] -= 1
                if not new_cigar[0][0]:
                    new_cigar = new_cigar[1:]
        else:
            return blank_record(rec["qname"], seq, qual, contigs)
        new_cigar = [[hard, "H"]] + new_cigar
    if rs & 4:   # art_right, filter.d:55-86
        art = cigar_ops(am.split(";")[1].split(",")[2])
        to_trim = aligned_length(art)
        hard = 0
        if to_trim < aligned_length(new_cigar):                      # new_cigar (filter.d:63)
            while to_trim:
                if new_cigar[-1][1] in QUERY_OPS:
                    seq, qual = seq[:-1], qual[:-1]
                    hard += 1
                if new_cigar[-1][1] in REF_OPS:
                    to_trim -= 1
                new_cigar[-1][0] -= 1
                if not new_cigar[-1][0]:
                    new_cigar = new_cigar[:-1]
        else:
            return blank_record(rec["qname"], seq, qual, contigs)
        new_cigar = new_cigar + [[hard, "H"]]
    out = dict(rec)
    out.update(cigar=cigar_str(new_cigar), seq=seq, qual=qual, pos=pos)
    return out


def natural_compare(a: str, b: str) -> int:
    """filter.d:127-165 (parse!long consumes the leading digits; on failure the value stays -1)."""
    while a and b:
        if not a[0].isdigit() and not b[0].isdigit():
            if a[0] == b[0]:
                a, b = a[1:], b[1:]
                continue
            return -1 if a[0] < b[0] else 1
        def take(s):
            m = re.match(r"\d+", s)
            if not m:
                return -1, s
            return int(m.group()), s[m.end():]
        ai, a2 = take(a)
        bi, b2 = take(b)
        if ai == bi:
            if (a2, b2) == (a, b):     # neither side consumed anything: the D loop would spin; cannot happen for mixed input
                return 0
            a, b = a2, b2
            continue
        return -1 if ai < bi else 1
    return 0 if len(a) == len(b) else (-1 if len(a) < len(b) else 1)


class Stats:
    """stats.d:16-72 (only the fields `fade out` prints)."""

    def __init__(self):
        self.read_count = self.clipped = self.sup = self.art_sup = self.art = self.aln_l = self.aln_r = 0

    def parse(self, rs: int):
        al, ar = (rs >> 1) & 1, (rs >> 2) & 1
        self.clipped += rs & 1
        self.art += al | ar
        self.sup += (rs >> 5) & 1
        self.art_sup += (al | ar) & ((rs >> 5) & 1)
        self.aln_l += al
        self.aln_r += ar

    def lines(self):
        n = float(self.read_count) if self.read_count else float("nan")
        return [f"read count:\t{self.read_count}", f"Clipped %:\t{self.clipped / n:g}",
                f"% With Supplementary alns:\t{self.sup / n:g}", f"Artifact rate:\t{self.art / n:g}",
                f"% With Supplementary alns and artifacts:\t{self.art_sup / n:g}",
                f"Artifact rate left only:\t{self.aln_l / n:g}", f"Artifact rate right only:\t{self.aln_r / n:g}"]


def pg_line(header: list[str], prog_id: str, version: str, cl: str) -> str:
    last = None
    for h in header:
        if h.startswith("@PG"):
            for x in h.split("\t"):
                if x.startswith("ID:"):
                    last = x[3:]
    s = f"@PG\tID:{prog_id}\tPN:fade\tVN:{version}"
    if last is not None:
        s += f"\tPP:{last}"
    return s + f"\tCL:{cl}"


def _rs_of(rec):
    t = rec["tags"].get("rs")
    return None if t is None else int(t[1]) & 0xff


def fade_out(records: list[dict], clip: bool, contigs: list[str]):
    """filter.d:167-269 -> (output records, Stats)."""
    st = Stats()
    out = []
    if clip:
        for rec in records:
            st.read_count += 1
            rs = _rs_of(rec)
            if rs is None:
                out.append(rec)
                continue
            st.parse(rs)
            out.append(rec if not (rs & 6) else clip_read(rec, rs, contigs))
        return out, st
    head = records[:10]
    is_sorted = all(not (natural_compare(head[k + 1]["qname"], head[k]["qname"]) < 0) for k in range(len(head) - 1))
    if is_sorted:
        k = 0
        while k < len(records):
            e = k + 1
            while e < len(records) and records[e]["qname"] == records[e - 1]["qname"]:
                e += 1
            group = records[k:e]
            art = False
            for rec in group:
                st.read_count += 1
                rs = _rs_of(rec)
                if rs is None:
                    continue
                st.parse(rs)
                art |= bool(rs & 6)
            if not art:
                out.extend(group)
            k = e
    else:
        for rec in records:
            st.read_count += 1
            rs = _rs_of(rec)
            if rs is None:
                continue
            st.parse(rs)
            if not (rs & 6):
                out.append(rec)
    return out, st


def fade_extract(records: list[dict], contigs: list[str]):
    """remap.d:29-85: one new record per artifact side."""
    out = []
    for rec in records:
        rs = _rs_of(rec)
        if rs is None or not (rs & 6) or "am" not in rec["tags"]:
            continue
        am = rec["tags"]["am"][1].split(";")
        for side, bit in ((0, 2), (1, 4)):
            if not (rs & bit):
                continue
            chrom, pos0, cig = am[side].split(",")
            tid = contigs.index(chrom) if chrom in contigs else -1
            seq = "".join(COMP.get(c, "N") for c in reversed(rec["seq"]))       # remap.d:59 / util.d:23-34
            new = dict(qname=rec["qname"], flag=0 if (rec["flag"] & 16) else 16, rname=chrom if tid >= 0 else "*",
                       pos=int(pos0) + 1, mapq=0, cigar=cig,
                       rnext=("=" if tid == 0 else (contigs[0] if contigs else "*")), pnext=1 if contigs else 0,   # U9
                       tlen=0, seq=seq, qual=rec["qual"][::-1], tags=OrderedDict())
            out.append(new)
    return out
