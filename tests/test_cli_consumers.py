"""`fade-b200 out [-c]` and `fade-b200 extract` (C++, SAM text) against the
independent Python restatement oracle/consumers.py.  Pure host code: runs without a GPU.  PARITY
UNPINNED with respect to real fade (filter.d / remap.d cannot be executed here)."""
import os
import random
import subprocess

import pytest

import samio
from fade_b200 import sim
from oracle import consumers as cons
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "fade_b200", "bin", "fade-b200")


def annotated_sam(tmp_path, n=2500, name_sorted=True, with_untagged=True):
    """An annotated SAM as `fade annotate` would write it, produced with the ORACLE (no GPU)."""
    names, contigs, cfg, _ = sim.config_c1()
    contigs = [contigs[0][:300_000]]
    rd = sim.make_reads(cfg, 0, n, contigs)
    path = tmp_path / "anno.sam"
    L = rd.read_len
    stride = (L + 1) // 2
    refb = contigs[0].tobytes()
    lines = ["@HD\tVN:1.6\tSO:queryname", f"@SQ\tSN:{names[0]}\tLN:{len(contigs[0])}", "@SQ\tSN:chrOther\tLN:1000",
             "@PG\tID:fade-annotate\tPN:fade"]
    order = list(range(n))
    if not name_sorted:
        random.Random(3).shuffle(order)
    for k in order:
        seq = orc.decode_nt16(rd.seq4[k * stride:(k + 1) * stride], L)
        qual = "".join(chr(int(q) + 33) for q in rd.qual[k * L:(k + 1) * L])
        cig = orc.cigar_string(rd.cigar[k, : rd.n_cigar[k]]) or "*"
        t = orc.annotate_record(is_mapped=not (rd.flag[k] & 4), has_sa=bool(rd.has_sa[k]), cigar=rd.cigar[k, : rd.n_cigar[k]],
                                seq4=rd.seq4[k * stride:(k + 1) * stride], qual=rd.qual[k * L:(k + 1) * L], l_qseq=L,
                                pos=int(rd.pos[k]), contig_name=names[0], ref_seq=refb)
        f = [f"r{k // 2}", str(int(rd.flag[k])), names[0], str(int(rd.pos[k]) + 1), "60", cig, "=", "1", "0", seq, qual, "NM:i:1"]
        if not (with_untagged and k % 97 == 0):
            f.append(f"rs:i:{t['rs']}")
            for tag in ("am", "as", "ar", "ab"):
                if tag in t:
                    f.append(f"{tag}:Z:{t[tag]}")
        lines.append("\t".join(f))
    path.write_text("\n".join(lines) + "\n")
    return path, [names[0], "chrOther"]


def run_cli(args, path):
    p = subprocess.run([BIN, *args, str(path)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    body = [ln for ln in p.stdout.splitlines() if not ln.startswith("@")]
    head = [ln for ln in p.stdout.splitlines() if ln.startswith("@")]
    return head, body, p.stderr


def load(path):
    recs = [cons.parse_sam_line(ln) for ln in open(path) if not ln.startswith("@")]
    return recs


@pytest.mark.parametrize("name_sorted", [True, False])
def test_out_filter_mode(tmp_path, name_sorted):
    path, contigs = annotated_sam(tmp_path, name_sorted=name_sorted)
    head, body, err = run_cli(["out"], path)
    exp, st = cons.fade_out(load(path), clip=False, contigs=contigs)
    assert body == [cons.format_sam_line(r) for r in exp]
    assert 0 < len(body) < 2500
    assert head[-1].startswith("@PG\tID:fade-extract\tPN:fade\tVN:") and "PP:fade-annotate" in head[-1]   # filter.d:173 (sic)
    assert ("looks name-sorted" in err) == name_sorted
    assert err.strip().splitlines()[-7:] == st.lines()


def test_out_clip_mode(tmp_path):
    path, contigs = annotated_sam(tmp_path)
    head, body, err = run_cli(["out", "-c"], path)
    recs = load(path)
    exp, st = cons.fade_out(recs, clip=True, contigs=contigs)
    assert body == [cons.format_sam_line(r) for r in exp] and len(body) == len(recs)
    changed = [(a, b) for a, b in zip(recs, exp) if cons.format_sam_line(a) != cons.format_sam_line(b)]
    assert len(changed) > 50
    for a, b in changed[:200]:
        if b["cigar"] == "*":
            continue                                    # whole alignment was artifact: blank record (U9)
        ops = cons.cigar_ops(b["cigar"])
        assert "H" in (ops[0][1], ops[-1][1])
        assert sum(n for n, op in ops if op in cons.QUERY_OPS) == len(b["seq"]) == len(b["qual"])
        assert sum(n for n, op in ops if op in "MIS=XH") == len(a["seq"])
    assert err.strip().splitlines()[-7:] == st.lines()


def test_extract(tmp_path):
    path, contigs = annotated_sam(tmp_path)
    head, body, err = run_cli(["extract"], path)
    exp = cons.fade_extract(load(path), contigs)
    assert body == [cons.format_sam_line(r) for r in exp] and len(body) > 50
    for ln in body[:100]:
        r = cons.parse_sam_line(ln)
        ops = cons.cigar_ops(r["cigar"])
        assert sum(n for n, op in ops if op in cons.QUERY_OPS) == len(r["seq"])      # am CIGAR covers the RC'd read
        assert r["flag"] in (0, 16) and r["rname"] == contigs[0]


def test_natural_compare_and_clip_units():
    c = cons.natural_compare
    assert c("r2", "r10") < 0 and c("r10", "r2") > 0 and c("a", "a") == 0 and c("r1a", "r1b") < 0 and c("r1", "r1x") < 0
    r = cons.parse_sam_line("q\t0\tchr1\t101\t60\t30S120M\t*\t0\t0\t" + "A" * 150 + "\t" + "I" * 150 +
                            "\trs:i:3\tam:Z:chr1,50,126S24=;")
    out = cons.clip_read(r, 3, ["chr1"])
    assert (out["cigar"], out["pos"], len(out["seq"])) == ("54H96M", 125, 96)
    r = cons.parse_sam_line("q\t16\tchr1\t101\t60\t110M40S\t*\t0\t0\t" + "A" * 150 + "\t" + "I" * 150 +
                            "\trs:i:5\tam:Z:;chr1,500,30=120S")
    out = cons.clip_read(r, 5, ["chr1"])
    assert (out["cigar"], out["pos"], len(out["seq"])) == ("80M70H", 101, 80)
    r["tags"]["am"] = ("Z", ";chr1,500,120=30S")          # artifact spans the whole alignment -> blank record
    out = cons.clip_read(r, 5, ["chr1"])
    assert out["cigar"] == "*" and out["flag"] == 0 and not out["tags"]


def test_consumers_stream_short_empty_and_malformed_inputs(tmp_path):
    """The consumers pull records one at a time (the first ten decide the name-sorted question, filter.d:215-217):
    inputs shorter than that, header-only inputs and a malformed record in the middle of the stream."""
    path, contigs = annotated_sam(tmp_path, n=300)
    lines = path.read_text().splitlines()
    head = [ln for ln in lines if ln.startswith("@")]
    recs = [ln for ln in lines if not ln.startswith("@")]
    for k in (0, 1, 3, 9, 10, 11):
        p = tmp_path / f"short{k}.sam"
        p.write_text("\n".join(head + recs[:k]) + "\n")
        for args, clip in ((["out"], False), (["out", "-c"], True)):
            _, body, err = run_cli(args, p)
            exp, st = cons.fade_out(load(p), clip=clip, contigs=contigs)
            assert body == [cons.format_sam_line(r) for r in exp], (k, args)
            # (0 reads: the rates are 0/0; the sign a NaN prints with is not pinned by anything)
            assert [ln.replace("-nan", "nan") for ln in err.strip().splitlines()[-7:]] == st.lines()
        _, body, _ = run_cli(["extract"], p)
        assert body == [cons.format_sam_line(r) for r in cons.fade_extract(load(p), contigs)]
    bad = tmp_path / "bad.sam"
    bad.write_text("\n".join(head + recs[:40] + ["only\tthree\tfields"] + recs[40:60]) + "\n")
    for args in (["out"], ["out", "-c"], ["extract"]):
        p = subprocess.run([BIN, *args, str(bad)], capture_output=True, text=True)
        assert p.returncode == 1 and "malformed" in p.stderr


def test_sort_by_name_feeds_out(tmp_path):
    """`fade-b200 sort -n` (the `samtools sort -n` step of BASELINE configs[4]): a shuffled annotated file comes
    back in the natural name order `fade out` tests for, mates adjacent (first of pair first), and `out` then
    ejects whole read groups exactly as on the originally sorted file."""
    path, contigs = annotated_sam(tmp_path, n=2000, name_sorted=False)
    (tmp_path / "s").mkdir()
    sorted_path, _ = annotated_sam(tmp_path / "s", n=2000, name_sorted=True)
    p = subprocess.run([BIN, "sort", "-n", str(path)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    out_lines = p.stdout.splitlines()
    head = [ln for ln in out_lines if ln.startswith("@")]
    body = [ln for ln in out_lines if not ln.startswith("@")]
    assert head[0].startswith("@HD") and "SO:queryname" in head[0]
    assert sorted(body) == sorted(ln for ln in open(path).read().splitlines() if not ln.startswith("@"))
    names = [ln.split("\t")[0] for ln in body]
    assert all(cons.natural_compare(a, b) <= 0 for a, b in zip(names, names[1:]))
    flags = [int(ln.split("\t")[1]) & 0xc0 for ln in body]
    assert all(fa <= fb for (na, fa), (nb, fb) in zip(zip(names, flags), zip(names[1:], flags[1:])) if na == nb)
    srt = tmp_path / "sorted.sam"
    srt.write_text(p.stdout)
    _, got, err = run_cli(["out"], srt)
    _, exp, _ = run_cli(["out"], sorted_path)
    assert "looks name-sorted" in err and sorted(got) == sorted(exp) and 0 < len(got) < 2000
    # BAM in / BAM out, and a header without @HD
    import bamcodec
    nohd = [ln for ln in open(path).read().splitlines() if not ln.startswith("@HD")]
    p2 = subprocess.run([BIN, "sort", "-n", "-b", "-"], input=bamcodec.encode(nohd), capture_output=True)
    assert p2.returncode == 0
    dec = bamcodec.decode(p2.stdout)
    assert dec[0] == "@HD\tVN:1.6\tSO:queryname" and [ln for ln in dec if not ln.startswith("@")] == body


def test_sort_by_name_order_is_the_reference_comparison_on_awkward_names(tmp_path):
    """The sort works on binary records with its own in-place restatement of filter.d:127-165: names mixing
    letters, digit runs, leading zeros and shared prefixes must come out in the order of the Python restatement,
    names that compare equal ("a01" / "a1") keep their input order, first of pair precedes second, and the SAM and
    BAM routes agree."""
    import functools
    import random
    import bamcodec
    rng = random.Random(5)
    alphabet = ["a", "b", ":", "_", "0", "1", "9", "10", "007"]
    names = ["".join(rng.choice(alphabet) for _ in range(rng.randint(1, 7))) for _ in range(1500)]
    names += ["a1", "a01", "a001", "x", "x", "9", "09"]
    head = ["@HD\tVN:1.6\tSO:unsorted", "@SQ\tSN:c\tLN:1000"]
    recs = []
    for k, nm in enumerate(names):
        flag = rng.choice([0x41, 0x81, 0])
        recs.append(f"{nm}\t{flag}\tc\t{1 + k % 900}\t60\t4M\t*\t0\t0\tACGT\tIIII\tXI:i:{k}")
    path = tmp_path / "in.sam"
    path.write_text("\n".join(head + recs) + "\n")
    p = subprocess.run([BIN, "sort", "-n", str(path)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    body = [ln for ln in p.stdout.splitlines() if not ln.startswith("@")]

    def key_cmp(x, y):
        fx, fy = x.split("\t"), y.split("\t")
        c = cons.natural_compare(fx[0], fy[0])
        return c if c else ((int(fx[1]) & 0xc0) > (int(fy[1]) & 0xc0)) - ((int(fx[1]) & 0xc0) < (int(fy[1]) & 0xc0))
    assert body == sorted(recs, key=functools.cmp_to_key(key_cmp))     # sorted() is stable, like the tool
    p2 = subprocess.run([BIN, "sort", "-nb", "-"], input=bamcodec.encode(head + recs), capture_output=True)
    assert p2.returncode == 0 and [ln for ln in bamcodec.decode(p2.stdout) if not ln.startswith("@")] == body


def test_sort_refuses_damaged_input(tmp_path):
    import bamcodec
    head = ["@HD\tVN:1.6", "@SQ\tSN:c\tLN:1000"]
    recs = [f"r{k}\t0\tc\t{1 + k}\t60\t4M\t*\t0\t0\tACGT\tIIII" for k in range(50)]
    bad = tmp_path / "bad.sam"
    bad.write_text("\n".join(head + recs[:20] + ["only\tthree\tfields"] + recs[20:]) + "\n")
    p = subprocess.run([BIN, "sort", "-n", str(bad)], capture_output=True)
    assert p.returncode == 1 and b"malformed" in p.stderr and b"r0" not in p.stdout
    raw = bamcodec.encode(head + recs)
    cut = subprocess.run([BIN, "sort", "-n", "-"], input=raw[: len(raw) - 60], capture_output=True)
    assert cut.returncode == 1 and b"r0" not in cut.stdout
    # intact BGZF, but one record declares more bases than it holds
    import struct
    payload = bytearray(bamcodec.bgzf_decode(raw))
    at = payload.index(b"r7\0") - 36                       # block_size of record r7
    struct.pack_into("<i", payload, at + 4 + 16, 5000)      # l_seq
    for args in (["sort", "-n"], ["sort", "-nb"]):
        p = subprocess.run([BIN, *args, "-"], input=bamcodec.bgzf_blocks(bytes(payload)) + raw[-28:], capture_output=True)
        assert p.returncode == 1 and b"malformed" in p.stderr


def test_out_and_extract_on_bam_input_equal_those_on_sam_input(tmp_path, monkeypatch):
    """BAM input takes a binary route through `out` (records copied as bytes, only clipped ones through text); SAM input
    the text route.  Both must give the same records, statistics and warnings: name-sorted and unsorted input, -c, all
    three containers, the file read in pieces of a single BGZF block (read groups then straddle the pieces), a giant
    read group, fewer than ten records, none at all, and rs tags of unusual types."""
    import bamcodec

    def both(lines, args):
        sam = ("\n".join(lines) + "\n").encode()
        bam = bamcodec.encode(lines)
        outs = []
        for src in (sam, bam):
            p = subprocess.run([BIN, *args, "-"], input=src, capture_output=True)
            assert p.returncode == 0, p.stderr
            raw = p.stdout
            if raw[:2] == b"\x1f\x8b":
                raw = "\n".join(bamcodec.decode(raw)).encode()
            outs.append(([ln for ln in raw.decode().splitlines() if not ln.startswith("@PG")], p.stderr))
        assert outs[0] == outs[1], args
        return outs[0][0]

    monkeypatch.setenv("FADE_IO_BLOCKS", "1")
    for name_sorted in (True, False):
        d = tmp_path / str(name_sorted)
        d.mkdir()
        path, _ = annotated_sam(d, n=3000, name_sorted=name_sorted)
        lines = open(path).read().splitlines()
        n_in = sum(not ln.startswith("@") for ln in lines)
        for args in (["out"], ["out", "-c"], ["out", "-b"], ["out", "-cb"], ["out", "-u"]):
            got = both(lines, args)
            n_out = sum(not ln.startswith("@") for ln in got)
            assert (n_out == n_in) if "-c" in args or "-cb" in args else (0 < n_out < n_in)
    head = [ln for ln in lines if ln.startswith("@")]
    body = [ln for ln in lines if not ln.startswith("@")]
    giant = [("same\t" + ln.split("\t", 1)[1]) for ln in body[:1500]]          # one read group larger than any piece
    both(head + giant, ["out", "-b"]); both(head + giant, ["out"])
    for few in (body[:3], []):
        both(head + few, ["out"]); both(head + few, ["out", "-cb"])
    odd = []
    for k, ln in enumerate(body[:40]):
        f = [x for x in ln.split("\t") if not x.startswith("rs:")]
        odd.append("\t".join(f + [["rs:Z:3", "rs:A:7", "rs:f:5", "rs:i:300", "rs:Z:junk"][k % 5]]))
    both(head + odd, ["out"]); both(head + odd, ["out", "-b"])
    # extract skips BAM records without artifact bits before they are converted to text: same output as from SAM input
    for args in (["extract"], ["extract", "-b"]):
        assert sum(not ln.startswith("@") for ln in both(lines, args)) > 50
        both(head + odd, args); both(head, args)
