"""`fade-b200 annotate` (C++ host driver over the C ABI, mirror of source/anno.d:16-52) on SAM text:
every record's rs / am / as / ar / ab equal the oracle's annotateTask, the other fields and tags are
untouched, and the @PG line of anno.d:25-32 is appended."""
import os
import subprocess

import pytest

import samio
from fade_b200 import sim
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "fade_b200", "bin", "fade-b200")


@pytest.mark.parametrize("extra,min_length,window", [([], 5, 300), (["--min-length", "12", "-w", "100", "--batch", "700"], 12, 100)])
def test_cli_annotate_matches_oracle(tmp_path, extra, min_length, window):
    names, contigs, cfg, _ = sim.config_c1()
    contigs = [contigs[0][:300_000]]
    rd = sim.make_reads(cfg, 0, 3000, contigs)
    fa, sam, out = tmp_path / "ref.fa", tmp_path / "in.sam", tmp_path / "out.sam"
    samio.write_fasta(fa, names, contigs)
    samio.write_sam(sam, names, contigs, rd)
    with open(out, "w") as fo:
        p = subprocess.run([BIN, "annotate", *extra, str(sam), str(fa)], stdout=fo, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    header, recs = samio.read_sam_tags(out)
    assert header[-1].startswith("@PG\tID:fade-annotate\tPN:fade\tVN:") and "\tPP:simulator\t" in header[-1]
    assert "CL:" in header[-1] and len(recs) == rd.n
    prm = orc.default_params(min_length=min_length, window_size=window)
    L = rd.read_len
    stride = (L + 1) // 2
    refb = contigs[0].tobytes()
    n_art = 0
    for k in range(rd.n):
        fields, tags = recs[f"r{k}"]
        exp = orc.annotate_record(is_mapped=not (rd.flag[k] & 4), has_sa=bool(rd.has_sa[k]),
                                  cigar=rd.cigar[k, : rd.n_cigar[k]], seq4=rd.seq4[k * stride:(k + 1) * stride],
                                  qual=rd.qual[k * L:(k + 1) * L], l_qseq=L, pos=int(rd.pos[k]), contig_name=names[0],
                                  ref_seq=refb, params=prm)
        got = {t: tags[t] for t in ("rs", "am", "as", "ar", "ab") if t in tags}
        assert got == exp, (k, got, exp)
        assert tags["NM"] == 0 and fields[3] == str(int(rd.pos[k]) + 1)
        n_art += "am" in exp
    assert n_art > 50


def test_cli_reannotation_replaces_old_tags(tmp_path):
    names, contigs, cfg, _ = sim.config_c1()
    contigs = [contigs[0][:200_000]]
    rd = sim.make_reads(cfg, 0, 800, contigs)
    fa, sam, o1, o2 = tmp_path / "ref.fa", tmp_path / "in.sam", tmp_path / "o1.sam", tmp_path / "o2.sam"
    samio.write_fasta(fa, names, contigs)
    samio.write_sam(sam, names, contigs, rd)
    for src, dst in ((sam, o1), (o1, o2)):
        with open(dst, "w") as fo:
            assert subprocess.run([BIN, "annotate", str(src), str(fa)], stdout=fo, stderr=subprocess.DEVNULL).returncode == 0
    _, a = samio.read_sam_tags(o1)
    h2, b = samio.read_sam_tags(o2)
    assert {k: v[1] for k, v in a.items()} == {k: v[1] for k, v in b.items()}
    assert h2[-1].split("\t")[1:4:2] == ["ID:fade-annotate", "VN:fade-b200-0.1"] and "PP:fade-annotate" in h2[-1]


def test_end_to_end_chain_annotate_out_extract(tmp_path):
    """BASELINE configs[4] in miniature: annotate (GPU) -> out -c / out / extract give the same
    records as the oracle's annotation pushed through the Python restatement of the consumers."""
    from oracle import consumers as cons
    names, contigs, cfg, _ = sim.config_c1()
    contigs = [contigs[0][:300_000]]
    rd = sim.make_reads(cfg, 0, 2000, contigs)
    fa, sam, anno = tmp_path / "ref.fa", tmp_path / "in.sam", tmp_path / "anno.sam"
    samio.write_fasta(fa, names, contigs)
    samio.write_sam(sam, names, contigs, rd)
    with open(anno, "w") as fo:
        assert subprocess.run([BIN, "annotate", str(sam), str(fa)], stdout=fo, stderr=subprocess.DEVNULL).returncode == 0
    recs = [cons.parse_sam_line(ln) for ln in open(anno) if not ln.startswith("@")]
    # the annotated records carry exactly the oracle's tags (checked above); now the consumers
    for args, exp in ((["out", "-c"], cons.fade_out(recs, True, names)[0]), (["out"], cons.fade_out(recs, False, names)[0]),
                      (["extract"], cons.fade_extract(recs, names))):
        p = subprocess.run([BIN, *args, str(anno)], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        body = [ln for ln in p.stdout.splitlines() if not ln.startswith("@")]
        assert body == [cons.format_sam_line(r) for r in exp], args
        assert len(body) > 50


def test_cli_annotate_reads_and_writes_bam(tmp_path):
    """the same annotation whether the records arrive as SAM text or BAM and leave as SAM,
    uncompressed BAM (-u) or BAM (-b), as util.d:65-76; checked with the independent codec tests/bamcodec.py.
    The rs tag travels as a uint8 (`rs:C`, anno.d:94 assigns a ubyte)."""
    import bamcodec
    names, contigs, cfg, _ = sim.config_c1()
    contigs = [contigs[0][:300_000]]
    rd = sim.make_reads(cfg, 0, 2500, contigs)
    fa, sam, bam = tmp_path / "ref.fa", tmp_path / "in.sam", tmp_path / "in.bam"
    samio.write_fasta(fa, names, contigs)
    samio.write_sam(sam, names, contigs, rd)
    bam.write_bytes(bamcodec.encode(open(sam).read().splitlines()))

    def run(src, *flags):
        p = subprocess.run([BIN, "annotate", *flags, str(src), str(fa)], capture_output=True)
        assert p.returncode == 0, p.stderr.decode()
        return p.stdout

    body = lambda ls: [ln for ln in ls if not ln.startswith("@PG\tID:fade-annotate")]
    ref = body(run(sam).decode().splitlines())                           # SAM text converted on the way in
    assert sum("\tam:Z:" in ln for ln in ref) > 50
    assert body(run(sam, "--text-path").decode().splitlines()) == ref    # the line-by-line loop
    assert body(run(bam).decode().splitlines()) == ref                   # binary records end to end (bamfast.hpp)
    assert body(run(bam, "--text-path").decode().splitlines()) == ref    # BAM through the SAM text loop
    assert body(run(bam, "--batch", "700", "-t", "3").decode().splitlines()) == ref   # several double-buffered batches
    out_b = run(bam, "-b")
    assert body(bamcodec.decode(out_b)) == ref
    assert body(bamcodec.decode(run(sam, "-u"))) == ref
    assert b"rsC" in bamcodec.bgzf_decode(out_b)
    # re-annotating the annotated BAM replaces the five tags, nothing else (PP chain aside)
    again = tmp_path / "again.bam"
    again.write_bytes(out_b)
    assert body(bamcodec.decode(run(again, "-b"))) == ref
    # and the chain stays in BAM: annotate -b | out -c -b | view
    p1 = subprocess.run([BIN, "out", "-c", "-b", "-"], input=out_b, capture_output=True)
    p2 = subprocess.run([BIN, "out", "-c", "-"], input=run(sam), capture_output=True)
    assert p1.returncode == 0 and p2.returncode == 0
    strip = lambda ls: [ln for ln in ls if not ln.startswith("@PG")]
    assert strip(bamcodec.decode(p1.stdout)) == strip(p2.stdout.decode().splitlines())


def test_binary_bam_path_equals_text_path_on_rich_records(tmp_path):
    """bamfast.hpp (binary records end to end) against the SAM text loop on records of every shape: unmapped,
    '*' SEQ / QUAL / CIGAR, all aux types incl. B arrays, pre-existing rs / am / ab tags, several contigs."""
    import random

    import bamcodec
    from test_bam_io import rich_lines
    rng = random.Random(21)
    fa = tmp_path / "ref.fa"
    with open(fa, "w") as f:
        for name, ln in (("chrA", 100000), ("chrB", 5000)):
            f.write(f">{name}\n" + "".join(rng.choice("ACGT") for _ in range(ln)) + "\n")
    lines = rich_lines(3000, seed=22)
    bam = tmp_path / "rich.bam"
    bam.write_bytes(bamcodec.encode(lines))
    outs = {}
    for tag, flags in (("fast", []), ("text", ["--text-path"]), ("fast_b", ["-b", "--batch", "512"]), ("text_u", ["-u", "--text-path"])):
        p = subprocess.run([BIN, "annotate", *flags, str(bam), str(fa)], capture_output=True)
        assert p.returncode == 0, p.stderr.decode()
        ls = p.stdout.decode().splitlines() if not flags or flags == ["--text-path"] else bamcodec.decode(p.stdout)
        outs[tag] = [ln for ln in ls if not ln.startswith("@PG\tID:fade-annotate")]
    assert outs["fast"] == outs["text"] == outs["fast_b"] == outs["text_u"]
    recs = [ln for ln in outs["fast"] if not ln.startswith("@")]
    assert len(recs) == 3000 and all("\trs:i:" in ln for ln in recs)
    assert sum("\trs:i:1" in ln or "\trs:i:33" in ln for ln in recs) > 300          # soft-clipped reads went to the GPU
    assert not any(ln.count("\tam:Z:") > 1 or ln.count("\trs:i:") > 1 for ln in recs)   # old tags were replaced


def test_cli_multi_gpu_equals_single_gpu(tmp_path):
    """`fade-b200 annotate --gpus N` (one ctx per GPU, the packed reference copied GPU to GPU, batches dealt round-robin,
    one writer emitting in input order -- the reference's merge is its mutex-guarded writer, anno.d:47-49): the same
    records as a single GPU, in the same order."""
    from fade_b200 import api
    n_dev = api.device_count()
    if n_dev < 2:
        pytest.skip("needs at least two GPUs")
    names, contigs, cfg, _ = sim.config_c1()
    contigs = [contigs[0][:300_000]]
    rd = sim.make_reads(cfg, 0, 6000, contigs)
    fa, bam = tmp_path / "ref.fa", tmp_path / "in.bam"
    sim.write_fasta(str(fa), names, contigs)
    sim.write_bam(str(bam), names, contigs, rd)

    def run(*flags):
        p = subprocess.run([BIN, "annotate", *flags, str(bam), str(fa)], capture_output=True)
        assert p.returncode == 0, p.stderr.decode()
        return [ln for ln in p.stdout.decode().splitlines() if not ln.startswith("@PG\tID:fade-annotate")]

    one = run("--batch", "500")
    assert sum("\tam:Z:" in ln for ln in one) > 200
    for g in sorted({2, min(n_dev, 4), n_dev}):
        assert run("--batch", "500", "--gpus", str(g)) == one, g          # 12 batches over g GPUs
    assert run("--gpus=2") == one                                           # one batch only: the second GPU stays idle
    assert run("--batch", "700", "--gpus", "2", "--device", str(n_dev - 2)) == one
    p = subprocess.run([BIN, "annotate", "--gpus", str(n_dev + 1), str(bam), str(fa)], capture_output=True)
    assert p.returncode == 1 and b"does not have" in p.stderr
