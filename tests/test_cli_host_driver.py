"""The HOST side of `fade-b200 annotate` (record parsing, compact batch layout, ring of batches over several GPUs,
tag assembly, SAM / BAM writers; mirror of source/anno.d:16-52,55-110) on the CPU: the driver runs against
tests/native/standin_device.cpp, a stand-in for the GPU side of the C ABI that is LD_PRELOADed by this test alone and
answers with the oracle's results.  The same assertions run against the real library on a B200 in tests/test_gpu_cli.py;
here they cover what the driver does around the device.  (The product has no CPU path: without the preload the driver
fails with FADEGPU_E_NODEV, which test_abi_and_host.py checks.)"""
import os
import subprocess

import pytest

import samio
from fade_b200 import sim
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "fade_b200", "bin", "fade-b200")


@pytest.fixture(scope="module")
def standin(tmp_path_factory):
    orc.lib()   # builds oracle/libfadeoracle.so when needed
    d = tmp_path_factory.mktemp("standin")
    so = d / "libstandin_device.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "native", "standin_device.cpp"),
                           "-o", str(so), "-L", os.path.join(ROOT, "oracle"), "-lfadeoracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return str(so)


@pytest.fixture
def preload(standin, monkeypatch):
    monkeypatch.setenv("LD_PRELOAD", standin)
    monkeypatch.delenv("FADE_STANDIN_DEVICES", raising=False)
    return monkeypatch


def test_without_the_standin_the_driver_refuses_to_run(tmp_path):
    names, contigs, cfg, _ = sim.config_c1()
    contigs = [contigs[0][:50_000]]
    rd = sim.make_reads(cfg, 0, 50, contigs)
    samio.write_fasta(tmp_path / "ref.fa", names, contigs)
    samio.write_sam(tmp_path / "in.sam", names, contigs, rd)
    env = {k: v for k, v in os.environ.items() if k != "LD_PRELOAD"}
    env["CUDA_VISIBLE_DEVICES"] = ""
    p = subprocess.run([BIN, "annotate", str(tmp_path / "in.sam"), str(tmp_path / "ref.fa")], capture_output=True, env=env)
    assert p.returncode != 0 and b"\tr0\t" not in p.stdout and p.stderr


@pytest.mark.parametrize("extra,min_length,window", [([], 5, 300), (["--min-length", "12", "-w", "100", "--batch", "700"], 12, 100),
                                                     (["--text-path"], 5, 300)])
def test_records_carry_the_oracles_tags(preload, tmp_path, extra, min_length, window):
    import test_gpu_cli
    test_gpu_cli.test_cli_annotate_matches_oracle(tmp_path, extra, min_length, window)


def test_reannotation_and_the_consumer_chain(preload, tmp_path):
    import test_gpu_cli
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    test_gpu_cli.test_cli_reannotation_replaces_old_tags(tmp_path / "a")
    test_gpu_cli.test_end_to_end_chain_annotate_out_extract(tmp_path / "b")


def test_bam_routes_of_the_driver(preload, tmp_path):
    """BAM in / out, --text-path, re-annotation of an annotated BAM, `annotate -b | out -c -b`, and records of every
    shape (all aux types, '*' fields, several contigs) through the binary and the text loop"""
    import test_gpu_cli
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    test_gpu_cli.test_cli_annotate_reads_and_writes_bam(tmp_path / "a")
    test_gpu_cli.test_binary_bam_path_equals_text_path_on_rich_records(tmp_path / "b")


def test_ring_over_several_devices_keeps_input_order(preload, tmp_path):
    """anno.d:44-50's one writer: with 1, 2 and 3 devices and batches far smaller than the file (so that the ring wraps
    many times and ends on a partial round) the output is byte for byte the same, for SAM, uBAM and BAM output and for
    SAM and BAM input."""
    names, contigs, cfg, _ = sim.config_c1()
    contigs = [contigs[0][:300_000]]
    rd = sim.make_reads(cfg, 0, 5000, contigs)
    fa, sam, bam = tmp_path / "ref.fa", tmp_path / "in.sam", tmp_path / "in.bam"
    samio.write_fasta(fa, names, contigs)
    samio.write_sam(sam, names, contigs, rd)
    with open(bam, "wb") as fo:
        assert subprocess.run([BIN, "view", "-b", str(sam)], stdout=fo).returncode == 0

    def run(src, con, gpus, batch):
        preload.setenv("FADE_STANDIN_DEVICES", str(gpus))
        p = subprocess.run([BIN, "annotate", *con, "--gpus", str(gpus), "--batch", str(batch), str(src), str(fa)], capture_output=True)
        assert p.returncode == 0, p.stderr.decode()
        return p.stdout

    def records(raw, con):   # the @PG line records the command line, which differs between the runs
        if con:
            with open(tmp_path / "x.bam", "wb") as fo:
                fo.write(raw)
            raw = subprocess.run([BIN, "view", str(tmp_path / "x.bam")], capture_output=True, check=True).stdout
        return [ln for ln in raw.split(b"\n") if not ln.startswith(b"@PG")]

    for con in ([], ["-u"], ["-b"]):
        base = records(run(sam, con, 1, 100_000), con)
        assert sum(b"\tam:Z:" in ln for ln in base) > 100
        for src, gpus, batch in ((sam, 2, 300), (bam, 3, 257), (bam, 1, 64), (sam, 3, 5000)):
            assert records(run(src, con, gpus, batch), con) == base, (con, src, gpus, batch)
    # more devices asked for than there are
    preload.setenv("FADE_STANDIN_DEVICES", "2")
    p = subprocess.run([BIN, "annotate", "--gpus", "3", str(sam), str(fa)], capture_output=True)
    assert p.returncode == 1 and b"devices this machine does not have" in p.stderr


def test_awkward_records_get_the_oracles_tags(preload, tmp_path):
    """Records the simulator never writes: hard clips, clips on both sides, insertions / deletions / N skips in the read's
    own CIGAR, unmapped and secondary / supplementary records, SA tags, reads at both contig ends (window clamp,
    analysis.d:45-59), clips at and below the length floor, lower-case reference, planted fold-back artifacts on either
    side.  Every record's rs / am / as / ar / ab must be what annotateTask (anno.d:55-110) gives in the oracle; every
    other field and tag must pass through untouched; SAM and BAM routes must agree."""
    import random
    import numpy as np
    rng = random.Random(99)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    rc = lambda s: "".join(comp[c] for c in reversed(s))
    contigs = {"c1": "".join(rng.choice("ACGT") for _ in range(6000)), "c2": "".join(rng.choice("ACGT") for _ in range(900))}
    contigs["c1"] = contigs["c1"][:3000] + contigs["c1"][3000:3400].lower() + contigs["c1"][3400:]   # soft-masked stretch
    names = list(contigs)
    rand = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    recs = []

    def add(name, flag, contig, pos0, cigar, seq, extra=()):
        qual = "".join(chr(33 + rng.randint(2, 40)) for _ in seq)
        recs.append([name, str(flag), contig, str(pos0 + 1), "60", cigar, "*", "0", "0", seq, qual, "NM:i:0", *extra])

    def ref(contig, a, n):
        return contigs[contig][a:a + n].upper()

    k = 0
    for contig, L in (("c1", 6000), ("c2", 900)):
        for _ in range(60):
            cl, cr = rng.choice([0, 0, 3, 5, 6, 12, 25, 40]), rng.choice([0, 0, 4, 5, 7, 15, 30])
            m = rng.randint(30, 90)
            pos0 = rng.choice([0, 1, 5, L - m - 1, L - m, rng.randint(0, L - m)])
            pos0 = max(0, min(pos0, L - m))
            body = ref(contig, pos0, m)
            # planted artifact: the clip is the reverse complement of reference text inside the window
            def clip_seq(n):
                if n == 0:
                    return ""
                if rng.random() < 0.6:
                    a = rng.randint(max(0, pos0 - 250), max(0, min(L - n, pos0 + m + 250 - n)))
                    s = rc(ref(contig, a, n))
                    if rng.random() < 0.3 and n > 8:
                        s = s[:n // 2] + comp[s[n // 2]] + s[n // 2 + 1:]
                    return s
                return rand(n)
            seq = clip_seq(cl) + body + clip_seq(cr)
            cigar = (f"{cl}S" if cl else "") + f"{m}M" + (f"{cr}S" if cr else "")
            style = rng.randrange(8)
            extra = []
            flag = rng.choice([0, 16, 99, 147])
            if style == 0:
                cigar = f"{rng.randint(1, 9)}H" + cigar + f"{rng.randint(1, 9)}H"
            elif style == 1 and m > 40:      # insertion + deletion inside the aligned part
                cigar = (f"{cl}S" if cl else "") + f"20M3I{m - 23 - 10}M4D10M" + (f"{cr}S" if cr else "")
            elif style == 2 and m > 40 and pos0 + m + 700 < L:   # spliced
                cigar = (f"{cl}S" if cl else "") + f"20M700N{m - 20}M" + (f"{cr}S" if cr else "")
            elif style == 3:
                extra.append("SA:Z:c2,10,+,30M70S,60,0;")
            elif style == 4:
                flag = rng.choice([256, 2048, 2064])
            elif style == 5:
                flag, cigar = 4, "*"
            add(f"q{k}", flag, contig if style != 5 else "*", pos0 if style != 5 else -1, cigar, seq, extra)
            k += 1
    add("qN", 0, "c1", 100, "10S40M", rc(ref("c1", 200, 10))[:5] + "NNNNN" + ref("c1", 100, 40))
    add("qx", 0, "c1", 100, "40M", ref("c1", 100, 40), ["rs:i:7", "am:Z:stale", "XX:Z:keep"])      # stale tags are replaced
    head = ["@HD\tVN:1.6\tSO:unsorted"] + [f"@SQ\tSN:{n}\tLN:{len(s)}" for n, s in contigs.items()]
    sam, fa, out = tmp_path / "in.sam", tmp_path / "ref.fa", tmp_path / "out.sam"
    sam.write_text("\n".join(head + ["\t".join(r) for r in recs]) + "\n")
    fa.write_text("".join(f">{n}\n" + "\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + "\n" for n, s in contigs.items()))
    with open(out, "w") as fo:
        p = subprocess.run([BIN, "annotate", "--batch", "37", str(sam), str(fa)], stdout=fo, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    _, got = samio.read_sam_tags(out)
    assert len(got) == len(recs)
    n_tags = n_sc = 0
    for r in recs:
        fields, tags = got[r[0]]
        assert fields == r[:11], r[0]
        flag = int(r[1])
        cigar = orc.cigar_from_string(r[5]) if r[5] != "*" else np.zeros(0, np.uint32)
        exp = orc.annotate_record(is_mapped=not (flag & 4), has_sa=any(x.startswith("SA:") for x in r[11:]), cigar=cigar,
                                  seq4=orc.pack_nt16(r[9]), qual=np.frombuffer(r[10].encode(), np.uint8) - 33, l_qseq=len(r[9]),
                                  pos=int(r[3]) - 1, contig_name=r[2], ref_seq=contigs.get(r[2], "").encode())
        assert {t: tags[t] for t in ("rs", "am", "as", "ar", "ab") if t in tags} == exp, (r[0], r[5], tags, exp)
        keep = [x for x in r[11:] if x[:2] not in ("rs", "am", "as", "ar", "ab")]
        assert all(tags[x[:2]] == (int(x[5:]) if x[3] == "i" else x[5:]) for x in keep), r[0]
        n_tags += "am" in exp
        n_sc += exp["rs"] & 1
    assert n_tags > 15 and n_sc > 60, (n_tags, n_sc)
    assert got["qx"][1]["XX"] == "keep" and "am" not in got["qx"][1]
    # BAM in, BAM out: the same records
    with open(tmp_path / "in.bam", "wb") as fo:
        assert subprocess.run([BIN, "view", "-b", str(sam)], stdout=fo).returncode == 0
    with open(tmp_path / "out.bam", "wb") as fo:
        assert subprocess.run([BIN, "annotate", "-b", "--batch", "64", str(tmp_path / "in.bam"), str(fa)], stdout=fo).returncode == 0
    back = subprocess.run([BIN, "view", str(tmp_path / "out.bam")], capture_output=True, text=True, check=True).stdout
    strip = lambda text: [ln for ln in text.splitlines() if not ln.startswith("@PG")]
    assert strip(back) == strip(out.read_text())
