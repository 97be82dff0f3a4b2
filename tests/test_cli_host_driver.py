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
