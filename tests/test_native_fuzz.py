"""fade_b200/csrc/host/fastdeflate.hpp / fastinflate.hpp (the DEFLATE encoder and decoder of the BAM writers and readers)
fuzzed against zlib under AddressSanitizer + UBSan: tests/native/fuzz_deflate.cpp is compiled here and run."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_deflate_encoder_and_decoder_against_zlib_under_sanitizers(tmp_path):
    exe = tmp_path / "fuzz_deflate"
    src = os.path.join(ROOT, "tests", "native", "fuzz_deflate.cpp")
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                         "-o", str(exe), src, "-lz"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-2000:]
    p = subprocess.run([str(exe), "-", "1200"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "fails 0" in p.stdout and "ERROR" not in p.stderr, p.stdout[-1500:] + p.stderr[-3000:]
    # the corrupted streams were both refused and (harmlessly) decoded: the sanitizers saw every path
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("cases")][0]
    assert int(line.split("cases ")[1].split()[0]) > 4000


def test_record_level_damage_through_the_cli_under_sanitizers(tmp_path):
    """The binary-record commands (sort -n, out, extract on BAM input, view --bulk) built with AddressSanitizer + UBSan and
    fed intact BGZF whose PAYLOAD has random bytes overwritten (names, sizes, CIGARs, aux fields) and ends inside a
    record: every run must end with exit code 0 or 1 and no sanitizer report."""
    import random
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bamcodec
    import test_cli_consumers as tc
    exe = tmp_path / "fade-asan"
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fopenmp", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                         "-o", str(exe), os.path.join(ROOT, "fade_b200", "csrc", "host", "fade_cli.cpp"),
                         "-L", os.path.join(ROOT, "fade_b200"), "-lfadegpu", "-lz", "-Wl,-rpath," + os.path.join(ROOT, "fade_b200")],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-2000:]
    path, _ = tc.annotated_sam(tmp_path, n=1500, name_sorted=True)
    good = bamcodec.encode(open(path).read().splitlines())
    payload = bytearray(bamcodec.bgzf_decode(good))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0", FADE_IO_BLOCKS="2")
    rng = random.Random(7)
    outcomes = set()
    for trial in range(10):
        p = bytearray(payload[: len(payload) - (rng.randint(0, 200) if trial else 0)])     # trial 0: the intact file
        if trial:
            for _ in range(rng.randint(1, 6)):
                p[rng.randrange(300, len(p))] = rng.randrange(256)
        data = bamcodec.bgzf_blocks(bytes(p)) + good[-28:]
        for cmd in (["sort", "-n", "-b"], ["out", "-b"], ["out", "-c"], ["extract"], ["view", "--bulk"]):
            r = subprocess.run([str(exe), *cmd, "-"], input=data, capture_output=True, env=env, timeout=300)
            assert r.returncode in (0, 1) and b"Sanitizer" not in r.stderr and b"runtime error" not in r.stderr, (trial, cmd, r.stderr[-1500:])
            outcomes.add(r.returncode)
    assert outcomes == {0, 1}      # both the accepting and the refusing paths ran
