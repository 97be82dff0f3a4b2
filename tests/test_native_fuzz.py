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
