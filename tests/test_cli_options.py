"""Command-line conventions of the reference (std.getopt with config.bundling, /root/reference/source/app.d:73-107) in the
C++ driver: bundled short flags, --name=value, unknown options rejected; and the file readers refuse size fields
a well-formed BGZF / BAM file cannot contain.  No GPU needed (view / sort / out / extract are host-only)."""
import os
import struct
import subprocess
import zlib

import pytest

import samio
from fade_b200 import sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "fade_b200", "bin", "fade-b200")


@pytest.fixture(scope="module")
def sam(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    names, contigs, cfg, _ = sim.config_c1()
    contigs = [contigs[0][:100_000]]
    rd = sim.make_reads(cfg, 0, 400, contigs)
    p = d / "in.sam"
    samio.write_sam(p, names, contigs, rd)
    # give every record an rs tag so that `out` keeps them (filter.d:228-231 drops untagged records)
    lines = [ln if ln.startswith("@") else ln.rstrip("\n") + "\trs:i:0\n" for ln in open(p)]
    open(p, "w").writelines(lines)
    return str(p)


def run(*args, **kw):
    return subprocess.run([BIN, *args], capture_output=True, **kw)


def test_bundled_short_flags_and_equals_forms(sam):
    a = run("out", "-c", "-b", sam)
    b = run("out", "-cb", sam)
    c = run("out", "--clip", "--bam", "--threads=3", sam)
    d = run("out", "-bc", "-t3", sam)
    assert a.returncode == b.returncode == c.returncode == d.returncode == 0
    body = lambda p: [ln for ln in run("view", "-", input=p.stdout).stdout.decode().splitlines() if not ln.startswith("@PG")]  # noqa: E731
    assert a.stdout[:4] == b"\x1f\x8b\x08\x04" and body(a) == body(b) == body(c) == body(d) and len(body(a)) > 400
    # sort: -n bundled with the container flag; view: value attached or separate
    s1 = run("sort", "-nb", sam).stdout
    s2 = run("sort", "-n", "--bam", sam).stdout
    assert run("view", "-", input=s1).stdout == run("view", "-t", "2", "-", input=s2).stdout != b""


@pytest.mark.parametrize("argv", [["out", "-x", "IN"], ["out", "--bogus", "IN"], ["extract", "-cb", "IN"], ["view", "--clip", "IN"],
                                  ["annotate", "--window", "5", "IN", "ref.fa"], ["out", "-c=1", "IN"], ["out", "IN", "extra"],
                                  ["annotate", "-t"], ["sort", "IN"]])
def test_unknown_or_malformed_options_are_errors(sam, argv):
    p = run(*[sam if a == "IN" else a for a in argv])
    assert p.returncode == 1 and p.stdout == b"" and p.stderr != b""


def test_b_and_u_are_exclusive(sam):
    p = run("view", "-bu", sam)
    assert p.returncode == 1 and b"exclusive" in p.stderr


def _bgzf_block(payload: bytes, isize=None) -> bytes:
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    c = co.compress(payload) + co.flush()
    head = b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(c) + 25)
    return head + c + struct.pack("<II", zlib.crc32(payload), len(payload) if isize is None else isize)


@pytest.mark.parametrize("bulk", [[], ["--bulk"]])
def test_block_that_claims_more_than_64k_is_damaged_input(tmp_path, bulk):
    hdr = b"BAM\x01" + struct.pack("<I", 0) + struct.pack("<I", 0)
    good = tmp_path / "ok.bam"
    good.write_bytes(_bgzf_block(hdr) + _bgzf_block(b""))
    assert run("view", *bulk, str(good)).returncode == 0
    bad = tmp_path / "bad.bam"
    bad.write_bytes(_bgzf_block(hdr, isize=0xF0000000) + _bgzf_block(b""))     # would be a 4 GB allocation
    p = run("view", *bulk, str(bad))
    assert p.returncode == 1 and (b"damaged" in p.stderr or b"cannot read" in p.stderr or b"not a BAM" in p.stderr)


def test_record_whose_fields_overrun_it_is_damaged(tmp_path):
    # one record: l_seq far beyond the record size
    name = b"r0\0"
    core = struct.pack("<iiBBHHHiiii", 0, 10, len(name), 60, 4680, 0, 0, 0x7ffffff0, -1, -1, 0) + name
    rec = struct.pack("<I", len(core)) + core
    text = b"@SQ\tSN:c\tLN:1000\n"
    hdr = b"BAM\x01" + struct.pack("<I", len(text)) + text + struct.pack("<I", 1) + struct.pack("<I", 2) + b"c\0" + struct.pack("<I", 1000)
    f = tmp_path / "overrun.bam"
    f.write_bytes(_bgzf_block(hdr + rec) + _bgzf_block(b""))
    p = run("view", str(f))
    assert p.returncode == 1


def test_simulator_bam_writer_matches_text_writer(tmp_path):
    """sim.write_bam (C++, parallel BGZF) writes the records tests/samio.py:write_sam writes as text"""
    names, contigs, cfg, _ = sim.config_c1()
    contigs = [contigs[0][:100_000]]
    rd = sim.make_reads(cfg, 0, 1500, contigs)
    sam, bam = tmp_path / "a.sam", tmp_path / "a.bam"
    samio.write_sam(sam, names, contigs, rd)
    sim.write_bam(str(bam), names, contigs, rd)
    p = subprocess.run([BIN, "view", str(bam)], capture_output=True)
    assert p.returncode == 0 and p.stdout.decode().splitlines() == open(sam).read().splitlines()


def test_builtin_deflate_encoder_is_read_back_by_zlib(tmp_path):
    """fastdeflate.hpp (the default block compressor of BAM output) on data of every kind: what zlib inflates from its
    BGZF stream is the input, block sizes respect the format, and it is not larger than zlib level 1 by more than a few
    per cent on BAM-like data."""
    import random
    import bamcodec
    rnd = random.Random(7)
    acgt = bytes(rnd.choice(b"ACGT") for _ in range(200_000))
    cases = {
        "empty": b"",
        "tiny": b"abc",
        "zeros": bytes(300_000),
        "random": bytes(rnd.getrandbits(8) for _ in range(150_000)),
        "acgt": acgt,
        "qualities": bytes(33 + rnd.randrange(40) for _ in range(150_000)),
        "periodic": bytes(i % 251 for i in range(200_000)),
        "far_copies": acgt[:40_000] + acgt[:40_000] + acgt[5_000:60_000],
        "block_edge": bytes(rnd.getrandbits(8) for _ in range(0xff00)) + b"x",
        "text": (b"@SQ\tSN:chr1\tLN:248956422\n" * 5000) + b"r1\t99\tchr1\t100\t60\t10S140M\t=\t300\t350\t" + acgt[:150] + b"\n",
    }
    for name, data in cases.items():
        src = tmp_path / f"{name}.bin"
        src.write_bytes(data)
        p = run("bgzf", str(src))
        assert p.returncode == 0, name
        assert bamcodec.bgzf_decode(p.stdout) == data, name
        assert p.stdout.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")), name   # EOF marker
        z1 = run("bgzf", "--level", "1", str(src))
        assert bamcodec.bgzf_decode(z1.stdout) == data
        if len(data) > 100_000:
            assert len(p.stdout) <= 1.08 * len(z1.stdout) + 1024, (name, len(p.stdout), len(z1.stdout))
    # through a pipe, and the same bytes whatever the thread count (blocks are compressed independently)
    a = run("bgzf", "-", input=cases["acgt"]).stdout
    b = run("bgzf", "-t", "1", "-", input=cases["acgt"]).stdout
    assert a == b and bamcodec.bgzf_decode(a) == cases["acgt"]


def test_reader_decodes_every_kind_of_deflate_block(sam, tmp_path):
    """BGZF blocks as other writers may produce them -- stored, fixed-Huffman, Huffman-only, several DEFLATE blocks per BGZF
    block -- are read identically by both readers (built-in decoder first, zlib behind it); a block whose CRC does not match
    its payload is damaged input."""
    import bamcodec
    ref = run("view", sam).stdout
    payload = bamcodec.bgzf_decode(run("view", "-u", sam).stdout)

    def block(data, level, strategy, split=False):
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        c = b""
        if split and len(data) > 100:
            c += co.compress(data[: len(data) // 3]) + co.flush(zlib.Z_FULL_FLUSH)
            c += co.compress(data[len(data) // 3:]) + co.flush(zlib.Z_SYNC_FLUSH)
            c += co.flush()
        else:
            c = co.compress(data) + co.flush()
        head = b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(c) + 25)
        return head + c + struct.pack("<II", zlib.crc32(data), len(data))

    eof = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    for name, (level, strategy, split, size) in {"stored": (0, zlib.Z_DEFAULT_STRATEGY, False, 60000), "fixed": (6, zlib.Z_FIXED, False, 30000),
                                                 "huffman": (6, zlib.Z_HUFFMAN_ONLY, False, 65000), "rle": (9, zlib.Z_RLE, False, 1000),
                                                 "multi": (6, zlib.Z_DEFAULT_STRATEGY, True, 50000), "tiny": (1, zlib.Z_DEFAULT_STRATEGY, False, 7)}.items():
        f = tmp_path / f"{name}.bam"
        f.write_bytes(b"".join(block(payload[a:a + size], level, strategy, split) for a in range(0, len(payload), size)) + eof)
        for bulk in ([], ["--bulk"]):
            p = run("view", *bulk, str(f))
            assert p.returncode == 0 and p.stdout == ref, (name, bulk, p.stderr)
        assert run("view", "--count", str(f)).stdout.strip() == b"400"
    bad = bytearray(block(payload[:40000], 6, zlib.Z_DEFAULT_STRATEGY) + eof)
    bad[-28 - 8] ^= 0x01                                   # the CRC of the first block
    g = tmp_path / "crc.bam"
    g.write_bytes(bytes(bad))
    for bulk in ([], ["--bulk"]):
        p = run("view", *bulk, str(g))
        assert p.returncode == 1 and p.stdout == b"" and (b"damaged" in p.stderr or b"cannot read" in p.stderr)
