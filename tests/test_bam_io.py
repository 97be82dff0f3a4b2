"""Record I/O of the C++ driver without htslib (fade_b200/csrc/host/samio.hpp): SAM text
and BGZF/BAM both ways, against the independent codec tests/bamcodec.py.  Pure host code, no GPU.
The reference gets this from dhtslib/htslib (anno.d:22, util.d:65-76); PARITY UNPINNED with respect to
htslib's exact bytes -- what is checked is the SAM/BAM specification and htslib's text conventions."""
import os
import random
import subprocess

import bamcodec
from test_cli_consumers import BIN, annotated_sam

HEADER = ["@HD\tVN:1.6\tSO:unsorted", "@SQ\tSN:chrA\tLN:100000", "@SQ\tSN:chrB\tLN:5000", "@RG\tID:g1\tSM:s",
          "@PG\tID:bwa\tPN:bwa\tCL:bwa mem ref.fa r.fq"]


def rich_lines(n=400, seed=9):
    """records exercising every field shape: unmapped, '*' SEQ/QUAL/CIGAR, RNEXT '=', long reads, every aux type"""
    rng = random.Random(seed)
    out = list(HEADER)
    for k in range(n):
        L = rng.choice([0, 1, 2, 7, 50, 151, 600])
        seq = "".join(rng.choice("ACGTNRYKM") for _ in range(L)) or "*"
        qual = "*" if L == 0 or k % 7 == 0 else "".join(chr(33 + rng.randrange(42)) for _ in range(L))
        unmapped = k % 11 == 0
        if unmapped or L == 0:
            cig = "*"
        elif L > 100 and k % 3 == 0:
            a, b = rng.randrange(1, 40), rng.randrange(0, 40)          # clips long enough to be realigned
            cig = f"{a}S{L - a - b}M" + (f"{b}S" if b else "")
        else:
            cig = f"3S{L - 5}M2S" if (L > 6 and k % 3 == 1) else f"{L}M"
        rname = "*" if unmapped else rng.choice(["chrA", "chrB"])
        pos = 0 if unmapped else rng.randrange(1, 4000)
        rnext = rng.choice(["*", "=", "chrB"]) if not unmapped else "*"
        if rnext == rname and rname != "*":
            rnext = "="                      # what htslib prints when the mate is on the same contig
        f = [f"read{k}/x", str(rng.choice([0, 4, 16, 99, 147, 2048, 2064])), rname, str(pos), str(rng.randrange(61)), cig, rnext,
             str(0 if rnext == "*" else rng.randrange(1, 5000)), str(rng.randrange(-900, 900)), seq, qual]
        f += [f"NM:i:{rng.choice([0, 3, 200, 255, 256, 65535, 65536, -1, -128, -129, -32768, -32769, 2000000000])}",
              f"XA:A:{rng.choice('abXY+')}", f"XF:f:{rng.choice(['0.5', '3', '-1.25', '1e+10', '0.001'])}",
              f"SA:Z:chrA,{k + 1},+,50M100S,60,0;", f"XH:H:{rng.choice(['1AE301', '', 'FF'])}",
              f"XB:B:{rng.choice(['c,-1,2,127', 'C,0,255', 's,-300,300', 'S,65535', 'i,-70000,70000', 'I,4000000000', 'f,0.5,-2'])}"]
        if k % 5 == 0:
            f += ["rs:i:3", "am:Z:chrA,100,50S100=;", "ab:Z:III;II"]
        out.append("\t".join(f))
    return out


def cli(args, data=None, path=None):
    p = subprocess.run([BIN, *args, str(path) if path else "-"], input=data, capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    return p.stdout


def test_sam_to_bam_matches_independent_codec():
    lines = rich_lines()
    sam = ("\n".join(lines) + "\n").encode()
    for flag, level in (("-b", 6), ("-u", 0)):
        bam = cli(["view", flag], sam)
        assert bamcodec.decode(bam) == lines                                     # our writer, their reader
        assert bgzf_payload(bam) == bgzf_payload(bamcodec.encode(lines, level))  # same uncompressed BAM stream, byte for byte
    assert bam[-28:] == bamcodec.EOF_BLOCK


def bgzf_payload(b):
    return bamcodec.bgzf_decode(b)


def test_bam_to_sam_matches_independent_codec():
    lines = rich_lines(seed=10)
    bam = bamcodec.encode(lines)
    assert cli(["view"], bam).decode().splitlines() == lines                     # their writer, our reader
    again = cli(["view", "-b"], bam)
    assert cli(["view"], again).decode().splitlines() == lines


def test_many_blocks_and_stdin_detection(tmp_path):
    lines = rich_lines(6000, seed=11)                                            # several MB: hundreds of BGZF blocks
    sam = ("\n".join(lines) + "\n").encode()
    p = tmp_path / "x.bam"
    p.write_bytes(cli(["view", "-b"], sam))
    assert cli(["view"], path=p).decode().splitlines() == lines
    assert cli(["view"], p.read_bytes()).decode().splitlines() == lines          # BAM on stdin is recognised by content


def test_bam_without_sq_text_uses_binary_reference_list():
    lines = [HEADER[0]] + [ln for ln in rich_lines(50, seed=12) if not ln.startswith("@")]
    full = bamcodec.encode(HEADER[:3] + lines[1:])
    # rebuild the header block with a text that lacks the @SQ lines but keep the binary reference list
    import struct
    d = bamcodec.bgzf_decode(full)
    l_text = struct.unpack_from("<I", d, 4)[0]
    text = (HEADER[0] + "\n").encode()
    d2 = b"BAM\1" + struct.pack("<I", len(text)) + text + d[8 + l_text:]
    got = cli(["view"], bamcodec.bgzf_blocks(d2)).decode().splitlines()
    assert got[0] == HEADER[0] and got[1:3] == HEADER[1:3]
    assert [ln for ln in got if not ln.startswith("@")] == lines[1:]


def test_damaged_bam_fails_loudly():
    bam = bytearray(cli(["view", "-b"], ("\n".join(rich_lines(300)) + "\n").encode()))
    bam[len(bam) // 2] ^= 0x55
    p = subprocess.run([BIN, "view", "-"], input=bytes(bam), capture_output=True)
    assert p.returncode != 0 and b"damaged" in p.stderr


def test_consumers_read_and_write_bam(tmp_path):
    """`out -c`, `out` and `extract` give the same records whether they read SAM or BAM and write SAM or BAM."""
    path, _ = annotated_sam(tmp_path, n=1500)
    sam = path.read_bytes()
    bam_in = tmp_path / "anno.bam"
    bam_in.write_bytes(bamcodec.encode(sam.decode().splitlines()))
    for args in (["out", "-c"], ["out"], ["extract"]):
        ref = cli(args, path=path).decode().splitlines()
        body = lambda ls: [ln for ln in ls if not ln.startswith("@PG\tID:fade-extract")]
        from_bam = cli(args, path=bam_in).decode().splitlines()
        assert body(from_bam) == body(ref)
        to_bam = bamcodec.decode(cli(args + ["-b"], path=path))
        assert body(to_bam) == body(ref)
        to_ubam = bamcodec.decode(cli(args + ["-u"], path=bam_in))
        assert body(to_ubam) == body(ref)


def test_parallel_readers_and_writers_of_the_annotate_loop():
    """`view --bulk` pushes every record through bamfast.hpp's I/O (parallel BGZF inflate / deflate, SAM text cut
    into lines and converted side by side) -- the part of `annotate` that needs no GPU.  Inputs large enough to
    cross the 8 MiB text refills and the 512-block inflate groups; results equal the single-threaded samio path."""
    lines = rich_lines(200000, seed=13)
    sam = ("\n".join(lines) + "\n").encode()
    bam = cli(["view", "-b"], sam)
    payload = bgzf_payload(bam)
    assert len(sam) > 48 << 20 and len(payload) > 520 * 0xff00                  # several text refills, > 512 BGZF blocks
    assert bamcodec.decode(bam[:0] + bam)[:2000] == lines[:2000]
    for src in (sam, bam):
        for t in ("1", "5"):
            assert cli(["view", "--bulk", "-t", t], src) == sam
            assert bgzf_payload(cli(["view", "--bulk", "-b", "-t", t], src)) == payload
        assert bgzf_payload(cli(["view", "--bulk", "-u"], src)) == payload
    # CRLF line ends and blank lines in SAM text
    crlf = ("\r\n".join(lines[:300]) + "\r\n\r\n").encode()
    assert cli(["view", "--bulk"], crlf).decode().splitlines() == lines[:300]
    # loud failures
    broken = bytearray(bam)
    broken[len(broken) // 3] ^= 0x41
    p = subprocess.run([BIN, "view", "--bulk", "-"], input=bytes(broken), capture_output=True)
    assert p.returncode != 0 and b"damaged" in p.stderr
    p = subprocess.run([BIN, "view", "--bulk", "-"], input=bam[: len(bam) // 2], capture_output=True)
    assert p.returncode != 0
    bad_sam = ("\n".join(lines[:50] + ["name\t0\tchrA"] + lines[50:60]) + "\n").encode()
    p = subprocess.run([BIN, "view", "--bulk", "-b", "-"], input=bad_sam, capture_output=True)
    assert p.returncode != 0 and b"malformed" in p.stderr


def test_codec_fuzz_against_independent_codec():
    """hypothesis: random well-formed SAM records (field extremes, every aux type, boundary integers) survive
    SAM -> BAM (ours) -> SAM (theirs) and SAM -> BAM (theirs) -> SAM (ours), and both BAM payloads are identical."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    name = st.text(alphabet=st.characters(min_codepoint=33, max_codepoint=126, blacklist_characters="@\t"), min_size=1, max_size=40)
    ints = st.one_of(st.sampled_from([0, 1, 255, 256, 65535, 65536, 2**31 - 1, 2**32 - 1, -1, -128, -129, -32768, -32769, -2**31]),
                     st.integers(-2**31, 2**32 - 1))
    floats = st.integers(-4000, 4000).map(lambda k: "%g" % (k / 8))
    ztext = st.text(alphabet=st.characters(min_codepoint=32, max_codepoint=126), max_size=30)
    tag = st.text(alphabet="ABCDEFGHXYZabcxyz", min_size=2, max_size=2)

    @st.composite
    def aux(draw):
        kind = draw(st.sampled_from("AifZHB"))
        t = draw(tag)
        if kind == "A":
            return f"{t}:A:{draw(st.sampled_from('!aZ9~+'))}"
        if kind == "i":
            return f"{t}:i:{draw(ints)}"
        if kind == "f":
            return f"{t}:f:{draw(floats)}"
        if kind == "Z":
            return f"{t}:Z:{draw(ztext)}"
        if kind == "H":
            return f"{t}:H:{draw(st.binary(max_size=8)).hex().upper()}"
        sub = draw(st.sampled_from("cCsSiIf"))
        lo, hi = {"c": (-128, 127), "C": (0, 255), "s": (-32768, 32767), "S": (0, 65535), "i": (-2**31, 2**31 - 1), "I": (0, 2**32 - 1), "f": (0, 0)}[sub]
        vals = draw(st.lists(floats if sub == "f" else st.integers(lo, hi).map(str), max_size=6))
        return f"{t}:B:{sub}" + "".join("," + v for v in vals)

    @st.composite
    def record(draw):
        L = draw(st.sampled_from([0, 1, 2, 3, 10, 33, 150]))
        seq = "".join(draw(st.lists(st.sampled_from("=ACMGRSVTWYHKDBN"), min_size=L, max_size=L))) or "*"
        qual = "*" if L == 0 or draw(st.booleans()) else "".join(chr(33 + q) for q in draw(st.lists(st.integers(0, 93), min_size=L, max_size=L)))
        mapped = L > 0 and draw(st.booleans())
        if mapped:
            s1 = draw(st.integers(0, L - 1))
            cig = (f"{s1}S" if s1 else "") + f"{L - s1}M" + draw(st.sampled_from(["", "5D", "100N2P"]))
        else:
            cig = "*"
        rname = draw(st.sampled_from(["chrA", "chrB"])) if mapped else "*"
        rnext = draw(st.sampled_from(["*", "=", "chrA", "chrB"])) if mapped else draw(st.sampled_from(["*", "chrB"]))
        if rnext == rname and rname != "*":
            rnext = "="
        tags, seen = [], set()
        for a in draw(st.lists(aux(), max_size=6)):
            if a[:2] not in seen:
                seen.add(a[:2]); tags.append(a)
        return "\t".join([draw(name), str(draw(st.integers(0, 65535))), rname, str(draw(st.integers(1, 2**29)) if mapped else 0),
                          str(draw(st.integers(0, 255))), cig, rnext, str(draw(st.integers(0, 2**29))), str(draw(st.integers(-2**29, 2**29))),
                          seq, qual] + tags)

    @settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.lists(record(), min_size=1, max_size=40))
    def run(recs):
        lines = HEADER + recs
        sam = ("\n".join(lines) + "\n").encode()
        ours = cli(["view", "-u"], sam)
        assert bamcodec.decode(ours) == lines
        theirs = bamcodec.encode(lines, 0)
        assert bgzf_payload(ours) == bgzf_payload(theirs)
        assert cli(["view"], theirs).decode().splitlines() == lines
        assert cli(["view", "--bulk"], theirs).decode().splitlines() == lines

    run()


def test_fasta_loader_with_and_without_fai(tmp_path):
    """The reference (anno.d:23 IndexedFastaFile) is read whole: through <fasta>.fai when there is one (one read
    per contig), line by line otherwise.  Same contigs either way: LF / CRLF, last line with and without a line
    end, contig lengths around the line width."""
    import random
    rng = random.Random(3)
    for nl, final_nl in (("\n", True), ("\n", False), ("\r\n", True)):
        contigs = [(f"c{k} some description", "".join(rng.choice("ACGTNacgtn") for _ in range(n)))
                   for k, n in enumerate((1, 59, 60, 61, 120, 1234, 60 * 50))]
        text, fai, off = "", [], 0
        for name, seq in contigs:
            head = f">{name}{nl}"
            off += len(head)
            body = nl.join(seq[a:a + 60] for a in range(0, len(seq), 60)) + nl
            fai.append(f"{name.split()[0]}\t{len(seq)}\t{off}\t60\t{60 + len(nl)}")
            text += head + body
            off += len(body)
        if not final_nl:
            text = text[: -len(nl)]
        fa = tmp_path / f"r{len(nl)}{int(final_nl)}.fa"
        fa.write_bytes(text.encode())
        plain = subprocess.run([BIN, "fasta-digest", str(fa)], capture_output=True, text=True)
        (tmp_path / (fa.name + ".fai")).write_text("\n".join(fai) + "\n")
        indexed = subprocess.run([BIN, "fasta-digest", str(fa)], capture_output=True, text=True)
        assert plain.returncode == 0 and indexed.returncode == 0
        assert plain.stdout == indexed.stdout
        rows = [ln.split("\t") for ln in indexed.stdout.splitlines()]
        assert {r[0]: int(r[1]) for r in rows} == {n.split()[0]: len(s) for n, s in contigs}
    # a .fai that does not match the file is refused and the text is read instead
    (tmp_path / (fa.name + ".fai")).write_text("c0\t5\t999999\t60\t61\n")
    again = subprocess.run([BIN, "fasta-digest", str(fa)], capture_output=True, text=True)
    assert again.stdout == plain.stdout


def test_golden_bam_bytes():
    """tests/golden/tiny_bam.json (written by tests/golden/make_bam_golden.py): both codecs still produce the
    committed uncompressed BAM stream from the committed SAM lines, and decode it to the committed text
    (lower-case bases come back upper-case, as through htslib)."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_bam.json")))
    sam = ("\n".join(g["sam"]) + "\n").encode()
    want = bytes.fromhex(g["bam_payload_hex"])
    assert bgzf_payload(bamcodec.encode(g["sam"], 0)) == want
    assert bgzf_payload(cli(["view", "-u"], sam)) == want
    assert bgzf_payload(cli(["view", "--bulk", "-b"], sam)) == want
    assert cli(["view"], bamcodec.bgzf_blocks(want)).decode().splitlines() == g["sam_after_round_trip"]
    assert g["sam_after_round_trip"][-1].split("\t")[9] == "ACG" and g["sam"][-1].split("\t")[9] == "acg"
