#!/usr/bin/env python
"""Writes tests/golden/tiny_bam.json: four SAM records and the uncompressed BAM stream (hex) that
tests/bamcodec.py encodes them to.  tests/test_bam_io.py::test_golden_bam_bytes checks that BOTH codecs
(the Python one and fade_b200/csrc/host/samio.hpp) still produce exactly these bytes.  Pins drift of the two
codecs; it is derived from the SAM/BAM specification, not from htslib output (no htslib here)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import bamcodec  # noqa: E402

LINES = [
    "@HD\tVN:1.6\tSO:unsorted",
    "@SQ\tSN:chrA\tLN:1000",
    "@SQ\tSN:chrB\tLN:500",
    "r1\t99\tchrA\t101\t60\t5S10M2D5M\t=\t301\t215\tACGTNACGTNACGTNACGTN\tIIIIIIIIIIJJJJJJJJJJ\tNM:i:2\trs:i:3\tam:Z:chrA,95,5S15=;",
    "r1\t147\tchrA\t301\t60\t20M\t=\t101\t-215\tTTTTTTTTTTGGGGGGGGGG\t*\tXA:A:x\tXF:f:0.5\tXB:B:s,-3,300",
    "r2\t4\t*\t0\t0\t*\t*\t0\t0\tACG\t!!#\tXH:H:1AE3\tXI:i:-70000",
    "r3\t16\tchrB\t1\t255\t3M\tchrA\t7\t0\tacg\tABC\tXU:i:4000000000\tXS:i:65535",
]

if __name__ == "__main__":
    payload = bamcodec.bgzf_decode(bamcodec.encode(LINES, 0))
    norm = bamcodec.decode(bamcodec.encode(LINES, 0))
    with open(os.path.join(HERE, "tiny_bam.json"), "w") as f:
        json.dump({"sam": LINES, "sam_after_round_trip": norm, "bam_payload_hex": payload.hex()}, f, indent=1)
    print(len(payload), "bytes")
