"""Hand-built read sets for the parity tests (ragged lengths, wildcard letters, contig ends...)."""
from __future__ import annotations

import random
from types import SimpleNamespace

import numpy as np

from oracle import oracle as orc

COMP = str.maketrans("ACGTNRYKMSWBDHVacgtn", "TGCANYRMKSWVHDBtgcan")


def revcomp(s: str) -> str:
    return s.translate(COMP)[::-1]


def build(reads):
    """reads: list of dict(seq=str, tid=int, pos=int, aligned_len=int, clip_left=int, clip_right=int)."""
    n = len(reads)
    packed = [orc.pack_nt16(r["seq"]) for r in reads]
    seq_off = np.zeros(n + 1, dtype=np.int64)
    for k, p in enumerate(packed):
        seq_off[k + 1] = seq_off[k] + len(p)
    seq4 = np.concatenate(packed) if packed else np.zeros(0, dtype=np.uint8)
    g = lambda key, dt: np.array([r[key] for r in reads], dtype=dt)  # noqa: E731
    return SimpleNamespace(n=n, seq4=seq4, seq_off=seq_off, l_qseq=np.array([len(r["seq"]) for r in reads], dtype=np.int32),
                           tid=g("tid", np.int32), pos=g("pos", np.int64), aligned_len=g("aligned_len", np.int32),
                           clip_left=g("clip_left", np.int32), clip_right=g("clip_right", np.int32))


def random_ref(rng: random.Random, n: int, alpha="ACGT") -> bytes:
    return "".join(rng.choice(alpha) for _ in range(n)).encode()


def ragged_reads(rng: random.Random, contigs: list[bytes], n: int, min_len=20, max_len=250, window=300,
                 wild_read_rate=0.0, alpha="ACGT"):
    """Reads of random length with artifact / random clips at random places, including reads
    hanging over contig starts and ends and clips at or below the length floor."""
    out = []
    for _ in range(n):
        tid = rng.randrange(len(contigs))
        ref = contigs[tid].decode()
        L = rng.randint(min_len, max_len)
        cl = rng.choice([0, 0, rng.randint(1, min(60, L // 2))])
        cr = rng.choice([0, 0, rng.randint(1, min(60, L // 2 - 1))]) if L > 4 else 0
        m = L - cl - cr
        if m < 1:
            cl, cr, m = 0, 0, L
        edge = rng.random()
        if edge < 0.1:
            pos = rng.randint(0, min(40, len(ref) - m))
        elif edge < 0.2:
            pos = rng.randint(max(0, len(ref) - m - 40), len(ref) - m)
        else:
            pos = rng.randint(0, len(ref) - m)
        body = ref[pos:pos + m].upper()
        parts = []
        for side, ln in ((0, cl), (1, cr)):
            if ln == 0:
                parts.append("")
                continue
            if rng.random() < 0.6:
                lo = max(0, pos - window)
                hi = min(len(ref), pos + m + window)
                o = rng.randint(lo, max(lo, hi - ln))
                seg = ref[o:o + ln].upper()
                seg = seg + "".join(rng.choice("ACGT") for _ in range(ln - len(seg)))
                parts.append(revcomp(seg))
            else:
                parts.append("".join(rng.choice(alpha) for _ in range(ln)))
        seq = list(parts[0] + body + parts[1])
        for i in range(len(seq)):
            if rng.random() < 0.01:
                seq[i] = rng.choice("ACGT")
            if wild_read_rate and rng.random() < wild_read_rate:
                seq[i] = rng.choice("RYKMSWN")
        out.append(dict(seq="".join(seq), tid=tid, pos=pos, aligned_len=m, clip_left=cl, clip_right=cr))
    return out
