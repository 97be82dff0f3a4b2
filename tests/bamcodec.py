"""A small, independent BAM codec for the tests (struct + zlib only; no htslib, no code shared with
fade_b200/csrc/host/samio.hpp): SAM text lines <-> BGZF/BAM bytes following SAMv1 sections 4.1-4.2,
with htslib's text conventions (smallest integer aux type, non-negative values unsigned; %g floats;
RNEXT '=' when equal to RNAME; '*' qualities = 0xff)."""
from __future__ import annotations

import struct
import zlib

OPS = "MIDNSHP=XB"
NT16 = "=ACMGRSVTWYHKDBN"
EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_blocks(data: bytes, level: int = 6, block: int = 0xff00) -> bytes:
    out = bytearray()
    for a in range(0, len(data), block):
        chunk = data[a:a + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        c = co.compress(chunk) + co.flush()
        out += struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, ord("B"), ord("C"), 2, len(c) + 25)
        out += c + struct.pack("<II", zlib.crc32(chunk), len(chunk))
    return bytes(out) + EOF_BLOCK


def bgzf_decode(raw: bytes) -> bytes:
    out = bytearray()
    p = 0
    while p < len(raw):
        assert raw[p:p + 4] == b"\x1f\x8b\x08\x04", "not a BGZF block"
        xlen = struct.unpack_from("<H", raw, p + 10)[0]
        extra = raw[p + 12:p + 12 + xlen]
        bsize = None
        q = 0
        while q + 4 <= len(extra):
            si1, si2, slen = struct.unpack_from("<BBH", extra, q)
            if (si1, si2, slen) == (66, 67, 2):
                bsize = struct.unpack_from("<H", extra, q + 4)[0]
            q += 4 + slen
        assert bsize is not None
        cdata = raw[p + 12 + xlen:p + bsize + 1 - 8]
        crc, isize = struct.unpack_from("<II", raw, p + bsize + 1 - 8)
        d = zlib.decompress(cdata, -15)
        assert len(d) == isize and zlib.crc32(d) == crc
        out += d
        p += bsize + 1
    return bytes(out)


def _reg2bin(beg: int, end: int) -> int:
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def _int_aux(v: int) -> bytes:
    if v < 0:
        return (b"c" + struct.pack("<b", v)) if v >= -128 else (b"s" + struct.pack("<h", v)) if v >= -32768 else b"i" + struct.pack("<i", v)
    return (b"C" + struct.pack("<B", v)) if v <= 255 else (b"S" + struct.pack("<H", v)) if v <= 65535 else b"I" + struct.pack("<I", v)


def encode_record(line: str, tid_of: dict) -> bytes:
    f = line.split("\t")
    name, flag, rname, pos, mapq, cig, rnext, pnext, tlen, seq, qual = f[:11]
    tid = -1 if rname == "*" else tid_of.get(rname, -1)
    mtid = -1 if rnext == "*" else tid if rnext == "=" else tid_of.get(rnext, -1)
    ops, ref_len, num = [], 0, ""
    if cig != "*":
        for ch in cig:
            if ch.isdigit():
                num += ch
            else:
                op = OPS.index(ch)
                ops.append((int(num) << 4) | op)
                if op in (0, 2, 3, 7, 8):
                    ref_len += int(num)
                num = ""
    l_seq = 0 if seq == "*" else len(seq)
    p0 = int(pos) - 1
    b = struct.pack("<iiBBHHHiiii", tid, p0, len(name) + 1, int(mapq), _reg2bin(p0, p0 + (ref_len or 1)), len(ops), int(flag),
                    l_seq, mtid, int(pnext) - 1, int(tlen))
    b += name.encode() + b"\0" + b"".join(struct.pack("<I", o) for o in ops)
    packed = bytearray((l_seq + 1) // 2)
    for i in range(l_seq):
        c = NT16.find(seq[i].upper())
        packed[i >> 1] |= (c if c >= 0 else 15) << (4 if i % 2 == 0 else 0)
    b += bytes(packed)
    b += (b"\xff" * l_seq) if qual == "*" else bytes(ord(c) - 33 for c in qual)
    for a in f[11:]:
        tag, ty, val = a[:2], a[3], a[5:]
        b += tag.encode()
        if ty == "A":
            b += b"A" + val[0].encode()
        elif ty == "i":
            b += _int_aux(int(val))
        elif ty == "f":
            b += b"f" + struct.pack("<f", float(val))
        elif ty in "ZH":
            b += ty.encode() + val.encode() + b"\0"
        elif ty == "B":
            st, items = val[0], [x for x in val[2:].split(",") if x] if len(val) > 1 else []
            fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[st]
            b += b"B" + st.encode() + struct.pack("<I", len(items))
            b += b"".join(struct.pack("<" + fmt, float(x) if st == "f" else int(x)) for x in items)
        else:
            raise ValueError(a)
    return struct.pack("<I", len(b)) + b


def encode(lines, level: int = 6) -> bytes:
    """SAM text lines (header first) -> BAM file bytes."""
    head = [ln for ln in lines if ln.startswith("@")]
    names, lens = [], []
    for ln in head:
        if ln.startswith("@SQ"):
            d = dict(x.split(":", 1) for x in ln.split("\t")[1:])
            names.append(d["SN"]); lens.append(int(d["LN"]))
    text = "".join(ln + "\n" for ln in head).encode()
    data = b"BAM\1" + struct.pack("<I", len(text)) + text + struct.pack("<I", len(names))
    for n, ln in zip(names, lens):
        data += struct.pack("<I", len(n) + 1) + n.encode() + b"\0" + struct.pack("<I", ln)
    tid_of = {n: i for i, n in enumerate(names)}
    first = bgzf_blocks(data, level)[:-len(EOF_BLOCK)]
    body = b"".join(encode_record(ln, tid_of) for ln in lines if not ln.startswith("@"))
    return first + bgzf_blocks(body, level)


def _g(v: float) -> str:
    return "%g" % v


def decode(raw: bytes):
    """BAM file bytes -> SAM text lines (header first)."""
    d = bgzf_decode(raw)
    assert d[:4] == b"BAM\1"
    l_text = struct.unpack_from("<I", d, 4)[0]
    text = d[8:8 + l_text].split(b"\0")[0].decode()
    p = 8 + l_text
    n_ref = struct.unpack_from("<I", d, p)[0]
    p += 4
    names = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<I", d, p)[0]
        names.append(d[p + 4:p + 4 + ln - 1].decode())
        p += 4 + ln + 4
    lines = [x for x in text.split("\n") if x]
    nm = lambda t: names[t] if 0 <= t < len(names) else "*"
    while p < len(d):
        bs = struct.unpack_from("<I", d, p)[0]
        r = d[p + 4:p + 4 + bs]
        p += 4 + bs
        tid, pos, l_name, mapq, _bin, n_cig, flag, l_seq, mtid, mpos, tlen = struct.unpack_from("<iiBBHHHiiii", r, 0)
        q = 32
        name = r[q:q + l_name - 1].decode(); q += l_name
        cig = "".join(f"{c >> 4}{OPS[c & 15]}" for c in struct.unpack_from(f"<{n_cig}I", r, q)) or "*"; q += 4 * n_cig
        seq = "".join(NT16[(r[q + (i >> 1)] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq)) or "*"; q += (l_seq + 1) // 2
        qual = "*" if l_seq == 0 or r[q] == 0xff else "".join(chr(c + 33) for c in r[q:q + l_seq]); q += l_seq
        f = [name, str(flag), nm(tid), str(pos + 1), str(mapq), cig, "*" if mtid < 0 else "=" if mtid == tid else nm(mtid),
             str(mpos + 1), str(tlen), seq, qual]
        while q < len(r):
            tag, ty = r[q:q + 2].decode(), chr(r[q + 2]); q += 3
            if ty == "A":
                f.append(f"{tag}:A:{chr(r[q])}"); q += 1
            elif ty in "cCsSiI":
                fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I"}[ty]
                f.append(f"{tag}:i:{struct.unpack_from('<' + fmt, r, q)[0]}"); q += struct.calcsize(fmt)
            elif ty == "f":
                f.append(f"{tag}:f:{_g(struct.unpack_from('<f', r, q)[0])}"); q += 4
            elif ty in "ZH":
                e = r.index(b"\0", q)
                f.append(f"{tag}:{ty}:{r[q:e].decode()}"); q = e + 1
            elif ty == "B":
                st, cnt = chr(r[q]), struct.unpack_from("<I", r, q + 1)[0]; q += 5
                fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[st]
                vals = struct.unpack_from(f"<{cnt}{fmt}", r, q); q += cnt * struct.calcsize(fmt)
                f.append(f"{tag}:B:{st}" + "".join("," + (_g(v) if st == "f" else str(v)) for v in vals))
            else:
                raise ValueError(ty)
        lines.append("\t".join(f))
    return lines
