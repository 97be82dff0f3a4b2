// fastdeflate.hpp -- a small, fast DEFLATE encoder for BGZF blocks (RFC 1951, one dynamic-Huffman block per call).
//
// Why: the record loop of `fade-b200 annotate -b` spent 60 % of its time in zlib's deflate (level 6: 15-25 MB/s per
// core on BAM records, whose qualities and packed bases are close to incompressible, so the longer match search buys
// 2-3 % of size).  This encoder does what matters for such data -- entropy coding with per-block optimal Huffman codes
// and a single-probe LZ77 match finder for the repetitive parts (names, tags, CIGARs) -- at several times the speed.
// Any inflater reads its output (the tests inflate it with zlib).  zlib stays in use for inflate and for --level 1..9.
//
// compress(): src[0, n) with n <= 65535 -> raw DEFLATE stream in dst (capacity >= n + 16); falls back to a stored
// block when the data does not compress.
#pragma once
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace fastdeflate {

namespace detail {

constexpr int kMaxLitLen = 286, kMaxDist = 30, kMaxCl = 19;

struct BitWriter {
    uint8_t *p, *end;
    uint64_t acc = 0;
    int nbits = 0;
    bool overflow = false;
    BitWriter(uint8_t *dst, size_t cap) : p(dst), end(dst + cap) {}
    inline void put(uint32_t v, int n)   // n <= 32, LSB first
    {
        acc |= (uint64_t)v << nbits;
        nbits += n;
        if (nbits >= 32) {
            if (p + 4 > end) { overflow = true; nbits -= 32; acc >>= 32; return; }
            const uint32_t w = (uint32_t)acc;
            memcpy(p, &w, 4);            // little endian host (x86-64 / aarch64)
            p += 4;
            acc >>= 32;
            nbits -= 32;
        }
    }
    inline void flush_byte()
    {
        while (nbits > 0) {
            if (p >= end) { overflow = true; return; }
            *p++ = (uint8_t)acc;
            acc >>= 8;
            nbits -= 8;
        }
        nbits = 0;
        acc = 0;
    }
};

inline uint32_t rev_bits(uint32_t code, int len)
{
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) { r = (r << 1) | (code & 1); code >>= 1; }
    return r;
}

// Code lengths of a minimum-redundancy prefix code, limited to max_bits (Moffat & Katajainen in-place algorithm on the
// sorted frequencies, then the Kraft-sum repair miniz / zlib use for the length limit).
inline void huffman_lengths(const uint32_t *freq, int n_sym, int max_bits, uint8_t *len)
{
    struct SF { uint32_t f; uint16_t s; };
    SF a[kMaxLitLen + 2];
    int n = 0;
    for (int s = 0; s < n_sym; ++s) { len[s] = 0; if (freq[s]) { a[n].f = freq[s]; a[n].s = (uint16_t)s; ++n; } }
    if (n == 0) return;
    if (n == 1) { len[a[0].s] = 1; return; }
    std::sort(a, a + n, [](const SF &x, const SF &y) { return x.f < y.f || (x.f == y.f && x.s < y.s); });
    // Moffat: w[i] holds, in turn, internal-node weights, parent indices and depths
    uint32_t w[kMaxLitLen + 2] = { 0 };
    for (int i = 0; i < n; ++i) w[i] = a[i].f;
    w[0] += w[1];
    int root = 0, leaf = 2, next;
    for (next = 1; next < n - 1; ++next) {
        if (leaf >= n || w[root] < w[leaf]) { w[next] = w[root]; w[root++] = (uint32_t)next; } else w[next] = w[leaf++];
        if (leaf >= n || (root < next && w[root] < w[leaf])) { w[next] += w[root]; w[root++] = (uint32_t)next; } else w[next] += w[leaf++];
    }
    w[n - 2] = 0;
    for (next = n - 3; next >= 0; --next) w[next] = w[w[next]] + 1;
    int avbl = 1, used = 0, dpth = 0;
    root = n - 2; next = n - 1;
    while (avbl > 0) {
        while (root >= 0 && (int)w[root] == dpth) { ++used; --root; }
        while (avbl > used) { w[next--] = (uint32_t)dpth; --avbl; }
        avbl = 2 * used; ++dpth; used = 0;
    }
    // w[i] = depth of the i-th least frequent symbol; enforce the limit
    int num[33] = { 0 };
    for (int i = 0; i < n; ++i) ++num[std::min<int>((int)w[i], 32)];
    for (int i = max_bits + 1; i <= 32; ++i) { num[max_bits] += num[i]; num[i] = 0; }
    uint32_t total = 0;
    for (int i = max_bits; i > 0; --i) total += (uint32_t)num[i] << (max_bits - i);
    while (total != (1u << max_bits)) {
        --num[max_bits];
        for (int i = max_bits - 1; i > 0; --i)
            if (num[i]) { --num[i]; num[i + 1] += 2; break; }
        --total;
    }
    // longest codes to the least frequent symbols
    int k = 0;
    for (int l = max_bits; l >= 1; --l)
        for (int c = 0; c < num[l]; ++c) len[a[k++].s] = (uint8_t)l;
}

inline void canonical_codes(const uint8_t *len, int n_sym, uint16_t *code)
{
    int bl_count[16] = { 0 };
    for (int s = 0; s < n_sym; ++s) ++bl_count[len[s]];
    bl_count[0] = 0;
    uint32_t next_code[16];
    uint32_t c = 0;
    for (int b = 1; b <= 15; ++b) { c = (c + (uint32_t)bl_count[b - 1]) << 1; next_code[b] = c; }
    for (int s = 0; s < n_sym; ++s) code[s] = len[s] ? (uint16_t)rev_bits(next_code[len[s]]++, len[s]) : 0;
}

struct Tables {
    uint8_t len_sym[259];      // match length -> symbol - 257
    uint8_t len_ebits[29];
    uint16_t len_base[29];
    uint8_t dist_ebits[30];
    uint16_t dist_base[30];
    uint8_t dist_sym_lo[257];  // distance 1..256 -> symbol
    uint8_t dist_sym_hi[256];  // (distance - 1) >> 7 -> symbol, for distances 257..32768
    Tables()
    {
        static const uint8_t leb[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
        int base = 3;
        for (int s = 0; s < 29; ++s) {
            len_ebits[s] = leb[s];
            len_base[s] = (uint16_t)(s == 28 ? 258 : base);
            if (s < 28) { for (int k = 0; k < (1 << leb[s]) && base + k <= 258; ++k) len_sym[base + k] = (uint8_t)s; base += 1 << leb[s]; }
        }
        len_sym[258] = 28;
        int db = 1;
        for (int s = 0; s < 30; ++s) {
            const int eb = s < 4 ? 0 : (s - 2) / 2;
            dist_ebits[s] = (uint8_t)eb;
            dist_base[s] = (uint16_t)db;
            for (int k = 0; k < (1 << eb); ++k) {
                const int d = db + k;
                if (d <= 256) dist_sym_lo[d] = (uint8_t)s;
                else dist_sym_hi[(d - 1) >> 7] = (uint8_t)s;
            }
            db += 1 << eb;
        }
    }
    inline int dist_sym(uint32_t d) const { return d <= 256 ? dist_sym_lo[d] : dist_sym_hi[(d - 1) >> 7]; }
};

inline const Tables &tables()
{
    static const Tables t;
    return t;
}

inline uint32_t load32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }

}  // namespace detail

inline size_t stored_block(const uint8_t *src, size_t n, uint8_t *dst)
{
    dst[0] = 1;   // BFINAL = 1, BTYPE = 00
    dst[1] = (uint8_t)(n & 0xff); dst[2] = (uint8_t)(n >> 8);
    dst[3] = (uint8_t)(~n & 0xff); dst[4] = (uint8_t)((~n >> 8) & 0xff);
    if (n) memcpy(dst + 5, src, n);
    return n + 5;
}

// returns the number of bytes written to dst; dst must hold n + 16 bytes
inline size_t compress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap)
{
    using namespace detail;
    if (n > 65535 || cap < n + 16) return 0;
    if (n < 32) return stored_block(src, n, dst);
    const Tables &T = tables();
    // ---- 1. LZ77: single-probe hash of 4-byte sequences, greedy ----
    constexpr int HB = 13;
    uint16_t head[1 << HB];
    memset(head, 0, sizeof(head));
    static thread_local uint32_t tok[65536];     // literal: byte | match: 1<<31 | (len-3) << 16 | (dist-1)
    uint32_t lfreq[kMaxLitLen] = { 0 }, dfreq[kMaxDist] = { 0 };
    size_t nt = 0, i = 0;
    const size_t last = n >= 4 ? n - 4 : 0;
    uint32_t miss = 0;
    while (i < n) {
        if (i <= last) {
            const uint32_t v = load32(src + i);
            const uint32_t h = (v * 2654435761u) >> (32 - HB);
            const uint32_t cand = head[h];
            head[h] = (uint16_t)(i + 1);
            if (cand && load32(src + cand - 1) == v && i + 1 - cand <= 32768) {
                const size_t c0 = cand - 1;
                const size_t lim = std::min<size_t>(258, n - i);
                size_t len = 4;
                while (len + 8 <= lim) {
                    uint64_t x, y;
                    memcpy(&x, src + i + len, 8); memcpy(&y, src + c0 + len, 8);
                    if (x != y) { len += (size_t)(__builtin_ctzll(x ^ y) >> 3); goto matched; }
                    len += 8;
                }
                while (len < lim && src[i + len] == src[c0 + len]) ++len;
matched:
                if (len > lim) len = lim;
                const uint32_t dist = (uint32_t)(i - c0);
                tok[nt++] = 0x80000000u | ((uint32_t)(len - 3) << 16) | (dist - 1);
                ++lfreq[257 + T.len_sym[len]];
                ++dfreq[T.dist_sym(dist)];
                // a few positions of the match enter the hash table so that following text can refer to it
                if (i + len <= last) {
                    const size_t e = i + len;
                    for (size_t k = i + 1; k < e && k < i + 4; ++k) head[(load32(src + k) * 2654435761u) >> (32 - HB)] = (uint16_t)(k + 1);
                    head[(load32(src + e - 1) * 2654435761u) >> (32 - HB)] = (uint16_t)e;
                }
                i += len;
                miss = 0;
                continue;
            }
        }
        // literal(s); in incompressible stretches (qualities, packed bases) probe less and less often
        const size_t step = 1 + (miss >> 6);
        ++miss;
        for (size_t k = 0; k < step && i < n; ++k, ++i) { tok[nt++] = src[i]; ++lfreq[src[i]]; }
    }
    lfreq[256] = 1;
    // at least two distance codes (a decoder wants a complete, non-trivial code; zlib does the same)
    { int used = 0; for (int s = 0; s < kMaxDist; ++s) used += dfreq[s] != 0; for (int s = 0; used < 2 && s < 2; ++s) if (!dfreq[s]) { dfreq[s] = 1; ++used; } }
    // ---- 2. Huffman codes ----
    uint8_t llen[kMaxLitLen], dlen[kMaxDist];
    uint16_t lcode[kMaxLitLen], dcode[kMaxDist];
    huffman_lengths(lfreq, kMaxLitLen, 15, llen);
    huffman_lengths(dfreq, kMaxDist, 15, dlen);
    canonical_codes(llen, kMaxLitLen, lcode);
    canonical_codes(dlen, kMaxDist, dcode);
    int hlit = kMaxLitLen, hdist = kMaxDist;
    while (hlit > 257 && !llen[hlit - 1]) --hlit;
    while (hdist > 1 && !dlen[hdist - 1]) --hdist;
    // code-length sequence, run-length coded (symbols 16 / 17 / 18)
    uint8_t seq[kMaxLitLen + kMaxDist];
    int ns = 0;
    for (int s = 0; s < hlit; ++s) seq[ns++] = llen[s];
    for (int s = 0; s < hdist; ++s) seq[ns++] = dlen[s];
    uint8_t rsym[kMaxLitLen + kMaxDist], rext[kMaxLitLen + kMaxDist];
    int nr = 0;
    uint32_t cfreq[kMaxCl] = { 0 };
    for (int k = 0; k < ns;) {
        const int v = seq[k];
        int run = 1;
        while (k + run < ns && seq[k + run] == v) ++run;
        int left = run;
        if (v == 0) {
            while (left >= 11) { const int r = std::min(left, 138); rsym[nr] = 18; rext[nr++] = (uint8_t)(r - 11); ++cfreq[18]; left -= r; }
            if (left >= 3) { rsym[nr] = 17; rext[nr++] = (uint8_t)(left - 3); ++cfreq[17]; left = 0; }
            while (left-- > 0) { rsym[nr] = 0; rext[nr++] = 0; ++cfreq[0]; }
        } else {
            rsym[nr] = (uint8_t)v; rext[nr++] = 0; ++cfreq[v]; --left;
            while (left >= 3) { const int r = std::min(left, 6); rsym[nr] = 16; rext[nr++] = (uint8_t)(r - 3); ++cfreq[16]; left -= r; }
            while (left-- > 0) { rsym[nr] = (uint8_t)v; rext[nr++] = 0; ++cfreq[v]; }
        }
        k += run;
    }
    uint8_t clen[kMaxCl];
    uint16_t ccode[kMaxCl];
    { int used = 0; for (int s = 0; s < kMaxCl; ++s) used += cfreq[s] != 0; if (used < 2) { for (int s = 0; used < 2 && s < kMaxCl; ++s) if (!cfreq[s]) { cfreq[s] = 1; ++used; } } }
    huffman_lengths(cfreq, kMaxCl, 7, clen);
    canonical_codes(clen, kMaxCl, ccode);
    static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    int hclen = 19;
    while (hclen > 4 && !clen[order[hclen - 1]]) --hclen;
    // ---- 3. bit stream ----
    BitWriter bw(dst, cap);
    bw.put(1, 1);            // BFINAL
    bw.put(2, 2);            // BTYPE = dynamic Huffman
    bw.put((uint32_t)(hlit - 257), 5);
    bw.put((uint32_t)(hdist - 1), 5);
    bw.put((uint32_t)(hclen - 4), 4);
    for (int k = 0; k < hclen; ++k) bw.put(clen[order[k]], 3);
    for (int k = 0; k < nr; ++k) {
        bw.put(ccode[rsym[k]], clen[rsym[k]]);
        if (rsym[k] == 16) bw.put(rext[k], 2);
        else if (rsym[k] == 17) bw.put(rext[k], 3);
        else if (rsym[k] == 18) bw.put(rext[k], 7);
    }
    for (size_t k = 0; k < nt; ++k) {
        const uint32_t t = tok[k];
        if (!(t & 0x80000000u)) { bw.put(lcode[t], llen[t]); continue; }
        const uint32_t len = ((t >> 16) & 0x1ff) + 3, dist = (t & 0xffff) + 1;
        const int ls = T.len_sym[len];
        bw.put(lcode[257 + ls], llen[257 + ls]);
        if (T.len_ebits[ls]) bw.put(len - T.len_base[ls], T.len_ebits[ls]);
        const int ds = T.dist_sym(dist);
        bw.put(dcode[ds], dlen[ds]);
        if (T.dist_ebits[ds]) bw.put(dist - T.dist_base[ds], T.dist_ebits[ds]);
    }
    bw.put(lcode[256], llen[256]);
    bw.flush_byte();
    const size_t out = (size_t)(bw.p - dst);
    if (bw.overflow || out >= n + 5) return stored_block(src, n, dst);
    return out;
}

}  // namespace fastdeflate
