// fastinflate.hpp -- a table-driven DEFLATE decoder (RFC 1951) for BGZF blocks whose inflated size is known.
//
// Why: after the built-in encoder (fastdeflate.hpp) zlib's inflate was half of the record loop of `fade-b200 annotate`.
// This decoder keeps a 64-bit bit buffer that is refilled eight bytes at a time, decodes literal / length symbols through
// a 10-bit primary table (longer codes through sub-tables) and copies matches word-wise: about twice zlib's speed on BAM.
// Safety net: a BGZF block carries the CRC-32 and the size of its payload; the callers verify both and fall back to
// zlib's inflate when this decoder reports an error or the check fails, so a decoder bug can cost time, not data.
//
// inflate(): raw DEFLATE stream in[0, n_in) -> out[0, n_out); `in` must be READABLE up to in + n_in + 16 (the bit buffer is
// refilled eight bytes at a time; what lies past n_in is never used by a well-formed stream); returns true iff the stream
// is well formed, ends with a final block inside in[0, n_in) and produces exactly n_out bytes.  Nothing is written outside
// out[0, n_out).
#pragma once
#include <cstdint>
#include <cstring>

namespace fastinflate {

namespace detail {

constexpr int LBITS = 10, DBITS = 8;
constexpr int LSIZE = (1 << LBITS) + 2048, DSIZE = (1 << DBITS) + 1024;
// table entry: value << 16 | op << 8 | bits     (bits = code bits to drop; for a link: bits of the primary index)
constexpr uint32_t OP_LIT = 0x00, OP_LEN = 0x10 /* + extra bits */, OP_EOB = 0x20, OP_LINK = 0x40 /* + sub-table bits */, OP_BAD = 0x80;

struct Tables {
    uint32_t lt[LSIZE];
    uint32_t dt[DSIZE];
};

inline uint32_t rev(uint32_t c, int len)
{
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) { r = (r << 1) | (c & 1); c >>= 1; }
    return r;
}

// Decode table of a canonical prefix code.  lens[0, n): code lengths (0 = unused); entry(sym) gives value << 16 | op << 8.
// Returns false for an over-subscribed code or when the sub-tables do not fit.
template <class EntryOf>
inline bool build(const uint8_t *lens, int n, int root, uint32_t *tab, int tab_size, EntryOf entry)
{
    int count[16] = { 0 };
    for (int s = 0; s < n; ++s) ++count[lens[s]];
    count[0] = 0;
    uint32_t code = 0, next[16];
    int64_t left = 1;
    for (int l = 1; l <= 15; ++l) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return false;                       // over-subscribed
        code = (code + (uint32_t)count[l - 1]) << 1;
        next[l] = code;
    }
    const int rsize = 1 << root;
    for (int i = 0; i < rsize; ++i) tab[i] = (OP_BAD << 8);
    // long codes: the widest code below every primary index decides the size of its sub-table
    uint8_t sub_bits[1 << LBITS];
    memset(sub_bits, 0, (size_t)rsize);
    uint16_t rc[288];
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        rc[s] = (uint16_t)rev(next[l]++, l);
        if (l > root) {
            const int p = rc[s] & (rsize - 1);
            if (l - root > sub_bits[p]) sub_bits[p] = (uint8_t)(l - root);
        }
    }
    int used = rsize;
    for (int p = 0; p < rsize; ++p)
        if (sub_bits[p]) {
            const int sz = 1 << sub_bits[p];
            if (used + sz > tab_size) return false;
            tab[p] = ((uint32_t)used << 16) | ((OP_LINK + sub_bits[p]) << 8) | (uint32_t)root;
            for (int i = 0; i < sz; ++i) tab[used + i] = (OP_BAD << 8);
            used += sz;
        }
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t e = entry(s);
        if (l <= root) {
            for (int i = rc[s]; i < rsize; i += 1 << l) tab[i] = e | (uint32_t)l;
        } else {
            const int p = rc[s] & (rsize - 1);
            const uint32_t link = tab[p];
            const int base = (int)(link >> 16), sb = (int)((link >> 8) & 0xff) - (int)OP_LINK;
            for (int i = rc[s] >> root; i < (1 << sb); i += 1 << (l - root)) tab[base + i] = e | (uint32_t)(l - root);
        }
    }
    return true;
}

inline uint32_t litlen_entry(int s)
{
    static const uint16_t base[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
    static const uint8_t extra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
    if (s < 256) return ((uint32_t)s << 16) | (OP_LIT << 8);
    if (s == 256) return OP_EOB << 8;
    if (s > 285) return OP_BAD << 8;
    return ((uint32_t)base[s - 257] << 16) | ((OP_LEN + extra[s - 257]) << 8);
}

inline uint32_t dist_entry(int s)
{
    static const uint16_t base[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                                       4097, 6145, 8193, 12289, 16385, 24577 };
    static const uint8_t extra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
    if (s > 29) return OP_BAD << 8;
    return ((uint32_t)base[s] << 16) | ((OP_LEN + extra[s]) << 8);
}

inline const Tables &fixed_tables()
{
    static const Tables t = [] {
        Tables x;
        uint8_t ll[288], dl[30];
        for (int s = 0; s < 144; ++s) ll[s] = 8;
        for (int s = 144; s < 256; ++s) ll[s] = 9;
        for (int s = 256; s < 280; ++s) ll[s] = 7;
        for (int s = 280; s < 288; ++s) ll[s] = 8;
        for (int s = 0; s < 30; ++s) dl[s] = 5;
        build(ll, 288, LBITS, x.lt, LSIZE, litlen_entry);
        build(dl, 30, DBITS, x.dt, DSIZE, dist_entry);
        return x;
    }();
    return t;
}

inline uint64_t load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

}  // namespace detail

inline bool inflate(const uint8_t *in, size_t n_in, uint8_t *out, size_t n_out)
{
    using namespace detail;
    const uint8_t *ip = in, *const in_end = in + n_in, *const in_lim = in + n_in + 8;
    uint8_t *op = out, *const out_end = out + n_out;
    uint64_t bb = 0;     // bit buffer, LSB first
    int bc = 0;          // valid bits in bb
    Tables dyn;
    // at least 56 bits in the buffer; near the end of the input some of them come from past in_end, which a valid stream
    // never consumes (checked at the end: the bits consumed must lie inside in[0, n_in))
#define FI_REFILL()                                       \
    do {                                                  \
        if (ip > in_lim) return false;                    \
        bb |= load64(ip) << bc;                           \
        ip += (63 - bc) >> 3;                             \
        bc |= 56;                                         \
    } while (0)
#define FI_DROP(n) do { bb >>= (n); bc -= (n); } while (0)
    for (;;) {
        FI_REFILL();
        const uint32_t hdr = (uint32_t)bb & 7;
        FI_DROP(3);
        const bool final_block = hdr & 1;
        const uint32_t type = hdr >> 1;
        const Tables *T = nullptr;
        if (type == 0) {
            // stored: back to a byte boundary of the input
            FI_DROP(bc & 7);
            ip -= bc >> 3;
            bb = 0; bc = 0;
            if (ip + 4 > in_end) return false;
            const uint32_t len = ip[0] | (ip[1] << 8), nlen = ip[2] | (ip[3] << 8);
            ip += 4;
            if ((len ^ nlen) != 0xffffu || ip + len > in_end || op + len > out_end) return false;
            memcpy(op, ip, len);
            ip += len; op += len;
            if (final_block) break;
            continue;
        } else if (type == 1) {
            T = &fixed_tables();
        } else if (type == 2) {
            const int hlit = (int)(bb & 31) + 257, hdist = (int)((bb >> 5) & 31) + 1, hclen = (int)((bb >> 10) & 15) + 4;
            FI_DROP(14);
            if (hlit > 286 || hdist > 30) return false;
            static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
            uint8_t cl[19] = { 0 };
            for (int k = 0; k < hclen; ++k) {
                if (bc < 3) FI_REFILL();
                cl[order[k]] = (uint8_t)(bb & 7);
                FI_DROP(3);
            }
            uint32_t ct[1 << 7];
            if (!build(cl, 19, 7, ct, 1 << 7, [](int s) { return (uint32_t)s << 16; })) return false;
            uint8_t lens[286 + 30 + 138];
            int n = 0;
            const int total = hlit + hdist;
            while (n < total) {
                FI_REFILL();
                const uint32_t e = ct[bb & 127];
                if ((e >> 8) & OP_BAD) return false;
                FI_DROP((int)(e & 0xff));
                const int sym = (int)(e >> 16);
                if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
                int rep;
                uint8_t v = 0;
                if (sym == 16) { if (n == 0) return false; v = lens[n - 1]; rep = 3 + (int)(bb & 3); FI_DROP(2); }
                else if (sym == 17) { rep = 3 + (int)(bb & 7); FI_DROP(3); }
                else { rep = 11 + (int)(bb & 127); FI_DROP(7); }
                if (n + rep > total) return false;
                memset(lens + n, v, (size_t)rep);
                n += rep;
            }
            if (lens[256] == 0) return false;            // no end-of-block code
            if (!build(lens, hlit, LBITS, dyn.lt, LSIZE, litlen_entry)) return false;
            if (!build(lens + hlit, hdist, DBITS, dyn.dt, DSIZE, dist_entry)) return false;
            T = &dyn;
        } else return false;

        const uint32_t *lt = T->lt, *dt = T->dt;
        // ---- fast loop: far enough from both ends that neither the refill nor a 258-byte match needs a bounds check ----
        if (n_out >= 280 && n_in >= 16) {
            const uint8_t *const in_fast = in_end - 8;
            uint8_t *const out_fast = out_end - 274;
            bool block_done = false;
            while (op < out_fast && ip < in_fast) {
#define FI_REFILL_FAST() do { bb |= load64(ip) << bc; ip += (63 - bc) >> 3; bc |= 56; } while (0)
                FI_REFILL_FAST();
                uint32_t e = lt[bb & ((1u << LBITS) - 1)];
                if (!((e >> 8) & 0xff)) {          // up to three literals per refill (45 of the 56 bits)
                    FI_DROP((int)(e & 0xff));
                    *op++ = (uint8_t)(e >> 16);
                    e = lt[bb & ((1u << LBITS) - 1)];
                    if (!((e >> 8) & 0xff)) {
                        FI_DROP((int)(e & 0xff));
                        *op++ = (uint8_t)(e >> 16);
                        e = lt[bb & ((1u << LBITS) - 1)];
                        if (!((e >> 8) & 0xff)) {
                            FI_DROP((int)(e & 0xff));
                            *op++ = (uint8_t)(e >> 16);
                            continue;
                        }
                    }
                    FI_REFILL_FAST();
                }
                uint32_t o = (e >> 8) & 0xff;
                if (o & OP_LINK) {
                    FI_DROP(LBITS);
                    e = lt[(e >> 16) + (bb & ((1u << (o - OP_LINK)) - 1))];
                    o = (e >> 8) & 0xff;
                }
                if (o & OP_BAD) return false;
                FI_DROP((int)(e & 0xff));
                if (o == OP_LIT) { *op++ = (uint8_t)(e >> 16); continue; }
                if (o == OP_EOB) { block_done = true; break; }
                const int le = (int)(o - OP_LEN);
                const uint32_t len = (e >> 16) + ((uint32_t)bb & ((1u << le) - 1));
                FI_DROP(le);
                uint32_t d = dt[bb & ((1u << DBITS) - 1)];
                uint32_t od = (d >> 8) & 0xff;
                if (od & OP_LINK) {
                    FI_DROP(DBITS);
                    d = dt[(d >> 16) + (bb & ((1u << (od - OP_LINK)) - 1))];
                    od = (d >> 8) & 0xff;
                }
                if ((od & OP_BAD) || od < OP_LEN) return false;
                FI_DROP((int)(d & 0xff));
                const int de = (int)(od - OP_LEN);
                const uint32_t dist = (d >> 16) + ((uint32_t)bb & ((1u << de) - 1));
                FI_DROP(de);
                if (dist > (size_t)(op - out)) return false;
                const uint8_t *src = op - dist;
                if (dist >= 8) {
                    uint8_t *dst = op;
                    const uint8_t *const stop = op + len;
                    do { memcpy(dst, src, 8); dst += 8; src += 8; } while (dst < stop);
                } else {
                    for (uint32_t k = 0; k < len; ++k) op[k] = src[k];
                }
                op += len;
#undef FI_REFILL_FAST
            }
            if (block_done) { if (final_block) break; continue; }
        }
        // ---- careful loop: the ends of the buffers ----
        for (;;) {
            FI_REFILL();
            uint32_t e = lt[bb & ((1u << LBITS) - 1)];
            // two literals per refill are common in BAM payloads
            if (!((e >> 8) & 0xff)) {
                if (op >= out_end) return false;
                FI_DROP((int)(e & 0xff));
                *op++ = (uint8_t)(e >> 16);
                e = lt[bb & ((1u << LBITS) - 1)];
                if (!((e >> 8) & 0xff)) {
                    if (op >= out_end) return false;
                    FI_DROP((int)(e & 0xff));
                    *op++ = (uint8_t)(e >> 16);
                    continue;
                }
                // (at most 30 bits are gone: 26 + 15 + 5 + 15 + 13 > 56 -- refill before a possible match)
                FI_REFILL();
            }
            uint32_t o = (e >> 8) & 0xff;
            if (o & OP_LINK) {
                FI_DROP(LBITS);
                e = lt[(e >> 16) + (bb & ((1u << (o - OP_LINK)) - 1))];
                o = (e >> 8) & 0xff;
            }
            if (o & OP_BAD) return false;
            FI_DROP((int)(e & 0xff));
            if (o == OP_LIT) {
                if (op >= out_end) return false;
                *op++ = (uint8_t)(e >> 16);
                continue;
            }
            if (o == OP_EOB) break;
            // length + distance
            const int le = (int)(o - OP_LEN);
            const uint32_t len = (e >> 16) + ((uint32_t)bb & ((1u << le) - 1));
            FI_DROP(le);
            uint32_t d = dt[bb & ((1u << DBITS) - 1)];
            uint32_t od = (d >> 8) & 0xff;
            if (od & OP_LINK) {
                FI_DROP(DBITS);
                d = dt[(d >> 16) + (bb & ((1u << (od - OP_LINK)) - 1))];
                od = (d >> 8) & 0xff;
            }
            if ((od & OP_BAD) || od < OP_LEN) return false;
            FI_DROP((int)(d & 0xff));
            const int de = (int)(od - OP_LEN);
            const uint32_t dist = (d >> 16) + ((uint32_t)bb & ((1u << de) - 1));
            FI_DROP(de);
            if (dist > (size_t)(op - out) || len > (size_t)(out_end - op)) return false;
            const uint8_t *src = op - dist;
            if (dist >= 8 && (size_t)(out_end - op) >= len + 8) {
                uint8_t *dst = op;                        // word-wise; the overshoot stays inside out[0, n_out)
                const uint8_t *const stop = op + len;
                do { memcpy(dst, src, 8); dst += 8; src += 8; } while (dst < stop);
            } else {
                for (uint32_t k = 0; k < len; ++k) op[k] = src[k];
            }
            op += len;
        }
        if (final_block) break;
    }
#undef FI_REFILL
#undef FI_DROP
    return op == out_end && (int64_t)(ip - in) * 8 - bc <= (int64_t)n_in * 8;
}

}  // namespace fastinflate
