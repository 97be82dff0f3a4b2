// samio.hpp -- record I/O for the C++ driver without htslib: SAM text and
// BGZF/BAM (zlib only), both directions, behind a line-oriented interface.
//
// The reference reads any of SAM/BAM through dhtslib's SAMReader (anno.d:22) and writes SAM, uBAM
// or BAM through getWriter (util.d:65-76: 0 = SAM, 1 = uncompressed BAM, 2 = BAM).  Here
//   LineSource  yields SAM text lines (header lines, then records) from SAM text or from BAM,
//   LineSink    takes SAM text lines and writes SAM text, BAM or level-0 BAM,
// so the commands in fade_cli.cpp stay format-agnostic.  Wire formats follow the SAM/BAM
// specification (SAMv1 sections 1.4, 4.1, 4.2): BGZF blocks are gzip members with a `BC` extra
// subfield; a BAM record is block_size + 32 fixed bytes + name + cigar + 4-bit bases + quals + aux.
// Text <-> binary conversions follow htslib's conventions: integer aux values are stored in the
// smallest type that holds them (non-negative ones unsigned), printed as `i`; floats print with %g;
// RNEXT prints "=" when it equals RNAME; a missing quality string is 0xff bytes.
#pragma once
#include <zlib.h>
#include "fastdeflate.hpp"
#include "fastinflate.hpp"
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace samio {

inline void put_u16(std::string &o, uint32_t v) { o.push_back((char)(v & 0xff)); o.push_back((char)((v >> 8) & 0xff)); }
inline void put_u32(std::string &o, uint32_t v) { put_u16(o, v & 0xffff); put_u16(o, v >> 16); }
inline uint32_t get_u16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
inline uint32_t get_u32(const uint8_t *p) { return get_u16(p) | (get_u16(p + 2) << 16); }
inline int32_t get_i32(const uint8_t *p) { return (int32_t)get_u32(p); }

// ---- BGZF -----------------------------------------------------------------------------------
// Payload of one BGZF block: cdata[0, clen) is the raw DEFLATE stream and cdata must be readable up to clen + 16 (the
// block's 8-byte trailer follows it in the file; callers allocate 8 bytes more).  The built-in decoder runs first; its
// result is accepted only if size and CRC-32 match, otherwise zlib decides (a damaged block fails there as well).
inline bool inflate_block(const uint8_t *cdata, size_t clen, uint8_t *out, uint32_t isize, uint32_t crc)
{
    if (isize == 0) return true;
    if (fastinflate::inflate(cdata, clen, out, isize) && (uint32_t)crc32(crc32(0, nullptr, 0), out, isize) == crc) return true;
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<uint8_t *>(cdata); zs.avail_in = (uInt)clen;
    zs.next_out = out; zs.avail_out = isize;
    const int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    return rc == Z_STREAM_END && zs.avail_out == 0 && (uint32_t)crc32(crc32(0, nullptr, 0), out, isize) == crc;
}

class BgzfReader {
public:
    explicit BgzfReader(FILE *f, const std::string &prefetched) : f_(f), raw_(prefetched) {}
    // exactly n bytes, false at a clean EOF before the first byte; throws nothing, sets bad() on damage
    bool read(void *dst, size_t n)
    {
        uint8_t *d = static_cast<uint8_t *>(dst);
        while (n) {
            if (pos_ == blk_.size() && !next_block()) return false;
            const size_t k = std::min(n, blk_.size() - pos_);
            memcpy(d, blk_.data() + pos_, k);
            d += k; pos_ += k; n -= k;
        }
        return true;
    }
    bool bad() const { return bad_; }

private:
    bool raw_read(uint8_t *d, size_t n)
    {
        size_t got = 0;
        if (raw_pos_ < raw_.size()) {
            got = std::min(n, raw_.size() - raw_pos_);
            memcpy(d, raw_.data() + raw_pos_, got);
            raw_pos_ += got;
        }
        if (got < n) got += fread(d + got, 1, n - got, f_);
        return got == n;
    }
    // Refills blk_ with the payload of the next group of blocks (up to kGroup): the raw blocks are read in order and
    // inflated side by side.  Empty blocks (the EOF marker) contribute nothing.
    bool next_block()
    {
        struct Raw { std::vector<uint8_t> c; uint32_t isize, crc; size_t off; };
        for (;;) {
            std::vector<Raw> grp;
            size_t total = 0;
            bool eof = false;
            while (grp.size() < kGroup) {
                uint8_t h[12];
                if (!raw_read(h, 12)) { eof = true; break; }
                if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { bad_ = true; return false; }
                const uint32_t xlen = get_u16(h + 10);
                std::vector<uint8_t> extra(xlen);
                if (!raw_read(extra.data(), xlen)) { bad_ = true; return false; }
                int64_t bsize = -1;
                for (size_t i = 0; i + 4 <= xlen;) {
                    const uint32_t sl = get_u16(&extra[i + 2]);
                    if (extra[i] == 'B' && extra[i + 1] == 'C' && sl == 2 && i + 6 <= xlen) bsize = get_u16(&extra[i + 4]);
                    i += 4 + sl;
                }
                if (bsize < 0) { bad_ = true; return false; }
                const int64_t clen = bsize - xlen - 19;
                if (clen < 0 || clen > 0x10000) { bad_ = true; return false; }
                Raw r;
                r.c.resize((size_t)clen + 16);     // trailer + the read slack of the built-in decoder
                if (!raw_read(r.c.data(), (size_t)clen + 8)) { bad_ = true; return false; }
                r.crc = get_u32(&r.c[(size_t)clen]);
                r.isize = get_u32(&r.c[(size_t)clen + 4]);
                if (r.isize > 0x10000) { bad_ = true; return false; }   // SAMv1 4.1: a block inflates to at most 64 KiB
                r.off = total;
                total += r.isize;
                grp.push_back(std::move(r));
                if (first_) break;    // the first block alone: a caller sniffing the magic should not wait for a group
            }
            first_ = false;
            if (grp.empty()) return false;
            blk_.resize(total);
            pos_ = 0;
            int bad = 0;
            const long ng = (long)grp.size();
#pragma omp parallel for schedule(dynamic, 4) reduction(| : bad) if (ng > 8)
            for (long k = 0; k < ng; ++k) {
                const Raw &r = grp[(size_t)k];
                if (!inflate_block(r.c.data(), r.c.size() - 16, blk_.data() + r.off, r.isize, r.crc)) bad = 1;
            }
            if (bad) { bad_ = true; return false; }
            if (total) return true;
            if (eof) return false;
        }
    }
    static constexpr size_t kGroup = 256;
    bool first_ = true;
    FILE *f_;
    std::string raw_;
    size_t raw_pos_ = 0;
    std::vector<uint8_t> blk_;
    size_t pos_ = 0;
    bool bad_ = false;
};

// Compression levels of BAM output: kFastLevel = the encoder of fastdeflate.hpp (the default of -b: the size of zlib
// level 6 on BAM records at six times its speed), 0 = stored, 1..9 = zlib.
constexpr int kFastLevel = -1;

// raw DEFLATE stream of src[0, n) (n <= 0xff00) into dst (capacity >= n + 1024); returns its size
inline size_t deflate_block(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, int level)
{
    if (level < 0) return fastdeflate::compress(src, n, dst, cap);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = const_cast<uint8_t *>(src); zs.avail_in = (uInt)n;
    zs.next_out = dst; zs.avail_out = (uInt)cap;
    deflate(&zs, Z_FINISH);
    const size_t clen = zs.total_out;
    deflateEnd(&zs);
    return clen;
}

class BgzfWriter {
public:
    BgzfWriter(FILE *f, int level) : f_(f), level_(level) { buf_.reserve(kBlock); }
    void write(const void *src, size_t n)
    {
        const uint8_t *s = static_cast<const uint8_t *>(src);
        while (n) {
            const size_t k = std::min(n, kBlock - buf_.size());
            buf_.insert(buf_.end(), s, s + k);
            s += k; n -= k;
            if (buf_.size() == kBlock) flush();
        }
    }
    void flush()   // closes the current block; blocks are deflated side by side, kPending at a time, and written in order
    {
        if (buf_.empty()) return;
        pend_.emplace_back();
        pend_.back().swap(buf_);
        buf_.reserve(kBlock);
        if (pend_.size() >= kPending) drain();
    }
    void finish()   // remaining data + the 28-byte EOF marker block
    {
        flush();
        drain();
        static const uint8_t eof[28] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        fwrite(eof, 1, sizeof(eof), f_);   // SAMv1 4.1.2
        fflush(f_);
    }

private:
    static constexpr size_t kBlock = 0xff00;
    static constexpr size_t kPending = 256;
    void drain()
    {
        const long nb = (long)pend_.size();
        if (!nb) return;
        std::vector<std::string> out((size_t)nb);
#pragma omp parallel for schedule(dynamic, 4) if (nb > 8)
        for (long k = 0; k < nb; ++k) {
            const std::vector<uint8_t> &in = pend_[(size_t)k];
            std::string &o = out[(size_t)k];
            o.resize(0x10000 + 64);
            const size_t clen = deflate_block(in.data(), in.size(), reinterpret_cast<uint8_t *>(&o[18]), o.size() - 18 - 8, level_);
            static const uint8_t head[16] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0 };
            memcpy(&o[0], head, 16);
            const uint32_t bsize = (uint32_t)(clen + 18 + 8 - 1);
            o[16] = (char)(bsize & 0xff); o[17] = (char)(bsize >> 8);
            const uint32_t crc = (uint32_t)crc32(crc32(0, nullptr, 0), in.data(), (uInt)in.size());
            for (int i = 0; i < 4; ++i) { o[18 + clen + (size_t)i] = (char)(crc >> (8 * i)); o[22 + clen + (size_t)i] = (char)((uint32_t)in.size() >> (8 * i)); }
            o.resize(18 + clen + 8);
        }
        for (const auto &o : out) fwrite(o.data(), 1, o.size(), f_);
        pend_.clear();
    }
    FILE *f_;
    int level_;
    std::vector<uint8_t> buf_;
    std::vector<std::vector<uint8_t>> pend_;
};

// ---- header ---------------------------------------------------------------------------------
struct Header {
    std::vector<std::string> lines;      // SAM header lines, no newline
    std::vector<std::string> names;      // @SQ SN in order (= refID)
    std::vector<int64_t> lens;
    std::map<std::string, int> tid_of;
    void add_line(const std::string &l)
    {
        lines.push_back(l);
        if (l.compare(0, 3, "@SQ") != 0) return;
        std::string sn;
        int64_t ln = 0;
        size_t a = 0;
        while (a <= l.size()) {
            size_t b = l.find('\t', a);
            if (b == std::string::npos) b = l.size();
            if (l.compare(a, 3, "SN:") == 0) sn = l.substr(a + 3, b - a - 3);
            if (l.compare(a, 3, "LN:") == 0) ln = atoll(l.substr(a + 3, b - a - 3).c_str());
            a = b + 1;
        }
        tid_of[sn] = (int)names.size();
        names.push_back(sn);
        lens.push_back(ln);
    }
};

// ---- BAM record <-> SAM line ----------------------------------------------------------------
inline int reg2bin(int64_t beg, int64_t end)   // SAMv1 5.3
{
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

inline void fmt_g(std::string &o, double v)
{
    char b[64];
    snprintf(b, sizeof(b), "%g", v);
    o += b;
}

// the bytes after block_size -> one SAM text line (no newline); false on a damaged record
inline bool bam_to_sam(const uint8_t *p, size_t n, const Header &h, std::string &o)
{
    if (n < 32) return false;
    const int32_t tid = get_i32(p), pos = get_i32(p + 4);
    const uint32_t l_name = p[8], mapq = p[9], n_cig = get_u16(p + 12), flag = get_u16(p + 14);
    const int32_t l_seq = get_i32(p + 16), mtid = get_i32(p + 20), mpos = get_i32(p + 24), tlen = get_i32(p + 28);
    size_t q = 32;
    if (l_seq < 0 || l_name == 0 || q + l_name + 4ull * n_cig + (size_t)(l_seq + 1) / 2 + (size_t)l_seq > n) return false;
    o.clear();
    o.append(reinterpret_cast<const char *>(p + q), strnlen(reinterpret_cast<const char *>(p + q), l_name));
    q += l_name;
    auto name_of = [&](int32_t t) -> std::string { return (t >= 0 && (size_t)t < h.names.size()) ? h.names[(size_t)t] : "*"; };
    o += '\t'; o += std::to_string(flag);
    o += '\t'; o += name_of(tid);
    o += '\t'; o += std::to_string((int64_t)pos + 1);
    o += '\t'; o += std::to_string(mapq);
    o += '\t';
    if (n_cig == 0) o += '*';
    for (uint32_t k = 0; k < n_cig; ++k) {
        const uint32_t c = get_u32(p + q + 4 * k);
        o += std::to_string(c >> 4);
        o += "MIDNSHP=XB??????"[c & 15];
    }
    q += 4ull * n_cig;
    o += '\t';
    if (mtid < 0) o += '*'; else if (mtid == tid) o += '='; else o += name_of(mtid);
    o += '\t'; o += std::to_string((int64_t)mpos + 1);
    o += '\t'; o += std::to_string(tlen);
    o += '\t';
    if (l_seq == 0) o += '*';
    for (int32_t i = 0; i < l_seq; ++i) o += "=ACMGRSVTWYHKDBN"[(p[q + (size_t)(i >> 1)] >> ((~i & 1) << 2)) & 15];
    q += (size_t)(l_seq + 1) / 2;
    o += '\t';
    if (l_seq == 0 || p[q] == 0xff) o += '*';
    else for (int32_t i = 0; i < l_seq; ++i) o += (char)(p[q + (size_t)i] + 33);
    q += (size_t)l_seq;
    while (q + 3 <= n) {   // aux
        o += '\t'; o += (char)p[q]; o += (char)p[q + 1]; o += ':';
        const char ty = (char)p[q + 2];
        q += 3;
        auto need = [&](size_t k) { return q + k <= n; };
        switch (ty) {
        case 'A': if (!need(1)) return false; o += "A:"; o += (char)p[q]; q += 1; break;
        case 'c': if (!need(1)) return false; o += "i:" + std::to_string((int)(int8_t)p[q]); q += 1; break;
        case 'C': if (!need(1)) return false; o += "i:" + std::to_string((unsigned)p[q]); q += 1; break;
        case 's': if (!need(2)) return false; o += "i:" + std::to_string((int)(int16_t)get_u16(p + q)); q += 2; break;
        case 'S': if (!need(2)) return false; o += "i:" + std::to_string(get_u16(p + q)); q += 2; break;
        case 'i': if (!need(4)) return false; o += "i:" + std::to_string(get_i32(p + q)); q += 4; break;
        case 'I': if (!need(4)) return false; o += "i:" + std::to_string(get_u32(p + q)); q += 4; break;
        case 'f': { if (!need(4)) return false; float f; memcpy(&f, p + q, 4); o += "f:"; fmt_g(o, f); q += 4; break; }
        case 'd': { if (!need(8)) return false; double d; memcpy(&d, p + q, 8); o += "d:"; fmt_g(o, d); q += 8; break; }
        case 'Z': case 'H': {
            const size_t l = strnlen(reinterpret_cast<const char *>(p + q), n - q);
            if (q + l >= n) return false;
            o += ty; o += ':';
            o.append(reinterpret_cast<const char *>(p + q), l);
            q += l + 1;
            break;
        }
        case 'B': {
            if (!need(5)) return false;
            const char st = (char)p[q];
            const uint32_t cnt = get_u32(p + q + 1);
            q += 5;
            const size_t w = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : (st == 'i' || st == 'I' || st == 'f') ? 4 : 0;
            if (!w || !need((size_t)cnt * w)) return false;
            o += "B:"; o += st;
            for (uint32_t k = 0; k < cnt; ++k, q += w) {
                o += ',';
                switch (st) {
                case 'c': o += std::to_string((int)(int8_t)p[q]); break;
                case 'C': o += std::to_string((unsigned)p[q]); break;
                case 's': o += std::to_string((int)(int16_t)get_u16(p + q)); break;
                case 'S': o += std::to_string(get_u16(p + q)); break;
                case 'i': o += std::to_string(get_i32(p + q)); break;
                case 'I': o += std::to_string(get_u32(p + q)); break;
                default: { float f; memcpy(&f, p + q, 4); fmt_g(o, f); }
                }
            }
            break;
        }
        default: return false;
        }
    }
    return q == n;
}

// one SAM text line -> the bytes after block_size; false on a malformed line
inline bool sam_to_bam(const std::string &line, const Header &h, std::string &o)
{
    std::vector<std::pair<size_t, size_t>> f;   // (offset, length) of the tab-separated fields
    for (size_t a = 0;;) {
        const size_t b = line.find('\t', a);
        if (b == std::string::npos) { f.push_back({ a, line.size() - a }); break; }
        f.push_back({ a, b - a });
        a = b + 1;
    }
    if (f.size() < 11) return false;
    auto fs = [&](size_t k) { return line.substr(f[k].first, f[k].second); };
    auto tid_of = [&](const std::string &s, int32_t same) -> int32_t {
        if (s == "*") return -1;
        if (s == "=") return same;
        auto it = h.tid_of.find(s);
        return it == h.tid_of.end() ? -1 : it->second;
    };
    const std::string name = fs(0), cig = fs(5), seq = fs(9), qual = fs(10);
    if (name.empty() || name.size() > 254) return false;
    const int32_t tid = tid_of(fs(2), -1);
    const int64_t pos = atoll(fs(3).c_str()) - 1;
    std::vector<uint32_t> ops;
    int64_t ref_len = 0;
    if (cig != "*") {
        uint64_t num = 0;
        bool have = false;
        for (char c : cig) {
            if (c >= '0' && c <= '9') { num = num * 10 + (uint64_t)(c - '0'); have = true; continue; }
            const char *pp = strchr("MIDNSHP=XB", c);
            if (!pp || !have || num >= (1u << 28)) return false;
            const uint32_t op = (uint32_t)(pp - "MIDNSHP=XB");
            ops.push_back((uint32_t)(num << 4) | op);
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += (int64_t)num;
            num = 0; have = false;
        }
        if (have || ops.size() > 65535) return false;
    }
    const int32_t l_seq = seq == "*" ? 0 : (int32_t)seq.size();
    if (qual != "*" && (int32_t)qual.size() != l_seq) return false;
    o.clear();
    put_u32(o, (uint32_t)tid);
    put_u32(o, (uint32_t)(int32_t)pos);
    o.push_back((char)(name.size() + 1));
    o.push_back((char)(atoi(fs(4).c_str()) & 0xff));
    put_u16(o, (uint32_t)reg2bin(pos, pos + (ref_len > 0 ? ref_len : 1)));
    put_u16(o, (uint32_t)ops.size());
    put_u16(o, (uint32_t)atoi(fs(1).c_str()));
    put_u32(o, (uint32_t)l_seq);
    put_u32(o, (uint32_t)tid_of(fs(6), tid));
    put_u32(o, (uint32_t)(int32_t)(atoll(fs(7).c_str()) - 1));
    put_u32(o, (uint32_t)(int32_t)atoll(fs(8).c_str()));
    o += name; o.push_back('\0');
    for (uint32_t c : ops) put_u32(o, c);
    {
        static const char tbl[] = "=ACMGRSVTWYHKDBN";
        const size_t base = o.size();
        o.append((size_t)(l_seq + 1) / 2, '\0');
        for (int32_t i = 0; i < l_seq; ++i) {
            const char c = (char)toupper((unsigned char)seq[(size_t)i]);
            const char *pp = c ? strchr(tbl, c) : nullptr;
            o[base + (size_t)(i >> 1)] = (char)((uint8_t)o[base + (size_t)(i >> 1)] | (uint8_t)((pp ? (int)(pp - tbl) : 15) << ((~i & 1) << 2)));
        }
    }
    if (qual == "*") o.append((size_t)l_seq, (char)0xff);
    else for (char c : qual) o.push_back((char)(c - 33));
    for (size_t k = 11; k < f.size(); ++k) {   // aux: TG:T:value
        const std::string a = fs(k);
        if (a.size() < 5 || a[2] != ':' || a[4] != ':') return false;
        o += a[0]; o += a[1];
        const char ty = a[3];
        const char *val = a.c_str() + 5;
        auto put_int = [&](long long v, bool allow_tag) {   // htslib: smallest type; non-negative -> unsigned
            char t;
            if (v < 0) t = v >= -128 ? 'c' : v >= -32768 ? 's' : 'i';
            else t = v <= 255 ? 'C' : v <= 65535 ? 'S' : 'I';
            if (allow_tag) o += t;
            if (t == 'c' || t == 'C') o.push_back((char)(v & 0xff));
            else if (t == 's' || t == 'S') put_u16(o, (uint32_t)(v & 0xffff));
            else put_u32(o, (uint32_t)(v & 0xffffffffll));
        };
        switch (ty) {
        case 'A': o += 'A'; o += val[0]; break;
        case 'i': put_int(atoll(val), true); break;
        case 'f': { o += 'f'; const float v = strtof(val, nullptr); o.append(reinterpret_cast<const char *>(&v), 4); break; }
        case 'Z': case 'H': o += ty; o += val; o.push_back('\0'); break;
        case 'B': {
            const char st = val[0];
            if (!strchr("cCsSiIf", st) || !st) return false;
            o += 'B'; o += st;
            std::vector<std::string> items;
            for (const char *s = val + 1; *s == ',';) {
                const char *e = strchr(s + 1, ',');
                items.emplace_back(s + 1, e ? (size_t)(e - s - 1) : strlen(s + 1));
                if (!e) break;
                s = e;
            }
            put_u32(o, (uint32_t)items.size());
            for (const auto &it : items) {
                if (st == 'f') { const float v = strtof(it.c_str(), nullptr); o.append(reinterpret_cast<const char *>(&v), 4); continue; }
                const long long v = atoll(it.c_str());
                if (st == 'c' || st == 'C') o.push_back((char)(v & 0xff));
                else if (st == 's' || st == 'S') put_u16(o, (uint32_t)(v & 0xffff));
                else put_u32(o, (uint32_t)(v & 0xffffffffll));
            }
            break;
        }
        default: return false;
        }
    }
    return true;
}

// ---- line-oriented source / sink ------------------------------------------------------------
class LineSource {
public:
    // path "-" = stdin.  BAM is recognised by the gzip magic.
    bool open(const std::string &path)
    {
        FILE *f = path == "-" ? stdin : fopen(path.c_str(), "rb");
        if (!f) return false;
        uint8_t m[2];
        const size_t got = fread(m, 1, 2, f);
        return open(f, std::string(reinterpret_cast<char *>(m), got));
    }
    // an already opened file whose first bytes (`pre`, at least 2 unless the file is shorter) were read
    bool open(FILE *f, const std::string &pre)
    {
        f_ = f;
        const size_t got = pre.size();
        const uint8_t *m = reinterpret_cast<const uint8_t *>(pre.data());
        if (got >= 2 && m[0] == 0x1f && m[1] == 0x8b) {
            bam_ = true;
            bz_ = new BgzfReader(f_, pre);
            return read_bam_header();
        }
        text_pre_ = pre;
        return true;
    }
    ~LineSource()
    {
        delete bz_;
        if (f_ && f_ != stdin) fclose(f_);
    }
    bool is_bam() const { return bam_; }
    // BAM input only: records for which skip(record bytes after block_size, length) holds are not converted to text and
    // come out of getline() as empty lines (a caller that needs one record in ten saves nine conversions); set before open()
    void skip_records_if(bool (*skip)(const uint8_t *, size_t)) { skip_ = skip; }
    bool failed() const { return fail_ || (bz_ && bz_->bad()); }
    // next SAM text line (header lines first); false at the end
    bool getline(std::string &line)
    {
        if (!bam_) return text_line(line);
        if (hdr_pos_ < hdr_.lines.size()) { line = hdr_.lines[hdr_pos_++]; return true; }
        if (q_pos_ == q_.size() && !refill()) return false;
        line.swap(q_[q_pos_++]);
        return true;
    }

private:
    // BAM input: the records of the next stretch of the file are read in order and converted to SAM text side by side
    bool refill()
    {
        q_.clear();
        q_pos_ = 0;
        if (fail_) return false;
        raw_.clear();
        off_.clear();
        while (off_.size() < kBatch) {
            uint8_t b4[4];
            if (!bz_->read(b4, 4)) break;
            const uint32_t bs = get_u32(b4);
            if (bs > (1u << 29)) { fail_ = true; break; }
            const size_t at = raw_.size();
            raw_.resize(at + bs);
            if (!bz_->read(raw_.data() + at, bs)) { fail_ = true; break; }
            off_.push_back(at);
        }
        off_.push_back(raw_.size());
        const long n = (long)off_.size() - 1;
        if (n <= 0) return false;
        q_.resize((size_t)n);
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) if (n > 256)
        for (long k = 0; k < n; ++k) {
            const uint8_t *rec = raw_.data() + off_[(size_t)k];
            const size_t len = off_[(size_t)k + 1] - off_[(size_t)k];
            if (skip_ && skip_(rec, len)) { q_[(size_t)k].clear(); continue; }   // served as an empty line, which the callers pass over
            if (!bam_to_sam(rec, len, hdr_, q_[(size_t)k])) bad |= 1;
        }
        if (bad) {   // serve the records before the first damaged one, then fail
            fail_ = true;
            std::string tmp;
            size_t good = 0;
            while (good < (size_t)n && bam_to_sam(raw_.data() + off_[good], off_[good + 1] - off_[good], hdr_, tmp)) ++good;
            q_.resize(good);
        }
        return !q_.empty();
    }
    static constexpr size_t kBatch = 1 << 15;
    bool (*skip_)(const uint8_t *, size_t) = nullptr;
    std::vector<std::string> q_;
    size_t q_pos_ = 0;
    std::vector<uint8_t> raw_;
    std::vector<size_t> off_;

    bool text_line(std::string &line)
    {
        line.clear();
        for (;;) {
            if (tpos_ == tbuf_.size()) {
                if (!text_pre_.empty()) { tbuf_.assign(text_pre_.begin(), text_pre_.end()); text_pre_.clear(); }
                else {
                    tbuf_.resize(1 << 20);
                    tbuf_.resize(fread(tbuf_.data(), 1, tbuf_.size(), f_));
                }
                tpos_ = 0;
                if (tbuf_.empty()) return !line.empty();
            }
            const char *s = tbuf_.data() + tpos_;
            const char *e = static_cast<const char *>(memchr(s, '\n', tbuf_.size() - tpos_));
            if (e) { line.append(s, (size_t)(e - s)); tpos_ += (size_t)(e - s) + 1; return true; }
            line.append(s, tbuf_.size() - tpos_);
            tpos_ = tbuf_.size();
        }
    }
    bool read_bam_header()
    {
        uint8_t b[8];
        if (!bz_->read(b, 8) || memcmp(b, "BAM\1", 4) != 0) { fail_ = true; return false; }
        const uint32_t l_text = get_u32(b + 4);
        std::string text(l_text, '\0');
        if (l_text && !bz_->read(&text[0], l_text)) { fail_ = true; return false; }
        text.resize(strnlen(text.c_str(), text.size()));
        bool has_sq = false;
        for (size_t a = 0; a < text.size();) {
            size_t e = text.find('\n', a);
            if (e == std::string::npos) e = text.size();
            if (e > a) { hdr_.add_line(text.substr(a, e - a)); has_sq |= text.compare(a, 3, "@SQ") == 0; }
            a = e + 1;
        }
        if (!bz_->read(b, 4)) { fail_ = true; return false; }
        const uint32_t n_ref = get_u32(b);
        for (uint32_t r = 0; r < n_ref; ++r) {
            if (!bz_->read(b, 4)) { fail_ = true; return false; }
            const uint32_t l = get_u32(b);
            std::string nm(l, '\0');
            if (!bz_->read(&nm[0], l) || !bz_->read(b, 4)) { fail_ = true; return false; }
            nm.resize(strnlen(nm.c_str(), nm.size()));
            // the binary reference list is authoritative when the text has no @SQ lines
            if (!has_sq) hdr_.add_line("@SQ\tSN:" + nm + "\tLN:" + std::to_string(get_u32(b)));
        }
        return true;
    }
    FILE *f_ = nullptr;
    bool bam_ = false, fail_ = false;
    BgzfReader *bz_ = nullptr;
    Header hdr_;
    size_t hdr_pos_ = 0;
    std::vector<uint8_t> rec_;
    std::string text_pre_;
    std::vector<char> tbuf_;
    size_t tpos_ = 0;
};

class LineSink {
public:
    enum Mode { SAM = 0, UBAM = 1, BAM = 2 };   // util.d:65-76
    LineSink(FILE *f, Mode m) : f_(f), mode_(m) {}
    ~LineSink() { close(); }
    bool put(const std::string &line)
    {
        if (mode_ == SAM) { fwrite(line.data(), 1, line.size(), f_); fputc('\n', f_); return true; }
        if (!line.empty() && line[0] == '@' && !started_) { hdr_.add_line(line); return true; }
        if (!started_) start();
        if (failed_) return false;
        // lines are converted to binary records side by side, kBatch at a time; a line that cannot be encoded makes
        // this and every later put() fail (the caller stops: the output would lack a record)
        pend_.push_back(line);
        if (pend_.size() >= kBatch) flush_lines();
        return !failed_;
    }
    bool close()
    {
        if (closed_) return !failed_;
        closed_ = true;
        if (mode_ == SAM) { fflush(f_); return true; }
        if (!started_) start();
        flush_lines();
        bz_->finish();
        delete bz_;
        bz_ = nullptr;
        return !failed_;
    }
    const std::string &bad_line() const { return bad_line_; }

private:
    void flush_lines()
    {
        const long n = (long)pend_.size();
        if (!n) return;
        recs_.resize((size_t)n);
        long first_bad = n;
#pragma omp parallel for schedule(static) reduction(min : first_bad) if (n > 256)
        for (long k = 0; k < n; ++k)
            if (!sam_to_bam(pend_[(size_t)k], hdr_, recs_[(size_t)k])) first_bad = std::min(first_bad, k);
        for (long k = 0; k < first_bad; ++k) {
            std::string bs;
            put_u32(bs, (uint32_t)recs_[(size_t)k].size());
            bz_->write(bs.data(), 4);
            bz_->write(recs_[(size_t)k].data(), recs_[(size_t)k].size());
        }
        if (first_bad < n) { failed_ = true; bad_line_ = pend_[(size_t)first_bad]; }
        pend_.clear();
    }
    static constexpr size_t kBatch = 1 << 15;
    std::vector<std::string> pend_, recs_;
    bool failed_ = false;
    std::string bad_line_;
    void start()
    {
        started_ = true;
        bz_ = new BgzfWriter(f_, mode_ == UBAM ? 0 : kFastLevel);
        std::string text;
        for (const auto &l : hdr_.lines) { text += l; text += '\n'; }
        std::string o("BAM\1", 4);
        put_u32(o, (uint32_t)text.size());
        o += text;
        put_u32(o, (uint32_t)hdr_.names.size());
        for (size_t r = 0; r < hdr_.names.size(); ++r) {
            put_u32(o, (uint32_t)hdr_.names[r].size() + 1);
            o += hdr_.names[r]; o.push_back('\0');
            put_u32(o, (uint32_t)hdr_.lens[r]);
        }
        bz_->write(o.data(), o.size());
        bz_->flush();   // htslib starts the records in a fresh block
    }
    FILE *f_;
    Mode mode_;
    Header hdr_;
    bool started_ = false, closed_ = false;
    BgzfWriter *bz_ = nullptr;
    std::string rec_;
};

}  // namespace samio
