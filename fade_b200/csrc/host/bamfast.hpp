// bamfast.hpp -- the record loop of `fade-b200 annotate` ("required for real-file
// end-to-end runs and for feeding 8 GPUs; the likely end-to-end bottleneck").
//
// The batched mirror of anno.d:36-52 on binary records: BGZF blocks are inflated in parallel (SAM text
// input is cut into lines and converted to the same binary records in parallel), the
// records of a batch are parsed in parallel (parse_clips / alignedLength / sc / sup through
// fadehost_prepare, anno.d:61-74), their 4-bit bases are copied as they are into the pinned view of
// a batch (the memcpy of INTEGRATION.md section 2), fadegpu_submit returns at once and the next
// batch is read while the GPU works; after fadegpu_wait the tags are written into the binary records
// (rs as a uint8 `C`, am/as/ar/ab as `Z`, anno.d:94-107) and the output is deflated in parallel (BAM, -b),
// stored (uncompressed BAM, -u) or formatted as SAM text.  Records keep their input order.
#pragma once
#include <omp.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../../../include/fadegpu.h"
#include "../../../include/fadehost.h"
#include "samio.hpp"

namespace bamfast {

using samio::get_i32;
using samio::get_u16;
using samio::get_u32;

// parallel BGZF inflate: raw blocks are read sequentially, decompressed side by side
class BulkReader {
public:
    BulkReader(FILE *f, const std::string &pre, int threads) : f_(f), pre_(pre), threads_(std::max(1, threads)) {}
    // appends the payload of up to max_blocks further blocks to out; false when the file is exhausted
    bool more(std::vector<uint8_t> &out, int max_blocks)
    {
        struct Blk { std::vector<uint8_t> c; uint32_t isize, crc; size_t off; };
        std::vector<Blk> blks;
        size_t total = 0;
        while ((int)blks.size() < max_blocks) {
            uint8_t h[12];
            const size_t got = raw(h, 12);
            if (got == 0) break;
            if (got != 12 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { bad_ = true; return false; }
            const uint32_t xlen = get_u16(h + 10);
            std::vector<uint8_t> extra(xlen);
            if (raw(extra.data(), xlen) != xlen) { bad_ = true; return false; }
            int64_t bsize = -1;
            for (size_t i = 0; i + 4 <= xlen;) {
                const uint32_t sl = get_u16(&extra[i + 2]);
                if (extra[i] == 'B' && extra[i + 1] == 'C' && sl == 2 && i + 6 <= xlen) bsize = get_u16(&extra[i + 4]);
                i += 4 + sl;
            }
            const int64_t clen = bsize - xlen - 19;
            if (bsize < 0 || clen < 0 || clen > 0x10000) { bad_ = true; return false; }
            Blk b;
            b.c.resize((size_t)clen + 16);     // trailer + the read slack of the built-in decoder
            if (raw(b.c.data(), (size_t)clen + 8) != (size_t)clen + 8) { bad_ = true; return false; }
            b.crc = get_u32(&b.c[(size_t)clen]);
            b.isize = get_u32(&b.c[(size_t)clen + 4]);
            if (b.isize > 0x10000) { bad_ = true; return false; }   // SAMv1 4.1: a block inflates to at most 64 KiB
            b.off = total;
            total += b.isize;
            blks.push_back(std::move(b));
        }
        if (blks.empty()) return false;
        const size_t base = out.size();
        out.resize(base + total);
        int bad = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(| : bad) num_threads(threads_)
        for (long k = 0; k < (long)blks.size(); ++k) {
            const Blk &b = blks[(size_t)k];
            if (!samio::inflate_block(b.c.data(), b.c.size() - 16, out.data() + base + b.off, b.isize, b.crc)) bad = 1;
        }
        if (bad) bad_ = true;
        return !bad;
    }
    bool bad() const { return bad_; }

private:
    size_t raw(uint8_t *d, size_t n)
    {
        size_t got = 0;
        if (pre_pos_ < pre_.size()) {
            got = std::min(n, pre_.size() - pre_pos_);
            memcpy(d, pre_.data() + pre_pos_, got);
            pre_pos_ += got;
        }
        if (got < n) got += fread(d + got, 1, n - got, f_);
        return got;
    }
    FILE *f_;
    std::string pre_;
    size_t pre_pos_ = 0;
    int threads_;
    bool bad_ = false;
};

// SAM text as a source of binary records: lines are cut sequentially and converted side by side
// (samio::sam_to_bam), appended to the stream as [block_size][record] exactly like a BAM payload
class SamTextSource {
public:
    SamTextSource(FILE *f, const std::string &pre, int threads) : f_(f), threads_(std::max(1, threads)) { buf_.assign(pre.begin(), pre.end()); }
    // header lines (up to the first record) into hdr
    bool header(samio::Header &hdr)
    {
        std::string line;
        for (;;) {
            const int rc = peek_line(line);
            if (rc <= 0 || line.empty() || line[0] != '@') return rc >= 0;
            hdr.add_line(line);
            pos_ = next_;
        }
    }
    // up to max_lines further records appended to out; false when the input is exhausted (or bad())
    bool more(std::vector<uint8_t> &out, const samio::Header &hdr, int max_lines)
    {
        std::vector<std::string> lines;
        std::string line;
        while ((int)lines.size() < max_lines) {
            const int rc = peek_line(line);
            if (rc <= 0) break;
            pos_ = next_;
            if (!line.empty()) lines.push_back(line);
        }
        if (lines.empty()) return false;
        std::vector<std::string> recs(lines.size());
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) num_threads(threads_)
        for (long k = 0; k < (long)lines.size(); ++k)
            if (!samio::sam_to_bam(lines[(size_t)k], hdr, recs[(size_t)k])) bad = 1;
        if (bad) {
            for (size_t k = 0; k < lines.size(); ++k)
                if (recs[k].empty()) { fprintf(stderr, "fade-b200: malformed SAM record: %.200s\n", lines[k].c_str()); break; }
            bad_ = true;
            return false;
        }
        size_t tot = 0;
        for (const auto &r : recs) tot += 4 + r.size();
        size_t w = out.size();
        out.resize(w + tot);
        for (const auto &r : recs) {
            const uint32_t n = (uint32_t)r.size();
            for (int i = 0; i < 4; ++i) out[w + (size_t)i] = (uint8_t)(n >> (8 * i));
            memcpy(&out[w + 4], r.data(), r.size());
            w += 4 + r.size();
        }
        return true;
    }
    bool bad() const { return bad_; }

private:
    // the line starting at pos_ (without CR/LF) and where the next one starts; 0 = end of input
    int peek_line(std::string &line)
    {
        for (;;) {
            const char *s = buf_.data() + pos_;
            const char *e = static_cast<const char *>(memchr(s, '\n', buf_.size() - pos_));
            if (e) {
                line.assign(s, (size_t)(e - s));
                next_ = pos_ + (size_t)(e - s) + 1;
                break;
            }
            if (eof_) {
                if (pos_ == buf_.size()) return 0;
                line.assign(s, buf_.size() - pos_);
                next_ = buf_.size();
                break;
            }
            buf_.erase(buf_.begin(), buf_.begin() + (long)pos_);   // keep the partial line, read on
            pos_ = 0;
            const size_t old = buf_.size();
            buf_.resize(old + (8u << 20));
            const size_t got = fread(buf_.data() + old, 1, 8u << 20, f_);
            buf_.resize(old + got);
            if (got == 0) eof_ = true;
        }
        if (!line.empty() && line.back() == '\r') line.pop_back();
        return 1;
    }
    FILE *f_;
    int threads_;
    std::vector<char> buf_;
    size_t pos_ = 0, next_ = 0;
    bool eof_ = false, bad_ = false;
};

// BGZF blocks of `data`, compressed side by side, written in order (no EOF marker)
inline void write_blocks(FILE *f, const uint8_t *data, size_t n, int level, int threads)
{
    constexpr size_t kBlock = 0xff00;
    const long nb = (long)((n + kBlock - 1) / kBlock);
    std::vector<std::string> out((size_t)nb);
#pragma omp parallel for schedule(dynamic, 4) num_threads(std::max(1, threads))
    for (long k = 0; k < nb; ++k) {
        const uint8_t *src = data + (size_t)k * kBlock;
        const size_t len = std::min(kBlock, n - (size_t)k * kBlock);
        std::string &o = out[(size_t)k];
        o.resize(0x10000 + 64);
        const size_t clen = samio::deflate_block(src, len, reinterpret_cast<uint8_t *>(&o[18]), o.size() - 26, level);
        static const uint8_t head[16] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0 };
        memcpy(&o[0], head, 16);
        const uint32_t bsize = (uint32_t)(clen + 25);
        o[16] = (char)(bsize & 0xff); o[17] = (char)(bsize >> 8);
        const uint32_t crc = (uint32_t)crc32(crc32(0, nullptr, 0), src, (uInt)len);
        for (int i = 0; i < 4; ++i) { o[18 + clen + (size_t)i] = (char)(crc >> (8 * i)); o[22 + clen + (size_t)i] = (char)((uint32_t)len >> (8 * i)); }
        o.resize(26 + clen);
    }
    for (const auto &o : out) fwrite(o.data(), 1, o.size(), f);
}

// bytes of one aux field starting at its 2-letter tag; 0 = damaged
inline size_t aux_field_size(const uint8_t *p, const uint8_t *end)
{
    if (p + 3 > end) return 0;
    const char ty = (char)p[2];
    size_t v;
    switch (ty) {
    case 'A': case 'c': case 'C': v = 1; break;
    case 's': case 'S': v = 2; break;
    case 'i': case 'I': case 'f': v = 4; break;
    case 'd': v = 8; break;
    case 'Z': case 'H': {
        const void *z = memchr(p + 3, 0, (size_t)(end - (p + 3)));
        if (!z) return 0;
        v = (size_t)(static_cast<const uint8_t *>(z) - (p + 3)) + 1;
        break;
    }
    case 'B': {
        if (p + 8 > end) return 0;
        const char st = (char)p[3];
        const size_t w = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : (st == 'i' || st == 'I' || st == 'f') ? 4 : 0;
        if (!w) return 0;
        v = 5 + w * (size_t)get_u32(p + 4);
        break;
    }
    default: return 0;
    }
    return (p + 3 + v <= end) ? 3 + v : 0;
}

struct RecMeta {
    size_t off;                 // of the record's block_size field inside the slot buffer
    uint32_t size;              // block_size
    int32_t aligned_len, clip_left, clip_right;
    uint8_t rs_base;
};

struct Slot {
    std::vector<uint8_t> buf;   // the batch's records, as in the file
    std::vector<RecMeta> rec;
    fadegpu_ctx *ctx = nullptr; // the GPU this slot's batches go to
    fadegpu_batch *bt = nullptr;
    fadegpu_batch_view v{};
    bool in_flight = false;
};

struct Job {
    fadegpu_params prm;
    int device = 0;             // first device
    int n_gpus = 1;             // devices device .. device + n_gpus - 1, batches dealt round-robin
    int64_t batch_n = 1 << 20;
    int level = samio::kFastLevel;   // -b output: fastdeflate.hpp by default, --level 1..9 = zlib
    int con = 0;                // util.d:65-76: 0 SAM, 1 uBAM, 2 BAM
    std::string cl, version;
    std::string fasta_path;
};

// Records of a SAM or BAM input as one byte stream of [block_size][record] entries, plus its header
struct RecordInput {
    RecordInput(FILE *f, const std::string &pre, bool bam, int threads)
        : is_bam(bam), rd(f, bam ? pre : std::string(), threads), txt(f, bam ? std::string() : pre, threads)
    {
        if (const char *e = getenv("FADE_IO_BLOCKS")) blocks = std::max(1, std::min(4096, atoi(e)));   // tests: small pieces
    }
    int blocks = 512;            // BGZF blocks inflated per refill
    bool is_bam;
    BulkReader rd;
    SamTextSource txt;
    samio::Header hdr;
    std::vector<uint8_t> stream;
    size_t spos = 0;

    bool need(size_t bytes)   // make stream[spos, spos + bytes) available
    {
        while (stream.size() - spos < bytes)
            if (!(is_bam ? rd.more(stream, blocks) : txt.more(stream, hdr, 1 << 16))) return false;
        return true;
    }
    bool bad() const { return is_bam ? rd.bad() : txt.bad(); }
    // header (SAMv1 1.3 / 4.2); false (with a message) when it cannot be read
    bool read_header()
    {
        if (!is_bam) {
            if (!txt.header(hdr)) { fprintf(stderr, "fade-b200: cannot read the SAM header\n"); return false; }
            return true;
        }
        auto damaged = [] { fprintf(stderr, "fade-b200: damaged BAM input\n"); return false; };
        if (!need(12) || memcmp(&stream[spos], "BAM\1", 4) != 0) {
            if (bad()) return damaged();
            fprintf(stderr, "fade-b200: not a BAM file\n");
            return false;
        }
        const uint32_t l_text = get_u32(&stream[spos + 4]);
        if (!need(12 + (size_t)l_text)) return damaged();
        std::string text(reinterpret_cast<const char *>(&stream[spos + 8]), l_text);
        text.resize(strnlen(text.c_str(), text.size()));
        bool has_sq = false;
        for (size_t a = 0; a < text.size();) {
            size_t e = text.find('\n', a);
            if (e == std::string::npos) e = text.size();
            if (e > a) { hdr.add_line(text.substr(a, e - a)); has_sq |= text.compare(a, 3, "@SQ") == 0; }
            a = e + 1;
        }
        spos += 8 + l_text;
        const uint32_t n_ref = get_u32(&stream[spos]);
        spos += 4;
        for (uint32_t r = 0; r < n_ref; ++r) {
            if (!need(4)) return damaged();
            const uint32_t l = get_u32(&stream[spos]);
            if (!need(8 + (size_t)l)) return damaged();
            std::string nm(reinterpret_cast<const char *>(&stream[spos + 4]), l);
            nm.resize(strnlen(nm.c_str(), nm.size()));
            // the binary reference list is authoritative when the text has no @SQ lines
            if (!has_sq) hdr.add_line("@SQ\tSN:" + nm + "\tLN:" + std::to_string(get_u32(&stream[spos + 4 + l])));
            spos += 8 + l;
        }
        return true;
    }
};

// the header as SAM text (con 0) or as the first BGZF block(s) of a BAM file (util.d:65-76)
inline void write_header(const samio::Header &hdr, int con, int threads, int bam_level = samio::kFastLevel)
{
    if (con == 0) {
        for (const auto &l : hdr.lines) { fwrite(l.data(), 1, l.size(), stdout); fputc('\n', stdout); }
        return;
    }
    std::string text, o("BAM\1", 4);
    for (const auto &l : hdr.lines) { text += l; text += '\n'; }
    samio::put_u32(o, (uint32_t)text.size());
    o += text;
    samio::put_u32(o, (uint32_t)hdr.names.size());
    for (size_t r = 0; r < hdr.names.size(); ++r) {
        samio::put_u32(o, (uint32_t)hdr.names[r].size() + 1);
        o += hdr.names[r]; o.push_back('\0');
        samio::put_u32(o, (uint32_t)hdr.lens[r]);
    }
    write_blocks(stdout, reinterpret_cast<const uint8_t *>(o.data()), o.size(), con == 1 ? 0 : bam_level, threads);
}

inline void write_eof_marker()
{
    static const uint8_t eof[28] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    fwrite(eof, 1, sizeof(eof), stdout);   // SAMv1 4.1.2
}

// `fade-b200 view --count`: number of records (parallel inflate, nothing decoded), for checks at scale
inline int count_records(FILE *fin, const std::string &pre, bool is_bam, int threads)
{
    RecordInput in(fin, pre, is_bam, threads);
    if (!in.read_header()) return 1;
    long long n = 0;
    for (;;) {
        while (in.spos + 4 <= in.stream.size()) {
            const uint32_t bs = get_u32(&in.stream[in.spos]);
            if (in.spos + 4 + (size_t)bs > in.stream.size()) break;
            in.spos += 4 + (size_t)bs;
            ++n;
        }
        in.stream.erase(in.stream.begin(), in.stream.begin() + (long)in.spos);
        in.spos = 0;
        const size_t have = in.stream.size();
        if (!in.need(have + 1)) {
            if (in.bad() || have) { fprintf(stderr, "fade-b200: damaged or truncated input\n"); return 1; }
            break;
        }
    }
    printf("%lld\n", n);
    return 0;
}

// `fade-b200 view --bulk`: every record through the parallel readers and writers of the annotate loop
// (format conversion only; exercises this file's I/O without a GPU)
inline int copy_records(FILE *fin, const std::string &pre, bool is_bam, int con, int threads)
{
    RecordInput in(fin, pre, is_bam, threads);
    if (!in.read_header()) return 1;
    write_header(in.hdr, con, threads);
    for (;;) {
        // whole records available in the stream
        size_t end = in.spos;
        while (end + 4 <= in.stream.size()) {
            const uint32_t bs = get_u32(&in.stream[end]);
            if (end + 4 + (size_t)bs > in.stream.size()) break;
            end += 4 + (size_t)bs;
        }
        if (end > in.spos) {
            if (con != 0) write_blocks(stdout, &in.stream[in.spos], end - in.spos, con == 1 ? 0 : samio::kFastLevel, threads);
            else {
                std::string line;
                for (size_t p = in.spos; p < end;) {
                    const uint32_t bs = get_u32(&in.stream[p]);
                    if (!samio::bam_to_sam(&in.stream[p + 4], bs, in.hdr, line)) { fprintf(stderr, "fade-b200: damaged BAM record\n"); return 1; }
                    fwrite(line.data(), 1, line.size(), stdout); fputc('\n', stdout);
                    p += 4 + (size_t)bs;
                }
            }
            in.stream.erase(in.stream.begin(), in.stream.begin() + (long)end);
            in.spos = 0;
        }
        const size_t have = in.stream.size() - in.spos;   // bytes of an incomplete record (none at the end of a good file)
        if (!in.need(have + 1)) {   // nothing more to read
            if (in.bad()) { if (is_bam) fprintf(stderr, "fade-b200: damaged BAM input\n"); return 1; }
            if (have) { fprintf(stderr, "fade-b200: truncated BAM record\n"); return 1; }
            break;
        }
    }
    if (con != 0) write_eof_marker();
    fflush(stdout);
    return 0;
}

// the whole command; returns the process exit code
inline int annotate_records(FILE *fin, const std::string &pre, bool is_bam, const Job &job,
                            bool (*read_fasta)(const std::string &, std::map<std::string, std::string> &))
{
    const int threads = job.prm.host_threads > 0 ? job.prm.host_threads : omp_get_max_threads();
    RecordInput in(fin, pre, is_bam, threads);
    if (!in.read_header()) return 1;
    samio::Header &hdr = in.hdr;
    std::vector<uint8_t> &stream = in.stream;
    size_t &spos = in.spos;
    auto need = [&](size_t bytes) { return in.need(bytes); };
    auto source_bad = [&]() { return in.bad(); };
    std::string last_pg;
    for (const auto &l : hdr.lines)
        if (l.compare(0, 3, "@PG") == 0) {
            const size_t a = l.find("\tID:");
            if (a != std::string::npos) last_pg = l.substr(a + 4, l.find('\t', a + 4) - a - 4);
        }
    std::string pg = "@PG\tID:fade-annotate\tPN:fade\tVN:" + job.version;   // anno.d:25-32
    if (!last_pg.empty()) pg += "\tPP:" + last_pg;
    pg += "\tCL:" + job.cl;
    hdr.add_line(pg);
    const int level = job.con == 1 ? 0 : job.level;
    write_header(hdr, job.con, threads, job.level);

    // ---- reference (anno.d:23) ----
    std::map<std::string, std::string> fasta;
    if (!read_fasta(job.fasta_path, fasta)) { fprintf(stderr, "fade-b200: cannot read %s\n", job.fasta_path.c_str()); return 1; }
    std::vector<const char *> cnames, cseqs;
    for (size_t t = 0; t < hdr.names.size(); ++t) {
        auto it = fasta.find(hdr.names[t]);
        if (it == fasta.end() || (int64_t)it->second.size() < hdr.lens[t]) {
            fprintf(stderr, "fade-b200: contig %s missing or shorter than @SQ LN in the FASTA\n", hdr.names[t].c_str());
            return 1;
        }
        cnames.push_back(hdr.names[t].c_str());
        cseqs.push_back(it->second.data());
    }
    fadegpu_params prm = job.prm;
    prm.flags |= FADEGPU_F_NO_SCATTER | FADEGPU_F_TAGS_ONLY;   // only rs / am / as / ar / ab leave this loop
    // One ctx (and its submit thread) per GPU; the reference is packed and uploaded once and copied GPU to GPU
    // (NVLink peer copy) to the others: every GPU holds its own copy, reads are independent, nothing is exchanged
    // on the data path.  The reference's merge is its mutex-guarded writer (anno.d:47-49); here the one writer
    // below emits the batches in input order.
    int n_dev = 0;
    if (fadegpu_device_count(&n_dev) != 0 || job.device < 0 || job.device + job.n_gpus > n_dev) {
        fprintf(stderr, "fade-b200: --device %d --gpus %d asks for devices this machine does not have (%d visible)\n", job.device, job.n_gpus, n_dev);
        return 1;
    }
    std::vector<fadegpu_ctx *> ctxs((size_t)job.n_gpus, nullptr);
    const double t_ref0 = omp_get_wtime();
    for (int g = 0; g < job.n_gpus; ++g) {
        if (fadegpu_create(job.device + g, &prm, &ctxs[(size_t)g]) != 0) { fprintf(stderr, "fade-b200: %s\n", fadegpu_last_error(nullptr)); return 1; }
        if (hdr.names.empty()) continue;
        const int rc = g == 0 ? fadegpu_load_reference(ctxs[0], (int32_t)hdr.names.size(), cnames.data(), hdr.lens.data(), cseqs.data())
                              : fadegpu_share_reference(ctxs[(size_t)g], ctxs[0]);
        if (rc != 0) { fprintf(stderr, "fade-b200: %s\n", fadegpu_last_error(ctxs[(size_t)g])); return 1; }
    }
    const double t_ref = omp_get_wtime() - t_ref0;
    fasta.clear();

    const int64_t max_seq = job.batch_n * 160;
    const int n_slots = 2 * job.n_gpus;          // two batches per GPU: one computing, one being read / written
    std::vector<Slot> slot((size_t)n_slots);
    for (int i = 0; i < n_slots; ++i) {
        Slot &s = slot[(size_t)i];
        s.ctx = ctxs[(size_t)(i % job.n_gpus)];
        if (fadegpu_alloc_batch(s.ctx, job.batch_n, max_seq, &s.bt) != 0 || fadegpu_get_batch_view(s.bt, &s.v) != 0) {
            fprintf(stderr, "fade-b200: %s\n", fadegpu_last_error(s.ctx));
            return 1;
        }
    }
    long long n_total = 0, n_art = 0, n_sc = 0, n_oversize = 0;
    int rc_all = 0;
    double t_read = 0, t_parse = 0, t_wait = 0, t_tag = 0, t_write = 0;   // FADE_TIMING=1 prints them
    const double t_start = omp_get_wtime();

    // records of the next batch -> slot (parse + fill the pinned view); false when there are none
    auto load = [&](Slot &s) -> bool {
        s.rec.clear();
        double t0 = omp_get_wtime();
        int64_t seq_bytes = 0;
        while ((int64_t)s.rec.size() < job.batch_n) {
            if (!need(4)) break;
            const uint32_t bs = get_u32(&stream[spos]);
            if (bs < 32 || !need(4 + (size_t)bs)) { if (!source_bad()) fprintf(stderr, "fade-b200: truncated BAM record\n"); rc_all = 1; break; }
            const int32_t l_seq = get_i32(&stream[spos + 4 + 16]);
            {   // the variable-length fields must fit the record (64-bit arithmetic: the sizes come from the file)
                const uint64_t l_name = stream[spos + 4 + 8], n_cig = get_u16(&stream[spos + 4 + 12]);
                if (l_seq < 0 || 32 + l_name + 4 * n_cig + ((uint64_t)l_seq + 1) / 2 + (uint64_t)l_seq > (uint64_t)bs) {
                    fprintf(stderr, "fade-b200: damaged BAM record\n"); rc_all = 1; break;
                }
            }
            if (seq_bytes + (l_seq + 1) / 2 + 1024 > max_seq) {
                if (s.rec.empty()) { fprintf(stderr, "fade-b200: a read of %d bases does not fit a batch (raise --batch)\n", l_seq); rc_all = 1; }
                break;
            }
            seq_bytes += (l_seq + 1) / 2;
            RecMeta m{};
            m.off = spos; m.size = bs;
            s.rec.push_back(m);
            spos += 4 + (size_t)bs;
        }
        if (source_bad()) { if (is_bam) fprintf(stderr, "fade-b200: damaged BAM input\n"); rc_all = 1; }
        if (s.rec.empty()) return false;
        t_read += omp_get_wtime() - t0; t0 = omp_get_wtime();
        // the slot takes the buffer (no copy, its capacity is reused two batches later); the unread tail moves on
        s.buf.swap(stream);
        stream.assign(s.buf.begin() + (long)spos, s.buf.end());
        spos = 0;
        const long n = (long)s.rec.size();
        fadegpu_batch_view &v = s.v;
        int64_t off = 0;
        for (long k = 0; k < n; ++k) {   // offsets of the bases inside the view
            v.meta[k].seq_off = (uint32_t)off;
            off += (get_i32(&s.buf[s.rec[(size_t)k].off + 4 + 16]) + 1) / 2;
        }
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) num_threads(threads)
        for (long k = 0; k < n; ++k) {
            RecMeta &m = s.rec[(size_t)k];
            const uint8_t *p = &s.buf[m.off + 4], *end = p + m.size;
            const uint32_t l_name = p[8], n_cig = get_u16(p + 12), flag = get_u16(p + 14);
            const int32_t l_seq = get_i32(p + 16);
            const uint8_t *cig = p + 32 + l_name, *seq = cig + 4ull * n_cig, *qual = seq + (size_t)(l_seq + 1) / 2, *aux = qual + (size_t)l_seq;
            if (aux > end) { bad = 1; continue; }
            uint32_t cigar[64];
            std::vector<uint32_t> big;
            uint32_t *cg = cigar;
            if (n_cig > 64) { big.resize(n_cig); cg = big.data(); }
            memcpy(cg, cig, 4ull * n_cig);
            bool has_sa = false;
            for (const uint8_t *a = aux; a < end;) {
                const size_t sz = aux_field_size(a, end);
                if (!sz) { bad = 1; break; }
                if (a[0] == 'S' && a[1] == 'A') has_sa = true;
                a += sz;
            }
            fadehost_record hr;
            hr.flag = (int32_t)flag; hr.has_sa = has_sa; hr.cigar = cg; hr.n_cigar = (int32_t)n_cig;
            hr.seq4 = seq; hr.qual = qual; hr.l_qseq = l_seq; hr.tid = get_i32(p); hr.pos = get_i32(p + 4);
            fadehost_prepare(&hr, &m.aligned_len, &m.clip_left, &m.clip_right, &m.rs_base);   // anno.d:61-74
            // the compact layout of the batch: one gate byte + one 32-byte record per read, bases as they are in the file
            fadegpu_read_meta &mm = v.meta[k];
            memcpy(v.seq4 + mm.seq_off, seq, (size_t)(l_seq + 1) / 2);
            mm.pos = hr.pos; mm.l_qseq = l_seq; mm.tid = hr.tid; mm.aligned_len = m.aligned_len;
            mm.clip_left = (uint32_t)m.clip_left; mm.clip_right = (uint32_t)m.clip_right;
            v.gate[k] = (uint8_t)std::min<uint32_t>(255u, std::max(mm.clip_left, mm.clip_right));
        }
        if (bad) { fprintf(stderr, "fade-b200: damaged BAM record\n"); rc_all = 1; return false; }
        if (fadegpu_submit_compact(s.ctx, s.bt, n, off) != 0) { fprintf(stderr, "fade-b200: %s\n", fadegpu_last_error(s.ctx)); rc_all = 1; return false; }
        t_parse += omp_get_wtime() - t0;
        s.in_flight = true;
        return true;
    };

    // results of the slot -> tagged records -> stdout
    auto emit = [&](Slot &s) -> bool {
        s.in_flight = false;
        fadegpu_results_view rv;
        double t0 = omp_get_wtime();
        if (fadegpu_wait(s.ctx, s.bt) != 0 || fadegpu_get_results(s.bt, &rv) != 0) {
            fprintf(stderr, "fade-b200: %s\n", fadegpu_last_error(s.ctx));
            return false;
        }
        {
            fadegpu_stats st;
            if (fadegpu_get_stats(s.bt, &st) == 0) n_oversize += st.n_oversize;
        }
        t_wait += omp_get_wtime() - t0; t0 = omp_get_wtime();
        const long n = (long)s.rec.size();
        const int T = (int)std::max<long>(1, std::min<long>(threads, n / 256));
        std::vector<std::string> part((size_t)T);
        std::vector<long long> art((size_t)T, 0), sc((size_t)T, 0);
        int bad = 0;
#pragma omp parallel for schedule(static, 1) reduction(| : bad) num_threads(T)
        for (int t = 0; t < T; ++t) {
            std::string &o = part[(size_t)t];
            std::string am, as_, ar, ab, rec, line;
            std::vector<uint32_t> big;
            for (long k = n * t / T; k < n * (t + 1) / T; ++k) {
                const RecMeta &m = s.rec[(size_t)k];
                const uint8_t *p = &s.buf[m.off + 4], *end = p + m.size;
                const uint32_t l_name = p[8], n_cig = get_u16(p + 12);
                const int32_t l_seq = get_i32(p + 16), tid = get_i32(p);
                const uint8_t *cig = p + 32 + l_name, *seq = cig + 4ull * n_cig, *qual = seq + (size_t)(l_seq + 1) / 2, *aux = qual + (size_t)l_seq;
                big.resize(n_cig);
                memcpy(big.data(), cig, 4ull * n_cig);
                fadehost_record hr;
                hr.flag = (int32_t)get_u16(p + 14); hr.has_sa = (m.rs_base & FADE_RS_SUP) != 0; hr.cigar = big.data(); hr.n_cigar = (int32_t)n_cig;
                hr.seq4 = seq; hr.qual = qual; hr.l_qseq = l_seq; hr.tid = tid; hr.pos = get_i32(p + 4);
                const char *cname = (tid >= 0 && (size_t)tid < hdr.names.size()) ? hdr.names[(size_t)tid].c_str() : "";
                const size_t cap = (size_t)4 * (size_t)std::max(l_seq, 0) + 512 + strlen(cname);
                am.resize(cap); as_.resize(cap); ar.resize(cap); ab.resize(cap);
                const int32_t ri = rv.result_index[k];
                static const uint32_t no_ops[1] = { 0 };
                const fadegpu_result *res = ri >= 0 ? &rv.results[ri] : nullptr;
                uint8_t rs = 0;
                const int rc = fadehost_finish(&hr, cname, m.rs_base, m.clip_left, m.clip_right, m.aligned_len, s.v.flags[k],
                                               res ? rv.win_start[ri] : 0, res ? res->beg_ref : 0, res ? res->n_ops : 0,
                                               res ? res->ops : no_ops, &rs, &am[0], &as_[0], &ar[0], &ab[0], cap);
                if (rc < 0) { bad = 1; continue; }
                // the record without the tags annotate (re)writes, then rs / am / as / ar / ab (anno.d:94-106)
                rec.assign(reinterpret_cast<const char *>(p), (size_t)(aux - p));
                for (const uint8_t *a = aux; a < end;) {
                    const size_t sz = aux_field_size(a, end);
                    if (!sz) { bad = 1; break; }
                    const bool ours = (a[0] == 'r' && a[1] == 's') || (a[0] == 'a' && (a[1] == 'm' || a[1] == 's' || a[1] == 'r' || a[1] == 'b'));
                    if (!ours) rec.append(reinterpret_cast<const char *>(a), sz);
                    a += sz;
                }
                rec += "rsC"; rec.push_back((char)rs);
                if (rc == 1) {
                    rec += "amZ"; rec += am.c_str(); rec.push_back('\0');
                    rec += "asZ"; rec += as_.c_str(); rec.push_back('\0');
                    rec += "arZ"; rec += ar.c_str(); rec.push_back('\0');
                    rec += "abZ"; rec += ab.c_str(); rec.push_back('\0');
                    ++art[(size_t)t];
                }
                sc[(size_t)t] += rs & 1;
                if (job.con == 0) {
                    if (!samio::bam_to_sam(reinterpret_cast<const uint8_t *>(rec.data()), rec.size(), hdr, line)) { bad = 1; continue; }
                    o += line; o += '\n';
                } else {
                    samio::put_u32(o, (uint32_t)rec.size());
                    o += rec;
                }
            }
        }
        if (bad) { fprintf(stderr, "fade-b200: damaged BAM record\n"); return false; }
        for (int t = 0; t < T; ++t) { n_art += art[(size_t)t]; n_sc += sc[(size_t)t]; }
        n_total += n;
        t_tag += omp_get_wtime() - t0; t0 = omp_get_wtime();
        if (job.con == 0) for (const auto &o : part) fwrite(o.data(), 1, o.size(), stdout);
        else {
            std::string all;
            size_t tot = 0;
            for (const auto &o : part) tot += o.size();
            all.reserve(tot);
            for (const auto &o : part) all += o;
            write_blocks(stdout, reinterpret_cast<const uint8_t *>(all.data()), all.size(), level, threads);
        }
        t_write += omp_get_wtime() - t0;
        return true;
    };

    // anno.d:44-50, batched: the slots form a ring that is filled in order (slot i goes to GPU i mod N) and emitted in
    // the same order, so the records keep their input order; a slot is emitted only when the ring wraps around to
    // the one before it, i.e. every GPU always has one batch computing while the host reads and writes others
    int cur = 0;
    for (;;) {
        const bool got = load(slot[(size_t)cur]);
        Slot &oldest = slot[(size_t)((cur + 1) % n_slots)];
        if (oldest.in_flight && !emit(oldest)) { rc_all = 1; break; }
        if (!got) break;
        cur = (cur + 1) % n_slots;
    }
    for (int k = 1; k <= n_slots && rc_all == 0; ++k) {      // drain in ring order, oldest first
        Slot &s = slot[(size_t)((cur + k) % n_slots)];
        if (s.in_flight && !emit(s)) rc_all = 1;
    }
    if (job.con != 0 && rc_all == 0) write_eof_marker();
    fflush(stdout);
    fprintf(stderr, "[fade-b200 annotate] %lld records, %lld soft-clipped, %lld with artifact tags\n", n_total, n_sc, n_art);
    if (n_oversize)
        fprintf(stderr, "[W::fade-b200 annotate] %lld reads were NOT realigned: their window exceeds 2^31 DP cells (rs lacks the artifact bits for them)\n",
                n_oversize);
    if (getenv("FADE_TIMING"))
        fprintf(stderr, "[fade-b200 annotate] %d threads, %d GPU(s); reference to the GPUs %.3f s; record loop %.3f s: inflate+split %.3f, "
                        "parse+fill+submit %.3f, wait for GPU %.3f, tag %.3f, deflate+write %.3f\n", threads, job.n_gpus, t_ref,
                omp_get_wtime() - t_start, t_read, t_parse, t_wait, t_tag, t_write);
    for (auto &s : slot) { if (s.in_flight) fadegpu_wait(s.ctx, s.bt); fadegpu_free_batch(s.bt); }
    for (auto *cx : ctxs) fadegpu_destroy(cx);
    return rc_all;
}

}  // namespace bamfast
