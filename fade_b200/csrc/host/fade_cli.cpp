// fade_cli.cpp -- C++ host driver `fade-b200 annotate`: the batched mirror of fade's annotate()
// (source/anno.d:16-52) on top of the C ABI (include/fadegpu.h, include/fadehost.h).
//
// The reference host is D + dhtslib/htslib; neither is available here, so record I/O is samio.hpp:
// SAM text or BAM in (recognised by content), SAM text (default), BAM (-b) or uncompressed BAM (-u)
// out, as util.d:65-76; the reference is a plain FASTA.  Flags and tag schema are the reference's:
//     fade-b200 annotate [-t N] [--min-length N] [-w N | --window-size N] [-b|-u] <in.sam|in.bam|-> <ref.fa>  > out
// Every record gets rs:i (anno.d:94); artifact records get am/as/ar/ab:Z (anno.d:98-107); a
// @PG ID:fade-annotate line is appended to the header (anno.d:25-32).  Records are written in
// input order (the reference's order is unspecified, anno.d:19).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include "../../../include/fadegpu.h"
#include "../../../include/fadehost.h"
#include <parallel/algorithm>
#include "samio.hpp"
#include "bamfast.hpp"

namespace {

const char *kVersion = "fade-b200-0.1";

samio::LineSink *g_out = nullptr;   // stdout as SAM text, uBAM or BAM (util.d:65-76)

// BAM output converts its lines in batches, so a line that cannot be encoded is reported by a LATER put() or by the close
void report_unencodable() { fprintf(stderr, "fade-b200: cannot encode record for BAM output: %s\n", g_out->bad_line().c_str()); }

bool out_line(const std::string &l)
{
    if (g_out->put(l)) return true;
    report_unencodable();
    return false;
}

// flushes and closes stdout; the command's exit code
int close_output()
{
    if (g_out->close()) return 0;
    report_unencodable();
    return 1;
}

// Command-line options the way std.getopt with config.bundling reads them in the reference (app.d:73-107):
// short flags may be bundled (-cb), a value may be attached (-t4, -t=4, --threads=4) or follow (-t 4,
// --threads 4), "--" ends the options, "-" is a positional (stdin), and an unknown option is an error.
struct OptSpec { char s; const char *l; bool value; };
struct Options {
    std::map<std::string, std::string> val;   // by long name; flags map to "1"
    std::vector<std::string> pos;
    bool ok = true;
    bool has(const char *l) const { return val.count(l) != 0; }
    long long num(const char *l, long long dflt) const { auto it = val.find(l); return it == val.end() ? dflt : atoll(it->second.c_str()); }
};

Options parse_options(int argc, char **argv, int first, const std::vector<OptSpec> &spec)
{
    Options o;
    auto by_long = [&](const std::string &l) -> const OptSpec * { for (auto &x : spec) if (l == x.l) return &x; return nullptr; };
    auto by_short = [&](char c) -> const OptSpec * { for (auto &x : spec) if (x.s && x.s == c) return &x; return nullptr; };
    auto bad = [&](const std::string &what) { fprintf(stderr, "fade-b200: %s\n", what.c_str()); o.ok = false; };
    bool opts_done = false;
    for (int i = first; i < argc && o.ok; ++i) {
        const std::string a = argv[i];
        if (opts_done || a == "-" || a.empty() || a[0] != '-') { o.pos.push_back(a); continue; }
        if (a == "--") { opts_done = true; continue; }
        if (a[1] == '-') {
            const size_t eq = a.find('=');
            const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            const OptSpec *sp = by_long(name);
            if (!sp) { bad("unrecognized option --" + name); break; }
            if (!sp->value) { if (eq != std::string::npos) bad("option --" + name + " takes no value"); else o.val[sp->l] = "1"; continue; }
            if (eq != std::string::npos) o.val[sp->l] = a.substr(eq + 1);
            else if (i + 1 < argc) o.val[sp->l] = argv[++i];
            else bad("missing value for --" + name);
            continue;
        }
        for (size_t k = 1; k < a.size() && o.ok; ++k) {   // a bundle of short options
            const OptSpec *sp = by_short(a[k]);
            if (!sp) { bad(std::string("unrecognized option -") + a[k]); break; }
            if (!sp->value) { o.val[sp->l] = "1"; continue; }
            std::string v = a.substr(k + 1);
            if (!v.empty() && v[0] == '=') v.erase(0, 1);
            if (v.empty()) {
                if (i + 1 < argc) v = argv[++i];
                else { bad(std::string("missing value for -") + a[k]); break; }
            }
            o.val[sp->l] = v;
            break;
        }
    }
    return o;
}

// -b / --bam, -u / --ubam as in app.d:82-83,94 (con = bam << 1 | ubam); false (with a message) when both are given
bool output_container(const Options &o, int &con)
{
    con = (o.has("bam") ? 2 : 0) | (o.has("ubam") ? 1 : 0);
    if (con > 2) { fprintf(stderr, "fade-b200: -b and -u are exclusive\n"); return false; }
    return true;
}

void open_output(int con)
{
    static samio::LineSink sink(stdout, con == 0 ? samio::LineSink::SAM : con == 1 ? samio::LineSink::UBAM : samio::LineSink::BAM);
    g_out = &sink;
}

struct Rec {
    std::string line;            // the record without trailing newline (tags rs/am/as/ar/ab stripped)
    int32_t flag = 0, tid = -1, l_qseq = 0;
    int64_t pos = 0;
    bool has_sa = false;
    std::vector<uint32_t> cigar;
    std::vector<uint8_t> seq4, qual;
    int32_t aligned_len = 0, clip_left = 0, clip_right = 0;
    uint8_t rs_base = 0;
};

int nt16_of(char c)
{
    static const char tbl[] = "=ACMGRSVTWYHKDBN";
    c = (char)toupper((unsigned char)c);
    const char *p = strchr(tbl, c);
    return (p && c) ? (int)(p - tbl) : 15;
}

bool parse_cigar(const std::string &s, std::vector<uint32_t> &out)
{
    out.clear();
    if (s == "*") return true;
    static const char ops[] = "MIDNSHP=XB";
    uint64_t num = 0;
    bool have = false;
    for (char c : s) {
        if (c >= '0' && c <= '9') { num = num * 10 + (uint64_t)(c - '0'); have = true; continue; }
        const char *p = strchr(ops, c);
        if (!p || !have || num >= (1u << 28)) return false;
        out.push_back((uint32_t)(num << 4) | (uint32_t)(p - ops));
        num = 0; have = false;
    }
    return !have;
}

std::vector<std::string> split_tab(const std::string &s)
{
    std::vector<std::string> f;
    size_t a = 0;
    for (;;) {
        const size_t b = s.find('\t', a);
        if (b == std::string::npos) { f.push_back(s.substr(a)); break; }
        f.push_back(s.substr(a, b - a));
        a = b + 1;
    }
    return f;
}

// <path>.fai (samtools faidx: name, length, offset, linebases, linewidth): every contig is read with one
// fread and its line ends squeezed out in place, instead of line by line
bool read_fasta_indexed(const std::string &path, std::map<std::string, std::string> &seqs)
{
    std::ifstream fai(path + ".fai");
    if (!fai) return false;
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    std::string line;
    bool ok = true;
    while (ok && std::getline(fai, line)) {
        const auto c = split_tab(line);
        if (c.size() < 5) { ok = false; break; }
        const long long len = atoll(c[1].c_str()), off = atoll(c[2].c_str()), lb = atoll(c[3].c_str()), lw = atoll(c[4].c_str());
        if (len < 0 || off < 0 || lb <= 0 || lw < lb) { ok = false; break; }
        const long long full = len / lb, rest = len % lb;
        const size_t span = (size_t)(full * lw + rest);
        std::string &dst = seqs[c[0]];
        dst.resize(span);
        if (fseeko(f, (off_t)off, SEEK_SET) != 0) { ok = false; break; }
        const size_t got = fread(&dst[0], 1, span, f);
        // bytes up to the last base (the file may end without a line end)
        const size_t min_bytes = rest > 0 ? span : (full > 0 ? (size_t)((full - 1) * lw + lb) : 0);
        if (got < min_bytes) { ok = false; break; }
        size_t w = 0;
        for (size_t r = 0; r < got;) {   // keep linebases, skip the line end
            const size_t k = std::min<size_t>((size_t)lb, std::min<size_t>(got - r, (size_t)len - w));
            memmove(&dst[w], &dst[r], k);
            w += k; r += (size_t)lw;
            if (w == (size_t)len) break;
        }
        if (w != (size_t)len) { ok = false; break; }
        dst.resize((size_t)len);
    }
    fclose(f);
    if (!ok) seqs.clear();
    return ok;
}

bool read_fasta(const std::string &path, std::map<std::string, std::string> &seqs)
{
    if (read_fasta_indexed(path, seqs)) return true;   // no (usable) .fai: read the text
    std::ifstream in(path);
    if (!in) return false;
    std::string line, name;
    std::string *cur = nullptr;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '>') {
            name = line.substr(1, line.find_first_of(" \t", 1) - 1);
            cur = &seqs[name];
            cur->clear();
        } else if (cur) *cur += line;
    }
    return true;
}

int usage()
{
    fprintf(stderr,
            "fade-b200 <annotate|out|extract> ...\n"
            "fade-b200 out [-c] <annotated SAM or ->      removes (or with -c hard-clips) artifact reads\n"
            "fade-b200 extract <annotated SAM or ->       emits the artifacts in their re-mapped state\n"
            "fade-b200 annotate: marks artifact reads in bam tags (B200 implementation of `fade annotate`)\n"
            "fade-b200 view <SAM/BAM or ->                 copies the records (format conversion)\n"
            "fade-b200 sort -n <SAM/BAM or ->              name-sorted copy (for `out` without -c)\n"
            "every command: -b / --bam writes BAM, -u / --ubam uncompressed BAM, default SAM text; input may be SAM or BAM\n"
            "usage: fade-b200 annotate [options] <SAM/BAM or -> <FASTA>\n"
            "  -t, --threads N      host threads (default: all cores)\n"
            "      --min-length N   minimum soft-clip length considered (default 5)\n"
            "  -w, --window-size N  bases considered outside of the read region (default 300)\n"
            "      --batch N        records per GPU batch (default 1048576)\n"
            "      --level N        compression of -b output: fast (default: built-in encoder, the size of zlib level 6 on BAM\n"
            "                       records at several times its speed) or a zlib level 0-9\n"
            "      --device N       first CUDA device (default 0)\n"
            "      --gpus N         GPUs: batches are dealt round-robin to devices N0..N0+N-1, every GPU holds the reference\n"
            "                       (packed once, copied GPU to GPU), the records keep their input order (default 1)\n"
            "options may be bundled (-cb) and written --name=value, as in fade\n");
    return 0;
}

}  // namespace

static int cmd_annotate(int argc, char **argv, const std::string &cl)
{
    fadegpu_params prm;
    fadegpu_default_params(&prm);
    const Options opt = parse_options(argc, argv, 2, { { 't', "threads", true }, { 0, "min-length", true }, { 'w', "window-size", true },
                                                      { 0, "batch", true }, { 0, "device", true }, { 0, "gpus", true }, { 0, "level", true }, { 0, "text-path", false },
                                                      { 'h', "help", false }, { 'b', "bam", false }, { 'u', "ubam", false } });
    if (!opt.ok) { usage(); return 1; }
    if (opt.has("help")) return usage();
    prm.host_threads = (int32_t)opt.num("threads", prm.host_threads);
    prm.min_length = (int32_t)opt.num("min-length", prm.min_length);
    prm.window_size = (int32_t)opt.num("window-size", prm.window_size);
    const int64_t batch_n = opt.num("batch", 1 << 20);
    const int device = (int)opt.num("device", 0), n_gpus = (int)opt.num("gpus", 1);
    const int level = (!opt.has("level") || opt.val.at("level") == "fast") ? samio::kFastLevel : (int)opt.num("level", 6);
    const bool text_path = opt.has("text-path");   // the line-by-line SAM text loop (A/B check of bamfast.hpp)
    const std::vector<std::string> &pos_args = opt.pos;
    int con = 0;
    if (pos_args.size() != 2) { usage(); return pos_args.empty() ? 0 : 1; }
    if (!output_container(opt, con)) return 1;
    if (batch_n <= 0 || n_gpus < 1 || (text_path && n_gpus != 1) || level < samio::kFastLevel || level > 9) { fprintf(stderr, "fade-b200: bad --batch / --gpus / --level\n"); return 1; }
    fprintf(stderr, "[W::fade annotate] Output will keep the input order\n");

    FILE *fin_raw = pos_args[0] == "-" ? stdin : fopen(pos_args[0].c_str(), "rb");
    if (!fin_raw) { fprintf(stderr, "fade-b200: cannot open %s\n", pos_args[0].c_str()); return 1; }
    std::string pre(2, '\0');
    pre.resize(fread(&pre[0], 1, 2, fin_raw));
    if (!text_path) {
        // the record loop on binary records (bamfast.hpp): BAM as it is, SAM text converted on the way in;
        // --text-path keeps the first, line-by-line implementation below (A/B check)
        const bool is_bam = pre.size() == 2 && (uint8_t)pre[0] == 0x1f && (uint8_t)pre[1] == 0x8b;
        bamfast::Job job;
        job.prm = prm; job.device = device; job.n_gpus = n_gpus; job.batch_n = batch_n; job.level = level; job.con = con; job.cl = cl; job.version = kVersion;
        job.fasta_path = pos_args[1];
        return bamfast::annotate_records(fin_raw, pre, is_bam, job, read_fasta);
    }
    open_output(con);

    // ---- header ----
    samio::LineSource src;
    if (!src.open(fin_raw, pre)) { fprintf(stderr, "fade-b200: cannot read %s\n", pos_args[0].c_str()); return 1; }
    samio::LineSource *in = &src;
    std::vector<std::string> header;
    std::vector<std::string> sq_names;
    std::vector<int64_t> sq_len;
    std::string last_pg_id, line;
    bool have_line = false;
    while (in->getline(line)) {
        if (line.empty() || line[0] != '@') { have_line = true; break; }
        header.push_back(line);
        const auto f = split_tab(line);
        if (f[0] == "@SQ") {
            std::string sn; int64_t ln = 0;
            for (const auto &x : f) { if (x.rfind("SN:", 0) == 0) sn = x.substr(3); if (x.rfind("LN:", 0) == 0) ln = atoll(x.c_str() + 3); }
            sq_names.push_back(sn); sq_len.push_back(ln);
        } else if (f[0] == "@PG") {
            for (const auto &x : f) if (x.rfind("ID:", 0) == 0) last_pg_id = x.substr(3);
        }
    }
    // anno.d:25-32
    std::string pg = std::string("@PG\tID:fade-annotate\tPN:fade\tVN:") + kVersion;
    if (!last_pg_id.empty()) pg += "\tPP:" + last_pg_id;
    pg += "\tCL:" + cl;
    header.push_back(pg);
    for (const auto &h : header) out_line(h);

    // ---- reference: anno.d:23; contigs in @SQ order so that tid indexes them ----
    std::map<std::string, std::string> fasta;
    if (!read_fasta(pos_args[1], fasta)) { fprintf(stderr, "fade-b200: cannot read %s\n", pos_args[1].c_str()); return 1; }
    std::vector<const char *> cnames, cseqs;
    std::vector<int64_t> clens;
    std::map<std::string, int> tid_of;
    for (size_t t = 0; t < sq_names.size(); ++t) {
        auto it = fasta.find(sq_names[t]);
        if (it == fasta.end() || (int64_t)it->second.size() < sq_len[t]) {
            fprintf(stderr, "fade-b200: contig %s missing or shorter than @SQ LN in the FASTA\n", sq_names[t].c_str());
            return 1;
        }
        tid_of[sq_names[t]] = (int)t;
        cnames.push_back(sq_names[t].c_str());
        cseqs.push_back(it->second.data());
        clens.push_back(sq_len[t]);   // rec.h.targetLength(tid), analysis.d:56-58
    }
    fadegpu_ctx *ctx = nullptr;
    if (fadegpu_create(device, &prm, &ctx) != 0) { fprintf(stderr, "fade-b200: %s\n", fadegpu_last_error(nullptr)); return 1; }
    if (!sq_names.empty() &&
        fadegpu_load_reference(ctx, (int32_t)sq_names.size(), cnames.data(), clens.data(), cseqs.data()) != 0) {
        fprintf(stderr, "fade-b200: %s\n", fadegpu_last_error(ctx));
        return 1;
    }

    // ---- batched record loop: anno.d:44-50 ----
    fadegpu_batch *bt = nullptr;
    fadegpu_batch_view v;
    const int64_t max_seq = batch_n * 160;
    if (fadegpu_alloc_batch(ctx, batch_n, max_seq, &bt) != 0 || fadegpu_get_batch_view(bt, &v) != 0) {
        fprintf(stderr, "fade-b200: %s\n", fadegpu_last_error(ctx));
        return 1;
    }
    std::vector<Rec> recs;
    recs.reserve((size_t)std::min<int64_t>(batch_n, 1 << 16));
    int64_t n_total = 0, n_art = 0, n_sc = 0;
    auto flush = [&]() -> int {
        const int64_t n = (int64_t)recs.size();
        if (n == 0) return 0;
        int64_t off = 0;
        for (int64_t k = 0; k < n; ++k) {
            Rec &r = recs[(size_t)k];
            v.seq_off[k] = off;
            memcpy(v.seq4 + off, r.seq4.data(), r.seq4.size());
            off += (int64_t)r.seq4.size();
            v.l_qseq[k] = r.l_qseq; v.tid[k] = r.tid; v.pos[k] = r.pos;
            v.aligned_len[k] = r.aligned_len; v.clip_left[k] = r.clip_left; v.clip_right[k] = r.clip_right;
        }
        v.seq_off[n] = off;
        if (fadegpu_submit(ctx, bt, n) != 0 || fadegpu_wait(ctx, bt) != 0) {
            fprintf(stderr, "fade-b200: %s\n", fadegpu_last_error(ctx));
            return 1;
        }
        std::string am, as_, ar, ab;
        for (int64_t k = 0; k < n; ++k) {
            Rec &r = recs[(size_t)k];
            fadehost_record hr;
            hr.flag = r.flag; hr.has_sa = r.has_sa; hr.cigar = r.cigar.data(); hr.n_cigar = (int32_t)r.cigar.size();
            hr.seq4 = r.seq4.data(); hr.qual = r.qual.data(); hr.l_qseq = r.l_qseq; hr.tid = r.tid; hr.pos = r.pos;
            const size_t cap = (size_t)4 * r.l_qseq + 512 + (r.tid >= 0 ? sq_names[(size_t)r.tid].size() : 0);
            am.resize(cap); as_.resize(cap); ar.resize(cap); ab.resize(cap);
            uint8_t rs = 0;
            const int rc = fadehost_finish(&hr, r.tid >= 0 ? sq_names[(size_t)r.tid].c_str() : "", r.rs_base, r.clip_left,
                                           r.clip_right, r.aligned_len, v.flags[k], v.win_start[k], v.beg_ref[k],
                                           v.n_ops[k], v.ops + (size_t)k * FADEGPU_MAX_OPS, &rs, &am[0], &as_[0], &ar[0],
                                           &ab[0], cap);
            if (rc < 0) { fprintf(stderr, "fade-b200: tag buffer too small\n"); return 1; }
            std::string o = r.line;
            o += "\trs:i:" + std::to_string((unsigned)rs);                                // anno.d:94
            if (rc == 1) {                                                                // anno.d:98-107
                o += "\tam:Z:"; o += am.c_str(); o += "\tas:Z:"; o += as_.c_str();
                o += "\tar:Z:"; o += ar.c_str(); o += "\tab:Z:"; o += ab.c_str();
            }
            if (!out_line(o)) return 1;
            n_art += rc == 1;
            n_sc += rs & 1;
        }
        n_total += n;
        recs.clear();
        return 0;
    };

    int64_t seq_bytes = 0;
    while (have_line) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty()) {
            auto f = split_tab(line);
            if (f.size() < 11) { fprintf(stderr, "fade-b200: malformed SAM record: %s\n", line.c_str()); return 1; }
            Rec r;
            r.flag = atoi(f[1].c_str());
            auto it = tid_of.find(f[2]);
            r.tid = it == tid_of.end() ? -1 : it->second;
            r.pos = atoll(f[3].c_str()) - 1;
            if (!parse_cigar(f[5], r.cigar)) { fprintf(stderr, "fade-b200: bad CIGAR %s\n", f[5].c_str()); return 1; }
            const std::string &seq = f[9];
            r.l_qseq = seq == "*" ? 0 : (int32_t)seq.size();
            r.seq4.assign((size_t)(r.l_qseq + 1) / 2, 0);
            for (int i = 0; i < r.l_qseq; ++i) r.seq4[(size_t)i >> 1] |= (uint8_t)(nt16_of(seq[(size_t)i]) << ((~i & 1) << 2));
            r.qual.assign((size_t)r.l_qseq, 0xff);
            if (f[10] != "*") for (int i = 0; i < r.l_qseq && i < (int)f[10].size(); ++i) r.qual[(size_t)i] = (uint8_t)(f[10][(size_t)i] - 33);
            // keep the mandatory fields and every tag except the ones annotate (re)writes
            std::string out;
            for (size_t k = 0; k < f.size(); ++k) {
                if (k >= 11) {
                    const std::string tag = f[k].substr(0, 2);
                    if (tag == "SA") r.has_sa = true;
                    if (tag == "rs" || tag == "am" || tag == "as" || tag == "ar" || tag == "ab") continue;
                }
                if (k) out += '\t';
                out += f[k];
            }
            r.line.swap(out);
            fadehost_record hr;
            hr.flag = r.flag; hr.has_sa = r.has_sa; hr.cigar = r.cigar.data(); hr.n_cigar = (int32_t)r.cigar.size();
            hr.seq4 = r.seq4.data(); hr.qual = r.qual.data(); hr.l_qseq = r.l_qseq; hr.tid = r.tid; hr.pos = r.pos;
            fadehost_prepare(&hr, &r.aligned_len, &r.clip_left, &r.clip_right, &r.rs_base);   // anno.d:61-74
            // the bases must fit the pinned view: flush BEFORE a record that would not, refuse one that never can
            const int64_t nb = (int64_t)r.seq4.size();
            if (nb > max_seq) { fprintf(stderr, "fade-b200: a read of %d bases does not fit a batch (raise --batch)\n", r.l_qseq); return 1; }
            if (seq_bytes + nb > max_seq) { if (flush()) return 1; seq_bytes = 0; }
            seq_bytes += nb;
            recs.push_back(std::move(r));
            if ((int64_t)recs.size() == batch_n) { if (flush()) return 1; seq_bytes = 0; }
        }
        have_line = in->getline(line);
    }
    if (src.failed()) { fprintf(stderr, "fade-b200: damaged BAM input\n"); return 1; }
    if (flush()) return 1;
    if (close_output()) return 1;
    fprintf(stderr, "[fade-b200 annotate] %lld records, %lld soft-clipped, %lld with artifact tags\n", (long long)n_total,
            (long long)n_sc, (long long)n_art);
    fadegpu_free_batch(bt);
    fadegpu_destroy(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Consumers of the tags, pure host code on SAM text:
//   fade-b200 out [-c]   source/filter.d:15-91,127-269 + source/stats.d:45-72
//   fade-b200 extract    source/remap.d:11-87
// Uncertain point U9 (dhtslib / htslib are not available here): a blank `SAMRecord(header)` is
// htslib's bam_init1() = calloc, so every core field is 0 -> RNAME first contig, POS 1, RNEXT "=",
// PNEXT 1 in SAM text.  oracle/consumers.py restates the same functions independently.
// ------------------------------------------------------------------------------------------------
namespace {

struct SamRec {
    std::vector<std::string> f;   // 11 mandatory fields
    std::vector<std::string> tags;
    bool has(const char *k) const { for (auto &t : tags) if (t.compare(0, 2, k) == 0) return true; return false; }
    std::string tag(const char *k) const { for (auto &t : tags) if (t.compare(0, 2, k) == 0) return t.substr(5); return ""; }
    std::string line() const
    {
        std::string o;
        for (size_t k = 0; k < f.size(); ++k) { if (k) o += '\t'; o += f[k]; }
        for (auto &t : tags) { o += '\t'; o += t; }
        return o;
    }
};

// One record line split only as far as `out` needs to decide what happens to it: the name and where the optional
// fields start.  A record that passes through unchanged is written from `line` as it came in.
struct LineRec {
    std::string line;
    size_t name_len = 0, tags_at = 0;   // tags_at = offset of the first optional field, line.size() when there is none
    bool index()                        // false: fewer than 11 fields
    {
        const char *p = line.data(), *const e = p + line.size();
        const char *t = p;
        for (int k = 0; k < 11; ++k) {
            const char *q = (const char *)memchr(t, '\t', (size_t)(e - t));
            if (k == 0) name_len = (size_t)((q ? q : e) - p);
            if (!q) { if (k < 10) return false; t = e; break; }
            t = q + 1;
        }
        tags_at = (size_t)(t - p);
        return true;
    }
    bool same_name(const LineRec &o) const { return name_len == o.name_len && memcmp(line.data(), o.line.data(), name_len) == 0; }
    // value text of the first optional field whose two-letter name is k (what SamRec::has / SamRec::tag find)
    bool tag(const char *k, std::string &v) const
    {
        for (size_t a = tags_at; a < line.size();) {
            size_t b = line.find('\t', a);
            if (b == std::string::npos) b = line.size();
            if (b - a >= 2 && line[a] == k[0] && line[a + 1] == k[1]) { v = b - a > 5 ? line.substr(a + 5, b - a - 5) : std::string(); return true; }
            a = b + 1;
        }
        return false;
    }
    SamRec split() const
    {
        SamRec r;
        auto f = split_tab(line);
        r.f.assign(f.begin(), f.begin() + 11);
        r.tags.assign(f.begin() + 11, f.end());
        return r;
    }
};

struct Sam {   // header of the stream; the records are pulled one at a time (inputs need not fit in memory)
    std::vector<std::string> header, contigs;
    std::string last_pg;
    samio::LineSource in;
    std::string pending;      // first record line, read while looking for the end of the header
    bool have_pending = false;

    bool open(const std::string &path)
    {
        if (!in.open(path)) return false;
        std::string line;
        while (in.getline(line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            if (line.empty()) continue;
            if (line[0] != '@') { pending.swap(line); have_pending = true; break; }
            header.push_back(line);
            const auto f = split_tab(line);
            if (f[0] == "@SQ") for (auto &x : f) if (x.rfind("SN:", 0) == 0) contigs.push_back(x.substr(3));
            if (f[0] == "@PG") for (auto &x : f) if (x.rfind("ID:", 0) == 0) last_pg = x.substr(3);
        }
        return !in.failed();
    }
    // next record; 0 = end of input, -1 = malformed record or damaged input
    int next(SamRec &r)
    {
        std::string line;
        for (;;) {
            if (have_pending) { line.swap(pending); have_pending = false; }
            else if (!in.getline(line)) return in.failed() ? -1 : 0;
            if (!line.empty() && line.back() == '\r') line.pop_back();
            if (!line.empty()) break;
        }
        auto f = split_tab(line);
        if (f.size() < 11) return -1;
        r.f.assign(f.begin(), f.begin() + 11);
        r.tags.assign(f.begin() + 11, f.end());
        return 1;
    }
    int next(LineRec &r)   // the same, without splitting the line
    {
        for (;;) {
            if (have_pending) { r.line.swap(pending); have_pending = false; }
            else if (!in.getline(r.line)) return in.failed() ? -1 : 0;
            if (!r.line.empty() && r.line.back() == '\r') r.line.pop_back();
            if (!r.line.empty()) break;
        }
        return r.index() ? 1 : -1;
    }
};

void write_header(const Sam &sam, const char *id, const std::string &cl)
{
    for (auto &h : sam.header) out_line(h);
    std::string pg = std::string("@PG\tID:") + id + "\tPN:fade\tVN:" + kVersion;   // filter.d:171-178, remap.d:19-26
    if (!sam.last_pg.empty()) pg += "\tPP:" + sam.last_pg;
    pg += "\tCL:" + cl;
    out_line(pg);
}

typedef std::vector<std::pair<long, char>> Cig;
Cig cig_parse(const std::string &s)
{
    Cig o;
    if (s == "*") return o;
    long n = 0;
    for (char c : s) { if (c >= '0' && c <= '9') n = n * 10 + (c - '0'); else { o.push_back({ n, c }); n = 0; } }
    return o;
}
std::string cig_str(const Cig &c)
{
    std::string o;
    for (auto &x : c) o += std::to_string(x.first) + x.second;
    return o.empty() ? "*" : o;
}
bool q_consuming(char op) { return strchr("MIS=X", op) != nullptr; }
bool r_consuming(char op) { return strchr("MDN=X", op) != nullptr; }
long cig_span(const Cig &c) { long s = 0; for (auto &x : c) if (r_consuming(x.second)) s += x.first; return s; }

std::string am_field(const std::string &am, int side, int idx)
{
    std::vector<std::string> sides;
    size_t a = 0;
    for (;;) { size_t b = am.find(';', a); if (b == std::string::npos) { sides.push_back(am.substr(a)); break; } sides.push_back(am.substr(a, b - a)); a = b + 1; }
    if ((size_t)side >= sides.size()) return "";
    const std::string &s = sides[(size_t)side];
    a = 0;
    for (int k = 0;; ++k) {
        size_t b = s.find(',', a);
        if (k == idx) return s.substr(a, b == std::string::npos ? std::string::npos : b - a);
        if (b == std::string::npos) return "";
        a = b + 1;
    }
}

SamRec blank_record(const SamRec &r, const std::string &seq, const std::string &qual, const Sam &sam)   // U9
{
    SamRec o;
    const bool hc = !sam.contigs.empty();
    o.f = { r.f[0], "0", hc ? sam.contigs[0] : "*", hc ? "1" : "0", "0", "*", hc ? "=" : "*", hc ? "1" : "0", "0", seq, qual };
    return o;
}

// source/filter.d:15-91
SamRec clip_read(const SamRec &rec, int rs, const Sam &sam)
{
    Cig nc = cig_parse(rec.f[5]);
    long pos = atol(rec.f[3].c_str());
    std::string seq = rec.f[9], qual = rec.f[10];
    const std::string am = rec.tag("am");
    if (rs & 2) {
        long to_trim = cig_span(cig_parse(am_field(am, 0, 2)));
        long hard = 0;
        if (to_trim < cig_span(cig_parse(rec.f[5]))) {
            while (to_trim) {
                if (q_consuming(nc[0].second)) { seq.erase(0, 1); qual.erase(0, 1); ++hard; }
                if (r_consuming(nc[0].second)) { ++pos; --to_trim; }
                if (--nc[0].first == 0) nc.erase(nc.begin());
            }
        } else return blank_record(rec, seq, qual, sam);
        nc.insert(nc.begin(), { hard, 'H' });
    }
    if (rs & 4) {
        long to_trim = cig_span(cig_parse(am_field(am, 1, 2)));
        long hard = 0;
        if (to_trim < cig_span(nc)) {
            while (to_trim) {
                if (q_consuming(nc.back().second)) { seq.pop_back(); qual.pop_back(); ++hard; }
                if (r_consuming(nc.back().second)) --to_trim;
                if (--nc.back().first == 0) nc.pop_back();
            }
        } else return blank_record(rec, seq, qual, sam);
        nc.push_back({ hard, 'H' });
    }
    SamRec o = rec;
    o.f[3] = std::to_string(pos); o.f[5] = cig_str(nc); o.f[9] = seq; o.f[10] = qual;
    return o;
}

// source/filter.d:127-165
// numericallyAwareStringComparison of source/filter.d:127-165, on the strings in place: while both have characters
// left -- two non-digits are compared as characters; otherwise each side's leading digits are read as a number (-1 when
// there are none, which consumes nothing) and compared; when everything compared equal the shorter string comes first.
int natural_compare(const char *a, size_t na, const char *b, size_t nb)
{
    auto isd = [](char c) { return c >= '0' && c <= '9'; };
    const char *const ae = a + na, *const be = b + nb;
    while (a < ae && b < be) {
        if (!isd(*a) && !isd(*b)) {
            if (*a == *b) { ++a; ++b; continue; }
            return *a < *b ? -1 : 1;
        }
        auto take = [&](const char *&p, const char *e) {
            if (p == e || !isd(*p)) return -1L;
            unsigned long long v = 0;
            while (p < e && isd(*p)) { if (v < (1ull << 62) / 10) v = v * 10 + (unsigned)(*p - '0'); else v = (1ull << 62); ++p; }
            return (long)v;
        };
        const char *a0 = a, *b0 = b;
        const long ai = take(a, ae), bi = take(b, be);
        if (ai == bi) { if (a == a0 && b == b0) return 0; continue; }
        return ai < bi ? -1 : 1;
    }
    const size_t ra = (size_t)(ae - a), rb = (size_t)(be - b);
    return ra == rb ? 0 : (ra < rb ? -1 : 1);
}

struct OutStats {   // source/stats.d:16-72
    long read_count = 0, clipped = 0, sup = 0, art_sup = 0, art = 0, aln_l = 0, aln_r = 0;
    void parse(int rs)
    {
        const int al = (rs >> 1) & 1, ar = (rs >> 2) & 1, sp = (rs >> 5) & 1;
        clipped += rs & 1; art += al | ar; sup += sp; art_sup += (al | ar) & sp; aln_l += al; aln_r += ar;
    }
    void print() const
    {
        const double n = read_count ? (double)read_count : 0.0 / 0.0;
        fprintf(stderr, "read count:\t%ld\nClipped %%:\t%g\n%% With Supplementary alns:\t%g\nArtifact rate:\t%g\n"
                        "%% With Supplementary alns and artifacts:\t%g\nArtifact rate left only:\t%g\nArtifact rate right only:\t%g\n",
                read_count, clipped / n, sup / n, art / n, art_sup / n, aln_l / n, aln_r / n);
    }
};

int rs_of(const SamRec &r, bool &have)
{
    have = r.has("rs");
    return have ? (atoi(r.tag("rs").c_str()) & 0xff) : 0;
}

int rs_of(const LineRec &r, bool &have)
{
    std::string v;
    have = r.tag("rs", v);
    return have ? (atoi(v.c_str()) & 0xff) : 0;
}

// `out` on BAM input: the records stay binary.  A record that passes through is copied as the bytes it came in as (or
// converted to text for SAM output), only the records `out -c` clips go through the text form (clip_read); the
// decisions -- rs of every record, the first ten names, whole read groups when the input looks name-sorted -- are those
// of the text path below (filter.d:167-269), on which SAM input still runs.
int out_bam_input(FILE *f, const std::string &path, bool clip, int con, int threads, const std::string &cl)
{
    using bamfast::get_u32;
    auto damaged = [&] { fprintf(stderr, "fade-b200: malformed record or damaged input in %s\n", path.c_str()); return 1; };
    bamfast::RecordInput in(f, std::string(), true, threads);
    if (!in.read_header()) return 1;
    Sam sam;                          // what clip_read and the @PG line need of it
    sam.contigs = in.hdr.names;
    for (const auto &h : in.hdr.lines)
        if (h.compare(0, 3, "@PG") == 0) for (const auto &x : split_tab(h)) if (x.rfind("ID:", 0) == 0) sam.last_pg = x.substr(3);
    {
        std::string pg = std::string("@PG\tID:fade-extract\tPN:fade\tVN:") + kVersion;   // sic: filter.d:173 uses the ID of extract
        if (!sam.last_pg.empty()) pg += "\tPP:" + sam.last_pg;
        in.hdr.add_line(pg + "\tCL:" + cl);
    }
    bamfast::write_header(in.hdr, con, threads);
    OutStats st;
    const std::vector<uint8_t> &sb = in.stream;
    auto name_of = [&](size_t off, size_t &len) { const char *nm = reinterpret_cast<const char *>(&sb[off + 36]); len = strnlen(nm, sb[off + 12]); return nm; };
    auto same_name = [&](size_t a, size_t b) { size_t la, lb; const char *na = name_of(a, la), *nb = name_of(b, lb); return la == lb && memcmp(na, nb, la) == 0; };
    bool eof = false, decided = false, sorted = true, unencodable = false;
    std::vector<size_t> offs;
    std::vector<uint8_t> have, keep;
    std::vector<int> rsv;
    std::vector<std::string> part;
    std::string all;
    for (;;) {
        if (!eof) {
            const size_t got = in.stream.size() - in.spos;
            if (!in.need(got + 1)) { eof = true; if (in.bad()) return damaged(); }
        }
        // whole records available
        offs.clear();
        size_t end = in.spos;
        while (end + 4 <= sb.size()) {
            const uint32_t bs = get_u32(&sb[end]);
            if (bs < 32) return damaged();
            if (end + 4 + (size_t)bs > sb.size()) break;
            const uint64_t ln = sb[end + 12], n_cig = bamfast::get_u16(&sb[end + 16]);
            const int64_t l_seq = bamfast::get_i32(&sb[end + 20]);
            if (ln < 1 || l_seq < 0 || 32 + ln + 4 * n_cig + (uint64_t)(l_seq + 1) / 2 + (uint64_t)l_seq > bs) return damaged();
            offs.push_back(end);
            end += 4 + (size_t)bs;
        }
        if (eof && end != sb.size()) return damaged();
        if (!decided) {   // filter.d:209-266: the first ten records decide whether the input is taken as name-sorted
            if (offs.size() < 10 && !eof) continue;
            for (size_t k = 0; k + 1 < offs.size() && k + 1 < 10; ++k) {
                size_t la, lb;
                const char *a = name_of(offs[k], la), *b = name_of(offs[k + 1], lb);
                if (natural_compare(b, lb, a, la) < 0) sorted = false;
            }
            decided = true;
            if (!clip)
                fprintf(stderr, sorted ? "[W::fade-out] Output looks name-sorted, ejecting all reads with same readname if any have an artifact\n"
                                       : "[W::fade-out] Output doesn't look name-sorted, ejecting by only reads with an artifact\n");
        }
        size_t n = offs.size();
        if (n == 0 && !eof) continue;    // only the front of one record so far
        if (!clip && sorted && !eof) {   // the last read group may continue in the part of the file not read yet
            const size_t tail = offs.back();
            while (n > 0 && same_name(offs[n - 1], tail)) --n;
            if (n == 0) continue;
        }
        // rs of every record (SamRec::has / tag: the first optional field named rs, read as a number)
        have.assign(n, 0); keep.assign(n, 0); rsv.assign(n, 0);
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) num_threads(threads)
        for (long k = 0; k < (long)n; ++k) {
            const uint8_t *p = &sb[offs[(size_t)k] + 4], *const e = p + get_u32(&sb[offs[(size_t)k]]);
            const uint8_t *a = p + 32 + p[8] + 4ull * bamfast::get_u16(p + 12) + (size_t)(bamfast::get_i32(p + 16) + 1) / 2 + (size_t)bamfast::get_i32(p + 16);
            while (a < e) {
                const size_t sz = bamfast::aux_field_size(a, e);
                if (!sz) { bad = 1; break; }
                if (a[0] == 'r' && a[1] == 's') {
                    have[(size_t)k] = 1;
                    long long v = 0;
                    switch (a[2]) {
                    case 'c': v = (int8_t)a[3]; break;
                    case 'C': v = a[3]; break;
                    case 's': v = (int16_t)bamfast::get_u16(a + 3); break;
                    case 'S': v = bamfast::get_u16(a + 3); break;
                    case 'i': v = bamfast::get_i32(a + 3); break;
                    case 'I': v = get_u32(a + 3); break;
                    default: {   // any other type: through the text form, like the text path
                        LineRec lr;
                        if (!samio::bam_to_sam(p, (size_t)(e - p), in.hdr, lr.line) || !lr.index()) { bad = 1; break; }
                        std::string t;
                        lr.tag("rs", t);
                        v = atoi(t.c_str());
                    }
                    }
                    rsv[(size_t)k] = (int)(v & 0xff);
                    break;
                }
                a += sz;
            }
        }
        if (bad) return damaged();
        // what happens to each record: 0 dropped, 1 as it is, 2 clipped
        if (clip) {           // filter.d:182-208
            for (size_t k = 0; k < n; ++k) {
                ++st.read_count;
                if (have[k]) st.parse(rsv[k]);
                keep[k] = (have[k] && (rsv[k] & 6)) ? 2 : 1;
            }
        } else if (sorted) {  // whole read groups
            for (size_t a = 0; a < n;) {
                size_t b = a + 1;
                while (b < n && same_name(offs[b], offs[a])) ++b;
                bool art = false;
                for (size_t k = a; k < b; ++k) { ++st.read_count; if (have[k]) { st.parse(rsv[k]); if (rsv[k] & 6) art = true; } }
                for (size_t k = a; k < b; ++k) keep[k] = art ? 0 : 1;
                a = b;
            }
        } else {
            for (size_t k = 0; k < n; ++k) {
                ++st.read_count;
                if (!have[k]) continue;
                st.parse(rsv[k]);
                keep[k] = (rsv[k] & 6) ? 0 : 1;
            }
        }
        // the surviving records, converted side by side and written in order
        const int T = (int)std::max<long>(1, std::min<long>(threads, (long)n / 256));
        part.assign((size_t)T, std::string());
        int enc = 0;
#pragma omp parallel for schedule(static, 1) reduction(| : bad, enc) num_threads(T)
        for (int t = 0; t < T; ++t) {
            std::string &o = part[(size_t)t];
            std::string line, rec;
            for (size_t k = n * (size_t)t / (size_t)T; k < n * (size_t)(t + 1) / (size_t)T; ++k) {
                if (!keep[k]) continue;
                const uint8_t *p = &sb[offs[k]];
                const uint32_t bs = get_u32(p);
                if (keep[k] == 1 && con != 0) { o.append(reinterpret_cast<const char *>(p), 4 + (size_t)bs); continue; }
                LineRec lr;
                if (!samio::bam_to_sam(p + 4, bs, in.hdr, lr.line)) { bad = 1; continue; }
                if (keep[k] == 2) {
                    if (!lr.index()) { bad = 1; continue; }
                    lr.line = clip_read(lr.split(), rsv[k], sam).line();
                }
                if (con == 0) { o += lr.line; o += '\n'; continue; }
                if (!samio::sam_to_bam(lr.line, in.hdr, rec)) { enc = 1; continue; }
                samio::put_u32(o, (uint32_t)rec.size());
                o += rec;
            }
        }
        if (bad) return damaged();
        if (enc) unencodable = true;
        if (con == 0) for (const auto &o : part) fwrite(o.data(), 1, o.size(), stdout);
        else {
            size_t tot = 0;
            for (const auto &o : part) tot += o.size();
            all.clear();
            all.reserve(tot);
            for (const auto &o : part) all += o;
            if (!all.empty()) bamfast::write_blocks(stdout, reinterpret_cast<const uint8_t *>(all.data()), all.size(), con == 1 ? 0 : samio::kFastLevel, threads);
        }
        const size_t consumed = n < offs.size() ? offs[n] : end;
        in.stream.erase(in.stream.begin(), in.stream.begin() + (long)consumed);
        in.spos = 0;
        if (eof && n == offs.size()) break;
    }
    if (con != 0) bamfast::write_eof_marker();
    if (fflush(stdout) != 0 || ferror(stdout)) { fprintf(stderr, "fade-b200: write error\n"); return 1; }
    st.print();
    if (unencodable) { fprintf(stderr, "fade-b200: a clipped record could not be encoded for BAM output\n"); return 1; }
    return 0;
}

int cmd_out(int argc, char **argv, const std::string &cl)
{
    const Options opt = parse_options(argc, argv, 2, { { 'c', "clip", false }, { 't', "threads", true }, { 'h', "help", false },
                                                      { 'b', "bam", false }, { 'u', "ubam", false } });
    if (!opt.ok || opt.pos.size() > 1) { usage(); return 1; }
    if (opt.has("help") || opt.pos.empty()) return usage();
    const bool clip = opt.has("clip");
    const std::string path = opt.pos[0];
    int con = 0;
    if (!output_container(opt, con)) return 1;
    {   // BAM input (a gzip member starts with 0x1f, SAM text never does) takes the binary route
        FILE *f = path == "-" ? stdin : fopen(path.c_str(), "rb");
        if (!f) { fprintf(stderr, "fade-b200: cannot read %s\n", path.c_str()); return 1; }
        const int c0 = fgetc(f);
        if (c0 != EOF) ungetc(c0, f);
        if (c0 == 0x1f) {
            const int threads = (int)opt.num("threads", 0) > 0 ? (int)opt.num("threads", 0) : omp_get_max_threads();
            return out_bam_input(f, path, clip, con, threads, cl);
        }
        if (f != stdin) fclose(f);
    }
    open_output(con);
    Sam sam;
    if (!sam.open(path)) { fprintf(stderr, "fade-b200: cannot read %s\n", path.c_str()); return 1; }
    write_header(sam, "fade-extract", cl);   // sic: filter.d:173 uses the ID of extract
    OutStats st;
    bool put_failed = false;
    auto put = [&put_failed](const std::string &l) { if (!out_line(l)) put_failed = true; };
    int rc = 0;
    LineRec r;
    if (clip) {   // filter.d:182-208
        while ((rc = sam.next(r)) == 1) {
            ++st.read_count;
            bool have;
            const int rs = rs_of(r, have);
            if (!have) { put(r.line); continue; }
            st.parse(rs);
            if (!(rs & 6)) put(r.line); else put(clip_read(r.split(), rs, sam).line());
        }
    } else {      // filter.d:209-266: the first ten records decide whether the input is taken as name-sorted
        std::vector<LineRec> head;
        while (head.size() < 10 && (rc = sam.next(r)) == 1) head.push_back(r);
        bool sorted = true;
        for (size_t k = 0; k + 1 < head.size(); ++k)
            if (natural_compare(head[k + 1].line.data(), head[k + 1].name_len, head[k].line.data(), head[k].name_len) < 0) sorted = false;
        size_t hp = 0;
        auto next = [&](LineRec &o) -> int {   // the buffered head first, then the rest of the stream
            if (hp < head.size()) { o = head[hp++]; return 1; }
            if (rc != 1) return rc;
            return rc = sam.next(o);
        };
        if (sorted) {
            fprintf(stderr, "[W::fade-out] Output looks name-sorted, ejecting all reads with same readname if any have an artifact\n");
            std::vector<LineRec> group;
            size_t ng = 0;        // records of the current group (slots of `group` are reused)
            bool art = false;
            auto flush_group = [&]() {
                if (!art) for (size_t g = 0; g < ng; ++g) put(group[g].line);
                ng = 0;
                art = false;
            };
            while (next(r) == 1) {
                if (ng && !group[ng - 1].same_name(r)) flush_group();
                ++st.read_count;
                bool have;
                const int rs = rs_of(r, have);
                if (have) { st.parse(rs); if (rs & 6) art = true; }
                if (ng == group.size()) group.emplace_back();
                std::swap(group[ng], r);
                ++ng;
            }
            flush_group();
        } else {
            fprintf(stderr, "[W::fade-out] Output doesn't look name-sorted, ejecting by only reads with an artifact\n");
            while (next(r) == 1) {
                ++st.read_count;
                bool have;
                const int rs = rs_of(r, have);
                if (!have) continue;
                st.parse(rs);
                if (!(rs & 6)) put(r.line);
            }
        }
    }
    if (rc < 0) { fprintf(stderr, "fade-b200: malformed record or damaged input in %s\n", path.c_str()); return 1; }
    st.print();
    return (close_output() || put_failed) ? 1 : 0;   // a record that cannot be encoded for the chosen container is an error, not a silent drop
}

// BAM records `extract` has no use for: an integer rs tag without an artifact bit, or no rs tag at all (remap.d:29-33).
// Anything unusual (damaged sizes, an rs tag of another type) is NOT skipped, so the text route below sees and judges it.
bool no_artifact_bits(const uint8_t *p, size_t n)
{
    if (n < 32) return false;
    const uint64_t l_name = p[8], n_cig = bamfast::get_u16(p + 12);
    const int64_t l_seq = bamfast::get_i32(p + 16);
    if (l_seq < 0 || 32 + l_name + 4 * n_cig + (uint64_t)(l_seq + 1) / 2 + (uint64_t)l_seq > n) return false;
    const uint8_t *a = p + 32 + l_name + 4 * n_cig + (uint64_t)(l_seq + 1) / 2 + (uint64_t)l_seq, *const e = p + n;
    while (a < e) {
        const size_t sz = bamfast::aux_field_size(a, e);
        if (!sz) return false;
        if (a[0] == 'r' && a[1] == 's') {
            switch (a[2]) {
            case 'c': case 'C': return !(a[3] & 6);
            case 's': case 'S': case 'i': case 'I': return !(a[3] & 6);   // little endian: the low byte comes first
            default: return false;
            }
        }
        a += sz;
    }
    return true;
}

int cmd_extract(int argc, char **argv, const std::string &cl)
{
    const Options opt = parse_options(argc, argv, 2, { { 't', "threads", true }, { 'h', "help", false }, { 'b', "bam", false },
                                                      { 'u', "ubam", false } });
    if (!opt.ok || opt.pos.size() > 1) { usage(); return 1; }
    if (opt.has("help") || opt.pos.empty()) return usage();
    const std::string path = opt.pos[0];
    int con = 0;
    if (!output_container(opt, con)) return 1;
    open_output(con);
    Sam sam;
    sam.in.skip_records_if(no_artifact_bits);
    if (!sam.open(path)) { fprintf(stderr, "fade-b200: cannot read %s\n", path.c_str()); return 1; }
    write_header(sam, "fade-extract", cl);
    static const char comp[] = "=TGKCYSBAWRDMHVN";   // complement of "=ACMGRSVTWYHKDBN" (util.d:18-21)
    static const char nt16[] = "=ACMGRSVTWYHKDBN";
    SamRec r;
    int rc;
    while ((rc = sam.next(r)) == 1) {   // remap.d:29-85
        bool have;
        const int rs = rs_of(r, have);
        if (!have || !(rs & 6) || !r.has("am")) continue;
        const std::string am = r.tag("am");
        for (int side = 0; side < 2; ++side) {
            if (!(rs & (side == 0 ? 2 : 4))) continue;
            const std::string chrom = am_field(am, side, 0), pos0 = am_field(am, side, 1), cig = am_field(am, side, 2);
            int tid = -1;
            for (size_t t = 0; t < sam.contigs.size(); ++t) if (sam.contigs[t] == chrom) tid = (int)t;
            std::string seq(r.f[9].rbegin(), r.f[9].rend());
            for (auto &ch : seq) { const char *p = strchr(nt16, toupper((unsigned char)ch)); ch = (p && ch) ? comp[p - nt16] : 'N'; }
            const std::string qual(r.f[10].rbegin(), r.f[10].rend());
            const int flag = (atoi(r.f[1].c_str()) & 16) ? 0 : 16;
            const bool hc = !sam.contigs.empty();
            std::string o = r.f[0] + "\t" + std::to_string(flag) + "\t" + (tid >= 0 ? chrom : std::string("*")) + "\t" +
                            std::to_string(atol(pos0.c_str()) + 1) + "\t0\t" + cig + "\t" +
                            (tid == 0 ? std::string("=") : (hc ? sam.contigs[0] : std::string("*"))) + "\t" + (hc ? "1" : "0") +
                            "\t0\t" + seq + "\t" + qual;
            if (!out_line(o)) return 1;
        }
    }
    if (rc < 0) { fprintf(stderr, "fade-b200: malformed record or damaged input in %s\n", path.c_str()); return 1; }
    return close_output();
}

// `fade-b200 sort -n`: name-sorted copy of the input (what config 5 of BASELINE.json uses `samtools sort -n`
// for, between annotate and out).  Order: the natural order `fade out` itself tests for (filter.d:127-165),
// then first-of-pair before second-of-pair (FLAG & 0xC0), then input order.  In memory.
int cmd_sort(int argc, char **argv)
{
    const Options opt = parse_options(argc, argv, 2, { { 'n', "name", false }, { 't', "threads", true }, { 'b', "bam", false },
                                                      { 'u', "ubam", false } });
    if (!opt.ok || opt.pos.size() != 1 || !opt.has("name")) {
        fprintf(stderr, "usage: fade-b200 sort -n [-b|-u] <SAM/BAM or ->   (only name order is implemented)\n");
        return 1;
    }
    const std::string path = opt.pos[0];
    int con = 0;
    if (!output_container(opt, con)) return 1;
    const int threads = (int)opt.num("threads", 0) > 0 ? (int)opt.num("threads", 0) : omp_get_max_threads();
    FILE *f = path == "-" ? stdin : fopen(path.c_str(), "rb");
    if (!f) { fprintf(stderr, "fade-b200: cannot read %s\n", path.c_str()); return 1; }
    std::string pre(2, '\0');
    pre.resize(fread(&pre[0], 1, 2, f));
    const bool is_bam = pre.size() == 2 && (uint8_t)pre[0] == 0x1f && (uint8_t)pre[1] == 0x8b;
    // every record as BAM bytes in one buffer (SAM lines are converted on the way in); only a 16-byte key per record
    // is sorted, and the records are gathered in that order on the way out
    bamfast::RecordInput in(f, pre, is_bam, threads);
    if (!in.read_header()) return 1;
    for (;;) {
        const size_t have = in.stream.size() - in.spos;
        if (!in.need(have + 1)) break;
    }
    if (in.bad()) { fprintf(stderr, "fade-b200: malformed record or damaged input in %s\n", path.c_str()); return 1; }
    struct Key { uint64_t off; uint32_t name; uint8_t len, pair; };
    std::vector<Key> keys;
    std::vector<char> names;
    const std::vector<uint8_t> &st = in.stream;
    for (size_t p2 = in.spos; p2 < st.size();) {
        const size_t left = st.size() - p2;
        const uint32_t bs = left >= 4 ? bamfast::get_u32(&st[p2]) : 0;
        const uint32_t ln = left >= 36 ? st[p2 + 4 + 8] : 0;
        // the fixed part and the sizes it declares must fit the record (the check samio::bam_to_sam starts with)
        const uint64_t n_cig = left >= 36 ? bamfast::get_u16(&st[p2 + 4 + 12]) : 0;
        const int64_t l_seq = left >= 36 ? bamfast::get_i32(&st[p2 + 4 + 16]) : 0;
        if (left < 36 || bs < 32 || (size_t)bs + 4 > left || ln < 1 || l_seq < 0 ||
            32 + ln + 4 * n_cig + (uint64_t)(l_seq + 1) / 2 + (uint64_t)l_seq > bs || names.size() > 0xffffff00u) {
            fprintf(stderr, "fade-b200: malformed record or damaged input in %s\n", path.c_str());
            return 1;
        }
        const char *nm = reinterpret_cast<const char *>(&st[p2 + 36]);
        const size_t nl = strnlen(nm, ln);
        keys.push_back({ (uint64_t)p2, (uint32_t)names.size(), (uint8_t)nl, (uint8_t)(bamfast::get_u16(&st[p2 + 4 + 14]) & 0xc0) });
        names.insert(names.end(), nm, nm + nl);
        p2 += 4 + (size_t)bs;
    }
    // name, then first / second of the pair, input order among equals
    omp_set_num_threads(threads);
    __gnu_parallel::stable_sort(keys.begin(), keys.end(), [&](const Key &x, const Key &y) {
        const int c = natural_compare(&names[x.name], x.len, &names[y.name], y.len);
        if (c) return c < 0;
        return x.pair < y.pair;
    });
    bool have_hd = false;
    for (auto &h : in.hdr.lines) {   // @HD SO:queryname, as samtools sort -n writes it
        if (h.compare(0, 3, "@HD") != 0) continue;
        have_hd = true;
        const size_t a = h.find("\tSO:");
        if (a == std::string::npos) h += "\tSO:queryname";
        else { const size_t e = h.find('\t', a + 1); h.replace(a, (e == std::string::npos ? h.size() : e) - a, "\tSO:queryname"); }
    }
    if (!have_hd) in.hdr.lines.insert(in.hdr.lines.begin(), "@HD\tVN:1.6\tSO:queryname");
    bamfast::write_header(in.hdr, con, threads);
    constexpr size_t kBatch = 1 << 17;   // records gathered (BAM) or converted to text (SAM) side by side per write
    std::vector<uint8_t> gathered;
    std::vector<size_t> at(kBatch + 1);
    std::vector<std::string> lines(con == 0 ? kBatch : 0);
    for (size_t k0 = 0; k0 < keys.size(); k0 += kBatch) {
        const size_t nk = std::min(kBatch, keys.size() - k0);
        if (con != 0) {
            at[0] = 0;
            for (size_t k = 0; k < nk; ++k) at[k + 1] = at[k] + 4 + bamfast::get_u32(&st[keys[k0 + k].off]);
            gathered.resize(at[nk]);
#pragma omp parallel for schedule(static) num_threads(threads)
            for (long k = 0; k < (long)nk; ++k) memcpy(&gathered[at[(size_t)k]], &st[keys[k0 + (size_t)k].off], at[(size_t)k + 1] - at[(size_t)k]);
            bamfast::write_blocks(stdout, gathered.data(), gathered.size(), con == 1 ? 0 : samio::kFastLevel, threads);
        } else {
            int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) num_threads(threads)
            for (long k = 0; k < (long)nk; ++k) {
                const size_t o = (size_t)keys[k0 + (size_t)k].off;
                if (!samio::bam_to_sam(&st[o + 4], bamfast::get_u32(&st[o]), in.hdr, lines[(size_t)k])) bad = 1;
                else lines[(size_t)k].push_back('\n');
            }
            if (bad) { fprintf(stderr, "fade-b200: malformed record or damaged input in %s\n", path.c_str()); return 1; }
            for (size_t k = 0; k < nk; ++k) fwrite(lines[k].data(), 1, lines[k].size(), stdout);
        }
    }
    if (con != 0) bamfast::write_eof_marker();
    if (fflush(stdout) != 0 || ferror(stdout)) { fprintf(stderr, "fade-b200: write error\n"); return 1; }
    return 0;
}

// `fade-b200 fasta-digest <FASTA>`: name, length and FNV-1a hash of every contig as the loader sees it
// (with or without a .fai); a check of the loader that needs no GPU
int cmd_fasta_digest(int argc, char **argv)
{
    if (argc < 3) return usage();
    std::map<std::string, std::string> fasta;
    if (!read_fasta(argv[2], fasta)) { fprintf(stderr, "fade-b200: cannot read %s\n", argv[2]); return 1; }
    for (const auto &kv : fasta) {
        uint64_t h = 1469598103934665603ull;
        for (unsigned char ch : kv.second) { h ^= ch; h *= 1099511628211ull; }
        printf("%s\t%zu\t%016llx\n", kv.first.c_str(), kv.second.size(), (unsigned long long)h);
    }
    return 0;
}

// `fade-b200 bgzf [--level fast|0..9] <file or ->`: the bytes of the file as a BGZF stream (the block compressor of the
// BAM writers on arbitrary data; used by the tests to check the built-in DEFLATE encoder against zlib's inflate)
int cmd_bgzf(int argc, char **argv)
{
    const Options opt = parse_options(argc, argv, 2, { { 0, "level", true }, { 't', "threads", true } });
    if (!opt.ok || opt.pos.size() != 1) { usage(); return 1; }
    const int level = (!opt.has("level") || opt.val.at("level") == "fast") ? samio::kFastLevel : (int)opt.num("level", 6);
    if (level < samio::kFastLevel || level > 9) { fprintf(stderr, "fade-b200: bad --level\n"); return 1; }
    FILE *f = opt.pos[0] == "-" ? stdin : fopen(opt.pos[0].c_str(), "rb");
    if (!f) { fprintf(stderr, "fade-b200: cannot read %s\n", opt.pos[0].c_str()); return 1; }
    samio::BgzfWriter bz(stdout, level);
    std::vector<uint8_t> buf(1 << 20);
    size_t got;
    while ((got = fread(buf.data(), 1, buf.size(), f)) > 0) bz.write(buf.data(), got);
    bz.finish();
    return 0;
}

// format conversion only: every header line and record, unchanged (SAM <-> BAM)
int cmd_view(int argc, char **argv)
{
    const Options opt = parse_options(argc, argv, 2, { { 0, "bulk", false }, { 0, "count", false }, { 't', "threads", true }, { 'h', "help", false },
                                                      { 'b', "bam", false }, { 'u', "ubam", false } });
    if (!opt.ok || opt.pos.size() > 1) { usage(); return 1; }
    if (opt.has("help") || opt.pos.empty()) return usage();
    const std::string path = opt.pos[0];
    const int threads = (int)opt.num("threads", 0);
    const bool bulk = opt.has("bulk");      // through the parallel readers / writers of the annotate loop
    int con = 0;
    if (!output_container(opt, con)) return 1;
    {   // BAM input always takes the parallel route (--bulk asks for it on SAM text as well)
        FILE *f = path == "-" ? stdin : fopen(path.c_str(), "rb");
        if (!f) { fprintf(stderr, "fade-b200: cannot read %s\n", path.c_str()); return 1; }
        const int c0 = fgetc(f);
        if (c0 != EOF) ungetc(c0, f);      // a gzip member starts with 0x1f, SAM text never does
        if (c0 == 0x1f || bulk || opt.has("count")) {
            std::string pre(2, '\0');
            pre.resize(fread(&pre[0], 1, 2, f));
            const bool is_bam = pre.size() == 2 && (uint8_t)pre[0] == 0x1f && (uint8_t)pre[1] == 0x8b;
            if (opt.has("count")) return bamfast::count_records(f, pre, is_bam, threads > 0 ? threads : omp_get_max_threads());
            return bamfast::copy_records(f, pre, is_bam, con, threads > 0 ? threads : omp_get_max_threads());
        }
        if (f != stdin) fclose(f);
    }
    open_output(con);
    samio::LineSource in;
    if (!in.open(path)) { fprintf(stderr, "fade-b200: cannot read %s\n", path.c_str()); return 1; }
    std::string line;
    while (in.getline(line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty() && !out_line(line)) return 1;
    }
    if (in.failed()) { fprintf(stderr, "fade-b200: damaged BAM input\n"); return 1; }
    return close_output();
}

}  // namespace

int main(int argc, char **argv)
{
    std::string cl;
    for (int i = 0; i < argc; ++i) { if (i) cl += " "; cl += argv[i]; }
    if (argc < 2) { usage(); return 0; }
    if (strcmp(argv[1], "annotate") == 0) return cmd_annotate(argc, argv, cl);
    if (strcmp(argv[1], "out") == 0) return cmd_out(argc, argv, cl);
    if (strcmp(argv[1], "extract") == 0) return cmd_extract(argc, argv, cl);
    if (strcmp(argv[1], "view") == 0) return cmd_view(argc, argv);
    if (strcmp(argv[1], "sort") == 0) return cmd_sort(argc, argv);
    if (strcmp(argv[1], "fasta-digest") == 0) return cmd_fasta_digest(argc, argv);
    if (strcmp(argv[1], "bgzf") == 0) return cmd_bgzf(argc, argv);
    usage();
    return 1;
}
