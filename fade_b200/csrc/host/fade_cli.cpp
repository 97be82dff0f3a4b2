// fade_cli.cpp -- C++ host driver `fade-b200 annotate`: the batched mirror of fade's annotate()
// (source/anno.d:16-52) on top of the C ABI (include/fadegpu.h, include/fadehost.h).
//
// The reference host is D + dhtslib/htslib; neither is available here, so this harness speaks SAM
// text only (the reference's default output container, util.d:65-76 case 0) and a plain FASTA.
// It keeps the reference's flags and tag schema:
//     fade-b200 annotate [-t N] [--min-length N] [-w N | --window-size N] <in.sam|-> <ref.fa>  > out.sam
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

namespace {

const char *kVersion = "fade-b200-0.1";

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

bool read_fasta(const std::string &path, std::map<std::string, std::string> &seqs)
{
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
            "fade-b200 annotate: marks artifact reads in bam tags (B200 implementation of `fade annotate`)\n"
            "usage: fade-b200 annotate [options] <SAM or -> <FASTA>   (SAM text in, SAM text out)\n"
            "  -t, --threads N      host threads (default: all cores)\n"
            "      --min-length N   minimum soft-clip length considered (default 5)\n"
            "  -w, --window-size N  bases considered outside of the read region (default 300)\n"
            "      --batch N        records per GPU batch (default 1048576)\n"
            "      --device N       CUDA device (default 0)\n");
    return 0;
}

}  // namespace

int main(int argc, char **argv)
{
    std::string cl;
    for (int i = 0; i < argc; ++i) { if (i) cl += " "; cl += argv[i]; }
    if (argc < 2 || strcmp(argv[1], "annotate") != 0) { usage(); return argc < 2 ? 0 : 1; }
    fadegpu_params prm;
    fadegpu_default_params(&prm);
    int64_t batch_n = 1 << 20;
    int device = 0;
    std::vector<std::string> pos_args;
    for (int i = 2; i < argc; ++i) {
        const std::string a = argv[i];
        auto need = [&](const char *what) -> const char * {
            if (i + 1 >= argc) { fprintf(stderr, "fade-b200: %s needs a value\n", what); exit(1); }
            return argv[++i];
        };
        if (a == "-t" || a == "--threads") prm.host_threads = atoi(need("--threads"));
        else if (a == "--min-length") prm.min_length = atoi(need("--min-length"));
        else if (a == "-w" || a == "--window-size") prm.window_size = atoi(need("--window-size"));
        else if (a == "--batch") batch_n = atoll(need("--batch"));
        else if (a == "--device") device = atoi(need("--device"));
        else if (a == "-h" || a == "--help") return usage();
        else if (a == "-b" || a == "--bam" || a == "-u" || a == "--ubam") {
            fprintf(stderr, "fade-b200: BAM output needs htslib, which this harness does not link; SAM text only\n");
            return 1;
        } else pos_args.push_back(a);
    }
    if (pos_args.size() < 2) { usage(); return 0; }
    fprintf(stderr, "[W::fade annotate] Output SAM will keep the input order\n");

    // ---- header ----
    std::istream *in = &std::cin;
    std::ifstream fin;
    if (pos_args[0] != "-") {
        fin.open(pos_args[0]);
        if (!fin) { fprintf(stderr, "fade-b200: cannot open %s\n", pos_args[0].c_str()); return 1; }
        in = &fin;
    }
    std::vector<std::string> header;
    std::vector<std::string> sq_names;
    std::vector<int64_t> sq_len;
    std::string last_pg_id, line;
    bool have_line = false;
    while (std::getline(*in, line)) {
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
    for (const auto &h : header) { fputs(h.c_str(), stdout); fputc('\n', stdout); }

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
            fputs(r.line.c_str(), stdout);
            fprintf(stdout, "\trs:i:%u", (unsigned)rs);                                   // anno.d:94
            if (rc == 1)                                                                  // anno.d:98-107
                fprintf(stdout, "\tam:Z:%s\tas:Z:%s\tar:Z:%s\tab:Z:%s", am.c_str(), as_.c_str(), ar.c_str(), ab.c_str());
            fputc('\n', stdout);
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
            seq_bytes += (int64_t)r.seq4.size();
            recs.push_back(std::move(r));
            if ((int64_t)recs.size() == batch_n || seq_bytes + 1024 > max_seq) { if (flush()) return 1; seq_bytes = 0; }
        }
        have_line = (bool)std::getline(*in, line);
    }
    if (flush()) return 1;
    fprintf(stderr, "[fade-b200 annotate] %lld records, %lld soft-clipped, %lld with artifact tags\n", (long long)n_total,
            (long long)n_sc, (long long)n_art);
    fadegpu_free_batch(bt);
    fadegpu_destroy(ctx);
    return 0;
}
