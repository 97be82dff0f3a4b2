// fadesim.cpp -- deterministic synthetic inputs for tests and bench.py:
// a random reference with N runs / lower-case runs and simulated, already-aligned 2xL paired
// reads with injected inverted-repeat ("fragmentase-style") soft clips.  Measurement / test
// infrastructure: not part of the hot path and not an oracle (it produces inputs only).
//
// RNG: SplitMix64 -> xoshiro256**; read k draws from a stream derived from (read_seed, k) and its
// fragment from (read_seed, k/2), so any sharding of the read range yields identical reads.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>
#include <zlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x)
    {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed, uint64_t stream)
    {
        uint64_t x = seed * 0xD1342543DE82EF95ull + stream * 0x9E3779B97F4A7C15ull + 0x2545F4914F6CDD1Dull;
        for (auto &v : s) v = splitmix(x);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next()
    {
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    int64_t below(int64_t n) { return n <= 0 ? 0 : (int64_t)(uni() * (double)n); }
    int64_t range(int64_t lo, int64_t hi) { return lo + below(hi - lo + 1); }  // inclusive
    double normal()
    {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

const char kBase[4] = { 'A', 'C', 'G', 'T' };
inline char comp(char c)
{
    switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    default: return 'N';
    }
}
inline int nt16(char c)
{
    switch (c) {
    case 'A': case 'a': return 1; case 'C': case 'c': return 2; case 'G': case 'g': return 4;
    case 'T': case 't': return 8; default: return 15;
    }
}
inline char up(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }

}  // namespace

extern "C" {

// host threads used by the generators (0 = leave the OpenMP default; torchrun sets OMP_NUM_THREADS=1)
void fadesim_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// Fill `out` (len bytes, no NUL) with a random contig.  n_run_every > 0 plants one N run of
// n_run_len bases in every n_run_every bases (at a random offset); lower_frac of the 1-kb
// tiles are lower-cased (soft-masking must be a no-op after upper-casing, analysis.d:63).
void fadesim_contig(uint64_t ref_seed, int32_t contig_index, int64_t len, int64_t n_run_every,
                    int64_t n_run_len, double lower_frac, char *out)
{
    const int64_t CH = 1 << 20;
    const int64_t nch = (len + CH - 1) / CH;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t c = 0; c < nch; ++c) {
        Rng r(ref_seed, ((uint64_t)contig_index << 40) + (uint64_t)c);
        const int64_t a = c * CH, b = std::min(len, a + CH);
        for (int64_t p = a; p < b; p += 32) {
            const uint64_t w = r.next();
            const int lim = (int)std::min<int64_t>(32, b - p);
            for (int k = 0; k < lim; ++k) out[p + k] = kBase[(w >> (2 * k)) & 3];
        }
        for (int64_t p = a; p < b; p += 1024)
            if (lower_frac > 0 && r.uni() < lower_frac)
                for (int64_t k = p; k < std::min(b, p + 1024); ++k) out[k] = (char)(out[k] + 32);
    }
    if (n_run_every > 0 && n_run_len > 0) {
        Rng r(ref_seed ^ 0xABCDEFull, 0x7777ull + (uint64_t)contig_index);
        for (int64_t a = 0; a + n_run_every <= len; a += n_run_every) {
            const int64_t o = a + r.below(n_run_every - n_run_len);
            memset(out + o, 'N', (size_t)n_run_len);
        }
    }
}

struct fadesim_cfg {
    uint64_t read_seed;
    int32_t read_len;        // L
    int32_t window;          // W used to place in-window artifacts
    int32_t clip_min_art;    // 6
    int32_t clip_max;        // 60
    double frag_mean, frag_sd;
    int32_t frag_max;        // 800
    double p_artifact;       // 0.10
    double p_random_clip;    // 0.05
    double p_both;           // 0.02
    double p_indel;          // 0.01
    double p_unmapped;       // 0.005
    double p_sa;             // 0.02
    double p_outside;        // 0.10 of artifacts come from outside the window
    double sub_rate;         // 0.005
    int32_t short_clip_law;  // 0: as above; 1: all clip lengths U{1..40} (stress config C4)
};

void fadesim_default_cfg(fadesim_cfg *c)
{
    c->read_seed = 2001; c->read_len = 150; c->window = 300; c->clip_min_art = 6; c->clip_max = 60;
    c->frag_mean = 400; c->frag_sd = 50; c->frag_max = 800;
    c->p_artifact = 0.10; c->p_random_clip = 0.05; c->p_both = 0.02; c->p_indel = 0.01;
    c->p_unmapped = 0.005; c->p_sa = 0.02; c->p_outside = 0.10; c->sub_rate = 0.005; c->short_clip_law = 0;
}

// Simulate reads [first, first+n).  All outputs are optional (may be NULL) except l_qseq.
//   seq4: n * ceil(L/2) bytes (fixed stride), seq_off: n+1, qual: n*L bytes,
//   cigar: n*6 words, n_cigar: n, flag/tid/pos/aligned_len/clip_left/clip_right/has_sa/truth: n.
// truth: bit0 = left clip is an in-window artifact, bit1 = right clip is one.
void fadesim_reads(const fadesim_cfg *cfg, int64_t first, int64_t n, int32_t n_contigs,
                   const char *const *contigs, const int64_t *contig_len,
                   uint8_t *seq4, int64_t *seq_off, int32_t *l_qseq, uint8_t *qual,
                   uint32_t *cigar, int32_t *n_cigar, int32_t *flag, int32_t *tid, int64_t *pos,
                   int32_t *aligned_len, int32_t *clip_left, int32_t *clip_right, uint8_t *has_sa,
                   uint8_t *truth)
{
    const int L = cfg->read_len;
    const int stride = (L + 1) / 2;
    int64_t total = 0;
    for (int t = 0; t < n_contigs; ++t) total += contig_len[t];
    if (seq_off) for (int64_t k = 0; k <= n; ++k) seq_off[k] = k * stride;
#pragma omp parallel for schedule(static)
    for (int64_t kk = 0; kk < n; ++kk) {
        const int64_t k = first + kk;
        Rng pr(cfg->read_seed, (uint64_t)(k >> 1) * 2 + 1);   // fragment stream (shared by the pair)
        Rng rr(cfg->read_seed ^ 0x5bd1e995ull, (uint64_t)k * 2);  // per-read stream
        // fragment
        int64_t g = pr.below(total);
        int t = 0;
        while (t + 1 < n_contigs && g >= contig_len[t]) { g -= contig_len[t]; ++t; }
        const int64_t clen = contig_len[t];
        const char *ref = contigs[t];
        int flen = (int)std::llround(cfg->frag_mean + cfg->frag_sd * pr.normal());
        flen = std::max(L + 20, std::min(cfg->frag_max, flen));
        if (flen > clen) flen = (int)clen;
        int64_t fstart = std::min<int64_t>(g, std::max<int64_t>(0, clen - flen));
        const bool second = (k & 1) != 0;
        int64_t p0 = second ? fstart + flen - L : fstart;   // leftmost reference base covered
        if (p0 < 0) p0 = 0;
        if (p0 + L + 8 > clen) p0 = std::max<int64_t>(0, clen - L - 8);
        int fl = 1 | (second ? (128 | 16) : (64 | 32));
        // events
        const bool unmapped = rr.uni() < cfg->p_unmapped;
        const double ev = rr.uni();
        int cl = 0, cr = 0;            // clip lengths
        int art_l = 0, art_r = 0;      // artifact clip on that side?
        auto clip_len_art = [&]() { return cfg->short_clip_law ? (int)rr.range(1, 40) : (int)rr.range(cfg->clip_min_art, cfg->clip_max); };
        auto clip_len_rnd = [&]() { return cfg->short_clip_law ? (int)rr.range(1, 40) : (int)rr.range(1, cfg->clip_max); };
        if (ev < cfg->p_artifact) {
            if (rr.uni() < 0.5) { cl = clip_len_art(); art_l = 1; } else { cr = clip_len_art(); art_r = 1; }
        } else if (ev < cfg->p_artifact + cfg->p_random_clip) {
            if (rr.uni() < 0.5) cl = clip_len_rnd(); else cr = clip_len_rnd();
        } else if (ev < cfg->p_artifact + cfg->p_random_clip + cfg->p_both) {
            cl = clip_len_art(); cr = clip_len_rnd();
            if (rr.uni() < 0.5) art_l = 1; else { art_r = 1; std::swap(cl, cr); }
        }
        const bool indel = rr.uni() < cfg->p_indel;
        const bool sa = rr.uni() < cfg->p_sa;
        // aligned part
        char buf[1024];
        uint32_t cg[6];
        int ncg = 0;
        const int m = L - cl - cr;     // read bases in the aligned part
        const int64_t apos = p0 + cl;  // rec.pos
        int64_t A = m;
        {
            int w = cl;
            if (indel && m > 20) {
                const int d = (int)rr.range(1, 3);
                const int a = (int)rr.range(5, m - 10);
                if (rr.uni() < 0.5) {  // insertion: d random read bases
                    for (int i = 0; i < a; ++i) buf[w++] = up(ref[apos + i]);
                    for (int i = 0; i < d; ++i) buf[w++] = kBase[rr.below(4)];
                    const int b = m - a - d;
                    for (int i = 0; i < b; ++i) buf[w++] = up(ref[apos + a + i]);
                    A = a + b;
                    if (cl) cg[ncg++] = ((uint32_t)cl << 4) | 4;
                    cg[ncg++] = ((uint32_t)a << 4) | 0; cg[ncg++] = ((uint32_t)d << 4) | 1; cg[ncg++] = ((uint32_t)b << 4) | 0;
                } else {               // deletion: skip d reference bases
                    for (int i = 0; i < a; ++i) buf[w++] = up(ref[apos + i]);
                    const int b = m - a;
                    for (int i = 0; i < b; ++i) buf[w++] = up(ref[std::min<int64_t>(clen - 1, apos + a + d + i)]);
                    A = m + d;
                    if (cl) cg[ncg++] = ((uint32_t)cl << 4) | 4;
                    cg[ncg++] = ((uint32_t)a << 4) | 0; cg[ncg++] = ((uint32_t)d << 4) | 2; cg[ncg++] = ((uint32_t)b << 4) | 0;
                }
            } else {
                for (int i = 0; i < m; ++i) buf[w++] = up(ref[apos + i]);
                if (cl) cg[ncg++] = ((uint32_t)cl << 4) | 4;
                cg[ncg++] = ((uint32_t)m << 4) | 0;
            }
            if (cr) cg[ncg++] = ((uint32_t)cr << 4) | 4;
        }
        // clips
        uint8_t tr = 0;
        for (int side = 0; side < 2; ++side) {
            const int len = side == 0 ? cl : cr;
            if (!len) continue;
            char *dst = side == 0 ? buf : buf + (L - cr);
            const bool art = side == 0 ? art_l : art_r;
            if (art) {
                const int64_t wlo = std::max<int64_t>(0, apos - cfg->window);
                const int64_t whi = std::min<int64_t>(clen, apos + A + cfg->window);
                int64_t o;
                bool inside = true;
                if (rr.uni() < cfg->p_outside) {
                    inside = false;
                    const int64_t dist = rr.range(50, 500);
                    o = rr.uni() < 0.5 ? wlo - dist - len : whi + dist;
                    if (o < 0 || o + len > clen) { o = std::max<int64_t>(0, std::min<int64_t>(clen - len, o)); inside = (o >= wlo && o + len <= whi); }
                } else {
                    o = wlo + rr.below(std::max<int64_t>(1, whi - wlo - len + 1));
                }
                for (int i = 0; i < len; ++i) dst[i] = up(comp(ref[o + len - 1 - i]));   // reverse complement
                if (inside) tr |= (uint8_t)(1u << side);
            } else {
                for (int i = 0; i < len; ++i) dst[i] = kBase[rr.below(4)];
            }
        }
        // substitution errors
        for (int i = 0; i < L; ++i)
            if (rr.uni() < cfg->sub_rate) buf[i] = kBase[rr.below(4)];
        if (unmapped) { fl = (fl & ~16) | 4; ncg = 0; cl = cr = 0; A = 0; tr = 0; }
        // outputs
        l_qseq[kk] = L;
        if (seq4) {
            uint8_t *o = seq4 + kk * stride;
            for (int i = 0; i < stride; ++i) {
                const int hi = nt16(buf[2 * i]), lo = (2 * i + 1 < L) ? nt16(buf[2 * i + 1]) : 0;
                o[i] = (uint8_t)((hi << 4) | lo);
            }
        }
        if (qual) for (int i = 0; i < L; ++i) qual[kk * L + i] = (uint8_t)rr.range(2, 40);
        if (cigar) { for (int i = 0; i < 6; ++i) cigar[kk * 6 + i] = i < ncg ? cg[i] : 0; }
        if (n_cigar) n_cigar[kk] = ncg;
        if (flag) flag[kk] = fl;
        if (tid) tid[kk] = t;
        if (pos) pos[kk] = apos;
        if (aligned_len) aligned_len[kk] = (int32_t)A;
        if (clip_left) clip_left[kk] = cl;
        if (clip_right) clip_right[kk] = cr;
        if (has_sa) has_sa[kk] = sa ? 1 : 0;
        if (truth) truth[kk] = tr;
    }
}


// ---- the simulated records as a BAM file (bench.py's e2e_file leg, the C5 chain at scale, tools/) --------------
// Same records as tests/samio.py:write_sam writes as text: QNAME r<name_base+k>, MAPQ 60, RNEXT *, NM:i:0 and, for
// has_sa reads, the SA:Z string of that writer.  BGZF blocks are deflated side by side (zlib `level`).
namespace {

void put32(std::string &o, uint32_t v) { for (int i = 0; i < 4; ++i) o.push_back((char)(v >> (8 * i))); }
void put16(std::string &o, uint32_t v) { o.push_back((char)(v & 0xff)); o.push_back((char)((v >> 8) & 0xff)); }

int reg2bin(int64_t beg, int64_t end)   // SAMv1 5.3
{
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

bool write_bgzf(FILE *f, const uint8_t *data, size_t n, int level)
{
    constexpr size_t kBlock = 0xff00;
    const long nb = (long)((n + kBlock - 1) / kBlock);
    constexpr long kGroup = 4096;
    for (long g0 = 0; g0 < nb; g0 += kGroup) {
        const long g1 = std::min(nb, g0 + kGroup);
        std::vector<std::string> out((size_t)(g1 - g0));
#pragma omp parallel for schedule(dynamic, 8)
        for (long k = g0; k < g1; ++k) {
            const uint8_t *src = data + (size_t)k * kBlock;
            const size_t len = std::min(kBlock, n - (size_t)k * kBlock);
            std::string &o = out[(size_t)(k - g0)];
            o.resize(0x10000 + 64);
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
            zs.next_in = const_cast<uint8_t *>(src); zs.avail_in = (uInt)len;
            zs.next_out = reinterpret_cast<uint8_t *>(&o[18]); zs.avail_out = (uInt)(o.size() - 26);
            deflate(&zs, Z_FINISH);
            const size_t clen = zs.total_out;
            deflateEnd(&zs);
            static const uint8_t head[16] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0 };
            memcpy(&o[0], head, 16);
            const uint32_t bsize = (uint32_t)(clen + 25);
            o[16] = (char)(bsize & 0xff); o[17] = (char)(bsize >> 8);
            const uint32_t crc = (uint32_t)crc32(crc32(0, nullptr, 0), src, (uInt)len);
            for (int i = 0; i < 4; ++i) { o[18 + clen + (size_t)i] = (char)(crc >> (8 * i)); o[22 + clen + (size_t)i] = (char)((uint32_t)len >> (8 * i)); }
            o.resize(26 + clen);
        }
        for (const auto &o : out) if (fwrite(o.data(), 1, o.size(), f) != o.size()) return false;
    }
    return true;
}

}  // namespace

// returns 0, or -1 when the file cannot be written
int fadesim_write_bam(const char *path, int32_t n_contigs, const char *const *names, const int64_t *lens, int64_t n,
                      int32_t read_len, int64_t name_base, const uint8_t *seq4, const uint8_t *qual, const uint32_t *cigar,
                      const int32_t *n_cigar, const int32_t *flag, const int32_t *tid, const int64_t *pos,
                      const int32_t *aligned_len, const uint8_t *has_sa, int32_t level)
{
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    std::string text = "@HD\tVN:1.6\tSO:unsorted\n";
    for (int t = 0; t < n_contigs; ++t) text += std::string("@SQ\tSN:") + names[t] + "\tLN:" + std::to_string(lens[t]) + "\n";
    text += "@PG\tID:simulator\tPN:fadesim\n";
    std::string hdr("BAM\1", 4);
    put32(hdr, (uint32_t)text.size());
    hdr += text;
    put32(hdr, (uint32_t)n_contigs);
    for (int t = 0; t < n_contigs; ++t) {
        put32(hdr, (uint32_t)strlen(names[t]) + 1);
        hdr += names[t]; hdr.push_back('\0');
        put32(hdr, (uint32_t)lens[t]);
    }
    bool ok = write_bgzf(f, reinterpret_cast<const uint8_t *>(hdr.data()), hdr.size(), level);
    const std::string sa = std::string("SAZ") + names[0] + ",1,+,50M100S,60,0;";
    const int64_t stride = (read_len + 1) / 2;
    constexpr int64_t kChunk = 1 << 18;                     // records encoded (in parallel) and written per round
    std::vector<std::string> part;
    for (int64_t c0 = 0; c0 < n && ok; c0 += kChunk) {
        const int64_t c1 = std::min(n, c0 + kChunk);
        int T = 1;
#ifdef _OPENMP
        T = std::max(1, omp_get_max_threads());
#endif
        part.assign((size_t)T, std::string());
#pragma omp parallel for schedule(static, 1) num_threads(T)
        for (int t = 0; t < T; ++t) {
            std::string &o = part[(size_t)t];
            char nm[32];
            for (int64_t k = c0 + (c1 - c0) * t / T; k < c0 + (c1 - c0) * (t + 1) / T; ++k) {
                const int ln = snprintf(nm, sizeof(nm), "r%lld", (long long)(name_base + k));
                const int nc = n_cigar[k];
                const size_t bs = 32 + (size_t)ln + 1 + 4u * (size_t)nc + (size_t)stride + (size_t)read_len + 4 + (has_sa[k] ? sa.size() + 1 : 0);
                put32(o, (uint32_t)bs);
                put32(o, (uint32_t)tid[k]);
                put32(o, (uint32_t)pos[k]);
                o.push_back((char)(ln + 1));
                o.push_back((char)60);
                const int64_t end = pos[k] + std::max<int64_t>(1, nc ? aligned_len[k] : 1);
                put16(o, (uint32_t)reg2bin(pos[k], end));
                put16(o, (uint32_t)nc);
                put16(o, (uint32_t)flag[k]);
                put32(o, (uint32_t)read_len);
                put32(o, 0xffffffffu); put32(o, 0xffffffffu); put32(o, 0);
                o.append(nm, (size_t)ln + 1);
                o.append(reinterpret_cast<const char *>(cigar + 6 * k), 4u * (size_t)nc);
                o.append(reinterpret_cast<const char *>(seq4 + stride * k), (size_t)stride);
                o.append(reinterpret_cast<const char *>(qual + (int64_t)read_len * k), (size_t)read_len);
                o.append("NMC", 3); o.push_back('\0');
                if (has_sa[k]) { o += sa; o.push_back('\0'); }
            }
        }
        std::string all;
        size_t tot = 0;
        for (const auto &p : part) tot += p.size();
        all.reserve(tot);
        for (const auto &p : part) all += p;
        ok = write_bgzf(f, reinterpret_cast<const uint8_t *>(all.data()), all.size(), level);
    }
    static const uint8_t eof[28] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    ok = ok && fwrite(eof, 1, sizeof(eof), f) == sizeof(eof);
    ok = (fclose(f) == 0) && ok;
    return ok ? 0 : -1;
}

// FASTA text of the contigs, `width` bases per line (what tests/samio.py:write_fasta writes)
int fadesim_write_fasta(const char *path, int32_t n_contigs, const char *const *names, const int64_t *lens,
                        const uint8_t *const *seqs, int32_t width)
{
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    bool ok = true;
    std::string buf;
    for (int t = 0; t < n_contigs && ok; ++t) {
        fprintf(f, ">%s synthetic\n", names[t]);
        const int64_t len = lens[t];
        constexpr int64_t kPiece = 64 << 20;
        for (int64_t a = 0; a < len && ok; a += kPiece - kPiece % width) {
            const int64_t e = std::min(len, a + (kPiece - kPiece % width));
            buf.clear();
            buf.reserve((size_t)(e - a) + (size_t)((e - a) / width) + 2);
            for (int64_t p = a; p < e; p += width) {
                const int64_t q = std::min(e, p + width);
                buf.append(reinterpret_cast<const char *>(seqs[t] + p), (size_t)(q - p));
                buf.push_back('\n');
            }
            ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
        }
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? 0 : -1;
}

}  // extern "C"
