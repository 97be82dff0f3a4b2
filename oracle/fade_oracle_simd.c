/*
 * fade_oracle_simd.c -- AVX2 (16 x int16 lanes) / AVX-512BW (32 lanes), one alignment per lane: CPU implementation
 * of the same path (the widest the host supports is chosen at run time; FADE_ORACLE_SIMD=avx2 forces 16 lanes), used ONLY as the reported CPU baseline of bench.py (cpu_baseline / --impl reference)
 * and validated against the scalar oracle in tests/.  TEST / MEASUREMENT INFRASTRUCTURE, not
 * product code.  PARITY UNPINNED like the scalar oracle it mirrors (see fade_oracle.h).
 *
 * Why it exists: the reference's SW is parasail's SIMD kernel (sw_trace_striped_16, call site
 * source/analysis.d:67), so a scalar port would understate the CPU.  parasail itself cannot be
 * built here; this is a from-scratch vectorisation with the same per-call work: DP fill with a
 * full trace table (one byte per cell and lane), end-cell selection, traceback to a CIGAR, S
 * padding and the accept predicates -- and, more favourable to the CPU than real fade, an
 * in-memory reference (no faidx mutex) and no per-call allocations.
 * Alignments whose read or window contains a letter outside ACGTN use the scalar oracle.
 */
#include "fade_oracle.h"
#include <immintrin.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_LANES 32
enum { T_ZERO = 0, T_DIAG = 1, T_F = 2, T_E = 3, T_EOPEN = 4, T_FOPEN = 8 };

static inline int map5(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'T': case 't': return 2;
    case 'G': case 'g': return 3; case 'N': case 'n': return 4; default: return 5;
    }
}

typedef struct {
    int16_t *q;      /* [qmax][lanes] query codes (pad = 100+lane-independent mismatch code) */
    int16_t *t;      /* [tmax][lanes] target codes */
    int16_t *H, *E;  /* [qmax][lanes] */
    uint8_t *tr;     /* [tmax][qmax][lanes] */
    size_t cap_q, cap_t, cap_tr;
} simd_ws;

static void *alloc64(size_t bytes) { return aligned_alloc(64, (bytes + 63) & ~(size_t)63); }

static void ws_reserve(simd_ws *w, int qmax, int tmax, int lanes)
{
    if ((size_t)qmax > w->cap_q) {
        free(w->q); free(w->H); free(w->E);
        w->cap_q = (size_t)qmax + 64;
        w->q = (int16_t *)alloc64(w->cap_q * lanes * 2);
        w->H = (int16_t *)alloc64(w->cap_q * lanes * 2);
        w->E = (int16_t *)alloc64(w->cap_q * lanes * 2);
    }
    if ((size_t)tmax > w->cap_t) {
        free(w->t);
        w->cap_t = (size_t)tmax + 256;
        w->t = (int16_t *)alloc64(w->cap_t * lanes * 2);
    }
    const size_t need = (size_t)qmax * tmax * lanes;
    if (need > w->cap_tr) {
        free(w->tr);
        w->cap_tr = need + need / 4;
        w->tr = (uint8_t *)alloc64(w->cap_tr);
    }
}

typedef void (*simd_kernel)(simd_ws *w, int qmax, int tmax, const fo_params *p, int *score, int *end_q, int *end_r);

__attribute__((target("avx2")))
static void simd_sw16(simd_ws *w, int qmax, int tmax, const fo_params *p,
                      int *score, int *end_q, int *end_r)
{
    const __m256i vo = _mm256_set1_epi16((short)p->gap_open), ve = _mm256_set1_epi16((short)p->gap_extend);
    const __m256i vmatch = _mm256_set1_epi16((short)p->match), vmis = _mm256_set1_epi16((short)p->mismatch);
    const __m256i zero = _mm256_setzero_si256();
    const __m256i neg = _mm256_set1_epi16(-20000);
    const __m256i c1 = _mm256_set1_epi16(1), c3 = _mm256_set1_epi16(3);
    const __m256i c4 = _mm256_set1_epi16(4), c8 = _mm256_set1_epi16(8);
    enum { LANES = 16 };
    /* locals: the byte stores into the trace table may alias *w as far as the compiler knows */
    __m256i *const H = (__m256i *)w->H, *const E = (__m256i *)w->E;
    const int16_t *const wq = w->q, *const wt = w->t;
    uint8_t *const wtr = w->tr;
    __m256i best = zero, bi = zero, bj = zero;
    for (int i = 0; i < qmax; ++i) { H[i] = zero; E[i] = neg; }
    for (int j = 0; j < tmax; ++j) {
        const __m256i tv = _mm256_load_si256((const __m256i *)(wt + (size_t)j * LANES));
        const __m256i vj = _mm256_set1_epi16((short)j);
        __m256i hdiag = zero, hup = zero, f = neg;
        /* end cell, first column then first row, strictly greater only: inside a column only the first row that beats
         * everything before it is tracked (cbest starts at the best of the earlier columns); the column number and the
         * global best are updated once per column */
        __m256i cbest = best, ci = zero, vi = zero;
        uint8_t *trj = wtr + (size_t)j * qmax * LANES;   /* column after column: the stores are sequential */
        for (int i = 0; i < qmax; ++i) {
            const __m256i qv = _mm256_load_si256((const __m256i *)(wq + (size_t)i * LANES));
            const __m256i hleft = H[i];
            /* E[i][j] = max(H[i][j-1]-o, E[i][j-1]-e); open iff strictly greater */
            const __m256i e_opn = _mm256_subs_epi16(hleft, vo), e_ext = _mm256_subs_epi16(E[i], ve);
            const __m256i eo = _mm256_cmpgt_epi16(e_opn, e_ext);
            const __m256i ev = _mm256_max_epi16(e_opn, e_ext);
            const __m256i f_opn = _mm256_subs_epi16(hup, vo), f_ext = _mm256_subs_epi16(f, ve);
            const __m256i fo = _mm256_cmpgt_epi16(f_opn, f_ext);
            const __m256i fv = _mm256_max_epi16(f_opn, f_ext);
            const __m256i s = _mm256_blendv_epi8(vmis, vmatch, _mm256_cmpeq_epi16(qv, tv));
            const __m256i hd = _mm256_max_epi16(_mm256_adds_epi16(hdiag, s), zero);
            const __m256i h = _mm256_max_epi16(_mm256_max_epi16(hd, ev), fv);
            /* source priority DIAG/ZERO > F > E; the comparison masks are 0 / -1, so 3 + is_f is E or F and 1 + is_z
             * is DIAG or ZERO */
            const __m256i is_d = _mm256_cmpeq_epi16(h, hd), is_f = _mm256_cmpeq_epi16(h, fv);
            const __m256i is_z = _mm256_cmpeq_epi16(h, zero);
            const __m256i src = _mm256_blendv_epi8(_mm256_add_epi16(c3, is_f), _mm256_add_epi16(c1, is_z), is_d);
            __m256i tb = _mm256_or_si256(src, _mm256_or_si256(_mm256_and_si256(eo, c4), _mm256_and_si256(fo, c8)));
            /* 16 x int16 -> 16 bytes */
            const __m256i pk = _mm256_packus_epi16(tb, tb);
            const __m128i lo = _mm256_castsi256_si128(pk), hi = _mm256_extracti128_si256(pk, 1);
            _mm_storeu_si128((__m128i *)(trj + (size_t)i * LANES), _mm_unpacklo_epi64(lo, hi));
            const __m256i gt = _mm256_cmpgt_epi16(h, cbest);
            cbest = _mm256_max_epi16(cbest, h);
            ci = _mm256_blendv_epi8(ci, vi, gt);
            vi = _mm256_add_epi16(vi, c1);
            hdiag = hleft; hup = h; f = fv;
            E[i] = ev; H[i] = h;
        }
        const __m256i upd = _mm256_cmpgt_epi16(cbest, best);
        best = cbest;
        bi = _mm256_blendv_epi8(bi, ci, upd);
        bj = _mm256_blendv_epi8(bj, vj, upd);
    }
    int16_t b[LANES], xi[LANES], xj[LANES];
    _mm256_storeu_si256((__m256i *)b, best); _mm256_storeu_si256((__m256i *)xi, bi); _mm256_storeu_si256((__m256i *)xj, bj);
    for (int l = 0; l < LANES; ++l) { score[l] = b[l]; end_q[l] = xi[l]; end_r[l] = xj[l]; }
}

/* the same recurrence on 32 lanes; comparisons land in mask registers and the trace byte is assembled in the byte domain */
__attribute__((target("avx512f,avx512bw,avx512vl")))
static void simd_sw32(simd_ws *w, int qmax, int tmax, const fo_params *p,
                      int *score, int *end_q, int *end_r)
{
    enum { LANES = 32 };
    const __m512i vo = _mm512_set1_epi16((short)p->gap_open), ve = _mm512_set1_epi16((short)p->gap_extend);
    const __m512i vmatch = _mm512_set1_epi16((short)p->match), vmis = _mm512_set1_epi16((short)p->mismatch);
    const __m512i zero = _mm512_setzero_si512();
    const __m512i neg = _mm512_set1_epi16(-20000), one = _mm512_set1_epi16(1);
    const __m256i b1 = _mm256_set1_epi8(1), b2 = _mm256_set1_epi8(2), b3 = _mm256_set1_epi8(3);
    const __m256i b4 = _mm256_set1_epi8(4), b8 = _mm256_set1_epi8(8);
    __m512i *const H = (__m512i *)w->H, *const E = (__m512i *)w->E;
    const int16_t *const wq = w->q, *const wt = w->t;
    uint8_t *const wtr = w->tr;
    __m512i best = zero, bi = zero, bj = zero;
    for (int i = 0; i < qmax; ++i) { H[i] = zero; E[i] = neg; }
    for (int j = 0; j < tmax; ++j) {
        const __m512i tv = _mm512_load_si512((const void *)(wt + (size_t)j * LANES));
        const __m512i vj = _mm512_set1_epi16((short)j);
        __m512i hdiag = zero, hup = zero, f = neg;
        __m512i cbest = best, ci = zero, vi = zero;      /* end cell as in simd_sw16 */
        uint8_t *trj = wtr + (size_t)j * qmax * LANES;   /* column after column: the stores are sequential */
        for (int i = 0; i < qmax; ++i) {
            const __m512i qv = _mm512_load_si512((const void *)(wq + (size_t)i * LANES));
            const __m512i hleft = H[i];
            const __m512i e_opn = _mm512_subs_epi16(hleft, vo), e_ext = _mm512_subs_epi16(E[i], ve);
            const __mmask32 eo = _mm512_cmpgt_epi16_mask(e_opn, e_ext);
            const __m512i ev = _mm512_max_epi16(e_opn, e_ext);
            const __m512i f_opn = _mm512_subs_epi16(hup, vo), f_ext = _mm512_subs_epi16(f, ve);
            const __mmask32 fo = _mm512_cmpgt_epi16_mask(f_opn, f_ext);
            const __m512i fv = _mm512_max_epi16(f_opn, f_ext);
            const __m512i s = _mm512_mask_blend_epi16(_mm512_cmpeq_epi16_mask(qv, tv), vmis, vmatch);
            const __m512i hd = _mm512_max_epi16(_mm512_adds_epi16(hdiag, s), zero);
            const __m512i h = _mm512_max_epi16(_mm512_max_epi16(hd, ev), fv);
            /* source priority DIAG/ZERO > F > E */
            const __mmask32 is_d = _mm512_cmpeq_epi16_mask(h, hd), is_f = _mm512_cmpeq_epi16_mask(h, fv);
            const __mmask32 is_nz = _mm512_cmpneq_epi16_mask(h, zero);
            __m256i tb = _mm256_mask_blend_epi8(is_f, b3, b2);
            tb = _mm256_mask_mov_epi8(tb, is_d, _mm256_maskz_mov_epi8(is_nz, b1));
            tb = _mm256_or_si256(tb, _mm256_or_si256(_mm256_maskz_mov_epi8(eo, b4), _mm256_maskz_mov_epi8(fo, b8)));
            _mm256_storeu_si256((__m256i *)(trj + (size_t)i * LANES), tb);
            const __mmask32 gt = _mm512_cmpgt_epi16_mask(h, cbest);
            cbest = _mm512_max_epi16(cbest, h);
            ci = _mm512_mask_mov_epi16(ci, gt, vi);
            vi = _mm512_add_epi16(vi, one);
            hdiag = hleft; hup = h; f = fv;
            E[i] = ev; H[i] = h;
        }
        const __mmask32 upd = _mm512_cmpgt_epi16_mask(cbest, best);
        best = cbest;
        bi = _mm512_mask_mov_epi16(bi, upd, ci);
        bj = _mm512_mask_mov_epi16(bj, upd, vj);
    }
    int16_t b[LANES], xi[LANES], xj[LANES];
    _mm512_storeu_si512((void *)b, best); _mm512_storeu_si512((void *)xi, bi); _mm512_storeu_si512((void *)xj, bj);
    for (int l = 0; l < LANES; ++l) { score[l] = b[l]; end_q[l] = xi[l]; end_r[l] = xj[l]; }
}

/* lanes of the widest kernel this host runs (16 or 32; 0 = no AVX2) */
int fo_simd_lanes(void)
{
    if (!__builtin_cpu_supports("avx2")) return 0;
    const char *force = getenv("FADE_ORACLE_SIMD");
    if (force && strcmp(force, "avx2") == 0) return 16;
    return (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl")) ? 32 : 16;
}

static int accept_side(int left, int score, int n_ops, const uint32_t *ops, int n_have, uint32_t clip_len)
{
    if (n_ops == 0 || n_ops > 10 || n_have < n_ops) return 0;
    const uint32_t edge = left ? ops[n_ops - 1] : ops[0];
    if ((edge & 0xf) != FO_EQ) return 0;
    const float cutoff = (float)(clip_len * 0.9 * 2);
    if (!((float)score > cutoff)) return 0;
    uint32_t clips[2];
    fo_parse_clips(ops, n_ops, clips);
    if (left) return !((clips[1] >> 4) != 0 || (clips[0] >> 4) == 0);
    return !((clips[0] >> 4) != 0 || (clips[1] >> 4) == 0);
}

/* same contract as fo_align_batch */
int fo_align_batch_simd(int64_t n, const uint8_t *seq4, const int64_t *seq_off, const int32_t *l_qseq,
                        const int32_t *tid, const int64_t *pos, const int32_t *aligned_len,
                        const int32_t *clip_left, const int32_t *clip_right,
                        int n_contigs, const char *const *contigs, const int64_t *contig_len,
                        const fo_params *p, fo_read_result *res, uint32_t *ops_out, int ops_cap,
                        int n_threads)
{
    const int LANES = fo_simd_lanes();
    if (LANES == 0)
        return fo_align_batch(n, seq4, seq_off, l_qseq, tid, pos, aligned_len, clip_left, clip_right, n_contigs,
                              contigs, contig_len, p, res, ops_out, ops_cap, n_threads);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    /* which reads need SW (analysis.d:34) and their windows (analysis.d:45-59) */
    int64_t *list = (int64_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
    int64_t m = 0;
    for (int64_t k = 0; k < n; ++k) {
        fo_read_result *r = &res[k];
        memset(r, 0, sizeof(*r));
        if (tid[k] < 0 || tid[k] >= n_contigs) continue;
        const uint32_t cl = (uint32_t)clip_left[k], cr = (uint32_t)clip_right[k];
        const int dl = cl != 0 && !(cl <= (uint32_t)p->min_length), dr = cr != 0 && !(cr <= (uint32_t)p->min_length);
        if (!dl && !dr) continue;
        int64_t start = pos[k] - p->window_size;
        if (start < 0) start = 0;
        int64_t end = pos[k] + aligned_len[k] + p->window_size;
        if (end > contig_len[tid[k]]) end = contig_len[tid[k]];
        r->win_start = start;
        if (end <= start || l_qseq[k] <= 0) continue;
        r->tlen = (int)(end - start);
        r->aligned = 1;
        list[m++] = k;
    }
    const simd_kernel kernel = LANES == 32 ? simd_sw32 : simd_sw16;
    const int64_t nb = (m + LANES - 1) / LANES;
    int err = 0;
#pragma omp parallel
    {
        simd_ws w;
        memset(&w, 0, sizeof(w));
        uint32_t *rev = NULL;
        size_t rev_cap = 0;
#pragma omp for schedule(dynamic, 4)
        for (int64_t bx = 0; bx < nb; ++bx) {
            const int cnt = (int)((bx + 1) * LANES <= m ? LANES : m - bx * LANES);
            int qmax = 1, tmax = 1;
            for (int l = 0; l < cnt; ++l) {
                const int64_t k = list[bx * LANES + l];
                if (l_qseq[k] > qmax) qmax = l_qseq[k];
                if (res[k].tlen > tmax) tmax = res[k].tlen;
            }
            ws_reserve(&w, qmax, tmax, LANES);
            int wild[MAX_LANES];
            for (int l = 0; l < LANES; ++l) wild[l] = 0;
            /* codes: query pad 100, target pad 200 -> never equal */
            for (int i = 0; i < qmax; ++i)
                for (int l = 0; l < LANES; ++l) w.q[(size_t)i * LANES + l] = 100;
            for (int j = 0; j < tmax; ++j)
                for (int l = 0; l < LANES; ++l) w.t[(size_t)j * LANES + l] = 200;
            for (int l = 0; l < cnt; ++l) {
                const int64_t k = list[bx * LANES + l];
                const int L = l_qseq[k];
                char qbuf[1024];
                char *qq = L <= 1024 ? qbuf : (char *)malloc((size_t)L);
                fo_revcomp_nt16(seq4 + seq_off[k], L, qq);                    /* analysis.d:40 */
                for (int i = 0; i < L; ++i) {
                    const int c = map5((unsigned char)qq[i]);
                    if (c == 5) wild[l] = 1;
                    w.q[(size_t)i * LANES + l] = (int16_t)c;
                }
                if (qq != qbuf) free(qq);
                const char *ref = contigs[tid[k]] + res[k].win_start;
                for (int j = 0; j < res[k].tlen; ++j) {
                    const int c = map5((unsigned char)ref[j]);                /* .toUpper is implied by map5 */
                    if (c == 5) wild[l] = 1;
                    w.t[(size_t)j * LANES + l] = (int16_t)c;
                }
            }
            int score[MAX_LANES], eq[MAX_LANES], er[MAX_LANES];
            kernel(&w, qmax, tmax, p, score, eq, er);
            for (int l = 0; l < cnt; ++l) {
                const int64_t k = list[bx * LANES + l];
                fo_read_result *r = &res[k];
                uint32_t *ops = ops_out + (size_t)k * ops_cap;
                if (wild[l]) {   /* wildcard letters: scalar oracle */
                    int rc = fo_align_read(seq4 + seq_off[k], l_qseq[k], pos[k], aligned_len[k], (uint32_t)clip_left[k],
                                           (uint32_t)clip_right[k], contigs[tid[k]], contig_len[tid[k]], p, r, ops, ops_cap);
                    if (rc) {
#pragma omp atomic write
                        err = rc;
                    }
                    continue;
                }
                const int L = l_qseq[k], T = r->tlen;
                memset(&r->sw, 0, sizeof(r->sw));
                r->sw.score = score[l];
                if (score[l] <= 0) continue;
                if ((size_t)(L + T + 4) > rev_cap) { free(rev); rev_cap = (size_t)(L + T + 64); rev = (uint32_t *)malloc(rev_cap * 4); }
                int i = eq[l], j = er[l], state = 0, nrev = 0, span = 0;
                while (i >= 0 && j >= 0) {
                    const uint8_t tb = w.tr[((size_t)j * qmax + i) * LANES + l];
                    uint32_t op;
                    if (state == 0) {
                        const int src = tb & 3;
                        if (src == T_ZERO) break;
                        if (src == T_DIAG) {
                            op = w.q[(size_t)i * LANES + l] == w.t[(size_t)j * LANES + l] ? FO_EQ : FO_X;
                            ++span; --i; --j;
                        } else { state = src == T_F ? 1 : 2; continue; }
                    } else if (state == 1) { op = FO_I; --i; if (tb & T_FOPEN) state = 0; }
                    else { op = FO_D; ++span; --j; if (tb & T_EOPEN) state = 0; }
                    if (nrev > 0 && (rev[nrev - 1] & 0xf) == op) rev[nrev - 1] += 16;
                    else rev[nrev++] = (1u << 4) | op;
                }
                const int lead = i + 1, trail = L - 1 - eq[l];
                int nn = 0;
                if (lead > 0) { if (nn < ops_cap) ops[nn] = ((uint32_t)lead << 4) | FO_S; ++nn; }
                for (int x = nrev - 1; x >= 0; --x) { if (nn < ops_cap) ops[nn] = rev[x]; ++nn; }
                if (trail > 0) { if (nn < ops_cap) ops[nn] = ((uint32_t)trail << 4) | FO_S; ++nn; }
                r->sw.end_query = eq[l]; r->sw.end_ref = er[l]; r->sw.beg_query = i + 1; r->sw.beg_ref = j + 1;
                r->sw.n_ops = nn; r->sw.ref_span = span;
                const int have = nn < ops_cap ? nn : ops_cap;
                const uint32_t cl = (uint32_t)clip_left[k], cr = (uint32_t)clip_right[k];
                if (cl != 0 && !(cl <= (uint32_t)p->min_length)) r->art_left = accept_side(1, score[l], nn, ops, have, cl);
                if (cr != 0 && !(cr <= (uint32_t)p->min_length)) r->art_right = accept_side(0, score[l], nn, ops, have, cr);
            }
        }
        free(w.q); free(w.t); free(w.H); free(w.E); free(w.tr); free(rev);
    }
    free(list);
    return err;
}
