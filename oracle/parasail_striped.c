/*
 * parasail_striped.c -- STRUCTURAL restatement of parasail 2.4.3's
 *     parasail_sw_trace_striped_{sse2,sse41}_128_16 / _avx2_256_16   (Farrar striped layout,
 *     lazy-F correction loop WITH its trace-table rewrites, column-max end-cell bookkeeping)
 *   + parasail_result_get_cigar over the striped trace table
 *   + dparasail's result wrapper (S padding),
 * i.e. the code reached from `p.sw_striped(q_seq, ref_seq)` / `res.cigar` at
 * /root/reference/source/analysis.d:67,69.
 *
 * TEST INFRASTRUCTURE ONLY (see fade_oracle.h).  Why it exists: fade_oracle.c restates the
 * RECURRENCE (rules P1-P5 of SURVEY.md 8a) as a clean scalar Gotoh DP and argues (SURVEY 8a,
 * "lane-width independence") that parasail's striped kernel gives the same answers for every SIMD
 * width.  This file replaces the argument by a check: it re-implements the striped kernel the way
 * upstream structures it -- vectors of `lanes` int16 values emulated lane by lane, segLen =
 * ceil(qlen / lanes), the query profile with zero padding, the score-carrying E array that is NOT
 * corrected by lazy-F next to the "accurate" Ea array that only feeds trace bits, the F seed of
 * -open, the lazy-F loop that re-derives the H source and the E / F trace bits of every cell it
 * touches, the H-column triple buffering that keeps the column of the best score, the per-column
 * `vMaxH > vMaxHUnit` test, the end_query scan in striped order -- and tests/test_striped_parity.py
 * fuzzes it against fo_sw_trace at 8 lanes (SSE2/SSE4.1/NEON builds), 16 lanes (AVX2 builds) and 32 lanes (the
 * lane count of the 8-bit AVX2 kernel: U10 -- dparasail's sw_striped may be parasail's saturating dispatcher, which
 * tries the 8-bit kernel first and keeps its answer when the score stays below 127 - match; without saturation the
 * 8-bit kernels compute the same recurrence, so only their lane count could matter).
 *
 * PROVENANCE / LIMIT: parasail's source is not in /root/reference nor installable here (no
 * network); this is restated from the author's knowledge of upstream src/sw_trace_striped.c,
 * src/cigar.c and src/matrix_lookup / parasail_matrix_create, pinned version 2.4.3
 * (/root/reference/Dockerfile:4).  It cannot be diffed against the upstream text offline, so it
 * narrows "parity unpinned" (the striped structure no longer rests on an argument) but does not
 * remove it: tools/export_external_check.py packages the one external run that would.
 */
#include "fade_oracle.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PS_MAX_LANES 32
/* vectors are ps_vec of PS_MAX_LANES slots of which only the first `lanes` are ever written or read */
#pragma GCC diagnostic ignored "-Wmaybe-uninitialized"

/* parasail.h trace bits */
enum {
    PS_ZERO = 0, PS_INS = 1, PS_DEL = 2, PS_DIAG = 4,
    PS_DIAG_E = 8, PS_INS_E = 16, PS_DIAG_F = 32, PS_DEL_F = 64,
    PS_ZERO_MASK = 120, /* all bits except the H source */
    PS_E_MASK = 103,    /* all bits except the E bits */
    PS_F_MASK = 31      /* all bits except the F bits */
};

typedef struct { int16_t v[PS_MAX_LANES]; } ps_vec;

static inline int16_t ps_sat(int x) { return (int16_t)(x > 32767 ? 32767 : (x < -32768 ? -32768 : x)); }

/* the vector instructions upstream uses, over `L` lanes */
static inline ps_vec ps_set1(int L, int x) { ps_vec r; for (int l = 0; l < L; ++l) r.v[l] = (int16_t)x; return r; }
static inline ps_vec ps_adds(int L, ps_vec a, ps_vec b) { ps_vec r; for (int l = 0; l < L; ++l) r.v[l] = ps_sat(a.v[l] + b.v[l]); return r; }
static inline ps_vec ps_subs(int L, ps_vec a, ps_vec b) { ps_vec r; for (int l = 0; l < L; ++l) r.v[l] = ps_sat(a.v[l] - b.v[l]); return r; }
static inline ps_vec ps_max(int L, ps_vec a, ps_vec b) { ps_vec r; for (int l = 0; l < L; ++l) r.v[l] = a.v[l] > b.v[l] ? a.v[l] : b.v[l]; return r; }
static inline ps_vec ps_or(int L, ps_vec a, ps_vec b) { ps_vec r; for (int l = 0; l < L; ++l) r.v[l] = (int16_t)(a.v[l] | b.v[l]); return r; }
static inline ps_vec ps_and(int L, ps_vec a, ps_vec b) { ps_vec r; for (int l = 0; l < L; ++l) r.v[l] = (int16_t)(a.v[l] & b.v[l]); return r; }
static inline ps_vec ps_andnot(int L, ps_vec a, ps_vec b) { ps_vec r; for (int l = 0; l < L; ++l) r.v[l] = (int16_t)(~a.v[l] & b.v[l]); return r; }
static inline ps_vec ps_cmpeq(int L, ps_vec a, ps_vec b) { ps_vec r; for (int l = 0; l < L; ++l) r.v[l] = (int16_t)(a.v[l] == b.v[l] ? -1 : 0); return r; }
static inline ps_vec ps_cmpgt(int L, ps_vec a, ps_vec b) { ps_vec r; for (int l = 0; l < L; ++l) r.v[l] = (int16_t)(a.v[l] > b.v[l] ? -1 : 0); return r; }
/* blendv(a, b, mask): mask ? b : a */
static inline ps_vec ps_blendv(int L, ps_vec a, ps_vec b, ps_vec m) { ps_vec r; for (int l = 0; l < L; ++l) r.v[l] = m.v[l] ? b.v[l] : a.v[l]; return r; }
/* _mm_slli_si128(v, 2): lane l <- lane l-1, lane 0 <- 0 */
static inline ps_vec ps_shift(int L, ps_vec a) { ps_vec r; r.v[0] = 0; for (int l = 1; l < L; ++l) r.v[l] = a.v[l - 1]; return r; }
static inline ps_vec ps_insert0(ps_vec a, int x) { a.v[0] = (int16_t)x; return a; }
static inline int ps_any(int L, ps_vec a) { for (int l = 0; l < L; ++l) if (a.v[l]) return 1; return 0; }
static inline int ps_hmax(int L, ps_vec a) { int m = a.v[0]; for (int l = 1; l < L; ++l) if (a.v[l] > m) m = a.v[l]; return m; }

/* parasail_matrix_create(alphabet, match, mismatch): size+1 symbols, the extra row / column (the
 * wildcard every unlisted byte maps to) scores 0; mapper is case-insensitive.  Call site:
 * /root/reference/source/anno.d:36 with "ACTGN". */
typedef struct {
    int size;              /* strlen(alphabet) + 1 */
    int mapper[256];
    int matrix[8 * 8];
    int max, min;
} ps_matrix;

static void ps_matrix_create(ps_matrix *m, const char *alphabet, int match, int mismatch)
{
    const int n = (int)strlen(alphabet), n1 = n + 1;
    int c = 0;
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) m->matrix[c++] = i == j ? match : mismatch;
        m->matrix[c++] = 0;
    }
    for (int j = 0; j < n1; ++j) m->matrix[c++] = 0;
    for (int i = 0; i < 256; ++i) m->mapper[i] = n;
    for (int i = 0; i < n; ++i) {
        int ch = (unsigned char)alphabet[i];
        int up = (ch >= 'a' && ch <= 'z') ? ch - 32 : ch, lo = (ch >= 'A' && ch <= 'Z') ? ch + 32 : ch;
        m->mapper[up] = i;
        m->mapper[lo] = i;
    }
    m->size = n1;
    m->max = match > mismatch ? match : mismatch;
    m->min = match > mismatch ? mismatch : match;
    if (m->max < 0) m->max = 0;
    if (m->min > 0) m->min = 0;
}

typedef struct {
    int32_t score, end_query, end_ref, beg_query, beg_ref;
    int32_t n_ops;        /* parasail's ops (no S padding) */
    int32_t saturated;
} ps_result;

/* trace table element (i, j) in striped order: vector (j*segLen + i % segLen), lane i / segLen */
#define PS_T(vi, j) (tt[(size_t)(j) * (size_t)segLen + (size_t)(vi)])

/*
 * parasail_sw_trace_striped_*_16(s1 = query, s2 = database) + parasail_result_get_cigar.
 * ops receives parasail's BAM-encoded CIGAR (no clipping ops) in forward order.
 * lanes = 8 (128-bit builds) or 16 (AVX2); any value in [1, 32] is accepted for experiments.
 */
static inline __attribute__((always_inline)) int ps_core(const char *s1, int s1Len, const char *s2, int s2Len, int open, int gap,
                        int match, int mismatch, const int lanes, ps_result *res, uint32_t *ops, int ops_cap)
{
    const int L = lanes;
    ps_matrix mat;
    ps_matrix_create(&mat, "ACTGN", match, mismatch);
    const int n = mat.size;
    const int segWidth = L;
    const int segLen = (s1Len + segWidth - 1) / segWidth;

    /* parasail_profile_create_*_16: profile[k][i] lane s = matrix[k][s1[i + s*segLen]], 0 beyond s1Len */
    ps_vec *vProfile = (ps_vec *)malloc(sizeof(ps_vec) * (size_t)n * (size_t)segLen);
    ps_vec *bufs = (ps_vec *)malloc(sizeof(ps_vec) * (size_t)segLen * 7);
    ps_vec *tt = (ps_vec *)calloc((size_t)segLen * (size_t)s2Len, sizeof(ps_vec));
    if (!vProfile || !bufs || !tt) { free(vProfile); free(bufs); free(tt); return -1; }
    {
        size_t index = 0;
        for (int k = 0; k < n; ++k)
            for (int i = 0; i < segLen; ++i) {
                ps_vec t;
                int j = i;
                for (int s = 0; s < segWidth; ++s) {
                    t.v[s] = (int16_t)(j >= s1Len ? 0 : mat.matrix[n * k + mat.mapper[(unsigned char)s1[j]]]);
                    j += segLen;
                }
                vProfile[index++] = t;
            }
    }
    ps_vec *pvHStore = bufs, *pvHLoad = bufs + segLen, *pvE = bufs + 2 * segLen, *pvEaStore = bufs + 3 * segLen,
           *pvEaLoad = bufs + 4 * segLen, *pvHT = bufs + 5 * segLen, *pvHMax = bufs + 6 * segLen;
    const ps_vec vGapO = ps_set1(L, open), vGapE = ps_set1(L, gap), vZero = ps_set1(L, 0);
    const int NEG_LIMIT = (-open < mat.min ? INT16_MIN + open : INT16_MIN - mat.min) + 1;
    int score = NEG_LIMIT;
    ps_vec vMaxH = vZero, vMaxHUnit = vZero;
    const int maxp = INT16_MAX - (mat.max + 1);
    const ps_vec vTZero = ps_set1(L, PS_ZERO), vTIns = ps_set1(L, PS_INS), vTDel = ps_set1(L, PS_DEL),
                 vTDiag = ps_set1(L, PS_DIAG), vTDiagE = ps_set1(L, PS_DIAG_E), vTInsE = ps_set1(L, PS_INS_E),
                 vTDiagF = ps_set1(L, PS_DIAG_F), vTDelF = ps_set1(L, PS_DEL_F), vTMask = ps_set1(L, PS_ZERO_MASK),
                 vFTMask = ps_set1(L, PS_F_MASK);
    int end_query = 0, end_ref = 0, saturated = 0;
    int i, j, k;

    for (i = 0; i < segLen; ++i) {
        pvHStore[i] = vZero;
        pvHLoad[i] = vZero;
        pvHMax[i] = vZero;
        pvE[i] = ps_set1(L, -open);
        pvEaStore[i] = ps_set1(L, -open);
        pvEaLoad[i] = ps_set1(L, -open);
        PS_T(i, 0) = vTDiagE;        /* column 0: every E was "opened" */
    }

    /* outer loop over the database sequence */
    for (j = 0; j < s2Len; ++j) {
        ps_vec vEF_opn = vZero, vE, vE_ext, vF, vF_ext = vZero, vFa, vFa_ext, vH, vH_dag;
        const ps_vec *vP;
        /* F is seeded with -open; errors this causes in H are repaired by the lazy-F loop */
        vF = ps_subs(L, vZero, vGapO);
        /* last segment of the previous column, shifted by one lane */
        vH = ps_shift(L, pvHStore[segLen - 1]);
        vP = vProfile + (size_t)mat.mapper[(unsigned char)s2[j]] * (size_t)segLen;
        if (end_ref == j - 2) {          /* keep the column that holds the best score so far */
            ps_vec *t = pvHMax; pvHMax = pvHLoad; pvHLoad = pvHStore; pvHStore = t;
        } else {
            ps_vec *t = pvHLoad; pvHLoad = pvHStore; pvHStore = t;
        }
        { ps_vec *t = pvEaLoad; pvEaLoad = pvEaStore; pvEaStore = t; }

        /* inner loop over the query segments */
        for (i = 0; i < segLen; ++i) {
            vE = pvE[i];
            vH_dag = ps_max(L, ps_adds(L, vH, vP[i]), vZero);
            vH = ps_max(L, ps_max(L, vH_dag, vE), vF);
            pvHStore[i] = vH;
            {   /* H source: DIAG (ZERO when the value is 0) > DEL (F) > INS (E) */
                const ps_vec vTAll = PS_T(i, j);
                const ps_vec cond_zero = ps_cmpeq(L, vH, vZero);
                const ps_vec case1 = ps_cmpeq(L, vH, vH_dag);
                const ps_vec case2 = ps_cmpeq(L, vH, vF);
                ps_vec vT = ps_blendv(L, ps_blendv(L, vTIns, vTDel, case2), ps_blendv(L, vTDiag, vTZero, cond_zero), case1);
                pvHT[i] = vT;
                PS_T(i, j) = ps_or(L, vT, vTAll);
            }
            vMaxH = ps_max(L, vH, vMaxH);
            vEF_opn = ps_subs(L, vH, vGapO);
            /* E of the next column: the score-carrying copy ... */
            vE_ext = ps_subs(L, vE, vGapE);
            vE = ps_max(L, vEF_opn, vE_ext);
            pvE[i] = vE;
            {   /* ... and the "accurate" copy that only decides the trace bit */
                ps_vec vEa = pvEaLoad[i];
                const ps_vec vEa_ext = ps_subs(L, vEa, vGapE);
                vEa = ps_max(L, vEF_opn, vEa_ext);
                pvEaStore[i] = vEa;
                if (j + 1 < s2Len) PS_T(i, j + 1) = ps_blendv(L, vTInsE, vTDiagE, ps_cmpgt(L, vEF_opn, vEa_ext));
            }
            /* F of the next segment row */
            vF_ext = ps_subs(L, vF, vGapE);
            vF = ps_max(L, vEF_opn, vF_ext);
            if (i + 1 < segLen) {
                const ps_vec vTAll = PS_T(i + 1, j);
                const ps_vec vT = ps_blendv(L, vTDelF, vTDiagF, ps_cmpgt(L, vEF_opn, vF_ext));
                PS_T(i + 1, j) = ps_or(L, vT, vTAll);
            }
            vH = pvHLoad[i];
        }

        /* lazy-F loop (does not update the score-carrying E: "disallow adjacent insertion and then
         * deletion"); rewrites H, the H source, the F bit and the accurate E of what it touches */
        vFa_ext = vF_ext;
        vFa = vF;
        for (k = 0; k < segWidth; ++k) {
            ps_vec vHp = ps_shift(L, pvHLoad[segLen - 1]);
            vEF_opn = ps_insert0(ps_shift(L, vEF_opn), -open);
            vF_ext = ps_insert0(ps_shift(L, vF_ext), NEG_LIMIT);
            vF = ps_insert0(ps_shift(L, vF), -open);
            vFa_ext = ps_insert0(ps_shift(L, vFa_ext), NEG_LIMIT);
            vFa = ps_insert0(ps_shift(L, vFa), -open);
            for (i = 0; i < segLen; ++i) {
                vH = ps_max(L, pvHStore[i], vF);
                pvHStore[i] = vH;
                {
                    ps_vec vTAll, vT, case1, case2, cond;
                    vHp = ps_max(L, ps_adds(L, vHp, vP[i]), vZero);
                    case1 = ps_cmpeq(L, vH, vHp);
                    case2 = ps_cmpeq(L, vH, vF);
                    cond = ps_andnot(L, case1, case2);
                    vTAll = PS_T(i, j);
                    vT = ps_blendv(L, pvHT[i], vTDel, cond);
                    pvHT[i] = vT;
                    vTAll = ps_or(L, ps_and(L, vTAll, vTMask), vT);
                    PS_T(i, j) = vTAll;
                }
                vMaxH = ps_max(L, vH, vMaxH);
                {   /* F bit of this cell from the accurate F */
                    ps_vec vTAll = PS_T(i, j);
                    const ps_vec vT = ps_blendv(L, vTDelF, vTDiagF, ps_cmpgt(L, vEF_opn, vFa_ext));
                    vTAll = ps_or(L, ps_and(L, vTAll, vFTMask), vT);
                    PS_T(i, j) = vTAll;
                }
                vEF_opn = ps_subs(L, vH, vGapO);
                vF_ext = ps_subs(L, vF, vGapE);
                {
                    ps_vec vEa = pvEaLoad[i];
                    const ps_vec vEa_ext = ps_subs(L, vEa, vGapE);
                    vEa = ps_max(L, vEF_opn, vEa_ext);
                    pvEaStore[i] = vEa;
                    if (j + 1 < s2Len) PS_T(i, j + 1) = ps_blendv(L, vTInsE, vTDiagE, ps_cmpgt(L, vEF_opn, vEa_ext));
                }
                if (!ps_any(L, ps_or(L, ps_cmpgt(L, vF_ext, vEF_opn), ps_cmpeq(L, vF_ext, vEF_opn)))) goto end;
                vF = vF_ext;
                vFa_ext = ps_subs(L, vFa, vGapE);
                vFa = ps_max(L, vEF_opn, vFa_ext);
                vHp = pvHLoad[i];
            }
        }
end:
        if (ps_any(L, ps_cmpgt(L, vMaxH, vMaxHUnit))) {
            score = ps_hmax(L, vMaxH);
            if (score > maxp) { saturated = 1; break; }
            vMaxHUnit = ps_set1(L, score);
            end_ref = j;
        }
    }

    if (score == INT16_MAX) saturated = 1;
    memset(res, 0, sizeof(*res));
    res->saturated = saturated;
    if (saturated) { free(vProfile); free(bufs); free(tt); return 0; }
    if (end_ref == j - 1) { ps_vec *t = pvHMax; pvHMax = pvHStore; pvHStore = t; }
    else if (end_ref == j - 2) { ps_vec *t = pvHMax; pvHMax = pvHLoad; pvHLoad = t; }
    {   /* end_query: the smallest query index holding the score in that column (striped scan) */
        const int column_len = segLen * segWidth;
        end_query = s1Len - 1;
        for (i = 0; i < column_len; ++i) {
            const int16_t hv = pvHMax[i / segWidth].v[i % segWidth];
            if (hv == score) {
                const int temp = i / segWidth + i % segWidth * segLen;
                if (temp < end_query) end_query = temp;
            }
        }
    }
    res->score = score;
    res->end_query = end_query;
    res->end_ref = end_ref;

    /* parasail_result_get_cigar (cigar.c, striped flavour): case-insensitive byte equality decides
     * '=' / 'X'; INS walks along the database (written 'D'), DEL along the query (written 'I') */
    {
        const int cap = s1Len + s2Len + 4;
        uint32_t *rev = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)cap);
        int nrev = 0;
        int where = PS_DIAG;
        int64_t ci = end_query, cj = end_ref;
        if (!rev) { free(vProfile); free(bufs); free(tt); return -1; }
#define PS_PUSH(op_)                                                                    \
    do {                                                                                \
        if (nrev > 0 && (rev[nrev - 1] & 0xf) == (uint32_t)(op_)) rev[nrev - 1] += 16;  \
        else rev[nrev++] = (1u << 4) | (uint32_t)(op_);                                 \
    } while (0)
        while (ci >= 0 && cj >= 0) {
            const int ht = PS_T(ci % segLen, cj).v[ci / segLen];
            if (where == PS_DIAG) {
                if (ht & PS_DIAG) {
                    int a = (unsigned char)s1[ci], b = (unsigned char)s2[cj];
                    if (a >= 'a' && a <= 'z') a -= 32;
                    if (b >= 'a' && b <= 'z') b -= 32;
                    PS_PUSH(a == b ? FO_EQ : FO_X);
                    --ci; --cj;
                } else if (ht & PS_INS) where = PS_INS;
                else if (ht & PS_DEL) where = PS_DEL;
                else break;                                  /* PS_ZERO */
            } else if (where == PS_INS) {
                PS_PUSH(FO_D);
                --cj;
                if (ht & PS_DIAG_E) where = PS_DIAG;
                else if (ht & PS_INS_E) where = PS_INS;
                else break;
            } else {
                PS_PUSH(FO_I);
                --ci;
                if (ht & PS_DIAG_F) where = PS_DIAG;
                else if (ht & PS_DEL_F) where = PS_DEL;
                else break;
            }
        }
#undef PS_PUSH
        res->beg_query = (int32_t)(ci + 1);
        res->beg_ref = (int32_t)(cj + 1);
        res->n_ops = nrev;
        for (int q = 0; q < nrev && q < ops_cap; ++q) ops[q] = rev[nrev - 1 - q];
        free(rev);
    }
    free(vProfile); free(bufs); free(tt);
    return 0;
}

int ps_sw_trace_striped(const char *s1, int s1Len, const char *s2, int s2Len, int open, int gap,
                        int match, int mismatch, int lanes, ps_result *res, uint32_t *ops, int ops_cap)
{
    if (!s1 || !s2 || s1Len <= 0 || s2Len <= 0 || lanes < 1 || lanes > PS_MAX_LANES || !res) return -1;
    switch (lanes) {   /* the widths upstream ships get a specialised (vectorisable) copy */
    case 8: return ps_core(s1, s1Len, s2, s2Len, open, gap, match, mismatch, 8, res, ops, ops_cap);
    case 16: return ps_core(s1, s1Len, s2, s2Len, open, gap, match, mismatch, 16, res, ops, ops_cap);
    case 32: return ps_core(s1, s1Len, s2, s2Len, open, gap, match, mismatch, 32, res, ops, ops_cap);
    default: return ps_core(s1, s1Len, s2, s2Len, open, gap, match, mismatch, lanes, res, ops, ops_cap);
    }
}

/* dparasail result wrapper on top (U1/U2): S(beg_query) first, S(qlen-1-end_query) last; same output
 * contract as fo_sw_trace so the two can be compared field by field. */
int ps_sw_trace(const char *q, int qlen, const char *t, int tlen, const fo_params *p, int lanes,
                fo_sw_result *r, uint32_t *ops, int ops_cap)
{
    ps_result pr;
    const int cap = qlen + tlen + 4;
    uint32_t *raw = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)cap);
    if (!raw) return -1;
    int rc = ps_sw_trace_striped(q, qlen, t, tlen, p->gap_open, p->gap_extend, p->match, p->mismatch, lanes, &pr, raw, cap);
    if (rc) { free(raw); return rc; }
    memset(r, 0, sizeof(*r));
    r->score = pr.score;
    if (pr.saturated || pr.score <= 0 || pr.n_ops == 0) {
        r->score = pr.saturated ? 0 : pr.score;
        free(raw);
        return 0;
    }
    r->end_query = pr.end_query; r->end_ref = pr.end_ref; r->beg_query = pr.beg_query; r->beg_ref = pr.beg_ref;
    int n = 0, span = 0;
    const int lead = pr.beg_query, trail = qlen - 1 - pr.end_query;
    if (lead > 0) { if (n < ops_cap && ops) ops[n] = ((uint32_t)lead << 4) | FO_S; ++n; }
    for (int k = 0; k < pr.n_ops; ++k) {
        const uint32_t op = raw[k] & 0xf;
        if (op == FO_EQ || op == FO_X || op == FO_D) span += (int)(raw[k] >> 4);
        if (n < ops_cap && ops) ops[n] = raw[k];
        ++n;
    }
    if (trail > 0) { if (n < ops_cap && ops) ops[n] = ((uint32_t)trail << 4) | FO_S; ++n; }
    r->n_ops = n;
    r->ref_span = span;
    free(raw);
    return 0;
}

/* ---- fuzz driver: striped restatement vs the scalar oracle ------------------------------------- */
static inline uint64_t ps_rng(uint64_t *s)
{
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

typedef struct {
    int64_t n_pairs, n_diverged, n_gapped, n_multi_max, n_zero_ef;   /* what the sample exercised */
    int64_t first_div;         /* index of the first diverging pair, -1 = none */
    int32_t first_qlen, first_tlen;
    char first_q[512], first_t[2048];
    uint32_t explained_by;     /* FO_SW_* switch (single bit) that makes the oracle agree on it, 0 = none */
} ps_fuzz_report;

/* pair `idx` of stream `seed`: mode 0 random ACGT(+N), 1 planted copy with substitutions, 2 planted
 * with indels, 3 low-complexity (2-letter / short-period repeats: ties and 0-valued cells), 4 mixed
 * with wildcard letters and lower case */
static void ps_make_pair(uint64_t seed, int64_t idx, int qmax, int tmax, char *q, int *qlen, char *t, int *tlen)
{
    uint64_t s = seed * 0x2545f4914f6cdd1dull + (uint64_t)idx * 0x9e3779b97f4a7c15ull + 1;
    const int mode = (int)(ps_rng(&s) % 5);
    static const char acgt[] = "ACGT";
    int ql = 8 + (int)(ps_rng(&s) % (uint64_t)(qmax - 7));
    int tl = 8 + (int)(ps_rng(&s) % (uint64_t)(tmax - 7));
    if (mode == 3) {
        const int period = 1 + (int)(ps_rng(&s) % 4);
        char unit[4];
        for (int k = 0; k < period; ++k) unit[k] = acgt[ps_rng(&s) & 3];
        for (int k = 0; k < ql; ++k) q[k] = (ps_rng(&s) % 16 == 0) ? acgt[ps_rng(&s) & 3] : unit[k % period];
        for (int k = 0; k < tl; ++k) t[k] = (ps_rng(&s) % 16 == 0) ? acgt[ps_rng(&s) & 3] : unit[(k + 1) % period];
    } else {
        for (int k = 0; k < ql; ++k) q[k] = acgt[ps_rng(&s) & 3];
        for (int k = 0; k < tl; ++k) t[k] = acgt[ps_rng(&s) & 3];
    }
    if (mode == 1 || mode == 2 || mode == 4) {
        /* plant a piece of q (a prefix, a suffix or an inner piece) into t */
        int pl = 6 + (int)(ps_rng(&s) % (uint64_t)(ql - 5));
        int qa = (ps_rng(&s) % 3 == 0) ? 0 : (int)(ps_rng(&s) % (uint64_t)(ql - pl + 1));
        if (ps_rng(&s) % 3 == 0) qa = ql - pl;
        char piece[600];
        int m = 0;
        for (int k = 0; k < pl && m < 590; ++k) {
            const uint64_t r = ps_rng(&s) % 100;
            if (mode != 1 && r < 3) { const int g = 1 + (int)(ps_rng(&s) % 6); for (int x = 0; x < g && m < 590; ++x) piece[m++] = acgt[ps_rng(&s) & 3]; }
            else if (mode != 1 && r < 6) { k += (int)(ps_rng(&s) % 6); continue; }
            piece[m++] = (r >= 94) ? acgt[ps_rng(&s) & 3] : q[qa + k];
        }
        if (m > tl) m = tl;
        const int ta = (int)(ps_rng(&s) % (uint64_t)(tl - m + 1));
        memcpy(t + ta, piece, (size_t)m);
    }
    if (mode == 4) {
        static const char odd[] = "NNNNRYKMSWnacgt";
        for (int k = 0; k < ql; ++k) if (ps_rng(&s) % 25 == 0) q[k] = odd[ps_rng(&s) % 15];
        for (int k = 0; k < tl; ++k) if (ps_rng(&s) % 25 == 0) t[k] = odd[ps_rng(&s) % 15];
    }
    *qlen = ql; *tlen = tl;
}

static int ps_same(const fo_sw_result *a, const uint32_t *oa, const fo_sw_result *b, const uint32_t *ob, int cap)
{
    if (a->score != b->score || a->n_ops != b->n_ops) return 0;
    if (a->score <= 0) return 1;
    if (a->end_query != b->end_query || a->end_ref != b->end_ref || a->beg_query != b->beg_query ||
        a->beg_ref != b->beg_ref || a->ref_span != b->ref_span) return 0;
    const int n = a->n_ops < cap ? a->n_ops : cap;
    return memcmp(oa, ob, sizeof(uint32_t) * (size_t)n) == 0;
}

/* does the clean DP of this pair contain a cell with H == 0 and E == 0 or F == 0 (the U8 situation),
 * or several cells holding the maximum (the P3 tie-breaks)? */
static void ps_census(const char *q, int ql, const char *t, int tl, const fo_params *p, int *multi_max, int *zero_ef)
{
    ps_matrix mat;
    ps_matrix_create(&mat, "ACTGN", p->match, p->mismatch);
    int *H = (int *)calloc((size_t)ql, sizeof(int)), *E = (int *)malloc(sizeof(int) * (size_t)ql);
    int best = 0, nbest = 0, z = 0;
    for (int i = 0; i < ql; ++i) E[i] = -(1 << 28);
    for (int j = 0; j < tl; ++j) {
        int hd = 0, hu = 0, f = -(1 << 28);
        for (int i = 0; i < ql; ++i) {
            const int hl = H[i];
            const int ev = hl - p->gap_open > E[i] - p->gap_extend ? hl - p->gap_open : E[i] - p->gap_extend;
            const int fv = hu - p->gap_open > f - p->gap_extend ? hu - p->gap_open : f - p->gap_extend;
            int d = hd + mat.matrix[mat.size * mat.mapper[(unsigned char)t[j]] + mat.mapper[(unsigned char)q[i]]];
            if (d < 0) d = 0;
            int h = d;
            if (ev > h) h = ev;
            if (fv > h) h = fv;
            if (h == 0 && (ev == 0 || fv == 0)) z = 1;
            if (h > best) { best = h; nbest = 1; } else if (h == best && h > 0) ++nbest;
            hd = hl; hu = h; f = fv; E[i] = ev; H[i] = h;
        }
    }
    free(H); free(E);
    *multi_max = nbest > 1;
    *zero_ef = z;
}

int ps_fuzz(uint64_t seed, int64_t n_pairs, int lanes, int qmax, int tmax, const fo_params *p, int n_threads,
            ps_fuzz_report *rep)
{
    if (qmax < 8 || qmax > 500 || tmax < 8 || tmax > 2000 || !rep) return -1;
    memset(rep, 0, sizeof(*rep));
    rep->n_pairs = n_pairs;
    rep->first_div = -1;
    int64_t n_div = 0, n_gap = 0, n_mm = 0, n_z = 0, first = -1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : n_div, n_gap, n_mm, n_z)
    for (int64_t idx = 0; idx < n_pairs; ++idx) {
        char q[512], t[2048];
        int ql, tl;
        ps_make_pair(seed, idx, qmax, tmax, q, &ql, t, &tl);
        fo_sw_result a, b;
        uint32_t oa[64], ob[64];
        fo_sw_trace(q, ql, t, tl, p, &a, oa, 64);
        ps_sw_trace(q, ql, t, tl, p, lanes, &b, ob, 64);
        for (int k = 0; k < a.n_ops && k < 64; ++k) if ((oa[k] & 0xf) == FO_I || (oa[k] & 0xf) == FO_D) { ++n_gap; break; }
        int mm, z;
        ps_census(q, ql, t, tl, p, &mm, &z);
        n_mm += mm; n_z += z;
        if (!ps_same(&a, oa, &b, ob, 64)) {
            ++n_div;
#pragma omp critical
            if (first < 0 || idx < first) first = idx;
        }
    }
    rep->n_diverged = n_div; rep->n_gapped = n_gap; rep->n_multi_max = n_mm; rep->n_zero_ef = n_z;
    rep->first_div = first;
    if (first >= 0) {
        int ql, tl;
        ps_make_pair(seed, first, qmax, tmax, rep->first_q, &ql, rep->first_t, &tl);
        rep->first_q[ql] = 0; rep->first_t[tl] = 0;
        rep->first_qlen = ql; rep->first_tlen = tl;
        fo_sw_result b;
        uint32_t ob[64];
        ps_sw_trace(rep->first_q, ql, rep->first_t, tl, p, lanes, &b, ob, 64);
        for (uint32_t sw = 1; sw <= FO_SW_WILD_MISMATCH; sw <<= 1) {
            fo_params p2 = *p;
            p2.switches ^= sw;
            fo_sw_result a;
            uint32_t oa[64];
            fo_sw_trace(rep->first_q, ql, rep->first_t, tl, &p2, &a, oa, 64);
            if (ps_same(&a, oa, &b, ob, 64)) { rep->explained_by = sw; break; }
        }
    }
    return 0;
}
