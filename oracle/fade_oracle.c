/*
 * fade_oracle.c -- scalar CPU restatement of the `fade annotate` realignment path.
 * TEST INFRASTRUCTURE ONLY (see fade_oracle.h).  PARITY UNPINNED (no reference vectors exist).
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * The Smith-Waterman itself restates parasail 2.4.3 (Dockerfile:4) as reached through
 * dparasail's Parasail.sw_striped (source/analysis.d:67): rules P1-P5 of SURVEY.md 8(a).
 */
#include "fade_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void fo_default_params(fo_params *p)
{
    p->gap_open = 10;  /* source/anno.d:36  Parasail("ACTGN", 10, 2, 2, -3) */
    p->gap_extend = 2;
    p->match = 2;
    p->mismatch = -3;
    p->window_size = 300; /* source/app.d:18 */
    p->min_length = 5;    /* source/app.d:17 */
    p->switches = 0;
}

/* P1: parasail_matrix_create("ACTGN", match, mismatch) -- call site source/anno.d:36.
 * index 0..4 = A C T G N (either case), 5 = wildcard for every other byte. */
static inline int fo_map(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'T': case 't': return 2;
    case 'G': case 'g': return 3;
    case 'N': case 'n': return 4;
    default: return 5;
    }
}

static inline int fo_sub(const fo_params *p, int a, int b)
{
    if (a == 5 || b == 5) return (p->switches & FO_SW_WILD_MISMATCH) ? p->mismatch : 0;
    return a == b ? p->match : p->mismatch;
}

static inline int fo_upper(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

/* trace byte: bits 0-1 source of H, bit 2 = E came from "open", bit 3 = F came from "open" */
enum { T_ZERO = 0, T_DIAG = 1, T_F = 2, T_E = 3, T_EOPEN = 4, T_FOPEN = 8 };

/*
 * P2 (fill), P3 (end cell), P4 (traceback + CIGAR), P5 (dparasail result wrapper).
 * Reference call site: source/analysis.d:67 `p.sw_striped(q_seq, ref_seq)`; fields consumed at
 * source/analysis.d:69-85,98-111 (res.cigar, res.score, res.position).
 */
int fo_sw_trace(const char *q, int qlen, const char *t, int tlen, const fo_params *p,
                fo_sw_result *r, uint32_t *ops, int ops_cap)
{
    if (!q || !t || !p || !r || qlen <= 0 || tlen <= 0) return -1;
    const int o = p->gap_open, e = p->gap_extend;
    const uint32_t sw = p->switches;
    const int NEG = -(1 << 28);

    int *Hcol = (int *)calloc((size_t)qlen, sizeof(int));  /* H[i][j-1] then H[i][j] */
    int *E = (int *)malloc((size_t)qlen * sizeof(int));    /* E[i][j] for the current column */
    int *Hbest = (int *)calloc((size_t)qlen, sizeof(int)); /* copy of the best column (P3) */
    uint8_t *tr = (uint8_t *)malloc((size_t)qlen * (size_t)tlen);
    uint8_t *qi = (uint8_t *)malloc((size_t)qlen);
    if (!Hcol || !E || !Hbest || !tr || !qi) {
        free(Hcol); free(E); free(Hbest); free(tr); free(qi);
        return -1;
    }
    for (int i = 0; i < qlen; ++i) { qi[i] = (uint8_t)fo_map((unsigned char)q[i]); E[i] = NEG; }

    int score = 0, end_ref = -1;
    for (int j = 0; j < tlen; ++j) {
        const int tj = fo_map((unsigned char)t[j]);
        int hdiag = 0;      /* H[i-1][j-1], boundary 0 */
        int hup = 0;        /* H[i-1][j],  boundary 0 */
        int f = NEG;        /* F[i-1][j] */
        int colmax = 0;
        for (int i = 0; i < qlen; ++i) {
            const int hleft = Hcol[i];   /* H[i][j-1] (0 for j==0 thanks to calloc) */
            /* E[i][j] = max(H[i][j-1]-o, E[i][j-1]-e): gap that consumes TARGET ('D') */
            const int e_opn = hleft - o, e_ext = E[i] - e;
            uint8_t tb = 0;
            int ev;
            if ((sw & FO_SW_GAP_TIE_OPEN) ? (e_opn >= e_ext) : (e_opn > e_ext)) { ev = e_opn; tb |= T_EOPEN; }
            else ev = e_ext;
            /* F[i][j] = max(H[i-1][j]-o, F[i-1][j]-e): gap that consumes QUERY ('I') */
            const int f_opn = hup - o, f_ext = f - e;
            int fv;
            if ((sw & FO_SW_GAP_TIE_OPEN) ? (f_opn >= f_ext) : (f_opn > f_ext)) { fv = f_opn; tb |= T_FOPEN; }
            else fv = f_ext;
            int hd = hdiag + fo_sub(p, qi[i], tj);
            if (hd < 0) hd = 0;
            int h = hd;
            if (ev > h) h = ev;
            if (fv > h) h = fv;
            /* P4 source priority: DIAG (ZERO when the value is 0) > F > E */
            if (h == hd) tb |= (h == 0) ? T_ZERO : T_DIAG;
            else if (sw & FO_SW_E_BEFORE_F) tb |= (h == ev) ? T_E : T_F;
            else tb |= (h == fv) ? T_F : T_E;
            tr[(size_t)i * tlen + j] = tb;
            hdiag = hleft;
            hup = h;
            f = fv;
            E[i] = ev;
            Hcol[i] = h;
            if (h > colmax) colmax = h;
        }
        /* P3: the column index is taken the first time the running maximum strictly grows */
        if ((sw & FO_SW_END_LAST_COL) ? (colmax >= score && colmax > 0) : (colmax > score)) {
            score = colmax;
            end_ref = j;
            memcpy(Hbest, Hcol, (size_t)qlen * sizeof(int));
        }
    }

    memset(r, 0, sizeof(*r));
    r->score = score;
    if (score <= 0 || end_ref < 0) {
        /* nothing aligned: res.cigar.length == 0 -> rejected at source/analysis.d:69 */
        r->end_query = r->end_ref = r->beg_query = r->beg_ref = 0;
        r->n_ops = 0;
        free(Hcol); free(E); free(Hbest); free(tr); free(qi);
        return 0;
    }
    /* P3: smallest row holding the score in that column */
    int end_query = qlen - 1;
    if (sw & FO_SW_END_LAST_COL) {
        for (int i = 0; i < qlen; ++i) if (Hbest[i] == score) end_query = i;
    } else {
        for (int i = 0; i < qlen; ++i) if (Hbest[i] == score && i < end_query) end_query = i;
    }

    /* P4: traceback */
    int cap = 2 * (qlen + tlen) + 4;
    uint32_t *rev = (uint32_t *)malloc((size_t)cap * sizeof(uint32_t));
    int nrev = 0;
    int i = end_query, j = end_ref;
    int state = 0; /* 0 = H, 1 = F (emits I), 2 = E (emits D) */
    int ref_span = 0;
#define PUSH(op_)                                                              \
    do {                                                                       \
        if (nrev > 0 && (rev[nrev - 1] & 0xf) == (uint32_t)(op_)) rev[nrev - 1] += 16; \
        else rev[nrev++] = (1u << 4) | (uint32_t)(op_);                        \
    } while (0)
    while (i >= 0 && j >= 0) {
        const uint8_t tb = tr[(size_t)i * tlen + j];
        if (state == 0) {
            const int src = tb & 3;
            if (src == T_ZERO) break;
            if (src == T_DIAG) {
                int eq;
                if (sw & FO_SW_EQ_BY_MATRIX) eq = fo_sub(p, qi[i], fo_map((unsigned char)t[j])) > 0;
                else eq = fo_upper((unsigned char)q[i]) == fo_upper((unsigned char)t[j]);
                PUSH(eq ? FO_EQ : FO_X);
                --i; --j;
            } else if (src == T_F) state = 1;
            else state = 2;
        } else if (state == 1) {
            PUSH((sw & FO_SW_SWAP_ID) ? FO_D : FO_I);
            --i;
            if (tb & T_FOPEN) state = 0;
        } else {
            PUSH((sw & FO_SW_SWAP_ID) ? FO_I : FO_D);
            --j;
            if (tb & T_EOPEN) state = 0;
        }
    }
#undef PUSH
    const int beg_query = i + 1, beg_ref = j + 1;
    /* res.cigar.alignedLength as dhtslib computes it: by op letter (= X D), source/analysis.d:110-113 */
    for (int k = 0; k < nrev; ++k) {
        const uint32_t op = rev[k] & 0xf;
        if (op == FO_EQ || op == FO_X || op == FO_D) ref_span += (int)(rev[k] >> 4);
    }

    /* P5: dparasail result wrapper adds S ops for the unaligned query ends (U1) */
    int n = 0;
    const int lead = beg_query, trail = qlen - 1 - end_query;
    const int pad = !(sw & FO_SW_NO_SOFTCLIP_PAD);
    if (pad && lead > 0) { if (n < ops_cap && ops) ops[n] = ((uint32_t)lead << 4) | FO_S; ++n; }
    for (int k = nrev - 1; k >= 0; --k) { if (n < ops_cap && ops) ops[n] = rev[k]; ++n; }
    if (pad && trail > 0) { if (n < ops_cap && ops) ops[n] = ((uint32_t)trail << 4) | FO_S; ++n; }

    r->end_query = end_query;
    r->end_ref = end_ref;
    r->beg_query = beg_query;
    r->beg_ref = beg_ref;
    r->n_ops = n;
    r->ref_span = ref_span;
    free(rev); free(Hcol); free(E); free(Hbest); free(tr); free(qi);
    return 0;
}

/* source/util.d:18-21 seq_comp_table and htslib's seq_nt16_str */
static const uint8_t fo_comp16[16] = { 0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15 };
static const char fo_nt16_str[] = "=ACMGRSVTWYHKDBN";

/* source/util.d:23-34 */
void fo_revcomp_nt16(const uint8_t *seq4, int l_qseq, char *out)
{
    int j = l_qseq - 1;
    for (int i = 0; i < l_qseq; ++i) {
        const int nib = (seq4[i >> 1] >> ((~i & 1) << 2)) & 0xf;
        out[j--] = fo_nt16_str[fo_comp16[nib]];
    }
}

void fo_decode_nt16(const uint8_t *seq4, int l_qseq, char *out)
{
    for (int i = 0; i < l_qseq; ++i) out[i] = fo_nt16_str[(seq4[i >> 1] >> ((~i & 1) << 2)) & 0xf];
}

/* source/util.d:37-62 */
void fo_parse_clips(const uint32_t *cigar, int n_cigar, uint32_t clips[2])
{
    clips[0] = clips[1] = 0;
    int first = 1;
    for (int k = 0; k < n_cigar; ++k) {
        const uint32_t op = cigar[k] & 0xf;
        if (op == FO_H) continue;
        const int is_sc = (op == FO_S);
        if (first && !is_sc) first = 0;
        else if (first && is_sc) clips[0] = cigar[k];
        else if (is_sc) clips[1] = cigar[k];
    }
}

/* dhtslib Cigar.alignedLength as used at source/analysis.d:53,110-113 */
int64_t fo_cigar_ref_span(const uint32_t *cigar, int n_cigar)
{
    int64_t s = 0;
    for (int k = 0; k < n_cigar; ++k) {
        const uint32_t op = cigar[k] & 0xf;
        if (op == FO_M || op == FO_D || op == FO_N || op == FO_EQ || op == FO_X) s += cigar[k] >> 4;
    }
    return s;
}

/* the accept predicate of source/analysis.d:69-80 (left) and :98-104 (right) */
static int fo_accept(int left, const fo_sw_result *sw, const uint32_t *ops, int n_have,
                     uint32_t clip_len)
{
    if (sw->n_ops == 0 || sw->n_ops > 10) return 0; /* analysis.d:69-70 */
    if (n_have < sw->n_ops) return 0;                /* cannot happen: ops_cap >= 10 enforced */
    const uint32_t edge = left ? ops[sw->n_ops - 1] : ops[0];
    if ((edge & 0xf) != FO_EQ) return 0;             /* analysis.d:74 / :98 */
    const float cutoff = (float)(clip_len * 0.9 * 2); /* analysis.d:43 (float cutoff) */
    if (!((float)sw->score > cutoff)) return 0;      /* analysis.d:76 / :100 */
    uint32_t clips[2];
    fo_parse_clips(ops, sw->n_ops, clips);           /* analysis.d:78 / :102 */
    if (left) return !((clips[1] >> 4) != 0 || (clips[0] >> 4) == 0);  /* analysis.d:79 */
    return !((clips[0] >> 4) != 0 || (clips[1] >> 4) == 0);            /* analysis.d:103 */
}

int fo_align_read(const uint8_t *seq4, int l_qseq, int64_t pos, int64_t aligned_len,
                  uint32_t clip_left, uint32_t clip_right,
                  const char *ref_seq, int64_t ref_len, const fo_params *p,
                  fo_read_result *r, uint32_t *ops, int ops_cap)
{
    memset(r, 0, sizeof(*r));
    if (ops_cap < 10) return -1;
    /* analysis.d:34  `clip_len <= artifact_floor_length` is uint <= int -> unsigned compare */
    const int do_left = clip_left != 0 && !(clip_left <= (uint32_t)p->min_length);
    const int do_right = clip_right != 0 && !(clip_right <= (uint32_t)p->min_length);
    if (!do_left && !do_right) return 0;
    /* analysis.d:45-59 */
    int64_t start = pos - p->window_size;
    if (start < 0) start = 0;
    int64_t end = pos + aligned_len + p->window_size;
    if (end > ref_len) end = ref_len;
    r->win_start = start;
    if (end <= start || l_qseq <= 0) return 0; /* empty window: nothing to align */
    const int tlen = (int)(end - start);
    r->tlen = tlen;
    char *q = (char *)malloc((size_t)l_qseq);
    char *t = (char *)malloc((size_t)tlen);
    fo_revcomp_nt16(seq4, l_qseq, q);                            /* analysis.d:40 */
    for (int k = 0; k < tlen; ++k) t[k] = (char)fo_upper((unsigned char)ref_seq[start + k]); /* :63 */
    int rc = fo_sw_trace(q, l_qseq, t, tlen, p, &r->sw, ops, ops_cap);   /* :67 */
    free(q); free(t);
    if (rc) return rc;
    r->aligned = 1;
    const int have = r->sw.n_ops < ops_cap ? r->sw.n_ops : ops_cap;
    if (do_left) r->art_left = fo_accept(1, &r->sw, ops, have, clip_left);
    if (do_right) r->art_right = fo_accept(0, &r->sw, ops, have, clip_right);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */

static char *fo_cigar_string(const uint32_t *ops, int n)
{
    static const char opc[] = "MIDNSHP=XB";
    char *s = (char *)malloc((size_t)n * 12 + 1);
    size_t w = 0;
    for (int k = 0; k < n; ++k) w += (size_t)sprintf(s + w, "%u%c", ops[k] >> 4, opc[ops[k] & 0xf]);
    s[w] = 0;
    return s;
}

static char *fo_substr(const char *s, int64_t a, int64_t b)
{
    if (b < a) b = a;
    char *o = (char *)malloc((size_t)(b - a) + 1);
    memcpy(o, s + a, (size_t)(b - a));
    o[b - a] = 0;
    return o;
}

static char *fo_join(const char *a, const char *b)
{
    const size_t la = a ? strlen(a) : 0, lb = b ? strlen(b) : 0;
    char *o = (char *)malloc(la + lb + 2);
    if (la) memcpy(o, a, la);
    o[la] = ';';
    if (lb) memcpy(o + la + 1, b, lb);
    o[la + 1 + lb] = 0;
    return o;
}

/* source/anno.d:55-110 with align_clip (source/analysis.d:22-124) inlined for both sides */
int fo_annotate_record(const fo_record *rec, const fo_params *p, fo_tags *out)
{
    memset(out, 0, sizeof(*out));
    int n_s = 0;
    for (int k = 0; k < rec->n_cigar; ++k) n_s += (rec->cigar[k] & 0xf) == FO_S;
    if (!rec->is_mapped || n_s == 0) return 0;       /* anno.d:61-65: rs = 0 */
    uint32_t clips[2];
    fo_parse_clips(rec->cigar, rec->n_cigar, clips);  /* anno.d:68 */
    uint8_t rs = 0;
    if ((clips[0] >> 4) != 0 || (clips[1] >> 4) != 0) rs |= 1;   /* sc,  anno.d:69-70 */
    if (rec->has_sa) rs |= 32;                                   /* sup, anno.d:73-74 */
    const int64_t A = fo_cigar_ref_span(rec->cigar, rec->n_cigar);
    const int L = rec->l_qseq;
    char *seq = (char *)malloc((size_t)L + 1), *qrc = (char *)malloc((size_t)L + 1);
    char *bq = (char *)malloc((size_t)L + 1);
    fo_decode_nt16(rec->seq4, L, seq); seq[L] = 0;
    fo_revcomp_nt16(rec->seq4, L, qrc); qrc[L] = 0;
    for (int k = 0; k < L; ++k) bq[k] = (char)(rec->qual[k] + 33);
    bq[L] = 0;

    char *al[2] = { NULL, NULL }, *sl[2] = { NULL, NULL }, *slrc[2] = { NULL, NULL }, *q[2] = { NULL, NULL };
    for (int side = 0; side < 2; ++side) {
        const uint32_t clip_len = clips[side] >> 4;
        if (clip_len == 0) continue;                 /* anno.d:79 / :87 */
        fo_read_result rr;
        uint32_t ops[16];
        /* one side at a time, exactly like the two align_clip calls */
        int rc = fo_align_read(rec->seq4, L, rec->pos, A, side == 0 ? clip_len : 0,
                               side == 1 ? clip_len : 0, rec->ref_seq, rec->ref_len, p, &rr, ops, 16);
        if (rc) { free(seq); free(qrc); free(bq); return rc; }
        const int art = side == 0 ? rr.art_left : rr.art_right;
        if (!art) continue;
        rs |= side == 0 ? 2 : 4;                     /* analysis.d:82 / :106 */
        uint32_t rclips[2];
        fo_parse_clips(ops, rr.sw.n_ops, rclips);
        char *cs = fo_cigar_string(ops, rr.sw.n_ops);
        const int64_t apos = rr.win_start + rr.sw.beg_ref;
        al[side] = (char *)malloc(strlen(rec->contig_name) + strlen(cs) + 32);
        sprintf(al[side], "%s,%lld,%s", rec->contig_name, (long long)apos, cs);  /* :84-85 */
        free(cs);
        int64_t plen;
        if (side == 0) {
            /* analysis.d:86-92 */
            const int64_t lim = rec->pos - (int64_t)clip_len;
            const int64_t overlap = apos >= lim ? apos - lim : 0;
            plen = ((int64_t)L - (int64_t)(rclips[0] >> 4)) + overlap;
            if (plen > L) plen = L;
            if (plen < 0) plen = 0;
            sl[side] = fo_substr(seq, 0, plen);
            slrc[side] = fo_substr(qrc, L - plen, L);
            q[side] = fo_substr(bq, 0, plen);
        } else {
            /* analysis.d:110-118 */
            const int64_t a = rec->pos + A + (int64_t)clip_len;
            const int64_t b = apos + rr.sw.ref_span;
            const int64_t overlap = a >= b ? a - b : 0;
            plen = ((int64_t)L - (int64_t)(rclips[1] >> 4)) + overlap;
            if (plen > L) plen = L;
            if (plen < 0) plen = 0;
            sl[side] = fo_substr(seq, L - plen, L);
            slrc[side] = fo_substr(qrc, 0, plen);
            q[side] = fo_substr(bq, L - plen, L);
        }
    }
    out->rs = rs;                                    /* anno.d:94 */
    if (rs & 6) {                                    /* anno.d:98-107 */
        out->has_tags = 1;
        out->am = fo_join(al[0], al[1]);
        out->as_ = fo_join(sl[0], sl[1]);
        out->ar = fo_join(slrc[0], slrc[1]);
        out->ab = fo_join(q[0], q[1]);
    }
    for (int s = 0; s < 2; ++s) { free(al[s]); free(sl[s]); free(slrc[s]); free(q[s]); }
    free(seq); free(qrc); free(bq);
    return 0;
}

void fo_free_tags(fo_tags *t)
{
    free(t->am); free(t->as_); free(t->ar); free(t->ab);
    t->am = t->as_ = t->ar = t->ab = NULL;
}

int fo_align_batch(int64_t n, const uint8_t *seq4, const int64_t *seq_off, const int32_t *l_qseq,
                   const int32_t *tid, const int64_t *pos, const int32_t *aligned_len,
                   const int32_t *clip_left, const int32_t *clip_right,
                   int n_contigs, const char *const *contigs, const int64_t *contig_len,
                   const fo_params *p, fo_read_result *res, uint32_t *ops_out, int ops_cap,
                   int n_threads)
{
    int err = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < n; ++k) {
        fo_read_result *r = &res[k];
        memset(r, 0, sizeof(*r));
        if (tid[k] < 0 || tid[k] >= n_contigs) continue;
        int rc = fo_align_read(seq4 + seq_off[k], l_qseq[k], pos[k], aligned_len[k],
                               (uint32_t)clip_left[k], (uint32_t)clip_right[k], contigs[tid[k]],
                               contig_len[tid[k]], p, r, ops_out + (size_t)k * ops_cap, ops_cap);
        if (rc) {
#pragma omp atomic write
            err = rc;
        }
    }
    return err;
}
