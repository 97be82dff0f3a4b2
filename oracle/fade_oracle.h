/*
 * fade_oracle.h -- CPU restatement of `fade annotate`'s soft-clip realignment path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or the reported CPU baseline.
 *
 * PARITY UNPINNED: the reference (blachlylab/fade) ships no tests, golden vectors or
 * fixtures, and its arithmetic lives in parasail 2.4.3 / dparasail ~>0.3.3 / dhtslib@c51b842,
 * none of which is present in /root/reference or installable here (no D toolchain, no
 * network).  This file restates the published algorithm of those libraries (rules P1-P5 of
 * SURVEY.md section 8a) and is anchored on the reference's own call sites:
 *   source/analysis.d:22-124  align_clip (window, cutoff, SW call, accept predicate, tags)
 *   source/anno.d:55-110      annotateTask (early-outs, rs bits, tag emission)
 *   source/util.d:18-62       reverse_complement_sam_record, parse_clips
 *   source/readstatus.d:5-26  ReadStatus bit positions
 *   README.md:135-140,161-164 scoring (open 10, extend 2, match +2, mismatch -3), 90 % rule
 * The unverifiable points (U1-U8) are switchable through fo_params.switches.
 */
#ifndef FADE_ORACLE_H
#define FADE_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* BAM CIGAR op codes (SAM spec; dhtslib Ops) */
enum { FO_M = 0, FO_I = 1, FO_D = 2, FO_N = 3, FO_S = 4, FO_H = 5, FO_P = 6, FO_EQ = 7, FO_X = 8 };

/* switches: a set bit flips the default of the corresponding uncertainty */
enum {
    FO_SW_NO_SOFTCLIP_PAD = 1 << 0, /* U1: do not add S ops for unaligned query ends          */
    FO_SW_END_LAST_COL    = 1 << 1, /* U4: pick the LAST column / row among maxima            */
    FO_SW_E_BEFORE_F      = 1 << 2, /* U5: trace priority DIAG > E(D) > F(I)                  */
    FO_SW_GAP_TIE_OPEN    = 1 << 3, /* U5: gap open wins ties against gap extend              */
    FO_SW_EQ_BY_MATRIX    = 1 << 4, /* U7: '=' iff matrix score > 0 (instead of byte equality)*/
    FO_SW_SWAP_ID         = 1 << 5, /* U3: the gap along the target is written 'I', along the query 'D' */
    FO_SW_WILD_MISMATCH   = 1 << 6  /* P1: a letter outside "ACTGN" scores `mismatch` instead of 0    */
};
/* Also assumed, not switchable (they would change the shape of the outputs, not a tie-break):
 * U2 res.position == parasail beg_ref (0-based, window relative); U6 Cigar.alignedLength counts the
 * reference-consuming ops; ZERO has priority over every other source (U8 below). */
/* U8 (not switchable): a cell whose H is 0 always ends the traceback (ZERO has priority), even
 * when F or E is exactly 0 there.  parasail's lazy-F pass may label such a cell differently
 * depending on the SIMD lane layout; this cannot be modelled width-independently. */

typedef struct {
    int32_t gap_open;     /* 10  anno.d:36 */
    int32_t gap_extend;   /* 2   */
    int32_t match;        /* 2   */
    int32_t mismatch;     /* -3  */
    int32_t window_size;  /* 300 app.d:18 (align_buffer_size) */
    int32_t min_length;   /* 5   app.d:17 (artifact_floor_length) */
    uint32_t switches;
} fo_params;

void fo_default_params(fo_params *p);

typedef struct {
    int32_t score;
    int32_t end_query, end_ref;   /* 0-based inclusive, window relative */
    int32_t beg_query, beg_ref;   /* 0-based, window relative (res.position == beg_ref) */
    int32_t n_ops;                /* number of ops of res.cigar (with S padding, U1)     */
    int32_t ref_span;             /* res.cigar.alignedLength: '=' 'X' 'D' 'M' 'N' lengths */
} fo_sw_result;

/* P1-P5: parasail_sw_trace_striped_16 + parasail_result_get_cigar + dparasail wrapper.
 * q = query (rows), t = target (columns), both ASCII.  ops receives up to ops_cap BAM-encoded
 * ops (len<<4|op) in forward order; r->n_ops is the full count.  Returns 0, or -1 on bad args. */
int fo_sw_trace(const char *q, int qlen, const char *t, int tlen, const fo_params *p,
                fo_sw_result *r, uint32_t *ops, int ops_cap);

/* util.d:18-34: reverse complement of a BAM 4-bit packed sequence, ASCII out (no NUL). */
void fo_revcomp_nt16(const uint8_t *seq4, int l_qseq, char *out);
/* ASCII decode of a BAM 4-bit packed sequence (htslib seq_nt16_str). */
void fo_decode_nt16(const uint8_t *seq4, int l_qseq, char *out);
/* util.d:37-62 on BAM-encoded ops; clips[k] is the raw op word (0 = none). */
void fo_parse_clips(const uint32_t *cigar, int n_cigar, uint32_t clips[2]);
/* dhtslib Cigar.alignedLength: sum of reference-consuming op lengths (M D N = X). */
int64_t fo_cigar_ref_span(const uint32_t *cigar, int n_cigar);

/* ---- read-level: steps a-d of align_clip (analysis.d:34-80, 98-104), both sides at once ---- */
typedef struct {
    int32_t aligned;          /* 1 iff at least one side passed the length floor and SW ran */
    int32_t art_left, art_right;
    int64_t win_start;        /* `start` of analysis.d:45-51 */
    int32_t tlen;             /* end - start */
    fo_sw_result sw;
} fo_read_result;

/* ref_seq/ref_len: the contig the read maps to (any case; upper-cased like analysis.d:63). */
int fo_align_read(const uint8_t *seq4, int l_qseq, int64_t pos, int64_t aligned_len,
                  uint32_t clip_left, uint32_t clip_right,
                  const char *ref_seq, int64_t ref_len, const fo_params *p,
                  fo_read_result *r, uint32_t *ops, int ops_cap);

/* ---- record-level: annotateTask (anno.d:55-110) incl. tag strings ---- */
typedef struct {
    int32_t is_mapped;         /* !(flag & 4) */
    int32_t has_sa;            /* rec["SA"].exists */
    const uint32_t *cigar;     /* BAM-encoded */
    int32_t n_cigar;
    const uint8_t *seq4;       /* BAM packed bases */
    const uint8_t *qual;       /* raw phred, l_qseq bytes */
    int32_t l_qseq;
    int64_t pos;               /* 0-based */
    const char *contig_name;
    const char *ref_seq;
    int64_t ref_len;
} fo_record;

typedef struct {
    uint8_t rs;                /* ReadStatus.raw */
    int32_t has_tags;          /* 1 iff am/as/ar/ab are written (anno.d:98) */
    char *am, *as_, *ar, *ab;  /* malloc'd NUL-terminated strings when has_tags; else NULL */
} fo_tags;

int fo_annotate_record(const fo_record *rec, const fo_params *p, fo_tags *out);
void fo_free_tags(fo_tags *t);

/* batch helper (OpenMP when compiled with it): same as fo_align_read over n reads laid out as
 * struct-of-arrays; ops_out is n*ops_cap words.  Used by tests and the CPU baseline timing. */
int fo_align_batch(int64_t n, const uint8_t *seq4, const int64_t *seq_off, const int32_t *l_qseq,
                   const int32_t *tid, const int64_t *pos, const int32_t *aligned_len,
                   const int32_t *clip_left, const int32_t *clip_right,
                   int n_contigs, const char *const *contigs, const int64_t *contig_len,
                   const fo_params *p, fo_read_result *res, uint32_t *ops_out, int ops_cap,
                   int n_threads);

/* AVX2 / AVX-512BW inter-sequence implementation of the same contract (fade_oracle_simd.c): the CPU baseline
 * of bench.py.  Uses the widest of the two the host supports (fo_simd_lanes(): 32, 16, or 0 = neither, in which
 * case it falls back to fo_align_batch); FADE_ORACLE_SIMD=avx2 in the environment forces 16 lanes. */
int fo_simd_lanes(void);
int fo_align_batch_simd(int64_t n, const uint8_t *seq4, const int64_t *seq_off, const int32_t *l_qseq,
                        const int32_t *tid, const int64_t *pos, const int32_t *aligned_len,
                        const int32_t *clip_left, const int32_t *clip_right,
                        int n_contigs, const char *const *contigs, const int64_t *contig_len,
                        const fo_params *p, fo_read_result *res, uint32_t *ops_out, int ops_cap,
                        int n_threads);

#ifdef __cplusplus
}
#endif
#endif
