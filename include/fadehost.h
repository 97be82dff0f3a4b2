/*
 * fadehost.h -- host-side mirror of the parts of `fade annotate` that stay on the CPU around the
 * libfadegpu seam: what annotateTask does before and after align_clip.  Plain C ABI so that the
 * C++ harness, the Python tests and (optionally) the D host can share one implementation; the D
 * host may equally keep its own code for these few lines.
 *
 *   fadehost_parse_clips     <- parse_clips,                     source/util.d:37-62
 *   fadehost_aligned_length  <- dhtslib Cigar.alignedLength,     used at source/analysis.d:53
 *   fadehost_prepare         <- annotateTask early-outs + sc/sup bits, source/anno.d:61-74
 *   fadehost_finish          <- rs byte and am/as/ar/ab strings, source/anno.d:94-107,
 *                               source/analysis.d:82-92,106-118, source/readstatus.d:5-26
 */
#ifndef FADEHOST_H
#define FADEHOST_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ReadStatus bits, source/readstatus.d:5-26 / TAGS.md */
#define FADE_RS_SC 1u
#define FADE_RS_ART_LEFT 2u
#define FADE_RS_ART_RIGHT 4u
#define FADE_RS_MATE_LEFT 8u   /* never set by this version of fade (analysis.d:83) */
#define FADE_RS_MATE_RIGHT 16u /* never set (analysis.d:107) */
#define FADE_RS_SUP 32u

typedef struct fadehost_record {
    int32_t flag;          /* SAM FLAG; bit 0x4 = unmapped (rec.isMapped) */
    int32_t has_sa;        /* rec["SA"].exists */
    const uint32_t *cigar; /* BAM-encoded ops (len<<4|op) */
    int32_t n_cigar;
    const uint8_t *seq4;   /* BAM 4-bit packed bases */
    const uint8_t *qual;   /* raw phred values, l_qseq bytes */
    int32_t l_qseq;
    int32_t tid;
    int64_t pos;           /* 0-based */
} fadehost_record;

void fadehost_parse_clips(const uint32_t *cigar, int32_t n_cigar, uint32_t clips[2]);
int64_t fadehost_aligned_length(const uint32_t *cigar, int32_t n_cigar);

/* Before the device call.  Returns 1 when the record goes to the device (it has a soft clip and
 * is mapped) and 0 for the early-out records of anno.d:61-65 (rs = 0, clip outputs = 0).
 * rs_base receives the sc / sup bits. */
int fadehost_prepare(const fadehost_record *rec, int32_t *aligned_len, int32_t *clip_left,
                     int32_t *clip_right, uint8_t *rs_base);

/* After the device call.  `flags`, `win_start`, `beg_ref`, `n_ops`, `ops` are the read's outputs
 * of fadegpu_batch_view.  Writes the rs byte; when an artifact side was accepted also writes the
 * four NUL-terminated tag strings into the caller's buffers (each of capacity `cap`) and returns
 * 1; returns 0 when no am/as/ar/ab tags are to be written, -1 when `cap` is too small. */
int fadehost_finish(const fadehost_record *rec, const char *contig_name, uint8_t rs_base,
                    int32_t clip_left, int32_t clip_right, int32_t aligned_len,
                    uint8_t flags, int64_t win_start, int32_t beg_ref, int32_t n_ops,
                    const uint32_t *ops, uint8_t *rs_out,
                    char *am, char *as_, char *ar, char *ab, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
