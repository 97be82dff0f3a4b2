// fadehost.cpp -- host-side mirror of annotateTask's bookkeeping (include/fadehost.h).
// Reference: source/anno.d:55-110, source/analysis.d:82-92,106-118, source/util.d:37-62.
#include "../../include/fadehost.h"
#include <cstdio>
#include <cstring>
#include <string>

namespace {
const char kNt16[] = "=ACMGRSVTWYHKDBN";                     // htslib seq_nt16_str
const unsigned char kComp[16] = { 0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15 };  // util.d:18-21
const char kOps[] = "MIDNSHP=XB";
inline int nib(const uint8_t *s, int i) { return (s[i >> 1] >> ((~i & 1) << 2)) & 0xf; }

bool put(char *dst, size_t cap, const std::string &s)
{
    if (s.size() + 1 > cap) return false;
    memcpy(dst, s.c_str(), s.size() + 1);
    return true;
}
}  // namespace

extern "C" {

void fadehost_parse_clips(const uint32_t *cigar, int32_t n_cigar, uint32_t clips[2])
{
    clips[0] = clips[1] = 0;
    bool first = true;
    for (int32_t k = 0; k < n_cigar; ++k) {
        const uint32_t op = cigar[k] & 0xf;
        if (op == 5) continue;             // skip hard clips
        const bool sc = op == 4;
        if (first && !sc) first = false;
        else if (first && sc) clips[0] = cigar[k];
        else if (sc) clips[1] = cigar[k];
    }
}

int64_t fadehost_aligned_length(const uint32_t *cigar, int32_t n_cigar)
{
    int64_t s = 0;
    for (int32_t k = 0; k < n_cigar; ++k) {
        const uint32_t op = cigar[k] & 0xf;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) s += cigar[k] >> 4;  // M D N = X
    }
    return s;
}

int fadehost_prepare(const fadehost_record *rec, int32_t *aligned_len, int32_t *clip_left,
                     int32_t *clip_right, uint8_t *rs_base)
{
    *aligned_len = 0; *clip_left = 0; *clip_right = 0; *rs_base = 0;
    int n_s = 0;
    for (int32_t k = 0; k < rec->n_cigar; ++k) n_s += (rec->cigar[k] & 0xf) == 4;
    if ((rec->flag & 4) || n_s == 0) return 0;                       // anno.d:61-65
    uint32_t clips[2];
    fadehost_parse_clips(rec->cigar, rec->n_cigar, clips);            // anno.d:68
    uint8_t rs = 0;
    if ((clips[0] >> 4) != 0 || (clips[1] >> 4) != 0) rs |= FADE_RS_SC;  // anno.d:69-70
    if (rec->has_sa) rs |= FADE_RS_SUP;                               // anno.d:73-74
    *rs_base = rs;
    *clip_left = (int32_t)(clips[0] >> 4);
    *clip_right = (int32_t)(clips[1] >> 4);
    *aligned_len = (int32_t)fadehost_aligned_length(rec->cigar, rec->n_cigar);
    return 1;
}

int fadehost_finish(const fadehost_record *rec, const char *contig_name, uint8_t rs_base,
                    int32_t clip_left, int32_t clip_right, int32_t aligned_len,
                    uint8_t flags, int64_t win_start, int32_t beg_ref, int32_t n_ops,
                    const uint32_t *ops, uint8_t *rs_out,
                    char *am, char *as_, char *ar, char *ab, size_t cap)
{
    uint8_t rs = rs_base;
    if (flags & 2u) rs |= FADE_RS_ART_LEFT;    // analysis.d:82
    if (flags & 4u) rs |= FADE_RS_ART_RIGHT;   // analysis.d:106
    *rs_out = rs;                              // anno.d:94
    if (!(rs & (FADE_RS_ART_LEFT | FADE_RS_ART_RIGHT))) return 0;   // anno.d:98
    const int L = rec->l_qseq;
    std::string seq((size_t)L, 'N'), qrc((size_t)L, 'N'), bq((size_t)L, '!');
    for (int i = 0; i < L; ++i) {
        const int nb = nib(rec->seq4, i);
        seq[(size_t)i] = kNt16[nb];
        qrc[(size_t)(L - 1 - i)] = kNt16[kComp[nb]];               // util.d:23-34
        bq[(size_t)i] = (char)(rec->qual[i] + 33);                   // qscoresPhredScaled
    }
    // the result CIGAR and the values derived from it (valid: accepted implies n_ops <= 10)
    std::string cig;
    int64_t span = 0;
    for (int k = 0; k < n_ops; ++k) {
        char buf[24];
        snprintf(buf, sizeof buf, "%u%c", ops[k] >> 4, kOps[ops[k] & 0xf]);
        cig += buf;
        const uint32_t op = ops[k] & 0xf;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += ops[k] >> 4;
    }
    uint32_t rclips[2];
    fadehost_parse_clips(ops, n_ops, rclips);                        // analysis.d:78 / :102
    const int64_t apos = win_start + beg_ref;                        // start + res.position
    std::string aln = std::string(contig_name ? contig_name : "") + "," + std::to_string(apos) + "," + cig;
    std::string al[2], sl[2], slrc[2], q[2];
    if (rs & FADE_RS_ART_LEFT) {                                     // analysis.d:84-92
        al[0] = aln;
        const int64_t lim = rec->pos - (int64_t)(uint32_t)clip_left;
        const int64_t overlap = apos >= lim ? apos - lim : 0;
        int64_t plen = ((int64_t)L - (int64_t)(rclips[0] >> 4)) + overlap;
        if (plen > L) plen = L;
        if (plen < 0) plen = 0;
        sl[0] = seq.substr(0, (size_t)plen);
        slrc[0] = qrc.substr((size_t)(L - plen));
        q[0] = bq.substr(0, (size_t)plen);
    }
    if (rs & FADE_RS_ART_RIGHT) {                                    // analysis.d:108-118
        al[1] = aln;
        const int64_t a = rec->pos + aligned_len + (int64_t)(uint32_t)clip_right;
        const int64_t b = apos + span;
        const int64_t overlap = a >= b ? a - b : 0;
        int64_t plen = ((int64_t)L - (int64_t)(rclips[1] >> 4)) + overlap;
        if (plen > L) plen = L;
        if (plen < 0) plen = 0;
        sl[1] = seq.substr((size_t)(L - plen));
        slrc[1] = qrc.substr(0, (size_t)plen);
        q[1] = bq.substr((size_t)(L - plen));
    }
    // anno.d:100-106: both sides joined by ';', an absent side is the empty string
    if (!put(am, cap, al[0] + ";" + al[1]) || !put(as_, cap, sl[0] + ";" + sl[1]) ||
        !put(ar, cap, slrc[0] + ";" + slrc[1]) || !put(ab, cap, q[0] + ";" + q[1]))
        return -1;
    return 1;
}

}  // extern "C"
