// standin_device.cpp -- TEST INFRASTRUCTURE ONLY: a stand-in for the GPU side of the C ABI, so that the HOST driver
// (`fade-b200 annotate`: record parsing, the compact batch layout, the ring of batches over several GPUs, tag
// assembly, the writers) can be exercised by the CPU test-suite.  It is built by tests/test_cli_host_driver.py into a
// temporary directory and LD_PRELOADed into the driver by that test alone; nothing in fade_b200/, bench.py or
// __graft_entry__.py builds, links or loads it, and the product has no CPU path: libfadegpu.so fails with
// FADEGPU_E_NODEV without a CUDA device (tests/test_abi_and_host.py::test_no_gpu_means_loud_failure_not_fallback).
// The "device" work is done by the oracle (oracle/fade_oracle.h), i.e. the driver's output is checked against the
// oracle end to end -- the same comparison tests/test_gpu_cli.py makes on a B200, minus the kernels.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <memory>
#include "fadegpu.h"
#include "fade_oracle.h"

struct StandinRef {
    std::vector<std::string> seqs;
    std::vector<int64_t> lens;
};

struct fadegpu_ctx {
    fadegpu_params prm;
    std::shared_ptr<StandinRef> ref;
    int device;
};

struct fadegpu_batch {
    fadegpu_ctx *ctx;
    int64_t max_reads, max_seq;
    std::vector<uint8_t> seq4, flags, gate;
    std::vector<int64_t> seq_off, pos, win_start_pr;
    std::vector<int32_t> l_qseq, tid, aligned_len, clip_left, clip_right, n_ops_pr, beg_ref_pr, score, begq, endq, endr;
    std::vector<uint32_t> ops_pr;
    std::vector<fadegpu_read_meta> meta;
    // results of the last submit
    std::vector<fadegpu_result> results;
    std::vector<int64_t> win_start;
    std::vector<int32_t> result_index;
    int64_t n_reads = 0;
    bool pending = false, compact = false;
    int64_t seq_bytes = 0;
    fadegpu_stats st;
};

static thread_local std::string g_err;
static int fail(int code, const char *msg) { g_err = msg; return code; }
static int devices()
{
    const char *e = getenv("FADE_STANDIN_DEVICES");
    return e ? atoi(e) : 1;
}

extern "C" {

int fadegpu_device_count(int *n) { if (!n) return fail(FADEGPU_E_ARG, "null"); *n = devices(); return 0; }

int fadegpu_create(int device, const fadegpu_params *p, fadegpu_ctx **out)
{
    if (!p || !out || device < 0 || device >= devices()) return fail(FADEGPU_E_NODEV, "stand-in: no such device");
    *out = new fadegpu_ctx{ *p, nullptr, device };
    return 0;
}
void fadegpu_destroy(fadegpu_ctx *c) { delete c; }
const char *fadegpu_last_error(const fadegpu_ctx *) { return g_err.c_str(); }

int fadegpu_load_reference(fadegpu_ctx *c, int32_t n, const char *const *, const int64_t *lengths, const char *const *seqs)
{
    if (!c || n < 0) return fail(FADEGPU_E_ARG, "stand-in: bad reference");
    auto r = std::make_shared<StandinRef>();
    for (int32_t k = 0; k < n; ++k) { r->seqs.emplace_back(seqs[k], (size_t)lengths[k]); r->lens.push_back(lengths[k]); }
    c->ref = r;
    return 0;
}
int fadegpu_share_reference(fadegpu_ctx *dst, const fadegpu_ctx *src)
{
    if (!dst || !src || !src->ref) return fail(FADEGPU_E_STATE, "stand-in: source has no reference");
    dst->ref = src->ref;
    return 0;
}

int fadegpu_alloc_batch(fadegpu_ctx *c, int64_t max_reads, int64_t max_seq, fadegpu_batch **out)
{
    if (!c || !out || max_reads <= 0 || max_seq <= 0) return fail(FADEGPU_E_ARG, "stand-in: bad batch size");
    auto *b = new fadegpu_batch();
    b->ctx = c; b->max_reads = max_reads; b->max_seq = max_seq;
    const size_t n = (size_t)max_reads;
    b->seq4.resize((size_t)max_seq + 16); b->flags.resize(n); b->gate.resize(n); b->meta.resize(n);
    b->seq_off.resize(n + 1); b->pos.resize(n); b->l_qseq.resize(n); b->tid.resize(n); b->aligned_len.resize(n);
    b->clip_left.resize(n); b->clip_right.resize(n);
    if (!(c->prm.flags & FADEGPU_F_NO_SCATTER)) {
        b->win_start_pr.resize(n); b->n_ops_pr.resize(n); b->beg_ref_pr.resize(n); b->score.resize(n); b->begq.resize(n);
        b->endq.resize(n); b->endr.resize(n); b->ops_pr.resize(n * FADEGPU_MAX_OPS);
    }
    memset(&b->st, 0, sizeof(b->st));
    *out = b;
    return 0;
}
void fadegpu_free_batch(fadegpu_batch *b) { delete b; }

int fadegpu_get_batch_view(fadegpu_batch *b, fadegpu_batch_view *v)
{
    if (!b || !v) return fail(FADEGPU_E_ARG, "null");
    memset(v, 0, sizeof(*v));
    v->max_reads = b->max_reads; v->max_seq_bytes = b->max_seq;
    v->seq4 = b->seq4.data(); v->seq_off = b->seq_off.data(); v->l_qseq = b->l_qseq.data(); v->tid = b->tid.data();
    v->pos = b->pos.data(); v->aligned_len = b->aligned_len.data(); v->clip_left = b->clip_left.data();
    v->clip_right = b->clip_right.data(); v->flags = b->flags.data(); v->gate = b->gate.data(); v->meta = b->meta.data();
    if (!(b->ctx->prm.flags & FADEGPU_F_NO_SCATTER)) {
        v->score = b->score.data(); v->beg_query = b->begq.data(); v->end_query = b->endq.data(); v->beg_ref = b->beg_ref_pr.data();
        v->end_ref = b->endr.data(); v->win_start = b->win_start_pr.data(); v->n_ops = b->n_ops_pr.data(); v->ops = b->ops_pr.data();
    }
    return 0;
}

static int queue(fadegpu_ctx *c, fadegpu_batch *b, int64_t n, bool compact, int64_t seq_bytes)
{
    if (!c || !b || b->ctx != c || n < 0 || n > b->max_reads) return fail(FADEGPU_E_ARG, "stand-in: bad submit");
    if (!c->ref) return fail(FADEGPU_E_STATE, "stand-in: no reference loaded");
    if (b->pending) return fail(FADEGPU_E_STATE, "stand-in: batch in flight");
    b->n_reads = n; b->pending = true; b->compact = compact; b->seq_bytes = seq_bytes;
    return 0;
}
int fadegpu_submit(fadegpu_ctx *c, fadegpu_batch *b, int64_t n) { return queue(c, b, n, false, 0); }
int fadegpu_submit_compact(fadegpu_ctx *c, fadegpu_batch *b, int64_t n, int64_t seq_bytes)
{
    if (seq_bytes < 0 || (b && seq_bytes > b->max_seq)) return fail(FADEGPU_E_ARG, "stand-in: seq_bytes out of range");
    return queue(c, b, n, true, seq_bytes);
}

// the work happens here (the real library does it between submit and wait, on the GPU)
int fadegpu_wait(fadegpu_ctx *c, fadegpu_batch *b)
{
    if (!c || !b || !b->pending) return fail(FADEGPU_E_STATE, "stand-in: nothing submitted");
    b->pending = false;
    const int64_t n = b->n_reads;
    if (b->compact) {   // unpack the compact layout into the seven arrays
        for (int64_t k = 0; k < n; ++k) {
            const fadegpu_read_meta &m = b->meta[(size_t)k];
            if ((int64_t)m.seq_off + (m.l_qseq + 1) / 2 > b->seq_bytes) return fail(FADEGPU_E_ARG, "stand-in: seq_off outside seq_bytes");
            if (b->gate[(size_t)k] != (uint8_t)std::min<uint32_t>(255u, std::max(m.clip_left, m.clip_right)))
                return fail(FADEGPU_E_ARG, "stand-in: gate byte does not match the record");
            b->seq_off[(size_t)k] = m.seq_off; b->l_qseq[(size_t)k] = m.l_qseq; b->tid[(size_t)k] = m.tid; b->pos[(size_t)k] = m.pos;
            b->aligned_len[(size_t)k] = m.aligned_len; b->clip_left[(size_t)k] = (int32_t)m.clip_left;
            b->clip_right[(size_t)k] = (int32_t)m.clip_right;
        }
    }
    if (getenv("FADE_STANDIN_NO_COMPUTE")) {   // host-pipeline timing: every read comes back unaligned
        std::fill(b->flags.begin(), b->flags.begin() + n, (uint8_t)0);
        b->results.clear(); b->win_start.clear();
        b->result_index.assign((size_t)n, -1);
        return 0;
    }
    fo_params p;
    fo_default_params(&p);
    p.gap_open = c->prm.gap_open; p.gap_extend = c->prm.gap_extend; p.match = c->prm.match; p.mismatch = c->prm.mismatch;
    p.window_size = c->prm.window_size; p.min_length = c->prm.min_length;
    std::vector<const char *> contigs;
    for (auto &s : c->ref->seqs) contigs.push_back(s.data());
    std::vector<fo_read_result> res((size_t)std::max<int64_t>(n, 1));
    std::vector<uint32_t> ops((size_t)std::max<int64_t>(n, 1) * FADEGPU_MAX_OPS);
    if (fo_align_batch(n, b->seq4.data(), b->seq_off.data(), b->l_qseq.data(), b->tid.data(), b->pos.data(), b->aligned_len.data(),
                       b->clip_left.data(), b->clip_right.data(), (int)contigs.size(), contigs.data(), c->ref->lens.data(), &p,
                       res.data(), ops.data(), FADEGPU_MAX_OPS, 0) != 0)
        return fail(FADEGPU_E_ARG, "stand-in: oracle rejected the batch");
    b->results.clear(); b->win_start.clear();
    b->result_index.assign((size_t)n, -1);
    const bool scatter = !(c->prm.flags & FADEGPU_F_NO_SCATTER);
    for (int64_t k = n - 1; k >= 0; --k) {   // records in REVERSE read order: callers must go through result_index
        const fo_read_result &r = res[(size_t)k];
        uint8_t fl = 0;
        if (r.aligned) {
            fl = FADEGPU_R_ALIGNED | (r.art_left ? FADEGPU_R_ART_LEFT : 0) | (r.art_right ? FADEGPU_R_ART_RIGHT : 0) |
                 (r.sw.n_ops > FADEGPU_MAX_OPS ? FADEGPU_R_OPS_TRUNC : 0);
            fadegpu_result o;
            memset(&o, 0, sizeof(o));
            o.score = r.sw.score; o.end_query = r.sw.end_query; o.end_ref = r.sw.end_ref; o.beg_query = r.sw.beg_query;
            o.beg_ref = r.sw.beg_ref; o.n_ops = r.sw.n_ops; o.flags = fl; o.read = (int32_t)k;
            for (int x = 0; x < FADEGPU_MAX_OPS && x < r.sw.n_ops; ++x) o.ops[x] = ops[(size_t)k * FADEGPU_MAX_OPS + (size_t)x];
            b->result_index[(size_t)k] = (int32_t)b->results.size();
            b->results.push_back(o);
            b->win_start.push_back(r.win_start);
            if (scatter) {
                b->score[(size_t)k] = o.score; b->begq[(size_t)k] = o.beg_query; b->endq[(size_t)k] = o.end_query;
                b->beg_ref_pr[(size_t)k] = o.beg_ref; b->endr[(size_t)k] = o.end_ref; b->win_start_pr[(size_t)k] = r.win_start;
                b->n_ops_pr[(size_t)k] = o.n_ops;
                memcpy(&b->ops_pr[(size_t)k * FADEGPU_MAX_OPS], o.ops, sizeof(o.ops));
            }
        }
        b->flags[(size_t)k] = fl;
    }
    memset(&b->st, 0, sizeof(b->st));
    b->st.n_reads = n; b->st.n_aligned = (int64_t)b->results.size();
    return 0;
}

int fadegpu_get_results(const fadegpu_batch *b, fadegpu_results_view *r)
{
    if (!b || !r) return fail(FADEGPU_E_ARG, "null");
    r->n_results = (int64_t)b->results.size(); r->results = b->results.data(); r->win_start = b->win_start.data();
    r->result_index = b->result_index.data();
    return 0;
}
int fadegpu_get_stats(const fadegpu_batch *b, fadegpu_stats *s) { if (!b || !s) return fail(FADEGPU_E_ARG, "null"); *s = b->st; return 0; }

}   // extern "C"
