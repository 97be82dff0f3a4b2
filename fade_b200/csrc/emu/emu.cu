// emu.cu -- HOST lock-step emulation of one 8-thread group of the realignment kernels.
//
// TEST INFRASTRUCTURE ONLY: built into libfadeemu.so and loaded only by tests/ (CPU, no GPU) to
// check the shared arithmetic core (sw_core.cuh: packed DP step, checkpoint/replay geometry,
// end-cell rule, traceback state machine, accept predicates) against the oracle.  The product
// library (libfadegpu.so) never links or loads this file; its orchestration of the same core is
// the CUDA code in kernels.cu.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../sw_core.cuh"

using namespace fade;

namespace {

template <int R>
struct ThreadState {
    uint32_t H[R], E[R], qs[R];
    uint32_t hu_prev, fout, M;
};

template <int R>
int align_pair(const uint8_t *qa, int qlen_a, const uint8_t *ta, int tlen_a,
               const uint8_t *qb, int qlen_b, const uint8_t *tb, int tlen_b,
               const SwConsts &k, int extra_blocks, int tagged, int32_t min_length,
               uint32_t clipl_a, uint32_t clipr_a, uint32_t clipl_b, uint32_t clipr_b,
               AlnOut *out_a, AlnOut *out_b)
{
    const int rows = FG * R;
    if (qlen_a > rows || qlen_b > rows) return -1;
    const int tmax = tlen_a > tlen_b ? tlen_a : tlen_b;
    const int nblk = num_blocks(tmax) + extra_blocks;
    const int TW = FBLK * nblk + FG + 1;
    std::vector<uint16_t> tw(TW);
    for (int idx = 0; idx < TW; ++idx) {
        const int j = idx - FG;
        const int ca = (j >= 0 && j < tlen_a) ? ta[j] : C_TPAD;
        const int cb = (j >= 0 && j < tlen_b) ? tb[j] : C_TPAD;
        tw[idx] = (uint16_t)t_sel(ca, cb);
    }
    std::vector<uint8_t> qc(rows);
    std::vector<ThreadState<R>> st(FG);
    for (int g = 0; g < FG; ++g)
        for (int r = 0; r < R; ++r) {
            const int i = g * R + r;
            const int ca = i < qlen_a ? qa[i] : C_QPAD;
            const int cb = i < qlen_b ? qb[i] : C_QPAD;
            st[g].qs[r] = q_sel(ca, cb);
            qc[i] = (uint8_t)(ca | (cb << 4));
        }
    auto init_state = [&](ThreadState<R> &s) {
        for (int r = 0; r < R; ++r) { s.H[r] = 0; s.E[r] = k.neg_o; }
        s.hu_prev = 0; s.fout = k.neg_o; s.M = 0;
    };
    for (int g = 0; g < FG; ++g) init_state(st[g]);

    // ---- score-only pass with checkpoints ----
    const int CW = ck_words<R>();
    std::vector<uint32_t> ck((size_t)nblk * FG * CW);
    LaneCtl ctl[2];
    memset(ctl, 0, sizeof(ctl));
    uint32_t Mprev[FG];
    for (int g = 0; g < FG; ++g) Mprev[g] = 0;
    for (int c = 0; c < nblk; ++c) {
        for (int u = 0; u < FBLK; ++u) {
            const int t = c * FBLK + u;
            uint32_t hu[FG], fin[FG];
            for (int g = 0; g < FG; ++g) {
                hu[g] = g == 0 ? 0u : st[g - 1].H[R - 1];   // __shfl_up of the bottom-row H
                fin[g] = g == 0 ? k.neg_o : st[g - 1].fout; // __shfl_up of the running F
            }
            for (int g = 0; g < FG; ++g) {
                const uint32_t ts = tw[t - g + FG];
                const uint32_t hd = st[g].hu_prev;
                fill_step<R>(st[g].H, st[g].E, st[g].qs, st[g].M, ts, hd, fin[g], st[g].fout, k);
                st[g].hu_prev = hu[g];
            }
        }
        for (int g = 0; g < FG; ++g) {
            uint32_t *p = &ck[((size_t)c * FG + g) * CW];
            for (int r = 0; r < R; ++r) { p[r] = st[g].H[r]; p[R + r] = st[g].E[r]; }
            p[2 * R] = st[g].hu_prev;
            p[2 * R + 1] = st[g].fout;
            if (st[g].M != Mprev[g]) {
                if (lane_lo(st[g].M) != lane_lo(Mprev[g])) ctl[0].blk[g] = c;
                if (lane_hi(st[g].M) != lane_hi(Mprev[g])) ctl[1].blk[g] = c;
                Mprev[g] = st[g].M;
            }
        }
    }
    for (int g = 0; g < FG; ++g) { ctl[0].best[g] = lane_lo(st[g].M); ctl[1].best[g] = lane_hi(st[g].M); }

    // ---- traceback: replay blocks with trace recording ----
    ctl_init(ctl[0], qlen_a, tlen_a);
    ctl_init(ctl[1], qlen_b, tlen_b);
    constexpr int RW = trace_words<R>();
    std::vector<uint32_t> tr((size_t)tile_words<R>());
    const bool tg = (tagged & 1) && k.tagged_ok;
    int guard = 0;
    bool first_iter = true;
    while (ctl[0].phase != 2 || ctl[1].phase != 2) {
        if (++guard > 4 * nblk + 16) return -2;
        int blk[2];
        for (int L = 0; L < 2; ++L) blk[L] = (ctl[L].phase != 2 && ctl[L].next_blk >= 0) ? ctl[L].next_blk : 0;
        bool scan[2][FG];
        for (int L = 0; L < 2; ++L)
            for (int g = 0; g < FG; ++g) scan[L][g] = (ctl_scanmask(ctl[L]) >> g) & 1u;
        // load per-lane state
        for (int g = 0; g < FG; ++g) {
            ThreadState<R> s0, s1;
            init_state(s0); init_state(s1);
            auto load = [&](ThreadState<R> &s, int b) {
                if (b == 0) return;
                const uint32_t *p = &ck[((size_t)(b - 1) * FG + g) * CW];
                for (int r = 0; r < R; ++r) { s.H[r] = p[r]; s.E[r] = p[R + r]; }
                s.hu_prev = p[2 * R]; s.fout = p[2 * R + 1];
            };
            load(s0, blk[0]); load(s1, blk[1]);
            auto mix = [&](uint32_t a, uint32_t b) {
                const uint32_t w = (a & 0xffffu) | (b & 0xffff0000u);
                return tg ? to_tagged(w) : w;
            };
            for (int r = 0; r < R; ++r) { st[g].H[r] = mix(s0.H[r], s1.H[r]); st[g].E[r] = mix(s0.E[r], s1.E[r]); }
            st[g].hu_prev = mix(s0.hu_prev, s1.hu_prev);
            st[g].fout = mix(s0.fout, s1.fout);
        }
        const uint32_t f_top = tg ? k.neg_o16 : k.neg_o;
        const int mul = tg ? 16 : 1;
        bool found[2][FG];
        for (int L = 0; L < 2; ++L) for (int g = 0; g < FG; ++g) found[L][g] = false;
        for (int u = 0; u < FBLK; ++u) {
            const int t0 = blk[0] * FBLK + u, t1 = blk[1] * FBLK + u;
            uint32_t hu[FG], fin[FG];
            for (int g = 0; g < FG; ++g) {
                hu[g] = g == 0 ? 0u : st[g - 1].H[R - 1];
                fin[g] = g == 0 ? f_top : st[g - 1].fout;
            }
            for (int g = 0; g < FG; ++g) {
                const uint32_t ts = (tw[t0 - g + FG] & 0x00ffu) | (tw[t1 - g + FG] & 0xff00u);
                uint32_t cmax = 0;
                uint32_t trw[RW];
                if (tg) trace_step_tagged<R>(st[g].H, st[g].E, st[g].qs, ts, st[g].hu_prev, fin[g], st[g].fout, k, trw, cmax);
                else trace_step_plain<R>(st[g].H, st[g].E, st[g].qs, ts, st[g].hu_prev, fin[g], st[g].fout, k, trw, cmax);
                st[g].hu_prev = hu[g];
                for (int w = 0; w < RW; ++w) tr[(size_t)tile_index<R>(u, g, w)] = trw[w];
                for (int L = 0; L < 2; ++L) {
                    if (!scan[L][g] || found[L][g]) continue;
                    const int cm = L == 0 ? lane_lo(cmax) : lane_hi(cmax);
                    if (cm != ctl[L].S * mul) continue;
                    for (int r = 0; r < R; ++r) {
                        const int hv = L == 0 ? lane_lo(st[g].H[r]) : lane_hi(st[g].H[r]);
                        if (hv == ctl[L].S * mul) {
                            found[L][g] = true;
                            ctl[L].fj[g] = (L == 0 ? t0 : t1) - g;
                            ctl[L].fr[g] = r;
                            break;
                        }
                    }
                }
            }
        }
        for (int L = 0; L < 2; ++L)
            for (int g = 0; g < FG; ++g)
                if (scan[L][g] && !found[L][g]) return -3;  // a candidate must find its cell
        // bit 2 of `tagged`: the first (scan) replay of a lane records no trace, like round 0 on the device
        const bool so0 = (tagged & 4) && ctl[0].phase == 0 && first_iter, so1 = (tagged & 4) && ctl[1].phase == 0 && first_iter;
        ctl_advance<R>(ctl[0], so0 ? nullptr : tr.data(), 0, StagedAcc{ tw.data(), qc.data(), 0 }, k);
        ctl_advance<R>(ctl[1], so1 ? nullptr : tr.data(), 1, StagedAcc{ tw.data(), qc.data(), 1 }, k);
        first_iter = false;
    }
    finalize_result(ctl[0], *out_a, 0, clipl_a, clipr_a, min_length);
    finalize_result(ctl[1], *out_b, 1, clipl_b, clipr_b, min_length);
    return 0;
}

}  // namespace

extern "C" int fadeemu_align_pair(int R, const uint8_t *qa, int qlen_a, const uint8_t *ta, int tlen_a,
                                  const uint8_t *qb, int qlen_b, const uint8_t *tb, int tlen_b,
                                  int open, int extend, int match, int mismatch, int extra_blocks,
                                  int tagged, int min_length, const uint32_t *clips /*[4]: la ra lb rb*/,
                                  AlnOut *out_a, AlnOut *out_b)
{
    SwConsts k = make_consts(open, extend, match, mismatch);
    k.shortcut = (tagged & 2) ? 0 : 1;   // bit 1 of `tagged`: disable the ungapped-diagonal shortcut
#define CASE(RR)                                                                                   \
    case RR:                                                                                       \
        return align_pair<RR>(qa, qlen_a, ta, tlen_a, qb, qlen_b, tb, tlen_b, k, extra_blocks,    \
                              tagged, min_length, clips[0], clips[1], clips[2], clips[3], out_a, out_b);
    switch (R) {
        CASE(1) CASE(2) CASE(3) CASE(5) CASE(13) CASE(19) CASE(25) CASE(32) CASE(38)
    default: return -10;
    }
#undef CASE
}

// helpers exposed for unit tests of the decoding primitives
extern "C" int fadeemu_comp_code_of_nt16(int nib) { return comp_code_of_nt16(nib); }
extern "C" uint32_t fadeemu_prmt(uint32_t a, uint32_t b, uint32_t s) { return prmt_sx(a, b, s); }
extern "C" int fadeemu_sizeof_alnout(void) { return (int)sizeof(AlnOut); }
