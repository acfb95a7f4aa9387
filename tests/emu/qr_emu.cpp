// CPU EMULATION HARNESS -- test infrastructure only, never loaded by the product package.
//
// Compiles the product's __host__ __device__ work-item functions (csrc/qr_decode_core.cuh,
// csrc/qr_mapper_core.cuh) with g++ and drives them with plain loops in place of the CUDA
// grid, so the lane state machine, the index arithmetic and the host-side math can be checked
// against the oracle in the container that has no GPU.  It says nothing about launch
// configuration, barriers or memory ordering: those are covered by the -m gpu tests.
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../qam-reconciliation_b200/csrc/qr_common.h"
#include "../../qam-reconciliation_b200/csrc/qr_graph_build.h"
#include "../../qam-reconciliation_b200/csrc/qr_decode_core.cuh"
#include "../../qam-reconciliation_b200/csrc/qr_decode_fused.cuh"
#include "../../qam-reconciliation_b200/csrc/qr_mapper_core.cuh"

namespace qr {
static std::string g_err;
void set_error(const std::string &m) { g_err = m; }
int fail(int code, const std::string &m) { g_err = m; return code; }
}  // namespace qr

using namespace qr;

template <typename T, int VEC, int D>
static uint32_t emu_bin(const DecodeParams<T> &P, const LaneInfo<VEC> &L, const CheckBin &bin)
{
    // split the bin over a few emulated "threads" so the prefetch pipeline sees first/stride > trivial
    uint32_t bad = 0;
    for (int32_t t = 0; t < 3; ++t) bad |= run_check_bin<T, VEC, D>(P, L, bin, t, 3);
    return bad;
}

template <typename T, int VEC>
static int emu_decode_t(const qr_graph &g, int lanes, bool generic, const void *llr, int llr_dtype,
                        const uint8_t *synd, int64_t frames, int maxiter, uint8_t *success, int32_t *iters,
                        void *post, int post_dtype, int64_t *steps_out)
{
    std::vector<T> c2v((size_t)g.E * lanes), postw((size_t)g.N * lanes), llrw((size_t)g.N * lanes);
    std::vector<uint8_t> syndw((size_t)g.C * lanes);
    std::vector<LaneState> st(2 * lanes);
    std::vector<int32_t> unsat(2 * lanes, 0), ctrl(CTRL_WORDS, 0);
    unsigned long long stats[2] = {0, 0};
    // poison the workspace: nothing may depend on its initial contents
    for (auto &v : c2v) v = (T)1e30;
    for (auto &v : postw) v = (T)-7e29;
    for (auto &v : llrw) v = (T)3e29;
    DecodeParams<T> P;
    P.bins = g.bins.data(); P.n_bins = (int32_t)g.bins.size();
    P.chk_order = g.chk_order.data(); P.slot_var = g.slot_var.data();
    P.var_ptr = g.var_ptr.data(); P.var_slot = g.var_slot.data();
    P.var_work = g.var_work.empty() ? nullptr : g.var_work.data();
    P.var_bins = g.var_bins.empty() ? nullptr : g.var_bins.data(); P.n_var_bins = (int32_t)g.var_bins.size();
    P.vslot_sorted = g.vslot_sorted.empty() ? nullptr : g.vslot_sorted.data();
    P.N = g.N; P.C = g.C; P.E = g.E; P.lanes = lanes; P.var_deg = g.var_deg;
    P.c2v = c2v.data(); P.post = postw.data(); P.llr = llrw.data(); P.synd = syndw.data();
    P.st[0] = st.data(); P.st[1] = st.data() + lanes;
    P.unsat[0] = unsat.data(); P.unsat[1] = unsat.data() + lanes;
    P.llr_in = llr; P.llr_in_f64 = llr_dtype == QR_F64; P.synd_in = synd;
    P.frames = frames; P.maxiter = maxiter; P.success = success; P.iters = iters;
    P.post_out = post; P.post_out_f64 = post_dtype == QR_F64;
    P.ctrl = ctrl.data(); P.stats = stats; P.work = nullptr; P.refill_list = nullptr;
    for (int l = 0; l < lanes; ++l) {
        LaneState s;
        s.frame = l < frames ? l : -1; s.iter = 0; s.fresh = s.frame >= 0; s.retire = -1;
        st[l] = s; st[lanes + l] = s;
    }
    ctrl[CTRL_NEXT_FRAME] = (int32_t)std::min<int64_t>(lanes, frames);
    ctrl[CTRL_REMAINING] = (int32_t)frames;
    ctrl[CTRL_FIN_STEP] = -1;
    auto refill = [&](int buf) {
        for (int lane = 0; lane < lanes; ++lane) {
            const LaneState s = P.st[buf][lane];
            if (!lane_needs_refill(s)) continue;
            for (int32_t n = 0; n < g.N; ++n) refill_var_elem<T>(P, s, lane, n);
            for (int32_t ci = 0; ci < g.C; ++ci) refill_chk_elem<T>(P, s, lane, ci);
        }
    };
    refill(0);
    const int LV = lanes / VEC;
    int64_t step = 0;
    for (; ctrl[CTRL_REMAINING] > 0; ++step) {
        if (step > (frames + lanes) * (int64_t)(maxiter + 3)) return -1;  // runaway guard
        const int cur = step & 1;
        for (int jv = 0; jv < LV; ++jv) {
            LaneInfo<VEC> L = load_lane_info<T, VEC>(P, cur, jv);
            if (!L.active) continue;
            uint32_t bad = 0;
            for (const CheckBin &bin : g.bins) {
                if (generic) { bad |= emu_bin<T, VEC, 0>(P, L, bin); continue; }
                switch (bin.degree) {
                case 2: bad |= emu_bin<T, VEC, 2>(P, L, bin); break;
                case 3: bad |= emu_bin<T, VEC, 3>(P, L, bin); break;
                case 4: bad |= emu_bin<T, VEC, 4>(P, L, bin); break;
                case 5: bad |= emu_bin<T, VEC, 5>(P, L, bin); break;
                case 6: bad |= emu_bin<T, VEC, 6>(P, L, bin); break;
                case 7: bad |= emu_bin<T, VEC, 7>(P, L, bin); break;
                case 8: bad |= emu_bin<T, VEC, 8>(P, L, bin); break;
                default: bad |= emu_bin<T, VEC, 0>(P, L, bin); break;
                }
            }
            for (int k = 0; k < VEC; ++k)
                if (bad >> k & 1) P.unsat[cur][L.l0 + k] = 1;
        }
        for (int jv = 0; jv < LV; ++jv) {
            LaneInfo<VEC> L = load_lane_info<T, VEC>(P, cur, jv);
            decide_lanes<T, VEC>(P, cur, L);
            for (int32_t t = 0; t < 5; ++t) run_var_range<T, VEC>(P, L, t, 5, (int32_t)g.N);
            bookkeep_lanes<T, VEC>(P, cur, (int32_t)step, L);
        }
        if (ctrl[CTRL_FIN_STEP] == step) refill(cur ^ 1);
    }
    if (steps_out) *steps_out = step;
    return 0;
}

// QR_SCHED_FUSED as a tile pipeline (qr_decode_fused.cuh): the stream of work items executed one after the other in
// stream order -- sweep F(t), bookkeeping BK(t) (what the last finisher of the sweep does on the device), and the
// post-processing PP one tile behind.  Run sequentially, every wait of the device code is trivially satisfied; what
// is checked here is the lane state machine, the refill lists, the record decoding and the arithmetic.
template <typename T, int VEC>
static int emu_decode_fused_t(const qr_graph &g, int lanes, int tl, const void *llr, int llr_dtype,
                              const uint8_t *synd, int64_t frames, int maxiter, uint8_t *success, int32_t *iters,
                              void *post, int post_dtype, int64_t *steps_out, int store_post, int64_t *shipped_out,
                              int park_rounds = 2)
{
    if (g.slot_nbr.empty() || g.max_cdeg > kFusedMaxCheckDegree || lanes % tl) return -2;
    std::vector<T> c2v0((size_t)g.E * lanes, (T)1e30), c2v1((size_t)g.E * lanes, (T)-3e30), llrw((size_t)g.N * lanes, (T)3e29);
    std::vector<uint8_t> syndw((size_t)g.C * lanes, 0xff);
    std::vector<LaneState> st(lanes);
    std::vector<int32_t> unsat(lanes, 0), ctrl(CTRL_WORDS, 0);
    unsigned long long stats[2] = {0, 0};
    FusedParams<T> F;
    DecodeParams<T> &P = F.P;
    P.bins = g.bins.data(); P.n_bins = (int32_t)g.bins.size();
    P.chk_order = g.chk_order.data(); P.slot_var = g.slot_var.data();
    P.var_ptr = g.var_ptr.data(); P.var_slot = g.var_slot.data();
    P.var_work = g.var_work.empty() ? nullptr : g.var_work.data();
    P.var_bins = nullptr; P.n_var_bins = 0; P.vslot_sorted = nullptr;
    P.N = g.N; P.C = g.C; P.E = g.E; P.lanes = lanes; P.var_deg = g.var_deg;
    P.c2v = nullptr; P.post = nullptr; P.llr = llrw.data(); P.synd = syndw.data();
    P.st[0] = st.data(); P.st[1] = nullptr;
    P.unsat[0] = unsat.data(); P.unsat[1] = nullptr;
    P.llr_in = llr; P.llr_in_f64 = llr_dtype == QR_F64; P.synd_in = synd;
    P.frames = frames; P.maxiter = maxiter; P.success = success; P.iters = iters;
    P.post_out = post; P.post_out_f64 = post_dtype == QR_F64;
    P.ctrl = ctrl.data(); P.stats = stats; P.work = nullptr; P.refill_list = nullptr;
    F.nbr = reinterpret_cast<const Nbr4 *>(g.slot_nbr.data());
    F.nbr_lean = nullptr;
    F.c2v[0] = c2v0.data(); F.c2v[1] = c2v1.data();
    const int tiles = lanes / tl;
    F.tl = tl; F.tiles = tiles; F.hints = 0; F.rows_per_claim = 2; F.pp_items = 3; F.park_rounds = park_rounds;
    std::vector<T> postw((size_t)g.N * lanes, (T)-7e29);
    F.post = store_post ? postw.data() : nullptr;
    std::vector<int32_t> tile_minfin(tiles, 0x7fffffff), rcount(tiles, 0);
    std::vector<RefillEntry> rlist(lanes);
    F.tile_minfin = tile_minfin.data(); F.rcount = rcount.data(); F.rlist = rlist.data();
    F.f_done = F.pp_done = F.pp_expect = nullptr; F.bk_word = nullptr; F.lane_flags = nullptr; F.ppq = nullptr; F.ppq_size = 0;
    // initial fill: lane l takes frame l
    for (int l = 0; l < lanes; ++l) {
        LaneState s;
        s.frame = l < frames ? l : -1; s.iter = 0; s.fresh = s.frame >= 0; s.retire = -1;
        st[l] = s;
        if (s.frame < 0) continue;
        const RefillEntry e{l % tl, -1, s.frame, 0};
        for (int32_t n = 0; n < g.N; ++n) fused_pp_var_elem<T>(F, 0, e, l / tl, n);
        for (int32_t ci = 0; ci < g.C; ++ci) fused_pp_chk_elem<T>(F, e, l / tl, ci);
    }
    ctrl[CTRL_NEXT_FRAME] = (int32_t)std::min<int64_t>(lanes, frames);
    ctrl[CTRL_REMAINING] = (int32_t)frames;
    ctrl[CTRL_MINFIN] = 0x7fffffff;
    int64_t shipped_from_post = 0, completed = 0;
    // float mode on a (3,6)-regular graph: the lean item with premultiplied-offset records, as on the device
    std::vector<NbrL> lean;
    if constexpr (std::is_same<T, float>::value) {
        if (g.var_deg == 3 && g.bins.size() == 1 && g.bins[0].degree == 6) {
            lean.resize(g.E);
            for (int64_t e = 0; e < g.E; ++e) lean[e] = make_lean_record(F.nbr[e], tl);
            F.nbr_lean = lean.data();
        }
    }
    auto sweep = [&](int tile, int cur) {
        const TileView<T> V = tile_view(F, cur, tile);
        for (int tx = 0; tx < tl / VEC; ++tx) {
            const LaneInfo<VEC> L = load_tile_lanes<T, VEC>(F, (tile * tl) / VEC + tx, tile_minfin[tile]);
            if (!L.active) continue;
            uint32_t bad = 0;
            if constexpr (std::is_same<T, float>::value) {
                if (!lean.empty()) {
                    const int32_t lt4 = tx * VEC * 4;
                    for (int32_t ci = 0; ci < g.C; ++ci) {
                        const char *llr_t = reinterpret_cast<const char *>(V.llr) + lt4;
                        const char *cold_t = reinterpret_cast<const char *>(V.c_old) + lt4;
                        char *cnew_t = reinterpret_cast<char *>(V.c_new) + lt4;
                        char *post_t = reinterpret_cast<char *>(V.post) + lt4;
                        if (L.fresh) bad |= fused_item_lean<6, true>(llr_t, cold_t, cnew_t, post_t, V.synd + tx * VEC, lean.data(), ci, 6 * ci, tl, L.fresh, L.wpost, L.active, false);
                        else bad |= fused_item_lean<6, false>(llr_t, cold_t, cnew_t, post_t, V.synd + tx * VEC, lean.data(), ci, 6 * ci, tl, 0u, L.wpost, L.active, false);
                    }
                    for (int k = 0; k < VEC; ++k)
                        if (bad >> k & 1) P.unsat[0][L.l0 + k] = 1;
                    continue;
                }
            }
            for (const CheckBin &bin : g.bins)
                for (int32_t t = 0; t < 3; ++t)
                    bad |= run_fused_bin_any<T, VEC>(V, F.nbr, L, tx * VEC, bin, t, 3, 0, 0);
            for (int k = 0; k < VEC; ++k)
                if (bad >> k & 1) P.unsat[0][L.l0 + k] = 1;
        }
    };
    auto bookkeep = [&](int tile) {          // tile_bookkeep of qr_decode_fused.cu, one lane after the other
        const int32_t minfin_used = tile_minfin[tile];
        int32_t listed = 0;
        for (int base = 0; base < tl; base += 32) {
            uint32_t m_run = 0, m_park = 0, m_aged = 0;
            for (int l = 0; l < 32; ++l) {
                const int lane = tile * tl + base + l;
                LaneState s = st[lane];
                const BkDecision d = bk_decide(s, unsat[lane], maxiter);
                if (d.fin_ok || d.fin_fail) {
                    success[s.frame] = d.fin_ok ? 1 : 0;
                    iters[s.frame] = d.fin_ok ? s.iter : maxiter;
                    if (d.fin_ok && s.iter > 0 && s.iter < ctrl[CTRL_MINFIN]) ctrl[CTRL_MINFIN] = s.iter;
                    stats[0] += (unsigned long long)s.iter;
                    ctrl[CTRL_REMAINING] -= 1;
                    const int32_t pv = (F.post && stores_post(s.iter, maxiter, minfin_used)) ? 1 : 0;
                    s.retire = s.frame; s.frame = -1; s.iter = pv; s.fresh = pv ? 0 : F.park_rounds;
                } else if (s.frame >= 0) {
                    s.iter += 1; s.fresh = 0; s.retire = -1;
                } else if (s.retire >= 0) {
                    s.fresh += 1;
                }
                const bool parked = s.frame < 0 && s.retire >= 0;
                if (s.frame >= 0) m_run |= 1u << l;
                if (parked) m_park |= 1u << l;
                if (parked && s.fresh >= F.park_rounds) m_aged |= 1u << l;
                st[lane] = s;
                unsat[lane] = 0;
            }
            const bool no_more = ctrl[CTRL_NEXT_FRAME] >= frames;
            const uint32_t m_rel = octets_to_release(m_run, m_park, m_aged, no_more);
            for (int l = 0; l < 32; ++l) {
                if (!(m_rel >> l & 1)) continue;
                const int lane = tile * tl + base + l;
                LaneState s = st[lane];
                RefillEntry e;
                e.lane = base + l; e.retire = s.retire;
                const int32_t nf = ctrl[CTRL_NEXT_FRAME]++;
                e.frame = (int64_t)nf < frames ? nf : -1;
                e.post_valid = s.iter;
                rlist[(size_t)tile * tl + listed++] = e;
                s.frame = e.frame; s.iter = 0; s.fresh = e.frame >= 0 ? 1 : 0; s.retire = -1;
                st[lane] = s;
            }
        }
        rcount[tile] = listed;
        tile_minfin[tile] = ctrl[CTRL_MINFIN];
    };
    auto post_process = [&](int tile, int cur) {
        for (int32_t ei = 0; ei < rcount[tile]; ++ei) {
            const RefillEntry e = rlist[(size_t)tile * tl + ei];
            if (e.post_valid && iters[e.retire] != 0) ++shipped_from_post;
            for (int32_t n = 0; n < g.N; ++n) fused_pp_var_elem<T>(F, cur, e, tile, n);
            for (int32_t ci = 0; ci < g.C; ++ci) fused_pp_chk_elem<T>(F, e, tile, ci);
            ++completed;
        }
        rcount[tile] = 0;
    };
    const int lag = std::min(1, tiles - 1);    // (on the device PP follows BK through a ready queue: any order after BK is legal)
    int64_t round = 0;
    for (; completed < frames; ++round) {
        if (round > (frames + lanes) * (int64_t)(maxiter + 3)) return -1;
        const int cur = (int)(round & 1);
        for (int t = 0; t < tiles; ++t) {
            sweep(t, cur);
            bookkeep(t);
            const int pt = t >= lag ? t - lag : t - lag + tiles;
            const int64_t pr = t >= lag ? round : round - 1;
            if (pr >= 0) post_process(pt, (int)(pr & 1));
        }
    }
    if (shipped_out) *shipped_out = shipped_from_post;
    if (steps_out) *steps_out = round;
    return 0;
}

extern "C" {

const char *emu_last_error() { return qr::g_err.c_str(); }

int emu_graph_tables(const int64_t *vid, const int64_t *cid, int64_t E, int64_t *dims, int32_t *chk_order,
                     int32_t *slot_edge, int32_t *slot_var, int32_t *var_ptr, int32_t *var_slot)
{
    qr_graph g;
    int rc = build_host_tables(g, vid, cid, E);
    if (rc) return rc;
    dims[0] = g.N; dims[1] = g.C; dims[2] = g.E; dims[3] = g.max_cdeg; dims[4] = g.max_vdeg;
    dims[5] = (int64_t)g.bins.size();
    if (chk_order) memcpy(chk_order, g.chk_order.data(), g.C * 4);
    if (slot_edge) memcpy(slot_edge, g.slot_edge.data(), g.E * 4);
    if (slot_var) memcpy(slot_var, g.slot_var.data(), g.E * 4);
    if (var_ptr) memcpy(var_ptr, g.var_ptr.data(), (g.N + 1) * 4);
    if (var_slot) memcpy(var_slot, g.var_slot.data(), g.E * 4);
    return 0;
}

int emu_decode(const int64_t *vid, const int64_t *cid, int64_t E, int precision, int lanes, int generic,
               const void *llr, int llr_dtype, const uint8_t *synd, int64_t frames, int maxiter,
               uint8_t *success, int32_t *iters, void *post, int post_dtype, int64_t *steps)
{
    qr_graph g;
    int rc = build_host_tables(g, vid, cid, E);
    if (rc) return rc;
    if (precision == QR_F64)
        return emu_decode_t<double, 2>(g, lanes, generic != 0, llr, llr_dtype, synd, frames, maxiter, success,
                                       iters, post, post_dtype, steps);
    return emu_decode_t<float, 4>(g, lanes, generic != 0, llr, llr_dtype, synd, frames, maxiter, success, iters,
                                  post, post_dtype, steps);
}

int emu_decode_fused(const int64_t *vid, const int64_t *cid, int64_t E, int precision, int lanes, int tile_lanes,
                     const void *llr, int llr_dtype, const uint8_t *synd, int64_t frames, int maxiter,
                     uint8_t *success, int32_t *iters, void *post, int post_dtype, int64_t *steps, int store_post,
                     int64_t *shipped_from_post)
{
    qr_graph g;
    int rc = build_host_tables(g, vid, cid, E);
    if (rc) return rc;
    if (precision == QR_F64)
        return emu_decode_fused_t<double, 2>(g, lanes, tile_lanes, llr, llr_dtype, synd, frames, maxiter, success,
                                             iters, post, post_dtype, steps, store_post, shipped_from_post);
    return emu_decode_fused_t<float, 4>(g, lanes, tile_lanes, llr, llr_dtype, synd, frames, maxiter, success, iters,
                                        post, post_dtype, steps, store_post, shipped_from_post);
}

// mapper arithmetic: mode bit 0 = fast inverse, bit 1 = corrected exponent
void emu_demap(int bps, const double *a, const double *thr, const double *p, double noise_var,
               const uint8_t *sign, const double *n_hat, const int64_t *tx, int64_t n, int mode, double *llr,
               double *yhat_out)
{
    const int M = 1 << bps;
    const double sigma = sqrt(noise_var), s2 = sqrt(2.0) * sigma;
    std::vector<double> FYt(M + 1), delta(M);
    FYt[0] = 0; FYt[M] = 1;
    for (int i = 1; i < M; ++i) FYt[i] = mixture_cdf(a, p, M, s2, thr[i]);
    for (int i = 0; i < M; ++i) delta[i] = FYt[i + 1] - FYt[i];
    std::vector<double> yh(M);
    // the same starting table libqamrecon builds on the device (qr_mapper_create)
    const int32_t tn = 16385;
    const double ty0 = a[0] - 9.0 * sigma, th = (a[M - 1] + 9.0 * sigma - ty0) / (tn - 1);
    std::vector<double> tabF(tn), tabf(tn);
    for (int32_t j = 0; j < tn; ++j) {
        tabF[j] = mixture_cdf(a, p, M, s2, ty0 + j * th);
        tabf[j] = mixture_pdf(a, p, M, sigma, ty0 + j * th);
    }
    // mode bit 2: no table at all; bit 3: table of F only (no Hermite solve)
    const int32_t jn = 8192;
    std::vector<int32_t> jump(jn + 2, 0);
    for (int32_t t = 0; t <= jn; ++t) {
        const double v = (double)t / (double)jn;
        int32_t lo = 0, hi = tn;
        while (hi - lo > 1) {
            const int32_t mid = (lo + hi) >> 1;
            if (tabF[mid] <= v) lo = mid; else hi = mid;
        }
        jump[t] = lo;
    }
    // mode bit 4: without the jump table (plain binary search over the region)
    const InvTable tab{(mode & 4) ? nullptr : tabF.data(), (mode & 12) ? nullptr : tabf.data(), (mode & 4) ? 0 : tn, ty0, th,
                       (mode & 16) ? nullptr : jump.data(), (mode & 16) ? 0 : jn};
    for (int64_t s = 0; s < n; ++s) {
        for (int i = 0; i < M; ++i) {
            const double target = inv_target(sign, FYt.data(), delta.data(), n_hat[s], i);
            yh[i] = (mode & 1) ? g_inv_fast(a, p, thr, FYt.data(), M, sigma, s2, target, 1e-9, i, tab)
                               : g_inv_exact(a, p, M, s2, target, 1e-9);
            if (yhat_out) yhat_out[s * M + i] = yh[i];
        }
        demap_from_yhat(a, p, delta.data(), M, bps, 2 * noise_var, yh.data(), (int32_t)tx[s], (mode & 2) != 0,
                        llr + s * bps);
    }
}

// demap_symbol (the function k_demap runs per symbol) with host-built tables: every QR_DEMAP_* mode, including the
// fp32-grade path (host stand-ins for the MUFU instructions: exp2f / log2f / float division)
void emu_demap_symbol(int bps, const double *a, const double *thr, const double *p, double noise_var,
                      const uint8_t *sign, const double *n_hat, const int64_t *tx, int64_t n, int mode, double alpha,
                      double *llr)
{
    const int M = 1 << bps;
    const double sigma = sqrt(noise_var), s2 = sqrt(2.0) * sigma;
    std::vector<double> FYt(M + 1), delta(M), ghi(M), glo(M), pz(3 * M);
    std::vector<float> g2hi(M), g2lo(M), pzf(3 * M);
    FYt[0] = 0; FYt[M] = 1;
    for (int i = 1; i < M; ++i) FYt[i] = mixture_cdf(a, p, M, s2, thr[i]);
    for (int i = 0; i < M; ++i) delta[i] = FYt[i + 1] - FYt[i];
    const double step = M > 1 ? a[1] - a[0] : 0.0;
    for (int i = 0; i < M; ++i) {
        const double t2 = (i * step) * (i * step);
        ghi[i] = exp(-t2 / (2 * noise_var));
        glo[i] = exp(-t2);
        g2hi[i] = (float)(t2 / (2 * noise_var) * 1.4426950408889634);
        g2lo[i] = (float)(t2 * 1.4426950408889634);
    }
    for (int i = 0; i < 3 * M; ++i) {
        const int k = i - (M - 1);
        pz[i] = (k >= 0 && k < M) ? p[k] : 0.0;
        pzf[i] = (float)pz[i];
    }
    const int32_t tn = 16385, jn = 8192;
    const double ty0 = a[0] - 9.0 * sigma, th = (a[M - 1] + 9.0 * sigma - ty0) / (tn - 1);
    std::vector<double> tabF(tn), tabf(tn);
    for (int32_t j = 0; j < tn; ++j) {
        tabF[j] = mixture_cdf(a, p, M, s2, ty0 + j * th);
        tabf[j] = mixture_pdf(a, p, M, sigma, ty0 + j * th);
    }
    std::vector<int32_t> jump(jn + 2, 0);
    for (int32_t t = 0; t <= jn; ++t) {
        const double v = (double)t / (double)jn;
        int32_t lo = 0, hi = tn;
        while (hi - lo > 1) {
            const int32_t mid = (lo + hi) >> 1;
            if (tabF[mid] <= v) lo = mid; else hi = mid;
        }
        jump[t] = lo;
    }
    MapperView m{};
    m.order = M; m.bps = bps; m.noise_var = noise_var; m.sigma = sigma; m.s2 = s2;
    m.constellation = a; m.thresholds = thr; m.probabilities = p; m.sign_config = sign; m.sign_g = sign;
    m.FY_thr = FYt.data(); m.delta = delta.data(); m.bare = nullptr;
    m.inv_tab = tabF.data(); m.inv_pdf = tabf.data(); m.inv_n = tn; m.inv_y0 = ty0; m.inv_h = th;
    m.inv_jump = jump.data(); m.inv_jn = jn; m.uniform = 1; m.index_errors = nullptr;
    TablesRef t{a, p, thr, FYt.data(), delta.data(), sign, ghi.data(), glo.data(), pz.data(), g2hi.data(), g2lo.data(), pzf.data()};
    // the coarse copy k_demap32 stages in shared memory
    const double h32 = (a[M - 1] + 9.0 * sigma - ty0) / (kInv32N - 1);
    std::vector<double> F32(kInv32N);
    std::vector<float> f32(kInv32N);
    std::vector<uint16_t> j32(kInv32J + 2);
    for (int32_t j = 0; j < kInv32N; ++j) {
        F32[j] = mixture_cdf(a, p, M, s2, ty0 + j * h32);
        f32[j] = (float)mixture_pdf(a, p, M, sigma, ty0 + j * h32);
    }
    for (int32_t tt = 0; tt <= kInv32J + 1; ++tt) {
        const double v = (double)tt / (double)kInv32J;
        int32_t lo = 0, hi = kInv32N;
        while (hi - lo > 1) {
            const int32_t mid = (lo + hi) >> 1;
            if (F32[mid] <= v) lo = mid; else hi = mid;
        }
        j32[tt] = (uint16_t)lo;
    }
    t.t32 = InvTable32{F32.data(), f32.data(), j32.data(), ty0, h32};
    for (int64_t s = 0; s < n; ++s) demap_symbol_any(m, t, n_hat[s], (int32_t)tx[s], mode, alpha, llr + s * bps);
}

void emu_front(int bps, const double *a, const double *thr, const double *p, double noise_var,
               const uint8_t *sign, const double *y, int64_t n, int64_t *idx, double *n_hat, uint8_t *bits)
{
    const int M = 1 << bps;
    const double sigma = sqrt(noise_var), s2 = sqrt(2.0) * sigma;
    std::vector<double> FYt(M + 1), delta(M);
    FYt[0] = 0; FYt[M] = 1;
    for (int i = 1; i < M; ++i) FYt[i] = mixture_cdf(a, p, M, s2, thr[i]);
    for (int i = 0; i < M; ++i) delta[i] = FYt[i + 1] - FYt[i];
    for (int64_t j = 0; j < n; ++j) {
        const int32_t i = hard_decide(thr, M, y[j]);
        idx[j] = i;
        const double F = mixture_cdf(a, p, M, s2, y[j]);
        n_hat[j] = sign[i] ? (FYt[i + 1] - F) / delta[i] : (F - FYt[i]) / delta[i];
        for (int k = 0; k < bps; ++k) bits[j * bps + k] = gray_bit(i, k);
    }
}

void emu_direct(int bps, const double *a, double two_variance, const double *y, int64_t n, double *llr)
{
    for (int64_t j = 0; j < n; ++j) direct_llr(a, 1 << bps, bps, two_variance, y[j], llr + j * bps);
}

}  // extern "C"
