// QR_SCHED_FUSED: persistent cooperative kernel running the fused flooding iteration of
// qr_decode_fused.cuh (reference: Decoder._decode, decoder.pyx:391-436).
//
// One step = fused phase (all tiles, work-stealing claims handed out in tile order so the grid sweeps
// one L2-sized tile at a time) -> the LAST CTA to finish it advances the lane state machine for all lanes
// -> grid barrier -> (only if a frame finished) refill phase -> grid barrier.
//
// Thread mapping of the fused phase: bx = TL / VEC threads share a check (together they move one contiguous
// TL * w byte row per access), a warp is 32 / bx checks wide, and WARPS claim work from one counter (see
// fused_phase); fp32: 128 registers, two CTAs of 256 threads per SM.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

// register-lean fused item: row loads issued in batches of FUSED_BATCH_EDGES edges (4 loads each), no index
// prefetch into registers.  Measured on B200 (2048 frames x 50 iterations, 1024 lanes, 128 registers, 2 CTAs/SM):
// batches of 3 edges 63.7 ms, 4 edges 59.8 ms, 5 edges 67.3 ms, all 6 at once 64.5 ms (spills); 4 it is.
#define FUSED_HALF_BATCH 1
#define FUSED_BATCH_EDGES 4
#include "qr_decode_fused.cuh"
#include "qr_handles.h"

namespace cg = cooperative_groups;

namespace qr {

constexpr int kFBlock = 256;
constexpr int kFusedMaxTile = 128;

__device__ __forceinline__ uint64_t l2_policy(int kind)
{
    uint64_t p;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// Work distribution: WARPS claim work, not CTAs.  A warp is 32 / bx checks wide (bx = TL / VEC threads share a
// check) and takes rows_per_claim passes per claim from one global counter, in tile order; the atomic of the
// NEXT claim is issued before the current one is processed.  No CTA barrier inside the phase, so the warps of
// an SM drift apart and cover each other's memory round trips (measured: a CTA-wide claim with its barrier cost
// 2.1 us per claim, profiles/r1_fused_experiments.txt f7/f8).  Per-lane "some check unsatisfied" flags are
// OR-reduced inside the warp and stored per tile (idempotent stores of 1).
template <typename T, int VEC, int DSEL>
__device__ __forceinline__ void fused_phase(const FusedParams<T> &F, int cur, uint64_t pol_ld, uint64_t pol_st)
{
    const DecodeParams<T> &P = F.P;
    const int32_t bx = F.tl / VEC;                    // threads per check row (<= 32)
    const int32_t lane = threadIdx.x & 31;
    const int32_t tx = lane % bx, tyw = lane / bx, wy = 32 / bx;
    const int32_t claim_rows = wy * F.rows_per_claim, C = (int32_t)P.C;
    const int32_t cpt = (C + claim_rows - 1) / claim_rows, total = cpt * F.tiles;
    int32_t tile_prev = -1;
    uint32_t bad = 0;
    LaneInfo<VEC> L;
    L.active = 0;
    TileView<T> V;
    const int32_t minfin = ld_stream(&P.ctrl[CTRL_MINFIN]);   // fixed for the whole phase (updated between phases)
    auto flush = [&]() {
        uint32_t b = bad;
        for (int32_t o = bx; o < 32; o <<= 1) b |= __shfl_xor_sync(0xffffffffu, b, o);
        if (tyw == 0) {
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                if (b >> k & 1) P.unsat[cur][tile_prev * F.tl + tx * VEC + k] = 1;
        }
        bad = 0;
    };
    // static_share (per mille) of the claims of a step is dealt round-robin to the warps without any atomic
    // (claims gw, gw + W, ...); the rest is claimed dynamically so that a slow SM does not hold the barrier
    const int32_t W = (int32_t)(gridDim.x * (kFBlock / 32)), gw = (int32_t)(blockIdx.x * (kFBlock / 32) + (threadIdx.x >> 5));
    const int32_t n_static = (int32_t)((int64_t)total * F.static_share / 1000) / W * W;
    int32_t nxt = gw < n_static ? gw : -1;
    if (nxt < 0) {
        if (lane == 0) nxt = n_static + atomicAdd(&P.work[0], 1);
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
    }
    for (;;) {
        const int32_t cl = nxt;
        if (cl >= total) break;
        const bool dyn = cl + W >= n_static;                   // the next claim comes from the counter
        if (!dyn) nxt = cl + W;
        else if (lane == 0) nxt = n_static + atomicAdd(&P.work[0], 1);   // in flight while this claim is processed
        const int32_t tile = cl / cpt;
        if (tile != tile_prev) {
            if (tile_prev >= 0) flush();
            L = load_lane_info<T, VEC>(P, cur, (tile * F.tl) / VEC + tx);
            mark_post_lanes<T, VEC>(F, L, minfin);
            V = tile_view(F, cur, tile);
            tile_prev = tile;
        }
        if (L.active) {
            const int32_t c0 = (cl - tile * cpt) * claim_rows, c1 = min(c0 + claim_rows, C);
            for (int32_t b = 0; b < P.n_bins; ++b) {
                const CheckBin bin = P.bins[b];
                const int32_t lo = max(c0, bin.chk_begin), hi = min(c1, bin.chk_begin + bin.count);
                if (lo >= hi) continue;
                const CheckBin sub{bin.degree, lo, hi - lo, bin.slot_begin + (lo - bin.chk_begin) * bin.degree};
                if constexpr (DSEL > 0) bad |= run_fused_bin<T, VEC, DSEL>(V, F.nbr, L, tx * VEC, sub, tyw, wy, pol_ld, pol_st);
                else bad |= run_fused_bin_any<T, VEC>(V, F.nbr, L, tx * VEC, sub, tyw, wy, pol_ld, pol_st);
            }
        }
        if (dyn) nxt = __shfl_sync(0xffffffffu, nxt, 0);
    }
    if (tile_prev >= 0) flush();
}

// REFILL PHASE of the fused schedule.  `buf` = lane-state buffer (and refill list) the bookkeeping of this
// step wrote; `cur` = message buffer the step READ (finished frames' posteriors are llr + its columns).
// Work item = (listed lane, block of kFBlock * kFRefillRows rows), all loads of a thread issued before its stores.
constexpr int kFRefillRows = 4;

template <typename T>
__device__ __forceinline__ void fused_refill_phase(const FusedParams<T> &F, int buf, int cur)
{
    const DecodeParams<T> &P = F.P;
    const int32_t cnt = ld_stream(&P.ctrl[CTRL_REFILL_CNT + buf]);
    const int32_t span = kFBlock * kFRefillRows;
    const int32_t vitems = (int32_t)((P.N + span - 1) / span), citems = (int32_t)((P.C + span - 1) / span);
    const int64_t total = (int64_t)cnt * (vitems + citems);
    const int32_t tl = F.tl;
    for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
        const int32_t e = (int32_t)(item % cnt), r = (int32_t)(item / cnt);
        const int32_t lane = ld_stream(&P.refill_list[(int64_t)buf * P.lanes + e]);
        const LaneState s = ld_stream(&P.st[buf][lane]);
        const int32_t tile = lane / tl, lt = lane % tl;
        if (r >= vitems) {
            if (s.frame >= 0 && s.fresh) {
                const int32_t c0 = (r - vitems) * span + threadIdx.x;
                uint8_t sy[kFRefillRows];
#pragma unroll
                for (int i = 0; i < kFRefillRows; ++i) {
                    const int32_t ci = c0 + i * kFBlock;
                    if (ci < P.C) sy[i] = P.synd_in[(int64_t)s.frame * P.C + P.chk_order[ci]];
                }
#pragma unroll
                for (int i = 0; i < kFRefillRows; ++i) {
                    const int32_t ci = c0 + i * kFBlock;
                    if (ci < P.C) P.synd[((int64_t)tile * P.C + ci) * tl + lt] = sy[i];
                }
            }
            continue;
        }
        const int32_t n0 = r * span + threadIdx.x;
        T *llr_col = P.llr + (int64_t)tile * P.N * tl + lt;
        const bool post_valid = F.post && s.retire >= 0 && ld_stream(&F.postok[(int64_t)buf * P.lanes + lane]) != 0;
        if (s.retire >= 0 && P.post_out && post_valid && ld_stream(&P.iters[s.retire]) != 0) {
            // the phase stored this frame's posteriors: one column of N elements
            const T *pc = F.post + (int64_t)tile * P.N * tl + lt;
            T v[kFRefillRows];
#pragma unroll
            for (int i = 0; i < kFRefillRows; ++i) {
                const int32_t n = n0 + i * kFBlock;
                if (n < P.N) v[i] = ld_stream(&pc[(int64_t)n * tl]);
            }
#pragma unroll
            for (int i = 0; i < kFRefillRows; ++i) {
                const int32_t n = n0 + i * kFBlock;
                if (n < P.N) store_output_llr(P.post_out, P.post_out_f64, (int64_t)s.retire * P.N + n, (double)v[i]);
            }
        } else if (s.retire >= 0 && P.post_out) {
            if (ld_stream(&P.iters[s.retire]) == 0) {
                LaneState only_retire = s;
                only_retire.frame = -1;
                for (int i = 0; i < kFRefillRows; ++i) {
                    const int32_t n = n0 + i * kFBlock;
                    if (n < P.N) fused_refill_var_elem<T>(F, cur, only_retire, lane, n);
                }
            } else {
                const T *c = F.c2v[cur] + (int64_t)tile * P.E * tl + lt;
                int32_t sl[kFRefillRows][3];
                T v[kFRefillRows][4];
#pragma unroll
                for (int i = 0; i < kFRefillRows; ++i) {
                    const int32_t n = n0 + i * kFBlock;
                    if (n < P.N) {
#pragma unroll
                        for (int j = 0; j < 3; ++j) sl[i][j] = P.var_slot[3 * n + j];
                    }
                }
#pragma unroll
                for (int i = 0; i < kFRefillRows; ++i) {
                    const int32_t n = n0 + i * kFBlock;
                    if (n < P.N) {
                        v[i][0] = ld_stream(&llr_col[(int64_t)n * tl]);
#pragma unroll
                        for (int j = 0; j < 3; ++j) v[i][1 + j] = ld_stream(&c[(int64_t)sl[i][j] * tl]);
                    }
                }
#pragma unroll
                for (int i = 0; i < kFRefillRows; ++i) {
                    const int32_t n = n0 + i * kFBlock;
                    if (n < P.N) {
                        T acc = v[i][0] + v[i][1];     // decoder.pyx:291-293, ascending edge id
                        acc = acc + v[i][2];
                        acc = acc + v[i][3];
                        store_output_llr(P.post_out, P.post_out_f64, (int64_t)s.retire * P.N + n, (double)acc);
                    }
                }
            }
        }
        if (s.frame >= 0 && s.fresh) {
            T v[kFRefillRows];
#pragma unroll
            for (int i = 0; i < kFRefillRows; ++i) {
                const int32_t n = n0 + i * kFBlock;
                if (n < P.N) v[i] = load_input_llr<T>(P.llr_in, P.llr_in_f64, (int64_t)s.frame * P.N + n);
            }
#pragma unroll
            for (int i = 0; i < kFRefillRows; ++i) {
                const int32_t n = n0 + i * kFBlock;
                if (n < P.N) llr_col[(int64_t)n * tl] = v[i];
            }
        }
    }
}

template <typename T, int VEC, int DSEL>
__global__ void __launch_bounds__(kFBlock, DSEL > 0 ? 2 : 1) k_fused(FusedParams<T> F)
{
    __shared__ int32_t s_last;
    const DecodeParams<T> &P = F.P;
    cg::grid_group grid = cg::this_grid();
    const uint64_t pol_ld = l2_policy(F.hints >= 2 ? 2 : 0), pol_st = l2_policy(F.hints >= 1 ? 1 : 0);
    fused_refill_phase<T>(F, 0, 0);      // first generation of frames into the lanes
    grid.sync();
    for (int step = 0;; ++step) {
        const int cur = step & 1;
        fused_phase<T, VEC, DSEL>(F, cur, pol_ld, pol_st);
        // the last CTA to get here advances the lane state machine (decoder.pyx:431-436) for every lane
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(&P.ctrl[CTRL_ARRIVE], 1) == (int32_t)gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (threadIdx.x == 0) {
                P.ctrl[CTRL_ARRIVE] = 0;
                P.work[0] = 0;
                P.ctrl[CTRL_REFILL_CNT + (cur ^ 1)] = 0;
                atomicAdd(&P.stats[1], 1ULL);
            }
            __syncthreads();
            const int32_t minfin = ld_stream(&P.ctrl[CTRL_MINFIN]);    // what the phase of this step went by
            for (int32_t jv = threadIdx.x; jv < P.lanes / VEC; jv += kFBlock) {
                LaneInfo<VEC> L = load_lane_info<T, VEC>(P, cur, jv);
                decide_lanes<T, VEC>(P, cur, L);
                fused_note_finishers<T, VEC>(F, cur ^ 1, L, minfin);
                bookkeep_lanes<T, VEC>(P, cur, step, L);
            }
            __syncthreads();
            if (threadIdx.x == 0) P.ctrl[CTRL_MINFIN] = *(volatile int32_t *)&P.ctrl[CTRL_MINFIN_NEXT];
        }
        grid.sync();
        if (*(volatile int32_t *)&P.ctrl[CTRL_FIN_STEP] == step) {
            fused_refill_phase<T>(F, cur ^ 1, cur);
            grid.sync();
        }
        if (*(volatile int32_t *)&P.ctrl[CTRL_REMAINING] <= 0) break;
    }
}

template <typename T, int VEC, int DSEL>
static int launch_fused(qr_decoder *d, const FusedParams<T> &F, cudaStream_t stream)
{
    const size_t smem = 0;
    auto kern = k_fused<T, VEC, DSEL>;
    int per_sm = 0;
    QR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFBlock, smem));
    if (per_sm < 1) return fail(QR_ERR_CUDA, "fused decoder kernel does not fit on an SM");
    const int grid = per_sm * d->sm_count;
    d->coop_grid = grid;
    FusedParams<T> Fc = F;
    void *args[] = {&Fc};
    QR_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(kFBlock), args, smem, stream));
    return QR_OK;
}

bool fused_eligible(const qr_graph *g)
{
    return g->d_slot_nbr != nullptr && g->max_cdeg <= kFusedMaxCheckDegree;
}

template <typename T, int VEC>
static int run_batch_fused_v(qr_decoder *d, const DecodeParams<T> &P, cudaStream_t stream)
{
    const qr_graph *g = d->g;
    if (!fused_eligible(g))
        return fail(QR_ERR_INVALID, "fused schedule needs every variable of degree 3 and check degrees <= 8");
    const size_t w = sizeof(T);
    if (!d->c2v2) {
        QR_CUDA_CHECK(cudaMalloc(&d->c2v2, (size_t)g->E * (size_t)d->lanes * w));
        QR_CUDA_CHECK(cudaMemset(d->c2v2, 0, (size_t)g->E * (size_t)d->lanes * w));
    }
    FusedParams<T> F;
    F.P = P;
    F.nbr = reinterpret_cast<const Nbr4 *>(g->d_slot_nbr);
    F.c2v[0] = static_cast<T *>(d->c2v);
    F.c2v[1] = static_cast<T *>(d->c2v2);
    F.post = d->fused_store_post ? static_cast<T *>(d->post) : nullptr;
    F.postok = d->postok;
    int32_t tl = d->fused_tile > 0 ? d->fused_tile : 32;
    tl = std::min<int32_t>(tl, kFusedMaxTile);
    while (tl > 32 && (P.lanes % tl || (tl / VEC) > 32)) tl /= 2;   // a check row is shared by at most one warp
    F.tl = tl;
    F.tiles = P.lanes / tl;
    F.hints = d->fused_hints;
    F.prefetch = d->fused_prefetch;
    F.rows_per_claim = d->fused_rpc > 0 ? d->fused_rpc : 4;
    F.static_share = d->fused_static;
    if (d->regular_degree == 6) return launch_fused<T, VEC, 6>(d, F, stream);
    return launch_fused<T, VEC, 0>(d, F, stream);
}

template <typename T>
int run_batch_fused(qr_decoder *d, const DecodeParams<T> &P, cudaStream_t stream)
{
    return run_batch_fused_v<T, 16 / sizeof(T)>(d, P, stream);
}

template int run_batch_fused<float>(qr_decoder *, const DecodeParams<float> &, cudaStream_t);
template int run_batch_fused<double>(qr_decoder *, const DecodeParams<double> &, cudaStream_t);

}  // namespace qr
