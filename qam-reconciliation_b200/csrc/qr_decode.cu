// Decoder kernels and batch orchestration (reference: Decoder._decode, decoder.pyx:391-436).
//
// Two schedules run the same phase code (qr_decode_core.cuh):
//   QR_SCHED_PERSISTENT  one cooperative kernel per batch; CTAs stay resident, phases are separated
//                        by grid barriers, finished lanes pull the next frame of the batch.
//   QR_SCHED_LAUNCH      one check kernel + one variable kernel per step (debug / comparison).
//
// Thread mapping of a phase: a CTA of 256 threads is a (bx x by) tile, bx along lane-vectors
// (tx fastest, so a warp touches contiguous bytes of a row), by along nodes.  The grid is split
// into `gx` lane tiles times `gy` node groups; each thread keeps its lanes for the whole phase, so
// lane state is read once and the per-lane "some check unsatisfied" flag is reduced in registers,
// then in shared memory, and reaches global memory once per CTA and lane.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <new>

#include "qr_decode_core.cuh"
#include "qr_handles.h"

namespace cg = cooperative_groups;

namespace qr {

constexpr int kBlock = 256;

// resident CTAs per SM the kernels are compiled for (caps registers per thread at 65536 / (256 * n))
template <typename T, int VEC>
constexpr int min_ctas()
{
    return sizeof(T) == 8 ? 1 : (VEC >= 4 ? 2 : (VEC == 2 ? 3 : 4));
}

struct Tiling {
    int32_t bx, by;   // CTA tile
    int32_t nxt;      // lane tiles = (lanes / VEC) / bx
};

// DSEL > 0: the graph is check-regular with that degree (only that code path is compiled in, so the
// register allocation is the one of the hot case); DSEL == 0: any mix of degrees.
template <typename T, int VEC, int DSEL>
__device__ __forceinline__ void check_phase(const DecodeParams<T> &P, int cur, const Tiling tl,
                                            int32_t *s_flags)
{
    const int32_t tx = threadIdx.x % tl.bx, ty = threadIdx.x / tl.bx;
    const int32_t G = gridDim.x;
    const int32_t gx = min(tl.nxt, G), gy = G / gx;
    const int32_t bxid = blockIdx.x % gx, byid = blockIdx.x / gx;
    if (byid >= gy) return;  // left-over CTAs (grid not a multiple of gx): no work, no barrier inside
    for (int32_t xt = bxid; xt < tl.nxt; xt += gx) {
        const int32_t jv = xt * tl.bx + tx;
        const LaneInfo<VEC> L = load_lane_info<T, VEC>(P, cur, jv);
        if (blockIdx.x == 0 && threadIdx.x == 0) P.ctrl[CTRL_REFILL_CNT + (cur ^ 1)] = 0;   // list the variable phase fills
        uint32_t bad = 0;
        if (L.active) {
            const int32_t first = byid * tl.by + ty, stride = gy * tl.by;
            for (int32_t b = 0; b < P.n_bins; ++b) {
                const CheckBin bin = P.bins[b];
                if (DSEL > 0) {
                    bad |= run_check_bin<T, VEC, DSEL>(P, L, bin, first, stride);
                } else {
                    switch (bin.degree) {
                    case 2: bad |= run_check_bin<T, VEC, 2>(P, L, bin, first, stride); break;
                    case 3: bad |= run_check_bin<T, VEC, 3>(P, L, bin, first, stride); break;
                    case 4: bad |= run_check_bin<T, VEC, 4>(P, L, bin, first, stride); break;
                    case 5: bad |= run_check_bin<T, VEC, 5>(P, L, bin, first, stride); break;
                    case 6: bad |= run_check_bin<T, VEC, 6>(P, L, bin, first, stride); break;
                    case 7: bad |= run_check_bin<T, VEC, 7>(P, L, bin, first, stride); break;
                    case 8: bad |= run_check_bin<T, VEC, 8>(P, L, bin, first, stride); break;
                    default: bad |= run_check_bin<T, VEC, 0>(P, L, bin, first, stride); break;
                    }
                }
            }
        }
        // reduce the per-lane flags over the CTA's node rows, then one global store per lane
        for (int32_t i = threadIdx.x; i < tl.bx * VEC; i += blockDim.x) s_flags[i] = 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < VEC; ++k)
            if (bad >> k & 1) s_flags[tx * VEC + k] = 1;
        __syncthreads();
        if (ty == 0) {
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                if (s_flags[tx * VEC + k]) P.unsat[cur][L.l0 + k] = 1;
        }
        __syncthreads();
    }
}

template <typename T, int VEC, bool UNROLLED>
__device__ __forceinline__ void var_phase(const DecodeParams<T> &P, int cur, int32_t step, const Tiling tl)
{
    const int32_t tx = threadIdx.x % tl.bx, ty = threadIdx.x / tl.bx;
    const int32_t G = gridDim.x;
    const int32_t gx = min(tl.nxt, G), gy = G / gx;
    const int32_t bxid = blockIdx.x % gx, byid = blockIdx.x / gx;
    if (byid >= gy) return;
    for (int32_t xt = bxid; xt < tl.nxt; xt += gx) {
        const int32_t jv = xt * tl.bx + tx;
        LaneInfo<VEC> L = load_lane_info<T, VEC>(P, cur, jv);
        decide_lanes<T, VEC>(P, cur, L);
        run_var_range<T, VEC, UNROLLED>(P, L, byid * tl.by + ty, gy * tl.by, (int32_t)P.N);
        if (byid == 0 && ty == 0) bookkeep_lanes<T, VEC>(P, cur, step, L);
    }
}

// ---- work-stealing variants for the persistent kernel: rows are claimed in chunks of by*kRowsPerClaim from
// a per-lane-tile counter, so a slow SM (far L2 die, unlucky DRAM pages) does not hold the grid barrier.
constexpr int kRowsPerClaim = 8;
// (irregular graphs, binned variable phase: a claim also pays the pipeline prologue of every degree bin it touches)
constexpr int kVarRowsPerClaimBinned = 32;

// Rows per claim: at least by * rows_min, and large enough that a CTA makes about 16 claims per phase -- every claim
// costs two CTA barriers and an atomic round trip (config 4, 944k checks: 100 claims per CTA and phase at the
// minimum size, 11 % of the phase).
__device__ __forceinline__ int32_t claim_rows(int32_t rows, int32_t ctas_per_lane_tile, int32_t by, int32_t rows_min)
{
    const int32_t want = rows / max(1, ctas_per_lane_tile * 16);
    return max(by * rows_min, (want + by - 1) / by * by);
}

template <typename T, int VEC, int DSEL>
__device__ __forceinline__ void check_phase_dyn(const DecodeParams<T> &P, int cur, const Tiling tl,
                                                int32_t *s_flags, int32_t *s_base)
{
    const int32_t tx = threadIdx.x % tl.bx, ty = threadIdx.x / tl.bx;
    const int32_t G = gridDim.x;
    const int32_t gx = min(tl.nxt, G);
    const int32_t bxid = blockIdx.x % gx, byid = blockIdx.x / gx;
    const int32_t C = (int32_t)P.C;
    const int32_t claim = claim_rows(C, G / gx, tl.by, kRowsPerClaim);
    for (int32_t xt = bxid; xt < tl.nxt; xt += gx) {
        const int32_t jv = xt * tl.bx + tx;
        const LaneInfo<VEC> L = load_lane_info<T, VEC>(P, cur, jv);
        if (byid == 0 && threadIdx.x == 0) P.work[kMaxLaneTiles + xt] = 0;   // variable-phase counter of this step
        if (blockIdx.x == 0 && threadIdx.x == 0) P.ctrl[CTRL_REFILL_CNT + (cur ^ 1)] = 0;   // list the variable phase fills
        uint32_t bad = 0;
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) *s_base = atomicAdd(&P.work[xt], claim);
            __syncthreads();
            const int32_t c0 = *s_base;
            if (c0 >= C) break;
            const int32_t c1 = min(c0 + claim, C);
            if (!L.active) continue;
            for (int32_t b = 0; b < P.n_bins; ++b) {
                const CheckBin bin = P.bins[b];
                const int32_t lo = max(c0, bin.chk_begin), hi = min(c1, bin.chk_begin + bin.count);
                if (lo >= hi) continue;
                const CheckBin sub{bin.degree, lo, hi - lo, bin.slot_begin + (lo - bin.chk_begin) * bin.degree};
                if (DSEL > 0) {
                    bad |= run_check_bin<T, VEC, DSEL>(P, L, sub, ty, tl.by);
                } else {
                    switch (bin.degree) {
                    case 2: bad |= run_check_bin<T, VEC, 2>(P, L, sub, ty, tl.by); break;
                    case 3: bad |= run_check_bin<T, VEC, 3>(P, L, sub, ty, tl.by); break;
                    case 4: bad |= run_check_bin<T, VEC, 4>(P, L, sub, ty, tl.by); break;
                    case 5: bad |= run_check_bin<T, VEC, 5>(P, L, sub, ty, tl.by); break;
                    case 6: bad |= run_check_bin<T, VEC, 6>(P, L, sub, ty, tl.by); break;
                    case 7: bad |= run_check_bin<T, VEC, 7>(P, L, sub, ty, tl.by); break;
                    case 8: bad |= run_check_bin<T, VEC, 8>(P, L, sub, ty, tl.by); break;
                    default: bad |= run_check_bin<T, VEC, 0>(P, L, sub, ty, tl.by); break;
                    }
                }
            }
        }
        for (int32_t i = threadIdx.x; i < tl.bx * VEC; i += blockDim.x) s_flags[i] = 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < VEC; ++k)
            if (bad >> k & 1) s_flags[tx * VEC + k] = 1;
        __syncthreads();
        if (ty == 0) {
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                if (s_flags[tx * VEC + k]) P.unsat[cur][L.l0 + k] = 1;
        }
        __syncthreads();
    }
}

template <typename T, int VEC, bool UNROLLED>
__device__ __forceinline__ void var_phase_dyn(const DecodeParams<T> &P, int cur, int32_t step, const Tiling tl,
                                              int32_t *s_base)
{
    const int32_t tx = threadIdx.x % tl.bx, ty = threadIdx.x / tl.bx;
    const int32_t G = gridDim.x;
    const int32_t gx = min(tl.nxt, G);
    const int32_t bxid = blockIdx.x % gx, byid = blockIdx.x / gx;
    const int32_t N = (int32_t)P.N;
    const int32_t claim = claim_rows(N, G / gx, tl.by, (UNROLLED && P.var_bins) ? kVarRowsPerClaimBinned : kRowsPerClaim);
    for (int32_t xt = bxid; xt < tl.nxt; xt += gx) {
        const int32_t jv = xt * tl.bx + tx;
        LaneInfo<VEC> L = load_lane_info<T, VEC>(P, cur, jv);
        decide_lanes<T, VEC>(P, cur, L);
        if (byid == 0 && threadIdx.x == 0) P.work[xt] = 0;                    // check-phase counter of the next step
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) *s_base = atomicAdd(&P.work[kMaxLaneTiles + xt], claim);
            __syncthreads();
            const int32_t n0 = *s_base;
            if (n0 >= N) break;
            if (UNROLLED && P.var_bins) run_var_binned<T, VEC>(P, L, n0, min(n0 + claim, N), ty, tl.by);
            else run_var_range<T, VEC, UNROLLED>(P, L, n0 + ty, tl.by, min(n0 + claim, N));
        }
        if (byid == 0 && ty == 0) bookkeep_lanes<T, VEC>(P, cur, step, L);
    }
}

// REFILL PHASE: ship finished frames' posterior columns, load newly assigned frames' LLR and syndrome
// columns.  `buf` = the state buffer (and refill list) the variable phase of this step wrote.
// Work item = (listed lane, block of kBlock*kRefillRows rows); a thread moves kRefillRows rows with all its
// loads issued before the stores (the column side of every access is one 32-byte sector per element, the
// caller's rows are contiguous across the threads of a warp).
constexpr int kRefillRows = 8;

template <typename T>
__device__ __forceinline__ void refill_phase(const DecodeParams<T> &P, int buf)
{
    const int32_t cnt = ld_stream(&P.ctrl[CTRL_REFILL_CNT + buf]);
    const int32_t span = kBlock * kRefillRows;
    const int32_t vitems = (int32_t)((P.N + span - 1) / span), citems = (int32_t)((P.C + span - 1) / span);
    const int64_t total = (int64_t)cnt * (vitems + citems);
    const int32_t lanes = P.lanes;
    for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
        // row-block-major order: CTAs resident at the same time serve the SAME rows of neighbouring listed lanes,
        // so when whole 8-lane sectors refill together (frames finishing in lock step) the sector traffic of the
        // lane-interleaved rows is shared through L2 instead of being paid once per lane
        const int32_t e = (int32_t)(item % cnt), r = (int32_t)(item / cnt);
        const int32_t lane = ld_stream(&P.refill_list[(int64_t)buf * lanes + e]);
        const LaneState s = ld_stream(&P.st[buf][lane]);
        if (r >= vitems) {
            if (s.frame >= 0 && s.fresh) {
                const int32_t c0 = (r - vitems) * span + threadIdx.x;
                uint8_t sy[kRefillRows];
#pragma unroll
                for (int i = 0; i < kRefillRows; ++i) {
                    const int32_t ci = c0 + i * kBlock;
                    if (ci < P.C) sy[i] = P.synd_in[(int64_t)s.frame * P.C + P.chk_order[ci]];
                }
#pragma unroll
                for (int i = 0; i < kRefillRows; ++i) {
                    const int32_t ci = c0 + i * kBlock;
                    if (ci < P.C) P.synd[(int64_t)ci * lanes + lane] = sy[i];
                }
            }
            continue;
        }
        const int32_t n0 = r * span + threadIdx.x;
        if (s.retire >= 0 && P.post_out) {
            if (ld_stream(&P.iters[s.retire]) == 0) {
                // frames that never iterated: rare, element-wise path (input copy / input + 0.0 semantics)
                LaneState only_retire = s;
                only_retire.frame = -1;
                for (int i = 0; i < kRefillRows; ++i) {
                    const int32_t n = n0 + i * kBlock;
                    if (n < P.N) refill_var_elem<T>(P, only_retire, lane, n);
                }
            } else {
                T v[kRefillRows];
#pragma unroll
                for (int i = 0; i < kRefillRows; ++i) {
                    const int32_t n = n0 + i * kBlock;
                    if (n < P.N) v[i] = ld_stream(&P.post[(int64_t)n * lanes + lane]);
                }
#pragma unroll
                for (int i = 0; i < kRefillRows; ++i) {
                    const int32_t n = n0 + i * kBlock;
                    if (n < P.N) store_output_llr(P.post_out, P.post_out_f64, (int64_t)s.retire * P.N + n, (double)v[i]);
                }
            }
        }
        if (s.frame >= 0 && s.fresh) {
            T v[kRefillRows];
#pragma unroll
            for (int i = 0; i < kRefillRows; ++i) {
                const int32_t n = n0 + i * kBlock;
                if (n < P.N) v[i] = load_input_llr<T>(P.llr_in, P.llr_in_f64, (int64_t)s.frame * P.N + n);
            }
#pragma unroll
            for (int i = 0; i < kRefillRows; ++i) {
                const int32_t n = n0 + i * kBlock;
                if (n < P.N) {
                    P.llr[(int64_t)n * lanes + lane] = v[i];
                    P.post[(int64_t)n * lanes + lane] = v[i];
                }
            }
        }
    }
}

template <typename T, int VEC, int DSEL>
__global__ void __launch_bounds__(kBlock, min_ctas<T, VEC>()) k_check(DecodeParams<T> P, int step, Tiling tl)
{
    __shared__ int32_t s_flags[32 * VEC];
    // Nothing changes CTRL_REMAINING while a check kernel runs, so every CTA sees the same value; the
    // variable kernel of this step must NOT read it (its own bookkeeping threads decrement it while
    // later CTAs are still starting), it reads the snapshot taken here instead.
    const int32_t remaining = *(volatile int32_t *)&P.ctrl[CTRL_REMAINING];
    if (blockIdx.x == 0 && threadIdx.x == 0) P.ctrl[CTRL_SNAPSHOT] = remaining;
    if (remaining == 0) return;
    check_phase<T, VEC, DSEL>(P, step & 1, tl, s_flags);
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock, min_ctas<T, VEC>()) k_var(DecodeParams<T> P, int step, Tiling tl)
{
    if (*(volatile int32_t *)&P.ctrl[CTRL_SNAPSHOT] == 0) return;
    var_phase<T, VEC, true>(P, step & 1, step, tl);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.stats[1], 1ULL);
}

// step < 0: the initial load of the first generation of frames; otherwise only if a frame finished in `step`
template <typename T>
__global__ void __launch_bounds__(kBlock) k_refill(DecodeParams<T> P, int step)
{
    if (step >= 0 && *(volatile int32_t *)&P.ctrl[CTRL_FIN_STEP] != step) return;
    refill_phase<T>(P, step < 0 ? 0 : (step & 1) ^ 1);
}

template <typename T, int VEC, int DSEL>
__global__ void __launch_bounds__(kBlock, min_ctas<T, VEC>()) k_persistent(DecodeParams<T> P, Tiling tl)
{
    __shared__ int32_t s_flags[32 * VEC];
    __shared__ int32_t s_base;
    cg::grid_group grid = cg::this_grid();
    refill_phase<T>(P, 0);      // first generation of frames into the lanes
    grid.sync();
    for (int step = 0;; ++step) {
        check_phase_dyn<T, VEC, DSEL>(P, step & 1, tl, s_flags, &s_base);
        grid.sync();
        var_phase_dyn<T, VEC, DSEL == 0>(P, step & 1, step, tl, &s_base);
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.stats[1], 1ULL);
        grid.sync();
        if (*(volatile int32_t *)&P.ctrl[CTRL_FIN_STEP] == step) {   // a frame finished: ship it, refill its lane
            refill_phase<T>(P, (step & 1) ^ 1);
            grid.sync();
        }
        if (*(volatile int32_t *)&P.ctrl[CTRL_REMAINING] <= 0) break;
    }
}

// lane l starts on frame l (if there is one); everything else is idle
__global__ void k_init_batch(LaneState *st0, LaneState *st1, int32_t *unsat0, int32_t *unsat1,
                             int32_t lanes, int64_t frames, int32_t *ctrl, int32_t *work,
                             unsigned long long *stats, int32_t *refill_list)
{
    const int32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < lanes) {
        LaneState s;
        s.frame = (int64_t)l < frames ? l : -1;
        s.iter = 0;
        s.fresh = s.frame >= 0;
        s.retire = -1;
        st0[l] = s;
        st1[l] = s;
        unsat0[l] = 0;
        unsat1[l] = 0;
        if (s.frame >= 0) refill_list[l] = l;      // first generation: lanes 0 .. min(lanes, frames) - 1
    }
    for (int32_t i = l; i < 2 * kMaxLaneTiles; i += gridDim.x * blockDim.x) work[i] = 0;
    if (l == 0) {
        ctrl[CTRL_NEXT_FRAME] = (int32_t)min((int64_t)lanes, frames);
        ctrl[CTRL_REMAINING] = (int32_t)frames;
        ctrl[CTRL_FIN_STEP] = -1;
        ctrl[CTRL_REFILL_CNT + 0] = (int32_t)min((int64_t)lanes, frames);
        ctrl[CTRL_REFILL_CNT + 1] = 0;
        stats[0] = 0;
        stats[1] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
struct Prec;
template <>
struct Prec<float> { static constexpr int VEC = 4; };
template <>
struct Prec<double> { static constexpr int VEC = 2; };

}  // namespace qr

namespace qr {

template <typename T>
static DecodeParams<T> make_params(const qr_decoder *d, const void *llr, int llr_dtype,
                                   const uint8_t *synd, int64_t frames, int32_t maxiter,
                                   uint8_t *success, int32_t *iters, void *post, int post_dtype)
{
    DecodeParams<T> P;
    const qr_graph *g = d->g;
    P.bins = g->d_bins;
    P.n_bins = (int32_t)g->bins.size();
    P.chk_order = g->d_chk_order;
    P.slot_var = g->d_slot_var;
    P.var_ptr = g->d_var_ptr;
    P.var_slot = g->d_var_slot;
    P.var_work = g->d_var_work;
    P.var_bins = g->d_var_bins; P.n_var_bins = (int32_t)g->var_bins.size(); P.vslot_sorted = g->d_vslot_sorted;
    P.N = g->N; P.C = g->C; P.E = g->E;
    P.var_deg = g->var_deg;
    // a batch smaller than the workspace uses a narrower layout, so no CTA is left with idle lanes only
    P.lanes = (int32_t)std::min<int64_t>(d->lanes, (frames + 31) / 32 * 32);
    P.c2v = static_cast<T *>(d->c2v);
    P.post = static_cast<T *>(d->post);
    P.llr = static_cast<T *>(d->llr);
    P.synd = d->synd;
    P.st[0] = d->st; P.st[1] = d->st + P.lanes;
    P.unsat[0] = d->unsat; P.unsat[1] = d->unsat + P.lanes;
    P.llr_in = llr; P.llr_in_f64 = llr_dtype == QR_F64;
    P.synd_in = synd;
    P.frames = frames; P.maxiter = maxiter;
    P.success = success; P.iters = iters;
    P.post_out = post; P.post_out_f64 = post_dtype == QR_F64;
    P.ctrl = d->ctrl; P.stats = d->stats; P.work = d->work; P.refill_list = d->refill_list;
    return P;
}

template <int VEC>
static Tiling make_tiling(int32_t lanes)
{
    Tiling tl;
    const int32_t lv = lanes / VEC;
    tl.bx = 32;
    while (lv % tl.bx) tl.bx >>= 1;
    tl.by = kBlock / tl.bx;
    tl.nxt = lv / tl.bx;
    return tl;
}

static int grid_for(int capacity, int32_t nxt)
{
    if (nxt >= capacity) return capacity;
    return (capacity / nxt) * nxt;
}

template <typename T, int DSEL, int VEC = Prec<T>::VEC>
static int run_batch_t(qr_decoder *d, const DecodeParams<T> &P, cudaStream_t stream)
{
    const Tiling tl = make_tiling<VEC>(P.lanes);
    // QR_SCHED_AUTO on a graph the fused schedule is not preferred for: the persistent kernel, except on very large
    // graphs, where its per-claim barriers and atomics cost more than two launches per iteration do (config 4,
    // 3.7 M edges: 5.09 TB/s launched against 4.69 TB/s persistent on B200)
    const bool launch_mode = d->schedule == QR_SCHED_LAUNCH || (d->schedule == QR_SCHED_AUTO && d->g->E >= (int64_t(1) << 21));
    if (!launch_mode) {
        int per_sm = 0;
        QR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_persistent<T, VEC, DSEL>,
                                                                    kBlock, 0));
        if (per_sm < 1) return fail(QR_ERR_CUDA, "persistent decoder kernel does not fit on an SM");
        const int grid = grid_for(per_sm * d->sm_count, tl.nxt);
        d->coop_grid = grid;
        DecodeParams<T> Pc = P;
        Tiling tlc = tl;
        void *args[] = {&Pc, &tlc};
        QR_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)k_persistent<T, VEC, DSEL>, dim3(grid),
                                                  dim3(kBlock), args, 0, stream));
        return QR_OK;
    }
    // launch-per-phase schedule: the host looks at the remaining-frames word every few steps
    const int grid = grid_for(4 * d->sm_count, tl.nxt);
    const int64_t rounds = (P.frames + P.lanes - 1) / P.lanes;
    const int64_t max_steps = rounds * ((int64_t)P.maxiter + 2) + 2;
    const int chunk = P.maxiter + 1;   // a lane finishes a frame at least every maxiter + 1 steps
    const int rgrid = 2 * d->sm_count;
    k_refill<T><<<rgrid, kBlock, 0, stream>>>(P, -1);
    for (int64_t step = 0; step < max_steps;) {
        for (int k = 0; k < chunk && step < max_steps; ++k, ++step) {
            k_check<T, VEC, DSEL><<<grid, kBlock, 0, stream>>>(P, (int)step, tl);
            k_var<T, VEC><<<grid, kBlock, 0, stream>>>(P, (int)step, tl);
            k_refill<T><<<rgrid, kBlock, 0, stream>>>(P, (int)step);
        }
        QR_CUDA_CHECK(cudaGetLastError());
        QR_CUDA_CHECK(cudaMemcpyAsync(d->h_ctrl, d->ctrl, CTRL_WORDS * sizeof(int32_t),
                                      cudaMemcpyDeviceToHost, stream));
        QR_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (d->h_ctrl[CTRL_REMAINING] <= 0) break;
    }
    return QR_OK;
}

template <typename T>
int run_batch_fused(qr_decoder *d, const DecodeParams<T> &P, cudaStream_t stream);   // qr_decode_fused.cu
bool fused_eligible(const qr_graph *g);
bool fused_preferred(const qr_graph *g, size_t w);

template <typename T>
static int run_batch(qr_decoder *d, const DecodeParams<T> &P, cudaStream_t stream)
{
    if (d->schedule == QR_SCHED_FUSED || (d->schedule == QR_SCHED_AUTO && fused_preferred(d->g, sizeof(T))))
        return run_batch_fused<T>(d, P, stream);
    if constexpr (sizeof(T) == 4) {
        // narrower lane vectors (more, lighter threads) -- experimental, QAMRECON_VEC=1|2
        if (d->regular_degree == 6 && d->vec == 1) return run_batch_t<T, 6, 1>(d, P, stream);
        if (d->regular_degree == 6 && d->vec == 2) return run_batch_t<T, 6, 2>(d, P, stream);
    }
    switch (d->regular_degree) {
    case 6: return run_batch_t<T, 6>(d, P, stream);
    default: return run_batch_t<T, 0>(d, P, stream);
    }
}

}  // namespace qr

extern "C" {

int qr_decoder_create(const qr_graph *g, int precision, int64_t lanes, qr_decoder **out)
{
    if (!out) return qr::fail(QR_ERR_INVALID, "null output pointer");
    *out = nullptr;
    if (!g) return qr::fail(QR_ERR_INVALID, "null graph");
    if (g->device < 0) return qr::fail(QR_ERR_INVALID, "graph was built without a device");
    if (!g->decodable)
        return qr::fail(QR_ERR_GRAPH, "the decoder needs every check node of degree 2..64 (graph built with qr_graph_create_any)");
    if (precision != QR_F32 && precision != QR_F64)
        return qr::fail(QR_ERR_INVALID, "precision must be QR_F32 or QR_F64");
    if (lanes < 0 || lanes > (1 << 20)) return qr::fail(QR_ERR_INVALID, "bad lane count");
    int prev = 0;
    cudaGetDevice(&prev);
    qr_decoder *d = new (std::nothrow) qr_decoder();
    if (!d) return qr::fail(QR_ERR_NOMEM, "out of host memory");
    auto body = [&]() -> int {
        QR_CUDA_CHECK(cudaSetDevice(g->device));
        cudaDeviceProp prop;
        QR_CUDA_CHECK(cudaGetDeviceProperties(&prop, g->device));
        if (!prop.cooperativeLaunch) return qr::fail(QR_ERR_CUDA, "device lacks cooperative launch");
        d->g = g; d->precision = precision; d->device = g->device;
        d->sm_count = prop.multiProcessorCount;
        const size_t w = precision == QR_F64 ? 8 : 4;
        const size_t per_lane = (size_t)g->E * w + 2 * (size_t)g->N * w + (size_t)g->C;
        if (lanes == 0) {
            // default: 512 frames in flight (measured optimum of the HBM-streaming regime on B200:
            // enough rows per phase to hide launch/barrier cost, workspace still small), less if
            // device memory is short
            size_t free_b = 0, total_b = 0;
            QR_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
            lanes = 512;
            while (lanes > 32 && (size_t)lanes * per_lane > free_b / 4) lanes /= 2;
        }
        lanes = (lanes + 31) / 32 * 32;
        d->lanes = (int32_t)lanes;
        d->regular_degree = (g->bins.size() == 1 && g->bins[0].degree == 6) ? 6 : 0;
        if (const char *v = getenv("QAMRECON_VEC")) d->vec = atoi(v);
        if (const char *v = getenv("QAMRECON_FUSED_TILE")) d->fused_tile = atoi(v);
        if (const char *v = getenv("QAMRECON_FUSED_HINTS")) d->fused_hints = atoi(v);
        if (const char *v = getenv("QAMRECON_FUSED_RPC")) d->fused_rpc = atoi(v);
        if (const char *v = getenv("QAMRECON_FUSED_PP_ITEMS")) d->fused_pp_items = atoi(v);
        if (const char *v = getenv("QAMRECON_FUSED_LEAN")) d->fused_lean = atoi(v);
        if (const char *v = getenv("QAMRECON_FUSED_PARK")) d->fused_park = std::max(0, atoi(v));
        if (const char *v = getenv("QAMRECON_FUSED_STORE_POST")) d->fused_store_post = atoi(v);
        const size_t L = (size_t)lanes;
        QR_CUDA_CHECK(cudaMalloc(&d->c2v, (size_t)g->E * L * w));
        QR_CUDA_CHECK(cudaMalloc(&d->post, (size_t)g->N * L * w));
        QR_CUDA_CHECK(cudaMalloc(&d->llr, (size_t)g->N * L * w));
        QR_CUDA_CHECK(cudaMalloc((void **)&d->synd, (size_t)g->C * L));
        QR_CUDA_CHECK(cudaMalloc((void **)&d->st, 2 * L * sizeof(qr::LaneState)));
        QR_CUDA_CHECK(cudaMalloc((void **)&d->unsat, 2 * L * sizeof(int32_t)));
        QR_CUDA_CHECK(cudaMalloc((void **)&d->ctrl, qr::CTRL_WORDS * sizeof(int32_t)));
        QR_CUDA_CHECK(cudaMalloc((void **)&d->stats, 2 * sizeof(unsigned long long)));
        QR_CUDA_CHECK(cudaMalloc((void **)&d->work, 2 * qr::kMaxLaneTiles * sizeof(int32_t)));
        QR_CUDA_CHECK(cudaMalloc((void **)&d->refill_list, 2 * L * sizeof(int32_t)));
        QR_CUDA_CHECK(cudaMemset(d->c2v, 0, (size_t)g->E * L * w));
        QR_CUDA_CHECK(cudaMemset(d->post, 0, (size_t)g->N * L * w));
        QR_CUDA_CHECK(cudaMemset(d->llr, 0, (size_t)g->N * L * w));
        QR_CUDA_CHECK(cudaMemset(d->synd, 0, (size_t)g->C * L));
        QR_CUDA_CHECK(cudaMemset(d->ctrl, 0, qr::CTRL_WORDS * sizeof(int32_t)));
        QR_CUDA_CHECK(cudaMemset(d->stats, 0, 2 * sizeof(unsigned long long)));
        QR_CUDA_CHECK(cudaHostAlloc((void **)&d->h_ctrl, qr::CTRL_WORDS * sizeof(int32_t), cudaHostAllocDefault));
        QR_CUDA_CHECK(cudaHostAlloc((void **)&d->h_stats, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
        return QR_OK;
    };
    int rc = body();
    cudaSetDevice(prev);
    if (rc != QR_OK) { qr_decoder_destroy(d); return rc; }
    *out = d;
    return QR_OK;
}

void qr_decoder_destroy(qr_decoder *d)
{
    if (!d) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(d->device);
    cudaFree(d->c2v); cudaFree(d->c2v2); cudaFree(d->post); cudaFree(d->llr); cudaFree(d->synd);
    cudaFree(d->st); cudaFree(d->unsat); cudaFree(d->ctrl); cudaFree(d->stats); cudaFree(d->work); cudaFree(d->refill_list); cudaFree(d->fused_ctl); cudaFree(d->fused_rlist); cudaFree(d->fused_ppq); cudaFree(d->fused_nbrl);
    cudaFree(d->pipe_buf);
    cudaFree(d->dev_buf);
    if (d->pipe_streams_ready) {
        cudaStreamDestroy(d->s_in); cudaStreamDestroy(d->s_out);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(d->ev_in[i]); cudaEventDestroy(d->ev_compute[i]); cudaEventDestroy(d->ev_out[i]); }
        cudaEventDestroy(d->ev_start);
    }
    if (d->h_ctrl) cudaFreeHost(d->h_ctrl);
    if (d->h_stats) cudaFreeHost(d->h_stats);
    cudaSetDevice(prev);
    delete d;
}

int qr_decoder_set_schedule(qr_decoder *d, int schedule)
{
    if (!d) return qr::fail(QR_ERR_INVALID, "null decoder");
    if (schedule != QR_SCHED_PERSISTENT && schedule != QR_SCHED_LAUNCH && schedule != QR_SCHED_FUSED &&
        schedule != QR_SCHED_AUTO)
        return qr::fail(QR_ERR_INVALID, "unknown schedule");
    if (schedule == QR_SCHED_FUSED && !qr::fused_eligible(d->g))
        return qr::fail(QR_ERR_INVALID, "fused schedule needs check degrees <= 8, variable degrees 1..64 and fewer than 2^27 variables");
    d->schedule = schedule;
    return QR_OK;
}

int qr_decode_batch(qr_decoder *d, const void *d_llr, int llr_dtype, const uint8_t *d_synd,
                    int64_t frames, int32_t max_iterations, uint8_t *d_success, int32_t *d_iters,
                    void *d_post, int post_dtype, void *stream_)
{
    if (!d) return qr::fail(QR_ERR_INVALID, "null decoder");
    if (frames < 0 || frames >= (int64_t(1) << 31)) return qr::fail(QR_ERR_INVALID, "bad frame count");
    if (max_iterations < 0) return qr::fail(QR_ERR_INVALID, "max_iterations must be >= 0");
    if (llr_dtype != QR_F32 && llr_dtype != QR_F64) return qr::fail(QR_ERR_INVALID, "bad llr dtype");
    if (d_post && post_dtype != QR_F32 && post_dtype != QR_F64)
        return qr::fail(QR_ERR_INVALID, "bad posterior dtype");
    if (frames == 0) return QR_OK;
    if (!d_llr || !d_synd || !d_success || !d_iters) return qr::fail(QR_ERR_INVALID, "null array");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int prev = 0;
    cudaGetDevice(&prev);
    auto body = [&]() -> int {
        QR_CUDA_CHECK(cudaSetDevice(d->device));
        const int32_t lanes = (int32_t)std::min<int64_t>(d->lanes, (frames + 31) / 32 * 32);
        qr::k_init_batch<<<(lanes + 255) / 256, 256, 0, stream>>>(
            d->st, d->st + lanes, d->unsat, d->unsat + lanes, lanes, frames, d->ctrl, d->work, d->stats, d->refill_list);
        QR_CUDA_CHECK(cudaGetLastError());
        d->last_stream = stream;
        if (d->precision == QR_F64) {
            auto P = qr::make_params<double>(d, d_llr, llr_dtype, d_synd, frames, max_iterations,
                                             d_success, d_iters, d_post, post_dtype);
            return qr::run_batch<double>(d, P, stream);
        }
        auto P = qr::make_params<float>(d, d_llr, llr_dtype, d_synd, frames, max_iterations, d_success,
                                        d_iters, d_post, post_dtype);
        return qr::run_batch<float>(d, P, stream);
    };
    int rc = body();
    cudaSetDevice(prev);
    return rc;
}

int qr_decoder_last_stats(qr_decoder *d, int64_t *frame_iterations, int64_t *steps)
{
    if (!d) return qr::fail(QR_ERR_INVALID, "null decoder");
    int prev = 0;
    cudaGetDevice(&prev);
    auto body = [&]() -> int {
        QR_CUDA_CHECK(cudaSetDevice(d->device));
        QR_CUDA_CHECK(cudaMemcpyAsync(d->h_stats, d->stats, 2 * sizeof(unsigned long long),
                                      cudaMemcpyDeviceToHost, d->last_stream));
        QR_CUDA_CHECK(cudaStreamSynchronize(d->last_stream));
        if (frame_iterations) *frame_iterations = (int64_t)d->h_stats[0];
        if (steps) *steps = (int64_t)d->h_stats[1];
        return QR_OK;
    };
    int rc = body();
    cudaSetDevice(prev);
    return rc;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Single-node entry points (reference decoder.pyx:190-217, 301-319, 372-388): one thread, the same
// device arithmetic as the QR_F64 decoder, on caller arrays indexed by ORIGINAL edge id.
namespace qr {

struct NodeEdges {
    int32_t deg;
    int32_t edge[kMaxCheckDegree];
    int32_t var[kMaxCheckDegree];
};

__global__ void k_process_check_node(NodeEdges ne, const uint8_t *synd,
                                     int64_t check, double *c2v, const double *v2c)
{
    if (threadIdx.x || blockIdx.x) return;
    double x[kMaxCheckDegree], bwd[kMaxCheckDegree];
    for (int i = 0; i < ne.deg; ++i) x[i] = v2c[ne.edge[i]];
    MathRef::check_node<1>(ne.deg, x, bwd, synd[check] != 0);
    for (int i = 0; i < ne.deg; ++i) c2v[ne.edge[i]] = x[i];
}

__global__ void k_process_var_node(const int32_t *edges, int32_t deg, int64_t var, const double *llr,
                                   const double *c2v, double *v2c, double *post)
{
    if (threadIdx.x || blockIdx.x) return;
    double acc = llr[var];
    for (int i = 0; i < deg; ++i) acc = acc + c2v[edges[i]];
    post[var] = acc;
    for (int i = 0; i < deg; ++i) v2c[edges[i]] = acc - c2v[edges[i]];
}

__global__ void k_check_synd_node(NodeEdges ne, int64_t check, const uint8_t *word, const uint8_t *synd,
                                  uint8_t *ok)
{
    if (threadIdx.x || blockIdx.x) return;
    uint8_t parity = synd[check];
    for (int i = 0; i < ne.deg; ++i) parity ^= word[ne.var[i]];
    *ok = parity ^ 1;
}

static int node_edges(const qr_graph *g, int64_t check, NodeEdges &ne)
{
    if (check < 0 || check >= g->C) return fail(QR_ERR_INVALID, "check node index out of range");
    if (!g->decodable) return fail(QR_ERR_GRAPH, "the decoder needs every check node of degree 2..64");
    int32_t slot = -1;
    for (int64_t s = 0; s < g->C; ++s)
        if (g->chk_order[s] == check) { slot = (int32_t)s; break; }
    ne.deg = g->chk_ptr[slot + 1] - g->chk_ptr[slot];
    for (int i = 0; i < ne.deg; ++i) {
        ne.edge[i] = g->slot_edge[g->chk_ptr[slot] + i];
        ne.var[i] = g->slot_var[g->chk_ptr[slot] + i];
    }
    return QR_OK;
}

}  // namespace qr

extern "C" {

int qr_process_check_node(const qr_graph *g, int64_t check, const uint8_t *d_synd, double *d_c2v,
                          const double *d_v2c, void *stream)
{
    if (!g || g->device < 0) return qr::fail(QR_ERR_INVALID, "graph has no device");
    qr::NodeEdges ne;
    int rc = qr::node_edges(g, check, ne);
    if (rc) return rc;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(g->device);
    qr::k_process_check_node<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(ne, d_synd, check, d_c2v, d_v2c);
    cudaError_t e = cudaGetLastError();
    cudaSetDevice(prev);
    if (e != cudaSuccess) return qr::fail(QR_ERR_CUDA, cudaGetErrorString(e));
    return QR_OK;
}

int qr_process_var_node(const qr_graph *g, int64_t var, const double *d_llr, const double *d_c2v,
                        double *d_v2c, double *d_post, void *stream_)
{
    if (!g || g->device < 0) return qr::fail(QR_ERR_INVALID, "graph has no device");
    if (var < 0 || var >= g->N) return qr::fail(QR_ERR_INVALID, "variable node index out of range");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int32_t q0 = g->var_ptr[var], q1 = g->var_ptr[var + 1];
    std::vector<int32_t> edges;
    for (int32_t q = q0; q < q1; ++q) edges.push_back(g->slot_edge[g->var_slot[q]]);
    int prev = 0;
    cudaGetDevice(&prev);
    auto body = [&]() -> int {
        QR_CUDA_CHECK(cudaSetDevice(g->device));
        int32_t *d_edges = nullptr;
        QR_CUDA_CHECK(cudaMalloc((void **)&d_edges, std::max<size_t>(edges.size(), 1) * sizeof(int32_t)));
        cudaError_t e = cudaMemcpyAsync(d_edges, edges.data(), edges.size() * sizeof(int32_t),
                                        cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) {
            qr::k_process_var_node<<<1, 32, 0, stream>>>(d_edges, (int32_t)edges.size(), var, d_llr, d_c2v,
                                                         d_v2c, d_post);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        cudaFree(d_edges);
        if (e != cudaSuccess) return qr::fail(QR_ERR_CUDA, cudaGetErrorString(e));
        return QR_OK;
    };
    int rc = body();
    cudaSetDevice(prev);
    return rc;
}

int qr_check_synd_node(const qr_graph *g, int64_t check, const uint8_t *d_word, const uint8_t *d_synd,
                       uint8_t *d_ok, void *stream)
{
    if (!g || g->device < 0) return qr::fail(QR_ERR_INVALID, "graph has no device");
    qr::NodeEdges ne;
    int rc = qr::node_edges(g, check, ne);
    if (rc) return rc;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(g->device);
    qr::k_check_synd_node<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(ne, check, d_word, d_synd, d_ok);
    cudaError_t e = cudaGetLastError();
    cudaSetDevice(prev);
    if (e != cudaSuccess) return qr::fail(QR_ERR_CUDA, cudaGetErrorString(e));
    return QR_OK;
}

}  // extern "C"
