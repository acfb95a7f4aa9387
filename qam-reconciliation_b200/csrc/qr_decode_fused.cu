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
[[maybe_unused]] constexpr int kFusedRowsPerClaim = 4;   // checks per thread and claim
constexpr int kFusedMaxTile = 128;

__device__ __forceinline__ uint64_t l2_policy(int kind)
{
    uint64_t p;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// per-lane "some check unsatisfied" bits of this CTA's threads -> one global store per lane
template <int VEC>
__device__ __forceinline__ void flush_bad(int32_t *unsat_cur, int32_t lane0, int32_t tx, int32_t ty, int32_t tl,
                                          uint32_t bad, int32_t *s_flags)
{
    for (int32_t i = threadIdx.x; i < tl; i += blockDim.x) s_flags[i] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < VEC; ++k)
        if (bad >> k & 1) s_flags[tx * VEC + k] = 1;
    __syncthreads();
    if (ty == 0) {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
            if (s_flags[tx * VEC + k]) unsat_cur[lane0 + tx * VEC + k] = 1;
    }
    __syncthreads();
}

// The gathers of a tile touch its rows in random order; left to them, the first touch of every row is a DRAM
// round trip that a thread waits for with its registers tied up.  Instead, whoever takes claim `chunk` of a
// tile pulls the SAME chunk of the NEXT tile into L2 in address order (prefetch.global.L2 holds no register
// and no scoreboard): DRAM sees long sequential bursts one tile ahead of the sweep, the gathers hit L2.
template <typename T>
__device__ __forceinline__ void prefetch_next_tile(const FusedParams<T> &F, int cur, int32_t tile, int32_t chunk,
                                                   int32_t cpt, int32_t c0, int32_t c1, int32_t deg_slots0,
                                                   int32_t deg_slots1)
{
    const DecodeParams<T> &P = F.P;
    const bool wrap = tile + 1 >= F.tiles;            // the next sweep starts over on the other buffer
    const int32_t pt = wrap ? 0 : tile + 1;
    const char *pc = reinterpret_cast<const char *>(F.c2v[wrap ? (cur ^ 1) : cur] + ((int64_t)pt * P.E + deg_slots0) * F.tl);
    const int64_t cbytes = (int64_t)(deg_slots1 - deg_slots0) * F.tl * (int64_t)sizeof(T);
    for (int64_t o = (int64_t)threadIdx.x * 128; o < cbytes; o += (int64_t)kFBlock * 128)
        asm volatile("prefetch.global.L2 [%0];" :: "l"(pc + o));
    const int64_t v0 = P.N * chunk / cpt, v1 = P.N * (chunk + 1) / cpt;
    const char *pl = reinterpret_cast<const char *>(P.llr + ((int64_t)pt * P.N + v0) * F.tl);
    const int64_t lbytes = (v1 - v0) * F.tl * (int64_t)sizeof(T);
    for (int64_t o = (int64_t)threadIdx.x * 128; o < lbytes; o += (int64_t)kFBlock * 128)
        asm volatile("prefetch.global.L2 [%0];" :: "l"(pl + o));
    (void)c0; (void)c1;
}

// Work distribution: WARPS claim work, not CTAs.  A warp is 32 / bx checks wide (bx = TL / VEC threads share a
// check) and takes rows_per_claim passes per claim from one global counter, in tile order; the atomic of the
// NEXT claim is issued before the current one is processed.  No CTA barrier inside the phase, so the warps of
// an SM drift apart and cover each other's memory round trips (measured: a CTA-wide claim with its barrier cost
// 2.1 us per claim, profiles/r1_fused_experiments.txt f7/f8).  Per-lane "some check unsatisfied" flags are
// OR-reduced inside the warp and stored per tile (idempotent stores of 1).
template <typename T, int VEC, int DSEL>
__device__ __forceinline__ void fused_phase(const FusedParams<T> &F, int cur, int32_t *s_flags, int32_t *s_claim,
                                            uint64_t pol_ld, uint64_t pol_st)
{
    (void)s_flags; (void)s_claim;
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

// ---------------------------------------------------------------------------------------------
// PIPELINED fused phase (check-regular graphs of degree D): the 4 D rows of an item go straight from L2
// to a thread-private shared-memory slot with cp.async (no registers held while they fly), two items
// deep -- while item k is being computed item k+1 is in flight, and item k+2 is issued as soon as k's
// slot is free.  Per SM 2 x 256 x 4 D x 16 B (192 KB for D = 6) of loads can be outstanding, against
// the ~96 KB (and only between compute bursts) of the register-staged version.
// Shared-memory slot of (stage s, row j) for thread t: uint4 index (s * 4 D + j) * kFBlock + t --
// consecutive threads, consecutive 16 B: conflict-free LDS.128 / cp.async.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint64_t pol)
{
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct ItemMeta {
    int32_t tile;      // tile of the claim this pipeline slot belongs to (the same for every thread of the CTA); -1: none
    int32_t ci;        // internal check slot (row of the syndrome tile; first CSR slot = ci * D); -1: this thread has no item
    uint32_t own;      // 2 bits per edge: position of the own edge among the variable's three
    uint32_t sy;       // the VEC syndrome bytes of the thread's lanes
    uint32_t active, fresh;
};

template <typename T, int VEC, int D, int STAGES>
__device__ __forceinline__ void fused_phase_pipe(const FusedParams<T> &F, int cur, int32_t *s_flags,
                                                 int32_t *s_claim, uint4 *stage, uint64_t pol_ld, uint64_t pol_st)
{
    using VT = Vec<T, VEC>;
    constexpr int ROWS = 4 * D;
    const DecodeParams<T> &P = F.P;
    const int32_t tl = F.tl, bx = tl / VEC, by = kFBlock / bx;
    const int32_t tx = threadIdx.x % bx, ty = threadIdx.x / bx, lt = tx * VEC;
    const int32_t claim_rows = by * kFusedRowsPerClaim, C = (int32_t)P.C;
    const int32_t cpt = (C + claim_rows - 1) / claim_rows, total = cpt * F.tiles;
    const int64_t E = P.E, N = P.N;
    const uint32_t stage_base = (uint32_t)__cvta_generic_to_shared(stage + threadIdx.x);

    // issue side: the claim being handed out row by row
    int32_t cl_tile = -1, cl_c0 = 0, r = kFusedRowsPerClaim;
    bool exhausted = false;
    uint32_t L_active = 0, L_fresh = 0;
    // compute side
    ItemMeta meta0, meta1;   // (two named slots, selected at compile time: no local-memory array)
    meta0.tile = meta1.tile = -1;
    meta0.ci = meta1.ci = -1;
    int32_t comp_tile = -1;
    uint32_t bad = 0;

    // item i of this thread; a new claim is fetched by the whole CTA at the same i
    auto next_item = [&](ItemMeta &m, Nbr4 (&q)[D]) {
        if (!exhausted && r == kFusedRowsPerClaim) {
            __syncthreads();
            if (threadIdx.x == 0) *s_claim = atomicAdd(&P.work[0], 1);
            __syncthreads();
            const int32_t cl = *s_claim;
            if (cl >= total) {
                exhausted = true;
            } else {
                const int32_t tile = cl / cpt;
                cl_c0 = (cl - tile * cpt) * claim_rows;
                r = 0;
                if (tile != cl_tile) {
                    const LaneInfo<VEC> L = load_lane_info<T, VEC>(P, cur, (tile * tl) / VEC + tx);
                    L_active = L.active; L_fresh = L.fresh;
                    cl_tile = tile;
                }
                if (F.prefetch & 1) {
                    // The gathers of a tile touch its rows in random order; left to them, the first touch of
                    // every row is a random 128..256-byte DRAM access (row-buffer miss each time).  Instead the
                    // SAME chunk of the NEXT tile is pulled into L2 here in address order, one claim ahead of
                    // the sweep: DRAM sees long sequential bursts, the gathers hit L2.
                    const int32_t chunk = cl - tile * cpt;
                    const bool wrap = tile + 1 >= F.tiles;            // next sweep starts over on the other buffer
                    const int32_t pt = wrap ? 0 : tile + 1;
                    const T *pc = F.c2v[wrap ? (cur ^ 1) : cur] + ((int64_t)pt * E + (int64_t)cl_c0 * D) * tl;
                    const int32_t c_hi = min(cl_c0 + claim_rows, C);
                    const int64_t cbytes = (int64_t)(c_hi - cl_c0) * D * tl * (int64_t)sizeof(T);
                    for (int64_t o = (int64_t)threadIdx.x * 128; o < cbytes; o += (int64_t)kFBlock * 128)
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char *>(pc) + o));
                    const int64_t v0 = N * chunk / cpt, v1 = N * (chunk + 1) / cpt;
                    const T *pl = P.llr + ((int64_t)pt * N + v0) * tl;
                    const int64_t lbytes = (v1 - v0) * tl * (int64_t)sizeof(T);
                    for (int64_t o = (int64_t)threadIdx.x * 128; o < lbytes; o += (int64_t)kFBlock * 128)
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char *>(pl) + o));
                }
            }
        }
        m.tile = -1; m.ci = -1; m.own = 0; m.sy = 0; m.active = 0; m.fresh = 0;
        if (!exhausted) {
            const int32_t row = cl_c0 + ty + r * by;
            ++r;
            m.tile = cl_tile;
            if (row < C && L_active) {
                m.ci = row; m.active = L_active; m.fresh = L_fresh;
                load_nbr_row<D>(F.nbr, row * D, q);
                const uint8_t *sp = P.synd + ((int64_t)cl_tile * C + row) * tl + lt;
                if constexpr (VEC == 4) m.sy = *reinterpret_cast<const uint32_t *>(sp);
                else m.sy = *reinterpret_cast<const uint16_t *>(sp);
#pragma unroll
                for (int e = 0; e < D; ++e) m.own |= ((uint32_t)q[e].vp >> 28) << (2 * e);
            }
        }
    };
    // the rows of item `c` have landed in stage s
    auto compute = [&](const ItemMeta &c, int s) {
        if (c.tile != comp_tile) {      // CTA-uniform: publish the flags of the tile just finished
            if (comp_tile >= 0) flush_bad<VEC>(P.unsat[cur], comp_tile * tl, tx, ty, tl, bad, s_flags);
            bad = 0;
            comp_tile = c.tile;
        }
        if (c.ci < 0) return;
        T *c_new = F.c2v[cur ^ 1] + (int64_t)c.tile * E * tl + lt;
        uint32_t par = 0;
#pragma unroll
        for (int k = 0; k < VEC; ++k) par |= ((c.sy >> (8 * k)) & 1u) << k;
        VT x[D];
#pragma unroll
        for (int e = 0; e < D; ++e) {
            VT ch, m0, m1, m2;
            const uint4 *sp = stage + ((size_t)(s * ROWS + 4 * e) * kFBlock + threadIdx.x);
            const uint4 t0 = sp[0], t1 = sp[kFBlock], t2 = sp[2 * kFBlock], t3 = sp[3 * kFBlock];
            memcpy(&ch, &t0, 16); memcpy(&m0, &t1, 16); memcpy(&m1, &t2, 16); memcpy(&m2, &t3, 16);
            const int own = (int)((c.own >> (2 * e)) & 3u);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const bool fresh = (c.fresh >> k & 1) != 0;          // first half-iteration: c2v == 0 (decoder.pyx:408)
                const T c0 = fresh ? (T)0 : m0.v[k];
                const T c1 = fresh ? (T)0 : m1.v[k];
                const T c2 = fresh ? (T)0 : m2.v[k];
                T post = ch.v[k] + c0;                               // decoder.pyx:291-293, ascending edge id
                post = post + c1;
                post = post + c2;
                par ^= (uint32_t)(post < (T)0) << k;                 // decoder.pyx:244 (strict <)
                const T mine = own == 0 ? c0 : (own == 1 ? c1 : c2);
                x[e].v[k] = post - mine;                             // decoder.pyx:295-297
            }
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            T xs[D];
#pragma unroll
            for (int e = 0; e < D; ++e) xs[e] = x[e].v[k];
            MathOf<T>::template run<D, D>(D, xs, ((c.sy >> (8 * k)) & 1u) != 0);
#pragma unroll
            for (int e = 0; e < D; ++e) x[e].v[k] = xs[e];
        }
#pragma unroll
        for (int e = 0; e < D; ++e)
            st_pol(reinterpret_cast<VT *>(c_new + ((int64_t)c.ci * D + e) * tl), x[e], pol_st);
        bad |= par & c.active;
    };
    auto issue = [&](const ItemMeta &m, const Nbr4 (&q)[D], int s) {
        if (m.ci >= 0) {
            const T *c_old = F.c2v[cur] + (int64_t)m.tile * E * tl + lt;
            const T *llr_t = P.llr + (int64_t)m.tile * N * tl + lt;
#pragma unroll
            for (int e = 0; e < D; ++e) {
                const uint32_t dst = stage_base + (uint32_t)((s * ROWS + 4 * e) * kFBlock * 16);
                cp_async16(dst, llr_t + (int64_t)(q[e].vp & 0x0fffffff) * tl, pol_ld);
                cp_async16(dst + kFBlock * 16, c_old + (int64_t)q[e].n0 * tl, pol_ld);
                cp_async16(dst + 2 * kFBlock * 16, c_old + (int64_t)q[e].n1 * tl, pol_ld);
                cp_async16(dst + 3 * kFBlock * 16, c_old + (int64_t)q[e].n2 * tl, pol_ld);
            }
        }
        cp_async_commit();
    };

    if constexpr (STAGES == 1) {
        // one item per thread in flight; latency is covered by the other warps of the SM (2 CTAs resident)
        for (;;) {
            ItemMeta m;
            Nbr4 q[D];
            next_item(m, q);
            if (exhausted) break;
            issue(m, q, 0);
            cp_async_wait0();
            compute(m, 0);
        }
    } else {
        // two items deep: compute i - 2, then refill its stage with item i
        auto step = [&](auto S) -> bool {
            constexpr int s = decltype(S)::value;
            ItemMeta &slot = s ? meta1 : meta0;
            ItemMeta m;
            Nbr4 q[D];
            next_item(m, q);
            const ItemMeta c = slot;
            cp_async_wait1();
            compute(c, s);
            issue(m, q, s);
            slot = m;
            return exhausted && meta0.tile < 0 && meta1.tile < 0;
        };
        for (;;) {
            if (step(std::integral_constant<int, 0>{})) break;
            if (step(std::integral_constant<int, 1>{})) break;
        }
        cp_async_wait0();
    }
    if (comp_tile >= 0) flush_bad<VEC>(P.unsat[cur], comp_tile * tl, tx, ty, tl, bad, s_flags);
}

// ---------------------------------------------------------------------------------------------
// BULK-COPY fused phase (PIPE == 3): the bx threads that share a check (one lane-tile row each) stage the
// 4 D rows of their item with TMA bulk copies (cp.async.bulk, one whole TL * w byte row per copy, completion
// on the group's mbarrier) instead of per-thread 16-byte cp.async -- measured on B200, per-thread 16-byte
// cp.async.cg requests are NOT merged into shared sectors (every thread pulls its own 32-byte sector: twice
// the L2 -> SM traffic, profiles/r1_fused_*), whole-row bulk copies move exactly the row.
// Stage of group g: smem[g][4 D rows][TL * w bytes]; thread tx reads bytes [16 tx, 16 tx + 16) of each row.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase)
{
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_row(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}

template <typename T, int VEC, int D>
__device__ __forceinline__ void fused_phase_bulk(const FusedParams<T> &F, int cur, int32_t *s_flags,
                                                 int32_t *s_claim, unsigned char *stage, uint64_t *bars,
                                                 uint32_t &phase, uint64_t pol_ld, uint64_t pol_st)
{
    using VT = Vec<T, VEC>;
    constexpr int ROWS = 4 * D;
    const DecodeParams<T> &P = F.P;
    const int32_t tl = F.tl, bx = tl / VEC, by = kFBlock / bx;
    const int32_t tx = threadIdx.x % bx, ty = threadIdx.x / bx, lt = tx * VEC;
    const int32_t claim_rows = by * kFusedRowsPerClaim, C = (int32_t)P.C;
    const int32_t cpt = (C + claim_rows - 1) / claim_rows, total = cpt * F.tiles;
    const int64_t E = P.E, N = P.N;
    const uint32_t row_bytes = (uint32_t)(tl * sizeof(T));
    unsigned char *my_stage = stage + (size_t)ty * ROWS * row_bytes;
    const uint32_t stage_addr = (uint32_t)__cvta_generic_to_shared(my_stage);
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bars + ty);
    // lanes of the warp that belong to this thread's group (bx <= 32 consecutive lanes)
    const uint32_t lane_id = threadIdx.x & 31u;
    const uint32_t grp_mask = (bx >= 32 ? 0xffffffffu : ((1u << bx) - 1u) << (lane_id / bx * bx));

    int32_t cl_tile = -1, cl_c0 = 0, r = kFusedRowsPerClaim;
    bool exhausted = false;
    uint32_t L_active = 0, L_fresh = 0;
    int32_t comp_tile = -1;
    uint32_t bad = 0;

    for (;;) {
        if (r == kFusedRowsPerClaim) {
            __syncthreads();
            if (threadIdx.x == 0) *s_claim = atomicAdd(&P.work[0], 1);
            __syncthreads();
            const int32_t cl = *s_claim;
            if (cl >= total) {
                exhausted = true;
            } else {
                const int32_t tile = cl / cpt;
                cl_c0 = (cl - tile * cpt) * claim_rows;
                r = 0;
                if (tile != cl_tile) {
                    const LaneInfo<VEC> L = load_lane_info<T, VEC>(P, cur, (tile * tl) / VEC + tx);
                    L_active = L.active; L_fresh = L.fresh;
                    cl_tile = tile;
                }
            }
        }
        if (exhausted) break;
        const int32_t row = cl_c0 + ty + r * by;
        ++r;
        if (cl_tile != comp_tile) {      // CTA-uniform: publish the flags of the tile just finished
            if (comp_tile >= 0) flush_bad<VEC>(P.unsat[cur], comp_tile * tl, tx, ty, tl, bad, s_flags);
            bad = 0;
            comp_tile = cl_tile;
        }
        // the group has an item if its check exists and any of its lanes runs a frame
        const bool grp_valid = row < C && (__ballot_sync(0xffffffffu, L_active != 0) & grp_mask) != 0;
        Nbr4 q[D];
        uint32_t own = 0, sy = 0;
        if (grp_valid) {
            load_nbr_row<D>(F.nbr, row * D, q);
            const uint8_t *sp = P.synd + ((int64_t)cl_tile * C + row) * tl + lt;
            if constexpr (VEC == 4) sy = *reinterpret_cast<const uint32_t *>(sp);
            else sy = *reinterpret_cast<const uint16_t *>(sp);
        }
        __syncwarp();                   // everyone is done reading the stage of the previous item
        if (grp_valid && tx == 0) mbar_expect_tx(bar, ROWS * row_bytes);
        __syncwarp();
        if (grp_valid) {
            const char *c_old = reinterpret_cast<const char *>(F.c2v[cur] + (int64_t)cl_tile * E * tl);
            const char *llr_t = reinterpret_cast<const char *>(P.llr + (int64_t)cl_tile * N * tl);
#pragma unroll
            for (int e = 0; e < D; ++e) {
                own |= ((uint32_t)q[e].vp >> 28) << (2 * e);
                // row j = 4 e + {0: llr, 1..3: the variable's three messages}; thread tx issues rows j % bx == tx
                const int32_t idx[4] = {q[e].vp & 0x0fffffff, q[e].n0, q[e].n1, q[e].n2};
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = 4 * e + jj;
                    if ((j & (bx - 1)) == tx)
                        bulk_row(stage_addr + j * row_bytes, (jj == 0 ? llr_t : c_old) + (int64_t)idx[jj] * row_bytes,
                                 row_bytes, bar, pol_ld);
                }
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
        }
        if (grp_valid && L_active) {
            T *c_new = F.c2v[cur ^ 1] + (int64_t)cl_tile * E * tl + lt;
            uint32_t par = 0;
#pragma unroll
            for (int k = 0; k < VEC; ++k) par |= ((sy >> (8 * k)) & 1u) << k;
            VT x[D];
#pragma unroll
            for (int e = 0; e < D; ++e) {
                VT ch, m0, m1, m2;
                const unsigned char *sp = my_stage + (size_t)(4 * e) * row_bytes + tx * 16;
                const uint4 t0 = *reinterpret_cast<const uint4 *>(sp);
                const uint4 t1 = *reinterpret_cast<const uint4 *>(sp + row_bytes);
                const uint4 t2 = *reinterpret_cast<const uint4 *>(sp + 2 * row_bytes);
                const uint4 t3 = *reinterpret_cast<const uint4 *>(sp + 3 * row_bytes);
                memcpy(&ch, &t0, 16); memcpy(&m0, &t1, 16); memcpy(&m1, &t2, 16); memcpy(&m2, &t3, 16);
                const int ow = (int)((own >> (2 * e)) & 3u);
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const bool fresh = (L_fresh >> k & 1) != 0;          // first half-iteration: c2v == 0 (decoder.pyx:408)
                    const T c0 = fresh ? (T)0 : m0.v[k];
                    const T c1 = fresh ? (T)0 : m1.v[k];
                    const T c2 = fresh ? (T)0 : m2.v[k];
                    T post = ch.v[k] + c0;                               // decoder.pyx:291-293, ascending edge id
                    post = post + c1;
                    post = post + c2;
                    par ^= (uint32_t)(post < (T)0) << k;                 // decoder.pyx:244 (strict <)
                    const T mine = ow == 0 ? c0 : (ow == 1 ? c1 : c2);
                    x[e].v[k] = post - mine;                             // decoder.pyx:295-297
                }
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                T xs[D];
#pragma unroll
                for (int e = 0; e < D; ++e) xs[e] = x[e].v[k];
                MathOf<T>::template run<D, D>(D, xs, ((sy >> (8 * k)) & 1u) != 0);
#pragma unroll
                for (int e = 0; e < D; ++e) x[e].v[k] = xs[e];
            }
#pragma unroll
            for (int e = 0; e < D; ++e)
                st_pol(reinterpret_cast<VT *>(c_new + ((int64_t)row * D + e) * tl), x[e], pol_st);
            bad |= par & L_active;
        }
    }
    if (comp_tile >= 0) flush_bad<VEC>(P.unsat[cur], comp_tile * tl, tx, ty, tl, bad, s_flags);
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

template <typename T, int VEC, int DSEL, int PIPE>   // PIPE: 0 register-staged, 1 / 2 = cp.async stages per thread
__global__ void __launch_bounds__(kFBlock, (PIPE == 1 || (PIPE == 3 && sizeof(T) == 4) || (PIPE == 0 && DSEL > 0)) ? ((sizeof(T) == 4 && VEC == 2) ? 3 : 2) : 1) k_fused(FusedParams<T> F)
{
    extern __shared__ uint4 dyn_stage[];   // PIPE stages x 4 DSEL rows x kFBlock threads x 16 B
    __shared__ uint64_t s_bars[kFBlock / 2];   // PIPE == 3: one mbarrier per group of bx threads
    uint32_t bulk_phase = 0;
    __shared__ int32_t s_flags[kFusedMaxTile];
    __shared__ int32_t s_claim[2], s_last;
    const DecodeParams<T> &P = F.P;
    cg::grid_group grid = cg::this_grid();
    const uint64_t pol_ld = l2_policy(F.hints >= 2 ? 2 : 0), pol_st = l2_policy(F.hints >= 1 ? 1 : 0);
    if constexpr (PIPE == 3) {
        if (threadIdx.x < kFBlock / 2) mbar_init((uint32_t)__cvta_generic_to_shared(s_bars + threadIdx.x), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    fused_refill_phase<T>(F, 0, 0);      // first generation of frames into the lanes
    grid.sync();
    for (int step = 0;; ++step) {
        const int cur = step & 1;
        if constexpr (PIPE == 3)
            fused_phase_bulk<T, VEC, DSEL>(F, cur, s_flags, s_claim, reinterpret_cast<unsigned char *>(dyn_stage), s_bars,
                                           bulk_phase, pol_ld, pol_st);
        else if constexpr (PIPE > 0) fused_phase_pipe<T, VEC, DSEL, PIPE>(F, cur, s_flags, s_claim, dyn_stage, pol_ld, pol_st);
        else fused_phase<T, VEC, DSEL>(F, cur, s_flags, s_claim, pol_ld, pol_st);
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

template <typename T, int VEC, int DSEL, int PIPE>
static int launch_fused(qr_decoder *d, const FusedParams<T> &F, cudaStream_t stream)
{
    const size_t smem = (size_t)(PIPE == 3 ? 1 : PIPE) * 4 * DSEL * kFBlock * 16;
    auto kern = k_fused<T, VEC, DSEL, PIPE>;
    if (smem) QR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
#ifdef QR_FUSED_EXPERIMENTS
    // staging variants measured on B200 and found slower than the register-staged phase:
    // 1 / 2 = per-thread cp.async stages, 3 = TMA bulk copies of whole rows
    if (d->regular_degree == 6 && d->fused_pipe == 3) return launch_fused<T, VEC, 6, 3>(d, F, stream);
    if (d->regular_degree == 6 && d->fused_pipe == 2) return launch_fused<T, VEC, 6, 2>(d, F, stream);
    if (d->regular_degree == 6 && d->fused_pipe == 1 && sizeof(T) == 4) return launch_fused<T, VEC, 6, 1>(d, F, stream);
#endif
    if (d->regular_degree == 6) return launch_fused<T, VEC, 6, 0>(d, F, stream);
    return launch_fused<T, VEC, 0, 0>(d, F, stream);
}

template <typename T>
int run_batch_fused(qr_decoder *d, const DecodeParams<T> &P, cudaStream_t stream)
{
#ifdef QR_FUSED_EXPERIMENTS
    // 2 lanes per thread (8-byte accesses, 80 registers, 3 CTAs/SM = 24 warps): measured 69.6 ms against 61.7 ms
    if constexpr (sizeof(T) == 4) {
        if (d->vec == 2 && d->regular_degree == 6) return run_batch_fused_v<T, 2>(d, P, stream);   // QAMRECON_VEC=2
    }
#endif
    return run_batch_fused_v<T, 16 / sizeof(T)>(d, P, stream);
}

template int run_batch_fused<float>(qr_decoder *, const DecodeParams<float> &, cudaStream_t);
template int run_batch_fused<double>(qr_decoder *, const DecodeParams<double> &, cudaStream_t);

}  // namespace qr
