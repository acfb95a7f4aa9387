// QR_SCHED_FUSED: persistent kernel running the fused flooding iteration of qr_decode_fused.cuh as a TILE
// PIPELINE (reference: Decoder._decode, decoder.pyx:391-436; the stream of work items, BK and PP are described
// at the top of qr_decode_fused.cuh).
//
// Thread mapping of a sweep claim: bx = TL / VEC threads share a check (together they move one contiguous
// TL * w byte row per access), a warp is 32 / bx checks wide and takes rows_per_claim passes per claim.  WARPS
// claim work, not CTAs: no CTA barrier anywhere after the initial fill, so the warps of an SM drift apart and cover
// each other's memory round trips.  fp32: 128 registers, two CTAs of 256 threads per SM.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <thread>
#include <type_traits>
#include <vector>

// Row loads of a fused item are issued in batches of FUSED_BATCH_EDGES edges (4 loads each).  Measured on B200
// (2048 frames x 50 iterations, 1024 lanes, 128 registers, 2 CTAs/SM): batches of 3 edges 63.7 ms, 4 edges 59.8 ms,
// 5 edges 67.3 ms, all 6 at once 64.5 ms (spills); 4 it is.
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

__device__ __forceinline__ int32_t ld_volatile(const int32_t *p) { return *reinterpret_cast<const volatile int32_t *>(p); }

// ---------------------------------------------------------------------------------------------
// BK(tile, round): advance the lane state machine of one tile (decoder.pyx:431-436), one warp, lane i of the warp
// = lane i (+32, +64 ...) of the tile.  Finished frames get their flags, the lane its next frame; the tile's
// refill list is written for PP.
template <typename T>
__device__ __forceinline__ void tile_bookkeep(const FusedParams<T> &F, int32_t tile, int32_t round)
{
    const DecodeParams<T> &P = F.P;
    const int32_t lid = threadIdx.x & 31;
    const int32_t minfin_used = ld_volatile(&F.tile_minfin[tile]);      // what the sweep just finished went by
    int32_t listed = 0;
    int32_t next_iter[kFusedMaxTile / 32], next_fresh[kFusedMaxTile / 32];   // this thread's lanes after the update
    RefillEntry *list = F.rlist + (int64_t)tile * F.tl;
    for (int32_t base = 0; base < F.tl; base += 32) {
        const int32_t lane = tile * F.tl + base + lid;
        LaneState s = ld_stream(&P.st[0][lane]);
        const int32_t unsat = ld_stream(&P.unsat[0][lane]);
        const BkDecision d = bk_decide(s, unsat, P.maxiter);
        const bool fin = d.fin_ok || d.fin_fail;
        const uint32_t m_fin = __ballot_sync(0xffffffffu, fin);
        const int32_t n_fin = __popc(m_fin);
        unsigned long long it_sum = fin ? (unsigned long long)s.iter : 0ULL;
        for (int o = 16; o > 0; o >>= 1) it_sum += __shfl_xor_sync(0xffffffffu, it_sum, o);
        if (fin) {
            // decoder.pyx:431-436: the frame is done; its lane is parked until the sector group is released
            P.success[s.frame] = d.fin_ok ? 1 : 0;
            P.iters[s.frame] = d.fin_ok ? s.iter : P.maxiter;
            if (d.fin_ok && s.iter > 0) atomicMin(&P.ctrl[CTRL_MINFIN], s.iter);   // (0 iterations = input already consistent: no signal)
            const int32_t pv = (F.post && stores_post(s.iter, P.maxiter, minfin_used)) ? 1 : 0;
            // (a frame whose posterior was NOT stored is rebuilt from this sweep's messages: it cannot wait)
            s.retire = s.frame; s.frame = -1; s.iter = pv; s.fresh = pv ? 0 : F.park_rounds;
        } else if (s.frame >= 0) {
            s.iter += 1;
            s.fresh = 0;
            s.retire = -1;
        } else if (s.retire >= 0) {
            s.fresh += 1;                                           // parked: one more round waited
        }
        if (lid == 0 && n_fin) {
            atomicAdd(&P.stats[0], it_sum);
            atomicAdd(&P.ctrl[CTRL_REMAINING], -n_fin);
        }
        // which parked lanes are released now
        const bool parked = s.frame < 0 && s.retire >= 0;
        const uint32_t m_run = __ballot_sync(0xffffffffu, s.frame >= 0), m_park = __ballot_sync(0xffffffffu, parked);
        const uint32_t m_aged = __ballot_sync(0xffffffffu, parked && s.fresh >= F.park_rounds);
        const bool no_more = ld_volatile(&P.ctrl[CTRL_NEXT_FRAME]) >= (int32_t)min(P.frames, (int64_t)0x7fffffff);
        const uint32_t m_rel = octets_to_release(m_run, m_park, m_aged, no_more);
        const int32_t n_rel = __popc(m_rel), rank = __popc(m_rel & ((1u << lid) - 1u));
        int32_t first_frame = 0;
        if (n_rel) {
            if (lid == 0) first_frame = atomicAdd(&P.ctrl[CTRL_NEXT_FRAME], n_rel);
            first_frame = __shfl_sync(0xffffffffu, first_frame, 0);
        }
        if (m_rel >> lid & 1) {
            RefillEntry e;
            e.lane = base + lid;
            e.retire = s.retire;
            const int32_t nf = first_frame + rank;
            e.frame = (int64_t)nf < P.frames ? nf : -1;
            e.post_valid = s.iter;
            list[listed + rank] = e;
            s.frame = e.frame; s.iter = 0; s.fresh = e.frame >= 0 ? 1 : 0; s.retire = -1;
        }
        P.st[0][lane] = s;
        P.unsat[0][lane] = 0;
        next_iter[base / 32] = s.frame >= 0 ? s.iter : -1;
        next_fresh[base / 32] = s.fresh;
        listed += n_rel;
    }
    __syncwarp();
    // PP of this tile: R items into the ready queue (nothing when no frame finished: the common case below the
    // waterfall).  Ticket h lives in ring slot h mod size; the sequence number h + 1 in the upper half marks it
    // published.  At most one PP per tile is outstanding (the tile's next sweep waits for it), so the ring, sized
    // tiles * R, cannot be lapped.
    int32_t first_ticket = 0, new_minfin = 0, expect = 0;
    if (lid == 0) {
        if (listed) first_ticket = atomicAdd(&P.ctrl[CTRL_PP_RESERVE], F.pp_items);
        new_minfin = ld_volatile(&P.ctrl[CTRL_MINFIN]);
        expect = ld_volatile(&F.pp_expect[tile]) + (listed ? F.pp_items : 0);
        F.rcount[tile] = listed;
        F.tile_minfin[tile] = new_minfin;                             // fixed for the tile's next sweep
        F.pp_expect[tile] = expect;
        if (tile == 0) atomicAdd(&P.stats[1], 1ULL);                  // rounds executed
    }
    first_ticket = __shfl_sync(0xffffffffu, first_ticket, 0);
    new_minfin = __shfl_sync(0xffffffffu, new_minfin, 0);
    expect = __shfl_sync(0xffffffffu, expect, 0);
    // what the next sweep reads per lane: one byte, bit 0 runs a frame, bit 1 first half-iteration (c2v counts as
    // zero), bit 2 store the posterior (the lane may finish in that sweep)
    for (int32_t base = 0; base < F.tl; base += 32) {
        const int32_t it = next_iter[base / 32];
        uint8_t fl = 0;
        if (it >= 0) fl = (uint8_t)(1 | (next_fresh[base / 32] ? 2 : 0) | ((F.post && stores_post(it, P.maxiter, new_minfin)) ? 4 : 0));
        F.lane_flags[tile * F.tl + base + lid] = fl;
    }
    __threadfence();                                                  // release: states, flags, lists, counters
    __syncwarp();
    if (listed) {
        for (int32_t it = lid; it < F.pp_items; it += 32) {
            const uint32_t h = (uint32_t)(first_ticket + it);
            const unsigned long long ent = ((unsigned long long)(h + 1u) << 32) |
                                           (unsigned long long)(((uint32_t)tile << 8) | ((uint32_t)it << 1) | (uint32_t)(round & 1));
            *reinterpret_cast<volatile unsigned long long *>(&F.ppq[h % (uint32_t)F.ppq_size]) = ent;
        }
    }
    // rounds book-kept (low word) and PP items published so far (high word) in ONE word: a sweep claim reads both
    // with a single load and compares the second with pp_done
    if (lid == 0)
        *reinterpret_cast<volatile unsigned long long *>(&F.bk_word[tile]) =
            ((unsigned long long)(uint32_t)expect << 32) | (unsigned long long)(uint32_t)(round + 1);
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// PP item `item` of (tile, round): rows [N item / R, N (item + 1) / R) of the variables and the same share of the
// checks, for every lane of the tile's refill list.  One warp; lane i takes rows r0 + i, r0 + i + 32, ...; all loads
// of a batch of kPPRows rows are issued before their stores.  `cur` = message buffer the sweep read.
constexpr int kPPRows = 4;    // rows in flight per thread (measured on B200: 16 is SLOWER -- the column accesses of a refill
                              // are 32-byte sectors at DRAM, and more of them in flight only crowd out the sweeps)

template <typename T>
__device__ __forceinline__ int32_t tile_pp_item(const FusedParams<T> &F, int32_t tile, int32_t item, int cur)
{
    const DecodeParams<T> &P = F.P;
    const int32_t cnt = ld_stream(&F.rcount[tile]);     // (stable until this PP is complete: the next BK of the tile comes after it)
    if (cnt == 0) return 0;
    const int32_t lid = threadIdx.x & 31, tl = F.tl, R = F.pp_items;
    const int32_t v0 = (int32_t)(P.N * item / R), v1 = (int32_t)(P.N * (item + 1) / R);
    const int32_t c0 = (int32_t)(P.C * item / R), c1 = (int32_t)(P.C * (item + 1) / R);
    const RefillEntry *list = F.rlist + (int64_t)tile * tl;
    for (int32_t ei = 0; ei < cnt; ++ei) {
        const RefillEntry e = ld_stream(&list[ei]);
        T *llr_col = P.llr + (int64_t)tile * P.N * tl + e.lane;
        // ---- ship the retired frame's posteriors
        if (e.retire >= 0 && P.post_out) {
            const bool never_iterated = ld_stream(&P.iters[e.retire]) == 0;
            if (never_iterated || !e.post_valid) {
                // rare paths, element by element: the copy / +0.0 semantics of a frame that never iterated, and the
                // rebuild llr + sum c2v of a frame that converged earlier than anything seen before
                RefillEntry only = e;
                only.frame = -1;
                for (int32_t n = v0 + lid; n < v1; n += 32) fused_pp_var_elem<T>(F, cur, only, tile, n);
            } else {
                const T *pc = F.post + (int64_t)tile * P.N * tl + e.lane;
                for (int32_t n0 = v0 + lid; n0 < v1; n0 += 32 * kPPRows) {
                    T v[kPPRows];
#pragma unroll
                    for (int i = 0; i < kPPRows; ++i) {
                        const int32_t n = n0 + 32 * i;
                        if (n < v1) v[i] = ld_stream(&pc[(int64_t)n * tl]);
                    }
#pragma unroll
                    for (int i = 0; i < kPPRows; ++i) {
                        const int32_t n = n0 + 32 * i;
                        if (n < v1) store_output_llr(P.post_out, P.post_out_f64, (int64_t)e.retire * P.N + n, (double)v[i]);
                    }
                }
            }
        }
        // ---- admit the next frame: channel LLRs and syndrome into the lane's columns
        if (e.frame >= 0) {
            for (int32_t n0 = v0 + lid; n0 < v1; n0 += 32 * kPPRows) {
                T v[kPPRows];
#pragma unroll
                for (int i = 0; i < kPPRows; ++i) {
                    const int32_t n = n0 + 32 * i;
                    if (n < v1) v[i] = load_input_llr<T>(P.llr_in, P.llr_in_f64, (int64_t)e.frame * P.N + n);
                }
#pragma unroll
                for (int i = 0; i < kPPRows; ++i) {
                    const int32_t n = n0 + 32 * i;
                    if (n < v1) llr_col[(int64_t)n * tl] = v[i];
                }
            }
            uint8_t *synd_col = P.synd + (int64_t)tile * P.C * tl + e.lane;
            for (int32_t k0 = c0 + lid; k0 < c1; k0 += 32 * kPPRows) {
                uint8_t sy[kPPRows];
#pragma unroll
                for (int i = 0; i < kPPRows; ++i) {
                    const int32_t ci = k0 + 32 * i;
                    if (ci < c1) sy[i] = P.synd_in[(int64_t)e.frame * P.C + P.chk_order[ci]];
                }
#pragma unroll
                for (int i = 0; i < kPPRows; ++i) {
                    const int32_t ci = k0 + 32 * i;
                    if (ci < c1) synd_col[(int64_t)ci * tl] = sy[i];
                }
            }
        }
    }
    return cnt;
}

// ---------------------------------------------------------------------------------------------
// INITIAL FILL: first generation of frames into the lanes (lane l takes frame l), whole grid, before the stream
// starts.  Work item = (lane, block of kFBlock * kFillRows rows), row-block-major so that CTAs resident at the same
// time write the same rows of neighbouring lanes and share the sectors of the lane-interleaved rows through L2.
constexpr int kFillRows = 4;

template <typename T>
__device__ __forceinline__ void fused_initial_fill(const FusedParams<T> &F)
{
    const DecodeParams<T> &P = F.P;
    const int32_t cnt = (int32_t)min((int64_t)P.lanes, P.frames);
    const int32_t span = kFBlock * kFillRows, tl = F.tl;
    const int32_t vitems = (int32_t)((P.N + span - 1) / span), citems = (int32_t)((P.C + span - 1) / span);
    const int64_t total = (int64_t)cnt * (vitems + citems);
    for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
        const int32_t lane = (int32_t)(item % cnt), r = (int32_t)(item / cnt);
        const int32_t tile = lane / tl, lt = lane % tl, frame = lane;
        if (r >= vitems) {
            const int32_t c0 = (r - vitems) * span + threadIdx.x;
            uint8_t sy[kFillRows];
#pragma unroll
            for (int i = 0; i < kFillRows; ++i) {
                const int32_t ci = c0 + i * kFBlock;
                if (ci < P.C) sy[i] = P.synd_in[(int64_t)frame * P.C + P.chk_order[ci]];
            }
#pragma unroll
            for (int i = 0; i < kFillRows; ++i) {
                const int32_t ci = c0 + i * kFBlock;
                if (ci < P.C) P.synd[((int64_t)tile * P.C + ci) * tl + lt] = sy[i];
            }
            continue;
        }
        const int32_t n0 = r * span + threadIdx.x;
        T *llr_col = P.llr + (int64_t)tile * P.N * tl + lt;
        T v[kFillRows];
#pragma unroll
        for (int i = 0; i < kFillRows; ++i) {
            const int32_t n = n0 + i * kFBlock;
            if (n < P.N) v[i] = load_input_llr<T>(P.llr_in, P.llr_in_f64, (int64_t)frame * P.N + n);
        }
#pragma unroll
        for (int i = 0; i < kFillRows; ++i) {
            const int32_t n = n0 + i * kFBlock;
            if (n < P.N) llr_col[(int64_t)n * tl] = v[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One claim of F(tile, round): checks [c0, c1) of the tile for the warp.
template <typename T, int VEC, int DSEL, int VDEG>
__device__ __forceinline__ uint32_t fused_claim(const FusedParams<T> &F, int32_t tile, int32_t round, uint32_t flags,
                                                int32_t c0, int32_t c1, int32_t lid)
{
    // Everything the rows need is derived HERE from (tile, round, flags) and the constant bank: nothing of it is
    // live across the passes of a claim, where the item wants every register (spilled control state costs L2
    // bandwidth -- local memory is write-through -- which is what bounds this kernel).
    if constexpr (std::is_same<T, float>::value && VEC == 4 && DSEL == 6 && VDEG == 3) {
        // (this instantiation is only launched with lean records: the generic item is not compiled into it)
        // check-regular graph: internal check ci has CSR slots [6 ci, 6 ci + 6)
        int32_t tl = F.tl;
        asm volatile("" : "+r"(tl));          // opaque: no tile pointer arithmetic is hoisted out of the pass loop and spilled
        const int32_t bx = tl >> 2, tx = lid % bx, tyw = lid / bx;
        uint32_t active = 0, fresh = 0, wpost = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t f = flags >> (8 * k);
            active |= (f & 1u) << k;
            fresh |= ((f >> 1) & 1u) << k;
            wpost |= ((f >> 2) & 1u) << k;
        }
        const int32_t ci = c0 + tyw;          // (a pass is one row per thread: c1 - c0 <= 32 / bx)
        if (!active || ci >= c1) return 0;
        const int cur = round & 1;
        const NbrL *rec = static_cast<const NbrL *>(F.nbr_lean);
        const int64_t tile_b = (int64_t)tile * tl * 4;
        const char *llr_t = reinterpret_cast<const char *>(F.P.llr) + tile_b * F.P.N + tx * 16;
        const char *cold_t = reinterpret_cast<const char *>(F.c2v[cur]) + tile_b * F.P.E + tx * 16;
        char *cnew_t = reinterpret_cast<char *>(F.c2v[cur ^ 1]) + tile_b * F.P.E + tx * 16;
        char *post_t = F.post ? reinterpret_cast<char *>(F.post) + tile_b * F.P.N + tx * 16 : nullptr;
        const uint8_t *synd_t = F.P.synd + (int64_t)tile * tl * F.P.C + tx * 4;
        if (fresh) return fused_item_lean<6, true>(llr_t, cold_t, cnew_t, post_t, synd_t, rec, ci, 6 * ci, tl, fresh, wpost, active, F.hints >= 1);
        return fused_item_lean<6, false>(llr_t, cold_t, cnew_t, post_t, synd_t, rec, ci, 6 * ci, tl, 0u, wpost, active, F.hints >= 1);
    } else {
    const int32_t bx = F.tl / VEC, tx = lid % bx, tyw = lid / bx, wy = 32 / bx;
    LaneInfo<VEC> L;
    L.l0 = tile * F.tl + tx * VEC;
    L.active = L.fresh = L.wpost = L.fin_ok = L.fin_fail = L.upd = 0;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const uint32_t f = flags >> (8 * k);
        L.active |= (f & 1u) << k;
        L.fresh |= ((f >> 1) & 1u) << k;
        L.wpost |= ((f >> 2) & 1u) << k;
    }
    if (!L.active) return 0;
    const TileView<T> V = tile_view(F, round & 1, tile);
    const uint64_t pol_ld = l2_policy(F.hints == 2 ? 2 : 0), pol_st = l2_policy(F.hints >= 1 ? 1 : 0);
    const DecodeParams<T> &P = F.P;
    uint32_t bad = 0;
    for (int32_t b = 0; b < P.n_bins; ++b) {
        const CheckBin bin = P.bins[b];
        const int32_t lo = max(c0, bin.chk_begin), hi = min(c1, bin.chk_begin + bin.count);
        if (lo >= hi) continue;
        const CheckBin sub{bin.degree, lo, hi - lo, bin.slot_begin + (lo - bin.chk_begin) * bin.degree};
        if constexpr (DSEL > 0) bad |= run_fused_bin<T, VEC, DSEL, VDEG>(V, F.nbr, L, tx * VEC, sub, tyw, wy, pol_ld, pol_st);
        else bad |= run_fused_bin_any<T, VEC>(V, F.nbr, L, tx * VEC, sub, tyw, wy, pol_ld, pol_st);
    }
    return bad;
    }
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Per-warp control block in shared memory.  The sweep item needs every register it can get (16 row loads in
// flight + 24 message values + the check-node recursion at 128 registers per thread); whatever is only touched
// BETWEEN items lives here, not in registers.  Lane 0 writes, __syncwarp(), every lane reads (broadcast).
struct WarpCtl {
    unsigned long long slot, bk;      // prefetched for the NEXT claim: ticket slot, the tile's {PP published | rounds book-kept}
    uint32_t ticket;
    int32_t done, ppd;                // prefetched: CTRL_COMPLETED, the next tile's PP items done
    int32_t round, tile, chunk;       // current claim
    int32_t n_round, n_tile, n_chunk; // next claim
    int32_t sig_tile, sig_round;      // finished claim not yet counted
    int32_t cnt_tile, cnt_round, cnt_val;   // count issued, result arrives during the next claim
    int32_t n_have_flags;
    uint32_t cpt;                     // claims per tile sweep
};

// (cold paths out of line; F is a __grid_constant__ kernel parameter, so the reference is a pointer into the
// constant bank -- no per-thread copy of the parameter block is ever made)
template <typename T>
__device__ __noinline__ void tile_bookkeep_cold(const FusedParams<T> &F, int32_t tile, int32_t round)
{
    tile_bookkeep<T>(F, tile, round);
}
template <typename T>
__device__ __noinline__ int32_t tile_pp_item_cold(const FusedParams<T> &F, int32_t tile, int32_t item, int cur)
{
    return tile_pp_item<T>(F, tile, item, cur);
}

// The kernel.  Two sources of work for a warp:
//   * the PP ready queue: every warp holds one ticket; when the item behind its ticket has been published (by the BK
//     of some tile) it is processed before the next sweep claim -- the tile's rows are then still in L2;
//   * the static stream of sweep claims: claim q of P.work[0], round = q / (T cpt), tile = (q / cpt) mod T,
//     chunk = q mod cpt (q < 2^32: the host cuts larger batches).
// Consecutive claims of a warp belong to DIFFERENT tiles (the grid covers about one tile per wave), so everything a
// claim needs to know about its tile is fetched while the PREVIOUS claim computes, between its passes:
//   after pass 0:  the count of the claim before is issued (release fence + atomic: who finishes a sweep LAST runs
//                  BK), the next claim id (atomic issued at the top, in flight during pass 0) is taken, and lane 0
//                  loads, for it: exit flag, this warp's ticket slot, the next tile's {PP items published, rounds
//                  book-kept} word and its PP-items-done counter, straight into the control block (lane 0 waits for
//                  that one batch of loads; the other 15 warps of the SM keep the memory system busy);
//   after pass 1:  if those say the next tile is ready, its lane flag bytes are loaded.
// A claim therefore starts without a round trip of its own; only when the next tile is NOT ready (few tiles, or the
// end of a batch) does the warp fall back to loading and waiting in place.  Nothing but the flag words, the pass
// counter and two lane-0 words is live across an item: 0 bytes of spills (spilled control state was 13 % of the
// kernel's L2 traffic).
template <typename T, int VEC, int DSEL, int VDEG>
__global__ void __launch_bounds__(kFBlock, DSEL > 0 ? 2 : 1) k_fused(const __grid_constant__ FusedParams<T> F)
{
    const DecodeParams<T> &P = F.P;
    cg::grid_group grid = cg::this_grid();
    fused_initial_fill<T>(F);
    grid.sync();

    __shared__ WarpCtl s_ctl[kFBlock / 32];
    // DISCIPLINE for the control block: lane 0 writes, __syncwarp(), every lane reads.  (Lanes of a warp are not in
    // lock step: "all lanes write the same value" races with the next update of the same field.)
    volatile WarpCtl &w = s_ctl[threadIdx.x >> 5];
    const int32_t lid = threadIdx.x & 31;
    if (lid == 0)
        w.cpt = (uint32_t)(((int32_t)P.C + (32 / (F.tl / VEC)) * F.rows_per_claim - 1) / ((32 / (F.tl / VEC)) * F.rows_per_claim));
    __syncwarp();
    const int32_t frames = (int32_t)P.frames;

    // claim id -> coordinates (lane 0)
    auto set_coord0 = [&](uint32_t q, bool next) {
        const uint32_t cpt = w.cpt, q32 = (uint32_t)q, per_round = cpt * (uint32_t)F.tiles;
        const uint32_t round = q32 / per_round, rem = q32 - round * per_round, tile = rem / cpt;
        if (next) { w.n_round = (int32_t)round; w.n_tile = (int32_t)tile; w.n_chunk = (int32_t)(rem - tile * cpt); }
        else { w.round = (int32_t)round; w.tile = (int32_t)tile; w.chunk = (int32_t)(rem - tile * cpt); }
    };
    // what a claim needs to know before it starts (lane 0): exit flag, ticket slot, the tile's words
    auto load_pre0 = [&](int32_t tile) {
        const int32_t done = ld_volatile(&P.ctrl[CTRL_COMPLETED]);
        const unsigned long long slot = *reinterpret_cast<const volatile unsigned long long *>(&F.ppq[w.ticket % (uint32_t)F.ppq_size]);
        const unsigned long long bk = ld_acquire_u64(&F.bk_word[tile]);
        const int32_t ppd = ld_volatile(&F.pp_done[tile]);
        w.done = done; w.slot = slot; w.bk = bk; w.ppd = ppd;
    };
    auto tile_ready = [&](int32_t round) {         // (uniform: reads only)
        const unsigned long long bk = w.bk;
        return round == 0 || ((int32_t)(uint32_t)bk >= round && w.ppd >= (int32_t)(bk >> 32));
    };
    // ---- completion counts: sig_* = finished claim not yet counted; cnt_* = count issued
    auto settle_count = [&]() {          // whoever finished a sweep LAST runs its bookkeeping
        __syncwarp();
        const int32_t ct = w.cnt_tile, cr = w.cnt_round;
        if (ct < 0) return;
        const bool last = (uint32_t)(w.cnt_val + 1) == w.cpt * (uint32_t)(cr + 1);
        __syncwarp();
        if (lid == 0) w.cnt_tile = -1;
        __syncwarp();
        if (last) {
            __threadfence();                                       // acquire: every claim of the sweep is visible
            tile_bookkeep_cold<T>(F, ct, cr);
        }
    };
    int32_t cnt_r = 0;                   // lane 0: result of the count just issued, not yet in the control block
    auto issue_count = [&](bool commit) {
        __syncwarp();
        const int32_t st = w.sig_tile, sr = w.sig_round;
        if (st < 0) return;
        settle_count();
        __threadfence();                                           // release: messages and flags before the count
        __syncwarp();
        if (lid == 0) {
            cnt_r = atomicAdd(&F.f_done[st], 1);                   // (its round trip overlaps the next pass unless committed here)
            if (commit) w.cnt_val = cnt_r;
            w.cnt_tile = st; w.cnt_round = sr;
            w.sig_tile = -1;
        }
        __syncwarp();
    };
    // ---- PP service: the item behind this warp's ticket, if it has been published
    auto serve_pp = [&]() -> bool {
        __syncwarp();
        const unsigned long long slot = w.slot;
        if ((uint32_t)(slot >> 32) != w.ticket + 1u) return false;
        __threadfence();                                           // acquire (the publisher released before writing the slot)
        const uint32_t e = (uint32_t)slot;                         // bits 8.. tile, bits 1..7 item, bit 0 parity of the sweep's round
        const int32_t tile = (int32_t)(e >> 8);
        const int32_t cnt = tile_pp_item_cold<T>(F, tile, (int32_t)((e >> 1) & 127u), (int)(e & 1u));
        __threadfence();                                           // release: shipped posteriors, new columns
        __syncwarp();
        if (lid == 0) {
            const int32_t done = atomicAdd(&F.pp_done[tile], 1) + 1;
            // every item of this PP complete: its retired frames are fully written out
            if (done == ld_volatile(&F.pp_expect[tile]) && cnt) atomicAdd(&P.ctrl[CTRL_COMPLETED], cnt);
            w.ticket = (uint32_t)atomicAdd(&P.ctrl[CTRL_PP_HEAD], 1);
        }
        __syncwarp();
        return true;
    };
    auto load_flags = [&](int32_t tile) {
        const uint8_t *p = F.lane_flags + tile * F.tl + (lid % (F.tl / VEC)) * VEC;
        if constexpr (VEC == 4) return (uint32_t)__ldcg(reinterpret_cast<const unsigned int *>(p));
        else return (uint32_t)__ldcg(reinterpret_cast<const unsigned short *>(p));
    };

    if (lid == 0) {
        const unsigned long long q = atomicAdd(reinterpret_cast<unsigned long long *>(P.work), 1ULL);
        set_coord0((uint32_t)q, false);
        w.ticket = (uint32_t)atomicAdd(&P.ctrl[CTRL_PP_HEAD], 1);
        w.sig_tile = -1; w.cnt_tile = -1;
        load_pre0(w.tile);
    }
    __syncwarp();
    bool have_flags = false;
    uint32_t flags = 0, n_flags = 0;          // (per lane: the flag bytes of this thread's VEC lanes)
    for (;;) {
        // ---- A. exit, PP service, readiness of the claim's tile
        __syncwarp();
        if (w.done >= frames) break;
        if (serve_pp()) {
            if (lid == 0) load_pre0(w.tile);
            have_flags = false;
            continue;
        }
        if (!tile_ready(w.round)) {
            // not ready: few tiles, or the tail of a batch.  Count what is pending (it may be what the tile waits
            // for), then poll, serving the PP queue meanwhile
            issue_count(true);
            settle_count();
            bool gone = false;
            for (;;) {
                __syncwarp();
                if (lid == 0) load_pre0(w.tile);
                __syncwarp();
                if (w.done >= frames) { gone = true; break; }
                if (tile_ready(w.round)) break;
                if (!serve_pp()) __nanosleep(100);
            }
            if (gone) break;
            have_flags = false;
            continue;                                              // (re-check the slot with the fresh loads)
        }
        if (!have_flags) flags = load_flags(w.tile);
        uint32_t nq = 0;                                               // (lane 0; claim ids stay below 2^32)
        if (lid == 0) nq = (uint32_t)atomicAdd(reinterpret_cast<unsigned long long *>(P.work), 1ULL);   // in flight during pass 0
        // ---- B. the claim.  Per pass, coordinates come from the control block again (volatile shared memory: the
        // compiler cannot keep views of the tile alive across the item and spill them)
        uint32_t bad = 0;
        const int32_t passes = F.rows_per_claim;
        for (int32_t r = 0; r < max(passes, 2); ++r) {
            if (r < passes) {
                const int32_t wy = 32 / (F.tl / VEC), claim_rows = wy * passes;
                const int32_t c0 = w.chunk * claim_rows, c1 = min(c0 + claim_rows, (int32_t)P.C);
                const int32_t lo = c0 + r * wy, hi = min(lo + wy, c1);
                if (lo < hi) bad |= fused_claim<T, VEC, DSEL, VDEG>(F, w.tile, w.round, flags, lo, hi, lid);
            }
            if (r == 0) {
                issue_count(false);                                    // the claim before this one
                if (lid == 0) {
                    set_coord0(nq, true);
                    // one batch of loads for the next claim (lane 0 waits for them here; the other warps of the SM
                    // keep the memory system busy meanwhile)
                    const int32_t nt = w.n_tile;
                    const int32_t p_done = ld_volatile(&P.ctrl[CTRL_COMPLETED]);
                    const unsigned long long p_slot = *reinterpret_cast<const volatile unsigned long long *>(&F.ppq[w.ticket % (uint32_t)F.ppq_size]);
                    const unsigned long long p_bk = ld_acquire_u64(&F.bk_word[nt]);
                    const int32_t p_ppd = ld_volatile(&F.pp_done[nt]);
                    w.done = p_done; w.slot = p_slot; w.bk = p_bk; w.ppd = p_ppd;
                }
            } else if (r == 1) {
                if (lid == 0) w.cnt_val = cnt_r;
                __syncwarp();
                const bool pf = w.done < frames && (uint32_t)(w.slot >> 32) != w.ticket + 1u && tile_ready(w.n_round);
                if (pf) n_flags = load_flags(w.n_tile);
                __syncwarp();
                if (lid == 0) w.n_have_flags = pf ? 1 : 0;
                __syncwarp();
            }
        }
        {
            // per-lane "some check unsatisfied" flags: OR over the warp's check rows, idempotent stores of 1
            const int32_t bx = F.tl / VEC, tx = lid % bx, tyw = lid / bx, tile = w.tile;
            for (int32_t o = bx; o < 32; o <<= 1) bad |= __shfl_xor_sync(0xffffffffu, bad, o);
            if (tyw == 0) {
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    if (bad >> k & 1) P.unsat[0][tile * F.tl + tx * VEC + k] = 1;
            }
        }
        __syncwarp();
        have_flags = w.n_have_flags != 0; flags = n_flags;
        __syncwarp();
        if (lid == 0) {
            w.sig_tile = w.tile; w.sig_round = w.round;                // counted after pass 0 of the next claim
            w.round = w.n_round; w.tile = w.n_tile; w.chunk = w.n_chunk;
        }
    }
}

template <typename T, int VEC, int DSEL, int VDEG>
static int launch_fused(qr_decoder *d, const FusedParams<T> &F, cudaStream_t stream)
{
    auto kern = k_fused<T, VEC, DSEL, VDEG>;
    int per_sm = 0;
    QR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFBlock, 0));
    if (per_sm < 1) return fail(QR_ERR_CUDA, "fused decoder kernel does not fit on an SM");
    const int grid = per_sm * d->sm_count;
    d->coop_grid = grid;
    FusedParams<T> Fc = F;
    void *args[] = {&Fc};
    QR_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(kFBlock), args, 0, stream));
    if (const char *wd = getenv("QAMRECON_FUSED_WATCHDOG")) {
        // debugging aid: if the kernel is still running after the given number of seconds, print the pipeline's
        // control words (read on a second stream while the kernel runs)
        const double limit = atof(wd);
        cudaStream_t side;
        cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
        const auto t0 = std::chrono::steady_clock::now();
        while (cudaStreamQuery(stream) == cudaErrorNotReady) {
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (el > limit) {
                std::vector<int32_t> ctrl(CTRL_WORDS), ctl(8 * (size_t)F.tiles);
                unsigned long long q = 0;
                cudaMemcpyAsync(ctrl.data(), F.P.ctrl, CTRL_WORDS * 4, cudaMemcpyDeviceToHost, side);
                cudaMemcpyAsync(&q, F.P.work, 8, cudaMemcpyDeviceToHost, side);
                cudaStreamSynchronize(side);
                fprintf(stderr, "[fused watchdog] %.1f s: frames %lld next_frame %d remaining %d minfin %d completed %d pp_head %d pp_reserve %d queue %llu cpt*tiles ? tiles %d R %d\n",
                        el, (long long)F.P.frames, ctrl[CTRL_NEXT_FRAME], ctrl[CTRL_REMAINING], ctrl[CTRL_MINFIN], ctrl[CTRL_COMPLETED],
                        ctrl[CTRL_PP_HEAD], ctrl[CTRL_PP_RESERVE], q, F.tiles, F.pp_items);
                for (int t = 0; t < F.tiles && t < 8; ++t) {
                    int32_t v[4]; unsigned long long bk = 0;
                    cudaMemcpyAsync(&v[0], F.f_done + t, 4, cudaMemcpyDeviceToHost, side);
                    cudaMemcpyAsync(&v[1], F.pp_done + t, 4, cudaMemcpyDeviceToHost, side);
                    cudaMemcpyAsync(&v[2], F.pp_expect + t, 4, cudaMemcpyDeviceToHost, side);
                    cudaMemcpyAsync(&v[3], F.rcount + t, 4, cudaMemcpyDeviceToHost, side);
                    cudaMemcpyAsync(&bk, F.bk_word + t, 8, cudaMemcpyDeviceToHost, side);
                    cudaStreamSynchronize(side);
                    fprintf(stderr, "[fused watchdog] tile %d: f_done %d pp_done %d pp_expect %d rcount %d bk_rounds %u bk_expect %u\n", t, v[0], v[1],
                            v[2], v[3], (unsigned)bk, (unsigned)(bk >> 32));
                }
                break;
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(50));
        }
        cudaStreamDestroy(side);
    }
    return QR_OK;
}

bool fused_eligible(const qr_graph *g)
{
    return !g->slot_nbr.empty() && g->max_cdeg <= kFusedMaxCheckDegree;
}

// QR_SCHED_AUTO.  The fused schedule pays off when a lane tile's live set (messages read + LLRs) stays in L2 between
// its d_v re-reads; beyond that it moves MORE bytes than the two-phase schedule (config 3: 75 MB per 32-lane tile,
// config 4: 604 MB -- both run the two-phase kernel).  It is also only tuned for graphs whose variables all have
// degree 3 (lean item, register-resident records); other graphs are eligible (QR_SCHED_FUSED) but not preferred.
bool fused_preferred(const qr_graph *g, size_t w)
{
    if (!fused_eligible(g) || g->var_deg != 3) return false;
    const size_t tile_bytes = (size_t)(g->E + g->N) * 32 * w;
    return tile_bytes <= ((size_t)72 << 20);
}

__global__ void k_build_lean(const Nbr4 *nbr, int64_t E, int32_t tl, NbrL *out)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < E) out[s] = make_lean_record(nbr[s], tl);
}

__global__ void k_init_fused(int32_t *ctl, int32_t n_ctl, int32_t *tile_minfin, int32_t tiles, int32_t *ctrl,
                             uint8_t *lane_flags, int32_t lanes, int64_t frames, int wpost0)
{
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_ctl) ctl[i] = 0;
    if (i < tiles) tile_minfin[i] = 0x7fffffff;
    // first generation: lane l runs frame l, first half-iteration; its posterior is stored only if 0 iterations are allowed
    if (i < lanes) lane_flags[i] = (int64_t)i < frames ? (uint8_t)(1 | 2 | (wpost0 ? 4 : 0)) : (uint8_t)0;
    if (i == 0) { ctrl[CTRL_MINFIN] = 0x7fffffff; ctrl[CTRL_COMPLETED] = 0; ctrl[CTRL_PP_HEAD] = 0; ctrl[CTRL_PP_RESERVE] = 0; }
}

template <typename T, int VEC>
static int run_batch_fused_v(qr_decoder *d, const DecodeParams<T> &P, cudaStream_t stream)
{
    const qr_graph *g = d->g;
    if (!fused_eligible(g))
        return fail(QR_ERR_INVALID, "fused schedule needs check degrees <= 8 (and fewer than 2^27 variables)");
    const size_t w = sizeof(T);
    if (!d->c2v2) {
        QR_CUDA_CHECK(cudaMalloc(&d->c2v2, (size_t)g->E * (size_t)d->lanes * w));
        QR_CUDA_CHECK(cudaMemset(d->c2v2, 0, (size_t)g->E * (size_t)d->lanes * w));
    }
    int32_t tl = d->fused_tile > 0 ? d->fused_tile : 32;
    tl = std::min<int32_t>(tl, kFusedMaxTile);
    while (tl > 32 && (P.lanes % tl || (tl / VEC) > 32)) tl /= 2;   // a check row is shared by at most one warp
    const int32_t tiles = P.lanes / tl, tiles_max = d->lanes / 32;
    constexpr int32_t kMaxPPItems = 128;   // (7 bits in a ready-queue entry)
    if (!d->fused_ctl) {
        // per tile: f_done, pp_done, pp_expect, rcount, tile_minfin, bk_word (2 words); the per-lane flag bytes; the
        // refill lists; the PP ready queue
        QR_CUDA_CHECK(cudaMalloc((void **)&d->fused_ctl, (size_t)8 * tiles_max * sizeof(int32_t) + (size_t)d->lanes));
        QR_CUDA_CHECK(cudaMalloc((void **)&d->fused_rlist, (size_t)d->lanes * sizeof(RefillEntry)));
        QR_CUDA_CHECK(cudaMalloc((void **)&d->fused_ppq, (size_t)tiles_max * kMaxPPItems * sizeof(unsigned long long)));
    }
    FusedParams<T> F;
    F.P = P;
    F.nbr = reinterpret_cast<const Nbr4 *>(g->d_slot_nbr);
    F.c2v[0] = static_cast<T *>(d->c2v);
    F.c2v[1] = static_cast<T *>(d->c2v2);
    F.post = d->fused_store_post ? static_cast<T *>(d->post) : nullptr;
    F.nbr_lean = nullptr;
    if (std::is_same<T, float>::value && d->regular_degree == 6 && g->var_deg == 3 && d->fused_lean &&
        (size_t)(g->E + g->N) * tl * 4 < ((size_t)1 << 32)) {
        if (!d->fused_nbrl || d->fused_nbrl_tl != tl) {
            if (!d->fused_nbrl) QR_CUDA_CHECK(cudaMalloc(&d->fused_nbrl, (size_t)g->E * sizeof(NbrL)));
            k_build_lean<<<(unsigned)((g->E + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const Nbr4 *>(g->d_slot_nbr), g->E, tl,
                                                                          static_cast<NbrL *>(d->fused_nbrl));
            QR_CUDA_CHECK(cudaGetLastError());
            d->fused_nbrl_tl = tl;
        }
        F.nbr_lean = d->fused_nbrl;
    }
    F.tl = tl;
    F.tiles = tiles;
    F.hints = d->fused_hints;
    F.rows_per_claim = d->fused_rpc > 0 ? d->fused_rpc : 4;
    F.park_rounds = d->fused_park;
    F.pp_items = d->fused_pp_items > 0 ? std::min(d->fused_pp_items, kMaxPPItems)
                                       : (int32_t)std::min<int64_t>(kMaxPPItems, std::max<int64_t>(4, g->N / 512));
    F.f_done = d->fused_ctl;
    F.pp_done = d->fused_ctl + tiles_max;
    F.pp_expect = d->fused_ctl + 2 * tiles_max;
    F.rcount = d->fused_ctl + 3 * tiles_max;
    F.bk_word = reinterpret_cast<unsigned long long *>(d->fused_ctl + 4 * tiles_max);
    F.tile_minfin = d->fused_ctl + 6 * tiles_max;
    F.lane_flags = reinterpret_cast<uint8_t *>(d->fused_ctl + 8 * tiles_max);
    F.rlist = static_cast<RefillEntry *>(d->fused_rlist);
    F.ppq = static_cast<unsigned long long *>(d->fused_ppq);
    F.ppq_size = tiles_max * kMaxPPItems;
    QR_CUDA_CHECK(cudaMemsetAsync(d->fused_ppq, 0, (size_t)F.ppq_size * sizeof(unsigned long long), stream));
    k_init_fused<<<(std::max<int32_t>(6 * tiles_max, P.lanes) + 255) / 256, 256, 0, stream>>>(
        d->fused_ctl, 6 * tiles_max, F.tile_minfin, tiles_max, P.ctrl, F.lane_flags, P.lanes, P.frames,
        (F.post && P.maxiter == 0) ? 1 : 0);
    QR_CUDA_CHECK(cudaGetLastError());
    // <.., 6, 3>: check-regular degree 6, every variable of degree 3 (config 2); in float it runs the lean item and
    // needs the lean records, otherwise the float kernel for any variable degree takes over
    const bool reg3 = g->var_deg == 3 && (!std::is_same<T, float>::value || F.nbr_lean != nullptr);
    if (d->regular_degree == 6 && reg3) return launch_fused<T, VEC, 6, 3>(d, F, stream);
    if (d->regular_degree == 6) return launch_fused<T, VEC, 6, 0>(d, F, stream);
    return launch_fused<T, VEC, 0, 0>(d, F, stream);
}

template <typename T>
int run_batch_fused(qr_decoder *d, const DecodeParams<T> &P, cudaStream_t stream)
{
    return run_batch_fused_v<T, 16 / sizeof(T)>(d, P, stream);
}

template int run_batch_fused<float>(qr_decoder *, const DecodeParams<float> &, cudaStream_t);
template int run_batch_fused<double>(qr_decoder *, const DecodeParams<double> &, cudaStream_t);

}  // namespace qr
