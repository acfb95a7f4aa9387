// FUSED FLOODING ITERATION (QR_SCHED_FUSED): check update, variable sums and the syndrome test of
// one flooding iteration in ONE pass, for graphs whose variables all have degree 3
// (reference: decoder.pyx:285-298, 322-369, 235-257, 424-433 -- same arithmetic, same order).
//
// No posterior array is stored.  A check rebuilds, for each of its variables,
//     post[v] = ((llr[v] + c2v[e0]) + c2v[e1]) + c2v[e2]       (e0 < e1 < e2: the variable's edges,
//                                                                ascending edge id, decoder.pyx:291-293)
// from the PREVIOUS iteration's messages, takes v2c = post[v] - c2v[own edge] (decoder.pyx:295-297),
// runs the same check-node arithmetic as the two-phase schedule and writes its outgoing messages to
// the OTHER of two message buffers (flooding = Jacobi: nobody may see this iteration's messages
// before the next one).  fp64 results are bit-identical to the two-phase schedule and the reference.
//
// Why: the two-phase schedule streams 3 E w + (N..3N) w + 2 N w bytes of HBM per frame-iteration.
// Here every c2v row is read by three checks (its own and the two others of its variable) and every
// llr row by three, so if those re-reads hit L2 the HBM traffic is (2 E + N) w + C -- about half.
// To make them hit, lanes are stored TILE-MAJOR,
//     c2v[buf] [tile][E][TL]     llr [tile][N][TL]     synd [tile][C][TL]       (TL = 32 or 64 lanes)
// and the whole grid sweeps one tile after the other (work-stealing claims are handed out in tile
// order), so the live set is one tile: (E + N) TL w = 33 MB (fp32, TL = 32) of the 126 MB L2.
// Outgoing messages are stored with an evict-first L2 policy (they are not needed again before the
// next sweep), loads keep the default / evict-last policy.
//
// The lane state machine, continuous batching and result semantics are those of
// qr_decode_core.cuh; only the phases differ: one fused phase per step, lane bookkeeping by the last
// CTA to finish it, and -- when a frame finished -- a refill phase that ships
// post = llr + sum c2v of the finished lanes and loads the next frames' columns.
#pragma once

#include "qr_decode_core.cuh"

#ifndef FUSED_HALF_BATCH
#define FUSED_HALF_BATCH 0
#endif

namespace qr {

struct alignas(16) Nbr4 {
    int32_t vp;          // variable id | (position of the own edge among n0..n2) << 28
    int32_t n0, n1, n2;  // CSR slots of the variable's edges, ascending edge id
};

// CTRL_MINFIN: smallest iteration count a frame of this batch finished successfully with, as the fused phase of
// the current step sees it; CTRL_MINFIN_NEXT: the same, being updated by the bookkeeping of the current step
enum : int { CTRL_ARRIVE = 6, CTRL_MINFIN = 7, CTRL_MINFIN_NEXT = 8 };

// A lane's posterior is STORED (by the check that holds the variable's first edge) in the steps it may finish in:
// its last allowed iteration, or any iteration from one below the earliest success seen so far in the batch.
// A frame that finishes in such a step ships from the stored column (N elements); one that finishes earlier than
// ever seen falls back to rebuilding llr + sum c2v (E + N elements).  Either way the same values.
QR_HD bool stores_post(int32_t iter, int32_t maxiter, int32_t minfin)
{
    return iter >= maxiter || iter >= minfin - 1;
}

template <typename T>
struct FusedParams {
    DecodeParams<T> P;   // graph, lane state, batch, control words; P.llr / P.synd are the tile-major arrays
    const Nbr4 *nbr;     // [E]
    T *c2v[2];           // message buffers, step parity selects the one being read
    T *post;             // [tile][N][TL] posteriors of the lanes that may finish this step (null: always rebuild)
    int32_t *postok;     // [2][lanes] per state buffer: the retiring frame's posterior column is valid
    int32_t tl;          // lanes per tile
    int32_t tiles;
    int32_t hints;       // L2 policy: 0 none, 1 stores evict-first, 2 + loads evict-last
    int32_t static_share;     // per mille of a step's claims dealt statically (no atomic), the rest is work-stolen
    int32_t rows_per_claim;   // checks per thread and work-stealing claim (register-staged phase)
    int32_t prefetch;    // bit 0: stream the next tile into L2 in address order, one claim ahead of the sweep; bit 1: next item
};

// ---- 16-byte row accesses with an L2 policy (device) / plain (host emulation)
template <typename V>
QR_HD V ld_pol(const V *p, uint64_t pol)
{
#if defined(__CUDA_ARCH__)
    static_assert(sizeof(V) == 16 || sizeof(V) == 8, "fused schedule moves 16- or 8-byte lane vectors");
    V v;
    if constexpr (sizeof(V) == 16) {
        uint32_t r0, r1, r2, r3;
        asm volatile("ld.global.cg.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "l"(p), "l"(pol));
        uint32_t t[4] = {r0, r1, r2, r3};
        memcpy(&v, t, 16);
    } else {
        uint32_t r0, r1;
        asm volatile("ld.global.cg.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(r0), "=r"(r1) : "l"(p), "l"(pol));
        uint32_t t[2] = {r0, r1};
        memcpy(&v, t, 8);
    }
    return v;
#else
    (void)pol;
    return *p;
#endif
}
template <typename V>
QR_HD void st_pol(V *p, const V &val, uint64_t pol)
{
#if defined(__CUDA_ARCH__)
    static_assert(sizeof(V) == 16 || sizeof(V) == 8, "fused schedule moves 16- or 8-byte lane vectors");
    if constexpr (sizeof(V) == 16) {
        uint32_t t[4];
        memcpy(t, &val, 16);
        asm volatile("st.global.cg.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;"
                     :: "l"(p), "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "l"(pol) : "memory");
    } else {
        uint32_t t[2];
        memcpy(t, &val, 8);
        asm volatile("st.global.cg.L2::cache_hint.v2.u32 [%0], {%1, %2}, %3;" :: "l"(p), "r"(t[0]), "r"(t[1]), "l"(pol)
                     : "memory");
    }
#else
    (void)pol;
    *p = val;
#endif
}

// pointers of one tile
template <typename T>
struct TileView {
    const T *c_old;
    T *c_new;
    const T *llr;
    const uint8_t *synd;
    T *post;
    int32_t tl;
    int32_t item_prefetch;   // pull the rows of the thread group's next item into L2 (FusedParams::prefetch bit 1)
};

template <typename T>
QR_HD TileView<T> tile_view(const FusedParams<T> &F, int cur, int32_t tile)
{
    TileView<T> V;
    V.tl = F.tl;
    V.item_prefetch = (F.prefetch >> 1) & 3;   // bit 0: next item's rows into L2, bit 1: next item's records into L1
    V.c_old = F.c2v[cur] + (int64_t)tile * F.P.E * F.tl;
    V.c_new = F.c2v[cur ^ 1] + (int64_t)tile * F.P.E * F.tl;
    V.llr = F.P.llr + (int64_t)tile * F.P.N * F.tl;
    V.synd = F.P.synd + (int64_t)tile * F.P.C * F.tl;
    V.post = F.post ? F.post + (int64_t)tile * F.P.N * F.tl : nullptr;
    return V;
}

template <typename T, int VEC>
QR_HD void mark_post_lanes(const FusedParams<T> &F, LaneInfo<VEC> &L, int32_t minfin)
{
    L.wpost = 0;
    if (!F.post) return;
#pragma unroll
    for (int k = 0; k < VEC; ++k)
        if ((L.active >> k & 1) && stores_post(L.iter[k], F.P.maxiter, minfin)) L.wpost |= 1u << k;
}

// FUSED item: internal check `ci` (first CSR slot slot0, degree D) for the thread's VEC lanes at
// offset `lt` inside the tile.  q[i] = neighbour record of slot slot0 + i.  Returns, per lane, 1 if
// this check is NOT satisfied by the posteriors (decoder.pyx:235-257).
template <typename T, int VEC, int D, bool ANYFRESH = true>
QR_HD uint32_t fused_item(const TileView<T> &V, const LaneInfo<VEC> &L, int32_t lt, int32_t ci, int32_t slot0,
                          const Nbr4 (&q)[D], uint64_t pol_ld, uint64_t pol_st)
{
    using VT = Vec<T, VEC>;
    const int32_t tl = V.tl;
    const Vec<uint8_t, VEC> sy = *reinterpret_cast<const Vec<uint8_t, VEC> *>(V.synd + (int64_t)ci * tl + lt);
    uint32_t par = 0;
#pragma unroll
    for (int k = 0; k < VEC; ++k) par |= (uint32_t)(sy.v[k] & 1u) << k;
    VT x[D];
    // the 4 D row loads go out in batches of HB edges (HB = D: all at once; smaller: fewer registers held
    // by loads in flight, so two CTAs fit an SM)
#ifdef FUSED_BATCH_EDGES
    constexpr int HB = FUSED_BATCH_EDGES < D ? FUSED_BATCH_EDGES : D;
#else
    constexpr int HB = (FUSED_HALF_BATCH && D > 3) ? (D + 1) / 2 : D;
#endif
#pragma unroll
    for (int i0 = 0; i0 < D; i0 += HB) {
        VT m[HB][3], ch[HB];
#pragma unroll
        for (int j = 0; j < HB; ++j) {
            const int i = i0 + j;
            if (i < D) {
                ch[j] = ld_pol(reinterpret_cast<const VT *>(V.llr + (int64_t)(q[i].vp & 0x0fffffff) * tl + lt), pol_ld);
                m[j][0] = ld_pol(reinterpret_cast<const VT *>(V.c_old + (int64_t)q[i].n0 * tl + lt), pol_ld);
                m[j][1] = ld_pol(reinterpret_cast<const VT *>(V.c_old + (int64_t)q[i].n1 * tl + lt), pol_ld);
                m[j][2] = ld_pol(reinterpret_cast<const VT *>(V.c_old + (int64_t)q[i].n2 * tl + lt), pol_ld);
            }
        }
#pragma unroll
        for (int j = 0; j < HB; ++j) {
            const int i = i0 + j;
            if (i < D) {
                const int own = (int)((uint32_t)q[i].vp >> 28);
                VT pv;
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    const bool fresh = ANYFRESH && (L.fresh >> k & 1) != 0;   // first half-iteration: c2v == 0 (decoder.pyx:408)
                    const T c0 = fresh ? (T)0 : m[j][0].v[k];
                    const T c1 = fresh ? (T)0 : m[j][1].v[k];
                    const T c2 = fresh ? (T)0 : m[j][2].v[k];
                    T post = ch[j].v[k] + c0;                            // decoder.pyx:291-293, ascending edge id
                    post = post + c1;
                    post = post + c2;
                    pv.v[k] = post;
                    par ^= (uint32_t)(post < (T)0) << k;                 // decoder.pyx:244 (strict <)
                    const T mine = own == 0 ? c0 : (own == 1 ? c1 : c2);
                    x[i].v[k] = post - mine;                             // decoder.pyx:295-297
                }
                // the check holding the variable's FIRST edge keeps the posterior of lanes that may finish now
                if (L.wpost && own == 0)
                    *reinterpret_cast<VT *>(V.post + (int64_t)(q[i].vp & 0x0fffffff) * tl + lt) = pv;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        T xs[D];
#pragma unroll
        for (int i = 0; i < D; ++i) xs[i] = x[i].v[k];
        MathOf<T>::template run<D, D>(D, xs, (sy.v[k] & 1u) != 0);
#pragma unroll
        for (int i = 0; i < D; ++i) x[i].v[k] = xs[i];
    }
#pragma unroll
    for (int i = 0; i < D; ++i)
        st_pol(reinterpret_cast<VT *>(V.c_new + (int64_t)(slot0 + i) * tl + lt), x[i], pol_st);
    return par & L.active;
}

template <int D>
QR_HD void load_nbr_row(const Nbr4 *__restrict__ tab, int32_t first, Nbr4 (&q)[D])
{
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int i = 0; i < D; ++i) {
        const int4 t = __ldg(reinterpret_cast<const int4 *>(tab + first + i));
        q[i].vp = t.x; q[i].n0 = t.y; q[i].n1 = t.z; q[i].n2 = t.w;
    }
#else
    for (int i = 0; i < D; ++i) q[i] = tab[first + i];
#endif
}

// checks k = first, first + stride, ... of one degree bin; the neighbour rows of the NEXT check are
// fetched while the current one is in flight
template <typename T, int VEC, int D>
QR_HD uint32_t run_fused_bin(const TileView<T> &V, const Nbr4 *nbr, const LaneInfo<VEC> &L, int32_t lt,
                             const CheckBin &bin, int32_t first, int32_t stride, uint64_t pol_ld, uint64_t pol_st)
{
    uint32_t bad = 0;
    if constexpr (FUSED_HALF_BATCH != 0) {
        // register-lean variant: no index prefetch into registers (latency is covered by the second resident
        // CTA).  Instead the rows of the thread group's NEXT item are pulled into L2 while this one is
        // processed: the bx threads of a group split the 4 D row ids of the next neighbour record row among
        // themselves (ids read straight from the table, L1), one prefetch.global.L2 each -- no register is
        // held, and the demand loads of the next item find their compulsory DRAM misses already in L2.
        for (int32_t k = first; k < bin.count; k += stride) {
#if defined(__CUDA_ARCH__)
            if ((V.item_prefetch & 1) && k + stride < bin.count) {
                const int32_t *raw = reinterpret_cast<const int32_t *>(nbr + bin.slot_begin + (k + stride) * D);
                const int32_t bx = V.tl / VEC, tx = lt / VEC;
                for (int32_t r = tx; r < 4 * D; r += bx) {
                    const int32_t id = __ldg(raw + r);
                    const T *row = (r & 3) == 0 ? V.llr + (int64_t)(id & 0x0fffffff) * V.tl : V.c_old + (int64_t)id * V.tl;
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(row));
                }
            }
#endif
#if defined(__CUDA_ARCH__)
            if ((V.item_prefetch & 2) && k + stride < bin.count) {
                // the neighbour records of the NEXT item into L1: its first dependent load then costs an L1 hit
                const Nbr4 *nx = nbr + bin.slot_begin + (k + stride) * D;
                asm volatile("prefetch.global.L1 [%0];" :: "l"(nx));
                asm volatile("prefetch.global.L1 [%0];" :: "l"(nx + D - 1));
            }
#endif
            Nbr4 cur[D];
            load_nbr_row<D>(nbr, bin.slot_begin + k * D, cur);
            // steady state (no lane of the thread on its first half-iteration): no per-element selects
            if (L.fresh) bad |= fused_item<T, VEC, D, true>(V, L, lt, bin.chk_begin + k, bin.slot_begin + k * D, cur, pol_ld, pol_st);
            else bad |= fused_item<T, VEC, D, false>(V, L, lt, bin.chk_begin + k, bin.slot_begin + k * D, cur, pol_ld, pol_st);
        }
    } else {
        Nbr4 cur[D], nxt[D];
        if (first < bin.count) load_nbr_row<D>(nbr, bin.slot_begin + first * D, cur);
        for (int32_t k = first; k < bin.count; k += stride) {
            const int32_t kn = k + stride;
            if (kn < bin.count) load_nbr_row<D>(nbr, bin.slot_begin + kn * D, nxt);
            bad |= fused_item<T, VEC, D>(V, L, lt, bin.chk_begin + k, bin.slot_begin + k * D, cur, pol_ld, pol_st);
#pragma unroll
            for (int i = 0; i < D; ++i) cur[i] = nxt[i];
        }
    }
    return bad;
}

template <typename T, int VEC>
QR_HD uint32_t run_fused_bin_any(const TileView<T> &V, const Nbr4 *nbr, const LaneInfo<VEC> &L, int32_t lt,
                                 const CheckBin &bin, int32_t first, int32_t stride, uint64_t pol_ld,
                                 uint64_t pol_st)
{
    switch (bin.degree) {
    case 2: return run_fused_bin<T, VEC, 2>(V, nbr, L, lt, bin, first, stride, pol_ld, pol_st);
    case 3: return run_fused_bin<T, VEC, 3>(V, nbr, L, lt, bin, first, stride, pol_ld, pol_st);
    case 4: return run_fused_bin<T, VEC, 4>(V, nbr, L, lt, bin, first, stride, pol_ld, pol_st);
    case 5: return run_fused_bin<T, VEC, 5>(V, nbr, L, lt, bin, first, stride, pol_ld, pol_st);
    case 6: return run_fused_bin<T, VEC, 6>(V, nbr, L, lt, bin, first, stride, pol_ld, pol_st);
    case 7: return run_fused_bin<T, VEC, 7>(V, nbr, L, lt, bin, first, stride, pol_ld, pol_st);
    case 8: return run_fused_bin<T, VEC, 8>(V, nbr, L, lt, bin, first, stride, pol_ld, pol_st);
    default: return 0;   // not reachable: the fused schedule is only chosen when max check degree <= 8
    }
}

constexpr int kFusedMaxCheckDegree = 8;

// ---- refill phase, one (lane, variable) element: ship post = llr + sum c2v[cur] of a finished frame
// (what the two-phase schedule keeps in its post array), bring in the next frame's channel LLR
template <typename T>
QR_HD void fused_refill_var_elem(const FusedParams<T> &F, int cur, const LaneState &s, int32_t lane, int32_t n,
                                 bool post_valid = false)
{
    const DecodeParams<T> &P = F.P;
    const int32_t tile = lane / F.tl, lt = lane % F.tl;
    const int64_t at = ((int64_t)tile * P.N + n) * F.tl + lt;
    if (s.retire >= 0 && P.post_out) {
        const int64_t idx = (int64_t)s.retire * P.N + n;
        if (ld_stream(&P.iters[s.retire]) == 0) {
            // never iterated: same semantics as refill_var_elem (copy / llr + 0.0)
            const bool copied = *static_cast<const volatile uint8_t *>(&P.success[s.retire]) != 0;
            if (copied && P.llr_in_f64 && P.post_out_f64) {
                static_cast<double *>(P.post_out)[idx] = static_cast<const double *>(P.llr_in)[idx];
            } else {
                double val = (double)load_input_llr<T>(P.llr_in, P.llr_in_f64, idx);
                if (!copied) val = val + 0.0;
                store_output_llr(P.post_out, P.post_out_f64, idx, val);
            }
        } else if (post_valid) {
            store_output_llr(P.post_out, P.post_out_f64, idx, (double)ld_stream(&F.post[at]));
        } else {
            const T *c = F.c2v[cur] + (int64_t)tile * P.E * F.tl + lt;
            T acc = ld_stream(&P.llr[at]);
            for (int j = 0; j < 3; ++j) acc = acc + ld_stream(&c[(int64_t)P.var_slot[3 * n + j] * F.tl]);
            store_output_llr(P.post_out, P.post_out_f64, idx, (double)acc);
        }
    }
    if (s.frame >= 0 && s.fresh)
        P.llr[at] = load_input_llr<T>(P.llr_in, P.llr_in_f64, (int64_t)s.frame * P.N + n);
}

// Before bookkeep_lanes of a step: remember, for the lanes that finish in it, whether the phase stored their
// posterior, and fold their iteration counts into the running minimum.  `nxt` = state buffer being written.
template <typename T, int VEC>
QR_HD void fused_note_finishers(const FusedParams<T> &F, int nxt, const LaneInfo<VEC> &L, int32_t minfin)
{
    if (!F.post) return;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        if (!((L.fin_ok | L.fin_fail) >> k & 1)) continue;
        F.postok[(int64_t)nxt * F.P.lanes + L.l0 + k] = stores_post(L.iter[k], F.P.maxiter, minfin) ? 1 : 0;
        if ((L.fin_ok >> k & 1) && L.iter[k] > 0) {   // (0 iterations = input already consistent: copy path, no signal)
#if defined(__CUDA_ARCH__)
            atomicMin(&F.P.ctrl[CTRL_MINFIN_NEXT], L.iter[k]);
#else
            if (L.iter[k] < F.P.ctrl[CTRL_MINFIN_NEXT]) F.P.ctrl[CTRL_MINFIN_NEXT] = L.iter[k];
#endif
        }
    }
}

template <typename T>
QR_HD void fused_refill_chk_elem(const FusedParams<T> &F, const LaneState &s, int32_t lane, int32_t ci)
{
    const DecodeParams<T> &P = F.P;
    const int32_t tile = lane / F.tl, lt = lane % F.tl;
    if (s.frame >= 0 && s.fresh)
        P.synd[((int64_t)tile * P.C + ci) * F.tl + lt] = P.synd_in[(int64_t)s.frame * P.C + P.chk_order[ci]];
}

}  // namespace qr
