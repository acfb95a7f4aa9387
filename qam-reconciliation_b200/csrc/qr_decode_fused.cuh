// FUSED FLOODING ITERATION (QR_SCHED_FUSED): check update, variable sums and the syndrome test of
// one flooding iteration in ONE pass (reference: decoder.pyx:285-298, 322-369, 235-257, 424-433 --
// same arithmetic, same order), for graphs with check degrees <= 8 and any variable degrees.
//
// No posterior array is stored.  A check rebuilds, for each of its variables,
//     post[v] = ((llr[v] + c2v[e0]) + c2v[e1]) + ...          (e0 < e1 < ...: the variable's edges,
//                                                               ascending edge id, decoder.pyx:291-293)
// from the PREVIOUS iteration's messages, takes v2c = post[v] - c2v[own edge] (decoder.pyx:295-297),
// runs the same check-node arithmetic as the two-phase schedule and writes its outgoing messages to
// the OTHER of two message buffers (flooding = Jacobi: nobody may see this iteration's messages
// before the next one).  fp64 results are bit-identical to the two-phase schedule and the reference.
//
// Why: the two-phase schedule streams 3 E w + (N..3N) w + 2 N w bytes of HBM per frame-iteration.
// Here every c2v row is read by d_v checks (its own and the others of its variable) and every llr row
// by d_v, so if those re-reads hit L2 the HBM traffic is (2 E + N) w + C -- about half.  To make them
// hit, lanes are stored TILE-MAJOR,
//     c2v[buf] [tile][E][TL]     llr [tile][N][TL]     synd [tile][C][TL]       (TL = 32 .. 128 lanes)
// and work is handed out in tile order, so the live set is one or two tiles: (E + N) TL w = 33 MB
// (config 2, fp32, TL = 32) of the 126 MB L2.  Outgoing messages are stored with an evict-first L2
// policy (they are not needed again before the next sweep).
//
// TILE PIPELINE (round 2).  Frames in different lanes never interact, so nothing in the algorithm needs a
// grid-wide barrier: the only ordering is per tile -- sweep r+1 of a tile after sweep r of the SAME tile, its
// lane bookkeeping and its refill.  Work comes from two queues:
//   F(t,r)   the fused sweep of tile t in round r, cut into claims of a few checks per warp; claims are handed out
//            from one counter in the order (r, t, chunk), so the grid sweeps tile after tile;
//   BK(t,r)  lane bookkeeping of tile t (decoder.pyx:431-436: finished? iteration limit? next frame), run by
//            whichever warp finishes the LAST claim of F(t,r); if a frame finished it publishes
//   PP(t,r)  post-processing of tile t, R items in a ready queue that warps serve before taking their next claim:
//            ship the posteriors of the frames BK retired, load the channel LLRs and syndromes of the frames it
//            admitted.  This happens right after the sweep, when the tile's rows are still in L2 -- the column
//            accesses of a refill (one 32-byte sector per element) hit L2 instead of DRAM -- and it overlaps the
//            sweeps of the following tiles.  Nothing is published for a tile without finished frames.
// The only waits: F(t,r+1) starts after BK(t,r) and PP(t,r), both a whole round earlier in the stream.
// No cooperative barrier after the initial fill; converged frames leave and new ones enter tile by tile.
#pragma once

#include "qr_decode_core.cuh"

#ifndef FUSED_BATCH_EDGES
#define FUSED_BATCH_EDGES 4
#endif

namespace qr {

// Neighbour record of one CSR slot (16 bytes, one LDG.128 per edge):
//   variable degree <= 3:  vp = variable | own position << 28 | degree << 30,  n0..n2 = CSR slots of the
//                          variable's edges in ascending edge id (unused ones = -1)
//   variable degree  > 3:  vp = variable | 1 << 31 ... written as  vp = variable | kNbrExt,
//                          n0 = first index of the variable's slot list in var_slot, n1 = degree, n2 = own position
struct alignas(16) Nbr4 {
    int32_t vp;
    int32_t n0, n1, n2;
};
constexpr int32_t kNbrVarMask = 0x07ffffff;      // variable ids < 2^27
constexpr int32_t kNbrExt = 0x08000000;          // bit 27: extended record
constexpr int kFusedMaxCheckDegree = 8;
constexpr int kFusedMaxVarDegree = 64;

QR_HD int32_t nbr_var(const Nbr4 &q) { return q.vp & kNbrVarMask; }
QR_HD bool nbr_ext(const Nbr4 &q) { return (q.vp & kNbrExt) != 0; }
QR_HD int nbr_own(const Nbr4 &q) { return nbr_ext(q) ? q.n2 : (int)(((uint32_t)q.vp >> 28) & 3u); }
QR_HD int nbr_deg(const Nbr4 &q) { return nbr_ext(q) ? q.n1 : (int)((uint32_t)q.vp >> 30); }

// control words of the fused schedule (DecodeParams::ctrl)
enum : int { CTRL_MINFIN = 7, CTRL_COMPLETED = 9, CTRL_PP_HEAD = 10, CTRL_PP_RESERVE = 11 };

// A lane's posterior is STORED (by the check that holds the variable's first edge) in the sweeps it may finish in:
// its last allowed iteration, or any iteration from one below the earliest success seen so far in the batch.
// A frame that finishes in such a sweep ships from the stored column (N elements); one that finishes earlier than
// ever seen falls back to rebuilding llr + sum c2v (E + N elements).  Either way the same values.
QR_HD bool stores_post(int32_t iter, int32_t maxiter, int32_t minfin)
{
    return iter >= maxiter || iter >= minfin - 1;
}

// one lane of a tile that PP has to serve
struct RefillEntry {
    int32_t lane;        // lane inside the tile
    int32_t retire;      // frame whose posteriors are shipped (-1: none)
    int32_t frame;       // frame admitted to the lane (-1: none, the lane goes idle)
    int32_t post_valid;  // the sweep stored the retiring frame's posterior column
};

template <typename T>
struct FusedParams {
    DecodeParams<T> P;   // graph, lane state (st[0], unsat[0]), batch, control words; P.llr / P.synd are tile-major
    const Nbr4 *nbr;     // [E]
    const void *nbr_lean;   // [E] NbrL records (float mode, every variable of degree 3, check-regular degree 6), or null
    T *c2v[2];           // message buffers, round parity selects the one being read
    T *post;             // [tile][N][TL] posteriors of the lanes that may finish in the sweep (null: always rebuild)
    int32_t tl;          // lanes per tile
    int32_t tiles;
    int32_t hints;       // L2 policy: 0 none, 1 stores evict-first, 2 + loads evict-last
    int32_t rows_per_claim;   // checks per thread and claim
    int32_t park_rounds; // a finished lane waits up to this many rounds for the other lanes of its 8-lane sector group, so
                         // that the group is shipped and refilled together (one sector per row for 8 frames); 0: never
    // tile pipeline (device arrays, [tiles] each; monotonic counters)
    int32_t pp_items;    // R: items per PP
    int32_t *f_done;     // finished claims of the tile, all rounds
    unsigned long long *bk_word;   // PP items published so far << 32 | rounds whose bookkeeping is done
    uint8_t *lane_flags; // [lanes] what a sweep reads per lane: bit 0 runs a frame, bit 1 first half-iteration, bit 2 store the posterior
    int32_t *pp_done;    // finished PP items of the tile, all rounds
    int32_t *pp_expect;  // PP items published for the tile, all rounds (the next sweep waits for pp_done to reach it)
    unsigned long long *ppq;   // ready queue of PP items: ring of (ticket + 1) << 32 | tile << 8 | item << 1 | round parity
    int32_t ppq_size;
    int32_t *tile_minfin;   // CTRL_MINFIN as the tile's current sweep uses it (fixed from BK to BK)
    int32_t *rcount;     // entries in the tile's refill list
    RefillEntry *rlist;  // [tiles][TL]
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

// 16-byte row accesses WITHOUT a policy register (the lean item: three 64-bit policy operands are six registers it
// does not have): loads bypass L1 (.cg), stores are either plain or streaming (.cs: evict-first in L2, the same
// effect as the evict_first policy of st_pol)
QR_HD Vec<float, 4> ld_row16(const void *p)
{
#if defined(__CUDA_ARCH__)
    const float4 r = __ldcg(reinterpret_cast<const float4 *>(p));
    Vec<float, 4> v;
    v.v[0] = r.x; v.v[1] = r.y; v.v[2] = r.z; v.v[3] = r.w;
    return v;
#else
    return *reinterpret_cast<const Vec<float, 4> *>(p);
#endif
}
QR_HD void st_row16(void *p, const Vec<float, 4> &val, bool streaming)
{
#if defined(__CUDA_ARCH__)
    const float4 r = make_float4(val.v[0], val.v[1], val.v[2], val.v[3]);
    if (streaming) __stcs(reinterpret_cast<float4 *>(p), r);
    else __stcg(reinterpret_cast<float4 *>(p), r);
#else
    (void)streaming;
    *reinterpret_cast<Vec<float, 4> *>(p) = val;
#endif
}

// Posterior of the lanes in `mask` only.  A PARKED lane (finished, waiting for its sector group to be refilled) keeps
// its posterior column until it is shipped: neighbours in the same thread's lane vector must not overwrite it.
template <typename T, int VEC>
QR_HD void store_post_lanes(T *p, const Vec<T, VEC> &v, uint32_t mask, uint64_t pol)
{
    if (mask == (1u << VEC) - 1u) {
        st_pol(reinterpret_cast<Vec<T, VEC> *>(p), v, pol);
    } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
            if (mask >> k & 1) p[k] = v.v[k];
    }
}

// pointers of one tile
template <typename T>
struct TileView {
    const T *c_old;
    T *c_new;
    const T *llr;
    const uint8_t *synd;
    T *post;
    const int32_t *var_slot;
    int32_t tl;
};

template <typename T>
QR_HD TileView<T> tile_view(const FusedParams<T> &F, int cur, int32_t tile)
{
    TileView<T> V;
    V.tl = F.tl;
    V.c_old = F.c2v[cur] + (int64_t)tile * F.P.E * F.tl;
    V.c_new = F.c2v[cur ^ 1] + (int64_t)tile * F.P.E * F.tl;
    V.llr = F.P.llr + (int64_t)tile * F.P.N * F.tl;
    V.synd = F.P.synd + (int64_t)tile * F.P.C * F.tl;
    V.post = F.post ? F.post + (int64_t)tile * F.P.N * F.tl : nullptr;
    V.var_slot = F.P.var_slot;
    return V;
}

// per-thread view of its VEC lanes of a tile, single-buffered state (the tile pipeline orders writer and readers)
template <typename T, int VEC>
QR_HD LaneInfo<VEC> load_tile_lanes(const FusedParams<T> &F, int32_t jv, int32_t minfin)
{
    LaneInfo<VEC> L = load_lane_info<T, VEC>(F.P, 0, jv);
    L.wpost = 0;
    if (F.post) {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
            if ((L.active >> k & 1) && stores_post(L.iter[k], F.P.maxiter, minfin)) L.wpost |= 1u << k;
    }
    return L;
}

// posterior and own message of one edge for the thread's lanes, variable of degree > 3 (extended record):
// one dependent slot lookup per message; the minority path of irregular graphs
template <typename T, int VEC>
QR_HD void fused_edge_ext(const TileView<T> &V, const LaneInfo<VEC> &L, int32_t lt, const Nbr4 &q, uint64_t pol_ld,
                          Vec<T, VEC> &post, Vec<T, VEC> &mine)
{
    using VT = Vec<T, VEC>;
    const int32_t tl = V.tl;
    post = ld_pol(reinterpret_cast<const VT *>(V.llr + (int64_t)nbr_var(q) * tl + lt), pol_ld);
    const int dv = q.n1, own = q.n2;
    for (int j = 0; j < dv; ++j) {
        const int32_t s = V.var_slot[q.n0 + j];
        const VT c = ld_pol(reinterpret_cast<const VT *>(V.c_old + (int64_t)s * tl + lt), pol_ld);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const T cv = (L.fresh >> k & 1) ? (T)0 : c.v[k];       // first half-iteration: c2v == 0 (decoder.pyx:408)
            post.v[k] = post.v[k] + cv;                            // decoder.pyx:291-293, ascending edge id
            if (j == own) mine.v[k] = cv;
        }
    }
}

// FUSED item: internal check `ci` (first CSR slot slot0, degree D) for the thread's VEC lanes at
// offset `lt` inside the tile.  q[i] = neighbour record of slot slot0 + i.  Returns, per lane, 1 if
// this check is NOT satisfied by the posteriors (decoder.pyx:235-257).
// VDEG = 3: every variable of the graph has degree 3 (no extended records, no degree tests: the config-2 kernel).
template <typename T, int VEC, int D, bool ANYFRESH = true, int VDEG = 0>
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
    // the 4 D row loads go out in batches of HB edges (fewer registers held by loads in flight than with all D
    // at once, so two CTAs fit an SM; measured best of 3 / 4 / 5 / 6 on B200)
    constexpr int HB = FUSED_BATCH_EDGES < D ? FUSED_BATCH_EDGES : D;
#pragma unroll
    for (int i0 = 0; i0 < D; i0 += HB) {
        VT m[HB][3], ch[HB];
#pragma unroll
        for (int j = 0; j < HB; ++j) {
            const int i = i0 + j;
            if (i < D && (VDEG == 3 || !nbr_ext(q[i]))) {
                ch[j] = ld_pol(reinterpret_cast<const VT *>(V.llr + (int64_t)nbr_var(q[i]) * tl + lt), pol_ld);
                // (a slot id of -1 = the variable has fewer than three edges: row 0 is loaded and not used)
                m[j][0] = ld_pol(reinterpret_cast<const VT *>(V.c_old + (int64_t)((VDEG != 3 && q[i].n0 < 0) ? 0 : q[i].n0) * tl + lt), pol_ld);
                m[j][1] = ld_pol(reinterpret_cast<const VT *>(V.c_old + (int64_t)((VDEG != 3 && q[i].n1 < 0) ? 0 : q[i].n1) * tl + lt), pol_ld);
                m[j][2] = ld_pol(reinterpret_cast<const VT *>(V.c_old + (int64_t)((VDEG != 3 && q[i].n2 < 0) ? 0 : q[i].n2) * tl + lt), pol_ld);
            }
        }
#pragma unroll
        for (int j = 0; j < HB; ++j) {
            const int i = i0 + j;
            if (i < D) {
                VT pv;
                if (VDEG != 3 && nbr_ext(q[i])) {
                    VT mine;
                    fused_edge_ext<T, VEC>(V, L, lt, q[i], pol_ld, pv, mine);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) {
                        par ^= (uint32_t)(pv.v[k] < (T)0) << k;
                        x[i].v[k] = pv.v[k] - mine.v[k];
                    }
                } else {
                    const int own = (int)(((uint32_t)q[i].vp >> 28) & 3u), dv = VDEG == 3 ? 3 : (int)((uint32_t)q[i].vp >> 30);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) {
                        const bool fresh = ANYFRESH && (L.fresh >> k & 1) != 0;   // first half-iteration: c2v == 0 (decoder.pyx:408)
                        const T c0 = fresh ? (T)0 : m[j][0].v[k];
                        const T c1 = fresh ? (T)0 : m[j][1].v[k];
                        const T c2 = fresh ? (T)0 : m[j][2].v[k];
                        T post = ch[j].v[k] + c0;                            // decoder.pyx:291-293, ascending edge id
                        if (dv > 1) post = post + c1;
                        if (dv > 2) post = post + c2;
                        pv.v[k] = post;
                        par ^= (uint32_t)(post < (T)0) << k;                 // decoder.pyx:244 (strict <)
                        const T mine = own == 0 ? c0 : (own == 1 ? c1 : c2);
                        x[i].v[k] = post - mine;                             // decoder.pyx:295-297
                    }
                }
                // the check holding the variable's FIRST edge keeps the posterior of lanes that may finish now
                if (L.wpost && (VDEG == 3 ? (((uint32_t)q[i].vp >> 28) & 3u) == 0 : nbr_own(q[i]) == 0))
                    store_post_lanes<T, VEC>(V.post + (int64_t)nbr_var(q[i]) * tl + lt, pv, L.wpost, 0);
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

// ---------------------------------------------------------------------------------------------
// LEAN fp32 item (every variable of degree 3).  The float mode has no bit-exactness contract, so the variable side
// is restated for the fewest instructions:
//     v2c = (llr + c[other 1]) + c[other 2]            (decoder.pyx:295-297 without the add-then-subtract of the own message)
//     post = v2c + c[own]                              (decoder.pyx:291-293; only its sign is used, decoder.pyx:244)
// The own messages of a check are its own CSR rows (contiguous: immediate offsets), the record holds what is left --
// the byte offsets, inside a tile, of the variable's LLR row and of its two OTHER message rows, premultiplied, so an
// address is one 32-bit add-to-base -- and no "which of the three is mine" select exists.
struct alignas(16) NbrL {
    uint32_t llr_off, o1_off, o2_off;   // row * TL * sizeof(float)
    uint32_t first;                     // 1: this slot is its variable's first edge (it stores the posterior when asked)
};

template <int D, bool ANYFRESH>
QR_HD uint32_t fused_item_lean(const char *llr_t, const char *cold_t, char *cnew_t, char *post_t,
                               const uint8_t *synd_t, const NbrL *rec, int32_t ci, int32_t slot0, int32_t tl,
                               uint32_t fresh, uint32_t wpost, uint32_t active, bool stream_st)
{
    using VT = Vec<float, 4>;
    const uint32_t rowb = (uint32_t)tl * 4u;
    const Vec<uint8_t, 4> sy = *reinterpret_cast<const Vec<uint8_t, 4> *>(synd_t + (int64_t)ci * tl);
    uint32_t par = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) par |= (uint32_t)(sy.v[k] & 1u) << k;
    const char *own_t = cold_t + (size_t)slot0 * rowb;
    VT x[D];
    constexpr int HB = FUSED_BATCH_EDGES < D ? FUSED_BATCH_EDGES : D;
#pragma unroll
    for (int i0 = 0; i0 < D; i0 += HB) {
        VT ch[HB], m1[HB], m2[HB], mo[HB];
        uint32_t first[HB];
#pragma unroll
        for (int j = 0; j < HB; ++j) {
            const int i = i0 + j;
            if (i < D) {
                NbrL q;
#if defined(__CUDA_ARCH__)
                const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(rec + slot0 + i));
                q.llr_off = raw.x; q.o1_off = raw.y; q.o2_off = raw.z; q.first = raw.w;
#else
                q = rec[slot0 + i];
#endif
                first[j] = q.first;
                ch[j] = ld_row16(llr_t + q.llr_off);
                m1[j] = ld_row16(cold_t + q.o1_off);
                m2[j] = ld_row16(cold_t + q.o2_off);
                mo[j] = ld_row16(own_t + (uint32_t)i * rowb);
            }
        }
#pragma unroll
        for (int j = 0; j < HB; ++j) {
            const int i = i0 + j;
            if (i < D) {
                VT pv;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool fr = ANYFRESH && (fresh >> k & 1) != 0;      // first half-iteration: c2v == 0 (decoder.pyx:408)
                    const float a = fr ? 0.0f : m1[j].v[k], b = fr ? 0.0f : m2[j].v[k], o = fr ? 0.0f : mo[j].v[k];
                    const float v2c = (ch[j].v[k] + a) + b;
                    const float post = v2c + o;
                    pv.v[k] = post;
                    par ^= (uint32_t)(post < 0.0f) << k;                    // decoder.pyx:244 (strict <)
                    x[i].v[k] = v2c;
                }
                if (wpost && first[j]) {
#if defined(__CUDA_ARCH__)
                    const uint32_t off = __ldg(&rec[slot0 + i].llr_off);
#else
                    const uint32_t off = rec[slot0 + i].llr_off;
#endif
                    if (wpost == 0xfu) st_row16(post_t + off, pv, false);
                    else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (wpost >> k & 1) reinterpret_cast<float *>(post_t + off)[k] = pv.v[k];
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float xs[D];
#pragma unroll
        for (int i = 0; i < D; ++i) xs[i] = x[i].v[k];
        MathOf<float>::template run<D, D>(D, xs, (sy.v[k] & 1u) != 0);
#pragma unroll
        for (int i = 0; i < D; ++i) x[i].v[k] = xs[i];
    }
    char *out_t = cnew_t + (size_t)slot0 * rowb;
#pragma unroll
    for (int i = 0; i < D; ++i) st_row16(out_t + (uint32_t)i * rowb, x[i], stream_st);
    return par & active;
}

// lean records from the graph's neighbour records (every variable of degree 3), for lane tiles of `tl` floats
QR_HD NbrL make_lean_record(const Nbr4 &q, int32_t tl)
{
    const uint32_t rowb = (uint32_t)tl * 4u;
    const int own = (int)(((uint32_t)q.vp >> 28) & 3u);
    const int32_t n[3] = {q.n0, q.n1, q.n2};
    NbrL r;
    r.llr_off = (uint32_t)nbr_var(q) * rowb;
    r.o1_off = (uint32_t)n[own == 0 ? 1 : 0] * rowb;
    r.o2_off = (uint32_t)n[own == 2 ? 1 : 2] * rowb;
    r.first = own == 0 ? 1u : 0u;
    return r;
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

// checks k = first, first + stride, ... of one degree bin (latency is covered by the other resident warps)
template <typename T, int VEC, int D, int VDEG = 0>
QR_HD uint32_t run_fused_bin(const TileView<T> &V, const Nbr4 *nbr, const LaneInfo<VEC> &L, int32_t lt,
                             const CheckBin &bin, int32_t first, int32_t stride, uint64_t pol_ld, uint64_t pol_st)
{
    uint32_t bad = 0;
    for (int32_t k = first; k < bin.count; k += stride) {
        Nbr4 cur[D];
        load_nbr_row<D>(nbr, bin.slot_begin + k * D, cur);
        // steady state (no lane of the thread on its first half-iteration): no per-element selects
        if (L.fresh) bad |= fused_item<T, VEC, D, true, VDEG>(V, L, lt, bin.chk_begin + k, bin.slot_begin + k * D, cur, pol_ld, pol_st);
        else bad |= fused_item<T, VEC, D, false, VDEG>(V, L, lt, bin.chk_begin + k, bin.slot_begin + k * D, cur, pol_ld, pol_st);
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

// ---- BK: the decision of decoder.pyx:431-436 for one lane after a sweep
struct BkDecision {
    bool fin_ok, fin_fail;
};
QR_HD BkDecision bk_decide(const LaneState &s, int32_t unsat, int32_t maxiter)
{
    BkDecision d{false, false};
    if (s.frame < 0) return d;
    if (unsat == 0) d.fin_ok = true;                  // syndrome test passed on post_t (decoder.pyx:402-405, :431-433)
    else if (s.iter >= maxiter) d.fin_fail = true;    // decoder.pyx:435-436
    return d;
}

// ---- BATCHED REFILLS.  A refill moves COLUMNS of the lane-interleaved arrays: one 32-byte sector per element, i.e.
// 8 lanes' worth of traffic for one lane.  So a finished lane is PARKED (LaneState: frame = -1, retire = the finished
// frame, iter = its post_valid flag, fresh = rounds waited) and its group of 8 lanes -- the lanes that share sectors
// -- is released together: when none of the 8 is running any more, when a parked lane has waited park_rounds, or
// when the batch has no frame left to admit (then nothing is gained by waiting).  `running` / `parked` / `aged`: bit
// per lane of the tile's 32-lane block, after this round's update.  Returns the lanes to release.
QR_HD uint32_t octets_to_release(uint32_t running, uint32_t parked, uint32_t aged, bool no_more_frames)
{
    uint32_t rel = 0;
    for (int o = 0; o < 4; ++o) {
        const uint32_t m = 0xffu << (8 * o);
        if (!(parked & m)) continue;
        if (no_more_frames || !(running & m) || (aged & m)) rel |= parked & m;
    }
    return rel;
}

// ---- PP, one (listed lane, variable) element: ship post = llr + sum c2v[cur] of a retired frame (what the
// two-phase schedule keeps in its post array), bring in the admitted frame's channel LLR.  `cur` = message
// buffer the sweep READ.
template <typename T>
QR_HD void fused_pp_var_elem(const FusedParams<T> &F, int cur, const RefillEntry &e, int32_t tile, int32_t n)
{
    const DecodeParams<T> &P = F.P;
    const int64_t at = ((int64_t)tile * P.N + n) * F.tl + e.lane;
    if (e.retire >= 0 && P.post_out) {
        const int64_t idx = (int64_t)e.retire * P.N + n;
        if (ld_stream(&P.iters[e.retire]) == 0) {
            // A frame that never iterated.  Input already consistent: the reference copies it (decoder.pyx:404), bit
            // for bit when both sides are fp64.  max_iterations == 0: the reference still ran its first variable
            // pass with c2v == 0 (decoder.pyx:420-421), i.e. llr + 0.0 per edge, which turns -0.0 into +0.0.
            const bool copied = *static_cast<const volatile uint8_t *>(&P.success[e.retire]) != 0;
            if (copied && P.llr_in_f64 && P.post_out_f64) {
                static_cast<double *>(P.post_out)[idx] = static_cast<const double *>(P.llr_in)[idx];
            } else {
                double val = (double)load_input_llr<T>(P.llr_in, P.llr_in_f64, idx);
                if (!copied && P.var_ptr[n + 1] > P.var_ptr[n]) val = val + 0.0;
                store_output_llr(P.post_out, P.post_out_f64, idx, val);
            }
        } else if (e.post_valid) {
            store_output_llr(P.post_out, P.post_out_f64, idx, (double)ld_stream(&F.post[at]));
        } else {
            const T *c = F.c2v[cur] + (int64_t)tile * P.E * F.tl + e.lane;
            T acc = ld_stream(&P.llr[at]);
            const int32_t q0 = P.var_ptr[n], q1 = P.var_ptr[n + 1];
            if (F.nbr_lean && q1 - q0 == 3) {
                // float mode, lean item: the stored posterior is the first edge's ((llr + c[e1]) + c[e2]) + c[e0];
                // the rebuilt one takes the same association, so both shipping paths give the same bits
                acc = acc + ld_stream(&c[(int64_t)P.var_slot[q0 + 1] * F.tl]);
                acc = acc + ld_stream(&c[(int64_t)P.var_slot[q0 + 2] * F.tl]);
                acc = acc + ld_stream(&c[(int64_t)P.var_slot[q0] * F.tl]);
            } else {
                for (int32_t q = q0; q < q1; ++q)
                    acc = acc + ld_stream(&c[(int64_t)P.var_slot[q] * F.tl]);    // decoder.pyx:291-293, ascending edge id
            }
            store_output_llr(P.post_out, P.post_out_f64, idx, (double)acc);
        }
    }
    if (e.frame >= 0) P.llr[at] = load_input_llr<T>(P.llr_in, P.llr_in_f64, (int64_t)e.frame * P.N + n);
}

template <typename T>
QR_HD void fused_pp_chk_elem(const FusedParams<T> &F, const RefillEntry &e, int32_t tile, int32_t ci)
{
    const DecodeParams<T> &P = F.P;
    if (e.frame >= 0)
        P.synd[((int64_t)tile * P.C + ci) * F.tl + e.lane] = P.synd_in[(int64_t)e.frame * P.C + P.chk_order[ci]];
}

}  // namespace qr
