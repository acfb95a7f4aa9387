// Per-work-item device functions of the syndrome sum-product decoder
// (reference: qamreconciliation/decoder.pyx:285-298, 322-369, 235-257, 391-436).
//
// DATA LAYOUT (one decoder workspace, `lanes` = L frames resident, L % 32 == 0):
//   c2v  [E][L]  check-to-variable messages, row = CSR slot (edges grouped by check, checks
//                sorted by degree; within a check ascending original edge id), frames innermost
//   post [N][L]  a-posteriori LLRs             llr [N][L]  channel LLRs
//   synd [C][L]  syndrome bytes, row = internal check slot
// A work item is (node, lane-vector): one thread owns VEC consecutive lanes of one node and moves
// them with one 128-bit access per row, so a warp reads/writes whole 128 B lines of a row.
//
// Only ONE message array is stored.  The reference's variable update writes
// v2c[e] = post[v] - c2v[e] (decoder.pyx:295-297); the check phase recomputes exactly that
// expression from post and the previous c2v, so the v2c array never exists in memory and the
// fp64 results are bit-identical to the two-array schedule.
//
// A lane runs one frame at a time.  State per lane (double-buffered by step parity, because the
// variable phase reads the old state while one thread per lane writes the new one):
//   frame  frame id (-1: idle)      iter  completed variable updates
//   fresh  1 until the first variable update of the frame (the lane's c2v column still holds the
//          previous frame's messages and counts as zero)
//   retire frame id whose posteriors still have to be written out of the lane's column (-1: none)
// One schedule step = check phase, variable phase and -- only when a frame finished -- a refill phase:
//   check phase, lane at iteration t: tests the syndrome on post_t (decoder.pyx:235-257) and
//       computes c2v_{t+1} from v2c_t = post_t - c2v_t (decoder.pyx:322-369).
//   variable phase: if the syndrome test passed -> frame done, (success=1, iters=t, post_t)
//       (decoder.pyx:402-405 for t=0, :431-433 for t>0); else if t == max_iterations -> done,
//       (0, max_iterations, post_t) (:435-436); else post_{t+1} = llr + sum c2v_{t+1}
//       (decoder.pyx:291-293).  A finished lane is handed the next frame of the batch at once.
//   refill phase: finished lanes' posterior columns go to the caller's rows, newly assigned frames'
//       LLR / syndrome rows come into the lane's columns (post = llr).  Message columns are never
//       moved or zeroed: converged frames are compacted out by refilling their lane (continuous
//       batching).  Check and variable phases therefore only ever touch the lane-interleaved arrays
//       and run ONE code path whatever the mix of fresh, running and finishing lanes in a thread.
//
// Everything here is __host__ __device__ so tests/emu can run the same code on the CPU.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#include "qr_common.h"

namespace qr {

struct LaneState {
    int32_t frame;
    int32_t iter;
    int32_t fresh;
    int32_t retire;
};

// CTRL_FIN_STEP: index of the last step in which a frame finished (so a refill phase is due)
// CTRL_REFILL_CNT[2]: lanes listed for the refill phase, double-buffered like the lane state
enum : int { CTRL_NEXT_FRAME = 0, CTRL_REMAINING = 1, CTRL_SNAPSHOT = 2, CTRL_FIN_STEP = 3, CTRL_REFILL_CNT = 4,
             CTRL_WORDS = 16 };
constexpr int kMaxLaneTiles = 1 << 15;

template <typename T>
struct DecodeParams {
    // graph (device pointers)
    const CheckBin *bins;
    int32_t n_bins;
    const int32_t *chk_order, *slot_var, *var_ptr, *var_slot;
    const int32_t *var_work;   // variable ids sorted by degree (null: every variable has degree var_deg)
    const CheckBin *var_bins;  // one bin per variable degree over var_work positions, slot lists in vslot_sorted (or null)
    int32_t n_var_bins;
    const int32_t *vslot_sorted;
    int64_t N, C, E;
    int32_t var_deg;  // > 0: every variable has this degree (var_slot row of n starts at n * var_deg)
    // workspace
    int32_t lanes;
    T *c2v, *post, *llr;
    uint8_t *synd;
    LaneState *st[2];
    int32_t *unsat[2];
    // batch
    const void *llr_in;
    int32_t llr_in_f64;
    const uint8_t *synd_in;
    int64_t frames;
    int32_t maxiter;
    uint8_t *success;
    int32_t *iters;
    void *post_out;
    int32_t post_out_f64;
    // control words and counters
    int32_t *ctrl;
    int32_t *work;              // [2][kMaxLaneTiles] row counters of the persistent kernel's work stealing
    int32_t *refill_list;       // [2][lanes] lanes the refill phase has to serve (null: scan all lanes)
    unsigned long long *stats;  // [0] flooding iterations summed over finished frames
};

// ---------------------------------------------------------------------------------------------
// vectors of VEC lanes
template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) Vec {
    T v[VEC];
};

// L2-only (.cg) accesses for data other SMs wrote in the previous phase (messages, posteriors):
// each value is read once per phase, so keeping it out of L1 leaves L1 to the index tables.
template <typename V>
QR_HD V ld_stream(const V *p)
{
#if defined(__CUDA_ARCH__)
    static_assert(sizeof(V) == 4 || sizeof(V) == 8 || sizeof(V) == 16, "vector width");
    V r;
    if constexpr (sizeof(V) == 16) {
        int4 t = __ldcg(reinterpret_cast<const int4 *>(p));
        memcpy(&r, &t, 16);
    } else if constexpr (sizeof(V) == 8) {
        int2 t = __ldcg(reinterpret_cast<const int2 *>(p));
        memcpy(&r, &t, 8);
    } else {
        int t = __ldcg(reinterpret_cast<const int *>(p));
        memcpy(&r, &t, 4);
    }
    return r;
#else
    return *p;
#endif
}
template <typename V>
QR_HD void st_stream(V *p, const V &val)
{
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(V) == 16) {
        int4 t; memcpy(&t, &val, 16);
        __stcg(reinterpret_cast<int4 *>(p), t);
    } else if constexpr (sizeof(V) == 8) {
        int2 t; memcpy(&t, &val, 8);
        __stcg(reinterpret_cast<int2 *>(p), t);
    } else {
        int t; memcpy(&t, &val, 4);
        __stcg(reinterpret_cast<int *>(p), t);
    }
#else
    *p = val;
#endif
}

template <typename T, int VEC>
QR_HD Vec<T, VEC> ld_row(const T *base, int64_t row, int32_t lanes, int32_t l0)
{
    return ld_stream(reinterpret_cast<const Vec<T, VEC> *>(base + row * lanes + l0));
}
template <typename T, int VEC>
QR_HD void st_row(T *base, int64_t row, int32_t lanes, int32_t l0, const Vec<T, VEC> &v)
{
    st_stream(reinterpret_cast<Vec<T, VEC> *>(base + row * lanes + l0), v);
}

template <typename T>
QR_HD T load_input_llr(const void *llr_in, int is_f64, int64_t idx);

template <>
QR_HD double load_input_llr<double>(const void *llr_in, int is_f64, int64_t idx)
{
    return is_f64 ? static_cast<const double *>(llr_in)[idx]
                  : (double)static_cast<const float *>(llr_in)[idx];
}

// fp32 mode saturates what it reads: +-inf or 1e300 (the reference's bare_llr_table sentinel,
// noisemapper.pyx:217-218) would otherwise turn into inf - inf = NaN in the sums.
constexpr float kLlrInClamp = 1.0e30f;
constexpr float kMsgClamp = 80.0f;  // |c2v| cap: exp(-80) is still a normal float

template <>
QR_HD float load_input_llr<float>(const void *llr_in, int is_f64, int64_t idx)
{
    float x = is_f64 ? (float)static_cast<const double *>(llr_in)[idx]
                     : static_cast<const float *>(llr_in)[idx];
    return fminf(fmaxf(x, -kLlrInClamp), kLlrInClamp);
}

QR_HD void store_output_llr(void *out, int is_f64, int64_t idx, double v)
{
    if (is_f64) static_cast<double *>(out)[idx] = v;
    else static_cast<float *>(out)[idx] = (float)v;
}

// ---------------------------------------------------------------------------------------------
// check-node arithmetic
//
// MathRef: the reference's box-plus, evaluated exactly as written at decoder.pyx:41-45,
//   sgn(a)sgn(b)min(|a|,|b|) + log(1+exp(-|a+b|)) - log(1+exp(-|a-b|))   (left to right),
// in the forward/backward recursion of decoder.pyx:341-367.
struct MathRef {
    using T = double;
    static QR_HD int sgn(double x) { return (0.0 < x) - (x < 0.0); }
    static QR_HD double box_plus(double a, double b)
    {
        double fa = fabs(a), fb = fabs(b);
        double mn = (fb < fa) ? fb : fa;
        double r = (double)(sgn(a) * sgn(b)) * mn;
        r = r + log(1 + exp(-fabs(a + b)));
        r = r - log(1 + exp(-fabs(a - b)));
        return r;
    }
    // x[0..deg) holds v2c on entry, c2v on exit; bwd is scratch of >= deg entries
    template <int UNR>
    static QR_HD void check_node(int deg, double *x, double *bwd, bool synd)
    {
        bwd[deg - 1] = x[deg - 1];
#pragma unroll UNR
        for (int i = deg - 2; i >= 1; --i) bwd[i] = box_plus(bwd[i + 1], x[i]);
        const double pre = synd ? -1.0 : 1.0;
        double fwd = x[0];
        x[0] = pre * bwd[1];
#pragma unroll UNR
        for (int i = 1; i < deg - 1; ++i) {
            double out = pre * box_plus(fwd, bwd[i + 1]);
            fwd = box_plus(fwd, x[i]);
            x[i] = out;
        }
        x[deg - 1] = pre * fwd;
    }
};

// MathFast: float.  A message of magnitude m is carried as u = exp(-m) in [0,1]; the box-plus of
// magnitudes is then u_a (+) u_b = (u_a + u_b) / (1 + u_a u_b)  [tanh(m/2) = (1-u)/(1+u)], which has
// no cancellation at large m.  Partial results stay as fractions n/d so the recursion is FMAs only;
// one ex2 per incoming and two lg2 per outgoing message are the only special-function ops.
// Signs: product of the other edges' sign bits, times the syndrome bit (decoder.pyx:358-367).
struct MathFast {
    using T = float;
    static QR_HD float ex2(float x)
    {
#if defined(__CUDA_ARCH__)
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
#else
        return exp2f(x);
#endif
    }
    static QR_HD float lg2(float x)
    {
#if defined(__CUDA_ARCH__)
        float r;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
#else
        return log2f(x);
#endif
    }
    static QR_HD uint32_t bits(float x)
    {
        uint32_t b;
        memcpy(&b, &x, 4);
        return b;
    }
    static QR_HD float emit(float n, float d, uint32_t sign_bit)
    {
        float mag = 0.69314718056f * (lg2(d) - lg2(n));
        mag = fminf(fmaxf(mag, 0.0f), kMsgClamp);  // also maps NaN/inf to the clamp
        uint32_t b = bits(mag) | sign_bit;
        float r;
        memcpy(&r, &b, 4);
        return r;
    }
    // x[0..deg): v2c in, c2v out.  u/nb/db: scratch of >= deg entries each.
    template <int UNR>
    static QR_HD void check_node(int deg, float *x, float *u, float *nb, float *db, bool synd)
    {
        uint32_t total = synd ? 0x80000000u : 0u;
#pragma unroll UNR
        for (int i = 0; i < deg; ++i) {
            total ^= bits(x[i]) & 0x80000000u;
            u[i] = ex2(-1.44269504089f * fabsf(x[i]));
        }
        // backward fractions B_i = u_i (+) ... (+) u_{deg-1} = nb[i] / db[i]
        nb[deg - 1] = u[deg - 1];
        db[deg - 1] = 1.0f;
#pragma unroll UNR
        for (int i = deg - 2; i >= 1; --i) {
            nb[i] = fmaf(u[i], db[i + 1], nb[i + 1]);
            db[i] = fmaf(u[i], nb[i + 1], db[i + 1]);
            if (UNR == 1 && db[i] > 1.0e18f) {  // run-time high degrees only: renormalise
                nb[i] *= 1.0e-18f;
                db[i] *= 1.0e-18f;
            }
        }
        float nf = u[0], df = 1.0f;  // forward fraction F_{i-1}
        x[0] = emit(nb[1], db[1], total ^ (bits(x[0]) & 0x80000000u));
#pragma unroll UNR
        for (int i = 1; i < deg - 1; ++i) {
            const float n = fmaf(nf, db[i + 1], nb[i + 1] * df);
            const float d = fmaf(nf, nb[i + 1], df * db[i + 1]);
            x[i] = emit(n, d, total ^ (bits(x[i]) & 0x80000000u));
            const float nf2 = fmaf(u[i], df, nf);
            df = fmaf(u[i], nf, df);
            nf = nf2;
            if (UNR == 1 && df > 1.0e18f) {
                nf *= 1.0e-18f;
                df *= 1.0e-18f;
            }
        }
        x[deg - 1] = emit(nf, df, total ^ (bits(x[deg - 1]) & 0x80000000u));
    }
};

template <typename T>
struct MathOf;
template <>
struct MathOf<double> {
    using M = MathRef;
    template <int CAP, int UNR>
    static QR_HD void run(int deg, double *x, bool synd)
    {
        double bwd[CAP];
        MathRef::check_node<UNR>(deg, x, bwd, synd);
    }
};
template <>
struct MathOf<float> {
    using M = MathFast;
    template <int CAP, int UNR>
    static QR_HD void run(int deg, float *x, bool synd)
    {
        float u[CAP], nb[CAP], db[CAP];
        MathFast::check_node<UNR>(deg, x, u, nb, db, synd);
    }
};

// ---------------------------------------------------------------------------------------------
// per-thread view of its VEC lanes, loaded once per phase
template <int VEC>
struct LaneInfo {
    int32_t l0;
    int32_t frame[VEC];
    int32_t iter[VEC];
    uint32_t active;  // bit k: lane runs a frame
    uint32_t fresh;   // bit k: no variable update yet for this frame: its c2v column counts as zero
    // variable phase decisions
    uint32_t fin_ok, fin_fail, upd;
    uint32_t wpost;   // fused schedule: lanes whose posterior is stored this step (they may finish in it)
};

template <typename T, int VEC>
QR_HD LaneInfo<VEC> load_lane_info(const DecodeParams<T> &P, int cur, int32_t jv)
{
    LaneInfo<VEC> L;
    L.l0 = jv * VEC;
    L.active = L.fresh = L.fin_ok = L.fin_fail = L.upd = L.wpost = 0;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const LaneState s = ld_stream(&P.st[cur][L.l0 + k]);
        L.frame[k] = s.frame;
        L.iter[k] = s.iter;
        if (s.frame >= 0) {
            L.active |= 1u << k;
            if (s.fresh) L.fresh |= 1u << k;
        }
    }
    return L;
}

template <typename T, int VEC>
QR_HD void decide_lanes(const DecodeParams<T> &P, int cur, LaneInfo<VEC> &L)
{
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        if (!(L.active >> k & 1)) continue;
        if (ld_stream(&P.unsat[cur][L.l0 + k]) == 0) L.fin_ok |= 1u << k;
        else if (L.iter[k] >= P.maxiter) L.fin_fail |= 1u << k;
        else L.upd |= 1u << k;
    }
}

template <int D>
QR_HD void load_index_row(const int32_t *__restrict__ tab, int32_t first, int32_t (&v)[D])
{
#pragma unroll
    for (int i = 0; i < D; ++i) v[i] = tab[first + i];
}

// ---------------------------------------------------------------------------------------------
// CHECK PHASE item: internal check `ci` (first CSR slot slot0) for the thread's lanes, given the
// variable ids of its edges.  All 2*deg 128-bit row loads are issued back to back before anything
// consumes them.  Returns a bit per lane: 1 = this check is NOT satisfied by the lane's posteriors.
// D > 0: compile-time degree (registers); D == 0: run-time degree <= kMaxCheckDegree.
template <typename T, int VEC, int D>
QR_HD uint32_t check_item(const DecodeParams<T> &P, const LaneInfo<VEC> &L, int32_t ci, int32_t slot0,
                          int32_t deg_rt, const int32_t *v)
{
    constexpr int CAP = D > 0 ? D : kMaxCheckDegree;
    constexpr int UNR = D > 0 ? D : 1;  // full unroll for compile-time degrees, none otherwise
    const int deg = D > 0 ? D : deg_rt;
    const int32_t lanes = P.lanes;
    Vec<T, VEC> pv[CAP], x[CAP];
#pragma unroll UNR
    for (int i = 0; i < deg; ++i) pv[i] = ld_row<T, VEC>(P.post, D > 0 ? v[i] : P.slot_var[slot0 + i], lanes, L.l0);
#pragma unroll UNR
    for (int i = 0; i < deg; ++i) x[i] = ld_row<T, VEC>(P.c2v, slot0 + i, lanes, L.l0);
    const Vec<uint8_t, VEC> sy = *reinterpret_cast<const Vec<uint8_t, VEC> *>(P.synd + (int64_t)ci * lanes + L.l0);
    uint32_t par = 0;   // bit k: syndrome XOR parity of the negative posteriors
#pragma unroll
    for (int k = 0; k < VEC; ++k) par |= (uint32_t)(sy.v[k] & 1u) << k;
#pragma unroll UNR
    for (int i = 0; i < deg; ++i) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            par ^= (uint32_t)(pv[i].v[k] < (T)0) << k;                      // decoder.pyx:244 (strict <)
            const T old = (L.fresh >> k & 1) ? (T)0 : x[i].v[k];            // first half-iteration: c2v == 0 (:408)
            x[i].v[k] = pv[i].v[k] - old;                                   // decoder.pyx:295-297
        }
    }
    // node update, one lane at a time (keeps only one lane's scratch live)
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        T xs[CAP];
#pragma unroll UNR
        for (int i = 0; i < deg; ++i) xs[i] = x[i].v[k];
        MathOf<T>::template run<CAP, UNR>(deg, xs, (sy.v[k] & 1u) != 0);
#pragma unroll UNR
        for (int i = 0; i < deg; ++i) x[i].v[k] = xs[i];
    }
#pragma unroll UNR
    for (int i = 0; i < deg; ++i) st_row<T, VEC>(P.c2v, slot0 + i, lanes, L.l0, x[i]);
    return par & L.active;
}

// syndrome byte semantics: the reference XORs the whole byte and tests (parity ^ 1) != 0
// (decoder.pyx:243-249); for the 0/1 bytes every caller passes that is the low bit, which is
// what the kernels use.

// All checks k = first, first+stride, ... of one degree bin, for the thread's lanes.  The index row of
// the NEXT check is fetched while the current one computes.
template <typename T, int VEC, int D>
QR_HD uint32_t run_check_bin(const DecodeParams<T> &P, const LaneInfo<VEC> &L, const CheckBin &bin,
                             int32_t first, int32_t stride)
{
    uint32_t bad = 0;
    if constexpr (D > 0) {
        int32_t cur[D], nxt[D];
        if (first < bin.count) load_index_row<D>(P.slot_var, bin.slot_begin + first * D, cur);
        for (int32_t k = first; k < bin.count; k += stride) {
            const int32_t kn = k + stride;
            if (kn < bin.count) load_index_row<D>(P.slot_var, bin.slot_begin + kn * D, nxt);
            bad |= check_item<T, VEC, D>(P, L, bin.chk_begin + k, bin.slot_begin + k * D, D, cur);
#pragma unroll
            for (int i = 0; i < D; ++i) cur[i] = nxt[i];
        }
    } else {
        for (int32_t k = first; k < bin.count; k += stride)
            bad |= check_item<T, VEC, 0>(P, L, bin.chk_begin + k, bin.slot_begin + k * bin.degree, bin.degree,
                                         nullptr);
    }
    return bad;
}

// ---------------------------------------------------------------------------------------------
// VARIABLE PHASE.  post_{t+1}[n] = llr[n] + sum of c2v over the variable's edges in ascending edge id
// (decoder.pyx:291-293) for the lanes that iterate on; lanes that just finished keep post_t (the refill
// phase ships it), idle lanes are don't-care.  MASKED = some lane of the vector does not update.
template <typename T, int VEC, int DV>
QR_HD void var_item_fixed(const DecodeParams<T> &P, const LaneInfo<VEC> &L, int32_t n, const int32_t (&slot)[DV],
                          bool masked)
{
    const int32_t lanes = P.lanes;
    Vec<T, VEC> acc = ld_row<T, VEC>(P.llr, n, lanes, L.l0);
    Vec<T, VEC> m[DV], old;
#pragma unroll
    for (int i = 0; i < DV; ++i) m[i] = ld_row<T, VEC>(P.c2v, slot[i], lanes, L.l0);
    if (masked) old = ld_row<T, VEC>(P.post, n, lanes, L.l0);
#pragma unroll
    for (int i = 0; i < DV; ++i) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc.v[k] = acc.v[k] + m[i].v[k];
    }
    if (masked) {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
            if (!(L.upd >> k & 1)) acc.v[k] = old.v[k];
    }
    st_row<T, VEC>(P.post, n, lanes, L.l0, acc);
}

// any variable degree (run-time loop over the variable's slot list)
template <typename T, int VEC>
QR_HD void var_item(const DecodeParams<T> &P, const LaneInfo<VEC> &L, int32_t n)
{
    const int32_t lanes = P.lanes;
    Vec<T, VEC> acc = ld_row<T, VEC>(P.llr, n, lanes, L.l0);
    const bool masked = L.upd != (1u << VEC) - 1u;
    Vec<T, VEC> old;
    if (masked) old = ld_row<T, VEC>(P.post, n, lanes, L.l0);
    const int32_t q0 = P.var_ptr[n], q1 = P.var_ptr[n + 1];
    for (int32_t q = q0; q < q1; ++q) {
        const Vec<T, VEC> m = ld_row<T, VEC>(P.c2v, P.var_slot[q], lanes, L.l0);
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc.v[k] = acc.v[k] + m.v[k];
    }
    if (masked) {
#pragma unroll
        for (int k = 0; k < VEC; ++k)
            if (!(L.upd >> k & 1)) acc.v[k] = old.v[k];
    }
    st_row<T, VEC>(P.post, n, lanes, L.l0, acc);
}

template <typename T, int VEC, int DV, int U>
QR_HD void run_var_fixed(const DecodeParams<T> &P, const LaneInfo<VEC> &L, int32_t first, int32_t stride,
                         int32_t n_end, bool masked)
{
    // U variables per trip (n, n+stride, ...): U*(DV+1) independent row loads in flight per thread;
    // the slot rows of the next trip are fetched while this one is summed.
    const int32_t N = n_end;
    int32_t cur[U][DV], nxt[U][DV];
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (first + u * stride < N) load_index_row<DV>(P.var_slot, (first + u * stride) * DV, cur[u]);
    for (int32_t n = first; n < N; n += U * stride) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int32_t nn = n + (U + u) * stride;
            if (nn < N) load_index_row<DV>(P.var_slot, nn * DV, nxt[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (n + u * stride < N) var_item_fixed<T, VEC, DV>(P, L, n + u * stride, cur[u], masked);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int i = 0; i < DV; ++i) cur[u][i] = nxt[u][i];
    }
}

// Variables at positions first, first + stride, ... < count of one degree bin (irregular graphs): the variable id and
// the DV slot ids of the NEXT item are fetched while the current one is summed, all DV + 1 (+1) row loads of an item
// are issued back to back -- the same pipeline as the regular case, which the degree-sorted order makes possible.
template <typename T, int VEC, int DV, int U>
QR_HD void run_var_bin(const DecodeParams<T> &P, const LaneInfo<VEC> &L, const CheckBin &bin, int32_t first,
                       int32_t stride, bool masked)
{
    // U items per trip (positions k, k + stride, ...): U (DV + 1) independent row loads in flight per thread
    int32_t cur[U][DV], nxt[U][DV], n_cur[U], n_nxt[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        n_cur[u] = n_nxt[u] = 0;
        const int32_t k = first + u * stride;
        if (k < bin.count) {
            n_cur[u] = P.var_work[bin.chk_begin + k];
            load_index_row<DV>(P.vslot_sorted, bin.slot_begin + k * DV, cur[u]);
        }
    }
    for (int32_t k0 = first; k0 < bin.count; k0 += U * stride) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int32_t kn = k0 + (U + u) * stride;
            if (kn < bin.count) {
                n_nxt[u] = P.var_work[bin.chk_begin + kn];
                load_index_row<DV>(P.vslot_sorted, bin.slot_begin + kn * DV, nxt[u]);
            }
        }
        if constexpr (U == 2) {
            // both items' rows loaded before either is summed
            const bool two = k0 + stride < bin.count;
            const int32_t lanes = P.lanes;
            Vec<T, VEC> a0 = ld_row<T, VEC>(P.llr, n_cur[0], lanes, L.l0), a1, m0[DV], m1[DV], o0, o1;
#pragma unroll
            for (int i = 0; i < DV; ++i) m0[i] = ld_row<T, VEC>(P.c2v, cur[0][i], lanes, L.l0);
            if (masked) o0 = ld_row<T, VEC>(P.post, n_cur[0], lanes, L.l0);
            if (two) {
                a1 = ld_row<T, VEC>(P.llr, n_cur[1], lanes, L.l0);
#pragma unroll
                for (int i = 0; i < DV; ++i) m1[i] = ld_row<T, VEC>(P.c2v, cur[1][i], lanes, L.l0);
                if (masked) o1 = ld_row<T, VEC>(P.post, n_cur[1], lanes, L.l0);
            }
#pragma unroll
            for (int i = 0; i < DV; ++i) {
#pragma unroll
                for (int kk = 0; kk < VEC; ++kk) a0.v[kk] = a0.v[kk] + m0[i].v[kk];
            }
            if (masked) {
#pragma unroll
                for (int kk = 0; kk < VEC; ++kk)
                    if (!(L.upd >> kk & 1)) a0.v[kk] = o0.v[kk];
            }
            st_row<T, VEC>(P.post, n_cur[0], lanes, L.l0, a0);
            if (two) {
#pragma unroll
                for (int i = 0; i < DV; ++i) {
#pragma unroll
                    for (int kk = 0; kk < VEC; ++kk) a1.v[kk] = a1.v[kk] + m1[i].v[kk];
                }
                if (masked) {
#pragma unroll
                    for (int kk = 0; kk < VEC; ++kk)
                        if (!(L.upd >> kk & 1)) a1.v[kk] = o1.v[kk];
                }
                st_row<T, VEC>(P.post, n_cur[1], lanes, L.l0, a1);
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (k0 + u * stride < bin.count) var_item_fixed<T, VEC, DV>(P, L, n_cur[u], cur[u], masked);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            n_cur[u] = n_nxt[u];
#pragma unroll
            for (int i = 0; i < DV; ++i) cur[u][i] = nxt[u][i];
        }
    }
}

// a bin of any degree: the slot list in chunks of 8 rows, summed in list order (ascending edge id, decoder.pyx:291-293)
template <typename T, int VEC>
QR_HD void run_var_bin_any(const DecodeParams<T> &P, const LaneInfo<VEC> &L, const CheckBin &bin, int32_t first,
                           int32_t stride, bool masked)
{
    const int32_t lanes = P.lanes, deg = bin.degree;
    for (int32_t k = first; k < bin.count; k += stride) {
        const int32_t n = P.var_work[bin.chk_begin + k];
        const int32_t *sl = P.vslot_sorted + bin.slot_begin + k * deg;
        Vec<T, VEC> acc = ld_row<T, VEC>(P.llr, n, lanes, L.l0), old;
        if (masked) old = ld_row<T, VEC>(P.post, n, lanes, L.l0);
        for (int32_t j0 = 0; j0 < deg; j0 += 8) {
            Vec<T, VEC> m[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j0 + j < deg) m[j] = ld_row<T, VEC>(P.c2v, sl[j0 + j], lanes, L.l0);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j0 + j < deg) {
#pragma unroll
                    for (int kk = 0; kk < VEC; ++kk) acc.v[kk] = acc.v[kk] + m[j].v[kk];
                }
        }
        if (masked) {
#pragma unroll
            for (int kk = 0; kk < VEC; ++kk)
                if (!(L.upd >> kk & 1)) acc.v[kk] = old.v[kk];
        }
        st_row<T, VEC>(P.post, n, lanes, L.l0, acc);
    }
}

// positions [p0, p1) of the degree-sorted variable order, rows first, first + stride, ... of that range
template <typename T, int VEC>
QR_HD void run_var_binned(const DecodeParams<T> &P, const LaneInfo<VEC> &L, int32_t p0, int32_t p1, int32_t first,
                          int32_t stride)
{
    if (!L.upd) return;
    const bool masked = L.upd != (1u << VEC) - 1u;
    for (int32_t b = 0; b < P.n_var_bins; ++b) {
        const CheckBin bin = P.var_bins[b];
        const int32_t lo = p0 > bin.chk_begin ? p0 : bin.chk_begin;
        const int32_t hi = p1 < bin.chk_begin + bin.count ? p1 : bin.chk_begin + bin.count;
        if (lo >= hi) continue;
        const CheckBin sub{bin.degree, lo, hi - lo, bin.slot_begin + (lo - bin.chk_begin) * bin.degree};
        switch (bin.degree) {
        case 1: run_var_bin<T, VEC, 1, 2>(P, L, sub, first, stride, masked); break;
        case 2: run_var_bin<T, VEC, 2, 2>(P, L, sub, first, stride, masked); break;
        case 3: run_var_bin<T, VEC, 3, 2>(P, L, sub, first, stride, masked); break;
        case 4: run_var_bin<T, VEC, 4, 2>(P, L, sub, first, stride, masked); break;
        case 5: run_var_bin<T, VEC, 5, 2>(P, L, sub, first, stride, masked); break;
        case 6: run_var_bin<T, VEC, 6, 1>(P, L, sub, first, stride, masked); break;
        case 8: run_var_bin<T, VEC, 8, 1>(P, L, sub, first, stride, masked); break;
        default: run_var_bin_any<T, VEC>(P, L, sub, first, stride, masked); break;
        }
    }
}

// All variables n = first, first+stride, ... < n_end for the thread's lanes (decisions already in L).
// UNROLLED: compile the per-degree unrolled variants for irregular variable degrees (kept out of the kernels
// specialised for one check degree, whose register allocation they would disturb).
template <typename T, int VEC, bool UNROLLED = true>
QR_HD void run_var_range(const DecodeParams<T> &P, const LaneInfo<VEC> &L, int32_t first, int32_t stride,
                         int32_t n_end)
{
    if (!L.upd) return;
    constexpr int U = 2;  // 4 was measured slower on B200 (register spills in the persistent kernel)
    // the common regular degree gets an unrolled, index-prefetching loop; anything else the generic one
    if (P.var_deg == 3) {
        run_var_fixed<T, VEC, 3, U>(P, L, first, stride, n_end, L.upd != (1u << VEC) - 1u);
        return;
    }
    if constexpr (UNROLLED) {
        if (P.var_bins) {          // irregular graph: degree bins over the sorted order (positions 0 .. n_end)
            run_var_binned<T, VEC>(P, L, 0, n_end, first, stride);
            return;
        }
    }
    // irregular variable degrees: positions first, first + stride, ... of the degree-sorted work list.  Within a
    // warp the degree is (almost always) the same, so the unrolled variants issue all their row loads at once
    // instead of one dependent load per edge.
    if constexpr (!UNROLLED) {
        for (int32_t pos = first; pos < n_end; pos += stride) var_item<T, VEC>(P, L, P.var_work ? P.var_work[pos] : pos);
        return;
    }
    const bool masked = L.upd != (1u << VEC) - 1u;
    for (int32_t pos = first; pos < n_end; pos += stride) {
        const int32_t n = P.var_work ? P.var_work[pos] : pos;
        const int32_t q0 = P.var_ptr[n], deg = P.var_ptr[n + 1] - q0;
        switch (deg) {
        case 2: { int32_t sl[2]; load_index_row<2>(P.var_slot, q0, sl); var_item_fixed<T, VEC, 2>(P, L, n, sl, masked); break; }
        case 3: { int32_t sl[3]; load_index_row<3>(P.var_slot, q0, sl); var_item_fixed<T, VEC, 3>(P, L, n, sl, masked); break; }
        case 4: { int32_t sl[4]; load_index_row<4>(P.var_slot, q0, sl); var_item_fixed<T, VEC, 4>(P, L, n, sl, masked); break; }
        case 5: { int32_t sl[5]; load_index_row<5>(P.var_slot, q0, sl); var_item_fixed<T, VEC, 5>(P, L, n, sl, masked); break; }
        case 6: { int32_t sl[6]; load_index_row<6>(P.var_slot, q0, sl); var_item_fixed<T, VEC, 6>(P, L, n, sl, masked); break; }
        case 8: { int32_t sl[8]; load_index_row<8>(P.var_slot, q0, sl); var_item_fixed<T, VEC, 8>(P, L, n, sl, masked); break; }
        default: var_item<T, VEC>(P, L, n); break;
        }
    }
}


// One thread per lane-vector advances the lane state machine, once per step.
#if defined(__CUDA_ARCH__)
#define QR_ATOMIC_ADD_I32(p, v) atomicAdd((p), (v))
#define QR_ATOMIC_ADD_U64(p, v) atomicAdd((p), (v))
#else
#define QR_ATOMIC_ADD_I32(p, v) ([&] { int32_t _o = *(p); *(p) += (v); return _o; }())
#define QR_ATOMIC_ADD_U64(p, v) ([&] { unsigned long long _o = *(p); *(p) += (v); return _o; }())
#endif

template <typename T, int VEC>
QR_HD void bookkeep_lanes(const DecodeParams<T> &P, int cur, int32_t step, const LaneInfo<VEC> &L)
{
    const int nxt = cur ^ 1;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        const int32_t lane = L.l0 + k;
        LaneState s;
        s.frame = L.frame[k];
        s.iter = L.iter[k];
        s.fresh = (L.fresh >> k) & 1;
        s.retire = -1;
        if ((L.fin_ok | L.fin_fail) >> k & 1) {
            const bool ok = (L.fin_ok >> k & 1) != 0;
            P.success[s.frame] = ok ? 1 : 0;
            P.iters[s.frame] = ok ? s.iter : P.maxiter;
            QR_ATOMIC_ADD_U64(&P.stats[0], (unsigned long long)s.iter);
            QR_ATOMIC_ADD_I32(&P.ctrl[CTRL_REMAINING], -1);
            s.retire = s.frame;                      // the refill phase ships this frame's posteriors
            P.ctrl[CTRL_FIN_STEP] = step;            // ... and is due after this step
            if (P.refill_list) {
                const int32_t slot = QR_ATOMIC_ADD_I32(&P.ctrl[CTRL_REFILL_CNT + nxt], 1);
                P.refill_list[(int64_t)nxt * P.lanes + slot] = lane;
            }
            const int32_t nf = QR_ATOMIC_ADD_I32(&P.ctrl[CTRL_NEXT_FRAME], 1);
            if ((int64_t)nf < P.frames) { s.frame = nf; s.iter = 0; s.fresh = 1; }
            else { s.frame = -1; s.iter = 0; s.fresh = 0; }
        } else if (L.upd >> k & 1) {
            s.iter += 1;
            s.fresh = 0;
        }
        P.st[nxt][lane] = s;
        P.unsat[nxt][lane] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// REFILL PHASE, one (lane, variable) element: ship the finished frame's posterior, bring in the new
// frame's channel LLR.  `s` is the lane's state as the variable phase of this step left it.
template <typename T>
QR_HD void refill_var_elem(const DecodeParams<T> &P, const LaneState &s, int32_t lane, int32_t n)
{
    const int64_t at = (int64_t)n * P.lanes + lane;
    if (s.retire >= 0 && P.post_out) {
        const int64_t idx = (int64_t)s.retire * P.N + n;
        if (ld_stream(&P.iters[s.retire]) == 0) {   // (L2 loads: written by another SM in the phase before)
            // A frame that never iterated.  Input already consistent: the reference copies it
            // (decoder.pyx:404), bit for bit when both sides are fp64.  max_iterations == 0: the
            // reference still ran its first variable pass with c2v == 0 (decoder.pyx:420-421), i.e.
            // llr + 0.0 per edge, which turns -0.0 into +0.0.
            const bool copied = *static_cast<const volatile uint8_t *>(&P.success[s.retire]) != 0;
            if (copied && P.llr_in_f64 && P.post_out_f64) {
                static_cast<double *>(P.post_out)[idx] = static_cast<const double *>(P.llr_in)[idx];
            } else {
                double val = (double)load_input_llr<T>(P.llr_in, P.llr_in_f64, idx);
                if (!copied && P.var_ptr[n + 1] > P.var_ptr[n]) val = val + 0.0;
                store_output_llr(P.post_out, P.post_out_f64, idx, val);
            }
        } else {
            store_output_llr(P.post_out, P.post_out_f64, idx, (double)ld_stream(&P.post[at]));
        }
    }
    if (s.frame >= 0 && s.fresh) {
        const T v = load_input_llr<T>(P.llr_in, P.llr_in_f64, (int64_t)s.frame * P.N + n);
        P.llr[at] = v;
        P.post[at] = v;
    }
}

template <typename T>
QR_HD void refill_chk_elem(const DecodeParams<T> &P, const LaneState &s, int32_t lane, int32_t ci)
{
    if (s.frame >= 0 && s.fresh)
        P.synd[(int64_t)ci * P.lanes + lane] = P.synd_in[(int64_t)s.frame * P.C + P.chk_order[ci]];
}

QR_HD bool lane_needs_refill(const LaneState &s) { return s.retire >= 0 || (s.frame >= 0 && s.fresh); }

}  // namespace qr
