// Per-symbol arithmetic of the softening noise mapper / LLR demapper, fp64
// (reference: qamreconciliation/noisemapper.pyx, alphabet.pyx:98-107, sims/reconciliation.pyx:25-51).
// __host__ __device__ so tests/emu can run the same code on the CPU.
#pragma once

#include <cmath>
#include <cstdint>

#include "qr_common.h"

namespace qr {

constexpr int kMaxBps = 8;
constexpr int kMaxOrder = 1 << kMaxBps;

// What a kernel needs of the mapper, by value (pointers are device pointers).
struct MapperView {
    int32_t order, bps;
    double noise_var, sigma, s2;  // s2 = sqrt(2) * sigma, as at noisemapper.pyx:66-67
    const double *constellation;  // [order]
    const double *thresholds;     // [order+1]
    const double *probabilities;  // [order]
    const uint8_t *sign_config;   // [order]   ctor sign_config: g_inv_search / demap_lappr
    const uint8_t *sign_g;        // [order]   sign rule of g / map_noise (== sign_config except for the FlipSign subclasses)
    const double *FY_thr;         // [order+1]  F_Y_thresholds
    const double *delta;          // [order]    delta_F_Y
    const double *bare;           // [order*bps]
    // F_Y sampled on a uniform grid y = inv_y0 + j * inv_h, j < inv_n: starting points of the fast inverse
    const double *inv_tab;
    const double *inv_pdf;
    int32_t inv_n;
    double inv_y0, inv_h;
    const int32_t *inv_jump;      // see InvTable::jump
    int32_t inv_jn;
    const double *inv32_F;        // InvTable32 (global copies, staged into shared memory by k_demap32)
    const float *inv32_f;
    const uint16_t *inv32_jump;
    double inv32_h;
    int32_t uniform;              // constellation points equally spaced (always true for PAMAlphabet): enables E^m G[m]
    int32_t *index_errors;        // device counter of caller-supplied symbol / region indices found out of range
};

// Caller-supplied symbol / region index -> [0, order): the reference's bounds-checked Cython raises IndexError
// for anything else; a kernel cannot raise, so it counts the offence (read back by qr_mapper_index_errors, which
// the Python classes turn into IndexError) and computes on index 0 instead of reading out of bounds.
QR_HD int32_t checked_index(const MapperView &m, long long idx)
{
    if (idx >= 0 && idx < (long long)m.order) return (int32_t)idx;
#if defined(__CUDA_ARCH__)
    if (m.index_errors) atomicAdd(m.index_errors, 1);
#else
    if (m.index_errors) ++*m.index_errors;
#endif
    return 0;
}

struct InvTable {
    const double *F;   // F_Y at y0 + j*h
    const double *f;   // its density at the same points (may be NULL: no Hermite solve)
    int32_t n;
    double y0, h;
    // optional: jump[t] = largest grid index g with F[g] <= t / jn (t = 0..jn, jn a power of two): narrows the
    // search for a target to the few grid cells of its bin instead of a binary search over the whole region
    const int32_t *jump = nullptr;
    int32_t jn = 0;
};

// the fp32-grade demapper's own, COARSE copy of that grid (1025 points: 14 KB, staged in shared memory by k_demap32,
// where a lookup costs ~30 cycles instead of an L2 round trip): F as double (target - F is a cancellation), the
// density as float, the jump table as 16-bit indices
constexpr int kInv32N = 1025, kInv32J = 1024;
struct InvTable32 {
    const double *F;
    const float *f;
    const uint16_t *jump;     // [kInv32J + 2]
    double y0, h;
};

// rounding-exact multiply/add: the reference is compiled without FMA contraction
QR_HD double mul_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
QR_HD double add_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

// __F_Z (noisemapper.pyx:66-67)
QR_HD double gauss_cdf(double z, double mu, double s2)
{
    return 0.5 * (1 + erf((z - mu) / s2));
}

// _single_F_Y (noisemapper.pyx:278-286): mixture CDF, summed k = 0 upward
QR_HD double mixture_cdf(const double *a, const double *p, int order, double s2, double y)
{
    double res = mul_rn(gauss_cdf(y, a[0], s2), p[0]);
    for (int i = 1; i < order; ++i) res = add_rn(res, mul_rn(gauss_cdf(y, a[i], s2), p[i]));
    return res;
}

// mixture pdf (derivative of the above), for the Newton solver
QR_HD double mixture_pdf(const double *a, const double *p, int order, double sigma, double y)
{
    double res = 0;
    for (int i = 0; i < order; ++i) {
        const double z = (y - a[i]) / sigma;
        res += p[i] * exp(-0.5 * z * z);
    }
    return res * (0.3989422804014327 / sigma);
}

// __binsearch (noisemapper.pyx:27-44) on an (offset, length) window instead of array slices;
// the comparisons come in the same order, so NaN and out-of-range inputs resolve identically.
QR_HD int32_t region_search(const double *dom, int32_t len, double val)
{
    int32_t base = 0;
    for (;;) {
        if (len == 1) return base;
        if (val < dom[0]) return base;
        if (val > dom[len - 1]) return base + len - 1;
        const int32_t mid = len / 2 - 1;
        if (val < dom[mid]) { len = mid; continue; }
        if (val >= dom[mid + 1]) { base += mid + 1; dom += mid + 1; len -= mid + 1; continue; }
        return base + mid;
    }
}

// hard_decide_index (noisemapper.pyx:349-359)
QR_HD int32_t hard_decide(const double *thresholds, int order, double y)
{
    int32_t r = region_search(thresholds, order + 1, y);
    return r == order ? order - 1 : r;
}

// Gray bit k of symbol i (bicm.pyx:26-41; closed form as at noisemapper.pyx:208-215)
QR_HD uint8_t gray_bit(int32_t i, int k)
{
    const int32_t q = i >> k;
    return (uint8_t)(((q * (q + 1)) & 3) != 0);
}

// g (noisemapper.pyx:289-292)
QR_HD double soften(const MapperView &m, const double *FYt, const double *delta, double F, int32_t i)
{
    return m.sign_config[i] ? (FYt[i + 1] - F) / delta[i] : (F - FYt[i]) / delta[i];
}

QR_HD double inv_target(const uint8_t *sign_config, const double *FYt, const double *delta, double n_hat,
                        int32_t i)
{
    // noisemapper.pyx:314-317
    return sign_config[i] ? add_rn(FYt[i + 1], -mul_rn(n_hat, delta[i]))
                          : add_rn(mul_rn(n_hat, delta[i]), FYt[i]);
}

constexpr int kMaxDoublings = 1100;  // 2^1100 overflows to inf: the reference would spin forever
constexpr int kMaxHalvings = 1200;

struct Bracket {
    double lo, hi;
    double f_end;  // F at the end the doubling loop last evaluated
};

// the doubling part of g_inv_search (noisemapper.pyx:319-334)
QR_HD Bracket bracket_root(const double *a, const double *p, int order, double s2, double target)
{
    Bracket b;
    if (target > .5) {
        b.hi = 1; b.lo = 0;
        b.f_end = mixture_cdf(a, p, order, s2, b.hi);
        for (int it = 0; b.f_end < target && it < kMaxDoublings; ++it) {
            b.lo = b.hi;
            b.hi *= 2.;
            b.f_end = mixture_cdf(a, p, order, s2, b.hi);
        }
    } else {
        b.lo = -1; b.hi = 0;
        b.f_end = mixture_cdf(a, p, order, s2, b.lo);
        for (int it = 0; b.f_end > target && it < kMaxDoublings; ++it) {
            b.hi = b.lo;
            b.lo *= 2.;
            b.f_end = mixture_cdf(a, p, order, s2, b.lo);
        }
    }
    return b;
}

// g_inv_search (noisemapper.pyx:310-345), exact replay: bisect on F until the bracket is <= 1e-9
QR_HD double g_inv_exact(const double *a, const double *p, int order, double s2, double target,
                         double accuracy)
{
    Bracket b = bracket_root(a, p, order, s2, target);
    double lo = b.lo, hi = b.hi;
    for (int it = 0; (hi - lo) > accuracy && it < kMaxHalvings; ++it) {
        const double mid = (hi + lo) / 2;
        if (mixture_cdf(a, p, order, s2, mid) > target) hi = mid;
        else lo = mid;
    }
    return (hi + lo) / 2;
}

// Acklam's rational approximation of the standard normal quantile (|rel err| < 1.2e-9); only a
// starting point for Newton, so the accuracy is ample.
QR_HD double norm_quantile(double q)
{
    const double a1 = -3.969683028665376e+01, a2 = 2.209460984245205e+02, a3 = -2.759285104469687e+02,
                 a4 = 1.383577518672690e+02, a5 = -3.066479806614716e+01, a6 = 2.506628277459239e+00;
    const double b1 = -5.447609879822406e+01, b2 = 1.615858368580409e+02, b3 = -1.556989798598866e+02,
                 b4 = 6.680131188771972e+01, b5 = -1.328068155288572e+01;
    const double c1 = -7.784894002430293e-03, c2 = -3.223964580411365e-01, c3 = -2.400758277161838e+00,
                 c4 = -2.549732539343734e+00, c5 = 4.374664141464968e+00, c6 = 2.938163982698783e+00;
    const double d1 = 7.784695709041462e-03, d2 = 3.224671290700398e-01, d3 = 2.445134137142996e+00,
                 d4 = 3.754408661907416e+00;
    if (q < 0.02425) {
        const double t = sqrt(-2 * log(q));
        return (((((c1 * t + c2) * t + c3) * t + c4) * t + c5) * t + c6) /
               ((((d1 * t + d2) * t + d3) * t + d4) * t + 1);
    }
    if (q > 1 - 0.02425) {
        const double t = sqrt(-2 * log(1 - q));
        return -(((((c1 * t + c2) * t + c3) * t + c4) * t + c5) * t + c6) /
               ((((d1 * t + d2) * t + d3) * t + d4) * t + 1);
    }
    const double t = q - 0.5, r = t * t;
    return (((((a1 * r + a2) * r + a3) * r + a4) * r + a5) * r + a6) * t /
           (((((b1 * r + b2) * r + b3) * r + b4) * r + b5) * r + 1);
}

// Fast variant of g_inv_search.  The bisection's outcome is a function of the root y* of
// F_Y(y) = target alone (F_Y is monotone: F(mid) > target  <=>  mid > y*; the doubling loop likewise
// compares powers of two with y*).  So: bracket y* analytically (interior regions: the decision
// thresholds, whose F values are the stored F_Y_thresholds; outer regions: single-Gaussian bounds),
// solve with a safeguarded Halley iteration (2-3 evaluations of F, f, f' instead of ~35 of F), then
// replay the reference's doubling and halving with comparisons against y* -- no F evaluations --
// which lands in the same 1e-9 cell and returns the same midpoint unless y* is within the solver
// error (~1e-14) of a midpoint.  Saturated targets (<= 0 or >= 1), whose result is defined by where
// erf rounds to +-1, take the exact path.
QR_HD double g_inv_fast(const double *a, const double *p, const double *thr, const double *FYt, int order,
                        double sigma, double s2, double target, double accuracy, int32_t region,
                        const InvTable tab)
{
    if (!(target > 0.0 && target < 1.0)) return g_inv_exact(a, p, order, s2, target, accuracy);
    double root = 0, f = 1;
    bool solved = false;
    // n_hat exactly 0 or 1 puts the target exactly on a stored threshold value: the root is the
    // decision threshold itself (a dyadic point for the usual constellations, where "mid > root"
    // must be decided exactly as the reference's "F(mid) > target" is)
    if (region > 0 && target == FYt[region]) { root = thr[region]; solved = true; }
    else if (region < order - 1 && target == FYt[region + 1]) { root = thr[region + 1]; solved = true; }
    if (!solved && tab.n > 1 && tab.f && target >= tab.F[0] && target < tab.F[tab.n - 1]) {
        // Table solve, no erf/exp at all: F_Y and its density are tabulated on a uniform grid of step
        // h ~ 2e-3; the cubic Hermite interpolant through (F, f) at the two grid points around the root
        // reproduces F_Y to h^4 |F^(4)| / 384 ~ 1e-14, so solving the cubic for the target gives the
        // root to ~1e-14 / f -- far inside the 1e-9 cell the result is snapped to.
        int32_t lo = 0, hi = tab.n - 1;
        if (region > 0) {            // the root lies in the decision region: start from its grid bracket
            const int32_t g = (int32_t)((thr[region] - tab.y0) / tab.h) - 1;
            if (g > lo && g < hi && tab.F[g] <= target) lo = g;
        }
        if (region < order - 1) {
            const int32_t g = (int32_t)((thr[region + 1] - tab.y0) / tab.h) + 2;
            if (g > lo && g < hi && tab.F[g] > target) hi = g;
        }
        if (tab.jump) {
            const int32_t t = (int32_t)(target * tab.jn);       // exact: jn is a power of two
            const int32_t g0 = tab.jump[t], g1 = tab.jump[t + 1] + 1;
            if (g0 > lo) lo = g0;                               // F[g0] <= t / jn <= target
            if (g1 < hi) hi = g1;                               // F[g1] > (t + 1) / jn > target
        }
        while (hi - lo > 1) {
            const int32_t mid = (lo + hi) >> 1;
            if (tab.F[mid] <= target) lo = mid; else hi = mid;
        }
        const double F0 = tab.F[lo], F1 = tab.F[hi], d0 = tab.f[lo], d1 = tab.f[hi];
        const double dF = F1 - F0, d = target - F0, A = tab.h * d0, B = tab.h * d1;
        if (dF > 0 && (d0 < d1 ? d0 : d1) * accuracy > 1e-13) {
            // G(s) = F_hermite(y_lo + s h) - F0 on s in [0, 1]
            const double c2 = 3 * dF - 2 * A - B, c3 = A + B - 2 * dF;
            double sx = d / dF;
            for (int it = 0; it < 3; ++it) {
                const double G = sx * (A + sx * (c2 + sx * c3));
                const double Gp = A + sx * (2 * c2 + 3 * sx * c3);
                sx -= (G - d) / Gp;
                sx = sx < 0 ? 0 : (sx > 1 ? 1 : sx);
            }
            root = tab.y0 + lo * tab.h + sx * tab.h;
            f = d0 + sx * (d1 - d0);
            solved = true;
        }
    }
    if (!solved) {
    double rlo, rhi, y;
    bool presolved = false;
    if (tab.n > 1 && target >= tab.F[0] && target < tab.F[tab.n - 1]) {
        int32_t lo = 0, hi = tab.n - 1;
        while (hi - lo > 1) {
            const int32_t mid = (lo + hi) >> 1;
            if (tab.F[mid] <= target) lo = mid; else hi = mid;
        }
        const double f0 = tab.F[lo], f1 = tab.F[hi];
        rlo = tab.y0 + lo * tab.h;
        rhi = rlo + tab.h;
        y = rlo + (f1 > f0 ? (target - f0) / (f1 - f0) : 0.5) * tab.h;
        rlo -= 1e-9; rhi += 1e-9;     // the table was rounded: leave the safeguard a hair of slack
        presolved = true;
    } else if (region > 0 && region < order - 1) {
        rlo = thr[region];
        rhi = thr[region + 1];
        const double t0 = FYt[region], t1 = FYt[region + 1];
        if (!(target >= t0 && target <= t1)) return g_inv_exact(a, p, order, s2, target, accuracy);
        y = rlo + (target - t0) / (t1 - t0) * (rhi - rlo);
    } else if (region == 0) {
        // p_0 Phi((y-a_0)/s) <= F(y) <= Phi((y-a_0)/s)
        rhi = thr[1];
        rlo = a[0] + sigma * norm_quantile(target) - 1e-6 * sigma;
        const double q = target / p[0];
        if (q < 1.0) rhi = fmin(rhi, a[0] + sigma * norm_quantile(q) + 1e-6 * sigma);
        if (!(FYt[1] >= target)) return g_inv_exact(a, p, order, s2, target, accuracy);
        y = rhi;
    } else {
        // Phi((y-a_top)/s) <= F(y) <= 1 - p_top + p_top Phi((y-a_top)/s)
        const int top = order - 1;
        rlo = thr[top];
        rhi = a[top] + sigma * norm_quantile(target) + 1e-6 * sigma;
        const double q = (target - (1.0 - p[top])) / p[top];
        if (q > 0.0) rlo = fmax(rlo, a[top] + sigma * norm_quantile(q) - 1e-6 * sigma);
        if (!(FYt[top] <= target)) return g_inv_exact(a, p, order, s2, target, accuracy);
        y = rlo;
    }
    if (!(rlo < rhi)) return g_inv_exact(a, p, order, s2, target, accuracy);
    if (!(y > rlo && y < rhi)) y = 0.5 * (rlo + rhi);
    // Cheap float pre-solve (3 Newton steps with erff/expf bring y within ~1e-6 of the root where the
    // density is not tiny), so that the double-precision Halley loop below usually needs ONE
    // evaluation: its cubic step from 1e-6 lands at ~1e-18.  Anything the float solve gets wrong
    // (tails, flat spots) is caught by the bracket and the convergence test of the double loop.
    if (!presolved) {
        const float s2f = (float)s2, tf = (float)target, c0f = (float)(0.3989422804014327 / sigma);
        float yf = (float)y;
        const float lof = (float)rlo, hif = (float)rhi;
        for (int it = 0; it < 3; ++it) {
            float F = 0.f, f = 0.f;
            for (int k = 0; k < order; ++k) {
                const float z = (yf - (float)a[k]) / s2f;
                F += (float)p[k] * (0.5f * (1.f + erff(z)));
                f += (float)p[k] * expf(-z * z);
            }
            const float yn = yf - (F - tf) / (f * c0f);
            if (!(yn > lof && yn < hif)) break;
            yf = yn;
        }
        if ((double)yf > rlo && (double)yf < rhi) y = (double)yf;
    }
    const double c0 = 0.3989422804014327 / sigma, c1 = c0 / (sigma * sigma);
    root = y;
    for (int it = 0; it < 100; ++it) {
        double F = 0, fp = 0;
        f = 0;
        for (int k = 0; k < order; ++k) {
            const double d = y - a[k];
            const double z = d / s2;
            const double e = p[k] * exp(-z * z);
            F += p[k] * (0.5 * (1 + erf(z)));
            f += e;
            fp -= e * d;
        }
        f *= c0;
        fp *= c1;
        const double g = F - target;
        if (g > 0) rhi = y; else rlo = y;
        const double den = 2 * f * f - g * fp;
        bool cubic = den > 0 && f > 0;
        double step = cubic ? 2 * g * f / den : g / f;
        double yn = y - step;
        bool safe = yn > rlo && yn < rhi;
        if (!safe) { yn = 0.5 * (rlo + rhi); step = y - yn; }
        y = yn;
        root = y;
        // Halley: the next error is ~ step^3 / sigma^2, Newton: ~ step^2 / sigma
        if (safe && fabs(step) <= (cubic ? 1e-5 : 3e-8) * sigma) break;
        if ((rhi - rlo) <= 4e-16 * fmax(1.0, fabs(y))) break;
    }
    }   // !solved
    // Deep in the tails F is flat at the resolution of a double (one ulp of F spans f^-1 * 1e-16 in y):
    // there the reference's answer is set by the rounding of F, not by the root.  Replay it exactly.
    if (!(f * accuracy > 1e-14)) return g_inv_exact(a, p, order, s2, target, accuracy);
    // replay of noisemapper.pyx:319-344 against the root
    if (accuracy == 1e-9 && fabs(root) < 4.0e6) {
        // Default accuracy, closed form.  The bracket [lo, hi] the doubling loop leaves has a power-of-two width
        // W = 2^e and ends that are multiples of it; 2^-30 <= 1e-9 < 2^-29, so the halving loop always stops at
        // cells of width exactly 2^-30 aligned at 0, whatever e, and returns the midpoint of the cell with
        // lo_cell <= root < hi_cell:  (floor(root * 2^30) + 1/2) * 2^-30  (all operations exact in double).
        // One exception: a root that IS the upper end of the bracket (target > 1/2 and root a power of two >= 1:
        // the doubling loop stops at hi == root) belongs to the last cell of the bracket, the one below it.
        double q = floor(root * 0x1p30);
        if (target > .5 && root >= 1.0) {
            int ex = 0;
            if (frexp(root, &ex) == 0.5) q -= 1.0;
        }
        if (target > .5 && root <= 0.0) q = 0.0;                    // bracket [0, 1]: roots below it go to its first cell
        if (!(target > .5) && root >= 0.0) q = -1.0;                // bracket [-1, 0]: roots above it go to its last cell
        return (q + 0.5) * 0x1p-30;
    }
    double lo, hi;
    if (target > .5) {
        hi = 1; lo = 0;
        for (int it = 0; hi < root && it < kMaxDoublings; ++it) { lo = hi; hi *= 2.; }
    } else {
        lo = -1; hi = 0;
        for (int it = 0; lo > root && it < kMaxDoublings; ++it) { hi = lo; lo *= 2.; }
    }
    // The halving loop picks, among the 2^k cells of width W / 2^k (k = halvings until the width is
    // <= accuracy), the one with lo_cell <= root < hi_cell (a root on a boundary goes to the upper cell,
    // root >= hi to the last, root < lo to the first).  lo, hi and every midpoint are dyadic, so the
    // cell can be computed directly and the result is bit-identical to running the loop.
    const double W = hi - lo;
    int k = 0;
    {
        // halvings until W / 2^k <= accuracy.  W is a power of two here (0/1 or doubled bounds), so k
        // follows from the exponents; the two probes settle the rounding of the estimate exactly.
        int ew = 0, ea = 0;
        (void)frexp(W, &ew);
        (void)frexp(accuracy, &ea);
        k = ew - ea + 1;
        if (k < 0) k = 0;
        while (k > 0 && ldexp(W, -(k - 1)) <= accuracy) --k;
        while (k < kMaxHalvings && ldexp(W, -k) > accuracy) ++k;
    }
    if (k > 52) {   // not reachable for accuracy = 1e-9 and finite brackets; keep the loop for safety
        for (int it = 0; (hi - lo) > accuracy && it < kMaxHalvings; ++it) {
            const double mid = (hi + lo) / 2;
            if (mid > root) hi = mid;
            else lo = mid;
        }
        return (hi + lo) / 2;
    }
    const double cell = ldexp(W, -k), last = ldexp(1.0, k) - 1.0;
    double idx = floor((root - lo) / cell);
    idx = idx < 0.0 ? 0.0 : (idx > last ? last : idx);
    // (root - lo) may have been rounded across a boundary: settle with exact comparisons
    if (idx > 0.0 && lo + idx * cell > root) idx -= 1.0;
    else if (idx < last && lo + (idx + 1.0) * cell <= root) idx += 1.0;
    return lo + (idx + 0.5) * cell;
}

// ---------------------------------------------------------------------------------------------
// fp32-GRADE demapper (QR_DEMAP_F32GRADE): for LLRs that are consumed as float.  The reference's result
// is a function of the root y* of F_Y(y) = target, snapped to a 2^-30 cell and pushed through fp64
// exp / log / divisions -- fidelity a float LLR cannot show.  Here the root is taken from the same
// Hermite table (one Newton step on the cubic: |dy| ~ 1e-11, no snap), exponentials are a double range
// reduction + ONE MUFU.EX2, reciprocals a float seed + one Newton step, the final log a MUFU.LG2 on the
// mantissa.  Relative error of the LLR ~1e-6 (tests: 1e-5 relative + 1e-6 absolute against the exact replay).
QR_HD float ex2_f32(float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return exp2f(x);
#endif
}
QR_HD float rcp_f32(float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
// 2^e as a double, |e| <= 1022 (exact: written into the exponent field)
QR_HD double pow2_int(int32_t e)
{
    e = e < -1022 ? -1022 : (e > 1023 ? 1023 : e);
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)(e + 1023) << 52);
#else
    return ldexp(1.0, e);
#endif
}

QR_HD double rcp_f32grade(double x)
{
    if (!(fabs(x) > 1e-30 && fabs(x) < 1e30)) return 1.0 / x;     // outside the float range (rare): the real division
#if defined(__CUDA_ARCH__)
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"((float)x));
#else
    const float r0 = 1.0f / (float)x;
#endif
    const double r = (double)r0;
    return fma(r, fma(-x, r, 1.0), r);      // one Newton step: ~1e-13 relative
}

// e^u, |u| < 700, relative error ~2^-22 whatever |u| (the range reduction is done in double)
QR_HD double exp_f32grade(double u)
{
    const double t = u * 1.4426950408889634;
#if defined(__CUDA_ARCH__)
    const int32_t ni = __double2int_rn(t);
    float m;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(m) : "f"((float)(t - (double)ni)));
    return (double)m * __longlong_as_double((long long)(ni + 1023) << 52);
#else
    const int32_t ni = (int32_t)nearbyint(t);
    const float m = exp2f((float)(t - (double)ni));
    return ldexp((double)m, ni);
#endif
}

// ln(r) for a positive finite double r: exponent taken from the bits, MUFU.LG2 on the mantissa in [0.75, 1.5)
// (so that r ~ 1, an LLR near zero, keeps the small absolute error of the approximation)
QR_HD double log_f32grade(double r)
{
#if defined(__CUDA_ARCH__)
    long long b = __double_as_longlong(r);
    int32_t e = (int32_t)((b >> 52) & 0x7ff) - 1023;
    double m = __longlong_as_double((b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);   // [1, 2)
    if (m >= 1.5) { m *= 0.5; e += 1; }
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"((float)m));
    return 0.6931471805599453 * ((double)e + (double)l);
#else
    int e = 0;
    double m = 2.0 * frexp(r, &e);      // [1, 2)
    e -= 1;
    if (m >= 1.5) { m *= 0.5; e += 1; }
    return 0.6931471805599453 * ((double)e + (double)log2f((float)m));
#endif
}

// (out of line: rare, and its register needs must not shape the callers)
static QR_HD_NOINLINE double g_inv_slow_path(const double *a, const double *p, const double *thr, const double *FYt,
                                             int order, double sigma, double s2, double target, int32_t region,
                                             const InvTable tab)
{
    return g_inv_fast(a, p, thr, FYt, order, sigma, s2, target, 1e-9, region, tab);
}

// root of F_Y(y) = target to ~1e-10 where the density is not vanishing; everything else (saturated targets, deep
// tails, targets outside the table) goes through g_inv_fast, which ends on the reference's cell
QR_HD double g_inv_f32grade(const double *a, const double *p, const double *thr, const double *FYt, int order,
                            double sigma, double s2, double target, int32_t region, const InvTable32 t32,
                            const InvTable tab)
{
    if (region > 0 && target == FYt[region]) return thr[region];
    if (region < order - 1 && target == FYt[region + 1]) return thr[region + 1];
    if (target > 0.0 && target < 1.0 && t32.F && target >= t32.F[0] && target < t32.F[kInv32N - 1]) {
        const int32_t t = (int32_t)(target * kInv32J);     // exact: a power of two
        int32_t lo = t32.jump[t], hi = (int32_t)t32.jump[t + 1] + 1;   // F[lo] <= t / J <= target < (t + 1) / J < F[hi]
        if (hi > kInv32N - 1) hi = kInv32N - 1;
        while (hi - lo > 1) {
            const int32_t mid = (lo + hi) >> 1;
            if (t32.F[mid] <= target) lo = mid; else hi = mid;
        }
        const double F0 = t32.F[lo], F1 = t32.F[hi];
        const float d0 = t32.f[lo], d1 = t32.f[hi];
        const double dF = F1 - F0, d = target - F0;
        if (dF > 1e-30 && (d0 < d1 ? d0 : d1) > 1e-5f) {
            // the two differences above are the cancellations that need double; the cubic itself is solved in float
            const float hf = (float)t32.h, dFf = (float)dF, df = (float)d;
            const float A = hf * d0, B = hf * d1;
            const float c2 = 3.0f * dFf - 2.0f * A - B, c3 = A + B - 2.0f * dFf;     // cubic Hermite through (F, f) at both ends
            float sx = df * rcp_f32(dFf);
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const float G = sx * fmaf(sx, fmaf(sx, c3, c2), A);
                const float Gp = fmaf(sx, fmaf(3.0f * sx, c3, 2.0f * c2), A);
                sx -= (G - df) * rcp_f32(Gp);
            }
            sx = fminf(fmaxf(sx, 0.0f), 1.0f);
            return fma((double)sx, t32.h, fma((double)lo, t32.h, t32.y0));
        }
    }
    return g_inv_slow_path(a, p, thr, FYt, order, sigma, s2, target, region, tab);
}

// demap_lappr (noisemapper.pyx:450-540) given the `order` reconstructed samples y_hat[i].
// NOTE the reference divides the exponent by 2*sigma^2 only for k > j (:511-515), not for k < j
// (:503-507); `corrected` divides in both.
QR_HD void demap_from_yhat(const double *a, const double *p, const double *delta, int order, int bps,
                           double two_s2, const double *y_hat, int32_t j, bool corrected, double *lappr)
{
    double N[kMaxBps], D[kMaxBps];
    for (int k = 0; k < bps; ++k) { N[k] = 0; D[k] = 0; }
    for (int i = 0; i < order; ++i) {
        const double yh = y_hat[i];
        double s = 0;
        for (int k = 0; k < j; ++k) {
            double ex = mul_rn(add_rn(add_rn(mul_rn(2, yh), -a[k]), -a[j]), add_rn(a[k], -a[j]));
            if (corrected) ex = ex / two_s2;
            s = add_rn(s, mul_rn(exp(ex), p[k]));
        }
        s = add_rn(s, p[j]);
        for (int k = j + 1; k < order; ++k) {
            const double ex = mul_rn(add_rn(add_rn(mul_rn(2, yh), -a[k]), -a[j]), add_rn(a[k], -a[j])) / two_s2;
            s = add_rn(s, mul_rn(exp(ex), p[k]));
        }
        const double w = delta[i] / s;
        int q = i;
        for (int k = 0; k < bps; ++k) {
            if ((q * (q + 1)) & 3) D[k] = add_rn(D[k], w);
            else N[k] = add_rn(N[k], w);
            q >>= 1;
        }
    }
    for (int k = 0; k < bps; ++k) lappr[k] = log(N[k]) - log(D[k]);
}

// The small per-alphabet tables a kernel works from (shared memory on the device, plain arrays on
// the host).
struct TablesRef {
    const double *a, *p, *thr, *FYt, *delta;
    const uint8_t *sign;
    const double *ghi = nullptr, *glo = nullptr, *pz = nullptr;   // see SharedTables (fast demapper only)
    // fp32-grade demapper: (m step)^2 c log2(e) for c = 1 / 2 sigma^2 (g2hi) and c = 1 (g2lo, the reference's
    // undivided k < j exponent), and the zero-padded probabilities, as floats
    const float *g2hi = nullptr, *g2lo = nullptr, *pzf = nullptr;
    InvTable32 t32{nullptr, nullptr, nullptr, 0.0, 0.0};     // coarse inverse table (shared memory in k_demap32)
};

// demap_lappr (noisemapper.pyx:450-540) for ONE symbol: Bob's metric n_hat, Alice's symbol j -> bps
// LLRs (scaled by alpha, sims/reconciliation.pyx:144-145).  mode: QR_DEMAP_* bits.
QR_HD void demap_symbol(const MapperView &m, const TablesRef &s, double nv, int32_t j, int mode, double alpha,
                        double *out)
{
    const bool fast = (mode & 1) != 0, corrected = (mode & QR_DEMAP_CORRECTED) != 0;
    const bool grade32 = fast && (mode & QR_DEMAP_F32GRADE) != 0;
    const double two_s2 = 2 * m.noise_var;
    // fast mode (tolerance 1e-7): multiply by the reciprocal instead of dividing twelve times
    // per symbol, one log of the ratio instead of two logs; exact mode keeps the reference's operations
    const double inv_two_s2 = 1.0 / two_s2;
    double N[kMaxBps], D[kMaxBps];
    for (int k = 0; k < m.bps; ++k) { N[k] = 0; D[k] = 0; }
    for (int i = 0; i < m.order; ++i) {
        const double target = inv_target(s.sign, s.FYt, s.delta, nv, i);
        const InvTable tab{m.inv_tab, m.inv_pdf, m.inv_n, m.inv_y0, m.inv_h, m.inv_jump, m.inv_jn};
        const double yh = grade32 ? g_inv_f32grade(s.a, s.p, s.thr, s.FYt, m.order, m.sigma, m.s2, target, i, s.t32, tab)
                          : fast  ? g_inv_fast(s.a, s.p, s.thr, s.FYt, m.order, m.sigma, m.s2, target, 1e-9, i, tab)
                                  : g_inv_exact(s.a, s.p, m.order, m.s2, target, 1e-9);
        double sum = 0;
        // Fast mode, uniform constellation: with d = y_hat - a_j and a_k - a_j = m step the exponent is
        // (2 d - m step)(m step) c = m (2 d step c) - (m step)^2 c, so exp of it = E^m G[m] with ONE exp per side
        // (E = exp(2 d step c); c = 1 / 2 sigma^2 for k > j and -- the reference's quirk -- 1 for k < j) and
        // G tabulated.  Falls back to the direct form if a power could overflow.
        const int M = m.order;
        const double step = M > 1 ? s.a[1] - s.a[0] : 0.0, dd = yh - s.a[j];
        const double u_hi = 2 * dd * step * inv_two_s2, u_lo = -2 * dd * step * (corrected ? inv_two_s2 : 1.0);
        if (fast && s.ghi && fabs(u_hi) * (M - 1) < 600.0 && fabs(u_lo) * (M - 1) < 600.0) {
            // the same instruction stream whatever j (threads of a warp hold different j): both sides always run
            // all M - 1 powers, against probabilities zero-padded outside the alphabet
            const double *glo = corrected ? s.ghi : s.glo;
            const double *pj = s.pz + (M - 1) + j;
            const double Eh = grade32 ? exp_f32grade(u_hi) : exp(u_hi), El = grade32 ? exp_f32grade(u_lo) : exp(u_lo);
            double ph = 1.0, pl = 1.0;
            sum = s.p[j];
            for (int mm = 1; mm < M; ++mm) {
                ph *= Eh; pl *= El;
                sum += pj[mm] * (ph * s.ghi[mm]) + pj[-mm] * (pl * glo[mm]);
            }
            const double w = grade32 ? s.delta[i] * rcp_f32grade(sum) : s.delta[i] / sum;
            int q = i;
            for (int k = 0; k < m.bps; ++k) {
                if ((q * (q + 1)) & 3) D[k] += w;
                else N[k] += w;
                q >>= 1;
            }
            continue;
        }
        for (int k = 0; k < j; ++k) {
            double ex = mul_rn(add_rn(add_rn(mul_rn(2, yh), -s.a[k]), -s.a[j]), add_rn(s.a[k], -s.a[j]));
            if (corrected) ex = fast ? ex * inv_two_s2 : ex / two_s2;
            sum = add_rn(sum, mul_rn(exp(ex), s.p[k]));
        }
        sum = add_rn(sum, s.p[j]);
        for (int k = j + 1; k < m.order; ++k) {
            const double pr = mul_rn(add_rn(add_rn(mul_rn(2, yh), -s.a[k]), -s.a[j]), add_rn(s.a[k], -s.a[j]));
            const double ex = fast ? pr * inv_two_s2 : pr / two_s2;
            sum = add_rn(sum, mul_rn(exp(ex), s.p[k]));
        }
        const double w = s.delta[i] / sum;
        int q = i;
        for (int k = 0; k < m.bps; ++k) {
            if ((q * (q + 1)) & 3) D[k] = add_rn(D[k], w);
            else N[k] = add_rn(N[k], w);
            q >>= 1;
        }
    }
    for (int k = 0; k < m.bps; ++k) {
        double v;
        if (grade32) {
            const double r = N[k] * rcp_f32grade(D[k]);
            // (a ratio outside the normal doubles -- one side vanished -- takes the library log)
            v = (r > 1e-300 && r < 1e300) ? log_f32grade(r) : log(N[k] / D[k]);
        } else {
            v = fast ? log(N[k] / D[k]) : log(N[k]) - log(D[k]);
        }
        if (alpha != 1.0) v = mul_rn(v, alpha);
        out[k] = v;
    }
}

// The sum over Alice's symbols of one hypothesis in the direct (non E^m) form, library exp: the rare case where a power
// of E could overflow (see demap_symbol)
static QR_HD_NOINLINE double demap_sum_direct(const TablesRef &s, int order, int32_t j, double yh, bool corrected,
                                              double inv_two_s2)
{
    double sum = 0;
    for (int k = 0; k < order; ++k) {
        if (k == j) { sum += s.p[j]; continue; }
        double ex = (2 * yh - s.a[k] - s.a[j]) * (s.a[k] - s.a[j]);
        if (k > j || corrected) ex *= inv_two_s2;
        sum += exp(ex) * s.p[k];
    }
    return sum;
}

// demap_lappr for ONE symbol at fp32 grade, alphabet size known at compile time (everything unrolled, accumulators
// in registers): equally spaced constellations only (PAMAlphabet always is).  Same formulas as demap_symbol's fast
// path, evaluated in FLOAT after the two cancellations that need double (target - F0 on the table, y_hat - a_j):
//   * root: linear guess + one Newton step on the Hermite cubic, in float (error ~1e-7 of a 1.5e-3 wide cell);
//   * the sum over Alice's symbols: term k has the exponent m u - (m step)^2 c, m = k - j, u = 2 (y_hat - a_j) step c;
//     all exponents of a hypothesis are formed in log2 units, their integer maximum e_max is factored out (the
//     undivided k < j exponents of the reference reach several hundred for 8-PAM: 2^e_max is applied to the weight
//     as an exact exponent-field operation on a double), each term is one FADD + MUFU.EX2 + FFMA;
//   * weights delta_i / sum accumulate in double (their range spans hundreds of binades), the LLR is one
//     MUFU.LG2 on the mantissa of N / D.
template <int BPS>
QR_HD void demap_symbol_f32grade(const MapperView &m, const TablesRef &s, double nv, int32_t j, bool corrected,
                                 double alpha, double *out)
{
    constexpr int M = 1 << BPS;
    const double inv_two_s2 = rcp_f32grade(2 * m.noise_var);
    const double step = M > 1 ? s.a[1] - s.a[0] : 0.0;
    const InvTable tab{m.inv_tab, m.inv_pdf, m.inv_n, m.inv_y0, m.inv_h, m.inv_jump, m.inv_jn};
    const float *g2lo = corrected ? s.g2hi : s.g2lo;
    const float *pj = s.pzf + (M - 1) + j;
    const double aj = s.a[j];
    const double k_hi = 2.0 * step * inv_two_s2 * 1.4426950408889634;                       // u (log2 units) per unit of y_hat - a_j
    const double k_lo = -2.0 * step * (corrected ? inv_two_s2 : 1.0) * 1.4426950408889634;
    const float pjj = pj[0];
    double N[BPS], D[BPS];
#pragma unroll
    for (int k = 0; k < BPS; ++k) { N[k] = 0; D[k] = 0; }
#pragma unroll
    for (int i = 0; i < M; ++i) {
        const double target = inv_target(s.sign, s.FYt, s.delta, nv, i);
        const double yh = g_inv_f32grade(s.a, s.p, s.thr, s.FYt, M, m.sigma, m.s2, target, i, s.t32, tab);
        const double dd = yh - aj;
        const float uh = (float)(dd * k_hi), ul = (float)(dd * k_lo);
        double w;
        if (fabsf(uh) < 300.0f && fabsf(ul) < 300.0f) {
            float th[M], tlo[M];
            float tmax = 0.0f;                                   // (the k = j term has exponent 0)
#pragma unroll
            for (int mm = 1; mm < M; ++mm) {
                th[mm] = fmaf((float)mm, uh, -s.g2hi[mm]);
                tlo[mm] = fmaf((float)mm, ul, -g2lo[mm]);
                // (terms outside the alphabet carry probability 0: keep them out of the maximum)
                if (pj[mm] > 0.0f) tmax = fmaxf(tmax, th[mm]);
                if (pj[-mm] > 0.0f) tmax = fmaxf(tmax, tlo[mm]);
            }
            const float emax = ceilf(tmax);
            float sum = pjj * ex2_f32(-emax);
#pragma unroll
            for (int mm = 1; mm < M; ++mm) {
                // (min: a zero-probability term may lie above the maximum; 0 * inf must not happen)
                sum = fmaf(pj[mm], ex2_f32(fminf(th[mm] - emax, 0.0f)), sum);
                sum = fmaf(pj[-mm], ex2_f32(fminf(tlo[mm] - emax, 0.0f)), sum);
            }
            w = s.delta[i] * (double)rcp_f32(sum) * pow2_int(-(int32_t)emax);      // delta_i / (sum 2^emax)
        } else {
            // a reconstructed sample absurdly far out (saturated metric): the reference's double arithmetic, out of line
            w = s.delta[i] / demap_sum_direct(s, M, j, yh, corrected, inv_two_s2);
        }
#pragma unroll
        for (int k = 0; k < BPS; ++k) {
            const int q = i >> k;
            if ((q * (q + 1)) & 3) D[k] += w;        // Gray bit k of hypothesis i (compile-time after unrolling)
            else N[k] += w;
        }
    }
#pragma unroll
    for (int k = 0; k < BPS; ++k) {
        const double r = N[k] * rcp_f32grade(D[k]);
        double v = (r > 1e-300 && r < 1e300) ? log_f32grade(r) : log(N[k] / D[k]);
        if (alpha != 1.0) v *= alpha;
        out[k] = v;
    }
}

// one symbol, any mode: what k_demap / k_demap32 run
QR_HD void demap_symbol_any(const MapperView &m, const TablesRef &s, double nv, int32_t j, int mode, double alpha,
                            double *out)
{
    const bool corrected = (mode & QR_DEMAP_CORRECTED) != 0;
    if ((mode & QR_DEMAP_F32GRADE) && (mode & 1) && m.uniform && s.ghi) {
        switch (m.bps) {
        case 1: demap_symbol_f32grade<1>(m, s, nv, j, corrected, alpha, out); return;
        case 2: demap_symbol_f32grade<2>(m, s, nv, j, corrected, alpha, out); return;
        case 3: demap_symbol_f32grade<3>(m, s, nv, j, corrected, alpha, out); return;
        case 4: demap_symbol_f32grade<4>(m, s, nv, j, corrected, alpha, out); return;
        default: break;
        }
    }
    demap_symbol(m, s, nv, j, mode, alpha, out);
}

// direct-reconciliation LLR (sims/reconciliation.pyx:25-51)
QR_HD void direct_llr(const double *a, int order, int bps, double two_variance, double y, double *lappr)
{
    double N[kMaxBps], D[kMaxBps];
    for (int l = 0; l < bps; ++l) { N[l] = 0; D[l] = 0; }
    for (int i = 0; i < order; ++i) {
        const double d = y - a[i];
        const double term = exp(-mul_rn(d, d) / two_variance);
        int q = i;
        for (int l = 0; l < bps; ++l) {
            if ((q * (q + 1)) & 3) D[l] = add_rn(D[l], term);
            else N[l] = add_rn(N[l], term);
            q >>= 1;
        }
    }
    for (int l = 0; l < bps; ++l) lappr[l] = log(N[l]) - log(D[l]);
}

}  // namespace qr
