// Rest of the NoiseMapper surface: the dense F_Y grid, g_inv / demap_noise (grid
// interpolation), the "simplified" and "sofisticated" LLR formulations, F_Y, F_Z, and the sign rule of the
// FlipSign / AntiFlipSign subclasses (reference: noisemapper.pyx:47-98, :135-144, :264-307, :391-404,
// :563-816).  Elementwise, fp64, one thread per element; the grid (a few 10^4 points) is read through L1/L2.
#include <cuda_runtime.h>

#include <cmath>

#include "qr_handles.h"
#include "qr_mapper_core.cuh"
#include "qr_mapper_device.cuh"

namespace qr {

// __F_Z (noisemapper.pyx:66-67) with the reference's operand order: (z - mu) / (sqrt(2) * sigma)
__device__ __forceinline__ double f_z(double z, double mu, double s2) { return 0.5 * (1 + erf((z - mu) / s2)); }

// NoiseMapper.F_Y (noisemapper.pyx:264-275): UNIFORM weights, summed i = 0 upward, divided by the order last
__device__ __forceinline__ double f_y_uniform(const double *a, int order, double s2, double y)
{
    double res = f_z(y, a[0], s2);
    for (int i = 1; i < order; ++i) res = add_rn(res, f_z(y, a[i], s2));
    return res / order;
}

// __interp (noisemapper.pyx:47-63)
__device__ __forceinline__ double interp(const double *__restrict__ dom, const double *__restrict__ cod, int32_t n,
                                         double val)
{
    if (val >= dom[n - 1]) return cod[n - 1];
    const int32_t i = region_search(dom, n, val);
    if (i == n - 1) return cod[i];
    const double d0 = dom[i], d1 = dom[i + 1];
    if (d1 == d0) return cod[i];
    const double c0 = cod[i];
    return add_rn(c0, mul_rn(add_rn(cod[i + 1], -c0), add_rn(val, -d0)) / add_rn(d1, -d0));
}

// g_inv (noisemapper.pyx:295-307; subclasses :786-797, :805-816 through the sign vector)
__device__ __forceinline__ double g_inv_grid(const SharedTables &s, const uint8_t *sign_g, const double *gF,
                                             const double *gy, int32_t npts, double n_hat, int32_t i)
{
    const double prod = mul_rn(n_hat, s.delta[i]);
    const double target = sign_g[i] ? add_rn(s.FYt[i + 1], -prod) : add_rn(prod, s.FYt[i]);
    return interp(gF, gy, npts, target);
}

__global__ void k_grid(MapperView m, double y_low, double y_high, int32_t n, double *__restrict__ y,
                       double *__restrict__ F)
{
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // numpy.linspace: arange(n) * step + low (two roundings), last point = high
    const double step = (y_high - y_low) / (double)(n - 1);
    double v = add_rn(mul_rn((double)i, step), y_low);
    if (i == n - 1 && n > 1) v = y_high;
    y[i] = v;
    F[i] = f_y_uniform(m.constellation, m.order, m.s2, v);
}

__global__ void k_F_Y(MapperView m, const double *__restrict__ y, int64_t n, double *__restrict__ out)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
        out[j] = f_y_uniform(m.constellation, m.order, m.s2, y[j]);
}

__global__ void k_F_Z(const double *__restrict__ z, int64_t n, double mu, double s2, double *__restrict__ out)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
        out[j] = f_z(z[j], mu, s2);
}

__global__ void __launch_bounds__(128) k_demap_noise(MapperView m, const uint8_t *__restrict__ sign_g,
                                                     const double *__restrict__ gF, const double *__restrict__ gy,
                                                     int32_t npts, const double *__restrict__ n_hat,
                                                     const long long *__restrict__ symb, int64_t n,
                                                     double *__restrict__ y_hat)
{
    __shared__ SharedTables s;
    __shared__ uint8_t sg[kMaxOrder];
    stage_tables(m, s);
    for (int i = threadIdx.x; i < m.order; i += blockDim.x) sg[i] = sign_g[i];
    __syncthreads();
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
        y_hat[j] = g_inv_grid(s, sg, gF, gy, npts, n_hat[j], checked_index(m, symb[j]));
}

// variant 1: demap_lappr_simplified (noisemapper.pyx:563-601); variant 2: demap_lappr_sofisticated (:624-748)
template <int VARIANT>
__global__ void __launch_bounds__(128) k_demap_variant(MapperView m, const uint8_t *__restrict__ sign_g,
                                                       const double *__restrict__ gF, const double *__restrict__ gy,
                                                       int32_t npts, const double *__restrict__ inf_erf,
                                                       const double *__restrict__ n_hat,
                                                       const long long *__restrict__ tx, int64_t n,
                                                       double *__restrict__ lappr)
{
    __shared__ SharedTables s;
    __shared__ uint8_t sg[kMaxOrder];
    stage_tables(m, s);
    for (int i = threadIdx.x; i < m.order; i += blockDim.x) sg[i] = sign_g[i];
    __syncthreads();
    const int M = m.order, bps = m.bps;
    const double two_s2 = 2 * m.noise_var;
    for (int64_t sidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; sidx < n;
         sidx += (int64_t)gridDim.x * blockDim.x) {
        const int32_t j = checked_index(m, tx[sidx]);
        const double nv = n_hat[sidx], a_j = s.a[j];
        double N[kMaxBps], D[kMaxBps];
        for (int k = 0; k < bps; ++k) { N[k] = 0; D[k] = 0; }
        if (VARIANT == 1) {
            for (int i = 0; i < M; ++i) {
                const double yh = g_inv_grid(s, sg, gF, gy, npts, nv, i);
                const double d = add_rn(yh, -a_j);
                const double e = exp(-mul_rn(d, d) / two_s2);
                for (int k = 0; k < bps; ++k) {
                    if (gray_bit(i, k)) D[k] = add_rn(D[k], e);
                    else N[k] = add_rn(N[k], e);
                }
            }
        } else {
            // as written in the reference: every hypothetical sample is g_inv(n, j) (:656-657)
            const double yh = g_inv_grid(s, sg, gF, gy, npts, nv, j);
            const double sqrt2sigma = sqrt(two_s2);
            // e_coeff does not depend on i (all y_hat[i] are equal): computed once, same operation order
            double e = s.p[j];
            for (int q = 0; q < M; ++q) {
                if (q == j) continue;
                const double pr = mul_rn(add_rn(add_rn(mul_rn(2, yh), -s.a[q]), -a_j), add_rn(s.a[q], -a_j));
                e = add_rn(e, mul_rn(s.p[q], exp(pr / two_s2)));
            }
            const double er = erf(add_rn(yh, -a_j) / sqrt2sigma);
            double S = 0, B = 0;
            for (int i = 0; i < M; ++i) {
                B = add_rn(B, s.delta[i] / e);
                S = add_rn(S, mul_rn(0.5, add_rn(er, -inf_erf[i * M + j])));
            }
            for (int i = 0; i < M; ++i) {
                const double beta = s.delta[i] / e, dFZ = mul_rn(0.5, add_rn(er, -inf_erf[i * M + j]));
                const double A = add_rn(mul_rn(beta, S), -mul_rn(dFZ, B));
                for (int k = 0; k < bps; ++k) {
                    if (gray_bit(i, k)) D[k] = add_rn(D[k], A);
                    else N[k] = add_rn(N[k], A);
                }
            }
        }
        for (int k = 0; k < bps; ++k) lappr[sidx * bps + k] = log(N[k]) - log(D[k]);
    }
}

// ---- mutual-information Monte Carlo: the per-sample terms of
// montecarlo_information (mutual_information.pyx:241-298) for given samples (x_ind, y), summed.
// T(y_hat) = p_x + sum_{m != x} p_m exp((2 y_hat - a_x - a_m)(a_m - a_x) / 2 sigma^2)   (:274-277, :281-284)
__device__ __forceinline__ double mi_inner(const SharedTables &s, int M, double two_s2, double yh, int32_t xi)
{
    const double x = s.a[xi];
    double tmp = s.p[xi];
    for (int q = 0; q < M; ++q) {
        if (q == xi) continue;
        const double pr = mul_rn(add_rn(add_rn(mul_rn(2, yh), -x), -s.a[q]), add_rn(s.a[q], -x));
        tmp = add_rn(tmp, mul_rn(s.p[q], exp(pr / two_s2)));
    }
    return tmp;
}

__global__ void __launch_bounds__(128) k_information(MapperView m, const uint8_t *__restrict__ sign_g,
                                                     const double *__restrict__ gF, const double *__restrict__ gy,
                                                     int32_t npts, const double *__restrict__ fwrd,
                                                     const double *__restrict__ p_Xhat,
                                                     const long long *__restrict__ x_ind,
                                                     const double *__restrict__ y, int64_t n, int which, int mode,
                                                     double *__restrict__ sums)
{
    __shared__ SharedTables s;
    __shared__ uint8_t sg[kMaxOrder];
    __shared__ double red[3][4];
    stage_tables(m, s);
    for (int i = threadIdx.x; i < m.order; i += blockDim.x) sg[i] = sign_g[i];
    __syncthreads();
    const int M = m.order;
    const double two_s2 = 2.0 * m.noise_var;
    double I0 = 0, I1 = 0, I2 = 0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const int32_t xi = checked_index(m, x_ind[j]);
        const double yv = y[j], x = s.a[xi];
        const int32_t xh = hard_decide(s.thr, M, yv);                                   // :241
        const double F = mixture_cdf(s.a, s.p, M, m.s2, yv);                            // :242 (g)
        const double nv = sg[xh] ? add_rn(s.FYt[xh + 1], -F) / s.delta[xh] : add_rn(F, -s.FYt[xh]) / s.delta[xh];
        if (which & 1) I0 += log2(p_Xhat[xh] / fwrd[xi * M + xh]);                      // :257-258
        if (which & 2) {                                                                // :262-268
            double tmp = s.p[xi];
            for (int k = 0; k < M; ++k) {
                if (k == xi) continue;
                const double pr = mul_rn(add_rn(add_rn(mul_rn(2, yv), -s.a[k]), -x), add_rn(s.a[k], -x));
                tmp = add_rn(tmp, mul_rn(s.p[k], exp(pr / two_s2)));
            }
            I1 += log2(tmp);
        }
        if (which & 4) {                                                                // :272-290
            double acc = 0;
            for (int k = 0; k < M; ++k) {
                if (k == xh) continue;
                const double yh = g_inv_grid(s, sg, gF, gy, npts, nv, k);
                acc = add_rn(acc, s.delta[k] / mi_inner(s, M, two_s2, yh, xi));
            }
            const double target = inv_target(s.sign, s.FYt, s.delta, nv, xh);
            const double yh = (mode & 1) ? g_inv_fast(s.a, s.p, s.thr, s.FYt, M, m.sigma, m.s2, target, 1e-9, xh,
                                                      InvTable{m.inv_tab, m.inv_pdf, m.inv_n, m.inv_y0, m.inv_h, m.inv_jump, m.inv_jn})
                                         : g_inv_exact(s.a, s.p, M, m.s2, target, 1e-9);
            acc = mul_rn(acc, mi_inner(s, M, two_s2, yh, xi) / s.delta[xh]);
            acc = add_rn(acc, 1.0);
            acc = mul_rn(acc, p_Xhat[xh]);
            I2 -= log2(acc);
        }
    }
    // block reduction, then one atomic per block and term
    for (int o = 16; o > 0; o >>= 1) {
        I0 += __shfl_down_sync(0xffffffffu, I0, o);
        I1 += __shfl_down_sync(0xffffffffu, I1, o);
        I2 += __shfl_down_sync(0xffffffffu, I2, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[0][w] = I0; red[1][w] = I1; red[2][w] = I2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[threadIdx.x][i];
        atomicAdd(&sums[threadIdx.x], t);
    }
}

static unsigned grid_for(int64_t n, int per_block)
{
    int64_t g = (n + per_block - 1) / per_block;
    return (unsigned)std::min<int64_t>(std::max<int64_t>(g, 1), 148 * 32);
}

}  // namespace qr

extern "C" {

int qr_mapper_set_g_sign(qr_mapper *m, const uint8_t *h_sign_g)
{
    if (!m || !h_sign_g) return qr::fail(QR_ERR_INVALID, "null argument");
    qr::DeviceGuard guard(m->device);
    QR_CUDA_CHECK(cudaMemcpy(m->d_sign_g, h_sign_g, m->order, cudaMemcpyHostToDevice));
    return QR_OK;
}

int qr_mapper_build_grid(qr_mapper *m, double y_low, double y_high, int64_t n_points)
{
    if (!m) return qr::fail(QR_ERR_INVALID, "null mapper");
    if (n_points < 2 || n_points > (int64_t(1) << 30) || !(y_high > y_low))
        return qr::fail(QR_ERR_INVALID, "bad grid");
    qr::DeviceGuard guard(m->device);
    cudaFree(m->grid_y);
    m->grid_y = m->grid_F = nullptr;
    m->grid_n = 0;
    QR_CUDA_CHECK(cudaMalloc((void **)&m->grid_y, 2 * (size_t)n_points * sizeof(double)));
    m->grid_F = m->grid_y + n_points;
    m->grid_n = (int32_t)n_points;
    qr::k_grid<<<(unsigned)((n_points + 255) / 256), 256>>>(qr::mapper_view(m), y_low, y_high, (int32_t)n_points,
                                                            m->grid_y, m->grid_F);
    QR_CUDA_CHECK(cudaGetLastError());
    QR_CUDA_CHECK(cudaDeviceSynchronize());
    return QR_OK;
}

int qr_mapper_grid(const qr_mapper *m, int64_t *n_points, double *h_y_range, double *h_F_Y)
{
    if (!m) return qr::fail(QR_ERR_INVALID, "null mapper");
    if (n_points) *n_points = m->grid_n;
    if (!m->grid_n) return (h_y_range || h_F_Y) ? qr::fail(QR_ERR_INVALID, "grid not built") : QR_OK;
    qr::DeviceGuard guard(m->device);
    if (h_y_range) QR_CUDA_CHECK(cudaMemcpy(h_y_range, m->grid_y, m->grid_n * sizeof(double), cudaMemcpyDeviceToHost));
    if (h_F_Y) QR_CUDA_CHECK(cudaMemcpy(h_F_Y, m->grid_F, m->grid_n * sizeof(double), cudaMemcpyDeviceToHost));
    return QR_OK;
}

int qr_F_Y(const qr_mapper *m, const double *d_y, int64_t n, double *d_out, void *stream)
{
    if (!m) return qr::fail(QR_ERR_INVALID, "null mapper");
    if (n < 0) return qr::fail(QR_ERR_INVALID, "negative length");
    if (n == 0) return QR_OK;
    if (!d_y || !d_out) return qr::fail(QR_ERR_INVALID, "null array");
    qr::DeviceGuard guard(m->device);
    qr::k_F_Y<<<qr::grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(qr::mapper_view(m), d_y, n, d_out);
    QR_CUDA_CHECK(cudaGetLastError());
    return QR_OK;
}

int qr_F_Z(const double *d_z, int64_t n, double mu, double sigma, double *d_out, void *stream)
{
    if (n < 0) return qr::fail(QR_ERR_INVALID, "negative length");
    if (n == 0) return QR_OK;
    if (!d_z || !d_out) return qr::fail(QR_ERR_INVALID, "null array");
    qr::k_F_Z<<<qr::grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_z, n, mu, sqrt(2.0) * sigma, d_out);
    QR_CUDA_CHECK(cudaGetLastError());
    return QR_OK;
}

int qr_demap_noise(const qr_mapper *m, const double *d_n_hat, const int64_t *d_symb, int64_t n, double *d_y_hat,
                   void *stream)
{
    if (!m) return qr::fail(QR_ERR_INVALID, "null mapper");
    if (!m->grid_n) return qr::fail(QR_ERR_INVALID, "F_Y grid not built (qr_mapper_build_grid)");
    if (n < 0) return qr::fail(QR_ERR_INVALID, "negative length");
    if (n == 0) return QR_OK;
    if (!d_n_hat || !d_symb || !d_y_hat) return qr::fail(QR_ERR_INVALID, "null array");
    qr::DeviceGuard guard(m->device);
    qr::k_demap_noise<<<qr::grid_for(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        qr::mapper_view(m), m->d_sign_g, m->grid_F, m->grid_y, m->grid_n, d_n_hat,
        reinterpret_cast<const long long *>(d_symb), n, d_y_hat);
    QR_CUDA_CHECK(cudaGetLastError());
    return QR_OK;
}

int qr_demap_lappr_variant(const qr_mapper *m, int variant, const double *d_n_hat, const int64_t *d_tx_index,
                           int64_t n, double *d_llr, void *stream)
{
    if (!m) return qr::fail(QR_ERR_INVALID, "null mapper");
    if (variant != 1 && variant != 2) return qr::fail(QR_ERR_INVALID, "variant must be 1 (simplified) or 2 (sofisticated)");
    if (!m->grid_n) return qr::fail(QR_ERR_INVALID, "F_Y grid not built (qr_mapper_build_grid)");
    if (n < 0) return qr::fail(QR_ERR_INVALID, "negative length");
    if (n == 0) return QR_OK;
    if (!d_n_hat || !d_tx_index || !d_llr) return qr::fail(QR_ERR_INVALID, "null array");
    qr::DeviceGuard guard(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const auto tx = reinterpret_cast<const long long *>(d_tx_index);
    if (variant == 1)
        qr::k_demap_variant<1><<<qr::grid_for(n, 128), 128, 0, st>>>(qr::mapper_view(m), m->d_sign_g, m->grid_F, m->grid_y,
                                                                      m->grid_n, m->inf_erf, d_n_hat, tx, n, d_llr);
    else
        qr::k_demap_variant<2><<<qr::grid_for(n, 128), 128, 0, st>>>(qr::mapper_view(m), m->d_sign_g, m->grid_F, m->grid_y,
                                                                      m->grid_n, m->inf_erf, d_n_hat, tx, n, d_llr);
    QR_CUDA_CHECK(cudaGetLastError());
    return QR_OK;
}

int qr_information_sums(const qr_mapper *m, const double *d_p_Xhat, const int64_t *d_x_ind, const double *d_y,
                        int64_t n, int which, int demap_mode, double *d_sums, void *stream)
{
    if (!m) return qr::fail(QR_ERR_INVALID, "null mapper");
    if (n < 0) return qr::fail(QR_ERR_INVALID, "negative length");
    if (which < 0 || which > 7) return qr::fail(QR_ERR_INVALID, "which is a 3-bit mask");
    if ((which & 4) && !m->grid_n) return qr::fail(QR_ERR_INVALID, "F_Y grid not built (qr_mapper_build_grid)");
    if (n == 0) return QR_OK;
    if (!d_p_Xhat || !d_x_ind || !d_y || !d_sums) return qr::fail(QR_ERR_INVALID, "null array");
    qr::DeviceGuard guard(m->device);
    qr::k_information<<<qr::grid_for(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        qr::mapper_view(m), m->d_sign_g, m->grid_F, m->grid_y, m->grid_n, m->fwrd, d_p_Xhat,
        reinterpret_cast<const long long *>(d_x_ind), d_y, n, which, demap_mode, d_sums);
    QR_CUDA_CHECK(cudaGetLastError());
    return QR_OK;
}

}  // extern "C"
