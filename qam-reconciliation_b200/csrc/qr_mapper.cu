// Noise-mapper kernels: one thread per symbol, fp64, alphabet tables staged in shared memory.
// (reference: qamreconciliation/noisemapper.pyx; see qr_mapper_core.cuh for the arithmetic.)
#include <cuda_runtime.h>

#include <cmath>
#include <new>

#include "qr_handles.h"
#include "qr_mapper_core.cuh"
#include "qr_mapper_device.cuh"

namespace qr {

static MapperView view_of(const qr_mapper *m)
{
    MapperView v;
    v.order = m->order; v.bps = m->bps;
    v.noise_var = m->noise_var; v.sigma = m->sigma; v.s2 = m->s2;
    v.constellation = m->constellation; v.thresholds = m->thresholds; v.probabilities = m->probabilities;
    v.sign_config = m->d_sign; v.sign_g = m->d_sign_g; v.FY_thr = m->FY_thr; v.delta = m->delta; v.bare = m->bare;
    v.inv_tab = m->inv_tab; v.inv_pdf = m->inv_pdf; v.inv_n = m->inv_n; v.inv_y0 = m->inv_y0; v.inv_h = m->inv_h;
    v.inv_jump = m->inv_jump; v.inv_jn = m->inv_jn;
    v.uniform = m->uniform;
    v.inv32_F = m->inv32_F; v.inv32_f = m->inv32_f; v.inv32_jump = m->inv32_jump; v.inv32_h = m->inv32_h;
    v.index_errors = m->d_index_errors;
    return v;
}

MapperView mapper_view(const qr_mapper *m) { return view_of(m); }

// NoiseMapper.__cinit__ tables (noisemapper.pyx:149-235), one thread: tiny and done once
__global__ void k_mapper_tables(MapperView m, double *FY_thr, double *delta, double *fwrd, double *back,
                                double *bare, double *inf_erf)
{
    if (threadIdx.x || blockIdx.x) return;
    const int M = m.order, bps = m.bps;
    const double *a = m.constellation, *thr = m.thresholds, *p = m.probabilities;
    FY_thr[0] = 0;
    FY_thr[M] = 1;
    for (int i = 1; i < M; ++i) FY_thr[i] = mixture_cdf(a, p, M, m.s2, thr[i]);
    for (int i = 0; i < M; ++i) delta[i] = add_rn(FY_thr[i + 1], -FY_thr[i]);
    for (int j = 0; j < M; ++j) {
        fwrd[j * M + 0] = 0.5 * (erf((thr[1] - a[j]) / m.s2) + 1);
        fwrd[j * M + M - 1] = 0.5 * (1 - erf((thr[M - 1] - a[j]) / m.s2));
        for (int i = 1; i < M - 1; ++i)
            fwrd[j * M + i] = 0.5 * add_rn(erf((thr[i + 1] - a[j]) / m.s2), -erf((thr[i] - a[j]) / m.s2));
    }
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) {
            double tot = 0;
            for (int k = 0; k < M; ++k) tot = add_rn(tot, mul_rn(p[k], fwrd[k * M + i]));
            back[i * M + j] = mul_rn(p[j], fwrd[j * M + i]) / tot;
        }
    for (int j = 0; j < M; ++j)
        for (int k = 0; k < bps; ++k) {
            double num = 0, den = 0;
            for (int i = 0; i < M; ++i) {
                if (gray_bit(i, k)) den = add_rn(den, fwrd[j * M + i]);
                else num = add_rn(num, fwrd[j * M + i]);
            }
            bare[j * bps + k] = (den == 0) ? 1e300 : log(num / den);
        }
    for (int j = 0; j < M; ++j) {
        inf_erf[0 * M + j] = -1;
        for (int i = 1; i < M; ++i) inf_erf[i * M + j] = erf((thr[i] - a[j]) / m.s2);
    }
}

// F_Y on a uniform grid (starting points of the fast inverse); one thread per grid point
__global__ void k_fill_inv_table(MapperView m, double *tab, double *pdf, int32_t n, double y0, double h)
{
    const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) {
        tab[j] = mixture_cdf(m.constellation, m.probabilities, m.order, m.s2, y0 + j * h);
        pdf[j] = mixture_pdf(m.constellation, m.probabilities, m.order, m.sigma, y0 + j * h);
    }
}

__global__ void k_fill_inv32(MapperView m, double *F, float *f, double y0, double h)
{
    const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < kInv32N) {
        F[j] = mixture_cdf(m.constellation, m.probabilities, m.order, m.s2, y0 + j * h);
        f[j] = (float)mixture_pdf(m.constellation, m.probabilities, m.order, m.sigma, y0 + j * h);
    }
}

__global__ void k_fill_jump32(const double *__restrict__ F, uint16_t *__restrict__ jump)
{
    const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > kInv32J + 1) return;
    const double v = (double)t / (double)kInv32J;
    int32_t lo = 0, hi = kInv32N;           // invariant: F[lo] <= v (or lo == 0), F[hi] > v (or hi == N)
    while (hi - lo > 1) {
        const int32_t mid = (lo + hi) >> 1;
        if (F[mid] <= v) lo = mid; else hi = mid;
    }
    jump[t] = (uint16_t)lo;
}

// jump[t] = largest grid index g with tab[g] <= t / jn (0 if none): see InvTable::jump
__global__ void k_fill_jump_table(const double *__restrict__ tab, int32_t n, int32_t *__restrict__ jump, int32_t jn)
{
    const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > jn) return;
    const double v = (double)t / (double)jn;
    int32_t lo = 0, hi = n;           // invariant: tab[lo] <= v (or lo == 0), tab[hi] > v (or hi == n)
    while (hi - lo > 1) {
        const int32_t mid = (lo + hi) >> 1;
        if (tab[mid] <= v) lo = mid; else hi = mid;
    }
    jump[t] = lo;
}

// hard decision (+ softening metric) (+ Gray bits) in one pass over y
// (noisemapper.pyx:349-359, :373-388; alphabet.pyx:98-107).  idx_in != NULL: use the given indices
// instead of deciding (map_noise / demap_symbols_to_bits on caller-provided indices).
__global__ void __launch_bounds__(256) k_front_end(MapperView m, const double *__restrict__ y,
                                                   const long long *__restrict__ idx_in, int64_t n,
                                                   long long *__restrict__ idx_out,
                                                   double *__restrict__ n_hat, uint8_t *__restrict__ bits)
{
    __shared__ SharedTables s;
    stage_tables(m, s);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n;
         j += (int64_t)gridDim.x * blockDim.x) {
        int32_t i;
        if (idx_in) i = checked_index(m, idx_in[j]);
        else i = hard_decide(s.thr, m.order, y[j]);
        if (idx_out) idx_out[j] = i;
        if (n_hat) {
            const double F = mixture_cdf(s.a, s.p, m.order, m.s2, y[j]);
            n_hat[j] = m.sign_g[i] ? add_rn(s.FYt[i + 1], -F) / s.delta[i] : add_rn(F, -s.FYt[i]) / s.delta[i];
        }
        if (bits) {
            for (int k = 0; k < m.bps; ++k) bits[j * m.bps + k] = gray_bit(i, k);
        }
    }
}

static unsigned grid_for_elems(int64_t n, int per_block = 256, int cap = 148 * 32);

template <typename OUT>
__global__ void __launch_bounds__(128) k_demap(MapperView m, const double *__restrict__ n_hat,
                                               const long long *__restrict__ tx, int64_t n, int mode,
                                               double alpha, OUT *__restrict__ llr)
{
    __shared__ SharedTables s;
    stage_tables(m, s);
    TablesRef t = tables_ref(s);
    if (!m.uniform) t.ghi = t.glo = nullptr;
    for (int64_t sidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; sidx < n;
         sidx += (int64_t)gridDim.x * blockDim.x) {
        double out[kMaxBps];
        demap_symbol(m, t, n_hat[sidx], checked_index(m, tx[sidx]), mode, alpha, out);
        for (int k = 0; k < m.bps; ++k) llr[sidx * m.bps + k] = (OUT)out[k];
    }
}

// fp32-grade demapper (QR_DEMAP_FAST | QR_DEMAP_F32GRADE), one instantiation per alphabet size: everything unrolled,
// accumulators in registers, the rare slow paths out of line
template <typename OUT, int BPS>
__global__ void __launch_bounds__(256) k_demap32(MapperView m, const double *__restrict__ n_hat,
                                                 const long long *__restrict__ tx, int64_t n, int corrected,
                                                 double alpha, OUT *__restrict__ llr)
{
    __shared__ SharedTables s;
    __shared__ double s_F[kInv32N];
    __shared__ float s_f[kInv32N];
    __shared__ uint16_t s_jump[kInv32J + 2];
    for (int i = threadIdx.x; i < kInv32N; i += blockDim.x) { s_F[i] = m.inv32_F[i]; s_f[i] = m.inv32_f[i]; }
    for (int i = threadIdx.x; i < kInv32J + 2; i += blockDim.x) s_jump[i] = m.inv32_jump[i];
    stage_tables(m, s);
    TablesRef t = tables_ref(s);
    t.t32 = InvTable32{s_F, s_f, s_jump, m.inv_y0, m.inv32_h};
    for (int64_t sidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; sidx < n;
         sidx += (int64_t)gridDim.x * blockDim.x) {
        double out[BPS];
        demap_symbol_f32grade<BPS>(m, t, n_hat[sidx], checked_index(m, tx[sidx]), corrected != 0, alpha, out);
        if constexpr (BPS == 2 && sizeof(OUT) == 4) {
            *reinterpret_cast<float2 *>(llr + sidx * 2) = make_float2((float)out[0], (float)out[1]);
        } else {
#pragma unroll
            for (int k = 0; k < BPS; ++k) llr[sidx * BPS + k] = (OUT)out[k];
        }
    }
}

template <typename OUT>
static bool launch_demap32(const qr_mapper *m, const double *n_hat, const long long *tx, int64_t n, int mode, double alpha,
                           OUT *llr, cudaStream_t st)
{
    if (!m->uniform || m->bps > 4) return false;
    const unsigned grid = grid_for_elems(n, 256, 148 * 16);
    const int corr = (mode & QR_DEMAP_CORRECTED) != 0;
    switch (m->bps) {
    case 1: k_demap32<OUT, 1><<<grid, 256, 0, st>>>(view_of(m), n_hat, tx, n, corr, alpha, llr); break;
    case 2: k_demap32<OUT, 2><<<grid, 256, 0, st>>>(view_of(m), n_hat, tx, n, corr, alpha, llr); break;
    case 3: k_demap32<OUT, 3><<<grid, 256, 0, st>>>(view_of(m), n_hat, tx, n, corr, alpha, llr); break;
    default: k_demap32<OUT, 4><<<grid, 256, 0, st>>>(view_of(m), n_hat, tx, n, corr, alpha, llr); break;
    }
    return true;
}

__global__ void __launch_bounds__(128) k_g_inv(MapperView m, const double *__restrict__ n_hat,
                                               const long long *__restrict__ region, int64_t n, int mode,
                                               double *__restrict__ y_hat)
{
    __shared__ SharedTables s;
    stage_tables(m, s);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n;
         j += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = checked_index(m, region[j]);
        const double target = inv_target(s.sign, s.FYt, s.delta, n_hat[j], i);
        y_hat[j] = (mode & 1) ? g_inv_fast(s.a, s.p, s.thr, s.FYt, m.order, m.sigma, m.s2, target, 1e-9, i, InvTable{m.inv_tab, m.inv_pdf, m.inv_n, m.inv_y0, m.inv_h, m.inv_jump, m.inv_jn})
                              : g_inv_exact(s.a, s.p, m.order, m.s2, target, 1e-9);
    }
}

template <typename OUT>
__global__ void __launch_bounds__(256) k_bare_llr(MapperView m, const long long *__restrict__ tx,
                                                  int64_t n, OUT *__restrict__ llr)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n;
         j += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = checked_index(m, tx[j]);
        for (int k = 0; k < m.bps; ++k) llr[j * m.bps + k] = (OUT)m.bare[i * m.bps + k];
    }
}

template <typename OUT>
__global__ void __launch_bounds__(256) k_direct_llr(MapperView m, const double *__restrict__ y, int64_t n,
                                                    double two_variance, OUT *__restrict__ llr)
{
    __shared__ double a[kMaxOrder];
    for (int i = threadIdx.x; i < m.order; i += blockDim.x) a[i] = m.constellation[i];
    __syncthreads();
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n;
         j += (int64_t)gridDim.x * blockDim.x) {
        double out[kMaxBps];
        direct_llr(a, m.order, m.bps, two_variance, y[j], out);
        for (int k = 0; k < m.bps; ++k) llr[j * m.bps + k] = (OUT)out[k];
    }
}

static unsigned grid_for_elems(int64_t n, int per_block, int cap)
{
    int64_t g = (n + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (unsigned)g;
}

}  // namespace qr

extern "C" {

int qr_mapper_create(int bits_per_symbol, const double *h_constellation, const double *h_thresholds,
                     const double *h_probabilities, double noise_var, const uint8_t *h_sign_config,
                     int device, qr_mapper **out)
{
    if (!out) return qr::fail(QR_ERR_INVALID, "null output pointer");
    *out = nullptr;
    if (bits_per_symbol < 1 || bits_per_symbol > qr::kMaxBps)
        return qr::fail(QR_ERR_INVALID, "bits per symbol must be in 1..8");
    if (!(noise_var > 0)) return qr::fail(QR_ERR_INVALID, "noise variance must be strictly positive");
    if (!h_constellation || !h_thresholds || !h_probabilities)
        return qr::fail(QR_ERR_INVALID, "null alphabet array");
    if (device < 0) return qr::fail(QR_ERR_INVALID, "the noise mapper needs a CUDA device");
    qr_mapper *m = new (std::nothrow) qr_mapper();
    if (!m) return qr::fail(QR_ERR_NOMEM, "out of host memory");
    const int M = 1 << bits_per_symbol;
    m->device = device; m->order = M; m->bps = bits_per_symbol;
    m->noise_var = noise_var;
    m->sigma = sqrt(noise_var);
    m->s2 = sqrt(2.0) * m->sigma;
    m->uniform = 1;
    for (int i = 2; i < M; ++i)
        if (fabs((h_constellation[i] - h_constellation[i - 1]) - (h_constellation[1] - h_constellation[0])) >
            1e-12 * fabs(h_constellation[1] - h_constellation[0])) m->uniform = 0;
    qr::DeviceGuard guard(device);
    auto body = [&]() -> int {
        const size_t nd = (size_t)M + (M + 1) + M + (M + 1) + M + 3 * (size_t)M * M + (size_t)M * bits_per_symbol;
        m->n_table_doubles = nd;
        QR_CUDA_CHECK(cudaMalloc((void **)&m->d_tables, nd * sizeof(double)));
        QR_CUDA_CHECK(cudaMalloc((void **)&m->d_sign, M));
        QR_CUDA_CHECK(cudaMalloc((void **)&m->d_sign_g, M));
        QR_CUDA_CHECK(cudaMalloc((void **)&m->d_index_errors, sizeof(int32_t)));
        QR_CUDA_CHECK(cudaMemset(m->d_index_errors, 0, sizeof(int32_t)));
        double *p = m->d_tables;
        m->constellation = p; p += M;
        m->thresholds = p; p += M + 1;
        m->probabilities = p; p += M;
        m->FY_thr = p; p += M + 1;
        m->delta = p; p += M;
        m->fwrd = p; p += (size_t)M * M;
        m->back = p; p += (size_t)M * M;
        m->inf_erf = p; p += (size_t)M * M;
        m->bare = p;
        QR_CUDA_CHECK(cudaMemcpy(m->constellation, h_constellation, M * sizeof(double), cudaMemcpyHostToDevice));
        QR_CUDA_CHECK(cudaMemcpy(m->thresholds, h_thresholds, (M + 1) * sizeof(double), cudaMemcpyHostToDevice));
        QR_CUDA_CHECK(cudaMemcpy(m->probabilities, h_probabilities, M * sizeof(double), cudaMemcpyHostToDevice));
        if (h_sign_config) QR_CUDA_CHECK(cudaMemcpy(m->d_sign, h_sign_config, M, cudaMemcpyHostToDevice));
        else QR_CUDA_CHECK(cudaMemset(m->d_sign, 0, M));
        QR_CUDA_CHECK(cudaMemcpy(m->d_sign_g, m->d_sign, M, cudaMemcpyDeviceToDevice));
        qr::k_mapper_tables<<<1, 32>>>(qr::view_of(m), m->FY_thr, m->delta, m->fwrd, m->back, m->bare, m->inf_erf);
        QR_CUDA_CHECK(cudaGetLastError());
        // inverse-CDF starting table: +-9 sigma around the constellation, ~2e-3 sigma-free grid step
        m->inv_n = 16385;
        m->inv_y0 = h_constellation[0] - 9.0 * m->sigma;
        m->inv_h = (h_constellation[M - 1] + 9.0 * m->sigma - m->inv_y0) / (m->inv_n - 1);
        QR_CUDA_CHECK(cudaMalloc((void **)&m->inv_tab, 2 * (size_t)m->inv_n * sizeof(double)));
        m->inv_pdf = m->inv_tab + m->inv_n;
        {
            qr::MapperView v = qr::view_of(m);
            v.inv_tab = nullptr;
            qr::k_fill_inv_table<<<(m->inv_n + 255) / 256, 256>>>(v, m->inv_tab, m->inv_pdf, m->inv_n, m->inv_y0, m->inv_h);
        }
        QR_CUDA_CHECK(cudaGetLastError());
        m->inv_jn = 8192;
        QR_CUDA_CHECK(cudaMalloc((void **)&m->inv_jump, (size_t)(m->inv_jn + 2) * sizeof(int32_t)));
        qr::k_fill_jump_table<<<(m->inv_jn + 1 + 255) / 256, 256>>>(m->inv_tab, m->inv_n, m->inv_jump, m->inv_jn);
        QR_CUDA_CHECK(cudaGetLastError());
        // coarse copy for the fp32-grade demapper (same span, kInv32N points)
        m->inv32_h = (h_constellation[M - 1] + 9.0 * m->sigma - m->inv_y0) / (qr::kInv32N - 1);
        QR_CUDA_CHECK(cudaMalloc((void **)&m->inv32_F, qr::kInv32N * sizeof(double)));
        QR_CUDA_CHECK(cudaMalloc((void **)&m->inv32_f, qr::kInv32N * sizeof(float)));
        QR_CUDA_CHECK(cudaMalloc((void **)&m->inv32_jump, (qr::kInv32J + 2) * sizeof(uint16_t)));
        {
            qr::MapperView v = qr::view_of(m);
            qr::k_fill_inv32<<<(qr::kInv32N + 255) / 256, 256>>>(v, m->inv32_F, m->inv32_f, m->inv_y0, m->inv32_h);
            QR_CUDA_CHECK(cudaGetLastError());
            qr::k_fill_jump32<<<(qr::kInv32J + 2 + 255) / 256, 256>>>(m->inv32_F, m->inv32_jump);
            QR_CUDA_CHECK(cudaGetLastError());
        }
        QR_CUDA_CHECK(cudaDeviceSynchronize());
        return QR_OK;
    };
    int rc = body();
    if (rc != QR_OK) { qr_mapper_destroy(m); return rc; }
    *out = m;
    return QR_OK;
}

void qr_mapper_destroy(qr_mapper *m)
{
    if (!m) return;
    {
        qr::DeviceGuard guard(m->device);
        cudaFree(m->d_tables);
        cudaFree(m->d_sign);
        cudaFree(m->d_sign_g);
        cudaFree(m->d_index_errors);
        cudaFree(m->grid_y);
        cudaFree(m->inv_tab);
        cudaFree(m->inv_jump);
        cudaFree(m->inv32_F); cudaFree(m->inv32_f); cudaFree(m->inv32_jump);
    }
    delete m;
}

int qr_mapper_index_errors(qr_mapper *m, int64_t *count, void *stream)
{
    if (!m || !count) return qr::fail(QR_ERR_INVALID, "null pointer");
    qr::DeviceGuard guard(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int32_t h = 0;
    QR_CUDA_CHECK(cudaMemcpyAsync(&h, m->d_index_errors, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    QR_CUDA_CHECK(cudaStreamSynchronize(st));
    if (h) {
        QR_CUDA_CHECK(cudaMemsetAsync(m->d_index_errors, 0, sizeof(int32_t), st));
        QR_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    *count = h;
    return QR_OK;
}

int qr_mapper_tables(const qr_mapper *m, double *F_Y_thresholds, double *delta_F_Y, double *fwrd,
                     double *back, double *bare_llr_table, double *inf_erf_table)
{
    if (!m) return qr::fail(QR_ERR_INVALID, "null mapper");
    qr::DeviceGuard guard(m->device);
    const int M = m->order;
    if (F_Y_thresholds) QR_CUDA_CHECK(cudaMemcpy(F_Y_thresholds, m->FY_thr, (M + 1) * sizeof(double), cudaMemcpyDeviceToHost));
    if (delta_F_Y) QR_CUDA_CHECK(cudaMemcpy(delta_F_Y, m->delta, M * sizeof(double), cudaMemcpyDeviceToHost));
    if (fwrd) QR_CUDA_CHECK(cudaMemcpy(fwrd, m->fwrd, (size_t)M * M * sizeof(double), cudaMemcpyDeviceToHost));
    if (back) QR_CUDA_CHECK(cudaMemcpy(back, m->back, (size_t)M * M * sizeof(double), cudaMemcpyDeviceToHost));
    if (bare_llr_table) QR_CUDA_CHECK(cudaMemcpy(bare_llr_table, m->bare, (size_t)M * m->bps * sizeof(double), cudaMemcpyDeviceToHost));
    if (inf_erf_table) QR_CUDA_CHECK(cudaMemcpy(inf_erf_table, m->inf_erf, (size_t)M * M * sizeof(double), cudaMemcpyDeviceToHost));
    return QR_OK;
}

static int front(const qr_mapper *m, const double *d_y, const int64_t *d_idx_in, int64_t n,
                 int64_t *d_idx_out, double *d_n_hat, uint8_t *d_bits, void *stream)
{
    if (!m) return qr::fail(QR_ERR_INVALID, "null mapper");
    if (n < 0) return qr::fail(QR_ERR_INVALID, "negative length");
    if (n == 0) return QR_OK;
    if ((!d_y && (!d_idx_in || d_n_hat))) return qr::fail(QR_ERR_INVALID, "null sample array");
    qr::DeviceGuard guard(m->device);
    qr::k_front_end<<<qr::grid_for_elems(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        qr::view_of(m), d_y, reinterpret_cast<const long long *>(d_idx_in), n,
        reinterpret_cast<long long *>(d_idx_out), d_n_hat, d_bits);
    QR_CUDA_CHECK(cudaGetLastError());
    return QR_OK;
}

int qr_hard_decide_index(const qr_mapper *m, const double *d_y, int64_t n, int64_t *d_index, void *stream)
{
    if (n > 0 && !d_index) return qr::fail(QR_ERR_INVALID, "null output");
    return front(m, d_y, nullptr, n, d_index, nullptr, nullptr, stream);
}

int qr_symbols_to_bits(const qr_mapper *m, const int64_t *d_index, int64_t n, uint8_t *d_bits, void *stream)
{
    if (n > 0 && (!d_index || !d_bits)) return qr::fail(QR_ERR_INVALID, "null array");
    return front(m, nullptr, d_index, n, nullptr, nullptr, d_bits, stream);
}

int qr_map_noise(const qr_mapper *m, const double *d_y, const int64_t *d_index, int64_t n,
                 double *d_n_hat, void *stream)
{
    if (n > 0 && (!d_index || !d_n_hat)) return qr::fail(QR_ERR_INVALID, "null array");
    return front(m, d_y, d_index, n, nullptr, d_n_hat, nullptr, stream);
}

int qr_front_end(const qr_mapper *m, const double *d_y, int64_t n, int64_t *d_index, double *d_n_hat,
                 uint8_t *d_bits, void *stream)
{
    return front(m, d_y, nullptr, n, d_index, d_n_hat, d_bits, stream);
}

int qr_demap_lappr(const qr_mapper *m, const double *d_n_hat, const int64_t *d_tx_index, int64_t n,
                   int mode, double alpha, void *d_llr, int llr_dtype, void *stream)
{
    if (!m) return qr::fail(QR_ERR_INVALID, "null mapper");
    if (n < 0) return qr::fail(QR_ERR_INVALID, "negative length");
    if (llr_dtype != QR_F32 && llr_dtype != QR_F64) return qr::fail(QR_ERR_INVALID, "bad llr dtype");
    if (mode < 0 || mode > 7 || ((mode & QR_DEMAP_F32GRADE) && !(mode & QR_DEMAP_FAST)))
        return qr::fail(QR_ERR_INVALID, "bad demap mode");
    if (n == 0) return QR_OK;
    if (!d_n_hat || !d_tx_index || !d_llr) return qr::fail(QR_ERR_INVALID, "null array");
    qr::DeviceGuard guard(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long *tx = reinterpret_cast<const long long *>(d_tx_index);
    if ((mode & QR_DEMAP_F32GRADE) &&
        (llr_dtype == QR_F64 ? qr::launch_demap32<double>(m, d_n_hat, tx, n, mode, alpha, static_cast<double *>(d_llr), st)
                             : qr::launch_demap32<float>(m, d_n_hat, tx, n, mode, alpha, static_cast<float *>(d_llr), st))) {
        QR_CUDA_CHECK(cudaGetLastError());
        return QR_OK;
    }
    const unsigned grid = qr::grid_for_elems(n, 128, 148 * 64);
    if (llr_dtype == QR_F64)
        qr::k_demap<double><<<grid, 128, 0, st>>>(qr::view_of(m), d_n_hat, reinterpret_cast<const long long *>(d_tx_index),
                                                   n, mode, alpha, static_cast<double *>(d_llr));
    else
        qr::k_demap<float><<<grid, 128, 0, st>>>(qr::view_of(m), d_n_hat, reinterpret_cast<const long long *>(d_tx_index),
                                                  n, mode, alpha, static_cast<float *>(d_llr));
    QR_CUDA_CHECK(cudaGetLastError());
    return QR_OK;
}

int qr_g_inv_search(const qr_mapper *m, const double *d_n_hat, const int64_t *d_region, int64_t n,
                    int mode, double *d_y_hat, void *stream)
{
    if (!m) return qr::fail(QR_ERR_INVALID, "null mapper");
    if (n < 0) return qr::fail(QR_ERR_INVALID, "negative length");
    if (n == 0) return QR_OK;
    if (!d_n_hat || !d_region || !d_y_hat) return qr::fail(QR_ERR_INVALID, "null array");
    qr::DeviceGuard guard(m->device);
    qr::k_g_inv<<<qr::grid_for_elems(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        qr::view_of(m), d_n_hat, reinterpret_cast<const long long *>(d_region), n, mode, d_y_hat);
    QR_CUDA_CHECK(cudaGetLastError());
    return QR_OK;
}

int qr_bare_llr(const qr_mapper *m, const int64_t *d_tx_index, int64_t n, void *d_llr, int llr_dtype,
                void *stream)
{
    if (!m) return qr::fail(QR_ERR_INVALID, "null mapper");
    if (n < 0) return qr::fail(QR_ERR_INVALID, "negative length");
    if (llr_dtype != QR_F32 && llr_dtype != QR_F64) return qr::fail(QR_ERR_INVALID, "bad llr dtype");
    if (n == 0) return QR_OK;
    if (!d_tx_index || !d_llr) return qr::fail(QR_ERR_INVALID, "null array");
    qr::DeviceGuard guard(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (llr_dtype == QR_F64)
        qr::k_bare_llr<double><<<qr::grid_for_elems(n), 256, 0, st>>>(qr::view_of(m), reinterpret_cast<const long long *>(d_tx_index), n, static_cast<double *>(d_llr));
    else
        qr::k_bare_llr<float><<<qr::grid_for_elems(n), 256, 0, st>>>(qr::view_of(m), reinterpret_cast<const long long *>(d_tx_index), n, static_cast<float *>(d_llr));
    QR_CUDA_CHECK(cudaGetLastError());
    return QR_OK;
}

int qr_direct_llr(const qr_mapper *m, const double *d_y, int64_t n, double two_variance, void *d_llr,
                  int llr_dtype, void *stream)
{
    if (!m) return qr::fail(QR_ERR_INVALID, "null mapper");
    if (n < 0) return qr::fail(QR_ERR_INVALID, "negative length");
    if (llr_dtype != QR_F32 && llr_dtype != QR_F64) return qr::fail(QR_ERR_INVALID, "bad llr dtype");
    if (n == 0) return QR_OK;
    if (!d_y || !d_llr) return qr::fail(QR_ERR_INVALID, "null array");
    qr::DeviceGuard guard(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (llr_dtype == QR_F64)
        qr::k_direct_llr<double><<<qr::grid_for_elems(n), 256, 0, st>>>(qr::view_of(m), d_y, n, two_variance, static_cast<double *>(d_llr));
    else
        qr::k_direct_llr<float><<<qr::grid_for_elems(n), 256, 0, st>>>(qr::view_of(m), d_y, n, two_variance, static_cast<float *>(d_llr));
    QR_CUDA_CHECK(cudaGetLastError());
    return QR_OK;
}

}  // extern "C"
