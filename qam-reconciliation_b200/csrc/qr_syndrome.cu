// Integer kernels around the decoder: syndrome evaluation, word / LLR consistency, error counting.
//   Matrix.eval_syndrome      reference matrix.pyx:55-60
//   Decoder.check_word        reference decoder.pyx:211-232
//   Decoder.check_lappr       reference decoder.pyx:235-281
//   count_errors_from_lappr   reference utils.pyx:27-40
// All operate on caller layout [frames][N] / [frames][C]; one thread per (frame, check), checks
// fastest, so the per-check index reads coalesce and a frame's word (N bytes) stays in L1.
#include <cuda_runtime.h>

#include "qr_handles.h"

namespace qr {

__global__ void k_eval_syndrome(const int32_t *__restrict__ chk_ptr, const int32_t *__restrict__ slot_var,
                                const int32_t *__restrict__ chk_order, int64_t N, int64_t C,
                                const uint8_t *__restrict__ word, uint8_t *__restrict__ synd)
{
    const int64_t b = blockIdx.y;
    const int64_t ci = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= C) return;
    const uint8_t *w = word + b * N;
    uint8_t acc = 0;
    for (int32_t s = chk_ptr[ci]; s < chk_ptr[ci + 1]; ++s) acc ^= w[slot_var[s]];
    synd[b * C + chk_order[ci]] = acc;
}

// The same, one CTA per frame with the frame's word packed to BITS in shared memory (N / 8 bytes): the per-check
// gathers become shared-memory reads instead of one L1 wavefront per byte (the byte gathers of k_eval_syndrome are
// L1-wavefront bound: 2.05 ms for 4096 frames of config 2, against 0.07 ms of HBM time for its bytes).
__global__ void __launch_bounds__(512) k_eval_syndrome_smem(const int32_t *__restrict__ chk_ptr, const int32_t *__restrict__ slot_var,
                                                            const int32_t *__restrict__ chk_order, int64_t N, int64_t C,
                                                            const uint8_t *__restrict__ word, uint8_t *__restrict__ synd)
{
    extern __shared__ uint8_t s_bits[];
    const int64_t b = blockIdx.x;
    const uint8_t *w = word + b * N;
    const int64_t nb = (N + 7) / 8;
    int not_bits = 0;                     // a byte other than 0 / 1: the XOR of BYTES is asked for (matrix.pyx:58)
    if ((reinterpret_cast<uintptr_t>(w) & 7) == 0) {
        const int64_t full = N / 8;
        for (int64_t i = threadIdx.x; i < full; i += blockDim.x) {
            const unsigned long long q = reinterpret_cast<const unsigned long long *>(w)[i];
            not_bits |= (q & ~0x0101010101010101ull) != 0;
            s_bits[i] = (uint8_t)(((q & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56);   // byte k's bit 0 -> bit k
        }
        if (threadIdx.x == 0 && full < nb) {
            uint8_t v = 0;
            for (int64_t j = full * 8; j < N; ++j) { v |= (uint8_t)((w[j] & 1u) << (j & 7)); not_bits |= w[j] > 1; }
            s_bits[full] = v;
        }
    } else {
        for (int64_t i = threadIdx.x; i < nb; i += blockDim.x) {
            uint8_t v = 0;
            for (int k = 0; k < 8 && i * 8 + k < N; ++k) { v |= (uint8_t)((w[i * 8 + k] & 1u) << k); not_bits |= w[i * 8 + k] > 1; }
            s_bits[i] = v;
        }
    }
    if (__syncthreads_or(not_bits)) {     // (uniform) byte gathers, as k_eval_syndrome
        for (int64_t ci = threadIdx.x; ci < C; ci += blockDim.x) {
            uint8_t acc = 0;
            for (int32_t s = chk_ptr[ci]; s < chk_ptr[ci + 1]; ++s) acc ^= w[slot_var[s]];
            synd[b * C + chk_order[ci]] = acc;
        }
        return;
    }
    for (int64_t ci = threadIdx.x; ci < C; ci += blockDim.x) {
        uint32_t acc = 0;
        for (int32_t s = chk_ptr[ci]; s < chk_ptr[ci + 1]; ++s) {
            const int32_t v = slot_var[s];
            acc ^= (uint32_t)s_bits[v >> 3] >> (v & 7);
        }
        synd[b * C + chk_order[ci]] = (uint8_t)(acc & 1u);
    }
}

// MODE 0: bits are bytes of `word`; MODE 1: bit = (lappr < 0) on float; MODE 2: on double
template <int MODE>
__global__ void k_check_frames(const int32_t *__restrict__ chk_ptr, const int32_t *__restrict__ slot_var,
                               const int32_t *__restrict__ chk_order, int64_t N, int64_t C,
                               const void *__restrict__ data, const uint8_t *__restrict__ synd,
                               uint8_t *__restrict__ ok)
{
    const int64_t b = blockIdx.y;
    const int64_t ci = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0;
    if (ci < C) {
        uint8_t parity = synd[b * C + chk_order[ci]];
        for (int32_t s = chk_ptr[ci]; s < chk_ptr[ci + 1]; ++s) {
            const int64_t v = slot_var[s];
            if (MODE == 0) parity ^= static_cast<const uint8_t *>(data)[b * N + v];
            else if (MODE == 1) parity ^= (uint8_t)(static_cast<const float *>(data)[b * N + v] < 0.0f);
            else parity ^= (uint8_t)(static_cast<const double *>(data)[b * N + v] < 0.0);
        }
        bad = ((uint8_t)(parity ^ 1)) == 0;  // decoder.pyx:207-208, :249
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) ok[b] = 0;
}

template <typename T>
__global__ void k_count_errors(const T *__restrict__ lappr, const uint8_t *__restrict__ word,
                               int64_t frame_len, int64_t k, int32_t *__restrict__ errors)
{
    const int64_t b = blockIdx.y;
    int32_t cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < k;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t w = word[b * frame_len + i];
        cnt += (lappr[b * frame_len + i] >= (T)0) ? (int32_t)w : 1 - (int32_t)w;  // utils.pyx:35-38
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&errors[b], cnt);
}

}  // namespace qr

extern "C" {

int qr_eval_syndrome(const qr_graph *g, const uint8_t *d_word, uint8_t *d_synd, int64_t frames,
                     void *stream)
{
    if (!g || g->device < 0) return qr::fail(QR_ERR_INVALID, "graph has no device");
    if (frames < 0 || frames > 65535 * 1024) return qr::fail(QR_ERR_INVALID, "bad frame count");
    if (frames == 0) return QR_OK;
    if (!d_word || !d_synd) return qr::fail(QR_ERR_INVALID, "null array");
    qr::DeviceGuard guard(g->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // enough frames to fill the GPU with one CTA per frame, and the packed word fits shared memory: staged kernel
    const size_t bits_bytes = (size_t)((g->N + 7) / 8 + 8);
    if (frames >= 256 && bits_bytes <= 200 * 1024 && frames <= 0x7fffffff) {
        QR_CUDA_CHECK(cudaFuncSetAttribute(qr::k_eval_syndrome_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        qr::k_eval_syndrome_smem<<<(unsigned)frames, 512, bits_bytes, st>>>(g->d_chk_ptr, g->d_slot_var, g->d_chk_order,
                                                                            g->N, g->C, d_word, d_synd);
        QR_CUDA_CHECK(cudaGetLastError());
        return QR_OK;
    }
    for (int64_t b0 = 0; b0 < frames; b0 += 65535) {
        const int64_t nb = frames - b0 < 65535 ? frames - b0 : 65535;
        dim3 grid((unsigned)((g->C + 255) / 256), (unsigned)nb);
        qr::k_eval_syndrome<<<grid, 256, 0, st>>>(g->d_chk_ptr, g->d_slot_var, g->d_chk_order, g->N, g->C,
                                                  d_word + b0 * g->N, d_synd + b0 * g->C);
    }
    QR_CUDA_CHECK(cudaGetLastError());
    return QR_OK;
}

static int check_frames(const qr_graph *g, int mode, const void *data, size_t elt, const uint8_t *d_synd,
                        int64_t frames, uint8_t *d_ok, void *stream)
{
    if (!g || g->device < 0) return qr::fail(QR_ERR_INVALID, "graph has no device");
    if (frames < 0) return qr::fail(QR_ERR_INVALID, "bad frame count");
    if (frames == 0) return QR_OK;
    if (!data || !d_synd || !d_ok) return qr::fail(QR_ERR_INVALID, "null array");
    qr::DeviceGuard guard(g->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    QR_CUDA_CHECK(cudaMemsetAsync(d_ok, 1, (size_t)frames, st));
    for (int64_t b0 = 0; b0 < frames; b0 += 65535) {
        const int64_t nb = frames - b0 < 65535 ? frames - b0 : 65535;
        dim3 grid((unsigned)((g->C + 255) / 256), (unsigned)nb);
        const void *p = static_cast<const char *>(data) + (size_t)b0 * g->N * elt;
        if (mode == 0)
            qr::k_check_frames<0><<<grid, 256, 0, st>>>(g->d_chk_ptr, g->d_slot_var, g->d_chk_order, g->N,
                                                        g->C, p, d_synd + b0 * g->C, d_ok + b0);
        else if (mode == 1)
            qr::k_check_frames<1><<<grid, 256, 0, st>>>(g->d_chk_ptr, g->d_slot_var, g->d_chk_order, g->N,
                                                        g->C, p, d_synd + b0 * g->C, d_ok + b0);
        else
            qr::k_check_frames<2><<<grid, 256, 0, st>>>(g->d_chk_ptr, g->d_slot_var, g->d_chk_order, g->N,
                                                        g->C, p, d_synd + b0 * g->C, d_ok + b0);
    }
    QR_CUDA_CHECK(cudaGetLastError());
    return QR_OK;
}

int qr_check_word(const qr_graph *g, const uint8_t *d_word, const uint8_t *d_synd, int64_t frames,
                  uint8_t *d_ok, void *stream)
{
    return check_frames(g, 0, d_word, 1, d_synd, frames, d_ok, stream);
}

int qr_check_lappr(const qr_graph *g, const void *d_lappr, int dtype, const uint8_t *d_synd,
                   int64_t frames, uint8_t *d_ok, void *stream)
{
    if (dtype != QR_F32 && dtype != QR_F64) return qr::fail(QR_ERR_INVALID, "bad dtype");
    return check_frames(g, dtype == QR_F32 ? 1 : 2, d_lappr, dtype == QR_F32 ? 4 : 8, d_synd, frames, d_ok,
                        stream);
}

int qr_count_errors(const void *d_lappr, int dtype, const uint8_t *d_word, int64_t frames,
                    int64_t frame_len, int64_t k, int32_t *d_errors, void *stream)
{
    if (dtype != QR_F32 && dtype != QR_F64) return qr::fail(QR_ERR_INVALID, "bad dtype");
    if (frames < 0 || k < 0 || k > frame_len) return qr::fail(QR_ERR_INVALID, "bad sizes");
    if (frames == 0) return QR_OK;
    if (!d_lappr || !d_word || !d_errors) return qr::fail(QR_ERR_INVALID, "null array");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    QR_CUDA_CHECK(cudaMemsetAsync(d_errors, 0, (size_t)frames * sizeof(int32_t), st));
    if (k == 0) return QR_OK;
    const unsigned gx = (unsigned)((k + 1023) / 1024 < 64 ? (k + 1023) / 1024 : 64);
    for (int64_t b0 = 0; b0 < frames; b0 += 65535) {
        const int64_t nb = frames - b0 < 65535 ? frames - b0 : 65535;
        dim3 grid(gx, (unsigned)nb);
        if (dtype == QR_F32)
            qr::k_count_errors<float><<<grid, 256, 0, st>>>(static_cast<const float *>(d_lappr) + b0 * frame_len,
                                                            d_word + b0 * frame_len, frame_len, k, d_errors + b0);
        else
            qr::k_count_errors<double><<<grid, 256, 0, st>>>(static_cast<const double *>(d_lappr) + b0 * frame_len,
                                                             d_word + b0 * frame_len, frame_len, k, d_errors + b0);
    }
    QR_CUDA_CHECK(cudaGetLastError());
    return QR_OK;
}

}  // extern "C"
