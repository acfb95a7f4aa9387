// Whole reverse-reconciliation pass over host buffers: the chain of the reference's Monte-Carlo
// loop body (sims/reconciliation.pyx:129-153 soft reverse, :300-308 hard reverse, :214-227 direct)
// for a batch of frames, host -> device -> host inside one call.
//
// The batch is cut into chunks that flow through a three-stage pipeline on three streams --
// host-to-device copy of chunk c+1, kernels of chunk c, device-to-host copy of chunk c-1 -- with
// two sets of device buffers, so with pinned host memory the PCIe transfers hide behind the
// decoder.  The decoder workspace is shared: kernels of consecutive chunks serialise on the
// caller's stream.
#include <cuda_runtime.h>

#include <vector>

#include <algorithm>
#include <cstring>

#include "qr_handles.h"

namespace {

// compact wire format: float samples / byte symbols widened on the device, hard decisions packed 8 per byte
__global__ void k_widen(const float *__restrict__ y32, const uint8_t *__restrict__ tx8, int64_t n, double *__restrict__ y,
                        long long *__restrict__ tx)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        y[i] = (double)y32[i];
        tx[i] = (long long)tx8[i];
    }
}

// bit (i & 7) of byte i >> 3 of a frame's row = (posterior[i] < 0), the hard decision of decoder.pyx:244
template <typename T>
__global__ void k_pack_decisions(const T *__restrict__ post, int64_t frames, int64_t N, int64_t row_bytes,
                                 uint8_t *__restrict__ packed)
{
    const int64_t total = frames * row_bytes;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = j / row_bytes, b = j - f * row_bytes;
        const T *p = post + f * N + b * 8;
        uint32_t v = 0;
        for (int k = 0; k < 8; ++k)
            if (b * 8 + k < N && p[k] < (T)0) v |= 1u << k;
        packed[j] = (uint8_t)v;
    }
}

struct Carver {
    char *base;
    size_t off = 0;
    explicit Carver(void *b) : base(static_cast<char *>(b)) {}
    template <typename T>
    T *take(size_t count)
    {
        off = (off + 255) / 256 * 256;
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

struct Buffers {
    double *y, *n_hat;
    int64_t *tx, *idx;
    uint8_t *word, *synd, *success;
    void *llr, *post;
    int32_t *iters, *errors;
    float *y32;          // compact wire format only
    uint8_t *tx8, *packed;
};

Buffers carve(Carver &c, int64_t frames, int64_t S, int64_t N, int64_t C, size_t wl, size_t wp, bool compact = false)
{
    Buffers b;
    b.y32 = compact ? c.take<float>(frames * S) : nullptr;
    b.tx8 = compact ? c.take<uint8_t>(frames * S) : nullptr;
    b.packed = compact ? c.take<uint8_t>(frames * ((N + 7) / 8)) : nullptr;
    b.y = c.take<double>(frames * S);
    b.n_hat = c.take<double>(frames * S);
    b.tx = c.take<int64_t>(frames * S);
    b.idx = c.take<int64_t>(frames * S);
    b.word = c.take<uint8_t>(frames * N);
    b.synd = c.take<uint8_t>(frames * C);
    b.success = c.take<uint8_t>(frames);
    b.llr = c.take<char>(frames * N * wl);
    b.post = c.take<char>(frames * N * wp);
    b.iters = c.take<int32_t>(frames);
    b.errors = c.take<int32_t>(frames);
    return b;
}

// kernels of one chunk: y/tx (device) -> word, synd, llr -> decode -> (errors)
int run_chain(qr_decoder *d, const qr_mapper *m, int mode, int demap_mode, double alpha, const double *y,
              const int64_t *tx, int64_t nf, int64_t S, int32_t max_iterations, int64_t k_info, double *n_hat,
              uint8_t *word, uint8_t *synd, void *llr, int llr_dtype, uint8_t *success, int32_t *iters, void *post,
              int post_dtype, int32_t *errors, cudaStream_t st)
{
    const qr_graph *g = d->g;
    int rc;
    if (mode == 0) {
        if ((rc = qr_front_end(m, y, nf * S, nullptr, n_hat, word, st))) return rc;
        if ((rc = qr_eval_syndrome(g, word, synd, nf, st))) return rc;
        if ((rc = qr_demap_lappr(m, n_hat, tx, nf * S, demap_mode, alpha, llr, llr_dtype, st))) return rc;
    } else if (mode == 1) {
        if ((rc = qr_front_end(m, y, nf * S, nullptr, nullptr, word, st))) return rc;
        if ((rc = qr_eval_syndrome(g, word, synd, nf, st))) return rc;
        if ((rc = qr_bare_llr(m, tx, nf * S, llr, llr_dtype, st))) return rc;
    } else {
        if ((rc = qr_symbols_to_bits(m, tx, nf * S, word, st))) return rc;
        if ((rc = qr_eval_syndrome(g, word, synd, nf, st))) return rc;
        if ((rc = qr_direct_llr(m, y, nf * S, 2 * m->noise_var, llr, llr_dtype, st))) return rc;
    }
    if ((rc = qr_decode_batch(d, llr, llr_dtype, synd, nf, max_iterations, success, iters, post, post_dtype, st)))
        return rc;
    if (errors && (rc = qr_count_errors(post, post_dtype, word, nf, g->N, k_info, errors, st))) return rc;
    return QR_OK;
}

}  // namespace

extern "C" int qr_reconcile_device(qr_decoder *d, const qr_mapper *m, int mode, int demap_mode, double alpha,
                                   const double *d_y, const int64_t *d_tx_index, int64_t frames,
                                   int32_t max_iterations, int64_t k_info, uint8_t *d_success, int32_t *d_iters,
                                   void *d_post, int post_dtype, uint8_t *d_word, uint8_t *d_synd,
                                   int32_t *d_bit_errors, void *stream_)
{
    if (!d || !m) return qr::fail(QR_ERR_INVALID, "null handle");
    if (mode < 0 || mode > 2) return qr::fail(QR_ERR_INVALID, "reconciliation mode must be 0, 1 or 2");
    if (frames < 0) return qr::fail(QR_ERR_INVALID, "bad frame count");
    if (d->device != m->device) return qr::fail(QR_ERR_INVALID, "decoder and mapper live on different devices");
    const qr_graph *g = d->g;
    const int64_t N = g->N, C = g->C;
    if (N % m->bps) return qr::fail(QR_ERR_INVALID, "codeword length is not a multiple of bits per symbol");
    if (k_info < 0 || k_info > N) return qr::fail(QR_ERR_INVALID, "bad information length");
    if (d_post && post_dtype != QR_F32 && post_dtype != QR_F64) return qr::fail(QR_ERR_INVALID, "bad dtype");
    if (frames == 0) return QR_OK;
    if (!d_y || !d_tx_index || !d_success || !d_iters) return qr::fail(QR_ERR_INVALID, "null array");
    const int64_t S = N / m->bps;
    const int llr_dtype = d->precision == QR_F64 ? QR_F64 : QR_F32;
    const size_t wl = llr_dtype == QR_F64 ? 8 : 4;
    if (!d_post) post_dtype = llr_dtype;
    const size_t wp = post_dtype == QR_F64 ? 8 : 4;
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    qr::DeviceGuard guard(d->device);
    // scratch: n_hat, llr, and whatever output the caller did not ask for
    Carver sizing(nullptr);
    auto layout = [&](Carver &c, double *&n_hat, void *&llr, uint8_t *&word, uint8_t *&synd, void *&post) {
        n_hat = c.take<double>(frames * S);
        llr = c.take<char>(frames * N * wl);
        word = d_word ? d_word : c.take<uint8_t>(frames * N);
        synd = d_synd ? d_synd : c.take<uint8_t>(frames * C);
        post = d_post ? d_post : (d_bit_errors ? (void *)c.take<char>(frames * N * wp) : nullptr);
    };
    double *n_hat; void *llr, *post; uint8_t *word, *synd;
    layout(sizing, n_hat, llr, word, synd, post);
    const size_t need = sizing.off + 256;
    if (need > d->dev_cap) {
        QR_CUDA_CHECK(cudaStreamSynchronize(st));
        cudaFree(d->dev_buf);
        d->dev_buf = nullptr; d->dev_cap = 0;
        QR_CUDA_CHECK(cudaMalloc(&d->dev_buf, need));
        d->dev_cap = need;
    }
    Carver carver(d->dev_buf);
    layout(carver, n_hat, llr, word, synd, post);
    return run_chain(d, m, mode, demap_mode, alpha, d_y, d_tx_index, frames, S, max_iterations, k_info, n_hat, word,
                     synd, llr, llr_dtype, d_success, d_iters, post, post_dtype, d_bit_errors, st);
}

// How qr_reconcile_host cuts a batch into chunks that flow through upload / kernels / download.
//  * Size: about eight chunks per batch and two decoder fills per chunk (with fewer lanes than frames, continuous
//    batching needs more frames than lanes to keep the lanes busy), never fewer than four chunks when the batch
//    allows it; and when a quarter of the batch fits the resident lanes, exactly that: four chunks hide the copies
//    and a chunk that fits its lanes never refills one (refill generations cost ~8 % at 3 dB).
//  * Ends: the first chunk's upload and the last chunk's download cannot hide behind kernels, so the batch ramps
//    up and down.  On B200 an upload in the reference's types runs ~3x as fast as the kernels consume frames and a
//    download ~6x: each head piece is 3x the one before it (its upload hides behind the kernels of the previous
//    piece), the tail pieces shrink 5x (64, 192, 576, full ..., 320, 64 for quarters of 1024 frames).  Batches
//    too small for the ramp, and the compact wire format (a third of the upload, next to no download: the ramp's
//    extra launches cost more than they hide), get one small piece at either end (a quarter of a chunk, 32 to 256
//    frames).
static std::vector<int64_t> host_chunk_cuts(int64_t frames, int64_t lanes, bool compact)
{
    std::vector<int64_t> cuts;
    cuts.push_back(0);
    if (frames <= 0) return cuts;
    int64_t chunk = std::max<int64_t>(2 * lanes, (frames + 7) / 8);
    chunk = std::min<int64_t>(chunk, std::max<int64_t>(lanes, (frames + 3) / 4));
    const int64_t quarter = ((frames + 3) / 4 + 31) / 32 * 32;
    const int64_t piece = std::max<int64_t>(quarter, 256);     // (below 8 lane tiles a launch is latency-bound: do not cut finer)
    if (lanes >= piece) chunk = piece;
    chunk = std::min(chunk, frames);
    if (chunk < frames) {
        std::vector<int64_t> up, down;
        for (int64_t x = 64; x < chunk; x *= 3) up.push_back(x);
        for (int64_t x = 64; x < chunk && down.size() < 2; x *= 5) down.push_back(x);
        int64_t ramp = 0;
        for (int64_t x : up) ramp += x;
        for (int64_t x : down) ramp += x;
        if (!compact && !up.empty() && frames >= ramp + chunk) {
            int64_t tail = 0;
            for (int64_t x : down) tail += x;
            for (int64_t x : up) cuts.push_back(cuts.back() + x);
            while (frames - tail - cuts.back() > chunk) cuts.push_back(cuts.back() + chunk);
            if (frames - tail > cuts.back()) cuts.push_back(frames - tail);
            for (size_t i = down.size(); i-- > 0;) cuts.push_back(cuts.back() + down[i]);
        } else {
            const int64_t small = std::max<int64_t>(32, std::min<int64_t>(256, chunk / 4 / 32 * 32));
            cuts.push_back(std::min(small, frames));
            while (frames - cuts.back() > chunk + small) cuts.push_back(cuts.back() + chunk);
            if (frames - cuts.back() > small) cuts.push_back(frames - small);
        }
    }
    if (cuts.back() < frames) cuts.push_back(frames);
    return cuts;
}

extern "C" int qr_host_chunk_cuts(int64_t frames, int64_t lanes, int compact, int64_t *cuts, int32_t max_cuts,
                                  int32_t *n_cuts)
{
    if (frames < 0 || lanes <= 0 || !cuts || !n_cuts || max_cuts < 2) return qr::fail(QR_ERR_INVALID, "bad arguments");
    const std::vector<int64_t> c = host_chunk_cuts(frames, lanes, compact != 0);
    if ((int64_t)c.size() > max_cuts) return qr::fail(QR_ERR_INVALID, "cut array too small");
    for (size_t i = 0; i < c.size(); ++i) cuts[i] = c[i];
    *n_cuts = (int32_t)c.size();
    return QR_OK;
}

static int reconcile_host_impl(qr_decoder *d, const qr_mapper *m, int mode, int demap_mode, double alpha,
                               const double *h_y, const int64_t *h_tx_index, const float *h_y32, const uint8_t *h_tx8,
                               int64_t frames, int32_t max_iterations, int64_t k_info, uint8_t *h_success,
                               int32_t *h_iters, void *h_post, int post_dtype, uint8_t *h_word, uint8_t *h_packed,
                               int32_t *h_bit_errors, void *stream_);

extern "C" int qr_reconcile_host(qr_decoder *d, const qr_mapper *m, int mode, int demap_mode, double alpha,
                                 const double *h_y, const int64_t *h_tx_index, int64_t frames,
                                 int32_t max_iterations, int64_t k_info, uint8_t *h_success,
                                 int32_t *h_iters, void *h_post, int post_dtype, uint8_t *h_word,
                                 int32_t *h_bit_errors, void *stream_)
{
    if (frames > 0 && (!h_y || !h_tx_index)) return qr::fail(QR_ERR_INVALID, "null input array");
    return reconcile_host_impl(d, m, mode, demap_mode, alpha, h_y, h_tx_index, nullptr, nullptr, frames, max_iterations,
                               k_info, h_success, h_iters, h_post, post_dtype, h_word, nullptr, h_bit_errors, stream_);
}

extern "C" int qr_reconcile_host_compact(qr_decoder *d, const qr_mapper *m, int mode, int demap_mode, double alpha,
                                         const float *h_y32, const uint8_t *h_tx8, int64_t frames,
                                         int32_t max_iterations, int64_t k_info, uint8_t *h_success, int32_t *h_iters,
                                         uint8_t *h_decisions_packed, int32_t *h_bit_errors, void *stream_)
{
    if (frames > 0 && (!h_y32 || !h_tx8)) return qr::fail(QR_ERR_INVALID, "null input array");
    if (m && m->order > 256) return qr::fail(QR_ERR_INVALID, "byte symbols need an alphabet of at most 256 points");
    return reconcile_host_impl(d, m, mode, demap_mode, alpha, nullptr, nullptr, h_y32, h_tx8, frames, max_iterations,
                               k_info, h_success, h_iters, nullptr, QR_F32, nullptr, h_decisions_packed, h_bit_errors,
                               stream_);
}

static int reconcile_host_impl(qr_decoder *d, const qr_mapper *m, int mode, int demap_mode, double alpha,
                               const double *h_y, const int64_t *h_tx_index, const float *h_y32, const uint8_t *h_tx8,
                               int64_t frames, int32_t max_iterations, int64_t k_info, uint8_t *h_success,
                               int32_t *h_iters, void *h_post, int post_dtype, uint8_t *h_word, uint8_t *h_packed,
                               int32_t *h_bit_errors, void *stream_)
{
    const bool compact = h_y32 != nullptr;
    if (!d || !m) return qr::fail(QR_ERR_INVALID, "null handle");
    if (mode < 0 || mode > 2) return qr::fail(QR_ERR_INVALID, "reconciliation mode must be 0, 1 or 2");
    if (frames < 0) return qr::fail(QR_ERR_INVALID, "bad frame count");
    if (d->device != m->device) return qr::fail(QR_ERR_INVALID, "decoder and mapper live on different devices");
    const qr_graph *g = d->g;
    const int64_t N = g->N, C = g->C;
    if (N % m->bps) return qr::fail(QR_ERR_INVALID, "codeword length is not a multiple of bits per symbol");
    if (k_info < 0 || k_info > N) return qr::fail(QR_ERR_INVALID, "bad information length");
    if (h_post && post_dtype != QR_F32 && post_dtype != QR_F64) return qr::fail(QR_ERR_INVALID, "bad dtype");
    if (frames == 0) return QR_OK;
    const int64_t S = N / m->bps;
    const int llr_dtype = d->precision == QR_F64 ? QR_F64 : QR_F32;
    const size_t wl = llr_dtype == QR_F64 ? 8 : 4;
    if (!h_post) post_dtype = llr_dtype;
    const size_t wp = post_dtype == QR_F64 ? 8 : 4;
    const int64_t row_bytes = (N + 7) / 8;
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    qr::DeviceGuard guard(d->device);

    const std::vector<int64_t> cuts = host_chunk_cuts(frames, d->lanes, compact);    // chunk boundaries: cuts[c] .. cuts[c + 1]
    int64_t chunk = 0;                                                        // largest piece: what the buffers hold
    for (size_t c = 0; c + 1 < cuts.size(); ++c) chunk = std::max(chunk, cuts[c + 1] - cuts[c]);
    const int64_t n_chunks = (int64_t)cuts.size() - 1;
    const int n_sets = n_chunks > 1 ? 2 : 1;

    if (!d->pipe_streams_ready) {
        QR_CUDA_CHECK(cudaStreamCreateWithFlags(&d->s_in, cudaStreamNonBlocking));
        QR_CUDA_CHECK(cudaStreamCreateWithFlags(&d->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            QR_CUDA_CHECK(cudaEventCreateWithFlags(&d->ev_in[i], cudaEventDisableTiming));
            QR_CUDA_CHECK(cudaEventCreateWithFlags(&d->ev_compute[i], cudaEventDisableTiming));
            QR_CUDA_CHECK(cudaEventCreateWithFlags(&d->ev_out[i], cudaEventDisableTiming));
        }
        QR_CUDA_CHECK(cudaEventCreateWithFlags(&d->ev_start, cudaEventDisableTiming));
        d->pipe_streams_ready = true;
    }

    Carver sizing(nullptr);
    for (int s = 0; s < n_sets; ++s) carve(sizing, chunk, S, N, C, wl, wp, compact);
    const size_t need = sizing.off + 256;
    if (need > d->pipe_cap) {
        QR_CUDA_CHECK(cudaDeviceSynchronize());
        cudaFree(d->pipe_buf);
        d->pipe_buf = nullptr;
        d->pipe_cap = 0;
        QR_CUDA_CHECK(cudaMalloc(&d->pipe_buf, need));
        d->pipe_cap = need;
    }
    Carver carver(d->pipe_buf);
    Buffers sets[2];
    for (int s = 0; s < n_sets; ++s) sets[s] = carve(carver, chunk, S, N, C, wl, wp, compact);

    // work queued on the caller's stream before this call must precede our copies
    QR_CUDA_CHECK(cudaEventRecord(d->ev_start, st));
    QR_CUDA_CHECK(cudaStreamWaitEvent(d->s_in, d->ev_start, 0));
    QR_CUDA_CHECK(cudaStreamWaitEvent(d->s_out, d->ev_start, 0));

    for (int64_t c = 0; c < n_chunks; ++c) {
        const int s = (int)(c % n_sets);
        const Buffers &b = sets[s];
        const int64_t f0 = cuts[c], nf = cuts[c + 1] - cuts[c];
        // stage 1: inputs of chunk c (the set is free once the kernels of chunk c-2 are done)
        if (c >= n_sets) QR_CUDA_CHECK(cudaStreamWaitEvent(d->s_in, d->ev_compute[s], 0));
        if (compact) {
            QR_CUDA_CHECK(cudaMemcpyAsync(b.y32, h_y32 + f0 * S, nf * S * sizeof(float), cudaMemcpyHostToDevice, d->s_in));
            QR_CUDA_CHECK(cudaMemcpyAsync(b.tx8, h_tx8 + f0 * S, nf * S, cudaMemcpyHostToDevice, d->s_in));
        } else {
            QR_CUDA_CHECK(cudaMemcpyAsync(b.y, h_y + f0 * S, nf * S * sizeof(double), cudaMemcpyHostToDevice, d->s_in));
            QR_CUDA_CHECK(cudaMemcpyAsync(b.tx, h_tx_index + f0 * S, nf * S * sizeof(int64_t), cudaMemcpyHostToDevice, d->s_in));
        }
        QR_CUDA_CHECK(cudaEventRecord(d->ev_in[s], d->s_in));
        // stage 2: kernels (outputs of the set are free once chunk c-2 has been copied out)
        QR_CUDA_CHECK(cudaStreamWaitEvent(st, d->ev_in[s], 0));
        if (c >= n_sets) QR_CUDA_CHECK(cudaStreamWaitEvent(st, d->ev_out[s], 0));
        if (compact) {
            k_widen<<<(unsigned)std::min<int64_t>((nf * S + 255) / 256, 148 * 16), 256, 0, st>>>(
                b.y32, b.tx8, nf * S, b.y, reinterpret_cast<long long *>(b.tx));
            QR_CUDA_CHECK(cudaGetLastError());
        }
        int rc = run_chain(d, m, mode, demap_mode, alpha, b.y, b.tx, nf, S, max_iterations, k_info, b.n_hat, b.word,
                           b.synd, b.llr, llr_dtype, b.success, b.iters, b.post, post_dtype,
                           h_bit_errors ? b.errors : nullptr, st);   // (b.post: also what the packed decisions are taken from)
        if (rc) {
            // copies of this and earlier chunks may still be in flight into / out of caller memory
            cudaStreamSynchronize(d->s_in); cudaStreamSynchronize(st); cudaStreamSynchronize(d->s_out);
            return rc;
        }
        if (h_packed) {
            const unsigned pg = (unsigned)std::min<int64_t>((nf * row_bytes + 255) / 256, 148 * 16);
            if (post_dtype == QR_F64) k_pack_decisions<double><<<pg, 256, 0, st>>>(static_cast<const double *>(b.post), nf, N, row_bytes, b.packed);
            else k_pack_decisions<float><<<pg, 256, 0, st>>>(static_cast<const float *>(b.post), nf, N, row_bytes, b.packed);
            QR_CUDA_CHECK(cudaGetLastError());
        }
        QR_CUDA_CHECK(cudaEventRecord(d->ev_compute[s], st));
        // stage 3: results of chunk c
        QR_CUDA_CHECK(cudaStreamWaitEvent(d->s_out, d->ev_compute[s], 0));
        if (h_bit_errors)
            QR_CUDA_CHECK(cudaMemcpyAsync(h_bit_errors + f0, b.errors, nf * sizeof(int32_t), cudaMemcpyDeviceToHost, d->s_out));
        if (h_success) QR_CUDA_CHECK(cudaMemcpyAsync(h_success + f0, b.success, nf, cudaMemcpyDeviceToHost, d->s_out));
        if (h_iters) QR_CUDA_CHECK(cudaMemcpyAsync(h_iters + f0, b.iters, nf * sizeof(int32_t), cudaMemcpyDeviceToHost, d->s_out));
        if (h_post)
            QR_CUDA_CHECK(cudaMemcpyAsync(static_cast<char *>(h_post) + (size_t)f0 * N * wp, b.post, (size_t)nf * N * wp,
                                          cudaMemcpyDeviceToHost, d->s_out));
        if (h_word) QR_CUDA_CHECK(cudaMemcpyAsync(h_word + f0 * N, b.word, nf * N, cudaMemcpyDeviceToHost, d->s_out));
        if (h_packed)
            QR_CUDA_CHECK(cudaMemcpyAsync(h_packed + f0 * row_bytes, b.packed, nf * row_bytes, cudaMemcpyDeviceToHost, d->s_out));
        QR_CUDA_CHECK(cudaEventRecord(d->ev_out[s], d->s_out));
    }
    QR_CUDA_CHECK(cudaStreamSynchronize(d->s_out));
    QR_CUDA_CHECK(cudaStreamSynchronize(st));
    return QR_OK;
}
