// membench -- what the B200 memory system gives for the access patterns of the decoder kernels
// (development / evidence tool: builder-measured L2 and HBM bandwidth).
//
//   stream   : every thread streams 16-byte loads over a buffer of the given size (L2-resident when it fits)
//   gather   : 8 threads share one ROW-byte row (ROW = 128: one fp32 lane tile row), rows picked by a hash
//              over a working set of the given size, U independent LDG.128 in flight per thread --
//              the fused decoder's gather pattern (ld.global.cg = L2 only, no L1 allocation)
//   copy     : read + write streaming (the MEASURED_PEAKS.json hbm_gbs pattern)
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/membench tools/membench.cu
// usage: membench            (prints one line per pattern and working-set size)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint4 ldcg(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int U>
__global__ void __launch_bounds__(256) k_stream(const uint4 *buf, size_t n_vec, int trips, uint32_t *sink)
{
    // every thread makes `trips` trips of U independent 16-byte loads, walking the buffer with the grid's stride and
    // wrapping around: the whole grid re-reads the working set over and over (L2-resident when it fits)
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) % n_vec;
    for (int r = 0; r < trips; ++r) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            v[u] = ldcg(buf + i);
            i += stride;
            if (i >= n_vec) i -= n_vec * (i / n_vec);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

// rows of ROWB bytes; TPR = ROWB / 16 threads per row; each thread does `iters` rounds of U gathered rows
template <int U>
__global__ void __launch_bounds__(256) k_gather(const uint4 *buf, uint32_t n_rows, int tpr, int iters, uint32_t seed,
                                                uint32_t *sink)
{
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t grp = gt / tpr, tx = gt % tpr;
    uint32_t acc = 0, h = hash32(grp * 2654435761u + seed);
    for (int it = 0; it < iters; ++it) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            h = hash32(h + 0x9e3779b9u);
            const uint32_t row = (uint32_t)(((uint64_t)h * n_rows) >> 32);
            v[u] = ldcg(buf + (size_t)row * tpr + tx);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void __launch_bounds__(256) k_copy(const uint4 *src, uint4 *dst, size_t n_vec)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) dst[i] = src[i];
}

template <typename F>
static float time_ms(F launch, int reps = 5)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("# %s, %d SMs, L2 %d MB\n", prop.name, sms, prop.l2CacheSize >> 20);
    const size_t cap = (size_t)4 << 30;
    uint4 *buf, *dst;
    uint32_t *sink;
    CK(cudaMalloc(&buf, cap)); CK(cudaMalloc(&dst, cap)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 1, cap)); CK(cudaMemset(dst, 0, cap));
    const int grid = sms * 8;

    // copy (read + write), the MEASURED_PEAKS pattern
    {
        const size_t n_vec = cap / 16;
        float ms = time_ms([&] { k_copy<<<grid, 256>>>(buf, dst, n_vec); });
        printf("copy      4096 MB  read+write               %8.1f GB/s\n", 2.0 * cap / ms / 1e6);
    }
    // streaming reads, working sets from L2-resident to HBM
    for (size_t mb : {16, 32, 64, 96, 256, 2048}) {
        const size_t bytes = mb << 20, n_vec = bytes / 16;
        const int trips = 256;                                    // 8 x 256 loads of 16 B per thread: 9.9 GB per launch
        float ms = time_ms([&] { k_stream<8><<<grid, 256>>>(buf, n_vec, trips, sink); });
        printf("stream  %6zu MB  LDG.128.cg x8 in flight    %8.1f GB/s\n", mb, (double)grid * 256 * trips * 8 * 16 / ms / 1e6);
    }
    // gathers of 128-byte rows (8 threads per row), 2 CTAs of 256 threads per SM like k_fused
    for (int tpr : {8, 4, 32}) {
        for (size_t mb : {32, 64, 96, 512, 4096}) {
            const size_t bytes = mb << 20;
            const uint32_t n_rows = (uint32_t)(bytes / (16 * tpr));
            const int g2 = sms * 2, iters = 3072;
            const double moved = (double)g2 * 256 * iters * 16;
            float m8 = time_ms([&] { k_gather<8><<<g2, 256>>>(buf, n_rows, tpr, iters / 8, 1, sink); });
            float m16 = time_ms([&] { k_gather<16><<<g2, 256>>>(buf, n_rows, tpr, iters / 16, 2, sink); });
            float m24 = time_ms([&] { k_gather<24><<<g2, 256>>>(buf, n_rows, tpr, iters / 24, 3, sink); });
            const int g4 = sms * 4;
            float m16b = time_ms([&] { k_gather<16><<<g4, 256>>>(buf, n_rows, tpr, iters / 16, 4, sink); });
            printf("gather  %6zu MB  rows of %4d B: 16 warps/SM U=8 %7.1f  U=16 %7.1f  U=24 %7.1f GB/s; 32 warps/SM U=16 %7.1f GB/s\n",
                   mb, 16 * tpr, moved / m8 / 1e6, (double)g2 * 256 * (iters / 16) * 16 * 16 / m16 / 1e6,
                   (double)g2 * 256 * (iters / 24) * 24 * 16 / m24 / 1e6, (double)g4 * 256 * (iters / 16) * 16 * 16 / m16b / 1e6);
        }
    }
    return 0;
}
