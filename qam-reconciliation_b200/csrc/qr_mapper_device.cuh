// Device-only helpers of the mapper kernels: the alphabet tables staged in shared memory.
#pragma once

#include "qr_mapper_core.cuh"

namespace qr {

struct SharedTables {
    double a[kMaxOrder], p[kMaxOrder], thr[kMaxOrder + 1], FYt[kMaxOrder + 1], delta[kMaxOrder];
    uint8_t sign[kMaxOrder];
};

__device__ __forceinline__ void stage_tables(const MapperView &m, SharedTables &s)
{
    for (int i = threadIdx.x; i < m.order; i += blockDim.x) {
        s.a[i] = m.constellation[i];
        s.p[i] = m.probabilities[i];
        s.delta[i] = m.delta[i];
        s.sign[i] = m.sign_config[i];
    }
    for (int i = threadIdx.x; i <= m.order; i += blockDim.x) {
        s.thr[i] = m.thresholds[i];
        s.FYt[i] = m.FY_thr[i];
    }
    __syncthreads();
}

__device__ __forceinline__ TablesRef tables_ref(const SharedTables &s)
{
    return TablesRef{s.a, s.p, s.thr, s.FYt, s.delta, s.sign};
}

}  // namespace qr
