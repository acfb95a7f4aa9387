// Device-only helpers of the mapper kernels: the alphabet tables staged in shared memory.
#pragma once

#include "qr_mapper_core.cuh"

namespace qr {

struct SharedTables {
    double a[kMaxOrder], p[kMaxOrder], thr[kMaxOrder + 1], FYt[kMaxOrder + 1], delta[kMaxOrder];
    uint8_t sign[kMaxOrder];
    // fast demapper: exp(-(m step)^2 / 2 sigma^2) and exp(-(m step)^2) (the reference's undivided k < j exponent)
    double ghi[kMaxOrder], glo[kMaxOrder];
    double pz[3 * kMaxOrder];   // probabilities zero-padded on both sides: pz[order - 1 + k] = p[k], 0 elsewhere
    float g2hi[kMaxOrder], g2lo[kMaxOrder], pzf[3 * kMaxOrder];   // fp32-grade demapper, see TablesRef
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
    if (m.order > 1) {
        const double step = m.constellation[1] - m.constellation[0];
        for (int i = threadIdx.x; i < m.order; i += blockDim.x) {
            const double t2 = (i * step) * (i * step);
            s.ghi[i] = exp(-t2 / (2 * m.noise_var));
            s.glo[i] = exp(-t2);
            s.g2hi[i] = (float)(t2 / (2 * m.noise_var) * 1.4426950408889634);
            s.g2lo[i] = (float)(t2 * 1.4426950408889634);
        }
    }
    for (int i = threadIdx.x; i < 3 * m.order; i += blockDim.x) {
        const int k = i - (m.order - 1);
        s.pz[i] = (k >= 0 && k < m.order) ? m.probabilities[k] : 0.0;
        s.pzf[i] = (float)s.pz[i];
    }
    __syncthreads();
}

__device__ __forceinline__ TablesRef tables_ref(const SharedTables &s)
{
    return TablesRef{s.a, s.p, s.thr, s.FYt, s.delta, s.sign, s.ghi, s.glo, s.pz, s.g2hi, s.g2lo, s.pzf};   // (callers drop ghi/glo when !uniform)
}

}  // namespace qr
