// Opaque handle layouts shared by the translation units of libqamrecon.
#pragma once

#include <cuda_runtime.h>

#include "qr_common.h"
#include "qr_mapper_core.cuh"

struct qr_mapper;
struct qr_decoder;

namespace qr {
struct LaneState;

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); cudaSetDevice(dev); }
    ~DeviceGuard() { cudaSetDevice(prev); }
};
}  // namespace qr

namespace qr {
MapperView mapper_view(const qr_mapper *m);
}  // namespace qr

struct qr_decoder {
    const qr_graph *g = nullptr;
    int precision = QR_F32;
    int schedule = QR_SCHED_AUTO;   // fused flooding iteration where the graph allows it, else the two-phase kernel
    int32_t lanes = 0;
    int device = 0;
    int regular_degree = 0;  // > 0: check-regular graph with a specialised kernel
    int vec = 0;             // 0: default lanes per thread; 1 or 2: narrower (fp32 only, experimental)
    void *c2v = nullptr, *post = nullptr, *llr = nullptr;
    void *c2v2 = nullptr;    // second message buffer of QR_SCHED_FUSED (allocated on first use)
    int fused_tile = 32;     // lanes per L2 tile of the fused schedule (32, 64 or 128)
    int fused_store_post = 1;   // store posteriors of lanes that may finish (cheap shipping), 0: always rebuild them
    int fused_rpc = 4;       // checks per thread and claim of a fused sweep
    int fused_pp_items = 0;  // items per tile post-processing (0: from the code length)
    int32_t *fused_ctl = nullptr;    // per-tile counters of the tile pipeline
    void *fused_rlist = nullptr;     // per-tile refill lists
    void *fused_ppq = nullptr;       // ready queue of post-processing items
    void *fused_nbrl = nullptr;      // lean neighbour records of the float mode (built for fused_nbrl_tl lanes per tile)
    int fused_nbrl_tl = 0;
    int fused_park = 1;              // rounds a finished lane waits for its sector group before it is refilled alone (QAMRECON_FUSED_PARK)
    int fused_lean = 1;              // 0: the generic item in float mode too (QAMRECON_FUSED_LEAN)
    int fused_hints = 1;     // L2 policy of the fused schedule: 0 none, 1 stores evict-first, 2 + loads evict-last
    uint8_t *synd = nullptr;
    qr::LaneState *st = nullptr;          // [2][lanes]
    int32_t *unsat = nullptr;             // [2][lanes]
    int32_t *ctrl = nullptr;              // [CTRL_WORDS]
    unsigned long long *stats = nullptr;  // [2]
    int32_t *work = nullptr;              // [2][kMaxLaneTiles]
    int32_t *refill_list = nullptr;       // [2][lanes]
    int32_t *h_ctrl = nullptr;            // pinned mirrors
    unsigned long long *h_stats = nullptr;
    cudaStream_t last_stream = nullptr;
    int coop_grid = 0;
    int sm_count = 0;
    // scratch of qr_reconcile_host (grow-only)
    void *pipe_buf = nullptr;
    size_t pipe_cap = 0;
    void *dev_buf = nullptr;   // scratch of qr_reconcile_device (grow-only)
    size_t dev_cap = 0;
    bool pipe_streams_ready = false;
    cudaStream_t s_in = nullptr, s_out = nullptr;   // copy streams of the host pipeline
    cudaEvent_t ev_in[2] = {}, ev_compute[2] = {}, ev_out[2] = {}, ev_start = nullptr;
};

struct qr_mapper {
    int device = 0;
    int order = 0, bps = 0;
    double noise_var = 0, sigma = 0, s2 = 0;
    double *d_tables = nullptr;  // one allocation holding every table below
    uint8_t *d_sign = nullptr;
    int32_t *d_index_errors = nullptr;   // see MapperView::index_errors
    uint8_t *d_sign_g = nullptr;   // sign rule of g / g_inv / map_noise (differs from d_sign only for the FlipSign subclasses)
    double *grid_y = nullptr, *grid_F = nullptr;   // dense F_Y grid of noisemapper.pyx:135-144 (one allocation), built on demand
    int32_t grid_n = 0;
    double *constellation = nullptr, *thresholds = nullptr, *probabilities = nullptr;
    double *FY_thr = nullptr, *delta = nullptr, *fwrd = nullptr, *back = nullptr, *bare = nullptr,
           *inf_erf = nullptr;
    size_t n_table_doubles = 0;
    double *inv_tab = nullptr;   // F_Y on a uniform grid, see MapperView
    double *inv_pdf = nullptr;   // its density on the same grid (second half of the same allocation)
    int32_t inv_n = 0;
    int32_t uniform = 1;           // equally spaced constellation
    int32_t *inv_jump = nullptr;   // InvTable::jump
    int32_t inv_jn = 0;
    double inv_y0 = 0, inv_h = 0;
    double *inv32_F = nullptr;     // coarse copy for the fp32-grade demapper (InvTable32)
    float *inv32_f = nullptr;
    uint16_t *inv32_jump = nullptr;
    double inv32_h = 0;
};
