// Shared declarations of libqamrecon (host side): status codes, error reporting, handles.
#pragma once

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/qamrecon.h"

#if defined(__CUDACC__)
#define QR_HD __host__ __device__ __forceinline__
#define QR_HD_NOINLINE __host__ __device__ __noinline__
#else
#define QR_HD inline
#define QR_HD_NOINLINE inline
#endif

namespace qr {

void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define QR_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            char _b[512];                                                                \
            snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                \
            return qr::fail(_e == cudaErrorMemoryAllocation ? QR_ERR_NOMEM : QR_ERR_CUDA, _b); \
        }                                                                                \
    } while (0)

// One degree class of check nodes: `count` internal check slots starting at `chk_begin`, all of
// degree `degree`, whose edges occupy CSR slots slot_begin + (ci - chk_begin) * degree + i.
struct CheckBin {
    int32_t degree;
    int32_t chk_begin;
    int32_t count;
    int32_t slot_begin;
};

constexpr int kMaxCheckDegree = 64;  // larger checks are rejected at graph creation

}  // namespace qr

// Tanner graph handle.  Host tables are always built; device copies only when device >= 0.
struct qr_graph {
    int64_t N = 0, C = 0, E = 0;
    int device = -1;
    int32_t max_cdeg = 0, max_vdeg = 0;
    int32_t var_deg = 0;  // > 0 when every variable node has this degree
    bool decodable = true;  // every check has degree 2..kMaxCheckDegree (what the decoder needs; Matrix takes any graph)
    std::vector<int32_t> chk_order;  // [C]   internal check slot -> original check id
    std::vector<int32_t> chk_ptr;    // [C+1] internal check slot -> first CSR slot
    std::vector<int32_t> slot_edge;  // [E]   CSR slot -> original edge id
    std::vector<int32_t> slot_var;   // [E]   CSR slot -> variable id
    std::vector<int32_t> var_ptr;    // [N+1]
    std::vector<int32_t> var_slot;   // [E]   per variable, CSR slots in ascending edge id
    std::vector<qr::CheckBin> bins;
    // variable ids sorted by degree (stable): neighbouring work items of the variable phase have the same degree,
    // so a warp runs one unrolled code path; empty when every variable has the same degree
    std::vector<int32_t> var_work;   // [N]
    int32_t *d_var_work = nullptr;
    // ... and, in that order, one bin per variable degree (CheckBin reused: degree, first position in var_work,
    // count, first entry of vslot_sorted) with the variables' slot lists stored contiguously per bin: the variable
    // phase walks a bin with compile-time degree and prefetches the next item's indices (irregular graphs only)
    std::vector<qr::CheckBin> var_bins;
    std::vector<int32_t> vslot_sorted;   // [E]
    qr::CheckBin *d_var_bins = nullptr;
    int32_t *d_vslot_sorted = nullptr;
    // fused schedule: per CSR slot a 16-byte neighbour record (Nbr4, qr_decode_fused.cuh); empty when the graph is
    // outside its limits (2^27 variables, variable degree 64)
    std::vector<int32_t> slot_nbr;   // [4 * E]
    int32_t *d_slot_nbr = nullptr;
    // device copies
    int32_t *d_chk_order = nullptr, *d_chk_ptr = nullptr, *d_slot_var = nullptr;
    int32_t *d_var_ptr = nullptr, *d_var_slot = nullptr;
    qr::CheckBin *d_bins = nullptr;
};
