// Tanner-graph ingestion: (vid, cid) edge arrays -> check-grouped (CSR) and variable-grouped
// (CSC) edge permutations, on the host in O(E), then copied to the device.
//
// Stands in for Decoder.__cinit__ (reference decoder.pyx:93-146, whose __build_table :60-89 is an
// O(nodes*E) scan) and Matrix.__cinit__ (matrix.pyx:21-38).  The reference's per-node edge lists
// are in ascending edge id (:73-76); that order fixes the floating-point association of every
// sum and of the forward/backward recursion, so it is preserved: CSR slots of a check and the
// CSC list of a variable both enumerate edges by ascending original edge id.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

#include "qr_common.h"
#include "qr_graph_build.h"

namespace qr {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

const char *last_error() { return g_last_error.c_str(); }

bool fused_eligible(const qr_graph *g);   // qr_decode_fused.cu

template <typename T>
static int upload(T **dst, const std::vector<T> &src)
{
    QR_CUDA_CHECK(cudaMalloc((void **)dst, std::max<size_t>(src.size(), 1) * sizeof(T)));
    if (!src.empty())
        QR_CUDA_CHECK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return QR_OK;
}

}  // namespace qr

extern "C" {

int qr_abi_version(void) { return 1; }

const char *qr_last_error(void) { return qr::last_error(); }

int qr_device_count(int *count)
{
    if (!count) return qr::fail(QR_ERR_INVALID, "null pointer");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return QR_OK;
}

static int graph_create(const int64_t *h_vid, const int64_t *h_cid, int64_t n_edges, int device, bool strict,
                        qr_graph **out);

int qr_graph_create(const int64_t *h_vid, const int64_t *h_cid, int64_t n_edges, int device,
                    qr_graph **out)
{
    return graph_create(h_vid, h_cid, n_edges, device, true, out);
}

int qr_graph_create_any(const int64_t *h_vid, const int64_t *h_cid, int64_t n_edges, int device,
                        qr_graph **out)
{
    return graph_create(h_vid, h_cid, n_edges, device, false, out);
}

int qr_graph_fused_eligible(const qr_graph *g, int *eligible)
{
    if (!g || !eligible) return qr::fail(QR_ERR_INVALID, "null pointer");
    *eligible = qr::fused_eligible(g) ? 1 : 0;
    return QR_OK;
}

static int graph_create(const int64_t *h_vid, const int64_t *h_cid, int64_t n_edges, int device, bool strict,
                        qr_graph **out)
{
    if (!out) return qr::fail(QR_ERR_INVALID, "null output pointer");
    *out = nullptr;
    qr_graph *g = new (std::nothrow) qr_graph();
    if (!g) return qr::fail(QR_ERR_NOMEM, "out of host memory");
    int rc;
    try {
        rc = qr::build_host_tables(*g, h_vid, h_cid, n_edges, strict);
    } catch (const std::bad_alloc &) {
        rc = qr::fail(QR_ERR_NOMEM, "out of host memory building the graph tables");
    }
    if (rc != QR_OK) { delete g; return rc; }
    g->device = device;
    if (device >= 0) {
        int prev = 0;
        cudaGetDevice(&prev);
        auto body = [&]() -> int {
            QR_CUDA_CHECK(cudaSetDevice(device));
            int r;
            if ((r = qr::upload(&g->d_chk_order, g->chk_order))) return r;
            if ((r = qr::upload(&g->d_chk_ptr, g->chk_ptr))) return r;
            if ((r = qr::upload(&g->d_slot_var, g->slot_var))) return r;
            if ((r = qr::upload(&g->d_var_ptr, g->var_ptr))) return r;
            if ((r = qr::upload(&g->d_var_slot, g->var_slot))) return r;
            if ((r = qr::upload(&g->d_bins, g->bins))) return r;
            if (!g->slot_nbr.empty() && (r = qr::upload(&g->d_slot_nbr, g->slot_nbr))) return r;
            if (!g->var_work.empty() && (r = qr::upload(&g->d_var_work, g->var_work))) return r;
            if (!g->var_bins.empty() && (r = qr::upload(&g->d_var_bins, g->var_bins))) return r;
            if (!g->vslot_sorted.empty() && (r = qr::upload(&g->d_vslot_sorted, g->vslot_sorted))) return r;
            return QR_OK;
        };
        rc = body();
        cudaSetDevice(prev);
        if (rc != QR_OK) { qr_graph_destroy(g); return rc; }
    }
    *out = g;
    return QR_OK;
}

void qr_graph_destroy(qr_graph *g)
{
    if (!g) return;
    if (g->device >= 0) {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(g->device);
        cudaFree(g->d_chk_order); cudaFree(g->d_chk_ptr); cudaFree(g->d_slot_var);
        cudaFree(g->d_var_ptr); cudaFree(g->d_var_slot); cudaFree(g->d_bins); cudaFree(g->d_slot_nbr); cudaFree(g->d_var_work); cudaFree(g->d_var_bins); cudaFree(g->d_vslot_sorted);
        cudaSetDevice(prev);
    }
    delete g;
}

int qr_graph_info(const qr_graph *g, int64_t *n_vars, int64_t *n_checks, int64_t *n_edges,
                  int32_t *max_check_degree, int32_t *max_var_degree)
{
    if (!g) return qr::fail(QR_ERR_INVALID, "null graph");
    if (n_vars) *n_vars = g->N;
    if (n_checks) *n_checks = g->C;
    if (n_edges) *n_edges = g->E;
    if (max_check_degree) *max_check_degree = g->max_cdeg;
    if (max_var_degree) *max_var_degree = g->max_vdeg;
    return QR_OK;
}

int qr_graph_export(const qr_graph *g, int32_t *chk_order, int32_t *slot_edge, int32_t *slot_var,
                    int32_t *var_ptr, int32_t *var_slot)
{
    if (!g) return qr::fail(QR_ERR_INVALID, "null graph");
    if (chk_order) memcpy(chk_order, g->chk_order.data(), g->chk_order.size() * sizeof(int32_t));
    if (slot_edge) memcpy(slot_edge, g->slot_edge.data(), g->slot_edge.size() * sizeof(int32_t));
    if (slot_var) memcpy(slot_var, g->slot_var.data(), g->slot_var.size() * sizeof(int32_t));
    if (var_ptr) memcpy(var_ptr, g->var_ptr.data(), g->var_ptr.size() * sizeof(int32_t));
    if (var_slot) memcpy(var_slot, g->var_slot.data(), g->var_slot.size() * sizeof(int32_t));
    return QR_OK;
}

}  // extern "C"
