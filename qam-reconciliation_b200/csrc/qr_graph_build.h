// Host-side construction of the graph tables (pure C++, no CUDA): shared by libqamrecon
// (qr_graph.cu) and by the CPU emulation harness under tests/emu.
#pragma once

#include <algorithm>

#include "qr_common.h"

namespace qr {

// strict: the decoder's requirements (every check of degree 2..kMaxCheckDegree) are enforced here, as
// Decoder.__cinit__ would have to; !strict (Matrix: any edge list, matrix.pyx:21-38 accepts checks of any degree,
// unused check ids included) only records whether the graph is decodable.
inline int build_host_tables(qr_graph &g, const int64_t *vid, const int64_t *cid, int64_t E, bool strict = true)
{
    if (!vid || !cid) return fail(QR_ERR_INVALID, "null edge array");
    if (E <= 0) return fail(QR_ERR_GRAPH, "edge list is empty");
    if (E >= (int64_t(1) << 31)) return fail(QR_ERR_GRAPH, "more than 2^31-1 edges");
    int64_t C = 0, N = 0;
    for (int64_t e = 0; e < E; ++e) {
        if (vid[e] < 0 || cid[e] < 0) return fail(QR_ERR_GRAPH, "negative node id in edge list");
        C = std::max(C, cid[e] + 1);
        N = std::max(N, vid[e] + 1);
    }
    if (C >= (int64_t(1) << 31) || N >= (int64_t(1) << 31))
        return fail(QR_ERR_GRAPH, "node id does not fit 31 bits");
    g.N = N; g.C = C; g.E = E;

    std::vector<int32_t> cdeg(C, 0), vdeg(N, 0);
    for (int64_t e = 0; e < E; ++e) { cdeg[cid[e]]++; vdeg[vid[e]]++; }
    g.max_cdeg = *std::max_element(cdeg.begin(), cdeg.end());
    g.max_vdeg = *std::max_element(vdeg.begin(), vdeg.end());
    g.var_deg = (*std::min_element(vdeg.begin(), vdeg.end()) == g.max_vdeg) ? g.max_vdeg : 0;
    g.decodable = g.max_cdeg <= kMaxCheckDegree;
    for (int64_t c = 0; c < C; ++c) g.decodable = g.decodable && cdeg[c] >= 2;
    for (int64_t c = 0; strict && c < C; ++c) {
        // the reference indexes its 2*(deg-1) scratch out of bounds for degree 1 and dereferences a
        // failed malloc for degree 0 (decoder.pyx:131-135, :337-342): reject instead
        if (cdeg[c] < 2) {
            char b[160];
            snprintf(b, sizeof(b), "check node %lld has degree %d; every check needs degree >= 2",
                     (long long)c, cdeg[c]);
            return fail(QR_ERR_GRAPH, b);
        }
    }
    if (strict && g.max_cdeg > kMaxCheckDegree) {
        char b[160];
        snprintf(b, sizeof(b), "check degree %d exceeds the supported maximum %d", g.max_cdeg,
                 kMaxCheckDegree);
        return fail(QR_ERR_GRAPH, b);
    }

    // internal check order: by degree, ties by id (counting sort) -> one contiguous bin per degree
    std::vector<int32_t> deg_count(g.max_cdeg + 2, 0);
    for (int64_t c = 0; c < C; ++c) deg_count[cdeg[c] + 1]++;
    for (int d = 0; d <= g.max_cdeg; ++d) deg_count[d + 1] += deg_count[d];
    g.chk_order.assign(C, 0);
    std::vector<int32_t> chk_slot(C);  // original check id -> internal slot
    {
        std::vector<int32_t> fill(deg_count.begin(), deg_count.end() - 1);
        for (int64_t c = 0; c < C; ++c) {
            int32_t s = fill[cdeg[c]]++;
            g.chk_order[s] = (int32_t)c;
            chk_slot[c] = s;
        }
    }
    g.chk_ptr.assign(C + 1, 0);
    for (int64_t s = 0; s < C; ++s) g.chk_ptr[s + 1] = g.chk_ptr[s] + cdeg[g.chk_order[s]];
    g.bins.clear();
    for (int d = 2; d <= g.max_cdeg; ++d) {
        int32_t cnt = deg_count[d + 1] - deg_count[d];
        if (cnt > 0) g.bins.push_back(CheckBin{d, deg_count[d], cnt, g.chk_ptr[deg_count[d]]});
    }

    g.slot_edge.assign(E, 0);
    g.slot_var.assign(E, 0);
    std::vector<int32_t> edge_slot(E);
    {
        std::vector<int32_t> fill(C, 0);
        for (int64_t e = 0; e < E; ++e) {
            int32_t s = chk_slot[cid[e]];
            int32_t slot = g.chk_ptr[s] + fill[s]++;
            g.slot_edge[slot] = (int32_t)e;
            g.slot_var[slot] = (int32_t)vid[e];
            edge_slot[e] = slot;
        }
    }
    g.var_ptr.assign(N + 1, 0);
    for (int64_t v = 0; v < N; ++v) g.var_ptr[v + 1] = g.var_ptr[v] + vdeg[v];
    g.var_slot.assign(E, 0);
    {
        std::vector<int32_t> fill(N, 0);
        for (int64_t e = 0; e < E; ++e) {
            int64_t v = vid[e];
            g.var_slot[g.var_ptr[v] + fill[v]++] = edge_slot[e];
        }
    }
    g.var_work.clear();
    if (g.var_deg == 0) {
        g.var_work.resize(N);
        for (int64_t v = 0; v < N; ++v) g.var_work[v] = (int32_t)v;
        std::stable_sort(g.var_work.begin(), g.var_work.end(),
                         [&](int32_t a, int32_t b) { return vdeg[a] < vdeg[b]; });
    }
    g.var_bins.clear(); g.vslot_sorted.clear();
    if (g.var_deg == 0) {
        g.vslot_sorted.resize(E);
        int32_t at = 0;
        for (int64_t pos = 0; pos < N; ++pos) {
            const int32_t v = g.var_work[pos], d = vdeg[v];
            if (g.var_bins.empty() || g.var_bins.back().degree != d) g.var_bins.push_back(CheckBin{d, (int32_t)pos, 0, at});
            g.var_bins.back().count++;
            for (int j = 0; j < d; ++j) g.vslot_sorted[at++] = g.var_slot[g.var_ptr[v] + j];
        }
    }
    // neighbour table of the fused schedule: what a check needs to rebuild post[v] = llr[v] + sum c2v
    // (decoder.pyx:291-293) of each of its variables without a stored posterior (record layout: Nbr4,
    // qr_decode_fused.cuh)
    g.slot_nbr.clear();
    // (every variable needs an edge: the posterior of a lane about to finish is stored by the check holding the
    // variable's first edge)
    if (N < (int64_t(1) << 27) && g.max_vdeg <= 64 && g.decodable && *std::min_element(vdeg.begin(), vdeg.end()) >= 1) {
        g.slot_nbr.assign(4 * (size_t)E, -1);
        for (int64_t s = 0; s < E; ++s) {
            const int32_t v = g.slot_var[s];
            const int32_t q0 = g.var_ptr[v], dv = g.var_ptr[v + 1] - q0;
            int32_t own = -1;
            for (int j = 0; j < dv; ++j)
                if (g.var_slot[q0 + j] == s) own = j;
            if (dv <= 3) {
                for (int j = 0; j < dv; ++j) g.slot_nbr[4 * s + 1 + j] = g.var_slot[q0 + j];
                g.slot_nbr[4 * s] = (int32_t)((uint32_t)v | ((uint32_t)own << 28) | ((uint32_t)dv << 30));
            } else {
                g.slot_nbr[4 * s] = v | 0x08000000;
                g.slot_nbr[4 * s + 1] = q0;
                g.slot_nbr[4 * s + 2] = dv;
                g.slot_nbr[4 * s + 3] = own;
            }
        }
    }
    return QR_OK;
}


}  // namespace qr
