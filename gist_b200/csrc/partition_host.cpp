// Host-side partition producer: k-way METIS over the training graph, the step the
// reference delegates to dgl.transform.metis_partition (cluster_gcn/partition_utils.py:11-18).
//
// METIS 5.x ships inside the CUDA toolkit as libmetis_static.a (built with 64-bit idx_t and
// 32-bit real_t; no header is installed, so the two entry points used are declared here).
// This library is host-only (no CUDA): it produces the node -> part vector that
// ClusterIter turns into `par_li`; everything downstream of it runs on the GPU.
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/gist_b200.h"
#include "../../include/gist_partition.h"

extern "C" {
typedef int64_t metis_idx_t;
typedef float metis_real_t;
int METIS_SetDefaultOptions(metis_idx_t *options);
int METIS_PartGraphKway(metis_idx_t *nvtxs, metis_idx_t *ncon, metis_idx_t *xadj, metis_idx_t *adjncy,
                        metis_idx_t *vwgt, metis_idx_t *vsize, metis_idx_t *adjwgt, metis_idx_t *nparts,
                        metis_real_t *tpwgts, metis_real_t *ubvec, metis_idx_t *options, metis_idx_t *objval,
                        metis_idx_t *part);
}

namespace {
constexpr int kMetisNOptions = 40;       // METIS_NOPTIONS
constexpr int kOptSeed = 8;              // METIS_OPTION_SEED
constexpr int kOptNumbering = 17;        // METIS_OPTION_NUMBERING
constexpr int kMetisOk = 1;              // METIS_OK
}  // namespace

// xadj[n+1] / adjncy[xadj[n]]: CSR of an UNDIRECTED simple graph (both directions present, no
// self loops) — gist_partition_symmetrize produces it from any in-CSR.  part[n] receives ids in
// [0, nparts); *edgecut the number of cut (undirected) edges.  Deterministic for a given seed.
extern "C" int gist_metis_part_kway(int64_t n, const int64_t *xadj, const int64_t *adjncy, int64_t nparts,
                                    int64_t seed, int64_t *part, int64_t *edgecut) {
    if (n < 0 || nparts <= 0 || !part || (n > 0 && !xadj) || (n > 0 && xadj[n] > 0 && !adjncy))
        return GIST_ERR_BADARG;
    if (n == 0) return GIST_OK;
    if (nparts == 1 || xadj[n] == 0) {            // METIS rejects k = 1; an empty graph has no cut
        for (int64_t i = 0; i < n; ++i) part[i] = nparts == 1 ? 0 : i % nparts;
        if (edgecut) *edgecut = 0;
        return GIST_OK;
    }
    metis_idx_t options[kMetisNOptions];
    METIS_SetDefaultOptions(options);
    options[kOptSeed] = seed;
    options[kOptNumbering] = 0;
    metis_idx_t nv = n, ncon = 1, np = nparts, objval = 0;
    const int r = METIS_PartGraphKway(&nv, &ncon, const_cast<metis_idx_t *>(xadj), const_cast<metis_idx_t *>(adjncy),
                                      nullptr, nullptr, nullptr, &np, nullptr, nullptr, options, &objval, part);
    if (edgecut) *edgecut = objval;
    return r == kMetisOk ? GIST_OK : GIST_ERR_UNSUPPORTED;
}

// Undirected simple view of a directed in-CSR (rowptr[n+1], col[nnz], int32 as the GPU graph
// stores it): edge {u, v} is present if u->v or v->u is, self loops and multi-edges are dropped.
// Two-call protocol: out_adjncy == NULL fills out_xadj[n+1] only (so the caller can size the
// second array), then the same call with out_adjncy fills the neighbour lists (sorted).
extern "C" int gist_partition_symmetrize(int64_t n, const int32_t *rowptr, const int32_t *col, int64_t *out_xadj,
                                         int64_t *out_adjncy) {
    if (n < 0 || !out_xadj || (n > 0 && (!rowptr || !col))) return GIST_ERR_BADARG;
    std::vector<int64_t> cnt((size_t)n + 1, 0);
    for (int64_t v = 0; v < n; ++v)
        for (int32_t e = rowptr[v]; e < rowptr[v + 1]; ++e) {
            const int32_t u = col[e];
            if (u < 0 || u >= n) return GIST_ERR_BADARG;
            if (u != v) { ++cnt[(size_t)v + 1]; ++cnt[(size_t)u + 1]; }
        }
    for (int64_t v = 0; v < n; ++v) cnt[(size_t)v + 1] += cnt[(size_t)v];
    std::vector<int64_t> adj((size_t)cnt[(size_t)n]);
    {
        std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
        for (int64_t v = 0; v < n; ++v)
            for (int32_t e = rowptr[v]; e < rowptr[v + 1]; ++e) {
                const int32_t u = col[e];
                if (u != v) { adj[(size_t)pos[(size_t)v]++] = u; adj[(size_t)pos[(size_t)u]++] = v; }
            }
    }
    // sort + unique per row
    int64_t w = 0;
    std::vector<int64_t> start((size_t)n + 1, 0);
    for (int64_t v = 0; v < n; ++v) {
        int64_t *b = adj.data() + cnt[(size_t)v], *e = adj.data() + cnt[(size_t)v + 1];
        std::sort(b, e);
        start[(size_t)v] = w;
        int64_t last = -1;
        for (int64_t *p = b; p < e; ++p)
            if (*p != last) { last = *p; adj[(size_t)w++] = last; }     // w <= p - adj.data(): in place is safe
    }
    start[(size_t)n] = w;
    for (int64_t v = 0; v <= n; ++v) out_xadj[v] = start[(size_t)v];
    if (out_adjncy)
        for (int64_t i = 0; i < w; ++i) out_adjncy[i] = adj[(size_t)i];
    return GIST_OK;
}
