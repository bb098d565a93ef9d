/* gist_partition.h — host-side partition producer (no CUDA): libgist_partition.so.
 *
 * Replaces dgl.transform.metis_partition as called by the reference at
 * cluster_gcn/partition_utils.py:11-18 (get_partition_list).  METIS 5.x itself comes from
 * the CUDA toolkit's libmetis_static.a (64-bit idx_t, 32-bit real_t), linked into this
 * library; its symbols are not re-exported.  The reference binding a maintainer adds is the
 * ctypes stub in gist_b200/partition.py.
 *
 * Conventions as include/gist_b200.h: extern "C", plain pointers and sizes, integer status
 * (0 = OK, negative = GIST_ERR_*), nothing thrown, no global state.  All buffers are HOST
 * memory owned by the caller.
 */
#ifndef GIST_PARTITION_H
#define GIST_PARTITION_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Undirected simple view of a directed in-CSR (rowptr[n+1], col[nnz], int32 as GistGraph stores
 * it): {u, v} is an edge if u->v or v->u is; self loops and multi-edges are dropped; neighbour
 * lists come out sorted.  Call once with out_adjncy == NULL to obtain out_xadj[n+1] (and hence
 * the size out_xadj[n] of the second array), then again with both. */
int gist_partition_symmetrize(int64_t n, const int32_t *rowptr, const int32_t *col, int64_t *out_xadj,
                              int64_t *out_adjncy);

/* k-way METIS (METIS_PartGraphKway, default options, given seed, edge-cut objective) of the
 * undirected graph (xadj, adjncy).  part[n] receives part ids in [0, nparts); *edgecut (optional)
 * the number of undirected edges cut.  Deterministic for a fixed (graph, nparts, seed). */
int gist_metis_part_kway(int64_t n, const int64_t *xadj, const int64_t *adjncy, int64_t nparts, int64_t seed,
                         int64_t *part, int64_t *edgecut);

#ifdef __cplusplus
}
#endif
#endif
