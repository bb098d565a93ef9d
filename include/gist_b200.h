/*
 * gist_b200 — C ABI of the B200-native GIST aggregation hot path.
 *
 * The reference (wolfecameron/GIST) has no FFI of its own: its hot path is a
 * handful of DGL 0.5.3 / PyTorch calls made from Python modules.  Each entry
 * point below names the reference call site(s) it replaces (paths relative to
 * the reference root).  The Python side of this repo (gist_b200/_lib.py) binds
 * these symbols with ctypes; INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t) and returns;
 *     no implicit synchronisation, no allocation, no global mutable state
 *     (except the launch counter) — the caller owns all buffers;
 *   - return value: 0 = OK, <0 = argument error (GIST_ERR_*), >0 = cudaError_t;
 *   - graph structure is int32 CSR over in-edges: row v lists the sources u of
 *     all edges u->v (multi-edges repeated, self loops allowed), no values;
 *   - dense matrices are row-major fp32 with an explicit leading dimension (in
 *     elements), so a kernel can read/write a column block of a wider buffer.
 */
#ifndef GIST_B200_H
#define GIST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *gist_stream_t; /* cudaStream_t */

/* Counter-based dropout (nn.Dropout as used at cluster_gcn/modules.py:228-229), fused into the
 * kernels on either side of it instead of running as a kernel of its own.  Whether element
 * (row, col) of the logical [n, width] matrix is kept is a pure function (Philox4x32-10) of
 * (seed, *step, stream_id, row, col): the forward kernel that applies the mask and the backward
 * kernel that needs it again regenerate it, nothing is stored.  `step` is a DEVICE counter read by
 * the kernel, so a captured CUDA graph draws a fresh mask at every replay (advance it with
 * gist_counter_add_i64 between steps); `stream_id` separates layers drawing from the same step.
 * Kept elements are scaled by 1 / (1 - p).  A NULL descriptor or p == 0 disables dropout. */
typedef struct gist_dropout {
    float p;               /* drop probability in [0, 1) */
    uint64_t seed;
    uint32_t stream_id;
    const int64_t *step;   /* device pointer; NULL = step 0 */
    int64_t *step_saved;   /* device pointer, optional: the applying kernel stores the step value it
                              read, to be passed as `step` to the matching backward call */
} gist_dropout_t;

#define GIST_ABI_VERSION 1

#define GIST_OK 0
#define GIST_ERR_BADARG (-1)      /* null pointer / negative size / bad enum      */
#define GIST_ERR_ALIGN (-2)       /* pointer not aligned for its element type     */
#define GIST_ERR_WORKSPACE (-3)   /* workspace or output capacity too small       */
#define GIST_ERR_UNSUPPORTED (-4) /* combination not implemented                  */

/* flags of gist_spmm_csr_f32 */
#define GIST_SPMM_RELU 1u        /* y = max(y, 0) as the last epilogue step       */
#define GIST_SPMM_NARROW 2u      /* force 1 vector / lane (narrow feature chunks) */
#define GIST_SPMM_WIDE 4u        /* force 2 vectors / lane                        */
#define GIST_SPMM_COOP_ON 8u     /* force CTA-cooperative hub rows (row split)    */
#define GIST_SPMM_COOP_OFF 16u   /* never split rows (default: on iff n_dst<=32k) */
#define GIST_SPMM_SLAB_OFF 32u   /* scheduled launches: never stage column slabs in shared memory (A/B switch) */
/* bits 8..11: background mode for gist_spmm_csr_ex_f32's extended epilogue (0 = off): at most
 * this many CTAs per SM are launched and walk the work with a grid stride, so a kernel that runs
 * in the shadow of a latency-critical branch (the next batch's layer-0 aggregation beside the
 * training step) never occupies every register file / warp slot of the chip. */
#define GIST_SPMM_BG_SHIFT 8
/* bits 12..14: force the lane-group width of the row-per-group kernel (0 = chosen from d and, for
 * operands larger than L2, from the L2 slab budget): 1, 2, 3, 4 -> 4, 8, 16, 32 lanes per row, i.e.
 * feature chunks of lanes x vector-width floats.  Tuning / measurement switch. */
#define GIST_SPMM_LANES_SHIFT 12
#define GIST_SPMM_LANES(code) ((uint32_t)(code) << GIST_SPMM_LANES_SHIFT)
/* bits 16..17: cap the gather width (0 = widest the alignment allows, narrowed for operands far
 * larger than L2 — see gist_spmm_csr_ex_f32): 1, 2, 3 -> 32-, 64-, 128-bit.  Measurement switch. */
#define GIST_SPMM_VEC_SHIFT 16

/* modes of gist_degree_norm_f32 */
#define GIST_NORM_INV 0       /* 1/deg, deg==0 -> 0   (ISTSAGELayer.get_norm, cluster_gcn/modules.py:239-243) */
#define GIST_NORM_RSQRT_CLAMP 1 /* clamp(deg,1)^-1/2  (DGL GraphConv norm='both')           */

int gist_abi_version(void);
const char *gist_status_string(int status);

/* Bind the calling thread's runtime to `device` (one process per GPU; call once
 * after torch.cuda.set_device). */
int gist_set_device(int device);

/* Number of kernels this library has launched in this process (bench.py's
 * gpu_launches evidence). */
uint64_t gist_launch_count(void);

/* Programmatic dependent launch for the kernels of the training step's critical chain (K4 and its
 * split-K second pass, the layer-norm kernels, the segment SpMM, the fused cross entropy, Adam, the
 * 3xTF32 weight split): launched with cudaLaunchAttributeProgrammaticStreamSerialization, each of
 * them becomes resident and runs its prologue while its predecessor on the stream drains, and blocks
 * in griddepcontrol.wait — before its first global-memory access — until the predecessor has
 * completed.  Results are unchanged (same stream-order dependencies); under stream capture the
 * attribute becomes a programmatic edge of the graph.  Process-wide switch; the initial value comes
 * from GIST_PDL in the environment (default off).  There is no reference counterpart: the reference
 * launches one eager PyTorch kernel after the other (cluster_gcn_ist_distrib.py:408-417). */
int gist_set_pdl(int32_t enabled);
int gist_get_pdl(void);

/* ------------------------------------------------------------------------
 * K1 / K2  CSR SpMM with fused epilogue.
 *
 *   acc[v, c]  = sum_{e in [rowptr[v], rowptr[v+1])}  s(col[e]) * X[col[e], c]
 *   Y[v, c]    = act( t(v) * acc[v, c] + addend[v, c] + bias[c] ),  c in [0, d)
 *   self_out[v, c] = X[v, c]                     (optional, needs n_dst == n_src)
 *
 * s = src_scale (or 1), t = dst_scale (or 1); addend / bias / self_out optional
 * (NULL).  Edges of a row are accumulated sequentially in CSR order, so the
 * result is run-to-run deterministic; no atomics.
 *
 * Replaces: g.update_all(fn.copy_src, fn.sum)   cluster_gcn/modules.py:136-137, :224-225,
 *           cluster_gcn/sampler.py:64-66; `ah * norm` + torch.cat  modules.py:226-227;
 *           DGL GraphConv's src/dst normalisation, bias and activation
 *           (gcn/gcn.py:30-56, cluster_gcn/modules.py:331-338).
 * Backward (K2): call it with the CSC arrays (out-edge lists) of the same graph
 * and src/dst scales swapped — dX[u] = s(u) * sum_{u->v} t(v) dY[v]  (what DGL's
 * GSpMM.backward does on the reversed graph).
 *
 * Vector width (128/64/32-bit gathers) is picked from the alignment of every
 * pointer / leading dimension involved.
 */
int gist_spmm_csr_f32(const int32_t *rowptr, const int32_t *col, int32_t n_dst, int32_t n_src,
                      const float *X, int64_t ldx, int32_t d,
                      float *Y, int64_t ldy,
                      const float *src_scale, const float *dst_scale, const float *bias,
                      const float *addend, int64_t ld_addend,
                      float *self_out, int64_t ld_self,
                      uint32_t flags, gist_stream_t stream);

/* Extended epilogue for the training step (ex == NULL: identical to gist_spmm_csr_f32):
 *   - dropout applied to Y and to self_out as they are written; the mask is that of the logical
 *     matrix whose columns [drop_col0_y, +d) / [drop_col0_self, +d) the two outputs are — for the
 *     SAGE concat z = [h | mean-agg(h)] (cluster_gcn/modules.py:226-229) that is d and 0;
 *   - y_lo / self_lo (optional): the 3xTF32 low halves tf32(x - trunc_tf32(x)) of the values
 *     written, so the nn.Linear that consumes z needs no separate split pass.
 * Replaces nn.Dropout on z (modules.py:228-229) and gist_split_tf32_f32 on z. */
/* Segment schedule of a (small, power-law) graph: every row cut into segments of <= seg_len edges
 * so that all warps of a launch gather the same amount (gist_spmm_schedule_build, once per cluster
 * batch; shared by the forward and — for a symmetric pattern — the transpose launches).
 *   seg_ptr[n+1]  first segment of each row (seg_ptr[n] = number of segments, read on the device)
 *   seg_row[max_segments]  row of each segment;  max_segments >= n + nnz / seg_len (host bound)
 *   counters[1 + chunks * n]  uint32, ZERO on entry, left zero on exit; chunks = ceil(d / 8) covers
 *                  every kernel variant: [0] head of the launch's work queue, then one arrival
 *                  counter per (feature chunk, row)
 *   workspace[max_segments][ld_workspace >= d, % 4 == 0]  fp32 scratch, no initialisation needed
 * Two launches that share `counters` / `workspace` must not run concurrently. */
typedef struct gist_spmm_schedule {
    const int32_t *seg_ptr;
    const int32_t *seg_row;
    int32_t seg_len;
    int64_t max_segments;
    uint32_t *counters;
    float *workspace;
    int64_t ld_workspace;
    /* optional (NULL / 0 = the behaviour above): */
    const int32_t *seg_meta; /* [max_segments][4] from gist_spmm_schedule_build_meta, 16-byte aligned: per segment
                                {row, first edge, first segment of the row, (edges << 20) | segments of the row} — a
                                warp item then starts with ONE 16-byte load instead of seg_row -> (seg_ptr, rowptr) */
    uint32_t flags;          /* GIST_SPMM_SCHED_* */
} gist_spmm_schedule_t;
#define GIST_SPMM_SCHED_PREFETCH 1u /* fetch the next work item from the queue before processing the current one */

typedef struct gist_spmm_ex {
    float *y_lo;
    int64_t ld_y_lo;
    float *self_lo;
    int64_t ld_self_lo;
    const gist_dropout_t *drop;
    int32_t drop_col0_y, drop_col0_self;
    const gist_spmm_schedule_t *schedule; /* NULL: one row per lane group (+ CTA-cooperative hub rows) */
} gist_spmm_ex_t;

size_t gist_spmm_schedule_workspace_bytes(int32_t n);
int gist_spmm_schedule_build(const int32_t *rowptr, int32_t n, int32_t seg_len, int32_t *seg_ptr,
                             int32_t *seg_row, int64_t max_segments, void *scan_ws, size_t scan_ws_bytes,
                             gist_stream_t stream);
/* The same, also filling seg_meta ([max_segments][4] int32, 16-byte aligned; NULL = skip).  The packed record
 * needs seg_len <= 4095 and max_segments < 2^20 (GIST_ERR_UNSUPPORTED otherwise: use the plain schedule). */
int gist_spmm_schedule_build_meta(const int32_t *rowptr, int32_t n, int32_t seg_len, int32_t *seg_ptr,
                                  int32_t *seg_row, int32_t *seg_meta, int64_t max_segments, void *scan_ws,
                                  size_t scan_ws_bytes, gist_stream_t stream);
int gist_spmm_csr_ex_f32(const int32_t *rowptr, const int32_t *col, int32_t n_dst, int32_t n_src,
                         const float *X, int64_t ldx, int32_t d,
                         float *Y, int64_t ldy,
                         const float *src_scale, const float *dst_scale, const float *bias,
                         const float *addend, int64_t ld_addend,
                         float *self_out, int64_t ld_self,
                         uint32_t flags, const gist_spmm_ex_t *ex, gist_stream_t stream);

/* K2 spelled out: identical kernel, arguments named for the transpose.
 * colptr/row = CSC of the forward graph. */
int gist_spmm_csc_f32(const int32_t *colptr, const int32_t *row, int32_t n_src, int32_t n_dst,
                      const float *dY, int64_t lddy, int32_t d,
                      float *dX, int64_t lddx,
                      const float *dst_scale, const float *src_scale,
                      const float *addend, int64_t ld_addend,
                      uint32_t flags, gist_stream_t stream);

/* out[v] = f(rowptr[v+1]-rowptr[v]);  replaces g.in_degrees() + get_norm
 * (cluster_gcn/modules.py:155-159, :239-243; sampler.py:72-76) and GraphConv's
 * degree clamps. */
int gist_degree_norm_f32(const int32_t *rowptr, int32_t n, int32_t mode, float *out,
                         gist_stream_t stream);

/* Exclusive prefix sum of n int32 values into out[0..n] (out[n] = total).
 * workspace: gist_scan_workspace_bytes(n) bytes. in may alias out. */
size_t gist_scan_workspace_bytes(int32_t n);
int gist_exclusive_scan_i32(const int32_t *in, int32_t n, int32_t *out, void *workspace,
                            size_t workspace_bytes, gist_stream_t stream);

/* ------------------------------------------------------------------------
 * K3  device-side cluster-batch builder (node-induced subgraph + relabel).
 *
 * new node i <-> nids[i] (order kept; nids must be unique); keeps every parent
 * edge whose two endpoints are selected, with multiplicity; within a row the
 * parent's edge order is kept, so the output is deterministic.
 *
 *   node_map   [n_parent] int32 scratch, all -1 on entry, all -1 again on exit
 *   out_rowptr [n_b + 1], out_col [col_capacity]  (capacity >= sum of parent
 *              degrees of nids is always enough); out_rowptr[n_b] = nnz_b
 *   out_inv_deg optional [n_b]: 1/in-degree of the batch graph, 0 if isolated
 *   scan_ws    gist_scan_workspace_bytes(n_b) bytes
 * If the capacity is exceeded no out-of-bounds write happens and
 * *overflow_flag (device int32, optional) is set to 1.
 *
 * Replaces: DGLGraph.subgraph on the CPU + per-step H2D of structure
 * (cluster_gcn/partition_utils.py:20-25, cluster_gcn_ist_distrib.py:409).
 */
int gist_cluster_batch_build(const int32_t *parent_rowptr, const int32_t *parent_col,
                             int32_t n_parent, const int64_t *nids, int32_t n_b,
                             int32_t *node_map, int32_t *out_rowptr, int32_t *out_col,
                             int64_t col_capacity, float *out_inv_deg, int32_t *overflow_flag,
                             void *scan_ws, size_t scan_ws_bytes, gist_stream_t stream);

/* gist_cluster_batch_build with the parent rows cut into chunks of 128 edges and one warp per CHUNK
 * instead of one warp per row (hub rows: one CTA): the same CSR, bit for bit, but the two walks over
 * the batch's parent rows (~1 M edges for a Reddit-shape batch) are one round of loads per warp on a
 * chip-wide grid instead of a launch as long as its unluckiest CTA.
 * max_chunks : capacity of the chunk arrays, >= sum_i ceil(parent_degree(nids[i]) / 128); any bound W
 *              on the batch's parent-degree sum gives W / 128 + n_b.  More chunks than that set
 *              *overflow_flag (if given) and the result is truncated.
 * workspace  : gist_cluster_batch_build_v2_workspace_bytes(n_b, max_chunks) bytes, 16-byte aligned.
 * Replaces the same reference calls as gist_cluster_batch_build. */
size_t gist_cluster_batch_build_v2_workspace_bytes(int32_t n_b, int64_t max_chunks);
int gist_cluster_batch_build_v2(const int32_t *parent_rowptr, const int32_t *parent_col, int32_t n_parent,
                                const int64_t *nids, int32_t n_b, int32_t *node_map, int32_t *out_rowptr,
                                int32_t *out_col, int64_t col_capacity, float *out_inv_deg,
                                int32_t *overflow_flag, int64_t max_chunks, void *workspace,
                                size_t workspace_bytes, gist_stream_t stream);

/* dst[i, :] = src[idx[i], :] for rows of row_bytes bytes (features, labels,
 * masks — the ndata row-gather of DGLGraph.subgraph). Strides in bytes. */
int gist_gather_rows(const void *src, int64_t src_stride_bytes, const int64_t *idx, int64_t n,
                     void *dst, int64_t dst_stride_bytes, int64_t row_bytes, gist_stream_t stream);

/* ------------------------------------------------------------------------
 * K5  GIST sub-model split / merge (2-D slice gather / scatter).
 *
 *   gather : dst[r, c] = src[ridx[r], cidx[c]]      dst is [n_rows, n_cols]
 *   scatter: dst[ridx[r], cidx[c]] = src[r, c]      src is [n_rows, n_cols]
 * ridx / cidx are int64 (torch.LongTensor, as create_partition returns) and
 * may be NULL = identity.  Indices must be unique for scatter.
 *
 * Replaces the advanced-indexing copies of dispatch / sync:
 * cluster_gcn/cluster_gcn_ist_distrib.py:107-133, :204-226, :291-313 and
 * gcn/train_ist.py:179-191, :244-285.
 */
int gist_slice_gather_f32(const float *src, int64_t ld_src, const int64_t *ridx, int64_t n_rows,
                          const int64_t *cidx, int64_t n_cols, float *dst, int64_t ld_dst,
                          gist_stream_t stream);
int gist_slice_scatter_f32(const float *src, int64_t ld_src, const int64_t *ridx, int64_t n_rows,
                           const int64_t *cidx, int64_t n_cols, float *dst, int64_t ld_dst,
                           gist_stream_t stream);

/* Several slice gathers (scatter == 0) or scatters (scatter != 0) in ONE launch: a GIST round boundary
 * moves 2 (L + 1) tensors per site (cluster_gcn_ist_distrib.py:100-195, 285-367), and with eight sites
 * the issue time of 48 small launches is what the step pipeline waits for.  `jobs` is a HOST array;
 * pointers inside are device pointers.  A source may be peer memory (an NVLink-mapped buffer of another
 * rank): the merge then reads every site's packed slices straight from that site's HBM while
 * scattering into the local replica — the all-gather and the scatter as one kernel. */
typedef struct gist_slice_job {
    const float *src;
    int64_t ld_src;
    const int64_t *ridx;   /* NULL = identity */
    int64_t n_rows;
    const int64_t *cidx;   /* NULL = identity */
    int64_t n_cols;
    float *dst;
    int64_t ld_dst;
} gist_slice_job_t;
int gist_slice_multi_f32(int32_t scatter, int32_t n_jobs, const gist_slice_job_t *jobs, gist_stream_t stream);

/* The merge of the ULTRA-WIDE layers (sync_model, cluster_gcn_ist_distrib.py:285-367; hidden 32768 split
 * 8 ways: every site's [4096, 8192] slice goes to random columns of its 4096 rows of a [32768, 65536]
 * replica): as 4-byte scatters that is one 32-byte sector per element, touched in random order over an
 * 8.6 GB matrix (measured: 11 ms at ~1 TB/s of DRAM traffic).  Here a CTA owns a few destination ROWS and
 * walks them left to right: per 32-byte sector it looks up which source column (if any) lands on each of
 * its 8 elements (inv_col), skips sectors nothing lands on, and otherwise reads the sector, patches it and
 * writes it back whole — the same bytes in ascending address order, whole sectors only.  When two source
 * rows fit 72 KB the CTA stages them in shared memory first: every source element is used once, at a random
 * moment of the sweep, and would otherwise have to survive in L2 next to the destination stream.
 *   dst[ridx[r], cc] = src[r, inv_col[cc]]   for every cc with inv_col[cc] >= 0       (ridx NULL = identity)
 * inv_col: int32 [dst_cols], the inverse of the slice's column index (gist_index_invert_i32), -1 where no
 * source column lands.  Jobs that share a destination must own DISJOINT ROWS (the sites' output-dimension
 * partitions are): sectors are rewritten whole.  dst rows 32-byte aligned (dst, ld_dst % 8 == 0). */
typedef struct gist_slice_rows_job {
    const float *src;       /* [n_rows, ld_src] */
    int64_t ld_src;
    const int64_t *ridx;    /* [n_rows] destination rows, or NULL */
    int64_t n_rows;
    int64_t n_cols;         /* source columns (every value of inv_col is < n_cols) */
    const int32_t *inv_col; /* [dst_cols] */
    float *dst;
    int64_t ld_dst;
    int64_t dst_cols;
} gist_slice_rows_job_t;
int gist_slice_scatter_rows_f32(int32_t n_jobs, const gist_slice_rows_job_t *jobs, gist_stream_t stream);
/* inv[0 .. size) = -1, then inv[idx[k]] = k for k in [0, n) (idx unique, in range). */
int gist_index_invert_i32(const int64_t *idx, int64_t n, int32_t *inv, int64_t size, gist_stream_t stream);


/* ------------------------------------------------------------------------
 * K4  tensor-core GEMM (tcgen05 / TMEM, TF32 inputs, fp32 accumulate).
 *
 *   C[M,N] = op(A)[M,K] * op(B)[N,K]^T (+ bias[N]) (ReLU)        row-major fp32
 *
 * a_layout / b_layout say how each operand is stored:
 *   GIST_GEMM_K_MAJOR  : A is [M, K] row-major (K contiguous), lda >= K   (B: [N, K], ldb >= K)
 *   GIST_GEMM_MN_MAJOR : A is [K, M] row-major (M contiguous), lda >= M   (B: [K, N], ldb >= N)
 * so the three contractions of a linear layer need no transposed copies:
 *   y  = z W^T + b : A = z  [n, in]  K-major,  B = W [out, in] K-major
 *   dz = dy W      : A = dy [n, out] K-major,  B = W [out, in] MN-major (N = in, K = out)
 *   dW = dy^T z    : A = dy [n, out] MN-major (M = out, K = n), B = z [n, in] MN-major
 *
 * Replaces the cuBLAS sgemm behind nn.Linear / th.matmul and its autograd
 * (cluster_gcn/modules.py:144, :233; DGL GraphConv's matmul).  Requires 16-byte aligned A, B
 * and lda, ldb multiples of 4 (TMA global strides); M, N, K themselves are arbitrary (TMA
 * zero-fills the edges).  TF32 keeps 10 mantissa bits of the inputs: ~1e-3 relative on the
 * products; the aggregation (K1/K2) stays full fp32.
 *
 * workspace: gist_gemm_tf32_workspace_bytes(M, N, K, flags) bytes, 16-byte aligned; used for
 * deterministic split-K when the output has too few tiles to fill the chip (the dW
 * contraction).  NULL is allowed: the GEMM then runs unsplit.
 */
#define GIST_GEMM_RELU 1u
#define GIST_GEMM_NO_SPLITK 2u
#define GIST_GEMM_TILE_N64 4u     /* force the output-tile width (default: chosen from the grid) */
#define GIST_GEMM_TILE_N128 8u
#define GIST_GEMM_TILE_N256 16u
#define GIST_GEMM_BACKGROUND 32u  /* plan and launch for a third of the SMs: a contraction off the critical path
                                     (dW beside the dz chain) must not queue 140 long CTAs in front of it */
#define GIST_GEMM_KC(n) ((uint32_t)(n) << 8) /* 3xTF32 only: K blocks (of 32) chained per TMEM accumulator, 1..255; 0 = default 4 */
#define GIST_GEMM_K_MAJOR 0
#define GIST_GEMM_MN_MAJOR 1
size_t gist_gemm_tf32_workspace_bytes(int32_t M, int32_t N, int32_t K, uint32_t flags);
int gist_gemm_tf32(const float *A, int64_t lda, int32_t a_layout, const float *B, int64_t ldb,
                   int32_t b_layout, float *C, int64_t ldc, int32_t M, int32_t N, int32_t K,
                   const float *bias, uint32_t flags, void *workspace, size_t workspace_bytes,
                   gist_stream_t stream);

/* The launch plan the two GEMMs use for a shape (introspection for tests / tuning): output-tile
 * width, number of K splits and K blocks (of 32) per split.  Host-only, no launch. */
int gist_gemm_plan(int32_t M, int32_t N, int32_t K, uint32_t flags, int32_t three_pass, int32_t *tile_n,
                   int32_t *splits, int32_t *kblocks_per_split);

/* Diagnostic: per-CTA phase stamps of every following GEMM launch (tools/gemm_trace.py).  `buffer`
 * = device memory for max_ctas * gist_gemm_trace_slots() int64 (max_ctas >= the SM count), or NULL to
 * switch tracing off (the default; production launches carry a NULL pointer and skip the stamps).
 * Not thread-safe: a process-wide switch for profiling sessions. */
int gist_gemm_set_trace(void *buffer, int32_t max_ctas);
int gist_gemm_trace_slots(void);

/* Both operands K-major, no workspace (never splits K). */
int gist_gemm_tn_tf32(const float *A, int64_t lda, const float *B, int64_t ldb, float *C, int64_t ldc,
                      int32_t M, int32_t N, int32_t K, const float *bias, uint32_t flags,
                      gist_stream_t stream);

/* 3xTF32: the same contraction with fp32-level accuracy on the TF32 tensor cores.  Each operand
 * is passed as (x, x_lo) where x_lo = tf32(x - trunc_tf32(x)) comes from gist_split_tf32_f32 (same
 * shape and layout as x, own leading dimension, 16-byte aligned, ld % 4 == 0).  The tensor core
 * reads only the top 19 bits of x, so  A_lo*B + A*B_lo + A*B  (three MMAs per K step, fp32
 * accumulation in TMEM) equals the fp32 product up to the dropped A_lo*B_lo term: ~2^-20 relative
 * per product.  This is the mode that meets the reference-parity tolerance (1e-5 relative,
 * BASELINE north_star) for nn.Linear forward and gradients; gist_gemm_tf32 is the faster,
 * 1e-3-accurate variant.  The tensor core's fp32 accumulation truncates, so the kernel closes the
 * TMEM accumulator every GIST_GEMM_KC K blocks and sums the chunks with ordinary fp32 adds in the
 * epilogue warps' registers: the error does not grow with K.  Tiles are at most 128 wide
 * (GIST_GEMM_TILE_N256 is served as 128).  Same flags as gist_gemm_tf32; workspace from
 * gist_gemm_3xtf32_workspace_bytes. */
size_t gist_gemm_3xtf32_workspace_bytes(int32_t M, int32_t N, int32_t K, uint32_t flags);
int gist_gemm_3xtf32(const float *A, const float *A_lo, int64_t lda, int64_t lda_lo, int32_t a_layout,
                     const float *B, const float *B_lo, int64_t ldb, int64_t ldb_lo, int32_t b_layout,
                     float *C, int64_t ldc, int32_t M, int32_t N, int32_t K, const float *bias,
                     uint32_t flags, void *workspace, size_t workspace_bytes, gist_stream_t stream);

/* Extended epilogue of K4 — what the training step needs besides the product, done where the product is
 * produced instead of in kernels of their own (every field optional; ex == NULL or all-zero: the plain
 * contraction).  A_lo == B_lo == NULL selects single-pass TF32, both given 3xTF32.
 *   tile_counters / n_counters : (honoured only by builds with -DGIST_GEMM_INKERNEL_SPLITK, see
 *       gist_gemm_has_inkernel_splitk(); the default build ignores them and runs the two-kernel split-K,
 *       which measured faster) uint32 array, ZERO on entry and left zero (one counter per output
 *       tile; n_counters >= ceil(M/128) * ceil(N/64) always suffices).  With it a split-K launch is ONE
 *       kernel: every split writes its partial tile to the workspace and the CTA that arrives last at
 *       the tile's counter adds the partials in split order and runs the epilogue (deterministic, same
 *       order as the two-kernel form).  Two launches in flight must not share counters.
 *   rowsum : [M] — sum over K of op(A)[m, k].  For dW = dy^T z (A = dy, MN-major) that is the bias
 *       gradient db = colsum(dy) (cluster_gcn/modules.py:233's nn.Linear), computed by the tensor core
 *       as one extra 16-column MMA per K step against a tile of ones.  Tiles are then <= 128 wide;
 *       both operands MN-major only (the dy^T z layout), GIST_ERR_UNSUPPORTED otherwise.
 *   ln_out / ld_ln / ln_stats / ln_eps / ln_flags : LayerNorm without affine (+ ReLU with
 *       GIST_ACT_RELU) over the rows of C (modules.py:234-236), y = z W^T layout, 3xTF32, N <= 256:
 *       C receives the pre-norm values (+ bias), ln_out the normalised / activated ones, ln_stats
 *       [M][2] = (mean, rstd) for gist_layernorm_act_bwd_f32.  N <= 128: one tile holds the row and the
 *       norm runs in the GEMM's own epilogue (or in the split-K second pass); 128 < N <= 256: in the
 *       split-K second pass when the planner splits K, else as the row-wise kernel behind the GEMM.
 *   drop : as in gist_gemm_dropmask_f32 (forces a single K split).
 * Workspace: gist_gemm_ex_workspace_bytes with the same arguments.  GIST_ERR_UNSUPPORTED when a
 * requested fusion does not apply to the shape / layout (the caller runs the separate kernel). */
typedef struct gist_gemm_ex {
    uint32_t *tile_counters;
    int64_t n_counters;
    float *rowsum;
    float *ln_out;
    int64_t ld_ln;
    float *ln_stats;
    float ln_eps;
    uint32_t ln_flags;
    const gist_dropout_t *drop;
} gist_gemm_ex_t;
int gist_gemm_has_inkernel_splitk(void);
size_t gist_gemm_ex_workspace_bytes(int32_t M, int32_t N, int32_t K, uint32_t flags, int32_t three_pass,
                                    const gist_gemm_ex_t *ex);
int gist_gemm_ex_f32(const float *A, const float *A_lo, int64_t lda, int64_t lda_lo, int32_t a_layout,
                     const float *B, const float *B_lo, int64_t ldb, int64_t ldb_lo, int32_t b_layout,
                     float *C, int64_t ldc, int32_t M, int32_t N, int32_t K, const float *bias,
                     uint32_t flags, void *workspace, size_t workspace_bytes, const gist_gemm_ex_t *ex,
                     gist_stream_t stream);

/* C = dropout_mask(seed, *step, stream_id) * (op(A) op(B)^T): the dz = dy W contraction of the layer
 * whose INPUT z went through dropout (cluster_gcn/modules.py:228-233).  The multiplier of C[r, c] is
 * the one gist_spmm_csr_ex_f32 applied to z[r, c] in the forward pass (same descriptor, with
 * step = the forward's step_saved), regenerated in the GEMM epilogue — replaces the masked_scale
 * kernel of nn.Dropout's backward.  A_lo == B_lo == NULL: single-pass TF32; both given: 3xTF32.
 * Never splits K. */
int gist_gemm_dropmask_f32(const float *A, const float *A_lo, int64_t lda, int64_t lda_lo, int32_t a_layout,
                           const float *B, const float *B_lo, int64_t ldb, int64_t ldb_lo, int32_t b_layout,
                           float *C, int64_t ldc, int32_t M, int32_t N, int32_t K, uint32_t flags,
                           const gist_dropout_t *drop, gist_stream_t stream);

/* lo[r,c] = tf32_round(x[r,c] - trunc_tf32(x[r,c])) and, if hi != NULL, hi[r,c] = trunc_tf32(x[r,c])
 * (x with its 13 low mantissa bits cleared) for a [rows, cols] row-major matrix. */
int gist_split_tf32_f32(const float *x, int64_t ld_x, int32_t rows, int32_t cols, float *hi, int64_t ld_hi,
                        float *lo, int64_t ld_lo, gist_stream_t stream);

/* gist_split_tf32_f32 (lo only) for up to 16 matrices in one launch: the weights of every layer at
 * the head of a training step (nn.Linear weights of cluster_gcn/modules.py:191-211). */
int gist_split_tf32_multi_f32(int32_t n_tensors, const float *const *x, const int64_t *ld_x,
                              const int32_t *rows, const int32_t *cols, float *const *lo,
                              const int64_t *ld_lo, gist_stream_t stream);

/* dst[c, r] = src[r, c] for a [rows, cols] row-major matrix (dst is [cols, rows]). */
int gist_transpose_f32(const float *src, int64_t ld_src, int32_t rows, int32_t cols, float *dst,
                       int64_t ld_dst, gist_stream_t stream);

/* ------------------------------------------------------------------------
 * Row-wise pieces of the training step between the SpMM and the GEMM (fused.cu).
 */
#define GIST_ACT_RELU 1u

/* y[r,:] = act((x[r,:] - mean_r) * rstd_r): nn.LayerNorm(d, elementwise_affine=False) followed by
 * the layer's ReLU (cluster_gcn/modules.py:234-236).  stats (optional) receives (mean_r, rstd_r)
 * pairs, [n, 2] floats, 8-byte aligned, for the backward. */
int gist_layernorm_act_fwd_f32(const float *x, int64_t ldx, int32_t n, int32_t d, float eps,
                               uint32_t flags, float *y, int64_t ldy, float *stats, gist_stream_t stream);
/* dx = d/dx of the above given dy; x is the forward INPUT, stats from the forward. */
int gist_layernorm_act_bwd_f32(const float *dy, int64_t lddy, const float *x, int64_t ldx,
                               const float *stats, int32_t n, int32_t d, uint32_t flags, float *dx,
                               int64_t lddx, float *dx_lo, int64_t ld_lo, gist_stream_t stream);

/* Whole-tensor layer norm: y = (x - mean) * rstd with ONE mean / biased variance over all n*d
 * elements, no affine — F.layer_norm(h, h.shape) as gcn/gcn.py:65-66 and
 * cluster_gcn/modules.py:347-348 (BaselineGCN) apply it between GraphConv layers.
 * stats (optional, 2 floats) receives (mean, rstd) for the backward.  Two launches, per-CTA fp64
 * partial moments folded in a fixed order.  workspace: gist_tensor_layernorm_workspace_bytes(n, d)
 * bytes, 16-byte aligned. */
size_t gist_tensor_layernorm_workspace_bytes(int32_t n, int32_t d);
int gist_tensor_layernorm_fwd_f32(const float *x, int64_t ldx, int32_t n, int32_t d, float eps, float *y,
                                  int64_t ldy, float *stats, void *workspace, size_t workspace_bytes,
                                  gist_stream_t stream);
/* dx = rstd * (dy - mean(dy) - xhat * mean(dy * xhat)); x is the forward INPUT, stats from the forward. */
int gist_tensor_layernorm_bwd_f32(const float *dy, int64_t lddy, const float *x, int64_t ldx,
                                  const float *stats, int32_t n, int32_t d, float *dx, int64_t lddx,
                                  void *workspace, size_t workspace_bytes, gist_stream_t stream);

/* out[c] = sum_r x[r, c] (bias gradient of nn.Linear), two-phase fixed-order reduction.
 * workspace: gist_colsum_workspace_bytes(n, d) bytes. */
size_t gist_colsum_workspace_bytes(int32_t n, int32_t d);
int gist_colsum_f32(const float *x, int64_t ldx, int32_t n, int32_t d, float *out, void *workspace,
                    size_t workspace_bytes, gist_stream_t stream);

/* CrossEntropyLoss()(pred[mask], labels[mask]) (cluster_gcn_ist_distrib.py:413-414) without the
 * boolean-index gather: loss_out[0] = mean over masked rows of (logsumexp(logits[r]) -
 * logits[r, labels[r]]), loss_out[1] = 1 / #masked rows.  mask: one byte per row (torch.bool) or
 * NULL = all rows.  lse [n] and row_loss [n] are outputs kept for the backward. */
int gist_masked_ce_fwd_f32(const float *logits, int64_t ld, int32_t n, int32_t C, const int64_t *labels,
                           const uint8_t *mask, float *lse, float *row_loss, float *loss_out,
                           gist_stream_t stream);
/* dlogits[r,c] = mask_r * (softmax(logits[r])_c - [c == labels[r]]) * grad_out[0] / #masked rows;
 * columns [C, fill_cols) of every row are zeroed (row padding for the TMA-fed GEMMs). */
int gist_masked_ce_bwd_f32(const float *logits, int64_t ld, int32_t n, int32_t C, const int64_t *labels,
                           const uint8_t *mask, const float *lse, const float *loss_out,
                           const float *grad_out, float *dlogits, int64_t ldd, int32_t fill_cols,
                           float *dlogits_lo, gist_stream_t stream);

/* Loss and its gradient seed in ONE launch — what `loss = CrossEntropyLoss()(pred[mask], y[mask]);
 * loss.backward()` (cluster_gcn_ist_distrib.py:413-415) puts at the head of the backward pass:
 * loss_out[0] = mean masked-row loss, loss_out[1] = 1 / #masked rows, and
 * dlogits[r,c] = mask_r * (softmax(logits[r])_c - [c == labels[r]]) / #masked rows (d loss / d logits
 * for an upstream gradient of 1), padding columns zeroed, optional 3xTF32 low half as in
 * gist_masked_ce_bwd_f32.  Requires n > 0.  workspace: gist_masked_ce_fused_workspace_bytes(n);
 * sync: one zeroed uint32 the kernel leaves zero (never shared by two launches in flight). */
size_t gist_masked_ce_fused_workspace_bytes(int32_t n);
int gist_masked_ce_fused_f32(const float *logits, int64_t ld, int32_t n, int32_t C, const int64_t *labels,
                             const uint8_t *mask, float *dlogits, int64_t ldd, int32_t fill_cols,
                             float *dlogits_lo, float *loss_out, void *workspace, size_t workspace_bytes,
                             uint32_t *sync, gist_stream_t stream);

/* dx_lo (layer norm backward) / dlogits_lo (cross entropy backward): optional (NULL) 3xTF32 low
 * halves tf32(v - trunc_tf32(v)) of the gradient just written, in the same layout (dlogits_lo: same
 * leading dimension and padding as dlogits) — the nn.Linear backward that consumes the gradient
 * then needs no gist_split_tf32_f32 pass over it.
 *
 * out[r, c] = m(r, col0 + c) * x[r, c]  (x == NULL: the multiplier itself): nn.Dropout as a
 * stand-alone kernel with the SAME mask the fused kernels apply (tests; eager fall-backs). */
int gist_dropout_f32(const float *x, int64_t ldx, int32_t n, int32_t d, int32_t col0, float *out, int64_t ldo,
                     const gist_dropout_t *drop, gist_stream_t stream);

/* *counter += delta on the device: advances the dropout step between training steps (one node of
 * the captured step graph, before the build / train branches fork). */
int gist_counter_add_i64(int64_t *counter, int64_t delta, gist_stream_t stream);

/* torch.optim.Adam(lr, betas, eps, weight_decay) (amsgrad=False; cluster_gcn_ist_distrib.py:405-407,
 * :417) over n_tensors tensors in one launch (per 24 tensors).  params / grads / exp_avg /
 * exp_avg_sq / numel are HOST arrays of device pointers / element counts (read during the call).
 * step: device float, the number of updates done so far; incremented on the device.
 * counter: device uint32, zero on entry, zero again on exit. */
int gist_adam_multi_f32(int32_t n_tensors, float *const *params, const float *const *grads,
                        float *const *exp_avg, float *const *exp_avg_sq, const int64_t *numel, float lr,
                        float beta1, float beta2, float eps, float weight_decay, float *step,
                        uint32_t *counter, gist_stream_t stream);

/* gist_adam_multi_f32 plus the two things that follow optimizer.step() at the tail of a training step
 * (cluster_gcn_ist_distrib.py:417 is the last statement of the reference's loop body), done by the same
 * launch instead of as two more links of the step's dependency chain:
 *   params_lo : HOST array (or NULL) of optional device pointers — the 3xTF32 low half
 *               tf32(p - trunc_tf32(p)) of every UPDATED parameter, same flat layout as the parameter
 *               (what gist_split_tf32_multi_f32 would compute at the head of the next step);
 *   tick      : device int64 (or NULL), incremented once by the last CTA of the last launch — the
 *               dropout clock of the next step (gist_counter_add_i64).  The caller orders everything that
 *               still reads the clock (the batch-preparation branch) before this call. */
int gist_adam_multi_ex_f32(int32_t n_tensors, float *const *params, const float *const *grads,
                           float *const *exp_avg, float *const *exp_avg_sq, const int64_t *numel, float lr,
                           float beta1, float beta2, float eps, float weight_decay, float *step,
                           uint32_t *counter, float *const *params_lo, int64_t *tick, gist_stream_t stream);

/* ------------------------------------------------------------------------
 * K6  graph attention (GAT) message passing: fused edge-softmax + weighted SpMM.
 *
 * Replaces GATLayer.forward's apply_edges(edge_attention) + update_all(message_func,
 * reduce_func) (cluster_gcn/modules.py:10-65; Python UDFs over DGL's degree-bucketed mailboxes).
 * For one head with z = fc(h) [n, D] and attn = attn_fc.weight [2D] = [a_l | a_r]:
 *   scores[v] = (a_l.z[v], a_r.z[v])
 *   e_uv      = leaky_relu(scores[u].x + scores[v].y, negative_slope)     for every edge u -> v
 *   out[v]    = sum_u softmax_u(e_uv) z[u]          (rows without in-edges: zeros)
 *   lse[v]    = logsumexp_u e_uv                    (kept for the backward; 0 if no in-edges)
 * No per-edge tensor is materialised.  D <= 1024.
 */
/* All `heads` attention heads of a MultiHeadGATLayer (cluster_gcn/modules.py:67-76) in one launch per
 * kernel: z / out / dout / dz are [n, heads * D] (head h = columns [h D, (h + 1) D): the projections of all
 * heads are one GEMM), attn [heads, 2D], scores [heads, n, 2], lse [heads, n], dattn [heads, 2D].  heads = 1
 * is the single-head API below. */
int gist_gat_scores_heads_f32(const float *z, int64_t ldz, int32_t n, int32_t D, int32_t heads, const float *attn,
                              float *scores, gist_stream_t stream);
int gist_gat_aggregate_heads_f32(const int32_t *rowptr, const int32_t *col, int32_t n, const float *z, int64_t ldz,
                                 int32_t D, int32_t heads, const float *scores, float negative_slope, float *out,
                                 int64_t ldo, float *lse, gist_stream_t stream);
size_t gist_gat_backward_heads_workspace_bytes(int32_t n, int32_t D, int32_t heads);
int gist_gat_backward_heads_f32(const int32_t *rowptr, const int32_t *col, const int32_t *colptr, const int32_t *row,
                                int32_t n, const float *z, int64_t ldz, int32_t D, int32_t heads, const float *scores,
                                const float *lse, const float *attn, float negative_slope, const float *out,
                                int64_t ldo, const float *dout, int64_t lddo, float *dz, int64_t lddz, float *dattn,
                                void *workspace, size_t workspace_bytes, gist_stream_t stream);
int gist_gat_scores_f32(const float *z, int64_t ldz, int32_t n, int32_t D, const float *attn,
                        float *scores /* [n,2], 8-byte aligned */, gist_stream_t stream);
int gist_gat_aggregate_f32(const int32_t *rowptr, const int32_t *col, int32_t n, const float *z,
                           int64_t ldz, int32_t D, const float *scores, float negative_slope,
                           float *out, int64_t ldo, float *lse, gist_stream_t stream);
/* Backward of scores + aggregate together, given dout = dL/dout:
 *   dz    [n, D] = dL/dz   (through the aggregation AND through the attention scores)
 *   dattn [2D]   = dL/dattn
 * (rowptr, col) = in-edge CSR, (colptr, row) = out-edge lists (CSC) of the same graph.
 * Deterministic: two gather passes (CSR order, then CSC order), no atomics.
 * workspace: gist_gat_backward_workspace_bytes(n, D) bytes. */
size_t gist_gat_backward_workspace_bytes(int32_t n, int32_t D);
int gist_gat_backward_f32(const int32_t *rowptr, const int32_t *col, const int32_t *colptr,
                          const int32_t *row, int32_t n, const float *z, int64_t ldz, int32_t D,
                          const float *scores, const float *lse, const float *attn,
                          float negative_slope, const float *out, int64_t ldo, const float *dout,
                          int64_t lddo, float *dz, int64_t lddz, float *dattn, void *workspace,
                          size_t workspace_bytes, gist_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GIST_B200_H */
