// K3 (cluster-batch builder), exclusive scan, row gather and K5 (GIST slice
// gather / scatter).  All HBM/L2-bound integer and copy work: coalesced index
// streams, one warp per adjacency row, order-preserving ballot compaction.
#include "common.cuh"

namespace gist {

// ---------------------------------------------------------------- scan ----
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048 elements per CTA pass

// Inclusive scan of one value per thread across a 256-thread CTA; returns the
// exclusive prefix of this thread and the CTA total.
__device__ __forceinline__ int block_exclusive_scan(int val, int &total) {
    __shared__ int warp_sums[kScanThreads / 32];
    __shared__ int s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = val;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = winc - w;  // exclusive warp offset
        if (lane == 31) s_total = winc;
    }
    __syncthreads();
    const int excl = inc - val + warp_sums[warp];
    total = s_total;
    __syncthreads();  // shared arrays are reused by the caller's next pass
    return excl;
}

// Single-CTA scan for small inputs (every cluster batch): loops over tiles with
// a running carry.  out has n+1 entries.
__global__ void __launch_bounds__(kScanThreads) scan_small_kernel(const int32_t *in, int32_t n,
                                                                 int32_t *out) {
    int carry = 0;
    for (int base = 0; base < n; base += kScanTile) {
        int v[kScanItems];
        int sum = 0;
        const int i0 = base + threadIdx.x * kScanItems;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            v[k] = (i0 + k < n) ? in[i0 + k] : 0;
            sum += v[k];
        }
        int total;
        int excl = block_exclusive_scan(sum, total) + carry;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (i0 + k < n) out[i0 + k] = excl;
            excl += v[k];
        }
        carry += total;
    }
    if (threadIdx.x == 0) out[n] = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const int32_t *in, int32_t n,
                                                                     int32_t *tile_sums) {
    const int64_t i0 = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) sum += (i0 + k < n) ? in[i0 + k] : 0;
    int total;
    block_exclusive_scan(sum, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int32_t *in, int32_t n,
                                                                 const int32_t *tile_offsets,
                                                                 int32_t n_tiles, int32_t *out) {
    const int64_t i0 = (int64_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (i0 + k < n) ? in[i0 + k] : 0;
        sum += v[k];
    }
    int total;
    int excl = block_exclusive_scan(sum, total) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (i0 + k < n) out[i0 + k] = excl;
        excl += v[k];
    }
    if (blockIdx.x == n_tiles - 1 && threadIdx.x == 0) out[n] = tile_offsets[n_tiles];
}

static int scan_launch(const int32_t *in, int32_t n, int32_t *out, void *ws, size_t ws_bytes,
                       cudaStream_t s) {
    if (n <= 4 * kScanTile) {
        scan_small_kernel<<<1, kScanThreads, 0, s>>>(in, n, out);
        count_launch();
        return last_error();
    }
    const int n_tiles = (n + kScanTile - 1) / kScanTile;
    if (ws_bytes < gist_scan_workspace_bytes(n) || !ws) return GIST_ERR_WORKSPACE;
    int32_t *tile_sums = reinterpret_cast<int32_t *>(ws);  // [n_tiles + 1]
    scan_tile_sums_kernel<<<n_tiles, kScanThreads, 0, s>>>(in, n, tile_sums);
    scan_small_kernel<<<1, kScanThreads, 0, s>>>(tile_sums, n_tiles, tile_sums);
    scan_apply_kernel<<<n_tiles, kScanThreads, 0, s>>>(in, n, tile_sums, n_tiles, out);
    count_launch(3);
    return last_error();
}

// ------------------------------------------------------- batch builder ----
// nids[i] < 0 is a padding sentinel: the row exists but is isolated (lets a caller pad
// every batch to a fixed node count, e.g. for CUDA-graph replay).
constexpr int kHeavyParentDeg = 1024;   // parent rows longer than this are split over the CTA
constexpr int kBuildUnroll = 4;         // 4 x 32 independent (col -> node_map) loads in flight

__global__ void batch_mark_kernel(const int64_t *__restrict__ nids, int32_t n_b,
                                  int32_t *__restrict__ node_map, int32_t value_is_index) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_b) {
        const int64_t p = nids[i];
        if (p >= 0) node_map[p] = value_is_index ? i : -1;
    }
}

__device__ __forceinline__ int count_range(const int32_t *__restrict__ pcol,
                                           const int32_t *__restrict__ node_map, int eb, int ee,
                                           int lane) {
    int cnt = 0;
    for (int e0 = eb; e0 < ee; e0 += 32 * kBuildUnroll) {
        int m[kBuildUnroll];
#pragma unroll
        for (int k = 0; k < kBuildUnroll; ++k) {
            const int e = e0 + k * 32 + lane;
            m[k] = (e < ee) ? __ldg(node_map + __ldg(pcol + e)) : -1;
        }
#pragma unroll
        for (int k = 0; k < kBuildUnroll; ++k) cnt += m[k] >= 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    return cnt;
}

// Order-preserving compaction of the in-batch neighbours of edges [eb, ee) to out_col[pos...].
__device__ __forceinline__ void fill_range(const int32_t *__restrict__ pcol,
                                           const int32_t *__restrict__ node_map, int eb, int ee,
                                           int lane, int64_t pos, int32_t *__restrict__ out_col,
                                           int64_t col_capacity) {
    for (int e0 = eb; e0 < ee; e0 += 32 * kBuildUnroll) {
        int m[kBuildUnroll];
#pragma unroll
        for (int k = 0; k < kBuildUnroll; ++k) {
            const int e = e0 + k * 32 + lane;
            m[k] = (e < ee) ? __ldg(node_map + __ldg(pcol + e)) : -1;
        }
#pragma unroll
        for (int k = 0; k < kBuildUnroll; ++k) {
            const unsigned bal = __ballot_sync(0xffffffffu, m[k] >= 0);
            if (m[k] >= 0) {
                const int64_t w = pos + __popc(bal & ((1u << lane) - 1u));
                if (w < col_capacity) out_col[w] = m[k];
            }
            pos += __popc(bal);
        }
    }
}

__device__ __forceinline__ void heavy_segment(int rs, int re, int warp, int &eb, int &ee) {
    int seg = (re - rs + 7) / 8;
    seg = (seg + 31) / 32 * 32;
    eb = min(re, rs + warp * seg);
    ee = min(re, eb + seg);
}

// One warp per batch row (8 rows per CTA); hub rows of the parent are counted by all
// 8 warps of the CTA in a second phase.
__global__ void __launch_bounds__(256) batch_count_kernel(const int32_t *__restrict__ prow,
                                                          const int32_t *__restrict__ pcol,
                                                          const int64_t *__restrict__ nids,
                                                          int32_t n_b,
                                                          const int32_t *__restrict__ node_map,
                                                          int32_t *__restrict__ deg) {
    __shared__ int s_heavy[8], s_cnt[8], s_nheavy;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (threadIdx.x == 0) s_nheavy = 0;
    __syncthreads();
    if (i < n_b) {
        const int64_t pnode = nids[i];
        int rs = 0, re = 0;
        if (pnode >= 0) { rs = prow[pnode]; re = prow[pnode + 1]; }
        if (re - rs > kHeavyParentDeg) {
            if (lane == 0) s_heavy[atomicAdd(&s_nheavy, 1)] = i;
        } else {
            const int cnt = count_range(pcol, node_map, rs, re, lane);
            if (lane == 0) deg[i] = cnt;
        }
    }
    __syncthreads();
    const int nheavy = s_nheavy;
    for (int h = 0; h < nheavy; ++h) {
        const int row = s_heavy[h];
        const int64_t pnode = nids[row];
        int eb, ee;
        heavy_segment(prow[pnode], prow[pnode + 1], warp, eb, ee);
        const int cnt = count_range(pcol, node_map, eb, ee, lane);
        if (lane == 0) s_cnt[warp] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w) t += s_cnt[w];
            deg[row] = t;
        }
        __syncthreads();
    }
}

// One warp per batch row: write relabelled neighbours, parent order preserved.
__global__ void __launch_bounds__(256) batch_fill_kernel(const int32_t *__restrict__ prow,
                                                         const int32_t *__restrict__ pcol,
                                                         const int64_t *__restrict__ nids,
                                                         int32_t n_b,
                                                         const int32_t *__restrict__ node_map,
                                                         const int32_t *__restrict__ out_rowptr,
                                                         int32_t *__restrict__ out_col,
                                                         int64_t col_capacity,
                                                         float *__restrict__ out_inv_deg,
                                                         int32_t *__restrict__ overflow_flag) {
    __shared__ int s_heavy[8], s_cnt[8], s_nheavy;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (threadIdx.x == 0) s_nheavy = 0;
    __syncthreads();
    if (i < n_b) {
        const int64_t pnode = nids[i];
        int rs = 0, re = 0;
        if (pnode >= 0) { rs = prow[pnode]; re = prow[pnode + 1]; }
        const int obeg = out_rowptr[i], oend = out_rowptr[i + 1];
        if (lane == 0) {
            const int dg = oend - obeg;
            if (out_inv_deg) out_inv_deg[i] = dg > 0 ? 1.0f / (float)dg : 0.f;
            if (oend > col_capacity && overflow_flag) *overflow_flag = 1;
        }
        if (re - rs > kHeavyParentDeg) {
            if (lane == 0) s_heavy[atomicAdd(&s_nheavy, 1)] = i;
        } else {
            fill_range(pcol, node_map, rs, re, lane, obeg, out_col, col_capacity);
        }
    }
    __syncthreads();
    const int nheavy = s_nheavy;
    for (int h = 0; h < nheavy; ++h) {
        const int row = s_heavy[h];
        const int64_t pnode = nids[row];
        int eb, ee;
        heavy_segment(prow[pnode], prow[pnode + 1], warp, eb, ee);
        const int cnt = count_range(pcol, node_map, eb, ee, lane);
        if (lane == 0) s_cnt[warp] = cnt;
        __syncthreads();
        int64_t pos = out_rowptr[row];
        for (int w = 0; w < warp; ++w) pos += s_cnt[w];
        fill_range(pcol, node_map, eb, ee, lane, pos, out_col, col_capacity);
        __syncthreads();
    }
}

// ----------------------------------------------------------- row gather ----
template <typename V>
__global__ void __launch_bounds__(256) gather_rows_kernel(const char *__restrict__ src,
                                                          int64_t src_stride,
                                                          const int64_t *__restrict__ idx, int64_t n,
                                                          char *__restrict__ dst, int64_t dst_stride,
                                                          int64_t vecs_per_row) {
    // One warp per destination row, lanes stride over the row's vectors.
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int64_t r = idx[i];
    V *d = reinterpret_cast<V *>(dst + i * dst_stride);
    if (r < 0) {   // padding sentinel: zero row
        V z;
        memset(&z, 0, sizeof(V));
        for (int64_t k = lane; k < vecs_per_row; k += 32) d[k] = z;
        return;
    }
    const V *s = reinterpret_cast<const V *>(src + r * src_stride);
    for (int64_t k = lane; k < vecs_per_row; k += 32) d[k] = __ldg(s + k);
}

// Narrow rows (labels, masks): one thread per row.
template <typename V>
__global__ void gather_elems_kernel(const char *__restrict__ src, int64_t src_stride,
                                    const int64_t *__restrict__ idx, int64_t n,
                                    char *__restrict__ dst, int64_t dst_stride) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t r = idx[i];
    V v;
    if (r < 0) memset(&v, 0, sizeof(V));
    else v = __ldg(reinterpret_cast<const V *>(src + r * src_stride));
    *reinterpret_cast<V *>(dst + i * dst_stride) = v;
}

// ------------------------------------------------------- slice (K5) -------
template <bool SCATTER>
__global__ void __launch_bounds__(256) slice_kernel(const float *__restrict__ src, int64_t ld_src,
                                                    const int64_t *__restrict__ ridx, int64_t n_rows,
                                                    const int64_t *__restrict__ cidx, int64_t n_cols,
                                                    float *__restrict__ dst, int64_t ld_dst) {
    // blockIdx.y walks rows (grid-stride), x covers columns: the dense side is
    // always accessed contiguously, the indexed side stays inside one row.
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const int64_t cc = cidx ? cidx[c] : c;
    for (int64_t r = blockIdx.y; r < n_rows; r += gridDim.y) {
        const int64_t rr = ridx ? ridx[r] : r;
        if (SCATTER) dst[rr * ld_dst + cc] = src[r * ld_src + c];
        else dst[r * ld_dst + c] = __ldg(src + rr * ld_src + cc);
    }
}

// Many slices in ONE launch: a GIST round boundary moves 2 (L + 1) tensors per site — with eight sites
// that is 48 launches whose host-side issue time, not their bytes, is what the step pipeline waits for
// (measured: 0.9 ms of stream time for 12 scatters of 0.8 MB at m = 2).  Sources may be PEER memory
// (NVLink-mapped buffers of the other ranks): the merge then reads every site's packed slices straight
// from that site's HBM and scatters them into the local replica — all-gather and scatter in one kernel.
constexpr int kSliceMaxJobs = 96;
struct SliceJobDev {
    const float *src;
    float *dst;
    const int64_t *ridx;
    const int64_t *cidx;
    int64_t ld_src, ld_dst;
    int32_t n_rows, n_cols;
    int32_t block0;     // first CTA of this job
    int32_t gx;         // column blocks (256 columns each); the job's other CTAs stride over rows
    int32_t gy;
    int32_t pad;
};
struct SliceJobsDev {
    SliceJobDev j[kSliceMaxJobs];
    int32_t n_jobs;
};

template <bool SCATTER>
__global__ void __launch_bounds__(256) slice_multi_kernel(const __grid_constant__ SliceJobsDev a) {
    int lo = 0, hi = a.n_jobs - 1;          // last job whose block0 <= blockIdx.x
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int)blockIdx.x >= a.j[mid].block0) lo = mid;
        else hi = mid - 1;
    }
    const SliceJobDev &J = a.j[lo];
    const int b = (int)blockIdx.x - J.block0;
    const int bx = b % J.gx, by = b / J.gx;
    const int64_t c = (int64_t)bx * 256 + threadIdx.x;
    if (c >= J.n_cols) return;
    const int64_t cc = J.cidx ? __ldg(J.cidx + c) : c;
    for (int64_t r = by; r < J.n_rows; r += J.gy) {
        const int64_t rr = J.ridx ? __ldg(J.ridx + r) : r;
        if (SCATTER) J.dst[rr * J.ld_dst + cc] = J.src[r * J.ld_src + c];     // plain load: src may be peer memory
        else J.dst[r * J.ld_dst + c] = J.src[rr * J.ld_src + cc];
    }
}

}  // namespace gist

using namespace gist;

extern "C" int gist_slice_multi_f32(int32_t scatter, int32_t n_jobs, const gist_slice_job_t *jobs,
                                    gist_stream_t stream) {
    if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return GIST_ERR_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    int done = 0;
    while (done < n_jobs) {
        SliceJobsDev a;
        a.n_jobs = 0;
        int64_t blocks = 0;
        // size the launch so that the whole batch of jobs is a few waves of CTAs
        int64_t total_elems = 0;
        const int first = done;
        int last = done;
        while (last < n_jobs && last - first < kSliceMaxJobs) {
            const gist_slice_job_t &j = jobs[last];
            if (j.n_rows < 0 || j.n_cols < 0) return GIST_ERR_BADARG;
            total_elems += j.n_rows * j.n_cols;
            ++last;
        }
        // a few waves for small rounds; for the ultra-wide slices (10^8 elements) ~16 rows x 256 columns per CTA,
        // so that the scattered 4-byte writes of a row block have thousands of CTAs in flight to hide behind
        int64_t target_blocks = 16LL * kNumSMs;
        if (total_elems / 4096 > target_blocks) target_blocks = total_elems / 4096;
        for (int i = first; i < last; ++i) {
            const gist_slice_job_t &j = jobs[i];
            if (j.n_rows == 0 || j.n_cols == 0) continue;
            if (!j.src || !j.dst || j.n_rows > 0x7fffffffLL || j.n_cols > 0x7fffffffLL) return GIST_ERR_BADARG;
            SliceJobDev &d = a.j[a.n_jobs++];
            d.src = j.src; d.dst = j.dst; d.ridx = j.ridx; d.cidx = j.cidx;
            d.ld_src = j.ld_src; d.ld_dst = j.ld_dst;
            d.n_rows = (int32_t)j.n_rows; d.n_cols = (int32_t)j.n_cols;
            d.gx = (int32_t)((j.n_cols + 255) / 256);
            // this job's share of the launch, by element count; at least one CTA per column block
            int64_t share = total_elems > 0 ? target_blocks * (j.n_rows * j.n_cols) / total_elems : 1;
            int64_t gy = share / d.gx;
            if (gy < 1) gy = 1;
            if (gy > j.n_rows) gy = j.n_rows;
            d.gy = (int32_t)gy;
            d.block0 = (int32_t)blocks;
            d.pad = 0;
            blocks += (int64_t)d.gx * d.gy;
            if (blocks > 0x7fffffffLL) return GIST_ERR_UNSUPPORTED;
        }
        done = last;
        if (a.n_jobs == 0) continue;
        if (scatter) slice_multi_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(a);
        else slice_multi_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(a);
        count_launch();
        const int st = last_error();
        if (st != GIST_OK) return st;
    }
    return GIST_OK;
}

extern "C" size_t gist_scan_workspace_bytes(int32_t n) {
    if (n <= 4 * kScanTile) return 0;
    const size_t n_tiles = ((size_t)n + kScanTile - 1) / kScanTile;
    return (n_tiles + 1) * sizeof(int32_t);
}

extern "C" int gist_exclusive_scan_i32(const int32_t *in, int32_t n, int32_t *out, void *workspace,
                                       size_t workspace_bytes, gist_stream_t stream) {
    if (n < 0 || !out || (n > 0 && !in)) return GIST_ERR_BADARG;
    return scan_launch(in, n, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int gist_cluster_batch_build(const int32_t *parent_rowptr, const int32_t *parent_col,
                                        int32_t n_parent, const int64_t *nids, int32_t n_b,
                                        int32_t *node_map, int32_t *out_rowptr, int32_t *out_col,
                                        int64_t col_capacity, float *out_inv_deg,
                                        int32_t *overflow_flag, void *scan_ws, size_t scan_ws_bytes,
                                        gist_stream_t stream) {
    if (n_parent < 0 || n_b < 0 || col_capacity < 0) return GIST_ERR_BADARG;
    if (!out_rowptr) return GIST_ERR_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (n_b == 0) {
        cudaError_t e = cudaMemsetAsync(out_rowptr, 0, sizeof(int32_t), s);
        return e == cudaSuccess ? GIST_OK : (int)e;
    }
    if (!parent_rowptr || !nids || !node_map) return GIST_ERR_BADARG;
    if (col_capacity > 0 && !out_col) return GIST_ERR_BADARG;
    if (scan_ws_bytes < gist_scan_workspace_bytes(n_b)) return GIST_ERR_WORKSPACE;
    const int tb = 256;
    const int g_thread = (n_b + tb - 1) / tb;
    const int g_warp = (n_b + 7) / 8;   // 8 warps (rows) per CTA
    batch_mark_kernel<<<g_thread, tb, 0, s>>>(nids, n_b, node_map, 1);
    // degrees land in out_rowptr[0..n_b) and are scanned in place
    batch_count_kernel<<<g_warp, tb, 0, s>>>(parent_rowptr, parent_col, nids, n_b, node_map,
                                             out_rowptr);
    count_launch(2);
    int st = scan_launch(out_rowptr, n_b, out_rowptr, scan_ws, scan_ws_bytes, s);
    if (st != GIST_OK) return st;
    batch_fill_kernel<<<g_warp, tb, 0, s>>>(parent_rowptr, parent_col, nids, n_b, node_map,
                                            out_rowptr, out_col, col_capacity, out_inv_deg,
                                            overflow_flag);
    batch_mark_kernel<<<g_thread, tb, 0, s>>>(nids, n_b, node_map, 0);
    count_launch(2);
    return last_error();
}

extern "C" int gist_gather_rows(const void *src, int64_t src_stride_bytes, const int64_t *idx,
                                int64_t n, void *dst, int64_t dst_stride_bytes, int64_t row_bytes,
                                gist_stream_t stream) {
    if (n < 0 || row_bytes < 0 || src_stride_bytes < 0 || dst_stride_bytes < row_bytes)
        return GIST_ERR_BADARG;
    if (n == 0 || row_bytes == 0) return GIST_OK;
    if (!src || !idx || !dst) return GIST_ERR_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    const char *sp = reinterpret_cast<const char *>(src);
    char *dp = reinterpret_cast<char *>(dst);
    auto ok = [&](int64_t a) {
        return row_bytes % a == 0 && src_stride_bytes % a == 0 && dst_stride_bytes % a == 0 &&
               aligned(src, (size_t)a) && aligned(dst, (size_t)a);
    };
    if (row_bytes <= 8 && (row_bytes == 8 || row_bytes == 4 || row_bytes == 2 || row_bytes == 1) &&
        ok(row_bytes)) {
        const unsigned g = (unsigned)((n + 255) / 256);
        if (row_bytes == 8) gather_elems_kernel<int2><<<g, 256, 0, s>>>(sp, src_stride_bytes, idx, n, dp, dst_stride_bytes);
        else if (row_bytes == 4) gather_elems_kernel<int><<<g, 256, 0, s>>>(sp, src_stride_bytes, idx, n, dp, dst_stride_bytes);
        else if (row_bytes == 2) gather_elems_kernel<short><<<g, 256, 0, s>>>(sp, src_stride_bytes, idx, n, dp, dst_stride_bytes);
        else gather_elems_kernel<char><<<g, 256, 0, s>>>(sp, src_stride_bytes, idx, n, dp, dst_stride_bytes);
        count_launch();
        return last_error();
    }
    const int64_t blocks = (n * 32 + 255) / 256;
    if (blocks > 0x7fffffffLL) return GIST_ERR_UNSUPPORTED;
    const unsigned g = (unsigned)blocks;
    if (ok(16)) gather_rows_kernel<int4><<<g, 256, 0, s>>>(sp, src_stride_bytes, idx, n, dp, dst_stride_bytes, row_bytes / 16);
    else if (ok(8)) gather_rows_kernel<int2><<<g, 256, 0, s>>>(sp, src_stride_bytes, idx, n, dp, dst_stride_bytes, row_bytes / 8);
    else if (ok(4)) gather_rows_kernel<int><<<g, 256, 0, s>>>(sp, src_stride_bytes, idx, n, dp, dst_stride_bytes, row_bytes / 4);
    else gather_rows_kernel<char><<<g, 256, 0, s>>>(sp, src_stride_bytes, idx, n, dp, dst_stride_bytes, row_bytes);
    count_launch();
    return last_error();
}

static int slice_launch(bool scatter, const float *src, int64_t ld_src, const int64_t *ridx,
                        int64_t n_rows, const int64_t *cidx, int64_t n_cols, float *dst,
                        int64_t ld_dst, cudaStream_t s) {
    if (n_rows < 0 || n_cols < 0) return GIST_ERR_BADARG;
    if (n_rows == 0 || n_cols == 0) return GIST_OK;
    if (!src || !dst) return GIST_ERR_BADARG;
    const int64_t gx = (n_cols + 255) / 256;
    if (gx > 0x7fffffffLL) return GIST_ERR_UNSUPPORTED;
    // enough row-CTAs to fill the chip a few times over, at most 65535
    int64_t gy = n_rows;
    const int64_t target = 8LL * kNumSMs;
    if (gx * gy > target * 4) gy = (target * 4 + gx - 1) / gx;
    if (gy > n_rows) gy = n_rows;
    if (gy > 65535) gy = 65535;
    if (gy < 1) gy = 1;
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (scatter) slice_kernel<true><<<grid, 256, 0, s>>>(src, ld_src, ridx, n_rows, cidx, n_cols, dst, ld_dst);
    else slice_kernel<false><<<grid, 256, 0, s>>>(src, ld_src, ridx, n_rows, cidx, n_cols, dst, ld_dst);
    count_launch();
    return last_error();
}

extern "C" int gist_slice_gather_f32(const float *src, int64_t ld_src, const int64_t *ridx,
                                     int64_t n_rows, const int64_t *cidx, int64_t n_cols, float *dst,
                                     int64_t ld_dst, gist_stream_t stream) {
    return slice_launch(false, src, ld_src, ridx, n_rows, cidx, n_cols, dst, ld_dst,
                        (cudaStream_t)stream);
}

extern "C" int gist_slice_scatter_f32(const float *src, int64_t ld_src, const int64_t *ridx,
                                      int64_t n_rows, const int64_t *cidx, int64_t n_cols,
                                      float *dst, int64_t ld_dst, gist_stream_t stream) {
    return slice_launch(true, src, ld_src, ridx, n_rows, cidx, n_cols, dst, ld_dst,
                        (cudaStream_t)stream);
}
