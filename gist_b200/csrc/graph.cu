// K3 (cluster-batch builder), exclusive scan, row gather and K5 (GIST slice
// gather / scatter).  All HBM/L2-bound integer and copy work: coalesced index
// streams, one warp per adjacency row, order-preserving ballot compaction.
#include <stdlib.h>

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

// ------------------------------------------------- batch builder, v2 ----
// v1 gives every batch row one warp (hub rows one CTA): the launch is as long as its unluckiest CTA
// — a parent row of 15 k edges is 15 rounds of two dependent loads for each of 8 warps, behind up to 8
// rounds of ordinary rows — so count and fill take 23-36 us each for ~1 M parent edges (4 MB) that the
// chip could walk in a few microseconds.  v2 cuts every parent row into CHUNKS of 128 edges (one round
// of 4 x 32 independent col -> node_map loads) and gives every chunk a warp of a chip-wide grid:
//   mark + chunks per row -> scan (chunk_ptr) -> count per chunk -> one-CTA scan of the chunk counts
//   (+ rowptr, 1 / degree) -> fill per chunk -> unmark.
// Chunks of a row are consecutive and a chunk's hits keep their order, so the CSR is identical to v1's
// (parent order preserved).  The walks stay latency-bound, but every warp's share is ONE round.
constexpr int kChunk = 32 * kBuildUnroll;

__global__ void batch_mark_chunks_kernel(const int64_t *__restrict__ nids, int32_t n_b,
                                         int32_t *__restrict__ node_map, const int32_t *__restrict__ prow,
                                         int32_t *__restrict__ nck) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_b) {
        const int64_t p = nids[i];
        int c = 0;
        if (p >= 0) {
            node_map[p] = i;
            c = (prow[p + 1] - prow[p] + kChunk - 1) / kChunk;
        }
        nck[i] = c;
    }
}

// chunk c belongs to the last row i with chunk_ptr[i] <= c (rows without chunks are skipped by construction);
// 32 probes per round: three rounds for 4 k rows instead of twelve dependent loads
__device__ __forceinline__ int chunk_row_search(const int32_t *__restrict__ chunk_ptr, int n_b, int c, int lane) {
    int lo = 0, hi = n_b;
    while (hi - lo > 1) {
        const int step = (hi - lo + 31) / 32;
        const int idx = lo + (lane + 1) * step;
        const bool le = idx < hi && __ldg(chunk_ptr + idx) <= c;
        const int k = __popc(__ballot_sync(0xffffffffu, le));
        lo += k * step;
        hi = min(hi, lo + step);
    }
    return lo;
}

__global__ void __launch_bounds__(256) chunk_count_kernel(const int32_t *__restrict__ prow,
                                                          const int32_t *__restrict__ pcol,
                                                          const int64_t *__restrict__ nids, int32_t n_b,
                                                          const int32_t *__restrict__ node_map,
                                                          const int32_t *__restrict__ chunk_ptr, int32_t max_chunks,
                                                          int32_t *__restrict__ chunk_row,
                                                          int32_t *__restrict__ chunk_cnt) {
    const int lane = threadIdx.x & 31;
    const int n_warps = gridDim.x * 8;
    const int T = min(__ldg(chunk_ptr + n_b), max_chunks);
    for (int c = blockIdx.x * 8 + (threadIdx.x >> 5); c < T; c += n_warps) {
        const int row = chunk_row_search(chunk_ptr, n_b, c, lane);
        const int64_t pnode = nids[row];
        const int rs = prow[pnode], re = prow[pnode + 1];
        const int eb = rs + (c - __ldg(chunk_ptr + row)) * kChunk;
        const int cnt = count_range(pcol, node_map, eb, min(re, eb + kChunk), lane);
        if (lane == 0) {
            chunk_row[c] = row;
            chunk_cnt[c] = cnt;
        }
    }
}

// One CTA: exclusive scan of the chunk counts in place (chunk_cnt[T] = total), then the batch's rowptr
// (row i starts where its first chunk does) and 1 / degree.
constexpr int kChunkScanThreads = 1024;
__global__ void __launch_bounds__(kChunkScanThreads) chunk_scan_rows_kernel(
    const int32_t *__restrict__ chunk_ptr, int32_t n_b, int32_t max_chunks, int32_t *__restrict__ chunk_cnt,
    int32_t *__restrict__ out_rowptr, float *__restrict__ out_inv_deg, int64_t col_capacity,
    int32_t *__restrict__ overflow_flag) {
    __shared__ int s_warp[kChunkScanThreads / 32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T_all = chunk_ptr[n_b];
    const int T = min(T_all, max_chunks);
    const int per = (T + kChunkScanThreads - 1) / kChunkScanThreads;
    const int b = min(T, tid * per), e = min(T, b + per);
    int sum = 0;
    for (int k = b; k < e; ++k) sum += chunk_cnt[k];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        s_warp[lane] = wi - w;          // exclusive offset of warp `lane`
        if (lane == 31) s_total = wi;
    }
    __syncthreads();
    int run = s_warp[warp] + incl - sum;
    for (int k = b; k < e; ++k) {
        const int v = chunk_cnt[k];
        chunk_cnt[k] = run;
        run += v;
    }
    const int total = s_total;
    if (tid == 0) {
        chunk_cnt[T] = total;
        if (overflow_flag && (total > col_capacity || T_all > max_chunks)) *overflow_flag = 1;
    }
    __syncthreads();                    // the positions (global memory, this CTA's own writes) are complete
    for (int i = tid; i <= n_b; i += kChunkScanThreads) out_rowptr[i] = chunk_cnt[min(chunk_ptr[i], T)];
    if (out_inv_deg) {
        __syncthreads();
        for (int i = tid; i < n_b; i += kChunkScanThreads) {
            const int dg = out_rowptr[i + 1] - out_rowptr[i];
            out_inv_deg[i] = dg > 0 ? 1.0f / (float)dg : 0.f;
        }
    }
}

__global__ void __launch_bounds__(256) chunk_fill_kernel(const int32_t *__restrict__ prow,
                                                         const int32_t *__restrict__ pcol,
                                                         const int64_t *__restrict__ nids, int32_t n_b,
                                                         const int32_t *__restrict__ node_map,
                                                         const int32_t *__restrict__ chunk_ptr, int32_t max_chunks,
                                                         const int32_t *__restrict__ chunk_row,
                                                         const int32_t *__restrict__ chunk_pos,
                                                         int32_t *__restrict__ out_col, int64_t col_capacity) {
    const int lane = threadIdx.x & 31;
    const int n_warps = gridDim.x * 8;
    const int T = min(__ldg(chunk_ptr + n_b), max_chunks);
    for (int c = blockIdx.x * 8 + (threadIdx.x >> 5); c < T; c += n_warps) {
        const int row = __ldg(chunk_row + c);
        const int64_t pnode = nids[row];
        const int rs = prow[pnode], re = prow[pnode + 1];
        const int eb = rs + (c - __ldg(chunk_ptr + row)) * kChunk;
        fill_range(pcol, node_map, eb, min(re, eb + kChunk), lane, (int64_t)__ldg(chunk_pos + c), out_col, col_capacity);
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

// Row-streaming merge (gist_slice_scatter_rows_f32): a CTA owns kRowsPerCta destination rows of one job and
// walks them left to right, one 32-byte sector (8 floats) per thread per step; the inverse column map of a
// sector is loaded once and serves all of the CTA's rows.
constexpr int kRowsL2 = 4, kRowsStaged = 2;
constexpr size_t kStageSmemMax = 72 * 1024;     // per CTA: three staged CTAs per SM
struct SliceRowsJobDev {
    const float *src;
    float *dst;
    const int64_t *ridx;
    const int32_t *inv;
    int64_t ld_src, ld_dst;
    int32_t n_rows, dst_cols;
    int32_t block0, n_cols;
};
struct SliceRowsJobsDev {
    SliceRowsJobDev j[kSliceMaxJobs];
    int32_t n_jobs;
};

// STAGED: the CTA's source rows are copied into shared memory first (coalesced) and patched in from there.
// Every source element is used once, at a random moment of the CTA's sweep over its destination rows, so
// without staging a source sector has to survive in L2 for the CTA's whole life next to the destination
// stream — at the ultra-wide size the resident CTAs' source rows alone are ~110 MB.
template <int ROWS, bool STAGED>
__global__ void __launch_bounds__(256) slice_scatter_rows_kernel(const __grid_constant__ SliceRowsJobsDev a) {
    extern __shared__ float s_src[];        // STAGED: [ROWS][n_cols]
    int lo = 0, hi = a.n_jobs - 1;          // last job whose block0 <= blockIdx.x
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int)blockIdx.x >= a.j[mid].block0) lo = mid;
        else hi = mid - 1;
    }
    const SliceRowsJobDev &J = a.j[lo];
    constexpr int kRowsPerCta = ROWS;
    const int r0 = ((int)blockIdx.x - J.block0) * kRowsPerCta;
    const float *srow[kRowsPerCta];
    float *drow[kRowsPerCta];
#pragma unroll
    for (int k = 0; k < kRowsPerCta; ++k) {
        const int r = min(r0 + k, J.n_rows - 1);                // rows past the end repeat the last one, never stored
        const int64_t rr = J.ridx ? __ldg(J.ridx + r) : (int64_t)r;
        srow[k] = J.src + (int64_t)r * J.ld_src;
        drow[k] = J.dst + rr * J.ld_dst;
    }
    const int n_live = min(kRowsPerCta, J.n_rows - r0);
    if constexpr (STAGED) {
#pragma unroll
        for (int k = 0; k < kRowsPerCta; ++k) {
            float *sk = s_src + (size_t)k * J.n_cols;
            if ((J.n_cols & 3) == 0 && (reinterpret_cast<uintptr_t>(srow[k]) & 15) == 0) {
                for (int c = threadIdx.x * 4; c < J.n_cols; c += 256 * 4)
                    *reinterpret_cast<float4 *>(sk + c) = *reinterpret_cast<const float4 *>(srow[k] + c);
            } else {
                for (int c = threadIdx.x; c < J.n_cols; c += 256) sk[c] = srow[k][c];
            }
            srow[k] = sk;
        }
        __syncthreads();
    }
    for (int c0 = threadIdx.x * 8; c0 < J.dst_cols; c0 += 256 * 8) {
        int iv[8];
        if (c0 + 8 <= J.dst_cols) {
            const int4 i0 = __ldg(reinterpret_cast<const int4 *>(J.inv + c0));
            const int4 i1 = __ldg(reinterpret_cast<const int4 *>(J.inv + c0 + 4));
            iv[0] = i0.x; iv[1] = i0.y; iv[2] = i0.z; iv[3] = i0.w;
            iv[4] = i1.x; iv[5] = i1.y; iv[6] = i1.z; iv[7] = i1.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) iv[i] = (c0 + i < J.dst_cols) ? __ldg(J.inv + c0 + i) : -1;
        }
        int any = -1;
#pragma unroll
        for (int i = 0; i < 8; ++i) any = max(any, iv[i]);
        if (any < 0) continue;                                  // nothing lands on this sector
        const bool whole = c0 + 8 <= J.dst_cols;
#pragma unroll
        for (int k = 0; k < kRowsPerCta; ++k) {
            if (k >= n_live) break;
            float *d = drow[k] + c0;
            if (whole) {
                float4 v0 = *reinterpret_cast<const float4 *>(d);
                float4 v1 = *reinterpret_cast<const float4 *>(d + 4);
                float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (iv[i] >= 0) v[i] = srow[k][iv[i]];      // plain load: src may be peer memory
                *reinterpret_cast<float4 *>(d) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4 *>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (iv[i] >= 0) d[i] = srow[k][iv[i]];
            }
        }
    }
}

__global__ void index_invert_kernel(const int64_t *__restrict__ idx, int64_t n, int32_t *__restrict__ inv, int64_t size) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const int64_t c = idx[k];
        if (c >= 0 && c < size) inv[c] = (int32_t)k;
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

extern "C" int gist_slice_scatter_rows_f32(int32_t n_jobs, const gist_slice_rows_job_t *jobs, gist_stream_t stream) {
    if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return GIST_ERR_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    int done = 0;
    while (done < n_jobs) {
        SliceRowsJobsDev a;
        a.n_jobs = 0;
        int64_t blocks = 0, max_cols = 0;
        for (; done < n_jobs && a.n_jobs < kSliceMaxJobs; ++done) {
            const gist_slice_rows_job_t &j = jobs[done];
            if (j.n_rows < 0 || j.dst_cols < 0 || j.n_cols < 0) return GIST_ERR_BADARG;
            if (j.n_rows == 0 || j.dst_cols == 0 || j.n_cols == 0) continue;
            if (!j.src || !j.dst || !j.inv_col || j.n_rows > 0x7fffffffLL || j.dst_cols > 0x7fffffffLL ||
                j.n_cols > 0x7fffffffLL || j.ld_dst < j.dst_cols || j.ld_src < j.n_cols)
                return GIST_ERR_BADARG;
            if (!aligned(j.dst, 32) || j.ld_dst % 8 || !aligned(j.inv_col, 16)) return GIST_ERR_ALIGN;
            SliceRowsJobDev &d = a.j[a.n_jobs++];
            d.src = j.src; d.dst = j.dst; d.ridx = j.ridx; d.inv = j.inv_col;
            d.ld_src = j.ld_src; d.ld_dst = j.ld_dst;
            d.n_rows = (int32_t)j.n_rows; d.dst_cols = (int32_t)j.dst_cols;
            d.n_cols = (int32_t)j.n_cols;
            if (j.n_cols > max_cols) max_cols = j.n_cols;
        }
        if (a.n_jobs == 0) continue;
        // source rows staged in shared memory when two of them fit a third of an SM's shared memory
        const size_t smem = (size_t)kRowsStaged * (size_t)max_cols * sizeof(float);
        const bool staged = smem <= kStageSmemMax && !getenv("GIST_MERGE_NO_STAGE");
        const int rows_per_cta = staged ? kRowsStaged : kRowsL2;
        for (int k = 0; k < a.n_jobs; ++k) {
            a.j[k].block0 = (int32_t)blocks;
            blocks += (a.j[k].n_rows + rows_per_cta - 1) / rows_per_cta;
            if (blocks > 0x7fffffffLL) return GIST_ERR_UNSUPPORTED;
        }
        if (staged) {
            static bool configured = false;
            if (!configured) {
                cudaError_t e = cudaFuncSetAttribute(slice_scatter_rows_kernel<kRowsStaged, true>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStageSmemMax);
                if (e != cudaSuccess) return (int)e;
                configured = true;
            }
            slice_scatter_rows_kernel<kRowsStaged, true><<<(unsigned)blocks, 256, smem, s>>>(a);
        } else {
            slice_scatter_rows_kernel<kRowsL2, false><<<(unsigned)blocks, 256, 0, s>>>(a);
        }
        count_launch();
        const int st = last_error();
        if (st != GIST_OK) return st;
    }
    return GIST_OK;
}

extern "C" int gist_index_invert_i32(const int64_t *idx, int64_t n, int32_t *inv, int64_t size, gist_stream_t stream) {
    if (n < 0 || size < 0 || (size > 0 && !inv) || (n > 0 && !idx) || size > 0x7fffffffLL) return GIST_ERR_BADARG;
    if (size == 0) return GIST_OK;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(inv, 0xFF, (size_t)size * sizeof(int32_t), s);
    if (e != cudaSuccess) return (int)e;
    if (n > 0) {
        index_invert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(idx, n, inv, size);
        count_launch();
    }
    return last_error();
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

static size_t align16(size_t x) { return (x + 15) / 16 * 16; }

extern "C" size_t gist_cluster_batch_build_v2_workspace_bytes(int32_t n_b, int64_t max_chunks) {
    if (n_b < 0 || max_chunks < 0) return 0;
    return align16(((size_t)n_b + 1) * 4) + align16((size_t)max_chunks * 4) + align16(((size_t)max_chunks + 1) * 4) +
           align16(gist_scan_workspace_bytes(n_b));
}

extern "C" int gist_cluster_batch_build_v2(const int32_t *parent_rowptr, const int32_t *parent_col,
                                           int32_t n_parent, const int64_t *nids, int32_t n_b,
                                           int32_t *node_map, int32_t *out_rowptr, int32_t *out_col,
                                           int64_t col_capacity, float *out_inv_deg, int32_t *overflow_flag,
                                           int64_t max_chunks, void *workspace, size_t workspace_bytes,
                                           gist_stream_t stream) {
    if (n_parent < 0 || n_b < 0 || col_capacity < 0 || max_chunks < 0 || max_chunks > 0x3fffffffLL) return GIST_ERR_BADARG;
    if (!out_rowptr) return GIST_ERR_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (n_b == 0) {
        cudaError_t e = cudaMemsetAsync(out_rowptr, 0, sizeof(int32_t), s);
        return e == cudaSuccess ? GIST_OK : (int)e;
    }
    if (!parent_rowptr || !nids || !node_map) return GIST_ERR_BADARG;
    if (col_capacity > 0 && !out_col) return GIST_ERR_BADARG;
    if (!workspace || !aligned(workspace, 16) ||
        workspace_bytes < gist_cluster_batch_build_v2_workspace_bytes(n_b, max_chunks))
        return GIST_ERR_WORKSPACE;
    char *w = reinterpret_cast<char *>(workspace);
    int32_t *chunk_ptr = reinterpret_cast<int32_t *>(w);
    w += align16(((size_t)n_b + 1) * 4);
    int32_t *chunk_row = reinterpret_cast<int32_t *>(w);
    w += align16((size_t)max_chunks * 4);
    int32_t *chunk_cnt = reinterpret_cast<int32_t *>(w);
    w += align16(((size_t)max_chunks + 1) * 4);
    void *scan_ws = w;
    const size_t scan_ws_bytes = gist_scan_workspace_bytes(n_b);
    const int tb = 256;
    const int g_thread = (n_b + tb - 1) / tb;
    int64_t g_chunk = (max_chunks + 7) / 8;
    if (g_chunk > 8LL * kNumSMs) g_chunk = 8LL * kNumSMs;
    if (g_chunk < 1) g_chunk = 1;
    batch_mark_chunks_kernel<<<g_thread, tb, 0, s>>>(nids, n_b, node_map, parent_rowptr, chunk_ptr);
    count_launch();
    int st = scan_launch(chunk_ptr, n_b, chunk_ptr, scan_ws, scan_ws_bytes, s);
    if (st != GIST_OK) return st;
    chunk_count_kernel<<<(unsigned)g_chunk, tb, 0, s>>>(parent_rowptr, parent_col, nids, n_b, node_map, chunk_ptr,
                                                       (int32_t)max_chunks, chunk_row, chunk_cnt);
    chunk_scan_rows_kernel<<<1, kChunkScanThreads, 0, s>>>(chunk_ptr, n_b, (int32_t)max_chunks, chunk_cnt, out_rowptr,
                                                           out_inv_deg, col_capacity, overflow_flag);
    chunk_fill_kernel<<<(unsigned)g_chunk, tb, 0, s>>>(parent_rowptr, parent_col, nids, n_b, node_map, chunk_ptr,
                                                      (int32_t)max_chunks, chunk_row, chunk_cnt, out_col, col_capacity);
    batch_mark_kernel<<<g_thread, tb, 0, s>>>(nids, n_b, node_map, 0);
    count_launch(4);
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
