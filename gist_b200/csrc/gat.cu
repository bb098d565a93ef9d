// K6: graph attention (GAT) message passing — fused edge-softmax + weighted SpMM, sm_100a.
//
// Replaces GATLayer.forward's apply_edges(edge_attention) + update_all(message_func,
// reduce_func) (cluster_gcn/modules.py:10-65): per-edge Python UDFs over degree-bucketed
// mailboxes in DGL.  For one head, with z = fc(h) [n, D] and attn_fc.weight = [a_l ‖ a_r]:
//
//   el[u] = a_l·z[u],  er[v] = a_r·z[v]                                     (scores kernel)
//   raw_uv = el[u] + er[v];  e_uv = leaky_relu(raw_uv, slope)               (modules.py:40-44)
//   alpha_uv = softmax over the in-edges (u -> v) of v                      (modules.py:53)
//   out[v] = sum_u alpha_uv z[u]                                            (modules.py:55)
//
// No per-edge tensor is ever materialised: a score needs only two per-node scalars, so the
// forward keeps lse[v] = logsumexp_u e_uv and the backward recomputes alpha_uv on the fly.
// Backward (all fixed-order, no atomics), with g = dOut:
//   c[v]   = g[v]·out[v]
//   ds_uv  = alpha_uv (g[v]·z[u] - c[v]) leaky'(raw_uv)
//   der[v] = sum_u ds_uv                      pass A: CSR order (rows v), gathers z[u]
//   del[u] = sum_v ds_uv                      pass B: CSC order (rows u), gathers g[v]
//   dz[u]  = sum_v alpha_uv g[v] + del[u] a_l + der[u] a_r          (pass B epilogue)
//   da_l   = sum_u del[u] z[u],  da_r = sum_u der[u] z[u]            (weighted column sums)
// One warp per row; a lane holds D/32 features in registers (D <= 1024).
#include "common.cuh"

namespace gist {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int VEC>
__device__ __forceinline__ void ldv(float (&r)[VEC], const float *p) {
    if constexpr (VEC == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
        r[0] = __ldg(p);
    }
}

template <int VEC>
__device__ __forceinline__ void stv(float *p, const float (&r)[VEC]) {
    if constexpr (VEC == 4) *reinterpret_cast<float4 *>(p) = make_float4(r[0], r[1], r[2], r[3]);
    else *p = r[0];
}

__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : slope * x; }

// Row `row` of a [*, D] matrix into the lane's register slots (zeros beyond D).
template <int VEC, int SLOTS>
__device__ __forceinline__ void load_row(float (&r)[SLOTS][VEC], const float *base, int D, int lane) {
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
        const int c = (k * 32 + lane) * VEC;
        if (c < D) {
            ldv<VEC>(r[k], base + c);
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) r[k][i] = 0.f;
        }
    }
}

template <int VEC, int SLOTS>
__device__ __forceinline__ float dot_slots(const float (&a)[SLOTS][VEC], const float (&b)[SLOTS][VEC]) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k)
#pragma unroll
        for (int i = 0; i < VEC; ++i) s = fmaf(a[k][i], b[k][i], s);
    return s;
}

// s[r] = (a_l·z[r], a_r·z[r])
__global__ void __launch_bounds__(256) gat_scores_kernel(const float *__restrict__ z, int64_t ldz, int n, int D,
                                                         const float *__restrict__ attn,
                                                         float2 *__restrict__ s) {
    // blockIdx.y = attention head: head h owns columns [h D, (h + 1) D) of z and its own attn / score rows
    z += (int64_t)blockIdx.y * D;
    attn += (int64_t)blockIdx.y * 2 * D;
    s += (int64_t)blockIdx.y * n;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const float *zr = z + (int64_t)r * ldz;
    float a = 0.f, b = 0.f;
    for (int c = lane; c < D; c += 32) {
        const float x = __ldg(zr + c);
        a = fmaf(x, __ldg(attn + c), a);
        b = fmaf(x, __ldg(attn + D + c), b);
    }
    a = wsum(a);
    b = wsum(b);
    if (lane == 0) s[r] = make_float2(a, b);
}

// Forward: out[v] = sum_u softmax_u(leaky(el[u] + er[v])) z[u];  lse[v] kept for the backward.
template <int VEC, int SLOTS>
__global__ void __launch_bounds__(256) gat_fwd_kernel(const int32_t *__restrict__ rowptr,
                                                      const int32_t *__restrict__ col, int n,
                                                      const float *__restrict__ z, int64_t ldz, int D,
                                                      const float2 *__restrict__ s, float slope,
                                                      float *__restrict__ out, int64_t ldo,
                                                      float *__restrict__ lse) {
    z += (int64_t)blockIdx.y * D;        // all heads of a layer in one launch: head = blockIdx.y
    s += (int64_t)blockIdx.y * n;
    out += (int64_t)blockIdx.y * D;
    lse += (int64_t)blockIdx.y * n;
    const int v = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (v >= n) return;
    const int rs = __ldg(rowptr + v), re = __ldg(rowptr + v + 1);
    const float er = s[v].y;
    float m = -INFINITY;
    for (int e = rs + lane; e < re; e += 32) m = fmaxf(m, leaky(s[__ldg(col + e)].x + er, slope));
    m = wmax(m);
    float den = 0.f;
    for (int e = rs + lane; e < re; e += 32) den += __expf(leaky(s[__ldg(col + e)].x + er, slope) - m);
    den = wsum(den);
    const float L = re > rs ? m + __logf(den) : 0.f;
    float acc[SLOTS][VEC];
#pragma unroll
    for (int k = 0; k < SLOTS; ++k)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[k][i] = 0.f;
    constexpr int U = SLOTS >= 4 ? 2 : 4;
    for (int e0 = rs; e0 < re; e0 += 32) {
        const int my = e0 + lane;
        int u_m = 0;
        float w_m = 0.f;
        if (my < re) {
            u_m = __ldg(col + my);
            w_m = __expf(leaky(s[u_m].x + er, slope) - L);
        }
        const int cnt = min(32, re - e0);
        for (int j = 0; j < cnt; j += U) {
            float x[U][SLOTS][VEC];
            float w[U];
#pragma unroll
            for (int jj = 0; jj < U; ++jj) {
                const int u = __shfl_sync(0xffffffffu, u_m, (j + jj) & 31);
                w[jj] = (j + jj) < cnt ? __shfl_sync(0xffffffffu, w_m, (j + jj) & 31) : 0.f;
                load_row<VEC, SLOTS>(x[jj], z + (int64_t)u * ldz, D, lane);
            }
#pragma unroll
            for (int jj = 0; jj < U; ++jj)
#pragma unroll
                for (int k = 0; k < SLOTS; ++k)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[k][i] = fmaf(w[jj], x[jj][k][i], acc[k][i]);
        }
    }
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
        const int c = (k * 32 + lane) * VEC;
        if (c < D) stv<VEC>(out + (int64_t)v * ldo + c, acc[k]);
    }
    if (lane == 0) lse[v] = L;
}

// Backward pass A (CSR order): c[v] = g[v]·out[v];  der[v] = sum_u alpha_uv (g[v]·z[u] - c[v]) leaky'(raw_uv)
template <int VEC, int SLOTS>
__global__ void __launch_bounds__(256) gat_bwd_dst_kernel(const int32_t *__restrict__ rowptr,
                                                          const int32_t *__restrict__ col, int n,
                                                          const float *__restrict__ z, int64_t ldz, int D,
                                                          const float2 *__restrict__ s,
                                                          const float *__restrict__ lse,
                                                          const float *__restrict__ g, int64_t ldg,
                                                          const float *__restrict__ out, int64_t ldo,
                                                          float slope, float *__restrict__ cvec,
                                                          float *__restrict__ der) {
    z += (int64_t)blockIdx.y * D;
    g += (int64_t)blockIdx.y * D;
    out += (int64_t)blockIdx.y * D;
    s += (int64_t)blockIdx.y * n;
    lse += (int64_t)blockIdx.y * n;
    cvec += (int64_t)blockIdx.y * n;
    der += (int64_t)blockIdx.y * n;
    const int v = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (v >= n) return;
    const int rs = __ldg(rowptr + v), re = __ldg(rowptr + v + 1);
    float gv[SLOTS][VEC], ov[SLOTS][VEC];
    load_row<VEC, SLOTS>(gv, g + (int64_t)v * ldg, D, lane);
    load_row<VEC, SLOTS>(ov, out + (int64_t)v * ldo, D, lane);
    const float c = wsum(dot_slots<VEC, SLOTS>(gv, ov));
    const float er = s[v].y, L = lse[v];
    float acc = 0.f;       // lane j accumulates the ds of the edges it owns
    constexpr int U = SLOTS >= 4 ? 2 : 4;
    for (int e0 = rs; e0 < re; e0 += 32) {
        const int my = e0 + lane;
        int u_m = 0;
        float f_m = 0.f;   // alpha_uv * leaky'(raw_uv) of my edge
        if (my < re) {
            u_m = __ldg(col + my);
            const float raw = s[u_m].x + er;
            f_m = __expf(leaky(raw, slope) - L) * (raw > 0.f ? 1.f : slope);
        }
        const int cnt = min(32, re - e0);
        float t_m = 0.f;   // g[v]·z[u] of my edge
        for (int j = 0; j < cnt; j += U) {
            float x[U][SLOTS][VEC];
#pragma unroll
            for (int jj = 0; jj < U; ++jj) {
                const int u = __shfl_sync(0xffffffffu, u_m, (j + jj) & 31);
                load_row<VEC, SLOTS>(x[jj], z + (int64_t)u * ldz, D, lane);
            }
#pragma unroll
            for (int jj = 0; jj < U; ++jj) {
                const float t = wsum(dot_slots<VEC, SLOTS>(gv, x[jj]));
                if (lane == j + jj) t_m = t;
            }
        }
        acc += f_m * (t_m - c);      // f_m = 0 for lanes without an edge
    }
    acc = wsum(acc);
    if (lane == 0) {
        cvec[v] = c;
        der[v] = acc;
    }
}

// Backward pass B (CSC order, rows u over out-edges u -> v):
//   del[u] = sum_v alpha_uv (g[v]·z[u] - c[v]) leaky'(raw_uv)
//   dz[u]  = sum_v alpha_uv g[v] + del[u] a_l + der[u] a_r
template <int VEC, int SLOTS>
__global__ void __launch_bounds__(256) gat_bwd_src_kernel(const int32_t *__restrict__ colptr,
                                                          const int32_t *__restrict__ row, int n,
                                                          const float *__restrict__ z, int64_t ldz, int D,
                                                          const float2 *__restrict__ s,
                                                          const float *__restrict__ lse,
                                                          const float *__restrict__ cvec,
                                                          const float *__restrict__ der,
                                                          const float *__restrict__ g, int64_t ldg,
                                                          const float *__restrict__ attn, float slope,
                                                          float *__restrict__ dz, int64_t lddz,
                                                          float *__restrict__ del) {
    z += (int64_t)blockIdx.y * D;
    g += (int64_t)blockIdx.y * D;
    dz += (int64_t)blockIdx.y * D;
    attn += (int64_t)blockIdx.y * 2 * D;
    s += (int64_t)blockIdx.y * n;
    lse += (int64_t)blockIdx.y * n;
    cvec += (int64_t)blockIdx.y * n;
    der += (int64_t)blockIdx.y * n;
    del += (int64_t)blockIdx.y * n;
    const int u = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (u >= n) return;
    const int rs = __ldg(colptr + u), re = __ldg(colptr + u + 1);
    float zu[SLOTS][VEC], acc[SLOTS][VEC];
    load_row<VEC, SLOTS>(zu, z + (int64_t)u * ldz, D, lane);
#pragma unroll
    for (int k = 0; k < SLOTS; ++k)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[k][i] = 0.f;
    const float el = s[u].x;
    float dacc = 0.f;
    constexpr int U = SLOTS >= 4 ? 2 : 4;
    for (int e0 = rs; e0 < re; e0 += 32) {
        const int my = e0 + lane;
        int v_m = 0;
        float a_m = 0.f, f_m = 0.f, c_m = 0.f;
        if (my < re) {
            v_m = __ldg(row + my);
            const float raw = el + s[v_m].y;
            a_m = __expf(leaky(raw, slope) - lse[v_m]);
            f_m = a_m * (raw > 0.f ? 1.f : slope);
            c_m = cvec[v_m];
        }
        const int cnt = min(32, re - e0);
        float t_m = 0.f;
        for (int j = 0; j < cnt; j += U) {
            float x[U][SLOTS][VEC];
            float a[U];
#pragma unroll
            for (int jj = 0; jj < U; ++jj) {
                const int v = __shfl_sync(0xffffffffu, v_m, (j + jj) & 31);
                a[jj] = (j + jj) < cnt ? __shfl_sync(0xffffffffu, a_m, (j + jj) & 31) : 0.f;
                load_row<VEC, SLOTS>(x[jj], g + (int64_t)v * ldg, D, lane);
            }
#pragma unroll
            for (int jj = 0; jj < U; ++jj) {
                const float t = wsum(dot_slots<VEC, SLOTS>(zu, x[jj]));
                if (lane == j + jj) t_m = t;
#pragma unroll
                for (int k = 0; k < SLOTS; ++k)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[k][i] = fmaf(a[jj], x[jj][k][i], acc[k][i]);
            }
        }
        dacc += f_m * (t_m - c_m);
    }
    dacc = wsum(dacc);
    const float dr = der[u];
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
        const int c = (k * 32 + lane) * VEC;
        if (c < D) {
            float al[VEC], ar[VEC], o[VEC];
            ldv<VEC>(al, attn + c);
            ldv<VEC>(ar, attn + D + c);
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] = acc[k][i] + dacc * al[i] + dr * ar[i];
            stv<VEC>(dz + (int64_t)u * lddz + c, o);
        }
    }
    if (lane == 0) del[u] = dacc;
}

// dattn partials: part[by][0:D] = sum_r del[r] z[r,:], part[by][D:2D] = sum_r der[r] z[r,:]
__global__ void __launch_bounds__(256) gat_dattn_partial_kernel(const float *__restrict__ z, int64_t ldz, int n,
                                                                int D, const float *__restrict__ del,
                                                                const float *__restrict__ der, int rows_per,
                                                                float *__restrict__ part) {
    __shared__ float sl[8][33], sr[8][33];
    z += (int64_t)blockIdx.z * D;            // head = blockIdx.z
    del += (int64_t)blockIdx.z * n;
    der += (int64_t)blockIdx.z * n;
    part += (int64_t)blockIdx.z * gridDim.y * 2 * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const int r0 = blockIdx.y * rows_per;
    const int r1 = min(n, r0 + rows_per);
    float a = 0.f, b = 0.f;
    if (c < D) {
        for (int r = r0 + warp; r < r1; r += 8) {
            const float x = __ldg(z + (int64_t)r * ldz + c);
            a = fmaf(__ldg(del + r), x, a);
            b = fmaf(__ldg(der + r), x, b);
        }
    }
    sl[warp][lane] = a;
    sr[warp][lane] = b;
    __syncthreads();
    if (warp == 0 && c < D) {
        float ta = 0.f, tb = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { ta += sl[w][lane]; tb += sr[w][lane]; }
        part[(int64_t)blockIdx.y * 2 * D + c] = ta;
        part[(int64_t)blockIdx.y * 2 * D + D + c] = tb;
    }
}

__global__ void __launch_bounds__(256) gat_dattn_final_kernel(const float *__restrict__ part, int nparts, int D2,
                                                              float *__restrict__ out) {
    part += (int64_t)blockIdx.y * nparts * D2;   // head = blockIdx.y
    out += (int64_t)blockIdx.y * D2;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D2) return;
    float t = 0.f;
#pragma unroll 8
    for (int p = 0; p < nparts; ++p) t += __ldg(part + (int64_t)p * D2 + c);
    out[c] = t;
}

static int dattn_parts(int n, int D) {
    const int cx = (D + 31) / 32;
    int by = (2 * kNumSMs + cx - 1) / cx;
    const int max_by = (n + 63) / 64;
    if (by > max_by) by = max_by;
    return by < 1 ? 1 : by;
}

struct GatVec {
    int vec, slots;
};

// widest vector the pointers allow, and the register slots a lane needs for D features
static bool gat_pick(int D, bool v4, GatVec *o) {
    const int per_slot = 32 * (v4 ? 4 : 1);
    int slots = (D + per_slot - 1) / per_slot;
    if (v4) {
        if (slots <= 1) slots = 1; else if (slots <= 2) slots = 2; else if (slots <= 4) slots = 4;
        else if (slots <= 8) slots = 8; else return false;
    } else {
        if (slots <= 2) slots = 2; else if (slots <= 8) slots = 8; else if (slots <= 32) slots = 32;
        else return false;
    }
    o->vec = v4 ? 4 : 1;
    o->slots = slots;
    return true;
}

#define GAT_DISPATCH(KERNEL, gv, ...)                                                      \
    do {                                                                                   \
        if (gv.vec == 4) {                                                                 \
            if (gv.slots == 1) KERNEL<4, 1><<<grid, 256, 0, s>>>(__VA_ARGS__);             \
            else if (gv.slots == 2) KERNEL<4, 2><<<grid, 256, 0, s>>>(__VA_ARGS__);        \
            else if (gv.slots == 4) KERNEL<4, 4><<<grid, 256, 0, s>>>(__VA_ARGS__);        \
            else KERNEL<4, 8><<<grid, 256, 0, s>>>(__VA_ARGS__);                           \
        } else {                                                                           \
            if (gv.slots == 2) KERNEL<1, 2><<<grid, 256, 0, s>>>(__VA_ARGS__);             \
            else if (gv.slots == 8) KERNEL<1, 8><<<grid, 256, 0, s>>>(__VA_ARGS__);        \
            else KERNEL<1, 32><<<grid, 256, 0, s>>>(__VA_ARGS__);                          \
        }                                                                                  \
    } while (0)

}  // namespace gist

using namespace gist;

// All H heads of a MultiHeadGATLayer in ONE launch per kernel (cluster_gcn/modules.py:67-76 loops over the
// heads): head h owns columns [h D, (h + 1) D) of z / out / dout / dz (one [n, H D] buffer each, the
// projection of all heads being one GEMM) and row h of attn [H, 2D], scores [H, n, 2], lse [H, n],
// dattn [H, 2D].  The grid's y dimension is the head.  H = 1 is the single-head API below.
extern "C" int gist_gat_scores_heads_f32(const float *z, int64_t ldz, int32_t n, int32_t D, int32_t heads,
                                         const float *attn, float *scores, gist_stream_t stream) {
    if (n < 0 || D <= 0 || heads <= 0 || heads > 65535) return GIST_ERR_BADARG;
    if (n == 0) return GIST_OK;
    if (!z || !attn || !scores || ldz < (int64_t)D * heads) return GIST_ERR_BADARG;
    if (!aligned(scores, 8)) return GIST_ERR_ALIGN;
    dim3 grid((unsigned)((n + 7) / 8), (unsigned)heads);
    gat_scores_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, ldz, n, D, attn, reinterpret_cast<float2 *>(scores));
    count_launch();
    return last_error();
}

extern "C" int gist_gat_scores_f32(const float *z, int64_t ldz, int32_t n, int32_t D, const float *attn,
                                   float *scores, gist_stream_t stream) {
    return gist_gat_scores_heads_f32(z, ldz, n, D, 1, attn, scores, stream);
}

extern "C" int gist_gat_aggregate_heads_f32(const int32_t *rowptr, const int32_t *col, int32_t n, const float *z,
                                            int64_t ldz, int32_t D, int32_t heads, const float *scores,
                                            float negative_slope, float *out, int64_t ldo, float *lse,
                                            gist_stream_t stream) {
    if (n < 0 || D <= 0 || heads <= 0 || heads > 65535) return GIST_ERR_BADARG;
    if (n == 0) return GIST_OK;
    if (!rowptr || !z || !scores || !out || !lse || ldz < (int64_t)D * heads || ldo < (int64_t)D * heads)
        return GIST_ERR_BADARG;
    if (!aligned(scores, 8)) return GIST_ERR_ALIGN;
    const bool v4 = D % 4 == 0 && ldz % 4 == 0 && ldo % 4 == 0 && aligned(z, 16) && aligned(out, 16);
    GatVec gv;
    if (!gat_pick(D, v4, &gv)) return GIST_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((unsigned)((n + 7) / 8), (unsigned)heads);
    const float2 *sc = reinterpret_cast<const float2 *>(scores);
    GAT_DISPATCH(gat_fwd_kernel, gv, rowptr, col, n, z, ldz, D, sc, negative_slope, out, ldo, lse);
    count_launch();
    return last_error();
}

extern "C" int gist_gat_aggregate_f32(const int32_t *rowptr, const int32_t *col, int32_t n, const float *z,
                                      int64_t ldz, int32_t D, const float *scores, float negative_slope,
                                      float *out, int64_t ldo, float *lse, gist_stream_t stream) {
    return gist_gat_aggregate_heads_f32(rowptr, col, n, z, ldz, D, 1, scores, negative_slope, out, ldo, lse, stream);
}

extern "C" size_t gist_gat_backward_heads_workspace_bytes(int32_t n, int32_t D, int32_t heads) {
    if (n <= 0 || D <= 0 || heads <= 0) return 0;
    // per head: c[n], der[n], del[n], dattn partials [parts][2D]
    return (size_t)heads * ((size_t)3 * n + (size_t)dattn_parts(n, D) * 2 * D) * sizeof(float);
}

extern "C" size_t gist_gat_backward_workspace_bytes(int32_t n, int32_t D) {
    return gist_gat_backward_heads_workspace_bytes(n, D, 1);
}

extern "C" int gist_gat_backward_heads_f32(const int32_t *rowptr, const int32_t *col, const int32_t *colptr,
                                           const int32_t *row, int32_t n, const float *z, int64_t ldz, int32_t D,
                                           int32_t heads, const float *scores, const float *lse, const float *attn,
                                           float negative_slope, const float *out, int64_t ldo, const float *dout,
                                           int64_t lddo, float *dz, int64_t lddz, float *dattn, void *workspace,
                                           size_t workspace_bytes, gist_stream_t stream) {
    if (n < 0 || D <= 0 || heads <= 0 || heads > 65535) return GIST_ERR_BADARG;
    if (!dattn) return GIST_ERR_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(dattn, 0, (size_t)heads * 2 * D * sizeof(float), s);
        return e == cudaSuccess ? GIST_OK : (int)e;
    }
    if (!rowptr || !colptr || !z || !scores || !lse || !attn || !out || !dout || !dz) return GIST_ERR_BADARG;
    const int64_t W = (int64_t)D * heads;
    if (ldz < W || ldo < W || lddo < W || lddz < W) return GIST_ERR_BADARG;
    if (!workspace || workspace_bytes < gist_gat_backward_heads_workspace_bytes(n, D, heads)) return GIST_ERR_WORKSPACE;
    if (!aligned(scores, 8) || !aligned(workspace, 4)) return GIST_ERR_ALIGN;
    const bool v4 = D % 4 == 0 && ldz % 4 == 0 && ldo % 4 == 0 && lddo % 4 == 0 && lddz % 4 == 0 &&
                    aligned(z, 16) && aligned(out, 16) && aligned(dout, 16) && aligned(dz, 16) &&
                    aligned(attn, 16);
    GatVec gv;
    if (!gat_pick(D, v4, &gv)) return GIST_ERR_UNSUPPORTED;
    float *cvec = reinterpret_cast<float *>(workspace);
    float *der = cvec + (size_t)heads * n, *del = der + (size_t)heads * n, *part = del + (size_t)heads * n;
    dim3 grid((unsigned)((n + 7) / 8), (unsigned)heads);
    const float2 *sc = reinterpret_cast<const float2 *>(scores);
    GAT_DISPATCH(gat_bwd_dst_kernel, gv, rowptr, col, n, z, ldz, D, sc, lse, dout, lddo, out, ldo,
                 negative_slope, cvec, der);
    GAT_DISPATCH(gat_bwd_src_kernel, gv, colptr, row, n, z, ldz, D, sc, lse, cvec, der, dout, lddo, attn,
                 negative_slope, dz, lddz, del);
    const int by = dattn_parts(n, D);
    const int rows_per = (n + by - 1) / by;
    dim3 g2((D + 31) / 32, by, (unsigned)heads);
    gat_dattn_partial_kernel<<<g2, 256, 0, s>>>(z, ldz, n, D, del, der, rows_per, part);
    dim3 g3((2 * D + 255) / 256, (unsigned)heads);
    gat_dattn_final_kernel<<<g3, 256, 0, s>>>(part, by, 2 * D, dattn);
    count_launch(4);
    return last_error();
}

extern "C" int gist_gat_backward_f32(const int32_t *rowptr, const int32_t *col, const int32_t *colptr,
                                     const int32_t *row, int32_t n, const float *z, int64_t ldz, int32_t D,
                                     const float *scores, const float *lse, const float *attn,
                                     float negative_slope, const float *out, int64_t ldo, const float *dout,
                                     int64_t lddo, float *dz, int64_t lddz, float *dattn, void *workspace,
                                     size_t workspace_bytes, gist_stream_t stream) {
    return gist_gat_backward_heads_f32(rowptr, col, colptr, row, n, z, ldz, D, 1, scores, lse, attn, negative_slope,
                                       out, ldo, dout, lddo, dz, lddz, dattn, workspace, workspace_bytes, stream);
}
