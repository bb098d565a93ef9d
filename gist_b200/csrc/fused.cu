// Row-wise / elementwise pieces of the training step that sit between the SpMM and GEMM
// kernels (cluster_gcn/modules.py:233-236, cluster_gcn_ist_distrib.py:405-417), fused so a
// cluster-batch step is a few launches instead of dozens:
//   * LayerNorm(affine=False) + ReLU, forward and backward          (modules.py:234-236)
//   * column sums (bias gradient of nn.Linear)
//   * masked cross entropy, forward (mean over masked rows) + backward (…distrib.py:413-415)
//   * Adam over a list of tensors in ONE launch                       (…distrib.py:405-407, :417)
// All HBM/L2-bound streaming work: one warp per row, 128-bit accesses where the layout allows,
// fixed-order reductions (deterministic, no float atomics).
#include "common.cuh"

namespace gist {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------- layer norm ----
// y[r, :] = act((x[r, :] - mean_r) * rstd_r),  var biased (torch.nn.LayerNorm), no affine.
// stats[r] = (mean_r, rstd_r) kept for the backward.
template <bool VEC4>
__global__ void __launch_bounds__(256) ln_act_fwd_kernel(const float *__restrict__ x, int64_t ldx, int n,
                                                         int d, float eps, int relu,
                                                         float *__restrict__ y, int64_t ldy,
                                                         float2 *__restrict__ stats) {
    pdl_sync();
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const float *xr = x + (int64_t)r * ldx;
    float *yr = y + (int64_t)r * ldy;
    float s = 0.f;
    if constexpr (VEC4) {
        for (int c = lane * 4; c < d; c += 128) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(xr + c));
            s += (v.x + v.y) + (v.z + v.w);
        }
    } else {
        for (int c = lane; c < d; c += 32) s += __ldg(xr + c);
    }
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;     // two-pass variance: the row is in L1 after the first pass
    if constexpr (VEC4) {
        for (int c = lane * 4; c < d; c += 128) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(xr + c));
            const float a = v.x - mean, b = v.y - mean, e = v.z - mean, f = v.w - mean;
            q += (a * a + b * b) + (e * e + f * f);
        }
    } else {
        for (int c = lane; c < d; c += 32) {
            const float a = __ldg(xr + c) - mean;
            q += a * a;
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
    if (lane == 0 && stats) stats[r] = make_float2(mean, rstd);
    if constexpr (VEC4) {
        for (int c = lane * 4; c < d; c += 128) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(xr + c));
            float4 o = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd,
                                   (v.w - mean) * rstd);
            if (relu) {
                o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            }
            *reinterpret_cast<float4 *>(yr + c) = o;
        }
    } else {
        for (int c = lane; c < d; c += 32) {
            float o = (__ldg(xr + c) - mean) * rstd;
            if (relu) o = fmaxf(o, 0.f);
            yr[c] = o;
        }
    }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * [xhat > 0] (ReLU) or dy.
template <bool VEC4>
__global__ void __launch_bounds__(256) ln_act_bwd_kernel(const float *__restrict__ dy, int64_t lddy,
                                                         const float *__restrict__ x, int64_t ldx,
                                                         const float2 *__restrict__ stats, int n, int d,
                                                         int relu, float *__restrict__ dx, int64_t lddx,
                                                         float *__restrict__ dx_lo, int64_t ld_lo) {
    pdl_sync();
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const float *xr = x + (int64_t)r * ldx;
    const float *gr = dy + (int64_t)r * lddy;
    float *or_ = dx + (int64_t)r * lddx;
    float *lo_ = dx_lo ? dx_lo + (int64_t)r * ld_lo : nullptr;     // 3xTF32 low half of dx
    const float2 st = stats[r];
    const float mean = st.x, rstd = st.y;
    float s1 = 0.f, s2 = 0.f;
    auto acc = [&](float xv, float gv) {
        const float xh = (xv - mean) * rstd;
        const float g = (relu && xh <= 0.f) ? 0.f : gv;
        s1 += g;
        s2 += g * xh;
    };
    if constexpr (VEC4) {
        for (int c = lane * 4; c < d; c += 128) {
            const float4 xv = __ldg(reinterpret_cast<const float4 *>(xr + c));
            const float4 gv = __ldg(reinterpret_cast<const float4 *>(gr + c));
            acc(xv.x, gv.x); acc(xv.y, gv.y); acc(xv.z, gv.z); acc(xv.w, gv.w);
        }
    } else {
        for (int c = lane; c < d; c += 32) acc(__ldg(xr + c), __ldg(gr + c));
    }
    const float m1 = warp_sum(s1) / (float)d, m2 = warp_sum(s2) / (float)d;
    auto out = [&](float xv, float gv) {
        const float xh = (xv - mean) * rstd;
        const float g = (relu && xh <= 0.f) ? 0.f : gv;
        return rstd * (g - m1 - xh * m2);
    };
    if constexpr (VEC4) {
        for (int c = lane * 4; c < d; c += 128) {
            const float4 xv = __ldg(reinterpret_cast<const float4 *>(xr + c));
            const float4 gv = __ldg(reinterpret_cast<const float4 *>(gr + c));
            const float4 o = make_float4(out(xv.x, gv.x), out(xv.y, gv.y), out(xv.z, gv.z), out(xv.w, gv.w));
            *reinterpret_cast<float4 *>(or_ + c) = o;
            if (lo_)
                *reinterpret_cast<float4 *>(lo_ + c) =
                    make_float4(tf32_lo(o.x), tf32_lo(o.y), tf32_lo(o.z), tf32_lo(o.w));
        }
    } else {
        for (int c = lane; c < d; c += 32) {
            const float o = out(__ldg(xr + c), __ldg(gr + c));
            or_[c] = o;
            if (lo_) lo_[c] = tf32_lo(o);
        }
    }
}

// ------------------------------------------------- whole-tensor layer norm ----
// gcn/gcn.py:65-66 normalises with F.layer_norm(h, h.shape): ONE mean / variance over all n*d
// elements, no affine.  (ATen runs that as a single-CTA row reduction.)  Two launches each way:
// per-CTA partial moments in fp64 -> every CTA of the second launch folds the partials in the
// same fixed order (deterministic, no atomics) and streams its share of the elements.
constexpr int kTlnThreads = 256;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of two doubles; result valid in every thread
__device__ __forceinline__ void block_sum2_d(double &a, double &b) {
    __shared__ double sa[kTlnThreads / 32], sb[kTlnThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warp_sum_d(a);
    b = warp_sum_d(b);
    __syncthreads();                       // protects reuse of sa/sb between calls
    if (lane == 0) { sa[warp] = a; sb[warp] = b; }
    __syncthreads();
    double ta = 0.0, tb = 0.0;
#pragma unroll
    for (int w = 0; w < kTlnThreads / 32; ++w) { ta += sa[w]; tb += sb[w]; }
    a = ta;
    b = tb;
}

// fold the per-CTA partials (a_p, b_p) in a fixed order
__device__ __forceinline__ void fold_partials(const double2 *__restrict__ part, int nparts, double &a, double &b) {
    a = 0.0;
    b = 0.0;
    for (int p = threadIdx.x; p < nparts; p += kTlnThreads) {
        const double2 v = part[p];
        a += v.x;
        b += v.y;
    }
    block_sum2_d(a, b);
}

// Elements are visited in units of VEC floats (VEC = 4 needs d, ld % 4 == 0 and 16-byte bases);
// unit u of the tensor is row u / (d / VEC), columns VEC * (u % (d / VEC)) ...
template <int VEC>
struct TlnVec;
template <>
struct TlnVec<4> {
    using T = float4;
    __device__ static void get(const float *p, float (&v)[4]) {
        const float4 q = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    __device__ static void put(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct TlnVec<1> {
    using T = float;
    __device__ static void get(const float *p, float (&v)[1]) { v[0] = __ldg(p); }
    __device__ static void put(float *p, const float (&v)[1]) { p[0] = v[0]; }
};

template <int VEC>
__global__ void __launch_bounds__(kTlnThreads) tln_moments_kernel(const float *__restrict__ x, int64_t ldx,
                                                                  int64_t units, int upr,
                                                                  double2 *__restrict__ part) {
    double s = 0.0, q = 0.0;
    for (int64_t u = (int64_t)blockIdx.x * kTlnThreads + threadIdx.x; u < units;
         u += (int64_t)gridDim.x * kTlnThreads) {
        const int64_t r = u / upr;
        const int c = (int)(u - r * upr) * VEC;
        float v[VEC];
        TlnVec<VEC>::get(x + r * ldx + c, v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            s += (double)v[i];
            q += (double)v[i] * (double)v[i];
        }
    }
    block_sum2_d(s, q);
    if (threadIdx.x == 0) part[blockIdx.x] = make_double2(s, q);
}

template <int VEC>
__global__ void __launch_bounds__(kTlnThreads) tln_apply_kernel(const float *__restrict__ x, int64_t ldx,
                                                                int64_t units, int upr, double inv_count,
                                                                float eps, const double2 *__restrict__ part,
                                                                int nparts, float *__restrict__ y, int64_t ldy,
                                                                float *__restrict__ stats) {
    double s, q;
    fold_partials(part, nparts, s, q);
    const double mean_d = s * inv_count;
    const double var_d = fmax(q * inv_count - mean_d * mean_d, 0.0);
    const float mean = (float)mean_d;
    const float rstd = (float)(1.0 / sqrt(var_d + (double)eps));
    if (blockIdx.x == 0 && threadIdx.x == 0 && stats) {
        stats[0] = mean;
        stats[1] = rstd;
    }
    for (int64_t u = (int64_t)blockIdx.x * kTlnThreads + threadIdx.x; u < units;
         u += (int64_t)gridDim.x * kTlnThreads) {
        const int64_t r = u / upr;
        const int c = (int)(u - r * upr) * VEC;
        float v[VEC];
        TlnVec<VEC>::get(x + r * ldx + c, v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = (v[i] - mean) * rstd;
        TlnVec<VEC>::put(y + r * ldy + c, v);
    }
}

// backward partials: (sum dy, sum dy * xhat), xhat recomputed from the forward input
template <int VEC>
__global__ void __launch_bounds__(kTlnThreads) tln_bwd_moments_kernel(const float *__restrict__ dy, int64_t lddy,
                                                                      const float *__restrict__ x, int64_t ldx,
                                                                      const float *__restrict__ stats,
                                                                      int64_t units, int upr,
                                                                      double2 *__restrict__ part) {
    const float mean = __ldg(stats), rstd = __ldg(stats + 1);
    double s = 0.0, q = 0.0;
    for (int64_t u = (int64_t)blockIdx.x * kTlnThreads + threadIdx.x; u < units;
         u += (int64_t)gridDim.x * kTlnThreads) {
        const int64_t r = u / upr;
        const int c = (int)(u - r * upr) * VEC;
        float xv[VEC], gv[VEC];
        TlnVec<VEC>::get(x + r * ldx + c, xv);
        TlnVec<VEC>::get(dy + r * lddy + c, gv);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            s += (double)gv[i];
            q += (double)gv[i] * (double)((xv[i] - mean) * rstd);
        }
    }
    block_sum2_d(s, q);
    if (threadIdx.x == 0) part[blockIdx.x] = make_double2(s, q);
}

// dx = rstd * (dy - mean(dy) - xhat * mean(dy * xhat))
template <int VEC>
__global__ void __launch_bounds__(kTlnThreads) tln_bwd_apply_kernel(const float *__restrict__ dy, int64_t lddy,
                                                                    const float *__restrict__ x, int64_t ldx,
                                                                    const float *__restrict__ stats, int64_t units,
                                                                    int upr, double inv_count,
                                                                    const double2 *__restrict__ part, int nparts,
                                                                    float *__restrict__ dx, int64_t lddx) {
    double s, q;
    fold_partials(part, nparts, s, q);
    const float m1 = (float)(s * inv_count), m2 = (float)(q * inv_count);
    const float mean = __ldg(stats), rstd = __ldg(stats + 1);
    for (int64_t u = (int64_t)blockIdx.x * kTlnThreads + threadIdx.x; u < units;
         u += (int64_t)gridDim.x * kTlnThreads) {
        const int64_t r = u / upr;
        const int c = (int)(u - r * upr) * VEC;
        float xv[VEC], gv[VEC];
        TlnVec<VEC>::get(x + r * ldx + c, xv);
        TlnVec<VEC>::get(dy + r * lddy + c, gv);
#pragma unroll
        for (int i = 0; i < VEC; ++i) gv[i] = rstd * (gv[i] - m1 - (xv[i] - mean) * rstd * m2);
        TlnVec<VEC>::put(dx + r * lddx + c, gv);
    }
}

// CTAs of the two-launch scheme: enough to fill the chip twice, never more than the work needs
static int tln_grid(int64_t count) {
    int64_t g = ceil_div64(count, (int64_t)kTlnThreads * 16);
    if (g > 2 * kNumSMs) g = 2 * kNumSMs;
    if (g < 1) g = 1;
    return (int)g;
}

// ---------------------------------------------------------------- colsum ----
// Phase 1: CTA (bx, by) sums rows [by*rows_per, ...) of columns [32*bx, 32*bx+32) -> part[by][c].
// Phase 2: out[c] = sum_by part[by][c] in index order.
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float *__restrict__ x, int64_t ldx, int n,
                                                             int d, int rows_per, float *__restrict__ part) {
    __shared__ float s[8][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const int r0 = blockIdx.y * rows_per;
    const int r1 = min(n, r0 + rows_per);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (c < d) {
        int r = r0 + warp;
        for (; r + 24 < r1; r += 32) {      // 4 independent loads in flight per thread
            a0 += __ldg(x + (int64_t)r * ldx + c);
            a1 += __ldg(x + (int64_t)(r + 8) * ldx + c);
            a2 += __ldg(x + (int64_t)(r + 16) * ldx + c);
            a3 += __ldg(x + (int64_t)(r + 24) * ldx + c);
        }
        for (; r < r1; r += 8) a0 += __ldg(x + (int64_t)r * ldx + c);
    }
    s[warp][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (warp == 0 && c < d) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s[w][lane];
        part[(int64_t)blockIdx.y * d + c] = t;
    }
}

__global__ void __launch_bounds__(256) colsum_final_kernel(const float *__restrict__ part, int nparts, int d,
                                                           float *__restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    float t = 0.f;
#pragma unroll 8
    for (int p = 0; p < nparts; ++p) t += __ldg(part + (int64_t)p * d + c);
    out[c] = t;
}

// ---------------------------------------------------------- cross entropy ----
// One warp per row: lse_r = logsumexp(logits[r, :]); row_loss[r] = mask_r ? lse_r - logits[r, y_r] : 0.
__global__ void __launch_bounds__(256) ce_rows_kernel(const float *__restrict__ logits, int64_t ld, int n,
                                                      int C, const int64_t *__restrict__ labels,
                                                      const uint8_t *__restrict__ mask,
                                                      float *__restrict__ lse, float *__restrict__ row_loss) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const float *xr = logits + (int64_t)r * ld;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, __ldg(xr + c));
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(__ldg(xr + c) - mx);
    s = warp_sum(s);
    if (lane == 0) {
        const float l = mx + logf(s);
        lse[r] = l;
        const bool m = mask ? (mask[r] != 0) : true;
        const int64_t y = labels[r];
        row_loss[r] = (m && y >= 0 && y < C) ? l - __ldg(xr + y) : 0.f;
    }
}

// Single CTA: loss = sum(row_loss) / count(mask); out[0] = loss, out[1] = 1 / count.
// A row counts iff it is masked in AND its label is a class index: rows labelled outside [0, C)
// (torch's ignore_index = -100 included) are excluded from the mean's denominator and get a zero
// gradient, as CrossEntropyLoss's ignore_index rows do.
__global__ void __launch_bounds__(1024) ce_reduce_kernel(const float *__restrict__ row_loss,
                                                         const uint8_t *__restrict__ mask,
                                                         const int64_t *__restrict__ labels, int C, int n,
                                                         float *__restrict__ out) {
    __shared__ float s_l[32], s_c[32];
    float l = 0.f, c = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        l += row_loss[i];
        const bool ok = (mask ? mask[i] != 0 : true) && (uint64_t)labels[i] < (uint64_t)C;
        c += ok ? 1.f : 0.f;
    }
    l = warp_sum(l);
    c = warp_sum(c);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_l[warp] = l; s_c[warp] = c; }
    __syncthreads();
    if (warp == 0) {
        l = lane < (blockDim.x >> 5) ? s_l[lane] : 0.f;
        c = lane < (blockDim.x >> 5) ? s_c[lane] : 0.f;
        l = warp_sum(l);
        c = warp_sum(c);
        if (lane == 0) {
            out[0] = l / c;          // 0/0 = nan, as torch's mean over an empty selection
            out[1] = 1.f / c;
        }
    }
}

// dlogits[r, c] = mask_r * (softmax(logits[r])_c - [c == y_r]) * gout / count
__global__ void __launch_bounds__(256) ce_bwd_kernel(const float *__restrict__ logits, int64_t ld, int n, int C,
                                                     const int64_t *__restrict__ labels,
                                                     const uint8_t *__restrict__ mask,
                                                     const float *__restrict__ lse,
                                                     const float *__restrict__ loss_out,
                                                     const float *__restrict__ gout,
                                                     float *__restrict__ dlogits, int64_t ldd, int ldd_fill,
                                                     float *__restrict__ dlo) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const float *xr = logits + (int64_t)r * ld;
    float *dr = dlogits + (int64_t)r * ldd;
    const int64_t y = labels[r];
    const bool m = (mask ? (mask[r] != 0) : true) && (uint64_t)y < (uint64_t)C;
    const float scale = m ? __ldg(gout) * __ldg(loss_out + 1) : 0.f;
    const float l = lse[r];
    for (int c = lane; c < ldd_fill; c += 32) {     // columns [C, ldd_fill) are row padding: zeroed
        float v = 0.f;
        if (c < C && m) v = (expf(__ldg(xr + c) - l) - (c == y ? 1.f : 0.f)) * scale;
        dr[c] = v;
        if (dlo) dlo[(int64_t)r * ldd + c] = tf32_lo(v);       // same layout as dlogits
    }
}

// Forward AND gradient seed in one launch (the trainers' loss + loss.backward() seed): every CTA
// counts the masked rows itself (n bytes, L2-resident), so dlogits is final as it is written and
// nothing waits on a reduction; the scalar loss is folded by the last CTA to arrive, over the
// per-CTA partial sums in CTA order (deterministic).  sync[0] is the arrival counter (left 0).
__global__ void __launch_bounds__(256) ce_fused_kernel(const float *__restrict__ logits, int64_t ld, int n,
                                                       int C, const int64_t *__restrict__ labels,
                                                       const uint8_t *__restrict__ mask,
                                                       float *__restrict__ dlogits, int64_t ldd, int ldd_fill,
                                                       float *__restrict__ dlo, float *__restrict__ partial,
                                                       unsigned int *__restrict__ sync,
                                                       float *__restrict__ out) {
    __shared__ int s_cnt[8];
    __shared__ float s_loss[8];
    __shared__ bool s_last;
    pdl_sync();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // This warp's row first: its loads are requested BEFORE the counting loop below and the barrier
    // behind it, so the kernel is one round of loads, not a chain of them (the launch is latency-bound:
    // 2 k rows of 41 classes).  Rows of up to 64 classes live in two registers per lane; wider rows are
    // re-read from L1 in the loops further down.
    const int r = blockIdx.x * 8 + warp;
    const bool fast = C <= 64;
    const float *xr = logits + (int64_t)(r < n ? r : 0) * ld;
    float x0 = -INFINITY, x1 = -INFINITY;
    int64_t y = -1;
    bool m = false;
    if (r < n) {
        if (fast) {
            if (lane < C) x0 = __ldg(xr + lane);
            if (lane + 32 < C) x1 = __ldg(xr + lane + 32);
        }
        y = __ldg(labels + r);
        m = mask ? (__ldg(mask + r) != 0) : true;
    }
    // rows that count: masked in AND labelled with a class index (see ce_reduce_kernel).  All loads of
    // a thread are independent (unrolled), so the loop is a couple of L2 round trips, not a chain.
    int cnt = 0;
    for (int i0 = threadIdx.x; i0 < n; i0 += 256 * 8) {     // 16 predicated loads in flight per round
        uint8_t mk[8];
        int64_t lb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = i0 + k * 256;
            mk[k] = (i < n) ? (mask ? __ldg(mask + i) : (uint8_t)1) : (uint8_t)0;
            lb[k] = (i < n) ? __ldg(labels + i) : (int64_t)-1;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) cnt += (mk[k] != 0 && (uint64_t)lb[k] < (uint64_t)C) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    cnt = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) cnt += s_cnt[w];
    const float inv = 1.f / (float)cnt;         // cnt == 0: inf, and no row is masked in
    float row_loss = 0.f;
    if (r < n) {
        float mx = fmaxf(x0, x1);
        if (!fast) {
            mx = -INFINITY;
            for (int c = lane; c < C; c += 32) mx = fmaxf(mx, __ldg(xr + c));
        }
        mx = warp_max(mx);
        float se = 0.f;
        if (fast) {
            if (lane < C) se += expf(x0 - mx);
            if (lane + 32 < C) se += expf(x1 - mx);
        } else {
            for (int c = lane; c < C; c += 32) se += expf(__ldg(xr + c) - mx);
        }
        se = warp_sum(se);
        const float l = mx + logf(se);
        m = m && (uint64_t)y < (uint64_t)C;
        if (m) {
            float xy;
            if (fast) {     // the label's logit sits in lane y % 32, register y / 32
                const float a = __shfl_sync(0xffffffffu, x0, (int)(y & 31)), b2 = __shfl_sync(0xffffffffu, x1, (int)(y & 31));
                xy = y < 32 ? a : b2;
            } else {
                xy = __ldg(xr + y);
            }
            row_loss = l - xy;
        }
        float *dr = dlogits + (int64_t)r * ldd;
        for (int c = lane; c < ldd_fill; c += 32) {     // columns [C, ldd_fill) are row padding: zeroed
            float v = 0.f;
            if (c < C && m) {
                const float xc = fast ? (c < 32 ? x0 : x1) : __ldg(xr + c);
                v = (expf(xc - l) - (c == y ? 1.f : 0.f)) * inv;
            }
            dr[c] = v;
            if (dlo) dlo[(int64_t)r * ldd + c] = tf32_lo(v);
        }
    }
    if (lane == 0) s_loss[warp] = row_loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_loss[w];
        partial[blockIdx.x] = t;
        __threadfence();
        s_last = atomicAdd(sync, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (warp == 0) {                                    // fixed order: lane-strided, then the shuffle tree
        float t = 0.f;
        for (int i = lane; i < (int)gridDim.x; i += 32) t += __ldcg(partial + i);
        t = warp_sum(t);
        if (lane == 0) {
            out[0] = t * inv;        // 0 * inf = nan, as torch's mean over an empty selection
            out[1] = inv;
            *sync = 0u;
        }
    }
}

// -------------------------------------------------------------------- adam ----
struct AdamTensor {
    float *p;
    const float *g;
    float *m;
    float *v;
    float *lo;          // optional: 3xTF32 low half of the UPDATED parameter (same flat layout as p)
    int64_t n;
    int32_t block0;     // first CTA of this tensor
    int32_t vec4;       // all four arrays 16-byte aligned
};
constexpr int kAdamMaxTensors = 24;
constexpr int kAdamChunk = 1024;    // elements per CTA (256 threads x float4)
struct AdamArgs {
    AdamTensor t[kAdamMaxTensors];
    int32_t n_tensors;
    float lr, beta1, beta2, eps, weight_decay;
};

// torch.optim.Adam (amsgrad=False, L2 weight decay): t = step + 1;
//   g += wd * p;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// step (a device float, shared by all tensors of the launch) is incremented by the LAST CTA to
// finish, after every CTA has read it.
__global__ void __launch_bounds__(256) adam_multi_kernel(const __grid_constant__ AdamArgs a,
                                                         float *__restrict__ step, int advance_step,
                                                         unsigned int *__restrict__ counter,
                                                         int64_t *__restrict__ tick) {
    pdl_sync();
    int ti = 0;
#pragma unroll 1
    for (int k = 1; k < a.n_tensors; ++k)
        if ((int)blockIdx.x >= a.t[k].block0) ti = k;
    const AdamTensor &T = a.t[ti];
    const float t = *step + 1.f;
    const float bc1 = 1.f - powf(a.beta1, t);
    const float bc2 = 1.f - powf(a.beta2, t);
    const float step_size = a.lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    const int64_t i0 = ((int64_t)(blockIdx.x - T.block0) * 256 + threadIdx.x) * 4;
    auto upd = [&](float &p, float g, float &m, float &v) {
        g += a.weight_decay * p;
        m = a.beta1 * m + (1.f - a.beta1) * g;
        v = a.beta2 * v + (1.f - a.beta2) * g * g;
        p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + a.eps);
    };
    if (T.vec4 && i0 + 3 < T.n) {          // 128-bit accesses: the ultra-wide weights are HBM streams
        float4 p = *reinterpret_cast<const float4 *>(T.p + i0);
        const float4 g = __ldg(reinterpret_cast<const float4 *>(T.g + i0));
        float4 m = *reinterpret_cast<const float4 *>(T.m + i0);
        float4 v = *reinterpret_cast<const float4 *>(T.v + i0);
        upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
        *reinterpret_cast<float4 *>(T.m + i0) = m;
        *reinterpret_cast<float4 *>(T.v + i0) = v;
        *reinterpret_cast<float4 *>(T.p + i0) = p;
        if (T.lo)
            *reinterpret_cast<float4 *>(T.lo + i0) = make_float4(tf32_lo(p.x), tf32_lo(p.y), tf32_lo(p.z), tf32_lo(p.w));
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = i0 + k;
            if (i < T.n) {
                float p = T.p[i], m = T.m[i], v = T.v[i];
                upd(p, T.g[i], m, v);
                T.m[i] = m;
                T.v[i] = v;
                T.p[i] = p;
                if (T.lo) T.lo[i] = tf32_lo(p);
            }
        }
    }
    if (advance_step) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned prev = atomicAdd(counter, 1u);
            if (prev == gridDim.x - 1) {
                *step = t;
                *counter = 0u;
                if (tick) *tick += 1;       // the dropout clock of the next training step (gist_counter_add_i64)
            }
        }
    }
}

// ------------------------------------------------------- dropout utilities ----
// out[r, c] = multiplier (0 or 1/(1-p)) of element (r, col0 + c): the mask the fused kernels apply,
// materialised (tests, and eager paths that want nn.Dropout as a stand-alone op).
__global__ void __launch_bounds__(256) dropout_mask_kernel(const DropParams dp, int n, int d, int col0,
                                                           const float *__restrict__ x, int64_t ldx,
                                                           float *__restrict__ out, int64_t ldo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n * d) return;
    const int r = (int)(i / d), c = (int)(i % d);
    const int64_t step = drop_step(dp);
    if (dp.step_saved && i == 0) *dp.step_saved = step;
    const float m = dp.p != 0.f ? drop_mult(dp, step, (uint32_t)r, (uint32_t)(col0 + c)) : 1.f;
    out[(int64_t)r * ldo + c] = x ? m * __ldg(x + (int64_t)r * ldx + c) : m;
}

__global__ void counter_add_kernel(int64_t *counter, int64_t delta) { *counter += delta; }

}  // namespace gist

using namespace gist;

extern "C" int gist_dropout_f32(const float *x, int64_t ldx, int32_t n, int32_t d, int32_t col0, float *out,
                                int64_t ldo, const gist_dropout_t *drop, gist_stream_t stream) {
    if (n < 0 || d < 0) return GIST_ERR_BADARG;
    if (n == 0 || d == 0) return GIST_OK;
    if (!out || ldo < d || (x && ldx < d)) return GIST_ERR_BADARG;
    DropParams dp;
    const int st = make_drop_params(drop, &dp);
    if (st != GIST_OK) return st;
    const int64_t items = (int64_t)n * d;
    dropout_mask_kernel<<<(unsigned)((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dp, n, d, col0, x, ldx,
                                                                                          out, ldo);
    count_launch();
    return last_error();
}

extern "C" int gist_counter_add_i64(int64_t *counter, int64_t delta, gist_stream_t stream) {
    if (!counter) return GIST_ERR_BADARG;
    counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, delta);
    count_launch();
    return last_error();
}

extern "C" int gist_layernorm_act_fwd_f32(const float *x, int64_t ldx, int32_t n, int32_t d, float eps,
                                          uint32_t flags, float *y, int64_t ldy, float *stats,
                                          gist_stream_t stream) {
    if (n < 0 || d < 0) return GIST_ERR_BADARG;
    if (n == 0 || d == 0) return GIST_OK;
    if (!x || !y || ldx < d || ldy < d) return GIST_ERR_BADARG;
    if (stats && !aligned(stats, 8)) return GIST_ERR_ALIGN;
    const bool v4 = d % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && aligned(x, 16) && aligned(y, 16);
    const int relu = (flags & GIST_ACT_RELU) ? 1 : 0;
    const unsigned grid = (unsigned)((n + 7) / 8);
    cudaStream_t s = (cudaStream_t)stream;
    float2 *st2 = reinterpret_cast<float2 *>(stats);
    const cudaError_t le = v4 ? launch_pdl(ln_act_fwd_kernel<true>, dim3(grid), dim3(256), 0, s, x, ldx, n, d, eps, relu, y, ldy, st2)
                              : launch_pdl(ln_act_fwd_kernel<false>, dim3(grid), dim3(256), 0, s, x, ldx, n, d, eps, relu, y, ldy, st2);
    count_launch();
    return le == cudaSuccess ? last_error() : (int)le;
}

extern "C" int gist_layernorm_act_bwd_f32(const float *dy, int64_t lddy, const float *x, int64_t ldx,
                                          const float *stats, int32_t n, int32_t d, uint32_t flags,
                                          float *dx, int64_t lddx, float *dx_lo, int64_t ld_lo,
                                          gist_stream_t stream) {
    if (n < 0 || d < 0) return GIST_ERR_BADARG;
    if (n == 0 || d == 0) return GIST_OK;
    if (!dy || !x || !stats || !dx || lddy < d || ldx < d || lddx < d || (dx_lo && ld_lo < d)) return GIST_ERR_BADARG;
    if (!aligned(stats, 8)) return GIST_ERR_ALIGN;
    const bool v4 = d % 4 == 0 && ldx % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0 && aligned(x, 16) &&
                    aligned(dy, 16) && aligned(dx, 16) && (!dx_lo || (aligned(dx_lo, 16) && ld_lo % 4 == 0));
    const int relu = (flags & GIST_ACT_RELU) ? 1 : 0;
    const unsigned grid = (unsigned)((n + 7) / 8);
    cudaStream_t s = (cudaStream_t)stream;
    const float2 *st = reinterpret_cast<const float2 *>(stats);
    const cudaError_t le = v4 ? launch_pdl(ln_act_bwd_kernel<true>, dim3(grid), dim3(256), 0, s, dy, lddy, x, ldx, st, n, d, relu, dx, lddx, dx_lo, ld_lo)
                              : launch_pdl(ln_act_bwd_kernel<false>, dim3(grid), dim3(256), 0, s, dy, lddy, x, ldx, st, n, d, relu, dx, lddx, dx_lo, ld_lo);
    count_launch();
    return le == cudaSuccess ? last_error() : (int)le;
}

extern "C" size_t gist_tensor_layernorm_workspace_bytes(int32_t n, int32_t d) {
    if (n <= 0 || d <= 0) return 0;
    return (size_t)tln_grid((int64_t)n * d) * sizeof(double2);
}

extern "C" int gist_tensor_layernorm_fwd_f32(const float *x, int64_t ldx, int32_t n, int32_t d, float eps,
                                             float *y, int64_t ldy, float *stats, void *workspace,
                                             size_t workspace_bytes, gist_stream_t stream) {
    if (n < 0 || d < 0) return GIST_ERR_BADARG;
    if (n == 0 || d == 0) return GIST_OK;
    if (!x || !y || ldx < d || ldy < d) return GIST_ERR_BADARG;
    if (workspace_bytes < gist_tensor_layernorm_workspace_bytes(n, d) || !workspace) return GIST_ERR_WORKSPACE;
    if (!aligned(workspace, 16)) return GIST_ERR_ALIGN;
    const int64_t count = (int64_t)n * d;
    const int grid = tln_grid(count);
    double2 *part = reinterpret_cast<double2 *>(workspace);
    const double inv = 1.0 / (double)count;
    cudaStream_t s = (cudaStream_t)stream;
    const bool v4 = d % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && aligned(x, 16) && aligned(y, 16);
    if (v4) {
        tln_moments_kernel<4><<<grid, kTlnThreads, 0, s>>>(x, ldx, count / 4, d / 4, part);
        tln_apply_kernel<4><<<grid, kTlnThreads, 0, s>>>(x, ldx, count / 4, d / 4, inv, eps, part, grid, y, ldy, stats);
    } else {
        tln_moments_kernel<1><<<grid, kTlnThreads, 0, s>>>(x, ldx, count, d, part);
        tln_apply_kernel<1><<<grid, kTlnThreads, 0, s>>>(x, ldx, count, d, inv, eps, part, grid, y, ldy, stats);
    }
    count_launch(2);
    return last_error();
}

extern "C" int gist_tensor_layernorm_bwd_f32(const float *dy, int64_t lddy, const float *x, int64_t ldx,
                                             const float *stats, int32_t n, int32_t d, float *dx, int64_t lddx,
                                             void *workspace, size_t workspace_bytes, gist_stream_t stream) {
    if (n < 0 || d < 0) return GIST_ERR_BADARG;
    if (n == 0 || d == 0) return GIST_OK;
    if (!dy || !x || !stats || !dx || lddy < d || ldx < d || lddx < d) return GIST_ERR_BADARG;
    if (workspace_bytes < gist_tensor_layernorm_workspace_bytes(n, d) || !workspace) return GIST_ERR_WORKSPACE;
    if (!aligned(workspace, 16)) return GIST_ERR_ALIGN;
    const int64_t count = (int64_t)n * d;
    const int grid = tln_grid(count);
    double2 *part = reinterpret_cast<double2 *>(workspace);
    const double inv = 1.0 / (double)count;
    cudaStream_t s = (cudaStream_t)stream;
    const bool v4 = d % 4 == 0 && ldx % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0 && aligned(x, 16) &&
                    aligned(dy, 16) && aligned(dx, 16);
    if (v4) {
        tln_bwd_moments_kernel<4><<<grid, kTlnThreads, 0, s>>>(dy, lddy, x, ldx, stats, count / 4, d / 4, part);
        tln_bwd_apply_kernel<4><<<grid, kTlnThreads, 0, s>>>(dy, lddy, x, ldx, stats, count / 4, d / 4, inv, part,
                                                            grid, dx, lddx);
    } else {
        tln_bwd_moments_kernel<1><<<grid, kTlnThreads, 0, s>>>(dy, lddy, x, ldx, stats, count, d, part);
        tln_bwd_apply_kernel<1><<<grid, kTlnThreads, 0, s>>>(dy, lddy, x, ldx, stats, count, d, inv, part, grid, dx,
                                                            lddx);
    }
    count_launch(2);
    return last_error();
}

static int colsum_parts(int32_t n, int32_t d) {
    const int cx = (d + 31) / 32;
    int by = (2 * kNumSMs + cx - 1) / cx;        // ~2 CTAs per SM in total
    const int max_by = (n + 63) / 64;            // at least 64 rows per CTA
    if (by > max_by) by = max_by;
    if (by < 1) by = 1;
    return by;
}

extern "C" size_t gist_colsum_workspace_bytes(int32_t n, int32_t d) {
    if (n <= 0 || d <= 0) return 0;
    return (size_t)colsum_parts(n, d) * d * sizeof(float);
}

extern "C" int gist_colsum_f32(const float *x, int64_t ldx, int32_t n, int32_t d, float *out,
                               void *workspace, size_t workspace_bytes, gist_stream_t stream) {
    if (n < 0 || d < 0) return GIST_ERR_BADARG;
    if (d == 0) return GIST_OK;
    if (!out) return GIST_ERR_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(out, 0, (size_t)d * sizeof(float), s);
        return e == cudaSuccess ? GIST_OK : (int)e;
    }
    if (!x || ldx < d) return GIST_ERR_BADARG;
    if (!workspace || workspace_bytes < gist_colsum_workspace_bytes(n, d)) return GIST_ERR_WORKSPACE;
    const int by = colsum_parts(n, d);
    const int rows_per = (n + by - 1) / by;
    dim3 grid((d + 31) / 32, by);
    float *part = reinterpret_cast<float *>(workspace);
    colsum_partial_kernel<<<grid, 256, 0, s>>>(x, ldx, n, d, rows_per, part);
    colsum_final_kernel<<<(d + 255) / 256, 256, 0, s>>>(part, by, d, out);
    count_launch(2);
    return last_error();
}

extern "C" int gist_masked_ce_fwd_f32(const float *logits, int64_t ld, int32_t n, int32_t C,
                                      const int64_t *labels, const uint8_t *mask, float *lse,
                                      float *row_loss, float *loss_out, gist_stream_t stream) {
    if (n < 0 || C <= 0) return GIST_ERR_BADARG;
    if (!loss_out || (n > 0 && (!logits || !labels || !lse || !row_loss || ld < C))) return GIST_ERR_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (n > 0) {
        ce_rows_kernel<<<(n + 7) / 8, 256, 0, s>>>(logits, ld, n, C, labels, mask, lse, row_loss);
        count_launch();
    }
    ce_reduce_kernel<<<1, 1024, 0, s>>>(row_loss, mask, labels, C, n, loss_out);
    count_launch();
    return last_error();
}

extern "C" int gist_masked_ce_bwd_f32(const float *logits, int64_t ld, int32_t n, int32_t C,
                                      const int64_t *labels, const uint8_t *mask, const float *lse,
                                      const float *loss_out, const float *grad_out, float *dlogits,
                                      int64_t ldd, int32_t fill_cols, float *dlogits_lo, gist_stream_t stream) {
    if (n < 0 || C <= 0) return GIST_ERR_BADARG;
    if (n == 0) return GIST_OK;
    if (!logits || !labels || !lse || !loss_out || !grad_out || !dlogits || ld < C || fill_cols < C ||
        ldd < fill_cols)
        return GIST_ERR_BADARG;
    ce_bwd_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(logits, ld, n, C, labels, mask, lse, loss_out,
                                                                grad_out, dlogits, ldd, fill_cols, dlogits_lo);
    count_launch();
    return last_error();
}

extern "C" size_t gist_masked_ce_fused_workspace_bytes(int32_t n) {
    return (size_t)((n + 7) / 8 + 1) * sizeof(float);
}

extern "C" int gist_masked_ce_fused_f32(const float *logits, int64_t ld, int32_t n, int32_t C,
                                        const int64_t *labels, const uint8_t *mask, float *dlogits,
                                        int64_t ldd, int32_t fill_cols, float *dlogits_lo, float *loss_out,
                                        void *workspace, size_t workspace_bytes, uint32_t *sync,
                                        gist_stream_t stream) {
    if (n <= 0 || C <= 0) return GIST_ERR_BADARG;
    if (!logits || !labels || !dlogits || !loss_out || !workspace || !sync || ld < C || fill_cols < C ||
        ldd < fill_cols || workspace_bytes < gist_masked_ce_fused_workspace_bytes(n))
        return GIST_ERR_BADARG;
    const cudaError_t le = launch_pdl(ce_fused_kernel, dim3((unsigned)((n + 7) / 8)), dim3(256), 0, (cudaStream_t)stream,
                                      logits, ld, n, C, labels, mask, dlogits, ldd, (int)fill_cols, dlogits_lo,
                                      (float *)workspace, sync, loss_out);
    count_launch();
    return le == cudaSuccess ? last_error() : (int)le;
}

extern "C" int gist_adam_multi_ex_f32(int32_t n_tensors, float *const *params, const float *const *grads,
                                      float *const *exp_avg, float *const *exp_avg_sq, const int64_t *numel,
                                      float lr, float beta1, float beta2, float eps, float weight_decay,
                                      float *step, uint32_t *counter, float *const *params_lo, int64_t *tick,
                                      gist_stream_t stream) {
    if (n_tensors < 0) return GIST_ERR_BADARG;
    if (n_tensors == 0) return GIST_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !numel || !step || !counter) return GIST_ERR_BADARG;
    if (tick && !aligned(tick, 8)) return GIST_ERR_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    int n_live = 0;
    for (int i = 0; i < n_tensors; ++i) {
        if (numel[i] < 0) return GIST_ERR_BADARG;
        if (numel[i] == 0) continue;
        if (!params[i] || !grads[i] || !exp_avg[i] || !exp_avg_sq[i]) return GIST_ERR_BADARG;
        ++n_live;
    }
    // non-empty tensors in launches of up to kAdamMaxTensors; only the last launch advances `step` (and `tick`)
    int i = 0, done = 0;
    while (done < n_live) {
        AdamArgs a;
        a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
        int k = 0;
        int64_t blocks = 0;
        for (; i < n_tensors && k < kAdamMaxTensors; ++i) {
            if (numel[i] == 0) continue;
            a.t[k].p = params[i]; a.t[k].g = grads[i]; a.t[k].m = exp_avg[i]; a.t[k].v = exp_avg_sq[i];
            a.t[k].lo = params_lo ? params_lo[i] : nullptr;
            a.t[k].n = numel[i]; a.t[k].block0 = (int32_t)blocks;
            a.t[k].vec4 = (aligned(params[i], 16) && aligned(grads[i], 16) && aligned(exp_avg[i], 16) &&
                           aligned(exp_avg_sq[i], 16) && (!a.t[k].lo || aligned(a.t[k].lo, 16))) ? 1 : 0;
            blocks += (numel[i] + kAdamChunk - 1) / kAdamChunk;
            if (blocks > 0x7fffffffLL) return GIST_ERR_UNSUPPORTED;
            ++k;
        }
        a.n_tensors = k;
        done += k;
        const int last = done >= n_live ? 1 : 0;
        const cudaError_t le = launch_pdl(adam_multi_kernel, dim3((unsigned)blocks), dim3(256), 0, s, a, step, last, counter,
                                          last ? tick : (int64_t *)nullptr);
        count_launch();
        if (le != cudaSuccess) return (int)le;
        const int st = last_error();
        if (st != GIST_OK) return st;
    }
    return GIST_OK;
}

extern "C" int gist_adam_multi_f32(int32_t n_tensors, float *const *params, const float *const *grads,
                                   float *const *exp_avg, float *const *exp_avg_sq, const int64_t *numel,
                                   float lr, float beta1, float beta2, float eps, float weight_decay,
                                   float *step, uint32_t *counter, gist_stream_t stream) {
    return gist_adam_multi_ex_f32(n_tensors, params, grads, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps,
                                  weight_decay, step, counter, nullptr, nullptr, stream);
}
