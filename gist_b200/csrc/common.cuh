// Shared helpers for the gist_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <utility>

#include "../../include/gist_b200.h"

namespace gist {

extern std::atomic<uint64_t> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

inline int last_error() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GIST_OK : (int)e;
}

inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------- programmatic dependent launch ----
// The training step is a chain of ~25 dependent kernels of 3-20 us each, so what a kernel costs
// besides its bytes — grid launch, barrier / TMEM set-up, descriptor prefetch — sits on the critical
// path once per link.  With programmatic stream serialization (gist_set_pdl / GIST_PDL=1) kernel k+1
// is made resident while kernel k drains and runs its prologue there; it blocks in
// griddepcontrol.wait until kernel k has COMPLETED and its writes are visible, so the data
// dependency is exactly the stream-order one.  Rules every kernel launched through launch_pdl
// follows: pdl_wait() comes before the first global-memory access (reads AND writes: the previous
// kernel may still be reading what this one overwrites) and before any early return; pdl_trigger()
// comes after it, so at most one dependent grid is ever pre-launched.  Under stream capture the
// attribute becomes a programmatic edge of the graph.
bool pdl_enabled();

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
    pdl_wait();
    pdl_trigger();
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (pdl_enabled()) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

// ---------------------------------------------------------------- 3xTF32 ----
// lo = tf32_rn(x - trunc_tf32(x)); *hi_out = trunc_tf32(x).  trunc_tf32 clears the 13 low
// mantissa bits, which is what tcgen05 kind::tf32 does to an fp32 operand, so the GEMM uses x
// itself as the leading term.  x - trunc(x) is exact in fp32; rounding it to TF32 here (instead
// of letting the MMA truncate it) halves and unbiases the residual error.  Non-finite x -> 0.
struct DropParams {
    float p;                 // 0 = disabled
    float scale;             // 1 / (1 - p)
    uint32_t thresh;         // keep iff rnd >= thresh
    uint32_t seed_lo, seed_hi, stream_id;
    const int64_t *step;     // device counter (NULL = 0)
    int64_t *step_saved;     // optional: the forward kernel records the step value it used
};

#ifdef __CUDACC__
__device__ __forceinline__ float tf32_lo(float x, float *hi_out) {
    const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    *hi_out = h;
    float r = x - h;
    if (!(fabsf(x) <= 3.402823466e38f)) r = 0.f;
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(r));
    return __uint_as_float(t);
}
__device__ __forceinline__ float tf32_lo(float x) {
    float h;
    return tf32_lo(x, &h);
}

// --------------------------------------------------------------- dropout ----
// Counter-based dropout mask (Philox4x32-10): keep(row, col) is a pure function of
// (seed, step, stream_id, row, col), so the forward kernel that applies it and the backward
// kernel that needs it again (the dz GEMM epilogue) regenerate it instead of storing it.
// One Philox call yields the four decisions of the aligned column group [4g, 4g+4).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ int64_t drop_step(const DropParams &d) { return d.step ? *d.step : 0; }

__device__ __forceinline__ uint4 drop_rand4(const DropParams &d, int64_t step, uint32_t row, uint32_t colgroup) {
    const uint32_t s_lo = (uint32_t)step, s_hi = (uint32_t)((uint64_t)step >> 32);
    return philox4x32_10(make_uint4(colgroup, row, s_lo, d.stream_id ^ (s_hi * 0x85EBCA6Bu)),
                         make_uint2(d.seed_lo, d.seed_hi));
}

__device__ __forceinline__ uint32_t pick4(const uint4 &r, int i) {
    return i == 0 ? r.x : (i == 1 ? r.y : (i == 2 ? r.z : r.w));
}

// multiplier (0 or scale) of element (row, col) of the logical matrix
__device__ __forceinline__ float drop_mult(const DropParams &d, int64_t step, uint32_t row, uint32_t col) {
    const uint4 r = drop_rand4(d, step, row, col >> 2);
    return pick4(r, col & 3) >= d.thresh ? d.scale : 0.f;
}
#endif

// host: fill the device-side parameter block from the public descriptor (NULL / p == 0 -> off)
inline int make_drop_params(const gist_dropout_t *in, DropParams *out) {
    out->p = 0.f; out->scale = 1.f; out->thresh = 0u;
    out->seed_lo = out->seed_hi = out->stream_id = 0u;
    out->step = nullptr; out->step_saved = nullptr;
    if (!in || in->p == 0.f) return GIST_OK;
    if (!(in->p > 0.f && in->p < 1.f)) return GIST_ERR_BADARG;
    out->p = in->p;
    out->scale = 1.f / (1.f - in->p);
    double t = (double)in->p * 4294967296.0;
    out->thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
    out->seed_lo = (uint32_t)in->seed;
    out->seed_hi = (uint32_t)(in->seed >> 32);
    out->stream_id = in->stream_id;
    out->step = in->step;
    out->step_saved = in->step_saved;
    return GIST_OK;
}

}  // namespace gist
