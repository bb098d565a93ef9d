// K4: TF32 tensor-core GEMM for the dense X·Wᵀ / ∂X / ∂W contractions, sm_100a.
//
//   C[M,N] = A[M,K] · B[N,K]ᵀ (+ bias[N]) (ReLU)      all row-major fp32, K contiguous
//
// tcgen05.mma.kind::tf32 (fp32 operands read straight from shared memory, mantissa
// truncated to 10 bits by the tensor core, fp32 accumulation in TMEM).  One CTA
// computes one 128 x BN output tile:
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D tiles (128-byte swizzle) of A and B
//               into a STAGES-deep shared-memory ring, mbarrier complete_tx signalling;
//   warp 1      TMEM allocator + MMA issuer: one elected lane issues 4 x (128 x BN x 8)
//               tcgen05.mma per 32-float K block, tcgen05.commit frees the ring slot;
//   warps 2-5   epilogue: tcgen05.ld the fp32 accumulator (lane = row), bias / ReLU,
//               64-byte row segments to global memory.
// TMA zero-fills out-of-bounds rows / K columns, so M, N, K need no padding; only the
// leading dimensions must be multiples of 4 floats (16-byte global strides).
#include <cuda.h>

#include "common.cuh"

namespace gist {

constexpr int kBM = 128;       // UMMA M (cta_group::1)
constexpr int kBK = 32;        // floats per K block = 128 bytes = one swizzle row
constexpr int kUmmaK = 8;      // tf32: 32 bytes of K per MMA
constexpr int kStages = 4;
constexpr int kGemmThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0,
                                            int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// Shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 128 bytes,
// 8-row groups 1024 bytes apart (SBO); LBO unused (one swizzle atom along K).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // [0,14)  start address >> 4
    d |= (uint64_t)0 << 16;                              // [16,30) leading byte offset >> 4
    d |= (uint64_t)(1024 >> 4) << 32;                    // [32,46) stride byte offset >> 4
    d |= (uint64_t)1 << 46;                              // [46,48) descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                              // [61,64) layout: SWIZZLE_128B
    return d;
}

// Instruction descriptor, kind::tf32: D = F32, A = B = TF32, both K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
          "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct GemmParams {
    float *C;
    int64_t ldc;
    const float *bias;
    int32_t M, N, K;
    int32_t relu;
    int32_t vec4;   // C rows and bias are 16-byte aligned: float4 epilogue stores
};

template <int BN>
struct GemmSmem {
    float a[kStages][kBM * kBK];   // 16 KB per stage, 1024-byte aligned (swizzle atom)
    float b[kStages][BN * kBK];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full;
    uint32_t tmem_base;
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    GemmSmem<BN> &sm = *reinterpret_cast<GemmSmem<BN> *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * kBM;
    const int n0 = blockIdx.x * BN;
    const int kblocks = (p.K + kBK - 1) / kBK;
    constexpr uint32_t kStageBytes = (kBM + BN) * kBK * sizeof(float);
    constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;      // power of two >= 32

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], 1);
        }
        mbar_init(&sm.tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // whole warp: allocate the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&sm.tmem_base)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = sm.tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&sm.empty[s], ph ^ 1);
                mbar_expect_tx(&sm.full[s], kStageBytes);
                tma_load_2d(&tmA, &sm.full[s], sm.a[s], kb * kBK, m0);
                tma_load_2d(&tmB, &sm.full[s], sm.b[s], kb * kBK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&sm.full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(sm.a[s]);
                const uint32_t b_addr = smem_u32(sm.b[s]);
#pragma unroll
                for (int k = 0; k < kBK / kUmmaK; ++k) {
                    const uint64_t ad = umma_desc_sw128(a_addr + k * kUmmaK * sizeof(float));
                    const uint64_t bd = umma_desc_sw128(b_addr + k * kUmmaK * sizeof(float));
                    umma_tf32(tmem_d, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&sm.empty[s]);   // slot reusable once these MMAs have read it
            }
            umma_commit(&sm.tmem_full);      // accumulator complete
        }
    } else {
        // epilogue warps 2..5: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)
        const int q = warp & 3;
        mbar_wait(&sm.tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
            float v[16];
            tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            __syncwarp();
            const bool full = p.vec4 && (n0 + c + 16 <= p.N);   // warp-uniform
            if (row < p.M) {
                float *out = p.C + (int64_t)row * p.ldc + n0 + c;
                if (full) {                                   // 16-byte aligned 64-byte segment
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        float4 r = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        if (p.bias) {
                            const float4 bb = __ldg(reinterpret_cast<const float4 *>(p.bias + n0 + c + i));
                            r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
                        }
                        if (p.relu) {
                            r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f);
                            r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
                        }
                        *reinterpret_cast<float4 *>(out + i) = r;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int col = n0 + c + i;
                        if (col < p.N) {
                            float r = v[i];
                            if (p.bias) r += __ldg(p.bias + col);
                            if (p.relu) r = fmaxf(r, 0.f);
                            out[i] = r;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols)
                     : "memory");
    }
}

// ---------------------------------------------------------------- transpose ----
__global__ void __launch_bounds__(256) transpose_kernel(const float *__restrict__ src, int64_t ld_src,
                                                        int rows, int cols, float *__restrict__ dst,
                                                        int64_t ld_dst) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;   // x walks the (long) row dimension
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int r = r0 + ty + j, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + j][tx] = __ldg(src + (int64_t)r * ld_src + c);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int c = c0 + ty + j, r = r0 + tx;      // dst[c, r]
        if (r < rows && c < cols) dst[(int64_t)c * ld_dst + r] = tile[tx][ty + j];
    }
}

// ------------------------------------------------------------ host helpers ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 2-D fp32 tensor map over a row-major [rows, k] matrix: box = [box_rows, 32 floats], 128B swizzle.
static int make_map(CUtensorMap *map, const float *base, int64_t rows, int64_t k, int64_t ld, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return GIST_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GIST_OK : GIST_ERR_BADARG;
}

template <int BN>
static int launch_gemm(const CUtensorMap &ta, const CUtensorMap &tb, const GemmParams &p, cudaStream_t s) {
    constexpr size_t smem = sizeof(GemmSmem<BN>) + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tn_tf32_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    dim3 grid((p.N + BN - 1) / BN, (p.M + kBM - 1) / kBM);
    gemm_tn_tf32_kernel<BN><<<grid, kGemmThreads, smem, s>>>(ta, tb, p);
    count_launch();
    return last_error();
}

}  // namespace gist

using namespace gist;

extern "C" int gist_gemm_tn_tf32(const float *A, int64_t lda, const float *B, int64_t ldb, float *C,
                                 int64_t ldc, int32_t M, int32_t N, int32_t K, const float *bias,
                                 uint32_t flags, gist_stream_t stream) {
    if (M < 0 || N < 0 || K < 0) return GIST_ERR_BADARG;
    if (M == 0 || N == 0) return GIST_OK;
    if (!A || !B || !C || K == 0) return GIST_ERR_BADARG;
    if (lda < K || ldb < K || ldc < N) return GIST_ERR_BADARG;
    // TMA: 16-byte aligned base and 16-byte multiple row stride
    if (!aligned(A, 16) || !aligned(B, 16) || (lda % 4) || (ldb % 4) || !aligned(C, 4)) return GIST_ERR_ALIGN;
    GemmParams p;
    p.C = C; p.ldc = ldc; p.bias = bias; p.M = M; p.N = N; p.K = K;
    p.relu = (flags & GIST_GEMM_RELU) ? 1 : 0;
    p.vec4 = (aligned(C, 16) && ldc % 4 == 0 && (!bias || aligned(bias, 16))) ? 1 : 0;
    // narrow tiles when the grid would otherwise leave most of the 148 SMs idle
    const int64_t tiles128 = (int64_t)((M + kBM - 1) / kBM) * ((N + 127) / 128);
    const bool wide = tiles128 >= 2 * kNumSMs;
    CUtensorMap ta, tb;
    int st = make_map(&ta, A, M, K, lda, kBM);
    if (st != GIST_OK) return st;
    st = make_map(&tb, B, N, K, ldb, wide ? 128 : 64);
    if (st != GIST_OK) return st;
    return wide ? launch_gemm<128>(ta, tb, p, (cudaStream_t)stream)
                : launch_gemm<64>(ta, tb, p, (cudaStream_t)stream);
}

extern "C" int gist_transpose_f32(const float *src, int64_t ld_src, int32_t rows, int32_t cols, float *dst,
                                  int64_t ld_dst, gist_stream_t stream) {
    if (rows < 0 || cols < 0) return GIST_ERR_BADARG;
    if (rows == 0 || cols == 0) return GIST_OK;
    if (!src || !dst || ld_src < cols || ld_dst < rows) return GIST_ERR_BADARG;
    dim3 grid((rows + 31) / 32, (cols + 31) / 32);
    if (grid.y > 65535) return GIST_ERR_UNSUPPORTED;
    transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, rows, cols, dst, ld_dst);
    count_launch();
    return last_error();
}
