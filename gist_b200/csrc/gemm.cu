// K4: TF32 tensor-core GEMM for the dense X·Wᵀ / ∂X / ∂W contractions, sm_100a.
//
//   C[M,N] = op(A)[M,K] · op(B)[N,K]ᵀ (+ bias[N]) (ReLU)        fp32 in HBM, fp32 accumulate
//
// tcgen05.mma.kind::tf32: fp32 operands are read straight from shared memory (the tensor
// core keeps 10 mantissa bits), accumulators live in TMEM.  Each operand may be stored
// K-major ([rows, K] row-major) or MN-major ([K, rows] row-major), so the three
// contractions of a linear layer — y = z Wᵀ, dz = dy W, dW = dyᵀ z — run on the SAME
// kernel without materialising a transpose.
//
// Persistent, warp-specialised CTA (one per SM, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes (128-byte swizzle) of A and B into
//               a STAGES-deep shared-memory ring, mbarrier complete_tx signalling;
//   warp 1      TMEM allocator + MMA issuer: one lane issues 4 x (128 x BN x 8) tcgen05.mma per
//               32-float K block; tcgen05.commit frees the ring slot / publishes the accumulator;
//   warps 2-5   epilogue: tcgen05.ld (lane = row) -> bias / ReLU -> global.  The accumulator is
//               double-buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps
//               the main loop of tile i+1.
// Work unit = (m-tile, n-tile, k-split).  Split-K (for the dW contraction: few output tiles,
// K = batch nodes) writes fp32 partials to a caller-provided workspace which a second kernel
// sums in a fixed order -> deterministic, no atomics.
// TMA zero-fills out-of-bounds rows / K columns, so M, N, K need no padding; only base
// pointers (16 B) and leading dimensions (multiple of 4 floats) are constrained.
//
// NT = 3 ("3xTF32", fp32-accurate): every operand comes as a pair (x, x_lo) with
// x_lo = tf32(x - trunc_tf32(x)) produced by split_tf32_kernel.  The tensor core drops the low
// 13 mantissa bits of x, i.e. multiplies trunc_tf32(x); three MMAs per K step
//     D += A_lo·B + A·B_lo + A·B        (only the A_lo·B_lo term, <= 2^-20 relative, is dropped)
// recover ~2^-20 relative accuracy per product with fp32 accumulation in TMEM, which keeps the
// layer outputs and gradients within the 1e-5 parity tolerance of the reference's fp32 sgemm.
#include <cuda.h>

#include "common.cuh"

namespace gist {

constexpr int kBM = 128;       // UMMA M (cta_group::1)
constexpr int kBK = 32;        // floats per K block = 128 bytes = one swizzle row
constexpr int kUmmaK = 8;      // tf32: 32 bytes of K per MMA
constexpr int kGemmThreads = 192;
constexpr int kSmemBudget = 200 * 1024;   // operand ring; barriers etc. on top (227 KB max per CTA)

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

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// 1-D bulk copy global -> shared (TMA engine), completion on an mbarrier; size % 16 == 0, 16-byte aligned
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Shared-memory matrix descriptors (sm_100 format, 128-byte swizzle).
//   K-major operand  : rows of 32 floats (128 B), 8-row groups 1024 B apart (SBO); LBO unused.
//                      One MMA reads 8 floats of K: start address + 32 B per K step.
//   MN-major operand : 32-bit MN-major operands exist only in the "128B swizzle, 32B atom" layout
//                      (layout type 1; TMA mode SWIZZLE_128B_ATOM_32B: the four 32-byte chunks of
//                      a 128-byte row are XORed with the row index mod 4).  The tile is a row of
//                      [32 floats of MN] x [32 K rows] boxes of 4096 B (what one TMA box writes);
//                      inside a box K row r is at r*128 B; a swizzle atom is 4 K rows (512 B).
//                      One MMA reads 8 K rows: start address + 1024 B per K step, SBO = 512
//                      (next 4 K rows), LBO = 4096 (next 32 floats of MN).
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // [0,14)  start address >> 4
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;    // [16,30) leading byte offset >> 4
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // [32,46) stride byte offset >> 4
    d |= (uint64_t)1 << 46;                              // [46,48) descriptor version (sm_100)
    d |= (uint64_t)layout_type << 61;                    // [61,64) swizzle layout
    return d;
}

// Instruction descriptor, kind::tf32: D = F32, A = B = TF32, M x N tile; bit 15 / 16 = A / B is
// MN-major ("transposed").
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
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

// One leader lane of a fully converged warp (the same lane every time: the MMAs and the commits that
// track them must come from one thread).  The warp-role loops run on ALL lanes with warp-uniform values
// so that descriptors and addresses live in uniform registers; only the async instruction itself sits
// under this predicate.  (Issuing from `if (lane == 0)` makes ptxas wrap every UTCHMMA / UTMALDG in an
// ELECT / R2UR.BROADCAST loop: ~10 extra instructions per MMA on a single thread, which made the main
// loop instruction-issue bound at ~0.57 us per K block.)
#ifdef GIST_GEMM_OLD_ISSUE      // A/B build: round 1's single-lane role loops
#define GIST_ROLE_LANES(lane) ((lane) == 0)
#define GIST_ROLE_SYNC()
__device__ __forceinline__ bool elect_one() { return true; }
#else
#define GIST_ROLE_LANES(lane) true
#define GIST_ROLE_SYNC() __syncwarp()
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
#endif

__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
          "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
          "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
          "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct GemmParams {
    float *C;           // final output, or the split-K partial buffer [splits][M][ldc]
    int64_t ldc;
    const float *bias;  // applied here only when splits == 1
    int32_t M, N, K;
    int32_t tiles_m, tiles_n, splits, kb_per_split, kblocks;
    int32_t relu;
    int32_t vec4;       // output rows and bias are 16-byte aligned: float4 epilogue stores
    int32_t kc;         // NT == 3: K blocks chained into one TMEM accumulator before it is drained
    int32_t background; // host only: launch on a third of the SMs (GIST_GEMM_BACKGROUND)
    DropParams drop;    // p != 0: C[r, c] *= dropout multiplier of (r, c) (the dz = dy W contraction)
    long long *trace;   // diagnostic (gist_gemm_set_trace): kTraceSlots clock stamps per CTA, or NULL
    // ---- extended epilogue (gist_gemm_ex_f32) ----
    // In-kernel split-K: partial tiles go to `ws` ([tile][split][BN/4][128] float4, lanes = rows:
    // coalesced), the CTA that arrives LAST at a tile's counter adds them in split order and runs the
    // epilogue — no second launch, same summation order as splitk_reduce_kernel.
    uint32_t *tile_counters;    // [tiles_m * tiles_n], zero on entry, left zero;  NULL = legacy two-kernel split-K
    float *ws;
    float *ws_rowsum;           // [tiles_m][splits][128] partial row sums
    // Row sums of op(A) over K (the bias gradient db = colsum(dy) of the dW = dy^T z contraction) from
    // the tensor core: one extra N = 16 MMA per K step against a tile of ones, in n-tile 0 only.
    float *rowsum;
    // LayerNorm (no affine) + optional ReLU over the rows of C when one tile holds the whole row
    // (N <= BN): C keeps the pre-norm values (the backward needs them), ln_y = act(LN(C)), ln_stats =
    // (mean, rstd).  Replaces ln_act_fwd_kernel behind a layer's projection.
    float *ln_y;
    int64_t ld_ln;
    float2 *ln_stats;
    float ln_eps;
    int32_t ln_relu;
};

// Phase stamps of a CTA's FIRST work unit (tools/gemm_trace.py): SM clock at  0 entry, 1 set-up done,
// 2 first TMA issued, 3 first stage landed, 4 last MMA committed, 5 accumulator seen by the epilogue,
// 6 epilogue stores done, 7 exit;  8 / 9 = %globaltimer (ns) at entry / exit;  10 = SM id.
constexpr int kTraceSlots = 12;
__device__ __forceinline__ void trace_stamp(const GemmParams &p, int slot) {
    if (p.trace) p.trace[(size_t)blockIdx.x * kTraceSlots + slot] = clock64();
}
__device__ __forceinline__ void trace_wall(const GemmParams &p, int slot) {
    if (p.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[(size_t)blockIdx.x * kTraceSlots + slot] = (long long)t;
    }
}

template <int BN, int NT>
struct GemmCfg {
    static constexpr int kOperandCopies = NT == 3 ? 2 : 1;               // (x) or (x, x_lo)
    static constexpr int kStageBytes = (kBM + BN) * kBK * 4 * kOperandCopies;
    static constexpr int kStages = kSmemBudget / kStageBytes;          // NT=1: 64: 8, 128: 6, 256: 4
    // double-buffered accumulator + (BN <= 128) two 16-column row-sum accumulators; power of two
#ifdef GIST_GEMM_TMEM_2BN       // A/B build: round 1's allocation (no row-sum accumulators)
    static constexpr int kTmemCols = 2 * BN;
#else
    static constexpr int kTmemCols = BN == 256 ? 512 : 4 * BN;
#endif
    static constexpr int kAuxCol = 2 * BN;                              // first row-sum column (BN <= 128)
    static constexpr size_t kOnesBytes = 2048;                          // 16 rows x 128 B of 1.0f (K-major, any swizzle)
    static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align*/ + kOnesBytes + 256 /*barriers*/;
};

__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return __uint_as_float(r);
}

// named barrier among the 128 epilogue threads (barrier 0 is __syncthreads)
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// MASK: the dropout-mask epilogue (gist_gemm_dropmask_f32) is compiled in.  The 3xTF32 epilogue holds the tile's
// row in registers and stores it fully unrolled, so the Philox code is replicated per 4-column group: the
// BN = 128 instantiations were 11.0 k SASS instructions with it against 4.6 k for the LayerNorm variant, which
// never masks (profiles/r2c_sass_evidence.txt) — and a one-tile CTA runs its epilogue once, cold (carrying the
// unused in-kernel split-K code there cost the replayed step 4.7 %).  Only the dz = mask * (dy W) launches use the
// MASK = true instantiations; every other launch has p.drop.p == 0 and takes the lean ones — same arithmetic.
template <int BN, bool A_MN, bool B_MN, int NT, bool LN = false, bool MASK = true>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmAl, const __grid_constant__ CUtensorMap tmBl,
                 const GemmParams p) {
    using Cfg = GemmCfg<BN, NT>;
    static_assert(!LN || BN <= 128, "the LayerNorm epilogue keeps the whole row of the tile in registers");
    // row sums of op(A) exist only where they are used — the dW = dy^T z layout (both operands MN-major),
    // tiles <= 128 wide — so every other instantiation's MMA issue loop carries no trace of them
    constexpr bool kRowsum = A_MN && B_MN && BN <= 128;
    constexpr int kStages = Cfg::kStages;
    constexpr size_t kABytes = (size_t)kStages * kBM * kBK * 4, kBBytes = (size_t)kStages * BN * kBK * 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                                ~static_cast<uintptr_t>(1023));
    float *sA = reinterpret_cast<float *>(base);                     // [stages][128*32]
    float *sB = reinterpret_cast<float *>(base + kABytes);           // [stages][BN*32]
    float *sAl = reinterpret_cast<float *>(base + kABytes + kBBytes);              // NT == 3 only
    float *sBl = reinterpret_cast<float *>(base + 2 * kABytes + kBBytes);
    float *sOnes = reinterpret_cast<float *>(base + (size_t)kStages * Cfg::kStageBytes);     // 1024-byte aligned
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + (size_t)kStages * Cfg::kStageBytes + Cfg::kOnesBytes);
    __shared__ int s_last;      // in-kernel split-K: this CTA arrived last at its tile
    uint64_t *full = bars, *empty = bars + kStages;
    uint64_t *tmem_full = bars + 2 * kStages, *tmem_empty = bars + 2 * kStages + 2;
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 4);
    uint64_t *red_bar = bars + 2 * kStages + 5;      // in-kernel split-K: partial tiles landed in shared memory

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);     // warp-uniform for the compiler too
    const int lane = threadIdx.x & 31;
    const int n_work = p.tiles_m * p.tiles_n * p.splits;

    if (threadIdx.x == 0) {
        trace_stamp(p, 0);
        trace_wall(p, 8);
        if (p.trace) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.trace[(size_t)blockIdx.x * kTraceSlots + 10] = smid;
        }
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);     // one arrival per epilogue warp
        }
        mbar_init(red_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // whole warp: allocate the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_base_slot)),
                     "r"((uint32_t)Cfg::kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (kRowsum && p.rowsum && warp >= 2) {       // the ones tile, visible to the tensor core (async proxy)
        for (int i = threadIdx.x - 64; i < (int)(Cfg::kOnesBytes / 4); i += 128) sOnes[i] = 1.0f;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_base_slot;
    // everything above touched shared memory, TMEM and kernel parameters only: under programmatic
    // dependent launch it ran while the previous kernel of the stream was draining
    pdl_sync();
    if (threadIdx.x == 0) trace_stamp(p, 1);

    if (warp == 0) {
      if (GIST_ROLE_LANES(lane)) {
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
            if constexpr (NT == 3) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAl)) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBl)) : "memory");
            }
        }
        int s = 0;
        uint32_t ph = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int split = w % p.splits;
            const int t = w / p.splits;
            const int m0 = (t % p.tiles_m) * kBM;
            const int n0 = (t / p.tiles_m) * BN;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&full[s], (uint32_t)Cfg::kStageBytes);
#pragma unroll
                    for (int part = 0; part < Cfg::kOperandCopies; ++part) {
                        const CUtensorMap *ma = part ? &tmAl : &tmA, *mb = part ? &tmBl : &tmB;
                        float *a = (part ? sAl : sA) + (size_t)s * kBM * kBK;
                        float *b = (part ? sBl : sB) + (size_t)s * BN * kBK;
                        if constexpr (A_MN) {
#pragma unroll
                            for (int j = 0; j < kBM / 32; ++j)
                                tma_load_2d(ma, &full[s], a + j * 32 * kBK, m0 + 32 * j, kb * kBK);
                        } else {
                            tma_load_2d(ma, &full[s], a, kb * kBK, m0);
                        }
                        if constexpr (B_MN) {
#pragma unroll
                            for (int j = 0; j < BN / 32; ++j)
                                tma_load_2d(mb, &full[s], b + j * 32 * kBK, n0 + 32 * j, kb * kBK);
                        } else {
                            tma_load_2d(mb, &full[s], b, kb * kBK, n0);
                        }
                    }
                    if (w == (int)blockIdx.x && kb == kb0) trace_stamp(p, 2);
                }
                GIST_ROLE_SYNC();
                if (++s == kStages) { s = 0; ph ^= 1; }
            }
        }
      }
    } else if (warp == 1) {
      if (GIST_ROLE_LANES(lane)) {
        constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN, A_MN, B_MN);
        constexpr uint32_t idesc_aux = umma_idesc_tf32(kBM, 16, A_MN, false);
        // descriptor = constant high part | (shared address >> 4); one K step (8 floats) advances the start
        // address by 32 B (K-major) or by 8 K rows = 1024 B (MN-major), i.e. adds 2 or 64 to the descriptor
        constexpr uint64_t kHiK = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)kLayoutSw128 << 61);
        constexpr uint64_t kHiMN = ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
                                   ((uint64_t)kLayoutSw128Base32 << 61);
        constexpr uint64_t kHiA = A_MN ? kHiMN : kHiK, kHiB = B_MN ? kHiMN : kHiK;
        constexpr uint32_t kStepA = A_MN ? 64 : 2, kStepB = B_MN ? 64 : 2;
        const uint32_t a0 = smem_u32(sA) >> 4, b0 = smem_u32(sB) >> 4;
        const uint32_t al0 = smem_u32(sAl) >> 4, bl0 = smem_u32(sBl) >> 4;
        const uint64_t ones_desc = kHiK | (uint64_t)(smem_u32(sOnes) >> 4);
        int s = 0;
        uint32_t ph = 0;
        int acc = 0;
        uint32_t acc_ph = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int split = w % p.splits;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
            const bool do_rowsum = kRowsum && p.rowsum != nullptr && (w / p.splits) / p.tiles_m == 0;
            // The tensor core's fp32 accumulate truncates, so error grows linearly with the
            // number of MMAs chained into one TMEM accumulator.  NT == 3 therefore closes the
            // accumulator every p.kc K blocks; the epilogue warps add the chunks in registers
            // (round-to-nearest fp32), alternating the two TMEM buffers.  NT == 1: one chunk.
            const int kc = NT == 3 ? p.kc : (kb1 - kb0);
            for (int kc0 = kb0; kc0 < kb1; kc0 += kc) {
                const int kc1 = min(kb1, kc0 + kc);
                mbar_wait(&tmem_empty[acc], acc_ph ^ 1);        // epilogue has drained this buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_addr = tmem_d + (uint32_t)(acc * BN);
                const uint32_t x_addr = tmem_d + (uint32_t)(Cfg::kAuxCol + acc * 16);
                for (int kb = kc0; kb < kc1; ++kb) {
                    mbar_wait(&full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        if (w == (int)blockIdx.x && kb == kb0) trace_stamp(p, 3);
                        const uint64_t ad0 = kHiA | (uint64_t)(a0 + (uint32_t)s * (kBM * kBK * 4 / 16));
                        const uint64_t bd0 = kHiB | (uint64_t)(b0 + (uint32_t)s * (BN * kBK * 4 / 16));
                        const uint64_t adl0 = kHiA | (uint64_t)(al0 + (uint32_t)s * (kBM * kBK * 4 / 16));
                        const uint64_t bdl0 = kHiB | (uint64_t)(bl0 + (uint32_t)s * (BN * kBK * 4 / 16));
#pragma unroll
                        for (int k = 0; k < kBK / kUmmaK; ++k) {
                            const uint64_t ad = ad0 + (uint64_t)(k * kStepA), bd = bd0 + (uint64_t)(k * kStepB);
                            const uint32_t first = (kb > kc0 || k > 0) ? 1u : 0u;
                            if constexpr (NT == 3) {     // small terms first, then the leading product
                                const uint64_t adl = adl0 + (uint64_t)(k * kStepA), bdl = bdl0 + (uint64_t)(k * kStepB);
                                umma_tf32(d_addr, adl, bd, idesc, first);
                                umma_tf32(d_addr, ad, bdl, idesc, 1u);
                                umma_tf32(d_addr, ad, bd, idesc, 1u);
                                if constexpr (kRowsum) {
                                    if (do_rowsum) {     // row sums of op(A): op(A) (lo, then hi) against ones
                                        umma_tf32(x_addr, adl, ones_desc + (uint64_t)(k * 2), idesc_aux, first);
                                        umma_tf32(x_addr, ad, ones_desc + (uint64_t)(k * 2), idesc_aux, 1u);
                                    }
                                }
                            } else {
                                umma_tf32(d_addr, ad, bd, idesc, first);
                                if constexpr (kRowsum) {
                                    if (do_rowsum) umma_tf32(x_addr, ad, ones_desc + (uint64_t)(k * 2), idesc_aux, first);
                                }
                            }
                        }
                        umma_commit(&empty[s]);          // slot reusable once these MMAs have read it
                        if (kb + 1 == kc1) {
                            umma_commit(&tmem_full[acc]);        // this chunk's accumulator is complete
                            if (w == (int)blockIdx.x && kc1 == kb1) trace_stamp(p, 4);
                        }
                    }
                    GIST_ROLE_SYNC();
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_ph ^= 1;
            }
        }
      }
    } else {
        // epilogue warps 2..5: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)
        const int q = warp & 3;
        const int trow = q * 32 + lane;              // row inside the tile
        int acc = 0;
        uint32_t acc_ph = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int split = w % p.splits;
            const int t = w / p.splits;
            const int tile_m = t % p.tiles_m;
            const int m0 = tile_m * kBM;
            const int n0 = (t / p.tiles_m) * BN;
            const int row = m0 + trow;
            // In-kernel split-K (last-arriver fold) is compiled in only with -DGIST_GEMM_INKERNEL_SPLITK: it
            // measured slower than the two-kernel form on every shape of the training step, and merely
            // carrying its code in the epilogue cost the replayed Reddit-shape step 4.7 % (0.2621 vs 0.2499
            // ms; the epilogue of a one-tile CTA runs once, cold, and is as long as its code).
#ifdef GIST_GEMM_INKERNEL_SPLITK
            const bool inker = p.splits > 1 && p.tile_counters != nullptr;     // last-arriver reduction in this kernel
#else
            constexpr bool inker = false;
#endif
            const bool legacy = p.splits > 1 && !inker;                        // partials for splitk_reduce_kernel
            const bool do_rowsum = kRowsum && p.rowsum != nullptr && n0 == 0;
            float *crow = legacy ? p.C + ((int64_t)split * p.M + row) * p.ldc : p.C + (int64_t)row * p.ldc;
            const bool fused = !legacy;
            const int ncols = min(BN, p.N - n0);     // warp-uniform
            // one 32-column segment of the finished row: bias / ReLU, then 128-bit or scalar stores
            const bool masked = MASK && fused && p.drop.p != 0.f;
            const int64_t dstep = masked ? drop_step(p.drop) : 0;
            auto store32 = [&](const float (&v)[32], int c) {
                if (row >= p.M) return;
                float *out = crow + n0 + c;
                if (p.vec4 && c + 32 <= ncols) {             // 16-byte aligned 128-byte segment
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        float4 r = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        if (fused) {
                            if (p.bias) {
                                const float4 bb = __ldg(reinterpret_cast<const float4 *>(p.bias + n0 + c + i));
                                r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
                            }
                            if (p.relu) {
                                r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f);
                                r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
                            }
                            if (masked) {        // n0 + c + i is a multiple of 4: one Philox call
                                const uint4 rnd = drop_rand4(p.drop, dstep, (uint32_t)row, (uint32_t)(n0 + c + i) >> 2);
                                r.x = rnd.x >= p.drop.thresh ? r.x * p.drop.scale : 0.f;
                                r.y = rnd.y >= p.drop.thresh ? r.y * p.drop.scale : 0.f;
                                r.z = rnd.z >= p.drop.thresh ? r.z * p.drop.scale : 0.f;
                                r.w = rnd.w >= p.drop.thresh ? r.w * p.drop.scale : 0.f;
                            }
                        }
                        *reinterpret_cast<float4 *>(out + i) = r;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (c + i < ncols) {
                            float r = v[i];
                            if (fused) {
                                if (p.bias) r += __ldg(p.bias + n0 + c + i);
                                if (p.relu) r = fmaxf(r, 0.f);
                                if (masked) r *= drop_mult(p.drop, dstep, (uint32_t)row, (uint32_t)(n0 + c + i));
                            }
                            out[i] = r;
                        }
                    }
                }
            };
            // in-kernel split-K: this unit's partial tile, [BN/4][128] float4 with lanes = rows (coalesced)
            float4 *part = reinterpret_cast<float4 *>(p.ws) + ((int64_t)t * p.splits + split) * (BN / 4) * kBM;
            auto store_part32 = [&](const float (&v)[32], int c) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    part[(int64_t)((c + i) >> 2) * kBM + trow] = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            };
            // The reducing CTA pulls the partial tiles into shared memory with bulk copies (the operand
            // ring is free: in-kernel reduction is used only when every CTA has a single work unit) —
            // all of a batch's bytes in flight at once instead of a chain of register loads — and adds
            // them in split order.  `fold(j, v, c)`: v += staged partial j, columns [c, c + 32).
            constexpr uint32_t kTileBytes = (uint32_t)kBM * BN * 4;
            constexpr int kStageTiles = kSmemBudget / kTileBytes >= 1 ? (int)(kSmemBudget / kTileBytes) : 1;
            const float4 *stage = reinterpret_cast<const float4 *>(base);
            uint32_t red_ph = 0;
            auto stage_partials = [&](int s_first, int count) {      // all 128 epilogue threads call this
                epi_bar_sync();                                      // previous batch fully consumed
                if (threadIdx.x == 64) {
                    asm volatile("fence.proxy.async;" ::: "memory"); // partials were written through the generic proxy
                    mbar_expect_tx(red_bar, (uint32_t)count * kTileBytes);
                    for (int j = 0; j < count; ++j)
                        bulk_load(base + (size_t)j * kTileBytes,
                                  reinterpret_cast<const float4 *>(p.ws) + ((int64_t)t * p.splits + s_first + j) * (BN / 4) * kBM,
                                  kTileBytes, red_bar);
                }
                mbar_wait(red_bar, red_ph);
                red_ph ^= 1;
            };
            auto fold = [&](int j, float (&v)[32], int c) {
                const float4 *ps = stage + (size_t)j * (BN / 4) * kBM;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 x = ps[((c >> 2) + i) * kBM + trow];
                    v[4 * i] += x.x; v[4 * i + 1] += x.y; v[4 * i + 2] += x.z; v[4 * i + 3] += x.w;
                }
            };
            // LayerNorm epilogue over a complete row held in registers (tiles_n == 1, n0 == 0)
            auto ln_finalize = [&](float (&x)[LN ? BN : 1]) {
                if constexpr (LN) {
                    if (row >= p.M) return;
                    float mean = 0.f;
#pragma unroll
                    for (int cc = 0; cc < BN; ++cc)
                        if (cc < ncols) {
                            if (p.bias) x[cc] += __ldg(p.bias + cc);
                            mean += x[cc];
                        }
                    mean /= (float)ncols;
                    float var = 0.f;
#pragma unroll
                    for (int cc = 0; cc < BN; ++cc)
                        if (cc < ncols) {
                            const float dlt = x[cc] - mean;
                            var += dlt * dlt;
                        }
                    const float rstd = rsqrtf(var / (float)ncols + p.ln_eps);
                    if (p.ln_stats) p.ln_stats[row] = make_float2(mean, rstd);
                    float *yrow = p.ln_y + (int64_t)row * p.ld_ln;
                    const bool v4 = p.vec4 && (reinterpret_cast<uintptr_t>(yrow) & 15) == 0;
#pragma unroll
                    for (int cc = 0; cc < BN; cc += 4) {
                        if (cc >= ncols) continue;
                        float o[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            o[i] = (x[cc + i] - mean) * rstd;
                            if (p.ln_relu) o[i] = fmaxf(o[i], 0.f);
                        }
                        if (v4 && cc + 4 <= ncols) {
                            *reinterpret_cast<float4 *>(crow + cc) = make_float4(x[cc], x[cc + 1], x[cc + 2], x[cc + 3]);
                            *reinterpret_cast<float4 *>(yrow + cc) = make_float4(o[0], o[1], o[2], o[3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (cc + i < ncols) {
                                    crow[cc + i] = x[cc + i];
                                    yrow[cc + i] = o[i];
                                }
                        }
                    }
                }
            };
            // all four epilogue warps have published their partials: one thread takes the tile's ticket
            auto arrive_is_last = [&]() -> bool {
                __threadfence();
                epi_bar_sync();
                if (threadIdx.x == 64) {
                    const unsigned prev = atomicAdd(p.tile_counters + t, 1u);
                    const int last = prev == (unsigned)(p.splits - 1);
                    if (last) p.tile_counters[t] = 0u;        // every other split has arrived: re-armed for the next launch
                    s_last = last;
                }
                epi_bar_sync();
                const bool last = s_last != 0;
                if (last) __threadfence();
                return last;
            };
            float rs = 0.f;                                   // row sum of op(A) (n-tile 0 only)
            float *rs_part = p.ws_rowsum ? p.ws_rowsum + ((int64_t)tile_m * p.splits + split) * kBM + trow : nullptr;

            if constexpr (NT == 3 || LN) {
                // the whole row of the tile in registers: lane = row, BN fp32 accumulators per thread;
                // NT == 3 sums the K chunks here (round-to-nearest fp32 adds)
                static_assert(BN <= 128, "this epilogue keeps the whole row of the tile in registers");
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.kblocks, kb0 + p.kb_per_split);
                const int kc = NT == 3 ? p.kc : (kb1 - kb0);
                float sum[BN];
#pragma unroll
                for (int i = 0; i < BN; ++i) sum[i] = 0.f;
                for (int kc0 = kb0; kc0 < kb1; kc0 += kc) {
                    mbar_wait(&tmem_full[acc], acc_ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (warp == 2 && lane == 0 && w == (int)blockIdx.x && kc0 + kc >= kb1) trace_stamp(p, 5);
#pragma unroll
                    for (int c = 0; c < BN; c += 32) {
                        if (c < ncols) {
                            float v[32];
                            tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c), v);
#pragma unroll
                            for (int i = 0; i < 32; ++i) sum[c + i] += v[i];
                        }
                    }
                    if constexpr (kRowsum) {
                        if (do_rowsum) rs += tmem_ld1(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(Cfg::kAuxCol + acc * 16));
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                    acc ^= 1;
                    if (acc == 0) acc_ph ^= 1;
                }
                bool finish = true;
                if (inker) {
#pragma unroll
                    for (int c = 0; c < BN; c += 32) {
                        if (c < ncols) {
                            float v[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = sum[c + i];
                            store_part32(v, c);
                        }
                    }
                    if (do_rowsum) *rs_part = rs;
                    finish = arrive_is_last();
                    if (finish) {                              // add the partials of all splits, in split order
#pragma unroll
                        for (int i = 0; i < BN; ++i) sum[i] = 0.f;
                        rs = 0.f;
                        for (int s1 = 0; s1 < p.splits; s1 += kStageTiles) {
                            const int nb = min(kStageTiles, p.splits - s1);
                            stage_partials(s1, nb);
                            for (int j = 0; j < nb; ++j) {
#pragma unroll
                                for (int c = 0; c < BN; c += 32) {
                                    if (c < ncols) {
                                        float v[32];
#pragma unroll
                                        for (int i = 0; i < 32; ++i) v[i] = sum[c + i];
                                        fold(j, v, c);
#pragma unroll
                                        for (int i = 0; i < 32; ++i) sum[c + i] = v[i];
                                    }
                                }
                            }
                        }
                        if (do_rowsum)
                            for (int s2 = 0; s2 < p.splits; ++s2)
                                rs += __ldcg(p.ws_rowsum + ((int64_t)tile_m * p.splits + s2) * kBM + trow);
                    }
                }
                if (finish) {
                    if constexpr (LN) {
                        ln_finalize(sum);
                    } else {
#pragma unroll
                        for (int c = 0; c < BN; c += 32) {
                            if (c < ncols) {
                                float v[32];
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] = sum[c + i];
                                store32(v, c);
                            }
                        }
                    }
                    if (do_rowsum && !legacy && row < p.M) p.rowsum[row] = rs;
                    if (do_rowsum && legacy) *rs_part = rs;          // folded by splitk_reduce_kernel
                }
            } else {
                mbar_wait(&tmem_full[acc], acc_ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp == 2 && lane == 0 && w == (int)blockIdx.x) trace_stamp(p, 5);
#pragma unroll 1
                for (int c = 0; c < ncols; c += 32) {
                    float v[32];
                    tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c), v);
                    if (inker) store_part32(v, c);
                    else store32(v, c);
                }
                if constexpr (kRowsum) {
                    if (do_rowsum) rs = tmem_ld1(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(Cfg::kAuxCol + acc * 16));
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                acc ^= 1;
                if (acc == 0) acc_ph ^= 1;
                bool finish = true;
                if (inker) {
                    if (do_rowsum) *rs_part = rs;
                    finish = arrive_is_last();
                    if (finish) {
                        if (p.splits <= kStageTiles) {         // every partial fits the staging area at once
                            stage_partials(0, p.splits);
#pragma unroll 1
                            for (int c = 0; c < ncols; c += 32) {
                                float v[32];
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] = 0.f;
                                for (int j = 0; j < p.splits; ++j) fold(j, v, c);
                                store32(v, c);
                            }
                        } else {                               // batches: running sums go through the output rows
#pragma unroll 1
                            for (int c = 0; c < ncols; c += 32) {
                                float v[32];
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] = 0.f;
                                for (int s2 = 0; s2 < p.splits; ++s2) {
                                    const float4 *ps = reinterpret_cast<const float4 *>(p.ws) + ((int64_t)t * p.splits + s2) * (BN / 4) * kBM;
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const float4 x = __ldcg(ps + (int64_t)((c >> 2) + i) * kBM + trow);
                                        v[4 * i] += x.x; v[4 * i + 1] += x.y; v[4 * i + 2] += x.z; v[4 * i + 3] += x.w;
                                    }
                                }
                                store32(v, c);
                            }
                        }
                        if (do_rowsum) {
                            rs = 0.f;
                            for (int s2 = 0; s2 < p.splits; ++s2)
                                rs += __ldcg(p.ws_rowsum + ((int64_t)tile_m * p.splits + s2) * kBM + trow);
                        }
                    }
                }
                if (finish && do_rowsum && !legacy && row < p.M) p.rowsum[row] = rs;
                if (do_rowsum && legacy) *rs_part = rs;
            }
            if (warp == 2 && lane == 0 && w == (int)blockIdx.x) trace_stamp(p, 6);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        trace_stamp(p, 7);
        trace_wall(p, 9);
    }
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d),
                     "r"((uint32_t)Cfg::kTmemCols)
                     : "memory");
    }
}

// Split-K second pass: C[r, c] = act(sum_s part[s][r][c] + bias[c]), fixed summation order.
// Row sums of op(A) (gist_gemm_ex_t.rowsum) from the per-split partials the GEMM CTAs of n-tile 0 left
// in the workspace: out[r] = sum_s part[(r / 128) * splits + s][r % 128], split order.  Runs in the tail
// blocks (blockIdx.x >= main_blocks) of the split-K second-pass kernels.
__device__ __forceinline__ void rowsum_fold(const float *__restrict__ part, int splits, int M, float *__restrict__ out,
                                            int r) {
    if (r >= M) return;
    const float *p0 = part + (int64_t)(r / kBM) * splits * kBM + (r % kBM);
    float t = 0.f;
#pragma unroll 4
    for (int s = 0; s < splits; ++s) t += __ldg(p0 + (int64_t)s * kBM);
    out[r] = t;
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float *__restrict__ part, int64_t ldp,
                                                            int splits, int M, int N,
                                                            const float *__restrict__ bias, int relu,
                                                            float *__restrict__ C, int64_t ldc, int main_blocks,
                                                            const float *__restrict__ rs_part,
                                                            float *__restrict__ rs_out) {
    pdl_sync();
    if ((int)blockIdx.x >= main_blocks) {
        rowsum_fold(rs_part, splits, M, rs_out, ((int)blockIdx.x - main_blocks) * 256 + (int)threadIdx.x);
        return;
    }
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nq = (N + 3) >> 2;
    if (i >= (int64_t)M * nq) return;
    const int r = (int)(i / nq), c = (int)(i % nq) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *p0 = part + (int64_t)r * ldp + c;      // ldp is a multiple of 4, buffer 16-byte aligned
    const int64_t sstride = (int64_t)M * ldp;
    int s = 0;
    for (; s + 4 <= splits; s += 4) {                   // four loads in flight, summed in split order
        const float4 v0 = __ldg(reinterpret_cast<const float4 *>(p0 + (int64_t)s * sstride));
        const float4 v1 = __ldg(reinterpret_cast<const float4 *>(p0 + (int64_t)(s + 1) * sstride));
        const float4 v2 = __ldg(reinterpret_cast<const float4 *>(p0 + (int64_t)(s + 2) * sstride));
        const float4 v3 = __ldg(reinterpret_cast<const float4 *>(p0 + (int64_t)(s + 3) * sstride));
        acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
        acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
        acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
        acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
    }
    for (; s < splits; ++s) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p0 + (int64_t)s * sstride));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float o[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (c + k < N) {
            float x = o[k];
            if (bias) x += __ldg(bias + c + k);
            if (relu) x = fmaxf(x, 0.f);
            C[(int64_t)r * ldc + c + k] = x;
        }
    }
}

// Split-K second pass WITH the layer norm (gist_gemm_ex_t.ln_out; N <= 128 * VPL, VPL = 1 or 2): one warp
// per row adds the row's partials in split order (lane = 4 columns of every 128-column group), adds the
// bias, and normalises the complete row with warp shuffles: C = pre-norm values, y = act(LN(C)), stats =
// (mean, rstd).  Replaces splitk_reduce_kernel + ln_act_fwd_kernel behind a layer's projection.  The
// moments are summed per lane in column order and folded with the same xor tree for either VPL.
template <int VPL>
__global__ void __launch_bounds__(256) splitk_reduce_ln_kernel(const float *__restrict__ part, int64_t ldp, int splits,
                                                               int M, int N, const float *__restrict__ bias, float eps,
                                                               int relu, float *__restrict__ C, int64_t ldc,
                                                               float *__restrict__ y, int64_t ldy,
                                                               float2 *__restrict__ stats) {
    pdl_sync();
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= M) return;
    const int64_t sstride = (int64_t)M * ldp;
    float x[VPL][4];
#pragma unroll
    for (int g = 0; g < VPL; ++g) {
        const int c = g * 128 + lane * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < N) {
            const float *p0 = part + (int64_t)r * ldp + c;
            int s = 0;
            for (; s + 4 <= splits; s += 4) {
                const float4 v0 = __ldg(reinterpret_cast<const float4 *>(p0 + (int64_t)s * sstride));
                const float4 v1 = __ldg(reinterpret_cast<const float4 *>(p0 + (int64_t)(s + 1) * sstride));
                const float4 v2 = __ldg(reinterpret_cast<const float4 *>(p0 + (int64_t)(s + 2) * sstride));
                const float4 v3 = __ldg(reinterpret_cast<const float4 *>(p0 + (int64_t)(s + 3) * sstride));
                acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
                acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
                acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
                acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
            }
            for (; s < splits; ++s) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(p0 + (int64_t)s * sstride));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        x[g][0] = acc.x; x[g][1] = acc.y; x[g][2] = acc.z; x[g][3] = acc.w;
    }
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < VPL; ++g)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = g * 128 + lane * 4 + k;
            if (c < N) {
                if (bias) x[g][k] += __ldg(bias + c);
                sum += x[g][k];
            } else {
                x[g][k] = 0.f;
            }
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)N;
    float q = 0.f;
#pragma unroll
    for (int g = 0; g < VPL; ++g)
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (g * 128 + lane * 4 + k < N) q += (x[g][k] - mean) * (x[g][k] - mean);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)N + eps);
    if (lane == 0 && stats) stats[r] = make_float2(mean, rstd);
#pragma unroll
    for (int g = 0; g < VPL; ++g) {
        const int c = g * 128 + lane * 4;
        if (c >= N) continue;
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[k] = (x[g][k] - mean) * rstd;
            if (relu) o[k] = fmaxf(o[k], 0.f);
        }
        float *cr = C + (int64_t)r * ldc + c, *yr = y + (int64_t)r * ldy + c;
        if (c + 4 <= N && ((reinterpret_cast<uintptr_t>(cr) | reinterpret_cast<uintptr_t>(yr)) & 15) == 0) {
            *reinterpret_cast<float4 *>(cr) = make_float4(x[g][0], x[g][1], x[g][2], x[g][3]);
            *reinterpret_cast<float4 *>(yr) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (c + k < N) {
                    cr[k] = x[g][k];
                    yr[k] = o[k];
                }
        }
    }
}

// ------------------------------------------------------------ 3xTF32 split ----
// lo = tf32_rn(x - trunc_tf32(x)); optionally hi = trunc_tf32(x).  trunc_tf32 clears the 13 low
// mantissa bits, which is what kind::tf32 does to an fp32 operand, so the GEMM can use x itself as
// the leading term.  x - trunc(x) is exact in fp32; rounding it to TF32 here (instead of letting
// the MMA truncate it) halves and unbiases the residual error.  Non-finite x -> lo = 0.
template <bool V4>
__global__ void __launch_bounds__(256) split_tf32_kernel(const float *__restrict__ x, int64_t ld_x, int rows,
                                                         int cols, float *__restrict__ hi, int64_t ld_hi,
                                                         float *__restrict__ lo, int64_t ld_lo) {
    const int per_row = V4 ? (cols + 3) >> 2 : cols;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)rows * per_row) return;
    const int r = (int)(i / per_row), c = (int)(i % per_row) * (V4 ? 4 : 1);
    if constexpr (V4) {
        // rows are padded to a multiple of 4 floats (ld % 4 == 0), so a whole float4 is in bounds
        const float4 v = __ldg(reinterpret_cast<const float4 *>(x + (int64_t)r * ld_x + c));
        float4 h, l;
        l.x = tf32_lo(v.x, &h.x); l.y = tf32_lo(v.y, &h.y); l.z = tf32_lo(v.z, &h.z); l.w = tf32_lo(v.w, &h.w);
        *reinterpret_cast<float4 *>(lo + (int64_t)r * ld_lo + c) = l;
        if (hi) *reinterpret_cast<float4 *>(hi + (int64_t)r * ld_hi + c) = h;
    } else {
        float h;
        lo[(int64_t)r * ld_lo + c] = tf32_lo(__ldg(x + (int64_t)r * ld_x + c), &h);
        if (hi) hi[(int64_t)r * ld_hi + c] = h;
    }
}

// Several matrices in ONE launch (the weights of a model at the head of a training step: three
// 2 us launches in a row are three launch latencies on the step's critical path).
constexpr int kSplitMaxTensors = 16;
struct SplitTensor {
    const float *x;
    float *lo;
    int64_t ld_x, ld_lo;
    int32_t rows, cols;
    int32_t block0;     // first CTA of this tensor
    int32_t vec4;
};
struct SplitArgs {
    SplitTensor t[kSplitMaxTensors];
    int32_t n_tensors;
};

__global__ void __launch_bounds__(256) split_tf32_multi_kernel(const __grid_constant__ SplitArgs a) {
    pdl_sync();
    int ti = 0;
#pragma unroll 1
    for (int k = 1; k < a.n_tensors; ++k)
        if ((int)blockIdx.x >= a.t[k].block0) ti = k;
    const SplitTensor &T = a.t[ti];
    const int per_row = T.vec4 ? (T.cols + 3) >> 2 : T.cols;
    const int64_t i = (int64_t)(blockIdx.x - T.block0) * 256 + threadIdx.x;
    if (i >= (int64_t)T.rows * per_row) return;
    const int r = (int)(i / per_row), c = (int)(i % per_row) * (T.vec4 ? 4 : 1);
    if (T.vec4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(T.x + (int64_t)r * T.ld_x + c));
        float4 l;
        l.x = tf32_lo(v.x); l.y = tf32_lo(v.y); l.z = tf32_lo(v.z); l.w = tf32_lo(v.w);
        *reinterpret_cast<float4 *>(T.lo + (int64_t)r * T.ld_lo + c) = l;
    } else {
        T.lo[(int64_t)r * T.ld_lo + c] = tf32_lo(__ldg(T.x + (int64_t)r * T.ld_x + c));
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

// 2-D fp32 tensor map over a row-major [outer, inner] matrix with row stride ld:
// box = [box_outer rows, 32 floats], 128B swizzle (16-byte atoms for K-major operands, 32-byte
// atoms for MN-major ones), zero fill out of bounds.
static int make_map(CUtensorMap *map, const float *base, int64_t outer, int64_t inner, int64_t ld,
                    int box_outer, bool mn_major) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return GIST_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GIST_OK : GIST_ERR_BADARG;
}

static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = kNumSMs;
    }
    return n;
}

// diagnostic trace buffer (gist_gemm_set_trace); NULL in production
static long long *g_gemm_trace = nullptr;
static int g_gemm_trace_ctas = 0;

struct GemmMaps {
    CUtensorMap a, b, al, bl;     // al / bl = the x_lo operands (NT == 3); copies of a / b otherwise
};

template <int BN, bool A_MN, bool B_MN, int NT, bool LN = false, bool MASK = true>
static int launch_gemm(const GemmMaps &m, const GemmParams &p, cudaStream_t s) {
    constexpr size_t smem = GemmCfg<BN, NT>::kSmemBytes;
    static_assert(GemmCfg<BN, NT>::kStages >= 2, "operand ring needs at least two stages");
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<BN, A_MN, B_MN, NT, LN, MASK>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int n_work = p.tiles_m * p.tiles_n * p.splits;
    const int cap = p.background ? max(sm_count() / 3, 1) : sm_count();
    const int grid = n_work < cap ? n_work : cap;
    const cudaError_t le = launch_pdl(gemm_tf32_kernel<BN, A_MN, B_MN, NT, LN, MASK>, dim3(grid), dim3(kGemmThreads), smem, s,
                                      m.a, m.b, m.al, m.bl, p);
    count_launch();
    return le == cudaSuccess ? last_error() : (int)le;
}

template <int BN, int NT, bool MASK>
static int launch_layout_m(bool a_mn, bool b_mn, const GemmMaps &m, const GemmParams &p, cudaStream_t s) {
    if (a_mn) return b_mn ? launch_gemm<BN, true, true, NT, false, MASK>(m, p, s)
                          : launch_gemm<BN, true, false, NT, false, MASK>(m, p, s);
    return b_mn ? launch_gemm<BN, false, true, NT, false, MASK>(m, p, s)
                : launch_gemm<BN, false, false, NT, false, MASK>(m, p, s);
}

template <int BN, int NT>
static int launch_layout(bool a_mn, bool b_mn, const GemmMaps &m, const GemmParams &p, cudaStream_t s) {
    if constexpr (NT == 3) {       // lean epilogue unless this launch applies a dropout mask (see gemm_tf32_kernel)
        if (p.drop.p == 0.f) return launch_layout_m<BN, NT, false>(a_mn, b_mn, m, p, s);
    }
    return launch_layout_m<BN, NT, true>(a_mn, b_mn, m, p, s);
}

template <int NT>
static int launch_tile(int bn, bool a_mn, bool b_mn, const GemmMaps &m, const GemmParams &p, cudaStream_t s) {
    if constexpr (NT == 1) {
        if (bn == 256) return launch_layout<256, NT>(a_mn, b_mn, m, p, s);
    }
    if (bn == 128) return launch_layout<128, NT>(a_mn, b_mn, m, p, s);
    return launch_layout<64, NT>(a_mn, b_mn, m, p, s);
}

constexpr uint32_t kPlanNo256 = 1u << 30;        // internal: tiles at most 128 wide (row sums live beside the accumulators)

struct GemmPlan {
    int bn, splits, kb_per_split;
    int64_t ldp;            // leading dimension of the split-K partial buffer
    size_t ws_bytes;
};

// Tile width and split-K factor: the widest tile that still gives most of the 148 SMs a work
// unit; when even 64-wide tiles leave more than half the chip idle and K is long (the dW
// contraction), K is split.
// Tile width and split-K factor from a small cost model (microseconds), fitted to ncu timelines
// of the training step on B200:
//   t = fixed + waves x max(K blocks per unit x stage bytes / (per-SM TMA rate), MMA time)
//       + split-K second pass (launch + partial-buffer traffic)
// Small problems are bound by the fixed cost and by how many SMs pull operands at once (one CTA
// streams ~110 GB/s whatever the tile), so the model prefers one full wave of short K loops over
// few long ones; large problems amortise everything and get the widest tile (least L2 traffic).
static GemmPlan plan_gemm(int M, int N, int K, uint32_t flags, bool x3, bool inkernel = false) {
    GemmPlan pl;
    // GIST_GEMM_BACKGROUND: a contraction nobody waits for yet (dW of a layer whose dz chain is still
    // running beside it) is planned for — and launched on — a third of the SMs, so the latency-critical
    // stream keeps finding free SMs instead of queueing behind 140 resident 20-us CTAs
    const int sms = (flags & GIST_GEMM_BACKGROUND) ? max(sm_count() / 3, 1) : sm_count();
    const int64_t tm = (M + kBM - 1) / kBM;
    const int kblocks = (K + kBK - 1) / kBK;
    int cand[3], nc = 0;
    // 3xTF32 sums K chunks in registers (one tile row per epilogue thread): tiles are <= 128 wide
    if (flags & GIST_GEMM_TILE_N64) cand[nc++] = 64;
    else if (flags & GIST_GEMM_TILE_N128) cand[nc++] = 128;
    else if (flags & GIST_GEMM_TILE_N256) cand[nc++] = x3 ? 128 : 256;
    else {
        if (!x3 && !(flags & kPlanNo256)) cand[nc++] = 256;
        cand[nc++] = 128;
        cand[nc++] = 64;
    }
    int max_splits = 1;
    if (!(flags & GIST_GEMM_NO_SPLITK)) {
        max_splits = kblocks / 4;               // at least 4 K blocks per split
        if (max_splits > 32) max_splits = 32;
        if (max_splits < 1) max_splits = 1;
    }
    double best = 1e30;
    pl.bn = cand[0]; pl.splits = 1; pl.kb_per_split = kblocks;
    for (int ci = 0; ci < nc; ++ci) {
        const int bn = cand[ci];
        const int64_t t = tm * ((N + bn - 1) / bn);
        const double stage_bytes = (double)(kBM + bn) * kBK * 4 * (x3 ? 2 : 1);
        const double kb_load_us = stage_bytes / 110e3;                               // ~110 GB/s per CTA
        const double kb_mma_us = (x3 ? 3 : 1) * (kBK / kUmmaK) * (bn / 64.0) * 0.017;  // 128x64x8 ~ 32 clk
        const double kb_us = kb_load_us > kb_mma_us ? kb_load_us : kb_mma_us;
        const double epi_us = 0.8 * bn / 64.0;
        for (int sp = 1; sp <= max_splits; ++sp) {
            const int kbps = (kblocks + sp - 1) / sp;
            const int s_eff = (kblocks + kbps - 1) / kbps;          // no empty split
            if (s_eff != sp) continue;
            const int64_t units = t * s_eff;
            if (sp > 1 && units > sms) break;                        // splitting past one wave never pays
            const double waves = (double)((units + sms - 1) / sms);
            double cost = 6.0 + waves * (kbps * kb_us + epi_us);
            // chip-wide operand traffic (L2 -> SM) bounds large problems
            const double agg_us = (double)units * kbps * stage_bytes / 14e6;
            if (agg_us > cost) cost = agg_us;
            if (s_eff > 1) {
                // second pass: a launch + chip-wide fold (legacy), or the tile's last CTA reading
                // s_eff partial tiles at one SM's L2 rate (in-kernel)
                if (inkernel) cost += 2.0 + (double)s_eff * kBM * bn * 4.0 / 80e3;
                else cost += 4.0 + (double)(s_eff + 1) * M * N * 4.0 / 3e6;
            }
            if (cost < best - 0.05) {
                best = cost;
                pl.bn = bn; pl.splits = s_eff; pl.kb_per_split = kbps;
            }
        }
    }
    pl.ldp = (N + 3) / 4 * 4;
    pl.ws_bytes = 0;
    if (pl.splits > 1) {
        const size_t tm2 = (size_t)tm, tn2 = (size_t)((N + pl.bn - 1) / pl.bn);
        const size_t legacy = ((size_t)pl.splits * M * pl.ldp + tm2 * pl.splits * kBM) * sizeof(float);
        const size_t inker = (tm2 * tn2 * pl.splits * kBM * pl.bn + tm2 * pl.splits * kBM) * sizeof(float);
        pl.ws_bytes = inkernel ? inker : legacy;
    }
    return pl;
}

}  // namespace gist

using namespace gist;

extern "C" size_t gist_gemm_tf32_workspace_bytes(int32_t M, int32_t N, int32_t K, uint32_t flags) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    return plan_gemm(M, N, K, flags, false).ws_bytes;
}

extern "C" int gist_gemm_plan(int32_t M, int32_t N, int32_t K, uint32_t flags, int32_t three_pass, int32_t *tile_n,
                              int32_t *splits, int32_t *kblocks_per_split) {
    if (M <= 0 || N <= 0 || K <= 0) return GIST_ERR_BADARG;
    const GemmPlan pl = plan_gemm(M, N, K, flags, three_pass != 0);
    if (tile_n) *tile_n = pl.bn;
    if (splits) *splits = pl.splits;
    if (kblocks_per_split) *kblocks_per_split = pl.kb_per_split;
    return GIST_OK;
}

extern "C" int gist_gemm_set_trace(void *buffer, int32_t max_ctas) {
    g_gemm_trace = reinterpret_cast<long long *>(buffer);
    g_gemm_trace_ctas = buffer ? max_ctas : 0;
    return GIST_OK;
}

extern "C" int gist_gemm_trace_slots(void) { return kTraceSlots; }

extern "C" size_t gist_gemm_3xtf32_workspace_bytes(int32_t M, int32_t N, int32_t K, uint32_t flags) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    return plan_gemm(M, N, K, flags, true).ws_bytes;
}

// tile-width restrictions of the extended epilogue, folded into the planner's flags
static uint32_t ex_plan_flags(int32_t N, uint32_t flags, const gist_gemm_ex_t *ex) {
    if (!ex) return flags;
    if (ex->ln_out && N <= 128) {          // one tile must hold the whole row
        flags &= ~(GIST_GEMM_TILE_N64 | GIST_GEMM_TILE_N128 | GIST_GEMM_TILE_N256);
        flags |= N <= 64 ? GIST_GEMM_TILE_N64 : GIST_GEMM_TILE_N128;
    } else if (ex->ln_out) {               // 128 < N <= 256: the norm runs in the split-K second pass (or behind the GEMM)
        flags |= kPlanNo256;
    } else if (ex->rowsum) {
        if (flags & GIST_GEMM_TILE_N256) flags = (flags & ~GIST_GEMM_TILE_N256) | GIST_GEMM_TILE_N128;
        flags |= kPlanNo256;
    }
    return flags;
}

// Shared body of the 1xTF32 and 3xTF32 entry points (A_lo == B_lo == nullptr selects 1x).
static int gemm_impl(const float *A, const float *A_lo, int64_t lda, int64_t lda_lo, int32_t a_layout,
                     const float *B, const float *B_lo, int64_t ldb, int64_t ldb_lo, int32_t b_layout, float *C,
                     int64_t ldc, int32_t M, int32_t N, int32_t K, const float *bias, uint32_t flags,
                     void *workspace, size_t workspace_bytes, const gist_dropout_t *drop, gist_stream_t stream,
                     const gist_gemm_ex_t *ex = nullptr) {
    const bool x3 = A_lo != nullptr;
    if (ex) {
        if (ex->ln_out && (N > 256 || a_layout != GIST_GEMM_K_MAJOR || b_layout != GIST_GEMM_K_MAJOR || !x3 ||
                           ex->ld_ln < N || (ex->ln_stats && !aligned(ex->ln_stats, 8))))
            return GIST_ERR_UNSUPPORTED;     // LayerNorm epilogue: y = z W^T, 3xTF32, rows of <= 256 (<= 128 in one tile)
        if (ex->ln_out && N > 128 && !ex->ln_stats) return GIST_ERR_BADARG;
        if (ex->ln_out && ex->rowsum) return GIST_ERR_UNSUPPORTED;
        if (ex->rowsum && (a_layout != GIST_GEMM_MN_MAJOR || b_layout != GIST_GEMM_MN_MAJOR))
            return GIST_ERR_UNSUPPORTED;     // row sums are compiled into the dy^T z layout only
        if (ex->tile_counters && !aligned(ex->tile_counters, 4)) return GIST_ERR_ALIGN;
        flags = ex_plan_flags(N, flags, ex);
    }
    // 128 < N <= 256 with the layer norm: no tile holds the row, so the norm rides in the two-kernel
    // split-K's second pass when the planner splits K, and runs as the row-wise kernel behind the GEMM
    // (same launches as the unfused form) when it does not
    const bool wide_ln = ex && ex->ln_out && N > 128;
#ifdef GIST_GEMM_INKERNEL_SPLITK
    const bool inkernel = ex && ex->tile_counters && !wide_ln;
#else
    const bool inkernel = false;
#endif
    DropParams dp;
    {
        const int st = make_drop_params(drop, &dp);
        if (st != GIST_OK) return st;
    }
    if (dp.p != 0.f) flags |= GIST_GEMM_NO_SPLITK;     // the mask is applied in the GEMM epilogue
    if (M < 0 || N < 0 || K < 0) return GIST_ERR_BADARG;
    if (M == 0 || N == 0) return GIST_OK;
    if (!A || !B || !C || K == 0 || (A_lo == nullptr) != (B_lo == nullptr)) return GIST_ERR_BADARG;
    if ((a_layout != GIST_GEMM_K_MAJOR && a_layout != GIST_GEMM_MN_MAJOR) ||
        (b_layout != GIST_GEMM_K_MAJOR && b_layout != GIST_GEMM_MN_MAJOR))
        return GIST_ERR_BADARG;
    const bool a_mn = a_layout == GIST_GEMM_MN_MAJOR, b_mn = b_layout == GIST_GEMM_MN_MAJOR;
    if (lda < (a_mn ? M : K) || ldb < (b_mn ? N : K) || ldc < N) return GIST_ERR_BADARG;
    if (x3 && (lda_lo < (a_mn ? M : K) || ldb_lo < (b_mn ? N : K))) return GIST_ERR_BADARG;
    // TMA: 16-byte aligned base and 16-byte multiple row stride
    if (!aligned(A, 16) || !aligned(B, 16) || (lda % 4) || (ldb % 4) || !aligned(C, 4)) return GIST_ERR_ALIGN;
    if (x3 && (!aligned(A_lo, 16) || !aligned(B_lo, 16) || (lda_lo % 4) || (ldb_lo % 4))) return GIST_ERR_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    GemmPlan pl = plan_gemm(M, N, K, flags, x3, inkernel);
    const int64_t n_tiles = (int64_t)((M + kBM - 1) / kBM) * ((N + pl.bn - 1) / pl.bn);
    const bool inker_ok = inkernel && n_tiles <= ex->n_counters && n_tiles * pl.splits <= sm_count();
    if (pl.splits > 1 && (!workspace || workspace_bytes < pl.ws_bytes || !aligned(workspace, 16) ||
                          (inkernel && !inker_ok))) {
        // no (or too small a) workspace / counter array: run unsplit rather than fail
        pl.splits = 1;
        pl.kb_per_split = (K + kBK - 1) / kBK;
    }
    GemmParams p;
    p.M = M; p.N = N; p.K = K;
    p.tiles_m = (M + kBM - 1) / kBM;
    p.tiles_n = (N + pl.bn - 1) / pl.bn;
    p.splits = pl.splits;
    p.kb_per_split = pl.kb_per_split;
    p.kblocks = (K + kBK - 1) / kBK;
    p.relu = (flags & GIST_GEMM_RELU) ? 1 : 0;
    p.background = (flags & GIST_GEMM_BACKGROUND) ? 1 : 0;
    p.drop = dp;
    p.kc = (int)((flags >> 8) & 0xFFu);
    if (p.kc == 0) p.kc = 4;          // 128 K elements = 48 chained MMAs per accumulator
    p.trace = (g_gemm_trace && g_gemm_trace_ctas >= sm_count()) ? g_gemm_trace : nullptr;
    p.tile_counters = nullptr; p.ws = nullptr; p.ws_rowsum = nullptr; p.rowsum = nullptr;
    p.ln_y = nullptr; p.ld_ln = 0; p.ln_stats = nullptr; p.ln_eps = 0.f; p.ln_relu = 0;
    if (pl.splits > 1 && !inkernel) {
        p.C = reinterpret_cast<float *>(workspace); p.ldc = pl.ldp; p.bias = nullptr; p.vec4 = 1;
        p.ws_rowsum = reinterpret_cast<float *>(workspace) + (size_t)pl.splits * M * pl.ldp;
    } else {
        p.C = C; p.ldc = ldc; p.bias = bias;
        p.vec4 = (aligned(C, 16) && ldc % 4 == 0 && (!bias || aligned(bias, 16))) ? 1 : 0;
        if (pl.splits > 1) {
            p.tile_counters = ex->tile_counters;
            p.ws = reinterpret_cast<float *>(workspace);
            p.ws_rowsum = p.ws + (size_t)n_tiles * pl.splits * kBM * pl.bn;
        }
    }
    if (ex) {
        p.rowsum = pl.bn <= 128 ? ex->rowsum : nullptr;
        if (ex->rowsum && !p.rowsum) return GIST_ERR_UNSUPPORTED;
        if (ex->ln_out) {
            if (pl.bn < N && !wide_ln) return GIST_ERR_UNSUPPORTED;
            p.relu = 0;      // the activation belongs to the normalised output
            if (!(pl.splits > 1 && !inkernel) && !wide_ln) {      // in the GEMM's own epilogue; else in the split-K second pass
                p.ln_y = ex->ln_out; p.ld_ln = ex->ld_ln; p.ln_stats = reinterpret_cast<float2 *>(ex->ln_stats);
                p.ln_eps = ex->ln_eps; p.ln_relu = (ex->ln_flags & GIST_ACT_RELU) ? 1 : 0;
            }
        }
    }
    auto map_a = [&](CUtensorMap *t, const float *ptr, int64_t ld) {
        return a_mn ? make_map(t, ptr, K, M, ld, kBK, true) : make_map(t, ptr, M, K, ld, kBM, false);
    };
    auto map_b = [&](CUtensorMap *t, const float *ptr, int64_t ld) {
        return b_mn ? make_map(t, ptr, K, N, ld, kBK, true) : make_map(t, ptr, N, K, ld, pl.bn, false);
    };
    GemmMaps m;
    int st = map_a(&m.a, A, lda);
    if (st != GIST_OK) return st;
    st = map_b(&m.b, B, ldb);
    if (st != GIST_OK) return st;
    if (x3) {
        st = map_a(&m.al, A_lo, lda_lo);
        if (st != GIST_OK) return st;
        st = map_b(&m.bl, B_lo, ldb_lo);
        if (st != GIST_OK) return st;
        if (p.ln_y) st = pl.bn == 64 ? launch_gemm<64, false, false, 3, true, false>(m, p, s)
                                     : launch_gemm<128, false, false, 3, true, false>(m, p, s);
        else st = launch_tile<3>(pl.bn, a_mn, b_mn, m, p, s);
    } else {
        m.al = m.a;
        m.bl = m.b;
        st = launch_tile<1>(pl.bn, a_mn, b_mn, m, p, s);
    }
    if (st != GIST_OK) return st;
    if (wide_ln && pl.splits == 1)      // unsplit: the row-wise kernel behind the GEMM (csrc/fused.cu)
        return gist_layernorm_act_fwd_f32(C, ldc, M, N, ex->ln_eps, ex->ln_flags, ex->ln_out, ex->ld_ln, ex->ln_stats,
                                          stream);
    if (pl.splits == 1 || inkernel) return st;
    cudaError_t le;
    if (ex && ex->ln_out) {
        const int ln_relu = (ex->ln_flags & GIST_ACT_RELU) ? 1 : 0;
        float2 *stats = reinterpret_cast<float2 *>(ex->ln_stats);
        const float *part = reinterpret_cast<const float *>(workspace);
        if (N <= 128)
            le = launch_pdl(splitk_reduce_ln_kernel<1>, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, s, part, pl.ldp,
                            pl.splits, M, N, bias, ex->ln_eps, ln_relu, C, ldc, ex->ln_out, ex->ld_ln, stats);
        else
            le = launch_pdl(splitk_reduce_ln_kernel<2>, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, s, part, pl.ldp,
                            pl.splits, M, N, bias, ex->ln_eps, ln_relu, C, ldc, ex->ln_out, ex->ld_ln, stats);
    } else {
        const int64_t items = (int64_t)M * ((N + 3) / 4);
        const int main_blocks = (int)((items + 255) / 256);
        const int rs_blocks = p.rowsum ? (M + 255) / 256 : 0;       // row sums ride in the tail blocks of the fold
        le = launch_pdl(splitk_reduce_kernel, dim3((unsigned)(main_blocks + rs_blocks)), dim3(256), 0, s,
                        reinterpret_cast<const float *>(workspace), pl.ldp, pl.splits, M, N, bias, p.relu, C, ldc,
                        main_blocks, (const float *)p.ws_rowsum, p.rowsum);
    }
    count_launch();
    return le == cudaSuccess ? last_error() : (int)le;
}

extern "C" int gist_gemm_tf32(const float *A, int64_t lda, int32_t a_layout, const float *B, int64_t ldb,
                              int32_t b_layout, float *C, int64_t ldc, int32_t M, int32_t N, int32_t K,
                              const float *bias, uint32_t flags, void *workspace, size_t workspace_bytes,
                              gist_stream_t stream) {
    return gemm_impl(A, nullptr, lda, 0, a_layout, B, nullptr, ldb, 0, b_layout, C, ldc, M, N, K, bias, flags,
                     workspace, workspace_bytes, nullptr, stream);
}

extern "C" int gist_gemm_3xtf32(const float *A, const float *A_lo, int64_t lda, int64_t lda_lo, int32_t a_layout,
                                const float *B, const float *B_lo, int64_t ldb, int64_t ldb_lo,
                                int32_t b_layout, float *C, int64_t ldc, int32_t M, int32_t N, int32_t K,
                                const float *bias, uint32_t flags, void *workspace, size_t workspace_bytes,
                                gist_stream_t stream) {
    if (!A_lo || !B_lo) return GIST_ERR_BADARG;
    return gemm_impl(A, A_lo, lda, lda_lo, a_layout, B, B_lo, ldb, ldb_lo, b_layout, C, ldc, M, N, K, bias, flags,
                     workspace, workspace_bytes, nullptr, stream);
}

extern "C" int gist_gemm_has_inkernel_splitk(void) {
#ifdef GIST_GEMM_INKERNEL_SPLITK
    return 1;
#else
    return 0;
#endif
}

extern "C" size_t gist_gemm_ex_workspace_bytes(int32_t M, int32_t N, int32_t K, uint32_t flags, int32_t three_pass,
                                               const gist_gemm_ex_t *ex) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    const bool inkernel = gist_gemm_has_inkernel_splitk() && ex && ex->tile_counters;
    return plan_gemm(M, N, K, ex_plan_flags(N, flags, ex), three_pass != 0, inkernel).ws_bytes;
}

extern "C" int gist_gemm_ex_f32(const float *A, const float *A_lo, int64_t lda, int64_t lda_lo, int32_t a_layout,
                                const float *B, const float *B_lo, int64_t ldb, int64_t ldb_lo, int32_t b_layout,
                                float *C, int64_t ldc, int32_t M, int32_t N, int32_t K, const float *bias,
                                uint32_t flags, void *workspace, size_t workspace_bytes, const gist_gemm_ex_t *ex,
                                gist_stream_t stream) {
    if ((A_lo == nullptr) != (B_lo == nullptr)) return GIST_ERR_BADARG;
    return gemm_impl(A, A_lo, lda, lda_lo, a_layout, B, B_lo, ldb, ldb_lo, b_layout, C, ldc, M, N, K, bias, flags,
                     workspace, workspace_bytes, ex ? ex->drop : nullptr, stream, ex);
}

extern "C" int gist_gemm_dropmask_f32(const float *A, const float *A_lo, int64_t lda, int64_t lda_lo,
                                      int32_t a_layout, const float *B, const float *B_lo, int64_t ldb,
                                      int64_t ldb_lo, int32_t b_layout, float *C, int64_t ldc, int32_t M,
                                      int32_t N, int32_t K, uint32_t flags, const gist_dropout_t *drop,
                                      gist_stream_t stream) {
    if ((A_lo == nullptr) != (B_lo == nullptr)) return GIST_ERR_BADARG;
    return gemm_impl(A, A_lo, lda, lda_lo, a_layout, B, B_lo, ldb, ldb_lo, b_layout, C, ldc, M, N, K, nullptr,
                     flags | GIST_GEMM_NO_SPLITK, nullptr, 0, drop, stream);
}

extern "C" int gist_split_tf32_f32(const float *x, int64_t ld_x, int32_t rows, int32_t cols, float *hi,
                                   int64_t ld_hi, float *lo, int64_t ld_lo, gist_stream_t stream) {
    if (rows < 0 || cols < 0) return GIST_ERR_BADARG;
    if (rows == 0 || cols == 0) return GIST_OK;
    if (!x || !lo || ld_x < cols || ld_lo < cols || (hi && ld_hi < cols)) return GIST_ERR_BADARG;
    const bool v4 = aligned(x, 16) && aligned(lo, 16) && (!hi || aligned(hi, 16)) && ld_x % 4 == 0 &&
                    ld_lo % 4 == 0 && (!hi || ld_hi % 4 == 0);
    const int per_row = v4 ? (cols + 3) / 4 : cols;
    const int64_t items = (int64_t)rows * per_row;
    const unsigned grid = (unsigned)((items + 255) / 256);
    if (v4) split_tf32_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, ld_x, rows, cols, hi, ld_hi, lo, ld_lo);
    else split_tf32_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, ld_x, rows, cols, hi, ld_hi, lo, ld_lo);
    count_launch();
    return last_error();
}

extern "C" int gist_split_tf32_multi_f32(int32_t n_tensors, const float *const *x, const int64_t *ld_x,
                                         const int32_t *rows, const int32_t *cols, float *const *lo,
                                         const int64_t *ld_lo, gist_stream_t stream) {
    if (n_tensors < 0 || n_tensors > kSplitMaxTensors) return GIST_ERR_BADARG;
    if (n_tensors == 0) return GIST_OK;
    if (!x || !ld_x || !rows || !cols || !lo || !ld_lo) return GIST_ERR_BADARG;
    SplitArgs a;
    a.n_tensors = 0;
    int64_t blocks = 0;
    for (int i = 0; i < n_tensors; ++i) {
        if (rows[i] < 0 || cols[i] < 0) return GIST_ERR_BADARG;
        if (rows[i] == 0 || cols[i] == 0) continue;
        if (!x[i] || !lo[i] || ld_x[i] < cols[i] || ld_lo[i] < cols[i]) return GIST_ERR_BADARG;
        SplitTensor &T = a.t[a.n_tensors++];
        T.x = x[i]; T.lo = lo[i]; T.ld_x = ld_x[i]; T.ld_lo = ld_lo[i]; T.rows = rows[i]; T.cols = cols[i];
        T.vec4 = aligned(x[i], 16) && aligned(lo[i], 16) && ld_x[i] % 4 == 0 && ld_lo[i] % 4 == 0;
        T.block0 = (int32_t)blocks;
        const int per_row = T.vec4 ? (cols[i] + 3) / 4 : cols[i];
        blocks += ((int64_t)rows[i] * per_row + 255) / 256;
        if (blocks > 0x7fffffffLL) return GIST_ERR_UNSUPPORTED;
    }
    if (blocks == 0) return GIST_OK;
    const cudaError_t le = launch_pdl(split_tf32_multi_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, a);
    count_launch();
    return le == cudaSuccess ? last_error() : (int)le;
}

extern "C" int gist_gemm_tn_tf32(const float *A, int64_t lda, const float *B, int64_t ldb, float *C,
                                 int64_t ldc, int32_t M, int32_t N, int32_t K, const float *bias,
                                 uint32_t flags, gist_stream_t stream) {
    return gist_gemm_tf32(A, lda, GIST_GEMM_K_MAJOR, B, ldb, GIST_GEMM_K_MAJOR, C, ldc, M, N, K, bias,
                          flags | GIST_GEMM_NO_SPLITK, nullptr, 0, stream);
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
