// K1 / K2: CSR SpMM (copy_u / sum message passing) with fused epilogue, sm_100a.
//
// Work decomposition:
//   * a GROUP of LPR lanes (4..32, power of two) owns one (row, feature-chunk);
//     a warp holds 32/LPR groups, a 256-thread CTA 8 warps of consecutive rows,
//     so neighbouring rows (which share neighbours inside a cluster) share L1;
//   * a feature chunk is LPR*VEC*VPL floats; the chunk index is the SLOW grid
//     dimension, so at any moment the whole chip gathers from one column slab
//     of X (n_src * chunk * 4 bytes) — for the full Reddit-shape graph that slab
//     (60-120 MB) is what lives in the 126 MB L2 while the slab is swept;
//   * the group loads LPR column indices with one coalesced load, then
//     broadcasts them with shuffles and issues U independent VEC-wide gathers
//     before touching the accumulators (memory-level parallelism);
//   * edges are accumulated in CSR order -> deterministic, no atomics.
#include <stdlib.h>

#include "common.cuh"

namespace gist {

std::atomic<uint64_t> g_launches{0};

// programmatic dependent launch (common.cuh): -1 = not decided yet (GIST_PDL in the environment)
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
    int v = g_pdl.load(std::memory_order_relaxed);
    if (v < 0) {
        const char *e = getenv("GIST_PDL");
        v = (e && e[0] && e[0] != '0') ? 1 : 0;
        g_pdl.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}

struct SpmmParams {
    const int32_t *rowptr;
    const int32_t *col;
    int32_t n_dst;
    int32_t n_src;
    const float *X;
    int32_t ldx;        // leading dimensions fit 32 bits (checked on the host): the row
    int32_t d;          // address is one IMAD.WIDE instead of a 64-bit multiply
    float *Y;
    int32_t ldy;
    const float *src_scale;
    const float *dst_scale;
    const float *bias;
    const float *addend;
    int32_t ld_add;
    float *self_out;
    int32_t ld_self;
    int32_t relu;
    int32_t row_blocks;
    int32_t heavy_deg;  // rows with more edges are split over the whole CTA (INT_MAX = never)
    int32_t total_blocks;   // EX variants: (row block, chunk) items; the grid may be smaller (background mode)
    // EX variants only (gist_spmm_csr_ex_f32): dropout on the two outputs and their 3xTF32 halves
    float *y_lo;
    int32_t ld_y_lo;
    float *self_lo;
    int32_t ld_self_lo;
    int32_t col0_y, col0_self;   // logical column of output column 0 in the dropout mask
    DropParams drop;
    // segment-balanced variant only (spmm_seg_kernel)
    const int32_t *seg_ptr;      // [n_dst + 1] first segment of every row; seg_ptr[n_dst] = #segments
    const int32_t *seg_row;      // [#segments] row of every segment
    int32_t seg_len;             // edges per segment
    int32_t seg_blocks;          // CTAs per feature chunk
    uint32_t *seg_count;         // [chunks][n_dst] arrival counters, zero between launches
    float *seg_ws;               // [#segments][ld_ws] partial sums of multi-segment rows
    const int4 *seg_meta;        // optional [#segments] {row, first edge, first segment of the row, (edges << 20) | segments of the row}
    uint32_t seg_flags;          // GIST_SPMM_SCHED_PREFETCH: request the next work item before processing the current one
    int32_t ld_ws;
};

template <int VEC>
__device__ __forceinline__ void ld_vec(float (&r)[VEC], const float *p) {
    if constexpr (VEC == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else if constexpr (VEC == 2) {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
        r[0] = v.x; r[1] = v.y;
    } else {
        r[0] = __ldg(p);
    }
}

template <int VEC>
__device__ __forceinline__ void st_vec(float *p, const float (&r)[VEC]) {
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4 *>(p) = make_float4(r[0], r[1], r[2], r[3]);
    } else if constexpr (VEC == 2) {
        *reinterpret_cast<float2 *>(p) = make_float2(r[0], r[1]);
    } else {
        *p = r[0];
    }
}

// Accumulate edges [eb, ee) of one row into acc (a group of LPR lanes, lane `lg`).
// Two-level summation: every block of <= LPR edges is summed on its own and then added
// to the running total, which keeps the fp32 rounding error of hub rows (10^4 edges)
// ~sqrt(LPR) lower than one long chain, at the cost of VPL*VEC adds per block.
template <int VEC, int LPR, int VPL, bool HAS_SS>
__device__ __forceinline__ void gather_range(const SpmmParams &p, int eb, int ee, int lg, int lane0,
                                             unsigned gmask, const int (&c)[VPL], const bool (&cv)[VPL],
                                             float (&acc)[VPL][VEC]) {
    constexpr int UMAX = (VPL == 1) ? 8 : 4;
    constexpr int U = (LPR < UMAX) ? LPR : UMAX;  // independent gathers in flight per lane
    // lanes past the end of the row (the last chunk's tail) gather column 0 instead of being
    // predicated off: their sums are never stored (every store checks cv), and the gathers of the
    // whole group stay free of per-lane predicates
    // One address = one IMAD.WIDE.U32: byte pointer of this lane's column(s) + u * (row pitch in
    // bytes, < 2^32 — checked on the host); and the index of gather jj of a round always comes from
    // lane jj of the group (the indices are rotated down by U after every round), so the shuffle's
    // source lane is an immediate.  The kernels are instruction-issue bound: per gathered vector
    // this is shuffle + address + load + packed adds.
    const char *xb[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) xb[k] = reinterpret_cast<const char *>(p.X + (cv[k] ? c[k] : 0));
    const uint32_t pitch = (uint32_t)p.ldx * 4u;
    // the column indices of block b+1 are requested before block b's gathers are issued, so the
    // index load of every block but the first is hidden behind a block of feature gathers
    int u_next = (eb + lg < ee) ? __ldg(p.col + eb + lg) : 0;
    for (int e0 = eb; e0 < ee; e0 += LPR) {
        const int my = e0 + lg;
        int u_rot = u_next;
        float s_rot = 0.f;
        if constexpr (HAS_SS) {
            if (my < ee) s_rot = __ldg(p.src_scale + u_rot);
        }
        u_next = (my + LPR < ee) ? __ldg(p.col + my + LPR) : 0;
        const int cnt = min(LPR, ee - e0);
        float blk[VPL][VEC];
#pragma unroll
        for (int k = 0; k < VPL; ++k)
#pragma unroll
            for (int i = 0; i < VEC; ++i) blk[k][i] = 0.f;
        for (int j = 0; j < cnt; j += U) {
            float x[U][VPL][VEC];
            float s[U];
            if (j + U <= cnt) {
                // full round (the common case): U unpredicated gathers, nothing to zero-fill
#pragma unroll
                for (int jj = 0; jj < U; ++jj) {
                    const uint32_t u = (uint32_t)__shfl_sync(gmask, u_rot, lane0 + jj);
                    if constexpr (HAS_SS) s[jj] = __shfl_sync(gmask, s_rot, lane0 + jj);
#pragma unroll
                    for (int k = 0; k < VPL; ++k)
                        ld_vec<VEC>(x[jj][k], reinterpret_cast<const float *>(xb[k] + (uint64_t)u * pitch));
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < U; ++jj) {
                    const uint32_t u = (uint32_t)__shfl_sync(gmask, u_rot, lane0 + jj);
                    if constexpr (HAS_SS) s[jj] = __shfl_sync(gmask, s_rot, lane0 + jj);
                    const bool ok = (j + jj) < cnt;
#pragma unroll
                    for (int k = 0; k < VPL; ++k) {
                        if (ok) {
                            ld_vec<VEC>(x[jj][k], reinterpret_cast<const float *>(xb[k] + (uint64_t)u * pitch));
                        } else {
#pragma unroll
                            for (int i = 0; i < VEC; ++i) x[jj][k][i] = 0.f;
                        }
                    }
                }
            }
            if constexpr (U < LPR) {      // next round's indices move down to lanes 0..U-1 of the group
                u_rot = __shfl_down_sync(gmask, u_rot, U, LPR);
                if constexpr (HAS_SS) s_rot = __shfl_down_sync(gmask, s_rot, U, LPR);
            }
#pragma unroll
            for (int jj = 0; jj < U; ++jj)
#pragma unroll
                for (int k = 0; k < VPL; ++k) {
                    if constexpr (VEC >= 2) {
                        // packed fp32x2 adds / fmas (sm_100): half the issue slots, each component
                        // an ordinary IEEE round-to-nearest operation — bit-identical results
#pragma unroll
                        for (int i = 0; i < VEC; i += 2) {
                            const float2 b = make_float2(blk[k][i], blk[k][i + 1]);
                            const float2 v = make_float2(x[jj][k][i], x[jj][k][i + 1]);
                            float2 r;
                            if constexpr (HAS_SS) r = __ffma2_rn(make_float2(s[jj], s[jj]), v, b);
                            else r = __fadd2_rn(b, v);
                            blk[k][i] = r.x;
                            blk[k][i + 1] = r.y;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            if constexpr (HAS_SS) blk[k][i] = fmaf(s[jj], x[jj][k][i], blk[k][i]);
                            else blk[k][i] += x[jj][k][i];
                        }
                    }
                }
        }
#pragma unroll
        for (int k = 0; k < VPL; ++k)
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[k][i] += blk[k][i];
    }
}

// EX: the forward records the dropout step it draws its mask at (once per launch), for the backward
__device__ __forceinline__ void record_drop_step(const SpmmParams &p) {
    if (p.drop.p != 0.f && p.drop.step_saved && blockIdx.x == 0 && threadIdx.x == 0)
        *p.drop.step_saved = drop_step(p.drop);
}

// EX: r <- dropout(r) for logical columns [col, col + VEC) of row v, then store r (and tf32_lo(r)).
template <int VEC>
__device__ __forceinline__ void drop_store(const SpmmParams &p, int64_t step, int v, int col, float (&r)[VEC],
                                           float *dst, float *dst_lo) {
    if (p.drop.p != 0.f) {
        uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
        uint32_t have = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const uint32_t cc = (uint32_t)(col + i);
            if ((cc >> 2) != have) {
                have = cc >> 2;
                rnd = drop_rand4(p.drop, step, (uint32_t)v, have);
            }
            r[i] = pick4(rnd, cc & 3) >= p.drop.thresh ? r[i] * p.drop.scale : 0.f;
        }
    }
    st_vec<VEC>(dst, r);
    if (dst_lo) {
        float l[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) l[i] = tf32_lo(r[i]);
        st_vec<VEC>(dst_lo, l);
    }
}

// Ragged tail (d % VEC != 0, e.g. d = 602 gathered with 128-bit loads from rows padded to 604
// floats): the last active lane of a row owns fewer than VEC valid columns.  Its GATHERS still read
// a whole vector (in bounds: every leading dimension is a multiple of VEC and >= d), its epilogue
// touches the valid columns only, one by one — pad columns are never written.
template <bool EX>
__device__ __forceinline__ void epilogue_tail_elem(const SpmmParams &p, int v, int c, float a, float t, int64_t step) {
    float r[1] = {a * t};
    if (p.addend) r[0] += __ldg(p.addend + (int64_t)v * p.ld_add + c);
    if (p.bias) r[0] += __ldg(p.bias + c);
    if (p.relu) r[0] = fmaxf(r[0], 0.f);
    if constexpr (EX) {
        drop_store<1>(p, step, v, p.col0_y + c, r, p.Y + (int64_t)v * p.ldy + c,
                      p.y_lo ? p.y_lo + (int64_t)v * p.ld_y_lo + c : nullptr);
    } else {
        p.Y[(int64_t)v * p.ldy + c] = r[0];
    }
    if (p.self_out) {
        float sx[1] = {__ldg(p.X + (int64_t)v * p.ldx + c)};
        if constexpr (EX) {
            drop_store<1>(p, step, v, p.col0_self + c, sx, p.self_out + (int64_t)v * p.ld_self + c,
                          p.self_lo ? p.self_lo + (int64_t)v * p.ld_self_lo + c : nullptr);
        } else {
            p.self_out[(int64_t)v * p.ld_self + c] = sx[0];
        }
    }
}

template <int VEC, int VPL, bool EX>
__device__ __forceinline__ void epilogue(const SpmmParams &p, int v, const int (&c)[VPL],
                                         const bool (&cv)[VPL], const float (&acc)[VPL][VEC]) {
    const float t = p.dst_scale ? __ldg(p.dst_scale + v) : 1.f;
    int64_t step = 0;
    if constexpr (EX) {
        if (p.drop.p != 0.f) step = drop_step(p.drop);
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        if (!cv[k]) continue;
        if constexpr (VEC > 1) {
            if (c[k] + VEC > p.d) {          // ragged tail of the row: valid columns only, one by one
#pragma unroll
                for (int i = 0; i < VEC - 1; ++i)
                    if (c[k] + i < p.d) epilogue_tail_elem<EX>(p, v, c[k] + i, acc[k][i], t, step);
                continue;
            }
        }
        float r[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) r[i] = acc[k][i] * t;
        if (p.addend) {
            float a[VEC];
            ld_vec<VEC>(a, p.addend + (int64_t)v * p.ld_add + c[k]);
#pragma unroll
            for (int i = 0; i < VEC; ++i) r[i] += a[i];
        }
        if (p.bias) {
            float b[VEC];
            ld_vec<VEC>(b, p.bias + c[k]);
#pragma unroll
            for (int i = 0; i < VEC; ++i) r[i] += b[i];
        }
        if (p.relu) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) r[i] = fmaxf(r[i], 0.f);
        }
        if constexpr (EX) {
            drop_store<VEC>(p, step, v, p.col0_y + c[k], r, p.Y + (int64_t)v * p.ldy + c[k],
                            p.y_lo ? p.y_lo + (int64_t)v * p.ld_y_lo + c[k] : nullptr);
        } else {
            st_vec<VEC>(p.Y + (int64_t)v * p.ldy + c[k], r);
        }
        if (p.self_out) {
            float sx[VEC];
            ld_vec<VEC>(sx, p.X + (int64_t)v * p.ldx + c[k]);
            if constexpr (EX) {
                drop_store<VEC>(p, step, v, p.col0_self + c[k], sx, p.self_out + (int64_t)v * p.ld_self + c[k],
                                p.self_lo ? p.self_lo + (int64_t)v * p.ld_self_lo + c[k] : nullptr);
            } else {
                st_vec<VEC>(p.self_out + (int64_t)v * p.ld_self + c[k], sx);
            }
        }
    }
}

// Resident CTAs per SM the register allocation must allow: the large-graph (non-COOP)
// variants are gather-latency bound and live on occupancy (40 regs -> 6 CTAs = 48 warps).
template <int VEC, int VPL, bool HAS_SS, bool COOP>
constexpr int min_ctas() {
    if (COOP) return VEC * VPL <= 2 ? 4 : (VEC * VPL <= 4 ? 3 : 2);   // cluster batches: short rows live on warps in flight too
    if (VEC * VPL <= 2) return HAS_SS ? 4 : 6;
    if (VEC * VPL <= 4) return 4;
    return 3;
}

// COOP = hub rows (more than heavy_deg edges) are split over all groups of the CTA and
// reduced through shared memory in a fixed order; used where rows are scarce.
template <int VEC, int LPR, int VPL, bool HAS_SS, bool COOP, bool EX = false>
__global__ void __launch_bounds__(256, min_ctas<VEC, VPL, HAS_SS, COOP>())
spmm_csr_kernel(const SpmmParams p) {
    constexpr int GPW = 32 / LPR;                 // row groups per warp
    constexpr int NGROUPS = 8 * GPW;              // row groups (= rows) per CTA
    constexpr int CHUNK = LPR * VEC * VPL;        // floats per feature chunk
    __shared__ int s_heavy[COOP ? NGROUPS : 1];
    __shared__ int s_nheavy;
    __shared__ float s_part[COOP ? NGROUPS : 1][COOP ? CHUNK : 1];   // 256*VEC*VPL floats = 1..8 KB

    static_assert(!EX || COOP, "the extended epilogue is launched with cooperative hub rows only");
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int grp = lane / LPR;
    const int lg = lane % LPR;
    const int gid = warp * GPW + grp;             // group index inside the CTA
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (grp * LPR));
    const int lane0 = grp * LPR;
    constexpr bool coop = COOP;
    if constexpr (EX) record_drop_step(p);
    // EX (training-step shapes): the grid may be smaller than the number of (row block, chunk)
    // items — background mode caps the CTAs resident per SM so a concurrent high-priority branch
    // always finds free slots — and each CTA then walks the items with a grid stride.
    unsigned bid = blockIdx.x;
    while (true) {
    const int chunk = bid / p.row_blocks;
    const int rb = bid - chunk * p.row_blocks;
    const int v = rb * NGROUPS + gid;

    int c[VPL];
    bool cv[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        c[k] = chunk * CHUNK + k * (LPR * VEC) + lg * VEC;
        cv[k] = c[k] < p.d;
    }
    float acc[VPL][VEC];
#pragma unroll
    for (int k = 0; k < VPL; ++k)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[k][i] = 0.f;

    if (coop) {
        if (threadIdx.x == 0) s_nheavy = 0;
        __syncthreads();
    }
    if (v < p.n_dst) {
        const int rs = __ldg(p.rowptr + v);
        const int re = __ldg(p.rowptr + v + 1);
        if (coop && re - rs > p.heavy_deg) {
            if (lg == 0) s_heavy[atomicAdd(&s_nheavy, 1)] = v;   // order irrelevant: rows are independent
        } else {
            gather_range<VEC, LPR, VPL, HAS_SS>(p, rs, re, lg, lane0, gmask, c, cv, acc);
            epilogue<VEC, VPL, EX>(p, v, c, cv, acc);
        }
    }
    if (!coop) return;
    __syncthreads();
    const int nheavy = s_nheavy;
    for (int h = 0; h < nheavy; ++h) {
        // every group of the CTA takes one contiguous, LPR-aligned slice of the hub row
        const int hv = s_heavy[h];
        const int rs = __ldg(p.rowptr + hv);
        const int re = __ldg(p.rowptr + hv + 1);
        int seg = (re - rs + NGROUPS - 1) / NGROUPS;
        seg = (seg + LPR - 1) / LPR * LPR;
        const int eb = min(re, rs + gid * seg);
        const int ee = min(re, eb + seg);
#pragma unroll
        for (int k = 0; k < VPL; ++k)
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[k][i] = 0.f;
        gather_range<VEC, LPR, VPL, HAS_SS>(p, eb, ee, lg, lane0, gmask, c, cv, acc);
#pragma unroll
        for (int k = 0; k < VPL; ++k)
#pragma unroll
            for (int i = 0; i < VEC; ++i) s_part[gid][(k * LPR + lg) * VEC + i] = acc[k][i];
        __syncthreads();
        if (gid == 0) {   // fixed summation order over the slices: deterministic
#pragma unroll
            for (int k = 0; k < VPL; ++k)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[k][i] = 0.f;
            for (int g2 = 0; g2 < NGROUPS; ++g2)
#pragma unroll
                for (int k = 0; k < VPL; ++k)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[k][i] += s_part[g2][(k * LPR + lg) * VEC + i];
            epilogue<VEC, VPL, EX>(p, hv, c, cv, acc);
        }
        __syncthreads();
    }
    if constexpr (!EX) {
        break;
    } else {
        bid += gridDim.x;
        if (bid >= (unsigned)p.total_blocks) break;
        __syncthreads();        // every thread has read s_nheavy / s_heavy of this item
    }
    }
}

// ---------------------------------------------------------------------------------------------
// Segment-balanced variant for cluster batches (rows are scarce and power-law: with one row per
// group the launch is as long as its unluckiest CTA — ncu: SMs idle 40-47 % of the kernel).
// Every row is cut into segments of <= seg_len edges (schedule built once per batch,
// gist_spmm_schedule_build); a group owns ONE (segment, feature-chunk), so all groups do the same
// amount of gathering.  A row with one segment is finished by its group.  A longer row's groups
// write their partial sums to a workspace and bump the row's arrival counter; the group that
// arrives last re-reads ALL partials and adds them in segment order — the sum does not depend on
// which group happens to be last, so the result is deterministic without a second launch or float
// atomics — then runs the epilogue and re-arms the counter.
#ifndef SEG_BATCH
#define SEG_BATCH 1    // warp items taken per queue fetch
#endif
#ifndef SEG_CTAS
#define SEG_CTAS 3     // resident CTAs per SM of the segment kernel (register budget 64K / (256 * SEG_CTAS))
#endif
template <int VEC, int LPR, int VPL, bool HAS_SS, bool EX>
__global__ void __launch_bounds__(256, SEG_CTAS) spmm_seg_kernel(const __grid_constant__ SpmmParams p) {
    constexpr int GPW = 32 / LPR;
    constexpr int CHUNK = LPR * VEC * VPL;
    const int lane = threadIdx.x & 31;
    const int grp = lane / LPR;
    const int lg = lane % LPR;
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (grp * LPR));
    const int lane0 = grp * LPR;
    pdl_sync();
    if constexpr (EX) record_drop_step(p);
    // The number of segments is known on the device only (the host sizes its buffers from the
    // batch's edge CAPACITY, several times the actual count), so the grid is a fixed number of
    // resident CTAs that pull work from a queue: a warp item is GPW consecutive segments of one
    // feature chunk — chunk-major, so the chip sweeps one column slab of X at a time.  Segments
    // differ in length (most rows of a cluster batch are one short segment), so a static
    // round-robin leaves warps idle behind the unlucky ones (ncu: 22 % of the warp slots active);
    // the queue hands the next item to whichever warp is free.  Its head is seg_count[0]: each of
    // the W warps makes exactly one failing fetch, so the fetch that returns total + W - 1 is the
    // last access of the launch and re-arms the head for the next one.
    const int n_seg = __ldg(p.seg_ptr + p.n_dst);
    const int wipc = (n_seg + GPW - 1) / GPW;              // warp items per chunk
    const int n_chunks = (p.d + CHUNK - 1) / CHUNK;
    const unsigned total = (unsigned)(wipc * n_chunks);
    const unsigned n_warps = gridDim.x * 8u;
    uint32_t *head = p.seg_count;
    uint32_t *arrivals = p.seg_count + 1;
    // Every warp's FIRST item is its own index — no fetch: at launch all W warps would otherwise
    // queue on one address (same-address atomics are served one at a time by their L2 slice) —
    // and the queue hands out items W, W + 1, ...  A fetch takes SEG_BATCH consecutive items
    // (1: measured 4 -> 40 us, 16 -> 108 us vs 23 us; the tail of a batch outweighs the atomics).
    static_assert(SEG_BATCH == 1, "one item per queue fetch (larger batches measured slower)");
    const unsigned q_total = total > n_warps ? total - n_warps : 0u;                     // items behind the queue
    unsigned work = blockIdx.x * 8u + (threadIdx.x >> 5);
    bool have = work < total;                            // fewer items than warps: straight to the failing fetch
    // GIST_SPMM_SCHED_PREFETCH: the fetch of the NEXT item is issued before the current one is processed —
    // the atomic's round trip (the queue head lives in one L2 slice) hides behind a segment of gathers
    // instead of sitting between two items; every warp still makes exactly one failing fetch
    const bool prefetch = (p.seg_flags & GIST_SPMM_SCHED_PREFETCH) != 0u;
    unsigned pre = 0u;
    if (prefetch && lane == 0) pre = atomicAdd(head, 1u);
    while (true) {
        if (!have) {
            if (!prefetch && lane == 0) pre = atomicAdd(head, 1u);
            work = __shfl_sync(0xffffffffu, pre, 0);
            if (work >= q_total) {
                // W failing fetches follow the last successful one; the last of them re-arms the head
                if (work == q_total + (n_warps - 1u) && lane == 0) *head = 0u;
                break;
            }
            work += n_warps;
            if (prefetch && lane == 0) pre = atomicAdd(head, 1u);
        }
        have = false;
        const unsigned item = work;
        const int chunk = (int)item / wipc;
        const int seg = ((int)item - chunk * wipc) * GPW + grp;
        if (seg >= n_seg) continue;                        // whole group skips together

        int c[VPL];
        bool cv[VPL];
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            c[k] = chunk * CHUNK + k * (LPR * VEC) + lg * VEC;
            cv[k] = c[k] < p.d;
        }
        float acc[VPL][VEC];
#pragma unroll
        for (int k = 0; k < VPL; ++k)
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[k][i] = 0.f;

        int v, s0, nseg, eb, ee;
        if (p.seg_meta) {       // one 16-byte record instead of seg_row -> (seg_ptr, rowptr): one round trip less per item
            const int4 mt = __ldg(p.seg_meta + seg);
            v = mt.x; eb = mt.y; s0 = mt.z;
            nseg = mt.w & 0xFFFFF;
            ee = eb + (int)((unsigned)mt.w >> 20);
        } else {
            v = __ldg(p.seg_row + seg);
            s0 = __ldg(p.seg_ptr + v);
            nseg = __ldg(p.seg_ptr + v + 1) - s0;
            const int rs = __ldg(p.rowptr + v), re = __ldg(p.rowptr + v + 1);
            eb = min(re, rs + (seg - s0) * p.seg_len);
            ee = min(re, eb + p.seg_len);
        }
        gather_range<VEC, LPR, VPL, HAS_SS>(p, eb, ee, lg, lane0, gmask, c, cv, acc);
        if (nseg == 1) {
            epilogue<VEC, VPL, EX>(p, v, c, cv, acc);
            continue;
        }
        float *mine = p.seg_ws + (int64_t)seg * p.ld_ws;
#pragma unroll
        for (int k = 0; k < VPL; ++k)
            if (cv[k]) st_vec<VEC>(mine + c[k], acc[k]);
        __threadfence();                                   // partials visible before the arrival
        __syncwarp(gmask);
        unsigned prev = 0;
        uint32_t *cnt = arrivals + (int64_t)chunk * p.n_dst + v;
        if (lg == 0) prev = atomicAdd(cnt, 1u);
        prev = __shfl_sync(gmask, prev, lane0);
        if (prev != (unsigned)(nseg - 1)) continue;
        __threadfence();
        if (lg == 0) *cnt = 0u;                            // re-armed for the next launch
#pragma unroll
        for (int k = 0; k < VPL; ++k)
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[k][i] = 0.f;
        // fixed order: segment 0, 1, 2, ...; four partial rows are requested before the first is
        // added (a hub row has tens of segments and this is the tail of the launch; a missing one
        // contributes +0, which changes nothing)
        for (int j = 0; j < nseg; j += 4) {
            float t[4][VPL][VEC];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float *part = p.seg_ws + (int64_t)(s0 + min(j + q, nseg - 1)) * p.ld_ws;
#pragma unroll
                for (int k = 0; k < VPL; ++k) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) t[q][k][i] = 0.f;
                    if (!cv[k] || j + q >= nseg) continue;
                    if constexpr (VEC == 4) {
                        const float4 w = __ldcg(reinterpret_cast<const float4 *>(part + c[k]));   // L2: other SMs wrote it
                        t[q][k][0] = w.x; t[q][k][1] = w.y; t[q][k][2] = w.z; t[q][k][3] = w.w;
                    } else if constexpr (VEC == 2) {
                        const float2 w = __ldcg(reinterpret_cast<const float2 *>(part + c[k]));
                        t[q][k][0] = w.x; t[q][k][1] = w.y;
                    } else {
                        t[q][k][0] = __ldcg(part + c[k]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int k = 0; k < VPL; ++k)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[k][i] += t[q][k][i];
        }
        epilogue<VEC, VPL, EX>(p, v, c, cv, acc);
    }
}


// ---------------------------------------------------------------------------------------------
// Slab variant for cluster batches: the gathered operand of a batch (<= ~3 k rows) is small enough
// that a COLUMN SLAB of it — SW floats of every source row — fits in one SM's shared memory.  A CTA
// (1024 threads) stages its slab once (coalesced 128-bit loads, src_scale folded in) and gathers every
// edge from shared memory instead of L2: 128 B/clk per SM against the ~50 B/clk an SM gets from L2 under
// load, and the L2 traffic of a launch drops from 4*nnz*d bytes to (CTAs per slab) x the operand.
// Grid = slabs x P parts; part p of a slab owns a contiguous range of the batch's segments (the schedule
// of spmm_seg_kernel: rows cut into <= seg_len edges), i.e. a contiguous range of the column-index
// array.  The CTA walks its range in tiles: per tile, all threads resolve the segments' metadata in
// parallel (two dependent L2 round trips for the whole tile instead of a chain per segment) and copy the
// tile's column indices into shared memory as 16-bit values (n_src < 65536); the gather loop itself then
// touches shared memory only.  A group of SW/4 lanes owns one (segment, slab): per edge one broadcast
// LDS.U16 (index) and one LDS.128 (data) per lane.  Rows of several segments are combined as in
// spmm_seg_kernel (partials to the workspace, last arriver adds them in segment order): deterministic,
// no float atomics, and independent of P.
constexpr int kSlabThreads = 1024;
constexpr int kSlabIdxCap = 16384;      // edges per tile (32 KB of 16-bit indices)
constexpr int kSlabSegCap = 768;        // segments per tile
struct SlabMeta {
    int v, eb, len, s0;                 // row, first edge (absolute), edges, first segment of the row
};

template <int SW, int OV, bool EX>
__global__ void __launch_bounds__(kSlabThreads, 1) spmm_slab_kernel(const __grid_constant__ SpmmParams p) {
    constexpr int LPR = SW / 4;                 // lanes per (segment, slab) group, one float4 each
    constexpr int NGROUPS = kSlabThreads / LPR;
    extern __shared__ float4 slab[];            // [n_src][LPR] | meta[kSlabSegCap] | nsegs[kSlabSegCap] | idx[kSlabIdxCap]
    SlabMeta *meta = reinterpret_cast<SlabMeta *>(slab + (size_t)p.n_src * LPR);
    int *nsegs = reinterpret_cast<int *>(meta + kSlabSegCap);
    uint16_t *idx = reinterpret_cast<uint16_t *>(nsegs + kSlabSegCap);
    __shared__ int s_te;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int lg = lane % LPR;
    const int gid = tid / LPR;
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << ((lane / LPR) * LPR));
    const int lane0 = (lane / LPR) * LPR;
    const int P = p.seg_blocks;
    const int slab_id = blockIdx.x / P;
    const int part = blockIdx.x - slab_id * P;
    const int c0 = slab_id * SW;
    if constexpr (EX) record_drop_step(p);

    const int n_seg = __ldg(p.seg_ptr + p.n_dst);
    const int s_begin = (int)((int64_t)n_seg * part / P);
    const int s_end = (int)((int64_t)n_seg * (part + 1) / P);

    // stage the slab.  Columns past d hold zeros (their sums are never stored).
    {
        const int nvec = p.n_src * LPR;
        const float *xb = p.X + c0;
#pragma unroll 4
        for (int i = tid; i < nvec; i += kSlabThreads) {
            const int row = i / LPR, q = i - row * LPR;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + 4 * q < p.d) v = __ldg(reinterpret_cast<const float4 *>(xb + (int64_t)row * p.ldx + 4 * q));
            if (p.src_scale) {
                const float sc = __ldg(p.src_scale + row);
                v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
            }
            slab[i] = v;
        }
    }

    int c[1];
    bool cv[1];
    c[0] = c0 + 4 * lg;
    cv[0] = c[0] < p.d;
    // the outputs may be less aligned than the staged operand (z = [h | agg] with an odd half width):
    // OV = 4 / 2 / 1 floats per store
    auto finish = [&](int v, const float (&a)[1][4]) {
        if constexpr (OV == 4) {
            epilogue<4, 1, EX>(p, v, c, cv, a);
        } else {
            constexpr int NV = 4 / OV;
            int c2[NV];
            bool cv2[NV];
            float a2[NV][OV];
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                c2[k] = c[0] + k * OV;
                cv2[k] = c2[k] < p.d;
#pragma unroll
                for (int i = 0; i < OV; ++i) a2[k][i] = a[0][k * OV + i];
            }
            epilogue<OV, NV, EX>(p, v, c2, cv2, a2);
        }
    };
    uint32_t *arrivals = p.seg_count + 1;

    for (int ts = s_begin; ts < s_end;) {
        // ---- tile metadata: every thread resolves segments of [ts, ts + kSlabSegCap) in parallel
        const int tcap = min(s_end - ts, kSlabSegCap);
        if (tid == 0) s_te = ts + tcap;
        for (int i = tid; i < tcap; i += kSlabThreads) {
            const int seg = ts + i;
            const int v = __ldg(p.seg_row + seg);
            const int s0 = __ldg(p.seg_ptr + v);
            const int s1 = __ldg(p.seg_ptr + v + 1);
            const int rs = __ldg(p.rowptr + v), re = __ldg(p.rowptr + v + 1);
            const int eb = min(re, rs + (seg - s0) * p.seg_len);
            const int ee = min(re, eb + p.seg_len);
            SlabMeta m;
            m.v = v; m.eb = eb; m.len = ee - eb; m.s0 = s0;
            meta[i] = m;
            nsegs[i] = s1 - s0;
        }
        __syncthreads();                                   // (also: the slab is staged, first time round)
        const int E0 = meta[0].eb;
        // segments are contiguous in the index array: the tile ends before the first one that does not
        // fit the index buffer (a single segment always fits: seg_len <= kSlabIdxCap)
        for (int i = tid; i < tcap; i += kSlabThreads)
            if (meta[i].eb - E0 + meta[i].len > kSlabIdxCap) atomicMin(&s_te, ts + i);
        __syncthreads();
        const int te = s_te;
        const int nt = te - ts;
        const int n_edges = meta[nt - 1].eb - E0 + meta[nt - 1].len;
        for (int e = tid; e < n_edges; e += kSlabThreads) idx[e] = (uint16_t)__ldg(p.col + E0 + e);
        __syncthreads();

        // ---- gather: shared memory only
        for (int i = gid; i < nt; i += NGROUPS) {
            const SlabMeta m = meta[i];
            const uint16_t *ix = idx + (m.eb - E0);
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f);
            int e = 0;
            for (; e + 4 <= m.len; e += 4) {               // four independent loads in flight, added in edge order
                const float4 x0 = slab[(int)ix[e] * LPR + lg];
                const float4 x1 = slab[(int)ix[e + 1] * LPR + lg];
                const float4 x2 = slab[(int)ix[e + 2] * LPR + lg];
                const float4 x3 = slab[(int)ix[e + 3] * LPR + lg];
                a0.x += x0.x; a0.y += x0.y; a0.z += x0.z; a0.w += x0.w;
                a0.x += x1.x; a0.y += x1.y; a0.z += x1.z; a0.w += x1.w;
                a0.x += x2.x; a0.y += x2.y; a0.z += x2.z; a0.w += x2.w;
                a0.x += x3.x; a0.y += x3.y; a0.z += x3.z; a0.w += x3.w;
            }
            for (; e < m.len; ++e) {
                const float4 x0 = slab[(int)ix[e] * LPR + lg];
                a0.x += x0.x; a0.y += x0.y; a0.z += x0.z; a0.w += x0.w;
            }
            float acc[1][4] = {{a0.x, a0.y, a0.z, a0.w}};
            const int v = m.v;
            const int nseg = nsegs[i];
            if (nseg == 1) {
                finish(v, acc);
                continue;
            }
            const int seg = ts + i;
            float *mine = p.seg_ws + (int64_t)seg * p.ld_ws;
            if (cv[0]) st_vec<4>(mine + c[0], acc[0]);
            __threadfence();                               // partials visible before the arrival
            __syncwarp(gmask);
            unsigned prev = 0;
            uint32_t *cnt_p = arrivals + (int64_t)slab_id * p.n_dst + v;
            if (lg == 0) prev = atomicAdd(cnt_p, 1u);
            prev = __shfl_sync(gmask, prev, lane0);
            if (prev != (unsigned)(nseg - 1)) continue;
            __threadfence();
            if (lg == 0) *cnt_p = 0u;                      // re-armed for the next launch
            acc[0][0] = acc[0][1] = acc[0][2] = acc[0][3] = 0.f;
            for (int j = 0; j < nseg; j += 4) {            // fixed order: segment 0, 1, 2, ...
                float4 t[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    t[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (cv[0] && j + q < nseg)
                        t[q] = __ldcg(reinterpret_cast<const float4 *>(p.seg_ws + (int64_t)(m.s0 + j + q) * p.ld_ws + c[0]));
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc[0][0] += t[q].x; acc[0][1] += t[q].y; acc[0][2] += t[q].z; acc[0][3] += t[q].w;
                }
            }
            finish(v, acc);
        }
        ts = te;
        if (ts < s_end) __syncthreads();                   // meta / idx are rewritten by the next tile
    }
}

constexpr size_t kSlabExtraSmem = kSlabSegCap * (sizeof(SlabMeta) + sizeof(int)) + kSlabIdxCap * sizeof(uint16_t);
constexpr size_t kSlabSmemBudget = 224 * 1024 - kSlabExtraSmem;     // bytes left for the slab itself

template <int SW, int OV>
static int launch_spmm_slab(const SpmmParams &p0, int bg, cudaStream_t stream) {
    SpmmParams p = p0;
    const int n_slabs = (p.d + SW - 1) / SW;
    int P = bg > 0 ? 1 : kNumSMs / n_slabs;       // background: one CTA per slab leaves the rest of the chip free
    if (P < 1) P = 1;
    p.seg_blocks = P;
    const size_t smem = (size_t)p.n_src * SW * sizeof(float) + kSlabExtraSmem;
    const bool ex = p.y_lo || p.self_lo || p.drop.p != 0.f;
    static bool configured[2] = {false, false};
    if (!configured[ex ? 1 : 0]) {
        cudaError_t e = ex ? cudaFuncSetAttribute(spmm_slab_kernel<SW, OV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  224 * 1024)
                           : cudaFuncSetAttribute(spmm_slab_kernel<SW, OV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  224 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured[ex ? 1 : 0] = true;
    }
    const unsigned grid = (unsigned)(n_slabs * P);
    if (ex) spmm_slab_kernel<SW, OV, true><<<grid, kSlabThreads, smem, stream>>>(p);
    else spmm_slab_kernel<SW, OV, false><<<grid, kSlabThreads, smem, stream>>>(p);
    count_launch();
    return last_error();
}

template <int VEC, int LPR, int VPL>
static int launch_spmm_seg(const SpmmParams &p0, int64_t max_segments, cudaStream_t stream) {
    SpmmParams p = p0;
    constexpr int NGROUPS = 8 * (32 / LPR);
    constexpr int CHUNK = LPR * VEC * VPL;
    constexpr int GPW = 32 / LPR;
    const int64_t max_work = ceil_div64(max_segments, GPW) * ceil_div64(p.d, CHUNK);     // warp items
    if (max_work <= 0) return GIST_OK;
    if (max_work > 0x3fffffffLL) return GIST_ERR_UNSUPPORTED;
    // resident CTAs only (SEG_CTAS per SM by the launch bounds); their warps pull from the device-side queue
    int64_t grid = (int64_t)SEG_CTAS * kNumSMs;
    if (grid > ceil_div64(max_work, 8)) grid = ceil_div64(max_work, 8);
    p.seg_blocks = 0;
    const bool ex = p.y_lo || p.self_lo || p.drop.p != 0.f;
    const dim3 g((unsigned)grid), b(256);
    cudaError_t le;
    if (p.src_scale) {
        le = ex ? launch_pdl(spmm_seg_kernel<VEC, LPR, VPL, true, true>, g, b, 0, stream, p)
                : launch_pdl(spmm_seg_kernel<VEC, LPR, VPL, true, false>, g, b, 0, stream, p);
    } else {
        le = ex ? launch_pdl(spmm_seg_kernel<VEC, LPR, VPL, false, true>, g, b, 0, stream, p)
                : launch_pdl(spmm_seg_kernel<VEC, LPR, VPL, false, false>, g, b, 0, stream, p);
    }
    count_launch();
    return le == cudaSuccess ? last_error() : (int)le;
}

template <int VEC>
static int dispatch_lanes_seg(const SpmmParams &p, int64_t max_segments, cudaStream_t stream) {
    const int lanes = (p.d + VEC - 1) / VEC;
    if (lanes <= 4) return launch_spmm_seg<VEC, 4, 1>(p, max_segments, stream);
    if (lanes <= 8) return launch_spmm_seg<VEC, 8, 1>(p, max_segments, stream);
    if (lanes <= 16) return launch_spmm_seg<VEC, 16, 1>(p, max_segments, stream);
    return launch_spmm_seg<VEC, 32, 1>(p, max_segments, stream);
}

// nseg[v] = max(1, ceil(deg(v) / seg_len)): every row gets at least one segment (its epilogue)
__global__ void seg_count_kernel(const int32_t *__restrict__ rowptr, int32_t n, int32_t seg_len,
                                 int32_t *__restrict__ nseg) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int deg = rowptr[v + 1] - rowptr[v];
    nseg[v] = max(1, (deg + seg_len - 1) / seg_len);
}

__global__ void seg_fill_kernel(const int32_t *__restrict__ seg_ptr, int32_t n, int64_t capacity,
                                int32_t *__restrict__ seg_row, const int32_t *__restrict__ rowptr, int32_t seg_len,
                                int4 *__restrict__ seg_meta) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int s0 = seg_ptr[v], s1 = seg_ptr[v + 1];
    for (int s = s0; s < s1 && s < capacity; ++s) seg_row[s] = v;
    if (seg_meta) {
        const int rs = rowptr[v], re = rowptr[v + 1];
        const int nseg = s1 - s0;
        for (int s = s0; s < s1 && s < capacity; ++s) {
            const int eb = min(re, rs + (s - s0) * seg_len);
            const int len = min(re, eb + seg_len) - eb;
            seg_meta[s] = make_int4(v, eb, s0, (int)(((unsigned)len << 20) | (unsigned)nseg));
        }
    }
}

template <int VEC, int LPR, int VPL>
static int launch_spmm(const SpmmParams &p0, cudaStream_t stream, int bg_ctas_per_sm = 0) {
    SpmmParams p = p0;
    constexpr int ROWS_PER_BLOCK = 8 * (32 / LPR);
    constexpr int CHUNK = LPR * VEC * VPL;
    const int64_t row_blocks = ceil_div64(p.n_dst, ROWS_PER_BLOCK);
    const int64_t chunks = ceil_div64(p.d, CHUNK);
    const int64_t grid = row_blocks * chunks;
    if (grid <= 0) return GIST_OK;
    if (grid > 0x7fffffffLL) return GIST_ERR_UNSUPPORTED;
    p.row_blocks = (int32_t)row_blocks;
    p.total_blocks = (int32_t)grid;
    const bool coop = p.heavy_deg != 0x7fffffff;
    if (p.y_lo || p.self_lo || p.drop.p != 0.f) {
        // extended epilogue (dropout / 3xTF32 halves): the training-step shapes only — rows are
        // scarce there, so always the CTA-cooperative variant
        if (!coop) p.heavy_deg = 128;
        int64_t g2 = grid;
        if (bg_ctas_per_sm > 0 && g2 > (int64_t)bg_ctas_per_sm * kNumSMs) g2 = (int64_t)bg_ctas_per_sm * kNumSMs;
        if (p.src_scale) spmm_csr_kernel<VEC, LPR, VPL, true, true, true><<<(unsigned)g2, 256, 0, stream>>>(p);
        else spmm_csr_kernel<VEC, LPR, VPL, false, true, true><<<(unsigned)g2, 256, 0, stream>>>(p);
        count_launch();
        return last_error();
    }
    if (p.src_scale) {
        if (coop) spmm_csr_kernel<VEC, LPR, VPL, true, true><<<(unsigned)grid, 256, 0, stream>>>(p);
        else spmm_csr_kernel<VEC, LPR, VPL, true, false><<<(unsigned)grid, 256, 0, stream>>>(p);
    } else {
        if (coop) spmm_csr_kernel<VEC, LPR, VPL, false, true><<<(unsigned)grid, 256, 0, stream>>>(p);
        else spmm_csr_kernel<VEC, LPR, VPL, false, false><<<(unsigned)grid, 256, 0, stream>>>(p);
    }
    count_launch();
    return last_error();
}

constexpr int64_t kL2SlabBytes = 64LL << 20;      // what one column slab of X may occupy of the 126 MB L2

template <int VEC>
static int dispatch_lanes(const SpmmParams &p, uint32_t flags, int32_t n_src, cudaStream_t stream) {
    const int lanes = (p.d + VEC - 1) / VEC;
    const int bg = (int)((flags >> GIST_SPMM_BG_SHIFT) & 15u);
    int force = (int)((flags >> GIST_SPMM_LANES_SHIFT) & 7u);       // 1..4 -> 4, 8, 16, 32 lanes per row
    if (force == 1) return launch_spmm<VEC, 4, 1>(p, stream, bg);
    if (force == 2) return launch_spmm<VEC, 8, 1>(p, stream, bg);
    if (force == 3) return launch_spmm<VEC, 16, 1>(p, stream, bg);
    if (force == 4) return launch_spmm<VEC, 32, 1>(p, stream, bg);
    if (lanes <= 4) return launch_spmm<VEC, 4, 1>(p, stream, bg);
    if (lanes <= 8) return launch_spmm<VEC, 8, 1>(p, stream, bg);
    if (lanes <= 16) return launch_spmm<VEC, 16, 1>(p, stream, bg);
    if (lanes <= 32) return launch_spmm<VEC, 32, 1>(p, stream, bg);
    bool wide;
    if (flags & GIST_SPMM_NARROW) wide = false;
    else if (flags & GIST_SPMM_WIDE) wide = true;
    else {
        // auto: two vectors per lane only when there are plenty of rows to fill the
        // chip AND the gathered operand is small enough that slab-by-slab L2
        // residency is not what we are after.
        const int64_t src_bytes = (int64_t)n_src * p.d * 4;
        const int64_t warps_wide = (int64_t)p.n_dst * ((lanes + 63) / 64);
        wide = lanes >= 64 && warps_wide >= 4LL * kNumSMs * 64 && src_bytes <= (64LL << 20);
    }
    return wide ? launch_spmm<VEC, 32, 2>(p, stream, bg) : launch_spmm<VEC, 32, 1>(p, stream, bg);
}

// alignment of everything the epilogue touches (outputs, addend, bias, low halves)
static bool vec_ok_out(int vec, const SpmmParams &p) {
    const size_t a = 4u * vec;
    if (!aligned(p.Y, a) || p.ldy % vec) return false;
    if (p.bias && !aligned(p.bias, a)) return false;
    if (p.addend && (!aligned(p.addend, a) || p.ld_add % vec)) return false;
    if (p.self_out && (!aligned(p.self_out, a) || p.ld_self % vec || !aligned(p.X, a) || p.ldx % vec)) return false;
    if (p.y_lo && (!aligned(p.y_lo, a) || p.ld_y_lo % vec)) return false;
    if (p.self_lo && (!aligned(p.self_lo, a) || p.ld_self_lo % vec)) return false;
    return true;
}

static bool vec_ok(int vec, const SpmmParams &p) {
    // d itself may be ragged (d % vec != 0): the row's last vector is gathered whole — in bounds,
    // because every leading dimension below is a multiple of vec and >= d — and its epilogue is
    // scalar (epilogue_tail).  What vector accesses need is alignment of every base and pitch.
    const size_t a = 4u * vec;
    if (!aligned(p.X, a) || p.ldx % vec) return false;
    if (!aligned(p.Y, a) || p.ldy % vec) return false;
    if (p.bias && !aligned(p.bias, a)) return false;
    if (p.addend && (!aligned(p.addend, a) || p.ld_add % vec)) return false;
    if (p.self_out && (!aligned(p.self_out, a) || p.ld_self % vec)) return false;
    if (p.y_lo && (!aligned(p.y_lo, a) || p.ld_y_lo % vec)) return false;
    if (p.self_lo && (!aligned(p.self_lo, a) || p.ld_self_lo % vec)) return false;
    return true;
}

__global__ void degree_norm_kernel(const int32_t *__restrict__ rowptr, int32_t n, int32_t mode,
                                   float *__restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int deg = rowptr[v + 1] - rowptr[v];
    float r;
    if (mode == GIST_NORM_INV) r = deg > 0 ? 1.0f / (float)deg : 0.f;
    else r = 1.0f / sqrtf((float)max(deg, 1));
    out[v] = r;
}

}  // namespace gist

using namespace gist;

extern "C" int gist_spmm_csr_f32(const int32_t *rowptr, const int32_t *col, int32_t n_dst,
                                 int32_t n_src, const float *X, int64_t ldx, int32_t d, float *Y,
                                 int64_t ldy, const float *src_scale, const float *dst_scale,
                                 const float *bias, const float *addend, int64_t ld_addend,
                                 float *self_out, int64_t ld_self, uint32_t flags,
                                 gist_stream_t stream) {
    return gist_spmm_csr_ex_f32(rowptr, col, n_dst, n_src, X, ldx, d, Y, ldy, src_scale, dst_scale, bias, addend,
                                ld_addend, self_out, ld_self, flags, nullptr, stream);
}

extern "C" int gist_spmm_csr_ex_f32(const int32_t *rowptr, const int32_t *col, int32_t n_dst,
                                    int32_t n_src, const float *X, int64_t ldx, int32_t d, float *Y,
                                    int64_t ldy, const float *src_scale, const float *dst_scale,
                                    const float *bias, const float *addend, int64_t ld_addend,
                                    float *self_out, int64_t ld_self, uint32_t flags,
                                    const gist_spmm_ex_t *ex, gist_stream_t stream) {
    if (n_dst < 0 || n_src < 0 || d < 0) return GIST_ERR_BADARG;
    if (n_dst == 0 || d == 0) return GIST_OK;
    if (!rowptr || !X || !Y) return GIST_ERR_BADARG;  // col may be NULL for an edgeless graph
    if (ldx < d || ldy < d) return GIST_ERR_BADARG;
    if (ldx > 0x3fffffffLL || ldy > 0x7fffffffLL || ld_addend > 0x7fffffffLL || ld_self > 0x7fffffffLL)   // ldx * 4 fits 32 bits
        return GIST_ERR_UNSUPPORTED;
    if (addend && ld_addend < d) return GIST_ERR_BADARG;
    if (self_out && (ld_self < d || n_dst != n_src)) return GIST_ERR_BADARG;
    if (!aligned(rowptr, 4) || !aligned(col, 4) || !aligned(X, 4) || !aligned(Y, 4))
        return GIST_ERR_ALIGN;
    SpmmParams p;
    p.rowptr = rowptr; p.col = col; p.n_dst = n_dst; p.n_src = n_src;
    p.X = X; p.ldx = (int32_t)ldx; p.d = d; p.Y = Y; p.ldy = (int32_t)ldy;
    p.src_scale = src_scale; p.dst_scale = dst_scale; p.bias = bias;
    p.addend = addend; p.ld_add = (int32_t)ld_addend; p.self_out = self_out; p.ld_self = (int32_t)ld_self;
    p.relu = (flags & GIST_SPMM_RELU) ? 1 : 0;
    p.row_blocks = 0;
    p.total_blocks = 0;
    p.y_lo = p.self_lo = nullptr;
    p.ld_y_lo = p.ld_self_lo = p.col0_y = p.col0_self = 0;
    make_drop_params(nullptr, &p.drop);
    if (ex) {
        if ((ex->y_lo && ex->ld_y_lo < d) || (ex->self_lo && (ex->ld_self_lo < d || !self_out))) return GIST_ERR_BADARG;
        if (ex->ld_y_lo > 0x7fffffffLL || ex->ld_self_lo > 0x7fffffffLL) return GIST_ERR_UNSUPPORTED;
        p.y_lo = ex->y_lo; p.ld_y_lo = (int32_t)ex->ld_y_lo;
        p.self_lo = ex->self_lo; p.ld_self_lo = (int32_t)ex->ld_self_lo;
        p.col0_y = ex->drop_col0_y; p.col0_self = ex->drop_col0_self;
        const int st = make_drop_params(ex->drop, &p.drop);
        if (st != GIST_OK) return st;
    }
    p.seg_ptr = p.seg_row = nullptr;
    p.seg_len = p.seg_blocks = p.ld_ws = 0;
    p.seg_count = nullptr;
    p.seg_ws = nullptr;
    p.seg_meta = nullptr;
    p.seg_flags = 0u;
    const gist_spmm_schedule_t *sch = ex ? ex->schedule : nullptr;
    if (sch) {
        if (!sch->seg_ptr || !sch->seg_row || !sch->counters || !sch->workspace || sch->seg_len <= 0 ||
            sch->max_segments < n_dst || sch->ld_workspace < d || sch->ld_workspace % 4 || !aligned(sch->workspace, 16))
            return GIST_ERR_BADARG;
        if (sch->ld_workspace > 0x7fffffffLL) return GIST_ERR_UNSUPPORTED;
        p.seg_ptr = sch->seg_ptr; p.seg_row = sch->seg_row; p.seg_len = sch->seg_len;
        p.seg_count = sch->counters; p.seg_ws = sch->workspace; p.ld_ws = (int32_t)sch->ld_workspace;
        if (sch->seg_meta && !aligned(sch->seg_meta, 16)) return GIST_ERR_ALIGN;
        p.seg_meta = reinterpret_cast<const int4 *>(sch->seg_meta);
        p.seg_flags = sch->flags;
        p.heavy_deg = 0x7fffffff;
        cudaStream_t s2 = (cudaStream_t)stream;
        // slab kernel: the whole batch's column slab in shared memory (needs 128-bit alignment everywhere
        // and n_src small enough); 16-float slabs, 8-float ones for batches up to twice as large
        if (!(flags & GIST_SPMM_SLAB_OFF) && aligned(p.X, 16) && p.ldx % 4 == 0 && n_src > 0 && n_src < 65536 &&
            sch->seg_len <= 64 && (size_t)n_src * 8 * sizeof(float) <= kSlabSmemBudget) {
            const int bg = (int)((flags >> GIST_SPMM_BG_SHIFT) & 15u);
            const bool wide = (size_t)n_src * 16 * sizeof(float) <= kSlabSmemBudget;
            if (vec_ok_out(4, p)) return wide ? launch_spmm_slab<16, 4>(p, bg, s2) : launch_spmm_slab<8, 4>(p, bg, s2);
            if (vec_ok_out(2, p)) return wide ? launch_spmm_slab<16, 2>(p, bg, s2) : launch_spmm_slab<8, 2>(p, bg, s2);
            return wide ? launch_spmm_slab<16, 1>(p, bg, s2) : launch_spmm_slab<8, 1>(p, bg, s2);
        }
        if (vec_ok(4, p)) return dispatch_lanes_seg<4>(p, sch->max_segments, s2);
        if (vec_ok(2, p)) return dispatch_lanes_seg<2>(p, sch->max_segments, s2);
        return dispatch_lanes_seg<1>(p, sch->max_segments, s2);
    }
    // Hub rows are split over the CTA only where rows are scarce (cluster batches): with
    // >32k rows there are enough warps that one long row is not the critical path.
    const bool coop = (flags & GIST_SPMM_COOP_ON) || (!(flags & GIST_SPMM_COOP_OFF) && n_dst <= 32768);
    p.heavy_deg = coop ? 128 : 0x7fffffff;
    cudaStream_t s = (cudaStream_t)stream;
    // Widest vector the alignment allows — except for operands far larger than L2 (the full-graph
    // SpMM of evaluate()): there the chip sweeps one column slab of X at a time (chunk index = slow
    // grid dimension) and a 32-lane x 128-bit chunk makes that slab n_src x 512 B.  Measured on the
    // Reddit shape (232 965 rows): d = 602 with 64-bit gathers (60 MB slab, 40 registers, 48 warps/SM)
    // 23.3 ms against 28.3 ms with 128-bit gathers (119 MB slab, 64 registers); narrower lane groups
    // (16 / 8 lanes x 128 bit: 60 / 30 MB slabs) 42.6 / 49.8 ms; d = 256: 8.2 ms against 11.6 ms.  So:
    // 64-bit gathers when the operand exceeds L2 and the 128-bit slab would not fit it but the 64-bit one
    // does.
    int max_vec = 4;
    const int force_vec = (int)((flags >> GIST_SPMM_VEC_SHIFT) & 3u);       // 1, 2, 3 -> at most 1, 2, 4 floats
    if (force_vec) max_vec = force_vec == 3 ? 4 : force_vec;
    else if ((int64_t)n_src * p.d * 4 > (96LL << 20) && p.d > 64 && (int64_t)n_src * 512 > kL2SlabBytes &&
             (int64_t)n_src * 256 <= kL2SlabBytes)
        max_vec = 2;
    if (max_vec >= 4 && vec_ok(4, p)) return dispatch_lanes<4>(p, flags, n_src, s);
    if (max_vec >= 2 && vec_ok(2, p)) return dispatch_lanes<2>(p, flags, n_src, s);
    return dispatch_lanes<1>(p, flags, n_src, s);
}

extern "C" int gist_spmm_csc_f32(const int32_t *colptr, const int32_t *row, int32_t n_src,
                                 int32_t n_dst, const float *dY, int64_t lddy, int32_t d, float *dX,
                                 int64_t lddx, const float *dst_scale, const float *src_scale,
                                 const float *addend, int64_t ld_addend, uint32_t flags,
                                 gist_stream_t stream) {
    // dX[u] = src_scale[u] * sum_{u->v} dst_scale[v] * dY[v]  (+ addend[u])
    return gist_spmm_csr_f32(colptr, row, n_src, n_dst, dY, lddy, d, dX, lddx, dst_scale, src_scale,
                             nullptr, addend, ld_addend, nullptr, 0, flags, stream);
}

extern "C" size_t gist_spmm_schedule_workspace_bytes(int32_t n) { return gist_scan_workspace_bytes(n); }

extern "C" int gist_spmm_schedule_build(const int32_t *rowptr, int32_t n, int32_t seg_len, int32_t *seg_ptr,
                                        int32_t *seg_row, int64_t max_segments, void *scan_ws,
                                        size_t scan_ws_bytes, gist_stream_t stream) {
    return gist_spmm_schedule_build_meta(rowptr, n, seg_len, seg_ptr, seg_row, nullptr, max_segments, scan_ws,
                                         scan_ws_bytes, stream);
}

extern "C" int gist_spmm_schedule_build_meta(const int32_t *rowptr, int32_t n, int32_t seg_len, int32_t *seg_ptr,
                                             int32_t *seg_row, int32_t *seg_meta, int64_t max_segments,
                                             void *scan_ws, size_t scan_ws_bytes, gist_stream_t stream) {
    if (n < 0 || seg_len <= 0 || max_segments < n) return GIST_ERR_BADARG;
    // the packed record holds <= 4095 edges and < 2^20 segments per row
    if (seg_meta && (seg_len > 4095 || max_segments >= (1LL << 20) || !aligned(seg_meta, 16))) return GIST_ERR_UNSUPPORTED;
    if (!seg_ptr) return GIST_ERR_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(seg_ptr, 0, sizeof(int32_t), s);
        return e == cudaSuccess ? GIST_OK : (int)e;
    }
    if (!rowptr || !seg_row) return GIST_ERR_BADARG;
    if (scan_ws_bytes < gist_scan_workspace_bytes(n)) return GIST_ERR_WORKSPACE;
    seg_count_kernel<<<(n + 255) / 256, 256, 0, s>>>(rowptr, n, seg_len, seg_ptr);
    count_launch();
    int st = gist_exclusive_scan_i32(seg_ptr, n, seg_ptr, scan_ws, scan_ws_bytes, stream);   // in place
    if (st != GIST_OK) return st;
    seg_fill_kernel<<<(n + 255) / 256, 256, 0, s>>>(seg_ptr, n, max_segments, seg_row, rowptr, seg_len,
                                                    reinterpret_cast<int4 *>(seg_meta));
    count_launch();
    return last_error();
}

extern "C" int gist_degree_norm_f32(const int32_t *rowptr, int32_t n, int32_t mode, float *out,
                                    gist_stream_t stream) {
    if (n < 0 || (mode != GIST_NORM_INV && mode != GIST_NORM_RSQRT_CLAMP)) return GIST_ERR_BADARG;
    if (n == 0) return GIST_OK;
    if (!rowptr || !out) return GIST_ERR_BADARG;
    degree_norm_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rowptr, n, mode, out);
    count_launch();
    return last_error();
}

extern "C" int gist_abi_version(void) { return GIST_ABI_VERSION; }

extern "C" uint64_t gist_launch_count(void) { return g_launches.load(); }

extern "C" int gist_set_pdl(int32_t enabled) {
    g_pdl.store(enabled ? 1 : 0, std::memory_order_relaxed);
    return GIST_OK;
}

extern "C" int gist_get_pdl(void) { return pdl_enabled() ? 1 : 0; }

extern "C" int gist_set_device(int device) {
    cudaError_t e = cudaSetDevice(device);
    return e == cudaSuccess ? GIST_OK : (int)e;
}

extern "C" const char *gist_status_string(int status) {
    switch (status) {
        case GIST_OK: return "ok";
        case GIST_ERR_BADARG: return "bad argument";
        case GIST_ERR_ALIGN: return "misaligned pointer";
        case GIST_ERR_WORKSPACE: return "workspace / capacity too small";
        case GIST_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown gist status";
    }
}
