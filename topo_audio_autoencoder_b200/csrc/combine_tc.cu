// S1 (b) on the 5th-generation tensor cores: the dense GEMMs of the SCCN message combine as
// tcgen05.mma kind::tf32 with 3xTF32 operand splitting (fp32-level accuracy), accumulators in tensor
// memory, one 128-row tile of target simplices per CTA iteration.  See tc.cuh for the primitives.
#include <algorithm>

#include "common.cuh"

#ifndef TOPO_DEBUG_KERNELS
#define TOPO_DEBUG_KERNELS 0
#endif
#include "tc.cuh"

namespace topo {
namespace {

using namespace tc;

constexpr int kTileRows = 128;
constexpr int kC = 64;
constexpr uint32_t kATile = kTileRows * kC * 4;   // 32 KB
constexpr uint32_t kBTile = kC * kC * 4;          // 16 KB

// [C_in x C_out] row-major weight (y = x W)  ->  B operand rows n = output column, K = input: W^T
// [C_out x C_in] row-major weight (y = x W^T, nn.Linear) -> B operand as stored
__device__ __forceinline__ void load_weight_tiles(const float* __restrict__ w, bool transpose, uint8_t* hi, uint8_t* lo,
                                                  int tid, int nthreads) {
    for (int idx = tid; idx < kC * (kC / 4); idx += nthreads) {
        const int n = idx / (kC / 4), chunk = idx % (kC / 4);
        float4 v;
        if (transpose) {
            v.x = __ldg(w + (chunk * 4 + 0) * kC + n);
            v.y = __ldg(w + (chunk * 4 + 1) * kC + n);
            v.z = __ldg(w + (chunk * 4 + 2) * kC + n);
            v.w = __ldg(w + (chunk * 4 + 3) * kC + n);
        } else {
            v = __ldg(reinterpret_cast<const float4*>(w + n * kC) + chunk);
        }
        store_split(hi, lo, kC, n, chunk, v);
    }
}

// One 3xTF32 product with explicit operand descriptors: `steps` MMAs of K = 8 each per pass; operand k-step
// offsets and LBO/SBO are given by the caller (K-major and MN-major operands differ only there).
struct OperandWalk {
    uint32_t hi, lo;        // shared-memory addresses of the hi / lo tile images
    uint32_t step_bytes;    // advance per k-step inside a group of `steps_per_jump`
    uint32_t jump_bytes;    // advance per group
    uint32_t steps_per_jump;
    uint32_t lbo, sbo;
};
__device__ __forceinline__ void gemm_3xtf32(uint32_t tmem_d, const OperandWalk& a, const OperandWalk& b, uint32_t idesc,
                                            int steps, uint32_t accumulate_into) {
    uint32_t acc = accumulate_into;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a_base = pass == 0 ? a.lo : a.hi;
        const uint32_t b_base = pass == 1 ? b.lo : b.hi;
#pragma unroll 1
        for (int k = 0; k < steps; ++k) {
            const uint32_t ao = (k / a.steps_per_jump) * a.jump_bytes + (k % a.steps_per_jump) * a.step_bytes;
            const uint32_t bo = (k / b.steps_per_jump) * b.jump_bytes + (k % b.steps_per_jump) * b.step_bytes;
            mma_tf32(tmem_d, smem_desc_sw128(a_base + ao, a.lbo, a.sbo), smem_desc_sw128(b_base + bo, b.lbo, b.sbo), idesc, acc);
            acc = 1;
        }
    }
}
// A [rows x 64] tile image contracted over its 64 columns (K-major): 4 steps of 32 B per 128-B atom, 2 atoms.
__device__ __forceinline__ OperandWalk walk_k_major(uint32_t hi, uint32_t lo, int rows) {
    return OperandWalk{hi, lo, 32u, static_cast<uint32_t>(rows) * 128u, 4u, 16u, 1024u};
}
// Debug / unit-test entry for the tcgen05 path: out[rows, 64] = a[rows, 64] @ w[64, 64].
//   mode 0: both operands from shared memory (K-major SWIZZLE_128B tiles)
//   mode 3: the A operand from tensor memory (written by the row threads with tcgen05.st)
// (MN-major fp32 operands need the SWIZZLE_128B_BASE32B layout, a different shared-memory image; the
//  weight-gradient products that would use them stay on the FFMA path for now.)
__global__ void __launch_bounds__(128) debug_gemm_kernel(const float* __restrict__ a, const float* __restrict__ w,
                                                         long long rows, int mode, float* __restrict__ out) {
    // 1024-byte aligned by declaration (SWIZZLE_128B atoms); no pointer arithmetic through integers, so the
    // compiler keeps every access in the shared address space (LDS/STS, not generic LD/ST)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw;
    uint8_t* a_hi = base;
    uint8_t* a_lo = a_hi + kATile;
    uint8_t* b_hi = a_lo + kATile;
    uint8_t* b_lo = b_hi + kBTile;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_smem, 256);
    load_weight_tiles(w, true, b_hi, b_lo, tid, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_smem;

    uint32_t parity = 0;
    const long long tiles = (rows + kTileRows - 1) / kTileRows;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row0 = tile * kTileRows;
        // 16 threads per row, one 16-byte chunk each: coalesced 256-byte rows
        for (int idx = tid; idx < kTileRows * 16; idx += 128) {
            const int r = idx >> 4, chunk = idx & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < rows) v = __ldg(reinterpret_cast<const float4*>(a + (row0 + r) * kC) + chunk);
            store_split(a_hi, a_lo, kTileRows, r, chunk, v);
        }
        if (mode == 3) {
            // A operand in tensor memory: thread t owns lane t; hi in columns [64,128), lo in [128,192)
            const long long arow = row0 + tid;
            for (int c8 = 0; c8 < 8; ++c8) {
                float hi8[8], lo8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float x = arow < rows ? __ldg(a + arow * kC + c8 * 8 + i) : 0.f;
                    hi8[i] = tf32_hi(x);
                    lo8[i] = x - hi8[i];
                }
                const uint32_t lane = tmem_base_smem + (static_cast<uint32_t>(warp * 32) << 16);
                tmem_st8(lane + 64 + c8 * 8, hi8);
                tmem_st8(lane + 128 + c8 * 8, lo8);
            }
            tmem_st_wait();
            tc_fence_before_sync();
        }
        fence_async_shared();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
            if (mode == 3) {
                uint32_t acc = 0;
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a_col = pass == 0 ? 128 : 64;
                    const uint32_t bb = smem_u32(pass == 1 ? b_lo : b_hi);
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t bo = (k >> 2) * (kC * 128) + (k & 3) * 32;
                        mma_tf32_ta(tmem_base, tmem_base + a_col + k * 8, smem_desc_sw128(bb + bo, 16, 1024),
                                    idesc_tf32(128, 64, 0, 0), acc);
                        acc = 1;
                    }
                }
            } else {
                gemm_3xtf32(tmem_base, walk_k_major(smem_u32(a_hi), smem_u32(a_lo), kTileRows),
                            walk_k_major(smem_u32(b_hi), smem_u32(b_lo), kC), idesc_tf32(128, 64, 0, 0), 8, 0);
            }
            mma_commit(&bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
        tc_fence_after_sync();
        const long long row = row0 + tid;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float v[32];
            tmem_ld32(taddr + half * 32, v);
            if (row < rows) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(out + row * kC + half * 32 + q * 4) =
                        make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            }
        }
        tc_fence_before_sync();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}


// ---------------------------------------------------------------------------------------------
// Fused combine forward on tensor cores (C = 64), 128-row tiles, 256 threads.
// Thread (h, r) = (tid / 128, tid % 128) owns columns [32h, 32h + 32) of tile row r (TMEM lane r; both
// warp groups reach the same lane quarter, warp % 4).  Everybody loads, everybody runs epilogues; thread 0
// issues the MMAs.  Tensor memory: T_k (conv outputs, 3 x 64 columns, kept for the final mix) and H
// (attention hidden layer, 64 columns).  Per message k:
//     stage A = agg_k (prefetched one message ahead into registers, split hi/lo, swizzled)
//     MMA1  T_k = agg_k W_k                    (24 x tcgen05.mma 128x64x8, 3xTF32)
//     epi1  m_k = scale_k T_k + x  -> A operand (TMEM -> registers -> smem)
//     MMA2  H   = m_k W1^T
//     epi2  s_k = w2 . GELU(H + b1) + b2       (half-row partials combined through shared memory)
// then  out = (sum a_k) x + sum_k a_k scale_k T_k,  LayerNorm,  store.
// The message loop and the 8-column epilogue chunks are real loops: the fully unrolled first version
// spent 19-35 % of its issue slots on instruction-cache misses (profiles/r01_combine_fwd_tc_v1.md).
// ---------------------------------------------------------------------------------------------
struct FwdSmem {
    static constexpr uint32_t kWeights = 0;                       // [4][hi, lo] x 16 KB
    static constexpr uint32_t kA = 8 * kBTile;                    // hi, lo x 32 KB
    static constexpr uint32_t kVecs = kA + 2 * kATile;            // b1, w2, gamma, beta
    static constexpr uint32_t kRed = kVecs + 4 * kC * 4;          // [NQ <= 4][128] partial row sums
    static constexpr uint32_t kScale = kRed + 4 * kTileRows * 4;  // scale[3], b2
    static constexpr uint32_t kTotal = kScale + 16;
};

// NQ column groups per row: thread (q, r) = (tid / 128, tid % 128) owns CW = 64 / NQ columns of tile row r.
// NQ = 4 (512 threads, 16 warps) doubles the warps that hide each other's epilogue latencies.
template <int NQ>
__global__ void __launch_bounds__(128 * NQ, 1) combine_fwd_tc_kernel(topo_combine_params P, long long rows,
                                                                     const int* __restrict__ n_rows_dev,
                                                                     float* __restrict__ out) {
    constexpr int NT_ = 128 * NQ;          // threads
    constexpr int CW = kC / NQ;            // columns per thread
    constexpr int NCH = CW / 8;            // 8-column TMEM chunks per thread
    constexpr int NV = CW / 4;             // float4 per thread row slice
    constexpr int PF = 2048 / NT_;         // prefetched 16-byte chunks of a 128 x 64 tile per thread
    // 1024-byte aligned by declaration (SWIZZLE_128B atoms); no pointer arithmetic through integers, so the
    // compiler keeps every access in the shared address space (LDS/STS, not generic LD/ST)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw;
    uint8_t* a_hi = base + FwdSmem::kA;
    uint8_t* a_lo = a_hi + kATile;
    float* vecs = reinterpret_cast<float*>(base + FwdSmem::kVecs);
    float* red = reinterpret_cast<float*>(base + FwdSmem::kRed);
    float* scale_s = reinterpret_cast<float*>(base + FwdSmem::kScale);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;

    const long long live = n_rows_dev ? min(static_cast<long long>(*n_rows_dev), rows) : rows;
    const long long tiles = (live + kTileRows - 1) / kTileRows;
    if (blockIdx.x >= tiles) return;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int h = tid >> 7, r = tid & 127, col0 = h * CW;
    const int n_msgs = P.n_msgs;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_smem, 256);
    for (int k = 0; k < n_msgs; ++k)
        load_weight_tiles(P.w[k], true, base + (2 * k) * kBTile, base + (2 * k + 1) * kBTile, tid, NT_);
    load_weight_tiles(P.att_w1, false, base + 6 * kBTile, base + 7 * kBTile, tid, NT_);
    for (int c = tid; c < kC; c += NT_) {
        vecs[c] = __ldg(P.att_b1 + c);
        vecs[kC + c] = __ldg(P.att_w2 + c);
        vecs[2 * kC + c] = P.apply_ln ? __ldg(P.ln_gamma + c) : 1.f;
        vecs[3 * kC + c] = P.apply_ln ? __ldg(P.ln_beta + c) : 0.f;
    }
    if (tid < 3) scale_s[tid] = tid < n_msgs ? __ldg(P.scale[tid]) : 0.f;
    if (tid == 3) scale_s[3] = __ldg(P.att_b2);
    fence_async_shared();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + col0;
    const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), w_s = smem_u32(base);
    const float b2 = scale_s[3];
    uint32_t parity = 0;

    float4 pre[PF];     // the next aggregate tile, PF of its 2048 16-byte chunks per thread
    auto prefetch = [&](const float* __restrict__ src, long long row0) {
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            const int idx = j * NT_ + tid, rr = idx >> 4, chunk = idx & 15;
            pre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + rr < live) pre[j] = __ldg(reinterpret_cast<const float4*>(src + (row0 + rr) * kC) + chunk);
        }
    };
    prefetch(P.agg[0], static_cast<long long>(blockIdx.x) * kTileRows);

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row0 = tile * kTileRows;
        const long long row = row0 + r;
        float4 xr[NV];  // this thread's slice of the residual row; later the output accumulator
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            xr[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (P.x != nullptr && row < live) xr[q] = __ldg(reinterpret_cast<const float4*>(P.x + row * kC + col0) + q);
        }
        float sc0 = 0.f, sc1 = 0.f, sc2 = 0.f;
#pragma unroll 1
        for (int k = 0; k < n_msgs; ++k) {
            const float scale_k = scale_s[k];
#pragma unroll
            for (int j = 0; j < PF; ++j) {
                const int idx = j * NT_ + tid;
                store_split(a_hi, a_lo, kTileRows, idx >> 4, idx & 15, pre[j]);
            }
            fence_async_shared();
            if (k + 1 < n_msgs) prefetch(P.agg[k + 1], row0);
            else if (tile + gridDim.x < tiles) prefetch(P.agg[0], (tile + gridDim.x) * kTileRows);
            __syncthreads();                                                    // S1: A = agg_k is staged
            if (warp == 0) {
                if (tid == 0) {
                    tc_fence_after_sync();
                    gemm_128x64x64_3xtf32(tmem_base + k * kC, a_hi_s, a_lo_s, w_s + (2 * k) * kBTile, w_s + (2 * k + 1) * kBTile, 0);
                    mma_commit(&bar);
                }
                __syncwarp();      // the other lanes park here instead of polling against the issuing lane
            }
            mbar_wait_backoff(&bar, parity);
            parity ^= 1;
            tc_fence_after_sync();
            // epilogue 1: m_k = scale_k T_k + x, re-staged as the A operand of the attention GEMM
#pragma unroll
            for (int c8 = 0; c8 < NCH; ++c8) {
                float v[8];
                tmem_ld8(lane_addr + k * kC + c8 * 8, v);
                const float4 x0 = xr[c8 * 2], x1 = xr[c8 * 2 + 1];
                const float4 m0 = make_float4(fmaf(scale_k, v[0], x0.x), fmaf(scale_k, v[1], x0.y), fmaf(scale_k, v[2], x0.z),
                                              fmaf(scale_k, v[3], x0.w));
                const float4 m1 = make_float4(fmaf(scale_k, v[4], x1.x), fmaf(scale_k, v[5], x1.y), fmaf(scale_k, v[6], x1.z),
                                              fmaf(scale_k, v[7], x1.w));
                store_split(a_hi, a_lo, kTileRows, r, h * NV + c8 * 2, m0);
                store_split(a_hi, a_lo, kTileRows, r, h * NV + c8 * 2 + 1, m1);
                if (P.saved_m[k] != nullptr && row < live) {
                    float4* dst = reinterpret_cast<float4*>(P.saved_m[k] + row * kC + col0) + c8 * 2;
                    dst[0] = m0;
                    dst[1] = m1;
                }
            }
            fence_async_shared();
            tc_fence_before_sync();
            __syncthreads();                                                    // S2: A = m_k is staged
            if (warp == 0) {
                if (tid == 0) {
                    tc_fence_after_sync();
                    gemm_128x64x64_3xtf32(tmem_base + 3 * kC, a_hi_s, a_lo_s, w_s + 6 * kBTile, w_s + 7 * kBTile, 0);
                    mma_commit(&bar);
                }
                __syncwarp();
            }
            mbar_wait_backoff(&bar, parity);
            parity ^= 1;
            tc_fence_after_sync();
            // epilogue 2: this half-row's share of the attention score
            float s = 0.f;
#pragma unroll 1
            for (int c8 = 0; c8 < NCH; ++c8) {
                float v[8];
                tmem_ld8(lane_addr + 3 * kC + c8 * 8, v);
                const float* b1p = vecs + col0 + c8 * 8;
                const float* w2p = vecs + kC + col0 + c8 * 8;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    v[i] += b1p[i];                                   // pre-GELU hidden activation
                    s = fmaf(gelu_exact(v[i]), w2p[i], s);
                }
                if (P.saved_pre[k] != nullptr && row < live) {
                    float4* dst = reinterpret_cast<float4*>(P.saved_pre[k] + row * kC + col0) + c8 * 2;
                    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
            red[h * kTileRows + r] = s;
            tc_fence_before_sync();
            __syncthreads();                                                    // S3: A, H and `red` are consistent
            float score = b2;
#pragma unroll
            for (int g = 0; g < NQ; ++g) score += red[g * kTileRows + r];
            if (k == 0) sc0 = score; else if (k == 1) sc1 = score; else sc2 = score;
            if (P.saved_score != nullptr && h == 0 && row < live) P.saved_score[k * rows + row] = score;
        }
        // softmax over the messages, mix straight from tensor memory
        float mx = sc0;
        if (n_msgs > 1) mx = fmaxf(mx, sc1);
        if (n_msgs > 2) mx = fmaxf(mx, sc2);
        const float e0 = expf(sc0 - mx), e1 = n_msgs > 1 ? expf(sc1 - mx) : 0.f, e2 = n_msgs > 2 ? expf(sc2 - mx) : 0.f;
        const float esum = e0 + e1 + e2;
        const float a0 = e0 / esum, a1 = e1 / esum, a2 = e2 / esum;
        const float asum = a0 + a1 + a2;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            xr[q].x *= asum; xr[q].y *= asum; xr[q].z *= asum; xr[q].w *= asum;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < n_msgs) {
                const float coef = (k == 0 ? a0 : (k == 1 ? a1 : a2)) * scale_s[k];
#pragma unroll
                for (int c8 = 0; c8 < NCH; ++c8) {
                    float v[8];
                    tmem_ld8(lane_addr + k * kC + c8 * 8, v);
                    float4& o0 = xr[c8 * 2];
                    float4& o1 = xr[c8 * 2 + 1];
                    o0.x = fmaf(coef, v[0], o0.x); o0.y = fmaf(coef, v[1], o0.y); o0.z = fmaf(coef, v[2], o0.z); o0.w = fmaf(coef, v[3], o0.w);
                    o1.x = fmaf(coef, v[4], o1.x); o1.y = fmaf(coef, v[5], o1.y); o1.z = fmaf(coef, v[6], o1.z); o1.w = fmaf(coef, v[7], o1.w);
                }
            }
        }
        tc_fence_before_sync();
        if (P.apply_ln) {
            float part = 0.f;
#pragma unroll
            for (int q = 0; q < NV; ++q) part += (xr[q].x + xr[q].y) + (xr[q].z + xr[q].w);
            __syncthreads();                       // every thread has consumed the last score from `red`
            red[h * kTileRows + r] = part;
            __syncthreads();
            float msum = 0.f;
#pragma unroll
            for (int g = 0; g < NQ; ++g) msum += red[g * kTileRows + r];
            const float mean = msum * (1.0f / kC);
            float var = 0.f;
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                var = fmaf(xr[q].x - mean, xr[q].x - mean, var);
                var = fmaf(xr[q].y - mean, xr[q].y - mean, var);
                var = fmaf(xr[q].z - mean, xr[q].z - mean, var);
                var = fmaf(xr[q].w - mean, xr[q].w - mean, var);
            }
            __syncthreads();
            red[h * kTileRows + r] = var;
            __syncthreads();
            float vsum = 0.f;
#pragma unroll
            for (int g = 0; g < NQ; ++g) vsum += red[g * kTileRows + r];
            const float rstd = 1.0f / sqrtf(vsum * (1.0f / kC) + P.ln_eps);
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const float* gm = vecs + 2 * kC + col0 + q * 4;
                const float* bt = vecs + 3 * kC + col0 + q * 4;
                xr[q].x = fmaf((xr[q].x - mean) * rstd, gm[0], bt[0]);
                xr[q].y = fmaf((xr[q].y - mean) * rstd, gm[1], bt[1]);
                xr[q].z = fmaf((xr[q].z - mean) * rstd, gm[2], bt[2]);
                xr[q].w = fmaf((xr[q].w - mean) * rstd, gm[3], bt[3]);
            }
        }
        if (row < live) {
#pragma unroll
            for (int q = 0; q < NV; ++q) *(reinterpret_cast<float4*>(out + row * kC + col0) + q) = xr[q];
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}


// ---------------------------------------------------------------------------------------------
// Backward through the conv weight on tensor cores, one message per blockIdx.y (C = 64):
//     dL/dagg_k = scale_k (dL/dm_k) W_k^T                      A = dm tile (K-major), B = W_k as stored [in][out]
//     g_wprod[k] += agg_k^T (dL/dm_k)                          contraction over the tile ROWS
// The row contraction is a K-major product of TRANSPOSED tiles: A^T = agg_k^T [64 x rows], B^T = dm^T
// [64 x rows], staged 64 rows at a time (lanes run along the rows, so every 128-byte line of the transposed
// tile is written by one warp, conflict free).  The MMA is issued with M = 128; rows 64..127 of its A operand
// alias whatever follows the 64-row image and produce accumulator rows that are never read.  The 64 x 64
// accumulator stays in tensor memory for the whole CTA and is added to global memory once at the end.
// ---------------------------------------------------------------------------------------------
struct ConvSmem {
    static constexpr uint32_t kW = 0;                              // W_k [in][out] hi, lo (2 x 16 KB)
    static constexpr uint32_t kA = 2 * kBTile;                     // dm tile hi, lo (2 x 32 KB)
    static constexpr uint32_t kAT = kA + 2 * kATile;               // agg^T chunk [64 x 64] hi, lo (2 x 16 KB)
    static constexpr uint32_t kBT = kAT + 2 * kBTile;              // dm^T chunk  [64 x 64] hi, lo (2 x 16 KB)
    static constexpr uint32_t kTotal = kBT + 2 * kBTile + 16 * 1024;   // slack: the M = 128 over-read stays in bounds
};

// Transposed staging of rows [row_lo, row_lo + 64) of a [rows, 64] global matrix: element (r, c) -> tile row c,
// K position r - row_lo.  Lane group = 64 consecutive rows; the NG = threads / 64 groups share the 16 column chunks.
template <int NG>
__device__ __forceinline__ void stage_transposed(const float* __restrict__ src, long long row_lo, long long live,
                                                 uint8_t* hi, uint8_t* lo, int tid) {
    const int rl = tid & 63, grp = tid >> 6;
    const long long row = row_lo + rl;
#pragma unroll
    for (int q = 0; q < 16 / NG; ++q) {
        const int chunk = grp * (16 / NG) + q;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < live) v = __ldg(reinterpret_cast<const float4*>(src + row * kC) + chunk);
        const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int trow = chunk * 4 + j;                       // tile row = source column
            const uint32_t off = tile_chunk_offset(kC, trow, rl >> 2) + (rl & 3) * 4;
            const float h = tf32_hi(vals[j]);
            *reinterpret_cast<float*>(hi + off) = h;
            *reinterpret_cast<float*>(lo + off) = vals[j] - h;
        }
    }
}

template <int NQ>
__global__ void __launch_bounds__(128 * NQ, 1) combine_bwd_conv_tc_kernel(topo_combine_params P, long long rows,
                                                                          const int* __restrict__ n_rows_dev,
                                                                          topo_combine_grads G, const float* __restrict__ dm_ws) {
    constexpr int NT_ = 128 * NQ, CW = kC / NQ, NCH = CW / 8, PF = 2048 / NT_;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw;
    uint8_t* w_hi = base + ConvSmem::kW;
    uint8_t* w_lo = w_hi + kBTile;
    uint8_t* a_hi = base + ConvSmem::kA;
    uint8_t* a_lo = a_hi + kATile;
    uint8_t* at_hi = base + ConvSmem::kAT;
    uint8_t* at_lo = at_hi + kBTile;
    uint8_t* bt_hi = base + ConvSmem::kBT;
    uint8_t* bt_lo = bt_hi + kBTile;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;

    const int k = blockIdx.y;
    const long long live = n_rows_dev ? min(static_cast<long long>(*n_rows_dev), rows) : rows;
    const long long tiles = (live + kTileRows - 1) / kTileRows;
    if (blockIdx.x >= tiles) return;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int h = tid >> 7, r = tid & 127, col0 = h * CW;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_smem, 128);
    load_weight_tiles(P.w[k], false, w_hi, w_lo, tid, NT_);      // B rows n = in, K = out: W_k as stored
    const float scale = __ldg(P.scale[k]);
    const float* __restrict__ agg = P.agg[k];
    const float* __restrict__ dm = dm_ws + static_cast<long long>(k) * rows * kC;
    float* __restrict__ g_agg = G.g_agg[k];
    fence_async_shared();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t tmem_dg = tmem_base, tmem_dp = tmem_base + kC;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t parity = 0, dp_started = 0;

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row0 = tile * kTileRows;
        // dm tile, K-major (coalesced rows)
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            const int idx = j * NT_ + tid, rr = idx >> 4, chunk = idx & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + rr < live) v = __ldg(reinterpret_cast<const float4*>(dm + (row0 + rr) * kC) + chunk);
            store_split(a_hi, a_lo, kTileRows, rr, chunk, v);
        }
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            stage_transposed<NT_ / 64>(agg, row0 + half * 64, live, at_hi, at_lo, tid);
            stage_transposed<NT_ / 64>(dm, row0 + half * 64, live, bt_hi, bt_lo, tid);
            fence_async_shared();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after_sync();
                if (half == 0)
                    gemm_3xtf32(tmem_dg, walk_k_major(smem_u32(a_hi), smem_u32(a_lo), kTileRows),
                                walk_k_major(smem_u32(w_hi), smem_u32(w_lo), kC), idesc_tf32(128, 64, 0, 0), 8, 0);
                gemm_3xtf32(tmem_dp, walk_k_major(smem_u32(at_hi), smem_u32(at_lo), kC),
                            walk_k_major(smem_u32(bt_hi), smem_u32(bt_lo), kC), idesc_tf32(128, 64, 0, 0), 8, dp_started);
                mma_commit(&bar);
            }
            dp_started = 1;
            mbar_wait(&bar, parity);
            parity ^= 1;
            tc_fence_after_sync();
            if (half == 0) {
                // dL/dagg_k = scale_k * DG: thread (h, r) writes its 32 columns of row r
                const long long row = row0 + r;
#pragma unroll
                for (int c8 = 0; c8 < NCH; ++c8) {
                    float v[8];
                    tmem_ld8(lane_addr + col0 + c8 * 8, v);
                    if (row < live) {
                        float4* dst = reinterpret_cast<float4*>(g_agg + row * kC + col0) + c8 * 2;
                        dst[0] = make_float4(scale * v[0], scale * v[1], scale * v[2], scale * v[3]);
                        dst[1] = make_float4(scale * v[4], scale * v[5], scale * v[6], scale * v[7]);
                    }
                }
            }
            tc_fence_before_sync();
            __syncthreads();          // the transposed chunks (and, after half 1, the dm tile) may be overwritten
        }
    }
    // the weight-gradient product of this CTA: rows 0..63 of DP (lanes 0..63), all 64 columns
    if (tid < 64) {
        float* dst = G.g_wprod[k] + tid * kC;
#pragma unroll 1
        for (int c8 = 0; c8 < 8; ++c8) {
            float v[8];
            tmem_ld8(lane_addr + kC + c8 * 8, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) atomicAdd(dst + c8 * 8 + i, v[i]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

}  // namespace
}  // namespace topo

using namespace topo;

#if TOPO_DEBUG_KERNELS      // unit-test GEMM: only in libtopo_b200_debug.so
extern "C" int topo_debug_gemm_tf32x3(const float* a, const float* w, int64_t rows, int mode, float* out,
                                      topo_stream_t stream) {
    TOPO_REQUIRE(a && w && out && rows >= 0, "bad argument");
    TOPO_REQUIRE(mode == 0 || mode == 3, "mode must be 0 (operands in shared memory) or 3 (A in tensor memory)");
    if (rows == 0) return TOPO_OK;
    const size_t smem = 2 * kATile + 2 * kBTile;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(debug_gemm_kernel), smem)) return rc;
    const int tiles = static_cast<int>((rows + kTileRows - 1) / kTileRows);
    debug_gemm_kernel<<<std::min(tiles, sm_count()), 128, smem, as_stream(stream)>>>(a, w, rows, mode, out);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
#endif  // TOPO_DEBUG_KERNELS

extern "C" int topo_sccn_combine_fwd_tc(const topo_combine_params* p, int64_t rows, const int32_t* n_rows_dev,
                                        float* out, topo_stream_t stream) {
    TOPO_REQUIRE(p && out && rows >= 0, "bad argument");
    TOPO_REQUIRE(p->n_msgs >= 1 && p->n_msgs <= 3, "n_msgs must be 1..3");
    if (p->channels != kC) {
        set_error("the tensor-core combine is instantiated for channels == 64");
        return TOPO_ERR_UNSUPPORTED;
    }
    for (int k = 0; k < p->n_msgs; ++k) TOPO_REQUIRE(p->agg[k] && p->w[k] && p->scale[k], "null message operand");
    TOPO_REQUIRE(p->att_w1 && p->att_b1 && p->att_w2 && p->att_b2, "null attention parameter");
    TOPO_REQUIRE(!p->apply_ln || (p->ln_gamma && p->ln_beta), "null LayerNorm parameter");
    if (rows == 0) return TOPO_OK;
    const size_t smem = FwdSmem::kTotal + 1024;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(combine_fwd_tc_kernel<4>), smem)) return rc;
    const int tiles = static_cast<int>((rows + kTileRows - 1) / kTileRows);
    combine_fwd_tc_kernel<4><<<std::min(tiles, p->max_ctas > 0 ? std::min(p->max_ctas, sm_count()) : sm_count()), 512, smem, as_stream(stream)>>>(*p, rows, n_rows_dev, out);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_sccn_combine_bwd_conv_tc(const topo_combine_params* p, int64_t rows, const int32_t* n_rows_dev,
                                             const topo_combine_grads* g, const float* workspace,
                                             topo_stream_t stream) {
    TOPO_REQUIRE(p && g && workspace && rows >= 0, "bad argument");
    TOPO_REQUIRE(p->n_msgs >= 1 && p->n_msgs <= 3, "n_msgs must be 1..3");
    if (p->channels != kC) {
        set_error("the tensor-core combine is instantiated for channels == 64");
        return TOPO_ERR_UNSUPPORTED;
    }
    for (int k = 0; k < p->n_msgs; ++k)
        TOPO_REQUIRE(p->agg[k] && p->w[k] && p->scale[k] && g->g_agg[k] && g->g_wprod[k], "null message operand");
    if (rows == 0) return TOPO_OK;
    const size_t smem = ConvSmem::kTotal;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(combine_bwd_conv_tc_kernel<4>), smem)) return rc;
    const int tiles = static_cast<int>((rows + kTileRows - 1) / kTileRows);
    const int per_msg = std::max(1, std::min(tiles, sm_count() / p->n_msgs));
    combine_bwd_conv_tc_kernel<4><<<dim3(per_msg, p->n_msgs), 512, smem, as_stream(stream)>>>(*p, rows, n_rows_dev, *g, workspace);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
