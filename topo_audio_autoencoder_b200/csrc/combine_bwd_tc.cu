// S1 (b) backward, fused, on the 5th-generation tensor cores (C = 64): LayerNorm / softmax / attention-MLP
// backward, the input gradients and all four weight-gradient products of one rank in ONE kernel, bf16x3
// operand images (tc16.cuh), accumulators in tensor memory.
//
// Per 128-row tile, thread (q, r) = (tid / 128, tid % 128) owns columns [16 q, 16 q + 16) of tile row r
// (TMEM lane r).  With the forward's saved m_k, pre_k = W1 m_k + b1 and score_k:
//   phase A   a = softmax_k(score), y = sum a_k m_k, LayerNorm backward -> dy, dscore_k = a_k (dy.m_k - sum_j a_j dy.m_j)
//   per message k, two MMA rounds on two operand images P and Q (48 KB each, staged once, read in both majors):
//     P = dpre_k = dscore_k w2 GELU'(pre_k),  Q = m_k
//       round 1:  T    = P  W1          (P K-major,  W1 [o][i] MN-major)
//                 DW1 += P^T Q          (both MN-major: contraction over the tile rows)
//     P = dm_k = a_k dy + T,  Q = agg_k
//       round 2:  Ga   = P  W_k^T       (P K-major,  W_k [in][out] K-major)        -> g_agg_k = scale_k Ga
//                 DWp_k += Q^T P        (both MN-major)                             -> g_wprod[k]
//   g_x = sum_k dm_k.  The GELU' of message k + 1 is evaluated while round 2 of message k runs.
// DW1 and DWp_k (64 x 64 each) stay in tensor memory for the whole CTA and are added to global memory once.
// The row-contraction MMAs are issued with M = 128: rows 64..127 of their A operand alias the next 16 KB of
// shared memory and produce accumulator rows that are never read.
// Column sums (b1, w2, gamma, beta gradients) are reduced over a warp's 32 rows with a halving shuffle
// exchange (16 shuffles per 16 columns), one register per quantity per thread.
#include <algorithm>

#include "common.cuh"
#include "tc16.cuh"

namespace topo {
namespace {

using namespace tc16;

constexpr int kTileRows = 128;
constexpr int kC = 64;
constexpr int kThreads = 512;
constexpr int kCW = 16;                                   // columns per thread
constexpr uint32_t kPart = kTileRows * 128;               // 16 KB
constexpr uint32_t kImg = 3 * kPart;                      // 48 KB
constexpr uint32_t kWPart = kC * 128;                     // 8 KB
constexpr uint32_t kWImg = 3 * kWPart;                    // 24 KB

struct BwdSmem {
    static constexpr uint32_t kQ = 0;                     // Q then P: the M = 128 over-read of Q lands in P, of P in W1
    static constexpr uint32_t kP = kImg;
    static constexpr uint32_t kW1 = 2 * kImg;
    static constexpr uint32_t kWk = kW1 + kWImg;          // 3 conv weights
    static constexpr uint32_t kVec = kWk + 3 * kWImg;     // w2[64], gamma[64]
    static constexpr uint32_t kRed = kVec + 2 * kC * 4;   // [7 slots][4 column groups][128 rows]: mean, var, c1, c2, da_0..2
    static constexpr uint32_t kTotal = kRed + 7 * 4 * kTileRows * 4;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void load_slice(const float* __restrict__ src, long long row, int col0, bool ok, float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) t = __ldg(reinterpret_cast<const float4*>(src + row * kC + col0) + j);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
}

__device__ __forceinline__ void store_slice_image(uint8_t* img, int r, int q, const float (&v)[16]) {
    float lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        lo[i] = v[i];
        hi[i] = v[8 + i];
    }
    store_split8(img, kPart, r, 2 * q, lo);
    store_split8(img, kPart, r, 2 * q + 1, hi);
}

// Sum over the warp's 32 lanes of each of 16 per-lane values; lane l returns column (l >> 1) & 15.
__device__ __forceinline__ float colsum16(const float (&v)[16], int lane) {
    float a[8], c[4], d[2];
    bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = up ? a[i] : a[i + 4], keep = up ? a[i + 4] : a[i];
        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = up ? c[i] : c[i + 2], keep = up ? c[i + 2] : c[i];
        d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    up = lane & 2;
    const float send = up ? d[0] : d[1], keep = up ? d[1] : d[0];
    float e = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    e += __shfl_xor_sync(0xffffffffu, e, 1);
    return e;
}

__device__ __forceinline__ void stage_weight16(const float* __restrict__ w, uint8_t* img, int tid) {
    for (int idx = tid; idx < kC * 8; idx += kThreads) {
        const int r = idx >> 3, chunk = idx & 7;
        const float4 a = __ldg(reinterpret_cast<const float4*>(w + r * kC) + chunk * 2);
        const float4 b = __ldg(reinterpret_cast<const float4*>(w + r * kC) + chunk * 2 + 1);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        store_split8(img, kWPart, r, chunk, v);
    }
}

__global__ void __launch_bounds__(kThreads, 1) combine_bwd_fused_kernel(topo_combine_params P, long long rows,
                                                                        const int* __restrict__ n_rows_dev,
                                                                        const float* __restrict__ grad_out,
                                                                        topo_combine_grads G) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw;
    uint8_t* q_img = base + BwdSmem::kQ;
    uint8_t* p_img = base + BwdSmem::kP;
    float* vecs = reinterpret_cast<float*>(base + BwdSmem::kVec);
    float* red = reinterpret_cast<float*>(base + BwdSmem::kRed);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;

    const long long live = n_rows_dev ? min(static_cast<long long>(*n_rows_dev), rows) : rows;
    const long long tiles = (live + kTileRows - 1) / kTileRows;
    if (blockIdx.x >= tiles) return;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = tid >> 7, r = tid & 127, col0 = q * kCW;
    const int n_msgs = P.n_msgs;
    const bool apply_ln = P.apply_ln != 0;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_smem, 512);
    stage_weight16(P.att_w1, base + BwdSmem::kW1, tid);
    for (int k = 0; k < n_msgs; ++k) stage_weight16(P.w[k], base + BwdSmem::kWk + k * kWImg, tid);
    for (int c = tid; c < kC; c += kThreads) {
        vecs[c] = __ldg(P.att_w2 + c);
        vecs[kC + c] = apply_ln ? __ldg(P.ln_gamma + c) : 1.f;
    }
    fence_async_shared();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_smem;
    // tensor-memory columns: T | Ga | DW1 | DWp_0 | DWp_1 | DWp_2
    const uint32_t tm_t = tmem_base, tm_ga = tmem_base + 64, tm_dw1 = tmem_base + 128, tm_dwp = tmem_base + 192;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t p_s = smem_u32(p_img), q_s = smem_u32(q_img);
    const uint32_t w1_s = smem_u32(base + BwdSmem::kW1), wk_s = smem_u32(base + BwdSmem::kWk);
    float scale_r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) scale_r[k] = k < n_msgs ? __ldg(P.scale[k]) : 0.f;

    uint32_t parity = 0, tiles_done = 0;
    float p_b1 = 0.f, p_w2 = 0.f, p_gamma = 0.f, p_beta = 0.f, p_b2 = 0.f;
    float w2r[kCW];
#pragma unroll
    for (int j = 0; j < kCW; ++j) w2r[j] = vecs[col0 + j];

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tiles_done) {
        const long long row = tile * kTileRows + r;
        const bool alive = row < live;
        // ---------------- phase A: attention weights, LayerNorm backward, dscore ----------------
        float dy[kCW];
        load_slice(grad_out, row, col0, alive, dy);
        float sc[3], att[3], dsc[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) sc[k] = (k < n_msgs && alive) ? __ldg(P.saved_score + k * rows + row) : 0.f;
        {
            float mx = sc[0];
            if (n_msgs > 1) mx = fmaxf(mx, sc[1]);
            if (n_msgs > 2) mx = fmaxf(mx, sc[2]);
            const float e0 = expf(sc[0] - mx), e1 = n_msgs > 1 ? expf(sc[1] - mx) : 0.f, e2 = n_msgs > 2 ? expf(sc[2] - mx) : 0.f;
            const float es = e0 + e1 + e2;
            att[0] = e0 / es; att[1] = e1 / es; att[2] = e2 / es;
        }
        if (apply_ln) {
            float y[kCW];
#pragma unroll
            for (int j = 0; j < kCW; ++j) y[j] = 0.f;
#pragma unroll 1
            for (int k = 0; k < n_msgs; ++k) {
                float mk[kCW];
                load_slice(P.saved_m[k], row, col0, alive, mk);
                const float a = k == 0 ? att[0] : (k == 1 ? att[1] : att[2]);
#pragma unroll
                for (int j = 0; j < kCW; ++j) y[j] = fmaf(a, mk[j], y[j]);
            }
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < kCW; ++j) part += y[j];
            red[(0 * 4 + q) * kTileRows + r] = part;
            __syncthreads();
            const float mean = (red[0 * kTileRows + r] + red[1 * kTileRows + r] + red[2 * kTileRows + r] + red[3 * kTileRows + r]) * (1.0f / kC);
            float var = 0.f;
#pragma unroll
            for (int j = 0; j < kCW; ++j) var = fmaf(y[j] - mean, y[j] - mean, var);
            red[(1 * 4 + q) * kTileRows + r] = var;
            __syncthreads();
            const float* rv = red + 4 * kTileRows;
            const float rstd = 1.0f / sqrtf((rv[r] + rv[kTileRows + r] + rv[2 * kTileRows + r] + rv[3 * kTileRows + r]) * (1.0f / kC) + P.ln_eps);
            float gx[kCW], c1 = 0.f, c2 = 0.f;
#pragma unroll
            for (int j = 0; j < kCW; ++j) {
                y[j] = (y[j] - mean) * rstd;                       // x-hat
                gx[j] = dy[j] * y[j];                              // gamma gradient term
                const float gy = dy[j] * vecs[kC + col0 + j];
                c1 += gy;
                c2 = fmaf(gy, y[j], c2);
            }
            p_gamma += colsum16(gx, lane);
            p_beta += colsum16(dy, lane);
            red[(2 * 4 + q) * kTileRows + r] = c1;
            red[(3 * 4 + q) * kTileRows + r] = c2;
            __syncthreads();
            const float* r1 = red + 8 * kTileRows;
            const float* r2 = red + 12 * kTileRows;
            c1 = (r1[r] + r1[kTileRows + r] + r1[2 * kTileRows + r] + r1[3 * kTileRows + r]) * (1.0f / kC);
            c2 = (r2[r] + r2[kTileRows + r] + r2[2 * kTileRows + r] + r2[3 * kTileRows + r]) * (1.0f / kC);
#pragma unroll
            for (int j = 0; j < kCW; ++j) dy[j] = rstd * (dy[j] * vecs[kC + col0 + j] - c1 - y[j] * c2);
        }
        // dscore_k = a_k (dy . m_k - sum_j a_j dy . m_j)
#pragma unroll 1
        for (int k = 0; k < n_msgs; ++k) {
            float mk[kCW];
            load_slice(P.saved_m[k], row, col0, alive, mk);
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < kCW; ++j) part = fmaf(dy[j], mk[j], part);
            red[((4 + k) * 4 + q) * kTileRows + r] = part;     // own slots: c1 / c2 may still be being read
        }
        __syncthreads();
        {
            float da[3], dot = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float* rk = red + (4 + k) * 4 * kTileRows;
                da[k] = k < n_msgs ? (rk[r] + rk[kTileRows + r] + rk[2 * kTileRows + r] + rk[3 * kTileRows + r]) : 0.f;
                dot = fmaf(att[k], da[k], dot);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dsc[k] = (k < n_msgs && alive) ? att[k] * (da[k] - dot) : 0.f;
                if (q == 0) p_b2 += dsc[k];
            }
        }

        // ---------------- per message: two MMA rounds ----------------
        float dx[kCW];
#pragma unroll
        for (int j = 0; j < kCW; ++j) dx[j] = 0.f;
        float pre[kCW];
        load_slice(P.saved_pre[0], row, col0, alive, pre);
#pragma unroll 1
        for (int k = 0; k < n_msgs; ++k) {
            const float dsk = k == 0 ? dsc[0] : (k == 1 ? dsc[1] : dsc[2]);
            const float ak = k == 0 ? att[0] : (k == 1 ? att[1] : att[2]);
            // dpre_k (evaluated while the previous round 2 is still running)
            float ge[kCW];
#pragma unroll
            for (int j = 0; j < kCW; ++j) {
                const float x = pre[j];
                const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
                const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
                ge[j] = dsk * (x * cdf);                           // dscore_k GELU(pre): w2 gradient term
                pre[j] = dsk * w2r[j] * (cdf + x * pdf);           // dpre_k
            }
            p_w2 += colsum16(ge, lane);
            p_b1 += colsum16(pre, lane);
            if (k > 0) {
                // round 2 of message k - 1: Ga -> g_agg_{k-1}
                mbar_wait(&bar, parity);
                parity ^= 1;
                tc_fence_after_sync();
                float ga[kCW];
                tmem_ld16(tm_ga + lane_addr + col0, ga);
                const float sp = k == 1 ? scale_r[0] : scale_r[1];
                if (alive) {
                    float4* dst = reinterpret_cast<float4*>(G.g_agg[k - 1] + row * kC + col0);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        dst[j] = make_float4(sp * ga[4 * j], sp * ga[4 * j + 1], sp * ga[4 * j + 2], sp * ga[4 * j + 3]);
                }
            }
            store_slice_image(p_img, r, q, pre);
            {
                float mk[kCW];
                load_slice(P.saved_m[k], row, col0, alive, mk);
                store_slice_image(q_img, r, q, mk);
            }
            fence_async_shared();
            tc_fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after_sync();
                gemm_bf16x3(tm_t, k_major(p_s, kTileRows), mn_major(w1_s, kC, kWPart), idesc_bf16(128, 64, 0, 1), kC / 16, 0);
                gemm_bf16x3(tm_dw1, mn_major(p_s, kTileRows, kPart), mn_major(q_s, kTileRows, kPart), idesc_bf16(128, 64, 1, 1),
                            kTileRows / 16, (tiles_done | static_cast<uint32_t>(k)) != 0u);
                mma_commit(&bar);
            }
            float ag[kCW];
            load_slice(P.agg[k], row, col0, alive, ag);
            mbar_wait(&bar, parity);
            parity ^= 1;
            tc_fence_after_sync();
            {
                float t[kCW];
                tmem_ld16(tm_t + lane_addr + col0, t);
#pragma unroll
                for (int j = 0; j < kCW; ++j) {
                    t[j] = alive ? fmaf(ak, dy[j], t[j]) : 0.f;    // dm_k
                    dx[j] += t[j];
                }
                store_slice_image(p_img, r, q, t);
            }
            store_slice_image(q_img, r, q, ag);
            fence_async_shared();
            tc_fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after_sync();
                gemm_bf16x3(tm_ga, k_major(p_s, kTileRows), k_major(wk_s + k * kWImg, kC), idesc_bf16(128, 64, 0, 0), kC / 16, 0);
                gemm_bf16x3(tm_dwp + k * 64, mn_major(q_s, kTileRows, kPart), mn_major(p_s, kTileRows, kPart),
                            idesc_bf16(128, 64, 1, 1), kTileRows / 16, tiles_done != 0u);
                mma_commit(&bar);
            }
            if (k + 1 < n_msgs) load_slice(P.saved_pre[k + 1], row, col0, alive, pre);
        }
        // round 2 of the last message
        mbar_wait(&bar, parity);
        parity ^= 1;
        tc_fence_after_sync();
        {
            float ga[kCW];
            tmem_ld16(tm_ga + lane_addr + col0, ga);
            const float sp = n_msgs == 1 ? scale_r[0] : (n_msgs == 2 ? scale_r[1] : scale_r[2]);
            if (alive) {
                float4* dst = reinterpret_cast<float4*>(G.g_agg[n_msgs - 1] + row * kC + col0);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    dst[j] = make_float4(sp * ga[4 * j], sp * ga[4 * j + 1], sp * ga[4 * j + 2], sp * ga[4 * j + 3]);
                if (G.g_x != nullptr) {
                    float4* dxp = reinterpret_cast<float4*>(G.g_x + row * kC + col0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dxp[j] = make_float4(dx[4 * j], dx[4 * j + 1], dx[4 * j + 2], dx[4 * j + 3]);
                }
            }
        }
        tc_fence_before_sync();
    }

    // ---------------- parameter gradients of this CTA ----------------
    __syncthreads();
    tc_fence_after_sync();
    if ((lane & 1) == 0) {
        const int c = col0 + ((lane >> 1) & 15);
        atomicAdd(G.g_att_b1 + c, p_b1);
        atomicAdd(G.g_att_w2 + c, p_w2);
        if (apply_ln) {
            atomicAdd(G.g_ln_gamma + c, p_gamma);
            atomicAdd(G.g_ln_beta + c, p_beta);
        }
    }
    if (q == 0) {
        p_b2 = warp_sum(p_b2);
        if (lane == 0 && p_b2 != 0.f) atomicAdd(G.g_att_b2, p_b2);
    }
    if ((warp & 3) < 2) {
        // accumulator rows 0..63 live in lanes 0..63; warp group g = warp / 4 drains DW1 (g = 0) or DWp_{g-1}
        const int g = warp >> 2;
        const int drow = (warp & 3) * 32 + lane;
        float* dst = g == 0 ? G.g_att_w1 : (g - 1 < n_msgs ? G.g_wprod[g - 1] : nullptr);
        if (dst != nullptr) {
            const uint32_t src = (g == 0 ? tm_dw1 : tm_dwp + (g - 1) * 64) + lane_addr;
#pragma unroll 1
            for (int c8 = 0; c8 < 8; ++c8) {
                float v[8];
                tmem_ld8(src + c8 * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) atomicAdd(dst + drow * kC + c8 * 8 + i, v[i]);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace
}  // namespace topo

using namespace topo;

extern "C" int topo_sccn_combine_bwd_tc(const topo_combine_params* p, int64_t rows, const int32_t* n_rows_dev,
                                        const float* grad_out, const topo_combine_grads* g, topo_stream_t stream) {
    TOPO_REQUIRE(p && g && grad_out && rows >= 0, "bad argument");
    TOPO_REQUIRE(p->n_msgs >= 1 && p->n_msgs <= 3, "n_msgs must be 1..3");
    if (p->channels != kC) {
        set_error("the tensor-core combine is instantiated for channels == 64");
        return TOPO_ERR_UNSUPPORTED;
    }
    TOPO_REQUIRE(p->saved_score, "the fused backward needs the forward's saved scores");
    for (int k = 0; k < p->n_msgs; ++k)
        TOPO_REQUIRE(p->agg[k] && p->w[k] && p->scale[k] && p->saved_m[k] && p->saved_pre[k] && g->g_agg[k] && g->g_wprod[k],
                     "null message operand (the fused backward needs the forward's saved activations)");
    TOPO_REQUIRE(p->att_w1 && p->att_w2 && g->g_att_w1 && g->g_att_b1 && g->g_att_w2 && g->g_att_b2, "null attention parameter");
    TOPO_REQUIRE(!p->apply_ln || (p->ln_gamma && g->g_ln_gamma && g->g_ln_beta), "null LayerNorm parameter");
    if (rows == 0) return TOPO_OK;
    const size_t smem = BwdSmem::kTotal + 1024;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(combine_bwd_fused_kernel), smem)) return rc;
    const int tiles = static_cast<int>((rows + kTileRows - 1) / kTileRows);
    combine_bwd_fused_kernel<<<std::min(tiles, sm_count()), kThreads, smem, as_stream(stream)>>>(*p, rows, n_rows_dev, grad_out, *g);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
