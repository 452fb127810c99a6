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
// DW1 and DWp_k (64 x 64 each) stay in tensor memory for the whole CTA and are added to global memory once
// (red.global.add.v4.f32).  The row-contraction MMAs are issued with M = 64 (accumulator row i in TMEM lane
// 32 (i / 16) + i % 16); the 72 MMAs of a round are issued by two or three threads, ONE per accumulator (the order in which
// the tensor pipe adds into an accumulator is then fixed: with per-CTA partial sums the parameter gradients are bit-reproducible).
// Column sums (b1, w2, gamma, beta gradients) accumulate in the chunk map's registers (eight columns per thread) and
// meet once per CTA: two shuffles per value, one pass through shared memory, one atomic per column.
// Every global load is issued at least one phase before its first use (phase-A operands of the next tile before the
// last MMA round is awaited, pre_{k+1} underneath round 1 of message k, pre_0 inside phase A) and everything is read
// exactly once, with the streaming cache policy.
#include <algorithm>

#include "common.cuh"
#include "layout.cuh"
#include "tc16.cuh"

#ifndef TOPO_DEBUG_KERNELS
#define TOPO_DEBUG_KERNELS 0
#endif

namespace topo {
namespace {

using namespace tc16;

constexpr bool kRowT = (TOPO_ROWMAP_TRANSPOSE & 2) != 0;

constexpr int kTileRows = 128;
constexpr int kC = 64;
constexpr int kThreads = 512;
constexpr int kCW = 16;                                   // columns per thread
constexpr uint32_t kPart = kTileRows * 128;               // 16 KB
constexpr uint32_t kImg = 3 * kPart;                      // 48 KB
constexpr uint32_t kWPart = kC * 128;                     // 8 KB
constexpr uint32_t kWImg = 3 * kWPart;                    // 24 KB

struct BwdSmem {
    static constexpr uint32_t kQ = 0;
    static constexpr uint32_t kP = kImg;
    static constexpr uint32_t kW1 = 2 * kImg;
    static constexpr uint32_t kWk = kW1 + kWImg;          // 3 conv weights
    static constexpr uint32_t kDy = kWk + 3 * kWImg;      // dy tile [128][64] fp32, 16-byte chunks XOR-swizzled by row
    static constexpr uint32_t kVec = kDy + kTileRows * kC * 4;   // w2[64], gamma[64]
    static constexpr uint32_t kAtt = kVec + 2 * kC * 4;   // softmax weights [3][128]
    static constexpr uint32_t kBar = kAtt + 3 * kTileRows * 4;   // mbarrier, tensor-memory base
    static constexpr uint32_t kTotal = kBar + 16;
};
static_assert(BwdSmem::kTotal <= 232448, "shared-memory budget of one CTA");

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

// eight consecutive columns of one row (two 128-bit loads; eight lanes cover a whole 256-byte row)
__device__ __forceinline__ void load_chunk(const float* __restrict__ src, long long row, int chunk, bool ok, float (&v)[8]) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (ok) {
        const float4* p = reinterpret_cast<const float4*>(src + row * kC) + chunk * 2;
        a = ldg_pinned_once(p);       // issued where written: these are prefetches into registers; every row-major
        b = ldg_pinned_once(p + 1);   // input of this kernel (dout, agg_k, row-major saved tensors) is read exactly once
    }
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// the same eight columns of a saved activation, in either layout (topo_combine_params.saved_layout)
__device__ __forceinline__ void load_saved(const float* __restrict__ src, bool tile_fragment, long long row0, int tile_row, int chunk,
                                           bool ok, float (&v)[8]) {
    if (!tile_fragment) {
        load_chunk(src, row0 + tile_row, chunk, ok, v);
        return;
    }
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (ok) {
        const float4* p = reinterpret_cast<const float4*>(src) + tf_index_chunk(row0, tile_row, chunk);
        a = ldg_pinned_once(p);
        b = ldg_pinned_once(p + kTileRows);
    }
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// sum over the eight lanes that share a tile row
__device__ __forceinline__ float row_sum8(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// the same for two values at once (independent shuffles interleave)
__device__ __forceinline__ void row_sum8_pair(float& a, float& b) {
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        const float ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o);
        a += ta;
        b += tb;
    }
}


__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ uint32_t dy_off(int row, int c16 /* 0..15 */) {
    return static_cast<uint32_t>(row * 256 + ((c16 ^ (row & 7)) << 4));
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

// Two thread <-> data mappings:
//   chunk map: thread t owns columns [8 c, 8 c + 8), c = t % 8, of tile rows t / 8 and t / 8 + 64.  Global
//              loads are coalesced (eight lanes = one 256-byte row), row reductions are three shuffles, and
//              the column sums accumulate in registers.  All element-wise work runs in this mapping.
//   row map:   thread (q, r) = (t / 128, t % 128) owns columns [16 q, 16 q + 16) of row r = TMEM lane r.  Only
//              the two tensor-memory epilogues use it; dy and the softmax weights cross over in shared memory.
template <int NM>
__global__ void __launch_bounds__(kThreads, 1) combine_bwd_fused_kernel(topo_combine_params P, long long rows,
                                                                        const int* __restrict__ n_rows_dev,
                                                                        const float* __restrict__ grad_out,
                                                                        topo_combine_grads G,
                                                                        unsigned long long* __restrict__ stamps) {
    // in-kernel timeline for scripts/ablate_bwd.py: compiled in only with -DTOPO_DEBUG_KERNELS=1 (build.py reads the
    // environment variable of that name); the shipped build has no trace of it
    int stamp_no = 0;
    auto stamp = [&]() {
#if TOPO_DEBUG_KERNELS
        if (stamps != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && stamp_no < 62) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            stamps[stamp_no] = t;
        }
        ++stamp_no;
#endif
    };
    stamp();                                                   // 0: kernel start
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw;
    uint8_t* q_img = base + BwdSmem::kQ;
    uint8_t* p_img = base + BwdSmem::kP;
    uint8_t* dy_s = base + BwdSmem::kDy;
    float* vecs = reinterpret_cast<float*>(base + BwdSmem::kVec);
    float* att_s = reinterpret_cast<float*>(base + BwdSmem::kAtt);
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + BwdSmem::kBar);
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(base + BwdSmem::kBar + 8);

    const long long live = n_rows_dev ? min(static_cast<long long>(*n_rows_dev), rows) : rows;
    const long long tiles = (live + kTileRows - 1) / kTileRows;
    if (blockIdx.x >= tiles) return;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = tid >> 7, r = tid & 127, col0 = q * kCW;     // row map
    const int c = tid & 7, ra = tid >> 3;                       // chunk map: rows ra and ra + 64
    constexpr int n_msgs = NM;
    // The tensor pipe adds the MMAs of DIFFERENT issuing threads into one accumulator in an order that varies from run to run,
    // so every weight-gradient accumulator belongs to ONE issuing thread.  With up to two messages the two halves of a row
    // contraction (tile rows 0-63 and 64-127) get an accumulator each -- three issuing threads, all 512 columns used -- and are
    // added when the CTA flushes; with three messages tensor memory has no room for that and one thread issues both halves.
    constexpr bool kSplit = NM <= 2;
    constexpr uint32_t kSecondHalf = kSplit ? 64u * (1 + NM) : 0u;       // column offset of the second-half accumulators
    const bool apply_ln = P.apply_ln != 0;
    const bool tf = P.saved_layout == TOPO_SAVED_TILE_FRAGMENT;

    if (tid == 0) {
        mbar_init(bar, kSplit ? 3 : 2);          // the issuing threads commit every round
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_base_smem, 512);
    if (P.weight_images != nullptr) {
        // images built once per layer (weight_images.cu): [W1 | k: W_k, V_k]; the backward uses W1 and the W_k
        const uint4* __restrict__ src = reinterpret_cast<const uint4*>(P.weight_images);
        constexpr int kImg16 = static_cast<int>(kWImg / 16);
        uint4* d1 = reinterpret_cast<uint4*>(base + BwdSmem::kW1);
        for (int idx = tid; idx < kImg16; idx += kThreads) d1[idx] = __ldg(src + idx);
        for (int k = 0; k < n_msgs; ++k) {
            uint4* dk = reinterpret_cast<uint4*>(base + BwdSmem::kWk + k * kWImg);
            const uint4* sk = src + kImg16 * (1 + 2 * k);
            for (int idx = tid; idx < kImg16; idx += kThreads) dk[idx] = __ldg(sk + idx);
        }
    } else {
        stage_weight16(P.att_w1, base + BwdSmem::kW1, tid);
        for (int k = 0; k < n_msgs; ++k) stage_weight16(P.w[k], base + BwdSmem::kWk + k * kWImg, tid);
    }
    for (int i = tid; i < kC; i += kThreads) {
        vecs[i] = __ldg(P.att_w2 + i);
        vecs[kC + i] = apply_ln ? __ldg(P.ln_gamma + i) : 1.f;
    }
    fence_async_shared();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_base_smem;
    // tensor-memory columns: T | Ga | DW1 | DWp_0 .. DWp_{NM-1} [| the same again for tile rows 64-127 when kSplit]
    const uint32_t tm_t = tmem_base, tm_ga = tmem_base + 64, tm_dw1 = tmem_base + 128, tm_dwp = tmem_base + 192;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    {
        // DW1 / DWp start at zero: every weight-gradient MMA accumulates, whoever issues it first
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) tmem_st8(tm_dw1 + lane_addr + q * 64 + c8 * 8, z);
        if (kSplit && NM == 2 && q < 2) {
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) tmem_st8(tm_dw1 + lane_addr + 256 + q * 64 + c8 * 8, z);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
    }
    const uint32_t p_s = smem_u32(p_img), q_s = smem_u32(q_img);
    const uint32_t w1_s = smem_u32(base + BwdSmem::kW1), wk_s = smem_u32(base + BwdSmem::kWk);
    float scale_r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) scale_r[k] = k < n_msgs ? __ldg(P.scale[k]) : 0.f;
    const float* w2c = vecs + 8 * c;                            // this thread's eight columns of w2 and gamma
    const float* gmc = vecs + kC + 8 * c;

    uint32_t parity = 0, tiles_done = 0;
    float p_b1[8], p_w2[8], p_gamma[8], p_beta[8], p_b2 = 0.f;  // column sums over this thread's rows
#pragma unroll
    for (int i = 0; i < 8; ++i) p_b1[i] = p_w2[i] = p_gamma[i] = p_beta[i] = 0.f;

    // phase-A operands of the NEXT tile are loaded before the current tile's last MMA round is awaited
    float n_dy[2][8], n_mk[3][2][8], n_sc[2][3];
    auto load_phase_a = [&](long long t) {
        const long long r0 = t * kTileRows;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const long long gr = r0 + ra + 64 * j;
            const bool ok = t < tiles && gr < live;
#pragma unroll
            for (int k = 0; k < 3; ++k) n_sc[j][k] = (k < n_msgs && ok) ? __ldg(P.saved_score + k * rows + gr) : 0.f;
            load_chunk(grad_out, gr, c, ok, n_dy[j]);
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (k < n_msgs) load_saved(P.saved_m[k], tf, r0, ra + 64 * j, c, ok, n_mk[k][j]);
        }
    };
    load_phase_a(blockIdx.x);
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tiles_done) {
        const long long row0 = tile * kTileRows;
        const long long grow[2] = {row0 + ra, row0 + ra + 64};
        const bool alive[2] = {grow[0] < live, grow[1] < live};
        stamp();                                               // tile start
        // ---------------- phase A (chunk map): softmax weights, LayerNorm backward, dscore ----------------
        // Both rows of the thread advance together (every statement is written for j = 0, 1) so that their
        // dependent chains -- loads, row reductions, exp, rsqrt -- overlap instead of running back to back.
        float dy[2][8], dsc[2][3], pre[2][8];
        {
            float mk[3][2][8], att[2][3], sc[2][3];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
                for (int k = 0; k < 3; ++k) sc[j][k] = n_sc[j][k];
#pragma unroll
                for (int i = 0; i < 8; ++i) dy[j][i] = n_dy[j][i];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (k < n_msgs) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) mk[k][j][i] = n_mk[k][j][i];
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float mx = sc[j][0];
                if (n_msgs > 1) mx = fmaxf(mx, sc[j][1]);
                if (n_msgs > 2) mx = fmaxf(mx, sc[j][2]);
                const float e0 = expf(sc[j][0] - mx), e1 = n_msgs > 1 ? expf(sc[j][1] - mx) : 0.f, e2 = n_msgs > 2 ? expf(sc[j][2] - mx) : 0.f;
                const float inv = 1.0f / (e0 + e1 + e2);
                att[j][0] = e0 * inv; att[j][1] = e1 * inv; att[j][2] = e2 * inv;
                if (c == 0) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) att_s[k * kTileRows + ra + 64 * j] = att[j][k];
                }
            }
            stamp();                                           // scores arrived, softmax done
            if (apply_ln) {
                float y[2][8], s0[2], mean[2], rstd[2], c1[2], c2[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[j][i] = 0.f;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (k < n_msgs) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) y[j][i] = fmaf(att[j][k], mk[k][j][i], y[j][i]);
                        }
                    }
                    s0[j] = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) s0[j] += y[j][i];
                }
                row_sum8_pair(s0[0], s0[1]);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    mean[j] = s0[j] * (1.0f / kC);
                    s0[j] = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) s0[j] = fmaf(y[j][i] - mean[j], y[j][i] - mean[j], s0[j]);
                }
                row_sum8_pair(s0[0], s0[1]);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    rstd[j] = 1.0f / sqrtf(s0[j] * (1.0f / kC) + P.ln_eps);
                    c1[j] = c2[j] = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        y[j][i] = (y[j][i] - mean[j]) * rstd[j];              // x-hat
                        p_gamma[i] = fmaf(dy[j][i], y[j][i], p_gamma[i]);
                        p_beta[i] += dy[j][i];
                        const float gy = dy[j][i] * gmc[i];
                        c1[j] += gy;
                        c2[j] = fmaf(gy, y[j][i], c2[j]);
                    }
                }
                row_sum8_pair(c1[0], c1[1]);
                row_sum8_pair(c2[0], c2[1]);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    c1[j] *= (1.0f / kC);
                    c2[j] *= (1.0f / kC);
#pragma unroll
                    for (int i = 0; i < 8; ++i) dy[j][i] = rstd[j] * (dy[j][i] * gmc[i] - c1[j] - y[j][i] * c2[j]);
                }
            }
            // the first message's hidden pre-activations: issued here, used right after phase A
#pragma unroll
            for (int j = 0; j < 2; ++j) load_saved(P.saved_pre[0], tf, row0, ra + 64 * j, c, alive[j], pre[j]);
            // dscore_k = a_k (dy . m_k - sum_j a_j dy . m_j)
            float da[2][3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    da[j][k] = 0.f;
                    if (k < n_msgs) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) da[j][k] = fmaf(dy[j][i], mk[k][j][i], da[j][k]);
                    }
                }
                if (k < n_msgs) row_sum8_pair(da[0][k], da[1][k]);
            }
            stamp();                                           // LayerNorm backward + dscore done
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float dot = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) dot = fmaf(att[j][k], da[j][k], dot);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    dsc[j][k] = (k < n_msgs && alive[j]) ? att[j][k] * (da[j][k] - dot) : 0.f;
                    if (c == 0) p_b2 += dsc[j][k];
                }
                // dy crosses to the row map through shared memory.  Lanes 0..3 of a row store their first 16-byte
                // chunk while lanes 4..7 store their second one (and vice versa): eight distinct bank groups.
                const int rr = ra + 64 * j;
                const float4 lo4 = make_float4(dy[j][0], dy[j][1], dy[j][2], dy[j][3]);
                const float4 hi4 = make_float4(dy[j][4], dy[j][5], dy[j][6], dy[j][7]);
                const bool first = c < 4;
                *reinterpret_cast<float4*>(dy_s + dy_off(rr, 2 * c + (first ? 0 : 1))) = first ? lo4 : hi4;
                *reinterpret_cast<float4*>(dy_s + dy_off(rr, 2 * c + (first ? 1 : 0))) = first ? hi4 : lo4;
                // message 0 goes straight into its operand image (Q is idle: the previous tile's last round was awaited)
                store_split8(q_img, kPart, rr, c, mk[0][j]);
            }
        }
        if (tile + gridDim.x < tiles) {
            // the next tile's operands start moving towards L2 now: one 128-byte line per thread and tensor slice
            const long long nrow0 = (tile + gridDim.x) * kTileRows;
            const long long nlines = min(static_cast<long long>(kTileRows), live - nrow0) * (kC * 4 / 128);   // row-major tensors
            const int line = tid & 255;
            const bool upper = tid >= 256;
            if (!upper && line < nlines) prefetch_l2_line(grad_out + nrow0 * kC + line * 32);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (k < n_msgs) {
                    if (upper && line < nlines) prefetch_l2_line(P.agg[k] + nrow0 * kC + line * 32);
                    if ((tf || line < nlines)) prefetch_l2_line((upper ? P.saved_pre[k] : P.saved_m[k]) + nrow0 * kC + line * 32);
                }
            }
        }
        stamp();                                               // phase A done
        // ---------------- per message: two MMA rounds ----------------
        float dx[kCW];                                                        // row map
#pragma unroll
        for (int i = 0; i < kCW; ++i) dx[i] = 0.f;
        const long long row = row0 + r;                                       // row map
        const bool row_alive = row < live;
#pragma unroll 1
        for (int k = 0; k < n_msgs; ++k) {
            float mq[2][8];
            if (k > 0) {
#pragma unroll
                for (int j = 0; j < 2; ++j) load_saved(P.saved_m[k], tf, row0, ra + 64 * j, c, alive[j], mq[j]);   // hidden behind the GELU math
            }
            // dpre_k (evaluated while the previous round 2 is still running)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float dsk = k == 0 ? dsc[j][0] : (k == 1 ? dsc[j][1] : dsc[j][2]);
#pragma unroll
                for (int i0 = 0; i0 < 8; i0 += 4) {
                    float cdf[4], pdf[4];
                    gelu_cdf_pdf_n<4>(pre[j] + i0, cdf, pdf);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int i = i0 + e;
                        const float x = pre[j][i];
                        p_w2[i] = fmaf(dsk, x * cdf[e], p_w2[i]);             // dscore_k GELU(pre): w2 gradient
                        const float d = dsk * w2c[i] * (cdf[e] + x * pdf[e]); // dpre_k
                        p_b1[i] += d;
                        pre[j][i] = d;
                    }
                }
            }
            if (k > 0) {
                // round 2 of message k - 1: Ga -> g_agg_{k-1}   (row map)
                mbar_wait_backoff(bar, parity);
                parity ^= 1;
                tc_fence_after_sync();
                float ga[kCW];
                tmem_ld16(tm_ga + lane_addr + col0, ga);
                const float sp = k == 1 ? scale_r[0] : scale_r[1];
#pragma unroll
                for (int i = 0; i < kCW; ++i) ga[i] *= sp;
                rowmap_store16<kRowT>(G.g_agg[k - 1], row, col0, live, ga, lane);       // 64 contiguous bytes per row and access
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                store_split8(p_img, kPart, ra + 64 * j, c, pre[j]);
                if (k > 0) store_split8(q_img, kPart, ra + 64 * j, c, mq[j]);
            }
            stamp();                                           // GELU' + images staged (before barrier)
            fence_async_shared();
            tc_fence_before_sync();
            __syncthreads();
            stamp();                                           // barrier 1 passed
            if (warp < (kSplit ? 3 : 2)) {
                // one issuing thread per accumulator: the 24 MMAs of the row product, the 48 of the row contraction in one or two halves
                if (lane == 0) {
                    tc_fence_after_sync();
                    if (warp == 0)
                        gemm_bf16x3_unrolled<kC / 16>(tm_t, k_major(p_s, kTileRows), mn_major(w1_s, kC, kWPart), idesc_bf16(128, 64, 0, 1), 0);
                    else if (!kSplit)
                        gemm_bf16x3_unrolled<8, 0>(tm_dw1, mn_major(p_s, kTileRows, kPart), mn_major(q_s, kTileRows, kPart), idesc_bf16(64, 64, 1, 1), 1);
                    else if (warp == 1)
                        gemm_bf16x3_unrolled<4, 0>(tm_dw1, mn_major(p_s, kTileRows, kPart), mn_major(q_s, kTileRows, kPart), idesc_bf16(64, 64, 1, 1), 1);
                    else
                        gemm_bf16x3_unrolled<4, 4>(tm_dw1 + kSecondHalf, mn_major(p_s, kTileRows, kPart), mn_major(q_s, kTileRows, kPart), idesc_bf16(64, 64, 1, 1), 1);
                    mma_commit(bar);
                }
                __syncwarp();      // the other lanes park here instead of polling against the issuing lane
            }
            float ag[2][8];
#pragma unroll
            for (int j = 0; j < 2; ++j) load_chunk(P.agg[k], grow[j], c, alive[j], ag[j]);
            // the next message's hidden pre-activations travel while this round runs (their image is staged: `pre` is free)
            if (k + 1 < n_msgs) {
#pragma unroll
                for (int j = 0; j < 2; ++j) load_saved(P.saved_pre[k + 1], tf, row0, ra + 64 * j, c, alive[j], pre[j]);
            }
            stamp();                                           // round 1 issued, agg loads issued
            mbar_wait_backoff(bar, parity);
            parity ^= 1;
            tc_fence_after_sync();
            stamp();                                           // round 1 complete
            {
                // the row map's view of dy and a_k
                float t[kCW];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(dy_s + dy_off(r, 4 * q + i));
                    t[4 * i] = v.x; t[4 * i + 1] = v.y; t[4 * i + 2] = v.z; t[4 * i + 3] = v.w;
                }
                const float ak = att_s[k * kTileRows + r];
                float tt[kCW];
                tmem_ld16(tm_t + lane_addr + col0, tt);
                float lo[8], hi[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    lo[i] = row_alive ? fmaf(ak, t[i], tt[i]) : 0.f;          // dm_k
                    hi[i] = row_alive ? fmaf(ak, t[8 + i], tt[8 + i]) : 0.f;
                    dx[i] += lo[i];
                    dx[8 + i] += hi[i];
                }
                store_split8(p_img, kPart, r, 2 * q, lo);
                store_split8(p_img, kPart, r, 2 * q + 1, hi);
                if (k == n_msgs - 1 && G.g_x != nullptr) rowmap_store16<kRowT>(G.g_x, row, col0, live, dx, lane);   // g_x = sum_k dm_k is complete here
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) store_split8(q_img, kPart, ra + 64 * j, c, ag[j]);
            stamp();                                           // dm + agg images staged
            fence_async_shared();
            tc_fence_before_sync();
            __syncthreads();
            if (warp < (kSplit ? 3 : 2)) {
                if (lane == 0) {
                    tc_fence_after_sync();
                    if (warp == 0)
                        gemm_bf16x3_unrolled<kC / 16>(tm_ga, k_major(p_s, kTileRows), k_major(wk_s + k * kWImg, kC), idesc_bf16(128, 64, 0, 0), 0);
                    else if (!kSplit)
                        gemm_bf16x3_unrolled<8, 0>(tm_dwp + k * 64, mn_major(q_s, kTileRows, kPart), mn_major(p_s, kTileRows, kPart), idesc_bf16(64, 64, 1, 1), 1);
                    else if (warp == 1)
                        gemm_bf16x3_unrolled<4, 0>(tm_dwp + k * 64, mn_major(q_s, kTileRows, kPart), mn_major(p_s, kTileRows, kPart), idesc_bf16(64, 64, 1, 1), 1);
                    else
                        gemm_bf16x3_unrolled<4, 4>(tm_dwp + k * 64 + kSecondHalf, mn_major(q_s, kTileRows, kPart), mn_major(p_s, kTileRows, kPart), idesc_bf16(64, 64, 1, 1), 1);
                    mma_commit(bar);
                }
                __syncwarp();
            }
        }
        // round 2 of the last message: the next tile's phase-A operands start moving before it is awaited
        stamp();                                               // last round 2 issued
        load_phase_a(tile + gridDim.x);
        mbar_wait_backoff(bar, parity);
        parity ^= 1;
        tc_fence_after_sync();
        {
            float ga[kCW];
            tmem_ld16(tm_ga + lane_addr + col0, ga);
            const float sp = n_msgs == 1 ? scale_r[0] : (n_msgs == 2 ? scale_r[1] : scale_r[2]);
#pragma unroll
            for (int i = 0; i < kCW; ++i) ga[i] *= sp;
            rowmap_store16<kRowT>(G.g_agg[n_msgs - 1], row, col0, live, ga, lane);
        }
        tc_fence_before_sync();
        __syncthreads();          // dy / a_k in shared memory are rewritten by the next tile's phase A
        stamp();                                               // tile done
    }

    // ---------------- parameter gradients of this CTA ----------------
    tc_fence_after_sync();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        // fold the four row groups of the warp (lanes with equal c): lanes 0..7 then hold the warp's column sums
        p_b1[i] += __shfl_xor_sync(0xffffffffu, p_b1[i], 8);
        p_b1[i] += __shfl_xor_sync(0xffffffffu, p_b1[i], 16);
        p_w2[i] += __shfl_xor_sync(0xffffffffu, p_w2[i], 8);
        p_w2[i] += __shfl_xor_sync(0xffffffffu, p_w2[i], 16);
        p_gamma[i] += __shfl_xor_sync(0xffffffffu, p_gamma[i], 8);
        p_gamma[i] += __shfl_xor_sync(0xffffffffu, p_gamma[i], 16);
        p_beta[i] += __shfl_xor_sync(0xffffffffu, p_beta[i], 8);
        p_beta[i] += __shfl_xor_sync(0xffffffffu, p_beta[i], 16);
    }
    {
        // one atomic per column and CTA: the 16 warps meet in shared memory first (the dy tile is idle now)
        float* part = reinterpret_cast<float*>(dy_s);          // [16 warps][4 quantities][64 columns]
        if (lane < 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                part[(warp * 4 + 0) * kC + 8 * c + i] = p_b1[i];
                part[(warp * 4 + 1) * kC + 8 * c + i] = p_w2[i];
                part[(warp * 4 + 2) * kC + 8 * c + i] = p_gamma[i];
                part[(warp * 4 + 3) * kC + 8 * c + i] = p_beta[i];
            }
        }
        p_b2 = warp_sum(p_b2);
        if (lane == 0) part[(kThreads / 32) * 4 * kC + warp] = p_b2;
        __syncthreads();
        // with a partials buffer this CTA's sums go to its own slot with plain stores (topo_sccn_finish_weight_grads adds the
        // slots in CTA order: bit-reproducible); without one they are added to the shared accumulators atomically
        float* slot = G.cta_partials != nullptr ? G.cta_partials + static_cast<size_t>(blockIdx.x) * kCtaPartialFloats : nullptr;
        if (tid < 4 * kC) {
            const int which = tid >> 6, col = tid & 63;
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) sum += part[(w * 4 + which) * kC + col];
            if (slot != nullptr) {
                slot[4 * kC * kC + which * kC + col] = sum;      // b1, w2, gamma, beta
            } else {
                float* dst = which == 0 ? G.g_att_b1 : (which == 1 ? G.g_att_w2 : (which == 2 ? G.g_ln_gamma : G.g_ln_beta));
                if (which < 2 || apply_ln) atomicAdd(dst + col, sum);
            }
        } else if (tid == 4 * kC) {
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) sum += part[(kThreads / 32) * 4 * kC + w];
            if (slot != nullptr) slot[4 * kC * kC + 4 * kC] = sum;
            else if (sum != 0.f) atomicAdd(G.g_att_b2, sum);
        }
    }
    {
        // M = 64 accumulators: row i lives in lane 32 (i / 16) + i % 16, i.e. lanes 0..15 of every lane quarter;
        // warp group g = warp / 4 drains DW1 (g = 0) or DWp_{g-1}
        const int g = warp >> 2;
        const int drow = (warp & 3) * 16 + (lane & 15);
        float* slot = G.cta_partials != nullptr ? G.cta_partials + static_cast<size_t>(blockIdx.x) * kCtaPartialFloats + g * kC * kC : nullptr;
        float* dst = g == 0 ? G.g_att_w1 : (g - 1 < n_msgs ? G.g_wprod[g - 1] : nullptr);
        if (slot != nullptr && (g == 0 || g - 1 < n_msgs)) {
            const uint32_t src = (g == 0 ? tm_dw1 : tm_dwp + (g - 1) * 64) + lane_addr;
#pragma unroll 1
            for (int c8 = 0; c8 < 8; ++c8) {
                float v[8];
                tmem_ld8(src + c8 * 8, v);
                if (kSplit) {
                    float v2[8];
                    tmem_ld8(src + kSecondHalf + c8 * 8, v2);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] += v2[i];
                }
                if (lane < 16) {
                    float4* d8 = reinterpret_cast<float4*>(slot + drow * kC + c8 * 8);      // slots are 16-byte aligned (16704 floats apart)
                    d8[0] = make_float4(v[0], v[1], v[2], v[3]);
                    d8[1] = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
        } else if (dst != nullptr) {
            const bool vec_ok = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
            const uint32_t src = (g == 0 ? tm_dw1 : tm_dwp + (g - 1) * 64) + lane_addr;
#pragma unroll 1
            for (int c8 = 0; c8 < 8; ++c8) {
                float v[8];
                tmem_ld8(src + c8 * 8, v);
                if (kSplit) {
                    float v2[8];
                    tmem_ld8(src + kSecondHalf + c8 * 8, v2);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] += v2[i];
                }
                if (lane < 16) {
                    float* d8 = dst + drow * kC + c8 * 8;
                    if (vec_ok) {                          // 16-byte aligned accumulators: two 4-wide reductions
                        red_add_v4(d8, v[0], v[1], v[2], v[3]);
                        red_add_v4(d8 + 4, v[4], v[5], v[6], v[7]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) atomicAdd(d8 + i, v[i]);
                    }
                }
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

static unsigned long long* g_bwd_stamps = nullptr;
#if TOPO_DEBUG_KERNELS
// globaltimer stamps of CTA 0 / thread 0 (up to 62 x uint64, see the stamp() calls) for scripts/ablate_bwd.py
extern "C" void topo_debug_bwd_stamps(unsigned long long* device_buffer) { g_bwd_stamps = device_buffer; }
#endif

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
        TOPO_REQUIRE(p->agg[k] && p->w[k] && p->scale[k] && p->saved_m[k] && p->saved_pre[k] && g->g_agg[k] &&
                         (g->g_wprod[k] || g->cta_partials),
                     "null message operand (the fused backward needs the forward's saved activations)");
    TOPO_REQUIRE(p->att_w1 && p->att_w2 && (g->cta_partials || (g->g_att_w1 && g->g_att_b1 && g->g_att_w2 && g->g_att_b2)),
                 "null attention parameter");
    TOPO_REQUIRE(!p->apply_ln || (p->ln_gamma && (g->cta_partials || (g->g_ln_gamma && g->g_ln_beta))), "null LayerNorm parameter");
    TOPO_REQUIRE((reinterpret_cast<uintptr_t>(g->cta_partials) & 15) == 0, "cta_partials must be 16-byte aligned");
    if (rows == 0) return TOPO_OK;
    const size_t smem = BwdSmem::kTotal;
    const int grid = combine_grid(rows, p->max_ctas);
#define TOPO_LAUNCH_BWD(NM)                                                                                                  \
    do {                                                                                                                     \
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(combine_bwd_fused_kernel<NM>), smem)) return rc;      \
        combine_bwd_fused_kernel<NM><<<grid, kThreads, smem, as_stream(stream)>>>(*p, rows, n_rows_dev, grad_out, *g, g_bwd_stamps); \
    } while (0)
    if (p->n_msgs == 1) TOPO_LAUNCH_BWD(1);
    else if (p->n_msgs == 2) TOPO_LAUNCH_BWD(2);
    else TOPO_LAUNCH_BWD(3);
#undef TOPO_LAUNCH_BWD
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
