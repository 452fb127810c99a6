// S1 (b) forward on the tensor cores, second generation (C = 64): bf16x3 operand images (tc16.cuh), one MMA
// round per message with NO dependent GEMM chain, operand slots and accumulators multi-buffered so the
// tensor pipe works underneath the epilogue of the previous message.
//
// The attention hidden layer of message k is  pre_k = W1 m_k + b1  with  m_k = s_k (agg_k W_k) + x.  Folding
// the two products,  pre_k = agg_k (s_k W_k W1^T) + (x W1^T + b1) = agg_k V_k + H0 + b1,  both GEMMs of a message
// read the SAME operand agg_k:   [ D1_k | H_k ] = agg_k [ W_k | V_k ]   is one 128 x 128 x 64 product, and
// H0 = x W1^T is one more product per tile.  V_k (64 x 64) is formed once per CTA in fp32.
//
// Units of a tile: X (stages x, issues H0), then one unit per message (stages agg_k, issues [D1|H]).  Staging is a
// stream of its own that runs ahead of the epilogues ACROSS tile borders: before the epilogue of message k the
// workers stage the next unit of the stream (the next message, or the next tile's X), and with a residual one more
// before the tile end (the next tile's first message), so every operand load and every MMA has an epilogue or the tile
// end to hide behind.  Tensor memory holds three [D1|H] accumulators (the tile end re-reads the last two messages
// while the next tile's first one is being written) and two H0 buffers: all 512 columns.
//   stage    (chunk map)  thread t owns columns [8c, 8c+8), c = t % 8, of tile rows t/8 and t/8 + 64: coalesced
//                         128-bit loads, prefetched one unit ahead into registers, split into the bf16x3 image
//   epilogue (row map)    thread (q, r) = (t / 128, t % 128) owns columns [16q, 16q+16) of row r = TMEM lane r:
//                         m_k = s_k D1 + x,  pre_k = H + H0 + b1,  partial score  sum GELU(pre) w2
//   tile end              scores meet in shared memory, softmax, mix, LayerNorm (per-thread sum / squared deviations,
//                         merged with one exchange), store
// m_k, pre_k and the scores are SAVED for the backward in the tile-fragment layout (layout.cuh): every store and
// every later load of these private tensors is a fully coalesced 128-bit access from the row map.
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

constexpr bool kRowT = (TOPO_ROWMAP_TRANSPOSE & 1) != 0;

constexpr int kTileRows = 128;
constexpr int kC = 64;
constexpr int kWorkers = 512;                             // 16 warps of stage / epilogue threads
constexpr int kThreads = kWorkers + 32;                   // + one warp that only issues the MMAs
constexpr int kCW = 16;
constexpr uint32_t kPart = kTileRows * 128;               // 16 KB
constexpr uint32_t kImg = 3 * kPart;                      // 48 KB
constexpr uint32_t kWPart = kC * 128;                     // 8 KB
constexpr uint32_t kWImg = 3 * kWPart;                    // 24 KB

struct FwdSmem16 {
    static constexpr uint32_t kA0 = 0;                    // operand slot 0
    static constexpr uint32_t kW1 = kImg;                 // W1 [o][i]
    static constexpr uint32_t kWV = kW1 + kWImg;          // per message: W_k [in][out] then V_k [in][hidden] (48 KB)
    static constexpr uint32_t kA1 = kWV + 2 * 2 * kWImg;  // operand slot 1 = the third message's weights (n_msgs < 3 only)
    static constexpr uint32_t kVec = kWV + 3 * 2 * kWImg; // b1, w2
    static constexpr uint32_t kRed = kVec + 2 * kC * 4;   // [3 score slots + sum + squared deviations][4 column groups][128 rows]
    static constexpr uint32_t kBar = kRed + 5 * 4 * kTileRows * 4;
    static constexpr uint32_t kTotal = kBar + 32;
};
static_assert(FwdSmem16::kA1 % 1024 == 0 && FwdSmem16::kTotal <= 232448, "shared-memory layout");

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

__device__ __forceinline__ void load_chunk(const float* __restrict__ src, long long row, int chunk, bool ok, bool once, float (&v)[8]) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (ok) {
        const float4* p = reinterpret_cast<const float4*>(src + row * kC) + chunk * 2;
        if (once) {                   // the aggregates are read once here (and once in the backward, far later)
            a = ldg_pinned_once(p);
            b = ldg_pinned_once(p + 1);
        } else {                      // x is read again by the row map
            a = ldg_pinned(p);
            b = ldg_pinned(p + 1);
        }
    }
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__global__ void __launch_bounds__(kThreads, 1) combine_fwd16_kernel(topo_combine_params P, long long rows,
                                                                    const int* __restrict__ n_rows_dev,
                                                                    float* __restrict__ out, int dbg, unsigned long long* __restrict__ stamps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw;
    float* vecs = reinterpret_cast<float*>(base + FwdSmem16::kVec);
    float* red = reinterpret_cast<float*>(base + FwdSmem16::kRed);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + FwdSmem16::kBar);       // one per operand slot
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(base + FwdSmem16::kBar + 16);

    // differential-timing knobs and the in-kernel timeline of scripts/ablate_fwd16.py exist only in builds with
    // -DTOPO_DEBUG_KERNELS=1 (build.py reads the environment variable of that name)
#if !TOPO_DEBUG_KERNELS
    dbg = 0;
#endif
    auto stamp = [&](int slot) {
#if TOPO_DEBUG_KERNELS
        if (stamps != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            stamps[slot] = t;
        }
#else
        (void)slot;
#endif
    };
    stamp(0);
    const long long live = n_rows_dev ? min(static_cast<long long>(*n_rows_dev), rows) : rows;
    const long long tiles = (live + kTileRows - 1) / kTileRows;
    if (blockIdx.x >= tiles) return;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = tid >> 7, r = tid & 127, col0 = q * kCW;     // row map
    const int c = tid & 7, ra = tid >> 3;                       // chunk map: rows ra and ra + 64
    const int n_msgs = P.n_msgs;
    const bool has_x = P.x != nullptr;
    const int n_units = n_msgs + (has_x ? 1 : 0);
    const int n_slots = n_msgs < 3 ? 2 : 1;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_base_smem, 512);
    const bool worker = tid < kWorkers;
    if (P.weight_images != nullptr) {
        // the images were built once for the whole layer (weight_images.cu): one coalesced copy
        const uint4* __restrict__ src = reinterpret_cast<const uint4*>(P.weight_images);
        uint4* dst = reinterpret_cast<uint4*>(base + FwdSmem16::kW1);
        const int n16 = static_cast<int>((kWImg + n_msgs * 2 * kWImg) / 16);
        for (int idx = tid; idx < n16; idx += kThreads) dst[idx] = __ldg(src + idx);
    } else {
    // weights: W1 as stored; per message W_k as stored and V_k = s_k W_k W1^T (thread t: row i = t/8, columns 8c..8c+7)
    for (int idx = tid; idx < kC * 8; idx += kThreads) {
        const int i = idx >> 3, ch = idx & 7;
        const float4 a = __ldg(reinterpret_cast<const float4*>(P.att_w1 + i * kC) + ch * 2);
        const float4 b = __ldg(reinterpret_cast<const float4*>(P.att_w1 + i * kC) + ch * 2 + 1);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        store_split8(base + FwdSmem16::kW1, kWPart, i, ch, v);
    }
    {
        // V_k[i][h] = s_k sum_o W_k[i][o] W1[h][o].  Thread t keeps row h = t % 64 of W1 in registers and forms
        // V_k[8 (t / 64) + e][h], e = 0..7 (the W_k rows are warp-uniform broadcast loads); the 64 x 64 result
        // crosses to the image writers through a staging tile in the still idle operand slot 0.
        const int h = tid & 63, ib = tid >> 6;
        float w1row[kC];
#pragma unroll
        for (int o4 = 0; o4 < kC / 4; ++o4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(P.att_w1 + h * kC) + o4);
            w1row[4 * o4] = t.x; w1row[4 * o4 + 1] = t.y; w1row[4 * o4 + 2] = t.z; w1row[4 * o4 + 3] = t.w;
        }
        float* vs = reinterpret_cast<float*>(base + FwdSmem16::kA0);
        for (int k = 0; k < n_msgs; ++k) {
            uint8_t* w_img = base + FwdSmem16::kWV + k * 2 * kWImg;
            uint8_t* v_img = w_img + kWImg;
            const float sk = __ldg(P.scale[k]);
            const float* __restrict__ wk = P.w[k];
            if (worker) {
                const int i = tid >> 3;
                const float4 a = __ldg(reinterpret_cast<const float4*>(wk + i * kC) + c * 2);
                const float4 b = __ldg(reinterpret_cast<const float4*>(wk + i * kC) + c * 2 + 1);
                const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                store_split8(w_img, kWPart, i, c, v);
            }
#pragma unroll 1
            for (int e = 0; e < 8 && worker; ++e) {
                const float4* __restrict__ wrow = reinterpret_cast<const float4*>(wk + (8 * ib + e) * kC);
                float acc = 0.f;
#pragma unroll
                for (int o4 = 0; o4 < kC / 4; ++o4) {
                    const float4 t = __ldg(wrow + o4);
                    acc = fmaf(t.x, w1row[4 * o4], acc);
                    acc = fmaf(t.y, w1row[4 * o4 + 1], acc);
                    acc = fmaf(t.z, w1row[4 * o4 + 2], acc);
                    acc = fmaf(t.w, w1row[4 * o4 + 3], acc);
                }
                vs[(8 * ib + e) * kC + h] = sk * acc;
            }
            __syncthreads();
            if (worker) {
                const int i = tid >> 3;
                const float4 a = *reinterpret_cast<const float4*>(vs + i * kC + 8 * c);
                const float4 b = *reinterpret_cast<const float4*>(vs + i * kC + 8 * c + 4);
                const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                store_split8(v_img, kWPart, i, c, v);
            }
            __syncthreads();
        }
    }
    }   // images built in place
    for (int i = tid; i < kC; i += kThreads) {
        vecs[i] = __ldg(P.att_b1 + i);
        vecs[kC + i] = __ldg(P.att_w2 + i);
    }
    stamp(1);
    fence_async_shared();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    stamp(2);
    const uint32_t tmem_base = *tmem_base_smem;
    // tensor-memory columns: three [D1 | H] accumulators, then two H0 buffers (512 in all)
    const uint32_t tm_h0 = tmem_base + 384;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t slot_s[2] = {smem_u32(base + FwdSmem16::kA0), smem_u32(base + FwdSmem16::kA1)};
    uint8_t* slot_p[2] = {base + FwdSmem16::kA0, base + FwdSmem16::kA1};
    const uint32_t w1_s = smem_u32(base + FwdSmem16::kW1), wv_s = smem_u32(base + FwdSmem16::kWV);
    const float b2 = __ldg(P.att_b2);
    float scale_r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) scale_r[k] = k < n_msgs ? __ldg(P.scale[k]) : 0.f;

    if (!worker) {
        // ================= MMA warp: waits for each staged unit, issues its product, commits to the slot's mbarrier =================
        // Units arrive in tile order (x, then the messages); message products rotate through three accumulators and the
        // x products through two, so that a unit can be issued while the epilogues of the previous tile still read theirs.
        uint32_t unit_no = 0, acc = 0, tile_no = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tile_no) {
#pragma unroll 1
            for (int u = 0; u < n_units; ++u, ++unit_no) {
                const int slot = n_slots == 2 ? static_cast<int>(unit_no & 1u) : 0;
                named_bar_sync(1 + (unit_no & 1u), kThreads);       // all 512 workers have staged this unit
                if (lane == 0) {
                    tc_fence_after_sync();
                    const int k = u - (has_x ? 1 : 0);
                    if (dbg & 16) {
                    } else if (k < 0) {
                        gemm_bf16x3_unrolled<kC / 16>(tm_h0 + (tile_no & 1u) * 64, k_major(slot_s[slot], kTileRows), k_major(w1_s, kC),
                                                      idesc_bf16(128, 64, 0, 0), 0);
                    } else {
                        gemm_bf16x3_unrolled<kC / 16>(tmem_base + acc * 128, k_major(slot_s[slot], kTileRows),
                                                      mn_major(wv_s + k * 2 * kWImg, kC, kWImg), idesc_bf16(128, 128, 0, 1), 0);
                    }
                    mma_commit(&bars[slot]);
                }
                __syncwarp();
                if (u >= (has_x ? 1 : 0)) acc = acc == 2 ? 0 : acc + 1;
            }
        }
    } else {
    // ================= workers =================
    uint32_t par0 = 0u, par1 = 0u;
    int pend0 = -1, pend1 = -1;     // number of the unit whose MMAs are in flight on operand slot 0 / 1 (-1: none)
    // wait until the MMAs that last read operand slot s (and wrote their accumulator) are complete
    auto drain = [&](int s) {
        if (s == 0) {
            if (pend0 >= 0) { mbar_wait_backoff(&bars[0], par0); par0 ^= 1u; pend0 = -1; tc_fence_after_sync(); }
        } else {
            if (pend1 >= 0) { mbar_wait_backoff(&bars[1], par1); par1 ^= 1u; pend1 = -1; tc_fence_after_sync(); }
        }
    };
    // source of unit u of a tile: x first (when present), then the aggregates
    auto unit_src = [&](int u) -> const float* {
        const int k = u - (has_x ? 1 : 0);
        return k < 0 ? P.x : (k == 0 ? P.agg[0] : (k == 1 ? P.agg[1] : P.agg[2]));
    };

    // ---- the staging stream: units in tile order, numbered from 0, running AHEAD of the epilogues across tile borders ----
    const int hx = has_x ? 1 : 0;
    float pf[2][8];                 // rows of the next unit to stage (chunk map), loaded when the previous one was staged
    long long s_tile = blockIdx.x;  // tile / unit of the next unit to stage
    int s_u = 0;
    int staged = 0;                 // units staged so far = number of the next one
    auto load_pf = [&]() {
        const long long nrow0 = s_tile * kTileRows;
        const float* __restrict__ src = unit_src(s_u);
#pragma unroll
        for (int j = 0; j < 2; ++j)
            load_chunk(src, nrow0 + ra + 64 * j, c, !(dbg & 8) && s_tile < tiles && (nrow0 + ra + 64 * j < live), s_u >= hx, pf[j]);
    };
    auto stage_next = [&]() {
        if (s_tile >= tiles) return;
        const int slot = n_slots == 2 ? (staged & 1) : 0;
        drain(slot);
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (!(dbg & 4)) store_split8(slot_p[slot], kPart, ra + 64 * j, c, pf[j]);
        fence_async_shared();
        tc_fence_before_sync();
        named_bar_arrive(1 + (staged & 1), kThreads);      // hand over to the MMA warp: arrive, do not wait
        if (slot == 0) pend0 = staged; else pend1 = staged;
        ++staged;
        if (++s_u == n_units) { s_u = 0; s_tile += gridDim.x; }
        load_pf();
    };
    load_pf();
    stage_next();
    if (has_x) stage_next();

    uint32_t acc0 = 0;              // accumulator of this tile's first message
    uint32_t tile_no = 0;
    float xr[kCW];                  // row map: this thread's slice of the residual row
    float4 xraw[4];                 // the same as it arrives from global memory (64 contiguous bytes per row and access, layout.cuh)
    {
        const long long row = static_cast<long long>(blockIdx.x) * kTileRows + r;
        rowmap_load16_issue<kRowT>(P.x, row, col0, live, has_x && !(dbg & 256), xraw, lane, [](const float4* p) { return ldg_pinned(p); });
    }
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tile_no) {
        const long long row0 = tile * kTileRows;
        const long long row = row0 + r;
        const bool row_alive = row < live;
        rowmap_load16_finish<kRowT>(xraw, xr, lane);              // issued a tile ago
        // every line this CTA reads two tiles from now goes to L2 (the register loads run up to a tile ahead)
        {
            const long long prow0 = (tile + 2 * static_cast<long long>(gridDim.x)) * kTileRows;
            if ((c & 3) == 0 && !(dbg & 8)) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const long long prow = prow0 + ra + 64 * j;
                    if (prow < live) {
                        for (int u = 0; u < n_units; ++u) prefetch_l2_line(unit_src(u) + prow * kC + 8 * c);
                    }
                }
            }
        }
        const uint32_t tm_h0_t = tm_h0 + (tile_no & 1u) * 64 + lane_addr + col0;
        const int id0 = static_cast<int>(tile_no) * n_units + hx;       // unit number of this tile's first message
#pragma unroll 1
        for (int k = 0; k < n_msgs; ++k) {
            stage_next();           // keeps the stream one unit ahead of this epilogue (two with a residual)
            // ---- epilogue of message k ----
            {
                const int id = id0 + k;
                const int es = n_slots == 2 ? (id & 1) : 0;
                if ((es == 0 ? pend0 : pend1) == id) drain(es);
                uint32_t ak = acc0 + k;
                if (ak >= 3) ak -= 3;
                const uint32_t tm = tmem_base + ak * 128 + lane_addr + col0;
                const float sk = k == 0 ? scale_r[0] : (k == 1 ? scale_r[1] : scale_r[2]);
                // the message first, so that its registers are free again before the GELU evaluations need theirs
                {
                    float m[kCW];
                    if (!(dbg & 64)) {
                        tmem_ld16(tm, m);
                    } else {
#pragma unroll
                        for (int i = 0; i < kCW; ++i) m[i] = 0.01f * i;
                    }
#pragma unroll
                    for (int i = 0; i < kCW; ++i) m[i] = fmaf(sk, m[i], xr[i]);
                    float4* pm = reinterpret_cast<float4*>(P.saved_m[k]) + tf_index(row0, q, r);
                    if (row_alive && !(dbg & 1)) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) __stcs(pm + j * kTileRows, make_float4(m[4 * j], m[4 * j + 1], m[4 * j + 2], m[4 * j + 3]));
                    }
                }
                float h[kCW];
                if (!(dbg & 64)) {
                    tmem_ld16(tm + 64, h);
                } else {
#pragma unroll
                    for (int i = 0; i < kCW; ++i) h[i] = 0.01f * i;
                }
                if (has_x && !(dbg & 64)) {
                    float h0[kCW];
                    tmem_ld16(tm_h0_t, h0);
#pragma unroll
                    for (int i = 0; i < kCW; ++i) h[i] += h0[i];
                }
#pragma unroll
                for (int i = 0; i < kCW; ++i) h[i] += vecs[col0 + i];          // pre-GELU hidden activation
                float4* pp = reinterpret_cast<float4*>(P.saved_pre[k]) + tf_index(row0, q, r);
                if (row_alive && !(dbg & 1)) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) __stcs(pp + j * kTileRows, make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]));
                }
                float part = 0.f;
#pragma unroll
                for (int i = 0; i < kCW; i += 4) {
                    float cdf[4], pdf[4];
                    gelu_cdf_pdf_n<4>(h + i, cdf, pdf);
#pragma unroll
                    for (int e = 0; e < 4; ++e) part = fmaf((dbg & 2) ? h[i + e] : h[i + e] * cdf[e], vecs[kC + col0 + i + e], part);
                }
                red[(k * 4 + q) * kTileRows + r] = part;
                tc_fence_before_sync();
            }
        }
        // with a residual the stream is two units ahead: the next tile's first message is issued underneath the tile end
        if (has_x) stage_next();
        // ---------------- tile end: softmax over the messages, mix, LayerNorm ----------------
        if (tile == blockIdx.x) stamp(3);
        named_bar_sync(3, kWorkers);                       // the partial scores are in `red`
        float att[3];
        {
            float sc[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float* rk = red + k * 4 * kTileRows;
                sc[k] = k < n_msgs ? b2 + ((rk[r] + rk[kTileRows + r]) + (rk[2 * kTileRows + r] + rk[3 * kTileRows + r])) : 0.f;
                if (k < n_msgs && q == 0 && row_alive) P.saved_score[k * rows + row] = sc[k];
            }
            float mx = sc[0];
            if (n_msgs > 1) mx = fmaxf(mx, sc[1]);
            if (n_msgs > 2) mx = fmaxf(mx, sc[2]);
            const float e0 = expf(sc[0] - mx), e1 = n_msgs > 1 ? expf(sc[1] - mx) : 0.f, e2 = n_msgs > 2 ? expf(sc[2] - mx) : 0.f;
            const float es = e0 + e1 + e2;
            att[0] = e0 / es; att[1] = e1 / es; att[2] = e2 / es;
        }
        float y[kCW];
#pragma unroll
        for (int i = 0; i < kCW; ++i) y[i] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < n_msgs) {
                if (k + 2 >= n_msgs) {
                    // the last two messages' products are still in tensor memory (an accumulator is only overwritten by
                    // the third message after it): m_k = s_k D1_k + x again, without a trip to L2
                    float d1[kCW];
                    uint32_t ak = acc0 + k;
                    if (ak >= 3) ak -= 3;
                    tmem_ld16(tmem_base + ak * 128 + lane_addr + col0, d1);
                    const float sk = k == 0 ? scale_r[0] : (k == 1 ? scale_r[1] : scale_r[2]);
#pragma unroll
                    for (int i = 0; i < kCW; ++i) y[i] = fmaf(att[k], fmaf(sk, d1[i], xr[i]), y[i]);
                } else {
                    // this thread wrote exactly these addresses a moment ago (same thread, program order)
                    const float4* pm = reinterpret_cast<const float4*>(P.saved_m[k]) + tf_index(row0, q, r);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 t = (row_alive && !(dbg & 32)) ? pm[j * kTileRows] : make_float4(0.f, 0.f, 0.f, 0.f);
                        y[4 * j] = fmaf(att[k], t.x, y[4 * j]);
                        y[4 * j + 1] = fmaf(att[k], t.y, y[4 * j + 1]);
                        y[4 * j + 2] = fmaf(att[k], t.z, y[4 * j + 2]);
                        y[4 * j + 3] = fmaf(att[k], t.w, y[4 * j + 3]);
                    }
                }
            }
        }
        tc_fence_before_sync();
        // the residual slice of this CTA's next tile (in L2 since two tiles ago) travels underneath the LayerNorm
        {
            const long long nrow = (tile + gridDim.x) * kTileRows + r;
            rowmap_load16_issue<kRowT>(P.x, nrow, col0, live, has_x && !(dbg & 256), xraw, lane, [](const float4* p) { return ldg_pinned(p); });
        }
        if (P.apply_ln) {
            // Row statistics with ONE exchange: every thread reduces its 16 columns to (sum, squared deviations from
            // its own mean) and the four partials of a row are merged exactly (Chan et al.): no E[y^2] - mean^2.
            float part = 0.f;
#pragma unroll
            for (int i = 0; i < kCW; ++i) part += y[i];
            const float lmean = part * (1.0f / kCW);
            float dev = 0.f;
#pragma unroll
            for (int i = 0; i < kCW; ++i) dev = fmaf(y[i] - lmean, y[i] - lmean, dev);
            red[(3 * 4 + q) * kTileRows + r] = part;
            red[(4 * 4 + q) * kTileRows + r] = dev;
            named_bar_sync(3, kWorkers);
            const float* rs = red + 3 * 4 * kTileRows;
            const float* rd = red + 4 * 4 * kTileRows;
            const float s0 = rs[r], s1 = rs[kTileRows + r], s2 = rs[2 * kTileRows + r], s3 = rs[3 * kTileRows + r];
            const float mean = ((s0 + s1) + (s2 + s3)) * (1.0f / kC);
            const float d0 = s0 * (1.0f / kCW) - mean, d1 = s1 * (1.0f / kCW) - mean, d2 = s2 * (1.0f / kCW) - mean, d3 = s3 * (1.0f / kCW) - mean;
            const float m2 = ((rd[r] + rd[kTileRows + r]) + (rd[2 * kTileRows + r] + rd[3 * kTileRows + r])) +
                             static_cast<float>(kCW) * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
            const float rstd = 1.0f / sqrtf(m2 * (1.0f / kC) + P.ln_eps);
#pragma unroll
            for (int j = 0; j < 4; ++j) {                  // gamma / beta: 64 floats, L1 resident
                const float4 gm = __ldg(reinterpret_cast<const float4*>(P.ln_gamma + col0) + j);
                const float4 bt = __ldg(reinterpret_cast<const float4*>(P.ln_beta + col0) + j);
                y[4 * j] = fmaf((y[4 * j] - mean) * rstd, gm.x, bt.x);
                y[4 * j + 1] = fmaf((y[4 * j + 1] - mean) * rstd, gm.y, bt.y);
                y[4 * j + 2] = fmaf((y[4 * j + 2] - mean) * rstd, gm.z, bt.z);
                y[4 * j + 3] = fmaf((y[4 * j + 3] - mean) * rstd, gm.w, bt.w);
            }
        } else {
            named_bar_sync(3, kWorkers);                   // without the exchange: the score slots must be read by everyone
        }
        rowmap_store16<kRowT>(out, row, col0, (dbg & 32) ? 0 : live, y, lane);       // 64 contiguous bytes per row and access
        // no barrier here: the score slots are not read after the exchange above, and the statistics slots are next
        // written behind the next tile's score barrier, which no thread passes before it has read them
        acc0 += static_cast<uint32_t>(n_msgs);
        while (acc0 >= 3) acc0 -= 3;
        if (tile == blockIdx.x) stamp(4);
    }
    }   // workers
    stamp(5);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
    stamp(6);
}

}  // namespace
}  // namespace topo

using namespace topo;

static int g_fwd16_debug_mask = 0;
static unsigned long long* g_fwd16_stamps = nullptr;
#if TOPO_DEBUG_KERNELS
// Differential-timing knob for scripts/ablate_fwd16.py (0 = the real kernel; results are WRONG for any other value).
extern "C" void topo_debug_fwd16_mask(int mask) { g_fwd16_debug_mask = mask; }
// globaltimer stamps of CTA 0 / thread 0 (7 x uint64): start, set-up done, after the set-up barrier, first tile's
// unit loop done, first tile done, all tiles done, end
extern "C" void topo_debug_fwd16_stamps(unsigned long long* device_buffer) { g_fwd16_stamps = device_buffer; }
#endif

extern "C" int topo_sccn_combine_fwd_tc2(const topo_combine_params* p, int64_t rows, const int32_t* n_rows_dev,
                                         float* out, topo_stream_t stream) {
    TOPO_REQUIRE(p && out && rows >= 0, "bad argument");
    TOPO_REQUIRE(p->n_msgs >= 1 && p->n_msgs <= 3, "n_msgs must be 1..3");
    if (p->channels != kC) {
        set_error("the tensor-core combine is instantiated for channels == 64");
        return TOPO_ERR_UNSUPPORTED;
    }
    for (int k = 0; k < p->n_msgs; ++k)
        TOPO_REQUIRE(p->agg[k] && p->w[k] && p->scale[k] && p->saved_m[k] && p->saved_pre[k],
                     "null message operand (this kernel always saves m_k and pre_k)");
    TOPO_REQUIRE(p->saved_score && p->saved_layout == TOPO_SAVED_TILE_FRAGMENT,
                 "saved activations must be requested in the tile-fragment layout");
    TOPO_REQUIRE(p->att_w1 && p->att_b1 && p->att_w2 && p->att_b2, "null attention parameter");
    TOPO_REQUIRE(!p->apply_ln || (p->ln_gamma && p->ln_beta), "null LayerNorm parameter");
    if (rows == 0) return TOPO_OK;
    const size_t smem = FwdSmem16::kTotal;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(combine_fwd16_kernel), smem)) return rc;
    combine_fwd16_kernel<<<combine_grid(rows, p->max_ctas), kThreads, smem, as_stream(stream)>>>(*p, rows, n_rows_dev, out, g_fwd16_debug_mask, g_fwd16_stamps);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
