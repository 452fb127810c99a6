// S1 (b) on the 5th-generation tensor cores: the dense GEMMs of the SCCN message combine as
// tcgen05.mma kind::tf32 with 3xTF32 operand splitting (fp32-level accuracy), accumulators in tensor
// memory, one 128-row tile of target simplices per CTA iteration.  See tc.cuh for the primitives.
#include <algorithm>

#include "common.cuh"
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

// out[rows, 64] = a[rows, 64] @ w[64, 64]   (debug / unit-test entry for the tcgen05 path)
__global__ void __launch_bounds__(128) debug_gemm_kernel(const float* __restrict__ a, const float* __restrict__ w,
                                                         long long rows, float* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
    if (warp == 0) tmem_alloc(&tmem_base_smem, 64);
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
        fence_async_shared();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
            gemm_128x64x64_3xtf32(tmem_base, smem_u32(a_hi), smem_u32(a_lo), smem_u32(b_hi), smem_u32(b_lo), 0);
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
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}


// ---------------------------------------------------------------------------------------------
// Fused combine forward on tensor cores (C = 64).  256 threads: warps 0-3 are "row threads" (thread t
// owns tile row t == TMEM lane t and runs the epilogues), warps 4-7 are "loader threads" (stream the
// aggregate tiles global -> registers -> split -> swizzled shared memory, one message ahead); thread 128
// issues every MMA.  Tensor memory: T_k (conv outputs, 3 x 64 columns, kept for the final mix) and H
// (attention hidden layer, 64 columns).  Per message:
//     MMA1  T_k = agg_k W_k                    (24 x tcgen05.mma 128x64x8)
//     epi1  m_k = scale_k T_k + x  -> A operand (row threads, TMEM -> registers -> smem)
//     MMA2  H   = m_k W1^T
//     epi2  s_k = w2 . GELU(H + b1) + b2
// then  out = sum_k softmax(s)_k scale_k T_k + x,  LayerNorm,  store.
// ---------------------------------------------------------------------------------------------
struct FwdSmem {
    static constexpr uint32_t kWeights = 0;                       // [4][hi, lo] x 16 KB
    static constexpr uint32_t kA = 8 * kBTile;                    // hi, lo x 32 KB
    static constexpr uint32_t kVecs = kA + 2 * kATile;            // b1, w2, gamma, beta
    static constexpr uint32_t kTotal = kVecs + 4 * kC * 4;
};

__global__ void __launch_bounds__(256, 1) combine_fwd_tc_kernel(topo_combine_params P, long long rows,
                                                                const int* __restrict__ n_rows_dev,
                                                                float* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_hi = base + FwdSmem::kA;
    uint8_t* a_lo = a_hi + kATile;
    float* vecs = reinterpret_cast<float*>(base + FwdSmem::kVecs);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;

    const long long live = n_rows_dev ? min(static_cast<long long>(*n_rows_dev), rows) : rows;
    const long long tiles = (live + kTileRows - 1) / kTileRows;
    if (blockIdx.x >= tiles) return;

    const int tid = threadIdx.x, warp = tid >> 5;
    const bool is_row = tid < 128;
    const int lt = tid & 127;                       // index inside the role group
    const int n_msgs = P.n_msgs;

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_smem, 256);
    for (int k = 0; k < n_msgs; ++k)
        load_weight_tiles(P.w[k], true, base + (2 * k) * kBTile, base + (2 * k + 1) * kBTile, tid, 256);
    load_weight_tiles(P.att_w1, false, base + 6 * kBTile, base + 7 * kBTile, tid, 256);
    for (int c = tid; c < kC; c += 256) {
        vecs[c] = __ldg(P.att_b1 + c);
        vecs[kC + c] = __ldg(P.att_w2 + c);
        vecs[2 * kC + c] = P.apply_ln ? __ldg(P.ln_gamma + c) : 1.f;
        vecs[3 * kC + c] = P.apply_ln ? __ldg(P.ln_beta + c) : 0.f;
    }
    float scale[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) scale[k] = k < n_msgs ? __ldg(P.scale[k]) : 0.f;
    const float b2 = __ldg(P.att_b2);
    fence_async_shared();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), w_s = smem_u32(base);
    uint32_t parity = 0;

    // 64 registers with a role-dependent meaning: loader threads keep 16 prefetched chunks of the next
    // aggregate tile in them, row threads keep their x row (and accumulate the output row into it)
    float4 buf[16];
    auto prefetch = [&](const float* __restrict__ src, long long row0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int idx = j * 128 + lt, r = idx >> 4, chunk = idx & 15;
            buf[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < live) buf[j] = __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * kC) + chunk);
        }
    };
    if (!is_row) prefetch(P.agg[0], static_cast<long long>(blockIdx.x) * kTileRows);

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row0 = tile * kTileRows;
        const long long row = row0 + lt;
        float sc[3] = {0.f, 0.f, 0.f};
        if (is_row) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                buf[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (P.x != nullptr && row < live) buf[q] = __ldg(reinterpret_cast<const float4*>(P.x + row * kC) + q);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k >= n_msgs) break;
            if (!is_row) {
                // stage the prefetched aggregate tile as the A operand, then start fetching the next one
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int idx = j * 128 + lt;
                    store_split(a_hi, a_lo, kTileRows, idx >> 4, idx & 15, buf[j]);
                }
                fence_async_shared();
                if (k + 1 < n_msgs) prefetch(P.agg[k + 1], row0);
                else if (tile + gridDim.x < tiles) prefetch(P.agg[0], (tile + gridDim.x) * kTileRows);
            }
            __syncthreads();                                                    // S1: A = agg_k is staged
            if (tid == 128) {
                tc_fence_after_sync();
                gemm_128x64x64_3xtf32(tmem_base + k * kC, a_hi_s, a_lo_s, w_s + (2 * k) * kBTile, w_s + (2 * k + 1) * kBTile, 0);
                mma_commit(&bar);
            }
            if (is_row) {
                mbar_wait(&bar, parity);
                tc_fence_after_sync();
                // epilogue 1: m_k = scale_k T_k + x, re-staged as the A operand of the attention GEMM
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float v[32];
                    tmem_ld32(lane_addr + k * kC + half * 32, v);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 xr = buf[half * 8 + q];
                        float4 m;
                        m.x = fmaf(scale[k], v[q * 4 + 0], xr.x);
                        m.y = fmaf(scale[k], v[q * 4 + 1], xr.y);
                        m.z = fmaf(scale[k], v[q * 4 + 2], xr.z);
                        m.w = fmaf(scale[k], v[q * 4 + 3], xr.w);
                        store_split(a_hi, a_lo, kTileRows, lt, half * 8 + q, m);
                    }
                }
                fence_async_shared();
                tc_fence_before_sync();
            }
            parity ^= 1;
            __syncthreads();                                                    // S2: A = m_k is staged
            if (tid == 128) {
                tc_fence_after_sync();
                gemm_128x64x64_3xtf32(tmem_base + 3 * kC, a_hi_s, a_lo_s, w_s + 6 * kBTile, w_s + 7 * kBTile, 0);
                mma_commit(&bar);
            }
            if (is_row) {
                mbar_wait(&bar, parity);
                tc_fence_after_sync();
                // epilogue 2: attention score of this message
                float s = 0.f;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float v[32];
                    tmem_ld32(lane_addr + 3 * kC + half * 32, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int col = half * 32 + i;
                        s = fmaf(gelu_exact(v[i] + vecs[col]), vecs[kC + col], s);
                    }
                }
                sc[k] = s + b2;
                tc_fence_before_sync();
            }
            parity ^= 1;
            __syncthreads();                                                    // S3: A and H are free again
        }
        if (is_row) {
            // softmax over the messages, mix straight from tensor memory, LayerNorm, store
            float mx = sc[0];
            for (int k = 1; k < n_msgs; ++k) mx = fmaxf(mx, sc[k]);
            float a[3], sum = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                a[k] = k < n_msgs ? expf(sc[k] - mx) : 0.f;
                sum += a[k];
            }
            // out = (sum_k a_k) x + sum_k (a_k scale_k) T_k, accumulated in place over the x registers
            float asum = 0.f, coef[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float ak = a[k] / sum;
                asum += ak;
                coef[k] = ak * scale[k];
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                buf[q].x *= asum; buf[q].y *= asum; buf[q].z *= asum; buf[q].w *= asum;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (k < n_msgs) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float v[32];
                        tmem_ld32(lane_addr + k * kC + half * 32, v);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4& o = buf[half * 8 + q];
                            o.x = fmaf(coef[k], v[q * 4 + 0], o.x);
                            o.y = fmaf(coef[k], v[q * 4 + 1], o.y);
                            o.z = fmaf(coef[k], v[q * 4 + 2], o.z);
                            o.w = fmaf(coef[k], v[q * 4 + 3], o.w);
                        }
                    }
                }
            }
            tc_fence_before_sync();
            if (P.apply_ln) {
                float mean = 0.f;
#pragma unroll
                for (int q = 0; q < 16; ++q) mean += (buf[q].x + buf[q].y) + (buf[q].z + buf[q].w);
                mean *= (1.0f / kC);
                float var = 0.f;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    var = fmaf(buf[q].x - mean, buf[q].x - mean, var);
                    var = fmaf(buf[q].y - mean, buf[q].y - mean, var);
                    var = fmaf(buf[q].z - mean, buf[q].z - mean, var);
                    var = fmaf(buf[q].w - mean, buf[q].w - mean, var);
                }
                const float rstd = 1.0f / sqrtf(var * (1.0f / kC) + P.ln_eps);
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float* gm = vecs + 2 * kC + q * 4;
                    const float* bt = vecs + 3 * kC + q * 4;
                    buf[q].x = fmaf((buf[q].x - mean) * rstd, gm[0], bt[0]);
                    buf[q].y = fmaf((buf[q].y - mean) * rstd, gm[1], bt[1]);
                    buf[q].z = fmaf((buf[q].z - mean) * rstd, gm[2], bt[2]);
                    buf[q].w = fmaf((buf[q].w - mean) * rstd, gm[3], bt[3]);
                }
            }
            if (row < live) {
#pragma unroll
                for (int q = 0; q < 16; ++q) *reinterpret_cast<float4*>(out + row * kC + q * 4) = buf[q];
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace
}  // namespace topo

using namespace topo;

extern "C" int topo_debug_gemm_tf32x3(const float* a, const float* w, int64_t rows, float* out, topo_stream_t stream) {
    TOPO_REQUIRE(a && w && out && rows >= 0, "bad argument");
    if (rows == 0) return TOPO_OK;
    const size_t smem = 2 * kATile + 2 * kBTile + 1024;
    TOPO_CUDA(cudaFuncSetAttribute(debug_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int tiles = static_cast<int>((rows + kTileRows - 1) / kTileRows);
    debug_gemm_kernel<<<std::min(tiles, sm_count()), 128, smem, as_stream(stream)>>>(a, w, rows, out);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

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
    TOPO_CUDA(cudaFuncSetAttribute(combine_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int tiles = static_cast<int>((rows + kTileRows - 1) / kTileRows);
    combine_fwd_tc_kernel<<<std::min(tiles, sm_count()), 256, smem, as_stream(stream)>>>(*p, rows, n_rows_dev, out);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
