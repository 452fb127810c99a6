// Unit-test entry for the bf16x3 tcgen05 path (tc16.cuh): the three operand arrangements the fused
// backward kernel relies on, each against an fp64 product in tests/test_gpu_tc.py.
#include <algorithm>

#include "common.cuh"
#include "tc16.cuh"

namespace topo {
namespace {

using namespace tc16;

constexpr int kRows = 128, kC = 64;
constexpr uint32_t kPart = kRows * 128;        // 16 KB: one part of a [128 x 64] image
constexpr uint32_t kImg = 3 * kPart;           // 48 KB
constexpr uint32_t kWPart = kC * 128;          // 8 KB
constexpr uint32_t kWImg = 3 * kWPart;         // 24 KB

// rows [row0, row0 + 128) of a [rows, 64] fp32 matrix -> image; rows past `rows` are zero
__device__ __forceinline__ void stage_tile(const float* __restrict__ src, long long row0, long long rows, uint8_t* img,
                                           int tid, int nthreads) {
    for (int idx = tid; idx < kRows * 8; idx += nthreads) {
        const int r = idx >> 3, chunk = idx & 7;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (row0 + r < rows) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * kC) + chunk * 2);
            const float4 b = __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * kC) + chunk * 2 + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
        store_split8(img, kPart, r, chunk, v);
    }
}

__device__ __forceinline__ void stage_weight(const float* __restrict__ w, uint8_t* img, int tid, int nthreads) {
    for (int idx = tid; idx < kC * 8; idx += nthreads) {
        const int r = idx >> 3, chunk = idx & 7;
        const float4 a = __ldg(reinterpret_cast<const float4*>(w + r * kC) + chunk * 2);
        const float4 b = __ldg(reinterpret_cast<const float4*>(w + r * kC) + chunk * 2 + 1);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        store_split8(img, kWPart, r, chunk, v);
    }
}

// mode 0: out[rows, 64] = a @ w      (w stored [in][out]: the B operand is read MN-major)
// mode 1: out[rows, 64] = a @ w^T    (w stored [out][in]: K-major)
// mode 2: out[64, 64]  += a^T b      (both operands MN-major, contraction over the rows; `w` is b [rows, 64])
__global__ void __launch_bounds__(128) gemm16_debug_kernel(const float* __restrict__ a, const float* __restrict__ w,
                                                           long long rows, int mode, uint32_t mn_lbo, uint32_t mn_sbo,
                                                           uint32_t mn_kstep, float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* a_img = smem_raw;
    uint8_t* b_img = smem_raw + kImg;          // weight image (modes 0, 1) or the second tile (mode 2)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_smem, 64);
    if (mode < 2) stage_weight(w, b_img, tid, 128);
    fence_async_shared();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_smem;
    if (mode == 3) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        // tensor memory starts undefined: clear all 128 lanes x 64 columns so untouched lanes read as zero
        for (int c8 = 0; c8 < 8; ++c8) tmem_st8(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c8 * 8, z);
        tmem_st_wait();
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
    }
    uint32_t parity = 0, started = 0;
    const long long tiles = (rows + kRows - 1) / kRows;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row0 = tile * kRows;
        stage_tile(a, row0, rows, a_img, tid, 128);
        if (mode >= 2) stage_tile(w, row0, rows, b_img, tid, 128);
        fence_async_shared();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
            if (mode == 3) {
                // layout probe: the row contraction with M = 64 (where do the 64 accumulator rows land?)
                Operand oa = mn_major(smem_u32(a_img), kRows, mn_lbo), ob = mn_major(smem_u32(b_img), kRows, mn_lbo);
                gemm_bf16x3(tmem_base, oa, ob, idesc_bf16(64, 64, 1, 1), kRows / 16, started);
            } else if (mode == 2) {
                Operand oa = mn_major(smem_u32(a_img), kRows, mn_lbo), ob = mn_major(smem_u32(b_img), kRows, mn_lbo);
                oa.sbo = ob.sbo = mn_sbo;
                oa.kstep = ob.kstep = mn_kstep;
                gemm_bf16x3(tmem_base, oa, ob, idesc_bf16(128, 64, 1, 1), kRows / 16, started);
            } else if (mode == 0) {
                Operand ob = mn_major(smem_u32(b_img), kC, mn_lbo);
                ob.sbo = mn_sbo;
                ob.kstep = mn_kstep;
                gemm_bf16x3(tmem_base, k_major(smem_u32(a_img), kRows), ob, idesc_bf16(128, 64, 0, 1), kC / 16, 0);
            } else {
                gemm_bf16x3(tmem_base, k_major(smem_u32(a_img), kRows), k_major(smem_u32(b_img), kC), idesc_bf16(128, 64, 0, 0),
                            kC / 16, 0);
            }
            mma_commit(&bar);
        }
        started = 1;
        mbar_wait(&bar, parity);
        parity ^= 1;
        tc_fence_after_sync();
        if (mode < 2) {
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
        }
        tc_fence_before_sync();
        __syncthreads();
    }
    if (mode == 3 && started) {
        // dump every lane: out[128, 64]
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        tc_fence_after_sync();
        for (int c8 = 0; c8 < 8; ++c8) {
            float v[8];
            tmem_ld8(taddr + c8 * 8, v);
            for (int i = 0; i < 8; ++i) atomicAdd(out + tid * kC + c8 * 8 + i, v[i]);
        }
    }
    if (mode == 2 && started && tid < 64) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        tc_fence_after_sync();
#pragma unroll 1
        for (int c8 = 0; c8 < 8; ++c8) {
            float v[8];
            tmem_ld8(taddr + c8 * 8, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) atomicAdd(out + tid * kC + c8 * 8 + i, v[i]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// ---------------------------------------------------------------------------------------------
// mode 4: the same product as mode 1 (out = a @ w^T, w stored [out][in]) issued as tcgen05.mma.cta_group::2 by a
// CLUSTER OF TWO CTAs.  Each CTA stages its own 128-row tile of `a` and only HALF of the weight image (the 32
// output columns 32 * rank .. 32 * rank + 31): the pair shares the B operand, the M = 256 product leaves each CTA's
// 128 rows in its own tensor memory.  Groundwork for halving the resident weight images of the combine kernels
// (DESIGN.md section 7); pinned by tests/test_gpu_tc.py.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t cta_rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_bar)), "r"(cta_rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) gemm16_pair_kernel(const float* __restrict__ a,
                                                                                   const float* __restrict__ w, long long rows,
                                                                                   float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* a_img = smem_raw;                 // this CTA's 128 rows (3 x 16 KB)
    uint8_t* b_img = smem_raw + kImg;          // this CTA's 32 of the 64 weight rows (3 x 4 KB)
    constexpr uint32_t kHalfPart = 32 * 128;
    __shared__ __align__(8) uint64_t full_bar;  // leader's: both CTAs have staged their operands
    __shared__ __align__(8) uint64_t mma_bar;   // in each CTA: the pair's MMAs are complete
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const long long tile = blockIdx.x;          // consecutive CTAs form a pair: tiles 2p and 2p + 1
    const long long row0 = tile * kRows;

    if (tid == 0) {
        mbar_init(&full_bar, 2);
        mbar_init(&mma_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    stage_tile(a, row0, rows, a_img, tid, 128);
    for (int idx = tid; idx < 32 * 8; idx += 128) {
        const int n = idx >> 3, chunk = idx & 7;
        const float* src = w + (32 * rank + n) * kC;
        const float4 x0 = __ldg(reinterpret_cast<const float4*>(src) + chunk * 2);
        const float4 x1 = __ldg(reinterpret_cast<const float4*>(src) + chunk * 2 + 1);
        const float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        store_split8(b_img, kHalfPart, n, chunk, v);
    }
    fence_async_shared();
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                         // both barriers are initialised before anybody arrives remotely
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_smem;
    if (tid == 0) mbar_arrive_remote(&full_bar, 0);      // "my operands are staged" -> the leader's barrier
    if (rank == 0 && tid == 0) {
        uint32_t spins = 0;
        while (!mbar_try_wait_cluster(&full_bar, 0)) {
            if (++spins > (1u << 24)) __trap();
        }
        tc_fence_after_sync();
        const uint64_t a0 = smem_desc_sw128(smem_u32(a_img), 16, 1024), b0 = smem_desc_sw128(smem_u32(b_img), 16, 1024);
        const uint32_t idesc = idesc_bf16(256, 64, 0, 0);
        uint32_t acc = 0;
        for (int pass = 0; pass < 6; ++pass) {
            const uint32_t pa = pass == 0 ? 2u : ((pass == 2 || pass == 3) ? 1u : 0u);
            const uint32_t pb = pass == 1 ? 2u : ((pass == 2 || pass == 4) ? 1u : 0u);
            for (int k = 0; k < 4; ++k) {
                mma_bf16_pair(tmem_base, a0 + ((pa * kPart + k * 32) >> 4), b0 + ((pb * kHalfPart + k * 32) >> 4), idesc, acc);
                acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&mma_bar)), "h"(static_cast<uint16_t>(3)) : "memory");
    }
    {
        uint32_t spins = 0;
        while (!mbar_try_wait_cluster(&mma_bar, 0)) {
            if (++spins > (1u << 24)) __trap();
        }
    }
    tc_fence_after_sync();
    {
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
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                         // nobody frees tensor memory while the peer may still be reading its half
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
}

}  // namespace
}  // namespace topo

using namespace topo;

extern "C" int topo_debug_gemm_bf16x3(const float* a, const float* w, int64_t rows, int mode, int mn_lbo, int mn_sbo,
                                      int mn_kstep, float* out, topo_stream_t stream) {
    TOPO_REQUIRE(a && w && out && rows >= 0, "bad argument");
    TOPO_REQUIRE(mode >= 0 && mode <= 4, "mode must be 0..4");
    if (rows == 0) return TOPO_OK;
    if (mode == 4) {
        // one 128-row tile per CTA, CTAs in pairs (an odd last tile gets an all-padding partner)
        const size_t smem_pair = kImg + 3 * 32 * 128 + 1024;
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(gemm16_pair_kernel), smem_pair)) return rc;
        const int n_tiles = static_cast<int>((rows + kRows - 1) / kRows);
        gemm16_pair_kernel<<<(n_tiles + 1) / 2 * 2, 128, smem_pair, as_stream(stream)>>>(a, w, rows, out);
        TOPO_LAUNCH_CHECK();
        return TOPO_OK;
    }
    const size_t smem = 2 * kImg + 1024;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(gemm16_debug_kernel), smem)) return rc;
    const int tiles = static_cast<int>((rows + kRows - 1) / kRows);
    gemm16_debug_kernel<<<std::min(tiles, sm_count()), 128, smem, as_stream(stream)>>>(
        a, w, rows, mode, static_cast<uint32_t>(mn_lbo), static_cast<uint32_t>(mn_sbo), static_cast<uint32_t>(mn_kstep), out);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
