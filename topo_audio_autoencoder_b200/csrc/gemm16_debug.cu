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

}  // namespace
}  // namespace topo

using namespace topo;

extern "C" int topo_debug_gemm_bf16x3(const float* a, const float* w, int64_t rows, int mode, int mn_lbo, int mn_sbo,
                                      int mn_kstep, float* out, topo_stream_t stream) {
    TOPO_REQUIRE(a && w && out && rows >= 0, "bad argument");
    TOPO_REQUIRE(mode >= 0 && mode <= 3, "mode must be 0..3");
    if (rows == 0) return TOPO_OK;
    const size_t smem = 2 * kImg + 1024;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(gemm16_debug_kernel), smem)) return rc;
    const int tiles = static_cast<int>((rows + kRows - 1) / kRows);
    gemm16_debug_kernel<<<std::min(tiles, sm_count()), 128, smem, as_stream(stream)>>>(
        a, w, rows, mode, static_cast<uint32_t>(mn_lbo), static_cast<uint32_t>(mn_sbo), static_cast<uint32_t>(mn_kstep), out);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
