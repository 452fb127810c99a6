// Operand images of the SCCN weights for the bf16x3 tensor-core kernels, built ONCE per layer and step.
//
// Every CTA of the combine kernels needs its rank's weights as bf16x3 SWIZZLE_128B images (tc16.cuh), plus, for
// the forward, the folded matrices V_k = s_k W_k W1^T.  Building them inside the kernels cost ~30 us of serial
// load latency per launch (48 launches per step).  This kernel writes them to global memory in exactly the byte
// layout the kernels keep in shared memory, so a CTA's set-up is one coalesced copy.
//   job with w:    dst = [ W_k image (24 KB) | V_k image (24 KB) ]
//   job without w: dst = [ W1 image (24 KB) ]
#include "common.cuh"
#include "tc16.cuh"

namespace topo {
namespace {

using namespace tc16;

constexpr int kC = 64;
constexpr uint32_t kWPart = kC * 128;
constexpr uint32_t kWImg = 3 * kWPart;
constexpr int kMaxJobs = 96;          // 96 x 32 bytes of kernel parameters: every layer of a 6-layer SCCN in one launch

struct ImageJobs {
    topo_image_job j[kMaxJobs];
};

__global__ void __launch_bounds__(512) prepare_images_kernel(const __grid_constant__ ImageJobs jobs) {
    __shared__ __align__(16) float w1s[kC][kC + 4];    // W1 [o][i], padded rows (bank-conflict-free column walks)
    __shared__ __align__(16) float wks[kC][kC + 4];    // W_k [in][out]
    const topo_image_job job = jobs.j[blockIdx.x];
    uint8_t* dst = static_cast<uint8_t*>(job.dst);
    const int tid = threadIdx.x, i = tid >> 3, c = tid & 7;
    // coalesced loads of both matrices (eight lanes = one 256-byte row)
    {
        const float4 a = __ldg(reinterpret_cast<const float4*>(job.att_w1 + i * kC) + c * 2);
        const float4 b = __ldg(reinterpret_cast<const float4*>(job.att_w1 + i * kC) + c * 2 + 1);
        *reinterpret_cast<float4*>(&w1s[i][8 * c]) = a;
        *reinterpret_cast<float4*>(&w1s[i][8 * c + 4]) = b;
        if (job.w == nullptr) {
            const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            store_split8(dst, kWPart, i, c, v);
            return;
        }
    }
    const float4 a = __ldg(reinterpret_cast<const float4*>(job.w + i * kC) + c * 2);
    const float4 b = __ldg(reinterpret_cast<const float4*>(job.w + i * kC) + c * 2 + 1);
    *reinterpret_cast<float4*>(&wks[i][8 * c]) = a;
    *reinterpret_cast<float4*>(&wks[i][8 * c + 4]) = b;
    {
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        store_split8(dst, kWPart, i, c, v);
    }
    const float sk = __ldg(job.scale);
    __syncthreads();
    // V[i][h] = s sum_o W_k[i][o] W1[h][o], h = 8c .. 8c+7
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 4
    for (int o = 0; o < kC; ++o) {
        const float w = wks[i][o];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, w1s[8 * c + e][o], acc[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= sk;
    store_split8(dst + kWImg, kWPart, i, c, acc);
}

struct WgradJobs {
    topo_wgrad_job j[kMaxJobs];
};

// One CTA of 1024 threads per job.  Elements are taken 256 at a time; with a partials buffer the four thread quarters each add
// every fourth CTA slot of an element (independent loads, fixed order) and meet in shared memory, again in a fixed order.
__global__ void __launch_bounds__(1024) finish_weight_grads_kernel(const __grid_constant__ WgradJobs jobs, int n_default, int grid_cap) {
    __shared__ float quarter[4][256];
    __shared__ float part[8];
    const topo_wgrad_job job = jobs.j[blockIdx.x];
    const int n = job.count > 0 ? job.count : n_default;
    const int e = threadIdx.x & 255, qd = threadIdx.x >> 8;
    const bool conv = job.w != nullptr;
    const float s = conv ? __ldg(job.scale) : 1.f;
    int n_active = 0;
    if (job.partials != nullptr) {
        const long long live = job.n_rows_dev ? min(static_cast<long long>(*job.n_rows_dev), static_cast<long long>(job.rows)) : job.rows;
        const long long tiles = (live + 127) / 128, launched = (job.rows + 127) / 128;
        const int cap = job.max_ctas > 0 ? min(job.max_ctas, grid_cap) : grid_cap;
        n_active = static_cast<int>(min(tiles, min(launched, static_cast<long long>(cap))));
    }
    float acc = 0.f;
    // 16 chunks of 256 elements at a time: every thread keeps 16 independent loads per CTA slot in flight (one chunk at a time the
    // kernel was a chain of slot loads: 84 us for a 64 x 64 matrix over 107 slots).  The order of every sum is unchanged.
    for (int i00 = 0; i00 < n; i00 += 16 * 256) {
        float t[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) t[c] = 0.f;
        if (job.partials != nullptr) {
            const float* src = job.partials + job.partial_offset + i00 + e;
            for (int b = qd; b < n_active; b += 4) {
                const float* sb = src + static_cast<size_t>(b) * kCtaPartialFloats;
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    if (i00 + c * 256 + e < n) t[c] += __ldcs(sb + c * 256);
            }
        }
#pragma unroll 1
        for (int c = 0; c < 16 && i00 + c * 256 < n; ++c) {
            const int i = i00 + c * 256 + e;
            float p = 0.f;
            if (job.partials != nullptr) {
                float tc = 0.f;      // select instead of t[c]: a runtime index would move the array to local memory
#pragma unroll
                for (int cc = 0; cc < 16; ++cc) tc = cc == c ? t[cc] : tc;
                quarter[qd][e] = tc;
                __syncthreads();
                p = (quarter[0][e] + quarter[1][e]) + (quarter[2][e] + quarter[3][e]);
                __syncthreads();
            } else if (i < n) {
                p = job.wprod[i];
            }
            if (qd == 0 && i < n) {
                job.g_w[i] = s * p;
                if (conv) acc = fmaf(p, __ldg(job.w + i), acc);
            }
        }
    }
    if (!conv) return;
    if (qd == 0) {
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += part[w];
        job.g_scale[0] = t;
    }
}

}  // namespace
}  // namespace topo

using namespace topo;

extern "C" int topo_sccn_finish_weight_grads(const topo_wgrad_job* jobs, int n_jobs, int channels, topo_stream_t stream) {
    TOPO_REQUIRE(jobs && n_jobs >= 0 && n_jobs <= kMaxJobs && channels > 0, "at most 96 jobs per call");
    if (n_jobs == 0) return TOPO_OK;
    WgradJobs packed{};
    for (int q = 0; q < n_jobs; ++q) {
        TOPO_REQUIRE((jobs[q].wprod || jobs[q].partials) && jobs[q].g_w, "null pointer in a weight-gradient job");
        TOPO_REQUIRE(jobs[q].w == nullptr || (jobs[q].scale && jobs[q].g_scale), "a conv-weight job needs scale and g_scale");
        TOPO_REQUIRE(jobs[q].count >= 0 && jobs[q].partial_offset >= 0 && jobs[q].partial_offset + jobs[q].count <= kCtaPartialFloats &&
                         jobs[q].rows >= 0, "bad partials geometry in a weight-gradient job");
        packed.j[q] = jobs[q];
    }
    finish_weight_grads_kernel<<<n_jobs, 1024, 0, as_stream(stream)>>>(packed, channels * channels, sm_count());
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_sccn_combine_grid(int64_t rows, int max_ctas) { return rows > 0 ? combine_grid(rows, max_ctas) : 0; }

extern "C" int topo_sccn_prepare_images(const topo_image_job* jobs, int n_jobs, int channels, topo_stream_t stream) {
    TOPO_REQUIRE(jobs && n_jobs >= 0 && n_jobs <= kMaxJobs, "at most 96 jobs per call");
    if (channels != kC) {
        set_error("weight images exist for channels == 64 (the tensor-core kernels)");
        return TOPO_ERR_UNSUPPORTED;
    }
    if (n_jobs == 0) return TOPO_OK;
    ImageJobs packed{};
    for (int q = 0; q < n_jobs; ++q) {
        TOPO_REQUIRE(jobs[q].att_w1 && jobs[q].dst && (jobs[q].w == nullptr || jobs[q].scale), "null pointer in an image job");
        TOPO_REQUIRE((reinterpret_cast<uintptr_t>(jobs[q].dst) & 15u) == 0, "image destinations must be 16-byte aligned");
        packed.j[q] = jobs[q];
    }
    prepare_images_kernel<<<n_jobs, 512, 0, as_stream(stream)>>>(packed);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
