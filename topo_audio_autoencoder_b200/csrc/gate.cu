// G1: stochastic gates over the flat [B, N] simplex axis.
//   - Hard Concrete sample / stretch / clamp / STE, forward and backward, fused, 128-bit HBM access.
//     (README.md:15-18 names the ingredients; the reference ships no code, spec in DESIGN.md.)
//   - BinaryGumbel training branch (encoder.py:34-41).
// HBM-bound streaming kernels: 12 B/element forward, 16 B/element backward.
#include "common.cuh"

namespace topo {
namespace {

struct RankOffsets {
    long long o[5];
};

__device__ __forceinline__ int rank_of(long long col, const RankOffsets& ro) {
    return (col >= ro.o[1]) + (col >= ro.o[2]) + (col >= ro.o[3]);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// One element of the forward.  Returns z; *s_out is the sigmoid, *inside whether the clamp passes
// gradient (torch.clamp: min <= x <= max).
__device__ __forceinline__ float hc_element(float logit, float u, float loc, float beta, float gamma,
                                            float zeta, bool training, float* s_out, bool* inside,
                                            float* x_out) {
    float x = logit + loc;
    if (training) x = (logf(u) - logf(1.0f - u) + x) / beta;
    const float s = sigmoidf_(x);
    const float sbar = s * (zeta - gamma) + gamma;
    *s_out = s;
    *x_out = x;
    *inside = (sbar >= 0.0f) && (sbar <= 1.0f);
    return fminf(fmaxf(sbar, 0.0f), 1.0f);
}

template <bool TRAINING>
__global__ void __launch_bounds__(256) hard_concrete_fwd_kernel(
    const float* __restrict__ logits, const float* __restrict__ u, const float* __restrict__ params,
    RankOffsets ro, long long total, int ste, float* __restrict__ z) {
    const float beta = params[0], gamma = params[1], zeta = params[2];
    const float loc[4] = {params[3], params[4], params[5], params[6]};
    const long long n_cols = ro.o[4];
    const long long n_vec = total >> 2;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_vec; i += stride) {
        const float4 l4 = __ldg(reinterpret_cast<const float4*>(logits) + i);
        float4 u4 = make_float4(0.5f, 0.5f, 0.5f, 0.5f);
        if (TRAINING) u4 = __ldg(reinterpret_cast<const float4*>(u) + i);
        const float l[4] = {l4.x, l4.y, l4.z, l4.w}, uu[4] = {u4.x, u4.y, u4.z, u4.w};
        float out[4];
        long long col = (i << 2) % n_cols;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float s, x;
            bool inside;
            float zz = hc_element(l[k], uu[k], loc[rank_of(col, ro)], beta, gamma, zeta, TRAINING, &s, &inside, &x);
            if (ste) zz = (zz > 0.5f) ? 1.0f : 0.0f;
            out[k] = zz;
            if (++col == n_cols) col = 0;
        }
        reinterpret_cast<float4*>(z)[i] = make_float4(out[0], out[1], out[2], out[3]);
    }
    // scalar tail (total not a multiple of 4)
    const long long tail = (n_vec << 2) + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (tail < total) {
        float s, x;
        bool inside;
        float zz = hc_element(logits[tail], TRAINING ? u[tail] : 0.5f, loc[rank_of(tail % n_cols, ro)], beta, gamma,
                              zeta, TRAINING, &s, &inside, &x);
        if (ste) zz = (zz > 0.5f) ? 1.0f : 0.0f;
        z[tail] = zz;
    }
}

// Backward: dL/dlogit = g * inside * (zeta-gamma) * s(1-s) / beta      (training; /1 in eval)
//           dL/dbeta  = sum g * inside * (zeta-gamma) * s(1-s) * (-x/beta)
//           dL/dgamma = sum g * inside * (1 - s);  dL/dzeta = sum g * inside * s
//           dL/dloc_r = sum over rank r of dL/dlogit
template <bool TRAINING>
__global__ void __launch_bounds__(256) hard_concrete_bwd_kernel(
    const float* __restrict__ logits, const float* __restrict__ u, const float* __restrict__ params,
    RankOffsets ro, long long total, const float* __restrict__ grad_z, float* __restrict__ grad_logits,
    float* __restrict__ grad_params, float* __restrict__ cta_partials) {
    const float beta = params[0], gamma = params[1], zeta = params[2];
    const float loc[4] = {params[3], params[4], params[5], params[6]};
    const long long n_cols = ro.o[4];
    const long long n_vec = total >> 2;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

    auto element = [&](float logit, float uu, float g, long long col) -> float {
        const int r = rank_of(col, ro);
        float s, x;
        bool inside;
        hc_element(logit, uu, loc[r], beta, gamma, zeta, TRAINING, &s, &inside, &x);
        const float gs = inside ? g : 0.0f;            // gradient w.r.t. sbar
        const float gx = gs * (zeta - gamma) * (s * (1.0f - s));
        const float gl = TRAINING ? gx / beta : gx;
        if (TRAINING) acc[0] += gx * (-x / beta);
        acc[1] += gs * (1.0f - s);
        acc[2] += gs * s;
        acc[3 + r] += gl;
        return gl;
    };

    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_vec; i += stride) {
        const float4 l4 = __ldg(reinterpret_cast<const float4*>(logits) + i);
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(grad_z) + i);
        float4 u4 = make_float4(0.5f, 0.5f, 0.5f, 0.5f);
        if (TRAINING) u4 = __ldg(reinterpret_cast<const float4*>(u) + i);
        const float l[4] = {l4.x, l4.y, l4.z, l4.w}, uu[4] = {u4.x, u4.y, u4.z, u4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
        float out[4];
        long long col = (i << 2) % n_cols;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            out[k] = element(l[k], uu[k], g[k], col);
            if (++col == n_cols) col = 0;
        }
        reinterpret_cast<float4*>(grad_logits)[i] = make_float4(out[0], out[1], out[2], out[3]);
    }
    const long long tail = (n_vec << 2) + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (tail < total)
        grad_logits[tail] = element(logits[tail], TRAINING ? u[tail] : 0.5f, grad_z[tail], tail % n_cols);

    // block reduction of the 7 parameter gradients, one atomic per block per parameter
    __shared__ float red[8][7];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const float v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        float v = 0.f;
        for (int w = 0; w < (blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
        if (cta_partials != nullptr) cta_partials[blockIdx.x * 8 + threadIdx.x] = v;      // added in CTA order by the kernel below
        else atomicAdd(&grad_params[threadIdx.x], v);
    }
}

// grad_params[k] = sum over the CTAs of the launch above, in a fixed order: warp k adds every 32nd slot per lane, then a shuffle tree
__global__ void __launch_bounds__(224) hard_concrete_bwd_reduce_kernel(const float* __restrict__ cta_partials, int n_ctas,
                                                                       float* __restrict__ grad_params) {
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float v = 0.f;
    for (int b = lane; b < n_ctas; b += 32) v += cta_partials[b * 8 + k];
    v = warp_sum(v);
    if (lane == 0) grad_params[k] = v;
}

// BinaryGumbel training branch: softmax over the pair ([l, 1-l] + g) / T, component 0.  Written in
// softmax form (max-subtracted exponentials) so exact zeros appear where torch's softmax has them.
__device__ __forceinline__ float gumbel_prob(float l, float g0, float g1, float temp, float* p1) {
    const float a0 = (l + g0) / temp, a1 = ((1.0f - l) + g1) / temp;
    const float m = fmaxf(a0, a1);
    const float e0 = expf(a0 - m), e1 = expf(a1 - m);
    const float inv = 1.0f / (e0 + e1);
    *p1 = e1 * inv;
    return e0 * inv;
}

__global__ void __launch_bounds__(256) binary_gumbel_fwd_kernel(const float* __restrict__ logits,
                                                                 const float* __restrict__ gumbels, float temp,
                                                                 long long count, float* __restrict__ probs) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count; i += stride) {
        float p1;
        probs[i] = gumbel_prob(__ldg(logits + i), __ldg(gumbels + i), __ldg(gumbels + count + i), temp, &p1);
    }
}

// Backward written in the order torch's autograd evaluates it (softmax backward over the pair, then the
// division by temp, then stack([l, 1 - l])), so that the fp32 cancellation in (g - g p0) matches:
//   d a0 = p0 (g - g p0),  d a1 = p1 (0 - g p0),  d l = d a0 / T - d a1 / T
__global__ void __launch_bounds__(256) binary_gumbel_bwd_kernel(const float* __restrict__ logits,
                                                                 const float* __restrict__ gumbels, float temp,
                                                                 long long count, const float* __restrict__ grad_probs,
                                                                 float* __restrict__ grad_logits) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count; i += stride) {
        float p1;
        const float p0 = gumbel_prob(__ldg(logits + i), __ldg(gumbels + i), __ldg(gumbels + count + i), temp, &p1);
        const float g = __ldg(grad_probs + i);
        const float dot = __fmul_rn(g, p0);
        const float da0 = __fmul_rn(p0, __fsub_rn(g, dot));
        const float da1 = __fmul_rn(p1, __fsub_rn(0.0f, dot));
        grad_logits[i] = __fsub_rn(da0 / temp, da1 / temp);
    }
}

int stream_grid(long long work_items) {
    const long long blocks = (work_items + 255) / 256;
    const long long cap = static_cast<long long>(sm_count()) * 8;   // 8 resident 256-thread CTAs per SM
    return static_cast<int>(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace topo

using namespace topo;

extern "C" int topo_hard_concrete_fwd(const float* logits, const float* u, const float* params,
                                      const int64_t offsets[5], int64_t batch, int training, int ste, float* z,
                                      topo_stream_t stream) {
    TOPO_REQUIRE(logits && params && offsets && z, "null argument");
    TOPO_REQUIRE(!training || u, "u is required in training mode");
    TOPO_REQUIRE(batch >= 0 && offsets[4] > 0, "bad sizes");
    TOPO_REQUIRE(aligned16(logits) && aligned16(z) && (!training || aligned16(u)), "buffers must be 16-byte aligned");
    const long long total = batch * offsets[4];
    if (total == 0) return TOPO_OK;
    RankOffsets ro;
    for (int i = 0; i < 5; ++i) ro.o[i] = offsets[i];
    const int grid = stream_grid((total >> 2) + 4);
    if (training)
        hard_concrete_fwd_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(logits, u, params, ro, total, ste, z);
    else
        hard_concrete_fwd_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(logits, u, params, ro, total, ste, z);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_hard_concrete_bwd(const float* logits, const float* u, const float* params,
                                      const int64_t offsets[5], int64_t batch, int training, const float* grad_z,
                                      float* grad_logits, float* grad_params, float* workspace, topo_stream_t stream) {
    TOPO_REQUIRE(logits && params && offsets && grad_z && grad_logits && grad_params, "null argument");
    TOPO_REQUIRE(!training || u, "u is required in training mode");
    TOPO_REQUIRE(batch >= 0 && offsets[4] > 0, "bad sizes");
    TOPO_REQUIRE(aligned16(logits) && aligned16(grad_z) && aligned16(grad_logits) && (!training || aligned16(u)),
                 "buffers must be 16-byte aligned");
    TOPO_CUDA(cudaMemsetAsync(grad_params, 0, 7 * sizeof(float), as_stream(stream)));
    const long long total = batch * offsets[4];
    if (total == 0) return TOPO_OK;
    RankOffsets ro;
    for (int i = 0; i < 5; ++i) ro.o[i] = offsets[i];
    const int grid = stream_grid((total >> 2) + 4);
    if (training)
        hard_concrete_bwd_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(logits, u, params, ro, total, grad_z,
                                                                            grad_logits, grad_params, workspace);
    else
        hard_concrete_bwd_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(logits, u, params, ro, total, grad_z,
                                                                             grad_logits, grad_params, workspace);
    if (workspace != nullptr) hard_concrete_bwd_reduce_kernel<<<1, 224, 0, as_stream(stream)>>>(workspace, grid, grad_params);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int64_t topo_hard_concrete_bwd_workspace_floats(const int64_t offsets[5], int64_t batch) {
    if (!offsets || batch < 0 || offsets[4] <= 0) return -1;
    return 8ll * stream_grid(((batch * offsets[4]) >> 2) + 4);
}

extern "C" int topo_binary_gumbel_fwd(const float* logits, const float* gumbels, float temp, int64_t count,
                                      float* probs, topo_stream_t stream) {
    TOPO_REQUIRE(logits && gumbels && probs, "null argument");
    TOPO_REQUIRE(count >= 0 && temp > 0.f, "bad arguments");
    if (count == 0) return TOPO_OK;
    binary_gumbel_fwd_kernel<<<stream_grid(count), 256, 0, as_stream(stream)>>>(logits, gumbels, temp, count, probs);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_binary_gumbel_bwd(const float* logits, const float* gumbels, float temp, int64_t count,
                                      const float* grad_probs, float* grad_logits, topo_stream_t stream) {
    TOPO_REQUIRE(logits && gumbels && grad_probs && grad_logits, "null argument");
    TOPO_REQUIRE(count >= 0 && temp > 0.f, "bad arguments");
    if (count == 0) return TOPO_OK;
    binary_gumbel_bwd_kernel<<<stream_grid(count), 256, 0, as_stream(stream)>>>(logits, gumbels, temp, count,
                                                                                grad_probs, grad_logits);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
