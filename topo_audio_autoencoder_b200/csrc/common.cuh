// Shared helpers for the topo_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "topo_b200.h"

namespace topo {

void set_error(const std::string& msg);

#define TOPO_REQUIRE(cond, msg)                                        \
    do {                                                               \
        if (!(cond)) {                                                 \
            ::topo::set_error(std::string(__func__) + ": " + (msg));   \
            return TOPO_ERR_INVALID;                                   \
        }                                                              \
    } while (0)

#define TOPO_CUDA(expr)                                                                 \
    do {                                                                                \
        cudaError_t err__ = (expr);                                                     \
        if (err__ != cudaSuccess) {                                                     \
            ::topo::set_error(std::string(__func__) + ": " + cudaGetErrorString(err__)); \
            return TOPO_ERR_CUDA;                                                       \
        }                                                                               \
    } while (0)

#define TOPO_LAUNCH_CHECK() TOPO_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(topo_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel (never again, e.g. not inside a graph capture)
int ensure_dynamic_smem(const void* kernel, size_t bytes);
// CTAs of a tensor-core combine launch (128-row tiles, persistent CTAs, at most one per SM and at most max_ctas when set)
inline int combine_grid(int64_t rows, int max_ctas) {
    const int64_t tiles = (rows + 127) / 128;
    const int cap = max_ctas > 0 ? (max_ctas < sm_count() ? max_ctas : sm_count()) : sm_count();
    return static_cast<int>(tiles < cap ? tiles : cap);
}
constexpr int kCtaPartialFloats = 16704;      // TOPO_CTA_PARTIAL_FLOATS: [w1][wprod x 3] 4 x 4096, [b1][w2][gamma][beta] 4 x 64, [b2] 1 (+ 63)

constexpr int kMaxRank = 3;

// Device-resident static tables of one vertex count.  All ids are int32 and local to their rank.
struct DeviceTables {
    int n;                 // vertices
    int cnt[4];            // n_r
    int off[5];            // rank offsets on the simplex axis
    int ncof[3];           // cofaces per simplex of rank r: n-1-r
    const int* faces[4];   // faces[r]: [n_r][r+1] ids in rank r-1 (r>=1); faces[0] = nullptr
    const int* cofaces[3]; // cofaces[r]: [n_r][ncof[r]] ids in rank r+1, ascending
    // explicit adjacency neighbour lists, sorted by neighbour id (operator builder only)
    int adj_w[4];          // slots per row: n-1, 2(n-2), 3(n-3), 4(n-4)
    const int* adj_nbr[4]; // [n_r][adj_w[r]] neighbour id in rank r
    const int* adj_via[4]; // [n_r][adj_w[r]] id of the simplex the pair shares (edge | coface | coface | face)
};

}  // namespace topo

struct topo_tables {
    topo::DeviceTables d;                 // device pointers
    int device;
    std::vector<int64_t> h_verts[4];      // [n_r][r+1] vertex ids (host)
    std::vector<int> h_faces[4];          // [n_r][r+1] (host)
    std::vector<int> h_cofaces[3];        // [n_r][n-1-r] (host)
    std::vector<void*> dev_blocks;        // everything cudaMalloc'ed for this object
};

namespace topo {
// The tables live on the device that was current when they were created: refuse a launch from any other device
// instead of dereferencing foreign pointers (one ConstraintMatrices per device; nn.Module.to() does not move them).
inline bool tables_on_current_device(const topo_tables* t) {
    int dev = -1;
    cudaGetDevice(&dev);
    return t->device >= 0 && t->device == dev;
}
}  // namespace topo
#define TOPO_REQUIRE_TABLES_DEVICE(t) \
    TOPO_REQUIRE(::topo::tables_on_current_device(t), "the tables were uploaded to another device (or built host-only): create them on the current device")

namespace topo {

// ---- device helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the 16 lanes that share (lane >> 4)
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float gelu_exact(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// d/dx [0.5 x (1 + erf(x/sqrt2))] = 0.5 (1 + erf(x/sqrt2)) + x * exp(-x^2/2) / sqrt(2 pi)
__device__ __forceinline__ float gelu_exact_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// The same two functions without erff (about 20 instructions instead of 60, no divergent branch): for
// z = |x| / sqrt 2,  erfc(z) = t P(t) exp(-z^2)  with  t = 1 / (1 + 0.42 z)  and P a degree-9 minimax fit of
// erfcx(z) / t on z in [0, 10] (relative error 2.3e-9; fitted with scipy).  exp(-z^2) = exp(-x^2/2)
// is also the Gaussian density the derivative needs.  Phi(x) = 1 - erfc(z)/2 for x >= 0 and erfc(z)/2 below, so
// the negative tail keeps full relative accuracy (0.5 (1 + erf) cancels there).  In fp32: |gelu error| < 4e-7.
__device__ __forceinline__ void gelu_cdf_pdf(float x, float& cdf, float& pdf) {
    const float ax = fabsf(x);
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(ax, 0.42f * 0.70710678118654752440f, 1.0f)));
    float p = -0.04113744501622546f;
    p = fmaf(p, t, 0.23070530236756076f);
    p = fmaf(p, t, -0.4866430497963537f);
    p = fmaf(p, t, 0.437150493790159f);
    p = fmaf(p, t, -0.19732979111351331f);
    p = fmaf(p, t, 0.2137338489469954f);
    p = fmaf(p, t, 0.150040600633153f);
    p = fmaf(p, t, 0.21989449846757936f);
    p = fmaf(p, t, 0.23661221813005356f);
    p = fmaf(p, t, 0.23697332130814744f);
    float e;                                           // exp(-x^2/2) = 2^(x * (x * -log2(e)/2)); flush-to-zero form: one MUFU
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * (x * -0.72134752044448170368f)));
    const float hq = (0.5f * t) * p * e;               // erfc(|x| / sqrt 2) / 2 = Phi(-|x|)
    cdf = x >= 0.0f ? 1.0f - hq : hq;
    pdf = 0.39894228040143267794f * e;
}
__device__ __forceinline__ float gelu_fast_exact(float x) {
    float cdf, pdf;
    gelu_cdf_pdf(x, cdf, pdf);
    return x * cdf;
}
// N independent evaluations written stage by stage: ptxas keeps a single evaluation as one dependent chain
// (about twenty instructions at four cycles each), so the interleaving has to be in the source.
template <int N>
__device__ __forceinline__ void gelu_cdf_pdf_n(const float* x, float* cdf, float* pdf) {
    float t[N], p[N], e[N];
#pragma unroll
    for (int i = 0; i < N; ++i)
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t[i]) : "f"(fmaf(fabsf(x[i]), 0.42f * 0.70710678118654752440f, 1.0f)));
#pragma unroll
    for (int i = 0; i < N; ++i)
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[i]) : "f"(x[i] * (x[i] * -0.72134752044448170368f)));
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fmaf(-0.04113744501622546f, t[i], 0.23070530236756076f);
    constexpr float kCoef[8] = {-0.4866430497963537f, 0.437150493790159f, -0.19732979111351331f, 0.2137338489469954f,
                                0.150040600633153f, 0.21989449846757936f, 0.23661221813005356f, 0.23697332130814744f};
#pragma unroll
    for (int s = 0; s < 8; ++s) {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = fmaf(p[i], t[i], kCoef[s]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float hq = (0.5f * t[i]) * p[i] * e[i];
        cdf[i] = x[i] >= 0.0f ? 1.0f - hq : hq;
        pdf[i] = 0.39894228040143267794f * e[i];
    }
}

template <int VEC>
struct Vec;
template <>
struct Vec<1> {
    float v[1];
    __device__ __forceinline__ void load(const float* p) { v[0] = __ldg(p); }
    __device__ __forceinline__ void store(float* p) const { p[0] = v[0]; }
};
template <>
struct Vec<2> {
    float v[2];
    __device__ __forceinline__ void load(const float* p) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    }
};
template <>
struct Vec<4> {
    float v[4];
    __device__ __forceinline__ void load(const float* p) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};

}  // namespace topo
