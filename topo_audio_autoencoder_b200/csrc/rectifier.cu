// R2 / G3 / L1 / L2: geometric-mean face rectifier, active-set compaction, structural penalties.
//
// Rectifier (replaces enforce_constraints, rectifier.py:75-127): one thread per (sample, simplex),
// one launch per level, faces gathered through the int32 face table (an int4 load for a
// tetrahedron) from the already-rectified level below.  HBM-bound and tiny (16 B / simplex fwd+bwd);
// the face table and the level below are L2-resident.
#include "common.cuh"

namespace topo {
namespace {

struct Split {
    float own;   // gradient reaching the simplex's own probability
    float gsum;  // gradient reaching the log-sum S of its faces
};

// torch.minimum(a, b): NaN-propagating.
__device__ __forceinline__ float torch_minimum(float a, float b) {
    return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b);
}

template <int ARITY>
__device__ __forceinline__ float constraint_value(const float (&pf)[ARITY], float eps, bool* any_zero, float* gm_out) {
    float s = 0.f;
    bool z = false;
#pragma unroll
    for (int a = 0; a < ARITY; ++a) {
        z |= (pf[a] == 0.0f);
        s += logf(pf[a] + eps);
    }
    const float gm = expf(s / static_cast<float>(ARITY));
    *any_zero = z;
    *gm_out = gm;
    return z ? (gm - gm) : gm;   // rectifier.py:94-96: exact zero with a zero-gradient path
}

// Gradient split of y = minimum(own, c), c = where(any_zero, gm - gm, gm), gm = exp(S / ARITY).
// Mirrors torch autograd: ties share the gradient 50/50; the masked branch passes (g - g) to gm.
template <int ARITY>
__device__ __forceinline__ Split split_gradient(float g, float own, float c, bool any_zero, float gm) {
    const float half_or_full = (own == c) ? g * 0.5f : g;
    Split r;
    r.own = (own > c) ? 0.0f : half_or_full;
    const float g_c = (own < c) ? 0.0f : half_or_full;
    const float g_gm = any_zero ? (g_c - g_c) : g_c;
    r.gsum = (g_gm * gm) / static_cast<float>(ARITY);
    return r;
}

template <int ARITY>
__device__ __forceinline__ void load_faces(const int* __restrict__ faces, int s, const float* __restrict__ below,
                                           float (&pf)[ARITY]) {
    int f[ARITY];
    if constexpr (ARITY == 4) {
        const int4 t = __ldg(reinterpret_cast<const int4*>(faces) + s);
        f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
    } else if constexpr (ARITY == 2) {
        const int2 t = __ldg(reinterpret_cast<const int2*>(faces) + s);
        f[0] = t.x; f[1] = t.y;
    } else {
#pragma unroll
        for (int a = 0; a < ARITY; ++a) f[a] = __ldg(faces + s * ARITY + a);
    }
#pragma unroll
    for (int a = 0; a < ARITY; ++a) pf[a] = below[f[a]];
}

// Forward level r (ARITY = r + 1).  probs_below points at the rectified rank r-1 slice base of the
// OUTPUT array (for edges: the vertex slice, which the same kernel copies from the input first).
template <int ARITY>
__global__ void __launch_bounds__(256) rectify_level_fwd(DeviceTables d, const float* __restrict__ probs_in,
                                                         float* __restrict__ probs_out, float eps, int batch) {
    constexpr int R = ARITY - 1;
    const int n_r = d.cnt[R];
    const long long total = static_cast<long long>(batch) * n_r;
    const long long N = d.off[4];
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / n_r), s = static_cast<int>(i % n_r);
        // vertices are never rectified: the edge level reads them from the input
        const float* below = (R == 1 ? probs_in : probs_out) + b * N + d.off[R - 1];
        float pf[ARITY];
        load_faces<ARITY>(d.faces[R], s, below, pf);
        bool z;
        float gm;
        const float c = constraint_value<ARITY>(pf, eps, &z, &gm);
        probs_out[b * N + d.off[R] + s] = torch_minimum(probs_in[b * N + d.off[R] + s], c);
    }
}

__global__ void __launch_bounds__(256) copy_vertices_kernel(DeviceTables d, const float* __restrict__ in,
                                                            float* __restrict__ out, int batch) {
    const long long total = static_cast<long long>(batch) * d.cnt[0];
    const long long N = d.off[4];
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / d.cnt[0], s = i % d.cnt[0];
        out[b * N + s] = in[b * N + s];
    }
}

// Backward level r.  total gradient on the rectified simplex = upstream + (sum over cofaces of the
// gradient of their log-sums) / (p' + eps); then the minimum / where / exp chain of this level.
// ws holds, per simplex, the gradient w.r.t. its log-sum S (written here, read by the level below).
template <int ARITY, bool TOP>
__global__ void __launch_bounds__(256) rectify_level_bwd(DeviceTables d, const float* __restrict__ probs_in,
                                                         const float* __restrict__ probs_out,
                                                         const float* __restrict__ grad_out, float eps, int batch,
                                                         float* __restrict__ grad_in, float* __restrict__ ws) {
    constexpr int R = ARITY - 1;
    const int n_r = d.cnt[R];
    const long long total = static_cast<long long>(batch) * n_r;
    const long long N = d.off[4];
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / n_r), s = static_cast<int>(i % n_r);
        const long long row = b * N;
        float g = grad_out[row + d.off[R] + s];
        if (!TOP) {
            const int w = d.ncof[R];
            const int* cof = d.cofaces[R] + static_cast<long long>(s) * w;
            const float* ws_up = ws + row + d.off[R + 1];
            float acc = 0.f;
            for (int j = 0; j < w; ++j) acc += ws_up[__ldg(cof + j)];
            g += acc / (probs_out[row + d.off[R] + s] + eps);
        }
        const float* below = (R == 1 ? probs_in : probs_out) + row + d.off[R - 1];
        float pf[ARITY];
        load_faces<ARITY>(d.faces[R], s, below, pf);
        bool z;
        float gm;
        const float c = constraint_value<ARITY>(pf, eps, &z, &gm);
        const Split sp = split_gradient<ARITY>(g, probs_in[row + d.off[R] + s], c, z, gm);
        grad_in[row + d.off[R] + s] = sp.own;
        ws[row + d.off[R] + s] = sp.gsum;
    }
}

// Vertices: the edge level indexes vertex pairs (rectifier.py:88), so autograd divides per term.
__global__ void __launch_bounds__(256) rectify_vertex_bwd(DeviceTables d, const float* __restrict__ probs_in,
                                                          const float* __restrict__ grad_out, float eps, int batch,
                                                          float* __restrict__ grad_in, const float* __restrict__ ws) {
    const int n0 = d.cnt[0];
    const long long total = static_cast<long long>(batch) * n0;
    const long long N = d.off[4];
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / n0), s = static_cast<int>(i % n0);
        const long long row = b * N;
        const float denom = probs_in[row + s] + eps;
        const int w = d.ncof[0];
        const int* cof = d.cofaces[0] + static_cast<long long>(s) * w;
        float g = grad_out[row + s];
        for (int j = 0; j < w; ++j) g += ws[row + d.off[1] + __ldg(cof + j)] / denom;
        grad_in[row + s] = g;
    }
}

// ---------------------------------------------------------------------------------------------
// Active sets: flags -> exclusive scan -> pos / act_idx / counts.  One CTA per (rank, sample).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) active_sets_kernel(DeviceTables d, const float* __restrict__ probs,
                                                          int* __restrict__ pos, int* __restrict__ act_idx,
                                                          int* __restrict__ counts) {
    const int r = blockIdx.x, b = blockIdx.y;
    const long long base = static_cast<long long>(b) * d.off[4] + d.off[r];
    const int n_r = d.cnt[r];
    __shared__ int warp_tot[8];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < n_r; start += 256) {
        const int s = start + threadIdx.x;
        const bool on = (s < n_r) && (probs[base + s] != 0.0f);   // nonzero(): NaN counts, -0.0 does not
        const unsigned m = __ballot_sync(0xffffffffu, on);
        const int in_warp = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int before = carry;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        const int p = before + in_warp;
        if (s < n_r) pos[base + s] = on ? p : -1;
        if (on) act_idx[base + p] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += warp_tot[w];
            carry += tot;
        }
        __syncthreads();
    }
    const int count = carry;
    for (int s = count + threadIdx.x; s < n_r; s += 256) act_idx[base + s] = -1;
    if (threadIdx.x == 0) counts[b * 4 + r] = count;
}

__global__ void row_offsets_kernel(const int* __restrict__ counts, int batch, int* __restrict__ row_off) {
    const int r = threadIdx.x;
    if (r >= 4) return;
    int acc = 0;
    for (int b = 0; b < batch; ++b) {
        row_off[r * (batch + 1) + b] = acc;
        acc += counts[b * 4 + r];
    }
    row_off[r * (batch + 1) + batch] = acc;
}

// ---------------------------------------------------------------------------------------------
// Penalties: one CTA per sample.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rank_sums(const DeviceTables& d, const float* __restrict__ p, float (&sum)[4],
                                          float* smem /* [4][4] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float a = 0.f;
        for (int s = threadIdx.x; s < d.cnt[r]; s += blockDim.x) a += p[d.off[r] + s];
        a = warp_sum(a);
        if (lane == 0) smem[r * 4 + warp] = a;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) sum[r] = smem[r * 4] + smem[r * 4 + 1] + smem[r * 4 + 2] + smem[r * 4 + 3];
}

constexpr float kEntEps = 1e-10f;

__global__ void __launch_bounds__(128) penalties_fwd_kernel(DeviceTables d, const float* __restrict__ probs,
                                                            float min_active, float max_active,
                                                            float* __restrict__ vertex_penalty,
                                                            float* __restrict__ entropy_loss) {
    __shared__ float smem[16];
    const float* p = probs + static_cast<long long>(blockIdx.x) * d.off[4];
    float sum[4];
    rank_sums(d, p, sum, smem);
    if (threadIdx.x == 0) {
        vertex_penalty[blockIdx.x] = fmaxf(min_active - sum[0], 0.f) + fmaxf(sum[0] - max_active, 0.f);
        float a[4], tot = 0.f;
        for (int r = 0; r < 4; ++r) { a[r] = sum[r] / static_cast<float>(d.cnt[r]); tot += a[r]; }
        float ent = 0.f;
        for (int r = 0; r < 4; ++r) { const float q = a[r] / (tot + kEntEps); ent += q * logf(q + kEntEps); }
        entropy_loss[blockIdx.x] = -0.1f * (-ent);
    }
}

__global__ void __launch_bounds__(128) penalties_bwd_kernel(DeviceTables d, const float* __restrict__ probs,
                                                            float min_active, float max_active,
                                                            const float* __restrict__ g_vp,
                                                            const float* __restrict__ g_ent,
                                                            float* __restrict__ grad_probs) {
    __shared__ float smem[16];
    __shared__ float coef[4];
    const long long row = static_cast<long long>(blockIdx.x) * d.off[4];
    float sum[4];
    rank_sums(d, probs + row, sum, smem);
    if (threadIdx.x == 0) {
        const float gv = g_vp ? g_vp[blockIdx.x] : 0.f, ge = g_ent ? g_ent[blockIdx.x] : 0.f;
        float a[4], tot = 0.f;
        for (int r = 0; r < 4; ++r) { a[r] = sum[r] / static_cast<float>(d.cnt[r]); tot += a[r]; }
        const float T = tot + kEntEps;
        float dq[4], dot = 0.f;
        for (int r = 0; r < 4; ++r) {
            const float q = a[r] / T;
            dq[r] = 0.1f * (logf(q + kEntEps) + q / (q + kEntEps));
            dot += dq[r] * q;
        }
        for (int r = 0; r < 4; ++r) coef[r] = ge * (dq[r] - dot) / T / static_cast<float>(d.cnt[r]);
        coef[0] += gv * (((sum[0] - max_active) > 0.f ? 1.f : 0.f) - ((min_active - sum[0]) > 0.f ? 1.f : 0.f));
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r)
        for (int s = threadIdx.x; s < d.cnt[r]; s += blockDim.x) grad_probs[row + d.off[r] + s] = coef[r];
}

int grid_for(long long items) {
    const long long blocks = (items + 255) / 256;
    const long long cap = static_cast<long long>(sm_count()) * 8;
    return static_cast<int>(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace
}  // namespace topo

using namespace topo;

extern "C" int topo_rectify_fwd(const topo_tables* t, const float* probs_in, float eps, int64_t batch,
                                float* probs_out, topo_stream_t stream) {
    TOPO_REQUIRE(t && probs_in && probs_out, "null argument");
    TOPO_REQUIRE_TABLES_DEVICE(t);
    TOPO_REQUIRE(batch >= 0 && batch < (1 << 24), "bad batch");
    TOPO_REQUIRE(probs_in != probs_out, "in-place rectification is not supported");
    if (batch == 0) return TOPO_OK;
    const DeviceTables& d = t->d;
    cudaStream_t s = as_stream(stream);
    const int B = static_cast<int>(batch);
    copy_vertices_kernel<<<grid_for(batch * d.cnt[0]), 256, 0, s>>>(d, probs_in, probs_out, B);
    if (d.cnt[1]) rectify_level_fwd<2><<<grid_for(batch * d.cnt[1]), 256, 0, s>>>(d, probs_in, probs_out, eps, B);
    if (d.cnt[2]) rectify_level_fwd<3><<<grid_for(batch * d.cnt[2]), 256, 0, s>>>(d, probs_in, probs_out, eps, B);
    if (d.cnt[3]) rectify_level_fwd<4><<<grid_for(batch * d.cnt[3]), 256, 0, s>>>(d, probs_in, probs_out, eps, B);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_rectify_bwd(const topo_tables* t, const float* probs_in, const float* probs_out,
                                const float* grad_out, float eps, int64_t batch, float* grad_in, float* workspace,
                                topo_stream_t stream) {
    TOPO_REQUIRE(t && probs_in && probs_out && grad_out && grad_in && workspace, "null argument");
    TOPO_REQUIRE_TABLES_DEVICE(t);
    TOPO_REQUIRE(batch >= 0 && batch < (1 << 24), "bad batch");
    if (batch == 0) return TOPO_OK;
    const DeviceTables& d = t->d;
    cudaStream_t s = as_stream(stream);
    const int B = static_cast<int>(batch);
    // top-down; the highest populated rank has no cofaces
    const int top = d.cnt[3] ? 3 : (d.cnt[2] ? 2 : (d.cnt[1] ? 1 : 0));
#define LEVEL(ARITY, R)                                                                                          \
    if (d.cnt[R]) {                                                                                              \
        if (top == R)                                                                                            \
            rectify_level_bwd<ARITY, true><<<grid_for(batch * d.cnt[R]), 256, 0, s>>>(d, probs_in, probs_out,    \
                                                                                      grad_out, eps, B, grad_in, \
                                                                                      workspace);                \
        else                                                                                                     \
            rectify_level_bwd<ARITY, false><<<grid_for(batch * d.cnt[R]), 256, 0, s>>>(d, probs_in, probs_out,   \
                                                                                       grad_out, eps, B, grad_in,\
                                                                                       workspace);               \
    }
    LEVEL(4, 3)
    LEVEL(3, 2)
    LEVEL(2, 1)
#undef LEVEL
    if (top == 0)
        TOPO_CUDA(cudaMemcpyAsync(grad_in, grad_out, sizeof(float) * batch * d.off[4], cudaMemcpyDeviceToDevice, s));
    else
        rectify_vertex_bwd<<<grid_for(batch * d.cnt[0]), 256, 0, s>>>(d, probs_in, grad_out, eps, B, grad_in, workspace);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_active_sets(const topo_tables* t, const float* probs, int64_t batch, int32_t* pos,
                                int32_t* act_idx, int32_t* counts, int32_t* row_off, topo_stream_t stream) {
    TOPO_REQUIRE(t && probs && pos && act_idx && counts && row_off, "null argument");
    TOPO_REQUIRE_TABLES_DEVICE(t);
    TOPO_REQUIRE(batch >= 0 && batch <= 65535, "batch must be in [0, 65535]");
    if (batch == 0) return TOPO_OK;
    cudaStream_t s = as_stream(stream);
    active_sets_kernel<<<dim3(4, static_cast<unsigned>(batch)), 256, 0, s>>>(t->d, probs, pos, act_idx, counts);
    row_offsets_kernel<<<1, 32, 0, s>>>(counts, static_cast<int>(batch), row_off);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_penalties_fwd(const topo_tables* t, const float* probs, int64_t batch, float min_active,
                                  float max_active, float* vertex_penalty, float* entropy_loss,
                                  topo_stream_t stream) {
    TOPO_REQUIRE(t && probs && vertex_penalty && entropy_loss, "null argument");
    TOPO_REQUIRE_TABLES_DEVICE(t);
    TOPO_REQUIRE(batch >= 0, "bad batch");
    if (batch == 0) return TOPO_OK;
    penalties_fwd_kernel<<<static_cast<unsigned>(batch), 128, 0, as_stream(stream)>>>(
        t->d, probs, min_active, max_active, vertex_penalty, entropy_loss);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_penalties_bwd(const topo_tables* t, const float* probs, int64_t batch, float min_active,
                                  float max_active, const float* g_vp, const float* g_ent, float* grad_probs,
                                  topo_stream_t stream) {
    TOPO_REQUIRE(t && probs && grad_probs, "null argument");
    TOPO_REQUIRE_TABLES_DEVICE(t);
    TOPO_REQUIRE(batch >= 0, "bad batch");
    if (batch == 0) return TOPO_OK;
    penalties_bwd_kernel<<<static_cast<unsigned>(batch), 128, 0, as_stream(stream)>>>(
        t->d, probs, min_active, max_active, g_vp, g_ent, grad_probs);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
