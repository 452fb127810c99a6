// (f)1, the consumer of the stage: the decoder's 4-head cross-attention (reference decoder.py:58-63, 158-162) over the
// COMPACT per-rank rows of a whole batch.
//
// The reference attends one sample at a time: queries = the L = 250 rows made from the sample's vertices, memory = the
// sample's active edges, triangles and tetrahedra (decoder.py:144-156).  A batched nn.MultiheadAttention needs that memory
// padded to the longest sample plus a key-padding mask, and its fp32 backward (the sm80 memory-efficient kernel, the only
// one that takes fp32 with a mask) spends 25 ms per 64 clips on a B200: a third of the whole training step.  Here the
// memory is never padded or copied: the projected keys / values stay in the concatenated compact layout the stage emits
// ([rows of all samples, rank after rank], C = heads x 16 columns) and `seg[b][r] = (first row, row count)` names the three
// runs of rows that belong to sample b.  Head dimension 16, everything fp32 on the FP32 pipes (13 GFLOP forward: a GEMM
// shape the tensor cores would need bf16x3 images and a shared-memory round trip of the probabilities for).
//
//   forward      CTA = 128 queries of one (sample, head); thread = 2 queries x 1 of 4 interleaved key subsets; keys and
//                values stream through shared memory in tiles of 128 (one broadcast LDS.128 feeds 8 FMAs of two queries);
//                online softmax in the log2 domain, the four partial states of a query meet in shuffles.  Saves
//                lse2 = log2 sum exp2(s2) per (sample, head, query).
//   backward Q   same mapping: recomputes p = exp2(s2 - lse2), dS = p (dO.v - D), dQ = scale sum dS k; also writes
//                D = dO.O per (sample, head, query) for the second pass.
//   backward KV  thread = one key of one head, its k, v, dK, dV in registers; the (sample, head)'s queries, dO, lse2 and D
//                stream through shared memory (every read a broadcast).  No atomics: both passes are deterministic.
#include <algorithm>

#include "common.cuh"

namespace topo {
namespace {

constexpr int kDh = 16;                 // head dimension
constexpr int kQPerCta = 128;           // queries per CTA (forward, backward Q)
constexpr int kKeyTile = 128;           // keys per shared-memory tile
constexpr int kPad = 20;                // floats per staged row: 80-byte rows keep four concurrent rows on distinct banks
constexpr int kKeysPerCta = 256;        // backward KV: one key per thread
constexpr int kQTile = 128;             // backward KV: queries per shared-memory tile
constexpr float kNegBig = -1.0e30f;
constexpr float kLn2 = 0.6931471805599453f;

struct Run {
    int start, len;
};
__device__ __forceinline__ Run run_of(const int* __restrict__ seg, long long b, int r) {
    return Run{__ldg(seg + (b * 3 + r) * 2), __ldg(seg + (b * 3 + r) * 2 + 1)};
}

__device__ __forceinline__ void load16(const float* __restrict__ p, float (&v)[kDh]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p) + j);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
}
__device__ __forceinline__ void lds16(const float* p, float (&v)[kDh]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 t = *(reinterpret_cast<const float4*>(p) + j);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
}
__device__ __forceinline__ float dot16(const float (&a)[kDh], const float (&b)[kDh]) {
    float s = a[0] * b[0];
#pragma unroll
    for (int c = 1; c < kDh; ++c) s = fmaf(a[c], b[c], s);
    return s;
}
__device__ __forceinline__ float sum4(float v) {          // over the four key subsets of a query (adjacent lanes)
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// stage rows [row0, row0 + n) of head h: k and v, 16 floats each, into padded shared-memory rows
__device__ __forceinline__ void stage_keys(const float* __restrict__ k, const float* __restrict__ v, long long row0, int n, int c,
                                           int h, float* ks, float* vs) {
    for (int idx = threadIdx.x; idx < kKeyTile * 4; idx += blockDim.x) {
        const int row = idx >> 2, c4 = idx & 3;
        if (row < n) {
            const long long off = (row0 + row) * c + h * kDh + c4 * 4;
            *reinterpret_cast<float4*>(ks + row * kPad + c4 * 4) = __ldg(reinterpret_cast<const float4*>(k + off));
            *reinterpret_cast<float4*>(vs + row * kPad + c4 * 4) = __ldg(reinterpret_cast<const float4*>(v + off));
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attention_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                            const float* __restrict__ v, const int* __restrict__ seg, int q_len,
                                                            int heads, float scale_log2e, float* __restrict__ out,
                                                            float* __restrict__ lse2) {
    __shared__ __align__(16) float ks[kKeyTile * kPad];
    __shared__ __align__(16) float vs[kKeyTile * kPad];
    const int c = heads * kDh;
    const long long b = blockIdx.z;
    const int h = blockIdx.y, sub = threadIdx.x & 3;
    const int qi[2] = {static_cast<int>(blockIdx.x) * kQPerCta + (threadIdx.x >> 2), static_cast<int>(blockIdx.x) * kQPerCta + 64 + (threadIdx.x >> 2)};
    float qr[2][kDh], acc[2][kDh], m[2], l[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        if (qi[t] < q_len) load16(q + (b * q_len + qi[t]) * c + h * kDh, qr[t]);
#pragma unroll
        for (int e = 0; e < kDh; ++e) {
            qr[t][e] = qi[t] < q_len ? qr[t][e] * scale_log2e : 0.f;
            acc[t][e] = 0.f;
        }
        m[t] = kNegBig;
        l[t] = 0.f;
    }
    for (int r = 0; r < 3; ++r) {
        const Run run = run_of(seg, b, r);
        for (int t0 = 0; t0 < run.len; t0 += kKeyTile) {
            const int n = min(kKeyTile, run.len - t0);
            __syncthreads();
            stage_keys(k, v, static_cast<long long>(run.start) + t0, n, c, h, ks, vs);
            __syncthreads();
            for (int j0 = sub; j0 < n; j0 += 16) {          // four keys of this thread's subset at a time
                float s[2][4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 4 * u;
                    float kr[kDh];
                    if (j < n) lds16(ks + j * kPad, kr);
#pragma unroll
                    for (int t = 0; t < 2; ++t) s[t][u] = j < n ? dot16(qr[t], kr) : kNegBig;
                }
                float p[2][4];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const float mx = fmaxf(fmaxf(s[t][0], s[t][1]), fmaxf(s[t][2], s[t][3]));
                    const float m_new = fmaxf(m[t], mx);
                    const float corr = exp2f(m[t] - m_new);
                    m[t] = m_new;
                    l[t] *= corr;
#pragma unroll
                    for (int e = 0; e < kDh; ++e) acc[t][e] *= corr;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        p[t][u] = (j0 + 4 * u < n) ? exp2f(s[t][u] - m_new) : 0.f;
                        l[t] += p[t][u];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 4 * u;
                    if (j < n) {
                        float vr[kDh];
                        lds16(vs + j * kPad, vr);
#pragma unroll
                        for (int t = 0; t < 2; ++t)
#pragma unroll
                            for (int e = 0; e < kDh; ++e) acc[t][e] = fmaf(p[t][u], vr[e], acc[t][e]);
                    }
                }
            }
        }
    }
    // the four key subsets of a query meet
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        float m_all = fmaxf(m[t], __shfl_xor_sync(0xffffffffu, m[t], 1));
        m_all = fmaxf(m_all, __shfl_xor_sync(0xffffffffu, m_all, 2));
        const float f = exp2f(m[t] - m_all);
        const float l_all = sum4(l[t] * f);
        const float inv = l_all > 0.f ? 1.0f / l_all : 0.f;
        float o[kDh];
#pragma unroll
        for (int e = 0; e < kDh; ++e) o[e] = sum4(acc[t][e] * f) * inv;
        if (qi[t] < q_len) {
            float* dst = out + (b * q_len + qi[t]) * c + h * kDh + 4 * sub;        // each subset lane stores a quarter of the row
            float4 w;
            if (sub == 0) w = make_float4(o[0], o[1], o[2], o[3]);
            else if (sub == 1) w = make_float4(o[4], o[5], o[6], o[7]);
            else if (sub == 2) w = make_float4(o[8], o[9], o[10], o[11]);
            else w = make_float4(o[12], o[13], o[14], o[15]);
            *reinterpret_cast<float4*>(dst) = w;
            if (sub == 0) lse2[(b * heads + h) * q_len + qi[t]] = m_all + log2f(fmaxf(l_all, 1e-38f));
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// backward, queries: dQ_i = scale sum_j dS_ij k_j,  dS_ij = p_ij (dO_i . v_j - D_i),  D_i = dO_i . O_i
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attention_bwd_q_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                              const float* __restrict__ v, const int* __restrict__ seg,
                                                              const float* __restrict__ out, const float* __restrict__ lse2,
                                                              const float* __restrict__ d_out, int q_len, int heads, float scale,
                                                              float* __restrict__ d_row /* [B, heads, L] */, float* __restrict__ dq) {
    __shared__ __align__(16) float ks[kKeyTile * kPad];
    __shared__ __align__(16) float vs[kKeyTile * kPad];
    const int c = heads * kDh;
    const long long b = blockIdx.z;
    const int h = blockIdx.y, sub = threadIdx.x & 3;
    const int qi[2] = {static_cast<int>(blockIdx.x) * kQPerCta + (threadIdx.x >> 2), static_cast<int>(blockIdx.x) * kQPerCta + 64 + (threadIdx.x >> 2)};
    float qr[2][kDh], go[2][kDh], acc[2][kDh], lse[2], dd[2];
    const float scale_log2e = scale * 1.4426950408889634f;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const bool ok = qi[t] < q_len;
        dd[t] = 0.f;
        lse[t] = 0.f;
        if (ok) {
            const long long row = (b * q_len + qi[t]) * c + h * kDh;
            float o[kDh];
            load16(q + row, qr[t]);
            load16(d_out + row, go[t]);
            load16(out + row, o);
            dd[t] = dot16(go[t], o);
            lse[t] = __ldg(lse2 + (b * heads + h) * q_len + qi[t]);
            if (sub == 0) d_row[(b * heads + h) * q_len + qi[t]] = dd[t];
        }
#pragma unroll
        for (int e = 0; e < kDh; ++e) {
            qr[t][e] = ok ? qr[t][e] * scale_log2e : 0.f;
            go[t][e] = ok ? go[t][e] : 0.f;
            acc[t][e] = 0.f;
        }
    }
    for (int r = 0; r < 3; ++r) {
        const Run run = run_of(seg, b, r);
        for (int t0 = 0; t0 < run.len; t0 += kKeyTile) {
            const int n = min(kKeyTile, run.len - t0);
            __syncthreads();
            stage_keys(k, v, static_cast<long long>(run.start) + t0, n, c, h, ks, vs);
            __syncthreads();
            for (int j = sub; j < n; j += 4) {
                float kr[kDh], vr[kDh];
                lds16(ks + j * kPad, kr);
                lds16(vs + j * kPad, vr);
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const float p = exp2f(dot16(qr[t], kr) - lse[t]);
                    const float ds = p * (dot16(go[t], vr) - dd[t]);
#pragma unroll
                    for (int e = 0; e < kDh; ++e) acc[t][e] = fmaf(ds, kr[e], acc[t][e]);
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        float o[kDh];
#pragma unroll
        for (int e = 0; e < kDh; ++e) o[e] = sum4(acc[t][e]) * scale;
        if (qi[t] < q_len) {
            float* dst = dq + (b * q_len + qi[t]) * c + h * kDh + 4 * sub;
            float4 w;
            if (sub == 0) w = make_float4(o[0], o[1], o[2], o[3]);
            else if (sub == 1) w = make_float4(o[4], o[5], o[6], o[7]);
            else if (sub == 2) w = make_float4(o[8], o[9], o[10], o[11]);
            else w = make_float4(o[12], o[13], o[14], o[15]);
            *reinterpret_cast<float4*>(dst) = w;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// backward, keys and values: dV_j = sum_i p_ij dO_i,  dK_j = scale sum_i dS_ij q_i
// grid.x = 3 runs x tiles_per_run tiles of 256 keys (tiles past a run's length exit at once)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kKeysPerCta) attention_bwd_kv_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                                       const float* __restrict__ v, const int* __restrict__ seg,
                                                                       const float* __restrict__ lse2, const float* __restrict__ d_out,
                                                                       const float* __restrict__ d_row, int q_len, int heads,
                                                                       float scale, int tiles_per_run, float* __restrict__ dk,
                                                                       float* __restrict__ dv) {
    __shared__ __align__(16) float qs[kQTile * kDh];
    __shared__ __align__(16) float gs[kQTile * kDh];
    __shared__ float ls[kQTile], ds_[kQTile];
    const int c = heads * kDh;
    const long long b = blockIdx.z;
    const int h = blockIdx.y;
    const Run run = run_of(seg, b, static_cast<int>(blockIdx.x) / tiles_per_run);
    const int key0 = (static_cast<int>(blockIdx.x) % tiles_per_run) * kKeysPerCta;
    if (key0 >= run.len) return;
    const int j = key0 + threadIdx.x;
    const bool ok = j < run.len;
    const long long krow = (static_cast<long long>(run.start) + j) * c + h * kDh;
    float kr[kDh], vr[kDh], gk[kDh], gv[kDh];
#pragma unroll
    for (int e = 0; e < kDh; ++e) kr[e] = vr[e] = gk[e] = gv[e] = 0.f;
    if (ok) {
        load16(k + krow, kr);
        load16(v + krow, vr);
    }
    const float scale_log2e = scale * 1.4426950408889634f;
    for (int i0 = 0; i0 < q_len; i0 += kQTile) {
        const int n = min(kQTile, q_len - i0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < kQTile * 4; idx += kKeysPerCta) {
            const int row = idx >> 2, c4 = idx & 3;
            if (row < n) {
                const long long off = (b * q_len + i0 + row) * c + h * kDh + c4 * 4;
                float4 t = __ldg(reinterpret_cast<const float4*>(q + off));
                t.x *= scale_log2e; t.y *= scale_log2e; t.z *= scale_log2e; t.w *= scale_log2e;
                *reinterpret_cast<float4*>(qs + row * kDh + c4 * 4) = t;
                *reinterpret_cast<float4*>(gs + row * kDh + c4 * 4) = __ldg(reinterpret_cast<const float4*>(d_out + off));
            }
        }
        for (int row = threadIdx.x; row < n; row += kKeysPerCta) {
            ls[row] = __ldg(lse2 + (b * heads + h) * q_len + i0 + row);
            ds_[row] = __ldg(d_row + (b * heads + h) * q_len + i0 + row);
        }
        __syncthreads();
#pragma unroll 2
        for (int i = 0; i < n; ++i) {
            float qv[kDh], gv_i[kDh];
            lds16(qs + i * kDh, qv);                       // every thread reads the same row: broadcasts
            lds16(gs + i * kDh, gv_i);
            const float p = exp2f(dot16(qv, kr) - ls[i]);
            const float ds = p * (dot16(gv_i, vr) - ds_[i]);
#pragma unroll
            for (int e = 0; e < kDh; ++e) {
                gv[e] = fmaf(p, gv_i[e], gv[e]);
                gk[e] = fmaf(ds, qv[e], gk[e]);
            }
        }
    }
    if (ok) {
        // qs holds q scale log2e: dK = scale sum dS q = ln2 sum dS (q scale log2e)
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            *reinterpret_cast<float4*>(dk + krow + 4 * j4) =
                make_float4(gk[4 * j4] * kLn2, gk[4 * j4 + 1] * kLn2, gk[4 * j4 + 2] * kLn2, gk[4 * j4 + 3] * kLn2);
            *reinterpret_cast<float4*>(dv + krow + 4 * j4) = make_float4(gv[4 * j4], gv[4 * j4 + 1], gv[4 * j4 + 2], gv[4 * j4 + 3]);
        }
    }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace topo

using namespace topo;

extern "C" int topo_cross_attention_fwd(const float* q, const float* k, const float* v, const int32_t* seg, int64_t batch,
                                        int64_t q_len, int heads, float* out, float* lse2, topo_stream_t stream) {
    TOPO_REQUIRE(q && k && v && seg && out && lse2, "null argument");
    TOPO_REQUIRE(batch >= 0 && batch < 65536 && q_len >= 1 && heads >= 1 && heads < 65536, "bad geometry");
    TOPO_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out), "buffers must be 16-byte aligned");
    if (batch == 0) return TOPO_OK;
    const dim3 grid(static_cast<unsigned>((q_len + kQPerCta - 1) / kQPerCta), static_cast<unsigned>(heads), static_cast<unsigned>(batch));
    attention_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(q, k, v, seg, static_cast<int>(q_len), heads,
                                                              1.4426950408889634f / sqrtf(static_cast<float>(kDh)), out, lse2);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_cross_attention_bwd(const float* q, const float* k, const float* v, const int32_t* seg, const float* out,
                                        const float* lse2, const float* d_out, int64_t batch, int64_t q_len, int heads,
                                        int64_t max_run_len, float* d_row, float* dq, float* dk, float* dv, topo_stream_t stream) {
    TOPO_REQUIRE(q && k && v && seg && out && lse2 && d_out && d_row && dq && dk && dv, "null argument");
    TOPO_REQUIRE(batch >= 0 && batch < 65536 && q_len >= 1 && heads >= 1 && heads < 65536 && max_run_len >= 0, "bad geometry");
    TOPO_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out) && aligned16(d_out) && aligned16(dq) && aligned16(dk) &&
                     aligned16(dv), "buffers must be 16-byte aligned");
    if (batch == 0) return TOPO_OK;
    const float scale = 1.0f / sqrtf(static_cast<float>(kDh));
    const dim3 grid_q(static_cast<unsigned>((q_len + kQPerCta - 1) / kQPerCta), static_cast<unsigned>(heads), static_cast<unsigned>(batch));
    attention_bwd_q_kernel<<<grid_q, 256, 0, as_stream(stream)>>>(q, k, v, seg, out, lse2, d_out, static_cast<int>(q_len), heads, scale,
                                                                  d_row, dq);
    if (max_run_len > 0) {
        const int tiles_per_run = static_cast<int>((max_run_len + kKeysPerCta - 1) / kKeysPerCta);
        const dim3 grid_kv(static_cast<unsigned>(3 * tiles_per_run), static_cast<unsigned>(heads), static_cast<unsigned>(batch));
        attention_bwd_kv_kernel<<<grid_kv, kKeysPerCta, 0, as_stream(stream)>>>(q, k, v, seg, lse2, d_out, d_row, static_cast<int>(q_len),
                                                                              heads, scale, tiles_per_run, dk, dv);
    }
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
