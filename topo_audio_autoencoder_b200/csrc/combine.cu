// S1 (b): the dense half of GradientSCCNLayer.forward (custom_sccn.py:73-136) for one rank, plus
// row LayerNorm and the active-embedding scaling (encoder.py:242-247).
//
// combine forward, one kernel, per 64-row tile of target simplices, everything on chip:
//     T_k = agg_k @ W_k            conv weight          (GEMM 64x64xC, fp32 FFMA, register tiled)
//     m_k = scale_k T_k + x        scale + residual     (custom_sccn.py:81-85, 98-102, 116-120)
//     s_k = w2 . GELU(W1 m_k + b1) + b2                 (message attention MLP, :129)
//     a   = softmax_k(s_k);  out = sum_k a_k m_k        (:130-132)
//     out = LayerNorm(out)                              (:133-134, training and not final layer)
// HBM traffic per row: (n_msgs + 2) * C * 4 bytes -- the aggregates and x in, out out.  Weights live
// in shared memory for the life of the (persistent) CTA.
//
// combine backward, two kernels: (1) recomputes the tile forward, back-propagates LayerNorm, the
// softmax and the attention MLP, emits dL/dm_k and all attention / LayerNorm parameter gradients;
// (2) per message, dL/dagg_k = scale_k dL/dm_k W_k^T and the weight-gradient product agg_k^T dL/dm_k.
//
// fp32 throughout (rtol 1e-5 against the PyTorch oracle rules out single-pass TF32).
#include <algorithm>

#include "common.cuh"

namespace topo {
namespace {

constexpr int TM = 64;        // rows per tile
constexpr int NT = 256;       // threads per CTA: 16 column groups x 16 row groups
constexpr int RPT = 4;        // rows per thread

template <int C>
struct Cfg {
    static constexpr int CPT = C / 16;      // columns per thread
    static constexpr int LD = C + 4;        // padded row stride of the activation tiles
    static constexpr int TILE = TM * LD;    // floats per activation tile
};

// acc[i][j] += sum_kk A[(ty*4+i)][kk] * B[kk][tx*CPT+j]      A: [TM][LD] tile, B: [C][C] row-major
template <int C>
__device__ __forceinline__ void gemm_nn(const float* __restrict__ As, const float* __restrict__ Bs,
                                        float (&acc)[RPT][Cfg<C>::CPT], int tx, int ty) {
    constexpr int CPT = Cfg<C>::CPT, LD = Cfg<C>::LD;
#pragma unroll 2
    for (int kk = 0; kk < C; kk += 4) {
        float4 a[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) a[i] = *reinterpret_cast<const float4*>(As + (ty * RPT + i) * LD + kk);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float b[CPT];
            const float* bp = Bs + (kk + q) * C + tx * CPT;
            if constexpr (CPT == 2) {
                const float2 t = *reinterpret_cast<const float2*>(bp);
                b[0] = t.x; b[1] = t.y;
            } else {
#pragma unroll
                for (int v = 0; v < CPT / 4; ++v) {
                    const float4 t = *reinterpret_cast<const float4*>(bp + v * 4);
                    b[v * 4] = t.x; b[v * 4 + 1] = t.y; b[v * 4 + 2] = t.z; b[v * 4 + 3] = t.w;
                }
            }
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const float av = q == 0 ? a[i].x : (q == 1 ? a[i].y : (q == 2 ? a[i].z : a[i].w));
#pragma unroll
                for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(av, b[j], acc[i][j]);
            }
        }
    }
}

// acc[i][j] += sum_r A[r][ty*4+i] * B[r][tx*CPT+j]           both [TM][LD] tiles (a K = rows product)
template <int C>
__device__ __forceinline__ void gemm_tn(const float* __restrict__ As, const float* __restrict__ Bs,
                                        float (&acc)[RPT][Cfg<C>::CPT], int tx, int ty) {
    constexpr int CPT = Cfg<C>::CPT, LD = Cfg<C>::LD;
    if (ty * RPT >= C) return;   // the C x C product needs only C / RPT row groups
#pragma unroll 4
    for (int r = 0; r < TM; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(As + r * LD + ty * RPT);
        float b[CPT];
        const float* bp = Bs + r * LD + tx * CPT;
        if constexpr (CPT == 2) {
            const float2 t = *reinterpret_cast<const float2*>(bp);
            b[0] = t.x; b[1] = t.y;
        } else {
#pragma unroll
            for (int v = 0; v < CPT / 4; ++v) {
                const float4 t = *reinterpret_cast<const float4*>(bp + v * 4);
                b[v * 4] = t.x; b[v * 4 + 1] = t.y; b[v * 4 + 2] = t.z; b[v * 4 + 3] = t.w;
            }
        }
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(av[i], b[j], acc[i][j]);
    }
}

template <int C>
__device__ __forceinline__ void zero_acc(float (&acc)[RPT][Cfg<C>::CPT]) {
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < Cfg<C>::CPT; ++j) acc[i][j] = 0.f;
}

// global [rows, C] tile -> smem [TM][LD]; rows past `live` become zeros
template <int C>
__device__ __forceinline__ void load_tile(const float* __restrict__ g, long long row0, long long live,
                                          float* __restrict__ s) {
    constexpr int V = C / 4;
    for (int idx = threadIdx.x; idx < TM * V; idx += NT) {
        const int r = idx / V, c4 = idx % V;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < live) v = __ldg(reinterpret_cast<const float4*>(g + (row0 + r) * C) + c4);
        *reinterpret_cast<float4*>(s + r * Cfg<C>::LD + c4 * 4) = v;
    }
}

// the thread's own [4][CPT] patch, global <-> registers
template <int C>
__device__ __forceinline__ void load_patch(const float* __restrict__ g, long long row0, long long live, int tx,
                                           int ty, float (&v)[RPT][Cfg<C>::CPT]) {
    constexpr int CPT = Cfg<C>::CPT;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const long long row = row0 + ty * RPT + i;
#pragma unroll
        for (int j = 0; j < CPT; ++j) v[i][j] = 0.f;
        if (g != nullptr && row < live) {
            const float* p = g + row * C + tx * CPT;
            if constexpr (CPT == 2) {
                const float2 t = __ldg(reinterpret_cast<const float2*>(p));
                v[i][0] = t.x; v[i][1] = t.y;
            } else {
#pragma unroll
                for (int q = 0; q < CPT / 4; ++q) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + q);
                    v[i][q * 4] = t.x; v[i][q * 4 + 1] = t.y; v[i][q * 4 + 2] = t.z; v[i][q * 4 + 3] = t.w;
                }
            }
        }
    }
}

template <int C>
__device__ __forceinline__ void store_patch(float* __restrict__ g, long long row0, long long live, int tx, int ty,
                                            const float (&v)[RPT][Cfg<C>::CPT]) {
    constexpr int CPT = Cfg<C>::CPT;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const long long row = row0 + ty * RPT + i;
        if (row < live) {
            float* p = g + row * C + tx * CPT;
            if constexpr (CPT == 2) {
                *reinterpret_cast<float2*>(p) = make_float2(v[i][0], v[i][1]);
            } else {
#pragma unroll
                for (int q = 0; q < CPT / 4; ++q)
                    *reinterpret_cast<float4*>(p + q * 4) = make_float4(v[i][q * 4], v[i][q * 4 + 1], v[i][q * 4 + 2], v[i][q * 4 + 3]);
            }
        }
    }
}

template <int C>
__device__ __forceinline__ void patch_to_tile(float* __restrict__ s, int tx, int ty,
                                              const float (&v)[RPT][Cfg<C>::CPT]) {
    constexpr int CPT = Cfg<C>::CPT;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        float* p = s + (ty * RPT + i) * Cfg<C>::LD + tx * CPT;
#pragma unroll
        for (int j = 0; j < CPT; ++j) p[j] = v[i][j];
    }
}

template <int C>
__device__ __forceinline__ void tile_to_patch(const float* __restrict__ s, int tx, int ty,
                                              float (&v)[RPT][Cfg<C>::CPT]) {
    constexpr int CPT = Cfg<C>::CPT;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const float* p = s + (ty * RPT + i) * Cfg<C>::LD + tx * CPT;
#pragma unroll
        for (int j = 0; j < CPT; ++j) v[i][j] = p[j];
    }
}

// [C][C] global matrix -> smem, optionally transposed
template <int C>
__device__ __forceinline__ void load_matrix(const float* __restrict__ g, float* __restrict__ s, bool transpose) {
    for (int idx = threadIdx.x; idx < C * C; idx += NT) {
        const float v = __ldg(g + idx);
        if (transpose) s[(idx % C) * C + idx / C] = v;
        else s[idx] = v;
    }
}

struct Softmax3 {
    float a[3];
};
__device__ __forceinline__ Softmax3 softmax_msgs(const float (&sc)[3], int n) {
    float mx = sc[0];
    for (int k = 1; k < n; ++k) mx = fmaxf(mx, sc[k]);
    Softmax3 r;
    float sum = 0.f;
    for (int k = 0; k < 3; ++k) {
        r.a[k] = (k < n) ? expf(sc[k] - mx) : 0.f;
        sum += r.a[k];
    }
    for (int k = 0; k < 3; ++k) r.a[k] /= sum;
    return r;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(NT) combine_fwd_kernel(topo_combine_params P, long long rows,
                                                         const int* __restrict__ n_rows_dev,
                                                         float* __restrict__ out) {
    constexpr int CPT = Cfg<C>::CPT;
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;                          // [3][C][C]
    float* W1t = Ws + 3 * C * C;               // [C][C]  W1 transposed: [in][out]
    float* vecs = W1t + C * C;                 // b1, w2, gamma, beta
    float* As = vecs + 4 * C;                  // [TM][LD]
    float* Ms = As + Cfg<C>::TILE;             // [TM][LD]

    const long long live = n_rows_dev ? min(static_cast<long long>(*n_rows_dev), rows) : rows;
    const long long tiles = (live + TM - 1) / TM;
    if (blockIdx.x >= tiles) return;

    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int k = 0; k < P.n_msgs; ++k) load_matrix<C>(P.w[k], Ws + k * C * C, false);
    load_matrix<C>(P.att_w1, W1t, true);
    for (int c = threadIdx.x; c < C; c += NT) {
        vecs[c] = __ldg(P.att_b1 + c);
        vecs[C + c] = __ldg(P.att_w2 + c);
        vecs[2 * C + c] = P.apply_ln ? __ldg(P.ln_gamma + c) : 1.f;
        vecs[3 * C + c] = P.apply_ln ? __ldg(P.ln_beta + c) : 0.f;
    }
    float scale[3];
    for (int k = 0; k < 3; ++k) scale[k] = k < P.n_msgs ? __ldg(P.scale[k]) : 0.f;
    const float b2 = __ldg(P.att_b2);
    __syncthreads();

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row0 = tile * TM;
        float xres[RPT][CPT];
        load_patch<C>(P.x, row0, live, tx, ty, xres);
        float m[3][RPT][CPT];
        float sc[RPT][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < P.n_msgs) {
                load_tile<C>(P.agg[k], row0, live, As);
                __syncthreads();
                float acc[RPT][CPT];
                zero_acc<C>(acc);
                gemm_nn<C>(As, Ws + k * C * C, acc, tx, ty);
#pragma unroll
                for (int i = 0; i < RPT; ++i)
#pragma unroll
                    for (int j = 0; j < CPT; ++j) m[k][i][j] = fmaf(scale[k], acc[i][j], xres[i][j]);
                patch_to_tile<C>(Ms, tx, ty, m[k]);
                if (P.saved_m[k] != nullptr) store_patch<C>(P.saved_m[k], row0, live, tx, ty, m[k]);
                __syncthreads();
                zero_acc<C>(acc);
                gemm_nn<C>(Ms, W1t, acc, tx, ty);
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    float part = 0.f;
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        const int col = tx * CPT + j;
                        acc[i][j] += vecs[col];                      // pre-GELU hidden activation
                        part = fmaf(gelu_exact(acc[i][j]), vecs[C + col], part);
                    }
                    sc[i][k] = half_warp_sum(part) + b2;
                }
                if (P.saved_pre[k] != nullptr) store_patch<C>(P.saved_pre[k], row0, live, tx, ty, acc);
            } else {
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    sc[i][k] = 0.f;
#pragma unroll
                    for (int j = 0; j < CPT; ++j) m[k][i][j] = 0.f;
                }
            }
        }
        float o[RPT][CPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const Softmax3 a = softmax_msgs(sc[i], P.n_msgs);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                o[i][j] = a.a[0] * m[0][i][j] + a.a[1] * m[1][i][j] + a.a[2] * m[2][i][j];
                sum += o[i][j];
            }
            if (P.apply_ln) {
                const float mean = half_warp_sum(sum) * (1.0f / C);
                float var = 0.f;
#pragma unroll
                for (int j = 0; j < CPT; ++j) var = fmaf(o[i][j] - mean, o[i][j] - mean, var);
                const float rstd = 1.0f / sqrtf(half_warp_sum(var) * (1.0f / C) + P.ln_eps);
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const int col = tx * CPT + j;
                    o[i][j] = fmaf((o[i][j] - mean) * rstd, vecs[2 * C + col], vecs[3 * C + col]);
                }
            }
        }
        store_patch<C>(out, row0, live, tx, ty, o);
        __syncthreads();   // As / Ms are rewritten by the next tile
    }
}

// ---------------------------------------------------------------------------------------------
// backward (1): LayerNorm, softmax and attention-MLP; emits dL/dm_k
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(NT) combine_bwd_attn_kernel(topo_combine_params P, long long rows,
                                                              const int* __restrict__ n_rows_dev,
                                                              const float* __restrict__ grad_out,
                                                              topo_combine_grads G, float* __restrict__ dm_ws) {
    constexpr int CPT = Cfg<C>::CPT;
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;                          // [3][C][C]
    float* W1t = Ws + 3 * C * C;               // [in][out]
    float* W1n = W1t + C * C;                  // [out][in] (as stored by nn.Linear)
    float* vecs = W1n + C * C;                 // b1, w2, gamma
    float* Mk = vecs + 4 * C;                  // [3][TM][LD]
    float* As = Mk + 3 * Cfg<C>::TILE;         // [TM][LD]
    float* red = As + Cfg<C>::TILE;            // [4][C] column partial sums at the end

    const long long live = n_rows_dev ? min(static_cast<long long>(*n_rows_dev), rows) : rows;
    const long long tiles = (live + TM - 1) / TM;
    if (blockIdx.x >= tiles) return;

    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int k = 0; k < P.n_msgs; ++k) load_matrix<C>(P.w[k], Ws + k * C * C, false);
    load_matrix<C>(P.att_w1, W1t, true);
    load_matrix<C>(P.att_w1, W1n, false);
    for (int c = threadIdx.x; c < C; c += NT) {
        vecs[c] = __ldg(P.att_b1 + c);
        vecs[C + c] = __ldg(P.att_w2 + c);
        vecs[2 * C + c] = P.apply_ln ? __ldg(P.ln_gamma + c) : 1.f;
    }
    float scale[3];
    for (int k = 0; k < 3; ++k) scale[k] = k < P.n_msgs ? __ldg(P.scale[k]) : 0.f;
    const float b2 = __ldg(P.att_b2);

    // per-thread parameter-gradient partials, reduced once per CTA
    float p_w1[RPT][CPT];                      // dW1[o = ty*4+i][in = tx*CPT+j]
    zero_acc<C>(p_w1);
    float p_b1[CPT], p_w2[CPT], p_gamma[CPT], p_beta[CPT], p_b2 = 0.f;
#pragma unroll
    for (int j = 0; j < CPT; ++j) p_b1[j] = p_w2[j] = p_gamma[j] = p_beta[j] = 0.f;
    __syncthreads();

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row0 = tile * TM;
        float sc[RPT][3];
        {
            float xres[RPT][CPT];
            load_patch<C>(P.x, row0, live, tx, ty, xres);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (k < P.n_msgs) {
                    load_tile<C>(P.agg[k], row0, live, As);
                    __syncthreads();
                    float acc[RPT][CPT];
                    zero_acc<C>(acc);
                    gemm_nn<C>(As, Ws + k * C * C, acc, tx, ty);
#pragma unroll
                    for (int i = 0; i < RPT; ++i)
#pragma unroll
                        for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(scale[k], acc[i][j], xres[i][j]);
                    patch_to_tile<C>(Mk + k * Cfg<C>::TILE, tx, ty, acc);
                    __syncthreads();
                    zero_acc<C>(acc);
                    gemm_nn<C>(Mk + k * Cfg<C>::TILE, W1t, acc, tx, ty);
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        float part = 0.f;
#pragma unroll
                        for (int j = 0; j < CPT; ++j) {
                            const int col = tx * CPT + j;
                            part = fmaf(gelu_exact(acc[i][j] + vecs[col]), vecs[C + col], part);
                        }
                        sc[i][k] = half_warp_sum(part) + b2;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < RPT; ++i) sc[i][k] = 0.f;
                }
            }
        }

        // mix, LayerNorm backward -> dmix; softmax backward -> dsc
        float dmix[RPT][CPT];
        load_patch<C>(grad_out, row0, live, tx, ty, dmix);
        float att[RPT][3], dsc[RPT][3];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const Softmax3 a = softmax_msgs(sc[i], P.n_msgs);
            float mk[3][CPT];
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int j = 0; j < CPT; ++j)
                    mk[k][j] = (k < P.n_msgs) ? Mk[k * Cfg<C>::TILE + (ty * RPT + i) * Cfg<C>::LD + tx * CPT + j] : 0.f;
            if (P.apply_ln) {
                float mix[CPT], sum = 0.f;
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    mix[j] = a.a[0] * mk[0][j] + a.a[1] * mk[1][j] + a.a[2] * mk[2][j];
                    sum += mix[j];
                }
                const float mean = half_warp_sum(sum) * (1.0f / C);
                float var = 0.f;
#pragma unroll
                for (int j = 0; j < CPT; ++j) var = fmaf(mix[j] - mean, mix[j] - mean, var);
                const float rstd = 1.0f / sqrtf(half_warp_sum(var) * (1.0f / C) + P.ln_eps);
                float c1 = 0.f, c2 = 0.f, xh[CPT], gy[CPT];
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    xh[j] = (mix[j] - mean) * rstd;
                    p_gamma[j] = fmaf(dmix[i][j], xh[j], p_gamma[j]);
                    p_beta[j] += dmix[i][j];
                    gy[j] = dmix[i][j] * vecs[2 * C + tx * CPT + j];
                    c1 += gy[j];
                    c2 = fmaf(gy[j], xh[j], c2);
                }
                c1 = half_warp_sum(c1) * (1.0f / C);
                c2 = half_warp_sum(c2) * (1.0f / C);
#pragma unroll
                for (int j = 0; j < CPT; ++j) dmix[i][j] = rstd * (gy[j] - c1 - xh[j] * c2);
            }
            float da[3], dot = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float part = 0.f;
#pragma unroll
                for (int j = 0; j < CPT; ++j) part = fmaf(dmix[i][j], mk[k][j], part);
                da[k] = half_warp_sum(part);
                dot = fmaf(a.a[k], da[k], dot);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                att[i][k] = a.a[k];
                dsc[i][k] = a.a[k] * (da[k] - dot);
                if (tx == 0 && k < P.n_msgs) p_b2 += dsc[i][k];
            }
        }

        // second pass over the messages: attention MLP backward, dL/dm_k
        float dxres[RPT][CPT];
        zero_acc<C>(dxres);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < P.n_msgs) {
                const float* Mt = Mk + k * Cfg<C>::TILE;
                float acc[RPT][CPT];
                zero_acc<C>(acc);
                gemm_nn<C>(Mt, W1t, acc, tx, ty);
#pragma unroll
                for (int i = 0; i < RPT; ++i)
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        const int col = tx * CPT + j;
                        const float pre = acc[i][j] + vecs[col];
                        p_w2[j] = fmaf(dsc[i][k], gelu_exact(pre), p_w2[j]);
                        const float dpre = dsc[i][k] * vecs[C + col] * gelu_exact_grad(pre);
                        p_b1[j] += dpre;
                        acc[i][j] = dpre;
                    }
                __syncthreads();                       // previous readers of As are done
                patch_to_tile<C>(As, tx, ty, acc);
                __syncthreads();
                zero_acc<C>(acc);
                gemm_nn<C>(As, W1n, acc, tx, ty);      // dpre @ W1  ([o][in])
#pragma unroll
                for (int i = 0; i < RPT; ++i)
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        acc[i][j] = fmaf(att[i][k], dmix[i][j], acc[i][j]);
                        dxres[i][j] += acc[i][j];
                    }
                store_patch<C>(dm_ws + static_cast<long long>(k) * rows * C, row0, live, tx, ty, acc);
                gemm_tn<C>(As, Mt, p_w1, tx, ty);      // dW1 += dpre^T m_k
            }
        }
        if (G.g_x != nullptr) store_patch<C>(G.g_x, row0, live, tx, ty, dxres);
        __syncthreads();
    }

    // ---- reduce the parameter gradients of this CTA ----
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < CPT; ++j)
            if (ty * RPT < C) atomicAdd(G.g_att_w1 + (ty * RPT + i) * C + tx * CPT + j, p_w1[i][j]);
    // column vectors: sum over the 16 row groups through shared memory
    __syncthreads();
    float* scratch = Mk;                                // [16][4][C] fits in the message tiles
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int col = tx * CPT + j;
        scratch[(ty * 4 + 0) * C + col] = p_b1[j];
        scratch[(ty * 4 + 1) * C + col] = p_w2[j];
        scratch[(ty * 4 + 2) * C + col] = p_gamma[j];
        scratch[(ty * 4 + 3) * C + col] = p_beta[j];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 4 * C; idx += NT) {
        const int which = idx / C, col = idx % C;
        float s = 0.f;
        for (int g = 0; g < 16; ++g) s += scratch[(g * 4 + which) * C + col];
        red[idx] = s;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 4 * C; idx += NT) {
        const int which = idx / C, col = idx % C;
        float* dst = which == 0 ? G.g_att_b1 : (which == 1 ? G.g_att_w2 : (which == 2 ? G.g_ln_gamma : G.g_ln_beta));
        if (dst != nullptr && (which < 2 || P.apply_ln)) atomicAdd(dst + col, red[idx]);
    }
    // b2: only tx == 0 threads hold a contribution
    p_b2 = warp_sum(p_b2);
    if ((threadIdx.x & 31) == 0 && p_b2 != 0.f) atomicAdd(G.g_att_b2, p_b2);
}

// ---------------------------------------------------------------------------------------------
// backward (1'), saved-activation variant: the forward left m_k and the pre-GELU hidden layer in HBM, so
// neither the conv-weight GEMM nor the first attention GEMM is recomputed -- two GEMM-equivalents per
// message (dpre W1 and the weight-gradient product dpre^T m_k) instead of five.
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(NT) combine_bwd_attn_saved_kernel(topo_combine_params P, long long rows,
                                                                    const int* __restrict__ n_rows_dev,
                                                                    const float* __restrict__ grad_out,
                                                                    topo_combine_grads G, float* __restrict__ dm_ws) {
    constexpr int CPT = Cfg<C>::CPT;
    extern __shared__ __align__(16) float smem[];
    float* W1n = smem;                         // [out][in] (as stored by nn.Linear)
    float* vecs = W1n + C * C;                 // w2, gamma
    float* Mk = vecs + 2 * C;                  // [3][TM][LD]
    float* As = Mk + 3 * Cfg<C>::TILE;         // [TM][LD]
    float* red = As + Cfg<C>::TILE;            // [4][C]

    const long long live = n_rows_dev ? min(static_cast<long long>(*n_rows_dev), rows) : rows;
    const long long tiles = (live + TM - 1) / TM;
    if (blockIdx.x >= tiles) return;

    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    load_matrix<C>(P.att_w1, W1n, false);
    for (int c = threadIdx.x; c < C; c += NT) {
        vecs[c] = __ldg(P.att_w2 + c);
        vecs[C + c] = P.apply_ln ? __ldg(P.ln_gamma + c) : 1.f;
    }
    const float b2 = __ldg(P.att_b2);
    float p_w1[RPT][CPT];
    zero_acc<C>(p_w1);
    float p_b1[CPT], p_w2[CPT], p_gamma[CPT], p_beta[CPT], p_b2 = 0.f;
#pragma unroll
    for (int j = 0; j < CPT; ++j) p_b1[j] = p_w2[j] = p_gamma[j] = p_beta[j] = 0.f;
    __syncthreads();

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row0 = tile * TM;
        float sc[RPT][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < P.n_msgs) {
                load_tile<C>(P.saved_m[k], row0, live, Mk + k * Cfg<C>::TILE);
                float pre[RPT][CPT];
                load_patch<C>(P.saved_pre[k], row0, live, tx, ty, pre);
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    float part = 0.f;
#pragma unroll
                    for (int j = 0; j < CPT; ++j) part = fmaf(gelu_exact(pre[i][j]), vecs[tx * CPT + j], part);
                    sc[i][k] = half_warp_sum(part) + b2;
                }
            } else {
#pragma unroll
                for (int i = 0; i < RPT; ++i) sc[i][k] = 0.f;
            }
        }
        __syncthreads();                               // the message tiles are in shared memory

        float dmix[RPT][CPT];
        load_patch<C>(grad_out, row0, live, tx, ty, dmix);
        float att[RPT][3], dsc[RPT][3];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const Softmax3 a = softmax_msgs(sc[i], P.n_msgs);
            float mk[3][CPT];
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int j = 0; j < CPT; ++j)
                    mk[k][j] = (k < P.n_msgs) ? Mk[k * Cfg<C>::TILE + (ty * RPT + i) * Cfg<C>::LD + tx * CPT + j] : 0.f;
            if (P.apply_ln) {
                float mix[CPT], sum = 0.f;
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    mix[j] = a.a[0] * mk[0][j] + a.a[1] * mk[1][j] + a.a[2] * mk[2][j];
                    sum += mix[j];
                }
                const float mean = half_warp_sum(sum) * (1.0f / C);
                float var = 0.f;
#pragma unroll
                for (int j = 0; j < CPT; ++j) var = fmaf(mix[j] - mean, mix[j] - mean, var);
                const float rstd = 1.0f / sqrtf(half_warp_sum(var) * (1.0f / C) + P.ln_eps);
                float c1 = 0.f, c2 = 0.f, xh[CPT], gy[CPT];
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    xh[j] = (mix[j] - mean) * rstd;
                    p_gamma[j] = fmaf(dmix[i][j], xh[j], p_gamma[j]);
                    p_beta[j] += dmix[i][j];
                    gy[j] = dmix[i][j] * vecs[C + tx * CPT + j];
                    c1 += gy[j];
                    c2 = fmaf(gy[j], xh[j], c2);
                }
                c1 = half_warp_sum(c1) * (1.0f / C);
                c2 = half_warp_sum(c2) * (1.0f / C);
#pragma unroll
                for (int j = 0; j < CPT; ++j) dmix[i][j] = rstd * (gy[j] - c1 - xh[j] * c2);
            }
            float da[3], dot = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float part = 0.f;
#pragma unroll
                for (int j = 0; j < CPT; ++j) part = fmaf(dmix[i][j], mk[k][j], part);
                da[k] = half_warp_sum(part);
                dot = fmaf(a.a[k], da[k], dot);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                att[i][k] = a.a[k];
                dsc[i][k] = a.a[k] * (da[k] - dot);
                if (tx == 0 && k < P.n_msgs) p_b2 += dsc[i][k];
            }
        }

        float dxres[RPT][CPT];
        zero_acc<C>(dxres);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < P.n_msgs) {
                const float* Mt = Mk + k * Cfg<C>::TILE;
                float acc[RPT][CPT];
                load_patch<C>(P.saved_pre[k], row0, live, tx, ty, acc);
#pragma unroll
                for (int i = 0; i < RPT; ++i)
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        const float pre = acc[i][j];
                        p_w2[j] = fmaf(dsc[i][k], gelu_exact(pre), p_w2[j]);
                        const float dpre = dsc[i][k] * vecs[tx * CPT + j] * gelu_exact_grad(pre);
                        p_b1[j] += dpre;
                        acc[i][j] = dpre;
                    }
                __syncthreads();                       // previous readers of As are done
                patch_to_tile<C>(As, tx, ty, acc);
                __syncthreads();
                zero_acc<C>(acc);
                gemm_nn<C>(As, W1n, acc, tx, ty);      // dpre @ W1  ([o][in])
#pragma unroll
                for (int i = 0; i < RPT; ++i)
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        acc[i][j] = fmaf(att[i][k], dmix[i][j], acc[i][j]);
                        dxres[i][j] += acc[i][j];
                    }
                store_patch<C>(dm_ws + static_cast<long long>(k) * rows * C, row0, live, tx, ty, acc);
                gemm_tn<C>(As, Mt, p_w1, tx, ty);      // dW1 += dpre^T m_k
            }
        }
        if (G.g_x != nullptr) store_patch<C>(G.g_x, row0, live, tx, ty, dxres);
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < CPT; ++j)
            if (ty * RPT < C) atomicAdd(G.g_att_w1 + (ty * RPT + i) * C + tx * CPT + j, p_w1[i][j]);
    __syncthreads();
    float* scratch = Mk;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int col = tx * CPT + j;
        scratch[(ty * 4 + 0) * C + col] = p_b1[j];
        scratch[(ty * 4 + 1) * C + col] = p_w2[j];
        scratch[(ty * 4 + 2) * C + col] = p_gamma[j];
        scratch[(ty * 4 + 3) * C + col] = p_beta[j];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 4 * C; idx += NT) {
        const int which = idx / C, col = idx % C;
        float s = 0.f;
        for (int g = 0; g < 16; ++g) s += scratch[(g * 4 + which) * C + col];
        red[idx] = s;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 4 * C; idx += NT) {
        const int which = idx / C, col = idx % C;
        float* dst = which == 0 ? G.g_att_b1 : (which == 1 ? G.g_att_w2 : (which == 2 ? G.g_ln_gamma : G.g_ln_beta));
        if (dst != nullptr && (which < 2 || P.apply_ln)) atomicAdd(dst + col, red[idx]);
    }
    p_b2 = warp_sum(p_b2);
    if ((threadIdx.x & 31) == 0 && p_b2 != 0.f) atomicAdd(G.g_att_b2, p_b2);
}

// ---------------------------------------------------------------------------------------------
// backward (2): through the conv weight, one message per blockIdx.y
//   dL/dagg_k = scale_k (dL/dm_k) W_k^T ;  g_wprod[k] += agg_k^T (dL/dm_k)
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(NT) combine_bwd_conv_kernel(topo_combine_params P, long long rows,
                                                              const int* __restrict__ n_rows_dev,
                                                              topo_combine_grads G, const float* __restrict__ dm_ws) {
    constexpr int CPT = Cfg<C>::CPT;
    extern __shared__ __align__(16) float smem[];
    float* Wt = smem;                           // [out][in] = W_k transposed
    float* Gs = Wt + C * C;                     // agg tile
    float* Ds = Gs + Cfg<C>::TILE;              // dL/dm tile
    const int k = blockIdx.y;
    const long long live = n_rows_dev ? min(static_cast<long long>(*n_rows_dev), rows) : rows;
    const long long tiles = (live + TM - 1) / TM;
    if (blockIdx.x >= tiles) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    load_matrix<C>(P.w[k], Wt, true);
    const float scale = __ldg(P.scale[k]);
    const float* dm = dm_ws + static_cast<long long>(k) * rows * C;
    float p_w[RPT][CPT];
    zero_acc<C>(p_w);
    __syncthreads();
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row0 = tile * TM;
        load_tile<C>(P.agg[k], row0, live, Gs);
        load_tile<C>(dm, row0, live, Ds);
        __syncthreads();
        float acc[RPT][CPT];
        zero_acc<C>(acc);
        gemm_nn<C>(Ds, Wt, acc, tx, ty);
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[i][j] *= scale;
        store_patch<C>(G.g_agg[k], row0, live, tx, ty, acc);
        gemm_tn<C>(Gs, Ds, p_w, tx, ty);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < CPT; ++j)
            if (ty * RPT < C) atomicAdd(G.g_wprod[k] + (ty * RPT + i) * C + tx * CPT + j, p_w[i][j]);
}

// ---------------------------------------------------------------------------------------------
// Row LayerNorm and the active-embedding scaling
//
// LayerNorm over [rows, C], C = 32 / 64 / 128: C / 4 lanes hold one row as float4s, a warp holds 32 / (C / 4) rows and
// works on two such groups per iteration (all of their loads are issued before the first shuffle), so a few thousand
// warps keep enough bytes in flight for HBM: the decoder consumer normalises 395,200 memory rows five times per
// micro-batch (decoder.py:70-83, 153-156), where a one-row-per-warp kernel reaches a tenth of the bandwidth.
// ---------------------------------------------------------------------------------------------
// 128-bit read-only load, issued where it is written (the two row groups of an iteration are fetched before the first shuffle)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

template <int LPR>
__device__ __forceinline__ float group_sum(float v) {      // over the LPR adjacent lanes that share a row
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int C>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(long long rows, const float* __restrict__ x,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps,
                                                            float* __restrict__ y) {
    constexpr int LPR = C / 4, RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane % LPR, rw = lane / LPR;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + sub);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + sub);
    const long long step = static_cast<long long>(gridDim.x) * 8 * RPW;
    for (long long row0 = (blockIdx.x * 8ll + (threadIdx.x >> 5)) * RPW; row0 < rows; row0 += 2 * step) {
        long long row[2] = {row0 + rw, row0 + step + rw};
        float4 v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
            v[u] = row[u] < rows ? ldg_stream(reinterpret_cast<const float4*>(x + row[u] * C) + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float mean = group_sum<LPR>((v[u].x + v[u].y) + (v[u].z + v[u].w)) * (1.0f / C);
            const float d0 = v[u].x - mean, d1 = v[u].y - mean, d2 = v[u].z - mean, d3 = v[u].w - mean;
            const float rstd = 1.0f / sqrtf(group_sum<LPR>(fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)))) * (1.0f / C) + eps);
            if (row[u] < rows)
                *(reinterpret_cast<float4*>(y + row[u] * C) + sub) =
                    make_float4(fmaf(d0 * rstd, g.x, b.x), fmaf(d1 * rstd, g.y, b.y), fmaf(d2 * rstd, g.z, b.z), fmaf(d3 * rstd, g.w, b.w));
        }
    }
}

template <int C>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(long long rows, const float* __restrict__ x,
                                                            const float* __restrict__ gamma, float eps,
                                                            const float* __restrict__ gy_in,
                                                            float* __restrict__ gx, float* __restrict__ g_gamma,
                                                            float* __restrict__ g_beta, float* __restrict__ cta_partials) {
    constexpr int LPR = C / 4, RPW = 32 / LPR;
    __shared__ float red[8][2][C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % LPR, rw = lane / LPR;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + sub);
    float pg[4] = {0.f, 0.f, 0.f, 0.f}, pb[4] = {0.f, 0.f, 0.f, 0.f};
    const long long step = static_cast<long long>(gridDim.x) * 8 * RPW;
    for (long long row0 = (blockIdx.x * 8ll + warp) * RPW; row0 < rows; row0 += 2 * step) {
        long long row[2] = {row0 + rw, row0 + step + rw};
        float4 v[2], dy[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const bool ok = row[u] < rows;
            v[u] = ok ? ldg_stream(reinterpret_cast<const float4*>(x + row[u] * C) + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
            dy[u] = ok ? ldg_stream(reinterpret_cast<const float4*>(gy_in + row[u] * C) + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float mean = group_sum<LPR>((v[u].x + v[u].y) + (v[u].z + v[u].w)) * (1.0f / C);
            const float d[4] = {v[u].x - mean, v[u].y - mean, v[u].z - mean, v[u].w - mean};
            const float rstd = 1.0f / sqrtf(group_sum<LPR>(fmaf(d[0], d[0], fmaf(d[1], d[1], fmaf(d[2], d[2], d[3] * d[3])))) * (1.0f / C) + eps);
            const float dyv[4] = {dy[u].x, dy[u].y, dy[u].z, dy[u].w};
            const float gv[4] = {g.x, g.y, g.z, g.w};
            float xh[4], gyv[4], c1 = 0.f, c2 = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                xh[k] = d[k] * rstd;
                pg[k] = fmaf(dyv[k], xh[k], pg[k]);        // rows past the end contribute dy = 0
                pb[k] += dyv[k];
                gyv[k] = dyv[k] * gv[k];
                c1 += gyv[k];
                c2 = fmaf(gyv[k], xh[k], c2);
            }
            c1 = group_sum<LPR>(c1) * (1.0f / C);
            c2 = group_sum<LPR>(c2) * (1.0f / C);
            if (row[u] < rows)
                *(reinterpret_cast<float4*>(gx + row[u] * C) + sub) =
                    make_float4(rstd * (gyv[0] - c1 - xh[0] * c2), rstd * (gyv[1] - c1 - xh[1] * c2),
                                rstd * (gyv[2] - c1 - xh[2] * c2), rstd * (gyv[3] - c1 - xh[3] * c2));
        }
    }
    // column sums: over the warp's row groups in shuffles, over the CTA's warps in shared memory, then either this CTA's slot of
    // the partials buffer (added in CTA order by layernorm_bwd_reduce_kernel: reproducible) or one atomic per column and CTA
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
            pg[k] += __shfl_xor_sync(0xffffffffu, pg[k], o);
            pb[k] += __shfl_xor_sync(0xffffffffu, pb[k], o);
        }
        if (rw == 0) {
            red[warp][0][4 * sub + k] = pg[k];
            red[warp][1][4 * sub + k] = pb[k];
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * C; idx += 256) {
        const int which = idx / C, col = idx % C;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][which][col];
        if (cta_partials != nullptr) cta_partials[static_cast<size_t>(blockIdx.x) * 2 * C + idx] = t;
        else atomicAdd((which == 0 ? g_gamma : g_beta) + col, t);
    }
}

// g_gamma / g_beta += the CTA slots in a fixed order: thread (group, column) adds every `groups`-th slot, the groups meet in
// shared memory in index order.  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024) layernorm_bwd_reduce_kernel(const float* __restrict__ cta_partials, int n_ctas, int c2,
                                                                    float* __restrict__ g_gamma, float* __restrict__ g_beta) {
    __shared__ float part[1024];
    const int groups = 1024 / c2, grp = threadIdx.x / c2, idx = threadIdx.x % c2;
    float t = 0.f;
    if (grp < groups)
        for (int b = grp; b < n_ctas; b += groups) t += cta_partials[static_cast<size_t>(b) * c2 + idx];
    part[threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.x < c2) {
        float s = 0.f;
        for (int g = 0; g < groups; ++g) s += part[g * c2 + threadIdx.x];
        float* dst = threadIdx.x < c2 / 2 ? g_gamma + threadIdx.x : g_beta + (threadIdx.x - c2 / 2);
        *dst += s;
    }
}

// X_r[row] = lne[id] * p[id]   for every active simplex of rank r in the batch
template <int VEC>
__global__ void __launch_bounds__(256) embed_fwd_kernel(DeviceTables d, topo_complex_view cv, int r,
                                                        const float* __restrict__ lne, float* __restrict__ x) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= cv.counts[b * 4 + r]) return;
    const long long axis = static_cast<long long>(b) * d.off[4] + d.off[r];
    const int id = cv.act_idx[axis + i];
    const float p = cv.probs[axis + id];
    const long long row = cv.row_off[r * (cv.batch + 1) + b] + i;
    Vec<VEC> v;
    v.load(lne + (static_cast<long long>(id) * 32 + lane) * VEC);
#pragma unroll
    for (int k = 0; k < VEC; ++k) v.v[k] *= p;
    v.store(x + (row * 32 + lane) * VEC);
}

// g_probs[b, id] += <g_x[row], lne[id]>
template <int VEC>
__global__ void __launch_bounds__(256) embed_bwd_probs_kernel(DeviceTables d, topo_complex_view cv, int r,
                                                              const float* __restrict__ lne,
                                                              const float* __restrict__ g_x,
                                                              float* __restrict__ g_probs) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= cv.counts[b * 4 + r]) return;
    const long long axis = static_cast<long long>(b) * d.off[4] + d.off[r];
    const int id = cv.act_idx[axis + i];
    const long long row = cv.row_off[r * (cv.batch + 1) + b] + i;
    Vec<VEC> v, g;
    v.load(lne + (static_cast<long long>(id) * 32 + lane) * VEC);
    g.load(g_x + (row * 32 + lane) * VEC);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) s = fmaf(v.v[k], g.v[k], s);
    s = warp_sum(s);
    if (lane == 0) g_probs[axis + id] += s;
}

// g_lne[id] = sum_b p[b, id] g_x[row(b, id)]   (deterministic: one warp per id walks the batch)
template <int VEC>
__global__ void __launch_bounds__(256) embed_bwd_table_kernel(DeviceTables d, topo_complex_view cv, int r,
                                                              const float* __restrict__ g_x,
                                                              float* __restrict__ g_lne) {
    const int lane = threadIdx.x & 31;
    const int id = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (id >= d.cnt[r]) return;
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    for (int b = 0; b < cv.batch; ++b) {
        const long long axis = static_cast<long long>(b) * d.off[4] + d.off[r];
        const int p_row = cv.pos[axis + id];
        if (p_row < 0) continue;
        const float p = cv.probs[axis + id];
        const long long row = cv.row_off[r * (cv.batch + 1) + b] + p_row;
        Vec<VEC> g;
        g.load(g_x + (row * 32 + lane) * VEC);
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = fmaf(p, g.v[k], acc[k]);
    }
    Vec<VEC> o;
#pragma unroll
    for (int k = 0; k < VEC; ++k) o.v[k] = acc[k];
    o.store(g_lne + (static_cast<long long>(id) * 32 + lane) * VEC);
}

int check_combine(const topo_combine_params* p, int64_t rows) {
    TOPO_REQUIRE(p != nullptr, "params is null");
    TOPO_REQUIRE(rows >= 0, "rows < 0");
    TOPO_REQUIRE(p->n_msgs >= 1 && p->n_msgs <= 3, "n_msgs must be 1..3");
    if (p->channels != 32 && p->channels != 64) {
        set_error("combine kernels are instantiated for channels 32 and 64");
        return TOPO_ERR_UNSUPPORTED;
    }
    for (int k = 0; k < p->n_msgs; ++k) TOPO_REQUIRE(p->agg[k] && p->w[k] && p->scale[k], "null message operand");
    TOPO_REQUIRE(p->att_w1 && p->att_b1 && p->att_w2 && p->att_b2, "null attention parameter");
    TOPO_REQUIRE(!p->apply_ln || (p->ln_gamma && p->ln_beta), "null LayerNorm parameter");
    return TOPO_OK;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    return ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), bytes);
}

template <int C>
size_t fwd_smem() { return sizeof(float) * (4 * C * C + 4 * C + 2 * Cfg<C>::TILE); }
template <int C>
size_t bwd_attn_smem() { return sizeof(float) * (5 * C * C + 4 * C + 4 * Cfg<C>::TILE + 4 * C); }
template <int C>
size_t bwd_conv_smem() { return sizeof(float) * (C * C + 2 * Cfg<C>::TILE); }
template <int C>
size_t bwd_attn_saved_smem() { return sizeof(float) * (C * C + 2 * C + 4 * Cfg<C>::TILE + 4 * C); }

}  // namespace
}  // namespace topo

using namespace topo;

extern "C" int topo_sccn_combine_fwd(const topo_combine_params* p, int64_t rows, const int32_t* n_rows_dev,
                                     float* out, topo_stream_t stream) {
    if (int rc = check_combine(p, rows)) return rc;
    TOPO_REQUIRE(out != nullptr, "out is null");
    if (rows == 0) return TOPO_OK;
    const int tiles = static_cast<int>((rows + TM - 1) / TM);
    cudaStream_t s = as_stream(stream);
    if (p->channels == 64) {
        const int grid = std::min(tiles, sm_count() * 2);
        if (int rc = set_smem(combine_fwd_kernel<64>, fwd_smem<64>())) return rc;
        combine_fwd_kernel<64><<<grid, NT, fwd_smem<64>(), s>>>(*p, rows, n_rows_dev, out);
    } else {
        const int grid = std::min(tiles, sm_count() * 4);
        if (int rc = set_smem(combine_fwd_kernel<32>, fwd_smem<32>())) return rc;
        combine_fwd_kernel<32><<<grid, NT, fwd_smem<32>(), s>>>(*p, rows, n_rows_dev, out);
    }
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

static int check_combine_bwd(const topo_combine_params* p, int64_t rows, const float* grad_out,
                             const topo_combine_grads* g, const float* workspace, bool attn) {
    if (int rc = check_combine(p, rows)) return rc;
    TOPO_REQUIRE(g && workspace, "null argument");
    if (attn) {
        TOPO_REQUIRE(grad_out != nullptr, "grad_out is null");
        TOPO_REQUIRE(g->g_att_w1 && g->g_att_b1 && g->g_att_w2 && g->g_att_b2, "null attention gradient buffer");
        TOPO_REQUIRE(!p->apply_ln || (g->g_ln_gamma && g->g_ln_beta), "null LayerNorm gradient buffer");
    } else {
        for (int k = 0; k < p->n_msgs; ++k) TOPO_REQUIRE(g->g_agg[k] && g->g_wprod[k], "null message gradient buffer");
    }
    return TOPO_OK;
}

extern "C" int topo_sccn_combine_bwd_attention(const topo_combine_params* p, int64_t rows, const int32_t* n_rows_dev,
                                               const float* grad_out, const topo_combine_grads* g, float* workspace,
                                               topo_stream_t stream) {
    if (int rc = check_combine_bwd(p, rows, grad_out, g, workspace, true)) return rc;
    if (rows == 0) return TOPO_OK;
    const int tiles = static_cast<int>((rows + TM - 1) / TM);
    cudaStream_t s = as_stream(stream);
    bool saved = true;
    for (int k = 0; k < p->n_msgs; ++k) saved = saved && p->saved_m[k] && p->saved_pre[k];
    if (saved) {
        if (p->channels == 64) {
            if (int rc = set_smem(combine_bwd_attn_saved_kernel<64>, bwd_attn_saved_smem<64>())) return rc;
            combine_bwd_attn_saved_kernel<64><<<std::min(tiles, sm_count() * 2), NT, bwd_attn_saved_smem<64>(), s>>>(
                *p, rows, n_rows_dev, grad_out, *g, workspace);
        } else {
            if (int rc = set_smem(combine_bwd_attn_saved_kernel<32>, bwd_attn_saved_smem<32>())) return rc;
            combine_bwd_attn_saved_kernel<32><<<std::min(tiles, sm_count() * 4), NT, bwd_attn_saved_smem<32>(), s>>>(
                *p, rows, n_rows_dev, grad_out, *g, workspace);
        }
        TOPO_LAUNCH_CHECK();
        return TOPO_OK;
    }
    if (p->channels == 64) {
        if (int rc = set_smem(combine_bwd_attn_kernel<64>, bwd_attn_smem<64>())) return rc;
        combine_bwd_attn_kernel<64><<<std::min(tiles, sm_count()), NT, bwd_attn_smem<64>(), s>>>(
            *p, rows, n_rows_dev, grad_out, *g, workspace);
    } else {
        if (int rc = set_smem(combine_bwd_attn_kernel<32>, bwd_attn_smem<32>())) return rc;
        combine_bwd_attn_kernel<32><<<std::min(tiles, sm_count() * 2), NT, bwd_attn_smem<32>(), s>>>(
            *p, rows, n_rows_dev, grad_out, *g, workspace);
    }
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_sccn_combine_bwd_conv(const topo_combine_params* p, int64_t rows, const int32_t* n_rows_dev,
                                          const topo_combine_grads* g, const float* workspace, topo_stream_t stream) {
    if (int rc = check_combine_bwd(p, rows, nullptr, g, workspace, false)) return rc;
    if (rows == 0) return TOPO_OK;
    const int tiles = static_cast<int>((rows + TM - 1) / TM);
    cudaStream_t s = as_stream(stream);
    if (p->channels == 64) {
        if (int rc = set_smem(combine_bwd_conv_kernel<64>, bwd_conv_smem<64>())) return rc;
        combine_bwd_conv_kernel<64><<<dim3(std::min(tiles, sm_count()), p->n_msgs), NT, bwd_conv_smem<64>(), s>>>(
            *p, rows, n_rows_dev, *g, workspace);
    } else {
        if (int rc = set_smem(combine_bwd_conv_kernel<32>, bwd_conv_smem<32>())) return rc;
        combine_bwd_conv_kernel<32><<<dim3(std::min(tiles, sm_count() * 2), p->n_msgs), NT, bwd_conv_smem<32>(), s>>>(
            *p, rows, n_rows_dev, *g, workspace);
    }
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_sccn_combine_bwd(const topo_combine_params* p, int64_t rows, const int32_t* n_rows_dev,
                                     const float* grad_out, const topo_combine_grads* g, float* workspace,
                                     topo_stream_t stream) {
    if (int rc = topo_sccn_combine_bwd_attention(p, rows, n_rows_dev, grad_out, g, workspace, stream)) return rc;
    return topo_sccn_combine_bwd_conv(p, rows, n_rows_dev, g, workspace, stream);
}

#define DISPATCH_VEC(channels, CALL)                      \
    switch (channels) {                                   \
        case 32: { constexpr int VEC = 1; CALL; } break;  \
        case 64: { constexpr int VEC = 2; CALL; } break;  \
        case 128: { constexpr int VEC = 4; CALL; } break; \
        default:                                          \
            set_error("channels must be 32, 64 or 128");  \
            return TOPO_ERR_UNSUPPORTED;                  \
    }

extern "C" int topo_layernorm_fwd(int64_t rows, int channels, const float* x, const float* gamma,
                                  const float* beta, float eps, float* y, topo_stream_t stream) {
    TOPO_REQUIRE(rows >= 0 && x && gamma && beta && y, "bad argument");
    if (rows == 0) return TOPO_OK;
    TOPO_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                   reinterpret_cast<uintptr_t>(beta)) & 15) == 0, "buffers must be 16-byte aligned");
    const int grid = static_cast<int>(std::min<int64_t>((rows + 7) / 8, sm_count() * 8));
    DISPATCH_VEC(channels, (layernorm_fwd_kernel<VEC * 32><<<grid, 256, 0, as_stream(stream)>>>(rows, x, gamma, beta, eps, y)));
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

static int layernorm_bwd_grid(int64_t rows) { return static_cast<int>(std::min<int64_t>((rows + 7) / 8, sm_count() * 4)); }

extern "C" int64_t topo_layernorm_bwd_workspace_floats(int64_t rows, int channels) {
    if (rows < 0 || (channels != 32 && channels != 64 && channels != 128)) return -1;
    return static_cast<int64_t>(layernorm_bwd_grid(rows)) * 2 * channels;
}

extern "C" int topo_layernorm_bwd(int64_t rows, int channels, const float* x, const float* gamma, float eps,
                                  const float* grad_y, float* grad_x, float* grad_gamma, float* grad_beta,
                                  float* workspace, topo_stream_t stream) {
    TOPO_REQUIRE(rows >= 0 && x && gamma && grad_y && grad_x && grad_gamma && grad_beta, "bad argument");
    if (rows == 0) return TOPO_OK;
    TOPO_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(grad_y) | reinterpret_cast<uintptr_t>(grad_x) |
                   reinterpret_cast<uintptr_t>(gamma)) & 15) == 0, "buffers must be 16-byte aligned");
    const int grid = layernorm_bwd_grid(rows);
    DISPATCH_VEC(channels, (layernorm_bwd_kernel<VEC * 32><<<grid, 256, 0, as_stream(stream)>>>(
                               rows, x, gamma, eps, grad_y, grad_x, grad_gamma, grad_beta, workspace)));
    if (workspace != nullptr)
        layernorm_bwd_reduce_kernel<<<1, 1024, 0, as_stream(stream)>>>(workspace, grid, 2 * channels, grad_gamma, grad_beta);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_embed_fwd(const topo_tables* t, const topo_complex_view* cv, int rank, int channels,
                              const float* lne, float* x_out, topo_stream_t stream) {
    TOPO_REQUIRE(t && cv && lne && x_out, "null argument");
    TOPO_REQUIRE_TABLES_DEVICE(t);
    TOPO_REQUIRE(rank >= 0 && rank <= 3, "rank out of range");
    if (cv->batch == 0 || t->d.cnt[rank] == 0) return TOPO_OK;
    const dim3 grid((t->d.cnt[rank] + 7) / 8, static_cast<unsigned>(cv->batch));
    DISPATCH_VEC(channels, (embed_fwd_kernel<VEC><<<grid, 256, 0, as_stream(stream)>>>(t->d, *cv, rank, lne, x_out)));
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_embed_bwd(const topo_tables* t, const topo_complex_view* cv, int rank, int channels,
                              const float* lne, const float* grad_x, float* grad_lne, float* grad_probs,
                              topo_stream_t stream) {
    TOPO_REQUIRE(t && cv && lne && grad_x && grad_lne && grad_probs, "null argument");
    TOPO_REQUIRE_TABLES_DEVICE(t);
    TOPO_REQUIRE(rank >= 0 && rank <= 3, "rank out of range");
    if (t->d.cnt[rank] == 0) return TOPO_OK;
    cudaStream_t s = as_stream(stream);
    if (cv->batch > 0) {
        const dim3 grid((t->d.cnt[rank] + 7) / 8, static_cast<unsigned>(cv->batch));
        DISPATCH_VEC(channels, (embed_bwd_probs_kernel<VEC><<<grid, 256, 0, s>>>(t->d, *cv, rank, lne, grad_x, grad_probs)));
    }
    DISPATCH_VEC(channels, (embed_bwd_table_kernel<VEC><<<(t->d.cnt[rank] + 7) / 8, 256, 0, s>>>(t->d, *cv, rank, grad_x, grad_lne)));
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
