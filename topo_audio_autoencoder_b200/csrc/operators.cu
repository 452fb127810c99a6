// B2: explicit weighted incidence / adjacency operators of ONE sample's active sub-complex.
// Replaces build_sparse_matrices (complex_builder.py:23-115): instead of dense n_r x n_r products
// followed by nonzero(), every row walks a static, neighbour-sorted slot list and emits the entries
// whose single-product value is non-zero.  Row-major sorted COO comes out directly.
//
//   op 0..3: adjacency rank_0..rank_3      op 4..6: incidence rank_1..rank_3
//   A0[v,v'] = p_e          (complex_builder.py:35-47)
//   I_r[f,s] = p_s          (complex_builder.py:52-59)
//   A1 = I2 I2^T, A2 = I3 I3^T, A3 = I3^T I3, diagonal removed (complex_builder.py:62-70):
//   A1[e,e'] = p_t*p_t,  A2[t,t'] = p_s*p_s,  A3[s,s'] = p_s*p_s'   (one product each)
#include "common.cuh"

namespace topo {
namespace {

__host__ __device__ __forceinline__ int op_rank(int op) { return op < 4 ? op : op - 4; }

// Visit the candidate entries of compact row i of operator op, in ascending column order.
// f(col, value, a, b): a / b are simplex-axis indices of the (at most two) probabilities the value
// is the product of (b < 0 when the value is a single probability; a == b for a square).
template <typename F>
__device__ __forceinline__ void visit_row(const DeviceTables& d, int op, int i, const float* __restrict__ probs,
                                          const int* __restrict__ pos, const int* __restrict__ act_idx, F&& f) {
    const int r = op_rank(op);
    const int id = act_idx[d.off[r] + i];
    if (op < 4) {
        const int w = d.adj_w[r];
        const int* nbr = d.adj_nbr[r] + static_cast<long long>(id) * w;
        const int* via = d.adj_via[r] + static_cast<long long>(id) * w;
        for (int j = 0; j < w; ++j) {
            const int c = __ldg(nbr + j), v = __ldg(via + j);
            const int col = pos[d.off[r] + c];
            if (col < 0) continue;
            if (op == 0) {
                const int a = d.off[1] + v;
                f(col, probs[a], a, -1);
            } else if (op == 3) {
                if (pos[d.off[2] + v] < 0) continue;              // shared triangle must be active
                const int a = d.off[3] + id, b = d.off[3] + c;
                f(col, probs[a] * probs[b], a, b);
            } else {
                const int a = d.off[r + 1] + v;                    // shared coface must be active
                if (pos[a] < 0) continue;
                f(col, probs[a] * probs[a], a, a);
            }
        }
    } else {
        const int w = d.ncof[r];
        const int* cof = d.cofaces[r] + static_cast<long long>(id) * w;
        for (int j = 0; j < w; ++j) {
            const int a = d.off[r + 1] + __ldg(cof + j);
            const int col = pos[a];
            if (col < 0) continue;
            f(col, probs[a], a, -1);
        }
    }
}

__global__ void __launch_bounds__(128) operators_count_kernel(DeviceTables d, const float* __restrict__ probs,
                                                              const int* __restrict__ pos,
                                                              const int* __restrict__ act_idx,
                                                              const int* __restrict__ counts, int max_rows,
                                                              int* __restrict__ row_ptr) {
    const int op = blockIdx.y;
    const int rows = counts[op_rank(op)];
    int* out = row_ptr + static_cast<long long>(op) * (max_rows + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
        int n = 0;
        visit_row(d, op, i, probs, pos, act_idx, [&](int, float v, int, int) { n += (v != 0.0f); });
        out[i + 1] = n;
    }
}

// in-place inclusive scan of out[1..rows], out[0] = 0; total also stored at out[max_rows]
__global__ void __launch_bounds__(1024) operators_scan_kernel(const int* __restrict__ counts, int max_rows,
                                                              int* __restrict__ row_ptr) {
    const int op = blockIdx.x;
    const int rows = counts[op_rank(op)];
    int* out = row_ptr + static_cast<long long>(op) * (max_rows + 1);
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) { carry = 0; out[0] = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < rows; start += 1024) {
        const int i = start + threadIdx.x;
        int v = (i < rows) ? out[i + 1] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (lane == 31) warp_tot[warp] = v;
        __syncthreads();
        int before = carry;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        if (i < rows) out[i + 1] = before + v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[max_rows] = carry;
}

struct OpBuffers {
    long long* rows[7];
    long long* cols[7];
    float* vals[7];
};

__global__ void __launch_bounds__(128) operators_fill_kernel(DeviceTables d, const float* __restrict__ probs,
                                                             const int* __restrict__ pos,
                                                             const int* __restrict__ act_idx,
                                                             const int* __restrict__ counts, int max_rows,
                                                             const int* __restrict__ row_ptr, OpBuffers out) {
    const int op = blockIdx.y;
    const int rows = counts[op_rank(op)];
    const int* rp = row_ptr + static_cast<long long>(op) * (max_rows + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
        int k = rp[i];
        visit_row(d, op, i, probs, pos, act_idx, [&](int col, float v, int, int) {
            if (v != 0.0f) {
                out.rows[op][k] = i;
                out.cols[op][k] = col;
                out.vals[op][k] = v;
                ++k;
            }
        });
    }
}

struct OpGrads {
    const float* g[7];
};

__global__ void __launch_bounds__(128) operators_bwd_kernel(DeviceTables d, const float* __restrict__ probs,
                                                            const int* __restrict__ pos,
                                                            const int* __restrict__ act_idx,
                                                            const int* __restrict__ counts, int max_rows,
                                                            const int* __restrict__ row_ptr, OpGrads grads,
                                                            float* __restrict__ grad_probs) {
    const int op = blockIdx.y;
    if (grads.g[op] == nullptr) return;
    const int rows = counts[op_rank(op)];
    const int* rp = row_ptr + static_cast<long long>(op) * (max_rows + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
        int k = rp[i];
        visit_row(d, op, i, probs, pos, act_idx, [&](int, float v, int a, int b) {
            if (v != 0.0f) {
                const float g = grads.g[op][k++];
                if (b < 0) {
                    atomicAdd(grad_probs + a, g);
                } else if (a == b) {
                    const float t = g * probs[a];
                    atomicAdd(grad_probs + a, t + t);
                } else {
                    atomicAdd(grad_probs + a, g * probs[b]);
                    atomicAdd(grad_probs + b, g * probs[a]);
                }
            }
        });
    }
}

int check_common(const topo_tables* t, int64_t max_rows) {
    TOPO_REQUIRE(t != nullptr, "tables is null");
    TOPO_REQUIRE_TABLES_DEVICE(t);
    TOPO_REQUIRE(max_rows >= 1, "max_rows must be >= 1");
    if (t->d.adj_w[0] < 0) {
        set_error("operator builder tables were not built for this vertex count (too large)");
        return TOPO_ERR_UNSUPPORTED;
    }
    return TOPO_OK;
}

}  // namespace
}  // namespace topo

using namespace topo;

extern "C" int topo_operators_count(const topo_tables* t, const float* probs, const int32_t* pos,
                                    const int32_t* act_idx, const int32_t* counts, int64_t max_rows,
                                    int32_t* row_ptr, topo_stream_t stream) {
    if (int rc = check_common(t, max_rows)) return rc;
    TOPO_REQUIRE(probs && pos && act_idx && counts && row_ptr, "null argument");
    cudaStream_t s = as_stream(stream);
    TOPO_CUDA(cudaMemsetAsync(row_ptr, 0, sizeof(int32_t) * 7 * (max_rows + 1), s));
    const unsigned gx = static_cast<unsigned>((max_rows + 127) / 128);
    operators_count_kernel<<<dim3(gx, 7), 128, 0, s>>>(t->d, probs, pos, act_idx, counts, static_cast<int>(max_rows), row_ptr);
    operators_scan_kernel<<<7, 1024, 0, s>>>(counts, static_cast<int>(max_rows), row_ptr);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_operators_fill(const topo_tables* t, const float* probs, const int32_t* pos,
                                   const int32_t* act_idx, const int32_t* counts, int64_t max_rows,
                                   const int32_t* row_ptr, int64_t* const dev_rows[7], int64_t* const dev_cols[7],
                                   float* const dev_vals[7], topo_stream_t stream) {
    if (int rc = check_common(t, max_rows)) return rc;
    TOPO_REQUIRE(probs && pos && act_idx && counts && row_ptr && dev_rows && dev_cols && dev_vals, "null argument");
    OpBuffers ob;
    for (int i = 0; i < 7; ++i) {
        ob.rows[i] = reinterpret_cast<long long*>(dev_rows[i]);
        ob.cols[i] = reinterpret_cast<long long*>(dev_cols[i]);
        ob.vals[i] = dev_vals[i];
    }
    const unsigned gx = static_cast<unsigned>((max_rows + 127) / 128);
    operators_fill_kernel<<<dim3(gx, 7), 128, 0, as_stream(stream)>>>(t->d, probs, pos, act_idx, counts,
                                                                      static_cast<int>(max_rows), row_ptr, ob);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_operators_bwd(const topo_tables* t, const float* probs, const int32_t* pos,
                                  const int32_t* act_idx, const int32_t* counts, int64_t max_rows,
                                  const int32_t* row_ptr, const float* const dev_grad_vals[7], float* grad_probs,
                                  topo_stream_t stream) {
    if (int rc = check_common(t, max_rows)) return rc;
    TOPO_REQUIRE(probs && pos && act_idx && counts && row_ptr && dev_grad_vals && grad_probs, "null argument");
    OpGrads og;
    for (int i = 0; i < 7; ++i) og.g[i] = dev_grad_vals[i];
    const unsigned gx = static_cast<unsigned>((max_rows + 127) / 128);
    operators_bwd_kernel<<<dim3(gx, 7), 128, 0, as_stream(stream)>>>(t->d, probs, pos, act_idx, counts,
                                                                     static_cast<int>(max_rows), row_ptr, og, grad_probs);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
