// S1 (a): neighbourhood aggregation for SCCN, forward and backward.
//
// Structured path (whole batch, matrix-free).  The reference applies ten sparse operators per layer
// (custom_sccn.py:78-81, 95-98, 113-116) that build_sparse_matrices materialised
// (complex_builder.py:52-70).  Here nothing is materialised: a group of C/4 lanes owns one active
// simplex (one float4 per lane, a 128-bit coalesced row), walks its static face / coface lists,
// weights the gathered rows with the rectified probabilities and writes one C-wide row.
// Every adjacency factors through the incidences (A1 = I2 I2^T, A2 = I3 I3^T, A3 = I3^T I3,
// complex_builder.py:62-64), so the same-rank aggregates are incidence gathers of the cross-rank
// aggregates minus a diagonal term:
//     down[r] = I_{r+1} X_{r+1}      up[r] = I_r^T X_{r-1}
//     same[1] = I_2 up[2] - q1 . X1          q1[e] = sum_{t > e} p_t^2
//     same[2] = I_3 up[3] - q2 . X2          q2[t] = sum_{s > t} p_s^2
//     same[3] = I_3^T down[2] - c p^2 . X3   c[s]  = number of active faces
// 88,920 row gathers per sample-layer for the full 20-vertex complex instead of 421,800.
// The gathers are L2-resident (one sample's features are 1.6 MB) and latency bound, so the loops are
// branch-free (predicated loads) and issue four independent 128-bit row loads per group at a time.
// All rows are single-owner: the backward needs no atomics and is deterministic.
//
// Generic path: CSR SpMM / SDDMM for caller-supplied operators.
#include <algorithm>

#include "common.cuh"

namespace topo {
namespace {

constexpr int kThreads = 256;

struct Feat {
    const float* p[4];
};
struct FeatMut {
    float* p[4];
};

struct Sections {
    int begin[5];   // first block of each rank's section (blocks are (rank, sample, row-chunk))
    int chunks[4];  // row chunks per sample for each rank
};

// V float4s per lane: lane l of a row's group holds float4 number l, l + L, ... of the row (every 128-bit access of
// the group is one contiguous run of 16 L bytes)
template <int V>
struct FV {
    float v[4 * V];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int k = 0; k < 4 * V; ++k) v[k] = 0.f;
    }
    __device__ __forceinline__ void fma(float w, const FV& x) {
#pragma unroll
        for (int k = 0; k < 4 * V; ++k) v[k] = fmaf(w, x.v[k], v[k]);
    }
    __device__ __forceinline__ void add(const FV& x) {
#pragma unroll
        for (int k = 0; k < 4 * V; ++k) v[k] += x.v[k];
    }
    __device__ __forceinline__ float dot(const FV& x) const {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 4 * V; ++k) s = fmaf(v[k], x.v[k], s);
        return s;
    }
};

// One group's view of its row.  L = lanes per row, V = float4s per lane: C = 4 L V.
template <int L, int V>
struct Group {
    using F4 = FV<V>;
    const float* probs;   // this sample's simplex axis
    const int* pos;
    int ro[4];            // compact row offset of this sample per rank
    int lane;             // lane within the group
    unsigned mask;        // the whole warp takes part in every shuffle
    unsigned dense;       // bit r: all simplices of rank r are active in this sample (the index lookups are identities)
    bool valid;           // this group owns a live row

    __device__ __forceinline__ F4 load(const float* base, int row) const {
        const float4* p = reinterpret_cast<const float4*>(base) + static_cast<long long>(row) * (L * V) + lane;
        F4 r;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float4 t = __ldg(p + j * L);
            r.v[4 * j] = t.x; r.v[4 * j + 1] = t.y; r.v[4 * j + 2] = t.z; r.v[4 * j + 3] = t.w;
        }
        return r;
    }
    __device__ __forceinline__ F4 load_if(bool ok, const float* base, int row) const {
        F4 r;
        r.zero();
        if (ok) r = load(base, row);
        return r;
    }
    __device__ __forceinline__ F4 load_rw(const float* base, int row) const {   // not through the read-only path
        const float4* p = reinterpret_cast<const float4*>(base) + static_cast<long long>(row) * (L * V) + lane;
        F4 r;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float4 t = p[j * L];
            r.v[4 * j] = t.x; r.v[4 * j + 1] = t.y; r.v[4 * j + 2] = t.z; r.v[4 * j + 3] = t.w;
        }
        return r;
    }
    __device__ __forceinline__ void store(float* base, int row, const F4& x) const {
        if (valid) {
            float4* p = reinterpret_cast<float4*>(base) + static_cast<long long>(row) * (L * V) + lane;
#pragma unroll
            for (int j = 0; j < V; ++j) p[j * L] = make_float4(x.v[4 * j], x.v[4 * j + 1], x.v[4 * j + 2], x.v[4 * j + 3]);
        }
    }
    __device__ __forceinline__ void accumulate(float* base, int row, const F4& x) const {
        if (valid) {
            float4* p = reinterpret_cast<float4*>(base) + static_cast<long long>(row) * (L * V) + lane;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                float4 t = p[j * L];
                t.x += x.v[4 * j]; t.y += x.v[4 * j + 1]; t.z += x.v[4 * j + 2]; t.w += x.v[4 * j + 3];
                p[j * L] = t;
            }
        }
    }
    __device__ __forceinline__ float group_sum(float v) const {
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, L);
        return v;
    }
    __device__ __forceinline__ int row_of(const DeviceTables& d, int r, int id) const {
        if (dense & (1u << r)) return ro[r] + id;       // every simplex of the rank is active: position == id
        const int p = pos[d.off[r] + id];
        return p < 0 ? -1 : ro[r] + p;
    }
};

// f(ok, row, p_s) for every coface slot of simplex `id` (rank R), four slots at a time; ok == false
// marks an inactive coface or a padding slot (its load is predicated off).
template <int L, int V, int R, typename F>
__device__ __forceinline__ void for_cofaces(const DeviceTables& d, const Group<L, V>& g, int id, F&& f) {
    const int w = d.ncof[R];
    const int* cof = d.cofaces[R] + static_cast<long long>(id) * w;
    for (int jb = 0; jb < w; jb += L) {
        const int j = jb + g.lane;
        int row = -1;
        float ps = 0.f;
        if (j < w) {
            const int s = __ldg(cof + j);
            ps = g.probs[d.off[R + 1] + s];
            row = (ps != 0.0f) ? g.row_of(d, R + 1, s) : -1;
        }
        const int m = min(L, w - jb);
        for (int j4 = 0; j4 < m; j4 += 4) {
            int rr[4];
            float pp[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                rr[q] = __shfl_sync(g.mask, row, (j4 + q) & (L - 1), L);
                pp[q] = __shfl_sync(g.mask, ps, (j4 + q) & (L - 1), L);
                if (j4 + q >= m) rr[q] = -1;
            }
            f(rr, pp);
        }
    }
}

// rows of the (R+1) faces of simplex `id` (rank R >= 1); -1 for an inactive face
template <int L, int V, int R>
__device__ __forceinline__ void face_rows(const DeviceTables& d, const Group<L, V>& g, int id, int (&rows)[R + 1]) {
    const int* fc = d.faces[R] + static_cast<long long>(id) * (R + 1);
#pragma unroll
    for (int a = 0; a <= R; ++a) rows[a] = g.row_of(d, R - 1, __ldg(fc + a));
}

template <int L, int V>
__device__ __forceinline__ void gather4(const Group<L, V>& g, const float* base, const int (&rr)[4], FV<V> (&v)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = g.load_if(rr[q] >= 0, base, rr[q]);
}

// Decode blockIdx.x -> (rank, sample, local row) for this group.  Groups past the live rows stay in the
// kernel (the shuffles are warp-wide) with valid == false and id == 0.
// kDense: test whether whole ranks are active and skip the position / id lookups for them (forward kernels; the
// backward kernels run at 32 registers and have none to spare for the flag).
// Returns false when not a single group of this block owns a live row (a sparse complex: the chunk lies past the
// sample's active count); the whole block then leaves at once -- nothing it could compute is ever stored.
template <int L, int V, bool kDense>
__device__ __forceinline__ bool locate(const DeviceTables& d, const Sections& sec, const topo_complex_view& cv,
                                       Group<L, V>* g, int* rank, int* row, int* id, long long* axis, int first_block = 0) {
    constexpr int kGroups = kThreads / L;
    const int blk = blockIdx.x + first_block;          // a launch may cover a suffix of the sections (one rank)
    const int r = (blk >= sec.begin[1]) + (blk >= sec.begin[2]) + (blk >= sec.begin[3]);
    // selects instead of sec.begin[r]: a runtime index would copy the parameter struct to local memory
    const int begin = r == 0 ? sec.begin[0] : (r == 1 ? sec.begin[1] : (r == 2 ? sec.begin[2] : sec.begin[3]));
    const int chunks = r == 0 ? sec.chunks[0] : (r == 1 ? sec.chunks[1] : (r == 2 ? sec.chunks[2] : sec.chunks[3]));
    const int off_r = r == 0 ? d.off[0] : (r == 1 ? d.off[1] : (r == 2 ? d.off[2] : d.off[3]));
    const int rel = blk - begin;
    // rel / chunks without the integer-division sequence: rel < 2^24 and (rel + 0.5) / chunks is at least 0.5 / chunks
    // away from an integer, far more than the rounding of one fp32 division
    const int b = kDense ? __float2int_rz(__fdividef(static_cast<float>(rel) + 0.5f, static_cast<float>(chunks))) : rel / chunks;
    const int i = (rel - b * chunks) * kGroups + threadIdx.x / L;
    const long long ax = static_cast<long long>(b) * d.off[4];
    g->probs = cv.probs + ax;
    g->pos = cv.pos + ax;
    g->lane = threadIdx.x % L;
    g->mask = 0xffffffffu;
    g->dense = 0u;
    int live_r;
    if constexpr (kDense) {
        const int4 cnt = __ldg(reinterpret_cast<const int4*>(cv.counts) + b);
        g->dense = (cnt.x == d.cnt[0] ? 1u : 0u) | (cnt.y == d.cnt[1] ? 2u : 0u) | (cnt.z == d.cnt[2] ? 4u : 0u) | (cnt.w == d.cnt[3] ? 8u : 0u);
        live_r = r == 0 ? cnt.x : (r == 1 ? cnt.y : (r == 2 ? cnt.z : cnt.w));
    } else {
        live_r = cv.counts[b * 4 + r];
    }
    g->valid = i < live_r;
    // block-uniform: the first row of this chunk against the sample's live count of the rank
    const int chunk_first = (rel - b * chunks) * kGroups;
    const int B1 = static_cast<int>(cv.batch) + 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) g->ro[q] = cv.row_off[q * B1 + b];
    const int ro_r = r == 0 ? g->ro[0] : (r == 1 ? g->ro[1] : (r == 2 ? g->ro[2] : g->ro[3]));
    *rank = r;
    *row = ro_r + (g->valid ? i : 0);
    *id = g->valid ? ((g->dense & (1u << r)) ? i : cv.act_idx[ax + off_r + i]) : 0;
    *axis = ax;
    return chunk_first < live_r;
}

// ------------------------------------------------------------------ forward, phase 1: cross-rank
template <int L, int V, int R>
__device__ __forceinline__ void cross_fwd_body(const DeviceTables& d, const Group<L, V>& g, int row, int id, const Feat& x,
                                               const FeatMut& down, const FeatMut& up) {
    if constexpr (R < 3) {             // down[R] = sum over cofaces p_s X_{R+1}[s]
        if (d.cnt[R + 1] > 0) {
            FV<V> acc;
            acc.zero();
            for_cofaces<L, V, R>(d, g, id, [&](const int (&rr)[4], const float (&pp)[4]) {
                FV<V> v[4];
                gather4<L, V>(g, x.p[R + 1], rr, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) acc.fma(pp[q], v[q]);
            });
            g.store(down.p[R], row, acc);
        }
    }
    if constexpr (R > 0) {             // up[R] = p_id * sum over active faces X_{R-1}[f]
        int fr[R + 1];
        face_rows<L, V, R>(d, g, id, fr);
        FV<V> acc;
        acc.zero();
        FV<V> v[R + 1];
#pragma unroll
        for (int a = 0; a <= R; ++a) v[a] = g.load_if(fr[a] >= 0, x.p[R - 1], fr[a]);
#pragma unroll
        for (int a = 0; a <= R; ++a) acc.add(v[a]);
        const float p = g.probs[d.off[R] + id];
#pragma unroll
        for (int k = 0; k < 4 * V; ++k) acc.v[k] *= p;
        g.store(up.p[R], row, acc);
    }
}

template <int L, int V>
__global__ void __launch_bounds__(kThreads, 6) agg_cross_fwd(const DeviceTables d, const Sections sec,
                                                          const topo_complex_view cv, const Feat x, const FeatMut down,
                                                          const FeatMut up) {
    Group<L, V> g;
    int r, row, id;
    long long axis;
    if (!locate<L, V, true>(d, sec, cv, &g, &r, &row, &id, &axis)) return;
    switch (r) {
        case 0: cross_fwd_body<L, V, 0>(d, g, row, id, x, down, up); break;
        case 1: cross_fwd_body<L, V, 1>(d, g, row, id, x, down, up); break;
        case 2: cross_fwd_body<L, V, 2>(d, g, row, id, x, down, up); break;
        default: cross_fwd_body<L, V, 3>(d, g, row, id, x, down, up); break;
    }
}

// ------------------------------------------------------------------ forward, phase 2: same-rank
// A0[v,v'] = p_e: walk the vertex's edges, gather the other endpoint.  f(rows[4], p_e[4]).
template <int L, int V, typename F>
__device__ __forceinline__ void for_vertex_neighbours(const DeviceTables& d, const Group<L, V>& g, int id, F&& f) {
    const int w = d.ncof[0];
    const int* cof = d.cofaces[0] + static_cast<long long>(id) * w;
    for (int jb = 0; jb < w; jb += L) {
        const int j = jb + g.lane;
        int orow = -1;
        float pe = 0.f;
        if (j < w) {
            const int e = __ldg(cof + j);
            pe = g.probs[d.off[1] + e];
            const int2 ends = __ldg(reinterpret_cast<const int2*>(d.faces[1]) + e);
            orow = (pe != 0.0f) ? g.row_of(d, 0, ends.x == id ? ends.y : ends.x) : -1;
        }
        const int m = min(L, w - jb);
        for (int j4 = 0; j4 < m; j4 += 4) {
            int rr[4];
            float pp[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                rr[q] = __shfl_sync(g.mask, orow, (j4 + q) & (L - 1), L);
                pp[q] = __shfl_sync(g.mask, pe, (j4 + q) & (L - 1), L);
                if (j4 + q >= m) rr[q] = -1;
            }
            f(rr, pp);
        }
    }
}

template <int L, int V, int R>
__device__ __forceinline__ void same_fwd_body(const DeviceTables& d, const Group<L, V>& g, int row, int id, const Feat& x,
                                              const Feat& down, const Feat& up, const FeatMut& same) {
    FV<V> acc;
    acc.zero();
    if constexpr (R == 0) {
        for_vertex_neighbours<L, V>(d, g, id, [&](const int (&rr)[4], const float (&pp)[4]) {
            FV<V> v[4];
            gather4<L, V>(g, x.p[0], rr, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc.fma(pp[q], v[q]);
        });
    } else if constexpr (R < 3) {
        // same[R] = sum_{s > id} p_s up[R+1][s]  -  (sum p_s^2) X_R[id]
        float qsum = 0.f;
        if (d.cnt[R + 1] > 0)
            for_cofaces<L, V, R>(d, g, id, [&](const int (&rr)[4], const float (&pp)[4]) {
                FV<V> v[4];
                gather4<L, V>(g, up.p[R + 1], rr, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc.fma(pp[q], v[q]);
                    if (rr[q] >= 0) qsum = fmaf(pp[q], pp[q], qsum);
                }
            });
        const FV<V> own = g.load_if(g.valid, x.p[R], row);
#pragma unroll
        for (int k = 0; k < 4 * V; ++k) acc.v[k] = fmaf(-qsum, own.v[k], acc.v[k]);
    } else {
        // same[3] = p (sum_{t < id, active} down[2][t] - c p X_3[id])
        int fr[4];
        face_rows<L, V, 3>(d, g, id, fr);
        FV<V> v[4];
        gather4<L, V>(g, down.p[2], fr, v);
        int n_faces = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            acc.add(v[a]);
            n_faces += fr[a] >= 0;
        }
        const float p = g.probs[d.off[3] + id];
        const FV<V> own = g.load_if(g.valid, x.p[3], row);
        const float cp = static_cast<float>(n_faces) * p;
#pragma unroll
        for (int k = 0; k < 4 * V; ++k) acc.v[k] = p * fmaf(-cp, own.v[k], acc.v[k]);
    }
    g.store(same.p[R], row, acc);
}

template <int L, int V>
__global__ void __launch_bounds__(kThreads, 6) agg_same_fwd(const DeviceTables d, const Sections sec,
                                                         const topo_complex_view cv, const Feat x, const Feat down,
                                                         const Feat up, const FeatMut same) {
    Group<L, V> g;
    int r, row, id;
    long long axis;
    if (!locate<L, V, true>(d, sec, cv, &g, &r, &row, &id, &axis)) return;
    switch (r) {
        case 0: same_fwd_body<L, V, 0>(d, g, row, id, x, down, up, same); break;
        case 1: same_fwd_body<L, V, 1>(d, g, row, id, x, down, up, same); break;
        case 2: same_fwd_body<L, V, 2>(d, g, row, id, x, down, up, same); break;
        default: same_fwd_body<L, V, 3>(d, g, row, id, x, down, up, same); break;
    }
}

// ------------------------------------------------------------------ backward, stage X: same-rank
// Updates g_up[2], g_up[3], g_down[2] in place (owner rows only), adds the diagonal and A0 terms to
// g_x, and the direct probability derivatives to g_probs.
template <int L, int V, int R>
__device__ __forceinline__ void same_bwd_body(const DeviceTables& d, const Group<L, V>& g, int row, int id, long long axis,
                                              const Feat& x, const Feat& down, const Feat& up, const Feat& g_same,
                                              const FeatMut& g_down, const FeatMut& g_up, const FeatMut& g_x,
                                              float* __restrict__ g_probs) {
    const float p = g.probs[d.off[R] + id];
    float gp = 0.f;            // d loss / d p_id collected by this group (lane-partial dot products)
    FV<V> gx;                     // addition to g_x[R][row]
    gx.zero();

    if constexpr (R == 0) {
        // same[0] = A0 X0, A0 symmetric:  g_x0[v] += sum_{v'} p_e g_same0[v']
        for_vertex_neighbours<L, V>(d, g, id, [&](const int (&rr)[4], const float (&pp)[4]) {
            FV<V> v[4];
            gather4<L, V>(g, g_same.p[0], rr, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) gx.fma(pp[q], v[q]);
        });
    }
    if constexpr (R == 1) {
        // owner of p_e for A0: d/dp_e = <g_same0[v], X0[v']> + <g_same0[v'], X0[v]>
        int fr[2];
        face_rows<L, V, 1>(d, g, id, fr);
        const bool both = fr[0] >= 0 && fr[1] >= 0;
        const FV<V> ga = g.load_if(both, g_same.p[0], fr[0]), gb = g.load_if(both, g_same.p[0], fr[1]);
        const FV<V> xa = g.load_if(both, x.p[0], fr[0]), xb = g.load_if(both, x.p[0], fr[1]);
        gp += ga.dot(xb) + gb.dot(xa);
    }
    if constexpr (R == 1 || R == 2) {
        // diagonal of same[R]: g_x[R] -= q g_same[R],  q = sum_{s > id} p_s^2
        if (d.cnt[R + 1] > 0) {
            float qsum = 0.f;
            for_cofaces<L, V, R>(d, g, id, [&](const int (&rr)[4], const float (&pp)[4]) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (rr[q] >= 0) qsum = fmaf(pp[q], pp[q], qsum);
            });
            const FV<V> gs = g.load_if(g.valid, g_same.p[R], row);
#pragma unroll
            for (int k = 0; k < 4 * V; ++k) gx.v[k] = fmaf(-qsum, gs.v[k], gx.v[k]);
        }
    }
    if constexpr (R == 2 || R == 3) {
        // owner of p_id for same[R-1] = I_R up[R] - q X_{R-1}:
        //   Rsum = sum_{f < id} g_same[R-1][f];  g_up[R][id] += p Rsum;
        //   d/dp = <Rsum, up[R][id]> - 2 p sum_f <g_same[R-1][f], X_{R-1}[f]>
        int fr[R + 1];
        face_rows<L, V, R>(d, g, id, fr);
        FV<V> gf[R + 1], xf[R + 1];
#pragma unroll
        for (int a = 0; a <= R; ++a) {
            gf[a] = g.load_if(fr[a] >= 0, g_same.p[R - 1], fr[a]);
            xf[a] = g.load_if(fr[a] >= 0, x.p[R - 1], fr[a]);
        }
        FV<V> rsum;
        rsum.zero();
        float diag = 0.f;
#pragma unroll
        for (int a = 0; a <= R; ++a) {
            rsum.add(gf[a]);
            diag += gf[a].dot(xf[a]);
        }
        const FV<V> u = g.load_if(g.valid, up.p[R], row);
        gp += rsum.dot(u) - 2.0f * p * diag;
#pragma unroll
        for (int k = 0; k < 4 * V; ++k) rsum.v[k] *= p;
        g.accumulate(g_up.p[R], row, rsum);
    }
    if constexpr (R == 2) {
        if (d.cnt[3] > 0) {
            // same[3] = I_3^T down[2] - ...:  g_down[2][t] += sum_{s > t} p_s g_same[3][s]
            FV<V> acc;
            acc.zero();
            for_cofaces<L, V, 2>(d, g, id, [&](const int (&rr)[4], const float (&pp)[4]) {
                FV<V> v[4];
                gather4<L, V>(g, g_same.p[3], rr, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) acc.fma(pp[q], v[q]);
            });
            g.accumulate(g_down.p[2], row, acc);
        }
    }
    if constexpr (R == 3) {
        // direct p dependence of same[3] = p sum_t down[2][t] - c p^2 X3, and its diagonal
        int fr[4];
        face_rows<L, V, 3>(d, g, id, fr);
        FV<V> v[4];
        gather4<L, V>(g, down.p[2], fr, v);
        FV<V> sum_d;
        sum_d.zero();
        int n_faces = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            sum_d.add(v[a]);
            n_faces += fr[a] >= 0;
        }
        const FV<V> gs = g.load_if(g.valid, g_same.p[3], row);
        const FV<V> own = g.load_if(g.valid, x.p[3], row);
        const float cf = static_cast<float>(n_faces);
        gp += gs.dot(sum_d) - 2.0f * cf * p * gs.dot(own);
#pragma unroll
        for (int k = 0; k < 4 * V; ++k) gx.v[k] = fmaf(-cf * p * p, gs.v[k], gx.v[k]);
    }

    g.accumulate(g_x.p[R], row, gx);
    if constexpr (R >= 1) {
        gp = g.group_sum(gp);
        if (g.lane == 0 && g.valid) g_probs[axis + d.off[R] + id] += gp;
    }
}

template <int L, int V>
__global__ void __launch_bounds__(kThreads, V == 1 ? 8 : 5) agg_same_bwd(const DeviceTables d, const Sections sec,
                                                         const topo_complex_view cv, const Feat x, const Feat down,
                                                         const Feat up, const Feat g_same, const FeatMut g_down,
                                                         const FeatMut g_up, const FeatMut g_x, float* __restrict__ g_probs) {
    Group<L, V> g;
    int r, row, id;
    long long axis;
    if (!locate<L, V, true>(d, sec, cv, &g, &r, &row, &id, &axis)) return;
    switch (r) {
        case 0: same_bwd_body<L, V, 0>(d, g, row, id, axis, x, down, up, g_same, g_down, g_up, g_x, g_probs); break;
        case 1: same_bwd_body<L, V, 1>(d, g, row, id, axis, x, down, up, g_same, g_down, g_up, g_x, g_probs); break;
        case 2: same_bwd_body<L, V, 2>(d, g, row, id, axis, x, down, up, g_same, g_down, g_up, g_x, g_probs); break;
        default: same_bwd_body<L, V, 3>(d, g, row, id, axis, x, down, up, g_same, g_down, g_up, g_x, g_probs); break;
    }
}

// ------------------------------------------------------------------ backward, stage Y: cross-rank
// With the TOTAL g_down / g_up:  down[R-1] = I_R X_R,  up[R] = I_R^T X_{R-1}.
template <int L, int V, int R>
__device__ __forceinline__ void cross_bwd_body(const DeviceTables& d, const Group<L, V>& g, int row, int id, long long axis,
                                               const Feat& x, const Feat& g_down, const Feat& g_up, const FeatMut& g_x,
                                               float* __restrict__ g_probs) {
    FV<V> gx;
    gx.zero();
    float gp = 0.f;
    if constexpr (R >= 1) {
        const float p = g.probs[d.off[R] + id];
        // as the coface s of down[R-1]:  Rsum = sum_f g_down[R-1][f];  g_x += p Rsum;  d/dp = <Rsum, X_R[s]>
        // as the target of up[R]:        d/dp = <g_up[R][s], sum_f X_{R-1}[f]>
        int fr[R + 1];
        face_rows<L, V, R>(d, g, id, fr);
        FV<V> gf[R + 1], xf[R + 1];
#pragma unroll
        for (int a = 0; a <= R; ++a) {
            gf[a] = g.load_if(fr[a] >= 0, g_down.p[R - 1], fr[a]);
            xf[a] = g.load_if(fr[a] >= 0, x.p[R - 1], fr[a]);
        }
        FV<V> rsum, qsum;
        rsum.zero();
        qsum.zero();
#pragma unroll
        for (int a = 0; a <= R; ++a) {
            rsum.add(gf[a]);
            qsum.add(xf[a]);
        }
        const FV<V> own = g.load_if(g.valid, x.p[R], row);
        const FV<V> gu = g.load_if(g.valid, g_up.p[R], row);
        gp = rsum.dot(own) + qsum.dot(gu);
#pragma unroll
        for (int k = 0; k < 4 * V; ++k) gx.v[k] = p * rsum.v[k];
    }
    if constexpr (R < 3) {
        if (d.cnt[R + 1] > 0) {
            // as a face of up[R+1]:  g_x[R][f] += sum_{s > f} p_s g_up[R+1][s]
            for_cofaces<L, V, R>(d, g, id, [&](const int (&rr)[4], const float (&pp)[4]) {
                FV<V> v[4];
                gather4<L, V>(g, g_up.p[R + 1], rr, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) gx.fma(pp[q], v[q]);
            });
        }
    }
    g.accumulate(g_x.p[R], row, gx);
    if constexpr (R >= 1) {
        gp = g.group_sum(gp);
        if (g.lane == 0 && g.valid) g_probs[axis + d.off[R] + id] += gp;
    }
}

template <int L, int V>
__global__ void __launch_bounds__(kThreads) agg_cross_bwd(const DeviceTables d, const Sections sec,
                                                          const topo_complex_view cv, const Feat x, const Feat g_down,
                                                          const Feat g_up, const FeatMut g_x, float* __restrict__ g_probs) {
    Group<L, V> g;
    int r, row, id;
    long long axis;
    if (!locate<L, V, true>(d, sec, cv, &g, &r, &row, &id, &axis)) return;
    switch (r) {
        case 0: cross_bwd_body<L, V, 0>(d, g, row, id, axis, x, g_down, g_up, g_x, g_probs); break;
        case 1: cross_bwd_body<L, V, 1>(d, g, row, id, axis, x, g_down, g_up, g_x, g_probs); break;
        case 2: cross_bwd_body<L, V, 2>(d, g, row, id, axis, x, g_down, g_up, g_x, g_probs); break;
        default: cross_bwd_body<L, V, 3>(d, g, row, id, axis, x, g_down, g_up, g_x, g_probs); break;
    }
}

// ------------------------------------------------------------------ backward, top rank: stages X and Y in one pass
// A tetrahedron's own row is all the two stages share across rows: stage Y needs the TOTAL g_down[2] of its faces
// (complete once stage X has run for the triangles) and the total g_up[3] of the row itself, which this kernel forms.
// So the order  stage X (ranks 0..2)  ->  this kernel (rank 3)  ->  stage Y (ranks 0..2, which gather the total
// g_up[3] written here)  reads x[3] once, reads and writes g_up[3] and g_x[3] once, gathers the faces' features once and
// never reads up[3] (it is p times the sum of the faces' features, which are in registers anyway).
template <int L, int V>
__global__ void __launch_bounds__(kThreads, 4) agg_top_bwd(const DeviceTables d, const Sections sec, const topo_complex_view cv,
                                                           const Feat x, const Feat down, const Feat g_same,
                                                           const Feat g_down, const FeatMut g_up, const FeatMut g_x,
                                                           float* __restrict__ g_probs) {
    Group<L, V> g;
    int r, row, id;
    long long axis;
    if (!locate<L, V, true>(d, sec, cv, &g, &r, &row, &id, &axis, sec.begin[3])) return;
    using F = FV<V>;
    const float p = g.probs[d.off[3] + id];
    int fr[4];
    face_rows<L, V, 3>(d, g, id, fr);
    int n_faces = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) n_faces += fr[a] >= 0;
    const float cf = static_cast<float>(n_faces);

    // stage X, as the owner of p for same[2] = I_3 up[3] - q X_2:  Rsum = sum_f g_same[2][f]
    F rsum, qsum;                       // qsum = sum_f X_2[f] (stage Y)
    rsum.zero();
    qsum.zero();
    float diag = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const F gf = g.load_if(fr[a] >= 0, g_same.p[2], fr[a]);
        const F xf = g.load_if(fr[a] >= 0, x.p[2], fr[a]);
        rsum.add(gf);
        qsum.add(xf);
        diag += gf.dot(xf);
    }
    // up[3][row] = p sum_f X_2[f] is what the forward wrote (same summation order): recomputed, not re-read
    float gp = p * rsum.dot(qsum) - 2.0f * p * diag;
    F gu;                               // total g_up[3][row]
    gu.zero();
    if (g.valid) gu = g.load_rw(g_up.p[3], row);
#pragma unroll
    for (int k = 0; k < 4 * V; ++k) gu.v[k] = fmaf(p, rsum.v[k], gu.v[k]);
    g.store(g_up.p[3], row, gu);        // stage Y of the triangles gathers it

    // stage X, the direct p dependence of same[3] = p sum_t down[2][t] - c p^2 X3, and its diagonal
    F sum_d;
    sum_d.zero();
#pragma unroll
    for (int a = 0; a < 4; ++a) sum_d.add(g.load_if(fr[a] >= 0, down.p[2], fr[a]));
    const F gs = g.load_if(g.valid, g_same.p[3], row);
    const F own = g.load_if(g.valid, x.p[3], row);
    gp += gs.dot(sum_d) - 2.0f * cf * p * gs.dot(own);

    // stage Y, as the coface of down[2] and the target of up[3]
    F rsum2;
    rsum2.zero();
#pragma unroll
    for (int a = 0; a < 4; ++a) rsum2.add(g.load_if(fr[a] >= 0, g_down.p[2], fr[a]));
    gp += rsum2.dot(own) + qsum.dot(gu);

    F gx;
#pragma unroll
    for (int k = 0; k < 4 * V; ++k) gx.v[k] = fmaf(-cf * p * p, gs.v[k], p * rsum2.v[k]);
    g.accumulate(g_x.p[3], row, gx);
    gp = g.group_sum(gp);
    if (g.lane == 0 && g.valid) g_probs[axis + d.off[3] + id] += gp;
}

// ------------------------------------------------------------------ generic CSR
constexpr int kWarpsPerBlock = 8;

template <int VEC>
struct Acc {
    float v[VEC];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = 0.f;
    }
    __device__ __forceinline__ void fma(float w, const Vec<VEC>& x) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = fmaf(w, x.v[k], v[k]);
    }
};

template <int VEC>
__device__ __forceinline__ Vec<VEC> load_row(const float* base, long long row, int lane) {
    Vec<VEC> x;
    x.load(base + (row * 32 + lane) * VEC);
    return x;
}

template <int VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) spmm_csr_kernel(long long rows, const int* __restrict__ row_ptr,
                                                                       const int* __restrict__ col_idx,
                                                                       const float* __restrict__ vals,
                                                                       const float* __restrict__ x,
                                                                       float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    for (long long row = blockIdx.x * static_cast<long long>(kWarpsPerBlock) + (threadIdx.x >> 5); row < rows;
         row += static_cast<long long>(gridDim.x) * kWarpsPerBlock) {
        const int e0 = row_ptr[row], e1 = row_ptr[row + 1];
        Acc<VEC> acc;
        acc.zero();
        for (int eb = e0; eb < e1; eb += 32) {
            const int e = eb + lane;
            const int col = e < e1 ? __ldg(col_idx + e) : 0;
            const float v = e < e1 ? __ldg(vals + e) : 0.f;
            const int m = min(32, e1 - eb);
#pragma unroll 4
            for (int jj = 0; jj < m; ++jj) {
                const int cc = __shfl_sync(0xffffffffu, col, jj);
                const float vv = __shfl_sync(0xffffffffu, v, jj);
                acc.fma(vv, load_row<VEC>(x, cc, lane));
            }
        }
        Vec<VEC> out;
#pragma unroll
        for (int k = 0; k < VEC; ++k) out.v[k] = acc.v[k];
        out.store(y + (row * 32 + lane) * VEC);
    }
}

template <int VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) sddmm_csr_kernel(long long rows, const int* __restrict__ row_ptr,
                                                                        const int* __restrict__ col_idx,
                                                                        const float* __restrict__ g_y,
                                                                        const float* __restrict__ x,
                                                                        float* __restrict__ g_vals) {
    const int lane = threadIdx.x & 31;
    for (long long row = blockIdx.x * static_cast<long long>(kWarpsPerBlock) + (threadIdx.x >> 5); row < rows;
         row += static_cast<long long>(gridDim.x) * kWarpsPerBlock) {
        const int e0 = row_ptr[row], e1 = row_ptr[row + 1];
        const Vec<VEC> g = load_row<VEC>(g_y, row, lane);
        for (int e = e0; e < e1; ++e) {
            const Vec<VEC> xv = load_row<VEC>(x, __ldg(col_idx + e), lane);
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < VEC; ++k) s = fmaf(g.v[k], xv.v[k], s);
            s = warp_sum(s);
            if (lane == 0) g_vals[e] = s;
        }
    }
}

Sections make_sections(const DeviceTables& d, int64_t batch, int groups_per_block) {
    Sections s;
    int acc = 0;
    for (int r = 0; r < 4; ++r) {
        s.begin[r] = acc;
        s.chunks[r] = std::max(1, (d.cnt[r] + groups_per_block - 1) / groups_per_block);
        acc += (d.cnt[r] ? s.chunks[r] : 0) * static_cast<int>(batch);
    }
    s.begin[4] = acc;   // empty (top) ranks own no blocks; begin[] stays monotone
    return s;
}

int check_view(const topo_tables* t, const topo_complex_view* cv, int channels) {
    TOPO_REQUIRE(t && cv, "null argument");
    TOPO_REQUIRE(t->device >= 0, "tables were built host-only");
    TOPO_REQUIRE_TABLES_DEVICE(t);
    TOPO_REQUIRE(cv->probs && cv->pos && cv->act_idx && cv->counts && cv->row_off, "null pointer in complex view");
    TOPO_REQUIRE((reinterpret_cast<uintptr_t>(cv->counts) & 15) == 0, "complex view: counts must be 16-byte aligned");
    TOPO_REQUIRE(cv->batch >= 0 && cv->batch <= 65535, "batch out of range");
    if (channels != 32 && channels != 64 && channels != 128) {
        set_error("channels must be 32, 64 or 128");
        return TOPO_ERR_UNSUPPORTED;
    }
    return TOPO_OK;
}

}  // namespace
}  // namespace topo

using namespace topo;

// lanes per row and float4s per lane: C = 64 runs with 8 lanes x 2 float4s (four rows per warp: the index arithmetic of
// a row is paid once per warp instruction, so twice as many rows share it)
#define DISPATCH_LANES(channels, CALL)                                  \
    switch (channels) {                                                 \
        case 32: { constexpr int L = 8, V = 1; CALL; } break;           \
        case 64: { constexpr int L = 8, V = 2; CALL; } break;           \
        default: { constexpr int L = 32, V = 1; CALL; } break;          \
    }
static int lanes_per_row(int channels) { return channels == 32 ? 8 : (channels == 64 ? 8 : 32); }

#define DISPATCH_VEC(channels, CALL)                      \
    switch (channels) {                                   \
        case 32: { constexpr int VEC = 1; CALL; } break;  \
        case 64: { constexpr int VEC = 2; CALL; } break;  \
        default: { constexpr int VEC = 4; CALL; } break;  \
    }

extern "C" int topo_sccn_aggregate_fwd(const topo_tables* t, const topo_complex_view* cv, int channels,
                                       const float* const x[4], float* const down[4], float* const up[4],
                                       float* const same[4], topo_stream_t stream) {
    if (int rc = check_view(t, cv, channels)) return rc;
    TOPO_REQUIRE(x && down && up && same, "null argument");
    if (cv->batch == 0) return TOPO_OK;
    const DeviceTables& d = t->d;
    Feat fx, fdown, fup;
    FeatMut mdown, mup, msame;
    for (int r = 0; r < 4; ++r) {
        fx.p[r] = x[r]; fdown.p[r] = down[r]; fup.p[r] = up[r];
        mdown.p[r] = down[r]; mup.p[r] = up[r]; msame.p[r] = same[r];
        if (d.cnt[r]) {
            TOPO_REQUIRE(x[r] && same[r], "missing feature / same buffer for a populated rank");
            TOPO_REQUIRE(r == 0 || up[r], "missing up buffer");
            TOPO_REQUIRE(r == 3 || d.cnt[r + 1] == 0 || down[r], "missing down buffer");
        }
    }
    const Sections sec = make_sections(d, cv->batch, kThreads / lanes_per_row(channels));
    if (sec.begin[4] == 0) return TOPO_OK;
    cudaStream_t s = as_stream(stream);
    DISPATCH_LANES(channels, (agg_cross_fwd<L, V><<<sec.begin[4], kThreads, 0, s>>>(d, sec, *cv, fx, mdown, mup)));
    DISPATCH_LANES(channels, (agg_same_fwd<L, V><<<sec.begin[4], kThreads, 0, s>>>(d, sec, *cv, fx, fdown, fup, msame)));
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_sccn_aggregate_bwd(const topo_tables* t, const topo_complex_view* cv, int channels,
                                       const float* const x[4], const float* const down[4],
                                       const float* const up[4], float* const g_down[4], float* const g_up[4],
                                       const float* const g_same[4], float* const g_x[4], float* g_probs,
                                       topo_stream_t stream) {
    if (int rc = check_view(t, cv, channels)) return rc;
    TOPO_REQUIRE(x && down && up && g_down && g_up && g_same && g_x && g_probs, "null argument");
    if (cv->batch == 0) return TOPO_OK;
    const DeviceTables& d = t->d;
    Feat fx, fdown, fup, fgsame, fgdown, fgup;
    FeatMut mgdown, mgup, mgx;
    for (int r = 0; r < 4; ++r) {
        fx.p[r] = x[r]; fdown.p[r] = down[r]; fup.p[r] = up[r]; fgsame.p[r] = g_same[r];
        fgdown.p[r] = g_down[r]; fgup.p[r] = g_up[r];
        mgdown.p[r] = g_down[r]; mgup.p[r] = g_up[r]; mgx.p[r] = g_x[r];
        if (d.cnt[r]) TOPO_REQUIRE(x[r] && g_same[r] && g_x[r], "missing buffer for a populated rank");
    }
    const Sections sec = make_sections(d, cv->batch, kThreads / lanes_per_row(channels));
    if (sec.begin[4] == 0) return TOPO_OK;
    cudaStream_t s = as_stream(stream);
    // ranks 0..2: stage X, then stage Y; the top rank runs both stages in one pass in between (agg_top_bwd)
    const int top = d.cnt[3] > 0 ? sec.begin[4] - sec.begin[3] : 0;
    const int low = sec.begin[4] - top;
    if (low > 0)
        DISPATCH_LANES(channels, (agg_same_bwd<L, V><<<low, kThreads, 0, s>>>(
                                     d, sec, *cv, fx, fdown, fup, fgsame, mgdown, mgup, mgx, g_probs)));
    if (top > 0)
        DISPATCH_LANES(channels, (agg_top_bwd<L, V><<<top, kThreads, 0, s>>>(
                                     d, sec, *cv, fx, fdown, fgsame, fgdown, mgup, mgx, g_probs)));
    if (low > 0)
        DISPATCH_LANES(channels, (agg_cross_bwd<L, V><<<low, kThreads, 0, s>>>(
                                     d, sec, *cv, fx, fgdown, fgup, mgx, g_probs)));
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_spmm_csr(int64_t rows, const int32_t* row_ptr, const int32_t* col_idx, const float* vals,
                             const float* x, int channels, float* y, topo_stream_t stream) {
    TOPO_REQUIRE(rows >= 0 && row_ptr && y, "bad argument");
    if (channels != 32 && channels != 64 && channels != 128) {
        set_error("channels must be 32, 64 or 128");
        return TOPO_ERR_UNSUPPORTED;
    }
    if (rows == 0) return TOPO_OK;
    const int grid = static_cast<int>(std::min<int64_t>((rows + kWarpsPerBlock - 1) / kWarpsPerBlock, sm_count() * 16));
    DISPATCH_VEC(channels, (spmm_csr_kernel<VEC><<<grid, kWarpsPerBlock * 32, 0, as_stream(stream)>>>(
                               rows, row_ptr, col_idx, vals, x, y)));
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_sddmm_csr(int64_t rows, const int32_t* row_ptr, const int32_t* col_idx, const float* g_y,
                              const float* x, int channels, float* g_vals, topo_stream_t stream) {
    TOPO_REQUIRE(rows >= 0 && row_ptr && g_y && x, "bad argument");
    if (channels != 32 && channels != 64 && channels != 128) {
        set_error("channels must be 32, 64 or 128");
        return TOPO_ERR_UNSUPPORTED;
    }
    if (rows == 0) return TOPO_OK;
    const int grid = static_cast<int>(std::min<int64_t>((rows + kWarpsPerBlock - 1) / kWarpsPerBlock, sm_count() * 16));
    DISPATCH_VEC(channels, (sddmm_csr_kernel<VEC><<<grid, kWarpsPerBlock * 32, 0, as_stream(stream)>>>(
                               rows, row_ptr, col_idx, g_y, x, g_vals)));
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
