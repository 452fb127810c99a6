// S1 (a): neighbourhood aggregation for SCCN, forward and backward.
//
// Structured path (whole batch, matrix-free).  The reference applies ten sparse operators per layer
// (custom_sccn.py:78-81, 95-98, 113-116) that build_sparse_matrices materialised
// (complex_builder.py:52-70).  Here nothing is materialised: one warp owns one active simplex,
// walks its static face / coface lists, weights the gathered rows with the rectified probabilities
// and writes one C-wide row.  Every adjacency factors through the incidences
// (A1 = I2 I2^T, A2 = I3 I3^T, A3 = I3^T I3, complex_builder.py:62-64), so the same-rank aggregates
// are incidence gathers of the cross-rank aggregates minus a diagonal term:
//     down[r] = I_{r+1} X_{r+1}      up[r] = I_r^T X_{r-1}
//     same[1] = I_2 up[2] - q1 . X1          q1[e] = sum_{t > e} p_t^2
//     same[2] = I_3 up[3] - q2 . X2          q2[t] = sum_{s > t} p_s^2
//     same[3] = I_3^T down[2] - c p^2 . X3   c[s]  = number of active faces
// 88,920 row gathers per sample-layer for the full 20-vertex complex instead of 421,800.
// All rows are single-owner, so the backward needs no atomics and is deterministic.
//
// Generic path: CSR SpMM / SDDMM for caller-supplied operators.
#include "common.cuh"

namespace topo {
namespace {

constexpr int kWarpsPerBlock = 8;

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
    __device__ __forceinline__ void add(const Vec<VEC>& x) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] += x.v[k];
    }
    __device__ __forceinline__ float dot(const Vec<VEC>& x) const {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < VEC; ++k) s = fmaf(v[k], x.v[k], s);
        return s;
    }
};

template <int VEC>
__device__ __forceinline__ Vec<VEC> load_row(const float* base, int row, int lane) {
    Vec<VEC> x;
    x.load(base + (static_cast<long long>(row) * 32 + lane) * VEC);
    return x;
}
// plain (non read-only-path) load for arrays another launch phase of the same kernel may write
template <int VEC>
__device__ __forceinline__ Vec<VEC> load_row_rw(const float* base, int row, int lane) {
    Vec<VEC> x;
    const float* p = base + (static_cast<long long>(row) * 32 + lane) * VEC;
#pragma unroll
    for (int k = 0; k < VEC; ++k) x.v[k] = p[k];
    return x;
}
template <int VEC>
__device__ __forceinline__ void store_row(float* base, int row, int lane, const float (&v)[VEC]) {
    Vec<VEC> x;
#pragma unroll
    for (int k = 0; k < VEC; ++k) x.v[k] = v[k];
    x.store(base + (static_cast<long long>(row) * 32 + lane) * VEC);
}
template <int VEC>
__device__ __forceinline__ void add_row(float* base, int row, int lane, const float (&v)[VEC]) {
    float* p = base + (static_cast<long long>(row) * 32 + lane) * VEC;
    Vec<VEC> x;
#pragma unroll
    for (int k = 0; k < VEC; ++k) x.v[k] = p[k] + v[k];
    x.store(p);
}

// One warp's view of its sample.
struct WarpCtx {
    const DeviceTables* d;
    const float* probs;   // this sample's simplex axis
    const int* pos;
    int ro[4];            // compact row offset of this sample per rank
    int lane;

    __device__ __forceinline__ int row_of(int r, int id) const {
        const int p = pos[d->off[r] + id];
        return p < 0 ? -1 : ro[r] + p;
    }
    __device__ __forceinline__ float prob(int r, int id) const { return probs[d->off[r] + id]; }
};

// f(row, p_s) for every active coface s (p_s != 0) of simplex `id` of rank r, ascending.
template <typename F>
__device__ __forceinline__ void for_cofaces(const WarpCtx& c, int r, int id, F&& f) {
    const int w = c.d->ncof[r];
    const int* cof = c.d->cofaces[r] + static_cast<long long>(id) * w;
    for (int jb = 0; jb < w; jb += 32) {
        const int j = jb + c.lane;
        int row = -1;
        float ps = 0.f;
        if (j < w) {
            const int s = __ldg(cof + j);
            ps = c.prob(r + 1, s);
            row = (ps != 0.0f) ? c.row_of(r + 1, s) : -1;
        }
        const int m = min(32, w - jb);
#pragma unroll 4
        for (int jj = 0; jj < m; ++jj) {
            const int rr = __shfl_sync(0xffffffffu, row, jj);
            const float pp = __shfl_sync(0xffffffffu, ps, jj);
            if (rr >= 0) f(rr, pp);
        }
    }
}

// f(row) for every active face of simplex `id` of rank r >= 1, ascending.  Returns the count.
template <typename F>
__device__ __forceinline__ int for_faces(const WarpCtx& c, int r, int id, F&& f) {
    const int* fc = c.d->faces[r] + static_cast<long long>(id) * (r + 1);
    int n = 0;
    for (int a = 0; a <= r; ++a) {
        const int row = c.row_of(r - 1, __ldg(fc + a));
        if (row >= 0) { f(row); ++n; }
    }
    return n;
}

// Decode blockIdx.x -> (rank, sample, first local row); returns false when this warp has no row.
__device__ __forceinline__ bool locate(const DeviceTables& d, const Sections& sec, const topo_complex_view& cv,
                                       WarpCtx* c, int* rank, int* local, int* id) {
    const int blk = blockIdx.x;
    const int r = (blk >= sec.begin[1]) + (blk >= sec.begin[2]) + (blk >= sec.begin[3]);
    const int rel = blk - sec.begin[r];
    const int b = rel / sec.chunks[r];
    const int i = (rel % sec.chunks[r]) * kWarpsPerBlock + (threadIdx.x >> 5);
    if (i >= cv.counts[b * 4 + r]) return false;
    const long long axis = static_cast<long long>(b) * d.off[4];
    c->d = &d;
    c->probs = cv.probs + axis;
    c->pos = cv.pos + axis;
    c->lane = threadIdx.x & 31;
    const int B1 = static_cast<int>(cv.batch) + 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) c->ro[q] = cv.row_off[q * B1 + b];
    *rank = r;
    *local = i;
    *id = cv.act_idx[axis + d.off[r] + i];
    return true;
}

// ------------------------------------------------------------------ forward, phase 1: cross-rank
template <int VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) agg_cross_fwd(DeviceTables d, Sections sec,
                                                                     topo_complex_view cv, Feat x, FeatMut down,
                                                                     FeatMut up) {
    WarpCtx c;
    int r, i, id;
    if (!locate(d, sec, cv, &c, &r, &i, &id)) return;
    const int row = c.ro[r] + i;
    if (r < 3 && d.cnt[r + 1] > 0) {   // down[r] = sum over cofaces p_s X_{r+1}[s]
        Acc<VEC> acc;
        acc.zero();
        for_cofaces(c, r, id, [&](int rr, float ps) { acc.fma(ps, load_row<VEC>(x.p[r + 1], rr, c.lane)); });
        store_row<VEC>(down.p[r], row, c.lane, acc.v);
    }
    if (r > 0) {                       // up[r] = p_id * sum over active faces X_{r-1}[f]
        Acc<VEC> acc;
        acc.zero();
        for_faces(c, r, id, [&](int rr) { acc.add(load_row<VEC>(x.p[r - 1], rr, c.lane)); });
        const float p = c.prob(r, id);
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc.v[k] *= p;
        store_row<VEC>(up.p[r], row, c.lane, acc.v);
    }
}

// ------------------------------------------------------------------ forward, phase 2: same-rank
template <int VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) agg_same_fwd(DeviceTables d, Sections sec,
                                                                    topo_complex_view cv, Feat x, Feat down, Feat up,
                                                                    FeatMut same) {
    WarpCtx c;
    int r, i, id;
    if (!locate(d, sec, cv, &c, &r, &i, &id)) return;
    const int row = c.ro[r] + i;
    Acc<VEC> acc;
    acc.zero();
    if (r == 0) {
        // A0[v,v'] = p_e: walk the vertex's edges, gather the other endpoint
        const int w = d.ncof[0];
        const int* cof = d.cofaces[0] + static_cast<long long>(id) * w;
        for (int jb = 0; jb < w; jb += 32) {
            const int j = jb + c.lane;
            int orow = -1;
            float pe = 0.f;
            if (j < w) {
                const int e = __ldg(cof + j);
                pe = c.prob(1, e);
                const int2 ends = __ldg(reinterpret_cast<const int2*>(d.faces[1]) + e);
                orow = (pe != 0.0f) ? c.row_of(0, ends.x == id ? ends.y : ends.x) : -1;
            }
            const int m = min(32, w - jb);
#pragma unroll 4
            for (int jj = 0; jj < m; ++jj) {
                const int rr = __shfl_sync(0xffffffffu, orow, jj);
                const float pp = __shfl_sync(0xffffffffu, pe, jj);
                if (rr >= 0) acc.fma(pp, load_row<VEC>(x.p[0], rr, c.lane));
            }
        }
    } else if (r < 3) {
        // same[r] = sum_{s > id} p_s up[r+1][s]  -  (sum p_s^2) X_r[id]
        float q = 0.f;
        if (d.cnt[r + 1] > 0)
            for_cofaces(c, r, id, [&](int rr, float ps) {
                acc.fma(ps, load_row<VEC>(up.p[r + 1], rr, c.lane));
                q = fmaf(ps, ps, q);
            });
        const Vec<VEC> own = load_row<VEC>(x.p[r], row, c.lane);
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc.v[k] = fmaf(-q, own.v[k], acc.v[k]);
    } else {
        // same[3] = p (sum_{t < id, active} down[2][t] - c p X_3[id])
        const int n_faces = for_faces(c, 3, id, [&](int rr) { acc.add(load_row<VEC>(down.p[2], rr, c.lane)); });
        const float p = c.prob(3, id);
        const Vec<VEC> own = load_row<VEC>(x.p[3], row, c.lane);
        const float cp = static_cast<float>(n_faces) * p;
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc.v[k] = p * fmaf(-cp, own.v[k], acc.v[k]);
    }
    store_row<VEC>(same.p[r], row, c.lane, acc.v);
}

// ------------------------------------------------------------------ backward, stage X: same-rank
// Updates g_up[2], g_up[3], g_down[2] in place (owner rows only), adds the diagonal and A0 terms to
// g_x, and the direct probability derivatives to g_probs.
template <int VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) agg_same_bwd(DeviceTables d, Sections sec,
                                                                    topo_complex_view cv, Feat x, Feat down, Feat up,
                                                                    Feat g_same, FeatMut g_down, FeatMut g_up,
                                                                    FeatMut g_x, float* __restrict__ g_probs) {
    WarpCtx c;
    int r, i, id;
    if (!locate(d, sec, cv, &c, &r, &i, &id)) return;
    const int row = c.ro[r] + i;
    const long long axis = c.probs - cv.probs;
    const float p = c.prob(r, id);
    float gp = 0.f;            // d loss / d p_id collected by this warp (lane-partial dot products)
    Acc<VEC> gx;               // addition to g_x[r][row]
    gx.zero();

    if (r == 0) {
        // same[0] = A0 X0, A0 symmetric:  g_x0[v] += sum_{v'} p_e g_same0[v']
        const int w = d.ncof[0];
        const int* cof = d.cofaces[0] + static_cast<long long>(id) * w;
        for (int jb = 0; jb < w; jb += 32) {
            const int j = jb + c.lane;
            int orow = -1;
            float pe = 0.f;
            if (j < w) {
                const int e = __ldg(cof + j);
                pe = c.prob(1, e);
                const int2 ends = __ldg(reinterpret_cast<const int2*>(d.faces[1]) + e);
                orow = (pe != 0.0f) ? c.row_of(0, ends.x == id ? ends.y : ends.x) : -1;
            }
            const int m = min(32, w - jb);
#pragma unroll 4
            for (int jj = 0; jj < m; ++jj) {
                const int rr = __shfl_sync(0xffffffffu, orow, jj);
                const float pp = __shfl_sync(0xffffffffu, pe, jj);
                if (rr >= 0) gx.fma(pp, load_row<VEC>(g_same.p[0], rr, c.lane));
            }
        }
    }
    if (r == 1) {
        // owner of p_e for A0: d/dp_e = <g_same0[v], X0[v']> + <g_same0[v'], X0[v]>
        const int2 ends = __ldg(reinterpret_cast<const int2*>(d.faces[1]) + id);
        const int ra = c.row_of(0, ends.x), rb = c.row_of(0, ends.y);
        if (ra >= 0 && rb >= 0) {
            const Vec<VEC> ga = load_row<VEC>(g_same.p[0], ra, c.lane), gb = load_row<VEC>(g_same.p[0], rb, c.lane);
            const Vec<VEC> xa = load_row<VEC>(x.p[0], ra, c.lane), xb = load_row<VEC>(x.p[0], rb, c.lane);
#pragma unroll
            for (int k = 0; k < VEC; ++k) gp += ga.v[k] * xb.v[k] + gb.v[k] * xa.v[k];
        }
    }
    if (r == 1 || r == 2) {
        // diagonal of same[r]: g_x[r] -= q g_same[r],  q = sum_{s > id} p_s^2
        if (d.cnt[r + 1] > 0) {
            float q = 0.f;
            for_cofaces(c, r, id, [&](int, float ps) { q = fmaf(ps, ps, q); });
            const Vec<VEC> g = load_row<VEC>(g_same.p[r], row, c.lane);
#pragma unroll
            for (int k = 0; k < VEC; ++k) gx.v[k] = fmaf(-q, g.v[k], gx.v[k]);
        }
    }
    if (r == 2 || r == 3) {
        // owner of p_id for same[r-1] = I_r up[r] - q X_{r-1}:
        //   R = sum_{f < id} g_same[r-1][f];  g_up[r][id] += p R;
        //   d/dp = <R, up[r][id]> - 2 p sum_f <g_same[r-1][f], X_{r-1}[f]>
        Acc<VEC> R;
        R.zero();
        float diag = 0.f;
        for_faces(c, r, id, [&](int rr) {
            const Vec<VEC> g = load_row<VEC>(g_same.p[r - 1], rr, c.lane);
            const Vec<VEC> xf = load_row<VEC>(x.p[r - 1], rr, c.lane);
            R.add(g);
#pragma unroll
            for (int k = 0; k < VEC; ++k) diag = fmaf(g.v[k], xf.v[k], diag);
        });
        const Vec<VEC> u = load_row<VEC>(up.p[r], row, c.lane);
        gp += R.dot(u) - 2.0f * p * diag;
        float add[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) add[k] = p * R.v[k];
        add_row<VEC>(g_up.p[r], row, c.lane, add);
    }
    if (r == 2 && d.cnt[3] > 0) {
        // same[3] = I_3^T down[2] - ...:  g_down[2][t] += sum_{s > t} p_s g_same[3][s]
        Acc<VEC> acc;
        acc.zero();
        for_cofaces(c, 2, id, [&](int rr, float ps) { acc.fma(ps, load_row<VEC>(g_same.p[3], rr, c.lane)); });
        add_row<VEC>(g_down.p[2], row, c.lane, acc.v);
    }
    if (r == 3) {
        // direct p dependence of same[3] = p sum_t down[2][t] - c p^2 X3, and its diagonal
        Acc<VEC> sum_d;
        sum_d.zero();
        const int n_faces = for_faces(c, 3, id, [&](int rr) { sum_d.add(load_row<VEC>(down.p[2], rr, c.lane)); });
        const Vec<VEC> g = load_row<VEC>(g_same.p[3], row, c.lane);
        const Vec<VEC> own = load_row<VEC>(x.p[3], row, c.lane);
        const float cf = static_cast<float>(n_faces);
        float d1 = 0.f, d2 = 0.f;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            d1 = fmaf(g.v[k], sum_d.v[k], d1);
            d2 = fmaf(g.v[k], own.v[k], d2);
            gx.v[k] = fmaf(-cf * p * p, g.v[k], gx.v[k]);
        }
        gp += d1 - 2.0f * cf * p * d2;
    }

    add_row<VEC>(g_x.p[r], row, c.lane, gx.v);
    if (r >= 1) {
        gp = warp_sum(gp);
        if (c.lane == 0) g_probs[axis + d.off[r] + id] += gp;
    }
}

// ------------------------------------------------------------------ backward, stage Y: cross-rank
// With the TOTAL g_down / g_up:  down[r-1] = I_r X_r,  up[r] = I_r^T X_{r-1}.
template <int VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) agg_cross_bwd(DeviceTables d, Sections sec,
                                                                     topo_complex_view cv, Feat x, Feat g_down,
                                                                     Feat g_up, FeatMut g_x,
                                                                     float* __restrict__ g_probs) {
    WarpCtx c;
    int r, i, id;
    if (!locate(d, sec, cv, &c, &r, &i, &id)) return;
    const int row = c.ro[r] + i;
    const long long axis = c.probs - cv.probs;
    Acc<VEC> gx;
    gx.zero();
    float gp = 0.f;

    if (r >= 1) {
        const float p = c.prob(r, id);
        // as the coface s of down[r-1]:  R = sum_f g_down[r-1][f];  g_x += p R;  d/dp = <R, X_r[s]>
        // as the target of up[r]:        d/dp = <g_up[r][s], sum_f X_{r-1}[f]>
        Acc<VEC> R, Q;
        R.zero();
        Q.zero();
        for_faces(c, r, id, [&](int rr) {
            R.add(load_row<VEC>(g_down.p[r - 1], rr, c.lane));
            Q.add(load_row<VEC>(x.p[r - 1], rr, c.lane));
        });
        const Vec<VEC> own = load_row<VEC>(x.p[r], row, c.lane);
        const Vec<VEC> gu = load_row<VEC>(g_up.p[r], row, c.lane);
        gp = R.dot(own) + Q.dot(gu);
#pragma unroll
        for (int k = 0; k < VEC; ++k) gx.v[k] = p * R.v[k];
    }
    if (r < 3 && d.cnt[r + 1] > 0) {
        // as a face of up[r+1]:  g_x[r][f] += sum_{s > f} p_s g_up[r+1][s]
        for_cofaces(c, r, id, [&](int rr, float ps) { gx.fma(ps, load_row<VEC>(g_up.p[r + 1], rr, c.lane)); });
    }
    add_row<VEC>(g_x.p[r], row, c.lane, gx.v);
    if (r >= 1) {
        gp = warp_sum(gp);
        if (c.lane == 0) g_probs[axis + d.off[r] + id] += gp;
    }
}

// ------------------------------------------------------------------ generic CSR
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
        store_row<VEC>(y, static_cast<int>(row), lane, acc.v);
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
        const Vec<VEC> g = load_row<VEC>(g_y, static_cast<int>(row), lane);
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

Sections make_sections(const DeviceTables& d, int64_t batch) {
    Sections s;
    int acc = 0;
    for (int r = 0; r < 4; ++r) {
        s.begin[r] = acc;
        s.chunks[r] = (d.cnt[r] + kWarpsPerBlock - 1) / kWarpsPerBlock;
        if (s.chunks[r] == 0) s.chunks[r] = 1;
        acc += (d.cnt[r] ? s.chunks[r] : 0) * static_cast<int>(batch);
    }
    s.begin[4] = acc;
    // empty ranks own no blocks: make their section empty but keep begin[] monotone
    return s;
}

int check_view(const topo_tables* t, const topo_complex_view* cv, int channels) {
    TOPO_REQUIRE(t && cv, "null argument");
    TOPO_REQUIRE(cv->probs && cv->pos && cv->act_idx && cv->counts && cv->row_off, "null pointer in complex view");
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

#define DISPATCH_VEC(channels, CALL)      \
    switch (channels) {                   \
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
    const Sections sec = make_sections(d, cv->batch);
    if (sec.begin[4] == 0) return TOPO_OK;
    cudaStream_t s = as_stream(stream);
    DISPATCH_VEC(channels, (agg_cross_fwd<VEC><<<sec.begin[4], kWarpsPerBlock * 32, 0, s>>>(d, sec, *cv, fx, mdown, mup)));
    DISPATCH_VEC(channels, (agg_same_fwd<VEC><<<sec.begin[4], kWarpsPerBlock * 32, 0, s>>>(d, sec, *cv, fx, fdown, fup, msame)));
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
    const Sections sec = make_sections(d, cv->batch);
    if (sec.begin[4] == 0) return TOPO_OK;
    cudaStream_t s = as_stream(stream);
    DISPATCH_VEC(channels, (agg_same_bwd<VEC><<<sec.begin[4], kWarpsPerBlock * 32, 0, s>>>(
                               d, sec, *cv, fx, fdown, fup, fgsame, mgdown, mgup, mgx, g_probs)));
    DISPATCH_VEC(channels, (agg_cross_bwd<VEC><<<sec.begin[4], kWarpsPerBlock * 32, 0, s>>>(
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
