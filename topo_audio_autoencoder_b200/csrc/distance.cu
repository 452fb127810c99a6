// D1-D3: tiled pairwise multi-scale spectral distance.
// Replaces the pair loop of compute_distances (precompute_distances.py:89-115), which recomputes
// both multi-scale STFTs for every pair and scatters results in a Python loop.  Here the N
// spectrograms are computed once (front half, host side), laid out as one padded row per clip, and a
// 64 x 64 tile of pairs is reduced per CTA:
//     d(i,j) = sum_s [ mean_s((x-y)^2) / (mean_s(x^2) + 1e-7) + mean_s |log(x+eps) - log(y+eps)| ]
// with x the LOWER-index clip (precompute_distances.py:89, 106-110), mirrored (:114-115).
// FP32-ALU bound (4 instructions per pair-element; the L1-of-logs term is not a GEMM); operands are
// staged through shared memory, k-major, so one LDS.128 feeds four pairs.  Two-level summation
// (16-element stage sums folded into the per-scale accumulator) keeps fp32 error ~1e-6 relative.
#include <algorithm>

#include "common.cuh"

namespace topo {
namespace {

constexpr int KC = 16;           // spectrogram bins per stage; every scale segment is padded to it
constexpr int TI = 64;           // rows of the pair tile
constexpr int TJ = 64;           // columns of the pair tile
constexpr int LDI = TI + 4;      // padded strides of the k-major stages
constexpr int LDJ = TJ + 4;
constexpr int kMaxScales = 8;

struct Segments {
    long long len[kMaxScales];      // true length of each scale segment
    long long pad_off[kMaxScales];  // offset in the padded row
    long long pad_len[kMaxScales];  // padded length (multiple of KC)
    long long src_off[kMaxScales];  // offset in the unpadded row
    int n;
    long long dp;                   // padded row length
};

Segments make_segments(const int64_t* seg_len, int n_scales) {
    Segments s;
    s.n = n_scales;
    long long po = 0, so = 0;
    for (int i = 0; i < n_scales; ++i) {
        s.len[i] = seg_len[i];
        s.pad_len[i] = (seg_len[i] + KC - 1) / KC * KC;
        s.pad_off[i] = po;
        s.src_off[i] = so;
        po += s.pad_len[i];
        so += seg_len[i];
    }
    s.dp = po;
    return s;
}

// one CTA per (clip, scale): padded copy, log(x + eps), mean of squares
__global__ void __launch_bounds__(256) distance_prepare_kernel(const float* __restrict__ spec, long long d,
                                                               Segments seg, float log_eps,
                                                               float* __restrict__ spec_p,
                                                               float* __restrict__ logspec_p,
                                                               float* __restrict__ sq_mean) {
    const long long clip = blockIdx.x;
    const int s = blockIdx.y;
    const float* src = spec + clip * d + seg.src_off[s];
    float* dst = spec_p + clip * seg.dp + seg.pad_off[s];
    float* ldst = logspec_p + clip * seg.dp + seg.pad_off[s];
    float acc = 0.f;
    for (long long k = threadIdx.x; k < seg.pad_len[s]; k += blockDim.x) {
        float v = 0.f, l = 0.f;
        if (k < seg.len[s]) {
            v = src[k];
            l = logf(v + log_eps);
            acc = fmaf(v, v, acc);
        }
        dst[k] = v;     // padding: both arrays 0, so padded bins contribute nothing to either sum
        ldst[k] = l;
    }
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        sq_mean[clip * seg.n + s] = t / static_cast<float>(seg.len[s]);
    }
}

struct Stage {
    float x[KC * LDI];
    float lx[KC * LDI];
    float y[KC * LDJ];
    float ly[KC * LDJ];
};

// 256 threads: ty = tid / 16 owns rows ty*4..+3, tx = tid % 16 owns columns tx*4..+3
// Rows and columns may come from DIFFERENT prepared blocks (streaming sweep: a resident row block against column blocks
// prepared on the fly): `rows` holds clips [row_g0, row_g0 + n_rows) of the collection, `cols` clips [col_g0, col_g0 +
// n_cols); the kernel reduces rows [row_begin, row_end) x columns [col_begin, col_end) given as LOCAL indices.
struct DistOperand {
    const float* spec;      // [count, dp] padded spectrogram rows
    const float* logspec;   // [count, dp] log(x + eps)
    const float* sq_mean;   // [count, n_scales]
    long long count;        // clips in this block
    long long g0;           // index of its first clip in the whole collection
};

__global__ void __launch_bounds__(256) distance_rows_kernel(DistOperand rows, DistOperand cols, Segments seg,
                                                               long long row_begin, long long row_end,
                                                               long long col_begin, long long col_end,
                                                               float* __restrict__ out, long long ld_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Stage* st = reinterpret_cast<Stage*>(smem_raw);   // [2]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long i0 = row_begin + static_cast<long long>(blockIdx.y) * TI;
    const long long j0 = col_begin + static_cast<long long>(blockIdx.x) * TJ;

    // loader role: 4 threads per row fetch 16 consecutive bins (one float4 each) of that row
    const int l_row = tid >> 2, l_k = (tid & 3) * 4;
    const long long gi = min(i0 + l_row, rows.count - 1);
    const long long gj = min(j0 + l_row, cols.count - 1);
    const float* px = rows.spec + gi * seg.dp + l_k;
    const float* plx = rows.logspec + gi * seg.dp + l_k;
    const float* py = cols.spec + gj * seg.dp + l_k;
    const float* ply = cols.logspec + gj * seg.dp + l_k;

    float4 r[4];
    auto fetch = [&](long long k0) {
        r[0] = __ldg(reinterpret_cast<const float4*>(px + k0));
        r[1] = __ldg(reinterpret_cast<const float4*>(plx + k0));
        r[2] = __ldg(reinterpret_cast<const float4*>(py + k0));
        r[3] = __ldg(reinterpret_cast<const float4*>(ply + k0));
    };
    auto scatter = [&](float* base, int ld, int row, const float4& v) {
        base[(l_k + 0) * ld + row] = v.x;
        base[(l_k + 1) * ld + row] = v.y;
        base[(l_k + 2) * ld + row] = v.z;
        base[(l_k + 3) * ld + row] = v.w;
    };
    auto deposit = [&](Stage& s) {
        scatter(s.x, LDI, l_row, r[0]);
        scatter(s.lx, LDI, l_row, r[1]);
        scatter(s.y, LDJ, l_row, r[2]);
        scatter(s.ly, LDJ, l_row, r[3]);
    };

    float dist[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) dist[a][c] = 0.f;

    const long long stages_total = seg.dp / KC;
    fetch(0);
    deposit(st[0]);
    __syncthreads();
    long long g = 0;
    for (int s = 0; s < seg.n; ++s) {
        float sq[4][4], l1[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) sq[a][c] = l1[a][c] = 0.f;
        const long long stages = seg.pad_len[s] / KC;
        for (long long q = 0; q < stages; ++q, ++g) {
            const Stage& cur = st[g & 1];
            const bool more = (g + 1) < stages_total;
            if (more) fetch((g + 1) * KC);
            float sq_i[4][4], l1_i[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) sq_i[a][c] = l1_i[a][c] = 0.f;
#pragma unroll 4
            for (int k = 0; k < KC; ++k) {
                const float4 xv = *reinterpret_cast<const float4*>(cur.x + k * LDI + ty * 4);
                const float4 lxv = *reinterpret_cast<const float4*>(cur.lx + k * LDI + ty * 4);
                const float4 y0 = *reinterpret_cast<const float4*>(cur.y + k * LDJ + tx * 4);
                const float4 ly0 = *reinterpret_cast<const float4*>(cur.ly + k * LDJ + tx * 4);
                const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, lxs[4] = {lxv.x, lxv.y, lxv.z, lxv.w};
                const float ys[4] = {y0.x, y0.y, y0.z, y0.w};
                const float lys[4] = {ly0.x, ly0.y, ly0.z, ly0.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float dd = xs[a] - ys[c];
                        sq_i[a][c] = fmaf(dd, dd, sq_i[a][c]);
                        l1_i[a][c] += fabsf(lxs[a] - lys[c]);
                    }
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    sq[a][c] += sq_i[a][c];
                    l1[a][c] += l1_i[a][c];
                }
            if (more) deposit(st[(g + 1) & 1]);
            __syncthreads();
        }
        // fold this scale: the normaliser belongs to the lower-index clip of the pair (collection-wide indices)
        const float inv_len = 1.0f / static_cast<float>(seg.len[s]);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const long long i = min(i0 + ty * 4 + a, rows.count - 1);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const long long j = min(j0 + tx * 4 + c, cols.count - 1);
                const bool row_is_lower = rows.g0 + i <= cols.g0 + j;
                const float norm = (row_is_lower ? __ldg(rows.sq_mean + i * seg.n + s) : __ldg(cols.sq_mean + j * seg.n + s)) + 1e-7f;
                dist[a][c] += (sq[a][c] * inv_len) / norm + l1[a][c] * inv_len;
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const long long i = i0 + ty * 4 + a;
        if (i >= row_end) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const long long j = j0 + tx * 4 + c;
            if (j < col_end) out[(i - row_begin) * ld_out + (j - col_begin)] = (rows.g0 + i == cols.g0 + j) ? 0.f : dist[a][c];
        }
    }
}

}  // namespace
}  // namespace topo

using namespace topo;

extern "C" int64_t topo_distance_padded_size(const int64_t* seg_len, int n_scales) {
    if (!seg_len || n_scales < 1 || n_scales > kMaxScales) return -1;
    return make_segments(seg_len, n_scales).dp;
}

extern "C" int topo_distance_prepare(const float* spec, int64_t n, int64_t d, const int64_t* seg_len, int n_scales,
                                     float log_eps, float* spec_p, float* logspec_p, float* sq_mean,
                                     topo_stream_t stream) {
    TOPO_REQUIRE(spec && seg_len && spec_p && logspec_p && sq_mean, "null argument");
    TOPO_REQUIRE(n_scales >= 1 && n_scales <= kMaxScales, "n_scales must be in [1, 8]");
    TOPO_REQUIRE(n >= 0 && n < (int64_t(1) << 31), "bad n");
    const Segments seg = make_segments(seg_len, n_scales);
    int64_t total = 0;
    for (int i = 0; i < n_scales; ++i) {
        TOPO_REQUIRE(seg_len[i] > 0, "empty scale segment");
        total += seg_len[i];
    }
    TOPO_REQUIRE(total == d, "segment lengths do not add up to d");
    if (n == 0) return TOPO_OK;
    distance_prepare_kernel<<<dim3(static_cast<unsigned>(n), n_scales), 256, 0, as_stream(stream)>>>(
        spec, d, seg, log_eps, spec_p, logspec_p, sq_mean);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

static int launch_distance(const DistOperand& rows, const DistOperand& cols, const Segments& seg, int64_t row_begin,
                           int64_t row_end, int64_t col_begin, int64_t col_end, float* out, int64_t ld_out, topo_stream_t stream) {
    const dim3 grid(static_cast<unsigned>((col_end - col_begin + TJ - 1) / TJ),
                    static_cast<unsigned>((row_end - row_begin + TI - 1) / TI));
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(distance_rows_kernel), 2 * sizeof(Stage))) return rc;
    distance_rows_kernel<<<grid, 256, 2 * sizeof(Stage), as_stream(stream)>>>(rows, cols, seg, row_begin, row_end, col_begin,
                                                                               col_end, out, ld_out);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_distance_rows(const float* spec_p, const float* logspec_p, const float* sq_mean, int64_t n,
                                  const int64_t* seg_len, int n_scales, int64_t row_begin, int64_t row_end,
                                  int64_t col_begin, int64_t col_end, float* out, topo_stream_t stream) {
    TOPO_REQUIRE(spec_p && logspec_p && sq_mean && seg_len && out, "null argument");
    TOPO_REQUIRE(n_scales >= 1 && n_scales <= kMaxScales, "n_scales must be in [1, 8]");
    TOPO_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= n, "bad row range");
    TOPO_REQUIRE(0 <= col_begin && col_begin <= col_end && col_end <= n, "bad column range");
    if (row_begin == row_end || col_begin == col_end) return TOPO_OK;
    const Segments seg = make_segments(seg_len, n_scales);
    const DistOperand all{spec_p, logspec_p, sq_mean, n, 0};
    return launch_distance(all, all, seg, row_begin, row_end, col_begin, col_end, out, col_end - col_begin, stream);
}

extern "C" int topo_distance_block(const float* row_spec, const float* row_logspec, const float* row_sq_mean, int64_t n_rows,
                                   int64_t row_global0, const float* col_spec, const float* col_logspec,
                                   const float* col_sq_mean, int64_t n_cols, int64_t col_global0, const int64_t* seg_len,
                                   int n_scales, float* out, int64_t ld_out, topo_stream_t stream) {
    TOPO_REQUIRE(row_spec && row_logspec && row_sq_mean && col_spec && col_logspec && col_sq_mean && seg_len && out, "null argument");
    TOPO_REQUIRE(n_scales >= 1 && n_scales <= kMaxScales, "n_scales must be in [1, 8]");
    TOPO_REQUIRE(n_rows >= 0 && n_cols >= 0 && row_global0 >= 0 && col_global0 >= 0 && ld_out >= n_cols, "bad block geometry");
    if (n_rows == 0 || n_cols == 0) return TOPO_OK;
    const Segments seg = make_segments(seg_len, n_scales);
    const DistOperand rows{row_spec, row_logspec, row_sq_mean, n_rows, row_global0};
    const DistOperand cols{col_spec, col_logspec, col_sq_mean, n_cols, col_global0};
    return launch_distance(rows, cols, seg, 0, n_rows, 0, n_cols, out, ld_out, stream);
}
