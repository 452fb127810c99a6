// Tile-fragment layout of the activations the combine forward saves for its own backward (m_k, pre_k).
//
// These [rows, 64] tensors are private to the two tensor-core kernels, whose epilogues own data as
// "thread (q, r) holds columns [16 q, 16 q + 16) of tile row r" (r = the tensor-memory lane).  Stored
// row-major, every 128-bit access of a warp touches 32 different 128-byte lines.  Stored as
//     float4 index = tile * 2048 + (q * 4 + j) * 128 + r        (j = 0..3: the thread's j-th float4)
// i.e. element (row, col) -> tile = row / 128, r = row % 128, q = col / 16, j = (col % 16) / 4, e = col % 4,
// consecutive lanes (r) are 16 bytes apart: each warp access is one contiguous 512-byte run.  The layout is a
// permutation inside each 128-row tile, so buffers are allocated with rows rounded up to a multiple of 128.
#pragma once

// bit 0: the forward's row-map accesses go through the lane transpose below, bit 1: the backward's
#ifndef TOPO_ROWMAP_TRANSPOSE
#define TOPO_ROWMAP_TRANSPOSE 1
#endif

namespace topo {

// float4 index of (q, r, j = 0); add j * 128 for the other three
__device__ __forceinline__ long long tf_index(long long row0 /* multiple of 128 */, int q, int r) {
    return (row0 >> 7) * 2048 + q * 512 + r;
}

// the same element range seen from the chunk map: columns [8 c, 8 c + 8) of tile row r are the float4 pair
// j = 2 (c % 2), 2 (c % 2) + 1 of column group q = c / 2
__device__ __forceinline__ long long tf_index_chunk(long long row0, int r, int c) {
    return (row0 >> 7) * 2048 + ((c >> 1) * 4 + (c & 1) * 2) * 128 + r;
}

// Row map <-> row-major global memory.  In the row map lane l of a warp holds 16 consecutive floats (four 16-byte chunks) of
// tile row l: a 128-bit access per lane touches 16 bytes in each of 32 different 128-byte lines (32 L1 wavefronts).  A 4 x 4
// transpose of the chunks inside every group of four lanes (two butterfly steps, 16 shuffles) leaves lane 4g + j with chunk j of
// rows 4g .. 4g + 3, so that access number i of the warp covers 64 contiguous bytes of rows 4g + i, g = 0..7: 8 wavefronts.
// The transpose is an involution: loads fetch in the transposed arrangement and apply it once more.
__device__ __forceinline__ void rowmap_transpose4(float4 (&v)[4], int lane) {
    const bool hi2 = (lane & 2) != 0, hi1 = (lane & 1) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {                       // distance 2: slot i + 2 of lanes 0, 1 <-> slot i of lanes 2, 3
        const float4 s = hi2 ? v[i] : v[i + 2];
        float4 t;
        t.x = __shfl_xor_sync(0xffffffffu, s.x, 2);
        t.y = __shfl_xor_sync(0xffffffffu, s.y, 2);
        t.z = __shfl_xor_sync(0xffffffffu, s.z, 2);
        t.w = __shfl_xor_sync(0xffffffffu, s.w, 2);
        if (hi2) v[i] = t; else v[i + 2] = t;
    }
#pragma unroll
    for (int i = 0; i < 4; i += 2) {                    // distance 1: slot i + 1 of even lanes <-> slot i of odd lanes
        const float4 s = hi1 ? v[i] : v[i + 1];
        float4 t;
        t.x = __shfl_xor_sync(0xffffffffu, s.x, 1);
        t.y = __shfl_xor_sync(0xffffffffu, s.y, 1);
        t.z = __shfl_xor_sync(0xffffffffu, s.z, 1);
        t.w = __shfl_xor_sync(0xffffffffu, s.w, 1);
        if (hi1) v[i] = t; else v[i + 1] = t;
    }
}

// v = this lane's 16 floats of row `row` (columns [col0, col0 + 16) of a [rows, 64] fp32 tensor); rows >= live are not written.
// Must be called by all 32 lanes of the warp.  TRANSPOSED = false is the direct access (one 16-byte piece of 32 rows per access).
template <bool TRANSPOSED>
__device__ __forceinline__ void rowmap_store16(float* __restrict__ base, long long row, int col0, long long live, const float (&v)[16], int lane) {
    float4 c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    if (TRANSPOSED) {
        rowmap_transpose4(c, lane);
        const long long g0 = row - (lane & 3);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (g0 + i < live) reinterpret_cast<float4*>(base + (g0 + i) * 64 + col0)[lane & 3] = c[i];
    } else if (row < live) {
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(base + row * 64 + col0)[i] = c[i];
    }
}

// the matching load in two halves, so that the loads can be issued long before their first use: rowmap_load16_issue fetches
// the transposed arrangement (rows >= live read as zero; LOAD is the 128-bit load to use), rowmap_load16_finish turns it into
// this lane's own 16 floats (all 32 lanes of the warp)
template <bool TRANSPOSED, typename LOAD>
__device__ __forceinline__ void rowmap_load16_issue(const float* __restrict__ base, long long row, int col0, long long live, bool enabled,
                                                    float4 (&c)[4], int lane, LOAD load) {
    const long long g0 = TRANSPOSED ? row - (lane & 3) : row;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        c[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (TRANSPOSED) {
            if (enabled && g0 + i < live) c[i] = load(reinterpret_cast<const float4*>(base + (g0 + i) * 64 + col0) + (lane & 3));
        } else {
            if (enabled && row < live) c[i] = load(reinterpret_cast<const float4*>(base + row * 64 + col0) + i);
        }
    }
}

template <bool TRANSPOSED>
__device__ __forceinline__ void rowmap_load16_finish(float4 (&c)[4], float (&v)[16], int lane) {
    if (TRANSPOSED) rowmap_transpose4(c, lane);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[4 * i] = c[i].x; v[4 * i + 1] = c[i].y; v[4 * i + 2] = c[i].z; v[4 * i + 3] = c[i].w; }
}

}  // namespace topo
