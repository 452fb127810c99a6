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

}  // namespace topo
