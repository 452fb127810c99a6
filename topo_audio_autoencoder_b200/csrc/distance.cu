// D1-D3: tiled pairwise multi-scale spectral distance.
// Replaces the pair loop of compute_distances (precompute_distances.py:89-115), which recomputes both multi-scale STFTs
// for every pair and scatters results in a Python loop.  Here the N spectrograms are computed once (front half, host
// side) and every pair is reduced from prepared operands:
//     d(i,j) = sum_s [ mean_s((x-y)^2) / (mean_s(x^2) + 1e-7) + mean_s |log(x+eps) - log(y+eps)| ]
// with x the LOWER-index clip (precompute_distances.py:89, 106-110), mirrored (:114-115).
//
// The two terms have different shapes and run on different pipes:
//   * relative L2:  sum (x-y)^2 = sum x^2 + sum y^2 - 2 <x, y>.  The Gram term <x, y> is a GEMM with K = 645,864: it runs
//     on the 5th-generation tensor cores as bf16x3 (tc16.cuh: three bf16 parts per fp32 value, six part products,
//     fp32-accurate) -- `gram_kernel`: operand tiles are pre-swizzled images in global memory that the TMA engine streams
//     into a two-stage shared-memory ring (cp.async.bulk + mbarrier transaction counts, one producer thread), one thread
//     issues tcgen05.mma into a ping-pong pair of tensor-memory accumulators, eight warps drain them.  Accumulation is
//     two-level so that a 130,000-term sum keeps fp32 accuracy: tensor memory holds 64 terms, registers 32 x 64,
//     the output tile the rest.
//   * L1 of logs:  sum |log x - log y| is not a GEMM -- `l1_kernel`: ONE integer instruction per pair-element (VABSDIFF with
//     accumulate) on Q6.20 fixed-point logs, 128 x 64 pair tile, operands pre-transposed in global memory (`logq`) so that a
//     stage of 32 bins is three contiguous 8 KB pieces the TMA engine drops into a four-stage shared-memory ring.
//   * `combine_kernel` folds both per scale with the lower-index clip's normaliser.
#include <algorithm>

#include "common.cuh"
#include "tc16.cuh"

namespace topo {
namespace {

using namespace tc16;

constexpr int KC = 16;            // bins per stage of the L1 kernel
constexpr int kChunk = 64;        // bins per k-chunk of the Gram kernel = one SWIZZLE_128B atom of bf16; segments are padded to it
constexpr int kRowsPerBlock = 128;    // clips per image block = M and N of the Gram tile
constexpr int kMaxScales = 8;
constexpr uint32_t kPartBytes = kRowsPerBlock * 128;     // one bf16 part of a [128 x 64] chunk: 16 KB
constexpr uint32_t kChunkBytes = 3 * kPartBytes;         // 48 KB per (clip block, k-chunk)

struct Segments {
    long long len[kMaxScales];      // true length of each scale segment
    long long pad_off[kMaxScales];  // offset in the padded row
    long long pad_len[kMaxScales];  // padded length (multiple of kChunk)
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
        s.pad_len[i] = (seg_len[i] + kChunk - 1) / kChunk * kChunk;
        s.pad_off[i] = po;
        s.src_off[i] = so;
        po += s.pad_len[i];
        so += seg_len[i];
    }
    s.dp = po;
    return s;
}

// ------------------------------------------------------------------------------------------------------------------
// prepare: one CTA per (clip, scale).  Mean of squares and the bf16x3 image of the clip's row inside its 128-clip block:
// image[block][k_chunk][part][row][64 bf16], rows swizzled exactly as tcgen05 wants them in shared memory, so that a
// (block, k_chunk) operand tile is 48 contiguous kilobytes.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) distance_prepare_kernel(const float* __restrict__ spec, long long d, Segments seg,
                                                               uint8_t* __restrict__ image, float* __restrict__ sq_mean) {
    const long long clip = blockIdx.x;
    const int s = blockIdx.y;
    const float* src = spec + clip * d + seg.src_off[s];
    const long long n_chunks = seg.dp / kChunk;
    const long long blk = clip / kRowsPerBlock;
    const int row = static_cast<int>(clip % kRowsPerBlock);
    uint8_t* img_block = image + blk * n_chunks * kChunkBytes;
    float acc = 0.f;
    // one thread per group of eight consecutive bins = one 16-byte chunk of each bf16 part
    for (long long k8 = threadIdx.x; k8 < seg.pad_len[s] / 8; k8 += blockDim.x) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const long long k = k8 * 8 + e;
            v[e] = k < seg.len[s] ? src[k] : 0.f;                     // padding: 0, contributes to neither sum
            acc = fmaf(v[e], v[e], acc);
        }
        const long long kk = seg.pad_off[s] + k8 * 8;                 // position in the padded row
        store_split8(img_block + (kk / kChunk) * kChunkBytes, kPartBytes, row, static_cast<int>((kk % kChunk) / 8), v);
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

// The L1 term's operand: log(x + eps) as unsigned Q6.20 fixed point,
//     q(v) = round((v + 40) * 2^20), clamped to [0, 2^26)          window [-40, 24): magnitudes from 4e-18 to 2.6e10
// stored k-major inside 64-clip blocks, logq[clip / 64][padded bin][clip % 64], so that the 64 clips x 32 bins one stage of
// `l1_kernel` consumes are 8 contiguous kilobytes.  q is exact for every fp32 value with |v| >= 8 and rounds to 2^-20 =
// 9.5e-7 (one fp32 ulp of the values in [8, 16)) below; the rounding is unbiased and averages over the >= 129,000 bins of a
// scale to ~1e-9 of the mean.  One CTA per (64-bin chunk, 64-clip block): coalesced reads along the bins, a transposition
// through shared memory, coalesced 256-byte rows out.  Padding bins and the clips past n in the last block hold q(0).
constexpr float kFixOffset = 40.0f, kFixScale = 1048576.0f, kFixMax = 67108863.0f;     // Q6.20
constexpr int kQBlock = 64;           // clips per block of logq

__device__ __forceinline__ uint32_t to_fixed(float v) {
    return __float2uint_rn(fminf(fmaf(v, kFixScale, kFixOffset * kFixScale), kFixMax));   // negative and NaN saturate to 0
}

__global__ void __launch_bounds__(256) distance_logq_kernel(const float* __restrict__ spec, long long n, long long d, Segments seg,
                                                            float log_eps, uint32_t* __restrict__ logq) {
    __shared__ uint32_t tile[kChunk][kQBlock + 1];
    const long long kp0 = static_cast<long long>(blockIdx.x) * kChunk;      // padded bin of this chunk (inside ONE scale)
    const long long blk = blockIdx.y;
    int s = 0;
    while (s + 1 < seg.n && kp0 >= seg.pad_off[s + 1]) ++s;
    const long long k_in_seg = kp0 - seg.pad_off[s] + (threadIdx.x & 63);
    const bool k_ok = k_in_seg < seg.len[s];
    const float* src = spec + seg.src_off[s] + k_in_seg;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int c = (threadIdx.x >> 6) + 4 * i;
        const long long clip = blk * kQBlock + c;
        const float v = (k_ok && clip < n) ? logf(__ldg(src + clip * d) + log_eps) : 0.f;
        tile[threadIdx.x & 63][c] = to_fixed(v);
    }
    __syncthreads();
    uint32_t* dst = logq + (blk * seg.dp + kp0) * kQBlock;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int k = (threadIdx.x >> 6) + 4 * i;
        dst[k * kQBlock + (threadIdx.x & 63)] = tile[k][threadIdx.x & 63];
    }
}

// Rows and columns may come from DIFFERENT prepared blocks (streaming sweep: a resident row block against column blocks
// prepared on the fly): an operand holds clips [g0, g0 + count) of the collection.
struct DistOperand {
    const uint32_t* logq;   // [ceil(count / 64)][dp][64] Q6.20 logs
    const uint8_t* image;   // [ceil(count / 128)][dp / 64] chunks of 48 KB
    const float* sq_mean;   // [count, n_scales]
    long long count;        // clips in this block
    long long g0;           // index of its first clip in the whole collection
};

// ------------------------------------------------------------------------------------------------------------------
// Gram term on the tensor cores.  grid = (column tiles, row tiles, scales); a CTA owns the 128 x 128 tile of <x, y> of one
// scale.  Warp 8 lane 0: producer (cp.async.bulk of the two 48 KB operand chunks of a stage).  Warp 9 lane 0: issues the
// 24 MMAs of a chunk (4 k-steps x 6 part products) into accumulator (chunk & 1).  Warps 0..7: drain that accumulator into
// registers (warp w: TMEM lane quarter w % 4, columns 64 (w / 4) ..), fold the registers into the output tile every 32
// chunks.  Nothing waits for anything but its own mbarrier.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kGramStages = 2;
constexpr int kGramThreads = 320;
constexpr int kFold = 32;          // chunks accumulated in registers between two folds into the output tile

struct GramSmem {
    static constexpr uint32_t kStage = 2 * kChunkBytes;                  // A then B
    static constexpr uint32_t kBar = kGramStages * kStage;               // full[2], empty[2], acc_full[2], acc_empty[2], tmem base
    static constexpr uint32_t kTotal = kBar + 128;
};

__global__ void __launch_bounds__(kGramThreads, 1) gram_kernel(DistOperand rows, DistOperand cols, Segments seg,
                                                               long long row_tile0, long long col_tile0, long long n_rows_out,
                                                               long long n_cols_out, long long row_begin, long long col_begin,
                                                               float* __restrict__ gram /* [scales][n_rows_out][n_cols_out] */) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw;
    uint64_t* full = reinterpret_cast<uint64_t*>(base + GramSmem::kBar);
    uint64_t* empty = full + 2;
    uint64_t* acc_full = full + 4;
    uint64_t* acc_empty = full + 6;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(full + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int s = blockIdx.z;
    const long long rt = row_tile0 + blockIdx.y, ct = col_tile0 + blockIdx.x;       // 128-clip blocks of the two operands
    const long long n_chunks_total = seg.dp / kChunk;
    const long long c0 = seg.pad_off[s] / kChunk, n_chunks = seg.pad_len[s] / kChunk;
    const uint8_t* a_src = rows.image + (rt * n_chunks_total + c0) * kChunkBytes;
    const uint8_t* b_src = cols.image + (ct * n_chunks_total + c0) * kChunkBytes;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 8);          // one arrival per draining warp
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_base_smem, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp == 8) {
        if (lane == 0) {
            for (long long c = 0; c < n_chunks; ++c) {
                const int st = static_cast<int>(c & 1);
                if (c >= kGramStages) mbar_wait_backoff(&empty[st], static_cast<uint32_t>(((c >> 1) - 1) & 1));
                uint8_t* dst = base + st * GramSmem::kStage;
                mbar_arrive_expect_tx(&full[st], 2 * kChunkBytes);
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    bulk_copy_g2s(dst + p * kPartBytes, a_src + c * kChunkBytes + p * kPartBytes, kPartBytes, &full[st]);
                    bulk_copy_g2s(dst + kChunkBytes + p * kPartBytes, b_src + c * kChunkBytes + p * kPartBytes, kPartBytes, &full[st]);
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            for (long long c = 0; c < n_chunks; ++c) {
                const int st = static_cast<int>(c & 1);
                mbar_wait_backoff(&full[st], static_cast<uint32_t>((c >> 1) & 1));
                if (c >= 2) mbar_wait_backoff(&acc_empty[st], static_cast<uint32_t>(((c >> 1) - 1) & 1));
                tc_fence_after_sync();
                const uint32_t a_s = smem_u32(base + st * GramSmem::kStage);
                gemm_bf16x3_unrolled<kChunk / 16>(tmem_base + st * 128, k_major(a_s, kRowsPerBlock), k_major(a_s + kChunkBytes, kRowsPerBlock),
                                                  idesc_bf16(128, 128, 0, 0), 0);
                mma_commit(&empty[st]);           // the stage's operand tiles may be overwritten
                mma_commit(&acc_full[st]);        // the accumulator holds this chunk's 64-term partial products
            }
        }
    } else {
        // ---- drain: thread (lane quarter w % 4, lane) = tile row, columns [64 (w / 4), +64) ----
        const int half = warp >> 2;
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const long long row = rt * kRowsPerBlock + (warp & 3) * 32 + lane - row_begin;         // row of the output tile
        const long long col0 = ct * kRowsPerBlock + half * 64 - col_begin;
        const bool row_ok = row >= 0 && row < n_rows_out;
        float* out_row = gram + (static_cast<long long>(s) * n_rows_out + (row_ok ? row : 0)) * n_cols_out;
        float lo[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) lo[i] = 0.f;
        bool first_fold = true;
        auto fold = [&]() {
            if (row_ok) {
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const long long col = col0 + i;
                    if (col >= 0 && col < n_cols_out) out_row[col] = first_fold ? lo[i] : out_row[col] + lo[i];
                }
            }
            first_fold = false;
#pragma unroll
            for (int i = 0; i < 64; ++i) lo[i] = 0.f;
        };
        for (long long c = 0; c < n_chunks; ++c) {
            const int st = static_cast<int>(c & 1);
            mbar_wait_backoff(&acc_full[st], static_cast<uint32_t>((c >> 1) & 1));
            tc_fence_after_sync();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v[16];
                uint32_t r[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(tmem_base + st * 128 + lane_addr + half * 64 + q * 16)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i] = __uint_as_float(r[i]);
                    lo[q * 16 + i] += v[i];
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[st])) : "memory");
            }
            if (((c + 1) % kFold) == 0) fold();
        }
        fold();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------------------------
// L1 of logs: grid = (64-clip column blocks, 128-clip row tiles, scales); a CTA reduces the 128 x 64 pair tile of one scale,
// l1[scale][i][j] = sum_k |lx_ik - ly_jk| (not yet divided by the segment length).
//
// sum |a - b| costs two FP32 instructions per pair-element (FADD, FADD with |.|) and one on the integer pipe: VABSDIFF.U32
// d = |a - b| + c on the fixed-point logs.  Measured on one B200 (scripts/probes/sad_rate.cu, operands from shared memory,
// 8 x 4 pairs per thread): 54.5 pair-elements per clock and SM with FADDs, 62.2 with VABSDIFF (the integer pipe runs at half
// the FADD rate but needs half the instructions); in this kernel the integer form also removes the accumulators' second
// level, and it makes the sum exact inside a 64-bin chunk and independent of where a pair sits in its tile.
//   warp 8, lane 0   producer: per stage of 32 bins three cp.async.bulk copies of 8 KB (row blocks 2t and 2t + 1, the
//                    column block) into a four-stage ring, armed with mbarrier transaction counts
//   warps 0..7       thread (ty, tx) = (tid / 16, tid % 16) owns rows 8 ty .. +7 and columns 4 tx .. +3: three LDS.128 and
//                    32 VABSDIFF per bin; 64 differences of 26 bits fit a 32-bit accumulator exactly (segments are padded
//                    to 64 bins), which is then added to the pair's fp32 sum; one mbarrier arrival per warp frees the stage
// No thread touches a global operand, no __syncthreads in the loop.
// ------------------------------------------------------------------------------------------------------------------
constexpr int TI = 128, TJ = 64;
constexpr int kL1KC = 32;                                   // bins per stage
constexpr int kL1Stages = 4;
constexpr int kL1Threads = 288;
constexpr uint32_t kL1Piece = kL1KC * kQBlock * 4;          // 8 KB: [32 bins][64 clips]
static_assert(kChunk % kL1KC == 0, "a 64-bin chunk is a whole number of stages");

struct L1Smem {
    static constexpr uint32_t kStage = 3 * kL1Piece;        // row block 2t, row block 2t + 1, column block
    static constexpr uint32_t kBar = kL1Stages * kStage;    // full[4], empty[4]
    static constexpr uint32_t kTotal = kBar + 64;
};

__global__ void __launch_bounds__(kL1Threads, 2) l1_kernel(DistOperand rows, DistOperand cols, Segments seg, long long row_tile0,
                                                           long long col_block0, long long row_begin, long long row_end,
                                                           long long col_begin, long long col_end,
                                                           float* __restrict__ l1_out /* [scales][rows][cols] */) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + L1Smem::kBar);
    uint64_t* empty = full + kL1Stages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int s = blockIdx.z;
    const long long rt = row_tile0 + blockIdx.y, cb = col_block0 + blockIdx.x;
    const long long stages = seg.pad_len[s] / kL1KC;            // even: the segments are padded to 64 bins

    if (tid == 0) {
        for (int i = 0; i < kL1Stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 8);                            // one arrival per consumer warp
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == 8) {
        if (lane == 0) {
            // blocks past the operand's last one (a tile that hangs over the end) are clamped: their pairs are never stored
            const long long n_rb = (rows.count + kQBlock - 1) / kQBlock, n_cb = (cols.count + kQBlock - 1) / kQBlock;
            const uint32_t* x0 = rows.logq + (min(2 * rt, n_rb - 1) * seg.dp + seg.pad_off[s]) * kQBlock;
            const uint32_t* x1 = rows.logq + (min(2 * rt + 1, n_rb - 1) * seg.dp + seg.pad_off[s]) * kQBlock;
            const uint32_t* y = cols.logq + (min(cb, n_cb - 1) * seg.dp + seg.pad_off[s]) * kQBlock;
            for (long long g = 0; g < stages; ++g) {
                const int st = static_cast<int>(g % kL1Stages);
                if (g >= kL1Stages) mbar_wait_backoff(&empty[st], static_cast<uint32_t>((g / kL1Stages - 1) & 1));
                uint8_t* dst = smem_raw + st * L1Smem::kStage;
                const long long off = g * (kL1KC * kQBlock);
                mbar_arrive_expect_tx(&full[st], L1Smem::kStage);
                bulk_copy_g2s(dst, x0 + off, kL1Piece, &full[st]);
                bulk_copy_g2s(dst + kL1Piece, x1 + off, kL1Piece, &full[st]);
                bulk_copy_g2s(dst + 2 * kL1Piece, y + off, kL1Piece, &full[st]);
            }
        }
        return;
    }

    const int tx = tid & 15, ty = tid >> 4;
    const uint32_t x_word = (ty >> 3) * (kL1Piece / 4) + (ty & 7) * 8, y_word = 2 * (kL1Piece / 4) + tx * 4;
    float l1[8][4];
    uint32_t acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            l1[a][c] = 0.f;
            acc[a][c] = 0u;
        }
    for (long long g = 0; g < stages; ++g) {
        const int st = static_cast<int>(g % kL1Stages);
        mbar_wait(&full[st], static_cast<uint32_t>((g / kL1Stages) & 1));
        const uint32_t* sx = reinterpret_cast<const uint32_t*>(smem_raw + st * L1Smem::kStage) + x_word;
        const uint32_t* sy = reinterpret_cast<const uint32_t*>(smem_raw + st * L1Smem::kStage) + y_word;
#pragma unroll 8
        for (int k = 0; k < kL1KC; ++k) {
            const uint4 xa = *reinterpret_cast<const uint4*>(sx + k * kQBlock);
            const uint4 xb = *reinterpret_cast<const uint4*>(sx + k * kQBlock + 4);
            const uint4 yv = *reinterpret_cast<const uint4*>(sy + k * kQBlock);
            const uint32_t xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const uint32_t ys[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] = __usad(xs[a], ys[c], acc[a][c]);
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[st])) : "memory");
        if ((g % (kChunk / kL1KC)) == kChunk / kL1KC - 1) {
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    l1[a][c] = fmaf(__uint2float_rn(acc[a][c]), 1.0f / kFixScale, l1[a][c]);
                    acc[a][c] = 0u;
                }
        }
    }
    const long long n_rows_out = row_end - row_begin, n_cols_out = col_end - col_begin;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const long long i = rt * TI + ty * 8 + a;
        if (i < row_begin || i >= row_end) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const long long j = cb * TJ + tx * 4 + c;
            if (j >= col_begin && j < col_end)
                l1_out[(static_cast<long long>(s) * n_rows_out + (i - row_begin)) * n_cols_out + (j - col_begin)] = l1[a][c];
        }
    }
}

// d(i,j) = sum_s [ (Sx_i + Sy_j - 2 G_ij) / len_s / (norm_lower + 1e-7) + L1_ij / len_s ],  S = len_s * mean square
__global__ void __launch_bounds__(256) combine_kernel(DistOperand rows, DistOperand cols, Segments seg, long long row_begin,
                                                      long long n_rows_out, long long col_begin, long long n_cols_out,
                                                      const float* __restrict__ gram, const float* __restrict__ l1,
                                                      float* __restrict__ out, long long ld_out) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= n_rows_out * n_cols_out) return;
    const long long i = idx / n_cols_out, j = idx % n_cols_out;
    const long long li = row_begin + i, lj = col_begin + j;
    const bool row_is_lower = rows.g0 + li <= cols.g0 + lj;
    float dist = 0.f;
    for (int s = 0; s < seg.n; ++s) {
        const float mx = __ldg(rows.sq_mean + li * seg.n + s), my = __ldg(cols.sq_mean + lj * seg.n + s);
        const float len = static_cast<float>(seg.len[s]);
        const float g = gram[(static_cast<long long>(s) * n_rows_out + i) * n_cols_out + j];
        // mean((x - y)^2) = mean x^2 + mean y^2 - 2 <x, y> / len, never negative
        const float msq = fmaxf((mx + my) - 2.0f * (g / len), 0.f);
        const float norm = (row_is_lower ? mx : my) + 1e-7f;
        dist += msq / norm + l1[(static_cast<long long>(s) * n_rows_out + i) * n_cols_out + j] / len;
    }
    out[i * ld_out + j] = (rows.g0 + li == cols.g0 + lj) ? 0.f : dist;
}

// The Gram kernel (tensor pipe + TMA engine, one CTA per SM for its 192 KB operand ring) and the L1 kernel (FP32 pipe) are
// independent until the combine: they are launched on two streams that fork from and join the caller's, so the block
// scheduler may run them side by side.  One side stream and two events per device, created on first use.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
SideStream* side_stream() {
    static SideStream per_device[64];
    int dev = 0;
    cudaGetDevice(&dev);
    SideStream& ss = per_device[dev >= 0 && dev < 64 ? dev : 0];
    if (ss.stream == nullptr) {
        if (cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming);
    }
    return &ss;
}

int launch_distance(const DistOperand& rows, const DistOperand& cols, const Segments& seg, int64_t row_begin, int64_t row_end,
                    int64_t col_begin, int64_t col_end, float* workspace, float* out, int64_t ld_out, topo_stream_t stream) {
    const int64_t nr = row_end - row_begin, nc = col_end - col_begin;
    float* gram = workspace;
    float* l1 = workspace + static_cast<int64_t>(seg.n) * nr * nc;
    cudaStream_t main_stream = as_stream(stream);
    SideStream* ss = side_stream();
    TOPO_REQUIRE(ss != nullptr, "could not create the side stream of the distance sweep");
    TOPO_CUDA(cudaEventRecord(ss->fork, main_stream));
    TOPO_CUDA(cudaStreamWaitEvent(ss->stream, ss->fork, 0));
    cudaStream_t s = ss->stream;
    {
        const int64_t rt0 = row_begin / kRowsPerBlock, rt1 = (row_end + kRowsPerBlock - 1) / kRowsPerBlock;
        const int64_t ct0 = col_begin / kRowsPerBlock, ct1 = (col_end + kRowsPerBlock - 1) / kRowsPerBlock;
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(gram_kernel), GramSmem::kTotal)) return rc;
        const dim3 grid(static_cast<unsigned>(ct1 - ct0), static_cast<unsigned>(rt1 - rt0), static_cast<unsigned>(seg.n));
        gram_kernel<<<grid, kGramThreads, GramSmem::kTotal, s>>>(rows, cols, seg, rt0, ct0, nr, nc, row_begin, col_begin, gram);
    }
    TOPO_CUDA(cudaEventRecord(ss->join, ss->stream));
    s = main_stream;
    {
        const int64_t rt0 = row_begin / TI, rt1 = (row_end + TI - 1) / TI;
        const int64_t cb0 = col_begin / TJ, cb1 = (col_end + TJ - 1) / TJ;
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(l1_kernel), L1Smem::kTotal)) return rc;
        const dim3 grid(static_cast<unsigned>(cb1 - cb0), static_cast<unsigned>(rt1 - rt0), static_cast<unsigned>(seg.n));
        l1_kernel<<<grid, kL1Threads, L1Smem::kTotal, s>>>(rows, cols, seg, rt0, cb0, row_begin, row_end, col_begin, col_end, l1);
    }
    TOPO_CUDA(cudaStreamWaitEvent(main_stream, ss->join, 0));
    {
        const int64_t total = nr * nc;
        combine_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(rows, cols, seg, row_begin, nr, col_begin, nc, gram, l1,
                                                                                   out, ld_out);
    }
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

}  // namespace
}  // namespace topo

using namespace topo;

extern "C" int64_t topo_distance_padded_size(const int64_t* seg_len, int n_scales) {
    if (!seg_len || n_scales < 1 || n_scales > kMaxScales) return -1;
    return make_segments(seg_len, n_scales).dp;
}

extern "C" int64_t topo_distance_image_bytes(int64_t n, const int64_t* seg_len, int n_scales) {
    if (!seg_len || n_scales < 1 || n_scales > kMaxScales || n < 0) return -1;
    const Segments seg = make_segments(seg_len, n_scales);
    return (n + kRowsPerBlock - 1) / kRowsPerBlock * (seg.dp / kChunk) * static_cast<int64_t>(kChunkBytes);
}

extern "C" int64_t topo_distance_logq_words(int64_t n, const int64_t* seg_len, int n_scales) {
    if (!seg_len || n_scales < 1 || n_scales > kMaxScales || n < 0) return -1;
    return (n + kQBlock - 1) / kQBlock * kQBlock * make_segments(seg_len, n_scales).dp;
}

extern "C" int64_t topo_distance_workspace_floats(int64_t n_rows, int64_t n_cols, int n_scales) {
    if (n_rows < 0 || n_cols < 0 || n_scales < 1 || n_scales > kMaxScales) return -1;
    return 2 * static_cast<int64_t>(n_scales) * n_rows * n_cols;
}

extern "C" int topo_distance_prepare(const float* spec, int64_t n, int64_t d, const int64_t* seg_len, int n_scales,
                                     float log_eps, uint32_t* logq, void* image, float* sq_mean, topo_stream_t stream) {
    TOPO_REQUIRE(spec && seg_len && logq && image && sq_mean, "null argument");
    TOPO_REQUIRE(log_eps >= 1e-17f, "log_eps below 1e-17 leaves the fixed-point window of the L1 term ([-40, 24) in log units)");
    TOPO_REQUIRE((reinterpret_cast<uintptr_t>(logq) & 15) == 0, "logq must be 16-byte aligned");
    TOPO_REQUIRE(n_scales >= 1 && n_scales <= kMaxScales, "n_scales must be in [1, 8]");
    TOPO_REQUIRE(n >= 0 && n < (int64_t(1) << 31), "bad n");
    TOPO_REQUIRE((reinterpret_cast<uintptr_t>(image) & 1023) == 0, "the operand image must be 1024-byte aligned");
    const Segments seg = make_segments(seg_len, n_scales);
    int64_t total = 0;
    for (int i = 0; i < n_scales; ++i) {
        TOPO_REQUIRE(seg_len[i] > 0, "empty scale segment");
        total += seg_len[i];
    }
    TOPO_REQUIRE(total == d, "segment lengths do not add up to d");
    if (n == 0) return TOPO_OK;
    distance_prepare_kernel<<<dim3(static_cast<unsigned>(n), n_scales), 256, 0, as_stream(stream)>>>(
        spec, d, seg, static_cast<uint8_t*>(image), sq_mean);
    distance_logq_kernel<<<dim3(static_cast<unsigned>(seg.dp / kChunk), static_cast<unsigned>((n + kQBlock - 1) / kQBlock)), 256, 0,
                           as_stream(stream)>>>(spec, n, d, seg, log_eps, logq);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}

extern "C" int topo_distance_rows(const uint32_t* logq, const void* image, const float* sq_mean, int64_t n,
                                  const int64_t* seg_len, int n_scales, int64_t row_begin, int64_t row_end,
                                  int64_t col_begin, int64_t col_end, float* workspace, float* out, topo_stream_t stream) {
    TOPO_REQUIRE(logq && image && sq_mean && seg_len && workspace && out, "null argument");
    TOPO_REQUIRE(n_scales >= 1 && n_scales <= kMaxScales, "n_scales must be in [1, 8]");
    TOPO_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= n, "bad row range");
    TOPO_REQUIRE(0 <= col_begin && col_begin <= col_end && col_end <= n, "bad column range");
    if (row_begin == row_end || col_begin == col_end) return TOPO_OK;
    const Segments seg = make_segments(seg_len, n_scales);
    const DistOperand all{logq, static_cast<const uint8_t*>(image), sq_mean, n, 0};
    return launch_distance(all, all, seg, row_begin, row_end, col_begin, col_end, workspace, out, col_end - col_begin, stream);
}

extern "C" int topo_distance_block(const uint32_t* row_logq, const void* row_image, const float* row_sq_mean, int64_t n_rows,
                                   int64_t row_global0, const uint32_t* col_logq, const void* col_image,
                                   const float* col_sq_mean, int64_t n_cols, int64_t col_global0, const int64_t* seg_len,
                                   int n_scales, float* workspace, float* out, int64_t ld_out, topo_stream_t stream) {
    TOPO_REQUIRE(row_logq && row_image && row_sq_mean && col_logq && col_image && col_sq_mean && seg_len && workspace && out,
                 "null argument");
    TOPO_REQUIRE(n_scales >= 1 && n_scales <= kMaxScales, "n_scales must be in [1, 8]");
    TOPO_REQUIRE(n_rows >= 0 && n_cols >= 0 && row_global0 >= 0 && col_global0 >= 0 && ld_out >= n_cols, "bad block geometry");
    if (n_rows == 0 || n_cols == 0) return TOPO_OK;
    const Segments seg = make_segments(seg_len, n_scales);
    const DistOperand rows{row_logq, static_cast<const uint8_t*>(row_image), row_sq_mean, n_rows, row_global0};
    const DistOperand cols{col_logq, static_cast<const uint8_t*>(col_image), col_sq_mean, n_cols, col_global0};
    return launch_distance(rows, cols, seg, 0, n_rows, 0, n_cols, workspace, out, ld_out, stream);
}
