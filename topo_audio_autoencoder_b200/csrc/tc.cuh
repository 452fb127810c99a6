// tcgen05 / TMEM / mbarrier primitives for sm_100a, written as inline PTX.
//
// Dense fp32-accurate GEMMs of the SCCN combine run on the 5th-generation tensor cores as
// 3xTF32: every fp32 operand is split into a TF32 "hi" part and a TF32 "lo" residual,
//     x = hi + lo,   hi = rna_tf32(x),   lo = x - hi   (exact in fp32)
// and  A.B  ~=  A_lo.B_hi + A_hi.B_lo + A_hi.B_hi  is accumulated in fp32 in tensor memory.
// The dropped A_lo.B_lo term is ~2^-22 relative, the same order as an fp32 FFMA chain's rounding.
//
// Operand tiles live in shared memory in the canonical K-major SWIZZLE_128B layout: rows of 128 bytes
// (32 fp32), groups of 8 rows (1024 B, "swizzle atom"), the 16-byte chunk index XOR-ed with the row
// index inside the atom; a 64-wide K extent is two such atoms.  The A operand may instead live in tensor
// memory (lane = row, column = k), written by the threads that own the rows with tcgen05.st.
// (MN-major fp32 operands would need SWIZZLE_128B_BASE32B, a different image; not used.)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace topo {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol error traps (the launch fails) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}

// Poll with a short back-off: the waiting threads must not starve the one thread that issues the MMAs.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(32);
        if (++spins > (1u << 24)) __trap();
    }
}

// ---- bulk asynchronous copies (the TMA engine's linear mode: cp.async.bulk, SASS UBLKCP) -------------------------
// One thread arms the barrier with the byte count and issues the copies; the engine moves the data without occupying
// a single LSU slot and completes the transaction count on the barrier.  dst / src / bytes: multiples of 16.
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- named barriers: producers arrive without blocking, the consumer warp syncs ----------------
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// 128-bit read-only load that the compiler may not sink towards its first use (a prefetch into registers
// must be ISSUED where it is written: plain __ldg loads get moved past the work they were meant to overlap)
__device__ __forceinline__ float4 ldg_pinned(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// the same for data that is read exactly once (saved activations in the backward): evict-first in L1 and L2, so that
// the streams that ARE re-read soon (gradients handed to the next kernel) keep their L2 lines
__device__ __forceinline__ float4 ldg_pinned_once(const float4* p) {
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// ---- proxy / tcgen05 fences ------------------------------------------------------------------
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- tensor memory ----------------------------------------------------------------------------
// One full warp allocates `cols` (power of two >= 32) columns and publishes the base address in smem.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}

// 32 lanes x 32 columns of fp32: thread i of the warp receives columns [col, col+32) of TMEM lane
// (lane_base + i).  A warp may only touch the lane quarter 32*(warp_id % 4).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 8 columns (small code footprint for looped epilogues)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// registers -> tensor memory, 32 lanes x 8 columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors --------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor:
// start >> 4 in [0,14), LBO >> 4 in [16,30), SBO >> 4 in [32,46), version 1 in [46,48), layout in [61,64)).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
           (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (static_cast<uint64_t>(1) << 46) |
           (static_cast<uint64_t>(2) << 61);
}

// Instruction descriptor for kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor).
// a_mn / b_mn: 1 = the operand is MN-major, 0 = K-major.
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
           (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same with the A operand in tensor memory (128 lanes x 8 columns of 32-bit values per MMA).
__device__ __forceinline__ void mma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// All previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- 3xTF32 operand tiles --------------------------------------------------------------------------
__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// Byte offset of the 16-byte chunk holding columns [4*chunk, 4*chunk+4) of `row` in a [rows x 64] fp32 tile
// (two 32-column swizzle atoms, each rows*128 bytes).
__device__ __forceinline__ uint32_t tile_chunk_offset(int rows, int row, int chunk /* 0..15 */) {
    const int atom = chunk >> 3, c = chunk & 7;
    return static_cast<uint32_t>(atom * rows * 128 + (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4));
}

// Split four consecutive values and store them into the hi and lo tiles.
__device__ __forceinline__ void store_split(uint8_t* hi_tile, uint8_t* lo_tile, int rows, int row, int chunk, float4 v) {
    float4 h, l;
    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    const uint32_t off = tile_chunk_offset(rows, row, chunk);
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
}

// D[128 x 64] (+)= A[128 x 64] . B^T with B stored as [64(N) x 64(K)], both K-major tiles; 3xTF32.
// `first` = 0 overwrites D.  Issued by one thread: 24 MMAs of shape 128x64x8.
__device__ __forceinline__ void gemm_128x64x64_3xtf32(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                                      uint32_t b_lo, uint32_t accumulate_into) {
    constexpr uint32_t idesc = idesc_tf32(128, 64, 0, 0);
    uint32_t acc = accumulate_into;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {       // small terms first
        const uint32_t a = pass == 0 ? a_lo : a_hi;
        const uint32_t b = pass == 1 ? b_lo : b_hi;
#pragma unroll
        for (int k = 0; k < 8; ++k) {            // 8 columns (32 B) per step; 4 steps per 128-B atom
            const uint32_t a_off = (k >> 2) * (128 * 128) + (k & 3) * 32;
            const uint32_t b_off = (k >> 2) * (64 * 128) + (k & 3) * 32;
            mma_tf32(tmem_d, smem_desc_sw128(a + a_off, 16, 1024), smem_desc_sw128(b + b_off, 16, 1024), idesc, acc);
            acc = 1;
        }
    }
}

}  // namespace tc
}  // namespace topo
