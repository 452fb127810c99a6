// bf16x3 operand images for tcgen05.mma kind::f16 (fp32-level accuracy, one image for both transposes).
//
// Every fp32 operand is split into three bf16 parts,  x = p0 + p1 + p2  (8 + 8 + 8 mantissa bits, exact),
// and a product A.B is accumulated in fp32 tensor memory from the six part products of relative weight
// >= 2^-16:  p2.q0 + p0.q2 + p1.q1 + p1.q0 + p0.q1 + p0.q0  (each bf16 x bf16 product is exact in fp32; the
// dropped terms are <= 2^-24 relative).  Six passes of K = 16 cost the same tensor time as the three K = 8
// passes of 3xTF32, the image is 6 bytes per element instead of 8, and -- the reason it exists -- a 16-bit
// SWIZZLE_128B image is a valid operand in BOTH majors:
//   * rows = M/N index, columns = K index  -> K-major  (y = x W^T style products, contraction over columns)
//   * rows = K index,  columns = M/N index -> MN-major (x^T y weight-gradient products, contraction over rows;
//                                                       y = x W with W stored [in][out])
// so dL/dpre, dL/dm, the messages and the aggregates are staged ONCE and feed both the row GEMM and the
// row-contraction GEMM.  (kind::tf32 MN-major needs the separate SWIZZLE_128B_BASE32B image.)
//
// Image of a [R x 64] tile part: R rows of 128 bytes (64 bf16), 8-row groups of 1024 bytes, the 16-byte
// chunk index XOR-ed with (row & 7); parts are R * 128 bytes apart.
#pragma once

#include <cuda_bf16.h>

#include "tc.cuh"

namespace topo {
namespace tc16 {

using namespace tc;

__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
           (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// byte offset of the 16-byte chunk holding columns [8 * chunk, 8 * chunk + 8) of `row` inside one part
__device__ __forceinline__ uint32_t img_off(int row, int chunk /* 0..7 */) {
    return static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ uint32_t pack2(float a, float b, float& ra, float& rb) {
    uint32_t h;                                                 // a -> low half -> lower address
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
    ra = a - __uint_as_float(h << 16);                          // a bf16 is the upper half of its float
    rb = b - __uint_as_float(h & 0xffff0000u);
    return h;
}

// eight consecutive columns -> one 16-byte chunk in each of the three parts
__device__ __forceinline__ void store_split8(uint8_t* img, uint32_t part_stride, int row, int chunk, const float (&v)[8]) {
    uint32_t p[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a = v[2 * j], b = v[2 * j + 1], ra, rb;
        p[0][j] = pack2(a, b, ra, rb);
        p[1][j] = pack2(ra, rb, a, b);
        p[2][j] = pack2(a, b, ra, rb);
    }
    const uint32_t off = img_off(row, chunk);
#pragma unroll
    for (int q = 0; q < 3; ++q)
        *reinterpret_cast<uint4*>(img + q * part_stride + off) = make_uint4(p[q][0], p[q][1], p[q][2], p[q][3]);
}

// How one operand image is walked by the MMA: k-steps of 16.
struct Operand {
    uint32_t base;         // shared-memory address of part 0
    uint32_t part_stride;  // bytes between parts
    uint32_t kstep;        // bytes per k-step (16 elements of K)
    uint32_t lbo, sbo;     // descriptor fields
};
// rows = M/N index, the 64 columns = K: one 128-byte row holds the whole K extent, 32 bytes per k-step
__device__ __forceinline__ Operand k_major(uint32_t base, int rows) {
    return Operand{base, static_cast<uint32_t>(rows) * 128u, 32u, 16u, 1024u};
}
// rows = K index (k-step = 16 rows = 2 groups of 8 rows), the 64 columns = one 64-wide M/N block;
// `next_block` = byte distance to the image that supplies M/N indices 64..127 (for a 128-wide operand)
__device__ __forceinline__ Operand mn_major(uint32_t base, int rows, uint32_t next_block) {
    return Operand{base, static_cast<uint32_t>(rows) * 128u, 2048u, next_block, 1024u};
}

// D (+)= A . B over `ksteps` k-steps, fp32-accurate.  Issued by one thread: 6 * ksteps MMAs.
__device__ __forceinline__ void gemm_bf16x3(uint32_t tmem_d, const Operand& a, const Operand& b, uint32_t idesc, int ksteps,
                                            uint32_t accumulate_into) {
    uint32_t acc = accumulate_into;
#pragma unroll 1
    for (int pass = 0; pass < 6; ++pass) {       // small terms first
        const uint32_t pa = pass == 0 ? 2u : ((pass == 2 || pass == 3) ? 1u : 0u);
        const uint32_t pb = pass == 1 ? 2u : ((pass == 2 || pass == 4) ? 1u : 0u);
        const uint32_t a0 = a.base + pa * a.part_stride, b0 = b.base + pb * b.part_stride;
#pragma unroll 1
        for (int k = 0; k < ksteps; ++k) {
            mma_bf16(tmem_d, smem_desc_sw128(a0 + k * a.kstep, a.lbo, a.sbo), smem_desc_sw128(b0 + k * b.kstep, b.lbo, b.sbo), idesc, acc);
            acc = 1;
        }
    }
}

// Same product with the pass / k-step structure known at compile time: the issuing thread executes one
// 32-bit add per operand and the MMA itself (the descriptor's 14-bit start-address field advances by the byte
// offset >> 4; shared-memory addresses stay below 256 KB, so the add never carries out of the field).
// K0: first k-step (several threads may each issue a k-range of one product into the same, already initialised
// accumulator: the sum is order independent up to fp32 rounding).
template <int KSTEPS, int K0 = 0>
__device__ __forceinline__ void gemm_bf16x3_unrolled(uint32_t tmem_d, const Operand& a, const Operand& b, uint32_t idesc,
                                                     uint32_t accumulate_into) {
    const uint64_t a0 = smem_desc_sw128(a.base, a.lbo, a.sbo), b0 = smem_desc_sw128(b.base, b.lbo, b.sbo);
    uint32_t acc = accumulate_into;
#pragma unroll
    for (int pass = 0; pass < 6; ++pass) {
        const uint32_t pa = pass == 0 ? 2u : ((pass == 2 || pass == 3) ? 1u : 0u);
        const uint32_t pb = pass == 1 ? 2u : ((pass == 2 || pass == 4) ? 1u : 0u);
#pragma unroll
        for (int k = K0; k < K0 + KSTEPS; ++k) {
            mma_bf16(tmem_d, a0 + ((pa * a.part_stride + k * a.kstep) >> 4), b0 + ((pb * b.part_stride + k * b.kstep) >> 4), idesc, acc);
            acc = 1;
        }
    }
}

// hint: bring `bytes` (multiple of 16) starting at `p` into L2 ahead of the loads that will want them
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ void prefetch_l2_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

}  // namespace tc16
}  // namespace topo
