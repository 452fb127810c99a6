// Probe (GPU): pair-element rate of the L1-of-logs inner loop in three forms on one B200:
//   F  two FP32 instructions per pair-element (FADD, FADD with |.|)              -- what l1_kernel does
//   I  one integer instruction per pair-element (VABSDIFF.U32 with accumulate) on fixed-point operands
//   M  two k-steps of I to one of F (the ALU pipe and the two FMA pipes side by side)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sad_rate sad_rate.cu ; prints pair-elements per clock and SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int KC = 48, LDI = 128, LDJ = 64;

template <int MODE>
__global__ void __launch_bounds__(256, 2) probe(const float* __restrict__ src, float* __restrict__ out, int iters) {
    __shared__ __align__(16) float lx[KC * LDI];
    __shared__ __align__(16) float ly[KC * LDJ];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    for (int i = tid; i < KC * LDI; i += 256) lx[i] = src[i];
    for (int i = tid; i < KC * LDJ; i += 256) ly[i] = src[KC * LDI + i];
    __syncthreads();
    float accf[8][4];
    unsigned acci[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) { accf[a][c] = 0.f; acci[a][c] = 0u; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll 6
        for (int k = 0; k < KC; ++k) {
            const float4 xa = *reinterpret_cast<const float4*>(lx + k * LDI + ty * 8);
            const float4 xb = *reinterpret_cast<const float4*>(lx + k * LDI + ty * 8 + 4);
            const float4 yv = *reinterpret_cast<const float4*>(ly + k * LDJ + tx * 4);
            const float xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const float ys[4] = {yv.x, yv.y, yv.z, yv.w};
            const bool as_int = MODE == 1 || (MODE == 2 && (k % 3) != 2);
            if (MODE >= 10) {
                constexpr int NI = MODE - 10;      // rows 0..NI-1 of the register tile on the integer pipe, the rest on the FP32 pipes
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (a < NI) acci[a][c] = __usad(__float_as_uint(xs[a]), __float_as_uint(ys[c]), acci[a][c]);
                        else accf[a][c] += fabsf(xs[a] - ys[c]);
                    }
            } else if (as_int) {
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acci[a][c] = __usad(__float_as_uint(xs[a]), __float_as_uint(ys[c]), acci[a][c]);
            } else {
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) accf[a][c] += fabsf(xs[a] - ys[c]);
            }
        }
    }
    float t = 0.f;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) t += accf[a][c] + static_cast<float>(acci[a][c]);
    out[blockIdx.x * 256 + tid] = t;
}

template <int MODE>
void run(const char* name, const float* src, float* out, int sms, double mhz) {
    const int iters = 2000, grid = sms * 2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<MODE><<<grid, 256>>>(src, out, 10);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        probe<MODE><<<grid, 256>>>(src, out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double pair_el = double(grid) * 256 * 32 * KC * iters;
    printf("%s: %.3f ms, %.3e pair-elements/s, %.1f per clock and SM at %.0f MHz (nominal FP32 lanes: 128)\n", name, best,
           pair_el / (best * 1e-3), pair_el / (best * 1e-3) / (mhz * 1e6) / sms, mhz);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *src, *out;
    cudaMalloc(&src, (KC * (LDI + LDJ)) * 4);
    cudaMalloc(&out, prop.multiProcessorCount * 2 * 256 * 4);
    cudaMemset(src, 0x3c, (KC * (LDI + LDJ)) * 4);
    run<0>("F  (2 FADD)", src, out, prop.multiProcessorCount, khz / 1e3);
    run<1>("I  (1 VABSDIFF)", src, out, prop.multiProcessorCount, khz / 1e3);
    run<2>("M  (2 k-steps I : 1 k-step F)", src, out, prop.multiProcessorCount, khz / 1e3);
    run<12>("rows 2 I : 6 F", src, out, prop.multiProcessorCount, khz / 1e3);
    run<13>("rows 3 I : 5 F", src, out, prop.multiProcessorCount, khz / 1e3);
    run<14>("rows 4 I : 4 F", src, out, prop.multiProcessorCount, khz / 1e3);
    run<15>("rows 5 I : 3 F", src, out, prop.multiProcessorCount, khz / 1e3);
    run<16>("rows 6 I : 2 F", src, out, prop.multiProcessorCount, khz / 1e3);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
