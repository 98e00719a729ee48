// Issue / pipe rates on sm_100a that the epilogue cycle models in DESIGN.md rest on: cycles per warp instruction per
// sub-partition for FFMA, FFMA2 (packed f32x2), FMUL2 + FADD2, MUFU.EX2, MUFU.RCP, I2IP, VIMNMX3, and mixes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, float a, float b) {
    float2 r[8];
    for (int i = 0; i < 8; ++i) r[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    int n[8];
    for (int i = 0; i < 8; ++i) n[i] = threadIdx.x + i;
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { r[i].x = __fmaf_rn(r[i].x, a, b); r[i].y = __fmaf_rn(r[i].y, a, b); }          // 2 FFMA
            if (MODE == 1) { r[i] = __ffma2_rn(r[i], a2, b2); }                                                // 1 FFMA2
            if (MODE == 2) { r[i] = __ffma2_rn(r[i], a2, b2); r[i] = __ffma2_rn(r[i], b2, a2); }               // 2 FFMA2
            if (MODE == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i].x)); }                      // 1 MUFU
            if (MODE == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i].x)); r[i] = __ffma2_rn(r[i], a2, b2); r[i] = __ffma2_rn(r[i], b2, a2); }
            if (MODE == 5) { n[i] = max(n[i] + it, max(n[(i + 1) & 7], n[(i + 2) & 7])); }                      // IADD + VIMNMX3
            if (MODE == 6) { r[i] = __fmul2_rn(r[i], a2); r[i] = __fadd2_rn(r[i], b2); }                       // FMUL2 + FADD2 (or a fused FFMA2)
            if (MODE == 7) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(r[i].x)); }
            if (MODE == 8) { r[i].x = fmaxf(r[i].x, a); r[i].y = fminf(r[i].y, b); }                           // 2 FMNMX
            if (MODE == 9) { n[i] = n[i] + it; n[i] ^= 0x4b000000; }                                           // IADD + LOP3
            if (MODE == 10) { r[i] = __ffma2_rn(r[i], a2, b2); n[i] = n[i] + it; r[i] = __ffma2_rn(r[i], b2, a2); n[i] ^= 0x4b000000; }   // 2 FFMA2 + 2 ALU
            if (MODE == 11) {                                                                                  // epilogue mix 4 FFMA2 : 3 ALU : 1 MUFU
                r[i] = __ffma2_rn(r[i], a2, b2); n[i] = n[i] + it; r[i] = __ffma2_rn(r[i], b2, a2); n[i] ^= 0x4b000000;
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i].x));
                r[i] = __ffma2_rn(r[i], a2, b2); n[(i + 3) & 7] = max(n[(i + 3) & 7], n[i]); r[i] = __ffma2_rn(r[i], b2, a2);
            }
            if (MODE == 12) { r[i].x = __fmaf_rn(r[i].x, a, b); n[i] = n[i] + it; r[i].y = __fmaf_rn(r[i].y, a, b); n[i] ^= 0x4b000000; }   // 2 FFMA + 2 ALU
            if (MODE == 14) { r[i].x += __int2float_rn(n[i]); n[i] += it; }                                   // I2FP + FADD + IADD
            if (MODE == 15) { r[i].x += __int_as_float(n[i]); n[i] += it; }                                   // FADD + IADD (reference for 14)
            if (MODE == 13) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i].x)); n[i] = n[i] + it; n[i] ^= 0x4b000000; n[(i + 3) & 7] = max(n[(i + 3) & 7], n[i]); n[i] += 3; }   // MUFU + 4 ALU
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += r[i].x + r[i].y + n[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int per_iter_instr, int threads) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    k<MODE><<<148, threads>>>(out, cyc, 1.0001f, 0.5f);
    k<MODE><<<148, threads>>>(out, cyc, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    const int warps_per_smsp = threads / 32 / 4;
    const double per = (double)h[0] / ITERS / 8.0 / per_iter_instr / warps_per_smsp;
    printf("%-40s %2d warps/SMSP: %.2f cycles per warp instruction per SMSP (%s)\n", name, warps_per_smsp, per, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int th : {128, 256, 512}) {
        run<0>("FFMA (3 reg) x2", 2, th);
        run<1>("FFMA2 x1", 1, th);
        run<2>("FFMA2 x2 dependent pair", 2, th);
        run<3>("MUFU.EX2", 1, th);
        run<4>("MUFU.EX2 + 2 FFMA2", 3, th);
        run<5>("IADD + VIMNMX3", 2, th);
        run<6>("FMUL2 + FADD2 (-> FFMA2?)", 2, th);
        run<7>("MUFU.RCP", 1, th);
        run<8>("FMNMX x2", 2, th);
        run<9>("IADD + LOP3", 2, th);
        run<10>("2 FFMA2 + 2 ALU interleaved", 4, th);
        run<11>("mix 4 FFMA2 + 3 ALU + 1 MUFU", 8, th);
        run<12>("2 FFMA + 2 ALU interleaved", 4, th);
        run<13>("1 MUFU + 4 ALU", 5, th);
        run<14>("I2FP + FADD + IADD", 3, th);
        run<15>("FADD + IADD", 2, th);
    }
    return 0;
}
