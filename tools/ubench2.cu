// ubench2.cu -- pipe / issue rates of the compositor's instruction kinds with REGISTER operands (ubench.cu used immediates for the
// scalar forms, which issue faster), alone and mixed, at the compositor's occupancy (3 CTAs x 8 warps per SM = 6 warps per partition).
// Reports warp-instructions per clock per SM partition.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench2 ubench2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define ITER 1024
#define NCH 8

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int MODE>
__global__ void __launch_bounds__(256, 3) k(float* out, const float* in, long long* cyc) {
    __shared__ __align__(16) unsigned char sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (unsigned char)(i * 7);
    __syncthreads();
    float x[NCH], m[NCH];
    unsigned long long y[NCH], m2[NCH];
    unsigned u[NCH], a[NCH];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        x[i] = in[threadIdx.x + i];
        m[i] = in[256 + i + (threadIdx.x & 1)];
        float lo = x[i], hi = x[i] + 0.5f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(y[i]) : "f"(lo), "f"(hi));
        asm("mov.b64 %0, {%1, %2};" : "=l"(m2[i]) : "f"(m[i]), "f"(m[i]));
        u[i] = __float_as_uint(in[i + 300 + threadIdx.x]);
        a[i] = sbase + ((threadIdx.x * 5 + i * 257) & 8191);
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MODE == 0) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(m[i]));
            if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(m[i]), "f"(m[(i + 1) & 7]));
            if (MODE == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(m2[i]));
            if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(m2[i]), "l"(m2[(i + 1) & 7]));
            if (MODE == 4) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(m[i]));
            if (MODE == 5) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(x[i]) : "r"(u[i]));  // I2FP, independent
            if (MODE == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(a[i]), "r"(a[(i + 1) & 7]));
            if (MODE == 7) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(a[i]), "r"(a[(i + 1) & 7]));
            if (MODE == 8) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(a[i]));
            if (MODE == 9) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(u[i]) : "r"(a[i]));
            if (MODE == 10) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u[i]) : "r"(a[i] & ~3u));
            if (MODE == 11) {  // FMUL2 + IADD 1:1
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(m2[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(a[i]));
            }
            if (MODE == 12) {  // FMUL2 + LDS.U8 1:1
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(m2[i]));
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(u[i]) : "r"(a[i]));
            }
            if (MODE == 13) {  // FMUL + IADD 1:1
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(m[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(a[i]));
            }
            if (MODE == 14) {  // the tap chain: LDS.U8 -> I2FP -> (pairs) FMUL2 + FFMA2   = 2 + 2 + 2 instructions per 2 taps
                unsigned b0, b1;
                float f0, f1;
                unsigned long long p, q;
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b0) : "r"(a[i]));
                asm volatile("ld.shared.u8 %0, [%1+1];" : "=r"(b1) : "r"(a[i]));
                asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f0) : "r"(b0));
                asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f1) : "r"(b1));
                asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(f0), "f"(f1));
                asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(q) : "l"(p), "l"(m2[i]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(y[i]) : "l"(q), "l"(m2[(i + 1) & 7]));
            }
            if (MODE == 15) {  // the same chain, scalar arithmetic: 2 + 2 + 4
                unsigned b0, b1;
                float f0, f1, q0, q1;
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b0) : "r"(a[i]));
                asm volatile("ld.shared.u8 %0, [%1+1];" : "=r"(b1) : "r"(a[i]));
                asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f0) : "r"(b0));
                asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f1) : "r"(b1));
                asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(q0) : "f"(f0), "f"(m[i]));
                asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(q1) : "f"(f1), "f"(m[i]));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(x[i]) : "f"(q0), "f"(m[(i + 1) & 7]));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(x[(i + 4) & 7]) : "f"(q1), "f"(m[(i + 1) & 7]));
            }
            if (MODE == 16) {  // FMUL2 + FMUL 1:1 (do the scalar and the packed forms share the pipe?)
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(m2[i]));
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(m[i]));
            }
            if (MODE == 17) {  // FMUL2 : IADD 1:2
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(m2[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(a[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(u[(i + 1) & 7]));
            }
            if (MODE == 18) {  // I2FP + IADD 1:1 (independent)
                asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(x[i]) : "r"(u[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(u[(i + 1) & 7]));
            }
            if (MODE == 19) {  // FMUL + I2FP 1:1 (is I2FP on the FMA side?)
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(m[(i + 1) & 7]));
                asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(x[i]) : "r"(u[i]));
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(y[i]));
        s += x[i] + lo + hi + __uint_as_float(u[i]) + m[i] + __uint_as_float(a[i]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_op) {
    float *out, *in;
    long long* cyc;
    const int blocks = 148 * 3;
    cudaMalloc(&out, blocks * 256 * 4), cudaMalloc(&in, 4096), cudaMalloc(&cyc, blocks * 8);
    float h_in[1024];
    for (int i = 0; i < 1024; ++i) h_in[i] = 1.0f + (i % 17) * 1e-7f;
    cudaMemcpy(in, h_in, 4096, cudaMemcpyHostToDevice);
    k<MODE><<<blocks, 256>>>(out, in, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, in, cyc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 3];
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += h[i];
    avg /= blocks;
    // per SM: 3 CTAs x 8 warps = 24 warps = 6 per partition; each warp issues ITER*NCH*instr_per_op
    double per_part = 6.0 * ITER * NCH * instr_per_op;
    printf("%-34s %s  clk %.0f  warp-instr/clk/partition %.3f   (event: %.3f ms -> %.3f at 1.965 GHz)\n", name, e ? cudaGetErrorString(e) : "ok", avg, per_part / avg, ms,
           per_part / (ms * 1e-3 * 1.965e9));
    cudaFree(out), cudaFree(in), cudaFree(cyc);
}

int main() {
    run<0>("FMUL rrr", 1), run<4>("FADD rrr", 1), run<1>("FFMA rrrr", 1), run<2>("FMUL2 rr", 1), run<3>("FFMA2 rrr", 1);
    run<5>("I2FP", 1), run<6>("LOP3", 1), run<7>("PRMT", 1), run<8>("IADD", 1), run<9>("LDS.U8", 1), run<10>("LDS.32", 1);
    run<11>("FMUL2 + IADD", 2), run<17>("FMUL2 + 2 IADD", 3), run<12>("FMUL2 + LDS.U8", 2), run<13>("FMUL + IADD", 2), run<16>("FMUL2 + FMUL", 2);
    run<18>("I2FP + IADD", 2), run<19>("FMUL + I2FP", 2);
    run<14>("tap chain packed (6 per 2 taps)", 6), run<15>("tap chain scalar (8 per 2 taps)", 8);
    return 0;
}
