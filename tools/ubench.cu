// ubench.cu -- issue/pipe throughput of the instructions the compositor's inner loop is made of (B200, sm_100a).
// Each kernel runs ITER x UNROLL independent ops per thread on 8 warps x 148x4 CTAs and reports warp-instructions
// per cycle per SM partition (1.0 = one issue slot per clock).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define ITER 2048
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const unsigned* in, long long* cyc) {
    __shared__ unsigned char sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (unsigned char)(i * 7);
    __syncthreads();
    float x[8];
    float2 y[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = 1.0f + threadIdx.x * 1e-7f + i; y[i] = make_float2(x[i], x[i] + 0.5f); u[i] = in[(threadIdx.x + i) & 255]; }
    const float m = 1.0000001f;
    const float2 m2 = make_float2(m, m);
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) x[i] = __fmul_rn(x[i], m);                                  // FMUL
            if (MODE == 1) y[i] = mul2(y[i], m2);                                      // FMUL2
            if (MODE == 2) y[i] = fma2(y[i], m2, m2);                                  // FFMA2
            if (MODE == 3) x[i] = __fmaf_rn(x[i], m, m);                               // FFMA
            if (MODE == 4) { x[i] = __uint2float_rn(u[i]); u[i] = u[i] + __float_as_uint(x[i]); }      // I2FP + IADD
            if (MODE == 5) { y[i] = mul2(y[i], m2); u[i] = (u[i] | 0x4B000000u) ^ (unsigned)it; }        // FMUL2 + LOP3
            if (MODE == 6) { x[i] = __fmul_rn(x[i], m); u[i] = (u[i] | 0x4B000000u) ^ (unsigned)it; }    // FMUL + LOP3
            if (MODE == 7) { u[i] = sm[(u[i] + it) & 4095]; }                           // dependent LDS.U8 (latency-ish)
            if (MODE == 8) { u[i] += sm[(threadIdx.x * 8 + i + it) & 4095]; }           // independent LDS.U8 + IADD
            if (MODE == 9) { y[i] = add2(y[i], m2); }                                  // FADD2
            if (MODE == 10) { x[i] = (float)(unsigned char)u[i]; u[i] = u[i] + __float_as_uint(x[i]); }  // I2F.U8 + IADD
            if (MODE == 11) { x[i] = __uint_as_float(__byte_perm(u[i], 0x4B000000u, 0x7651)); u[i] += __float_as_uint(x[i]); } // PRMT + IADD
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i].x + y[i].y + __uint_as_float(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_op) {
    float* out; unsigned* in; long long* cyc;
    const int blocks = 148 * 2;
    cudaMalloc(&out, blocks * 256 * 4); cudaMalloc(&in, 1024); cudaMalloc(&cyc, blocks * 8);
    cudaMemset(in, 1, 1024);
    k<MODE><<<blocks, 256>>>(out, in, cyc);
    k<MODE><<<blocks, 256>>>(out, in, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[296]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    // per SM: 2 CTAs x 8 warps = 16 warps = 4 per partition; each warp issues ITER*8*instr_per_op
    double per_part = 4.0 * ITER * 8 * instr_per_op;
    printf("%-28s %s  cycles %.0f  warp-instr/clk/partition %.3f\n", name, e ? cudaGetErrorString(e) : "ok", avg, per_part / avg);
}

int main() {
    run<0>("FMUL", 1); run<3>("FFMA", 1); run<1>("FMUL2", 1); run<2>("FFMA2", 1); run<9>("FADD2", 1);
    run<4>("I2FP.U32 + IADD", 2); run<10>("I2F.U8 + IADD", 2); run<11>("PRMT + IADD", 2);
    run<5>("FMUL2 + LOP3(2)", 3); run<6>("FMUL + LOP3(2)", 3);
    run<7>("LDS.U8 dependent", 1); run<8>("LDS.U8 + IADD", 2);
    return 0;
}
