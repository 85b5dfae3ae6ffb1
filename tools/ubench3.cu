// ubench3.cu -- what an instruction of the compositor's inner loop costs by OPERAND FORM: register, broadcast scalar, immediate.
// ubench.cu / ubench2.cu left one question open: inside svb_mix_ring's layer bodies a scheduler issues 0.73 instructions per
// clock with a ready warp nearly always at hand (profiles/r2_history.md section 8) -- which resource is that?  Every probe below is a
// stream of INDEPENDENT instructions (destinations are never sources: no dependency stalls, 4 warps per scheduler), so what is
// measured is the issue / pipe / register-read rate of the form itself.  (Destinations feed back as first sources: twelve independent
// chains per thread, or nvcc folds the loop's identical iterations into one.)
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench3 ubench3.cu && ./ubench3
#include <cuda_runtime.h>
#include <cstdio>

#define ITER 512
#define N 12  // independent destinations per thread

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, const float* in, long long* cyc) {
    float s[N], t[N], d[N];
    float Dx[N], Dy[N], Tx[N], Ty[N];
    unsigned u[N], v[N], w[N], x2[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        s[i] = in[threadIdx.x + i], t[i] = in[64 + i + (threadIdx.x & 3)], d[i] = 1.f;
        Dx[i] = s[i], Dy[i] = s[i] + 0.25f, Tx[i] = t[i], Ty[i] = t[i] + 0.5f;
        u[i] = __float_as_uint(in[128 + i + threadIdx.x]), v[i] = __float_as_uint(in[200 + i + (threadIdx.x & 7)]), w[i] = u[i];
    }
    const float one = in[500];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int j = (i + 5) % N;
            if (MODE == 0) asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; mul.rn.f32x2 a, a, b; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]) : "f"(Tx[j]), "f"(Ty[j]));
            if (MODE == 1) asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 k, 0x3b8080813b808081; mul.rn.f32x2 a, a, k; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]));
            if (MODE == 2) asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %2}; mul.rn.f32x2 a, a, b; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]) : "f"(t[j]));
            if (MODE == 3) asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; mov.b64 c, {%4, %5}; fma.rn.f32x2 a, a, b, c; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]) : "f"(Tx[j]), "f"(Ty[j]), "f"(Tx[(j + 3) % N]), "f"(Ty[(j + 3) % N]));
            if (MODE == 4) asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 k, 0x3b8080813b808081; mov.b64 c, {%2, %3}; fma.rn.f32x2 a, a, k, c; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]) : "f"(Tx[j]), "f"(Ty[j]));
            if (MODE == 5) asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %2}; mov.b64 c, {%3, %4}; fma.rn.f32x2 a, a, b, c; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]) : "f"(one), "f"(Tx[j]), "f"(Ty[j]));
            if (MODE == 6) asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %2}; mov.b64 k, 0x4b0000004b000000; fma.rn.f32x2 a, a, b, k; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]) : "f"(one));
            if (MODE == 7) asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; add.rn.f32x2 a, a, b; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]) : "f"(Tx[j]), "f"(Ty[j]));
            if (MODE == 10) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(d[i]) : "f"(t[j]));
            if (MODE == 11) asm volatile("mul.rn.f32 %0, %0, 0f3b808081;" : "+f"(d[i]));
            if (MODE == 12) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(d[i]) : "f"(t[j]), "f"(s[j]));
            if (MODE == 13) asm volatile("fma.rn.f32 %0, %0, 0f3b808081, %1;" : "+f"(d[i]) : "f"(t[j]));
            if (MODE == 14) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(d[i]) : "f"(t[j]));
            if (MODE == 20) asm volatile("add.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(v[j]));
            if (MODE == 21) asm volatile("add.u32 %0, %0, 77;" : "+r"(w[i]));
            if (MODE == 22) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i]) : "r"(v[j]), "r"(u[j]));
            if (MODE == 23) asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(w[i]));
            if (MODE == 24) asm volatile("prmt.b32 %0, %0, %1, 0x0040;" : "+r"(w[i]) : "r"(v[j]));
            if (MODE == 25) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(v[j]), "r"(u[j]));
            if (MODE == 26) asm volatile("mov.b32 %0, %0;" : "+r"(w[i]));
            if (MODE == 30) {  // FMUL2 rr + IADD rr
                asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; mul.rn.f32x2 a, a, b; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]) : "f"(Tx[j]), "f"(Ty[j]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(v[j]));
            }
            if (MODE == 31) {  // FFMA2 rrr + IADD rr
                asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; mov.b64 c, {%4, %5}; fma.rn.f32x2 a, a, b, c; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]) : "f"(Tx[j]), "f"(Ty[j]), "f"(Tx[(j + 3) % N]), "f"(Ty[(j + 3) % N]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(v[j]));
            }
            if (MODE == 32) {  // FMUL2 r,imm + IADD rr
                asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 k, 0x3b8080813b808081; mul.rn.f32x2 a, a, k; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(v[j]));
            }
            if (MODE == 33) {  // FMUL2 rr + FMUL rr
                asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; mul.rn.f32x2 a, a, b; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]) : "f"(Tx[j]), "f"(Ty[j]));
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(d[i]) : "f"(t[j]));
            }
            if (MODE == 34) {  // FMUL2 rr + I2FP
                asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; mul.rn.f32x2 a, a, b; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]) : "f"(Tx[j]), "f"(Ty[j]));
                asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(w[i]));
            }
            if (MODE == 35) {  // FFMA2 rrr + 2 x IADD r,imm
                asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; mov.b64 c, {%4, %5}; fma.rn.f32x2 a, a, b, c; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]) : "f"(Tx[j]), "f"(Ty[j]), "f"(Tx[(j + 3) % N]), "f"(Ty[(j + 3) % N]));
                asm volatile("add.u32 %0, %0, 77;" : "+r"(w[i]));
                asm volatile("add.u32 %0, %1, 78;" : "=r"(v[i]) : "r"(u[j]));
            }
            if (MODE == 36) {  // FMUL2 rr + 2 x IADD rr
                asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; mul.rn.f32x2 a, a, b; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]) : "f"(Tx[j]), "f"(Ty[j]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(v[j]));
                asm volatile("add.u32 %0, %1, %2;" : "=r"(x2[i]) : "r"(u[j]), "r"(v[i]));
            }
            if (MODE == 37) {  // FMUL2 rr + IMAD rrr
                asm volatile("{.reg .b64 a, b, c, k; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %3}; mul.rn.f32x2 a, a, b; mov.b64 {%0, %1}, a;}" : "+f"(Dx[i]), "+f"(Dy[i]) : "f"(Tx[j]), "f"(Ty[j]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(v[j]), "r"(u[j]));
            }
            if (MODE == 38) {  // FMUL rr + IADD rr
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(d[i]) : "f"(t[j]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(v[j]));
            }
            if (MODE == 39) {  // FFMA rrr + IADD rr
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(d[i]) : "f"(t[j]), "f"(s[j]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(v[j]));
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float lo = Dx[i], hi = Dy[i];
        acc += d[i] + lo + hi + __uint_as_float(w[i]) + __uint_as_float(v[i]) + (MODE == 36 ? __uint_as_float(x2[i]) : 0.f);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_op) {
    float *out, *in;
    long long* cyc;
    const int blocks = 148;
    cudaMalloc(&out, blocks * 512 * 4), cudaMalloc(&in, 8192), cudaMalloc(&cyc, blocks * 8);
    float h_in[2048];
    for (int i = 0; i < 2048; ++i) h_in[i] = 1.0f + (i % 17) * 1e-7f;
    h_in[500] = 1.0f;
    cudaMemcpy(in, h_in, 8192, cudaMemcpyHostToDevice);
    k<MODE><<<blocks, 512>>>(out, in, cyc);
    k<MODE><<<blocks, 512>>>(out, in, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += h[i];
    avg /= blocks;
    const double per_part = 4.0 * ITER * N * per_op;  // 16 warps per SM = 4 per scheduler
    printf("%-28s %s  clocks %8.0f   instructions per clock per scheduler %.3f   clocks per instruction %.2f\n", name, e ? cudaGetErrorString(e) : "ok", avg, per_part / avg,
           avg / per_part);
    cudaFree(out), cudaFree(in), cudaFree(cyc);
}

int main() {
    run<0>("FMUL2 rr", 1), run<1>("FMUL2 r,imm", 1), run<2>("FMUL2 r,scalar", 1), run<7>("FADD2 rr", 1);
    run<3>("FFMA2 rrr", 1), run<4>("FFMA2 r,imm,r", 1), run<5>("FFMA2 r,scalar,r", 1), run<6>("FFMA2 r,scalar,imm", 1);
    run<10>("FMUL rr", 1), run<11>("FMUL r,imm", 1), run<14>("FADD rr", 1), run<12>("FFMA rrr", 1), run<13>("FFMA r,imm,r", 1);
    run<20>("IADD rr", 1), run<21>("IADD r,imm", 1), run<22>("LOP3 rrr", 1), run<23>("I2FP r", 1), run<24>("PRMT rr", 1), run<25>("IMAD rrr", 1);
    run<30>("FMUL2 rr + IADD rr", 2), run<31>("FFMA2 rrr + IADD rr", 2), run<32>("FMUL2 r,imm + IADD rr", 2), run<33>("FMUL2 rr + FMUL rr", 2);
    run<34>("FMUL2 rr + I2FP", 2), run<35>("FFMA2 rrr + 2 IADD r,imm", 3), run<36>("FMUL2 rr + 2 IADD rr", 3), run<37>("FMUL2 rr + IMAD rrr", 2);
    run<38>("FMUL rr + IADD rr", 2), run<39>("FFMA rrr + IADD rr", 2);
    return 0;
}
