// tma_probe.cu -- which ways of handing a host-encoded CUtensorMap to cp.async.bulk.tensor work on this box?
//   variant 0: __grid_constant__ kernel parameter   1: global memory, no fence   2: global memory + tensormap acquire fence
//   box/elem/extents given on the command line.   build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int V>
__global__ void probe(const __grid_constant__ CUtensorMap pm, const CUtensorMap* gm, unsigned char* out, int bytes, int x, int y) {
    __shared__ alignas(128) unsigned char box[32768];
    __shared__ alignas(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const void* tm = V == 0 ? (const void*)&pm : (const void*)gm;
        if (V == 2) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(box)),
                     "l"(tm), "r"(x), "r"(y), "r"(s32(&bar))
                     : "memory");
    }
    unsigned ok = 0;
    for (int spin = 0; !ok && spin < (1 << 22); ++spin)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(s32(&bar)) : "memory");
    if (!ok) { if (threadIdx.x == 0) printf("TIMEOUT\n"); return; }
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = box[i];
}

#define CK(x) do { auto r = (x); if (r != 0) { printf("FAIL %s -> %d\n", #x, (int)r); return 1; } } while (0)

int main(int argc, char** argv) {
    int variant = atoi(argv[1]), elem = atoi(argv[2]), W = atoi(argv[3]), H = atoi(argv[4]), stride = atoi(argv[5]), bw = atoi(argv[6]), bh = atoi(argv[7]);
    int x = argc > 8 ? atoi(argv[8]) : 0, y = argc > 9 ? atoi(argv[9]) : 0;
    CK(cudaSetDevice(0));
    unsigned char* img;
    CK(cudaMalloc(&img, (size_t)stride * H));
    std::vector<unsigned char> h((size_t)stride * H);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(i * 7 + (i >> 8));
    CK(cudaMemcpy(img, h.data(), h.size(), cudaMemcpyHostToDevice));
    alignas(64) CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H}, gstr[1] = {(cuuint64_t)stride};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
    CK(cuTensorMapEncodeTiled(&tm, elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, img, gdim, gstr, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    CUtensorMap* gm;
    CK(cudaMalloc(&gm, 4096));
    CK(cudaMemcpy(gm, &tm, 128, cudaMemcpyHostToDevice));
    int bytes = bw * bh * elem;
    unsigned char* out;
    CK(cudaMalloc(&out, bytes));
    CK(cudaMemset(out, 0xEE, bytes));
    if (variant == 0) probe<0><<<1, 128>>>(tm, gm, out, bytes, x, y);
    if (variant == 1) probe<1><<<1, 128>>>(tm, gm, out, bytes, x, y);
    if (variant == 2) probe<2><<<1, 128>>>(tm, gm, out, bytes, x, y);
    cudaError_t e = cudaDeviceSynchronize();
    if (e) { printf("variant %d elem %d %dx%d stride %d box %dx%d at (%d,%d): ERROR %s\n", variant, elem, W, H, stride, bw, bh, x, y, cudaGetErrorString(e)); return 2; }
    std::vector<unsigned char> o(bytes);
    CK(cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int j = 0; j < bh; ++j)
        for (int i = 0; i < bw * elem; ++i) {
            int gx = x * elem + i, gy = y + j;
            unsigned char want = (gx < W * elem && gy < H) ? h[(size_t)gy * stride + gx] : 0;
            if (o[j * bw * elem + i] != want) ++bad;
        }
    printf("variant %d elem %d %dx%d stride %d box %dx%d at (%d,%d): ok, %d mismatches\n", variant, elem, W, H, stride, bw, bh, x, y, bad);
    return 0;
}
