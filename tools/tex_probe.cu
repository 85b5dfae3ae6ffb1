// tex_probe.cu -- does the texture unit's UNORM8 -> float conversion equal c / 255.0f correctly rounded, and which texels does
// tex2Dgather return for a coordinate on a texel corner, clamped at the edges?  (Decides whether a gather-based compositor can
// be bit-exact.)   nvcc -arch=sm_100a -o tex_probe tex_probe.cu && ./tex_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_unorm(cudaTextureObject_t t, unsigned* bad, float* out) {
    const int c = threadIdx.x;
    // texel c of row 0: gather around the corner between texels (c, 0) and (c+1, 1): x = c + 1, y = 1
    const float4 g = tex2Dgather<float4>(t, (float)c + 1.0f, 1.0f, 0);
    const float want = __fdiv_rn((float)c, 255.0f);
    out[c] = g.w;
    if (__float_as_uint(g.w) != __float_as_uint(want)) atomicAdd(bad, 1u);
}
__global__ void k_foot(cudaTextureObject_t t, float x, float y, float4* out) { *out = tex2Dgather<float4>(t, x, y, 0); }
__global__ void k_foot2(cudaTextureObject_t t, float x, float y, float4* out) { out[0] = tex2Dgather<float4>(t, x, y, 0); out[1] = tex2Dgather<float4>(t, x, y, 1); }

int main() {
    const int W = 256, H = 4;
    uint8_t host[H][W];
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) host[y][x] = (uint8_t)((x + 64 * y) & 255);
    uint8_t* dev; size_t pitch;
    cudaMallocPitch(&dev, &pitch, W, H);
    cudaMemcpy2D(dev, pitch, host, W, W, H, cudaMemcpyHostToDevice);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypePitch2D; rd.res.pitch2D.devPtr = dev; rd.res.pitch2D.desc = cudaCreateChannelDesc<unsigned char>();
    rd.res.pitch2D.width = W; rd.res.pitch2D.height = H; rd.res.pitch2D.pitchInBytes = pitch;
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
    cudaTextureObject_t tex; cudaError_t e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    printf("create: %s (pitch %zu)\n", cudaGetErrorString(e), pitch);
    unsigned* bad; float* out; float4* f4;
    cudaMallocManaged(&bad, 4); cudaMallocManaged(&out, 256 * 4); cudaMallocManaged(&f4, 64);
    *bad = 0;
    k_unorm<<<1, 256>>>(tex, bad, out);
    e = cudaDeviceSynchronize();
    printf("unorm: %s, %u of 256 differ from __fdiv_rn(c, 255)\n", cudaGetErrorString(e), *bad);
    for (int c : {1, 3, 127, 254}) printf("  c=%d tex=%a div=%a\n", c, out[c], (float)c / 255.0f);
    // footprint order: texel value = x + 64 y; ask around corner (10.0+1, 1.0+1) -> texels (10..11, 1..2)
    k_foot<<<1, 1>>>(tex, 11.0f, 2.0f, f4); cudaDeviceSynchronize();
    printf("gather at (11,2): x=%g y=%g z=%g w=%g   (texels (10,1)=74 (11,1)=75 (10,2)=138 (11,2)=139)\n", f4->x * 255, f4->y * 255, f4->z * 255, f4->w * 255);
    // clamping: left/top edge (x = 0 -> texels (-1,0) clamp), right/bottom edge
    k_foot<<<1, 1>>>(tex, 0.0f, 0.0f, f4); cudaDeviceSynchronize();
    printf("gather at (0,0): %g %g %g %g   (all texel (0,0)=0 expected)\n", f4->x * 255, f4->y * 255, f4->z * 255, f4->w * 255);
    k_foot<<<1, 1>>>(tex, 256.0f, 4.0f, f4); cudaDeviceSynchronize();
    printf("gather at (256,4): %g %g %g %g   (all texel (255,3)=191 expected)\n", f4->x * 255, f4->y * 255, f4->z * 255, f4->w * 255);
    k_foot<<<1, 1>>>(tex, 0.0f, 2.0f, f4); cudaDeviceSynchronize();
    printf("gather at (0,2): %g %g %g %g   (texels (0,1)=64 twice, (0,2)=128 twice)\n", f4->x * 255, f4->y * 255, f4->z * 255, f4->w * 255);
    // two-channel texture over the same bytes (NV12 chroma): W/2 texels of (U,V)
    cudaResourceDesc r2 = rd; r2.res.pitch2D.desc = cudaCreateChannelDesc<uchar2>(); r2.res.pitch2D.width = W / 2;
    cudaTextureObject_t tex2; e = cudaCreateTextureObject(&tex2, &r2, &td, nullptr);
    k_foot2<<<1, 1>>>(tex2, 6.0f, 1.0f, f4); cudaDeviceSynchronize();
    printf("uchar2 create %s; gather comp0 at (6,1): %g %g %g %g  comp1: %g %g %g %g   (texels 5,6 of row 0: U=10,12 V=11,13; row 1 +64)\n", cudaGetErrorString(e), f4[0].x * 255, f4[0].y * 255,
           f4[0].z * 255, f4[0].w * 255, f4[1].x * 255, f4[1].y * 255, f4[1].z * 255, f4[1].w * 255);
    return 0;
}
