// kernels_gather.cuh -- svb_mix_gather: the fused compositor with the texture unit doing the taps.
//
// The bilinear footprint of a separable YUV layer is fetched by ONE tex2Dgather per output sample: the four texels
// (i0, j0) .. (i0+1, j0+1) that the reference's OpenCL sampler reads (kernels.cl.swift:61: normalised coordinates,
// CLAMP_TO_EDGE, LINEAR), each index clamped to the plane on its own by the unit's clamp addressing -- exactly the sampler's
// rule -- and converted UNORM8 -> float by the unit.  tools/tex_probe.cu established on the B200 that this conversion equals
// c / 255.0f correctly rounded for all 256 codes and which component holds which texel:
//     x = (i0, j0+1)   y = (i0+1, j0+1)   z = (i0+1, j0)   w = (i0, j0)
// The filter itself stays in fp32 with the reference's weights and summation order (the unit's own bilinear filter has 8-bit
// weights and is not used), so the bytes are the same as the TMA-staged kernel's; per sample the gather replaces four
// shared-memory byte reads, four conversions and four UNORM8 reads, and nothing is staged: no shared
// memory boxes, no mbarriers, no CTA barrier -- a warp owns a 128 x 4 strip of a tile and never waits for another warp.
//
// Work units = (tile, strip) claimed by warps from the launch's counter; the tile's plan (SvbTilePlan, written by the
// pre-pass) is read in place through the read-only path, all lanes the same address.
#pragma once
#include "kernels_tiled.cuh"

#ifndef SVB_GATHER_CHROMA_LDG
#define SVB_GATHER_CHROMA_LDG 0  // 1: chroma taps as plain byte loads instead of gathers -- measured slower (0.470 vs 0.394 ms), kept for A/B
#endif

namespace svb {

// One separable YUV layer over this warp's 128 x 4 strip.  MODE as in fast_layer (0: inside the picture, opacity 1; 1: inside,
// 0 <= opacity <= 1; 2: lean edge; 3: anything).  N12: chroma is one two-channel plane (texU), else two planes.
template <int MODE, bool N12>
__device__ __forceinline__ void gather_layer(const SvbLayerDesc* __restrict__ L, unsigned long long texY, unsigned long long texU, unsigned long long texV, const uint32_t* __restrict__ colblk,
                                             const uint32_t* __restrict__ rowblk, int lane, int strip, float alpha, float onef, const FillTerms& ft,
                                             float2 (&Yi)[4][2], float2 (&Ui)[2], float2 (&Vi)[2]) {
    constexpr bool UNIT = MODE == 0, EDGE = MODE == 2, GEN = MODE == 3, PK = true;
    const float2 AL = splat(alpha), NAL = splat(sub(1.f, alpha)), ONE = splat(onef);
    float xg[4];
    float A[4], NA[4];  // a and 1 - a of the column
    int okc[4] = {7, 7, 7, 7};
    float2 M[2] = {splat(1.f), splat(1.f)};
#pragma unroll
    for (int p = 0; p < 2; ++p) {  // pair p = luma columns 64p + 2*lane, +1 of the tile
        const float2 a = __ldg(reinterpret_cast<const float2*>(colblk + 64 * p + 2 * lane));
        const uint2 e = __ldg(reinterpret_cast<const uint2*>(colblk + SVB_TILE_W + 64 * p + 2 * lane));
        xg[2 * p] = gather_coord(e.x), xg[2 * p + 1] = gather_coord(e.y);
        A[2 * p] = a.x, A[2 * p + 1] = a.y, NA[2 * p] = sub(1.f, a.x), NA[2 * p + 1] = sub(1.f, a.y);
        if (GEN) okc[2 * p] = (int)(e.x >> 17), okc[2 * p + 1] = (int)(e.y >> 17);
        if (EDGE) M[p] = make_float2((e.x >> 17) == 7u ? 1.f : 0.f, (e.y >> 17) == 7u ? 1.f : 0.f);
    }
    // blend -> UNORM8 write -> the next layer's UNORM8 read stays an integer-valued float (fast_layer's settle)
    auto settle = [&](float2 cur_i, float2 v, float2 fillc, float lo, int ok0, int ok1, float2 m) -> float2 {
        if (UNIT) return quant2<false, PK>(v, ONE);
        const float2 cur = unorm2<PK>(cur_i);
        const float2 qi = quant2<GEN, PK>(add2<PK>(mul2<PK>(cur, NAL), mul2<PK>(v, AL), ONE), ONE);
        if (EDGE) return fma2<PK>(m, fma2<PK>(cur_i, splat(-1.f), qi), cur_i);
        if (!GEN) return qi;
        float2 rf = add2<PK>(mul2<PK>(cur, ft.naf), mul2<PK>(fillc, ft.af), ONE);
        rf.x = fminf(fmaxf(rf.x, lo), 1.f), rf.y = fminf(fmaxf(rf.y, lo), 1.f);
        const float2 qf = quant2<true, PK>(rf, ONE);
        float2 out;
        out.x = ok0 == 7 ? qi.x : ((ok0 & 1) ? qf.x : cur_i.x);
        out.y = ok1 == 7 ? qi.y : ((ok1 & 1) ? qf.y : cur_i.y);
        return out;
    };
    auto filter = [](const float4 g, float na, float a, float b, float nb) -> float { return gather_filter(g, na, a, b, nb); };
#pragma unroll
    for (int r = 0; r < SVB_GATHER_ROWS; ++r) {
        const uint2 ry = __ldg(reinterpret_cast<const uint2*>(rowblk + 2 * (SVB_GATHER_ROWS * strip + r)));
        const int okr = (EDGE || GEN) ? (int)(ry.y >> 17) : 7;
        if (EDGE && okr != 7) continue;  // a row outside the picture (warp-uniform): untouched
        const float yg = gather_coord(ry.y), b = __uint_as_float(ry.x), nb = sub(1.f, b);
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = filter(tex2Dgather<float4>((cudaTextureObject_t)texY, xg[c], yg, 0), NA[c], A[c], b, nb);
#pragma unroll
        for (int p = 0; p < 2; ++p) Yi[r][p] = settle(Yi[r][p], make_float2(v[2 * p], v[2 * p + 1]), ft.fy, 0.f, okc[2 * p] & okr, okc[2 * p + 1] & okr, M[p]);
    }
    // chroma texel columns lane and 32 + lane of the tile, chroma rows (ROWS/2)*strip ...
    const uint32_t pc0 = __ldg(colblk + 2 * SVB_TILE_W + SVB_TILE_W / 2 + lane), pc1 = __ldg(colblk + 2 * SVB_TILE_W + SVB_TILE_W / 2 + 32 + lane);
    const float ac0 = __uint_as_float(__ldg(colblk + 2 * SVB_TILE_W + lane)), ac1 = __uint_as_float(__ldg(colblk + 2 * SVB_TILE_W + 32 + lane));
    const float xc0 = gather_coord(pc0), xc1 = gather_coord(pc1);
    const float nac0 = sub(1.f, ac0), nac1 = sub(1.f, ac1);
    const int okc0 = GEN ? (int)(pc0 >> 17) : 7, okc1 = GEN ? (int)(pc1 >> 17) : 7;
    const float2 MC = EDGE ? make_float2((pc0 >> 17) == 7u ? 1.f : 0.f, (pc1 >> 17) == 7u ? 1.f : 0.f) : splat(1.f);
#pragma unroll
    for (int k = 0; k < SVB_GATHER_ROWS / 2; ++k) {
        const uint2 rc = __ldg(reinterpret_cast<const uint2*>(rowblk + 2 * SVB_TILE_H + 2 * ((SVB_GATHER_ROWS / 2) * strip + k)));
        const int okq = (EDGE || GEN) ? (int)(rc.y >> 17) : 7;
        if (EDGE && okq != 7) continue;
        const float yc = gather_coord(rc.y), b = __uint_as_float(rc.x), nb = sub(1.f, b);
#if SVB_GATHER_CHROMA_LDG
        // chroma taps as plain byte loads through L1 (the texture unit is the busiest unit of this kernel: a third of its fetches
        // moved to the load path).  The clamped indices of the table entries address the plane directly.
        float u0, u1, v0, v1;
        {
            const unsigned j0 = rc.y & 0xffffu, j1 = j0 + ((rc.y >> 16) & 1u);
            const unsigned i00 = pc0 & 0xffffu, i01 = i00 + ((pc0 >> 16) & 1u), i10 = pc1 & 0xffffu, i11 = i10 + ((pc1 >> 16) & 1u);
            if (N12) {
                const uint8_t* __restrict__ r0 = (const uint8_t*)L->plane[1] + (size_t)j0 * L->stride[1];
                const uint8_t* __restrict__ r1 = (const uint8_t*)L->plane[1] + (size_t)j1 * L->stride[1];
                const unsigned a00 = ldg_u16(r0 + 2 * i00), a10 = ldg_u16(r0 + 2 * i01), a01 = ldg_u16(r1 + 2 * i00), a11 = ldg_u16(r1 + 2 * i01);
                const unsigned b00 = ldg_u16(r0 + 2 * i10), b10 = ldg_u16(r0 + 2 * i11), b01 = ldg_u16(r1 + 2 * i10), b11 = ldg_u16(r1 + 2 * i11);
                u0 = filter(make_float4(unorm(opaque(a01 & 0xffu)), unorm(opaque(a11 & 0xffu)), unorm(opaque(a10 & 0xffu)), unorm(opaque(a00 & 0xffu))), nac0, ac0, b, nb);
                v0 = filter(make_float4(unorm(opaque(a01 >> 8)), unorm(opaque(a11 >> 8)), unorm(opaque(a10 >> 8)), unorm(opaque(a00 >> 8))), nac0, ac0, b, nb);
                u1 = filter(make_float4(unorm(opaque(b01 & 0xffu)), unorm(opaque(b11 & 0xffu)), unorm(opaque(b10 & 0xffu)), unorm(opaque(b00 & 0xffu))), nac1, ac1, b, nb);
                v1 = filter(make_float4(unorm(opaque(b01 >> 8)), unorm(opaque(b11 >> 8)), unorm(opaque(b10 >> 8)), unorm(opaque(b00 >> 8))), nac1, ac1, b, nb);
            } else {
                const uint8_t* __restrict__ ur0 = (const uint8_t*)L->plane[1] + (size_t)j0 * L->stride[1];
                const uint8_t* __restrict__ ur1 = (const uint8_t*)L->plane[1] + (size_t)j1 * L->stride[1];
                const uint8_t* __restrict__ vr0 = (const uint8_t*)L->plane[2] + (size_t)j0 * L->stride[2];
                const uint8_t* __restrict__ vr1 = (const uint8_t*)L->plane[2] + (size_t)j1 * L->stride[2];
                u0 = filter(make_float4(unorm(ldg_u8(ur1 + i00)), unorm(ldg_u8(ur1 + i01)), unorm(ldg_u8(ur0 + i01)), unorm(ldg_u8(ur0 + i00))), nac0, ac0, b, nb);
                v0 = filter(make_float4(unorm(ldg_u8(vr1 + i00)), unorm(ldg_u8(vr1 + i01)), unorm(ldg_u8(vr0 + i01)), unorm(ldg_u8(vr0 + i00))), nac0, ac0, b, nb);
                u1 = filter(make_float4(unorm(ldg_u8(ur1 + i10)), unorm(ldg_u8(ur1 + i11)), unorm(ldg_u8(ur0 + i11)), unorm(ldg_u8(ur0 + i10))), nac1, ac1, b, nb);
                v1 = filter(make_float4(unorm(ldg_u8(vr1 + i10)), unorm(ldg_u8(vr1 + i11)), unorm(ldg_u8(vr0 + i11)), unorm(ldg_u8(vr0 + i10))), nac1, ac1, b, nb);
            }
        }
#else
        const cudaTextureObject_t tu = (cudaTextureObject_t)texU, tv = (cudaTextureObject_t)(N12 ? texU : texV);
        const float u0 = filter(tex2Dgather<float4>(tu, xc0, yc, 0), nac0, ac0, b, nb), u1 = filter(tex2Dgather<float4>(tu, xc1, yc, 0), nac1, ac1, b, nb);
        const float v0 = filter(tex2Dgather<float4>(tv, xc0, yc, N12 ? 1 : 0), nac0, ac0, b, nb), v1 = filter(tex2Dgather<float4>(tv, xc1, yc, N12 ? 1 : 0), nac1, ac1, b, nb);
#endif
        Ui[k] = settle(Ui[k], make_float2(u0, u1), ft.fu, -1.f, okc0 & okq, okc1 & okq, MC);
        Vi[k] = settle(Vi[k], make_float2(v0, v1), ft.fv, -1.f, okc0 & okq, okc1 & okq, MC);
    }
}

}  // namespace svb

#ifndef SVB_GATHER_MIN_CTAS
#define SVB_GATHER_MIN_CTAS 4  // 64 registers, 32 warps per SM: 0.377 ms per launch against 0.394 (3 CTAs, 80 registers) and 0.454 (2 CTAs, 126)
#endif

extern "C" __global__ void __launch_bounds__(SVB_TILED_THREADS, SVB_GATHER_MIN_CTAS)
    svb_mix_gather(const SvbFrameDesc* __restrict__ frames, const SvbTilePlan* __restrict__ plans, int total_units, float one, int* __restrict__ unit_counter) {
    using namespace svb;
    const int lane = threadIdx.x & 31;
    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(unit_counter, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= total_units) return;
        const int tile = u / SVB_GATHER_STRIPS, strip = u % SVB_GATHER_STRIPS;
        const int4* __restrict__ hdr = reinterpret_cast<const int4*>(plans + tile);
        const int4(*plan)[5] = reinterpret_cast<const int4(*)[5]>(plans + tile) + 1;  // plan[i] = the i-th layer that touches this tile
        const int4 h0 = __ldg(hdr), h1 = __ldg(hdr + 1);
        const int nact = h0.x, x0 = h0.z, y0 = h0.w, W = h1.x, H = h1.y;
        const SvbFrameDesc* __restrict__ F = frames + h0.y;
        const int xt = x0 + 2 * lane, yt = y0 + SVB_GATHER_ROWS * strip;  // this thread's columns xt, xt+1, xt+64, xt+65 x rows yt..yt+ROWS-1
        if (yt >= H) continue;                               // the whole strip lies below the frame
        const bool live = xt < W;                            // W and H even are planner preconditions
        const bool live1 = xt + SVB_TILE_W / 2 < W;
        const int4 h2 = __ldg(hdr + 2), h3 = __ldg(hdr + 3), h4 = __ldg(hdr + 4);
        uint8_t* const oY = (uint8_t*)(((unsigned long long)(unsigned)h2.y << 32) | (unsigned)h2.x);
        uint8_t* const oU = (uint8_t*)(((unsigned long long)(unsigned)h2.w << 32) | (unsigned)h2.z);
        uint8_t* const oV = (uint8_t*)(((unsigned long long)(unsigned)h3.y << 32) | (unsigned)h3.x);
        const int sY = h3.z, sU = h3.w, sV = h4.x;

        float2 Yi[4][2], Ui[2], Vi[2];
#pragma unroll
        for (int r = 0; r < 4; ++r) Yi[r][0] = Yi[r][1] = splat(0.f);  // img_clear_*: Y = 0, chroma = 0.5 -> 128 (rows beyond SVB_GATHER_ROWS stay unused)
        Ui[0] = Ui[1] = Vi[0] = Vi[1] = splat(128.f);
        if ((h1.w & SVB_FRAME_LOAD_CUR) && live) {
#pragma unroll
            for (int r = 0; r < SVB_GATHER_ROWS; ++r)
                if (yt + r < H) {
                    const uint8_t* row = oY + (size_t)(yt + r) * sY + xt;
                    const unsigned w0 = *(const unsigned short*)row, w1 = live1 ? *(const unsigned short*)(row + SVB_TILE_W / 2) : 0u;
                    Yi[r][0] = bytes2(opaque(w0 & 0xff), opaque(w0 >> 8)), Yi[r][1] = bytes2(opaque(w1 & 0xff), opaque(w1 >> 8));
                }
#pragma unroll
            for (int k = 0; k < SVB_GATHER_ROWS / 2; ++k)
                if (yt + 2 * k < H) {
                    if (h1.z == SVB_NV12) {
                        const uint8_t* row = oU + (size_t)((yt >> 1) + k) * sU + xt;
                        const unsigned w0 = *(const unsigned short*)row, w1 = live1 ? *(const unsigned short*)(row + SVB_TILE_W / 2) : 0x8080u;
                        Ui[k] = bytes2(opaque(w0 & 0xff), opaque(w1 & 0xff)), Vi[k] = bytes2(opaque(w0 >> 8), opaque(w1 >> 8));
                    } else {
                        const uint8_t* ru = oU + (size_t)((yt >> 1) + k) * sU + (xt >> 1);
                        const uint8_t* rv = oV + (size_t)((yt >> 1) + k) * sV + (xt >> 1);
                        Ui[k] = bytes2(opaque(ru[0]), opaque(live1 ? ru[SVB_TILE_W / 4] : 128u)), Vi[k] = bytes2(opaque(rv[0]), opaque(live1 ? rv[SVB_TILE_W / 4] : 128u));
                    }
                }
        }

        for (int i = 0; i < nact; ++i) {
            const int4 p0 = __ldg(&plan[i][0]);
            const int mode = p0.x & 0xff;
            const SvbLayerDesc* __restrict__ L = &F->layers[p0.x >> 8];
            if (mode >= PLAN_STAGED) {
                const int4 p1 = __ldg(&plan[i][1]), p3 = __ldg(&plan[i][3]), p4 = __ldg(&plan[i][4]);
                const int fmt = p1.y >> 8, lflags = p1.y & 0xff;
                const float alpha = __int_as_float(p1.w);
                const uint32_t* __restrict__ colblk = (const uint32_t*)(((unsigned long long)(unsigned)p3.w << 32) | (unsigned)p3.z);
                const uint32_t* __restrict__ rowblk = (const uint32_t*)(((unsigned long long)(unsigned)p4.y << 32) | (unsigned)p4.x);
                const unsigned long long tY = L->tex[0], tU = L->tex[1], tV = L->tex[2];
                FillTerms ft;
                if (mode == PLAN_STAGED_EDGE || !(lflags & SVB_LAYER_OPACITY_01)) {
                    const float4 fc = ldrow(L->u.fillColor, 0);
                    const float3 fill = rgb2yuv(fc.x, fc.y, fc.z);
                    const float af = mul(alpha, fc.w);
                    ft.fy = splat(fill.x), ft.fu = splat(fill.y), ft.fv = splat(fill.z), ft.af = splat(af), ft.naf = splat(sub(1.f, af));
                    // a sample inside the border rectangle but outside the picture anywhere in this warp's strip? (fast_layer's dispatch)
                    auto odd = [](uint32_t p) { const unsigned ok = p >> 17; return (ok & 1u) != 0u && ok != 7u; };
                    bool mixed = odd(__ldg(colblk + SVB_TILE_W + 2 * lane)) || odd(__ldg(colblk + SVB_TILE_W + 2 * lane + 1)) || odd(__ldg(colblk + SVB_TILE_W + 64 + 2 * lane)) ||
                                 odd(__ldg(colblk + SVB_TILE_W + 65 + 2 * lane)) || odd(__ldg(colblk + 2 * SVB_TILE_W + SVB_TILE_W / 2 + lane)) ||
                                 odd(__ldg(colblk + 2 * SVB_TILE_W + SVB_TILE_W / 2 + 32 + lane));
#pragma unroll
                    for (int r = 0; r < SVB_GATHER_ROWS; ++r) mixed = mixed || odd(__ldg(rowblk + 2 * (SVB_GATHER_ROWS * strip + r) + 1));
#pragma unroll
                    for (int k = 0; k < SVB_GATHER_ROWS / 2; ++k) mixed = mixed || odd(__ldg(rowblk + 2 * SVB_TILE_H + 2 * ((SVB_GATHER_ROWS / 2) * strip + k) + 1));
                    const bool lean = (lflags & SVB_LAYER_OPACITY_01) && !__any_sync(0xffffffffu, mixed);
                    // (the V component of an NV12 chroma plane is picked by the instruction, so the format is a template argument here too)
                    if (lean) {
                        if (fmt == SVB_NV12) gather_layer<2, true>(L, tY, tU, tV, colblk, rowblk, lane, strip, alpha, one, ft, Yi, Ui, Vi);
                        else gather_layer<2, false>(L, tY, tU, tV, colblk, rowblk, lane, strip, alpha, one, ft, Yi, Ui, Vi);
                    } else {
                        if (fmt == SVB_NV12) gather_layer<3, true>(L, tY, tU, tV, colblk, rowblk, lane, strip, alpha, one, ft, Yi, Ui, Vi);
                        else gather_layer<3, false>(L, tY, tU, tV, colblk, rowblk, lane, strip, alpha, one, ft, Yi, Ui, Vi);
                    }
                } else if (lflags & SVB_LAYER_UNIT_OPACITY) {
                    if (fmt == SVB_NV12) gather_layer<0, true>(L, tY, tU, tV, colblk, rowblk, lane, strip, alpha, one, ft, Yi, Ui, Vi);
                    else gather_layer<0, false>(L, tY, tU, tV, colblk, rowblk, lane, strip, alpha, one, ft, Yi, Ui, Vi);
                } else {
                    if (fmt == SVB_NV12) gather_layer<1, true>(L, tY, tU, tV, colblk, rowblk, lane, strip, alpha, one, ft, Yi, Ui, Vi);
                    else gather_layer<1, false>(L, tY, tU, tV, colblk, rowblk, lane, strip, alpha, one, ft, Yi, Ui, Vi);
                }
            } else if (live) {  // PLAN_GENERIC / PLAN_TABLE_RGBA: out of line
                float st[24];
                state_to_array(Yi, Ui, Vi, st);
                if (mode == PLAN_TABLE_RGBA) {
                    const int4 p3 = __ldg(&plan[i][3]), p4 = __ldg(&plan[i][4]);
                    rgba_table_layer(L, (const uint32_t*)(((unsigned long long)(unsigned)p3.w << 32) | (unsigned)p3.z),
                                     (const uint32_t*)(((unsigned long long)(unsigned)p4.y << 32) | (unsigned)p4.x), xt, yt, W, H, st, strip, SVB_GATHER_ROWS);
                } else {
                    generic_layer(L, xt, yt, (float)W, (float)H, W, H, st, SVB_GATHER_ROWS);
                }
                array_to_state(st, Yi, Ui, Vi);
            }
        }

        if (live) {
#pragma unroll
            for (int r = 0; r < SVB_GATHER_ROWS; ++r)
                if (yt + r < H) {
                    uint8_t* row = oY + (size_t)(yt + r) * sY + xt;
                    *(unsigned short*)row = pack2(Yi[r][0].x, Yi[r][0].y);
                    if (live1) *(unsigned short*)(row + SVB_TILE_W / 2) = pack2(Yi[r][1].x, Yi[r][1].y);
                }
#pragma unroll
            for (int k = 0; k < SVB_GATHER_ROWS / 2; ++k)
                if (yt + 2 * k < H) {
                    if (h1.z == SVB_NV12) {
                        uint8_t* row = oU + (size_t)((yt >> 1) + k) * sU + xt;
                        *(unsigned short*)row = pack2(Ui[k].x, Vi[k].x);
                        if (live1) *(unsigned short*)(row + SVB_TILE_W / 2) = pack2(Ui[k].y, Vi[k].y);
                    } else {
                        uint8_t* ru = oU + (size_t)((yt >> 1) + k) * sU + (xt >> 1);
                        uint8_t* rv = oV + (size_t)((yt >> 1) + k) * sV + (xt >> 1);
                        ru[0] = (uint8_t)__float2uint_rn(Ui[k].x), rv[0] = (uint8_t)__float2uint_rn(Vi[k].x);
                        if (live1) ru[SVB_TILE_W / 4] = (uint8_t)__float2uint_rn(Ui[k].y), rv[SVB_TILE_W / 4] = (uint8_t)__float2uint_rn(Vi[k].y);
                    }
                }
        }
    }
}
