// kernels_tiled.cuh -- svb_mix_tiled: the fused compositor's fast path.
//
// A CTA owns a 128x32 luma tile of one output frame (tiles of every frame of the batch are dealt round-robin
// to a grid sized to the SM count).  The running picture of the tile lives in registers as packed bytes
// (a thread owns 4 columns x 4 rows of luma and the 2x2 chroma texels under them) and is re-quantised to
// 8 bits after every layer, so the bytes equal the reference's clear-then-fold over an 8-bit target
// (mix.video.swift:113-125).  Layers whose rectangle misses the tile are skipped.
//
// Separable layers (no rotation: x outputs depend on x only, y outputs on y only -- the planner proves it
// from the uniforms) take the table path: 128 threads evaluate the reference's per-pixel coordinate chain
// (kernels.cl.swift:70-78 + the OpenCL 1.2 linear sampler) once per COLUMN and 32 once per ROW, bit-exactly,
// into shared memory; the source footprint of the tile is staged by one TMA 2-D tensor copy per plane
// (cp.async.bulk.tensor, mbarrier completion) and every pixel then costs four byte taps, the bilinear
// sum and the blend.  Everything else (rotated layers, BGRA/RGBA sources, tiles straddling a layer edge)
// runs the generic per-pixel evaluator of svb_device.cuh for that layer on that tile.
#pragma once
#include "svb_device.cuh"

#define SVB_TILED_THREADS 256

namespace svb {

struct __align__(16) Ent {
    float a;  // fractional weight of the i1 tap
    int i0, i1;  // clamped source indices (plane coordinates)
    int ok;      // bit0 border in [0,1], bit1 tx in [0,1], bit2 uv in [0,1]
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// A tensor copy that never lands (bad descriptor) must not hang the GPU: trap after ~seconds instead.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    for (unsigned spin = 0; !mbar_try(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}
// descriptors are rewritten by the host between launches: make the tensormap proxy re-read them
__device__ __forceinline__ void tmap_acquire(const void* tmap) {
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

// The reference's coordinate chain for one output column (axis 0) or row (axis 1) of a separable layer:
// out_uv = g/size; normpos = out_uv*2-1; tx = vecmat4(normpos, T); border = vecmat4(normpos, B);
// uv = vecmat4(tx, X)  (kernels.cl.swift:72-78), then the sampler's i0/i1/a for a plane `n` texels long.
// The cross-axis terms multiply exact zeros in a separable layer, so they are dropped; the value is identical.
struct Axis {
    float tB, tT, tU;  // border, tx, uv along this axis
};
__device__ __forceinline__ Axis axis_chain(const SvbUniforms* __restrict__ U, int axis, int g, float size) {
    const float n = sub(mul(__fdiv_rn((float)g, size), 2.f), 1.f);
    const float nx = axis == 0 ? n : 0.f, ny = axis == 0 ? 0.f : n;
    Axis r;
    r.tB = dot4(nx, ny, 0.f, 1.f, ldrow(U->borderMatrix, axis));
    const float t0 = dot4(nx, ny, 0.f, 1.f, ldrow(U->transform, 0));
    const float t1 = dot4(nx, ny, 0.f, 1.f, ldrow(U->transform, 1));
    const float t2 = dot4(nx, ny, 0.f, 1.f, ldrow(U->transform, 2));
    const float t3 = dot4(nx, ny, 0.f, 1.f, ldrow(U->transform, 3));
    r.tT = axis == 0 ? t0 : t1;
    r.tU = dot4(axis == 0 ? t0 : 0.f, axis == 0 ? 0.f : t1, t2, t3, ldrow(U->textureTx, axis));
    return r;
}
__device__ __forceinline__ Ent axis_entry(const Axis& c, int n) {
    Ent e;
    e.ok = (c.tB >= 0.f && c.tB <= 1.f ? 1 : 0) | (c.tT >= 0.f && c.tT <= 1.f ? 2 : 0) | (c.tU >= 0.f && c.tU <= 1.f ? 4 : 0);
    const float um = sub(mul(c.tU, (float)n), 0.5f);
    const float fu = floorf(um);
    e.a = sub(um, fu);
    const int i = (int)fu;
    e.i0 = min(max(i, 0), n - 1);
    e.i1 = min(max(i + 1, 0), n - 1);
    return e;
}

__device__ __forceinline__ unsigned get8(unsigned w, int k) { return (w >> (8 * k)) & 0xffu; }
__device__ __forceinline__ unsigned put8(unsigned w, int k, unsigned v) { return (w & ~(0xffu << (8 * k))) | (v << (8 * k)); }

struct TiledSmem {
    alignas(128) uint8_t boxY[SVB_BOX_Y_BYTES];
    alignas(128) uint8_t boxC[SVB_BOX_C_BYTES];
    Ent colY[SVB_TILE_W];
    Ent colC[SVB_TILE_W / 2];
    Ent rowY[SVB_TILE_H];
    Ent rowC[SVB_TILE_H / 2];
    alignas(8) uint64_t bar;
};

}  // namespace svb

extern "C" __global__ void __launch_bounds__(SVB_TILED_THREADS, 2)
    svb_mix_tiled(const SvbFrameDesc* __restrict__ frames, int nframes, int total_tiles) {
    using namespace svb;
    __shared__ TiledSmem sm;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned phase = 0;
    if (t == 0) {
        mbar_init(&sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int f = 0;
        while (f + 1 < nframes && frames[f + 1].first_tile <= tile) ++f;
        const SvbFrameDesc* __restrict__ F = frames + f;
        const int local = tile - F->first_tile;
        const int W = F->width, H = F->height;
        const int x0 = (local % F->tiles_x) * SVB_TILE_W, y0 = (local / F->tiles_x) * SVB_TILE_H;
        const int xt = x0 + 4 * lane, yt = y0 + 4 * warp;  // this thread's 4x4 block
        const bool live = xt < W && yt < H;                 // W % 4 == 0 and H even are planner preconditions
        const bool nv12 = F->format == SVB_NV12;
        const float fW = (float)W, fH = (float)H;
        uint8_t* const oY = (uint8_t*)F->out_plane[0];
        uint8_t* const oU = (uint8_t*)F->out_plane[1];
        uint8_t* const oV = (uint8_t*)F->out_plane[2];
        const int sY = F->out_stride[0], sU = F->out_stride[1], sV = F->out_stride[2];

        // running picture: Yb[r] = 4 luma bytes of row yt+r; Cb[k] = (u0,v0,u1,v1) of chroma row (yt/2)+k
        unsigned Yb[4] = {0u, 0u, 0u, 0u}, Cb[2] = {0x80808080u, 0x80808080u};  // img_clear_*: Y=0, C=0.5->128
        if ((F->flags & SVB_FRAME_LOAD_CUR) && live) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (yt + r < H) Yb[r] = *(const unsigned*)(oY + (size_t)(yt + r) * sY + xt);
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (yt + 2 * k < H) {
                    if (nv12) {
                        Cb[k] = *(const unsigned*)(oU + (size_t)((yt >> 1) + k) * sU + xt);
                    } else {
                        const uchar2 u = *(const uchar2*)(oU + (size_t)((yt >> 1) + k) * sU + (xt >> 1));
                        const uchar2 v = *(const uchar2*)(oV + (size_t)((yt >> 1) + k) * sV + (xt >> 1));
                        Cb[k] = u.x | (v.x << 8) | (u.y << 16) | (v.y << 24);
                    }
                }
        }

        for (int l = 0; l < F->nlayers; ++l) {
            const SvbLayerDesc* __restrict__ L = &F->layers[l];
            if (L->rect[0] >= x0 + SVB_TILE_W || L->rect[2] <= x0 || L->rect[1] >= y0 + SVB_TILE_H || L->rect[3] <= y0) continue;
            const SvbUniforms* __restrict__ U = &L->u;
            const int fmt = L->format, lflags = L->flags;
            const bool yuv = fmt == SVB_NV12 || fmt == SVB_Y420P;
            bool full = false;
            if ((lflags & SVB_LAYER_SEPARABLE) && yuv) {
                // ---- tables: one column / row of the reference's coordinate chain per thread ----------
                bool mine_ok = true;
                if (t < SVB_TILE_W) {
                    const Axis c = axis_chain(U, 0, x0 + t, fW);
                    const Ent e = axis_entry(c, L->width);
                    sm.colY[t] = e;
                    if ((t & 1) == 0) sm.colC[t >> 1] = axis_entry(c, L->width / 2);
                    mine_ok = e.ok == 7 || x0 + t >= W;
                } else if (t < SVB_TILE_W + SVB_TILE_H) {
                    const int r = t - SVB_TILE_W;
                    const Axis c = axis_chain(U, 1, y0 + r, fH);
                    const Ent e = axis_entry(c, L->height);
                    sm.rowY[r] = e;
                    if ((r & 1) == 0) sm.rowC[r >> 1] = axis_entry(c, L->height / 2);
                    mine_ok = e.ok == 7 || y0 + r >= H;
                }
                full = __syncthreads_and(mine_ok);  // every pixel of the tile is inside the picture
            }
            if (full) {
                // ---- source footprint of the tile (indices are monotone in x and y) ---------------------
                const int lastc = min(SVB_TILE_W, W - x0) - 1, lastr = min(SVB_TILE_H, H - y0) - 1;
                const int iy0 = min(sm.colY[0].i0, sm.colY[lastc].i0), iy1 = max(sm.colY[0].i1, sm.colY[lastc].i1);
                const int jy0 = min(sm.rowY[0].i0, sm.rowY[lastr].i0), jy1 = max(sm.rowY[0].i1, sm.rowY[lastr].i1);
                const int lc = lastc >> 1, lr = lastr >> 1;
                const int ic0 = min(sm.colC[0].i0, sm.colC[lc].i0), ic1 = max(sm.colC[0].i1, sm.colC[lc].i1);
                const int jc0 = min(sm.rowC[0].i0, sm.rowC[lr].i0), jc1 = max(sm.rowC[0].i1, sm.rowC[lr].i1);
                const bool staged = (lflags & SVB_LAYER_STAGED) && iy1 - iy0 < L->box_w && jy1 - jy0 < L->box_h &&
                                    ic1 - ic0 < L->box_cw && jc1 - jc0 < L->box_ch;
                const uint8_t *pY, *pU, *pV;  // tap(i,j) = p[j*pitch + i*step]
                int pitchY, pitchC, stepC;
                if (staged) {
                    const int cbytes = fmt == SVB_NV12 ? L->box_cw * L->box_ch * 2 : L->box_cw * L->box_ch;
                    if (t == 0) {
                        tmap_acquire(L->tmap[0]);
                        tmap_acquire(L->tmap[1]);
                        if (fmt == SVB_Y420P) tmap_acquire(L->tmap[2]);
                        mbar_expect_tx(&sm.bar, L->box_w * L->box_h + cbytes * (fmt == SVB_NV12 ? 1 : 2));
                        tma_load_2d(sm.boxY, L->tmap[0], iy0, jy0, &sm.bar);
                        tma_load_2d(sm.boxC, L->tmap[1], ic0, jc0, &sm.bar);
                        if (fmt == SVB_Y420P) tma_load_2d(sm.boxC + SVB_BOX_C_BYTES / 2, L->tmap[2], ic0, jc0, &sm.bar);
                    }
                    pitchY = L->box_w;
                    pY = sm.boxY - ((size_t)jy0 * pitchY + iy0);
                    if (fmt == SVB_NV12) {
                        pitchC = L->box_cw * 2, stepC = 2;
                        pU = sm.boxC - ((size_t)jc0 * pitchC + ic0 * 2);
                        pV = pU + 1;
                    } else {
                        pitchC = L->box_cw, stepC = 1;
                        pU = sm.boxC - ((size_t)jc0 * pitchC + ic0);
                        pV = pU + SVB_BOX_C_BYTES / 2;
                    }
                    mbar_wait(&sm.bar, phase);
                    phase ^= 1;
                } else {
                    pitchY = L->stride[0];
                    pY = (const uint8_t*)L->plane[0];
                    pitchC = L->stride[1];
                    pU = (const uint8_t*)L->plane[1];
                    if (fmt == SVB_NV12) {
                        stepC = 2, pV = pU + 1;
                    } else {
                        stepC = 1, pV = (const uint8_t*)L->plane[2];  // stride[2] == stride[1] is a planner precondition
                    }
                }
                // ---- pixels ----------------------------------------------------------------------------
                if (live) {
                    const float alpha = U->opacity, nalpha = sub(1.f, alpha);
                    Ent cy[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) cy[c] = sm.colY[4 * lane + c];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const Ent ry = sm.rowY[4 * warp + r];
                        const float b = ry.a, nb = sub(1.f, b);
                        const uint8_t* q0 = pY + (size_t)ry.i0 * pitchY;
                        const uint8_t* q1 = pY + (size_t)ry.i1 * pitchY;
                        unsigned out = 0;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float a = cy[c].a, na = sub(1.f, a);
                            const float v = add(add(add(mul(mul(na, nb), unorm(q0[cy[c].i0])), mul(mul(a, nb), unorm(q0[cy[c].i1]))),
                                                    mul(mul(na, b), unorm(q1[cy[c].i0]))),
                                                mul(mul(a, b), unorm(q1[cy[c].i1])));
                            const float res = add(mul(unorm(get8(Yb[r], c)), nalpha), mul(v, alpha));
                            out |= rte8(res) << (8 * c);
                        }
                        Yb[r] = out;
                    }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const Ent rc = sm.rowC[2 * warp + k];
                        const float b = rc.a, nb = sub(1.f, b);
                        const size_t o0 = (size_t)rc.i0 * pitchC, o1 = (size_t)rc.i1 * pitchC;
                        unsigned out = 0;
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const Ent cc = sm.colC[2 * lane + c];
                            const float a = cc.a, na = sub(1.f, a);
                            const float w00 = mul(na, nb), w10 = mul(a, nb), w01 = mul(na, b), w11 = mul(a, b);
                            const int e0 = cc.i0 * stepC, e1 = cc.i1 * stepC;
                            const float vu = add(add(add(mul(w00, unorm(pU[o0 + e0])), mul(w10, unorm(pU[o0 + e1]))), mul(w01, unorm(pU[o1 + e0]))),
                                                 mul(w11, unorm(pU[o1 + e1])));
                            const float vv = add(add(add(mul(w00, unorm(pV[o0 + e0])), mul(w10, unorm(pV[o0 + e1]))), mul(w01, unorm(pV[o1 + e0]))),
                                                 mul(w11, unorm(pV[o1 + e1])));
                            const float ru = add(mul(unorm(get8(Cb[k], 2 * c)), nalpha), mul(vu, alpha));
                            const float rv = add(mul(unorm(get8(Cb[k], 2 * c + 1)), nalpha), mul(vv, alpha));
                            out |= (rte8(ru) << (16 * c)) | (rte8(rv) << (16 * c + 8));
                        }
                        Cb[k] = out;
                    }
                }
            } else if (live) {
                // ---- generic per-pixel evaluation of this layer on this thread's 4x4 block --------------
                const Src s = layer_src(L);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const bool chroma = ((r | c) & 1) == 0;
                        const int k = r >> 1, cc = c >> 1;
                        float oy, ou, ov;
                        const float cu = chroma ? unorm(get8(Cb[k], 2 * cc)) : 0.f, cv = chroma ? unorm(get8(Cb[k], 2 * cc + 1)) : 0.f;
                        if (yt + r < H && eval_pixel(U, s, xt + c, yt + r, fW, fH, chroma, unorm(get8(Yb[r], c)), cu, cv, oy, ou, ov)) {
                            Yb[r] = put8(Yb[r], c, rte8(oy));
                            if (chroma) Cb[k] = put8(put8(Cb[k], 2 * cc, rte8(ou)), 2 * cc + 1, rte8(ov));
                        }
                    }
                }
            }
            __syncthreads();  // tables and boxes are free for the next layer
        }

        if (live) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (yt + r < H) *(unsigned*)(oY + (size_t)(yt + r) * sY + xt) = Yb[r];
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (yt + 2 * k < H) {
                    if (nv12) {
                        *(unsigned*)(oU + (size_t)((yt >> 1) + k) * sU + xt) = Cb[k];
                    } else {
                        *(uchar2*)(oU + (size_t)((yt >> 1) + k) * sU + (xt >> 1)) = make_uchar2(get8(Cb[k], 0), get8(Cb[k], 2));
                        *(uchar2*)(oV + (size_t)((yt >> 1) + k) * sV + (xt >> 1)) = make_uchar2(get8(Cb[k], 1), get8(Cb[k], 3));
                    }
                }
        }
    }
}
