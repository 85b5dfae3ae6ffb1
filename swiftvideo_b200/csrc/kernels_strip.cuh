// kernels_strip.cuh -- svb_mix_strip: the fused compositor's fast path, second design (round 2).
//
// What the profile of svb_mix_tiled said (profiles/r2_history.md): its layer bodies would saturate the issue port on their
// own, but a warp spent 56 % of its time outside them -- at the CTA barrier of every staged layer (20 %), waiting for plan /
// descriptor loads, and in per-layer and per-tile set-up code (a third of all instructions).  Hence here:
//   * a WARP is the unit of work and never waits for another warp: it owns a 64x8 unit of an output frame (claimed from a
//     counter in row-major order) and stages ITS OWN source footprint -- one TMA 2-D tensor copy per plane into a private
//     double buffer, completion on a private mbarrier; the copies of the next layer (or of the next unit's first layer) fly
//     while the current one is computed;
//   * a warp plans its own units, one ahead, lane = layer: which layers touch the unit, how (interior / edge / per-pixel), where
//     their boxes start -- two 8-byte loads per layer, because the plan is separable (svb_strip_tables leaves a record per unit
//     column and per unit row of every layer); a layer's table slices (the reference's coordinate chain, kernels.cl.swift:70-78,
//     per output column and row, in the form the inner loop consumes -- weights with their complements, row offsets already
//     multiplied by the box pitch) arrive with its boxes by two bulk copies;
//   * free-running warps do not share an instruction stream, so the hot code must fit the instruction cache of a scheduler on its
//     own: ONE loop of two output rows (~2 KB) serves luma and chroma rows, NV12 and planar sources alike;
//   * a lane owns two adjacent luma columns x 8 rows and the chroma texel column under them, and walks DOWN its rows: the
//     converted taps of a source row stay in registers and serve the next output row when it continues from there (always at
//     1:1, four rows in five at the 1.2:1 of the headline workload), so a sample costs two or three new taps, not four;
//   * the running picture of the unit lives in shared memory (3 KB per warp, integer-valued floats, re-quantised after
//     every layer exactly like the reference's 8-bit target, mix.video.swift:113-125).
// The per-sample arithmetic is fast_layer's (kernels_tiled.cuh), operation for operation: packed fp32x2, every multiply and
// add rounded on its own.
#pragma once
#include "kernels_tiled.cuh"

#ifndef SVB_STRIP_MIN_CTAS
#define SVB_STRIP_MIN_CTAS 5
#endif

namespace svb {

struct TapRow {  // the converted taps of one source row under a lane's two samples: (s0, s1) at tap i0 and at tap i1
    float2 p0, p1;
};

__device__ __forceinline__ float4 lds_f4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u4(unsigned addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds_u2(unsigned addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds_u1(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// ---- interior layer: the unit lies wholly inside the picture and no tap is clamped along x -------------------------------
//   OPAQUE: opacity == 1 (cur*(1-1) + v*1 == v: the running picture is not read); else 0 <= opacity <= 1.
//   HALF:   every weight of the unit's table blocks is exactly 1/2 (a picture shown 1:1, or scaled by an exact power of two, at
//           whole-pixel positions: the reference's sampler then averages four texels).  The four weights are 1/4 each, products by
//           1/4 are exact and scaling by a power of two commutes with rounding, so
//               ((w00 t00 + w10 t10) + w01 t01) + w11 t11  ==  ((t00 + t10) + t01 + t11) / 4      bit for bit
//           (three additions instead of four weight products, four tap products and three additions), and an opaque layer's
//           rint(v * 255) is rint(sum * 63.75).
//   colY / colC: shared-memory address of the staged box minus its origin (ix0 + jy0 * pitch), so that a tap's address is
//   colY + i0 + row offset;  cstep: bytes from a chroma texel to the next (2: NV12, 1: planar);  vofs: from a U byte to its V byte
//   tab / rows: shared-memory addresses of the unit's column block and row block;  st: shared-memory address of this lane's slot in state row 0
// Rows 0..7 are the luma rows, 8..11 the chroma rows of the unit; a row's two samples are (col 2l, col 2l+1) or (U, V).
template <bool OPAQUE, bool HALF>
__device__ __forceinline__ void strip_layer(unsigned colY, unsigned colC, unsigned cstep, unsigned vofs, unsigned tab, unsigned rows, int lane, float alpha, float onef, unsigned st) {
    constexpr bool PK = true;
    const float2 AL = splat(alpha), NAL = splat(sub(1.f, alpha)), ONE = splat(onef);
    // which rows must fetch their upper source row: the first luma and the first chroma row, and every row whose upper row is not
    // the lower row of the row before (never at 1:1, one row in five at 1.2:1).  One vote per layer; the row loop tests a bit.
    unsigned reload;
    {
        const unsigned r = (unsigned)lane < 12u ? (unsigned)lane : 0u;
        const unsigned top = lds_u1(rows + 16u * r + 8u), prev = lds_u1(rows + 16u * (r ? r - 1u : 0u) + 12u);
        reload = __ballot_sync(0xffffffffu, (unsigned)lane < 12u && (r == 0u || r == 8u || top != prev));
    }
    const uint2 aw = lds_u2(tab + 8u * lane), e = lds_u2(tab + 4u * SVB_UNIT_W + 8u * lane);
    float2 A = make_float2(__uint_as_float(aw.x), __uint_as_float(aw.y));
    float2 NA = make_float2(sub(1.f, A.x), sub(1.f, A.y));
    // tap addresses of the row at offset 0: first tap of sample 0 / sample 1, second tap of sample 0 / sample 1
    unsigned b00 = colY + (e.x & 0xffffu), b10 = colY + (e.y & 0xffffu), b01 = b00 + 1u, b11 = b10 + 1u;
    TapRow T, B;
    T.p0 = T.p1 = B.p0 = B.p1 = splat(0.f);

    auto raw = [&](unsigned (&q)[4], unsigned off) { q[0] = lds_u8(b00 + off), q[1] = lds_u8(b10 + off), q[2] = lds_u8(b01 + off), q[3] = lds_u8(b11 + off); };
    auto conv = [&](TapRow& t, const unsigned (&q)[4]) { t.p0 = unorm2<PK>(bytes2(q[0], q[1])), t.p1 = unorm2<PK>(bytes2(q[2], q[3])); };
    // one output row from its upper and lower source rows (converted taps), its table entry and the running picture's pair
    auto out_row = [&](const TapRow& top, const TapRow& bot, const uint4 w, float2 ci, unsigned sa) {
        float2 v;
        if (HALF) {
            v = add2<PK>(add2<PK>(add2<PK>(top.p0, top.p1, ONE), bot.p0, ONE), bot.p1, ONE);
        } else {
            const float2 Bf = splat(__uint_as_float(w.x)), NB = splat(__uint_as_float(w.y));
            v = bilin2<PK>(mul2<PK>(NA, NB), mul2<PK>(A, NB), mul2<PK>(NA, Bf), mul2<PK>(A, Bf), top.p0, top.p1, bot.p0, bot.p1, ONE);
        }
        float2 out;
        if (OPAQUE) {
            if (HALF) out = add2<PK>(add2<PK>(mul2<PK>(v, splat(63.75f)), splat(8388608.f), ONE), splat(-8388608.f), ONE);
            else out = quant2<false, PK>(v, ONE);
        } else {
            if (HALF) v = mul2<PK>(v, splat(0.25f));
            const float2 cur = unorm2<PK>(ci);
            out = quant2<false, PK>(add2<PK>(mul2<PK>(cur, NAL), mul2<PK>(v, AL), ONE), ONE);
        }
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sa), "f"(out.x), "f"(out.y) : "memory");
    };
    auto lds_state = [&](unsigned sa) {
        float2 ci = splat(0.f);
        if (!OPAQUE) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(ci.x), "=f"(ci.y) : "r"(sa) : "memory");
        return ci;
    };
    // Two output rows per pass.  The upper source row of a row is the lower one of the row before unless its reload bit says otherwise
    // (warp-uniform, from a vote: a plain branch -- predicated off, the fetch would still take its issue slots), so the tap registers
    // trade roles row by row and are back in place after a pass.  Every shared-memory read of the pass is issued before its arithmetic.
    auto pass = [&](unsigned ent, unsigned sa, bool fetch1, bool fetch2) {
        const uint4 w1 = lds_u4(ent), w2 = lds_u4(ent + 16u);  // b, 1-b, j0 * pitch, j1 * pitch
        unsigned q1[4], q2[4];
        if (fetch1) {
            unsigned q0[4];
            raw(q0, w1.z);
            conv(T, q0);
        }
        raw(q1, w1.w);
        raw(q2, w2.w);
        const float2 ci1 = lds_state(sa), ci2 = lds_state(sa + 256u);
        conv(B, q1);
        out_row(T, B, w1, ci1, sa);
        if (fetch2) {
            unsigned q0[4];
            raw(q0, w2.z);
            conv(B, q0);
        }
        conv(T, q2);
        out_row(B, T, w2, ci2, sa + 256u);
    };
    unsigned ent = rows, sa = st;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {  // luma rows, then chroma rows
        if (half) {  // the lane's chroma texel column, samples (U, V)
            const unsigned pc = lds_u1(tab + 4u * (2 * SVB_UNIT_W + SVB_UNIT_W / 2) + 4u * lane);
            const float ac = __uint_as_float(lds_u1(tab + 4u * (2 * SVB_UNIT_W) + 4u * lane));
            A = splat(ac), NA = splat(sub(1.f, ac));
            b00 = colC + (pc & 0xffffu) * cstep, b10 = b00 + vofs, b01 = b00 + cstep, b11 = b10 + cstep;
        }
#pragma unroll 1
        for (int n = half ? 2 : 4; n > 0; --n) {  // two rows per pass: the tap registers trade roles row by row and are back in place after a pass
            pass(ent, sa, (reload & 1u) != 0u, (reload & 2u) != 0u);
            reload >>= 2, ent += 32u, sa += 512u;
        }
    }
}

// ---- edge layer: the unit straddles the picture's (or the border rectangle's) edge: taps may be clamped along x, samples may lie
// outside.  LEAN (fast_layer's MODE 2): every sample is inside the picture or untouched, 0 <= opacity <= 1 -- rows outside are
// skipped (warp-uniform), columns outside keep their value through a 0/1 mask.  Otherwise (MODE 3): per-sample class from the ok
// bits -- picture / fill / untouched (kernels.cl.swift:77,84-85,96-105) -- and saturating stores.  One unit-layer in ten: no tap re-use.
template <bool LEAN>
__device__ __noinline__ void strip_layer_edge(unsigned colY, unsigned colC, unsigned cstep, unsigned vofs, unsigned tab, unsigned rows, int lane, float alpha, float onef, const float4 fill,
                                              float af, float2* __restrict__ sY) {
    constexpr bool PK = true, GEN = !LEAN;
    const float2 AL = splat(alpha), NAL = splat(sub(1.f, alpha)), ONE = splat(onef), AF = splat(af), NAF = splat(sub(1.f, af));
    const uint2 aw = lds_u2(tab + 8u * lane), e = lds_u2(tab + 4u * SVB_UNIT_W + 8u * lane);
    const unsigned pc = lds_u1(tab + 4u * (2 * SVB_UNIT_W + SVB_UNIT_W / 2) + 4u * lane);
    const float ac = __uint_as_float(lds_u1(tab + 4u * (2 * SVB_UNIT_W) + 4u * lane));
#pragma unroll 1
    for (int r = 0; r < 12; ++r) {
        const bool chroma = r >= 8;
        const float2 A = chroma ? splat(ac) : make_float2(__uint_as_float(aw.x), __uint_as_float(aw.y));
        const float2 NA = make_float2(sub(1.f, A.x), sub(1.f, A.y));
        unsigned b00, b10, b01, b11;
        int ok0, ok1;
        if (chroma) {
            b00 = colC + (pc & 0xffffu) * cstep, b10 = b00 + vofs, b01 = b00 + ((pc >> 16) & 1u) * cstep, b11 = b01 + vofs;
            ok0 = ok1 = (int)(pc >> 17);
        } else {
            b00 = colY + (e.x & 0xffffu), b10 = colY + (e.y & 0xffffu), b01 = b00 + ((e.x >> 16) & 1u), b11 = b10 + ((e.y >> 16) & 1u);
            ok0 = (int)(e.x >> 17), ok1 = (int)(e.y >> 17);
        }
        const uint4 w = lds_u4(rows + 16u * r);
        const int okr = (int)(lds_u1(rows + 192u + 4u * r) & 7u);
        if (LEAN && okr != 7) continue;  // a row outside the picture (warp-uniform): untouched
        const float2 Bf = splat(__uint_as_float(w.x)), NB = splat(__uint_as_float(w.y));
        const float2 t00 = unorm2<PK>(bytes2(lds_u8(b00 + w.z), lds_u8(b10 + w.z))), t10 = unorm2<PK>(bytes2(lds_u8(b01 + w.z), lds_u8(b11 + w.z)));
        const float2 t01 = unorm2<PK>(bytes2(lds_u8(b00 + w.w), lds_u8(b10 + w.w))), t11 = unorm2<PK>(bytes2(lds_u8(b01 + w.w), lds_u8(b11 + w.w)));
        const float2 v = bilin2<PK>(mul2<PK>(NA, NB), mul2<PK>(A, NB), mul2<PK>(NA, Bf), mul2<PK>(A, Bf), t00, t10, t01, t11, ONE);
        float2* __restrict__ st = sY + r * 32 + lane;
        const float2 cur_i = *st;
        const float2 cur = unorm2<PK>(cur_i);
        const float2 qi = quant2<GEN, PK>(add2<PK>(mul2<PK>(cur, NAL), mul2<PK>(v, AL), ONE), ONE);
        float2 out;
        if (LEAN) {  // columns outside keep cur_i: cur_i + m*(qi - cur_i) in integer-valued floats, every step exact
            const float2 m = make_float2((ok0 & okr) == 7 ? 1.f : 0.f, (ok1 & okr) == 7 ? 1.f : 0.f);
            out = fma2<PK>(m, fma2<PK>(cur_i, splat(-1.f), qi), cur_i);
        } else {
            const float2 fillc = chroma ? make_float2(fill.y, fill.z) : splat(fill.x);
            const float lo = chroma ? -1.f : 0.f;
            float2 rf = add2<PK>(mul2<PK>(cur, NAF), mul2<PK>(fillc, AF), ONE);
            rf.x = fminf(fmaxf(rf.x, lo), 1.f), rf.y = fminf(fmaxf(rf.y, lo), 1.f);
            const float2 qf = quant2<true, PK>(rf, ONE);
            const int k0 = ok0 & okr, k1 = ok1 & okr;
            out.x = k0 == 7 ? qi.x : ((k0 & 1) ? qf.x : cur_i.x);
            out.y = k1 == 7 ? qi.y : ((k1 & 1) ? qf.y : cur_i.y);
        }
        *st = out;
    }
}

// ---- layers that are not staged (rotation, footprint too large; BGRA / RGBA overlays): per-pixel evaluators over the lane's 2x8
// block of the running picture in shared memory (sY: the lane's first luma pair; rows are SVB_UNIT_W floats apart.  sC likewise). ------
__device__ __noinline__ void strip_generic_layer(const SvbLayerDesc* __restrict__ L, int xt, int yt, int W, int H, float* __restrict__ sY, float* __restrict__ sC) {
    const Src s = layer_src(L);
    const SvbUniforms* __restrict__ U = &L->u;
    const float fW = (float)W, fH = (float)H;
#pragma unroll 1
    for (int q = 0; q < 2 * SVB_UNIT_H; ++q) {
        const int r = q >> 1, c = q & 1;
        if (yt + r >= H) break;
        if (xt + c >= W) continue;
        const bool chroma = ((r | c) & 1) == 0;
        float* __restrict__ py = sY + r * SVB_UNIT_W + c;
        float* __restrict__ pc = sC + (r >> 1) * SVB_UNIT_W;
        float oy, ou, ov;
        if (eval_pixel(U, s, xt + c, yt + r, fW, fH, chroma, unorm_f(*py), chroma ? unorm_f(pc[0]) : 0.f, chroma ? unorm_f(pc[1]) : 0.f, oy, ou, ov)) {
            *py = quantf(oy);
            if (chroma) pc[0] = quantf(ou), pc[1] = quantf(ov);
        }
    }
}

// A separable BGRA / RGBA layer: coordinate chain from the layer's tables (read in place), the four RGBA taps from global memory, rgba_pixel's arithmetic
__device__ __noinline__ void strip_rgba_layer(const SvbLayerDesc* __restrict__ L, const uint32_t* __restrict__ colblk, const uint32_t* __restrict__ rowblk, int lane, int xt, int yt,
                                              int W, int H, float* __restrict__ sY, float* __restrict__ sC) {
    const Src s = layer_src(L);
    const float opacity = __ldg(&L->u.opacity);
    const float4 fc = ldrow(L->u.fillColor, 0);
    Ent ce[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) ce[c] = unpack_ent(__ldg(colblk + 2 * lane + c), __ldg(colblk + SVB_UNIT_W + 2 * lane + c));
#pragma unroll 1
    for (int r = 0; r < SVB_UNIT_H; ++r) {
        if (yt + r >= H) break;
        const float b = __uint_as_float(__ldg(rowblk + 4 * r)), nb = __uint_as_float(__ldg(rowblk + 4 * r + 1));
        const unsigned q = __ldg(rowblk + 48 + r);  // ok | dj << 3 | j0 << 4
        const int okr = (int)(q & 7u);
        if ((okr & 3) != 3) continue;  // the row lies outside the border rectangle or the picture's rectangle: untouched (kernels.cl.swift:77,509)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int ok = ce[c].ok & okr;
            if ((ok & 3) != 3 || xt + c >= W) continue;
            const bool chroma = ((r | c) & 1) == 0;
            Taps k;
            k.i0 = ce[c].i0, k.i1 = ce[c].i1, k.j0 = (int)(q >> 4), k.j1 = (int)(q >> 4) + (int)((q >> 3) & 1u);
            const float na = sub(1.f, ce[c].a);
            k.w00 = mul(na, nb), k.w10 = mul(ce[c].a, nb), k.w01 = mul(na, b), k.w11 = mul(ce[c].a, b);  // make_taps' weights
            float* __restrict__ py = sY + r * SVB_UNIT_W + c;
            float* __restrict__ pc = sC + (r >> 1) * SVB_UNIT_W;
            float oy, ou, ov;
            rgba_pixel(s, opacity, fc, (ok & 4) != 0, k, unorm_f(*py), chroma ? unorm_f(pc[0]) : 0.f, chroma ? unorm_f(pc[1]) : 0.f, oy, ou, ov);
            *py = quantf(oy);
            if (chroma) pc[0] = quantf(ou), pc[1] = quantf(ov);
        }
    }
}

__device__ __forceinline__ size_t strip_layer_words(const SvbFrameDesc* __restrict__ F) {  // tiles_x / tiles_y hold the unit counts in a strip batch
    return (size_t)(F->tiles_x * (SVB_UCOL_WORDS + 4) + F->tiles_y * (SVB_UROW_WORDS + 4));
}

// exactly one lane of the (converged) warp: ptxas then issues the TMA instructions below without its one-lane-at-a-time loop
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

}  // namespace svb

// ---- pre-pass: the coordinate tables of every separable layer of every frame of the batch, unit-blocked (svb_desc.h), and the
// column / row records a warp plans its units from.  grid (unit columns + unit rows of the largest frame, max layers, frames),
// 96 threads: a block fills one column block (64 luma + 32 chroma entries) or one row block (8 + 4 entries) and writes its record.
// (at most 32 registers: 96 x 32 fit beside the resident CTAs of the compositor launched before, so this pre-pass runs UNDER that launch)
extern "C" __global__ void __launch_bounds__(96, 21) svb_strip_tables(const SvbFrameDesc* __restrict__ frames, uint32_t* __restrict__ tables, int* __restrict__ unit_counter) {
    using namespace svb;
    if ((blockIdx.x | blockIdx.y | blockIdx.z | threadIdx.x) == 0) *unit_counter = 0;  // svb_mix_strip claims its units from it
    const SvbFrameDesc* __restrict__ F = frames + blockIdx.z;
    const int l = (int)blockIdx.y;
    if (l >= F->nlayers) return;
    const SvbLayerDesc* __restrict__ L = &F->layers[l];
    if (!(L->flags & SVB_LAYER_SEPARABLE)) return;
    const int ux_n = F->tiles_x, uy_n = F->tiles_y, W = F->width, H = F->height;
    const int b = (int)blockIdx.x, t = (int)threadIdx.x;
    if (b >= ux_n + uy_n) return;
    const bool yuv = L->format == SVB_NV12 || L->format == SVB_Y420P;  // BGRA / RGBA layers only use the luma-resolution entries
    const bool ring = (F->flags & SVB_FRAME_RING) != 0;  // svb_mix_ring stages tile-sized boxes (box_*), svb_mix_strip unit-sized ones (sbox_*)
    const bool staged = (L->flags & (ring ? SVB_LAYER_STAGED : SVB_LAYER_STAGED_S)) != 0;
    const int bx_w = ring ? L->box_w : L->sbox_w, bx_h = ring ? L->box_h : L->sbox_h, bx_cw = ring ? L->box_cw : L->sbox_cw, bx_ch = ring ? L->box_ch : L->sbox_ch;
    uint32_t* __restrict__ base = tables + F->table_base + (size_t)l * strip_layer_words(F);
    uint32_t* __restrict__ rbase = base + ux_n * SVB_UCOL_WORDS;
    uint32_t* __restrict__ crec = rbase + uy_n * SVB_UROW_WORDS;
    uint32_t* __restrict__ rrec = crec + 4 * ux_n;
    __shared__ int s_i0[96], s_i1[96], s_ok[96];
    auto odd = [](int ok) { return (ok & 1) != 0 && ok != 7; };
    const bool colblk = b < ux_n;
    Ent e;
    e.a = 0.5f, e.i0 = e.i1 = 0, e.ok = 7;
    bool have = false;
    if (colblk) {
        if (t < SVB_UNIT_W) {
            e = ent_col_y(L, W, b * SVB_UNIT_W + t), have = true;
            uint32_t* __restrict__ o = base + b * SVB_UCOL_WORDS + t;
            o[0] = __float_as_uint(e.a), o[SVB_UNIT_W] = pack_ent(e);
        } else if (yuv) {
            e = ent_col_c(L, W, b * (SVB_UNIT_W / 2) + t - SVB_UNIT_W), have = true;
            uint32_t* __restrict__ o = base + b * SVB_UCOL_WORDS + 2 * SVB_UNIT_W + (t - SVB_UNIT_W);
            o[0] = __float_as_uint(e.a), o[SVB_UNIT_W / 2] = pack_ent(e);
        }
    } else {
        const int r = b - ux_n;
        const unsigned pitchY = staged ? (unsigned)bx_w : 0u, pitchC = staged ? (unsigned)(L->format == SVB_NV12 ? 2 * bx_cw : bx_cw) : 0u;
        uint32_t* __restrict__ blk = rbase + r * SVB_UROW_WORDS;
        if (t < SVB_UNIT_H) e = ent_row_y(L, H, r * SVB_UNIT_H + t), have = true;
        else if (t < SVB_UNIT_H + SVB_UNIT_H / 2 && yuv) e = ent_row_c(L, H, r * (SVB_UNIT_H / 2) + t - SVB_UNIT_H), have = true;
        if (have) {
            const unsigned pitch = t < SVB_UNIT_H ? pitchY : pitchC;
            reinterpret_cast<uint4*>(blk)[t] = make_uint4(__float_as_uint(e.a), __float_as_uint(sub(1.f, e.a)), (unsigned)e.i0 * pitch, (unsigned)e.i1 * pitch);
            blk[48 + t] = (unsigned)e.ok | ((unsigned)(e.i1 - e.i0) << 3) | ((unsigned)e.i0 << 4);
        }
    }
    s_i0[t] = e.i0, s_i1[t] = e.i1, s_ok[t] = e.ok;
    const int half = __syncthreads_and(!have || __float_as_uint(e.a) == 0x3f000000u);  // (also the barrier before s_* are read)
    const int mixed = __syncthreads_or(have && odd(e.ok));
    if (t != 0) return;
    // first and last entry of the valid range: the clamped tap indices and the inside / outside classes are monotone along an axis, so
    // the ends bound the footprint and decide for the whole range; x origins are rounded down to 16 bytes for TMA
    unsigned flags = (half ? SVB_UREC_HALF : 0u) | (mixed ? SVB_UREC_MIXED : 0u);
    if (colblk) {
        const int lastc = min(SVB_UNIT_W, W - b * SVB_UNIT_W) - 1, c0 = SVB_UNIT_W, c1 = SVB_UNIT_W + (lastc >> 1);
        const int ix0 = min(s_i0[0], s_i0[lastc]) & ~15, ic0 = yuv ? (min(s_i0[c0], s_i0[c1]) & (L->format == SVB_NV12 ? ~7 : ~15)) : 0;
        bool fits = staged && max(s_i1[0], s_i1[lastc]) - ix0 < bx_w, full = s_ok[0] == 7 && s_ok[lastc] == 7;
        bool xfree = s_i1[0] != s_i0[0] && s_i1[lastc] != s_i0[lastc];
        if (yuv) fits = fits && max(s_i1[c0], s_i1[c1]) - ic0 < bx_cw, xfree = xfree && s_i1[c0] != s_i0[c0] && s_i1[c1] != s_i0[c1];
        flags |= (full ? SVB_UREC_FULL : 0u) | (xfree ? SVB_UREC_XFREE : 0u) | (fits ? SVB_UREC_FITS : 0u);
        const int ic1 = yuv ? max(s_i1[c0], s_i1[c1]) : 0;
        reinterpret_cast<uint4*>(crec)[b] = make_uint4((unsigned)ix0 | ((unsigned)ic0 << 16), (unsigned)max(s_i1[0], s_i1[lastc]) | ((unsigned)ic1 << 16), flags, 0u);
    } else {
        const int r = b - ux_n, lastr = min(SVB_UNIT_H, H - r * SVB_UNIT_H) - 1, c0 = SVB_UNIT_H, c1 = SVB_UNIT_H + (lastr >> 1);
        const int jy0 = min(s_i0[0], s_i0[lastr]), jc0 = yuv ? min(s_i0[c0], s_i0[c1]) : 0;
        bool fits = staged && max(s_i1[0], s_i1[lastr]) - jy0 < bx_h, full = s_ok[0] == 7 && s_ok[lastr] == 7;
        if (yuv) fits = fits && max(s_i1[c0], s_i1[c1]) - jc0 < bx_ch;
        flags |= (full ? SVB_UREC_FULL : 0u) | (fits ? SVB_UREC_FITS : 0u);
        const int jc1 = yuv ? max(s_i1[c0], s_i1[c1]) : 0;
        reinterpret_cast<uint4*>(rrec)[r] = make_uint4((unsigned)jy0 | ((unsigned)jc0 << 16), (unsigned)max(s_i1[0], s_i1[lastr]) | ((unsigned)jc1 << 16), flags, 0u);
    }
}

// ---- the compositor ---------------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(SVB_STRIP_THREADS, SVB_STRIP_MIN_CTAS)
    svb_mix_strip(const SvbFrameDesc* __restrict__ frames, const uint32_t* __restrict__ tables, int nframes, int total_units, float one, int* __restrict__ unit_counter, int box_y_bytes,
                  int box_c_bytes, int plan_slot_bytes) {
    using namespace svb;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* const bar = reinterpret_cast<uint64_t*>(smem_raw) + 2 * warp;  // this warp's two stages
    const unsigned stage_bytes = (unsigned)(box_y_bytes + box_c_bytes) + SVB_STRIP_TAB_BYTES;
    unsigned char* const mine = smem_raw + SVB_STRIP_HDR_BYTES + (size_t)warp * (size_t)(SVB_STRIP_STATE_BYTES + 2 * plan_slot_bytes + 2 * stage_bytes);
    float2* const sY = reinterpret_cast<float2*>(mine);  // [12 rows][32 lanes]: luma pairs of rows 0..7, then (U, V) of chroma rows 0..3
    const unsigned state = smem_u32(mine) + 8u * lane;
    const unsigned plan0 = smem_u32(mine + SVB_STRIP_STATE_BYTES);    // two plan slots
    const unsigned stage0 = plan0 + 2u * (unsigned)plan_slot_bytes;  // stage s: luma box, chroma box, table blocks
    const unsigned tab_off = (unsigned)(box_y_bytes + box_c_bytes);
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const unsigned mb0 = smem_u32(&bar[0]);
    // first unit of every frame of the batch (at most 64 frames: two per lane), to find a unit's frame with two votes
    const int firstA = lane < nframes ? frames[lane].first_tile : 0x7fffffff, firstB = lane + 32 < nframes ? frames[lane + 32].first_tile : 0x7fffffff;
    int fenced = -1;  // (after the tensor-map table has wrapped) frame whose maps the issuing lane has acquired
    unsigned phase0 = 0, phase1 = 0;
    int stage = 0;        // buffer that holds (or is about to receive) the next staged layer to consume
    bool primed = false;  // this unit's first staged layer was put in flight by the previous unit

    // ---- the plan of unit u, into plan slot s: lane = layer (svb_desc.h: header + one record per layer that touches the unit) ----
    auto plan_unit = [&](int u, int s) {
        const int f = __popc(__ballot_sync(0xffffffffu, u >= firstA)) + __popc(__ballot_sync(0xffffffffu, u >= firstB)) - 1;
        const int first = f < 32 ? __shfl_sync(0xffffffffu, firstA, f) : __shfl_sync(0xffffffffu, firstB, f - 32);
        const SvbFrameDesc* __restrict__ F = frames + f;
        const int ux_n = F->tiles_x, nl = F->nlayers, local = u - first;
        const int uy = local / ux_n, ux = local - uy * ux_n, x0 = ux * SVB_UNIT_W, y0 = uy * SVB_UNIT_H;
        unsigned mode = PLAN_SKIP;
        bool covers = false;
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, r2 = r0, r3 = r0;
        if (lane < nl) {
            const uint4* __restrict__ pc = reinterpret_cast<const uint4*>(&F->layers[lane].pc);
            const uint4 c2 = __ldg(pc + 2), c3 = __ldg(pc + 3);
            const unsigned fmt = c2.y & 0xffu, lflags = c2.y >> 8;
            if ((int)c3.x < x0 + SVB_UNIT_W && (int)c3.z > x0 && (int)c3.y < y0 + SVB_UNIT_H && (int)c3.w > y0) {  // the layer's rectangle touches the unit
                r0.w = c2.x;
                r1.z = c2.z + (unsigned)ux * SVB_UCOL_WORDS, r1.w = c2.z + (unsigned)ux_n * SVB_UCOL_WORDS + (unsigned)uy * SVB_UROW_WORDS;
                if (!(lflags & SVB_LAYER_SEPARABLE)) {
                    mode = PLAN_GENERIC;
                } else if (fmt != SVB_NV12 && fmt != SVB_Y420P) {
                    mode = PLAN_TABLE_RGBA;
                } else {
                    const uint4* __restrict__ rec = reinterpret_cast<const uint4*>(tables + c2.w);
                    const uint4 cq = __ldg(rec + ux), rq = __ldg(rec + ux_n + uy);
                    const uint2 cr = make_uint2(cq.x, cq.z), rr = make_uint2(rq.x, rq.z);
                    const uint4 c1 = __ldg(pc + 1);
                    r2 = __ldg(pc);
                    const unsigned both = cr.y & rr.y;
                    const bool full = (both & SVB_UREC_FULL) != 0u;
                    mode = !(both & SVB_UREC_FITS) ? PLAN_GENERIC : (full && (cr.y & SVB_UREC_XFREE) ? PLAN_STAGED : PLAN_STAGED_EDGE);
                    covers = full && (lflags & SVB_LAYER_UNIT_OPACITY);
                    r0.y = (cr.x & 0xffffu) | (rr.x << 16), r0.z = (cr.x >> 16) | (rr.x & 0xffff0000u);
                    r1.x = c1.z, r1.y = c1.w;
                    r3 = make_uint4(c1.x, c1.y, (both & SVB_UREC_HALF) | ((cr.y | rr.y) & SVB_UREC_MIXED), 0u);
                }
                r0.x = mode | ((unsigned)lane << 8) | (fmt << 16) | (lflags << 20);
            }
        }
        unsigned act = __ballot_sync(0xffffffffu, mode != PLAN_SKIP);
        const unsigned cov = __ballot_sync(0xffffffffu, covers), stg = __ballot_sync(0xffffffffu, mode >= PLAN_STAGED), inner = __ballot_sync(0xffffffffu, mode == PLAN_STAGED);
        unsigned first_covers = 0;
        if (cov) {
            const unsigned top = 31u - (unsigned)__clz(cov);
            act &= ~((1u << top) - 1u);  // drop what the topmost covering layer hides
            first_covers = (inner >> top) & 1u;
        }
        const unsigned slot_a = plan0 + (unsigned)s * (unsigned)plan_slot_bytes, below = act & ((1u << lane) - 1u);
        auto sts4 = [](unsigned a, const uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); };
        if ((act >> lane) & 1u) {
            const unsigned a = slot_a + SVB_UPLAN_HDR_BYTES + SVB_UPLAN_REC_BYTES * (unsigned)__popc(below);
            sts4(a, r0), sts4(a + 16, r1), sts4(a + 32, r2), sts4(a + 48, r3);
        }
        // bit i of the staged mask: the i-th listed layer is staged -- every listed staged layer sets the bit of its own position
        const unsigned smask = __reduce_or_sync(0xffffffffu, ((act & stg) >> lane) & 1u ? 1u << __popc(below) : 0u);
        if (lane == 0) sts4(slot_a, make_uint4((unsigned)__popc(act) | (smask << 16), (unsigned)x0 | ((unsigned)y0 << 16), (unsigned)f, first_covers));
        __syncwarp();
    };
    auto bulk = [&](unsigned dst, const void* src, unsigned bytes, unsigned mb) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
    };
    // the async copies of the staged layer whose plan record lies at shared-memory address ra, into stage b (one elected lane); f = its frame
    auto issue = [&](unsigned ra, int f, int b) {
        const uint4 r0 = lds_u4(ra), r1 = lds_u4(ra + 16), r2 = lds_u4(ra + 32), r3 = lds_u4(ra + 48);
        if (elect_one()) {
            if (f != fenced) {
                const SvbFrameDesc* __restrict__ TF = frames + f;
                if (TF->flags & SVB_FRAME_TMAP_FENCE)
                    for (int q = 0; q < TF->nlayers; ++q)
                        if (TF->layers[q].flags & SVB_LAYER_STAGED_S) {
                            tmap_acquire((const void*)TF->layers[q].stmap[0]);
                            tmap_acquire((const void*)TF->layers[q].stmap[1]);
                            if (TF->layers[q].format != SVB_NV12) tmap_acquire((const void*)TF->layers[q].stmap[2]);
                        }
                fenced = f;
            }
            const unsigned dst = stage0 + (unsigned)b * stage_bytes, mb = mb0 + 8u * (unsigned)b;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(r1.x) : "memory");
            auto tma = [&](unsigned d, unsigned long long tmap, unsigned x, unsigned y) {
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(d), "l"(tmap), "r"(x), "r"(y), "r"(mb)
                             : "memory");
            };
            tma(dst, ((unsigned long long)r2.y << 32) | r2.x, r0.y & 0xffffu, r0.y >> 16);
            tma(dst + (unsigned)box_y_bytes, ((unsigned long long)r2.w << 32) | r2.z, r0.z & 0xffffu, r0.z >> 16);
            if (((r0.x >> 16) & 0xfu) != SVB_NV12) tma(dst + (unsigned)box_y_bytes + (unsigned)box_c_bytes / 2u, ((unsigned long long)r3.y << 32) | r3.x, r0.z & 0xffffu, r0.z >> 16);
            bulk(dst + tab_off, tables + r1.z, SVB_UCOL_WORDS * 4, mb);
            bulk(dst + tab_off + SVB_UCOL_WORDS * 4, tables + r1.w, SVB_UROW_WORDS * 4, mb);
        }
        __syncwarp();
    };

    // Units are claimed two ahead by lane 0 (the atomic's latency is nobody's wait).  Every pass of the loop below plans the NEXT unit
    // (one call site: the planning code exists once) and then computes the current one, whose plan the pass before left in `slot`.
    int c0 = 0, c1 = 0;  // lane 0's: the next two units to plan
    if (lane == 0) {
        c0 = atomicAdd(unit_counter, 1);
        c1 = atomicAdd(unit_counter, 1);
    }
    bool have_cur = false;
    int slot = 1;  // plan slot of the current unit (the first pass has none and plans into slot 0)
    for (;;) {
        const int un = __shfl_sync(0xffffffffu, c0, 0);  // (claimed at least a whole unit ago, but for the first two)
        const bool have_nxt = un < total_units;
        if (lane == 0) c0 = c1, c1 = atomicAdd(unit_counter, 1);
        const unsigned plan = plan0 + (unsigned)slot * (unsigned)plan_slot_bytes, nplan = plan0 + (unsigned)(slot ^ 1) * (unsigned)plan_slot_bytes;
        uint4 h0 = make_uint4(0, 0, 0, 0);
        if (have_cur) {  // this unit's first staged layer goes out before anything else
            h0 = lds_u4(plan);
            if (!primed && (h0.x >> 16)) issue(plan + SVB_UPLAN_HDR_BYTES + SVB_UPLAN_REC_BYTES * (unsigned)(__ffs(h0.x >> 16) - 1), (int)h0.z, stage);
        }
        primed = false;
        if (have_nxt) plan_unit(un, slot ^ 1);  // (every lane is past its reads of that slot: it held the unit before this one)
        if (!have_cur) {
            if (!have_nxt) break;
            have_cur = true, slot ^= 1;
            continue;
        }
        const int nact = (int)(h0.x & 0xffffu), x0 = (int)(h0.y & 0xffffu), y0 = (int)(h0.y >> 16), f = (int)h0.z;
        const unsigned smask = h0.x >> 16;
        const SvbFrameDesc* __restrict__ F = frames + f;
        const int W = F->width, H = F->height, ofmt = F->format, fflags = F->flags;
        const int xt = x0 + 2 * lane, yt = y0;  // this lane's columns xt, xt+1 x rows yt .. yt+7
        const bool live = xt < W;               // W and H even are planner preconditions

        // ---- running picture: img_clear_* (Y = 0, chroma = 0.5 -> 128), or the target's bytes when an earlier pass left them ----
        if (fflags & SVB_FRAME_LOAD_CUR) {
            const uint8_t* const oY = (const uint8_t*)F->out_plane[0];
            const uint8_t* const oU = (const uint8_t*)F->out_plane[1];
            const uint8_t* const oV = (const uint8_t*)F->out_plane[2];
            const int sYb = F->out_stride[0], sUb = F->out_stride[1], sVb = F->out_stride[2];
#pragma unroll
            for (int r = 0; r < SVB_UNIT_H; ++r) {
                unsigned w0 = 0;
                if (live && yt + r < H) w0 = *(const unsigned short*)(oY + (size_t)(yt + r) * sYb + xt);
                sY[r * 32 + lane] = bytes2(opaque(w0 & 0xff), opaque(w0 >> 8));
            }
#pragma unroll
            for (int k = 0; k < SVB_UNIT_H / 2; ++k) {
                unsigned cu = 128, cv = 128;
                if (live && yt + 2 * k < H) {
                    if (ofmt == SVB_NV12) {
                        const unsigned w0 = *(const unsigned short*)(oU + (size_t)((yt >> 1) + k) * sUb + xt);
                        cu = w0 & 0xff, cv = w0 >> 8;
                    } else {
                        cu = oU[(size_t)((yt >> 1) + k) * sUb + (xt >> 1)], cv = oV[(size_t)((yt >> 1) + k) * sVb + (xt >> 1)];
                    }
                }
                sY[(SVB_UNIT_H + k) * 32 + lane] = bytes2(opaque(cu), opaque(cv));
            }
        } else if (!(h0.w & 1u)) {  // (bit 0: the first listed layer overwrites every sample without reading it)
#pragma unroll
            for (int r = 0; r < SVB_UNIT_H; ++r) sY[r * 32 + lane] = splat(0.f);
#pragma unroll
            for (int k = 0; k < SVB_UNIT_H / 2; ++k) sY[(SVB_UNIT_H + k) * 32 + lane] = splat(128.f);
        }

#pragma unroll 1
        for (int i = 0; i < nact; ++i) {
            const unsigned ra = plan + SVB_UPLAN_HDR_BYTES + SVB_UPLAN_REC_BYTES * (unsigned)i;
            const uint4 r0 = lds_u4(ra), r1 = lds_u4(ra + 16);
            const int mode = (int)(r0.x & 0xffu);
            if (mode >= PLAN_STAGED) {
                __syncwarp();  // every lane is past its reads of the other stage
                // refill the other stage: the next staged layer of this unit, else the first one of the next unit
                const unsigned rest = smask >> (i + 1);
                if (rest) {
                    issue(ra + SVB_UPLAN_REC_BYTES * (unsigned)__ffs(rest), f, stage ^ 1);
                } else if (have_nxt) {
                    const uint4 g0 = lds_u4(nplan);
                    const unsigned smn = g0.x >> 16;
                    if (smn) {
                        issue(nplan + SVB_UPLAN_HDR_BYTES + SVB_UPLAN_REC_BYTES * (unsigned)(__ffs(smn) - 1), (int)g0.z, stage ^ 1);
                        primed = true;
                    }
                }
                if (stage == 0) mbar_wait(&bar[0], phase0), phase0 ^= 1;
                else mbar_wait(&bar[1], phase1), phase1 ^= 1;
                const unsigned fmt = (r0.x >> 16) & 0xfu, lflags = r0.x >> 20;
                const unsigned pitchY = r1.y & 0xffffu, pitchC = r1.y >> 16, cstep = fmt == SVB_NV12 ? 2u : 1u;
                const unsigned bY = stage0 + (unsigned)stage * stage_bytes, bC = bY + (unsigned)box_y_bytes, tab = bY + tab_off;
                const unsigned vofs = fmt == SVB_NV12 ? 1u : (unsigned)box_c_bytes / 2u;
                const unsigned colY = bY - (r0.y & 0xffffu) - (r0.y >> 16) * pitchY, colC = bC - (r0.z & 0xffffu) * cstep - (r0.z >> 16) * pitchC;
                const float alpha = __uint_as_float(r0.w);
                const unsigned uflags = lds_u1(ra + 56);  // SVB_UREC_HALF / SVB_UREC_MIXED of the unit
                if (mode == PLAN_STAGED_EDGE || !(lflags & SVB_LAYER_OPACITY_01)) {
                    const SvbLayerDesc* __restrict__ L = &F->layers[(r0.x >> 8) & 0xffu];
                    const float4 fc = ldrow(L->u.fillColor, 0);
                    const float3 fl = rgb2yuv(fc.x, fc.y, fc.z);
                    const float af = mul(alpha, fc.w);
                    // lean: no sample of the unit lies inside the border rectangle but outside the picture (without a border or letterbox: none ever does)
                    if ((lflags & SVB_LAYER_OPACITY_01) && !(uflags & SVB_UREC_MIXED)) strip_layer_edge<true>(colY, colC, cstep, vofs, tab, tab + 4u * SVB_UCOL_WORDS, lane, alpha, one, make_float4(fl.x, fl.y, fl.z, 0.f), af, sY);
                    else strip_layer_edge<false>(colY, colC, cstep, vofs, tab, tab + 4u * SVB_UCOL_WORDS, lane, alpha, one, make_float4(fl.x, fl.y, fl.z, 0.f), af, sY);
                } else if (lflags & SVB_LAYER_UNIT_OPACITY) {
                    if (uflags & SVB_UREC_HALF) strip_layer<true, true>(colY, colC, cstep, vofs, tab, tab + 4u * SVB_UCOL_WORDS, lane, alpha, one, state);
                    else strip_layer<true, false>(colY, colC, cstep, vofs, tab, tab + 4u * SVB_UCOL_WORDS, lane, alpha, one, state);
                } else {
                    if (uflags & SVB_UREC_HALF) strip_layer<false, true>(colY, colC, cstep, vofs, tab, tab + 4u * SVB_UCOL_WORDS, lane, alpha, one, state);
                    else strip_layer<false, false>(colY, colC, cstep, vofs, tab, tab + 4u * SVB_UCOL_WORDS, lane, alpha, one, state);
                }
                stage ^= 1;
            } else {
                const SvbLayerDesc* __restrict__ L = &F->layers[(r0.x >> 8) & 0xffu];
                float* const py = reinterpret_cast<float*>(sY + lane);
                if (mode == PLAN_TABLE_RGBA) strip_rgba_layer(L, tables + r1.z, tables + r1.w, lane, xt, yt, W, H, py, py + SVB_UNIT_W * SVB_UNIT_H);
                else strip_generic_layer(L, xt, yt, W, H, py, py + SVB_UNIT_W * SVB_UNIT_H);
            }
        }

        // ---- the unit's bytes: two luma bytes per lane and row, one (U, V) pair per lane and chroma row -----------------------
        if (live) {
            const int sYb = F->out_stride[0], sUb = F->out_stride[1], sVb = F->out_stride[2];
            const float2 ONE = splat(one);
            auto pack = [&](float2 v) {  // two integer-valued floats in 0..255 -> two bytes: + 2^23 leaves them in the low mantissa bits
                const float2 x = add2<true>(v, splat(8388608.f), ONE);
                return (unsigned short)__byte_perm(__float_as_uint(x.x), __float_as_uint(x.y), 0x0040);
            };
            uint8_t* pY = (uint8_t*)F->out_plane[0] + (size_t)yt * sYb + xt;
            const int nrow = min(SVB_UNIT_H, H - yt);
#pragma unroll 1
            for (int r = 0; r < nrow; ++r, pY += sYb) *(unsigned short*)pY = pack(sY[r * 32 + lane]);
            if (ofmt == SVB_NV12) {
                uint8_t* pC = (uint8_t*)F->out_plane[1] + (size_t)(yt >> 1) * sUb + xt;
#pragma unroll 1
                for (int k = 0; 2 * k < nrow; ++k, pC += sUb) *(unsigned short*)pC = pack(sY[(SVB_UNIT_H + k) * 32 + lane]);
            } else {
                uint8_t* pU = (uint8_t*)F->out_plane[1] + (size_t)(yt >> 1) * sUb + (xt >> 1);
                uint8_t* pV = (uint8_t*)F->out_plane[2] + (size_t)(yt >> 1) * sVb + (xt >> 1);
#pragma unroll 1
                for (int k = 0; 2 * k < nrow; ++k, pU += sUb, pV += sVb) {
                    const unsigned short p = pack(sY[(SVB_UNIT_H + k) * 32 + lane]);
                    *pU = (uint8_t)(p & 0xff), *pV = (uint8_t)(p >> 8);
                }
            }
        }
        __syncwarp();  // the plan slot and the state are free
        if (!have_nxt) break;
        slot ^= 1;
    }
}
