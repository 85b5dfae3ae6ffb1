// kernels_strip.cuh -- the unit-level machinery of the fused compositor's fast path (round 2): the layer bodies that composite one
// 64x8 unit of an output frame, and svb_strip_tables, the pre-pass that feeds them.  svb_mix_ring (kernels_ring.cuh) is the
// compositor built on them.
//
//   * a warp owns a 64x8 unit; a lane owns two adjacent luma columns x 8 rows and the chroma texel column under them, and walks DOWN
//     its rows: the converted taps of a source row stay in registers and serve the next output row when it continues from there
//     (always at 1:1, four rows in five at the 1.2:1 of the headline workload), so a sample costs two or three new taps, not four;
//   * warps that do not run in lock-step do not share an instruction stream, so the hot code must fit the instruction cache of a
//     scheduler on its own: ONE loop of two output rows (~2 KB) serves luma and chroma rows, NV12 and planar sources alike (a first
//     version with a 5 KB body per mode spent a quarter of its stall samples on instruction fetch, profiles/r2_history.md);
//   * the running picture of the unit lives in shared memory (3.2 KB per warp -- twelve rows 272 bytes apart, svb_desc.h -- integer-valued floats, re-quantised after every layer
//     exactly like the reference's 8-bit target, mix.video.swift:113-125): the row loop stays rolled without register rotation;
//   * everything a row needs arrives in the form the loop consumes: svb_strip_tables evaluates the reference's coordinate chain
//     (kernels.cl.swift:70-78) once per output column and row, bit-exactly, and stores weights with their complements and row
//     offsets already multiplied by the staged box pitch, blocked per unit; it also leaves one record per unit column and unit row
//     (footprint, inside / edge / fill classes) from which a tile is planned with a handful of loads, because the plan is separable.
// The per-sample arithmetic is fast_layer's (kernels_tiled.cuh), operation for operation: packed fp32x2, every multiply and add
// rounded on its own.
#pragma once
#include "kernels_tiled.cuh"

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
    // the lower row of the row before (never at 1:1, one row in five at 1.2:1) -- one bit per row, left in the row block by the
    // pre-pass; the row loop tests a bit.
    unsigned reload = lds_u1(rows + 4u * SVB_UROW_RELOAD_WORD);
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
            if (HALF) out = __fadd2_rn(add2<PK>(mul2<PK>(v, splat(63.75f)), splat(8388608.f), ONE), splat(-8388608.f));
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
        const float2 ci1 = lds_state(sa), ci2 = lds_state(sa + SVB_STATE_PITCH_B);
        conv(B, q1);
        out_row(T, B, w1, ci1, sa);
        if (fetch2) {
            unsigned q0[4];
            raw(q0, w2.z);
            conv(B, q0);
        }
        conv(T, q2);
        out_row(B, T, w2, ci2, sa + SVB_STATE_PITCH_B);
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
            reload >>= 2, ent += 32u, sa += 2u * SVB_STATE_PITCH_B;
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
        float2* __restrict__ st = sY + r * SVB_STATE_PITCH_F2 + lane;
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
// block of the running picture in shared memory (sY: the lane's first luma pair; rows are SVB_STATE_PITCH_F floats apart.  sC likewise). ------
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
        float* __restrict__ py = sY + r * SVB_STATE_PITCH_F + c;
        float* __restrict__ pc = sC + (r >> 1) * SVB_STATE_PITCH_F;
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
            float* __restrict__ py = sY + r * SVB_STATE_PITCH_F + c;
            float* __restrict__ pc = sC + (r >> 1) * SVB_STATE_PITCH_F;
            float oy, ou, ov;
            rgba_pixel(s, opacity, fc, (ok & 4) != 0, k, unorm_f(*py), chroma ? unorm_f(pc[0]) : 0.f, chroma ? unorm_f(pc[1]) : 0.f, oy, ou, ov);
            *py = quantf(oy);
            if (chroma) pc[0] = quantf(ou), pc[1] = quantf(ov);
        }
    }
}

__device__ __forceinline__ size_t strip_layer_words(const SvbFrameDesc* __restrict__ F) {  // tiles_x / tiles_y hold the unit counts in a ring batch
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
// (at most 32 registers, so that many blocks start at once in the tail of the compositor launched before, as its CTAs retire.  It does
// not run UNDER that launch's resident CTAs: three CTAs of nine warps at 72 registers leave no scheduler of an SM room that the
// block scheduler will use -- not even for one-warp blocks; with the compositor held to 64 registers they do run under it --
// measured in round 2, profiles/r2_history.md section 9.  About half of the 4 - 8 us between two compositor launches is this pre-pass.)
extern "C" __global__ void __launch_bounds__(96, 21) svb_strip_tables(const SvbFrameDesc* __restrict__ frames, uint32_t* __restrict__ tables, int* __restrict__ unit_counter) {
    using namespace svb;
    if ((blockIdx.x | blockIdx.y | blockIdx.z | threadIdx.x) == 0) unit_counter[0] = 0, unit_counter[1] = 0;  // svb_mix_ring claims its tiles from word 0 and counts its finished CTAs in word 1 (it leaves both at zero: a batch that finds its tables in place skips this pre-pass)
    const SvbFrameDesc* __restrict__ F = frames + blockIdx.z;
    const int l = (int)blockIdx.y;
    if (l >= F->nlayers) return;
    const SvbLayerDesc* __restrict__ L = &F->layers[l];
    if (!(L->flags & SVB_LAYER_SEPARABLE)) return;
    const int ux_n = F->tiles_x, uy_n = F->tiles_y, W = F->width, H = F->height;
    const int b = (int)blockIdx.x, t = (int)threadIdx.x;
    if (b >= ux_n + uy_n) return;
    const bool yuv = L->format == SVB_NV12 || L->format == SVB_Y420P;  // BGRA / RGBA layers only use the luma-resolution entries
    const bool staged = (L->flags & SVB_LAYER_STAGED) != 0;
    const int bx_w = L->box_w, bx_h = L->box_h, bx_cw = L->box_cw, bx_ch = L->box_ch;  // the tile-sized staged boxes (row offsets carry their pitch)
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
        unsigned reload = 0x101u;  // (strip_layer: the rows that cannot take over the converted taps of the row before)
#pragma unroll 1
        for (int k = 1; k < SVB_UNIT_H + SVB_UNIT_H / 2; ++k)
            if (k != SVB_UNIT_H && s_i0[k] != s_i1[k - 1]) reload |= 1u << k;
        rbase[r * SVB_UROW_WORDS + SVB_UROW_RELOAD_WORD] = reload;
    }
}
