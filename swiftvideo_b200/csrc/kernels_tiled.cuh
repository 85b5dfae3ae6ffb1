// kernels_tiled.cuh -- svb_mix_tiled: the fused compositor's fast path.
//
// svb_mix_tables (a small pre-pass per batch): for every separable YUV layer (no rotation: x outputs depend on x
// only and y outputs on y only -- the planner proves it from the uniforms) the reference's per-pixel coordinate
// chain (kernels.cl.swift:70-78 followed by the OpenCL 1.2 linear sampler's i0/i1/frac) is evaluated bit-exactly
// once per output COLUMN and once per output ROW, into a table in global memory (L2-resident; layout: svb_desc.h).
//
// svb_mix_tiled: a CTA owns 128x32 luma tiles of the output frames of the batch, dealt round-robin in row-major
// order to a grid sized to the SM count (neighbouring tiles run at the same time, so the source lines they share
// are fetched from HBM once).  The running picture of a tile lives in registers -- a thread owns two column pairs
// (2*lane, 2*lane+1) and (64+2*lane, 65+2*lane) x 4 rows of luma and the 2x2 chroma texels under them, held as
// integer-valued floats -- and is re-quantised to 8 bits after every layer, so the bytes equal the reference's
// clear-then-fold over an 8-bit target (mix.video.swift:113-125).  (Column pairs rather than four adjacent
// columns: neighbouring lanes then read neighbouring taps, 2*scale bytes apart, so a warp's byte taps fall in
// distinct banks up to a 2:1 downscale -- with four columns per lane 43 % of the shared-memory wavefronts of the
// v6 kernel were bank-conflict replays.)  Per tile every layer is planned first (one thread per layer):
//   skip         the layer's rectangle misses the tile
//   staged       the tile lies wholly inside the layer's picture: the source footprint is staged by one TMA 2-D
//                tensor copy per plane (cp.async.bulk.tensor + mbarrier) and the tile's table blocks by two bulk
//                copies, double-buffered so that the copies of the next staged layer fly while this one is
//                computed; a pixel then costs four byte taps from shared memory, the UNORM8 reads, the bilinear
//                sum, the blend and the re-quantisation -- all as packed fp32x2 instructions (FMUL2 / FFMA2: two
//                pixels per issue slot, each lane rounded separately)
//   staged edge  the tile straddles the picture's edge: the same, with a per-pixel class (picture / fill /
//                untouched) from the tables' ok bits
//   generic      rotated layers, BGRA/RGBA sources and footprints too large to stage: the per-pixel evaluator of
//                svb_device.cuh.
#pragma once
#include "svb_device.cuh"

#ifndef SVB_ROLL_ROWS
#define SVB_ROLL_ROWS 1  // the layer bodies run their two row pairs as a rolled loop (0: unrolled, 0.7 % slower at three CTAs per SM)
#endif
#ifndef SVB_DYNAMIC_TILES
// 1: CTAs claim tiles from a counter (row-major order, so neighbouring tiles still run together); 0: static round-robin.
// Tiles cost between zero and eight layers, and with a static deal the slowest CTA's share decided the kernel time
// (v7 profile: SM cycles active 1.01 M on average against 1.13 M elapsed).
#define SVB_DYNAMIC_TILES 1
#endif
#ifndef SVB_CLAIM_WARP
// The warp whose lane 0 claims this CTA's tiles and fetches their plans.  Not warp 0: thread 0 issues the TMA copies, and
// whatever a warp does alone makes it late for the next CTA barrier -- two warps late by one chore each cost less than one
// warp late by both.
#define SVB_CLAIM_WARP (SVB_TILED_COMPUTE_WARPS - 1)
#endif


namespace svb {

struct __align__(16) Ent {
    float a;     // fractional weight of the i1 tap
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
// a tensor-map slot that the host has REWRITTEN (SVB_FRAME_TMAP_FENCE, svb_desc.h): make the tensormap proxy re-read it
__device__ __forceinline__ void tmap_acquire(const void* tmap) {
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
}
// The box origin must be 16-byte aligned along the row (x * element size); y is free.
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

// The table entries of a layer, by output column / row (luma) or chroma texel column / row: chroma is produced by the even
// luma column / row (kernels.cl.swift:76).  svb_mix_tables stores them, and plans tiles from the same functions.
__device__ __forceinline__ Ent ent_col_y(const SvbLayerDesc* __restrict__ L, int W, int x) { return axis_entry(axis_chain(&L->u, 0, min(x, W - 1), (float)W), L->width); }
__device__ __forceinline__ Ent ent_col_c(const SvbLayerDesc* __restrict__ L, int W, int c) { return axis_entry(axis_chain(&L->u, 0, min(2 * c, W - 2), (float)W), L->width / 2); }
__device__ __forceinline__ Ent ent_row_y(const SvbLayerDesc* __restrict__ L, int H, int r) { return axis_entry(axis_chain(&L->u, 1, min(r, H - 1), (float)H), L->height); }
__device__ __forceinline__ Ent ent_row_c(const SvbLayerDesc* __restrict__ L, int H, int c) { return axis_entry(axis_chain(&L->u, 1, min(2 * c, H - 2), (float)H), L->height / 2); }

// ---- packed fp32x2 (sm_100): two independent IEEE operations per instruction, each rounded on its own --------
// PK=false spells the same operations as two scalar instructions (identical results; kept for A/B timing:
// SVB_FP32X2=0 in the environment of the host selects it).
template <bool PK> __device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    return PK ? __fmul2_rn(a, b) : make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
}
// ptxas 12.9 contracts  mul.rn.f32x2 + add.rn.f32x2  into FFMA2 even under -fmad=false (it does not for the scalar
// .rn forms), which would round once where the reference rounds twice.  The packed add is therefore issued as
// fma(a, ONE, b) with ONE == 1.0f arriving as a kernel argument: bit-identical to a + b (a*1 is exact), one
// instruction like FADD2, and opaque to the contraction because the multiplicand is not a compile-time constant.  (Measured in round 2:
// with the multiplicand as the IMMEDIATE 1.0 -- one register operand less to read, tools/ubench3.cu -- ptxas folds the fma back into
// FADD2 and contracts it with the FMUL2 in front: 19 FFMA2 with three register operands appear in svb_mix_ring.  Not usable.)
template <bool PK> __device__ __forceinline__ float2 add2(float2 a, float2 b, float2 one) {
    return PK ? __ffma2_rn(a, one, b) : make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
}
template <bool PK> __device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    return PK ? __ffma2_rn(a, b, c) : make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y));
}
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }
// UNORM8 read of two integer-valued floats (same identity as unorm_f)
template <bool PK> __device__ __forceinline__ float2 unorm2(float2 c) { return fma2<PK>(c, splat(SVB_K1), mul2<PK>(c, splat(SVB_K2))); }
__device__ __forceinline__ float2 bytes2(unsigned b0, unsigned b1) { return make_float2(__uint2float_rn(b0), __uint2float_rn(b1)); }
// UNORM8 write of two values, kept as integer-valued floats: rint(clamp(v*255, 0, 255)).  Adding 2^23 rounds to
// the nearest integer, ties to even, exactly like rint; without CLAMP the caller guarantees 0 <= v*255 < 255.5.
template <bool CLAMP, bool PK>
__device__ __forceinline__ float2 quant2(float2 v, float2 one) {
    float2 x = mul2<PK>(v, splat(255.f));
    if (CLAMP) {
        x.x = fminf(fmaxf(x.x, 0.f), 255.f);
        x.y = fminf(fmaxf(x.y, 0.f), 255.f);
    }
    // (the second addition takes a SUM, not a product: nothing for ptxas to contract it with, so it may be a plain packed add --
    // one register operand less to read than fma(x, ONE, -2^23); the layer bodies are bound by register reads, tools/ubench3.cu)
    const float2 y = add2<PK>(x, splat(8388608.f), one);
    return PK ? __fadd2_rn(y, splat(-8388608.f)) : make_float2(__fadd_rn(y.x, -8388608.f), __fadd_rn(y.y, -8388608.f));
}
template <bool PK>
__device__ __forceinline__ float2 bilin2(float2 w00, float2 w10, float2 w01, float2 w11, float2 t00, float2 t10, float2 t01, float2 t11,
                                         float2 one) {
    return add2<PK>(add2<PK>(add2<PK>(mul2<PK>(w00, t00), mul2<PK>(w10, t10), one), mul2<PK>(w01, t01), one), mul2<PK>(w11, t11), one);
}

#define SVB_TILE_TAB_WORDS (SVB_TAB_COL_WORDS + SVB_TAB_ROW_WORDS)  // the column block and the row block of one tile
struct TiledSmem {  // the fixed part; the boxes follow at SVB_TILED_FIXED_BYTES: Y[0] Y[1] C[0] C[1]
    alignas(128) uint32_t tabs[2][SVB_TILE_TAB_WORDS];  // the staged layer's table blocks, copied with its boxes
    alignas(16) SvbTilePlan plan[2];                    // [tile parity]: the tile's plan, fetched by one bulk copy (svb_mix_plan wrote it)
    alignas(8) uint64_t bar[2];                         // box buffers
    uint64_t pbar;                                      // plan slot
    int tile_idx[3];                                    // ring over this CTA's tile sequence: index of its k-th tile in slot k % 3 (>= total: none)
};
static_assert(sizeof(TiledSmem) <= SVB_TILED_FIXED_BYTES && SVB_TILED_FIXED_BYTES % 128 == 0, "SVB_TILED_FIXED_BYTES (svb_desc.h) must cover TiledSmem");
enum { PLAN_SKIP = 0, PLAN_GENERIC = 1, PLAN_TABLE_RGBA = 2, PLAN_STAGED = 4, PLAN_STAGED_EDGE = 5 };  // >= PLAN_STAGED: boxes come by TMA

__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {  // 16-byte granules
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// table words of an entry (svb_desc.h): a, and p = i0 | (i1 - i0) << 16 | ok << 17  (i1 - i0 is 0 or 1: both are clamps
// of consecutive integers; the planner keeps planes wider than 65535 texels off this path)
__device__ __forceinline__ uint32_t pack_ent(const Ent& e) { return (uint32_t)e.i0 | ((uint32_t)(e.i1 - e.i0) << 16) | ((uint32_t)e.ok << 17); }
__device__ __forceinline__ Ent unpack_ent(uint32_t a, uint32_t p) {
    Ent e;
    e.a = __uint_as_float(a), e.i0 = (int)(p & 0xffffu), e.i1 = e.i0 + (int)((p >> 16) & 1u), e.ok = (int)(p >> 17);
    return e;
}
__device__ __forceinline__ unsigned lds_u8(unsigned addr) {  // opaque u32 (see ldg_u8)
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

template <int OFF> __device__ __forceinline__ unsigned lds_u8o(unsigned addr) {  // [addr + OFF] with OFF in the instruction
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}

// gather coordinate of a table entry: the footprint of tex2Dgather at coordinate c is texels floor(c - 1/2) and the next,
// clamped one by one.  For an unclamped index i the table holds i0 = clamp(i), i1 = clamp(i + 1): c = i0 + 1 reproduces
// (i0, i1) in every case but the left / top clamp (i < 0: i0 == i1 == 0), where c = 0 does.
__device__ __forceinline__ float gather_coord(uint32_t p) {
    const unsigned i0 = p & 0xffffu, d = (p >> 16) & 1u;
    return (float)(int)(i0 + ((d | i0) != 0u ? 1u : 0u));
}

// (1-a)(1-b) T00 + a(1-b) T10 + (1-a)b T01 + ab T11 over a gathered footprint g = (T01, T11, T10, T00) -- tools/tex_probe.cu --
// with the sampler's weights and summation order (svb_device.cuh: make_taps, filt).  Scalar on purpose: pairing weights to the
// gather's component order costs two moves per use, and scalar multiplies and adds issue on both FMA pipes.
__device__ __forceinline__ float gather_filter(const float4 g, float na, float a, float b, float nb) {
    const float w00 = mul(na, nb), w10 = mul(a, nb), w01 = mul(na, b), w11 = mul(a, b);
    return add(add(add(mul(w00, g.w), mul(w10, g.z)), mul(w01, g.x)), mul(w11, g.y));
}

// Tables of one layer of one frame: tiles_x column blocks, then tiles_y row blocks
struct Tabs {
    const uint32_t* col;
    const uint32_t* row;
};
__device__ __forceinline__ Tabs layer_tabs(const uint32_t* __restrict__ tables, const SvbFrameDesc* __restrict__ F, int l) {
    Tabs t;
    t.col = tables + F->table_base + (size_t)l * (size_t)(F->tiles_x * SVB_TAB_COL_WORDS + F->tiles_y * SVB_TAB_ROW_WORDS);
    t.row = t.col + F->tiles_x * SVB_TAB_COL_WORDS;
    return t;
}
// word offsets of an entry's `a` inside its layer's tables, and from there to its `p`
__device__ __forceinline__ int col_y_word(int x) { return (x / SVB_TILE_W) * SVB_TAB_COL_WORDS + (x % SVB_TILE_W); }
__device__ __forceinline__ int col_c_word(int c) { return (c / (SVB_TILE_W / 2)) * SVB_TAB_COL_WORDS + 2 * SVB_TILE_W + (c % (SVB_TILE_W / 2)); }
__device__ __forceinline__ int row_y_word(int r) { return (r / SVB_TILE_H) * SVB_TAB_ROW_WORDS + 2 * (r % SVB_TILE_H); }
__device__ __forceinline__ int row_c_word(int c) { return (c / (SVB_TILE_H / 2)) * SVB_TAB_ROW_WORDS + 2 * SVB_TILE_H + 2 * (c % (SVB_TILE_H / 2)); }

// One separable YUV layer over one tile, taps staged in shared memory.
//   Yi / Ui / Vi: the running picture as integer-valued floats; pairs hold two horizontally adjacent samples.
//   MODE 0: tile wholly inside the picture, opacity == 1 (cur*(1-1) + v*1 == v exactly: no blend)
//   MODE 1: tile wholly inside the picture, 0 <= opacity <= 1 (blended values stay in [0,1]: no store clamp)
//   MODE 2: edge tile whose samples are either inside the picture or untouched (no fill sample in this warp's block, every
//           row either wholly inside or outside the border rectangle; 0 <= opacity <= 1) -- the common edge.  Rows outside
//           are skipped by their warp (rows are warp-uniform), columns outside keep their value through a 0/1 mask.
//   MODE 3: anything: per-pixel class from the tables' ok bits -- picture / fill / untouched
//           (kernels.cl.swift:77,84-85,96-105) -- and saturating stores
struct FillTerms {
    float2 fy, fu, fv, af, naf;  // RGB2YUV(fillColor.rgb, 1) splat; opacity*fillColor.w and its complement
};
//   MODE 0 and 1 are only planned for tiles whose taps are never clamped along x (i1 == i0 + 1 for every column of
//   the tile, luma and chroma): the second tap of a row is then the first one's address plus an immediate, and a pixel
//   costs two address adds (one per tap row) instead of four.  N12 tells them the chroma layout (NV12: U and V
//   interleaved, so all eight chroma taps of a texel hang off two addresses); MODE 2 reads it from pitchC / stepC.
template <int MODE, bool PK, bool N12>
__device__ __forceinline__ void fast_layer(unsigned boxY, unsigned boxU, unsigned boxV, const uint32_t* __restrict__ tabs, int lane, int warp, int iy0, int jy0,
                                           int ic0, int jc0, int pitchY, int pitchC, int stepC, float alpha, float onef, const FillTerms& ft,
                                           float2 (&Yi)[4][2], float2 (&Ui)[2], float2 (&Vi)[2]) {
    constexpr bool UNIT = MODE == 0, EDGE = MODE == 2, GEN = MODE == 3, XCL = MODE >= 2;  // XCL: taps may be clamped along x
    const float2 AL = splat(alpha), NAL = splat(sub(1.f, alpha)), ONE = splat(onef);
    unsigned o0[4], o1[4];
    int okc[4] = {7, 7, 7, 7};
    float2 A[2], NA[2], M[2];  // M (EDGE): 1 where the column is inside the picture, else 0
#pragma unroll
    for (int p = 0; p < 2; ++p) {  // pair p = luma columns 64p + 2*lane, +1 of the tile: 8-byte reads, lane after lane
        const float2 a = *reinterpret_cast<const float2*>(tabs + 64 * p + 2 * lane);
        const uint2 e = *reinterpret_cast<const uint2*>(tabs + SVB_TILE_W + 64 * p + 2 * lane);
        o0[2 * p] = boxY + ((e.x & 0xffffu) - iy0), o0[2 * p + 1] = boxY + ((e.y & 0xffffu) - iy0);
        if (XCL) o1[2 * p] = o0[2 * p] + ((e.x >> 16) & 1u), o1[2 * p + 1] = o0[2 * p + 1] + ((e.y >> 16) & 1u);
        if (GEN) okc[2 * p] = (int)(e.x >> 17), okc[2 * p + 1] = (int)(e.y >> 17);
        if (EDGE) M[p] = make_float2((e.x >> 17) == 7u ? 1.f : 0.f, (e.y >> 17) == 7u ? 1.f : 0.f);
        A[p] = a;
        NA[p] = make_float2(sub(1.f, a.x), sub(1.f, a.y));
    }
    // chroma columns lane and 32 + lane of the tile (the texels under the two luma pairs)
    const uint32_t pc0 = tabs[2 * SVB_TILE_W + SVB_TILE_W / 2 + lane], pc1 = tabs[2 * SVB_TILE_W + SVB_TILE_W / 2 + 32 + lane];
    const unsigned sC = XCL ? (unsigned)stepC : (N12 ? 2u : 1u);
    const unsigned oc00 = ((pc0 & 0xffffu) - ic0) * sC, oc10 = ((pc1 & 0xffffu) - ic0) * sC;
    unsigned oc01 = 0, oc11 = 0;
    int okc0 = 7, okc1 = 7;
    float2 MC = splat(1.f);
    if (XCL) oc01 = oc00 + ((pc0 >> 16) & 1u) * sC, oc11 = oc10 + ((pc1 >> 16) & 1u) * sC;
    if (GEN) okc0 = (int)(pc0 >> 17), okc1 = (int)(pc1 >> 17);
    if (EDGE) MC = make_float2((pc0 >> 17) == 7u ? 1.f : 0.f, (pc1 >> 17) == 7u ? 1.f : 0.f);
    const float2 AC = make_float2(__uint_as_float(tabs[2 * SVB_TILE_W + lane]), __uint_as_float(tabs[2 * SVB_TILE_W + 32 + lane]));
    const float2 NAC = make_float2(sub(1.f, AC.x), sub(1.f, AC.y));
    // blend -> UNORM8 write -> the next layer's UNORM8 read stays an integer-valued float
    auto settle = [&](float2 cur_i, float2 v, float2 fillc, float lo, int ok0, int ok1, float2 m) -> float2 {
        if (UNIT) return quant2<false, PK>(v, ONE);
        const float2 cur = unorm2<PK>(cur_i);
        const float2 qi = quant2<GEN, PK>(add2<PK>(mul2<PK>(cur, NAL), mul2<PK>(v, AL), ONE), ONE);
        // columns outside the picture keep cur_i: cur_i + m*(qi - cur_i) in integer-valued floats, every step exact
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
    const uint32_t* __restrict__ rows = tabs + SVB_TAB_COL_WORDS;  // rows past the frame's bottom hold the last row's entry
    // two luma rows (4*warp + 2*k, +1) and the chroma row under them (2*warp + k)
    auto row_pair = [&](int k, float2 (&Y0)[2], float2 (&Y1)[2], float2& Uk, float2& Vk) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float2(&Yr)[2] = r == 0 ? Y0 : Y1;
            const uint2 ry = *reinterpret_cast<const uint2*>(rows + 2 * (4 * warp + 2 * k + r));
            const int okr = XCL ? (int)(ry.y >> 17) : 7;
            if (EDGE && okr != 7) continue;  // a row outside the picture (warp-uniform): untouched
            const unsigned r0 = ((ry.y & 0xffffu) - jy0) * pitchY, r1 = r0 + ((ry.y >> 16) & 1u) * pitchY;
            const float2 B = splat(__uint_as_float(ry.x)), NB = splat(sub(1.f, __uint_as_float(ry.x)));
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                float2 t00, t10, t01, t11;
                if (XCL) {
                    const unsigned a0 = o0[2 * p], a1 = o1[2 * p], b0 = o0[2 * p + 1], b1 = o1[2 * p + 1];
                    t00 = unorm2<PK>(bytes2(lds_u8(r0 + a0), lds_u8(r0 + b0)));
                    t10 = unorm2<PK>(bytes2(lds_u8(r0 + a1), lds_u8(r0 + b1)));
                    t01 = unorm2<PK>(bytes2(lds_u8(r1 + a0), lds_u8(r1 + b0)));
                    t11 = unorm2<PK>(bytes2(lds_u8(r1 + a1), lds_u8(r1 + b1)));
                } else {
                    const unsigned a0 = r0 + o0[2 * p], b0 = r0 + o0[2 * p + 1], a1 = r1 + o0[2 * p], b1 = r1 + o0[2 * p + 1];
                    t00 = unorm2<PK>(bytes2(lds_u8o<0>(a0), lds_u8o<0>(b0)));
                    t10 = unorm2<PK>(bytes2(lds_u8o<1>(a0), lds_u8o<1>(b0)));
                    t01 = unorm2<PK>(bytes2(lds_u8o<0>(a1), lds_u8o<0>(b1)));
                    t11 = unorm2<PK>(bytes2(lds_u8o<1>(a1), lds_u8o<1>(b1)));
                }
                const float2 v = bilin2<PK>(mul2<PK>(NA[p], NB), mul2<PK>(A[p], NB), mul2<PK>(NA[p], B), mul2<PK>(A[p], B), t00, t10, t01, t11, ONE);
                Yr[p] = settle(Yr[p], v, ft.fy, 0.f, okc[2 * p] & okr, okc[2 * p + 1] & okr, M[p]);
            }
        }
        const uint2 rc = *reinterpret_cast<const uint2*>(rows + 2 * SVB_TILE_H + 2 * (2 * warp + k));
        const int okq = XCL ? (int)(rc.y >> 17) : 7;
        if (EDGE && okq != 7) return;
        const unsigned q0 = ((rc.y & 0xffffu) - jc0) * pitchC, q1 = q0 + ((rc.y >> 16) & 1u) * pitchC;
        const float2 BC = splat(__uint_as_float(rc.x)), NBC = splat(sub(1.f, __uint_as_float(rc.x)));
        const float2 w00 = mul2<PK>(NAC, NBC), w10 = mul2<PK>(AC, NBC), w01 = mul2<PK>(NAC, BC), w11 = mul2<PK>(AC, BC);
        const unsigned u0 = boxU + q0, u1 = boxU + q1;
        float2 u, v;
        if (XCL) {
            const unsigned v0 = boxV + q0, v1 = boxV + q1;
            u = bilin2<PK>(w00, w10, w01, w11, unorm2<PK>(bytes2(lds_u8(u0 + oc00), lds_u8(u0 + oc10))), unorm2<PK>(bytes2(lds_u8(u0 + oc01), lds_u8(u0 + oc11))),
                           unorm2<PK>(bytes2(lds_u8(u1 + oc00), lds_u8(u1 + oc10))), unorm2<PK>(bytes2(lds_u8(u1 + oc01), lds_u8(u1 + oc11))), ONE);
            v = bilin2<PK>(w00, w10, w01, w11, unorm2<PK>(bytes2(lds_u8(v0 + oc00), lds_u8(v0 + oc10))), unorm2<PK>(bytes2(lds_u8(v0 + oc01), lds_u8(v0 + oc11))),
                           unorm2<PK>(bytes2(lds_u8(v1 + oc00), lds_u8(v1 + oc10))), unorm2<PK>(bytes2(lds_u8(v1 + oc01), lds_u8(v1 + oc11))), ONE);
        } else if (N12) {  // (U, V) byte pairs: texel i0 at [c], [c+1], texel i0 + 1 at [c+2], [c+3]
            const unsigned c00 = u0 + oc00, c10 = u0 + oc10, c01 = u1 + oc00, c11 = u1 + oc10;
            u = bilin2<PK>(w00, w10, w01, w11, unorm2<PK>(bytes2(lds_u8o<0>(c00), lds_u8o<0>(c10))), unorm2<PK>(bytes2(lds_u8o<2>(c00), lds_u8o<2>(c10))),
                           unorm2<PK>(bytes2(lds_u8o<0>(c01), lds_u8o<0>(c11))), unorm2<PK>(bytes2(lds_u8o<2>(c01), lds_u8o<2>(c11))), ONE);
            v = bilin2<PK>(w00, w10, w01, w11, unorm2<PK>(bytes2(lds_u8o<1>(c00), lds_u8o<1>(c10))), unorm2<PK>(bytes2(lds_u8o<3>(c00), lds_u8o<3>(c10))),
                           unorm2<PK>(bytes2(lds_u8o<1>(c01), lds_u8o<1>(c11))), unorm2<PK>(bytes2(lds_u8o<3>(c01), lds_u8o<3>(c11))), ONE);
        } else {  // planar chroma: the V box lies boxV - boxU bytes behind the U box
            const unsigned c00 = u0 + oc00, c10 = u0 + oc10, c01 = u1 + oc00, c11 = u1 + oc10, dv = boxV - boxU;
            u = bilin2<PK>(w00, w10, w01, w11, unorm2<PK>(bytes2(lds_u8o<0>(c00), lds_u8o<0>(c10))), unorm2<PK>(bytes2(lds_u8o<1>(c00), lds_u8o<1>(c10))),
                           unorm2<PK>(bytes2(lds_u8o<0>(c01), lds_u8o<0>(c11))), unorm2<PK>(bytes2(lds_u8o<1>(c01), lds_u8o<1>(c11))), ONE);
            v = bilin2<PK>(w00, w10, w01, w11, unorm2<PK>(bytes2(lds_u8o<0>(c00 + dv), lds_u8o<0>(c10 + dv))), unorm2<PK>(bytes2(lds_u8o<1>(c00 + dv), lds_u8o<1>(c10 + dv))),
                           unorm2<PK>(bytes2(lds_u8o<0>(c01 + dv), lds_u8o<0>(c11 + dv))), unorm2<PK>(bytes2(lds_u8o<1>(c01 + dv), lds_u8o<1>(c11 + dv))), ONE);
        }
        Uk = settle(Uk, u, ft.fu, -1.f, okc0 & okq, okc1 & okq, MC);
        Vk = settle(Vk, v, ft.fv, -1.f, okc0 & okq, okc1 & okq, MC);
    };
#if SVB_ROLL_ROWS
    // one copy of the body, run twice; the two halves of the register block trade places after each pass (and are back
    // where they were after the second): half the code for the instruction cache to hold, which three resident CTAs
    // drifting apart make scarce (no_instruction stalls went from 0.36 to 1.34 per issue when the third CTA came in)
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
        row_pair(k, Yi[0], Yi[1], Ui[0], Vi[0]);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            float2 x = Yi[0][p];
            Yi[0][p] = Yi[2][p], Yi[2][p] = x;
            x = Yi[1][p], Yi[1][p] = Yi[3][p], Yi[3][p] = x;
        }
        float2 x = Ui[0];
        Ui[0] = Ui[1], Ui[1] = x;
        x = Vi[0], Vi[0] = Vi[1], Vi[1] = x;
    }
#else
    row_pair(0, Yi[0], Yi[1], Ui[0], Vi[0]);
    row_pair(1, Yi[2], Yi[3], Ui[1], Vi[1]);
#endif
}

// The running picture of a thread as a local array (for the out-of-line layer bodies): st[0..15] luma, row-major over
// columns (xt, xt+1, xt+64, xt+65); st[16..19] U texels, st[20..23] V texels.
__device__ __forceinline__ void state_to_array(const float2 (&Yi)[4][2], const float2 (&Ui)[2], const float2 (&Vi)[2], float* __restrict__ st) {
#pragma unroll
    for (int r = 0; r < 4; ++r) st[4 * r] = Yi[r][0].x, st[4 * r + 1] = Yi[r][0].y, st[4 * r + 2] = Yi[r][1].x, st[4 * r + 3] = Yi[r][1].y;
#pragma unroll
    for (int k = 0; k < 2; ++k) st[16 + 2 * k] = Ui[k].x, st[17 + 2 * k] = Ui[k].y, st[20 + 2 * k] = Vi[k].x, st[21 + 2 * k] = Vi[k].y;
}
__device__ __forceinline__ void array_to_state(const float* __restrict__ st, float2 (&Yi)[4][2], float2 (&Ui)[2], float2 (&Vi)[2]) {
#pragma unroll
    for (int r = 0; r < 4; ++r) Yi[r][0] = make_float2(st[4 * r], st[4 * r + 1]), Yi[r][1] = make_float2(st[4 * r + 2], st[4 * r + 3]);
#pragma unroll
    for (int k = 0; k < 2; ++k) Ui[k] = make_float2(st[16 + 2 * k], st[17 + 2 * k]), Vi[k] = make_float2(st[20 + 2 * k], st[21 + 2 * k]);
}

// Any layer, any tile: the per-pixel evaluator of svb_device.cuh over this thread's 4x4 block.  Kept compact (one
// copy of the evaluator, pixels in a rolled loop over a local copy of the block) so that it does not crowd the
// instruction cache of the fast path; rotated layers, BGRA/RGBA sources and footprints too large to stage come here.
__device__ __noinline__ void generic_layer(const SvbLayerDesc* __restrict__ L, int xt, int yt, float fW, float fH, int W, int H, float* __restrict__ st, int rows = 4) {
    const Src s = layer_src(L);
    const SvbUniforms* __restrict__ U = &L->u;
#pragma unroll 1
    for (int q = 0; q < 4 * rows; ++q) {
        const int r = q >> 2, c = q & 3;
        if (yt + r >= H) break;
        const int x = xt + (c & 1) + (SVB_TILE_W / 2) * (c >> 1);  // columns xt, xt+1, xt+64, xt+65
        if (x >= W) continue;
        const bool chroma = ((r | c) & 1) == 0;
        const int ci = 16 + (r >> 1) * 2 + (c >> 1);  // st[16..19] = U texels, st[20..23] = V texels
        float oy, ou, ov;
        if (eval_pixel(U, s, x, yt + r, fW, fH, chroma, unorm_f(st[q]), chroma ? unorm_f(st[ci]) : 0.f, chroma ? unorm_f(st[ci + 4]) : 0.f, oy, ou,
                       ov)) {
            st[q] = quantf(oy);
            if (chroma) st[ci] = quantf(ou), st[ci + 4] = quantf(ov);
        }
    }
}

// A separable BGRA / RGBA layer (text and logo overlays: upright, any scale) over this thread's 4x4 block: the coordinate
// chain comes from the layer's tables (two divisions and five dot products per pixel in the generic evaluator), the four
// RGBA taps straight from global memory (L1/L2: neighbouring pixels share them), the arithmetic is rgba_pixel's.  Out of
// line like generic_layer; 1.8x faster than it per RGBA picture-in-picture of the bench's geometry (profiles/r1_history.md).
__device__ __noinline__ void rgba_table_layer(const SvbLayerDesc* __restrict__ L, const uint32_t* __restrict__ colblk, const uint32_t* __restrict__ rowblk, int xt, int yt,
                                              int W, int H, float* __restrict__ st, int strip, int rows = 4) {  // strip: which `rows`-row strip of the tile this warp owns
    const Src s = layer_src(L);
    const float opacity = __ldg(&L->u.opacity);
    const float4 fc = ldrow(L->u.fillColor, 0);
    const int lane = threadIdx.x & 31, warp = strip;
    Ent ce[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int x = 2 * lane + (c & 1) + (SVB_TILE_W / 2) * (c >> 1);  // tile-local columns 2l, 2l+1, 64+2l, 65+2l
        ce[c] = unpack_ent(__ldg(colblk + x), __ldg(colblk + SVB_TILE_W + x));
    }
#pragma unroll 1
    for (int r = 0; r < rows; ++r) {
        if (yt + r >= H) break;
        const Ent re = unpack_ent(__ldg(rowblk + 2 * (rows * warp + r)), __ldg(rowblk + 2 * (rows * warp + r) + 1));
        if ((re.ok & 3) != 3) continue;  // the row lies outside the border rectangle or the picture's rectangle: untouched (kernels.cl.swift:77,509)
        const float nb = sub(1.f, re.a);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int ok = ce[c].ok & re.ok;
            if ((ok & 3) != 3 || xt + (c & 1) + (SVB_TILE_W / 2) * (c >> 1) >= W) continue;
            const bool chroma = ((r | c) & 1) == 0;
            const int q = 4 * r + c, ci = 16 + (r >> 1) * 2 + (c >> 1);
            Taps k;
            k.i0 = ce[c].i0, k.i1 = ce[c].i1, k.j0 = re.i0, k.j1 = re.i1;
            const float na = sub(1.f, ce[c].a);
            k.w00 = mul(na, nb), k.w10 = mul(ce[c].a, nb), k.w01 = mul(na, re.a), k.w11 = mul(ce[c].a, re.a);  // make_taps' weights
            float oy, ou, ov;
            rgba_pixel(s, opacity, fc, (ok & 4) != 0, k, unorm_f(st[q]), chroma ? unorm_f(st[ci]) : 0.f, chroma ? unorm_f(st[ci + 4]) : 0.f, oy, ou, ov);
            st[q] = quantf(oy);
            if (chroma) st[ci] = quantf(ou), st[ci + 4] = quantf(ov);
        }
    }
}

// two integer-valued floats in 0..255 -> two bytes: adding 2^23 leaves the integer in the low mantissa bits (exact), one
// PRMT gathers the two low bytes (F2I runs on the quarter-rate pipe, and the tile epilogue had 24 of them)
__device__ __forceinline__ unsigned short pack2(float a, float b) {
    return (unsigned short)__byte_perm(__float_as_uint(__fadd_rn(a, 8388608.f)), __float_as_uint(__fadd_rn(b, 8388608.f)), 0x0040);
}

}  // namespace svb

#ifndef SVB_TILED_MIN_CTAS
#define SVB_TILED_MIN_CTAS 3  // 80 registers per thread (a few spills) and 24 resident warps beat 128 registers and 16 warps by 5.6 %
#endif

namespace svb {

struct TileGeo {
    const SvbFrameDesc* F;
    int frame, x0, y0, lastc, lastr;
};

// Plan of layer `l` on one tile (one thread per layer).  Returns the mode; *covers = the layer overwrites every sample of
// the tile whatever lies below: the tile is wholly inside the picture and opacity == 1, so every pixel takes
// cur*(1-1) + v*1 == v (kernels.cl.swift:86-92).
// out[0] = (mode | l << 8, iy0, jy0, ic0)   out[1] = (jc0, format << 8 | flags, box_w | box_cw << 16, opacity bits)
// and, ready for the one thread that issues the copies (every instruction on its path makes its warp late for the next
// barrier): out[2] = (&tmap Y, &tmap C)   out[3] = (&tmap V or 0, table column block)   out[4] = (table row block, tx bytes, frame)
__device__ __forceinline__ int plan_layer(const uint32_t* __restrict__ tables, const TileGeo& g, int l, int4* __restrict__ out, bool* __restrict__ covers) {
    const SvbFrameDesc* __restrict__ F = g.F;
    const SvbLayerDesc* __restrict__ L = &F->layers[l];
    const int x0 = g.x0, y0 = g.y0, W = F->width, H = F->height;
    int mode, iy0 = 0, jy0 = 0, ic0 = 0, jc0 = 0;
    *covers = false;
    if (L->rect[0] >= x0 + SVB_TILE_W || L->rect[2] <= x0 || L->rect[1] >= y0 + SVB_TILE_H || L->rect[3] <= y0) {
        mode = PLAN_SKIP;
    } else if (!(L->flags & SVB_LAYER_SEPARABLE)) {
        mode = PLAN_GENERIC;
    } else if (L->format != SVB_NV12 && L->format != SVB_Y420P) {
        mode = PLAN_TABLE_RGBA;
    } else {
        // the table entries of the tile's first and last column / row, recomputed (the same functions fill the tables, in this
        // same launch: nothing to wait for)
        const Ent cA = ent_col_y(L, W, x0), cB = ent_col_y(L, W, x0 + g.lastc), rA = ent_row_y(L, H, y0), rB = ent_row_y(L, H, y0 + g.lastr);
        const Ent ccA = ent_col_c(L, W, x0 >> 1), ccB = ent_col_c(L, W, (x0 + g.lastc) >> 1);
        const Ent rcA = ent_row_c(L, H, y0 >> 1), rcB = ent_row_c(L, H, (y0 + g.lastr) >> 1);
        // source footprint of the tile: the clamped tap indices are monotone along each axis, so the ends bound it;
        // x origins are rounded down to 16 bytes for TMA
        iy0 = min(cA.i0, cB.i0) & ~15;
        jy0 = min(rA.i0, rB.i0);
        ic0 = min(ccA.i0, ccB.i0) & (L->format == SVB_NV12 ? ~7 : ~15);
        jc0 = min(rcA.i0, rcB.i0);
        const bool fits = max(cA.i1, cB.i1) - iy0 < L->box_w && max(rA.i1, rB.i1) - jy0 < L->box_h && max(ccA.i1, ccB.i1) - ic0 < L->box_cw &&
                          max(rcA.i1, rcB.i1) - jc0 < L->box_ch;
        // border, tx and uv are monotone too: both ends inside [0,1] means every pixel of the tile is inside the picture
        const bool full = cA.ok == 7 && cB.ok == 7 && rA.ok == 7 && rB.ok == 7;
        // i1 - i0 is 0 only where the tap index was clamped, and the unclamped index is monotone: the ends decide for the tile.
        // The interior layer bodies (MODE 0 / 1) address the second tap of a row as "first + 1".
        const bool xfree = cA.i1 != cA.i0 && cB.i1 != cB.i0 && ccA.i1 != ccA.i0 && ccB.i1 != ccB.i0;
        if (F->flags & SVB_FRAME_GATHER)  // texture path: nothing is staged and the unit clamps, so only inside / edge matters
            mode = !(L->flags & SVB_LAYER_TEX) ? PLAN_GENERIC : (full ? PLAN_STAGED : PLAN_STAGED_EDGE);
        else
            mode = !((L->flags & SVB_LAYER_STAGED) && fits) ? PLAN_GENERIC : (full && xfree ? PLAN_STAGED : PLAN_STAGED_EDGE);
        *covers = full && (L->flags & SVB_LAYER_UNIT_OPACITY);
    }
    out[0] = make_int4(mode | (l << 8), iy0, jy0, ic0);
    out[1] = make_int4(jc0, (L->format << 8) | (L->flags & 0xff), L->box_w | (L->box_cw << 16), __float_as_int(L->u.opacity));
    out[2] = out[3] = out[4] = make_int4(0, 0, 0, 0);
    if (mode == PLAN_TABLE_RGBA) {  // the tile's table blocks, read in place
        const Tabs tb = layer_tabs(tables, F, l);
        const unsigned long long cb = (unsigned long long)(tb.col + (x0 / SVB_TILE_W) * SVB_TAB_COL_WORDS), rb = (unsigned long long)(tb.row + (y0 / SVB_TILE_H) * SVB_TAB_ROW_WORDS);
        out[3] = make_int4(0, 0, (int)(unsigned)cb, (int)(cb >> 32));
        out[4] = make_int4((int)(unsigned)rb, (int)(rb >> 32), 0, g.frame);
    }
    if (mode >= PLAN_STAGED) {
        const bool n12 = L->format == SVB_NV12;
        const int cbytes = n12 ? L->box_cw * L->box_ch * 2 : L->box_cw * L->box_ch;
        const Tabs tb = layer_tabs(tables, F, l);
        const unsigned long long m0 = L->tmap[0], m1 = L->tmap[1], m2 = n12 ? 0ull : L->tmap[2];
        const unsigned long long cb = (unsigned long long)(tb.col + (x0 / SVB_TILE_W) * SVB_TAB_COL_WORDS), rb = (unsigned long long)(tb.row + (y0 / SVB_TILE_H) * SVB_TAB_ROW_WORDS);
        out[2] = make_int4((int)(unsigned)m0, (int)(m0 >> 32), (int)(unsigned)m1, (int)(m1 >> 32));
        out[3] = make_int4((int)(unsigned)m2, (int)(m2 >> 32), (int)(unsigned)cb, (int)(cb >> 32));
        out[4] = make_int4((int)(unsigned)rb, (int)(rb >> 32), L->box_w * L->box_h + cbytes * (n12 ? 1 : 2) + SVB_TILE_TAB_WORDS * 4, g.frame);
    }
    return mode;
}

// The plan of one tile: one warp, lane l plans layer l of the tile's frame.  The plan that reaches the compositor
// (SvbTilePlan, svb_desc.h) lists only the layers that touch the tile, bottom to top, starting at the topmost layer that
// hides everything under it (occlusion: hidden layers are neither fetched nor computed, the bytes are the same), and carries
// what a tile needs of its frame -- so that the compositor's CTAs fetch a tile's plan with ONE bulk copy and no warp of theirs
// computes anything for it.  (Planning inside the compositor made its planning warp late for the first barrier of every
// tile: barrier waits were 27 % of the stall samples, profiles/r1_history.md.)
__device__ __forceinline__ void plan_tile(const SvbFrameDesc* __restrict__ F, int frame, int local, const uint32_t* __restrict__ tables, SvbTilePlan* __restrict__ plans, int lane) {
    TileGeo g;
    g.F = F, g.frame = frame;
    g.x0 = (local % F->tiles_x) * SVB_TILE_W;
    g.y0 = (local / F->tiles_x) * SVB_TILE_H;
    g.lastc = min(SVB_TILE_W, F->width - g.x0) - 1;
    g.lastr = min(SVB_TILE_H, F->height - g.y0) - 1;
    int4 e[5];
    bool covers = false;
    int mode = PLAN_SKIP;
    if (lane < F->nlayers) mode = plan_layer(tables, g, lane, e, &covers);
    unsigned act = __ballot_sync(0xffffffffu, mode != PLAN_SKIP);
    const unsigned cov = __ballot_sync(0xffffffffu, covers);
    if (cov) act &= ~((1u << (31 - __clz(cov))) - 1u);  // drop what the topmost covering layer hides
    int4* __restrict__ out = reinterpret_cast<int4*>(plans + F->first_tile + local);
    if ((act >> lane) & 1u) {
        int4* __restrict__ o = out + 5 * (1 + __popc(act & ((1u << lane) - 1u)));
#pragma unroll
        for (int q = 0; q < 5; ++q) o[q] = e[q];
    }
    if (lane == 0) {
        const unsigned long long p0 = F->out_plane[0], p1 = F->out_plane[1], p2 = F->out_plane[2];
        out[0] = make_int4(__popc(act), frame, g.x0, g.y0);
        out[1] = make_int4(F->width, F->height, F->format, F->flags);
        out[2] = make_int4((int)(unsigned)p0, (int)(p0 >> 32), (int)(unsigned)p1, (int)(p1 >> 32));
        out[3] = make_int4((int)(unsigned)p2, (int)(p2 >> 32), F->out_stride[0], F->out_stride[1]);
        out[4] = make_int4(F->out_stride[2], 0, 0, 0);
    }
}

}  // namespace svb

// ---- pre-pass: coordinate tables of every separable YUV layer of every frame of the batch, and the plan of every tile ---
// grid (table_blocks * max layers + plan_blocks, 1, frames), 256 threads; blockIdx.z = frame.
//   blockIdx.x <  table_blocks * max layers: layer blockIdx.x / table_blocks, one table entry (two words) per thread, blocks
//                                            padded to whole tiles
//   beyond:                                  one tile of the frame per warp, see plan_tile
extern "C" __global__ void __launch_bounds__(256) svb_mix_tables(const SvbFrameDesc* __restrict__ frames, uint32_t* __restrict__ tables, int* __restrict__ tile_counter,
                                                                SvbTilePlan* __restrict__ plans, int table_blocks, int max_layers) {
    using namespace svb;
    if ((blockIdx.x | blockIdx.z | threadIdx.x) == 0) *tile_counter = 0;  // svb_mix_tiled claims its tiles from it
    const SvbFrameDesc* __restrict__ F = frames + blockIdx.z;
    if ((int)blockIdx.x >= table_blocks * max_layers) {
        const int local = ((int)blockIdx.x - table_blocks * max_layers) * 8 + (int)(threadIdx.x >> 5);
        if (local < F->tiles_x * F->tiles_y) plan_tile(F, (int)blockIdx.z, local, tables, plans, (int)(threadIdx.x & 31));
        return;
    }
    const int l = (int)blockIdx.x / table_blocks;
    if (l >= F->nlayers) return;
    const SvbLayerDesc* __restrict__ L = &F->layers[l];
    if (!(L->flags & SVB_LAYER_SEPARABLE)) return;
    const bool yuv = L->format == SVB_NV12 || L->format == SVB_Y420P;  // BGRA / RGBA layers only use the luma-resolution entries
    const int W = F->width, H = F->height;
    const int ncy = F->tiles_x * SVB_TILE_W, ncc = ncy / 2, nry = F->tiles_y * SVB_TILE_H, nrc = nry / 2;
    int e = ((int)blockIdx.x % table_blocks) * (int)blockDim.x + (int)threadIdx.x;
    if (e >= ncy + ncc + nry + nrc) return;
    uint32_t* __restrict__ base = tables + F->table_base + (size_t)l * (size_t)(F->tiles_x * SVB_TAB_COL_WORDS + F->tiles_y * SVB_TAB_ROW_WORDS);
    uint32_t* __restrict__ rbase = base + F->tiles_x * SVB_TAB_COL_WORDS;
    if (e < ncy) {
        const Ent t = ent_col_y(L, W, e);
        base[col_y_word(e)] = __float_as_uint(t.a), base[col_y_word(e) + SVB_TILE_W] = pack_ent(t);
    } else if ((e -= ncy) < ncc) {
        if (!yuv) return;
        const Ent t = ent_col_c(L, W, e);
        base[col_c_word(e)] = __float_as_uint(t.a), base[col_c_word(e) + SVB_TILE_W / 2] = pack_ent(t);
    } else if ((e -= ncc) < nry) {
        const Ent t = ent_row_y(L, H, e);
        rbase[row_y_word(e)] = __float_as_uint(t.a), rbase[row_y_word(e) + 1] = pack_ent(t);
    } else if (yuv) {
        e -= nry;
        const Ent t = ent_row_c(L, H, e);
        rbase[row_c_word(e)] = __float_as_uint(t.a), rbase[row_c_word(e) + 1] = pack_ent(t);
    }
}

// ---- the compositor --------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(SVB_TILED_THREADS, SVB_TILED_MIN_CTAS)
    svb_mix_tiled(const SvbFrameDesc* __restrict__ frames, const SvbTilePlan* __restrict__ plans, int total_tiles, float one, int* __restrict__ tile_counter, int box_y_bytes,
                  int box_c_bytes) {
    using namespace svb;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TiledSmem& sm = *reinterpret_cast<TiledSmem*>(smem_raw);
    uint8_t* const boxes = smem_raw + SVB_TILED_FIXED_BYTES;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int fenced = -1;  // frame whose tensor maps this CTA's issuing thread has acquired
    unsigned phase0 = 0, phase1 = 0, pphase = 0;
    if (t == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        mbar_init(&sm.pbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int cur = 0;
    int stage = 0;        // box buffer that holds (or is about to receive) the next staged layer to consume
    bool primed = false;  // this tile's first staged layer was already put in flight by the previous tile
    // Tiles are claimed two ahead by one lane (a global atomic whose latency nobody waits for); the same lane fetches the plan
    // of the CTA's next tile while the current one is computed.
    const bool claimer = t == SVB_CLAIM_WARP * 32;
    int nclaimed = 0;
    auto claim = [&]() -> int {
        if (SVB_DYNAMIC_TILES) return atomicAdd(tile_counter, 1);
        return (int)blockIdx.x + (int)gridDim.x * nclaimed++;
    };
    auto fetch_plan = [&](int tile_index, int slot) {
        mbar_expect_tx(&sm.pbar, (unsigned)sizeof(SvbTilePlan));
        bulk_load(&sm.plan[slot], plans + tile_index, (unsigned)sizeof(SvbTilePlan), &sm.pbar);
    };
    __syncthreads();  // the mbarriers are initialised
    if (claimer) {
        const int t0 = claim();
        sm.tile_idx[0] = t0;
        sm.tile_idx[1] = claim();
        if (t0 < total_tiles) fetch_plan(t0, 0);
    }
    __syncthreads();
    int tile = sm.tile_idx[0];
    int k3 = 0;  // ordinal of the tile in this CTA's sequence, modulo 3
    if (tile < total_tiles) {
        mbar_wait(&sm.pbar, pphase);
        pphase ^= 1;
    }

    while (tile < total_tiles) {
        const int4(*plan)[5] = reinterpret_cast<const int4(*)[5]>(&sm.plan[cur]) + 1;  // plan[i] = the i-th layer that touches this tile
        const int4* const hdr = reinterpret_cast<const int4*>(&sm.plan[cur]);
        const int4 h0 = hdr[0], h1 = hdr[1];
        const int nact = h0.x, x0 = h0.z, y0 = h0.w, W = h1.x, H = h1.y;
        const SvbFrameDesc* __restrict__ F = frames + h0.y;
        const int xt = x0 + 2 * lane, yt = y0 + 4 * warp;  // this thread's columns xt, xt+1, xt+64, xt+65 x rows yt..yt+3
        const bool live = xt < W && yt < H;                // W and H even are planner preconditions
        const bool live1 = xt + SVB_TILE_W / 2 < W;         // the second column pair is inside the frame

        // TMA of one staged layer into buffer `b`: the source boxes and the tile's blocks of the coordinate tables, all
        // addresses and sizes precomputed by the planner
        auto issue = [&](const int4(*pl)[5], int i, int b) {
            const int4 q0 = pl[i][0], q1 = pl[i][1], q2 = pl[i][2], q3 = pl[i][3], q4 = pl[i][4];
            auto ptr = [](int lo, int hi) { return (const void*)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo); };
            if (q4.w != fenced) {  // (only after the context's tensor-map table has wrapped) acquire a frame's maps once per CTA
                const SvbFrameDesc* __restrict__ TF = frames + q4.w;
                if (TF->flags & SVB_FRAME_TMAP_FENCE)
                    for (int q = 0; q < TF->nlayers; ++q)
                        if (TF->layers[q].flags & SVB_LAYER_STAGED) {
                            tmap_acquire((const void*)TF->layers[q].tmap[0]);
                            tmap_acquire((const void*)TF->layers[q].tmap[1]);
                            if (TF->layers[q].format != SVB_NV12) tmap_acquire((const void*)TF->layers[q].tmap[2]);
                        }
                fenced = q4.w;
            }
            uint8_t* const by = boxes + b * box_y_bytes;
            uint8_t* const bc = boxes + 2 * box_y_bytes + b * box_c_bytes;
            mbar_expect_tx(&sm.bar[b], q4.z);
            tma_load_2d(by, ptr(q2.x, q2.y), q0.y, q0.z, &sm.bar[b]);
            tma_load_2d(bc, ptr(q2.z, q2.w), q0.w, q1.x, &sm.bar[b]);
            if (q3.x | q3.y) tma_load_2d(bc + box_c_bytes / 2, ptr(q3.x, q3.y), q0.w, q1.x, &sm.bar[b]);
            bulk_load(sm.tabs[b], ptr(q3.z, q3.w), SVB_TAB_COL_WORDS * 4, &sm.bar[b]);
            bulk_load(sm.tabs[b] + SVB_TAB_COL_WORDS, ptr(q4.x, q4.y), SVB_TAB_ROW_WORDS * 4, &sm.bar[b]);
        };
        // first staged layer of `pl` from entry `from` on, or -1
        auto first_staged = [&](const int4(*pl)[5], int from, int n) {
            for (int i = from; i < n; ++i)
                if ((pl[i][0].x & 0xff) >= PLAN_STAGED) return i;
            return -1;
        };
        if (!primed && t == 0) {
            const int i = first_staged(plan, 0, nact);
            if (i >= 0) issue(plan, i, stage);
        }
        primed = false;
        // this CTA's next tile: its plan is fetched now (one bulk copy), and the tile after it is claimed
        const int k3n = k3 == 2 ? 0 : k3 + 1;  // slot of the next tile's index; the slot after it receives the new claim
        const int next_tile = sm.tile_idx[k3n];
        const bool has_next = next_tile < total_tiles;
        if (claimer) {
            if (has_next) fetch_plan(next_tile, cur ^ 1);
            sm.tile_idx[k3n == 2 ? 0 : k3n + 1] = claim();
        }

        // ---- running picture: integer-valued floats ------------------------------------------------------------
        float2 Yi[4][2], Ui[2], Vi[2];  // Yi[r][p] = columns (2p, 2p+1) of row yt+r; Ui/Vi[k] = the two chroma texels of chroma row yt/2+k
#pragma unroll
        for (int r = 0; r < 4; ++r) Yi[r][0] = Yi[r][1] = splat(0.f);  // img_clear_*: Y = 0, chroma = 0.5 -> 128
        Ui[0] = Ui[1] = Vi[0] = Vi[1] = splat(128.f);
        if ((h1.w & SVB_FRAME_LOAD_CUR) && live) {
            const int4 h2 = hdr[2], h3 = hdr[3], h4 = hdr[4];
            const uint8_t* const oY = (const uint8_t*)(((unsigned long long)(unsigned)h2.y << 32) | (unsigned)h2.x);
            const uint8_t* const oU = (const uint8_t*)(((unsigned long long)(unsigned)h2.w << 32) | (unsigned)h2.z);
            const uint8_t* const oV = (const uint8_t*)(((unsigned long long)(unsigned)h3.y << 32) | (unsigned)h3.x);
            const int sY = h3.z, sU = h3.w, sV = h4.x;
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (yt + r < H) {
                    const uint8_t* row = oY + (size_t)(yt + r) * sY + xt;
                    const unsigned w0 = *(const unsigned short*)row, w1 = live1 ? *(const unsigned short*)(row + SVB_TILE_W / 2) : 0u;
                    Yi[r][0] = bytes2(opaque(w0 & 0xff), opaque(w0 >> 8)), Yi[r][1] = bytes2(opaque(w1 & 0xff), opaque(w1 >> 8));
                }
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (yt + 2 * k < H) {
                    if (h1.z == SVB_NV12) {  // chroma texel xt/2 = the (U, V) byte pair at byte xt of the row; texel xt/2 + 32 is 64 bytes on
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
            const int4 p0 = plan[i][0];
            const int mode = p0.x & 0xff;
            if (mode >= PLAN_STAGED) {
                const int4 p1 = plan[i][1];
                const int jc0 = p1.x;
                __syncthreads();  // every warp is past its reads of the other buffer
                if (t == 0) {  // refill the other buffer: the next staged layer of this tile, else the first one of the next tile
                    const int j = first_staged(plan, i + 1, nact);
                    if (j >= 0) {
                        issue(plan, j, stage ^ 1);
                    } else if (has_next) {
                        mbar_wait(&sm.pbar, pphase);  // the next tile's plan has landed (it was requested a whole tile ago)
                        const int4(*pn)[5] = reinterpret_cast<const int4(*)[5]>(&sm.plan[cur ^ 1]) + 1;
                        const int jn = first_staged(pn, 0, sm.plan[cur ^ 1].e[0][0][0]);
                        if (jn >= 0) {
                            issue(pn, jn, stage ^ 1);
                            primed = true;  // only this thread looks at it
                        }
                    }
                }
                if (stage == 0) mbar_wait(&sm.bar[0], phase0), phase0 ^= 1;
                else mbar_wait(&sm.bar[1], phase1), phase1 ^= 1;
                if (live) {
                    const int fmt = p1.y >> 8, lflags = p1.y & 0xff, box_w = p1.z & 0xffff, box_cw = p1.z >> 16;
                    const int pitchC = fmt == SVB_NV12 ? box_cw * 2 : box_cw, stepC = fmt == SVB_NV12 ? 2 : 1;
                    const unsigned bY = smem_u32(boxes + stage * box_y_bytes), bU = smem_u32(boxes + 2 * box_y_bytes + stage * box_c_bytes);
                    const unsigned bV = bU + (fmt == SVB_NV12 ? 1 : box_c_bytes / 2);
                    const float alpha = __int_as_float(p1.w);
                    FillTerms ft;
                    if (mode == PLAN_STAGED_EDGE || !(lflags & SVB_LAYER_OPACITY_01)) {
                        const SvbLayerDesc* __restrict__ L = &F->layers[p0.x >> 8];
                        const float4 fc = ldrow(L->u.fillColor, 0);
                        const float3 fill = rgb2yuv(fc.x, fc.y, fc.z);
                        const float af = mul(alpha, fc.w);
                        ft.fy = splat(fill.x), ft.fu = splat(fill.y), ft.fv = splat(fill.z), ft.af = splat(af), ft.naf = splat(sub(1.f, af));
                        // Does this warp's block hold a sample inside the border rectangle but outside the picture (fill), or a row that is
                        // so along its whole length?  Without a border or letterbox no tile does, and the lean edge body applies.
                        const uint32_t* __restrict__ tb = sm.tabs[stage];
                        auto odd = [](uint32_t p) { const unsigned ok = p >> 17; return (ok & 1u) != 0u && ok != 7u; };
                        bool mixed = odd(tb[SVB_TILE_W + 2 * lane]) || odd(tb[SVB_TILE_W + 2 * lane + 1]) || odd(tb[SVB_TILE_W + 64 + 2 * lane]) ||
                                     odd(tb[SVB_TILE_W + 65 + 2 * lane]) || odd(tb[2 * SVB_TILE_W + SVB_TILE_W / 2 + lane]) || odd(tb[2 * SVB_TILE_W + SVB_TILE_W / 2 + 32 + lane]);
#pragma unroll
                        for (int r = 0; r < 4; ++r) mixed = mixed || odd(tb[SVB_TAB_COL_WORDS + 2 * (4 * warp + r) + 1]);
#pragma unroll
                        for (int k = 0; k < 2; ++k) mixed = mixed || odd(tb[SVB_TAB_COL_WORDS + 2 * SVB_TILE_H + 2 * (2 * warp + k) + 1]);
                        const bool lean = (lflags & SVB_LAYER_OPACITY_01) && !__any_sync(__activemask(), mixed);
                        // (both inlined: called out of line, with the running picture through a local array, the edge bodies cost 8 % of
                        // the whole kernel although edge tiles are one layer-tile in ten -- profiles/r1_history.md)
                        if (lean) fast_layer<2, true, false>(bY, bU, bV, sm.tabs[stage], lane, warp, p0.y, p0.z, p0.w, jc0, box_w, pitchC, stepC, alpha, one, ft, Yi, Ui, Vi);
                        else fast_layer<3, true, false>(bY, bU, bV, sm.tabs[stage], lane, warp, p0.y, p0.z, p0.w, jc0, box_w, pitchC, stepC, alpha, one, ft, Yi, Ui, Vi);
                    } else if (lflags & SVB_LAYER_UNIT_OPACITY) {
                        if (fmt == SVB_NV12) fast_layer<0, true, true>(bY, bU, bV, sm.tabs[stage], lane, warp, p0.y, p0.z, p0.w, jc0, box_w, pitchC, stepC, alpha, one, ft, Yi, Ui, Vi);
                        else fast_layer<0, true, false>(bY, bU, bV, sm.tabs[stage], lane, warp, p0.y, p0.z, p0.w, jc0, box_w, pitchC, stepC, alpha, one, ft, Yi, Ui, Vi);
                    } else {
                        if (fmt == SVB_NV12) fast_layer<1, true, true>(bY, bU, bV, sm.tabs[stage], lane, warp, p0.y, p0.z, p0.w, jc0, box_w, pitchC, stepC, alpha, one, ft, Yi, Ui, Vi);
                        else fast_layer<1, true, false>(bY, bU, bV, sm.tabs[stage], lane, warp, p0.y, p0.z, p0.w, jc0, box_w, pitchC, stepC, alpha, one, ft, Yi, Ui, Vi);
                    }
                }
                stage ^= 1;
            } else if (live) {  // PLAN_GENERIC / PLAN_TABLE_RGBA: no staging, out of line
                float st[24];
                state_to_array(Yi, Ui, Vi, st);
                if (mode == PLAN_TABLE_RGBA) {
                    const int4 p3 = plan[i][3], p4 = plan[i][4];
                    rgba_table_layer(&F->layers[p0.x >> 8], (const uint32_t*)(((unsigned long long)(unsigned)p3.w << 32) | (unsigned)p3.z),
                                     (const uint32_t*)(((unsigned long long)(unsigned)p4.y << 32) | (unsigned)p4.x), xt, yt, W, H, st, warp);
                } else {
                    generic_layer(&F->layers[p0.x >> 8], xt, yt, (float)W, (float)H, W, H, st);
                }
                array_to_state(st, Yi, Ui, Vi);
            }
        }

        if (live) {
            const int4 h2 = hdr[2], h3 = hdr[3], h4 = hdr[4];
            uint8_t* const oY = (uint8_t*)(((unsigned long long)(unsigned)h2.y << 32) | (unsigned)h2.x);
            uint8_t* const oU = (uint8_t*)(((unsigned long long)(unsigned)h2.w << 32) | (unsigned)h2.z);
            uint8_t* const oV = (uint8_t*)(((unsigned long long)(unsigned)h3.y << 32) | (unsigned)h3.x);
            const int sY = h3.z, sU = h3.w, sV = h4.x;
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (yt + r < H) {
                    uint8_t* row = oY + (size_t)(yt + r) * sY + xt;
                    *(unsigned short*)row = pack2(Yi[r][0].x, Yi[r][0].y);
                    if (live1) *(unsigned short*)(row + SVB_TILE_W / 2) = pack2(Yi[r][1].x, Yi[r][1].y);
                }
#pragma unroll
            for (int k = 0; k < 2; ++k)
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
        if (has_next) {  // every thread sees the next plan's arrival for itself before anyone re-arms the barrier
            mbar_wait(&sm.pbar, pphase);
            pphase ^= 1;
        }
        __syncthreads();  // the boxes and this tile's plan slot are free
        tile = next_tile, k3 = k3n, cur ^= 1;
    }
}
