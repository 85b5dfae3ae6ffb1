// svb_desc.h -- frame/layer descriptors shared by the host planner (mix_video.cpp) and the fused kernels.
// Plain C structs: compiled by g++ and nvcc alike.
#pragma once
#include <stdint.h>

#define SVB_MAX_LAYERS 16

// tile of the tiled kernel (luma pixels); chroma is TILE_W/2 x TILE_H/2
#define SVB_TILE_W 128
#ifndef SVB_TILE_H
#define SVB_TILE_H 32
#endif
#define SVB_TILED_COMPUTE_WARPS (SVB_TILE_H / 4)              // a warp covers 128 columns x 4 rows
#define SVB_TILED_THREADS (SVB_TILED_COMPUTE_WARPS * 32)
// svb_mix_gather: luma rows of a tile one warp owns (4 or 2: fewer rows = less state per thread = more warps in flight to cover
// the texture latency); a work unit is one such strip of one tile
#ifndef SVB_GATHER_ROWS
#define SVB_GATHER_ROWS 4
#endif
#define SVB_GATHER_STRIPS (SVB_TILE_H / SVB_GATHER_ROWS)
// largest source footprint the tiled kernel stages in shared memory per tile and layer
#ifndef SVB_BOX_Y_BYTES
#define SVB_BOX_Y_BYTES (SVB_TILE_H * 640)
#endif
#ifndef SVB_BOX_C_BYTES
#define SVB_BOX_C_BYTES (SVB_TILE_H * 384)
#endif

// 0..3: the formats the reference has kernels for.  4..6: sources the reference's PixelFormat names (sample.pict.swift:22-27) but has no kernel
// for -- composed by the generic and per-layer kernels: NV21 = NV12 with (Cr, Cb) pairs, Y422P / Y444P = three planes with
// chroma at half width / full size.
enum SvbFormat { SVB_NV12 = 0, SVB_Y420P = 1, SVB_BGRA = 2, SVB_RGBA = 3, SVB_NV21 = 4, SVB_Y422P = 5, SVB_Y444P = 6 };
#define SVB_FORMAT_IS_YUV(f) ((f) == SVB_NV12 || (f) == SVB_Y420P || (f) >= SVB_NV21)
#define SVB_FORMAT_IS_SEMIPLANAR(f) ((f) == SVB_NV12 || (f) == SVB_NV21)
#define SVB_FORMAT_CHROMA_W(f, w) ((f) == SVB_Y444P ? (w) : (w) / 2)
#define SVB_FORMAT_CHROMA_H(f, h) ((f) == SVB_Y444P || (f) == SVB_Y422P ? (h) : (h) / 2)

enum {
    SVB_FRAME_LOAD_CUR = 1,   // continue an earlier pass: start from the target's bytes, not from clear
    SVB_FRAME_SCALAR_FP = 2,  // tuning aid: spell the packed fp32x2 arithmetic as scalar instructions (same results)
    SVB_FRAME_GATHER = 4,     // the batch goes to svb_mix_gather: plan separable YUV layers for the texture path (no staging limits)
    SVB_FRAME_RING = 16,      // the batch goes to svb_mix_ring (unit-blocked tables, tiles planned in the compositor)
    SVB_FRAME_TMAP_FENCE = 8  // tensor-map slots of the context's table have been rewritten: acquire the frame's maps before the first copy
};
enum {
    SVB_LAYER_SEPARABLE = 1,     // x outputs depend only on x and y outputs only on y (no rotation/shear)
    SVB_LAYER_UNIT_OPACITY = 2,  // opacity == 1: cur*(1-1) + v*1 == v exactly, the blend is skipped
    SVB_LAYER_STAGED = 4,        // tensor maps below are valid: source footprints are staged by TMA
    SVB_LAYER_OPACITY_01 = 8,    // 0 <= opacity <= 1: blended values cannot leave [0,1], store clamps are no-ops
    SVB_LAYER_TEX = 16           // tex[] below holds texture objects over the source planes (svb_mix_gather)
};

// ImageUniforms as uploaded by applyComputeImage (reference compute.swift:76-86; device mirror
// kernels.cl.swift:49-59): 236 bytes of payload, padded to 240 so the float4 rows stay 16-byte aligned.
typedef struct __attribute__((aligned(16))) SvbUniforms {
    float transform[16];
    float textureTx[16];
    float borderMatrix[16];
    float fillColor[4];
    float inSize[2];
    float outSize[2];
    float opacity;
    float sampleTime;
    float targetTime;
    float pad_;
} SvbUniforms;

// What the producer warp of svb_mix_ring needs of a layer to plan and stage a tile, gathered in 64 bytes (four 16-byte loads by the planning lane)
typedef struct __attribute__((aligned(16))) SvbStripConsts {
    uint32_t stmapY[2], stmapC[2];                 // device addresses of the tensor maps (luma, chroma / U)
    uint32_t stmapV[2], stx_bytes, pitches;        // (V plane of a planar source); bytes of the staged boxes; box pitch luma | chroma << 16, in bytes
    uint32_t opacity_bits, fmtflags, tab, rec;     // format | SVB_LAYER_* << 4 | box_h << 12 | box_ch << 22; word offsets into the batch's table buffer: the layer's tables, its column records (row records follow)
    int32_t rect[4];                               // = SvbLayerDesc::rect
} SvbStripConsts;

typedef struct __attribute__((aligned(64))) SvbLayerDesc {
    // Tensor maps live in a table in device memory that belongs to the context (mix_video.cpp): a slot is written once, before
    // the first launch that uses it, and never rewritten, so no tensormap-proxy fence stands in front of the copies (the profile of
    // round 2 had a sixth of all stall samples at those fences).  Here: the slots' device addresses.
    unsigned long long tmap[3];  // CUtensorMap per source plane (valid when SVB_LAYER_STAGED): boxes of svb_mix_tiled
    SvbUniforms u;               // 240
    unsigned long long plane[3]; // device pointers
    int32_t stride[3];
    int32_t width, height;       // luma / RGBA size; chroma planes are (width/2, height/2)
    int32_t format;
    int32_t flags;               // SVB_LAYER_*
    int32_t rect[4];             // x0,y0,x1,y1: outside this canvas rectangle the layer touches nothing
    int32_t box_w, box_h;        // staged luma box (elements); chroma box is box_cw x box_ch
    int32_t box_cw, box_ch;
    int32_t pad0_;
    unsigned long long tex[3];   // CUtexObject per source plane (valid when SVB_LAYER_TEX): UNORM8, clamp, unnormalised coordinates
    int32_t pad_[16];
    SvbStripConsts pc;           // svb_mix_ring: filled by the host at launch (mix_video.cpp: launchFrames)
} SvbLayerDesc;

typedef struct __attribute__((aligned(64))) SvbFrameDesc {
    unsigned long long out_plane[3];
    int32_t out_stride[3];
    int32_t width, height;
    int32_t format;
    int32_t nlayers;
    int32_t flags;       // SVB_FRAME_*
    int32_t first_tile;  // prefix sum of tiles over the batch
    int32_t tiles_x, tiles_y;
    int32_t table_base;  // first word of this frame's coordinate tables in the batch's table buffer (svb_mix_tables)
    int32_t pad_[4];
    SvbLayerDesc layers[SVB_MAX_LAYERS];
} SvbFrameDesc;

// Coordinate tables of one layer of a WxH frame (svb_mix_tables), in 4-byte words.  An entry is a pair
// (a, p): a = the fractional weight of the i1 tap (fp32), p = i0 | (i1 - i0) << 16 | ok << 17.  The entries are laid
// out in tile-sized blocks, so that a tile's slice is ONE contiguous bulk copy per axis and the per-lane reads of
// the compositor are conflict-free in shared memory:
//   column block (one per tile column): aY[TILE_W] pY[TILE_W] aC[TILE_W/2] pC[TILE_W/2]
//   row block    (one per tile row)   : (aY,pY)[TILE_H] (aC,pC)[TILE_H/2]
// Blocks are padded to whole tiles with copies of the last valid entry.
#define SVB_TAB_COL_WORDS (3 * SVB_TILE_W)
#define SVB_TAB_ROW_WORDS (3 * SVB_TILE_H)
#define SVB_TILES_X(W) (((W) + SVB_TILE_W - 1) / SVB_TILE_W)
#define SVB_TILES_Y(H) (((H) + SVB_TILE_H - 1) / SVB_TILE_H)
#define SVB_TABLE_WORDS(W, H) (SVB_TILES_X(W) * SVB_TAB_COL_WORDS + SVB_TILES_Y(H) * SVB_TAB_ROW_WORDS)

// ---- svb_mix_ring: a warp owns a 64x8 unit (a lane: two adjacent luma columns x 8 rows and the chroma texel column under
// them).  Tables of one layer (svb_strip_tables), blocked by unit so that a unit's slices arrive in the warp's shared memory by
// two bulk copies:
//   column block (one per unit column, 192 words): aY[64] pY[64] aC[32] pC[32]   (a = weight of the i1 tap, p as above)
//   row block    (one per unit row, 64 words)    : 12 entries of 4 words, luma rows 0..7 then chroma rows 0..3:
//                                                  b, 1-b, j0*pitch, j1*pitch      (pitch = the layer's staged box pitch in bytes:
//                                                  a tap's shared-memory address is a per-lane column term plus the row's offset)
//                                                  then 12 words  ok | (j1-j0) << 3 | j0 << 4  (edge and RGBA bodies), then one word: bit r set = row r
//                                                  starts from a source row that is not the lower source row of row r-1 (SVB_UROW_RELOAD_WORD), 3 words of padding
//   column records (one per unit column, 4 words) and row records (one per unit row, 4 words): what a plan needs of the
//   blocks -- (first source index luma | chroma << 16, last source index luma | chroma << 16, SVB_UREC_* flags, 0) -- so that a unit
//   (or a tile of units) is planned from two 16-byte loads per layer (the plan is separable: a unit is inside a picture iff its column range and its row range are).
#define SVB_UNIT_W 64
#define SVB_UNIT_H 8
#define SVB_UCOL_WORDS (3 * SVB_UNIT_W)
#define SVB_UROW_WORDS 64
#define SVB_UROW_RELOAD_WORD 60
#define SVB_UNITS_X(W) (((W) + SVB_UNIT_W - 1) / SVB_UNIT_W)
#define SVB_UNITS_Y(H) (((H) + SVB_UNIT_H - 1) / SVB_UNIT_H)
// words of one layer's tables: column blocks, row blocks, column records, row records
#define SVB_UTABLE_WORDS(W, H) (SVB_UNITS_X(W) * (SVB_UCOL_WORDS + 4) + SVB_UNITS_Y(H) * (SVB_UROW_WORDS + 4))
enum {
    SVB_UREC_FULL = 1,   // every entry of the block has border, tx and uv inside [0,1]: the range lies inside the picture
    SVB_UREC_XFREE = 2,  // (columns) no tap is clamped: i1 == i0 + 1 throughout
    SVB_UREC_FITS = 4,   // the footprint of the range fits the layer's staged box along this axis
    SVB_UREC_HALF = 8,   // every weight of the block is exactly 1/2: the four bilinear weights are 1/4 each
    SVB_UREC_MIXED = 16  // the block holds an entry inside the border rectangle but outside the picture (a fill sample)
};
// a unit's running picture as floats: 12 rows (8 luma, 4 chroma) of 32 lanes x 8 bytes, rows 272 bytes apart -- 16 bytes of padding per
// row turn the rows by four banks each, so that the epilogue's transposed reads (a lane fetches 16 consecutive floats of ONE row, eight
// rows per quarter-warp) are free of bank conflicts while the layer bodies' row-wise accesses stay so
#define SVB_STATE_PITCH_F 68                         // floats from a state row to the next
#define SVB_STATE_PITCH_F2 (SVB_STATE_PITCH_F / 2)   // the same in (pair) slots
#define SVB_STATE_PITCH_B (SVB_STATE_PITCH_F * 4)    // and in bytes
#define SVB_STRIP_STATE_BYTES ((SVB_UNIT_H + SVB_UNIT_H / 2) * SVB_STATE_PITCH_B)

// Plan of one tile, written by svb_mix_plan and fetched by the compositor's CTAs with one bulk copy.  Five 16-byte words
// per entry; entry 0 is the header, entries 1..n the layers that touch the tile, bottom to top (kernels_tiled.cuh):
//   header  [0] = (n, frame, x0, y0)  [1] = (width, height, format, frame flags)  [2] = (&Y, &U)  [3] = (&V, strideY, strideU)
//           [4] = (strideV, 0, 0, 0)
typedef struct __attribute__((aligned(16))) SvbTilePlan {
    int32_t e[1 + SVB_MAX_LAYERS][5][4];
} SvbTilePlan;

// dynamic shared memory of svb_mix_tiled: a fixed part (two table slices, two tile plans, mbarriers, the tile ring) and
// behind it two luma boxes and two chroma boxes whose size the host picks PER LAUNCH from the largest staged footprint
// of the batch (box_y_bytes / box_c_bytes kernel arguments, multiples of 256, at most SVB_BOX_*_BYTES): shared memory
// not taken stays L1, and the headline workload needs 11 KB per stage, not the 32 KB worst case.
#define SVB_TILED_FIXED_BYTES ((2 * (SVB_TAB_COL_WORDS + SVB_TAB_ROW_WORDS) * 4 + 2 * (1 + SVB_MAX_LAYERS) * 80 + 64 + 127) / 128 * 128)
#define SVB_TILED_SMEM_BYTES(boxY, boxC) (SVB_TILED_FIXED_BYTES + 2 * (boxY) + 2 * (boxC))
#define SVB_TILED_SMEM_MAX SVB_TILED_SMEM_BYTES(SVB_BOX_Y_BYTES, SVB_BOX_C_BYTES)

// output tile of svb_scale_convert and the most filter taps per axis it takes (Lanczos-3 down to 1 : 2.66, bilinear down to 1 : 8)
#define SVB_SCALE_TW 64
#ifndef SVB_SCALE_TH
#define SVB_SCALE_TH 16  // 16 rows: 47 KB of shared memory per CTA at 4K -> 1080p Lanczos (4 CTAs per SM); 32 rows halve the halo but leave 2 CTAs per SM (92 vs 66 us)
#endif
// row pitches (floats) of the transposed source windows when the tile's vertical footprint allows the compile-time ones: multiples of 4
// (16-byte aligned columns) whose quarter is odd (columns spread over the banks).  44 = 16 rows x 2 + 12 taps: 2 : 1 Lanczos-3.
// pitch (floats) of the horizontally filtered rows: the tile width + 4, so that the two row groups a warp of the horizontal pass writes
// (rows 4 apart) land 16 banks apart
#define SVB_SCALE_HP (SVB_SCALE_TW + 4)
#define SVB_SCALE_PITCH_Y 44
#define SVB_SCALE_PITCH_C 28
#define SVB_SCALE_MAX_TAPS 16  // tap counts up to this are compiled per count; larger ones loop at run time

// svb_scale_convert (kernels_scale.cuh): NV12 / P010 -> BGRA with a separable resize, passed by value as the kernel argument.
// Tables per axis and plane kind (Y = luma plane, C = chroma plane): first[dstN] = first source index of each output
// column / row (may lie outside the plane: indices are clamped when sampling), w[dstN * n] = its n tap weights.
typedef struct SvbScaleDesc {
    unsigned long long srcY, srcC, dst;
    unsigned long long fYx, wYx, fYy, wYy, fCx, wCx, fCy, wCy;
    int32_t strideY, strideC, dstStride;
    int32_t srcW, srcH, dstW, dstH;
    int32_t format;          // 0 NV12, 1 P010 (10 bits in the MSBs of little-endian 16-bit words)
    int32_t nYx, nYy, nCx, nCy;
    int32_t spanYy, spanCy;  // most source rows (luma / chroma) one tileH-row output tile reaches: shared-memory sizing
    int32_t spanYx, spanCx;  // most source columns (chunk-aligned) one 64-column output tile stages
    int32_t pitchY, pitchC;  // row pitch (floats) of the transposed windows: a multiple of 4 whose quarter is odd
    int32_t tileH;           // output rows per CTA (<= SVB_SCALE_TH)
    int32_t vecY, vecC, vecDst;  // plane base and stride allow 16-byte accesses (else sample-by-sample staging / 4-byte stores)
} SvbScaleDesc;

#ifdef __cplusplus
static_assert(sizeof(SvbUniforms) == 240, "SvbUniforms layout");
static_assert(sizeof(SvbLayerDesc) % 64 == 0, "SvbLayerDesc alignment");
static_assert(sizeof(SvbStripConsts) == 64 && sizeof(SvbLayerDesc) == 512, "SvbStripConsts / SvbLayerDesc layout");
static_assert(sizeof(SvbFrameDesc) % 64 == 0, "SvbFrameDesc alignment");
static_assert(sizeof(SvbTilePlan) == (1 + SVB_MAX_LAYERS) * 80, "SvbTilePlan layout");
#endif
