// ring_layout.h -- shared-memory layout of svb_mix_ring (kernels_ring.cuh), shared with the host launcher (mix_video.cpp).
#pragma once
#include "svb_desc.h"

// Plan of one tile in shared memory (written by the planning warp): a 16-byte header and SVB_RPLAN_REC_BYTES per layer that touches
// the tile, bottom to top.
//   header  [1] = (width, height, format | frame flags << 8, strideY)   [2] = (&Y, &U)   [3] = (&V, strideU, strideV)   and [0] =
//           (n | staged mask << 16, x0 | y0 << 16, frame, bit0: the first listed layer overwrites every sample without reading it; n = 0xffff: no more tiles)
//   layer   the consumers' part, written ready to use (so that a consumer warp decodes nothing per layer):
//           [0] = (address of the staged luma box minus the footprint's origin, the same of the chroma box, address of the stage's table
//                  blocks, address of the stage's `full` mbarrier | 1 << 30 | the parity to wait for << 31)      -- all zero for a layer that is not staged
//           [1] = (opacity bits, chroma texel step | U-to-V distance << 8, the body of each of the tile's eight units: four bits each (SVB_BODY_*), word 0 of [2])
//           the producer's part:
//           [2] = (mode | layer << 8 | format << 16 | layer flags << 20, ix0 | jy0 << 16, ic0 | jc0 << 16, opacity bits)
//           [3] = (bytes the stage receives, pitchY | pitchC << 16, first column block, first row block)   -- word offsets into the table buffer
//           [4] = (&tensor map Y, &tensor map C)   [5] = (&tensor map V, column-block bytes, row-block bytes)
//           [6] = (flags of the tile's two unit columns: byte each, flags of its four unit rows: byte each, 0, 0)   -- SVB_UREC_* | SVB_RREC_TOUCH
#define SVB_RPLAN_HDR_BYTES 64
#define SVB_RPLAN_REC_BYTES 112
enum { SVB_BODY_NONE = 0, SVB_BODY_BLEND = 1, SVB_BODY_BLEND_HALF = 2, SVB_BODY_OPAQUE = 3, SVB_BODY_OPAQUE_HALF = 4, SVB_BODY_EDGE_LEAN = 5, SVB_BODY_EDGE = 6 };
#define SVB_RREC_TOUCH 0x80u  // the layer's rectangle reaches into this unit column / unit row
#define SVB_RPLAN_SLOT_BYTES(layers) ((SVB_RPLAN_HDR_BYTES + SVB_RPLAN_REC_BYTES * (layers) + 127) / 128 * 128)
#ifndef SVB_RING_STAGES
#define SVB_RING_STAGES 3
#endif
#define SVB_RING_WARPS 8
#define SVB_RING_THREADS (32 * (SVB_RING_WARPS + 1))  // eight consumer warps and the producer warp
#define SVB_RING_PLANS 3
#define SVB_RING_TAB_BYTES (2 * SVB_UCOL_WORDS * 4 + 4 * SVB_UROW_WORDS * 4)  // a stage's table blocks: two unit columns, four unit rows
#define SVB_RING_HDR_BYTES 128
#define SVB_RING_SMEM_BYTES(boxY, boxC, layers) \
    (SVB_RING_HDR_BYTES + SVB_RING_WARPS * SVB_STRIP_STATE_BYTES + SVB_RING_PLANS * SVB_RPLAN_SLOT_BYTES(layers) + SVB_RING_STAGES * ((boxY) + (boxC) + SVB_RING_TAB_BYTES))

