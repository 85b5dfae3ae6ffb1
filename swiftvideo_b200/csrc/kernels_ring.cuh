// kernels_ring.cuh -- svb_mix_ring: the fused compositor's fast path, third design (round 2).
//
// svb_mix_tiled (round 1) amortises planning and staging over a CTA -- one plan, one set of TMA copies per 128x32 tile and layer --
// but its eight warps meet at a CTA barrier in front of every layer (20 % of its stall samples) and its layer bodies are large.
// A second design of this round made the warp autonomous -- every warp planned, staged and composited its own 64x8 unit with the
// compact loops of kernels_strip.cuh, no barrier at all -- but paid planning, table copies and TMA issue per unit: 43 % of its
// instructions were overhead, and it only drew level with svb_mix_tiled (profiles/r2_history.md).  This kernel takes the halves that worked:
//   * a CTA of eight warps owns a 128x32 tile: ONE warp plans it (lane = layer, from the column / row records svb_strip_tables
//     leaves), ONE elected lane issues its copies -- per tile and layer two or three TMA tensor copies of the source footprint and two
//     bulk copies of the table blocks;
//   * each warp composites its own 64x8 unit of the tile with the layer bodies of kernels_strip.cuh, at its own pace: the staged layers go
//     through a ring of THREE stages with a `full` mbarrier (the copies' bytes) and an `empty` mbarrier (eight warp arrivals) per
//     stage, so a warp waits for data, never for another warp, and the issuing lane only waits for a warp that lags by more than a layer;
//   * the ring runs across tiles: the plan of the next tile is ready (its own pair of mbarriers) long before this tile's last
//     layers are computed, so the first layers of the next tile are in flight by then.
// Per-sample arithmetic: strip_layer / strip_layer_edge (kernels_strip.cuh), i.e. fast_layer's, operation for operation.
#pragma once
#include "kernels_strip.cuh"

#ifndef SVB_RING_MIN_CTAS
#define SVB_RING_MIN_CTAS 3
#endif

namespace svb {

__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_u32(unsigned bar, unsigned parity, unsigned hint_ns) {  // one try, sleeping up to hint_ns for the phase
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
#ifndef SVB_RING_WAIT_HINT
#define SVB_RING_WAIT_HINT 2000u
#endif
__device__ __forceinline__ void mbar_wait_u32(unsigned bar, unsigned parity) {
#ifdef SVB_RING_LEAN_WAIT
    // the whole loop in PTX: try-wait, branch -- no counter, no trap (three instructions a round instead of six)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SVB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra SVB_DONE;\n"
        "bra SVB_WAIT;\n"
        "SVB_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity), "r"(SVB_RING_WAIT_HINT)
        : "memory");
    return;
#endif
    unsigned ok;
    unsigned spin = 0;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"  // (suspend-time hint: a waiting warp sleeps instead of spinning through issue slots)
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(SVB_RING_WAIT_HINT)
            : "memory");
        if (++spin > (1u << 22)) __trap();  // a copy that never lands must not hang the GPU
#ifdef SVB_RING_WAIT_SLEEP
        if (!ok) __nanosleep(SVB_RING_WAIT_SLEEP);
#endif
    } while (!ok);
}

}  // namespace svb

#include "ring_layout.h"

extern "C" __global__ void __launch_bounds__(SVB_RING_THREADS, SVB_RING_MIN_CTAS)
    svb_mix_ring(const SvbFrameDesc* __restrict__ frames, const uint32_t* __restrict__ tables, int nframes, int total_tiles, float one, int* __restrict__ tile_counter, int box_y_bytes,
                 int box_c_bytes, int plan_slot_bytes) {
    using namespace svb;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // mbarriers: full[3] (copies landed), empty[3] (eight warps have left the stage), pfull[3] (plan written), pempty[3] (eight warps have left the plan)
    const unsigned bars = smem_u32(smem_raw);
    const unsigned full0 = bars, empty0 = full0 + 8u * SVB_RING_STAGES, pfull0 = empty0 + 8u * SVB_RING_STAGES, pempty0 = pfull0 + 8u * SVB_RING_PLANS;
    const unsigned stage_bytes = (unsigned)(box_y_bytes + box_c_bytes) + SVB_RING_TAB_BYTES;
    unsigned char* const my_state = smem_raw + SVB_RING_HDR_BYTES + (size_t)(warp & 7) * SVB_STRIP_STATE_BYTES;
    float2* const sY = reinterpret_cast<float2*>(my_state);  // [12 rows][SVB_STATE_PITCH_F2: 32 lanes + padding]: luma pairs of rows 0..7, then (U, V) of chroma rows 0..3
    const unsigned state = smem_u32(my_state) + 8u * lane;
    const unsigned plan0 = smem_u32(smem_raw + SVB_RING_HDR_BYTES + SVB_RING_WARPS * SVB_STRIP_STATE_BYTES);
    const unsigned stage0 = plan0 + (unsigned)SVB_RING_PLANS * (unsigned)plan_slot_bytes;
    const unsigned tab_off = (unsigned)(box_y_bytes + box_c_bytes);
    if (threadIdx.x == 0) {
        for (int i = 0; i < SVB_RING_STAGES; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8u * i), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8u * i), "r"(SVB_RING_WARPS));
        }
        for (int i = 0; i < SVB_RING_PLANS; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pfull0 + 8u * i), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pempty0 + 8u * i), "r"(SVB_RING_WARPS));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // the only CTA-wide barrier of the kernel
    const int firstA = lane < nframes ? frames[lane].first_tile : 0x7fffffff, firstB = lane + 32 < nframes ? frames[lane + 32].first_tile : 0x7fffffff;
    auto sts4 = [](unsigned a, const uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); };

    // ---- the plan of tile t (t >= total_tiles: the end marker) into plan slot s: the planning warp, lane = layer ----------------
    auto plan_tile = [&](int t, int s, unsigned gbase) -> unsigned {  // gbase: staged layers planned before this tile; returns those of this tile
        const unsigned slot_a = plan0 + (unsigned)s * (unsigned)plan_slot_bytes;
        if (t >= total_tiles) {
            if (lane == 0) sts4(slot_a, make_uint4(0xffffu, 0u, 0u, 0u));
            return 0u;
        }
        const int f = __popc(__ballot_sync(0xffffffffu, t >= firstA)) + __popc(__ballot_sync(0xffffffffu, t >= firstB)) - 1;
        const int first = f < 32 ? __shfl_sync(0xffffffffu, firstA, f) : __shfl_sync(0xffffffffu, firstB, f - 32);
        const SvbFrameDesc* __restrict__ F = frames + f;
        const int W = F->width, H = F->height, nl = F->nlayers, local = t - first;
        const int tx_n = (W + 2 * SVB_UNIT_W - 1) / (2 * SVB_UNIT_W), ux_n = F->tiles_x, uy_n = F->tiles_y;  // (tiles_x / tiles_y hold the unit counts)
        const int ty = local / tx_n, tx = local - ty * tx_n, x0 = tx * 2 * SVB_UNIT_W, y0 = ty * 4 * SVB_UNIT_H;
        const int ncol = min(2, ux_n - 2 * tx), nrow = min(4, uy_n - 4 * ty);
        unsigned mode = PLAN_SKIP;
        bool covers = false, inner = false;
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, r2 = r0, r3 = r0, r4 = r0;
        unsigned relY = 0, relC = 0, c1y = 0, codes = 0;  // the consumers' part of the record (below)
        if (lane < nl) {
            // every load of the plan is issued here, before anything is looked at (one trip to L2, not six): the layer's constants and
            // the records of the tile's two unit columns and four unit rows (indices clamped: an absent unit is masked below)
            const uint4* __restrict__ pc = reinterpret_cast<const uint4*>(&F->layers[lane].pc);
            const size_t lw = (size_t)F->table_base + (size_t)lane * strip_layer_words(F);
            const uint4* __restrict__ rec = reinterpret_cast<const uint4*>(tables + lw + (size_t)ux_n * SVB_UCOL_WORDS + (size_t)uy_n * SVB_UROW_WORDS);
            const uint4 c0 = __ldg(pc), c1 = __ldg(pc + 1), c2 = __ldg(pc + 2), c3 = __ldg(pc + 3);
            uint4 q[6];
#pragma unroll
            for (int k = 0; k < 2; ++k) q[k] = __ldg(rec + min(2 * tx + k, ux_n - 1));
#pragma unroll
            for (int k = 0; k < 4; ++k) q[2 + k] = __ldg(rec + ux_n + min(4 * ty + k, uy_n - 1));
            const unsigned fmt = c2.y & 0xfu, lflags = (c2.y >> 4) & 0xffu;
            if ((int)c3.x < x0 + 2 * SVB_UNIT_W && (int)c3.z > x0 && (int)c3.y < y0 + 4 * SVB_UNIT_H && (int)c3.w > y0) {  // the layer's rectangle touches the tile
                r0.w = c2.x;
                r1.z = (unsigned)lw + (unsigned)(2 * tx) * SVB_UCOL_WORDS, r1.w = (unsigned)lw + (unsigned)ux_n * SVB_UCOL_WORDS + (unsigned)(4 * ty) * SVB_UROW_WORDS;
                r3.z = (unsigned)ncol * SVB_UCOL_WORDS * 4u, r3.w = (unsigned)nrow * SVB_UROW_WORDS * 4u;
                if (!(lflags & SVB_LAYER_SEPARABLE)) {
                    mode = PLAN_GENERIC;
                } else if (fmt != SVB_NV12 && fmt != SVB_Y420P) {
                    mode = PLAN_TABLE_RGBA;
                } else {
                    // first / last source index over the tile's unit columns and unit rows; their flags
                    unsigned ix0 = 0xffffu, ic0 = 0xffffu, ix1 = 0u, ic1 = 0u, jy0 = 0xffffu, jc0 = 0xffffu, jy1 = 0u, jc1 = 0u, all = ~0u, cf = 0u, rf = 0u;
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        if (k < ncol) {
                            const bool touch = (int)c3.x < x0 + (k + 1) * SVB_UNIT_W && (int)c3.z > x0 + k * SVB_UNIT_W;
                            ix0 = min(ix0, q[k].x & 0xffffu), ic0 = min(ic0, q[k].x >> 16), ix1 = max(ix1, q[k].y & 0xffffu), ic1 = max(ic1, q[k].y >> 16);
                            all &= q[k].z;
                            cf |= ((q[k].z & 0x7fu) | (touch ? SVB_RREC_TOUCH : 0u)) << (8 * k);
                        }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < nrow) {
                            const bool touch = (int)c3.y < y0 + (k + 1) * SVB_UNIT_H && (int)c3.w > y0 + k * SVB_UNIT_H;
                            jy0 = min(jy0, q[2 + k].x & 0xffffu), jc0 = min(jc0, q[2 + k].x >> 16), jy1 = max(jy1, q[2 + k].y & 0xffffu), jc1 = max(jc1, q[2 + k].y >> 16);
                            all &= q[2 + k].z | SVB_UREC_XFREE;
                            rf |= ((q[2 + k].z & 0x7fu) | (touch ? SVB_RREC_TOUCH : 0u)) << (8 * k);
                        }
                    const unsigned box_w = c1.w & 0xffffu, box_h = (c2.y >> 12) & 0x3ffu, box_ch = c2.y >> 22;
                    const unsigned box_cw = (c1.w >> 16) / (fmt == SVB_NV12 ? 2u : 1u);
                    // (column records hold origins already rounded down for TMA; the footprint must fit the tile-sized boxes)
                    const bool fits = (lflags & SVB_LAYER_STAGED) && ix1 - ix0 < box_w && jy1 - jy0 < box_h && ic1 - ic0 < box_cw && jc1 - jc0 < box_ch;
                    mode = fits ? PLAN_STAGED : PLAN_GENERIC;
                    covers = (all & SVB_UREC_FULL) && (lflags & SVB_LAYER_UNIT_OPACITY);
                    inner = fits && (all & SVB_UREC_XFREE);  // every unit of the tile takes the interior body
                    r0.y = ix0 | (jy0 << 16), r0.z = ic0 | (jc0 << 16);
                    r1.x = c1.z + r3.z + r3.w, r1.y = c1.w;
                    r2 = c0;
                    r3.x = c1.x, r3.y = c1.y;
                    r4 = make_uint4(cf, rf, 0u, 0u);
                    if (fits) {
                        // what a consumer warp would otherwise derive per layer and unit: the staged boxes' addresses minus the footprint's
                        // origin (relative to the stage), and the body each of the tile's eight units takes (four bits per unit, SVB_BODY_*)
                        const unsigned pitchY = c1.w & 0xffffu, pitchC = c1.w >> 16, cstep = fmt == SVB_NV12 ? 2u : 1u;
                        relY = 0u - ix0 - jy0 * pitchY, relC = (unsigned)box_y_bytes - ic0 * cstep - jc0 * pitchC;
                        c1y = cstep | ((fmt == SVB_NV12 ? 1u : (unsigned)box_c_bytes / 2u) << 8);
                        const bool op01 = (lflags & SVB_LAYER_OPACITY_01) != 0, unit = (lflags & SVB_LAYER_UNIT_OPACITY) != 0;
#pragma unroll
                        for (int u = 0; u < SVB_RING_WARPS; ++u) {
                            const unsigned cfl = (cf >> (8 * (u & 1))) & 0xffu, rfl = (rf >> (8 * (u >> 1))) & 0xffu, both = cfl & rfl;
                            unsigned code = SVB_BODY_NONE;
                            if (both & SVB_RREC_TOUCH) {
                                if (!((both & SVB_UREC_FULL) && (cfl & SVB_UREC_XFREE)) || !op01) code = op01 && !((cfl | rfl) & SVB_UREC_MIXED) ? SVB_BODY_EDGE_LEAN : SVB_BODY_EDGE;
                                else code = (unit ? SVB_BODY_OPAQUE : SVB_BODY_BLEND) + ((both & SVB_UREC_HALF) ? 1u : 0u);
                            }
                            codes |= code << (4 * u);
                        }
                    }
                }
                r0.x = mode | ((unsigned)lane << 8) | (fmt << 16) | (lflags << 20);
            }
        }
        unsigned act = __ballot_sync(0xffffffffu, mode != PLAN_SKIP);
        const unsigned cov = __ballot_sync(0xffffffffu, covers), stg = __ballot_sync(0xffffffffu, mode >= PLAN_STAGED), inn = __ballot_sync(0xffffffffu, inner);
        unsigned first_covers = 0;
        if (cov) {
            const unsigned top = 31u - (unsigned)__clz(cov);
            act &= ~((1u << top) - 1u);  // drop what the topmost covering layer hides
            first_covers = (inn >> top) & 1u;
        }
        const unsigned below = act & ((1u << lane) - 1u);
        if ((act >> lane) & 1u) {
            const unsigned a = slot_a + SVB_RPLAN_HDR_BYTES + SVB_RPLAN_REC_BYTES * (unsigned)__popc(below);
            uint4 k0 = make_uint4(0u, 0u, 0u, 0u);
            if (mode >= PLAN_STAGED) {  // staged layer number gs of this CTA lives in stage gs % 3 and its barriers are in their (gs / 3)-th use
                const unsigned gs = gbase + (unsigned)__popc(below & stg), b = gs % SVB_RING_STAGES, bY = stage0 + b * stage_bytes;
                k0 = make_uint4(bY + relY, bY + relC, bY + tab_off, (full0 + 8u * b) | 0x40000000u | (((gs / SVB_RING_STAGES) & 1u) << 31));
            }
            sts4(a, k0), sts4(a + 16, make_uint4(r0.w, c1y, codes, r0.x));
            sts4(a + 32, r0), sts4(a + 48, r1), sts4(a + 64, r2), sts4(a + 80, r3), sts4(a + 96, r4);
        }
        const unsigned smask = __reduce_or_sync(0xffffffffu, ((act & stg) >> lane) & 1u ? 1u << __popc(below) : 0u);
        if (lane == 0) {
            const unsigned long long p0 = F->out_plane[0], p1 = F->out_plane[1], p2 = F->out_plane[2];
            sts4(slot_a, make_uint4((unsigned)__popc(act) | (smask << 16), (unsigned)x0 | ((unsigned)y0 << 16), (unsigned)f, first_covers));
            sts4(slot_a + 16, make_uint4((unsigned)W, (unsigned)H, (unsigned)F->format | ((unsigned)F->flags << 8), (unsigned)F->out_stride[0]));
            sts4(slot_a + 32, make_uint4((unsigned)p0, (unsigned)(p0 >> 32), (unsigned)p1, (unsigned)(p1 >> 32)));
            sts4(slot_a + 48, make_uint4((unsigned)p2, (unsigned)(p2 >> 32), (unsigned)F->out_stride[1], (unsigned)F->out_stride[2]));
        }
        return (unsigned)__popc(act & stg);
    };
    // the async copies of the staged layer whose plan record lies at shared-memory address ra, into stage b (one elected lane of the issuing warp)
    auto issue = [&](unsigned ra, unsigned b) {
        const uint4 r0 = lds_u4(ra + 32), r1 = lds_u4(ra + 48), r2 = lds_u4(ra + 64), r3 = lds_u4(ra + 80);
        if (elect_one()) {
            const unsigned dst = stage0 + b * stage_bytes, mb = full0 + 8u * b;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(r1.x) : "memory");
            auto tma = [&](unsigned d, unsigned long long tmap, unsigned x, unsigned y) {
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(d), "l"(tmap), "r"(x), "r"(y), "r"(mb)
                             : "memory");
            };
            auto bulk = [&](unsigned d, const void* src, unsigned bytes) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src), "r"(bytes), "r"(mb) : "memory");
            };
            tma(dst, ((unsigned long long)r2.y << 32) | r2.x, r0.y & 0xffffu, r0.y >> 16);
            tma(dst + (unsigned)box_y_bytes, ((unsigned long long)r2.w << 32) | r2.z, r0.z & 0xffffu, r0.z >> 16);
            if (((r0.x >> 16) & 0xfu) != SVB_NV12) tma(dst + (unsigned)box_y_bytes + (unsigned)box_c_bytes / 2u, ((unsigned long long)r3.y << 32) | r3.x, r0.z & 0xffffu, r0.z >> 16);
            bulk(dst + tab_off, tables + r1.z, r3.z);
            bulk(dst + tab_off + 2u * SVB_UCOL_WORDS * 4u, tables + r1.w, r3.w);
        }
        __syncwarp();
    };

    // Warp 8 is the producer: it claims tiles, plans them (three plan slots: up to three tiles ahead of the slowest consumer) and issues
    // the copies of their staged layers in order, as far ahead as the ring allows.  Warps 0..7 are the consumers, one 64x8 unit each.
    // `g` counts the staged layers of this CTA: staged layer g lives in stage g % 3 and its barriers are in their (g / 3)-th use -- the
    // producer and every consumer count alike because they read the same plans, so nobody tracks barrier phases.  Tile n of this CTA
    // has its plan in slot n % 3.
    if (warp == SVB_RING_WARPS) {
        // The producer serves two queues and never blocks on one while the other has work: a stage that the consumers have left is
        // refilled at once (first priority: a consumer may be waiting for those bytes), and whenever a plan slot is free the next tile
        // is claimed and planned -- up to three tiles ahead, so a consumer never waits for a plan.
        unsigned gi = 0;       // staged layers issued
        unsigned planned = 0;  // tiles planned
        unsigned gplan = 0;    // staged layers planned
        unsigned n = 0, i = 0, smask = 0;  // the tile whose staged layers are being issued, its next listed layer, the staged layers left (bit 0 = layer i)
        bool have_tile = false, last_planned = false;
        for (unsigned idle = 0;;) {
            if (!have_tile && planned > n) {
                const unsigned hx = lds_u4(plan0 + (n % SVB_RING_PLANS) * (unsigned)plan_slot_bytes).x;
                if ((hx & 0xffffu) == 0xffffu) {  // the end marker: this CTA claims no more.  The last CTA to get here leaves the batch's
                    // counters at zero, so that the next launch on this buffer can start without the pre-pass (mix_video.cpp: tabSig)
                    if (lane == 0 && atomicAdd(tile_counter + 1, 1) == (int)gridDim.x - 1) tile_counter[0] = 0, tile_counter[1] = 0;
                    return;
                }
                smask = hx >> 16, i = 0, have_tile = true;
            }
            if (have_tile) {
                if (!smask) {
                    ++n, have_tile = false;
                    continue;
                }
                const unsigned skip = (unsigned)__ffs(smask) - 1u;
                smask >>= skip, i += skip;
            }
            bool did = false;
            const unsigned b = gi % SVB_RING_STAGES;
            if (have_tile && (gi < SVB_RING_STAGES || mbar_try_u32(empty0 + 8u * b, ((gi / SVB_RING_STAGES) - 1u) & 1u, 0u))) {  // every consumer has left the stage's previous tenant
                issue(plan0 + (n % SVB_RING_PLANS) * (unsigned)plan_slot_bytes + SVB_RPLAN_HDR_BYTES + SVB_RPLAN_REC_BYTES * i, b);
                ++gi, smask >>= 1, ++i, did = true;
            } else if (!last_planned && planned < n + SVB_RING_PLANS) {
                const unsigned slot = planned % SVB_RING_PLANS;
                if (planned < SVB_RING_PLANS || mbar_try_u32(pempty0 + 8u * slot, ((planned / SVB_RING_PLANS) - 1u) & 1u, 0u)) {  // every consumer has left the slot's previous plan
                    int t = 0;
                    if (lane == 0) t = atomicAdd(tile_counter, 1);
                    t = __shfl_sync(0xffffffffu, t, 0);
                    gplan += plan_tile(t, (int)slot, gplan);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(pfull0 + 8u * slot);
                    ++planned, did = true, last_planned = t >= total_tiles;
                }
            }
            if (did) {
                idle = 0;
            } else {  // nothing to do right now: sleep on the barrier that matters most
                if (have_tile) mbar_try_u32(empty0 + 8u * b, ((gi / SVB_RING_STAGES) - 1u) & 1u, 1000u);
                else mbar_try_u32(pempty0 + 8u * (planned % SVB_RING_PLANS), ((planned / SVB_RING_PLANS) - 1u) & 1u, 1000u);
                if (++idle > (1u << 22)) __trap();  // a barrier that never completes must not hang the GPU
            }
        }
    }
    // (per-warp constants of the layer loop: this warp's unit of a tile and where its table blocks lie inside a stage)
    const int ux = warp & 1, uy = warp >> 1;
    const unsigned tabc_off = (unsigned)ux * SVB_UCOL_WORDS * 4u, tabr_off = 2u * SVB_UCOL_WORDS * 4u + (unsigned)uy * SVB_UROW_WORDS * 4u, code_shift = 4u * (unsigned)warp;
    for (unsigned n = 0;; ++n) {  // tile ordinal of this CTA
        const unsigned slot = n % SVB_RING_PLANS, plan = plan0 + slot * (unsigned)plan_slot_bytes;
        mbar_wait_u32(pfull0 + 8u * slot, (n / SVB_RING_PLANS) & 1u);  // this tile's plan
        const uint4 h0 = lds_u4(plan);
        const unsigned nact = h0.x & 0xffffu;
        if (nact == 0xffffu) break;
        const int x0 = (int)(h0.y & 0xffffu), y0 = (int)(h0.y >> 16), f = (int)h0.z;
        const SvbFrameDesc* __restrict__ F = frames + f;
        const uint4 h1 = lds_u4(plan + 16);
        const int W = (int)h1.x, H = (int)h1.y, ofmt = (int)(h1.z & 0xffu), fflags = (int)(h1.z >> 8);
#define SVB_RING_TARGET()                                                                    \
    const uint4 h2 = lds_u4(plan + 32), h3 = lds_u4(plan + 48);                              \
    const int sYb = (int)lds_u1(plan + 28), sUb = (int)h3.z, sVb = (int)h3.w;                \
    uint8_t* const oY = (uint8_t*)(((unsigned long long)h2.y << 32) | h2.x);                 \
    uint8_t* const oU = (uint8_t*)(((unsigned long long)h2.w << 32) | h2.z);                 \
    uint8_t* const oV = (uint8_t*)(((unsigned long long)h3.y << 32) | h3.x);
        const int xt = x0 + ux * SVB_UNIT_W + 2 * lane, yt = y0 + uy * SVB_UNIT_H;  // this lane's columns xt, xt+1 x rows yt .. yt+7
        const bool live = xt < W && yt < H;                                      // W and H even are planner preconditions

        // ---- running picture: img_clear_* (Y = 0, chroma = 0.5 -> 128), or the target's bytes when an earlier pass left them ----
        if (fflags & SVB_FRAME_LOAD_CUR) {
            SVB_RING_TARGET()
#pragma unroll
            for (int r = 0; r < SVB_UNIT_H; ++r) {
                unsigned w0 = 0;
                if (live && yt + r < H) w0 = *(const unsigned short*)(oY + (size_t)(yt + r) * sYb + xt);
                sY[r * SVB_STATE_PITCH_F2 + lane] = bytes2(opaque(w0 & 0xff), opaque(w0 >> 8));
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
                sY[(SVB_UNIT_H + k) * SVB_STATE_PITCH_F2 + lane] = bytes2(opaque(cu), opaque(cv));
            }
        } else if (!(h0.w & 1u)) {  // (bit 0: the first listed layer overwrites every sample without reading it)
#pragma unroll
            for (int r = 0; r < SVB_UNIT_H; ++r) sY[r * SVB_STATE_PITCH_F2 + lane] = splat(0.f);
#pragma unroll
            for (int k = 0; k < SVB_UNIT_H / 2; ++k) sY[(SVB_UNIT_H + k) * SVB_STATE_PITCH_F2 + lane] = splat(128.f);
        }

#pragma unroll 1
        for (unsigned i = 0; i < nact; ++i) {
            const unsigned ra = plan + SVB_RPLAN_HDR_BYTES + SVB_RPLAN_REC_BYTES * i;
            const uint4 k0 = lds_u4(ra), k1 = lds_u4(ra + 16);  // the consumers' part of the record: everything below is ready to use (ring_layout.h)
            if (k0.w) {                                         // a staged layer
                const unsigned code = (k1.z >> code_shift) & 0xfu, mb = k0.w & 0x00ffffffu;
                // (a warp the layer does not reach waits as well: an arrival on `empty` is only in the right phase once the stage's copies were issued)
                mbar_wait_u32(mb, k0.w >> 31);
                if (code != SVB_BODY_NONE) {  // the layer reaches into this warp's unit
                    const unsigned colY = k0.x, colC = k0.y, tabc = k0.z + tabc_off, tabr = k0.z + tabr_off, cstep = k1.y & 0xffu, vofs = k1.y >> 8;
                    const float alpha = __uint_as_float(k1.x);
                    if (code == SVB_BODY_BLEND) strip_layer<false, false>(colY, colC, cstep, vofs, tabc, tabr, lane, alpha, one, state);
                    else if (code == SVB_BODY_OPAQUE) strip_layer<true, false>(colY, colC, cstep, vofs, tabc, tabr, lane, alpha, one, state);
                    else if (code == SVB_BODY_OPAQUE_HALF) strip_layer<true, true>(colY, colC, cstep, vofs, tabc, tabr, lane, alpha, one, state);
                    else if (code == SVB_BODY_BLEND_HALF) strip_layer<false, true>(colY, colC, cstep, vofs, tabc, tabr, lane, alpha, one, state);
                    else {
                        const SvbLayerDesc* __restrict__ L = &F->layers[(k1.w >> 8) & 0xffu];
                        const float4 fc = ldrow(L->u.fillColor, 0);
                        const float3 fl = rgb2yuv(fc.x, fc.y, fc.z);
                        const float af = mul(alpha, fc.w);
                        // lean: no sample of the unit lies inside the border rectangle but outside the picture (without a border or letterbox: none ever does)
                        if (code == SVB_BODY_EDGE_LEAN) strip_layer_edge<true>(colY, colC, cstep, vofs, tabc, tabr, lane, alpha, one, make_float4(fl.x, fl.y, fl.z, 0.f), af, sY);
                        else strip_layer_edge<false>(colY, colC, cstep, vofs, tabc, tabr, lane, alpha, one, make_float4(fl.x, fl.y, fl.z, 0.f), af, sY);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(mb + 8u * SVB_RING_STAGES);  // this warp has left the stage (`empty` lies SVB_RING_STAGES barriers behind `full`)
            } else {
                const uint4 r1 = lds_u4(ra + 48);
                const SvbLayerDesc* __restrict__ L = &F->layers[(k1.w >> 8) & 0xffu];
                float* const py = reinterpret_cast<float*>(sY + lane);
                const uint32_t* __restrict__ colblk = tables + r1.z + ux * SVB_UCOL_WORDS;
                const uint32_t* __restrict__ rowblk = tables + r1.w + uy * SVB_UROW_WORDS;
                if (yt < H) {
                    if ((k1.w & 0xffu) == PLAN_TABLE_RGBA) strip_rgba_layer(L, colblk, rowblk, lane, xt, yt, W, H, py, py + SVB_STATE_PITCH_F * SVB_UNIT_H);
                    else strip_generic_layer(L, xt, yt, W, H, py, py + SVB_STATE_PITCH_F * SVB_UNIT_H);
                }
            }
        }

        // ---- the unit's bytes.  A whole unit of an NV12 target whose planes allow 16-byte stores leaves as TWO stores per warp: the
        // state is read transposed -- a lane takes 16 consecutive samples of one row (rows are padded so that these reads spread over
        // all banks, svb_desc.h) -- so that one st.v4 covers the eight luma rows (64 bytes each) and a second one, of the lower half
        // warp, the four chroma rows.  Anything else (a unit cut by the picture's edge, planar chroma, odd alignment): two bytes per
        // lane and row.
        __syncwarp();
        if (live) {
            SVB_RING_TARGET()
            const float2 ONE = splat(one);
            const int xu = x0 + ux * SVB_UNIT_W;
            if (ofmt == SVB_NV12 && xu + SVB_UNIT_W <= W && yt + SVB_UNIT_H <= H && ((h2.x | h2.z | (unsigned)sYb | (unsigned)sUb) & 15u) == 0u) {
                auto pack4 = [&](const float4 v) {  // four integer-valued floats in 0..255 -> four bytes: + 2^23 leaves them in the low mantissa bits
                    const float2 a = add2<true>(make_float2(v.x, v.y), splat(8388608.f), ONE), b = add2<true>(make_float2(v.z, v.w), splat(8388608.f), ONE);
                    return __byte_perm(__byte_perm(__float_as_uint(a.x), __float_as_uint(a.y), 0x0040), __byte_perm(__float_as_uint(b.x), __float_as_uint(b.y), 0x0040), 0x5410);
                };
                auto row16 = [&](unsigned a) {
                    const float4 v0 = lds_f4(a), v1 = lds_f4(a + 16u), v2 = lds_f4(a + 32u), v3 = lds_f4(a + 48u);
                    return make_uint4(pack4(v0), pack4(v1), pack4(v2), pack4(v3));
                };
                const unsigned sbase = smem_u32(my_state);
                {
                    const int r = lane & 7, seg = lane >> 3;
                    *reinterpret_cast<uint4*>(oY + (size_t)(yt + r) * sYb + xu + 16 * seg) = row16(sbase + (unsigned)r * SVB_STATE_PITCH_B + 64u * (unsigned)seg);
                }
                if (lane < 16) {
                    const int r = lane & 3, seg = lane >> 2;
                    *reinterpret_cast<uint4*>(oU + (size_t)((yt >> 1) + r) * sUb + xu + 16 * seg) = row16(sbase + (unsigned)(SVB_UNIT_H + r) * SVB_STATE_PITCH_B + 64u * (unsigned)seg);
                }
            } else {
                auto pack = [&](float2 v) {  // two integer-valued floats in 0..255 -> two bytes
                    const float2 x = add2<true>(v, splat(8388608.f), ONE);
                    return (unsigned short)__byte_perm(__float_as_uint(x.x), __float_as_uint(x.y), 0x0040);
                };
                uint8_t* pY = oY + (size_t)yt * sYb + xt;
                const int nrow = min(SVB_UNIT_H, H - yt);
#pragma unroll 1
                for (int r = 0; r < nrow; ++r, pY += sYb) *(unsigned short*)pY = pack(sY[r * SVB_STATE_PITCH_F2 + lane]);
                if (ofmt == SVB_NV12) {
                    uint8_t* pC = oU + (size_t)(yt >> 1) * sUb + xt;
#pragma unroll 1
                    for (int k = 0; 2 * k < nrow; ++k, pC += sUb) *(unsigned short*)pC = pack(sY[(SVB_UNIT_H + k) * SVB_STATE_PITCH_F2 + lane]);
                } else {
                    uint8_t* pU = oU + (size_t)(yt >> 1) * sUb + (xt >> 1);
                    uint8_t* pV = oV + (size_t)(yt >> 1) * sVb + (xt >> 1);
#pragma unroll 1
                    for (int k = 0; 2 * k < nrow; ++k, pU += sUb, pV += sVb) {
                        const unsigned short p = pack(sY[(SVB_UNIT_H + k) * SVB_STATE_PITCH_F2 + lane]);
                        *pU = (uint8_t)(p & 0xff), *pV = (uint8_t)(p >> 8);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(pempty0 + 8u * slot);  // this warp has left the tile's plan
    }
#undef SVB_RING_TARGET
}
