// mix_video.h -- VideoMixer, mirroring /root/reference/Sources/SwiftVideo/mix.video.swift:21-184.
//
// What is kept: the constructor's parameters that matter to compositing (:22-30), the ingest rule of the
// closure (:57-75: a sample of another asset is stored by revision, the mixer's own samples pass through),
// the two-generation sample store (:104-107,114), z-ordering (:115), the backing ring of 10 GPU targets
// (:148-167), findKernel's name construction (:142-146) with its error behaviour, clear-then-fold (:116-125),
// the emitted sample (:127-131).  What is not rebuilt: the Clock/Bus/Source plumbing around it (SURVEY.md
// section 2.1 #11, out of scope) -- the caller drives mix(at:) instead of a WallClock timer.
//
// What is ours: the fold is one fused launch (svb_mix_tiled / svb_mix_generic) instead of L+1 launches,
// 2L allocations and 2L synchronous copies; several mixers of one GPU can be folded into
// the same launch (mixMany); waiting for completion is optional.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "compute.h"
#include "svb_desc.h"

namespace svb {

class VideoMixer {
  public:
    enum class Mode : int {
        fused = 0,     // svb_mix_ring where the tile preconditions hold, else the generic fused kernel
        perLayer = 1,  // the reference's own sequence: clear kernel + one applyComputeImage per layer
        generic = 2,   // the generic fused kernel only
        fusedGather = 3,  // fused, with svb_mix_gather (taps through the texture unit) wherever every staged layer can be bound as a texture
        fusedTiled = 4,   // fused, with svb_mix_tiled (the CTA-per-tile TMA compositor of round 1)
        fusedRing = 5     // fused, with svb_mix_ring (tiles planned and staged per CTA, eight free-running warps, three-stage ring)
    };

    VideoMixer(const ComputeContext* computeContext, Vector2 outputSize, PixelFormat outputFormat = PixelFormat::nv12,
               const std::string& assetId = "", const std::string& workspaceId = "", int64_t frameDuration = 1000,
               int64_t timescale = 30000, int64_t epoch = 0);

    const std::string& assetId() const { return idAsset; }
    const std::string& workspaceId() const { return idWorkspace; }
    ComputeContext* computeContext() { return hasContext ? &clContext : nullptr; }
    void setMode(Mode m) { mode = m; }

    // The body of the ingest closure (mix.video.swift:57-75).  Returns true when the sample was stored as a
    // layer (`.nothing`), false when it is the mixer's own asset and simply passes through (`.just`).
    bool push(const PictureSample& pic);
    bool push(std::shared_ptr<const PictureSample> pic);

    // mix(at:) (:95-140): compose the stored layers into the next backing image and return the emitted
    // sample.  `time` is the tick time in `timescale` units.  wait=true is upstream's endComputePass(ctx, true).
    PictureSample mix(int64_t time, bool wait = true);

    // Fold the ticks of several mixers that share one GPU context into one launch.
    static void mixMany(VideoMixer* const* mixers, int n, int64_t time, PictureSample* outs, bool wait);

    // Low-level fold used by mix(): clear + layers (already z-sorted) into `target`, uniforms given.
    static ComputeContext composeRaw(const ComputeContext& ctx, const PictureSample& target,
                                     const std::vector<const PictureSample*>& layers, const ImageUniforms* uniforms, Mode mode);

    int64_t frameDuration, timescale, epoch;

  private:
    struct Tick {
        PictureSample backing;
        std::vector<std::shared_ptr<const PictureSample>> images;  // z-sorted, kept alive until the launch is queued
    };
    Tick beginTick();                                                          // :113-115
    ComputeKernel findKernel(const PictureSample* image, const PictureSample& target) const;  // :142-146
    PictureSample getBacking();                                                // :148-165
    void endTick();                                                            // :104-107

    static const int numberBackingImages = 10;  // :167
    std::vector<PictureSample> backing;
    int currentBacking = 0;
    PixelFormat backingFormat;
    Vector2 backingSize;
    ComputeContext clContext;
    bool hasContext = false;
    std::map<std::string, std::shared_ptr<const PictureSample>> samples[2];
    // what was derived from a sample the last time it was composed (its uniforms and its planned layer descriptor), by identity; an
    // entry lives while its sample does and is seen every few ticks
    struct Planned {
        std::weak_ptr<const PictureSample> who;
        ImageUniforms uniforms;
        SvbLayerDesc desc;
        unsigned tick = 0;
    };
    std::map<const PictureSample*, Planned> planned;
    unsigned tickNo = 0;
    std::string idAsset, idWorkspace;
    Mode mode = Mode::fused;
};

// Device time of the fused launches (CUDA events around each kernel on the compute stream), for the roofline.
void setLaunchTiming(const ComputeContext& ctx, bool on);
// The fused compositor keeps the coordinate tables of a batch whose geometry (frame sizes, layer uniforms, source sizes and formats) did
// not change since the batch's buffer was last filled, and launches its table pre-pass only otherwise (default: on).
void setTableCache(const ComputeContext& ctx, bool on);
void readLaunchTiming(const ComputeContext& ctx, double* totalMs, unsigned long long* launches);
// Host time spent inside the fused compose calls since timing was enabled (plan + driver calls: what the caller's thread pays per tick)
void readHostTiming(const ComputeContext& ctx, double* totalMs, unsigned long long* calls, double* waitMs = nullptr);

// Planner: fills frame descriptors (one per pass of 16 layers) for one target.  Exposed for tests.
struct FramePlan {
    std::vector<SvbFrameDesc> passes;
    bool tiledOk = false;
};
// `cached` (optional, one entry per layer, nullptr = plan it): a layer descriptor planned earlier for the SAME sample on a target of the
// same size -- samples are immutable, so a mixer that sees a sample again (a still picture, a layer that persists one extra tick) skips
// the tensor-map / texture lookups and the rectangle arithmetic for it.
FramePlan planFrame(const ComputeContext& ctx, const PictureSample& target, const std::vector<const PictureSample*>& layers,
                    const ImageUniforms* uniforms, const SvbLayerDesc* const* cached = nullptr);

}  // namespace svb
