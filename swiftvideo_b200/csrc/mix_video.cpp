// mix_video.cpp -- VideoMixer and the fused-frame planner (see mix_video.h for the reference map).
#include "mix_video.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "cu_driver.h"
#include "host_prof.h"
#include "ring_layout.h"

namespace svb {

static const CuDriver& drv() {
    const CuDriver& d = cu();
    if (!d.ok) throw ComputeError(ErrorCode::deviceNotAvailable, std::string("CUDA driver unavailable: ") + d.why);
    return d;
}

// ---- per-context state: descriptor ring + tensor-map cache ----------------------------------------------

namespace {

constexpr int kSegments = 8;     // batches in flight before the host has to wait for the oldest
constexpr int kSegFrames = 64;   // frame descriptors per batch
#ifndef SVB_DEFAULT_GATHER
#define SVB_DEFAULT_GATHER 0     // which fused compositor runs unless SVB_COMPOSITOR says otherwise: 0 = svb_mix_tiled (TMA), 1 = svb_mix_gather
#endif

struct MixerShared {
    uint8_t* host = nullptr;  // pinned staging, kSegments * kSegFrames descriptors
    CUdeviceptr dev = 0;
    CUevent ev[kSegments] = {};
    bool used[kSegments] = {};
    // The descriptor copy and the pre-pass of a batch run on their own stream: nothing in them depends on the compositor launch
    // queued before them: the copy runs ahead, the pre-pass's blocks start as that launch's CTAs retire (nothing of this stream starts
    // beside its resident CTAs, profiles/r2_history.md section 9): the step is the compositor + 4 - 8 us.  Each segment
    // owns its table / plan buffer (free again once ev[seg] has fired, which the host waits for before reusing the segment).
    CUstream prep = nullptr;
    CUevent evPrep[kSegments] = {};
    CUdeviceptr tabBuf[kSegments] = {};
    size_t tabBytes[kSegments] = {};
    // What a segment's buffer holds: everything svb_strip_tables' output depends on for the ring batch that last filled it (frame
    // sizes, per layer the uniforms, source size, format, flags, staged boxes and the tables' place in the buffer).  A batch with the
    // same signature -- a mixer whose layout did not change since this segment's last turn, the steady state of a live mix -- finds
    // its tables in place and goes without the pre-pass; anything else that writes the buffer clears the signature.
    std::vector<uint8_t> tabSig[kSegments];
    std::vector<uint8_t> sigScratch;
    bool tableCache = true;  // (setTableCache: off = every ring batch runs its pre-pass, as before round 2's second session)
    int next = 0;
    // Tensor maps: a table in device memory, grown by chunks.  A slot is written once (a synchronous 128-byte copy when a new
    // plane geometry first appears) and never again, so kernels need no tensormap-proxy fence; only when the table would exceed
    // kTmapMaxChunks does it wrap around, and from then on every launch carries SVB_FRAME_TMAP_FENCE.
    static constexpr int kTmapChunk = 4096, kTmapMaxChunks = 64;
    std::map<std::array<uint64_t, 4>, CUdeviceptr> tmaps;
    std::vector<CUdeviceptr> tmapChunks;
    int tmapChunkAt = -1, tmapUsed = 0;
    bool tmapFence = false;
    CUfunction fTiled = nullptr, fGather = nullptr, fGeneric = nullptr, fTables = nullptr, fStripTables = nullptr, fRing = nullptr;
    int gatherCtasPerSm = 0;
    int texAlign = 512, texPitchAlign = 32;
    std::map<std::array<uint64_t, 3>, CUtexObject> texs;  // texture objects over source planes by (pointer, size, pitch | channels)
    std::map<size_t, int> ringCtasPerSm;  // the same for svb_mix_ring
    std::map<size_t, int> tiledCtasPerSm;  // resident CTAs of svb_mix_tiled per SM by dynamic shared memory size: the persistent grid is smCount times this
    std::mutex mu;
    std::mutex launchMu;  // one launch at a time per context: the descriptor ring, the table buffers and the timing lists are per context
    // optional per-launch device timing of the fused kernels (bench.py's roofline leg)
    bool timing = false;
    std::vector<std::pair<CUevent, CUevent>> timed, spare;
    double timedMs = 0.0;
    unsigned long long timedLaunches = 0;
    // host time spent inside mixMany / composeRaw while timing is on (plan + driver calls; what the caller's thread pays per tick)
    double hostMs = 0.0;      // host time inside composeFused while timing is on, WITHOUT ...
    double hostWaitMs = 0.0;  // ... the time spent blocked on a descriptor segment the GPU has not released yet (back-pressure)
    unsigned long long hostCalls = 0;
};

void harvest(MixerShared& s) {  // caller holds a CtxGuard
    for (auto& pr : s.timed) {
        float ms = 0.f;
        if (cu().cuEventSynchronize(pr.second) == CUDA_SUCCESS && cu().cuEventElapsedTime(&ms, pr.first, pr.second) == CUDA_SUCCESS) {
            s.timedMs += ms;
            ++s.timedLaunches;
        }
        s.spare.push_back(pr);
    }
    s.timed.clear();
}

void freeShared(InternalContext* ic) {
    auto* s = (MixerShared*)ic->mixerShared;
    if (!s) return;
    if (s->host) cu().cuMemFreeHost(s->host);
    if (s->dev) cu().cuMemFree(s->dev);
    for (CUevent e : s->ev)
        if (e) cu().cuEventDestroy(e);
    for (CUevent e : s->evPrep)
        if (e) cu().cuEventDestroy(e);
    for (CUdeviceptr b : s->tabBuf)
        if (b) cu().cuMemFree(b);
    for (auto& kv : s->texs) cu().cuTexObjectDestroy(kv.second);
    for (CUdeviceptr c : s->tmapChunks) cu().cuMemFree(c);
    if (s->prep) cu().cuStreamDestroy(s->prep);
    for (auto* v : {&s->timed, &s->spare})
        for (auto& pr : *v) {
            cu().cuEventDestroy(pr.first);
            cu().cuEventDestroy(pr.second);
        }
    delete s;
    ic->mixerShared = nullptr;
}

MixerShared& shared(const std::shared_ptr<InternalContext>& ic) {  // caller holds a CtxGuard
    std::lock_guard<std::mutex> g(ic->mu);
    if (!ic->mixerShared) {
        auto* s = new MixerShared();
        ic->mixerShared = s;
        ic->mixerSharedFree = freeShared;
        const size_t bytes = (size_t)kSegments * kSegFrames * sizeof(SvbFrameDesc);
        check(drv().cuMemHostAlloc((void**)&s->host, bytes, 0), "cuMemHostAlloc");
        check(drv().cuMemAlloc(&s->dev, bytes), "cuMemAlloc");
        for (CUevent& e : s->ev) check(drv().cuEventCreate(&e, CU_EVENT_DISABLE_TIMING), "cuEventCreate");
        for (CUevent& e : s->evPrep) check(drv().cuEventCreate(&e, CU_EVENT_DISABLE_TIMING), "cuEventCreate");
        check(drv().cuStreamCreate(&s->prep, CU_STREAM_NON_BLOCKING), "cuStreamCreate");
        s->fTiled = ic->builtin("svb_mix_tiled");
        check(drv().cuFuncSetAttribute(s->fTiled, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, SVB_TILED_SMEM_MAX),
              "cuFuncSetAttribute(max dynamic shared memory)");
        s->fTables = ic->builtin("svb_mix_tables");
        s->fGather = ic->builtin("svb_mix_gather");
        {
            int n = 0;
            check(drv().cuOccupancyMaxActiveBlocksPerMultiprocessor(&n, s->fGather, SVB_TILED_THREADS, 0), "cuOccupancyMaxActiveBlocksPerMultiprocessor");
            s->gatherCtasPerSm = std::max(1, n);
            drv().cuDeviceGetAttribute(&s->texAlign, CU_DEVICE_ATTRIBUTE_TEXTURE_ALIGNMENT, ic->device);
            drv().cuDeviceGetAttribute(&s->texPitchAlign, CU_DEVICE_ATTRIBUTE_TEXTURE_PITCH_ALIGNMENT, ic->device);
            if (s->texAlign <= 0) s->texAlign = 512;
            if (s->texPitchAlign <= 0) s->texPitchAlign = 32;
        }
        s->fGeneric = ic->builtin("svb_mix_generic");
        s->fStripTables = ic->builtin("svb_strip_tables");
        s->fRing = ic->builtin("svb_mix_ring");
        check(drv().cuFuncSetAttribute(s->fRing, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, 200 * 1024), "cuFuncSetAttribute(max dynamic shared memory)");
    }
    return *(MixerShared*)ic->mixerShared;
}

int svbFormat(PixelFormat f) {
    switch (f) {
    case PixelFormat::nv12: return SVB_NV12;
    case PixelFormat::y420p: return SVB_Y420P;
    case PixelFormat::BGRA: return SVB_BGRA;
    case PixelFormat::RGBA: return SVB_RGBA;
    case PixelFormat::nv21: return SVB_NV21;
    case PixelFormat::y422p: return SVB_Y422P;
    case PixelFormat::y444p: return SVB_Y444P;
    default: return -1;
    }
}

bool finite16(const float* m) {
    for (int i = 0; i < 16; ++i)
        if (!std::isfinite(m[i])) return false;
    return true;
}

int roundUp(int v, int m) { return (v + m - 1) / m * m; }

// 2-D tensor map over one plane, in the context's device table; cached by plane and box (the encode costs ~1 us, the first use of a
// geometry one small synchronous copy; the device pool recycles blocks, so the same few planes come round).  out = the slot's address.
bool tensorMap(MixerShared& sh, unsigned long long* out, CUdeviceptr ptr, int elemBytes, int w, int h, int strideBytes, int boxW, int boxH) {
    if ((ptr & 15) || (strideBytes & 15) || w <= 0 || h <= 0 || boxW > 256 || boxH > 256 || (boxW * elemBytes) % 16) return false;
    std::lock_guard<std::mutex> lock(sh.mu);  // mixers of one context may plan from different threads
    const std::array<uint64_t, 4> key = {(uint64_t)ptr, ((uint64_t)(uint32_t)w << 32) | (uint32_t)h,
                                         ((uint64_t)(uint32_t)strideBytes << 32) | (uint32_t)elemBytes,
                                         ((uint64_t)(uint32_t)boxW << 32) | (uint32_t)boxH};
    auto it = sh.tmaps.find(key);
    if (it == sh.tmaps.end()) {
        alignas(64) CUtensorMap tm;
        static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
        const cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)h};
        const cuuint64_t gstride[1] = {(cuuint64_t)strideBytes};
        const cuuint32_t box[2] = {(cuuint32_t)boxW, (cuuint32_t)boxH};
        const cuuint32_t estride[2] = {1, 1};
        CUresult r = drv().cuTensorMapEncodeTiled(&tm, elemBytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                                                  (void*)ptr, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
        if (sh.tmapChunkAt < 0 || sh.tmapUsed == MixerShared::kTmapChunk) {
            if ((int)sh.tmapChunks.size() < MixerShared::kTmapMaxChunks) {
                CUdeviceptr c = 0;
                if (drv().cuMemAlloc(&c, (size_t)MixerShared::kTmapChunk * 128) != CUDA_SUCCESS) return false;
                sh.tmapChunks.push_back(c);
                sh.tmapChunkAt = (int)sh.tmapChunks.size() - 1;
            } else {  // wrap: slots are rewritten from here on -- launches in flight finish first, later ones fence
                if (drv().cuCtxSynchronize() != CUDA_SUCCESS) return false;
                sh.tmapChunkAt = (sh.tmapChunkAt + 1) % MixerShared::kTmapMaxChunks;
                sh.tmaps.clear();
                sh.tmapFence = true;
            }
            sh.tmapUsed = 0;
        }
        const CUdeviceptr slot = sh.tmapChunks[sh.tmapChunkAt] + (size_t)sh.tmapUsed * 128;
        if (drv().cuMemcpyHtoD(slot, &tm, 128) != CUDA_SUCCESS) return false;
        ++sh.tmapUsed;
        it = sh.tmaps.emplace(key, slot).first;
    }
    *out = (unsigned long long)it->second;
    return true;
}

// A texture object over one source plane for svb_mix_gather: 8-bit channels read as UNORM floats, unnormalised coordinates,
// clamp addressing (the OpenCL sampler's CLAMP_TO_EDGE, kernels.cl.swift:61).  0 when the plane cannot be bound (base or pitch
// alignment).  Cached by pointer and shape: the device pool recycles blocks, so the same few planes come round.
CUtexObject textureObject(MixerShared& sh, CUdeviceptr ptr, int channels, int w, int h, int pitchBytes) {
    if (!ptr || w <= 0 || h <= 0 || (ptr % (CUdeviceptr)sh.texAlign) || (pitchBytes % sh.texPitchAlign) || pitchBytes < w * channels) return 0;
    const std::array<uint64_t, 3> key = {(uint64_t)ptr, ((uint64_t)(uint32_t)w << 32) | (uint32_t)h, ((uint64_t)(uint32_t)pitchBytes << 8) | (uint32_t)channels};
    std::lock_guard<std::mutex> lock(sh.mu);  // mixers of one context may plan from different threads
    auto it = sh.texs.find(key);
    if (it != sh.texs.end()) return it->second;
    CUDA_RESOURCE_DESC rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.resType = CU_RESOURCE_TYPE_PITCH2D;
    rd.res.pitch2D.devPtr = ptr;
    rd.res.pitch2D.format = CU_AD_FORMAT_UNSIGNED_INT8;
    rd.res.pitch2D.numChannels = (unsigned)channels;
    rd.res.pitch2D.width = (size_t)w;
    rd.res.pitch2D.height = (size_t)h;
    rd.res.pitch2D.pitchInBytes = (size_t)pitchBytes;
    CUDA_TEXTURE_DESC td;
    std::memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = CU_TR_ADDRESS_MODE_CLAMP;
    td.filterMode = CU_TR_FILTER_MODE_POINT;  // tex2Dgather fetches the footprint; the filter runs in fp32 in the kernel
    td.flags = 0;                             // UNORM8 -> float in the unit, unnormalised coordinates
    // bounded: the device pool recycles blocks, so a long run meets the same few hundred planes again; past the bound a plane simply
    // goes without (its batch then takes the TMA compositor).  Nothing is destroyed before the context goes: a descriptor already
    // handed to a launch may hold the handle.
    if (sh.texs.size() >= 4096) return 0;
    CUtexObject obj = 0;
    if (drv().cuTexObjectCreate(&obj, &rd, &td, nullptr) != CUDA_SUCCESS) return 0;
    sh.texs.emplace(key, obj);
    return obj;
}

// Canvas rectangle outside which `border` cannot land in [0,1]^2: the unit square pushed through the forward
// border map, grown by 2 px (fp32 rounding of the kernel's chain moves an edge by far less than a pixel).
void layerRect(const SvbUniforms& u, int W, int H, int32_t rect[4]) {
    rect[0] = 0, rect[1] = 0, rect[2] = W, rect[3] = H;
    const double a = u.borderMatrix[0], b = u.borderMatrix[1], c = u.borderMatrix[4], d = u.borderMatrix[5];
    const double e = u.borderMatrix[3], f = u.borderMatrix[7];
    const double det = a * d - b * c;
    if (!finite16(u.borderMatrix) || !(std::fabs(det) > 1e-30)) return;
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int k = 0; k < 4; ++k) {
        const double bx = (k & 1) - e, by = (k >> 1) - f;  // border = (k&1, k>>1)
        const double nx = (d * bx - b * by) / det, ny = (-c * bx + a * by) / det;
        const double px = (nx + 1.0) * 0.5 * W, py = (ny + 1.0) * 0.5 * H;
        xmin = std::min(xmin, px), xmax = std::max(xmax, px), ymin = std::min(ymin, py), ymax = std::max(ymax, py);
    }
    if (!(std::isfinite(xmin) && std::isfinite(xmax) && std::isfinite(ymin) && std::isfinite(ymax))) return;
    auto clampi = [](double v, int lo, int hi) { return (int)std::min<double>(std::max<double>(v, lo), hi); };
    rect[0] = clampi(std::floor(xmin) - 2, 0, W);
    rect[1] = clampi(std::floor(ymin) - 2, 0, H);
    rect[2] = clampi(std::ceil(xmax) + 3, 0, W);
    rect[3] = clampi(std::ceil(ymax) + 3, 0, H);
}

}  // namespace

// ---- planner --------------------------------------------------------------------------------------------

FramePlan planFrame(const ComputeContext& ctx, const PictureSample& target, const std::vector<const PictureSample*>& layers,
                    const ImageUniforms* uniforms, const SvbLayerDesc* const* cached) {
    FramePlan plan;
    const int tf = svbFormat(target.pixelFormat());
    if (tf != SVB_NV12 && tf != SVB_Y420P) throw ComputeError(ErrorCode::badTarget, "fused compose needs an nv12 or y420p target");
    const int W = (int)target.size().x, H = (int)target.size().y;
    if (W <= 0 || H <= 0 || (W & 1) || (H & 1)) throw ComputeError(ErrorCode::badTarget, "target size must be even");
    const auto& tex = target.imgBuffer.computeTextures;
    const auto& planes = target.imgBuffer.planes;
    const size_t np = tf == SVB_NV12 ? 2 : 3;
    if (tex.size() < np || planes.size() < np) throw ComputeError(ErrorCode::badTarget, "badTarget");

    SvbFrameDesc base;
    std::memset(&base, 0, sizeof(base));
    for (size_t i = 0; i < np; ++i) {
        base.out_plane[i] = tex[i]->mem;
        base.out_stride[i] = planes[i].stride;
    }
    base.width = W, base.height = H, base.format = tf;
    static const bool scalarFp = std::getenv("SVB_FP32X2") && std::getenv("SVB_FP32X2")[0] == '0';  // A/B timing aid
    if (scalarFp) base.flags |= SVB_FRAME_SCALAR_FP;
    base.tiles_x = (W + SVB_TILE_W - 1) / SVB_TILE_W;
    base.tiles_y = (H + SVB_TILE_H - 1) / SVB_TILE_H;
    // the generic kernel stores luma (and NV12 chroma) two bytes at a time
    if ((planes[0].stride & 1) || (tf == SVB_NV12 && (planes[1].stride & 1))) throw ComputeError(ErrorCode::badTarget, "target strides must be even");
    // the tiled kernel stores luma and NV12 chroma two bytes at a time (Y420P chroma byte by byte)
    plan.tiledOk = (W % 2 == 0) && (H % 2 == 0) && (base.out_plane[0] % 2 == 0) && (tf != SVB_NV12 || base.out_plane[1] % 2 == 0);

    MixerShared& sh = shared(ctx.ctx);
    const size_t n = layers.size();
    const size_t npass = std::max<size_t>(1, (n + SVB_MAX_LAYERS - 1) / SVB_MAX_LAYERS);
    plan.passes.assign(npass, base);
    for (size_t p = 0; p < npass; ++p)
        if (p) plan.passes[p].flags |= SVB_FRAME_LOAD_CUR;
    for (size_t k = 0; k < n; ++k) {
        SvbFrameDesc& F = plan.passes[k / SVB_MAX_LAYERS];
        SvbLayerDesc& L = F.layers[F.nlayers++];
        if (cached && cached[k]) {  // planned before for this very sample and a target of this size
            L = *cached[k];
            if (L.format >= SVB_NV21) plan.tiledOk = false;
            continue;
        }
        const PictureSample& img = *layers[k];
        const int sf = svbFormat(img.pixelFormat());
        if (sf < 0) throw ComputeError(ErrorCode::computeKernelNotFound, std::string("computeKernelNotFound(img_") + pixelFormatName(img.pixelFormat()) + "_" + pixelFormatName(target.pixelFormat()) + ")");
        if (img.bufferType() != BufferType::gpu) throw ComputeError(ErrorCode::badInputData, "Input images must be uploaded to GPU");
        const size_t snp = SVB_FORMAT_IS_SEMIPLANAR(sf) ? 2 : (SVB_FORMAT_IS_YUV(sf) ? 3 : 1);
        if (img.imgBuffer.computeTextures.size() < snp || img.imgBuffer.planes.size() < snp) throw ComputeError(ErrorCode::badInputData, "Bad input image");
        std::memcpy(&L.u, &uniforms[k], sizeof(ImageUniforms));
        L.u.pad_ = 0.f;
        for (size_t i = 0; i < snp; ++i) {
            L.plane[i] = img.imgBuffer.computeTextures[i]->mem;
            L.stride[i] = img.imgBuffer.planes[i].stride;
        }
        L.width = (int)img.imgBuffer.planes[0].size.x;
        L.height = (int)img.imgBuffer.planes[0].size.y;
        L.format = sf;
        if (sf >= SVB_NV21) plan.tiledOk = false;  // the tile compositors know the reference's four source formats; these go through svb_mix_generic
        const SvbUniforms& u = L.u;
        const float *T = u.transform, *X = u.textureTx, *B = u.borderMatrix;
        const bool finite = finite16(T) && finite16(X) && finite16(B) && std::isfinite(u.opacity);
        int flags = 0;
        static const bool noTables = std::getenv("SVB_NO_TABLES") != nullptr, noTma = std::getenv("SVB_NO_TMA") != nullptr;  // debugging aids
        if (!noTables && finite && T[1] == 0.f && T[4] == 0.f && T[8] == 0.f && T[9] == 0.f && T[12] == 0.f && T[13] == 0.f && B[1] == 0.f &&
            B[4] == 0.f && X[1] == 0.f && X[4] == 0.f && (sf != SVB_Y420P || L.stride[1] == L.stride[2]) && L.width <= 65535 && L.height <= 65535)
            flags |= SVB_LAYER_SEPARABLE;
        if (u.opacity == 1.0f) flags |= SVB_LAYER_UNIT_OPACITY;
        if (u.opacity >= 0.f && u.opacity <= 1.f) flags |= SVB_LAYER_OPACITY_01;
        layerRect(u, W, H, L.rect);
        if ((flags & SVB_LAYER_SEPARABLE) && (sf == SVB_NV12 || sf == SVB_Y420P) && L.width >= 2 && L.height >= 2) {
            // source texels per output pixel along each axis (double is plenty: the device re-checks the fit per tile)
            const double sx = std::fabs((double)L.width * X[0] * T[0] * 2.0 / W), sy = std::fabs((double)L.height * X[5] * T[5] * 2.0 / H);
            const int cw = L.width / 2, ch = L.height / 2;
            // +15 bytes of slack: the kernel rounds the box's x origin down to a 16-byte boundary (a TMA requirement)
            int bw = roundUp((int)std::ceil(sx * (SVB_TILE_W - 1)) + 3 + 15, 16), bh = (int)std::ceil(sy * (SVB_TILE_H - 1)) + 3;
            int bcw = sf == SVB_NV12 ? roundUp((int)std::ceil(sx * 0.5 * (SVB_TILE_W - 2)) + 3 + 7, 8)
                                     : roundUp((int)std::ceil(sx * 0.5 * (SVB_TILE_W - 2)) + 3 + 15, 16);
            int bch = (int)std::ceil(sy * 0.5 * (SVB_TILE_H - 2)) + 3;
            const int cbytes = bcw * bch * (sf == SVB_NV12 ? 2 : 1);
            const bool fits = std::isfinite(sx) && std::isfinite(sy) && bw <= 256 && bh <= 256 && bw * bh <= SVB_BOX_Y_BYTES &&
                              bcw <= 256 && bch <= 256 && cbytes <= (sf == SVB_NV12 ? SVB_BOX_C_BYTES : SVB_BOX_C_BYTES / 2);
            if (!noTma && fits && tensorMap(sh, &L.tmap[0], L.plane[0], 1, L.width, L.height, L.stride[0], bw, bh) &&
                (sf == SVB_NV12 ? tensorMap(sh, &L.tmap[1], L.plane[1], 2, cw, ch, L.stride[1], bcw, bch)
                                : (tensorMap(sh, &L.tmap[1], L.plane[1], 1, cw, ch, L.stride[1], bcw, bch) &&
                                   tensorMap(sh, &L.tmap[2], L.plane[2], 1, cw, ch, L.stride[2], bcw, bch)))) {
                flags |= SVB_LAYER_STAGED;
                L.box_w = bw, L.box_h = bh, L.box_cw = bcw, L.box_ch = bch;
            }
        }
        if ((flags & SVB_LAYER_SEPARABLE) && (sf == SVB_NV12 || sf == SVB_Y420P) && L.width >= 2 && L.height >= 2) {
            const int cw = L.width / 2, ch = L.height / 2;
            const CUtexObject t0 = textureObject(sh, L.plane[0], 1, L.width, L.height, L.stride[0]);
            const CUtexObject t1 = sf == SVB_NV12 ? textureObject(sh, L.plane[1], 2, cw, ch, L.stride[1]) : textureObject(sh, L.plane[1], 1, cw, ch, L.stride[1]);
            const CUtexObject t2 = sf == SVB_NV12 ? t1 : textureObject(sh, L.plane[2], 1, cw, ch, L.stride[2]);
            if (t0 && t1 && t2) {
                L.tex[0] = t0, L.tex[1] = t1, L.tex[2] = t2;
                flags |= SVB_LAYER_TEX;
            }
        }
        L.flags = flags;
    }
    return plan;
}

// ---- launch ---------------------------------------------------------------------------------------------

namespace {

// SVB_COMPOSITOR=ring|tma|gather picks the fused compositor (default: ring = svb_mix_ring; tma = svb_mix_tiled, the round-1 kernel;
// gather = svb_mix_gather, taps through the texture unit); read once.
int compositorChoice() {  // 1 tiled, 2 gather, 3 ring
    static const int v = [] {
        const char* e = std::getenv("SVB_COMPOSITOR");
        if (e && std::strcmp(e, "ring") == 0) return 3;
        if (e && std::strcmp(e, "gather") == 0) return 2;
        if (e && (std::strcmp(e, "tma") == 0 || std::strcmp(e, "tiled") == 0)) return 1;
        return SVB_DEFAULT_GATHER ? 2 : 3;
    }();
    return v;
}
bool gatherByDefault() { return compositorChoice() == 2; }

// One launch over `frames` (all tiled-capable, or all generic).  Caller holds a CtxGuard.
void launchFrames(const ComputeContext& ctx, std::vector<SvbFrameDesc>& frames, bool tiled, bool wantGather, int want) {  // want: 0 default, 1 svb_mix_tiled, 3 svb_mix_ring
    const CuDriver& d = drv();
    MixerShared& sh = shared(ctx.ctx);
    InternalContext& ic = *ctx.ctx;
    std::lock_guard<std::mutex> launchLock(sh.launchMu);  // mixers of one context may be driven from different threads
    for (size_t start = 0; start < frames.size(); start += kSegFrames) {
        const int n = (int)std::min<size_t>(kSegFrames, frames.size() - start);
        int seg;
        {
            std::lock_guard<std::mutex> g(sh.mu);
            seg = sh.next;
            sh.next = (sh.next + 1) % kSegments;
        }
        if (sh.used[seg]) {  // eight segments: the host may run eight launches ahead of the GPU, then it waits here
            SVB_PROF(8, "launch: wait for segment (GPU back-pressure)");
            const auto w0 = std::chrono::steady_clock::now();
            check(d.cuEventSynchronize(sh.ev[seg]), "cuEventSynchronize");
            sh.hostWaitMs += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
        }
        SVB_PROF(9, "launch: fill + copy + 2 launches");
        SvbFrameDesc* host = (SvbFrameDesc*)(sh.host + (size_t)seg * kSegFrames * sizeof(SvbFrameDesc));
        int total = 0, maxW = 0, maxH = 0, maxLayers = 1, maxEnts = 1, maxTiles = 1;
        size_t tableEnts = 0;
        int boxY = 0, boxC = 0;   // largest staged footprint of the batch (bytes per box; chroma = both planes of a Y420P source)
        // Which compositor: svb_mix_gather (taps through the texture unit) when every layer the TMA kernel would stage can be
        // bound as a texture too (base / pitch alignment); SVB_COMPOSITOR=tma|gather overrides the default.
        bool gather = tiled && (wantGather || gatherByDefault());
        for (int i = 0; gather && i < n; ++i) {
            const SvbFrameDesc& fr = frames[start + i];
            for (int l = 0; l < fr.nlayers; ++l)
                if ((fr.layers[l].flags & SVB_LAYER_STAGED) && !(fr.layers[l].flags & SVB_LAYER_TEX)) gather = false;
        }
        // which TMA compositor: svb_mix_ring by default (SVB_COMPOSITOR / the mixer's mode can ask for svb_mix_tiled)
        const int pick = want ? want : (compositorChoice() == 3 ? 3 : 1);
        const bool ring = tiled && !gather && pick == 3;  // same tile boxes as svb_mix_tiled: whatever that kernel stages, this one does
        for (int i = 0; i < n; ++i) {
            SvbFrameDesc& fr = frames[start + i];
            if (gather) fr.flags |= SVB_FRAME_GATHER;
            else fr.flags &= ~SVB_FRAME_GATHER;
            if (sh.tmapFence) fr.flags |= SVB_FRAME_TMAP_FENCE;
            if (ring) fr.flags |= SVB_FRAME_RING;
            else fr.flags &= ~SVB_FRAME_RING;
            for (int l = 0; l < fr.nlayers; ++l) {
                const SvbLayerDesc& L = fr.layers[l];
                if (!(L.flags & SVB_LAYER_STAGED)) continue;
                boxY = std::max(boxY, roundUp(L.box_w * L.box_h, 256));
                boxC = std::max(boxC, L.format == SVB_NV12 ? roundUp(L.box_cw * L.box_ch * 2, 256) : 2 * roundUp(L.box_cw * L.box_ch, 128));
            }
            // (a ring batch counts its 64x8 units here -- its tables are unit-blocked -- where a tiled batch counts 128x32 tiles: same fields)
            fr.tiles_x = ring ? SVB_UNITS_X(fr.width) : SVB_TILES_X(fr.width);
            fr.tiles_y = ring ? SVB_UNITS_Y(fr.height) : SVB_TILES_Y(fr.height);
            fr.first_tile = total;
            total += ring ? SVB_TILES_X(fr.width) * SVB_TILES_Y(fr.height) : fr.tiles_x * fr.tiles_y;  // (svb_mix_ring claims 128x32 tiles; its tables are unit-blocked all the same)
            maxTiles = std::max(maxTiles, fr.tiles_x * fr.tiles_y);
            maxW = std::max(maxW, fr.width), maxH = std::max(maxH, fr.height);
            const int ents = ring ? SVB_UTABLE_WORDS(fr.width, fr.height) : SVB_TABLE_WORDS(fr.width, fr.height);  // 4-byte words per layer
            fr.table_base = (int32_t)tableEnts;
            if (ring)
                for (int l = 0; l < fr.nlayers; ++l) {  // what a planning lane of svb_mix_ring reads of its layer
                    SvbLayerDesc& L = fr.layers[l];
                    SvbStripConsts& c = L.pc;
                    const unsigned long long* tm = L.tmap;
                    c.stmapY[0] = (uint32_t)tm[0], c.stmapY[1] = (uint32_t)(tm[0] >> 32);
                    c.stmapC[0] = (uint32_t)tm[1], c.stmapC[1] = (uint32_t)(tm[1] >> 32);
                    c.stmapV[0] = (uint32_t)tm[2], c.stmapV[1] = (uint32_t)(tm[2] >> 32);
                    std::memcpy(&c.opacity_bits, &L.u.opacity, 4);
                    const int cb = L.box_cw * L.box_ch * (L.format == SVB_NV12 ? 2 : 1);
                    c.stx_bytes = (L.flags & SVB_LAYER_STAGED) ? (uint32_t)(L.box_w * L.box_h + (L.format == SVB_NV12 ? cb : 2 * cb)) : 0u;  // the table blocks are added per tile
                    c.pitches = (uint32_t)L.box_w | ((uint32_t)(L.format == SVB_NV12 ? 2 * L.box_cw : L.box_cw) << 16);
                    c.fmtflags = (uint32_t)L.format | ((uint32_t)(L.flags & 0xff) << 4) | ((uint32_t)L.box_h << 12) | ((uint32_t)L.box_ch << 22);
                    c.tab = (uint32_t)(tableEnts + (size_t)l * (size_t)ents);
                    c.rec = c.tab + (uint32_t)(fr.tiles_x * SVB_UCOL_WORDS + fr.tiles_y * SVB_UROW_WORDS);
                    std::memcpy(c.rect, L.rect, sizeof(c.rect));
                }
            tableEnts += (size_t)ents * (size_t)fr.nlayers;
            maxLayers = std::max(maxLayers, fr.nlayers), maxEnts = std::max(maxEnts, ents);
            // copy only the header and the layers in use
            const size_t used = offsetof(SvbFrameDesc, layers) + sizeof(SvbLayerDesc) * (size_t)fr.nlayers;
            std::memcpy(&host[i], &fr, used);
        }
        CUdeviceptr dev = sh.dev + (size_t)seg * kSegFrames * sizeof(SvbFrameDesc);
        check(d.cuMemcpyHtoDAsync(dev, host, (size_t)n * sizeof(SvbFrameDesc), sh.prep), "cuMemcpyHtoDAsync");
        std::pair<CUevent, CUevent> tev{nullptr, nullptr};
        if (sh.timing) {
            if (sh.timed.size() >= 256) harvest(sh);
            if (!sh.spare.empty()) {
                tev = sh.spare.back();
                sh.spare.pop_back();
            } else {
                check(d.cuEventCreate(&tev.first, CU_EVENT_DEFAULT), "cuEventCreate");
                check(d.cuEventCreate(&tev.second, CU_EVENT_DEFAULT), "cuEventCreate");
            }
        }
        if (tiled) {
            // pre-pass (one launch): per-column / per-row coordinate tables of the batch (two words per entry) and the plan of every
            // tile; then the compositor (+ the tile counter it claims its work from, zeroed by the pre-pass)
            const size_t counterOff = (std::max<size_t>(tableEnts, 4) * 4 + 15) & ~(size_t)15;
            const size_t plansOff = counterOff + 16;
            const size_t tableBytes = plansOff + (ring ? 0 : (size_t)total * sizeof(SvbTilePlan));  // (a ring batch has no plans in global memory: its producer warps plan their own tiles)
            if (sh.tabBytes[seg] < tableBytes) {
                // cuMemAlloc / cuMemFree synchronise the device: when one segment's buffer must grow, grow them all, once, instead of
                // stalling the next kSegments - 1 launches as well (a mixer's first ticks are often lighter than its steady state)
                check(d.cuCtxSynchronize(), "cuCtxSynchronize");
                const size_t want = tableBytes + tableBytes / 4;
                for (int k = 0; k < kSegments; ++k) {
                    if (sh.tabBytes[k] >= want) continue;
                    if (sh.tabBuf[k]) check(d.cuMemFree(sh.tabBuf[k]), "cuMemFree");
                    sh.tabBuf[k] = 0, sh.tabBytes[k] = 0;
                    check(d.cuMemAlloc(&sh.tabBuf[k], want), "cuMemAlloc");
                    sh.tabBytes[k] = want;
                    sh.tabSig[k].clear();
                }
            }
            CUdeviceptr tables = sh.tabBuf[seg];
            CUdeviceptr counter = tables + counterOff, plans = tables + plansOff;
            if (ring) {
                // the batch's signature (see tabSig): when the segment's buffer already holds these tables the pre-pass is not launched --
                // svb_mix_ring leaves the tile counter at zero behind itself
                std::vector<uint8_t>& sig = sh.sigScratch;
                sig.clear();
                auto put = [&sig](const void* p, size_t bytes) { sig.insert(sig.end(), (const uint8_t*)p, (const uint8_t*)p + bytes); };
                for (int i = 0; i < n; ++i) {
                    const SvbFrameDesc& fr = frames[start + i];
                    const int32_t head[8] = {fr.width, fr.height, fr.nlayers, fr.tiles_x, fr.tiles_y, fr.table_base, fr.first_tile, fr.format};
                    put(head, sizeof(head));
                    for (int l = 0; l < fr.nlayers; ++l) {
                        const SvbLayerDesc& L = fr.layers[l];
                        put(&L.u, sizeof(L.u));
                        const int32_t geo[10] = {L.width, L.height, L.format, L.flags, L.box_w, L.box_h, L.box_cw, L.box_ch, (int32_t)L.pc.tab, (int32_t)L.pc.rec};
                        put(geo, sizeof(geo));
                    }
                }
                if (!sh.tableCache || sig != sh.tabSig[seg]) {
                    // one pre-pass launch: a block per unit column and per unit row of every layer fills its table block and its plan record
                    int blocks = 1;
                    for (int i = 0; i < n; ++i) blocks = std::max(blocks, frames[start + i].tiles_x + frames[start + i].tiles_y);
                    void* targs[] = {&dev, &tables, &counter};
                    check(d.cuLaunchKernel(sh.fStripTables, (unsigned)blocks, (unsigned)maxLayers, (unsigned)n, 96, 1, 1, 0, sh.prep, targs, nullptr), "cuLaunchKernel(svb_strip_tables)");
                    noteKernelLaunch();
                    sh.tabSig[seg] = sig;
                }
            } else {
                sh.tabSig[seg].clear();  // (the other compositors lay the buffer out differently)
                int tableBlocks = (maxEnts / 2 + 255) / 256;
                void* targs[] = {&dev, &tables, &counter, &plans, &tableBlocks, &maxLayers};
                check(d.cuLaunchKernel(sh.fTables, (unsigned)(tableBlocks * maxLayers + (maxTiles + 7) / 8), 1, (unsigned)n, 256, 1, 1, 0, sh.prep, targs, nullptr),
                      "cuLaunchKernel(svb_mix_tables)");
                noteKernelLaunch();
            }
            check(d.cuEventRecord(sh.evPrep[seg], sh.prep), "cuEventRecord");
            check(d.cuStreamWaitEvent(ic.compute, sh.evPrep[seg], 0), "cuStreamWaitEvent");
            if (tev.first) check(d.cuEventRecord(tev.first, ic.compute), "cuEventRecord");  // time svb_mix_tiled alone
            float one = 1.0f;  // see add2() in kernels_tiled.cuh
            if (ring) {
                int nf = n, slotBytes = SVB_RPLAN_SLOT_BYTES(maxLayers);
                void* args[] = {&dev, &tables, &nf, &total, &one, &counter, &boxY, &boxC, &slotBytes};
                size_t smem = SVB_RING_SMEM_BYTES((size_t)boxY, (size_t)boxC, maxLayers);
                int perSm;
                {
                    std::lock_guard<std::mutex> g(sh.mu);
                    auto it = sh.ringCtasPerSm.find(smem);
                    if (it == sh.ringCtasPerSm.end()) {
                        int nb = 0;
                        check(d.cuOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sh.fRing, SVB_RING_THREADS, smem), "cuOccupancyMaxActiveBlocksPerMultiprocessor");
                        it = sh.ringCtasPerSm.emplace(smem, std::max(1, nb)).first;
                    }
                    perSm = it->second;
                }
                const unsigned grid = (unsigned)std::min(total, ic.smCount * perSm);
                check(d.cuLaunchKernel(sh.fRing, grid, 1, 1, SVB_RING_THREADS, 1, 1, (unsigned)smem, ic.compute, args, nullptr), "cuLaunchKernel(svb_mix_ring)");
                noteKernelLaunch();
            } else if (gather) {
                int units = total * SVB_GATHER_STRIPS;  // (tile, strip) pairs, claimed by warps
                void* args[] = {&dev, &plans, &units, &one, &counter};
                const unsigned grid = (unsigned)std::min((units + SVB_TILED_COMPUTE_WARPS - 1) / SVB_TILED_COMPUTE_WARPS, ic.smCount * sh.gatherCtasPerSm);
                check(d.cuLaunchKernel(sh.fGather, grid, 1, 1, SVB_TILED_THREADS, 1, 1, 0, ic.compute, args, nullptr), "cuLaunchKernel(svb_mix_gather)");
                noteKernelLaunch();
            } else {
                void* args[] = {&dev, &plans, &total, &one, &counter, &boxY, &boxC};
                // shared memory: the fixed part plus two box pairs sized for the largest staged footprint of this batch
                size_t smem = SVB_TILED_SMEM_BYTES((size_t)boxY, (size_t)boxC);
                int perSm;
                {
                    std::lock_guard<std::mutex> g(sh.mu);
                    auto it = sh.tiledCtasPerSm.find(smem);
                    if (it == sh.tiledCtasPerSm.end()) {
                        int n = 0;
                        check(d.cuOccupancyMaxActiveBlocksPerMultiprocessor(&n, sh.fTiled, SVB_TILED_THREADS, smem), "cuOccupancyMaxActiveBlocksPerMultiprocessor");
                        it = sh.tiledCtasPerSm.emplace(smem, std::max(1, n)).first;
                    }
                    perSm = it->second;
                }
                const unsigned grid = (unsigned)std::min(total, ic.smCount * perSm);
                check(d.cuLaunchKernel(sh.fTiled, grid, 1, 1, SVB_TILED_THREADS, 1, 1, (unsigned)smem, ic.compute, args, nullptr), "cuLaunchKernel(svb_mix_tiled)");
                noteKernelLaunch();
            }
        } else {
            check(d.cuEventRecord(sh.evPrep[seg], sh.prep), "cuEventRecord");  // the descriptor copy
            check(d.cuStreamWaitEvent(ic.compute, sh.evPrep[seg], 0), "cuStreamWaitEvent");
            if (tev.first) check(d.cuEventRecord(tev.first, ic.compute), "cuEventRecord");
            void* args[] = {&dev};
            check(d.cuLaunchKernel(sh.fGeneric, (unsigned)((maxW / 2 + 31) / 32), (unsigned)((maxH / 2 + 7) / 8), (unsigned)n, 32, 8, 1, 0,
                                   ic.compute, args, nullptr),
                  "cuLaunchKernel(svb_mix_generic)");
            noteKernelLaunch();
        }
        if (tev.first) {
            check(d.cuEventRecord(tev.second, ic.compute), "cuEventRecord");
            sh.timed.push_back(tev);
        }
        check(d.cuEventRecord(sh.ev[seg], ic.compute), "cuEventRecord");
        sh.used[seg] = true;
    }
}

void waitInputs(const ComputeContext& ctx, const PictureSample& p, bool willWrite = false) {
    for (const auto& t : p.imgBuffer.computeTextures) {
        waitReady(ctx.ctx->compute, *t);
        // a target: an asynchronous download of what the plane held before may still be reading it (the backing ring comes round
        // every ten ticks and a compose is an order of magnitude faster than the copy over PCIe)
        if (willWrite && t->lastRead) check(drv().cuStreamWaitEvent(ctx.ctx->compute, t->lastRead->e, 0), "cuStreamWaitEvent");
        if (willWrite && t->consumerRead) check(drv().cuStreamWaitEvent(ctx.ctx->compute, t->consumerRead->e, 0), "cuStreamWaitEvent");
    }
}

struct Job {
    const PictureSample* target;
    std::vector<const PictureSample*> layers;
    std::vector<ImageUniforms> uniforms;
    std::vector<const SvbLayerDesc*> cached;    // per layer: a descriptor planned earlier for the same sample, or nullptr (empty: none)
    std::vector<SvbLayerDesc>* planned = nullptr;  // out: the descriptors as planned now, in layer order
};

// Fused compose of several independent targets on one context.
void composeFused(const ComputeContext& ctx, std::vector<Job>& jobs, bool forceGeneric, bool wantGather = false, int want = 0) {
    CtxGuard g(ctx.ctx);
    struct HostClock {  // (only while launch timing is on)
        MixerShared& sh;
        double waited0;
        std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        ~HostClock() {
            if (sh.timing) sh.hostMs += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() - (sh.hostWaitMs - waited0), ++sh.hostCalls;
        }
    } hostClock{shared(ctx.ctx), shared(ctx.ctx).hostWaitMs};
    SVB_PROF(3, "composeFused");
    std::vector<FramePlan> plans;
    bool allTiled = !forceGeneric;
    size_t maxPass = 0;
    for (Job& j : jobs) {
        SVB_PROF(4, "compose: planFrame + waitInputs (per frame)");
        plans.push_back(planFrame(ctx, *j.target, j.layers, j.uniforms.data(), j.cached.empty() ? nullptr : j.cached.data()));
        if (j.planned) {
            j.planned->clear();
            for (const SvbFrameDesc& f : plans.back().passes) j.planned->insert(j.planned->end(), f.layers, f.layers + f.nlayers);
        }
        allTiled = allTiled && plans.back().tiledOk;
        maxPass = std::max(maxPass, plans.back().passes.size());
        waitInputs(ctx, *j.target, true);
        for (const PictureSample* l : j.layers) waitInputs(ctx, *l);
    }
    for (size_t p = 0; p < maxPass; ++p) {  // a later pass of a frame reads what its earlier pass wrote: separate launches
        std::vector<SvbFrameDesc> frames;
        for (FramePlan& pl : plans)
            if (p < pl.passes.size()) frames.push_back(pl.passes[p]);
        launchFrames(ctx, frames, allTiled, wantGather, want);
    }
    SVB_PROF(10, "compose: markWritten + lastUse");
    for (Job& j : jobs) markWritten(ctx, *j.target);
    // every layer plane learns the point on the compute stream after which this compose no longer reads it (ComputeBuffer::lastUse)
    auto used = std::make_shared<Event>(ctx.ctx);
    check(drv().cuEventRecord(used->e, ctx.ctx->compute), "cuEventRecord");
    for (Job& j : jobs)
        for (const PictureSample* l : j.layers)
            for (const auto& t : l->imgBuffer.computeTextures) t->lastUse = used;
}

}  // namespace

void setTableCache(const ComputeContext& ctx, bool on) {
    MixerShared& sh = shared(ctx.ctx);
    std::lock_guard<std::mutex> launchLock(sh.launchMu);
    sh.tableCache = on;
}

void setLaunchTiming(const ComputeContext& ctx, bool on) {
    CtxGuard g(ctx.ctx);
    MixerShared& sh = shared(ctx.ctx);
    harvest(sh);
    sh.timing = on;
    sh.timedMs = 0.0;
    sh.timedLaunches = 0;
    sh.hostMs = 0.0;
    sh.hostWaitMs = 0.0;
    sh.hostCalls = 0;
}
void readHostTiming(const ComputeContext& ctx, double* totalMs, unsigned long long* calls, double* waitMs) {
    CtxGuard g(ctx.ctx);
    MixerShared& sh = shared(ctx.ctx);
    *totalMs = sh.hostMs;
    *calls = sh.hostCalls;
    if (waitMs) *waitMs = sh.hostWaitMs;
}
void readLaunchTiming(const ComputeContext& ctx, double* totalMs, unsigned long long* launches) {
    CtxGuard g(ctx.ctx);
    MixerShared& sh = shared(ctx.ctx);
    harvest(sh);
    *totalMs = sh.timedMs;
    *launches = sh.timedLaunches;
}

// ---- VideoMixer -----------------------------------------------------------------------------------------

VideoMixer::VideoMixer(const ComputeContext* computeContext, Vector2 outputSize, PixelFormat outputFormat, const std::string& assetId,
                       const std::string& workspaceId, int64_t frameDuration, int64_t timescale, int64_t epoch)
    : frameDuration(frameDuration), timescale(timescale), epoch(epoch), backingFormat(outputFormat), backingSize(outputSize),
      idWorkspace(workspaceId) {
    if (computeContext) {  // mix.video.swift:32-40
        clContext = createComputeContext(*computeContext);
        hasContext = (bool)clContext.ctx;
    } else {
        try {
            clContext = makeComputeContext(ComputeDeviceType::GPU);
            hasContext = true;
        } catch (const ComputeError&) {
            hasContext = false;  // upstream prints "Error making compute context!" and carries on without one
        }
    }
    static std::atomic<int> counter{0};
    idAsset = assetId.empty() ? "mixer-" + std::to_string(++counter) : assetId;  // upstream: UUID().uuidString
}

bool VideoMixer::push(const PictureSample& pic) { return push(std::make_shared<const PictureSample>(pic)); }
bool VideoMixer::push(std::shared_ptr<const PictureSample> pic) {  // :57-75
    if (!hasContext) throw ComputeError(ErrorCode::badContextState, "No Compute Context");
    if (pic->assetId() != idAsset) {
        samples[0][pic->revision()] = std::move(pic);
        return true;
    }
    return false;
}

ComputeKernel VideoMixer::findKernel(const PictureSample* image, const PictureSample& target) const {  // :142-146
    const std::string inp = image ? pixelFormatName(image->pixelFormat()) : "clear";
    return defaultComputeKernelFromString("img_" + inp + "_" + pixelFormatName(target.pixelFormat()));
}

PictureSample VideoMixer::getBacking() {  // :148-165
    if (!hasContext) throw ComputeError(ErrorCode::badContextState, "No context");
    if ((int)backing.size() < numberBackingImages) {
        // upstream uploads an (uninitialised) CPU image to obtain the GPU planes; the bytes are never read because every compose starts
        // with the clear kernel, so only the plane layout and the device allocations are made here -- no host side at all (a download
        // brings its own page-locked buffers, compute.cpp: downloadComputePicture).
        PictureSample gpu;
        gpu.imgBuffer.planes = planesForFormat(backingFormat, backingSize);
        gpu.imgBuffer.pixelFormat = backingFormat;
        gpu.imgBuffer.size = backingSize;
        gpu.idAsset = idAsset, gpu.idWorkspace = idWorkspace, gpu.idRevision = idAsset;
        CtxGuard g(clContext.ctx);
        gpu.imgBuffer.computeTextures = allocPictureTextures(clContext, gpu.imgBuffer.planes);  // one block: an emitted frame downloads as one copy
        gpu.imgBuffer.bufferType = BufferType::gpu;
        gpu.done = std::make_shared<Event>(clContext.ctx);
        backing.push_back(gpu);
        return gpu;
    }
    PictureSample image = backing[currentBacking];
    currentBacking = (currentBacking + 1) % (int)backing.size();
    return image;
}

VideoMixer::Tick VideoMixer::beginTick() {  // :113-115
    Tick tk;
    tk.backing = getBacking();
    // merging { lhs, _ in lhs }: generation 0 wins; both maps are ordered by revision, so a two-way merge visits the union in
    // revision order without building a third map
    std::vector<std::pair<int, const std::shared_ptr<const PictureSample>*>> order;
    order.reserve(samples[0].size() + samples[1].size());
    auto a = samples[0].begin(), b = samples[1].begin();
    while (a != samples[0].end() || b != samples[1].end()) {
        const int c = a == samples[0].end() ? 1 : b == samples[1].end() ? -1 : a->first.compare(b->first);
        const std::shared_ptr<const PictureSample>& pick = c <= 0 ? a->second : b->second;
        order.emplace_back(pick->zIndex(), &pick);
        if (c <= 0) ++a;
        if (c >= 0) ++b;
    }
    // upstream's sort is unstable and the source is a dictionary: equal zIndex has no defined order there.
    // Here ties keep the revision order so that a frame is reproducible.
    std::stable_sort(order.begin(), order.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    tk.images.reserve(order.size());
    for (const auto& o : order) tk.images.push_back(*o.second);
    return tk;
}

void VideoMixer::endTick() {  // the `defer` block, :104-107
    samples[1].swap(samples[0]);
    samples[0].clear();
}

ComputeContext VideoMixer::composeRaw(const ComputeContext& ctxIn, const PictureSample& target, const std::vector<const PictureSample*>& layers,
                                      const ImageUniforms* uniforms, Mode mode) {
    const std::string tname = pixelFormatName(target.pixelFormat());
    // findKernel for the clear pass and every layer first: an unsupported pair fails before any work is queued
    const ComputeKernel clearKernel = defaultComputeKernelFromString("img_clear_" + tname);
    std::vector<ComputeKernel> kernels;
    for (const PictureSample* l : layers) kernels.push_back(defaultComputeKernelFromString(std::string("img_") + pixelFormatName(l->pixelFormat()) + "_" + tname));
    const int tf = svbFormat(target.pixelFormat());
    if (mode == Mode::perLayer || (tf != SVB_NV12 && tf != SVB_Y420P)) {
        // the reference's own sequence (mix.video.swift:117-124): clear, then one launch per layer
        ComputeContext ctx = runComputeKernel(ctxIn, {}, target, clearKernel, "", 3, nullptr, 0, false);
        for (size_t k = 0; k < layers.size(); ++k) ctx = runComputeKernel(ctx, {layers[k]}, target, kernels[k], "", 3, &uniforms[k], sizeof(ImageUniforms), true);
        return ctx;
    }
    for (ComputeKernel k : kernels)
        if (k == ComputeKernel::img_bgra_bgra) throw ComputeError(ErrorCode::computeKernelNotFound, "computeKernelNotFound(img_bgra_bgra)");
    std::vector<Job> jobs(1);
    jobs[0].target = &target;
    jobs[0].layers = layers;
    jobs[0].uniforms.assign(uniforms, uniforms + layers.size());
    composeFused(ctxIn, jobs, mode == Mode::generic, mode == Mode::fusedGather, mode == Mode::fusedTiled ? 1 : mode == Mode::fusedRing ? 3 : 0);
    return ctxIn;
}

PictureSample VideoMixer::mix(int64_t time, bool wait) {
    VideoMixer* self = this;
    PictureSample out;
    mixMany(&self, 1, time, &out, wait);
    return out;
}

void VideoMixer::mixMany(VideoMixer* const* mixers, int n, int64_t time, PictureSample* outs, bool wait) {
    if (n <= 0) return;
    for (int i = 0; i < n; ++i) {
        if (!mixers[i]->hasContext) throw ComputeError(ErrorCode::badContextState, "No context");
        if (mixers[i]->clContext.ctx != mixers[0]->clContext.ctx) throw ComputeError(ErrorCode::invalidContext, "mixMany: mixers must share one compute context");
    }
    struct EndTicks {  // the `defer` of mix(at:): generations rotate even when compose throws
        VideoMixer* const* m;
        int n;
        ~EndTicks() {
            for (int i = 0; i < n; ++i) m[i]->endTick();
        }
    } defer{mixers, n};

    SVB_PROF(0, "mixMany");
    std::vector<Tick> ticks;
    ticks.reserve(n);
    std::vector<Job> jobs;
    bool fusedAll = true;
    for (int i = 0; i < n; ++i) {
        SVB_PROF(1, "mixMany: beginTick (per mixer)");
        ticks.push_back(mixers[i]->beginTick());
        const int tf = svbFormat(ticks.back().backing.pixelFormat());
        fusedAll = fusedAll && mixers[i]->mode != Mode::perLayer && (tf == SVB_NV12 || tf == SVB_Y420P);
    }
    ComputeContext ctx0 = mixers[0]->clContext;
    if (fusedAll) {
        jobs.resize(n);
        std::vector<std::vector<SvbLayerDesc>> plannedNow(n);
        bool generic = false, wantGather = false;
        int want = 0;
        for (int i = 0; i < n; ++i) {
            SVB_PROF(2, "mixMany: plan-cache lookup (per mixer)");
            Tick& tk = ticks[i];
            VideoMixer& mx = *mixers[i];
            ++mx.tickNo;
            jobs[i].target = &tk.backing;
            jobs[i].planned = &plannedNow[i];
            mx.findKernel(nullptr, tk.backing);
            for (const auto& im : tk.images) {
                auto hit = mx.planned.find(im.get());
                if (hit != mx.planned.end() && hit->second.who.lock().get() == im.get()) {  // seen before (and still the same object): nothing to derive again
                    hit->second.tick = mx.tickNo;
                    jobs[i].layers.push_back(im.get());
                    jobs[i].uniforms.push_back(hit->second.uniforms);
                    jobs[i].cached.push_back(&hit->second.desc);
                    continue;
                }
                const ComputeKernel k = mx.findKernel(im.get(), tk.backing);  // throws invalidValue for an unknown pair, as upstream
                if (k == ComputeKernel::img_bgra_bgra) throw ComputeError(ErrorCode::computeKernelNotFound, "computeKernelNotFound(img_bgra_bgra)");
                jobs[i].layers.push_back(im.get());
                jobs[i].uniforms.push_back(makeImageUniforms(*im, tk.backing));
                jobs[i].cached.push_back(nullptr);
            }
            generic = generic || mixers[i]->mode == Mode::generic;
            wantGather = wantGather || mixers[i]->mode == Mode::fusedGather;
            if (mixers[i]->mode == Mode::fusedTiled) want = 1;
            if (mixers[i]->mode == Mode::fusedRing) want = 3;
        }
        composeFused(ctx0, jobs, generic, wantGather, want);
        SVB_PROF(11, "mixMany: cache update");
        for (int i = 0; i < n; ++i) {  // remember what was derived; forget the samples that are gone or have not been seen for a while
            VideoMixer& mx = *mixers[i];
            const Tick& tk = ticks[i];
            for (size_t k = 0; k < tk.images.size() && k < plannedNow[i].size(); ++k)
                if (!jobs[i].cached[k]) {
                    VideoMixer::Planned& p = mx.planned[tk.images[k].get()];
                    p.who = tk.images[k], p.uniforms = jobs[i].uniforms[k], p.desc = plannedNow[i][k], p.tick = mx.tickNo;
                }
            for (auto it = mx.planned.begin(); it != mx.planned.end();)
                it = (mx.tickNo - it->second.tick > 8u || it->second.who.expired()) ? mx.planned.erase(it) : std::next(it);
        }
    } else {
        for (int i = 0; i < n; ++i) {
            Tick& tk = ticks[i];
            std::vector<const PictureSample*> layers;
            std::vector<ImageUniforms> us;
            for (const auto& im : tk.images) {
                layers.push_back(im.get());
                us.push_back(makeImageUniforms(*im, tk.backing));
            }
            mixers[i]->clContext = composeRaw(mixers[i]->clContext, tk.backing, layers, us.data(), mixers[i]->mode);
        }
    }
    {
        SVB_PROF(12, "mixMany: outputs");
        CtxGuard g(ctx0.ctx);
        for (int i = 0; i < n; ++i) {
            PictureSample& b = ticks[i].backing;
            if (b.done) check(drv().cuEventRecord(b.done->e, ctx0.ctx->compute), "cuEventRecord");
            outs[i] = b;  // PictureSample(backing, pts:, time:) :127-130
            outs[i].timeValue = time;
            outs[i].ptsValue = time - mixers[i]->epoch;
            outs[i].timescale = mixers[i]->timescale;
        }
        if (wait) check(drv().cuStreamSynchronize(ctx0.ctx->compute), "cuStreamSynchronize");  // endComputePass(ctx, true)
    }
}

}  // namespace svb
