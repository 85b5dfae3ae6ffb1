// scale.cpp -- host side of the convert+scale operator (svb_scale_convert, kernels_scale.cuh): filter tables, launch.
//
// OURS, not upstream's: the reference has no high-bit-depth format (sample.pict.swift:19), no filter but the linear
// sampler (kernels.cl.swift:61) and no BGRA-writing kernel on Linux (compute.swift:54 names img_bgra_bgra only).
// BASELINE.json's configs 2 and 5 name the operator; oracle/scale_oracle.c defines what it computes.
#include <array>
#include <cmath>
#include <cstring>
#include <map>

#include "compute.h"
#include "cu_driver.h"
#include "svb_desc.h"

namespace svb {

static const CuDriver& drv() {
    const CuDriver& d = cu();
    if (!d.ok) throw ComputeError(ErrorCode::deviceNotAvailable, std::string("CUDA driver unavailable: ") + d.why);
    return d;
}

namespace {

constexpr int kTileW = SVB_SCALE_TW, kTileH = SVB_SCALE_TH;

double filterWeight(ScaleFilter f, double t) {
    t = std::fabs(t);
    if (f == ScaleFilter::bilinear) return t < 1.0 ? 1.0 - t : 0.0;
    if (t >= 3.0) return 0.0;
    if (t < 1e-9) return 1.0;
    const double pt = M_PI * t;
    return (std::sin(pt) / pt) * (std::sin(pt / 3.0) / (pt / 3.0));
}

struct DeviceTable {
    ScaleTable host;
    CUdeviceptr first = 0, weights = 0;
    // most source samples any aligned run of `run` consecutive outputs reaches; with chunk > 1 the window starts at a multiple
    // of `chunk` samples and covers whole chunks (the kernel stages 16-byte chunks)
    int span(int run, int chunk = 1) const {
        int best = 0;
        const int n = (int)host.first.size();
        for (int a = 0; a < n; a += run) {
            const int b = std::min(a + run, n) - 1;
            const int lo = host.first[(size_t)a] & ~(chunk - 1), hi = host.first[(size_t)b] + host.taps - 1;
            best = std::max(best, ((hi - lo) / chunk + 1) * chunk);
        }
        return best;
    }
};
struct ScaleShared {
    std::mutex mu;
    std::map<std::array<int, 3>, DeviceTable> tables;
    CUfunction fn = nullptr, fnAny = nullptr;  // compile-time window pitches / pitches from the descriptor
};

void freeScaleShared(InternalContext* ic) {
    auto* s = (ScaleShared*)ic->scaleShared;
    if (!s) return;
    for (auto& kv : s->tables) {
        if (kv.second.first) cu().cuMemFree(kv.second.first);
        if (kv.second.weights) cu().cuMemFree(kv.second.weights);
    }
    delete s;
    ic->scaleShared = nullptr;
}

ScaleShared& shared(const std::shared_ptr<InternalContext>& ic) {  // caller holds a CtxGuard
    std::lock_guard<std::mutex> g(ic->mu);
    if (!ic->scaleShared) {
        auto* s = new ScaleShared();
        s->fn = ic->builtin("svb_scale_convert");
        s->fnAny = ic->builtin("svb_scale_convert_any");
        check(drv().cuFuncSetAttribute(s->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, 200 * 1024), "cuFuncSetAttribute(max dynamic shared memory)");
        check(drv().cuFuncSetAttribute(s->fnAny, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, 200 * 1024), "cuFuncSetAttribute(max dynamic shared memory)");
        ic->scaleShared = s;
        ic->scaleSharedFree = freeScaleShared;
    }
    return *(ScaleShared*)ic->scaleShared;
}

// The table of one axis on the device (built once per (filter, srcN, dstN) and context; a synchronous upload).
const DeviceTable& deviceTable(const std::shared_ptr<InternalContext>& ic, ScaleShared& sh, ScaleFilter f, int srcN, int dstN) {
    std::lock_guard<std::mutex> g(sh.mu);
    const std::array<int, 3> key{(int)f, srcN, dstN};
    auto it = sh.tables.find(key);
    if (it != sh.tables.end()) return it->second;
    DeviceTable t;
    t.host = makeScaleTable(f, srcN, dstN);
    check(drv().cuMemAlloc(&t.first, sizeof(int32_t) * (size_t)dstN), "cuMemAlloc");
    check(drv().cuMemAlloc(&t.weights, sizeof(float) * t.host.weights.size()), "cuMemAlloc");
    check(drv().cuMemcpyHtoD(t.first, t.host.first.data(), sizeof(int32_t) * (size_t)dstN), "cuMemcpyHtoD");
    check(drv().cuMemcpyHtoD(t.weights, t.host.weights.data(), sizeof(float) * t.host.weights.size()), "cuMemcpyHtoD");
    (void)ic;
    return sh.tables.emplace(key, std::move(t)).first->second;
}

}  // namespace

// swscale's filter construction (libswscale/utils.c initFilter) in floating point: centre c = (x + 1/2) r - 1/2 with
// r = srcN / dstN, the kernel stretched by max(1, r) when minifying, 2 ceil(support) taps from floor(c - support) + 1,
// weights normalised to 1 in double and rounded to float.
ScaleTable makeScaleTable(ScaleFilter filter, int srcN, int dstN) {
    if (srcN < 1 || dstN < 1) throw ComputeError(ErrorCode::invalidValue, "makeScaleTable: empty axis");
    const double r = (double)srcN / (double)dstN, s = r > 1.0 ? r : 1.0;
    const double support = (filter == ScaleFilter::bilinear ? 1.0 : 3.0) * s;
    ScaleTable t;
    t.taps = 2 * (int)std::ceil(support);
    t.first.resize((size_t)dstN);
    t.weights.resize((size_t)dstN * (size_t)t.taps);
    std::vector<double> w((size_t)t.taps);
    for (int x = 0; x < dstN; ++x) {
        const double c = ((double)x + 0.5) * r - 0.5;
        const int f = (int)std::floor(c - support) + 1;
        double sum = 0.0;
        for (int k = 0; k < t.taps; ++k) {
            w[(size_t)k] = filterWeight(filter, ((double)(f + k) - c) / s);
            sum += w[(size_t)k];
        }
        t.first[(size_t)x] = f;
        for (int k = 0; k < t.taps; ++k) t.weights[(size_t)x * (size_t)t.taps + (size_t)k] = (float)(w[(size_t)k] / sum);
    }
    return t;
}

PictureSample scaleConvertPicture(const ComputeContext& ctx, const PictureSample& src, Vector2 dstSize, PixelFormat dstFormat, ScaleFilter filter, bool wait) {
    if (!ctx.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
    if (src.bufferType() != BufferType::gpu) throw ComputeError(ErrorCode::badInputData, "Input images must be uploaded to GPU");
    const PixelFormat sf = src.pixelFormat();
    if (sf != PixelFormat::nv12 && sf != PixelFormat::p010)
        throw ComputeError(ErrorCode::computeKernelNotFound, std::string("computeKernelNotFound(img_scale_") + pixelFormatName(sf) + "_" + pixelFormatName(dstFormat) + ")");
    if (dstFormat != PixelFormat::BGRA)
        throw ComputeError(ErrorCode::computeKernelNotFound, std::string("computeKernelNotFound(img_scale_") + pixelFormatName(sf) + "_" + pixelFormatName(dstFormat) + ")");
    const int srcW = (int)src.size().x, srcH = (int)src.size().y, dstW = (int)dstSize.x, dstH = (int)dstSize.y;
    if (srcW < 2 || srcH < 2 || (srcW & 1) || (srcH & 1)) throw ComputeError(ErrorCode::badInputData, "scaleConvertPicture: the source needs even, non-zero dimensions");
    if (dstW < 1 || dstH < 1) throw ComputeError(ErrorCode::badTarget, "badTarget");
    if (src.imgBuffer.computeTextures.size() < 2 || src.imgBuffer.planes.size() < 2) throw ComputeError(ErrorCode::badInputData, "Missing image buffer");

    CtxGuard g(ctx.ctx);
    const CuDriver& d = drv();
    InternalContext& ic = *ctx.ctx;
    ScaleShared& sh = shared(ctx.ctx);
    const DeviceTable& yx = deviceTable(ctx.ctx, sh, filter, srcW, dstW);
    const DeviceTable& yy = deviceTable(ctx.ctx, sh, filter, srcH, dstH);
    const DeviceTable& cx = deviceTable(ctx.ctx, sh, filter, srcW / 2, dstW);
    const DeviceTable& cy = deviceTable(ctx.ctx, sh, filter, srcH / 2, dstH);
    // Shared memory of one CTA: the horizontally filtered rows (luma + U + V, 64 floats each, row counts rounded up to 4) and the
    // transposed source window (luma, reused for the two chroma components).  The tile is 16 output rows high unless the vertical
    // footprint of a strong minification needs a shorter one to fit; beyond that the ratio is refused.
    const bool p010 = sf == PixelFormat::p010;
    const int chunkY = p010 ? 8 : 16, chunkC = p010 ? 4 : 8;  // samples / (U, V) pairs per 16 bytes
    const int spanYx = yx.span(kTileW, chunkY), spanCx = cx.span(kTileW, chunkC);
    auto pitchOf = [](int rows) {  // 16-byte aligned columns; an odd number of 16-byte groups spreads the columns over the banks
        int p = (rows + 3) & ~3;
        return ((p >> 2) & 1) ? p : p + 4;
    };
    int tileH = kTileH, spanYy = 0, spanCy = 0, pitchY = 0, pitchC = 0;
    size_t smem = 0;
    bool fixedPitch = false;
    for (;; tileH >>= 1) {
        spanYy = yy.span(tileH), spanCy = cy.span(tileH);
        // the kernel with compile-time pitches when the tile's rows fit them (every ratio down to 2 : 1 Lanczos-3 at 16 rows)
        fixedPitch = ((spanYy + 3) & ~3) <= SVB_SCALE_PITCH_Y && ((spanCy + 3) & ~3) <= SVB_SCALE_PITCH_C;
        pitchY = fixedPitch ? SVB_SCALE_PITCH_Y : pitchOf(spanYy), pitchC = fixedPitch ? SVB_SCALE_PITCH_C : pitchOf(spanCy);
        const size_t window = std::max((size_t)spanYx * (size_t)pitchY, 2 * (size_t)spanCx * (size_t)pitchC);
        smem = ((size_t)(2 * kTileH * 16) + (size_t)(((spanYy + 3) & ~3) + 2 * ((spanCy + 3) & ~3)) * SVB_SCALE_HP + window) * sizeof(float);
        if (smem <= 200 * 1024 || tileH == 1) break;
    }
    if (smem > 200 * 1024)
        throw ComputeError(ErrorCode::notImplemented, "scaleConvertPicture: the filter footprint of this ratio (" + std::to_string(yy.host.taps) + " x " +
                                                          std::to_string(yx.host.taps) + " taps) does not fit one tile");

    PictureSample out;
    out.imgBuffer.planes = planesForFormat(PixelFormat::BGRA, dstSize);
    out.imgBuffer.pixelFormat = PixelFormat::BGRA;
    out.imgBuffer.bufferType = BufferType::gpu;
    out.imgBuffer.size = dstSize;
    const size_t bytes = (size_t)out.imgBuffer.planes[0].stride * (size_t)dstH;
    auto tex = std::make_shared<ComputeBuffer>(ic.alloc(bytes), bytes, ctx.ctx);
    out.imgBuffer.computeTextures = {tex};
    out.idAsset = src.idAsset, out.idWorkspace = src.idWorkspace, out.idRevision = src.idRevision;
    out.ptsValue = src.ptsValue, out.timeValue = src.timeValue, out.timescale = src.timescale;

    SvbScaleDesc desc;
    std::memset(&desc, 0, sizeof(desc));
    desc.srcY = src.imgBuffer.computeTextures[0]->mem, desc.srcC = src.imgBuffer.computeTextures[1]->mem, desc.dst = tex->mem;
    desc.fYx = yx.first, desc.wYx = yx.weights, desc.fYy = yy.first, desc.wYy = yy.weights;
    desc.fCx = cx.first, desc.wCx = cx.weights, desc.fCy = cy.first, desc.wCy = cy.weights;
    desc.strideY = src.imgBuffer.planes[0].stride, desc.strideC = src.imgBuffer.planes[1].stride, desc.dstStride = out.imgBuffer.planes[0].stride;
    desc.srcW = srcW, desc.srcH = srcH, desc.dstW = dstW, desc.dstH = dstH;
    desc.format = sf == PixelFormat::p010 ? 1 : 0;
    desc.nYx = yx.host.taps, desc.nYy = yy.host.taps, desc.nCx = cx.host.taps, desc.nCy = cy.host.taps;
    desc.spanYy = spanYy, desc.spanCy = spanCy, desc.spanYx = spanYx, desc.spanCx = spanCx;
    desc.pitchY = pitchY, desc.pitchC = pitchC, desc.tileH = tileH;
    // 16-byte accesses need an aligned base and stride, and rows long enough to read the chunk that holds the last sample
    desc.vecY = (desc.srcY % 16 == 0 && desc.strideY % 16 == 0 && desc.strideY >= ((srcW + chunkY - 1) / chunkY) * 16) ? 1 : 0;
    desc.vecC = (desc.srcC % 16 == 0 && desc.strideC % 16 == 0 && desc.strideC >= ((srcW / 2 + chunkC - 1) / chunkC) * 16) ? 1 : 0;
    desc.vecDst = (desc.dst % 16 == 0 && desc.dstStride % 16 == 0) ? 1 : 0;

    if (src.done) check(d.cuStreamWaitEvent(ic.compute, src.done->e, 0), "cuStreamWaitEvent");
    for (const auto& t : src.imgBuffer.computeTextures) {
        waitReady(ic.compute, *t);
        t->lastUse = nullptr, t->lastUseUnknown = true;
    }
    void* args[] = {&desc};
    check(d.cuLaunchKernel(fixedPitch ? sh.fn : sh.fnAny, (unsigned)((dstW + kTileW - 1) / kTileW), (unsigned)((dstH + tileH - 1) / tileH), 1, 256, 1, 1, (unsigned)smem, ic.compute, args, nullptr),
          "cuLaunchKernel(svb_scale_convert)");
    noteKernelLaunch();
    markWritten(ctx, out);
    if (wait) {
        check(d.cuStreamSynchronize(ic.compute), "cuStreamSynchronize");
    } else {
        out.done = std::make_shared<Event>(ctx.ctx);
        check(d.cuEventRecord(out.done->e, ic.compute), "cuEventRecord");
    }
    return out;
}

}  // namespace svb
