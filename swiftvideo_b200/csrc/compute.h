// compute.h -- host side of the compute path, mirroring the reference's operator surface name for name:
//   compute.swift            ComputeError, ComputeKernel, ImageUniforms, defaultComputeKernelFromString,
//                            makeComputeContext, usingContext, applyComputeImage                (:22-170)
//   compute.cuda.swift       ComputeDevice, ComputeContext, ComputeBuffer, availableComputeDevices,
//                            createComputeContext, buildComputeKernel, runComputeKernel,
//                            begin/endComputePass, upload/downloadComputeBuffer/Picture         (:23-440)
//   sample.pict*.swift       PixelFormat, Component, Plane, BufferType, ImageBuffer, PictureSample,
//                            createPictureSample                                                (sample.pict.swift:20-101,
//                                                                                                sample.pict.linux.swift:23-311)
// The reference is Swift; no Swift toolchain exists in this image, so the host side is C++ above the same
// driver API and below the C ABI in include/svb200.h.
#pragma once
#include <atomic>
#include <cuda.h>
#include <stdint.h>

#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "vector_math.h"

namespace svb {

// ---- ComputeError (compute.swift:22-39) ---------------------------------------------------------------
enum class ErrorCode : int {
    ok = 0,
    invalidPlatform = 1,
    invalidDevice = 2,
    invalidOperation = 3,
    invalidValue = 4,
    invalidProgram = 5,
    invalidContext = 6,
    deviceNotAvailable = 7,
    outOfMemory = 8,
    compilerNotAvailable = 9,
    computeKernelNotFound = 10,
    badTarget = 11,
    badInputData = 12,
    badContextState = 13,
    compilerError = 14,
    unknownError = 15,
    notImplemented = 16,
};
struct ComputeError : std::runtime_error {
    ErrorCode code;
    ComputeError(ErrorCode c, const std::string& what) : std::runtime_error(what), code(c) {}
};
void check(CUresult r, const char* where);  // CUresult -> ComputeError, compute.cuda.swift:102-112

// ---- ComputeKernel (compute.swift:49-74) ---------------------------------------------------------------
enum class ComputeKernel : int {
    img_nv12_nv12 = 0,
    img_bgra_nv12,
    img_rgba_nv12,
    img_bgra_bgra,
    img_y420p_y420p,
    img_y420p_nv12,
    img_clear_nv12,
    img_clear_yuvs,
    img_clear_bgra,
    img_clear_y420p,
    img_clear_rgba,
    img_rgba_y420p,
    img_bgra_y420p,
    snd_s16i_s16i,
    me_fullsearch,
    custom,
    // ours: sources upstream's PixelFormat names but has no kernel for, under findKernel's own img_<source>_<target> rule;
    // after `custom` so that the upstream cases keep their order
    img_nv21_nv12,
    img_y422p_nv12,
    img_y444p_nv12,
    img_y422p_y420p,
    img_y444p_y420p,
    count_
};
const char* computeKernelName(ComputeKernel k);                          // String(describing:)
ComputeKernel defaultComputeKernelFromString(const std::string& name);  // compute.swift:90-110 (throws invalidValue)

// ---- picture data model --------------------------------------------------------------------------------
// p010 is ours (upstream: "TODO: Higher bit-depth formats", sample.pict.swift:19); it comes after `invalid` so that the upstream cases keep their values
enum class PixelFormat : int { nv12 = 0, nv21, yuvs, zvuy, y420p, y422p, y444p, RGBA, BGRA, shape, text, invalid, p010 };
enum class BufferType : int { shared = 0, cpu, gpu, invalid };
enum class Component : int { r, g, b, a, y, cr, cb };
const char* pixelFormatName(PixelFormat f);  // lower-cased case name, as VideoMixer.findKernel builds it

struct Plane {  // sample.pict.swift:46-56
    Vector2 size;
    int stride = 0;
    int bitDepth = 8;
    std::vector<Component> components;
};

struct InternalContext;  // owns the CUcontext + streams + pools
struct Event {           // a CUevent that dies with its last holder
    CUevent e = nullptr;
    std::shared_ptr<InternalContext> ctx;
    explicit Event(std::shared_ptr<InternalContext> c);
    ~Event();
    Event(const Event&) = delete;
};
class ComputeBuffer {    // compute.cuda.swift:75-92: owns device memory, released in the destructor
  public:
    ComputeBuffer(CUdeviceptr p, size_t size, std::shared_ptr<InternalContext> ctx) : mem(p), size(size), ctx(std::move(ctx)) {}
    // a plane inside a picture's single allocation: `whole` owns the memory and returns it to the pool when its last plane is gone.
    // Upstream allocates per plane (createTexture, compute.cuda.swift:413-431); one block per picture lets a picture whose host planes
    // lie back to back travel as ONE copy -- over a link busy in both directions, 64 copies a tick reach 1350 frames/s where 128 reach
    // 1240 (profiles/r2_copy_probe.log)
    ComputeBuffer(const std::shared_ptr<ComputeBuffer>& owner, size_t offset, size_t size) : mem(owner->mem + offset), size(size), whole(owner), ctx(owner->ctx) {}
    ~ComputeBuffer();
    ComputeBuffer(const ComputeBuffer&) = delete;
    CUdeviceptr mem;  // non-const: its address is what cuLaunchKernel's param array points at (:297)
    const size_t size;
    std::shared_ptr<Event> ready;  // recorded after the last async write (upload); consumers wait on it (waitReady)
    CUstream readyStream = nullptr;          // the stream `ready` was recorded on: work queued on that stream is ordered already
    std::atomic<bool> readyFired{false};     // `ready` was seen complete: a resident layer costs no driver call per tick after that
    void noteWrite(CUstream s) { readyStream = s, readyFired.store(false, std::memory_order_relaxed); }
    bool usedByDownload = false;   // the download stream has read this block (set by downloadComputeBuffer)
    std::shared_ptr<Event> lastRead;  // recorded after the last async READ on another stream (download): a writer waits on it before it overwrites (the mixer's backing ring)
    std::shared_ptr<Event> consumerRead;  // recorded by a reader outside this context's streams (an encoder's stream, a peer GPU's gather): writers and the pool wait on it
    // recorded on the compute stream right after the last compose that read this block (mix_video.cpp); nullptr = unknown.  With it the pool
    // orders a recycled block behind THAT launch instead of behind whatever the streams hold when the block is released -- two ticks later
    // for a layer, by when the compute stream's tail is a compose that itself waits for uploads still in flight (a 0.5 ms bubble per tick)
    std::shared_ptr<Event> lastUse;
    bool lastUseUnknown = false;  // some launch other than a compose read the block (per-layer kernels, the scale operator): use the tails
    std::shared_ptr<ComputeBuffer> whole;  // set on a plane carved out of a picture's allocation
    std::shared_ptr<uint8_t> hostKeep;  // source of an in-flight async upload stays alive with the texture
    std::shared_ptr<InternalContext> ctx;
};

// Host bytes of one plane: a slice of a shared allocation (buffersForPlanes slices one ByteBuffer, :296-311)
struct HostData {
    std::shared_ptr<uint8_t> base;  // keeps the allocation alive
    uint8_t* ptr = nullptr;
    size_t size = 0;
};

struct ImageBuffer {  // sample.pict.linux.swift:23-72
    PixelFormat pixelFormat = PixelFormat::invalid;
    BufferType bufferType = BufferType::invalid;
    Vector2 size;
    std::vector<std::shared_ptr<ComputeBuffer>> computeTextures;
    std::vector<HostData> buffers;
    std::vector<Plane> planes;
};

class PictureSample {  // sample.pict.linux.swift:105-249 (immutable upstream; copied-with-changes here too)
  public:
    ImageBuffer imgBuffer;
    Matrix4 transform, texTransform, borderTransform;
    Vector4 bgColor{0, 0, 0, 1};  // default fillColor, :158
    float alpha = 1.0f;
    std::string idAsset, idWorkspace, idRevision;
    int64_t ptsValue = 0, timeValue = 0;
    int64_t timescale = 1;
    std::shared_ptr<Event> done;  // set on samples produced asynchronously (mix / download): wait before reading

    Matrix4 matrix() const { return transform; }
    Matrix4 textureMatrix() const { return texTransform; }
    Matrix4 borderMatrix() const { return borderTransform; }
    Vector4 fillColor() const { return bgColor; }
    float opacity() const { return alpha; }
    Vector2 size() const { return imgBuffer.size; }
    PixelFormat pixelFormat() const { return imgBuffer.pixelFormat; }
    BufferType bufferType() const { return imgBuffer.bufferType; }
    const std::string& revision() const { return idRevision; }
    const std::string& assetId() const { return idAsset; }
    int zIndex() const { return (int)std::round((Vector3{0, 0, 0} * transform).z); }  // :116
};

std::vector<Plane> planesForFormat(PixelFormat f, Vector2 size);  // :275-294
// createPictureSample (:254-273): one contiguous CPU allocation sliced into planes. `pinned` (ours) asks
// for page-locked memory from `ctx` so uploads/downloads run at PCIe rate; the layout is unchanged.
PictureSample createPictureSample(Vector2 size, PixelFormat format, const std::string& assetId,
                                  const std::string& workspaceId, struct ComputeContext* pinnedFrom = nullptr, CUstream writer = nullptr);

// ImageBuffer(pixelFormat:bufferType:size:buffers:planes:) + PictureSample(img, ...) (sample.pict.linux.swift:23-39,160-189):
// a CPU sample over caller-described planes -- what the FFmpeg decoder builds with its own linesize as stride
// (SwiftVideo_FFmpeg/dec.video.ffmpeg.swift:144-220).  The bytes are copied (stride * rows per plane).
PictureSample pictureSampleFromPlanes(PixelFormat format, Vector2 size, const uint8_t* const* planes, const int* strides, int planeCount,
                                      const std::string& assetId, const std::string& workspaceId, struct ComputeContext* pinnedFrom = nullptr);

// ---- devices / context ---------------------------------------------------------------------------------
enum class ComputeDeviceType : int { GPU, CPU, Accelerator, Default };
struct ComputeDevice {  // compute.cuda.swift:23-30
    CUdevice device = 0;
    int index = 0;
    bool available = false;
    ComputeDeviceType deviceType = ComputeDeviceType::GPU;
    bool supportsImages = true;
};

struct CUDAProgram {  // compute.cuda.swift:43-58 -- a loaded module + one entry point
    CUmodule module = nullptr;
    CUfunction function = nullptr;
    bool ownsModule = false;
    std::shared_ptr<InternalContext> ctx;
    ~CUDAProgram();
};

struct ComputeContext {  // compute.cuda.swift:60-73: shared InternalContext + this holder's kernel library
    std::shared_ptr<InternalContext> ctx;
    std::map<std::string, std::shared_ptr<CUDAProgram>> library;
};

std::vector<ComputeDevice> availableComputeDevices();  // :132-153
// makeComputeContext(forType:) (compute.swift:121-129) picks devices.first; `deviceIndex` is the minimal
// extension SURVEY.md section 8e asks for so that streams can be placed one-per-GPU (default 0 = upstream behaviour).
ComputeContext makeComputeContext(ComputeDeviceType type, int deviceIndex = 0);
ComputeContext createComputeContext(const ComputeDevice& device);  // :159-165
ComputeContext createComputeContext(const ComputeContext& sharing);  // :155-157 (empty kernel library)
void destroyComputeContext(ComputeContext& ctx);                     // :167-169 (no-op upstream too)

// The sm_100a module holding every kernel (what buildComputeKernel's NVRTC step produced upstream).
void kernelModuleImage(const void** image, size_t* size);
// buildComputeKernel (:171-201).  Upstream compiles `source` with NVRTC then cuModuleLoadData +
// cuModuleGetFunction(name); here `image` is a ready cubin/PTX (NULL = our built-in module).
ComputeContext buildComputeKernel(const ComputeContext& ctx, const std::string& name, const void* image);
// buildComputeKernel(_:name:source:) (compute.cuda.swift:171-201) as upstream has it: CUDA C source -> NVRTC (--fmad=false, :177; the
// architecture option is sm_100a here: upstream's compute_30 is rejected by CUDA 12.9) -> cuModuleLoadData -> cuModuleGetFunction(name).
// The compiled kernel joins the context's library under `name` (merging { $1 }), where a custom(name:) launch or a built-in's name finds it.
// Fails with compilerNotAvailable when libnvrtc cannot be loaded, compilerError (with the NVRTC log) when the source does not compile.
ComputeContext buildComputeKernelFromSource(const ComputeContext& ctx, const std::string& name, const std::string& source);

ComputeContext beginComputePass(const ComputeContext& ctx);                        // :308-311
ComputeContext endComputePass(const ComputeContext& ctx, bool waitForCompletion);  // :313-319
template <class F>
ComputeContext usingContext(const ComputeContext& ctx, F&& fn) {  // compute.swift:131-134
    return endComputePass(fn(beginComputePass(ctx)), true);
}

std::shared_ptr<ComputeBuffer> uploadComputeBuffer(const ComputeContext& ctx, const void* src, size_t size,
                                                   std::shared_ptr<ComputeBuffer> dst);  // :330-342
void downloadComputeBuffer(const ComputeContext& ctx, const ComputeBuffer& src, void* dst, size_t dstSize);  // :344-357
// `wait` = upstream's behaviour (a synchronous cuMemcpyHtoD, then endComputePass(ctx, true), :339,:379): the source bytes may be
// reused as soon as the call returns.  wait=false (ours) queues the copies on the upload stream and returns at once with `done` set:
// the SOURCE buffers -- above all a page-locked staging buffer that a decoder refills -- must stay untouched until waitPicture(result).
PictureSample uploadComputePicture(const ComputeContext& ctx, const PictureSample& pict, int maxPlanes = 3,
                                   bool retainCpuBuffer = true, bool wait = true);        // :359-381
// `wait` = upstream's endComputePass(ctx, true) at :396; wait=false (ours) returns at once with `done` set.
PictureSample downloadComputePicture(const ComputeContext& ctx, const PictureSample& pict,
                                     bool retainGpuBuffer = false, bool wait = true);     // :383-402
// uploadComputePicture for several pictures at once (a tick's layers): CPU pictures that lie next to each other in page-locked host
// memory -- createPictureSample(pinnedFrom:) hands out neighbours when called in sequence -- and whose planes are 256-byte aligned in the
// tight layout share ONE device block and travel as ONE copy; the others are uploaded one by one.  GPU pictures pass through.
// (over a link busy in both directions a tick of 64 pictures reaches 1 350 frames/s picture by picture and 1 570 as one copy,
// profiles/r2_copy_probe.log)
std::vector<PictureSample> uploadComputePictures(const ComputeContext& ctx, const std::vector<const PictureSample*>& picts, int maxPlanes = 3,
                                                 bool retainCpuBuffer = true, bool wait = true);
void waitPicture(const PictureSample& pict);
// order stream `s` behind the last asynchronous write of `t`: nothing to do when that write was queued on `s` itself or is known to be
// complete (the caller holds a CtxGuard)
void waitReady(CUstream s, ComputeBuffer& t);
// Device planes for a picture with the given plane shapes: one allocation when every plane starts at a multiple of 256 bytes in the
// tight back-to-back layout (true of every even-sized 8-bit picture of video size), else one allocation per plane.
std::vector<std::shared_ptr<ComputeBuffer>> allocPictureTextures(const ComputeContext& ctx, const std::vector<Plane>& planes, int maxPlanes = 3);
// ---- device hand-off (upstream's commented-out h264_nvenc path, enc.video.ffmpeg.swift:169-170) ----------
// A consumer that reads a GPU sample where it lies (an encoder session on the same CUDA context) orders itself behind
// pictureReadyEvent() with cuStreamWaitEvent and, once its reads are queued, calls pictureConsumedOn(its stream): the mixer's
// backing ring and the block pool then wait for that point instead of the ten-tick convention of mix.video.swift:152-164.
CUevent pictureReadyEvent(const PictureSample& pict);                  // nullptr: nothing pending
void pictureConsumedOn(const PictureSample& pict, CUstream consumer);  // consumer: a stream of the sample's own context
// A GPU sample that lives on another device copied into `dst` (cuMemcpyPeerAsync on dst's upload stream -- NVLink when the
// devices are peers -- ordered behind the sample's completion; the source planes' next writer waits for the copy).
// A sample already on dst's device is returned as it is.  wait=false returns at once with `done` set.
PictureSample gatherComputePicture(const ComputeContext& dst, const PictureSample& pict, bool wait = true);  // block until `done` (if any) has fired

// GPUBarrierUpload / GPUBarrierDownload (compute.swift:175-198, :232-255): the pipeline stages around the two calls above.  A sample
// that already lives on the right side passes through untouched (`.just($0)`); a failure becomes the event error upstream emits:
// EventError("barrier.upload" | "barrier.download", -1, "<error>", assetId:).
struct EventError {
    std::string domain;
    int code = 0;
    std::string description;
    std::string assetId;
};
struct BarrierResult {
    bool ok = true;        // .just(sample) / .error(error)
    PictureSample sample;
    EventError error;
};
class GPUBarrierUpload {
  public:
    explicit GPUBarrierUpload(const ComputeContext& context, bool retainCpuBuffer = true) : context(createComputeContext(context)), retainCpuBuffer(retainCpuBuffer) {}
    BarrierResult operator()(const PictureSample& sample, bool wait = true) const;
  private:
    ComputeContext context;
    bool retainCpuBuffer;
};
class GPUBarrierDownload {
  public:
    explicit GPUBarrierDownload(const ComputeContext& context, bool retainGpuBuffer = true) : context(createComputeContext(context)), retainGpuBuffer(retainGpuBuffer) {}
    BarrierResult operator()(const PictureSample& sample, bool wait = true) const;
  private:
    ComputeContext context;
    bool retainGpuBuffer;
};

// runComputeKernel<T> (:260-306): params [outputs..., inputs..., uniforms, inStride[]], block (gcd(W,16),
// gcd(H,16)), grid (W/bx, H/by), 0 B smem.  `blends` is accepted and ignored exactly as upstream (:267).
ComputeContext runComputeKernel(const ComputeContext& ctx, const std::vector<const PictureSample*>& images,
                                const PictureSample& target, ComputeKernel kernel, const std::string& customName,
                                int maxPlanes, const void* uniforms, size_t uniformsSize, bool blends);

// ImageUniforms (compute.swift:76-86), 236 bytes, and applyComputeImage (:145-170)
struct ImageUniforms {
    float transform[16], textureTransform[16], borderMatrix[16];
    float fillColor[4];
    float inputSize[2], outputSize[2];
    float opacity, imageTime, targetTime;
};
static_assert(sizeof(ImageUniforms) == 236, "ImageUniforms must be 236 bytes");
ImageUniforms makeImageUniforms(const PictureSample& image, const PictureSample& target);
ComputeContext applyComputeImage(const ComputeContext& ctx, const PictureSample& image, const PictureSample& target,
                                 ComputeKernel kernel);

// ---- convert + scale (ours: no upstream counterpart; BASELINE.json configs 2 and 5 name it, oracle/scale_oracle.c defines it)
enum class ScaleFilter : int { bilinear = 0, lanczos3 = 1 };
// One axis of the resize: first[dstN] source index and n weights per output sample (swscale's filter construction in
// floating point; scale.cpp).  Exposed so that the tests can hold the library's tables against the oracle's.
struct ScaleTable {
    int taps = 0;
    std::vector<int32_t> first;
    std::vector<float> weights;
};
ScaleTable makeScaleTable(ScaleFilter filter, int srcN, int dstN);
// `src`: a GPU nv12 or p010 sample.  Returns a GPU BGRA sample of `dstSize` (a new texture); wait=false returns at once
// with `done` set.  Throws notImplemented for ratios whose filter needs more than 16 taps.
PictureSample scaleConvertPicture(const ComputeContext& ctx, const PictureSample& src, Vector2 dstSize, PixelFormat dstFormat, ScaleFilter filter,
                                  bool wait = true);

// After queueing a kernel that writes `target` on the compute stream: later consumers on other streams
// (downloads) order themselves behind it through the textures' `ready` events.
void markWritten(const ComputeContext& ctx, const PictureSample& target);
unsigned long long kernelLaunchCount();  // launches of our kernels issued by this process
void noteKernelLaunch();

// ---- plumbing shared with mix_video.cpp ----------------------------------------------------------------
struct InternalContext {
    CUcontext ctx = nullptr;
    CUdevice device = 0;
    int deviceIndex = 0;
    int smCount = 0;
    CUstream compute = nullptr, upload = nullptr, download = nullptr;
    CUmodule module = nullptr;  // built-in kernel module, loaded once per context
    std::mutex mu;
    struct Block {
        CUdeviceptr p = 0;
        CUevent after[3] = {nullptr, nullptr, nullptr};  // tails of compute/upload/download at release time
        std::shared_ptr<Event> consumer;                   // a foreign reader's completion (ComputeBuffer::consumerRead)
        std::shared_ptr<Event> lastUse;                    // ComputeBuffer::lastUse: stands in for the compute and upload tails
    };
    std::multimap<size_t, Block> pool;  // freed device blocks by size (upstream cuMemAllocs per upload)
    // page-locked host blocks by size (createPictureSample(pinnedFrom:), and the destination of a download that has none):
    // cuMemHostAlloc costs milliseconds, and a download into pageable memory runs at a fraction of the link rate
    struct HostBlock {
        void* p = nullptr;
        CUevent after[2] = {nullptr, nullptr};  // tails of upload/download at release time: an async copy may still touch it
    };
    std::multimap<size_t, HostBlock> hostPool;
    // page-locked memory is taken from the driver in large chunks and handed out front to back, so that pictures created one after the
    // other lie next to each other in host memory: uploadComputePictures then moves a whole run of them with ONE copy
    std::vector<void*> hostChunks;
    uint8_t* hostArena = nullptr;
    size_t hostArenaLeft = 0;
    // a pooled page-locked block.  writer == nullptr: the HOST will write it (waits until no copy still touches the recycled block);
    // writer = a stream: that stream's copies will write it (the stream is ordered behind those copies instead: nobody blocks)
    void* allocHost(size_t size, CUstream writer = nullptr);
    void releaseHost(void* p, size_t size);
    std::vector<CUevent> spareEvents;
    // state shared by every VideoMixer of this context (descriptor ring, tensor-map cache); owned by mix_video.cpp
    void* mixerShared = nullptr;
    void (*mixerSharedFree)(InternalContext*) = nullptr;
    // filter tables of the convert+scale operator by (filter, srcN, dstN); owned by scale.cpp
    void* scaleShared = nullptr;
    void (*scaleSharedFree)(InternalContext*) = nullptr;
    ~InternalContext();
    CUdeviceptr alloc(size_t size);
    void release(CUdeviceptr p, size_t size, bool usedByDownload = true, std::shared_ptr<Event> consumer = nullptr, std::shared_ptr<Event> lastUse = nullptr);  // usedByDownload=false: the block never met the download stream, whose tail it then need not wait for
    CUfunction builtin(const char* name);
};
struct CtxGuard {  // cuCtxPushCurrent / cuCtxPopCurrent pair
    explicit CtxGuard(const std::shared_ptr<InternalContext>& c);
    ~CtxGuard();
};

}  // namespace svb
