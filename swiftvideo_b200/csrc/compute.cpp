// compute.cpp -- host side of the compute path (see compute.h for the reference map).
#include "compute.h"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "cu_driver.h"

extern "C" {
extern const unsigned char svb200_kernels_cubin[];  // generated from kernels.cu by build.py
extern const unsigned long long svb200_kernels_cubin_len;
}

namespace svb {

static std::atomic<unsigned long long> g_launches{0};
unsigned long long kernelLaunchCount() { return g_launches.load(); }
void noteKernelLaunch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- errors ---------------------------------------------------------------------------------------------

void check(CUresult r, const char* where) {  // compute.cuda.swift:102-112
    if (r == CUDA_SUCCESS) return;
    const char* s = nullptr;
    if (cu().ok) cu().cuGetErrorString(r, &s);
    std::string msg = std::string(where) + ": " + (s ? s : "CUDA error") + " (" + std::to_string((int)r) + ")";
    switch (r) {
    case CUDA_ERROR_INVALID_VALUE: throw ComputeError(ErrorCode::invalidValue, msg);
    case CUDA_ERROR_OUT_OF_MEMORY: throw ComputeError(ErrorCode::outOfMemory, msg);
    case CUDA_ERROR_INVALID_CONTEXT: throw ComputeError(ErrorCode::invalidContext, msg);
    case CUDA_ERROR_ILLEGAL_ADDRESS: throw ComputeError(ErrorCode::badContextState, "Illegal address access: " + msg);
    case CUDA_ERROR_NOT_FOUND: throw ComputeError(ErrorCode::badInputData, "Symbol not found: " + msg);
    default: throw ComputeError(ErrorCode::unknownError, msg);
    }
}

static const CuDriver& drv() {
    const CuDriver& d = cu();
    if (!d.ok) throw ComputeError(ErrorCode::deviceNotAvailable, std::string("CUDA driver unavailable: ") + d.why);
    return d;
}

// ---- kernel names ---------------------------------------------------------------------------------------

static const char* kKernelNames[] = {"img_nv12_nv12",  "img_bgra_nv12",  "img_rgba_nv12",  "img_bgra_bgra",   "img_y420p_y420p",
                                     "img_y420p_nv12", "img_clear_nv12", "img_clear_yuvs", "img_clear_bgra",  "img_clear_y420p",
                                     "img_clear_rgba", "img_rgba_y420p", "img_bgra_y420p", "snd_s16i_s16i",   "me_fullsearch",
                                     "custom",         "img_nv21_nv12",  "img_y422p_nv12", "img_y444p_nv12",  "img_y422p_y420p",
                                     "img_y444p_y420p"};

const char* computeKernelName(ComputeKernel k) { return kKernelNames[(int)k]; }

ComputeKernel defaultComputeKernelFromString(const std::string& name) {  // compute.swift:90-110
    static const std::map<std::string, ComputeKernel> m = {
        {"img_nv12_nv12", ComputeKernel::img_nv12_nv12},     {"img_bgra_nv12", ComputeKernel::img_bgra_nv12},
        {"img_rgba_nv12", ComputeKernel::img_rgba_nv12},     {"img_bgra_bgra", ComputeKernel::img_bgra_bgra},
        {"img_y420p_y420p", ComputeKernel::img_y420p_y420p}, {"img_y420p_nv12", ComputeKernel::img_y420p_nv12},
        {"img_clear_nv12", ComputeKernel::img_clear_nv12},   {"img_clear_yuvs", ComputeKernel::img_clear_yuvs},
        {"img_clear_bgra", ComputeKernel::img_clear_bgra},   {"img_clear_rgba", ComputeKernel::img_clear_bgra},  // :101
        {"img_rgba_y420p", ComputeKernel::img_rgba_y420p},   {"img_bgra_y420p", ComputeKernel::img_bgra_y420p},
        {"img_clear_y420p", ComputeKernel::img_clear_y420p},
        // ours: the sources upstream names without a kernel
        {"img_nv21_nv12", ComputeKernel::img_nv21_nv12},     {"img_y422p_nv12", ComputeKernel::img_y422p_nv12},
        {"img_y444p_nv12", ComputeKernel::img_y444p_nv12},   {"img_y422p_y420p", ComputeKernel::img_y422p_y420p},
        {"img_y444p_y420p", ComputeKernel::img_y444p_y420p}};
    auto it = m.find(name);
    if (it == m.end()) throw ComputeError(ErrorCode::invalidValue, "no default compute kernel named " + name);
    return it->second;
}

const char* pixelFormatName(PixelFormat f) {
    static const char* n[] = {"nv12", "nv21", "yuvs", "zvuy", "y420p", "y422p", "y444p", "rgba", "bgra", "shape", "text", "invalid", "p010"};
    return n[(int)f];
}

// ---- context --------------------------------------------------------------------------------------------

CtxGuard::CtxGuard(const std::shared_ptr<InternalContext>& c) { check(drv().cuCtxPushCurrent(c->ctx), "cuCtxPushCurrent"); }
CtxGuard::~CtxGuard() {
    CUcontext old;
    cu().cuCtxPopCurrent(&old);
}

InternalContext::~InternalContext() {
    if (!ctx || !cu().ok) return;
    cu().cuCtxPushCurrent(ctx);
    cu().cuCtxSynchronize();
    if (mixerSharedFree) mixerSharedFree(this);
    if (scaleSharedFree) scaleSharedFree(this);
    for (auto& kv : pool) {
        cu().cuMemFree(kv.second.p);
        for (CUevent e : kv.second.after)
            if (e) cu().cuEventDestroy(e);
    }
    for (auto& kv : hostPool)  // (the blocks are pieces of the chunks below)
        for (CUevent e : kv.second.after)
            if (e) cu().cuEventDestroy(e);
    for (void* c : hostChunks) cu().cuMemFreeHost(c);
    for (CUevent e : spareEvents) cu().cuEventDestroy(e);
    if (module) cu().cuModuleUnload(module);
    if (compute) cu().cuStreamDestroy(compute);
    if (upload) cu().cuStreamDestroy(upload);
    if (download) cu().cuStreamDestroy(download);
    CUcontext old;
    cu().cuCtxPopCurrent(&old);
    cu().cuDevicePrimaryCtxRelease(device);
}

// Freed blocks are recycled (upstream cuMemAllocs per upload and frees in deinit).  A block can be released
// while work that touches it is still queued, so release() records the tail of each stream and alloc()
// makes every stream wait for those tails before the block's next user can touch it.  By the time a block
// comes round again the events have normally fired and the waits cost nothing.
CUdeviceptr InternalContext::alloc(size_t size) {
    size = (size + 255) & ~(size_t)255;
    Block b;
    bool hit = false;
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = pool.find(size);
        if (it != pool.end()) {
            b = it->second;
            pool.erase(it);
            hit = true;
        }
    }
    if (hit) {
        CUstream all[3] = {compute, upload, download};
        for (CUevent e : b.after) {
            if (!e) continue;
            if (cu().cuEventQuery(e) != CUDA_SUCCESS)
                for (CUstream s : all) cu().cuStreamWaitEvent(s, e, 0);
            std::lock_guard<std::mutex> g(mu);
            spareEvents.push_back(e);
        }
        if (b.consumer && cu().cuEventQuery(b.consumer->e) != CUDA_SUCCESS)
            for (CUstream s : all) cu().cuStreamWaitEvent(s, b.consumer->e, 0);
        if (b.lastUse && cu().cuEventQuery(b.lastUse->e) != CUDA_SUCCESS)
            for (CUstream s : all) cu().cuStreamWaitEvent(s, b.lastUse->e, 0);
        return b.p;
    }
    CUdeviceptr p = 0;
    check(drv().cuMemAlloc(&p, size), "cuMemAlloc");
    if (std::getenv("SVB_DEBUG_POOL")) fprintf(stderr, "[svb] cuMemAlloc %zu\n", size);
    return p;
}
void InternalContext::release(CUdeviceptr p, size_t size, bool usedByDownload, std::shared_ptr<Event> consumer, std::shared_ptr<Event> lastUse) {
    size = (size + 255) & ~(size_t)255;
    Block b;
    b.p = p;
    b.consumer = std::move(consumer);
    b.lastUse = std::move(lastUse);
    bool lastUseKnown = (bool)b.lastUse;
    if (cu().ok && ctx) {
        cu().cuCtxPushCurrent(ctx);
        // (an Event refers to its context: one that has fired is dropped here rather than parked in the context's own pool)
        if (b.consumer && cu().cuEventQuery(b.consumer->e) == CUDA_SUCCESS) b.consumer = nullptr;
        if (b.lastUse && cu().cuEventQuery(b.lastUse->e) == CUDA_SUCCESS) b.lastUse = nullptr;
        CUstream all[3] = {compute, upload, download};
        for (int i = 0; i < 3; ++i) {
            // (a block the download stream never read -- an uploaded layer -- does not wait for that stream's tail: it would order the next
            // tick's uploads behind this tick's downloads and halve the link's duplex rate)
            if (i == 2 && !usedByDownload) continue;
            if (i < 2 && lastUseKnown) continue;  // the last compose that read the block is known: later work on these streams never touched it
            CUevent e = nullptr;
            {
                std::lock_guard<std::mutex> g(mu);
                if (!spareEvents.empty()) {
                    e = spareEvents.back();
                    spareEvents.pop_back();
                }
            }
            if (!e && cu().cuEventCreate(&e, CU_EVENT_DISABLE_TIMING) != CUDA_SUCCESS) e = nullptr;
            if (e && cu().cuEventRecord(e, all[i]) == CUDA_SUCCESS) b.after[i] = e;
        }
        CUcontext old;
        cu().cuCtxPopCurrent(&old);
    }
    std::lock_guard<std::mutex> g(mu);
    pool.emplace(size, b);
}

// caller holds a CtxGuard
void* InternalContext::allocHost(size_t size, CUstream writer) {
    size = (size + 4095) & ~(size_t)4095;
    HostBlock b;
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = hostPool.find(size);
        if (it != hostPool.end()) {
            b = it->second;
            hostPool.erase(it);
        }
    }
    if (b.p) {
        for (CUevent e : b.after) {
            if (!e) continue;
            if (writer) cu().cuStreamWaitEvent(writer, e, 0);  // the block is written by `writer`'s copies: order them behind whatever still reads it
            else cu().cuEventSynchronize(e);                   // the host writes it: normally long fired
            std::lock_guard<std::mutex> g(mu);
            spareEvents.push_back(e);
        }
        return b.p;
    }
    std::lock_guard<std::mutex> g(mu);
    if (hostArenaLeft < size) {  // (what is left of the old chunk stays unused: at most one picture's worth)
        const size_t chunk = std::max<size_t>(size, (size_t)128 << 20);
        void* c = nullptr;
        check(drv().cuMemHostAlloc(&c, chunk, 0), "cuMemHostAlloc");
        if (std::getenv("SVB_DEBUG_POOL")) fprintf(stderr, "[svb] cuMemHostAlloc %zu\n", chunk);
        hostChunks.push_back(c);
        hostArena = (uint8_t*)c, hostArenaLeft = chunk;
    }
    void* p = hostArena;
    hostArena += size, hostArenaLeft -= size;
    return p;
}
void InternalContext::releaseHost(void* p, size_t size) {
    size = (size + 4095) & ~(size_t)4095;
    if (!cu().ok || !ctx) return;
    HostBlock b;
    b.p = p;
    cu().cuCtxPushCurrent(ctx);
    CUstream two[2] = {upload, download};
    for (int i = 0; i < 2; ++i) {
        CUevent e = nullptr;
        {
            std::lock_guard<std::mutex> g(mu);
            if (!spareEvents.empty()) {
                e = spareEvents.back();
                spareEvents.pop_back();
            }
        }
        if (!e && cu().cuEventCreate(&e, CU_EVENT_DISABLE_TIMING) != CUDA_SUCCESS) e = nullptr;
        if (e && cu().cuEventRecord(e, two[i]) == CUDA_SUCCESS) b.after[i] = e;
    }
    CUcontext old;
    cu().cuCtxPopCurrent(&old);
    std::lock_guard<std::mutex> g(mu);
    hostPool.emplace(size, b);
}

CUfunction InternalContext::builtin(const char* name) {
    CUfunction f = nullptr;
    check(drv().cuModuleGetFunction(&f, module, name), name);
    return f;
}

// Events are recycled through the context's spare list: a tick creates and drops a handful (completion of the mix, last use of the
// layers, last read of a download) and cuEventCreate / cuEventDestroy each cost a context push and a driver call.  Re-recording a
// recycled event is safe: a wait already queued on it refers to the record that was current when the wait was queued.
Event::Event(std::shared_ptr<InternalContext> c) : ctx(std::move(c)) {
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        if (!ctx->spareEvents.empty()) {
            e = ctx->spareEvents.back();
            ctx->spareEvents.pop_back();
        }
    }
    if (e) return;
    CtxGuard g(ctx);
    check(drv().cuEventCreate(&e, CU_EVENT_DISABLE_TIMING), "cuEventCreate");
}
Event::~Event() {
    if (e && ctx) {
        std::lock_guard<std::mutex> g(ctx->mu);
        if (ctx->ctx && ctx->spareEvents.size() < 4096) {
            ctx->spareEvents.push_back(e);
            e = nullptr;
        }
    }
    if (e && cu().ok) {
        cu().cuCtxPushCurrent(ctx->ctx);
        cu().cuEventDestroy(e);
        CUcontext old;
        cu().cuCtxPopCurrent(&old);
    }
}

ComputeBuffer::~ComputeBuffer() {  // compute.cuda.swift:82-88
    if (whole) {  // a plane of a picture's allocation: what the pool must know travels to the owner
        whole->usedByDownload = whole->usedByDownload || usedByDownload;
        if (consumerRead) whole->consumerRead = consumerRead;
        if (lastUse) whole->lastUse = lastUse;
        else whole->lastUseUnknown = true;
        return;
    }
    if (mem && ctx) ctx->release(mem, size, usedByDownload, consumerRead, lastUseUnknown ? nullptr : lastUse);
}

CUDAProgram::~CUDAProgram() {
    if (ownsModule && module && cu().ok && ctx) {
        cu().cuCtxPushCurrent(ctx->ctx);
        cu().cuModuleUnload(module);
        CUcontext old;
        cu().cuCtxPopCurrent(&old);
    }
}

std::vector<ComputeDevice> availableComputeDevices() {  // compute.cuda.swift:132-153
    std::vector<ComputeDevice> out;
    if (!cu().ok) return out;  // upstream: catch -> []
    int n = 0;
    if (cu().cuDeviceGetCount(&n) != CUDA_SUCCESS) return out;
    for (int i = 0; i < n; ++i) {
        ComputeDevice d;
        int mode = 0;
        if (cu().cuDeviceGet(&d.device, i) != CUDA_SUCCESS) return {};
        if (cu().cuDeviceGetAttribute(&mode, CU_DEVICE_ATTRIBUTE_COMPUTE_MODE, d.device) != CUDA_SUCCESS) return {};
        d.index = i;
        d.available = mode == CU_COMPUTEMODE_DEFAULT;
        out.push_back(d);
    }
    return out;
}

void kernelModuleImage(const void** image, size_t* size) {
    *image = svb200_kernels_cubin;
    *size = (size_t)svb200_kernels_cubin_len;
}

ComputeContext createComputeContext(const ComputeDevice& device) {  // compute.cuda.swift:159-165
    const CuDriver& d = drv();
    auto ic = std::make_shared<InternalContext>();
    ic->device = device.device;
    ic->deviceIndex = device.index;
    // Upstream: cuCtxCreate_v2.  We retain the device's primary context instead so that memory and streams
    // interoperate with anything else in the process that uses the CUDA runtime (torch in tests/bench).
    check(d.cuDevicePrimaryCtxRetain(&ic->ctx, device.device), "cuDevicePrimaryCtxRetain");
    CtxGuard g(ic);
    check(d.cuDeviceGetAttribute(&ic->smCount, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, device.device), "cuDeviceGetAttribute");
    check(d.cuStreamCreate(&ic->compute, CU_STREAM_NON_BLOCKING), "cuStreamCreate");
    check(d.cuStreamCreate(&ic->upload, CU_STREAM_NON_BLOCKING), "cuStreamCreate");
    check(d.cuStreamCreate(&ic->download, CU_STREAM_NON_BLOCKING), "cuStreamCreate");
    CUresult r = d.cuModuleLoadData(&ic->module, svb200_kernels_cubin);
    if (r != CUDA_SUCCESS)
        throw ComputeError(ErrorCode::invalidProgram,
                           "cuModuleLoadData(svb200 sm_100a module) failed (" + std::to_string((int)r) +
                               "): this build runs on B200 (sm_100a) only");
    ComputeContext c;
    c.ctx = ic;
    return c;
}

ComputeContext createComputeContext(const ComputeContext& sharing) {  // :155-157
    ComputeContext c;
    c.ctx = sharing.ctx;
    return c;
}

void destroyComputeContext(ComputeContext& ctx) {  // :167-169: released with the last reference
    ctx.library.clear();
    ctx.ctx.reset();
}

ComputeContext makeComputeContext(ComputeDeviceType type, int deviceIndex) {  // compute.swift:121-129
    std::vector<ComputeDevice> devs;
    for (const ComputeDevice& d : availableComputeDevices())
        if (d.deviceType == type && d.available) devs.push_back(d);
    if (deviceIndex < 0 || deviceIndex >= (int)devs.size())
        throw ComputeError(ErrorCode::deviceNotAvailable, cu().ok ? "no such compute device" : std::string("no compute device: ") + cu().why);
    return createComputeContext(devs[deviceIndex]);
}

ComputeContext buildComputeKernel(const ComputeContext& ctx, const std::string& name, const void* image) {  // :171-201
    if (!ctx.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
    CtxGuard g(ctx.ctx);
    auto prog = std::make_shared<CUDAProgram>();
    prog->ctx = ctx.ctx;
    if (image) {
        check(drv().cuModuleLoadData(&prog->module, image), "cuModuleLoadData");
        prog->ownsModule = true;
    } else {
        prog->module = ctx.ctx->module;
    }
    CUresult r = drv().cuModuleGetFunction(&prog->function, prog->module, name.c_str());
    if (r == CUDA_ERROR_NOT_FOUND) throw ComputeError(ErrorCode::badInputData, "Symbol not found: " + name);
    check(r, "cuModuleGetFunction");
    ComputeContext out = ctx;
    out.library[name] = prog;  // merging { $1 }: the new program replaces an older one of the same name
    return out;
}

// ---- NVRTC, bound lazily like the driver (dlopen: the library itself must load on a box without the toolkit) ------------------
namespace {
struct Nvrtc {
    void* h = nullptr;
    int (*createProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*compileProgram)(void*, int, const char* const*) = nullptr;
    int (*getPTXSize)(void*, size_t*) = nullptr;
    int (*getPTX)(void*, char*) = nullptr;
    int (*getCUBINSize)(void*, size_t*) = nullptr;
    int (*getCUBIN)(void*, char*) = nullptr;
    int (*getLogSize)(void*, size_t*) = nullptr;
    int (*getLog)(void*, char*) = nullptr;
    int (*destroyProgram)(void**) = nullptr;
    bool ok = false;
};
const Nvrtc& nvrtc() {
    static const Nvrtc n = [] {
        Nvrtc r;
        for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.13"})
            if ((r.h = dlopen(name, RTLD_NOW | RTLD_LOCAL))) break;
        if (!r.h) return r;
        auto sym = [&](const char* s) { return dlsym(r.h, s); };
        r.createProgram = (decltype(r.createProgram))sym("nvrtcCreateProgram");
        r.compileProgram = (decltype(r.compileProgram))sym("nvrtcCompileProgram");
        r.getPTXSize = (decltype(r.getPTXSize))sym("nvrtcGetPTXSize");
        r.getPTX = (decltype(r.getPTX))sym("nvrtcGetPTX");
        r.getCUBINSize = (decltype(r.getCUBINSize))sym("nvrtcGetCUBINSize");
        r.getCUBIN = (decltype(r.getCUBIN))sym("nvrtcGetCUBIN");
        r.getLogSize = (decltype(r.getLogSize))sym("nvrtcGetProgramLogSize");
        r.getLog = (decltype(r.getLog))sym("nvrtcGetProgramLog");
        r.destroyProgram = (decltype(r.destroyProgram))sym("nvrtcDestroyProgram");
        r.ok = r.createProgram && r.compileProgram && r.getPTXSize && r.getPTX && r.getLogSize && r.getLog && r.destroyProgram;
        return r;
    }();
    return n;
}
}  // namespace

ComputeContext buildComputeKernelFromSource(const ComputeContext& ctx, const std::string& name, const std::string& source) {  // :171-201
    if (!ctx.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
    const Nvrtc& rt = nvrtc();
    if (!rt.ok) throw ComputeError(ErrorCode::compilerNotAvailable, "compilerNotAvailable: libnvrtc could not be loaded");
    void* prog = nullptr;
    if (rt.createProgram(&prog, source.c_str(), name.c_str(), 0, nullptr, nullptr) != 0) throw ComputeError(ErrorCode::invalidProgram, "nvrtcCreateProgram failed");
    struct Destroy {
        const Nvrtc& rt;
        void** p;
        ~Destroy() { rt.destroyProgram(p); }
    } destroy{rt, &prog};
    // upstream: ["--gpu-architecture=compute_30", "--fmad=false"] (:177).  A real architecture (sm_100a) yields a cubin directly; every
    // multiply and add must round on its own, as in every other kernel of the path.
    const char* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "-default-device"};
    const int rc = rt.compileProgram(prog, 3, opts);
    if (rc != 0) {
        size_t n = 0;
        std::string log;
        if (rt.getLogSize(prog, &n) == 0 && n > 1) {
            log.resize(n);
            rt.getLog(prog, &log[0]);
        }
        throw ComputeError(ErrorCode::compilerError, "compilerError(" + name + "): " + log);
    }
    std::vector<char> image;
    size_t n = 0;
    if (rt.getCUBINSize && rt.getCUBIN && rt.getCUBINSize(prog, &n) == 0 && n > 0) {
        image.resize(n);
        if (rt.getCUBIN(prog, image.data()) != 0) image.clear();
    }
    if (image.empty()) {  // (a virtual architecture: PTX, finished by the driver's JIT)
        if (rt.getPTXSize(prog, &n) != 0 || n == 0) throw ComputeError(ErrorCode::invalidProgram, "nvrtcGetPTXSize failed");
        image.resize(n);
        if (rt.getPTX(prog, image.data()) != 0) throw ComputeError(ErrorCode::invalidProgram, "nvrtcGetPTX failed");
    }
    return buildComputeKernel(ctx, name, image.data());
}

// maybeBuildKernel (:203-218): a name already in the library wins; otherwise the built-in of that name is
// loaded; a name with no built-in throws computeKernelNotFound.
static ComputeContext maybeBuildKernel(const ComputeContext& ctx, ComputeKernel kernel, const std::string& customName) {
    const std::string name = kernel == ComputeKernel::custom ? customName : computeKernelName(kernel);
    if (ctx.library.count(name)) return ctx;
    static const char* builtins[] = {"img_clear_nv12", "img_clear_y420p", "img_clear_bgra", "img_nv12_nv12",
                                     "img_y420p_nv12", "img_y420p_y420p", "img_bgra_nv12",  "img_rgba_nv12",
                                     "img_bgra_y420p", "img_rgba_y420p",
                                     // ours: the operators upstream names without a Linux kernel (kernels_dropin.cuh)
                                     "img_bgra_bgra",  "img_clear_yuvs",  "img_nv21_nv12",  "img_y422p_nv12",
                                     "img_y444p_nv12", "img_y422p_y420p", "img_y444p_y420p"};
    bool have = false;
    for (const char* b : builtins) have = have || name == b;
    if (!have) throw ComputeError(ErrorCode::computeKernelNotFound, "computeKernelNotFound(" + name + ")");
    return buildComputeKernel(ctx, name, nullptr);
}

ComputeContext beginComputePass(const ComputeContext& ctx) {  // :308-311
    if (!ctx.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
    check(drv().cuCtxPushCurrent(ctx.ctx->ctx), "cuCtxPushCurrent");
    return ctx;
}
ComputeContext endComputePass(const ComputeContext& ctx, bool waitForCompletion) {  // :313-319
    if (waitForCompletion) {
        // upstream: cuCtxSynchronize. Equivalent for our work: drain the streams this context launches on.
        cu().cuStreamSynchronize(ctx.ctx->upload);
        cu().cuStreamSynchronize(ctx.ctx->compute);
        cu().cuStreamSynchronize(ctx.ctx->download);
    }
    CUcontext old;
    cu().cuCtxPopCurrent(&old);
    return ctx;
}

// ---- buffers --------------------------------------------------------------------------------------------

static std::shared_ptr<ComputeBuffer> createBuffer(const ComputeContext& ctx, size_t size) {  // :404-410
    CtxGuard g(ctx.ctx);
    return std::make_shared<ComputeBuffer>(ctx.ctx->alloc(size), size, ctx.ctx);
}

std::shared_ptr<ComputeBuffer> uploadComputeBuffer(const ComputeContext& ctx, const void* src, size_t size,
                                                   std::shared_ptr<ComputeBuffer> dst) {  // :330-342
    if (!src) throw ComputeError(ErrorCode::invalidValue, "uploadComputeBuffer: null source");
    auto buf = dst ? dst : createBuffer(ctx, size);
    if (buf->size < size) throw ComputeError(ErrorCode::badInputData, "Compute buffer needs to be >= to data.count");
    CtxGuard g(ctx.ctx);
    check(drv().cuMemcpyHtoDAsync(buf->mem, src, size, ctx.ctx->upload), "cuMemcpyHtoDAsync");
    if (!buf->ready) buf->ready = std::make_shared<Event>(ctx.ctx);
    check(drv().cuEventRecord(buf->ready->e, ctx.ctx->upload), "cuEventRecord");
    buf->noteWrite(ctx.ctx->upload);
    return buf;
}

void downloadComputeBuffer(const ComputeContext& ctx, const ComputeBuffer& src, void* dst, size_t dstSize) {  // :344-357
    if (dstSize < src.size) throw ComputeError(ErrorCode::badInputData, "Destination data buffer must be >= buffer.size");
    CtxGuard g(ctx.ctx);
    const_cast<ComputeBuffer&>(src).usedByDownload = true;
    check(drv().cuMemcpyDtoHAsync(dst, src.mem, src.size, ctx.ctx->download), "cuMemcpyDtoHAsync");
}

// ---- pictures -------------------------------------------------------------------------------------------

std::vector<Plane> planesForFormat(PixelFormat f, Vector2 size) {  // sample.pict.linux.swift:275-294
    const int width = (int)size.x;
    const Vector2 half{size.x / 2, size.y / 2};
    switch (f) {
    case PixelFormat::nv12:
        return {Plane{size, width, 8, {Component::y}}, Plane{half, width, 8, {Component::cb, Component::cr}}};
    case PixelFormat::p010:  // ours: 16-bit little-endian words, ten bits in the MSBs; same plane shapes as nv12
        return {Plane{size, width * 2, 10, {Component::y}}, Plane{half, width * 2, 10, {Component::cb, Component::cr}}};
    case PixelFormat::BGRA:
    case PixelFormat::RGBA:
        return {Plane{size, width * 4, 8, {Component::r, Component::g, Component::b, Component::a}}};
    case PixelFormat::yuvs:
        return {Plane{size, width * 2, 8, {Component::cr, Component::y, Component::cb, Component::y}}};
    case PixelFormat::zvuy:
        return {Plane{size, width * 2, 8, {Component::y, Component::cb, Component::y, Component::cr}}};
    case PixelFormat::y420p:
        return {Plane{size, width, 8, {Component::y}}, Plane{half, width / 2, 8, {Component::cb}}, Plane{half, width / 2, 8, {Component::cr}}};
    // ours: upstream's planesForFormat throws for these three (it has no kernels for them); the layouts follow componentsForPlane
    // (sample.pict.swift:84-90) and the tight strides of the cases above
    case PixelFormat::nv21:
        return {Plane{size, width, 8, {Component::y}}, Plane{half, width, 8, {Component::cr, Component::cb}}};
    case PixelFormat::y422p: {
        const Vector2 c{size.x / 2, size.y};
        return {Plane{size, width, 8, {Component::y}}, Plane{c, width / 2, 8, {Component::cb}}, Plane{c, width / 2, 8, {Component::cr}}};
    }
    case PixelFormat::y444p:
        return {Plane{size, width, 8, {Component::y}}, Plane{size, width, 8, {Component::cb}}, Plane{size, width, 8, {Component::cr}}};
    default:
        throw ComputeError(ErrorCode::badInputData, "Invalid pixel format");
    }
}

PictureSample createPictureSample(Vector2 size, PixelFormat format, const std::string& assetId,
                                  const std::string& workspaceId, ComputeContext* pinnedFrom, CUstream writer) {  // :254-273
    if (!(size.x > 0 && size.y > 0)) throw ComputeError(ErrorCode::invalidOperation, "createPictureSample: empty size");
    PictureSample s;
    s.imgBuffer.planes = planesForFormat(format, size);
    size_t total = 0;
    for (const Plane& p : s.imgBuffer.planes) total += (size_t)p.stride * (size_t)(int)p.size.y;
    std::shared_ptr<uint8_t> base;
    if (pinnedFrom && pinnedFrom->ctx) {
        CtxGuard g(pinnedFrom->ctx);
        auto ic = pinnedFrom->ctx;
        void* p = ic->allocHost(std::max<size_t>(total, 1), writer);
        const size_t n = std::max<size_t>(total, 1);
        base = std::shared_ptr<uint8_t>((uint8_t*)p, [ic, n](uint8_t* q) { ic->releaseHost(q, n); });
    } else {
        void* p = nullptr;
        if (posix_memalign(&p, 4096, std::max<size_t>(total, 1)) != 0) throw ComputeError(ErrorCode::outOfMemory, "host allocation failed");
        base = std::shared_ptr<uint8_t>((uint8_t*)p, [](uint8_t* q) { free(q); });
    }
    size_t off = 0;
    for (const Plane& p : s.imgBuffer.planes) {
        size_t len = (size_t)p.stride * (size_t)(int)p.size.y;
        s.imgBuffer.buffers.push_back(HostData{base, base.get() + off, len});
        off += len;
    }
    s.imgBuffer.pixelFormat = format;
    s.imgBuffer.bufferType = BufferType::cpu;
    s.imgBuffer.size = size;
    s.idAsset = assetId;
    s.idWorkspace = workspaceId;
    s.idRevision = assetId;  // :214
    return s;
}

PictureSample pictureSampleFromPlanes(PixelFormat format, Vector2 size, const uint8_t* const* planes, const int* strides, int planeCount,
                                      const std::string& assetId, const std::string& workspaceId, ComputeContext* pinnedFrom) {
    if (!(size.x > 0 && size.y > 0)) throw ComputeError(ErrorCode::invalidOperation, "empty size");
    std::vector<Plane> layout = planesForFormat(format, size);
    if (planeCount != (int)layout.size()) throw ComputeError(ErrorCode::badInputData, "Input image must have the same number of buffers as planes");
    for (int i = 0; i < planeCount; ++i) {
        const int minStride = (int)layout[i].size.x * (int)layout[i].components.size() * (layout[i].bitDepth > 8 ? 2 : 1);
        if (!planes[i] || strides[i] < minStride) throw ComputeError(ErrorCode::badInputData, "plane stride smaller than its row");
        layout[i].stride = strides[i];
    }
    // one allocation, planes back to back at their own strides (each plane start 64-byte aligned)
    size_t total = 0;
    std::vector<size_t> offs;
    for (const Plane& p : layout) {
        total = (total + 63) & ~(size_t)63;
        offs.push_back(total);
        total += (size_t)p.stride * (size_t)(int)p.size.y;
    }
    PictureSample proto = createPictureSample(Vector2{(float)((total + 3) / 4 + 1), 1.f}, PixelFormat::RGBA, assetId, workspaceId, pinnedFrom);
    std::shared_ptr<uint8_t> base = proto.imgBuffer.buffers[0].base;  // borrow the allocator (pinned or not)
    PictureSample s;
    s.imgBuffer.pixelFormat = format;
    s.imgBuffer.bufferType = BufferType::cpu;
    s.imgBuffer.size = size;
    s.imgBuffer.planes = layout;
    for (int i = 0; i < planeCount; ++i) {
        const size_t len = (size_t)layout[i].stride * (size_t)(int)layout[i].size.y;
        std::memcpy(base.get() + offs[i], planes[i], len);
        s.imgBuffer.buffers.push_back(HostData{base, base.get() + offs[i], len});
    }
    s.idAsset = assetId;
    s.idWorkspace = workspaceId;
    s.idRevision = assetId;
    return s;
}

std::vector<std::shared_ptr<ComputeBuffer>> allocPictureTextures(const ComputeContext& ctx, const std::vector<Plane>& planes, int maxPlanes) {
    const int n = std::min((int)planes.size(), maxPlanes);
    std::vector<size_t> off((size_t)n + 1, 0);
    bool together = n > 1;
    for (int i = 0; i < n; ++i) {
        off[(size_t)i + 1] = off[(size_t)i] + (size_t)(int)planes[(size_t)i].size.y * (size_t)planes[(size_t)i].stride;
        together = together && off[(size_t)i] % 256 == 0;
    }
    std::vector<std::shared_ptr<ComputeBuffer>> out;
    if (together) {
        auto whole = createBuffer(ctx, off[(size_t)n]);
        for (int i = 0; i < n; ++i) out.push_back(std::make_shared<ComputeBuffer>(whole, off[(size_t)i], off[(size_t)i + 1] - off[(size_t)i]));
    } else {
        for (int i = 0; i < n; ++i) out.push_back(createBuffer(ctx, off[(size_t)i + 1] - off[(size_t)i]));
    }
    return out;
}

// createTexture (compute.cuda.swift:413-431): device memory of stride*height per plane.
static std::vector<std::shared_ptr<ComputeBuffer>> createTexture(const ComputeContext& ctx, const ImageBuffer& image, int maxPlanes) {
    if (image.bufferType != BufferType::cpu) return image.computeTextures;
    const int planeCount = (int)image.planes.size();
    if (!(planeCount <= 3 && planeCount > 0)) throw ComputeError(ErrorCode::badInputData, "Input image must have 1, 2, or 3 planes");
    if (planeCount != (int)image.buffers.size())
        throw ComputeError(ErrorCode::badInputData, "Input image must have the same number of buffers as planes");
    return allocPictureTextures(ctx, image.planes, maxPlanes);
}

// the planes of `textures` are one allocation and `buffers` lie back to back in host memory at the same offsets: one copy moves the picture
static bool oneCopy(const std::vector<std::shared_ptr<ComputeBuffer>>& textures, const std::vector<HostData>& buffers) {
    if (textures.size() < 2 || buffers.size() < textures.size() || !textures[0]->whole) return false;
    for (size_t i = 0; i < textures.size(); ++i) {
        if (textures[i]->whole != textures[0]->whole) return false;
        if (buffers[i].size < textures[i]->size) return false;
        if ((size_t)(buffers[i].ptr - buffers[0].ptr) != (size_t)(textures[i]->mem - textures[0]->mem)) return false;
    }
    return textures[0]->mem == textures[0]->whole->mem;
}

PictureSample uploadComputePicture(const ComputeContext& ctx, const PictureSample& pict, int maxPlanes, bool retainCpuBuffer, bool wait) {  // :359-381
    if (pict.bufferType() != BufferType::cpu) return pict;
    if (pict.imgBuffer.planes.empty()) throw ComputeError(ErrorCode::badInputData, "Missing image buffer");
    if (!ctx.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
    auto textures = createTexture(ctx, pict.imgBuffer, maxPlanes);
    if (oneCopy(textures, pict.imgBuffer.buffers)) {
        CtxGuard g(ctx.ctx);
        const auto& last = textures.back();
        check(drv().cuMemcpyHtoDAsync(textures[0]->mem, pict.imgBuffer.buffers[0].ptr, (size_t)(last->mem - textures[0]->mem) + last->size, ctx.ctx->upload), "cuMemcpyHtoDAsync");
        auto ready = std::make_shared<Event>(ctx.ctx);
        check(drv().cuEventRecord(ready->e, ctx.ctx->upload), "cuEventRecord");
        for (size_t i = 0; i < textures.size(); ++i) textures[i]->ready = ready, textures[i]->noteWrite(ctx.ctx->upload), textures[i]->hostKeep = pict.imgBuffer.buffers[i].base;
    } else {
        for (size_t i = 0; i < textures.size(); ++i) {
            uploadComputeBuffer(ctx, pict.imgBuffer.buffers[i].ptr, std::min(pict.imgBuffer.buffers[i].size, textures[i]->size), textures[i]);
            textures[i]->hostKeep = pict.imgBuffer.buffers[i].base;
        }
    }
    PictureSample out = pict;
    out.imgBuffer.computeTextures = textures;
    if (!retainCpuBuffer) out.imgBuffer.buffers.clear();
    out.imgBuffer.bufferType = BufferType::gpu;
    out.done = nullptr;
    {
        CtxGuard g(ctx.ctx);
        if (wait) {
            check(drv().cuStreamSynchronize(ctx.ctx->upload), "cuStreamSynchronize");  // upstream: synchronous copies + endComputePass(ctx, true)
        } else {  // the caller learns from `done` when the source bytes may be reused
            out.done = std::make_shared<Event>(ctx.ctx);
            check(drv().cuEventRecord(out.done->e, ctx.ctx->upload), "cuEventRecord");
        }
    }
    return out;
}

namespace {
// the tight back-to-back layout of a CPU picture's planes: offsets, total; false when the planes do not lie that way in host memory
// or would not be 256-byte aligned on the device
bool tightLayout(const PictureSample& p, int maxPlanes, std::vector<size_t>& off) {
    const int n = std::min((int)p.imgBuffer.planes.size(), maxPlanes);
    if (n < 1 || (int)p.imgBuffer.buffers.size() < n) return false;
    off.assign((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i) {
        const size_t bytes = (size_t)(int)p.imgBuffer.planes[(size_t)i].size.y * (size_t)p.imgBuffer.planes[(size_t)i].stride;
        if (off[(size_t)i] % 256 || p.imgBuffer.buffers[(size_t)i].size < bytes || p.imgBuffer.buffers[(size_t)i].ptr != p.imgBuffer.buffers[0].ptr + off[(size_t)i]) return false;
        off[(size_t)i + 1] = off[(size_t)i] + bytes;
    }
    return true;
}
}  // namespace

std::vector<PictureSample> uploadComputePictures(const ComputeContext& ctx, const std::vector<const PictureSample*>& picts, int maxPlanes, bool retainCpuBuffer,
                                                 bool wait) {
    if (!ctx.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
    const size_t n = picts.size();
    std::vector<PictureSample> out(n);
    constexpr size_t kMaxRun = (size_t)96 << 20;  // bytes per copy: large enough to amortise, small enough that the first layers land early
    bool any = false;
    size_t i = 0;
    std::vector<size_t> off, offj;
    while (i < n) {
        const PictureSample& p = *picts[i];
        if (p.bufferType() != BufferType::cpu) {
            out[i] = p;
            ++i;
            continue;
        }
        any = true;
        size_t j = i;
        std::vector<std::vector<size_t>> offs;
        std::vector<size_t> at;  // byte offset of each picture of the run from the first one's first byte
        if (tightLayout(p, maxPlanes, off)) {
            offs.push_back(off), at.push_back(0);
            const uint8_t* base = p.imgBuffer.buffers[0].ptr;
            size_t end = off.back();
            for (j = i + 1; j < n; ++j) {
                const PictureSample& q = *picts[j];
                if (q.bufferType() != BufferType::cpu || !tightLayout(q, maxPlanes, offj)) break;
                const size_t next = (end + 4095) & ~(size_t)4095;  // the page-locked pool hands out multiples of 4 KB
                if (q.imgBuffer.buffers[0].ptr != base + next || next + offj.back() > kMaxRun) break;
                offs.push_back(offj), at.push_back(next);
                end = next + offj.back();
            }
            if (j - i < 2) j = i;  // a run of one: the ordinary path
            else {
                CtxGuard g(ctx.ctx);
                auto whole = createBuffer(ctx, end);
                check(drv().cuMemcpyHtoDAsync(whole->mem, base, end, ctx.ctx->upload), "cuMemcpyHtoDAsync");
                auto ready = std::make_shared<Event>(ctx.ctx);
                check(drv().cuEventRecord(ready->e, ctx.ctx->upload), "cuEventRecord");
                for (size_t k = i; k < j; ++k) {
                    const PictureSample& q = *picts[k];
                    PictureSample& o = out[k];
                    o = q;
                    o.imgBuffer.computeTextures.clear();
                    const std::vector<size_t>& ok = offs[k - i];
                    for (size_t pl = 0; pl + 1 < ok.size(); ++pl) {
                        auto t = std::make_shared<ComputeBuffer>(whole, at[k - i] + ok[pl], ok[pl + 1] - ok[pl]);
                        t->ready = ready, t->noteWrite(ctx.ctx->upload), t->hostKeep = q.imgBuffer.buffers[pl].base;
                        o.imgBuffer.computeTextures.push_back(t);
                    }
                    if (!retainCpuBuffer) o.imgBuffer.buffers.clear();
                    o.imgBuffer.bufferType = BufferType::gpu;
                    o.done = nullptr;
                }
            }
        }
        if (j == i) {
            out[i] = uploadComputePicture(ctx, p, maxPlanes, retainCpuBuffer, false);
            out[i].done = nullptr;
            j = i + 1;
        }
        i = j;
    }
    if (any) {
        CtxGuard g(ctx.ctx);
        if (wait) {
            check(drv().cuStreamSynchronize(ctx.ctx->upload), "cuStreamSynchronize");
        } else {  // one completion for the whole batch: the last copy queued
            auto done = std::make_shared<Event>(ctx.ctx);
            check(drv().cuEventRecord(done->e, ctx.ctx->upload), "cuEventRecord");
            for (size_t k = 0; k < n; ++k)
                if (picts[k]->bufferType() == BufferType::cpu) out[k].done = done;
        }
    }
    return out;
}

void waitReady(CUstream s, ComputeBuffer& t) {
    if (!t.ready || t.readyStream == s || t.readyFired.load(std::memory_order_relaxed)) return;
    if (cu().cuEventQuery(t.ready->e) == CUDA_SUCCESS) {
        t.readyFired.store(true, std::memory_order_relaxed);
        return;
    }
    check(drv().cuStreamWaitEvent(s, t.ready->e, 0), "cuStreamWaitEvent");
}

void waitPicture(const PictureSample& pict) {
    if (!pict.done) return;
    CtxGuard g(pict.done->ctx);
    check(drv().cuEventSynchronize(pict.done->e), "cuEventSynchronize");
}

CUevent pictureReadyEvent(const PictureSample& pict) {
    if (pict.done) return pict.done->e;
    // the planes of one sample are written in plane order by one stream: the last plane's event covers them all
    for (auto t = pict.imgBuffer.computeTextures.rbegin(); t != pict.imgBuffer.computeTextures.rend(); ++t)
        if ((*t)->ready) return (*t)->ready->e;
    return nullptr;
}

void pictureConsumedOn(const PictureSample& pict, CUstream consumer) {
    if (pict.bufferType() != BufferType::gpu) throw ComputeError(ErrorCode::badInputData, "not a GPU sample");
    for (const auto& t : pict.imgBuffer.computeTextures) {
        CtxGuard g(t->ctx);
        auto e = std::make_shared<Event>(t->ctx);
        check(drv().cuEventRecord(e->e, consumer), "cuEventRecord");
        t->consumerRead = e;
    }
}

PictureSample gatherComputePicture(const ComputeContext& dst, const PictureSample& pict, bool wait) {
    if (pict.bufferType() != BufferType::gpu) throw ComputeError(ErrorCode::badInputData, "not a GPU sample");
    if (!dst.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
    if (pict.imgBuffer.computeTextures.empty()) throw ComputeError(ErrorCode::badInputData, "Missing image buffer");
    const auto& src = pict.imgBuffer.computeTextures[0]->ctx;
    if (src->device == dst.ctx->device) return pict;
    CtxGuard g(dst.ctx);
    {   // direct access when the two devices are peers (NVLink / NVSwitch); otherwise the driver stages the copy
        int can = 0;
        if (drv().cuDeviceCanAccessPeer(&can, dst.ctx->device, src->device) == CUDA_SUCCESS && can) {
            CUresult r = drv().cuCtxEnablePeerAccess(src->ctx, 0);
            if (r != CUDA_SUCCESS && r != CUDA_ERROR_PEER_ACCESS_ALREADY_ENABLED) check(r, "cuCtxEnablePeerAccess");
        }
    }
    PictureSample out = pict;
    out.imgBuffer.computeTextures.clear();
    out.imgBuffer.buffers.clear();
    CUstream st = dst.ctx->upload;
    if (pict.done) check(drv().cuStreamWaitEvent(st, pict.done->e, 0), "cuStreamWaitEvent");
    for (const auto& t : pict.imgBuffer.computeTextures) {
        if (t->ready) check(drv().cuStreamWaitEvent(st, t->ready->e, 0), "cuStreamWaitEvent");  // (another context's event: always waited for)
        auto copy = createBuffer(dst, t->size);
        check(drv().cuMemcpyPeerAsync(copy->mem, dst.ctx->ctx, t->mem, src->ctx, t->size, st), "cuMemcpyPeerAsync");
        auto e = std::make_shared<Event>(dst.ctx);
        check(drv().cuEventRecord(e->e, st), "cuEventRecord");
        t->consumerRead = e;  // the source's next writer (its mixer's ring, its pool) waits for this copy
        copy->ready = e;      // and readers on dst's other streams order themselves behind it
        copy->noteWrite(st);
        out.imgBuffer.computeTextures.push_back(copy);
    }
    out.done = nullptr;
    if (wait) {
        check(drv().cuStreamSynchronize(st), "cuStreamSynchronize");
    } else {
        out.done = std::make_shared<Event>(dst.ctx);
        check(drv().cuEventRecord(out.done->e, st), "cuEventRecord");
    }
    return out;
}

PictureSample downloadComputePicture(const ComputeContext& ctx, const PictureSample& pict, bool retainGpuBuffer, bool wait) {  // :383-402
    if (pict.bufferType() != BufferType::gpu) return pict;
    if (pict.imgBuffer.planes.empty()) throw ComputeError(ErrorCode::badInputData, "Missing image buffer");
    if (!ctx.ctx) throw ComputeError(ErrorCode::badContextState, "No context");
    PictureSample out = pict;
    const size_t n = pict.imgBuffer.computeTextures.size();
    CtxGuard g(ctx.ctx);
    {
        // Always fresh host buffers: upstream's `dst` is a value-type Data (copy on write), so a download never changes the bytes of the
        // sample that was uploaded, nor of its copies; the buffers a GPU sample retained from its upload alias that allocation here.
        // Page-locked and pooled: a copy into pageable memory is staged by the driver at a fraction of the link rate.
        ComputeContext pin = ctx;
        PictureSample host = createPictureSample(pict.size(), pict.pixelFormat(), pict.idAsset, pict.idWorkspace, &pin, ctx.ctx->download);
        out.imgBuffer.buffers = host.imgBuffer.buffers;
        for (size_t i = 0; i < n && i < out.imgBuffer.buffers.size(); ++i)  // (decoder-style padded planes: the device plane may be larger than the tight layout)
            if (out.imgBuffer.buffers[i].size < pict.imgBuffer.computeTextures[i]->size) {
                const size_t len = pict.imgBuffer.computeTextures[i]->size;
                void* q = ctx.ctx->allocHost(len, ctx.ctx->download);
                auto ic = ctx.ctx;
                std::shared_ptr<uint8_t> base((uint8_t*)q, [ic, len](uint8_t* z) { ic->releaseHost(z, len); });
                out.imgBuffer.buffers[i] = HostData{base, base.get(), len};
            }
    }
    if (pict.done) check(drv().cuStreamWaitEvent(ctx.ctx->download, pict.done->e, 0), "cuStreamWaitEvent");
    if (oneCopy(pict.imgBuffer.computeTextures, out.imgBuffer.buffers)) {
        const auto& tx = pict.imgBuffer.computeTextures;
        for (const auto& tex : tx)
            waitReady(ctx.ctx->download, *tex);
        check(drv().cuMemcpyDtoHAsync(out.imgBuffer.buffers[0].ptr, tx[0]->mem, (size_t)(tx.back()->mem - tx[0]->mem) + tx.back()->size, ctx.ctx->download), "cuMemcpyDtoHAsync");
        auto read = std::make_shared<Event>(ctx.ctx);
        check(drv().cuEventRecord(read->e, ctx.ctx->download), "cuEventRecord");
        for (const auto& tex : tx) tex->usedByDownload = true, tex->lastRead = read;
    } else
    for (size_t i = 0; i < n; ++i) {
        const auto& tex = pict.imgBuffer.computeTextures[i];
        waitReady(ctx.ctx->download, *tex);
        downloadComputeBuffer(ctx, *tex, out.imgBuffer.buffers[i].ptr, out.imgBuffer.buffers[i].size);
        // whoever overwrites this plane next (the mixer recycles its backing ring) waits for this copy first
        if (!tex->lastRead) tex->lastRead = std::make_shared<Event>(ctx.ctx);
        check(drv().cuEventRecord(tex->lastRead->e, ctx.ctx->download), "cuEventRecord");
    }
    out.done = nullptr;
    if (wait) {
        check(drv().cuStreamSynchronize(ctx.ctx->download), "cuStreamSynchronize");  // endComputePass(ctx, true), :396
    } else {
        out.done = std::make_shared<Event>(ctx.ctx);
        check(drv().cuEventRecord(out.done->e, ctx.ctx->download), "cuEventRecord");
    }
    if (!retainGpuBuffer) out.imgBuffer.computeTextures.clear();
    out.imgBuffer.bufferType = BufferType::cpu;
    return out;
}

BarrierResult GPUBarrierUpload::operator()(const PictureSample& sample, bool wait) const {  // compute.swift:183-195
    BarrierResult r;
    if (sample.bufferType() != BufferType::cpu) {
        r.sample = sample;  // .just($0)
        return r;
    }
    try {
        r.sample = uploadComputePicture(context, sample, 3, retainCpuBuffer, wait);
    } catch (const std::exception& e) {
        r.ok = false;
        r.error = EventError{"barrier.upload", -1, e.what(), sample.assetId()};
    }
    return r;
}
BarrierResult GPUBarrierDownload::operator()(const PictureSample& sample, bool wait) const {  // compute.swift:240-252
    BarrierResult r;
    if (sample.bufferType() != BufferType::gpu) {
        r.sample = sample;
        return r;
    }
    try {
        r.sample = downloadComputePicture(context, sample, retainGpuBuffer, wait);
    } catch (const std::exception& e) {
        r.ok = false;
        r.error = EventError{"barrier.download", -1, e.what(), sample.assetId()};
    }
    return r;
}

// ---- kernels --------------------------------------------------------------------------------------------

static unsigned gcdu(unsigned a, unsigned b) {  // clock.swift:197, used for block sizing at :290-291
    while (b) {
        unsigned t = a % b;
        a = b;
        b = t;
    }
    return a;
}

ComputeContext runComputeKernel(const ComputeContext& ctxIn, const std::vector<const PictureSample*>& images,
                                const PictureSample& target, ComputeKernel kernel, const std::string& customName,
                                int maxPlanes, const void* uniforms, size_t uniformsSize, bool /*blends*/) {  // :260-306
    (void)maxPlanes;
    for (const PictureSample* im : images)
        if (im->bufferType() != BufferType::gpu) throw ComputeError(ErrorCode::badInputData, "Input images must be uploaded to GPU");
    if (target.imgBuffer.planes.empty() || target.imgBuffer.computeTextures.empty()) throw ComputeError(ErrorCode::badTarget, "badTarget");
    if (kernel != ComputeKernel::custom)  // the built-in kernels derive the output pitch from the launch size, as upstream's do
        for (const Plane& p : target.imgBuffer.planes)
            // (the packed 4:2:2 formats list the four components of a two-pixel group, sample.pict.linux.swift:283-285: two bytes per pixel)
            if (p.stride != (int)p.size.x * (target.pixelFormat() == PixelFormat::yuvs || target.pixelFormat() == PixelFormat::zvuy ? 2 : (int)p.components.size()))
                throw ComputeError(ErrorCode::badTarget, "badTarget: the per-layer kernels need a target without row padding");
    ComputeContext ctx = maybeBuildKernel(ctxIn, kernel, customName);
    const std::string name = kernel == ComputeKernel::custom ? customName : computeKernelName(kernel);
    auto it = ctx.library.find(name);
    if (it == ctx.library.end()) throw ComputeError(ErrorCode::computeKernelNotFound, "computeKernelNotFound(" + name + ")");
    const CUDAProgram& prog = *it->second;

    CtxGuard g(ctx.ctx);
    const CuDriver& d = drv();
    std::vector<std::shared_ptr<ComputeBuffer>> keep;  // buffers in marshalling order (:294-297)
    for (const auto& t : target.imgBuffer.computeTextures) keep.push_back(t);
    std::vector<int32_t> inputStride;
    for (const PictureSample* im : images) {
        for (const auto& t : im->imgBuffer.computeTextures) {
            waitReady(ctx.ctx->compute, *t);
            t->lastUse = nullptr, t->lastUseUnknown = true;  // read by this launch: the pool falls back to the stream tails
            keep.push_back(t);
        }
        for (const Plane& p : im->imgBuffer.planes) inputStride.push_back((int32_t)p.stride);
    }
    for (const auto& t : target.imgBuffer.computeTextures) waitReady(ctx.ctx->compute, *t);
    // uniforms and the stride array travel as two small device buffers, as upstream (:278-289); both copies are
    // ordered on the compute stream ahead of the launch.
    if (uniforms && uniformsSize) {
        auto ub = createBuffer(ctx, uniformsSize);
        check(d.cuMemcpyHtoDAsync(ub->mem, uniforms, uniformsSize, ctx.ctx->compute), "cuMemcpyHtoDAsync");
        keep.push_back(ub);
    }
    if (!inputStride.empty()) {
        auto sb = createBuffer(ctx, inputStride.size() * sizeof(int32_t));
        check(d.cuMemcpyHtoDAsync(sb->mem, inputStride.data(), inputStride.size() * sizeof(int32_t), ctx.ctx->compute), "cuMemcpyHtoDAsync");
        keep.push_back(sb);
    }
    // pageable sources: the async copies above have completed their host read on return (staged by the driver)
    const unsigned W = (unsigned)target.size().x, H = (unsigned)target.size().y;
    const unsigned bx = gcdu(W, 16), by = gcdu(H, 16);
    std::vector<void*> args;
    for (auto& b : keep) args.push_back(&b->mem);
    check(d.cuLaunchKernel(prog.function, W / bx, H / by, 1, bx, by, 1, 0, ctx.ctx->compute, args.data(), nullptr), "cuLaunchKernel");
    noteKernelLaunch();
    markWritten(ctx, target);
    // `keep` may drop its temporaries now: release() orders every stream behind the launch before reuse.
    return ctx;
}

void markWritten(const ComputeContext& ctx, const PictureSample& target) {  // caller holds a CtxGuard
    for (const auto& t : target.imgBuffer.computeTextures) {
        if (!t->ready) t->ready = std::make_shared<Event>(ctx.ctx);
        check(drv().cuEventRecord(t->ready->e, ctx.ctx->compute), "cuEventRecord");
        t->noteWrite(ctx.ctx->compute);
    }
}

ImageUniforms makeImageUniforms(const PictureSample& image, const PictureSample& target) {  // compute.swift:149-161
    ImageUniforms u;
    std::memcpy(u.transform, image.matrix().inverse().transpose().data(), 64);
    std::memcpy(u.textureTransform, image.textureMatrix().inverse().transpose().data(), 64);
    std::memcpy(u.borderMatrix, image.borderMatrix().inverse().transpose().data(), 64);
    const Vector4 f = image.fillColor();
    u.fillColor[0] = f.x, u.fillColor[1] = f.y, u.fillColor[2] = f.z, u.fillColor[3] = f.w;
    u.inputSize[0] = image.size().x, u.inputSize[1] = image.size().y;
    u.outputSize[0] = target.size().x, u.outputSize[1] = target.size().y;
    u.opacity = image.opacity();
    u.imageTime = image.timescale ? (float)((double)image.timeValue / (double)image.timescale) : 0.f;
    u.targetTime = target.timescale ? (float)((double)target.timeValue / (double)target.timescale) : 0.f;
    return u;
}

ComputeContext applyComputeImage(const ComputeContext& ctx, const PictureSample& image, const PictureSample& target,
                                 ComputeKernel kernel) {  // compute.swift:145-170
    const ImageUniforms u = makeImageUniforms(image, target);
    return runComputeKernel(ctx, {&image}, target, kernel, "", 3, &u, sizeof(u), true);
}

}  // namespace svb
