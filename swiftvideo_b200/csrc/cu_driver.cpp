// cu_driver.cpp -- dlopen-based binding of the CUDA driver API (see cu_driver.h).
#include "cu_driver.h"

#include <dlfcn.h>

#include <mutex>

namespace svb {

#define SVB_STR2(x) #x
#define SVB_STR(x) SVB_STR2(x)  // expands cuda.h's versioning macros first: cuMemAlloc -> "cuMemAlloc_v2"

static CuDriver g_drv;
static std::once_flag g_once;

static void load() {
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        g_drv.why = "libcuda.so.1 not found (no NVIDIA driver on this machine)";
        return;
    }
    bool all = true;
#define SVB_CU_FN(name)                                                  \
    g_drv.name = reinterpret_cast<decltype(g_drv.name)>(dlsym(h, SVB_STR(name))); \
    if (!g_drv.name) {                                                   \
        all = false;                                                     \
        g_drv.why = "libcuda.so.1 lacks " SVB_STR(name);                 \
    }
    SVB_CU_FN(cuInit)
    SVB_CU_FN(cuGetErrorString)
    SVB_CU_FN(cuDeviceGetCount)
    SVB_CU_FN(cuDeviceGet)
    SVB_CU_FN(cuDeviceGetAttribute)
    SVB_CU_FN(cuDeviceGetName)
    SVB_CU_FN(cuDevicePrimaryCtxRetain)
    SVB_CU_FN(cuDevicePrimaryCtxRelease)
    SVB_CU_FN(cuCtxPushCurrent)
    SVB_CU_FN(cuCtxPopCurrent)
    SVB_CU_FN(cuCtxSynchronize)
    SVB_CU_FN(cuMemAlloc)
    SVB_CU_FN(cuMemFree)
    SVB_CU_FN(cuMemHostAlloc)
    SVB_CU_FN(cuMemFreeHost)
    SVB_CU_FN(cuMemcpyHtoD)
    SVB_CU_FN(cuMemcpyDtoH)
    SVB_CU_FN(cuMemcpyHtoDAsync)
    SVB_CU_FN(cuMemcpyDtoHAsync)
    SVB_CU_FN(cuMemcpy2DAsync)
    SVB_CU_FN(cuMemsetD8Async)
    SVB_CU_FN(cuMemcpyPeerAsync)
    SVB_CU_FN(cuDeviceCanAccessPeer)
    SVB_CU_FN(cuCtxEnablePeerAccess)
    SVB_CU_FN(cuModuleLoadData)
    SVB_CU_FN(cuModuleUnload)
    SVB_CU_FN(cuModuleGetFunction)
    SVB_CU_FN(cuFuncSetAttribute)
    SVB_CU_FN(cuFuncGetAttribute)
    SVB_CU_FN(cuOccupancyMaxActiveBlocksPerMultiprocessor)
    SVB_CU_FN(cuTexObjectCreate)
    SVB_CU_FN(cuTexObjectDestroy)
    SVB_CU_FN(cuLaunchKernel)
    SVB_CU_FN(cuStreamCreate)
    SVB_CU_FN(cuStreamDestroy)
    SVB_CU_FN(cuStreamSynchronize)
    SVB_CU_FN(cuStreamWaitEvent)
    SVB_CU_FN(cuEventCreate)
    SVB_CU_FN(cuEventDestroy)
    SVB_CU_FN(cuEventRecord)
    SVB_CU_FN(cuEventSynchronize)
    SVB_CU_FN(cuEventQuery)
    SVB_CU_FN(cuEventElapsedTime)
    SVB_CU_FN(cuTensorMapEncodeTiled)
#undef SVB_CU_FN
    if (!all) return;
    CUresult r = g_drv.cuInit(0);  // compute.cuda.swift:94-100
    if (r != CUDA_SUCCESS) {
        g_drv.why = "cuInit failed (no usable GPU)";
        return;
    }
    g_drv.ok = true;
    g_drv.why = "";
}

const CuDriver& cu() {
    std::call_once(g_once, load);
    return g_drv;
}

}  // namespace svb
