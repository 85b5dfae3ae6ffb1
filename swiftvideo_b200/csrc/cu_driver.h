// cu_driver.h -- the CUDA driver API, bound at run time.
//
// The reference reaches the GPU through its CCUDA shim, which is nothing but <cuda.h> + <nvrtc.h>
// (/root/reference/Sources/CCUDA/shim.h:1-4): the driver API *is* its FFI.  We call the same entry points
// (the list compute.cuda.swift uses: cuInit :97, cuDeviceGet* :136-141, cuCtx* :309-317, cuMemAlloc :407,
// cuMemcpyHtoD/DtoH :339,:353, cuModuleLoadData :193, cuModuleGetFunction :194, cuLaunchKernel :299) but
// resolve them with dlopen("libcuda.so.1") so that libsvb200.so itself loads on a machine without a
// driver (symbol/ABI tests run there); any compute call then fails with SVB_ERROR_DEVICE_NOT_AVAILABLE.
#pragma once
#include <cuda.h>

namespace svb {

struct CuDriver {
    bool ok = false;
    const char* why = "not loaded";
#define SVB_CU_FN(name) decltype(&::name) name = nullptr;
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
};

// Loads libcuda.so.1 once (thread-safe) and calls cuInit(0).
const CuDriver& cu();

}  // namespace svb
