"""How much of the host->device link does one tick's upload pattern reach?  One step of the headline workload uploads 8 x (4K NV12) + 56 x
(1080p NV12) pictures, one cuMemcpyHtoDAsync per plane: 128 copies of 1-8 MB.  This probe times that pattern from page-locked memory on
1, 2, 3 and 4 streams, plane by plane and picture by picture, against one copy of the same bytes.  Raw driver API, no library code.
    python tools/copy_probe.py   ->  JSON lines (GB/s), not bench values"""
import ctypes as C
import json
import time

cu = C.CDLL("libcuda.so.1")


def ck(rc, what):
    assert rc == 0, f"{what}: {rc}"


ck(cu.cuInit(0), "cuInit")
dev = C.c_int()
ck(cu.cuDeviceGet(C.byref(dev), 0), "cuDeviceGet")
ctx = C.c_void_p()
ck(cu.cuDevicePrimaryCtxRetain(C.byref(ctx), dev), "retain")
ck(cu.cuCtxPushCurrent_v2(ctx), "push")
planes = []
for _ in range(8):
    planes += [3840 * 2160, 3840 * 1080]
    for _ in range(7):
        planes += [1920 * 1080, 1920 * 540]
total = sum(planes)
host = C.c_void_p()
ck(cu.cuMemHostAlloc(C.byref(host), C.c_size_t(total), 0), "cuMemHostAlloc")
devp = C.c_uint64()
ck(cu.cuMemAlloc_v2(C.byref(devp), C.c_size_t(total)), "cuMemAlloc")
streams = []
for _ in range(4):
    s = C.c_void_p()
    ck(cu.cuStreamCreate(C.byref(s), 1), "cuStreamCreate")
    streams.append(s)


def run(sizes, ns, reps=6):
    best = None
    for _ in range(reps):
        for s in streams:
            cu.cuStreamSynchronize(s)
        t0 = time.perf_counter()
        for _step in range(4):
            off = 0
            for k, n in enumerate(sizes):
                ck(cu.cuMemcpyHtoDAsync_v2(C.c_uint64(devp.value + off), C.c_void_p(host.value + off), C.c_size_t(n), streams[k % ns]), "copy")
                off += n
        for s in streams:
            cu.cuStreamSynchronize(s)
        dt = (time.perf_counter() - t0) / 4
        best = dt if best is None else min(best, dt)
    return best


pictures = [planes[i] + planes[i + 1] for i in range(0, len(planes), 2)]
for name, sizes in (("one copy", [total]), ("per picture (64 copies)", pictures), ("per plane (128 copies)", planes)):
    for ns in (1, 2, 3, 4):
        if name == "one copy" and ns > 1:
            continue
        dt = run(sizes, ns)
        print(json.dumps({"pattern": name, "streams": ns, "ms_per_step": round(dt * 1e3, 3), "gbs": round(total / dt / 1e9, 1), "frames_per_s_at_this_rate": round(8 / dt, 1)}))

# the same upload pattern with the step's downloads (8 composited 4K NV12 frames, one copy per plane) running the other way at the same time
down = [3840 * 2160, 3840 * 1080] * 8
dtotal = sum(down)
hostd = C.c_void_p()
ck(cu.cuMemHostAlloc(C.byref(hostd), C.c_size_t(dtotal), 0), "cuMemHostAlloc")
devd = C.c_uint64()
ck(cu.cuMemAlloc_v2(C.byref(devd), C.c_size_t(dtotal)), "cuMemAlloc")


def duplex(sizes, ns=1, reps=6, steps=6):
    best = None
    for _ in range(reps):
        for s in streams:
            cu.cuStreamSynchronize(s)
        t0 = time.perf_counter()
        for _step in range(steps):
            off = 0
            for k, n in enumerate(sizes):
                ck(cu.cuMemcpyHtoDAsync_v2(C.c_uint64(devp.value + off), C.c_void_p(host.value + off), C.c_size_t(n), streams[k % ns]), "h2d")
                off += n
            off = 0
            for n in down:
                ck(cu.cuMemcpyDtoHAsync_v2(C.c_void_p(hostd.value + off), C.c_uint64(devd.value + off), C.c_size_t(n), streams[3]), "d2h")
                off += n
        for s in streams:
            cu.cuStreamSynchronize(s)
        dt = (time.perf_counter() - t0) / steps
        best = dt if best is None else min(best, dt)
    return best


for name, sizes in (("one copy", [total]), ("per picture (64 copies)", pictures), ("per plane (128 copies)", planes)):
  for ns in (1, 2, 3):
    dt = duplex(sizes, ns)
    print(json.dumps({"pattern": name + " + the step's 16 download copies at the same time", "h2d_streams": ns, "ms_per_step": round(dt * 1e3, 3), "h2d_gbs": round(total / dt / 1e9, 1),
                      "d2h_gbs": round(dtotal / dt / 1e9, 1), "frames_per_s_at_this_rate": round(8 / dt, 1)}))
