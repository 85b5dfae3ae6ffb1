import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, swiftvideo_b200 as sv
ctx = sv.make_compute_context(0)
h = sv.create_picture_sample(1920, 1080, sv.NV12, "a", "b", pinned_from=ctx)
h.set_host_bytes(np.zeros(1920*1080*3//2, np.uint8))
d = h.upload(ctx)
ctx.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    outs = [d.scale_convert(ctx, 1280, 720, sv.BGRA, 0, wait=False) for _ in range(8)]
    t1 = time.perf_counter()
    ctx.synchronize()
    t2 = time.perf_counter()
    print("queue 8 launches %.3f ms, drain %.3f ms" % ((t1-t0)*1e3, (t2-t1)*1e3))
    del outs

# the bench's loop shape: outputs of the previous step are released while this step's launches are queued
srcs = []
for i in range(8):
    hh = sv.create_picture_sample(1920, 1080, sv.NV12, f"s{i}", "b", pinned_from=ctx)
    hh.set_host_bytes(np.full(1920*1080*3//2, i, np.uint8))
    srcs.append(hh.upload(ctx))
ctx.synchronize()
for variant in ("keep-last", "drop-at-once"):
    for rep in range(2):
        t0 = time.perf_counter()
        last = None
        for s in range(20):
            o = [srcs[k].scale_convert(ctx, 1280, 720, sv.BGRA, 0, wait=False) for k in range(8)]
            if variant == "keep-last":
                last = o
            del o
        t1 = time.perf_counter()
        ctx.synchronize()
        t2 = time.perf_counter()
        print(variant, "20 steps: queue %.3f ms, drain %.3f ms" % ((t1-t0)*1e3, (t2-t1)*1e3))
t = sv.Timer(ctx)
t.start()
o = [srcs[k].scale_convert(ctx, 1280, 720, sv.BGRA, 0, wait=False) for k in range(8)]
t.stop()
print("timer 8 launches: %.3f ms" % t.elapsed_ms())
