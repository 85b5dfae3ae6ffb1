"""-m gpu: VideoMixer behaviour through the C ABI (mix.video.swift:21-184) and the reference's own launch path."""
import ctypes as C

import numpy as np
import pytest

import scenes
import swiftvideo_b200 as sv
from gpu_util import FMT, context, fetch, first_diff, gpu_target, to_gpu
from oracle import oracle as O
from swiftvideo_b200 import animator, api

pytestmark = pytest.mark.gpu


def _place(pict, canvas, src_size, pos, size, z, opacity=1.0, revision=None, **kw):
    m, t, b = animator.picture_state(canvas, src_size, pos, size, z=z, **kw)
    return pict.with_(matrix=m, texture_matrix=t, border_matrix=b, opacity=opacity, revision=revision)


def _oracle_mix(target_fmt, canvas, placed, images, lib=None):
    """Oracle fold using the uniforms the host side itself derives from the pictures' matrices (lib: the reference-text build unless given)."""
    tgt = sv.create_picture_sample(canvas[0], canvas[1], FMT[target_fmt], "t", "w")
    us = []
    for p in placed:
        u = api.make_image_uniforms(p, tgt)
        ou = O.Uniforms()
        C.memmove(C.byref(ou), C.byref(u), 236)
        us.append(ou)
    want = O.Image(target_fmt, canvas[0], canvas[1])
    assert (lib or O.best()[0]).mix(want, images, us) == 0
    return want.data


@pytest.mark.parametrize("mode", [sv.MixMode.FUSED, sv.MixMode.FUSED_RING, sv.MixMode.FUSED_TILED, sv.MixMode.PER_LAYER, sv.MixMode.GENERIC])
def test_mixer_z_order_and_generations(mode):
    ctx = context()
    canvas = (256, 128)
    mixer = sv.VideoMixer(ctx, canvas[0], canvas[1], sv.NV12, asset_id="mixer", workspace_id="ws")
    mixer.set_mode(mode)
    imgs = [scenes.random_image(O.NV12, 128, 72, 900 + i) for i in range(3)]
    gpu = [to_gpu(ctx, im, f"asset{i}") for i, im in enumerate(imgs)]
    # pushed in scrambled order; z decides the fold order (mix.video.swift:115)
    placed = [_place(gpu[0], canvas, (128, 72), (0, 0), (256, 128), z=0, opacity=1.0),
              _place(gpu[1], canvas, (128, 72), (30, 10), (160, 90), z=2, opacity=0.5),
              _place(gpu[2], canvas, (128, 72), (60, 30), (160, 90), z=1, opacity=0.7)]
    assert [p.z_index() for p in placed] == [1, 3, 2]
    assert mixer.push(placed[1]) and mixer.push(placed[0]) and mixer.push(placed[2])
    out = mixer.mix(1000)
    want = _oracle_mix(O.NV12, canvas, [placed[0], placed[2], placed[1]], [imgs[0], imgs[2], imgs[1]])
    got = fetch(ctx, out)
    assert (got == want).all(), first_diff(got, want)
    assert out.info().time == 1000 and out.info().buffer_type == api.BUFFER_GPU
    # nothing pushed this tick: the previous generation is still composed (samples[1], :104-107,114) ...
    got2 = fetch(ctx, mixer.mix(2000))
    assert (got2 == want).all()
    # ... and is gone one tick later: only the clear remains
    got3 = fetch(ctx, mixer.mix(3000))
    clear = O.Image(O.NV12, *canvas)
    O.best()[0].clear(clear)
    assert (got3 == clear.data).all()
    # a newer sample of the same revision replaces the older one
    mixer.push(placed[0])
    mixer.push(_place(gpu[2], canvas, (128, 72), (0, 0), (256, 128), z=0, opacity=1.0, revision="asset0"))
    got4 = fetch(ctx, mixer.mix(4000))
    want4 = _oracle_mix(O.NV12, canvas, [_place(gpu[2], canvas, (128, 72), (0, 0), (256, 128), z=0)], [imgs[2]])
    assert (got4 == want4).all()
    mixer.close()


def test_mixer_backing_ring_and_passthrough():
    ctx = context()
    mixer = sv.VideoMixer(ctx, 64, 32, sv.NV12, asset_id="ring")
    seen = []
    for t in range(12):
        out = mixer.mix(t)
        seen.append(tuple(out.device_planes()))
    assert len(set(seen[:10])) == 10            # numberBackingImages = 10 (mix.video.swift:167)
    assert seen[10] == seen[0] and seen[11] == seen[1]
    own = sv.create_picture_sample(64, 32, sv.NV12, "ring", "w")
    assert mixer.push(own) is False             # the mixer's own asset passes through (:70-72)
    mixer.close()


def test_mixer_errors():
    ctx = context()
    # a CPU sample reaches the fold: "Input images must be uploaded to GPU" (compute.cuda.swift:268-270)
    mixer = sv.VideoMixer(ctx, 64, 32, sv.NV12, asset_id="m1")
    cpu = sv.create_picture_sample(64, 32, sv.NV12, "cpu", "w")
    mixer.push(cpu)
    with pytest.raises(sv.ComputeError) as e:
        mixer.mix(0)
    assert e.value.name == "badInputData"
    # the generations still rotate after a failed compose (the defer block, :104-107): the bad sample is seen
    # once more from generation 1, then it is gone
    with pytest.raises(sv.ComputeError):
        mixer.mix(1)
    out = mixer.mix(2)
    clear = O.Image(O.NV12, 64, 32)
    O.best()[0].clear(clear)
    assert (fetch(ctx, out) == clear.data).all()
    # nv12 -> y420p has no kernel name in the map: defaultComputeKernelFromString throws invalidValue (compute.swift:105-108)
    m2 = sv.VideoMixer(ctx, 64, 32, sv.Y420P, asset_id="m2")
    m2.push(to_gpu(ctx, scenes.random_image(O.NV12, 64, 32, 1), "n"))
    with pytest.raises(sv.ComputeError) as e:
        m2.mix(0)
    assert e.value.name == "invalidValue"
    # BGRA target: clear works (img_clear_bgra); a BGRA layer goes through img_bgra_bgra (tests/test_gpu_formats.py); an NV12 layer has
    # no name in the map (img_nv12_bgra): invalidValue
    m3 = sv.VideoMixer(ctx, 64, 32, sv.BGRA, asset_id="m3")
    got = fetch(ctx, m3.mix(0))
    want = O.Image(O.BGRA, 64, 32)
    O.best()[0].clear(want)
    assert (got == want.data).all()
    m3.push(to_gpu(ctx, scenes.random_image(O.NV12, 64, 32, 2), "b"))
    with pytest.raises(sv.ComputeError) as e:
        m3.mix(1)
    assert e.value.name == "invalidValue"
    for m in (mixer, m2, m3):
        m.close()


def test_mix_many_equals_single():
    """Several mixers folded into one launch give the bytes each gives alone."""
    ctx = context()
    canvas = (384, 160)
    mixers, wants = [], []
    for s in range(5):
        m = sv.VideoMixer(ctx, canvas[0], canvas[1], sv.NV12 if s % 2 == 0 else sv.Y420P, asset_id=f"mm{s}")
        fmt = O.NV12 if s % 2 == 0 else O.Y420P
        imgs = [scenes.random_image(O.Y420P, 200, 120, 1000 + 10 * s + i) for i in range(3)]
        placed = [_place(to_gpu(ctx, im, f"s{s}l{i}"), canvas, (200, 120), (20 * i + 5 * s, 10 * i), (300 - 20 * i, 140 - 10 * i), z=i,
                         opacity=1.0 - 0.2 * i) for i, im in enumerate(imgs)]
        m.push_many(placed)
        mixers.append(m)
        wants.append(_oracle_mix(fmt, canvas, placed, imgs))
    outs = sv.VideoMixer.mix_many(mixers, 5, wait=False)
    for s, out in enumerate(outs):
        got = fetch(ctx, out)
        assert (got == wants[s]).all(), f"stream {s}: {first_diff(got, wants[s])}"
    for m in mixers:
        m.close()


def test_reference_launch_path():
    """The drop-in contract: a kernel image registered under a ComputeKernel name wins over the built-in
    (compute.cuda.swift:210-212), and runComputeKernel/applyComputeImage launch it with the reference's geometry."""
    ctx = context().sharing()              # createComputeContext(sharing:): empty kernel library
    image = sv.kernel_module_image()
    for name in ("img_clear_y420p", "img_y420p_y420p"):   # what Examples/Mixing uses on Linux
        ctx.build_compute_kernel(name, image)
    canvas = (320, 180)                    # 16x4 blocks: gcd(320,16)=16, gcd(180,16)=4
    src = scenes.random_image(O.Y420P, 160, 90, 77)
    g = to_gpu(ctx, src, "src")
    m, t, b = animator.picture_state(canvas, (160, 90), (0, 0), (160, 180), z=0, aspect="fill")
    placed = g.with_(matrix=m, texture_matrix=t, border_matrix=b, opacity=0.75)
    target = gpu_target(ctx, O.Y420P, *canvas)
    api.run_compute_kernel(ctx, [], target, api.default_compute_kernel_from_string("img_clear_y420p"))
    api.apply_compute_image(ctx, placed, target, api.default_compute_kernel_from_string("img_y420p_y420p"))
    ctx.synchronize()
    want = _oracle_mix(O.Y420P, canvas, [placed], [src])
    got = fetch(ctx, target)
    assert (got == want).all(), first_diff(got, want)


def test_async_upload_mix_download_pipeline():
    """Host buffers in, host buffers out, nothing waited on until the end (the e2e path bench.py times)."""
    ctx = context()
    canvas = (512, 256)
    mixer = sv.VideoMixer(ctx, canvas[0], canvas[1], sv.NV12, asset_id="pipe")
    frames = []
    for f in range(6):
        imgs = [scenes.random_image(O.NV12, 256, 144, 2000 + 10 * f + i) for i in range(3)]
        hosts = []
        for i, im in enumerate(imgs):
            h = sv.create_picture_sample(256, 144, sv.NV12, f"a{i}", "w", pinned_from=ctx)
            h.set_host_bytes(im.data)
            hosts.append(h)
        placed = [_place(h.upload(ctx, retain_cpu_buffer=False, wait=False), canvas, (256, 144), (40 * i, 20 * i), (400, 200), z=i, opacity=0.9 - 0.2 * i)
                  for i, h in enumerate(hosts)]
        mixer.push_many(placed)
        out = mixer.mix(f, wait=False).download(ctx, retain_gpu_buffer=True, wait=False)
        frames.append((out, placed, imgs))
    for out, placed, imgs in frames:
        out.wait()
        want = _oracle_mix(O.NV12, canvas, placed, imgs)
        got = out.host_bytes()
        assert (got == want).all(), first_diff(got, want)
    mixer.close()


def test_barriers_are_idempotent():
    """GPUBarrierUpload / GPUBarrierDownload (compute.swift:175-198, :232-255): a sample already on the right side passes through as the
    SAME sample (`.just($0)`), a CPU sample is uploaded, a GPU sample downloaded; bytes survive the round trip."""
    ctx = context()
    img = scenes.random_image(O.NV12, 128, 72, 4242)
    cpu = sv.create_picture_sample(128, 72, sv.NV12, "bar", "w")
    cpu.set_host_bytes(img.data)
    gpu, err = cpu.barrier_upload(ctx)
    assert err is None and gpu.info().buffer_type == sv.BUFFER_GPU
    again, err = gpu.barrier_upload(ctx)          # already on the GPU: untouched
    assert err is None and again.info().buffer_type == sv.BUFFER_GPU and again.same_sample(gpu)
    back, err = gpu.barrier_download(ctx)
    assert err is None and back.info().buffer_type == sv.BUFFER_CPU
    assert (back.host_bytes() == img.data).all()
    still, err = back.barrier_download(ctx)       # already on the CPU: untouched
    assert err is None and still.same_sample(back)


def test_async_upload_sets_done_and_source_is_reusable_after_wait():
    """uploadComputePicture(wait=false) returns at once with `done` set (a page-locked staging buffer that the host refills
    raced the copy): after wait() the source bytes may be overwritten, and every uploaded sample keeps the bytes it was given."""
    ctx = context()
    staging = sv.create_picture_sample(1920, 1080, sv.NV12, "stage", "w", pinned_from=ctx)
    frames = [scenes.random_image(O.NV12, 1920, 1080, 5000 + i).data for i in range(4)]
    ups = []
    for f in frames:
        staging.set_host_bytes(f)
        up = staging.upload(ctx, retain_cpu_buffer=False, wait=False)
        up.wait()                                  # the copy has left the staging buffer
        ups.append(up)
    for f, up in zip(frames, ups):
        assert (fetch(ctx, up) == f).all()


def test_download_never_touches_the_uploaded_sample():
    """downloadComputePicture allocates its own host buffers (upstream: a value-type Data): the CPU sample that was uploaded, and its
    copies, keep their bytes when the GPU sample is composed into and downloaded."""
    ctx = context()
    canvas = (256, 144)
    stale = sv.create_picture_sample(canvas[0], canvas[1], sv.NV12, "t", "w")
    stale.set_host_bytes(np.full(canvas[0] * canvas[1] * 3 // 2, 0xA5, dtype=np.uint8))
    target = stale.upload(ctx)                     # retain_cpu_buffer=True: the GPU sample still refers to stale's host bytes
    layer = to_gpu(ctx, scenes.random_image(O.NV12, 256, 144, 77), "l")
    u = scenes.layer_uniforms(canvas, (256, 144), (0, 0), canvas, z=1, opacity=1.0)
    sv.compose(ctx, target, [layer], [u], sv.MixMode.FUSED)
    out = target.download(ctx, retain_gpu_buffer=True)
    assert (out.host_bytes() != 0xA5).any()
    assert (stale.host_bytes() == 0xA5).all(), "the download wrote into the uploaded sample's host bytes"


def test_backing_ring_lapped_by_async_downloads():
    """25 ticks with mix(wait=false) + download(wait=false) and nothing waited on: the backing ring of 10 comes round twice while earlier
    downloads may still be reading (a compose is an order of magnitude faster than the copy over PCIe).  Every downloaded frame must
    hold its own tick's bytes (write-after-read on a recycled backing)."""
    ctx = context()
    canvas = (1920, 1088)
    mixer = sv.VideoMixer(ctx, canvas[0], canvas[1], sv.NV12, asset_id="lap")
    imgs = [scenes.random_image(O.NV12, 1920, 1088, 6000 + i) for i in range(5)]
    # (one revision for all five: a sample of the tick before would otherwise persist one extra frame, mix.video.swift:104-107,114)
    gl = [_place(to_gpu(ctx, im, f"lap{i}"), canvas, (1920, 1088), (0, 0), canvas, z=0, opacity=1.0, revision="lap") for i, im in enumerate(imgs)]
    outs = []
    for t in range(25):
        mixer.push_many([gl[t % 5]])
        outs.append((t, mixer.mix(t, wait=False).download(ctx, retain_gpu_buffer=True, wait=False)))
    want = {}
    for t, o in outs:
        o.wait()
        k = t % 5
        if k not in want:
            want[k] = _oracle_mix(O.NV12, canvas, [gl[k]], [imgs[k]])
        got = o.host_bytes()
        assert (got == want[k]).all(), f"tick {t}: {first_diff(got, want[k])}"
    mixer.close()


def test_tick_many_one_call_end_to_end():
    """svb_video_mixer_tick_many: upload + push + mix + download of two mixers in one call, nothing waited for until the end; the bytes are
    the oracle's fold of each mixer's own layers."""
    ctx = context()
    canvas = (512, 256)
    mixers = [sv.VideoMixer(ctx, canvas[0], canvas[1], sv.NV12, asset_id=f"tick{m}") for m in range(2)]
    results = []
    for f in range(4):
        per_mixer, keep = [], []
        for m in range(2):
            imgs = [scenes.random_image(O.NV12, 256, 144, 3000 + 100 * f + 10 * m + i) for i in range(3)]
            hosts = []
            for i, im in enumerate(imgs):
                h = sv.create_picture_sample(256, 144, sv.NV12, f"t{m}a{i}", "w", pinned_from=ctx)
                h.set_host_bytes(im.data)
                hosts.append(_place(h, canvas, (256, 144), (30 * i + 17 * m, 25 * i), (380, 190), z=i, opacity=0.95 - 0.2 * i))
            per_mixer.append(hosts)
            keep.append((hosts, imgs))
        outs = sv.VideoMixer.tick_many(mixers, per_mixer, f, wait=False)
        results.append((outs, keep))
    for outs, keep in results:
        for o, (hosts, imgs) in zip(outs, keep):
            o.wait()
            want = _oracle_mix(O.NV12, canvas, hosts, imgs)
            got = o.host_bytes()
            assert (got == want).all(), first_diff(got, want)
    for m in mixers:
        m.close()


CUSTOM_SOURCE = r"""
// a caller's own kernel in the reference's argument convention (compute.cuda.swift:294-297): [out planes..., in planes..., uniforms, inStride];
// the launch is gcd(W,16) x gcd(H,16) blocks over the target, the output pitch is the launch width (kernels.cuda.swift:151)
struct ImageUniforms { float transform[16], textureTx[16], borderMatrix[16], fillColor[4], inSize[2], outSize[2], opacity, sampleTime, targetTime; };
extern "C" __global__ void my_invert_nv12(unsigned char* oY, unsigned char* oC, const unsigned char* iY, const unsigned char* iC,
                                          const ImageUniforms* u, const int* inStride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, W = gridDim.x * blockDim.x;
    const float v = 255.0f - (float)iY[y * inStride[0] + x];
    oY[y * W + x] = (unsigned char)(v * u->opacity + 0.5f);            // --fmad=false: the product is rounded before the sum
    if (!(x & 1) && !(y & 1)) {
        oC[(y / 2) * W + x] = iC[(y / 2) * inStride[1] + x + 1];         // swap U and V
        oC[(y / 2) * W + x + 1] = iC[(y / 2) * inStride[1] + x];
    }
}
"""


def test_custom_kernel_from_source():
    """buildComputeKernel(_:name:source:) (compute.cuda.swift:171-201) and ComputeKernel.custom(name:) (compute.swift:73): CUDA C source is
    compiled by NVRTC for sm_100a with --fmad=false, registered under its name, and launched by runComputeKernel with the reference's
    argument order and geometry."""
    ctx = context().sharing()
    ctx.build_compute_kernel_from_source("my_invert_nv12", CUSTOM_SOURCE)
    W, H = 320, 180
    src = scenes.random_image(O.NV12, W, H, 31337)
    g = to_gpu(ctx, src, "src")
    target = gpu_target(ctx, O.NV12, W, H)
    u = api.ImageUniforms()
    u.opacity = 0.5
    api.run_compute_kernel(ctx, [g], target, api.KERNEL_CUSTOM, uniforms=u, custom_name="my_invert_nv12", blends=True)
    ctx.synchronize()
    got = fetch(ctx, target)
    y = src.data[: W * H].astype(np.float32)
    want_y = ((np.float32(255.0) - y) * np.float32(0.5) + np.float32(0.5)).astype(np.uint8)
    c = src.data[W * H:].reshape(H // 2, W // 2, 2)
    want_c = c[:, :, ::-1].reshape(-1)
    assert (got[: W * H] == want_y).all(), first_diff(got[: W * H], want_y)
    assert (got[W * H:] == want_c).all(), first_diff(got[W * H:], want_c)
    # a name that was never built is not found; source that does not compile reports the compiler's log
    with pytest.raises(sv.ComputeError) as e:
        api.run_compute_kernel(ctx, [g], target, api.KERNEL_CUSTOM, uniforms=u, custom_name="nobody", blends=True)
    assert e.value.name == "computeKernelNotFound"
    with pytest.raises(sv.ComputeError) as e:
        ctx.build_compute_kernel_from_source("broken", 'extern "C" __global__ void broken() { this does not compile; }')
    assert e.value.name == "compilerError" and "error" in str(e.value)


@pytest.mark.parametrize("mode", [sv.MixMode.FUSED_RING, sv.MixMode.FUSED_TILED, sv.MixMode.PER_LAYER])
def test_layers_placed_by_the_native_animator(mode):
    """layers placed the way bench.py and a SwiftVideo composition place them -- svb_animate_picture and
    PictureAnimator (parent, anchors, a transition under way) -- composed by the mixer and compared byte for byte with the
    oracle fold; the uniforms the host derives are also checked against a float64 inverse of the restated animator's matrices."""
    from oracle import animator_ref as R
    ctx = context()
    canvas = (384, 216)
    mixer = sv.VideoMixer(ctx, canvas[0], canvas[1], sv.NV12, asset_id="mixer", workspace_id="ws")
    mixer.set_mode(mode)
    fmts = [O.NV12, O.Y420P, O.BGRA, O.NV12]
    imgs = [scenes.random_image(f, 128, 72, 4200 + i) for i, f in enumerate(fmts)]
    gpu = [to_gpu(ctx, im, f"asset{i}") for i, im in enumerate(imgs)]
    base = dict(rotation=0.0, border=(0, 0, 0, 0), tex_offset=(0, 0), transparency=0.0)
    # layer 0: the bench's placement call, full canvas
    placed = [gpu[0].animate(canvas, (0, 0, 0.0), canvas, transparency=0.0)]
    # layer 1: a panel in mid-transition (grows and fades), layer 2 rides its bottom-right corner, layer 3 stretches with it
    panel = sv.PictureAnimator(canvas)
    p0, p1 = dict(base, pos=(20, 16, 1.0), size=(160, 90)), dict(base, pos=(20, 16, 1.0), size=(240, 150), transparency=0.4)
    panel.set_state(sv.element_state(p0["pos"], p0["size"]), 0.0, 0.0)
    badge = sv.PictureAnimator(canvas, parent=panel)
    bs = dict(base, pos=(100, 50, 2.0), size=(48, 28), transparency=0.25, border=(2, 2, 2, 2), fill=(0.9, 0.1, 0.1, 1.0))
    badge.set_state(sv.element_state(bs["pos"], bs["size"], transparency=0.25, border=bs["border"], fill=bs["fill"], anchors=sv.ANCHOR_BOTTOM_RIGHT))
    strip = sv.PictureAnimator(canvas, parent=panel)
    ss = dict(base, pos=(8, 60, 3.0), size=(140, 24), aspect=2)
    strip.set_state(sv.element_state(ss["pos"], ss["size"], aspect=2, anchors=sv.ANCHOR_TOP_LEFT | sv.ANCHOR_TOP_RIGHT))
    for a, g in ((panel, gpu[1]), (badge, gpu[2]), (strip, gpu[3])):   # first frame latches the initial parent state
        assert a.apply(g, 0.0) is not None
    panel.set_state(sv.element_state(p1["pos"], p1["size"], transparency=0.4), 2.0, 1.0)
    now = 1.5
    placed += [panel.apply(gpu[1], now), badge.apply(gpu[2], now), strip.apply(gpu[3], now)]
    assert [p.z_index() for p in placed] == [1, 2, 3, 4]
    # the restated animator says where they are
    pm0 = R.picture_state((128, 72), p0)["matrix"]
    pst = R.picture_state((128, 72), p0, nxt=p1, pct=0.25)
    want_states = [R.picture_state((128, 72), dict(base, pos=(0, 0, 0.0), size=canvas)), pst,
                   R.picture_state((128, 72), bs, anchors=[R.BR], parent=pst["matrix"], initial_parent=pm0),
                   R.picture_state((128, 72), ss, anchors=[R.TL, R.TR], parent=pst["matrix"], initial_parent=pm0)]
    tgt = sv.create_picture_sample(canvas[0], canvas[1], sv.NV12, "t", "w")
    for p, w, op in zip(placed, want_states, (1.0, 0.9, 0.75 * 0.9, 1.0 * 0.9)):
        i = p.info()
        assert np.allclose(np.array(i.matrix[:]), R.project(canvas, w["matrix"]), rtol=1e-5, atol=1e-6)
        assert abs(i.opacity - op) < 1e-6
        u = api.make_image_uniforms(p, tgt)
        for got, mem in ((u.transform, R.project(canvas, w["matrix"])), (u.border_matrix, R.project(canvas, w["border_matrix"])), (u.texture_transform, w["texture_matrix"])):
            inv = np.linalg.inv(np.asarray(mem, dtype=np.float64).reshape(4, 4).T)     # standard (row, col); uniforms hold its rows
            assert np.allclose(np.array(got[:]).reshape(4, 4), inv, rtol=3e-6, atol=3e-6)
    # the badge sits at panel origin + its own position + the panel's growth so far (20, 15)
    bm = np.array(placed[2].info().matrix[:]).reshape(4, 4)
    assert abs((bm[3, 0] + 1) * canvas[0] / 2 - (20 + 100 + 20)) < 1e-3 and abs((bm[3, 1] + 1) * canvas[1] / 2 - (16 + 50 + 15)) < 1e-3
    for p in placed:
        assert mixer.push(p)
    got = fetch(ctx, mixer.mix(1000))
    want = _oracle_mix(O.NV12, canvas, placed, imgs)
    assert (got == want).all(), first_diff(got, want)
    mixer.close()


def test_two_threads_drive_mixers_of_one_context():
    """any thread may call; mixers sharing one context are serialised on its compute stream.  Two host threads tick two
    mixers of the same context concurrently (asynchronous mixes, downloads joined at the end): every frame of both is bit-exact."""
    import threading
    ctx = context()
    canvas = (256, 144)
    results, errors = {}, []

    def worker(idx):
        try:
            mixer = sv.VideoMixer(ctx, canvas[0], canvas[1], sv.NV12, asset_id=f"mixer{idx}", workspace_id="ws")
            imgs = [scenes.random_image(O.NV12, 128, 72, 9000 + 10 * idx + k) for k in range(3)]
            gpu = [to_gpu(ctx, im, f"t{idx}a{k}") for k, im in enumerate(imgs)]
            frames = []
            for tick in range(24):
                placed = [_place(gpu[0], canvas, (128, 72), (0, 0), canvas, z=0, revision="a"),
                          _place(gpu[1], canvas, (128, 72), (8 + 4 * tick, 6 + idx), (140, 80), z=1, opacity=0.6, revision="b"),
                          _place(gpu[2], canvas, (128, 72), (60, 20 + 2 * tick), (150, 90), z=2, opacity=0.85, revision="c")]
                for p in placed:
                    mixer.push(p)
                out = mixer.mix(1000 * (tick + 1), wait=False).download(ctx, retain_gpu_buffer=True, wait=False)
                frames.append((out, placed))
            got = []
            for out, placed in frames:
                out.wait()
                got.append((out.host_bytes().copy(), placed))
            results[idx] = (got, imgs)
            mixer.close()
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for idx in range(2):
        got, imgs = results[idx]
        for tick, (bytes_, placed) in enumerate(got):
            want = _oracle_mix(O.NV12, canvas, placed, imgs)
            assert (bytes_ == want).all(), (idx, tick, first_diff(bytes_, want))


def test_batch_upload_coalesces_neighbours():
    """svb_upload_compute_pictures: pictures created one after the other from the context's page-locked pool lie next to each other in host
    memory and go up as ONE copy into ONE device block (device offsets mirror host offsets); a pageable picture in between and a GPU picture
    take the ordinary paths; every picture's bytes arrive intact and compose bit-exactly."""
    ctx = sv.make_compute_context(0)  # a context of its own: its page-locked pool holds no recycled blocks yet, so neighbours are neighbours
    canvas = (256, 144)
    fmts = [O.NV12, O.NV12, O.BGRA, O.Y420P, O.NV12, O.NV12]
    sizes = [(256, 144), (128, 72), (64, 32), (128, 72), (130, 70), (128, 72)]   # 130x70: planes not 256-byte aligned -> its own copy
    imgs = [scenes.random_image(f, w, h, 9500 + i) for i, (f, (w, h)) in enumerate(zip(fmts, sizes))]
    host = []
    for i, im in enumerate(imgs):
        p = sv.create_picture_sample(im.width, im.height, FMT[im.format], f"p{i}", "w", pinned_from=None if i == 3 else ctx)  # 3: pageable
        p.set_host_bytes(im.data)
        host.append(p)
    already = to_gpu(ctx, imgs[0], "gpu0")
    ups = sv.upload_many(ctx, host + [already], retain_cpu_buffer=False, wait=False)
    assert len(ups) == 7 and ups[6].device_frame().planes[0].ptr == already.device_frame().planes[0].ptr
    for u in ups[:6]:
        u.wait()
    hp = [h.info().planes[0].host for h in host]
    dp = [u.device_frame().planes[0].ptr for u in ups]
    # 0, 1, 2 are neighbours in the pool (created in sequence, all tightly laid out): one block, device offsets = host offsets
    assert dp[1] - dp[0] == hp[1] - hp[0] > 0 and dp[2] - dp[0] == hp[2] - hp[0]
    f0 = ups[0].device_frame()
    assert f0.planes[1].ptr - f0.planes[0].ptr == 256 * 144
    for u, im in zip(ups[:6], imgs):
        assert (fetch(ctx, u) == im.data).all()
    placed = [_place(ups[0], canvas, sizes[0], (0, 0), canvas, z=0), _place(ups[1], canvas, sizes[1], (10, 8), (150, 84), z=1, opacity=0.7),
              _place(ups[2], canvas, sizes[2], (100, 40), (96, 48), z=2, opacity=0.9), _place(ups[3], canvas, sizes[3], (30, 60), (128, 72), z=3, opacity=0.5),
              _place(ups[4], canvas, sizes[4], (140, 10), (100, 54), z=4), _place(ups[5], canvas, sizes[5], (4, 90), (96, 50), z=5, opacity=0.8)]
    mixer = sv.VideoMixer(ctx, canvas[0], canvas[1], sv.NV12, asset_id="m", workspace_id="w")
    for p in placed:
        assert mixer.push(p)
    got = fetch(ctx, mixer.mix(1000))
    want = _oracle_mix(O.NV12, canvas, placed, imgs)
    assert (got == want).all(), first_diff(got, want)
    mixer.close()
