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


def _oracle_mix(target_fmt, canvas, placed, images):
    """Oracle fold using the uniforms the host side itself derives from the pictures' matrices."""
    tgt = sv.create_picture_sample(canvas[0], canvas[1], FMT[target_fmt], "t", "w")
    us = []
    for p in placed:
        u = api.make_image_uniforms(p, tgt)
        ou = O.Uniforms()
        C.memmove(C.byref(ou), C.byref(u), 236)
        us.append(ou)
    want = O.Image(target_fmt, canvas[0], canvas[1])
    assert O.best()[0].mix(want, images, us) == 0
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
    # BGRA target: clear works (img_clear_bgra), any layer has no kernel
    m3 = sv.VideoMixer(ctx, 64, 32, sv.BGRA, asset_id="m3")
    got = fetch(ctx, m3.mix(0))
    want = O.Image(O.BGRA, 64, 32)
    O.best()[0].clear(want)
    assert (got == want.data).all()
    m3.push(to_gpu(ctx, scenes.random_image(O.BGRA, 64, 32, 2), "b"))
    with pytest.raises(sv.ComputeError) as e:
        m3.mix(1)
    assert e.value.name == "computeKernelNotFound"   # img_bgra_bgra is in the enum but has no kernel on Linux
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
        placed = [_place(h.upload(ctx, retain_cpu_buffer=False), canvas, (256, 144), (40 * i, 20 * i), (400, 200), z=i, opacity=0.9 - 0.2 * i)
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
