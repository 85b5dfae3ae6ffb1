"""CPU: the C-ABI library loads, exports every symbol include/svb200.h declares, and its host-side logic
(kernel-name map, picture layouts, zIndex, uniforms) behaves like the reference's.  No compute calls."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import scenes
import swiftvideo_b200 as sv
from oracle import oracle as O
from swiftvideo_b200 import animator, api

ROOT = Path(__file__).resolve().parent.parent


def test_every_declared_symbol_is_exported():
    header = (ROOT / "include" / "svb200.h").read_text()
    names = set(re.findall(r"^(?:svb_status|int|void|const char\*|unsigned long long)\s+(svb_[a-z0-9_]+)\s*\(", header, re.M))
    assert len(names) >= 40
    for n in sorted(names):
        assert hasattr(sv.lib, n), f"{n} declared in include/svb200.h but not exported by libsvb200.so"


def test_kernel_name_map():
    """The reference's only compute test (Tests/swiftVideoInternalTests/computeTests.swift:9-39): name -> enum."""
    names = ["img_nv12_nv12", "img_bgra_nv12", "img_rgba_nv12", "img_bgra_bgra", "img_y420p_y420p", "img_y420p_nv12", "img_clear_nv12",
             "img_clear_yuvs", "img_clear_bgra", "img_clear_y420p", "img_rgba_y420p", "img_bgra_y420p"]
    for n in names:
        k = api.default_compute_kernel_from_string(n)
        assert sv.lib.svb_compute_kernel_name(k).decode() == n
    # img_clear_rgba maps to img_clear_bgra (compute.swift:101)
    assert sv.lib.svb_compute_kernel_name(api.default_compute_kernel_from_string("img_clear_rgba")).decode() == "img_clear_bgra"
    with pytest.raises(sv.ComputeError) as e:
        api.default_compute_kernel_from_string("img_nv12_y420p")
    assert e.value.name == "invalidValue"


def test_module_image_has_the_reference_entry_points():
    img = sv.kernel_module_image()
    assert img[:4] == b"\x7fELF"
    for n in ("img_clear_nv12", "img_clear_y420p", "img_clear_bgra", "img_nv12_nv12", "img_y420p_nv12", "img_y420p_y420p",
              "img_bgra_nv12", "img_rgba_nv12", "img_bgra_y420p", "img_rgba_y420p", "svb_mix_tiled", "svb_mix_tables", "svb_mix_ring", "svb_strip_tables", "svb_mix_generic",
              "svb_scale_convert", "svb_scale_convert_any",
              # the operators upstream names without a Linux kernel: img_bgra_bgra after its Metal text, the yuvs clear,
              # and the NV21 / 4:2:2 / 4:4:4 sources under findKernel's naming rule
              "img_bgra_bgra", "img_clear_yuvs", "img_nv21_nv12", "img_y422p_nv12", "img_y444p_nv12", "img_y422p_y420p", "img_y444p_y420p"):
        assert n.encode() in img


def test_picture_layouts():
    """planesForFormat / buffersForPlanes (sample.pict.linux.swift:275-311): one allocation, planes back to back."""
    for fmt, ofmt in ((sv.NV12, O.NV12), (sv.Y420P, O.Y420P), (sv.BGRA, O.BGRA), (sv.RGBA, O.RGBA), (sv.NV21, O.NV21), (sv.Y422P, O.Y422P),
                      (sv.Y444P, O.Y444P)):
        p = sv.create_picture_sample(64, 36, fmt, "a", "w")
        i = p.info()
        layout, total = O.plane_layout(ofmt, 64, 36)
        assert i.plane_count == len(layout)
        base = i.planes[0].host
        for k, (off, w, h, stride, nc) in enumerate(layout):
            pl = i.planes[k]
            assert (int(pl.width), int(pl.height), pl.stride, pl.components, pl.bit_depth) == (w, h, stride, nc, 8)
            assert pl.host - base == off
        assert i.buffer_type == api.BUFFER_CPU and list(i.fill_color) == [0, 0, 0, 1] and i.opacity == 1.0
    # P010 (ours): NV12's plane shapes, two bytes per component, ten significant bits
    i = sv.create_picture_sample(64, 36, sv.P010, "a", "w").info()
    assert i.plane_count == 2 and i.pixel_format == sv.P010
    assert [(int(pl.width), int(pl.height), pl.stride, pl.components, pl.bit_depth) for pl in (i.planes[0], i.planes[1])] == [(64, 36, 128, 1, 10), (32, 18, 128, 2, 10)]
    assert i.planes[1].host - i.planes[0].host == 128 * 36
    with pytest.raises(sv.ComputeError) as e:
        sv.create_picture_sample(0, 10, sv.NV12)
    assert e.value.name == "invalidOperation"
    with pytest.raises(sv.ComputeError) as e:
        sv.create_picture_sample(16, 16, 13)  # past the last pixel format
    assert e.value.name == "badInputData"
    with pytest.raises(sv.ComputeError) as e:
        sv.create_picture_sample(16, 16, api.SHAPE)  # planesForFormat's default branch (sample.pict.linux.swift:289-291)
    assert e.value.name == "badInputData"


def test_z_index_and_with():
    p = sv.create_picture_sample(16, 16, sv.NV12, "asset", "w")
    for z in (0.0, 1.4, 2.5, -1.0):
        m, t, b = animator.picture_state((64, 64), (16, 16), (3, 4), (20, 20), z=z)
        q = p.with_(matrix=m)
        assert q.z_index() == int(np.floor(abs(z + 1) + 0.5) * np.sign(z + 1))  # round(pos.z + 1), half away from zero
    q = p.with_(opacity=0.25, fill_color=(0.1, 0.2, 0.3, 0.4), revision="rev")
    i = q.info()
    assert i.opacity == 0.25 and np.allclose(list(i.fill_color), [0.1, 0.2, 0.3, 0.4])
    assert list(p.info().fill_color) == [0, 0, 0, 1]  # the original is immutable


def test_uniforms_host_logic():
    """applyComputeImage's uniforms (inverse.transpose of the three matrices, compute.swift:149-161) against an
    independent float64 construction; fp32 inverse rounding is unpinned upstream, hence a tolerance."""
    canvas = (1280, 720)
    tgt = sv.create_picture_sample(canvas[0], canvas[1], sv.NV12, "t", "w")
    src = sv.create_picture_sample(640, 360, sv.Y420P, "s", "w")
    for kw in (dict(pos=(0, 0), size=(640, 720), aspect="fill"), dict(pos=(100, 50), size=(320, 200), rotation=0.4, border=(3, 4, 5, 6)),
               dict(pos=(-20, 600), size=(900, 300), aspect="fit", z=3.0)):
        m, t, b = animator.picture_state(canvas, (640, 360), **kw)
        q = src.with_(matrix=m, texture_matrix=t, border_matrix=b, opacity=0.5, fill_color=(1, 0, 0, 1))
        u = api.make_image_uniforms(q, tgt)
        want = scenes.layer_uniforms(canvas, (640, 360), kw["pos"], kw["size"], rotation=kw.get("rotation", 0.0), z=kw.get("z", 0.0),
                                     opacity=0.5, fill=(1, 0, 0, 1), border=kw.get("border", (0, 0, 0, 0)), aspect=kw.get("aspect", "none"))
        for a, w in ((u.transform, want.transform), (u.texture_transform, want.textureTx), (u.border_matrix, want.borderMatrix)):
            assert np.allclose(np.array(a[:]), np.array(w[:]), rtol=2e-5, atol=2e-5)
        assert list(u.input_size) == [640, 360] and list(u.output_size) == [1280, 720] and u.opacity == 0.5
        if not kw.get("rotation"):  # axis-aligned layers must give exact zeros: the separable fast path keys on them
            tr = np.array(u.transform[:]).reshape(4, 4)
            assert tr[0, 1] == 0 and tr[1, 0] == 0 and tr[2, 0] == 0 and tr[2, 1] == 0 and tr[3, 0] == 0 and tr[3, 1] == 0


@pytest.mark.skipif(sv.available_compute_devices() > 0, reason="a GPU is present")
def test_no_cpu_fallback():
    """Without a device every compute entry point fails loudly (ComputeError.deviceNotAvailable)."""
    with pytest.raises(sv.ComputeError) as e:
        sv.make_compute_context()
    assert e.value.name == "deviceNotAvailable"
    with pytest.raises(sv.ComputeError):
        sv.VideoMixer(None, 64, 64).mix(0)


def test_animate_picture_matches_python_animator():
    """PictureAnimator.impl in native code (svb_animate_picture) against the numpy helper the tests place layers with."""
    canvas = (1280, 720)
    src = sv.create_picture_sample(640, 360, sv.NV12, "s", "w")
    for kw, asp in ((dict(pos=(0, 0), size=(640, 720)), "fill"), (dict(pos=(100.5, 50.25), size=(320, 200), rotation=0.4, border=(3, 4, 5, 6)), "none"),
                    (dict(pos=(-20, 600, 3.0), size=(900, 300)), "fit")):
        pos = kw["pos"]
        m, t, b = animator.picture_state(canvas, (640, 360), pos[:2], kw["size"], rotation=kw.get("rotation", 0.0), z=pos[2] if len(pos) > 2 else 0.0,
                                         border=kw.get("border", (0, 0, 0, 0)), aspect=asp)
        q = src.animate(canvas, pos, kw["size"], rotation=kw.get("rotation", 0.0), border=kw.get("border", (0, 0, 0, 0)),
                        aspect={"none": 0, "fit": 1, "fill": 2}[asp], fill=(0.2, 0.4, 0.6, 0.8), transparency=0.25, parent_opacity=0.5, revision="r")
        i = q.info()
        assert np.allclose(np.array(i.matrix[:]), m, rtol=1e-5, atol=1e-6)
        assert np.allclose(np.array(i.texture_matrix[:]), t, rtol=1e-5, atol=1e-6)
        assert np.allclose(np.array(i.border_matrix[:]), b, rtol=1e-5, atol=1e-6)
        assert abs(i.opacity - 0.375) < 1e-7 and np.allclose(list(i.fill_color), [0.2, 0.4, 0.6, 0.8])
        assert i.z_index == int(round((pos[2] if len(pos) > 2 else 0.0) + 1))       # ortho's m43 = 1
    # no fill colour set -> (0,0,0,0) (animator.pic.swift:334-342); centre origin shifts by half the size
    q = src.animate(canvas, (100, 100), (50, 40), top_left=False)
    assert list(q.info().fill_color) == [0, 0, 0, 0]
    mm = np.array(q.info().matrix[:]).reshape(4, 4)
    assert abs(mm[3, 0] - (2.0 / 1280 * 75 - 1)) < 1e-6 and abs(mm[3, 1] - (2.0 / 720 * 80 - 1)) < 1e-6


def test_picture_sample_from_planes():
    """Decoder-style planes with their own linesize (the reference wraps AVFrame.data/linesize into planes,
    dec.video.ffmpeg.swift:176-190): strides are kept, bytes are copied, short strides are rejected."""
    import swiftvideo_b200 as sv
    from swiftvideo_b200 import api
    rng = np.random.default_rng(5)
    y = rng.integers(0, 256, (48, 64 + 16), dtype=np.uint8)
    u = rng.integers(0, 256, (24, 32 + 16), dtype=np.uint8)
    v = rng.integers(0, 256, (24, 32 + 8), dtype=np.uint8)
    p = api.picture_sample_from_planes(64, 48, sv.Y420P, [y, u, v], "cam", "ws")
    i = p.info()
    assert i.plane_count == 3 and i.pixel_format == sv.Y420P
    assert [i.planes[k].stride for k in range(3)] == [80, 48, 40]
    assert [i.planes[k].width for k in range(3)] == [64, 32, 32]
    got = p.host_planes()
    assert (got[0] == y).all() and (got[1] == u).all() and (got[2] == v).all()
    y[:] = 0  # the sample owns a copy
    assert got[0].any()
    with pytest.raises(sv.ComputeError) as e:
        api.picture_sample_from_planes(64, 48, sv.NV12, [np.zeros((48, 60), np.uint8), np.zeros((24, 64), np.uint8)])
    assert "stride" in str(e.value)
    with pytest.raises(sv.ComputeError):
        api.picture_sample_from_planes(64, 48, sv.NV12, [np.zeros((48, 64), np.uint8)])  # plane count


def test_gpu_barriers_pass_through_and_error_shape():
    """GPUBarrierUpload / GPUBarrierDownload (compute.swift:175-198, :232-255) without a GPU: a CPU sample goes through the download barrier
    untouched (`.just($0)`); the upload barrier on a box without a device fails with upstream's event error
    EventError("barrier.upload", -1, "<error>", assetId:)."""
    p = sv.create_picture_sample(64, 36, sv.NV12, "asset-7", "w")
    if sv.available_compute_devices() > 0:
        pytest.skip("a GPU is present: covered by tests/test_gpu_mixer.py::test_barriers_are_idempotent")
    h, err = C.c_void_p(), api.EventError()
    rc = sv.lib.svb_gpu_barrier_upload(None, p._h, 1, 1, C.byref(h), C.byref(err))  # a NULL context is an invalid value, not a crash
    assert rc == 4 and not h.value
    try:
        ctx = sv.make_compute_context(0)
    except sv.ComputeError as e:
        assert e.name == "deviceNotAvailable"
        return
    same, err = p.barrier_download(ctx)  # (only reached on a box with a driver but no usable device)
    assert err is None and same.same_sample(p)
