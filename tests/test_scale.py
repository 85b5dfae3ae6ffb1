"""The convert+scale operator (NV12 / P010 -> BGRA, bilinear / Lanczos-3): BASELINE.json configs 2 and 5.

The reference has no such operator (SURVEY.md "five facts" 2-3), so parity is pinned in two steps:
  * oracle/scale_oracle.c is the definition; it is held against libswscale 9.1 -- the comparator BASELINE.json names -- in
    swscale's accurate mode on smooth content: at most ONE code value apart (swscale accumulates in 14/15-bit fixed point),
  * the CUDA kernel must equal the definition bit for bit (random content, ragged sizes, both formats, both filters, up- and
    down-scaling, BASELINE's full sizes), called through the C ABI.
"""
import numpy as np
import pytest

import swscale_util as S
from oracle import oracle as O

needs_sws = pytest.mark.skipif(not S.available(), reason="libswscale (OpenCV wheel) not loadable")


# ---- CPU: tables, the definition against libswscale ---------------------------------------------------------------------

@pytest.mark.parametrize("filt", [O.SC_BILINEAR, O.SC_LANCZOS3])
@pytest.mark.parametrize("src_n,dst_n", [(3840, 1920), (2160, 1080), (1920, 1280), (1080, 720), (360, 640), (1080, 1080), (959, 413), (2, 1), (2, 7)])
def test_filter_tables_match_oracle(filt, src_n, dst_n):
    """The library's filter tables (host code, no device) equal the oracle's bit for bit, sum to one and stay within reach."""
    import swiftvideo_b200 as sv
    f1, w1 = sv.scale_filter_table(filt, src_n, dst_n)
    f2, w2 = O.scale_table(filt, src_n, dst_n)
    assert w1.shape == w2.shape and (f1 == f2).all()
    assert (w1.view(np.uint32) == w2.view(np.uint32)).all()
    assert np.allclose(w1.sum(axis=1), 1.0, atol=1e-6)
    assert (np.diff(f1) >= 0).all()  # firsts are monotone: a tile's footprint is bounded by its end rows / columns


def test_filter_table_errors():
    import swiftvideo_b200 as sv
    with pytest.raises(sv.ComputeError) as e:
        sv.scale_filter_table(7, 10, 10)
    assert e.value.name == "invalidValue"
    with pytest.raises(sv.ComputeError):
        sv.scale_filter_table(0, 0, 10)


SWS_CASES = [
    ("nv12", 8, O.SC_NV12, O.SC_BILINEAR, S.SWS_BILINEAR, (640, 360), (426, 240)),
    ("nv12", 8, O.SC_NV12, O.SC_BILINEAR, S.SWS_BILINEAR, (640, 360), (640, 360)),
    ("nv12", 8, O.SC_NV12, O.SC_LANCZOS3, S.SWS_LANCZOS, (320, 180), (640, 360)),
    ("p010le", 10, O.SC_P010, O.SC_LANCZOS3, S.SWS_LANCZOS, (960, 540), (480, 270)),
    ("p010le", 10, O.SC_P010, O.SC_BILINEAR, S.SWS_BILINEAR, (480, 270), (854, 480)),
    ("nv12", 8, O.SC_NV12, O.SC_BILINEAR, S.SWS_BILINEAR, (1920, 1080), (1280, 720)),  # BASELINE cfg 2's scale
]


@needs_sws
@pytest.mark.parametrize("name,bits,fmt,filt,flag,src,dst", SWS_CASES)
def test_definition_within_one_code_of_swscale(name, bits, fmt, filt, flag, src, dst):
    pic = S.smooth_picture(bits, src[0], src[1], seed=3)
    ours = O.scale_convert(fmt, filt, pic, src[0], src[1], dst[0], dst[1]).astype(np.int32)
    sc = S.Scaler(name, src[0], src[1], dst[0], dst[1], flag | S.SWS_ACCURATE_RND | S.SWS_FULL_CHR_H_INT)
    ref = sc.run(pic).astype(np.int32)
    sc.close()
    d = np.abs(ours - ref)
    assert d[:, :, 3].max() == 0  # alpha 255 on both sides
    assert d.max() <= 1, f"max |diff| {d.max()}"  # TOLERANCE: one 8-bit code value
    assert d.mean() < 0.02


def _scale_golden():
    from pathlib import Path
    z = np.load(Path(__file__).resolve().parent / "golden" / "scale_cases.npz")
    return z, [str(n) for n in z["names"]]


@pytest.mark.parametrize("name", _scale_golden()[1])
def test_definition_against_committed_fixtures(name):
    """tests/golden/scale_cases.npz (made by make_scale_golden.py): the definition still produces its stored bytes, and they lie within
    one code value of the stored libswscale bytes -- checked without libswscale at hand."""
    z, _ = _scale_golden()
    fmt, filt, sw, sh, dw, dh = (int(v) for v in z[f"{name}/meta"])
    got = O.scale_convert(fmt, filt, z[f"{name}/src"], sw, sh, dw, dh)
    assert (got == z[f"{name}/definition"]).all()
    d = np.abs(got.astype(np.int32) - z[f"{name}/swscale"].astype(np.int32))
    assert d.max() <= 1  # TOLERANCE: one 8-bit code value


def test_definition_known_answers():
    """Flat pictures: a resize of a constant is that constant, and the colour matrix maps video black / white / grey as BT.601
    limited range says."""
    for bits, fmt in ((8, O.SC_NV12), (10, O.SC_P010)):
        for (y, u, v), want in (((16, 128, 128), (0, 0, 0)), ((235, 128, 128), (255, 255, 255)), ((126, 128, 128), (128, 128, 128))):
            w, h = 64, 36
            if bits == 8:
                pic = np.concatenate([np.full(w * h, y, np.uint8), np.tile(np.array([u, v], np.uint8), w * h // 4)])
            else:
                pic = np.concatenate([np.full(w * h, (y * 4) << 6, np.uint16), np.tile(np.array([(u * 4) << 6, (v * 4) << 6], np.uint16), w * h // 4)]).view(np.uint8)
            for filt in (O.SC_BILINEAR, O.SC_LANCZOS3):
                out = O.scale_convert(fmt, filt, pic, w, h, 40, 30)
                assert (out[:, :, 3] == 255).all()
                assert (out[:, :, :3] == np.array(want[::-1], np.uint8)).all(), (bits, filt, out[0, 0])


def test_definition_bad_arguments():
    pic = np.zeros(64 * 36 * 3 // 2, np.uint8)
    with pytest.raises(ValueError):
        O.scale_convert(5, O.SC_BILINEAR, pic, 64, 36, 32, 18)
    with pytest.raises(ValueError):
        O.scale_convert(O.SC_NV12, 9, pic, 64, 36, 32, 18)


# ---- GPU: the kernel against the definition -------------------------------------------------------------------------------

def _random_picture(fmt, w, h, seed):
    rng = np.random.default_rng(seed)
    if fmt == O.SC_NV12:
        return rng.integers(0, 256, size=w * h * 3 // 2, dtype=np.uint8)
    return (rng.integers(0, 1024, size=w * h * 3 // 2, dtype=np.uint16) << 6).view(np.uint8)  # every 10-bit code, full range


def _gpu_scale(ctx, fmt, filt, pic, src, dst):
    import swiftvideo_b200 as sv
    p = sv.create_picture_sample(src[0], src[1], sv.NV12 if fmt == O.SC_NV12 else sv.P010, "src", "test")
    p.set_host_bytes(pic)
    out = p.upload(ctx).scale_convert(ctx, dst[0], dst[1], sv.BGRA, filt)
    info = out.info()
    assert info.pixel_format == sv.BGRA and int(info.width) == dst[0] and int(info.height) == dst[1]
    return out.download(ctx).host_bytes().reshape(dst[1], dst[0], 4)


GPU_CASES = [
    (O.SC_NV12, O.SC_BILINEAR, (640, 360), (426, 240)),
    (O.SC_NV12, O.SC_BILINEAR, (640, 360), (640, 360)),
    (O.SC_NV12, O.SC_LANCZOS3, (320, 180), (640, 360)),
    (O.SC_P010, O.SC_LANCZOS3, (960, 540), (480, 270)),
    (O.SC_P010, O.SC_BILINEAR, (64, 36), (200, 100)),
    (O.SC_NV12, O.SC_BILINEAR, (130, 70), (97, 53)),      # ragged: partial tiles on both axes, odd destination
    (O.SC_P010, O.SC_LANCZOS3, (258, 130), (101, 67)),    # ragged, 2.55 : 1 (16 taps: the most the kernel takes)
    (O.SC_NV12, O.SC_LANCZOS3, (2, 2), (5, 3)),           # the smallest source
    (O.SC_NV12, O.SC_BILINEAR, (256, 64), (33, 9)),       # 7.8 : 1 bilinear (16 taps)
    (O.SC_NV12, O.SC_LANCZOS3, (512, 256), (128, 64)),    # 4 : 1 Lanczos: 24 taps, the run-time tap loops
    (O.SC_P010, O.SC_LANCZOS3, (768, 384), (128, 64)),    # 6 : 1 Lanczos: 36 taps, the tile shrinks to 8 rows to fit shared memory
    (O.SC_NV12, O.SC_BILINEAR, (768, 96), (64, 8)),       # 12 : 1 bilinear: 24 taps
    (O.SC_P010, O.SC_BILINEAR, (3840, 64), (1280, 48)),   # wide: many column tiles, 3 : 1 across, 1.33 : 1 down
]


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,filt,src,dst", GPU_CASES)
def test_gpu_scale_bit_exact(fmt, filt, src, dst):
    import gpu_util
    pic = _random_picture(fmt, src[0], src[1], seed=5000 + src[0] + dst[1])
    got = _gpu_scale(gpu_util.context(), fmt, filt, pic, src, dst)
    want = O.scale_convert(fmt, filt, pic, src[0], src[1], dst[0], dst[1])
    bad = int((got != want).sum())
    assert bad == 0, f"{bad} bytes differ; first rows got {got[0, :2]} want {want[0, :2]}"


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,filt,strides", [(O.SC_NV12, O.SC_LANCZOS3, (650, 652)),    # neither 16-byte chunks nor aligned rows: sample-by-sample staging
                                              (O.SC_NV12, O.SC_BILINEAR, (704, 672)),    # padded but aligned: 16-byte staging reads into the padding at the right edge
                                              (O.SC_P010, O.SC_LANCZOS3, (1284, 1300)),  # P010 rows that are only 4-byte aligned
                                              (O.SC_P010, O.SC_BILINEAR, (1344, 1280))])
def test_gpu_scale_decoder_strides(fmt, filt, strides):
    """Planes with the decoder's own linesize (dec.video.ffmpeg.swift:183): the aligned and the unaligned staging paths give the definition's bytes."""
    import gpu_util
    import swiftvideo_b200 as sv
    from swiftvideo_b200 import api
    w, h, dst = 640, 360, (300, 170)
    bps = 1 if fmt == O.SC_NV12 else 2
    pic = _random_picture(fmt, w, h, seed=77 + strides[0])
    rng = np.random.default_rng(9)
    y = rng.integers(0, 256, (h, strides[0]), dtype=np.uint8)          # the padding holds noise the kernel must never let through
    c = rng.integers(0, 256, (h // 2, strides[1]), dtype=np.uint8)
    y[:, :w * bps] = pic[:w * h * bps].reshape(h, w * bps)
    c[:, :w * bps] = pic[w * h * bps:].reshape(h // 2, w * bps)
    src = api.picture_sample_from_planes(w, h, sv.NV12 if fmt == O.SC_NV12 else sv.P010, [y, c], "cam", "ws").upload(gpu_util.context())
    got = src.scale_convert(gpu_util.context(), dst[0], dst[1], sv.BGRA, filt).download(gpu_util.context()).host_bytes().reshape(dst[1], dst[0], 4)
    want = O.scale_convert(fmt, filt, pic, w, h, dst[0], dst[1])
    assert int((got != want).sum()) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,filt,src,dst", [(O.SC_NV12, O.SC_BILINEAR, (1920, 1080), (1280, 720)),     # BASELINE cfg 2's convert + scale
                                              (O.SC_P010, O.SC_LANCZOS3, (3840, 2160), (1920, 1080))])  # BASELINE cfg 5
def test_gpu_scale_full_size(fmt, filt, src, dst):
    import gpu_util
    pic = _random_picture(fmt, src[0], src[1], seed=5)
    got = _gpu_scale(gpu_util.context(), fmt, filt, pic, src, dst)
    want = O.scale_convert(fmt, filt, pic, src[0], src[1], dst[0], dst[1])
    assert int((got != want).sum()) == 0


@pytest.mark.gpu
def test_cfg2_chain_full_size():
    """BASELINE cfg 2's reported extra, whole: 1920x1080 NV12 -> (convert + bilinear scale) -> 1280x720 BGRA -> the reference's img_bgra_nv12
    through the mixer -> 1280x720 NV12.  Each stage against its definition: the BGRA bytes equal oracle/scale_oracle.c's, the NV12 bytes equal
    the OpenCL-text oracle's clear + img_bgra_nv12 over those BGRA bytes."""
    import ctypes as C

    import gpu_util
    import swiftvideo_b200 as sv
    from oracle import oracle as MO
    from swiftvideo_b200 import api
    ctx = gpu_util.context()
    src, dst = (1920, 1080), (1280, 720)
    pic = _random_picture(O.SC_NV12, src[0], src[1], seed=2002)
    p = sv.create_picture_sample(src[0], src[1], sv.NV12, "cam", "test")
    p.set_host_bytes(pic)
    bgra = p.upload(ctx).scale_convert(ctx, dst[0], dst[1], sv.BGRA, O.SC_BILINEAR)
    want_bgra = O.scale_convert(O.SC_NV12, O.SC_BILINEAR, pic, src[0], src[1], dst[0], dst[1])
    assert (bgra.download(ctx, retain_gpu_buffer=True).host_bytes().reshape(dst[1], dst[0], 4) == want_bgra).all()
    mixer = sv.VideoMixer(ctx, dst[0], dst[1], sv.NV12, asset_id="mixer", workspace_id="ws")
    placed = bgra.animate(dst, (0, 0, 0.0), dst)
    mixer.push(placed)
    got = mixer.mix(1000).download(ctx).host_bytes()
    layer = MO.Image(MO.BGRA, dst[0], dst[1])
    layer.data[:] = want_bgra.reshape(-1)
    u = api.make_image_uniforms(placed, sv.create_picture_sample(dst[0], dst[1], sv.NV12, "t", "w"))
    ou = MO.Uniforms()
    C.memmove(C.byref(ou), C.byref(u), 236)
    want = MO.Image(MO.NV12, dst[0], dst[1])
    assert MO.best()[0].mix(want, [layer], [ou]) == 0
    assert (got == want.data).all()
    mixer.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", _scale_golden()[1])
def test_gpu_scale_against_committed_fixtures(name):
    """The kernel against the committed fixtures: the definition's bytes exactly, libswscale's within one code value."""
    import gpu_util
    z, _ = _scale_golden()
    fmt, filt, sw, sh, dw, dh = (int(v) for v in z[f"{name}/meta"])
    got = _gpu_scale(gpu_util.context(), fmt, filt, z[f"{name}/src"], (sw, sh), (dw, dh))
    assert (got == z[f"{name}/definition"]).all()
    assert np.abs(got.astype(np.int32) - z[f"{name}/swscale"].astype(np.int32)).max() <= 1


@pytest.mark.gpu
def test_gpu_scale_errors():
    import gpu_util
    import swiftvideo_b200 as sv
    ctx = gpu_util.context()
    host = sv.create_picture_sample(64, 36, sv.NV12, "a", "t")
    with pytest.raises(sv.ComputeError) as e:  # not uploaded
        host.scale_convert(ctx, 32, 18)
    assert e.value.name == "badInputData"
    nv = host.upload(ctx)
    with pytest.raises(sv.ComputeError) as e:  # only BGRA targets exist
        nv.scale_convert(ctx, 32, 18, sv.NV12)
    assert e.value.name == "computeKernelNotFound"
    wide = sv.create_picture_sample(2048, 64, sv.NV12, "w", "t").upload(ctx)
    with pytest.raises(sv.ComputeError) as e:  # Lanczos-3 at 128 : 1 is 768 taps a column: no tile of it fits shared memory
        wide.scale_convert(ctx, 16, 8, sv.BGRA, sv.FILTER_LANCZOS3)
    assert e.value.name == "notImplemented" and "768" in str(e.value)
    bgra = sv.create_picture_sample(64, 36, sv.BGRA, "b", "t").upload(ctx)
    with pytest.raises(sv.ComputeError) as e:  # sources are NV12 / P010
        bgra.scale_convert(ctx, 32, 18)
    assert e.value.name == "computeKernelNotFound"
    with pytest.raises(sv.ComputeError) as e:
        nv.scale_convert(ctx, 0, 18)
    assert e.value.name == "badTarget"
