"""ctypes binding of libsvb200.so (include/svb200.h), shaped like the reference's Swift surface:
makeComputeContext / createPictureSample / uploadComputePicture / VideoMixer ... (compute.swift, compute.cuda.swift,
sample.pict.linux.swift, mix.video.swift of unpause-live/SwiftVideo)."""
import ctypes as C
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
_LIB_PATH = _PKG / "libsvb200.so"

# PixelFormat / ComputeKernel / modes (svb200.h enums)
NV12, NV21, YUVS, ZVUY, Y420P, Y422P, Y444P, RGBA, BGRA, SHAPE, TEXT, INVALID, P010 = range(13)  # P010: ours (svb200.h)
FILTER_BILINEAR, FILTER_LANCZOS3 = 0, 1
BUFFER_SHARED, BUFFER_CPU, BUFFER_GPU, BUFFER_INVALID = range(4)
DEVICE_GPU = 0
KERNEL_CUSTOM = 15


class MixMode:
    FUSED, PER_LAYER, GENERIC, FUSED_GATHER, FUSED_TILED, FUSED_RING = 0, 1, 2, 3, 4, 5


STATUS_NAMES = ["ok", "invalidPlatform", "invalidDevice", "invalidOperation", "invalidValue", "invalidProgram", "invalidContext",
                "deviceNotAvailable", "outOfMemory", "compilerNotAvailable", "computeKernelNotFound", "badTarget", "badInputData",
                "badContextState", "compilerError", "unknownError", "notImplemented"]


class ComputeError(RuntimeError):
    """ComputeError (compute.swift:22-39)."""

    def __init__(self, code, message):
        super().__init__(f"{STATUS_NAMES[code] if 0 <= code < len(STATUS_NAMES) else code}: {message}")
        self.code = code
        self.name = STATUS_NAMES[code] if 0 <= code < len(STATUS_NAMES) else str(code)


class EventError(C.Structure):
    """svb_event_error: EventError(domain, code, description, assetId:) as the barriers emit it (compute.swift:190,247)."""

    _fields_ = [("domain", C.c_char * 32), ("code", C.c_int), ("description", C.c_char * 256), ("asset_id", C.c_char * 128)]


class ImageUniforms(C.Structure):
    """svb_image_uniforms == ImageUniforms (compute.swift:76-86), 236 bytes."""

    _fields_ = [("transform", C.c_float * 16), ("texture_transform", C.c_float * 16), ("border_matrix", C.c_float * 16),
                ("fill_color", C.c_float * 4), ("input_size", C.c_float * 2), ("output_size", C.c_float * 2),
                ("opacity", C.c_float), ("image_time", C.c_float), ("target_time", C.c_float)]


class ElementState(C.Structure):
    """svb_element_state: the ElementState fields PictureAnimator reads."""

    _fields_ = [("pic_pos", C.c_float * 3), ("size", C.c_float * 2), ("texture_offset", C.c_float * 2), ("border_size", C.c_float * 4),
                ("fill_color", C.c_float * 4), ("rotation", C.c_float), ("transparency", C.c_float), ("pic_aspect", C.c_int32),
                ("pic_origin", C.c_int32), ("has_fill_color", C.c_int32), ("hidden", C.c_int32), ("parent_anchors", C.c_uint32)]


class ComputedPictureState(C.Structure):
    """svb_computed_picture_state: unprojected matrices (Matrix4 memory order), fill colour, opacity."""

    _fields_ = [("matrix", C.c_float * 16), ("texture_matrix", C.c_float * 16), ("border_matrix", C.c_float * 16),
                ("fill_color", C.c_float * 4), ("opacity", C.c_float)]


ANCHOR_TOP_LEFT, ANCHOR_TOP_RIGHT, ANCHOR_BOTTOM_LEFT, ANCHOR_BOTTOM_RIGHT = 1, 2, 4, 8


class _DevicePlane(C.Structure):
    _fields_ = [("ptr", C.c_ulonglong), ("pitch", C.c_int32), ("width_bytes", C.c_int32), ("rows", C.c_int32), ("pad_", C.c_int32)]


class DeviceFrame(C.Structure):
    """svb_device_frame: a GPU sample's planes for an on-device consumer (device pointers, pitches, CUcontext, completion CUevent)."""

    _fields_ = [("device_index", C.c_int32), ("pixel_format", C.c_int32), ("plane_count", C.c_int32), ("width", C.c_float), ("height", C.c_float),
                ("planes", _DevicePlane * 3), ("context", C.c_void_p), ("ready_event", C.c_void_p)]


class _PlaneInfo(C.Structure):
    _fields_ = [("width", C.c_float), ("height", C.c_float), ("stride", C.c_int32), ("bit_depth", C.c_int32),
                ("components", C.c_int32), ("host", C.c_void_p), ("device", C.c_ulonglong), ("size", C.c_size_t)]


class _PictureInfo(C.Structure):
    _fields_ = [("pixel_format", C.c_int32), ("buffer_type", C.c_int32), ("width", C.c_float), ("height", C.c_float),
                ("plane_count", C.c_int32), ("planes", _PlaneInfo * 3), ("matrix", C.c_float * 16),
                ("texture_matrix", C.c_float * 16), ("border_matrix", C.c_float * 16), ("fill_color", C.c_float * 4),
                ("opacity", C.c_float), ("z_index", C.c_int32), ("pts", C.c_int64), ("time", C.c_int64), ("timescale", C.c_int64)]


def _load():
    if not _LIB_PATH.exists():
        raise ImportError(f"{_LIB_PATH} is missing: build it with `python -m swiftvideo_b200.build` "
                          "(there is no CPU fallback for the CUDA path)")
    l = C.CDLL(str(_LIB_PATH))
    l.svb_last_error.restype = C.c_char_p
    l.svb_version.restype = C.c_char_p
    l.svb_compute_kernel_name.restype = C.c_char_p
    l.svb_video_mixer_asset_id.restype = C.c_char_p
    l.svb_kernel_launch_count.restype = C.c_ulonglong
    l.svb_picture_release.restype = None
    l.svb_video_mixer_destroy.restype = None
    l.svb_timer_destroy.restype = None
    l.svb_create_picture_sample.argtypes = [C.c_float, C.c_float, C.c_int, C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p]
    l.svb_video_mixer_create.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_char_p, C.c_char_p, C.c_int64, C.c_int64,
                                         C.c_int64, C.c_void_p]
    l.svb_video_mixer_mix.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    l.svb_video_mixer_mix_many.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p]
    l.svb_video_mixer_push.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    l.svb_video_mixer_push_many.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    l.svb_video_mixer_set_mode.argtypes = [C.c_void_p, C.c_int]
    l.svb_picture_with.argtypes = [C.c_void_p] * 6 + [C.c_char_p, C.c_char_p, C.c_void_p]
    l.svb_picture_info_get.argtypes = [C.c_void_p, C.c_void_p]
    l.svb_picture_wait.argtypes = [C.c_void_p]
    l.svb_picture_release.argtypes = [C.c_void_p]
    l.svb_upload_compute_picture.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    l.svb_download_compute_picture.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    l.svb_video_mixer_tick_many.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    l.svb_picture_identity.restype = C.c_ulonglong
    for n in ("svb_picture_revision", "svb_picture_asset_id"):
        getattr(l, n).restype = C.c_char_p
        getattr(l, n).argtypes = [C.c_void_p]
    l.svb_picture_identity.argtypes = [C.c_void_p]
    l.svb_gpu_barrier_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    l.svb_gpu_barrier_download.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    l.svb_compose.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    l.svb_run_compute_kernel.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p,
                                         C.c_size_t, C.c_int]
    l.svb_apply_compute_image.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    l.svb_make_image_uniforms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    l.svb_make_compute_context.argtypes = [C.c_int, C.c_int, C.c_void_p]
    l.svb_create_compute_context_sharing.argtypes = [C.c_void_p, C.c_void_p]
    l.svb_destroy_compute_context.argtypes = [C.c_void_p]
    l.svb_begin_compute_pass.argtypes = [C.c_void_p]
    l.svb_end_compute_pass.argtypes = [C.c_void_p, C.c_int]
    l.svb_context_sm_count.argtypes = [C.c_void_p]
    l.svb_context_device_index.argtypes = [C.c_void_p]
    l.svb_build_compute_kernel.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    l.svb_default_compute_kernel_from_string.argtypes = [C.c_char_p, C.c_void_p]
    l.svb_kernel_module_image.argtypes = [C.c_void_p, C.c_void_p]
    l.svb_timer_create.argtypes = [C.c_void_p, C.c_void_p]
    for n in ("svb_timer_start", "svb_timer_stop", "svb_timer_destroy"):
        getattr(l, n).argtypes = [C.c_void_p]
    l.svb_timer_elapsed_ms.argtypes = [C.c_void_p, C.c_void_p]
    l.svb_video_mixer_destroy.argtypes = [C.c_void_p]
    l.svb_video_mixer_asset_id.argtypes = [C.c_void_p]
    l.svb_picture_sample_from_planes.argtypes = [C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_void_p,
                                                 C.c_void_p]
    l.svb_animate_picture.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_float, C.c_char_p, C.c_void_p]
    l.svb_upload_compute_pictures.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    l.svb_picture_device_frame.argtypes = [C.c_void_p, C.c_void_p]
    l.svb_picture_consumed_on.argtypes = [C.c_void_p, C.c_void_p]
    l.svb_gather_picture.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    l.svb_compute_picture_state.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    l.svb_animator_create.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_uint32, C.c_void_p]
    l.svb_animator_destroy.argtypes = [C.c_void_p]
    l.svb_animator_destroy.restype = None
    l.svb_animator_revision.argtypes = [C.c_void_p]
    l.svb_animator_revision.restype = C.c_char_p
    l.svb_animator_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double]
    l.svb_animator_set_parent.argtypes = [C.c_void_p, C.c_void_p]
    l.svb_animator_computed_state.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_double, C.c_void_p, C.c_void_p]
    l.svb_animator_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    l.svb_launch_timing.argtypes = [C.c_void_p, C.c_int]
    l.svb_launch_timing_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    l.svb_table_cache.argtypes = [C.c_void_p, C.c_int]
    l.svb_host_timing_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    l.svb_host_timing_read2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.svb_selftest_unorm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    l.svb_scale_convert_picture.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p]
    l.svb_scale_filter_table.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return l


lib = _load()


def _check(status):
    if status != 0:
        raise ComputeError(status, lib.svb_last_error().decode())


def _b(s):
    return None if s is None else s.encode()


def _f(arr, n):
    if arr is None:
        return None
    a = np.ascontiguousarray(np.asarray(arr, dtype=np.float32).reshape(-1))
    assert a.size == n
    return (C.c_float * n)(*a.tolist())


def available_compute_devices():
    return lib.svb_available_compute_devices()


def kernel_launch_count():
    return int(lib.svb_kernel_launch_count())


def kernel_module_image():
    """bytes of the sm_100a cubin a host can cuModuleLoadData itself (INTEGRATION.md level 1)."""
    p, n = C.c_void_p(), C.c_size_t()
    _check(lib.svb_kernel_module_image(C.byref(p), C.byref(n)))
    return C.string_at(p.value, n.value)


def default_compute_kernel_from_string(name):
    k = C.c_int()
    _check(lib.svb_default_compute_kernel_from_string(name.encode(), C.byref(k)))
    return k.value


class ComputeContext:
    def __init__(self, handle):
        self._h = handle

    @property
    def sm_count(self):
        return lib.svb_context_sm_count(self._h)

    @property
    def device_index(self):
        return lib.svb_context_device_index(self._h)

    def sharing(self):
        """createComputeContext(sharing:)"""
        h = C.c_void_p()
        _check(lib.svb_create_compute_context_sharing(self._h, C.byref(h)))
        return ComputeContext(h)

    def synchronize(self):
        """beginComputePass + endComputePass(ctx, true)"""
        _check(lib.svb_begin_compute_pass(self._h))
        _check(lib.svb_end_compute_pass(self._h, 1))

    def build_compute_kernel(self, name, image=None):
        buf = None if image is None else C.create_string_buffer(image, len(image))
        _check(lib.svb_build_compute_kernel(self._h, name.encode(), buf))

    def build_compute_kernel_from_source(self, name, source):
        """buildComputeKernel(_:name:source:): CUDA C source through NVRTC (sm_100a, --fmad=false) into the context's library."""
        _check(lib.svb_build_compute_kernel_from_source(self._h, name.encode(), source.encode()))

    def selftest_unorm(self):
        a, b = np.zeros(256, np.float32), np.zeros(256, np.float32)
        _check(lib.svb_selftest_unorm(self._h, a.ctypes.data, b.ctypes.data))
        return a, b

    def launch_timing(self, enable=True):
        _check(lib.svb_launch_timing(self._h, int(enable)))

    def table_cache(self, enable=True):
        """Keep a batch's coordinate tables while its geometry does not change (default on); off: the table pre-pass runs for every launch."""
        _check(lib.svb_table_cache(self._h, int(enable)))

    def launch_timing_read(self):
        """(total device ms, launches) of the fused kernels since timing was enabled."""
        ms, n = C.c_double(), C.c_ulonglong()
        _check(lib.svb_launch_timing_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def host_timing_read(self, with_wait=False):
        """(total host ms, calls[, back-pressure ms]) spent inside the library's fused compose calls since launch timing was enabled; the
        total does not count the time blocked on a GPU that is eight launches behind (that is the back-pressure figure)."""
        ms, n, w = C.c_double(), C.c_ulonglong(), C.c_double()
        _check(lib.svb_host_timing_read2(self._h, C.byref(ms), C.byref(n), C.byref(w)))
        return (ms.value, n.value, w.value) if with_wait else (ms.value, n.value)

    def close(self):
        if self._h:
            lib.svb_destroy_compute_context(self._h)
            self._h = None


def make_compute_context(device_index=0):
    """makeComputeContext(forType: .GPU) -- device_index is the one extension (upstream takes devices.first)."""
    h = C.c_void_p()
    _check(lib.svb_make_compute_context(DEVICE_GPU, device_index, C.byref(h)))
    return ComputeContext(h)


class PictureSample:
    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        try:
            if self._h:
                lib.svb_picture_release(self._h)
                self._h = None
        except Exception:
            pass

    def info(self):
        i = _PictureInfo()
        _check(lib.svb_picture_info_get(self._h, C.byref(i)))
        return i

    def pixel_format(self):
        return self.info().pixel_format

    def buffer_type(self):
        return self.info().buffer_type

    def z_index(self):
        return self.info().z_index

    def device_planes(self):
        i = self.info()
        return [i.planes[k].device for k in range(i.plane_count)]

    def host_planes(self):
        """numpy views (height x stride) over the CPU buffers; keep this sample alive while using them."""
        i = self.info()
        out = []
        for k in range(i.plane_count):
            p = i.planes[k]
            if not p.host:
                raise ComputeError(12, "sample has no CPU buffer")
            buf = (C.c_uint8 * p.size).from_address(p.host)
            a = np.frombuffer(buf, dtype=np.uint8).reshape(int(p.height), p.stride)
            out.append(a)
        return out

    def host_bytes(self):
        """All planes concatenated (the contiguous layout of createPictureSample)."""
        return np.concatenate([p.reshape(-1) for p in self.host_planes()])

    def set_host_bytes(self, data):
        data = np.asarray(data, dtype=np.uint8).reshape(-1)
        off = 0
        for p in self.host_planes():
            n = p.size
            p.reshape(-1)[:] = data[off : off + n]
            off += n
        assert off == data.size, (off, data.size)

    def with_(self, matrix=None, texture_matrix=None, border_matrix=None, fill_color=None, opacity=None, revision=None, asset_id=None):
        """PictureSample(other, matrix:textureMatrix:borderMatrix:fillColor:opacity:revision:assetId:)."""
        h = C.c_void_p()
        op = None if opacity is None else C.c_float(opacity)
        _check(lib.svb_picture_with(self._h, _f(matrix, 16), _f(texture_matrix, 16), _f(border_matrix, 16), _f(fill_color, 4),
                                    C.cast(C.pointer(op), C.c_void_p) if op is not None else None, _b(revision), _b(asset_id), C.byref(h)))
        return PictureSample(h)

    def revision(self):
        return lib.svb_picture_revision(self._h).decode()

    def device_frame(self):
        """svb_picture_device_frame: where the planes lie, for a consumer that reads them on the device."""
        f = DeviceFrame()
        _check(lib.svb_picture_device_frame(self._h, C.byref(f)))
        return f

    def consumed_on(self, stream):
        """svb_picture_consumed_on: the consumer's reads are queued on `stream` (a CUstream handle); writers wait for that point."""
        _check(lib.svb_picture_consumed_on(self._h, C.c_void_p(stream)))

    def gather(self, ctx, wait=True):
        """svb_gather_picture: this GPU sample copied onto ctx's device (peer copy); itself when it already lives there."""
        h = C.c_void_p()
        _check(lib.svb_gather_picture(ctx._h, self._h, 1 if wait else 0, C.byref(h)))
        return PictureSample(h)

    def asset_id(self):
        return lib.svb_picture_asset_id(self._h).decode()

    def animate(self, canvas, pos, size, rotation=0.0, border=(0, 0, 0, 0), aspect=0, tex_offset=(0, 0), fill=None, transparency=0.0,
                top_left=True, parent_opacity=1.0, revision=None):
        """PictureAnimator.impl in native code (svb_animate_picture)."""
        st = element_state(pos, size, rotation=rotation, border=border, aspect=aspect, tex_offset=tex_offset, fill=fill, transparency=transparency,
                           top_left=top_left)
        h = C.c_void_p()
        _check(lib.svb_animate_picture(self._h, float(canvas[0]), float(canvas[1]), C.byref(st), float(parent_opacity), _b(revision), C.byref(h)))
        return PictureSample(h)

    def upload(self, ctx, max_planes=3, retain_cpu_buffer=True, wait=True):
        """uploadComputePicture; wait=False queues the copies and returns at once (keep the source bytes untouched until result.wait())"""
        h = C.c_void_p()
        _check(lib.svb_upload_compute_picture(ctx._h, self._h, max_planes, int(retain_cpu_buffer), int(wait), C.byref(h)))
        return PictureSample(h)

    def download(self, ctx, retain_gpu_buffer=False, wait=True):
        """downloadComputePicture"""
        h = C.c_void_p()
        _check(lib.svb_download_compute_picture(ctx._h, self._h, int(retain_gpu_buffer), int(wait), C.byref(h)))
        return PictureSample(h)

    def wait(self):
        _check(lib.svb_picture_wait(self._h))

    def same_sample(self, other):
        """True when both handles refer to the same immutable sample (a barrier's pass-through)."""
        return lib.svb_picture_identity(self._h) == lib.svb_picture_identity(other._h) != 0

    def barrier_upload(self, ctx, retain_cpu_buffer=True, wait=True):
        """GPUBarrierUpload (compute.swift:175-198): (sample, None) or (None, EventError)."""
        return self._barrier(lib.svb_gpu_barrier_upload, ctx, retain_cpu_buffer, wait)

    def barrier_download(self, ctx, retain_gpu_buffer=True, wait=True):
        """GPUBarrierDownload (compute.swift:232-255): (sample, None) or (None, EventError)."""
        return self._barrier(lib.svb_gpu_barrier_download, ctx, retain_gpu_buffer, wait)

    def _barrier(self, fn, ctx, retain, wait):
        h, err = C.c_void_p(), EventError()
        rc = fn(ctx._h, self._h, int(retain), int(wait), C.byref(h), C.byref(err))
        return (PictureSample(h), None) if rc == 0 else (None, err)

    def scale_convert(self, ctx, width, height, pixel_format=BGRA, filter=FILTER_BILINEAR, wait=True):
        """svb_scale_convert_picture (ours; no upstream counterpart): a GPU NV12 / P010 sample -> a new GPU BGRA sample."""
        h = C.c_void_p()
        _check(lib.svb_scale_convert_picture(ctx._h, self._h, width, height, pixel_format, filter, int(wait), C.byref(h)))
        return PictureSample(h)


def element_state(pos, size, rotation=0.0, border=(0, 0, 0, 0), aspect=0, tex_offset=(0, 0), fill=None, transparency=0.0, top_left=True,
                  hidden=False, anchors=0):
    """An svb_element_state (Proto/Composition.proto ElementState, the fields the picture animator reads)."""
    st = ElementState()
    st.pic_pos[:] = [float(v) for v in (tuple(pos) + (0.0,))[:3]]
    st.size[:] = [float(size[0]), float(size[1])]
    st.texture_offset[:] = [float(tex_offset[0]), float(tex_offset[1])]
    st.border_size[:] = [float(v) for v in border]
    if fill is not None:
        st.fill_color[:] = [float(v) for v in fill]
        st.has_fill_color = 1
    st.rotation, st.transparency = float(rotation), float(transparency)
    st.pic_aspect, st.pic_origin = int(aspect), 1 if top_left else 0
    st.hidden, st.parent_anchors = int(bool(hidden)), int(anchors)
    return st


def compute_picture_state(sample_size, current, next=None, pct=None, anchors=ANCHOR_TOP_LEFT, parent_matrix=None, initial_parent_matrix=None):
    """computePictureState (svb_compute_picture_state): the unprojected ComputedPictureState of an element."""
    out = ComputedPictureState()
    p = None if pct is None else C.c_float(pct)
    _check(lib.svb_compute_picture_state(float(sample_size[0]), float(sample_size[1]), C.byref(current), C.byref(next) if next is not None else None,
                                         C.cast(C.pointer(p), C.c_void_p) if p is not None else None, int(anchors), _f(parent_matrix, 16),
                                         _f(initial_parent_matrix, 16), C.byref(out)))
    return out


class PictureAnimator:
    """PictureAnimator (animator.pic.swift:29-139) in native code; `now` is passed in seconds where upstream reads its Clock."""

    def __init__(self, canvas, parent=None, parent_anchors=ANCHOR_TOP_LEFT):
        self._h = C.c_void_p()
        self._parent = parent  # upstream holds the parent weakly; the caller keeps it alive, as here
        _check(lib.svb_animator_create(float(canvas[0]), float(canvas[1]), parent._h if parent is not None else None, int(parent_anchors), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib.svb_animator_destroy(self._h)
            self._h = None

    @property
    def revision(self):
        return lib.svb_animator_revision(self._h).decode()

    def set_state(self, state, duration=0.0, now=0.0):
        _check(lib.svb_animator_set_state(self._h, C.byref(state), float(duration), float(now)))

    def set_parent(self, parent):
        self._parent = parent
        _check(lib.svb_animator_set_parent(self._h, parent._h if parent is not None else None))

    def computed_state(self, sample_size, now=0.0, parent_state=None):
        out = ComputedPictureState()
        _check(lib.svb_animator_computed_state(self._h, float(sample_size[0]), float(sample_size[1]), float(now),
                                               C.byref(parent_state) if parent_state is not None else None, C.byref(out)))
        return out

    def apply(self, pict, now=0.0):
        """impl(): the re-issued sample, or None where upstream emits nothing (hidden / no state)."""
        h = C.c_void_p()
        _check(lib.svb_animator_apply(self._h, pict._h, float(now), C.byref(h)))
        return PictureSample(h) if h else None


def upload_many(ctx, picts, max_planes=3, retain_cpu_buffer=True, wait=True):
    """svb_upload_compute_pictures: a batch upload; neighbours in page-locked host memory travel as one copy."""
    n = len(picts)
    arr = (C.c_void_p * max(n, 1))(*[p._h for p in picts])
    outs = (C.c_void_p * max(n, 1))()
    _check(lib.svb_upload_compute_pictures(ctx._h, arr, n, int(max_planes), int(retain_cpu_buffer), int(wait), outs))
    return [PictureSample(C.c_void_p(outs[i])) for i in range(n)]


def scale_filter_table(filter, src_n, dst_n):
    """(first[dst_n] int32, weights[dst_n, taps] float32): one axis of the convert+scale operator's resize (host only)."""
    taps = C.c_int()
    _check(lib.svb_scale_filter_table(filter, src_n, dst_n, None, None, 0, C.byref(taps)))
    first = np.zeros(dst_n, np.int32)
    w = np.zeros((dst_n, taps.value), np.float32)
    _check(lib.svb_scale_filter_table(filter, src_n, dst_n, first.ctypes.data, w.ctypes.data, w.size, C.byref(taps)))
    return first, w


def create_picture_sample(width, height, pixel_format, asset_id="", workspace_id="", pinned_from=None):
    """createPictureSample; pinned_from=ctx makes the CPU buffer page-locked."""
    h = C.c_void_p()
    _check(lib.svb_create_picture_sample(width, height, pixel_format, asset_id.encode(), workspace_id.encode(),
                                         pinned_from._h if pinned_from else None, C.byref(h)))
    return PictureSample(h)


def picture_sample_from_planes(width, height, pixel_format, planes, asset_id="", workspace_id="", pinned_from=None):
    """CPU sample over numpy planes (2-D uint8 arrays whose row length in bytes is the stride) -- decoder-style strides."""
    arrs = [np.ascontiguousarray(p, dtype=np.uint8) for p in planes]
    ptrs = (C.c_void_p * 3)(*[a.ctypes.data for a in arrs] + [None] * (3 - len(arrs)))
    strides = (C.c_int32 * 3)(*[a.shape[1] for a in arrs] + [0] * (3 - len(arrs)))
    h = C.c_void_p()
    _check(lib.svb_picture_sample_from_planes(width, height, pixel_format, ptrs, strides, len(arrs), asset_id.encode(), workspace_id.encode(),
                                              pinned_from._h if pinned_from else None, C.byref(h)))
    return PictureSample(h)


def make_image_uniforms(image, target):
    u = ImageUniforms()
    _check(lib.svb_make_image_uniforms(image._h, target._h, C.byref(u)))
    return u


def apply_compute_image(ctx, image, target, kernel):
    _check(lib.svb_apply_compute_image(ctx._h, image._h, target._h, kernel))


def run_compute_kernel(ctx, images, target, kernel, uniforms=None, custom_name=None, blends=False):
    arr = (C.c_void_p * max(len(images), 1))(*[i._h for i in images])
    _check(lib.svb_run_compute_kernel(ctx._h, arr, len(images), target._h, kernel, _b(custom_name), 3,
                                      C.cast(C.pointer(uniforms), C.c_void_p) if uniforms is not None else None,
                                      C.sizeof(uniforms) if uniforms is not None else 0, int(blends)))


def compose(ctx, target, layers, uniforms, mode=MixMode.FUSED):
    """clear + fold `layers` (z-sorted) into `target` with explicit ImageUniforms (any 236-byte ctypes struct)."""
    n = len(layers)
    arr = (C.c_void_p * max(n, 1))(*[l._h for l in layers])
    us = (ImageUniforms * max(n, 1))()
    for i, u in enumerate(uniforms):
        C.memmove(C.byref(us[i]), C.byref(u), 236)
    _check(lib.svb_compose(ctx._h, target._h, arr, us, n, mode))


class VideoMixer:
    def __init__(self, ctx, width, height, pixel_format=NV12, asset_id=None, workspace_id="", frame_duration=1000, timescale=30000,
                 epoch=0):
        h = C.c_void_p()
        _check(lib.svb_video_mixer_create(ctx._h if ctx else None, width, height, pixel_format, _b(asset_id), workspace_id.encode(),
                                          frame_duration, timescale, epoch, C.byref(h)))
        self._h = h

    def asset_id(self):
        return lib.svb_video_mixer_asset_id(self._h).decode()

    def set_mode(self, mode):
        _check(lib.svb_video_mixer_set_mode(self._h, mode))

    def push(self, pict):
        stored = C.c_int()
        _check(lib.svb_video_mixer_push(self._h, pict._h, C.byref(stored)))
        return bool(stored.value)

    def push_many(self, picts):
        arr = (C.c_void_p * len(picts))(*[p._h for p in picts])
        _check(lib.svb_video_mixer_push_many(self._h, arr, len(picts)))

    def mix(self, time, wait=True):
        h = C.c_void_p()
        _check(lib.svb_video_mixer_mix(self._h, time, int(wait), C.byref(h)))
        return PictureSample(h)

    @staticmethod
    def mix_many(mixers, time, wait=True):
        n = len(mixers)
        arr = (C.c_void_p * n)(*[m._h for m in mixers])
        outs = (C.c_void_p * n)()
        _check(lib.svb_video_mixer_mix_many(arr, n, time, int(wait), outs))
        return [PictureSample(C.c_void_p(outs[i])) for i in range(n)]

    @staticmethod
    def tick_many(mixers, layers_per_mixer, time, wait=False):
        """svb_video_mixer_tick_many: upload + push + mix + download of one tick of several mixers in ONE call across the boundary.
        layers_per_mixer: a list of lists of (CPU or GPU) samples; returns the emitted frames as CPU samples (wait() before reading)."""
        n = len(mixers)
        flat = [p for ls in layers_per_mixer for p in ls]
        arr = (C.c_void_p * n)(*[m._h for m in mixers])
        larr = (C.c_void_p * max(1, len(flat)))(*[p._h for p in flat])
        counts = (C.c_int * n)(*[len(ls) for ls in layers_per_mixer])
        outs = (C.c_void_p * n)()
        _check(lib.svb_video_mixer_tick_many(arr, n, larr, counts, time, int(wait), outs))
        return [PictureSample(C.c_void_p(outs[i])) for i in range(n)]

    def close(self):
        if self._h:
            lib.svb_video_mixer_destroy(self._h)
            self._h = None


class Timer:
    """CUDA-event stopwatch over the context's three streams."""

    def __init__(self, ctx):
        h = C.c_void_p()
        _check(lib.svb_timer_create(ctx._h, C.byref(h)))
        self._h = h

    def start(self):
        _check(lib.svb_timer_start(self._h))

    def stop(self):
        _check(lib.svb_timer_stop(self._h))

    def elapsed_ms(self):
        ms = C.c_float()
        _check(lib.svb_timer_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def close(self):
        if self._h:
            lib.svb_timer_destroy(self._h)
            self._h = None
