"""swiftvideo_b200 -- B200-native VideoMixer compute path (NV12/YUV420P/BGRA convert, bilinear scale, N-layer
alpha composite) behind SwiftVideo's ComputeContext / PictureSample / VideoMixer operator surface.

The product is libsvb200.so (C ABI: include/svb200.h; sources: swiftvideo_b200/csrc).  This package is the thin
ctypes binding the tests and bench.py drive it through (swiftvideo_b200.api, re-exported here on first use so
that `python -m swiftvideo_b200.build` can run before the library exists).  There is no CPU fallback: if the
library has not been built, the binding raises ImportError; if no B200 is present, every compute call raises
ComputeError(deviceNotAvailable).
"""
import importlib

_API = ("BUFFER_CPU", "BUFFER_GPU", "EventError", "BGRA", "NV12", "RGBA", "Y420P", "P010", "NV21", "Y422P", "Y444P", "YUVS", "ZVUY", "FILTER_BILINEAR", "FILTER_LANCZOS3", "scale_filter_table", "ComputeContext", "ComputeError", "ImageUniforms", "MixMode", "PictureSample", "Timer",
        "VideoMixer", "PictureAnimator", "DeviceFrame", "ElementState", "ComputedPictureState", "element_state", "compute_picture_state",
        "ANCHOR_TOP_LEFT", "ANCHOR_TOP_RIGHT", "ANCHOR_BOTTOM_LEFT", "ANCHOR_BOTTOM_RIGHT", "available_compute_devices", "compose", "create_picture_sample", "upload_many", "kernel_launch_count",
        "kernel_module_image", "lib", "make_compute_context")


def __getattr__(name):
    if name in _API or name in ("api", "animator"):
        mod = importlib.import_module(".api" if name != "animator" else ".animator", __name__)
        return mod if name in ("api", "animator") else getattr(mod, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
