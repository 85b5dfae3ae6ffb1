"""Build libsvb200.so in-tree: kernels.cu -> sm_100a cubin (embedded) + the C++ host side + the C ABI.

    python -m swiftvideo_b200.build [--force]

nvcc cross-compiles without a GPU.  The cubin is also left beside the library (svb200_kernels.cubin) for
hosts that load it themselves with cuModuleLoad (level 1).
"""
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
BUILD = PKG / "build"
LIB = PKG / "libsvb200.so"
CUBIN = PKG / "svb200_kernels.cubin"
CUDA_HOME = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda"))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17"]
NVCC_FLAGS += os.environ.get("SVB_NVCC_DEFS", "").split()  # tuning builds, e.g. SVB_NVCC_DEFS="-DSVB_TILED_MIN_CTAS=3"
CXX_FLAGS = [f for f in os.environ.get("SVB_NVCC_DEFS", "").split() if f.startswith("-D")] + ["-O2", "-g1", "-fPIC", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wno-unused-function",
             "-fvisibility=hidden", f"-I{CUDA_HOME}/include"]


def _newer(target: Path, sources):
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _run(cmd):
    r = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build step failed: " + " ".join(str(c) for c in cmd) + "\n" + r.stdout + r.stderr)
    return r.stdout + r.stderr


def build(force=False, verbose=False):
    BUILD.mkdir(exist_ok=True)
    cu_sources = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [CSRC / "svb_desc.h"]
    log = ""
    if force or _newer(CUBIN, cu_sources):
        log += _run([CUDA_HOME / "bin" / "nvcc", "-cubin", *NVCC_FLAGS, "-Xptxas", "-v", "-o", CUBIN, CSRC / "kernels.cu"])
    host_sources = [CSRC / n for n in ("cu_driver.cpp", "compute.cpp", "mix_video.cpp", "scale.cpp", "animator.cpp", "abi.cpp")]
    headers = sorted(CSRC.glob("*.h")) + [PKG.parent / "include" / "svb200.h"]
    if force or _newer(LIB, host_sources + headers + [CUBIN]):
        blob = BUILD / "kernels_cubin.S"
        blob.write_text(
            '.section .rodata\n.global svb200_kernels_cubin\n.global svb200_kernels_cubin_len\n.balign 64\n'
            f'svb200_kernels_cubin:\n.incbin "{CUBIN}"\nsvb200_kernels_cubin_end:\n.byte 0\n.balign 8\n'
            'svb200_kernels_cubin_len:\n.quad svb200_kernels_cubin_end - svb200_kernels_cubin\n'
            '.section .note.GNU-stack,"",@progbits\n')
        log += _run(["g++", *CXX_FLAGS, "-shared", "-o", LIB, *host_sources, blob, "-ldl", "-lpthread"])
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
