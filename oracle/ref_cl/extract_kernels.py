#!/usr/bin/env python3
"""Extract the reference's OpenCL image-kernel text and wrap it for compilation as C++.

TEST INFRASTRUCTURE ONLY.  Reads  <reference>/Sources/SwiftVideo/kernels.cl.swift  where it lies,
writes ONLY into oracle/_ref/ (git-ignored): the reference's source is never copied into the repo.

The kernel bodies are emitted verbatim except for one mechanical, semantics-free rewrite that C++
needs: OpenCL vector literals  `(float4)(a, b, c, d)`  become constructor calls  `float4(a, b, c, d)`
(same for float2 / int2 / float3).  Everything else OpenCL-specific is supplied by cl_shim.hpp.
"""
import re
import sys
from pathlib import Path


def extract(swift_text: str):
    prelude = re.search(r'let kOpenCLKernelMatrixFuncs =\s*"""\n(.*?)\n"""', swift_text, re.S).group(1)
    kernels = {}
    for m in re.finditer(r'case (\w+) =\s*\n\s*"""\n(.*?)\n\s*"""', swift_text, re.S):
        name, body = m.group(1), m.group(2)
        if name.startswith("img_"):
            kernels[name] = body
    return prelude, kernels


def cxxify(src: str) -> str:
    return re.sub(r"\((float4|float3|float2|int2)\)\(", r"\1(", src)


def signature(name: str, body: str):
    m = re.search(r"__kernel\s+void\s+" + name + r"\s*\((.*?)\)\s*\{", body, re.S)
    params = [p.strip() for p in m.group(1).split(",")]
    nimg = sum(1 for p in params if "image2d_t" in p)
    has_uniforms = any("ImageUniforms" in p for p in params)
    assert nimg + (1 if has_uniforms else 0) == len(params), (name, params)
    return nimg, has_uniforms


def main():
    ref_root = Path(sys.argv[1])
    out_dir = Path(sys.argv[2])
    text = (ref_root / "Sources/SwiftVideo/kernels.cl.swift").read_text()
    prelude, kernels = extract(text)
    out = ["// GENERATED from the reference's kernels.cl.swift by oracle/ref_cl/extract_kernels.py -- do not commit.\n"]
    table = []
    for name, body in kernels.items():
        nimg, has_u = signature(name, body)
        out.append(f"namespace k_{name} {{\n{cxxify(prelude)}\n{cxxify(body)}\n}}\n")
        args = ", ".join(f"im[{i}]" for i in range(nimg))
        if has_u:
            args += f", (const k_{name}::ImageUniforms*)u"
        out.append(
            f"static void run_{name}(ClImage** im, const void* u, int W, int H, int y0, int y1) {{\n"
            f"  (void)u; g_gsize[0] = W; g_gsize[1] = H;\n"
            f"  for (int y = y0; y < y1; ++y) for (int x = 0; x < W; ++x) {{\n"
            f"    g_gid[0] = x; g_gid[1] = y; k_{name}::{name}({args});\n"
            f"  }}\n}}\n"
        )
        table.append(f'  {{"{name}", run_{name}, {nimg}, {int(has_u)}}},')
    out.append("static const RefKernel kRefKernels[] = {\n" + "\n".join(table) + "\n};\n")
    out_dir.mkdir(parents=True, exist_ok=True)
    (out_dir / "kernels_gen.inc").write_text("".join(out))
    print(f"extracted {len(kernels)} kernels: {' '.join(kernels)}")


if __name__ == "__main__":
    main()
