"""Synthetic pictures and ImageUniforms for the parity tests (host-side helper, numpy only).

Uniform construction mirrors how the reference produces them -- PictureAnimator
(animator.pic.swift:107-128,207-272,326-333) builds  matrix = ortho(canvas) * T(pos) * Rz(rot) * S(size),
textureMatrix from the aspect mode, borderMatrix from the rect grown by borderSize, and
applyComputeImage (compute.swift:145-170) uploads inverse.transpose of each -- but pixel parity does
not depend on it: the uniforms are an INPUT of both the oracle and the CUDA path.
"""
import numpy as np

from oracle import oracle as O


def _ortho(canvas):
    cx, cy = canvas
    return np.array([[2.0 / cx, 0, 0, -1], [0, 2.0 / cy, 0, -1], [0, 0, 1, 1], [0, 0, 0, 1]], dtype=np.float64)


def _translate(x, y, z=0.0):
    m = np.eye(4)
    m[:3, 3] = (x, y, z)
    return m


def _rotz(a):
    c, s = np.cos(a), np.sin(a)
    m = np.eye(4)
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
    return m


def _scale(x, y, z=1.0):
    return np.diag([x, y, z, 1.0])


def _inv_rows(m):
    """Row-major float32[16] of M^-1 (what `inverse.transpose` looks like in the kernel's float4[4])."""
    inv = np.linalg.inv(m)
    inv[np.abs(inv) < 1e-12] = 0.0  # structural zeros stay exact zeros
    return inv.astype(np.float32).reshape(-1)


def texture_matrix(src_size, geom_size, aspect="none", tex_offset=(0.0, 0.0)):
    """computeTextureMatrix, animator.pic.swift:207-227."""
    if aspect == "none":
        return np.eye(4)
    orig = src_size[0] / src_size[1]
    geom = geom_size[0] / geom_size[1]
    if aspect == "fit":
        sx = 1.0 if orig > geom else orig / geom
        sy = 1.0 if orig <= geom else geom / orig
    elif aspect == "fill":
        sx = 1.0 if orig <= geom else orig / geom
        sy = 1.0 if orig > geom else geom / orig
    else:
        raise ValueError(aspect)
    return _translate(tex_offset[0] + (1 - sx) / 2, tex_offset[1] + (1 - sy) / 2) @ _scale(sx, sy)


def layer_uniforms(canvas, src_size, pos, size, *, rotation=0.0, z=0.0, opacity=1.0, fill=(0.0, 0.0, 0.0, 0.0),
                   border=(0.0, 0.0, 0.0, 0.0), aspect="none", tex_offset=(0.0, 0.0)):
    """ImageUniforms for one layer placed at pixel `pos` with pixel `size` on `canvas`."""
    proj = _ortho(canvas)
    m = proj @ _translate(pos[0], pos[1], z) @ _rotz(rotation) @ _scale(size[0], size[1])
    bl, bt, br, bb = border
    bm = proj @ _translate(pos[0] - bl, pos[1] - bt, z) @ _rotz(rotation) @ _scale(bl + size[0] + br, bt + size[1] + bb)
    tm = texture_matrix(src_size, size, aspect, tex_offset)
    u = O.Uniforms()
    u.transform[:] = _inv_rows(m).tolist()
    u.textureTx[:] = _inv_rows(tm).tolist()
    u.borderMatrix[:] = _inv_rows(bm).tolist()
    u.fillColor[:] = [float(np.float32(v)) for v in fill]
    u.inSize[:] = [float(src_size[0]), float(src_size[1])]
    u.outSize[:] = [float(canvas[0]), float(canvas[1])]
    u.opacity = float(np.float32(opacity))
    u.sampleTime = 0.0
    u.targetTime = 0.0
    return u


def random_image(fmt, w, h, seed, dist="uniform"):
    """dist 'uniform': u8 over [0,255] (worst case for rounding ties); 'ramp': smooth ramp +-3 noise in video range."""
    rng = np.random.default_rng(seed)
    img = O.Image(fmt, w, h)
    if dist == "uniform":
        img.data[:] = rng.integers(0, 256, size=img.nbytes, dtype=np.uint8)
    elif dist == "ramp":
        for i, (off, pw, ph, stride, nc) in enumerate(img.layout):
            yy, xx = np.mgrid[0:ph, 0 : pw * nc]
            base = 16 + (xx * 200 // max(pw * nc - 1, 1) + yy * 19 // max(ph - 1, 1)) % 220
            noise = rng.integers(-3, 4, size=base.shape)
            img.plane(i)[:, :] = np.clip(base + noise, 16, 240 if i else 235).astype(np.uint8)
    else:
        raise ValueError(dist)
    return img


# ---- BASELINE.json configs -------------------------------------------------------

def cfg_seed(cfg, stream, layer):
    return 1000 * cfg + 16 * stream + layer


def cfg2_scene(stream=0, dist="uniform"):
    """1920x1080 NV12 -> 1280x720 NV12 target, full canvas, opacity 1."""
    canvas = (1280, 720)
    src = random_image(O.NV12, 1920, 1080, cfg_seed(2, stream, 0), dist)
    u = layer_uniforms(canvas, (1920, 1080), (0, 0), canvas, z=1.0)
    return canvas, O.NV12, [src], [u]


def cfg34_geometry(nlayers):
    """(src_size, pos, dst_size, opacity) per layer for the 4K composites, z = layer index."""
    geo = [((3840, 2160), (0, 0), (3840, 2160), 1.0)]
    for k in range(1, nlayers):
        if nlayers <= 4:
            pos = ((k - 1) * 640, 270 + (k - 1) * 180)
        else:
            pos = (((k - 1) % 4) * 560, 135 + ((k - 1) // 4) * 990)
        geo.append(((1920, 1080), pos, (1600, 900), 0.5 + 0.05 * k))
    return geo


def cfg34_scene(nlayers, stream=0, dist="uniform", cfg=None):
    cfg = cfg or (3 if nlayers <= 4 else 4)
    canvas = (3840, 2160)
    layers, us = [], []
    for k, (ssz, pos, dsz, op) in enumerate(cfg34_geometry(nlayers)):
        layers.append(random_image(O.NV12, ssz[0], ssz[1], cfg_seed(cfg, stream, k), dist))
        us.append(layer_uniforms(canvas, ssz, pos, dsz, z=float(k + 1), opacity=op))
    return canvas, O.NV12, layers, us


# ---- small parity scenes ----------------------------------------------------------------------------------

class Case:
    def __init__(self, name, target_fmt, canvas, layers, uniforms):
        self.name, self.target_fmt, self.canvas, self.layers, self.uniforms = name, target_fmt, canvas, layers, uniforms


def _standard_stack(target_fmt, src_fmt, canvas, seed, dist="uniform"):
    """Four layers exercising: upscale, aspect-fit letterbox with fill, border + opacity, rotation."""
    cw, ch = canvas
    l0 = random_image(src_fmt, 48, 28, seed + 0, dist)
    l1 = random_image(src_fmt, 80, 60, seed + 1, dist)
    l2 = random_image(src_fmt, 40, 40, seed + 2, dist)
    l3 = random_image(src_fmt, 36, 20, seed + 3, dist)
    us = [
        layer_uniforms(canvas, (48, 28), (0, 0), canvas, z=1, opacity=1.0),
        layer_uniforms(canvas, (80, 60), (cw * 0.1, ch * 0.15), (cw * 0.55, ch * 0.5), z=2, opacity=0.8,
                       fill=(0.2, 0.6, 0.9, 0.7), aspect="fit"),
        layer_uniforms(canvas, (40, 40), (cw * 0.5, ch * 0.4), (cw * 0.4, ch * 0.45), z=3, opacity=0.6,
                       fill=(1.0, 0.3, 0.1, 1.0), border=(3, 2, 4, 5), aspect="fill"),
        layer_uniforms(canvas, (36, 20), (cw * 0.3, ch * 0.3), (cw * 0.35, ch * 0.3), z=4, opacity=0.9,
                       rotation=0.3, fill=(0.0, 1.0, 0.0, 0.5), border=(2, 2, 2, 2)),
    ]
    return [l0, l1, l2, l3], us


PAIRS = [(O.NV12, O.NV12), (O.Y420P, O.NV12), (O.Y420P, O.Y420P), (O.BGRA, O.NV12), (O.RGBA, O.NV12),
         (O.BGRA, O.Y420P), (O.RGBA, O.Y420P)]


def parity_cases(canvas=(96, 64)):
    cases = []
    for i, (sf, tf) in enumerate(PAIRS):
        for dist in ("uniform", "ramp"):
            layers, us = _standard_stack(tf, sf, canvas, 100 + 10 * i, dist)
            cases.append(Case(f"stack_{O.FORMAT_NAMES[sf]}_{O.FORMAT_NAMES[tf]}_{dist}", tf, canvas, layers, us))
    # mixed source formats in one fold, NV12 target (y420p + bgra + rgba + nv12)
    mixed_l, mixed_u = [], []
    for k, sf in enumerate((O.NV12, O.Y420P, O.BGRA, O.RGBA)):
        l, u = _standard_stack(O.NV12, sf, canvas, 300 + 10 * k)
        mixed_l.append(l[k])
        mixed_u.append(u[k])
    cases.append(Case("mixed_sources_nv12", O.NV12, canvas, mixed_l, mixed_u))
    # opacity / fill-alpha edge values, including out-of-range opacity (the API does not clamp it)
    for op in (0.0, 1.0, 0.5, 1.5, -0.25):
        for sf in (O.NV12, O.BGRA):
            src = random_image(sf, 64, 48, 400)
            u = layer_uniforms(canvas, (64, 48), (8, 6), (70, 50), z=1, opacity=op, fill=(0.9, 0.1, 0.4, 0.5),
                               border=(4, 4, 4, 4), aspect="fit")
            base = random_image(O.NV12, canvas[0], canvas[1], 401)
            ub = layer_uniforms(canvas, canvas, (0, 0), canvas, z=0, opacity=1.0)
            cases.append(Case(f"opacity_{op}_{O.FORMAT_NAMES[sf]}", O.NV12, canvas, [base, src], [ub, u]))
    # identity: same size, full canvas (still a half-pixel-shifted bilinear -- out_uv has no +0.5)
    src = random_image(O.NV12, canvas[0], canvas[1], 500)
    cases.append(Case("identity_nv12", O.NV12, canvas, [src], [layer_uniforms(canvas, canvas, (0, 0), canvas, z=1)]))
    # partially off-canvas, negative origin, big downscale
    src = random_image(O.Y420P, 160, 120, 510)
    cases.append(Case("offcanvas_y420p", O.Y420P, canvas, [src],
                      [layer_uniforms(canvas, (160, 120), (-20, -10), (60, 40), z=1, opacity=0.75)]))
    # clear only
    cases.append(Case("clear_only_nv12", O.NV12, canvas, [], []))
    cases.append(Case("clear_only_y420p", O.Y420P, canvas, [], []))
    # smallest legal pictures
    src = random_image(O.NV12, 2, 2, 520)
    cases.append(Case("tiny_2x2", O.NV12, (2, 2), [src], [layer_uniforms((2, 2), (2, 2), (0, 0), (2, 2), z=1, opacity=0.5)]))
    # ragged sizes (not multiples of 16/32), 3 layers incl. rotation
    rc = (66, 38)
    layers, us = _standard_stack(O.NV12, O.NV12, rc, 530)
    cases.append(Case("ragged_66x38", O.NV12, rc, layers, us))
    layers, us = _standard_stack(O.Y420P, O.BGRA, rc, 540)
    cases.append(Case("ragged_66x38_bgra_y420p", O.Y420P, rc, layers, us))
    return cases


def run_case(lib, case, threads=0):
    target = O.Image(case.target_fmt, case.canvas[0], case.canvas[1])
    target.data[:] = 0xA5  # stale bytes: the clear pass must overwrite them
    rc = lib.mix(target, case.layers, case.uniforms, threads=threads)
    return rc, target


def golden_cases():
    """Cases stored in tests/golden/cases.npz (inputs + the bytes oracle/_ref produced): [(Case, expected bytes)]."""
    import ctypes as C
    from pathlib import Path
    z = np.load(Path(__file__).resolve().parent / "golden" / "cases.npz")
    out = []
    for name in z["names"]:
        name = str(name)
        tf, w, h, n = (int(v) for v in z[f"{name}/meta"])
        layers, us = [], []
        for i in range(n):
            lf, lw, lh = (int(v) for v in z[f"{name}/l{i}/meta"])
            layers.append(O.Image(lf, lw, lh, z[f"{name}/l{i}/data"]))
            u = O.Uniforms()
            C.memmove(C.byref(u), z[f"{name}/l{i}/uniforms"].tobytes(), 236)
            us.append(u)
        out.append((Case(name, tf, (w, h), layers, us), z[f"{name}/out"]))
    return out
