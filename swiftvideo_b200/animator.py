"""Host-side helper: the matrices PictureAnimator hands to the mixer, in VectorMath memory order.

Mirrors /root/reference/Sources/SwiftVideo/animator.pic.swift:
  :107-128  impl():  matrix = ortho(canvas) * state.matrix, borderMatrix = ortho(canvas) * state.borderMatrix
  :207-227  computeTextureMatrix (aspect fit / fill)
  :229-272  computePictureState:  T(pos) * Rz(rotation) * S(size);  border rect grown by borderSize (l, t, r, b)
  :326-333  Matrix4(ortho): scale (2/cx, 2/cy, 1), translate (-1, -1), m43 = 1   => zIndex = round(pos.z + 1)
Used by the tests and bench.py to place layers; the pixel path itself only ever sees the resulting matrices.
"""
import numpy as np

F = np.float32


def _mem(std):
    """standard (row, col) 4x4 -> 16 floats in VectorMath order m11,m12,m13,m14,m21,... (column by column)."""
    return np.ascontiguousarray(std.astype(F).T).reshape(-1)


def ortho(canvas):
    cx, cy = canvas
    return np.array([[F(2) / F(cx), 0, 0, -1], [0, F(2) / F(cy), 0, -1], [0, 0, 1, 1], [0, 0, 0, 1]], dtype=F)


def translation(x, y, z=0.0):
    m = np.eye(4, dtype=F)
    m[:3, 3] = (x, y, z)
    return m


def rotation_z(a):
    m = np.eye(4, dtype=F)
    c, s = F(np.cos(a)), F(np.sin(a))
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
    return m


def scale(x, y, z=1.0):
    return np.diag(np.array([x, y, z, 1.0], dtype=F))


def texture_matrix(src_size, geom_size, aspect="none", tex_offset=(0.0, 0.0)):
    if aspect == "none":
        return np.eye(4, dtype=F)
    orig = F(src_size[0]) / F(src_size[1])
    geom = F(geom_size[0]) / F(geom_size[1])
    if aspect == "fit":
        sx = F(1) if orig > geom else orig / geom
        sy = F(1) if orig <= geom else geom / orig
    elif aspect == "fill":
        sx = F(1) if orig <= geom else orig / geom
        sy = F(1) if orig > geom else geom / orig
    else:
        raise ValueError(aspect)
    return translation(F(tex_offset[0]) + (F(1) - sx) / F(2), F(tex_offset[1]) + (F(1) - sy) / F(2)) @ scale(sx, sy)


def picture_state(canvas, src_size, pos, size, rotation=0.0, z=0.0, border=(0.0, 0.0, 0.0, 0.0), aspect="none", tex_offset=(0.0, 0.0)):
    """(matrix, textureMatrix, borderMatrix) as 16-float arrays for PictureSample.with_()."""
    proj = ortho(canvas)
    rot = rotation_z(rotation) if rotation else np.eye(4, dtype=F)
    m = proj @ translation(pos[0], pos[1], z) @ rot @ scale(size[0], size[1])
    bl, bt, br, bb = border
    bm = proj @ translation(pos[0] - bl, pos[1] - bt, z) @ rot @ scale(bl + size[0] + br, bt + size[1] + bb)
    tm = texture_matrix(src_size, size, aspect, tex_offset)
    return _mem(m), _mem(tm), _mem(bm)
