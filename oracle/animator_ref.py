"""TEST INFRASTRUCTURE ONLY -- a numpy float32 restatement of the reference's picture animator, used to check the native
animator (swiftvideo_b200/csrc/animator.cpp).  Only tests/ may import it.

Follows /root/reference/Sources/SwiftVideo/animator.pic.swift step by step, vertex list and all, so that the native
edge-based formulation is checked against the reference's own case analysis:
  computePositionSize  :149-191      computeElementState :193-205      computeTextureMatrix :207-227
  computePictureState  :229-272      interpolate         :278-304      Matrix4(ortho)       :326-333
Parity unpinned: the reference has no animator test or golden vector; this file is the restated definition.
Matrices are returned as 16 floats in VectorMath memory order (m11 m12 m13 m14 m21 ...), like PictureSample carries them.
"""
import numpy as np

F = np.float32
TL, TR, BL, BR = 0, 1, 2, 3  # Proto/Composition.proto:31-36


def _v(*a):
    return np.array(a, dtype=F)


def position_size(base_pos, base_size, parent_pos, delta, anchors):
    """:149-191 -- three vertices (top-left, top-right, bottom-left) pushed around by the anchor set."""
    rel = _v(base_pos[0] + parent_pos[0], base_pos[1] + parent_pos[1], base_pos[2] + F(0))
    v = [rel.copy(), rel + _v(base_size[0], 0, 0), rel + _v(0, base_size[1], 0)]
    a = set(anchors)
    dx, dy = _v(delta[0], 0, 0), _v(0, delta[1], 0)
    if BR in a:
        v = [p + _v(*delta) for p in v]
        if BL in a:
            v[0][0] = rel[0]
            v[2][0] = rel[0]
        if TR in a:
            v[0][1] = rel[1]
            v[1][1] = rel[1]
        if TL in a:
            v[0] = rel.copy()
            v[1] = rel + _v(base_size[0], 0, 0) + dx
            v[2] = rel + _v(0, base_size[1], 0) + dy
    elif TR in a:
        v[1] = v[1] + dx
        if TL not in a and BL not in a:
            v[0] = v[0] + dx
            v[2] = v[2] + dx
        elif BL in a:
            v[2] = v[2] + dy
    elif BL in a:
        v[2] = v[2] + dy
        if TL not in a:
            v[1] = v[1] + dy
            v[0] = v[0] + dy
    return v[0], _v(v[1][0] - v[0][0], v[2][1] - v[0][1], 1.0)


def lerp(a, b, t):
    a, b = np.asarray(a, dtype=F), np.asarray(b, dtype=F)
    return a + (b - a) * F(t)  # :278-304


def element_state(cur, nxt, pct):
    """:193-205; states are dicts with the ElementState field names (fill None = unset = zeros, :335-342)."""
    z4 = (0.0, 0.0, 0.0, 0.0)
    out = dict(nxt)
    for k in ("pos", "size", "tex_offset", "rotation", "transparency", "border"):
        out[k] = lerp(cur[k], nxt[k], pct)
    out["fill"] = lerp(cur.get("fill") or z4, nxt.get("fill") or z4, pct)
    return out


def _translate(p):
    m = np.eye(4, dtype=F)
    m[:3, 3] = p
    return m


def _rotz(a):
    c, s = F(np.cos(F(a))), F(np.sin(F(a)))
    m = np.eye(4, dtype=F)
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
    return m


def _scale(s):
    return np.diag(np.array([s[0], s[1], s[2], 1.0], dtype=F))


def _mem(m):
    return np.ascontiguousarray(m.astype(F).T).reshape(-1)


def texture_matrix(sample_size, geom, offset, aspect):
    """:207-227; aspect 0 none, 1 fit, 2 fill"""
    if aspect == 0:
        return np.eye(4, dtype=F)
    orig, g = F(sample_size[0]) / F(sample_size[1]), F(geom[0]) / F(geom[1])
    if aspect == 1:
        sx = F(1) if orig > g else orig / g
        sy = F(1) if orig <= g else g / orig
    else:
        sx = F(1) if orig <= g else orig / g
        sy = F(1) if orig > g else g / orig
    return _translate(_v(F(offset[0]) + (F(1) - sx) / F(2), F(offset[1]) + (F(1) - sy) / F(2), 0)) @ _scale(_v(sx, sy, 1))


def _col_len(mem, c):
    return F(np.sqrt(F(mem[4 * c] * mem[4 * c] + mem[4 * c + 1] * mem[4 * c + 1])))


def picture_state(sample_size, cur, nxt=None, pct=None, anchors=(TL,), parent=None, initial_parent=None):
    """:229-272 -> dict(matrix, texture_matrix, border_matrix, fill, opacity); parent / initial_parent are 16-float matrices."""
    st = element_state(cur, nxt, pct) if (nxt is not None and pct is not None) else cur
    ppos, psize, isize = _v(0, 0, 0), _v(0, 0, 0), _v(0, 0, 0)
    if parent is not None:
        parent = np.asarray(parent, dtype=F)
        ppos = _v(parent[12], parent[13], parent[14])
        psize = _v(_col_len(parent, 0), _col_len(parent, 1), 0)
    if initial_parent is not None:
        ip = np.asarray(initial_parent, dtype=F)
        isize = _v(_col_len(ip, 0), _col_len(ip, 1), 0)
    delta = psize - isize
    size2 = np.asarray(st["size"], dtype=F)
    add = _v(0, 0, 0) if st.get("top_left", True) else -_v(size2[0] / F(2), size2[1] / F(2), 0)
    pos3 = np.asarray((tuple(st["pos"]) + (0.0,))[:3], dtype=F)
    rel, size = position_size(pos3, _v(size2[0], size2[1], 0), ppos, delta, anchors)
    pos = rel + add
    b = np.asarray(st["border"], dtype=F)
    bpos = pos - _v(b[0], b[1], 0)
    bsize = _v(b[0] + size[0] + b[2], b[1] + size[1] + b[3], 1)
    rot = _rotz(st["rotation"])
    fill = st.get("fill")
    return dict(matrix=_mem(_translate(pos) @ rot @ _scale(size)),
                texture_matrix=_mem(texture_matrix(sample_size, size, st["tex_offset"], st.get("aspect", 0))),
                border_matrix=_mem(_translate(bpos) @ rot @ _scale(bsize)),
                fill=np.asarray(fill if fill is not None else (0, 0, 0, 0), dtype=F), opacity=F(1) - F(st["transparency"]))


def ortho(canvas):
    """:326-333, as a standard (row, col) matrix"""
    return np.array([[F(2) / F(canvas[0]), 0, 0, -1], [0, F(2) / F(canvas[1]), 0, -1], [0, 0, 1, 1], [0, 0, 0, 1]], dtype=F)


def project(canvas, mem):
    """impl() :118-121: projection * matrix, both in memory order"""
    return _mem(ortho(canvas) @ np.asarray(mem, dtype=F).reshape(4, 4).T)
