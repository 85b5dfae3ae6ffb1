"""Host logic, no GPU: the native PictureAnimator (svb_compute_picture_state / svb_animator_*) against oracle/animator_ref.py,
the vertex-by-vertex restatement of animator.pic.swift:149-272, and the state machine of setState / impl (:54-128)."""
import itertools

import numpy as np
import pytest

import swiftvideo_b200 as sv
from oracle import animator_ref as R
from swiftvideo_b200 import api

TOL = dict(rtol=2e-6, atol=2e-6)


def _native_state(d):
    return sv.element_state(d["pos"], d["size"], rotation=d["rotation"], border=d["border"], aspect=d.get("aspect", 0),
                            tex_offset=d["tex_offset"], fill=d.get("fill"), transparency=d["transparency"], top_left=d.get("top_left", True),
                            hidden=d.get("hidden", False), anchors=d.get("anchors", 0))


def _rand_state(rng, **kw):
    d = dict(pos=tuple(rng.uniform(-50, 400, 2).round(2)) + (float(rng.integers(0, 4)),), size=tuple(rng.uniform(20, 500, 2).round(2)),
             rotation=float(rng.choice([0.0, 0.0, 0.3, -1.1])), border=tuple(rng.uniform(0, 9, 4).round(1)),
             tex_offset=tuple(rng.uniform(-0.2, 0.2, 2).round(3)), transparency=float(rng.uniform(0, 0.9)),
             aspect=int(rng.integers(0, 3)), top_left=bool(rng.integers(0, 2)), fill=tuple(rng.uniform(0, 1, 4)) if rng.integers(0, 2) else None)
    d.update(kw)
    return d


def _parent_matrix(rng):
    return R.picture_state((64, 64), _rand_state(rng, rotation=float(rng.choice([0.0, 0.5])), aspect=0))["matrix"]


def _same(cs, want):
    assert np.allclose(np.array(cs.matrix[:]), want["matrix"], **TOL)
    assert np.allclose(np.array(cs.texture_matrix[:]), want["texture_matrix"], **TOL)
    assert np.allclose(np.array(cs.border_matrix[:]), want["border_matrix"], **TOL)
    assert np.allclose(np.array(cs.fill_color[:]), want["fill"], **TOL) and abs(cs.opacity - want["opacity"]) < 1e-6


def test_every_anchor_set_against_the_vertex_restatement():
    """computePositionSize :149-191: all 16 anchor subsets, with and without an initial parent state; rotation 0 is exact."""
    rng = np.random.default_rng(11)
    for bits in range(16):
        anchors = [a for a in range(4) if bits >> a & 1]
        for trial in range(6):
            st = _rand_state(rng)
            parent = _parent_matrix(rng)
            initial = _parent_matrix(rng) if trial % 2 else None
            want = R.picture_state((640, 360), st, anchors=anchors or [R.TL], parent=parent, initial_parent=initial)
            got = sv.compute_picture_state((640, 360), _native_state(st), anchors=bits, parent_matrix=parent, initial_parent_matrix=initial)
            _same(got, want)
            if st["rotation"] == 0.0:  # no trigonometry involved: the floats themselves agree
                assert (np.array(got.matrix[:], dtype=np.float32) == want["matrix"]).all()
                assert (np.array(got.border_matrix[:], dtype=np.float32) == want["border_matrix"]).all()


def test_anchor_semantics_by_example():
    """A 100x50 child at (10, 20) in a parent that grew by (+40, +30): which edges move for the common anchor sets."""
    child = sv.element_state((10, 20), (100, 50))
    p0 = R.picture_state((8, 8), dict(pos=(5, 7), size=(200, 100), rotation=0.0, border=(0, 0, 0, 0), tex_offset=(0, 0), transparency=0.0))["matrix"]
    p1 = R.picture_state((8, 8), dict(pos=(5, 7), size=(240, 130), rotation=0.0, border=(0, 0, 0, 0), tex_offset=(0, 0), transparency=0.0))["matrix"]

    def rect(anchors):
        m = np.array(sv.compute_picture_state((8, 8), child, anchors=anchors, parent_matrix=p1, initial_parent_matrix=p0).matrix[:]).reshape(4, 4)
        return (m[3, 0], m[3, 1], m[0, 0], m[1, 1])  # x, y, w, h

    A = sv
    assert rect(A.ANCHOR_TOP_LEFT) == (15, 27, 100, 50)                                               # pinned: follows the parent's origin only
    assert rect(A.ANCHOR_TOP_RIGHT) == (55, 27, 100, 50)                                              # rides the right edge
    assert rect(A.ANCHOR_BOTTOM_LEFT) == (15, 57, 100, 50)                                            # rides the bottom edge
    assert rect(A.ANCHOR_BOTTOM_RIGHT) == (55, 57, 100, 50)                                           # rides the corner
    assert rect(A.ANCHOR_TOP_LEFT | A.ANCHOR_TOP_RIGHT) == (15, 27, 140, 50)                          # stretches horizontally
    assert rect(A.ANCHOR_TOP_LEFT | A.ANCHOR_BOTTOM_LEFT) == (15, 27, 100, 80)                        # stretches vertically
    assert rect(A.ANCHOR_TOP_LEFT | A.ANCHOR_BOTTOM_RIGHT) == (15, 27, 140, 80)                       # stretches both ways
    assert rect(A.ANCHOR_TOP_RIGHT | A.ANCHOR_BOTTOM_RIGHT) == (55, 27, 100, 80)                      # right edge, full height
    assert rect(A.ANCHOR_BOTTOM_LEFT | A.ANCHOR_BOTTOM_RIGHT) == (15, 57, 140, 50)                    # bottom edge, full width
    assert rect(0) == rect(A.ANCHOR_TOP_LEFT)                                                         # empty set = [.anchorTopLeft] (:62)


def test_transition_interpolates_continuous_fields_and_jumps_discrete_ones():
    """computeElementState :193-205"""
    rng = np.random.default_rng(3)
    for pct in (0.0, 0.25, 0.5, 1.0, 1.2):
        a, b = _rand_state(rng), _rand_state(rng)
        want = R.picture_state((320, 240), a, nxt=b, pct=pct)
        got = sv.compute_picture_state((320, 240), _native_state(a), next=_native_state(b), pct=pct)
        _same(got, want)
    a, b = _rand_state(rng), _rand_state(rng)
    # next without pct (and pct without next) leaves the current state alone (:236-241)
    _same(sv.compute_picture_state((320, 240), _native_state(a), next=_native_state(b)), R.picture_state((320, 240), a))
    _same(sv.compute_picture_state((320, 240), _native_state(a), pct=0.5), R.picture_state((320, 240), a))


def test_animator_state_machine():
    """setState / computedState / impl :54-128 with the caller's clock"""
    canvas = (1280, 720)
    src = sv.create_picture_sample(64, 36, sv.NV12, "s", "w")
    base = dict(rotation=0.0, border=(0, 0, 0, 0), tex_offset=(0, 0), transparency=0.0)
    s0, s1, s2 = dict(base, pos=(0, 0, 1.0), size=(100, 100)), dict(base, pos=(200, 100, 1.0), size=(300, 200), transparency=0.5), dict(base, pos=(0, 0), size=(50, 50))
    a = sv.PictureAnimator(canvas)
    assert a.apply(src, 0.0) is None                                   # no state yet -> .nothing (:125-127)
    with pytest.raises(sv.ComputeError) as e:
        a.computed_state((64, 36), 0.0)
    assert "noCurrentState" in str(e.value)
    a.set_state(_native_state(s0), duration=5.0, now=10.0)            # first state lands at once whatever the duration (:56)
    _same(a.computed_state((64, 36), 10.0), R.picture_state((64, 36), s0))
    a.set_state(_native_state(s1), duration=4.0, now=20.0)            # a timed transition
    for now in (20.0, 21.0, 23.5):
        _same(a.computed_state((64, 36), now), R.picture_state((64, 36), s0, nxt=s1, pct=(now - 20.0) / 4.0))
    q = a.apply(src, 22.0)
    want = R.picture_state((64, 36), s0, nxt=s1, pct=0.5)
    i = q.info()
    assert np.allclose(np.array(i.matrix[:]), R.project(canvas, want["matrix"]), rtol=1e-5, atol=1e-6)
    assert np.allclose(np.array(i.border_matrix[:]), R.project(canvas, want["border_matrix"]), rtol=1e-5, atol=1e-6)
    assert abs(i.opacity - 0.75) < 1e-6 and i.z_index == 2 and q.revision() == a.revision
    _same(a.computed_state((64, 36), 24.0), R.picture_state((64, 36), s1))   # the transition has ended: next became current
    _same(a.computed_state((64, 36), 30.0), R.picture_state((64, 36), s1))
    a.set_state(_native_state(s2), duration=0.0, now=31.0)            # duration <= 0 replaces at once
    _same(a.computed_state((64, 36), 31.0), R.picture_state((64, 36), s2))
    a.set_state(_native_state(dict(s2, hidden=True)), duration=0.0, now=32.0)
    assert a.apply(src, 32.0) is None                                  # hidden -> .nothing (:108-110)
    b = sv.PictureAnimator(canvas)
    assert a.revision != b.revision and a.revision


def test_animator_parent_chain():
    """impl :107-128: the parent's computed state offsets the child and scales its opacity; the initial parent state is latched
    after the first frame, so the first frame sees the parent's whole size as its size change."""
    canvas = (640, 360)
    src = sv.create_picture_sample(64, 36, sv.NV12, "s", "w")
    base = dict(rotation=0.0, border=(0, 0, 0, 0), tex_offset=(0, 0))
    ps0, ps1 = dict(base, pos=(40, 30), size=(200, 100), transparency=0.5), dict(base, pos=(40, 30), size=(300, 160), transparency=0.5)
    cs = dict(base, pos=(10, 10, 2.0), size=(50, 20), transparency=0.2, anchors=sv.ANCHOR_BOTTOM_RIGHT)
    parent = sv.PictureAnimator(canvas)
    child = sv.PictureAnimator(canvas, parent=parent)
    child.set_state(_native_state(cs))                                # anchors come from the state (:62)
    assert child.apply(src, 0.0) is None                               # parent has no state -> .nothing
    parent.set_state(_native_state(ps0))
    pm0 = R.picture_state((64, 36), ps0)["matrix"]
    first = child.apply(src, 0.0).info()
    want = R.picture_state((64, 36), cs, anchors=[R.BR], parent=pm0, initial_parent=None)
    assert np.allclose(np.array(first.matrix[:]), R.project(canvas, want["matrix"]), rtol=1e-5, atol=1e-6)
    assert abs(first.opacity - 0.8 * 0.5) < 1e-6
    second = child.apply(src, 1.0).info()                              # initial parent state now latched: no size change
    want = R.picture_state((64, 36), cs, anchors=[R.BR], parent=pm0, initial_parent=pm0)
    assert np.allclose(np.array(second.matrix[:]), R.project(canvas, want["matrix"]), rtol=1e-5, atol=1e-6)
    parent.set_state(_native_state(ps1), duration=2.0, now=2.0)       # the parent grows; the child rides its corner
    mid = R.picture_state((64, 36), ps0, nxt=ps1, pct=0.5)["matrix"]
    third = child.apply(src, 3.0).info()
    want = R.picture_state((64, 36), cs, anchors=[R.BR], parent=mid, initial_parent=pm0)
    m = np.array(third.matrix[:])
    assert np.allclose(m, R.project(canvas, want["matrix"]), rtol=1e-5, atol=1e-6)
    un = np.linalg.inv(R.ortho(canvas).astype(np.float64)) @ m.reshape(4, 4).T
    assert np.allclose([un[0, 3], un[1, 3], un[0, 0], un[1, 1]], [40 + 10 + 50, 30 + 10 + 30, 50, 20], atol=1e-3)
    child.set_parent(None)
    alone = child.apply(src, 4.0).info()
    assert abs(alone.opacity - 0.8) < 1e-6


def test_element_state_validation():
    st = sv.element_state((0, 0), (10, 10))
    st.parent_anchors = 16
    with pytest.raises(sv.ComputeError):
        sv.compute_picture_state((8, 8), st)
    with pytest.raises(sv.ComputeError):
        sv.compute_picture_state((8, 8), sv.element_state((0, 0), (10, 10)), anchors=99)
