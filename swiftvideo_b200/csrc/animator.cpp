// animator.cpp -- see animator.h for the reference map.
#include "animator.h"

#include <atomic>
#include <cmath>

namespace svb {

namespace {
inline float mix(float a, float b, float t) { return a + (b - a) * t; }  // interpolate(), animator.pic.swift:302-304
inline Vector4 fillOf(const ElementState& s) { return s.hasFillColor ? s.fillColor : Vector4{0, 0, 0, 0}; }  // getFillColor :335-342
// length of a matrix column's xy part: the size a T*R*S matrix was built with, whatever its rotation (:243-250)
inline float extent(float a, float b) { return std::sqrt(a * a + b * b); }
}  // namespace

ElementState computeElementState(const ElementState& c, const ElementState& n, float t) {
    ElementState s;
    s.picPos = Vector3{mix(c.picPos.x, n.picPos.x, t), mix(c.picPos.y, n.picPos.y, t), mix(c.picPos.z, n.picPos.z, t)};
    s.size = Vector2{mix(c.size.x, n.size.x, t), mix(c.size.y, n.size.y, t)};
    s.textureOffset = Vector2{mix(c.textureOffset.x, n.textureOffset.x, t), mix(c.textureOffset.y, n.textureOffset.y, t)};
    s.rotation = mix(c.rotation, n.rotation, t);
    s.transparency = mix(c.transparency, n.transparency, t);
    s.picAspect = n.picAspect;  // the discrete fields jump to the target at once
    s.picOrigin = n.picOrigin;
    const Vector4 cf = fillOf(c), nf = fillOf(n);
    s.fillColor = Vector4{mix(cf.x, nf.x, t), mix(cf.y, nf.y, t), mix(cf.z, nf.z, t), mix(cf.w, nf.w, t)};
    s.hasFillColor = true;  // the interpolated message always carries a fill colour (possibly all zero)
    s.borderSize = Vector4{mix(c.borderSize.x, n.borderSize.x, t), mix(c.borderSize.y, n.borderSize.y, t),
                           mix(c.borderSize.z, n.borderSize.z, t), mix(c.borderSize.w, n.borderSize.w, t)};
    // hidden / parentAnchor are not part of the interpolated message (proto defaults): the animator reads them from its
    // current state and its own anchor list, never from this result
    return s;
}

// The reference moves three corner vertices case by case; the outcome per edge is:
//   the right (bottom) edge follows the parent's growth when any anchor names the right (bottom) side;
//   the left (top) edge follows it too unless an anchor holds the element to the left (top) side -- then the element stretches.
// Edges are formed as (rel + size) + delta and rel + delta so the float results are the reference's.
void computePositionSize(Vector3 basePos, Vector3 baseSize, Vector3 parentPos, Vector3 delta, unsigned anchors, Vector3* pos, Vector3* size) {
    const Vector3 rel{basePos.x + parentPos.x, basePos.y + parentPos.y, basePos.z + 0.f};
    const bool rightFollows = anchors & (anchorTopRight | anchorBottomRight), leftHeld = anchors & (anchorTopLeft | anchorBottomLeft);
    const bool bottomFollows = anchors & (anchorBottomLeft | anchorBottomRight), topHeld = anchors & (anchorTopLeft | anchorTopRight);
    float left = rel.x, right = rel.x + baseSize.x, top = rel.y, bottom = rel.y + baseSize.y, z = rel.z;
    if (rightFollows) {
        right += delta.x;
        if (!leftHeld) left += delta.x;
    }
    if (bottomFollows) {
        bottom += delta.y;
        if (!topHeld) top += delta.y;
    }
    if ((anchors & anchorBottomRight) && !(anchors & anchorTopLeft)) z += delta.z;  // whole-vertex adds carry the (zero) z delta
    *pos = Vector3{left, top, z};
    *size = Vector3{right - left, bottom - top, 1.0f};
}

Matrix4 computeTextureMatrix(Vector2 sampleSize, Vector3 geometrySize, Vector2 textureOffset, AspectMode aspect) {
    const float origAspect = sampleSize.x / sampleSize.y;
    const float geomAspect = geometrySize.x / geometrySize.y;
    float scalex, scaley;
    switch (aspect) {
    case AspectMode::aspectFit:
        scalex = origAspect > geomAspect ? 1.0f : origAspect / geomAspect;
        scaley = origAspect <= geomAspect ? 1.0f : geomAspect / origAspect;
        break;
    case AspectMode::aspectFill:
        scalex = origAspect <= geomAspect ? 1.0f : origAspect / geomAspect;
        scaley = origAspect > geomAspect ? 1.0f : geomAspect / origAspect;
        break;
    default:
        return Matrix4::identity();
    }
    return Matrix4::translation(Vector3{textureOffset.x + (1.0f - scalex) / 2, textureOffset.y + (1.0f - scaley) / 2, 0}) *
           Matrix4::scale(Vector3{scalex, scaley, 1.0f});
}

ComputedPictureState computePictureState(Vector2 sampleSize, const ElementState& current, const PictureStateInputs& in) {
    const ElementState state = (in.next && in.pct) ? computeElementState(current, *in.next, *in.pct) : current;
    Vector3 parentPos{0, 0, 0}, parentSize{0, 0, 0}, initialSize{0, 0, 0};
    if (in.parent) {
        parentPos = Vector3{in.parent->m41, in.parent->m42, in.parent->m43};
        parentSize = Vector3{extent(in.parent->m11, in.parent->m12), extent(in.parent->m21, in.parent->m22), 0};
    }
    if (in.initialParent) initialSize = Vector3{extent(in.initialParent->m11, in.initialParent->m12), extent(in.initialParent->m21, in.initialParent->m22), 0};
    const Vector3 delta{parentSize.x - initialSize.x, parentSize.y - initialSize.y, parentSize.z - initialSize.z};
    Vector3 rel, size;
    computePositionSize(state.picPos, Vector3{state.size.x, state.size.y, 0}, parentPos, delta, in.anchors, &rel, &size);
    // a centre origin shifts by half the state's own size, not the anchored one (:252)
    const Vector3 add = state.picOrigin == PicOrigin::originTopLeft ? Vector3{0, 0, 0} : Vector3{-(state.size.x / 2), -(state.size.y / 2), -0.f};
    const Vector3 pos{rel.x + add.x, rel.y + add.y, rel.z + add.z};
    const Vector3 borderPos{pos.x - state.borderSize.x, pos.y - state.borderSize.y, pos.z - 0.f};
    const Vector3 borderSize{state.borderSize.x + size.x + state.borderSize.z, state.borderSize.y + size.y + state.borderSize.w, 1};
    const Matrix4 rot = Matrix4::rotation(Vector4{0, 0, 1, state.rotation});
    ComputedPictureState out;
    out.matrix = Matrix4::translation(pos) * rot * Matrix4::scale(size);
    out.textureMatrix = computeTextureMatrix(sampleSize, size, state.textureOffset, state.picAspect);
    out.borderMatrix = Matrix4::translation(borderPos) * rot * Matrix4::scale(borderSize);
    out.fillColor = fillOf(state);
    out.opacity = 1.0f - state.transparency;
    return out;
}

PictureSample projectPicture(const PictureSample& sample, Vector2 canvasSize, const ComputedPictureState& cs, float parentOpacity,
                             const std::string& revision) {
    const Matrix4 projection = Matrix4::ortho(canvasSize);
    PictureSample out = sample;
    out.transform = projection * cs.matrix;
    out.texTransform = cs.textureMatrix;
    out.borderTransform = projection * cs.borderMatrix;
    out.bgColor = cs.fillColor;
    out.alpha = cs.opacity * parentOpacity;
    if (!revision.empty()) out.idRevision = revision;
    return out;
}

PictureSample animatePicture(const PictureSample& sample, Vector2 canvasSize, const ElementState& state, float parentOpacity,
                             const std::string& revision) {
    return projectPicture(sample, canvasSize, computePictureState(sample.size(), state), parentOpacity, revision);
}

// ---- the stateful animator ---------------------------------------------------------------------------------------------

static std::atomic<unsigned long long> g_animatorSerial{0};

PictureAnimator::PictureAnimator(Vector2 canvasSize, std::shared_ptr<PictureAnimator> parent, unsigned parentAnchors)
    : canvasSize_(canvasSize), parent_(parent), anchors_(parentAnchors ? parentAnchors : (unsigned)anchorTopLeft),
      revision_("animator-" + std::to_string(++g_animatorSerial)) {}  // upstream: UUID().uuidString (:39)

void PictureAnimator::settle(double now) {
    if (next_ && start_ && duration_ && now >= *start_ + *duration_) {
        anchors_ = next_->parentAnchor ? next_->parentAnchor : (unsigned)anchorTopLeft;
        current_ = next_;
        next_.reset(), start_.reset(), duration_.reset(), initialParentState_.reset();
    }
}

void PictureAnimator::setState(const ElementState& state, double durationSeconds, double now) {
    std::lock_guard<std::mutex> g(mu_);
    settle(now);
    if (!current_ || durationSeconds <= 0) {
        current_ = state;
        next_.reset(), start_.reset(), duration_.reset(), initialParentState_.reset();
        anchors_ = state.parentAnchor ? state.parentAnchor : (unsigned)anchorTopLeft;
    } else {
        start_ = now;
        next_ = state;
        duration_ = durationSeconds;
    }
}

void PictureAnimator::setParent(std::shared_ptr<PictureAnimator> parent) {
    std::lock_guard<std::mutex> g(mu_);
    parent_ = parent;
}

ComputedPictureState PictureAnimator::computedState(Vector2 sampleSize, double now, const ComputedPictureState* parentState) {
    std::lock_guard<std::mutex> g(mu_);
    settle(now);
    if (!current_) throw ComputeError(ErrorCode::invalidValue, "noCurrentState");
    PictureStateInputs in;
    if (start_ && duration_) in.pct = (float)(now - *start_) / (float)*duration_;
    in.next = next_ ? &*next_ : nullptr;
    in.anchors = anchors_;
    in.parent = parentState ? &parentState->matrix : nullptr;
    in.initialParent = initialParentState_ ? &initialParentState_->matrix : nullptr;
    return computePictureState(sampleSize, *current_, in);
}

bool PictureAnimator::apply(const PictureSample& sample, double now, PictureSample* out) {
    std::shared_ptr<PictureAnimator> parent;
    {
        std::lock_guard<std::mutex> g(mu_);
        settle(now);
        if (!current_ || current_->hidden) return false;
        parent = parent_.lock();
    }
    try {
        std::optional<ComputedPictureState> parentState;
        if (parent) parentState = parent->computedState(sample.size(), now);  // the parent's own parent is not consulted (:113)
        const ComputedPictureState cs = computedState(sample.size(), now, parentState ? &*parentState : nullptr);
        {
            std::lock_guard<std::mutex> g(mu_);
            if (parentState && !initialParentState_) initialParentState_ = parentState;  // latched after the first use (:116-118)
        }
        *out = projectPicture(sample, canvasSize_, cs, parentState ? parentState->opacity : 1.0f, revision_);
        return true;
    } catch (const ComputeError&) {
        return false;  // .nothing (:125-127)
    }
}

}  // namespace svb
