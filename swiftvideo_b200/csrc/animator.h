// animator.h -- PictureAnimator in native code: an element's state (or a transition between two states) becomes the matrices
// the mixer consumes (/root/reference/Sources/SwiftVideo/animator.pic.swift):
//   computePositionSize :149-191 (the parent's size change moved through the element's anchors), computeElementState :193-205
//   (linear interpolation of a transition), computeTextureMatrix :207-227, computePictureState :229-272 (parent position and
//   size read off the parent's unprojected matrix), impl() :107-128 (Matrix4(ortho) :326-333, opacities multiplied,
//   initialParentState latched on first use), setState :54-80 (immediate, or a timed transition), computedState :82-102.
// the host-side step right before the hot path, here in native code so that a tick's uniforms are
// produced without leaving the library.  The reference drives transitions from its Clock; this class is clock-free: the
// caller passes "now" (seconds) and a transition is promoted the first time now reaches its end.
#pragma once
#include <memory>
#include <mutex>
#include <optional>

#include "compute.h"

namespace svb {

enum class AspectMode : int { aspectNone = 0, aspectFit = 1, aspectFill = 2 };  // Proto/Composition.proto AspectMode
enum class PicOrigin : int { originCenter = 0, originTopLeft = 1 };
enum PictureAnchor : unsigned {  // Proto/Composition.proto:31-36, as a set: bit n = enum value n
    anchorTopLeft = 1u << 0,
    anchorTopRight = 1u << 1,
    anchorBottomLeft = 1u << 2,
    anchorBottomRight = 1u << 3,
};

struct ElementState {  // the fields of ElementState that the picture animator reads (Proto/Composition.proto:56-71)
    Vector3 picPos;
    Vector2 size;
    Vector2 textureOffset;
    Vector4 borderSize;  // x = left, y = top, z = right, w = bottom (animator.pic.swift:257-259)
    Vector4 fillColor;
    float rotation = 0.f;
    float transparency = 0.f;
    AspectMode picAspect = AspectMode::aspectNone;
    PicOrigin picOrigin = PicOrigin::originTopLeft;
    bool hasFillColor = false;  // getFillColor(): (0,0,0,0) when unset (:334-342)
    bool hidden = false;        // impl() emits nothing for a hidden element (:108-110)
    unsigned parentAnchor = 0;  // set of PictureAnchor; empty = [.anchorTopLeft] (:62)
};

struct ComputedPictureState {  // :141-147
    Matrix4 matrix, textureMatrix, borderMatrix;
    Vector4 fillColor;
    float opacity = 1.f;
};

struct PictureStateInputs {                     // computePictureState's optional arguments (:229-235)
    const Matrix4* parent = nullptr;            // the parent's ComputedPictureState.matrix (unprojected)
    const Matrix4* initialParent = nullptr;     // the parent's matrix when this element first saw it (initialParentState)
    const ElementState* next = nullptr;         // the transition's target ...
    std::optional<float> pct;                   // ... and how far along it is; both must be set to interpolate (:236-241)
    unsigned anchors = anchorTopLeft;
};

ElementState computeElementState(const ElementState& current, const ElementState& next, float pct);                 // :193-205
// (position, size) of the element after the parent's offset and size change are applied through the anchors :149-191
void computePositionSize(Vector3 basePos, Vector3 baseSize, Vector3 parentPos, Vector3 parentSizeDelta, unsigned anchors, Vector3* pos,
                         Vector3* size);
Matrix4 computeTextureMatrix(Vector2 sampleSize, Vector3 geometrySize, Vector2 textureOffset, AspectMode aspect);   // :207-227
ComputedPictureState computePictureState(Vector2 sampleSize, const ElementState& state, const PictureStateInputs& in = {});  // :229-272
// One-shot PictureAnimator.impl (:107-128) for a parent-less element in a settled state.
PictureSample animatePicture(const PictureSample& sample, Vector2 canvasSize, const ElementState& state, float parentOpacity,
                             const std::string& revision);
// The projection impl() applies to a computed state (:118-124).
PictureSample projectPicture(const PictureSample& sample, Vector2 canvasSize, const ComputedPictureState& cs, float parentOpacity,
                             const std::string& revision);

class PictureAnimator {  // animator.pic.swift:29-139
public:
    PictureAnimator(Vector2 canvasSize, std::shared_ptr<PictureAnimator> parent = nullptr, unsigned parentAnchors = anchorTopLeft);
    // setState (:54-80): duration <= 0, or no current state yet, replaces the state at once; otherwise starts a transition
    // at `now` that ends (next becomes current, anchors re-read from it) at now + duration.
    void setState(const ElementState& state, double durationSeconds, double now);
    void setParent(std::shared_ptr<PictureAnimator> parent);                                          // :104-106
    // computedState (:82-102); throws ComputeError(invalidValue, "noCurrentState") like AnimatorError.noCurrentState
    ComputedPictureState computedState(Vector2 sampleSize, double now, const ComputedPictureState* parentState = nullptr);
    // impl (:107-128): false = nothing emitted (hidden element, or no state anywhere on the chain)
    bool apply(const PictureSample& sample, double now, PictureSample* out);
    const std::string& revision() const { return revision_; }

private:
    void settle(double now);  // promote a finished transition (the closure setState schedules on the clock :67-75)
    std::mutex mu_;
    Vector2 canvasSize_;
    std::weak_ptr<PictureAnimator> parent_;  // weak like the reference (:138)
    std::optional<ElementState> current_, next_;
    std::optional<double> start_, duration_;
    std::optional<ComputedPictureState> initialParentState_;
    unsigned anchors_;
    std::string revision_;
};

}  // namespace svb
