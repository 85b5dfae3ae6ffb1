// animator.h -- the part of PictureAnimator that turns an element's state into the matrices the mixer consumes
// (/root/reference/Sources/SwiftVideo/animator.pic.swift): computeTextureMatrix :207-227, computePictureState :229-272
// (without the parent/anchor bookkeeping of computePositionSize :154-198, which only moves pos/size before this point),
// and impl() :107-128, which projects with Matrix4(ortho) :326-333 and multiplies the opacities.
// the host-side step right before the hot path, here in native code so that a tick's uniforms are
// produced without leaving the library.
#pragma once
#include "compute.h"

namespace svb {

enum class AspectMode : int { aspectNone = 0, aspectFit = 1, aspectFill = 2 };  // Proto/Composition.proto AspectMode
enum class PicOrigin : int { originCenter = 0, originTopLeft = 1 };

struct ElementState {  // the fields of ElementState that computePictureState reads (Proto/Composition.proto:56-71)
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
};

struct ComputedPictureState {  // :141-147
    Matrix4 matrix, textureMatrix, borderMatrix;
    Vector4 fillColor;
    float opacity = 1.f;
};

Matrix4 computeTextureMatrix(Vector2 sampleSize, Vector3 geometrySize, Vector2 textureOffset, AspectMode aspect);  // :207-227
ComputedPictureState computePictureState(Vector2 sampleSize, const ElementState& state);                              // :229-272
// PictureAnimator.impl (:107-128): the sample re-issued with projected matrices, fill colour, opacity and revision.
PictureSample animatePicture(const PictureSample& sample, Vector2 canvasSize, const ElementState& state, float parentOpacity,
                             const std::string& revision);

}  // namespace svb
