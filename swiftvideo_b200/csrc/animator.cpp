// animator.cpp -- see animator.h for the reference map.
#include "animator.h"

namespace svb {

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

ComputedPictureState computePictureState(Vector2 sampleSize, const ElementState& state) {
    const Vector3 add = state.picOrigin == PicOrigin::originTopLeft ? Vector3{0, 0, 0} : Vector3{-state.size.x / 2, -state.size.y / 2, 0};
    const Vector3 size{state.size.x, state.size.y, 1.0f};  // computePositionSize returns z = 1 (:197)
    const Vector3 pos{state.picPos.x + add.x, state.picPos.y + add.y, state.picPos.z + add.z};
    const Vector3 borderPos{pos.x - state.borderSize.x, pos.y - state.borderSize.y, pos.z};
    const Vector3 borderSize{state.borderSize.x + size.x + state.borderSize.z, state.borderSize.y + size.y + state.borderSize.w, 1};
    const Matrix4 rot = Matrix4::rotation(Vector4{0, 0, 1, state.rotation});
    ComputedPictureState out;
    out.matrix = Matrix4::translation(pos) * rot * Matrix4::scale(size);
    out.textureMatrix = computeTextureMatrix(sampleSize, size, state.textureOffset, state.picAspect);
    out.borderMatrix = Matrix4::translation(borderPos) * rot * Matrix4::scale(borderSize);
    out.fillColor = state.hasFillColor ? state.fillColor : Vector4{0, 0, 0, 0};
    out.opacity = 1.0f - state.transparency;
    return out;
}

PictureSample animatePicture(const PictureSample& sample, Vector2 canvasSize, const ElementState& state, float parentOpacity,
                             const std::string& revision) {
    const ComputedPictureState cs = computePictureState(sample.size(), state);
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

}  // namespace svb
