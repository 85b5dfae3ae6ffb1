// vector_math.h -- the small part of VectorMath (nicklockwood/VectorMath, reference Package.swift:61,
// `from: "0.4.0"`, version unpinned upstream) that the hot path's host side uses: Matrix4 products,
// inverse, transpose, and the row-vector product behind PictureSample.zIndex().
//
// Conventions inferred from the reference's call sites: fields mCR are stored column by
// column (m41,m42,m43 = translation, animator.pic.swift:244,326-333), `A * B` applies B first, `v * M`
// treats v as a row vector.  After `.inverse.transpose` the 16 floats in memory are, row by row, the
// float4[4] the kernels dot against (compute.swift:151-155, kernels.cl.swift:27).
// PARITY UNPINNED for the fp32 rounding of inverse(): the reference has no test that fixes it.  Pixel
// parity does not depend on it -- ImageUniforms is an input of the oracle and of the kernels alike.
#pragma once
#include <cmath>

namespace svb {

struct Vector2 {
    float x = 0, y = 0;
};
struct Vector3 {
    float x = 0, y = 0, z = 0;
};
struct Vector4 {
    float x = 0, y = 0, z = 0, w = 0;
};

struct Matrix4 {
    // memory order == VectorMath's: m11 m12 m13 m14 | m21 ... (column 1 first)
    float m11 = 1, m12 = 0, m13 = 0, m14 = 0;
    float m21 = 0, m22 = 1, m23 = 0, m24 = 0;
    float m31 = 0, m32 = 0, m33 = 1, m34 = 0;
    float m41 = 0, m42 = 0, m43 = 0, m44 = 1;

    static Matrix4 identity() { return Matrix4(); }
    static Matrix4 from_array(const float* a) {
        Matrix4 m;
        float* d = &m.m11;
        for (int i = 0; i < 16; ++i) d[i] = a[i];
        return m;
    }
    static Matrix4 translation(Vector3 t) {
        Matrix4 m;
        m.m41 = t.x, m.m42 = t.y, m.m43 = t.z;
        return m;
    }
    static Matrix4 scale(Vector3 s) {
        Matrix4 m;
        m.m11 = s.x, m.m22 = s.y, m.m33 = s.z;
        return m;
    }
    // axis-angle (x, y, z, angle); the animator only ever passes the z axis (animator.pic.swift:264)
    static Matrix4 rotation(Vector4 r) {
        float len = std::sqrt(r.x * r.x + r.y * r.y + r.z * r.z);
        if (len == 0.f) return Matrix4();
        float x = r.x / len, y = r.y / len, z = r.z / len, c = std::cos(r.w), s = std::sin(r.w), t = 1.f - c;
        Matrix4 m;
        m.m11 = t * x * x + c, m.m12 = t * x * y + s * z, m.m13 = t * x * z - s * y;
        m.m21 = t * x * y - s * z, m.m22 = t * y * y + c, m.m23 = t * y * z + s * x;
        m.m31 = t * x * z + s * y, m.m32 = t * y * z - s * x, m.m33 = t * z * z + c;
        return m;
    }
    // Matrix4(_ ortho: Vector2), animator.pic.swift:326-333 (m43 = 1: zIndex = round(pos.z + 1))
    static Matrix4 ortho(Vector2 c) {
        Matrix4 m;
        m.m11 = 2.f / c.x, m.m22 = 2.f / c.y, m.m33 = 1.f;
        m.m41 = -1.f, m.m42 = -1.f, m.m43 = 1.f, m.m44 = 1.f;
        return m;
    }
    const float* data() const { return &m11; }
    float at(int col, int row) const { return (&m11)[col * 4 + row]; }
    float& at(int col, int row) { return (&m11)[col * 4 + row]; }

    Matrix4 transpose() const {
        Matrix4 t;
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r) t.at(c, r) = at(r, c);
        return t;
    }
    // adjugate / determinant, fp32
    Matrix4 inverse() const {
        const float* m = &m11;
        float inv[16];
        inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
        inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
        inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
        inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
        inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
        inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
        inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
        inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
        inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
        inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
        inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
        inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
        inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
        inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
        inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
        inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
        float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
        float s = 1.0f / det;
        Matrix4 r;
        float* d = &r.m11;
        for (int i = 0; i < 16; ++i) d[i] = inv[i] * s;
        return r;
    }
};

// (A * B): apply B first.  result.m(c,r) = sum_k A.m(k,r) * B.m(c,k)
inline Matrix4 operator*(const Matrix4& a, const Matrix4& b) {
    Matrix4 m;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s += a.at(k, r) * b.at(c, k);
            m.at(c, r) = s;
        }
    return m;
}

// row-vector product, w ignored (PictureSample.zIndex, sample.pict.linux.swift:116)
inline Vector3 operator*(Vector3 v, const Matrix4& m) {
    return Vector3{v.x * m.m11 + v.y * m.m21 + v.z * m.m31 + m.m41, v.x * m.m12 + v.y * m.m22 + v.z * m.m32 + m.m42,
                   v.x * m.m13 + v.y * m.m23 + v.z * m.m33 + m.m43};
}

}  // namespace svb
