// Scalar 4x4 helpers for the host pass graph. Column-major like nrd::CommonSettings' matrices.
// Semantics follow the MathLib routines the reference host calls (ml.h:737-781 MvpToPlanes,
// ml.h:1031-1135 DecomposeProjection, Guts/f32.h:931-961 operator*, :1710-1720 InvertOrtho);
// op order follows the SSE originals (separate multiply and add — build with -ffp-contract=off) so the
// constant buffers come out bit-identical to a reference built without -mfma.
#pragma once
#include <cmath>

#include "constants.h"

namespace nrdb {
namespace mat {

inline float& at(Mat4& a, int r, int c) { return a.m[c * 4 + r]; }
inline float at(const Mat4& a, int r, int c) { return a.m[c * 4 + r]; }

inline Mat4 identity() {
    Mat4 r = {};
    r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
    return r;
}

inline Mat4 mul(const Mat4& a, const Mat4& b) {  // a * b, vectors are columns
    Mat4 r;
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++) {
            float s = b.m[c * 4 + 0] * a.m[0 * 4 + row];
            s = b.m[c * 4 + 1] * a.m[1 * 4 + row] + s;
            s = b.m[c * 4 + 2] * a.m[2 * 4 + row] + s;
            s = b.m[c * 4 + 3] * a.m[3 * 4 + row] + s;
            r.m[c * 4 + row] = s;
        }
    return r;
}

inline void negateColumn(Mat4& a, int c) {
    for (int r = 0; r < 4; r++) a.m[c * 4 + r] = -a.m[c * 4 + r];
}
inline void negateRow(Mat4& a, int r) {
    for (int c = 0; c < 4; c++) a.m[c * 4 + r] = -a.m[c * 4 + r];
}
inline void setTranslation(Mat4& a, float x, float y, float z) {
    a.m[12] = x; a.m[13] = y; a.m[14] = z; a.m[15] = 1.0f;
}

// Inverse of a rigid transform: transpose the rotation, translation = -(R^T t)
inline Mat4 invertOrtho(const Mat4& a) {
    Mat4 r = {};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) at(r, i, j) = at(a, j, i);
    float t[3] = {a.m[12], a.m[13], a.m[14]};
    for (int i = 0; i < 3; i++) {
        float s = t[0] * r.m[0 * 4 + i];
        s = t[1] * r.m[1 * 4 + i] + s;
        s = t[2] * r.m[2 * 4 + i] + s;
        r.m[12 + i] = -s;
    }
    r.m[15] = 1.0f;
    return r;
}

struct Projection {
    float frustum[4];  // (-x0, -y1, x0 - x1, y1 - y0): view-space xy = (uv * frustum.zw + frustum.xy) * z
    float projectY;    // |2 / (y1 - y0)|
    bool ortho, reversedZ, leftHanded;
};

inline Projection decomposeProjection(const Mat4& p) {
    // Frustum planes from the rows of the projection (D3D depth range)
    float row[4][4];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) row[r][c] = at(p, r, c);
    float l[4], r[4], b[4], t[4], f[4], n[4];
    for (int i = 0; i < 4; i++) {
        l[i] = row[3][i] + row[0][i];
        r[i] = row[3][i] - row[0][i];
        b[i] = row[3][i] + row[1][i];
        t[i] = row[3][i] - row[1][i];
        f[i] = row[3][i] - row[2][i];
        n[i] = row[2][i];
    }
    auto dot3 = [](const float* a, const float* b2) { return a[0] * b2[0] + a[1] * b2[1] + a[2] * b2[2]; };
    auto scale = [](float* a, float s) { for (int i = 0; i < 4; i++) a[i] *= s; };
    scale(l, 1.0f / std::sqrt(dot3(l, l)));
    scale(r, 1.0f / std::sqrt(dot3(r, r)));
    scale(b, 1.0f / std::sqrt(dot3(b, b)));
    scale(t, 1.0f / std::sqrt(dot3(t, t)));
    const float eps = 1e-7f;
    scale(n, 1.0f / std::fmax(std::sqrt(dot3(n, n)), eps));
    scale(f, 1.0f / std::fmax(std::sqrt(dot3(f, f)), eps));

    Projection out = {};
    out.reversedZ = std::fabs(n[3]) > std::fabs(f[3]);
    if (out.reversedZ)
        for (int i = 0; i < 4; i++) std::swap(n[i], f[i]);
    out.ortho = at(p, 3, 3) == 1.0f;
    float nearZ = -n[3];

    float x0, x1, y0, y1;
    if (out.ortho) {
        x0 = -l[3]; x1 = r[3]; y0 = -b[3]; y1 = t[3];
        if (at(p, 1, 1) < 0.0f) std::swap(y0, y1);
    } else {
        x0 = l[2] / l[0]; x1 = r[2] / r[0]; y0 = b[2] / b[1]; y1 = t[2] / t[1];
    }

    // Handedness: does the basis formed by the projection's x, y columns and its view direction have positive volume?
    float clipW = at(p, 3, 2) * nearZ + at(p, 3, 3);
    float c0[3] = {p.m[0], p.m[1], p.m[2]}, c1[3] = {p.m[4], p.m[5], p.m[6]}, c2[3];
    if (out.ortho) {
        float s = out.reversedZ ? -1.0f : 1.0f;
        c2[0] = p.m[8] * s; c2[1] = p.m[9] * s; c2[2] = p.m[10] * s;
    } else {
        c2[0] = 0.0f; c2[1] = 0.0f; c2[2] = clipW > 0.0f ? 1.0f : -1.0f;
    }
    float cr[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2], c0[0] * c1[1] - c0[1] * c1[0]};
    bool cmp = dot3(cr, c2) > 0.0f;
    out.leftHanded = at(p, 1, 1) > 0.0f ? cmp : !cmp;

    out.projectY = std::fabs(2.0f / (y1 - y0));
    out.frustum[0] = -x0;
    out.frustum[1] = -y1;
    out.frustum[2] = x0 - x1;
    out.frustum[3] = y1 - y0;
    return out;
}

}  // namespace mat
}  // namespace nrdb
