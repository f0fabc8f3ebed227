// TEST INFRASTRUCTURE ONLY (oracle/): a minimal, self-written stand-in for the
// un-vendored glm dependency of the reference (CMakeLists.txt:17,25 take a
// user-supplied GLM_DIR; no version is pinned and glm is not in this image).
//
// Only the subset the reference's per-pixel hot path touches is provided, with
// glm's published component-wise definitions:
//   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z
//   normalize(v)  = v * (1 / sqrt(dot(v,v)))           (inversesqrt = 1/sqrt)
//   sign(x)       = (0 < x) - (x < 0)
//   v * mat3      = (dot(m[0],v), dot(m[1],v), dot(m[2],v))   (row vector)
//   rotate(m,a,v) = Rodrigues form, columns combined left to right
// Call sites this serves: lsvo.hpp:47,149  raycaster.hpp:152,156,192,193,197,200
// camera_controller.hpp:39,42,53  utils.cpp:96-97  volumetric.hpp:31-34  svo.hpp:142
// "parity unpinned at the glm boundary"  — see DESIGN.md.
#pragma once
#include <cmath>
#include <cstdint>
#include <string>

namespace glm {

template <typename T> struct tvec2 {
    T x, y;
    tvec2() : x(0), y(0) {}
    explicit tvec2(T s) : x(s), y(s) {}
    tvec2(T x_, T y_) : x(x_), y(y_) {}
    template <typename A, typename B> tvec2(A x_, B y_) : x(T(x_)), y(T(y_)) {}
};

template <typename T> struct tvec3 {
    T x, y, z;
    tvec3() : x(0), y(0), z(0) {}
    explicit tvec3(T s) : x(s), y(s), z(s) {}
    tvec3(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
    template <typename A, typename B, typename C> tvec3(A x_, B y_, C z_) : x(T(x_)), y(T(y_)), z(T(z_)) {}
    template <typename U> tvec3(const tvec2<U>& v, T z_) : x(T(v.x)), y(T(v.y)), z(z_) {}
    template <typename U> tvec3(const tvec3<U>& v) : x(T(v.x)), y(T(v.y)), z(T(v.z)) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    tvec3& operator-=(const tvec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    tvec3& operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
};

template <typename T> struct tvec4 {
    T x, y, z, w;
    tvec4() : x(0), y(0), z(0), w(0) {}
    tvec4(T x_, T y_, T z_, T w_) : x(x_), y(y_), z(z_), w(w_) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec3<int32_t> ivec3;
typedef tvec2<int32_t> ivec2;

// ---- vec2 ----
inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(float s, const vec2& a) { return vec2(s * a.x, s * a.y); }
inline vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }

// ---- vec3 ----
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator/(float s, const vec3& a) { return vec3(s / a.x, s / a.y, s / a.z); }
inline vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }

inline vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

inline vec3 abs(const vec3& v) { return vec3(std::fabs(v.x), std::fabs(v.y), std::fabs(v.z)); }
inline float sign(float x) { return float(0.0f < x) - float(x < 0.0f); }
inline vec3 sign(const vec3& v) { return vec3(sign(v.x), sign(v.y), sign(v.z)); }
inline float dot(const vec3& a, const vec3& b) {
    const vec3 t(a * b);
    return t.x + t.y + t.z;
}
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }

// ---- matrices (column major, m[col][row]) ----
struct mat4 {
    vec4 c[4];
    mat4() {}
    explicit mat4(float d) {
        c[0] = vec4(d, 0, 0, 0); c[1] = vec4(0, d, 0, 0); c[2] = vec4(0, 0, d, 0); c[3] = vec4(0, 0, 0, d);
    }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};

struct mat3 {
    vec3 c[3];
    mat3() {}
    explicit mat3(float d) { c[0] = vec3(d, 0, 0); c[1] = vec3(0, d, 0); c[2] = vec3(0, 0, d); }
    mat3(const mat4& m) {  // upper-left 3x3; implicit, as glm without GLM_FORCE_EXPLICIT_CTOR
        for (int i = 0; i < 3; ++i) c[i] = vec3(m[i].x, m[i].y, m[i].z);
    }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};

inline mat4 operator*(const mat4& a, const mat4& b) {
    mat4 r;
    for (int j = 0; j < 4; ++j)
        r[j] = a[0] * b[j][0] + a[1] * b[j][1] + a[2] * b[j][2] + a[3] * b[j][3];
    return r;
}

// row-vector times matrix
inline vec3 operator*(const vec3& v, const mat3& m) {
    return vec3(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z,
                m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
                m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z);
}

inline mat4 rotate(const mat4& m, float angle, const vec3& v) {
    const float c = std::cos(angle);
    const float s = std::sin(angle);
    const vec3 axis(normalize(v));
    const vec3 temp((1.0f - c) * axis);
    float R[3][3];
    R[0][0] = c + temp[0] * axis[0];
    R[0][1] = temp[0] * axis[1] + s * axis[2];
    R[0][2] = temp[0] * axis[2] - s * axis[1];
    R[1][0] = temp[1] * axis[0] - s * axis[2];
    R[1][1] = c + temp[1] * axis[1];
    R[1][2] = temp[1] * axis[2] + s * axis[0];
    R[2][0] = temp[2] * axis[0] + s * axis[1];
    R[2][1] = temp[2] * axis[1] - s * axis[0];
    R[2][2] = c + temp[2] * axis[2];
    mat4 r;
    r[0] = m[0] * R[0][0] + m[1] * R[0][1] + m[2] * R[0][2];
    r[1] = m[0] * R[1][0] + m[1] * R[1][1] + m[2] * R[1][2];
    r[2] = m[0] * R[2][0] + m[1] * R[2][1] + m[2] * R[2][2];
    r[3] = m[3];
    return r;
}

}  // namespace glm
