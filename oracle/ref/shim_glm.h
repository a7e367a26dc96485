// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// Minimal stand-in for the handful of glm types the reference's hot-path files use
// (glm itself is not vendored by the reference and is absent from this image).
// It exists only so that /root/reference/Code/*.{h,cpp} compile VERBATIM into
// oracle/_ref/libgcref.so (see oracle/Makefile). Semantics that matter for mesher
// parity, and that this file reproduces exactly:
//   * converting constructors between vec<N,T> of different T narrow per component
//     with a plain static_cast (u16vec3 -> u8vec3 at WorldRender.cpp:475,
//     ivec4 -> u8vec4 at WorldRender.cpp:304-312);
//   * ivec4 + ivec4, ivec4 >> int, ivec4 / int are per-component int arithmetic.
#pragma once
#include <cmath>
#include <cstdint>

namespace glm
{
template <int N, typename T> struct vec;

template <typename T> struct vec<2, T>
{
    union { T x, r; };
    union { T y, g; };
    vec() = default;
    template <typename A> explicit vec(A s) : x((T)s), y((T)s) {}
    template <typename A, typename B> vec(A a, B b) : x((T)a), y((T)b) {}
    template <typename U> vec(const vec<2, U>& o) : x((T)o.x), y((T)o.y) {}
};

template <typename T> struct vec<3, T>
{
    union { T x, r; };
    union { T y, g; };
    union { T z, b; };
    vec() = default;
    template <typename A> explicit vec(A s) : x((T)s), y((T)s), z((T)s) {}
    template <typename A, typename B, typename C> vec(A a, B b_, C c) : x((T)a), y((T)b_), z((T)c) {}
    template <typename U> vec(const vec<3, U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    template <typename U> vec& operator+=(const vec<3, U>& o) { x += (T)o.x; y += (T)o.y; z += (T)o.z; return *this; }
    template <typename U> vec& operator-=(const vec<3, U>& o) { x -= (T)o.x; y -= (T)o.y; z -= (T)o.z; return *this; }
    vec& operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

template <typename T> struct vec<4, T>
{
    union { T x, r; };
    union { T y, g; };
    union { T z, b; };
    union { T w, a; };
    vec() = default;
    template <typename A> explicit vec(A s) : x((T)s), y((T)s), z((T)s), w((T)s) {}
    template <typename A, typename B, typename C, typename D> vec(A a_, B b_, C c, D d) : x((T)a_), y((T)b_), z((T)c), w((T)d) {}
    template <typename U, typename D> vec(const vec<3, U>& o, D d) : x((T)o.x), y((T)o.y), z((T)o.z), w((T)d) {}
    template <typename U> vec(const vec<4, U>& o) : x((T)o.x), y((T)o.y), z((T)o.z), w((T)o.w) {}
};

#define GCREF_VEC_BINOP(op) \
    template <typename T> inline vec<2, T> operator op(const vec<2, T>& a, const vec<2, T>& b) { return vec<2, T>(a.x op b.x, a.y op b.y); } \
    template <typename T> inline vec<3, T> operator op(const vec<3, T>& a, const vec<3, T>& b) { return vec<3, T>(a.x op b.x, a.y op b.y, a.z op b.z); } \
    template <typename T> inline vec<4, T> operator op(const vec<4, T>& a, const vec<4, T>& b) { return vec<4, T>(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); } \
    template <typename T> inline vec<2, T> operator op(const vec<2, T>& a, T s) { return vec<2, T>(a.x op s, a.y op s); } \
    template <typename T> inline vec<3, T> operator op(const vec<3, T>& a, T s) { return vec<3, T>(a.x op s, a.y op s, a.z op s); } \
    template <typename T> inline vec<4, T> operator op(const vec<4, T>& a, T s) { return vec<4, T>(a.x op s, a.y op s, a.z op s, a.w op s); } \
    template <typename T> inline vec<3, T> operator op(T s, const vec<3, T>& a) { return vec<3, T>(s op a.x, s op a.y, s op a.z); }

GCREF_VEC_BINOP(+)
GCREF_VEC_BINOP(-)
GCREF_VEC_BINOP(*)
GCREF_VEC_BINOP(/)
#undef GCREF_VEC_BINOP

template <typename T> inline vec<4, T> operator>>(const vec<4, T>& a, int s) { return vec<4, T>(a.x >> s, a.y >> s, a.z >> s, a.w >> s); }
template <typename T> inline vec<3, T> operator-(const vec<3, T>& a) { return vec<3, T>(-a.x, -a.y, -a.z); }

template <typename T> inline bool operator==(const vec<2, T>& a, const vec<2, T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> inline bool operator==(const vec<3, T>& a, const vec<3, T>& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <typename T> inline bool operator==(const vec<4, T>& a, const vec<4, T>& b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
template <int N, typename T> inline bool operator!=(const vec<N, T>& a, const vec<N, T>& b) { return !(a == b); }

typedef vec<2, float> vec2;
typedef vec<3, float> vec3;
typedef vec<4, float> vec4;
typedef vec<2, int> ivec2;
typedef vec<3, int> ivec3;
typedef vec<4, int> ivec4;
typedef vec<3, uint8_t> u8vec3;
typedef vec<4, uint8_t> u8vec4;
typedef vec<2, uint16_t> u16vec2;
typedef vec<3, uint16_t> u16vec3;

struct mat4 { float m[16]; mat4() = default; explicit mat4(float) {} };

inline float length2(const vec3& v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
inline float length2(const vec2& v) { return v.x * v.x + v.y * v.y; }
inline float distance2(const vec2& a, const vec2& b) { return length2(a - b); }
inline float distance2(const vec3& a, const vec3& b) { return length2(a - b); }
inline float length(const vec3& v) { return std::sqrt(length2(v)); }
inline vec3 normalize(const vec3& v) { float l = length(v); return vec3(v.x / l, v.y / l, v.z / l); }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline mat4 translate(const mat4& m, const vec3&) { return m; }
}
