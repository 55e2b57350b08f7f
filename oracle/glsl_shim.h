// glsl_shim.h -- just enough of GLSL's vector types and built-ins, in C++, to compile functions of the
// reference's compute shaders from the text where it lies (oracle/build_ref_shaders.py extracts them
// from /root/reference/data/shaders into oracle/_ref/; nothing of the reference is copied into this
// repository).  Test infrastructure, like everything under oracle/: it exists to pin the CPU checker
// (lucid_oracle.cpp) against the reference's own source for the arithmetic that decides coverage.
//
// Conventions of the translation (applied by the extractor, stated here because they are part of the pin):
//   * every floating literal gets an `f` suffix (GLSL literals are 32-bit floats);
//   * swizzles become calls (`v.xyz` -> `v.xyz()`), `out` / `inout` parameters become references;
//   * int(x) / uint(x) of a float become saturating conversions (GLSL leaves out-of-range values
//     undefined; saturation is what the CUDA kernels and the checker do);
//   * inversesqrt(x) := 1 / sqrt(x), length(v) := sqrt(dot(v, v)), each operation rounded to binary32
//     (compiled with -ffp-contract=off): the floating-point contract of DESIGN.md section 4.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

namespace glsl {

typedef uint32_t uint;

inline int glsl_int(float x) {
	if(!(x == x))
		return 0;
	if(x >= 2147483648.0f)
		return 2147483647;
	if(x <= -2147483648.0f)
		return -2147483647 - 1;
	return (int)x;
}
inline int glsl_int(int x) { return x; }
inline int glsl_int(uint x) { return (int)x; }
inline uint glsl_uint(float x) {
	if(!(x == x) || x <= 0.0f)
		return 0u;
	if(x >= 4294967296.0f)
		return 0xffffffffu;
	return (uint)x;
}
inline uint glsl_uint(uint x) { return x; }
inline uint glsl_uint(int x) { return (uint)x; }

template <class T> struct tvec2;
template <class T> struct tvec3;
template <class T> struct tvec4;

template <class T> struct tvec2 {
	T x, y;
	tvec2() : x(), y() {}
	tvec2(T a, T b) : x(a), y(b) {}
	explicit tvec2(T s) : x(s), y(s) {}
	template <class U> explicit tvec2(const tvec2<U> &o) : x((T)o.x), y((T)o.y) {}
	T &operator[](int i) { return (&x)[i]; }
	const T &operator[](int i) const { return (&x)[i]; }
};
template <class T> struct tvec3 {
	T x, y, z;
	tvec3() : x(), y(), z() {}
	tvec3(T a, T b, T c) : x(a), y(b), z(c) {}
	explicit tvec3(T s) : x(s), y(s), z(s) {}
	T &operator[](int i) { return (&x)[i]; }
	const T &operator[](int i) const { return (&x)[i]; }
	tvec2<T> xy() const { return tvec2<T>(x, y); }
};
template <class T> struct tvec4 {
	T x, y, z, w;
	tvec4() : x(), y(), z(), w() {}
	tvec4(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
	explicit tvec4(T s) : x(s), y(s), z(s), w(s) {}
	tvec4(const tvec3<T> &v, T d) : x(v.x), y(v.y), z(v.z), w(d) {}
	tvec4(const tvec2<T> &a, const tvec2<T> &b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
	template <class U> tvec4(const tvec2<U> &a, const tvec2<U> &b) : x((T)a.x), y((T)a.y), z((T)b.x), w((T)b.y) {}
	template <class U> explicit tvec4(const tvec4<U> &o);
	T &operator[](int i) { return (&x)[i]; }
	const T &operator[](int i) const { return (&x)[i]; }
	tvec3<T> xyz() const { return tvec3<T>(x, y, z); }
	tvec3<T> xzw() const { return tvec3<T>(x, z, w); }
	tvec2<T> xy() const { return tvec2<T>(x, y); }
	tvec2<T> zw() const { return tvec2<T>(z, w); }
	void mul_xyz(T s) { x *= s, y *= s, z *= s; }
	void set_xyz(const tvec3<T> &v) { x = v.x, y = v.y, z = v.z; }
};
// conversions between component types: float -> uint / int saturate, everything else is a plain cast
template <class T, class U> inline T convertComponent(U v) { return (T)v; }
template <> inline uint convertComponent<uint, float>(float v) { return glsl_uint(v); }
template <> inline int convertComponent<int, float>(float v) { return glsl_int(v); }
template <class T> template <class U>
tvec4<T>::tvec4(const tvec4<U> &o)
	: x(convertComponent<T, U>(o.x)), y(convertComponent<T, U>(o.y)), z(convertComponent<T, U>(o.z)),
	  w(convertComponent<T, U>(o.w)) {}

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec2<int> ivec2;
typedef tvec3<int> ivec3;
typedef tvec4<int> ivec4;
typedef tvec2<uint> uvec2;
typedef tvec3<uint> uvec3;
typedef tvec4<uint> uvec4;
typedef tvec3<bool> bvec3;

// uvec4(uvec3, uint) with an int-typed second argument, ivec4(uint) broadcast
inline uvec4 make_uvec4(const uvec3 &v, uint w) { return uvec4(v, w); }

#define GLSL_VEC_OPS(V, N, EXPR2, EXPRS)                                                            \
	template <class T> inline V<T> operator+(const V<T> &a, const V<T> &b) { return EXPR2(+); }    \
	template <class T> inline V<T> operator-(const V<T> &a, const V<T> &b) { return EXPR2(-); }    \
	template <class T> inline V<T> operator*(const V<T> &a, const V<T> &b) { return EXPR2(*); }    \
	template <class T> inline V<T> operator+(const V<T> &a, T b) { return EXPRS(+); }               \
	template <class T> inline V<T> operator-(const V<T> &a, T b) { return EXPRS(-); }               \
	template <class T> inline V<T> operator*(const V<T> &a, T b) { return EXPRS(*); }               \
	template <class T> inline V<T> &operator+=(V<T> &a, const V<T> &b) { return a = a + b; }        \
	template <class T> inline V<T> &operator-=(V<T> &a, const V<T> &b) { return a = a - b; }        \
	template <class T> inline V<T> &operator*=(V<T> &a, const V<T> &b) { return a = a * b; }        \
	template <class T> inline V<T> &operator*=(V<T> &a, T b) { return a = a * b; }

#define E2_2(op) tvec2<T>(a.x op b.x, a.y op b.y)
#define ES_2(op) tvec2<T>(a.x op b, a.y op b)
#define E2_3(op) tvec3<T>(a.x op b.x, a.y op b.y, a.z op b.z)
#define ES_3(op) tvec3<T>(a.x op b, a.y op b, a.z op b)
#define E2_4(op) tvec4<T>(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w)
#define ES_4(op) tvec4<T>(a.x op b, a.y op b, a.z op b, a.w op b)
GLSL_VEC_OPS(tvec2, 2, E2_2, ES_2)
GLSL_VEC_OPS(tvec3, 3, E2_3, ES_3)
GLSL_VEC_OPS(tvec4, 4, E2_4, ES_4)

template <class T> inline tvec2<T> operator*(T s, const tvec2<T> &a) { return a * s; }
template <class T> inline tvec3<T> operator*(T s, const tvec3<T> &a) { return a * s; }
template <class T> inline tvec4<T> operator*(T s, const tvec4<T> &a) { return a * s; }
template <class T> inline tvec3<T> operator-(const tvec3<T> &a) { return tvec3<T>(-a.x, -a.y, -a.z); }
// GLSL's == on vectors is "all components equal"
template <class T> inline bool operator==(const tvec3<T> &a, const tvec3<T> &b) {
	return a.x == b.x && a.y == b.y && a.z == b.z;
}
// mixed int / float arithmetic (implicit int -> float conversion of GLSL)
inline vec4 operator*(const vec4 &a, const ivec4 &b) {
	return vec4(a.x * float(b.x), a.y * float(b.y), a.z * float(b.z), a.w * float(b.w));
}
inline vec3 operator*(const vec3 &a, int b) { return a * float(b); }
inline vec3 &operator*=(vec3 &a, int b) { return a = a * float(b); }
inline ivec4 operator>>(const ivec4 &a, const ivec4 &b) { return ivec4(a.x >> b.x, a.y >> b.y, a.z >> b.z, a.w >> b.w); }
inline uvec4 operator>>(const uvec4 &a, int b) { return uvec4(a.x >> b, a.y >> b, a.z >> b, a.w >> b); }
inline ivec4 operator&(const ivec4 &a, int b) { return ivec4(a.x & b, a.y & b, a.z & b, a.w & b); }

inline float min(float a, float b) { return b < a ? b : a; } // GLSL: y < x ? y : x
inline float max(float a, float b) { return a < b ? b : a; } // GLSL: x < y ? y : x
inline float min(float a, int b) { return min(a, float(b)); }
inline float max(float a, int b) { return max(a, float(b)); }
inline int min(int a, int b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline vec2 min(const vec2 &a, const vec2 &b) { return vec2(min(a.x, b.x), min(a.y, b.y)); }
inline vec2 max(const vec2 &a, const vec2 &b) { return vec2(max(a.x, b.x), max(a.y, b.y)); }
inline ivec4 min(const ivec4 &a, int b) { return ivec4(min(a.x, b), min(a.y, b), min(a.z, b), min(a.w, b)); }
inline ivec4 max(const ivec4 &a, int b) { return ivec4(max(a.x, b), max(a.y, b), max(a.z, b), max(a.w, b)); }
inline float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }
inline int clamp(int v, int lo, int hi) { return min(max(v, lo), hi); }
inline vec4 clamp(const vec4 &v, const vec4 &lo, const vec4 &hi) {
	return vec4(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y), clamp(v.z, lo.z, hi.z), clamp(v.w, lo.w, hi.w));
}
inline float floor(float x) { return std::floor(x); }
inline float ceil(float x) { return std::ceil(x); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline float dot(const vec3 &a, const vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(const vec3 &a, const vec3 &b) {
	return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline float length(const vec3 &v) { return std::sqrt(dot(v, v)); }

inline uint floatBitsToUint(float f) {
	uint u;
	memcpy(&u, &f, 4);
	return u;
}
inline float uintBitsToFloat(uint u) {
	float f;
	memcpy(&f, &u, 4);
	return f;
}
inline vec2 uintBitsToFloat(const uvec2 &v) { return vec2(uintBitsToFloat(v.x), uintBitsToFloat(v.y)); }
inline vec2 fract(const vec2 &v) { return vec2(v.x - std::floor(v.x), v.y - std::floor(v.y)); }
inline uvec3 floatBitsToUint(const vec3 &v) { return uvec3(floatBitsToUint(v.x), floatBitsToUint(v.y), floatBitsToUint(v.z)); }
inline uvec4 floatBitsToUint(const vec4 &v) {
	return uvec4(floatBitsToUint(v.x), floatBitsToUint(v.y), floatBitsToUint(v.z), floatBitsToUint(v.w));
}
inline vec3 uintBitsToFloat(const uvec3 &v) { return vec3(uintBitsToFloat(v.x), uintBitsToFloat(v.y), uintBitsToFloat(v.z)); }

// column-major 4x4 matrix times column vector, each product and sum rounded on its own, in the order
// GLSL defines for M * v: sum over columns c of M[c] * v[c]
struct mat4 {
	vec4 col[4];
};
inline vec4 operator*(const mat4 &m, const vec4 &v) {
	vec4 r = m.col[0] * v.x;
	r = r + m.col[1] * v.y;
	r = r + m.col[2] * v.z;
	r = r + m.col[3] * v.w;
	return r;
}

inline vec3 clamp(const vec3 &v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline int bitCount(uint v) { return __builtin_popcount(v); }
inline int findLSB(uint v) { return v == 0 ? -1 : __builtin_ctz(v); }
inline void swap(float &a, float &b) {
	float t = a;
	a = b, b = t;
}
inline void swap(uint &a, uint &b) {
	uint t = a;
	a = b, b = t;
}

// pow: driver-defined in GLSL.  This is the polynomial contract of DESIGN.md section 4 (the same
// definition as orc_pow in lucid_oracle.cpp and pow_poly in lucid_b200/csrc/common.cuh): what the pin of
// shadeSample covers is the reference's arithmetic *around* pow, not pow itself.
inline float contract_log2(float x) {
	uint ix = floatBitsToUint(x);
	int e = (int)(ix - 0x3f3504f3u) >> 23;
	float m = uintBitsToFloat(ix - ((uint)e << 23));
	float f = m - 1.0f;
	float p = -0.146203533f;
	p = fmaf(p, f, 0.23420985f);
	p = fmaf(p, f, -0.24882181f);
	p = fmaf(p, f, 0.287075609f);
	p = fmaf(p, f, -0.360241979f);
	p = fmaf(p, f, 0.48092404f);
	p = fmaf(p, f, -0.721352756f);
	p = fmaf(p, f, 1.4426949f);
	return fmaf(p, f, (float)e);
}
inline float contract_exp2(float t) {
	float n = std::floor(t + 0.5f);
	float r = t - n;
	float p = 0.00134004327f;
	p = fmaf(p, r, 0.00967603736f);
	p = fmaf(p, r, 0.0555032715f);
	p = fmaf(p, r, 0.240221068f);
	p = fmaf(p, r, 0.693147182f);
	p = fmaf(p, r, 1.0f);
	int ni = glsl_int(n);
	if(ni < -126)
		return 0.0f;
	if(ni > 127)
		ni = 127;
	return uintBitsToFloat(floatBitsToUint(p) + ((uint)ni << 23));
}
inline float pow(float x, float y) {
	if(!(x > 0.0f))
		return 0.0f;
	return contract_exp2(y * contract_log2(x));
}

// shared-memory atomics of a single emulated invocation
inline uint atomicAdd(uint &mem, uint v) {
	uint old = mem;
	mem += v;
	return old;
}
inline int atomicAdd(int &mem, int v) {
	int old = mem;
	mem += v;
	return old;
}

} // namespace glsl
