// glsl_shim.hpp -- just enough of GLSL 4.60 in C++17 to compile the reference's shading pass
// (src/shaders/shading_pass.frag.glsl and what it includes) for the CPU, unmodified except for
// the mechanical token rewrites done by oracle/build_ref.py. TEST INFRASTRUCTURE.
//
// Built-in semantics are the ones stated at the top of oracle/risltc_oracle.c (dot / cross /
// normalize / matrix products as plain fp32 sums in index order, no contraction), so the compiled
// reference and the C restatement can be compared bit for bit. Everything lives in namespace glsl;
// the shader code is compiled inside the same namespace so that these definitions hide <cmath>'s.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#undef M_PI
#undef M_INV_PI
#undef M_HALF_PI
#undef M_INFINITY

namespace glsl {

typedef unsigned int uint;

struct vec2; struct vec3; struct vec4;

// ---- swizzle proxies: a view onto N consecutive-or-not members of the parent's float array
template <typename V, int N, int A, int B>
struct swz2 {
	float d[N];
	operator V() const { return V(d[A], d[B]); }
	swz2& operator=(const V& v) { float a = v.x, b = v.y; d[A] = a; d[B] = b; return *this; }
	swz2& operator=(const swz2& o) { return *this = V(o); }
	swz2& operator*=(float s) { d[A] *= s; d[B] *= s; return *this; }
};
template <typename V, int N, int A, int B, int C>
struct swz3 {
	float d[N];
	operator V() const { return V(d[A], d[B], d[C]); }
	swz3& operator=(const V& v) { float a = v.x, b = v.y, c = v.z; d[A] = a; d[B] = b; d[C] = c; return *this; }
	swz3& operator=(const swz3& o) { return *this = V(o); }
};

struct ivec2 { int x, y; ivec2() : x(0), y(0) {} ivec2(int a, int b) : x(a), y(b) {} explicit ivec2(const vec2& v); };
struct uvec2 {
	uint x, y;
	uvec2() : x(0), y(0) {}
	uvec2(uint a, uint b) : x(a), y(b) {}
	uvec2(const ivec2& v) : x((uint) v.x), y((uint) v.y) {}
	uint& operator[](int i) { return i ? y : x; }
	uint operator[](int i) const { return i ? y : x; }
};
struct uvec4 {
	struct rg_t { uint v[4]; operator uvec2() const { return uvec2(v[0], v[1]); } };
	union { struct { uint x, y, z, w; }; struct { uint r, g, b, a; }; rg_t rg; };
	uvec4() : x(0), y(0), z(0), w(0) {}
	uvec4(uint a_, uint b_, uint c_, uint d_) : x(a_), y(b_), z(c_), w(d_) {}
};

struct vec2 {
	union { struct { float x, y; }; struct { float r, g; }; swz2<vec2, 2, 0, 1> xy; swz2<vec2, 2, 1, 0> yx; };
	vec2() : x(0.0f), y(0.0f) {}
	explicit vec2(float a) : x(a), y(a) {}
	vec2(float a, float b) : x(a), y(b) {}
	vec2(const vec2& o) : x(o.x), y(o.y) {}
	vec2& operator=(const vec2& o) { x = o.x; y = o.y; return *this; }
	float& operator[](int i) { return i ? y : x; }
	float operator[](int i) const { return i ? y : x; }
	vec2& operator*=(float s) { x *= s; y *= s; return *this; }
	vec2& operator+=(const vec2& o) { x += o.x; y += o.y; return *this; }
};
inline ivec2::ivec2(const vec2& v) : x((int) v.x), y((int) v.y) {}

struct vec3 {
	union {
		struct { float x, y, z; }; struct { float r, g, b; };
		swz2<vec2, 3, 0, 1> xy; swz2<vec2, 3, 1, 0> yx; swz2<vec2, 3, 1, 2> yz; swz2<vec2, 3, 0, 1> rg; swz3<vec3, 3, 0, 1, 2> xyz; swz3<vec3, 3, 0, 1, 2> rgb;
	};
	vec3() : x(0.0f), y(0.0f), z(0.0f) {}
	explicit vec3(float a) : x(a), y(a), z(a) {}
	vec3(float a, float b_, float c) : x(a), y(b_), z(c) {}
	vec3(const vec2& v, float c) : x(v.x), y(v.y), z(c) {}
	vec3(const ivec2& v, float c) : x((float) v.x), y((float) v.y), z(c) {}
	vec3(const vec3& o) : x(o.x), y(o.y), z(o.z) {}
	vec3& operator=(const vec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
	float& operator[](int i) { return (&x)[i]; }
	float operator[](int i) const { return (&x)[i]; }
	float& operator[](uint i) { return (&x)[i]; }
	float operator[](uint i) const { return (&x)[i]; }
	vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
	vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
	vec3& operator*=(const vec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
	vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
	vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
};

struct vec4 {
	union {
		struct { float x, y, z, w; }; struct { float r, g, b, a; };
		swz2<vec2, 4, 0, 1> xy; swz2<vec2, 4, 2, 3> zw; swz2<vec2, 4, 0, 1> rg; swz2<vec2, 4, 2, 3> ba; swz3<vec3, 4, 0, 1, 2> xyz; swz3<vec3, 4, 0, 1, 2> rgb;
	};
	vec4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
	vec4(float a_, float b_, float c, float d) : x(a_), y(b_), z(c), w(d) {}
	vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
	vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
	vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
	float& operator[](int i) { return (&x)[i]; }
	float operator[](int i) const { return (&x)[i]; }
};

// ---- component-wise arithmetic
inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(vec2 a, vec2 b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator*(vec2 a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, vec2 a) { return vec2(s * a.x, s * a.y); }
inline vec2 operator/(vec2 a, float s) { return vec2(a.x / s, a.y / s); }
inline vec2 operator-(float s, vec2 a) { return vec2(s - a.x, s - a.y); }
inline vec2 operator-(vec2 a) { return vec2(-a.x, -a.y); }
inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec4 operator*(vec4 a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }

// ---- scalar built-ins (hide <cmath>)
inline float abs(float x) { return std::fabs(x); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline float max(float a, float b) { return std::fmax(a, b); }
inline float min(float a, float b) { return std::fmin(a, b); }
inline float clamp(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
inline float fma(float a, float b, float c) { return std::fmaf(a, b, c); }
// transcendental functions: GLSL prescribes no rounding, the oracle defines them as the correctly rounded fp32 value
// (double-precision libm rounded once; see glsl_atan in risltc_oracle.c)
inline float sin(float x) { return (float) std::sin((double) x); }
inline float cos(float x) { return (float) std::cos((double) x); }
inline float atan(float x) { return (float) std::atan((double) x); }
inline float atan(float y, float x) { return (float) std::atan2((double) y, (double) x); }
inline float acos(float x) { return (float) std::acos((double) x); }
inline float asin(float x) { return (float) std::asin((double) x); }
inline float tan(float x) { return (float) std::tan((double) x); }
inline float pow(float a, float b) { return std::pow(a, b); }
inline float log2(float x) { return std::log2(x); }
inline float exp2(float x) { return std::exp2(x); }
inline float floor(float x) { return std::floor(x); }
inline float fract(float x) { return x - std::floor(x); }
inline float sign(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }
inline float step(float edge, float x) { return (x < edge) ? 0.0f : 1.0f; }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline bool isnan(float x) { return std::isnan(x); }
inline bool isinf(float x) { return std::isinf(x); }
inline uint floatBitsToUint(float f) { uint u; std::memcpy(&u, &f, 4); return u; }
inline float uintBitsToFloat(uint u) { float f; std::memcpy(&f, &u, 4); return f; }

// ---- vector built-ins
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(vec4 a, vec4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(vec2 a) { return std::sqrt(dot(a, a)); }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec2 normalize(vec2 a) { return a * inversesqrt(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a * inversesqrt(dot(a, a)); }
inline vec2 abs(vec2 a) { return vec2(std::fabs(a.x), std::fabs(a.y)); }
inline vec3 abs(vec3 a) { return vec3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
inline vec3 max(vec3 a, vec3 b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline vec3 min(vec3 a, vec3 b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 clamp(vec3 v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline vec2 fma(vec2 a, vec2 b, vec2 c) { return vec2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
inline vec3 fma(vec3 a, vec3 b, vec3 c) { return vec3(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y), fma(a.z, b.z, c.z)); }
inline vec3 mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
inline vec2 mix(vec2 a, vec2 b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 sqrt(vec3 a) { return vec3(std::sqrt(a.x), std::sqrt(a.y), std::sqrt(a.z)); }

// ---- matrices, column-major like GLSL: m[column][row]
struct mat2 {
	vec2 c[2];
	mat2() {}
	mat2(vec2 a, vec2 b) { c[0] = a; c[1] = b; }
	vec2& operator[](int i) { return c[i]; }
	const vec2& operator[](int i) const { return c[i]; }
	mat2& operator-=(const mat2& o) { c[0] = c[0] - o.c[0]; c[1] = c[1] - o.c[1]; return *this; }
};
inline mat2 operator-(const mat2& a, const mat2& b) { return mat2(a.c[0] - b.c[0], a.c[1] - b.c[1]); }
inline mat2 operator-(const mat2& a) { return mat2(-a.c[0], -a.c[1]); }
inline mat2 transpose(const mat2& m) { return mat2(vec2(m.c[0].x, m.c[1].x), vec2(m.c[0].y, m.c[1].y)); }
inline float determinant(const mat2& m) { return m.c[0].x * m.c[1].y - m.c[1].x * m.c[0].y; }
inline mat2 outerProduct(vec2 col, vec2 row) { return mat2(col * row.x, col * row.y); }

struct mat3 {
	vec3 c[3];
	mat3() {}
	mat3(vec3 a, vec3 b, vec3 d) { c[0] = a; c[1] = b; c[2] = d; }
	mat3(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2) { c[0] = vec3(a0, a1, a2); c[1] = vec3(b0, b1, b2); c[2] = vec3(c0, c1, c2); }
	vec3& operator[](int i) { return c[i]; }
	const vec3& operator[](int i) const { return c[i]; }
	vec3& operator[](uint i) { return c[i]; }
	const vec3& operator[](uint i) const { return c[i]; }
};
inline vec3 operator*(const mat3& m, vec3 v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z; }
inline mat3 operator-(const mat3& m) { return mat3(-m.c[0], -m.c[1], -m.c[2]); }
inline mat3 transpose(const mat3& m) { return mat3(vec3(m.c[0].x, m.c[1].x, m.c[2].x), vec3(m.c[0].y, m.c[1].y, m.c[2].y), vec3(m.c[0].z, m.c[1].z, m.c[2].z)); }
inline mat3 operator*(const mat3& a, const mat3& b) { return mat3(a * b.c[0], a * b.c[1], a * b.c[2]); }
inline float determinant(const mat3& m) { return dot(m.c[0], cross(m.c[1], m.c[2])); }

struct mat3x4 {   // 3 columns of vec4
	vec4 c[3];
	vec4& operator[](int i) { return c[i]; }
};
struct mat4x3 {   // 4 columns of vec3
	vec3 c[4];
	mat4x3() {}
	mat4x3(vec3 a, vec3 b, vec3 d, vec3 e) { c[0] = a; c[1] = b; c[2] = d; c[3] = e; }
	vec3& operator[](int i) { return c[i]; }
	const vec3& operator[](int i) const { return c[i]; }
	vec3& operator[](uint i) { return c[i]; }
	const vec3& operator[](uint i) const { return c[i]; }
};
inline vec3 operator*(const mat4x3& m, vec4 v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z + m.c[3] * v.w; }
inline mat4x3 operator*(const mat3& a, const mat4x3& b) { return mat4x3(a * b.c[0], a * b.c[1], a * b.c[2], a * b.c[3]); }
inline mat3x4 transpose(const mat4x3& m) {
	mat3x4 r;
	r.c[0] = vec4(m.c[0].x, m.c[1].x, m.c[2].x, m.c[3].x);
	r.c[1] = vec4(m.c[0].y, m.c[1].y, m.c[2].y, m.c[3].y);
	r.c[2] = vec4(m.c[0].z, m.c[1].z, m.c[2].z, m.c[3].z);
	return r;
}
inline vec4 operator*(const mat3x4& m, vec3 v) {
	return vec4(m.c[0].x * v.x + m.c[1].x * v.y + m.c[2].x * v.z, m.c[0].y * v.x + m.c[1].y * v.y + m.c[2].y * v.z,
	            m.c[0].z * v.x + m.c[1].z * v.y + m.c[2].z * v.z, m.c[0].w * v.x + m.c[1].w * v.y + m.c[2].w * v.z);
}
struct mat4 {
	vec4 c[4];
	vec4& operator[](int i) { return c[i]; }
};
inline vec4 operator*(const mat4& m, vec4 v) {
	vec4 r;
	for (int i = 0; i != 4; ++i) r[i] = m.c[0][i] * v.x + m.c[1][i] * v.y + m.c[2][i] * v.z + m.c[3][i] * v.w;
	return r;
}

// ---- resources. Texel buffers and input attachments read what the harness points them at; the
// LTC array uses the oracle's definition of the sampler (exact fp32 bilinear, nearest layer).
struct utextureBuffer { const void* data; int bytes_per_texel; };
struct textureBuffer { const uint16_t* data; };
struct usubpassInput { uint value; };
// a material texture: either flat (texel) or a texture object of the caller that the caller's sampler reads (the oracle's
// definition of textureGrad, orc_sample_texture_grad: texture filtering is driver code, not in the tree)
typedef void (*sample_texture_fn)(const void* texture, const float* uv, const float* ddx, const float* ddy, float* rgba);
extern sample_texture_fn g_sample_texture;
struct sampler2D { float texel[4]; const void* texture; };
struct sampler2DArray { const uint16_t* data; int channels, res, layers; };
struct accelerationStructureEXT { int unused; };
struct rayQueryEXT { bool hit; vec3 origin, direction; float t_min, t_max; };

inline uvec4 texelFetch(const utextureBuffer& b, int i) {
	if (b.bytes_per_texel == 8) { const uint* p = (const uint*) b.data; return uvec4(p[2 * (size_t) i], p[2 * (size_t) i + 1], 0, 1); }
	return uvec4(((const uint8_t*) b.data)[i], 0, 0, 1);
}
inline vec4 texelFetch(const textureBuffer& b, int i) {
	const uint16_t* p = b.data + 4 * (size_t) i;
	return vec4((float) p[0] / 65535.0f, (float) p[1] / 65535.0f, (float) p[2] / 65535.0f, (float) p[3] / 65535.0f);
}
inline uvec4 subpassLoad(const usubpassInput& s) { return uvec4(s.value, 0, 0, 0); }
inline vec4 textureGrad(const sampler2D& s, vec2 uv, vec2 ddx, vec2 ddy) {
	if (!s.texture) return vec4(s.texel[0], s.texel[1], s.texel[2], s.texel[3]);
	float c[2] = { uv.x, uv.y }, dx[2] = { ddx.x, ddx.y }, dy[2] = { ddy.x, ddy.y }, out[4];
	g_sample_texture(s.texture, c, dx, dy, out);
	return vec4(out[0], out[1], out[2], out[3]);
}
inline vec4 textureLod(const sampler2DArray& s, vec3 coord, float) {
	int res = s.res;
	int layer = (int) std::floor(coord.z + 0.5f);
	if (layer < 0) layer = 0;
	if (layer > s.layers - 1) layer = s.layers - 1;
	float x = coord.x * (float) res - 0.5f, y = coord.y * (float) res - 0.5f;
	float fx0 = std::floor(x), fy0 = std::floor(y), fx = x - fx0, fy = y - fy0;
	int x0 = (int) fx0, y0 = (int) fy0, x1 = x0 + 1, y1 = y0 + 1;
	auto cl = [res](int v) { return v < 0 ? 0 : (v > res - 1 ? res - 1 : v); };
	x0 = cl(x0); x1 = cl(x1); y0 = cl(y0); y1 = cl(y1);
	float w00 = (1.0f - fx) * (1.0f - fy), w10 = fx * (1.0f - fy), w01 = (1.0f - fx) * fy, w11 = fx * fy;
	size_t base = (size_t) layer * res * res;
	size_t i00 = base + (size_t) y0 * res + x0, i10 = base + (size_t) y0 * res + x1, i01 = base + (size_t) y1 * res + x0, i11 = base + (size_t) y1 * res + x1;
	vec4 r(0.0f, 0.0f, 0.0f, 1.0f);
	for (int ch = 0; ch != s.channels; ++ch)
		r[ch] = w00 * ((float) s.data[i00 * s.channels + ch] / 65535.0f) + w10 * ((float) s.data[i10 * s.channels + ch] / 65535.0f)
		      + w01 * ((float) s.data[i01 * s.channels + ch] / 65535.0f) + w11 * ((float) s.data[i11 * s.channels + ch] / 65535.0f);
	return r;
}
#define nonuniformEXT(x) (x)

// ---- ray queries: forwarded to the any-hit callback the harness was given (the oracle's BVH)
enum { gl_RayFlagsTerminateOnFirstHitEXT = 4, gl_RayFlagsOpaqueEXT = 1, gl_RayFlagsSkipClosestHitShaderEXT = 8,
       gl_RayQueryCommittedIntersectionNoneEXT = 0, gl_RayQueryCommittedIntersectionTriangleEXT = 1 };
typedef int (*any_hit_fn)(const void* scene, const float* origin, const float* dir, float t_min, float t_max);
extern any_hit_fn g_any_hit;
extern const void* g_any_hit_scene;
extern thread_local unsigned long long g_ray_count;
inline void rayQueryInitializeEXT(rayQueryEXT& q, const accelerationStructureEXT&, int, int, vec3 o, float t_min, vec3 d, float t_max) {
	q.hit = false; q.origin = o; q.direction = d; q.t_min = t_min; q.t_max = t_max;
}
inline bool rayQueryProceedEXT(rayQueryEXT& q) {
	float o[3] = { q.origin.x, q.origin.y, q.origin.z }, d[3] = { q.direction.x, q.direction.y, q.direction.z };
	++g_ray_count;
	q.hit = g_any_hit(g_any_hit_scene, o, d, q.t_min, q.t_max) != 0;
	return false;
}
inline int rayQueryGetIntersectionTypeEXT(const rayQueryEXT& q, bool) { return q.hit ? gl_RayQueryCommittedIntersectionTriangleEXT : gl_RayQueryCommittedIntersectionNoneEXT; }

}  // namespace glsl
