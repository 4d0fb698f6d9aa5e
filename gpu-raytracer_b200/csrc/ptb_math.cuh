// Small vector algebra + bit tricks for the path tracing kernels (device only).
//
// The arithmetic contract: results must match the reference's kernels, which are compiled with
// --use_fast_math (approximate division / sqrt / rsqrt, FMA contraction) on top of NVIDIA's helper_math
// conventions (Src/CUDA/cudart/cuda_math.h: component-wise operators, normalize = v * rsqrtf(dot(v,v)),
// float3 / float = three divisions).  Expression *trees* therefore follow the reference formulas; the
// scheduling, memory layout and kernel structure around them are ours.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PTB_PI 3.14159265359f
#define PTB_ONE_OVER_PI 0.31830988618f
#define PTB_TWO_PI 6.28318530718f
#define PTB_ONE_OVER_TWO_PI 0.15915494309f
#define PTB_EPSILON 0.0001f
#define PTB_INVALID (-1)
#define PTB_INF (__int_as_float(0x7f800000))

#define PTB_DI __device__ __forceinline__

PTB_DI float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
PTB_DI float3 f3(float s) { return make_float3(s, s, s); }
PTB_DI float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
PTB_DI float4 f4(float3 v) { return make_float4(v.x, v.y, v.z, 0.0f); }
PTB_DI float4 f4(float3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
PTB_DI float4 f4(float s) { return make_float4(s, s, s, s); }
PTB_DI float2 f2(float x, float y) { return make_float2(x, y); }

PTB_DI float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
PTB_DI float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
PTB_DI float3 operator-(float3 a, float b) { return f3(a.x - b, a.y - b, a.z - b); }
PTB_DI float3 operator+(float3 a, float b) { return f3(a.x + b, a.y + b, a.z + b); }
PTB_DI float3 operator-(float b, float3 a) { return f3(b - a.x, b - a.y, b - a.z); }
PTB_DI float3 operator+(float b, float3 a) { return f3(b + a.x, b + a.y, b + a.z); }
PTB_DI float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
PTB_DI float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
PTB_DI float3 operator*(float3 a, float b) { return f3(a.x * b, a.y * b, a.z * b); }
PTB_DI float3 operator*(float b, float3 a) { return f3(b * a.x, b * a.y, b * a.z); }
PTB_DI float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
PTB_DI float3 operator/(float3 a, float b) { return f3(a.x / b, a.y / b, a.z / b); }
PTB_DI float3 operator/(float b, float3 a) { return f3(b / a.x, b / a.y, b / a.z); }
PTB_DI void operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
PTB_DI void operator-=(float3& a, float3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
PTB_DI void operator*=(float3& a, float3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; }
PTB_DI void operator*=(float3& a, float b) { a.x *= b; a.y *= b; a.z *= b; }
PTB_DI void operator/=(float3& a, float b) { a.x /= b; a.y /= b; a.z /= b; }

PTB_DI float2 operator+(float2 a, float2 b) { return f2(a.x + b.x, a.y + b.y); }
PTB_DI float2 operator-(float2 a, float2 b) { return f2(a.x - b.x, a.y - b.y); }
PTB_DI float2 operator*(float2 a, float b) { return f2(a.x * b, a.y * b); }
PTB_DI float2 operator*(float b, float2 a) { return f2(b * a.x, b * a.y); }

PTB_DI float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
PTB_DI float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
PTB_DI float4 operator*(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
PTB_DI float4 operator*(float4 a, float b) { return make_float4(a.x * b, a.y * b, a.z * b, a.w * b); }
PTB_DI float4 operator*(float b, float4 a) { return make_float4(b * a.x, b * a.y, b * a.z, b * a.w); }
PTB_DI float4 operator/(float4 a, float b) { return make_float4(a.x / b, a.y / b, a.z / b, a.w / b); }
PTB_DI void operator+=(float4& a, float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

PTB_DI float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PTB_DI float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
PTB_DI float3 cross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PTB_DI float length(float3 v) { return sqrtf(dot(v, v)); }
PTB_DI float3 normalize(float3 v) { float inv = rsqrtf(dot(v, v)); return v * inv; }
PTB_DI float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
PTB_DI float square(float x) { return x * x; }
PTB_DI float safe_sqrt(float x) { return sqrtf(fmaxf(0.0f, x)); }
PTB_DI float3 safe_sqrt(float3 v) { return f3(safe_sqrt(v.x), safe_sqrt(v.y), safe_sqrt(v.z)); }
PTB_DI float abs_dot(float3 a, float3 b) { return fabsf(dot(a, b)); }
PTB_DI float sign1(float x) { return copysignf(1.0f, x); }
PTB_DI float luminance(float r, float g, float b) { return 0.299f * r + 0.587f * g + 0.114f * b; }
// NVIDIA helper_math lerp (a + t*(b-a)): in the reference the non-template helper_math overload wins over Util.h's template
PTB_DI float lerpf(float a, float b, float t) { return a + t * (b - a); }
PTB_DI float2 sincos2(float x) { float s, c; __sincosf(x, &s, &c); return f2(s, c); }
template <typename T> PTB_DI T barycentric(float u, float v, T base, T e1, T e2) { return base + u * e1 + v * e2; }

// integer-ordered min/max on float bit patterns (reference: vmin/vmax.s32 video instructions, Util.h:305-341)
PTB_DI float imax3(float a, float b, float c) { return __int_as_float(max(max(__float_as_int(a), __float_as_int(b)), __float_as_int(c))); }
PTB_DI float imin3(float a, float b, float c) { return __int_as_float(min(min(__float_as_int(a), __float_as_int(b)), __float_as_int(c))); }
PTB_DI float imin_max(float a, float b, float c) { return __int_as_float(max(min(__float_as_int(a), __float_as_int(b)), __float_as_int(c))); } // max(min(a,b),c)
PTB_DI float imax_min(float a, float b, float c) { return __int_as_float(min(max(__float_as_int(a), __float_as_int(b)), __float_as_int(c))); } // min(max(a,b),c)

PTB_DI unsigned byte_of(unsigned x, unsigned i) { return (x >> (i * 8)) & 0xffu; }
PTB_DI unsigned msb(unsigned x) { return 31u - __clz(x); }            // x != 0
PTB_DI unsigned sign_extend_s8x4(unsigned x) { unsigned r; asm("prmt.b32 %0, %1, 0x0, 0x0000ba98;" : "=r"(r) : "r"(x)); return r; } // byte -> 0xff if its MSB is set, else 0x00

PTB_DI void orthonormal_basis(float3 n, float3& t, float3& b) {
    float sg = copysignf(1.0f, n.z);
    float a = -1.0f / (sg + n.z);
    float bb = n.x * n.y * a;
    t = f3(1.0f + sg * n.x * n.x * a, sg * bb, -sg * n.x);
    b = f3(bb, sg + n.y * n.y * a, -n.y);
}
PTB_DI float3 local_to_world(float3 v, float3 t, float3 b, float3 n) {
    return f3(t.x * v.x + b.x * v.y + n.x * v.z, t.y * v.x + b.y * v.y + n.y * v.z, t.z * v.x + b.z * v.y + n.z * v.z);
}
PTB_DI float3 world_to_local(float3 v, float3 t, float3 b, float3 n) { return f3(dot(t, v), dot(b, v), dot(n, v)); }

struct Mat3x4 { float4 r0, r1, r2; };
PTB_DI float3 xform_pos(const Mat3x4& m, float3 p) {
    return f3(m.r0.x * p.x + m.r0.y * p.y + m.r0.z * p.z + m.r0.w,
              m.r1.x * p.x + m.r1.y * p.y + m.r1.z * p.z + m.r1.w,
              m.r2.x * p.x + m.r2.y * p.y + m.r2.z * p.z + m.r2.w);
}
PTB_DI float3 xform_dir(const Mat3x4& m, float3 d) {
    return f3(m.r0.x * d.x + m.r0.y * d.y + m.r0.z * d.z,
              m.r1.x * d.x + m.r1.y * d.y + m.r1.z * d.z,
              m.r2.x * d.x + m.r2.y * d.y + m.r2.z * d.z);
}
PTB_DI Mat3x4 load_mat(const float4* base, int id) {
    Mat3x4 m; m.r0 = __ldg(base + 3 * id); m.r1 = __ldg(base + 3 * id + 1); m.r2 = __ldg(base + 3 * id + 2); return m;
}

PTB_DI float3 ray_origin_epsilon_offset(float3 origin, float3 direction, float3 gn) {
    return origin + sign1(dot(direction, gn)) * PTB_EPSILON * gn;
}
PTB_DI float3 reflect_direction(float3 d, float3 n) { return 2.0f * dot(d, n) * n - d; }
PTB_DI float3 refract_direction(float3 d, float3 n, float eta) {
    float c = dot(d, n);
    float k = 1.0f - eta * eta * (1.0f - square(c));
    return (eta * c - safe_sqrt(k)) * n - eta * d;
}

// hashes (Util.h:104-149)
PTB_DI unsigned pcg_hash(unsigned seed) {
    unsigned state = seed * 747796405u + 2891336453u;
    unsigned word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}
PTB_DI unsigned hash_with(unsigned seed, unsigned hash) {
    seed = (seed ^ 61u) ^ hash; seed += seed << 3; seed ^= seed >> 4; seed *= 0x27d4eb2du; return seed;
}
PTB_DI unsigned permute_index(unsigned index, unsigned length, unsigned seed) {
    unsigned mask = length - 1;
    index ^= seed; index *= 0xe170893d; index ^= seed >> 16; index ^= (index & mask) >> 4; index ^= seed >> 8;
    index *= 0x0929eb3f; index ^= seed >> 23; index ^= (index & mask) >> 1; index *= 1 | seed >> 27;
    index *= 0x6935fa69; index ^= (index & mask) >> 11; index *= 0x74dcb303; index ^= (index & mask) >> 2;
    index *= 0x9e501cc3; index ^= (index & mask) >> 2; index *= 0xc860a3df; index &= mask; index ^= index >> 5;
    return (index + seed) & mask;
}
