// Device-side building blocks shared by the kernels: RNG, camera, triangle tests, CWBVH node test,
// light sampling, microfacet helpers.  Formulas follow the reference files cited per function.
#pragma once
#include "ptb_math.cuh"
#include "ptb_types.cuh"

// ------------------------------------------------------------------------------------------ RNG
// PMJ02 points + blue-noise Cranley-Patterson shift, keyed by (pixel_index, dimension, bounce, sample_index).
// Reference: Src/CUDA/Sampling.h:30-84.
enum SampleDim { DIM_FILTER = 0, DIM_APERTURE, DIM_RUSSIAN_ROULETTE, DIM_NEE_LIGHT, DIM_NEE_TRIANGLE, DIM_BSDF_0, DIM_BSDF_1, DIM_COUNT, DIM_PER_BOUNCE = 5 };

template <int Dim>
PTB_DI float2 rng2(const Frame& P, unsigned pixel_index, unsigned bounce, unsigned sample_index) {
    unsigned hash = pcg_hash((pixel_index * unsigned(DIM_COUNT) + unsigned(Dim)) * PTB_MAX_BOUNCES + bounce);
    if (sample_index >= PTB_PMJ_SAMPLES) {
        const float one_over_max = __uint_as_float(0x2f7fffffu);
        float x = hash_with(sample_index, hash) * one_over_max;
        float y = hash_with(sample_index + 0xdeadbeefu, hash) * one_over_max;
        return f2(x, y);
    }
    unsigned dim = unsigned(Dim) + unsigned(DIM_PER_BOUNCE) * bounce;
    if (dim >= PTB_PMJ_SEQUENCES) sample_index = permute_index(sample_index, PTB_PMJ_SAMPLES, hash);
    const float2* seq = P.pmj + (dim % PTB_PMJ_SEQUENCES) * PTB_PMJ_SAMPLES;
    float2 s = seq[sample_index];
    const uchar2* tile = P.blue_noise + (dim % PTB_BLUE_NOISE_TEXTURES) * (PTB_BLUE_NOISE_DIM * PTB_BLUE_NOISE_DIM);
    int x = (pixel_index % P.pitch) % PTB_BLUE_NOISE_DIM;
    int y = (pixel_index / P.pitch) % PTB_BLUE_NOISE_DIM;
    uchar2 bn = tile[x + y * PTB_BLUE_NOISE_DIM];
    s = s + f2(bn.x * (1.0f / 255.0f), bn.y * (1.0f / 255.0f));
    if (s.x >= 1.0f) s.x -= 1.0f;
    if (s.y >= 1.0f) s.y -= 1.0f;
    return s;
}

PTB_DI bool pdf_is_valid(float pdf) { return isfinite(pdf) && pdf > 1e-4f; }
PTB_DI float power_heuristic(float f, float g) { return (f * f) / (f * f + g * g); }

PTB_DI float sample_tent(float u) { return u < 0.5f ? safe_sqrt(2.0f * u) - 1.0f : 1.0f - safe_sqrt(2.0f - 2.0f * u); }
PTB_DI float2 sample_gaussian(float u1, float u2) { float f = sqrtf(-2.0f * logf(u1)); float a = PTB_TWO_PI * u2; return f * sincos2(a); }
PTB_DI float sample_exp(float lambda, float u) { return -logf(u) / lambda; }
PTB_DI float2 sample_triangle(float u1, float u2) {
    if (u2 > u1) { u1 *= 0.5f; u2 -= u1; } else { u2 *= 0.5f; u1 -= u2; }
    return f2(u1, u2);
}
PTB_DI float2 sample_disk(float u1, float u2) {
    float a = 2.0f * u1 - 1.0f, b = 2.0f * u2 - 1.0f;
    float phi, r;
    if (a * a > b * b) { r = a; phi = 0.25f * PTB_PI * (b / a); }
    else               { r = b; phi = 0.5f * PTB_PI - 0.25f * PTB_PI * (a / b); }
    return r * sincos2(phi);
}
PTB_DI float3 sample_cosine_hemisphere(float u1, float u2) { float2 d = sample_disk(u1, u2); return f3(d.x, d.y, safe_sqrt(1.0f - dot(d, d))); }
PTB_DI float3 spherical_to_cartesian(float st, float ct, float sp, float cp) { return f3(st * sp, st * cp, ct); }
PTB_DI float3 sample_henyey_greenstein(float3 omega, float g, float u1, float u2) {
    float ct;
    if (fabsf(g) < 1e-3f) ct = 1.0f - 2.0f * u1;
    else ct = -(1.0f + g * g - square((1.0f - g * g) / (1.0f + g - 2.0f * g * u1))) / (2.0f * g);
    float st = safe_sqrt(1.0f - square(ct));
    float2 sc = sincos2(PTB_TWO_PI * u2);
    float3 d = spherical_to_cartesian(st, ct, sc.x, sc.y);
    float3 v1, v2; orthonormal_basis(omega, v1, v2);
    return local_to_world(d, v1, v2, omega);
}
// Heitz 2018 visible-normal sampling (Sampling.h:154-178)
PTB_DI float3 sample_vndf_ggx(float3 omega, float ax, float ay, float u1, float u2) {
    float3 v = normalize(f3(ax * omega.x, ay * omega.y, omega.z));
    float lsq = v.x * v.x + v.y * v.y;
    float3 a1 = lsq > 0.0f ? f3(-v.y, v.x, 0.0f) / sqrtf(lsq) : f3(1.0f, 0.0f, 0.0f);
    float3 a2 = cross(v, a1);
    float2 d = sample_disk(u1, u2);
    float t1 = d.x;
    float t2 = lerpf(safe_sqrt(1.0f - t1 * t1), d.y, 0.5f + 0.5f * v.z);
    float3 nh = t1 * a1 + t2 * a2 + safe_sqrt(1.0f - t1 * t1 - t2 * t2) * v;
    return normalize(f3(ax * nh.x, ay * nh.y, nh.z));
}

// ------------------------------------------------------------------------------------------ camera
// Thin-lens ray with reconstruction-filter jitter. Reference: Src/CUDA/Camera.h:20-62.
struct Ray { float3 o, d; };

PTB_DI Ray camera_ray(const Frame& P, int pixel_index, int sample_index, int x, int y) {
    float2 rf = rng2<DIM_FILTER>(P, pixel_index, 0, sample_index);
    float2 ra = rng2<DIM_APERTURE>(P, pixel_index, 0, sample_index);
    float2 jitter;
    if (P.config.enable_svgf) {
        const float hx[4] = { 0.3f, 0.7f, 0.2f, 0.8f };
        const float hy[4] = { 0.2f, 0.8f, 0.7f, 0.3f };
        jitter.x = hx[sample_index & 3]; jitter.y = hy[sample_index & 3];
    } else if (P.config.reconstruction_filter == 0) {
        jitter = rf;
    } else if (P.config.reconstruction_filter == 1) {
        jitter.x = sample_tent(rf.x); jitter.y = sample_tent(rf.y);
    } else {
        float2 g = sample_gaussian(rf.x, rf.y);
        jitter.x = 0.5f + 0.5f * g.x; jitter.y = 0.5f + 0.5f * g.y;
    }
    float xj = float(x) + jitter.x, yj = float(y) + jitter.y;
    float3 pos = f3(P.camera.position[0], P.camera.position[1], P.camera.position[2]);
    float3 blc = f3(P.camera.bottom_left_corner[0], P.camera.bottom_left_corner[1], P.camera.bottom_left_corner[2]);
    float3 xa = f3(P.camera.x_axis[0], P.camera.x_axis[1], P.camera.x_axis[2]);
    float3 ya = f3(P.camera.y_axis[0], P.camera.y_axis[1], P.camera.y_axis[2]);
    float3 focal_point = P.camera.focal_distance * normalize(blc + xj * xa + yj * ya);
    float2 lens = P.camera.aperture_radius * sample_disk(ra.x, ra.y);
    float3 offset = xa * lens.x + ya * lens.y;
    Ray r; r.o = pos + offset; r.d = normalize(focal_point - offset);
    return r;
}

// ------------------------------------------------------------------------------------------ triangles
// 96-byte records, Src/CUDA/Raytracing/Triangle.h:4-102.
struct TriPos { float3 p0, e1, e2; };
PTB_DI TriPos load_tri_pos(const Frame& P, int id) {
    const float4* t = P.triangles + 6 * size_t(id);
    float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2);
    TriPos r; r.p0 = f3(a.x, a.y, a.z); r.e1 = f3(a.w, b.x, b.y); r.e2 = f3(b.z, b.w, c.x);
    return r;
}
// merged static BVH: compact 48-byte records, remap in the spare words.  A slot whose instance has started moving is RETIRED by
// poisoning its records once (k_retire_merged_slots: p0.x = NaN, so u is NaN and every comparison of the test fails) -- the
// triangle tests themselves carry no retirement check (a per-hit table lookup here cost 3.7 % of the frame, 27.05 -> 28.05 ms).
PTB_DI TriPos load_tri_pos_flat(const Frame& P, int id, float4& c) {
    const float4* t = P.flat_tris + 3 * size_t(id);
    float4 a = __ldg(t), b = __ldg(t + 1); c = __ldg(t + 2);
    TriPos r; r.p0 = f3(a.x, a.y, a.z); r.e1 = f3(a.w, b.x, b.y); r.e2 = f3(b.z, b.w, c.x);
    return r;
}
struct TriFull { float3 p0, e1, e2, n0, ne1, ne2; float2 t0, te1, te2; };
PTB_DI TriFull load_tri_full(const Frame& P, int id) {
    const float4* t = P.triangles + 6 * size_t(id);
    float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2), d = __ldg(t + 3), e = __ldg(t + 4), g = __ldg(t + 5);
    TriFull r;
    r.p0 = f3(a.x, a.y, a.z); r.e1 = f3(a.w, b.x, b.y); r.e2 = f3(b.z, b.w, c.x);
    r.n0 = f3(c.y, c.z, c.w); r.ne1 = f3(d.x, d.y, d.z); r.ne2 = f3(d.w, e.x, e.y);
    r.t0 = f2(e.z, e.w); r.te1 = f2(g.x, g.y); r.te2 = f2(g.z, g.w);
    return r;
}

struct Hit { float t, u, v; int mesh_id, triangle_id; };
PTB_DI uint4 pack_hit(const Hit& h) {
    unsigned uv = int(h.u * 65535.0f) | (int(h.v * 65535.0f) << 16);
    return make_uint4(h.mesh_id, h.triangle_id, __float_as_uint(h.t), uv);
}
PTB_DI Hit unpack_hit(uint4 w) {
    Hit h; h.mesh_id = w.x; h.triangle_id = w.y; h.t = __uint_as_float(w.z);
    h.u = float(w.w & 0xffff) / 65535.0f; h.v = float(w.w >> 16) / 65535.0f;
    return h;
}

// Woop's test (Woop 2004, "unit triangle"): the ray is mapped by the reference's precomputed affine map, t comes from the z
// components alone and (u, v) are the mapped hit point.  Opt-in (ptb_set_intersector): north_star names it; the reference itself
// uses Moeller-Trumbore, and u, v, t differ from it in the last ulps, so this mode is held to the 1e-4 rel-L2 bar, not to bit parity.
PTB_DI bool woop_test(const Frame& P, int ref, const Ray& ray, float t_max, float& t, float& u, float& v) {
    const float4* m = P.flat_woop + 3 * size_t(ref);
    float4 r2 = __ldg(m + 2);
    float oz = r2.x * ray.o.x + r2.y * ray.o.y + r2.z * ray.o.z + r2.w;
    float dz = r2.x * ray.d.x + r2.y * ray.d.y + r2.z * ray.d.z;
    t = -oz / dz;
    if (!(t > 0.0f && t < t_max)) return false;
    float4 r0 = __ldg(m), r1 = __ldg(m + 1);
    u = (r0.x * ray.o.x + r0.y * ray.o.y + r0.z * ray.o.z + r0.w) + t * (r0.x * ray.d.x + r0.y * ray.d.y + r0.z * ray.d.z);
    v = (r1.x * ray.o.x + r1.y * ray.o.y + r1.z * ray.o.z + r1.w) + t * (r1.x * ray.d.x + r1.y * ray.d.y + r1.z * ray.d.z);
    return u >= 0.0f && v >= 0.0f && u + v <= 1.0f;
}

// Moeller-Trumbore, closest hit (Triangle.h:148-174)
PTB_DI void intersect_triangle(const Frame& P, int mesh_id, int tri_id, const Ray& ray, Hit& hit) {
    float4 c = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const bool flat = mesh_id == PTB_FLAT_MESH;
    if (flat && P.flat_woop) {
        float t, u, v;
        if (woop_test(P, tri_id, ray, hit.t, t, u, v)) {
            int2 who = __ldg(P.flat_who + tri_id);
            hit.t = t; hit.u = u; hit.v = v; hit.mesh_id = -(2 + who.y); hit.triangle_id = who.x;
        }
        return;
    }
    TriPos tr = flat ? load_tri_pos_flat(P, tri_id, c) : load_tri_pos(P, tri_id);
    if (flat) { tri_id = __float_as_int(c.y); mesh_id = -(2 + __float_as_int(c.z)); }   // original triangle id; slot resolved when the hit is stored
    float3 h = cross(ray.d, tr.e2);
    float a = dot(tr.e1, h);
    float f = 1.0f / a;
    float3 s = ray.o - tr.p0;
    float u = f * dot(s, h);
    if (u >= 0.0f && u <= 1.0f) {
        float3 q = cross(s, tr.e1);
        float v = f * dot(ray.d, q);
        if (v >= 0.0f && u + v <= 1.0f) {
            float t = f * dot(tr.e2, q);
            if (t > 0.0f && t < hit.t) { hit.t = t; hit.u = u; hit.v = v; hit.mesh_id = mesh_id; hit.triangle_id = tri_id; }
        }
    }
}
// any hit (Triangle.h:176-198)
PTB_DI bool occludes_triangle(const Frame& P, int mesh_id, int tri_id, const Ray& ray, float max_distance) {
    float4 c;
    const bool flat = mesh_id == PTB_FLAT_MESH;
    if (flat && P.flat_woop) {
        float t, u, v;
        return woop_test(P, tri_id, ray, max_distance, t, u, v);
    }
    TriPos tr = flat ? load_tri_pos_flat(P, tri_id, c) : load_tri_pos(P, tri_id);
    float3 h = cross(ray.d, tr.e2);
    float a = dot(tr.e1, h);
    float f = 1.0f / a;
    float3 s = ray.o - tr.p0;
    float u = f * dot(s, h);
    if (u >= 0.0f && u <= 1.0f) {
        float3 q = cross(s, tr.e1);
        float v = f * dot(ray.d, q);
        if (v >= 0.0f && u + v <= 1.0f) {
            float t = f * dot(tr.e2, q);
            if (t > 0.0f && t < max_distance) return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------------------------------ CWBVH node test
// Src/CUDA/Raytracing/BVH8.h:5-107.  Returns bits 24..31 = hit internal children in traversal priority order,
// bits 0..23 = hit triangle (or instance) slots.
PTB_DI unsigned ray_octant_inv4(float3 d) {
    return (d.x < 0.0f ? 0u : 0x04040404u) | (d.y < 0.0f ? 0u : 0x02020202u) | (d.z < 0.0f ? 0u : 0x01010101u);
}

// byte j of x as a float.  Two exact routes: I2F.U8 (XU pipe, quarter rate -- ncu showed the XU pipe 73 % busy with the 48
// conversions per node) or the 2^23 magic number: PRMT builds the bits of (8388608 + byte), one FADD (FMA pipe) removes the bias.
// Both give the same float bit pattern; PTB_CVT_MAGIC_MASK picks, per use site, which pipe pays (bit 0..5 = xmin,ymin,zmin,xmax,ymax,zmax).
#ifndef PTB_CVT_MAGIC_MASK
#define PTB_CVT_MAGIC_MASK 0x09   // x_min and x_max bytes (16 of 48) through the ALU+FMA route: XU 84 % -> balanced; frame 33.56 -> 32.77 ms (tools/gpu_variants_wave.py)
#endif
template <int SITE>
PTB_DI float byte_to_float(unsigned x, int j) {
    if ((PTB_CVT_MAGIC_MASK >> SITE) & 1) return __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7540u | unsigned(j))) - 8388608.0f;
    return float(byte_of(x, j));
}

PTB_DI unsigned cwbvh_node_intersect(const Ray& ray, unsigned oct_inv4, float max_distance, float4 n0, float4 n1, float4 n2, float4 n3, float4 n4) {
    float3 p = f3(n0.x, n0.y, n0.z);
    unsigned e_imask = __float_as_uint(n0.w);
    float3 adj_inv = f3(__uint_as_float(byte_of(e_imask, 0) << 23) / ray.d.x,
                        __uint_as_float(byte_of(e_imask, 1) << 23) / ray.d.y,
                        __uint_as_float(byte_of(e_imask, 2) << 23) / ray.d.z);
    float3 adj_org = (p - ray.o) / ray.d;
    unsigned hit_mask = 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        unsigned meta4 = __float_as_uint(i == 0 ? n1.z : n1.w);
        unsigned is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        unsigned inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        unsigned bit_index4 = (meta4 ^ (oct_inv4 & inner_mask4)) & 0x1f1f1f1fu;
        unsigned child_bits4 = (meta4 >> 5) & 0x07070707u;
        unsigned qlx = __float_as_uint(i == 0 ? n2.x : n2.y), qhx = __float_as_uint(i == 0 ? n2.z : n2.w);
        unsigned qly = __float_as_uint(i == 0 ? n3.x : n3.y), qhy = __float_as_uint(i == 0 ? n3.z : n3.w);
        unsigned qlz = __float_as_uint(i == 0 ? n4.x : n4.y), qhz = __float_as_uint(i == 0 ? n4.z : n4.w);
        unsigned x_min = ray.d.x < 0.0f ? qhx : qlx, x_max = ray.d.x < 0.0f ? qlx : qhx;
        unsigned y_min = ray.d.y < 0.0f ? qhy : qly, y_max = ray.d.y < 0.0f ? qly : qhy;
        unsigned z_min = ray.d.z < 0.0f ? qhz : qlz, z_max = ray.d.z < 0.0f ? qlz : qhz;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float3 tmin3 = f3(byte_to_float<0>(x_min, j), byte_to_float<1>(y_min, j), byte_to_float<2>(z_min, j));
            float3 tmax3 = f3(byte_to_float<3>(x_max, j), byte_to_float<4>(y_max, j), byte_to_float<5>(z_max, j));
            tmin3 = tmin3 * adj_inv + adj_org;
            tmax3 = tmax3 * adj_inv + adj_org;
            float tmin = imax3(tmin3.x, tmin3.y, fmaxf(tmin3.z, 0.0f));
            float tmax = imin3(tmax3.x, tmax3.y, fminf(tmax3.z, max_distance));
            if (tmin < tmax) hit_mask |= byte_of(child_bits4, j) << byte_of(bit_index4, j);
        }
    }
    return hit_mask;
}

// ------------------------------------------------------------------------------------------ pixel word: pixel | slot << pix_bits | flags
PTB_DI int word_pixel(const Frame& P, unsigned w) { return int(w & ((1u << P.pix_bits) - 1u)); }
PTB_DI int word_slot(const Frame& P, unsigned w) { return int((w & ~PTB_FLAGS_ALL) >> P.pix_bits); }
PTB_DI int word_fb_index(const Frame& P, unsigned w) { return word_slot(P, w) * P.fb_stride + word_pixel(P, w); }

// ------------------------------------------------------------------------------------------ AOV helpers (AOV.h:4-46); px = framebuffer index (slot plane + pixel)
PTB_DI void aov_set(const Frame& P, int k, int px, float4 v) { if (P.aov[k].fb) P.aov[k].fb[px] = v; }
PTB_DI void aov_add(const Frame& P, int k, int px, float4 v) { if (P.aov[k].fb) { float4 c = P.aov[k].fb[px]; c += v; P.aov[k].fb[px] = c; } }
PTB_DI float4 aov_get(const Frame& P, int k, int px) { return P.aov[k].fb[px]; }

// ------------------------------------------------------------------------------------------ materials (Material.h:21-144)
PTB_DI float roughness_to_alpha(float r) { return fmaxf(1e-6f, square(r)); }
#define PTB_ROUGHNESS_CUTOFF 0.05f

PTB_DI float3 sample_sky(const Frame& P, float3 d) {     // Sky.h:7-16
    float phi = atan2f(-d.z, d.x);
    float theta = acosf(clampf(d.y, -1.0f, 1.0f));
    float u = phi * PTB_ONE_OVER_TWO_PI + 0.5f;
    float v = theta * PTB_ONE_OVER_PI;
    return P.sky_scale * f3(tex2D<float4>(P.sky_tex, u, v));
}

PTB_DI int binary_search_cdf(const float* cdf, int first, int last, float value) {  // Util.h:87-102
    int l = first, r = last;
    while (true) {
        int m = (l + r) / 2;
        if (m > first && value <= cdf[m - 1]) r = m - 1;
        else if (value > cdf[m]) l = m + 1;
        else return m;
    }
}
PTB_DI int sample_light(const Frame& P, float u1, float u2, int& transform_id) {   // Sampling.h:180-190
    int lm = binary_search_cdf(P.light_mesh_cdf, 0, P.light_mesh_count - 1, u1);
    transform_id = P.light_mesh_transform_indices[lm];
    int2 span = P.light_mesh_triangle_span[lm];
    int lt = binary_search_cdf(P.light_triangle_cdf, span.x, span.y, u2);
    return P.light_triangle_indices[lt];
}

PTB_DI float fresnel_dielectric(float cos_i, float eta) {
    float sin_o2 = eta * eta * (1.0f - square(cos_i));
    if (sin_o2 >= 1.0f) return 1.0f;
    float cos_o = safe_sqrt(1.0f - sin_o2);
    float p = (eta * cos_i - cos_o) / (eta * cos_i + cos_o);
    float s = (cos_i - eta * cos_o) / (cos_i + eta * cos_o);
    return 0.5f * (p * p + s * s);
}
PTB_DI float3 fresnel_conductor(float cos_i, float3 eta, float3 k) {
    float c2 = square(cos_i);
    float s2 = 1.0f - c2;
    float3 inner = eta * eta - k * k - s2;
    float3 a2b2 = safe_sqrt(inner * inner + 4.0f * k * k * eta * eta);
    float3 a = safe_sqrt(0.5f * (a2b2 + inner));
    float3 x0 = a2b2 + c2, y0 = 2.0f * a * cos_i;
    float3 sq = (x0 - y0) / (x0 + y0);
    float3 x1 = a2b2 * c2 + square(s2), y1 = 2.0f * a * cos_i * s2;
    float3 pq = (x1 - y1) / (x1 + y1) * sq;
    return 0.5f * (pq + sq);
}
PTB_DI float average_fresnel(float ior) { return (ior - 1.0f) / (4.08567f + 1.00071f * ior); }
PTB_DI float3 average_fresnel(float3 eta, float3 k) {
    float3 num = eta * (133.736f - 98.9833f * eta) + k * (eta * (59.5617f - 3.98288f * eta) - 182.37f) + ((0.30818f * eta - 13.1093f) * eta - 62.5919f) * k * k - 8.21474f;
    float3 den = k * (eta * (94.6517f - 15.8558f * eta) - 187.166f) + (-78.476f * eta - 395.268f) * eta + (eta * (eta - 15.4387f) - 62.0752f) * k * k;
    return num / den;
}
PTB_DI float ggx_D(float3 m, float ax, float ay) {
    if (m.z < 1e-6f) return 0.0f;
    float sx = -m.x / (m.z * ax), sy = -m.y / (m.z * ay);
    float sl = 1.0f + sx * sx + sy * sy;
    float c2 = m.z * m.z, c4 = c2 * c2;
    return 1.0f / (sl * sl * PTB_PI * ax * ay * c4);
}
PTB_DI float ggx_lambda(float3 w, float ax, float ay) { return 0.5f * (sqrtf(1.0f + (square(ax * w.x) + square(ay * w.y)) / square(w.z)) - 1.0f); }
PTB_DI float ggx_G1(float3 w, float ax, float ay) { return 1.0f / (1.0f + ggx_lambda(w, ax, ay)); }
PTB_DI float ggx_G2(float3 wo, float3 wi, float3 wm, float ax, float ay) {
    bool bi = dot(wi, wm) * wi.z <= 0.0f, bo = dot(wo, wm) * wo.z <= 0.0f;
    if (bi || bo) return 0.0f;
    return 1.0f / (1.0f + ggx_lambda(wo, ax, ay) + ggx_lambda(wi, ax, ay));
}

// Kulla-Conty lookups (KullaConty.h:12-65)
PTB_DI float remapf(float v, float a, float b, float c, float d) { return c + (v - a) / (b - a) * (d - c); }
PTB_DI float3 fresnel_multiscatter(float3 F_avg, float E_avg) { return F_avg * F_avg * E_avg / (f3(1.0f) - F_avg * (1.0f - E_avg)); }
// NOTE (ptxas 12.9 / sm_100a): a texture fetch whose bindless handle is NOT warp-uniform is compiled to a "waterfall" loop
// (R2UR + BRA.U.ANY) and in that loop ptxas re-materialises the coordinates with an unpredicated MOV into the register that
// is also the TEX destination -- lanes served in an earlier iteration get their RESULT overwritten by a coordinate whenever
// the loop runs more than once.  The reference's kernel_material_dielectric (`(entering ? lut_enter : lut_leave).get(...)`,
// KullaConty.h:16-35) is hit by this in the cubin built here.  We therefore never fetch through a
// lane-dependent handle: both LUTs are fetched through their (kernel-parameter, hence uniform) handles and the result selected.
PTB_DI float dielectric_directional_albedo(const Frame& P, float ior, float rough, float cos_theta, bool entering) {
    ior = remapf(ior, PTB_LUT_DIELECTRIC_MIN_IOR, PTB_LUT_DIELECTRIC_MAX_IOR, 0.0f, 1.0f);
    cos_theta = fabsf(cos_theta);
    float e = tex3D<float>(P.lut_dielectric_dir_enter, ior, rough, cos_theta);
    float l = tex3D<float>(P.lut_dielectric_dir_leave, ior, rough, cos_theta);
    return entering ? e : l;
}
PTB_DI float dielectric_albedo(const Frame& P, float ior, float rough, bool entering) {
    ior = remapf(ior, PTB_LUT_DIELECTRIC_MIN_IOR, PTB_LUT_DIELECTRIC_MAX_IOR, 0.0f, 1.0f);
    float e = tex2D<float>(P.lut_dielectric_enter, ior, rough);
    float l = tex2D<float>(P.lut_dielectric_leave, ior, rough);
    return entering ? e : l;
}
// Fetch through a per-lane texture handle (albedo textures are per material): serve one distinct handle at a time with the
// handle broadcast from the first pending lane, so every hardware fetch sees a warp-uniform handle (see the note above).
template <typename Fetch>
PTB_DI float4 fetch_per_handle(cudaTextureObject_t handle, Fetch fetch) {
    float4 r = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    bool pending = true;
    while (pending) {
        unsigned m = __activemask();
        unsigned long long first = __shfl_sync(m, (unsigned long long)handle, __ffs(m) - 1);
        if (first == (unsigned long long)handle) { r = fetch((cudaTextureObject_t)first); pending = false; }
    }
    return r;
}
PTB_DI float conductor_directional_albedo(const Frame& P, float rough, float cos_theta) { return tex2D<float>(P.lut_conductor_dir, rough, fabsf(cos_theta)); }
PTB_DI float conductor_albedo(const Frame& P, float rough) { return tex1D<float>(P.lut_conductor, rough); }
PTB_DI float kulla_conty_lobe(float E_i, float E_o, float E_avg) { return (1.0f - E_i) * (1.0f - E_o) / fmaxf(0.0001f, PTB_PI * (1.0f - E_avg)); }
PTB_DI float kulla_conty_reciprocity(float E_enter, float E_leave) { return (1.0f - E_leave) / fmaxf(0.0001f, 2.0f - E_enter - E_leave); }

// ------------------------------------------------------------------------------------------ warp-aggregated queue append
// One atomic per warp per queue instead of one per lane (the reference appends with per-lane atomicAdd,
// Pathtracer.cu:294,432,544,729).  `want` may differ per lane; all lanes of the warp must call this.
PTB_DI int warp_append(int* counter, bool want) {
    unsigned mask = __ballot_sync(0xffffffffu, want);
    if (mask == 0) return 0;
    unsigned lane = threadIdx.x & 31u;
    int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == unsigned(leader)) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}
