/* pt_oracle.c -- CPU restatement of the reference's wavefront path-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gpu-raytracer_b200/) links, imports or calls this
 * file; it exists so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg have an independent
 * checker that runs without a GPU.
 *
 * PARITY STATUS: "parity pinned by reference run" -- the reference ships no tests or golden vectors
 * (SURVEY.md section 4), so this restatement is pinned against outputs of the reference's own kernels
 * (Src/CUDA/Pathtracer.cu compiled unmodified into oracle/_ref/, driven by oracle/ref_harness.cpp on a B200):
 * tests/golden/ holds primary-hit tables and framebuffers produced by that run (tests/golden/make_golden.py).
 * The reference arithmetic is --use_fast_math on the GPU (approximate div/sqrt/rsqrt/sin/cos); this file uses
 * IEEE libm, so comparisons against GPU results are tolerance based (hit ids: mismatch rate; images: rel-L2).
 *
 * Per pixel this walks exactly the state machine the wavefront kernels implement, one path per pixel per pass:
 *   kernel_generate            Src/CUDA/Pathtracer.cu:122-139, Camera.h:20-62
 *   random<Dim>                Src/CUDA/Sampling.h:44-84, Util.h:105-149
 *   bvh8_trace / node test     Src/CUDA/Raytracing/BVH8.h:29-274   (closest hit, TLAS -> BLAS)
 *   bvh8_trace_shadow          Src/CUDA/Raytracing/BVH8.h:276-444
 *   bvh2_trace(_shadow)        Src/CUDA/Raytracing/BVH2.h:4-244
 *   triangle_intersect         Src/CUDA/Raytracing/Triangle.h:148-198 (Moeller-Trumbore)
 *   kernel_sort                Src/CUDA/Pathtracer.cu:220-463 (sky, emitters + MIS, russian roulette)
 *   shade_material<Diffuse>    Src/CUDA/Pathtracer.cu:557-757, BSDF.h:8-70
 *   next_event_estimation      Src/CUDA/Pathtracer.cu:465-555, Sampling.h:180-190
 *   kernel_accumulate          Src/CUDA/Pathtracer.cu:775-796, AOV.h:35-46
 * Scope of this restatement: diffuse + emissive materials, sky, NEE + MIS, russian roulette, hit packing with
 * 16-bit barycentrics (Buffers.h:28-47).  Textures are not filtered on the CPU (hardware anisotropic filtering
 * is not reproducible); textured materials use their constant colour and are checked only against oracle/_ref.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PO_PI 3.14159265359f
#define PO_ONE_OVER_PI 0.31830988618f
#define PO_TWO_PI 6.28318530718f
#define PO_ONE_OVER_TWO_PI 0.15915494309f
#define PO_EPSILON 0.0001f
#define PO_MAX_BOUNCES 128
#define PO_INVALID (-1)
#define PO_STACK 64

typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 scl(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline float dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 cross3(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float len3(v3 a) { return sqrtf(dot3(a, a)); }
static inline v3 norm3(v3 a) { return scl(a, 1.0f / sqrtf(dot3(a, a))); }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int32_t f2i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
static inline float i2f(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
/* integer-compare min/max on float bit patterns (Util.h:305-341) */
static inline float imax3(float a, float b, float c) { int32_t x = f2i(a), y = f2i(b), z = f2i(c); int32_t m = x > y ? x : y; return i2f(m > z ? m : z); }
static inline float imin3(float a, float b, float c) { int32_t x = f2i(a), y = f2i(b), z = f2i(c); int32_t m = x < y ? x : y; return i2f(m < z ? m : z); }
static inline float iminmax(float a, float b, float c) { int32_t x = f2i(a), y = f2i(b), z = f2i(c); int32_t m = x < y ? x : y; return i2f(m > z ? m : z); } /* max(min(a,b),c) */
static inline float imaxmin(float a, float b, float c) { int32_t x = f2i(a), y = f2i(b), z = f2i(c); int32_t m = x > y ? x : y; return i2f(m < z ? m : z); } /* min(max(a,b),c) */

typedef struct {
    /* geometry */
    const float*    triangles;      /* T * 24 floats */
    const uint8_t*  bvh_nodes;      /* 80-byte CWBVH nodes or 32-byte binary nodes */
    int             bvh_kind;       /* 8 or 2 */
    const int32_t*  mesh_roots;     /* root | identity << 31 */
    const int32_t*  mesh_material_ids;
    const float*    mesh_transforms;     /* M * 12 */
    const float*    mesh_transforms_inv; /* M * 12 */
    /* materials */
    const int8_t*   material_types;
    const float*    materials;      /* K * 8 floats */
    /* lights */
    float           lights_total_weight;
    const int32_t*  light_triangle_indices;
    const float*    light_triangle_cdf;
    int             light_mesh_count;
    const float*    light_mesh_cdf;
    const int32_t*  light_mesh_triangle_span; /* pairs */
    const int32_t*  light_mesh_transform_indices;
    /* sky */
    const float*    sky;            /* H * W * 4 */
    int             sky_width, sky_height;
    float           sky_scale;
    /* rng tables */
    const float*    pmj;            /* 64 * 4096 * 2 */
    const uint8_t*  blue_noise;     /* 16 * 128 * 128 * 2 */
    /* camera: position, bottom_left, x_axis, y_axis, spread, aperture, focal */
    float           camera[15];
    int             width, height, pitch;
} po_scene;

typedef struct {
    int   reconstruction_filter;    /* 0 box, 1 tent, 2 gaussian */
    int   num_bounces;
    int   enable_nee, enable_mis, enable_rr;
} po_config;

/* ------------------------------------------------------------------ RNG (Sampling.h:44-84) */
static inline uint32_t pcg_hash(uint32_t seed) {
    uint32_t state = seed * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}
static inline uint32_t hash_with(uint32_t seed, uint32_t hash) {
    seed = (seed ^ 61u) ^ hash; seed += seed << 3; seed ^= seed >> 4; seed *= 0x27d4eb2du; return seed;
}
static uint32_t permute(uint32_t index, uint32_t length, uint32_t seed) {
    uint32_t mask = length - 1;
    index ^= seed; index *= 0xe170893d; index ^= seed >> 16; index ^= (index & mask) >> 4; index ^= seed >> 8;
    index *= 0x0929eb3f; index ^= seed >> 23; index ^= (index & mask) >> 1; index *= 1 | seed >> 27;
    index *= 0x6935fa69; index ^= (index & mask) >> 11; index *= 0x74dcb303; index ^= (index & mask) >> 2;
    index *= 0x9e501cc3; index ^= (index & mask) >> 2; index *= 0xc860a3df; index &= mask; index ^= index >> 5;
    return (index + seed) & mask;
}
enum { DIM_FILTER, DIM_APERTURE, DIM_RR, DIM_NEE_LIGHT, DIM_NEE_TRIANGLE, DIM_BSDF_0, DIM_BSDF_1, DIM_COUNT, DIM_PER_BOUNCE = 5 };

static void random2(const po_scene* s, int dim_id, uint32_t pixel_index, uint32_t bounce, uint32_t sample_index, float out[2]) {
    uint32_t hash = pcg_hash((pixel_index * (uint32_t)DIM_COUNT + (uint32_t)dim_id) * PO_MAX_BOUNCES + bounce);
    if (sample_index >= 4096u) {
        const float k = u2f(0x2f7fffffu);
        out[0] = (float)hash_with(sample_index, hash) * k;
        out[1] = (float)hash_with(sample_index + 0xdeadbeefu, hash) * k;
        return;
    }
    uint32_t dim = (uint32_t)dim_id + (uint32_t)DIM_PER_BOUNCE * bounce;
    if (dim >= 64u) sample_index = permute(sample_index, 4096u, hash);
    const float* seq = s->pmj + (size_t)(dim % 64u) * 4096u * 2u;
    float sx = seq[2 * sample_index], sy = seq[2 * sample_index + 1];
    const uint8_t* bn = s->blue_noise + (size_t)(dim % 16u) * 128u * 128u * 2u;
    int x = (int)((pixel_index % (uint32_t)s->pitch) % 128u);
    int y = (int)((pixel_index / (uint32_t)s->pitch) % 128u);
    sx += (float)bn[2 * (x + y * 128)] * (1.0f / 255.0f);
    sy += (float)bn[2 * (x + y * 128) + 1] * (1.0f / 255.0f);
    if (sx >= 1.0f) sx -= 1.0f;
    if (sy >= 1.0f) sy -= 1.0f;
    out[0] = sx; out[1] = sy;
}

static inline float safe_sqrt(float x) { return sqrtf(fmaxf(0.0f, x)); }
static inline float sample_tent(float u) { return u < 0.5f ? safe_sqrt(2.0f * u) - 1.0f : 1.0f - safe_sqrt(2.0f - 2.0f * u); }
static void sample_disk(float u1, float u2, float out[2]) {
    float a = 2.0f * u1 - 1.0f, b = 2.0f * u2 - 1.0f, phi, r;
    if (a * a > b * b) { r = a; phi = 0.25f * PO_PI * (b / a); }
    else               { r = b; phi = 0.5f * PO_PI - 0.25f * PO_PI * (a / b); }
    out[0] = r * sinf(phi); out[1] = r * cosf(phi);
}

/* ------------------------------------------------------------------ camera (Camera.h:20-62) */
typedef struct { v3 o, d; } ray_t;

static ray_t camera_ray(const po_scene* s, const po_config* cfg, int pixel_index, int sample_index, int x, int y) {
    float rf[2], ra[2], jx = 0.5f, jy = 0.5f;
    random2(s, DIM_FILTER, pixel_index, 0, sample_index, rf);
    random2(s, DIM_APERTURE, pixel_index, 0, sample_index, ra);
    if (cfg->reconstruction_filter == 0) { jx = rf[0]; jy = rf[1]; }
    else if (cfg->reconstruction_filter == 1) { jx = sample_tent(rf[0]); jy = sample_tent(rf[1]); }
    else {
        float f = sqrtf(-2.0f * logf(rf[0])), a = PO_TWO_PI * rf[1];
        jx = 0.5f + 0.5f * (f * sinf(a)); jy = 0.5f + 0.5f * (f * cosf(a));
    }
    const float* c = s->camera;
    v3 pos = V(c[0], c[1], c[2]), blc = V(c[3], c[4], c[5]), xa = V(c[6], c[7], c[8]), ya = V(c[9], c[10], c[11]);
    float aperture = c[13], focal = c[14];
    float xj = (float)x + jx, yj = (float)y + jy;
    v3 focal_point = scl(norm3(add(add(blc, scl(xa, xj)), scl(ya, yj))), focal);
    float d2[2]; sample_disk(ra[0], ra[1], d2);
    float lx = aperture * d2[0], ly = aperture * d2[1];
    v3 offset = add(scl(xa, lx), scl(ya, ly));
    ray_t r; r.o = add(pos, offset); r.d = norm3(sub(focal_point, offset));
    return r;
}

/* ------------------------------------------------------------------ triangles / transforms */
typedef struct { float t, u, v; int mesh_id, triangle_id; } hit_t;

static inline void tri_pos(const po_scene* s, int id, v3* p0, v3* e1, v3* e2) {
    const float* t = s->triangles + (size_t)id * 24;
    *p0 = V(t[0], t[1], t[2]); *e1 = V(t[3], t[4], t[5]); *e2 = V(t[6], t[7], t[8]);
}
static inline v3 xf_pos(const float* m, v3 p) {
    return V(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7], m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
}
static inline v3 xf_dir(const float* m, v3 d) {
    return V(m[0] * d.x + m[1] * d.y + m[2] * d.z, m[4] * d.x + m[5] * d.y + m[6] * d.z, m[8] * d.x + m[9] * d.y + m[10] * d.z);
}

static inline void tri_closest(const po_scene* s, int mesh_id, int tri_id, const ray_t* r, hit_t* h) {
    v3 p0, e1, e2; tri_pos(s, tri_id, &p0, &e1, &e2);
    v3 hh = cross3(r->d, e2);
    float a = dot3(e1, hh), f = 1.0f / a;
    v3 sv = sub(r->o, p0);
    float u = f * dot3(sv, hh);
    if (u >= 0.0f && u <= 1.0f) {
        v3 q = cross3(sv, e1);
        float v = f * dot3(r->d, q);
        if (v >= 0.0f && u + v <= 1.0f) {
            float t = f * dot3(e2, q);
            if (t > 0.0f && t < h->t) { h->t = t; h->u = u; h->v = v; h->mesh_id = mesh_id; h->triangle_id = tri_id; }
        }
    }
}
static inline int tri_any(const po_scene* s, int tri_id, const ray_t* r, float max_distance) {
    v3 p0, e1, e2; tri_pos(s, tri_id, &p0, &e1, &e2);
    v3 hh = cross3(r->d, e2);
    float a = dot3(e1, hh), f = 1.0f / a;
    v3 sv = sub(r->o, p0);
    float u = f * dot3(sv, hh);
    if (u >= 0.0f && u <= 1.0f) {
        v3 q = cross3(sv, e1);
        float v = f * dot3(r->d, q);
        if (v >= 0.0f && u + v <= 1.0f) {
            float t = f * dot3(e2, q);
            if (t > 0.0f && t < max_distance) return 1;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ CWBVH traversal (BVH8.h) */
static inline uint32_t oct_inv4(v3 d) {
    return (d.x < 0.0f ? 0u : 0x04040404u) | (d.y < 0.0f ? 0u : 0x02020202u) | (d.z < 0.0f ? 0u : 0x01010101u);
}
static inline uint32_t byte_of(uint32_t x, int i) { return (x >> (i * 8)) & 0xffu; }
static inline uint32_t sign_extend_s8x4(uint32_t x) { /* byte -> 0xff if its MSB is set, else 0 */
    uint32_t r = 0; for (int i = 0; i < 4; i++) if (x & (0x80u << (8 * i))) r |= 0xffu << (8 * i); return r;
}
static inline int msb(uint32_t x) { return 31 - __builtin_clz(x); }

static uint32_t node8_intersect(const uint32_t* n, const ray_t* r, uint32_t oi4, float tmax_ray) {
    v3 p = V(u2f(n[0]), u2f(n[1]), u2f(n[2]));
    uint32_t e_imask = n[3];
    float aix = u2f(byte_of(e_imask, 0) << 23) / r->d.x;
    float aiy = u2f(byte_of(e_imask, 1) << 23) / r->d.y;
    float aiz = u2f(byte_of(e_imask, 2) << 23) / r->d.z;
    float aox = (p.x - r->o.x) / r->d.x, aoy = (p.y - r->o.y) / r->d.y, aoz = (p.z - r->o.z) / r->d.z;
    uint32_t hit_mask = 0;
    for (int i = 0; i < 2; i++) {
        uint32_t meta4 = n[6 + i];
        uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        uint32_t bit_index4 = (meta4 ^ (oi4 & inner_mask4)) & 0x1f1f1f1fu;
        uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        uint32_t qlx = n[8 + i], qhx = n[10 + i], qly = n[12 + i], qhy = n[14 + i], qlz = n[16 + i], qhz = n[18 + i];
        uint32_t xmin = r->d.x < 0.0f ? qhx : qlx, xmax = r->d.x < 0.0f ? qlx : qhx;
        uint32_t ymin = r->d.y < 0.0f ? qhy : qly, ymax = r->d.y < 0.0f ? qly : qhy;
        uint32_t zmin = r->d.z < 0.0f ? qhz : qlz, zmax = r->d.z < 0.0f ? qlz : qhz;
        for (int j = 0; j < 4; j++) {
            float t0x = (float)byte_of(xmin, j) * aix + aox, t0y = (float)byte_of(ymin, j) * aiy + aoy, t0z = (float)byte_of(zmin, j) * aiz + aoz;
            float t1x = (float)byte_of(xmax, j) * aix + aox, t1y = (float)byte_of(ymax, j) * aiy + aoy, t1z = (float)byte_of(zmax, j) * aiz + aoz;
            float tmin = imax3(t0x, t0y, fmaxf(t0z, 0.0f));
            float tmax = imin3(t1x, t1y, fminf(t1z, tmax_ray));
            if (tmin < tmax) hit_mask |= byte_of(child_bits4, j) << byte_of(bit_index4, j);
        }
    }
    return hit_mask;
}

typedef struct { uint32_t x, y; } u2;

/* closest == 1: fills *hit; closest == 0: returns 1 if anything is hit before max_distance */
static int trace8(const po_scene* s, ray_t ray, hit_t* hit, float max_distance, int closest) {
    u2 stack[PO_STACK]; int sp = 0;
    u2 cur = { 0u, 0x80000000u };
    ray_t world = ray;
    uint32_t oi4 = oct_inv4(ray.d);
    int tlas_sp = PO_INVALID, mesh_id = 0, identity = 1;
    float tbest = closest ? hit->t : max_distance;
    for (;;) {
        u2 tri;
        if (cur.y & 0xff000000u) {
            uint32_t hits_imask = cur.y;
            int child_off = msb(hits_imask);
            uint32_t base = cur.x;
            cur.y &= ~(1u << child_off);
            if (cur.y & 0xff000000u) stack[sp++] = cur;
            uint32_t slot = (uint32_t)(child_off - 24) ^ (oi4 & 0xffu);
            uint32_t rel = (uint32_t)__builtin_popcount(hits_imask & ~(0xffffffffu << slot));
            const uint32_t* n = (const uint32_t*)(s->bvh_nodes + (size_t)(base + rel) * 80);
            if (closest) tbest = hit->t;
            uint32_t hm = node8_intersect(n, &ray, oi4, tbest);
            cur.x = n[4]; tri.x = n[5];
            cur.y = (hm & 0xff000000u) | byte_of(n[3], 3);
            tri.y = hm & 0x00ffffffu;
        } else { tri = cur; cur.x = 0; cur.y = 0; }

        while (tri.y != 0) {
            if (tlas_sp == PO_INVALID) {
                int off = msb(tri.y); tri.y &= ~(1u << off);
                mesh_id = (int)tri.x + off;
                if (tri.y != 0) stack[sp++] = tri;
                if (cur.y & 0xff000000u) stack[sp++] = cur;
                tlas_sp = sp;
                uint32_t root = (uint32_t)s->mesh_roots[mesh_id];
                identity = (int)(root >> 31);
                if (!identity) {
                    const float* m = s->mesh_transforms_inv + (size_t)mesh_id * 12;
                    ray.o = xf_pos(m, ray.o); ray.d = xf_dir(m, ray.d);
                    oi4 = oct_inv4(ray.d);
                }
                cur.x = root & 0x7fffffffu; cur.y = 0x80000000u;
                break;
            } else {
                int ti = msb(tri.y); tri.y &= ~(1u << ti);
                if (closest) tri_closest(s, mesh_id, (int)tri.x + ti, &ray, hit);
                else if (tri_any(s, (int)tri.x + ti, &ray, max_distance)) return 1;
            }
        }
        if ((cur.y & 0xff000000u) == 0) {
            if (sp == 0) return 0;
            if (sp == tlas_sp) {
                tlas_sp = PO_INVALID;
                if (!identity) { ray = world; oi4 = oct_inv4(ray.d); }
            }
            cur = stack[--sp];
        }
    }
}

/* ------------------------------------------------------------------ binary BVH traversal (BVH2.h) */
typedef struct { float lo[3], hi[3]; int32_t left_or_first; uint32_t count_axis; } node2_t;

static inline int box2_hit(const node2_t* n, const ray_t* r, float tmax) {
    float t0x = (n->lo[0] - r->o.x) / r->d.x, t0y = (n->lo[1] - r->o.y) / r->d.y, t0z = (n->lo[2] - r->o.z) / r->d.z;
    float t1x = (n->hi[0] - r->o.x) / r->d.x, t1y = (n->hi[1] - r->o.y) / r->d.y, t1z = (n->hi[2] - r->o.z) / r->d.z;
    float tn = iminmax(t0x, t1x, iminmax(t0y, t1y, iminmax(t0z, t1z, 0.0f)));
    float tf = imaxmin(t0x, t1x, imaxmin(t0y, t1y, imaxmin(t0z, t1z, tmax)));
    return tn < tf;
}

static int trace2(const po_scene* s, ray_t ray, hit_t* hit, float max_distance, int closest) {
    int stack[PO_STACK]; int sp = 0;
    ray_t world = ray;
    int tlas_sp = PO_INVALID, mesh_id = 0, identity = 1;
    const node2_t* nodes = (const node2_t*)s->bvh_nodes;
    stack[sp++] = 0;
    for (;;) {
        if (sp == tlas_sp) { tlas_sp = PO_INVALID; if (!identity) ray = world; }
        int ni = stack[--sp];
        const node2_t* n = nodes + ni;
        if (box2_hit(n, &ray, closest ? hit->t : max_distance)) {
            uint32_t count = n->count_axis & 0x3fffffffu, axis = n->count_axis >> 30;
            if (count > 0) {
                if (tlas_sp == PO_INVALID) {
                    tlas_sp = sp;
                    mesh_id = n->left_or_first;
                    uint32_t root = (uint32_t)s->mesh_roots[mesh_id];
                    identity = (int)(root >> 31);
                    if (!identity) {
                        const float* m = s->mesh_transforms_inv + (size_t)mesh_id * 12;
                        ray.o = xf_pos(m, ray.o); ray.d = xf_dir(m, ray.d);
                    }
                    stack[sp++] = (int)(root & 0x7fffffffu);
                } else {
                    for (int i = n->left_or_first; i < n->left_or_first + (int)count; i++) {
                        if (closest) tri_closest(s, mesh_id, i, &ray, hit);
                        else if (tri_any(s, i, &ray, max_distance)) return 1;
                    }
                }
            } else {
                float da = axis == 0 ? ray.d.x : (axis == 1 ? ray.d.y : ray.d.z);
                int first, second;
                if (da > 0.0f) { second = n->left_or_first + 1; first = n->left_or_first; }
                else           { second = n->left_or_first;     first = n->left_or_first + 1; }
                stack[sp++] = second; stack[sp++] = first;
            }
        }
        if (sp == 0) return 0;
    }
}

static void trace_closest(const po_scene* s, ray_t ray, hit_t* hit) {
    hit->t = INFINITY; hit->triangle_id = PO_INVALID; hit->mesh_id = 0; hit->u = hit->v = 0.0f;
    if (s->bvh_kind == 8) trace8(s, ray, hit, 0.0f, 1); else trace2(s, ray, hit, 0.0f, 1);
}
static int trace_any(const po_scene* s, ray_t ray, float max_distance) {
    return s->bvh_kind == 8 ? trace8(s, ray, NULL, max_distance, 0) : trace2(s, ray, NULL, max_distance, 0);
}

/* ------------------------------------------------------------------ shading helpers */
static v3 sample_sky(const po_scene* s, v3 d) {
    float phi = atan2f(-d.z, d.x), theta = acosf(fminf(fmaxf(d.y, -1.0f), 1.0f));
    float u = phi * PO_ONE_OVER_TWO_PI + 0.5f, v = theta * PO_ONE_OVER_PI;
    /* bilinear, clamp addressing, normalised coordinates */
    float fx = u * (float)s->sky_width - 0.5f, fy = v * (float)s->sky_height - 0.5f;
    int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
    float ax = fx - (float)x0, ay = fy - (float)y0;
    int x1 = x0 + 1, y1 = y0 + 1;
#define CL(v, n) ((v) < 0 ? 0 : ((v) >= (n) ? (n) - 1 : (v)))
    x0 = CL(x0, s->sky_width); x1 = CL(x1, s->sky_width); y0 = CL(y0, s->sky_height); y1 = CL(y1, s->sky_height);
    const float* a = s->sky + ((size_t)y0 * s->sky_width + x0) * 4; const float* b = s->sky + ((size_t)y0 * s->sky_width + x1) * 4;
    const float* c = s->sky + ((size_t)y1 * s->sky_width + x0) * 4; const float* e = s->sky + ((size_t)y1 * s->sky_width + x1) * 4;
    v3 r;
    r.x = (a[0] * (1 - ax) + b[0] * ax) * (1 - ay) + (c[0] * (1 - ax) + e[0] * ax) * ay;
    r.y = (a[1] * (1 - ax) + b[1] * ax) * (1 - ay) + (c[1] * (1 - ax) + e[1] * ax) * ay;
    r.z = (a[2] * (1 - ax) + b[2] * ax) * (1 - ay) + (c[2] * (1 - ax) + e[2] * ax) * ay;
    return scl(r, s->sky_scale);
}
static inline int pdf_ok(float pdf) { return isfinite(pdf) && pdf > 1e-4f; }
static inline float power_heuristic(float f, float g) { return (f * f) / (f * f + g * g); }
static inline float lum(v3 c) { return 0.299f * c.x + 0.587f * c.y + 0.114f * c.z; }
static inline float signf1(float x) { return copysignf(1.0f, x); }
static inline v3 eps_offset(v3 o, v3 d, v3 gn) { return add(o, scl(gn, signf1(dot3(d, gn)) * PO_EPSILON)); }

static int bsearch_cdf(const float* cdf, int first, int last, float value) {
    int l = first, r = last;
    for (;;) {
        int m = (l + r) / 2;
        if (m > first && value <= cdf[m - 1]) r = m - 1;
        else if (value > cdf[m]) l = m + 1;
        else return m;
    }
}
static void onb(v3 n, v3* t, v3* b) {
    float sg = copysignf(1.0f, n.z), a = -1.0f / (sg + n.z), bb = n.x * n.y * a;
    *t = V(1.0f + sg * n.x * n.x * a, sg * bb, -sg * n.x);
    *b = V(bb, sg + n.y * n.y * a, -n.y);
}

typedef struct {
    float* fb[6]; /* RADIANCE, DIRECT, INDIRECT, ALBEDO, NORMAL, POSITION framebuffers (float4 per pixel) or NULL */
} po_aovs;

static inline void fb_set(const po_aovs* a, int k, int px, v3 c) { if (a->fb[k]) { float* p = a->fb[k] + (size_t)px * 4; p[0] = c.x; p[1] = c.y; p[2] = c.z; p[3] = 0.0f; } }
static inline void fb_add(const po_aovs* a, int k, int px, v3 c) { if (a->fb[k]) { float* p = a->fb[k] + (size_t)px * 4; p[0] += c.x; p[1] += c.y; p[2] += c.z; } }
enum { AOV_RADIANCE, AOV_DIRECT, AOV_INDIRECT, AOV_ALBEDO, AOV_NORMAL, AOV_POSITION };

static inline uint32_t pack_uv(float u, float v) { return (uint32_t)(int)(u * 65535.0f) | ((uint32_t)(int)(v * 65535.0f) << 16); }

/* one path for one pixel, one pass (= one trip through generate/trace/sort/shade/shadow per bounce) */
static void trace_pixel(const po_scene* s, const po_config* cfg, const po_aovs* aov, int sample_index, int x, int y,
                        uint32_t* primary_hit /* 4 words or NULL */, long long counters[2 * PO_MAX_BOUNCES]) {
    int pixel_index = x + y * s->pitch;
    ray_t ray = camera_ray(s, cfg, pixel_index, sample_index, x, y);
    v3 throughput = V(1.0f, 1.0f, 1.0f);
    int allow_nee = 0; float last_pdf = 0.0f;
    for (int bounce = 0; bounce < cfg->num_bounces; bounce++) {
        hit_t hit; trace_closest(s, ray, &hit);
        counters[bounce]++;
        /* hits travel through the queue with 16-bit barycentrics (Buffers.h:28-47) */
        uint32_t uv = hit.triangle_id == PO_INVALID ? 0u : pack_uv(hit.u, hit.v);
        if (bounce == 0 && primary_hit) { primary_hit[0] = (uint32_t)hit.mesh_id; primary_hit[1] = (uint32_t)hit.triangle_id; primary_hit[2] = f2u(hit.t); primary_hit[3] = uv; }
        hit.u = (float)(uv & 0xffffu) / 65535.0f; hit.v = (float)(uv >> 16) / 65535.0f;

        /* ---- kernel_sort ---- */
        if (hit.triangle_id == PO_INVALID) {
            v3 ill = mul(throughput, sample_sky(s, ray.d));
            if (bounce == 0) { fb_set(aov, AOV_ALBEDO, pixel_index, V(1, 1, 1)); fb_set(aov, AOV_RADIANCE, pixel_index, ill); fb_set(aov, AOV_DIRECT, pixel_index, ill); }
            else if (bounce == 1) { fb_add(aov, AOV_RADIANCE, pixel_index, ill); fb_add(aov, AOV_DIRECT, pixel_index, ill); }
            else { fb_add(aov, AOV_RADIANCE, pixel_index, ill); fb_add(aov, AOV_INDIRECT, pixel_index, ill); }
            return;
        }
        int material_id = s->mesh_material_ids[hit.mesh_id];
        int mtype = s->material_types[material_id];
        const float* mat = s->materials + (size_t)material_id * 8;
        const float* world = s->mesh_transforms + (size_t)hit.mesh_id * 12;
        if (mtype == 0) { /* LIGHT */
            v3 p0, e1, e2; tri_pos(s, hit.triangle_id, &p0, &e1, &e2);
            v3 lgn = norm3(xf_dir(world, cross3(e1, e2)));
            v3 emission = V(mat[0], mat[1], mat[2]);
            int count_it = cfg->enable_nee ? !allow_nee : 1;
            if (count_it) {
                v3 ill = mul(throughput, emission);
                if (bounce == 0) { fb_set(aov, AOV_ALBEDO, pixel_index, V(1, 1, 1)); fb_set(aov, AOV_RADIANCE, pixel_index, emission); fb_set(aov, AOV_DIRECT, pixel_index, emission); }
                else if (bounce == 1) { fb_add(aov, AOV_RADIANCE, pixel_index, ill); fb_add(aov, AOV_DIRECT, pixel_index, ill); }
                else { fb_add(aov, AOV_RADIANCE, pixel_index, ill); fb_add(aov, AOV_INDIRECT, pixel_index, ill); }
                return;
            }
            if (cfg->enable_mis) {
                float cos_l = fabsf(dot3(ray.d, lgn));
                float light_pdf = lum(emission) * (hit.t * hit.t) / (cos_l * s->lights_total_weight);
                if (!pdf_ok(light_pdf)) return;
                float w = power_heuristic(last_pdf, light_pdf);
                v3 ill = scl(mul(throughput, emission), w);
                fb_add(aov, AOV_RADIANCE, pixel_index, ill);
                fb_add(aov, bounce == 1 ? AOV_DIRECT : AOV_INDIRECT, pixel_index, ill);
            }
            return;
        }
        /* russian roulette (Pathtracer.cu:199-218) */
        if (bounce == cfg->num_bounces - 1) return;
        if (cfg->enable_rr && bounce > 0) {
            float p = fminf(fmaxf(imax3(throughput.x, throughput.y, throughput.z), 0.0f), 1.0f);
            float rr[2]; random2(s, DIM_RR, pixel_index, bounce, sample_index, rr);
            if (rr[0] > p) return;
            throughput = V(throughput.x / p, throughput.y / p, throughput.z / p);
        }
        if (mtype != 1) return; /* only the diffuse BSDF is restated on the CPU */

        /* ---- kernel_material_diffuse ---- */
        const float* t = s->triangles + (size_t)hit.triangle_id * 24;
        v3 p0 = V(t[0], t[1], t[2]), e1 = V(t[3], t[4], t[5]), e2 = V(t[6], t[7], t[8]);
        v3 n0 = V(t[9], t[10], t[11]), ne1 = V(t[12], t[13], t[14]), ne2 = V(t[15], t[16], t[17]);
        v3 hp = add(p0, add(scl(e1, hit.u), scl(e2, hit.v)));
        v3 nrm = add(n0, add(scl(ne1, hit.u), scl(ne2, hit.v)));
        hp = xf_pos(world, hp);
        nrm = norm3(xf_dir(world, nrm));
        v3 we1 = xf_dir(world, e1), we2 = xf_dir(world, e2);
        v3 gn = cross3(we1, we2);
        gn = scl(gn, 1.0f / len3(gn));
        if (!(dot3(ray.d, gn) < 0.0f)) nrm = scl(nrm, -1.0f);
        v3 tg, bt; onb(nrm, &tg, &bt);
        v3 wi = V(dot3(tg, scl(ray.d, -1.0f)), dot3(bt, scl(ray.d, -1.0f)), dot3(nrm, scl(ray.d, -1.0f)));
        if (wi.z <= 0.0f) return;
        v3 albedo = V(mat[0], mat[1], mat[2]);
        if (bounce == 0) { fb_set(aov, AOV_ALBEDO, pixel_index, albedo); fb_set(aov, AOV_NORMAL, pixel_index, nrm); fb_set(aov, AOV_POSITION, pixel_index, hp); }
        throughput = mul(throughput, albedo);

        /* next event estimation */
        if (cfg->enable_nee && s->lights_total_weight > 0.0f) {
            float rl[2], rt[2];
            random2(s, DIM_NEE_LIGHT, pixel_index, bounce, sample_index, rl);
            random2(s, DIM_NEE_TRIANGLE, pixel_index, bounce, sample_index, rt);
            int lm = bsearch_cdf(s->light_mesh_cdf, 0, s->light_mesh_count - 1, rl[0]);
            int lmesh = s->light_mesh_transform_indices[lm];
            int ltri = s->light_triangle_indices[bsearch_cdf(s->light_triangle_cdf, s->light_mesh_triangle_span[2 * lm], s->light_mesh_triangle_span[2 * lm + 1], rl[1])];
            float u1 = rt[0], u2 = rt[1];
            if (u2 > u1) { u1 *= 0.5f; u2 -= u1; } else { u2 *= 0.5f; u1 -= u2; }
            v3 lp0, le1, le2; tri_pos(s, ltri, &lp0, &le1, &le2);
            v3 lp = add(lp0, add(scl(le1, u1), scl(le2, u2)));
            const float* lw = s->mesh_transforms + (size_t)lmesh * 12;
            lp = xf_pos(lw, lp);
            v3 lgn = norm3(xf_dir(lw, cross3(le1, le2)));
            v3 hp_o = eps_offset(hp, sub(lp, hp), gn);
            v3 lp_o = eps_offset(lp, sub(hp_o, lp), lgn);
            v3 to_l = sub(lp_o, hp_o);
            float dist = len3(to_l);
            to_l = V(to_l.x / dist, to_l.y / dist, to_l.z / dist);
            float cos_l = fabsf(dot3(to_l, lgn)), cos_h = dot3(to_l, nrm);
            const float* lmat = s->materials + (size_t)s->mesh_material_ids[lmesh] * 8;
            v3 emission = V(lmat[0], lmat[1], lmat[2]);
            if (cos_h > 0.0f) {
                float bsdf_pdf = cos_h * PO_ONE_OVER_PI;
                if (pdf_ok(bsdf_pdf)) {
                    float light_pdf = lum(emission) * (dist * dist) / (cos_l * s->lights_total_weight);
                    if (pdf_ok(light_pdf)) {
                        float w = cfg->enable_mis ? power_heuristic(light_pdf, bsdf_pdf) : 1.0f;
                        v3 ill = mul(mul(throughput, V(bsdf_pdf, bsdf_pdf, bsdf_pdf)), emission);
                        ill = scl(ill, w); ill = V(ill.x / light_pdf, ill.y / light_pdf, ill.z / light_pdf);
                        counters[PO_MAX_BOUNCES + bounce]++;
                        ray_t sr; sr.o = hp_o; sr.d = to_l;
                        if (!trace_any(s, sr, dist)) {
                            fb_add(aov, AOV_RADIANCE, pixel_index, ill);
                            if (bounce == 0) fb_set(aov, AOV_DIRECT, pixel_index, ill); else fb_add(aov, AOV_INDIRECT, pixel_index, ill);
                        }
                    }
                }
            }
        }
        /* sample the BSDF: cosine weighted */
        float rb[2]; random2(s, DIM_BSDF_0, pixel_index, bounce, sample_index, rb);
        float dk[2]; sample_disk(rb[0], rb[1], dk);
        v3 wo = V(dk[0], dk[1], safe_sqrt(1.0f - (dk[0] * dk[0] + dk[1] * dk[1])));
        v3 dir = V(tg.x * wo.x + bt.x * wo.y + nrm.x * wo.z, tg.y * wo.x + bt.y * wo.y + nrm.y * wo.z, tg.z * wo.x + bt.z * wo.y + nrm.z * wo.z);
        float pdf = wo.z * PO_ONE_OVER_PI;
        if (!pdf_ok(pdf)) return;
        ray.o = eps_offset(hp, dir, gn); ray.d = dir;
        allow_nee = 1; last_pdf = pdf;
    }
}

/* ------------------------------------------------------------------ exported entry points */
/* Renders ONE pass (sample_index) over pixels [y0,y1) x [0,width) into the framebuffers (which the caller zeroed),
   like one Pathtracer::render() before kernel_accumulate.  counters: [0..127] closest-hit rays per bounce,
   [128..255] shadow rays per bounce.  primary_hits: 4 uint32 per pixel (mesh, triangle, t bits, uv16|uv16) or NULL. */
int po_render_pass(const po_scene* s, const po_config* cfg, int sample_index, int y0, int y1,
                   float* fb_radiance, float* fb_direct, float* fb_indirect, float* fb_albedo, float* fb_normal, float* fb_position,
                   uint32_t* primary_hits, long long* counters) {
    po_aovs aov = { { fb_radiance, fb_direct, fb_indirect, fb_albedo, fb_normal, fb_position } };
    long long total[2 * PO_MAX_BOUNCES]; memset(total, 0, sizeof(total));
#pragma omp parallel
    {
        long long local[2 * PO_MAX_BOUNCES]; memset(local, 0, sizeof(local));
#pragma omp for schedule(dynamic, 4)
        for (int y = y0; y < y1; y++)
            for (int x = 0; x < s->width; x++)
                trace_pixel(s, cfg, &aov, sample_index, x, y, primary_hits ? primary_hits + (size_t)(x + y * s->pitch) * 4 : NULL, local);
#pragma omp critical
        for (int i = 0; i < 2 * PO_MAX_BOUNCES; i++) total[i] += local[i];
    }
    if (counters) for (int i = 0; i < 2 * PO_MAX_BOUNCES; i++) counters[i] += total[i];
    return 0;
}

/* kernel_accumulate (AOV.h:35-46): acc += (fb - acc) / n for n > 0, acc = fb for n == 0 */
void po_accumulate(float* acc, const float* fb, long long count, float n) {
    for (long long i = 0; i < count; i++) {
        if (n > 0.0f) acc[i] += (fb[i] - acc[i]) / n; else acc[i] = fb[i];
    }
}

/* Brute-force closest hit over every instance x triangle (no BVH): validates the BVH builders. */
void po_brute_force_primary(const po_scene* s, const po_config* cfg, int sample_index, int mesh_count, const int32_t* mesh_tri_first,
                            const int32_t* mesh_tri_count, uint32_t* hits_out) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < s->height; y++) for (int x = 0; x < s->width; x++) {
        int px = x + y * s->pitch;
        ray_t ray = camera_ray(s, cfg, px, sample_index, x, y);
        hit_t h; h.t = INFINITY; h.triangle_id = PO_INVALID; h.mesh_id = 0; h.u = h.v = 0;
        for (int m = 0; m < mesh_count; m++) {
            ray_t r = ray;
            if (!((uint32_t)s->mesh_roots[m] >> 31)) { const float* mi = s->mesh_transforms_inv + (size_t)m * 12; r.o = xf_pos(mi, ray.o); r.d = xf_dir(mi, ray.d); }
            for (int t = mesh_tri_first[m]; t < mesh_tri_first[m] + mesh_tri_count[m]; t++) tri_closest(s, m, t, &r, &h);
        }
        uint32_t* o = hits_out + (size_t)px * 4;
        o[0] = (uint32_t)h.mesh_id; o[1] = (uint32_t)h.triangle_id; o[2] = f2u(h.t); o[3] = h.triangle_id == PO_INVALID ? 0u : pack_uv(h.u, h.v);
    }
}

void po_primary_hits(const po_scene* s, const po_config* cfg, int sample_index, uint32_t* hits_out) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < s->height; y++) for (int x = 0; x < s->width; x++) {
        int px = x + y * s->pitch;
        ray_t ray = camera_ray(s, cfg, px, sample_index, x, y);
        hit_t h; trace_closest(s, ray, &h);
        uint32_t* o = hits_out + (size_t)px * 4;
        o[0] = (uint32_t)h.mesh_id; o[1] = (uint32_t)h.triangle_id; o[2] = f2u(h.t); o[3] = h.triangle_id == PO_INVALID ? 0u : pack_uv(h.u, h.v);
    }
}

int po_sizeof_scene(void) { return (int)sizeof(po_scene); }
