// TLAS refit on the device.
//
// The reference rebuilds the top-level BVH on the CPU every frame that an instance moves (Integrator::build_tlas,
// Src/Renderer/Integrators/Integrator.cpp:399-430: SAH build over the instance boxes, conversion, blocking upload).  Here the topology
// of the last host-built TLAS is kept and only its boxes are recomputed, on the GPU, from the new instance transforms:
//
//   phase A  world box of every instance = its BLAS root box moved by the new transform (the centre / extent form of
//            Src/Math/AABB.cpp's transform, like Mesh::update), padded to a minimum size like Mesh.cpp:16-33;
//   phase B  bottom-up over the TLAS nodes: a node whose internal children are done takes the union of its child boxes and is
//            re-quantised in place (origin p, per-axis exponent e, 8-bit child boxes -- the arithmetic of the CPU converter,
//            host/bvh_build.cpp WideConverter::emit / BVH8Converter.cpp:193-246); imask, child / instance bases, meta and the slot
//            assignment are left alone.  BVH2 TLAS nodes (config 0) get their float box replaced.
//
// One CTA: the passes of phase B are separated by __syncthreads (one pass per tree level, a TLAS of 2 M nodes needs ~7).  Slots the
// static merge blanked (meta = 0) stay blank and do not contribute, so the pruned TLAS gets tighter, not looser.
// The result is a valid BVH over the same instances: closest hits are those of a rebuilt TLAS (order of equal-distance hits aside);
// only traversal cost can differ after large motions, and the host may upload a rebuilt TLAS (ptb_update_instances) whenever it likes.
#pragma once
#include "ptb_types.cuh"

struct RefitArgs {
    float4*       nodes;          // TLAS at the front of the node array (80-byte CWBVH nodes, or 32-byte BVH2 nodes)
    int           node_count;
    int           kind;           // 8 or 2
    int           instance_count;
    const float*  local_boxes;    // instance_count x 6: BLAS root box in mesh space
    const float4* transforms;     // instance_count x 3 rows
    float*        instance_boxes; // scratch, instance_count x 6
    float*        node_boxes;     // scratch, node_count x 6
    int*          done;           // scratch, node_count: pass number + 1 in which the node was finished, 0 = not yet
};

struct RBox { float3 lo, hi; };
PTB_DI RBox rbox_empty() { RBox b; b.lo = f3(PTB_INF); b.hi = f3(-PTB_INF); return b; }
PTB_DI bool rbox_valid(const RBox& b) { return b.lo.x <= b.hi.x; }
PTB_DI void rbox_grow(RBox& b, const RBox& o) {
    b.lo = f3(fminf(b.lo.x, o.lo.x), fminf(b.lo.y, o.lo.y), fminf(b.lo.z, o.lo.z));
    b.hi = f3(fmaxf(b.hi.x, o.hi.x), fmaxf(b.hi.y, o.hi.y), fmaxf(b.hi.z, o.hi.z));
}
PTB_DI RBox rbox_load(const float* p) { RBox b; b.lo = f3(p[0], p[1], p[2]); b.hi = f3(p[3], p[4], p[5]); return b; }
PTB_DI void rbox_store(float* p, const RBox& b) { p[0] = b.lo.x; p[1] = b.lo.y; p[2] = b.lo.z; p[3] = b.hi.x; p[4] = b.hi.y; p[5] = b.hi.z; }

// smallest power of two 2^k with 255 * 2^k >= extent, as the biased exponent byte the node stores (float bits >> 23).
// Integer construction instead of exp2(ceil(log2(x))): exact under --use_fast_math, and checked against the one-off rounding of
// extent / 255 so that no quantised child coordinate can exceed 255.
PTB_DI unsigned refit_exponent(float extent) {
    float x = __fmul_rn(extent, 1.0f / 255.0f);
    unsigned bits = __float_as_uint(x);
    unsigned e = (bits >> 23) & 0xffu;
    if (bits & 0x007fffffu) e++;
    if (e < 1u) e = 1u;                         // zero / denormal extent: the smallest normal step
    if (e > 253u) e = 253u;
    while (e < 253u && __fmul_rn(extent, __uint_as_float((254u - e) << 23)) > 255.0f) e++;      // extent / 2^k <= 255
    return e;
}

__global__ void __launch_bounds__(1024) k_refit_tlas(const RefitArgs A) {
    const int tid = threadIdx.x, nth = blockDim.x;
    // ---- phase A: instance boxes in world space
    for (int i = tid; i < A.instance_count; i += nth) {
        RBox l = rbox_load(A.local_boxes + 6 * size_t(i));
        float4 r0 = A.transforms[3 * size_t(i)], r1 = A.transforms[3 * size_t(i) + 1], r2 = A.transforms[3 * size_t(i) + 2];
        float3 c = 0.5f * (l.lo + l.hi), h = 0.5f * (l.hi - l.lo);
        float3 nc = f3(r0.x * c.x + r0.y * c.y + r0.z * c.z + r0.w, r1.x * c.x + r1.y * c.y + r1.z * c.z + r1.w, r2.x * c.x + r2.y * c.y + r2.z * c.z + r2.w);
        float3 nh = f3(fabsf(r0.x) * h.x + fabsf(r0.y) * h.y + fabsf(r0.z) * h.z, fabsf(r1.x) * h.x + fabsf(r1.y) * h.y + fabsf(r1.z) * h.z,
                       fabsf(r2.x) * h.x + fabsf(r2.y) * h.y + fabsf(r2.z) * h.z);
        // one ulp-scale pad: the centre / extent form rounds, and the box must contain the transformed BLAS root box
        float3 pad = f3(1e-6f * (fabsf(nc.x) + nh.x), 1e-6f * (fabsf(nc.y) + nh.y), 1e-6f * (fabsf(nc.z) + nh.z));
        RBox w; w.lo = nc - nh - pad; w.hi = nc + nh + pad;
        float lo[3] = { w.lo.x, w.lo.y, w.lo.z }, hi[3] = { w.hi.x, w.hi.y, w.hi.z };
        for (int k = 0; k < 3; k++) {           // Mesh.cpp:24-33: no axis thinner than 0.001
            float eps = 0.001f;
            while (hi[k] - lo[k] < eps) { lo[k] -= eps; hi[k] += eps; eps *= 2.0f; }
        }
        w.lo = f3(lo[0], lo[1], lo[2]); w.hi = f3(hi[0], hi[1], hi[2]);
        rbox_store(A.instance_boxes + 6 * size_t(i), w);
    }
    for (int i = tid; i < A.node_count; i += nth) A.done[i] = 0;
    __syncthreads();

    // ---- phase B: bottom-up, one tree level per pass
    for (int pass = 1; pass <= 64; pass++) {
        for (int i = tid; i < A.node_count; i += nth) {
            if (A.done[i]) continue;
            if (A.kind == 2 && i == 1) { A.done[i] = pass; continue; }               // BVH2: node 1 is the unused alignment slot
            RBox box = rbox_empty();
            bool ready = true;
            if (A.kind == 8) {
                float4* n = A.nodes + 5 * size_t(i);
                float4 n0 = n[0], n1 = n[1];
                unsigned imask = byte_of(__float_as_uint(n0.w), 3);
                unsigned base_child = __float_as_uint(n1.x), base_inst = __float_as_uint(n1.y);
                unsigned meta_lo = __float_as_uint(n1.z), meta_hi = __float_as_uint(n1.w);
                RBox child[8];
                unsigned internal = 0;
                for (int k = 0; k < 8; k++) {
                    unsigned meta = byte_of(k < 4 ? meta_lo : meta_hi, k & 3);
                    child[k] = rbox_empty();
                    if (imask & (1u << k)) {
                        unsigned c = base_child + internal++;
                        if (!meta) continue;                                         // blanked by the static merge
                        if (c >= unsigned(A.node_count)) continue;
                        int d = A.done[c];
                        if (d == 0 || d >= pass) { ready = false; break; }           // finished in an EARLIER pass only (ordered by the barrier)
                        child[k] = rbox_load(A.node_boxes + 6 * size_t(c));
                    } else if (meta) {
                        unsigned first = meta & 31u, count = __popc(meta >> 5);
                        for (unsigned t = 0; t < count; t++) {
                            unsigned inst = base_inst + first + t;
                            if (inst < unsigned(A.instance_count)) rbox_grow(child[k], rbox_load(A.instance_boxes + 6 * size_t(inst)));
                        }
                    }
                    if (rbox_valid(child[k])) rbox_grow(box, child[k]);
                }
                if (!ready) continue;
                if (rbox_valid(box)) {
                    unsigned ex = refit_exponent(__fsub_ru(box.hi.x, box.lo.x)), ey = refit_exponent(__fsub_ru(box.hi.y, box.lo.y)), ez = refit_exponent(__fsub_ru(box.hi.z, box.lo.z));
                    float rx = __uint_as_float((254u - ex) << 23), ry = __uint_as_float((254u - ey) << 23), rz = __uint_as_float((254u - ez) << 23);   // 2^-k
                    unsigned q[6][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } };    // lo x,y,z  hi x,y,z ; two words of 4 children
                    for (int k = 0; k < 8; k++) {
                        if (!rbox_valid(child[k])) continue;
                        // differences rounded outwards (down for the low side, up for the high side); the scale is a power of two
                        float v[6] = { floorf(__fmul_rn(__fsub_rd(child[k].lo.x, box.lo.x), rx)), floorf(__fmul_rn(__fsub_rd(child[k].lo.y, box.lo.y), ry)),
                                       floorf(__fmul_rn(__fsub_rd(child[k].lo.z, box.lo.z), rz)),
                                       ceilf(__fmul_rn(__fsub_ru(child[k].hi.x, box.lo.x), rx)), ceilf(__fmul_rn(__fsub_ru(child[k].hi.y, box.lo.y), ry)),
                                       ceilf(__fmul_rn(__fsub_ru(child[k].hi.z, box.lo.z), rz)) };
                        for (int c = 0; c < 6; c++) q[c][k >> 2] |= unsigned(fminf(fmaxf(v[c], 0.0f), 255.0f)) << (8 * (k & 3));
                    }
                    n[0] = make_float4(box.lo.x, box.lo.y, box.lo.z, __uint_as_float(ex | (ey << 8) | (ez << 16) | (imask << 24)));
                    n[2] = make_float4(__uint_as_float(q[0][0]), __uint_as_float(q[0][1]), __uint_as_float(q[3][0]), __uint_as_float(q[3][1]));
                    n[3] = make_float4(__uint_as_float(q[1][0]), __uint_as_float(q[1][1]), __uint_as_float(q[4][0]), __uint_as_float(q[4][1]));
                    n[4] = make_float4(__uint_as_float(q[2][0]), __uint_as_float(q[2][1]), __uint_as_float(q[5][0]), __uint_as_float(q[5][1]));
                }
            } else {
                float4* n = A.nodes + 2 * size_t(i);
                float4 a = n[0], b = n[1];
                int left_or_first = __float_as_int(b.z);
                unsigned count = __float_as_uint(b.w) & 0x3fffffffu;
                if (count) {
                    for (unsigned t = 0; t < count; t++) {
                        unsigned inst = unsigned(left_or_first) + t;
                        if (inst < unsigned(A.instance_count)) rbox_grow(box, rbox_load(A.instance_boxes + 6 * size_t(inst)));
                    }
                } else {
                    for (int c = left_or_first; c <= left_or_first + 1; c++) {
                        if (c < 0 || c >= A.node_count) continue;
                        int d = A.done[c];
                        if (d == 0 || d >= pass) { ready = false; break; }
                        RBox cb = rbox_load(A.node_boxes + 6 * size_t(c));
                        if (rbox_valid(cb)) rbox_grow(box, cb);
                    }
                }
                if (!ready) continue;
                if (rbox_valid(box)) { n[0] = make_float4(box.lo.x, box.lo.y, box.lo.z, box.hi.x); n[1] = make_float4(box.hi.y, box.hi.z, b.z, b.w); }
                (void)a;
            }
            rbox_store(A.node_boxes + 6 * size_t(i), box);
            A.done[i] = pass;
        }
        __syncthreads();
        if (A.done[0]) break;               // uniform: every thread reads it after the barrier
        __syncthreads();                    // nobody may start the next pass (and write done[0]) before everyone has read it
    }
}
