// Wavefront kernels (sm_100a).  One pass = generate -> per bounce { trace, sort, shade x types, shadow trace } -> accumulate.
// All 1-D kernels are persistent grid-stride loops sized to the machine (148 SMs x resident CTAs) and read their
// work count from device memory, so the host never reads a queue size back (the reference launches full-width
// grids and early-outs; Pathtracer.cpp:284-289).
#pragma once
#include "ptb_device.cuh"

#ifndef PTB_TRACE_BLOCK
#define PTB_TRACE_BLOCK 256
#endif
#ifndef PTB_TRACE_MIN_BLOCKS
#define PTB_TRACE_MIN_BLOCKS 4          // 64 registers/thread, 1024 threads/SM: measured 4 % faster than 3 x 80 registers on Sponza
#endif
#ifndef PTB_SM_STACK
#define PTB_SM_STACK 10                   // stack entries per thread kept in shared memory
#endif
#define PTB_STACK_TOTAL 32                // BVH_STACK_SIZE, Common.h:104
#define PTB_LOCAL_STACK (PTB_STACK_TOTAL - PTB_SM_STACK)
#ifndef PTB_TLAS_STAGE_MAX_NODES
#define PTB_TLAS_STAGE_MAX_NODES 128      // up to 10 KB of nodes bulk-copied (TMA) into shared memory per CTA; 256 / 384 were slower (they shrink the L1:
#endif                                    // profiles/r2_variants_smem_vs_l1.log), 8 .. 128 within noise of each other
#ifndef PTB_DYNFETCH_ND
#define PTB_DYNFETCH_ND 2                 // dynamic fetch heuristic, Ylitie et al. 2017 section 4.4 (BVH8.h:109-111 uses 4 / 16; 2 / 8 measured 1.6 % faster here)
#endif
#ifndef PTB_DYNFETCH_NW
#define PTB_DYNFETCH_NW 8
#endif
#ifndef PTB_POSTPONE_DIVISOR
#define PTB_POSTPONE_DIVISOR 5            // triangle postponing threshold (BVH8.h:12-15)
#endif
#ifndef PTB_STAGE_MERGED_TOP
#define PTB_STAGE_MERGED_TOP 1            // stage the merged tree's top levels whenever a merged tree exists (0: only when the TLAS is never walked)
#endif
#ifndef PTB_SHADE_MIN_BLOCKS_DIFFUSE
#define PTB_SHADE_MIN_BLOCKS_DIFFUSE 4    // 64 registers (120 B of spills) but twice the gathers in flight: frame 32.72 -> 32.21 ms
#endif
#ifndef PTB_TILED_GENERATE
#define PTB_TILED_GENERATE 1              // primary rays enumerated so that a warp covers an 8x4 pixel tile (not a 32x1 strip)
#endif

// ------------------------------------------------------------------------------------------ tile mapping
// local index -> global pixel: rows are dealt to ranks in interleaved bands of band_rows rows.
PTB_DI void local_to_pixel(const Frame& P, int local, int& x, int& y) {
    int row = local / P.width;
    x = local - row * P.width;
    int band = row / P.band_rows;
    y = (band * P.world + P.rank) * P.band_rows + (row - band * P.band_rows);
}

// Enumeration order of the primary rays: same pixel set as local_to_pixel, but consecutive groups of 32 indices cover an
// 8x4 pixel tile, so a warp's rays (and every later bounce spawned from them: queue order is inherited) start out spatially
// compact.  Purely a scheduling choice: pixel_index, the RNG key, is unchanged.
PTB_DI void generate_order_to_pixel(const Frame& P, int local, int& x, int& y) {
    const int band_px = P.band_rows * P.width;
    int band = local / band_px;
    int j = local - band * band_px;
    int y0 = (band * P.world + P.rank) * P.band_rows;
    int rows = min(P.band_rows, P.height - y0);
    int row, xx;
    if (((P.width & 7) | (rows & 3)) == 0) {
        int tile = j >> 5, t = j & 31;
        int tiles_x = P.width >> 3;
        int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        xx = tx * 8 + (t & 7); row = ty * 4 + (t >> 3);
    } else {
        row = j / P.width; xx = j - row * P.width;
    }
    x = xx; y = y0 + row;
}

// ------------------------------------------------------------------------------------------ generate
__global__ void __launch_bounds__(256) k_generate(const __grid_constant__ Frame P) {
    const RayQueue& q = P.q[0];
    const int total = P.local_pixels * P.wave_samples;          // slot-major: all pixels of pass slot 0, then slot 1, ...
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int slot = i / P.local_pixels;
        int x, y;
#if PTB_TILED_GENERATE
        generate_order_to_pixel(P, i - slot * P.local_pixels, x, y);
#else
        local_to_pixel(P, i - slot * P.local_pixels, x, y);
#endif
        int pixel_index = x + y * P.pitch;
        Ray r = camera_ray(P, pixel_index, P.first_sample + slot, x, y);
        q.od0[i] = make_float4(r.o.x, r.o.y, r.o.z, r.d.x);
        q.od1[i] = make_float4(r.d.y, r.d.z, 0.0f, 0.0f);
        q.pix[i] = unsigned(pixel_index) | (unsigned(slot) << P.pix_bits);
    }
}

// ------------------------------------------------------------------------------------------ ray ordering
// Secondary and shadow rays leave k_shade in pixel order: a warp's 32 rays start close together but point anywhere, and the
// warp-synchronous traversal then runs at half occupancy of its lanes (ncu: 15.9 of 32 threads active on shadow rays vs 25.5 on
// primary rays).  Before such a queue is traced it is counting-sorted by DIRECTION BIN (8 octants or 8x8 octahedral cells) --
// inside a bin the pixel order survives, so a warp gets rays that start close together AND point the same way.  The sort only
// produces a 4-byte index per ray (`order`), the 48-byte rays stay where they are; the reference has no such step (its
// kernel_trace_* consume the queues in emission order, Pathtracer.cu:165-197).  Results do not depend on the order rays are traced in.
PTB_DI unsigned direction_bin(const float3& d, int bins) {
    if (bins <= 8) return (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
    float inv = 1.0f / (fabsf(d.x) + fabsf(d.y) + fabsf(d.z));
    float u = d.x * inv, v = d.y * inv;
    if (d.z < 0.0f) {
        float uu = (1.0f - fabsf(v)) * (u >= 0.0f ? 1.0f : -1.0f);
        float vv = (1.0f - fabsf(u)) * (v >= 0.0f ? 1.0f : -1.0f);
        u = uu; v = vv;
    }
    int iu = min(7, max(0, int((u * 0.5f + 0.5f) * 8.0f)));
    int iv = min(7, max(0, int((v * 0.5f + 0.5f) * 8.0f)));
    return unsigned(iv * 8 + iu);
}
PTB_DI int* bin_row(const Frame& P, int bounce, bool shadow) { return P.bin_counts + ((bounce * 2 + (shadow ? 1 : 0)) * PTB_ORDER_MAX_BINS); }

// pass 1: bin of every ray + its rank inside the bin (CTA-aggregated: one global atomic per bin per 256 rays)
template <bool SHADOW>
__global__ void __launch_bounds__(256) k_bin_count(const __grid_constant__ Frame P, int bounce) {
    __shared__ int s_cnt[PTB_ORDER_MAX_BINS];
    __shared__ int s_base[PTB_ORDER_MAX_BINS];
    const int count = SHADOW ? P.counters->shadow[bounce] : P.counters->trace[bounce];
    const float4* od0 = SHADOW ? P.sq.od0 : P.q[bounce & 1].od0;
    const float4* od1 = SHADOW ? P.sq.od1 : P.q[bounce & 1].od1;
    int* row = bin_row(P, bounce, SHADOW);
    for (int start = blockIdx.x * blockDim.x; start < count; start += gridDim.x * blockDim.x) {    // CTA-uniform trip count
        if (threadIdx.x < PTB_ORDER_MAX_BINS) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        int i = start + threadIdx.x;
        unsigned key = 0; int local = 0;
        if (i < count) {
            float4 a = od0[i], b = od1[i];
            key = direction_bin(f3(a.w, b.x, b.y), P.order_bins);
            local = atomicAdd(&s_cnt[key], 1);
        }
        __syncthreads();
        if (threadIdx.x < PTB_ORDER_MAX_BINS && s_cnt[threadIdx.x] > 0) s_base[threadIdx.x] = atomicAdd(&row[threadIdx.x], s_cnt[threadIdx.x]);
        __syncthreads();
        if (i < count) { P.bin_key[i] = (unsigned char)key; P.bin_rank[i] = unsigned(s_base[key] + local); }
    }
}
// pass 2: order[bin start + rank] = ray index
template <bool SHADOW>
__global__ void __launch_bounds__(256) k_bin_scatter(const __grid_constant__ Frame P, int bounce) {
    __shared__ int s_start[PTB_ORDER_MAX_BINS];
    const int count = SHADOW ? P.counters->shadow[bounce] : P.counters->trace[bounce];
    const int* row = bin_row(P, bounce, SHADOW);
    if (threadIdx.x == 0) { int acc = 0; for (int b = 0; b < PTB_ORDER_MAX_BINS; b++) { s_start[b] = acc; acc += row[b]; } }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        P.order[unsigned(s_start[P.bin_key[i]]) + P.bin_rank[i]] = unsigned(i);
}

// ------------------------------------------------------------------------------------------ TMA staging of the TLAS
PTB_DI unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// Thread 0 issues one 1-D bulk async copy (cp.async.bulk, the non-tensor TMA path) of the first `bytes` of the
// node array into shared memory, completion tracked by an mbarrier; everybody waits on phase 0.
PTB_DI void stage_tlas_nodes(float4* dst, const float4* src, unsigned bytes, unsigned long long* bar) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
    }
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ CWBVH traversal
// Persistent threads; a warp stays converged and re-fills idle lanes with ONE atomic per refill (warp-aggregated
// dynamic fetch).  Algorithm = compressed-wide-BVH traversal of Src/CUDA/Raytracing/BVH8.h:113-444 (node groups /
// triangle groups, octant-ordered child pops, TLAS->BLAS instancing, triangle postponing), restructured so that all
// ballots are taken in uniform control flow instead of sampling __activemask().
struct TraceShared {
    const float4* tlas;     // staged nodes in shared memory
    unsigned      stage_base; // node index of the first staged node
    int           staged;   // number of staged nodes
    uint2*        stack;    // PTB_SM_STACK x blockDim entries, strided by blockDim (conflict free)
};

PTB_DI void stack_push(const TraceShared& S, uint2* local, int& sp, uint2 v) {
    if (sp < PTB_SM_STACK) S.stack[sp * PTB_TRACE_BLOCK + threadIdx.x] = v; else local[sp - PTB_SM_STACK] = v;
    sp++;
}
PTB_DI uint2 stack_pop(const TraceShared& S, const uint2* local, int& sp) {
    sp--;
    return sp < PTB_SM_STACK ? S.stack[sp * PTB_TRACE_BLOCK + threadIdx.x] : local[sp - PTB_SM_STACK];
}

// STATS = true additionally counts node visits / triangle tests / instance transforms per ray kind (roofline accounting:
// algorithmic bytes = 80 B per node + 48 B per triangle + 48 B per instance transform + the ray/hit streams).
template <bool SHADOW, bool STATS>
__global__ void __launch_bounds__(PTB_TRACE_BLOCK, PTB_TRACE_MIN_BLOCKS) k_trace8(const __grid_constant__ Frame P, int bounce, const unsigned* __restrict__ order) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw);
    float4* tlas_sm = reinterpret_cast<float4*>(smem_raw + 16);
    TraceShared S;
    // staged in shared memory: the top levels of the merged BVH when there is one (breadth-first layout, so its first 128 nodes are its
    // first ~3 levels: every ray visits several of them), otherwise the TLAS
    const bool flat_only = P.flat_root >= 0 && (P.flat_all || PTB_STAGE_MERGED_TOP);
    S.stage_base = flat_only ? unsigned(P.flat_root) : 0u;
    S.staged = flat_only ? min(P.flat_node_count, PTB_TLAS_STAGE_MAX_NODES) : min(P.tlas_nodes, PTB_TLAS_STAGE_MAX_NODES);
    S.tlas = tlas_sm;
    S.stack = reinterpret_cast<uint2*>(smem_raw + 16 + size_t(PTB_TLAS_STAGE_MAX_NODES) * 80);
    stage_tlas_nodes(tlas_sm, P.nodes8 + 5 * size_t(S.stage_base), unsigned(S.staged) * 80u, bar);

    const int count = SHADOW ? P.counters->shadow[bounce] : P.counters->trace[bounce];
    int* retired = SHADOW ? &P.counters->retired_shadow[bounce] : &P.counters->retired[bounce];
    const RayQueue& q = P.q[bounce & 1];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned FULL = 0xffffffffu;

    uint2 local_stack[PTB_LOCAL_STACK];
    int sp = 0;
    uint2 cur = make_uint2(0, 0);
    int ray_index = 0;
    Ray ray; ray.o = f3(0.0f); ray.d = f3(1.0f);
    unsigned oct4 = 0;
    Hit hit; hit.t = 0.0f; hit.u = hit.v = 0.0f; hit.mesh_id = 0; hit.triangle_id = PTB_INVALID;   // shadow rays: hit.t = max distance
    int tlas_sp = PTB_INVALID, mesh_id = 0;
    bool identity = true, live = false, exhausted = false;
    unsigned long long st_nodes = 0, st_tris = 0, st_xf = 0, st_rays = 0, st_miss = 0;

    while (true) {
        // ---- refill idle lanes (one atomic per warp)
        unsigned idle = __ballot_sync(FULL, !live && !exhausted);
        if (idle) {
            int leader = __ffs(idle) - 1;
            int base = 0;
            if (lane == unsigned(leader)) base = atomicAdd(retired, __popc(idle));
            base = __shfl_sync(FULL, base, leader);
            if (!live && !exhausted) {
                ray_index = base + __popc(idle & ((1u << lane) - 1u));
                if (ray_index < count) {
                    if (order) ray_index = int(order[ray_index]);          // direction-binned order (k_bin_scatter)
                    float4 a, b;
                    if (SHADOW) { a = P.sq.od0[ray_index]; b = P.sq.od1[ray_index]; hit.t = b.z; }
                    else        { a = q.od0[ray_index];    b = q.od1[ray_index];    hit.t = PTB_INF; hit.triangle_id = PTB_INVALID; }
                    ray.o = f3(a.x, a.y, a.z); ray.d = f3(a.w, b.x, b.y);
                    oct4 = ray_octant_inv4(ray.d);
                    sp = 0; live = true;
                    if (P.flat_root >= 0) {
                        // merged static BVH first (it sets a tight hit.t early); the TLAS root waits on the stack for what is not merged
                        if (!P.flat_all) stack_push(S, local_stack, sp, make_uint2(0u, 0x80000000u));
                        tlas_sp = sp; mesh_id = PTB_FLAT_MESH; identity = true;
                        cur = make_uint2(unsigned(P.flat_root), 0x80000000u);
                    } else {
                        cur = make_uint2(0u, 0x80000000u);
                        tlas_sp = PTB_INVALID;
                    }
                    if (STATS) st_rays++;
                } else exhausted = true;
            }
        }
        if (__ballot_sync(FULL, live) == 0) {          // nothing in flight and the queue is empty
            if (STATS) {
                for (int o = 16; o > 0; o >>= 1) {
                    st_nodes += __shfl_down_sync(FULL, st_nodes, o); st_tris += __shfl_down_sync(FULL, st_tris, o);
                    st_xf += __shfl_down_sync(FULL, st_xf, o); st_rays += __shfl_down_sync(FULL, st_rays, o); st_miss += __shfl_down_sync(FULL, st_miss, o);
                }
                if (lane == 0) {
                    TraceStats* ts = P.trace_stats + (SHADOW ? 1 : 0);
                    atomicAdd(&ts->nodes, st_nodes); atomicAdd(&ts->triangles, st_tris); atomicAdd(&ts->instance_transforms, st_xf);
                    atomicAdd(&ts->rays, st_rays); atomicAdd(&ts->misses, st_miss);
                }
            }
            return;
        }

        int lost = 0;
        while (true) {
            uint2 tri = make_uint2(0, 0);
            if (live) {
                if (cur.y & 0xff000000u) {
                    unsigned hits_imask = cur.y;
                    unsigned child_off = msb(hits_imask);
                    unsigned child_base = cur.x;
                    cur.y &= ~(1u << child_off);
                    if (cur.y & 0xff000000u) stack_push(S, local_stack, sp, cur);
                    unsigned slot = (child_off - 24u) ^ (oct4 & 0xffu);
                    unsigned rel = __popc(hits_imask & ~(0xffffffffu << slot));
                    unsigned ni = child_base + rel;
                    float4 n0, n1, n2, n3, n4;
                    if (ni - S.stage_base < unsigned(S.staged)) {
                        const float4* n = S.tlas + 5 * (ni - S.stage_base);
                        n0 = n[0]; n1 = n[1]; n2 = n[2]; n3 = n[3]; n4 = n[4];
                    } else {
                        const float4* n = P.nodes8 + 5 * size_t(ni);
                        n0 = __ldg(n); n1 = __ldg(n + 1); n2 = __ldg(n + 2); n3 = __ldg(n + 3); n4 = __ldg(n + 4);
                    }
                    if (STATS) st_nodes++;
                    unsigned hm = cwbvh_node_intersect(ray, oct4, hit.t, n0, n1, n2, n3, n4);
                    unsigned imask = byte_of(__float_as_uint(n0.w), 3);
                    cur.x = __float_as_uint(n1.x);
                    tri.x = __float_as_uint(n1.y);
                    cur.y = (hm & 0xff000000u) | imask;
                    tri.y = hm & 0x00ffffffu;
                } else {
                    tri = cur; cur = make_uint2(0, 0);
                }
            }
            const int postpone_threshold = __popc(__ballot_sync(FULL, live)) / PTB_POSTPONE_DIVISOR;

            // TLAS level: a "triangle" is an instance; enter its BLAS (BVH8.h:204-232)
            if (live && tri.y != 0 && tlas_sp == PTB_INVALID) {
                unsigned off = msb(tri.y);
                tri.y &= ~(1u << off);
                int inst = int(tri.x + off);
                unsigned root = unsigned(__ldg(P.mesh_roots + inst));
                if (tri.y != 0) stack_push(S, local_stack, sp, tri);
                tri.y = 0;
                if (!(root & PTB_ROOT_MERGED)) {            // merged instances were already intersected through the merged BVH
                    mesh_id = inst;
                    if (cur.y & 0xff000000u) stack_push(S, local_stack, sp, cur);
                    tlas_sp = sp;
                    identity = (root & PTB_ROOT_IDENTITY) != 0;
                    if (!identity) {
                        Mat3x4 inv = load_mat(P.mesh_transforms_inv, mesh_id);
                        ray.o = xform_pos(inv, ray.o);
                        ray.d = xform_dir(inv, ray.d);
                        oct4 = ray_octant_inv4(ray.d);
                        if (STATS) st_xf++;
                    }
                    cur = make_uint2(root & 0x3fffffffu, 0x80000000u);
                }
            }
            // BLAS level: test triangles, or postpone when too few lanes want to (BVH8.h:233-246)
            bool terminated = false;
            while (true) {
                bool want = live && tri.y != 0;
                unsigned wanting = __ballot_sync(FULL, want);
                if (wanting == 0) break;
                if (__popc(wanting) < postpone_threshold) {
                    if (want) { stack_push(S, local_stack, sp, tri); tri.y = 0; }
                    break;
                }
                if (want) {
                    unsigned ti = msb(tri.y);
                    tri.y &= ~(1u << ti);
                    if (STATS) st_tris++;
                    if (SHADOW) {
                        if (occludes_triangle(P, mesh_id, int(tri.x + ti), ray, hit.t)) { terminated = true; tri.y = 0; }
                    } else {
                        intersect_triangle(P, mesh_id, int(tri.x + ti), ray, hit);
                    }
                }
            }
            if (live) {
                if (SHADOW && terminated) {
                    live = false; sp = 0; cur = make_uint2(0, 0);         // occluded: drop the ray
                } else if ((cur.y & 0xff000000u) == 0) {
                    if (sp == 0) {
                        if (SHADOW) {
                            // unoccluded: deposit the light sample (Pathtracer.cu:183-196)
                            if (STATS) st_miss++;
                            float4 ill = P.sq.illum[ray_index];
                            int px = word_fb_index(P, __float_as_uint(P.sq.od1[ray_index].w));
                            float4 v = make_float4(ill.x, ill.y, ill.z, ill.w);
                            aov_add(P, PTB_AOV_RADIANCE, px, v);
                            if (bounce == 0) aov_set(P, PTB_AOV_RADIANCE_DIRECT, px, v);
                            else             aov_add(P, PTB_AOV_RADIANCE_INDIRECT, px, v);
                        } else {
                            if (hit.mesh_id < -1) hit.mesh_id = __ldg(P.flat_slot_instance + (-(hit.mesh_id) - 2));   // merged slot -> instance
                            q.hit[ray_index] = pack_hit(hit);
                        }
                        live = false; cur = make_uint2(0, 0);
                    } else {
                        if (sp == tlas_sp) {
                            tlas_sp = PTB_INVALID;
                            if (!identity) {   // back to world space: re-read the ray (L2 resident) instead of holding 6 registers
                                float4 a = SHADOW ? P.sq.od0[ray_index] : q.od0[ray_index];
                                float4 b = SHADOW ? P.sq.od1[ray_index] : q.od1[ray_index];
                                ray.o = f3(a.x, a.y, a.z); ray.d = f3(a.w, b.x, b.y);
                                oct4 = ray_octant_inv4(ray.d);
                            }
                        }
                        cur = stack_pop(S, local_stack, sp);
                    }
                }
            }
            unsigned alive = __ballot_sync(FULL, live);
            if (alive == 0) break;
            lost += 32 - __popc(alive) - PTB_DYNFETCH_ND;
            if (lost >= PTB_DYNFETCH_NW) break;
        }
    }
}

// Retirement of merged slots (ptb_update_instances / ptb_refit_instances: an identity instance started moving).  Every triangle record of
// the merged tree whose slot maps to no instance any more is made unhittable: Moeller-Trumbore record p0.x = NaN (s = o - p0 is NaN,
// u is NaN, `u >= 0` fails), Woop record row 2 = NaN (t is NaN, `t > 0` fails).  Runs once per retirement, not per frame.
__global__ void __launch_bounds__(256) k_retire_merged_slots(float4* flat_tris, float4* flat_woop, const int2* flat_who, int n_refs, const int* slot_instance) {
    const float nan = __int_as_float(0x7fc00000);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_refs; i += gridDim.x * blockDim.x) {
        int slot = __float_as_int(flat_tris[3 * size_t(i) + 2].z);
        if (slot_instance[slot] >= 0) continue;
        flat_tris[3 * size_t(i)].x = nan;
        if (flat_woop) flat_woop[3 * size_t(i) + 2] = make_float4(nan, nan, nan, nan);
    }
    (void)flat_who;
}

// ------------------------------------------------------------------------------------------ binary BVH traversal (config 1)
// Src/CUDA/Raytracing/BVH2.h:4-244: ordered descent by split axis, TLAS leaf = instance.
PTB_DI bool node2_hit(float4 a, float4 b, const Ray& ray, float tmax) {
    float3 lo = f3(a.x, a.y, a.z), hi = f3(a.w, b.x, b.y);
    float3 t0 = (lo - ray.o) / ray.d;
    float3 t1 = (hi - ray.o) / ray.d;
    float tn = imin_max(t0.x, t1.x, imin_max(t0.y, t1.y, imin_max(t0.z, t1.z, 0.0f)));
    float tf = imax_min(t0.x, t1.x, imax_min(t0.y, t1.y, imax_min(t0.z, t1.z, tmax)));
    return tn < tf;
}

template <bool SHADOW, bool STATS>
__global__ void __launch_bounds__(PTB_TRACE_BLOCK, PTB_TRACE_MIN_BLOCKS) k_trace2(const __grid_constant__ Frame P, int bounce) {
    const int count = SHADOW ? P.counters->shadow[bounce] : P.counters->trace[bounce];
    int* retired = SHADOW ? &P.counters->retired_shadow[bounce] : &P.counters->retired[bounce];
    const RayQueue& q = P.q[bounce & 1];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned FULL = 0xffffffffu;
    int stack[PTB_STACK_TOTAL];
    unsigned long long st_nodes = 0, st_tris = 0, st_xf = 0, st_rays = 0, st_miss = 0;
    while (true) {
        // one ray per lane per round; refill is warp-aggregated
        int base = 0;
        if (lane == 0) base = atomicAdd(retired, 32);
        base = __shfl_sync(FULL, base, 0);
        if (base >= count) {
            if (STATS) {      // roofline accounting (32-byte nodes, 48-byte triangle tests), same counters as k_trace8<*, true>
                TraceStats* ts = P.trace_stats + (SHADOW ? 1 : 0);
                atomicAdd(&ts->nodes, st_nodes); atomicAdd(&ts->triangles, st_tris); atomicAdd(&ts->instance_transforms, st_xf);
                atomicAdd(&ts->rays, st_rays); atomicAdd(&ts->misses, st_miss);
            }
            return;
        }
        int ray_index = base + int(lane);
        if (ray_index >= count) continue;
        if (STATS) st_rays++;
        float4 a = SHADOW ? P.sq.od0[ray_index] : q.od0[ray_index];
        float4 b = SHADOW ? P.sq.od1[ray_index] : q.od1[ray_index];
        Ray world; world.o = f3(a.x, a.y, a.z); world.d = f3(a.w, b.x, b.y);
        Ray ray = world;
        Hit hit; hit.t = SHADOW ? b.z : PTB_INF; hit.u = hit.v = 0.0f; hit.mesh_id = 0; hit.triangle_id = PTB_INVALID;
        int sp = 0, tlas_sp = PTB_INVALID, mesh_id = 0;
        bool identity = true, occluded = false;
        stack[sp++] = 0;
        while (sp > 0 && !occluded) {
            if (sp == tlas_sp) { tlas_sp = PTB_INVALID; if (!identity) ray = world; }
            int ni = stack[--sp];
            float4 na = __ldg(P.nodes2 + 2 * size_t(ni)), nb = __ldg(P.nodes2 + 2 * size_t(ni) + 1);
            if (STATS) st_nodes++;
            if (!node2_hit(na, nb, ray, hit.t)) continue;
            int first = __float_as_int(nb.z);
            unsigned count_axis = __float_as_uint(nb.w);
            unsigned n = count_axis & 0x3fffffffu, axis = count_axis >> 30;
            if (n > 0) {
                if (tlas_sp == PTB_INVALID) {
                    tlas_sp = sp;
                    mesh_id = first;
                    unsigned root = unsigned(__ldg(P.mesh_roots + mesh_id));
                    identity = (root >> 31) != 0;
                    if (!identity) {
                        Mat3x4 inv = load_mat(P.mesh_transforms_inv, mesh_id);
                        ray.o = xform_pos(inv, ray.o); ray.d = xform_dir(inv, ray.d);
                        if (STATS) st_xf++;
                    }
                    stack[sp++] = int(root & 0x7fffffffu);
                } else {
                    for (int t = first; t < first + int(n); t++) {
                        if (STATS) st_tris++;
                        if (SHADOW) { if (occludes_triangle(P, mesh_id, t, ray, hit.t)) { occluded = true; break; } }
                        else intersect_triangle(P, mesh_id, t, ray, hit);
                    }
                }
            } else {
                float da = axis == 0 ? ray.d.x : (axis == 1 ? ray.d.y : ray.d.z);
                int near_child = da > 0.0f ? first : first + 1;
                int far_child  = da > 0.0f ? first + 1 : first;
                stack[sp++] = far_child; stack[sp++] = near_child;
            }
        }
        if (SHADOW) {
            if (!occluded) {
                if (STATS) st_miss++;
                float4 ill = P.sq.illum[ray_index];
                int px = word_fb_index(P, __float_as_uint(b.w));
                float4 v = make_float4(ill.x, ill.y, ill.z, ill.w);
                aov_add(P, PTB_AOV_RADIANCE, px, v);
                if (bounce == 0) aov_set(P, PTB_AOV_RADIANCE_DIRECT, px, v);
                else             aov_add(P, PTB_AOV_RADIANCE_INDIRECT, px, v);
            }
        } else {
            q.hit[ray_index] = pack_hit(hit);
        }
    }
}

// ------------------------------------------------------------------------------------------ 4-wide BVH traversal (--bvh bvh4)
// Src/CUDA/Raytracing/BVH4.h:4-295.  A stack entry names one child SLOT: (node index | slot << 30); popping it reads that slot's
// (index, count): a leaf is intersected (or, at TLAS level, entered), an internal child has its four boxes tested and the hit ones pushed
// far to near.  The order comes from the reference's trick: the slot number replaces the two low mantissa bits of t_near, and a three-pass
// bubble sort orders the four tagged floats (descending, so the nearest is pushed last); both are replicated exactly, ties included.
struct Quad4Hits { float t_near[4]; bool hit[4]; };
PTB_DI Quad4Hits node4_intersect(const float4* n, const Ray& ray, float tmax) {
    float4 lx = __ldg(n), ly = __ldg(n + 1), lz = __ldg(n + 2), hx = __ldg(n + 3), hy = __ldg(n + 4), hz = __ldg(n + 5);
    float tx0[4] = { (lx.x - ray.o.x) / ray.d.x, (lx.y - ray.o.x) / ray.d.x, (lx.z - ray.o.x) / ray.d.x, (lx.w - ray.o.x) / ray.d.x };
    float tx1[4] = { (hx.x - ray.o.x) / ray.d.x, (hx.y - ray.o.x) / ray.d.x, (hx.z - ray.o.x) / ray.d.x, (hx.w - ray.o.x) / ray.d.x };
    float ty0[4] = { (ly.x - ray.o.y) / ray.d.y, (ly.y - ray.o.y) / ray.d.y, (ly.z - ray.o.y) / ray.d.y, (ly.w - ray.o.y) / ray.d.y };
    float ty1[4] = { (hy.x - ray.o.y) / ray.d.y, (hy.y - ray.o.y) / ray.d.y, (hy.z - ray.o.y) / ray.d.y, (hy.w - ray.o.y) / ray.d.y };
    float tz0[4] = { (lz.x - ray.o.z) / ray.d.z, (lz.y - ray.o.z) / ray.d.z, (lz.z - ray.o.z) / ray.d.z, (lz.w - ray.o.z) / ray.d.z };
    float tz1[4] = { (hz.x - ray.o.z) / ray.d.z, (hz.y - ray.o.z) / ray.d.z, (hz.z - ray.o.z) / ray.d.z, (hz.w - ray.o.z) / ray.d.z };
    Quad4Hits r;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        float tn = imin_max(tx0[c], tx1[c], imin_max(ty0[c], ty1[c], imin_max(tz0[c], tz1[c], 0.0f)));
        float tf = imax_min(tx0[c], tx1[c], imax_min(ty0[c], ty1[c], imax_min(tz0[c], tz1[c], tmax)));
        r.hit[c] = tn < tf;
        r.t_near[c] = __uint_as_float((__float_as_uint(tn) & 0xfffffffcu) | unsigned(c));
    }
#pragma unroll
    for (int i = 1; i < 4; i++) {
#pragma unroll
        for (int j = i - 1; j >= 0; j--) {
            if (r.t_near[j] < r.t_near[j + 1]) { float t = r.t_near[j]; r.t_near[j] = r.t_near[j + 1]; r.t_near[j + 1] = t; }
        }
    }
    return r;
}

template <bool SHADOW, bool STATS>
__global__ void __launch_bounds__(PTB_TRACE_BLOCK, PTB_TRACE_MIN_BLOCKS) k_trace4(const __grid_constant__ Frame P, int bounce) {
    const int count = SHADOW ? P.counters->shadow[bounce] : P.counters->trace[bounce];
    int* retired = SHADOW ? &P.counters->retired_shadow[bounce] : &P.counters->retired[bounce];
    const RayQueue& q = P.q[bounce & 1];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned FULL = 0xffffffffu;
    unsigned stack[PTB_STACK_TOTAL];
    unsigned long long st_nodes = 0, st_tris = 0, st_xf = 0, st_rays = 0, st_miss = 0;
    while (true) {
        int base = 0;                                  // one ray per lane per round; the refill is warp-aggregated
        if (lane == 0) base = atomicAdd(retired, 32);
        base = __shfl_sync(FULL, base, 0);
        if (base >= count) {
            if (STATS) {
                TraceStats* ts = P.trace_stats + (SHADOW ? 1 : 0);
                atomicAdd(&ts->nodes, st_nodes); atomicAdd(&ts->triangles, st_tris); atomicAdd(&ts->instance_transforms, st_xf);
                atomicAdd(&ts->rays, st_rays); atomicAdd(&ts->misses, st_miss);
            }
            return;
        }
        int ray_index = base + int(lane);
        if (ray_index >= count) continue;
        if (STATS) st_rays++;
        float4 a = SHADOW ? P.sq.od0[ray_index] : q.od0[ray_index];
        float4 b = SHADOW ? P.sq.od1[ray_index] : q.od1[ray_index];
        Ray world; world.o = f3(a.x, a.y, a.z); world.d = f3(a.w, b.x, b.y);
        Ray ray = world;
        Hit hit; hit.t = SHADOW ? b.z : PTB_INF; hit.u = hit.v = 0.0f; hit.mesh_id = 0; hit.triangle_id = PTB_INVALID;
        int sp = 0, tlas_sp = PTB_INVALID, mesh_id = 0;
        bool identity = true, occluded = false;
        stack[sp++] = 1u;                              // (node 1, slot 0): the entry slot, pointing at the TLAS root
        while (sp > 0 && !occluded) {
            if (sp == tlas_sp) { tlas_sp = PTB_INVALID; if (!identity) ray = world; }
            unsigned packed = stack[--sp];
            const float4* slots = P.nodes4 + 8 * size_t(packed & 0x3fffffffu) + 6;
            float4 ic = __ldg(slots + ((packed >> 30) >> 1));                       // two (index, count) pairs per float4
            int index = __float_as_int((packed >> 30) & 1u ? ic.z : ic.x), n = __float_as_int((packed >> 30) & 1u ? ic.w : ic.y);
            if (n > 0) {
                if (tlas_sp == PTB_INVALID) {
                    tlas_sp = sp;
                    mesh_id = index;
                    unsigned root = unsigned(__ldg(P.mesh_roots + mesh_id));
                    identity = (root >> 31) != 0;
                    if (!identity) {
                        Mat3x4 inv = load_mat(P.mesh_transforms_inv, mesh_id);
                        ray.o = xform_pos(inv, ray.o); ray.d = xform_dir(inv, ray.d);
                        if (STATS) st_xf++;
                    }
                    stack[sp++] = (root & 0x3fffffffu) + 1u;                        // the BLAS's own entry slot
                } else {
                    for (int t = index; t < index + n; t++) {
                        if (STATS) st_tris++;
                        if (SHADOW) { if (occludes_triangle(P, mesh_id, t, ray, hit.t)) { occluded = true; break; } }
                        else intersect_triangle(P, mesh_id, t, ray, hit);
                    }
                }
            } else {
                if (STATS) st_nodes++;
                Quad4Hits h = node4_intersect(P.nodes4 + 8 * size_t(index), ray, hit.t);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    unsigned id = __float_as_uint(h.t_near[i]) & 3u;
                    bool take = id == 0 ? h.hit[0] : id == 1 ? h.hit[1] : id == 2 ? h.hit[2] : h.hit[3];
                    if (take) stack[sp++] = (id << 30) | unsigned(index);
                }
            }
        }
        if (SHADOW) {
            if (!occluded) {
                if (STATS) st_miss++;
                float4 ill = P.sq.illum[ray_index];
                int px = word_fb_index(P, __float_as_uint(b.w));
                float4 v = make_float4(ill.x, ill.y, ill.z, ill.w);
                aov_add(P, PTB_AOV_RADIANCE, px, v);
                if (bounce == 0) aov_set(P, PTB_AOV_RADIANCE_DIRECT, px, v);
                else             aov_add(P, PTB_AOV_RADIANCE_INDIRECT, px, v);
            }
        } else {
            q.hit[ray_index] = pack_hit(hit);
        }
    }
}

// ------------------------------------------------------------------------------------------ SVGF g-buffers (SVGF.h:61-98)
PTB_DI float2 oct_encode_normal(float3 n) {
    n /= (fabsf(n.x) + fabsf(n.y) + fabsf(n.z));
    if (n.z < 0.0f) {
        n.x = (1.0f - fabsf(n.y)) * (n.x >= 0.0f ? +1.0f : -1.0f);
        n.y = (1.0f - fabsf(n.x)) * (n.y >= 0.0f ? +1.0f : -1.0f);
    }
    return f2(0.5f + 0.5f * n.x, 0.5f + 0.5f * n.y);
}
PTB_DI float3 oct_decode_normal(float2 f) {
    f = f * 2.0f - f2(1.0f, 1.0f);
    float3 n = f3(f.x, f.y, 1.0f - fabsf(f.x) - fabsf(f.y));
    float t = __saturatef(-n.z);
    n.x += n.x >= 0.0 ? -t : t;
    n.y += n.y >= 0.0 ? -t : t;
    return normalize(n);
}
PTB_DI float4 mat4_mul(const float* m, float4 v) {
    return make_float4(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w, m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * v.w,
                       m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * v.w, m[12] * v.x + m[13] * v.y + m[14] * v.z + m[15] * v.w);
}
PTB_DI void svgf_set_gbuffers(const Frame& P, int x, int y, const Hit& hit, float3 hit_point, float3 normal, float3 hit_point_prev) {
    float4 proj = mat4_mul(P.svgf.view_projection, make_float4(hit_point.x, hit_point.y, hit_point.z, 1.0f));
    float4 proj_prev = mat4_mul(P.svgf.view_projection_prev, make_float4(hit_point_prev.x, hit_point_prev.y, hit_point_prev.z, 1.0f));
    float depth = proj.z;
    float depth_prev = proj_prev.z;
    float2 on = oct_encode_normal(normal);
    int px = x + y * P.pitch;
    P.svgf.gbuf_normal_depth[px] = make_float4(on.x, on.y, depth, depth_prev);
    P.svgf.gbuf_ids[px] = make_int2(hit.mesh_id, hit.triangle_id);
    P.svgf.gbuf_screen_prev[px] = f2(proj_prev.x / proj_prev.w, proj_prev.y / proj_prev.w);
}

// ------------------------------------------------------------------------------------------ sort: terminate or classify
// Src/CUDA/Pathtracer.cu:199-463.  Survivors are appended (by index) to their material queue.
PTB_DI bool russian_roulette(const Frame& P, int pixel_index, int fbi, int bounce, int sample_index, float3& throughput) {
    if (bounce == P.config.num_bounces - 1) return true;
    if (P.config.enable_russian_roulette && bounce > 0) {
        float3 t = throughput;
        if (P.config.enable_svgf) t *= f3(aov_get(P, PTB_AOV_ALBEDO, fbi));
        float survival = __saturatef(imax3(t.x, t.y, t.z));
        float r = rng2<DIM_RUSSIAN_ROULETTE>(P, pixel_index, bounce, sample_index).x;
        if (r > survival) return true;
        throughput /= survival;
    }
    return false;
}

PTB_DI void deposit(const Frame& P, int bounce, int px, float3 at_bounce0, float3 illumination) {
    if (bounce == 0) {
        aov_set(P, PTB_AOV_ALBEDO, px, f4(1.0f));
        aov_set(P, PTB_AOV_RADIANCE, px, f4(at_bounce0));
        aov_set(P, PTB_AOV_RADIANCE_DIRECT, px, f4(at_bounce0));
    } else if (bounce == 1) {
        aov_add(P, PTB_AOV_RADIANCE, px, f4(illumination));
        aov_add(P, PTB_AOV_RADIANCE_DIRECT, px, f4(illumination));
    } else {
        aov_add(P, PTB_AOV_RADIANCE, px, f4(illumination));
        aov_add(P, PTB_AOV_RADIANCE_INDIRECT, px, f4(illumination));
    }
}

#ifndef PTB_SORT_MIN_BLOCKS
#define PTB_SORT_MIN_BLOCKS 4
#endif
#ifndef PTB_SORT_PREFETCH
#define PTB_SORT_PREFETCH 1     // measured: k_sort 1.44 -> 1.33 ms per frame (profiles/r2_summary.md)
#endif
PTB_DI void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
// one atomic per group of lanes that are active at the call site (divergent code: the rare medium-scatter emission of k_sort)
PTB_DI int append_active(int* counter) {
    unsigned act = __activemask();
    unsigned lane = threadIdx.x & 31u;
    int leader = __ffs(act) - 1, base = 0;
    if (int(lane) == leader) base = atomicAdd(counter, __popc(act));
    base = __shfl_sync(act, base, leader);
    return base + __popc(act & ((1u << lane) - 1u));
}
__global__ void __launch_bounds__(256, PTB_SORT_MIN_BLOCKS) k_sort(const __grid_constant__ Frame P, int bounce) {
    const RayQueue& q = P.q[bounce & 1];
    const RayQueue& qn = P.q[(bounce + 1) & 1];
    const int count = P.counters->trace[bounce];
    const int rounded = (count + 31) & ~31;     // whole warps iterate together so the aggregated appends stay converged
    for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < rounded; index += gridDim.x * blockDim.x) {
        int dest = -1;                            // material queue to join: 0..3, -1 = path ended (or scattered)
#if PTB_SORT_PREFETCH
        {   // the kernel waits on HBM (ncu: long_scoreboard): pull the next grid-stride iteration's records into L2 while this one is processed
            const int nxt = index + PTB_SORT_PREFETCH * int(gridDim.x * blockDim.x);
            if (nxt < count) {
                prefetch_l2(q.od0 + nxt); prefetch_l2(q.od1 + nxt); prefetch_l2(q.hit + nxt);
                if (bounce > 0) prefetch_l2(q.path + nxt);
                if ((threadIdx.x & 7u) == 0) prefetch_l2(q.pix + nxt);
            }
        }
#endif
        if (index < count) {
            float4 a = q.od0[index], b = q.od1[index];
            float3 ray_direction = f3(a.w, b.x, b.y);
            Hit hit = unpack_hit(__ldg(q.hit + index));
            float cone_angle = b.z, cone_width = b.w;
            unsigned pf = q.pix[index];
            int pixel_index = word_pixel(P, pf);
            const int fbi = word_fb_index(P, pf);
            const int sample_index = P.first_sample + word_slot(P, pf);
            bool allow_nee = (pf & PTB_FLAG_ALLOW_NEE) != 0;
            bool inside_medium = (pf & PTB_FLAG_INSIDE_MEDIUM) != 0;
            float4 path = bounce == 0 ? make_float4(1.0f, 1.0f, 1.0f, 0.0f) : q.path[index];
            float3 throughput = f3(path.x, path.y, path.z);
            bool ended = false;

            int medium_id = PTB_INVALID;
            if (inside_medium) {                                       // Pathtracer.cu:252-325
                medium_id = q.medium[index];
                float4 m0 = __ldg(P.media + 2 * medium_id), m1 = __ldg(P.media + 2 * medium_id + 1);
                float3 sigma_a = f3(m0.x, m0.y, m0.z), sigma_s = f3(m1.x, m1.y, m1.z);
                float g = m0.w;
                if ((sigma_s.x + sigma_s.y + sigma_s.z) > 0.0f) {
                    float2 rs = rng2<DIM_BSDF_0>(P, pixel_index, bounce, sample_index);
                    float2 rp = rng2<DIM_BSDF_1>(P, pixel_index, bounce, sample_index);
                    float3 sigma_t = sigma_a + sigma_s;
                    float tsum = throughput.x + throughput.y + throughput.z;
                    float3 wpdf = throughput / tsum;
                    float st;
                    if (rs.x * tsum < throughput.x) st = sigma_t.x;
                    else if (rs.x * tsum < throughput.x + throughput.y) st = sigma_t.y;
                    else st = sigma_t.z;
                    float dist = sample_exp(st, rs.y);
                    float tt = fminf(dist, hit.t);
                    float3 tr = f3(expf(-sigma_t.x * tt), expf(-sigma_t.y * tt), expf(-sigma_t.z * tt));
                    if (dist < hit.t) {
                        float3 pdf = wpdf * sigma_t * tr;
                        throughput *= sigma_s * tr / (pdf.x + pdf.y + pdf.z);
                        if (!russian_roulette(P, pixel_index, fbi, bounce, sample_index, throughput)) {
                            float3 dir_out = sample_henyey_greenstein(-ray_direction, g, rp.x, rp.y);
                            float3 org = f3(a.x, a.y, a.z) + dist * ray_direction;
                            if (P.config.enable_mipmapping && bounce == 0) {
                                cone_angle = P.camera.pixel_spread_angle;
                                cone_width = P.camera.pixel_spread_angle * dist;
                            }
                            // medium scattering emits straight into the next trace queue (rare: appended from inside the branch, so
                            // that its 14 values are not live registers of the common path)
                            int slot = append_active(&P.counters->trace[bounce + 1]);
                            qn.od0[slot] = make_float4(org.x, org.y, org.z, dir_out.x);
                            qn.od1[slot] = make_float4(dir_out.y, dir_out.z, cone_angle, cone_width);
                            qn.path[slot] = make_float4(throughput.x, throughput.y, throughput.z, 0.0f);
                            qn.pix[slot] = (pf & ~PTB_FLAGS_ALL) | PTB_FLAG_INSIDE_MEDIUM;
                            qn.medium[slot] = medium_id;
                        }
                        ended = true;
                    } else {
                        float3 pdf = wpdf * tr;
                        throughput *= tr / (pdf.x + pdf.y + pdf.z);
                    }
                } else {
                    throughput *= f3(expf(-sigma_a.x * hit.t), expf(-sigma_a.y * hit.t), expf(-sigma_a.z * hit.t));
                }
            }

            if (!ended && hit.triangle_id == PTB_INVALID) {            // miss: sample the sky (Pathtracer.cu:327-343)
                float3 ill = throughput * sample_sky(P, ray_direction);
                deposit(P, bounce, fbi, ill, ill);
                ended = true;
            }
            if (!ended) {
                if (bounce == 0 && P.pixel_query[0] == pixel_index && word_slot(P, pf) == 0) {     // GUI picking (Pathtracer.cu:345-348); pixel_index is -1 unless a query is pending
                    P.pixel_query[1] = hit.mesh_id; P.pixel_query[2] = hit.triangle_id;
                }
                int material_id = P.mesh_material_ids[hit.mesh_id];
                int mtype = P.material_types[material_id];
                if (mtype == PTB_MAT_LIGHT) {                          // Pathtracer.cu:354-422
                    TriPos lt = load_tri_pos(P, hit.triangle_id);
                    float3 light_point = barycentric(hit.u, hit.v, lt.p0, lt.e1, lt.e2);
                    float3 light_point_prev = light_point;
                    float3 lgn = cross(lt.e1, lt.e2);
                    Mat3x4 world = load_mat(P.mesh_transforms, hit.mesh_id);
                    light_point = xform_pos(world, light_point);
                    lgn = xform_dir(world, lgn);
                    lgn = normalize(lgn);
                    if (bounce == 0 && P.config.enable_svgf) {
                        Mat3x4 wprev = load_mat(P.mesh_transforms_prev, hit.mesh_id);
                        light_point_prev = xform_pos(wprev, light_point_prev);
                        svgf_set_gbuffers(P, pixel_index % P.pitch, pixel_index / P.pitch, hit, light_point, lgn, light_point_prev);
                    }
                    float3 emission = f3(__ldg(P.materials + 2 * material_id));
                    bool count_it = P.config.enable_next_event_estimation ? !allow_nee : true;
                    if (count_it) {
                        float3 ill = throughput * emission;
                        deposit(P, bounce, fbi, emission, ill);
                    } else if (P.config.enable_multiple_importance_sampling) {
                        float cos_l = abs_dot(ray_direction, lgn);
                        float d2 = hit.t * hit.t;
                        float brdf_pdf = path.w;
                        float power = luminance(emission.x, emission.y, emission.z);
                        float light_pdf = power * d2 / (cos_l * P.lights_total_weight);
                        if (pdf_is_valid(light_pdf)) {
                            float w = power_heuristic(brdf_pdf, light_pdf);
                            float3 ill = throughput * emission * w;
                            aov_add(P, PTB_AOV_RADIANCE, fbi, f4(ill));
                            if (bounce == 1) aov_add(P, PTB_AOV_RADIANCE_DIRECT, fbi, f4(ill));
                            else             aov_add(P, PTB_AOV_RADIANCE_INDIRECT, fbi, f4(ill));
                        }
                    }
                    ended = true;
                } else if (!russian_roulette(P, pixel_index, fbi, bounce, sample_index, throughput)) {
                    dest = mtype - PTB_MAT_DIFFUSE;
                    if (bounce > 0) q.path[index] = make_float4(throughput.x, throughput.y, throughput.z, path.w);  // roulette rescale, in place
                    // flags for the shade pass: only the medium bit survives (Pathtracer.cu:450-451)
                    q.pix[index] = (pf & ~PTB_FLAGS_ALL) | (medium_id != PTB_INVALID ? PTB_FLAG_INSIDE_MEDIUM : 0u);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < 4; m++) {
            int slot = warp_append(&P.counters->mat[m][bounce], dest == m);
            if (dest == m) P.matq[m][slot] = index;
        }
    }
}

// ------------------------------------------------------------------------------------------ BSDFs (BSDF.h:8-525)
struct ShadeCtx {
    int pixel_index, bounce, sample_index;
    float3 tangent, bitangent, normal, omega_i;
};

struct BSDFDiffuse {
    static constexpr bool HAS_ALBEDO = true;
    static constexpr int QUEUE = 0;
    float3 diffuse; int texture_id; float3 albedo;
    PTB_DI void init(const Frame& P, const ShadeCtx&, bool, int material_id) {
        float4 m = __ldg(P.materials + 2 * material_id);
        diffuse = f3(m.x, m.y, m.z); texture_id = __float_as_int(m.w);
    }
    PTB_DI bool has_texture() const { return texture_id != PTB_INVALID; }
    PTB_DI bool allow_nee() const { return true; }
    PTB_DI void apply_albedo(const Frame& P, const ShadeCtx& c, float3& throughput) const {
        if (!(P.config.enable_svgf && c.bounce == 0)) throughput *= albedo;
    }
    PTB_DI bool eval(const Frame&, const ShadeCtx&, float3, float cos_o, float3& bsdf, float& pdf) const {
        if (cos_o <= 0.0f) return false;
        bsdf = f3(cos_o * PTB_ONE_OVER_PI);
        pdf = cos_o * PTB_ONE_OVER_PI;
        return pdf_is_valid(pdf);
    }
    PTB_DI bool sample(const Frame& P, const ShadeCtx& c, float3&, int&, float3& dir_out, float& pdf) const {
        float2 r = rng2<DIM_BSDF_0>(P, c.pixel_index, c.bounce, c.sample_index);
        float3 wo = sample_cosine_hemisphere(r.x, r.y);
        dir_out = local_to_world(wo, c.tangent, c.bitangent, c.normal);
        pdf = wo.z * PTB_ONE_OVER_PI;
        return pdf_is_valid(pdf);
    }
};

struct BSDFPlastic {
    static constexpr bool HAS_ALBEDO = true;
    static constexpr int QUEUE = 1;
    float3 diffuse; int texture_id; float roughness; float3 albedo;
    PTB_DI void init(const Frame& P, const ShadeCtx&, bool, int material_id) {
        float4 m = __ldg(P.materials + 2 * material_id);
        diffuse = f3(m.x, m.y, m.z); texture_id = __float_as_int(m.w);
        roughness = __ldg(reinterpret_cast<const float*>(P.materials + 2 * material_id + 1));
    }
    PTB_DI bool has_texture() const { return texture_id != PTB_INVALID; }
    PTB_DI bool allow_nee() const { return true; }
    PTB_DI void apply_albedo(const Frame&, const ShadeCtx&, float3&) const {}
    PTB_DI bool eval(const Frame&, const ShadeCtx& c, float3 to_light, float cos_o, float3& bsdf, float& pdf) const {
        if (cos_o <= 0.0f) return false;
        const float IOR = 1.5f, ETA = 1.0f / IOR;
        float3 wi = c.omega_i;
        float3 wo = world_to_local(to_light, c.tangent, c.bitangent, c.normal);
        float3 wm = normalize(wi + wo);
        float ax = roughness_to_alpha(roughness), ay = roughness_to_alpha(roughness);
        float F = fresnel_dielectric(dot(wi, wm), ETA);
        float D = ggx_D(wm, ax, ay);
        float G1 = ggx_G1(wi, ax, ay);
        float G2 = ggx_G2(wo, wi, wm, ax, ay);
        float3 spec = f3(F * G2 * D / (4.0f * wi.z));
        float F_i = fresnel_dielectric(wi.z, ETA);
        float F_o = fresnel_dielectric(wo.z, ETA);
        float F_avg = average_fresnel(IOR);
        float isf = 1.0f - (1.0f - F_avg) * square(ETA);
        float3 diff = ETA * ETA * (1.0f - F_i) * (1.0f - F_o) * albedo * PTB_ONE_OVER_PI / (1.0f - albedo * isf) * wo.z;
        float pdf_s = G1 * D / (4.0f * wi.z);
        float pdf_d = wo.z * PTB_ONE_OVER_PI;
        pdf = lerpf(pdf_d, pdf_s, F_i);
        bsdf = spec + diff;
        return pdf_is_valid(pdf);
    }
    PTB_DI bool sample(const Frame& P, const ShadeCtx& c, float3& throughput, int&, float3& dir_out, float& pdf) const {
        const float IOR = 1.5f, ETA = 1.0f / IOR;
        float rf = rng2<DIM_BSDF_0>(P, c.pixel_index, c.bounce, c.sample_index).x;
        float2 rb = rng2<DIM_BSDF_1>(P, c.pixel_index, c.bounce, c.sample_index);
        float3 wi = c.omega_i;
        float F_i = fresnel_dielectric(wi.z, ETA);
        float ax = roughness_to_alpha(roughness), ay = roughness_to_alpha(roughness);
        float3 wm, wo;
        if (rf < F_i) { wm = sample_vndf_ggx(wi, ax, ay, rb.x, rb.y); wo = reflect_direction(wi, wm); }
        else          { wo = sample_cosine_hemisphere(rb.x, rb.y);    wm = normalize(wi + wo); }
        if (wm.z < 0.0f) return false;
        float F = fresnel_dielectric(dot(wi, wm), ETA);
        float D = ggx_D(wm, ax, ay);
        float G1 = ggx_G1(wi, ax, ay);
        float G2 = ggx_G2(wo, wi, wm, ax, ay);
        float3 spec = f3(F * G2 * D / (4.0f * wi.z));
        float F_o = fresnel_dielectric(wo.z, ETA);
        float F_avg = average_fresnel(IOR);
        float isf = 1.0f - (1.0f - F_avg) * square(ETA);
        float3 diff = ETA * ETA * (1.0f - F_i) * (1.0f - F_o) * albedo * PTB_ONE_OVER_PI / (1.0f - albedo * isf) * wo.z;
        float pdf_s = G1 * D / (4.0f * wi.z);
        float pdf_d = wo.z * PTB_ONE_OVER_PI;
        pdf = lerpf(pdf_d, pdf_s, F_i);
        throughput *= (spec + diff) / pdf;
        dir_out = local_to_world(wo, c.tangent, c.bitangent, c.normal);
        return pdf_is_valid(pdf);
    }
};

struct BSDFDielectric {
    static constexpr bool HAS_ALBEDO = false;
    static constexpr int QUEUE = 2;
    int medium_id; float ior, roughness, eta; float3 albedo;
    PTB_DI void init(const Frame& P, const ShadeCtx&, bool entering, int material_id) {
        float4 m = __ldg(P.materials + 2 * material_id);
        medium_id = __float_as_int(m.x); ior = m.y; roughness = m.z;
        eta = entering ? 1.0f / ior : ior;
    }
    PTB_DI bool has_texture() const { return false; }
    PTB_DI bool allow_nee() const { return roughness >= PTB_ROUGHNESS_CUTOFF; }
    PTB_DI void apply_albedo(const Frame&, const ShadeCtx&, float3&) const {}

    // shared tail of eval() and sample(): single- and multi-scatter lobes for a given (wi, wo, wm)
    PTB_DI void lobes(const Frame& P, float3 wi, float3 wo, float3 wm, bool reflected, float F, float E_i, float ratio,
                      float E_avg_enter, float E_avg_leave, bool entering, float& bsdf_s, float& bsdf_m, float& pdf_s, float& pdf_m) const {
        float ax = roughness_to_alpha(roughness), ay = roughness_to_alpha(roughness);
        float D = ggx_D(wm, ax, ay);
        float G1 = ggx_G1(wi, ax, ay);
        float G2 = ggx_G2(wo, wi, wm, ax, ay);
        float i_m = abs_dot(wi, wm), o_m = abs_dot(wo, wm);
        if (reflected) {
            bsdf_s = F * G2 * D / (4.0f * wi.z);
            pdf_s = F * G1 * D / (4.0f * wi.z);
            float E_o = dielectric_directional_albedo(P, ior, roughness, wo.z, entering);
            float E_avg = entering ? E_avg_enter : E_avg_leave;
            bsdf_m = (1.0f - ratio) * fabsf(wo.z) * kulla_conty_lobe(E_i, E_o, E_avg);
            pdf_m = (1.0f - ratio) * fabsf(wo.z) * PTB_ONE_OVER_PI;
        } else {
            bsdf_s = (1.0f - F) * G2 * D * i_m * o_m / (wi.z * square(eta * i_m + o_m) * square(eta));
            pdf_s = (1.0f - F) * G1 * D * i_m * o_m / (wi.z * square(eta * i_m + o_m));
            float E_o = dielectric_directional_albedo(P, ior, roughness, wo.z, !entering);
            float E_avg = entering ? E_avg_leave : E_avg_enter;
            bsdf_m = ratio * fabsf(wo.z) * kulla_conty_lobe(E_i, E_o, E_avg);
            pdf_m = ratio * fabsf(wo.z) * PTB_ONE_OVER_PI;
        }
    }
    PTB_DI bool eval(const Frame& P, const ShadeCtx& c, float3 to_light, float, float3& bsdf, float& pdf) const {
        float3 wi = c.omega_i;
        float3 wo = world_to_local(to_light, c.tangent, c.bitangent, c.normal);
        bool reflected = wo.z >= 0.0f;
        float3 wm = reflected ? normalize(wi + wo) : normalize(eta * wi + wo);
        wm *= sign1(wm.z);
        float i_m = abs_dot(wi, wm);
        float F = fresnel_dielectric(i_m, eta);
        bool entering = eta < 1.0f;
        float F_avg = average_fresnel(ior);
        if (!entering) F_avg = 1.0f - (1.0f - F_avg) / square(ior);
        float Ee = dielectric_albedo(P, ior, roughness, true), El = dielectric_albedo(P, ior, roughness, false);
        float x = kulla_conty_reciprocity(Ee, El);
        float ratio = (entering ? x : (1.0f - x)) * (1.0f - F_avg);
        float E_i = dielectric_directional_albedo(P, ior, roughness, wi.z, entering);
        float bs, bm, ps, pm;
        lobes(P, wi, wo, wm, reflected, F, E_i, ratio, Ee, El, entering, bs, bm, ps, pm);
        bsdf = f3(bs + bm);
        pdf = lerpf(pm, ps, E_i);
        return pdf_is_valid(pdf);
    }
    PTB_DI bool sample(const Frame& P, const ShadeCtx& c, float3& throughput, int& medium, float3& dir_out, float& pdf) const {
        float2 r0 = rng2<DIM_BSDF_0>(P, c.pixel_index, c.bounce, c.sample_index);
        float2 r1 = rng2<DIM_BSDF_1>(P, c.pixel_index, c.bounce, c.sample_index);
        float3 wi = c.omega_i;
        float ax = roughness_to_alpha(roughness), ay = roughness_to_alpha(roughness);
        bool entering = eta < 1.0f;
        float E_i = dielectric_directional_albedo(P, ior, roughness, wi.z, entering);
        float F_avg = average_fresnel(ior);
        if (!entering) F_avg = 1.0f - (1.0f - F_avg) / square(ior);
        float Ee = dielectric_albedo(P, ior, roughness, true), El = dielectric_albedo(P, ior, roughness, false);
        float x = kulla_conty_reciprocity(Ee, El);
        float ratio = (entering ? x : (1.0f - x)) * (1.0f - F_avg);
        float F; bool reflected; float3 wm, wo;
        if (r0.x < E_i) {
            wm = sample_vndf_ggx(wi, ax, ay, r1.x, r1.y);
            F = fresnel_dielectric(abs_dot(wi, wm), eta);
            reflected = r0.y < F;
            wo = reflected ? reflect_direction(wi, wm) : refract_direction(wi, wm, eta);
        } else {
            wo = sample_cosine_hemisphere(r1.x, r1.y);
            reflected = r0.y > ratio;
            if (reflected) wm = normalize(wi + wo);
            else { wo = -wo; wm = normalize(eta * wi + wo); }
            wm *= sign1(wm.z);
            F = fresnel_dielectric(abs_dot(wi, wm), eta);
        }
        if (reflected ^ (wo.z >= 0.0f)) return false;
        float bs, bm, ps, pm;
        lobes(P, wi, wo, wm, reflected, F, E_i, ratio, Ee, El, entering, bs, bm, ps, pm);
        if (!reflected) medium = entering ? medium_id : PTB_INVALID;
        pdf = lerpf(pm, ps, E_i);
        throughput *= (bs + bm) / pdf;
        dir_out = local_to_world(wo, c.tangent, c.bitangent, c.normal);
        return pdf_is_valid(pdf);
    }
};

struct BSDFConductor {
    static constexpr bool HAS_ALBEDO = false;
    static constexpr int QUEUE = 3;
    float3 eta, k; float roughness; float3 albedo;
    PTB_DI void init(const Frame& P, const ShadeCtx&, bool, int material_id) {
        float4 m0 = __ldg(P.materials + 2 * material_id), m1 = __ldg(P.materials + 2 * material_id + 1);
        eta = f3(m0.x, m0.y, m0.z); roughness = m0.w; k = f3(m1.x, m1.y, m1.z);
    }
    PTB_DI bool has_texture() const { return false; }
    PTB_DI bool allow_nee() const { return roughness >= PTB_ROUGHNESS_CUTOFF; }
    PTB_DI void apply_albedo(const Frame&, const ShadeCtx&, float3&) const {}
    PTB_DI void lobes(const Frame& P, float3 wi, float3 wo, float3 wm, float o_m, float E_i, float3& bsdf, float& pdf) const {
        float ax = roughness_to_alpha(roughness), ay = roughness_to_alpha(roughness);
        float3 F = fresnel_conductor(o_m, eta, k);
        float D = ggx_D(wm, ax, ay);
        float G1 = ggx_G1(wi, ax, ay);
        float G2 = ggx_G2(wo, wi, wm, ax, ay);
        float3 single = F * G2 * D / (4.0f * wi.z);
        float pdf_s = G1 * D / (4.0f * wi.z);
        float E_o = conductor_directional_albedo(P, roughness, wo.z);
        float E_avg = conductor_albedo(P, roughness);
        float3 F_avg = average_fresnel(eta, k);
        float3 F_ms = fresnel_multiscatter(F_avg, E_avg);
        float3 multi = F_ms * kulla_conty_lobe(E_i, E_o, E_avg) * wo.z;
        float pdf_m = wo.z * PTB_ONE_OVER_PI;
        bsdf = single + multi;
        pdf = lerpf(pdf_m, pdf_s, E_i);
    }
    PTB_DI bool eval(const Frame& P, const ShadeCtx& c, float3 to_light, float cos_o, float3& bsdf, float& pdf) const {
        if (cos_o <= 0.0f) return false;
        float3 wi = c.omega_i;
        float3 wo = world_to_local(to_light, c.tangent, c.bitangent, c.normal);
        float3 wm = normalize(wo + wi);
        float o_m = dot(wo, wm);
        if (o_m <= 0.0f) return false;
        float E_i = conductor_directional_albedo(P, roughness, wi.z);
        lobes(P, wi, wo, wm, o_m, E_i, bsdf, pdf);
        return pdf_is_valid(pdf);
    }
    PTB_DI bool sample(const Frame& P, const ShadeCtx& c, float3& throughput, int&, float3& dir_out, float& pdf) const {
        float2 r0 = rng2<DIM_BSDF_0>(P, c.pixel_index, c.bounce, c.sample_index);
        float2 r1 = rng2<DIM_BSDF_1>(P, c.pixel_index, c.bounce, c.sample_index);
        float3 wi = c.omega_i;
        float ax = roughness_to_alpha(roughness), ay = roughness_to_alpha(roughness);
        float E_i = conductor_directional_albedo(P, roughness, wi.z);
        float3 wm, wo;
        if (r0.x < E_i) { wm = sample_vndf_ggx(wi, ax, ay, r1.x, r1.y); wo = reflect_direction(wi, wm); }
        else            { wo = sample_cosine_hemisphere(r1.x, r1.y);    wm = normalize(wi + wo); }
        float o_m = dot(wo, wm);
        if (o_m <= 0.0f || wo.z < 0.0f) return false;
        float3 bsdf;
        lobes(P, wi, wo, wm, o_m, E_i, bsdf, pdf);
        throughput *= bsdf / pdf;
        dir_out = local_to_world(wo, c.tangent, c.bitangent, c.normal);
        return pdf_is_valid(pdf);
    }
};

// ------------------------------------------------------------------------------------------ ray cones (RayCone.h)
PTB_DI float triangle_curvature(float3 pe1, float3 pe2, float3 ne1, float3 ne2) {
    float3 ne0 = ne1 - ne2, pe0 = pe1 - pe2;
    float k01 = dot(ne1, pe1) / dot(pe1, pe1);
    float k02 = dot(ne2, pe2) / dot(pe2, pe2);
    float k12 = dot(ne0, pe0) / dot(pe0, pe0);
    return (k01 + k02 + k12) * (1.0f / 3.0f);
}
PTB_DI void ray_cone_ellipse_axes(float3 rd, float3 gn, float cone_width, float3& a1, float3& a2) {
    float3 h1 = rd - dot(gn, rd) * gn;
    float3 h2 = cross(gn, h1);
    a1 = cone_width / fmaxf(0.0001f, length(h1 - dot(rd, h1) * rd)) * h1;
    a2 = cone_width / fmaxf(0.0001f, length(h2 - dot(rd, h2) * rd)) * h2;
}
PTB_DI float2 ellipse_axis_to_gradient(const TriFull& t, float inv_2area, float3 gn, float3 hit_point, float2 hit_uv, float3 axis) {
    float3 ep = hit_point + axis - t.p0;
    float u = dot(gn, cross(ep, t.e2)) * inv_2area;
    float v = dot(gn, cross(t.e1, ep)) * inv_2area;
    return barycentric(u, v, t.t0, t.te1, t.te2) - hit_uv;
}

// ------------------------------------------------------------------------------------------ shade + NEE + extend
// Src/CUDA/Pathtracer.cu:465-757.  One kernel instantiation per BSDF; shadow rays and extension rays are appended
// with warp-aggregated atomics.
template <typename BSDF> struct ShadeOccupancy { static constexpr int min_blocks = 2; };        // microfacet BSDFs: ~120 registers
template <> struct ShadeOccupancy<BSDFDiffuse> { static constexpr int min_blocks = PTB_SHADE_MIN_BLOCKS_DIFFUSE; };
template <typename BSDF>
__global__ void __launch_bounds__(256, ShadeOccupancy<BSDF>::min_blocks) k_shade(const __grid_constant__ Frame P, int bounce) {
    const RayQueue& q = P.q[bounce & 1];
    const RayQueue& qn = P.q[(bounce + 1) & 1];
    const int count = P.counters->mat[BSDF::QUEUE][bounce];
    const int rounded = (count + 31) & ~31;
    const int* queue = P.matq[BSDF::QUEUE];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += gridDim.x * blockDim.x) {
        bool emit_shadow = false, emit_next = false;
        float4 sh0, sh1, sh_ill, nx0, nx1, nx_path; unsigned nx_pix = 0; int nx_medium = PTB_INVALID;
        if (i < count) {
            const int index = queue[i];
            float4 a = q.od0[index], b = q.od1[index];
            float3 ray_direction = f3(a.w, b.x, b.y);
            // The reference copies the hit from the trace buffer into the material buffer (kernel_sort) and so packs the
            // barycentrics to 16 bits a SECOND time (HitBuffer::set after HitBuffer::get, Buffers.h:28-47); that round trip is
            // not the identity (int(q / 65535.0f * 65535.0f) can be q - 1).  We compact by index, so replay it here.
            Hit hit = unpack_hit(pack_hit(unpack_hit(__ldg(q.hit + index))));
            unsigned pf = q.pix[index];
            int pixel_index = word_pixel(P, pf);
            const int fbi = word_fb_index(P, pf);
            const int sample_index = P.first_sample + word_slot(P, pf);
            int medium_id = (pf & PTB_FLAG_INSIDE_MEDIUM) ? q.medium[index] : PTB_INVALID;
            float3 throughput = f3(1.0f);
            if (bounce > 0) { float4 p = q.path[index]; throughput = f3(p.x, p.y, p.z); }

            TriFull tri = load_tri_full(P, hit.triangle_id);
            float3 hit_point = barycentric(hit.u, hit.v, tri.p0, tri.e1, tri.e2);
            float3 normal = barycentric(hit.u, hit.v, tri.n0, tri.ne1, tri.ne2);
            float2 tex_coord = barycentric(hit.u, hit.v, tri.t0, tri.te1, tri.te2);
            float3 hit_point_local = hit_point;
            Mat3x4 world = load_mat(P.mesh_transforms, hit.mesh_id);
            hit_point = xform_pos(world, hit_point);
            normal = xform_dir(world, normal);
            normal = normalize(normal);
            float mesh_scale_inv = 1.0f / length(f3(world.r0.x, world.r0.y, world.r0.z));

            float cone_angle = 0.0f, cone_width = 0.0f, curvature = 0.0f;
            if (P.config.enable_mipmapping) {
                if (bounce == 0) { cone_angle = P.camera.pixel_spread_angle; cone_width = cone_angle * hit.t; }
                else             { cone_angle = b.z; cone_width = b.w + cone_angle * hit.t; }
                curvature = triangle_curvature(tri.e1, tri.e2, tri.ne1, tri.ne2) * mesh_scale_inv;
            }
            tri.e1 = xform_dir(world, tri.e1);
            tri.e2 = xform_dir(world, tri.e2);
            float3 gn = cross(tri.e1, tri.e2);
            float inv_2area = 1.0f / length(gn);
            gn *= inv_2area;
            bool entering = dot(ray_direction, gn) < 0.0f;
            if (!entering) { normal = -normal; curvature = -curvature; }

            ShadeCtx c;
            c.pixel_index = pixel_index; c.bounce = bounce; c.sample_index = sample_index;
            orthonormal_basis(normal, c.tangent, c.bitangent);
            c.normal = normal;
            c.omega_i = world_to_local(-ray_direction, c.tangent, c.bitangent, normal);

            if (c.omega_i.z > 0.0f) {
                int material_id = P.mesh_material_ids[hit.mesh_id];
                BSDF bsdf;
                bsdf.init(P, c, entering, material_id);

                if constexpr (BSDF::HAS_ALBEDO) {
                    // albedo texture through ray cones: anisotropic gradients at bounce 0, isotropic LOD after (RayCone.h:16-36)
                    float3 base = bsdf.diffuse;
                    int tid = bsdf.texture_id;
                    float3 albedo;
                    if (tid == PTB_INVALID) {
                        albedo = base;
                    } else {
                        TextureEntry te = P.textures[tid];
                        float4 texel;
                        if (P.config.enable_mipmapping) {
                            if (bounce == 0) {
                                float3 ax1, ax2;
                                ray_cone_ellipse_axes(ray_direction, gn, cone_width, ax1, ax2);
                                float2 g1 = ellipse_axis_to_gradient(tri, inv_2area, gn, hit_point, tex_coord, ax1);
                                float2 g2 = ellipse_axis_to_gradient(tri, inv_2area, gn, hit_point, tex_coord, ax2);
                                texel = fetch_per_handle(te.tex, [&](cudaTextureObject_t t) { return tex2DGrad<float4>(t, tex_coord.x, tex_coord.y, g1, g2); });
                            } else {
                                float lod_tri = sqrtf(fabsf(tri.te1.x * tri.te2.y - tri.te2.x * tri.te1.y) * inv_2area);
                                float lod_cone = fabsf(cone_width / dot(ray_direction, gn));
                                float lod = log2f(lod_tri * lod_cone);
                                float level = lod + te.lod_bias;
                                texel = fetch_per_handle(te.tex, [&](cudaTextureObject_t t) { return tex2DLod<float4>(t, tex_coord.x, tex_coord.y, level); });
                            }
                        } else {
                            texel = fetch_per_handle(te.tex, [&](cudaTextureObject_t t) { return tex2D<float4>(t, tex_coord.x, tex_coord.y); });
                        }
                        albedo = base * f3(texel);
                    }
                    bsdf.albedo = albedo;
                    if (bounce == 0) aov_set(P, PTB_AOV_ALBEDO, fbi, f4(albedo));
                    bsdf.apply_albedo(P, c, throughput);
                } else {
                    if (bounce == 0) aov_set(P, PTB_AOV_ALBEDO, fbi, f4(1.0f));
                }
                if (bounce == 0) {
                    aov_set(P, PTB_AOV_NORMAL, fbi, f4(normal));
                    aov_set(P, PTB_AOV_POSITION, fbi, f4(hit_point));
                }
                if (P.config.enable_mipmapping) cone_angle -= 2.0f * curvature * fabsf(cone_width) / dot(normal, ray_direction);

                if (bounce == 0 && P.config.enable_svgf) {
                    float3 hp_prev = hit_point_local;
                    Mat3x4 wprev = load_mat(P.mesh_transforms_prev, hit.mesh_id);
                    hp_prev = xform_pos(wprev, hp_prev);
                    svgf_set_gbuffers(P, pixel_index % P.pitch, pixel_index / P.pitch, hit, hit_point, normal, hp_prev);
                }

                // ---- next event estimation (Pathtracer.cu:465-555)
                if (P.config.enable_next_event_estimation && P.lights_total_weight > 0.0f && bsdf.allow_nee()) {
                    float2 rl = rng2<DIM_NEE_LIGHT>(P, pixel_index, bounce, sample_index);
                    float2 rt = rng2<DIM_NEE_TRIANGLE>(P, pixel_index, bounce, sample_index);
                    int light_mesh;
                    int light_tri = sample_light(P, rl.x, rl.y, light_mesh);
                    float2 luv = sample_triangle(rt.x, rt.y);
                    TriPos lt = load_tri_pos(P, light_tri);
                    float3 light_point = barycentric(luv.x, luv.y, lt.p0, lt.e1, lt.e2);
                    float3 lgn = cross(lt.e1, lt.e2);
                    Mat3x4 lw = load_mat(P.mesh_transforms, light_mesh);
                    light_point = xform_pos(lw, light_point);
                    lgn = xform_dir(lw, lgn);
                    lgn = normalize(lgn);
                    float3 hp = ray_origin_epsilon_offset(hit_point, light_point - hit_point, gn);
                    light_point = ray_origin_epsilon_offset(light_point, hp - light_point, lgn);
                    float3 to_light = light_point - hp;
                    float dist = length(to_light);
                    to_light /= dist;
                    float cos_l = abs_dot(to_light, lgn);
                    float cos_h = dot(to_light, normal);
                    int lmat = P.mesh_material_ids[light_mesh];
                    float3 emission = f3(__ldg(P.materials + 2 * lmat));
                    float3 bv; float bp;
                    if (bsdf.eval(P, c, to_light, cos_h, bv, bp)) {
                        float power = luminance(emission.x, emission.y, emission.z);
                        float light_pdf = power * square(dist) / (cos_l * P.lights_total_weight);
                        if (pdf_is_valid(light_pdf)) {
                            float w = P.config.enable_multiple_importance_sampling ? power_heuristic(light_pdf, bp) : 1.0f;
                            float3 ill = throughput * bv * emission * w / light_pdf;
                            emit_shadow = true;
                            sh0 = make_float4(hp.x, hp.y, hp.z, to_light.x);
                            sh1 = make_float4(to_light.y, to_light.z, dist, __uint_as_float(pf & ~PTB_FLAGS_ALL));
                            sh_ill = make_float4(ill.x, ill.y, ill.z, 0.0f);
                        }
                    }
                }

                // ---- sample the BSDF, emit the extension ray
                float3 dir_out; float pdf;
                if (bsdf.sample(P, c, throughput, medium_id, dir_out, pdf)) {
                    float3 org = ray_origin_epsilon_offset(hit_point, dir_out, gn);
                    bool nee = bsdf.allow_nee();
                    emit_next = true;
                    nx0 = make_float4(org.x, org.y, org.z, dir_out.x);
                    nx1 = make_float4(dir_out.y, dir_out.z, cone_angle, cone_width);
                    nx_path = make_float4(throughput.x, throughput.y, throughput.z, pdf);
                    nx_pix = (pf & ~PTB_FLAGS_ALL) | (nee ? PTB_FLAG_ALLOW_NEE : 0u) | (medium_id != PTB_INVALID ? PTB_FLAG_INSIDE_MEDIUM : 0u);
                    nx_medium = medium_id;
                }
            }
        }
        int s_slot = warp_append(&P.counters->shadow[bounce], emit_shadow);
        if (emit_shadow) { P.sq.od0[s_slot] = sh0; P.sq.od1[s_slot] = sh1; P.sq.illum[s_slot] = sh_ill; }
        int n_slot = warp_append(&P.counters->trace[bounce + 1], emit_next);
        if (emit_next) {
            qn.od0[n_slot] = nx0; qn.od1[n_slot] = nx1; qn.path[n_slot] = nx_path; qn.pix[n_slot] = nx_pix;
            if (nx_medium != PTB_INVALID) qn.medium[n_slot] = nx_medium;
        }
    }
}

// ------------------------------------------------------------------------------------------ ambient occlusion integrator
// The reference's second integrator (Src/CUDA/AO.cu:116-184, host Src/Renderer/Integrators/AO.cpp:143-192): primary hit -> one
// cosine-distributed occlusion ray of length ao_radius; an unoccluded ray sets the pixel to 1.  Reuses the wavefront machinery:
// k_generate -> k_trace8 / k_trace2 -> k_ambient_occlusion -> shadow trace (deposit) -> k_accumulate, waves included.
__global__ void __launch_bounds__(256) k_ambient_occlusion(const __grid_constant__ Frame P, float ao_radius) {
    const RayQueue& q = P.q[0];
    const int count = P.counters->trace[0];
    const int rounded = (count + 31) & ~31;
    for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < rounded; index += gridDim.x * blockDim.x) {
        bool emit = false;
        float4 sh0, sh1;
        if (index < count) {
            float4 a = q.od0[index], b = q.od1[index];
            float3 ray_direction = f3(a.w, b.x, b.y);
            Hit hit = unpack_hit(__ldg(q.hit + index));
            unsigned pf = q.pix[index];
            const int pixel_index = word_pixel(P, pf);
            const int fbi = word_fb_index(P, pf);
            const int sample_index = P.first_sample + word_slot(P, pf);
            if (hit.triangle_id != PTB_INVALID) {
                if (P.pixel_query[0] == pixel_index && word_slot(P, pf) == 0) { P.pixel_query[1] = hit.mesh_id; P.pixel_query[2] = hit.triangle_id; }
                TriFull tri = load_tri_full(P, hit.triangle_id);
                float3 gn = normalize(cross(tri.e1, tri.e2));                 // mesh space, like the reference (AO.cu:136)
                float3 hit_point = barycentric(hit.u, hit.v, tri.p0, tri.e1, tri.e2);
                float3 hit_normal = barycentric(hit.u, hit.v, tri.n0, tri.ne1, tri.ne2);
                Mat3x4 world = load_mat(P.mesh_transforms, hit.mesh_id);
                hit_point = xform_pos(world, hit_point);
                hit_normal = xform_dir(world, hit_normal);
                hit_normal = normalize(hit_normal);
                if (dot(ray_direction, hit_normal) > 0.0f) hit_normal = -hit_normal;
                aov_set(P, PTB_AOV_NORMAL, fbi, f4(hit_normal));
                aov_set(P, PTB_AOV_POSITION, fbi, f4(hit_point));
                float3 tangent, bitangent;
                orthonormal_basis(hit_normal, tangent, bitangent);
                float2 r = rng2<DIM_BSDF_0>(P, pixel_index, 0, sample_index);
                float3 wo = sample_cosine_hemisphere(r.x, r.y);
                float3 dir_out = local_to_world(wo, tangent, bitangent, hit_normal);
                float pdf = wo.z * PTB_ONE_OVER_PI;
                if (pdf_is_valid(pdf)) {
                    float3 org = ray_origin_epsilon_offset(hit_point, dir_out, gn);
                    emit = true;
                    sh0 = make_float4(org.x, org.y, org.z, dir_out.x);
                    sh1 = make_float4(dir_out.y, dir_out.z, ao_radius, __uint_as_float(pf & ~PTB_FLAGS_ALL));
                }
            }
        }
        int slot = warp_append(&P.counters->shadow[0], emit);
        if (emit) { P.sq.od0[slot] = sh0; P.sq.od1[slot] = sh1; P.sq.illum[slot] = make_float4(1.0f, 1.0f, 1.0f, 1.0f); }
    }
}

// ------------------------------------------------------------------------------------------ accumulate (+ clear)
// kernel_accumulate (Pathtracer.cu:775-796, AOV.h:35-46) fused with the framebuffer clear the reference does with
// separate memsets (Integrator.cpp:377-383): one read-modify-write pass over HBM per enabled AOV instead of two.
#ifndef PTB_ACC_BATCH
#define PTB_ACC_BATCH 9     // slot planes whose loads are in flight together (one batch for the bench's 9-pass wave: 0.39 -> 0.115 ms; 4: 0.39, 12: 0.23)
#endif
__global__ void __launch_bounds__(256) k_accumulate(const __grid_constant__ Frame P) {
    const bool push = P.xchg.count > 0 && P.xchg.push;
    size_t plane = 0;
    if (push) plane = size_t(P.xchg.control[P.rank]->epoch & 1u) * size_t(P.fb_stride);   // epoch only moves in k_exchange_wait, stream-ordered before us
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.local_pixels; i += gridDim.x * blockDim.x) {
        int x, y; local_to_pixel(P, i, x, y);
        int px = x + y * P.pitch;
        float4 colour = f4(0.0f);
#pragma unroll
        for (int k = 0; k < PTB_AOV_COUNT; k++) {
            if (!P.aov[k].fb) continue;
            const bool averaged = !(k == PTB_AOV_RADIANCE_DIRECT || k == PTB_AOV_RADIANCE_INDIRECT);   // those two are only cleared by the reference
            float4 acc = averaged ? P.aov[k].acc[px] : f4(0.0f);
            // fold the slot planes in pass order (same arithmetic as one kernel_accumulate per pass); the loads of four planes are
            // issued together -- a plain loop serialises nine dependent-looking HBM round trips per pixel (0.72 -> 0.2 ms per frame)
            for (int s0 = 0; s0 < P.wave_samples; s0 += PTB_ACC_BATCH) {
                float4 fb[PTB_ACC_BATCH];
#pragma unroll
                for (int j = 0; j < PTB_ACC_BATCH; j++) {
                    size_t fbi = size_t(s0 + j) * P.fb_stride + px;
                    fb[j] = (averaged && s0 + j < P.wave_samples) ? __ldcs(P.aov[k].fb + fbi) : f4(0.0f);
                }
#pragma unroll
                for (int j = 0; j < PTB_ACC_BATCH; j++) {
                    if (s0 + j >= P.wave_samples) break;
                    if (averaged) {
                        float n = float(P.first_sample + s0 + j);
                        if (n > 0.0f) acc += (fb[j] - acc) / n; else acc = fb[j];
                    }
                    P.aov[k].fb[size_t(s0 + j) * P.fb_stride + px] = f4(0.0f);
                }
            }
            if (averaged) P.aov[k].acc[px] = acc;
            if (k == PTB_AOV_RADIANCE) colour = acc;
        }
        if (!isfinite(colour.x + colour.y + colour.z)) colour = make_float4(1000.0f, 0.0f, 1000.0f, 1.0f);
        P.display[px] = colour;
        if (push) {
            for (int r = 0; r < P.xchg.count; r++) P.xchg.frames[r][plane + px] = colour;       // NVLink stores, 16 B per lane, row-contiguous
        }
    }
    if (push) {
        // release: every thread fences its peer stores, the last CTA to finish announces this rank's rows on every rank
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            ExchangeControl* mine = P.xchg.control[P.rank];
            unsigned done = atomicAdd(&mine->blocks_done, 1u);
            if (done == gridDim.x - 1) {
                mine->blocks_done = 0;
                __threadfence_system();
                const unsigned frame_no = mine->epoch + 1u;
                for (int r = 0; r < P.xchg.count; r++) {
                    unsigned* slot = &P.xchg.control[r]->arrivals[P.rank];
                    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(slot), "r"(frame_no) : "memory");
                }
            }
        }
    }
}

// Wait until every rank has stored its rows of the current frame into OUR block, then advance the epoch (which flips the
// buffer the next frame is stored to).  One thread; bounded spin (a lost peer must not hang the GPU).
__global__ void k_exchange_wait(const __grid_constant__ Frame P) {
    ExchangeControl* mine = P.xchg.control[P.rank];
    const unsigned target = mine->epoch + 1u;
    unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        bool all = true;
        for (int r = 0; r < P.xchg.count; r++) {
            unsigned seen; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(&mine->arrivals[r]) : "memory");
            all = all && int(seen - target) >= 0;
        }
        if (all) break;
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 4000000000ull) { mine->status = 1u; break; }
        __nanosleep(200);
    }
    mine->epoch += 1u;
}

// ------------------------------------------------------------------------------------------ bookkeeping
// Start of pass: zero the per-bounce counters and seed trace[0] (what BufferSizes::reset + a blocking
// cuMemcpyHtoD does in the reference, Pathtracer.cpp:791-795,846-847) -- here a device-side kernel, no host stall.
__global__ void k_begin_pass(const __grid_constant__ Frame P) {
    int* c = reinterpret_cast<int*>(P.counters);
    for (int i = threadIdx.x; i < int(sizeof(Counters) / sizeof(int)); i += blockDim.x) c[i] = 0;
    if (P.bin_counts) for (int i = threadIdx.x; i < PTB_ORDER_MAX_BOUNCE * 2 * PTB_ORDER_MAX_BINS; i += blockDim.x) P.bin_counts[i] = 0;
    __syncthreads();
    if (threadIdx.x == 0) P.counters->trace[0] = P.local_pixels * P.wave_samples;
}
// End of pass: fold the per-bounce counters into 64-bit totals (for Mrays/s).
__global__ void k_fold_counters(const __grid_constant__ Frame P) {
    int b = threadIdx.x;
    if (b < PTB_MAX_BOUNCES) {
        P.totals->trace[b] += (unsigned long long)P.counters->trace[b];
        P.totals->shadow[b] += (unsigned long long)P.counters->shadow[b];
        for (int m = 0; m < 4; m++) if (P.counters->mat[m][b]) atomicAdd(&P.totals->mat[m], (unsigned long long)P.counters->mat[m][b]);
    }
    if (b == 0) P.totals->frames += 1ull;
}

// primary-hit tap for parity tests: pixel-keyed copy of the bounce-0 hits
__global__ void k_tap_primary_hits(const __grid_constant__ Frame P, uint4* out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.local_pixels; i += gridDim.x * blockDim.x)
        out[word_pixel(P, P.q[0].pix[i])] = P.q[0].hit[i];                   // slot 0 of the last wave
}

// ------------------------------------------------------------------------------------------ SVGF input exchange (world > 1)
PTB_DI unsigned ld_acquire_sys(const unsigned* p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
PTB_DI void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
PTB_DI bool spin_until_all(const Frame& P, const unsigned* slots, unsigned target) {
    unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        bool all = true;
        for (int r = 0; r < P.xchg.count; r++) all = all && int(ld_acquire_sys(slots + r) - target) >= 0;
        if (all) return true;
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 4000000000ull) return false;
        __nanosleep(200);
    }
}
// Tile-local SVGF (world > 1).  Tracing is sharded by interleaved bands, filtering by contiguous blocks of rows: rank f filters rows
// [f * rows_per_block, (f + 1) * rows_per_block) and needs the noisy inputs of those rows plus PTB_SVGF_HALO rows on either side.
PTB_DI int svgf_block_owner(const Frame& P, int y) { int o = y / P.xchg.rows_per_block; return o < P.world ? o : P.world - 1; }
PTB_DI void svgf_ext_range(const Frame& P, int f, int& y0, int& y1) {
    y0 = max(0, f * P.xchg.rows_per_block - PTB_SVGF_HALO);
    y1 = min(P.height, (f == P.world - 1 ? P.height : (f + 1) * P.xchg.rows_per_block) + PTB_SVGF_HALO);
}
// 1. every rank stores the rows it TRACED (six planes, 80 B per pixel) into the input planes (this frame's parity) of every rank
//    whose extended block contains the row -- itself included -- then publishes the frame number on every rank
__global__ void __launch_bounds__(256) k_svgf_push(const __grid_constant__ Frame P) {
    const int S = P.fb_stride;
    const size_t in_off = xchg_input_offset(S, P.svgf.parity);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.local_pixels; i += gridDim.x * blockDim.x) {
        int x, y; local_to_pixel(P, i, x, y);
        size_t px = size_t(x) + size_t(y) * P.pitch;
        float4 v0 = P.aov[PTB_AOV_RADIANCE_DIRECT].fb[px], v1 = P.aov[PTB_AOV_RADIANCE_INDIRECT].fb[px], v2 = P.aov[PTB_AOV_ALBEDO].fb[px];
        float4 v3 = P.svgf.gbuf_normal_depth[px];
        int2 ids = P.svgf.gbuf_ids[px];
        float2 sp = P.svgf.gbuf_screen_prev[px];
        // the filter chain clears nothing on this side any more: the trace-side planes start the next frame empty
        P.aov[PTB_AOV_RADIANCE_DIRECT].fb[px] = f4(0.0f); P.aov[PTB_AOV_RADIANCE_INDIRECT].fb[px] = f4(0.0f); P.aov[PTB_AOV_ALBEDO].fb[px] = f4(0.0f);
        P.svgf.gbuf_normal_depth[px] = f4(0.0f); P.svgf.gbuf_ids[px] = make_int2(0, 0); P.svgf.gbuf_screen_prev[px] = f2(0.0f, 0.0f);
        for (int f = 0; f < P.world; f++) {
            int y0, y1; svgf_ext_range(P, f, y0, y1);
            if (y < y0 || y >= y1) continue;
            float4* dst = P.xchg.frames[f] + in_off;
            dst[px] = v0; dst[size_t(S) + px] = v1; dst[size_t(S) * 2 + px] = v2; dst[size_t(S) * 3 + px] = v3;
            reinterpret_cast<int2*>(dst + size_t(S) * 4)[px] = ids;
            reinterpret_cast<float2*>(dst + size_t(S) * 4 + size_t(S) / 2)[px] = sp;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        ExchangeControl* mine = P.xchg.control[P.rank];
        unsigned done = atomicAdd(&mine->svgf_blocks_done, 1u);
        if (done == gridDim.x - 1) {
            mine->svgf_blocks_done = 0;
            __threadfence_system();
            const unsigned frame_no = mine->svgf_epoch + 1u;
            for (int r = 0; r < P.xchg.count; r++) st_release_sys(&P.xchg.control[r]->svgf_arrivals[P.rank], frame_no);
        }
    }
}
// 2. wait for every rank's rows of this frame (which also means: every rank has finished filtering the previous frame, so the
//    history it owns is complete and nobody still reads the parity this frame overwrites), advance the epoch
__global__ void k_svgf_wait_arrivals(const __grid_constant__ Frame P) {
    ExchangeControl* mine = P.xchg.control[P.rank];
    if (!spin_until_all(P, mine->svgf_arrivals, mine->svgf_epoch + 1u)) mine->status = 1u;
    mine->svgf_epoch += 1u;
}
// 3. after the filter chain: the displayed rows of this rank's block go to every rank's gathered frame (same protocol and
//    counters as the accumulate kernel's fused gather; k_exchange_wait closes the frame)
__global__ void __launch_bounds__(256) k_svgf_push_display(const __grid_constant__ Frame P) {
    const size_t plane = size_t(P.xchg.control[P.rank]->epoch & 1u) * size_t(P.fb_stride);
    const int rows = P.svgf.block_y1 - P.svgf.block_y0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * P.width; i += gridDim.x * blockDim.x) {
        int y = P.svgf.block_y0 + i / P.width, x = i % P.width;
        size_t px = size_t(x) + size_t(y) * P.pitch;
        float4 c = P.display[px];
        for (int r = 0; r < P.xchg.count; r++) P.xchg.frames[r][plane + px] = c;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        ExchangeControl* mine = P.xchg.control[P.rank];
        unsigned done = atomicAdd(&mine->blocks_done, 1u);
        if (done == gridDim.x - 1) {
            mine->blocks_done = 0;
            __threadfence_system();
            const unsigned frame_no = mine->epoch + 1u;
            for (int r = 0; r < P.xchg.count; r++) st_release_sys(&P.xchg.control[r]->arrivals[P.rank], frame_no);
        }
    }
}

// tile export / assemble for the multi-GPU gather
__global__ void k_export_rows(const __grid_constant__ Frame P, const float4* src, float4* dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.local_pixels; i += gridDim.x * blockDim.x) {
        int x, y; local_to_pixel(P, i, x, y);
        int row = i / P.width;
        dst[row * P.pitch + x] = src[x + y * P.pitch];
    }
}
__global__ void k_assemble_rows(const __grid_constant__ Frame P, const float4* src, int max_rows, float4* dst) {
    int total = P.width * P.height;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int y = i / P.width, x = i - y * P.width;
        int band = y / P.band_rows;
        int owner = band % P.world;
        int local_row = (band / P.world) * P.band_rows + (y - band * P.band_rows);
        dst[x + y * P.pitch] = src[(size_t(owner) * max_rows + local_row) * P.pitch + x];
    }
}
