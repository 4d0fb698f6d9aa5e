/* ptb.h -- C ABI of the B200 wavefront path tracer (libptb.so).
 *
 * Drop-in boundary for ONE path of jan-van-bergen/GPU-Raytracer: the per-frame wavefront pipeline behind
 * `Integrator::update()` / `Pathtracer::render()`.  The reference has no FFI; its narrowest seam is the abstract
 * Integrator used by main (Src/Renderer/Integrators/Integrator.h:56-296, Src/Main.cpp:137-138) plus the
 * name-based host->module ABI (cuModuleGetGlobal names set in Integrator.cpp:15-30,97,154,171-175,276,292-303 and
 * Pathtracer.cpp:16-38,262-273,326-356,479-498).  Each entry point below names the reference interface it replaces.
 *
 * Conventions: plain C types only; every call returns 0 on success or a non-zero error (cudaError_t value, or a
 * negative PTB_E_* code) instead of the reference's print-and-__debugbreak (Src/Device/CUDACall.h:9-22).
 * The caller owns host memory (copied during the call); the ctx owns all device memory.  One ctx per GPU; calls on
 * one ctx are not thread-safe; different ctxs are independent.  No CPU fallback exists: every entry point fails
 * with an error if no CUDA device is usable.
 */
#ifndef PTB_H
#define PTB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ptb_ctx ptb_ctx;

#define PTB_E_BADARG  (-1)
#define PTB_E_NOSCENE (-2)
#define PTB_E_STATE   (-3)
#define PTB_E_EXCHANGE (-4)

/* AOVType, Src/CUDA/Common.h:27-37 */
enum { PTB_AOV_RADIANCE = 0, PTB_AOV_RADIANCE_DIRECT, PTB_AOV_RADIANCE_INDIRECT, PTB_AOV_ALBEDO, PTB_AOV_NORMAL, PTB_AOV_POSITION, PTB_AOV_COUNT };

/* Byte-identical to GPUConfig (44 bytes), Src/CUDA/Common.h:39-67.  Replaces the `config` module global. */
typedef struct ptb_config {
    int32_t  reconstruction_filter;      /* 0 box, 1 tent, 2 gaussian */
    uint32_t aov_mask;                   /* bit i enables AOV i (RADIANCE is always on) */
    int32_t  num_bounces;
    uint8_t  enable_mipmapping;
    uint8_t  enable_next_event_estimation;
    uint8_t  enable_multiple_importance_sampling;
    uint8_t  enable_russian_roulette;
    uint8_t  enable_svgf;
    uint8_t  enable_spatial_variance;
    uint8_t  enable_taa;
    uint8_t  pad_;
    float    alpha_colour;
    float    alpha_moment;
    int32_t  num_atrous_iterations;
    float    sigma_z, sigma_n, sigma_l;
} ptb_config;

/* Byte-identical to CUDACamera (60 bytes), Src/Renderer/Integrators/Integrator.cpp:456-472.  Replaces `camera`. */
typedef struct ptb_camera {
    float position[3];
    float bottom_left_corner[3];
    float x_axis[3];
    float y_axis[3];
    float pixel_spread_angle;
    float aperture_radius;
    float focal_distance;
} ptb_camera;

/* One texture: mip chain of RGBA8 texels or BC1 blocks (Src/Renderer/Texture.h, Integrator.cpp:42-94). */
typedef struct ptb_texture {
    int32_t             format;          /* 0 = RGBA8, 1 = BC1 */
    int32_t             width, height;   /* in texels */
    int32_t             num_levels;
    const void* const*  levels;          /* num_levels pointers, level l is max(w>>l,1) x max(h>>l,1) texels (BC1: 4x4 blocks of 8 bytes) */
    float               lod_bias;
} ptb_texture;

/* The flat scene: what Integrator::init_geometry / init_materials / init_sky / init_rng and
 * Pathtracer::calc_light_power upload (Integrator.cpp:21-99,101-283,285-304; Pathtracer.cpp:384-534). */
typedef struct ptb_scene {
    const void*    triangles;            /* triangle_count x 96 B, BVH leaf order (Integrator.h:139-151) */
    int32_t        triangle_count;
    const void*    bvh_nodes;            /* node_count x 80 B (CWBVH) or 32 B (binary); TLAS in [0, 2*mesh_count) */
    int32_t        bvh_node_count;
    int32_t        bvh_kind;             /* 8 (CWBVH, 80-byte nodes), 4 (BVH4, 128-byte nodes; BLAS entry = root + 1) or 2 (32-byte nodes) */
    int32_t        tlas_node_count;
    int32_t        mesh_count;           /* instances, in TLAS leaf order */
    const int32_t* mesh_bvh_root_indices;/* root | identity << 31 */
    const int32_t* mesh_material_ids;
    const float*   mesh_transforms;      /* mesh_count x 12 */
    const float*   mesh_transforms_inv;
    const float*   mesh_transforms_prev;
    int32_t        material_count;
    const int8_t*  material_types;
    const void*    materials;            /* material_count x 32 B */
    int32_t        medium_count;
    const void*    media;                /* medium_count x 32 B */
    int32_t        texture_count;
    const ptb_texture* textures;
    const float*   sky;                  /* sky_height x sky_width x float4 */
    int32_t        sky_width, sky_height;
    float          sky_scale;
    const float*   pmj_samples;          /* 64 x 4096 x float2 */
    const uint8_t* blue_noise;           /* 16 x 128 x 128 x uchar2 */
    float          lights_total_weight;
    int32_t        light_triangle_count;
    const int32_t* light_triangle_indices;
    const float*   light_triangle_cumulative_probability;
    int32_t        light_mesh_count;
    const float*   light_mesh_cumulative_probability;
    const int32_t* light_mesh_triangle_span;     /* pairs (first, last) */
    const int32_t* light_mesh_transform_indices;
} ptb_scene;

/* Rays traced since the last reset: what the reference only exposes as device `buffer_sizes` (Pathtracer.cu:103-116). */
typedef struct ptb_ray_stats {
    uint64_t trace[128];                 /* closest-hit rays per bounce */
    uint64_t shadow[128];                /* shadow rays per bounce */
    uint64_t shaded[4];                  /* diffuse, plastic, dielectric, conductor */
    uint64_t frames;
} ptb_ray_stats;

/* Traversal work per ray kind ([0] closest hit, [1] shadow) for ONE pass, counted by an instrumented variant of the
 * trace kernel: the inputs of SURVEY.md section 8d's algorithmic-bytes formula. */
typedef struct ptb_traversal_stats {
    uint64_t rays[2], nodes[2], triangles[2], instance_transforms[2], shadow_misses;
} ptb_traversal_stats;

/* Pathtracer(gl_tex, width, height, scene) + cuda_init/resize_init (Pathtracer.h:260, Pathtracer.cpp:9-41,255-301).
 * rank/world: this ctx traces the rows ((y / band_rows) % world == rank); world = 1 renders the whole frame. */
int  ptb_create(ptb_ctx** out, int device, int width, int height, int rank, int world, int band_rows);
/* cuda_free (Pathtracer.cpp:43-74) */
void ptb_destroy(ptb_ctx* ctx);
/* init_materials / init_geometry / init_sky / init_rng / light tables (see ptb_scene) */
int  ptb_upload_scene(ptb_ctx* ctx, const ptb_scene* scene);
/* global_config.set_value (Integrator.cpp:518-521); resets accumulation like invalidated_gpu_config does */
int  ptb_set_config(ptb_ctx* ctx, const ptb_config* config);
/* global_camera.set_value (Integrator.cpp:454-481); view_projection[_prev] feed svgf_data (Pathtracer.cpp:707-717), may be NULL */
int  ptb_set_camera(ptb_ctx* ctx, const ptb_camera* camera, const float* view_projection, const float* view_projection_prev);
/* build_tlas upload (Integrator.cpp:399-430): new TLAS nodes + per-instance tables after the host rebuilt the TLAS */
int  ptb_update_instances(ptb_ctx* ctx, const void* tlas_nodes, int tlas_node_count, int mesh_count,
                          const int32_t* mesh_bvh_root_indices, const int32_t* mesh_material_ids,
                          const float* transforms, const float* transforms_inv, const float* transforms_prev);
/* TLAS refit on the GPU (replaces the per-frame CPU rebuild of Integrator::build_tlas, Integrator.cpp:399-430,
 * for instances that only move): new 3x4 transforms (mesh_count x 12 floats each, in the order of the last uploaded tables), same TLAS
 * topology.  The device recomputes every instance box from its BLAS root box and re-quantises the TLAS nodes bottom-up in place
 * (csrc/ptb_refit.cuh); no host BVH work, no host <-> device synchronisation.  transforms_prev may be NULL: the transforms in use become
 * the previous ones (Mesh::transform_prev).  An instance whose transform stops being the identity leaves the merged static BVH (its
 * slot is retired) and is traced through the TLAS from then on.  Closest hits equal those of a TLAS rebuilt on the host. */
int  ptb_refit_instances(ptb_ctx* ctx, const float* transforms, const float* transforms_inv, const float* transforms_prev);
/* Pathtracer::render() for one pass with the given sample_index (Pathtracer.cpp:738-855). Asynchronous on the ctx stream. */
int  ptb_render(ptb_ctx* ctx, int sample_index);
/* Replaces the reference's batching (BATCH_SIZE = 1080 x 720 pixels, Src/CUDA/Common.h:69-71; a blocking 4-KB upload between batches, Pathtracer.cpp:789-795):
 * number of pass slots a wave can carry (default 1).  With `samples` > 1, ptb_render_frame traces up to that many consecutive
 * passes TOGETHER (every ray carries its pass slot; each slot has its own framebuffer plane; the accumulate pass folds the
 * planes in pass order, so accumulators are bit-identical to tracing pass by pass).  Re-allocates the ray queues. */
int  ptb_reserve_wave(ptb_ctx* ctx, int samples);
/* The reference's `-N <samples>` capture loop (Src/Main.cpp:137-142: update + render until sample_index == N):
 * `num_passes` consecutive ptb_render calls (sample_index = first_sample_index ...) replayed as ONE CUDA graph: the launch
 * sequence of a frame is static (queue sizes live in device memory), so the ~20 launches per pass cost one graph launch per
 * frame.  Graphs are cached per (first_sample_index, num_passes) and dropped whenever camera / config / instances change. */
int  ptb_render_frame(ptb_ctx* ctx, int first_sample_index, int num_passes);
/* Same as ptb_render, with the instrumented trace kernels (the reference has no counterpart; feeds SURVEY 8d's algorithmic bytes); blocks and returns the pass's traversal statistics. */
int  ptb_measure_traversal(ptb_ctx* ctx, int sample_index, ptb_traversal_stats* out);
/* cuStreamSynchronize equivalent (the reference's blocking set_value at the end of Pathtracer::render, Pathtracer.cpp:845-847, plays this role) */
int  ptb_sync(ptb_ctx* ctx);
/* get_aov(type).framebuffer / .accumulator (Integrator.h:247): device pointer to pitch x height float4 */
int  ptb_get_aov(ptb_ctx* ctx, int aov_type, int accumulated, void** device_ptr, int* pitch);
/* Device pointer of the displayed image (what kernel_accumulate / svgf_finalize / taa_finalize write to the GL surface) */
int  ptb_get_display(ptb_ctx* ctx, void** device_ptr, int* pitch);
/* Copies an AOV (or the display image when aov_type < 0) to host memory: pitch x height x float4 */
int  ptb_download(ptb_ctx* ctx, int aov_type, int accumulated, float* host_dst);
/* The window's post-processing pass (Src/Shaders/post.frag:18-41: clamp >= 0, ACES filmic curve, gamma 2.2) applied on the device to the
 * displayed frame (with several ranks: the gathered frame), 8-bit RGBA, pitch x height, row 0 = bottom.  device_rgba8 (may be NULL)
 * receives the device pointer; host_dst (may be NULL) a blocking copy.  What the reference captures as .ppm (Src/Main.cpp:195-225). */
int  ptb_present(ptb_ctx* ctx, void** device_rgba8, void* host_dst);
/* set_pixel_query(x, y) (Integrator.h:266-277): the next rendered pass records which (mesh_id, triangle_id) the primary ray of that
 * pixel hits (kernel_sort, Pathtracer.cu:345-348).  ptb_get_pixel_query blocks, returns them (-1, -1 = sky or nothing rendered yet)
 * and clears the query like Integrator::update does (Integrator.cpp:483-494); mesh_id is the instance index in TLAS leaf order. */
int  ptb_set_pixel_query(ptb_ctx* ctx, int x, int y);
int  ptb_get_pixel_query(ptb_ctx* ctx, int* mesh_id, int* triangle_id);
/* Ray counters (device buffer_sizes, summed over passes) */
int  ptb_get_ray_stats(ptb_ctx* ctx, ptb_ray_stats* out, int reset);
/* CUDA stream the ctx launches on (cudaStream_t as void*), e.g. to enqueue a collective after ptb_render */
int  ptb_get_stream(ptb_ctx* ctx, void** stream);
/* Rows owned by this rank, packed: writes owned_rows x pitch float4 of the accumulated AOV into dst (device pointer) */
int  ptb_export_rows(ptb_ctx* ctx, int aov_type, void* device_dst, int* owned_rows);
/* Inverse of ptb_export_rows for `world` packed tiles laid out rank-major in device_src (each max_rows x pitch float4):
 * scatters them into the full-frame image device_dst (pitch x height float4).  Used after the NCCL all-gather. */
int  ptb_assemble_rows(ptb_ctx* ctx, const void* device_src, int max_rows, void* device_dst);
/* Static merge (default on): instances with an identity transform (root bit 31) are additionally built into ONE CWBVH at
 * upload (CPU SAH build inside ptb_upload_scene / ptb_update_instances) and traced through it instead of TLAS -> BLAS
 * (BVH8.h:204-232); hits still report the original (mesh_id, triangle_id).  mode 0 = trace the reference's two-level hierarchy only
 * (every pixel bit-identical to the reference kernels); 1 = merged tree built with spatial splits (the algorithm of the reference's
 * SBVH builder, Src/BVH/Builders/SBVHBuilder.cpp; default); 2 = merged tree built with the plain full-sweep SAH builder. */
int  ptb_set_static_merge(ptb_ctx* ctx, int mode);
/* Ray-triangle test used inside the merged static BVH.  PTB_INTERSECT_MT (default) = Moeller-Trumbore, the reference's test
 * (Src/CUDA/Raytracing/Triangle.h:148-198), bit-exact.  PTB_INTERSECT_WOOP = Woop's precomputed unit-triangle map (one 48-byte affine
 * map per triangle reference, built at upload): u, v, t differ from the reference in the last ulps, so frames agree within the
 * 1e-4 rel-L2 bar instead of bit for bit (tests/test_gpu_configs.py reports the hit-id mismatch rate).  Instances outside the
 * merged tree (moving / transformed ones) keep Moeller-Trumbore. */
/* Pathtracer::resize_free + resize_init (Src/Renderer/Integrators/Pathtracer.cpp:255-314): new film size for the same scene.
 * Accumulators, SVGF / TAA history and cached frame graphs start over; set the camera for the new film afterwards (Camera::resize)
 * and restart sample_index at 0.  Not allowed while a frame exchange is connected (PTB_E_STATE). */
int  ptb_resize(ptb_ctx* ctx, int width, int height);
/* Which integrator ptb_render / ptb_render_frame run.  PTB_INTEGRATOR_PATHTRACER (default): Pathtracer::render.  PTB_INTEGRATOR_AO: the
 * reference's ambient-occlusion integrator (Src/CUDA/AO.cu:49-184, Src/Renderer/Integrators/AO.cpp:143-192): primary hit, one
 * cosine-distributed occlusion ray of length ao_radius (AO.h:94: 1.0), RADIANCE = 1 where it escapes; NORMAL / POSITION AOVs as there. */
#define PTB_INTEGRATOR_PATHTRACER 0
#define PTB_INTEGRATOR_AO         1
int  ptb_set_integrator(ptb_ctx* ctx, int kind, float ao_radius);
#define PTB_INTERSECT_MT   0
#define PTB_INTERSECT_WOOP 1
int  ptb_set_intersector(ptb_ctx* ctx, int kind);
/* Trace order of secondary and shadow rays: 0 = queue (emission) order like the reference's kernel_trace_* (Pathtracer.cu:165-197),
 * 8 / 64 = counting-sorted by direction bin (octants / 8x8 octahedral cells) before each trace launch.  A scheduling choice
 * only: every pixel gets the same rays and the same result, bit for bit. */
int  ptb_set_ray_ordering(ptb_ctx* ctx, int bins);
/* Frame exchange over NVLink peer memory -- the multi-GPU gather fused into the accumulate kernel (north_star: "a final gather
 * of the tile framebuffers").  The reference is single-GPU (Pathtracer.cpp:738-855 ends in kernel_accumulate writing the one
 * GL surface, Pathtracer.cu:775-796); here the last accumulate of ptb_render_frame stores every finished pixel straight into
 * the full-frame buffer of EVERY rank and the frame ends with a device-side wait for all peers, so after ptb_render_frame each
 * rank holds the complete image -- no export / all_gather / assemble launches.
 *   ptb_exchange_create       allocate this rank's block; returns its device base and/or its 64-byte CUDA IPC handle
 *   ptb_exchange_connect      peers living in THIS process: world base pointers, rank-major (own entry ignored)
 *   ptb_exchange_connect_ipc  peers in other processes: world x 64-byte IPC handles, rank-major (own entry ignored)
 *   ptb_exchange_disconnect   stop exchanging
 *   ptb_exchange_frame        device pointer to the complete frame (pitch x height float4) of the last ptb_render_frame; valid
 *                             until the next-but-one ptb_render_frame (two buffers alternate); read it on the ctx stream or
 *                             order the read before the next ptb_render_frame.
 * With SVGF enabled and world > 1 the exchange carries the filter's noisy inputs instead (every ptb_render stores this rank's rows
 * of them into every peer, then every rank filters the whole frame): ptb_exchange_frame then returns the display buffer.
 * Every rank must call ptb_render_frame / ptb_render the same number of times.  A peer that does not deliver within 4 s makes the next
 * ptb_sync return PTB_E_EXCHANGE instead of hanging the GPU. */
int  ptb_exchange_create(ptb_ctx* ctx, void** local_base, void* ipc_handle_out);
int  ptb_exchange_connect(ptb_ctx* ctx, void* const* peer_bases);
int  ptb_exchange_connect_ipc(ptb_ctx* ctx, const void* ipc_handles);
int  ptb_exchange_frame(ptb_ctx* ctx, void** device_ptr, int* pitch);
int  ptb_exchange_disconnect(ptb_ctx* ctx);   /* back to rank-local frames (the block stays allocated) */
/* Debug/parity taps: copy queue state of the LAST rendered pass to host.  which: 0 = primary hits (pitch*height uint4, pixel keyed) */
int  ptb_debug_read(ptb_ctx* ctx, int which, void* host_dst, int64_t bytes);
/* Number of kernels this library launched since creation (bench.py's gpu_launches) */
int64_t ptb_launch_count(ptb_ctx* ctx);
/* Per-stage device time of the last ptb_render with timing enabled: fills up to n floats (ms) in the order of ptb_stage_name() */
int  ptb_set_timing(ptb_ctx* ctx, int enabled);
int  ptb_get_stage_ms(ptb_ctx* ctx, float* ms, int n);
const char* ptb_stage_name(int i);
const char* ptb_error_string(int code);

#ifdef __cplusplus
}
#endif
#endif /* PTB_H */
