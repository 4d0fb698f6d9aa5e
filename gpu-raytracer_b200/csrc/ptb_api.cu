// libptb.so -- host side of the B200 wavefront path tracer: device memory, uploads, the per-pass launch
// sequence (the role of Pathtracer::render(), Src/Renderer/Integrators/Pathtracer.cpp:738-855) and the C ABI
// declared in include/ptb.h.  CUDA runtime API + a little driver API for mip-mapped block-compressed textures.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include <algorithm>

#include "ptb.h"
#include "ptb_kernels.cuh"
#include "ptb_post.cuh"
#include "../host/static_merge.h"
#include "ptb_svgf.cuh"
#include "ptb_refit.cuh"
#include <unordered_map>
#include <array>

#define CK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { ctx_fail(ctx, #expr, (int)e__); return (int)e__; } } while (0)
// inside ptb_create: a failure must not leak the half-built context
#define CKC(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { ctx_fail(ctx, #expr, (int)e__); ptb_destroy(ctx); return (int)e__; } } while (0)
#define CKD(expr) do { CUresult r__ = (expr); if (r__ != CUDA_SUCCESS) { ctx_fail(ctx, #expr, 1000 + (int)r__); return 1000 + (int)r__; } } while (0)

// Driver-API entry points are resolved through the runtime (cudaGetDriverEntryPoint) so libptb.so carries no link-time
// dependency on libcuda.so.1: the library still loads (and its symbol table can be checked) on a machine without a driver.
struct DriverApi {
    CUresult (*MipmappedArrayCreate)(CUmipmappedArray*, const CUDA_ARRAY3D_DESCRIPTOR*, unsigned) = nullptr;
    CUresult (*MipmappedArrayGetLevel)(CUarray*, CUmipmappedArray, unsigned) = nullptr;
    CUresult (*MipmappedArrayDestroy)(CUmipmappedArray) = nullptr;
    CUresult (*Memcpy2D)(const CUDA_MEMCPY2D*) = nullptr;
    CUresult (*TexObjectCreate)(CUtexObject*, const CUDA_RESOURCE_DESC*, const CUDA_TEXTURE_DESC*, const CUDA_RESOURCE_VIEW_DESC*) = nullptr;
    CUresult (*TexObjectDestroy)(CUtexObject) = nullptr;
    CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
    bool ok = false;
};
static DriverApi g_drv;
// 2-D tensor map over one pitch x height float4 plane, seen as float32 [height][pitch * 4]; box = box_w texels x box_h rows;
// out-of-range texels are filled with zeros by the copy engine
static int encode_tile_map_impl(CUtensorMap* map, const void* plane, int pitch, int height, int box_w, int box_h) {
    if (!g_drv.TensorMapEncodeTiled || (reinterpret_cast<size_t>(plane) & 15) || box_w * 4 > 256 || box_h > 256) return 1;
    cuuint64_t dims[2] = { (cuuint64_t)pitch * 4, (cuuint64_t)height };
    cuuint64_t strides[1] = { (cuuint64_t)pitch * 16 };
    cuuint32_t box[2] = { (cuuint32_t)box_w * 4, (cuuint32_t)box_h };
    cuuint32_t elem[2] = { 1, 1 };
    CUresult r = g_drv.TensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(plane), dims, strides, box, elem,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
}
static int load_driver_api() {
    if (g_drv.ok) return 0;
    struct { const char* name; void** slot; } want[] = {
        { "cuMipmappedArrayCreate", (void**)&g_drv.MipmappedArrayCreate }, { "cuMipmappedArrayGetLevel", (void**)&g_drv.MipmappedArrayGetLevel },
        { "cuMipmappedArrayDestroy", (void**)&g_drv.MipmappedArrayDestroy }, { "cuMemcpy2D", (void**)&g_drv.Memcpy2D },
        { "cuTexObjectCreate", (void**)&g_drv.TexObjectCreate }, { "cuTexObjectDestroy", (void**)&g_drv.TexObjectDestroy } };
    for (auto& w : want) {
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint(w.name, w.slot, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*w.slot) { fprintf(stderr, "[ptb] driver entry point %s unavailable\n", w.name); return PTB_E_STATE; }
    }
    g_drv.ok = true;
    {   // optional: tensor-map TMA for the SVGF tiles (the filter falls back to plain staging loads without it)
        cudaDriverEntryPointQueryResult q; void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && fn &&
            !(getenv("PTB_SVGF_TMA") && atoi(getenv("PTB_SVGF_TMA")) == 0)) {
            g_drv.TensorMapEncodeTiled = reinterpret_cast<decltype(g_drv.TensorMapEncodeTiled)>(fn);
            encode_tile_map = encode_tile_map_impl;
        }
        cudaGetLastError();
    }
    return 0;
}

enum Stage { ST_GENERATE = 0, ST_TRACE, ST_SORT, ST_SHADE, ST_SHADOW, ST_POST, ST_ORDER, ST_COUNT };
static const char* kStageNames[ST_COUNT] = { "generate", "trace", "sort", "shade", "shadow_trace", "accumulate_or_svgf", "ray_order" };

struct DeviceTexture {
    CUmipmappedArray array = nullptr;
    CUtexObject      tex = 0;
};

struct ptb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    Frame F;                         // host copy of the kernel parameter block
    bool has_scene = false;
    bool has_type[4] = { false, false, false, false };
    bool has_lights = false;
    int  mesh_capacity = 0, node_count = 0, bvh_kind = 8;
    int  sm_count = 148;
    int  owned_rows = 0;
    std::vector<void*> allocs;       // everything cudaMalloc'ed, freed in ptb_destroy
    std::vector<void*> film_allocs;  // everything sized by the film (display, accumulators, SVGF state): re-allocated by ptb_resize
    std::vector<void*> wave_allocs;  // ray queues + framebuffer planes, re-allocated by ptb_reserve_wave
    int wave_capacity = 1;           // pass slots a wave can carry
    std::vector<DeviceTexture> textures;
    cudaArray_t sky_array = nullptr;
    cudaArray_t lut_arrays[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    bool luts_ready = false;
    uint4* tap_hits = nullptr;
    uchar4* present_buf = nullptr;                    // tone-mapped 8-bit frame (ptb_present), allocated on first use
    long long launches = 0;
    bool stats_mode = false;
    struct FrameGraph { int first, passes; cudaGraphExec_t exec; long long launches; };
    std::vector<FrameGraph> graphs;
    bool capturing = false;
    bool timing = false;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> timed;
    std::vector<cudaEvent_t> event_pool;
    size_t event_used = 0;
    float stage_ms[ST_COUNT] = {};
    int last_sample_index = -1;
    int frames_since_reset = 0;
    std::string last_error;
    SVGFHistory local_hist[2] = {};                   // SVGF / TAA temporal state of a single-GPU context (two parities)
    unsigned svgf_frames = 0;                         // filtered frames so far: frame k writes parity k & 1
    int pixel_query_host[3] = { -1, -1, -1 };
    // shadow rays of bounce b and extension rays of bounce b+1 are independent until the next sort: the shadow trace runs on a
    // side stream so that the tail of one persistent trace kernel is filled by the CTAs of the other (render_wave)
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool overlap_enabled = true;
    // static merge: identity-transform instances re-built into ONE CWBVH at upload (rebuild_static_merge)
    bool merge_enabled = true;
    bool merge_spatial = true;
    bool half_nodes = true;                           // conservative packed-half node test outside the bit-exact mode (A/B: PTB_HALF_NODES=0)                        // merged BVH built with spatial splits (SBVH); false = plain full-sweep SAH
    std::vector<unsigned char> host_nodes;            // node array as uploaded by the host (for leaf walks)
    std::vector<float4> host_tri_pos;                 // first 3 float4 of every triangle record
    std::vector<int> host_roots;                      // roots as last given by the host
    std::unordered_map<unsigned, std::array<float, 6>> blas_root_boxes;    // BLAS root node -> its box in mesh space (ptb_refit_instances)
    float* refit_scratch = nullptr;                   // device: local boxes | instance boxes | node boxes | done flags
    std::vector<int> merge_slot_root;                 // merged slot -> BLAS root (identity bit included)
    std::vector<int> merge_slot_instance;             // merged slot -> instance index in the current TLAS leaf order
    std::vector<int> merge_decided_roots;             // sorted roots of the identity instances the last merge decision (built, disabled or skipped) was taken on
    float4* merge_nodes = nullptr;                    // device: [host node array | merged nodes], owned
    float4* merge_tris = nullptr;
    int     merge_ref_count = 0;                     // triangle references in the merged tree (>= triangles: spatial splits)
    float4* merge_woop = nullptr;                     // Woop maps of the merged references (ptb_set_intersector)
    int2*   merge_who = nullptr;
    int     intersector = PTB_INTERSECT_MT;
    int     integrator = PTB_INTEGRATOR_PATHTRACER;   // or PTB_INTEGRATOR_AO (ptb_set_integrator)
    float   ao_radius = 1.0f;                         // AO.h:94
    int*    merge_slot_instance_dev = nullptr;
    const float4* uploaded_nodes = nullptr;           // device copy of the host's array (kept: the fallback when nothing is merged)
    // pinned staging for the per-frame uploads (ptb_update_instances): the host arrays are copied here during the call and the
    // H2D copies run asynchronously behind the previous frame -- the call never waits for the GPU (two arenas alternate)
    unsigned char* stage_mem[2] = { nullptr, nullptr };
    unsigned char* stage_dev[2] = { nullptr, nullptr };
    int stage_segments = 0;
    size_t stage_cap[2] = { 0, 0 };
    cudaEvent_t stage_done[2] = { nullptr, nullptr };
    int stage_cur = 0; size_t stage_used = 0; bool stage_open = false;
    // frame exchange over peer memory
    void* xchg_block = nullptr;                       // {ExchangeControl (256 B), frame[2]} owned by this ctx
    void* xchg_ipc_opened[PTB_MAX_PEERS] = {};        // mappings opened with cudaIpcOpenMemHandle, closed in ptb_destroy
    unsigned xchg_frames = 0;                         // frames pushed + awaited so far (host mirror of ExchangeControl::epoch)
};

// CPU SAH + CWBVH builder (host/bvh_build.cpp, linked into this library): used for the merged static BVH
extern "C" {
void* ptbh_build_triangles(const float* pos, int n, int kind, float sah_node, float sah_leaf);
void* ptbh_build_triangles_sbvh(const float* pos, int n, float alpha, int bins, float max_dup);
void  ptbh_set_tri_cost(float c);
void  ptbh_set_optimizer(int passes, float fraction, int max_depth);
#ifndef PTB_MERGE_ALPHA_DEFAULT
#define PTB_MERGE_ALPHA_DEFAULT 3e-4f
#endif
#ifndef PTB_MERGE_BINS_DEFAULT
#define PTB_MERGE_BINS_DEFAULT 96
#endif
#ifndef PTB_MERGE_TRICOST_DEFAULT
#define PTB_MERGE_TRICOST_DEFAULT 1.0f
#endif
#ifndef PTB_MERGE_OPTIMIZE_DEFAULT
#define PTB_MERGE_OPTIMIZE_DEFAULT 1        // sweeps of the insertion-based optimiser over the merged tree's binary BVH (0 = off)
#endif
int   ptbh_node_count(void* h);
int   ptbh_index_count(void* h);
void  ptbh_export(void* h, void* nodes_out, int* indices_out, int node_offset, int index_offset);
void  ptbh_free(void* h);
}
#ifndef PTB_DEFAULT_ORDER_BINS
#define PTB_DEFAULT_ORDER_BINS 0
#endif

static void drop_graphs(ptb_ctx* ctx);

static void ctx_fail(ptb_ctx* ctx, const char* what, int code) {
    if (!ctx) return;
    char buf[512];
    snprintf(buf, sizeof(buf), "%s failed with %d (%s)", what, code, code < 1000 ? cudaGetErrorString((cudaError_t)code) : "driver error");
    ctx->last_error = buf;
    fprintf(stderr, "[ptb] %s\n", buf);
}

template <typename T>
static int dev_alloc(ptb_ctx* ctx, T** out, size_t count) {
    void* p = nullptr;
    CK(cudaMalloc(&p, (count ? count : 1) * sizeof(T)));
    ctx->allocs.push_back(p);
    *out = static_cast<T*>(p);
    return 0;
}
template <typename T>
static int film_alloc(ptb_ctx* ctx, T** out, size_t count) {
    void* p = nullptr;
    CK(cudaMalloc(&p, (count ? count : 1) * sizeof(T)));
    ctx->film_allocs.push_back(p);
    *out = static_cast<T*>(p);
    return 0;
}
template <typename T>
static int dev_upload(ptb_ctx* ctx, const T** out, const void* src, size_t count) {
    T* p = nullptr;
    int e = dev_alloc(ctx, &p, count);
    if (e) return e;
    if (count) CK(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *out = p;
    return 0;
}

static int owned_rows_of(int height, int rank, int world, int band) {
    int rows = 0;
    for (int y = 0; y < height; y++) if ((y / band) % world == rank) rows++;
    return rows;
}

static int grid_for(const ptb_ctx* ctx, int blocks_per_sm) { return ctx->sm_count * blocks_per_sm; }

// Ray queues, material/shadow queues and the per-slot framebuffer planes of every enabled AOV, for `samples` pass slots.
template <typename T>
static int wave_alloc(ptb_ctx* ctx, T** out, size_t count) {
    void* p = nullptr;
    CK(cudaMalloc(&p, (count ? count : 1) * sizeof(T)));
    ctx->wave_allocs.push_back(p);
    *out = static_cast<T*>(p);
    return 0;
}
static int allocate_wave_storage(ptb_ctx* ctx, int samples) {
    Frame& F = ctx->F;
    if (samples < 1 || F.pix_bits + 1 > 30 || samples > (1 << (30 - F.pix_bits))) return PTB_E_BADARG;
    CK(cudaStreamSynchronize(ctx->stream));
    for (void* p : ctx->wave_allocs) cudaFree(p);
    ctx->wave_allocs.clear();
    const size_t N = (size_t)F.local_pixels * samples;
    int e = 0;
    for (int i = 0; i < 2; i++) {
        e |= wave_alloc(ctx, &F.q[i].od0, N); e |= wave_alloc(ctx, &F.q[i].od1, N); e |= wave_alloc(ctx, &F.q[i].hit, N);
        e |= wave_alloc(ctx, &F.q[i].path, N); e |= wave_alloc(ctx, &F.q[i].pix, N); e |= wave_alloc(ctx, &F.q[i].medium, N);
    }
    e |= wave_alloc(ctx, &F.sq.od0, N); e |= wave_alloc(ctx, &F.sq.od1, N); e |= wave_alloc(ctx, &F.sq.illum, N);
    for (int m = 0; m < 4; m++) e |= wave_alloc(ctx, &F.matq[m], N);
    F.order = nullptr; F.bin_rank = nullptr; F.bin_key = nullptr;
    if (F.order_bins > 0) { e |= wave_alloc(ctx, &F.order, N); e |= wave_alloc(ctx, &F.bin_rank, N); e |= wave_alloc(ctx, &F.bin_key, N); }   // 9 B per slot, only when ordering is on
    const size_t plane = (size_t)F.pitch * F.height;
    for (int k = 0; k < PTB_AOV_COUNT; k++) {
        if (!(F.config.aov_mask & (1u << k))) { F.aov[k].fb = nullptr; continue; }
        e |= wave_alloc(ctx, &F.aov[k].fb, plane * samples);
        if (!e) CK(cudaMemsetAsync(F.aov[k].fb, 0, plane * samples * sizeof(float4), ctx->stream));
        if (!F.aov[k].acc) { e |= film_alloc(ctx, &F.aov[k].acc, plane); if (!e) CK(cudaMemsetAsync(F.aov[k].acc, 0, plane * sizeof(float4), ctx->stream)); }
    }
    if (e) return PTB_E_STATE;
    ctx->wave_capacity = samples;
    return 0;
}

// Everything whose size follows the film (the role of Pathtracer::resize_init, Pathtracer.cpp:255-301): display image, hit tap,
// accumulators, wave storage, SVGF state.  F.width / F.height must be set; called by ptb_create and ptb_resize.
static int ensure_svgf(ptb_ctx* ctx);
static int allocate_film(ptb_ctx* ctx) {
    Frame& F = ctx->F;
    F.pitch = (F.width + 31) / 32 * 32;
    ctx->owned_rows = owned_rows_of(F.height, F.rank, F.world, F.band_rows);
    F.local_pixels = ctx->owned_rows * F.width;
    const size_t pixels = (size_t)F.pitch * F.height;
    F.fb_stride = (int)pixels;
    F.pix_bits = 1; while ((1ull << F.pix_bits) < pixels) F.pix_bits++;
    F.wave_samples = 1; F.first_sample = 0;
    if (film_alloc(ctx, &F.display, pixels) || film_alloc(ctx, &ctx->tap_hits, pixels)) return PTB_E_STATE;
    CK(cudaMemsetAsync(F.display, 0, pixels * sizeof(float4), ctx->stream));
    int e = allocate_wave_storage(ctx, ctx->wave_capacity); if (e) return e;
    if (F.config.enable_svgf) { e = ensure_svgf(ctx); if (e) return e; }
    return 0;
}

// One pinned arena = [segment table | payloads].  stage_upload() only appends; staging_flush() ships the arena with ONE H2D copy
// and ONE small kernel that applies the segments in order (k_apply_uploads) -- nine separate small copies cost ~0.3 ms of
// serialised copy-engine latency per frame, measured in bench.py's end-to-end loop.
struct UploadSegment { unsigned long long dst; unsigned offset, words; };
#define PTB_STAGE_MAX_SEGMENTS 32
#define PTB_STAGE_TABLE_BYTES (16 + PTB_STAGE_MAX_SEGMENTS * sizeof(UploadSegment))
__global__ void __launch_bounds__(1024) k_apply_uploads(const unsigned char* arena) {
    const int n = *reinterpret_cast<const int*>(arena);
    const UploadSegment* seg = reinterpret_cast<const UploadSegment*>(arena + 16);
    for (int s = 0; s < n; s++) {                  // in order: later segments may overwrite earlier ones (raw TLAS, then the pruned copy)
        unsigned* dst = reinterpret_cast<unsigned*>(seg[s].dst);
        const unsigned* src = reinterpret_cast<const unsigned*>(arena + seg[s].offset);
        for (unsigned i = threadIdx.x; i < seg[s].words; i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
}

// CUDA loads kernels lazily, on their first launch, and that load synchronises the device.  A frame that ends in a device-side
// wait for a peer (k_exchange_wait, k_svgf_wait_*) must never meet such a load while the peer's work is still to be enqueued,
// so every kernel of the library is loaded when the first context is created.
template <typename K> static void preload(K kernel) { cudaFuncAttributes a; cudaFuncGetAttributes(&a, kernel); }
static void preload_kernels() {
    preload(k_generate); preload(k_begin_pass); preload(k_fold_counters); preload(k_sort); preload(k_accumulate);
    preload(k_trace8<false, false>); preload(k_trace8<true, false>); preload(k_trace8<false, true>); preload(k_trace8<true, true>);
    preload(k_trace2<false, false>); preload(k_trace2<true, false>); preload(k_trace2<false, true>); preload(k_trace2<true, true>);
    preload(k_trace4<false, false>); preload(k_trace4<true, false>); preload(k_trace4<false, true>); preload(k_trace4<true, true>);
    preload(k_shade<BSDFDiffuse>); preload(k_shade<BSDFPlastic>); preload(k_shade<BSDFDielectric>); preload(k_shade<BSDFConductor>);
    preload(k_bin_count<false>); preload(k_bin_count<true>); preload(k_bin_scatter<false>); preload(k_bin_scatter<true>);
    preload(k_tap_primary_hits); preload(k_refit_tlas); preload(k_retire_merged_slots); preload(k_export_rows); preload(k_assemble_rows);
    preload(k_exchange_wait); preload(k_svgf_push); preload(k_svgf_wait_arrivals); preload(k_svgf_push_display);
    preload(k_svgf_reproject); preload(k_svgf_variance); preload(k_svgf_atrous<0, false>); preload(k_svgf_atrous<1, false>); preload(k_svgf_atrous<2, false>); preload(k_svgf_atrous<1, true>); preload(k_svgf_atrous<2, true>); preload(k_svgf_finalize); preload(k_taa); preload(k_taa_finalize);
    preload(k_clear_framebuffers); preload(k_apply_uploads); preload(k_present); preload(k_ambient_occlusion);
    preload(k_integrate_dielectric); preload(k_average_dielectric); preload(k_integrate_conductor); preload(k_average_conductor); preload(k_dump_luts);
    cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- lifetime
extern "C" int ptb_create(ptb_ctx** out, int device, int width, int height, int rank, int world, int band_rows) {
    if (!out || width <= 0 || height <= 0 || world <= 0 || rank < 0 || rank >= world || band_rows <= 0) return PTB_E_BADARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { fprintf(stderr, "[ptb] no CUDA device: the product path has no CPU fallback\n"); return e != cudaSuccess ? (int)e : (int)cudaErrorNoDevice; }
    if (device < 0 || device >= n) return PTB_E_BADARG;
    ptb_ctx* ctx = new (std::nothrow) ptb_ctx();
    if (!ctx) return PTB_E_STATE;
    ctx->device = device;
    CKC(cudaSetDevice(device));
    CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
    CKC(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    if (const char* v = getenv("PTB_TRACE_OVERLAP")) ctx->overlap_enabled = atoi(v) != 0;     // A/B switch for tools/, default on
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;

    Frame& F = ctx->F;
    memset(&F, 0, sizeof(F));
    F.width = width; F.height = height;
    F.rank = rank; F.world = world; F.band_rows = band_rows;

    // defaults of GPUConfig (Common.h:39-67)
    F.config.reconstruction_filter = 2; F.config.aov_mask = 1u; F.config.num_bounces = 10;
    F.config.enable_mipmapping = 1; F.config.enable_next_event_estimation = 1; F.config.enable_multiple_importance_sampling = 1;
    F.config.enable_russian_roulette = 1; F.config.enable_svgf = 0; F.config.enable_spatial_variance = 1; F.config.enable_taa = 1;
    F.config.alpha_colour = 0.1f; F.config.alpha_moment = 0.1f; F.config.num_atrous_iterations = 6;
    F.config.sigma_z = 4.0f; F.config.sigma_n = 16.0f; F.config.sigma_l = 10.0f;

    if (dev_alloc(ctx, &F.counters, 1) || dev_alloc(ctx, &F.totals, 1) || dev_alloc(ctx, &F.trace_stats, 2)) { ptb_destroy(ctx); return PTB_E_STATE; }
    if (dev_alloc(ctx, &F.bin_counts, PTB_ORDER_MAX_BOUNCE * 2 * PTB_ORDER_MAX_BINS)) { ptb_destroy(ctx); return PTB_E_STATE; }
    CKC(cudaMemsetAsync(F.bin_counts, 0, sizeof(int) * PTB_ORDER_MAX_BOUNCE * 2 * PTB_ORDER_MAX_BINS, ctx->stream));
    F.order_bins = PTB_DEFAULT_ORDER_BINS;
    CKC(cudaMemsetAsync(F.trace_stats, 0, 2 * sizeof(TraceStats), ctx->stream));
    CKC(cudaMemsetAsync(F.counters, 0, sizeof(Counters), ctx->stream));
    CKC(cudaMemsetAsync(F.totals, 0, sizeof(RayTotals), ctx->stream));
    if (dev_alloc(ctx, &F.pixel_query, 4)) { ptb_destroy(ctx); return PTB_E_STATE; }
    CKC(cudaMemsetAsync(F.pixel_query, 0xff, 4 * sizeof(int), ctx->stream));      // {INVALID, INVALID, INVALID}: no query pending
    F.config.aov_mask = 1u;          // RADIANCE is always on (Pathtracer.cpp:267-268)
    ctx->wave_capacity = 1;
    { int fe = allocate_film(ctx); if (fe) { ptb_destroy(ctx); return fe; } }

    preload_kernels();
    CKC(cudaFuncSetAttribute(k_trace8<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CKC(cudaFuncSetAttribute(k_trace8<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CKC(cudaFuncSetAttribute(k_trace8<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CKC(cudaFuncSetAttribute(k_trace8<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CKC(cudaStreamSynchronize(ctx->stream));
    *out = ctx;
    return 0;
}

extern "C" void ptb_destroy(ptb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    for (void*& m : ctx->xchg_ipc_opened) if (m) { cudaIpcCloseMemHandle(m); m = nullptr; }
    if (ctx->xchg_block) { cudaFree(ctx->xchg_block); ctx->xchg_block = nullptr; }
    for (int c = 0; c < 2; c++) { if (ctx->stage_mem[c]) cudaFreeHost(ctx->stage_mem[c]); if (ctx->stage_dev[c]) cudaFree(ctx->stage_dev[c]); if (ctx->stage_done[c]) cudaEventDestroy(ctx->stage_done[c]); }
    if (ctx->merge_nodes) cudaFree(ctx->merge_nodes);
    if (ctx->merge_tris) cudaFree(ctx->merge_tris);
    if (ctx->merge_woop) cudaFree(ctx->merge_woop);
    if (ctx->merge_who) cudaFree(ctx->merge_who);
    if (ctx->merge_slot_instance_dev) cudaFree(ctx->merge_slot_instance_dev);
    if (ctx->refit_scratch) cudaFree(ctx->refit_scratch);
    if (g_drv.ok) for (auto& t : ctx->textures) { if (t.tex) g_drv.TexObjectDestroy(t.tex); if (t.array) g_drv.MipmappedArrayDestroy(t.array); }
    if (ctx->F.sky_tex) cudaDestroyTextureObject(ctx->F.sky_tex);
    if (ctx->sky_array) cudaFreeArray(ctx->sky_array);
    cudaTextureObject_t luts[6] = { ctx->F.lut_dielectric_dir_enter, ctx->F.lut_dielectric_dir_leave, ctx->F.lut_dielectric_enter,
                                    ctx->F.lut_dielectric_leave, ctx->F.lut_conductor_dir, ctx->F.lut_conductor };
    for (int i = 0; i < 6; i++) { if (luts[i]) cudaDestroyTextureObject(luts[i]); if (ctx->lut_arrays[i]) cudaFreeArray(ctx->lut_arrays[i]); }
    drop_graphs(ctx);
    for (void* p : ctx->wave_allocs) cudaFree(p);
    for (void* p : ctx->film_allocs) cudaFree(p);
    for (void* p : ctx->allocs) cudaFree(p);
    for (auto ev : ctx->event_pool) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// ---------------------------------------------------------------------------------------------- AOV management
static int ensure_aov(ptb_ctx* ctx, int k) {
    Frame& F = ctx->F;
    if (F.aov[k].fb) return 0;
    const size_t plane = (size_t)F.pitch * F.height;
    int e = wave_alloc(ctx, &F.aov[k].fb, plane * ctx->wave_capacity); if (e) return e;
    CK(cudaMemsetAsync(F.aov[k].fb, 0, plane * ctx->wave_capacity * sizeof(float4), ctx->stream));
    if (!F.aov[k].acc) { e = film_alloc(ctx, &F.aov[k].acc, plane); if (e) return e; CK(cudaMemsetAsync(F.aov[k].acc, 0, plane * sizeof(float4), ctx->stream)); }
    return 0;
}

static int ensure_svgf(ptb_ctx* ctx) {
    Frame& F = ctx->F;
    if (F.svgf.moment) return 0;
    const size_t pixels = (size_t)F.pitch * F.height;
    int e = 0;
    e |= film_alloc(ctx, &F.svgf.gbuf_normal_depth, pixels); e |= film_alloc(ctx, &F.svgf.gbuf_ids, pixels); e |= film_alloc(ctx, &F.svgf.gbuf_screen_prev, pixels);
    e |= film_alloc(ctx, &F.svgf.moment, pixels); e |= film_alloc(ctx, &F.svgf.taa_curr, pixels);
    if (e) return PTB_E_STATE;
    CK(cudaMemsetAsync(F.svgf.gbuf_normal_depth, 0, pixels * sizeof(float4), ctx->stream));
    CK(cudaMemsetAsync(F.svgf.gbuf_ids, 0, pixels * sizeof(int2), ctx->stream));
    CK(cudaMemsetAsync(F.svgf.gbuf_screen_prev, 0, pixels * sizeof(float2), ctx->stream));
    CK(cudaMemsetAsync(F.svgf.moment, 0, pixels * sizeof(float4), ctx->stream));
    CK(cudaMemsetAsync(F.svgf.taa_curr, 0, pixels * sizeof(float4), ctx->stream));
    // temporal state, two parities (one GPU; with several GPUs it lives in the exchange block, see render_wave)
    for (int p = 0; p < 2; p++) {
        SVGFHistory& h = ctx->local_hist[p];
        e |= film_alloc(ctx, &h.direct, pixels); e |= film_alloc(ctx, &h.indirect, pixels); e |= film_alloc(ctx, &h.moment, pixels);
        e |= film_alloc(ctx, &h.normal_depth, pixels); e |= film_alloc(ctx, &h.taa, pixels); e |= film_alloc(ctx, &h.length, pixels);
        if (e) return PTB_E_STATE;
        float4* planes[5] = { h.direct, h.indirect, h.moment, h.normal_depth, h.taa };
        for (float4* q : planes) CK(cudaMemsetAsync(q, 0, pixels * sizeof(float4), ctx->stream));
        CK(cudaMemsetAsync(h.length, 0, pixels * sizeof(int), ctx->stream));
    }
    ctx->svgf_frames = 0;
    return 0;
}

static void drop_graphs(ptb_ctx* ctx) {
    for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
    ctx->graphs.clear();
}

extern "C" int ptb_set_config(ptb_ctx* ctx, const ptb_config* config) {
    if (!ctx || !config) return PTB_E_BADARG;
    { ptb_config incoming = *config; incoming.aov_mask |= ctx->F.config.aov_mask & 1u;
      if (memcmp(&incoming, &ctx->F.config, sizeof(incoming)) != 0) drop_graphs(ctx); }
    if (config->num_bounces < 1 || config->num_bounces > PTB_MAX_BOUNCES - 1) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    ctx->F.config = *config;
    ctx->F.config.aov_mask |= 1u;
    if (config->enable_svgf) {   // svgf_init enables DIRECT / INDIRECT / ALBEDO (Pathtracer.cpp:331-333)
        ctx->F.config.aov_mask |= (1u << PTB_AOV_RADIANCE_DIRECT) | (1u << PTB_AOV_RADIANCE_INDIRECT) | (1u << PTB_AOV_ALBEDO);
        int e = ensure_svgf(ctx); if (e) return e;
    }
    for (int k = 0; k < PTB_AOV_COUNT; k++) if (ctx->F.config.aov_mask & (1u << k)) { if (!ctx->F.aov[k].fb) drop_graphs(ctx); int e = ensure_aov(ctx, k); if (e) return e; }
    ctx->frames_since_reset = 0;
    return 0;
}

extern "C" int ptb_set_camera(ptb_ctx* ctx, const ptb_camera* camera, const float* vp, const float* vp_prev) {
    if (!ctx || !camera) return PTB_E_BADARG;
    if (memcmp(&ctx->F.camera, camera, sizeof(*camera)) != 0 || (vp && memcmp(ctx->F.svgf.view_projection, vp, 64) != 0) ||
        (vp_prev && memcmp(ctx->F.svgf.view_projection_prev, vp_prev, 64) != 0)) drop_graphs(ctx);   // the parameter block is baked into captured graphs
    ctx->F.camera = *camera;
    if (vp) memcpy(ctx->F.svgf.view_projection, vp, 64);
    if (vp_prev) memcpy(ctx->F.svgf.view_projection_prev, vp_prev, 64);
    else if (vp) memcpy(ctx->F.svgf.view_projection_prev, vp, 64);
    return 0;
}

// ---------------------------------------------------------------------------------------------- textures
static int create_texture(ptb_ctx* ctx, const ptb_texture& t, DeviceTexture& out) {
    // Same object the reference builds (Integrator.cpp:42-94): mip-mapped array, wrap addressing, trilinear,
    // anisotropic; BC1 = 2 x uint32 per 4x4 block viewed through a BC1 resource view.
    const bool bc1 = t.format == 1;
    const int aw = bc1 ? (t.width + 3) / 4 : t.width, ah = bc1 ? (t.height + 3) / 4 : t.height;
    CUDA_ARRAY3D_DESCRIPTOR ad; memset(&ad, 0, sizeof(ad));
    ad.Width = aw; ad.Height = ah; ad.Depth = 0;
    ad.NumChannels = bc1 ? 2 : 4;
    ad.Format = bc1 ? CU_AD_FORMAT_UNSIGNED_INT32 : CU_AD_FORMAT_UNSIGNED_INT8;
    CKD(g_drv.MipmappedArrayCreate(&out.array, &ad, t.num_levels));
    for (int l = 0; l < t.num_levels; l++) {
        CUarray level;
        CKD(g_drv.MipmappedArrayGetLevel(&level, out.array, l));
        int lw = aw >> l; if (lw < 1) lw = 1;
        int lh = ah >> l; if (lh < 1) lh = 1;
        CUDA_MEMCPY2D cp; memset(&cp, 0, sizeof(cp));
        cp.srcMemoryType = CU_MEMORYTYPE_HOST; cp.srcHost = t.levels[l];
        cp.srcPitch = (size_t)lw * (bc1 ? 8 : 4);
        cp.dstMemoryType = CU_MEMORYTYPE_ARRAY; cp.dstArray = level;
        cp.WidthInBytes = cp.srcPitch; cp.Height = lh;
        CKD(g_drv.Memcpy2D(&cp));
    }
    CUDA_RESOURCE_DESC rd; memset(&rd, 0, sizeof(rd));
    rd.resType = CU_RESOURCE_TYPE_MIPMAPPED_ARRAY; rd.res.mipmap.hMipmappedArray = out.array;
    CUDA_TEXTURE_DESC td; memset(&td, 0, sizeof(td));
    td.addressMode[0] = CU_TR_ADDRESS_MODE_WRAP; td.addressMode[1] = CU_TR_ADDRESS_MODE_WRAP; td.addressMode[2] = CU_TR_ADDRESS_MODE_CLAMP;
    td.filterMode = CU_TR_FILTER_MODE_LINEAR; td.mipmapFilterMode = CU_TR_FILTER_MODE_LINEAR;
    td.mipmapLevelBias = 0.0f; td.maxAnisotropy = 16;   // the reference queries GL_MAX_TEXTURE_MAX_ANISOTROPY_EXT (16 on NVIDIA)
    td.minMipmapLevelClamp = 0.0f; td.maxMipmapLevelClamp = float(t.num_levels - 1);
    td.flags = CU_TRSF_NORMALIZED_COORDINATES;
    CUDA_RESOURCE_VIEW_DESC vd; memset(&vd, 0, sizeof(vd));
    vd.format = bc1 ? CU_RES_VIEW_FORMAT_UNSIGNED_BC1 : CU_RES_VIEW_FORMAT_UINT_4X8;
    vd.width = bc1 ? (size_t)aw * 4 : aw; vd.height = bc1 ? (size_t)ah * 4 : ah;
    vd.firstMipmapLevel = 0; vd.lastMipmapLevel = t.num_levels - 1;
    CKD(g_drv.TexObjectCreate(&out.tex, &rd, &td, &vd));
    return 0;
}

static int create_float_texture(ptb_ctx* ctx, cudaArray_t* arr, cudaTextureObject_t* tex, const float* host, int channels,
                                int w, int h, int d, bool device_src) {
    cudaChannelFormatDesc cd = channels == 4 ? cudaCreateChannelDesc<float4>() : cudaCreateChannelDesc<float>();
    if (d > 0) CK(cudaMalloc3DArray(arr, &cd, make_cudaExtent(w, h, d)));
    else       CK(cudaMallocArray(arr, &cd, w, h));
    size_t texel = sizeof(float) * channels;
    cudaMemcpyKind kind = device_src ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (d > 0) {
        cudaMemcpy3DParms p; memset(&p, 0, sizeof(p));
        p.srcPtr = make_cudaPitchedPtr(const_cast<float*>(host), w * texel, w, h);
        p.dstArray = *arr; p.extent = make_cudaExtent(w, h, d); p.kind = kind;
        CK(cudaMemcpy3D(&p));
    } else {
        CK(cudaMemcpy2DToArray(*arr, 0, 0, host, w * texel, w * texel, h > 0 ? h : 1, kind));
    }
    cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray; rd.res.array.array = *arr;
    cudaTextureDesc td; memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 1;
    CK(cudaCreateTextureObject(tex, &rd, &td, nullptr));
    return 0;
}

// ---------------------------------------------------------------------------------------------- Kulla-Conty LUT bake
static int bake_luts(ptb_ctx* ctx) {
    if (ctx->luts_ready) return 0;
    Frame& F = ctx->F;
    const int D = PTB_LUT_DIELECTRIC_DIM, C = PTB_LUT_CONDUCTOR_DIM;
    float *dir = nullptr, *avg = nullptr, *cdir = nullptr, *cavg = nullptr;
    struct Scratch { float **a, **b, **c, **d; ~Scratch() { cudaFree(*a); cudaFree(*b); cudaFree(*c); cudaFree(*d); } } scratch{ &dir, &avg, &cdir, &cavg };
    CK(cudaMalloc(&dir, sizeof(float) * D * D * D));
    CK(cudaMalloc(&avg, sizeof(float) * D * D));
    cudaTextureObject_t* dir_tex[2] = { &F.lut_dielectric_dir_enter, &F.lut_dielectric_dir_leave };
    cudaTextureObject_t* avg_tex[2] = { &F.lut_dielectric_enter, &F.lut_dielectric_leave };
    for (int pass = 0; pass < 2; pass++) {
        bool entering = pass == 0;
        k_integrate_dielectric<<<(D * D * D + 255) / 256, 256, 0, ctx->stream>>>(F, entering, dir);
        k_average_dielectric<<<(D * D + 255) / 256, 256, 0, ctx->stream>>>(dir, avg);
        ctx->launches += 2;
        CK(cudaStreamSynchronize(ctx->stream));
        int e = create_float_texture(ctx, &ctx->lut_arrays[pass], dir_tex[pass], dir, 1, D, D, D, true); if (e) return e;
        e = create_float_texture(ctx, &ctx->lut_arrays[2 + pass], avg_tex[pass], avg, 1, D, D, 0, true); if (e) return e;
    }
    CK(cudaMalloc(&cdir, sizeof(float) * C * C));
    CK(cudaMalloc(&cavg, sizeof(float) * C));
    k_integrate_conductor<<<(C * C + 255) / 256, 256, 0, ctx->stream>>>(F, cdir);
    k_average_conductor<<<1, 256, 0, ctx->stream>>>(cdir, cavg);
    ctx->launches += 2;
    CK(cudaStreamSynchronize(ctx->stream));
    int e = create_float_texture(ctx, &ctx->lut_arrays[4], &F.lut_conductor_dir, cdir, 1, C, C, 0, true); if (e) return e;
    e = create_float_texture(ctx, &ctx->lut_arrays[5], &F.lut_conductor, cavg, 1, C, 0, 0, true); if (e) return e;
    ctx->luts_ready = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------- scene upload
// ---------------------------------------------------------------------------------------------- static merge
// The reference traces a two-level hierarchy: a TLAS over instances, one BLAS per mesh (Integrator.cpp:101-283, BVH8.h:204-232).
// Sponza is 384 instances with heavily overlapping boxes, all with identity transforms: measured on the B200, the same triangles
// in ONE CWBVH cost 35 % less trace time with a bit-identical image (tools/gpu_flat.py).  So at upload every instance whose
// transform is the identity (root bit 31, i.e. the ray is NOT transformed on entry -- intersections are the same float
// operations either way) is merged: its triangles are collected by walking its BLAS leaves, one SAH/CWBVH build (the CPU
// builder of host/bvh_build.cpp) runs over all of them, the nodes are appended to the node array in breadth-first order and the
// triangles are stored as compact 48-byte records carrying the original triangle id + merged slot.  Rays walk the merged BVH
// first and skip merged instances in the TLAS (PTB_ROOT_MERGED); hits report the original (mesh_id, triangle_id).
static int staging_begin(ptb_ctx* ctx, size_t need) {
    need += PTB_STAGE_TABLE_BYTES;
    ctx->stage_cur ^= 1;
    const int c = ctx->stage_cur;
    if (ctx->stage_done[c]) CK(cudaEventSynchronize(ctx->stage_done[c]));      // the copy that last read this arena (two calls ago) is long done
    else CK(cudaEventCreateWithFlags(&ctx->stage_done[c], cudaEventDisableTiming));
    if (ctx->stage_cap[c] < need) {
        if (ctx->stage_mem[c]) cudaFreeHost(ctx->stage_mem[c]);
        if (ctx->stage_dev[c]) cudaFree(ctx->stage_dev[c]);
        ctx->stage_mem[c] = nullptr; ctx->stage_dev[c] = nullptr; ctx->stage_cap[c] = 0;
        CK(cudaMallocHost(reinterpret_cast<void**>(&ctx->stage_mem[c]), need));
        CK(cudaMalloc(reinterpret_cast<void**>(&ctx->stage_dev[c]), need));
        ctx->stage_cap[c] = need;
    }
    ctx->stage_used = PTB_STAGE_TABLE_BYTES; ctx->stage_segments = 0; ctx->stage_open = true;
    return 0;
}
// ship what has been appended so far (one H2D copy + one kernel); the arena stays open for more segments
static int staging_flush(ptb_ctx* ctx) {
    const int c = ctx->stage_cur;
    if (!ctx->stage_open || ctx->stage_segments == 0) return 0;
    *reinterpret_cast<int*>(ctx->stage_mem[c]) = ctx->stage_segments;
    // a flush in the middle of a call must not overwrite arena bytes that an earlier, still running flush reads: the device
    // mirror is only re-used from offset 0 because flushes of one arena are stream-ordered
    CK(cudaMemcpyAsync(ctx->stage_dev[c], ctx->stage_mem[c], ctx->stage_used, cudaMemcpyHostToDevice, ctx->stream));
    k_apply_uploads<<<1, 1024, 0, ctx->stream>>>(ctx->stage_dev[c]); ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->stage_done[c], ctx->stream));
    ctx->stage_segments = 0;
    return 0;
}
// host -> device copy whose source may be changed by the caller as soon as we return
static int stage_upload(ptb_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!bytes) return 0;
    const int c = ctx->stage_cur;
    const size_t padded = (bytes + 15) & ~size_t(15);
    if (!ctx->stage_open || (bytes & 3) || (reinterpret_cast<size_t>(dst) & 3)) {      // no arena / odd size: plain synchronous copy, in order
        int fe = staging_flush(ctx); if (fe) return fe;
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    if (ctx->stage_segments == PTB_STAGE_MAX_SEGMENTS || ctx->stage_used + padded > ctx->stage_cap[c]) {
        // segment table or arena full: ship this arena and go on in the other one (stream order keeps the segments in order); the
        // call stays asynchronous -- round 1 fell back to a blocking copy here
        int fe = staging_flush(ctx); if (fe) return fe;
        size_t want = ctx->stage_cap[c] > 2 * padded + 4096 ? ctx->stage_cap[c] : 2 * padded + 4096;
        int be = staging_begin(ctx, want); if (be) return be;
        return stage_upload(ctx, dst, src, bytes);
    }
    unsigned char* at = ctx->stage_mem[c] + ctx->stage_used;
    memcpy(at, src, bytes);
    UploadSegment* seg = reinterpret_cast<UploadSegment*>(ctx->stage_mem[c] + 16) + ctx->stage_segments++;
    seg->dst = reinterpret_cast<unsigned long long>(dst); seg->offset = (unsigned)ctx->stage_used; seg->words = (unsigned)(bytes / 4);
    ctx->stage_used += padded;
    return 0;
}
static int staging_end(ptb_ctx* ctx) {
    int fe = staging_flush(ctx);
    ctx->stage_open = false;
    return fe;
}
// Ship what is staged and keep staging in the OTHER arena: for work that must be enqueued between two groups of uploads of one call.
// (Appending to the same pinned arena after a flush would rewrite its segment table while the flush's asynchronous copy may still be
// pending.)
static int staging_split(ptb_ctx* ctx) {
    if (!ctx->stage_open) return 0;
    const size_t cap = ctx->stage_cap[ctx->stage_cur];
    int fe = staging_flush(ctx); if (fe) return fe;
    return staging_begin(ctx, cap > PTB_STAGE_TABLE_BYTES ? cap - PTB_STAGE_TABLE_BYTES : 4096);
}

static void upload_roots(ptb_ctx* ctx, std::vector<int>& device_roots) {
    stage_upload(ctx, (void*)ctx->F.mesh_roots, device_roots.data(), sizeof(int) * device_roots.size());
    if (ctx->merge_slot_instance_dev && !ctx->merge_slot_instance.empty())
        stage_upload(ctx, ctx->merge_slot_instance_dev, ctx->merge_slot_instance.data(), sizeof(int) * ctx->merge_slot_instance.size());
}

// The host's TLAS still lists the merged instances.  In OUR copy of it (the one rays walk) every leaf slot whose instances are
// all merged, and every internal child whose whole subtree is, gets meta = 0 -- "empty slot" to the node test, so rays never
// descend towards instances they already intersected through the merged BVH (imask is left alone: it drives child indexing).
static int upload_pruned_tlas(ptb_ctx* ctx) {
    Frame& F = ctx->F;
    if (F.flat_root < 0 || !ctx->merge_nodes || F.tlas_nodes <= 0) return 0;
    std::vector<char> merged(ctx->host_roots.size(), 0);
    for (int i : ctx->merge_slot_instance) if (i >= 0 && (size_t)i < merged.size()) merged[i] = 1;
    std::vector<unsigned char> tl(ctx->host_nodes.begin(), ctx->host_nodes.begin() + (size_t)F.tlas_nodes * 80);
    bool all = ptb_merge::prune_tlas(tl.data(), 0, merged);
    int flat_all = all ? 1 : 0;
    if (flat_all != F.flat_all) { drop_graphs(ctx); F.flat_all = flat_all; }
    return stage_upload(ctx, ctx->merge_nodes, tl.data(), tl.size());
}

static std::vector<int> identity_roots_sorted(const std::vector<int>& roots) {
    std::vector<int> r;
    for (int v : roots) if ((unsigned)v & PTB_ROOT_IDENTITY) r.push_back(v);
    std::sort(r.begin(), r.end());
    return r;
}

static int rebuild_static_merge(ptb_ctx* ctx) {
    Frame& F = ctx->F;
    drop_graphs(ctx);
    { int fe = staging_flush(ctx); if (fe) return fe; }      // staged TLAS nodes must be on the device before they are copied below
    CK(cudaStreamSynchronize(ctx->stream));
    const int M = (int)ctx->host_roots.size();
    std::vector<int> slots;
    if (ctx->merge_enabled && ctx->bvh_kind == 8)
        for (int i = 0; i < M; i++) if ((unsigned)ctx->host_roots[i] & PTB_ROOT_IDENTITY) slots.push_back(i);
    std::vector<int> device_roots = ctx->host_roots;
    ctx->merge_slot_root.clear(); ctx->merge_slot_instance.clear();
    ctx->merge_decided_roots = identity_roots_sorted(ctx->host_roots);
    if (ctx->merge_nodes) { cudaFree(ctx->merge_nodes); ctx->merge_nodes = nullptr; }
    if (ctx->merge_tris) { cudaFree(ctx->merge_tris); ctx->merge_tris = nullptr; ctx->merge_ref_count = 0; }
    if (ctx->merge_woop) { cudaFree(ctx->merge_woop); ctx->merge_woop = nullptr; }
    if (ctx->merge_who) { cudaFree(ctx->merge_who); ctx->merge_who = nullptr; }
    F.flat_woop = nullptr; F.flat_who = nullptr;
    if (ctx->merge_slot_instance_dev) { cudaFree(ctx->merge_slot_instance_dev); ctx->merge_slot_instance_dev = nullptr; }
    F.flat_root = -1; F.flat_all = 0; F.flat_node_count = 0; F.flat_tris = nullptr; F.flat_slot_instance = nullptr;
    if (ctx->bvh_kind == 8) F.nodes8 = ctx->uploaded_nodes;
    if (slots.empty()) { upload_roots(ctx, device_roots); return 0; }

    // triangles of the merged instances
    std::vector<float> pos; std::vector<int2> who;          // (original triangle id, merged slot)
    std::vector<int> tris;
    for (size_t k = 0; k < slots.size(); k++) {
        int root = ctx->host_roots[slots[k]];
        tris.clear();
        ptb_merge::collect_leaf_primitives(ctx->host_nodes.data(), (unsigned)root & 0x3fffffffu, tris);
        for (int t : tris) {
            const float4* r = &ctx->host_tri_pos[(size_t)t * 3];
            float3 p0 = make_float3(r[0].x, r[0].y, r[0].z), e1 = make_float3(r[0].w, r[1].x, r[1].y), e2 = make_float3(r[1].z, r[1].w, r[2].x);
            float v[9] = { p0.x, p0.y, p0.z, p0.x + e1.x, p0.y + e1.y, p0.z + e1.z, p0.x + e2.x, p0.y + e2.y, p0.z + e2.z };
            pos.insert(pos.end(), v, v + 9);
            who.push_back(make_int2(t, (int)k));
        }
        ctx->merge_slot_root.push_back(root); ctx->merge_slot_instance.push_back(slots[k]);
        device_roots[slots[k]] = int((unsigned)root | PTB_ROOT_MERGED);
    }
    const int n = (int)who.size();
    if (n == 0) { ctx->merge_slot_root.clear(); ctx->merge_slot_instance.clear(); upload_roots(ctx, ctx->host_roots); return 0; }
    // One CWBVH over all of them.  Default: split BVH (spatial splits, host/bvh_build.cpp SpatialBuilder; the reference's second
    // builder, Src/BVH/Builders/SBVHBuilder.cpp) -- Sponza's long wall / floor triangles overlap everything in a plain SAH tree;
    // measured with ptbh_trace_stats on bounce rays: 22.8 -> 17.6 node visits and 13.0 -> 8.4 triangle tests per ray for 11 % more
    // triangle references.  A reference is a (clipped box, triangle) pair: a triangle can sit in several leaves, every leaf entry gets
    // its own 48-byte record below.  The merged tree is deeper than any single BLAS: if it could outgrow the traversal stack
    // (BVH_STACK_SIZE = 32 entries, two per level, plus the TLAS root that waits underneath) the plain SAH tree is tried, then the
    // two-level walk is kept.
    void* h = nullptr; int nm = 0;
    std::vector<unsigned char> dfs;
    std::vector<int> order;
    for (int attempt = ctx->merge_spatial ? 0 : 1; attempt < 2; attempt++) {
        // builder parameters: A/B overrides for tools/ (PTB_MERGE_ALPHA / _BINS / _DUP / _TRICOST); the defaults are the measured best
        float alpha = PTB_MERGE_ALPHA_DEFAULT, dup = 2.0f, tri_cost = PTB_MERGE_TRICOST_DEFAULT; int bins = PTB_MERGE_BINS_DEFAULT;
        if (const char* v = getenv("PTB_MERGE_ALPHA")) alpha = (float)atof(v);
        if (const char* v = getenv("PTB_MERGE_BINS")) bins = atoi(v);
        if (const char* v = getenv("PTB_MERGE_DUP")) dup = (float)atof(v);
        if (const char* v = getenv("PTB_MERGE_TRICOST")) tri_cost = (float)atof(v);
        // insertion-based optimisation of the binary split BVH before the wide collapse (host/bvh_build.cpp ReinsertionOptimizer): on
        // Sponza 8 % less SAH cost and 9-25 % fewer node visits per ray from four different views, closest hits unchanged, +2 s of build;
        // the builder keeps the tree as built if the optimised one would not fit the traversal stack
        int opt_passes = PTB_MERGE_OPTIMIZE_DEFAULT; float opt_fraction = 1.0f;
        if (const char* v = getenv("PTB_MERGE_OPTIMIZE")) opt_passes = atoi(v);
        if (const char* v = getenv("PTB_MERGE_OPT_FRACTION")) opt_fraction = (float)atof(v);
        ptbh_set_tri_cost(tri_cost);
        ptbh_set_optimizer(opt_passes, opt_fraction, (PTB_STACK_TOTAL - 3) / 2);
        h = attempt == 0 ? ptbh_build_triangles_sbvh(pos.data(), n, alpha, bins, dup) : ptbh_build_triangles(pos.data(), n, 8, 4.0f, 1.0f);
        ptbh_set_tri_cost(1.0f); ptbh_set_optimizer(0, 1.0f, 0);
        if (!h) return PTB_E_STATE;
        nm = ptbh_node_count(h);
        dfs.assign((size_t)nm * 80, 0); order.assign((size_t)ptbh_index_count(h), 0);
        ptbh_export(h, dfs.data(), order.data(), 0, 0);
        ptbh_free(h);
        if (2 * ptb_merge::max_depth(dfs.data(), 0) + 3 <= PTB_STACK_TOTAL) break;
        nm = 0;
    }
    if (nm == 0) {
        fprintf(stderr, "[ptb] static merge skipped: merged BVH too deep for the %d-entry traversal stack\n", PTB_STACK_TOTAL);
        ctx->merge_slot_root.clear(); ctx->merge_slot_instance.clear();
        upload_roots(ctx, ctx->host_roots);
        return 0;
    }
    std::vector<unsigned char> bfs((size_t)nm * 80);
    const int n_refs = (int)order.size();
    // breadth-first re-layout (children of a node stay contiguous and in slot order); child indices become global
    const int base = ctx->node_count;
    ptb_merge::bfs_relayout(dfs.data(), nm, base, bfs.data());
    std::vector<float4> flat((size_t)n_refs * 3);
    for (int j = 0; j < n_refs; j++) {
        int src = order[j];
        const float4* r = &ctx->host_tri_pos[(size_t)who[src].x * 3];
        flat[(size_t)j * 3 + 0] = r[0]; flat[(size_t)j * 3 + 1] = r[1];
        float id_bits, slot_bits; memcpy(&id_bits, &who[src].x, 4); memcpy(&slot_bits, &who[src].y, 4);
        flat[(size_t)j * 3 + 2] = make_float4(r[2].x, id_bits, slot_bits, 0.0f);
    }
    CK(cudaMalloc(&ctx->merge_nodes, ((size_t)base + nm) * 80));
    CK(cudaMemcpyAsync(ctx->merge_nodes, ctx->uploaded_nodes, (size_t)base * 80, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(reinterpret_cast<unsigned char*>(ctx->merge_nodes) + (size_t)base * 80, bfs.data(), (size_t)nm * 80, cudaMemcpyHostToDevice, ctx->stream));
    // Woop maps (double precision on the host): columns (e1, e2, n, p0) of the triangle's frame, inverted; n = e1 x e2
    std::vector<float4> woop((size_t)n_refs * 3); std::vector<int2> who_ref((size_t)n_refs);
    for (int j = 0; j < n_refs; j++) {
        const float4* r = &ctx->host_tri_pos[(size_t)who[order[j]].x * 3];
        double p0[3] = { r[0].x, r[0].y, r[0].z }, e1[3] = { r[0].w, r[1].x, r[1].y }, e2[3] = { r[1].z, r[1].w, r[2].x };
        double nn[3] = { e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0] };
        double det = nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2];          // det [e1 e2 n] = |n|^2
        double inv = det > 0.0 ? 1.0 / det : 0.0;
        // rows of [e1 e2 n]^-1: (e2 x n) / det, (n x e1) / det, n / det
        double r0[3] = { (e2[1] * nn[2] - e2[2] * nn[1]) * inv, (e2[2] * nn[0] - e2[0] * nn[2]) * inv, (e2[0] * nn[1] - e2[1] * nn[0]) * inv };
        double r1[3] = { (nn[1] * e1[2] - nn[2] * e1[1]) * inv, (nn[2] * e1[0] - nn[0] * e1[2]) * inv, (nn[0] * e1[1] - nn[1] * e1[0]) * inv };
        double r2[3] = { nn[0] * inv, nn[1] * inv, nn[2] * inv };
        auto row = [&](const double* m) { return make_float4((float)m[0], (float)m[1], (float)m[2], (float)-(m[0] * p0[0] + m[1] * p0[1] + m[2] * p0[2])); };
        woop[(size_t)j * 3 + 0] = row(r0); woop[(size_t)j * 3 + 1] = row(r1); woop[(size_t)j * 3 + 2] = row(r2);
        who_ref[j] = who[order[j]];
    }
    CK(cudaMalloc(&ctx->merge_woop, woop.size() * sizeof(float4)));
    CK(cudaMemcpyAsync(ctx->merge_woop, woop.data(), woop.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMalloc(&ctx->merge_who, who_ref.size() * sizeof(int2)));
    CK(cudaMemcpyAsync(ctx->merge_who, who_ref.data(), who_ref.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMalloc(&ctx->merge_tris, flat.size() * sizeof(float4)));
    CK(cudaMemcpyAsync(ctx->merge_tris, flat.data(), flat.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMalloc(&ctx->merge_slot_instance_dev, sizeof(int) * slots.size()));
    CK(cudaStreamSynchronize(ctx->stream));       // `bfs` and `flat` are host temporaries
    F.nodes8 = ctx->merge_nodes;
    F.flat_root = base; F.flat_node_count = nm; F.flat_all = (int)slots.size() == M ? 1 : 0;
    F.flat_tris = ctx->merge_tris; F.flat_slot_instance = ctx->merge_slot_instance_dev; ctx->merge_ref_count = n_refs;
    F.flat_who = ctx->merge_who; F.flat_woop = ctx->intersector == PTB_INTERSECT_WOOP ? ctx->merge_woop : nullptr;
    upload_roots(ctx, device_roots);
    return upload_pruned_tlas(ctx);
}

// New roots from the host (TLAS leaf order may have changed): keep the merged BVH if it still covers exactly the identity
// instances, only refreshing slot -> instance; otherwise rebuild it.
static int apply_roots(ptb_ctx* ctx, const int32_t* roots, int mesh_count) {
    ctx->host_roots.assign(roots, roots + mesh_count);
    if (ctx->bvh_kind != 8) { std::vector<int> r = ctx->host_roots; upload_roots(ctx, r); return 0; }
    std::vector<int> ident;
    for (int i = 0; i < mesh_count; i++) if ((unsigned)roots[i] & PTB_ROOT_IDENTITY) ident.push_back(i);
    // Nothing is merged (merge switched off, skipped as too deep, or no identity instance) and the set of identity instances is the
    // one that decision was taken on: the roots go up as they are -- no rebuild, no stream synchronise, the frame graphs survive.
    if (ctx->merge_slot_root.empty() && identity_roots_sorted(ctx->host_roots) == ctx->merge_decided_roots) {
        std::vector<int> r = ctx->host_roots; upload_roots(ctx, r); return 0;
    }
    // Match the merged slots with the identity instances of the new table (TLAS leaf order may have changed).  A slot whose instance
    // is gone or no longer has an identity transform is RETIRED: its slot -> instance entry becomes -1, which the triangle tests
    // check before accepting a hit, and the instance is traced through the TLAS like any moving one.  No rebuild, no stream
    // synchronise (the round-1 code re-merged all 262 K triangles on the CPU, 0.65 s, whenever an identity instance started moving).
    // Instances that become identity later are simply not merged until ptb_set_static_merge() is called again.
    if (!ctx->merge_enabled) { ctx->merge_decided_roots = identity_roots_sorted(ctx->host_roots); std::vector<int> r = ctx->host_roots; upload_roots(ctx, r); return 0; }
    std::vector<int> slot_instance(ctx->merge_slot_root.size(), -1);
    std::vector<char> taken(ident.size(), 0);
    for (size_t k = 0; k < slot_instance.size(); k++) {
        if (ctx->merge_slot_instance[k] < 0) continue;                                               // retired earlier: stays retired
        size_t j = k < ident.size() && !taken[k] && roots[ident[k]] == ctx->merge_slot_root[k] ? k : ident.size();     // common case: order unchanged
        if (j == ident.size()) for (j = 0; j < ident.size(); j++) if (!taken[j] && roots[ident[j]] == ctx->merge_slot_root[k]) break;
        if (j == ident.size()) continue;                                                             // retire
        taken[j] = 1; slot_instance[k] = ident[j];
    }
    bool newly_retired = false;
    for (size_t k = 0; k < slot_instance.size(); k++) if (slot_instance[k] < 0 && ctx->merge_slot_instance[k] >= 0) newly_retired = true;
    ctx->merge_slot_instance = slot_instance;
    std::vector<int> device_roots = ctx->host_roots;
    for (int i : slot_instance) if (i >= 0) device_roots[i] = int((unsigned)device_roots[i] | PTB_ROOT_MERGED);
    upload_roots(ctx, device_roots);
    if (newly_retired && ctx->merge_tris && ctx->merge_ref_count > 0) {
        // the slot table must be on the device before the records are poisoned; still no host <-> device synchronisation
        int fe = staging_split(ctx); if (fe) return fe;
        k_retire_merged_slots<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->merge_tris, ctx->merge_woop, ctx->merge_who, ctx->merge_ref_count, ctx->merge_slot_instance_dev);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    return upload_pruned_tlas(ctx);
}

extern "C" int ptb_set_static_merge(ptb_ctx* ctx, int mode) {
    if (!ctx || mode < 0 || mode > 2) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    const bool enabled = mode != 0, spatial = mode != 2;
    if (enabled == ctx->merge_enabled && (!enabled || spatial == ctx->merge_spatial)) return 0;
    ctx->merge_enabled = enabled;
    if (enabled) ctx->merge_spatial = spatial;
    return ctx->has_scene ? rebuild_static_merge(ctx) : 0;
}

// Pathtracer::resize_free + resize_init (Pathtracer.cpp:255-314): new film size, same scene.  Accumulators, SVGF history and the
// frame graphs start over; the caller sets the camera for the new film (Camera::resize) and restarts sample_index at 0.
extern "C" int ptb_resize(ptb_ctx* ctx, int width, int height) {
    if (!ctx || width <= 0 || height <= 0) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->F.xchg.count > 0) return PTB_E_STATE;           // connected peers hold pointers into the old film: ptb_exchange_disconnect first
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->side_stream) CK(cudaStreamSynchronize(ctx->side_stream));
    drop_graphs(ctx);
    if (width == ctx->F.width && height == ctx->F.height) return 0;
    for (void*& m : ctx->xchg_ipc_opened) if (m) { cudaIpcCloseMemHandle(m); m = nullptr; }
    if (ctx->xchg_block) { cudaFree(ctx->xchg_block); ctx->xchg_block = nullptr; }
    for (void* p : ctx->wave_allocs) cudaFree(p);
    ctx->wave_allocs.clear();
    for (void* p : ctx->film_allocs) cudaFree(p);
    ctx->film_allocs.clear();
    Frame& F = ctx->F;
    for (int k = 0; k < PTB_AOV_COUNT; k++) { F.aov[k].fb = nullptr; F.aov[k].acc = nullptr; }
    F.display = nullptr; ctx->tap_hits = nullptr; ctx->present_buf = nullptr;
    { float vp[16], vpp[16]; memcpy(vp, F.svgf.view_projection, 64); memcpy(vpp, F.svgf.view_projection_prev, 64);
      memset(&F.svgf, 0, sizeof(F.svgf)); memcpy(F.svgf.view_projection, vp, 64); memcpy(F.svgf.view_projection_prev, vpp, 64); }
    F.width = width; F.height = height;
    int e = allocate_film(ctx); if (e) return e;
    ctx->frames_since_reset = 0; ctx->last_sample_index = -1;
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// The reference's second integrator (Src/Renderer/Integrators/AO.{h,cpp}, Src/CUDA/AO.cu): one cosine-distributed occlusion ray of
// length `ao_radius` per primary hit.  Same queues, traversal kernels, waves and accumulator as the path tracer.
extern "C" int ptb_set_integrator(ptb_ctx* ctx, int kind, float ao_radius) {
    if (!ctx || (kind != PTB_INTEGRATOR_PATHTRACER && kind != PTB_INTEGRATOR_AO) || !(ao_radius > 0.0f)) return PTB_E_BADARG;
    if (kind == PTB_INTEGRATOR_AO && ctx->F.config.enable_svgf) return PTB_E_STATE;
    if (kind != ctx->integrator || ao_radius != ctx->ao_radius) drop_graphs(ctx);
    ctx->integrator = kind; ctx->ao_radius = ao_radius;
    return 0;
}

extern "C" int ptb_set_intersector(ptb_ctx* ctx, int kind) {
    if (!ctx || (kind != PTB_INTERSECT_MT && kind != PTB_INTERSECT_WOOP)) return PTB_E_BADARG;
    if (kind == ctx->intersector) return 0;
    drop_graphs(ctx);
    ctx->intersector = kind;
    ctx->F.flat_woop = (kind == PTB_INTERSECT_WOOP && ctx->F.flat_root >= 0) ? ctx->merge_woop : nullptr;
    return 0;
}

extern "C" int ptb_upload_scene(ptb_ctx* ctx, const ptb_scene* s) {
    if (!ctx || !s) return PTB_E_BADARG;
    if (ctx->has_scene) return PTB_E_STATE;     // one scene per ctx (create a new ctx to switch scenes)
    if (!s->triangles || !s->bvh_nodes || s->mesh_count <= 0 || (s->bvh_kind != 8 && s->bvh_kind != 4 && s->bvh_kind != 2) || !s->pmj_samples || !s->blue_noise || !s->sky)
        return PTB_E_BADARG;
    // layout convention of the node array (Integrator.cpp:113,252-277): TLAS in slots [0, 2 * mesh_count), every BLAS root behind it --
    // ptb_update_instances overwrites the front of the array and relies on it
    if (s->tlas_node_count < 0 || s->tlas_node_count > 2 * s->mesh_count || s->bvh_node_count < 2 * s->mesh_count) return PTB_E_BADARG;
    for (int i = 0; i < s->mesh_count; i++) {
        unsigned root = (unsigned)s->mesh_bvh_root_indices[i] & 0x3fffffffu;
        if (root < 2u * (unsigned)s->mesh_count || root >= (unsigned)s->bvh_node_count) return PTB_E_BADARG;
    }
    CK(cudaSetDevice(ctx->device));
    { int de = load_driver_api(); if (de) return de; }
    Frame& F = ctx->F;
    int e = 0;
    e |= dev_upload<float4>(ctx, &F.triangles, s->triangles, (size_t)s->triangle_count * 6);
    ctx->bvh_kind = s->bvh_kind; ctx->node_count = s->bvh_node_count;
    if (s->bvh_kind == 8) { e |= dev_upload<float4>(ctx, &F.nodes8, s->bvh_nodes, (size_t)s->bvh_node_count * 5); ctx->uploaded_nodes = F.nodes8; }
    else if (s->bvh_kind == 4) e |= dev_upload<float4>(ctx, &F.nodes4, s->bvh_nodes, (size_t)s->bvh_node_count * 8);
    else                  e |= dev_upload<float4>(ctx, &F.nodes2, s->bvh_nodes, (size_t)s->bvh_node_count * 2);
    F.flat_root = -1;
    if (s->bvh_kind == 8) {
        const unsigned char* nb = static_cast<const unsigned char*>(s->bvh_nodes);
        ctx->host_nodes.assign(nb, nb + (size_t)s->bvh_node_count * 80);
        ctx->host_tri_pos.resize((size_t)s->triangle_count * 3);
        const float4* tr = static_cast<const float4*>(s->triangles);
        for (int t = 0; t < s->triangle_count; t++) for (int k = 0; k < 3; k++) ctx->host_tri_pos[(size_t)t * 3 + k] = tr[(size_t)t * 6 + k];
    }
    F.tlas_nodes = s->bvh_kind == 8 ? s->tlas_node_count : 0;
    ctx->mesh_capacity = s->mesh_count;
    e |= dev_upload<int>(ctx, &F.mesh_roots, s->mesh_bvh_root_indices, s->mesh_count);
    e |= dev_upload<int>(ctx, &F.mesh_material_ids, s->mesh_material_ids, s->mesh_count);
    e |= dev_upload<float4>(ctx, &F.mesh_transforms, s->mesh_transforms, (size_t)s->mesh_count * 3);
    e |= dev_upload<float4>(ctx, &F.mesh_transforms_inv, s->mesh_transforms_inv, (size_t)s->mesh_count * 3);
    e |= dev_upload<float4>(ctx, &F.mesh_transforms_prev, s->mesh_transforms_prev ? s->mesh_transforms_prev : s->mesh_transforms, (size_t)s->mesh_count * 3);
    e |= dev_upload<signed char>(ctx, &F.material_types, s->material_types, s->material_count);
    e |= dev_upload<float4>(ctx, &F.materials, s->materials, (size_t)s->material_count * 2);
    e |= dev_upload<float4>(ctx, &F.media, s->media, (size_t)(s->medium_count > 0 ? s->medium_count : 0) * 2);
    e |= dev_upload<float2>(ctx, &F.pmj, s->pmj_samples, (size_t)PTB_PMJ_SEQUENCES * PTB_PMJ_SAMPLES);
    e |= dev_upload<uchar2>(ctx, &F.blue_noise, s->blue_noise, (size_t)PTB_BLUE_NOISE_TEXTURES * PTB_BLUE_NOISE_DIM * PTB_BLUE_NOISE_DIM);
    if (e) return PTB_E_STATE;

    // which shade kernels run: scene.has_<type> is over ALL materials (Scene::check_materials, Scene.cpp)
    for (int i = 0; i < s->material_count; i++) {
        int t = s->material_types[i];
        if (t >= PTB_MAT_DIFFUSE && t <= PTB_MAT_CONDUCTOR) ctx->has_type[t - PTB_MAT_DIFFUSE] = true;
    }
    // lights
    F.lights_total_weight = s->light_mesh_count > 0 ? s->lights_total_weight : 0.0f;
    ctx->has_lights = s->light_mesh_count > 0 && s->lights_total_weight > 0.0f;
    F.light_mesh_count = s->light_mesh_count;
    e |= dev_upload<int>(ctx, &F.light_triangle_indices, s->light_triangle_indices, s->light_triangle_count);
    e |= dev_upload<float>(ctx, &F.light_triangle_cdf, s->light_triangle_cumulative_probability, s->light_triangle_count);
    e |= dev_upload<float>(ctx, &F.light_mesh_cdf, s->light_mesh_cumulative_probability, s->light_mesh_count);
    e |= dev_upload<int2>(ctx, &F.light_mesh_triangle_span, s->light_mesh_triangle_span, s->light_mesh_count);
    e |= dev_upload<int>(ctx, &F.light_mesh_transform_indices, s->light_mesh_transform_indices, s->light_mesh_count);
    if (e) return PTB_E_STATE;

    // textures
    std::vector<TextureEntry> table((size_t)(s->texture_count > 0 ? s->texture_count : 0));
    ctx->textures.resize(table.size());
    for (size_t i = 0; i < table.size(); i++) {
        int te = create_texture(ctx, s->textures[i], ctx->textures[i]); if (te) return te;
        table[i].tex = (cudaTextureObject_t)ctx->textures[i].tex; table[i].lod_bias = s->textures[i].lod_bias; table[i].pad = 0.0f;
    }
    e |= dev_upload<TextureEntry>(ctx, &F.textures, table.data(), table.size());
    CK(cudaStreamSynchronize(ctx->stream));   // `table` is a host temporary

    // sky: float4 array, bilinear, clamp (Integrator.cpp:285-296)
    e = create_float_texture(ctx, &ctx->sky_array, &F.sky_tex, s->sky, 4, s->sky_width, s->sky_height, 0, false); if (e) return e;
    F.sky_scale = s->sky_scale;

    // mesh-space box of every BLAS root (what a TLAS refit moves around): union of the root's child boxes
    for (int i = 0; i < s->mesh_count; i++) {
        unsigned root = (unsigned)s->mesh_bvh_root_indices[i] & 0x3fffffffu;
        if (ctx->blas_root_boxes.count(root)) continue;
        std::array<float, 6> b = { 1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f };
        if (s->bvh_kind == 8) {
            const unsigned char* n = static_cast<const unsigned char*>(s->bvh_nodes) + (size_t)root * 80;
            float p[3]; memcpy(p, n, 12);
            for (int k = 0; k < 8; k++) {
                if (!n[24 + k]) continue;
                for (int a = 0; a < 3; a++) {
                    unsigned eb = (unsigned)n[12 + a] << 23; float scale; memcpy(&scale, &eb, 4);
                    float lo = p[a] + scale * (float)n[32 + 16 * a + k], hi = p[a] + scale * (float)n[40 + 16 * a + k];
                    b[a] = lo < b[a] ? lo : b[a]; b[3 + a] = hi > b[3 + a] ? hi : b[3 + a];
                }
            }
        } else if (s->bvh_kind == 2) {
            memcpy(b.data(), static_cast<const unsigned char*>(s->bvh_nodes) + (size_t)root * 32, 24);
        }
        ctx->blas_root_boxes[root] = b;
    }
    ctx->has_scene = true;
    ctx->host_roots.assign(s->mesh_bvh_root_indices, s->mesh_bvh_root_indices + s->mesh_count);
    e = rebuild_static_merge(ctx); if (e) return e;
    if (ctx->has_type[2] || ctx->has_type[3]) { e = bake_luts(ctx); if (e) return e; }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int ptb_update_instances(ptb_ctx* ctx, const void* tlas_nodes, int tlas_node_count, int mesh_count,
                                    const int32_t* roots, const int32_t* material_ids, const float* xf, const float* xf_inv, const float* xf_prev) {
    if (!ctx || !ctx->has_scene) return PTB_E_NOSCENE;
    if (mesh_count != ctx->mesh_capacity || tlas_node_count > 2 * mesh_count || !tlas_nodes) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->bvh_kind == 8 && ctx->F.tlas_nodes != tlas_node_count) drop_graphs(ctx);
    Frame& F = ctx->F;
    size_t node_bytes = ctx->bvh_kind == 8 ? 80 : ctx->bvh_kind == 4 ? 128 : 32;
    void* dst = ctx->bvh_kind == 8 ? (void*)F.nodes8 : ctx->bvh_kind == 4 ? (void*)F.nodes4 : (void*)F.nodes2;
    { int se = staging_begin(ctx, 3 * node_bytes * (size_t)tlas_node_count + (size_t)mesh_count * 192 + 4096); if (se) return se; }
    { int se = stage_upload(ctx, dst, tlas_nodes, node_bytes * tlas_node_count); if (se) return se; }
    if (ctx->bvh_kind == 8 && ctx->uploaded_nodes != F.nodes8)      // keep the un-merged copy current as well
        { int se = stage_upload(ctx, (void*)ctx->uploaded_nodes, tlas_nodes, node_bytes * tlas_node_count); if (se) return se; }
    if (ctx->bvh_kind == 8) {
        F.tlas_nodes = tlas_node_count;
        memcpy(ctx->host_nodes.data(), tlas_nodes, node_bytes * tlas_node_count);
    }
    if (roots) { int re = apply_roots(ctx, roots, mesh_count); if (re) return re; }
    else       { int re = upload_pruned_tlas(ctx); if (re) return re; }
    int se = 0;
    if (material_ids) se |= stage_upload(ctx, (void*)F.mesh_material_ids, material_ids, sizeof(int) * mesh_count);
    if (xf) se |= stage_upload(ctx, (void*)F.mesh_transforms, xf, 48 * (size_t)mesh_count);
    if (xf_inv) se |= stage_upload(ctx, (void*)F.mesh_transforms_inv, xf_inv, 48 * (size_t)mesh_count);
    if (xf_prev) se |= stage_upload(ctx, (void*)F.mesh_transforms_prev, xf_prev, 48 * (size_t)mesh_count);
    if (se) return se;
    return staging_end(ctx);        // asynchronous: the copies run behind whatever the stream is still rendering
}

// TLAS refit on the device (ptb_refit.cuh): new transforms, same topology, no host BVH work and no host <-> device synchronisation.
static bool is_identity_3x4(const float* m) {
    static const float I[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
    return memcmp(m, I, sizeof(I)) == 0;
}
extern "C" int ptb_refit_instances(ptb_ctx* ctx, const float* xf, const float* xf_inv, const float* xf_prev) {
    if (!ctx || !ctx->has_scene) return PTB_E_NOSCENE;
    if (!xf || !xf_inv) return PTB_E_BADARG;
    if (ctx->bvh_kind == 4) return PTB_E_STATE;          // refit covers the CWBVH and the binary TLAS; a 4-wide TLAS is rebuilt on the host (ptb_update_instances)
    CK(cudaSetDevice(ctx->device));
    Frame& F = ctx->F;
    const int M = ctx->mesh_capacity;
    const int n_tlas = ctx->bvh_kind == 8 ? F.tlas_nodes : (2 * M < ctx->node_count ? 2 * M : ctx->node_count);
    if (M <= 0 || n_tlas <= 0) return PTB_E_STATE;
    // an instance that leaves (or returns to) the identity changes how it is traced: the ray transform is skipped for identity
    // instances (root bit 31), and a moving instance must leave the merged static BVH (its slot is retired, apply_roots)
    std::vector<int> roots = ctx->host_roots;
    bool flags_changed = false;
    std::vector<float> local((size_t)M * 6);
    for (int i = 0; i < M; i++) {
        unsigned r = (unsigned)roots[i], nr = is_identity_3x4(xf + 12 * (size_t)i) ? (r | PTB_ROOT_IDENTITY) : (r & ~PTB_ROOT_IDENTITY);
        if (nr != r) { roots[i] = (int)nr; flags_changed = true; }
        auto it = ctx->blas_root_boxes.find(r & 0x3fffffffu);
        if (it == ctx->blas_root_boxes.end()) return PTB_E_STATE;
        memcpy(&local[(size_t)i * 6], it->second.data(), 24);
    }
    const size_t node_slots = (size_t)(2 * M > n_tlas ? 2 * M : n_tlas);
    if (!ctx->refit_scratch) CK(cudaMalloc(reinterpret_cast<void**>(&ctx->refit_scratch), ((size_t)M * 12 + node_slots * 7) * sizeof(float)));
    float* d_local = ctx->refit_scratch; float* d_inst = d_local + (size_t)M * 6; float* d_node = d_inst + (size_t)M * 6;
    int* d_done = reinterpret_cast<int*>(d_node + node_slots * 6);
    // Integrator::update keeps last frame's transform for the motion vectors (Mesh::transform_prev).  Enqueued before anything is staged:
    // the staged uploads of this call land with ONE flush at the end (a flush in the middle of a call would let the host overwrite the
    // segment table of the pinned arena while its asynchronous copy may still be pending)
    if (!xf_prev) CK(cudaMemcpyAsync((void*)F.mesh_transforms_prev, F.mesh_transforms, 48 * (size_t)M, cudaMemcpyDeviceToDevice, ctx->stream));
    { int se = staging_begin(ctx, (size_t)M * (3 * 48 + 24 + 8) + (size_t)n_tlas * 80 + 8192); if (se) return se; }
    if (flags_changed) { int re = apply_roots(ctx, roots.data(), M); if (re) return re; }
    int se = stage_upload(ctx, (void*)F.mesh_transforms, xf, 48 * (size_t)M);
    se |= stage_upload(ctx, (void*)F.mesh_transforms_inv, xf_inv, 48 * (size_t)M);
    if (xf_prev) se |= stage_upload(ctx, (void*)F.mesh_transforms_prev, xf_prev, 48 * (size_t)M);
    se |= stage_upload(ctx, d_local, local.data(), 24 * (size_t)M);
    if (se) return se;
    { int fe = staging_end(ctx); if (fe) return fe; }
    RefitArgs A;
    A.node_count = n_tlas; A.kind = ctx->bvh_kind; A.instance_count = M; A.local_boxes = d_local; A.transforms = F.mesh_transforms;
    A.instance_boxes = d_inst; A.node_boxes = d_node; A.done = d_done;
    // every copy of the TLAS a ray may walk: the array in use (the pruned copy next to the merged BVH, or the uploaded one) and, when
    // they differ, the uploaded array the strict mode falls back to
    float4* in_use = const_cast<float4*>(ctx->bvh_kind == 8 ? F.nodes8 : F.nodes2);
    A.nodes = in_use;
    k_refit_tlas<<<1, 1024, 0, ctx->stream>>>(A); ctx->launches++;
    if (ctx->bvh_kind == 8 && ctx->uploaded_nodes && ctx->uploaded_nodes != F.nodes8) {
        A.nodes = const_cast<float4*>(ctx->uploaded_nodes);
        k_refit_tlas<<<1, 1024, 0, ctx->stream>>>(A); ctx->launches++;
    }
    CK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- render
struct StageTimer {
    ptb_ctx* ctx; int stage; cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t take(ptb_ctx* c) {
        if (c->event_used == c->event_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); c->event_pool.push_back(e); }
        return c->event_pool[c->event_used++];
    }
    StageTimer(ptb_ctx* c, int s) : ctx(c), stage(s) { if (c->timing) { a = take(c); b = take(c); cudaEventRecord(a, c->stream); } }
    ~StageTimer() { if (ctx->timing) { cudaEventRecord(b, ctx->stream); ctx->timed.push_back({ stage, { a, b } }); } }
};

static size_t trace8_smem() { return 16 + (size_t)PTB_TLAS_STAGE_MAX_NODES * 80 + (size_t)PTB_SM_STACK * PTB_TRACE_BLOCK * sizeof(uint2); }

template <bool SHADOW>
static void launch_trace8(ptb_ctx* ctx, const Frame& F, int grid, cudaStream_t st, int bounce, const unsigned* order) {
    if (ctx->stats_mode) k_trace8<SHADOW, true><<<grid, PTB_TRACE_BLOCK, trace8_smem(), st>>>(F, bounce, order);
    else                 k_trace8<SHADOW, false><<<grid, PTB_TRACE_BLOCK, trace8_smem(), st>>>(F, bounce, order);
}

// One wave: `samples` consecutive passes (first_sample ...) through the whole pipeline.  The role of one or several
// Pathtracer::render() calls (Pathtracer.cpp:738-855); asynchronous, no host<->device synchronisation.
static int render_wave(ptb_ctx* ctx, int first_sample, int samples, bool push = false) {
    if (samples < 1 || samples > ctx->wave_capacity) return PTB_E_BADARG;
    Frame F = ctx->F;
    F.first_sample = first_sample; F.wave_samples = samples;
    F.xchg.push = push && F.xchg.count > 0 && !F.config.enable_svgf;
    F.xchg.svgf = F.xchg.count > 0 && F.config.enable_svgf;
    if (F.config.enable_svgf) {
        SVGFBuffers& V = F.svgf;
        V.parity = int(ctx->svgf_frames & 1u);
        if (F.xchg.svgf) {
            // tile-local filter: this rank filters one block of rows; its inputs (stored by the tracing ranks) and the history it owns
            // live in its exchange block, where the neighbours can reach them
            const int S = F.fb_stride;
            F.xchg.rows_per_block = (F.height + F.world - 1) / F.world;
            V.block_y0 = F.rank * F.xchg.rows_per_block < F.height ? F.rank * F.xchg.rows_per_block : F.height;
            V.block_y1 = F.rank == F.world - 1 ? F.height : ((F.rank + 1) * F.xchg.rows_per_block < F.height ? (F.rank + 1) * F.xchg.rows_per_block : F.height);
            V.ext_y0 = V.block_y0 - PTB_SVGF_HALO > 0 ? V.block_y0 - PTB_SVGF_HALO : 0;
            V.ext_y1 = V.block_y1 + PTB_SVGF_HALO < F.height ? V.block_y1 + PTB_SVGF_HALO : F.height;
            float4* in = F.xchg.frames[F.rank] + xchg_input_offset(S, V.parity);
            V.in_direct = in; V.in_indirect = in + (size_t)S; V.in_albedo = in + (size_t)S * 2; V.in_normal_depth = in + (size_t)S * 3;
            V.in_ids = reinterpret_cast<int2*>(in + (size_t)S * 4); V.in_screen_prev = reinterpret_cast<float2*>(in + (size_t)S * 4 + (size_t)S / 2);
            for (int p = 0; p < 2; p++) {
                float4* h = F.xchg.frames[F.rank] + xchg_history_offset(S, p);
                V.hist[p].direct = h; V.hist[p].indirect = h + (size_t)S; V.hist[p].moment = h + (size_t)S * 2; V.hist[p].normal_depth = h + (size_t)S * 3;
                V.hist[p].taa = h + (size_t)S * 4; V.hist[p].length = reinterpret_cast<int*>(h + (size_t)S * 5);
            }
        } else {
            V.block_y0 = V.ext_y0 = 0; V.block_y1 = V.ext_y1 = F.height;
            V.in_direct = F.aov[PTB_AOV_RADIANCE_DIRECT].fb; V.in_indirect = F.aov[PTB_AOV_RADIANCE_INDIRECT].fb; V.in_albedo = F.aov[PTB_AOV_ALBEDO].fb;
            V.in_normal_depth = V.gbuf_normal_depth; V.in_ids = V.gbuf_ids; V.in_screen_prev = V.gbuf_screen_prev;
            V.hist[0] = ctx->local_hist[0]; V.hist[1] = ctx->local_hist[1];
        }
        ctx->svgf_frames++;
    }
    if (F.config.enable_svgf && samples != 1) return PTB_E_STATE;   // SVGF is temporal: one pass per displayed frame
    cudaStream_t st = ctx->stream;
    const int g1d = grid_for(ctx, 8);
    const int gtrace = grid_for(ctx, PTB_TRACE_MIN_BLOCKS);
    const bool nee = ctx->has_lights && F.config.enable_next_event_estimation;

    k_begin_pass<<<1, 256, 0, st>>>(F); ctx->launches++;
    { StageTimer t(ctx, ST_GENERATE); k_generate<<<g1d, 256, 0, st>>>(F); ctx->launches++; }
    // Dependencies inside a bounce: trace -> sort -> shade -> { shadow trace, next bounce's trace }.  The two traces only meet
    // again at the next sort (both deposit into the framebuffer: shadow first, to keep the reference's summation order), so the
    // shadow trace runs on a side stream; its CTAs and the next closest-hit trace's CTAs fill each other's ramp-down tails.
    // Per-stage timing (events on the main stream) and the ordering experiment keep everything on one stream.
    const bool overlap = ctx->overlap_enabled && nee && !ctx->timing && !ctx->stats_mode && F.order_bins == 0 && ctx->integrator != PTB_INTEGRATOR_AO;
    bool side_pending = false;
    const bool ao = ctx->integrator == PTB_INTEGRATOR_AO;
    for (int bounce = 0; bounce < (ao ? 1 : F.config.num_bounces); bounce++) {
        const bool ordered = F.order_bins > 0 && ctx->bvh_kind == 8 && bounce < PTB_ORDER_MAX_BOUNCE;
        const unsigned* order_c = ordered && bounce > 0 ? F.order : nullptr;       // primary rays are coherent as generated
        const unsigned* order_s = ordered ? F.order : nullptr;
        if (order_c) { StageTimer t(ctx, ST_ORDER);
          k_bin_count<false><<<g1d, 256, 0, st>>>(F, bounce); k_bin_scatter<false><<<g1d, 256, 0, st>>>(F, bounce); ctx->launches += 2; }
        { StageTimer t(ctx, ST_TRACE);
          if (ctx->bvh_kind == 8) { launch_trace8<false>(ctx, F, gtrace, st, bounce, order_c); }
          else if (ctx->bvh_kind == 4) { if (ctx->stats_mode) k_trace4<false, true><<<gtrace, PTB_TRACE_BLOCK, 0, st>>>(F, bounce); else k_trace4<false, false><<<gtrace, PTB_TRACE_BLOCK, 0, st>>>(F, bounce); }
          else if (ctx->stats_mode) k_trace2<false, true><<<gtrace, PTB_TRACE_BLOCK, 0, st>>>(F, bounce);
          else                    k_trace2<false, false><<<gtrace, PTB_TRACE_BLOCK, 0, st>>>(F, bounce);
          ctx->launches++; }
        if (side_pending) { CK(cudaStreamWaitEvent(st, ctx->ev_join, 0)); side_pending = false; }     // shadow[bounce-1] deposits before sort[bounce]
        if (ao) { StageTimer t(ctx, ST_SHADE); k_ambient_occlusion<<<g1d, 256, 0, st>>>(F, ctx->ao_radius); ctx->launches++; }
        else { StageTimer t(ctx, ST_SORT); k_sort<<<g1d, 256, 0, st>>>(F, bounce); ctx->launches++; }
        if (!ao) { StageTimer t(ctx, ST_SHADE);
          if (ctx->has_type[0]) { k_shade<BSDFDiffuse><<<g1d, 256, 0, st>>>(F, bounce); ctx->launches++; }
          if (ctx->has_type[1]) { k_shade<BSDFPlastic><<<g1d, 256, 0, st>>>(F, bounce); ctx->launches++; }
          if (ctx->has_type[2]) { k_shade<BSDFDielectric><<<g1d, 256, 0, st>>>(F, bounce); ctx->launches++; }
          if (ctx->has_type[3]) { k_shade<BSDFConductor><<<g1d, 256, 0, st>>>(F, bounce); ctx->launches++; } }
        if (nee && order_s) { StageTimer t(ctx, ST_ORDER);
          k_bin_count<true><<<g1d, 256, 0, st>>>(F, bounce); k_bin_scatter<true><<<g1d, 256, 0, st>>>(F, bounce); ctx->launches += 2; }
        if (nee || ao) {
            StageTimer t(ctx, ST_SHADOW);
            cudaStream_t ss = st;
            if (overlap) { CK(cudaEventRecord(ctx->ev_fork, st)); CK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0)); ss = ctx->side_stream; }
            if (ctx->bvh_kind == 8) { launch_trace8<true>(ctx, F, gtrace, ss, bounce, order_s); }
            else if (ctx->bvh_kind == 4) { if (ctx->stats_mode) k_trace4<true, true><<<gtrace, PTB_TRACE_BLOCK, 0, ss>>>(F, bounce); else k_trace4<true, false><<<gtrace, PTB_TRACE_BLOCK, 0, ss>>>(F, bounce); }
            else if (ctx->stats_mode) k_trace2<true, true><<<gtrace, PTB_TRACE_BLOCK, 0, ss>>>(F, bounce);
            else                    k_trace2<true, false><<<gtrace, PTB_TRACE_BLOCK, 0, ss>>>(F, bounce);
            ctx->launches++;
            if (overlap) { CK(cudaEventRecord(ctx->ev_join, ctx->side_stream)); side_pending = true; }
        }
    }
    if (side_pending) { CK(cudaStreamWaitEvent(st, ctx->ev_join, 0)); side_pending = false; }
    { StageTimer t(ctx, ST_POST);
      if (F.config.enable_svgf) {
          int e = launch_svgf(F, st, first_sample, g1d, &ctx->launches); if (e) return e;
          if (F.xchg.svgf) ctx->xchg_frames++;
      } else {
          k_accumulate<<<g1d, 256, 0, st>>>(F); ctx->launches++;
          if (F.xchg.push) { k_exchange_wait<<<1, 1, 0, st>>>(F); ctx->launches++; }
      } }
    k_fold_counters<<<1, PTB_MAX_BOUNCES, 0, st>>>(F); ctx->launches++;
    if (!ctx->capturing) CK(cudaGetLastError());
    ctx->last_sample_index = first_sample + samples - 1;
    ctx->frames_since_reset += samples;
    return 0;
}

extern "C" int ptb_render(ptb_ctx* ctx, int sample_index) {
    if (!ctx) return PTB_E_BADARG;
    if (!ctx->has_scene) return PTB_E_NOSCENE;
    CK(cudaSetDevice(ctx->device));
    if (ctx->timing) { ctx->timed.clear(); ctx->event_used = 0; }
    return render_wave(ctx, sample_index, 1);
}

extern "C" int ptb_set_ray_ordering(ptb_ctx* ctx, int bins) {
    if (!ctx || !(bins == 0 || bins == 8 || bins == 64)) return PTB_E_BADARG;
    if (bins != ctx->F.order_bins) {
        drop_graphs(ctx); ctx->F.order_bins = bins;
        if (bins > 0 && !ctx->F.order) return allocate_wave_storage(ctx, ctx->wave_capacity);
    }
    return 0;
}

extern "C" int ptb_reserve_wave(ptb_ctx* ctx, int samples) {
    if (!ctx) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    if (samples == ctx->wave_capacity) return 0;
    drop_graphs(ctx);
    return allocate_wave_storage(ctx, samples);
}

// passes [first, first + n) as waves of up to wave_capacity passes each
static int render_passes(ptb_ctx* ctx, int first, int n) {
    int cap = ctx->F.config.enable_svgf ? 1 : ctx->wave_capacity;
    for (int done = 0; done < n;) {
        int s = n - done < cap ? n - done : cap;
        int e = render_wave(ctx, first + done, s, /*push the finished frame to the peers*/ done + s == n); if (e) return e;
        done += s;
    }
    return 0;
}

extern "C" int ptb_render_frame(ptb_ctx* ctx, int first_sample_index, int num_passes) {
    if (!ctx || num_passes <= 0) return PTB_E_BADARG;
    if (!ctx->has_scene) return PTB_E_NOSCENE;
    CK(cudaSetDevice(ctx->device));
    if (ctx->timing || ctx->stats_mode || ctx->F.config.enable_svgf) {   // per-stage events are not captured into graphs; SVGF frames alternate history parity
        if (ctx->timing) { ctx->timed.clear(); ctx->event_used = 0; }
        int e = render_passes(ctx, first_sample_index, num_passes);
        if (!e && ctx->F.xchg.count > 0 && !ctx->F.config.enable_svgf) ctx->xchg_frames++;
        return e;
    }
    for (auto& g : ctx->graphs) if (g.first == first_sample_index && g.passes == num_passes) {
        CK(cudaGraphLaunch(g.exec, ctx->stream));
        if (ctx->F.xchg.count > 0 && !ctx->F.config.enable_svgf) ctx->xchg_frames++;
        ctx->launches += g.launches;
        ctx->last_sample_index = first_sample_index + num_passes - 1;
        ctx->frames_since_reset += num_passes;
        return 0;
    }
    long long before = ctx->launches;
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    int e = render_passes(ctx, first_sample_index, num_passes);
    ctx->capturing = false;
    cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    if (e) { if (graph) cudaGraphDestroy(graph); return e; }
    CK(ce);
    long long graph_launches = ctx->launches - before;
    ctx->launches = before;                           // capture itself executed nothing
    cudaGraphExec_t exec = nullptr;
    CK(cudaGraphInstantiate(&exec, graph, 0));
    cudaGraphDestroy(graph);
    if (ctx->graphs.size() >= 8) { cudaGraphExecDestroy(ctx->graphs.front().exec); ctx->graphs.erase(ctx->graphs.begin()); }
    ctx->graphs.push_back({ first_sample_index, num_passes, exec, graph_launches });
    return ptb_render_frame(ctx, first_sample_index, num_passes);
}

// ---------------------------------------------------------------------------------------------- frame exchange (peer memory)
extern "C" int ptb_exchange_create(ptb_ctx* ctx, void** local_base, void* ipc_handle_out) {
    if (!ctx) return PTB_E_BADARG;
    static_assert(sizeof(ExchangeControl) <= PTB_XCHG_HEADER, "control block");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size promised by ptb.h");
    CK(cudaSetDevice(ctx->device));
    if (ctx->F.world > PTB_MAX_PEERS) return PTB_E_BADARG;
    if (!ctx->xchg_block) {
        size_t bytes = PTB_XCHG_HEADER + PTB_XCHG_BLOCK_FLOAT4(ctx->F.fb_stride) * sizeof(float4);   // control | 2 frames | SVGF input planes
        CK(cudaMalloc(&ctx->xchg_block, bytes));
        CK(cudaMemset(ctx->xchg_block, 0, bytes));
    }
    if (local_base) *local_base = ctx->xchg_block;
    if (ipc_handle_out) CK(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(ipc_handle_out), ctx->xchg_block));
    return 0;
}

static int exchange_connect(ptb_ctx* ctx, void* const* bases) {
    Frame& F = ctx->F;
    drop_graphs(ctx);
    for (int r = 0; r < F.world; r++) {
        char* base = static_cast<char*>(r == F.rank ? ctx->xchg_block : bases[r]);
        if (!base) return PTB_E_BADARG;
        F.xchg.control[r] = reinterpret_cast<ExchangeControl*>(base);
        F.xchg.frames[r] = reinterpret_cast<float4*>(base + PTB_XCHG_HEADER);
    }
    F.xchg.count = F.world;
    F.xchg.push = 0;
    ctx->xchg_frames = 0;
    CK(cudaMemset(ctx->xchg_block, 0, PTB_XCHG_HEADER));
    return 0;
}

extern "C" int ptb_exchange_connect(ptb_ctx* ctx, void* const* peer_bases) {
    if (!ctx || !peer_bases || !ctx->xchg_block) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    for (int r = 0; r < ctx->F.world; r++) {                // peers of the same process on other devices: enable peer access
        if (r == ctx->F.rank || !peer_bases[r]) continue;
        cudaPointerAttributes at;
        CK(cudaPointerGetAttributes(&at, peer_bases[r]));
        if (at.device != ctx->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ctx_fail(ctx, "cudaDeviceEnablePeerAccess", (int)e); return (int)e; }
            cudaGetLastError();
        }
    }
    return exchange_connect(ctx, peer_bases);
}

extern "C" int ptb_exchange_connect_ipc(ptb_ctx* ctx, const void* ipc_handles) {
    if (!ctx || !ipc_handles || !ctx->xchg_block) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    void* bases[PTB_MAX_PEERS] = {};
    const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(ipc_handles);
    for (int r = 0; r < ctx->F.world; r++) {
        if (r == ctx->F.rank) continue;
        if (!ctx->xchg_ipc_opened[r]) {
            cudaIpcMemHandle_t hh; memcpy(&hh, &h[r], sizeof(hh));
            CK(cudaIpcOpenMemHandle(&ctx->xchg_ipc_opened[r], hh, cudaIpcMemLazyEnablePeerAccess));
        }
        bases[r] = ctx->xchg_ipc_opened[r];
    }
    return exchange_connect(ctx, bases);
}

extern "C" int ptb_exchange_disconnect(ptb_ctx* ctx) {
    if (!ctx) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    drop_graphs(ctx);
    ctx->F.xchg = Exchange{};
    ctx->xchg_frames = 0;
    return 0;
}

extern "C" int ptb_exchange_frame(ptb_ctx* ctx, void** device_ptr, int* pitch) {
    if (!ctx || !device_ptr) return PTB_E_BADARG;
    if (ctx->F.xchg.count == 0) return PTB_E_STATE;
    if (ctx->xchg_frames == 0) return PTB_E_STATE;
    *device_ptr = static_cast<char*>(ctx->xchg_block) + PTB_XCHG_HEADER + (size_t)((ctx->xchg_frames - 1) & 1u) * ctx->F.fb_stride * sizeof(float4);
    if (pitch) *pitch = ctx->F.pitch;
    return 0;
}

extern "C" int ptb_measure_traversal(ptb_ctx* ctx, int sample_index, ptb_traversal_stats* out) {
    if (!ctx || !out) return PTB_E_BADARG;
    if (!ctx->has_scene) return PTB_E_NOSCENE;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->F.trace_stats, 0, 2 * sizeof(TraceStats), ctx->stream));
    ctx->stats_mode = true;
    int e = ptb_render(ctx, sample_index);
    ctx->stats_mode = false;
    if (e) return e;
    TraceStats h[2];
    CK(cudaMemcpyAsync(h, ctx->F.trace_stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 2; k++) { out->rays[k] = h[k].rays; out->nodes[k] = h[k].nodes; out->triangles[k] = h[k].triangles; out->instance_transforms[k] = h[k].instance_transforms; }
    out->shadow_misses = h[1].misses;
    return 0;
}

extern "C" int ptb_sync(ptb_ctx* ctx) {
    if (!ctx) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->F.xchg.count > 0) {
        ExchangeControl c;
        CK(cudaMemcpy(&c, ctx->xchg_block, sizeof(c), cudaMemcpyDeviceToHost));
        if (c.status) { ctx->last_error = "frame exchange: a peer did not deliver its rows within 4 s"; fprintf(stderr, "[ptb] %s\n", ctx->last_error.c_str()); return PTB_E_EXCHANGE; }
    }
    if (ctx->timing && !ctx->timed.empty()) {
        for (int i = 0; i < ST_COUNT; i++) ctx->stage_ms[i] = 0.0f;
        for (auto& t : ctx->timed) { float ms = 0.0f; cudaEventElapsedTime(&ms, t.second.first, t.second.second); ctx->stage_ms[t.first] += ms; }
        ctx->timed.clear();
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------- readback
extern "C" int ptb_get_aov(ptb_ctx* ctx, int aov_type, int accumulated, void** device_ptr, int* pitch) {
    if (!ctx || aov_type < 0 || aov_type >= PTB_AOV_COUNT || !device_ptr) return PTB_E_BADARG;
    float4* p = accumulated ? ctx->F.aov[aov_type].acc : ctx->F.aov[aov_type].fb;
    if (!p) return PTB_E_STATE;
    *device_ptr = p;
    if (pitch) *pitch = ctx->F.pitch;
    return 0;
}
extern "C" int ptb_get_display(ptb_ctx* ctx, void** device_ptr, int* pitch) {
    if (!ctx || !device_ptr) return PTB_E_BADARG;
    *device_ptr = ctx->F.display;
    if (pitch) *pitch = ctx->F.pitch;
    return 0;
}
extern "C" int ptb_download(ptb_ctx* ctx, int aov_type, int accumulated, float* host_dst) {
    if (!ctx || !host_dst) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    void* src = nullptr;
    if (aov_type < 0) src = ctx->F.display;
    else { int e = ptb_get_aov(ctx, aov_type, accumulated, &src, nullptr); if (e) return e; }
    CK(cudaMemcpyAsync(host_dst, src, (size_t)ctx->F.pitch * ctx->F.height * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
// What the window does with a frame before showing or capturing it as LDR (Src/Shaders/post.frag, Src/Main.cpp:195-225): tone-map the
// displayed image (or, with several ranks, the gathered frame) to 8-bit RGBA on the device; optional copy to host (pitch x height x 4 B).
extern "C" int ptb_present(ptb_ctx* ctx, void** device_rgba8, void* host_dst) {
    if (!ctx) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    const size_t pixels = (size_t)ctx->F.pitch * ctx->F.height;
    if (!ctx->present_buf) { int e = film_alloc(ctx, &ctx->present_buf, pixels); if (e) return e; }
    const float4* src = ctx->F.display;
    if (ctx->F.xchg.count > 0 && ctx->xchg_frames > 0) { void* p = nullptr; int e = ptb_exchange_frame(ctx, &p, nullptr); if (e) return e; src = static_cast<const float4*>(p); }
    k_present<<<grid_for(ctx, 8), 256, 0, ctx->stream>>>(ctx->F, src, ctx->present_buf); ctx->launches++;
    CK(cudaGetLastError());
    if (device_rgba8) *device_rgba8 = ctx->present_buf;
    if (host_dst) {
        CK(cudaMemcpyAsync(host_dst, ctx->present_buf, pixels * sizeof(uchar4), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

// set_pixel_query (Integrator.h:266-277) / the read-back in Integrator::update (Integrator.cpp:483-494)
extern "C" int ptb_set_pixel_query(ptb_ctx* ctx, int x, int y) {
    if (!ctx || x < 0 || y < 0 || x >= ctx->F.width || y >= ctx->F.height) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    ctx->pixel_query_host[0] = x + y * ctx->F.pitch; ctx->pixel_query_host[1] = -1; ctx->pixel_query_host[2] = -1;
    CK(cudaMemcpyAsync(ctx->F.pixel_query, ctx->pixel_query_host, 3 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
extern "C" int ptb_get_pixel_query(ptb_ctx* ctx, int* mesh_id, int* triangle_id) {
    if (!ctx || !mesh_id || !triangle_id) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    int q[3];
    CK(cudaMemcpyAsync(q, ctx->F.pixel_query, sizeof(q), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *mesh_id = q[1]; *triangle_id = q[2];
    CK(cudaMemsetAsync(ctx->F.pixel_query, 0xff, 3 * sizeof(int), ctx->stream));  // query consumed: pixel_index back to INVALID
    return 0;
}

extern "C" int ptb_get_ray_stats(ptb_ctx* ctx, ptb_ray_stats* out, int reset) {
    if (!ctx || !out) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    static_assert(sizeof(RayTotals) == sizeof(ptb_ray_stats), "ray stats layout");
    CK(cudaMemcpyAsync(out, ctx->F.totals, sizeof(RayTotals), cudaMemcpyDeviceToHost, ctx->stream));
    if (reset) CK(cudaMemsetAsync(ctx->F.totals, 0, sizeof(RayTotals), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int ptb_get_stream(ptb_ctx* ctx, void** stream) {
    if (!ctx || !stream) return PTB_E_BADARG;
    *stream = ctx->stream;
    return 0;
}
extern "C" int ptb_export_rows(ptb_ctx* ctx, int aov_type, void* device_dst, int* owned_rows) {
    if (!ctx || !device_dst) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    const float4* src = aov_type < 0 ? ctx->F.display : ctx->F.aov[aov_type].acc;
    if (!src) return PTB_E_STATE;
    k_export_rows<<<grid_for(ctx, 8), 256, 0, ctx->stream>>>(ctx->F, src, static_cast<float4*>(device_dst)); ctx->launches++;
    if (owned_rows) *owned_rows = ctx->owned_rows;
    CK(cudaGetLastError());
    return 0;
}
extern "C" int ptb_assemble_rows(ptb_ctx* ctx, const void* device_src, int max_rows, void* device_dst) {
    if (!ctx || !device_src || !device_dst) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    k_assemble_rows<<<grid_for(ctx, 8), 256, 0, ctx->stream>>>(ctx->F, static_cast<const float4*>(device_src), max_rows, static_cast<float4*>(device_dst)); ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
extern "C" int ptb_debug_read(ptb_ctx* ctx, int which, void* host_dst, int64_t bytes) {
    if (!ctx || !host_dst) return PTB_E_BADARG;
    CK(cudaSetDevice(ctx->device));
    if (which == 0) {
        size_t need = (size_t)ctx->F.pitch * ctx->F.height * sizeof(uint4);
        if ((size_t)bytes < need) return PTB_E_BADARG;
        CK(cudaMemsetAsync(ctx->tap_hits, 0xff, need, ctx->stream));
        k_tap_primary_hits<<<grid_for(ctx, 8), 256, 0, ctx->stream>>>(ctx->F, ctx->tap_hits); ctx->launches++;
        CK(cudaMemcpyAsync(host_dst, ctx->tap_hits, need, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    if (which == 1) {   // per-bounce counters of the last pass
        if ((size_t)bytes < sizeof(Counters)) return PTB_E_BADARG;
        CK(cudaMemcpyAsync(host_dst, ctx->F.counters, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    if (which >= 10 && which <= 17) {   // SVGF state (pitch x height elements) as the last filtered frame left it; with several ranks only
                                        // the rows of this rank's filter block are meaningful
        const Frame& F = ctx->F;
        if (!F.svgf.moment || ctx->svgf_frames == 0) return PTB_E_STATE;
        const int parity = int((ctx->svgf_frames - 1u) & 1u);
        SVGFHistory h = ctx->local_hist[parity];
        if (F.xchg.count > 0) {
            float4* b = F.xchg.frames[F.rank] + xchg_history_offset(F.fb_stride, parity); const size_t S = (size_t)F.fb_stride;
            h.direct = b; h.indirect = b + S; h.moment = b + 2 * S; h.normal_depth = b + 3 * S; h.taa = b + 4 * S; h.length = reinterpret_cast<int*>(b + 5 * S);
        }
        const void* src[8] = { h.normal_depth, h.direct, h.indirect, h.moment, F.svgf.moment, F.svgf.taa_curr, h.taa, h.length };
        size_t elem = which == 17 ? sizeof(int) : sizeof(float4);
        size_t need = (size_t)F.pitch * F.height * elem;
        if (!src[which - 10] || (size_t)bytes < need) return PTB_E_BADARG;
        CK(cudaMemcpyAsync(host_dst, src[which - 10], need, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    if (which == 3) {   // the TLAS the rays walk (front of the node array in use), e.g. after ptb_refit_instances
        const size_t node_bytes = ctx->bvh_kind == 8 ? 80 : ctx->bvh_kind == 4 ? 128 : 32;
        const int n = ctx->bvh_kind == 8 ? ctx->F.tlas_nodes : (2 * ctx->mesh_capacity < ctx->node_count ? 2 * ctx->mesh_capacity : ctx->node_count);
        const void* src = ctx->bvh_kind == 8 ? (const void*)ctx->F.nodes8 : ctx->bvh_kind == 4 ? (const void*)ctx->F.nodes4 : (const void*)ctx->F.nodes2;
        if (!src || (size_t)bytes < node_bytes * (size_t)n) return PTB_E_BADARG;
        CK(cudaMemcpyAsync(host_dst, src, node_bytes * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    if (which == 2) {   // Kulla-Conty LUT contents
        const size_t n = 2 * 16 * 16 * 16 + 2 * 16 * 16 + 32 * 32 + 32;
        if ((size_t)bytes < n * sizeof(float) || !ctx->luts_ready) return PTB_E_BADARG;
        float* d = nullptr;
        CK(cudaMalloc(&d, n * sizeof(float)));
        k_dump_luts<<<(16 * 16 * 16 + 255) / 256, 256, 0, ctx->stream>>>(ctx->F, d); ctx->launches++;
        CK(cudaMemcpyAsync(host_dst, d, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(d);
        return 0;
    }
    return PTB_E_BADARG;
}
extern "C" int64_t ptb_launch_count(ptb_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int ptb_set_timing(ptb_ctx* ctx, int enabled) { if (!ctx) return PTB_E_BADARG; ctx->timing = enabled != 0; return 0; }
extern "C" int ptb_get_stage_ms(ptb_ctx* ctx, float* ms, int n) {
    if (!ctx || !ms) return PTB_E_BADARG;
    for (int i = 0; i < n && i < ST_COUNT; i++) ms[i] = ctx->stage_ms[i];
    return 0;
}
extern "C" const char* ptb_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : nullptr; }
extern "C" const char* ptb_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case PTB_E_BADARG: return "bad argument";
        case PTB_E_NOSCENE: return "no scene uploaded";
        case PTB_E_STATE: return "invalid state or allocation failure";
        case PTB_E_EXCHANGE: return "frame exchange timed out waiting for a peer";
        default: return code > 0 && code < 1000 ? cudaGetErrorString((cudaError_t)code) : "CUDA driver error";
    }
}
