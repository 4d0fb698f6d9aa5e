// Device-visible structures of the B200 wavefront path tracer.
//
// Unlike the reference (some 60 module globals set by name, Integrator.cpp:15-30 / Pathtracer.cpp:16-38),
// every kernel receives ONE `Frame` by value as a __grid_constant__ parameter: pointers into HBM,
// the 44-byte GPUConfig and the 60-byte camera.  Ray streams are SoA of 16-byte vectors so every
// queue access is a single 128-bit transaction per lane (coalesced 512 B per warp):
//
//   RayQueue   od0 = (origin.xyz, dir.x)  od1 = (dir.y, dir.z, cone_angle, cone_width)
//              hit = (mesh_id, triangle_id, t bits, u16|v16<<16)      (layout of Buffers.h:28-32)
//              path = (throughput.xyz, last_pdf)   pix = pixel_index | flags   medium = medium id
//   ShadowQueue od0 = (origin.xyz, dir.x)  od1 = (dir.y, dir.z, max_distance, pixel_index bits)
//              illum = (illumination.xyz, -)
//   material queues hold 4-byte INDICES into the current RayQueue (compaction by index: the sort pass moves
//   4 bytes per surviving ray instead of re-writing a 52-byte payload like Pathtracer.cu:426-456).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ptb.h"

#define PTB_MAX_BOUNCES 128          // also a multiplier inside the RNG hash (Sampling.h:46) -- must stay 128
#define PTB_PMJ_SEQUENCES 64
#define PTB_PMJ_SAMPLES 4096
#define PTB_BLUE_NOISE_TEXTURES 16
#define PTB_BLUE_NOISE_DIM 128

#define PTB_FLAG_ALLOW_NEE (1u << 31)
#define PTB_FLAG_INSIDE_MEDIUM (1u << 30)
#define PTB_FLAGS_ALL (PTB_FLAG_ALLOW_NEE | PTB_FLAG_INSIDE_MEDIUM)

#define PTB_LUT_DIELECTRIC_DIM 16
#define PTB_LUT_DIELECTRIC_MIN_IOR 1.0001f
#define PTB_LUT_DIELECTRIC_MAX_IOR 2.5f
#define PTB_LUT_CONDUCTOR_DIM 32

enum { PTB_MAT_LIGHT = 0, PTB_MAT_DIFFUSE = 1, PTB_MAT_PLASTIC = 2, PTB_MAT_DIELECTRIC = 3, PTB_MAT_CONDUCTOR = 4 };

struct RayQueue {
    float4*   od0;
    float4*   od1;
    uint4*    hit;
    float4*   path;
    unsigned* pix;
    int*      medium;
};

struct ShadowQueue {
    float4* od0;
    float4* od1;
    float4* illum;
};

// per-bounce queue sizes; all bounces kept so one reset per frame suffices (same idea as Pathtracer.cu:100-116)
struct Counters {
    int trace[PTB_MAX_BOUNCES];
    int mat[4][PTB_MAX_BOUNCES];     // diffuse, plastic, dielectric, conductor
    int shadow[PTB_MAX_BOUNCES];
    int retired[PTB_MAX_BOUNCES];
    int retired_shadow[PTB_MAX_BOUNCES];
};

struct RayTotals {                    // 64-bit running sums for Mrays/s, folded in by k_fold_counters
    unsigned long long trace[PTB_MAX_BOUNCES];
    unsigned long long shadow[PTB_MAX_BOUNCES];
    unsigned long long mat[4];
    unsigned long long frames;
};

struct TraceStats {                   // [0] closest-hit, [1] shadow; filled only by the STATS kernel variants
    unsigned long long rays, nodes, triangles, instance_transforms, misses;
};

struct AOVBuffers {
    float4* fb;   // written during the current pass
    float4* acc;  // running mean over passes
};

struct TextureEntry {                 // 16 bytes, Integrator.h:129-132
    cudaTextureObject_t tex;
    float lod_bias;
    float pad;
};

// Temporal state of SVGF + TAA, one parity.  The filter of frame k writes hist[k & 1] and reads hist[(k - 1) & 1]; with several ranks
// every rank filters -- and owns the history of -- one contiguous block of rows, and reads of the previous parity go to the OWNER of
// the row (peer memory), see ptb_svgf.cuh.
struct SVGFHistory {
    float4* direct;
    float4* indirect;
    float4* moment;
    float4* normal_depth;
    float4* taa;                      // taa_frame_prev
    int*    length;
};
struct SVGFBuffers {
    // trace side: written by k_sort / k_shade at the rows this rank TRACES (interleaved bands)
    float4* gbuf_normal_depth;        // (oct normal.xy, depth, depth gradient)
    int2*   gbuf_ids;                 // (mesh id, triangle id)
    float2* gbuf_screen_prev;         // previous-frame screen position
    // filter side: what the filter chain of this rank reads.  world == 1: the trace-side planes themselves; world > 1: planes in
    // this rank's exchange block that every tracing rank stores its rows into (k_svgf_push)
    float4* in_direct;
    float4* in_indirect;
    float4* in_albedo;
    float4* in_normal_depth;
    int2*   in_ids;
    float2* in_screen_prev;
    float4* moment;                   // frame_buffer_moment
    float4* taa_curr;
    SVGFHistory hist[2];
    int     parity;                   // hist[parity] is written this frame, hist[parity ^ 1] is last frame's
    int     block_y0, block_y1;       // rows this rank filters and whose history it owns
    int     ext_y0, ext_y1;           // block + halo (PTB_SVGF_HALO rows each side, clipped): rows the filter chain is evaluated on
    float   view_projection[16];
    float   view_projection_prev[16];
};
// Rows of halo a filter block needs around it so that its own rows come out exactly as a whole-frame filter computes them:
// TAA 1 + a-trous strides 32+16+8+4+2+1 + variance blur 1 + 7x7 variance 3 + depth gradient 1 = 69 (SVGF.h:284-554); rounded up.
#define PTB_SVGF_HALO 72

// Frame exchange over NVLink peer memory (multi-GPU).  Every rank owns one block {2 full frames, control words};
// the blocks of all ranks are mapped into every rank (CUDA IPC between processes, plain peer access inside one).  The last
// k_accumulate of a frame stores each finished pixel straight into the frame of EVERY rank -- the all-gather is fused into the
// accumulate kernel, pixel by pixel, instead of export -> NCCL all_gather -> assemble -- then publishes its frame number in its
// own arrival slot on every rank; k_exchange_wait spins until all slots of the local block carry the current frame number.  Frames alternate between two buffers so a fast rank never overwrites a
// frame its peer is still reading.
#define PTB_MAX_PEERS 16
struct ExchangeControl {
    unsigned arrivals[PTB_MAX_PEERS]; // slot s: number of frames whose rows rank s has finished storing into this block
    unsigned blocks_done;             // local: CTAs of k_accumulate that finished storing
    unsigned epoch;                   // local: frames completed
    unsigned status;                  // local: 0 ok, 1 = wait timed out
    // SVGF input exchange (k_svgf_push): the same protocol on its own counters.  The input planes are double-buffered by frame
    // parity; a rank cannot run two frames ahead of a peer (its own filter of frame k+1 needs that peer's rows of frame k+1)
    unsigned svgf_arrivals[PTB_MAX_PEERS];
    unsigned svgf_blocks_done;
    unsigned svgf_epoch;
};
struct Exchange {
    int     count;                    // ranks taking part (0 = off)
    int     push;                     // this launch is the last accumulate of a frame: store to the peers
    float4* frames[PTB_MAX_PEERS];    // peer-mapped base of rank r's block: frame parity p at frames[r] + p * pitch * height
    ExchangeControl* control[PTB_MAX_PEERS];
    // SVGF with world > 1 (tile-local filter): behind the two frames every block carries, per parity, the six filter-input planes
    // (stored into by the tracing ranks) and the temporal history planes of the rows this rank owns (read by its neighbours)
    int     svgf;                     // this pass runs the tile-local filter over the exchange blocks
    int     rows_per_block;           // filter block height: rank r owns rows [r * rows_per_block, (r + 1) * rows_per_block)
};
#define PTB_XCHG_HEADER 512
// Layout of a block behind the header, in float4 units (S = pitch * height, a multiple of 32):
//   [0, 2S)                       the two gathered frames
//   inputs of parity p  at 2S + p * 5S:        direct S | indirect S | albedo S | normal_depth S | ids S/2 | screen_prev S/2
//   history of parity p at 12S + p * 5.25S:    direct S | indirect S | moment S | normal_depth S | taa S | length S/4
__host__ __device__ inline size_t xchg_input_offset(int S, int parity) { return size_t(S) * 2 + size_t(parity) * (size_t(S) * 5); }
__host__ __device__ inline size_t xchg_history_offset(int S, int parity) { return size_t(S) * 12 + size_t(parity) * (size_t(S) * 5 + size_t(S) / 4); }
#define PTB_XCHG_BLOCK_FLOAT4(S) (size_t(S) * 12 + 2 * (size_t(S) * 5 + size_t(S) / 4))

struct Frame {
    // film + tile ownership (rows are dealt to ranks in interleaved bands)
    int width, height, pitch;
    int rank, world, band_rows;
    int local_pixels;                 // pixels this rank traces per pass
    // A "wave" carries wave_samples consecutive passes (sample indices first_sample ...) through the pipeline at once: with
    // 180 GB of HBM the 8 spp of a frame need not be traced one pass at a time (the reference's per-pass launches leave the late,
    // nearly empty bounces latency bound).  Every ray carries its pass slot next to its pixel index, each slot has its own
    // framebuffer plane, and k_accumulate folds the planes into the running mean in pass order, so the result is bit-identical
    // to tracing the passes one after another.
    int wave_samples, first_sample;
    int pix_bits;                     // pixel index occupies the low pix_bits of the `pix` word, the pass slot the bits above (below the 2 flag bits)
    int fb_stride;                    // pitch * height: distance between framebuffer planes of consecutive slots
    ptb_config config;
    ptb_camera camera;

    RayQueue    q[2];
    ShadowQueue sq;
    int*        matq[4];
    Counters*   counters;
    RayTotals*  totals;
    TraceStats* trace_stats;
    AOVBuffers  aov[PTB_AOV_COUNT];
    float4*     display;              // what the reference writes to its GL surface
    int*        pixel_query;          // {pixel_index, mesh_id, triangle_id} (PixelQuery, Pathtracer.cu:120): filled by k_sort at bounce 0

    // scene
    const float4*        triangles;   // 6 float4 per triangle
    const float4*        nodes8;      // 5 float4 per CWBVH node
    const float4*        nodes2;      // 2 float4 per binary node
    const float4*        nodes4;      // 8 float4 per 4-wide node (BVH.h:25-59): lo x,y,z / hi x,y,z of the four children, then 4 x (index, count)
    int                  tlas_nodes;  // nodes8 [0, tlas_nodes) are the TLAS
    const int*           mesh_roots;
    const int*           mesh_material_ids;
    const float4*        mesh_transforms;
    const float4*        mesh_transforms_inv;
    const float4*        mesh_transforms_prev;
    const signed char*   material_types;
    const float4*        materials;   // 2 float4 per material
    const float4*        media;       // 2 float4 per medium
    const TextureEntry*  textures;
    cudaTextureObject_t  sky_tex;
    float                sky_scale;
    const float2*        pmj;
    const uchar2*        blue_noise;

    // lights
    float        lights_total_weight;
    const int*   light_triangle_indices;
    const float* light_triangle_cdf;
    int          light_mesh_count;
    const float* light_mesh_cdf;
    const int2*  light_mesh_triangle_span;
    const int*   light_mesh_transform_indices;

    // Kulla-Conty LUTs
    cudaTextureObject_t lut_dielectric_dir_enter, lut_dielectric_dir_leave, lut_dielectric_enter, lut_dielectric_leave;
    cudaTextureObject_t lut_conductor_dir, lut_conductor;

    SVGFBuffers svgf;
    Exchange xchg;

    // Static merge (ptb_api.cu: rebuild_static_merge): every identity-transform instance is also reachable through ONE merged
    // CWBVH (nodes appended to nodes8 at flat_root); rays walk that first and skip those instances in the TLAS.
    int           flat_root;          // node index of the merged BVH's root, -1 = none
    int           flat_all;           // every instance is merged: the TLAS is not traversed at all
    int           flat_node_count;
    const float4* flat_tris;          // 3 float4 per merged triangle: p0, e1, e2 (36 B) + original triangle id + merged slot
    const int*    flat_slot_instance; // merged slot -> current instance index (TLAS leaf order can change between frames)
    // Woop intersection (ptb_set_intersector(PTB_INTERSECT_WOOP), merged BVH only): per triangle reference the 3x4 affine map that
    // takes the triangle to the unit triangle in the z = 0 plane (Woop 2004); ids live in a side table read only on an accepted hit
    const float4* flat_woop;          // 3 float4 per reference: rows of the map; nullptr = Moeller-Trumbore (what the reference does)
    const int2*   flat_who;           // (original triangle id, merged slot) per reference

    // Ray ordering (see k_bin_count / k_bin_scatter): incoherent queues are traced in direction-bin order through `order`
    int            order_bins;        // 0 = trace in queue order, else 8 (octants) or 64 (8x8 octahedral cells)
    unsigned*      order;             // position -> ray index, one per queue slot
    unsigned*      bin_rank;          // ray index -> rank inside its bin
    unsigned char* bin_key;           // ray index -> bin
    int*           bin_counts;        // [PTB_ORDER_MAX_BOUNCE][2][64] histogram rows, zeroed by k_begin_pass
};
#define PTB_ROOT_IDENTITY 0x80000000u   // mesh_roots bit 31: identity transform (Integrator.cpp, root | identity << 31)
#define PTB_ROOT_MERGED   0x40000000u   // mesh_roots bit 30 (device copy only): instance lives in the merged static BVH
#define PTB_FLAT_MESH     (-2147483647 - 1) // mesh_id while traversing the merged BVH; a hit found there carries -(2 + merged slot)
#define PTB_ORDER_MAX_BINS 64
#define PTB_ORDER_MAX_BOUNCE 16
