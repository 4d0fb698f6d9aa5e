// SVGF (temporal reprojection, spatial variance, a-trous wavelet filter, finalize) and TAA post passes.
// Reference: Src/CUDA/SVGF/SVGF.h:100-609, Src/CUDA/SVGF/TAA.h:10-172.  The g-buffers are plain pitch-linear
// HBM arrays (the reference uses CUDA surfaces bound to arrays); reads past the edge clamp like cudaBoundaryModeClamp.
// All kernels are 1-D grid-stride over the pitch x height pixel grid, rows contiguous -> coalesced 128-bit accesses.
#pragma once
#include <cuda.h>
#include "ptb_kernels.cuh"

#define PTB_SVGF_EPS 1e-8f
#define PTB_FEEDBACK_ITERATION 1

PTB_DI float4 gbuf_nd(const Frame& P, int x, int y) {
    x = min(max(x, 0), P.pitch - 1); y = min(max(y, 0), P.height - 1);
    return P.svgf.in_normal_depth[x + y * P.pitch];
}
// helper_math lerp semantics (see lerpf)
PTB_DI float4 lerp4(float4 a, float4 b, float t) { return a + t * (b - a); }
PTB_DI float3 lerp3(float3 a, float3 b, float t) { return a + t * (b - a); }

// Rows [ext_y0, ext_y1) of the frame, enumerated row-major over `cols` columns: the pixel grid every filter kernel runs over.
// One GPU: the whole frame.  Several: this rank's filter block + PTB_SVGF_HALO rows.
PTB_DI int svgf_rows(const Frame& P) { return P.svgf.ext_y1 - P.svgf.ext_y0; }

// Last frame's temporal state at row y.  One GPU: the other parity of the local buffers.  Several: every rank owns the history
// of the rows of ITS filter block; rows outside are read from the owner's exchange block over NVLink (L2-only loads: peer data
// must not be served from a stale L1 line).
struct HistoryView { const float4 *direct, *indirect, *moment, *normal_depth, *taa; const int* length; bool remote; };
PTB_DI HistoryView svgf_previous(const Frame& P, int y) {
    HistoryView v;
    const SVGFHistory& h = P.svgf.hist[P.svgf.parity ^ 1];
    if (!P.xchg.svgf || (y >= P.svgf.block_y0 && y < P.svgf.block_y1)) {
        v.direct = h.direct; v.indirect = h.indirect; v.moment = h.moment; v.normal_depth = h.normal_depth; v.taa = h.taa; v.length = h.length; v.remote = false;
    } else {
        const size_t S = size_t(P.fb_stride);
        const float4* b = P.xchg.frames[svgf_block_owner(P, y)] + xchg_history_offset(P.fb_stride, P.svgf.parity ^ 1);
        v.direct = b; v.indirect = b + S; v.moment = b + 2 * S; v.normal_depth = b + 3 * S; v.taa = b + 4 * S; v.length = reinterpret_cast<const int*>(b + 5 * S); v.remote = true;
    }
    return v;
}
PTB_DI float4 hist_load(const float4* p, bool remote) { return remote ? __ldcg(p) : *p; }
PTB_DI int hist_load(const int* p, bool remote) { return remote ? __ldcg(p) : *p; }

PTB_DI bool tap_consistent(const Frame& P, int x, int y, float3 normal, float depth) {
    if (x < 0 || x >= P.width) return false;
    if (y < 0 || y >= P.height) return false;
    HistoryView h = svgf_previous(P, y);
    float4 prev = hist_load(h.normal_depth + (x + y * P.pitch), h.remote);
    float3 pn = oct_decode_normal(f2(prev.x, prev.y));
    return dot(normal, pn) > 0.95f && fabsf(depth - prev.z) < 2.0f;
}

PTB_DI float2 edge_stopping_weights(const Frame& P, int dx, int dy, float2 cgrad, float cdepth, float depth, float3 cn, float3 n,
                                    float cl_d, float cl_i, float l_d, float l_i, float denom_d, float denom_i) {
    float d = cgrad.x * float(dx) + cgrad.y * float(dy);
    float ln_w_z = fabsf(cdepth - depth) / (P.config.sigma_z * fabsf(d) + PTB_SVGF_EPS);
    float w_n = powf(fmaxf(0.0f, dot(cn, n)), P.config.sigma_n);
    float w_d = w_n * expf(-fabsf(cl_d - l_d) * denom_d - ln_w_z);
    float w_i = w_n * expf(-fabsf(cl_i - l_i) * denom_i - ln_w_z);
    return f2(w_d, w_i);
}

__global__ void __launch_bounds__(256) k_svgf_reproject(const __grid_constant__ Frame P, int sample_index) {
    const int total = P.width * svgf_rows(P);
    const SVGFHistory& cur = P.svgf.hist[P.svgf.parity];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int y = P.svgf.ext_y0 + i / P.width, x = i % P.width;
        int px = x + y * P.pitch;
        float4 direct = P.svgf.in_direct[px];
        float4 indirect = P.svgf.in_indirect[px];
        float4 moment;
        moment.x = luminance(direct.x, direct.y, direct.z);
        moment.y = luminance(indirect.x, indirect.y, indirect.z);
        // the squares are rounded products of their own (mul.rn): left as plain `x * x`, ptxas folds them into the subtraction of the
        // moment lerp below (FFMA), which the reference build does not do -- its squares have a second use -- and every frame after
        // the first differed from the reference in the last bit of the second moments (tools/gpu_svgf_diag.py)
        moment.z = __fmul_rn(moment.x, moment.x);
        moment.w = __fmul_rn(moment.y, moment.y);
        float4 nd = P.svgf.in_normal_depth[px];
        float2 sprev = P.svgf.in_screen_prev[px];
        float3 normal = oct_decode_normal(f2(nd.x, nd.y));
        float depth = nd.z, depth_prev = nd.w;
        HistoryView own = svgf_previous(P, y);
        const int history_prev = hist_load(own.length + px, own.remote);
        if (depth == 0.0f) { cur.length[px] = history_prev; continue; }     // sky pixel: nothing is touched (SVGF.h:127-129)
        float u_prev = 0.5f + 0.5f * sprev.x, v_prev = 0.5f + 0.5f * sprev.y;
        float s_prev = u_prev * float(P.width), t_prev = v_prev * float(P.height);
        int x_prev = int(s_prev - 0.5f), y_prev = int(t_prev - 0.5f);
        float fs = s_prev - floorf(s_prev), ft = t_prev - floorf(t_prev);
        float ofs = 1.0f - fs, oft = 1.0f - ft;
        float w0 = ofs * oft, w1 = fs * oft, w2 = ofs * ft;
        float w3 = 1.0f - w0 - w1 - w2;
        float weights[4] = { w0, w1, w2, w3 };
        float wsum = 0.0f;
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                int tap = k + j * 2;
                if (tap_consistent(P, x_prev + k, y_prev + j, normal, depth_prev)) wsum += weights[tap]; else weights[tap] = 0.0f;
            }
        float4 pd = f4(0.0f), pi = f4(0.0f), pm = f4(0.0f);
        if (wsum > 0.0f) {
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    int tap = k + j * 2;
                    if (weights[tap] != 0.0f) {
                        HistoryView h = svgf_previous(P, y_prev + j);
                        int ti = (x_prev + k) + (y_prev + j) * P.pitch;
                        pd += weights[tap] * hist_load(h.direct + ti, h.remote);
                        pi += weights[tap] * hist_load(h.indirect + ti, h.remote);
                        pm += weights[tap] * hist_load(h.moment + ti, h.remote);
                    }
                }
        } else {
            for (int j = -1; j <= 1; j++)
                for (int k = -1; k <= 1; k++) {
                    int tx = x_prev + k, ty = y_prev + j;
                    if (tap_consistent(P, tx, ty, normal, depth_prev)) {
                        HistoryView h = svgf_previous(P, ty);
                        int ti = tx + ty * P.pitch;
                        pd += hist_load(h.direct + ti, h.remote); pi += hist_load(h.indirect + ti, h.remote); pm += hist_load(h.moment + ti, h.remote);
                        wsum += 1.0f;
                    }
                }
        }
        if (wsum > 0.0f) {
            pd = pd / wsum; pi = pi / wsum; pm = pm / wsum;
            int history = history_prev + 1;
            cur.length[px] = history;
            float inv_history = 1.0f / float(history);
            float a_c = fmaxf(P.config.alpha_colour, inv_history);
            float a_m = fmaxf(P.config.alpha_moment, inv_history);
            direct = lerp4(pd, direct, a_c);
            indirect = lerp4(pi, indirect, a_c);
            moment = lerp4(pm, moment, a_m);
            if (history >= 4 || !P.config.enable_spatial_variance) {
                direct.w = fmaxf(0.0f, moment.z - moment.x * moment.x);
                indirect.w = fmaxf(0.0f, moment.w - moment.y * moment.y);
            }
        } else {
            cur.length[px] = 0;
            direct.w = 1.0f; indirect.w = 1.0f;
        }
        P.svgf.in_direct[px] = direct;
        P.svgf.in_indirect[px] = indirect;
        P.svgf.moment[px] = moment;
    }
}

__global__ void __launch_bounds__(256) k_svgf_variance(const __grid_constant__ Frame P, const float4* din, const float4* iin, float4* dout, float4* iout) {
    const int total = P.pitch * svgf_rows(P);             // the reference runs this one over the padded pitch (SVGF.h:293)
    const int* history_length = P.svgf.hist[P.svgf.parity].length;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int y = P.svgf.ext_y0 + idx / P.pitch, x = idx % P.pitch;
        int px = x + y * P.pitch;
        int history = history_length[px];
        float4 cd = din[px], ci = iin[px];
        if (history >= 4) { dout[px] = cd; iout[px] = ci; continue; }
        float ldenom = 1.0f / P.config.sigma_l;
        float cl_d = luminance(cd.x, cd.y, cd.z), cl_i = luminance(ci.x, ci.y, ci.z);
        float4 cnd = gbuf_nd(P, x, y);
        float3 cn = oct_decode_normal(f2(cnd.x, cnd.y));
        float cdepth = cnd.z;
        float2 cgrad = f2(gbuf_nd(P, x + 1, y).z - cdepth, gbuf_nd(P, x, y + 1).z - cdepth);
        if (cdepth == 0.0f) { dout[px] = cd; iout[px] = ci; continue; }
        float sw_d = 1.0f, sw_i = 1.0f;
        float4 sc_d = cd, sc_i = ci, sm = f4(0.0f);
        for (int j = -3; j <= 3; j++) {
            int ty = y + j;
            if (ty < 0 || ty >= P.height) continue;
            for (int k = -3; k <= 3; k++) {
                int tx = x + k;
                if (tx < 0 || tx >= P.width) continue;
                if (k == 0 && j == 0) continue;
                int ti = tx + ty * P.pitch;
                float4 td = din[ti], tind = iin[ti], tm = P.svgf.moment[ti];
                float l_d = luminance(td.x, td.y, td.z), l_i = luminance(tind.x, tind.y, tind.z);
                float4 tnd = P.svgf.in_normal_depth[ti];
                float3 n = oct_decode_normal(f2(tnd.x, tnd.y));
                float2 w = edge_stopping_weights(P, k, j, cgrad, cdepth, tnd.z, cn, n, cl_d, cl_i, l_d, l_i, ldenom, ldenom);
                sw_d += w.x; sw_i += w.y;
                sc_d += w.x * td; sc_i += w.y * tind;
                sm += tm * make_float4(w.x, w.y, w.x, w.y);
            }
        }
        sw_d = fmaxf(sw_d, 1e-6f); sw_i = fmaxf(sw_i, 1e-6f);
        sc_d = sc_d / sw_d; sc_i = sc_i / sw_i;
        sm = make_float4(sm.x / sw_d, sm.y / sw_i, sm.z / sw_d, sm.w / sw_i);
        sc_d.w = fmaxf(0.0f, sm.z - sm.x * sm.x);
        sc_i.w = fmaxf(0.0f, sm.w - sm.y * sm.y);
        dout[px] = sc_d; iout[px] = sc_i;
    }
}

// One a-trous iteration.  Work-group shape: a 32 x 8 pixel tile per 256-thread CTA (a warp covers one 512-byte row segment of every
// plane); for the dense strides 1 and 2 the tile and its halo of the three input planes are staged in shared memory first
// (38 x 14 resp. 36 x 12 texels of 16 B per plane: every texel is fetched from L2 once per CTA instead of up to 9 times per warp).
// The arithmetic and its order are SVGF.h:416-554's -- the outputs are compared bit for bit with the reference kernels.
// The three tensor maps (2-D, float32 view of a pitch x height float4 plane, box = tile + halo) are only read by the TMA variant.
struct AtrousMaps { CUtensorMap direct, indirect, normal_depth; };
PTB_DI void tma_load_tile(void* dst, const CUtensorMap* map, int x_float, int y, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_u32(bar)), "r"(x_float), "r"(y) : "memory");
}
template <int STEP_TILED, bool TMA>
__global__ void __launch_bounds__(256) k_svgf_atrous(const __grid_constant__ Frame P, const float4* din, const float4* iin, float4* dout, float4* iout, int step,
                                                     const __grid_constant__ AtrousMaps maps) {
    constexpr int TW = 32, TH = 8;
    constexpr int HALO = STEP_TILED > 0 ? STEP_TILED : 1;             // STEP_TILED = 1 or 2: halo of the taps; the variance blur needs 1
    constexpr int SW = TW + 2 * HALO, SH = TH + 2 * HALO;
    __shared__ __align__(128) float4 s_d[STEP_TILED > 0 ? SW * SH : 1];
    __shared__ __align__(128) float4 s_i[STEP_TILED > 0 ? SW * SH : 1];
    __shared__ __align__(128) float4 s_nd[STEP_TILED > 0 ? (SW + 1) * (SH + 1) : 1];
    __shared__ __align__(8) unsigned long long s_bar;
    if (TMA && STEP_TILED > 0) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    unsigned phase = 0;
    const int tiles_x = (P.width + TW - 1) / TW, tiles_y = (svgf_rows(P) + TH - 1) / TH;
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const int tx0 = (tile % tiles_x) * TW, ty0 = P.svgf.ext_y0 + (tile / tiles_x) * TH;
        if (STEP_TILED > 0 && TMA) {
            // 2-D TMA: one elected thread posts the three tile + halo boxes (texels outside the plane arrive as zeros and are never
            // read: every tap below uses a clamped or range-checked coordinate); everybody waits on the mbarrier's phase
            __syncthreads();                                          // previous tile's readers are done
            if (threadIdx.x == 0) {
                constexpr unsigned bytes = unsigned(sizeof(float4)) * (2u * SW * SH + (SW + 1) * (SH + 1));
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(bytes) : "memory");
                tma_load_tile(s_d, &maps.direct, (tx0 - HALO) * 4, ty0 - HALO, &s_bar);
                tma_load_tile(s_i, &maps.indirect, (tx0 - HALO) * 4, ty0 - HALO, &s_bar);
                tma_load_tile(s_nd, &maps.normal_depth, (tx0 - HALO) * 4, ty0 - HALO, &s_bar);
            }
            unsigned done = 0;
            while (!done) {
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done) : "r"(smem_u32(&s_bar)), "r"(phase) : "memory");
            }
            phase ^= 1u;
        } else if (STEP_TILED > 0) {
            __syncthreads();                                          // previous tile's readers are done
            for (int t = threadIdx.x; t < SW * SH; t += 256) {
                int sx = t % SW, sy = t / SW;
                int gx = min(max(tx0 - HALO + sx, 0), P.pitch - 1), gy = min(max(ty0 - HALO + sy, 0), P.height - 1);      // clamped: only ever read through clamped or in-range taps
                s_d[t] = din[gx + gy * P.pitch]; s_i[t] = iin[gx + gy * P.pitch];
            }
            for (int t = threadIdx.x; t < (SW + 1) * (SH + 1); t += 256) {
                int sx = t % (SW + 1), sy = t / (SW + 1);
                int gx = min(max(tx0 - HALO + sx, 0), P.pitch - 1), gy = min(max(ty0 - HALO + sy, 0), P.height - 1);
                s_nd[t] = P.svgf.in_normal_depth[gx + gy * P.pitch];
            }
            __syncthreads();
        }
        const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
        const int x = tx0 + lx, y = ty0 + ly;
        if (x >= P.width || y >= P.svgf.ext_y1) continue;
        const int px = x + y * P.pitch;
        // staged reads: (gx, gy) must lie inside the staged window (all taps below do); global otherwise
        auto ld_d = [&](int gx, int gy) { return STEP_TILED > 0 ? s_d[(gx - tx0 + HALO) + (gy - ty0 + HALO) * SW] : din[gx + gy * P.pitch]; };
        auto ld_i = [&](int gx, int gy) { return STEP_TILED > 0 ? s_i[(gx - tx0 + HALO) + (gy - ty0 + HALO) * SW] : iin[gx + gy * P.pitch]; };
        auto ld_nd = [&](int gx, int gy) { return STEP_TILED > 0 ? s_nd[(gx - tx0 + HALO) + (gy - ty0 + HALO) * (SW + 1)] : P.svgf.in_normal_depth[gx + gy * P.pitch]; };
        float vb_d = 0.0f, vb_i = 0.0f;
#pragma unroll
        for (int j = -1; j <= 1; j++) {
            int ty = min(max(y + j, 0), P.height - 1);
#pragma unroll
            for (int k = -1; k <= 1; k++) {
                int tx = min(max(x + k, 0), P.width - 1);
                float v_d = ld_d(tx, ty).w, v_i = ld_i(tx, ty).w;
                float kw = scalbnf(0.25f, -(abs(k) + abs(j)));
                vb_d += v_d * kw; vb_i += v_i * kw;
            }
        }
        float denom_d = rsqrtf(P.config.sigma_l * P.config.sigma_l * fmaxf(0.0f, vb_d) + PTB_SVGF_EPS);
        float denom_i = rsqrtf(P.config.sigma_l * P.config.sigma_l * fmaxf(0.0f, vb_i) + PTB_SVGF_EPS);
        float4 cd = ld_d(x, y), ci = ld_i(x, y);
        float cl_d = luminance(cd.x, cd.y, cd.z), cl_i = luminance(ci.x, ci.y, ci.z);
        float4 cnd = ld_nd(x, y);
        float3 cn = oct_decode_normal(f2(cnd.x, cnd.y));
        float cdepth = cnd.z;
        if (cdepth == 0.0f) {
            // sky pixel: the reference leaves everything as it is (SVGF.h:447-449).  Its history buffers are updated in place, ours
            // alternate between two parities: carry the untouched feedback value over (only rows this rank owns are ever read)
            if (step == (1 << PTB_FEEDBACK_ITERATION) && y >= P.svgf.block_y0 && y < P.svgf.block_y1) {
                const SVGFHistory& prev = P.svgf.hist[P.svgf.parity ^ 1];
                P.svgf.hist[P.svgf.parity].direct[px] = prev.direct[px]; P.svgf.hist[P.svgf.parity].indirect[px] = prev.indirect[px];
            }
            continue;
        }
        float2 cgrad = f2(ld_nd(min(x + 1, P.pitch - 1), y).z - cdepth, ld_nd(x, min(y + 1, P.height - 1)).z - cdepth);
        float sw_d = 1.0f, sw_i = 1.0f;
        float4 sc_d = cd, sc_i = ci;
#pragma unroll
        for (int j = -1; j <= 1; j++) {
            int ty = y + j * step;
            if (ty < 0 || ty >= P.height) continue;
#pragma unroll
            for (int k = -1; k <= 1; k++) {
                int tx = x + k * step;
                if (tx < 0 || tx >= P.width) continue;
                if (k == 0 && j == 0) continue;
                float4 td = ld_d(tx, ty), tind = ld_i(tx, ty);
                float l_d = luminance(td.x, td.y, td.z), l_i = luminance(tind.x, tind.y, tind.z);
                float4 tnd = ld_nd(tx, ty);
                float3 n = oct_decode_normal(f2(tnd.x, tnd.y));
                float2 w = edge_stopping_weights(P, k * step, j * step, cgrad, cdepth, tnd.z, cn, n, cl_d, cl_i, l_d, l_i, denom_d, denom_i);
                sw_d += w.x; sw_i += w.y;
                sc_d += make_float4(w.x, w.x, w.x, w.x * w.x) * td;
                sc_i += make_float4(w.y, w.y, w.y, w.y * w.y) * tind;
            }
        }
        float inv_d = 1.0f / sw_d, inv_i = 1.0f / sw_i;
        sc_d = sc_d * inv_d; sc_i = sc_i * inv_i;
        sc_d.w *= inv_d; sc_i.w *= inv_i;
        dout[px] = sc_d; iout[px] = sc_i;
        if (step == (1 << PTB_FEEDBACK_ITERATION)) { P.svgf.hist[P.svgf.parity].direct[px] = sc_d; P.svgf.hist[P.svgf.parity].indirect[px] = sc_i; }
    }
}

__global__ void __launch_bounds__(256) k_svgf_finalize(const __grid_constant__ Frame P, const float4* cdirect, const float4* cindirect) {
    const int total = P.width * svgf_rows(P);
    const SVGFHistory& cur = P.svgf.hist[P.svgf.parity];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int y = P.svgf.ext_y0 + i / P.width, x = i % P.width;
        int px = x + y * P.pitch;
        float4 direct = cdirect[px], indirect = cindirect[px];
        float4 colour = (direct + indirect) * P.svgf.in_albedo[px];
        P.display[px] = colour;
        if (P.config.enable_taa) {
            colour = colour / (1.0f + luminance(colour.x, colour.y, colour.z));
            colour.x = safe_sqrt(colour.x); colour.y = safe_sqrt(colour.y); colour.z = safe_sqrt(colour.z);
            P.svgf.taa_curr[px] = colour;
        }
        float4 moment = P.svgf.moment[px];
        float4 nd = P.svgf.in_normal_depth[px];
        if (P.config.num_atrous_iterations <= PTB_FEEDBACK_ITERATION) { cur.direct[px] = direct; cur.indirect[px] = indirect; }
        cur.moment[px] = moment;
        cur.normal_depth[px] = nd;
    }
}

PTB_DI float mitchell_netravali(float x) {
    const float B = 1.0f / 3.0f, C = 1.0f / 3.0f;
    x = fabsf(x);
    float x2 = x * x, x3 = x2 * x;
    if (x < 1.0f) return (1.0f / 6.0f) * ((12.0f - 9.0f * B - 6.0f * C) * x3 + (-18.0f + 12.0f * B + 6.0f * C) * x2 + (6.0f - 2.0f * B));
    if (x < 2.0f) return (1.0f / 6.0f) * ((-B - 6.0f * C) * x3 + (6.0f * B + 30.0f * C) * x2 + (-12.0f * B - 48.0f * C) * x + (8.0f * B + 24.0f * C));
    return 0.0f;
}
PTB_DI float3 rgb_to_ycocg(float3 c) { return f3(0.25f * c.x + 0.5f * c.y + 0.25f * c.z, 0.5f * c.x - 0.5f * c.z, -0.25f * c.x + 0.5f * c.y - 0.25f * c.z); }
PTB_DI float3 ycocg_to_rgb(float3 c) { return f3(__saturatef(c.x + c.y - c.z), __saturatef(c.x + c.z), __saturatef(c.x - c.y - c.z)); }
PTB_DI float3 clamp3(float3 v, float3 a, float3 b) { return f3(clampf(v.x, a.x, b.x), clampf(v.y, a.y, b.y), clampf(v.z, a.z, b.z)); }

__global__ void __launch_bounds__(256) k_taa(const __grid_constant__ Frame P, int sample_index) {
    const int total = P.width * (P.svgf.block_y1 - P.svgf.block_y0);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int y = P.svgf.block_y0 + idx / P.width, x = idx % P.width;
        int px = x + y * P.pitch;
        const float4* curr = P.svgf.taa_curr;
        float4 colour = curr[px];
        if (sample_index == 0) { P.display[px] = colour; continue; }
        float2 sprev = P.svgf.in_screen_prev[px];
        float u_prev = 0.5f + 0.5f * sprev.x, v_prev = 0.5f + 0.5f * sprev.y;
        float s_prev = u_prev * float(P.width), t_prev = v_prev * float(P.height);
        int x_prev = int(s_prev + 0.5f), y_prev = int(t_prev + 0.5f);
        float sum_w = 0.0f; float4 sum = f4(0.0f);
        for (int j = y_prev - 2; j < y_prev + 2; j++) {
            if (j < 0 || j >= P.height) continue;
            HistoryView h = svgf_previous(P, j);
            for (int i = x_prev - 2; i < x_prev + 2; i++) {
                if (i < 0 || i >= P.width) continue;
                float w = mitchell_netravali(float(i) + 0.5f - s_prev) * mitchell_netravali(float(j) + 0.5f - t_prev);
                sum_w += w;
                sum += w * hist_load(h.taa + (i + j * P.pitch), h.remote);
            }
        }
        if (sum_w > 0.0f) {
            float3 c_curr = rgb_to_ycocg(f3(colour));
            float3 c_prev = rgb_to_ycocg(f3(sum / sum_w));
            float3 avg = c_curr, var = c_curr * c_curr;
            auto tap = [&](int off) { float3 f = rgb_to_ycocg(f3(curr[px + off])); avg += f; var += f * f; };
            if (x >= 1) { if (y >= 1) tap(-P.pitch - 1); tap(-1); if (y < P.height - 1) tap(P.pitch - 1); }
            if (y >= 1) tap(-P.pitch);
            if (y < P.height - 1) tap(P.pitch);
            if (x < P.width - 1) { if (y >= 1) tap(1 - P.pitch); tap(1); if (y < P.height - 1) tap(1 + P.pitch); }
            avg *= 1.0f / 9.0f; var *= 1.0f / 9.0f;
            float3 sigma2 = var - avg * avg;
            float3 sigma = f3(safe_sqrt(sigma2.x), safe_sqrt(sigma2.y), safe_sqrt(sigma2.z));
            float3 cmin = avg - 1.25f * sigma, cmax = avg + 1.25f * sigma;
            c_prev = clamp3(c_prev, cmin, cmax);
            float3 integrated = ycocg_to_rgb(lerp3(c_prev, c_curr, 0.1f));
            colour.x = integrated.x; colour.y = integrated.y; colour.z = integrated.z;
        }
        P.display[px] = colour;
    }
}

__global__ void __launch_bounds__(256) k_taa_finalize(const __grid_constant__ Frame P) {
    const int total = P.width * (P.svgf.block_y1 - P.svgf.block_y0);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int y = P.svgf.block_y0 + idx / P.width, x = idx % P.width;
        int px = x + y * P.pitch;
        float4 colour = P.display[px];
        P.svgf.hist[P.svgf.parity].taa[px] = colour;
        colour = colour * colour;
        colour = colour / (1.0f - luminance(colour.x, colour.y, colour.z));
        P.display[px] = colour;
    }
}

// End of a filtered frame on ONE GPU: the trace-side planes ARE the filter inputs; clear every enabled framebuffer
// (aovs_clear_to_zero, Integrator.cpp:377-383) and the three g-buffers (SVGF.h:600-606, TAA.h:168-171).  With several GPUs
// k_svgf_push has already cleared what it shipped; only the other AOV planes remain.
__global__ void __launch_bounds__(256) k_clear_framebuffers(const __grid_constant__ Frame P, int clear_gbuffers) {
    const int total = P.pitch * P.height;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < PTB_AOV_COUNT; k++) if (P.aov[k].fb) P.aov[k].fb[i] = f4(0.0f);
        if (clear_gbuffers) { P.svgf.gbuf_normal_depth[i] = f4(0.0f); P.svgf.gbuf_ids[i] = make_int2(0, 0); P.svgf.gbuf_screen_prev[i] = f2(0.0f, 0.0f); }
    }
}

// launch sequence of the SVGF branch of Pathtracer::render (Pathtracer.cpp:798-837)
// The window's post-processing pass (Src/Shaders/post.frag:18-41): clamp to >= 0, ACES filmic curve, gamma 2.2; written as 8-bit RGBA
// like the GL frame buffer the reference presents.  Row 0 = bottom of the image, as everywhere on the device.
__global__ void __launch_bounds__(256) k_present(const __grid_constant__ Frame P, const float4* src, uchar4* out) {
    const int total = P.pitch * P.height;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float4 c = src[i];
        float v[3] = { fmaxf(c.x, 0.0f), fmaxf(c.y, 0.0f), fmaxf(c.z, 0.0f) };
        unsigned char o[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float x = v[k];
            float t = __saturatef((x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f));
            o[k] = (unsigned char)__float2uint_rn(255.0f * powf(t, 1.0f / 2.2f));
        }
        out[i] = make_uchar4(o[0], o[1], o[2], 255);
    }
}

// host hook: builds the 2-D tensor map of one float4 plane for a box of box_w x box_h texels (set by ptb_api.cu when the driver offers
// cuTensorMapEncodeTiled; nullptr = stage the tiles with plain loads)
typedef int (*EncodeTileMapFn)(CUtensorMap* map, const void* plane, int pitch, int height, int box_w, int box_h);
static EncodeTileMapFn encode_tile_map = nullptr;

static int launch_svgf(Frame& F, cudaStream_t st, int sample_index, int grid, long long* launches) {
    // world > 1: tile-local filter.  Every rank stores the rows it traced into the input planes of the ranks that filter them
    // (block + halo), filters ITS block of rows, reads last frame's history across block borders from the owning rank, and ships
    // its displayed rows to every rank.
    const bool sharded = F.world != 1;
    if (sharded) {
        if (!(F.xchg.count > 0 && F.xchg.svgf)) { fprintf(stderr, "[ptb] SVGF with world > 1 needs the frame exchange (ptb_exchange_connect*)\n"); return PTB_E_STATE; }
        k_svgf_push<<<grid, 256, 0, st>>>(F); k_svgf_wait_arrivals<<<1, 1, 0, st>>>(F); (*launches) += 2;
    }
    k_svgf_reproject<<<grid, 256, 0, st>>>(F, sample_index); (*launches)++;
    float4* din = F.svgf.in_direct; float4* iin = F.svgf.in_indirect;
    float4* dout = F.aov[PTB_AOV_RADIANCE_DIRECT].acc; float4* iout = F.aov[PTB_AOV_RADIANCE_INDIRECT].acc;
    if (F.config.enable_spatial_variance) {
        k_svgf_variance<<<grid, 256, 0, st>>>(F, din, iin, dout, iout); (*launches)++;
        float4* t = din; din = dout; dout = t; t = iin; iin = iout; iout = t;
    }
    for (int i = 0; i < F.config.num_atrous_iterations; i++) {
        AtrousMaps maps; memset(&maps, 0, sizeof(maps));
        const int halo = i + 1;                    // strides 1 and 2 are staged in shared memory
        bool tma = i < 2 && encode_tile_map && encode_tile_map(&maps.direct, din, F.pitch, F.height, 32 + 2 * halo, 8 + 2 * halo) == 0 &&
                   encode_tile_map(&maps.indirect, iin, F.pitch, F.height, 32 + 2 * halo, 8 + 2 * halo) == 0 &&
                   encode_tile_map(&maps.normal_depth, F.svgf.in_normal_depth, F.pitch, F.height, 32 + 2 * halo + 1, 8 + 2 * halo + 1) == 0;
        if (i == 0)      { if (tma) k_svgf_atrous<1, true><<<grid, 256, 0, st>>>(F, din, iin, dout, iout, 1, maps); else k_svgf_atrous<1, false><<<grid, 256, 0, st>>>(F, din, iin, dout, iout, 1, maps); }
        else if (i == 1) { if (tma) k_svgf_atrous<2, true><<<grid, 256, 0, st>>>(F, din, iin, dout, iout, 2, maps); else k_svgf_atrous<2, false><<<grid, 256, 0, st>>>(F, din, iin, dout, iout, 2, maps); }
        else             k_svgf_atrous<0, false><<<grid, 256, 0, st>>>(F, din, iin, dout, iout, 1 << i, maps);
        (*launches)++;
        float4* t = din; din = dout; dout = t; t = iin; iin = iout; iout = t;
    }
    k_svgf_finalize<<<grid, 256, 0, st>>>(F, din, iin); (*launches)++;
    if (F.config.enable_taa) {
        k_taa<<<grid, 256, 0, st>>>(F, sample_index); (*launches)++;
        k_taa_finalize<<<grid, 256, 0, st>>>(F); (*launches)++;
    }
    k_clear_framebuffers<<<grid, 256, 0, st>>>(F, sharded ? 0 : 1); (*launches)++;
    if (sharded) { k_svgf_push_display<<<grid, 256, 0, st>>>(F); k_exchange_wait<<<1, 1, 0, st>>>(F); (*launches) += 2; }
    return 0;
}
