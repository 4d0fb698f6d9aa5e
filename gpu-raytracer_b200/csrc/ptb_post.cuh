// Init-time Kulla-Conty LUT baking (KullaConty.h:83-240).
#pragma once
#include "ptb_device.cuh"

// ------------------------------------------------------------------------------------------ LUT bake
// Directional albedo of the single-scatter GGX lobe, Monte-Carlo integrated with 100 000 samples per cell, then
// cosine-averaged.  Layout: index = ior + rough * D + cos * D * D (dielectric), rough + cos * C (conductor).
#define PTB_LUT_SAMPLES 100000

__global__ void __launch_bounds__(256) k_integrate_dielectric(const __grid_constant__ Frame P, bool entering, float* lut) {
    const int D = PTB_LUT_DIELECTRIC_DIM;
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= D * D * D) return;
    int i = tid % D, r = (tid / D) % D, c = (tid / (D * D)) % D;
    float ior = remapf((float(i) + 0.5f) / float(D), 0.0f, 1.0f, PTB_LUT_DIELECTRIC_MIN_IOR, PTB_LUT_DIELECTRIC_MAX_IOR);
    float eta = entering ? 1.0f / ior : ior;
    float rough = (float(r) + 0.5f) / float(D);
    float cos_theta = (float(c) + 0.5f) / float(D);
    float sin_theta = safe_sqrt(1.0f - square(cos_theta));
    float3 wi = f3(sin_theta, 0.0f, cos_theta);
    float avg = 0.0f;
    for (int s = 0; s < PTB_LUT_SAMPLES; s++) {
        float rf = rng2<DIM_BSDF_0>(P, tid, 0, s).y;
        float2 rb = rng2<DIM_BSDF_1>(P, tid, 0, s);
        float ax = roughness_to_alpha(rough), ay = roughness_to_alpha(rough);
        float3 wm = sample_vndf_ggx(wi, ax, ay, rb.x, rb.y);
        float F = fresnel_dielectric(abs_dot(wi, wm), eta);
        bool reflected = rf < F;
        float3 wo = reflected ? reflect_direction(wi, wm) : refract_direction(wi, wm, eta);
        float weight = 0.0f;
        if (!(reflected ^ (wo.z >= 0.0f))) {
            float Dm = ggx_D(wm, ax, ay);
            float G1 = ggx_G1(wi, ax, ay);
            float G2 = ggx_G2(wo, wi, wm, ax, ay);
            float i_m = abs_dot(wi, wm), o_m = abs_dot(wo, wm);
            float w = G2 / G1;
            float pdf;
            if (reflected) pdf = F * G1 * Dm / (4.0f * wi.z);
            else           pdf = (1.0f - F) * G1 * Dm * i_m * o_m / (wi.z * square(eta * i_m + o_m));
            weight = pdf_is_valid(pdf) ? w : 0.0f;
        }
        avg = avg + (weight - avg) / float(s + 1);
    }
    lut[tid] = avg;
}

__global__ void k_average_dielectric(const float* dir, float* out) {
    const int D = PTB_LUT_DIELECTRIC_DIM;
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= D * D) return;
    int i = tid % D, r = (tid / D) % D;
    float avg = 0.0f;
    for (int c = 0; c < D; c++) {
        float cos_theta = (float(c) + 0.5f) / float(D);
        float v = dir[i + r * D + c * D * D] * cos_theta;
        avg = avg + (v - avg) / float(c + 1);
    }
    out[tid] = 2.0f * avg;
}

__global__ void __launch_bounds__(256) k_integrate_conductor(const __grid_constant__ Frame P, float* lut) {
    const int C = PTB_LUT_CONDUCTOR_DIM;
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= C * C) return;
    int r = tid % C, c = (tid / C) % C;
    float rough = (float(r) + 0.5f) / float(C);
    float cos_theta = (float(c) + 0.5f) / float(C);
    float sin_theta = safe_sqrt(1.0f - square(cos_theta));
    float3 wi = f3(sin_theta, 0.0f, cos_theta);
    float avg = 0.0f;
    for (int s = 0; s < PTB_LUT_SAMPLES; s++) {
        float2 rb = rng2<DIM_BSDF_0>(P, tid, 0, s);
        float ax = roughness_to_alpha(rough), ay = roughness_to_alpha(rough);
        float3 wm = sample_vndf_ggx(wi, ax, ay, rb.x, rb.y);
        float3 wo = reflect_direction(wi, wm);
        float weight = 0.0f;
        if (!(dot(wo, wm) <= 0.0f || wo.z <= 0.0f)) {
            float Dm = ggx_D(wm, ax, ay);
            float G1 = ggx_G1(wi, ax, ay);
            float G2 = ggx_G2(wo, wi, wm, ax, ay);
            float w = G2 / G1;
            float pdf = G1 * Dm / (4.0f * wi.z);
            weight = pdf_is_valid(pdf) ? w : 0.0f;
        }
        avg = avg + (weight - avg) / float(s + 1);
    }
    lut[tid] = avg;
}

__global__ void k_average_conductor(const float* dir, float* out) {
    const int C = PTB_LUT_CONDUCTOR_DIM;
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= C) return;
    int r = tid % C;
    float avg = 0.0f;
    for (int c = 0; c < C; c++) {
        float cos_theta = (float(c) + 0.5f) / float(C);
        float v = dir[r + c * C] * cos_theta;
        avg = avg + (v - avg) / float(c + 1);
    }
    out[tid] = 2.0f * avg;
}


// debug/parity tap: the six LUTs sampled at their texel centres -> 2*16^3 + 2*16^2 + 32^2 + 32 floats
__global__ void k_dump_luts(const __grid_constant__ Frame P, float* out) {
    const int D = PTB_LUT_DIELECTRIC_DIM, C = PTB_LUT_CONDUCTOR_DIM;
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < D * D * D) {
        int i = tid % D, r = (tid / D) % D, c = tid / (D * D);
        float u = (i + 0.5f) / D, v = (r + 0.5f) / D, w = (c + 0.5f) / D;
        out[tid] = tex3D<float>(P.lut_dielectric_dir_enter, u, v, w);
        out[D * D * D + tid] = tex3D<float>(P.lut_dielectric_dir_leave, u, v, w);
    }
    if (tid < D * D) {
        float u = (tid % D + 0.5f) / D, v = (tid / D + 0.5f) / D;
        out[2 * D * D * D + tid] = tex2D<float>(P.lut_dielectric_enter, u, v);
        out[2 * D * D * D + D * D + tid] = tex2D<float>(P.lut_dielectric_leave, u, v);
    }
    if (tid < C * C) out[2 * D * D * D + 2 * D * D + tid] = tex2D<float>(P.lut_conductor_dir, (tid % C + 0.5f) / C, (tid / C + 0.5f) / C);
    if (tid < C) out[2 * D * D * D + 2 * D * D + C * C + tid] = tex1D<float>(P.lut_conductor, (tid + 0.5f) / C);
}
