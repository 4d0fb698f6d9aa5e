// Implementation of the C++ facade (ptb_pathtracer.h) over the C ABI.  Links against libptb.so only.
#include "ptb_pathtracer.h"

#include <cmath>
#include <cstring>

namespace ptb {

Pathtracer::Pathtracer(const ptb_scene& scene, const CameraDesc& cam, int device, int rank, int world, int band_rows)
    : camera(cam), scene_(scene), device_(device), rank_(rank), world_(world), band_rows_(band_rows), mesh_count_(scene.mesh_count) {
    // defaults of GPUConfig (Common.h:39-67)
    std::memset(&gpu_config, 0, sizeof(gpu_config));
    gpu_config.reconstruction_filter = 2; gpu_config.aov_mask = 1u; gpu_config.num_bounces = 10;
    gpu_config.enable_mipmapping = 1; gpu_config.enable_next_event_estimation = 1; gpu_config.enable_multiple_importance_sampling = 1;
    gpu_config.enable_russian_roulette = 1; gpu_config.enable_svgf = 0; gpu_config.enable_spatial_variance = 1; gpu_config.enable_taa = 1;
    gpu_config.alpha_colour = 0.1f; gpu_config.alpha_moment = 0.1f; gpu_config.num_atrous_iterations = 6;
    gpu_config.sigma_z = 4.0f; gpu_config.sigma_n = 16.0f; gpu_config.sigma_l = 10.0f;
}

Pathtracer::~Pathtracer() { cuda_free(); delete[] moved_xf_; }

void Pathtracer::check(int code, const char* what) {
    if (code != 0) throw Error(std::string(what) + " failed: " + ptb_error_string(code), code);
}

void Pathtracer::cuda_init(unsigned, int width, int height) {
    if (ctx_) cuda_free();
    check(ptb_create(&ctx_, device_, width, height, rank_, world_, band_rows_), "ptb_create");
    check(ptb_upload_scene(ctx_, &scene_), "ptb_upload_scene");
    screen_width = width; screen_height = height; screen_pitch = (width + 31) / 32 * 32; pixel_count = width * height;
    invalidated_scene = false;          // the TLAS arrives with the scene; ptb_update_instances moves instances afterwards
    invalidated_camera = true; invalidated_gpu_config = true; invalidated_aovs = true;
    have_view_projection_ = false;
    sample_index = 0;
}

void Pathtracer::cuda_free() {
    if (ctx_) { ptb_destroy(ctx_); ctx_ = nullptr; }
}

void Pathtracer::resize_free() {
    if (ctx_) check(ptb_sync(ctx_), "ptb_sync");        // the device memory itself is re-sized in resize_init (ptb_resize)
}

void Pathtracer::resize_init(unsigned, int width, int height) {
    if (!ctx_) throw Error("resize_init before cuda_init", PTB_E_STATE);
    check(ptb_resize(ctx_, width, height), "ptb_resize");
    screen_width = width; screen_height = height; screen_pitch = (width + 31) / 32 * 32; pixel_count = width * height;
    invalidated_camera = true;          // scene.camera.resize(width, height)
    invalidated_gpu_config = true;
    have_view_projection_ = false;
    sample_index = 0;
}

// Camera::recalibrate + Camera::update (Src/Renderer/Camera.cpp:20-42,86-95) and the block Integrator::update uploads (Integrator.cpp:456-474)
void Pathtracer::use_camera_block(const ptb_camera* block, const float* vp) {
    have_block_ = block != nullptr;
    if (block) { block_ = *block; std::memcpy(block_vp_, vp, 64); }
    invalidated_camera = true;
}

void Pathtracer::upload_camera() {
    if (have_block_) {
        ptb_camera c = block_;
        if (gpu_config.enable_svgf) c.aperture_radius = 0.0f;
        if (have_view_projection_) std::memcpy(view_projection_prev_, view_projection_, 64); else std::memcpy(view_projection_prev_, block_vp_, 64);
        std::memcpy(view_projection_, block_vp_, 64);
        have_view_projection_ = true;
        check(ptb_set_camera(ctx_, &c, view_projection_, view_projection_prev_), "ptb_set_camera");
        return;
    }
    const float half_w = 0.5f * float(screen_width), half_h = 0.5f * float(screen_height);
    const float tan_half = std::tan(0.5f * camera.fov);
    const float d = half_w / tan_half;
    const float* R = camera.rotation;
    auto rot = [&](float x, float y, float z, float* out) {
        out[0] = R[0] * x + R[1] * y + R[2] * z; out[1] = R[3] * x + R[4] * y + R[5] * z; out[2] = R[6] * x + R[7] * y + R[8] * z; };
    ptb_camera c;
    std::memcpy(c.position, camera.position, 12);
    rot(-half_w, -half_h, -d, c.bottom_left_corner);
    rot(1.0f, 0.0f, 0.0f, c.x_axis);
    rot(0.0f, 1.0f, 0.0f, c.y_axis);
    c.pixel_spread_angle = std::atan(2.0f * tan_half / float(screen_width));
    c.aperture_radius = gpu_config.enable_svgf ? 0.0f : camera.aperture_radius;     // "SVGF and DoF cannot simultaneously be enabled" (Integrator.cpp:433-437)
    c.focal_distance = camera.focal_distance;
    // view_projection = perspective(fov, aspect, near, far) * R^T * T(-position), row-major
    const float n = camera.near_plane, f = camera.far_plane, aspect = half_h / half_w;
    float P[16] = {}; P[0] = 1.0f / tan_half; P[5] = 1.0f / (aspect * tan_half); P[10] = -(f + n) / (f - n); P[11] = -2.0f * f * n / (f - n); P[14] = -1.0f;
    float V[16] = {};
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) V[i * 4 + j] = R[j * 3 + i];
        V[i * 4 + 3] = -(R[0 * 3 + i] * camera.position[0] + R[1 * 3 + i] * camera.position[1] + R[2 * 3 + i] * camera.position[2]);
    }
    V[15] = 1.0f;
    float VP[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { float s = 0.0f; for (int k = 0; k < 4; k++) s += P[i * 4 + k] * V[k * 4 + j]; VP[i * 4 + j] = s; }
    if (have_view_projection_) std::memcpy(view_projection_prev_, view_projection_, 64); else std::memcpy(view_projection_prev_, VP, 64);
    std::memcpy(view_projection_, VP, 64);
    have_view_projection_ = true;
    check(ptb_set_camera(ctx_, &c, view_projection_, view_projection_prev_), "ptb_set_camera");
}

void Pathtracer::update(float) {
    if (!ctx_) throw Error("update before cuda_init", PTB_E_STATE);
    if (pixel_query_status == PixelQueryStatus::PENDING && sample_index > 0) {     // read-back of a query the last pass answered (Integrator.cpp:483-494)
        check(ptb_get_pixel_query(ctx_, &pixel_query.mesh_id, &pixel_query.triangle_id), "ptb_get_pixel_query");
        pixel_query.pixel_index = -1;
        pixel_query_status = PixelQueryStatus::OUTPUT_READY;
    }
    bool scene_moved = false;
    if (invalidated_scene && moved_xf_) {                              // Integrator.cpp:441-452: build_tlas() -- here a refit on the device
        check(ptb_refit_instances(ctx_, moved_xf_, moved_xf_ + size_t(mesh_count_) * 12, nullptr), "ptb_refit_instances");
        delete[] moved_xf_; moved_xf_ = nullptr;
        scene_moved = true;
    }
    invalidated_scene = false;
    const bool camera_moved = invalidated_camera || scene_moved;
    if (invalidated_camera) {
        upload_camera();
        invalidated_camera = false;
    } else if (gpu_config.enable_svgf && have_view_projection_) {
        std::memcpy(view_projection_prev_, view_projection_, 64);      // static camera: the previous matrix catches up (Camera.cpp:90)
        upload_camera();
    }
    if (invalidated_gpu_config || invalidated_aovs) {
        invalidated_gpu_config = false; invalidated_aovs = false;
        sample_index = 0;
        check(ptb_set_config(ctx_, &gpu_config), "ptb_set_config");
    } else if (camera_moved && !gpu_config.enable_svgf) {
        sample_index = 0;
    } else {
        sample_index++;                                                // Integrator.cpp:518-526
    }
}

void Pathtracer::move_instances(const float* transforms, const float* transforms_inv) {
    if (!transforms || !transforms_inv || mesh_count_ <= 0) throw Error("move_instances without a scene", PTB_E_BADARG);
    const size_t n = size_t(mesh_count_) * 12;
    if (!moved_xf_) moved_xf_ = new float[2 * n];
    std::memcpy(moved_xf_, transforms, n * sizeof(float));
    std::memcpy(moved_xf_ + n, transforms_inv, n * sizeof(float));
    invalidated_scene = true;
}

void Pathtracer::render() {
    if (!ctx_) throw Error("render before cuda_init", PTB_E_STATE);
    check(ptb_render(ctx_, sample_index), "ptb_render");
}

void Pathtracer::set_pixel_query(int x, int y) {
    if (x < 0 || y < 0 || x >= screen_width || y >= screen_height) return;
    check(ptb_set_pixel_query(ctx_, x, y), "ptb_set_pixel_query");
    pixel_query.pixel_index = x + y * screen_pitch; pixel_query.mesh_id = -1; pixel_query.triangle_id = -1;
    pixel_query_status = PixelQueryStatus::PENDING;
}

void Pathtracer::synchronize() { check(ptb_sync(ctx_), "ptb_sync"); }

const void* Pathtracer::get_display(int* pitch) {
    void* p = nullptr; check(ptb_get_display(ctx_, &p, pitch), "ptb_get_display"); return p;
}
const void* Pathtracer::get_aov(AOVType t, bool accumulated, int* pitch) {
    void* p = nullptr; check(ptb_get_aov(ctx_, int(t), accumulated ? 1 : 0, &p, pitch), "ptb_get_aov"); return p;
}
void Pathtracer::download_display(float* dst) { check(ptb_download(ctx_, -1, 1, dst), "ptb_download"); }
void Pathtracer::download_aov(AOVType t, bool accumulated, float* dst) { check(ptb_download(ctx_, int(t), accumulated ? 1 : 0, dst), "ptb_download"); }

}  // namespace ptb
