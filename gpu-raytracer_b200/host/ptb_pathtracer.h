// C++ facade over the C ABI of include/ptb.h with the reference's Integrator / Pathtracer entry points
// (Src/Renderer/Integrators/Integrator.h:56-296, Pathtracer.h:146-286): cuda_init / cuda_free / resize_init / resize_free /
// update / render / set_pixel_query, the public state sample_index, invalidated_*, screen_width / height / pitch, gpu_config,
// pixel_query.  A maintainer of the reference derives (or replaces) his Integrator with this class; INTEGRATION.md shows where.
//
// Differences forced by the boundary: the scene arrives as the flat ptb_scene description (the arrays Integrator::init_* upload,
// Integrator.cpp:21-304) instead of a Scene&, there is no GL frame-buffer handle (get_display() returns the device pointer the
// reference would map its GL surface to), and errors are thrown as ptb::Error instead of __debugbreak (CUDACall.h:9-22).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "ptb.h"

namespace ptb {

struct Error : std::runtime_error { int code; Error(const std::string& what, int c) : std::runtime_error(what), code(c) {} };

enum class AOVType { RADIANCE = 0, RADIANCE_DIRECT, RADIANCE_INDIRECT, ALBEDO, NORMAL, POSITION, COUNT };   // AOV.h:4-12
enum class PixelQueryStatus { INACTIVE, PENDING, OUTPUT_READY };                                            // Integrator.h:58-62
struct PixelQuery { int pixel_index = -1, mesh_id = -1, triangle_id = -1; };                                 // Common.h:112-117

// camera as the loader leaves it (Src/Renderer/Camera.h): position, rotation (3x3, columns = right / up / back), fov, lens
struct CameraDesc {
    float position[3] = { 0, 0, 0 };
    float rotation[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };     // row-major 3x3: world = R * local; the camera looks along local -z
    float fov = 1.4835298f;                                // 85 degrees, Scene.cpp:17
    float aperture_radius = 0.0f, focal_distance = 10.0f;
    float near_plane = 0.1f, far_plane = 300.0f;
};

class Integrator {
public:
    virtual ~Integrator() {}
    virtual void cuda_init(unsigned frame_buffer_handle, int screen_width, int screen_height) = 0;
    virtual void cuda_free() = 0;
    virtual void resize_free() = 0;
    virtual void resize_init(unsigned frame_buffer_handle, int width, int height) = 0;
    virtual void update(float delta) = 0;
    virtual void render() = 0;

    ptb_config gpu_config;
    int screen_width = 0, screen_height = 0, screen_pitch = 0, pixel_count = 0;
    int sample_index = 0;
    bool invalidated_scene = true, invalidated_camera = true, invalidated_gpu_config = true, invalidated_aovs = true;
    PixelQuery pixel_query;
    PixelQueryStatus pixel_query_status = PixelQueryStatus::INACTIVE;
};

class Pathtracer : public Integrator {
public:
    // `scene` and everything it points to must stay alive until cuda_init() returns (the arrays are copied to the device there)
    Pathtracer(const ptb_scene& scene, const CameraDesc& camera, int device = 0, int rank = 0, int world = 1, int band_rows = 8);
    ~Pathtracer() override;

    void cuda_init(unsigned frame_buffer_handle, int screen_width, int screen_height) override;   // Pathtracer.cpp:9-41 + Integrator::cuda_init
    void cuda_free() override;                                                                      // Pathtracer.cpp:43-74
    void resize_init(unsigned frame_buffer_handle, int width, int height) override;               // Pathtracer.cpp:255-301
    void resize_free() override;                                                                    // Pathtracer.cpp:303-314
    void update(float delta) override;                                                              // Integrator.cpp:432-528
    void render() override;                                                                         // Pathtracer.cpp:738-855
    void set_pixel_query(int x, int y);                                                             // Integrator.h:266-277
    // Moving instances (the role of Mesh::update + invalidated_scene + build_tlas, Integrator.cpp:399-430,441-452): new object-to-world
    // and world-to-object matrices, mesh_count x 12 floats each in table order; the next update() refits the TLAS on the device
    // (ptb_refit_instances) and restarts the accumulation.
    void move_instances(const float* transforms, const float* transforms_inv);

    void aov_enable(AOVType t)  { gpu_config.aov_mask |=  (1u << int(t)); invalidated_aovs = true; invalidated_gpu_config = true; }
    void aov_disable(AOVType t) { gpu_config.aov_mask &= ~(1u << int(t)); invalidated_aovs = true; invalidated_gpu_config = true; }
    bool aov_is_enabled(AOVType t) const { return (gpu_config.aov_mask & (1u << int(t))) != 0; }

    CameraDesc camera;                    // change it, then set invalidated_camera (Camera::update does that in the reference)
    // a camera block + view-projection matrix computed elsewhere (e.g. by the scene loader) instead of `camera`; nullptr = back to `camera`
    void use_camera_block(const ptb_camera* block, const float* view_projection);
    void synchronize();                   // blocks until the queued passes are done
    const void* get_display(int* pitch = nullptr);                           // device pointer, pitch x height float4
    const void* get_aov(AOVType t, bool accumulated, int* pitch = nullptr);  // device pointer (Integrator.h:247)
    void download_display(float* host_dst);                                  // pitch x height x float4
    void download_aov(AOVType t, bool accumulated, float* host_dst);
    ptb_ctx* context() { return ctx_; }

private:
    void check(int code, const char* what);
    void upload_camera();
    ptb_scene scene_;
    ptb_ctx* ctx_ = nullptr;
    int device_, rank_, world_, band_rows_;
    float view_projection_[16], view_projection_prev_[16];
    bool have_view_projection_ = false;
    bool have_block_ = false;
    ptb_camera block_;
    float block_vp_[16];
    float* moved_xf_ = nullptr;           // pending transforms of move_instances (2 x mesh_count x 12 floats), owned
    int mesh_count_ = 0;
};

}  // namespace ptb
