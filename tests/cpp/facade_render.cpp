// C++ end of tests/test_gpu_facade.py: reads a raw scene dump (scene.dump_raw), renders it through the compiled C++ facade
// (host/ptb_pathtracer.h: cuda_init / update / render / resize_init ...) and writes the radiance accumulator back as raw floats.
// usage: facade_render <scene.raw> <passes> <out.f32> [<width2> <height2> <camera2.raw> <out2.f32>]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "ptb_pathtracer.h"

struct Raw { std::map<std::string, std::vector<unsigned char>> items; 
    bool load(const char* path) {
        FILE* f = fopen(path, "rb"); if (!f) return false;
        char magic[8]; int32_t count = 0;
        if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "PTBRAW1", 7) != 0 || fread(&count, 4, 1, f) != 1) { fclose(f); return false; }
        for (int i = 0; i < count; i++) {
            char name[32]; int64_t n = 0;
            if (fread(name, 1, 32, f) != 32 || fread(&n, 8, 1, f) != 1) { fclose(f); return false; }
            std::vector<unsigned char> d((size_t)n);
            if (n && fread(d.data(), 1, (size_t)n, f) != (size_t)n) { fclose(f); return false; }
            name[31] = 0; items[name] = std::move(d);
        }
        fclose(f); return true;
    }
    const void* ptr(const char* k) { auto it = items.find(k); return it == items.end() || it->second.empty() ? nullptr : it->second.data(); }
    size_t bytes(const char* k) { auto it = items.find(k); return it == items.end() ? 0 : it->second.size(); }
    int i32(const char* k) { int32_t v = 0; memcpy(&v, ptr(k), 4); return v; }
    float f32(const char* k) { float v = 0; memcpy(&v, ptr(k), 4); return v; }
};

static bool write_floats(const char* path, const std::vector<float>& v) {
    FILE* f = fopen(path, "wb"); if (!f) return false;
    bool ok = fwrite(v.data(), 4, v.size(), f) == v.size(); fclose(f); return ok;
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: facade_render scene.raw passes out.f32 [w2 h2 camera2.raw out2.f32]\n"); return 2; }
    Raw r; if (!r.load(argv[1])) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    const int passes = atoi(argv[2]);
    ptb_scene s; memset(&s, 0, sizeof(s));
    const int node_bytes = r.i32("bvh_kind") == 8 ? 80 : 32;
    s.triangles = r.ptr("triangles"); s.triangle_count = int(r.bytes("triangles") / 96);
    s.bvh_nodes = r.ptr("bvh_nodes"); s.bvh_node_count = int(r.bytes("bvh_nodes") / node_bytes); s.bvh_kind = r.i32("bvh_kind"); s.tlas_node_count = r.i32("tlas_node_count");
    s.mesh_count = int(r.bytes("mesh_bvh_root_indices") / 4);
    s.mesh_bvh_root_indices = (const int32_t*)r.ptr("mesh_bvh_root_indices"); s.mesh_material_ids = (const int32_t*)r.ptr("mesh_material_ids");
    s.mesh_transforms = (const float*)r.ptr("mesh_transforms"); s.mesh_transforms_inv = (const float*)r.ptr("mesh_transforms_inv"); s.mesh_transforms_prev = (const float*)r.ptr("mesh_transforms_prev");
    s.material_count = int(r.bytes("material_types")); s.material_types = (const int8_t*)r.ptr("material_types"); s.materials = (const float*)r.ptr("materials");
    s.medium_count = int(r.bytes("media") / 32); s.media = (const float*)r.ptr("media");
    std::vector<ptb_texture> tex((size_t)r.i32("texture_count"));
    std::vector<std::vector<const void*>> levels(tex.size());
    for (size_t i = 0; i < tex.size(); i++) {
        char k[32]; snprintf(k, sizeof(k), "tex%zu_meta", i);
        const int32_t* m = (const int32_t*)r.ptr(k);
        tex[i].format = m[0]; tex[i].width = m[1]; tex[i].height = m[2]; tex[i].num_levels = m[3]; memcpy(&tex[i].lod_bias, &m[4], 4);
        for (int l = 0; l < m[3]; l++) { snprintf(k, sizeof(k), "tex%zu_l%d", i, l); levels[i].push_back(r.ptr(k)); }
        tex[i].levels = levels[i].data();
    }
    s.texture_count = int(tex.size()); s.textures = tex.empty() ? nullptr : tex.data();
    s.sky = (const float*)r.ptr("sky"); s.sky_width = r.i32("sky_width"); s.sky_height = r.i32("sky_height"); s.sky_scale = r.f32("sky_scale");
    s.pmj_samples = (const float*)r.ptr("pmj"); s.blue_noise = (const unsigned char*)r.ptr("blue_noise");
    s.lights_total_weight = r.f32("lights_total_weight");
    s.light_triangle_count = int(r.bytes("light_triangle_indices") / 4);
    s.light_triangle_indices = (const int32_t*)r.ptr("light_triangle_indices"); s.light_triangle_cumulative_probability = (const float*)r.ptr("light_triangle_cdf");
    s.light_mesh_count = int(r.bytes("light_mesh_cdf") / 4); s.light_mesh_cumulative_probability = (const float*)r.ptr("light_mesh_cdf");
    s.light_mesh_triangle_span = (const int32_t*)r.ptr("light_mesh_triangle_span"); s.light_mesh_transform_indices = (const int32_t*)r.ptr("light_mesh_transform_indices");
    try {
        ptb::Pathtracer pt(s, ptb::CameraDesc());
        pt.gpu_config.num_bounces = r.i32("num_bounces");
        pt.cuda_init(0, r.i32("width"), r.i32("height"));
        ptb_camera block; memcpy(&block, r.ptr("camera"), sizeof(block));
        pt.use_camera_block(&block, (const float*)r.ptr("view_projection"));
        auto run = [&](const char* out) {
            for (int i = 0; i <= passes; i++) { pt.update(0.0f); pt.render(); }      // Src/Main.cpp:137-142: sample_index 0..passes
            pt.synchronize();
            if (pt.sample_index != passes) { fprintf(stderr, "sample_index %d after %d passes\n", pt.sample_index, passes + 1); return false; }
            std::vector<float> img((size_t)pt.screen_pitch * pt.screen_height * 4);
            pt.download_aov(ptb::AOVType::RADIANCE, true, img.data());
            return write_floats(out, img);
        };
        if (!run(argv[3])) return 1;
        if (argc >= 8) {                                     // Pathtracer::resize_free + resize_init, then the same loop on the new film
            Raw c2; if (!c2.load(argv[6])) { fprintf(stderr, "cannot read %s\n", argv[6]); return 2; }
            pt.resize_free();
            pt.resize_init(0, atoi(argv[4]), atoi(argv[5]));
            memcpy(&block, c2.ptr("camera"), sizeof(block));
            pt.use_camera_block(&block, (const float*)c2.ptr("view_projection"));
            if (!run(argv[7])) return 1;
        }
        pt.cuda_free();
    } catch (const ptb::Error& e) { fprintf(stderr, "facade error: %s\n", e.what()); return 1; }
    printf("FACADE-OK\n");
    return 0;
}
