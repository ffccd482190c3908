// rt_shell.cpp — headless host shell: the reference's main() + keyBoardInput() + screenshot()
// surface (RayTracing/src/rayTracing.cpp:1210-1443, :333-371, :124-283) without GLFW/GL.
//
// The window, the folder dialog and the GL plumbing are gone; everything else keeps its name,
// default and meaning: the scene comes from a model file plus an optional container
// (rayTracing.cpp:1285-1291), the five fixed materials are appended (:1268-1283), the camera starts
// at the reference values (:82-89), key events are replayed with the reference's 120 Hz frame step
// (:78-79) through the same key → action table (W/A/S/D, Shift+W/S, z/Z defocus, x/X rays per
// pixel, Ctrl+S screenshot), each interactive frame is one rt_render_frame with the preview
// uniforms (:1386-1400), and Ctrl+S runs rt_screenshot and writes Images/test.png (:261-264).
//
//   rt_shell [--folder RayTracing/Data/<model> | --model file.rtsc | --synthetic sphere:N] [--container none|classic|cornell|mirror|sidelit|sky]
//            [--width W --height H] [--camera x,y,z] [--keys "W:0.5,A:0.25,z:1,X:0.5"]
//            [--frames F --spp S --bounces B --env 0|1] [--rng pcg|philox] [--device D]
//            [--out RayTracing/Images/test.png] [--preview preview.png]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"
#include "rt_host.h"

static void die(const char* what, const char* detail) {
    fprintf(stderr, "rt_shell: %s: %s\n", what, detail ? detail : "");
    exit(1);
}
#define RT(call, ctx)                                         \
    do {                                                      \
        if ((call) != RT_OK) die(#call, rt_last_error(ctx));  \
    } while (0)

struct KeyEvent {
    char key;      // W A S D (move), U/J = Shift+W / Shift+S (up/down), z Z (defocus), x X (spp)
    float seconds;
};

int main(int argc, char** argv) {
    rth_defaults d;
    rth_get_defaults(&d);
    std::string folder, model, synthetic, container = "none", keys, out = "test.png", preview;
    int W = d.scr_width, H = d.scr_height, device = 0;
    int frames = d.screenshot_frames, spp = d.screenshot_rays_per_pixel, bounces = d.screenshot_max_bounce_count;
    int env = d.screenshot_environmental_light, rng = RT_RNG_PHILOX;
    float cam[3] = {d.camera_pos[0], d.camera_pos[1], d.camera_pos[2]};
    for (int i = 1; i < argc; i++) {
        auto arg = [&](const char* n) { return strcmp(argv[i], n) == 0 && i + 1 < argc; };
        if (arg("--folder")) folder = argv[++i];  // what the folder dialog returns in the reference (rayTracing.cpp:1249)
        else if (arg("--model")) model = argv[++i];
        else if (arg("--synthetic")) synthetic = argv[++i];
        else if (arg("--container")) container = argv[++i];
        else if (arg("--width")) W = atoi(argv[++i]);
        else if (arg("--height")) H = atoi(argv[++i]);
        else if (arg("--camera")) sscanf(argv[++i], "%f,%f,%f", &cam[0], &cam[1], &cam[2]);
        else if (arg("--keys")) keys = argv[++i];
        else if (arg("--frames")) frames = atoi(argv[++i]);
        else if (arg("--spp")) spp = atoi(argv[++i]);
        else if (arg("--bounces")) bounces = atoi(argv[++i]);
        else if (arg("--env")) env = atoi(argv[++i]);
        else if (arg("--rng")) rng = strcmp(argv[++i], "pcg") == 0 ? RT_RNG_REF_PCG : RT_RNG_PHILOX;
        else if (arg("--device")) device = atoi(argv[++i]);
        else if (arg("--out")) out = argv[++i];
        else if (arg("--preview")) preview = argv[++i];
        else die("unknown argument", argv[i]);
    }

    // ---- scene assembly (main(), rayTracing.cpp:1258-1291)
    rth_scene* scene = rth_scene_create();
    if (!folder.empty()) {
        printf("Loading model, please wait...\n");
        if (rth_load_model_folder(scene, folder.c_str()) != RT_OK) die("cannot load model folder", rth_last_error());
    } else if (!model.empty()) {
        if (rth_scene_load(scene, model.c_str()) != RT_OK) die("cannot load model", rth_last_error());
    } else if (!synthetic.empty()) {
        int n = 64;
        sscanf(synthetic.c_str(), "sphere:%d", &n);
        rth_set_procedural_texture(scene, 0, 1024);
        const int mat = rth_add_textured(scene, 0);
        const float c[3] = {0, 0, 0};
        rth_add_displaced_sphere(scene, n, c, 3.0f, 0.05f, mat);
    }
    const int red = rth_add_fixed_materials(scene);
    const int green = red + 1, wall = red + 2, light = red + 3, mirror = red + 4;
    if (container == "classic") rth_create_classic_cornell_box(scene, 10.0f, red, green, wall, light);
    else if (container == "cornell") rth_add_cornell_box(scene, d.cornell_light_size, d.cornell_padding, light, 1);
    else if (container == "mirror") rth_add_mirror_cornell_box(scene, d.cornell_light_size, d.cornell_padding, light, mirror);
    else if (container == "sidelit") rth_add_side_lit_cornell_box(scene, d.cornell_light_size, d.cornell_padding, light, wall, 1);
    else if (container == "sky") rth_add_sky_light_plane(scene, light);
    else if (container != "none") die("unknown container", container.c_str());
    printf("%lld triangles loaded\n", (long long)rth_scene_triangle_count(scene));

    // ---- upload + BVH (rayTracing.cpp:1293, :1323-1325)
    rt_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.device = device;
    cfg.rng_mode = rng;
    cfg.world_size = 1;
    rt_ctx* ctx = nullptr;
    if (rt_create(&ctx, &cfg) != RT_OK) die("rt_create", rt_last_error(nullptr));
    RT(rt_scene_set_triangles(ctx, rth_scene_triangles(scene), rth_scene_triangle_count(scene)), ctx);
    RT(rt_scene_set_materials(ctx, rth_scene_materials(scene), rth_scene_material_count(scene)), ctx);
    for (int t = 0; t < rth_scene_texture_count(scene); t++) {
        int tw, th, tc;
        const uint8_t* px = rth_scene_texture(scene, t, &tw, &th, &tc);
        RT(rt_scene_set_texture(ctx, t, px, tw, th, tc), ctx);
    }
    printf("Building BVH...\n");
    RT(rt_scene_build(ctx), ctx);
    printf("Built BVH.\n");

    // ---- camera + replayed input (rayTracing.cpp:1337, :1358-1419)
    rth_camera camera;
    rth_camera_init(&camera, W, H, d.max_speed, cam, d.hfov, d.pitch, d.yaw, d.focus_distance, d.defocus_angle, d.zoom);
    float numRaysPerPixel = d.num_rays_per_pixel;
    std::vector<KeyEvent> events;
    for (size_t p = 0; p < keys.size();) {
        KeyEvent e{keys[p], 0.0f};
        size_t colon = keys.find(':', p);
        if (colon == std::string::npos) break;
        e.seconds = (float)atof(keys.c_str() + colon + 1);
        events.push_back(e);
        size_t comma = keys.find(',', colon);
        if (comma == std::string::npos) break;
        p = comma + 1;
    }
    const float SPF = 1.0f / 120.0f;  // rayTracing.cpp:78-79
    uint32_t frameIndex = 0;
    rt_uniforms u;
    for (const KeyEvent& e : events) {
        for (float t = 0.0f; t < e.seconds; t += SPF) {
            uint8_t bits = 0;  // keyBoardInput(), rayTracing.cpp:333-371
            switch (e.key) {
                case 'W': bits |= RTH_FORWARD; break;
                case 'S': bits |= RTH_BACKWARD; break;
                case 'A': bits |= RTH_LEFT; break;
                case 'D': bits |= RTH_RIGHT; break;
                case 'U': bits |= RTH_UP; break;
                case 'J': bits |= RTH_DOWN; break;
                case 'z': bits |= RTH_DEFOCUS_UP; break;
                case 'Z': bits |= RTH_DEFOCUS_DOWN; break;
                case 'x': numRaysPerPixel = rth_adjust_rays_per_pixel(numRaysPerPixel, 1, SPF); break;
                case 'X': numRaysPerPixel = rth_adjust_rays_per_pixel(numRaysPerPixel, 0, SPF); break;
                default: break;
            }
            rth_camera_keyboard(&camera, bits, SPF);
            rth_fill_interactive_uniforms(scene, &camera, numRaysPerPixel, frameIndex++, &u);
            RT(rt_render_frame(ctx, &u), ctx);  // the interactive frame (preview shading)
        }
    }
    if (!preview.empty()) {
        rth_fill_interactive_uniforms(scene, &camera, numRaysPerPixel, frameIndex++, &u);
        RT(rt_render_frame(ctx, &u), ctx);
        std::vector<float> img((size_t)W * H * 4);
        RT(rt_read_frame_rgba32f(ctx, img.data()), ctx);
        std::vector<uint8_t> rgb((size_t)W * H * 3);
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                for (int c = 0; c < 3; c++) {
                    float v = img[((size_t)y * W + x) * 4 + c];
                    v = v < 0 ? 0 : (v > 1 ? 1 : v);
                    rgb[((size_t)(H - 1 - y) * W + x) * 3 + c] = (uint8_t)(v * 255.0f + 0.5f);
                }
        if (rth_write_png(preview.c_str(), W, H, 3, rgb.data()) != RT_OK) die("preview", rth_last_error());
    }

    // ---- Ctrl+S (screenshot(), rayTracing.cpp:124-283)
    printf("Performing Path Tracing, this will take a very long time and slow down your computer.\n");
    printf("Image dimensions: %dx%d\n", W, H);
    rth_fill_screenshot_uniforms(scene, &camera, &u);
    u.numRaysPerPixel = spp;
    u.maxBounceCount = bounces;
    u.environmentalLight = env;
    std::vector<uint8_t> pixels((size_t)W * H * 3);
    const auto t0 = std::chrono::steady_clock::now();
    RT(rt_screenshot(ctx, &u, frames, pixels.data()), ctx);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rth_write_png(out.c_str(), W, H, 3, pixels.data()) != RT_OK)
        fprintf(stderr, "Failed to write PNG file: %s\n", out.c_str());
    else
        printf("Screenshot saved to: %s\n", out.c_str());
    rt_counters c;
    RT(rt_get_counters(ctx, &c), ctx);
    printf("Total render time: %g minutes.  %.1f Mrays/s (%llu segments)\n", secs / 60.0, c.segments / secs * 1e-6,
           (unsigned long long)c.segments);
    rt_destroy(ctx);
    rth_scene_destroy(scene);
    return 0;
}
