/*
 * rt_host.h — C-ABI of the host-side scene assembly library (librt_host.so; pure C++, no CUDA).
 *
 * Mirrors the reference's host surface around the hot path so that a user of the reference finds
 * the same operations with the same names, argument meaning and arithmetic:
 *   containers   RayTracing/src/rayTracing.cpp:388-1118  (createClassicCornellBox, addCornellBox,
 *                addMirrorCornellBox, addSideLitCornellBox, addSkyLightPlane, addCube,
 *                createDiverseCornellBox) and the fixed materials of :1268-1283
 *   camera       RayTracing/Assets/headers/camera.h:38-227
 *   controls     RayTracing/src/rayTracing.cpp:333-371 (key bits → camera / spp / screenshot)
 * plus the synthetic meshes the benchmark configs name and PNG output
 * (the reference's Ctrl+S → Images/test.png, rayTracing.cpp:261-264).
 *
 * The arrays this library hands out are the wire structs of include/rt_b200.h, ready for
 * rt_scene_set_triangles / rt_scene_set_materials / rt_scene_set_texture.
 */
#ifndef RT_HOST_H
#define RT_HOST_H

#include "../../include/rt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rth_scene rth_scene;

rth_scene* rth_scene_create(void);   /* holds the `_default_` material (mesh.h:322-324) at index 0 */
void rth_scene_destroy(rth_scene* s);
const char* rth_last_error(void);
void rth_set_error(const char* msg);  /* used by the translation units of this library */

int64_t rth_scene_triangle_count(const rth_scene* s);
const rt_triangle* rth_scene_triangles(const rth_scene* s);
int32_t rth_scene_material_count(const rth_scene* s);
const rt_material* rth_scene_materials(const rth_scene* s);
int32_t rth_scene_texture_count(const rth_scene* s);
const uint8_t* rth_scene_texture(const rth_scene* s, int32_t slot, int32_t* w, int32_t* h, int32_t* ch);

/* model folder ----------------------------------------------------------------------------------
 * getTrianglesData_(folder, …) of mesh.h:279-613: first .obj of the folder, every .mtl, textures/
 * (PNG and baseline 4:4:4 JPEG decoded here, byte-identical to stb_image).  Fills triangles, the material table in the reference's std::map
 * order, and up to 5 textures (flipped vertically like stbi_set_flip_vertically_on_load(true)).
 * The scene must be fresh.  Follow with rth_add_fixed_materials + a container, as main() does. */
int rth_load_model_folder(rth_scene* s, const char* folder);
/* PNG → 8-bit pixels with stbi_load(…, 0) channel conventions (top-down).  out == NULL: sizes only. */
int rth_decode_png(const uint8_t* bytes, int64_t n, uint8_t* out, int64_t cap, int32_t* w, int32_t* h, int32_t* ch);
int rth_scene_replace_materials(rth_scene* s, const rt_material* mats, int32_t count);

/* materials ------------------------------------------------------------------------------------ */
int32_t rth_add_material(rth_scene* s, const rt_material* m); /* returns its index */
/* Material::make* (mesh.h:49-102) on a fresh Material(); returns the new index */
int32_t rth_add_diffuse(rth_scene* s, float r, float g, float b);
int32_t rth_add_light(rth_scene* s, float r, float g, float b, float strength);
int32_t rth_add_specular(rth_scene* s, float r, float g, float b, float sr, float sg, float sb,
                         float smoothness, float specular_probability);
int32_t rth_add_checker(rth_scene* s, float scale);
int32_t rth_add_glass(rth_scene* s, float r, float g, float b, float refractive_index);
int32_t rth_add_textured(rth_scene* s, int32_t texture_index);
/* red, green, wall, light(15.0), mirror — rayTracing.cpp:1268-1283.  Returns index of `red`. */
int32_t rth_add_fixed_materials(rth_scene* s);

/* geometry ------------------------------------------------------------------------------------- */
int rth_add_triangles(rth_scene* s, const rt_triangle* tris, int64_t n);
int rth_add_cube(rth_scene* s, const float center[3], const float size[3], const float rotation[3],
                 int32_t material);                                           /* :867-923  */
int rth_create_classic_cornell_box(rth_scene* s, float room_size, int32_t red, int32_t green,
                                   int32_t white, int32_t light);             /* :949-1041 */
int rth_create_diverse_cornell_box(rth_scene* s, float room_size, int32_t red, int32_t green,
                                   int32_t white, int32_t light, int32_t glass, int32_t mirror,
                                   int32_t checker, int32_t metal);           /* :1071-1118 */
int rth_add_cornell_box(rth_scene* s, float light_size, float pad, int32_t light,
                        int32_t light_enabled);                               /* :453-547  */
int rth_add_mirror_cornell_box(rth_scene* s, float light_size, float pad, int32_t light,
                               int32_t mirror);                               /* :569-664  */
int rth_add_side_lit_cornell_box(rth_scene* s, float light_size, float pad, int32_t light,
                                 int32_t wall, int32_t rotate);               /* :690-847  */
int rth_add_sky_light_plane(rth_scene* s, int32_t light);                     /* :388-432  */

/* Synthetic tessellated mesh of the benchmark configs: n×n quads = 2n² triangles,
 * r(θ,φ) = radius·(1 + amp·sin 8θ·sin 6φ), CCW outward, uv = (j/n, i/n). */
int rth_add_displaced_sphere(rth_scene* s, int32_t n, const float center[3], float radius, float amp,
                             int32_t material);
/* Procedural size×size RGB texture (gradient + checker) into `slot`. */
int rth_set_procedural_texture(rth_scene* s, int32_t slot, int32_t size);
int rth_set_texture(rth_scene* s, int32_t slot, const uint8_t* px, int32_t w, int32_t h, int32_t ch);

/* camera (camera.h) ---------------------------------------------------------------------------- */
typedef struct rth_camera {
    float scrWidth, scrHeight, aspectRatio;
    float mouseSensitivity;
    float position[3], right[3], up[3], front[3], worldUp[3];
    float pitch, yaw, speed;
    double lastX, lastY;
    int32_t firstMouse;
    float zoomSensitivity, zoom;
    float hfov;
    float viewportRight[3], viewportUp[3], viewportFront[3], pixelRight[3], pixelUp[3];
    float focusDistance;
    float defocusAngle, defocusSensitivity;
    float defocusDiskRight[3], defocusDiskUp[3];
} rth_camera;

enum { /* Camera::Movement, camera.h:88-97 */
    RTH_FORWARD = 0x80, RTH_BACKWARD = 0x40, RTH_LEFT = 0x20, RTH_RIGHT = 0x10,
    RTH_UP = 0x08, RTH_DOWN = 0x04, RTH_DEFOCUS_UP = 0x02, RTH_DEFOCUS_DOWN = 0x01
};

void rth_camera_init(rth_camera* c, int32_t width, int32_t height, float speed, const float pos[3],
                     float hfov, float pitch, float yaw, float focus_dist, float defocus_angle,
                     float zoom);                                             /* camera.h:99-118  */
void rth_camera_keyboard(rth_camera* c, uint8_t input_bits, float dt);        /* camera.h:121-147 */
void rth_camera_mouse(rth_camera* c, double xpos, double ypos);               /* camera.h:194-213 */
void rth_camera_scroll(rth_camera* c, float y_offset);                        /* camera.h:163-180 */
void rth_camera_update_uniforms(const rth_camera* c, rt_uniforms* u);         /* camera.h:182-192 */

/* Reference defaults (rayTracing.cpp:48-89). */
typedef struct rth_defaults {
    int32_t scr_width, scr_height;
    int32_t max_bounce_count;            /* 10 */
    float num_rays_per_pixel;            /* 5  */
    float rays_per_pixel_sensitivity;    /* 10 */
    int32_t basic_shading, basic_shading_shadow, basic_shading_environmental_light;
    float light_position[3];
    int32_t screenshot_basic_shading, screenshot_environmental_light, screenshot_max_bounce_count,
        screenshot_rays_per_pixel, screenshot_frames;
    float cornell_light_brightness, cornell_padding, cornell_light_size;
    float max_speed, hfov, pitch, yaw, focus_distance, defocus_angle, zoom;
    float camera_pos[3];
} rth_defaults;
void rth_get_defaults(rth_defaults* d);

/* Interactive-frame / screenshot uniform blocks exactly as main() and screenshot() fill them
 * (rayTracing.cpp:1386-1400 and :146-150). */
void rth_fill_interactive_uniforms(const rth_scene* s, const rth_camera* c, float num_rays_per_pixel,
                                   uint32_t frame_index, rt_uniforms* u);
void rth_fill_screenshot_uniforms(const rth_scene* s, const rth_camera* c, rt_uniforms* u);

/* x / X key handling (rayTracing.cpp:358-367): returns the new clamped rays-per-pixel */
float rth_adjust_rays_per_pixel(float current, int32_t increase, float dt);

/* output --------------------------------------------------------------------------------------- */
int rth_write_png(const char* path, int32_t w, int32_t h, int32_t channels, const uint8_t* pixels);

/* RTSC scene files (the exchange format of oracle/ref_harness.cpp): triangles + materials (+ textures) */
int rth_scene_save(const rth_scene* s, const char* path);
int rth_scene_load(rth_scene* s, const char* path);

#ifdef __cplusplus
}
#endif
#endif
