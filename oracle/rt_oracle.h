/*
 * rt_oracle.h — CPU oracle for the path-tracing hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under raytracing2-fork_b200/ includes, links or calls it.
 *
 * It is a line-by-line CPU restatement of the reference's GLSL compute shader
 * (RayTracing/Assets/Shaders/compute.glsl) and of the host pieces either side of it
 * (BVH.h builder, camera.h uniform derivation, screenshot() accumulation in rayTracing.cpp).
 *
 * PARITY — what pins this restatement:
 *   (1) the reference's OWN shader source: oracle/glsl2cpp.py rewrites compute.glsl syntactically and
 *       oracle/ref_shader.cpp compiles it against the reference's vendored glm (oracle/_ref/libref_shader.so);
 *       on the cases of tests/scenes.py (path-traced frames with every material type, textures, defocus, sky,
 *       both preview variants) this oracle's RGBA32F frames are BIT-IDENTICAL to it
 *       (tests/test_refshader_cpu.py; outputs committed as tests/golden/refshader_images.npz);
 *   (2) the real reference host code compiled from /root/reference (oracle/_ref/ref_host): BVH builder,
 *       camera uniforms, scene containers, loader, the integer part of the RNG.
 * PARITY STILL UNPINNED: what only a GL driver supplies and the reference tree does not contain — the
 * precision of cos / sin / exp / acos / pow, FMA contraction, the texture unit's filter arithmetic and the
 * evaluation order of the two jitter draws on compute.glsl:688 (GLSL leaves it undefined).  No OpenGL exists
 * in the build image and the reference ships no tests or golden vectors, so for those this oracle DEFINES the
 * arithmetic: IEEE-754 binary32, round to nearest even, no contraction, left-to-right jitter draws, the
 * elementary functions and the bilinear filter of DESIGN.md §4 (within 5 ulp of the same shader run with
 * glibc's float functions).
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include "../include/rt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

typedef struct orc_counters {
    uint64_t segments;    /* calculateRayCollisionBVH calls */
    uint64_t paths;
    uint64_t node_visits; /* inner nodes whose two children were fetched */
    uint64_t tri_tests;
} orc_counters;

/* scene ------------------------------------------------------------------------------------ */
orc_scene* orc_scene_create(const rt_triangle* tris, int64_t n, const rt_material* mats, int32_t k);
void orc_scene_destroy(orc_scene* s);
int orc_scene_set_texture(orc_scene* s, int32_t slot, const uint8_t* pixels, int32_t w, int32_t h,
                          int32_t channels);
/* BVH.h:150-220 restated (incl. its size() quirk, NaN costs, depth cap 32, 1e-4 padding).
 * Permutes an internal copy of the triangles; original indices are kept for ids and ties. */
int orc_scene_build_bvh(orc_scene* s);
int64_t orc_scene_node_count(const orc_scene* s);
int orc_scene_get_nodes(const orc_scene* s, rt_ref_node* out);
/* permuted triangle array (as the reference would upload it) + original index of each slot */
int orc_scene_get_permuted(const orc_scene* s, rt_triangle* tris_out, int32_t* orig_index_out);

/* closest hit ------------------------------------------------------------------------------ */
/* use_bvh = 0: brute force over all triangles (the exhaustive definition: min dst, then min index);
 * use_bvh = 1: reference traversal order over the reference BVH with non-strict pruning. */
int orc_trace_rays(const orc_scene* s, const float* origins, const float* dirs, int64_t count,
                   int use_bvh, int32_t* tri_id, float* dst, float* bary_u, float* bary_v);
int orc_first_hit(const orc_scene* s, const rt_uniforms* u, int32_t mode, int32_t rng_mode,
                  int use_bvh, int threads, int32_t* tri_id, float* dst);

/* render ----------------------------------------------------------------------------------- */
/* One frame → RGBA32F, row 0 = bottom (compute.glsl:660-701).  Region [x0,x1)×[y0,y1) only
 * (pixels outside are left untouched); pass 0,0,W,H for the whole image. */
int orc_render_frame(const orc_scene* s, const rt_uniforms* u, int32_t rng_mode, int threads,
                     int32_t x0, int32_t y0, int32_t x1, int32_t y1, float* rgba32f,
                     orc_counters* counters);
/* screenshot() (rayTracing.cpp:184-259): frames × (render, quantise to u8, sum), /frames, trunc,
 * flip → RGB8 top-down.  frame_list = NULL → frames 0..frames-1; otherwise only the listed frame
 * indices are rendered and `sum_out` (W*H*3 u32, bottom-up, may be NULL) receives the partial sums
 * so a multi-rank reduce can be emulated. */
int orc_screenshot(const orc_scene* s, const rt_uniforms* u, int32_t frames, int32_t rng_mode,
                   int threads, const int32_t* frame_list, int32_t n_list, uint32_t* sum_out,
                   uint8_t* rgb8_topdown, orc_counters* counters);
/* finalise stage alone: sums (bottom-up) → RGB8 top-down */
void orc_finalize(const uint32_t* sums, int32_t w, int32_t h, int32_t frames, uint8_t* rgb8_topdown);

/* camera.h:99-192 restated */
void orc_camera_uniforms(int32_t width, int32_t height, const float pos[3], float hfov, float pitch,
                         float yaw, float focus_dist, float defocus_angle, float zoom,
                         rt_uniforms* out);

/* unit-test hooks --------------------------------------------------------------------------- */
float orc_random(uint32_t* state);               /* compute.glsl:148-154 */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float orc_philox_draw(uint32_t pixel, uint32_t frame, uint32_t sample, uint32_t bounce, uint32_t j);
float orc_cos01(float x);
float orc_sin01(float x);
float orc_exp(float x);
float orc_acos(float x);
float orc_pow_gamma(float x);                    /* pow(x, 1/2.2) */
void orc_tonemap_srgb(const float in[3], float out[3]);
void orc_env_light(const float dir[3], float out[3]);
void orc_sample_texture(const orc_scene* s, int32_t tex, float u, float v, float out[3]);
/* returns 1 on hit; semantics of compute.glsl:302-340 */
int orc_ray_triangle(const float o[3], const float d[3], const rt_triangle* t, float* dst, float* u,
                     float* v);
float orc_ray_bounds(const float o[3], const float d[3], const float bmin[3], const float bmax[3]);

#ifdef __cplusplus
}
#endif
#endif
