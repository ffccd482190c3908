// ref_shader.cpp — runs the reference's OWN compute shader source on the CPU (TEST INFRASTRUCTURE ONLY).
//
// This file contains no reference code.  oracle/glsl2cpp.py rewrites
// /root/reference/RayTracing/Assets/Shaders/compute.glsl into oracle/_ref/compute_glsl.inc, which exists only
// while this file compiles (a syntactic rewrite: qualifiers dropped, `inout` → references, `f` suffixes on
// float literals, `.xyz` → `xyz()`, see the script), and this harness #includes it inside a namespace whose vocabulary is the reference's
// own vendored glm 0.9.9.7 (/root/reference/external/glm: vec2/3/4, dot, cross, normalize, reflect, mix,
// clamp, smoothstep, mod, …).  So every statement executed for a pixel — RNG, ray generation, the BVH
// walk over the reference's own node array, Möller–Trumbore, the material switch, Russian roulette,
// ACES + gamma — is the shader text itself, compiled with -O2 -ffp-contract=off (binary32, one rounding
// per operation).  What a GL driver would supply and the reference tree does not contain is supplied
// here and named:
//   * imageStore / imageSize / gl_GlobalInvocationID   the dispatch, one CPU thread per band of rows
//   * texture(sampler2D, vec2)                           GL_LINEAR / GL_REPEAT / unorm8 / no mips
//                                                        (textureClass.cpp:95-101), DESIGN.md §4.6 order
//   * cos / sin / exp / acos / pow                       REF_SPEC_MATH=1: the oracle's elementary functions
//                                                        (a driver's are unspecified);
//                                                        REF_SPEC_MATH=0: glm → libm (cosf, expf, powf …)
//   * vec3 / int, mod(float, int)                        GLSL's implicit int → float conversion
// Built by oracle/Makefile into oracle/_ref/libref_shader.so (+ libref_shader_libm.so); it pins the
// oracle (tests/test_refshader_cpu.py), mints tests/golden/refshader_images.npz for the GPU parity tests and is
// the CPU arm of bench.py (`cpu_baseline.kind = "reference"`).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#define GLM_FORCE_PURE          // no SIMD paths: the plain C++ expressions of func_geometric.inl
#define GLM_FORCE_XYZW_ONLY_OFF
#include <glm/glm.hpp>

#include "../include/rt_b200.h"

#ifndef REF_SPEC_MATH
#define REF_SPEC_MATH 1
#endif
#if REF_SPEC_MATH
extern "C" {
float orc_cos01(float x);
float orc_sin01(float x);
float orc_exp(float x);
float orc_acos(float x);
float orc_pow_gamma(float x);
}
#endif

namespace glsl {
using namespace glm;
typedef unsigned int uint;

// ---- what the GL side provides
struct image2D {
    float* px = nullptr;
    int w = 0, h = 0;
};
struct Texel {
    vec3 rgb;
};
struct sampler2D {
    const uint8_t* px = nullptr;
    int w = 0, h = 0, ch = 0;
};
struct InvocationID {
    uvec2 xy;
};
static thread_local InvocationID gl_GlobalInvocationID;

inline ivec2 imageSize(const image2D& img) { return ivec2(img.w, img.h); }
inline void imageStore(image2D& img, ivec2 p, vec4 v) {
    float* o = img.px + ((size_t)p.y * img.w + p.x) * 4;
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
// bilinear, REPEAT, texel centres at +0.5, unorm8 / 255 — the arithmetic of DESIGN.md §4.6
inline Texel texture(const sampler2D& t, vec2 uv) {
    Texel out;
    out.rgb = vec3(0.0f);
    if (t.w <= 0 || t.h <= 0 || !t.px) return out;
    float s = uv.x - std::floor(uv.x);
    float r = uv.y - std::floor(uv.y);
    if (!(s >= 0.0f && s <= 1.0f)) s = 0.0f;
    if (!(r >= 0.0f && r <= 1.0f)) r = 0.0f;
    const float fx = s * (float)t.w - 0.5f, fy = r * (float)t.h - 0.5f;
    const float flx = std::floor(fx), fly = std::floor(fy);
    const float ax = fx - flx, ay = fy - fly;
    auto wrap = [](int i, int n) { return ((i % n) + n) % n; };
    const int i0 = wrap((int)flx, t.w), i1 = wrap((int)flx + 1, t.w);
    const int j0 = wrap((int)fly, t.h), j1 = wrap((int)fly + 1, t.h);
    auto texel = [&](int i, int j) {
        const uint8_t* p = t.px + ((size_t)j * t.w + i) * t.ch;
        const float r8 = (float)p[0] / 255.0f;
        if (t.ch == 1) return vec3(r8, r8, r8);
        const float g8 = (float)p[1] / 255.0f;
        if (t.ch == 2) return vec3(r8, g8, 0.0f);
        return vec3(r8, g8, (float)p[2] / 255.0f);
    };
    const float w00 = (1.0f - ax) * (1.0f - ay), w10 = ax * (1.0f - ay), w01 = (1.0f - ax) * ay, w11 = ax * ay;
    out.rgb = ((texel(i0, j0) * w00 + texel(i1, j0) * w10) + texel(i0, j1) * w01) + texel(i1, j1) * w11;
    return out;
}

// ---- GLSL conveniences glm lacks
inline vec3 xyz(const vec3& v) { return v; }
inline vec3 xyz(const vec4& v) { return vec3(v.x, v.y, v.z); }
inline vec3 operator/(const vec3& v, int s) { return v / (float)s; }
inline float mod(float x, int y) { return glm::mod(x, (float)y); }

#if REF_SPEC_MATH
// the elementary functions a GL driver would bring, replaced by the spec'd ones
inline float cos(float x) { return orc_cos01(x); }
inline float sin(float x) { return orc_sin01(x); }
inline float exp(float x) { return orc_exp(x); }
inline float acos(float x) { return orc_acos(x); }
inline vec3 pow(const vec3& b, const vec3& e) {  // only use: toSRGB, pow(c, vec3(1/2.2))
    (void)e;
    return vec3(orc_pow_gamma(b.x), orc_pow_gamma(b.y), orc_pow_gamma(b.z));
}
#endif

#include "_ref/compute_glsl.inc"

}  // namespace glsl

// ------------------------------------------------------------------------------------------------ C API
namespace {
std::vector<glsl::Triangle> g_tris;
std::vector<glsl::Node> g_nodes;
std::vector<glsl::Material> g_mats;
std::vector<uint8_t> g_tex[5];
}  // namespace

extern "C" {

// triangles in the REFERENCE BVH's order (BVH.h permutes them), its node array, the material table
int refsh_set_scene(const rt_triangle* tris, int64_t n, const rt_ref_node* nodes, int64_t m, const rt_material* mats,
                    int32_t k) {
    using namespace glsl;
    g_tris.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        Triangle& t = g_tris[(size_t)i];
        t.a = vec3(tris[i].a[0], tris[i].a[1], tris[i].a[2]);
        t.b = vec3(tris[i].b[0], tris[i].b[1], tris[i].b[2]);
        t.c = vec3(tris[i].c[0], tris[i].c[1], tris[i].c[2]);
        t.aTex = vec2(tris[i].aTex[0], tris[i].aTex[1]);
        t.bTex = vec2(tris[i].bTex[0], tris[i].bTex[1]);
        t.cTex = vec2(tris[i].cTex[0], tris[i].cTex[1]);
        t.mtlIndex = tris[i].materialIndex;
        t.pad = 0;
    }
    g_nodes.resize((size_t)m);
    for (int64_t i = 0; i < m; i++) {
        Node& d = g_nodes[(size_t)i];
        d.bounds.bmin = vec3(nodes[i].bmin[0], nodes[i].bmin[1], nodes[i].bmin[2]);
        d.bounds.bmax = vec3(nodes[i].bmax[0], nodes[i].bmax[1], nodes[i].bmax[2]);
        d.bounds.pad0 = d.bounds.pad1 = 0.0f;
        d.triangleIndex = nodes[i].triangleIndex;
        d.triangleCount = nodes[i].triangleCount;
        d.childIndex = nodes[i].childIndex;
        d.pad0 = 0;
    }
    g_mats.resize((size_t)k);
    for (int32_t i = 0; i < k; i++) {
        Material& d = g_mats[(size_t)i];
        const rt_material& s = mats[i];
        d.color = vec4(s.color[0], s.color[1], s.color[2], s.color[3]);
        d.specularColor = vec4(s.specularColor[0], s.specularColor[1], s.specularColor[2], s.specularColor[3]);
        d.emissionColor = vec4(s.emissionColor[0], s.emissionColor[1], s.emissionColor[2], s.emissionColor[3]);
        d.textureIndex = s.textureIndex;
        d.emissionStrength = s.emissionStrength;
        d.smoothness = s.smoothness;
        d.specularProbability = s.specularProbability;
        d.checkerScale = s.checkerScale;
        d.refractiveIndex = s.refractiveIndex;
        d.materialType = s.materialType;
        d.index = s.index;
        d.isEdgeHighlight = s.isEdgeHighlight;
        d.pad1 = d.pad2 = d.pad3 = 0;
    }
    triangles = g_tris.data();
    allNodes = g_nodes.data();
    materials = g_mats.data();
    return 0;
}

int refsh_set_texture(int32_t slot, const uint8_t* px, int32_t w, int32_t h, int32_t ch) {
    if (slot < 0 || slot >= 5) return 1;
    g_tex[slot].assign(px, px + (size_t)w * h * ch);
    glsl::sampler2D s;
    s.px = g_tex[slot].data();
    s.w = w; s.h = h; s.ch = ch;
    glsl::sampler2D* dst[5] = {&glsl::texture0, &glsl::texture1, &glsl::texture2, &glsl::texture3, &glsl::texture4};
    *dst[slot] = s;
    return 0;
}

// UBO.Update + glDispatchCompute: one shader_main() per pixel of [x0,x1) x [y0,y1) into rgba32f (W*H*4 floats, row 0
// = bottom; pixels outside the region are left untouched)
int refsh_render_region(const rt_uniforms* u, float* rgba32f, int32_t x0, int32_t y0, int32_t x1, int32_t y1, int threads) {
    using namespace glsl;
    if (!u || !rgba32f || g_nodes.empty()) return 1;
    pad = u->pad; numTextures = u->numTextures; width = u->width; height = u->height;
    numSpheres = u->numSpheres; numTriangles = u->numTriangles;
    basicShading = u->basicShading != 0; basicShadingShadow = u->basicShadingShadow != 0;
    auto v4 = [](const float* p) { return vec4(p[0], p[1], p[2], p[3]); };
    basicShadingLightPosition = v4(u->basicShadingLightPosition);
    environmentalLight = u->environmentalLight != 0;
    maxBounceCount = u->maxBounceCount; numRaysPerPixel = u->numRaysPerPixel; frameIndex = u->frameIndex;
    cameraPos = v4(u->cameraPos); viewportRight = v4(u->viewportRight); viewportUp = v4(u->viewportUp);
    viewportFront = v4(u->viewportFront); pixelRight = v4(u->pixelRight); pixelUp = v4(u->pixelUp);
    defocusDiskRight = v4(u->defocusDiskRight); defocusDiskUp = v4(u->defocusDiskUp);
    imgOutput.px = rgba32f;
    imgOutput.w = (int)u->width;
    imgOutput.h = (int)u->height;
    const int W = (int)u->width, H = (int)u->height;
    if (x0 < 0) x0 = 0;
    if (y0 < 0) y0 = 0;
    if (x1 > W) x1 = W;
    if (y1 > H) y1 = H;
    if (x1 <= x0 || y1 <= y0) return 0;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > y1 - y0) threads = y1 - y0;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([=]() {
            for (int y = y0 + t; y < y1; y += threads)
                for (int x = x0; x < x1; x++) {
                    gl_GlobalInvocationID.xy = uvec2((unsigned)x, (unsigned)y);
                    shader_main();
                }
        });
    for (auto& th : pool) th.join();
    return 0;
}

int refsh_render(const rt_uniforms* u, float* rgba32f, int threads) {
    if (!u) return 1;
    return refsh_render_region(u, rgba32f, 0, 0, (int32_t)u->width, (int32_t)u->height, threads);
}

int refsh_spec_math(void) { return REF_SPEC_MATH; }

}  // extern "C"
