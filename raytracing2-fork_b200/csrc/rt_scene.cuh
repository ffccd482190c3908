// rt_scene.cuh — device-side scene layout and the closest-hit traversal shared by every kernel.
//
// HBM layout.  ncu showed the traversal kernel bound by the L1TEX data pipe
// (l1tex__data_pipe_lsu_wavefronts 89-95 % of peak): a gather where every lane reads its own node is
// served at ~16 bytes per lane per wavefront, so the only lever is BYTES PER VISIT.
//   nodes    : 32 B per inner node = ONE 256-bit load (LDG.E.256 on sm_100a).  Both child boxes
//              live in the parent, quantised to 16 bits per plane on a global grid over the padded
//              scene box and rounded OUTWARD (the boxes only prune, so a larger box is always
//              safe; triangle tests stay exact binary32):
//                w0..w2 = left  child (lo.x | hi.x << 16, lo.y | hi.y << 16, lo.z | hi.z << 16)
//                w3..w5 = right child, w6 = left, w7 = right
//              child >= 0: inner node index; child < 0: leaf, ~child = first | (count-1) << 27.
//              (The reference re-reads the 48-byte parent and then two 48-byte children per visit:
//              144 B, compute.glsl:425,443-444.)
//   nodes4   : 64 B per 4-wide node (two 256-bit loads), collapsed from the binary tree by k_collapse4: four child
//              boxes w0..w11 (three words each, same quantisation), four children w12..w15.  k_extend walks these by
//              default; the binary nodes serve the per-thread hooks, RT_BVH_WIDTH=2 and the deep-tree fallback.
//   tri_geom : 64 B stride per sorted triangle; traversal reads the first 48 B = a, e0 = b-a,
//              e1 = c-a, N = cross(e0,e1), precomputed with the very operations
//              compute.glsl:307-309 performs per test, so every bit of dst,u,v is unchanged
//   tri_shade: 32 B per sorted triangle = aTex,bTex,cTex, materialIndex, original index
//   tri_orig : 4 B per sorted triangle, the index in the caller's array (tie-break + reported id)
#pragma once
#include "rt_math.cuh"

namespace rt {

struct SceneView {
    const uint4* __restrict__ nodes;       // 2 per inner node (32 B, 32-byte aligned)
    const uint4* __restrict__ nodes4;      // 4 per 4-wide node (64 B), or NULL (RT_BVH_WIDTH=4 builds them for k_extend)
    const float4* __restrict__ tri_geom;   // 4 per sorted triangle (64 B, 32-byte aligned)
    const float4* __restrict__ tri_shade;  // 2 per sorted triangle
    const int32_t* __restrict__ tri_orig;
    const float4* __restrict__ materials;  // 6 float4 per material (the 96-byte reference struct)
    const uint8_t* tex_px[5];
    int32_t tex_w[5], tex_h[5], tex_ch[5];
    int32_t num_tris;
    int32_t num_nodes4;    // 4-wide nodes behind `nodes4`
    int32_t num_materials;
    int32_t root_is_leaf;  // scenes with a single triangle have no inner node
    float grid_lo[3];      // world position of quantised coordinate 0
    float grid_inv[3];     // cells per world unit (65535 cells span the padded scene box)
};

struct HitRec {
    float t;       // 1e38f = miss
    float u, v;
    int32_t slot;  // sorted slot, -1 = miss
};

#ifndef RT_SLAB_FMA
#define RT_SLAB_FMA 1
#endif
// traversal stack entries per ray: 16 in shared memory (k_extend), the rest in local memory.  rt_scene_build checks the
// exact worst case of the tree it built against this (binary: depth; 4-wide: sum of children-1 along the deepest path).
constexpr int kStackSize = 256;
// outward rounding margin of the quantised planes, in grid cells (bvh_build.cu k_emit_nodes)
#if RT_SLAB_FMA
constexpr float kGuardCells = 0.0625f;
#else
constexpr float kGuardCells = 1e-3f;
#endif

// 256-bit read-only global load (sm_100a: LDG.E.ENL2.256.CONSTANT): one instruction, one L1TEX
// wavefront per lane for 32 bytes, where four 128-bit loads of a 64-byte record would cost four.
__device__ __forceinline__ void ldg256u(const void* p, uint32_t (&v)[8]) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
// exact uint16 -> float.  RT_DEQ_I2F: one I2F.U16 per plane (reads the half register directly, XU
// pipe); otherwise splice into the mantissa of 2^23 and subtract 2^23 (LOP3 + FADD, ALU/FMA pipes).
#ifndef RT_DEQ_I2F
#define RT_DEQ_I2F 1
#endif
#if RT_DEQ_I2F
__device__ __forceinline__ float q_lo(uint32_t w) { return (float)(uint16_t)(w & 0xffffu); }
__device__ __forceinline__ float q_hi(uint32_t w) { return (float)(uint16_t)(w >> 16); }
#else
__device__ __forceinline__ float q_lo(uint32_t w) { return __uint_as_float(0x4B000000u | (w & 0xffffu)) - 8388608.0f; }
__device__ __forceinline__ float q_hi(uint32_t w) { return __uint_as_float(0x4B000000u | (w >> 16)) - 8388608.0f; }
#endif

// The ray in grid coordinates: per-axis scaling of origin and direction leaves t unchanged.
struct GridRay {
#if RT_SLAB_FMA
    float ox, oy, oz;     // -(o - grid_lo) * grid_inv * i: the constant term of t = q * i + c
#else
    float ox, oy, oz;     // (o - grid_lo) * grid_inv
#endif
    float ix, iy, iz;     // 1 / (d * grid_inv), clamped to +-1e28 (RT_SLAB_FMA=0: a zero component gives +-inf)
};
__device__ __forceinline__ GridRay make_grid_ray(const SceneView& sc, V3 o, V3 d) {
    GridRay g;
    g.ox = (o.x - sc.grid_lo[0]) * sc.grid_inv[0];
    g.oy = (o.y - sc.grid_lo[1]) * sc.grid_inv[1];
    g.oz = (o.z - sc.grid_lo[2]) * sc.grid_inv[2];
    g.ix = 1.0f / (d.x * sc.grid_inv[0]);
    g.iy = 1.0f / (d.y * sc.grid_inv[1]);
    g.iz = 1.0f / (d.z * sc.grid_inv[2]);
#if RT_SLAB_FMA
    // finite slopes keep q*i + c free of inf - inf; a ray parallel to a slab then decides by the sign of
    // (q - o) * 1e28, which the builder's guard band keeps right (kGuardCells)
    g.ix = fminf(fmaxf(g.ix, -1e28f), 1e28f);
    g.iy = fminf(fmaxf(g.iy, -1e28f), 1e28f);
    g.iz = fminf(fmaxf(g.iz, -1e28f), 1e28f);
    g.ox = -g.ox * g.ix;
    g.oy = -g.oy * g.iy;
    g.oz = -g.oz * g.iz;
#endif
    return g;
}
__device__ __forceinline__ float slab_t(float q, float o, float i) {
#if RT_SLAB_FMA
    return __fmaf_rn(q, i, o);
#else
    return (q - o) * i;
#endif
}
// Two slab tests against the quantised child boxes of one node: t = fma(plane, inv, -origin * inv), one FFMA
// per plane where (plane - origin) * inv needs an FADD and an FMUL (12 fewer instructions per node visit,
// +2.5 % on config 2).  Conservativeness: with o the origin in cells, the roundings of the grid transform,
// of the constant term, of the slope and of the fma itself displace a plane by at most
// 2^-24 * (6|o| + 3 * 65535) cells.  The builder rounds every plane outward by at least kGuardCells = 1/16
// cell (covers |o| <= 65535, i.e. any origin inside the grid: 0.035 cell) and the far side is widened by
// 8 ulp relative (covers origins outside it, where |plane - o| grows with |o|), so rounding can never cull
// a true hit.  Slopes are clamped to +-1e28: no inf - inf for rays parallel to a slab.
// (RT_SLAB_FMA=0 keeps the subtract-multiply form with a 1e-3 cell guard band.  Selecting near / far planes
// by direction sign with a byte permute instead of min / max was measured too: 3 more registers cost a
// resident block per SM and 1.5 %.)
// WIDEN = false drops the relative widening of the far side: it exists only for origins far outside the grid
// (|o| beyond ~2.1 grid extents, where the bound above exceeds the guard band); every bounce ray starts on a
// surface of the scene, i.e. inside the grid, and so does a camera that stands inside or near the scene box —
// k_extend is instantiated both ways and the host picks per launch (rt_api.cu: widen_needed).
constexpr float kWidenFar = 1.000001f;
template <bool WIDEN = true>
__device__ __forceinline__ void slab2(const uint32_t (&w)[8], const GridRay& g, float bestT, float& lNear,
                                      float& rNear, bool& hitL, bool& hitR) {
    const float lx0 = slab_t(q_lo(w[0]), g.ox, g.ix), lx1 = slab_t(q_hi(w[0]), g.ox, g.ix);
    const float ly0 = slab_t(q_lo(w[1]), g.oy, g.iy), ly1 = slab_t(q_hi(w[1]), g.oy, g.iy);
    const float lz0 = slab_t(q_lo(w[2]), g.oz, g.iz), lz1 = slab_t(q_hi(w[2]), g.oz, g.iz);
    const float rx0 = slab_t(q_lo(w[3]), g.ox, g.ix), rx1 = slab_t(q_hi(w[3]), g.ox, g.ix);
    const float ry0 = slab_t(q_lo(w[4]), g.oy, g.iy), ry1 = slab_t(q_hi(w[4]), g.oy, g.iy);
    const float rz0 = slab_t(q_lo(w[5]), g.oz, g.iz), rz1 = slab_t(q_hi(w[5]), g.oz, g.iz);
    lNear = fmaxf(fmaxf(fminf(lx0, lx1), fminf(ly0, ly1)), fmaxf(fminf(lz0, lz1), 0.0f));
    rNear = fmaxf(fmaxf(fminf(rx0, rx1), fminf(ry0, ry1)), fmaxf(fminf(rz0, rz1), 0.0f));
    float lFar = fminf(fminf(fmaxf(lx0, lx1), fmaxf(ly0, ly1)), fminf(fmaxf(lz0, lz1), bestT));
    float rFar = fminf(fminf(fmaxf(rx0, rx1), fmaxf(ry0, ry1)), fminf(fmaxf(rz0, rz1), bestT));
    if (WIDEN) { lFar *= kWidenFar; rFar *= kWidenFar; }
    hitL = lNear <= lFar;
    hitR = rNear <= rFar;
}
// One slab test against a quantised box (words lo | hi << 16 per axis): entry distance, or kSlabMiss.
constexpr float kSlabMiss = 3.0e38f;
template <bool WIDEN = true>
__device__ __forceinline__ float slab1(uint32_t wx, uint32_t wy, uint32_t wz, const GridRay& g, float bestT) {
    const float x0 = slab_t(q_lo(wx), g.ox, g.ix), x1 = slab_t(q_hi(wx), g.ox, g.ix);
    const float y0 = slab_t(q_lo(wy), g.oy, g.iy), y1 = slab_t(q_hi(wy), g.oy, g.iy);
    const float z0 = slab_t(q_lo(wz), g.oz, g.iz), z1 = slab_t(q_hi(wz), g.oz, g.iz);
    const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), bestT));
    if (WIDEN) tf *= kWidenFar;
    return tn <= tf ? tn : kSlabMiss;
}
// first 48 bytes of a triangle record: a, e0, e1, N
struct TriGeom {
    V3 a, e0, e1, N;
};
__device__ __forceinline__ TriGeom load_tri(const SceneView& sc, int32_t s) {
    float g[8];
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(g[0]), "=f"(g[1]), "=f"(g[2]), "=f"(g[3]), "=f"(g[4]), "=f"(g[5]), "=f"(g[6]), "=f"(g[7])
                 : "l"(sc.tri_geom + 4 * s));
    const float4 h = __ldg(sc.tri_geom + 4 * s + 2);
    TriGeom t;
    t.a = v3(g[0], g[1], g[2]);
    t.e0 = v3(g[3], g[4], g[5]);
    t.e1 = v3(g[6], g[7], h.x);
    t.N = v3(h.y, h.z, h.w);
    return t;
}
__device__ __forceinline__ void ldg256(const void* p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
constexpr float kMissT = 1e38f;

// compute.glsl:302-340 on the precomputed (a, e0, e1, N).  Returns true on a hit and the same
// dst/u/v bits the reference expression order produces.
__device__ __forceinline__ bool ray_triangle(V3 o, V3 d, V3 a, V3 e0, V3 e1, V3 N, float& dst,
                                             float& u, float& v) {
    const float det = -dot(d, N);
    if ((det < 1e-10f && det > -1e-10f) || det < 0.0f) return false;
    const float invDet = 1.0f / det;
    const V3 ao = o - a;
    dst = dot(ao, N) * invDet;
    if (dst <= 1e-6f) return false;
    const V3 dao = cross(d, ao);
    u = -dot(e1, dao) * invDet;
    v = dot(e0, dao) * invDet;
    if (u < 0.0f || v < 0.0f || 1.0f - u - v < 0.0f) return false;
    return true;
}

// Leaf children are packed into one negative int: child = ~(first | (count-1) << 27).
constexpr int kLeafCountShift = 27;
constexpr int32_t kLeafFirstMask = (1 << kLeafCountShift) - 1;
__host__ __device__ __forceinline__ int32_t pack_leaf(int32_t first, int32_t count) {
    return ~(first | ((count - 1) << kLeafCountShift));
}

// Closest hit = min dst, ties → lowest ORIGINAL triangle index.  The BVH only prunes:
// a subtree is skipped when its (conservatively widened) slab interval cannot contain a hit with
// t <= best, so the result equals the exhaustive minimum over all triangles regardless of the
// hierarchy or the visiting order.
template <bool COUNT>
__device__ __forceinline__ HitRec closest_hit(const SceneView& sc, V3 o, V3 d, uint32_t& nodeVisits,
                                              uint32_t& triTests) {
    HitRec best;
    best.t = kMissT;
    best.u = 0.0f;
    best.v = 0.0f;
    best.slot = -1;
    int32_t bestOrig = 0x7fffffff;
    if (sc.num_tris <= 0) return best;

    const GridRay g = make_grid_ray(sc, o, d);
    const float kWiden = 1.000001f;

    int32_t stack[kStackSize];
    float tstack[kStackSize];
    int sp = 0;
    int32_t cur = sc.root_is_leaf ? pack_leaf(0, sc.num_tris) : 0;

    for (;;) {
        if (cur >= 0) {
            uint32_t w[8];
            ldg256u(sc.nodes + 2 * cur, w);
            if (COUNT) nodeVisits++;
            float lNear, rNear;
            bool hitL, hitR;
            slab2(w, g, best.t, lNear, rNear, hitL, hitR);
            const int32_t cl = (int32_t)w[6], cr = (int32_t)w[7];
            if (hitL && hitR) {
                const bool leftFirst = lNear <= rNear;
                stack[sp] = leftFirst ? cr : cl;
                tstack[sp] = leftFirst ? rNear : lNear;
                sp++;
                cur = leftFirst ? cl : cr;
                continue;
            } else if (hitL) {
                cur = cl;
                continue;
            } else if (hitR) {
                cur = cr;
                continue;
            }
        } else {
            const int32_t packed = ~cur;
            const int32_t first = packed & kLeafFirstMask;
            const int32_t count = (packed >> kLeafCountShift) + 1;
            for (int32_t s = first; s < first + count; s++) {
                const TriGeom tg = load_tri(sc, s);
                if (COUNT) triTests++;
                float dst, u, v;
                if (ray_triangle(o, d, tg.a, tg.e0, tg.e1, tg.N, dst, u, v)) {
                    if (dst <= best.t && dst < kMissT) {
                        const int32_t orig = __ldg(&sc.tri_orig[s]);
                        if (dst < best.t || orig < bestOrig) {
                            best.t = dst;
                            best.u = u;
                            best.v = v;
                            best.slot = s;
                            bestOrig = orig;
                        }
                    }
                }
            }
        }
        // pop the next subtree that can still hold a hit with t <= best
        bool got = false;
        while (sp > 0) {
            --sp;
            if (tstack[sp] <= best.t * kWiden) {
                cur = stack[sp];
                got = true;
                break;
            }
        }
        if (!got) break;
    }
    return best;
}

}  // namespace rt
