// rt_scene.cuh — device-side scene layout and the closest-hit traversal shared by every kernel.
//
// HBM layout, all arrays 16-byte aligned and read with 128-bit loads:
//   nodes    : 64 B per inner node (4 x float4), BOTH child boxes stored in the parent, so one
//              node visit is one 64-byte fetch (the reference re-reads the parent and then two
//              48-byte children: 144 B per visit, compute.glsl:425,443-444)
//                n0 = (L.lo.x, L.hi.x, L.lo.y, L.hi.y)   n1 = (R.lo.x, R.hi.x, R.lo.y, R.hi.y)
//                n2 = (L.lo.z, L.hi.z, R.lo.z, R.hi.z)   n3 = (left, right, leftCount, rightCount) as int
//              child >= 0: inner node index; child < 0: leaf, first sorted slot = ~child
//   tri_geom : 48 B per sorted triangle (3 x float4) = a, e0 = b-a, e1 = c-a, N = cross(e0,e1),
//              precomputed with the very operations compute.glsl:307-309 performs per test, so the
//              per-ray arithmetic (and every bit of dst,u,v) is unchanged while the UVs and material
//              index no longer travel through the intersection loop
//   tri_shade: 32 B per sorted triangle (2 x float4) = aTex,bTex,cTex, materialIndex, original index
//   tri_orig : 4 B per sorted triangle, the index in the caller's array (tie-break + reported id)
#pragma once
#include "rt_math.cuh"

namespace rt {

struct SceneView {
    const float4* __restrict__ nodes;      // 4 per inner node
    const float4* __restrict__ tri_geom;   // 3 per sorted triangle
    const float4* __restrict__ tri_shade;  // 2 per sorted triangle
    const int32_t* __restrict__ tri_orig;
    const float4* __restrict__ materials;  // 6 float4 per material (the 96-byte reference struct)
    const uint8_t* tex_px[5];
    int32_t tex_w[5], tex_h[5], tex_ch[5];
    int32_t num_tris;
    int32_t num_materials;
    int32_t root_is_leaf;  // scenes with a single triangle have no inner node
};

struct HitRec {
    float t;       // 1e38f = miss
    float u, v;
    int32_t slot;  // sorted slot, -1 = miss
};

constexpr int kStackSize = 64;
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
constexpr int kMaxLeafTris = 16;
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

    // IEEE division: a zero component gives +-inf; NaNs (inf*0) drop out of fminf/fmaxf
    const float idx = 1.0f / d.x, idy = 1.0f / d.y, idz = 1.0f / d.z;
    const float kWiden = 1.000001f;  // ~8 ulp: rounding in the slab test can never cull a true hit

    int32_t stack[kStackSize];
    float tstack[kStackSize];
    int sp = 0;
    int32_t cur = sc.root_is_leaf ? pack_leaf(0, sc.num_tris) : 0;

    for (;;) {
        if (cur >= 0) {
            const float4 n0 = __ldg(&sc.nodes[4 * cur + 0]);
            const float4 n1 = __ldg(&sc.nodes[4 * cur + 1]);
            const float4 n2 = __ldg(&sc.nodes[4 * cur + 2]);
            const float4 n3 = __ldg(&sc.nodes[4 * cur + 3]);
            if (COUNT) nodeVisits++;
            // slabs as (plane - origin) * inv: the subtraction is exact or nearly so, which keeps
            // the test meaningful for rays almost parallel to a slab (an fma of two huge products
            // would cancel catastrophically there)
            const float lx0 = (n0.x - o.x) * idx, lx1 = (n0.y - o.x) * idx;
            const float ly0 = (n0.z - o.y) * idy, ly1 = (n0.w - o.y) * idy;
            const float lz0 = (n2.x - o.z) * idz, lz1 = (n2.y - o.z) * idz;
            const float rx0 = (n1.x - o.x) * idx, rx1 = (n1.y - o.x) * idx;
            const float ry0 = (n1.z - o.y) * idy, ry1 = (n1.w - o.y) * idy;
            const float rz0 = (n2.z - o.z) * idz, rz1 = (n2.w - o.z) * idz;
            const float lNear = fmaxf(fmaxf(fminf(lx0, lx1), fminf(ly0, ly1)), fmaxf(fminf(lz0, lz1), 0.0f));
            const float rNear = fmaxf(fmaxf(fminf(rx0, rx1), fminf(ry0, ry1)), fmaxf(fminf(rz0, rz1), 0.0f));
            const float lFar = fminf(fminf(fmaxf(lx0, lx1), fmaxf(ly0, ly1)), fminf(fmaxf(lz0, lz1), best.t)) * kWiden;
            const float rFar = fminf(fminf(fmaxf(rx0, rx1), fmaxf(ry0, ry1)), fminf(fmaxf(rz0, rz1), best.t)) * kWiden;
            const bool hitL = lNear <= lFar;
            const bool hitR = rNear <= rFar;
            const int32_t cl = __float_as_int(n3.x), cr = __float_as_int(n3.y);
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
                const float4 g0 = __ldg(&sc.tri_geom[3 * s + 0]);
                const float4 g1 = __ldg(&sc.tri_geom[3 * s + 1]);
                const float4 g2 = __ldg(&sc.tri_geom[3 * s + 2]);
                if (COUNT) triTests++;
                float dst, u, v;
                if (ray_triangle(o, d, v3(g0.x, g0.y, g0.z), v3(g0.w, g1.x, g1.y), v3(g1.z, g1.w, g2.x),
                                 v3(g2.y, g2.z, g2.w), dst, u, v)) {
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
