// rt_scene.cuh — device-side scene layout and the closest-hit traversal shared by every kernel.
//
// HBM layout, all arrays 32-byte aligned and read with 256-bit loads (LDG.E.256 on
// sm_100a).  ncu showed the traversal kernel bound by L1TEX wavefronts (l1tex data pipe 89 % busy):
// a gather where every lane reads its own node costs one wavefront per lane per load instruction,
// so the layouts minimise the NUMBER OF LOAD INSTRUCTIONS per visit: a node is two 32-byte loads,
// a triangle two.
//   nodes    : 64 B per inner node, BOTH child boxes in the parent (the reference re-reads the parent
//              and then two 48-byte children: 144 B per visit, compute.glsl:425,443-444)
//                v8 #0 = (L.lo.x, L.hi.x, L.lo.y, L.hi.y, L.lo.z, L.hi.z, left, right)   [ints as bits]
//                v8 #1 = (R.lo.x, R.hi.x, R.lo.y, R.hi.y, R.lo.z, R.hi.z, 0, 0)
//              child >= 0: inner node index; child < 0: leaf, ~child = first | (count-1) << 27
//   tri_geom : 64 B per sorted triangle = a, e0 = b-a, e1 = c-a, N = cross(e0,e1) (48 B, precomputed
//              with the very operations compute.glsl:307-309 performs per test, so every bit of
//              dst,u,v is unchanged), then original index, material index, 2 pad words
//   tri_shade: 32 B per sorted triangle = aTex,bTex,cTex, materialIndex, original index
//   tri_orig : 4 B per sorted triangle, the index in the caller's array (reported id)
#pragma once
#include "rt_math.cuh"

namespace rt {

struct SceneView {
    const float4* __restrict__ nodes;      // 4 per inner node (64 B, 32-byte aligned)
    const float4* __restrict__ tri_geom;   // 4 per sorted triangle (64 B, 32-byte aligned)
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

// 256-bit read-only global load (sm_100a: LDG.E.ENL2.256.CONSTANT): one instruction, one L1TEX
// wavefront per lane for 32 bytes, where four 128-bit loads of a 64-byte record would cost four.
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
            float nl[8], nr[8];
            ldg256(sc.nodes + 4 * cur, nl);
            ldg256(sc.nodes + 4 * cur + 2, nr);
            if (COUNT) nodeVisits++;
            // slabs as (plane - origin) * inv: the subtraction is exact or nearly so, which keeps
            // the test meaningful for rays almost parallel to a slab (an fma of two huge products
            // would cancel catastrophically there)
            const float lx0 = (nl[0] - o.x) * idx, lx1 = (nl[1] - o.x) * idx;
            const float ly0 = (nl[2] - o.y) * idy, ly1 = (nl[3] - o.y) * idy;
            const float lz0 = (nl[4] - o.z) * idz, lz1 = (nl[5] - o.z) * idz;
            const float rx0 = (nr[0] - o.x) * idx, rx1 = (nr[1] - o.x) * idx;
            const float ry0 = (nr[2] - o.y) * idy, ry1 = (nr[3] - o.y) * idy;
            const float rz0 = (nr[4] - o.z) * idz, rz1 = (nr[5] - o.z) * idz;
            const float lNear = fmaxf(fmaxf(fminf(lx0, lx1), fminf(ly0, ly1)), fmaxf(fminf(lz0, lz1), 0.0f));
            const float rNear = fmaxf(fmaxf(fminf(rx0, rx1), fminf(ry0, ry1)), fmaxf(fminf(rz0, rz1), 0.0f));
            const float lFar = fminf(fminf(fmaxf(lx0, lx1), fmaxf(ly0, ly1)), fminf(fmaxf(lz0, lz1), best.t)) * kWiden;
            const float rFar = fminf(fminf(fmaxf(rx0, rx1), fmaxf(ry0, ry1)), fminf(fmaxf(rz0, rz1), best.t)) * kWiden;
            const bool hitL = lNear <= lFar;
            const bool hitR = rNear <= rFar;
            const int32_t cl = __float_as_int(nl[6]), cr = __float_as_int(nl[7]);
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
                float g[8], h[8];
                ldg256(sc.tri_geom + 4 * s, g);
                ldg256(sc.tri_geom + 4 * s + 2, h);
                if (COUNT) triTests++;
                float dst, u, v;
                if (ray_triangle(o, d, v3(g[0], g[1], g[2]), v3(g[3], g[4], g[5]), v3(g[6], g[7], h[0]),
                                 v3(h[1], h[2], h[3]), dst, u, v)) {
                    if (dst <= best.t && dst < kMissT) {
                        const int32_t orig = __float_as_int(h[4]);
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
