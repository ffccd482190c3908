// rt_oracle.cpp — CPU oracle (TEST INFRASTRUCTURE ONLY; see rt_oracle.h for the rules).
//
// Restates, function by function, the reference hot path.  Citations are relative to the reference
// tree: S = RayTracing/Assets/Shaders/compute.glsl, B = RayTracing/Assets/headers/BVH.h,
// C = RayTracing/Assets/headers/camera.h, M = RayTracing/Assets/headers/mesh.h,
// R = RayTracing/src/rayTracing.cpp, G = external/glm/detail/func_geometric.inl.
//
// Build with -ffp-contract=off (both the parity and the speed build): every expression below is
// evaluated in binary32 in exactly the order written.
#include "rt_oracle.h"

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------
// vec3 with glm 0.9.9.7 operation order (G:48-110)
struct V3 {
    float x, y, z;
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 v3(const float* p) { return V3{p[0], p[1], p[2]}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
inline V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
// G:52-53  dot3 = (a.x*b.x + a.y*b.y) + a.z*b.z
inline float dot(V3 a, V3 b) {
    float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
    return (tx + ty) + tz;
}
// G:74-77
inline V3 cross(V3 x, V3 y) {
    return V3{x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
}
inline float length(V3 v) { return sqrtf(dot(v, v)); }  // G:12
// G:88 + func_exponential.inl:138  normalize(v) = v * (1/sqrt(dot(v,v)))
inline V3 normalize(V3 v) { return v * (1.0f / sqrtf(dot(v, v))); }
// G:108  reflect(I,N) = I - N*dot(N,I)*2
inline V3 reflect(V3 I, V3 N) { return I - (N * dot(N, I)) * 2.0f; }
inline float gmin(float x, float y) { return (y < x) ? y : x; }  // glm::min
inline float gmax(float x, float y) { return (x < y) ? y : x; }  // glm::max
inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
// glm mix: x*(1-a) + y*a
inline V3 mix(V3 x, V3 y, float a) { return x * (1.0f - a) + y * a; }
inline float smoothstep(float e0, float e1, float x) {
    float t = gclamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline uint32_t f2u(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
inline float u2f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// ------------------------------------------------------------------------------------------------
// Elementary functions.  GLSL leaves these to the driver; the oracle fixes them as
// short binary32 kernels so that CPU and GPU can agree bit for bit.

// cos / sin on [0,1] (the only argument range on the path: angle = random(), S:163-164).
// Taylor in z = x*x, Horner, 7 / 6 terms.
float cos01(float x) {
    const float z = x * x;
    float p = 2.08767570e-09f;      // 1/12!
    p = p * z + -2.75573192e-07f;   // -1/10!
    p = p * z + 2.48015873e-05f;    // 1/8!
    p = p * z + -1.38888889e-03f;   // -1/6!
    p = p * z + 4.16666667e-02f;    // 1/4!
    p = p * z + -0.5f;
    p = p * z + 1.0f;
    return p;
}
float sin01(float x) {
    const float z = x * x;
    float p = -2.50521084e-08f;     // -1/11!
    p = p * z + 2.75573192e-06f;    // 1/9!
    p = p * z + -1.98412698e-04f;   // -1/7!
    p = p * z + 8.33333333e-03f;    // 1/5!
    p = p * z + -1.66666667e-01f;   // -1/3!
    p = p * z + 1.0f;
    return x * p;
}

// e^r for |r| <= 0.35, Taylor degree 7
inline float exp_poly(float r) {
    float p = 1.98412698e-04f;      // 1/7!
    p = p * r + 1.38888889e-03f;    // 1/6!
    p = p * r + 8.33333333e-03f;    // 1/5!
    p = p * r + 4.16666667e-02f;    // 1/4!
    p = p * r + 1.66666667e-01f;    // 1/3!
    p = p * r + 0.5f;
    p = p * r + 1.0f;
    p = p * r + 1.0f;
    return p;
}
inline float pow2i(float n) {  // 2^n for integral n in [-126, 127]
    return u2f((uint32_t)((int32_t)n + 127) << 23);
}
float exp_(float x) {
    if (!(x >= -87.0f)) return 0.0f;  // also NaN -> 0
    if (x > 88.0f) x = 88.0f;
    const float n = floorf(x * 1.44269504f + 0.5f);
    const float r = (x - n * 0.693145752f) - n * 1.42860677e-06f;  // ln2 hi / lo
    return exp_poly(r) * pow2i(n);
}

// acos on [-1,1]: the classic three-range rational kernel (as in FreeBSD msun e_acosf.c)
inline float asin_R(float z) {
    const float pS0 = 1.6666586697e-01f, pS1 = -4.2743422091e-02f, pS2 = -8.6563630030e-03f,
                qS1 = -7.0662963390e-01f;
    const float p = z * (pS0 + z * (pS1 + z * pS2));
    const float q = 1.0f + z * qS1;
    return p / q;
}
float acos_(float x) {
    const float pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
    if (x >= 1.0f) return 0.0f;
    if (x <= -1.0f) return 3.14159274f;
    if (x < 0.5f && x > -0.5f) {
        const float z = x * x;
        const float r = asin_R(z);
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    if (x < 0.0f) {
        const float z = (1.0f + x) * 0.5f;
        const float s = sqrtf(z);
        const float w = asin_R(z) * s - pio2_lo;
        return 2.0f * (pio2_hi - (s + w));
    }
    const float z = (1.0f - x) * 0.5f;
    const float s = sqrtf(z);
    const float df = u2f(f2u(s) & 0xfffff000u);
    const float c = (z - df * df) / (s + df);
    const float w = asin_R(z) * s + c;
    return 2.0f * (df + w);
}

// pow(x, 1/2.2) for x in [0,1] (S:656-658): exp2(log2(x) * (1/2.2))
float pow_gamma(float x) {
    if (!(x >= 1.17549435e-38f)) return 0.0f;  // 0, denormals, negatives and NaN -> 0
    const uint32_t bits = f2u(x);
    int32_t e = (int32_t)((bits >> 23) & 0xffu) - 127;
    float m = u2f((bits & 0x007fffffu) | 0x3f800000u);  // [1,2)
    if (m > 1.41421354f) {
        m = m * 0.5f;
        e += 1;
    }
    const float s = (m - 1.0f) / (m + 1.0f);
    const float z = s * s;
    float p = 9.09090909e-02f;      // 1/11
    p = p * z + 1.11111111e-01f;    // 1/9
    p = p * z + 1.42857143e-01f;    // 1/7
    p = p * z + 0.2f;
    p = p * z + 3.33333333e-01f;
    p = p * z + 1.0f;
    const float lnm = (s + s) * p;
    const float log2x = (float)e + lnm * 1.44269504f;
    const float y = log2x * 0.45454547f;  // 1/2.2 rounded to binary32
    const float n = floorf(y + 0.5f);
    if (n < -126.0f) return 0.0f;
    const float r = (y - n) * 0.693147182f;
    return exp_poly(r) * pow2i(n);
}

// ------------------------------------------------------------------------------------------------
// RNG.  S:148-154: the literal 4294967295.0 is a float in GLSL and rounds to 2^32.
inline float pcg_next(uint32_t& state) {
    state = state * 747796405u + 2891336453u;
    uint32_t result = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    result = (result >> 22u) ^ result;
    return (float)result / 4294967296.0f;
}

// Philox4x32-10 (Salmon et al., SC'11), the counter-based generator of the B200 path.
inline void philox(const uint32_t c[4], const uint32_t k[2], uint32_t out[4]) {
    uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], k0 = k[0], k1 = k[1];
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// One random stream per (pixel, frame, sample, bounce).  Draw j of the stream is word (j & 3) of
// philox(ctr = (j >> 2, bounce, sample, 'RT20'), key = (pixel, frame)).
struct Rng {
    int mode;
    uint32_t state;  // PCG
    uint32_t pixel, frame, sample, bounce, j;
    uint32_t cache[4];
    uint32_t cached_block;
    void stream(uint32_t b) {
        bounce = b;
        j = 0;
        cached_block = 0xffffffffu;
    }
    float next() {
        if (mode == RT_RNG_REF_PCG) return pcg_next(state);
        const uint32_t blk = j >> 2;
        if (blk != cached_block) {
            const uint32_t c[4] = {blk, bounce, sample, 0x52543230u};
            const uint32_t k[2] = {pixel, frame};
            philox(c, k, cache);
            cached_block = blk;
        }
        const uint32_t r = cache[j & 3u];
        j++;
        return (float)r / 4294967296.0f;
    }
    float range(float l, float r) { return l + (r - l) * next(); }  // S:156-159
};

// S:174-185
V3 randomDirection(Rng& rng) {
    for (int i = 0; i < 100; i++) {
        float x = rng.next() * 2.0f - 1.0f;
        float y = rng.next() * 2.0f - 1.0f;
        float z = rng.next() * 2.0f - 1.0f;
        if (length(v3(x, y, z)) < 1.0f) return normalize(v3(x, y, z));
    }
    return v3(0, 0, 0);
}

// ------------------------------------------------------------------------------------------------
struct Texture {
    int w = 0, h = 0, ch = 0;
    std::vector<uint8_t> px;
};

struct BvhTri {  // M:141-154
    V3 mn, mx, center;
};

struct Hit {
    bool didHit = false;
    float dst = 1e38f;
    int slot = -1;  // index into the (permuted) triangle array = reference triangleIndex
    int orig = -1;  // index into the caller's array
    float u = 0, v = 0;
    V3 normal{0, 0, 0};
    V3 point{0, 0, 0};
    int mtl = 0;
};

struct Counters {
    uint64_t segments = 0, paths = 0, node_visits = 0, tri_tests = 0;
};

}  // namespace

struct orc_scene {
    std::vector<rt_triangle> tris;  // permuted by the BVH build exactly as B:194-195 does
    std::vector<int32_t> orig;      // original index of each slot
    std::vector<rt_material> mats;
    std::vector<rt_ref_node> nodes;
    Texture tex[RT_MAX_TEXTURES];
    bool built = false;
};

namespace {

// S:302-340
inline bool rayTriangle(V3 o, V3 d, const rt_triangle& t, float& dst, float& u, float& v, V3& N) {
    const V3 a = v3(t.a), b = v3(t.b), c = v3(t.c);
    const V3 e0 = b - a;
    const V3 e1 = c - a;
    N = cross(e0, e1);
    const float det = -dot(d, N);
    if ((det < 1e-10f && det > -1e-10f) || det < 0) return false;
    const float invDet = 1.0f / det;
    const V3 ao = o - a;
    dst = dot(ao, N) * invDet;
    if (dst <= 1e-6f) return false;
    const V3 dao = cross(d, ao);
    u = -dot(e1, dao) * invDet;
    v = dot(e0, dao) * invDet;
    if (u < 0 || v < 0 || 1.0f - u - v < 0) return false;
    return true;
}

// S:370-408
inline bool isCloseToZero(float val) { return (val < 1e-6f) && (val > -1e-6f); }
inline float rayBounds(V3 o, V3 d, const float* bmin, const float* bmax) {
    float tMin = -1e32f, tMax = 1e32f;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    for (int i = 0; i < 3; i++) {
        if (!isCloseToZero(dd[i])) {
            float t0 = (bmin[i] - oo[i]) / dd[i];
            float t1 = (bmax[i] - oo[i]) / dd[i];
            if (t0 > t1) {
                float tmp = t0;
                t0 = t1;
                t1 = tmp;
            }
            if (tMin < t0) tMin = t0;
            if (tMax > t1) tMax = t1;
            if (tMin >= tMax || tMax < 0) return 1e38f;
        }
    }
    return tMin;
}

// The closest-hit rule of this project: min dst, ties → lowest ORIGINAL index.
// (S:433 keeps the first-encountered among equal dst, which depends on traversal order.)
inline void consider(Hit& best, const orc_scene& s, int slot, V3 o, V3 d, Counters& cn) {
    float dst, u, v;
    V3 N;
    cn.tri_tests++;
    if (!rayTriangle(o, d, s.tris[slot], dst, u, v, N)) return;
    const int orig = s.orig[slot];
    if (dst < best.dst || (dst == best.dst && best.didHit && orig < best.orig)) {
        best.didHit = true;
        best.dst = dst;
        best.slot = slot;
        best.orig = orig;
        best.u = u;
        best.v = v;
        best.normal = N;  // normalised lazily by finishHit
        best.mtl = s.tris[slot].materialIndex;
    }
}
inline void finishHit(Hit& h, V3 o, V3 d) {
    if (!h.didHit) return;
    h.point = o + d * h.dst;          // S:330
    h.normal = normalize(h.normal);   // S:331
}

Hit closestBrute(const orc_scene& s, V3 o, V3 d, Counters& cn) {
    Hit best;
    const int n = (int)s.tris.size();
    for (int i = 0; i < n; i++) consider(best, s, i, o, d, cn);
    finishHit(best, o, d);
    return best;
}

// S:410-460 with two deliberate changes: lexicographic (dst, index) update and
// NON-strict pruning (a child whose entry distance equals the best distance may still hold a
// lower-index tie), which together make the result independent of traversal order.
Hit closestBVH(const orc_scene& s, V3 o, V3 d, Counters& cn) {
    int stack[64];
    int sp = 0;
    stack[sp++] = 0;
    Hit best;
    while (sp > 0) {
        const rt_ref_node& node = s.nodes[stack[--sp]];
        if (node.childIndex == -1) {
            for (int i = node.triangleIndex; i < node.triangleIndex + node.triangleCount; i++)
                consider(best, s, i, o, d, cn);
        } else {
            cn.node_visits++;
            const int ia = node.childIndex, ib = node.childIndex + 1;
            const rt_ref_node& A = s.nodes[ia];
            const rt_ref_node& Bn = s.nodes[ib];
            const float dstA = rayBounds(o, d, A.bmin, A.bmax);
            const float dstB = rayBounds(o, d, Bn.bmin, Bn.bmax);
            const bool nearA = dstA < dstB;
            const float dstNear = nearA ? dstA : dstB;
            const float dstFar = nearA ? dstB : dstA;
            const int iNear = nearA ? ia : ib;
            const int iFar = nearA ? ib : ia;
            if (dstFar < 1e38f && dstFar <= best.dst) stack[sp++] = iFar;
            if (dstNear < 1e38f && dstNear <= best.dst) stack[sp++] = iNear;
        }
    }
    finishHit(best, o, d);
    return best;
}

inline Hit closest(const orc_scene& s, V3 o, V3 d, bool bvh, Counters& cn) {
    cn.segments++;
    return bvh ? closestBVH(s, o, d, cn) : closestBrute(s, o, d, cn);
}

// GL_LINEAR / GL_REPEAT / unorm8 / no mips (textureClass.cpp:95-101); DESIGN.md §4.6
V3 sampleTexture(const Texture& t, float u, float v) {
    if (t.w <= 0 || t.h <= 0) return v3(0, 0, 0);
    float s = u - floorf(u);
    float r = v - floorf(v);
    if (!(s >= 0.0f && s <= 1.0f)) s = 0.0f;
    if (!(r >= 0.0f && r <= 1.0f)) r = 0.0f;
    const float fx = s * (float)t.w - 0.5f;
    const float fy = r * (float)t.h - 0.5f;
    const float flx = floorf(fx), fly = floorf(fy);
    const float ax = fx - flx, ay = fy - fly;
    int i0 = (int)flx, j0 = (int)fly;
    int i1 = i0 + 1, j1 = j0 + 1;
    i0 = ((i0 % t.w) + t.w) % t.w;
    i1 = ((i1 % t.w) + t.w) % t.w;
    j0 = ((j0 % t.h) + t.h) % t.h;
    j1 = ((j1 % t.h) + t.h) % t.h;
    auto texel = [&](int i, int j) {
        const uint8_t* p = &t.px[((size_t)j * t.w + i) * t.ch];
        const float r8 = (float)p[0] / 255.0f;
        if (t.ch == 1) return v3(r8, r8, r8);
        const float g8 = (float)p[1] / 255.0f;
        if (t.ch == 2) return v3(r8, g8, 0.0f);
        return v3(r8, g8, (float)p[2] / 255.0f);
    };
    const float w00 = (1.0f - ax) * (1.0f - ay), w10 = ax * (1.0f - ay), w01 = (1.0f - ax) * ay,
                w11 = ax * ay;
    return ((texel(i0, j0) * w00 + texel(i1, j0) * w10) + texel(i0, j1) * w01) + texel(i1, j1) * w11;
}

// S:342-368
V3 triangleTextureColor(const orc_scene& s, const rt_uniforms& un, int textureIndex, float w, float u,
                        float v, const rt_triangle& t) {
    const float uvx = (t.aTex[0] * u + t.bTex[0] * v) + t.cTex[0] * w;
    const float uvy = (t.aTex[1] * u + t.bTex[1] * v) + t.cTex[1] * w;
    if (textureIndex < 0 || textureIndex >= un.numTextures) return v3(0, 0, 0);
    if (textureIndex >= RT_MAX_TEXTURES) return v3(1.0f, 0.0f, 1.0f);
    return sampleTexture(s.tex[textureIndex], uvx, uvy);
}

// S:216-273
V3 envLight(V3 dir) {
    const V3 sunDir = normalize(v3(0.6f, 0.3f, -0.2f));
    const float sunDot = dot(dir, sunDir);
    const float horizonDot = dir.y;
    const V3 zenithColor = v3(0.15f, 0.25f, 0.65f), deepOrange = v3(1.2f, 0.4f, 0.1f),
             yellow = v3(1.0f, 0.8f, 0.3f), coolBlue = v3(0.3f, 0.4f, 0.7f),
             groundColor = v3(0.2f, 0.15f, 0.1f);
    const float sunToOpposite = (dot(dir, -sunDir) + 1.0f) * 0.5f;
    V3 horizonColor;
    if (sunToOpposite < 0.5f)
        horizonColor = mix(deepOrange, yellow, sunToOpposite * 2.0f);
    else
        horizonColor = mix(yellow, coolBlue, (sunToOpposite - 0.5f) * 2.0f);
    const float skyGradient = smoothstep(-0.2f, 0.8f, horizonDot);
    const V3 baseColor = mix(horizonColor, zenithColor, skyGradient);
    const V3 sunCenter = v3(15.0f, 15.0f, 10.0f);
    const float sunAngle = acos_(gclamp(sunDot, -1.0f, 1.0f));
    const float glow1 = exp_(-sunAngle * 600.0f);
    const float glow2 = exp_(-sunAngle * 150.0f) * 0.3f;
    const float glow3 = exp_(-sunAngle * 60.0f) * 0.1f;
    const float glow4 = exp_(-sunAngle * 15.0f) * 0.03f;
    const float totalGlow = ((glow1 + glow2) + glow3) + glow4;
    V3 finalColor = baseColor + sunCenter * totalGlow;
    if (horizonDot < 0.0f) {
        const float groundBlend = smoothstep(-0.1f, 0.0f, horizonDot);
        finalColor = mix(groundColor, finalColor, groundBlend);
        const float groundSunGlow = exp_(-sunAngle * 15.0f) * 0.2f;
        finalColor = finalColor + (sunCenter * groundSunGlow) * 0.05f;
    }
    return finalColor;
}

// S:201-214
V3 refract_(V3 I, V3 N, float eta, bool& isRefracted) {
    const float k = 1.0f - eta * eta * (1.0f - dot(N, I) * dot(N, I));
    if (k < 0.0f) {
        isRefracted = false;
        return reflect(I, N);
    }
    isRefracted = true;
    return I * eta - N * (eta * dot(N, I) + sqrtf(k));
}

inline bool blackChecker(V3 p, float scale) {  // S:524-527
    if (!(scale > 0.0f)) return false;
    const float sum = (floorf(p.x * scale) + floorf(p.y * scale)) + floorf(p.z * scale);
    const float m = sum - 2.0f * floorf(sum / 2.0f);  // mod(sum, 2)
    return m == 0.0f;
}

struct Ray {
    V3 origin, direction;
    bool insideGlass;
};

// S:472-563
V3 trace(const orc_scene& s, const rt_uniforms& un, Ray ray, Rng& rng, bool bvh, Counters& cn) {
    V3 rayColor = v3(1, 1, 1);
    V3 incomingLight = v3(0, 0, 0);
    int bounceCount = 0;
    while (bounceCount < un.maxBounceCount) {
        bounceCount++;
        rng.stream((uint32_t)bounceCount);
        Hit hit = closest(s, ray.origin, ray.direction, bvh, cn);
        if (hit.didHit) {
            const rt_material& material = s.mats[hit.mtl];
            if (material.materialType != RT_MAT_GLASS)
                ray.origin = hit.point - (ray.direction * hit.dst) * -1e-3f;
            else
                ray.origin = hit.point + (ray.direction * hit.dst) * -1e-3f;
            V3 attenuation = v3(0, 0, 0);
            const V3 prevDirection = ray.direction;
            const float bw = (1.0f - hit.u) - hit.v;  // S:336
            switch (material.materialType) {
                case RT_MAT_DIFFUSE:
                case RT_MAT_TEXTURE: {
                    ray.direction = normalize(hit.normal + randomDirection(rng));
                    const rt_triangle& tri = s.tris[hit.slot];
                    attenuation = material.materialType == RT_MAT_DIFFUSE
                                      ? v3(material.color)
                                      : triangleTextureColor(s, un, material.textureIndex, bw, hit.u,
                                                             hit.v, tri);
                    break;
                }
                case RT_MAT_SPECULAR: {
                    const V3 diffuseDirection = normalize(hit.normal + randomDirection(rng));
                    const V3 specularDirection = reflect(ray.direction, hit.normal);
                    const bool isSpecularBounce = material.specularProbability > rng.next();
                    ray.direction = mix(diffuseDirection, specularDirection,
                                        isSpecularBounce ? material.smoothness : 0.0f);
                    attenuation = isSpecularBounce ? v3(1, 1, 1) : v3(material.color);
                    break;
                }
                case RT_MAT_LIGHT: {
                    const V3 emitted = v3(material.emissionColor) * material.emissionStrength;
                    incomingLight = incomingLight + emitted * rayColor;
                    return incomingLight;
                }
                case RT_MAT_CHECKER: {
                    ray.direction = normalize(hit.normal + randomDirection(rng));
                    attenuation = blackChecker(ray.origin, material.checkerScale) ? v3(0, 0, 0)
                                                                                   : v3(1, 1, 1);
                    break;
                }
                case RT_MAT_GLASS: {
                    const float ri = ray.insideGlass ? material.refractiveIndex
                                                     : 1.0f / material.refractiveIndex;
                    bool isRefracted;
                    ray.direction = refract_(ray.direction, hit.normal, ri, isRefracted);
                    ray.insideGlass = isRefracted != ray.insideGlass;
                    attenuation = v3(material.color);
                    break;
                }
                default:
                    return v3(1.0f, 0.0f, 1.0f);
            }
            if (material.isEdgeHighlight != 0 && bounceCount > 1)
                ray.direction = prevDirection;
            else
                rayColor = rayColor * attenuation;
            const float p = gmax(rayColor.x, gmax(rayColor.y, rayColor.z));
            if (rng.next() > p) break;
            rayColor = rayColor * (1.0f / p);
        } else {
            if (un.environmentalLight != 0)
                incomingLight = incomingLight + envLight(ray.direction) * rayColor;
            return incomingLight;
        }
    }
    return incomingLight;
}

inline V3 normalizeColor(V3 c) {  // S:462-470
    const float m = gmax(gmax(c.x, c.y), c.z);
    if (m > 1.0f) return c / m;
    return c;
}

// S:565-645 (preview shading)
V3 traceBasic(const orc_scene& s, const rt_uniforms& un, Ray ray, bool bvh, Counters& cn) {
    V3 colorCumulative = v3(0, 0, 0);
    int bounceCount = 0;
    while (bounceCount < un.maxBounceCount) {
        bounceCount++;
        Hit hit = closest(s, ray.origin, ray.direction, bvh, cn);
        if (hit.didHit) {
            ray.origin = hit.point - hit.normal * 1e-4f;
            const rt_material& material = s.mats[hit.mtl];
            switch (material.materialType) {
                case RT_MAT_SPECULAR:
                    colorCumulative = colorCumulative + v3(material.color);
                    ray.direction = reflect(ray.direction, hit.normal);
                    break;
                case RT_MAT_DIFFUSE:
                case RT_MAT_TEXTURE:
                case RT_MAT_CHECKER: {
                    V3 color;
                    if (material.materialType == RT_MAT_TEXTURE) {
                        const float bw = (1.0f - hit.u) - hit.v;
                        color = triangleTextureColor(s, un, material.textureIndex, bw, hit.u, hit.v,
                                                     s.tris[hit.slot]);
                    } else if (material.materialType == RT_MAT_DIFFUSE) {
                        color = v3(material.color);
                    } else {
                        color = blackChecker(ray.origin, material.checkerScale) ? v3(0, 0, 0)
                                                                                 : v3(1, 1, 1);
                    }
                    colorCumulative = colorCumulative + color;
                    if (un.basicShadingShadow != 0) {
                        const V3 toLight = normalize(v3(un.basicShadingLightPosition) - hit.point);
                        Hit sh = closest(s, ray.origin, toLight, bvh, cn);
                        const V3 c = sh.didHit ? colorCumulative / 5.0f : colorCumulative;
                        return c / (float)bounceCount;
                    }
                    return colorCumulative / (float)bounceCount;
                }
                case RT_MAT_LIGHT:
                    return normalizeColor(v3(material.emissionColor));
                case RT_MAT_GLASS: {
                    const float ri = ray.insideGlass ? material.refractiveIndex
                                                     : 1.0f / material.refractiveIndex;
                    bool isRefracted;
                    ray.direction = refract_(ray.direction, hit.normal, ri, isRefracted);
                    ray.insideGlass = isRefracted != ray.insideGlass;
                    colorCumulative = v3(material.color);
                    break;
                }
                case RT_MAT_GLASS_HIGHLIGHT:
                    break;  // S:629-632: `bounceCount == 0` is never true
                default:
                    return v3(1.0f, 0.0f, 1.0f);
            }
        } else {
            colorCumulative = colorCumulative + envLight(ray.direction);
            break;
        }
    }
    return colorCumulative / (float)bounceCount;
}

// S:647-658
inline float aces1(float x) {
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    return gclamp((x * (a * x + b)) / (x * (c * x + d) + e), 0.0f, 1.0f);
}
inline V3 tonemapSRGB(V3 c) {
    return v3(pow_gamma(aces1(c.x)), pow_gamma(aces1(c.y)), pow_gamma(aces1(c.z)));
}

// Primary ray pieces of S:660-690
struct PixelSetup {
    V3 endPoint, centreDir;
    uint32_t seed;
};
inline PixelSetup pixelSetup(const rt_uniforms& un, int tx, int ty) {
    const int W = (int)un.width, H = (int)un.height;
    const float x = (float)(tx * 2 - W) / (float)W;
    const float y = (float)(ty * 2 - H) / (float)H;
    PixelSetup ps;
    ps.seed = (uint32_t)tx + (uint32_t)ty * (uint32_t)W + un.frameIndex * 968824447u;
    const V3 cam = v3(un.cameraPos), vf = v3(un.viewportFront), vr = v3(un.viewportRight),
             vu = v3(un.viewportUp);
    ps.endPoint = ((cam + vf) + vr * x) + vu * y;
    ps.centreDir = normalize((vf + vr * x) + vu * y);
    return ps;
}
inline Ray sampleRay(const rt_uniforms& un, const PixelSetup& ps, Rng& rng) {
    Ray r;
    const float angle = rng.next();  // S:163
    const float cx = cos01(angle), sy = sin01(angle);
    r.origin = (v3(un.cameraPos) + v3(un.defocusDiskRight) * cx) + v3(un.defocusDiskUp) * sy;
    const float jr = rng.range(-0.5f, 0.5f);
    const float ju = rng.range(-0.5f, 0.5f);
    const V3 endJ = (ps.endPoint + v3(un.pixelRight) * jr) + v3(un.pixelUp) * ju;
    r.direction = normalize(endJ - r.origin);
    r.insideGlass = false;
    return r;
}

void renderPixel(const orc_scene& s, const rt_uniforms& un, int rng_mode, bool bvh, int tx, int ty,
                 float* out4, Counters& cn) {
    const PixelSetup ps = pixelSetup(un, tx, ty);
    V3 color;
    if (un.basicShading != 0) {
        Ray ray;
        ray.origin = v3(un.cameraPos);
        ray.direction = ps.centreDir;
        ray.insideGlass = false;  // undefined in S:674-677; taken as false
        cn.paths++;
        color = traceBasic(s, un, ray, bvh, cn);
    } else {
        Rng rng;
        rng.mode = rng_mode;
        rng.state = ps.seed;
        rng.pixel = (uint32_t)tx + (uint32_t)ty * un.width;
        rng.frame = un.frameIndex;
        V3 cum = v3(0, 0, 0);
        for (int i = 0; i < un.numRaysPerPixel; i++) {
            rng.sample = (uint32_t)i;
            rng.stream(0);
            Ray ray = sampleRay(un, ps, rng);
            cn.paths++;
            cum = cum + trace(s, un, ray, rng, bvh, cn);
        }
        color = cum / (float)un.numRaysPerPixel;
        color = tonemapSRGB(color);
    }
    out4[0] = color.x;
    out4[1] = color.y;
    out4[2] = color.z;
    out4[3] = 1.0f;
}

template <class F>
void parallelRows(int y0, int y1, int threads, F&& f) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads == 1) {
        for (int y = y0; y < y1; y++) f(y, 0);
        return;
    }
    std::atomic<int> next(y0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t]() {
            for (;;) {
                const int y = next.fetch_add(1);
                if (y >= y1) break;
                f(y, t);
            }
        });
    for (auto& th : pool) th.join();
}

inline uint32_t quantize8(float c) {  // GL float → unorm8 of the RGB8 blit (R:194-217)
    if (!(c > 0.0f)) return 0;
    if (c >= 1.0f) return 255;
    return (uint32_t)(c * 255.0f + 0.5f);
}

// ------------------------------------------------------------------------------------------------
// BVH.h restated
struct BBox {
    V3 mn{1e30f, 1e30f, 1e30f}, mx{-1e30f, -1e30f, -1e30f};
    void grow(const BvhTri& t) {  // B:41-45
        mn = v3(gmin(mn.x, t.mn.x), gmin(mn.y, t.mn.y), gmin(mn.z, t.mn.z));
        mx = v3(gmax(mx.x, t.mx.x), gmax(mx.y, t.mx.y), gmax(mx.z, t.mx.z));
    }
    V3 size() const {  // B:30-33: x-extent in all three components (reference quirk)
        const float sx = mx.x - mn.x;
        return v3(sx, sx, sx);
    }
    void expand() {  // B:47-51
        mn = mn - v3(1e-4f, 1e-4f, 1e-4f);
        mx = mx + v3(1e-4f, 1e-4f, 1e-4f);
    }
};
inline float nodeCost(V3 size, int count) {  // B:79-90
    const float halfArea = size.x * (size.y + size.z) + size.y * size.z;
    return halfArea * (float)count;
}
inline float axisOf(V3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

struct Builder {
    std::vector<BvhTri> bt;
    std::vector<rt_triangle>& tris;
    std::vector<int32_t>& orig;
    std::vector<rt_ref_node>& nodes;
    std::vector<BBox> boxes;  // bounds per node (kept as V3 for arithmetic)
    Builder(std::vector<rt_triangle>& t, std::vector<int32_t>& o, std::vector<rt_ref_node>& n)
        : tris(t), orig(o), nodes(n) {}

    void push(const BBox& b, int triIdx, int count, int child) {
        rt_ref_node n;
        memset(&n, 0, sizeof n);
        n.bmin[0] = b.mn.x; n.bmin[1] = b.mn.y; n.bmin[2] = b.mn.z;
        n.bmax[0] = b.mx.x; n.bmax[1] = b.mx.y; n.bmax[2] = b.mx.z;
        n.triangleIndex = triIdx;
        n.triangleCount = count;
        n.childIndex = child;
        nodes.push_back(n);
        boxes.push_back(b);
    }
    float evaluateSplit(int node, int axis, float pos) {  // B:92-115
        BBox A, Bx;
        int nA = 0, nB = 0;
        const int i0 = nodes[node].triangleIndex, i1 = i0 + nodes[node].triangleCount;
        for (int i = i0; i < i1; i++) {
            if (axisOf(bt[i].center, axis) < pos) {
                A.grow(bt[i]);
                nA++;
            } else {
                Bx.grow(bt[i]);
                nB++;
            }
        }
        return nodeCost(A.size(), nA) + nodeCost(Bx.size(), nB);
    }
    void chooseSplit(int& axisOut, float& posOut, float& cost, int node) {  // B:117-143
        cost = 1e32f;
        posOut = 0;
        axisOut = 0;
        for (int axis = 0; axis < 3; axis++) {
            const float start = axisOf(boxes[node].mn, axis), end = axisOf(boxes[node].mx, axis);
            for (int i = 0; i < 10; i++) {
                const float t = (float)(i + 1) / (float)(10 + 1);
                const float pos = start + (end - start) * t;
                const float c = evaluateSplit(node, axis, pos);
                if (c < cost) {
                    cost = c;
                    posOut = pos;
                    axisOut = axis;
                }
            }
        }
    }
    void split(int root, int depth) {  // B:170-220
        if (depth == 32 || nodes[root].triangleCount < 1) return;
        int axis;
        float pos, cost;
        chooseSplit(axis, pos, cost, root);
        if (cost >= nodeCost(boxes[root].size(), nodes[root].triangleCount)) return;
        BBox bA, bB;
        int aIdx = nodes[root].triangleIndex, aCnt = 0;
        int bIdx = nodes[root].triangleIndex, bCnt = 0;
        const int i0 = nodes[root].triangleIndex, i1 = i0 + nodes[root].triangleCount;
        for (int i = i0; i < i1; i++) {
            const bool inA = axisOf(bt[i].center, axis) < pos;
            if (inA) {
                bA.grow(bt[i]);
                aCnt++;
                const int sw = aIdx + aCnt - 1;
                std::swap(bt[i], bt[sw]);
                std::swap(tris[i], tris[sw]);
                std::swap(orig[i], orig[sw]);
                bIdx += 1;
            } else {
                bB.grow(bt[i]);
                bCnt++;
            }
        }
        bA.expand();
        bB.expand();
        if (aCnt > 0 || bCnt > 0) {
            const int childA = (int)nodes.size();
            nodes[root].childIndex = childA;
            push(bA, aIdx, aCnt, -1);
            push(bB, bIdx, bCnt, -1);
            split(childA, depth + 1);
            split(childA + 1, depth + 1);
        }
    }
};

}  // namespace

// ================================================================================================
extern "C" {

orc_scene* orc_scene_create(const rt_triangle* tris, int64_t n, const rt_material* mats, int32_t k) {
    if (n < 0 || k <= 0 || (n > 0 && !tris) || !mats) return nullptr;
    for (int64_t i = 0; i < n; i++)
        if (tris[i].materialIndex < 0 || tris[i].materialIndex >= k) return nullptr;
    orc_scene* s = new orc_scene;
    s->tris.assign(tris, tris + n);
    s->mats.assign(mats, mats + k);
    s->orig.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) s->orig[(size_t)i] = (int32_t)i;
    return s;
}
void orc_scene_destroy(orc_scene* s) { delete s; }

int orc_scene_set_texture(orc_scene* s, int32_t slot, const uint8_t* pixels, int32_t w, int32_t h,
                          int32_t ch) {
    if (!s || slot < 0 || slot >= RT_MAX_TEXTURES || !pixels || w <= 0 || h <= 0 || ch < 1 || ch > 4)
        return RT_ERR_INVALID;
    Texture& t = s->tex[slot];
    t.w = w; t.h = h; t.ch = ch;
    t.px.assign(pixels, pixels + (size_t)w * h * ch);
    return RT_OK;
}

int orc_scene_build_bvh(orc_scene* s) {
    if (!s) return RT_ERR_INVALID;
    s->nodes.clear();
    Builder b(s->tris, s->orig, s->nodes);
    b.bt.resize(s->tris.size());
    BBox bounds;
    for (size_t i = 0; i < s->tris.size(); i++) {  // M:148-153
        const V3 a = v3(s->tris[i].a), bb = v3(s->tris[i].b), c = v3(s->tris[i].c);
        BvhTri& t = b.bt[i];
        t.mn = v3(gmin(gmin(a.x, bb.x), c.x), gmin(gmin(a.y, bb.y), c.y), gmin(gmin(a.z, bb.z), c.z));
        t.mx = v3(gmax(gmax(a.x, bb.x), c.x), gmax(gmax(a.y, bb.y), c.y), gmax(gmax(a.z, bb.z), c.z));
        t.center = ((a + bb) + c) / 3.0f;
        bounds.grow(t);
    }
    bounds.expand();
    b.push(bounds, 0, (int)s->tris.size(), -1);  // B:154-160
    b.split(0, 1);
    s->built = true;
    return RT_OK;
}
int64_t orc_scene_node_count(const orc_scene* s) { return s ? (int64_t)s->nodes.size() : 0; }
int orc_scene_get_nodes(const orc_scene* s, rt_ref_node* out) {
    if (!s || !out) return RT_ERR_INVALID;
    memcpy(out, s->nodes.data(), s->nodes.size() * sizeof(rt_ref_node));
    return RT_OK;
}
int orc_scene_get_permuted(const orc_scene* s, rt_triangle* tris_out, int32_t* orig_out) {
    if (!s) return RT_ERR_INVALID;
    if (tris_out) memcpy(tris_out, s->tris.data(), s->tris.size() * sizeof(rt_triangle));
    if (orig_out) memcpy(orig_out, s->orig.data(), s->orig.size() * sizeof(int32_t));
    return RT_OK;
}

int orc_trace_rays(const orc_scene* s, const float* origins, const float* dirs, int64_t count,
                   int use_bvh, int32_t* tri_id, float* dst, float* bu, float* bv) {
    if (!s || !origins || !dirs || count < 0) return RT_ERR_INVALID;
    if (use_bvh && !s->built) return RT_ERR_STATE;
    Counters cn;
    for (int64_t i = 0; i < count; i++) {
        Hit h = closest(*s, v3(origins + 3 * i), v3(dirs + 3 * i), use_bvh != 0, cn);
        if (tri_id) tri_id[i] = h.didHit ? h.orig : -1;
        if (dst) dst[i] = h.dst;
        if (bu) bu[i] = h.didHit ? h.u : 0.0f;
        if (bv) bv[i] = h.didHit ? h.v : 0.0f;
    }
    return RT_OK;
}

int orc_first_hit(const orc_scene* s, const rt_uniforms* u, int32_t mode, int32_t rng_mode,
                  int use_bvh, int threads, int32_t* tri_id, float* dst) {
    if (!s || !u) return RT_ERR_INVALID;
    if (use_bvh && !s->built) return RT_ERR_STATE;
    const int W = (int)u->width, H = (int)u->height;
    parallelRows(0, H, threads, [&](int ty, int) {
        Counters cn;
        for (int tx = 0; tx < W; tx++) {
            const PixelSetup ps = pixelSetup(*u, tx, ty);
            V3 o, d;
            if (mode == RT_FIRST_HIT_CENTRE) {
                o = v3(u->cameraPos);
                d = ps.centreDir;
            } else {
                Rng rng;
                rng.mode = rng_mode;
                rng.state = ps.seed;
                rng.pixel = (uint32_t)tx + (uint32_t)ty * u->width;
                rng.frame = u->frameIndex;
                rng.sample = 0;
                rng.stream(0);
                Ray r = sampleRay(*u, ps, rng);
                o = r.origin;
                d = r.direction;
            }
            Hit h = closest(*s, o, d, use_bvh != 0, cn);
            const size_t i = (size_t)ty * W + tx;
            if (tri_id) tri_id[i] = h.didHit ? h.orig : -1;
            if (dst) dst[i] = h.dst;
        }
    });
    return RT_OK;
}

int orc_render_frame(const orc_scene* s, const rt_uniforms* u, int32_t rng_mode, int threads,
                     int32_t x0, int32_t y0, int32_t x1, int32_t y1, float* rgba, orc_counters* out) {
    if (!s || !u || !rgba || !s->built) return RT_ERR_INVALID;
    const int W = (int)u->width;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    std::vector<Counters> cns((size_t)threads);
    parallelRows(y0, y1, threads, [&](int ty, int t) {
        for (int tx = x0; tx < x1; tx++)
            renderPixel(*s, *u, rng_mode, true, tx, ty, rgba + ((size_t)ty * W + tx) * 4, cns[t]);
    });
    if (out) {
        for (auto& c : cns) {
            out->segments += c.segments;
            out->paths += c.paths;
            out->node_visits += c.node_visits;
            out->tri_tests += c.tri_tests;
        }
    }
    return RT_OK;
}

void orc_finalize(const uint32_t* sums, int32_t W, int32_t H, int32_t frames, uint8_t* rgb8) {
    // R:248-259: average (float divide), min(255), truncate, flip
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W * 3; x++) {
            const float avg = (float)sums[(size_t)y * W * 3 + x] / (float)frames;
            const float c = avg < 255.0f ? avg : 255.0f;
            rgb8[(size_t)(H - 1 - y) * W * 3 + x] = (uint8_t)c;
        }
}

int orc_screenshot(const orc_scene* s, const rt_uniforms* u, int32_t frames, int32_t rng_mode,
                   int threads, const int32_t* frame_list, int32_t n_list, uint32_t* sum_out,
                   uint8_t* rgb8, orc_counters* counters) {
    if (!s || !u || frames <= 0) return RT_ERR_INVALID;
    const int W = (int)u->width, H = (int)u->height;
    std::vector<float> img((size_t)W * H * 4);
    std::vector<uint32_t> sums((size_t)W * H * 3, 0u);
    const int n = frame_list ? n_list : frames;
    for (int k = 0; k < n; k++) {
        rt_uniforms uf = *u;
        uf.frameIndex = (uint32_t)(frame_list ? frame_list[k] : k);  // R:187
        int rc = orc_render_frame(s, &uf, rng_mode, threads, 0, 0, W, H, img.data(), counters);
        if (rc) return rc;
        for (size_t p = 0; p < (size_t)W * H; p++)  // R:194-238
            for (int c = 0; c < 3; c++) sums[p * 3 + c] += quantize8(img[p * 4 + c]);
    }
    if (sum_out) memcpy(sum_out, sums.data(), sums.size() * sizeof(uint32_t));
    if (rgb8) orc_finalize(sums.data(), W, H, frames, rgb8);
    return RT_OK;
}

// C:99-192.  `cos`, `sin`, `exp` are the unqualified calls of camera.h (float overloads once
// <cmath> is in scope, as it is through glm); glm::tan is std::tan(float).
void orc_camera_uniforms(int32_t width, int32_t height, const float pos[3], float hfov, float pitch,
                         float yaw, float focusDistance, float defocusAngle, float zoom,
                         rt_uniforms* out) {
    const float aspect = (float)width / (float)height;  // C:102
    V3 front;
    front.x = std::cos(yaw) * std::cos(pitch);  // C:152-154 (float overloads)
    front.y = std::sin(pitch);
    front.z = std::sin(yaw) * std::cos(pitch);
    front = normalize(front);
    const V3 worldUp = v3(0, 1, 0);
    const V3 right = normalize(cross(worldUp, front));
    const V3 up = normalize(cross(front, right));
    const float h = std::tan(hfov / 2);
    const float viewportWidth = 2 * h / std::exp(zoom * 0.1f);
    const float viewportHeight = viewportWidth / aspect;
    const V3 viewportRight = (right * viewportWidth) * focusDistance;
    const V3 viewportUp = (up * viewportHeight) * focusDistance;
    const V3 viewportFront = (-front) * focusDistance;
    const V3 pixelRight = viewportRight / (float)width;   // scrWidth is a float member (C:42)
    const V3 pixelUp = viewportUp / (float)height;
    const float defocusRadius = focusDistance * std::tan(defocusAngle / 2.0f);
    const V3 ddr = right * defocusRadius, ddu = up * defocusRadius;
    auto put = [](float* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = 0.0f; };
    put(out->cameraPos, v3(pos));
    put(out->viewportRight, viewportRight);
    put(out->viewportUp, viewportUp);
    put(out->viewportFront, viewportFront);
    put(out->pixelRight, pixelRight);
    put(out->pixelUp, pixelUp);
    put(out->defocusDiskRight, ddr);
    put(out->defocusDiskUp, ddu);
    out->width = (uint32_t)width;
    out->height = (uint32_t)height;
}

float orc_random(uint32_t* state) { return pcg_next(*state); }
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox(ctr, key, out);
}
float orc_philox_draw(uint32_t pixel, uint32_t frame, uint32_t sample, uint32_t bounce, uint32_t j) {
    Rng r;
    r.mode = RT_RNG_PHILOX;
    r.pixel = pixel; r.frame = frame; r.sample = sample;
    r.stream(bounce);
    r.j = j;
    return r.next();
}
float orc_cos01(float x) { return cos01(x); }
float orc_sin01(float x) { return sin01(x); }
float orc_exp(float x) { return exp_(x); }
float orc_acos(float x) { return acos_(x); }
float orc_pow_gamma(float x) { return pow_gamma(x); }
void orc_tonemap_srgb(const float in[3], float out[3]) {
    V3 c = tonemapSRGB(v3(in));
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
void orc_env_light(const float dir[3], float out[3]) {
    V3 c = envLight(v3(dir));
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
void orc_sample_texture(const orc_scene* s, int32_t tex, float u, float v, float out[3]) {
    V3 c = (tex >= 0 && tex < RT_MAX_TEXTURES) ? sampleTexture(s->tex[tex], u, v) : v3(0, 0, 0);
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
int orc_ray_triangle(const float o[3], const float d[3], const rt_triangle* t, float* dst, float* u,
                     float* v) {
    float dd = 0, uu = 0, vv = 0;
    V3 N;
    const bool hit = rayTriangle(v3(o), v3(d), *t, dd, uu, vv, N);
    if (dst) *dst = dd;
    if (u) *u = uu;
    if (v) *v = vv;
    return hit ? 1 : 0;
}
float orc_ray_bounds(const float o[3], const float d[3], const float bmin[3], const float bmax[3]) {
    return rayBounds(v3(o), v3(d), bmin, bmax);
}

}  // extern "C"
