// rt_math.cuh — device arithmetic of the path tracer.
//
// The whole library is compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false: every
// binary32 operation below rounds once, in the order written, exactly like the CPU oracle's
// (-ffp-contract=off).  Where a fused multiply-add is wanted for speed and the result does not
// have to match the oracle bit for bit (box slabs only) the code says __fmaf_rn explicitly.
//
// Operation order follows glm 0.9.9.7 (external/glm/detail/func_geometric.inl:48-110), which the
// project takes as the normative definition of the GLSL built-ins.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rt {

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 v3(const float* p) { return V3{p[0], p[1], p[2]}; }
__device__ __forceinline__ V3 v3(float4 f) { return V3{f.x, f.y, f.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
// dot3 = (a.x*b.x + a.y*b.y) + a.z*b.z
__device__ __forceinline__ float dot(V3 a, V3 b) {
    const float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
    return (tx + ty) + tz;
}
__device__ __forceinline__ V3 cross(V3 x, V3 y) {
    return V3{x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
}
__device__ __forceinline__ float length(V3 v) { return sqrtf(dot(v, v)); }
__device__ __forceinline__ V3 normalize(V3 v) { return v * (1.0f / sqrtf(dot(v, v))); }
__device__ __forceinline__ V3 reflect(V3 I, V3 N) { return I - (N * dot(N, I)) * 2.0f; }
__device__ __forceinline__ float gmin(float x, float y) { return (y < x) ? y : x; }
__device__ __forceinline__ float gmax(float x, float y) { return (x < y) ? y : x; }
__device__ __forceinline__ float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
__device__ __forceinline__ V3 mix(V3 x, V3 y, float a) { return x * (1.0f - a) + y * a; }
__device__ __forceinline__ float smoothstep(float e0, float e1, float x) {
    const float t = gclamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

// ---------------------------------------------------------------------------------- elementary
// cos / sin on [0,1] (the defocus angle is random() in [0,1], compute.glsl:163-164)
__device__ __forceinline__ float cos01(float x) {
    const float z = x * x;
    float p = 2.08767570e-09f;
    p = p * z + -2.75573192e-07f;
    p = p * z + 2.48015873e-05f;
    p = p * z + -1.38888889e-03f;
    p = p * z + 4.16666667e-02f;
    p = p * z + -0.5f;
    p = p * z + 1.0f;
    return p;
}
__device__ __forceinline__ float sin01(float x) {
    const float z = x * x;
    float p = -2.50521084e-08f;
    p = p * z + 2.75573192e-06f;
    p = p * z + -1.98412698e-04f;
    p = p * z + 8.33333333e-03f;
    p = p * z + -1.66666667e-01f;
    p = p * z + 1.0f;
    return x * p;
}
__device__ __forceinline__ float exp_poly(float r) {
    float p = 1.98412698e-04f;
    p = p * r + 1.38888889e-03f;
    p = p * r + 8.33333333e-03f;
    p = p * r + 4.16666667e-02f;
    p = p * r + 1.66666667e-01f;
    p = p * r + 0.5f;
    p = p * r + 1.0f;
    p = p * r + 1.0f;
    return p;
}
__device__ __forceinline__ float pow2i(float n) {
    return __uint_as_float((uint32_t)((int32_t)n + 127) << 23);
}
__device__ __forceinline__ float exp_(float x) {
    if (!(x >= -87.0f)) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    const float n = floorf(x * 1.44269504f + 0.5f);
    const float r = (x - n * 0.693145752f) - n * 1.42860677e-06f;
    return exp_poly(r) * pow2i(n);
}
__device__ __forceinline__ float asin_R(float z) {
    const float pS0 = 1.6666586697e-01f, pS1 = -4.2743422091e-02f, pS2 = -8.6563630030e-03f,
                qS1 = -7.0662963390e-01f;
    const float p = z * (pS0 + z * (pS1 + z * pS2));
    const float q = 1.0f + z * qS1;
    return p / q;
}
__device__ __forceinline__ float acos_(float x) {
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
    const float df = __uint_as_float(__float_as_uint(s) & 0xfffff000u);
    const float c = (z - df * df) / (s + df);
    const float w = asin_R(z) * s + c;
    return 2.0f * (df + w);
}
// pow(x, 1/2.2), x in [0,1]
__device__ __forceinline__ float pow_gamma(float x) {
    if (!(x >= 1.17549435e-38f)) return 0.0f;
    const uint32_t bits = __float_as_uint(x);
    int32_t e = (int32_t)((bits >> 23) & 0xffu) - 127;
    float m = __uint_as_float((bits & 0x007fffffu) | 0x3f800000u);
    if (m > 1.41421354f) {
        m = m * 0.5f;
        e += 1;
    }
    const float s = (m - 1.0f) / (m + 1.0f);
    const float z = s * s;
    float p = 9.09090909e-02f;
    p = p * z + 1.11111111e-01f;
    p = p * z + 1.42857143e-01f;
    p = p * z + 0.2f;
    p = p * z + 3.33333333e-01f;
    p = p * z + 1.0f;
    const float lnm = (s + s) * p;
    const float log2x = (float)e + lnm * 1.44269504f;
    const float y = log2x * 0.45454547f;
    const float n = floorf(y + 0.5f);
    if (n < -126.0f) return 0.0f;
    const float r = (y - n) * 0.693147182f;
    return exp_poly(r) * pow2i(n);
}
// tonemapACES + toSRGB, compute.glsl:647-658
__device__ __forceinline__ float aces1(float x) {
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    return gclamp((x * (a * x + b)) / (x * (c * x + d) + e), 0.0f, 1.0f);
}
__device__ __forceinline__ V3 tonemap_srgb(V3 c) {
    return v3(pow_gamma(aces1(c.x)), pow_gamma(aces1(c.y)), pow_gamma(aces1(c.z)));
}
__device__ __forceinline__ uint32_t quantize8(float c) {  // float → unorm8 of the RGB8 blit
    if (!(c > 0.0f)) return 0u;
    if (c >= 1.0f) return 255u;
    return (uint32_t)(c * 255.0f + 0.5f);
}

// ---------------------------------------------------------------------------------- RNG
__device__ __forceinline__ float u32_to_unit(uint32_t r) {
    return (float)r / 4294967296.0f;  // compute.glsl:153 (the literal is a float: 2^32)
}
// random() * 2 - 1 (compute.glsl:177-179) straight from the 32-bit draw: float(r) / 2^32 and the doubling are exact
// scalings by powers of two, so the only rounding of `u32_to_unit(r) * 2.0f - 1.0f` is that of the subtraction, and
// fma(float(r), 2^-31, -1) rounds the same exact value once — identical bits, two instructions instead of four.
__device__ __forceinline__ float u32_to_signed_unit(uint32_t r) {
    return __fmaf_rn((float)r, 4.656612873077392578125e-10f, -1.0f);
}
__device__ __forceinline__ float pcg_next(uint32_t& state) {  // compute.glsl:148-154
    state = state * 747796405u + 2891336453u;
    uint32_t result = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    result = (result >> 22u) ^ result;
    return u32_to_unit(result);
}
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// One random stream per (pixel, frame, sample, bounce); draw j is word (j & 3) of
// philox(ctr = (j >> 2, bounce, sample, 'RT20'), key = (pixel, frame)).  In RT_RNG_REF_PCG mode the
// stream is the reference's sequential hash and `state` is the only live field.
template <int MODE>
struct Rng;
template <>
struct Rng<0> {  // RT_RNG_REF_PCG
    uint32_t state;
    __device__ __forceinline__ void init(uint32_t st, uint32_t, uint32_t, uint32_t) { state = st; }
    __device__ __forceinline__ void stream(uint32_t) {}
    __device__ __forceinline__ float next() { return pcg_next(state); }
    __device__ __forceinline__ uint32_t carry() const { return state; }
};
template <>
struct Rng<1> {  // RT_RNG_PHILOX
    uint32_t pixel, frame, sample, bounce, j;
    uint32_t cache[4];
    __device__ __forceinline__ void init(uint32_t smp, uint32_t pix, uint32_t frm, uint32_t) {
        sample = smp; pixel = pix; frame = frm; bounce = 0; j = 0;
    }
    __device__ __forceinline__ void stream(uint32_t b) { bounce = b; j = 0; }
    __device__ __forceinline__ float next() {
        if ((j & 3u) == 0u) philox4x32_10(j >> 2, bounce, sample, 0x52543230u, pixel, frame, cache);
        const uint32_t w = j & 3u;
        const uint32_t r = w == 0 ? cache[0] : (w == 1 ? cache[1] : (w == 2 ? cache[2] : cache[3]));
        j++;
        return u32_to_unit(r);
    }
    __device__ __forceinline__ uint32_t carry() const { return sample; }
};

// compute.glsl:174-185
template <class R>
__device__ __forceinline__ V3 random_direction(R& rng) {
    for (int i = 0; i < 100; i++) {
        const float x = rng.next() * 2.0f - 1.0f;
        const float y = rng.next() * 2.0f - 1.0f;
        const float z = rng.next() * 2.0f - 1.0f;
        if (length(v3(x, y, z)) < 1.0f) return normalize(v3(x, y, z));
    }
    return v3(0.0f, 0.0f, 0.0f);
}

}  // namespace rt
