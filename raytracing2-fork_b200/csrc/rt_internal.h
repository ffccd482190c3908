// rt_internal.h — declarations shared by the translation units of librt_b200.so (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rt_b200.h"
#include "rt_scene.cuh"

namespace rt {

// ---------------------------------------------------------------- bvh_build.cu
struct BuildArgs {
    const rt_triangle* tris;  // device copy of the caller's array (original order)
    int n;
    int use_ploc;             // 1 = PLOC hierarchy (default), 0 = Karras LBVH
    int dfs_layout;           // 1 = store PLOC nodes in depth-first order
    int ploc_radius;          // search window of the clustering, 0 = default
    int leaf_max_tris;        // most triangles a leaf may hold (PLOC; 0 = default 4, 1 = one triangle per leaf)
    float leaf_cb;            // SAH cost of a box test relative to a triangle test (0 = default)
    float4* centroid;         // n
    uint32_t* bounds;         // 12 order-preserving uints
    uint64_t* keys[2];        // n each
    uint32_t* vals[2];        // n each
    uint32_t* hist;           // build_scratch_words(n): sort histograms, later PLOC control block + tile states
    int32_t* children;        // 2*(n-1)
    int32_t* parent;          // 2n-1
    float4* boxes;            // 2*(2n-1)
    uint32_t* flags;          // n+1
    uint32_t* nodeDepth;      // 2n (height of the subtree under each inner node / entity)
    uint32_t* status;         // kBuildStatusWords: depth of the binary tree + error flags (see below)
    int sm_count;
    float* grid;              // 6: quantisation grid (output)
    uint4* nodes;             // 2*(n-1)   (output)
    uint4* nodes4;            // 4*(n-1) or NULL: 4-wide nodes collapsed from the binary tree (output)
    uint32_t* wide_count;     // 2: [0] number of 4-wide nodes written, [1] worst-case traversal stack entries (device)
    int* wide_levels;         // host: depth of the 4-wide tree = levels the collapse ran (output)
    float4* geom;             // 4*n       (output)
    float4* shade;            // 2*n       (output)
    int32_t* orig;            // n         (output)
};
// status words written by the build (device), read back once by rt_scene_build
constexpr int kBuildDepth = 0;       // depth of the binary tree
constexpr int kBuildNonFinite = 1;   // != 0: a vertex coordinate is inf or NaN
constexpr int kBuildPlocStuck = 2;   // != 0: PLOC found no mutual pair / ran out of rounds (caller falls back to Karras)
constexpr int kBuildPlocRounds = 3;  // rounds the clustering took (diagnostic)
constexpr int kBuildStatusWords = 4;
cudaError_t build_lbvh(const BuildArgs& a, cudaStream_t st, uint64_t* launches);
size_t build_scratch_words(int n);
// sets *flag (device word) to 1 if a vertex coordinate is NaN, inf or larger than 1e18 in magnitude
cudaError_t validate_triangles(const rt_triangle* tris, int n, uint32_t* flag, int sm_count, cudaStream_t st);

// ---------------------------------------------------------------- wavefront.cu
// Path state of the wavefront, structure of arrays, 16-byte records, ping-ponged between bounces
// so that every kernel reads and writes dense, coalesced arrays.
struct PathArrays {
    float4* od0;   // origin.xyz, dir.x
    float4* od1;   // dir.yz, throughput.xy
    float4* misc;  // throughput.z, slot id (bits), rng carry (bits), flags (bits: bounce | insideGlass<<16)
};
// Exact unsigned division by a launch constant (Granlund & Montgomery): q = (t + ((n - t) >> 1)) >> (l - 1) with
// t = umulhi(M, n); valid for every 32-bit n and 2 <= d < 2^31.  d == 1 is flagged by l == 0.
struct FastDiv {
    uint32_t M, l;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f{0u, 0u};
    if (d <= 1u) return f;
    uint32_t l = 0;
    while (((uint64_t)1 << l) < d) l++;
    f.M = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << l) - d)) / d + 1);
    f.l = l;
    return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, FastDiv f) {
    if (f.l == 0u) return n;
    const uint32_t t = __umulhi(f.M, n);
    return (t + ((n - t) >> 1)) >> (f.l - 1u);
}
#endif
struct FrameParams {
    rt_uniforms u;
    int32_t width, height;
    int32_t local_pixels;       // pixels this rank renders
    // One batch of the wavefront holds `lanes_active` lanes of every local pixel; lane L carries sample
    // sample_base + L % samples_in_batch of batch frame L / samples_in_batch, whose frameIndex is
    // u.frameIndex + (L / samples_in_batch) * frame_stride.  Several frames share a batch whenever a whole
    // frame needs fewer path slots than the budget (small images, tile split, RT_RNG_REF_PCG where a pixel's
    // samples are sequential), so the persistent kernels always see large wavefronts.
    int32_t sample_base;        // first sample index of this batch
    int32_t lanes_active;       // frames_in_batch * samples_in_batch
    int32_t frames_in_batch;    // >= 1
    int32_t samples_in_batch;   // samples of one pixel of one frame in flight together (1 in RT_RNG_REF_PCG mode)
    int32_t frame_stride;       // frameIndex step between batch frames (world_size under RT_SPLIT_FRAMES)
    int32_t debug_zero_contrib; // RT_DEBUG_ZERO_CONTRIB=1: raygen clears every contrib slot (debugging aid)
    const int32_t* rows;        // local row -> absolute row (RT_SPLIT_TILES), NULL = identity
    FastDiv div_pixels;         // / local_pixels
    FastDiv div_samples;        // / samples_in_batch
    FastDiv div_width;          // / width
};
struct WaveBuffers {
    PathArrays cur, next;
    float4* hit;          // t, u, v, slot(bits) per path of `cur`
    float4* contrib;      // radiance of each (lane, pixel) slot of the batch
    float4* accum;        // running sum per (batch frame, local pixel) (xyz)
    uint32_t* pix_rng;    // RT_RNG_REF_PCG: the per-(batch frame, pixel) stream state carried across samples
    uint32_t* counts;     // counts[b] = live paths entering bounce b (maxBounce + 2 entries), followed by
                          // the same number of k_extend work cursors
    unsigned long long* stats;  // [0] segments [1] paths [2] node visits [3] tri tests
    float4* image;        // W*H RGBA32F, bottom-up (the reference's image binding 0)
    uint32_t* frame_sum;  // W*H*3 sums of the 8-bit frames, bottom-up
    uint8_t* out_rgb8;    // W*H*3 final, top-down
};

// queue + results of the parity hooks (rt_first_hit, rt_trace_rays), which run k_extend on caller-defined rays
struct HookBuffers {
    PathArrays rays;            // od0 / od1 only
    float4* hit;
    uint32_t* counts;           // [0] rays, [1] work cursor
    unsigned long long* stats;  // the ctx counters (segments, node visits, triangle tests)
};

struct Launcher {
    cudaStream_t st;
    int sm_count;
    int rng_mode;
    bool instrument;
    int extend_grid;   // persistent k_extend grid: SMs x resident blocks per SM
    int extend_grid_wide;  // the same for the 4-wide variant (more registers)
    int node_steps_wide;   // node steps per vote of the 4-wide variant
    int leaf_vote;     // k_extend: leaf step when this many lanes wait at a leaf
    int refill;        // k_extend: refill when this many lanes are idle
    int node_steps;    // k_extend: node steps per vote
    bool speculative;  // k_extend: postponed-leaf variant
    bool top_smem;     // k_extend (4-wide): root + three levels of the tree staged in shared memory
    bool widen_always;   // debugging aid: widened slab test on every bounce
    bool widen_primary;  // camera rays may start far outside the quantisation grid (rt_scene.cuh, slab1)
    bool hooks_thread; // parity hooks walk the binary tree per thread instead of running k_extend
    int shade_blocks_per_sm;  // grid-stride k_shade: blocks per SM
    int shade_defer_bounces;      // auto policy: textured scenes defer the first this-many bounces
    int shade_defer_batch;        // windows per deferred reservation at bounce 0 (1 .. 3)
    int shade_defer_batch_later;  // ... at the later bounces
    int shade_defer;          // k_shade's deferred queue append: 0 never, 1 at bounce 0, 2 at every bounce, 3 auto
    uint64_t* kernel_launches;
    uint64_t* extend_launches;
    // optional per-class device timing
    cudaEvent_t* ev_pool;
    int ev_cap;
    int* ev_used;
    int* ev_tag;  // 0 extend, 1 other
    bool timing;
};

int wf_extend_blocks_per_sm(bool instrument, bool wide, bool top);
cudaError_t wf_clear_accum(const Launcher& L, const WaveBuffers& wb, long long entries);
cudaError_t wf_seed_pixels(const Launcher& L, const SceneView& sc, const WaveBuffers& wb, const FrameParams& fp);
cudaError_t wf_render_batch(const Launcher& L, const SceneView& sc, const WaveBuffers& wb, const FrameParams& fp);
cudaError_t wf_resolve_frame(const Launcher& L, const WaveBuffers& wb, const FrameParams& fp, bool add_to_sum);
cudaError_t wf_preview(const Launcher& L, const SceneView& sc, const WaveBuffers& wb, const FrameParams& fp);
cudaError_t wf_finalize(const Launcher& L, const uint32_t* frame_sum, uint8_t* out, int w, int h, int frames);
cudaError_t wf_first_hit(const Launcher& L, const SceneView& sc, const FrameParams& fp, int mode, const HookBuffers& hb,
                         int32_t* tri_id, float* dst);
cudaError_t wf_trace_rays(const Launcher& L, const SceneView& sc, const float* o, const float* d, int64_t n,
                          const HookBuffers& hb, int32_t* tri, float* dst, float* bu, float* bv);
cudaError_t wf_scatter_rows(const Launcher& L, const uint32_t* compact, uint32_t* full, const int32_t* rows,
                            int nrows, int width);
cudaError_t wf_gather_rows(const Launcher& L, const uint32_t* full, uint32_t* compact, const int32_t* rows,
                           int nrows, int width);

}  // namespace rt
