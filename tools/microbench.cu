// microbench.cu — the ceilings the traversal roofline is measured against ("bound = L2 bandwidth
// (builder microbenchmarks it) or FP32 issue, whichever is lower"; §6: "builder must microbenchmark" the FP32 peak).
//
// Built by `make -C tools` into tools/libmicrobench.so (sm_100a), loaded by bench.py through ctypes on rank 0 and
// run for about a second before the timed region.  Nothing in the product links it.
//
//   mb_gather_gbs   random gathers of 32 B or 64 B records (one or two 256-bit loads per record, the way k_extend
//                   fetches nodes) from a table of a given size: 8 MB = L2-resident like the config-2 BVH, 1 GB =
//                   HBM-resident like config 4's.  `dependent` = 1 chains every load's address on the previous
//                   record (a ray's walk down the tree); 0 issues four independent gathers per step (the most
//                   the memory system delivers to this access pattern).
//   mb_issue_tops   thread-instructions per second of one instruction class, 8 independent chains per thread,
//                   full occupancy: 0 FADD/FMUL alternating (what -fmad=false code issues), 1 FFMA, 2 FMNMX,
//                   3 I2F.U16 (the dequantisation of a box plane), 4 the slab mix of k_extend's node step
//                   (2 I2F + 2 FFMA + 2 FMNMX per plane pair).
//   mb_stream_gbs   a plain float4 copy (read + write bytes) — the same thing MEASURED_PEAKS.json's hbm_gbs is.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {  // lowbias32
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&v)[8]) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}

template <int REC, bool DEP>
__global__ void __launch_bounds__(128) k_gather(const uint4* __restrict__ table, uint32_t nrec, int steps,
                                                uint32_t* __restrict__ sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int MLP = DEP ? 1 : 4;
    uint32_t idx[MLP];
#pragma unroll
    for (int k = 0; k < MLP; k++) idx[k] = mix32(tid * MLP + k + 1u);
    uint32_t acc = 0;
    for (int s = 0; s < steps; s++) {
#pragma unroll
        for (int k = 0; k < MLP; k++) {
            const uint32_t r = (uint32_t)(((uint64_t)idx[k] * nrec) >> 32);
            const uint4* p = table + (size_t)r * (REC / 16);
            uint32_t a[8];
            ldg256(p, a);
            uint32_t x = a[0] ^ a[3] ^ a[7];
            if (REC == 64) {
                uint32_t b[8];
                ldg256(p + 2, b);
                x ^= b[1] ^ b[6];
            }
            acc ^= x;
            idx[k] = DEP ? mix32(idx[k] ^ x) : mix32(idx[k] + 0x9e3779b9u);
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;  // keeps the loads alive
}

template <int KIND>
__global__ void __launch_bounds__(256) k_issue(int iters, float seed, float* __restrict__ sink) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = seed + (float)(threadIdx.x + k) * 1e-3f;
    const float a = 1.0000001f, b = 1e-7f;
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] = (__float_as_uint(seed) | 0x10001u) + 77u * k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (KIND == 0) { v[k] = __fadd_rn(v[k], b); v[k] = __fmul_rn(v[k], a); }
            if (KIND == 1) { v[k] = __fmaf_rn(v[k], a, b); v[k] = __fmaf_rn(v[k], a, b); }
            if (KIND == 2) { v[k] = fminf(v[k], v[(k + 1) & 7]); v[k] = fmaxf(v[k], v[(k + 3) & 7]); }
            if (KIND == 3) {
                float f0, f1;
                asm volatile("cvt.rn.f32.u16 %0, %1;" : "=f"(f0) : "h"((unsigned short)w[k]));
                asm volatile("cvt.rn.f32.u16 %0, %1;" : "=f"(f1) : "h"((unsigned short)(w[k] >> 16)));
                v[k] = __uint_as_float(__float_as_uint(v[k]) ^ __float_as_uint(f0) ^ __float_as_uint(f1));
            }
            if (KIND == 4) {  // one plane pair of the node step: dequantise, t = fma(q, inv, c), order
                float q0, q1;
                asm volatile("cvt.rn.f32.u16 %0, %1;" : "=f"(q0) : "h"((unsigned short)w[k]));
                asm volatile("cvt.rn.f32.u16 %0, %1;" : "=f"(q1) : "h"((unsigned short)(w[k] >> 16)));
                const float t0 = __fmaf_rn(q0, a, v[k]), t1 = __fmaf_rn(q1, a, v[(k + 1) & 7]);
                v[k] = fminf(t0, v[(k + 2) & 7]);
                v[(k + 1) & 7] = fmaxf(t1, v[(k + 3) & 7]);
            }
        }
        if (KIND >= 3) {
#pragma unroll
            for (int k = 0; k < 8; k++) w[k] = w[k] * 1664525u + 1013904223u;  // one IMAD per two conversions
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += v[k];
    if (s == 123.456f) sink[0] = s;
}

__global__ void __launch_bounds__(256) k_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void k_fill(uint4* t, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        t[i] = make_uint4(mix32((uint32_t)i), mix32((uint32_t)i + 1u), (uint32_t)i, ~(uint32_t)i);
}

int sm_count() {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

template <class F>
double best_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();  // warm-up
    cudaDeviceSynchronize();
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

}  // namespace

extern "C" {

// GB/s of record bytes delivered; blocks_per_sm x 128 threads per SM resident (k_extend runs 9-10 x 128)
double mb_gather_gbs(size_t table_bytes, int record_bytes, int dependent, int blocks_per_sm, int steps) {
    if ((record_bytes != 32 && record_bytes != 64) || table_bytes < 4096 || blocks_per_sm <= 0 || steps <= 0) return -1.0;
    uint4* table = nullptr;
    uint32_t* sink = nullptr;
    if (cudaMalloc(&table, table_bytes) != cudaSuccess) { cudaGetLastError(); return -2.0; }
    cudaMalloc(&sink, 4);
    k_fill<<<sm_count() * 8, 256>>>(table, table_bytes / 16);
    const uint32_t nrec = (uint32_t)(table_bytes / record_bytes);
    const int grid = sm_count() * blocks_per_sm;
    auto launch = [&]() {
        if (record_bytes == 32) {
            if (dependent) k_gather<32, true><<<grid, 128>>>(table, nrec, steps, sink);
            else k_gather<32, false><<<grid, 128>>>(table, nrec, steps, sink);
        } else {
            if (dependent) k_gather<64, true><<<grid, 128>>>(table, nrec, steps, sink);
            else k_gather<64, false><<<grid, 128>>>(table, nrec, steps, sink);
        }
    };
    const double ms = best_ms(launch, 5);
    const double loads = (double)grid * 128.0 * steps * (dependent ? 1.0 : 4.0);
    cudaFree(table);
    cudaFree(sink);
    if (cudaGetLastError() != cudaSuccess) return -3.0;
    return loads * record_bytes / (ms * 1e-3) * 1e-9;
}

// 1e12 thread-instructions per second of the class (each FADD / FMUL / FFMA / FMNMX / I2F counts one)
double mb_issue_tops(int kind, int iters) {
    float* sink = nullptr;
    cudaMalloc(&sink, 4);
    const int grid = sm_count() * 8;
    auto launch = [&]() {
        switch (kind) {
            case 0: k_issue<0><<<grid, 256>>>(iters, 1.0f, sink); break;
            case 1: k_issue<1><<<grid, 256>>>(iters, 1.0f, sink); break;
            case 2: k_issue<2><<<grid, 256>>>(iters, 1.0f, sink); break;
            case 3: k_issue<3><<<grid, 256>>>(iters, 1.0f, sink); break;
            default: k_issue<4><<<grid, 256>>>(iters, 1.0f, sink); break;
        }
    };
    const double ms = best_ms(launch, 5);
    cudaFree(sink);
    static const double perIter[5] = {16.0, 16.0, 16.0, 16.0, 48.0};  // counted instructions per loop iteration per thread
    if (kind < 0 || kind > 4 || cudaGetLastError() != cudaSuccess) return -1.0;
    return (double)grid * 256.0 * iters * perIter[kind] / (ms * 1e-3) * 1e-12;
}

double mb_stream_gbs(size_t bytes) {
    float4 *a = nullptr, *b = nullptr;
    if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, bytes) != cudaSuccess) { cudaGetLastError(); cudaFree(a); return -2.0; }
    cudaMemset(a, 1, bytes);
    const size_t n = bytes / 16;
    auto launch = [&]() { k_copy<<<sm_count() * 16, 256>>>(a, b, n); };
    const double ms = best_ms(launch, 5);
    cudaFree(a);
    cudaFree(b);
    return 2.0 * (double)bytes / (ms * 1e-3) * 1e-9;
}

}  // extern "C"
