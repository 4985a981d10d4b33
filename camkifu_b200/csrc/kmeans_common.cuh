// Shared pieces of the k-means kernels (kmeans.cu: one CTA per (frame, attempt), any input; kmeans_cluster.cu: one
// thread-block cluster per (frame, attempt), uint8 input held in distributed shared memory).
#pragma once
#include <float.h>

#include "ckb_common.cuh"

#define KM_MAX_ITER 100                // criteria type has EPS only => maxCount = 100 (the "15" is ignored)
#define KM_EPS2 9.0                    // (eps = 3)^2
#define KM_ITERS_FALLBACK (-1)         // KmAttempt.iters: the cluster kernel declined this attempt (see kmeans_cluster.cu)

struct KmAttempt {
    double compactness;
    float centers[9];      // final centres
    float old_centers[9];  // the centres the returned labels were assigned against
    int n_fix;             // empty-cluster repairs of the last iteration (label overrides)
    int fix_idx[2];
    int fix_k[2];
    int iters;
};

struct Region {
    int x0, y0, h, w, N;   // bounding box of the zones [rs, re) x [cs, ce) in the canonical image (cluster_colors' sub-image)
    int rs, re, cs, ce;
};

#define CKB_MAX_REGIONS 16
struct RegionSet {         // the regions of one batched call (SfMeta: 3 x 3), passed to the kernels by value
    Region r[CKB_MAX_REGIONS];
    int n;
};

// ----------------------------------------------------------------------------------------------------------- helpers
__device__ __forceinline__ float dist3(float ax, float ay, float az, float bx, float by, float bz)
{
    // hal::normL2Sqr_ for n = 3: ((t0*t0 + t1*t1) + t2*t2) in float32 without contraction
    const float t0 = __fsub_rn(ax, bx), t1 = __fsub_rn(ay, by), t2 = __fsub_rn(az, bz);
    float d = __fmul_rn(t0, t0);
    d = __fadd_rn(d, __fmul_rn(t1, t1));
    d = __fadd_rn(d, __fmul_rn(t2, t2));
    return d;
}

template <bool F32>
__device__ __forceinline__ float3 load_px(const void *scratch, int i)
{
    if (F32) {
        const float4 v = __ldg((const float4 *)scratch + i);
        return make_float3(v.x, v.y, v.z);
    } else {
        const uchar4 v = __ldg((const uchar4 *)scratch + i);
        return make_float3((float)v.x, (float)v.y, (float)v.z);
    }
}

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_incl_scan_d(double v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

__device__ __forceinline__ uint32_t rng_next(uint64_t &s)
{
    s = (uint64_t)(uint32_t)s * 4164903690ULL + (s >> 32);
    return (uint32_t)s;
}

__device__ __forceinline__ double rng_double(uint64_t &s)
{
    const uint32_t t = rng_next(s);
    const uint64_t v = ((uint64_t)t << 32) | rng_next(s);
    return __dmul_rn(__ull2double_rn(v), 5.4210108624275221700372640043497e-20);
}

// label = first minimum of the three float32 distances (KMeansDistanceComputer: `if (min_dist > dist)`)
__device__ __forceinline__ int argmin3(float3 x, const float *c)
{
    const float d0 = dist3(x.x, x.y, x.z, c[0], c[1], c[2]);
    const float d1 = dist3(x.x, x.y, x.z, c[3], c[4], c[5]);
    const float d2 = dist3(x.x, x.y, x.z, c[6], c[7], c[8]);
    int k = 0;
    float m = d0;
    if (m > d1) { m = d1; k = 1; }
    if (m > d2) { k = 2; }
    return k;
}

