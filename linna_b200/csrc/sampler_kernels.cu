// On-device step of the affine-invariant ensemble sampler (Goodman & Weare stretch move; what
// emcee.EnsembleSampler runs by default, linna/sampler.py:493-503): proposal and accept/reject for one
// half-ensemble, so that walker positions, lnP and the random numbers never leave the GPU and a half-step is
// three launches (propose, fused likelihood, accept) instead of ~20 tensor ops.
//
// Random numbers are Philox4x32-10 keyed by (seed, walker slot, offset): the proposal and the acceptance draw use
// different offsets, every walker its own subsequence, so the stream is reproducible and independent of the grid.
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>

#include "../../include/linna_b200.h"

namespace {

__global__ void stretch_propose_kernel(const float *__restrict__ x, int d, const int64_t *__restrict__ first,
                                       const int64_t *__restrict__ second, int64_t ns, int64_t n_second, float a,
                                       unsigned long long seed, unsigned long long offset, float *__restrict__ y,
                                       float *__restrict__ z)
{
    const int64_t total = ns * d;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / d;
        const int k = (int)(e - i * d);
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)i, offset, &st);
        const float4 r = curand_uniform4(&st);                       // (0, 1]
        int64_t pick = (int64_t)(r.x * (float)n_second);
        if (pick >= n_second) pick = n_second - 1;
        const int64_t partner = second[pick];
        const float t = (a - 1.0f) * r.y + 1.0f;
        const float zz = t * t / a;                                  // g(z) ~ 1/sqrt(z) on [1/a, a]
        const float c = x[partner * d + k];
        y[e] = c + zz * (x[first[i] * d + k] - c);
        if (k == 0) z[i] = zz;
    }
}

__global__ void stretch_accept_kernel(float *__restrict__ x, float *__restrict__ lnp, float *__restrict__ naccepted, int d,
                                      const int64_t *__restrict__ first, int64_t ns, const float *__restrict__ y,
                                      const float *__restrict__ lnp_y, const float *__restrict__ z, unsigned long long seed,
                                      unsigned long long offset)
{
    const int64_t total = ns * d;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / d;
        const int k = (int)(e - i * d);
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)i, offset, &st);
        const float r = curand_uniform(&st);
        const int64_t w = first[i];
        float ly = lnp_y[i];
        if (ly != ly) ly = -INFINITY;                                // NaN -> -inf (linna/util.py:1015-1016)
        const float lnq = (float)(d - 1) * logf(z[i]) + ly - lnp[w];
        const bool acc = (logf(r) < lnq) && (fabsf(ly) <= 3.0e38f);
        // every thread of walker i reads lnp[w] before any of them may overwrite it: the writer is thread k == 0 of
        // the same walker and the read above precedes its write in program order; other walkers never touch w
        if (acc) x[w * d + k] = y[e];
        if (k == 0 && acc) {
            naccepted[w] += 1.0f;
        }
    }
}

// lnP of accepted walkers is written by a second tiny pass so that no thread can read an already-updated lnp[w]
__global__ void stretch_commit_lnp_kernel(float *__restrict__ lnp, const int64_t *__restrict__ first, int64_t ns, int d,
                                          const float *__restrict__ lnp_y, const float *__restrict__ z,
                                          unsigned long long seed, unsigned long long offset)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ns; i += (int64_t)gridDim.x * blockDim.x) {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)i, offset, &st);
        const float r = curand_uniform(&st);
        const int64_t w = first[i];
        float ly = lnp_y[i];
        if (ly != ly) ly = -INFINITY;
        const float lnq = (float)(d - 1) * logf(z[i]) + ly - lnp[w];
        if ((logf(r) < lnq) && (fabsf(ly) <= 3.0e38f)) lnp[w] = ly;
    }
}

int grid_for(int64_t work)
{
    int64_t g = (work + 255) / 256;
    return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

}  // namespace

extern "C" int linna_stretch_propose(const float *x, int32_t d, const int64_t *first, const int64_t *second, int64_t ns,
                                     int64_t n_second, float a, uint64_t seed, uint64_t offset, float *y, float *z,
                                     void *stream)
{
    if (!x || !first || !second || !y || !z || d <= 0 || ns < 0 || n_second <= 0 || a <= 1.0f) return LINNA_EINVAL;
    if (ns == 0) return LINNA_OK;
    stretch_propose_kernel<<<grid_for(ns * d), 256, 0, (cudaStream_t)stream>>>(x, d, first, second, ns, n_second, a, seed, offset,
                                                                               y, z);
    return cudaGetLastError() == cudaSuccess ? LINNA_OK : LINNA_ECUDA;
}

extern "C" int linna_stretch_accept(float *x, float *lnp, float *naccepted, int32_t d, const int64_t *first, int64_t ns,
                                    const float *y, const float *lnp_y, const float *z, uint64_t seed, uint64_t offset,
                                    void *stream)
{
    if (!x || !lnp || !naccepted || !first || !y || !lnp_y || !z || d <= 0 || ns < 0) return LINNA_EINVAL;
    if (ns == 0) return LINNA_OK;
    stretch_accept_kernel<<<grid_for(ns * d), 256, 0, (cudaStream_t)stream>>>(x, lnp, naccepted, d, first, ns, y, lnp_y, z, seed,
                                                                              offset);
    stretch_commit_lnp_kernel<<<grid_for(ns), 256, 0, (cudaStream_t)stream>>>(lnp, first, ns, d, lnp_y, z, seed, offset);
    return cudaGetLastError() == cudaSuccess ? LINNA_OK : LINNA_ECUDA;
}
