// On-device sampler steps: walker positions, momenta, lnP, accept/reject and the random numbers never leave the GPU.
//
//   * affine-invariant ensemble sampler (Goodman & Weare stretch move; what emcee.EnsembleSampler runs by default,
//     linna/sampler.py:493-503): proposal and accept/reject for one half-ensemble -- a half-step is three launches
//     (propose, fused likelihood, accept) instead of ~20 tensor ops;
//   * batched Hamiltonian Monte Carlo (linna/HMCSampler.py:23-66, every chain with its own Metropolis test): momentum
//     draw + Hamiltonian + first half kick + drift in one launch, the inner kick + drift in one, the last half kick +
//     Hamiltonian + Metropolis select in one -- a leapfrog step is the fused lnP+gradient launch plus ONE of these;
//   * the half-chain mean / std shift test of the convergence check (linna/sampler.py:370-387) as a two-pass
//     reduction in float64.
//
// One WARP owns one walker / chain: lane 0 draws the walker's random numbers once (Philox4x32-10 keyed by (seed, walker,
// offset): reproducible, independent of the grid) and takes the accept decision once; the lanes then move the d
// coordinates with coalesced accesses.  Position and lnP of a walker are therefore always updated by the same decision.
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/linna_b200.h"

namespace {

constexpr int WPB = 8;   // warps (walkers) per block

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(32 * WPB) stretch_propose_kernel(const float *__restrict__ x, int d, const int64_t *__restrict__ first,
                                                                   const int64_t *__restrict__ second, int64_t ns, int64_t n_second,
                                                                   float a, unsigned long long seed, unsigned long long offset,
                                                                   float *__restrict__ y, float *__restrict__ z)
{
    const int lane = threadIdx.x & 31;
    for (int64_t i = blockIdx.x * (int64_t)WPB + (threadIdx.x >> 5); i < ns; i += (int64_t)gridDim.x * WPB) {
        float zz = 0.f;
        long long partner = 0;
        if (lane == 0) {
            curandStatePhilox4_32_10_t st;
            curand_init(seed, (unsigned long long)i, offset, &st);
            const float4 r = curand_uniform4(&st);                       // (0, 1]
            int64_t pick = (int64_t)(r.x * (float)n_second);
            if (pick >= n_second) pick = n_second - 1;
            partner = second[pick];
            const float t = (a - 1.0f) * r.y + 1.0f;
            zz = t * t / a;                                              // g(z) ~ 1/sqrt(z) on [1/a, a]
            z[i] = zz;
        }
        zz = __shfl_sync(0xffffffffu, zz, 0);
        partner = __shfl_sync(0xffffffffu, partner, 0);
        const float *xc = x + partner * d, *xw = x + first[i] * d;
        for (int k = lane; k < d; k += 32) {
            const float c = xc[k];
            y[i * d + k] = c + zz * (xw[k] - c);
        }
    }
}

__global__ void __launch_bounds__(32 * WPB) stretch_accept_kernel(float *__restrict__ x, float *__restrict__ lnp, float *__restrict__ naccepted,
                                                                  int d, const int64_t *__restrict__ first, int64_t ns,
                                                                  const float *__restrict__ y, const float *__restrict__ lnp_y,
                                                                  const float *__restrict__ z, unsigned long long seed,
                                                                  unsigned long long offset)
{
    const int lane = threadIdx.x & 31;
    for (int64_t i = blockIdx.x * (int64_t)WPB + (threadIdx.x >> 5); i < ns; i += (int64_t)gridDim.x * WPB) {
        const int64_t w = first[i];
        int acc = 0;
        if (lane == 0) {   // ONE decision per walker: position and lnP below both follow it
            curandStatePhilox4_32_10_t st;
            curand_init(seed, (unsigned long long)i, offset, &st);
            const float r = curand_uniform(&st);
            float ly = lnp_y[i];
            if (ly != ly) ly = -INFINITY;                                // NaN -> -inf (linna/util.py:1015-1016)
            const float lnq = (float)(d - 1) * logf(z[i]) + ly - lnp[w];
            acc = ((logf(r) < lnq) && (fabsf(ly) <= 3.0e38f)) ? 1 : 0;
            if (acc) {
                lnp[w] = ly;
                naccepted[w] += 1.0f;
            }
        }
        acc = __shfl_sync(0xffffffffu, acc, 0);
        if (acc)
            for (int k = lane; k < d; k += 32) x[w * d + k] = y[i * d + k];
    }
}

// ------------------------------------------------------------------------------------------ batched HMC
// begin: p ~ N(0, m); H0 = sum p^2 / 2m - lnP; p += eps/2 * grad; xn = x + eps * p / m        (HMCSampler.py:25-36)
__global__ void __launch_bounds__(32 * WPB) hmc_begin_kernel(const float *__restrict__ x, const float *__restrict__ lnp,
                                                             const float *__restrict__ grad, const float *__restrict__ mass, int d,
                                                             int64_t nc, float eps, unsigned long long seed, unsigned long long offset,
                                                             float *__restrict__ p, float *__restrict__ xn, float *__restrict__ H0)
{
    const int lane = threadIdx.x & 31;
    for (int64_t c = blockIdx.x * (int64_t)WPB + (threadIdx.x >> 5); c < nc; c += (int64_t)gridDim.x * WPB) {
        float kin = 0.f;
        for (int k0 = 0; k0 < d; k0 += 128) {   // a lane draws four normals per 128 coordinates
            const int kb = k0 + 4 * lane;
            if (kb < d) {
                curandStatePhilox4_32_10_t st;
                curand_init(seed, (unsigned long long)(c * ((d + 127) / 128 * 32) + (k0 >> 2) + lane), offset, &st);
                const float4 nz = curand_normal4(&st);
                const float nn[4] = {nz.x, nz.y, nz.z, nz.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = kb + e;
                    if (k < d) {
                        const float mk = mass[k];
                        float pk = nn[e] * sqrtf(mk);
                        kin += 0.5f * pk * pk / mk;
                        pk += 0.5f * eps * grad[c * d + k];
                        p[c * d + k] = pk;
                        xn[c * d + k] = x[c * d + k] + eps * (pk / mk);
                    }
                }
            }
        }
        kin = warp_sum(kin);
        if (lane == 0) H0[c] = kin - lnp[c];
    }
}

// inner leapfrog step: p += eps * grad(xn); xn += eps * p / m                                  (HMCSampler.py:43-46)
__global__ void hmc_step_kernel(float *__restrict__ p, float *__restrict__ xn, const float *__restrict__ grad,
                                const float *__restrict__ mass, int d, int64_t total, float eps)
{
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const float mk = mass[e % d];
        const float pk = p[e] + eps * grad[e];
        p[e] = pk;
        xn[e] += eps * (pk / mk);
    }
}

// end: p += eps/2 * grad(xn); H1 = sum p^2 / 2m - lnP(xn); accept with min(1, exp(H0 - H1)); select  (HMCSampler.py:51-59)
__global__ void __launch_bounds__(32 * WPB) hmc_end_kernel(float *__restrict__ x, float *__restrict__ lnp, float *__restrict__ grad,
                                                           const float *__restrict__ xn, const float *__restrict__ lnp_n,
                                                           const float *__restrict__ grad_n, const float *__restrict__ p,
                                                           const float *__restrict__ mass, const float *__restrict__ H0, int d,
                                                           int64_t nc, float eps, unsigned long long seed, unsigned long long offset,
                                                           float *__restrict__ naccepted)
{
    const int lane = threadIdx.x & 31;
    for (int64_t c = blockIdx.x * (int64_t)WPB + (threadIdx.x >> 5); c < nc; c += (int64_t)gridDim.x * WPB) {
        float kin = 0.f;
        for (int k = lane; k < d; k += 32) {
            const float pk = p[c * d + k] + 0.5f * eps * grad_n[c * d + k];
            kin += 0.5f * pk * pk / mass[k];
        }
        kin = warp_sum(kin);
        int acc = 0;
        if (lane == 0) {
            curandStatePhilox4_32_10_t st;
            curand_init(seed, (unsigned long long)c, offset, &st);
            const float r = curand_uniform(&st);
            const float l = lnp_n[c];
            const float dH = H0[c] - (kin - l);
            acc = (isfinite(l) && isfinite(dH) && r < expf(fminf(dH, 0.f))) ? 1 : 0;
            if (acc) {
                lnp[c] = l;
                naccepted[c] += 1.0f;
            }
        }
        acc = __shfl_sync(0xffffffffu, acc, 0);
        if (acc)
            for (int k = lane; k < d; k += 32) {
                x[c * d + k] = xn[c * d + k];
                grad[c * d + k] = grad_n[c * d + k];
            }
    }
}

// ------------------------------------------------------------------------------------------ half-chain moments
// rows [r0, r1) of a row-major [rows][d] matrix: per-block partial sums of (x - shift)^power in float64.
// pass 1: power 1, shift 0 -> means; pass 2: power 2, shift = mean -> variances.  Partials are added in block order on
// the host side of the C ABI: deterministic.
template <typename T>
__global__ void __launch_bounds__(256) moments_kernel(const T *__restrict__ x, int64_t r0, int64_t r1, int d, const double *__restrict__ shift,
                                                      int power, double *__restrict__ partial)
{
    __shared__ double sh[256];
    const int cols = d < 256 ? d : 256;
    const int lanes = 256 / cols;                  // row lanes per block
    const int col = threadIdx.x % cols, rl = threadIdx.x / cols;
    for (int c0 = 0; c0 < d; c0 += cols) {
        const int c = c0 + col;
        double s = 0.0;
        if (rl < lanes && c < d) {
            const double sf = shift ? shift[c] : 0.0;
            for (int64_t r = r0 + blockIdx.x * (int64_t)lanes + rl; r < r1; r += (int64_t)gridDim.x * lanes) {
                const double v = (double)x[r * d + c] - sf;
                s += power == 1 ? v : v * v;
            }
        }
        sh[threadIdx.x] = s;
        __syncthreads();
        if (rl == 0 && c < d) {
            double t = 0.0;
            for (int q = 0; q < lanes; ++q) t += sh[q * cols + col];
            partial[(size_t)blockIdx.x * d + c] = t;
        }
        __syncthreads();
    }
}

int grid_warps(int64_t n)
{
    int64_t g = (n + WPB - 1) / WPB;
    return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace

extern "C" int linna_stretch_propose(const float *x, int32_t d, const int64_t *first, const int64_t *second, int64_t ns,
                                     int64_t n_second, float a, uint64_t seed, uint64_t offset, float *y, float *z,
                                     void *stream)
{
    if (!x || !first || !second || !y || !z || d <= 0 || ns < 0 || n_second <= 0 || a <= 1.0f) return LINNA_EINVAL;
    if (ns == 0) return LINNA_OK;
    stretch_propose_kernel<<<grid_warps(ns), 32 * WPB, 0, (cudaStream_t)stream>>>(x, d, first, second, ns, n_second, a, seed, offset, y, z);
    return cudaGetLastError() == cudaSuccess ? LINNA_OK : LINNA_ECUDA;
}

extern "C" int linna_stretch_accept(float *x, float *lnp, float *naccepted, int32_t d, const int64_t *first, int64_t ns,
                                    const float *y, const float *lnp_y, const float *z, uint64_t seed, uint64_t offset,
                                    void *stream)
{
    if (!x || !lnp || !naccepted || !first || !y || !lnp_y || !z || d <= 0 || ns < 0) return LINNA_EINVAL;
    if (ns == 0) return LINNA_OK;
    stretch_accept_kernel<<<grid_warps(ns), 32 * WPB, 0, (cudaStream_t)stream>>>(x, lnp, naccepted, d, first, ns, y, lnp_y, z, seed, offset);
    return cudaGetLastError() == cudaSuccess ? LINNA_OK : LINNA_ECUDA;
}

extern "C" int linna_hmc_begin(const float *x, const float *lnp, const float *grad, const float *mass, int32_t d, int64_t nc, float eps,
                               uint64_t seed, uint64_t offset, float *p, float *xn, float *H0, void *stream)
{
    if (!x || !lnp || !grad || !mass || !p || !xn || !H0 || d <= 0 || nc < 0) return LINNA_EINVAL;
    if (nc == 0) return LINNA_OK;
    hmc_begin_kernel<<<grid_warps(nc), 32 * WPB, 0, (cudaStream_t)stream>>>(x, lnp, grad, mass, d, nc, eps, seed, offset, p, xn, H0);
    return cudaGetLastError() == cudaSuccess ? LINNA_OK : LINNA_ECUDA;
}

extern "C" int linna_hmc_step(float *p, float *xn, const float *grad, const float *mass, int32_t d, int64_t nc, float eps, void *stream)
{
    if (!p || !xn || !grad || !mass || d <= 0 || nc < 0) return LINNA_EINVAL;
    if (nc == 0) return LINNA_OK;
    const int64_t total = nc * d;
    int64_t g = (total + 255) / 256;
    g = g > 148 * 16 ? 148 * 16 : g;
    hmc_step_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(p, xn, grad, mass, d, total, eps);
    return cudaGetLastError() == cudaSuccess ? LINNA_OK : LINNA_ECUDA;
}

extern "C" int linna_hmc_end(float *x, float *lnp, float *grad, const float *xn, const float *lnp_n, const float *grad_n, const float *p,
                             const float *mass, const float *H0, int32_t d, int64_t nc, float eps, uint64_t seed, uint64_t offset,
                             float *naccepted, void *stream)
{
    if (!x || !lnp || !grad || !xn || !lnp_n || !grad_n || !p || !mass || !H0 || !naccepted || d <= 0 || nc < 0) return LINNA_EINVAL;
    if (nc == 0) return LINNA_OK;
    hmc_end_kernel<<<grid_warps(nc), 32 * WPB, 0, (cudaStream_t)stream>>>(x, lnp, grad, xn, lnp_n, grad_n, p, mass, H0, d, nc, eps, seed,
                                                                          offset, naccepted);
    return cudaGetLastError() == cudaSuccess ? LINNA_OK : LINNA_ECUDA;
}

// mean and (population) standard deviation per column of rows [r0, r1) of a DEVICE matrix x[rows][d] (float32 when
// is_double == 0): two passes in float64, results on the HOST.
extern "C" int linna_column_moments(const void *x, int32_t is_double, int64_t r0, int64_t r1, int32_t d, double *mean_host,
                                    double *std_host, void *stream)
{
    if (!x || !mean_host || !std_host || d <= 0 || r1 <= r0) return LINNA_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = 296;
    double *partial = nullptr, *mean_dev = nullptr;
    if (cudaMalloc(&partial, (size_t)nblk * d * sizeof(double)) != cudaSuccess) return LINNA_ENOMEM;
    if (cudaMalloc(&mean_dev, (size_t)d * sizeof(double)) != cudaSuccess) { cudaFree(partial); return LINNA_ENOMEM; }
    double *ph = new double[(size_t)nblk * d];
    const double n = (double)(r1 - r0);
    int rc = LINNA_OK;
    for (int pass = 1; pass <= 2 && rc == LINNA_OK; ++pass) {
        if (is_double) moments_kernel<double><<<nblk, 256, 0, st>>>((const double *)x, r0, r1, d, pass == 2 ? mean_dev : nullptr, pass, partial);
        else moments_kernel<float><<<nblk, 256, 0, st>>>((const float *)x, r0, r1, d, pass == 2 ? mean_dev : nullptr, pass, partial);
        if (cudaMemcpyAsync(ph, partial, (size_t)nblk * d * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { rc = LINNA_ECUDA; break; }
        for (int c = 0; c < d; ++c) {
            double s = 0.0;
            for (int b = 0; b < nblk; ++b) s += ph[(size_t)b * d + c];
            if (pass == 1) mean_host[c] = s / n;
            else std_host[c] = sqrt(s / n);
        }
        if (pass == 1 && cudaMemcpyAsync(mean_dev, mean_host, (size_t)d * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = LINNA_ECUDA;
    }
    delete[] ph;
    cudaFree(partial), cudaFree(mean_dev);
    return rc;
}
