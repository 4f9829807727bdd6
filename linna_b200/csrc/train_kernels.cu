// Emulator-training kernels that follow the fused forward/loss/backward-data pass (fused_ffma.cu):
//
//   wgrad_kernel  : dW[n][k] = gscale * sum_b gz[b][n] * x[b][k]   for every linear map of the network in
//                   ONE launch (a flat list of 64x64 output tiles over all layers; the bias gradient is
//                   the extra column k == K of x, preset to 1).  With `fuse` the AdamW update
//                   (torch.optim.AdamW semantics, linna/predictor_gpu.py:267) runs in the epilogue and
//                   the new weight is written to the flat parameter vector and to both packed operand
//                   copies, so the gradient never touches HBM.
//   adamw_kernel  : the stand-alone update used after an NCCL all-reduce of the flat gradient.
//   mean_kernel   : deterministic mean of the per-row losses (Loss_fn, linna/util.py:1114-1115).
#include "linna_device.cuh"

namespace linna {

constexpr int WG_T = 64;   // output tile edge
constexpr int WG_BB = 32;  // batch rows per pipeline slab

__device__ __forceinline__ void wg_cp16(float *smem_dst, const float *gsrc, bool valid)
{
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz) : "memory");
}

__device__ __forceinline__ void adamw_apply(const AdamArgs &a, int idx, float g)
{
    float p = a.params[idx], m = a.m[idx], v = a.v[idx];
    p *= 1.0f - a.lr * a.wd;                       // decoupled weight decay
    m = m + (1.0f - a.beta1) * (g - m);            // exp_avg.lerp_(grad, 1-beta1)
    v = v * a.beta2 + (1.0f - a.beta2) * g * g;    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p = p - (a.lr / a.bc1) * (m / denom);
    a.params[idx] = p, a.m[idx] = m, a.v[idx] = v;
    a.blob[a.map_fwd[idx]] = p;
    const int mb = a.map_bwd[idx];
    if (mb >= 0) a.blob[mb] = p;
}

__global__ void __launch_bounds__(256) wgrad_kernel(const WgradLayer *__restrict__ layers, const WgradTile *__restrict__ tiles,
                                                    const float *__restrict__ rm_base, int B, const AdamArgs ad)
{
    __shared__ __align__(16) float Gs[2][WG_BB][WG_T];
    __shared__ __align__(16) float Xs[2][WG_BB][WG_T];
    const WgradTile t = tiles[blockIdx.x];
    const WgradLayer L = layers[t.layer];
    const int tid = threadIdx.x;
    const int tn = tid >> 4, tk = tid & 15;
    const float *G = rm_base + L.gz_off;
    const float *X = rm_base + L.x_off;

    auto load = [&](int slab, int stage) {
        const int b0 = slab * WG_BB;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int f = tid + i * 256;
            const int r = f >> 4, c4 = f & 15;
            const int b = b0 + r;
            const int cn = t.n0 + 4 * c4, ck = t.k0 + 4 * c4;
            const bool vg = b < B && cn < L.gz_ld;
            const bool vx = b < B && ck < L.x_ld;
            wg_cp16(&Gs[stage][r][4 * c4], vg ? G + (size_t)b * L.gz_ld + cn : G, vg);
            wg_cp16(&Xs[stage][r][4 * c4], vx ? X + (size_t)b * L.x_ld + ck : X, vx);
        }
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int nslab = (B + WG_BB - 1) / WG_BB;
    load(0, 0);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (int s = 0; s < nslab; ++s) {
        if (s + 1 < nslab) load(s + 1, (s + 1) & 1);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncthreads();
        const int st = s & 1;
#pragma unroll
        for (int b = 0; b < WG_BB; ++b) {
            const float4 a = *reinterpret_cast<const float4 *>(&Gs[st][b][4 * tn]);
            const float4 x = *reinterpret_cast<const float4 *>(&Xs[st][b][4 * tk]);
            const float av[4] = {a.x, a.y, a.z, a.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], xv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = t.n0 + 4 * tn + i;
        if (n >= L.N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = t.k0 + 4 * tk + j;
            int idx;
            if (k < L.K) idx = L.w_flat + n * L.K + k;
            else if (k == L.K && L.b_flat >= 0) idx = L.b_flat + n;
            else continue;
            const float g = L.gscale * acc[i][j];
            if (ad.fuse) adamw_apply(ad, idx, g);
            else ad.grads[idx] = g;
        }
    }
}

__global__ void adamw_kernel(const AdamArgs ad, int n_params)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_params; i += gridDim.x * blockDim.x)
        adamw_apply(ad, i, ad.grads[i]);
}

__global__ void mean_kernel(const float *__restrict__ x, int n, float *__restrict__ out)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += (double)x[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)(sh[0] / (double)n);
}

__global__ void fill_col_kernel(float *base, int ld, int col, int rows, float value)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x)
        base[(size_t)r * ld + col] = value;
}

__global__ void scatter_params_kernel(const float *__restrict__ params, float *blob, const int32_t *__restrict__ map_fwd,
                                      const int32_t *__restrict__ map_bwd, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float p = params[i];
        blob[map_fwd[i]] = p;
        const int mb = map_bwd[i];
        if (mb >= 0) blob[mb] = p;
    }
}

cudaError_t launch_scatter_params(const float *params, float *blob, const int32_t *map_fwd, const int32_t *map_bwd,
                                  int n_params, cudaStream_t stream)
{
    scatter_params_kernel<<<(n_params + 255) / 256, 256, 0, stream>>>(params, blob, map_fwd, map_bwd, n_params);
    return cudaGetLastError();
}

cudaError_t launch_wgrad(const WgradLayer *layers, const WgradTile *tiles, int n_tiles, const float *rm_base, int B,
                         const AdamArgs &ad, cudaStream_t stream)
{
    wgrad_kernel<<<n_tiles, 256, 0, stream>>>(layers, tiles, rm_base, B, ad);
    return cudaGetLastError();
}

cudaError_t launch_adamw(const AdamArgs &ad, int n_params, int num_sms, cudaStream_t stream)
{
    int blocks = (n_params + 255) / 256;
    if (blocks > num_sms * 8) blocks = num_sms * 8;
    adamw_kernel<<<blocks, 256, 0, stream>>>(ad, n_params);
    return cudaGetLastError();
}

cudaError_t launch_mean(const float *x, int n, float *out, cudaStream_t stream)
{
    mean_kernel<<<1, 256, 0, stream>>>(x, n, out);
    return cudaGetLastError();
}

cudaError_t launch_fill_col(float *base, int ld, int col, int rows, float value, cudaStream_t stream)
{
    fill_col_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(base, ld, col, rows, value);
    return cudaGetLastError();
}

}  // namespace linna
