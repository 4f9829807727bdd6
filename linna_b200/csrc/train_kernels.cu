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
#include "../../include/linna_b200.h"

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

// Data-parallel optimiser step with the gradient all-reduce INSIDE it (NVLink peer memory, no NCCL call): every rank's
// gradient-out step leaves its flat gradient in a symmetric buffer; this kernel (1) tells every peer that this rank's
// gradient is complete (a release store of the sequence number into the peer's signal pad -- the kernel is stream-ordered
// behind the weight-gradient launch), (2) waits until every peer has said the same, (3) averages the world's gradients in
// rank order -- straight from the peers' memory, 16-byte loads over NVLink -- and applies AdamW.  Every rank adds the
// same numbers in the same order, so the replicas stay bit-identical.
//   one phase (world <= 2): every rank reads all the peers' gradients itself.
//   two phases (avg_offset >= 0): rank r averages only ITS slice of the vector (reduce-scatter by pull) into the `avg`
//     region of its buffer; when the whole grid has done so (arrival counter) the last CTA tells the peers, and after that
//     second barrier every rank applies AdamW to the full vector, fetching each slice's average from its owner: ~2 x the
//     vector in remote reads whatever the world size, instead of (world - 1) x.
// The gradient buffers are double-buffered by the caller: a peer may still read this rank's buffer of step t while the
// rank writes step t + 1, and it cannot still read it at step t + 2 because this rank's barrier of step t + 1 needed that
// peer's signal, sent after its kernel of step t; the same argument covers the `avg` region.
__device__ __forceinline__ void peer_signal(const PeerReduce &pr, int slot)
{
    __threadfence_system();
    uint32_t *dst = pr.pads[threadIdx.x] + slot + pr.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(pr.token) : "memory");
}
__device__ __forceinline__ void peer_wait(const PeerReduce &pr, int slot)
{
    const uint32_t *src = pr.pads[pr.rank] + slot + threadIdx.x;
    const long long t0 = clock64();
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
        if ((int32_t)(v - pr.token) >= 0) break;              // a peer may already be one step ahead
        if (clock64() - t0 > 20000000000LL) __trap();         // ~10 s: a rank is missing
        __nanosleep(100);
    }
}

__global__ void __launch_bounds__(256) adamw_peer_kernel(const AdamArgs ad, int n_params, const PeerReduce pr)
{
    __shared__ const float *s_g[16];
    __shared__ int s_last;
    if (blockIdx.x == 0 && (int)threadIdx.x < pr.world) peer_signal(pr, pr.slot);
    if ((int)threadIdx.x < pr.world) {
        peer_wait(pr, pr.slot);
        s_g[threadIdx.x] = pr.grads[threadIdx.x];
    }
    __syncthreads();
    const float inv = 1.0f / (float)pr.world;
    const int n4 = n_params >> 2;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    if (pr.avg_offset < 0) {
        for (int i = tid; i < n4; i += nthr) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < pr.world; ++r) {
                const float4 v = __ldcg(reinterpret_cast<const float4 *>(s_g[r] + pr.offset) + i);
                g.x += v.x, g.y += v.y, g.z += v.z, g.w += v.w;
            }
            adamw_apply(ad, 4 * i, g.x * inv), adamw_apply(ad, 4 * i + 1, g.y * inv);
            adamw_apply(ad, 4 * i + 2, g.z * inv), adamw_apply(ad, 4 * i + 3, g.w * inv);
        }
        for (int i = 4 * n4 + tid; i < n_params; i += nthr) {
            float g = 0.f;
            for (int r = 0; r < pr.world; ++r) g += __ldcg(s_g[r] + pr.offset + i);
            adamw_apply(ad, i, g * inv);
        }
        return;
    }
    // ---- phase 1: this rank's slice, averaged into its own `avg` region (quads of 4 floats; the tail quad is padded)
    const int nq = (n_params + 3) >> 2;                               // quads of the (padded) vector
    const int qchunk = (nq + pr.world - 1) / pr.world;                // quads per rank
    const int q0 = pr.rank * qchunk, q1 = min(nq, q0 + qchunk);
    float4 *avg_mine = reinterpret_cast<float4 *>(const_cast<float *>(s_g[pr.rank]) + pr.avg_offset);
    for (int i = q0 + tid; i < q1; i += nthr) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < pr.world; ++r) {
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(s_g[r] + pr.offset) + i);
            g.x += v.x, g.y += v.y, g.z += v.z, g.w += v.w;
        }
        avg_mine[i] = make_float4(g.x * inv, g.y * inv, g.z * inv, g.w * inv);
    }
    // ---- the whole grid has written its part -> tell the peers; then wait for theirs
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(pr.ticket, 1);
        s_last = t == (int)gridDim.x - 1;
        if (s_last) *pr.ticket = 0;      // ready for the next launch (stream-ordered)
        __threadfence_system();
    }
    __syncthreads();
    if (s_last && (int)threadIdx.x < pr.world) peer_signal(pr, pr.slot + 16);
    if ((int)threadIdx.x < pr.world) peer_wait(pr, pr.slot + 16);
    __syncthreads();
    // ---- phase 2: AdamW on the full vector, every slice's average from its owner
    for (int i = tid; i < nq; i += nthr) {
        const int owner = i / qchunk;
        const float4 g = __ldcg(reinterpret_cast<const float4 *>(s_g[owner] + pr.avg_offset) + i);
        if (4 * i + 3 < n_params) {
            adamw_apply(ad, 4 * i, g.x), adamw_apply(ad, 4 * i + 1, g.y), adamw_apply(ad, 4 * i + 2, g.z), adamw_apply(ad, 4 * i + 3, g.w);
        } else {
            const float gv[4] = {g.x, g.y, g.z, g.w};
            for (int e = 0; e < 4 && 4 * i + e < n_params; ++e) adamw_apply(ad, 4 * i + e, gv[e]);
        }
    }
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

cudaError_t launch_adamw_peer(const AdamArgs &ad, int n_params, int num_sms, const PeerReduce &pr, cudaStream_t stream)
{
    int blocks = (n_params / 4 + 255) / 256;
    // the two-phase form holds a grid-wide barrier: every CTA must be resident (256 threads, few registers: 2 per SM is safe)
    const int cap = pr.avg_offset >= 0 ? num_sms * 2 : num_sms * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adamw_peer_kernel<<<blocks, 256, 0, stream>>>(ad, n_params, pr);
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

// ------------------------------------------------------------------------------------------------------------------
// Free-standing loss terms: Auxilleryfunc.__call__(y_pred, y_target) (linna/util.py:1070-1088) for tensors the caller
// already holds -- the three quadratic forms of the training loss in normalised space for rows of network output
// `y_pred` and physical targets `y_target`, plus d loss_b / d y_pred_b for autograd.  Eight rows per CTA: the three
// residual vectors of every row sit in shared memory, a thread owns output columns j = tid, tid + 256, ... and walks
// the (symmetrised) C_hat^-1 column-wise with coalesced loads.  Not a hot path (Predictor.train never calls it: the
// training kernels evaluate the loss in their own epilogues) -- it exists so that Loss_fn / Val_metric_fn work on
// free-standing tensors exactly like the reference's.
namespace linna {

constexpr int LT_ROWS = 8;

__global__ void __launch_bounds__(256) loss_terms_kernel(const float *__restrict__ y_pred, const float *__restrict__ y_target, int64_t n,
                                                         int n_out, const float *__restrict__ data_hat, const float *__restrict__ icov,
                                                         const float *__restrict__ sigma, const float *__restrict__ y_mean,
                                                         const float *__restrict__ y_std, int ypositive, float *__restrict__ loss,
                                                         float *__restrict__ chisq_md, float *__restrict__ chisq_nnd,
                                                         float *__restrict__ dloss)
{
    extern __shared__ float lt_smem[];                 // [3][LT_ROWS][n_out] residuals, then [3][LT_ROWS] reductions
    float *delta = lt_smem;
    double *red = reinterpret_cast<double *>(lt_smem + 3 * LT_ROWS * n_out);     // 96 n_out bytes in: 8-byte aligned
    __shared__ double chi[3][LT_ROWS];
    const int64_t row0 = (int64_t)blockIdx.x * LT_ROWS;
    const int tid = threadIdx.x;
    for (int e = tid; e < LT_ROWS * n_out; e += 256) {
        const int r = e / n_out, j = e - r * n_out;
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
        if (row0 + r < n) {
            const float Y = y_target[(row0 + r) * n_out + j], yp = y_pred[(row0 + r) * n_out + j], dh = data_hat[j];
            float t = Y / (sigma ? sigma[j] : 1.f);                                  // util.py:432
            if (ypositive) t = logf(t);                                              // util.py:567-568
            t = (t - y_mean[j]) / y_std[j];                                          // util.py:570
            if (!(Y == 1e-30f || Y == 1e10f || dh == 1e-30f)) d0 = t - yp, d1 = t - dh, d2 = yp - dh;   // util.py:1072
        }
        delta[(0 * LT_ROWS + r) * n_out + j] = d0, delta[(1 * LT_ROWS + r) * n_out + j] = d1, delta[(2 * LT_ROWS + r) * n_out + j] = d2;
    }
    if (tid < 3 * LT_ROWS) chi[tid / LT_ROWS][tid % LT_ROWS] = 0.0;
    __syncthreads();
    double part[3][LT_ROWS];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int r = 0; r < LT_ROWS; ++r) part[k][r] = 0.0;
    for (int j = tid; j < n_out; j += 256) {
        float q[3][LT_ROWS];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int r = 0; r < LT_ROWS; ++r) q[k][r] = 0.f;
        for (int i = 0; i < n_out; ++i) {
            const float c = icov[(size_t)i * n_out + j];
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int r = 0; r < LT_ROWS; ++r) q[k][r] = fmaf(delta[(k * LT_ROWS + r) * n_out + i], c, q[k][r]);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int r = 0; r < LT_ROWS; ++r) part[k][r] += (double)(q[k][r] * delta[(k * LT_ROWS + r) * n_out + j]);
        if (dloss)   // raw q of the target-vs-prediction form; scaled by -2 / chi2_Md below, once that is known
#pragma unroll
            for (int r = 0; r < LT_ROWS; ++r)
                if (row0 + r < n) dloss[(row0 + r) * n_out + j] = q[0][r];
    }
    // block reduction of the 24 partial sums
    for (int k = 0; k < 3; ++k)
        for (int r = 0; r < LT_ROWS; ++r) {
            double v = part[k][r];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((tid & 31) == 0) red[tid >> 5] = v;
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int w = 0; w < 8; ++w) s += red[w];
                chi[k][r] = s;
            }
            __syncthreads();
        }
    if (tid < LT_ROWS && row0 + tid < n) {
        const float mnn = (float)chi[0][tid], nnd = (float)chi[2][tid];
        float md = (float)chi[1][tid];
        if (md < 0.5f * (float)n_out) md = 0.5f * (float)n_out;                      // util.py:1086
        loss[row0 + tid] = mnn / md, chisq_md[row0 + tid] = md, chisq_nnd[row0 + tid] = nnd;   // util.py:1087
        chi[1][tid] = (double)md;
    }
    __syncthreads();
    if (dloss) {
        for (int e = tid; e < LT_ROWS * n_out; e += 256) {
            const int r = e / n_out, j = e - r * n_out;
            if (row0 + r >= n) continue;
            // d (delta^T A delta / md) / d y_pred = -2 A delta / md where the entry is not masked (delta = t - y_pred)
            const float Y = y_target[(row0 + r) * n_out + j];
            const bool masked = Y == 1e-30f || Y == 1e10f || data_hat[j] == 1e-30f;
            const float g = dloss[(row0 + r) * n_out + j] * (float)(-2.0 / chi[1][r]);
            dloss[(row0 + r) * n_out + j] = masked ? 0.f : g;
        }
    }
}

}  // namespace linna

extern "C" int linna_loss_terms(const float *y_pred, const float *y_target, int64_t n, int32_t n_out, const float *data_hat,
                                const float *icov_hat, const float *sigma, const float *y_mean, const float *y_std, int32_t ypositive,
                                float *loss, float *chisq_md, float *chisq_nnd, float *dloss, void *stream)
{
    if (!y_pred || !y_target || !data_hat || !icov_hat || !y_mean || !y_std || !loss || !chisq_md || !chisq_nnd || n < 0 || n_out <= 0)
        return -1;
    if (n == 0) return 0;
    const size_t smem = ((size_t)3 * linna::LT_ROWS * n_out + 2) * sizeof(float) + 8 * sizeof(double);
    if (smem > 200 * 1024) return -1;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(linna::loss_terms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -3;
        attr = true;
    }
    const int grid = (int)((n + linna::LT_ROWS - 1) / linna::LT_ROWS);
    linna::loss_terms_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(y_pred, y_target, n, n_out, data_hat, icov_hat, sigma, y_mean, y_std,
                                                                        ypositive, loss, chisq_md, chisq_nnd, dloss);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

// ------------------------------------------------------------------------------------------ training-set statistics
// Per-column LOWER median (what torch.median returns) of f(Y[r][c]) over the rows of a row-major [n][d] matrix, by radix
// selection on order-preserving 32-bit keys: four passes of a 256-bin histogram, one CTA per column, no sort and no
// shared-memory capacity limit.  f(v) = v / sigma_c, optionally log(.), optionally |. - centre_c|: the reference's
// normalisation statistics y_mean = median(y / sigma) and y_std = median |y / sigma - y_mean| (linna/util.py:1440-1450,
// :1308-1313), which it takes with two CPU sorts of the whole training set.  The matrix (21 MB at 10^4 x 500) is read
// from L2.
namespace linna {
__device__ __forceinline__ uint32_t stat_key(float v)
{
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float stat_unkey(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__global__ void __launch_bounds__(256) column_select_kernel(const float *__restrict__ Y, int64_t n, int d, const float *__restrict__ sigma,
                                                           int take_log, const float *__restrict__ centre, float *__restrict__ out)
{
    __shared__ unsigned hist[256];
    __shared__ uint32_t s_prefix, s_mask;
    __shared__ long long s_k;
    const int c = blockIdx.x;
    const float sg = sigma ? sigma[c] : 1.f, ctr = centre ? centre[c] : 0.f;
    if (threadIdx.x == 0) s_prefix = 0u, s_mask = 0u, s_k = (n - 1) / 2;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[threadIdx.x] = 0u;
        __syncthreads();
        const uint32_t prefix = s_prefix, mask = s_mask;
        for (int64_t r = threadIdx.x; r < n; r += 256) {
            float v = Y[r * d + c] / sg;                 // Y_transform_data, linna/util.py:432
            if (take_log) v = logf(v);
            if (centre) v = fabsf(v - ctr);
            const uint32_t key = stat_key(v);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long k = s_k, cum = 0;
            int b = 0;
            for (; b < 255; ++b) {
                if (cum + (long long)hist[b] > k) break;
                cum += hist[b];
            }
            s_k = k - cum;
            s_prefix = prefix | ((uint32_t)b << shift);
            s_mask = mask | (255u << shift);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = stat_unkey(s_prefix);
}
}  // namespace linna

extern "C" int linna_column_median_mad(const float *Y, int64_t n, int32_t d, const float *sigma, int32_t take_log, float *median,
                                       float *mad, void *stream)
{
    if (!Y || !median || n <= 0 || d <= 0) return LINNA_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    linna::column_select_kernel<<<d, 256, 0, st>>>(Y, n, d, sigma, take_log, nullptr, median);
    if (mad) linna::column_select_kernel<<<d, 256, 0, st>>>(Y, n, d, sigma, take_log, median, mad);
    return cudaGetLastError() == cudaSuccess ? LINNA_OK : LINNA_ECUDA;
}
