// Fused FP32 (FFMA) emulator-likelihood kernel for sm_100a.
//
// One persistent CTA owns a tile of BM = 8*RG walkers and runs the whole step program on it:
//   prologue : latent u -> prior CDF map -> log10 -> input normalisation          (a2, a3 of SURVEY 8a)
//   steps    : every layer / res-block of the MLP as a register-tiled GEMM         (a5)
//              head epilogue = inverse output transform + residual d = m - data    (a6, a7)
//              chi^2 step    = Cholesky-factor product r = L^T d, shuffle-reduced  (a8)
//              backward steps reuse the relu masks saved by the forward            (a11)
//   finish   : lnP = -chi^2/(2T) - |u|^2/2, NaN -> -inf                            (a1, a9, a10)
//
// GEMM mapping (per step, per 8*CG-column chunk):  256 threads = RG row-groups x KS k-slices x CG
// column-groups; every thread owns an 8x8 register tile (rows rg*8.., columns {4cg..4cg+3} and
// {4CG+4cg..}), so the inner loop is 4 LDS.128 per 64 FFMA for every layer width: narrow layers
// trade column groups for k-slices (split-K inside the CTA, reduced through shared memory).
// Weights ([k][n] packed) and activations ([feature][row] in the per-CTA L2-resident arena) are
// streamed through a cp.async ring of 24 KB stages.  A tile of 8*RG walkers uses every weight 8*RG times, so the
// CTA needs 64/RG bytes per clock from L2 at the full FFMA rate, and the bytes in flight decide what it gets
// (an L2 round trip is ~2000 cycles): the 8-row kernel (small batches, training; one CTA per SM) runs an
// 8-stage ring (112 KB in flight), the 16- and 32-row kernels 4 stages with two CTAs per SM.
#include "linna_device.cuh"
#include "../../include/linna_b200.h"

namespace linna {

__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gsrc, bool valid)
{
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

struct Geo {  // thread geometry of one GEMM step
    int cg_log2, CG, KS, ks_log2, NCH;
    int cg, ks, rg;
};

template <int RG>
struct Cfg {
    static constexpr int BM = 8 * RG, BK = 2 * RG;
    static constexpr int BK_LOG2 = (RG == 4) ? 3 : (RG == 2) ? 2 : 1;
    static constexpr int CGMAX_LOG2 = (RG == 4) ? 6 : (RG == 2) ? 7 : 8;  // 256/RG column groups at most
    static constexpr int CGMIN_LOG2 = CGMAX_LOG2 - 3;                     // at most 8 k-slices
    static constexpr int STAGES = (RG == 1) ? 8 : 4;                      // cp.async ring depth
};

template <int RG>
__device__ __forceinline__ Geo make_geo(int N, int tid)
{
    Geo g;
    int need = (N + 7) >> 3;  // column groups needed
    int l = Cfg<RG>::CGMIN_LOG2;
    while ((1 << l) < need && l < Cfg<RG>::CGMAX_LOG2) ++l;
    g.cg_log2 = l;
    g.CG = 1 << l;
    g.ks_log2 = Cfg<RG>::CGMAX_LOG2 - l;
    g.KS = 1 << g.ks_log2;
    g.NCH = 8 << l;
    g.cg = tid & (g.CG - 1);
    int t = tid >> l;
    g.ks = t & (g.KS - 1);
    g.rg = t >> g.ks_log2;
    return g;
}

// acc += A[K][BM]^T-tile @ Wt[K][ldw] columns [n0, n0+NCH).  The column-group count is a template
// parameter so that every shared-memory stride in the inner loop is an immediate.
template <int RG, int CGL>
__device__ __forceinline__ void gemm_phase(float (&acc)[8][8], float *smem, const float *__restrict__ A,
                                           const float *__restrict__ Wt, int K, int ldw, int n0, int tid)
{
    constexpr int BM = Cfg<RG>::BM, BK = Cfg<RG>::BK, BKL = Cfg<RG>::BK_LOG2, kStages = Cfg<RG>::STAGES;
    constexpr int CG = 1 << CGL, KSL = Cfg<RG>::CGMAX_LOG2 - CGL, KS = 1 << KSL, NCH = 8 * CG;
    constexpr int R4L = BKL;              // log2(BM/4) == log2(BK)
    constexpr int NA4 = KS * BK * (BM / 4);  // float4 copies of the activation slab
    const int kslice = (((K + KS - 1) >> KSL) + BK - 1) / BK * BK;
    const int nt = kslice / BK;

    // ---- per-thread copy descriptors, computed once
    const float *wp[4];
    int wk[4], wdst[4];
    bool wcol[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int f = tid + i * kThreads;
        const int c4 = f & (2 * CG - 1);
        const int t = f >> (CGL + 1);  // row of the [KS*BK][NCH] slab
        const int krow = (t >> BKL) * kslice + (t & (BK - 1));
        const int col = n0 + 4 * c4;
        wk[i] = krow;
        wcol[i] = col < ldw;
        wp[i] = Wt + (size_t)krow * ldw + (wcol[i] ? col : 0);
        wdst[i] = t * NCH + 4 * c4;
    }
    const float *ap[2];
    int ak[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int f = tid + i * kThreads;
        const int t = f >> R4L;
        const int krow = (t >> BKL) * kslice + (t & (BK - 1));
        ak[i] = krow;
        ap[i] = A + (size_t)krow * BM + 4 * (f & (BM / 4 - 1));
    }
    const size_t wstep = (size_t)BK * ldw;

    auto load_tile = [&](int kt, int stage) {
        float *sW = smem + stage * kStageFloats;
        float *sA = sW + kStageWFloats;
        const int k0 = kt * BK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool v = wcol[i] && (wk[i] + k0 < K);
            cp_async16(sW + wdst[i], v ? wp[i] : Wt, v);
            wp[i] += wstep;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (i * kThreads < NA4 && (NA4 >= (i + 1) * kThreads || tid < NA4 - i * kThreads)) {
                const bool v = ak[i] + k0 < K;
                cp_async16(sA + 4 * (tid + i * kThreads), v ? ap[i] : A, v);
                ap[i] += BK * BM;
            }
        }
    };

#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
        if (s < nt) load_tile(s, s);
        cp_async_commit();
    }
    const int cg = tid & (CG - 1), ks = (tid >> CGL) & (KS - 1), rg = tid >> (CGL + KSL);
    const int offW = (ks * BK) * NCH + 4 * cg;
    const int offA = kStageWFloats + (ks * BK) * BM + rg * 8;
    int stage = 0;
    for (int kt = 0; kt < nt; ++kt) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        {
            const int nx = kt + kStages - 1;
            int st2 = stage + kStages - 1;
            if (st2 >= kStages) st2 -= kStages;
            if (nx < nt) load_tile(nx, st2);
            cp_async_commit();
        }
        const float *sW = smem + stage * kStageFloats + offW;
        const float *sA = smem + stage * kStageFloats + offA;
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(sA + kk * BM);
            const float4 a1 = *reinterpret_cast<const float4 *>(sA + kk * BM + 4);
            const float4 b0 = *reinterpret_cast<const float4 *>(sW + kk * NCH);
            const float4 b1 = *reinterpret_cast<const float4 *>(sW + kk * NCH + 4 * CG);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (++stage == kStages) stage = 0;
    }
    cp_async_wait<0>();
    __syncthreads();
}

template <int RG>
__device__ __forceinline__ void gemm_dispatch(float (&acc)[8][8], float *smem, const float *A, const float *Wt, int K,
                                              int ldw, int n0, int cg_log2, int tid)
{
    constexpr int L0 = Cfg<RG>::CGMIN_LOG2;
    switch (cg_log2 - L0) {
    case 0: gemm_phase<RG, L0 + 0>(acc, smem, A, Wt, K, ldw, n0, tid); break;
    case 1: gemm_phase<RG, L0 + 1>(acc, smem, A, Wt, K, ldw, n0, tid); break;
    case 2: gemm_phase<RG, L0 + 2>(acc, smem, A, Wt, K, ldw, n0, tid); break;
    default: gemm_phase<RG, L0 + 3>(acc, smem, A, Wt, K, ldw, n0, tid); break;
    }
}

// Sum the KS partial tiles of one (rg, cg) into the ks == 0 thread.
__device__ __forceinline__ void reduce_ks(float (&acc)[8][8], float *smem, const Geo &g, int tid)
{
    if (g.KS == 1) return;
    float4 *red = reinterpret_cast<float4 *>(smem);  // [16][256] float4 = 64 KB <= ring size
    if (g.ks != 0) {
#pragma unroll
        for (int q = 0; q < 16; ++q)
            red[q * kThreads + tid] = make_float4(acc[q >> 1][(q & 1) * 4 + 0], acc[q >> 1][(q & 1) * 4 + 1],
                                                  acc[q >> 1][(q & 1) * 4 + 2], acc[q >> 1][(q & 1) * 4 + 3]);
    }
    __syncthreads();
    if (g.ks == 0) {
        for (int s = 1; s < g.KS; ++s) {
            int other = tid + s * g.CG;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                float4 v = red[q * kThreads + other];
                acc[q >> 1][(q & 1) * 4 + 0] += v.x;
                acc[q >> 1][(q & 1) * 4 + 1] += v.y;
                acc[q >> 1][(q & 1) * 4 + 2] += v.z;
                acc[q >> 1][(q & 1) * 4 + 3] += v.w;
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void store8(float *p, const float (&v)[8])
{
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4 *>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void load8_cg(const float *p, float (&v)[8])
{
    float4 a = __ldcg(reinterpret_cast<const float4 *>(p));
    float4 b = __ldcg(reinterpret_cast<const float4 *>(p + 4));
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}

// Row-major copy of an epilogue column (training: operands of the weight-gradient GEMM).
__device__ __forceinline__ void store_rm(float *rm_base, const Step &st, int64_t row0, int row_base, int nrows, int cidx,
                                         const float (&v)[8])
{
    float *p = rm_base + (size_t)st.rm_off + (size_t)(row0 + row_base) * st.rm_ld + cidx;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (row_base + i < nrows) p[(size_t)i * st.rm_ld] = v[i];
}

// Prior CDF map of one latent coordinate (Transform.__call__, linna/util.py:339-343).
__device__ __forceinline__ float prior_map(float u, int kind, float scale, float shift)
{
    float t = u;
    if (kind == LINNA_PRIOR_FLAT) t = 0.5f * (1.0f + erff(u / 1.41421356237309515f));  // gauss2unif, util.py:300
    return t * scale + shift;
}

template <int RG>
__global__ void __launch_bounds__(kThreads, RG == 1 ? 1 : 2) fused_ffma_kernel(const KernelArgs args)
{
    constexpr int BM = 8 * RG;
    extern __shared__ float4 smem4[];
    float *smem = reinterpret_cast<float *>(smem4);
    double *chi_part = reinterpret_cast<double *>(smem + Cfg<RG>::STAGES * kStageFloats);  // [RG*nsub<=8][8]
    double *chi_acc = chi_part + 64;                                               // [BM]
    float *lnprior = reinterpret_cast<float *>(chi_acc + 32);                      // [BM]

    __shared__ Step s_steps[kMaxSteps];  // the step program, staged once per CTA (broadcast LDS afterwards)

    const int tid = threadIdx.x;
    const Program *__restrict__ prog = args.prog;
    const Consts &c = args.c;
    float *arena = args.arena + (size_t)blockIdx.x * prog->arena_features * BM;
    uint8_t *masks = args.masks + (size_t)blockIdx.x * prog->mask_features * RG;
    const int n_in = c.n_in, n_out = c.n_out;
    // indexed launch: the row count lives in device memory (rows flagged by the tensor-core kernel)
    const int64_t n_rows = args.n_dev ? ((int64_t)__ldg(args.n_dev) < args.n ? (int64_t)__ldg(args.n_dev) : args.n) : args.n;
    const int32_t *__restrict__ ridx = args.row_index;
    if (args.zero_me && blockIdx.x == 0 && threadIdx.x == 0) *args.zero_me = 0;
    const int64_t ntiles = (n_rows + BM - 1) / BM;
    const int n_steps = prog->n_steps;
    {
        const int nwords = n_steps * (int)(sizeof(Step) / 4);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(prog->steps);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_steps);
        for (int i = tid; i < nwords; i += kThreads) dst[i] = src[i];
    }
    __syncthreads();

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * BM;
        const int nrows = (int)((n_rows - row0) < BM ? (n_rows - row0) : BM);
        // global row of tile row r (only dereferenced for r < nrows)
        auto grow = [&](int r) -> int64_t { return ridx ? (int64_t)ridx[row0 + r] : row0 + r; };

        // ---------------- prologue: u -> theta -> xhat (feature-major) ----------------
        {
            float *xb = arena + (size_t)prog->in_buf * BM;
            const float *in = args.in;
            for (int e = tid; e < BM * n_in; e += kThreads) {
                int r = e / n_in, i = e - r * n_in;
                float th = 0.f;
                if (r < nrows) {
                    float u = in[grow(r) * n_in + i];
                    th = args.input_theta ? u : prior_map(u, c.prior_kind[i], c.prior_scale[i], c.prior_shift[i]);
                    if (c.log10_flag && c.log10_flag[i]) th = log10f(th);            // util.py:491-496
                    th = (th - c.x_mean[i]) / c.x_std[i];                             // util.py:497
                }
                xb[(size_t)i * BM + r] = th;
                if (args.rm_base && prog->in_rm_off >= 0 && r < nrows)
                    args.rm_base[(size_t)prog->in_rm_off + (size_t)(row0 + args.rm_row0 + r) * prog->in_rm_ld + i] = th;
            }
            if (tid < BM) {
                float s = 0.f;
                if (tid < nrows && !args.input_theta) {
                    const float *ur = in + grow(tid) * n_in;
                    for (int i = 0; i < n_in; ++i) { float u = ur[i]; s = fmaf(u, u, s); }
                }
                lnprior[tid] = -0.5f * s;                                             // util.py:1165
                chi_acc[tid] = 0.0;
            }
        }
        __syncthreads();

        // ---------------- the step program ----------------
        for (int si = 0; si < n_steps; ++si) {
            const Step &st = s_steps[si];
            const int N = st.N;
            const Geo g = make_geo<RG>(N, tid);
            const int row_base = g.rg * 8;
            for (int n0 = 0; n0 < N; n0 += g.NCH) {
                float acc[8][8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
                if (st.K1 > 0)
                    gemm_dispatch<RG>(acc, smem, arena + (size_t)st.src1 * BM, st.wt1, st.K1, st.ldw1, n0, g.cg_log2, tid);
                if (st.scale != 1.0f) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] *= st.scale;
                }
                if (st.K2 > 0)
                    gemm_dispatch<RG>(acc, smem, arena + (size_t)st.src2 * BM, st.wt2, st.K2, st.ldw2, n0, g.cg_log2, tid);
                reduce_ks(acc, smem, g, tid);

                double part[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) part[i] = 0.0;

                if (g.ks == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int cidx = n0 + ((j < 4) ? (4 * g.cg + j) : (4 * g.CG + 4 * g.cg + (j - 4)));
                        if (cidx >= N) continue;
                        float v[8];
                        const float b = st.bias ? st.scale * __ldg(st.bias + cidx) : 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = acc[i][j] + b;
                        if (st.flags & F_ADD_SRC2) {
                            float s2[8];
                            load8_cg(arena + (size_t)(st.src2 + cidx) * BM + row_base, s2);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] += s2[i];
                        }
                        if (st.epi == EPI_ACT || st.epi == EPI_HEAD) {
                            if (st.flags & F_RELU) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
                            }
                            if (st.flags & F_SAVE_MASK) {
                                unsigned m = 0;
#pragma unroll
                                for (int i = 0; i < 8; ++i) m |= (v[i] > 0.f ? 1u : 0u) << i;
                                masks[(size_t)(st.mask_off + cidx) * RG + g.rg] = (uint8_t)m;
                            }
                        }
                        if (st.epi == EPI_ACT) {
                            store8(arena + (size_t)(st.dst + cidx) * BM + row_base, v);
                            if (st.rm_off >= 0) store_rm(args.rm_base, st, row0 + args.rm_row0, row_base, nrows, cidx, v);
                        } else if (st.epi == EPI_LOSSHEAD) {
                            const float ys = __ldg(c.y_std + cidx), ym = __ldg(c.y_mean + cidx);
                            const float sg = c.sigma ? __ldg(c.sigma + cidx) : 1.f;
                            const float dh = __ldg(c.data_hat + cidx);
                            float dl[8];
                            unsigned mb = 0;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int r = row_base + i;
                                float dv = 0.f;
                                if (r < nrows) {
                                    const float Y = __ldg(args.target + (row0 + r) * n_out + cidx);
                                    float t = Y / sg;                                              // util.py:432
                                    if (c.ypositive) t = logf(t);                                  // util.py:567-568
                                    t = (t - ym) / ys;                                             // util.py:570
                                    const bool ok = !(Y == 1e-30f || Y == 1e10f || dh == 1e-30f);   // util.py:1072
                                    dv = args.delta_kind == 0 ? t - v[i] : args.delta_kind == 1 ? t - dh : v[i] - dh;
                                    if (!ok) dv = 0.f;
                                    mb |= (ok ? 1u : 0u) << i;
                                }
                                dl[i] = dv;
                            }
                            store8(arena + (size_t)(st.dst + cidx) * BM + row_base, dl);
                            if (st.flags & F_SAVE_MASK) masks[(size_t)(st.mask_off + cidx) * RG + g.rg] = (uint8_t)mb;
                        } else if (st.epi == EPI_LOSSQ) {
                            float d[8];
                            load8_cg(arena + (size_t)(st.src1 + cidx) * BM + row_base, d);
#pragma unroll
                            for (int i = 0; i < 8; ++i) part[i] += (double)(v[i] * d[i]);
                            if (st.flags & F_LOSS_GRAD) {
                                const unsigned m = __ldcg(masks + (size_t)(st.mask_off + cidx) * RG + g.rg);
                                float gq[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const int r = row_base + i;
                                    const float rs = r < nrows ? -2.0f * args.loss_inv_B / __ldg(args.cmd + row0 + r) : 0.f;
                                    gq[i] = ((m >> i) & 1u) ? v[i] * rs : 0.f;
                                }
                                store8(arena + (size_t)(st.dst + cidx) * BM + row_base, gq);
                                if (st.rm_off >= 0) store_rm(args.rm_base, st, row0 + args.rm_row0, row_base, nrows, cidx, gq);
                            }
                        } else if (st.epi == EPI_HEAD) {
                            const float ys = __ldg(c.y_std + cidx), ym = __ldg(c.y_mean + cidx);
                            const float sg = c.sigma ? __ldg(c.sigma + cidx) : 1.f;
                            const float dt = c.data ? __ldg(c.data + cidx) : 0.f;
                            float y[8], d[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float yy = fmaf(v[i], ys, ym);                        // util.py:542
                                if (c.ypositive) yy = expf(yy);                      // util.py:540
                                y[i] = yy;
                                d[i] = yy * sg - dt;                                  // util.py:458, :954
                            }
                            if (st.flags & F_COT) {
                                // vector-Jacobian product: the cotangent of the requested output (yhat | y | m) pulled back
                                // to yhat: d y / d yhat = y_std (x y with the exp), d m / d y = sigma
                                float ct[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const int r = row_base + i;
                                    float g = r < nrows ? __ldg(args.target + (row0 + r) * n_out + cidx) : 0.f;
                                    if ((st.flags & F_RELU) && !(v[i] > 0.f)) g = 0.f;
                                    if (args.out_kind != LINNA_OUT_YHAT) g *= c.ypositive ? ys * y[i] : ys;
                                    if (args.out_kind == LINNA_OUT_M) g *= sg;
                                    ct[i] = g;
                                }
                                store8(arena + (size_t)(st.dst + cidx) * BM + row_base, ct);
                                if (st.rm_off >= 0) store_rm(args.rm_base, st, row0 + args.rm_row0, row_base, nrows, cidx, ct);
                            } else
                                store8(arena + (size_t)(st.dst + cidx) * BM + row_base, d);
                            if (st.flags & F_SAVE_Y) store8(arena + (size_t)(st.ybuf + cidx) * BM + row_base, y);
                            if ((st.flags & F_OUT_VEC) && args.out_vec) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    int r = row_base + i;
                                    if (r < nrows) {
                                        float o = args.out_kind == LINNA_OUT_YHAT ? v[i]
                                                  : args.out_kind == LINNA_OUT_Y  ? y[i]
                                                                                  : y[i] * sg;
                                        args.out_vec[grow(r) * n_out + cidx] = o;
                                    }
                                }
                            }
                        } else if (st.epi == EPI_CHI2) {
                            if (c.quad_kind == LINNA_QUAD_CHOL) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) part[i] += (double)(v[i] * v[i]);
                            } else {
                                float d[8];
                                load8_cg(arena + (size_t)(st.src1 + cidx) * BM + row_base, d);
#pragma unroll
                                for (int i = 0; i < 8; ++i) part[i] += (double)(v[i] * d[i]);
                            }
                            if (st.flags & F_STORE_DST) store8(arena + (size_t)(st.dst + cidx) * BM + row_base, v);
                        } else if (st.epi == EPI_BWD) {
                            if (st.colscale) {
                                const float cs = __ldg(st.colscale + cidx);
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] *= cs;
                            }
                            if (st.flags & F_MUL_YSAVE) {
                                float y[8];
                                load8_cg(arena + (size_t)(st.ybuf + cidx) * BM + row_base, y);
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] *= y[i];
                            }
                            if (st.flags & F_APPLY_MASK) {
                                const unsigned m = __ldcg(masks + (size_t)(st.mask_off + cidx) * RG + g.rg);
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = ((m >> i) & 1u) ? v[i] : 0.f;
                            }
                            store8(arena + (size_t)(st.dst + cidx) * BM + row_base, v);
                            if (st.rm_off >= 0) store_rm(args.rm_base, st, row0 + args.rm_row0, row_base, nrows, cidx, v);
                        } else if (st.epi == EPI_GRAD) {
                            // chain through xhat = (theta' - mean)/std, theta' = log10(theta), theta = prior(u)
                            const int kind = args.input_theta ? 0 : c.prior_kind[cidx];
                            const float ps = args.input_theta ? 1.f : c.prior_scale[cidx], psh = args.input_theta ? 0.f : c.prior_shift[cidx];
                            const float inv_std = 1.0f / c.x_std[cidx];
                            const bool lg = c.log10_flag && c.log10_flag[cidx];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                int r = row_base + i;
                                if (r < nrows) {
                                    const int64_t gr = grow(r);
                                    float u = args.in[gr * n_in + cidx];
                                    float gx = v[i] * inv_std;
                                    if (args.input_theta) {   // input is the physical parameter itself: no prior map, no prior term
                                        if (lg) gx /= (u * 2.30258509299404568f);
                                        args.grad[gr * n_in + cidx] = gx;
                                        continue;
                                    }
                                    if (lg) gx /= (prior_map(u, kind, ps, psh) * 2.30258509299404568f);
                                    float jac = ps;
                                    if (kind == LINNA_PRIOR_FLAT) jac *= 0.398942280401432678f * expf(-0.5f * u * u);
                                    args.grad[gr * n_in + cidx] = gx * jac - u;
                                }
                            }
                        }
                    }
                }

                if (st.epi == EPI_CHI2 || st.epi == EPI_LOSSQ) {
                    // shuffle-reduce the row partials over the column groups of one (rg, ks) lane group
                    const int width = g.CG < 32 ? g.CG : 32;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        double p = part[i];
                        for (int o = width >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
                        part[i] = p;
                    }
                    const int nsub = g.CG >> 5 ? g.CG >> 5 : 1;
                    if (g.ks == 0 && (g.cg & (width - 1)) == 0) {
                        const int sub = g.cg >> 5;
#pragma unroll
                        for (int i = 0; i < 8; ++i) chi_part[(g.rg * nsub + sub) * 8 + i] = part[i];
                    }
                    __syncthreads();
                    if (tid < BM) {
                        double s = chi_acc[tid];
                        for (int sub = 0; sub < nsub; ++sub) s += chi_part[((tid >> 3) * nsub + sub) * 8 + (tid & 7)];
                        chi_acc[tid] = s;
                    }
                }
                __syncthreads();  // dst visible to the next step; ring + chi_part reusable
            }
            if (st.epi == EPI_LOSSQ && tid < nrows && args.lnp) {
                const double chi = chi_acc[tid];
                args.lnp[grow(tid)] = args.cmd ? (float)chi / __ldg(args.cmd + row0 + tid) : (float)chi;   // util.py:1087
            }
            if (st.epi == EPI_CHI2 && tid < nrows && args.lnp) {
                float l = (float)(-0.5 * chi_acc[tid]) * c.inv_T + lnprior[tid];       // util.py:1013
                if (l != l) l = -INFINITY;                                            // util.py:1015-1016
                args.lnp[grow(tid)] = l;
            }
        }
        __syncthreads();
    }
}

static size_t fused_smem_bytes(int rg)
{
    const int stages = rg == 1 ? Cfg<1>::STAGES : rg == 2 ? Cfg<2>::STAGES : Cfg<4>::STAGES;
    return (size_t)stages * kStageFloats * sizeof(float) + 64 * 8 + 32 * 8 + 32 * 4;
}

cudaError_t launch_fused_ffma(const KernelArgs &args, int rg, int grid, cudaStream_t stream)
{
    static bool attr_set[3] = {false, false, false};
    const size_t smem = fused_smem_bytes(rg);
    cudaError_t e = cudaSuccess;
#define LINNA_LAUNCH(RGV, IDX)                                                                                  \
    {                                                                                                           \
        if (!attr_set[IDX]) {                                                                                   \
            e = cudaFuncSetAttribute(fused_ffma_kernel<RGV>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                     (int)smem);                                                                \
            if (e != cudaSuccess) return e;                                                                     \
            attr_set[IDX] = true;                                                                               \
        }                                                                                                       \
        fused_ffma_kernel<RGV><<<grid, kThreads, smem, stream>>>(args);                                         \
    }
    if (rg == 4) LINNA_LAUNCH(4, 0)
    else if (rg == 2) LINNA_LAUNCH(2, 1)
    else LINNA_LAUNCH(1, 2)
#undef LINNA_LAUNCH
    return cudaGetLastError();
}

int fused_ffma_max_ctas_per_sm(int rg)
{
    int nb = 0;
    const size_t smem = fused_smem_bytes(rg);
    if (rg == 4) {
        cudaFuncSetAttribute(fused_ffma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fused_ffma_kernel<4>, kThreads, smem);
    } else if (rg == 2) {
        cudaFuncSetAttribute(fused_ffma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fused_ffma_kernel<2>, kThreads, smem);
    } else {
        cudaFuncSetAttribute(fused_ffma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fused_ffma_kernel<1>, kThreads, smem);
    }
    return nb > 0 ? nb : 1;
}

}  // namespace linna
