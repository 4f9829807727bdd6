// Tensor-core emulator-likelihood kernel for large walker batches (sm_100a: tcgen05 + TMEM + TMA).
//
// Same step program as the FFMA kernel, but every GEMM  D[128 walkers][N] = A[128][K] . B[N][K]^T  runs
// on the 5th-generation tensor cores with the accumulator in tensor memory:
//   * 3xTF32 error-compensated product  D += A_lo.B_hi + A_hi.B_lo + A_hi.B_hi  (hi = rna_tf32(x),
//     lo = x - hi), fp32 accumulation in TMEM.  Single-pass TF32 (~5e-4) cannot meet the 1e-5 parity
//     bar; the split product carries ~2^-21 per term.
//   * operands are K-major 128-byte-swizzled tiles in shared memory, filled by TMA
//     (cp.async.bulk.tensor) from the packed hi/lo weights and from the row-major hi/lo activation
//     arena of this CTA; completion is tracked with mbarriers (expect_tx).
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread) + TMEM allocator,
//     warps 2-5 = epilogue: tcgen05.ld the 128x256 accumulator, bias/relu/inverse transform, split into
//     hi/lo and store the next layer's A operand; the chi^2 step reduces r^2 per TMEM lane (= walker)
//     without any cross-thread traffic.  Two 256-column accumulator buffers let the epilogue of one
//     column chunk overlap the MMAs of the next.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "linna_host.hpp"

namespace linna {

constexpr int TC_M = 128;       // walkers per tile
constexpr int TC_KC = 32;       // k-chunk in floats = one 128-byte swizzle row
constexpr int TC_NC = 256;      // accumulator columns per buffer
constexpr int TC_STAGES = 2;
constexpr int TC_A_BYTES = TC_M * TC_KC * 4;                      // 16 KB
constexpr int TC_B_BYTES = TC_NC * TC_KC * 4;                     // 32 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // 96 KB
constexpr int TC_THREADS = 320;       // TMA warp + MMA warp + 8 epilogue warps
constexpr int TC_EPI_THREADS = 256;
constexpr int TC_MAX_STEPS = 24;
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024;  // + alignment slack

enum TcEpi : int32_t { TC_EPI_ACT = 0, TC_EPI_HEAD = 1, TC_EPI_CHI2 = 2 };

struct TcStep {
    int32_t nphase;
    int32_t srcA[2], srcLo[2];     // arena column of the hi copy; lo copy at srcA + srcLo
    int32_t K[2];
    int32_t mapBhi[2], mapBlo[2];  // tensor-map indices of the weight operand (hi / lo)
    int32_t N;
    int32_t dst, dstLo, dstPad;    // output arena column (hi), lo offset, zero-padded width
    int32_t epi, relu, tri;
    float bias_scale;
    const float *bias;
};

struct TcProgram {
    int32_t n_steps, in_dst, in_lo, in_pad;
    int32_t ld, seg_kc, pad_[2];
    TcStep steps[TC_MAX_STEPS];
};

struct TcArgs {
    const TcProgram *prog;
    const CUtensorMap *maps;  // maps[0] = activation arena, then 2 per weight operand
    Consts c;
    const float *in;
    float *lnp;
    float *arena;
    int64_t n;
    int *err;
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap (and report) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int *err, int code)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 3000000000LL) {
            if (err) atomicExch(err, code);
            __threadfence_system();
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile, 128-byte swizzle: rows are 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);  // start address, 16-byte units
    d |= (uint64_t)1 << 16;                  // leading byte offset (unused with swizzle)
    d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc_tf32(int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float tf32_hi(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float tc_prior_map(float u, int kind, float scale, float shift)
{
    float t = u;
    if (kind == LINNA_PRIOR_FLAT) t = 0.5f * (1.0f + erff(u / 1.41421356237309515f));
    return t * scale + shift;
}

// ------------------------------------------------------------------------------------------ kernel
// Two-level accumulation.  The tensor core truncates the fp32 accumulator on every tcgen05.mma, so a
// chain of 3*K/8 instructions carries a systematic toward-zero bias of ~K * 2^-25 (measured: 6e-6 on
// chi^2 at K = 500), above the 1e-5 parity bar once a few layers stack.  Every `seg_kc` k-chunks
// (default 4 = 128 values of K = 48 instructions) the partial tile is therefore drained from tensor
// memory and added into fp32 REGISTER accumulators with round-to-nearest by the epilogue warps, while
// the MMA warp already fills the other TMEM buffer.
__global__ void __launch_bounds__(TC_THREADS, 1) tc_lnp_kernel(const TcArgs args)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], tfull_bar[2], tempty_bar[2], step_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ double chi_s[TC_M];
    __shared__ TcStep s_steps[TC_MAX_STEPS];

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const TcProgram *prog = args.prog;
    const int n_steps = prog->n_steps;
    const int ld = prog->ld;
    const int seg_kc = prog->seg_kc;
    const Consts &c = args.c;
    const CUtensorMap *maps = args.maps;

    for (int i = tid; i < n_steps * (int)(sizeof(TcStep) / 4); i += TC_THREADS)
        reinterpret_cast<uint32_t *>(s_steps)[i] = reinterpret_cast<const uint32_t *>(prog->steps)[i];
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
        for (int b = 0; b < 2; ++b) mbar_init(&tfull_bar[b], 1), mbar_init(&tempty_bar[b], TC_EPI_THREADS);
        mbar_init(&step_bar, TC_EPI_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    const int64_t ntiles = (args.n + TC_M - 1) / TC_M;
    const int arena_row0 = blockIdx.x * TC_M;   // this CTA's rows of the activation arena

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0, ev = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int si = 0; si < n_steps; ++si) {
                    const TcStep &st = s_steps[si];
                    bool input_ready = false;
                    for (int n0 = 0; n0 < st.N; n0 += TC_NC) {
                        for (int p = 0; p < st.nphase; ++p) {
                            const int nk = (st.K[p] + TC_KC - 1) / TC_KC;
                            const int k_first = st.tri ? n0 / TC_KC : 0;   // L^T: B[n][k] = 0 for k < n
                            for (int kc = k_first; kc < nk; ++kc) {
                                mbar_wait(&empty_bar[stage], ph ^ 1, args.err, 1);
                                uint8_t *sA = smem + stage * TC_STAGE_BYTES;
                                mbar_expect_tx(&full_bar[stage], TC_STAGE_BYTES);
                                tma_load_2d(sA + 2 * TC_A_BYTES, maps + st.mapBhi[p], &full_bar[stage], kc * TC_KC, n0);
                                tma_load_2d(sA + 2 * TC_A_BYTES + TC_B_BYTES, maps + st.mapBlo[p], &full_bar[stage], kc * TC_KC, n0);
                                if (!input_ready) {   // the activations this step reads were written by the previous one
                                    mbar_wait(&step_bar, ev & 1, args.err, 2);
                                    ++ev;
                                    input_ready = true;
                                }
                                tma_load_2d(sA, maps, &full_bar[stage], st.srcA[p] + kc * TC_KC, arena_row0);
                                tma_load_2d(sA + TC_A_BYTES, maps, &full_bar[stage], st.srcA[p] + st.srcLo[p] + kc * TC_KC, arena_row0);
                                if (++stage == TC_STAGES) stage = 0, ph ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 0, acc = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int si = 0; si < n_steps; ++si) {
                    const TcStep &st = s_steps[si];
                    for (int n0 = 0; n0 < st.N; n0 += TC_NC) {
                        const int nvalid = st.N - n0 < TC_NC ? st.N - n0 : TC_NC;
                        const uint32_t idesc = make_idesc_tf32((nvalid + 15) & ~15);
                        int in_seg = 0;          // k-chunks issued into the current TMEM buffer
                        uint32_t dcol = 0;
                        for (int p = 0; p < st.nphase; ++p) {
                            const int nk = (st.K[p] + TC_KC - 1) / TC_KC;
                            const int k_first = st.tri ? n0 / TC_KC : 0;
                            for (int kc = k_first; kc < nk; ++kc) {
                                if (in_seg == 0) {   // open a fresh accumulator buffer
                                    const int buf = acc & 1;
                                    mbar_wait(&tempty_bar[buf], ((acc >> 1) & 1) ^ 1, args.err, 3);
                                    tc_fence_after();
                                    dcol = tmem_base + buf * TC_NC;
                                }
                                mbar_wait(&full_bar[stage], ph, args.err, 4);
                                tc_fence_after();
                                const uint32_t a_hi = smem_u32(smem + stage * TC_STAGE_BYTES);
                                const uint32_t a_lo = a_hi + TC_A_BYTES, b_hi = a_hi + 2 * TC_A_BYTES, b_lo = b_hi + TC_B_BYTES;
                                const int kleft = st.K[p] - kc * TC_KC;
                                const int ksteps = kleft >= TC_KC ? 4 : (kleft + 7) >> 3;
                                for (int ks = 0; ks < ksteps; ++ks) {
                                    const uint32_t o = ks * 32;   // 8 tf32 = 32 bytes along K inside the swizzle row
                                    umma_tf32(dcol, make_sdesc(a_lo + o), make_sdesc(b_hi + o), idesc, (in_seg | ks) ? 1u : 0u);
                                    umma_tf32(dcol, make_sdesc(a_hi + o), make_sdesc(b_lo + o), idesc, 1);
                                    umma_tf32(dcol, make_sdesc(a_hi + o), make_sdesc(b_hi + o), idesc, 1);
                                }
                                umma_commit(&empty_bar[stage]);   // frees the smem stage when these MMAs retire
                                if (++stage == TC_STAGES) stage = 0, ph ^= 1;
                                const bool last_kc = (p == st.nphase - 1) && (kc == nk - 1);
                                if (++in_seg == seg_kc || last_kc) {
                                    umma_commit(&tfull_bar[acc & 1]);   // partial tile complete -> epilogue drains it
                                    ++acc;
                                    in_seg = 0;
                                }
                            }
                        }
                    }
                }
            }
        }
    } else {
        // =============================== epilogue warps ===============================
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                // which 128 columns of the 256-column chunk
        const int row = q * 32 + lane;                   // TMEM lane == walker of the tile
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        float *arow = args.arena + (size_t)(arena_row0 + row) * ld;
        uint32_t acc = 0;
        const int n_in = c.n_in;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t grow = tile * TC_M + row;
            const bool valid = grow < args.n;
            // ---- prologue: u -> theta -> xhat, split, store as the first A operand
            float lnprior = 0.f;
            if (half == 0) {
                const float *u = args.in + grow * n_in;
                for (int i = 0; i < prog->in_pad; ++i) {
                    float xh = 0.f;
                    if (i < n_in && valid) {
                        const float uu = u[i];
                        lnprior = fmaf(uu, uu, lnprior);
                        float th = tc_prior_map(uu, c.prior_kind[i], c.prior_scale[i], c.prior_shift[i]);
                        if (c.log10_flag && c.log10_flag[i]) th = log10f(th);
                        xh = (th - c.x_mean[i]) / c.x_std[i];
                    }
                    const float hi = tf32_hi(xh);
                    arow[prog->in_dst + i] = hi;
                    arow[prog->in_dst + prog->in_lo + i] = xh - hi;
                }
                lnprior *= -0.5f;
            }
            asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy stores -> visible to TMA reads
            mbar_arrive(&step_bar);
            double chi = 0.0;
            for (int si = 0; si < n_steps; ++si) {
                const TcStep &st = s_steps[si];
                for (int n0 = 0; n0 < st.N; n0 += TC_NC) {
                    const int nvalid = st.N - n0 < TC_NC ? st.N - n0 : TC_NC;
                    const int c0 = half * 128;                       // first chunk column of this thread
                    const int nmine = nvalid - c0 < 0 ? 0 : (nvalid - c0 > 128 ? 128 : nvalid - c0);
                    float racc[128];
#pragma unroll
                    for (int j = 0; j < 128; ++j) racc[j] = 0.f;
                    // number of TMEM segments of this chunk (must mirror the MMA warp's loop)
                    int nkc = 0;
                    for (int p = 0; p < st.nphase; ++p) nkc += (st.K[p] + TC_KC - 1) / TC_KC - (st.tri ? n0 / TC_KC : 0);
                    const int nseg = (nkc + seg_kc - 1) / seg_kc;
                    for (int sg = 0; sg < nseg; ++sg) {
                        const int buf = acc & 1;
                        mbar_wait(&tfull_bar[buf], (acc >> 1) & 1, args.err, 5);
                        tc_fence_after();
#pragma unroll
                        for (int cb = 0; cb < 128; cb += 32) {
                            if (cb < nmine) {
                                uint32_t r[32];
                                tmem_ld32(tmem_base + lane_addr + buf * TC_NC + c0 + cb, r);
#pragma unroll
                                for (int j = 0; j < 32; ++j) racc[cb + j] += __uint_as_float(r[j]);
                            }
                        }
                        tc_fence_before();
                        mbar_arrive(&tempty_bar[buf]);
                        ++acc;
                    }
                    if (st.epi == TC_EPI_CHI2) {
                        float chi_f = 0.f;
#pragma unroll
                        for (int j = 0; j < 128; ++j)
                            if (j < nmine) chi_f = fmaf(racc[j], racc[j], chi_f);
                        chi += (double)chi_f;
                    } else {
                        const int npad = st.dstPad - n0 - c0;        // columns to write incl. zero padding
#pragma unroll
                        for (int cb = 0; cb < 128; cb += 32) {
                            if (cb < npad) {
                                float hi[32], lo[32];
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const int col = n0 + c0 + cb + j;
                                    float v = 0.f;
                                    if (col < st.N) {
                                        v = racc[cb + j];
                                        if (st.bias) v += st.bias_scale * __ldg(st.bias + col);
                                        if (st.relu) v = fmaxf(v, 0.f);
                                        if (st.epi == TC_EPI_HEAD) {
                                            float y = fmaf(v, __ldg(c.y_std + col), __ldg(c.y_mean + col));
                                            if (c.ypositive) y = expf(y);
                                            v = y * __ldg(c.sigma + col) - __ldg(c.data + col);
                                        }
                                    }
                                    hi[j] = tf32_hi(v);
                                    lo[j] = v - hi[j];
                                }
                                float4 *ph4 = reinterpret_cast<float4 *>(arow + st.dst + n0 + c0 + cb);
                                float4 *pl4 = reinterpret_cast<float4 *>(arow + st.dst + st.dstLo + n0 + c0 + cb);
#pragma unroll
                                for (int qq = 0; qq < 8; ++qq) {
                                    ph4[qq] = make_float4(hi[4 * qq], hi[4 * qq + 1], hi[4 * qq + 2], hi[4 * qq + 3]);
                                    pl4[qq] = make_float4(lo[4 * qq], lo[4 * qq + 1], lo[4 * qq + 2], lo[4 * qq + 3]);
                                }
                            }
                        }
                    }
                }
                if (st.epi != TC_EPI_CHI2) {
                    asm volatile("fence.proxy.async;" ::: "memory");
                    mbar_arrive(&step_bar);
                }
            }
            // combine the two column halves of every walker and finish lnP
            if (half == 1) chi_s[row] = chi;
            asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
            if (half == 0 && valid) {
                float l = (float)(-0.5 * (chi + chi_s[row])) * c.inv_T + lnprior;
                if (l != l) l = -INFINITY;
                args.lnp[grow] = l;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcContext {
    float *blob = nullptr;
    float *arena = nullptr;
    CUtensorMap *maps_dev = nullptr;
    TcProgram *prog_dev = nullptr;
    int *err_dev = nullptr;
    int grid = 0, ld = 0;
    std::string error;
};

static inline int pad32(int n) { return (n + 31) & ~31; }

// k-chunks (of 32) accumulated in tensor memory between two promotions to the register accumulators;
// LINNA_TC_SEG_KC overrides it (experiments: a huge value reproduces plain single-level accumulation).
static int tc_seg_kc()
{
    const char *e = getenv("LINNA_TC_SEG_KC");
    int v = e ? atoi(e) : 4;
    return v > 0 ? v : 4;
}

static float host_tf32_hi(float x)
{
    uint32_t b;
    memcpy(&b, &x, 4);
    if ((b & 0x7f800000u) == 0x7f800000u) return x;   // inf / nan
    b = (b + 0x1000u) & 0xffffe000u;                  // round to nearest, ties away (cvt.rna)
    float r;
    memcpy(&r, &b, 4);
    return r;
}

void tc_destroy(TcContext *t)
{
    if (!t) return;
    if (t->blob) cudaFree(t->blob);
    if (t->arena) cudaFree(t->arena);
    if (t->maps_dev) cudaFree(t->maps_dev);
    if (t->prog_dev) cudaFree(t->prog_dev);
    if (t->err_dev) cudaFree(t->err_dev);
    delete t;
}

// Build the tensor-core context of a model that already has likelihood constants.  Returns nullptr and
// fills `why` when the shape is unsupported or the driver lacks the tensor-map entry point.
TcContext *tc_build(const linna_model *m, std::string &why)
{
    if (m->has_extra) { why = "extra linear branch not supported on the tensor-core path"; return nullptr; }
    if (!m->has_like) { why = "likelihood not set"; return nullptr; }
    if (m->quad_kind != LINNA_QUAD_CHOL) { why = "tensor-core path needs the Cholesky form of the quadratic"; return nullptr; }
    EncodeTiledFn encode = nullptr;
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
            qres != cudaDriverEntryPointSuccess) {
            why = "cuTensorMapEncodeTiled not available";
            return nullptr;
        }
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const int n_in = m->n_in, n_out = m->n_out;
    // ---- pack hi/lo weight operands ([N][ldk], K-major, the torch layout itself) and biases
    std::vector<float> h;
    auto alloc = [&](size_t n) { size_t o = (h.size() + 63) / 64 * 64; h.resize(o + n, 0.f); return o; };
    struct Mat { size_t hi, lo; int N, K, ldk; };
    std::vector<Mat> mats;
    auto put_mat = [&](const float *W, int N, int K, float scale, bool transpose) {
        Mat mt;
        mt.N = N, mt.K = K, mt.ldk = (K + 3) & ~3;
        mt.hi = alloc((size_t)N * mt.ldk), mt.lo = alloc((size_t)N * mt.ldk);
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) {
                const float w = scale * (transpose ? W[(size_t)k * N + n] : W[(size_t)n * K + k]);
                const float hi = host_tf32_hi(w);
                h[mt.hi + (size_t)n * mt.ldk + k] = hi;
                h[mt.lo + (size_t)n * mt.ldk + k] = w - hi;
            }
        mats.push_back(mt);
        return (int)mats.size() - 1;
    };
    auto put_vec = [&](const std::vector<float> &v) { size_t o = alloc(v.size()); std::copy(v.begin(), v.end(), h.begin() + o); return o; };

    int maxW = std::max(n_in, n_out), maxMid = 32;
    for (const OpHost &op : m->ops) maxW = std::max(maxW, std::max(op.in, op.out)), maxMid = std::max(maxMid, op.mid);
    // arena columns: each slot holds [hi | lo] of capacity cap
    const int capX = pad32(n_in), capW = pad32(maxW), capH = pad32(maxMid);
    const int slotX = 0, slotA = slotX + 2 * capX, slotB = slotA + 2 * capW, slotH = slotB + 2 * capW;
    const int ld = slotH + 2 * capH;

    TcProgram pg;
    memset(&pg, 0, sizeof pg);
    pg.in_dst = slotX, pg.in_lo = capX, pg.in_pad = capX, pg.ld = ld;
    pg.seg_kc = tc_seg_kc();
    int ns = 0;
    struct BiasRef { int step; size_t off; };
    std::vector<BiasRef> bias_refs;
    auto new_step = [&]() -> TcStep & {
        TcStep &s = pg.steps[ns++];
        memset(&s, 0, sizeof s);
        s.nphase = 1, s.bias_scale = 1.f;
        return s;
    };
    auto lo_of = [&](int slot) { return slot == slotX ? capX : slot == slotH ? capH : capW; };
    auto other = [&](int b) { return b == slotA ? slotB : slotA; };
    if ((int)m->ops.size() * 2 + 2 > TC_MAX_STEPS) { why = "too many layers"; return nullptr; }
    int cur = slotX;
    for (size_t i = 0; i < m->ops.size(); ++i) {
        const OpHost &op = m->ops[i];
        const bool last = i + 1 == m->ops.size();
        if (op.kind == LINNA_OP_LINEAR) {
            TcStep &s = new_step();
            const int mi = put_mat(op.w.data(), op.out, op.in, 1.f, false);
            s.srcA[0] = cur, s.srcLo[0] = lo_of(cur), s.K[0] = op.in, s.mapBhi[0] = 1 + 2 * mi, s.mapBlo[0] = 2 + 2 * mi;
            s.N = op.out, s.relu = op.act == LINNA_ACT_RELU, s.epi = last ? TC_EPI_HEAD : TC_EPI_ACT;
            s.dst = other(cur), s.dstLo = capW, s.dstPad = pad32(op.out);
            bias_refs.push_back({ns - 1, put_vec(op.b)});
            cur = s.dst;
        } else {
            if (!op.has_ws) { why = "identity skip not supported on the tensor-core path"; return nullptr; }
            TcStep &hs = new_step();
            const int m1 = put_mat(op.w.data(), op.mid, op.in, 1.f, false);
            hs.srcA[0] = cur, hs.srcLo[0] = lo_of(cur), hs.K[0] = op.in, hs.mapBhi[0] = 1 + 2 * m1, hs.mapBlo[0] = 2 + 2 * m1;
            hs.N = op.mid, hs.relu = 1, hs.epi = TC_EPI_ACT, hs.dst = slotH, hs.dstLo = capH, hs.dstPad = pad32(op.mid);
            bias_refs.push_back({ns - 1, put_vec(op.b)});
            TcStep &ys = new_step();
            const int m2 = put_mat(op.w2.data(), op.out, op.mid, op.alpha, false);   // alpha folded into W2
            const int m3 = put_mat(op.ws.data(), op.out, op.in, 1.f, false);
            ys.nphase = 2;
            ys.srcA[0] = slotH, ys.srcLo[0] = capH, ys.K[0] = op.mid, ys.mapBhi[0] = 1 + 2 * m2, ys.mapBlo[0] = 2 + 2 * m2;
            ys.srcA[1] = cur, ys.srcLo[1] = lo_of(cur), ys.K[1] = op.in, ys.mapBhi[1] = 1 + 2 * m3, ys.mapBlo[1] = 2 + 2 * m3;
            ys.N = op.out, ys.relu = 1, ys.epi = last ? TC_EPI_HEAD : TC_EPI_ACT, ys.bias_scale = op.alpha;
            ys.dst = other(cur), ys.dstLo = capW, ys.dstPad = pad32(op.out);
            bias_refs.push_back({ns - 1, put_vec(op.b2)});
            cur = ys.dst;
        }
    }
    {   // chi^2: r = d . Q with B[n][k] = Q[k][n]  (CHOL: Q = L => B = L^T, upper-trapezoidal k >= n)
        TcStep &q = new_step();
        const int mq = put_mat(m->quad.data(), n_out, n_out, 1.f, true);
        q.srcA[0] = cur, q.srcLo[0] = capW, q.K[0] = n_out, q.mapBhi[0] = 1 + 2 * mq, q.mapBlo[0] = 2 + 2 * mq;
        q.N = n_out, q.epi = TC_EPI_CHI2, q.tri = m->quad_kind == LINNA_QUAD_CHOL ? 1 : 0;
    }
    pg.n_steps = ns;

    TcContext *t = new TcContext();
    auto bail = [&](const std::string &msg) { why = msg; tc_destroy(t); return (TcContext *)nullptr; };
    t->ld = ld;
    t->grid = m->num_sms;
    if (cudaMalloc(&t->blob, h.size() * sizeof(float)) != cudaSuccess) return bail("cudaMalloc blob");
    if (cudaMemcpy(t->blob, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return bail("upload");
    const size_t arena_floats = (size_t)t->grid * TC_M * ld;
    if (cudaMalloc(&t->arena, arena_floats * sizeof(float)) != cudaSuccess) return bail("cudaMalloc arena");
    cudaMemset(t->arena, 0, arena_floats * sizeof(float));
    for (auto &br : bias_refs) pg.steps[br.step].bias = t->blob + br.off;

    // ---- tensor maps
    std::vector<CUtensorMap> maps(1 + 2 * mats.size());
    auto encode2d = [&](CUtensorMap *mp, float *base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_outer) {
        cuuint64_t dims[2] = {inner, outer};
        cuuint64_t strides[1] = {pitch_bytes};
        cuuint32_t box[2] = {(cuuint32_t)TC_KC, box_outer};
        cuuint32_t estr[2] = {1, 1};
        return encode(mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    if (encode2d(&maps[0], t->arena, (uint64_t)ld, (uint64_t)t->grid * TC_M, (uint64_t)ld * 4, TC_M) != CUDA_SUCCESS)
        return bail("cuTensorMapEncodeTiled(arena) failed");
    for (size_t i = 0; i < mats.size(); ++i) {
        const Mat &mt = mats[i];
        if (encode2d(&maps[1 + 2 * i], t->blob + mt.hi, (uint64_t)mt.K, (uint64_t)mt.N, (uint64_t)mt.ldk * 4, TC_NC) != CUDA_SUCCESS ||
            encode2d(&maps[2 + 2 * i], t->blob + mt.lo, (uint64_t)mt.K, (uint64_t)mt.N, (uint64_t)mt.ldk * 4, TC_NC) != CUDA_SUCCESS)
            return bail("cuTensorMapEncodeTiled(weights) failed");
    }
    if (cudaMalloc(&t->maps_dev, maps.size() * sizeof(CUtensorMap)) != cudaSuccess) return bail("cudaMalloc maps");
    cudaMemcpy(t->maps_dev, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    if (cudaMalloc(&t->prog_dev, sizeof(TcProgram)) != cudaSuccess) return bail("cudaMalloc prog");
    cudaMemcpy(t->prog_dev, &pg, sizeof(TcProgram), cudaMemcpyHostToDevice);
    if (cudaMalloc(&t->err_dev, sizeof(int)) != cudaSuccess) return bail("cudaMalloc err");
    cudaMemset(t->err_dev, 0, sizeof(int));
    if (cudaFuncSetAttribute(tc_lnp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) != cudaSuccess)
        return bail("cudaFuncSetAttribute(tc_lnp_kernel)");
    if (cudaDeviceSynchronize() != cudaSuccess) return bail("sync after tc_build");
    return t;
}

cudaError_t tc_launch_lnp(const linna_model *m, TcContext *t, const float *u, int64_t n, float *lnp, cudaStream_t stream)
{
    TcArgs a;
    memset(&a, 0, sizeof a);
    a.prog = t->prog_dev, a.maps = t->maps_dev, a.c = m->consts;
    a.in = u, a.lnp = lnp, a.arena = t->arena, a.n = n, a.err = t->err_dev;
    const int64_t tiles = (n + TC_M - 1) / TC_M;
    const int grid = (int)std::min<int64_t>(tiles, t->grid);
    tc_lnp_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace linna
