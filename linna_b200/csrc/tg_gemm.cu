// Tensor-core emulator TRAINING kernels (sm_100a: tcgen05 + TMEM + TMA): the forward pass, the loss, backward-data and
// the weight gradients (+ AdamW) of Predictor.train's inner loop (linna/predictor_gpu.py:273-288) as tiled GEMMs.
//
// The sampling kernel (tc_f16.cu) gives every CTA pair whole walkers and walks all layers on them; a training batch
// is only B = 500 rows, so here every LAYER is spread over the chip instead: one launch per layer, one CTA per
// 128 x 64 output tile (4 row tiles x N/64 column tiles; the weight-gradient launch covers every layer's 128 x 64
// tiles of dW at once, ~200 CTAs).
//
//   * bf16x3 products.  Every fp32 operand x is kept as three bf16 planes, h = bf16(x), m = bf16(x - h),
//     l = bf16(x - h - m) (24 significant bits, fp32's exponent range: gradients of any magnitude and weights that
//     drift during training need no scaling), and  D += A_m.B_m + A_h.B_l + A_l.B_h + A_h.B_m + A_m.B_h + A_h.B_h
//     with exact bf16 x bf16 products and fp32 accumulation in tensor memory.
//   * two-level accumulation as in tc_f16.cu: the tensor core truncates its fp32 accumulator on every instruction
//     (~0.5 ulp of the accumulator, toward zero), so every k-chunk (64 values of K = 24 instructions) is drained from
//     tensor memory and added with round-to-nearest into 64 register accumulators per epilogue thread -- and the leading
//     term goes to its OWN tensor-memory accumulator (h.h | everything else), so that the 20 small-term instructions of
//     a chunk truncate at THEIR magnitude (2^-8 of the sum and below) and only 4 instructions per chunk touch the leading
//     accumulator.  A segment is two k-chunks (8 leading instructions; two ring buffers of 2 x 64 columns): with three
//     accumulators drained every chunk the tcgen05.ld traffic, not the tensor core, set the pace.
//   * split-K: a layer of few tiles and a long contraction is cut into up to 4 k-ranges per tile, one CTA each; the
//     partial tiles go through an L2-resident workspace and the LAST CTA to arrive (atomic ticket) adds them in the
//     fixed order of the k-ranges and runs the epilogue -- deterministic, no floating-point atomics.
//   * every launch is a programmatic dependent launch: barrier set-up and tensor-memory allocation of layer l+1
//     overlap the tail of layer l (griddepcontrol.wait before the first dependent access).
//   * operands are 128-byte-swizzled tiles filled by TMA (3 stages x 72 KB): forward / backward-data read
//     activations [batch][features] and weights [out][in] (or the transposed copy) K-major; the weight-gradient GEMM
//     dW[n][k] = sum_b gz[b][n] x[b][k] contracts over the BATCH and reads the very same activation planes as
//     MN-major operands (instruction-descriptor transpose bits), so no transposed activation copy exists.  The
//     bias gradient is the extra column k == K of x, which every producer epilogue sets to 1.
//   * epilogues (one thread = one output row, 32 columns): bias + relu + mask bits + bf16x3 split (forward), the
//     normalised-space residual (loss head), chi^2 partials and d loss / d yhat (loss), relu-mask backward, and
//     AdamW with the refresh of every packed copy of the weight (weight gradients).
//   * warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, warps 2-9 = epilogue: a thread owns one output row
//     (tensor-memory lane) and HALF of the tile's 64 columns (warps 2-5 columns [0,32), warps 6-9 columns [32,64)).  The
//     epilogues are latency-bound chains run once per CTA (one warp per scheduler with four warps: 5-8 us per tile,
//     26 us for the AdamW epilogue, more than the k-loops); eight warps halve them.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <array>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "linna_host.hpp"

namespace linna {

constexpr int TG_BM = 128, TG_BN = 64, TG_KC = 64;            // output tile; k-chunk = 64 bf16 = one 128-byte swizzle row
constexpr int TG_STAGES = 3;
constexpr int TG_A_PLANE = TG_BM * 128;                        // 16 KB: 128 rows (or 2 x 64 contraction rows) x 128 B
constexpr int TG_B_PLANE = TG_BN * 128;                        // 8 KB
constexpr int TG_STAGE_BYTES = 3 * (TG_A_PLANE + TG_B_PLANE);  // 72 KB
constexpr int TG_SMEM_BYTES = TG_STAGES * TG_STAGE_BYTES + 1024;
constexpr int TG_EPI_THREADS = 256;   // 8 epilogue warps: warps w and w + 4 share a tensor-memory lane quarter and split the 64 columns
constexpr int TG_HN = TG_BN / 2;       // columns per epilogue thread
constexpr int TG_THREADS = 64 + TG_EPI_THREADS;
constexpr int TG_DEP_NT = 4;           // a fused res-block launch handles narrow steps of up to 4 column tiles

enum TgEpi : int32_t { TG_ACT = 0, TG_HEAD = 1, TG_LOSSQ = 2, TG_BWD = 3 };

struct TgStep {
    int32_t epi, nphase;
    int32_t mapA[2], mapB[2];     // tensor maps: A box 64 x 128 rows, B box 64 x 64 rows (both K-major)
    int32_t rowsA[2], rowsB[2];   // rows per bf16 plane of the operand (plane p starts at row p * rows)
    int32_t K[2];
    int32_t N;                    // valid output columns
    int32_t n_tiles;              // column tiles of 64
    int64_t out_off;              // element offset of the output's plane 0 inside the activation blob
    int32_t out_ld, out_rows;     // row pitch (elements) and rows per plane of the output
    int32_t mapOut, pad0_;        // tensor map of the output (box 64 x 128 rows, 128-byte swizzle): TMA store of the tile
    int32_t write_ones;           // column N of the output is the bias column of the next layer's input
    int32_t relu, save_mask, apply_mask;
    int64_t mask_off;             // word offset of this layer's relu bits (save) / of the producer's bits (apply)
    int32_t mask_ld;              // words per row
    float bias_scale;
    const float *bias;            // [N] fp32 bias inside the model's FP32 blob (kept current by every AdamW path), or nullptr
};

struct TgWLayer {
    int32_t mapA, mapB;           // gz planes / x planes as MN-major operands (box 64 x 64 batch rows)
    int32_t N, K;
    int32_t w_flat, b_flat;       // offsets into the flat parameter vector (b_flat < 0: no bias)
    float gscale, pack_scale;     // alpha of a res-block's second layer (gradient and packed copies), else 1
    int64_t wf_off, wb_off;       // plane 0 of the packed forward [N][K] / backward [K][N] copy (elements of the weight blob)
    int32_t wf_ld, wf_rows, wb_ld, wb_rows;
    // the FP32 kernels' packed copies inside the model blob (float offsets): Wt[k][ldf] + n, W[n][ldb] + k, bias[n]
    int64_t blob_f, blob_b, blob_bias;
    int32_t blob_ldf, blob_ldb;
};
struct TgWTile {
    int32_t layer, m0, n0, pad_;
};

struct TgArgs {
    const CUtensorMap *maps;
    const TgStep *step;
    __nv_bfloat16 *act;           // activation / gradient planes
    uint32_t *masks;
    int32_t B, B_pad;
    int32_t B_eff;                // rows the layer launches of this step covered: 128 x row tiles (contraction of the weight-gradient GEMM)
    // loss
    const float *target;          // [B][n_out] physical targets
    const float *cmd;             // [B] clamped chi2(target, data), or nullptr (chi^2 only)
    float *delta32;               // [B_pad][dld] fp32 copy of the residual
    int32_t dld;
    float *chi_part;              // [B_pad][chi_ld] chi^2 partial of every column tile
    int32_t chi_ld;
    Consts c;
    int32_t delta_kind;
    float loss_inv_B;
    int32_t want_grad;            // LOSSQ: also emit d loss / d yhat
    // weight gradients
    const TgWLayer *wlayers;
    const TgWTile *wtiles;
    __nv_bfloat16 *wblob;         // packed weight planes
    AdamArgs adam;
    int *err;
    float *ws;                    // split-K workspace: [tile][k-range][128][64] partial tiles
    int32_t *sem;                 // [tile] arrival tickets
    int32_t splitk;               // k-ranges per tile (CTAs per tile) of this launch: 1, 2 or 4
    long long *dbg;               // LINNA_TG_DEBUG: [launch slot][8] cycle stamps of CTA 0 (profiling aid), or nullptr
    int32_t dbg_slot;
    // Fused res-block launch: the first `dep_ctas` CTAs run the block's NARROW step (`step_dep`: hidden layer, or its
    // gradient), the others the two-phase step whose LAST phase contracts over that narrow output.  Both read the same
    // predecessor, so the wide CTAs run their first phase (the skip GEMM, K = layer width) at once and only the producer of
    // a CTA whose k-range reaches the second phase waits -- on per-tile flags that the narrow CTAs set to the launch's
    // sequence number when their tile has landed in global memory.  The narrow step's launch, its ~10 us and the gap
    // behind it disappear.
    const TgStep *step_dep;
    int32_t dep_ctas, dep_splitk;
    int32_t dep_ws_base, dep_sem_base;   // the narrow step's split-K workspace (partial-tile units) and tickets
    uint32_t *dep_cnt;            // [row tile][TG_DEP_NT] sequence number of the launch that last wrote the narrow tile
    uint32_t dep_target;          // this launch's sequence number
    int32_t dep_ntiles;           // column tiles of the narrow step (<= TG_DEP_NT)
};

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t tg_smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void tg_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tg_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tg_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tg_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tg_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool tg_mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(tg_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap (and report) instead of hanging the GPU.
__device__ __forceinline__ void tg_mbar_wait(uint64_t *bar, uint32_t parity, int *err, int code)
{
    if (tg_mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!tg_mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 2000000000LL) {
            if (err) atomicExch(err, code);
            __threadfence_system();
            __trap();
        }
    }
}
__device__ __forceinline__ bool tg_elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tg_tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(tg_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(tg_smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tg_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tg_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, 128-byte swizzle.  K-major operand: rows are 128 B apart, 8-row groups 1024 B
// (SBO); a K = 16 slice is a 32-byte step of the start address inside the swizzle row.  MN-major operand (the
// weight-gradient GEMM: a tile row is one CONTRACTION index holding 64 consecutive M/N values): 8 contraction rows
// are 1024 B (SBO), the next 64 M/N values are the next TMA box, `lbo` bytes further, and a K = 16 slice is 16 rows
// = 2048 B.
__device__ __forceinline__ uint64_t tg_sdesc(uint32_t saddr, uint32_t lbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16, bf16 inputs, fp32 accumulation, M = 128, N = 64; `mn_major`: both operands transposed (MN-major)
__device__ __forceinline__ uint32_t tg_idesc(bool mn_major)
{
    uint32_t d = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TG_BN >> 3) << 17) | ((uint32_t)(TG_BM >> 4) << 24);
    if (mn_major) d |= (1u << 15) | (1u << 16);
    return d;
}
__device__ __forceinline__ void tg_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tg_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tg_tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
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
}
__device__ __forceinline__ void tg_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// x = h + m + l, three bf16 values; two numbers per call (packed bf16x2 words)
__device__ __forceinline__ void tg_split3(float a, float b, uint32_t &h, uint32_t &m, uint32_t &l)
{
    const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(hh);
    const float ra = a - hf.x, rb = b - hf.y;
    const __nv_bfloat162 mm = __floats2bfloat162_rn(ra, rb);
    const float2 mf = __bfloat1622float2(mm);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(ra - mf.x, rb - mf.y);
    h = *reinterpret_cast<const uint32_t *>(&hh);
    m = *reinterpret_cast<const uint32_t *>(&mm);
    l = *reinterpret_cast<const uint32_t *>(&ll);
}

// ------------------------------------------------------------------------------------------ GEMM core
// One 128 x 64 output tile: sum over phases of A_p . B_p^T, result in `racc` of the epilogue threads (thread = tile
// row (warp & 3) * 32 + lane).  K-major: A box (kc, plane * rowsA + m0) of 128 rows, B box (kc, plane * rowsB + n0)
// of 64 rows.  MN-major (weight gradients): A boxes (m0 + 64 g, plane * rows + kc) for g = 0, 1, B box (n0, plane *
// rows + kc), 64 contraction rows each.
struct TgTileDesc {
    const CUtensorMap *mapA[2], *mapB[2];
    int32_t rowsA[2], rowsB[2], K[2];
    int32_t nphase, m0, n0;
    int32_t i0, i1;               // this CTA's range of the flattened (phase, k-chunk) list (split-K), [i0, i1)
    bool mn_major;
};

struct TgPipe {
    uint64_t full_bar[TG_STAGES], empty_bar[TG_STAGES], tfull_bar[2], tempty_bar[2];
    uint32_t tmem_slot;
};

__device__ __forceinline__ void tg_gemm_tile(const TgTileDesc &t, uint8_t *smem, TgPipe &pp, float (&racc)[TG_HN], int *err,
                                             long long *dbg = nullptr, const uint32_t *dep_flag = nullptr, uint32_t dep_target = 0,
                                             int dep_ntiles = 0)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (dbg && threadIdx.x == 64) dbg[0] = clock64();
    if (threadIdx.x == 0) {
        for (int s = 0; s < TG_STAGES; ++s) tg_mbar_init(&pp.full_bar[s], 1), tg_mbar_init(&pp.empty_bar[s], 1);
        for (int b = 0; b < 2; ++b) tg_mbar_init(&pp.tfull_bar[b], 1), tg_mbar_init(&pp.tempty_bar[b], TG_EPI_THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane < 2 * t.nphase) {   // descriptor fetch overlaps the predecessor's tail
        const CUtensorMap *mp = (lane & 1) ? t.mapB[lane >> 1] : t.mapA[lane >> 1];
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(mp)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tg_smem_u32(&pp.tmem_slot)), "r"(256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tg_fence_before();
    __syncthreads();
    tg_fence_after();
    const uint32_t tmem_base = pp.tmem_slot;
    const int nk0 = (t.K[0] + TG_KC - 1) / TG_KC;
    const int total = t.i1 - t.i0;
    // K-major GEMMs (forward / backward-data): the B operand is a packed WEIGHT matrix, which nothing in the launch chain
    // of a step writes (see the note on the epilogue constants in tg_layer_kernel), so the weight planes of the first
    // stages are requested before the wait for the predecessor; the activation planes follow after it.
    const int npre = t.mn_major ? 0 : (total < TG_STAGES ? total : TG_STAGES);
    if (warp == 0 && tg_elect_one()) {
        for (int s = 0; s < npre; ++s) {
            const int i = t.i0 + s, p = i < nk0 ? 0 : 1, kc = i < nk0 ? i : i - nk0;
            uint8_t *sb = smem + s * TG_STAGE_BYTES + 3 * TG_A_PLANE;
            tg_mbar_expect_tx(&pp.full_bar[s], TG_STAGE_BYTES);
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
                tg_tma_load_2d(sb + pl * TG_B_PLANE, t.mapB[p], &pp.full_bar[s], kc * TG_KC, pl * t.rowsB[p] + t.n0);
        }
    }
    __syncwarp();
    // everything above overlapped the previous kernel's tail (programmatic dependent launch); from here on its
    // results are read
    if (dbg && threadIdx.x == 64) dbg[1] = clock64();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (dbg && threadIdx.x == 64) dbg[2] = clock64();

    if (warp == 0) {
        // =============================== TMA producer ===============================
        int stage = 0;
        uint32_t ph = 0;
        bool dep_ok = dep_flag == nullptr;
        for (int i = t.i0; i < t.i1; ++i) {
            {
                const int p = i < nk0 ? 0 : 1, kc = i < nk0 ? i : i - nk0;
                if (p == 1 && !dep_ok) {
                    // fused res-block launch: the second phase reads the narrow step's output of this row tile, written
                    // by other CTAs of this launch (TMA stores, completed before the counter was bumped)
                    const long long t0 = clock64();
                    for (int q = 0; q < dep_ntiles; ++q)
                        for (;;) {
                            uint32_t v;
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(dep_flag + q) : "memory");
                            if (v == dep_target) break;
                            __nanosleep(40);
                            if (clock64() - t0 > 2000000000LL) {
                                if (err) atomicExch(err, 25);
                                __threadfence_system();
                                __trap();
                            }
                        }
                    asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy acquire -> the TMA loads below
                    dep_ok = true;
                }
                tg_mbar_wait(&pp.empty_bar[stage], ph ^ 1, err, 21);
                if (tg_elect_one()) {
                    uint8_t *sa = smem + stage * TG_STAGE_BYTES, *sb = sa + 3 * TG_A_PLANE;
                    const bool pre = i - t.i0 < npre;   // this stage's barrier is armed and its weight planes are on their way
                    if (!pre) tg_mbar_expect_tx(&pp.full_bar[stage], TG_STAGE_BYTES);
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl) {
                        if (!t.mn_major) {
                            tg_tma_load_2d(sa + pl * TG_A_PLANE, t.mapA[p], &pp.full_bar[stage], kc * TG_KC, pl * t.rowsA[p] + t.m0);
                            if (!pre) tg_tma_load_2d(sb + pl * TG_B_PLANE, t.mapB[p], &pp.full_bar[stage], kc * TG_KC, pl * t.rowsB[p] + t.n0);
                        } else {
                            tg_tma_load_2d(sa + pl * TG_A_PLANE, t.mapA[p], &pp.full_bar[stage], t.m0, pl * t.rowsA[p] + kc * TG_KC);
                            tg_tma_load_2d(sa + pl * TG_A_PLANE + TG_A_PLANE / 2, t.mapA[p], &pp.full_bar[stage], t.m0 + 64,
                                           pl * t.rowsA[p] + kc * TG_KC);
                            tg_tma_load_2d(sb + pl * TG_B_PLANE, t.mapB[p], &pp.full_bar[stage], t.n0, pl * t.rowsB[p] + kc * TG_KC);
                        }
                    }
                }
                __syncwarp();
                if (++stage == TG_STAGES) stage = 0, ph ^= 1;
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        int stage = 0;
        uint32_t ph = 0;
        const uint32_t idesc = tg_idesc(t.mn_major);
        const uint32_t kstep = t.mn_major ? (2048u >> 4) : (32u >> 4);   // descriptor start-address step of a K = 16 slice
        for (int it = 0; it < total; ++it) {
            const int seg = it >> 1, buf = seg & 1;
            const bool seg_first = (it & 1) == 0, seg_last = (it & 1) == 1 || it + 1 == total;
            if (seg_first) tg_mbar_wait(&pp.tempty_bar[buf], ((uint32_t)(seg >> 1) & 1u) ^ 1u, err, 23);
            tg_mbar_wait(&pp.full_bar[stage], ph, err, 22);
            tg_fence_after();
            if (tg_elect_one()) {
                const uint32_t sa = tg_smem_u32(smem + stage * TG_STAGE_BYTES), sb = sa + 3 * TG_A_PLANE;
                const uint64_t ah = tg_sdesc(sa, TG_A_PLANE / 2), am = tg_sdesc(sa + TG_A_PLANE, TG_A_PLANE / 2),
                               al = tg_sdesc(sa + 2 * TG_A_PLANE, TG_A_PLANE / 2);
                const uint64_t bh = tg_sdesc(sb, TG_B_PLANE), bm = tg_sdesc(sb + TG_B_PLANE, TG_B_PLANE),
                               bl = tg_sdesc(sb + 2 * TG_B_PLANE, TG_B_PLANE);
                const uint32_t d0 = tmem_base + buf * (2 * TG_BN), d1 = d0 + TG_BN;   // h.h | all smaller terms
#pragma unroll
                for (int ks = 0; ks < TG_KC / 16; ++ks) {
                    const uint64_t o = (uint64_t)(ks * kstep);
                    const uint32_t acc = (ks > 0 || !seg_first) ? 1u : 0u;
                    tg_mma(d1, am + o, bm + o, idesc, acc);
                    tg_mma(d1, ah + o, bl + o, idesc, 1u);
                    tg_mma(d1, al + o, bh + o, idesc, 1u);
                    tg_mma(d1, ah + o, bm + o, idesc, 1u);
                    tg_mma(d1, am + o, bh + o, idesc, 1u);
                    tg_mma(d0, ah + o, bh + o, idesc, acc);
                }
                tg_commit(&pp.empty_bar[stage]);                  // the operand stage is free when these MMAs retire
                if (seg_last) tg_commit(&pp.tfull_bar[buf]);      // and the partial tile can be drained
            }
            __syncwarp();
            if (++stage == TG_STAGES) stage = 0, ph ^= 1;
        }
    } else {
        // =============================== epilogue warps: drain every k-chunk ===============================
        const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp - 2) >> 2) * TG_HN);
#pragma unroll
        for (int j = 0; j < TG_HN; ++j) racc[j] = 0.f;
        const int nseg = (total + 1) >> 1;
        for (int seg = 0; seg < nseg; ++seg) {
            const int buf = seg & 1;
            tg_mbar_wait(&pp.tfull_bar[buf], (uint32_t)(seg >> 1) & 1u, err, 24);
            tg_fence_after();
            {
                uint32_t r0[32], r1[32];
                tg_tmem_ld32(tmem_lane + buf * (2 * TG_BN), r0);            // leading term
                tg_tmem_ld32(tmem_lane + buf * (2 * TG_BN) + TG_BN, r1);    // all smaller terms
                tg_tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < TG_HN; ++j) racc[j] += __uint_as_float(r1[j]) + __uint_as_float(r0[j]);
            }
            tg_fence_before();
            __syncwarp();
            if (lane == 0) tg_mbar_arrive(&pp.tempty_bar[buf]);
            if (dbg && threadIdx.x == 64 && seg == 0) dbg[3] = clock64();
        }
        if (dbg && threadIdx.x == 64) dbg[4] = clock64();
    }
}

__device__ __forceinline__ void tg_gemm_finish(TgPipe &pp)
{
    tg_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 1) {
        tg_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(pp.tmem_slot), "r"(256) : "memory");
    }
}

__device__ __forceinline__ void tg_tma_store_2d(const void *smem_src, const CUtensorMap *map, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :
                 : "l"(reinterpret_cast<uint64_t>(map)), "r"(tg_smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
// The output values [j0, j0 + 32) of tile row `row` -> the three bf16 planes of the staging tile in shared memory (3 x
// [128 rows x 128 B], 128-byte swizzle: 16-byte chunk c of a row sits at chunk c ^ (row & 7)), which one thread then
// hands to TMA.
__device__ __forceinline__ void tg_stage_planes(uint8_t *stg, int row, const float *v, int j0)
{
    const uint32_t base = tg_smem_u32(stg) + (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
#pragma unroll 1
    for (int j = j0; j < j0 + TG_HN; j += 8) {
        uint32_t h[4], m[4], l[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) tg_split3(v[j + e], v[j + e + 1], h[e >> 1], m[e >> 1], l[e >> 1]);
        const uint32_t o = base + ((((uint32_t)j >> 3) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o + TG_A_PLANE), "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o + 2 * TG_A_PLANE), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
    }
}
// barrier of the epilogue warps (the two service warps never join it)
__device__ __forceinline__ void tg_epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TG_EPI_THREADS) : "memory"); }
// the epilogue threads have staged their half rows: one of them sends the three planes of the tile to global memory
// `landed`: wait until the tile is in global memory (a fused launch signals other CTAs afterwards); otherwise only until
// the staging tile has been read -- the grid's completion makes the writes visible to the next launch.
__device__ __forceinline__ void tg_store_tile(uint8_t *stg, const CUtensorMap *map, int col0, int row0, int plane_rows, bool landed = false)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tg_epi_sync();
    if (threadIdx.x == 64) {
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) tg_tma_store_2d(stg + pl * TG_A_PLANE, map, col0, pl * plane_rows + row0);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (landed) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}
// [128 rows][64 floats] tile between global memory (row pitch `ld`) and shared memory (row pitch 65), coalesced: an
// epilogue warp moves half a row per instruction (lane = column; warps 2-5 the left half, warps 6-9 the right half of the
// 32 rows of their quarter), four rows in flight.  Rows >= nrows / columns >= ncols are read as 0 and not written.
constexpr int TG_FT_LD = TG_BN + 1;
__device__ __forceinline__ void tg_ftile_load(float *tile, const float *g, int64_t ld, int nrows, int ncols)
{
    const int warp = threadIdx.x >> 5, wq = warp & 3, j = TG_HN * ((warp - 2) >> 2) + (threadIdx.x & 31);
#pragma unroll 4
    for (int rr = 0; rr < 32; ++rr) {
        const int r = wq * 32 + rr;
        tile[r * TG_FT_LD + j] = (r < nrows && j < ncols) ? __ldg(g + (int64_t)r * ld + j) : 0.f;
    }
}
__device__ __forceinline__ void tg_ftile_store(const float *tile, float *g, int64_t ld, int nrows, int ncols)
{
    const int warp = threadIdx.x >> 5, wq = warp & 3, j = TG_HN * ((warp - 2) >> 2) + (threadIdx.x & 31);
#pragma unroll 4
    for (int rr = 0; rr < 32; ++rr) {
        const int r = wq * 32 + rr;
        if (r < nrows && j < ncols) g[(int64_t)r * ld + j] = tile[r * TG_FT_LD + j];
    }
}

// ------------------------------------------------------------------------------------------ layer kernel
__global__ void __launch_bounds__(TG_THREADS, 1) tg_layer_kernel(const TgArgs args)
{
    extern __shared__ uint8_t tg_smem_raw[];
    __shared__ TgPipe pp;
    __shared__ TgStep st;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(tg_smem_raw) + 1023) & ~(uintptr_t)1023);
    const bool is_dep = (int)blockIdx.x < args.dep_ctas;   // a narrow-step CTA of a fused res-block launch
    const int bid = is_dep ? (int)blockIdx.x : (int)blockIdx.x - args.dep_ctas;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(is_dep ? args.step_dep : args.step);
        for (int i = threadIdx.x; i < (int)(sizeof(TgStep) / 4); i += TG_THREADS) reinterpret_cast<uint32_t *>(&st)[i] = src[i];
    }
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next layer may set itself up while this one runs
    TgTileDesc t;
    t.nphase = st.nphase, t.mn_major = false;
    for (int p = 0; p < 2; ++p) {
        t.mapA[p] = args.maps + st.mapA[p], t.mapB[p] = args.maps + st.mapB[p];
        t.rowsA[p] = st.rowsA[p], t.rowsB[p] = st.rowsB[p], t.K[p] = st.K[p];
    }
    const int S = is_dep ? args.dep_splitk : args.splitk;
    const int tile = bid / S, kr = bid - tile * S;
    const int ws_base = is_dep ? args.dep_ws_base : 0;   // in partial tiles
    int32_t *sem = args.sem + (is_dep ? args.dep_sem_base : 0) + tile;
    const int nt = tile % st.n_tiles, mt = tile / st.n_tiles;
    t.m0 = mt * TG_BM, t.n0 = nt * TG_BN;
    {
        const int T = (st.K[0] + TG_KC - 1) / TG_KC + (st.nphase > 1 ? (st.K[1] + TG_KC - 1) / TG_KC : 0);
        t.i0 = (int)((int64_t)T * kr / S), t.i1 = (int)((int64_t)T * (kr + 1) / S);
    }
    // Per-column constants of the epilogue (the layer's bias; the output transform and the data vector for the loss head)
    // go to shared memory BEFORE the wait for the predecessor: read per column inside the epilogue loops they were a chain
    // of L2 round trips (3-5 k cycles per tile).  Nothing in the launch chain of a step writes them: parameters change
    // only in the weight-gradient / AdamW kernels, and the first kernel of every chain (tg_input_kernel) is an ordinary
    // launch that starts after those have completed.
    __shared__ float s_bias[TG_BN], s_cst[4][TG_BN];
    if (threadIdx.x < TG_BN) {
        const int col = t.n0 + (int)threadIdx.x;
        s_bias[threadIdx.x] = (st.bias && col < st.N) ? __ldg(st.bias + col) : 0.f;
    } else if (st.epi == TG_HEAD && threadIdx.x < 2 * TG_BN) {
        const int j = (int)threadIdx.x - TG_BN, col = t.n0 + j;
        const bool in = col < st.N;
        s_cst[0][j] = in ? __ldg(args.c.y_std + col) : 1.f;
        s_cst[1][j] = in ? __ldg(args.c.y_mean + col) : 0.f;
        s_cst[2][j] = (in && args.c.sigma) ? __ldg(args.c.sigma + col) : 1.f;
        s_cst[3][j] = in ? __ldg(args.c.data_hat + col) : 0.f;
    }
    float racc[TG_HN];
    long long *dbg = (args.dbg && bid == 0 && !is_dep) ? args.dbg + 8 * args.dbg_slot : nullptr;
    const bool needs_dep = !is_dep && args.dep_ctas > 0 && st.nphase > 1;
    tg_gemm_tile(t, smem, pp, racc, args.err, dbg, needs_dep ? args.dep_cnt + mt * TG_DEP_NT : nullptr, args.dep_target,
                 args.dep_ntiles);   // (its first barrier publishes the constants)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hh = (warp - 2) >> 2, j0 = TG_HN * hh;   // epilogue warps: this thread's half of the tile's columns
    __shared__ int s_last;
    __shared__ float s_part[TG_BM];
    bool do_epilogue = true;
    if (S > 1) {
        // split-K: park the partial tile in the workspace; the last k-range to arrive adds all of them in k-range order
        if (warp >= 2) {
            const int row = (warp & 3) * 32 + lane;
            // [tile][k-range][float4 column][row]: a warp instruction writes 32 consecutive rows = 512 contiguous bytes
            float4 *wsp = reinterpret_cast<float4 *>(args.ws) + (size_t)(ws_base + tile * S + kr) * (TG_BN / 4) * TG_BM + row;
#pragma unroll
            for (int j = 0; j < TG_HN; j += 4)
                wsp[(size_t)((j0 + j) >> 2) * TG_BM] = make_float4(racc[j], racc[j + 1], racc[j + 2], racc[j + 3]);
            __threadfence();
            tg_epi_sync();
            if (threadIdx.x == 64) {
                const int ticket = atomicAdd(sem, 1);
                s_last = ticket == S - 1 ? 1 : 0;
                if (ticket == S - 1) *sem = 0;   // ready for the next launch
                __threadfence();
            }
            tg_epi_sync();
            do_epilogue = s_last != 0;
            if (do_epilogue) {
#pragma unroll
                for (int j = 0; j < TG_HN; ++j) racc[j] = 0.f;
                // two k-ranges' loads in flight at a time (S is 2 or 4), added in k-range order
                for (int r = 0; r < S; r += 2) {
                    const float4 *src = reinterpret_cast<const float4 *>(args.ws) + (size_t)(ws_base + tile * S + r) * (TG_BN / 4) * TG_BM + row;
                    float4 va[TG_HN / 4], vb[TG_HN / 4];
#pragma unroll
                    for (int j = 0; j < TG_HN; j += 4) {
                        va[j >> 2] = __ldcg(src + (size_t)((j0 + j) >> 2) * TG_BM);
                        vb[j >> 2] = __ldcg(src + (size_t)(TG_BN / 4) * TG_BM + (size_t)((j0 + j) >> 2) * TG_BM);
                    }
#pragma unroll
                    for (int j = 0; j < TG_HN; j += 4) {
                        racc[j] += va[j >> 2].x, racc[j + 1] += va[j >> 2].y, racc[j + 2] += va[j >> 2].z, racc[j + 3] += va[j >> 2].w;
                        racc[j] += vb[j >> 2].x, racc[j + 1] += vb[j >> 2].y, racc[j + 2] += vb[j >> 2].z, racc[j + 3] += vb[j >> 2].w;
                    }
                }
            }
        }
    }
    if (warp >= 2 && do_epilogue) {
        // The operand stages are free (every MMA of this CTA has retired): stage 0 becomes the bf16x3 staging tile of the
        // output (48 KB, sent by TMA), stage 1 a [128][65] fp32 tile for the loss steps' coalesced global traffic.
        uint8_t *stg = smem;
        float *ftile = reinterpret_cast<float *>(smem + TG_STAGE_BYTES);
        const int row = (warp & 3) * 32 + lane;
        const int grow = t.m0 + row;
        const bool valid = grow < args.B;
        const int N = st.N, n0 = t.n0;
        const int nrows = args.B - t.m0 < TG_BM ? (args.B - t.m0 < 0 ? 0 : args.B - t.m0) : TG_BM;
        const int ncols = N - n0 < TG_BN ? (N - n0 < 0 ? 0 : N - n0) : TG_BN;
        const Consts &c = args.c;
        // The accumulators go to this thread's own half row of a [128][65] fp32 tile in shared memory and every epilogue is
        // a ROLLED loop over that half row: fully unrolled over the register accumulators the three epilogues were ~100 KB
        // of straight-line code executed once per CTA, i.e. bound by instruction fetch (measured: 8 - 12 us per tile).
        float *vrow = ftile + row * TG_FT_LD;
        float *yrow = reinterpret_cast<float *>(smem + TG_STAGE_BYTES + 36 * 1024) + row * TG_FT_LD;   // second tile (targets / residual)
#pragma unroll
        for (int j = 0; j < TG_HN; ++j) vrow[j0 + j] = racc[j];
        const int epi = st.epi;
        // this thread's word of the row's relu / validity bits: 32 columns = one word
        uint32_t *mword = args.masks + st.mask_off + (size_t)grow * st.mask_ld + (n0 >> 5) + hh;
        if (epi == TG_ACT || epi == TG_BWD) {
            const uint32_t mwh = st.apply_mask ? *mword : 0xffffffffu;
            const float *bias = st.bias;
            const float bscale = st.bias_scale;
            const bool relu = st.relu != 0, ones = st.write_ones != 0;
            uint32_t bits = 0u;
            // eight columns per trip, from the fp32 row straight into the three bf16 planes of the staging tile (128-byte
            // swizzle: 16-byte chunk c of a row sits at chunk c ^ (row & 7)): one pass, no write-back of the row
            const uint32_t stg_row = tg_smem_u32(stg) + (uint32_t)row * 128u, swz = (uint32_t)(row & 7);
#pragma unroll 1
            for (int g8 = 0; g8 < TG_HN; g8 += 8) {
                float y8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int jj = g8 + e, j = j0 + jj, col = n0 + j;
                    float y = vrow[j];
                    if (bias && col < N) y += bscale * s_bias[j];
                    if (relu) y = fmaxf(y, 0.f);
                    y = ((mwh >> jj) & 1u) ? y : 0.f;
                    bits |= (y > 0.f ? 1u : 0u) << jj;
                    if (!(valid && col < N)) y = (valid && col == N && ones) ? 1.f : 0.f;
                    y8[e] = y;
                }
                uint32_t h[4], m[4], l[4];
#pragma unroll
                for (int e = 0; e < 8; e += 2) tg_split3(y8[e], y8[e + 1], h[e >> 1], m[e >> 1], l[e >> 1]);
                const uint32_t o = stg_row + ((((uint32_t)(j0 + g8) >> 3) ^ swz) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o + TG_A_PLANE), "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o + 2 * TG_A_PLANE), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
            }
            if (st.save_mask) *mword = bits;
            tg_store_tile(stg, args.maps + st.mapOut, n0, t.m0, st.out_rows, is_dep);
            if (is_dep && threadIdx.x == 64) {   // the tile has landed (wait_group 0 above): tell the wide CTAs of this row tile
                asm volatile("fence.proxy.async;" ::: "memory");
                __threadfence();
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(args.dep_cnt + mt * TG_DEP_NT + nt), "r"(args.dep_target) : "memory");
            }
        } else if (epi == TG_HEAD) {
            // vrow = yhat - bias; residual in normalised space (Auxilleryfunc, linna/util.py:1070-1088)
            float *ytile = reinterpret_cast<float *>(smem + TG_STAGE_BYTES + 36 * 1024);
            tg_ftile_load(ytile, args.target + (size_t)t.m0 * N + n0, N, nrows, ncols);
            tg_epi_sync();
            const float *bias = st.bias;
            const float bscale = st.bias_scale;
            const int kind = args.delta_kind;
            uint32_t bits = 0u;
#pragma unroll 2
            for (int jj = 0; jj < TG_HN; ++jj) {
                const int j = j0 + jj, col = n0 + j;
                float dv = 0.f;
                if (valid && col < N) {
                    const float yh = vrow[j] + (bias ? bscale * s_bias[j] : 0.f);
                    const float ys = s_cst[0][j], ym = s_cst[1][j], sg = s_cst[2][j], dh = s_cst[3][j];
                    const float Y = yrow[j];
                    float tt = Y / sg;                                              // util.py:432
                    if (c.ypositive) tt = logf(tt);                                 // util.py:567-568
                    tt = (tt - ym) / ys;                                            // util.py:570
                    const bool ok = !(Y == 1e-30f || Y == 1e10f || dh == 1e-30f);    // util.py:1072
                    dv = kind == 0 ? tt - yh : kind == 1 ? tt - dh : yh - dh;
                    if (!ok) dv = 0.f;
                    bits |= (ok ? 1u : 0u) << jj;
                }
                vrow[j] = dv;
            }
            *mword = bits;
            tg_stage_planes(stg, row, vrow, j0);
            tg_store_tile(stg, args.maps + st.mapOut, n0, t.m0, st.out_rows);     // (its barrier also publishes the fp32 rows)
            tg_ftile_store(ftile, args.delta32 + (size_t)t.m0 * args.dld + n0, args.dld, TG_BM, TG_BN);
        } else {   // TG_LOSSQ: q = delta @ Chat^-1 ; chi2 += q . delta ; g_yhat = -2 q ok / (cmd B)
            float *dtile = reinterpret_cast<float *>(smem + TG_STAGE_BYTES + 36 * 1024);
            tg_ftile_load(dtile, args.delta32 + (size_t)t.m0 * args.dld + n0, args.dld, TG_BM, TG_BN);
            tg_epi_sync();
            const uint32_t okh = *mword;
            const float rs = (valid && args.cmd) ? -2.0f * args.loss_inv_B / __ldg(args.cmd + grow) : 0.f;
            float part = 0.f;
#pragma unroll 4
            for (int jj = 0; jj < TG_HN; ++jj) {
                const int j = j0 + jj, col = n0 + j;
                const float q = (valid && col < N) ? vrow[j] : 0.f;
                part = fmaf(q, yrow[j], part);
                vrow[j] = ((okh >> jj) & 1u) ? q * rs : 0.f;
            }
            // the row's chi^2 partial of this tile: left half + right half, in that order
            if (hh) s_part[row] = part;
            tg_epi_sync();
            if (!hh) args.chi_part[(size_t)grow * args.chi_ld + (n0 / TG_BN)] = part + s_part[row];
            if (args.want_grad) {
                tg_stage_planes(stg, row, vrow, j0);
                tg_store_tile(stg, args.maps + st.mapOut, n0, t.m0, st.out_rows);
            }
        }
    }
    if (dbg && threadIdx.x == 64) dbg[5] = clock64();
    tg_gemm_finish(pp);
    if (dbg && threadIdx.x == 64) dbg[6] = clock64();
}

// ------------------------------------------------------------------------------------------ weight gradients + AdamW
// torch.optim.AdamW (linna/predictor_gpu.py:267): decoupled weight decay, bias-corrected moments
__device__ __forceinline__ float tg_adamw(const AdamArgs &a, int idx, float g)
{
    float p = a.params[idx], m = a.m[idx], v = a.v[idx];
    p *= 1.0f - a.lr * a.wd;
    m = m + (1.0f - a.beta1) * (g - m);
    v = v * a.beta2 + (1.0f - a.beta2) * g * g;
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
    p = p - (a.lr / a.bc1) * (m / denom);
    a.params[idx] = p, a.m[idx] = m, a.v[idx] = v;
    return p;
}

__device__ __forceinline__ void tg_split3_1(float x, __nv_bfloat16 &h, __nv_bfloat16 &m, __nv_bfloat16 &l)
{
    h = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h);
    m = __float2bfloat16_rn(r1);
    l = __float2bfloat16_rn(r1 - __bfloat162float(m));
}

constexpr int TG_WT_LD = TG_BN + 1;   // padded row pitch of the gradient tile in shared memory

__global__ void __launch_bounds__(TG_THREADS, 1) tg_wgrad_kernel(const TgArgs args)
{
    extern __shared__ uint8_t tg_smem_raw[];
    __shared__ TgPipe pp;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(tg_smem_raw) + 1023) & ~(uintptr_t)1023);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const TgWTile wt = args.wtiles[blockIdx.x];
    const TgWLayer L = args.wlayers[wt.layer];
    TgTileDesc t;
    t.nphase = 1, t.mn_major = true;
    t.mapA[0] = t.mapA[1] = args.maps + L.mapA, t.mapB[0] = t.mapB[1] = args.maps + L.mapB;
    t.rowsA[0] = t.rowsA[1] = t.rowsB[0] = t.rowsB[1] = args.B_pad;
    t.K[0] = args.B_eff, t.K[1] = 0;
    t.m0 = wt.m0, t.n0 = wt.n0;
    t.i0 = 0, t.i1 = args.B_eff / TG_KC;
    float racc[TG_HN];
    long long *dbg = (args.dbg && blockIdx.x == 0) ? args.dbg + 8 * 40 : nullptr;
    tg_gemm_tile(t, smem, pp, racc, args.err, dbg);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= 2) {
        // The GEMM leaves thread = row n with 32 k-values; parameters, moments and the forward copies are contiguous in
        // k.  The tile goes through shared memory (the operand stages are free: every MMA has retired) so that pass 1
        // runs with lane = k (coalesced AdamW state, forward planes, backward FP32 copy) and pass 2 with lane = n
        // (backward planes, forward FP32 copy).  Warps w and w + 4 share the 32 rows of a quarter: 16 rows each in pass
        // 1, 32 of the 64 k-columns each in pass 2.
        float *tile = reinterpret_cast<float *>(smem);
        const int wq = warp & 3, row = wq * 32 + lane, hh = (warp - 2) >> 2;
#pragma unroll
        for (int j = 0; j < TG_HN; ++j) tile[row * TG_WT_LD + TG_HN * hh + j] = L.gscale * racc[j];
        tg_epi_sync();
        const AdamArgs &ad = args.adam;
        const int64_t pf = (int64_t)L.wf_rows * L.wf_ld;
        // four rows (eight coalesced 128-byte lines of each of p, m, v) in flight per batch
#pragma unroll 1
        for (int r0 = 16 * hh; r0 < 16 * hh + 16; r0 += 4) {
            int idx[8];
            float g[8], pp_[8], mm_[8], vv_[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int r = wq * 32 + r0 + (q >> 1), n = wt.m0 + r, j = 32 * (q & 1) + lane, k = wt.n0 + j;
                idx[q] = -1;
                if (n < L.N) {
                    if (k < L.K) idx[q] = L.w_flat + n * L.K + k;
                    else if (k == L.K && L.b_flat >= 0) idx[q] = L.b_flat + n;
                }
                g[q] = tile[r * TG_WT_LD + j];
            }
            if (ad.fuse) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (idx[q] >= 0) pp_[q] = ad.params[idx[q]], mm_[q] = ad.m[idx[q]], vv_[q] = ad.v[idx[q]];
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (idx[q] < 0) continue;
                const int r = wq * 32 + r0 + (q >> 1), n = wt.m0 + r, j = 32 * (q & 1) + lane, k = wt.n0 + j;
                if (!ad.fuse) { ad.grads[idx[q]] = g[q]; continue; }
                float p = pp_[q], m = mm_[q], v = vv_[q];
                p *= 1.0f - ad.lr * ad.wd;                       // torch.optim.AdamW (predictor_gpu.py:267)
                m = m + (1.0f - ad.beta1) * (g[q] - m);
                v = v * ad.beta2 + (1.0f - ad.beta2) * g[q] * g[q];
                const float denom = sqrtf(v) / ad.bc2_sqrt + ad.eps;
                p = p - (ad.lr / ad.bc1) * (m / denom);
                ad.params[idx[q]] = p, ad.m[idx[q]] = m, ad.v[idx[q]] = v;
                if (k < L.K) {
                    tile[r * TG_WT_LD + j] = p;
                    __nv_bfloat16 b0, b1, b2;
                    tg_split3_1(p * L.pack_scale, b0, b1, b2);
                    const int64_t of = L.wf_off + (int64_t)n * L.wf_ld + k;
                    args.wblob[of] = b0, args.wblob[of + pf] = b1, args.wblob[of + 2 * pf] = b2;
                    ad.blob[L.blob_b + (int64_t)n * L.blob_ldb + k] = p;
                } else
                    ad.blob[L.blob_bias + n] = p;
            }
        }
        if (ad.fuse) {
            tg_epi_sync();
            const int n = wt.m0 + row;
            if (n < L.N) {
                const int64_t pb = (int64_t)L.wb_rows * L.wb_ld;
#pragma unroll 8
                for (int j = TG_HN * hh; j < TG_HN * hh + TG_HN; ++j) {
                    const int k = wt.n0 + j;
                    if (k >= L.K) break;
                    const float p = tile[row * TG_WT_LD + j];
                    __nv_bfloat16 b0, b1, b2;
                    tg_split3_1(p * L.pack_scale, b0, b1, b2);
                    const int64_t ob = L.wb_off + (int64_t)k * L.wb_ld + n;
                    args.wblob[ob] = b0, args.wblob[ob + pb] = b1, args.wblob[ob + 2 * pb] = b2;
                    ad.blob[L.blob_f + (int64_t)k * L.blob_ldf + n] = p;
                }
            }
        }
    }
    if (dbg && threadIdx.x == 64) dbg[5] = clock64();
    tg_gemm_finish(pp);
    if (dbg && threadIdx.x == 64) dbg[6] = clock64();
}

// params -> packed bf16x3 planes of every weight matrix (set-up, load_state_dict, after the stand-alone AdamW of a
// data-parallel step).  One CTA per 128 x 64 tile of the weight-gradient tile table: the forward copy [n][k] is written
// with lane = k, the backward copy [k][n] with lane = n after a pass through shared memory -- written straight from a
// flat loop over the parameters, the backward copy was 32 separate sectors per warp store (4 M two-byte stores at C3).
__global__ void __launch_bounds__(256) tg_repack_kernel(const float *__restrict__ params, __nv_bfloat16 *wblob,
                                                        const TgWLayer *__restrict__ layers, const TgWTile *__restrict__ tiles)
{
    __shared__ float tile[TG_BM][TG_BN + 1];
    const TgWTile wt = tiles[blockIdx.x];
    const TgWLayer L = layers[wt.layer];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t pf = (int64_t)L.wf_rows * L.wf_ld, pb = (int64_t)L.wb_rows * L.wb_ld;
    for (int r = warp; r < TG_BM; r += 8) {
        const int n = wt.m0 + r;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 32 * h + lane, k = wt.n0 + j;
            float p = 0.f;
            if (n < L.N && k < L.K) {
                p = params[L.w_flat + n * L.K + k] * L.pack_scale;
                __nv_bfloat16 b0, b1, b2;
                tg_split3_1(p, b0, b1, b2);
                const int64_t of = L.wf_off + (int64_t)n * L.wf_ld + k;
                wblob[of] = b0, wblob[of + pf] = b1, wblob[of + 2 * pf] = b2;
            }
            tile[r][j] = p;
        }
    }
    __syncthreads();
    const int r = threadIdx.x & (TG_BM - 1), n = wt.m0 + r;
    if (n < L.N)
        for (int j = 32 * (threadIdx.x >> 7); j < 32 * (threadIdx.x >> 7) + 32; ++j) {
            const int k = wt.n0 + j;
            if (k >= L.K) break;
            __nv_bfloat16 b0, b1, b2;
            tg_split3_1(tile[r][j], b0, b1, b2);
            const int64_t ob = L.wb_off + (int64_t)k * L.wb_ld + n;
            wblob[ob] = b0, wblob[ob + pb] = b1, wblob[ob + 2 * pb] = b2;
        }
}

// physical parameters -> xhat = (theta' - mean)/std (util.py:483-497) as bf16x3 planes, bias column set to 1
__global__ void tg_input_kernel(const float *__restrict__ X, int B, int B_pad, int n_in, int ld, Consts c, __nv_bfloat16 *planes)
{
    const int64_t plane = (int64_t)B_pad * ld;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < B_pad * ld; e += gridDim.x * blockDim.x) {
        const int r = e / ld, i = e - r * ld;
        float x = 0.f;
        if (r < B) {
            if (i < n_in) {
                float th = X[(size_t)r * n_in + i];
                if (c.log10_flag && c.log10_flag[i]) th = log10f(th);
                x = (th - c.x_mean[i]) / c.x_std[i];
            } else if (i == n_in)
                x = 1.f;
        }
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        const float r1 = x - __bfloat162float(h);
        const __nv_bfloat16 m = __float2bfloat16_rn(r1);
        planes[e] = h, planes[e + plane] = m, planes[e + 2 * plane] = __float2bfloat16_rn(r1 - __bfloat162float(m));
    }
}

// loss rows and their mean from the per-tile chi^2 partials (fixed summation order: deterministic)
__global__ void tg_loss_kernel(const float *__restrict__ chi_part, int chi_ld, int n_tiles, const float *__restrict__ cmd, int B,
                               float *__restrict__ rows, float *__restrict__ mean)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (int b = threadIdx.x; b < B; b += 256) {
        double chi = 0.0;
        for (int t = 0; t < n_tiles; ++t) chi += (double)chi_part[(size_t)b * chi_ld + t];
        const float l = cmd ? (float)chi / cmd[b] : (float)chi;                      // util.py:1087
        rows[b] = l;
        s += (double)l;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0 && mean) mean[0] = (float)(sh[0] / (double)B);              // util.py:1114-1115
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*TgEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TgContext {
    int B_pad = 0, n_in = 0, n_out = 0;
    __nv_bfloat16 *act = nullptr, *wblob = nullptr;
    uint32_t *masks = nullptr;
    float *delta32 = nullptr, *chi_part = nullptr;
    int dld = 0, chi_ld = 0;
    CUtensorMap *maps_dev = nullptr;
    TgStep *steps_dev = nullptr;
    TgWLayer *wlayers_dev = nullptr;
    TgWTile *wtiles_dev = nullptr;
    int *err_dev = nullptr;
    std::vector<TgStep> steps;         // forward ..., HEAD, LOSSQ, backward ...
    int n_fwd = 0, i_head = 0, i_lossq = 0, n_steps = 0;
    int n_wlayers = 0, n_wtiles = 0;
    int64_t in_off = 0;
    int in_ld = 0;
    int lossq_tiles = 0;
    int64_t max_wn = 0;
    float *ws = nullptr;
    int32_t *sem = nullptr;
    int max_tiles = 0, num_sms = 148;
    long long *dbg = nullptr;     // LINNA_TG_DEBUG
    // Weight-gradient buckets.  The weight gradients of a layer need that layer's d loss / d z, and -- with AdamW fused
    // into the same launch -- every backward-data step that still reads the layer's weights to have run.  Both hold
    // once the backward chain has passed the layer, so the tiles are grouped into up to three buckets of layers (last
    // layers first) and every bucket but the last is launched on its own stream as soon as the chain has passed it,
    // overlapping the rest of the chain; the last bucket follows the chain on the caller's stream.  In a data-parallel
    // step the buckets are also the units of the gradient all-reduce (contiguous ranges of the flat parameter vector).
    int n_buckets = 1;
    int bucket_tile0[3] = {0, 0, 0}, bucket_ntiles[3] = {0, 0, 0};
    int bucket_ready_step[3] = {0, 0, 0};          // launch after this step of the program has been enqueued
    int64_t bucket_flat_lo[3] = {0, 0, 0}, bucket_flat_hi[3] = {0, 0, 0};
    cudaStream_t bucket_stream[3] = {nullptr, nullptr, nullptr};   // nullptr: the caller's stream
    cudaEvent_t bucket_ready[3] = {nullptr, nullptr, nullptr}, bucket_done[3] = {nullptr, nullptr, nullptr};
    bool buckets_unjoined = false;                  // the last grad-out step left the bucket streams for the caller to join
    // the loss reduction of a training step (one CTA; only the host reads it) runs beside the backward chain
    cudaStream_t loss_stream = nullptr;
    cudaEvent_t loss_ready = nullptr, loss_done = nullptr;
    // fused res-block launches (TgArgs::step_dep): fuse_with_next[si] = step si is the narrow step of a res-block and
    // step si + 1 the two-phase step that consumes it; dep_cnt [n_steps][m_tiles_max] counters, dep_seq[si] launches so far
    std::vector<uint8_t> fuse_with_next;
    std::vector<uint32_t> dep_seq;
    uint32_t *dep_cnt = nullptr;
    int m_tiles_max = 0;
};

cudaError_t tg_repack(TgContext *t, const float *params, cudaStream_t stream);

void tg_destroy(TgContext *t)
{
    if (!t) return;
    cudaFree(t->act), cudaFree(t->wblob), cudaFree(t->masks), cudaFree(t->delta32), cudaFree(t->chi_part);
    cudaFree(t->maps_dev), cudaFree(t->steps_dev), cudaFree(t->wlayers_dev), cudaFree(t->wtiles_dev), cudaFree(t->err_dev);
    cudaFree(t->ws), cudaFree(t->sem), cudaFree(t->dbg), cudaFree(t->dep_cnt);
    if (t->loss_stream) cudaStreamDestroy(t->loss_stream);
    if (t->loss_ready) cudaEventDestroy(t->loss_ready);
    if (t->loss_done) cudaEventDestroy(t->loss_done);
    for (int b = 0; b < 3; ++b) {
        if (t->bucket_stream[b]) cudaStreamDestroy(t->bucket_stream[b]);
        if (t->bucket_ready[b]) cudaEventDestroy(t->bucket_ready[b]);
        if (t->bucket_done[b]) cudaEventDestroy(t->bucket_done[b]);
    }
    delete t;
}

static inline int tg_pad(int n, int q) { return (n + q - 1) / q * q; }

// Build the tensor-core training context: plane buffers, tensor maps, the layer-step list and the weight-gradient
// tile table.  `flat_off` gives, per op, the offsets of (w, b, w2, b2, ws) in the flat parameter vector.  Returns
// nullptr and fills `why` when the network shape is not covered (identity skips, extra linear branch).
TgContext *tg_build(const linna_model *m, const std::vector<std::array<int, 5>> &flat_off,
                    const std::vector<std::array<const float *, 2>> &bias_ptr, const std::vector<std::array<int64_t, 8>> &blob_off,
                    std::string &why)
{
    if (m->has_extra) { why = "extra linear branch"; return nullptr; }
    for (const OpHost &op : m->ops)
        if (op.kind == LINNA_OP_RES && !op.has_ws) { why = "identity skip"; return nullptr; }
    if (m->ops.back().kind != LINNA_OP_LINEAR || m->ops.back().act != LINNA_ACT_NONE) { why = "last layer must be linear"; return nullptr; }
    TgEncodeFn encode = nullptr;
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
            qres != cudaDriverEntryPointSuccess) {
            why = "cuTensorMapEncodeTiled not available";
            return nullptr;
        }
        encode = reinterpret_cast<TgEncodeFn>(fn);
    }
    TgContext *t = new TgContext();
    auto bail = [&](const std::string &msg) { why = msg; tg_destroy(t); return (TgContext *)nullptr; };
    const int nops = (int)m->ops.size();
    const int B_pad = tg_pad(m->max_batch, 128);
    t->B_pad = B_pad, t->n_in = m->n_in, t->n_out = m->n_out;

    // ---- activation / gradient matrices: [3 planes][B_pad][ld], ld = pad64(width + 1) (the +1 is the bias column)
    struct Mat { int64_t off; int ld; int mapA, mapMN; };
    int64_t act_elems = 0;
    std::vector<Mat> mats;
    auto new_mat = [&](int width) {
        Mat a;
        a.ld = tg_pad(width + 1, 64), a.off = act_elems, a.mapA = a.mapMN = -1;
        act_elems += 3 * (int64_t)B_pad * a.ld;
        mats.push_back(a);
        return (int)mats.size() - 1;
    };
    const int mX = new_mat(m->n_in);
    std::vector<int> mAct(nops), mHid(nops, -1), mGz(nops), mGzh(nops, -1);
    for (int i = 0; i < nops; ++i) {
        mAct[i] = new_mat(m->ops[i].out);      // output of op i (for the last op: the residual delta)
        mGz[i] = new_mat(m->ops[i].out);       // d loss / d (pre-activation of op i's output)
        if (m->ops[i].kind == LINNA_OP_RES) mHid[i] = new_mat(m->ops[i].mid), mGzh[i] = new_mat(m->ops[i].mid);
    }
    // ---- packed weights: forward [N_pad64][pad64(K)] and backward [K_pad64][pad64(N)], three planes each
    struct WMat { int64_t off; int ld, rows; int map; };
    int64_t w_elems = 0;
    auto new_w = [&](int rows, int cols) {
        WMat w;
        w.ld = tg_pad(cols, 64), w.rows = tg_pad(rows, 64), w.off = w_elems, w.map = -1;
        w_elems += 3 * (int64_t)w.rows * w.ld;
        return w;
    };
    struct OpW { WMat f, b, f2, b2, fs, bs; };
    std::vector<OpW> ow(nops);
    for (int i = 0; i < nops; ++i) {
        const OpHost &op = m->ops[i];
        if (op.kind == LINNA_OP_LINEAR) {
            ow[i].f = new_w(op.out, op.in), ow[i].b = new_w(op.in, op.out);
        } else {
            ow[i].f = new_w(op.mid, op.in), ow[i].b = new_w(op.in, op.mid);
            ow[i].f2 = new_w(op.out, op.mid), ow[i].b2 = new_w(op.mid, op.out);
            ow[i].fs = new_w(op.out, op.in), ow[i].bs = new_w(op.in, op.out);
        }
    }
    WMat wQ = new_w(m->n_out, m->n_out);   // Chat^-1 (symmetrised): q = delta @ Chat^-1, B[n][k] = Chat^-1[k][n] = itself

    if (cudaMalloc(&t->act, act_elems * sizeof(__nv_bfloat16)) != cudaSuccess) return bail("cudaMalloc activations");
    cudaMemset(t->act, 0, act_elems * sizeof(__nv_bfloat16));
    if (cudaMalloc(&t->wblob, w_elems * sizeof(__nv_bfloat16)) != cudaSuccess) return bail("cudaMalloc weights");
    cudaMemset(t->wblob, 0, w_elems * sizeof(__nv_bfloat16));

    // ---- tensor maps
    std::vector<CUtensorMap> maps;
    auto add_map = [&](void *base, int ld, int rows_total, int box_rows) -> int {
        CUtensorMap mp;
        cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows_total};
        cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
        cuuint32_t estr[2] = {1, 1};
        if (encode(&mp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
        maps.push_back(mp);
        return (int)maps.size() - 1;
    };
    for (Mat &a : mats) {
        a.mapA = add_map(t->act + a.off, a.ld, 3 * B_pad, 128);
        a.mapMN = add_map(t->act + a.off, a.ld, 3 * B_pad, 64);
        if (a.mapA < 0 || a.mapMN < 0) return bail("cuTensorMapEncodeTiled(activations) failed");
    }
    auto map_w = [&](WMat &w) { w.map = add_map(t->wblob + w.off, w.ld, 3 * w.rows, 64); return w.map >= 0; };
    for (int i = 0; i < nops; ++i) {
        bool ok = map_w(ow[i].f) && map_w(ow[i].b);
        if (m->ops[i].kind == LINNA_OP_RES) ok = ok && map_w(ow[i].f2) && map_w(ow[i].b2) && map_w(ow[i].fs) && map_w(ow[i].bs);
        if (!ok) return bail("cuTensorMapEncodeTiled(weights) failed");
    }
    if (!map_w(wQ)) return bail("cuTensorMapEncodeTiled(Chat^-1) failed");

    // ---- masks: relu bits of every forward output (+ the ok bits of the loss), words per row
    int64_t mask_words = 0;
    std::vector<int64_t> mkY(nops, -1), mkH(nops, -1);
    std::vector<int> mkYld(nops, 0), mkHld(nops, 0);
    auto new_mask = [&](int width, int &ld) {
        ld = tg_pad(width + 1, 64) / 32;
        int64_t o = mask_words;
        mask_words += (int64_t)B_pad * ld;
        return o;
    };
    for (int i = 0; i < nops; ++i) {
        mkY[i] = new_mask(m->ops[i].out, mkYld[i]);
        if (m->ops[i].kind == LINNA_OP_RES) mkH[i] = new_mask(m->ops[i].mid, mkHld[i]);
    }
    if (cudaMalloc(&t->masks, std::max<int64_t>(mask_words, 64) * sizeof(uint32_t)) != cudaSuccess) return bail("cudaMalloc masks");
    cudaMemset(t->masks, 0, std::max<int64_t>(mask_words, 64) * sizeof(uint32_t));
    t->dld = tg_pad(m->n_out + 1, 64);
    t->chi_ld = t->dld / 64;
    if (cudaMalloc(&t->delta32, (size_t)B_pad * t->dld * sizeof(float)) != cudaSuccess) return bail("cudaMalloc delta");
    if (cudaMalloc(&t->chi_part, (size_t)B_pad * t->chi_ld * sizeof(float)) != cudaSuccess) return bail("cudaMalloc chi");
    cudaMemset(t->delta32, 0, (size_t)B_pad * t->dld * sizeof(float));
    cudaMemset(t->chi_part, 0, (size_t)B_pad * t->chi_ld * sizeof(float));

    // ---- steps
    auto base_step = [&](int epi, int out_mat, int N) {
        TgStep s;
        memset(&s, 0, sizeof s);
        s.epi = epi, s.nphase = 1, s.N = N, s.bias = nullptr, s.bias_scale = 1.f;
        s.out_off = mats[out_mat].off, s.out_ld = mats[out_mat].ld, s.out_rows = B_pad, s.mapOut = mats[out_mat].mapA;
        s.n_tiles = mats[out_mat].ld / 64;
        return s;
    };
    auto set_phase = [&](TgStep &s, int p, int a_mat, const WMat &w, int K) {
        s.mapA[p] = mats[a_mat].mapA, s.rowsA[p] = B_pad, s.mapB[p] = w.map, s.rowsB[p] = w.rows, s.K[p] = K;
    };
    auto producer_mask = [&](int i, int64_t &off, int &ld) -> bool {   // relu bits of the producer of op i's input
        if (i <= 0) return false;
        const OpHost &pv = m->ops[i - 1];
        if (!(pv.kind == LINNA_OP_RES || pv.act == LINNA_ACT_RELU)) return false;
        off = mkY[i - 1], ld = mkYld[i - 1];
        return true;
    };
    int cur = mX;
    for (int i = 0; i < nops; ++i) {
        const OpHost &op = m->ops[i];
        const bool last = i + 1 == nops;
        if (op.kind == LINNA_OP_LINEAR) {
            TgStep s = base_step(last ? TG_HEAD : TG_ACT, mAct[i], op.out);
            set_phase(s, 0, cur, ow[i].f, op.in);
            s.bias = bias_ptr[i][0], s.relu = (!last && op.act == LINNA_ACT_RELU) ? 1 : 0, s.save_mask = s.relu, s.write_ones = last ? 0 : 1;
            s.mask_off = mkY[i], s.mask_ld = mkYld[i];     // HEAD: the ok bits of the loss live in the last op's mask
            t->steps.push_back(s);
        } else {
            TgStep h = base_step(TG_ACT, mHid[i], op.mid);
            set_phase(h, 0, cur, ow[i].f, op.in);
            h.bias = bias_ptr[i][0], h.relu = 1, h.save_mask = 1, h.write_ones = 1, h.mask_off = mkH[i], h.mask_ld = mkHld[i];
            t->steps.push_back(h);
            TgStep y = base_step(TG_ACT, mAct[i], op.out);
            y.nphase = 2;
            set_phase(y, 0, cur, ow[i].fs, op.in);
            set_phase(y, 1, mHid[i], ow[i].f2, op.mid);
            y.bias = bias_ptr[i][1], y.bias_scale = op.alpha, y.relu = 1, y.save_mask = 1, y.write_ones = 1;
            y.mask_off = mkY[i], y.mask_ld = mkYld[i];
            t->steps.push_back(y);
        }
        cur = mAct[i];
    }
    t->n_fwd = (int)t->steps.size() - 1;
    t->i_head = t->n_fwd;
    {   // q = delta @ Chat^-1
        TgStep q = base_step(TG_LOSSQ, mGz[nops - 1], m->n_out);
        set_phase(q, 0, mAct[nops - 1], wQ, m->n_out);
        q.mask_off = mkY[nops - 1], q.mask_ld = mkYld[nops - 1];
        t->i_lossq = (int)t->steps.size();
        t->lossq_tiles = q.n_tiles;
        t->steps.push_back(q);
    }
    std::vector<int> op_last_step(nops, 0);   // last step of the program that belongs to op i (its weights are free afterwards)
    op_last_step[nops - 1] = t->i_lossq;
    for (int i = nops - 1; i >= 1; --i) {   // backward-data; d loss / d xhat is not needed
        const OpHost &op = m->ops[i];
        int64_t moff = 0;
        int mld = 0;
        const bool pm = producer_mask(i, moff, mld);
        if (op.kind == LINNA_OP_LINEAR) {
            TgStep s = base_step(TG_BWD, mGz[i - 1], op.in);
            set_phase(s, 0, mGz[i], ow[i].b, op.out);
            if (pm) s.apply_mask = 1, s.mask_off = moff, s.mask_ld = mld;
            t->steps.push_back(s);
        } else {
            TgStep h = base_step(TG_BWD, mGzh[i], op.mid);
            set_phase(h, 0, mGz[i], ow[i].b2, op.out);
            h.apply_mask = 1, h.mask_off = mkH[i], h.mask_ld = mkHld[i];
            t->steps.push_back(h);
            TgStep x = base_step(TG_BWD, mGz[i - 1], op.in);
            x.nphase = 2;
            set_phase(x, 0, mGz[i], ow[i].bs, op.out);
            set_phase(x, 1, mGzh[i], ow[i].b, op.mid);
            if (pm) x.apply_mask = 1, x.mask_off = moff, x.mask_ld = mld;
            t->steps.push_back(x);
        }
        op_last_step[i] = (int)t->steps.size() - 1;
    }
    if (m->ops[0].kind == LINNA_OP_RES) {   // the hidden gradient of a leading res-block is still needed for its weights
        const OpHost &op = m->ops[0];
        TgStep h = base_step(TG_BWD, mGzh[0], op.mid);
        set_phase(h, 0, mGz[0], ow[0].b2, op.out);
        h.apply_mask = 1, h.mask_off = mkH[0], h.mask_ld = mkHld[0];
        t->steps.push_back(h);
    }
    t->n_steps = (int)t->steps.size();
    op_last_step[0] = t->n_steps - 1;
    t->in_off = mats[mX].off, t->in_ld = mats[mX].ld;

    // ---- weight-gradient layers and tiles
    std::vector<TgWLayer> wl;
    auto pad4i = [](int n) { return (n + 3) & ~3; };
    auto add_wl = [&](int gz_mat, int x_mat, int N, int K, int wf, int bf, float gs, const WMat &f, const WMat &b, int64_t bl_f,
                      int64_t bl_b, int64_t bl_bias) {
        TgWLayer L;
        memset(&L, 0, sizeof L);
        L.blob_f = bl_f, L.blob_b = bl_b, L.blob_bias = bl_bias, L.blob_ldf = pad4i(N), L.blob_ldb = pad4i(K);
        L.mapA = mats[gz_mat].mapMN, L.mapB = mats[x_mat].mapMN, L.N = N, L.K = K, L.w_flat = wf, L.b_flat = bf;
        L.gscale = gs, L.pack_scale = gs;
        L.wf_off = f.off, L.wf_ld = f.ld, L.wf_rows = f.rows, L.wb_off = b.off, L.wb_ld = b.ld, L.wb_rows = b.rows;
        wl.push_back(L);
    };
    for (int i = 0; i < nops; ++i) {
        const OpHost &op = m->ops[i];
        const int xin = i == 0 ? mX : mAct[i - 1];
        if (op.kind == LINNA_OP_LINEAR) {
            add_wl(mGz[i], xin, op.out, op.in, flat_off[i][0], flat_off[i][1], 1.f, ow[i].f, ow[i].b, blob_off[i][0], blob_off[i][1],
                   blob_off[i][2]);
        } else {
            add_wl(mGzh[i], xin, op.mid, op.in, flat_off[i][0], flat_off[i][1], 1.f, ow[i].f, ow[i].b, blob_off[i][0], blob_off[i][1],
                   blob_off[i][2]);
            add_wl(mGz[i], mHid[i], op.out, op.mid, flat_off[i][2], flat_off[i][3], op.alpha, ow[i].f2, ow[i].b2, blob_off[i][3],
                   blob_off[i][4], blob_off[i][5]);
            add_wl(mGz[i], xin, op.out, op.in, flat_off[i][4], -1, 1.f, ow[i].fs, ow[i].bs, blob_off[i][6], blob_off[i][7], -1);
        }
    }
    // op of every weight-gradient layer, and the bucket of every op: the last two ops, the middle ones, the first two
    std::vector<int> wl_op;
    for (int i = 0; i < nops; ++i) wl_op.insert(wl_op.end(), m->ops[i].kind == LINNA_OP_LINEAR ? 1 : 3, i);
    const bool split = nops >= 5 && !getenv("LINNA_TG_NO_BUCKETS");
    auto bucket_of = [&](int op) { return !split ? 0 : op >= nops - 2 ? 0 : op >= 2 ? 1 : 2; };
    t->n_buckets = split ? 3 : 1;
    std::vector<TgWTile> tiles;
    for (size_t l = 0; l < wl.size(); ++l) {
        const int kmax = wl[l].K + (wl[l].b_flat >= 0 ? 1 : 0);
        for (int m0 = 0; m0 < wl[l].N; m0 += TG_BM)
            for (int n0 = 0; n0 < kmax; n0 += TG_BN) tiles.push_back(TgWTile{(int32_t)l, m0, n0, 0});
    }
    // bucket by bucket; inside a bucket large tiles first: the 64 tiles of the widest skip layer should not start last
    std::stable_sort(tiles.begin(), tiles.end(), [&](const TgWTile &a, const TgWTile &b) {
        const int ba = bucket_of(wl_op[a.layer]), bb = bucket_of(wl_op[b.layer]);
        if (ba != bb) return ba < bb;
        return (int64_t)wl[a.layer].N * wl[a.layer].K > (int64_t)wl[b.layer].N * wl[b.layer].K;
    });
    for (int b = 0; b < t->n_buckets; ++b) {
        t->bucket_tile0[b] = (int)tiles.size(), t->bucket_ntiles[b] = 0;
        t->bucket_flat_lo[b] = INT64_MAX, t->bucket_flat_hi[b] = 0, t->bucket_ready_step[b] = 0;
    }
    for (size_t i = 0; i < tiles.size(); ++i) {
        const int b = bucket_of(wl_op[tiles[i].layer]);
        t->bucket_tile0[b] = std::min(t->bucket_tile0[b], (int)i), ++t->bucket_ntiles[b];
    }
    for (size_t l = 0; l < wl.size(); ++l) {
        const int b = bucket_of(wl_op[l]);
        t->bucket_ready_step[b] = std::max(t->bucket_ready_step[b], op_last_step[wl_op[l]]);
        int64_t lo = wl[l].w_flat, hi = (int64_t)wl[l].w_flat + (int64_t)wl[l].N * wl[l].K;
        if (wl[l].b_flat >= 0) lo = std::min<int64_t>(lo, wl[l].b_flat), hi = std::max<int64_t>(hi, (int64_t)wl[l].b_flat + wl[l].N);
        t->bucket_flat_lo[b] = std::min(t->bucket_flat_lo[b], lo), t->bucket_flat_hi[b] = std::max(t->bucket_flat_hi[b], hi);
    }
    for (int b = 0; b + 1 < t->n_buckets; ++b) {   // every bucket but the last gets its own stream
        if (cudaStreamCreateWithFlags(&t->bucket_stream[b], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&t->bucket_ready[b], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&t->bucket_done[b], cudaEventDisableTiming) != cudaSuccess)
            return bail("bucket streams");
    }
    if (cudaStreamCreateWithFlags(&t->loss_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&t->loss_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&t->loss_done, cudaEventDisableTiming) != cudaSuccess)
        return bail("loss stream");
    t->n_wlayers = (int)wl.size(), t->n_wtiles = (int)tiles.size();
    for (const TgWLayer &L : wl) t->max_wn = std::max<int64_t>(t->max_wn, (int64_t)L.N * L.K);

    // Chat^-1 planes (constant): B[n][k] = Chat^-1[k][n]
    {
        std::vector<__nv_bfloat16> h((size_t)3 * wQ.rows * wQ.ld, __float2bfloat16_rn(0.f));
        const int n = m->n_out;
        for (int r = 0; r < n; ++r)
            for (int k = 0; k < n; ++k) {
                const float x = m->icov_hat[(size_t)k * n + r];
                const __nv_bfloat16 a = __float2bfloat16_rn(x);
                const float r1 = x - __bfloat162float(a);
                const __nv_bfloat16 b = __float2bfloat16_rn(r1);
                const size_t o = (size_t)r * wQ.ld + k, pl = (size_t)wQ.rows * wQ.ld;
                h[o] = a, h[o + pl] = b, h[o + 2 * pl] = __float2bfloat16_rn(r1 - __bfloat162float(b));
            }
        cudaMemcpy(t->wblob + wQ.off, h.data(), h.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
    }
    if (cudaMalloc(&t->maps_dev, maps.size() * sizeof(CUtensorMap)) != cudaSuccess) return bail("cudaMalloc maps");
    cudaMemcpy(t->maps_dev, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    if (cudaMalloc(&t->steps_dev, t->steps.size() * sizeof(TgStep)) != cudaSuccess) return bail("cudaMalloc steps");
    cudaMemcpy(t->steps_dev, t->steps.data(), t->steps.size() * sizeof(TgStep), cudaMemcpyHostToDevice);
    if (cudaMalloc(&t->wlayers_dev, wl.size() * sizeof(TgWLayer)) != cudaSuccess) return bail("cudaMalloc wlayers");
    cudaMemcpy(t->wlayers_dev, wl.data(), wl.size() * sizeof(TgWLayer), cudaMemcpyHostToDevice);
    if (cudaMalloc(&t->wtiles_dev, tiles.size() * sizeof(TgWTile)) != cudaSuccess) return bail("cudaMalloc wtiles");
    cudaMemcpy(t->wtiles_dev, tiles.data(), tiles.size() * sizeof(TgWTile), cudaMemcpyHostToDevice);
    if (cudaMalloc(&t->err_dev, sizeof(int)) != cudaSuccess) return bail("cudaMalloc err");
    cudaMemset(t->err_dev, 0, sizeof(int));
    t->num_sms = m->num_sms;
    if (getenv("LINNA_TG_DEBUG")) {
        if (cudaMalloc(&t->dbg, 64 * 8 * sizeof(long long)) != cudaSuccess) return bail("cudaMalloc dbg");
        cudaMemset(t->dbg, 0, 64 * 8 * sizeof(long long));
    }
    for (const TgStep &s : t->steps) t->max_tiles = std::max(t->max_tiles, (B_pad / TG_BM) * s.n_tiles);
    // (twice: the narrow step of a fused res-block launch parks its partial tiles behind the wide step's)
    if (cudaMalloc(&t->ws, (size_t)2 * t->max_tiles * 4 * TG_BM * TG_BN * sizeof(float)) != cudaSuccess) return bail("cudaMalloc split-K workspace");
    if (cudaMalloc(&t->sem, (size_t)2 * t->max_tiles * sizeof(int32_t)) != cudaSuccess) return bail("cudaMalloc split-K tickets");
    cudaMemset(t->sem, 0, (size_t)2 * t->max_tiles * sizeof(int32_t));
    t->m_tiles_max = B_pad / TG_BM;
    t->fuse_with_next.assign(t->steps.size(), 0), t->dep_seq.assign(t->steps.size(), 0);
    for (size_t si = 0; si + 1 < t->steps.size(); ++si) {
        const TgStep &a = t->steps[si], &b = t->steps[si + 1];
        // the narrow step's output planes are the A operand of the next step's second phase
        if (a.nphase == 1 && b.nphase == 2 && (a.epi == TG_ACT || a.epi == TG_BWD) && a.mapOut == b.mapA[1] && (int)si != t->i_lossq &&
            (int)si + 1 != t->i_lossq && !getenv("LINNA_TG_NO_FUSE"))
            t->fuse_with_next[si] = 1;
    }
    if (cudaMalloc(&t->dep_cnt, t->steps.size() * t->m_tiles_max * TG_DEP_NT * sizeof(uint32_t)) != cudaSuccess) return bail("cudaMalloc dep flags");
    cudaMemset(t->dep_cnt, 0, t->steps.size() * t->m_tiles_max * TG_DEP_NT * sizeof(uint32_t));
    if (cudaFuncSetAttribute(tg_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(tg_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_BYTES) != cudaSuccess)
        return bail("cudaFuncSetAttribute(tg kernels)");
    {   // initial packed planes from the host copies of the weights (flat state_dict order)
        size_t nflat = 0;
        for (const OpHost &op : m->ops) nflat += op.w.size() + op.b.size() + op.w2.size() + op.b2.size() + op.ws.size();
        std::vector<float> flat;
        flat.reserve(nflat);
        for (const OpHost &op : m->ops) {
            flat.insert(flat.end(), op.w.begin(), op.w.end()), flat.insert(flat.end(), op.b.begin(), op.b.end());
            if (op.kind == LINNA_OP_RES) {
                flat.insert(flat.end(), op.w2.begin(), op.w2.end()), flat.insert(flat.end(), op.b2.begin(), op.b2.end());
                if (op.has_ws) flat.insert(flat.end(), op.ws.begin(), op.ws.end());
            }
        }
        float *tmp = nullptr;
        if (cudaMalloc(&tmp, flat.size() * sizeof(float)) != cudaSuccess) return bail("cudaMalloc flat params");
        cudaMemcpy(tmp, flat.data(), flat.size() * sizeof(float), cudaMemcpyHostToDevice);
        cudaError_t e = tg_repack(t, tmp, 0);
        cudaDeviceSynchronize();
        cudaFree(tmp);
        if (e != cudaSuccess) return bail("initial weight pack failed");
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return bail("sync after tg_build");
    return t;
}

// packed planes <- flat parameter vector (device)
cudaError_t tg_repack(TgContext *t, const float *params, cudaStream_t stream)
{
    tg_repack_kernel<<<t->n_wtiles, 256, 0, stream>>>(params, t->wblob, t->wlayers_dev, t->wtiles_dev);
    return cudaGetLastError();
}

static TgArgs tg_args(const linna_model *m, TgContext *t, int64_t B)
{
    TgArgs a;
    memset(&a, 0, sizeof a);
    a.maps = t->maps_dev, a.act = t->act, a.masks = t->masks, a.B = (int)B, a.B_pad = t->B_pad;
    a.B_eff = (int)((B + TG_BM - 1) / TG_BM) * TG_BM;
    a.delta32 = t->delta32, a.dld = t->dld, a.chi_part = t->chi_part, a.chi_ld = t->chi_ld, a.c = m->consts;
    a.wlayers = t->wlayers_dev, a.wtiles = t->wtiles_dev, a.wblob = t->wblob, a.err = t->err_dev;
    a.ws = t->ws, a.sem = t->sem, a.splitk = 1;
    return a;
}

// Programmatic dependent launch: the kernel may start (barrier set-up, tensor-memory allocation) while its predecessor
// in the stream is still running; it executes griddepcontrol.wait before touching anything the predecessor wrote.
template <typename Kern>
static cudaError_t tg_launch_pdl(Kern kern, int grid, cudaStream_t stream, const TgArgs &a)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(TG_THREADS), cfg.dynamicSmemBytes = TG_SMEM_BYTES, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

// one layer step over `m_tiles` row tiles: up to 4 k-ranges per tile when the layer has few tiles and a long contraction
static int tg_splitk(const TgContext *t, const TgStep &s, int tiles)
{
    const int T = (s.K[0] + TG_KC - 1) / TG_KC + (s.nphase > 1 ? (s.K[1] + TG_KC - 1) / TG_KC : 0);
    static const int smax = getenv("LINNA_TG_MAX_SPLITK") ? std::max(1, atoi(getenv("LINNA_TG_MAX_SPLITK"))) : 4;
    // k-chunks per range: measured at C5 (scratch/splitk_sweep.sh) 0.261 ms per step with >= 4, 0.253 with >= 2, 0.257 with >= 1;
    // no split-K at all 0.277 -- the k-loops are a small part of a launch
    static const int tmin = getenv("LINNA_TG_SPLITK_MIN_CHUNKS") ? std::max(1, atoi(getenv("LINNA_TG_SPLITK_MIN_CHUNKS"))) : 2;
    int S = 1;
    while (S < smax && tiles * (2 * S) <= t->num_sms && T >= tmin * S) S *= 2;
    if (getenv("LINNA_TG_NO_SPLITK")) S = 1;
    return S;
}
static cudaError_t tg_launch_layer(TgContext *t, TgArgs &a, int si, int m_tiles, cudaStream_t stream)
{
    const TgStep &s = t->steps[si];
    const int tiles = m_tiles * s.n_tiles;
    a.step = t->steps_dev + si, a.splitk = tg_splitk(t, s, tiles);
    a.dbg = t->dbg, a.dbg_slot = si;
    a.step_dep = nullptr, a.dep_ctas = 0;
    return tg_launch_pdl(tg_layer_kernel, tiles * a.splitk, stream, a);
}
// Steps si (narrow) and si + 1 (two-phase) of a res-block in ONE launch when every CTA of both can be resident at once
// (one CTA per SM: the wide CTAs spin on the narrow ones, which have the lower block indices and are dispatched first).
// Returns false when the pair does not fit and has to be launched step by step.
static bool tg_launch_fused(TgContext *t, TgArgs &a, int si, int m_tiles, cudaStream_t stream, cudaError_t *ce)
{
    const TgStep &n = t->steps[si], &w = t->steps[si + 1];
    const int tiles_n = m_tiles * n.n_tiles, tiles_w = m_tiles * w.n_tiles;
    const int Sn = tg_splitk(t, n, tiles_n), Sw = tg_splitk(t, w, tiles_w);
    if (tiles_n * Sn + tiles_w * Sw > t->num_sms || tiles_n > t->max_tiles || n.n_tiles > TG_DEP_NT) return false;
    a.step = t->steps_dev + si + 1, a.splitk = Sw;
    a.dbg = t->dbg, a.dbg_slot = si + 1;
    a.step_dep = t->steps_dev + si, a.dep_ctas = tiles_n * Sn, a.dep_splitk = Sn;
    a.dep_ws_base = t->max_tiles * 4, a.dep_sem_base = t->max_tiles;
    a.dep_cnt = t->dep_cnt + (size_t)si * t->m_tiles_max * TG_DEP_NT, a.dep_ntiles = n.n_tiles;
    a.dep_target = ++t->dep_seq[si];                 // 1, 2, ...: the flags start at 0
    *ce = tg_launch_pdl(tg_layer_kernel, tiles_n * Sn + tiles_w * Sw, stream, a);
    a.step_dep = nullptr, a.dep_ctas = 0;
    return true;
}

// forward + loss head + quadratic form over B <= max_batch rows; `want_grad` also leaves d loss / d yhat for the
// backward pass.  Returns the number of kernels launched (negative: CUDA error).
// `side`: run the loss reduction on the context's loss stream (joined by the caller) instead of in line.
static int tg_forward_loss(const linna_model *m, TgContext *t, const float *X, const float *Y, const float *cmd,
                           int64_t B, int delta_kind, bool want_grad, float *rows, float *mean, cudaStream_t stream, bool side = false)
{
    TgArgs a = tg_args(m, t, B);
    a.target = Y, a.cmd = cmd, a.delta_kind = delta_kind, a.loss_inv_B = 1.0f / (float)B, a.want_grad = want_grad ? 1 : 0;
    const int m_tiles = (int)((B + TG_BM - 1) / TG_BM);
    int launches = 0;
    {
        const int total = t->B_pad * t->in_ld;
        tg_input_kernel<<<(total + 255) / 256, 256, 0, stream>>>(X, (int)B, t->B_pad, t->n_in, t->in_ld, m->consts, t->act + t->in_off);
        ++launches;
    }
    // delta_kind 1 (target vs data) does not need the network at all, but the chi^2 calls are not hot: same path
    for (int si = 0; si <= t->i_lossq; ++si) {
        cudaError_t ce = cudaSuccess;
        if (t->fuse_with_next[si] && tg_launch_fused(t, a, si, m_tiles, stream, &ce)) ++si;
        else ce = tg_launch_layer(t, a, si, m_tiles, stream);
        if (ce != cudaSuccess) return -1;
        ++launches;
    }
    cudaStream_t ls = stream;
    if (side) {
        if (cudaEventRecord(t->loss_ready, stream) != cudaSuccess || cudaStreamWaitEvent(t->loss_stream, t->loss_ready, 0) != cudaSuccess)
            return -1;
        ls = t->loss_stream;
    }
    tg_loss_kernel<<<1, 256, 0, ls>>>(t->chi_part, t->chi_ld, t->lossq_tiles, cmd, (int)B, rows, mean);
    if (side && cudaEventRecord(t->loss_done, ls) != cudaSuccess) return -1;
    ++launches;
    return cudaGetLastError() == cudaSuccess ? launches : -1;
}

// `leave_unjoined` is reserved (a caller that reduces the buckets behind their own streams); every caller passes false and
// the bucket streams are joined with `stream` before the step returns.
int tg_train_step(const linna_model *m, TgContext *t, const float *X, const float *Y, const float *cmd, int64_t B, const AdamArgs &ad,
                  float *loss_rows, float *loss_mean, cudaStream_t stream, bool leave_unjoined)
{
    int launches = tg_forward_loss(m, t, X, Y, cmd, B, 0, true, loss_rows, loss_mean, stream, true);
    if (launches < 0) return -1;
    TgArgs a = tg_args(m, t, B);
    const int m_tiles = (int)((B + TG_BM - 1) / TG_BM);
    TgArgs w = a;
    w.adam = ad, w.dbg = t->dbg;
    auto launch_bucket = [&](int b, cudaStream_t st) -> bool {
        if (t->bucket_ntiles[b] <= 0) return true;
        TgArgs wb = w;
        wb.wtiles = t->wtiles_dev + t->bucket_tile0[b];
        if (tg_launch_pdl(tg_wgrad_kernel, t->bucket_ntiles[b], st, wb) != cudaSuccess) return false;
        ++launches;
        return true;
    };
    auto side_buckets_after = [&](int si) -> bool {   // buckets whose layers the chain has just passed: onto their own streams
        for (int b = 0; b + 1 < t->n_buckets; ++b)
            if (t->bucket_ready_step[b] == si) {
                if (cudaEventRecord(t->bucket_ready[b], stream) != cudaSuccess) return false;
                if (cudaStreamWaitEvent(t->bucket_stream[b], t->bucket_ready[b], 0) != cudaSuccess) return false;
                if (!launch_bucket(b, t->bucket_stream[b])) return false;
            }
        return true;
    };
    if (!side_buckets_after(t->i_lossq)) return -1;
    for (int si = t->i_lossq + 1; si < t->n_steps; ++si) {
        cudaError_t ce = cudaSuccess;
        if (t->fuse_with_next[si] && tg_launch_fused(t, a, si, m_tiles, stream, &ce)) {
            if (ce != cudaSuccess || !side_buckets_after(si)) return -1;
            ++si;
        } else
            ce = tg_launch_layer(t, a, si, m_tiles, stream);
        if (ce != cudaSuccess) return -1;
        ++launches;
        if (!side_buckets_after(si)) return -1;
    }
    if (!launch_bucket(t->n_buckets - 1, stream)) return -1;
    if (cudaStreamWaitEvent(stream, t->loss_done, 0) != cudaSuccess) return -1;   // the loss of this step is part of the step
    t->buckets_unjoined = leave_unjoined && t->n_buckets > 1;
    if (!t->buckets_unjoined)
        for (int b = 0; b + 1 < t->n_buckets; ++b) {
            if (cudaEventRecord(t->bucket_done[b], t->bucket_stream[b]) != cudaSuccess) return -1;
            if (cudaStreamWaitEvent(stream, t->bucket_done[b], 0) != cudaSuccess) return -1;
        }
    return cudaGetLastError() == cudaSuccess ? launches : -1;
}

int tg_chisq(const linna_model *m, TgContext *t, const float *X, const float *Y, int64_t n, int kind, float *chi2,
             cudaStream_t stream)
{
    int launches = 0;
    for (int64_t done = 0; done < n; done += m->max_batch) {
        const int64_t B = std::min<int64_t>(m->max_batch, n - done);
        const int l = tg_forward_loss(m, t, X + done * m->n_in, Y + done * m->n_out, nullptr, B, kind, false, chi2 + done, nullptr,
                                      stream);
        if (l < 0) return -1;
        launches += l;
    }
    return launches;
}

// LINNA_TG_DEBUG: cycle stamps of CTA 0 of every layer launch of the last step: [step][8]
int tg_debug_read(TgContext *t, long long *out, int max_steps)
{
    if (!t || !t->dbg) return 0;
    const int n = std::min(max_steps, 41);   // slots 0 .. n_steps-1: layer launches; slot 40: the weight-gradient launch
    cudaDeviceSynchronize();
    cudaMemcpy(out, t->dbg, (size_t)n * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
    return n;
}

int tg_check(TgContext *t)
{
    int e = 0;
    cudaMemcpy(&e, t->err_dev, sizeof(int), cudaMemcpyDeviceToHost);
    return e;
}

}  // namespace linna
