// Tensor-core emulator-likelihood kernel (sm_100a: tcgen05 + TMEM + TMA), lnP and lnP+gradient.
//
// Same step program as the FFMA kernel (linna/nn.py:45-56, :110-133; linna/util.py:953-955, :990-1021), but
// every GEMM  D[256 walkers][N] = A[256][K] . B[N][K]^T  runs on the 5th-generation tensor cores:
//
//   * split-fp16 product.  Every fp32 operand x is stored as two halves, hi = fp16(x), lo = fp16(x - hi)
//     (22 significant bits; weights are pre-scaled by a power of two per step so that they sit in the
//     middle of the fp16 range), and  D += A_lo.B_hi + A_hi.B_lo + A_hi.B_hi  with exact fp16 x fp16
//     products and fp32 accumulation in tensor memory.  kind::f16 runs at twice the TF32 rate and the
//     operands are half as wide, so the three passes cost what 1.5 TF32 passes would.
//   * two-level accumulation.  The tensor core truncates its fp32 accumulator on every tcgen05.mma
//     (measured: a toward-zero bias of ~0.5 ulp per instruction), so every `seg_kc` k-chunks (default 6 = 192
//     values of K = 36 instructions; measured max relative lnP error over the goldens 2.7e-7 at 6, 3.2e-7 at 4,
//     3.7e-7 at 8, all at the level of the reference's own float32, 2.0e-7; mean relative error 6e-8 at 6) the
//     partial tile is drained from tensor memory and added with round-to-nearest into fp32 REGISTER
//     accumulators by the epilogue warps, while the MMA warp already fills another TMEM buffer (tensor memory
//     is a ring of four 128-column entries, TfAccRing: a chunk of more than 128 columns takes an aligned pair).
//   * CTA pairs.  One cluster = two CTAs = 2 x 128 walkers; rank 0 issues every MMA for the pair
//     (cta_group::2, M = 256, N <= 256).  Each CTA supplies its own 128 activation rows and HALF of the weight
//     tile and receives its own 128 x N accumulator rows: half the L2 weight traffic per walker, and the MMA
//     is off the shared-memory bandwidth limit a 128 x 128 single-CTA instruction sits on.
//   * operands are K-major 128-byte-swizzled tiles (4 stages x 36 KB) filled by TMA (cp.async.bulk.tensor,
//     completion on the leader's mbarrier) from the packed weights and from the row-major activation arena of
//     this CTA.  Both keep the hi and lo halves of a k-chunk of 32 values side by side ([32 hi | 32 lo] = 128
//     contiguous bytes per row), so that one TMA box row is one full cache line and a stage is two TMA
//     instructions (A tile, B half tile) rather than four of half-line rows; the MMA descriptors pick the hi or
//     lo half of the swizzle row by a 64-byte offset of the start address.  Layer outputs go back to the arena
//     through staging boxes of the same format and TMA stores.  The store warps publish, per 64-column box,
//     how far a layer's output is visible, and the producer fetches a k-chunk as soon as ITS columns are
//     there, so the next layer starts on the first columns of an activation while the epilogue is still
//     writing the last ones.  A narrow step (N <= 32) packs two k-chunks into one stage (TF_STAGE_BYTES).
//   * every cluster interleaves two walker pairs ("slots") layer by layer, so that the layer-to-layer
//     dependency bubble of one pair is filled with the other pair's MMAs.
//   * warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA) + TMEM allocator -- both walk the program
//     warp-uniformly and issue from one elected lane --, warps 2 / 3
//     = TMA store issuers of column group 0 / 1, warps 4-7 / 8-11 = epilogue of column group 0 / 1 (columns
//     [0,128) / [128,256) of every 256-column chunk): an epilogue thread owns one walker (TMEM lane) and 128
//     columns, so the chi^2 reduction, the relu masks of the backward pass and the final Jacobian need no
//     cross-thread traffic beyond one exchange of the two groups' chi^2 partials.  setmaxnreg moves registers
//     from the four service warps to the epilogue warps, whose 128 fp32 accumulators per thread are the
//     second accumulation level.
//
// Two programs: LNP (forward, chi^2) and GRAD (forward with saved relu masks, backward-data through every
// layer with the transposed weights at a per-walker power-of-two scale, prior-map Jacobian in the last
// epilogue).
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "linna_host.hpp"

namespace linna {

constexpr int TF_M = 128;        // walkers per tile
constexpr int TF_NC = 256;       // accumulator columns per chunk (two epilogue groups of 128)
constexpr int TF_KC = 32;        // k-chunk: 32 values of K = [32 hi halves | 32 lo halves] = one 128-byte swizzle row
constexpr int TF_STAGES = 4;
constexpr int TF_TILE_BYTES = TF_M * 2 * TF_KC * 2;              // 16 KB operand tile (hi and lo of one k-chunk, interleaved)
constexpr int TF_NARROW_B = 2048;                                // B-half tile of a narrow step: 16 rows x 128 bytes
// One operand stage: [A tile 16 KB][B-half tile 16 KB][4 KB].  A "narrow" step (N <= 32: the bottleneck layer of a
// res-block, the last backward step) needs 16 weight rows per CTA and k-chunk, and its k-loop is pure latency (one
// TMA round trip per four stages in flight, next to no MMA work): it packs TWO k-chunks into a stage -- the second
// A tile in the place of the B-half tile, the two 2 KB weight tiles in the last 4 KB -- so that twice as much of
// the k-loop is in flight.
constexpr int TF_STAGE_BYTES = 2 * TF_TILE_BYTES + 2 * TF_NARROW_B;   // 36 KB per CTA
constexpr int TF_BOX_BYTES = 128 * 64 * 2;                       // staging box: 128 rows x 64 halves = 16 KB = one k-chunk, hi | lo
constexpr int TF_STG_BYTES = 2 * TF_BOX_BYTES;                   // two k-chunks (64 activation columns) per column group
constexpr int TF_SMEM_BYTES = TF_STAGES * TF_STAGE_BYTES + 2 * TF_STG_BYTES + 1024;
constexpr int TF_THREADS = 384;  // TMA, MMA, 2 store warps + 2 x 4 epilogue warps
constexpr int TF_MAX_STEPS = 48;
// LINNA_TC_DEBUG counters per CTA: 16 role counters, then per program step the cycles of the MMA warp [16, 40), of
// epilogue group 0 [40, 64), and of that group's waits for accumulators [64, 88) and chunk epilogues [88, 112)
constexpr int TF_DBG_STRIDE = 128;

enum TfEpi : int32_t { TF_ACT = 0, TF_HEAD = 1, TF_CHI2 = 2, TF_BWD = 3, TF_GRADOUT = 4, TF_PREDICT = 5 };
enum TfFlags : int32_t { TFF_RELU = 1, TFF_SAVE_MASK = 2, TFF_APPLY_MASK = 4, TFF_TRI = 8,
                         TFF_LAST_USE0 = 16, TFF_LAST_USE1 = 32,     // phase 0 / 1 is the last reader of its activation slot
                         TFF_ROWSCALE = 64 };   // first backward step: apply the per-walker power-of-two gradient scale
enum TfVariant : int32_t { TFV_ACT = 0, TFV_ACT_SAVE, TFV_CHI2, TFV_CHI2_STORE, TFV_BWD, TFV_HEAD, TFV_HEAD_EXP, TFV_GRADOUT, TFV_PREDICT };

struct TfStep {
    int32_t nphase;
    int32_t src[2];       // arena column (in halves: 2 x the activation column) of this phase's A operand
    int32_t K[2];
    int32_t mapB[2];      // tensor-map index of the weight operand, box of 128 rows (full 256-column chunks)
    int32_t mapBt[2];     // the same matrix with a box of `tail_rows` rows, for the last chunk of the step
    int32_t tail_rows;    // weight rows each CTA supplies to the last chunk: round32(N - n0_last) / 2
    int32_t src_pub[2][2];   // [phase][group]: 64-column boxes the group published (per tile pass) before the producer of src
    int32_t N;
    int32_t dst;          // arena column (in halves) of the output, -1: none
    int32_t dst_pad;      // output width rounded up to 64 (pad columns are written as zeros)
    int32_t epi, flags;
    int32_t mask_word;    // first 32-bit word of this layer's relu bits inside a mask row
    float inv_scale;      // 2^-s of the weight pre-scale (and any uniform factor folded in)
    float clampv;         // lower clamp of the output: 0 (relu) or -inf
    int32_t variant;      // TfVariant: which specialised chunk epilogue runs
    const float *bias;    // [dst_pad] effective bias (zero-padded), or nullptr
    const float *vscale;  // [dst_pad] per-column scale (HEAD), or nullptr
    const float *sub;     // [dst_pad] data/sigma (HEAD with ypositive), or nullptr
};

struct TfProgram {
    int32_t n_steps;
    int32_t total_pub[2];  // 64-column boxes each column group publishes per tile pass (prologue included)
    int32_t in_col;        // arena column (in halves) of xhat
    int32_t pad0_;
    int32_t seg_kc;        // k-chunks accumulated in tensor memory between two drains
    int32_t mask_words;    // 32-bit words per mask row
    int32_t pad_[1];
    TfStep steps[TF_MAX_STEPS];
};

struct TfArgs {
    const TfProgram *prog;
    const CUtensorMap *maps;  // [0] arena load, [1] arena store, then one per weight operand
    Consts c;
    const float *in;
    float *lnp;
    float *grad;
    uint32_t *masks;
    int64_t n;
    int32_t slots;            // walker pairs interleaved per cluster: 2, or 1 when the batch cannot fill the GPU twice
    int *err;
    long long *dbg;           // optional [grid][8] cycle counters (LINNA_TC_DEBUG): where the service warps wait
    // rows whose result came out NaN (an activation beyond the fp16 range, |x| > 65504, turns into inf - inf): their
    // indices are appended here and the FP32 FFMA kernel recomputes exactly these rows right after this launch
    int32_t *fix_count;
    int32_t *fix_rows;
    int32_t fix_cap;
    // Predictor.predict (PREDICT program): physical parameters in (no prior map), the selected vector out
    float *out_vec;           // [n][n_out]
    int32_t out_kind;         // LINNA_OUT_*
    int32_t input_theta;
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
__device__ __noinline__ void tf_die(int *err, int code)
{
    if (err) atomicExch(err, code);
    __threadfence_system();
    __trap();
}
// Bounded wait: a protocol bug must trap (and report) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int *err, int code)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 2000000000LL) tf_die(err, code);
    }
}
// DBG (LINNA_TC_DEBUG) instantiation only: the same wait with its duration added to a cycle counter.  The
// production kernel does not read the clock on its service warps' critical path.
template <bool DBG>
__device__ __forceinline__ void mbar_wait_timed(uint64_t *bar, uint32_t parity, int *err, int code, long long &acc)
{
    if (!DBG) {
        mbar_wait(bar, parity, err, code);
        return;
    }
    const long long t0 = clock64();   // try_wait itself suspends the thread for a while: time it from the start
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 2000000000LL) tf_die(err, code);
    }
    acc += clock64() - t0;
}
// One lane of a converged warp.  Issuing tcgen05 / TMA instructions from `if (elect_one())` inside warp-uniform
// control flow keeps the service loops free of the per-instruction serialisation loops that the compiler has
// to wrap around uniform-datapath instructions in divergent code (`if (lane == 0)`).
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// CTA-pair load: the bytes land in this CTA's shared memory, the completion is signalled on the LEADER's
// barrier (peer bit of the address cleared), as cta_group::2 MMAs consume both CTAs' tiles together.
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    // plain arrive: a .release.cluster here costs MEMBAR.ALL.GPU + ERRBAR per call; the tensor-memory handoff is
    // ordered by tcgen05.fence::before_thread_sync / after_thread_sync on the two sides
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const void *smem_src, const CUtensorMap *map, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :
                 : "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile, 128-byte swizzle: rows are 128 B apart, 8-row groups 1024 B apart.  A row holds the hi
// halves of its k-chunk in bytes [0, 64) and the lo halves in [64, 128): a K = 16 slice of either is a 32-byte
// step of the start address inside the swizzle row.
__device__ __forceinline__ uint64_t make_sdesc128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);  // start address, 16-byte units
    d |= (uint64_t)1 << 16;                  // leading byte offset (unused with swizzle)
    d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
    return d;
}
// The six MMAs of one k-chunk (two K = 16 slices x three split-fp16 passes) from the descriptors of the A and
// B-half tiles of a stage: a slice is +32 bytes, the lo half +64 bytes (descriptor address units are 16 bytes;
// the tiles are 1024-byte aligned, so the additions never carry out of the address field).
__device__ __forceinline__ void umma_f16_pair_x6(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, t;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, %4, %4;\n\t"
        "add.u64 a1, %1, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 a3, %1, 6;\n\t"
        "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a2, %2, %3, p;\n\t"     // A_lo . B_hi, K slice 0
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, b2, %3, t;\n\t"     // A_hi . B_lo
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, t;\n\t"     // A_hi . B_hi
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b1, %3, t;\n\t"     // K slice 1
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b3, %3, t;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, t;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::f16 with fp16 inputs, fp32 accumulate, A and B K-major, M = 256 over the CTA pair
__device__ __forceinline__ uint32_t make_idesc_f16(int n)
{
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)((2 * TF_M) >> 4) << 24);
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
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
// x = hi + lo with hi = fp16(x), lo = fp16(x - hi); two values per call (packed half2 words)
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo)
{
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
// max that PROPAGATES NaN (fmaxf returns the other operand): a relu must not turn the NaN of an fp16 overflow into 0
__device__ __forceinline__ float max_nan(float a, float b)
{
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float tf_prior_map(float u, int kind, float scale, float shift)
{
    float t = u;
    if (kind == LINNA_PRIOR_FLAT) t = 0.5f * (1.0f + erff(u / 1.41421356237309515f));  // gauss2unif, util.py:300
    return t * scale + shift;
}
// k-chunks the next operand stage of a narrow step takes (producer and MMA issuer decide alike): two when both
// belong to the same 64-column publication box, the same phase and the same accumulation segment.
__device__ __forceinline__ int tf_stage_take(bool narrow, int kc, int nk, int in_seg, int seg_kc)
{
    return (narrow && !(kc & 1) && kc + 1 < nk && in_seg + 2 <= seg_kc) ? 2 : 1;
}
// Tensor memory is a ring of four 128-column accumulator entries.  A segment of a chunk of up to 128 columns takes
// one entry, a wider one an even-aligned pair -- so the narrow steps, whose k-loops are chains of short segments, have
// four accumulator buffers in flight instead of two.  The MMA issuer and the epilogue warps walk the ring alike:
// `r` is the next entry, bit e of `par` the parity of the uses of entry e so far.
struct TfAccRing {
    uint32_t r = 0, par = 0;
    // entry of the next segment (`wide`: two entries)
    __device__ __forceinline__ uint32_t open(bool wide)
    {
        if (wide && (r & 1u)) r = (r + 1u) & 3u;
        return r;
    }
    __device__ __forceinline__ uint32_t parity(uint32_t e) const { return (par >> e) & 1u; }
    __device__ __forceinline__ void close(uint32_t e, bool wide)
    {
        par ^= wide ? (3u << e) : (1u << e);
        r = (e + (wide ? 2u : 1u)) & 3u;
    }
};
// stages (k-chunks over all phases) of output chunk n0 of a step
__device__ __forceinline__ int tf_chunk_stages(const TfStep &st, int n0)
{
    int s = 0;
    const int k0 = (st.flags & TFF_TRI) ? n0 / TF_KC : 0;
    for (int p = 0; p < st.nphase; ++p) s += (st.K[p] + TF_KC - 1) / TF_KC - k0;
    return s;
}

// ------------------------------------------------------------------------------------------ chunk epilogue
struct TfEpiCtx {
    uint8_t *my_x, *my_y;       // this thread's 128-byte rows of the two staging boxes (k-chunk 0 / 1 of 64 columns)
    uint32_t sx, sw16;          // shared-space address of my_x (my_y = + TF_BOX_BYTES); swizzle phase << 4
    uint64_t *sfree, *sfull;
    uint32_t *mask_row;
    int *err;
    double chi;
    uint32_t sidx;              // staging boxes handed to the store warp so far
    int sw;                     // 128-byte swizzle phase of this row
    long long t_sfree;          // cycles spent waiting for the staging box (LINNA_TC_DEBUG)
    float row_scale;            // 2^-k of this walker's backward pass (k from its chi^2), 1 outside it
    bool timing;                // LINNA_TC_DEBUG instantiation
};

// 128 columns of one output chunk (the columns [c0, c0+128) of the layer, owned by one epilogue group):
// v = acc*scale + bias (clamped below for relu), optional exp / mask / chi^2, split into fp16 hi/lo, written to
// the swizzled staging box that the store warp sends to the arena.  All flags are compile-time so that each
// kind of step runs ~6 instructions per element out of a few KB of code.  Pad columns (>= N) come out as
// exact zeros: their accumulators are zero (TMA zero-fills the missing weight rows) and the bias / scale
// vectors are zero-padded.
template <bool BIAS, bool VSCALE, bool EXPY, bool SAVE, bool APPLY, bool CHI, bool STORE>
__device__ __forceinline__ void tf_chunk_epilogue(const float (&racc)[128], const TfStep &st, int c0, int mword,
                                                  const float *bias_s, TfEpiCtx &x)
{
    uint32_t mw[4] = {0u, 0u, 0u, 0u};
    if (APPLY) {
        if (st.flags & TFF_APPLY_MASK) {
            const uint4 m4 = *reinterpret_cast<const uint4 *>(x.mask_row + mword);
            mw[0] = m4.x, mw[1] = m4.y, mw[2] = m4.z, mw[3] = m4.w;
        } else {
            mw[0] = mw[1] = mw[2] = mw[3] = 0xffffffffu;
        }
    }
    const float clampv = st.clampv;
    const float inv_scale = (APPLY && (st.flags & TFF_ROWSCALE)) ? st.inv_scale * x.row_scale : st.inv_scale;
    float chi_f = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int col0 = c0 + 64 * h;
        const bool store = STORE && col0 < st.dst_pad;
        if (!store && !(CHI && col0 < st.N)) continue;
        if (store) {   // staging box free again
            if (x.timing) mbar_wait_timed<true>(x.sfree, (x.sidx & 1) ^ 1, x.err, 7, x.t_sfree);
            else mbar_wait(x.sfree, (x.sidx & 1) ^ 1, x.err, 7);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int cb = col0 + 8 * j;
            float b[8], s[8], v[8];
            if (BIAS) {   // this chunk's 128 bias values were staged in shared memory while the accumulators drained
                const float4 b0 = *reinterpret_cast<const float4 *>(bias_s + 64 * h + 8 * j);
                const float4 b1 = *reinterpret_cast<const float4 *>(bias_s + 64 * h + 8 * j + 4);
                b[0] = b0.x, b[1] = b0.y, b[2] = b0.z, b[3] = b0.w, b[4] = b1.x, b[5] = b1.y, b[6] = b1.z, b[7] = b1.w;
            }
            if (VSCALE) {
                const float4 s0 = __ldg(reinterpret_cast<const float4 *>(st.vscale + cb));
                const float4 s1 = __ldg(reinterpret_cast<const float4 *>(st.vscale + cb + 4));
                s[0] = s0.x, s[1] = s0.y, s[2] = s0.z, s[3] = s0.w, s[4] = s1.x, s[5] = s1.y, s[6] = s1.z, s[7] = s1.w;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int li = 64 * h + 8 * j + e;   // column inside this group's 128
                float y = fmaf(racc[li], VSCALE ? s[e] : inv_scale, BIAS ? b[e] : 0.f);
                if (EXPY) y = cb + e < st.N ? expf(y) - __ldg(st.sub + cb + e) : 0.f;
                y = max_nan(y, clampv);
                if (APPLY) y = ((mw[li >> 5] >> (li & 31)) & 1u) ? y : 0.f;
                if (SAVE) mw[li >> 5] |= (y > 0.f ? 1u : 0u) << (li & 31);
                if (CHI) chi_f = fmaf(y, y, chi_f);
                v[e] = y;
            }
            if (store) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int e = 0; e < 8; e += 2) split2(v[e], v[e + 1], hw[e >> 1], lw[e >> 1]);
                // columns [8j, 8j+8) of the box: 16-byte unit j&3 of the hi half, 4 + (j&3) of the lo half
                // (32-bit shared-space stores: no generic-address arithmetic in the hot loop)
                const uint32_t bx = x.sx + ((j & 4) ? (uint32_t)TF_BOX_BYTES : 0u);
                st_shared_v4(bx + ((uint32_t)((j & 3) << 4) ^ x.sw16), hw[0], hw[1], hw[2], hw[3]);
                st_shared_v4(bx + ((uint32_t)((4 + (j & 3)) << 4) ^ x.sw16), lw[0], lw[1], lw[2], lw[3]);
            }
        }
        if (store) {
            fence_async_smem();
            mbar_arrive(x.sfull);
            ++x.sidx;
        }
    }
    if (SAVE) *reinterpret_cast<uint4 *>(x.mask_row + mword) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
    if (CHI) x.chi += (double)chi_f;
}

// ------------------------------------------------------------------------------------------ kernel
// One cluster = one CTA pair = 2 x 128 walkers.  CTA rank 0 issues every tcgen05.mma for the pair
// (cta_group::2, M = 256): each CTA supplies its own 128 activation rows and HALF of the weight tile, and
// receives its own 128 x N accumulator rows in its own tensor memory.
template <bool DBG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TF_THREADS, 1) tc_f16_kernel(const TfArgs args)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[TF_STAGES], empty_bar[TF_STAGES], pfull_bar[4], pempty_bar[4];
    __shared__ __align__(8) uint64_t sfull_bar[2], sfree_bar[2];   // staging buffer of column group 0 / 1: written / read out
    __shared__ __align__(8) double chi_s[TF_M];
    __shared__ __align__(8) double chi_x[2][TF_M];       // chi^2 partials of the two column groups (backward scale)
    __shared__ __align__(16) float bias_stage[2][2][128];   // [chunk parity][column group][column]
    __shared__ uint32_t tmem_slot;
    __shared__ uint32_t ready_cnt[2][2];   // [slot][column group]
    __shared__ TfStep s_steps[TF_MAX_STEPS];

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *stg_all = smem + TF_STAGES * TF_STAGE_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const bool leader = cta_rank == 0;
    const TfProgram *prog = args.prog;
    const int n_steps = prog->n_steps;
    const int seg_kc = prog->seg_kc;
    const Consts &c = args.c;
    const CUtensorMap *maps = args.maps;

    for (int i = tid; i < n_steps * (int)(sizeof(TfStep) / 4); i += TF_THREADS)
        reinterpret_cast<uint32_t *>(s_steps)[i] = reinterpret_cast<const uint32_t *>(prog->steps)[i];
    if (tid == 0) {
        for (int s = 0; s < TF_STAGES; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
        for (int b = 0; b < 4; ++b) mbar_init(&pfull_bar[b], 1), mbar_init(&pempty_bar[b], 16);   // 8 warps x 2 CTAs
        for (int b = 0; b < 2; ++b) mbar_init(&sfull_bar[b], 128), mbar_init(&sfree_bar[b], 1);
        ready_cnt[0][0] = ready_cnt[0][1] = ready_cnt[1][0] = ready_cnt[1][1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // the peer's barriers are initialised before anything is signalled across the pair
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    const int64_t npairs = (args.n + 2 * TF_M - 1) / (2 * TF_M);
    // Every cluster works on TWO walker pairs ("slots") at a time, interleaved step by step: while the last
    // chunk of slot 0's layer goes through epilogue, store and publication, the tensor core already runs the same
    // layer for slot 1, so the layer-to-layer dependency never idles the MMA pipe.
    // Cluster c owns the walker pairs c, c + C, c + 2C, ... and walks through them `slots` at a time (the last
    // visit takes whatever is left), so every cluster gets floor or ceil of pairs / clusters whatever `slots` is.
    const int64_t n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
    const int64_t my_pairs = cluster_id < npairs ? (npairs - cluster_id + n_clusters - 1) / n_clusters : 0;
    const int64_t pair0 = 0, pair_step = args.slots;   // `pair` below counts this cluster's own pairs
    const int arena_row0 = blockIdx.x * 2 * TF_M;   // this CTA's rows of the activation arena (slot 0, then slot 1)

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(56));
    if (warp == 0) {
        // =============================== TMA producer (both CTAs) ===============================
        // The whole warp walks the program (warp-uniform control flow, every lane waits on the barriers); one
        // elected lane issues the TMA instructions.
        {
            int stage = 0;
            uint32_t ph = 0, pubA = 0, pubB = 0;
            uint32_t seen00 = 0, seen01 = 0, seen10 = 0, seen11 = 0;   // [slot][group]: publications seen so far (registers, not a local array)
            long long w_empty = 0, w_ready = 0;
            const long long t_begin = DBG ? clock64() : 0;
            for (int64_t pair = pair0; pair < my_pairs; pair += pair_step, pubA += prog->total_pub[0], pubB += prog->total_pub[1]) {
                const int nslots = (args.slots == 2 && pair + 1 < my_pairs) ? 2 : 1;
                for (int si = 0; si < n_steps; ++si) {
                    const TfStep &st = s_steps[si];
                    const int st_N = st.N, st_flags = st.flags, st_nphase = st.nphase;
                    for (int slot = 0; slot < nslots; ++slot)
                    for (int n0 = 0; n0 < st_N; n0 += TF_NC) {
                        // this CTA's half of the weight rows; the last chunk of a step uses a tensor map whose box is just
                        // that tall (a bottleneck layer of 16 columns loads 8 rows per CTA, not 128 rows of zero fill)
                        const bool tail = n0 + TF_NC >= st_N;
                        const int brows = tail ? st.tail_rows : TF_M;
                        const int nb = n0 + (int)cta_rank * brows;
                        const uint32_t stage_tx = 2u * (uint32_t)(TF_TILE_BYTES + brows * 4 * TF_KC);   // both CTAs' bytes
                        const int k0 = (st_flags & TFF_TRI) ? n0 / TF_KC : 0;   // L^T: B[n][k] = 0 for k < n
                        const int arow = arena_row0 + slot * TF_M;
                        const bool narrow = st_N <= 32;   // one chunk, 16 weight rows per CTA
                        int in_seg = 0;
                        for (int p = 0; p < st_nphase; ++p) {
                            const int nk = (st.K[p] + TF_KC - 1) / TF_KC;
                            const CUtensorMap *mb = maps + (tail ? st.mapBt[p] : st.mapB[p]);
                            const int src_col = st.src[p];
                            const uint32_t base_pub0 = pubA + (uint32_t)st.src_pub[p][0] + 1u, base_pub1 = pubB + (uint32_t)st.src_pub[p][1] + 1u;
                            for (int kc = k0, take; kc < nk; kc += take) {
                                take = tf_stage_take(narrow, kc, nk, in_seg, seg_kc);
                                in_seg += take;
                                if (in_seg >= seg_kc) in_seg = 0;
                                mbar_wait_timed<DBG>(&empty_bar[stage], ph ^ 1, args.err, 1, w_empty);
                                uint8_t *sb = smem + stage * TF_STAGE_BYTES;
                                uint64_t *fb = &full_bar[stage];
                                // the activations this stage reads must be visible: their producer chunk is published per
                                // 64-column box of its column group
                                const int col = kc * TF_KC;
                                const int grp = (col >> 7) & 1;
                                const uint32_t box = (uint32_t)((col >> 8) * 2 + ((col & 127) >> 6));
                                const uint32_t need = (grp ? base_pub1 : base_pub0) + box;
                                uint32_t seen_v = slot ? (grp ? seen11 : seen10) : (grp ? seen01 : seen00);
                                const int ca = src_col + 2 * col;
                                const bool wait_pub = seen_v < need;
                                if (elect_one()) {   // the weights, and in the usual case (activations published) the whole stage
                                    if (leader) mbar_expect_tx(fb, take == 2 ? 2 * stage_tx : stage_tx);
                                    if (narrow) {
                                        tma_load_2d_pair(sb + 2 * TF_TILE_BYTES, mb, fb, kc * 2 * TF_KC, nb);
                                        if (take == 2) tma_load_2d_pair(sb + 2 * TF_TILE_BYTES + TF_NARROW_B, mb, fb, (kc + 1) * 2 * TF_KC, nb);
                                    } else {
                                        tma_load_2d_pair(sb + TF_TILE_BYTES, mb, fb, kc * 2 * TF_KC, nb);
                                    }
                                    if (!wait_pub) {
                                        tma_load_2d_pair(sb, maps, fb, ca, arow);
                                        if (take == 2) tma_load_2d_pair(sb + TF_TILE_BYTES, maps, fb, ca + 2 * TF_KC, arow);
                                    }
                                }
                                if (wait_pub) {
                                    const long long t0 = clock64();
                                    while ((seen_v = ld_acquire_u32(&ready_cnt[slot][grp])) < need) {
                                        __nanosleep(64);
                                        if (clock64() - t0 > 2000000000LL) tf_die(args.err, 2);
                                    }
                                    if (DBG) w_ready += clock64() - t0;
                                    seen_v = __shfl_sync(0xffffffffu, seen_v, 0);   // every lane made its own acquire; keep the count warp-uniform
                                    if (slot) { if (grp) seen11 = seen_v; else seen10 = seen_v; }
                                    else { if (grp) seen01 = seen_v; else seen00 = seen_v; }
                                    fence_async_all();
                                    if (elect_one()) {
                                        tma_load_2d_pair(sb, maps, fb, ca, arow);
                                        if (take == 2) tma_load_2d_pair(sb + TF_TILE_BYTES, maps, fb, ca + 2 * TF_KC, arow);
                                    }
                                }
                                if (++stage == TF_STAGES) stage = 0, ph ^= 1;
                            }
                        }
                    }
                }
            }
            if (DBG && lane == 0) {
                long long *d = args.dbg + (size_t)blockIdx.x * TF_DBG_STRIDE;
                d[0] = clock64() - t_begin, d[1] = w_empty, d[2] = w_ready;
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer (leader CTA only) ===============================
        // Warp-uniform loop as above; the elected lane issues the tcgen05.mma / tcgen05.commit instructions.
        if (leader) {
            int stage = 0;
            uint32_t ph = 0, acc_e = 0;
            TfAccRing ring;
            long long w_full = 0, w_pempty = 0, w_full_head = 0;
            const long long t_begin = DBG ? clock64() : 0;
            const uint64_t desc0 = make_sdesc128(smem_u32(smem));
            for (int64_t pair = pair0; pair < my_pairs; pair += pair_step) {
                const int nslots = (args.slots == 2 && pair + 1 < my_pairs) ? 2 : 1;
                for (int si = 0; si < n_steps; ++si) {
                    const TfStep &st = s_steps[si];
                    const int st_N = st.N, st_flags = st.flags, st_nphase = st.nphase;
                    const long long t_step = DBG ? clock64() : 0;
                    int nk_p[2];
                    nk_p[0] = (st.K[0] + TF_KC - 1) / TF_KC, nk_p[1] = st_nphase > 1 ? (st.K[1] + TF_KC - 1) / TF_KC : 0;
                    for (int slot = 0; slot < nslots; ++slot)
                    for (int n0 = 0; n0 < st_N; n0 += TF_NC) {
                        const int nvalid = st_N - n0 < TF_NC ? st_N - n0 : TF_NC;
                        // N rounded up to the 32 columns one tcgen05.ld drains: the extra rows of B are TMA zero fill
                        const uint32_t idesc = make_idesc_f16((nvalid + 31) & ~31);
                        const int k0 = (st_flags & TFF_TRI) ? n0 / TF_KC : 0;
                        const int total = nk_p[0] - k0 + (st_nphase > 1 ? nk_p[1] - k0 : 0);
                        const bool narrow = st_N <= 32;
                        const bool wide = ((nvalid + 31) & ~31) > 128;   // accumulator columns of this chunk: one ring entry or two
                        int in_seg = 0, done = 0, nstage = 0;
                        uint32_t dcol = 0;
                        for (int p = 0; p < st_nphase; ++p) {
                            const int nk = nk_p[p];
                            for (int kc = k0, take; kc < nk; kc += take, ++nstage) {
                                take = tf_stage_take(narrow, kc, nk, in_seg, seg_kc);
                                if (in_seg == 0) {   // open a fresh accumulator buffer (in both CTAs)
                                    acc_e = ring.open(wide);
                                    mbar_wait_timed<DBG>(&pempty_bar[acc_e], ring.parity(acc_e) ^ 1, args.err, 3, w_pempty);
                                    if (wide) mbar_wait_timed<DBG>(&pempty_bar[acc_e + 1], ring.parity(acc_e + 1) ^ 1, args.err, 3, w_pempty);
                                    dcol = tmem_base + acc_e * 128;
                                }
                                if (DBG && n0 == 0 && nstage < TF_STAGES) mbar_wait_timed<DBG>(&full_bar[stage], ph, args.err, 4, w_full_head);
                                else mbar_wait_timed<DBG>(&full_bar[stage], ph, args.err, 4, w_full);
                                tc_fence_after();
                                const uint32_t acc = in_seg > 0 ? 1u : 0u;
                                in_seg += take, done += take;
                                const bool seg_end = in_seg >= seg_kc || done == total;
                                if (elect_one()) {
                                    const uint64_t adesc = desc0 + (uint64_t)(stage * (TF_STAGE_BYTES >> 4));
                                    if (narrow) {   // [A(kc)][A(kc+1)][B(kc) 2 KB][B(kc+1) 2 KB]
                                        umma_f16_pair_x6(dcol, adesc, adesc + (2 * TF_TILE_BYTES >> 4), idesc, acc);
                                        if (take == 2)
                                            umma_f16_pair_x6(dcol, adesc + (TF_TILE_BYTES >> 4), adesc + ((2 * TF_TILE_BYTES + TF_NARROW_B) >> 4), idesc, 1u);
                                    } else {
                                        umma_f16_pair_x6(dcol, adesc, adesc + (TF_TILE_BYTES >> 4), idesc, acc);
                                    }
                                    umma_commit_pair(&empty_bar[stage]);   // frees the smem stage in both CTAs when these MMAs retire
                                    if (seg_end) {   // partial tiles complete -> both epilogues drain them
                                        umma_commit_pair(&pfull_bar[acc_e]);
                                        if (wide) umma_commit_pair(&pfull_bar[acc_e + 1]);
                                    }
                                }
                                if (seg_end) ring.close(acc_e, wide), in_seg = 0;
                                if (++stage == TF_STAGES) stage = 0, ph ^= 1;
                            }
                        }
                    }
                    if (DBG && lane == 0 && si < 24) args.dbg[(size_t)blockIdx.x * TF_DBG_STRIDE + 16 + si] += clock64() - t_step;
                }
            }
            if (DBG && lane == 0) {
                long long *d = args.dbg + (size_t)blockIdx.x * TF_DBG_STRIDE;
                d[3] = clock64() - t_begin, d[4] = w_full + w_full_head, d[5] = w_pempty, d[14] = w_full_head;
            }
        }
    } else {
        // =============================== TMA store issuers (column group 0 / 1) ===============================
        if (lane == 0) {
            const int gi = warp - 2;
            uint8_t *stg_x = stg_all + gi * TF_STG_BYTES, *stg_y = stg_x + TF_BOX_BYTES;
            const CUtensorMap *map_st = maps + 1;
            uint32_t sidx = 0, pub[2] = {0, 0}, pend[2] = {0, 0};
            // Publication needs the stores to have LANDED (wait_group), which takes far longer than handing the
            // staging box back (wait_group.read).  Chunks are therefore published lazily, when a later box has been
            // issued and "all but the newest group complete" costs nothing -- except the last chunk of a layer pass,
            // whose consumer may be waiting for exactly these columns.
            auto publish = [&](bool blocking) {
                if (!(pend[0] | pend[1])) return;
                if (blocking) bulk_wait_all();
                else asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
                fence_async_all();
                for (int sl = 0; sl < 2; ++sl)
                    if (pend[sl]) {
                        pub[sl] += pend[sl], pend[sl] = 0;
                        st_release_u32(&ready_cnt[sl][gi], pub[sl]);
                    }
            };
            auto store_box = [&](int col, int slot) {
                mbar_wait(&sfull_bar[gi], sidx & 1, args.err, 6);      // the 128 epilogue threads have written the box
                tma_store_2d(stg_x, map_st, col, arena_row0 + slot * TF_M);          // k-chunk 0 of the 64 columns: hi | lo
                tma_store_2d(stg_y, map_st, col + 64, arena_row0 + slot * TF_M);     // k-chunk 1
                bulk_commit();
                bulk_wait_read();                                      // staging read out: hand it back
                mbar_arrive(&sfree_bar[gi]);
                ++sidx;
                publish(false);                                        // everything older than this box has landed
                ++pend[slot];                                          // publication unit = one 64-column box
            };
            for (int64_t pair = pair0; pair < my_pairs; pair += pair_step) {
                const int nslots = (args.slots == 2 && pair + 1 < my_pairs) ? 2 : 1;
                if (gi == 0)
                    for (int slot = 0; slot < nslots; ++slot) {
                        store_box(prog->in_col, slot);
                        publish(true);
                    }
                for (int si = 0; si < n_steps; ++si) {
                    const TfStep &st = s_steps[si];
                    if (st.dst < 0) continue;
                    for (int slot = 0; slot < nslots; ++slot)
                        for (int n0 = 0; n0 < st.N; n0 += TF_NC) {
                            const int c0 = n0 + 128 * gi;
                            if (c0 >= st.dst_pad) break;
                            store_box(st.dst + 2 * c0, slot);
                            if (c0 + 64 < st.dst_pad) store_box(st.dst + 2 * (c0 + 64), slot);
                            // The next box of this group is a whole chunk of MMAs away: flush now (the store warp has nothing
                            // else to do), so that the consumer layer can prefetch these columns at once.  Only between the
                            // two boxes of a chunk is publication lazy.
                            publish(true);
                        }
                }
            }
            publish(true);
        }
    }
    } else {
        // =============================== epilogue warps ===============================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(216));
        TfEpiCtx x;
        const int gi = (warp - 4) >> 2;                  // column group: columns [128 gi, 128 gi + 128) of every chunk
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;                   // TMEM lane == walker of the tile
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);   // this warp's lane quarter
        x.my_x = stg_all + gi * TF_STG_BYTES + row * 128, x.my_y = x.my_x + TF_BOX_BYTES;
        x.sx = smem_u32(x.my_x), x.sw16 = (uint32_t)(row & 7) << 4;
        x.sw = row & 7;
        x.sfree = &sfree_bar[gi], x.sfull = &sfull_bar[gi];
        x.sidx = 0;
        uint32_t *const mask_row0 = args.masks ? args.masks + (size_t)(arena_row0 + row) * prog->mask_words : nullptr;
        x.mask_row = mask_row0;
        x.err = args.err;
        x.chi = 0.0;
        x.t_sfree = 0;
        x.row_scale = 1.f;
        float row_unscale[2] = {1.f, 1.f}, row_scale2[2] = {1.f, 1.f};   // 2^k / 2^-k per walker-pair slot
        const int n_in = c.n_in;
        uint32_t nchunk = 0;
        TfAccRing ring;
        long long e_wait = 0, e_drain = 0, e_epi = 0;
        constexpr bool timing = DBG;
        x.timing = DBG;
        const long long e_begin = clock64();
        const uint32_t pempty_remote0 = map_to_cta(smem_u32(&pempty_bar[0]), 0);   // the leader's drain barriers, as cluster addresses

        for (int64_t pair = pair0; pair < my_pairs; pair += pair_step) {
            const int nslots = (args.slots == 2 && pair + 1 < my_pairs) ? 2 : 1;
            // ---- prologue: u -> theta -> xhat (util.py:323-347, :483-497), split, staged in group 0's boxes, TMA store.
            // Both column groups share the work (erf, log10 and a division per parameter: 8 k cycles on one group alone):
            // group g converts the parameters [16 j + 8 g, +8) of its walker.
            float lnprior2[2] = {0.f, 0.f};
            double chi2[2] = {0.0, 0.0};
            for (int slot = 0; slot < nslots; ++slot) {
                const int64_t grow = ((cluster_id + (pair + slot) * n_clusters) * 2 + cta_rank) * TF_M + row;
                const bool valid = grow < args.n;
                float lnprior = 0.f;
                if (gi == 0) mbar_wait(x.sfree, (x.sidx & 1) ^ 1, args.err, 7);
                asm volatile("bar.sync 1, 256;" ::: "memory");   // group 0's staging boxes are free
                uint8_t *px = stg_all + row * 128, *py = px + TF_BOX_BYTES;
                const float *u = args.in + grow * n_in;
#pragma unroll 1
                for (int j = gi; j < 8; j += 2) {
                    uint32_t hw[4], lw[4];
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                        float xv[2];
#pragma unroll
                        for (int z = 0; z < 2; ++z) {
                            const int i = 8 * j + e + z;
                            float xh = 0.f;
                            if (i < n_in && valid) {
                                const float uu = __ldg(u + i);
                                lnprior = fmaf(uu, uu, lnprior);
                                float th = args.input_theta ? uu : tf_prior_map(uu, c.prior_kind[i], c.prior_scale[i], c.prior_shift[i]);
                                if (c.log10_flag && c.log10_flag[i]) th = log10f(th);
                                xh = (th - c.x_mean[i]) / c.x_std[i];
                            }
                            xv[z] = xh;
                        }
                        split2(xv[0], xv[1], hw[e >> 1], lw[e >> 1]);
                    }
                    uint8_t *bx = (j & 4) ? py : px;
                    *reinterpret_cast<uint4 *>(bx + (((j & 3) ^ x.sw) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    *reinterpret_cast<uint4 *>(bx + (((4 + (j & 3)) ^ x.sw) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                }
                if (gi == 1) chi_s[row] = (double)lnprior;
                fence_async_smem();
                asm volatile("bar.sync 1, 256;" ::: "memory");   // both halves of the row are written
                if (gi == 0) {
                    lnprior2[slot] = -0.5f * (lnprior + (float)chi_s[row]);           // util.py:1165
                    mbar_arrive(x.sfull);
                    ++x.sidx;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");   // chi_s may be reused
            }
#pragma unroll 1
            for (int si = 0; si < n_steps; ++si) {
                const TfStep &st = s_steps[si];
                const int nch = (st.N + TF_NC - 1) / TF_NC;
                const long long t_step = timing ? clock64() : 0, w_step = e_wait, p_step = e_epi;
#pragma unroll 1
                for (int slot = 0; slot < nslots; ++slot) {
                const int64_t grow = ((cluster_id + (pair + slot) * n_clusters) * 2 + cta_rank) * TF_M + row;
                const bool valid = grow < args.n;
                x.mask_row = mask_row0 + (size_t)slot * TF_M * prog->mask_words;
                x.chi = 0.0;
                x.row_scale = (st.flags & TFF_ROWSCALE) ? row_scale2[slot] : 1.f;
#pragma unroll 1
                for (int ch = 0; ch < nch; ++ch) {
                    const int n0 = ch * TF_NC;
                    const int nvalid = st.N - n0 < TF_NC ? st.N - n0 : TF_NC;
                    const bool wide = ((nvalid + 31) & ~31) > 128;   // two ring entries (see TfAccRing)
                    int ncol = ((nvalid + 31) & ~31) - 128 * gi;   // accumulator columns of this group in this chunk
                    ncol = ncol < 0 ? 0 : (ncol > 128 ? 128 : ncol);
                    const int c0 = n0 + 128 * gi;
                    // this thread's share of the chunk's bias vector: fetched now, parked in shared memory after the drain
                    const float bpre = (st.bias && c0 < st.N + 64) ? __ldg(st.bias + c0 + (tid & 127)) : 0.f;
                    float racc[128];
#pragma unroll
                    for (int j = 0; j < 128; ++j) racc[j] = 0.f;
                    const int nseg = (tf_chunk_stages(st, n0) + seg_kc - 1) / seg_kc;
#pragma unroll 1
                    for (int sg = 0; sg < nseg; ++sg) {
                        const uint32_t acc_e = ring.open(wide);
                        const long long t_a = timing ? clock64() : 0;
                        mbar_wait(&pfull_bar[acc_e], ring.parity(acc_e), args.err, 5);
                        if (wide) mbar_wait(&pfull_bar[acc_e + 1], ring.parity(acc_e + 1), args.err, 5);
                        const long long t_b = timing ? clock64() : 0;
                        tc_fence_after();
#pragma unroll
                        for (int cb = 0; cb < 128; cb += 32) {
                            if (cb < ncol) {
                                uint32_t r[32];
                                tmem_ld32(tmem_lane + acc_e * 128 + gi * 128 + cb, r);
#pragma unroll
                                for (int j = 0; j < 32; ++j) racc[cb + j] += __uint_as_float(r[j]);
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {   // one arrival per warp at the leader
                            mbar_arrive_cluster(pempty_remote0 + 8 * acc_e);
                            if (wide) mbar_arrive_cluster(pempty_remote0 + 8 * (acc_e + 1));
                        }
                        ring.close(acc_e, wide);
                        if (timing) e_wait += t_b - t_a, e_drain += clock64() - t_b;
                    }
                    const long long t_c = timing ? clock64() : 0;
                    // ---------------- chunk epilogue (one compact specialisation per kind of step)
                    const int mword = st.mask_word + 8 * ch + 4 * gi;
                    float *bias_s = bias_stage[nchunk & 1][gi];
                    bias_s[tid & 127] = bpre;
                    ++nchunk;
                    asm volatile("bar.sync %0, 128;" ::"r"(2 + gi) : "memory");
                    if (c0 >= st.dst_pad && c0 >= st.N) continue;   // nothing of this chunk belongs to this group
                    switch (st.variant) {
                    case TFV_ACT: tf_chunk_epilogue<true, false, false, false, false, false, true>(racc, st, c0, mword, bias_s, x); break;
                    case TFV_ACT_SAVE: tf_chunk_epilogue<true, false, false, true, false, false, true>(racc, st, c0, mword, bias_s, x); break;
                    case TFV_CHI2: tf_chunk_epilogue<true, false, false, false, false, true, false>(racc, st, c0, mword, bias_s, x); break;
                    case TFV_CHI2_STORE: tf_chunk_epilogue<true, false, false, false, false, true, true>(racc, st, c0, mword, bias_s, x); break;
                    case TFV_BWD: tf_chunk_epilogue<false, false, false, false, true, false, true>(racc, st, c0, mword, bias_s, x); break;
                    case TFV_HEAD: tf_chunk_epilogue<true, true, false, false, false, false, true>(racc, st, c0, mword, bias_s, x); break;
                    case TFV_HEAD_EXP: tf_chunk_epilogue<true, true, true, false, false, false, true>(racc, st, c0, mword, bias_s, x); break;
                    case TFV_PREDICT: {
                        // yhat = W s + b ; y = yhat y_std + y_mean (exp) ; m = y sigma (linna/util.py:532-542, :457-458): the
                        // selected vector goes straight to global memory, 128 contiguous floats of this walker's row.
                        // (Static indices only: one dynamic index would move all 128 accumulators to local memory.)
                        bool bad = false;
                        if (valid) {
                            float *orow = args.out_vec + grow * c.n_out + c0;
                            const int nc = st.N - c0;   // columns of this group that exist (may exceed 128)
                            const bool vec4 = (c.n_out & 3) == 0;
                            const int okind = args.out_kind;
#pragma unroll
                            for (int g4 = 0; g4 < 32; ++g4) {
                                if (4 * g4 < nc) {
                                    float v[4];
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const int j = 4 * g4 + e;
                                        float t = fmaf(racc[j], st.inv_scale, bias_s[j]);
                                        t = max_nan(t, st.clampv);
                                        if (okind != LINNA_OUT_YHAT && j < nc) {
                                            t = fmaf(t, __ldg(c.y_std + c0 + j), __ldg(c.y_mean + c0 + j));   // util.py:542
                                            if (c.ypositive) t = expf(t);                                     // util.py:540
                                            if (okind == LINNA_OUT_M && c.sigma) t *= __ldg(c.sigma + c0 + j);  // util.py:458
                                        }
                                        bad |= (t != t) && j < nc;
                                        v[e] = t;
                                    }
                                    if (vec4 && 4 * g4 + 3 < nc) {
                                        *reinterpret_cast<float4 *>(orow + 4 * g4) = make_float4(v[0], v[1], v[2], v[3]);
                                    } else {
#pragma unroll
                                        for (int e = 0; e < 4; ++e)
                                            if (4 * g4 + e < nc) orow[4 * g4 + e] = v[e];
                                    }
                                }
                            }
                        }
                        if (bad && args.fix_rows) {   // an activation beyond the fp16 range: the FP32 kernel redoes this row
                            const int k = atomicAdd(args.fix_count, 1);
                            if (k < args.fix_cap) args.fix_rows[k] = (int32_t)grow;
                        }
                    } break;
                    default: {
                        // TFV_GRADOUT (n_in <= 64: group 0 only): chain through xhat = (theta' - mean)/std,
                        // theta' = log10(theta), theta = prior(u).  The accumulators go through this thread's own
                        // staging row so that the loop stays rolled.
                        mbar_wait(x.sfree, (x.sidx & 1) ^ 1, args.err, 7);
                        float *scr_a = reinterpret_cast<float *>(x.my_x), *scr_b = reinterpret_cast<float *>(x.my_y);
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            *reinterpret_cast<float4 *>(scr_a + i) = make_float4(racc[i], racc[i + 1], racc[i + 2], racc[i + 3]);
                            *reinterpret_cast<float4 *>(scr_b + i) = make_float4(racc[32 + i], racc[33 + i], racc[34 + i], racc[35 + i]);
                        }
                        if (valid) {
                            const float *u = args.in + grow * n_in;
                            float *gout = args.grad + grow * n_in;
#pragma unroll 1
                            for (int i = 0; i < n_in; ++i) {
                                const float uu = __ldg(u + i);
                                const int kind = c.prior_kind[i];
                                const float ps = c.prior_scale[i];
                                float gx = (i < 32 ? scr_a[i] : scr_b[i - 32]) * (st.inv_scale * row_unscale[slot]) / c.x_std[i];
                                if (c.log10_flag && c.log10_flag[i])
                                    gx /= (tf_prior_map(uu, kind, ps, c.prior_shift[i]) * 2.30258509299404568f);
                                float jac = ps;
                                if (kind == LINNA_PRIOR_FLAT) jac *= 0.398942280401432678f * expf(-0.5f * uu * uu);
                                gout[i] = gx * jac - uu;
                            }
                        }
                    } break;
                    }
                    if (timing) e_epi += clock64() - t_c;
                }
                chi2[slot] += x.chi;
                if (st.variant == TFV_CHI2_STORE) {
                    // The backward pass is linear in r: carry it at unit scale (r / 2^k, k = exponent of |r|) so that a
                    // walker far from the peak (|r| ~ 1e3) cannot push a gradient past the fp16 range; the last
                    // epilogue multiplies 2^k back.  Exact: powers of two only.
                    chi_x[gi][row] = x.chi;
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    const double tot = chi_x[0][row] + chi_x[1][row];
                    int k = (tot > 0.0 && tot < 1e300) ? (ilogb(tot) >> 1) : 0;
                    k = k < -60 ? -60 : (k > 60 ? 60 : k);
                    row_scale2[slot] = ldexpf(1.f, -k);
                    row_unscale[slot] = ldexpf(1.f, k);
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
                }
                if (timing && tid == 128 && si < 24) {
                    long long *d = args.dbg + (size_t)blockIdx.x * TF_DBG_STRIDE;
                    d[40 + si] += clock64() - t_step, d[64 + si] += e_wait - w_step, d[88 + si] += e_epi - p_step;
                }
            }
            // combine the two column groups of every walker and finish lnP
            for (int slot = 0; slot < nslots; ++slot) {
                const int64_t grow = ((cluster_id + (pair + slot) * n_clusters) * 2 + cta_rank) * TF_M + row;
                if (gi == 1) chi_s[row] = chi2[slot];
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (gi == 0 && grow < args.n && args.lnp) {
                    float l = (float)(-0.5 * (chi2[slot] + chi_s[row])) * c.inv_T + lnprior2[slot];   // util.py:1013
                    if (l != l) {
                        l = -INFINITY;                                                             // util.py:1015-1016
                        if (args.fix_rows) {   // fp16 overflow or a NaN input: the FP32 kernel decides (it redoes this row)
                            const int k = atomicAdd(args.fix_count, 1);
                            if (k < args.fix_cap) args.fix_rows[k] = (int32_t)grow;
                        }
                    }
                    args.lnp[grow] = l;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
        }
        if (DBG && (warp == 4 || warp == 8) && lane == 0) {
            long long *d = args.dbg + (size_t)blockIdx.x * TF_DBG_STRIDE + (warp == 4 ? 6 : 10);
            d[0] = clock64() - e_begin, d[1] = e_wait, d[2] = e_drain, d[3] = e_epi;
            if (warp == 8) args.dbg[(size_t)blockIdx.x * TF_DBG_STRIDE + 15] = x.t_sfree;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // neither CTA frees tensor memory (or exits) while its peer still computes on the pair
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcContext {
    __half *wblob = nullptr;      // packed hi/lo weight operands
    float *fblob = nullptr;       // effective biases
    __half *arena = nullptr;
    int ld = 0;                   // arena row pitch in halves
    uint32_t *masks = nullptr;
    CUtensorMap *maps_dev = nullptr;
    TfProgram *prog_dev = nullptr;   // [0] LNP, [1] GRAD
    int *err_dev = nullptr;
    long long *dbg_dev = nullptr;
    int32_t *fix_count = nullptr;     // [2]: ping-pong counters of flagged rows (launch k uses [k & 1])
    int32_t *fix_rows = nullptr;      // [fix_cap]
    int64_t fix_cap = 0;
    uint64_t launches = 0;
    bool has_lnp = false, has_grad = false, has_predict = false;
    int grid = 0;
    std::string error;
};

static inline int pad64(int n) { return (n + 63) & ~63; }

// k-chunks (of 32) accumulated in tensor memory between two promotions to the register accumulators
static int tc_seg_kc()
{
    const char *e = getenv("LINNA_TC_SEG_KC");
    int v = e ? atoi(e) : 6;
    return v > 0 ? v : 6;
}

void tc_destroy(TcContext *t)
{
    if (!t) return;
    if (t->wblob) cudaFree(t->wblob);
    if (t->fblob) cudaFree(t->fblob);
    if (t->arena) cudaFree(t->arena);
    if (t->masks) cudaFree(t->masks);
    if (t->maps_dev) cudaFree(t->maps_dev);
    if (t->prog_dev) cudaFree(t->prog_dev);
    if (t->err_dev) cudaFree(t->err_dev);
    if (t->dbg_dev) cudaFree(t->dbg_dev);
    if (t->fix_count) cudaFree(t->fix_count);
    if (t->fix_rows) cudaFree(t->fix_rows);
    delete t;
}

namespace {
struct MatSrc {
    const float *W;   // source matrix
    int N, K;         // operand shape: B[N][K]
    bool transpose;   // false: W is [N][K]; true: W is [K][N]
    float premul;
};
struct Packer {
    std::vector<__half> w;       // weight blob (halves)
    std::vector<float> f;        // float blob (biases)
    struct Mat { size_t off; int N, K, ldk; };   // row n: per k-chunk of 32, [32 hi halves | 32 lo halves]
    std::vector<Mat> mats;
    size_t walloc(size_t n) { size_t o = (w.size() + 127) / 128 * 128; w.resize(o + n, __float2half_rn(0.f)); return o; }
    size_t fput(const std::vector<float> &v, int padded)
    {
        size_t o = (f.size() + 63) / 64 * 64;
        f.resize(o + padded, 0.f);
        std::copy(v.begin(), v.end(), f.begin() + o);
        return o;
    }
    static float at(const MatSrc &s, int n, int k)
    {
        return s.premul * (s.transpose ? s.W[(size_t)k * s.N + n] : s.W[(size_t)n * s.K + k]);
    }
    // power-of-two pre-scale that puts the largest |w| of a step's operands into [256, 512)
    static int pick_shift(const std::vector<MatSrc> &srcs)
    {
        float mx = 0.f;
        for (const MatSrc &s : srcs)
            for (int n = 0; n < s.N; ++n)
                for (int k = 0; k < s.K; ++k) {
                    const float a = fabsf(at(s, n, k));
                    if (a > mx && std::isfinite(a)) mx = a;
                }
        if (mx <= 0.f) return 0;
        int e;
        frexpf(mx, &e);            // mx = f * 2^e, f in [0.5, 1)
        int s = 9 - e;             // mx * 2^s in [256, 512)
        return std::max(-100, std::min(100, s));
    }
    int put(const MatSrc &s, int shift)
    {
        Mat mt;
        mt.N = s.N, mt.K = s.K, mt.ldk = 2 * TF_KC * ((s.K + TF_KC - 1) / TF_KC);
        mt.off = walloc((size_t)s.N * mt.ldk);
        const float sc = ldexpf(1.f, shift);
        for (int n = 0; n < s.N; ++n)
            for (int k = 0; k < s.K; ++k) {
                const float x = at(s, n, k) * sc;
                const __half hi = __float2half_rn(x);
                const size_t o = mt.off + (size_t)n * mt.ldk + 2 * TF_KC * (k / TF_KC) + k % TF_KC;
                w[o] = hi;
                w[o + TF_KC] = __float2half_rn(x - __half2float(hi));
            }
        mats.push_back(mt);
        return (int)mats.size() - 1;
    }
};
}  // namespace

// Build the tensor-core context of a model that already has likelihood constants.  Returns nullptr and
// fills `why` when the shape is unsupported or the driver lacks the tensor-map entry point.
TcContext *tc_build(const linna_model *m, std::string &why)
{
    if (m->has_extra) { why = "extra linear branch not supported on the tensor-core path"; return nullptr; }
    // the likelihood programs need the likelihood constants in Cholesky form; the predict program needs neither
    const bool like_ok = m->has_like && m->quad_kind == LINNA_QUAD_CHOL;
    if (m->n_in > 64) { why = "tensor-core path supports at most 64 input parameters"; return nullptr; }
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
    const int n_out = m->n_out;
    const int nops = (int)m->ops.size();
    const bool fold = like_ok && linna_can_fold(m);
    std::vector<float> Af, cf;
    if (fold) linna_fold_tail(m, Af, cf);
    for (const OpHost &op : m->ops)
        if (op.kind != LINNA_OP_LINEAR && !op.has_ws) { why = "identity skip not supported on the tensor-core path"; return nullptr; }
    if (m->ops.empty() || m->ops.back().kind != LINNA_OP_LINEAR || m->ops.back().act != LINNA_ACT_NONE) {
        why = "tensor-core path needs a plain linear last layer";
        return nullptr;
    }

    Packer P;
    // arena slots: X = xhat, A/B = ping-pong layer outputs, H = res-block hidden
    enum { SLOT_X = 0, SLOT_A = 1, SLOT_B = 2, SLOT_H = 3, NSLOT = 4 };
    int slot_w[NSLOT] = {64, 64, 64, 64};
    struct BiasRef { int prog, step, what; size_t off; };   // what: 0 bias, 1 vscale, 2 sub
    std::vector<BiasRef> bias_refs;
    std::vector<float> sigL;   // (diag(sigma) L) for the unfolded chi^2
    if (!fold && like_ok) {
        sigL.resize((size_t)n_out * n_out);
        for (int k = 0; k < n_out; ++k)
            for (int n = 0; n < n_out; ++n) sigL[(size_t)k * n_out + n] = m->sigma[k] * m->quad[(size_t)k * n_out + n];
    }

    TfProgram pgs[3];   // [0] lnP, [1] lnP + gradient, [2] predict
    const int seg_kc = tc_seg_kc();
    int mask_words_total = 0;
    bool has_grad = fold;   // the backward program is built on the folded tail only
    for (int pk = 0; pk < 3; ++pk) {
        TfProgram &pg = pgs[pk];
        memset(&pg, 0, sizeof pg);
        if (pk == 1 && !has_grad) continue;
        if (pk < 2 && !like_ok) continue;
        const bool grad = pk == 1, predict = pk == 2;
        const bool fold_here = fold && !predict;
        int ns = 0, pubs[2] = {1, 0};   // the prologue publishes one chunk of column group 0
        int slot_pub[NSLOT][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        int mask_words = 0;
        std::vector<int> mask_y(nops, -1), mask_h(nops, -1);
        bool overflow = false;
        auto new_step = [&]() -> TfStep & {
            if (ns >= TF_MAX_STEPS) { overflow = true; ns = TF_MAX_STEPS - 1; }
            TfStep &s = pg.steps[ns++];
            memset(&s, 0, sizeof s);
            s.nphase = 1, s.inv_scale = 1.f, s.dst = -1, s.mask_word = 0;
            return s;
        };
        // src slots are stored in s.src[] as slot ids first and turned into columns once the widths are known
        auto set_phase = [&](TfStep &s, int p, int slot, const MatSrc &ms, int shift) {
            s.src[p] = slot, s.K[p] = ms.K, s.mapB[p] = P.put(ms, shift);   // matrix index; turned into tensor-map indices below
            s.src_pub[p][0] = slot_pub[slot][0], s.src_pub[p][1] = slot_pub[slot][1];
        };
        auto set_dst = [&](TfStep &s, int slot) {
            s.dst = slot, s.dst_pad = pad64(s.N);
            slot_w[slot] = std::max(slot_w[slot], s.dst_pad);
            for (int g = 0; g < 2; ++g) {   // column group g stores the 64-column boxes of [256 ch + 128 g, +128) below dst_pad
                slot_pub[slot][g] = pubs[g];
                for (int n0 = 0; n0 < s.N; n0 += TF_NC)
                    pubs[g] += (n0 + 128 * g < s.dst_pad ? 1 : 0) + (n0 + 128 * g + 64 < s.dst_pad ? 1 : 0);
            }
        };
        auto set_bias = [&](const std::vector<float> &b, float scale, int N) {
            std::vector<float> e(N, 0.f);
            for (size_t i = 0; i < b.size() && (int)i < N; ++i) e[i] = scale * b[i];
            bias_refs.push_back({pk, ns - 1, 0, P.fput(e, pad64(N) + 256)});
        };
        // HEAD: y = yhat*y_std + y_mean (exp) ; m = y*sigma ; d = m - data (util.py:532-542, :457-458), handed on
        // as d/sigma = y - data/sigma so that it stays in the fp16 range whatever the units are.  Without exp the
        // whole map is one per-column affine of the accumulator.
        auto set_head = [&](const std::vector<float> &b, float scale, int N, float inv_scale) {
            std::vector<float> vs(N), hb(N), sub(N);
            for (int i = 0; i < N; ++i) {
                const double dos = (double)m->data[i] / (double)m->sigma[i];
                vs[i] = inv_scale * m->y_std[i];
                hb[i] = (float)((double)scale * b[i] * m->y_std[i] + m->y_mean[i] - (m->ypositive ? 0.0 : dos));
                sub[i] = (float)dos;
            }
            bias_refs.push_back({pk, ns - 1, 0, P.fput(hb, pad64(N) + 256)});
            bias_refs.push_back({pk, ns - 1, 1, P.fput(vs, pad64(N) + 256)});
            bias_refs.push_back({pk, ns - 1, 2, P.fput(sub, pad64(N) + 256)});
        };
        auto new_mask = [&](int N) { int o = mask_words; mask_words += 8 * ((N + TF_NC - 1) / TF_NC); return o; };
        auto other = [&](int b) { return b == SLOT_A ? SLOT_B : SLOT_A; };
        int cur = SLOT_X;
        // ------------------------------ forward
        for (int i = 0; i < nops; ++i) {
            const OpHost &op = m->ops[i];
            const bool last = i + 1 == nops;
            if (last && fold_here) break;
            if (op.kind == LINNA_OP_LINEAR) {
                TfStep &s = new_step();
                MatSrc ms{op.w.data(), op.out, op.in, false, 1.f};
                const int sh = Packer::pick_shift({ms});
                set_phase(s, 0, cur, ms, sh);
                s.N = op.out, s.inv_scale = ldexpf(1.f, -sh);
                if (op.act == LINNA_ACT_RELU) {
                    s.flags |= TFF_RELU;
                    if (grad) s.flags |= TFF_SAVE_MASK, s.mask_word = mask_y[i] = new_mask(op.out);
                }
                s.epi = last ? (predict ? TF_PREDICT : TF_HEAD) : TF_ACT;
                if (!(last && predict)) set_dst(s, other(cur));
                if (last && !predict) set_head(op.b, 1.f, op.out, s.inv_scale);
                else set_bias(op.b, 1.f, op.out);
                if (!(last && predict)) cur = s.dst;
            } else {
                TfStep &hs = new_step();
                MatSrc m1{op.w.data(), op.mid, op.in, false, 1.f};
                const int sh1 = Packer::pick_shift({m1});
                set_phase(hs, 0, cur, m1, sh1);
                hs.N = op.mid, hs.inv_scale = ldexpf(1.f, -sh1), hs.flags = TFF_RELU, hs.epi = TF_ACT;
                if (grad) hs.flags |= TFF_SAVE_MASK, hs.mask_word = mask_h[i] = new_mask(op.mid);
                set_dst(hs, SLOT_H);
                set_bias(op.b, 1.f, op.mid);
                // y = relu(alpha*(W2 h + b2) + Ws x): the long skip product first, so that the MMAs do not
                // wait for the epilogue of h
                TfStep &ys = new_step();
                MatSrc m3{op.ws.data(), op.out, op.in, false, 1.f};
                MatSrc m2{op.w2.data(), op.out, op.mid, false, op.alpha};
                const int sh = Packer::pick_shift({m2, m3});
                ys.nphase = 2;
                set_phase(ys, 0, cur, m3, sh);
                set_phase(ys, 1, SLOT_H, m2, sh);
                ys.N = op.out, ys.inv_scale = ldexpf(1.f, -sh), ys.flags = TFF_RELU, ys.epi = TF_ACT;
                if (grad) ys.flags |= TFF_SAVE_MASK, ys.mask_word = mask_y[i] = new_mask(op.out);
                set_dst(ys, other(cur));
                set_bias(op.b2, op.alpha, op.out);
                cur = ys.dst;
            }
        }
        auto prev_mask = [&](int i) -> int {   // relu mask of the producer of op i's input
            if (i <= 0) return -1;
            const OpHost &pv = m->ops[i - 1];
            return (pv.kind == LINNA_OP_RES || pv.act == LINNA_ACT_RELU) ? mask_y[i - 1] : -1;
        };
        if (predict) {
            // nothing after the last layer: its epilogue writes the selected vector
        } else if (fold) {
            // r = Af s + cf ; chi^2 = |r|^2
            const int K = m->ops.back().in, sbuf = cur;
            TfStep &q = new_step();
            MatSrc mq{Af.data(), n_out, K, false, 1.f};
            const int sh = Packer::pick_shift({mq});
            set_phase(q, 0, sbuf, mq, sh);
            q.N = n_out, q.inv_scale = ldexpf(1.f, -sh), q.epi = TF_CHI2;
            if (grad) set_dst(q, other(sbuf));
            set_bias(cf, 1.f, n_out);
            if (grad) {
                // d lnL / d s = -(1/T) Af^T r, masked by the relu of the layer that produced s
                TfStep &g0 = new_step();
                MatSrc mg{Af.data(), K, n_out, true, 1.f};
                const int shg = Packer::pick_shift({mg});
                set_phase(g0, 0, q.dst, mg, shg);
                g0.N = K, g0.inv_scale = ldexpf(1.f, -shg) * (-1.0f / m->temperature), g0.epi = TF_BWD;
                g0.flags |= TFF_ROWSCALE;
                const int pm = prev_mask(nops - 1);
                if (pm >= 0) g0.flags |= TFF_APPLY_MASK, g0.mask_word = pm;
                set_dst(g0, sbuf);
                cur = g0.dst;
                for (int i = nops - 2; i >= 0; --i) {
                    const OpHost &op = m->ops[i];
                    const int pm2 = prev_mask(i);
                    if (op.kind == LINNA_OP_LINEAR) {
                        TfStep &s2 = new_step();
                        MatSrc mb{op.w.data(), op.in, op.out, true, 1.f};
                        const int shb = Packer::pick_shift({mb});
                        set_phase(s2, 0, cur, mb, shb);
                        s2.N = op.in, s2.inv_scale = ldexpf(1.f, -shb), s2.epi = i == 0 ? TF_GRADOUT : TF_BWD;
                        if (pm2 >= 0) s2.flags |= TFF_APPLY_MASK, s2.mask_word = pm2;
                        if (i > 0) { set_dst(s2, other(cur)); cur = s2.dst; }
                    } else {
                        TfStep &h = new_step();
                        MatSrc mh{op.w2.data(), op.mid, op.out, true, op.alpha};
                        const int shh = Packer::pick_shift({mh});
                        set_phase(h, 0, cur, mh, shh);
                        h.N = op.mid, h.inv_scale = ldexpf(1.f, -shh), h.epi = TF_BWD, h.flags = TFF_APPLY_MASK, h.mask_word = mask_h[i];
                        set_dst(h, SLOT_H);
                        TfStep &x = new_step();
                        MatSrc mxs{op.ws.data(), op.in, op.out, true, 1.f};
                        MatSrc mxa{op.w.data(), op.in, op.mid, true, 1.f};
                        const int shx = Packer::pick_shift({mxs, mxa});
                        x.nphase = 2;
                        set_phase(x, 0, cur, mxs, shx);
                        set_phase(x, 1, SLOT_H, mxa, shx);
                        x.N = op.in, x.inv_scale = ldexpf(1.f, -shx), x.epi = i == 0 ? TF_GRADOUT : TF_BWD;
                        if (pm2 >= 0) x.flags |= TFF_APPLY_MASK, x.mask_word = pm2;
                        if (i > 0) { set_dst(x, other(cur)); cur = x.dst; }
                    }
                }
            }
        } else {
            // r = (d/sigma) . (diag(sigma) L): B[n][k] = sigma_k L[k][n], zero for k < n
            TfStep &q = new_step();
            MatSrc mq{sigL.data(), n_out, n_out, true, 1.f};
            const int sh = Packer::pick_shift({mq});
            set_phase(q, 0, cur, mq, sh);
            q.N = n_out, q.inv_scale = ldexpf(1.f, -sh), q.epi = TF_CHI2, q.flags = TFF_TRI;
            set_bias(std::vector<float>(), 1.f, n_out);
        }
        if (overflow) { why = "too many layers"; return nullptr; }
        for (int i = 0; i < ns; ++i)        // is phase p of step i the last reader of its slot before it is overwritten?
            for (int p = 0; p < pg.steps[i].nphase; ++p) {
                const int slot = pg.steps[i].src[p];
                bool last = true;
                for (int p2 = p + 1; p2 < pg.steps[i].nphase; ++p2) last &= pg.steps[i].src[p2] != slot;
                for (int j = i + 1; j < ns && last; ++j) {
                    bool reads = false;
                    for (int p2 = 0; p2 < pg.steps[j].nphase; ++p2) reads |= pg.steps[j].src[p2] == slot;
                    if (reads) last = false;
                    if (pg.steps[j].dst == slot) break;
                }
                if (last) pg.steps[i].flags |= p ? TFF_LAST_USE1 : TFF_LAST_USE0;
            }
        for (int i = 0; i < ns; ++i) {
            TfStep &s = pg.steps[i];
            s.clampv = (s.flags & TFF_RELU) ? 0.f : -INFINITY;
            switch (s.epi) {
            case TF_ACT: s.variant = (s.flags & TFF_SAVE_MASK) ? TFV_ACT_SAVE : TFV_ACT; break;
            case TF_HEAD: s.variant = m->ypositive ? TFV_HEAD_EXP : TFV_HEAD; break;
            case TF_CHI2: s.variant = s.dst >= 0 ? TFV_CHI2_STORE : TFV_CHI2; break;
            case TF_BWD: s.variant = TFV_BWD; break;
            case TF_PREDICT: s.variant = TFV_PREDICT; break;
            default: s.variant = TFV_GRADOUT; break;
            }
        }
        pg.n_steps = ns, pg.total_pub[0] = pubs[0], pg.total_pub[1] = pubs[1], pg.seg_kc = seg_kc;
        mask_words_total = std::max(mask_words_total, (mask_words + 3) & ~3);
    }
    // ---- arena columns
    int slot_col[NSLOT], ncol = 0;
    for (int s = 0; s < NSLOT; ++s) slot_col[s] = ncol, ncol += slot_w[s];
    const int ld = 2 * ncol;   // halves per arena row: every activation column is a (hi, lo) pair
    for (int pk = 0; pk < 3; ++pk) {
        TfProgram &pg = pgs[pk];
        pg.in_col = 2 * slot_col[SLOT_X], pg.mask_words = std::max(mask_words_total, 4);
        for (int i = 0; i < pg.n_steps; ++i) {
            TfStep &s = pg.steps[i];
            for (int p = 0; p < s.nphase; ++p) s.src[p] = 2 * slot_col[s.src[p]];
            if (s.dst >= 0) s.dst = 2 * slot_col[s.dst];
        }
    }

    TcContext *t = new TcContext();
    auto bail = [&](const std::string &msg) { why = msg; tc_destroy(t); return (TcContext *)nullptr; };
    t->grid = m->num_sms & ~1;   // whole CTA pairs
    t->ld = ld;
    t->has_lnp = like_ok, t->has_grad = like_ok && has_grad;
    t->has_predict = m->ops.back().kind == LINNA_OP_LINEAR && pgs[2].n_steps > 0;
    if (cudaMalloc(&t->wblob, P.w.size() * sizeof(__half)) != cudaSuccess) return bail("cudaMalloc weights");
    if (cudaMemcpy(t->wblob, P.w.data(), P.w.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) return bail("upload");
    if (cudaMalloc(&t->fblob, std::max<size_t>(P.f.size(), 64) * sizeof(float)) != cudaSuccess) return bail("cudaMalloc biases");
    if (!P.f.empty()) cudaMemcpy(t->fblob, P.f.data(), P.f.size() * sizeof(float), cudaMemcpyHostToDevice);
    const size_t rows = (size_t)t->grid * 2 * TF_M;   // two slots of 128 walkers per CTA
    if (cudaMalloc(&t->arena, rows * ld * sizeof(__half)) != cudaSuccess) return bail("cudaMalloc arena");
    cudaMemset(t->arena, 0, rows * ld * sizeof(__half));
    if (cudaMalloc(&t->masks, rows * (size_t)pgs[0].mask_words * sizeof(uint32_t)) != cudaSuccess) return bail("cudaMalloc masks");
    for (auto &br : bias_refs) {
        TfStep &s = pgs[br.prog].steps[br.step];
        (br.what == 0 ? s.bias : br.what == 1 ? s.vscale : s.sub) = t->fblob + br.off;
    }

    // ---- tensor maps
    // per weight matrix: a 128-row box for full chunks and a box as tall as one CTA's share of the last chunk
    auto tail_rows_of = [](int N) { const int last = N - (N - 1) / TF_NC * TF_NC; return ((last + 31) & ~31) / 2; };
    std::vector<CUtensorMap> maps(2 + 2 * P.mats.size());
    for (int pk = 0; pk < 3; ++pk)
        for (int i = 0; i < pgs[pk].n_steps; ++i) {
            TfStep &s = pgs[pk].steps[i];
            s.tail_rows = tail_rows_of(s.N);
            for (int p = 0; p < s.nphase; ++p) {
                const int mi = s.mapB[p];
                if (P.mats[mi].N != s.N) return bail("internal: operand rows != step columns");
                s.mapB[p] = 2 + 2 * mi, s.mapBt[p] = 3 + 2 * mi;
            }
        }
    auto encode2d = [&](CUtensorMap *mp, void *base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner,
                        uint32_t box_outer, CUtensorMapSwizzle sw) {
        cuuint64_t dims[2] = {inner, outer};
        cuuint64_t strides[1] = {pitch_bytes};
        cuuint32_t box[2] = {box_inner, box_outer};
        cuuint32_t estr[2] = {1, 1};
        return encode(mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    if (encode2d(&maps[0], t->arena, (uint64_t)ld, rows, (uint64_t)ld * 2, 2 * TF_KC, TF_M, CU_TENSOR_MAP_SWIZZLE_128B) != CUDA_SUCCESS)
        return bail("cuTensorMapEncodeTiled(arena load) failed");
    if (encode2d(&maps[1], t->arena, (uint64_t)ld, rows, (uint64_t)ld * 2, 64, TF_M, CU_TENSOR_MAP_SWIZZLE_128B) != CUDA_SUCCESS)
        return bail("cuTensorMapEncodeTiled(arena store) failed");
    for (size_t i = 0; i < P.mats.size(); ++i) {
        const Packer::Mat &mt = P.mats[i];
        if (encode2d(&maps[2 + 2 * i], t->wblob + mt.off, (uint64_t)mt.ldk, (uint64_t)mt.N, (uint64_t)mt.ldk * 2, 2 * TF_KC, TF_M,
                     CU_TENSOR_MAP_SWIZZLE_128B) != CUDA_SUCCESS ||
            encode2d(&maps[3 + 2 * i], t->wblob + mt.off, (uint64_t)mt.ldk, (uint64_t)mt.N, (uint64_t)mt.ldk * 2, 2 * TF_KC,
                     (uint32_t)tail_rows_of(mt.N), CU_TENSOR_MAP_SWIZZLE_128B) != CUDA_SUCCESS)
            return bail("cuTensorMapEncodeTiled(weights) failed");
    }
    if (cudaMalloc(&t->maps_dev, maps.size() * sizeof(CUtensorMap)) != cudaSuccess) return bail("cudaMalloc maps");
    cudaMemcpy(t->maps_dev, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    if (cudaMalloc(&t->prog_dev, 3 * sizeof(TfProgram)) != cudaSuccess) return bail("cudaMalloc prog");
    cudaMemcpy(t->prog_dev, pgs, 3 * sizeof(TfProgram), cudaMemcpyHostToDevice);
    if (cudaMalloc(&t->err_dev, sizeof(int)) != cudaSuccess) return bail("cudaMalloc err");
    cudaMemset(t->err_dev, 0, sizeof(int));
    if (cudaMalloc(&t->fix_count, 2 * sizeof(int32_t)) != cudaSuccess) return bail("cudaMalloc fix_count");
    cudaMemset(t->fix_count, 0, 2 * sizeof(int32_t));
    if (getenv("LINNA_TC_DEBUG")) {
        if (cudaMalloc(&t->dbg_dev, (size_t)t->grid * TF_DBG_STRIDE * sizeof(long long)) != cudaSuccess) return bail("cudaMalloc dbg");
        cudaMemset(t->dbg_dev, 0, (size_t)t->grid * TF_DBG_STRIDE * sizeof(long long));
    }
    if (cudaFuncSetAttribute(tc_f16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(tc_f16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM_BYTES) != cudaSuccess)
        return bail("cudaFuncSetAttribute(tc_f16_kernel)");
    if (cudaDeviceSynchronize() != cudaSuccess) return bail("sync after tc_build");
    return t;
}

bool tc_has_grad(const TcContext *t) { return t && t->has_grad; }
bool tc_has_lnp(const TcContext *t) { return t && t->has_lnp; }
bool tc_has_predict(const TcContext *t) { return t && t->has_predict; }

// LINNA_TC_DEBUG: per-CTA cycle counters of the last launch -> host ([grid][8]); returns the grid size
int tc_debug_read(TcContext *t, long long *out, int max_ctas)
{
    if (!t || !t->dbg_dev) return 0;
    const int n = std::min(max_ctas, t->grid);
    cudaDeviceSynchronize();
    cudaMemcpy(out, t->dbg_dev, (size_t)n * TF_DBG_STRIDE * sizeof(long long), cudaMemcpyDeviceToHost);
    return n;
}

// Where the launch that is about to be made will list its NaN rows: (rows, count of this launch, count of the next
// launch -- the fix-up kernel clears it).
void tc_fix_buffers(TcContext *t, const int32_t **rows, const int32_t **count, int32_t **next_count)
{
    *rows = t->fix_rows, *count = t->fix_count + (t->launches & 1), *next_count = t->fix_count + ((t->launches + 1) & 1);
}

static cudaError_t tc_launch(const linna_model *m, TcContext *t, int pk, const float *u, int64_t n, float *lnp, float *grad,
                             cudaStream_t stream, float *out_vec = nullptr, int out_kind = 0)
{
    if (t->fix_cap < n) {   // stream-ordered with every launch that used the old list (one model, chained launches)
        if (t->fix_rows) cudaFreeAsync(t->fix_rows, stream);
        t->fix_rows = nullptr, t->fix_cap = 0;
        const int64_t want = 2 * n + 1024;   // the predict program may list a row once per column group
        cudaError_t e = cudaMallocAsync(&t->fix_rows, (size_t)want * sizeof(int32_t), stream);
        if (e != cudaSuccess) return e;
        t->fix_cap = want;
    }
    TfArgs a;
    memset(&a, 0, sizeof a);
    a.fix_count = t->fix_count + (t->launches & 1), a.fix_rows = t->fix_rows;
    a.fix_cap = (int32_t)std::min<int64_t>(t->fix_cap, 0x7fffffff);
    a.prog = t->prog_dev + pk, a.maps = t->maps_dev, a.c = m->consts;
    a.in = u, a.lnp = lnp, a.grad = grad, a.masks = t->masks, a.n = n, a.err = t->err_dev, a.dbg = t->dbg_dev;
    a.out_vec = out_vec, a.out_kind = out_kind, a.input_theta = pk == 2 ? 1 : 0;
    const int64_t pairs = (n + 2 * TF_M - 1) / (2 * TF_M);
    // One cluster of two CTAs per TWO walker pairs ("slots"), interleaved layer by layer: the layer-to-layer
    // dependency bubble of one pair is filled with the other pair's MMAs (measured +4% on lnP, +1% on lnP+grad at
    // C3; the activation arena in flight doubles and the L2 hit rate drops from 78% to 58%, which is why it is not
    // more).  Batches that cannot fill every cluster twice spread one pair per cluster; LINNA_TC_SLOTS=1 forces that.
    const int64_t clusters = t->grid / 2;
    const int want_slots = getenv("LINNA_TC_SLOTS") ? atoi(getenv("LINNA_TC_SLOTS")) : 2;
    a.slots = (want_slots == 2 && pairs > clusters) ? 2 : 1;
    const int grid = 2 * (int)std::min<int64_t>(pairs, clusters);
    if (a.dbg) cudaMemsetAsync(t->dbg_dev, 0, (size_t)t->grid * TF_DBG_STRIDE * sizeof(long long), stream);
    if (a.dbg) tc_f16_kernel<true><<<grid, TF_THREADS, TF_SMEM_BYTES, stream>>>(a);
    else tc_f16_kernel<false><<<grid, TF_THREADS, TF_SMEM_BYTES, stream>>>(a);
    return cudaGetLastError();
}

// the launch has been followed by its fix-up kernel: the next one uses the other counter
void tc_launch_done(TcContext *t) { ++t->launches; }

cudaError_t tc_launch_lnp(const linna_model *m, TcContext *t, const float *u, int64_t n, float *lnp, cudaStream_t stream)
{
    return tc_launch(m, t, 0, u, n, lnp, nullptr, stream);
}

cudaError_t tc_launch_grad(const linna_model *m, TcContext *t, const float *u, int64_t n, float *lnp, float *grad,
                           cudaStream_t stream)
{
    return tc_launch(m, t, 1, u, n, lnp, grad, stream);
}

// Predictor.predict: physical parameters in, the selected vector (yhat | y | m) out
cudaError_t tc_launch_predict(const linna_model *m, TcContext *t, const float *theta, int64_t n, float *out, int out_kind,
                              cudaStream_t stream)
{
    return tc_launch(m, t, 2, theta, n, nullptr, nullptr, stream, out, out_kind);
}

}  // namespace linna
