// Small-batch (latency) emulator-likelihood kernel for sm_100a: one thread-block CLUSTER per tile of 8 walkers.
//
// The 8-row FFMA kernel gives a tile of walkers to ONE CTA, which then streams every weight of the network
// (3.4 MB at the README shape, 5.3 MB at DES-Y3 shape) through one SM: 0.17 ms for an ensemble of 4 walkers while
// 147 SMs idle (profiles/r1_bench_c1_ffma_v9.json).  Here the same step program (linna_device.cuh: the layers of
// linna/nn.py:45-56, :110-133, the inverse output transform of linna/util.py:532-542 / :457-458, the chi^2 of
// :953-955, the backward steps of the gradient program) is executed by a cluster of CS CTAs (8, or 16 where the
// device schedules it) that split the COLUMNS of every layer:
//
//   * CTA `rank` owns the output columns [rank*cpc, rank*cpc + cpc) of every step.  A thread owns one quad of those
//     columns and one k-lane: 8 rows x 4 columns of FP32 accumulators, one 16-byte weight load per k.
//   * weights are independent of the activations, so every thread prefetches ITS OWN weight stream through a private
//     cp.async ring (D entries of 16 bytes, D x 4 KB per CTA) that runs ahead across layer boundaries: no barrier
//     guards the ring (a thread only reads what it copied itself), and the layer-to-layer dependency chain never
//     waits for a weight that has not been requested yet.
//   * activations live in shared memory, replicated in every CTA of the cluster ([feature][8 rows] floats): the
//     epilogue of a step broadcasts its columns to all CS arenas with st.shared::cluster (distributed shared
//     memory), and one barrier.cluster per step orders the layers.  Nothing but the input, lnP, the gradient and a
//     predicted vector touches global memory.
//   * k-lanes are combined by warp shuffles (power-of-two quad counts) and a shared-memory pass in a fixed order;
//     chi^2 partials are summed per CTA and then over the ranks by the leader, also in a fixed order: results are
//     run-to-run deterministic.
//
// Arithmetic is the FP32 FFMA arithmetic of fused_ffma.cu (fp32 accumulation, chi^2 row sums in fp64); the summation
// order over k differs (k-lanes), so the two kernels agree to rounding, and both are held to the same parity bar.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "linna_device.cuh"
#include "../../include/linna_b200.h"

namespace linna {

constexpr int CL_THREADS = 256;
constexpr int CL_ROWS = 8;
constexpr int CL_RED_FLOATS = 8192;   // k-lane partials: [P][Q][4 columns][8 rows], P*Q <= 256

struct ClArgs {
    KernelArgs a;
    int32_t chi_q;     // quads per CTA of an n_out-wide step (sizes the chi^2 partial buffer)
    int32_t pad_;
    long long *dbg;    // LINNA_CLUSTER_DEBUG: [kMaxSteps][4] cycles of thread 0 of cluster 0 / rank 0: k-loop, k-lane
                       // reduction, epilogue + broadcast, cluster barrier
};

__device__ __forceinline__ unsigned cl_smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cl_cp_async16(unsigned saddr, const float *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cl_lds_v4(unsigned saddr, float4 &v)
{
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
}
__device__ __forceinline__ unsigned cl_opaque(unsigned v)
{
    unsigned r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ void cl_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cl_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ unsigned cl_mapa(unsigned saddr, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cl_st_remote_v4(unsigned caddr, float a, float b, float c, float d)
{
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// Asynchronous remote stores: the data lands in the peer's shared memory and the same transaction decrements the
// transaction count of the peer's mbarrier, so the peer learns that the bytes are there (and sees them) by waiting on
// its own barrier -- no fence, no cluster-wide barrier on the sending side.
__device__ __forceinline__ void cl_st_async_v4(unsigned caddr, unsigned cbar, float a, float b, float c, float d)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(caddr), "f"(a),
                 "f"(b), "f"(c), "f"(d), "r"(cbar)
                 : "memory");
}
__device__ __forceinline__ void cl_st_async_f64(unsigned caddr, unsigned cbar, double v)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(caddr), "d"(v), "r"(cbar) : "memory");
}
__device__ __forceinline__ void cl_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cl_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool cl_mbar_try_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void cl_cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float cl_prior_map(float u, int kind, float scale, float shift)
{
    float t = u;
    if (kind == LINNA_PRIOR_FLAT) t = 0.5f * (1.0f + erff(u / 1.41421356237309515f));  // gauss2unif, util.py:300
    return t * scale + shift;
}

// Reduce-scatter of the 32 accumulators of one quad over the 32/Q lanes of a warp that share it (lane = kslane*Q + q):
// every level exchanges half of the values a lane still carries.  Afterwards lane kslane holds the finished sums of the
// elements [kslane*Q, kslane*Q + Q) and stores them to dst[32] (element e = 8*column + row).
template <int Q>
__device__ __forceinline__ void cl_reduce_scatter(float (&v)[32], int lane, float *dst)
{
    int n = 32;
#pragma unroll
    for (int o = 16; o >= Q; o >>= 1) {
        n >>= 1;
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < n) {
                const float send = up ? v[i] : v[i + n];
                const float keep = up ? v[i + n] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
    }
    float *d = dst + (lane / Q) * Q;   // Q == 32: one lane per quad, all 32 values
    if (Q >= 4) {
#pragma unroll
        for (int i = 0; i < Q; i += 4) *reinterpret_cast<float4 *>(d + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < Q; ++i) d[i] = v[i];
    }
}

// Work split of one step inside the cluster: CTA `rank` owns Q quads of columns starting at col0; thread `tid` owns
// quad q and k-lane s of S (threads beyond Q*S idle).  `pow2`: the k-lanes of a warp can be combined by shuffles.
// The CTA-level part (Q, col0) is worked out once per kernel for every step (ClStepGeo, shared memory): the weight
// stream crosses a phase boundary at every second or third load, and with the integer divisions of this split inside it
// the stream bookkeeping, not the FFMAs, set the pace of the k-loop (ncu: the division sequences were executed for every
// third refill).
struct ClStepGeo {
    int Q, col0, qshift;   // qshift = log2(Q) when Q is a power of two, else -1
};
struct ClGeo {
    int Q, S, q, s, col0;
    bool active, pow2;
};
__device__ __forceinline__ ClStepGeo cl_step_geo(int N, int rank, int cs)
{
    ClStepGeo g;
    const int cpc = (((N + cs - 1) / cs) + 3) & ~3;
    g.col0 = rank * cpc;
    int nc = N - g.col0;
    nc = nc < 0 ? 0 : (nc > cpc ? cpc : nc);
    g.Q = (nc + 3) >> 2;
    g.qshift = (g.Q > 0 && (g.Q & (g.Q - 1)) == 0) ? 31 - __clz(g.Q) : -1;
    return g;
}
__device__ __forceinline__ ClGeo cl_geo(const ClStepGeo &sg, int tid)
{
    ClGeo g;
    g.Q = sg.Q, g.col0 = sg.col0;
    g.pow2 = sg.qshift >= 0 && sg.Q <= 32;
    if (sg.qshift >= 0) {
        g.q = tid & (sg.Q - 1), g.s = tid >> sg.qshift, g.S = CL_THREADS >> sg.qshift;
    } else if (sg.Q > 0) {
        g.s = tid / sg.Q, g.q = tid - g.s * sg.Q, g.S = CL_THREADS / sg.Q;
    } else {
        g.q = 0, g.s = 0, g.S = 1;
    }
    g.active = sg.Q > 0 && g.s < g.S;
    return g;
}

// This thread's private weight stream: the float4 w[k][col0 + 4q .. +3] for k = s, s + S, ... of every phase of every
// step of every tile of the cluster, in program order.
struct ClStream {
    const Step *steps;
    const ClStepGeo *geo;
    int n_steps, tid;
    long long tiles_left;
    int si, ph;
    int k, K, S;
    const float *p;
    size_t stride;
    bool done;

    __device__ __forceinline__ void open_phase()
    {
        // find the next (tile, step, phase) in which this thread has at least one load
        for (;;) {
            if (tiles_left <= 0) { done = true; return; }
            if (si >= n_steps) { si = 0, ph = 0, --tiles_left; continue; }
            const Step &st = steps[si];
            const int Kp = ph == 0 ? st.K1 : st.K2;
            const float *w = ph == 0 ? st.wt1 : st.wt2;
            if (ph > 1) { ++si, ph = 0; continue; }
            if (Kp > 0 && w) {
                const ClGeo g = cl_geo(geo[si], tid);
                if (g.active && g.s < Kp) {
                    const int ldw = ph == 0 ? st.ldw1 : st.ldw2;
                    k = g.s, K = Kp, S = g.S;
                    p = w + (size_t)g.s * ldw + g.col0 + 4 * g.q;
                    stride = (size_t)g.S * ldw;
                    return;
                }
            }
            ++ph;
        }
    }
    __device__ __forceinline__ void init(const Step *steps_, const ClStepGeo *geo_, int n_steps_, int tid_, long long tiles)
    {
        steps = steps_, geo = geo_, n_steps = n_steps_, tid = tid_, tiles_left = tiles;
        si = 0, ph = 0, done = false, k = 0, K = 0, S = 1, p = nullptr, stride = 0;
        open_phase();
    }
    __device__ __forceinline__ void next()
    {
        k += S, p += stride;
        if (k >= K) { ++ph; open_phase(); }
    }
};

template <int D>
__global__ void __launch_bounds__(CL_THREADS, 1) cluster_ffma_kernel(const ClArgs cargs)
{
    extern __shared__ float4 cl_smem4[];
    __shared__ Step s_steps[kMaxSteps];
    __shared__ ClStepGeo s_geo[kMaxSteps];
    __shared__ __align__(8) uint64_t full_bar[2];   // "the outputs of step s have landed in this CTA's arena", by step parity
    const KernelArgs &args = cargs.a;
    const Program *__restrict__ prog = args.prog;
    const Consts &c = args.c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned rank, cs, cid, ncl;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(cs));
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(ncl));

    // ---- shared-memory carve-up
    float *ring = reinterpret_cast<float *>(cl_smem4);                    // [D][256] float4
    float *red = ring + (size_t)D * CL_THREADS * 4;                       // [CL_RED_FLOATS]
    float *arena = red + CL_RED_FLOATS;                                   // [arena_features][8]
    double *chi_part = reinterpret_cast<double *>(arena + (size_t)prog->arena_features * CL_ROWS);   // [chi_q*8][4]
    double *chi_all = chi_part + (size_t)cargs.chi_q * 32;                // [cs][8]   (leader's copy is the one that is read)
    double *chi_acc = chi_all + 16 * CL_ROWS;                             // [8]
    float *lnprior = reinterpret_cast<float *>(chi_acc + CL_ROWS);        // [8]
    uint8_t *masks = reinterpret_cast<uint8_t *>(lnprior + CL_ROWS);      // [mask_features][2]: one nibble of row bits per row half

    const int n_in = c.n_in, n_out = c.n_out;
    const int n_steps = prog->n_steps;
    const bool ktiming = cargs.dbg && tid == 0 && rank == 0 && cid == 0;   // LINNA_CLUSTER_DEBUG: phases outside the step loop
    const long long k_t0 = ktiming ? clock64() : 0;
    {
        const int nwords = n_steps * (int)(sizeof(Step) / 4);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(prog->steps);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_steps);
        for (int i = tid; i < nwords; i += CL_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    if (tid < n_steps) s_geo[tid] = cl_step_geo(s_steps[tid].N, (int)rank, (int)cs);
    __syncthreads();

    const int64_t n_rows = args.n;
    const int64_t ntiles = (n_rows + CL_ROWS - 1) / CL_ROWS;
    const int64_t my_tiles = (int64_t)cid < ntiles ? (ntiles - cid + ncl - 1) / ncl : 0;

    const unsigned arena_s = cl_opaque(cl_smem_u32(arena));
    const unsigned chi_all_s = cl_opaque(cl_smem_u32(chi_all));

    const unsigned bar_s = cl_opaque(cl_smem_u32(full_bar));
    if (tid == 0) {
        cl_mbar_init(bar_s, 1), cl_mbar_init(bar_s + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned gstep = 0, phase_bits = 0;   // steps executed so far (over all tiles): a step uses barrier gstep & 1; bit b of
                                          // phase_bits is the parity of the phases barrier b has completed on THIS CTA
    // The cluster barrier comes BEFORE the first weight loads: its release fence would otherwise wait for every one of
    // them (a full DRAM / L2 round trip on top of the time the slowest CTA needs to start).
    const long long k_t1 = ktiming ? clock64() : 0;
    cl_cluster_sync();    // every CTA of the cluster is running, with its barriers initialised, before anything is written across it
    const long long k_t2 = ktiming ? clock64() : 0;
    // ---- weight stream: D loads in flight per thread, across step and tile boundaries
    ClStream ws;
    ws.init(s_steps, s_geo, n_steps, tid, my_tiles);
    // shared-space addresses are taken once and laundered through an opaque move: ptxas otherwise rebuilds the shared
    // window base (S2UR SR_CgaCtaId + LEA) in front of every use, inside the k-loop
    const unsigned ring_s = cl_opaque(cl_smem_u32(ring) + (unsigned)tid * 16u);
    unsigned cons = 0;
#pragma unroll 1
    for (int i = 0; i < D; ++i) {
        if (!ws.done) {
            cl_cp_async16(ring_s + (unsigned)i * (CL_THREADS * 16u), ws.p);
            ws.next();
        }
        cl_commit();
    }

    for (int64_t tl = 0; tl < my_tiles; ++tl) {
        const int64_t row0 = ((int64_t)cid + tl * ncl) * CL_ROWS;
        const int nrows = (int)((n_rows - row0) < CL_ROWS ? (n_rows - row0) : CL_ROWS);

        // ---------------- prologue: u -> theta -> xhat (every CTA keeps its own copy) ----------------
        {
            float *xb = arena + (size_t)prog->in_buf * CL_ROWS;
            const float *in = args.in;
            for (int e = tid; e < CL_ROWS * n_in; e += CL_THREADS) {
                const int r = e / n_in, i = e - r * n_in;
                float th = 0.f;
                if (r < nrows) {
                    const float u = in[(row0 + r) * n_in + i];
                    th = args.input_theta ? u : cl_prior_map(u, c.prior_kind[i], c.prior_scale[i], c.prior_shift[i]);
                    if (c.log10_flag && c.log10_flag[i]) th = log10f(th);            // util.py:491-496
                    th = (th - c.x_mean[i]) / c.x_std[i];                             // util.py:497
                }
                xb[(size_t)i * CL_ROWS + r] = th;
            }
            {   // lnprior = -|u|^2 / 2 (util.py:1165): warp w sums row w (lanes stride over the parameters, fixed shuffle tree)
                float s = 0.f;
                if (warp < nrows && !args.input_theta) {
                    const float *ur = in + (row0 + warp) * n_in;
                    for (int i = lane; i < n_in; i += 32) { const float u = ur[i]; s = fmaf(u, u, s); }
                }
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) lnprior[warp] = -0.5f * s, chi_acc[warp] = 0.0;
            }
        }
        __syncthreads();

        // ---------------- the step program ----------------
#pragma unroll 1
        for (int si = 0; si < n_steps; ++si) {
            const Step &st = s_steps[si];
            const int N = st.N;
            const ClGeo g = cl_geo(s_geo[si], tid);
            const bool timing = cargs.dbg && tid == 0 && rank == 0 && cid == 0;
            const long long t0 = timing ? clock64() : 0;
            // Hand-over of the step's outputs.  Every CTA sends each of its columns to all arenas with st.async, which
            // completes bytes on the receiver's mbarrier: a CTA moves on to the next step when 32 bytes per column of the
            // layer have arrived (a CTA that runs ahead can be at most one step ahead -- it needs everybody's columns of
            // this step, which are sent after their k-loops -- so two barriers, by step parity, are enough, and nothing a
            // slower CTA still reads is overwritten).  Steps that broadcast nothing end in a plain cluster barrier.
            const bool chi_step = st.epi == EPI_CHI2;
            const bool sends = st.dst >= 0 && (st.epi == EPI_ACT || st.epi == EPI_HEAD || st.epi == EPI_BWD || (chi_step && (st.flags & F_STORE_DST)));
            const unsigned expect = (sends ? 32u * (unsigned)N : 0u) + (chi_step && rank == 0 ? cs * 64u : 0u);
            const unsigned my_bar = bar_s + (gstep & 1u) * 8u, my_par = (phase_bits >> (gstep & 1u)) & 1u;
            if (expect && tid == 0) cl_mbar_expect_tx(my_bar, expect);
            float acc[32];   // [column of the quad][row]: acc[8 cc + r]
#pragma unroll
            for (int e = 0; e < 32; ++e) acc[e] = 0.f;
            // this thread's first epilogue item is known now: fetch its bias while the k-loop runs
            float bpre = 0.f;
            {
                const int cidx0 = g.col0 + 4 * (tid >> 3) + ((tid >> 1) & 3);
                if (st.bias && tid < 8 * g.Q && cidx0 < N) bpre = __ldg(st.bias + cidx0);
            }
            if (g.active) {
#pragma unroll 1
                for (int ph = 0; ph < 2; ++ph) {
                    const int K = ph == 0 ? st.K1 : st.K2;
                    const float *wp = ph == 0 ? st.wt1 : st.wt2;
                    if (ph == 1 && st.scale != 1.0f) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) acc[e] *= st.scale;
                    }
                    if (K <= 0 || !wp) continue;
                    // shared-space addresses throughout the k-loop (a generic pointer makes the compiler rebuild the
                    // CTA's shared window base from %cluster_ctaid every iteration)
                    const unsigned ap_s = arena_s + (unsigned)(ph == 0 ? st.src1 : st.src2) * (CL_ROWS * 4u);
                    const int S = g.S;
                    int k = g.s;
                    // four ring entries per trip: the shared-memory latencies of one trip hide behind 128 FFMAs
#pragma unroll 1
                    for (; k + 3 * S < K; k += 4 * S) {
                        cl_wait<D - 4>();   // this thread's four oldest loads have landed
                        float4 w[4], a0[4], a1[4];
                        unsigned slot[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            slot[u] = ring_s + ((cons + u) & (unsigned)(D - 1)) * (CL_THREADS * 16u);
                            cl_lds_v4(slot[u], w[u]);
                            cl_lds_v4(ap_s + (unsigned)(k + u * S) * 32u, a0[u]);
                            cl_lds_v4(ap_s + (unsigned)(k + u * S) * 32u + 16u, a1[u]);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float a[8] = {a0[u].x, a0[u].y, a0[u].z, a0[u].w, a1[u].x, a1[u].y, a1[u].z, a1[u].w};
                            const float b[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                                for (int r = 0; r < 8; ++r) acc[8 * cc + r] = fmaf(a[r], b[cc], acc[8 * cc + r]);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {   // refill the slots that were just read
                            if (!ws.done) {
                                cl_cp_async16(slot[u], ws.p);
                                ws.next();
                            }
                            cl_commit();
                        }
                        cons += 4;
                    }
#pragma unroll 1
                    for (; k < K; k += S) {
                        cl_wait<D - 1>();   // this thread's oldest load has landed
                        const unsigned slot = ring_s + (cons & (unsigned)(D - 1)) * (CL_THREADS * 16u);
                        float4 w, a0, a1;
                        cl_lds_v4(slot, w);
                        cl_lds_v4(ap_s + (unsigned)k * 32u, a0);
                        cl_lds_v4(ap_s + (unsigned)k * 32u + 16u, a1);
                        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        const float b[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                            for (int r = 0; r < 8; ++r) acc[8 * cc + r] = fmaf(a[r], b[cc], acc[8 * cc + r]);
                        if (!ws.done) {   // refill the slot that was just read
                            cl_cp_async16(slot, ws.p);
                            ws.next();
                        }
                        cl_commit();
                        ++cons;
                    }
                }
            }
            const long long t1 = timing ? clock64() : 0;
            // ---- combine the k-lanes.  Power-of-two quad counts: the 32/Q lanes of a warp that share a quad run a
            //      reduce-scatter (each level halves the values a lane carries: 31 shuffles instead of 160 for Q = 1),
            //      and every lane leaves its Q finished sums of the warp in shared memory; otherwise every k-lane leaves
            //      its 32 partials.  The second pass adds the partial sets in a fixed order.
            int P;
            if (g.pow2) {
                P = CL_THREADS / 32;   // one partial set per warp
                float *dst = red + (size_t)(warp * g.Q + g.q) * 32;
                switch (g.Q) {
                case 1: cl_reduce_scatter<1>(acc, lane, dst); break;
                case 2: cl_reduce_scatter<2>(acc, lane, dst); break;
                case 4: cl_reduce_scatter<4>(acc, lane, dst); break;
                case 8: cl_reduce_scatter<8>(acc, lane, dst); break;
                case 16: cl_reduce_scatter<16>(acc, lane, dst); break;
                default: cl_reduce_scatter<32>(acc, lane, dst); break;
                }
            } else {
                P = g.S;
                if (g.active) {
                    float4 *dst = reinterpret_cast<float4 *>(red + (size_t)(g.s * g.Q + g.q) * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
                }
            }
            __syncthreads();

            const long long t2 = timing ? clock64() : 0;
            // ---- epilogue: one item = one column x one half of the rows
            for (int item = tid; item < 8 * g.Q; item += CL_THREADS) {
                const int q = item >> 3, cc = (item >> 1) & 3, rh = item & 1;
                const int cidx = g.col0 + 4 * q + cc;
                double part[4] = {0.0, 0.0, 0.0, 0.0};
                if (cidx < N) {
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int p = 0; p < P; ++p) {
                        const float4 t = *reinterpret_cast<const float4 *>(red + (size_t)(p * g.Q + q) * 32 + cc * 8 + rh * 4);
                        v[0] += t.x, v[1] += t.y, v[2] += t.z, v[3] += t.w;
                        if (timing && p == 0) cargs.dbg[4 * kMaxSteps + 4 * si + 2] += (v[0] != 12345.f ? clock64() : 0) - t2;
                    }
                    if (timing) cargs.dbg[4 * kMaxSteps + 4 * si + 0] += clock64() - t2;
                    const float b = st.bias ? st.scale * (item == tid ? bpre : __ldg(st.bias + cidx)) : 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] += b;
                    if (timing) cargs.dbg[4 * kMaxSteps + 4 * si + 1] += (v[0] != 12345.f ? clock64() : 0) - t2;
                    const int rb = 4 * rh;   // first row of this item
                    auto local4 = [&](int feature) -> float4 {
                        return *reinterpret_cast<const float4 *>(arena + (size_t)feature * CL_ROWS + rb);
                    };
                    auto bcast = [&](int feature, const float (&x)[4]) {
                        const unsigned la = arena_s + ((unsigned)feature * CL_ROWS + rb) * 4u;
                        for (unsigned r = 0; r < cs; ++r) cl_st_async_v4(cl_mapa(la, r), cl_mapa(my_bar, r), x[0], x[1], x[2], x[3]);
                    };
                    if (st.flags & F_ADD_SRC2) {
                        const float4 s2 = local4(st.src2 + cidx);
                        v[0] += s2.x, v[1] += s2.y, v[2] += s2.z, v[3] += s2.w;
                    }
                    if (st.epi == EPI_ACT || st.epi == EPI_HEAD) {
                        if (st.flags & F_RELU) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
                        }
                        if (st.flags & F_SAVE_MASK) {
                            unsigned mb = 0;
#pragma unroll
                            for (int i = 0; i < 4; ++i) mb |= (v[i] > 0.f ? 1u : 0u) << i;
                            masks[(size_t)(st.mask_off + cidx) * 2 + rh] = (uint8_t)mb;
                        }
                    }
                    if (st.epi == EPI_ACT) {
                        bcast(st.dst + cidx, v);
                    } else if (st.epi == EPI_HEAD) {
                        const float ys = __ldg(c.y_std + cidx), ym = __ldg(c.y_mean + cidx);
                        const float sg = c.sigma ? __ldg(c.sigma + cidx) : 1.f;
                        const float dt = c.data ? __ldg(c.data + cidx) : 0.f;
                        float y[4], d[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float yy = fmaf(v[i], ys, ym);                        // util.py:542
                            if (c.ypositive) yy = expf(yy);                      // util.py:540
                            y[i] = yy;
                            d[i] = yy * sg - dt;                                  // util.py:458, :954
                        }
                        bcast(st.dst + cidx, d);
                        if (st.flags & F_SAVE_Y)
                            *reinterpret_cast<float4 *>(arena + (size_t)(st.ybuf + cidx) * CL_ROWS + rb) = make_float4(y[0], y[1], y[2], y[3]);
                        if ((st.flags & F_OUT_VEC) && args.out_vec) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int r = rb + i;
                                if (r < nrows) {
                                    const float o = args.out_kind == LINNA_OUT_YHAT ? v[i] : args.out_kind == LINNA_OUT_Y ? y[i] : y[i] * sg;
                                    args.out_vec[(row0 + r) * n_out + cidx] = o;
                                }
                            }
                        }
                    } else if (st.epi == EPI_CHI2) {
                        if (c.quad_kind == LINNA_QUAD_CHOL) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) part[i] = (double)(v[i] * v[i]);
                        } else {
                            const float4 d4 = local4(st.src1 + cidx);
                            const float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) part[i] = (double)(v[i] * d[i]);
                        }
                        if (st.flags & F_STORE_DST) bcast(st.dst + cidx, v);
                    } else if (st.epi == EPI_BWD) {
                        if (st.colscale) {
                            const float csx = __ldg(st.colscale + cidx);
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[i] *= csx;
                        }
                        if (st.flags & F_MUL_YSAVE) {
                            const float4 y4 = local4(st.ybuf + cidx);
                            v[0] *= y4.x, v[1] *= y4.y, v[2] *= y4.z, v[3] *= y4.w;
                        }
                        if (st.flags & F_APPLY_MASK) {
                            const unsigned mb = masks[(size_t)(st.mask_off + cidx) * 2 + rh];
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[i] = ((mb >> i) & 1u) ? v[i] : 0.f;
                        }
                        bcast(st.dst + cidx, v);
                    } else if (st.epi == EPI_GRAD) {
                        // chain through xhat = (theta' - mean)/std, theta' = log10(theta), theta = prior(u)
                        const int kind = args.input_theta ? 0 : c.prior_kind[cidx];
                        const float ps = args.input_theta ? 1.f : c.prior_scale[cidx], psh = args.input_theta ? 0.f : c.prior_shift[cidx];
                        const float inv_std = 1.0f / c.x_std[cidx];
                        const bool lg = c.log10_flag && c.log10_flag[cidx];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = rb + i;
                            if (r < nrows) {
                                const int64_t gr = row0 + r;
                                const float u = args.in[gr * n_in + cidx];
                                float gx = v[i] * inv_std;
                                if (args.input_theta) {
                                    if (lg) gx /= (u * 2.30258509299404568f);
                                    args.grad[gr * n_in + cidx] = gx;
                                    continue;
                                }
                                if (lg) gx /= (cl_prior_map(u, kind, ps, psh) * 2.30258509299404568f);
                                float jac = ps;
                                if (kind == LINNA_PRIOR_FLAT) jac *= 0.398942280401432678f * expf(-0.5f * u * u);
                                args.grad[gr * n_in + cidx] = gx * jac - u;
                            }
                        }
                    }
                }
                if (chi_step) {
                    double *cp = chi_part + (size_t)item * 4;
                    cp[0] = part[0], cp[1] = part[1], cp[2] = part[2], cp[3] = part[3];
                }
            }
            if (chi_step) {
                __syncthreads();
                if (tid < CL_ROWS) {   // this CTA's columns, in column order; then to the leader
                    double s = 0.0;
                    const int rh = tid >> 2, ri = tid & 3;
                    for (int it = rh; it < 8 * g.Q; it += 2) s += chi_part[(size_t)it * 4 + ri];
                    cl_st_async_f64(cl_mapa(chi_all_s + (unsigned)(rank * CL_ROWS + tid) * 8u, 0), cl_mapa(my_bar, 0), s);
                }
            }
            const long long t3 = timing ? clock64() : 0;
            if (expect) {   // this step's columns (and, on the leader, the chi^2 partials) have landed here
                while (!cl_mbar_try_wait(my_bar, my_par)) {}
                phase_bits ^= 1u << (gstep & 1u);
            }
            if (!sends) cl_cluster_sync();   // nothing was handed over: no CTA may run ahead into buffers a peer still reads
            ++gstep;
            if (timing) {
                long long *d = cargs.dbg + 4 * si;
                d[0] += t1 - t0, d[1] += t2 - t1, d[2] += t3 - t2, d[3] += clock64() - t3;
            }
            if (chi_step && rank == 0 && tid < CL_ROWS) {
                double s = chi_acc[tid];
                for (unsigned r = 0; r < cs; ++r) s += chi_all[r * CL_ROWS + tid];
                chi_acc[tid] = s;
                if (tid < nrows && args.lnp) {
                    float l = (float)(-0.5 * s) * c.inv_T + lnprior[tid];       // util.py:1013
                    if (l != l) l = -INFINITY;                                    // util.py:1015-1016
                    args.lnp[row0 + tid] = l;
                }
            }
        }
        __syncthreads();
    }
    const long long k_t3 = ktiming ? clock64() : 0;
    cl_wait<0>();
    cl_cluster_sync();   // no CTA exits while a peer may still write into its shared memory
    if (ktiming) {
        long long *d = cargs.dbg + 4 * kMaxSteps + 4 * 60;   // [set-up + ring prefill, first cluster barrier, tiles, last barrier]
        d[0] += k_t1 - k_t0, d[1] += k_t2 - k_t1, d[2] += k_t3 - k_t2, d[3] += clock64() - k_t3;
    }
}

size_t cluster_ffma_smem_bytes(const Program &pg, int n_out, int cs, int depth, int *chi_q_out)
{
    const int cpc = (((n_out + cs - 1) / cs) + 3) & ~3;
    const int chi_q = cpc / 4;
    if (chi_q_out) *chi_q_out = chi_q;
    size_t b = (size_t)depth * CL_THREADS * 16 + (size_t)CL_RED_FLOATS * 4 + (size_t)pg.arena_features * CL_ROWS * 4;
    b += (size_t)chi_q * 32 * 8 + 16 * CL_ROWS * 8 + CL_ROWS * 8 + CL_ROWS * 4;
    b += (size_t)std::max(pg.mask_features, 1) * 2 + 16;
    return (b + 15) & ~(size_t)15;
}

// Cluster sizes this device schedules for the kernel at `smem` bytes (0: none).
int cluster_ffma_max_clusters(int cs, int depth, size_t smem)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(cs, 1, 1), cfg.blockDim = dim3(CL_THREADS, 1, 1), cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    int nc = 0;
    cudaError_t e;
    if (depth == 16) {
        cudaFuncSetAttribute(cluster_ffma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cs > 8) cudaFuncSetAttribute(cluster_ffma_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        e = cudaOccupancyMaxActiveClusters(&nc, cluster_ffma_kernel<16>, &cfg);
    } else {
        cudaFuncSetAttribute(cluster_ffma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cs > 8) cudaFuncSetAttribute(cluster_ffma_kernel<8>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        e = cudaOccupancyMaxActiveClusters(&nc, cluster_ffma_kernel<8>, &cfg);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return nc;
}

static long long *g_cl_dbg = nullptr;

// LINNA_CLUSTER_DEBUG: per-step cycle counters of the launches so far ([kMaxSteps][4]); returns the number of values
int cluster_ffma_debug_read(long long *out, int max_values)
{
    if (!g_cl_dbg) return 0;
    const int n = std::min(max_values, kMaxSteps * 8);
    cudaDeviceSynchronize();
    cudaMemcpy(out, g_cl_dbg, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemset(g_cl_dbg, 0, kMaxSteps * 8 * sizeof(long long));
    return n;
}

cudaError_t launch_cluster_ffma(const KernelArgs &args, int chi_q, int cs, int depth, int clusters, size_t smem, cudaStream_t stream)
{
    ClArgs ca;
    memset(&ca, 0, sizeof ca);
    ca.a = args, ca.chi_q = chi_q;
    if (!g_cl_dbg && getenv("LINNA_CLUSTER_DEBUG")) {
        if (cudaMalloc(&g_cl_dbg, kMaxSteps * 8 * sizeof(long long)) == cudaSuccess) cudaMemset(g_cl_dbg, 0, kMaxSteps * 8 * sizeof(long long));
        else g_cl_dbg = nullptr;
    }
    ca.dbg = g_cl_dbg;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(clusters * cs, 1, 1), cfg.blockDim = dim3(CL_THREADS, 1, 1), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    if (depth == 16) return cudaLaunchKernelEx(&cfg, cluster_ffma_kernel<16>, ca);
    return cudaLaunchKernelEx(&cfg, cluster_ffma_kernel<8>, ca);
}

}  // namespace linna
