// Microbenchmark: how fast can a few SMs stream column slices of an L2-resident fp32 matrix?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2stream l2stream.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ void cp16(unsigned s, const float *g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory"); }
__device__ __forceinline__ void cp16ca(unsigned s, const float *g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory"); }
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void waitg() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// variant 0: per-thread cp.async ring of depth D; thread (q, s): quad q of Q, k-lane s of S
template <int D, int MODE>
__global__ void __launch_bounds__(256, 1) k_ring(const float *W, int K, int ldw, int cols_per_cta, int nmat, size_t mat_stride, float *out, long long *cyc)
{
    extern __shared__ float4 sm[];
    const int tid = threadIdx.x;
    const int Q = cols_per_cta / 4, S = 256 / Q, q = tid % Q, s = tid / Q;
    const int col0 = blockIdx.x * cols_per_cta;
    const unsigned ring = (unsigned)__cvta_generic_to_shared(sm) + tid * 16;
    float acc[4] = {0, 0, 0, 0};
    const long long t0 = clock64();
    if (MODE == 0) {
        // stream position
        int pm = 0, pk = s;
        auto issue = [&](int slot) {
            if (pm < nmat) {
                cp16(ring + slot * 4096, W + pm * mat_stride + (size_t)pk * ldw + col0 + 4 * q);
                pk += S;
                if (pk >= K) pk = s, ++pm;
            }
            commit();
        };
        for (int i = 0; i < D; ++i) issue(i);
        unsigned cons = 0;
        for (int m = 0; m < nmat; ++m)
            for (int k = s; k < K; k += S) {
                waitg<D - 1>();
                const unsigned slot = cons & (D - 1);
                float4 w = sm[slot * 256 + tid];
                acc[0] += w.x, acc[1] += w.y, acc[2] += w.z, acc[3] += w.w;
                issue(slot);
                ++cons;
            }
    } else if (MODE == 2 || MODE == 3) {
        float a32[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) a32[e] = 0.f;
        float4 *act = sm + D * 256;   // [K][2] float4 = K x 8 rows
        for (int i = tid; i < 2 * K; i += 256) act[i] = make_float4(1.f, 2.f, 3.f, 4.f);
        __syncthreads();
        int pm = 0, pk = s;
        auto issue = [&](int slot) {
            if (MODE == 2 && pm < nmat) {
                cp16(ring + slot * 4096, W + pm * mat_stride + (size_t)pk * ldw + col0 + 4 * q);
                pk += S;
                if (pk >= K) pk = s, ++pm;
            }
            commit();
        };
        for (int i = 0; i < D; ++i) issue(i);
        unsigned cons = 0;
        for (int m = 0; m < nmat; ++m)
            for (int k = s; k < K; k += S) {
                waitg<D - 1>();
                const unsigned slot = cons & (D - 1);
                float4 w = sm[slot * 256 + tid];
                float4 a0 = act[2 * k], a1 = act[2 * k + 1];
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                    for (int r = 0; r < 8; ++r) a32[8 * cc + r] = fmaf(a[r], b[cc], a32[8 * cc + r]);
                issue(slot);
                ++cons;
            }
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e & 3] += a32[e];
    } else if (MODE == 1) {
        // direct register loads, unrolled by 8
        for (int m = 0; m < nmat; ++m) {
            const float *p = W + m * mat_stride + col0 + 4 * q;
            int k = s;
            for (; k + 7 * S < K; k += 8 * S) {
                float4 w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = __ldcg(reinterpret_cast<const float4 *>(p + (size_t)(k + j * S) * ldw));
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[0] += w[j].x, acc[1] += w[j].y, acc[2] += w[j].z, acc[3] += w[j].w;
            }
            for (; k < K; k += S) {
                float4 w = __ldcg(reinterpret_cast<const float4 *>(p + (size_t)k * ldw));
                acc[0] += w.x, acc[1] += w.y, acc[2] += w.z, acc[3] += w.w;
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 256 + tid] = acc[0] + acc[1] + acc[2] + acc[3];
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

int main(int argc, char **argv)
{
    const int K = 1000, N = 500, ldw = 500, nmat = 8;
    const size_t mat = (size_t)K * ldw;
    float *W, *out;
    long long *cyc;
    cudaMalloc(&W, nmat * mat * 4 + 4096);
    cudaMemset(W, 0, nmat * mat * 4 + 4096);
    cudaMalloc(&out, 148 * 256 * 4);
    cudaMalloc(&cyc, 148 * 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    int cluster = 1;
    auto run = [&](const char *name, auto kern, int ctas, int cpc, int smem) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(ctas), cfg.blockDim = dim3(256), cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cluster, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
        cfg.attrs = at, cfg.numAttrs = 1;
        for (int i = 0; i < 3; ++i) cudaLaunchKernelEx(&cfg, kern, (const float *)W, K, ldw, cpc, nmat, mat, out, cyc);
        cudaEventRecord(a);
        for (int i = 0; i < 10; ++i) cudaLaunchKernelEx(&cfg, kern, (const float *)W, K, ldw, cpc, nmat, mat, out, cyc);
        cudaEventRecord(b);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        long long h[148];
        cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
        const double bytes_cta = (double)nmat * K * cpc * 4;
        printf("%-28s ctas=%3d cpc=%3d: %.1f us/launch, cta0 %lld cycles, %.1f B/clk/CTA, total %.1f GB/s  err=%s\n", name, ctas, cpc, ms * 100, h[0],
               bytes_cta / h[0], bytes_cta * ctas / (ms / 10 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    run("ring16 + act LDS + 32 FFMA", k_ring<16, 2>, 16, 32, 16 * 4096 + 32768);
    run("no loads: LDS + 32 FFMA", k_ring<16, 3>, 16, 32, 16 * 4096 + 32768);
    for (int cl : {16}) {
        cluster = cl;
        printf("cluster size %d\n", cl);
        run("ring D=16", k_ring<16, 0>, 16, 32, 16 * 4096);
        run("ldg x8", k_ring<8, 1>, 16, 32, 4096);
        run("ring D=16 x2 clusters", k_ring<16, 0>, 32, 16, 16 * 4096);
    }
    cluster = 1;
    for (int ctas : {16}) {
        const int cpc = 32;
        run("ring D=8", k_ring<8, 0>, ctas, cpc, 8 * 4096);
        run("ring D=16", k_ring<16, 0>, ctas, cpc, 16 * 4096);
        run("ring D=32", k_ring<32, 0>, ctas, cpc, 32 * 4096);
        run("ldg x8", k_ring<8, 1>, ctas, cpc, 4096);
    }
    run("ring D=16 cpc=64 8 ctas", k_ring<16, 0>, 7, 64, 16 * 4096);
    run("ring D=32 cpc=64 8 ctas", k_ring<32, 0>, 7, 64, 32 * 4096);
    run("ring D=16 cpc=4 125 ctas", k_ring<16, 0>, 125, 4, 16 * 4096);
    return 0;
}
