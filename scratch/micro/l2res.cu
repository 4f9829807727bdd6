// Is a 4 MB weight matrix re-read from L2 or from DRAM by consecutive launches?  Same kernel, (a) the same 4 MB every
// launch, (b) rotating over 48 different 4 MB regions (192 MB > L2).
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
__device__ __forceinline__ void cp16(unsigned s, const float *g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory"); }
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void waitg() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int D>
__global__ void __launch_bounds__(256, 1) k(const float *W, int K, int ldw, int cpc, float *out, long long *cyc)
{
    extern __shared__ float4 sm[];
    const int tid = threadIdx.x;
    const int Q = cpc / 4, S = 256 / Q, q = tid % Q, s = tid / Q, col0 = blockIdx.x * cpc;
    const unsigned ring = (unsigned)__cvta_generic_to_shared(sm) + tid * 16;
    float acc = 0;
    const long long t0 = clock64();
    int pk = s;
    auto issue = [&](int slot) {
        if (pk < K) { cp16(ring + slot * 4096, W + (size_t)pk * ldw + col0 + 4 * q); pk += S; }
        commit();
    };
    for (int i = 0; i < D; ++i) issue(i);
    unsigned cons = 0;
    for (int kk = s; kk < K; kk += S) {
        waitg<D - 1>();
        const unsigned slot = cons & (D - 1);
        float4 w = sm[slot * 256 + tid];
        acc += w.x + w.y + w.z + w.w;
        issue(slot);
        ++cons;
    }
    out[blockIdx.x * 256 + tid] = acc;
    if (tid == 0) cyc[blockIdx.x] = clock64() - t0;
}
int main()
{
    const int K = 2000, ldw = 512, ctas = 16, cpc = 32, nreg = 48;
    const size_t reg = (size_t)K * ldw;   // 4 MB
    float *W, *out; long long *cyc;
    cudaMalloc(&W, nreg * reg * 4); cudaMemset(W, 0, nreg * reg * 4);
    cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(k<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 2; ++mode) {
        for (int i = 0; i < 5; ++i) k<16><<<ctas, 256, 65536>>>(W, K, ldw, cpc, out, cyc);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        for (int i = 0; i < 96; ++i) k<16><<<ctas, 256, 65536>>>(W + (mode ? (size_t)(i % nreg) * reg : 0), K, ldw, cpc, out, cyc);
        cudaEventRecord(b); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%s: %.2f us/launch, cta0 %lld cycles, %.1f B/clk/CTA\n", mode ? "rotating over 192 MB (DRAM)" : "same 4 MB (L2?)      ", ms / 96 * 1e3, h,
               (double)K * cpc * 4 / h);
    }
    return 0;
}
