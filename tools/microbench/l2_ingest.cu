// How many bytes per clock can one SM pull out of L2 with bulk async copies while all 148 SMs do the same?
// (sizing question for a contraction that would stream fp32 template features instead of prepared bf16)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int STAGES>
__global__ void __launch_bounds__(128, 1) ingest(const char* __restrict__ src, size_t footprint, int chunk, int iters, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t nchunks = footprint / chunk;
    size_t c = ((size_t)blockIdx.x * 977) % nchunks;
    const long long t0 = clock64();
    for (int i = 0; i < iters + STAGES; ++i) {
        const int s = i % STAGES;
        if (i >= STAGES) {
            const uint32_t parity = ((i / STAGES) - 1) & 1;
            uint32_t done = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(parity) : "memory");
            }
        }
        if (i < iters) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem + (size_t)s * chunk)), "l"(src + c * chunk), "r"(chunk), "r"(smem_u32(&bar[s])) : "memory");
            c += 148 * 3 + 1;
            if (c >= nchunks) c -= nchunks * (c / nchunks);
        }
    }
    cycles[blockIdx.x] = clock64() - t0;
}

int main(int argc, char** argv) {
    const size_t mb = argc > 1 ? atoi(argv[1]) : 48;
    const int chunk = argc > 2 ? atoi(argv[2]) : 32768;
    const int iters = argc > 3 ? atoi(argv[3]) : 4000;
    const size_t footprint = mb << 20;
    char* src;
    cudaMalloc(&src, footprint);
    cudaMemset(src, 1, footprint);
    long long* cyc;
    cudaMalloc(&cyc, 148 * 8);
    constexpr int ST = 4;
    cudaFuncSetAttribute(ingest<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * chunk + 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        ingest<ST><<<148, 128, ST * chunk + 1024>>>(src, footprint, chunk, iters, cyc);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += h[i];
        avg /= 148;
        const double bytes = (double)148 * iters * chunk;
        printf("footprint %zu MiB chunk %d: %.3f ms, %.2f TB/s aggregate, %.1f B/clk/SM (avg %.0f cycles)\n", mb, chunk, ms,
               bytes / ms * 1e-9, (double)iters * chunk / avg, avg);
    }
    return 0;
}
