// Microbenchmark: how much HBM bandwidth can a SUBSET of the SMs pull?
//
// P CTAs (one per SM: large dynamic shared memory), each streams its share of `in` (fp32, 680 MB) through a ring of
// shared-memory stages with cp.async.bulk and writes half as many bytes back out with cp.async.bulk stores -- the
// traffic shape of the stage-1 bank prologue (4 bytes read + 2 bytes written per element).  Reports GB/s against P,
// to decide whether the prologue could run on a few SMs beside the tensor-core kernel.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o subset_stream.bin subset_stream.cu
//   subset_stream.bin [stride]   (CTA i works iff i % stride == 0; the grid is always 148; no argument = sweep)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                              \
        }                                                                                         \
    } while (0)

constexpr int CHUNK = 8192;  // bytes per stage
constexpr int STAGES = 24;   // 192 KB in flight per SM

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

// workers = CTAs with blockIdx % stride == 0; chunk c of the input goes to worker (c % workers)
__global__ void __launch_bounds__(64) stream_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t nchunks,
                                                    int stride, int write_out) {
    extern __shared__ __align__(128) uint8_t ring[];
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
    if (blockIdx.x % stride != 0) return;
    const size_t worker = blockIdx.x / stride, workers = (gridDim.x + stride - 1) / stride;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // producer
        int s = 0;
        uint32_t ph = 0;
        for (size_t c = worker; c < nchunks; c += workers) {
            mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(ring + s * CHUNK)),
                         "l"(in + c * CHUNK), "r"(CHUNK), "r"(smem_u32(&full[s]))
                         : "memory");
            if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
    } else if (threadIdx.x == 32) {
        // consumer: store half of each chunk, then free the stage
        int s = 0;
        uint32_t ph = 0;
        for (size_t c = worker; c < nchunks; c += workers) {
            mbar_wait(smem_u32(&full[s]), ph);
            if (write_out) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * (CHUNK / 2)),
                             "r"(smem_u32(ring + s * CHUNK)), "r"(CHUNK / 2)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
            if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
    }
}

int main(int argc, char** argv) {
    const size_t bytes = 162ull * 1024 * 1024 * 4;  // the config-2 bank
    const size_t nchunks = bytes / CHUNK;
    uint8_t *in, *out;
    CK(cudaMalloc(&in, bytes));
    CK(cudaMalloc(&out, bytes / 2));
    CK(cudaMemset(in, 1, bytes));
    const size_t smem = (size_t)STAGES * CHUNK + 1024;
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int strides[] = {1, 2, 3, 4, 5, 6, 8};
    for (int wr = 1; wr >= 0; --wr)
        for (int stride : strides) {
            if (argc > 1 && atoi(argv[1]) != stride) continue;
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                stream_kernel<<<148, 64, smem>>>(in, out, nchunks, stride, wr);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep == 1)
                    printf("workers %3d (every %d. SM)  %s  %8.1f us  read %6.0f GB/s  read+write %6.0f GB/s\n",
                           (148 + stride - 1) / stride, stride, wr ? "read+write" : "read only ", ms * 1e3, bytes / ms * 1e-6,
                           (wr ? 1.5 : 1.0) * bytes / ms * 1e-6);
            }
        }
    return 0;
}
