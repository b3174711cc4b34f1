// Microbenchmark: can two persistent kernels split the 148 SMs of a B200 between them?
//
// Kernel A ("prologue stand-in"): PA clusters of 2 CTAs, one CTA per SM (large dynamic shared memory), launched
// first on stream 1.  Kernel B ("GEMM stand-in"): (148 - 2*PA)/2 clusters of 2 CTAs, one CTA per SM, launched right
// after on stream 2.  Every CTA records its SM id and its start time (globaltimer) and then spins for `spin_us`.
// If all CTAs of B start within a few microseconds of A's, the two grids run side by side; a CTA of B that
// starts ~spin_us late had to wait for an SM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o partition.bin partition.cu
//   partition.bin [PA clusters] [spin_us] [B clusters (default: the rest)]
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                              \
        }                                                                                         \
    } while (0)

struct Rec {
    unsigned long long t0;
    unsigned smid;
    unsigned pad;
};

__global__ void occupy(Rec* rec, unsigned long long spin_ns) {
    extern __shared__ unsigned char dyn[];
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        rec[blockIdx.x].t0 = t0;
        rec[blockIdx.x].smid = smid;
        dyn[0] = 1;
    }
    unsigned long long t;
    do {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < spin_ns);
}

static void launch(int clusters, int threads, size_t smem, cudaStream_t st, Rec* rec, unsigned long long spin_ns) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * 2);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, occupy, rec, spin_ns));
}

int main(int argc, char** argv) {
    const int PA = argc > 1 ? atoi(argv[1]) : 20;
    const int spin_us = argc > 2 ? atoi(argv[2]) : 200;
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    const int PB = argc > 3 ? atoi(argv[3]) : (nsm - 2 * PA) / 2;
    const size_t smem = 200 * 1024;
    CK(cudaFuncSetAttribute(occupy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    Rec *ra, *rb;
    CK(cudaMalloc(&ra, sizeof(Rec) * 2 * PA));
    CK(cudaMalloc(&rb, sizeof(Rec) * 2 * PB));
    for (int rep = 0; rep < 3; ++rep) {
        launch(PA, 1024, smem, s1, ra, (unsigned long long)spin_us * 1000ull);
        launch(PB, 640, smem, s2, rb, (unsigned long long)spin_us * 1000ull);
        CK(cudaDeviceSynchronize());
        std::vector<Rec> ha(2 * PA), hb(2 * PB);
        CK(cudaMemcpy(ha.data(), ra, sizeof(Rec) * 2 * PA, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hb.data(), rb, sizeof(Rec) * 2 * PB, cudaMemcpyDeviceToHost));
        unsigned long long tmin = ~0ull;
        for (auto& r : ha) tmin = std::min(tmin, r.t0);
        for (auto& r : hb) tmin = std::min(tmin, r.t0);
        int late_a = 0, late_b = 0, split_a = 0, split_b = 0;
        double max_a = 0, max_b = 0;
        for (auto& r : ha) { double d = (r.t0 - tmin) * 1e-3; max_a = std::max(max_a, d); late_a += d > spin_us * 0.5; }
        for (auto& r : hb) { double d = (r.t0 - tmin) * 1e-3; max_b = std::max(max_b, d); late_b += d > spin_us * 0.5; }
        for (int i = 0; i < PA; ++i) split_a += (ha[2 * i].smid >> 1) != (ha[2 * i + 1].smid >> 1);
        for (int i = 0; i < PB; ++i) split_b += (hb[2 * i].smid >> 1) != (hb[2 * i + 1].smid >> 1);
        printf("rep %d: A %d clusters (max start %.1f us, %d late CTAs, %d clusters not on SM pair 2k/2k+1)  "
               "B %d clusters (max start %.1f us, %d late CTAs, %d clusters not on SM pair 2k/2k+1)\n",
               rep, PA, max_a, late_a, split_a, PB, max_b, late_b, split_b);
        if (rep == 2) {
            printf("A smids:");
            for (auto& r : ha) printf(" %u", r.smid);
            printf("\n");
        }
    }
    return 0;
}
