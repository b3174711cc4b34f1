// Microbenchmark: how many DRAM bytes does a B200 move for the lookup's access pattern, per load path?
//
// Pattern = the stage-3 corr lookup at r=4 (config 4): Q independent 64x64 fp32 slices (16 KB each), one
// 10-row x 10-float window per slice at a random origin.  Each variant fetches exactly the 16-byte pieces
// that cover the window and folds them into a checksum.  tools/microbench/run.sh runs it plain and under
//   ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,... 
// to read DRAM bytes per query next to the 400 useful bytes (results: profiles/r1w_dram_granularity.md).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o dram_gran.bin dram_gran.cu
//   dram_gran.bin [l2_fetch_granularity(0=leave)] [Q] [pieces|like|param|bulk|all] [zero|random]
//
// Caveat found with it: cp.async copies that target a shared-memory address which an earlier, still
// pending cp.async of the same warp also targets are partly squashed (fewer sectors are requested), so
// the "ldgsts_cg_16" line of `pieces` and the same_dst=1 lines of `param` under-count; `like` and
// same_dst=0 are the valid cp.async measurements.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                              \
        }                                                                                         \
    } while (0)

constexpr int MAPW = 64, MAPH = 64, WIN = 10;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ void origin(uint32_t q, int& x0, int& y0) {
    const uint32_t h = hash32(q * 2654435761u + 12345u);
    x0 = (int)(h % (MAPW - WIN + 1));
    y0 = (int)((h >> 16) % (MAPH - WIN + 1));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

enum { V_LDG_NC = 0, V_LDG_CG, V_LDG_CV, V_LDG_LU, V_LDG_EVICT_FIRST, V_LDG_NOALLOC, V_LDG_SCALAR, V_LDGSTS, V_COUNT };
static const char* kNames[] = {"ldg_nc_v4", "ldg_cg_v4", "ldg_cv_v4", "ldg_lu_v4", "ldg_evict_first_v4", "ldg_L1noalloc_v4",
                               "ldg_nc_scalar", "ldgsts_cg_16"};

template <int V>
__device__ __forceinline__ float4 load16(const float* p, uint64_t pol) {
    float4 v;
    if (V == V_LDG_NC) asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (V == V_LDG_CG) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (V == V_LDG_CV) asm volatile("ld.global.cv.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (V == V_LDG_LU) asm volatile("ld.global.lu.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (V == V_LDG_EVICT_FIRST)
        asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    if (V == V_LDG_NOALLOC)
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// lane -> (row, 16-byte piece) of a warp-uniform query's window; 8 rows per pass
template <int V>
__global__ void __launch_bounds__(128) gather_pieces(const float* __restrict__ vol, uint32_t Q, float* __restrict__ sink) {
    __shared__ __align__(16) float stage[4][32 * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * 4 + warp, nw = gridDim.x * 4;
    uint64_t pol = 0;
    if (V == V_LDG_EVICT_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    float acc = 0.f;
    for (uint32_t q = gw; q < Q; q += nw) {
        int x0, y0;
        origin(q, x0, y0);
        const float* slice = vol + (size_t)q * (MAPW * MAPH);
        if (V == V_LDG_SCALAR) {
            // lane -> element: 100 elements in 4 passes
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                const int e = ps * 32 + lane;
                const int row = e / WIN, col = e - row * WIN;
                if (e < WIN * WIN) acc += __ldg(slice + (y0 + row) * MAPW + x0 + col);
            }
            continue;
        }
        const int xs = x0 & ~3;
        const int pieces = ((x0 + WIN - 1 - xs) >> 2) + 1;
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
            const int row = ps * 8 + (lane >> 2), v = lane & 3;
            if (row < WIN && v < pieces) {
                const float* src = slice + (y0 + row) * MAPW + xs + 4 * v;
                if (V == V_LDGSTS) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&stage[warp][lane * 4])), "l"(src) : "memory");
                } else {
                    const float4 t = load16<V>(src, pol);
                    acc += t.x + t.y + t.z + t.w;
                }
            }
        }
        if (V == V_LDGSTS) {
            // keep a few groups in flight: wait only every 4th query
            if (((q - gw) / nw & 3u) == 3u) {
                asm volatile("cp.async.wait_all;" ::: "memory");
                acc += stage[warp][lane * 4];
            }
        }
    }
    if (V == V_LDGSTS) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        acc += stage[warp][lane * 4];
    }
    if (acc == 123456.789f) sink[0] = acc;
}


// the lookup kernel's structure: a warp owns 32 consecutive queries, stages their windows one query per
// instruction (lane -> row/piece), waits, then writes 81 coalesced output lines (lane = query)
template <bool ZFILL, bool WRITES, bool READBACK>
__global__ void __launch_bounds__(128) gather_like_lookup(const float* __restrict__ vol, uint32_t Q, float* __restrict__ out,
                                                          float* __restrict__ sink, uint32_t srcsize) {
    extern __shared__ __align__(16) float dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* stage = dyn + warp * (32 * 164);  // 41 pieces (odd) per query
    const uint32_t gw = blockIdx.x * 4 + warp, nw = gridDim.x * 4;
    float acc = 0.f;
    for (uint32_t g = gw; g * 32 < Q; g += nw) {
        const uint32_t qb = g * 32;
        __syncwarp();
        for (int ql = 0; ql < 32; ++ql) {
            int x0, y0;
            origin(qb + ql, x0, y0);
            const float* slice = vol + (size_t)(qb + ql) * (MAPW * MAPH);
            const int xs = x0 & ~3;
            const int pieces = ((x0 + WIN - 1 - xs) >> 2) + 1;
#pragma unroll
            for (int ps = 0; ps < 2; ++ps) {
                const int row = ps * 8 + (lane >> 2), v = lane & 3;
                if (row < WIN && v < pieces) {
                    const float* src = slice + (y0 + row) * MAPW + xs + 4 * v;
                    const uint32_t dst = smem_u32(stage + ql * 164 + (row * 4 + v) * 4);
                    if (ZFILL) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(srcsize) : "memory");
                    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        if (READBACK) {
            for (int c = 0; c < 81; ++c) {
                const float v = stage[lane * 164 + (c / 9) * 16 + (c % 9)];
                if (WRITES) __stcs(out + (size_t)c * Q + qb + lane, v);
                else acc += v;
            }
        } else if (WRITES) {
            for (int c = 0; c < 81; ++c) __stcs(out + (size_t)c * Q + qb + lane, 1.0f);
        } else {
            acc += stage[lane * 164];
        }
    }
    if (acc == 123456.789f) sink[0] = acc;
}


// parametrised staging: G queries in flight per warp before the wait, three warp->query mappings, two
// shared-memory destination modes
__global__ void __launch_bounds__(128) gather_param(const float* __restrict__ vol, uint32_t Q, float* __restrict__ sink, int G,
                                                    int mapping, int same_dst) {
    extern __shared__ __align__(16) float dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* stage = dyn + warp * (32 * 164);
    const uint32_t gw = blockIdx.x * 4 + warp, nw = gridDim.x * 4;
    const uint32_t per_warp = (Q + nw - 1) / nw;
    float acc = 0.f;
    for (uint32_t it = 0; it * G < per_warp; ++it) {
        __syncwarp();
        for (int j = 0; j < G; ++j) {
            const uint32_t sidx = it * G + j;
            uint32_t q;
            if (mapping == 0) q = gw * per_warp + sidx;           // contiguous chunk per warp
            else if (mapping == 1) q = sidx * nw + gw;            // strided: neighbouring warps take neighbouring queries
            else q = (it * nw + gw) * G + j;                      // block-cyclic, block = G queries
            if (sidx >= per_warp || q >= Q) continue;
            int x0, y0;
            origin(q, x0, y0);
            const float* slice = vol + (size_t)q * (MAPW * MAPH);
            const int xs = x0 & ~3;
            const int pieces = ((x0 + WIN - 1 - xs) >> 2) + 1;
#pragma unroll
            for (int ps = 0; ps < 2; ++ps) {
                const int row = ps * 8 + (lane >> 2), v = lane & 3;
                if (row < WIN && v < pieces) {
                    const float* src = slice + (y0 + row) * MAPW + xs + 4 * v;
                    const uint32_t dst = smem_u32(stage + (same_dst ? 0 : j * 164) + (row * 4 + v) * 4);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        acc += stage[lane * 4];
    }
    if (acc == 123456.789f) sink[0] = acc;
}

__global__ void fill_random(float* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = (float)(hash32((uint32_t)i) & 0xFFFFFF) * (1.0f / 16777216.0f) - 0.5f;
}

// one 1-D bulk copy (48 or 64 bytes) per window row
__global__ void __launch_bounds__(128) gather_bulk_rows(const float* __restrict__ vol, uint32_t Q, float* __restrict__ sink) {
    __shared__ __align__(128) uint8_t slots[4][32][64];
    __shared__ __align__(8) uint64_t bars[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t bar = smem_u32(&bars[warp]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint32_t gw = blockIdx.x * 4 + warp, nw = gridDim.x * 4;
    uint32_t phase = 0;
    float acc = 0.f;
    // 3 queries x 10 rows = 30 lanes per step
    for (uint32_t q0 = gw * 3; q0 < Q; q0 += nw * 3) {
        const int qi = lane / WIN, row = lane - qi * WIN;
        const bool on = lane < 30 && q0 + qi < Q;
        int x0 = 0, y0 = 0;
        if (on) origin(q0 + qi, x0, y0);
        const int xs = x0 & ~3;
        const uint32_t bytes = on ? (uint32_t)((((x0 + WIN - 1 - xs) >> 2) + 1) * 16) : 0u;
        uint32_t total = bytes;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
        __syncwarp();
        if (on) {
            const float* src = vol + (size_t)(q0 + qi) * (MAPW * MAPH) + (y0 + row) * MAPW + xs;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(&slots[warp][lane][0])),
                         "l"(src), "r"(bytes), "r"(bar)
                         : "memory");
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done) : "r"(bar), "r"(phase) : "memory");
        }
        phase ^= 1u;
        acc += *reinterpret_cast<const float*>(&slots[warp][lane][0]);
        __syncwarp();
    }
    if (acc == 123456.789f) sink[0] = acc;
}

template <typename F>
static void timed(const char* name, uint32_t Q, F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();  // warm-up (the 4 GB volume is far larger than L2, so every launch is cold in DRAM terms)
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-28s %8.3f us  %7.1f Mquery/s  useful %6.1f GB/s\n", name, ms * 1e3, Q / ms * 1e-3, Q * 400.0 / ms * 1e-6);
    fflush(stdout);
}

int main(int argc, char** argv) {
    const int gran = argc > 1 ? atoi(argv[1]) : 0;
    const uint32_t Q = argc > 2 ? (uint32_t)atoi(argv[2]) : 262144u;
    if (gran) {
        CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran));
    }
    size_t now = 0;
    CK(cudaDeviceGetLimit(&now, cudaLimitMaxL2FetchGranularity));
    printf("cudaLimitMaxL2FetchGranularity = %zu (requested %d), Q = %u\n", now, gran, Q);
    float* vol;
    float* sink;
    CK(cudaMalloc(&vol, (size_t)Q * MAPW * MAPH * 4));
    CK(cudaMalloc(&sink, 256));
    CK(cudaMemset(vol, 0, (size_t)Q * MAPW * MAPH * 4));
    const int grid = 148 * 8;
    const char* sel = argc > 3 ? argv[3] : "all";
    auto want = [&](const char* k) { return !strcmp(sel, "all") || !strcmp(sel, k); };

#define RUN(V) timed(kNames[V], Q, [&] { gather_pieces<V><<<grid, 128>>>(vol, Q, sink); });
    float* outbuf;
    CK(cudaMalloc(&outbuf, (size_t)Q * 81 * 4));
    if (argc > 4 && !strcmp(argv[4], "random")) {
        fill_random<<<148 * 8, 256>>>(vol, (size_t)Q * MAPW * MAPH);
        CK(cudaDeviceSynchronize());
        printf("volume filled with random floats\n");
    }
    if (want("like")) {
        const int sm = 4 * 32 * 164 * 4;
        CK(cudaFuncSetAttribute(gather_like_lookup<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        CK(cudaFuncSetAttribute(gather_like_lookup<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        CK(cudaFuncSetAttribute(gather_like_lookup<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        CK(cudaFuncSetAttribute(gather_like_lookup<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        CK(cudaFuncSetAttribute(gather_like_lookup<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        timed("like_plain", Q, [&] { gather_like_lookup<false, false, false><<<grid, 128, sm>>>(vol, Q, outbuf, sink, 16); });
        timed("like_zfill", Q, [&] { gather_like_lookup<true, false, false><<<grid, 128, sm>>>(vol, Q, outbuf, sink, 16); });
        timed("like_writes", Q, [&] { gather_like_lookup<false, true, false><<<grid, 128, sm>>>(vol, Q, outbuf, sink, 16); });
        timed("like_readback", Q, [&] { gather_like_lookup<false, false, true><<<grid, 128, sm>>>(vol, Q, outbuf, sink, 16); });
        timed("like_zfill_writes_readback", Q, [&] { gather_like_lookup<true, true, true><<<grid, 128, sm>>>(vol, Q, outbuf, sink, 16); });
    }
    if (want("param")) {
        const int sm = 4 * 32 * 164 * 4;
        CK(cudaFuncSetAttribute(gather_param, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        const int Gs[4] = {1, 4, 8, 32};
        for (int mapping = 0; mapping < 3; ++mapping)
            for (int gi = 0; gi < 4; ++gi)
                for (int same = 0; same < 2; ++same) {
                    char name[64];
                    snprintf(name, sizeof name, "param_map%d_G%d_same%d", mapping, Gs[gi], same);
                    timed(name, Q, [&] { gather_param<<<grid, 128, sm>>>(vol, Q, sink, Gs[gi], mapping, same); });
                }
    }
    if (want("pieces")) {
    RUN(V_LDG_NC) RUN(V_LDG_CG) RUN(V_LDG_CV) RUN(V_LDG_LU) RUN(V_LDG_EVICT_FIRST) RUN(V_LDG_NOALLOC) RUN(V_LDG_SCALAR) RUN(V_LDGSTS)
    CK(cudaGetLastError());
    }

    if (want("bulk")) timed("bulk_rows_48_64B", Q, [&] { gather_bulk_rows<<<grid, 128>>>(vol, Q, sink); });
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
