// Shared host/device helpers for libpicopose_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "../../include/picopose_b200.h"

namespace pp {

// ---- host-side error plumbing -------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define PP_CHECK_ARG(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) return ::pp::fail(PP_ERR_ARG, __VA_ARGS__);  \
    } while (0)

#define PP_CUDA(call)                                                                            \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return ::pp::fail(PP_ERR_LAUNCH, "%s failed: %s (%s:%d)", #call,                     \
                              cudaGetErrorString(e__), __FILE__, __LINE__);                      \
    } while (0)

// Kernel-launch accounting (pp_launch_count) and the optional event pair recorded around the next
// tensor-core GEMM launch (pp_profile_gemm_events), used by bench.py for the roofline figure.
void count_launch();
bool take_profile_events(cudaEvent_t* start, cudaEvent_t* stop);

#define PP_LAUNCHED()                 \
    do {                              \
        PP_CUDA(cudaGetLastError());  \
        ::pp::count_launch();         \
    } while (0)

// Query-side bookkeeping produced by pp_match_prepare_query (see mask_compact_kernel).
// q_meta layout (4-byte elements): mrow[B*T] | rank[B*T] | rowmap[B*T] | tv[B] | fm[B]
struct QueryMeta {
    float* mrow;   // (B,T) nearest-resized query mask
    int* rank;     // (B,T) compact row of patch t, -1 if masked
    int* rowmap;   // (B,T) patch of compact row r
    int* tv;       // (B)   number of unmasked patches
    int* fm;       // (B)   first masked patch, -1 if none
};
QueryMeta split_query_meta(void* q_meta, int B, int T);

// Host-mapped fault record shared by all kernels (5 ints: code, block, thread, aux0, aux1); created on first use.
// codes: 1-4 contraction pipeline time-outs (the kernel traps), 5 exchange peer missing, 6 bank index out of range.
int fault_buffer(int** dev_ptr);

// 1x1 convolution applied to the block's lookup tile while it is still in shared memory: the first
// layer of the reference's MotionEncoder.corr_net (model/stage3/raft_decoder.py:113-116,127-129,157), corr_inch = L*D*D ->
// cout channels (+ bias, + ReLU), i.e. a (cout x L*D*D) matrix applied per query.
struct WConv {
    const float* weight = nullptr;  // (cout, L*D*D) row-major = Conv2d.weight[:, :, 0, 0]
    const float* bias = nullptr;    // (cout) or null
    float* out = nullptr;           // (N, cout, H, W)
    int cout = 0;
    int relu = 0;
};

// Verifies the current device is sm_100 (cached per device).
int require_sm100();
int sm_count();

// Programmatic dependent launch: a kernel launched through launch_dependent() may be scheduled while the kernel before it
// on the stream is still draining, so its launch latency and its own prologue overlap that tail; the data dependency is
// kept by the kernel itself, which calls grid_dependency_wait() before it touches anything the predecessor wrote (a no-op
// when there is no programmatic edge).  PICOPOSE_B200_PDL=0 launches plainly.
bool pdl_enabled();
// An early trigger (griddepcontrol.launch_dependents in front of the wait, so that the next kernel is scheduled during
// this grid's last wave and not only at its exit) was measured on the configs[1] step: 0.3557 ms against 0.3540 ms, three
// alternating runs each on one box (profiles/r2i_pdl_early_trigger_ab.jsonl) -- the early blocks hold SM resources while
// they wait.  Nobody triggers early.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- small device helpers -----------------------------------------------------------------
// Monotone map float -> uint32 (a < b  <=>  ord(a) < ord(b)); -0.0 must be canonicalised first.
__device__ __forceinline__ uint32_t f32_ord(float v) {
    uint32_t u = __float_as_uint(v);
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord_f32(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
    return __uint_as_float(u);
}
// (value, index) -> 64-bit key whose unsigned max is "largest value, then smallest index".
__device__ __forceinline__ unsigned long long pack_key(float v, uint32_t idx) {
    return ((unsigned long long)f32_ord(v) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ float key_value(unsigned long long k) { return ord_f32((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_index(unsigned long long k) { return 0xFFFFFFFFu - (uint32_t)k; }

// F.interpolate(mode='nearest') source index: min(floor(dst * in/out), in - 1), computed in fp32 like ATen.
__host__ __device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size) {
    float scale = (float)in_size / (float)out_size;
    int s = (int)floorf((float)dst * scale);
    return s < in_size - 1 ? s : in_size - 1;
}

}  // namespace pp
