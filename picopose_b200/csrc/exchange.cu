// Multi-GPU top-k exchange over NVLink peer memory (one node, one process per GPU).
//
// The template bank is sharded by views, so torch.topk over all views (utils/matching.py:68) becomes: local top-k per
// rank, exchange of the (B, k) candidate pairs, identical merge on every rank.  This file does all three in ONE kernel:
// block b ranks detection b's local scores, stores its k (score, global index) pairs straight into every peer's
// exchange buffer (plain stores through the NVLink peer mapping), releases a per-(rank, detection) flag on every peer,
// waits for the peers' flags of the same detection and merges the world*k candidates.  Nothing but detection b's own
// flags is waited for, so there is no grid-wide or host-side synchronisation and a call costs one launch instead of
// local top-k kernel + NCCL all-gather + merge kernel.
//
// Exchange buffer of a rank (allocated here with cudaMalloc so that it can be exported with cudaIpcGetMemHandle):
//   flags [2][world][max_b] u32        flag value = epoch of the call that wrote the slot (never 0)
//   pairs [2][world][max_b][k_max]     {float score, int64 index}, 16 bytes each
// The leading [2] is the epoch parity: a rank may run at most one call ahead of a peer (its merge of call e needs the
// peer's flags of call e, which the peer writes only after finishing call e-1 in stream order), so two slot sets
// never collide.
#include "pp_common.cuh"

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace pp {

struct XPair {
    float score;
    int pad;
    long long index;
};

struct XchgLayout {
    size_t flags_bytes, pairs_off, total;
};
static XchgLayout xchg_layout(int world, int max_b, int k_max) {
    XchgLayout l;
    l.flags_bytes = ((size_t)2 * world * max_b * sizeof(unsigned) + 255) / 256 * 256;
    l.pairs_off = l.flags_bytes;
    l.total = l.pairs_off + (size_t)2 * world * max_b * k_max * sizeof(XPair);
    return l;
}

// Waiting for a peer is bounded in wall time (PICOPOSE_B200_XCHG_TIMEOUT_S, default 120 s, 0 = wait for ever, like NCCL).
// A rank that gives up does NOT trap -- the CUDA context and the healthy ranks survive: it records fault code 5 in the
// host-mapped fault buffer (pp_check_device_faults reports it), poisons its outputs (NaN scores, index -1) and returns.
// Flags are compared wrap-safe with >= , so a peer that (against the one-step-ahead contract) has already overwritten a
// slot with a later epoch cannot dead-lock the waiter.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool epoch_reached(unsigned seen, unsigned epoch) { return (int)(seen - epoch) >= 0 && seen != 0u; }

__device__ __noinline__ void record_xchg_fault(int* fault, int peer, unsigned epoch) {
    if (fault) {
        fault[1] = blockIdx.x;
        fault[2] = threadIdx.x;
        fault[3] = peer;
        fault[4] = (int)epoch;
        __threadfence_system();
        fault[0] = 5;
        __threadfence_system();
    }
}

static unsigned long long xchg_timeout_ns() {
    static long long cached = -1;
    if (cached < 0) {
        const char* e = getenv("PICOPOSE_B200_XCHG_TIMEOUT_S");
        double s = e ? atof(e) : 120.0;
        if (!(s >= 0.0)) s = 120.0;
        cached = (long long)(s * 1e9);
    }
    return (unsigned long long)cached;
}

__global__ void __launch_bounds__(256)
topk_exchange_kernel(const float* __restrict__ scores, int N, int k, long long idx_offset, char* const* __restrict__ peers,
                     int rank, int world, int max_b, int k_max, size_t pairs_off, unsigned epoch, int b0,
                     unsigned long long timeout_ns, int* __restrict__ fault,
                     float* __restrict__ out_score, long long* __restrict__ out_idx) {
    extern __shared__ unsigned long long s_keys[];  // N packed keys + 8 partials
    unsigned long long* s_red = s_keys + N;
    __shared__ XPair s_loc[256];
    __shared__ unsigned long long s_merge[1024];
    __shared__ int s_gave_up;
    const int b = b0 + blockIdx.x;
    if (threadIdx.x == 0) s_gave_up = 0;
    const int par = (int)(epoch & 1u);
    // ---- local top-k (torch.topk on this rank's views; ties resolve to the lowest index) ----
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_keys[i] = pack_key(scores[(size_t)b * N + i] + 0.0f, (uint32_t)i);
    __syncthreads();
    const int kl = k < N ? k : N;
    for (int r = 0; r < k; ++r) {
        if (r >= kl) {  // the shard holds fewer than k views: padding that loses every comparison
            if (threadIdx.x == 0) s_loc[r] = XPair{-INFINITY, 0, -1};
            continue;
        }
        unsigned long long best = 0ull;
        for (int i = threadIdx.x; i < N; i += blockDim.x) best = s_keys[i] > best ? s_keys[i] : best;
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = s_red[w] > best ? s_red[w] : best;
            const uint32_t idx = key_index(best);
            s_loc[r] = XPair{key_value(best), 0, (long long)idx + idx_offset};
            s_keys[idx] = 0ull;
        }
        __syncthreads();
    }
    __syncthreads();
    // ---- push: pair j of this rank's slot in every peer's buffer (its own included) ----
    for (int i = threadIdx.x; i < world * k; i += blockDim.x) {
        const int p = i / k, j = i - p * k;
        XPair* dst = reinterpret_cast<XPair*>(peers[p] + pairs_off) + (((size_t)par * world + rank) * max_b + b) * k_max + j;
        const XPair v = s_loc[j];
        *reinterpret_cast<volatile float*>(&dst->score) = v.score;
        *reinterpret_cast<volatile long long*>(&dst->index) = v.index;
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        // release: the pairs above are visible system-wide before the flag is
        __threadfence_system();
        volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(peers[threadIdx.x]) + ((size_t)par * world + rank) * max_b + b;
        *flag = epoch;
        // acquire: wait for peer `threadIdx.x`'s pairs of this detection in this rank's own buffer
        volatile unsigned* mine = reinterpret_cast<volatile unsigned*>(peers[rank]) + ((size_t)par * world + threadIdx.x) * max_b + b;
        const unsigned long long t0 = globaltimer_ns();
        while (!epoch_reached(*mine, epoch)) {
            __nanosleep(200);
            if (timeout_ns && globaltimer_ns() - t0 > timeout_ns) {
                record_xchg_fault(fault, (int)threadIdx.x, epoch);
                s_gave_up = 1;
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (s_gave_up) {  // a peer never showed up: poison this detection's result instead of merging stale slots
        if ((int)threadIdx.x < k) {
            out_score[(size_t)b * k + threadIdx.x] = __int_as_float(0x7fc00000);
            out_idx[(size_t)b * k + threadIdx.x] = -1;
        }
        return;
    }
    // ---- merge: world * k candidates, ties go to the lowest rank / slot (= lowest view index for contiguous shards) ----
    const int n = world * k;
    const XPair* own = reinterpret_cast<const XPair*>(peers[rank] + pairs_off);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int r = i / k, j = i - r * k;
        const float sc = __ldcg(&own[(((size_t)par * world + r) * max_b + b) * k_max + j].score);
        s_merge[i] = pack_key(sc + 0.0f, (uint32_t)i);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int o = 0; o < k; ++o) {
            unsigned long long best = 0ull;
            for (int i = lane; i < n; i += 32) best = s_merge[i] > best ? s_merge[i] : best;
            for (int sft = 16; sft > 0; sft >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, sft);
                best = other > best ? other : best;
            }
            if (lane == 0) {
                const uint32_t pos = key_index(best);
                const int r = pos / k, j = pos - r * k;
                out_score[(size_t)b * k + o] = key_value(best);
                out_idx[(size_t)b * k + o] = __ldcg(&own[(((size_t)par * world + r) * max_b + b) * k_max + j].index);
                s_merge[pos] = 0ull;
            }
            __syncwarp();
        }
    }
}

// ---- flags for bulk pushes (PeerGather): one u32 per (parity, source rank) at the head of a buffer ----
__global__ void xchg_signal_kernel(char* const* __restrict__ peers, size_t flag_off, int rank, int world, unsigned epoch) {
    // runs after the pushes of this rank in stream order: the copies have landed before the flag is written
    if ((int)threadIdx.x < world) {
        __threadfence_system();
        volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(peers[threadIdx.x] + flag_off) + rank;
        *flag = epoch;
    }
}
__global__ void xchg_wait_kernel(const char* __restrict__ own, size_t flag_off, int world, unsigned epoch,
                                 unsigned long long timeout_ns, int* __restrict__ fault) {
    if ((int)threadIdx.x < world) {
        const volatile unsigned* flag = reinterpret_cast<const volatile unsigned*>(own + flag_off) + threadIdx.x;
        const unsigned long long t0 = globaltimer_ns();
        while (!epoch_reached(*flag, epoch)) {
            __nanosleep(500);
            if (timeout_ns && globaltimer_ns() - t0 > timeout_ns) {
                record_xchg_fault(fault, (int)threadIdx.x, epoch);  // reported by pp_check_device_faults; no trap
                break;
            }
        }
        __threadfence_system();
    }
}

}  // namespace pp

extern "C" int pp_xchg_push(const void* src, size_t bytes, const void* const* peers_host, int world, size_t dst_offset,
                            void* stream) {
    using namespace pp;
    PP_CHECK_ARG(src && peers_host && world > 0, "pp_xchg_push: bad arguments");
    for (int p = 0; p < world; ++p)
        PP_CUDA(cudaMemcpyAsync(static_cast<char*>(const_cast<void*>(peers_host[p])) + dst_offset, src, bytes,
                                cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return PP_OK;
}

// Optional fan-out of the bulk copies (PICOPOSE_B200_PUSH_STREAMS = n > 1; default 1 = all on the caller's stream):
// copies issued on one stream run one after the other, and a gather at 8 GPUs is 8 x 4 MB of them behind a host upload,
// all of which has to fit into one 0.4 ms step.  With n > 1 they are dealt over n internal non-blocking streams per
// device (fork / join with events on the caller's stream).  Measured (tools/microbench/push_chain.py,
// profiles/r2i_push_chain.md): on an idle GPU the 8-copy chain takes 86 us either way (one stream already fills the
// link); with the copy engines busy 4 streams take 328 us against 398 us; inside bench.py --gpus 2 the fork / join costs
// 4-8 % of the end-to-end rate.  Hence off by default.
namespace pp {
constexpr int kMaxPushStreams = 8;
struct PushLanes {
    int n = 0;     // 0 = not created yet, < 0 = creation failed (stay on the caller's stream)
    cudaStream_t st[kMaxPushStreams];
    cudaEvent_t fork, join[kMaxPushStreams];
};
static int push_lanes_wanted() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("PICOPOSE_B200_PUSH_STREAMS");
        int n = e ? atoi(e) : 1;
        cached = n < 1 ? 1 : (n > kMaxPushStreams ? kMaxPushStreams : n);
    }
    return cached;
}
static PushLanes* push_lanes() {
    static PushLanes lanes[64];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    PushLanes& l = lanes[dev];
    if (l.n == 0) {
        const int want = push_lanes_wanted();
        bool ok = cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; ok && i < want; ++i)
            ok = cudaStreamCreateWithFlags(&l.st[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&l.join[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) (void)cudaGetLastError();
        l.n = ok ? want : -1;
    }
    return l.n > 0 ? &l : nullptr;
}
}  // namespace pp

extern "C" int pp_xchg_push_signal(const void* src, size_t bytes, size_t dst_offset, const void* payload,
                                   size_t payload_bytes, size_t payload_offset, const void* const* peers_host,
                                   const void* const* peers_dev, size_t flag_offset, int rank, int world, uint32_t epoch,
                                   void* stream) {
    using namespace pp;
    PP_CHECK_ARG(src && peers_host && peers_dev && world > 0 && world <= 256 && rank >= 0 && rank < world && epoch != 0,
                 "pp_xchg_push_signal: bad arguments");
    PP_CHECK_ARG(payload_bytes == 0 || payload, "pp_xchg_push_signal: null payload");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PushLanes* lanes = push_lanes_wanted() > 1 && world > 1 ? push_lanes() : nullptr;
    if (lanes) {
        static std::mutex issue_mu;                         // the events are per device, not per caller
        std::lock_guard<std::mutex> lock(issue_mu);
        const int n = lanes->n < world ? lanes->n : world;
        PP_CUDA(cudaEventRecord(lanes->fork, st));
        for (int i = 0; i < n; ++i) PP_CUDA(cudaStreamWaitEvent(lanes->st[i], lanes->fork, 0));
        for (int p = 0; p < world; ++p) {
            const int q = (rank + 1 + p) % world;          // start with the neighbour: ranks do not all hit peer 0 first
            PP_CUDA(cudaMemcpyAsync(static_cast<char*>(const_cast<void*>(peers_host[q])) + dst_offset, src, bytes,
                                    cudaMemcpyDeviceToDevice, lanes->st[p % n]));
        }
        for (int i = 0; i < n; ++i) {
            PP_CUDA(cudaEventRecord(lanes->join[i], lanes->st[i]));
            PP_CUDA(cudaStreamWaitEvent(st, lanes->join[i], 0));
        }
    } else {
        if (int rc = pp_xchg_push(src, bytes, peers_host, world, dst_offset, stream)) return rc;
    }
    // The payload (the query masks, a few KB) takes the copy engines as well.  Storing it from the flag kernel instead
    // (world copies fewer; 115 -> 86 us for the 8-rank chain on an idle GPU) was built and measured inside the real loop:
    // bench.py --gpus 2 end to end 5363 / 5145 det/s with copies against 3998 / 4036 with kernel stores, same box,
    // alternating (profiles/r2i_push_chain.md).  The cause was not isolated; the suspect is the system-scope fence behind
    // peer stores on an SM whose memory pipes the contraction keeps full.
    if (payload_bytes)
        if (int rc = pp_xchg_push(payload, payload_bytes, peers_host, world, payload_offset, stream)) return rc;
    return pp_xchg_signal(peers_dev, flag_offset, rank, world, epoch, stream);
}

extern "C" int pp_xchg_signal(const void* const* peers_dev, size_t flag_offset, int rank, int world, uint32_t epoch,
                              void* stream) {
    using namespace pp;
    PP_CHECK_ARG(peers_dev && world > 0 && world <= 256 && rank >= 0 && rank < world && epoch != 0, "pp_xchg_signal: bad arguments");
    xchg_signal_kernel<<<1, 32 * ((world + 31) / 32), 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<char* const*>(const_cast<void* const*>(peers_dev)), flag_offset, rank, world, epoch);
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" int pp_xchg_wait(const void* own_buf, size_t flag_offset, int world, uint32_t epoch, void* stream) {
    using namespace pp;
    PP_CHECK_ARG(own_buf && world > 0 && world <= 256 && epoch != 0, "pp_xchg_wait: bad arguments");
    int* fault = nullptr;
    if (int rc = fault_buffer(&fault)) return rc;
    xchg_wait_kernel<<<1, 32 * ((world + 31) / 32), 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const char*>(own_buf), flag_offset, world, epoch,
                                                                       xchg_timeout_ns(), fault);
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" size_t pp_xchg_bytes(int world, int max_b, int k_max) {
    if (world <= 0 || max_b <= 0 || k_max <= 0) return 0;
    return pp::xchg_layout(world, max_b, k_max).total;
}

extern "C" int pp_xchg_create(size_t bytes, void** buf, void* handle64) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    PP_CHECK_ARG(bytes > 0 && buf && handle64, "pp_xchg_create: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    PP_CUDA(cudaMalloc(&p, bytes));
    PP_CUDA(cudaMemset(p, 0, bytes));
    PP_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(PP_ERR_DEVICE, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle64, &h, 64);
    *buf = p;
    return PP_OK;
}

extern "C" int pp_xchg_open(const void* handle64, void** peer_buf) {
    using namespace pp;
    PP_CHECK_ARG(handle64 && peer_buf, "pp_xchg_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(PP_ERR_DEVICE, "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    *peer_buf = p;
    return PP_OK;
}

extern "C" int pp_xchg_close(void* peer_buf) {
    using namespace pp;
    if (peer_buf) PP_CUDA(cudaIpcCloseMemHandle(peer_buf));
    return PP_OK;
}

extern "C" int pp_xchg_destroy(void* buf) {
    using namespace pp;
    if (buf) PP_CUDA(cudaFree(buf));
    return PP_OK;
}

extern "C" int pp_topk_exchange(const float* scores, int B, int N, int k, int64_t idx_offset, const void* const* peers_dev,
                                int rank, int world, int max_b, int k_max, uint32_t epoch, float* out_score,
                                int64_t* out_idx, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (B == 0 || k == 0) return PP_OK;
    PP_CHECK_ARG(peers_dev && out_score && out_idx && (scores || N == 0), "pp_topk_exchange: null pointer");
    PP_CHECK_ARG(world >= 1 && rank >= 0 && rank < world && world <= 256, "pp_topk_exchange: bad rank %d / world %d", rank, world);
    PP_CHECK_ARG(B <= max_b && k <= k_max && k > 0 && k <= 256 && world * k <= 1024 && N >= 0 && N <= 24000,
                 "pp_topk_exchange: bad sizes (B=%d of %d, k=%d of %d, N=%d)", B, max_b, k, k_max, N);
    PP_CHECK_ARG(epoch != 0, "pp_topk_exchange: epoch 0 is reserved for 'never written'");
    const XchgLayout l = xchg_layout(world, max_b, k_max);
    const size_t smem = ((size_t)N + 8) * sizeof(unsigned long long);
    if (smem > 48 * 1024)
        PP_CUDA(cudaFuncSetAttribute(topk_exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int* fault = nullptr;
    if (int rc = fault_buffer(&fault)) return rc;
    // A block spins until the peers' blocks of the SAME detection have run, so every block of a launch must be
    // co-resident with the ones it (transitively) waits for: launches are cut into chunks of at most the number of
    // blocks the device holds at once.  Chunks of one call share the epoch (flags are per detection) and every rank
    // issues them in the same order, so a chunk only ever waits for the same chunk of its peers.
    int per_sm = 0;
    PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, topk_exchange_kernel, 256, smem));
    PP_CHECK_ARG(per_sm >= 1, "pp_topk_exchange: kernel does not fit an SM with N=%d", N);
    const int resident = per_sm * sm_count();
    for (int b0 = 0; b0 < B; b0 += resident) {
        const int nb = B - b0 < resident ? B - b0 : resident;
        topk_exchange_kernel<<<nb, 256, smem, static_cast<cudaStream_t>(stream)>>>(
            scores, N, k, (long long)idx_offset, reinterpret_cast<char* const*>(const_cast<void* const*>(peers_dev)), rank,
            world, max_b, k_max, l.pairs_off, epoch, b0, xchg_timeout_ns(), fault, out_score,
            reinterpret_cast<long long*>(out_idx));
        PP_LAUNCHED();
    }
    return PP_OK;
}
