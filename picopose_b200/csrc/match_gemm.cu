// Stage-1 similarity contraction on the 5th-gen tensor cores with a fused reduction epilogue.
//
// Replaces the reference's  sim = einsum("b c t, b n c s -> b n t s")  and everything that sweeps
// the materialised sim tensor afterwards (utils/matching.py:47-51, and :22 for the stage-2 volume):
// the T x S similarity tile lives only in TMEM; the epilogue reduces it on the fly to
//   * per query row t   : max_s sim[t,s] and its first argmax            (utils/matching.py:50)
//   * per template col s: max_t m[t]*sim[t,s] and its first argmax       (utils/matching.py:48,51)
// published as packed 64-bit keys with atomicMax, so tiles of one (b, n) may run on any SM in any order.
//
// Structure (persistent, warp-specialised, one CTA or one CTA pair per 148 SMs):
//   warp 0     TMA producer: K-major bf16 tiles, 128-byte swizzle, STAGES-deep mbarrier ring
//   warp 1     MMA issuer  : tcgen05.mma kind::f16, M = 128*CL, N = 256, K = 16 per instruction,
//                            accumulators double-buffered in TMEM (2 x 256 columns)
//   warp 2     TMEM allocator
//   warps 4-19 epilogue    : tcgen05.ld 32x32b -> registers; warp w reads lane quarter w%4 and the
//                            32-column chunks (w-4)/4, (w-4)/4 + 4 of the accumulator
// CL = 2 pairs two SMs (cta_group::2): each CTA loads its own 128 A rows and half of the B tile, the
// leader issues M=256 MMAs, commits are multicast to both CTAs.
//
// Operand roles.  EPI_MATCH puts the TEMPLATE patches on the M side (A operand, TMEM lanes) and the compact
// query rows on the N side (B operand, TMEM columns): UMMA M is fixed at 128 per CTA but N may be any multiple
// of 16, so a detection with tv unmasked patches is cut into ceil(tv/256) column tiles of
// round_up(tv / tiles, 32) columns and (almost) no padding rows are multiplied -- 655 live patches cost
// 3 x 224 columns instead of 3 x 256 rows.  EPI_EMIT keeps the query on the M side (its rows are not compacted).
#include "pp_common.cuh"
#include "pp_ptx.cuh"

#include <cudaTypedefs.h>
#include <mutex>

namespace pp {

constexpr int BLOCK_M = 128;  // accumulator rows per CTA (= TMEM lanes)
constexpr int BLOCK_N = 256;  // accumulator columns per tile (UMMA N)
constexpr int BLOCK_K = 64;   // bf16 elements per k-block: one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_ACC = 2;
constexpr int NUM_EPI_WARPS = 16;                                   // 4 per TMEM lane quarter, 64 columns each
constexpr int EPI_COLS = BLOCK_N / (NUM_EPI_WARPS / 4);               // accumulator columns per epilogue warp
constexpr int EPI_CHUNKS = EPI_COLS / 32;
constexpr int FIRST_EPI_WARP = 4;
constexpr int GEMM_THREADS = (FIRST_EPI_WARP + NUM_EPI_WARPS) * 32;  // 640
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;                 // 16 KiB
constexpr long long WAIT_TIMEOUT_CYCLES = 4000000000LL;              // ~2 s: a stuck pipeline traps instead of hanging

constexpr int MAX_DETS_PER_LAUNCH = 1024;  // tile prefix table lives in shared memory

// EPI_MATCH_FAST: same reductions with cheaper keys (bf16 mode): the similarity is offset by 2.0 so that every key is a
// positive float -- no sign handling, no rounding add -- at a resolution of 2^-17 (7.6e-6 absolute), far below the bf16
// operands' own error; EPI_MATCH keeps 2^-19 relative for the fp32-accurate modes.
// EPI_SIM: the stage-2 similarity volume (utils/matching.py:22-25) straight from the accumulator: norms, template mask,
// clamp and the "(w h)" transposed layout applied in the epilogue, out[b, s, h, w] = max(0, sim[b, t = w*H + h, s] * m[s]).
enum { EPI_MATCH = 0, EPI_EMIT = 1, EPI_MATCH_FAST = 2, EPI_SIM = 3 };

struct GemmParams {
    int B, N, T;  // detections, views per bank, patches (T == S)
    int num_k_blocks;
    int num_mt, num_nt;  // tiles along T (per 128*CL rows) and S (per 256 columns)
    uint32_t num_groups;  // ceil(B / chunk) * N * num_mt work groups
    int tile_mode;        // 1: short launch, single TILES dealt round-robin (per-detection tile prefix in shared memory)
    int chunk;            // detections per group (>= 1): consecutive entries of det_order
    const int32_t* det_order;  // (B,) detections sorted by bank (null = identity): a chunk then shares its M-side tiles
    int n_banks;
    const int32_t* bank_of_det;    // (B,) or null = identity; entries outside [0, n_banks) are clamped and reported (fault 6)
    const float* mrow;             // (B, T) nearest-resized query mask, indexed by patch
    const int* tv;                 // (B) unmasked query patches per detection (EPI_MATCH: rows are compacted); null = T
    const int* rowmap;             // (B, T) patch index of compact query row r
    const float* ra;               // (B, T) EPI_EMIT only: inverse norms of the query rows (EPI_MATCH: folded into the operand)
    const float* rb;               // (n_banks, N, T) inverse norms of the template patches
    unsigned long long* rowkey;    // (B, N, T)
    unsigned long long* colkey;    // (B, N, T)
    float* emit;                   // EPI_EMIT: (B*N, T, T) raw products times emit_scale
    float emit_scale;
    int emit_tile_w;               // EPI_EMIT: 0 = rows of T keys (reference layout); W > 0 = every row is an (T/W x W) key map
                                   // stored as 4 x 8 tiles of 32 floats (one 128-byte line each), see corr_lookup.cu
    const float* cmask;            // EPI_SIM: (B, Hm, Wm) template masks, nearest-resized to the patch grid on the fly
    int Hm, Wm, gh, gw;            // EPI_SIM: mask size, patch grid (gh x gw, T = gh * gw)
    int* fault;                    // host-mapped fault record
};

template <int CL>
struct GemmCfg {
    static constexpr int STAGES = CL == 1 ? 4 : 6;
    static constexpr int B_ROWS = BLOCK_N / CL;
    static constexpr int B_STAGE_BYTES = B_ROWS * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int BAR_BYTES = 256;
    static constexpr int RB_BYTES = NUM_EPI_WARPS * EPI_COLS * 8;  // per epilogue warp: factor + patch index of its columns
    static constexpr int PREFIX_BYTES = (3 * MAX_DETS_PER_LAUNCH + 1) * 4;  // per-detection tile prefix (tile mode), live rows, order
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + RB_BYTES + PREFIX_BYTES + 1024;  // + alignment slack
};

__device__ __noinline__ void report_fault(int* fault, int code, int a, int b) {
    if (fault) {
        fault[1] = blockIdx.x;
        fault[2] = threadIdx.x;
        fault[3] = a;
        fault[4] = b;
        __threadfence_system();
        fault[0] = code;
        __threadfence_system();
    }
    __trap();
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* fault, int code, int aux) {
    if (ptx::mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!ptx::mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > WAIT_TIMEOUT_CYCLES) report_fault(fault, code, aux, (int)parity);
    }
}

struct TileCoord {
    int b, n, nt, mt;
    int nct;    // column tiles of this (b, n, mt) group
    int ncols;  // accumulator columns of a tile (UMMA N), a multiple of 32
    int rows;   // EPI_MATCH: live compact query rows of detection b
};

// Work is handed out in GROUPS (detection chunk, n, mt): one 128*CL-row tile of the M-side operand (a template-patch tile,
// 256 x Kp bf16) against ALL column tiles of up to `chunk` detections.  A cluster walks them back to back, so the M-side
// tile is read from HBM once and comes out of L2 for the rest of the group; with det_order sorting the detections by bank,
// detections that share an object bank share those reads (8 detections per bank in BASELINE configs[2]).
// (r1 dealt single tiles round-robin; the column tiles of one M-side tile then ran on neighbouring clusters, which drift
// apart over a long batch until the shared tile has left L2: ncu showed 25 GB of DRAM reads for 10.8 GB of operands on an
// 8 x 642 batch, and every detection streamed its whole bank again.)
// (launches of up to MAX_DETS_PER_LAUNCH detections keep copies of tv[] and det_order[] in shared memory: the per-tile /
// per-detection decode of all three roles then stays off the global-memory path, whose L1 lines the epilogue's streaming
// loads keep evicting -- measured +3 % on the configs[1] contraction, same box)
__device__ __forceinline__ int live_rows(const GemmParams& p, int b, const int* rows_s = nullptr) {
    return rows_s ? rows_s[b] : (p.tv ? __ldg(p.tv + b) : p.T);
}
__device__ __forceinline__ uint32_t col_tiles(int rows) { return (uint32_t)((rows + BLOCK_N - 1) / BLOCK_N); }
// bank of detection b, forced into range: a bad index must not become an out-of-bounds TMA coordinate / rnorm read
__device__ __forceinline__ int bank_of(const GemmParams& p, int b) {
    if (!p.bank_of_det) return b;
    const int v = __ldg(p.bank_of_det + b);
    return v < 0 ? 0 : (v >= p.n_banks ? p.n_banks - 1 : v);
}
// Work item of a cluster: detections [pos0, pos1) x column tiles [nt0, nt0 + nts) of (view n, patch tile mt).
//  * group mode (long launches): all column tiles of up to `chunk` detections (nts < 0 = "all of the detection's");
//  * tile mode (short launches, e.g. the 1 x 162 and 8 x 21 shapes of BASELINE configs[1]): ONE tile -- a launch of ~30
//    tile-times per cluster cannot afford groups of 3-6 tiles (28 vs 30 tile-times on the 8 x 21 shape, measured +6.6 %),
//    and it is over before neighbouring clusters can drift apart, so they still share the template tile in L2.
struct WorkItem {
    int pos0, pos1, n, mt, nt0, nts;
};
__device__ __forceinline__ WorkItem decode_work(uint32_t g, const GemmParams& p, const uint32_t* prefix, const int* rows_s) {
    WorkItem w;
    if (p.tile_mode) {
        // flat tile index -> detection by binary search in the per-detection tile prefix, then (n, mt, nt)
        int lo = 0, hi = p.B;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (prefix[mid] <= g) lo = mid; else hi = mid;
        }
        const uint32_t nct = col_tiles(live_rows(p, lo, rows_s));
        const uint32_t local = g - prefix[lo];
        const uint32_t r = local / nct;
        w.nt0 = (int)(local - r * nct);
        w.nts = 1;
        const uint32_t n = r / (uint32_t)p.num_mt;
        w.mt = (int)(r - n * (uint32_t)p.num_mt);
        w.n = (int)n;
        w.pos0 = lo;
        w.pos1 = lo + 1;
        return w;
    }
    const uint32_t r = g / (uint32_t)p.num_mt;
    w.mt = (int)(g - r * (uint32_t)p.num_mt);
    const uint32_t c = r / (uint32_t)p.N;
    w.n = (int)(r - c * (uint32_t)p.N);
    w.pos0 = (int)c * p.chunk;
    w.pos1 = min(w.pos0 + p.chunk, p.B);
    w.nt0 = 0;
    w.nts = -1;
    return w;
}
// detection at position `pos` of the (bank-sorted) order: its tile shape
template <bool MATCH>
__device__ __forceinline__ TileCoord det_tiles(int pos, int n, int mt, const GemmParams& p, const int* rows_s, const int* order_s) {
    TileCoord c;
    c.b = order_s ? order_s[pos] : (p.det_order ? __ldg(p.det_order + pos) : pos);
    c.n = n;
    c.mt = mt;
    c.nt = 0;
    if (MATCH) {
        // a detection with tv unmasked patches is cut into ceil(tv / 256) column tiles of round_up(tv / tiles, 32) columns
        c.rows = live_rows(p, c.b, rows_s);
        c.nct = (int)col_tiles(c.rows);
        c.ncols = c.nct ? (int)((((uint32_t)c.rows + (uint32_t)c.nct - 1) / (uint32_t)c.nct + 31u) & ~31u) : 0;
    } else {
        c.rows = p.T;
        c.nct = p.num_nt;
        c.ncols = BLOCK_N;
    }
    return c;
}

template <int CL, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
match_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const GemmParams p) {
    using Cfg = GemmCfg<CL>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    // 128-byte swizzle atoms need 1024-byte alignment
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_base + STAGES * A_STAGE_BYTES;
    const uint32_t bars = smem_base + STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + NUM_ACC + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + STAGES * Cfg::STAGE_BYTES + 8 * (2 * STAGES + 2 * NUM_ACC));
    float* rb_stage = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES);
    uint32_t* tile_prefix = reinterpret_cast<uint32_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES + Cfg::RB_BYTES);
    int* rows_tab = reinterpret_cast<int*>(tile_prefix + MAX_DETS_PER_LAUNCH + 1);
    constexpr bool MATCH = EPI == EPI_MATCH || EPI == EPI_MATCH_FAST;  // template patches on the M side, compact query rows on the N side
    constexpr bool FAST = EPI == EPI_MATCH_FAST;
    int* order_tab = rows_tab + MAX_DETS_PER_LAUNCH;
    const bool tabs = MATCH && p.B <= MAX_DETS_PER_LAUNCH;
    const int* rows_s = tabs ? rows_tab : nullptr;
    const int* order_s = tabs ? order_tab : nullptr;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = CL > 1 ? ptx::cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const uint32_t cluster_id = blockIdx.x / CL;
    const uint32_t num_clusters = gridDim.x / CL;

    if (CL > 1) ptx::cluster_sync();  // both CTAs resident before the paired TMEM allocation

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_a);
        ptx::prefetch_tensormap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(full_bar(s), CL);  // one producer arrival per CTA (on the leader's barrier)
            ptx::mbar_init(empty_bar(s), 1);  // one tcgen05.commit arrival
        }
        for (int s = 0; s < NUM_ACC; ++s) {
            ptx::mbar_init(tfull_bar(s), 1);
            ptx::mbar_init(tempty_bar(s), CL * NUM_EPI_WARPS);  // one arrival per epilogue warp of every CTA
        }
        ptx::fence_barrier_init();
    } else if (warp == 2) {
        ptx::tmem_alloc<CL>(ptx::smem_u32((const void*)tmem_slot), 512);
    } else if (warp == 3 && tabs) {
        // live rows and launch order of every detection; tile mode: inclusive scan of the per-detection tile counts,
        // 32 detections per step
        grid_dependency_wait();  // tv / det_order come from the kernels before this one
        uint32_t run = 0;
        if (lane == 0) tile_prefix[0] = 0;
        for (int b0 = 0; b0 < p.B; b0 += 32) {
            const int b = b0 + lane;
            uint32_t cnt = 0;
            if (b < p.B) {
                const int rows = live_rows(p, b);
                rows_tab[b] = rows;
                order_tab[b] = p.det_order ? __ldg(p.det_order + b) : b;
                cnt = col_tiles(rows) * (uint32_t)(p.N * p.num_mt);
            }
            uint32_t inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += up;
            }
            if (b < p.B) tile_prefix[b + 1] = run + inc;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    ptx::tc_fence_before();
    if (CL > 1) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t num_groups = (MATCH && p.tile_mode) ? tile_prefix[p.B] : p.num_groups;
    // programmatic dependent launch: barrier set-up and the TMEM allocation above overlap the tail of the previous kernel;
    // operands, norms and masks are only touched from here on
    grid_dependency_wait();

    if (warp == 0) {
        // ===================== TMA producer (every CTA) =====================
        // The whole warp walks the loop (converged code keeps addresses and descriptors in uniform registers);
        // one elected lane issues the copies and the barrier arrivals.
        int stage = 0;
        uint32_t phase = 0;
        for (uint32_t grp = cluster_id; grp < num_groups; grp += num_clusters) {
         const WorkItem wi = decode_work(grp, p, tile_prefix, rows_s);
         for (int pos = wi.pos0; pos < wi.pos1; ++pos) {
          TileCoord tc = det_tiles<MATCH>(pos, wi.n, wi.mt, p, rows_s, order_s);
          const int bank = bank_of(p, tc.b);
          const int q_row0 = tc.b * p.T;                                       // query operand: rows of detection b
          const int t_row0 = (int)(((long long)bank * p.N + tc.n) * p.T);      // bank operand: rows of view n
          const int nt_end = wi.nts < 0 ? tc.nct : wi.nt0 + wi.nts;
          for (tc.nt = wi.nt0; tc.nt < nt_end; ++tc.nt) {
            // A = M-side operand (128 rows per CTA), B = N-side operand (each CTA of a pair holds half of the columns;
            // the box is always B_ROWS rows, the MMA reads the first ncols / CL of them)
            const int a_row = (MATCH ? t_row0 : q_row0) + tc.mt * (BLOCK_M * CL) + (int)cta_rank * BLOCK_M;
            const int b_row = MATCH ? q_row0 + tc.nt * tc.ncols + (int)cta_rank * (tc.ncols / CL)
                                    : t_row0 + tc.nt * BLOCK_N + (int)cta_rank * Cfg::B_ROWS;
            for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u, p.fault, 1, stage);
                const uint32_t da = smem_a + stage * A_STAGE_BYTES;
                const uint32_t db = smem_b + stage * Cfg::B_STAGE_BYTES;
                if (ptx::elect_one_sync()) {
                    if (CL == 1) {
                        ptx::mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                        ptx::tma_load_2d(&tmap_a, full_bar(stage), da, kb * BLOCK_K, a_row);
                        ptx::tma_load_2d(&tmap_b, full_bar(stage), db, kb * BLOCK_K, b_row);
                    } else {
                        ptx::tma_load_2d_pair(&tmap_a, full_bar(stage), da, kb * BLOCK_K, a_row);
                        ptx::tma_load_2d_pair(&tmap_b, full_bar(stage), db, kb * BLOCK_K, b_row);
                        if (leader) ptx::mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES * CL);
                        else ptx::mbar_arrive_cluster(full_bar(stage), 0);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
          }
         }
        }
    } else if (warp == 1 && leader) {
        // ===================== MMA issuer (leader CTA; converged warp, one elected lane issues) =====================
        int stage = 0;
        uint32_t phase = 0;
        uint32_t iter = 0;
        for (uint32_t grp = cluster_id; grp < num_groups; grp += num_clusters) {
         const WorkItem wi = decode_work(grp, p, tile_prefix, rows_s);
         for (int pos = wi.pos0; pos < wi.pos1; ++pos) {
          const TileCoord gc = det_tiles<MATCH>(pos, wi.n, wi.mt, p, rows_s, order_s);
          const uint32_t idesc = ptx::idesc_bf16(BLOCK_M * CL, gc.ncols);
          const int nt_end = wi.nts < 0 ? gc.nct : wi.nt0 + wi.nts;
          for (int nt = wi.nt0; nt < nt_end; ++nt, ++iter) {
            const int as = (int)(iter & 1);
            const uint32_t aphase = (uint32_t)((iter >> 1) & 1);
            mbar_wait(tempty_bar(as), aphase ^ 1u, p.fault, 2, as);
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
            for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                mbar_wait(full_bar(stage), phase, p.fault, 3, stage);
                ptx::tc_fence_after();
                const uint64_t adesc = ptx::smem_desc_sw128(smem_a + stage * A_STAGE_BYTES);
                const uint64_t bdesc = ptx::smem_desc_sw128(smem_b + stage * Cfg::B_STAGE_BYTES);
                if (ptx::elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 16 elements = 32 bytes inside the swizzle row: +2 in the (addr >> 4) field
                        ptx::umma_bf16<CL>(adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), tmem_d,
                                           (kb > 0 || k > 0) ? 1u : 0u, idesc);
                    }
                    ptx::umma_commit<CL>(empty_bar(stage));  // smem slot reusable once these MMAs retire
                    if (kb == p.num_k_blocks - 1) ptx::umma_commit<CL>(tfull_bar(as));
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
          }
         }
        }
    } else if (warp >= FIRST_EPI_WARP) {
        // ===================== epilogue (every CTA) =====================
        const int e = warp - FIRST_EPI_WARP;
        const int q = warp & 3;   // TMEM lane quarter this warp may read
        const int hh = e >> 2;    // first 32-column chunk of the tile this warp reads (then hh + 4)
        uint32_t iter = 0;
        for (uint32_t grp = cluster_id; grp < num_groups; grp += num_clusters) {
         const WorkItem wi = decode_work(grp, p, tile_prefix, rows_s);
         for (int pos = wi.pos0; pos < wi.pos1; ++pos) {
          TileCoord tc = det_tiles<MATCH>(pos, wi.n, wi.mt, p, rows_s, order_s);
          const int nt_end = wi.nts < 0 ? tc.nct : wi.nt0 + wi.nts;
          // column maxima (over the query rows, lane-local) run across all column tiles of the group: one atomic per
          // (template patch, group) instead of one per tile
          float best = -INFINITY;
          int best_t = 0;
          for (tc.nt = wi.nt0; tc.nt < nt_end; ++tc.nt, ++iter) {
            const int as = (int)(iter & 1);
            const uint32_t aphase = (uint32_t)((iter >> 1) & 1);
            const int T = p.T;
            const size_t bn = (size_t)tc.b * p.N + tc.n;
            const int warp_row0 = tc.mt * (BLOCK_M * CL) + (int)cta_rank * BLOCK_M + q * 32;  // M-side row of lane 0
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N);
            auto release_tmem = [&]() {
                // this warp has drained its part of the accumulator: hand the TMEM stage back
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CL == 1) ptx::mbar_arrive(tempty_bar(as));
                    else ptx::mbar_arrive_cluster(tempty_bar(as), 0);
                }
            };

            if (MATCH) {
                // lane = template patch s (TMEM lane), columns = compact query rows r of detection b; t = rowmap[r].
                // The query operand is already normalised and multiplied by its mask value (pp_match_prepare_query), so
                // acc * rb[s] IS the masked similarity m[t] * sim[t, s]; no per-column factor is left.
                const int s_row = warp_row0 + lane;
                const bool s_ok = s_row < T;
                const int bank = bank_of(p, tc.b);
                const float rb_l = s_ok ? __ldg(p.rb + ((size_t)bank * p.N + tc.n) * T + s_row) : 0.f;
                // lanes past the last template patch must lose every comparison
                const float c_add = FAST ? (s_ok ? 2.0f : 0.0f) : (s_ok ? 0.0f : -3.0e38f);
                const int col0 = tc.nt * tc.ncols;  // first compact query row of the tile
                // patch index of this warp's <= 2 chunks of columns, staged before waiting for the accumulator so the
                // global latency hides behind the MMAs
                int* tp_s = reinterpret_cast<int*>(rb_stage + e * (2 * EPI_COLS));
                __syncwarp();  // previous tile's readers are done
#pragma unroll
                for (int i = 0; i < EPI_CHUNKS; ++i) {
                    const int r = col0 + (hh + 4 * i) * 32 + lane;
                    const bool r_ok = (hh + 4 * i) * 32 < tc.ncols && r < tc.rows;
                    tp_s[i * 32 + lane] = r_ok ? (p.rowmap ? __ldg(p.rowmap + (size_t)tc.b * T + r) : r) : 0;
                }
                __syncwarp();

                mbar_wait(tfull_bar(as), aphase, p.fault, 4, as);
                ptx::tc_fence_after();
                if (hh * 32 >= tc.ncols) release_tmem();  // narrow tile: nothing for this warp to read

#pragma unroll 1
                for (int i = 0; i < EPI_CHUNKS; ++i) {
                    const int c = hh + 4 * i;
                    if (c * 32 >= tc.ncols) break;  // warp-uniform
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
                    ptx::tmem_ld_wait();
                    if ((c + 4) * 32 >= tc.ncols) release_tmem();
                    const int r0 = col0 + c * 32;
                    const int ncols = min(32, tc.rows - r0);
                    if (ncols <= 0) continue;  // warp-uniform: chunk entirely past the last live query row
                    // ---- lane-local: first-argmax over t (this lane's columns, in increasing t) of the masked similarity.
                    // ---- across lanes: first-argmax over s (the lanes) for every column.  Each value becomes a float key
                    // whose 5 low mantissa bits carry the lane, so that a float max means "largest value, then lowest
                    // lane"; a 31-shuffle butterfly transposes and reduces the 32 x 32 keys of the warp.
                    float k[32];
                    float cbest;
                    int cj;
                    if (FAST) {
                        // keys are bits(sim + 2.0) with the low 5 bits replaced: all positive, ordered like the values;
                        // payload 31 - lane (rows) resp. 31 - j (columns) lets the lowest index win a tie
                        const uint32_t lane_pos = (uint32_t)(31 - lane);
                        float ck = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const uint32_t xb = __float_as_uint(fmaf(__uint_as_float(v[j]), rb_l, c_add)) & 0xFFFFFFE0u;
                            k[j] = __uint_as_float(xb | lane_pos);
                            ck = fmaxf(ck, __uint_as_float(xb | (uint32_t)(31 - j)));
                        }
                        if (ncols < 32) {  // ragged last chunk: columns past the last live row hold zeros (= 2.0): leave them out
                            ck = 0.f;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < ncols) ck = fmaxf(ck, __uint_as_float((__float_as_uint(k[j]) & 0xFFFFFFE0u) | (uint32_t)(31 - j)));
                        }
                        cj = 31 - (int)(__float_as_uint(ck) & 31u);
                        cbest = __uint_as_float(__float_as_uint(ck) & 0xFFFFFFE0u) - 2.0f;
                    } else {
                        // exact compare on the lane's own columns (rb[s] > 0 commutes with the max; strict > keeps the
                        // first index); row keys: value rounded to nearest at bit 5 (2^-19 relative), payload 31 - lane
                        // for x >= 0 and lane for x < 0 so that the lowest lane wins for either sign
                        cbest = -INFINITY;
                        cj = 0;
                        const uint32_t lane_pos = (uint32_t)(31 - lane), lane_neg = (uint32_t)lane;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float x = __uint_as_float(v[j]);
                            if (j < ncols && x > cbest) { cbest = x; cj = j; }
                            const uint32_t xb = __float_as_uint(fmaf(x, rb_l, c_add));
                            k[j] = __uint_as_float(((xb + 0x10u) & 0xFFFFFFE0u) | ((xb >> 31) ? lane_neg : lane_pos));
                        }
                        cbest *= rb_l;  // the masked similarity itself (the sign of the column maximum matters: masked
                                        // query rows compete with the value 0, see finalize_scores_kernel)
                    }
                    if (cbest > best) { best = cbest; best_t = tp_s[i * 32 + cj]; }
                    // butterfly transpose-reduce: after the 5 exchanges lane j holds the warp's winner of column j
#pragma unroll
                    for (int half = 16; half >= 1; half >>= 1) {
                        const bool up = (lane & half) != 0;
#pragma unroll
                        for (int ii = 0; ii < half; ++ii) {
                            const float send = up ? k[ii] : k[ii + half];
                            const float keep = up ? k[ii + half] : k[ii];
                            const float recv = __shfl_xor_sync(0xffffffffu, send, half);
                            k[ii] = fmaxf(keep, recv);
                        }
                    }
                    {
                        const uint32_t kb = __float_as_uint(k[0]);
                        const int win_lane = FAST ? 31 - (int)(kb & 31u) : ((kb >> 31) ? (int)(kb & 31u) : 31 - (int)(kb & 31u));
                        const float val = FAST ? __uint_as_float(kb & 0xFFFFFFE0u) - 2.0f : __uint_as_float(kb & 0xFFFFFFE0u);
                        if (lane < ncols && warp_row0 < T)
                            atomicMax(p.rowkey + bn * T + tp_s[i * 32 + lane], pack_key(val + 0.0f, (uint32_t)(warp_row0 + win_lane)));
                    }
                }
                if (tc.nt == nt_end - 1 && s_ok && best > -INFINITY)
                    atomicMax(p.colkey + bn * T + s_row, pack_key(best + 0.0f, (uint32_t)best_t));
            } else if (EPI == EPI_SIM) {
                // lane = query patch t, columns = template patches s: sim = acc * ra[t] * rb[s], times the template mask,
                // clamped at 0, written at out[b, s, h, w] with t = w * gh + h ("b (w h) c -> b c h w", utils/matching.py:25).
                // For a fixed column the lanes' addresses are gw floats apart: partial sectors that L2 merges (the whole
                // volume is written exactly once; it is 256 KB per detection at the native 16 x 16 grid).
                const int t = warp_row0 + lane;
                const bool row_ok = t < T;
                const float ra_t = row_ok ? __ldg(p.ra + (size_t)tc.b * T + t) : 0.f;
                const int th = row_ok ? t % p.gh : 0, tw = row_ok ? t / p.gh : 0;
                float* out_t = p.emit + (size_t)tc.b * T * T + (size_t)th * p.gw + tw;   // + s * T per column
                // per-column factor rb[s] * mask[s] of this warp's <= 2 chunks, staged before the accumulator is ready
                float* cf_s = rb_stage + e * (2 * EPI_COLS);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < EPI_CHUNKS; ++i) {
                    const int s = tc.nt * BLOCK_N + (hh + 4 * i) * 32 + lane;
                    float f = 0.f;
                    if (s < T) {
                        const int sy = s / p.gw, sx = s - sy * p.gw;
                        const float m = __ldg(p.cmask + ((size_t)tc.b * p.Hm + nearest_src(sy, p.Hm, p.gh)) * p.Wm + nearest_src(sx, p.Wm, p.gw));
                        f = __ldg(p.rb + (size_t)tc.b * T + s) * m;
                    }
                    cf_s[i * 32 + lane] = f;
                }
                __syncwarp();
                mbar_wait(tfull_bar(as), aphase, p.fault, 4, as);
                ptx::tc_fence_after();
#pragma unroll 1
                for (int i = 0; i < EPI_CHUNKS; ++i) {
                    const int c = hh + 4 * i;
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
                    ptx::tmem_ld_wait();
                    if (i == EPI_CHUNKS - 1) release_tmem();
                    const int s0 = tc.nt * BLOCK_N + c * 32;
                    if (s0 >= T || !row_ok) continue;
                    const int ncols = min(32, T - s0);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j < ncols) {
                            // (acc * ra) * (rb * m): same association as the reference up to the folding of m into rb
                            const float val = (__uint_as_float(v[j]) * ra_t) * cf_s[i * 32 + j];
                            out_t[(size_t)(s0 + j) * T] = val < 0.f ? 0.f : val;
                        }
                    }
                }
            } else {
                // EPI_EMIT: lane = query patch t, columns = template patches s; the scaled products go to HBM
                const int t = warp_row0 + lane;
                const bool row_ok = t < T;
                mbar_wait(tfull_bar(as), aphase, p.fault, 4, as);
                ptx::tc_fence_after();
#pragma unroll 1
                for (int i = 0; i < EPI_CHUNKS; ++i) {
                    const int c = hh + 4 * i;
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
                    ptx::tmem_ld_wait();
                    if (i == EPI_CHUNKS - 1) release_tmem();
                    const int s0 = tc.nt * BLOCK_N + c * 32;
                    if (s0 >= T) continue;  // warp-uniform: chunk entirely past the last template patch
                    const int ncols = min(32, T - s0);
                    if (row_ok && p.emit_tile_w > 0) {
                        // tiled key map: 8 consecutive keys (one row segment of a 4 x 8 tile) = 32 contiguous bytes
                        const int Wk = p.emit_tile_w, tpr = Wk >> 3;
                        float* slice = p.emit + (bn * T + t) * (size_t)T;
#pragma unroll
                        for (int gq = 0; gq < 4; ++gq) {
                            const int s = s0 + 8 * gq;
                            if (s < T) {  // T % 8 == 0 (checked on the host)
                                const int y = s / Wk, x = s - y * Wk;
                                float* dst = slice + (size_t)(((y >> 2) * tpr + (x >> 3)) * 32 + (y & 3) * 8);
                                *reinterpret_cast<float4*>(dst) =
                                    make_float4(__uint_as_float(v[8 * gq]) * p.emit_scale, __uint_as_float(v[8 * gq + 1]) * p.emit_scale,
                                                __uint_as_float(v[8 * gq + 2]) * p.emit_scale, __uint_as_float(v[8 * gq + 3]) * p.emit_scale);
                                *reinterpret_cast<float4*>(dst + 4) =
                                    make_float4(__uint_as_float(v[8 * gq + 4]) * p.emit_scale, __uint_as_float(v[8 * gq + 5]) * p.emit_scale,
                                                __uint_as_float(v[8 * gq + 6]) * p.emit_scale, __uint_as_float(v[8 * gq + 7]) * p.emit_scale);
                            }
                        }
                    } else if (row_ok) {
                        float* dst = p.emit + (bn * T + t) * (size_t)T + s0;
                        if (ncols == 32 && (T & 3) == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4*>(dst + j) =
                                    make_float4(__uint_as_float(v[j]) * p.emit_scale, __uint_as_float(v[j + 1]) * p.emit_scale,
                                                __uint_as_float(v[j + 2]) * p.emit_scale, __uint_as_float(v[j + 3]) * p.emit_scale);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < ncols) dst[j] = __uint_as_float(v[j]) * p.emit_scale;
                        }
                    }
                }
            }
          }
         }
        }
    }

    // ===================== teardown =====================
    ptx::tc_fence_before();
    if (CL > 1) ptx::cluster_sync(); else __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<CL>(tmem_base, 512);
}

// ---- host side ----------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

// 2-D bf16 row-major tensor (rows x Kp), box = (BLOCK_K x box_rows), 128-byte swizzle
static int make_tmap(CUtensorMap* map, const void* base, uint64_t rows, uint64_t kp, uint32_t box_rows) {
    auto fn = get_encode_fn();
    if (!fn) return fail(PP_ERR_DEVICE, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {kp, rows};
    cuuint64_t strides[1] = {kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PP_ERR_LAUNCH, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PP_OK;
}

static int* g_fault_host = nullptr;  // pinned, host-mapped: survives a trapped kernel
static int* g_fault_dev = nullptr;
static std::mutex g_fault_mutex;  // first use may come from several host threads at once
static int ensure_fault_buffer() {
    std::lock_guard<std::mutex> lock(g_fault_mutex);
    if (g_fault_dev) return PP_OK;
    int* host = nullptr;
    int* dev = nullptr;
    PP_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&host), 64 * sizeof(int), cudaHostAllocMapped));
    for (int i = 0; i < 64; ++i) host[i] = 0;
    const cudaError_t e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&dev), host, 0);
    if (e != cudaSuccess) {
        cudaFreeHost(host);
        return fail(PP_ERR_LAUNCH, "cudaHostGetDevicePointer failed: %s", cudaGetErrorString(e));
    }
    g_fault_host = host;
    g_fault_dev = dev;
    return PP_OK;
}

template <int CL, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
    using Cfg = GemmCfg<CL>;
    auto kern = match_gemm_kernel<CL, EPI>;
    PP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    long long clusters = sm_count() / CL;
    if (clusters > (long long)p.num_groups) clusters = p.num_groups;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(clusters * CL));
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    const bool prof = take_profile_events(&ev0, &ev1);
    if (prof) PP_CUDA(cudaEventRecord(ev0, st));
    PP_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
    if (prof) PP_CUDA(cudaEventRecord(ev1, st));
    count_launch();
    return PP_OK;
}

// Shared by pp_match_scores (EPI_MATCH) and pp_match_similarity (EPI_EMIT).
int run_match_gemm(int epi, const void* q_prep, const void* bank_prep, int64_t n_banks, const int32_t* bank_of_det,
                   int B, int N, int T, int Kp, const float* mrow, const int* tv, const int* rowmap, const float* ra,
                   const float* rb,
                   unsigned long long* rowkey, unsigned long long* colkey, float* emit, float emit_scale, int cluster,
                   cudaStream_t st, int emit_tile_w, const int32_t* det_order, const float* cmask, int Hm, int Wm, int gh) {
    PP_CHECK_ARG(emit_tile_w == 0 || (epi == EPI_EMIT && emit_tile_w % 8 == 0 && T % emit_tile_w == 0 && (T / emit_tile_w) % 4 == 0),
                 "tiled emit needs a key map of W %% 8 == 0 columns and H %% 4 == 0 rows (W=%d, T=%d)", emit_tile_w, T);
    PP_CHECK_ARG(Kp > 0 && Kp % BLOCK_K == 0, "Kp must be a positive multiple of %d (got %d)", BLOCK_K, Kp);
    PP_CHECK_ARG((reinterpret_cast<uintptr_t>(q_prep) & 127) == 0 && (reinterpret_cast<uintptr_t>(bank_prep) & 127) == 0,
                 "prepared operands must be 128-byte aligned");
    PP_CHECK_ARG((long long)n_banks * N * T < (1LL << 31) && (long long)B * T < (1LL << 31),
                 "operand row count exceeds the 2^31 TMA coordinate range; split the call");
    if (cluster & PP_MATCH_FAST_KEYS) {
        cluster &= ~PP_MATCH_FAST_KEYS;
        if (epi == EPI_MATCH) epi = EPI_MATCH_FAST;
    }
    if (cluster == 0) cluster = 2;
    PP_CHECK_ARG(cluster == 1 || cluster == 2, "cluster must be 0, 1 or 2 (got %d)", cluster);
    if (int rc = ensure_fault_buffer()) return rc;
    GemmParams p{};
    p.B = B;
    p.N = N;
    p.T = T;
    p.num_k_blocks = Kp / BLOCK_K;
    p.num_mt = (T + BLOCK_M * cluster - 1) / (BLOCK_M * cluster);
    p.num_nt = (T + BLOCK_N - 1) / BLOCK_N;
    // short launches deal single tiles (see decode_work): fewer than 64 tile-times per cluster on the dense upper bound
    const long long per_det = (long long)N * p.num_mt;
    const bool match_epi = epi == EPI_MATCH || epi == EPI_MATCH_FAST;
    p.tile_mode = match_epi && B <= MAX_DETS_PER_LAUNCH &&
                  (long long)B * per_det * p.num_nt < 64LL * (sm_count() / cluster);
    // detections per group: as many as share M-side tiles usefully (8) while leaving >= 4 groups per cluster
    int chunk = 1;
    if (det_order && bank_of_det && !p.tile_mode) {
        const long long want = (long long)B * per_det / (4LL * (sm_count() / cluster));
        chunk = (int)(want < 1 ? 1 : (want > 8 ? 8 : want));
    }
    p.chunk = chunk;
    p.det_order = chunk > 1 ? det_order : nullptr;
    // (tile mode: the dense upper bound of the tile count sizes the grid; the kernel takes the exact count from its prefix table)
    const long long groups_ll = p.tile_mode ? (long long)B * per_det * p.num_nt : (long long)((B + chunk - 1) / chunk) * per_det;
    PP_CHECK_ARG(groups_ll < (1LL << 31), "too many tile groups in one launch (%lld); split the detection batch", groups_ll);
    p.num_groups = (uint32_t)groups_ll;
    p.n_banks = (int)n_banks;
    p.bank_of_det = bank_of_det;
    p.mrow = mrow;
    p.tv = tv;
    p.rowmap = rowmap;
    p.ra = ra;
    p.rb = rb;
    p.rowkey = rowkey;
    p.colkey = colkey;
    p.emit = emit;
    p.emit_scale = emit_scale;
    p.emit_tile_w = emit_tile_w;
    p.cmask = cmask;
    p.Hm = Hm;
    p.Wm = Wm;
    p.gh = gh;
    p.gw = gh > 0 ? T / gh : 0;
    PP_CHECK_ARG(epi != EPI_SIM || (cmask && ra && rb && emit && gh > 0 && gh * p.gw == T && Hm > 0 && Wm > 0), "similarity epilogue: bad arguments");
    p.fault = g_fault_dev;
    if (p.num_groups == 0) return PP_OK;
    // A = M-side operand (box of 128 rows), B = N-side operand (box of 256 / cluster rows); EPI_MATCH puts the template
    // bank on the M side and the compact query rows on the N side, EPI_EMIT the other way round
    CUtensorMap ta, tb;
    const bool match = epi == EPI_MATCH || epi == EPI_MATCH_FAST;
    const void* m_op = match ? bank_prep : q_prep;
    const void* n_op = match ? q_prep : bank_prep;
    const uint64_t m_rows = match ? (uint64_t)n_banks * N * T : (uint64_t)B * T;
    const uint64_t n_rows = match ? (uint64_t)B * T : (uint64_t)n_banks * N * T;
    if (int rc = make_tmap(&ta, m_op, m_rows, (uint64_t)Kp, BLOCK_M)) return rc;
    if (int rc = make_tmap(&tb, n_op, n_rows, (uint64_t)Kp, BLOCK_N / cluster)) return rc;
    if (cluster == 1) {
        return epi == EPI_MATCH ? launch_gemm<1, EPI_MATCH>(ta, tb, p, st)
               : epi == EPI_MATCH_FAST ? launch_gemm<1, EPI_MATCH_FAST>(ta, tb, p, st)
               : epi == EPI_SIM ? launch_gemm<1, EPI_SIM>(ta, tb, p, st) : launch_gemm<1, EPI_EMIT>(ta, tb, p, st);
    }
    return epi == EPI_MATCH ? launch_gemm<2, EPI_MATCH>(ta, tb, p, st)
           : epi == EPI_MATCH_FAST ? launch_gemm<2, EPI_MATCH_FAST>(ta, tb, p, st)
           : epi == EPI_SIM ? launch_gemm<2, EPI_SIM>(ta, tb, p, st) : launch_gemm<2, EPI_EMIT>(ta, tb, p, st);
}

int fault_buffer(int** dev_ptr) {
    if (int rc = ensure_fault_buffer()) return rc;
    *dev_ptr = g_fault_dev;
    return PP_OK;
}

int read_fault_record(int* out5) {
    if (!g_fault_host) return 0;
    for (int i = 0; i < 5; ++i) out5[i] = g_fault_host[i];
    const int code = g_fault_host[0];
    g_fault_host[0] = 0;
    return code;
}

}  // namespace pp
