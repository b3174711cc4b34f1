// C-ABI plumbing of libpicopose_b200: error reporting, device gate, fault read-back.
#include "pp_common.cuh"

#include <atomic>
#include <cstring>

namespace pp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;
bool take_profile_events(cudaEvent_t* start, cudaEvent_t* stop) {
    if (!g_prof_start || !g_prof_stop) return false;
    *start = g_prof_start;
    *stop = g_prof_stop;
    g_prof_start = g_prof_stop = nullptr;
    return true;
}

static int g_sm_count[64];
static int g_dev_ok[64];  // 0 = unknown, 1 = sm_100, -1 = something else

static int probe_device(int* dev_out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(PP_ERR_DEVICE, "no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e));
    if (dev < 0 || dev >= 64) return fail(PP_ERR_DEVICE, "device ordinal %d out of range", dev);
    if (g_dev_ok[dev] == 0) {
        cudaDeviceProp prop;
        e = cudaGetDeviceProperties(&prop, dev);
        if (e != cudaSuccess) return fail(PP_ERR_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        g_sm_count[dev] = prop.multiProcessorCount;
        g_dev_ok[dev] = (prop.major == 10) ? 1 : -1;
        if (g_dev_ok[dev] < 0)
            set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, prop.major, prop.minor);
    }
    *dev_out = dev;
    return PP_OK;
}

int require_sm100() {
    int dev;
    if (int rc = probe_device(&dev)) return rc;
    if (g_dev_ok[dev] < 0) {
        return fail(PP_ERR_DEVICE, "current device is not sm_100 (B200); libpicopose_b200 has no fallback path");
    }
    return PP_OK;
}

bool pdl_enabled() {
    const char* e = getenv("PICOPOSE_B200_PDL");
    return !(e && e[0] == '0');
}

int sm_count() {
    int dev;
    if (probe_device(&dev)) return 148;
    return g_sm_count[dev] > 0 ? g_sm_count[dev] : 148;
}

int read_fault_record(int* out5);

}  // namespace pp

extern "C" int pp_version(void) { return PP_VERSION; }

extern "C" long long pp_launch_count(void) { return pp::g_launches.load(std::memory_order_relaxed); }

extern "C" void pp_profile_gemm_events(void* start_event, void* stop_event) {
    pp::g_prof_start = static_cast<cudaEvent_t>(start_event);
    pp::g_prof_stop = static_cast<cudaEvent_t>(stop_event);
}

extern "C" const char* pp_last_error(void) { return pp::g_err; }

extern "C" int pp_check_device_faults(void) {
    cudaError_t e = cudaDeviceSynchronize();
    int rec[5] = {0, 0, 0, 0, 0};
    const int code = pp::read_fault_record(rec);
    if (code == 5)
        return pp::fail(PP_ERR_KERNEL, "exchange timeout: peer %d never published epoch %u (block %d); the outputs of that call "
                                       "are poisoned (NaN / -1).  PICOPOSE_B200_XCHG_TIMEOUT_S sets the limit (%s)",
                        rec[3], (unsigned)rec[4], rec[1], cudaGetErrorString(e));
    if (code == 6)
        return pp::fail(PP_ERR_KERNEL, "bank index out of range: detection %d names bank %d of %d (clamped; results of that "
                                       "detection are meaningless)", rec[1], rec[2], rec[3]);
    if (code != 0) {
        // codes: 1 producer waits for a free smem stage, 2 MMA waits for a drained TMEM stage,
        //        3 MMA waits for TMA data, 4 epilogue waits for the accumulator
        return pp::fail(PP_ERR_KERNEL, "pipeline timeout: code %d block %d thread %d aux %d parity %d (%s)", code, rec[1],
                        rec[2], rec[3], rec[4], cudaGetErrorString(e));
    }
    if (e != cudaSuccess) return pp::fail(PP_ERR_LAUNCH, "device error: %s", cudaGetErrorString(e));
    return PP_OK;
}
