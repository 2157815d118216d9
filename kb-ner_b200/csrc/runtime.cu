// Library-level plumbing of the C ABI: error text, launch counter, device check, TMA maps.
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "tma_host.cuh"

namespace kbner {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int num_sms() {
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMs;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

static std::atomic<int> g_sm_budget{0};
int sm_budget() {
    const int n = num_sms(), b = g_sm_budget.load(std::memory_order_relaxed);
    return (b > 0 && b < n) ? b : n;
}

static bool nvtx_on() {
    static const bool on = [] {
        const char *e = getenv("KBNER_NVTX");
        return e && e[0] == '1';
    }();
    return on;
}
void nvtx_push(const char *name) {
    if (nvtx_on()) nvtxRangePushA(name);
}
void nvtx_pop() {
    if (nvtx_on()) nvtxRangePop();
}

bool pdl_enabled() {
    static const bool on = [] {
        const char *e = getenv("KBNER_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)p;
    });
    return fn;
}

struct TmapKey {
    const void *base;
    uint64_t rows, cols, ld;
    uint32_t br, bc, eb;
    bool operator==(const TmapKey &o) const {
        return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && br == o.br && bc == o.bc && eb == o.eb;
    }
};
struct TmapHash {
    size_t operator()(const TmapKey &k) const {
        size_t h = (size_t)k.base;
        h = h * 1000003u ^ k.rows;
        h = h * 1000003u ^ k.cols;
        h = h * 1000003u ^ k.ld;
        h = h * 1000003u ^ ((uint64_t)k.br << 32 | k.bc << 4 | k.eb);
        return h;
    }
};

int make_tmap_2d(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 uint32_t box_cols, uint32_t elem_bytes) {
    static std::mutex mu;
    static std::unordered_map<TmapKey, CUtensorMap, TmapHash> cache;
    TmapKey key{base, rows, cols, ld, box_rows, box_cols, elem_bytes};
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return KBNER_OK;
        }
    }
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)");
        return KBNER_ECUDA;
    }
    if (((uintptr_t)base & 15u) != 0 || (ld * elem_bytes) % 16 != 0 || box_cols * elem_bytes != 128 || box_rows > 256 ||
        (elem_bytes != 2 && elem_bytes != 4)) {
        set_error("tensor map: base must be 16-B aligned, row pitch a multiple of 16 B, box 128 B x <=256 rows "
                  "(got ld=%llu box=%ux%u elem=%u)", (unsigned long long)ld, box_rows, box_cols, elem_bytes);
        return KBNER_EINVAL;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * elem_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return KBNER_ECUDA;
    }
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
    return KBNER_OK;
}

}  // namespace kbner

extern "C" int kbner_abi_version(void) { return 2; }

extern "C" int kbner_set_sm_budget(int n_sms) {
    if (n_sms < 0) {
        kbner::set_error("set_sm_budget: %d", n_sms);
        return KBNER_EINVAL;
    }
    kbner::g_sm_budget.store(n_sms & ~1, std::memory_order_relaxed);      // CTA pairs: an even number of SMs
    return KBNER_OK;
}
extern "C" const char *kbner_last_error(void) { return kbner::g_err; }
extern "C" uint64_t kbner_launch_count(void) { return kbner::g_launches.load(); }
extern "C" void kbner_add_launches(uint64_t n) { kbner::count_launch((int)n); }
extern "C" int kbner_device_check(int dev) {
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) {
        kbner::set_error("cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
        return KBNER_ENODEVICE;
    }
    if (p.major != 10) {
        kbner::set_error("device %d is sm_%d%d; kbner_b200 kernels are built for sm_100a only", dev, p.major, p.minor);
        return KBNER_ENODEVICE;
    }
    return KBNER_OK;
}
