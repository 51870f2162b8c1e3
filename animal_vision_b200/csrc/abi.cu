// libavb200: version, error reporting and host-side table builders.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "avb_common.cuh"

namespace avb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    cudaGetLastError();  // clear the sticky-less error state
    return AVB_E_CUDA;
}

int sm_count() {
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cached[dev].load();
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev].store(v);
    }
    return v;
}

// ---- per-kernel timing recorder (thread local; bench.py's roofline leg)
struct ProfRec {
    const char *name;
    cudaEvent_t e0, e1;
};
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec> *g_prof = nullptr;

bool profiling_on() { return g_prof_on; }

void profile_mark(const char *name, cudaStream_t st, bool begin) {
    if (!g_prof) g_prof = new std::vector<ProfRec>();
    if (begin) {
        ProfRec r{name, nullptr, nullptr};
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, st);
        g_prof->push_back(r);
    } else if (!g_prof->empty()) {
        cudaEventRecord(g_prof->back().e1, st);
    }
}

}  // namespace avb

extern "C" {

int avb_profile_begin(void) {
    avb::g_prof_on = true;
    if (avb::g_prof) avb::g_prof->clear();
    return AVB_OK;
}

int avb_profile_end(char *names, int name_stride, float *ms, int capacity) {
    avb::g_prof_on = false;
    if (!avb::g_prof) return 0;
    int n = 0;
    for (auto &r : *avb::g_prof) {
        cudaEventSynchronize(r.e1);
        float t = 0.f;
        cudaEventElapsedTime(&t, r.e0, r.e1);
        if (n < capacity && names && ms) {
            snprintf(names + (size_t)n * name_stride, name_stride, "%s", r.name);
            ms[n] = t;
            ++n;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    avb::g_prof->clear();
    return n;
}

int avb_version(void) { return AVB_VERSION; }

#ifndef AVB_HEADER_SHA
#define AVB_HEADER_SHA "unknown"
#endif
const char *avb_header_sha(void) { return AVB_HEADER_SHA; }

const char *avb_last_error(void) { return avb::g_err; }

int avb_build_encode_table(const float *thr_host, uint32_t *table_host, int capacity) {
    if (!thr_host || !table_host) {
        avb::set_error("avb_build_encode_table: null pointer");
        return AVB_E_ARG;
    }
    uint32_t tb[255];
    for (int i = 0; i < 255; ++i) {
        memcpy(&tb[i], &thr_host[i], 4);
        bool ok = thr_host[i] > 0.0f && thr_host[i] <= 1.0f && (i == 0 || tb[i] > tb[i - 1]);
        if (!ok) {
            avb::set_error("avb_build_encode_table: thresholds must be strictly increasing in (0,1] (index %d)", i);
            return AVB_E_ARG;
        }
    }
    const uint32_t one = 0x3f800000u;
    // keep as few mantissa bits as possible while no bucket holds two thresholds
    for (int mant = 5; mant <= 12; ++mant) {
        const uint32_t shift = 23 - mant;
        bool clash = false;
        for (int i = 1; i < 255 && !clash; ++i) clash = (tb[i] >> shift) == (tb[i - 1] >> shift);
        if (clash) continue;
        const uint32_t key_min = (tb[0] >> shift) - 1;  // first bucket lies wholly below thr[1]
        const uint32_t nb = (one >> shift) - key_min + 1;
        if ((int)(nb + avb::ENC_HEADER) > capacity) break;
        table_host[0] = key_min;
        table_host[1] = shift;
        table_host[2] = nb;
        table_host[3] = 0;
        int next = 0;  // index of the first threshold not yet passed
        for (uint32_t k = 0; k < nb; ++k) {
            const uint32_t key = key_min + k;
            uint32_t low = 1u << shift;  // "no threshold in this bucket"
            const uint32_t byte_at_start = (uint32_t)next;
            if (next < 255 && (tb[next] >> shift) == key) {
                low = tb[next] & ((1u << shift) - 1u);
                ++next;
            }
            table_host[avb::ENC_HEADER + k] = (low << 8) | byte_at_start;
        }
        if (next != 255) {
            avb::set_error("avb_build_encode_table: internal error (%d thresholds placed)", next);
            return AVB_E_ARG;
        }
        return (int)(nb + avb::ENC_HEADER);
    }
    avb::set_error("avb_build_encode_table: thresholds too dense for the bucketed table");
    return AVB_E_UNSUPPORTED;
}

}  // extern "C"
