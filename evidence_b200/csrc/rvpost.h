// rvpost.h — per-device scratch shared by the post-processing entry points (rvfip.cu, rvorder.cu).
//
// rvl_fip_accumulate / rvl_order_planets are handle-less host-buffer calls that a post-processing
// script makes once per (run, k) block: allocating and freeing their device buffers and events on
// every call cost 6-13 ms around a 0.1-0.3 ms kernel.  The buffers are now kept per
// device and slot, grown on demand, and released by rvl_post_release() (or at process exit).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <utility>

namespace rvpost {

struct Slot {
    void *p = nullptr;
    size_t cap = 0;
};
struct DeviceScratch {
    std::map<int, Slot> slots;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};
inline std::mutex g_mutex;
inline std::map<int, DeviceScratch> g_scratch;

// the caller holds g_mutex for the duration of its call (one post-processing call at a time)
inline cudaError_t get(int device, int slot, size_t bytes, void **out)
{
    Slot &s = g_scratch[device].slots[slot];
    if (bytes > s.cap) {
        if (s.p) cudaFree(s.p);
        s.p = nullptr;
        s.cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        const cudaError_t e = cudaMalloc(&s.p, want);
        if (e != cudaSuccess) return e;
        s.cap = want;
    }
    *out = s.p;
    return cudaSuccess;
}
inline cudaError_t events(int device, cudaEvent_t *e0, cudaEvent_t *e1)
{
    DeviceScratch &d = g_scratch[device];
    if (!d.ev0) {
        cudaError_t e = cudaEventCreate(&d.ev0);
        if (e != cudaSuccess) return e;
        e = cudaEventCreate(&d.ev1);
        if (e != cudaSuccess) return e;
    }
    *e0 = d.ev0;
    *e1 = d.ev1;
    return cudaSuccess;
}
inline void release_all()
{
    std::lock_guard<std::mutex> lock(g_mutex);
    int prev = 0;
    cudaGetDevice(&prev);
    for (auto &kv : g_scratch) {
        cudaSetDevice(kv.first);
        for (auto &s : kv.second.slots)
            if (s.second.p) cudaFree(s.second.p);
        if (kv.second.ev0) cudaEventDestroy(kv.second.ev0);
        if (kv.second.ev1) cudaEventDestroy(kv.second.ev1);
    }
    g_scratch.clear();
    cudaSetDevice(prev);
}

}  // namespace rvpost
