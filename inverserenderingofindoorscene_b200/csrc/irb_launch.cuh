// irb_launch.cuh -- host-side launch helpers of the persistent kernels: everything a launch needs besides the kernel
// arguments is computed ONCE per (kernel, shared-memory size, device) or per (tensor base, shape) and cached, so that a
// small launch (2 images per GPU in config 3 over 8 GPUs) is not dominated by cudaFuncSetAttribute, an occupancy query
// and 7-9 cuTensorMapEncodeTiled calls.
#pragma once
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "irb_tma.cuh"

namespace irb {

inline int cached_sms() {
    static thread_local int dev_cached = -1, sms_cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != dev_cached) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        sms_cached = sms > 0 ? sms : 148;
        dev_cached = dev;
    }
    return sms_cached;
}

// resident CTAs per SM of `kernel` at (threads, dynamic smem) on the current device; sets the dynamic shared-memory
// attribute the first time.  Returns a cudaError_t (0 = ok).
template <typename KernelT> inline int persistent_ctas_per_sm(KernelT kernel, int threads, size_t smem, int* per_sm) {
    struct Key {
        const void* k;
        size_t smem;
        int threads, dev;
        bool operator==(const Key& o) const { return k == o.k && smem == o.smem && threads == o.threads && dev == o.dev; }
    };
    struct Hash {
        size_t operator()(const Key& x) const {
            return std::hash<const void*>()(x.k) ^ (x.smem * 0x9e3779b97f4a7c15ull) ^ ((size_t)x.threads << 40) ^ ((size_t)x.dev << 52);
        }
    };
    static std::mutex mu;
    static std::unordered_map<Key, int, Hash> cache;       // (kernel, smem, threads, device) -> resident CTAs per SM
    static std::unordered_map<Key, size_t, Hash> attr;     // (kernel, device) -> largest dynamic smem size set so far
    int dev = 0;
    cudaGetDevice(&dev);
    const Key key{(const void*)kernel, smem, threads, dev};
    const Key akey{(const void*)kernel, 0, 0, dev};
    bool need_attr = true;
    {
        std::lock_guard<std::mutex> g(mu);
        auto ia = attr.find(akey);
        need_attr = (ia == attr.end()) || ia->second < smem;   // the attribute is a MAXIMUM: only ever raise it
        auto it = cache.find(key);
        if (it != cache.end() && !need_attr) {
            *per_sm = it->second;
            return 0;
        }
    }
    if (need_attr) {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    int n = 1;
    const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem);
    if (e != cudaSuccess || n < 1) n = 1;
    {
        std::lock_guard<std::mutex> g(mu);
        cache[key] = n;
        if (need_attr) attr[akey] = smem;
    }
    *per_sm = n;
    return 0;
}

// make_tile_map with a small per-thread cache keyed by (base, plane, rows, nb, tile): PyTorch's caching allocator hands
// the same addresses back step after step, so a training loop encodes each descriptor once.
inline bool cached_tile_map(CUtensorMap* m, const float* base, size_t plane, unsigned rows, unsigned nb, unsigned tile) {
    struct Entry {
        const float* base;
        size_t plane;
        unsigned rows, nb, tile;
        bool valid;
        CUtensorMap map;
    };
    constexpr int N = 64;
    static thread_local Entry tab[N];
    const size_t h = ((uintptr_t)base >> 4) * 0x9e3779b97f4a7c15ull ^ plane * 31 ^ ((size_t)rows << 20) ^ ((size_t)nb << 32) ^ tile;
    Entry& e = tab[(h >> 17) % N];
    if (e.valid && e.base == base && e.plane == plane && e.rows == rows && e.nb == nb && e.tile == tile) {
        memcpy(m, &e.map, sizeof(CUtensorMap));
        return true;
    }
    if (!make_tile_map(m, base, plane, rows, nb, tile)) return false;
    e.base = base; e.plane = plane; e.rows = rows; e.nb = nb; e.tile = tile; e.valid = true;
    memcpy(&e.map, m, sizeof(CUtensorMap));
    return true;
}

}  // namespace irb
