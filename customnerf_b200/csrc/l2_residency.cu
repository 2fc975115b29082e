// l2_residency.cu -- keeping the hash table resident in the 126 MB L2 across the kernels of a train step.
//
// BASELINE.json north_star, kernel 1: "per-level tables staged so the working set stays L2-resident".  The reference relies on
// whatever the cache happens to keep (its comment at gridencoder.cu:387 claims one level is cached at a time; the launch puts all
// levels in one grid).  A B200 train step streams ~350 MB of activations through the L2 between two uses of the 47 MB table, so
// without help every gather of the next encode pass starts from HBM.  CUDA's access-policy window does the staging in
// hardware: a carve-out of the L2 is set aside for PERSISTING lines (cudaLimitPersistingL2CacheSize) and every kernel launched
// on a stream whose window covers [base, base + bytes) keeps a `hit_ratio` share of the lines it touches there in the carve-
// out; accesses outside the window are treated as streaming.  Kernel nodes captured into a CUDA graph inherit the window of
// the capturing stream.
#include "common.cuh"
#include <string.h>

extern "C" {

// carve-out for persisting lines; returns what the device granted in *granted (0: the device has no persisting L2)
int nb200_l2_persist_limit(uint64_t bytes, uint64_t *granted, uint64_t *max_window) {
    int dev = 0, max_persist = 0, max_win = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    if (max_window) *max_window = (uint64_t)max_win;
    if (bytes > (uint64_t)max_persist) bytes = (uint64_t)max_persist;
    cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)bytes);
    if (e != cudaSuccess) { cudaGetLastError(); if (granted) *granted = 0; return (int)e; }
    size_t got = 0;
    cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
    if (granted) *granted = (uint64_t)got;
    return 0;
}

// window of every later launch on `stream`: bytes == 0 removes it
int nb200_stream_access_window(void *stream, const void *base, uint64_t bytes, float hit_ratio) {
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = const_cast<void *>(base);
    attr.accessPolicyWindow.num_bytes = (size_t)bytes;
    attr.accessPolicyWindow.hitRatio = bytes ? hit_ratio : 0.0f;
    attr.accessPolicyWindow.hitProp = bytes ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
    attr.accessPolicyWindow.missProp = bytes ? cudaAccessPropertyStreaming : cudaAccessPropertyNormal;
    cudaError_t e = cudaStreamSetAttribute(nb_stream(stream), cudaStreamAttributeAccessPolicyWindow, &attr);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return 0;
}

// turn every persisting line back into a normal one (call when the table is freed / replaced)
int nb200_l2_persist_reset(void) {
    cudaError_t e = cudaCtxResetPersistingL2Cache();
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return 0;
}

}  // extern "C"
