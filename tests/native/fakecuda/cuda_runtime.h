// TEST INFRASTRUCTURE.  A stand-in for <cuda_runtime.h> that lets the HOST side of the library (classeq2_b200/csrc/
// capi.cu: planning, packing, chunked pipeline, scatter, multi-device fan-out, workspaces) compile with g++ and run
// without a GPU, under AddressSanitizer / ThreadSanitizer.  "Device" memory is malloc'ed host memory (so a copy that
// runs past a device buffer is an ASan report), every "asynchronous" call completes before it returns (stream order
// cannot be violated here - what is checked is the host logic, its buffer sizes and its threads), events carry the
// host clock.  The kernels' launch functions are supplied by tests/native/fake_kernels.cpp.  Nothing of this is ever
// linked into the product.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

typedef int cudaError_t;
enum : int { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorInvalidConfiguration = 9,
             cudaErrorInvalidDevice = 101, cudaErrorUnknown = 999 };
struct FakeCudaStream { int device; };
struct FakeCudaEvent { double t_ms; };
typedef FakeCudaStream *cudaStream_t;
typedef FakeCudaEvent *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
enum cudaLimit { cudaLimitMaxL2FetchGranularity = 5 };
constexpr unsigned cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1;
struct cudaDeviceProp { int major, minor, multiProcessorCount; };
struct cudaPointerAttributes { cudaMemoryType type; };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };

namespace fakecuda {
inline std::mutex &mu() { static std::mutex m; return m; }
inline std::map<const char *, size_t> &pinned() { static std::map<const char *, size_t> m; return m; }   // cudaHostAlloc'ed ranges
inline int device_count() { const char *e = getenv("FAKE_CUDA_DEVICES"); const int n = e ? atoi(e) : 4; return n > 0 ? n : 4; }
inline int &current() { static thread_local int d = 0; return d; }
inline double now_ms() { using namespace std::chrono; return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count(); }
inline long &live_allocs() { static long n = 0; return n; }    // device + pinned allocations not yet freed (leak check of the tests)
inline bool &skip_copies() { static bool b = false; return b; }   // host-overhead timing (tools/micro/host_floor.cpp): a DMA costs the host nothing
}  // namespace fakecuda

inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorInvalidConfiguration ? "invalid configuration" : "fake CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = fakecuda::device_count(); return cudaSuccess; }
inline cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= fakecuda::device_count()) return cudaErrorInvalidDevice; fakecuda::current() = d; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d) {
    if (d < 0 || d >= fakecuda::device_count()) return cudaErrorInvalidDevice;
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
    return cudaSuccess;
}
inline cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) {
    *p = malloc(n ? n : 1);
    if (!*p) return cudaErrorMemoryAllocation;
    std::lock_guard<std::mutex> lk(fakecuda::mu());
    ++fakecuda::live_allocs();
    return cudaSuccess;
}
inline cudaError_t cudaFree(void *p) {
    if (p) { std::lock_guard<std::mutex> lk(fakecuda::mu()); --fakecuda::live_allocs(); }
    free(p);
    return cudaSuccess;
}
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) {
    *p = malloc(n ? n : 1);
    if (!*p) return cudaErrorMemoryAllocation;
    std::lock_guard<std::mutex> lk(fakecuda::mu());
    fakecuda::pinned()[static_cast<const char *>(*p)] = n ? n : 1;
    ++fakecuda::live_allocs();
    return cudaSuccess;
}
inline cudaError_t cudaFreeHost(void *p) {
    if (p) { std::lock_guard<std::mutex> lk(fakecuda::mu()); fakecuda::pinned().erase(static_cast<const char *>(p)); --fakecuda::live_allocs(); }
    free(p);
    return cudaSuccess;
}
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) {
    std::lock_guard<std::mutex> lk(fakecuda::mu());
    a->type = cudaMemoryTypeUnregistered;
    auto it = fakecuda::pinned().upper_bound(static_cast<const char *>(p));
    if (it != fakecuda::pinned().begin()) {
        --it;
        if (static_cast<const char *>(p) < it->first + it->second) a->type = cudaMemoryTypeHost;
    }
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n && !fakecuda::skip_copies()) memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new FakeCudaStream{fakecuda::current()}; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new FakeCudaEvent{0.0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t_ms = fakecuda::now_ms(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return cudaSuccess; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
