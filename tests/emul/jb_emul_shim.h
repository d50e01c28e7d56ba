// TEST INFRASTRUCTURE ONLY -- never part of the product, never loaded by jiminy_b200/.
//
// Compiles the device code of jiminy_b200/csrc unchanged for the host and runs every CUDA thread
// of a block as a std::thread, with warp shuffles / votes / __syncwarp implemented as rendez-vous
// between the lanes named in the mask.  This lets the CPU test-suite (no GPU in the build
// container) exercise the *same* kernel source -- lane plan, trunk all-reduce, scheduler -- against
// the oracle before any GPU time is spent.  The numbers it produces are not performance data and
// are never reported as such.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <memory>
#include <thread>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)

struct EmulDim3 { unsigned x = 1, y = 1, z = 1; };
extern thread_local EmulDim3 threadIdx, blockIdx, blockDim, gridDim;
extern thread_local double* emul_smem;

// `extern __shared__ double smem[];` inside the kernel
#define __shared__
#define EMUL_SHARED_DECL
namespace emul {
struct Warp {
    // exchange slots for the 32 lanes + one barrier per distinct mask in flight; masks used by the
    // kernels are contiguous aligned groups, so a barrier per (group size, group index) suffices.
    double dslot[32];
    int islot[32];
    // watchdog (JB_EMUL_WATCHDOG=seconds): what every lane last waited on
    unsigned last_mask[32] = {};
    long long n_sync[32] = {};
    int waiting[32] = {};
    std::vector<std::unique_ptr<std::barrier<>>> bars;   // indexed by group id for the current L
    std::unique_ptr<std::barrier<>> full;                 // whole-warp barrier (mask 0xffffffff)
    std::map<unsigned, std::unique_ptr<std::barrier<>>> other;   // any other mask (unions of groups: the scheduler's votes)
    // "every lane of the warp has loaded its env state": one per pass of the step body (hot-path pass, full pass), see
    // load_fence_wait() below
    std::unique_ptr<std::barrier<>> load_fence[2];
    std::mutex other_mutex;
    int L = 1;
};
extern thread_local Warp* warp;
extern thread_local int lane_id;
inline int group_of(unsigned mask) { return __builtin_ctz(mask) / warp->L; }
inline void sync_group_(unsigned mask);
inline void sync_group(unsigned mask) {
    warp->last_mask[lane_id] = mask; ++warp->n_sync[lane_id]; warp->waiting[lane_id] = 1;
    sync_group_(mask);
    warp->waiting[lane_id] = 0;
}
inline void sync_group_(unsigned mask) {
    if (mask == 0xffffffffu && warp->L < 32) { warp->full->arrive_and_wait(); return; }
    const int g = group_of(mask);
    const unsigned gm = warp->L >= 32 ? 0xffffffffu : (((1u << warp->L) - 1u) << (g * warp->L));
    if (mask == gm) { warp->bars[g]->arrive_and_wait(); return; }
    std::barrier<>* b;
    {
        std::lock_guard<std::mutex> lock(warp->other_mutex);
        auto& slot = warp->other[mask];
        if (!slot) slot.reset(new std::barrier<>(__builtin_popcount(mask)));
        b = slot.get();
    }
    b->arrive_and_wait();
}
}  // namespace emul

// The padding env groups of the last warp re-read the state of the last real env (c.env = n_env - 1), which that env
// writes back at the end of its step.  On the GPU the loads of all the lanes of a warp are one instruction, issued long
// before any lane reaches the store phase; here the lanes are OS threads that can be descheduled for longer than a whole
// step, so two lanes of a padding group could load different values of (t, dt, ...) and run different numbers of loop
// iterations -- a deadlock of the emulation only.  The step body therefore tells the emulator when a lane is through its
// loads (wait: nobody goes on to compute, let alone store, before every lane has loaded) or leaves the body before them
// (drop).  Both are no-ops in the device build.
namespace emul {
inline void load_fence_wait(int pass) { warp->load_fence[pass]->arrive_and_wait(); }
inline void load_fence_drop(int pass) { warp->load_fence[pass]->arrive_and_drop(); }
}  // namespace emul

inline void __syncwarp(unsigned mask = 0xffffffffu) { emul::sync_group(mask); }
inline double __shfl_xor_sync(unsigned mask, double x, int o) {
    emul::warp->dslot[emul::lane_id] = x;
    emul::sync_group(mask);
    const double r = emul::warp->dslot[emul::lane_id ^ o];
    emul::sync_group(mask);
    return r;
}
inline int __shfl_xor_sync(unsigned mask, int x, int o) {
    emul::warp->islot[emul::lane_id] = x;
    emul::sync_group(mask);
    const int r = emul::warp->islot[emul::lane_id ^ o];
    emul::sync_group(mask);
    return r;
}
inline double __shfl_sync(unsigned mask, double x, int src) {
    emul::warp->dslot[emul::lane_id] = x;
    emul::sync_group(mask);
    const double r = emul::warp->dslot[src & 31];
    emul::sync_group(mask);
    return r;
}
inline bool __any_sync(unsigned mask, bool p) {
    emul::warp->islot[emul::lane_id] = p ? 1 : 0;
    emul::sync_group(mask);
    bool r = false;
    for (int l = 0; l < 32; ++l)
        if (mask & (1u << l)) r = r || (emul::warp->islot[l] != 0);
    emul::sync_group(mask);
    return r;
}
inline bool __all_sync(unsigned mask, bool p) { return !__any_sync(mask, !p); }
inline unsigned __ballot_sync(unsigned mask, bool p) {
    emul::warp->islot[emul::lane_id] = p ? 1 : 0;
    emul::sync_group(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if ((mask & (1u << l)) && emul::warp->islot[l] != 0) r |= 1u << l;
    emul::sync_group(mask);
    return r;
}
// blocks run one after the other in the emulator: plain read-modify-write is enough
inline unsigned int atomicOr(unsigned int* p, unsigned int v) { const unsigned int o = *p; *p = o | v; return o; }
inline unsigned int atomicAnd(unsigned int* p, unsigned int v) { const unsigned int o = *p; *p = o & v; return o; }
inline double __longlong_as_double(long long x) { double d; std::memcpy(&d, &x, sizeof d); return d; }
inline void sincos(double x, double* s, double* c) { *s = std::sin(x); *c = std::cos(x); }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }

// ---- minimal CUDA runtime stand-ins (host memory, synchronous "streams")
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(1, n); return cudaSuccess; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::calloc(1, n); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < height; ++r) std::memmove(static_cast<char*>(d) + r * dpitch, static_cast<const char*>(s) + r * spitch, width);
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = reinterpret_cast<void*>(1); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

namespace emul {
// Runs `body()` for every thread of every block, 32 lanes of a warp concurrently.
void launch(unsigned grid, unsigned block, size_t smem_bytes, int lanes_per_group, const std::function<void()>& body);
extern int current_L;   // lanes per env of the batch being launched (set by the launch macro)
}  // namespace emul

#define JB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emul::launch((grid), (block), (smem), emul::current_L, [&]() { kernel(__VA_ARGS__); })
