// TEST INFRASTRUCTURE ONLY -- see jb_emul_shim.h.  Builds tests/emul/libjiminy_b200_emul.so: the
// product's C ABI (jiminy_b200/csrc/jb_capi.cu + jb_plan.cpp, unchanged) on top of the thread
// emulation of a warp.
#include "jb_emul_shim.h"
#include <chrono>
#include <condition_variable>
#include <cstdio>

thread_local EmulDim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local double* emul_smem = nullptr;
namespace emul {
thread_local Warp* warp = nullptr;
thread_local int lane_id = 0;
int current_L = 1;

void launch(unsigned grid, unsigned block, size_t smem_bytes, int L, const std::function<void()>& body) {
    // shared memory of a real launch holds whatever the previous block left there: JB_EMUL_SMEM_FILL=nan (or a number)
    // poisons it so that a read-before-write shows up on the CPU
    double fill = 0.0;
    if (const char* e = std::getenv("JB_EMUL_SMEM_FILL")) fill = (e[0] == 'n' || e[0] == 'N') ? std::nan("") : std::atof(e);
    std::vector<double> smem((smem_bytes + 7) / 8 + 1, fill);
    const unsigned nwarps = (block + 31) / 32;
    for (unsigned bi = 0; bi < grid; ++bi) {
        std::vector<Warp> warps(nwarps);
        for (auto& w : warps) {
            w.L = L;
            for (int g = 0; g < 32 / L; ++g) w.bars.emplace_back(new std::barrier<>(L));
            w.full.reset(new std::barrier<>(32));
            for (auto& f : w.load_fence) f.reset(new std::barrier<>(32));
        }
        // The lanes of a block start together, like the threads of a real block do.  Without this the first lanes can be
        // through a whole (small-robot) step before the last ones are even created -- and the padding groups of the last
        // warp, which re-read the state of the last real env, would then read what that env has already written back
        // (different values on the two lanes of a group -> different trip counts -> a deadlock that no GPU can show:
        // there the loads of all 32 lanes are issued before any lane gets anywhere near the store phase).
        std::barrier<> start_line(block);
        std::vector<std::thread> threads;
        threads.reserve(block);
        for (unsigned ti = 0; ti < block; ++ti) {
            threads.emplace_back([&, ti]() {
                threadIdx.x = ti; blockIdx.x = bi; blockDim.x = block; gridDim.x = grid;
                emul_smem = smem.data();
                warp = &warps[ti / 32];
                lane_id = static_cast<int>(ti % 32);
                start_line.arrive_and_wait();
                body();
            });
        }
        // watchdog: woken as soon as the block is done (a condition variable, not a polling sleep: a launch must stay cheap)
        std::mutex dog_mutex;
        std::condition_variable dog_cv;
        bool finished = false;
        std::thread dog;
        if (const char* e = std::getenv("JB_EMUL_WATCHDOG")) {
            const int secs = std::atoi(e);
            if (secs > 0) dog = std::thread([&, secs]() {
                std::unique_lock<std::mutex> lock(dog_mutex);
                if (dog_cv.wait_for(lock, std::chrono::seconds(secs), [&] { return finished; })) return;
                std::fprintf(stderr, "[emul watchdog] block %u stuck (L = %d); lane: waiting mask syncs\n", bi, L);
                for (auto& w : warps)
                    for (int l = 0; l < 32; ++l)
                        std::fprintf(stderr, "  lane %2d: %d %08x %lld\n", l, w.waiting[l], w.last_mask[l], w.n_sync[l]);
                std::abort();
            });
        }
        for (auto& t : threads) t.join();
        if (dog.joinable()) {
            { std::lock_guard<std::mutex> lock(dog_mutex); finished = true; }
            dog_cv.notify_all();
            dog.join();
        }
    }
}
}  // namespace emul

#define JB_HOST_EMUL 1
#include "../../jiminy_b200/csrc/jb_capi.cu"
