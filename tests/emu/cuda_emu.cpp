#include "cuda_emu.h"

#include <cstdlib>
#include <new>

thread_local dim3 threadIdx;
thread_local dim3 blockIdx;
dim3 blockDim;
dim3 gridDim;
// every __shared__ variable of the kernel headers lives in the ELF section "emu_shared" (cuda_emu.h); the linker
// brackets it with these two symbols
#ifndef HYP_EMU_PLAIN_SHARED
extern "C" {
extern char __start_emu_shared[];
extern char __stop_emu_shared[];
}
#endif
namespace emu {
BlockState* g_block = nullptr;
void* g_dyn_smem = nullptr;

void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body) {
    gridDim = grid;
    blockDim = block;
    const int nt = (int)block.x;
    // exact-size heap block: under -fsanitize=address (HYP_EMU_ASAN=1) an overrun of the dynamic
    // shared-memory request is reported the way compute-sanitizer would report it on the device
    void* smem = nullptr;
    if (posix_memalign(&smem, 64, dyn_smem ? dyn_smem : 1) != 0) throw std::bad_alloc();
    g_dyn_smem = smem;
    for (unsigned by = 0; by < grid.y; by++)
        for (unsigned bx = 0; bx < grid.x; bx++) {
            BlockState st;
            st.nthreads = nt;
            pthread_barrier_init(&st.bar, nullptr, nt);
            const int nw = (nt + 31) / 32;
            st.warp_bar.resize(nw);
            for (int w = 0; w < nw; w++) pthread_barrier_init(&st.warp_bar[w], nullptr, std::min(32, nt - 32 * w));
            st.shf.assign(nt, 0.0);
            g_block = &st;
            // Shared memory is undefined at the start of a block on the device; the statics that stand in for it would
            // otherwise keep the previous block's (plausible) values.  All-ones bytes = NaN for doubles, -1 for ints: a
            // kernel that reads a shared entry it did not write in THIS block now fails its test, as the NaN-poisoned
            // output buffers of the wrappers do for global memory.
#ifndef HYP_EMU_PLAIN_SHARED
            memset(__start_emu_shared, 0xFF, (size_t)(__stop_emu_shared - __start_emu_shared));
#endif
            if (dyn_smem) memset(smem, 0xFF, dyn_smem);
            std::vector<std::thread> th;
            th.reserve(nt);
            for (int t = 0; t < nt; t++)
                th.emplace_back([&, t] {
                    threadIdx = dim3(t, 0, 0);
                    blockIdx = dim3(bx, by, 0);
                    body();
                });
            for (auto& x : th) x.join();
            pthread_barrier_destroy(&st.bar);
            for (auto& wb : st.warp_bar) pthread_barrier_destroy(&wb);
        }
    g_block = nullptr;
    g_dyn_smem = nullptr;
    free(smem);
}
}  // namespace emu
