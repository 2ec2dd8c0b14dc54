// simt_emu.cpp — runtime of the CPU SIMT emulator (TEST INFRASTRUCTURE ONLY, see simt_emu.h).
#include "simt_emu.h"

namespace simt {

State g;

static void fibre_main() {
    g.body();
    g.fibres[g.cur].done = true;
    swapcontext(&g.fibres[g.cur].ctx, &g.sched);
}

void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, std::function<void()> body) {
    const unsigned nthreads = block.x * block.y * block.z;
    if (nthreads % kWarp) { fprintf(stderr, "simt: block size must be a multiple of 32\n"); abort(); }
    g.bdim = block;
    g.gdim = grid;
    g.body = body;
    std::vector<uint8_t> smem(dyn_smem_bytes + 16);
    g.dyn_smem = (uint8_t*)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    std::vector<uint8_t*> stacks(nthreads);
    for (auto& s : stacks) s = (uint8_t*)malloc(kStack);
    for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++) {
        g.bid = uint3{bx, by, bz};
        g.fibres.assign(nthreads, Fibre());
        g.warps.assign(nthreads / kWarp, WarpSlot());
        g.bar_arrived = 0;
        g.bar_generation = 0;
        for (unsigned t = 0; t < nthreads; t++) {
            Fibre& f = g.fibres[t];
            f.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            f.stack = stacks[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = &g.sched;
            makecontext(&f.ctx, fibre_main, 0);
        }
        unsigned remaining = nthreads;
        uint64_t idle_rounds = 0;
        while (remaining) {
            const uint64_t before = g.collectives + g.bar_generation;
            const unsigned rem_before = remaining;
            for (unsigned t = 0; t < nthreads; t++) {
                if (g.fibres[t].done) continue;
                g.cur = (int)t;
                swapcontext(&g.sched, &g.fibres[t].ctx);
                if (g.fibres[t].done) remaining--;
            }
            if (g.collectives + g.bar_generation == before && remaining == rem_before) {
                if (++idle_rounds > 4) {
                    fprintf(stderr, "simt: deadlock (lanes waiting on a collective that never completes)\n");
                    for (size_t w = 0; w < g.warps.size(); w++) {
                        unsigned live = 0;
                        for (int l = 0; l < kWarp; l++) live += !g.fibres[w * kWarp + l].done;
                        fprintf(stderr, "  warp %zu: live %u, pending op %d with %u arrivals; cta barrier arrivals %u\n", w, live,
                                g.warps[w].op, g.warps[w].arrived, g.bar_arrived);
                    }
                    abort();
                }
            } else idle_rounds = 0;
        }
        g.cur = -1;
    }
    for (auto s : stacks) free(s);
    g.dyn_smem = nullptr;
}

}  // namespace simt
