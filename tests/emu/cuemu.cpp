// TEST INFRASTRUCTURE — fiber scheduler of the SIMT logic emulator (see cuemu.h).
#include "cuemu.h"

dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace cuemu {

static State g_state;
State& st() { return g_state; }

static const size_t kStack = 256 * 1024;

static void set_ids(unsigned t) {
    State& s = g_state;
    s.cur = t;
    threadIdx.x = t % blockDim.x;
    threadIdx.y = (t / blockDim.x) % blockDim.y;
    threadIdx.z = t / (blockDim.x * blockDim.y);
}

void yield() {
    State& s = g_state;
    swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}

void barrier_wait(Barrier& b) {
    unsigned gen = b.gen;
    if (++b.arrived >= b.expected) {
        b.arrived = 0;
        b.gen++;
        return;
    }
    while (b.gen == gen) yield();
}

static void on_exit(unsigned t) {
    State& s = g_state;
    s.fibers[t].done = true;
    Warp& w = s.warps[t / 32];
    w.xbuf[t & 31] = 0;
    // a thread that returns no longer takes part in barriers (lenient, like hardware for exited threads)
    Barrier* bars[2] = {&w.bar, &s.block_bar};
    for (Barrier* b : bars) {
        if (b->expected) b->expected--;
        if (b->expected && b->arrived >= b->expected) {
            b->arrived = 0;
            b->gen++;
        }
    }
}

static void trampoline() {
    State& s = g_state;
    unsigned t = s.cur;
    s.body();
    on_exit(t);
    swapcontext(&s.fibers[t].ctx, &s.sched);
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, std::function<void()> body) {
    State& s = g_state;
    unsigned nt = block.x * block.y * block.z;
    if (nt == 0 || nt > 1024) {
        fprintf(stderr, "cuemu: bad block size %u\n", nt);
        abort();
    }
    gridDim = grid;
    blockDim = block;
    s.nthreads = nt;
    s.body = std::move(body);
    if (s.stacks.size() < kStack * nt) s.stacks.resize(kStack * nt);
    s.fibers.resize(nt);
    s.dyn_smem.assign(smem_bytes + 16, 0xCD);  // poisoned, like uninitialised shared memory
    unsigned nwarps = (nt + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                blockIdx = dim3(bx, by, bz);
                s.block_bar = Barrier();
                s.block_bar.expected = nt;
                s.warps.assign(nwarps, Warp());
                for (unsigned w = 0; w < nwarps; w++) {
                    s.warps[w].bar.expected = std::min(32u, nt - w * 32);
                    memset(s.warps[w].xbuf, 0, sizeof(s.warps[w].xbuf));
                }
                for (unsigned t = 0; t < nt; t++) {
                    Fiber& f = s.fibers[t];
                    f.done = false;
                    f.tid = t;
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = s.stacks.data() + kStack * t;
                    f.ctx.uc_stack.ss_size = kStack;
                    f.ctx.uc_link = nullptr;
                    makecontext(&f.ctx, trampoline, 0);
                }
                unsigned live = nt;
                unsigned long long spins = 0;
                while (live) {
                    unsigned progressed = 0;
                    for (unsigned t = 0; t < nt; t++) {
                        Fiber& f = s.fibers[t];
                        if (f.done) continue;
                        set_ids(t);
                        swapcontext(&s.sched, &f.ctx);
                        if (f.done) {
                            live--;
                            progressed++;
                        }
                    }
                    // every live fiber is parked at a rendezvous: only a deadlock if this repeats with no exits
                    if (!progressed && ++spins > 50000000ull) {
                        fprintf(stderr, "cuemu: no progress (deadlocked barrier?)\n");
                        abort();
                    }
                }
            }
}

}  // namespace cuemu
