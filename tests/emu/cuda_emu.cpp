// TEST INFRASTRUCTURE ONLY: fiber scheduler for tests/emu/cuda_emu.h
#include "cuda_emu.h"
#include <stdexcept>

thread_local EmuState emu_cur;

namespace {
struct Fiber {
    ucontext_t ctx;
    unsigned char* stack = nullptr;
    EmuState st;
    bool done = false;
};
const size_t STACK = 128 * 1024;
thread_local std::vector<unsigned char*> stack_pool;    // allocated once, reused by every launch
unsigned char* get_stack(size_t i) {
    while (stack_pool.size() <= i) stack_pool.push_back((unsigned char*)malloc(STACK));
    return stack_pool[i];
}
thread_local ucontext_t sched_ctx;
thread_local Fiber* running = nullptr;
thread_local const std::function<void()>* cur_body = nullptr;

void fiber_entry() {
    (*cur_body)();
    running->done = true;
    swapcontext(&running->ctx, &sched_ctx);
}
}  // namespace

// one scheduling step: give the other fibers of the block a turn
void emu_yield() {
    Fiber* f = running;
    f->st = emu_cur;
    swapcontext(&f->ctx, &sched_ctx);
    emu_cur = f->st;
}

// block barrier: a real arrival count (threads that already exited count as arrived, as on the GPU), so that code
// between two barriers may yield a different number of times in different warps (warp shuffles, see below)
static thread_local size_t bar_arrived = 0, bar_generation = 0, bar_live = 0;
static void bar_release_if_complete() {
    if (bar_live > 0 && bar_arrived >= bar_live) { bar_arrived = 0; ++bar_generation; }
}
void emu_syncthreads() {
    const size_t gen = bar_generation;
    ++bar_arrived;
    bar_release_if_complete();
    while (bar_generation == gen) emu_yield();
}

static double shfl_slots[2][4096];
static unsigned char shfl_parity[4096];
double emu_shfl_exchange(double v, int src_lane) {
    // two slot sets used alternately: a lane overwrites set p again only in its call after next, i.e. after the yield of the
    // call in between, by which every lane of the warp has finished reading set p -- one yield per exchange is enough
    const unsigned tid = emu_cur.tid.x;            // 1-D blocks only
    const unsigned p = shfl_parity[tid] ^= 1;
    shfl_slots[p][tid] = v;
    emu_yield();                                   // every lane of the warp has published (lanes run in lockstep)
    return shfl_slots[p][(tid & ~31u) + (unsigned)src_lane];
}

void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const size_t nthreads = (size_t)block.x * block.y * block.z;
    // the launch limits of the real device (sm_100): a launch that violates them fails on the GPU with "invalid configuration
    // argument" while a host loop would just run zero or too many iterations
    if (grid.x == 0 || grid.y == 0 || grid.z == 0 || nthreads == 0 || nthreads > 1024 || block.z > 64 || grid.y > 65535 || grid.z > 65535 ||
        smem > 227 * 1024) {
        fprintf(stderr, "emu_launch: invalid CUDA launch configuration grid=(%u,%u,%u) block=(%u,%u,%u) smem=%zu\n",
                grid.x, grid.y, grid.z, block.x, block.y, block.z, smem);
        abort();
    }
    std::vector<unsigned char> shared(smem + 64);
    std::vector<Fiber> fibers(nthreads);
    for (size_t i = 0; i < nthreads; ++i) fibers[i].stack = get_stack(i);
    cur_body = &body;
    static const bool reverse_blocks = [] { const char* e = getenv("DB_EMU_ORDER"); return e && e[0] == 'r'; }();
    for (unsigned bzi = 0; bzi < grid.z; ++bzi)
    for (unsigned byi = 0; byi < grid.y; ++byi)
    for (unsigned bxi = 0; bxi < grid.x; ++bxi) {
        // blocks of a grid run one after the other here; on the GPU their order is arbitrary: reversed with DB_EMU_ORDER=reverse
        const unsigned bz = reverse_blocks ? grid.z - 1 - bzi : bzi, by = reverse_blocks ? grid.y - 1 - byi : byi,
                       bx = reverse_blocks ? grid.x - 1 - bxi : bxi;
        std::memset(shared.data(), 0xA5, shared.size());   // poison shared memory
        std::memset(shfl_parity, 0, sizeof(shfl_parity));
        size_t i = 0;
        for (unsigned tz = 0; tz < block.z; ++tz)
        for (unsigned ty = 0; ty < block.y; ++ty)
        for (unsigned tx = 0; tx < block.x; ++tx, ++i) {
            Fiber& f = fibers[i];
            f.done = false;
            f.st.tid = {tx, ty, tz}; f.st.bid = {bx, by, bz}; f.st.bdim = block; f.st.gdim = grid;
            f.st.smem = shared.data();
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = STACK;
            f.ctx.uc_link = &sched_ctx;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        size_t remaining = nthreads;
        bar_arrived = 0; bar_live = nthreads;
        // DB_EMU_ORDER=reverse: the threads of a block take their turns in descending order.  Code that is correct only because a
        // lower-numbered thread happens to run first (a missing __syncthreads / __syncwarp) passes in one order and fails in the other.
        static const bool reverse_order = [] { const char* e = getenv("DB_EMU_ORDER"); return e && e[0] == 'r'; }();
        while (remaining) {
            size_t finished_this_round = 0, waiting = 0;
            for (size_t fi = 0; fi < fibers.size(); ++fi) {
                Fiber& f = fibers[reverse_order ? fibers.size() - 1 - fi : fi];
                if (f.done) continue;
                running = &f;
                emu_cur = f.st;
                swapcontext(&sched_ctx, &f.ctx);
                if (f.done) { ++finished_this_round; --bar_live; bar_release_if_complete(); } else { ++waiting; }
            }
            remaining -= finished_this_round;
            if (finished_this_round && waiting) {
                // some threads exited while others wait at a barrier: allowed in CUDA only if exited threads
                // never reach the barrier; continue scheduling the waiting ones.
            }
        }
    }
    running = nullptr;
}
