// Block/fiber scheduler for the CPU emulation (TEST INFRASTRUCTURE ONLY, see emul_cuda.h).
#include "emul_cuda.h"
#include <atomic>

namespace emu {
thread_local BlockCtx* tl_block = nullptr;
static const size_t kStack = 96 * 1024;
static const size_t kDynSmem = 256 * 1024;

static void fiber_entry() {
    BlockCtx* b = tl_block;
    Fiber* f = b->cur;
    (*b->body)();
    f->done = true;
    swapcontext(&f->ctx, &b->sched);
}

void sync() {
    BlockCtx* b = tl_block;
    swapcontext(&b->cur->ctx, &b->sched);
}

static void run_block(BlockCtx& b, std::vector<Fiber>& fibers, char* stacks) {
    const unsigned nt = b.bdim.x * b.bdim.y * b.bdim.z;
    for (unsigned t = 0; t < nt; ++t) {
        Fiber& f = fibers[t];
        f.done = false;
        f.tid = dim3(t % b.bdim.x, (t / b.bdim.x) % b.bdim.y, t / (b.bdim.x * b.bdim.y));
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = stacks + (size_t)t * kStack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, fiber_entry, 0);
    }
    unsigned alive = nt;
    while (alive) {
        for (unsigned t = 0; t < nt; ++t) {
            Fiber& f = fibers[t];
            if (f.done) continue;
            b.cur = &f;
            swapcontext(&b.sched, &f.ctx);
            if (f.done) --alive;
        }
    }
}

void launch_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    const unsigned nt = block.x * block.y * block.z;
    if (smem > kDynSmem) { fprintf(stderr, "emu: dynamic smem %zu too large\n", smem); abort(); }
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    const char* env = getenv("PKB_EMUL_THREADS");
    if (env) hw = (unsigned)atoi(env);
    size_t nworkers = nblocks < hw ? nblocks : hw;
    if (nworkers == 0) return;
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        std::vector<Fiber> fibers(nt);
        char* stacks = (char*)malloc((size_t)nt * kStack);
        unsigned char* dyn = nullptr;
        if (posix_memalign((void**)&dyn, 128, kDynSmem)) abort();
        BlockCtx b;
        b.bdim = block;
        b.gdim = grid;
        b.dyn = dyn;
        b.body = &body;
        tl_block = &b;
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= nblocks) break;
            b.bid = dim3((unsigned)(i % grid.x), (unsigned)((i / grid.x) % grid.y), (unsigned)(i / ((size_t)grid.x * grid.y)));
            run_block(b, fibers, stacks);
        }
        tl_block = nullptr;
        free(stacks);
        free(dyn);
    };
    if (nworkers == 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (size_t w = 0; w < nworkers; ++w) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
}
}  // namespace emu
