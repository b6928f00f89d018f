// DEVELOPER / TEST HARNESS ONLY -- see cuda_emu.hpp.
#include "cuda_emu.hpp"
#include <algorithm>
#include <cstdio>
#include <condition_variable>
#include <mutex>
#include <sys/mman.h>
#include <thread>
#include <vector>

namespace zkemu {
thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;

namespace {
constexpr size_t kStack = 256 * 1024;
constexpr size_t kMaxThreads = 1024;
constexpr size_t kDynSmem = 232 * 1024;

struct Fiber {
    void *sp = nullptr;
    bool done = true;
    dim3 tid;
};

struct Worker {
    uint8_t *stack_base = nullptr;    // kMaxThreads * kStack, mmap'ed once (pages are touched lazily)
    std::vector<Fiber> fibers;
    void *sched_sp = nullptr;
    int cur = -1;
    const std::function<void()> *body = nullptr;
    std::vector<uint8_t> smem;
    std::vector<uint64_t> xbuf;
};
thread_local Worker *t_w = nullptr;

// switch_ctx(save_sp, load_sp): saves callee-saved registers on the current stack, stores rsp, loads the other.
extern "C" void zkemu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl zkemu_switch
.type zkemu_switch,@function
zkemu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size zkemu_switch,.-zkemu_switch
)");

void fiber_entry() {
    Worker *w = t_w;
    (*w->body)();
    Fiber &f = w->fibers[w->cur];
    f.done = true;
    zkemu_switch(&f.sp, w->sched_sp);
    abort();  // never resumed
}

void run_block(Worker *w, dim3 block, const std::function<void()> &body) {
    size_t n = (size_t) block.x * block.y * block.z;
    if (n > kMaxThreads) { fprintf(stderr, "zkemu: block too large\n"); abort(); }
    w->fibers.resize(n);
    w->body = &body;
    for (size_t i = 0; i < n; ++i) {
        Fiber &f = w->fibers[i];
        f.done = false;
        f.tid = dim3(i % block.x, (i / block.x) % block.y, i / ((size_t) block.x * block.y));
        // initial frame: 6 zeroed callee-saved registers, then the entry address as return address.
        uintptr_t top = (uintptr_t) (w->stack_base + (i + 1) * kStack);
        top &= ~(uintptr_t) 15;
        void **sp = (void **) top;
        *--sp = nullptr;                  // fake return address of fiber_entry (keeps 16-byte alignment at entry)
        *--sp = (void *) &fiber_entry;    // `ret` target
        for (int r = 0; r < 6; ++r) *--sp = nullptr;
        f.sp = sp;
    }
    size_t alive = n;
    while (alive) {
        for (size_t i = 0; i < n; ++i) {
            Fiber &f = w->fibers[i];
            if (f.done) continue;
            w->cur = (int) i;
            t_threadIdx = f.tid;
            zkemu_switch(&w->sched_sp, f.sp);
            if (f.done) --alive;
        }
    }
}
}  // namespace

void sync_threads() {
    Worker *w = t_w;
    Fiber &f = w->fibers[w->cur];
    zkemu_switch(&f.sp, w->sched_sp);
}

void *dyn_smem() { return t_w->smem.data(); }
uint64_t *exchange_buf() { return t_w->xbuf.data(); }

namespace {
// persistent worker pool: thread_local "shared memory" and fiber stacks live as long as the process
struct Pool {
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> threads;
    uint64_t generation = 0;
    unsigned active = 0;
    bool quit = false;
    // current job
    dim3 grid, block;
    const std::function<void()> *body = nullptr;
    std::atomic<size_t> next{0};
    size_t nblocks = 0;

    void worker_main() {
        Worker worker;
        t_w = &worker;
        worker.stack_base = (uint8_t *) mmap(nullptr, kMaxThreads * kStack, PROT_READ | PROT_WRITE,
                                              MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (worker.stack_base == MAP_FAILED) { perror("zkemu mmap"); abort(); }
        worker.smem.resize(kDynSmem);
        worker.xbuf.resize(kMaxThreads);
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv_work.wait(lk, [&] { return quit || generation != seen; });
                if (quit) return;
                seen = generation;
            }
            t_blockDim = block;
            t_gridDim = grid;
            for (;;) {
                size_t b = next.fetch_add(1);
                if (b >= nblocks) break;
                t_blockIdx = dim3(b % grid.x, (b / grid.x) % grid.y, b / ((size_t) grid.x * grid.y));
                run_block(&worker, block, *body);
            }
            {
                std::unique_lock<std::mutex> lk(m);
                if (--active == 0) cv_done.notify_all();
            }
        }
    }
    Pool() {
        unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        if (const char *e = getenv("ZKEMU_THREADS")) hw = std::max(1, atoi(e));
        for (unsigned i = 0; i < hw; ++i) threads.emplace_back([this] { worker_main(); });
    }
    ~Pool() {
        { std::unique_lock<std::mutex> lk(m); quit = true; }
        cv_work.notify_all();
        for (auto &t : threads) t.join();
    }
};
Pool &pool() { static Pool p; return p; }
}  // namespace

void launch(dim3 grid_, dim3 block_, size_t smem, const std::function<void()> &body_) {
    if (smem > kDynSmem) { fprintf(stderr, "zkemu: dynamic smem too large\n"); abort(); }
    size_t nblocks = (size_t) grid_.x * grid_.y * grid_.z;
    if (nblocks == 0) return;
    Pool &p = pool();
    std::unique_lock<std::mutex> lk(p.m);
    p.grid = grid_; p.block = block_; p.body = &body_; p.nblocks = nblocks;
    p.next.store(0);
    p.active = (unsigned) p.threads.size();
    ++p.generation;
    p.cv_work.notify_all();
    p.cv_done.wait(lk, [&] { return p.active == 0; });
}
}  // namespace zkemu
