// cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
#include "cuda_emu.h"

#include <ucontext.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

struct emuEvent {
    std::chrono::steady_clock::time_point t;
};

namespace emu {

thread_local uint3 t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;

namespace {

constexpr size_t kStack = 256 * 1024;
constexpr size_t kSmemMax = 256 * 1024;

struct WarpState {
    alignas(16) unsigned char slot[32][16];
    int arrive = 0;
    unsigned gen = 0;
    int lanes = 32;
};

struct Fibre {
    ucontext_t ctx;
    unsigned char* stack = nullptr;
    bool done = false;
};

struct Worker {
    std::vector<Fibre> fib;
    std::vector<WarpState> warps;
    ucontext_t sched;
    unsigned char* smem = nullptr;
    int cur = -1;
    int nthreads = 0;
    int bar_arrive = 0;
    unsigned bar_gen = 0;
    const std::function<void()>* fn = nullptr;
};

thread_local Worker* t_worker = nullptr;

void yield_to_sched() {
    Worker* w = t_worker;
    swapcontext(&w->fib[w->cur].ctx, &w->sched);
}

void fibre_entry() {
    Worker* w = t_worker;
    (*w->fn)();
    w->fib[w->cur].done = true;
    swapcontext(&w->fib[w->cur].ctx, &w->sched);
}

void run_block(Worker* w, unsigned blk, dim3 grid, dim3 block) {
    const int n = (int)(block.x * block.y * block.z);
    if ((int)w->fib.size() < n) w->fib.resize(n);
    w->nthreads = n;
    w->bar_arrive = 0;
    const int nwarp = (n + 31) / 32;
    if ((int)w->warps.size() < nwarp) w->warps.resize(nwarp);
    for (int i = 0; i < nwarp; ++i) {
        w->warps[i].arrive = 0;
        w->warps[i].lanes = (i == nwarp - 1) ? n - 32 * i : 32;
    }
    uint3 bidx;
    bidx.x = blk % grid.x;
    bidx.y = (blk / grid.x) % grid.y;
    bidx.z = blk / (grid.x * grid.y);
    for (int t = 0; t < n; ++t) {
        Fibre& f = w->fib[t];
        if (!f.stack) f.stack = (unsigned char*)aligned_alloc(64, kStack);
        f.done = false;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &w->sched;
        makecontext(&f.ctx, fibre_entry, 0);
    }
    int alive = n;
    while (alive > 0) {
        for (int t = 0; t < n; ++t) {
            Fibre& f = w->fib[t];
            if (f.done) continue;
            w->cur = t;
            t_blockIdx = bidx;
            t_blockDim = block;
            t_gridDim = grid;
            t_threadIdx.x = t % block.x;
            t_threadIdx.y = (t / block.x) % block.y;
            t_threadIdx.z = t / (block.x * block.y);
            swapcontext(&w->sched, &f.ctx);
            if (f.done) --alive;
        }
    }
}

struct Pool {
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::vector<std::thread> threads;
    std::vector<Worker*> workers;
    // current job
    unsigned long job_id = 0;
    std::function<void()> fn;
    dim3 grid, block;
    size_t smem = 0;
    std::atomic<unsigned> next{0};
    unsigned nblocks = 0;
    int active = 0;
    bool quit = false;

    Pool() {
        int n = (int)std::thread::hardware_concurrency();
        if (const char* e = getenv("SPCSC_EMU_THREADS")) n = atoi(e);
        if (n < 1) n = 1;
        for (int i = 0; i < n; ++i) {
            Worker* w = new Worker();
            w->smem = (unsigned char*)aligned_alloc(128, kSmemMax);
            workers.push_back(w);
            threads.emplace_back([this, w]() { loop(w); });
        }
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> l(mu);
            quit = true;
        }
        cv_job.notify_all();
        for (auto& t : threads) t.join();
    }
    void loop(Worker* w) {
        t_worker = w;
        unsigned long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(mu);
                cv_job.wait(l, [&]() { return quit || job_id != seen; });
                if (quit) return;
                seen = job_id;
            }
            w->fn = &fn;
            for (;;) {
                unsigned b = next.fetch_add(1);
                if (b >= nblocks) break;
                run_block(w, b, grid, block);
            }
            {
                std::lock_guard<std::mutex> l(mu);
                if (--active == 0) cv_done.notify_all();
            }
        }
    }
    void run(dim3 g, dim3 b, size_t s, std::function<void()> f) {
        std::unique_lock<std::mutex> l(mu);
        fn = std::move(f);
        grid = g;
        block = b;
        smem = s;
        nblocks = g.x * g.y * g.z;
        next = 0;
        active = (int)threads.size();
        ++job_id;
        cv_job.notify_all();
        cv_done.wait(l, [&]() { return active == 0; });
    }
};

Pool& pool() {
    static Pool p;
    return p;
}

}  // namespace

unsigned char* dyn_smem() { return t_worker->smem; }

void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> fn) {
    if (smem > kSmemMax) {
        fprintf(stderr, "emu: dynamic smem %zu too large\n", smem);
        abort();
    }
    if (grid.x * grid.y * grid.z == 0) return;
    pool().run(grid, block, smem, std::move(fn));
}

void sync_threads() {
    Worker* w = t_worker;
    unsigned g = w->bar_gen;
    if (++w->bar_arrive == w->nthreads) {
        w->bar_arrive = 0;
        ++w->bar_gen;
        return;
    }
    while (w->bar_gen == g) yield_to_sched();
}

static void warp_barrier(Worker* w, WarpState& ws) {
    unsigned g = ws.gen;
    if (++ws.arrive == ws.lanes) {
        ws.arrive = 0;
        ++ws.gen;
        return;
    }
    while (ws.gen == g) yield_to_sched();
}

void sync_warp() {
    Worker* w = t_worker;
    warp_barrier(w, w->warps[w->cur / 32]);
}

void warp_exchange(const void* mine, void* out, int src_lane, size_t nbytes) {
    Worker* w = t_worker;
    WarpState& ws = w->warps[w->cur / 32];
    memcpy(ws.slot[w->cur % 32], mine, nbytes);
    warp_barrier(w, ws);
    if (src_lane >= ws.lanes) src_lane = w->cur % 32;
    memcpy(out, ws.slot[src_lane], nbytes);
    warp_barrier(w, ws);
}

}  // namespace emu

cudaError_t cudaMalloc(void** p, size_t n) {
    *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
