// cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
//
// Execution model: a pool of OS threads; each takes one *cluster* of blocks at a time and
// runs all its CUDA threads as cooperative fibres (ucontext) in round-robin order.
// __syncthreads / __syncwarp / cluster barriers are generation counters on which fibres
// yield.  Warp shuffles go through a per-warp slot array.
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
constexpr int kMaxCluster = 16;

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

struct BlockCtx {
    std::vector<Fibre> fib;
    std::vector<WarpState> warps;
    unsigned char* smem = nullptr;
    int bar_arrive = 0;
    unsigned bar_gen = 0;
    int nb_arrive[16] = {0};
    unsigned nb_gen[16] = {0};
    uint3 bidx;
};

struct Worker {
    BlockCtx blk[kMaxCluster];
    ucontext_t sched;
    int cur_blk = 0, cur = -1;
    int nthreads = 0, cs = 1;
    int cl_arrive = 0;
    unsigned cl_gen = 0;
    std::vector<unsigned> cl_seen;     // per fibre: generation at which it arrived
    const std::function<void()>* fn = nullptr;
};

thread_local Worker* t_worker = nullptr;

void yield_to_sched() {
    Worker* w = t_worker;
    swapcontext(&w->blk[w->cur_blk].fib[w->cur].ctx, &w->sched);
}

void fibre_entry() {
    Worker* w = t_worker;
    (*w->fn)();
    Fibre& f = w->blk[w->cur_blk].fib[w->cur];
    f.done = true;
    swapcontext(&f.ctx, &w->sched);
}

void run_cluster(Worker* w, unsigned cl, dim3 grid, dim3 block, unsigned cs) {
    const int n = (int)(block.x * block.y * block.z);
    w->nthreads = n;
    w->cs = (int)cs;
    w->cl_arrive = 0;
    w->cl_seen.assign((size_t)cs * n, 0u);
    const int nwarp = (n + 31) / 32;
    for (unsigned b = 0; b < cs; ++b) {
        BlockCtx& B = w->blk[b];
        if (!B.smem) B.smem = (unsigned char*)aligned_alloc(128, kSmemMax);
        if ((int)B.fib.size() < n) B.fib.resize(n);
        if ((int)B.warps.size() < nwarp) B.warps.resize(nwarp);
        B.bar_arrive = 0;
        for (int i = 0; i < 16; ++i) B.nb_arrive[i] = 0;
        for (int i = 0; i < nwarp; ++i) {
            B.warps[i].arrive = 0;
            B.warps[i].lanes = (i == nwarp - 1) ? n - 32 * i : 32;
        }
        const unsigned blk = cl * cs + b;
        B.bidx.x = blk % grid.x;
        B.bidx.y = (blk / grid.x) % grid.y;
        B.bidx.z = blk / (grid.x * grid.y);
        for (int t = 0; t < n; ++t) {
            Fibre& f = B.fib[t];
            if (!f.stack) f.stack = (unsigned char*)aligned_alloc(64, kStack);
            f.done = false;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = &w->sched;
            makecontext(&f.ctx, fibre_entry, 0);
        }
    }
    int alive = n * (int)cs;
    while (alive > 0) {
        for (unsigned b = 0; b < cs; ++b) {
            BlockCtx& B = w->blk[b];
            for (int t = 0; t < n; ++t) {
                Fibre& f = B.fib[t];
                if (f.done) continue;
                w->cur_blk = (int)b;
                w->cur = t;
                t_blockIdx = B.bidx;
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
}

struct Pool {
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::vector<std::thread> threads;
    unsigned long job_id = 0;
    std::function<void()> fn;
    dim3 grid, block;
    unsigned cs = 1;
    std::atomic<unsigned> next{0};
    unsigned nclusters = 0;
    int active = 0;
    bool quit = false;

    Pool() {
        int n = (int)std::thread::hardware_concurrency();
        if (const char* e = getenv("SPCSC_EMU_THREADS")) n = atoi(e);
        if (n < 1) n = 1;
        for (int i = 0; i < n; ++i) threads.emplace_back([this]() { loop(new Worker()); });
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
                unsigned c = next.fetch_add(1);
                if (c >= nclusters) break;
                run_cluster(w, c, grid, block, cs);
            }
            {
                std::lock_guard<std::mutex> l(mu);
                if (--active == 0) cv_done.notify_all();
            }
        }
    }
    void run(dim3 g, dim3 b, unsigned c, std::function<void()> f) {
        std::unique_lock<std::mutex> l(mu);
        fn = std::move(f);
        grid = g;
        block = b;
        cs = c;
        nclusters = g.x * g.y * g.z / c;
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

void warp_barrier(WarpState& ws) {
    unsigned g = ws.gen;
    if (++ws.arrive == ws.lanes) {
        ws.arrive = 0;
        ++ws.gen;
        return;
    }
    while (ws.gen == g) yield_to_sched();
}

}  // namespace

unsigned char* dyn_smem() { return t_worker->blk[t_worker->cur_blk].smem; }

void launch_cluster(dim3 grid, dim3 block, unsigned cs, size_t smem, std::function<void()> fn) {
    if (smem > kSmemMax || cs < 1 || cs > (unsigned)kMaxCluster || grid.x % cs != 0) {
        fprintf(stderr, "emu: bad launch (smem %zu, cluster %u, grid.x %u)\n", smem, cs, grid.x);
        abort();
    }
    if (grid.x * grid.y * grid.z == 0) return;
    pool().run(grid, block, cs, std::move(fn));
}

void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> fn) {
    launch_cluster(grid, block, 1, smem, std::move(fn));
}

void sync_threads() {
    Worker* w = t_worker;
    BlockCtx& B = w->blk[w->cur_blk];
    unsigned g = B.bar_gen;
    if (++B.bar_arrive == w->nthreads) {
        B.bar_arrive = 0;
        ++B.bar_gen;
        return;
    }
    while (B.bar_gen == g) yield_to_sched();
}

void named_barrier(int id, int nthreads) {
    Worker* w = t_worker;
    BlockCtx& B = w->blk[w->cur_blk];
    unsigned g = B.nb_gen[id];
    if (++B.nb_arrive[id] == nthreads) {
        B.nb_arrive[id] = 0;
        ++B.nb_gen[id];
        return;
    }
    while (B.nb_gen[id] == g) yield_to_sched();
}

void sync_warp() {
    Worker* w = t_worker;
    warp_barrier(w->blk[w->cur_blk].warps[w->cur / 32]);
}

void warp_exchange(const void* mine, void* out, int src_lane, size_t nbytes) {
    Worker* w = t_worker;
    WarpState& ws = w->blk[w->cur_blk].warps[w->cur / 32];
    memcpy(ws.slot[w->cur % 32], mine, nbytes);
    warp_barrier(ws);
    if (src_lane >= ws.lanes) src_lane = w->cur % 32;
    memcpy(out, ws.slot[src_lane], nbytes);
    warp_barrier(ws);
}

void yield() { yield_to_sched(); }

unsigned cluster_rank() { return (unsigned)t_worker->cur_blk; }
unsigned cluster_size() { return (unsigned)t_worker->cs; }

// split cluster barrier: arrive records the generation, wait blocks until it has advanced
void cluster_arrive() {
    Worker* w = t_worker;
    w->cl_seen[(size_t)w->cur_blk * w->nthreads + w->cur] = w->cl_gen;
    if (++w->cl_arrive == w->nthreads * w->cs) {
        w->cl_arrive = 0;
        ++w->cl_gen;
    }
}
void cluster_wait() {
    Worker* w = t_worker;
    const unsigned g = w->cl_seen[(size_t)w->cur_blk * w->nthreads + w->cur];
    while (w->cl_gen == g) yield_to_sched();
}

void* map_shared_rank(void* p, unsigned rank) {
    Worker* w = t_worker;
    unsigned char* base = w->blk[w->cur_blk].smem;
    size_t off = (unsigned char*)p - base;
    if (off >= kSmemMax || rank >= (unsigned)w->cs) {
        fprintf(stderr, "emu: map_shared_rank out of range\n");
        abort();
    }
    return w->blk[rank].smem + off;
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
