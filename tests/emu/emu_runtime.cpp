// Fiber scheduler of the host emulation (see cuda_runtime.h in this directory) -- test infrastructure.
#include "cuda_runtime.h"
#include "cuda.h"

emu_uint3 threadIdx, blockIdx, blockDim, gridDim;

cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q) {
    const bool ok = strcmp(name, "cuTensorMapEncodeTiled") == 0;
    *fn = ok ? (void*)&emu_cuTensorMapEncodeTiled : nullptr;
    *q = ok ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
    return cudaSuccess;
}

namespace acme {
alignas(1024) unsigned char smem_raw[256 * 1024];  // the kernels' `extern __shared__ smem_raw[]`
}

namespace acme_emu {
Cta g_cta;
unsigned char* g_smem = acme::smem_raw;

namespace {
struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
};
std::vector<Fiber> fibers;
ucontext_t sched_ctx;
int cur = -1;
const std::function<void()>* g_body = nullptr;

void trampoline() {
    (*g_body)();
    fibers[cur].done = true;
    g_cta.live--;
    g_cta.warp_live[cur >> 5]--;
    g_cta.live_mask[cur >> 5] &= ~(1u << (cur & 31));
    g_cta.progress++;
    // a finished lane no longer takes part in its warp's barriers: release a barrier it was the last one missing from
    const int w = cur >> 5;
    if (g_cta.warp_live[w] && g_cta.warp_cnt[w] >= g_cta.warp_live[w]) { g_cta.warp_cnt[w] = 0; g_cta.warp_gen[w]++; }
    if (g_cta.live && g_cta.cta_cnt >= (unsigned)g_cta.live) { g_cta.cta_cnt = 0; g_cta.cta_gen++; }
    swapcontext(&fibers[cur].ctx, &sched_ctx);
}
}  // namespace

void yield() { swapcontext(&fibers[cur].ctx, &sched_ctx); }

void warp_barrier(unsigned mask) {
    const int w = cur >> 5;
    const unsigned m = mask & g_cta.live_mask[w];
    if (mask == 0xffffffffu) {  // the whole (live) warp; sub-warp masks always use their own barrier below
        const unsigned gen = g_cta.warp_gen[w];
        if (++g_cta.warp_cnt[w] >= g_cta.warp_live[w]) {
            g_cta.warp_cnt[w] = 0;
            g_cta.warp_gen[w]++;
            g_cta.progress++;
            return;
        }
        while (g_cta.warp_gen[w] == gen) yield();
        return;
    }
    // a sub-warp group: its own barrier, keyed by the mask
    Cta::GroupBar* gb = nullptr;
    for (auto& g : g_cta.group[w])
        if (g.mask == mask || g.mask == 0) { gb = &g; break; }
    if (!gb) { fprintf(stderr, "emu: more than 4 distinct sub-warp masks in one warp\n"); abort(); }
    gb->mask = mask;
    const unsigned gen = gb->gen;
    if (++gb->cnt >= (unsigned)__builtin_popcount(m)) {
        gb->cnt = 0;
        gb->gen++;
        g_cta.progress++;
        return;
    }
    while (gb->gen == gen) yield();
}

void cta_barrier() {
    const unsigned gen = g_cta.cta_gen;
    if (++g_cta.cta_cnt >= (unsigned)g_cta.live) {
        g_cta.cta_cnt = 0;
        g_cta.cta_gen++;
        g_cta.progress++;
        return;
    }
    while (g_cta.cta_gen == gen) yield();
}

void run_cta(const std::function<void()>& body, unsigned block, unsigned bx, unsigned grid) {
    if (block > MAX_THREADS) { fprintf(stderr, "emu: block of %u threads\n", block); abort(); }
    g_body = &body;
    g_cta = Cta();
    g_cta.nthreads = g_cta.live = (int)block;
    for (unsigned w = 0; w < MAX_THREADS / 32; w++) {
        g_cta.warp_gen[w] = g_cta.warp_cnt[w] = 0;
        const int lo = (int)w * 32;
        g_cta.warp_live[w] = lo >= (int)block ? 0u : (unsigned)std::min(32, (int)block - lo);
        g_cta.live_mask[w] = g_cta.warp_live[w] >= 32 ? 0xffffffffu : ((1u << g_cta.warp_live[w]) - 1u);
        for (auto& g : g_cta.group[w]) g = Cta::GroupBar{0, 0, 0};
    }
    blockIdx = {bx, 0, 0}; blockDim = {block, 1, 1}; gridDim = {grid, 1, 1};
    fibers.assign(block, Fiber());
    for (unsigned t = 0; t < block; t++) {
        Fiber& f = fibers[t];
        f.stack.resize(512 * 1024);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.data();
        f.ctx.uc_stack.ss_size = f.stack.size();
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, trampoline, 0);
    }
    while (g_cta.live > 0) {
        const unsigned long before = g_cta.progress;
        for (unsigned t = 0; t < block; t++) {
            if (fibers[t].done) continue;
            cur = (int)t;
            threadIdx = {t, 0, 0};
            swapcontext(&sched_ctx, &fibers[t].ctx);
        }
        if (g_cta.progress == before && g_cta.live > 0) {
            fprintf(stderr, "emu: deadlock -- a collective was not reached by every live thread of its warp/CTA\n");
            fprintf(stderr, "  block %u, live %d, cta barrier %u/%d\n", bx, g_cta.live, g_cta.cta_cnt, g_cta.live);
            for (unsigned w = 0; w * 32 < block; w++) {
                fprintf(stderr, "  warp %u: live mask %08x, full-warp barrier %u/%u", w, g_cta.live_mask[w], g_cta.warp_cnt[w], g_cta.warp_live[w]);
                for (auto& g : g_cta.group[w])
                    if (g.mask) fprintf(stderr, ", group %08x %u/%d", g.mask, g.cnt, __builtin_popcount(g.mask & g_cta.live_mask[w]));
                fprintf(stderr, "\n");
            }
            abort();
        }
    }
    cur = -1;
}
}  // namespace acme_emu
