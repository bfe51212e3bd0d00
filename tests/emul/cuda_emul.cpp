// tests/emul/cuda_emul.cpp -- TEST INFRASTRUCTURE ONLY: fiber scheduler behind cuda_emul.h.
#include "cuda_emul.h"
#include <dlfcn.h>
#include <map>

uint3e threadIdx, blockIdx;
dim3 blockDim, gridDim;
unsigned char* emul_dyn_smem = nullptr;

namespace emul {

static const size_t STACK = 256 * 1024;
struct Fiber { ucontext_t ctx; char* stack = nullptr; bool done = false; void* site = nullptr; int kind = 0; void* hist[8] = {0}; unsigned nbar = 0; };
static std::vector<Fiber> fibers;
static ucontext_t sched_ctx;
static int cur = -1;
static std::vector<void*> full_hist[2];   // complete barrier history of tid 0 and tid 32 (diagnostics)
static const std::function<void()>* cur_body = nullptr;

// block barrier state
static unsigned bar_gen = 0; static int bar_arrived = 0;
// named barriers (bar.sync id, nthreads): released when `want` fibers have arrived
static unsigned nb_gen[16]; static int nb_arrived[16], nb_want[16];
// warp exchange state
struct WarpEx { unsigned gen = 0; unsigned arrived = 0; unsigned want = 0xffffffffu; unsigned long long buf[2][32]; unsigned arrived_mask_done[2] = {0, 0}; };
static std::vector<WarpEx> warps;

static void yield() { int me = cur; swapcontext(&fibers[me].ctx, &sched_ctx); }

static void trampoline() { (*cur_body)(); fibers[cur].done = true; swapcontext(&fibers[cur].ctx, &sched_ctx); }

void block_barrier()
{
    fibers[cur].site = __builtin_return_address(0); fibers[cur].kind = 1; fibers[cur].hist[fibers[cur].nbar++ & 7] = fibers[cur].site;
    if (cur == 0) full_hist[0].push_back(fibers[cur].site); else if (cur == 32) full_hist[1].push_back(fibers[cur].site);
    unsigned g = bar_gen; bar_arrived++;
    while (bar_gen == g) yield();
}

void named_barrier(int id, int count)
{
    fibers[cur].site = __builtin_return_address(0); fibers[cur].kind = 3;
    unsigned g = nb_gen[id]; nb_arrived[id]++; nb_want[id] = count;
    while (nb_gen[id] == g) yield();
}

unsigned warp_exchange(unsigned mask, unsigned long long val, unsigned long long* all32)
{
    int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    WarpEx& W = warps[w];
    fibers[cur].site = __builtin_return_address(0); fibers[cur].kind = 2;
    unsigned g = W.gen; int slot = g & 1;
    W.buf[slot][lane] = val; W.arrived |= 1u << lane; W.want = mask;
    while (W.gen == g) yield();
    memcpy(all32, W.buf[slot], sizeof(W.buf[slot]));
    return W.arrived_mask_done[slot];
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body)
{
    int nt = (int)block.x;
    blockDim = block; gridDim = grid;
    std::vector<unsigned char> dyn(smem + 64);
    emul_dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 15) & ~(uintptr_t)15);
    if ((int)fibers.size() < nt) { size_t old = fibers.size(); fibers.resize(nt); for (size_t i = old; i < fibers.size(); ++i) fibers[i].stack = (char*)malloc(STACK); }
    cur_body = &body;
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        bar_gen = 0; bar_arrived = 0; full_hist[0].clear(); full_hist[1].clear();
        for (int i = 0; i < 16; ++i) { nb_gen[i] = 0; nb_arrived[i] = 0; nb_want[i] = 0; }
        warps.assign((nt + 31) / 32, WarpEx());
        for (int i = 0; i < nt; ++i) {
            Fiber& f = fibers[i]; f.done = false; f.nbar = 0;
            getcontext(&f.ctx); f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK; f.ctx.uc_link = &sched_ctx;
            makecontext(&f.ctx, trampoline, 0);
        }
        int alive = nt;
        while (alive > 0) {
            int progressed = 0;
            for (int i = 0; i < nt; ++i) {
                if (fibers[i].done) continue;
                cur = i; threadIdx.x = i; threadIdx.y = threadIdx.z = 0; blockIdx.x = bx; blockIdx.y = by; blockIdx.z = 0;
                swapcontext(&sched_ctx, &fibers[i].ctx);
                if (fibers[i].done) { alive--; progressed = 1; }
            }
            // release block barrier when every live fiber has arrived
            if (alive > 0 && bar_arrived >= alive) { bar_arrived = 0; bar_gen++; progressed = 1; }
            for (int i = 0; i < 16; ++i) if (nb_arrived[i] && nb_arrived[i] >= nb_want[i]) { nb_arrived[i] = 0; nb_gen[i]++; progressed = 1; }
            // release warp exchanges when every live lane of the warp has arrived
            for (size_t w = 0; w < warps.size(); ++w) {
                WarpEx& W = warps[w]; if (!W.arrived) continue;
                unsigned live = 0; for (int l = 0; l < 32; ++l) { int t = (int)w * 32 + l; if (t < nt && !fibers[t].done) live |= 1u << l; }
                if ((W.arrived & W.want & live) == (W.want & live)) { W.arrived_mask_done[W.gen & 1] = W.arrived; W.arrived = 0; W.gen++; progressed = 1; }
            }
            if (!progressed && alive > 0) {
                // every live fiber is parked: either all at a barrier (handled above) or a deadlock
                bool any = false;
                for (size_t w = 0; w < warps.size(); ++w) if (warps[w].arrived) any = true;
                for (int i = 0; i < 16; ++i) if (nb_arrived[i]) { fprintf(stderr, "cuda_emul: named barrier %d stuck: %d/%d arrived\n", i, nb_arrived[i], nb_want[i]); abort(); }
                if (!any && bar_arrived < alive) { fprintf(stderr, "cuda_emul: deadlock (block %u): %d/%d at barrier\n", bx, bar_arrived, alive); abort(); }
                if (any && bar_arrived > 0) {
                    fprintf(stderr, "cuda_emul: divergent barrier (block %u)\n", bx);
                    std::map<std::pair<void*, int>, std::vector<int>> sites;
                    for (int i = 0; i < nt; ++i) if (!fibers[i].done) sites[{fibers[i].site, fibers[i].kind}].push_back(i);
                    for (auto& kv : sites) {
                        Dl_info di; dladdr(kv.first.first, &di);
                        fprintf(stderr, "  %s at %s+0x%lx : %zu fibers, first tid %d last tid %d\n", kv.first.second == 1 ? "block barrier" : "warp sync",
                                di.dli_fname ? di.dli_fname : "?", (unsigned long)((char*)kv.first.first - (char*)di.dli_fbase), kv.second.size(), kv.second.front(), kv.second.back());
                    }
                    for (int i : {0, 1, 32, 63, 511}) if (i < nt) { Dl_info di; dladdr(fibers[i].site, &di); fprintf(stderr, "  tid %d nbar %u hist:", i, fibers[i].nbar);
                        for (unsigned h = 0; h < 8; ++h) fprintf(stderr, " 0x%lx", (unsigned long)((char*)fibers[i].hist[(fibers[i].nbar + h) & 7] - (char*)di.dli_fbase)); fprintf(stderr, "\n"); }
                    if (getenv("EMUL_DUMP_HIST")) { Dl_info di; dladdr(fibers[0].site, &di);
                        for (int k = 0; k < 2; ++k) { FILE* f = fopen(k ? "/tmp/emul_hist_32.txt" : "/tmp/emul_hist_0.txt", "w");
                            for (void* a : full_hist[k]) fprintf(f, "0x%lx\n", (unsigned long)((char*)a - (char*)di.dli_fbase)); fclose(f); } }
                    abort();
                }
            }
        }
    }
    emul_dyn_smem = nullptr;
}

}  // namespace emul
