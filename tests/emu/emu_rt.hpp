// emu_rt.hpp - a small SPMD emulator for the CUDA kernels of this repository (test infrastructure, CPU only).
//
//   * one OS thread per CTA (so CTAs really run concurrently and grid-wide spin barriers work), one ucontext fiber per CUDA
//     thread of the CTA; `__shared__` variables are thread_local, i.e. per CTA;
//   * __syncthreads / warp shuffles are rendezvous points between fibers;
//   * asynchronous copies are DEFERRED as long as the programming model allows: an LDGSTS copy (cp.async) is performed when the
//     issuing thread executes its wait, a bulk copy when some thread polls the mbarrier it completes on - a kernel that touches
//     a stage before it has waited reads the poison the dynamic shared memory is filled with;
//   * dynamic shared memory and every array handed to a kernel through emu_alloc end at an inaccessible guard page, so reads or
//     writes past their end fault instead of passing silently.
#pragma once
#include <sys/mman.h>
#include <ucontext.h>
#include <unistd.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace phw_emu {

// PHW_EMU_MODE: "late" (default) - asynchronous copies land as late as the programming model allows (a stage that is read before
// it was waited for is still poison); "early" - they land at issue time (a stage that is refilled while somebody still needs its
// old content is clobbered).  PHW_EMU_SEED != 0: the fibers of a CTA are scheduled in random order between rendezvous points.
inline bool mode_early() { static const bool e = [] { const char* m = std::getenv("PHW_EMU_MODE"); return m && std::string(m) == "early"; }(); return e; }
inline unsigned long long sched_seed() { static const unsigned long long v = [] { const char* m = std::getenv("PHW_EMU_SEED"); return m ? std::strtoull(m, nullptr, 10) : 0ull; }(); return v; }

struct dim3e { unsigned int x = 1, y = 1, z = 1; };
inline thread_local dim3e threadIdx, blockIdx, blockDim, gridDim;

struct Copy { void* dst; const void* src; size_t bytes; const void* bar; };
struct MBar { unsigned phases = 0; int pending = 0, count = 0; long long tx = 0; long long idle_polls = 0; };

struct Cta {
    int nt = 0, cur = 0, alive = 0;
    std::vector<ucontext_t> ctx;
    ucontext_t sched;
    std::vector<char*> stacks;
    std::vector<char> done;
    int bar_arrived = 0; unsigned bar_gen = 0;
    struct Warp { double val[32]; int arrived = 0; unsigned gen = 0; };
    std::vector<Warp> warps;
    std::vector<std::vector<Copy>> ldgsts;      // per thread
    std::vector<Copy> bulk;                     // per CTA
    std::unordered_map<const void*, MBar> bars;
    unsigned char* dyn = nullptr;
    const std::function<void()>* body = nullptr;
    long long ticks = 0, bulk_issued = 0;
};
inline thread_local Cta* g_cta = nullptr;

inline void yield() {
    Cta* c = g_cta;
    const int me = c->cur;
    swapcontext(&c->ctx[me], &c->sched);
}
// fault injection (tests of the failure path): PHW_EMU_DROP=n drops the n-th bulk copy of CTA 0 (its bytes never arrive);
// the emulated clock then advances fast enough for the kernels' wait limit to expire
inline long long drop_nth() { static const long long v = [] { const char* m = std::getenv("PHW_EMU_DROP"); return m ? std::atoll(m) : 0ll; }(); return v; }
inline long long clock() { g_cta->ticks += drop_nth() ? 50000000LL : 1LL; return g_cta->ticks; }

inline void syncthreads() {
    Cta* c = g_cta;
    const unsigned gen = c->bar_gen;
    if (++c->bar_arrived == c->alive) { c->bar_arrived = 0; c->bar_gen++; return; }
    while (c->bar_gen == gen) yield();
}
inline double shfl_xor(double v, int lane_mask) {
    Cta* c = g_cta;
    const int tid = (int)threadIdx.x, w = tid >> 5, l = tid & 31;
    const int lanes = std::min(32, c->nt - 32 * w);
    Cta::Warp& W = c->warps[w];
    W.val[l] = v;
    unsigned gen = W.gen;
    if (++W.arrived == lanes) { W.arrived = 0; W.gen++; } else while (W.gen == gen) yield();
    const double r = W.val[(l ^ lane_mask) < lanes ? (l ^ lane_mask) : l];
    gen = W.gen;
    if (++W.arrived == lanes) { W.arrived = 0; W.gen++; } else while (W.gen == gen) yield();
    return r;
}

// ---- asynchronous copies
inline void ldgsts(void* dst, const void* src, size_t bytes) {
    if (mode_early()) { std::memcpy(dst, src, bytes); return; }
    g_cta->ldgsts[threadIdx.x].push_back(Copy{dst, src, bytes, nullptr});
}
inline void ldgsts_wait_all() {
    auto& q = g_cta->ldgsts[threadIdx.x];
    for (const Copy& k : q) std::memcpy(k.dst, k.src, k.bytes);
    q.clear();
}
inline void mbar_init(const void* bar, unsigned count) { MBar& b = g_cta->bars[bar]; b = MBar{}; b.count = (int)count; b.pending = (int)count; }
inline void mbar_arrive_expect_tx(const void* bar, unsigned bytes) { MBar& b = g_cta->bars.at(bar); b.tx += bytes; b.pending -= 1; }
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, const void* bar) {
    if ((((size_t)dst) & 15) || (((size_t)src) & 15) || (bytes & 15) || bytes == 0) {
        std::fprintf(stderr, "emu: bulk copy violates the 16-byte rule (dst %p src %p bytes %u)\n", dst, src, bytes);
        std::abort();
    }
    if (drop_nth() && blockIdx.x == 0 && ++g_cta->bulk_issued == drop_nth()) return;      // injected fault: this copy is lost
    if (mode_early()) { std::memcpy(dst, src, bytes); g_cta->bars.at(bar).tx -= (long long)bytes; return; }
    g_cta->bulk.push_back(Copy{dst, src, bytes, bar});
}
inline bool mbar_try_wait(const void* bar, unsigned parity) {
    Cta* c = g_cta;
    MBar& b = c->bars.at(bar);
    for (size_t i = 0; i < c->bulk.size();) {
        if (c->bulk[i].bar == bar) {
            std::memcpy(c->bulk[i].dst, c->bulk[i].src, c->bulk[i].bytes);
            b.tx -= (long long)c->bulk[i].bytes;
            c->bulk[i] = c->bulk.back(); c->bulk.pop_back();
        } else ++i;
    }
    if (b.tx < 0) { std::fprintf(stderr, "emu: more bytes arrived on an mbarrier than were expected\n"); std::abort(); }
    if (b.pending == 0 && b.tx == 0) { b.phases++; b.pending = b.count; }
    const bool ok = (b.phases & 1u) != parity;
    if (ok) b.idle_polls = 0;
    else {
        if (++b.idle_polls > 20000000LL && !drop_nth()) {
            std::fprintf(stderr, "emu: an mbarrier phase never completes (pending arrivals %d, outstanding bytes %lld)\n", b.pending, b.tx);
            std::abort();
        }
        yield();
    }
    return ok;
}
inline unsigned char* dyn_smem() { return g_cta->dyn; }

// ---- guarded allocations: the usable range ends at a PROT_NONE page
struct Guarded { void* base = nullptr; size_t map_bytes = 0; void* ptr = nullptr; };
inline Guarded guarded_alloc(size_t bytes) {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t use = (bytes + 15) & ~(size_t)15;
    const size_t body = (use + page - 1) / page * page;
    Guarded g;
    g.map_bytes = body + page;
    g.base = mmap(nullptr, g.map_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (g.base == MAP_FAILED) { std::perror("mmap"); std::abort(); }
    mprotect((char*)g.base + body, page, PROT_NONE);
    g.ptr = (char*)g.base + body - use;
    return g;
}
inline void guarded_free(Guarded& g) { if (g.base) munmap(g.base, g.map_bytes); g = Guarded{}; }

// ---- launch: `grid` CTAs of `nt` threads, `smem` bytes of dynamic shared memory, body() is the kernel call
inline void fiber_entry() {
    Cta* c = g_cta;
    (*c->body)();
    c->done[c->cur] = 1;
    c->alive--;
    // a thread that exits no longer takes part in barriers: release a barrier the others are waiting at
    if (c->alive > 0 && c->bar_arrived == c->alive) { c->bar_arrived = 0; c->bar_gen++; }
    swapcontext(&c->ctx[c->cur], &c->sched);
}
inline void run_cta(int bid, int grid, int nt, size_t smem, const std::function<void()>& body) {
    Cta cta;
    g_cta = &cta;
    cta.nt = nt; cta.alive = nt; cta.body = &body;
    cta.ctx.resize(nt); cta.done.assign(nt, 0); cta.stacks.resize(nt);
    cta.warps.resize((nt + 31) / 32); cta.ldgsts.resize(nt);
    Guarded dyn = guarded_alloc(std::max<size_t>(smem, 16));
    std::memset(dyn.ptr, 0xFF, (smem + 15) & ~(size_t)15);          // poison: NaNs as doubles, huge ids as integers
    cta.dyn = (unsigned char*)dyn.ptr;
    blockIdx.x = (unsigned)bid; gridDim.x = (unsigned)grid; blockDim.x = (unsigned)nt;
    const size_t stack = 256 * 1024;
    for (int t = 0; t < nt; ++t) {
        cta.stacks[t] = (char*)std::malloc(stack);
        getcontext(&cta.ctx[t]);
        cta.ctx[t].uc_stack.ss_sp = cta.stacks[t];
        cta.ctx[t].uc_stack.ss_size = stack;
        cta.ctx[t].uc_link = &cta.sched;
        makecontext(&cta.ctx[t], (void (*)())fiber_entry, 0);
    }
    unsigned long long rng = sched_seed() ? sched_seed() * 0x9E3779B97F4A7C15ull + (unsigned long long)bid * 0xD1B54A32D192ED03ull + 1ull : 0ull;
    while (cta.alive > 0) {
        if (rng == 0) {
            for (int t = 0; t < nt; ++t) {
                if (cta.done[t]) continue;
                cta.cur = t; threadIdx.x = (unsigned)t;
                swapcontext(&cta.sched, &cta.ctx[t]);
            }
        } else {   // random runnable fiber (xorshift)
            rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
            int t = (int)(rng % (unsigned long long)nt);
            while (cta.done[t]) t = (t + 1) % nt;
            cta.cur = t; threadIdx.x = (unsigned)t;
            swapcontext(&cta.sched, &cta.ctx[t]);
        }
    }
    if (!cta.bulk.empty() && !drop_nth()) { std::fprintf(stderr, "emu: CTA %d exits with %zu bulk copies in flight\n", bid, cta.bulk.size()); std::abort(); }
    for (int t = 0; t < nt; ++t) {
        if (!cta.ldgsts[t].empty() && !drop_nth()) { std::fprintf(stderr, "emu: CTA %d thread %d exits with LDGSTS copies in flight\n", bid, t); std::abort(); }
        std::free(cta.stacks[t]);
    }
    guarded_free(dyn);
    g_cta = nullptr;
}
inline void launch(int grid, int nt, size_t smem, const std::function<void()>& body) {
    std::vector<std::thread> th;
    for (int b = 0; b < grid; ++b) th.emplace_back(run_cta, b, grid, nt, smem, std::cref(body));
    for (auto& t : th) t.join();
}

}  // namespace phw_emu
