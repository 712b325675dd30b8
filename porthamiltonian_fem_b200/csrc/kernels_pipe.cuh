// kernels_pipe.cuh - bulk-copy pipelined variants of the two Chebyshev mass solves (sm_100a).
//
// The lean solve kernels (kernels.cuh: solve_ell_coop_k, solve_mq_coop_k) are bound by the number of independent request
// streams an SM keeps in flight, not by the layout (phw_stream_probe): every patch goes through
// load -> barrier -> compute -> barrier phases and only the loads of the resident CTAs overlap.  Here ALL contiguous
// streams of a patch (cell records or ELL rows, b, x_{k-1}, the owned range of x_k, Jacobi weights, adjacency codes)
// are fetched by the copy engine: one thread issues 1-D bulk copies (cp.async.bulk.shared::cluster.global, the TMA path
// without a tensor map) that complete on an mbarrier, into the OTHER of two shared-memory stages while the CTA computes
// the current patch entirely from shared memory.  The irregular part (ghost values, <= 10 % of a patch) is gathered by
// 8-byte LDGSTS copies one patch ahead, from an index list that itself arrived by bulk copy with the previous stage.
// Per SM one stage (50-100 KB) is always in flight, independent of registers and resident warps.
//
// Bulk copies need 16-byte aligned addresses and sizes.  The owned ranges start at arbitrary dof ids, so a stream of n
// elements starting at global element `start` lands in shared memory shifted by (start mod 2) [doubles] or (start mod 4)
// [32-bit words]: the aligned middle part travels as one bulk copy, the (<= 1 / <= 3) elements before and after it as
// single LDGSTS copies - nothing outside the range is ever read.
//
// Same iteration, same accepted iterate and the same deterministic reductions as the lean kernels (wave_2D.py:820,854).
#pragma once
#include "kernels.cuh"

// the dynamic shared memory of a pipelined kernel (the emulator hands out a guarded host buffer instead)
#ifdef PHW_HOST_EMU
#define PHW_DYN_SMEM(name) unsigned char* name = phw_emu::dyn_smem()
#else
#define PHW_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

namespace phw {

#ifdef PHW_HOST_EMU
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) { phw_emu::mbar_init(bar, count); }
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) { phw_emu::mbar_arrive_expect_tx(bar, bytes); }
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) { return phw_emu::mbar_try_wait(bar, parity); }
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
#endif
// wait for the phase with the given parity; false if the stage did not arrive.  A wait that exceeds ~5 s (never seen; a stage
// lands within microseconds) marks the run as failed in the high word of ctl->nonconv, which phw_run / phw_mass_solve report
// as an error, and every later wait of the run gives up at once.  The callers then skip the patch (pipe_alive), so that a
// protocol error ends in a clean error code - not in a hung GPU and not in loads from indices that never arrived.
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, uint32_t parity, Ctl* ctl) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    unsigned int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((spins++ & 63u) == 0u && (*((volatile long long*)&ctl->nonconv) >> 32) != 0) return false;
        if (clock64() - t0 > 10000000000LL) {
            atomicAdd((unsigned long long*)&ctl->nonconv, 1ull << 32);
            return false;
        }
    }
    return true;
}
// top-of-patch protocol of every pipelined kernel: own gathers done, bulk copies landed, visible to the whole CTA.
// Returns false (for all threads of the CTA alike) if the stage must not be used.
__device__ __forceinline__ bool pipe_stage_ready(unsigned long long* bar, uint32_t parity, Ctl* ctl, int* s_dead) {
    cp_async_wait_all();
    if (!mbar_wait(bar, parity, ctl)) *s_dead = 1;
    __syncthreads();
    return *s_dead == 0;
}
#ifdef PHW_HOST_EMU
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* bar) { phw_emu::bulk_g2s(smem_dst, gmem_src, bytes, bar); }
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) { phw_emu::ldgsts(smem_dst, gmem_src, 4); }
__device__ __forceinline__ void fence_proxy_async() {}
#else
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
// orders generic-proxy accesses (ld/st) against the async-proxy accesses of the bulk copies, both ways
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
#endif

// ---- streams of doubles: element i of the range [start, start + n) lives at s[(start & 1) + i]
__device__ __forceinline__ uint32_t f64s_bytes(int start, int n) {
    const int m = n - (start & 1);
    return m > 0 ? (uint32_t)(m & ~1) * 8u : 0u;
}
__device__ __forceinline__ void f64s_bulk(double* s, const double* g, int start, int n, unsigned long long* bar) {
    const int head = start & 1;
    const uint32_t bytes = f64s_bytes(start, n);
    if (bytes) bulk_g2s(s + 2 * head, g + start + head, bytes, bar);
}
// j = 0: the element before the aligned part, j = 1: the element after it (each by one thread)
__device__ __forceinline__ void f64s_fix(double* s, const double* g, int start, int n, int j) {
    const int head = start & 1, m = n - head;
    if (j == 0) { if (head && n > 0) cp_async8(&s[head], &g[start]); }
    else if (m > 0 && (m & 1)) cp_async8(&s[head + n - 1], &g[start + n - 1]);
}
// ---- streams of 32-bit words: element i lives at s[(start & 3) + i]
__device__ __forceinline__ int u32s_head(int start, int n) { const int h = (4 - (start & 3)) & 3; return h < n ? h : n; }
__device__ __forceinline__ uint32_t u32s_bytes(int start, int n) { return n > 0 ? (uint32_t)((n - u32s_head(start, n)) & ~3) * 4u : 0u; }
__device__ __forceinline__ void u32s_bulk(uint32_t* s, const uint32_t* g, int start, int n, unsigned long long* bar) {
    if (n <= 0) return;
    const int head = u32s_head(start, n);
    const uint32_t bytes = u32s_bytes(start, n);
    if (bytes) bulk_g2s(s + (start & 3) + head, g + start + head, bytes, bar);
}
// j = 0..2: the elements before the aligned part, j = 3..5: those after it
__device__ __forceinline__ void u32s_fix(uint32_t* s, const uint32_t* g, int start, int n, int j) {
    if (n <= 0) return;
    const int head = u32s_head(start, n), nb = (n - head) & ~3;
    const int e = (j < 3) ? j : head + nb + (j - 3);
    const bool ok = (j < 3) ? (j < head) : (e < n);
    if (ok) cp_async4(&s[(start & 3) + e], &g[start + e]);
}

// patch headers of a CTA's next patches are kept in a ring of four in shared memory (loaded three patches ahead)
__device__ __forceinline__ void load_hdr_smem_at(PatchHdr* dst, const PatchHdr* src, int first_thread) {
    const int t = (int)threadIdx.x - first_thread;
    if (t >= 0 && t < 20) reinterpret_cast<int*>(dst)[t] = __ldg(reinterpret_cast<const int*>(src) + t);
}

// partitioned runs: the send-list range of patch P (PeerComm::psend_off) travels with its header into a parallel ring
__device__ __forceinline__ void load_snd_smem_at(int* dst2, const int* off, int P, int first_thread) {
    const int t = (int)threadIdx.x - first_thread - 20;
    if (t >= 0 && t < 2) dst2[t] = __ldg(off + P + t);
}

// grid-wide barrier of the pipelined kernels: grid_barrier plus the proxy fences that order the iterate written with
// ordinary stores before the barrier against the bulk copies that read it after the barrier
__device__ __forceinline__ void grid_barrier_px(unsigned int* bar, unsigned int& epoch) {
    fence_proxy_async();
    grid_barrier(bar, epoch);
    fence_proxy_async();
}

// ---- inter-GPU exchange of the pipelined solves (partitioned runs; protocol of round 2) -----------------------------
// (1) peer_push_patch: right after the rows of patch P are written, the threads of the CTA copy its interface rows straight into the
//     neighbours' iterate buffers (peer-mapped memory, NVLink stores) - while the other CTAs (and this one) go on with the sweep, so
//     the transfer overlaps the interior work.  A CTA that pushed anything issues ONE system-scope fence (thread 0, after the
//     block-wide barriers of the residual reduction, before it arrives at the grid barrier): the fences are what costs on this
//     machine (measured, profiles/r2d_exchange_ab.txt: ~5 us when every pushing thread fences, ~4 us when thread 0 of every CTA
//     fences at system scope, ~1 us for a single thread) - and barrier + fence are cumulative, so one per CTA is enough.
// (2) peer_finish_iter: after the grid-wide reduction of the local residual norms, ONE thread publishes the rank's norms and a
//     sequence flag in every rank's mailbox (system-scope release: everything this grid pushed before its barrier is visible to
//     whoever sees the flag); thread 0 of EVERY CTA then polls the flags in its OWN mailbox (local memory, volatile loads), reads the
//     norms with volatile loads and sums them in rank order - identical bits on every rank and CTA, no second grid barrier, no
//     broadcast through global memory.  Consumer side: the mailbox and the iterate buffers the peers write are LOCAL memory, whose
//     point of coherence is this GPU's L2 - once the flag is visible there, so is everything the sender fenced before it; what is
//     left to do is to drop stale L1 lines before the CTA gathers external values, which a gpu-scope fence does (PHW_XCH_DEBUG=1
//     restores system-scope fences on both sides: the conservative reading of the memory model, ~8 us per iteration slower).
// Per iteration this adds about one NVLink one-way latency to the critical path (the round-1 protocol, peer_exchange_iter: push
// after the barrier + system fence by every thread + two more grid barriers).
template <int NT>
__device__ __forceinline__ bool peer_push_patch(const PeerComm& pc, int field, int b, int e, const double* xnew, int nxt) {
    if (b >= e) return false;                              // uniform for the CTA
    const int4* ent = pc.psend_ent[field];
    for (int i = b + (int)threadIdx.x; i < e; i += NT) {
        const int4 t = ent[i];
        double* dst = field == 0 ? pc.nb_tp[t.y][nxt] : pc.nb_tq[t.y][nxt];
        dst[t.z] = xnew[t.x];
    }
    if (pc.xch_dbg & 1) __threadfence_system();
    return true;
}
// seq: exchange number of this iteration (the same on every rank); s_x: two doubles of shared memory.  Called by all threads.
template <int NT>
__device__ __forceinline__ void peer_finish_iter(const PeerComm& pc, Ctl* ctl, unsigned long long seq, double* s_x, double& rr, double& bb) {
    const int par = (int)(seq & 1ull);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int r = 0; r < pc.world; ++r) { pc.mbox[r]->part[par][pc.rank][0] = rr; pc.mbox[r]->part[par][pc.rank][1] = bb; }
        __threadfence_system();
        for (int r = 0; r < pc.world; ++r) *((volatile unsigned long long*)&pc.mbox[r]->flag[pc.rank]) = seq;
    }
    if (threadIdx.x == 0) {
        Mailbox* mine = pc.mbox[pc.rank];
        bool dead = *((volatile int*)&ctl->comm_dead) != 0;
        if (!dead) {
            const long long t0 = clock64();
            for (int r = 0; r < pc.world && !dead; ++r)
                while (ld_volatile_u64(&mine->flag[r]) < seq)
                    if (clock64() - t0 > pc.timeout_cycles) { ctl->comm_dead = 1; dead = true; break; }
        }
        if (pc.xch_dbg & 1) __threadfence_system(); else __threadfence();
        double s0 = 0.0, s1 = 0.0;
        for (int r = 0; r < pc.world; ++r) { s0 += *((volatile double*)&mine->part[par][r][0]); s1 += *((volatile double*)&mine->part[par][r][1]); }
        s_x[0] = s0; s_x[1] = s1;
    }
    __syncthreads();
    rr = s_x[0]; bb = s_x[1];
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// M_q solve (cell-based rows on the packed 32-byte cell records).  Stage = [x_k local | cell records | b | pinv |
// x_{k-1} | eadj]; besides the two stages: the per-cell contributions (3 doubles per cell) and two ghost-index lists.
// ------------------------------------------------------------------------------------------------
struct MqPipeArgs {
    int npatch;
    const PatchHdr* hdr;
    const uint4* cellq;
    const int* ghost_e; const uint32_t* eadj;
    const double* b; const double* pinv;
    double* buf[3];
    double cheb_d, cheb_c, tol2;
    int max_iter, which;
    double* partials; Ctl* ctl;
    // carve-up of the dynamic shared memory (bytes, multiples of 16; sized for the largest patch by the host)
    uint32_t off_cells, off_b, off_pinv, off_xprev, off_eadj, stage_bytes, off_sc3, off_gidx, gidx_bytes;
};

// put everything patch `hn` needs in flight into `stage` (all threads call this; thread 0 issues the bulk copies).
// gidx_this: ghost index list of this patch in shared memory (nullptr: read it from global memory - first patch of an
// iteration); hn2 / gidx_next: the patch after `hn` and the list slot its ghost indices go to (nullptr: none).
template <int NT, bool DIAG>
__device__ __forceinline__ void mq_issue(const MqPipeArgs& a, unsigned char* stage, const uint32_t* gidx_this, uint32_t* gidx_next,
                                         const PatchHdr& hn, const PatchHdr* hn2, const double* xk, const double* xprev,
                                         bool use_prev, unsigned long long* bar) {
    double* sq = reinterpret_cast<double*>(stage);
    double* sb = reinterpret_cast<double*>(stage + a.off_b);
    double* spv = reinterpret_cast<double*>(stage + a.off_pinv);
    double* sxp = reinterpret_cast<double*>(stage + a.off_xprev);
    uint32_t* sea = reinterpret_cast<uint32_t*>(stage + a.off_eadj);
    const int es = hn.e_start, ec = hn.e_cnt;
    const uint32_t* gh = reinterpret_cast<const uint32_t*>(a.ghost_e);
    if (threadIdx.x == 0) {
        const uint32_t cell_bytes = (uint32_t)hn.n_ecells * 32u;
        uint32_t bytes = f64s_bytes(es, ec) * ((use_prev ? 3u : 2u) + (DIAG ? 0u : 1u)) + cell_bytes + u32s_bytes(es, ec);
        if (hn2) bytes += u32s_bytes(hn2->ge_off, hn2->ge_cnt);
        mbar_arrive_expect_tx(bar, bytes);
        if (cell_bytes) bulk_g2s(stage + a.off_cells, a.cellq + 2 * (size_t)hn.cell_off, cell_bytes, bar);
        f64s_bulk(sq, xk, es, ec, bar);
        f64s_bulk(sb, a.b, es, ec, bar);
        if (!DIAG) f64s_bulk(spv, a.pinv, es, ec, bar);
        if (use_prev) f64s_bulk(sxp, xprev, es, ec, bar);
        u32s_bulk(sea, a.eadj, es, ec, bar);
        if (hn2) u32s_bulk(gidx_next, gh, hn2->ge_off, hn2->ge_cnt, bar);
    }
    const int lane = (int)threadIdx.x - (NT - 32);      // the last warp copies the unaligned ends of the streams
    if (lane >= 0) {
        if (lane < 2) f64s_fix(sq, xk, es, ec, lane);
        else if (lane < 4) f64s_fix(sb, a.b, es, ec, lane - 2);
        else if (lane < 6) { if (!DIAG) f64s_fix(spv, a.pinv, es, ec, lane - 4); }
        else if (lane < 8) { if (use_prev) f64s_fix(sxp, xprev, es, ec, lane - 6); }
        else if (lane < 14) u32s_fix(sea, a.eadj, es, ec, lane - 8);
        else if (lane < 20) { if (hn2) u32s_fix(gidx_next, gh, hn2->ge_off, hn2->ge_cnt, lane - 14); }
    }
    double* sg = sq + (es & 1) + ec;                     // ghost values follow the owned range
    if (gidx_this) {
        const uint32_t* gi = gidx_this + (hn.ge_off & 3);
        for (int i = threadIdx.x; i < hn.ge_cnt; i += NT) cp_async8(&sg[i], &xk[gi[i]]);
    } else {
        for (int i = threadIdx.x; i < hn.ge_cnt; i += NT) cp_async8(&sg[i], &xk[__ldg(&a.ghost_e[hn.ge_off + i])]);
    }
}

// DIAG: the Jacobi weight 1 / diag(M_q)_e is recomputed from the (<= 2) cell records of the edge, which are in shared memory
// anyway (diag contribution of slot i of a cell: 3 S - 4 g_i), instead of being streamed (12 of 99 B per cell and iteration)
// MULTI: with the inter-GPU exchange (peer_push_patch / peer_finish_iter above: interface rows pushed into the neighbours' iterate
// buffers patch by patch during the sweep, residual norms all-reduced through the mailboxes).  External dofs are never part of an
// owned range, so only the ghost gathers (ordinary LDGSTS loads, behind the acquiring fence of peer_finish_iter) read what the
// peers wrote.
template <int NT, bool DIAG, bool MULTI>
__global__ void __launch_bounds__(NT, 1) solve_mq_pipe_k(const __grid_constant__ MqPipeArgs a, const __grid_constant__ PeerComm pc) {
    PHW_DYN_SMEM(smraw);
    __shared__ double red[64];
    __shared__ double bc[2];
    __shared__ PatchHdr s_hdr[4];
    __shared__ __align__(8) unsigned long long full[2];
    __shared__ int s_dead;       // a stage did not arrive (see mbar_wait): the CTA stops touching the stages
    __shared__ int s_snd[4][2];  // MULTI: send-list range of the patches in s_hdr
    SolveCtl& sc = a.ctl->sc[a.which];
    int k = 0, cur = sc.cur;
    double rho = 0.0, rr = 0.0, bb = 0.0;
    bool conv = false;
    unsigned int epoch = 0;
    unsigned int seq = 0;       // patches this CTA has processed so far: stage = seq & 1, mbarrier parity = (seq >> 1) & 1
    const unsigned long long xseq0 = MULTI ? a.ctl->comm_seq : 0ull;   // exchanges done before this launch (updated only at its end)
    if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); s_dead = 0; }
    __syncthreads();
    double* sc3 = reinterpret_cast<double*>(smraw + a.off_sc3);
    const int n_my = ((int)blockIdx.x < a.npatch) ? (a.npatch - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    while (true) {
        double c1, c2, rho_k;
        cheb_coeffs(k, rho, a.cheb_d, a.cheb_c, c1, c2, rho_k);
        const int nxt = (cur + 1) % 3, prv = (cur + 2) % 3;
        const double* xk = cur == 0 ? a.buf[0] : (cur == 1 ? a.buf[1] : a.buf[2]);
        const double* xprev = prv == 0 ? a.buf[0] : (prv == 1 ? a.buf[1] : a.buf[2]);
        double* xnew = nxt == 0 ? a.buf[0] : (nxt == 1 ? a.buf[1] : a.buf[2]);
        const bool use_prev = k > 0;
        double r0 = 0.0, b0 = 0.0;
        for (int j = 0; j < 3 && j < n_my; ++j) {
            load_hdr_smem_at(&s_hdr[j], &a.hdr[blockIdx.x + j * gridDim.x], 32 * j);
            if (MULTI) load_snd_smem_at(s_snd[j], pc.psend_off[1], blockIdx.x + j * gridDim.x, 32 * j);
        }
        int snd_b = 0, snd_e = 0;       // MULTI: send-list range of the patch whose rows were written last
        bool pushed = false;
        __syncthreads();
        if (n_my > 0)
            mq_issue<NT, DIAG>(a, smraw + (seq & 1u) * a.stage_bytes, nullptr,
                         reinterpret_cast<uint32_t*>(smraw + a.off_gidx + ((seq + 1u) & 1u) * a.gidx_bytes),
                         s_hdr[0], n_my > 1 ? &s_hdr[1] : nullptr, xk, xprev, use_prev, &full[seq & 1u]);
        for (int it = 0; it < n_my; ++it, ++seq) {
            const unsigned int s = seq & 1u;
            if (!pipe_stage_ready(&full[s], (seq >> 1) & 1u, a.ctl, &s_dead)) continue;
            const PatchHdr& h = s_hdr[it & 3];
            if (MULTI) {   // the rows of the previous patch are visible to the whole CTA now: push its interface rows to the neighbours
                if (!(pc.xch_dbg & 32)) pushed |= peer_push_patch<NT>(pc, 1, snd_b, snd_e, xnew, nxt);
                snd_b = s_snd[it & 3][0]; snd_e = s_snd[it & 3][1];
            }
            if (it + 3 < n_my) {
                load_hdr_smem_at(&s_hdr[(it + 3) & 3], &a.hdr[blockIdx.x + (it + 3) * gridDim.x], 64);
                if (MULTI) load_snd_smem_at(s_snd[(it + 3) & 3], pc.psend_off[1], blockIdx.x + (it + 3) * gridDim.x, 64);
            }
            if (it + 1 < n_my)
                mq_issue<NT, DIAG>(a, smraw + (s ^ 1u) * a.stage_bytes,
                             reinterpret_cast<const uint32_t*>(smraw + a.off_gidx + (s ^ 1u) * a.gidx_bytes),
                             reinterpret_cast<uint32_t*>(smraw + a.off_gidx + s * a.gidx_bytes),
                             s_hdr[(it + 1) & 3], it + 2 < n_my ? &s_hdr[(it + 2) & 3] : nullptr, xk, xprev, use_prev, &full[s ^ 1u]);
            const unsigned char* stage = smraw + s * a.stage_bytes;
            const double* sq = reinterpret_cast<const double*>(stage) + (h.e_start & 1);
            const uint4* scell = reinterpret_cast<const uint4*>(stage + a.off_cells);
            for (int lc = threadIdx.x; lc < h.n_ecells; lc += NT) {
                const uint4 qa = scell[2 * lc], qb = scell[2 * lc + 1];
                const uint32_t e0 = qa.x & 0xFFFFu, e1 = qa.x >> 16, e2 = qa.y & 0xFFFFu;
                const double g0 = __hiloint2double((int)qa.w, (int)qa.z);
                const double g1 = __hiloint2double((int)qb.y, (int)qb.x), g2 = __hiloint2double((int)qb.w, (int)qb.z);
                const double S = g0 + g1 + g2;
                const double s0 = (e0 & 0x8000u) ? -1.0 : 1.0, s1 = (e1 & 0x8000u) ? -1.0 : 1.0, s2 = (e2 & 0x8000u) ? -1.0 : 1.0;
                const double t0 = s0 * sq[e0 & 0x3FFFu], t1 = s1 * sq[e1 & 0x3FFFu], t2 = s2 * sq[e2 & 0x3FFFu];
                const double ST = S * (t0 + t1 + t2);
                sc3[3 * lc] = s0 * (ST + 2.0 * S * t0 - 4.0 * (g0 * t0 + g2 * t1 + g1 * t2));
                sc3[3 * lc + 1] = s1 * (ST + 2.0 * S * t1 - 4.0 * (g2 * t0 + g1 * t1 + g0 * t2));
                sc3[3 * lc + 2] = s2 * (ST + 2.0 * S * t2 - 4.0 * (g1 * t0 + g0 * t1 + g2 * t2));
            }
            __syncthreads();
            const double* sb = reinterpret_cast<const double*>(stage + a.off_b) + (h.e_start & 1);
            const double* spv = reinterpret_cast<const double*>(stage + a.off_pinv) + (h.e_start & 1);
            const double* sxp = reinterpret_cast<const double*>(stage + a.off_xprev) + (h.e_start & 1);
            const uint32_t* sea = reinterpret_cast<const uint32_t*>(stage + a.off_eadj) + (h.e_start & 3);
            const double* sgeo = reinterpret_cast<const double*>(stage + a.off_cells);
            for (int le = threadIdx.x; le < h.e_cnt; le += NT) {
                const uint32_t adj = sea[le];
                const uint32_t ca = adj & 0xFFFFu, cb = adj >> 16;
                const double bv = sb[le];
                double val = sc3[3 * (ca >> 2) + (ca & 3u)];
                if (cb != 0xFFFFu) val += sc3[3 * (cb >> 2) + (cb & 3u)];
                double pv;
                if (DIAG) {
                    const double* ra = sgeo + 4 * (ca >> 2);          // record as doubles: [ids | g0 | g1 | g2]
                    double dg = 3.0 * (ra[1] + ra[2] + ra[3]) - 4.0 * ra[1 + (ca & 3u)];
                    if (cb != 0xFFFFu) { const double* rb = sgeo + 4 * (cb >> 2); dg += 3.0 * (rb[1] + rb[2] + rb[3]) - 4.0 * rb[1 + (cb & 3u)]; }
                    pv = 1.0 / dg;
                } else pv = spv[le];
                const double xc = sq[le];
                const double r_ = bv - val;
                double xn = xc + c2 * pv * r_;
                if (use_prev) xn += c1 * (xc - sxp[le]);
                xnew[h.e_start + le] = xn;
                r0 += pv * r_ * r_;
                b0 += pv * bv * bv;
            }
        }
        __syncthreads();
        if (MULTI && !(pc.xch_dbg & 32)) pushed |= peer_push_patch<NT>(pc, 1, snd_b, snd_e, xnew, nxt);      // the CTA's last patch
        block_sum2(r0, b0, red);
        double* part = a.partials + (size_t)(k & 1) * 2 * gridDim.x;
        const bool stamp = MULTI && (pc.xch_dbg & 64) && k < 32 && threadIdx.x == 0;
        if (threadIdx.x == 0) {
            if (MULTI && pushed) __threadfence_system();    // one fence for everything this CTA pushed (ordered before it by the barriers above)
            part[2 * blockIdx.x] = r0; part[2 * blockIdx.x + 1] = b0;
            if (stamp) {
                const unsigned long long t_ = phw_globaltimer();
#ifndef PHW_HOST_EMU
                atomicMax(&a.ctl->xts[1][k][4], t_);
#endif
                if (blockIdx.x == 0) a.ctl->xts[1][k][0] = t_;
            }
        }
        grid_barrier_px(&a.ctl->gbar[0], epoch);
        if (stamp && blockIdx.x == 0) a.ctl->xts[1][k][1] = phw_globaltimer();
        if (threadIdx.x < 32) {
            double s0 = 0.0, s1 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) { s0 += __ldcg(&part[2 * i]); s1 += __ldcg(&part[2 * i + 1]); }
            s0 = warp_sum(s0); s1 = warp_sum(s1);
            if (threadIdx.x == 0) { bc[0] = s0; bc[1] = s1; }
        }
        __syncthreads();
        rr = bc[0]; bb = bc[1];
        __syncthreads();
        if (stamp && blockIdx.x == 0) a.ctl->xts[1][k][2] = phw_globaltimer();
        if (MULTI) {
            if (pc.xch_dbg & 32) { peer_exchange_iter<NT>(pc, a.ctl, epoch, false, xnew, nxt, k, rr, bb); }
            else peer_finish_iter<NT>(pc, a.ctl, xseq0 + (unsigned long long)k + 1ull, bc, rr, bb);
            fence_proxy_async();
            if (stamp && blockIdx.x == 0) a.ctl->xts[1][k][3] = phw_globaltimer();
        }
        conv = rr <= a.tol2 * bb;
        if (blockIdx.x == 0 && threadIdx.x == 0 && k < 48) a.ctl->reshist[a.which][k] = (bb > 0.0) ? rr / bb : 0.0;
        k += 1;
        if (conv) break;
        rho = rho_k; cur = nxt;
        if (k >= a.max_iter) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc.cur = cur; sc.k = k; sc.done = 1; sc.rr = rr; sc.bb = bb; sc.rho = rho;
        sc.iters += k; sc.solves += 1;
        if (MULTI && !(pc.xch_dbg & 32)) a.ctl->comm_seq = xseq0 + (unsigned long long)k;
        if (!conv) a.ctl->nonconv += 1;
        double rel = (bb > 0.0) ? sqrt(rr / bb) : 0.0;
        if (rel > sc.maxres) sc.maxres = rel;
    }
}

// ------------------------------------------------------------------------------------------------
// M_p solve on the assembled sliced-ELL rows.  Stage = [slice offsets (ids | weights) | x_k local | b | x_{k-1} |
// weights | ids]; plus two ghost-index lists.
// ------------------------------------------------------------------------------------------------
struct EllPipeArgs {
    int npatch;
    const PatchHdr* hdr;
    const uint32_t* ell_ioff; const uint32_t* ell_woff; const uint16_t* ell_ids; const double* ell_w;
    const int* ghost_v;
    const double* b;
    double* buf[3];
    double cheb_d, cheb_c, tol2;
    int max_iter, which;
    double* partials; Ctl* ctl;
    uint32_t off_wo, off_x, off_b, off_xprev, off_w, off_ids, stage_bytes, off_gidx, gidx_bytes;
};

template <int NT>
__device__ __forceinline__ void ell_issue(const EllPipeArgs& a, unsigned char* stage, const uint32_t* gidx_this, uint32_t* gidx_next,
                                          const PatchHdr& hn, const PatchHdr* hn2, const double* xk, const double* xprev,
                                          bool use_prev, unsigned long long* bar) {
    uint32_t* s_io = reinterpret_cast<uint32_t*>(stage);
    uint32_t* s_wo = reinterpret_cast<uint32_t*>(stage + a.off_wo);
    double* sx = reinterpret_cast<double*>(stage + a.off_x);
    double* sb = reinterpret_cast<double*>(stage + a.off_b);
    double* sxp = reinterpret_cast<double*>(stage + a.off_xprev);
    const int vs = hn.v_start, vc = hn.v_cnt;
    const uint32_t* gh = reinterpret_cast<const uint32_t*>(a.ghost_v);
    if (threadIdx.x == 0) {
        const uint32_t id_bytes = (uint32_t)hn.ell_in * 2u, w_bytes = (uint32_t)hn.ell_wn * 8u;
        uint32_t bytes = f64s_bytes(vs, vc) * (use_prev ? 3u : 2u) + id_bytes + w_bytes;
        if (hn2) bytes += u32s_bytes(hn2->gv_off, hn2->gv_cnt);
        mbar_arrive_expect_tx(bar, bytes);
        if (w_bytes) bulk_g2s(stage + a.off_w, a.ell_w + (uint32_t)hn.ell_w0, w_bytes, bar);
        if (id_bytes) bulk_g2s(stage + a.off_ids, a.ell_ids + (uint32_t)hn.ell_i0, id_bytes, bar);
        f64s_bulk(sx, xk, vs, vc, bar);
        f64s_bulk(sb, a.b, vs, vc, bar);
        if (use_prev) f64s_bulk(sxp, xprev, vs, vc, bar);
        if (hn2) u32s_bulk(gidx_next, gh, hn2->gv_off, hn2->gv_cnt, bar);
    }
    const int lane = (int)threadIdx.x - (NT - 32);
    if (lane >= 0) {
        if (lane < 2) f64s_fix(sx, xk, vs, vc, lane);
        else if (lane < 4) f64s_fix(sb, a.b, vs, vc, lane - 2);
        else if (lane < 6) { if (use_prev) f64s_fix(sxp, xprev, vs, vc, lane - 4); }
        else if (lane < 12) { if (hn2) u32s_fix(gidx_next, gh, hn2->gv_off, hn2->gv_cnt, lane - 6); }
    }
    // slice offsets of the patch (raw; the patch bases ell_i0 / ell_w0 are subtracted where they are used)
    const int nsl = (vc + 31) >> 5;
    for (int i = threadIdx.x; i <= nsl; i += NT) { cp_async4(&s_io[i], &a.ell_ioff[hn.vs_off + i]); cp_async4(&s_wo[i], &a.ell_woff[hn.vs_off + i]); }
    double* sg = sx + (vs & 1) + vc;
    if (gidx_this) {
        const uint32_t* gi = gidx_this + (hn.gv_off & 3);
        for (int i = threadIdx.x; i < hn.gv_cnt; i += NT) cp_async8(&sg[i], &xk[gi[i]]);
    } else {
        for (int i = threadIdx.x; i < hn.gv_cnt; i += NT) cp_async8(&sg[i], &xk[__ldg(&a.ghost_v[hn.gv_off + i])]);
    }
}

template <int NT, int MINB, bool MULTI>
__global__ void __launch_bounds__(NT, MINB) solve_ell_pipe_k(const __grid_constant__ EllPipeArgs a, const __grid_constant__ PeerComm pc) {
    PHW_DYN_SMEM(smraw);
    __shared__ double red[64];
    __shared__ double bc[2];
    __shared__ PatchHdr s_hdr[4];
    __shared__ __align__(8) unsigned long long full[2];
    __shared__ int s_dead;       // a stage did not arrive (see mbar_wait): the CTA stops touching the stages
    __shared__ int s_snd[4][2];  // MULTI: send-list range of the patches in s_hdr
    SolveCtl& sc = a.ctl->sc[a.which];
    int k = 0, cur = sc.cur;
    double rho = 0.0, rr = 0.0, bb = 0.0;
    bool conv = false;
    unsigned int epoch = 0;
    unsigned int seq = 0;
    const unsigned long long xseq0 = MULTI ? a.ctl->comm_seq : 0ull;   // exchanges done before this launch (updated only at its end)
    if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); s_dead = 0; }
    __syncthreads();
    const int n_my = ((int)blockIdx.x < a.npatch) ? (a.npatch - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    while (true) {
        double c1, c2, rho_k;
        cheb_coeffs(k, rho, a.cheb_d, a.cheb_c, c1, c2, rho_k);
        const int nxt = (cur + 1) % 3, prv = (cur + 2) % 3;
        const double* xk = cur == 0 ? a.buf[0] : (cur == 1 ? a.buf[1] : a.buf[2]);
        const double* xprev = prv == 0 ? a.buf[0] : (prv == 1 ? a.buf[1] : a.buf[2]);
        double* xnew = nxt == 0 ? a.buf[0] : (nxt == 1 ? a.buf[1] : a.buf[2]);
        const bool use_prev = k > 0;
        double r0 = 0.0, b0 = 0.0;
        for (int j = 0; j < 3 && j < n_my; ++j) {
            load_hdr_smem_at(&s_hdr[j], &a.hdr[blockIdx.x + j * gridDim.x], 32 * j);
            if (MULTI) load_snd_smem_at(s_snd[j], pc.psend_off[0], blockIdx.x + j * gridDim.x, 32 * j);
        }
        int snd_b = 0, snd_e = 0;       // MULTI: send-list range of the patch whose rows were written last
        bool pushed = false;
        __syncthreads();
        if (n_my > 0)
            ell_issue<NT>(a, smraw + (seq & 1u) * a.stage_bytes, nullptr,
                          reinterpret_cast<uint32_t*>(smraw + a.off_gidx + ((seq + 1u) & 1u) * a.gidx_bytes),
                          s_hdr[0], n_my > 1 ? &s_hdr[1] : nullptr, xk, xprev, use_prev, &full[seq & 1u]);
        for (int it = 0; it < n_my; ++it, ++seq) {
            const unsigned int s = seq & 1u;
            if (!pipe_stage_ready(&full[s], (seq >> 1) & 1u, a.ctl, &s_dead)) continue;
            const PatchHdr& h = s_hdr[it & 3];
            if (MULTI) {   // the rows of the previous patch are visible to the whole CTA now: push its interface rows to the neighbours
                if (!(pc.xch_dbg & 32)) pushed |= peer_push_patch<NT>(pc, 0, snd_b, snd_e, xnew, nxt);
                snd_b = s_snd[it & 3][0]; snd_e = s_snd[it & 3][1];
            }
            if (it + 3 < n_my) {
                load_hdr_smem_at(&s_hdr[(it + 3) & 3], &a.hdr[blockIdx.x + (it + 3) * gridDim.x], 64);
                if (MULTI) load_snd_smem_at(s_snd[(it + 3) & 3], pc.psend_off[0], blockIdx.x + (it + 3) * gridDim.x, 64);
            }
            if (it + 1 < n_my)
                ell_issue<NT>(a, smraw + (s ^ 1u) * a.stage_bytes,
                              reinterpret_cast<const uint32_t*>(smraw + a.off_gidx + (s ^ 1u) * a.gidx_bytes),
                              reinterpret_cast<uint32_t*>(smraw + a.off_gidx + s * a.gidx_bytes),
                              s_hdr[(it + 1) & 3], it + 2 < n_my ? &s_hdr[(it + 2) & 3] : nullptr, xk, xprev, use_prev, &full[s ^ 1u]);
            const unsigned char* stage = smraw + s * a.stage_bytes;
            const uint32_t* s_io = reinterpret_cast<const uint32_t*>(stage);
            const uint32_t* s_wo = reinterpret_cast<const uint32_t*>(stage + a.off_wo);
            const double* sx = reinterpret_cast<const double*>(stage + a.off_x) + (h.v_start & 1);
            const double* sb = reinterpret_cast<const double*>(stage + a.off_b) + (h.v_start & 1);
            const double* sxp = reinterpret_cast<const double*>(stage + a.off_xprev) + (h.v_start & 1);
            const double* sw = reinterpret_cast<const double*>(stage + a.off_w);
            const uint4* sid4 = reinterpret_cast<const uint4*>(stage + a.off_ids);
            const uint32_t i0 = (uint32_t)h.ell_i0, w0 = (uint32_t)h.ell_w0;
            for (int lv = threadIdx.x; lv < h.v_cnt; lv += NT) {
                const int sl = lv >> 5, lane = lv & 31;
                const uint32_t io0 = s_io[sl] - i0, wo0 = s_wo[sl] - w0;
                const int W = (int)((s_wo[sl + 1] - w0 - wo0) >> 5);
                double val = 0.0, dsum = 0.0;
                for (int c = 0; 8 * c < W; ++c) {
                    const uint4 cd = sid4[(io0 >> 3) + c * 32 + lane];
                    const uint32_t w4[4] = {cd.x, cd.y, cd.z, cd.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double wa = (8 * c + 2 * j < W) ? sw[wo0 + (8 * c + 2 * j) * 32 + lane] : 0.0;
                        const double wb = (8 * c + 2 * j + 1 < W) ? sw[wo0 + (8 * c + 2 * j + 1) * 32 + lane] : 0.0;
                        val += wa * sx[w4[j] & 0xFFFFu];
                        val += wb * sx[w4[j] >> 16];
                        dsum += wa + wb;
                    }
                }
                // the P1 mass diagonal is the sum of the row's off-diagonal entries, so neither it nor the Jacobi weight is read
                const double xc = sx[lv], bv = sb[lv];
                const double pv = 1.0 / dsum;
                val += dsum * xc;
                const double r_ = bv - val;
                double xn = xc + c2 * pv * r_;
                if (use_prev) xn += c1 * (xc - sxp[lv]);
                xnew[h.v_start + lv] = xn;
                r0 += pv * r_ * r_;
                b0 += pv * bv * bv;
            }
        }
        __syncthreads();
        if (MULTI && !(pc.xch_dbg & 32)) pushed |= peer_push_patch<NT>(pc, 0, snd_b, snd_e, xnew, nxt);      // the CTA's last patch
        block_sum2(r0, b0, red);
        double* part = a.partials + (size_t)(k & 1) * 2 * gridDim.x;
        const bool stamp = MULTI && (pc.xch_dbg & 64) && k < 32 && threadIdx.x == 0;
        if (threadIdx.x == 0) {
            if (MULTI && pushed) __threadfence_system();    // one fence for everything this CTA pushed (ordered before it by the barriers above)
            part[2 * blockIdx.x] = r0; part[2 * blockIdx.x + 1] = b0;
            if (stamp) {
                const unsigned long long t_ = phw_globaltimer();
#ifndef PHW_HOST_EMU
                atomicMax(&a.ctl->xts[0][k][4], t_);
#endif
                if (blockIdx.x == 0) a.ctl->xts[0][k][0] = t_;
            }
        }
        grid_barrier_px(&a.ctl->gbar[0], epoch);
        if (stamp && blockIdx.x == 0) a.ctl->xts[0][k][1] = phw_globaltimer();
        if (threadIdx.x < 32) {
            double s0 = 0.0, s1 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) { s0 += __ldcg(&part[2 * i]); s1 += __ldcg(&part[2 * i + 1]); }
            s0 = warp_sum(s0); s1 = warp_sum(s1);
            if (threadIdx.x == 0) { bc[0] = s0; bc[1] = s1; }
        }
        __syncthreads();
        rr = bc[0]; bb = bc[1];
        __syncthreads();
        if (stamp && blockIdx.x == 0) a.ctl->xts[0][k][2] = phw_globaltimer();
        if (MULTI) {
            if (pc.xch_dbg & 32) { peer_exchange_iter<NT>(pc, a.ctl, epoch, true, xnew, nxt, k, rr, bb); }
            else peer_finish_iter<NT>(pc, a.ctl, xseq0 + (unsigned long long)k + 1ull, bc, rr, bb);
            fence_proxy_async();
            if (stamp && blockIdx.x == 0) a.ctl->xts[0][k][3] = phw_globaltimer();
        }
        conv = rr <= a.tol2 * bb;
        if (blockIdx.x == 0 && threadIdx.x == 0 && k < 48) a.ctl->reshist[a.which][k] = (bb > 0.0) ? rr / bb : 0.0;
        k += 1;
        if (conv) break;
        rho = rho_k; cur = nxt;
        if (k >= a.max_iter) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc.cur = cur; sc.k = k; sc.done = 1; sc.rr = rr; sc.bb = bb; sc.rho = rho;
        sc.iters += k; sc.solves += 1;
        if (MULTI && !(pc.xch_dbg & 32)) a.ctl->comm_seq = xseq0 + (unsigned long long)k;
        if (!conv) a.ctl->nonconv += 1;
        double rel = (bb > 0.0) ? sqrt(rr / bb) : 0.0;
        if (rel > sc.maxres) sc.maxres = rel;
    }
}

// ------------------------------------------------------------------------------------------------
// H_wave in ONE pipelined pass (replaces the two assemble(scalar) calls of wave_2D.py:882,907): both quadratic forms are
// sums of cell-local forms over the OWN cells of every patch,
//   p^T M_p p = sum_K |K|/12 (sum_i p_i^2 + (sum_i p_i)^2),     q^T M_q q = sum_K t^T M_K t,  t_i = s_i q_i,
// so no row gather (adjacency codes, per-cell contribution buffer) and no halo cells are needed: 64 B per cell
// (vertex ids 8, area 8, edge ids + metric 32, p 4, q 12) instead of the 88 B of the two row-based reductions.
// Stage = [p local | q local | vertex-id records | areas | edge records]; two ghost lists per field.
// ------------------------------------------------------------------------------------------------
struct EnergyPipeArgs {
    int npatch;
    const PatchHdr* hdr;
    const uint2* recv; const double* area; const uint4* cellq;
    const int* ghost_v; const int* ghost_e;
    const double* p; const double* q;
    double scale_p, scale_q;         // red[0] = scale_p * p^T M_p p, red[1] = scale_q * q^T M_q q
    int counter_id;
    double* partials; Ctl* ctl;
    uint32_t off_q, off_recv, off_area, off_cells, stage_bytes, off_gv, gv_bytes, off_ge, ge_bytes;
};

template <int NT>
__device__ __forceinline__ void en_issue(const EnergyPipeArgs& a, unsigned char* stage, const uint32_t* gv_this, const uint32_t* ge_this,
                                         uint32_t* gv_next, uint32_t* ge_next, const PatchHdr& hn, const PatchHdr* hn2,
                                         unsigned long long* bar) {
    double* sp = reinterpret_cast<double*>(stage);
    double* sq = reinterpret_cast<double*>(stage + a.off_q);
    double* srv = reinterpret_cast<double*>(stage + a.off_recv);
    double* sar = reinterpret_cast<double*>(stage + a.off_area);
    const double* grecv = reinterpret_cast<const double*>(a.recv);
    const uint32_t* ghv = reinterpret_cast<const uint32_t*>(a.ghost_v);
    const uint32_t* ghe = reinterpret_cast<const uint32_t*>(a.ghost_e);
    const int vs = hn.v_start, vc = hn.v_cnt, es = hn.e_start, ec = hn.e_cnt, co = hn.cell_off, no = hn.n_own;
    if (threadIdx.x == 0) {
        const uint32_t cell_bytes = (uint32_t)no * 32u;
        uint32_t bytes = f64s_bytes(vs, vc) + f64s_bytes(es, ec) + 2u * f64s_bytes(co, no) + cell_bytes;
        if (hn2) bytes += u32s_bytes(hn2->gv_off, hn2->gv_cnt) + u32s_bytes(hn2->ge_off, hn2->ge_cnt);
        mbar_arrive_expect_tx(bar, bytes);
        if (cell_bytes) bulk_g2s(stage + a.off_cells, a.cellq + 2 * (size_t)co, cell_bytes, bar);
        f64s_bulk(sq, a.q, es, ec, bar);
        f64s_bulk(sp, a.p, vs, vc, bar);
        f64s_bulk(srv, grecv, co, no, bar);
        f64s_bulk(sar, a.area, co, no, bar);
        if (hn2) { u32s_bulk(gv_next, ghv, hn2->gv_off, hn2->gv_cnt, bar); u32s_bulk(ge_next, ghe, hn2->ge_off, hn2->ge_cnt, bar); }
    }
    const int lane = (int)threadIdx.x - (NT - 32);
    if (lane >= 0) {
        if (lane < 2) f64s_fix(sp, a.p, vs, vc, lane);
        else if (lane < 4) f64s_fix(sq, a.q, es, ec, lane - 2);
        else if (lane < 6) f64s_fix(srv, grecv, co, no, lane - 4);
        else if (lane < 8) f64s_fix(sar, a.area, co, no, lane - 6);
        else if (lane < 14) { if (hn2) u32s_fix(gv_next, ghv, hn2->gv_off, hn2->gv_cnt, lane - 8); }
        else if (lane < 20) { if (hn2) u32s_fix(ge_next, ghe, hn2->ge_off, hn2->ge_cnt, lane - 14); }
    }
    double* sgp = sp + (vs & 1) + vc;
    double* sgq = sq + (es & 1) + ec;
    if (gv_this) {
        const uint32_t* gi = gv_this + (hn.gv_off & 3);
        for (int i = threadIdx.x; i < hn.gv_cnt; i += NT) cp_async8(&sgp[i], &a.p[gi[i]]);
        const uint32_t* gj = ge_this + (hn.ge_off & 3);
        for (int i = threadIdx.x; i < hn.ge_cnt; i += NT) cp_async8(&sgq[i], &a.q[gj[i]]);
    } else {
        for (int i = threadIdx.x; i < hn.gv_cnt; i += NT) cp_async8(&sgp[i], &a.p[__ldg(&a.ghost_v[hn.gv_off + i])]);
        for (int i = threadIdx.x; i < hn.ge_cnt; i += NT) cp_async8(&sgq[i], &a.q[__ldg(&a.ghost_e[hn.ge_off + i])]);
    }
}

template <int NT>
__global__ void __launch_bounds__(NT, 1) energy_pipe_k(const __grid_constant__ EnergyPipeArgs a) {
    PHW_DYN_SMEM(smraw);
    __shared__ double red[64];
    __shared__ PatchHdr s_hdr[4];
    __shared__ __align__(8) unsigned long long full[2];
    __shared__ int s_dead;       // a stage did not arrive (see mbar_wait): the CTA stops touching the stages
    __shared__ bool is_last;
    if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); s_dead = 0; }
    const int n_my = ((int)blockIdx.x < a.npatch) ? (a.npatch - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    for (int j = 0; j < 3 && j < n_my; ++j) load_hdr_smem_at(&s_hdr[j], &a.hdr[blockIdx.x + j * gridDim.x], 32 * j);
    __syncthreads();
    double accp = 0.0, accq = 0.0;
    auto gv_slot = [&](unsigned int i) { return reinterpret_cast<uint32_t*>(smraw + a.off_gv + (i & 1u) * a.gv_bytes); };
    auto ge_slot = [&](unsigned int i) { return reinterpret_cast<uint32_t*>(smraw + a.off_ge + (i & 1u) * a.ge_bytes); };
    if (n_my > 0)
        en_issue<NT>(a, smraw, nullptr, nullptr, gv_slot(1), ge_slot(1), s_hdr[0], n_my > 1 ? &s_hdr[1] : nullptr, &full[0]);
    for (int it = 0; it < n_my; ++it) {
        const unsigned int s = (unsigned int)it & 1u;
        if (!pipe_stage_ready(&full[s], ((unsigned int)it >> 1) & 1u, a.ctl, &s_dead)) continue;
        const PatchHdr& h = s_hdr[it & 3];
        if (it + 3 < n_my) load_hdr_smem_at(&s_hdr[(it + 3) & 3], &a.hdr[blockIdx.x + (it + 3) * gridDim.x], 64);
        if (it + 1 < n_my)
            en_issue<NT>(a, smraw + (s ^ 1u) * a.stage_bytes, gv_slot(it + 1), ge_slot(it + 1), gv_slot(it + 2), ge_slot(it + 2),
                         s_hdr[(it + 1) & 3], it + 2 < n_my ? &s_hdr[(it + 2) & 3] : nullptr, &full[s ^ 1u]);
        const unsigned char* stage = smraw + s * a.stage_bytes;
        const double* sp = reinterpret_cast<const double*>(stage) + (h.v_start & 1);
        const double* sq = reinterpret_cast<const double*>(stage + a.off_q) + (h.e_start & 1);
        const uint2* srv = reinterpret_cast<const uint2*>(stage + a.off_recv) + (h.cell_off & 1);
        const double* sar = reinterpret_cast<const double*>(stage + a.off_area) + (h.cell_off & 1);
        const uint4* scell = reinterpret_cast<const uint4*>(stage + a.off_cells);
        for (int lc = threadIdx.x; lc < h.n_own; lc += NT) {
            const uint2 rv = srv[lc];
            const double p0 = sp[rv.x & 0xFFFFu], p1 = sp[rv.x >> 16], p2 = sp[rv.y & 0xFFFFu];
            const double ps = p0 + p1 + p2;
            const double wgt = (rv.y & 0x400000u) ? 0.0 : 1.0;      // halo cell of another rank (partitioned run): counted there
            accp += wgt * sar[lc] * (p0 * p0 + p1 * p1 + p2 * p2 + ps * ps);
            const uint4 qa = scell[2 * lc], qb = scell[2 * lc + 1];
            const uint32_t e0 = qa.x & 0xFFFFu, e1 = qa.x >> 16, e2 = qa.y & 0xFFFFu;
            const double g0 = __hiloint2double((int)qa.w, (int)qa.z);
            const double g1 = __hiloint2double((int)qb.y, (int)qb.x), g2 = __hiloint2double((int)qb.w, (int)qb.z);
            const double S = g0 + g1 + g2;
            const double s0 = (e0 & 0x8000u) ? -1.0 : 1.0, s1 = (e1 & 0x8000u) ? -1.0 : 1.0, s2 = (e2 & 0x8000u) ? -1.0 : 1.0;
            const double t0 = s0 * sq[e0 & 0x3FFFu], t1 = s1 * sq[e1 & 0x3FFFu], t2 = s2 * sq[e2 & 0x3FFFu];
            const double ST = S * (t0 + t1 + t2);
            accq += wgt * (t0 * (ST + 2.0 * S * t0 - 4.0 * (g0 * t0 + g2 * t1 + g1 * t2))
                         + t1 * (ST + 2.0 * S * t1 - 4.0 * (g2 * t0 + g1 * t1 + g0 * t2))
                         + t2 * (ST + 2.0 * S * t2 - 4.0 * (g1 * t0 + g0 * t1 + g2 * t2)));
        }
    }
    __syncthreads();
    block_sum2(accp, accq, red);
    if (threadIdx.x == 0) {
        a.partials[2 * blockIdx.x] = accp; a.partials[2 * blockIdx.x + 1] = accq;
        __threadfence();
        is_last = (atomicAdd(&a.ctl->counters[a.counter_id], 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double tp = 0.0, tq = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += NT) { tp += __ldcg(&a.partials[2 * b]); tq += __ldcg(&a.partials[2 * b + 1]); }
    block_sum2(tp, tq, red);
    if (threadIdx.x == 0) { a.ctl->counters[a.counter_id] = 0; a.ctl->red[0] = a.scale_p * tp; a.ctl->red[1] = a.scale_q * tq; }
}

// ------------------------------------------------------------------------------------------------
// right-hand sides (replace assemble(b), wave_2D.py:814,849), pipelined:  y_e = cD (D x_p)_e  and  y_v = cD (D^T x_q)_v.
// Same per-cell contributions and the same fixed summation order as edge_rows_k / vertex_rows_k <STORE, HAS_D>.
// ------------------------------------------------------------------------------------------------
struct RhsPipeArgs {
    int npatch;
    const PatchHdr* hdr;
    const uint2* recv; const uint2* rece;
    const int* ghost;                 // ghost list of the INPUT field (vertices for the edge rows, edges for the vertex rows)
    const uint32_t* eadj;             // edge rows: (cell, slot) codes of every owned edge
    const uint32_t* vslice_off; const uint16_t* vadj;   // vertex rows: sliced-ELL adjacency
    const double* x;                  // input field
    double* y;                        // output rows
    double cD;
    Ctl* ctl;
    uint32_t off_a, off_b, off_c, stage_bytes, off_sc3, off_gidx, gidx_bytes;
};

// edge rows.  Stage = [x_p local | vertex-id records | edge-id records | eadj]
template <int NT>
__device__ __forceinline__ void rhs_e_issue(const RhsPipeArgs& a, unsigned char* stage, const uint32_t* gidx_this, uint32_t* gidx_next,
                                            const PatchHdr& hn, const PatchHdr* hn2, unsigned long long* bar) {
    double* sp = reinterpret_cast<double*>(stage);
    double* srv = reinterpret_cast<double*>(stage + a.off_a);
    double* sre = reinterpret_cast<double*>(stage + a.off_b);
    uint32_t* sea = reinterpret_cast<uint32_t*>(stage + a.off_c);
    const double* grv = reinterpret_cast<const double*>(a.recv);
    const double* gre = reinterpret_cast<const double*>(a.rece);
    const uint32_t* gh = reinterpret_cast<const uint32_t*>(a.ghost);
    const int vs = hn.v_start, vc = hn.v_cnt, es = hn.e_start, ec = hn.e_cnt, co = hn.cell_off, ncl = hn.n_ecells;
    if (threadIdx.x == 0) {
        uint32_t bytes = f64s_bytes(vs, vc) + 2u * f64s_bytes(co, ncl) + u32s_bytes(es, ec);
        if (hn2) bytes += u32s_bytes(hn2->gv_off, hn2->gv_cnt);
        mbar_arrive_expect_tx(bar, bytes);
        f64s_bulk(srv, grv, co, ncl, bar);
        f64s_bulk(sre, gre, co, ncl, bar);
        f64s_bulk(sp, a.x, vs, vc, bar);
        u32s_bulk(sea, a.eadj, es, ec, bar);
        if (hn2) u32s_bulk(gidx_next, gh, hn2->gv_off, hn2->gv_cnt, bar);
    }
    const int lane = (int)threadIdx.x - (NT - 32);
    if (lane >= 0) {
        if (lane < 2) f64s_fix(sp, a.x, vs, vc, lane);
        else if (lane < 4) f64s_fix(srv, grv, co, ncl, lane - 2);
        else if (lane < 6) f64s_fix(sre, gre, co, ncl, lane - 4);
        else if (lane < 12) u32s_fix(sea, a.eadj, es, ec, lane - 6);
        else if (lane < 18) { if (hn2) u32s_fix(gidx_next, gh, hn2->gv_off, hn2->gv_cnt, lane - 12); }
    }
    double* sg = sp + (vs & 1) + vc;
    if (gidx_this) {
        const uint32_t* gi = gidx_this + (hn.gv_off & 3);
        for (int i = threadIdx.x; i < hn.gv_cnt; i += NT) cp_async8(&sg[i], &a.x[gi[i]]);
    } else {
        for (int i = threadIdx.x; i < hn.gv_cnt; i += NT) cp_async8(&sg[i], &a.x[__ldg(&a.ghost[hn.gv_off + i])]);
    }
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) rhs_edge_pipe_k(const __grid_constant__ RhsPipeArgs a) {
    PHW_DYN_SMEM(smraw);
    __shared__ PatchHdr s_hdr[4];
    __shared__ __align__(8) unsigned long long full[2];
    __shared__ int s_dead;       // a stage did not arrive (see mbar_wait): the CTA stops touching the stages
    if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); s_dead = 0; }
    const int n_my = ((int)blockIdx.x < a.npatch) ? (a.npatch - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    for (int j = 0; j < 3 && j < n_my; ++j) load_hdr_smem_at(&s_hdr[j], &a.hdr[blockIdx.x + j * gridDim.x], 32 * j);
    __syncthreads();
    double* sc3 = reinterpret_cast<double*>(smraw + a.off_sc3);
    auto g_slot = [&](unsigned int i) { return reinterpret_cast<uint32_t*>(smraw + a.off_gidx + (i & 1u) * a.gidx_bytes); };
    const double cD = a.cD;
    if (n_my > 0) rhs_e_issue<NT>(a, smraw, nullptr, g_slot(1), s_hdr[0], n_my > 1 ? &s_hdr[1] : nullptr, &full[0]);
    for (int it = 0; it < n_my; ++it) {
        const unsigned int s = (unsigned int)it & 1u;
        if (!pipe_stage_ready(&full[s], ((unsigned int)it >> 1) & 1u, a.ctl, &s_dead)) continue;
        const PatchHdr& h = s_hdr[it & 3];
        if (it + 3 < n_my) load_hdr_smem_at(&s_hdr[(it + 3) & 3], &a.hdr[blockIdx.x + (it + 3) * gridDim.x], 64);
        if (it + 1 < n_my)
            rhs_e_issue<NT>(a, smraw + (s ^ 1u) * a.stage_bytes, g_slot(it + 1), g_slot(it + 2), s_hdr[(it + 1) & 3],
                            it + 2 < n_my ? &s_hdr[(it + 2) & 3] : nullptr, &full[s ^ 1u]);
        const unsigned char* stage = smraw + s * a.stage_bytes;
        const double* sp = reinterpret_cast<const double*>(stage) + (h.v_start & 1);
        const uint2* srv = reinterpret_cast<const uint2*>(stage + a.off_a) + (h.cell_off & 1);
        const uint2* sre = reinterpret_cast<const uint2*>(stage + a.off_b) + (h.cell_off & 1);
        const uint32_t* sea = reinterpret_cast<const uint32_t*>(stage + a.off_c) + (h.e_start & 3);
        for (int lc = threadIdx.x; lc < h.n_ecells; lc += NT) {
            const uint2 re = sre[lc], rv = srv[lc];
            const uint32_t e0 = re.x & 0xFFFFu, e1 = re.x >> 16, e2 = re.y & 0xFFFFu;
            const double s0 = (e0 & 0x8000u) ? -1.0 : 1.0, s1 = (e1 & 0x8000u) ? -1.0 : 1.0, s2 = (e2 & 0x8000u) ? -1.0 : 1.0;
            const double p0 = sp[rv.x & 0xFFFFu], p1 = sp[rv.x >> 16], p2 = sp[rv.y & 0xFFFFu];
            const double third = (p0 + p1 + p2) * (1.0 / 3.0);
            double d0 = third, d1 = third, d2 = third;
            if (e0 & 0x4000u) d0 -= 0.5 * (p1 + p2);      // zero-flux sides: -(1/2) (p_a + p_b) at the edge's end vertices
            if (e1 & 0x4000u) d1 -= 0.5 * (p0 + p2);
            if (e2 & 0x4000u) d2 -= 0.5 * (p0 + p1);
            double v0 = 0.0, v1 = 0.0, v2 = 0.0;
            v0 += cD * s0 * d0; v1 += cD * s1 * d1; v2 += cD * s2 * d2;
            sc3[3 * lc] = v0; sc3[3 * lc + 1] = v1; sc3[3 * lc + 2] = v2;
        }
        __syncthreads();
        for (int le = threadIdx.x; le < h.e_cnt; le += NT) {
            const uint32_t adj = sea[le];
            const uint32_t ca = adj & 0xFFFFu, cb = adj >> 16;
            double val = sc3[3 * (ca >> 2) + (ca & 3u)];
            if (cb != 0xFFFFu) val += sc3[3 * (cb >> 2) + (cb & 3u)];
            a.y[h.e_start + le] = val;
        }
    }
}

// vertex rows.  Stage = [slice offsets | x_q local | edge-id records | adjacency codes of the patch]
template <int NT>
__device__ __forceinline__ void rhs_v_issue(const RhsPipeArgs& a, unsigned char* stage, const uint32_t* gidx_this, uint32_t* gidx_next,
                                            const PatchHdr& hn, const PatchHdr* hn2, unsigned long long* bar) {
    uint32_t* s_off = reinterpret_cast<uint32_t*>(stage);
    double* sq = reinterpret_cast<double*>(stage + a.off_a);
    double* sre = reinterpret_cast<double*>(stage + a.off_b);
    const double* gre = reinterpret_cast<const double*>(a.rece);
    const uint32_t* gh = reinterpret_cast<const uint32_t*>(a.ghost);
    const int es = hn.e_start, ec = hn.e_cnt, co = hn.cell_off, ncl = hn.n_vcells;
    if (threadIdx.x == 0) {
        const uint32_t adj_bytes = (uint32_t)hn.adj_cnt * 2u;
        uint32_t bytes = f64s_bytes(es, ec) + f64s_bytes(co, ncl) + adj_bytes;
        if (hn2) bytes += u32s_bytes(hn2->ge_off, hn2->ge_cnt);
        mbar_arrive_expect_tx(bar, bytes);
        if (adj_bytes) bulk_g2s(stage + a.off_c, a.vadj + (uint32_t)hn.adj_off, adj_bytes, bar);
        f64s_bulk(sre, gre, co, ncl, bar);
        f64s_bulk(sq, a.x, es, ec, bar);
        if (hn2) u32s_bulk(gidx_next, gh, hn2->ge_off, hn2->ge_cnt, bar);
    }
    const int lane = (int)threadIdx.x - (NT - 32);
    if (lane >= 0) {
        if (lane < 2) f64s_fix(sq, a.x, es, ec, lane);
        else if (lane < 4) f64s_fix(sre, gre, co, ncl, lane - 2);
        else if (lane < 10) { if (hn2) u32s_fix(gidx_next, gh, hn2->ge_off, hn2->ge_cnt, lane - 4); }
    }
    const int nsl = (hn.v_cnt + 31) >> 5;
    for (int i = threadIdx.x; i <= nsl; i += NT) cp_async4(&s_off[i], &a.vslice_off[hn.vs_off + i]);
    double* sg = sq + (es & 1) + ec;
    if (gidx_this) {
        const uint32_t* gi = gidx_this + (hn.ge_off & 3);
        for (int i = threadIdx.x; i < hn.ge_cnt; i += NT) cp_async8(&sg[i], &a.x[gi[i]]);
    } else {
        for (int i = threadIdx.x; i < hn.ge_cnt; i += NT) cp_async8(&sg[i], &a.x[__ldg(&a.ghost[hn.ge_off + i])]);
    }
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) rhs_vertex_pipe_k(const __grid_constant__ RhsPipeArgs a) {
    PHW_DYN_SMEM(smraw);
    __shared__ PatchHdr s_hdr[4];
    __shared__ __align__(8) unsigned long long full[2];
    __shared__ int s_dead;       // a stage did not arrive (see mbar_wait): the CTA stops touching the stages
    if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); s_dead = 0; }
    const int n_my = ((int)blockIdx.x < a.npatch) ? (a.npatch - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    for (int j = 0; j < 3 && j < n_my; ++j) load_hdr_smem_at(&s_hdr[j], &a.hdr[blockIdx.x + j * gridDim.x], 32 * j);
    __syncthreads();
    double* sc3 = reinterpret_cast<double*>(smraw + a.off_sc3);
    auto g_slot = [&](unsigned int i) { return reinterpret_cast<uint32_t*>(smraw + a.off_gidx + (i & 1u) * a.gidx_bytes); };
    const double cD = a.cD;
    if (n_my > 0) rhs_v_issue<NT>(a, smraw, nullptr, g_slot(1), s_hdr[0], n_my > 1 ? &s_hdr[1] : nullptr, &full[0]);
    for (int it = 0; it < n_my; ++it) {
        const unsigned int s = (unsigned int)it & 1u;
        if (!pipe_stage_ready(&full[s], ((unsigned int)it >> 1) & 1u, a.ctl, &s_dead)) continue;
        const PatchHdr& h = s_hdr[it & 3];
        if (it + 3 < n_my) load_hdr_smem_at(&s_hdr[(it + 3) & 3], &a.hdr[blockIdx.x + (it + 3) * gridDim.x], 64);
        if (it + 1 < n_my)
            rhs_v_issue<NT>(a, smraw + (s ^ 1u) * a.stage_bytes, g_slot(it + 1), g_slot(it + 2), s_hdr[(it + 1) & 3],
                            it + 2 < n_my ? &s_hdr[(it + 2) & 3] : nullptr, &full[s ^ 1u]);
        const unsigned char* stage = smraw + s * a.stage_bytes;
        const uint32_t* s_off = reinterpret_cast<const uint32_t*>(stage);
        const double* sq = reinterpret_cast<const double*>(stage + a.off_a) + (h.e_start & 1);
        const uint2* sre = reinterpret_cast<const uint2*>(stage + a.off_b) + (h.cell_off & 1);
        const uint4* sadj4 = reinterpret_cast<const uint4*>(stage + a.off_c);
        for (int lc = threadIdx.x; lc < h.n_vcells; lc += NT) {
            const uint2 re = sre[lc];
            const uint32_t e0 = re.x & 0xFFFFu, e1 = re.x >> 16, e2 = re.y & 0xFFFFu;
            const double t0 = (e0 & 0x8000u) ? -sq[e0 & 0x3FFFu] : sq[e0 & 0x3FFFu];
            const double t1 = (e1 & 0x8000u) ? -sq[e1 & 0x3FFFu] : sq[e1 & 0x3FFFu];
            const double t2 = (e2 & 0x8000u) ? -sq[e2 & 0x3FFFu] : sq[e2 & 0x3FFFu];
            const double common = cD * (t0 + t1 + t2) * (1.0 / 3.0);
            double d0 = common, d1 = common, d2 = common;
            // zero-flux sides: -(1/2) s_e q_e at both end vertices of edge i (the vertices != i)
            if (e0 & 0x4000u) { const double hq = 0.5 * cD * t0; d1 -= hq; d2 -= hq; }
            if (e1 & 0x4000u) { const double hq = 0.5 * cD * t1; d0 -= hq; d2 -= hq; }
            if (e2 & 0x4000u) { const double hq = 0.5 * cD * t2; d0 -= hq; d1 -= hq; }
            double v0 = 0.0, v1 = 0.0, v2 = 0.0;
            v0 += d0; v1 += d1; v2 += d2;
            sc3[3 * lc] = v0; sc3[3 * lc + 1] = v1; sc3[3 * lc + 2] = v2;
        }
        __syncthreads();
        const uint32_t adj0 = (uint32_t)h.adj_off;
        for (int lv = threadIdx.x; lv < h.v_cnt; lv += NT) {
            const uint32_t o0 = s_off[lv >> 5] - adj0, o1 = s_off[(lv >> 5) + 1] - adj0;
            double val = 0.0;
            const uint32_t nchunk = (o1 - o0) >> 8;            // 32 lanes * 8 codes per chunk
            for (uint32_t k = 0; k < nchunk; ++k) {
                const uint4 cd = sadj4[(o0 >> 3) + k * 32 + (lv & 31)];
                const uint32_t w4[4] = {cd.x, cd.y, cd.z, cd.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t c0 = w4[j] & 0xFFFFu, c1 = w4[j] >> 16;
                    if (c0 != 0xFFFFu) val += sc3[3 * (c0 >> 2) + (c0 & 3u)];
                    if (c1 != 0xFFFFu) val += sc3[3 * (c1 >> 2) + (c1 & 3u)];
                }
            }
            a.y[h.v_start + lv] = val;
        }
    }
}

}  // namespace phw
