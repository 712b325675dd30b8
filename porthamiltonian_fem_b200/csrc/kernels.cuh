// kernels.cuh - sm_100a kernels of the port-Hamiltonian wave hot path (P1 x RT1, fp64).
//
// Two "row" kernel families evaluate, patch by patch and entirely from shared memory,
//   vertex rows:  val_v = cM (M_p x_p)_v + cD (D^T x_q)_v + cB (B_top x_p)_v
//   edge rows:    val_e = cM (M_q x_q)_e + cD (D x_p)_e   + cB (B_right x_q)_e
// i.e. every block of the reference's mixed system  [[M_p, dt K D^T], [-dt/rho D, M_q]]
// (wave_2D.py:449-518; casimir boundary blocks :327-351), followed by a fused epilogue:
//   STORE  y = val                      -> right-hand sides  (replaces assemble(b), wave_2D.py:814,849)
//   CHEB   x+ = x + c1 (x - x-) + c2 P^-1 (b - val), accumulating ||b - val||^2
//                                        -> one Chebyshev iteration (replaces solve(A, x, b), :820,854)
//   DOT    sum += x . val                -> quadratic energy forms (replaces assemble(scalar), :882,907)
// A patch = owned dof ranges + halo cells (layout.hpp).  Stage 1 stages the local input values in
// shared memory (owned range: coalesced; ghosts: gathered), stage 2 computes the three per-cell
// contributions into shared memory (cell records and geometry streamed once, coalesced), stage 3
// sums the contributions of each owned dof in a fixed order (deterministic; no atomics on data) and
// applies the epilogue with coalesced global traffic.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "layout.hpp"

// loads in flight per thread in the batched stages (memory-level parallelism vs registers)
#ifndef PHW_U_VCELLS
#define PHW_U_VCELLS 3
#endif
#ifndef PHW_U_ECELLS
#define PHW_U_ECELLS 2
#endif
#ifndef PHW_U_EROWS
#define PHW_U_EROWS 2
#endif

#ifndef PHW_NT_P
#define PHW_NT_P 256
#endif
#ifndef PHW_NT_Q
#define PHW_NT_Q 384
#endif
#ifndef PHW_MINB_P
#define PHW_MINB_P 4
#endif
#ifndef PHW_MINB_Q
#define PHW_MINB_Q 2
#endif
#ifndef PHW_NT_ELL
#define PHW_NT_ELL 256
#endif
#ifndef PHW_MINB_ELL
#define PHW_MINB_ELL 4
#endif
#ifndef PHW_NT_ELLS
#define PHW_NT_ELLS 512
#endif
#ifndef PHW_MINB_ELLS
#define PHW_MINB_ELLS 2
#endif

namespace phw {

#ifndef PHW_NT_ROWS
#define PHW_NT_ROWS 512
#endif
#ifndef PHW_MINB_ROWS
#define PHW_MINB_ROWS 2
#endif
constexpr int NT_ROWS = PHW_NT_ROWS;   // threads per CTA of the row kernels
constexpr int NT_COOP_P = PHW_NT_P;  // threads per CTA of the persistent M_p solve: 256 x 4 CTAs/SM (64 registers; 6 CTAs x 40 registers spill into L2)
constexpr int NT_COOP_Q = PHW_NT_Q;  // M_q solve: 2 CTAs/SM (shared-memory bound) x 384 threads leaves 85 registers per thread for batched loads

struct SolveCtl {
    int k, done, cur, pad;
    double rho, rr, bb, racc, bacc, maxres;
    long long iters, solves;
};

// device-resident control block: everything the step loop needs between kernels lives here, so the
// host only enqueues work and never reads anything back (wave_2D.py:787-982 loop state).
struct Ctl {
    double t, bE, inpE, H_em, H_wave, disp, pleft, u_now, h1, pad0;
    double x[3], xm[3], xt2, xi1;
    double flux[4];
    double red[8];
    long long step;
    long long nonconv;
    unsigned long long comm_seq;     // number of inter-GPU exchanges done (identical on every rank)
    int comm_dead, comm_pad;
    double xch[2][4];                // globally reduced values of the current exchange
    SolveCtl sc[3];
    unsigned int counters[8];
    unsigned int gbar[4];            // arrival counter of grid_barrier (zeroed by the host before each cooperative launch)
    double reshist[3][48];           // ||r_k||^2 / ||b||^2 of the first 48 iterations of the most recent solve (diagnostics)
    // PHW_XCH_DEBUG bit 6 (measurement only): globaltimer stamps of the first 32 iterations of the most recent pipelined M_p / M_q
    // solve of a partitioned run, taken by CTA 0: [0] its sweep is done, [1] grid barrier passed, [2] local norms reduced,
    // [3] norms of all ranks received; [4] = latest arrival of any CTA at the grid barrier (atomicMax)
    unsigned long long xts[2][32][5];
};
#ifdef PHW_HOST_EMU
__device__ __forceinline__ unsigned long long phw_globaltimer() { return 0ull; }
#else
__device__ __forceinline__ unsigned long long phw_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

struct DevMesh {
    int npatch, nv, ne;
    const PatchHdr* hdr;
    const uint2* recv;   // vertex half of the cell record: lv0|lv1<<16, lv2|flags<<16
    const uint2* rece;   // edge half: le0|le1<<16, le2
    const double* area;
    const double* g0; const double* g1; const double* g2;
    const int* ghost_v; const int* ghost_e;
    const uint32_t* eadj;
    const uint32_t* vslice_off;
    const uint16_t* vadj;
    // assembled P1 mass rows (sliced ELL, layout.hpp): neighbour local ids + weights (|K1| + |K2|) / 12
    const uint32_t* ell_ioff; const uint32_t* ell_woff;
    const uint16_t* ell_ids; const double* ell_w;
};

struct TriBuf { double* b[4]; };   // rotating Chebyshev iterates (the one-sweep kernels use the first three)

// ---- multi-GPU: peer-memory halo exchange + all-reduce fused into the persistent solve ----
constexpr int MAX_RANKS = 8;
struct Mailbox {                       // lives at the start of every rank's exchange arena (peer-mapped)
    unsigned long long flag[MAX_RANKS];        // flag[r] = last exchange sequence number rank r has published here
    double part[2][MAX_RANKS][4];              // per-rank partial sums, double-buffered by sequence parity
};
struct PeerComm {
    int rank, world, n_nb, pad;
    int nb_rank[MAX_RANKS];
    double* nb_tp[MAX_RANKS][3];               // neighbour's Chebyshev buffers (peer pointers)
    double* nb_tq[MAX_RANKS][3];
    const int* send_v_loc[MAX_RANKS]; const int* send_v_rem[MAX_RANKS]; int n_send_v[MAX_RANKS];
    const int* send_e_loc[MAX_RANKS]; const int* send_e_rem[MAX_RANKS]; int n_send_e[MAX_RANKS];
    Mailbox* mbox[MAX_RANKS];                  // every rank's mailbox (own one included)
    long long timeout_cycles;
    // exchange protocol of the pipelined solves (round 2): the send lists bucketed by owner patch, so that a CTA pushes the
    // interface rows of a patch right after computing them (overlapped with the rest of the sweep) instead of after the grid barrier.
    // psend_off[f][P] .. psend_off[f][P + 1] = entries of patch P for field f (0 vertices, 1 edges); entry = (local row id,
    // neighbour slot, position in the neighbour's vectors, 0)
    const int* psend_off[2];
    const int4* psend_ent[2];
    int xch_dbg;                               // A/B knobs of the exchange (PHW_XCH_DEBUG; measurement only, see peer_finish_iter)
};

// PHW_HOST_EMU: the kernels of this directory also compile for the host-side SPMD emulator of tests/emu (fibers for threads,
// OS threads for CTAs, deferred asynchronous copies) - a CPU check of indexing, staging protocol and arithmetic.  The emulator
// replaces the few inline-PTX helpers; nothing else differs.
#ifdef PHW_HOST_EMU
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }
#else
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
#endif

// all-reduce (sum, rank order => identical bits on every rank) of n <= 4 doubles through the peer mailboxes.
// Called by ONE thread per rank.  Every rank calls it the same number of times (seq is the call count).
__device__ inline void peer_allreduce(const PeerComm& pc, Ctl* ctl, double* vals, int n) {
    const unsigned long long seq = ctl->comm_seq + 1;
    ctl->comm_seq = seq;
    if (pc.world <= 1) return;
    const int par = (int)(seq & 1ull);
    for (int r = 0; r < pc.world; ++r)
        for (int i = 0; i < n; ++i) pc.mbox[r]->part[par][pc.rank][i] = vals[i];
    __threadfence_system();
    for (int r = 0; r < pc.world; ++r) *((volatile unsigned long long*)&pc.mbox[r]->flag[pc.rank]) = seq;
    Mailbox* mine = pc.mbox[pc.rank];
    if (!ctl->comm_dead) {
        const long long t0 = clock64();
        for (int r = 0; r < pc.world; ++r) {
            while (ld_volatile_u64(&mine->flag[r]) < seq) {
                if (clock64() - t0 > pc.timeout_cycles) { ctl->comm_dead = 1; break; }
            }
            if (ctl->comm_dead) break;
        }
    }
    __threadfence_system();
    for (int i = 0; i < n; ++i) {
        double sacc = 0.0;
        for (int r = 0; r < pc.world; ++r) sacc += *((volatile double*)&mine->part[par][r][i]);
        vals[i] = sacc;
    }
}

// small single-block kernel: all-reduce ctl->red[first..first+n) or ctl->flux[first..] across ranks
__global__ void allreduce_k(PeerComm pc, Ctl* ctl, int what, int first, int n) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double* v = (what == 0) ? &ctl->red[first] : &ctl->flux[first];
    double tmp[4];
    for (int i = 0; i < n; ++i) tmp[i] = v[i];
    peer_allreduce(pc, ctl, tmp, n);
    for (int i = 0; i < n; ++i) v[i] = tmp[i];
}

enum { EPI_STORE = 0, EPI_CHEB = 1, EPI_DOT = 2 };
enum { MODE_OP = 0, MODE_DIAG = 1 };

struct RowArgs {
    DevMesh m;
    const double* xp;        // vertex input (nullptr -> taken from tp.b[cur])
    const double* xq;        // edge input
    TriBuf tp, tq;           // Chebyshev buffers of the p and q unknowns
    double cM, cD, cB;
    // STORE
    double* y;
    // CHEB
    const double* b;
    const double* pinv;
    const double* mdiag;     // pure P1 mass diagonal sum_K |K|/6 (nullptr: use 1/pinv)
    const double* gscale;    // batched runs: per-group factor applied to cD (nullptr: 1)
    double* wpart;           // batched runs, DOT: per-(patch, warp) partial sums [npatch][NT/32] (nullptr: block reduction)
    double xscale;           // > 0 (strong Dirichlet rows): stop test ||r|| <= tol (||b|| + sqrt(xscale) ||x||_M), needs mdiag
    double cheb_d, cheb_c;   // centre and focal distance of the spectrum ellipse
    double tol2, weight;
    int which;               // SolveCtl index
    int use_cheb_input;      // 1: the row's own unknown is read from the TriBuf[cur]
    int coupled_from_tribuf; // 1: the *other* field is also read from its TriBuf[cur] (block solve)
    int finalize;            // 1: this launch closes the iteration (advance k, rotate, test convergence)
    int max_iter;
    // DOT
    int red_slot;            // ctl->red[red_slot] receives the sum (scaled by dot_scale)
    double dot_scale;
    int dot_accumulate;
    double* partials;        // [npatch*2] scratch
    Ctl* ctl;
    int counter_id;
    int gbar;                // 1: cooperative kernels synchronise the grid with grid_barrier (ctl->gbar[0] zeroed before the launch)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block reduction of two values; result valid in thread 0
__device__ __forceinline__ void block_sum2(double& a, double& b, double* red /* >= 2*32 doubles */) {
    a = warp_sum(a); b = warp_sum(b);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { red[w] = a; red[32 + w] = b; }
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        a = (l < nw) ? red[l] : 0.0;
        b = (l < nw) ? red[32 + l] : 0.0;
        a = warp_sum(a); b = warp_sum(b);
    }
}

// Chebyshev coefficients for iteration k given rho_{k-1} (Saad, Iterative Methods, Alg. 12.1;
// ellipse version: Manteuffel 1977).  x_{k+1} = x_k + c1 (x_k - x_{k-1}) + c2 P^-1 r_k.
__device__ __forceinline__ void cheb_coeffs(int k, double rho_prev, double d, double c, double& c1, double& c2, double& rho_k) {
    if (k == 0 || c <= 0.0) { c1 = 0.0; c2 = 1.0 / d; rho_k = c / d; return; }
    double sigma1 = d / c;
    rho_k = 1.0 / (2.0 * sigma1 - rho_prev);
    c1 = rho_k * rho_prev;
    c2 = 2.0 * rho_k / c;
}


// resolved per-launch (or per-iteration) pointers and Chebyshev coefficients
struct PassIO {
    const double* xp;      // vertex input
    const double* xq;      // edge input
    const double* xprev;   // previous iterate of the row's own unknown (CHEB, k > 0)
    double* xnew;          // next iterate (CHEB)
    double c1, c2;
    int use_prev;
};

__device__ __forceinline__ double* tb_sel(const TriBuf& t, int i) { return i == 0 ? t.b[0] : (i == 1 ? t.b[1] : t.b[2]); }

// patch headers are double-buffered in shared memory: the header of the CTA's next patch is fetched while the current
// patch is processed (no exposed header load at the top of a patch)
__device__ __forceinline__ void load_hdr_smem(PatchHdr* dst, const PatchHdr* src) {
    if (threadIdx.x < 20) reinterpret_cast<int*>(dst)[threadIdx.x] = __ldg(reinterpret_cast<const int*>(src) + threadIdx.x);
}

// asynchronous global -> shared copies (LDGSTS): the data of a whole patch is put in flight at once without holding
// registers; completion is awaited with cp_async_wait_all() before the block-wide barrier
#ifdef PHW_HOST_EMU
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) { phw_emu::ldgsts(smem_dst, gmem_src, 16); }
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) { phw_emu::ldgsts(smem_dst, gmem_src, 8); }
__device__ __forceinline__ void cp_async_wait_all() { phw_emu::ldgsts_wait_all(); }
#else
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
#endif

// ------------------------------------------------------------------------------------------------
// vertex rows: all patches of this block
// ------------------------------------------------------------------------------------------------
template <int EPI, int MODE, bool HAS_M, bool HAS_D, bool CAS, int NT>
__device__ __forceinline__ void vertex_pass(const RowArgs& a, const PassIO& io, double* smem, double& acc0, double& acc1) {
    const DevMesh& m = a.m;
    // pure P1 mass rows: (M_p x)_v = sum_K (|K|/12) sum(x_K) + x_v * diag_v / 2, diag_v = sum_K |K|/6 = 1/pinv_v,
    // so one value per cell suffices (a.pinv must be set)
    constexpr bool ONEVAL = HAS_M && !HAS_D && !CAS && MODE == MODE_OP;
    __shared__ PatchHdr s_hdr[2];
    if (blockIdx.x < m.npatch) load_hdr_smem(&s_hdr[0], &m.hdr[blockIdx.x]);
    __syncthreads();
    int it = 0;
    for (int P = blockIdx.x; P < m.npatch; P += gridDim.x, ++it) {
        const PatchHdr& h = s_hdr[it & 1];
        const int Pn = P + gridDim.x;
        if (Pn < m.npatch) load_hdr_smem(&s_hdr[(it + 1) & 1], &m.hdr[Pn]);
        const double cD = (HAS_D && a.gscale) ? a.cD * a.gscale[h.group] : a.cD;
        const int nlv = h.v_cnt + h.gv_cnt, nle = h.e_cnt + h.ge_cnt;
        const int nsl = (h.v_cnt + 31) >> 5;
        uint32_t* s_off = reinterpret_cast<uint32_t*>(smem);   // [nsl + 1] slice offsets of the sliced-ELL adjacency
        double* sp = smem + ((nsl + 2) >> 1);      // [nlv]
        double* sq = sp + ((nlv + 1) & ~1);        // [nle] (only if HAS_D)
        double* sc3 = HAS_D ? sq + ((nle + 1) & ~1) : sq;   // [CW*n_vcells]
        // ---- stage 1: local inputs
        for (int i = threadIdx.x; i <= nsl; i += NT) s_off[i] = __ldg(&m.vslice_off[h.vs_off + i]);
        if (MODE == MODE_OP && (HAS_M || CAS || EPI != EPI_STORE)) {
            for (int i = threadIdx.x; i < h.v_cnt; i += NT) sp[i] = io.xp[h.v_start + i];
            for (int i = threadIdx.x; i < h.gv_cnt; i += NT) sp[h.v_cnt + i] = io.xp[m.ghost_v[h.gv_off + i]];
        }
        if (HAS_D && MODE == MODE_OP) {
            for (int i = threadIdx.x; i < h.e_cnt; i += NT) sq[i] = io.xq[h.e_start + i];
            for (int i = threadIdx.x; i < h.ge_cnt; i += NT) sq[h.e_cnt + i] = io.xq[m.ghost_e[h.ge_off + i]];
        }
        __syncthreads();
        // ---- stage 2: per-cell contributions to its three vertices
        if (ONEVAL) {
            // batched: issue the loads of U cells before consuming them (memory-level parallelism)
            constexpr int U = PHW_U_VCELLS;
            for (int base = threadIdx.x; base < h.n_vcells; base += U * NT) {
                uint2 rvb[U]; double Ab[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int lc = base + u * NT;
                    if (lc < h.n_vcells) { rvb[u] = __ldg(&m.recv[h.cell_off + lc]); Ab[u] = __ldg(&m.area[h.cell_off + lc]); }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int lc = base + u * NT;
                    if (lc < h.n_vcells)
                        sc3[lc] = a.cM * Ab[u] * (1.0 / 12.0) * (sp[rvb[u].x & 0xFFFFu] + sp[rvb[u].x >> 16] + sp[rvb[u].y & 0xFFFFu]);
                }
            }
        }
        for (int lc = threadIdx.x; lc < (ONEVAL ? 0 : h.n_vcells); lc += NT) {
            const uint2 rv = __ldg(&m.recv[h.cell_off + lc]);
            double v0 = 0.0, v1 = 0.0, v2 = 0.0;
            if (HAS_M || CAS) {
                const double A = __ldg(&m.area[h.cell_off + lc]);
                if (MODE == MODE_DIAG) {
                    v0 = v1 = v2 = a.cM * A * (1.0 / 6.0);
                    if (CAS) {
                        const uint32_t fl = rv.y >> 16;
                        if (fl & 1u) { const double d = a.cB * sqrt(48.0 * A * __ldg(&m.g0[h.cell_off + lc])) * (1.0 / 3.0); v1 += d; v2 += d; }
                        if (fl & 2u) { const double d = a.cB * sqrt(48.0 * A * __ldg(&m.g1[h.cell_off + lc])) * (1.0 / 3.0); v0 += d; v2 += d; }
                        if (fl & 4u) { const double d = a.cB * sqrt(48.0 * A * __ldg(&m.g2[h.cell_off + lc])) * (1.0 / 3.0); v0 += d; v1 += d; }
                    }
                } else {
                    const double x0 = sp[rv.x & 0xFFFFu], x1 = sp[rv.x >> 16], x2 = sp[rv.y & 0xFFFFu];
                    if (HAS_M) {
                        const double w = a.cM * A * (1.0 / 12.0), s = x0 + x1 + x2;
                        v0 = w * (s + x0); v1 = w * (s + x1); v2 = w * (s + x2);
                    }
                    if (CAS) {
                        const uint32_t fl = rv.y >> 16;
                        if (fl & 7u) {   // impedance top side: P1 edge mass |e|/6 [2 1; 1 2]  (wave_2D.py:330,476)
                            const double gg[3] = {__ldg(&m.g0[h.cell_off + lc]), __ldg(&m.g1[h.cell_off + lc]), __ldg(&m.g2[h.cell_off + lc])};
                            const double xx[3] = {x0, x1, x2};
                            double vv[3] = {0.0, 0.0, 0.0};
#pragma unroll
                            for (int i = 0; i < 3; ++i)
                                if (fl & (1u << i)) {
                                    const double len = sqrt(48.0 * A * gg[i]);
                                    const int j = (i + 1) % 3, kk = (i + 2) % 3;
                                    vv[j] += a.cB * len * (1.0 / 6.0) * (2.0 * xx[j] + xx[kk]);
                                    vv[kk] += a.cB * len * (1.0 / 6.0) * (2.0 * xx[kk] + xx[j]);
                                }
                            v0 += vv[0]; v1 += vv[1]; v2 += vv[2];
                        }
                    }
                }
            }
            if (HAS_D && MODE == MODE_OP) {
                const uint2 re = __ldg(&m.rece[h.cell_off + lc]);
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
                v0 += d0; v1 += d1; v2 += d2;
            }
            sc3[3 * lc] = v0; sc3[3 * lc + 1] = v1; sc3[3 * lc + 2] = v2;
        }
        __syncthreads();
        // ---- stage 3: owned rows (sliced-ELL adjacency, 8 codes per 16-byte load, warp-uniform trip count),
        //      fixed summation order, fused epilogue; streaming operands are requested before the reduction
        const int vround = (h.v_cnt + 31) & ~31;
        const uint4* adj4 = reinterpret_cast<const uint4*>(m.vadj);
        for (int lv = threadIdx.x; lv < vround; lv += NT) {
            const uint32_t o0 = s_off[lv >> 5], o1 = s_off[(lv >> 5) + 1];
            const bool live = lv < h.v_cnt;
            const int gid = h.v_start + (live ? lv : 0);
            double bv = 0.0, pv = 0.0, xpv = 0.0, dgv = 0.0;
            if (EPI == EPI_CHEB && live) {
                bv = a.b[gid]; pv = a.pinv[gid];
                if (io.use_prev) xpv = io.xprev[gid];
                if ((ONEVAL || a.xscale > 0.0) && a.mdiag) dgv = a.mdiag[gid];
            }
            if (EPI != EPI_CHEB && ONEVAL && live) dgv = a.mdiag[gid];
            double val = 0.0;
            const uint32_t nchunk = (o1 - o0) >> 8;            // 32 lanes * 8 codes per chunk
            for (uint32_t k = 0; k < nchunk; ++k) {
                const uint4 cd = __ldg(&adj4[(o0 >> 3) + k * 32 + (lv & 31)]);
                const uint32_t w4[4] = {cd.x, cd.y, cd.z, cd.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t c0 = w4[j] & 0xFFFFu, c1 = w4[j] >> 16;
                    if (c0 != 0xFFFFu) val += ONEVAL ? sc3[c0 >> 2] : sc3[3 * (c0 >> 2) + (c0 & 3u)];
                    if (c1 != 0xFFFFu) val += ONEVAL ? sc3[c1 >> 2] : sc3[3 * (c1 >> 2) + (c1 & 3u)];
                }
            }
            if (!live) continue;
            if (EPI == EPI_STORE) {
                if (ONEVAL) val += 0.5 * a.cM * sp[lv] * dgv;
                a.y[gid] = val;
            } else if (EPI == EPI_CHEB) {
                const double xc = sp[lv];
                if (ONEVAL) val += 0.5 * a.cM * xc * (a.mdiag ? dgv : 1.0 / pv);
                const double r_ = bv - val;
                double xn = xc + io.c2 * pv * r_;
                if (io.use_prev) xn += io.c1 * (xc - xpv);
                io.xnew[gid] = xn;
                acc0 += pv * r_ * r_;
                acc1 += pv * bv * bv + a.xscale * dgv * xc * xc;
            } else {
                if (ONEVAL) val += 0.5 * a.cM * sp[lv] * dgv;
                acc0 += sp[lv] * val;
            }
        }
        if (EPI == EPI_DOT && a.wpart) {   // batched runs: per-(patch, warp) partials, reduced per group afterwards
            const double w0 = warp_sum(acc0);
            if ((threadIdx.x & 31) == 0) a.wpart[(size_t)P * (NT / 32) + (threadIdx.x >> 5)] = w0;
            acc0 = 0.0;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// vertex rows of the pure P1 mass solve in assembled form: (M_p x)_v = d_v x_v + sum_w m_vw x_w, d_v = sum_w m_vw, with the sliced-ELL
// rows of layout.hpp (weights m_vw = (|K1| + |K2|) / 12).  Same rows, same bounds and the same DRAM bytes as the cell-based
// pass (ids 16 B + weights 8 B x degree per vertex instead of cell records + areas + adjacency), but one shared-memory
// stage instead of two: no per-cell buffer, one barrier less, about half the shared-memory wavefronts.
// ------------------------------------------------------------------------------------------------
// COH: the right-hand side is read with ordinary (coherent) loads - required inside the persistent whole-loop kernel
// (kernels_loop.cuh), where b is rewritten by an earlier phase of the SAME launch, so the read-only path must not be used
template <int NT, bool COH = false>
__device__ __forceinline__ void vertex_ell_pass(const RowArgs& a, const PassIO& io, double* smem, double& acc0, double& acc1) {
    const DevMesh& m = a.m;
    __shared__ PatchHdr s_hdr[2];
    if (blockIdx.x < m.npatch) load_hdr_smem(&s_hdr[0], &m.hdr[blockIdx.x]);
    __syncthreads();
    const uint4* ids4 = reinterpret_cast<const uint4*>(m.ell_ids);
    int it = 0;
    for (int P = blockIdx.x; P < m.npatch; P += gridDim.x, ++it) {
        const PatchHdr& h = s_hdr[it & 1];
        if (P + (int)gridDim.x < m.npatch) load_hdr_smem(&s_hdr[(it + 1) & 1], &m.hdr[P + gridDim.x]);
        const int nsl = (h.v_cnt + 31) >> 5;
        uint32_t* s_io = reinterpret_cast<uint32_t*>(smem);   // [nsl + 1] id offsets
        uint32_t* s_wo = s_io + (nsl + 1);                     // [nsl + 1] weight offsets
        double* sx = smem + (nsl + 2);                         // [v_cnt + gv_cnt]
        for (int i = threadIdx.x; i <= nsl; i += NT) { s_io[i] = __ldg(&m.ell_ioff[h.vs_off + i]); s_wo[i] = __ldg(&m.ell_woff[h.vs_off + i]); }
        for (int i = threadIdx.x; i < h.v_cnt; i += NT) sx[i] = io.xp[h.v_start + i];
        for (int i = threadIdx.x; i < h.gv_cnt; i += NT) sx[h.v_cnt + i] = io.xp[__ldg(&m.ghost_v[h.gv_off + i])];
        __syncthreads();
        for (int lv = threadIdx.x; lv < h.v_cnt; lv += NT) {
            const int sl = lv >> 5, lane = lv & 31;
            const uint32_t io0 = s_io[sl], wo0 = s_wo[sl];
            const int W = (int)((s_wo[sl + 1] - wo0) >> 5);
            const int gid = h.v_start + lv;
            const double bv = COH ? a.b[gid] : __ldg(&a.b[gid]);
            const double xpv = io.use_prev ? io.xprev[gid] : 0.0;
            const double* __restrict__ wrow = m.ell_w + wo0 + lane;
            double val = 0.0, dsum = 0.0;
            for (int c = 0; 8 * c < W; ++c) {
                const uint4 cd = __ldg(&ids4[(io0 >> 3) + c * 32 + lane]);
                const uint32_t w4[4] = {cd.x, cd.y, cd.z, cd.w};
                double wv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) wv[j] = (8 * c + j < W) ? __ldg(&wrow[(8 * c + j) * 32]) : 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    val += wv[2 * j] * sx[w4[j] & 0xFFFFu];
                    val += wv[2 * j + 1] * sx[w4[j] >> 16];
                    dsum += wv[2 * j] + wv[2 * j + 1];
                }
            }
            // the P1 mass diagonal is the sum of the row's off-diagonal entries (sum_K |K|/6 = sum_w (|K1|+|K2|)/12), so
            // neither the diagonal nor the Jacobi weight has to be read
            const double xc = sx[lv];
            const double pv = 1.0 / dsum;
            val += dsum * xc;
            const double r_ = bv - val;
            double xn = xc + io.c2 * pv * r_;
            if (io.use_prev) xn += io.c1 * (xc - xpv);
            io.xnew[gid] = xn;
            acc0 += pv * r_ * r_;
            acc1 += pv * bv * bv;
        }
        __syncthreads();
    }
}

// The same rows with every stream of the patch staged in shared memory by asynchronous copies: ids and weights (two
// contiguous blocks), b and x_{k-1} (owned range) are requested for the WHOLE patch before anything is consumed, so a
// patch costs one memory latency instead of one per 256-row round; rows then run from shared memory.
template <int NT>
__device__ __forceinline__ void vertex_ell_pass_staged(const RowArgs& a, const PassIO& io, double* smem, double& acc0, double& acc1) {
    const DevMesh& m = a.m;
    __shared__ PatchHdr s_hdr[2];
    if (blockIdx.x < m.npatch) load_hdr_smem(&s_hdr[0], &m.hdr[blockIdx.x]);
    __syncthreads();
    int it = 0;
    for (int P = blockIdx.x; P < m.npatch; P += gridDim.x, ++it) {
        const PatchHdr& h = s_hdr[it & 1];
        if (P + (int)gridDim.x < m.npatch) load_hdr_smem(&s_hdr[(it + 1) & 1], &m.hdr[P + gridDim.x]);
        const int nsl = (h.v_cnt + 31) >> 5, nlv = h.v_cnt + h.gv_cnt;
        const uint32_t i0 = (uint32_t)h.ell_i0, in = (uint32_t)h.ell_in, w0 = (uint32_t)h.ell_w0, wn = (uint32_t)h.ell_wn;
        uint32_t* s_io = reinterpret_cast<uint32_t*>(smem);
        uint32_t* s_wo = s_io + (nsl + 1);
        double* sx = smem + ((nsl + 3) & ~1);
        double* sb = sx + ((nlv + 1) & ~1);
        double* sxp = sb + ((h.v_cnt + 1) & ~1);
        double* sw = sxp + ((h.v_cnt + 1) & ~1);
        uint4* sid4 = reinterpret_cast<uint4*>(sw + wn);
        {   // whole-patch requests
            const uint4* gi = reinterpret_cast<const uint4*>(m.ell_ids + i0);
            for (uint32_t i = threadIdx.x; i < (in >> 3); i += NT) cp_async16(&sid4[i], &gi[i]);
            const double2* gw = reinterpret_cast<const double2*>(m.ell_w + w0);
            double2* sw2 = reinterpret_cast<double2*>(sw);
            for (uint32_t i = threadIdx.x; i < (wn >> 1); i += NT) cp_async16(&sw2[i], &gw[i]);
            for (int i = threadIdx.x; i < h.v_cnt; i += NT) cp_async8(&sb[i], &a.b[h.v_start + i]);
            if (io.use_prev)
                for (int i = threadIdx.x; i < h.v_cnt; i += NT) cp_async8(&sxp[i], &io.xprev[h.v_start + i]);
        }
        for (int i = threadIdx.x; i <= nsl; i += NT) { s_io[i] = __ldg(&m.ell_ioff[h.vs_off + i]) - i0; s_wo[i] = __ldg(&m.ell_woff[h.vs_off + i]) - w0; }
        for (int i = threadIdx.x; i < h.v_cnt; i += NT) sx[i] = io.xp[h.v_start + i];
        for (int i = threadIdx.x; i < h.gv_cnt; i += NT) sx[h.v_cnt + i] = io.xp[__ldg(&m.ghost_v[h.gv_off + i])];
        cp_async_wait_all();
        __syncthreads();
        for (int lv = threadIdx.x; lv < h.v_cnt; lv += NT) {
            const int sl = lv >> 5, lane = lv & 31;
            const uint32_t io0 = s_io[sl], wo0 = s_wo[sl];
            const int W = (int)((s_wo[sl + 1] - wo0) >> 5);
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
            const double xc = sx[lv], bv = sb[lv];
            const double pv = 1.0 / dsum;
            val += dsum * xc;
            const double r_ = bv - val;
            double xn = xc + io.c2 * pv * r_;
            if (io.use_prev) xn += io.c1 * (xc - sxp[lv]);
            io.xnew[h.v_start + lv] = xn;
            acc0 += pv * r_ * r_;
            acc1 += pv * bv * bv;
        }
        __syncthreads();
    }
}

// diagnostics: reads exactly the streams of one ELL M_p iteration patch by patch (ids, weights, b, x_prev, x owned) and
// writes one value per owned vertex - no shared memory, no barriers, no dependent loads: the bandwidth the memory system
// delivers for this layout and access order, to compare the solve kernels with (phw_stream_probe)
template <int NT>
__global__ void __launch_bounds__(NT) stream_probe_k(DevMesh m, const double* __restrict__ b, const double* __restrict__ xprev,
                                                       const double* __restrict__ x, double* __restrict__ out) {
    for (int P = blockIdx.x; P < m.npatch; P += gridDim.x) {
        const PatchHdr h = m.hdr[P];
        double acc = 0.0;
        const uint4* gi = reinterpret_cast<const uint4*>(m.ell_ids + (uint32_t)h.ell_i0);
        for (uint32_t i = threadIdx.x; i < ((uint32_t)h.ell_in >> 3); i += NT) { const uint4 v = __ldg(&gi[i]); acc += (double)(v.x ^ v.y ^ v.z ^ v.w); }
        const double2* gw = reinterpret_cast<const double2*>(m.ell_w + (uint32_t)h.ell_w0);
        for (uint32_t i = threadIdx.x; i < ((uint32_t)h.ell_wn >> 1); i += NT) { const double2 v = __ldg(&gw[i]); acc += v.x + v.y; }
        for (int i = threadIdx.x; i < h.v_cnt; i += NT) {
            const int g = h.v_start + i;
            out[g] = acc + __ldg(&b[g]) + __ldg(&xprev[g]) + __ldg(&x[g]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// edge rows: all patches of this block
// ------------------------------------------------------------------------------------------------
template <int EPI, int MODE, bool HAS_M, bool HAS_D, bool CAS, int NT, bool COH = false>
__device__ __forceinline__ void edge_pass(const RowArgs& a, const PassIO& io, double* smem, double& acc0, double& acc1) {
    const DevMesh& m = a.m;
    __shared__ PatchHdr s_hdr[2];
    if (blockIdx.x < m.npatch) load_hdr_smem(&s_hdr[0], &m.hdr[blockIdx.x]);
    __syncthreads();
    int it = 0;
    for (int P = blockIdx.x; P < m.npatch; P += gridDim.x, ++it) {
        const PatchHdr& h = s_hdr[it & 1];
        const int Pn = P + gridDim.x;
        if (Pn < m.npatch) load_hdr_smem(&s_hdr[(it + 1) & 1], &m.hdr[Pn]);
        const double cD = (HAS_D && a.gscale) ? a.cD * a.gscale[h.group] : a.cD;
        const int nlv = h.v_cnt + h.gv_cnt, nle = h.e_cnt + h.ge_cnt;
        double* sq = smem;                          // [nle]
        double* sp = sq + ((nle + 1) & ~1);         // [nlv] (only if HAS_D)
        double* sc3 = HAS_D ? sp + ((nlv + 1) & ~1) : sp;   // [3*n_ecells]
        if (MODE == MODE_OP && (HAS_M || CAS || EPI != EPI_STORE)) {
            for (int i = threadIdx.x; i < h.e_cnt; i += NT) sq[i] = io.xq[h.e_start + i];
            for (int i = threadIdx.x; i < h.ge_cnt; i += NT) sq[h.e_cnt + i] = io.xq[m.ghost_e[h.ge_off + i]];
        }
        if (HAS_D && MODE == MODE_OP) {
            for (int i = threadIdx.x; i < h.v_cnt; i += NT) sp[i] = io.xp[h.v_start + i];
            for (int i = threadIdx.x; i < h.gv_cnt; i += NT) sp[h.v_cnt + i] = io.xp[m.ghost_v[h.gv_off + i]];
        }
        __syncthreads();
        constexpr bool MONLY = HAS_M && !HAS_D && !CAS && MODE == MODE_OP && EPI == EPI_CHEB;
        if (MONLY) {
            // pure RT1 mass rows: issue the loads of U cells before consuming them (memory-level parallelism)
            constexpr int U = PHW_U_ECELLS;
            for (int base = threadIdx.x; base < h.n_ecells; base += U * NT) {
                uint2 reb[U]; double gb[U][3];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int lc = base + u * NT;
                    if (lc < h.n_ecells) {
                        reb[u] = __ldg(&m.rece[h.cell_off + lc]);
                        gb[u][0] = __ldg(&m.g0[h.cell_off + lc]); gb[u][1] = __ldg(&m.g1[h.cell_off + lc]); gb[u][2] = __ldg(&m.g2[h.cell_off + lc]);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int lc = base + u * NT;
                    if (lc < h.n_ecells) {
                        const uint32_t e0 = reb[u].x & 0xFFFFu, e1 = reb[u].x >> 16, e2 = reb[u].y & 0xFFFFu;
                        const double g0 = gb[u][0], g1 = gb[u][1], g2 = gb[u][2], S = g0 + g1 + g2;
                        const double q0 = sq[e0 & 0x3FFFu], q1 = sq[e1 & 0x3FFFu], q2 = sq[e2 & 0x3FFFu];
                        const double s0 = (e0 & 0x8000u) ? -1.0 : 1.0, s1 = (e1 & 0x8000u) ? -1.0 : 1.0, s2 = (e2 & 0x8000u) ? -1.0 : 1.0;
                        const double t0 = s0 * q0, t1 = s1 * q1, t2 = s2 * q2;
                        const double ST = S * (t0 + t1 + t2);
                        sc3[3 * lc] = a.cM * s0 * (ST + 2.0 * S * t0 - 4.0 * (g0 * t0 + g2 * t1 + g1 * t2));
                        sc3[3 * lc + 1] = a.cM * s1 * (ST + 2.0 * S * t1 - 4.0 * (g2 * t0 + g1 * t1 + g0 * t2));
                        sc3[3 * lc + 2] = a.cM * s2 * (ST + 2.0 * S * t2 - 4.0 * (g1 * t0 + g0 * t1 + g2 * t2));
                    }
                }
            }
        }
        for (int lc = threadIdx.x; lc < (MONLY ? 0 : h.n_ecells); lc += NT) {
            const uint2 re = __ldg(&m.rece[h.cell_off + lc]);
            uint2 rv = make_uint2(0u, 0u);
            if (HAS_D || CAS) rv = __ldg(&m.recv[h.cell_off + lc]);
            const uint32_t e0 = re.x & 0xFFFFu, e1 = re.x >> 16, e2 = re.y & 0xFFFFu;
            const double s0 = (e0 & 0x8000u) ? -1.0 : 1.0, s1 = (e1 & 0x8000u) ? -1.0 : 1.0, s2 = (e2 & 0x8000u) ? -1.0 : 1.0;
            double v0 = 0.0, v1 = 0.0, v2 = 0.0;
            if (HAS_M || CAS) {
                const double g0 = __ldg(&m.g0[h.cell_off + lc]), g1 = __ldg(&m.g1[h.cell_off + lc]), g2 = __ldg(&m.g2[h.cell_off + lc]);
                const double S = g0 + g1 + g2;
                if (MODE == MODE_DIAG) {
                    v0 = a.cM * (3.0 * S - 4.0 * g0); v1 = a.cM * (3.0 * S - 4.0 * g1); v2 = a.cM * (3.0 * S - 4.0 * g2);
                    if (CAS) {
                        const uint32_t fl = (rv.y >> 16) >> 3;
                        if (fl & 7u) {
                            const double A = __ldg(&m.area[h.cell_off + lc]);
                            if (fl & 1u) v0 += a.cB * rsqrt(48.0 * A * g0);
                            if (fl & 2u) v1 += a.cB * rsqrt(48.0 * A * g1);
                            if (fl & 4u) v2 += a.cB * rsqrt(48.0 * A * g2);
                        }
                    }
                } else {
                    const double q0 = sq[e0 & 0x3FFFu], q1 = sq[e1 & 0x3FFFu], q2 = sq[e2 & 0x3FFFu];
                    if (HAS_M) {
                        // RT1 element mass from the metric: m_ii = 3S - 4 g_i, m_ij = S - 4 g_k (unsigned basis)
                        const double t0 = s0 * q0, t1 = s1 * q1, t2 = s2 * q2;
                        const double ST = S * (t0 + t1 + t2);
                        v0 = a.cM * s0 * (ST + 2.0 * S * t0 - 4.0 * (g0 * t0 + g2 * t1 + g1 * t2));
                        v1 = a.cM * s1 * (ST + 2.0 * S * t1 - 4.0 * (g2 * t0 + g1 * t1 + g0 * t2));
                        v2 = a.cM * s2 * (ST + 2.0 * S * t2 - 4.0 * (g1 * t0 + g0 * t1 + g2 * t2));
                    }
                    if (CAS) {
                        const uint32_t fl = (rv.y >> 16) >> 3;
                        if (fl & 7u) {   // impedance right side: <phi_e.n, phi_e.n> = 1/|e|   (wave_2D.py:350,482)
                            const double A = __ldg(&m.area[h.cell_off + lc]);
                            if (fl & 1u) v0 += a.cB * q0 * rsqrt(48.0 * A * g0);
                            if (fl & 2u) v1 += a.cB * q1 * rsqrt(48.0 * A * g1);
                            if (fl & 4u) v2 += a.cB * q2 * rsqrt(48.0 * A * g2);
                        }
                    }
                }
            }
            if (HAS_D && MODE == MODE_OP) {
                const double p0 = sp[rv.x & 0xFFFFu], p1 = sp[rv.x >> 16], p2 = sp[rv.y & 0xFFFFu];
                const double third = (p0 + p1 + p2) * (1.0 / 3.0);
                double d0 = third, d1 = third, d2 = third;
                if (e0 & 0x4000u) d0 -= 0.5 * (p1 + p2);
                if (e1 & 0x4000u) d1 -= 0.5 * (p0 + p2);
                if (e2 & 0x4000u) d2 -= 0.5 * (p0 + p1);
                v0 += cD * s0 * d0; v1 += cD * s1 * d1; v2 += cD * s2 * d2;
            }
            sc3[3 * lc] = v0; sc3[3 * lc + 1] = v1; sc3[3 * lc + 2] = v2;
        }
        __syncthreads();
        if (MONLY) {
            constexpr int U3 = PHW_U_EROWS;
            for (int base = threadIdx.x; base < h.e_cnt; base += U3 * NT) {
                uint32_t adjb[U3]; double bvb[U3], pvb[U3], xpb[U3];
#pragma unroll
                for (int u = 0; u < U3; ++u) {
                    const int le = base + u * NT;
                    if (le < h.e_cnt) {
                        const int gid = h.e_start + le;
                        adjb[u] = __ldg(&m.eadj[gid]); bvb[u] = COH ? a.b[gid] : __ldg(&a.b[gid]); pvb[u] = __ldg(&a.pinv[gid]);
                        xpb[u] = io.use_prev ? io.xprev[gid] : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < U3; ++u) {
                    const int le = base + u * NT;
                    if (le < h.e_cnt) {
                        const int gid = h.e_start + le;
                        const uint32_t ca = adjb[u] & 0xFFFFu, cb = adjb[u] >> 16;
                        double val = sc3[3 * (ca >> 2) + (ca & 3u)];
                        if (cb != 0xFFFFu) val += sc3[3 * (cb >> 2) + (cb & 3u)];
                        const double xc = sq[le];
                        const double r_ = bvb[u] - val;
                        double xn = xc + io.c2 * pvb[u] * r_;
                        if (io.use_prev) xn += io.c1 * (xc - xpb[u]);
                        io.xnew[gid] = xn;
                        acc0 += pvb[u] * r_ * r_;
                        acc1 += pvb[u] * bvb[u] * bvb[u];
                    }
                }
            }
        }
        for (int le = threadIdx.x; le < (MONLY ? 0 : h.e_cnt); le += NT) {
            const int gid = h.e_start + le;
            const uint32_t adj = m.eadj[gid];
            double bv = 0.0, pv = 0.0, xpv = 0.0;
            if (EPI == EPI_CHEB) { bv = a.b[gid]; pv = a.pinv[gid]; if (io.use_prev) xpv = io.xprev[gid]; }
            const uint32_t ca = adj & 0xFFFFu, cb = adj >> 16;
            double val = sc3[3 * (ca >> 2) + (ca & 3u)];
            if (cb != 0xFFFFu) val += sc3[3 * (cb >> 2) + (cb & 3u)];
            if (EPI == EPI_STORE) {
                a.y[gid] = val;
            } else if (EPI == EPI_CHEB) {
                const double xc = sq[le];
                const double r_ = bv - val;
                double xn = xc + io.c2 * pv * r_;
                if (io.use_prev) xn += io.c1 * (xc - xpv);
                io.xnew[gid] = xn;
                acc0 += pv * r_ * r_;
                acc1 += pv * bv * bv;
            } else {
                acc0 += sq[le] * val;
            }
        }
        if (EPI == EPI_DOT && a.wpart) {
            const double w0 = warp_sum(acc0);
            if ((threadIdx.x & 31) == 0) a.wpart[(size_t)P * (NT / 32) + (threadIdx.x >> 5)] = w0;
            acc0 = 0.0;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// last-block epilogue of the multi-launch path: deterministic sum of the per-block partials and the
// iteration control of the Chebyshev solve
// ------------------------------------------------------------------------------------------------
template <int EPI>
__device__ void finish_launch(const RowArgs& a, double acc0, double acc1, double* red) {
    __shared__ bool is_last;
    block_sum2(acc0, acc1, red);
    if (threadIdx.x == 0) {
        a.partials[2 * blockIdx.x] = acc0; a.partials[2 * blockIdx.x + 1] = acc1;
        __threadfence();
        unsigned int prev = atomicAdd(&a.ctl->counters[a.counter_id], 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double s0 = 0.0, s1 = 0.0;
    for (int p = threadIdx.x; p < (int)gridDim.x; p += blockDim.x) {
        s0 += __ldcg(&a.partials[2 * p]);
        s1 += __ldcg(&a.partials[2 * p + 1]);
    }
    block_sum2(s0, s1, red);
    if (threadIdx.x == 0) {
        Ctl* ctl = a.ctl;
        ctl->counters[a.counter_id] = 0;
        if (EPI == EPI_DOT) {
            double v = s0 * a.dot_scale;
            ctl->red[a.red_slot] = a.dot_accumulate ? ctl->red[a.red_slot] + v : v;
        } else if (EPI == EPI_CHEB) {
            SolveCtl& sc = ctl->sc[a.which];
            double rr = sc.racc + s0 * a.weight;
            double bb = sc.bacc + s1 * a.weight;
            if (!a.finalize) { sc.racc = rr; sc.bacc = bb; }
            else {
                sc.racc = 0.0; sc.bacc = 0.0;
                sc.rr = rr; sc.bb = bb;
                double c1, c2, rho_k;
                cheb_coeffs(sc.k, sc.rho, a.cheb_d, a.cheb_c, c1, c2, rho_k);
                const bool conv = rr <= a.tol2 * bb;
                sc.k += 1;
                sc.iters += 1;
                if (!conv) { sc.rho = rho_k; sc.cur = (sc.cur + 1) % 3; }   // on convergence keep the verified iterate x_k
                if (conv || sc.k >= a.max_iter) {
                    sc.done = 1;
                    if (!conv) ctl->nonconv += 1;
                    double rel = (bb > 0.0) ? sqrt(rr / bb) : 0.0;
                    if (rel > sc.maxres) sc.maxres = rel;
                }
            }
        }
    }
}

template <int EPI>
__device__ __forceinline__ bool resolve_io(const RowArgs& a, bool vertex, PassIO& io) {
    io.xp = a.xp; io.xq = a.xq; io.xprev = nullptr; io.xnew = nullptr; io.c1 = 0.0; io.c2 = 0.0; io.use_prev = 0;
    if (EPI == EPI_CHEB) {
        const SolveCtl& sc = a.ctl->sc[a.which];
        if (sc.done) return false;
        const int cur = sc.cur, k = sc.k;
        double rk;
        cheb_coeffs(k, sc.rho, a.cheb_d, a.cheb_c, io.c1, io.c2, rk);
        io.use_prev = (k > 0);
        const TriBuf& own = vertex ? a.tp : a.tq;
        const TriBuf& oth = vertex ? a.tq : a.tp;
        if (vertex) { io.xp = tb_sel(own, cur); if (a.coupled_from_tribuf) io.xq = tb_sel(oth, cur); }
        else        { io.xq = tb_sel(own, cur); if (a.coupled_from_tribuf) io.xp = tb_sel(oth, cur); }
        io.xprev = tb_sel(own, (cur + 2) % 3);
        io.xnew = tb_sel(own, (cur + 1) % 3);
    }
    return true;
}

template <int EPI, int MODE, bool HAS_M, bool HAS_D, bool CAS>
__global__ void __launch_bounds__(NT_ROWS, PHW_MINB_ROWS) vertex_rows_k(const __grid_constant__ RowArgs a) {
    extern __shared__ double smem[];
    __shared__ double red[64];
    PassIO io;
    if (!resolve_io<EPI>(a, true, io)) return;
    double acc0 = 0.0, acc1 = 0.0;
    vertex_pass<EPI, MODE, HAS_M, HAS_D, CAS, NT_ROWS>(a, io, smem, acc0, acc1);
    if (EPI != EPI_STORE && !(EPI == EPI_DOT && a.wpart)) finish_launch<EPI>(a, acc0, acc1, red);
}

template <int EPI, int MODE, bool HAS_M, bool HAS_D, bool CAS>
__global__ void __launch_bounds__(NT_ROWS, PHW_MINB_ROWS) edge_rows_k(const __grid_constant__ RowArgs a) {
    extern __shared__ double smem[];
    __shared__ double red[64];
    PassIO io;
    if (!resolve_io<EPI>(a, false, io)) return;
    double acc0 = 0.0, acc1 = 0.0;
    edge_pass<EPI, MODE, HAS_M, HAS_D, CAS, NT_ROWS>(a, io, smem, acc0, acc1);
    if (EPI != EPI_STORE && !(EPI == EPI_DOT && a.wpart)) finish_launch<EPI>(a, acc0, acc1, red);
}

// grid-wide barrier for co-resident grids (cooperative launch): one arrival per CTA on a monotonically increasing counter
// (zeroed by the host before the launch).  Back-to-back cooperative_groups grid syncs cost ~40 us each in these kernels
// (592 CTAs of 256 threads, measured); this barrier ~1 us - which is what makes the inter-GPU exchange, with its two extra
// barriers per iteration, affordable in the lean kernels.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        while (*((volatile unsigned int*)bar) < epoch) { }
        __threadfence();
    }
    __syncthreads();
}

// the generic cooperative kernels use the same barrier when RowArgs::gbar is set (phw_params.flags bit 10), else a
// cooperative-groups grid sync
#define PHW_GRID_SYNC(args_) do { if ((args_).gbar) grid_barrier(&(args_).ctl->gbar[0], epoch); else grid.sync(); } while (0)

// ------------------------------------------------------------------------------------------------
// persistent cooperative solve: the whole Chebyshev iteration loop of one linear solve in ONE launch.
// WHICH = 0: M_p x = b_p, 1: M_q x = b_q, 2: coupled block system (implicit schemes).  One grid-wide
// barrier per iteration; the residual norm is reduced deterministically (per-block partials, summed in
// the same order by every block) so all blocks take the same stop decision without host involvement.
// ------------------------------------------------------------------------------------------------
template <int WHICH, bool CAS, int NT, int MINB, int ELL = 0>
__global__ void __launch_bounds__(NT, MINB) solve_coop_k(const __grid_constant__ RowArgs av, const __grid_constant__ RowArgs ae,
                                                         const __grid_constant__ PeerComm pc) {
    extern __shared__ double smem[];
    __shared__ double red[64];
    __shared__ double bc[2];
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const RowArgs& a0 = (WHICH == 1) ? ae : av;
    SolveCtl& sc = a0.ctl->sc[a0.which];
    int k = 0, cur = sc.cur;
    double rho = 0.0, rr = 0.0, bb = 0.0;
    bool conv = false;
    unsigned int epoch = 0;
    double* partials = a0.partials;
    while (true) {
        PassIO iov, ioe;
        double c1, c2, rho_k;
        cheb_coeffs(k, rho, a0.cheb_d, a0.cheb_c, c1, c2, rho_k);
        const int nxt = (cur + 1) % 3, prv = (cur + 2) % 3;
        iov.xp = tb_sel(av.tp, cur); iov.xq = tb_sel(av.tq, cur); iov.xprev = tb_sel(av.tp, prv); iov.xnew = tb_sel(av.tp, nxt);
        ioe.xp = iov.xp; ioe.xq = iov.xq; ioe.xprev = tb_sel(av.tq, prv); ioe.xnew = tb_sel(av.tq, nxt);
        iov.c1 = ioe.c1 = c1; iov.c2 = ioe.c2 = c2; iov.use_prev = ioe.use_prev = (k > 0);
        double r0 = 0.0, b0 = 0.0;
        if (WHICH == 0 || WHICH == 2) {
            double a0_ = 0.0, a1_ = 0.0;
            if (ELL == 2 && WHICH == 0) vertex_ell_pass_staged<NT>(av, iov, smem, a0_, a1_);
            else if (ELL == 1 && WHICH == 0) vertex_ell_pass<NT>(av, iov, smem, a0_, a1_);
            else vertex_pass<EPI_CHEB, MODE_OP, true, WHICH == 2, CAS, NT>(av, iov, smem, a0_, a1_);
            r0 += av.weight * a0_; b0 += av.weight * a1_;
        }
        if (WHICH == 1 || WHICH == 2) {
            double a0_ = 0.0, a1_ = 0.0;
            edge_pass<EPI_CHEB, MODE_OP, true, WHICH == 2, CAS, NT>(ae, ioe, smem, a0_, a1_);
            r0 += ae.weight * a0_; b0 += ae.weight * a1_;
        }
        block_sum2(r0, b0, red);
        double* part = partials + (size_t)(k & 1) * 2 * gridDim.x;
        if (threadIdx.x == 0) { part[2 * blockIdx.x] = r0; part[2 * blockIdx.x + 1] = b0; }
        PHW_GRID_SYNC(a0);
        if (threadIdx.x < 32) {
            double s0 = 0.0, s1 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) { s0 += __ldcg(&part[2 * i]); s1 += __ldcg(&part[2 * i + 1]); }
            s0 = warp_sum(s0); s1 = warp_sum(s1);
            if (threadIdx.x == 0) { bc[0] = s0; bc[1] = s1; }
        }
        __syncthreads();
        rr = bc[0]; bb = bc[1];
        __syncthreads();
        if (pc.world > 1) {
            // (1) push the new iterate of the dofs my neighbours hold as external, straight into their buffers
            const int gtid = blockIdx.x * NT + threadIdx.x, gsz = gridDim.x * NT;
            for (int nb = 0; nb < pc.n_nb; ++nb) {
                if (WHICH == 0 || WHICH == 2) {
                    const double* src = iov.xnew; double* dst = pc.nb_tp[nb][nxt];
                    for (int i = gtid; i < pc.n_send_v[nb]; i += gsz) dst[pc.send_v_rem[nb][i]] = src[pc.send_v_loc[nb][i]];
                }
                if (WHICH == 1 || WHICH == 2) {
                    const double* src = ioe.xnew; double* dst = pc.nb_tq[nb][nxt];
                    for (int i = gtid; i < pc.n_send_e[nb]; i += gsz) dst[pc.send_e_rem[nb][i]] = src[pc.send_e_loc[nb][i]];
                }
            }
            __threadfence_system();
            PHW_GRID_SYNC(a0);
            // (2) one thread publishes this rank's residual norms + "iterate pushed" flag and waits for the others
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                double v2[2] = {rr, bb};
                peer_allreduce(pc, a0.ctl, v2, 2);
                a0.ctl->xch[k & 1][0] = v2[0]; a0.ctl->xch[k & 1][1] = v2[1];
                __threadfence();
            }
            PHW_GRID_SYNC(a0);
            rr = __ldcg(&a0.ctl->xch[k & 1][0]); bb = __ldcg(&a0.ctl->xch[k & 1][1]);
        }
        // the residual just measured belongs to x_k (buffer `cur`), not to the x_{k+1} produced in the same pass:
        // a Chebyshev iterate can be exact for the components present while its successor is not (the momentum
        // term), so on convergence the verified iterate x_k is the one that is returned.
        conv = rr <= a0.tol2 * bb;
        if (blockIdx.x == 0 && threadIdx.x == 0 && k < 48) a0.ctl->reshist[a0.which][k] = (bb > 0.0) ? rr / bb : 0.0;
        k += 1;
        if (conv) break;
        rho = rho_k; cur = nxt;
        if (k >= a0.max_iter) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc.cur = cur; sc.k = k; sc.done = 1; sc.rr = rr; sc.bb = bb; sc.rho = rho;
        sc.iters += k; sc.solves += 1;
        if (!conv) a0.ctl->nonconv += 1;
        double rel = (bb > 0.0) ? sqrt(rr / bb) : 0.0;
        if (rel > sc.maxres) sc.maxres = rel;
    }
}


// one inter-GPU exchange of a Chebyshev iteration, called by all threads of all CTAs after the grid-wide
// reduction: push the new iterate of the interface dofs into the neighbours' buffer `nxt`, all-reduce the residual norms
template <int NT>
__device__ __forceinline__ void peer_exchange_iter(const PeerComm& pc, Ctl* ctl, unsigned int& epoch, bool vertex_field,
                                                   const double* xnew, int nxt, int k, double& rr, double& bb) {
    const int gtid = blockIdx.x * NT + threadIdx.x, gsz = gridDim.x * NT;
    for (int nb = 0; nb < pc.n_nb; ++nb) {
        if (vertex_field) {
            double* dst = pc.nb_tp[nb][nxt];
            for (int i = gtid; i < pc.n_send_v[nb]; i += gsz) dst[pc.send_v_rem[nb][i]] = xnew[pc.send_v_loc[nb][i]];
        } else {
            double* dst = pc.nb_tq[nb][nxt];
            for (int i = gtid; i < pc.n_send_e[nb]; i += gsz) dst[pc.send_e_rem[nb][i]] = xnew[pc.send_e_loc[nb][i]];
        }
    }
    __threadfence_system();
    grid_barrier(&ctl->gbar[0], epoch);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double v2[2] = {rr, bb};
        peer_allreduce(pc, ctl, v2, 2);
        ctl->xch[k & 1][0] = v2[0]; ctl->xch[k & 1][1] = v2[1];
        __threadfence();
    }
    grid_barrier(&ctl->gbar[0], epoch);
    rr = __ldcg(&ctl->xch[k & 1][0]); bb = __ldcg(&ctl->xch[k & 1][1]);
}

// ------------------------------------------------------------------------------------------------
// lean variant of the ELL M_p solve (MULTI: with the inter-GPU exchange): the same iteration as solve_coop_k<0,..,ELL>, with a compact argument block
// and a register-light row loop, so that more threads fit on an SM (the streams need the parallelism: phw_stream_probe)
// ------------------------------------------------------------------------------------------------
struct EllSolveArgs {
    int npatch;
    const PatchHdr* hdr;
    const uint32_t* ell_ioff; const uint32_t* ell_woff; const uint16_t* ell_ids; const double* ell_w;
    const int* ghost_v;
    const double* b;
    double* buf[3];
    double cheb_d, cheb_c, tol2;
    int max_iter, which;
    double* partials; Ctl* ctl;
};

template <int NT, int MINB, bool MULTI>
__global__ void __launch_bounds__(NT, MINB) solve_ell_coop_k(const __grid_constant__ EllSolveArgs a, const __grid_constant__ PeerComm pc) {
    extern __shared__ double smem[];
    __shared__ double red[64];
    __shared__ double bc[2];
    __shared__ PatchHdr s_hdr[2];
    SolveCtl& sc = a.ctl->sc[a.which];
    int k = 0, cur = sc.cur;
    double rho = 0.0, rr = 0.0, bb = 0.0;
    bool conv = false;
    unsigned int epoch = 0;
    const uint4* ids4 = reinterpret_cast<const uint4*>(a.ell_ids);
    while (true) {
        double c1, c2, rho_k;
        cheb_coeffs(k, rho, a.cheb_d, a.cheb_c, c1, c2, rho_k);
        const int nxt = (cur + 1) % 3, prv = (cur + 2) % 3;
        typedef typename std::conditional<MULTI, const double*, const double* __restrict__>::type xin_t;   // peers write these buffers
        typedef typename std::conditional<MULTI, double*, double* __restrict__>::type xout_t;
        xin_t xk = cur == 0 ? a.buf[0] : (cur == 1 ? a.buf[1] : a.buf[2]);
        xin_t xprev = prv == 0 ? a.buf[0] : (prv == 1 ? a.buf[1] : a.buf[2]);
        xout_t xnew = nxt == 0 ? a.buf[0] : (nxt == 1 ? a.buf[1] : a.buf[2]);
        double r0 = 0.0, b0 = 0.0;
        if (blockIdx.x < a.npatch) load_hdr_smem(&s_hdr[0], &a.hdr[blockIdx.x]);
        __syncthreads();
        int it = 0;
        for (int P = blockIdx.x; P < a.npatch; P += gridDim.x, ++it) {
            const PatchHdr& h = s_hdr[it & 1];
            if (P + (int)gridDim.x < a.npatch) load_hdr_smem(&s_hdr[(it + 1) & 1], &a.hdr[P + gridDim.x]);
            const int nsl = (h.v_cnt + 31) >> 5;
            uint32_t* s_io = reinterpret_cast<uint32_t*>(smem);
            uint32_t* s_wo = s_io + (nsl + 1);
            double* sx = smem + (nsl + 2);
            for (int i = threadIdx.x; i <= nsl; i += NT) { s_io[i] = __ldg(&a.ell_ioff[h.vs_off + i]); s_wo[i] = __ldg(&a.ell_woff[h.vs_off + i]); }
            for (int i = threadIdx.x; i < h.v_cnt; i += NT) sx[i] = xk[h.v_start + i];
            for (int i = threadIdx.x; i < h.gv_cnt; i += NT) sx[h.v_cnt + i] = xk[__ldg(&a.ghost_v[h.gv_off + i])];
            __syncthreads();
            for (int lv = threadIdx.x; lv < h.v_cnt; lv += NT) {
                const int sl = lv >> 5, lane = lv & 31;
                const uint32_t wo0 = s_wo[sl];
                const int W = (int)((s_wo[sl + 1] - wo0) >> 5);
                const uint4* idp = ids4 + (s_io[sl] >> 3) + lane;
                const double* __restrict__ wrow = a.ell_w + wo0 + lane;
                const int gid = h.v_start + lv;
                const double bv = __ldg(&a.b[gid]);
                const double xpv = (k > 0) ? xprev[gid] : 0.0;
                double val = 0.0, dsum = 0.0;
                for (int c = 0; 8 * c < W; ++c) {
                    const uint4 cd = __ldg(idp + c * 32);
                    const int kk = 8 * c;
                    {
                        const double w0 = (kk + 0 < W) ? __ldg(&wrow[(kk + 0) * 32]) : 0.0, w1 = (kk + 1 < W) ? __ldg(&wrow[(kk + 1) * 32]) : 0.0;
                        const double w2 = (kk + 2 < W) ? __ldg(&wrow[(kk + 2) * 32]) : 0.0, w3 = (kk + 3 < W) ? __ldg(&wrow[(kk + 3) * 32]) : 0.0;
                        val += w0 * sx[cd.x & 0xFFFFu]; val += w1 * sx[cd.x >> 16]; val += w2 * sx[cd.y & 0xFFFFu]; val += w3 * sx[cd.y >> 16];
                        dsum += (w0 + w1) + (w2 + w3);
                    }
                    if (kk + 4 < W) {
                        const double w0 = __ldg(&wrow[(kk + 4) * 32]), w1 = (kk + 5 < W) ? __ldg(&wrow[(kk + 5) * 32]) : 0.0;
                        const double w2 = (kk + 6 < W) ? __ldg(&wrow[(kk + 6) * 32]) : 0.0, w3 = (kk + 7 < W) ? __ldg(&wrow[(kk + 7) * 32]) : 0.0;
                        val += w0 * sx[cd.z & 0xFFFFu]; val += w1 * sx[cd.z >> 16]; val += w2 * sx[cd.w & 0xFFFFu]; val += w3 * sx[cd.w >> 16];
                        dsum += (w0 + w1) + (w2 + w3);
                    }
                }
                const double xc = sx[lv];
                const double pv = 1.0 / dsum;
                val += dsum * xc;
                const double r_ = bv - val;
                double xn = xc + c2 * pv * r_;
                if (k > 0) xn += c1 * (xc - xpv);
                xnew[gid] = xn;
                r0 += pv * r_ * r_;
                b0 += pv * bv * bv;
            }
            __syncthreads();
        }
        block_sum2(r0, b0, red);
        double* part = a.partials + (size_t)(k & 1) * 2 * gridDim.x;
        if (threadIdx.x == 0) { part[2 * blockIdx.x] = r0; part[2 * blockIdx.x + 1] = b0; }
        grid_barrier(&a.ctl->gbar[0], epoch);
        if (threadIdx.x < 32) {
            double s0 = 0.0, s1 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) { s0 += __ldcg(&part[2 * i]); s1 += __ldcg(&part[2 * i + 1]); }
            s0 = warp_sum(s0); s1 = warp_sum(s1);
            if (threadIdx.x == 0) { bc[0] = s0; bc[1] = s1; }
        }
        __syncthreads();
        rr = bc[0]; bb = bc[1];
        __syncthreads();
        if (MULTI) peer_exchange_iter<NT>(pc, a.ctl, epoch, true, xnew, nxt, k, rr, bb);
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
        if (!conv) a.ctl->nonconv += 1;
        double rel = (bb > 0.0) ? sqrt(rr / bb) : 0.0;
        if (rel > sc.maxres) sc.maxres = rel;
    }
}

// lean variant of the cell-based M_q solve (same iteration as solve_coop_k<1>; MULTI: with the inter-GPU exchange): compact
// arguments, restrict-qualified streams, packed cell records
struct MqSolveArgs {
    int npatch;
    const PatchHdr* hdr;
    const uint4* cellq;          // [npc][2] packed edge-row cell records: {rece.x, rece.y, g0} {g1, g2}: one 32-byte stream
    const int* ghost_e; const uint32_t* eadj;
    const double* b; const double* pinv;
    double* buf[3];
    double cheb_d, cheb_c, tol2;
    int max_iter, which;
    double* partials; Ctl* ctl;
};

template <int NT, int MINB, int UC, int UR, bool MULTI>
__global__ void __launch_bounds__(NT, MINB) solve_mq_coop_k(const __grid_constant__ MqSolveArgs a, const __grid_constant__ PeerComm pc) {
    extern __shared__ double smem[];
    __shared__ double red[64];
    __shared__ double bc[2];
    __shared__ PatchHdr s_hdr[2];
    SolveCtl& sc = a.ctl->sc[a.which];
    int k = 0, cur = sc.cur;
    double rho = 0.0, rr = 0.0, bb = 0.0;
    bool conv = false;
    unsigned int epoch = 0;
    while (true) {
        double c1, c2, rho_k;
        cheb_coeffs(k, rho, a.cheb_d, a.cheb_c, c1, c2, rho_k);
        const int nxt = (cur + 1) % 3, prv = (cur + 2) % 3;
        typedef typename std::conditional<MULTI, const double*, const double* __restrict__>::type xin_t;   // peers write these buffers
        typedef typename std::conditional<MULTI, double*, double* __restrict__>::type xout_t;
        xin_t xk = cur == 0 ? a.buf[0] : (cur == 1 ? a.buf[1] : a.buf[2]);
        xin_t xprev = prv == 0 ? a.buf[0] : (prv == 1 ? a.buf[1] : a.buf[2]);
        xout_t xnew = nxt == 0 ? a.buf[0] : (nxt == 1 ? a.buf[1] : a.buf[2]);
        double r0 = 0.0, b0 = 0.0;
        if (blockIdx.x < a.npatch) load_hdr_smem(&s_hdr[0], &a.hdr[blockIdx.x]);
        __syncthreads();
        int it = 0;
        for (int P = blockIdx.x; P < a.npatch; P += gridDim.x, ++it) {
            const PatchHdr& h = s_hdr[it & 1];
            if (P + (int)gridDim.x < a.npatch) load_hdr_smem(&s_hdr[(it + 1) & 1], &a.hdr[P + gridDim.x]);
            const int nle = h.e_cnt + h.ge_cnt;
            double* __restrict__ sq = smem;
            double* __restrict__ sc3 = sq + ((nle + 1) & ~1);
            for (int i = threadIdx.x; i < h.e_cnt; i += NT) sq[i] = xk[h.e_start + i];
            for (int i = threadIdx.x; i < h.ge_cnt; i += NT) sq[h.e_cnt + i] = xk[__ldg(&a.ghost_e[h.ge_off + i])];
            __syncthreads();
            for (int base = threadIdx.x; base < h.n_ecells; base += UC * NT) {
                uint4 qa[UC], qb[UC];
#pragma unroll
                for (int u = 0; u < UC; ++u) {
                    const int lc = base + u * NT;
                    if (lc < h.n_ecells) { qa[u] = __ldg(&a.cellq[2 * (size_t)(h.cell_off + lc)]); qb[u] = __ldg(&a.cellq[2 * (size_t)(h.cell_off + lc) + 1]); }
                }
#pragma unroll
                for (int u = 0; u < UC; ++u) {
                    const int lc = base + u * NT;
                    if (lc < h.n_ecells) {
                        const uint32_t e0 = qa[u].x & 0xFFFFu, e1 = qa[u].x >> 16, e2 = qa[u].y & 0xFFFFu;
                        const double g0 = __hiloint2double((int)qa[u].w, (int)qa[u].z);
                        const double g1 = __hiloint2double((int)qb[u].y, (int)qb[u].x), g2 = __hiloint2double((int)qb[u].w, (int)qb[u].z);
                        const double S = g0 + g1 + g2;
                        const double s0 = (e0 & 0x8000u) ? -1.0 : 1.0, s1 = (e1 & 0x8000u) ? -1.0 : 1.0, s2 = (e2 & 0x8000u) ? -1.0 : 1.0;
                        const double t0 = s0 * sq[e0 & 0x3FFFu], t1 = s1 * sq[e1 & 0x3FFFu], t2 = s2 * sq[e2 & 0x3FFFu];
                        const double ST = S * (t0 + t1 + t2);
                        sc3[3 * lc] = s0 * (ST + 2.0 * S * t0 - 4.0 * (g0 * t0 + g2 * t1 + g1 * t2));
                        sc3[3 * lc + 1] = s1 * (ST + 2.0 * S * t1 - 4.0 * (g2 * t0 + g1 * t1 + g0 * t2));
                        sc3[3 * lc + 2] = s2 * (ST + 2.0 * S * t2 - 4.0 * (g1 * t0 + g0 * t1 + g2 * t2));
                    }
                }
            }
            __syncthreads();
            for (int base = threadIdx.x; base < h.e_cnt; base += UR * NT) {
                uint32_t adjb[UR]; double bvb[UR], pvb[UR], xpb[UR];
#pragma unroll
                for (int u = 0; u < UR; ++u) {
                    const int le = base + u * NT;
                    if (le < h.e_cnt) {
                        const int gid = h.e_start + le;
                        adjb[u] = __ldg(&a.eadj[gid]); bvb[u] = __ldg(&a.b[gid]); pvb[u] = __ldg(&a.pinv[gid]);
                        xpb[u] = (k > 0) ? xprev[gid] : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < UR; ++u) {
                    const int le = base + u * NT;
                    if (le < h.e_cnt) {
                        const uint32_t ca = adjb[u] & 0xFFFFu, cb = adjb[u] >> 16;
                        double val = sc3[3 * (ca >> 2) + (ca & 3u)];
                        if (cb != 0xFFFFu) val += sc3[3 * (cb >> 2) + (cb & 3u)];
                        const double xc = sq[le];
                        const double r_ = bvb[u] - val;
                        double xn = xc + c2 * pvb[u] * r_;
                        if (k > 0) xn += c1 * (xc - xpb[u]);
                        xnew[h.e_start + le] = xn;
                        r0 += pvb[u] * r_ * r_;
                        b0 += pvb[u] * bvb[u] * bvb[u];
                    }
                }
            }
            __syncthreads();
        }
        block_sum2(r0, b0, red);
        double* part = a.partials + (size_t)(k & 1) * 2 * gridDim.x;
        if (threadIdx.x == 0) { part[2 * blockIdx.x] = r0; part[2 * blockIdx.x + 1] = b0; }
        grid_barrier(&a.ctl->gbar[0], epoch);
        if (threadIdx.x < 32) {
            double s0 = 0.0, s1 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) { s0 += __ldcg(&part[2 * i]); s1 += __ldcg(&part[2 * i + 1]); }
            s0 = warp_sum(s0); s1 = warp_sum(s1);
            if (threadIdx.x == 0) { bc[0] = s0; bc[1] = s1; }
        }
        __syncthreads();
        rr = bc[0]; bb = bc[1];
        __syncthreads();
        if (MULTI) peer_exchange_iter<NT>(pc, a.ctl, epoch, false, xnew, nxt, k, rr, bb);
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
        if (!conv) a.ctl->nonconv += 1;
        double rel = (bb > 0.0) ? sqrt(rr / bb) : 0.0;
        if (rel > sc.maxres) sc.maxres = rel;
    }
}

// ------------------------------------------------------------------------------------------------
// lean energy reductions (replace assemble(scalar) of H_wave, wave_2D.py:882,907):  red[slot] = scale x^T M x  on the assembled
// ELL rows (M_p) / the packed cell records (M_q); per-block partials, summed in block order by the last block to finish
// ------------------------------------------------------------------------------------------------
struct DotArgs {
    int npatch;
    const PatchHdr* hdr;
    const uint32_t* ell_ioff; const uint32_t* ell_woff; const uint16_t* ell_ids; const double* ell_w;   // M_p
    const uint4* cellq; const uint32_t* eadj;                                                           // M_q
    const int* ghost;
    const double* x;
    double scale;
    int slot, counter_id;
    double* partials; Ctl* ctl;
};

__device__ __forceinline__ void dot_finish(const DotArgs& a, double acc, double* red) {
    __shared__ bool is_last;
    double z = 0.0;
    block_sum2(acc, z, red);
    if (threadIdx.x == 0) {
        a.partials[blockIdx.x] = acc;
        __threadfence();
        is_last = (atomicAdd(&a.ctl->counters[a.counter_id], 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double t = 0.0; z = 0.0;
    for (int p = threadIdx.x; p < (int)gridDim.x; p += blockDim.x) t += __ldcg(&a.partials[p]);
    block_sum2(t, z, red);
    if (threadIdx.x == 0) { a.ctl->counters[a.counter_id] = 0; a.ctl->red[a.slot] = a.scale * t; }
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) dot_ell_k(const __grid_constant__ DotArgs a) {
    extern __shared__ double smem[];
    __shared__ double red[64];
    __shared__ PatchHdr s_hdr[2];
    const uint4* ids4 = reinterpret_cast<const uint4*>(a.ell_ids);
    const double* __restrict__ x = a.x;
    double acc = 0.0;
    if (blockIdx.x < a.npatch) load_hdr_smem(&s_hdr[0], &a.hdr[blockIdx.x]);
    __syncthreads();
    int it = 0;
    for (int P = blockIdx.x; P < a.npatch; P += gridDim.x, ++it) {
        const PatchHdr& h = s_hdr[it & 1];
        if (P + (int)gridDim.x < a.npatch) load_hdr_smem(&s_hdr[(it + 1) & 1], &a.hdr[P + gridDim.x]);
        const int nsl = (h.v_cnt + 31) >> 5;
        uint32_t* s_io = reinterpret_cast<uint32_t*>(smem);
        uint32_t* s_wo = s_io + (nsl + 1);
        double* sx = smem + (nsl + 2);
        for (int i = threadIdx.x; i <= nsl; i += NT) { s_io[i] = __ldg(&a.ell_ioff[h.vs_off + i]); s_wo[i] = __ldg(&a.ell_woff[h.vs_off + i]); }
        for (int i = threadIdx.x; i < h.v_cnt; i += NT) sx[i] = x[h.v_start + i];
        for (int i = threadIdx.x; i < h.gv_cnt; i += NT) sx[h.v_cnt + i] = x[__ldg(&a.ghost[h.gv_off + i])];
        __syncthreads();
        for (int lv = threadIdx.x; lv < h.v_cnt; lv += NT) {
            const int sl = lv >> 5, lane = lv & 31;
            const uint32_t wo0 = s_wo[sl];
            const int W = (int)((s_wo[sl + 1] - wo0) >> 5);
            const uint4* idp = ids4 + (s_io[sl] >> 3) + lane;
            const double* __restrict__ wrow = a.ell_w + wo0 + lane;
            double val = 0.0, dsum = 0.0;
            for (int c = 0; 8 * c < W; ++c) {
                const uint4 cd = __ldg(idp + c * 32);
                const int kk = 8 * c;
                {
                    const double w0 = (kk + 0 < W) ? __ldg(&wrow[(kk + 0) * 32]) : 0.0, w1 = (kk + 1 < W) ? __ldg(&wrow[(kk + 1) * 32]) : 0.0;
                    const double w2 = (kk + 2 < W) ? __ldg(&wrow[(kk + 2) * 32]) : 0.0, w3 = (kk + 3 < W) ? __ldg(&wrow[(kk + 3) * 32]) : 0.0;
                    val += w0 * sx[cd.x & 0xFFFFu]; val += w1 * sx[cd.x >> 16]; val += w2 * sx[cd.y & 0xFFFFu]; val += w3 * sx[cd.y >> 16];
                    dsum += (w0 + w1) + (w2 + w3);
                }
                if (kk + 4 < W) {
                    const double w0 = __ldg(&wrow[(kk + 4) * 32]), w1 = (kk + 5 < W) ? __ldg(&wrow[(kk + 5) * 32]) : 0.0;
                    const double w2 = (kk + 6 < W) ? __ldg(&wrow[(kk + 6) * 32]) : 0.0, w3 = (kk + 7 < W) ? __ldg(&wrow[(kk + 7) * 32]) : 0.0;
                    val += w0 * sx[cd.z & 0xFFFFu]; val += w1 * sx[cd.z >> 16]; val += w2 * sx[cd.w & 0xFFFFu]; val += w3 * sx[cd.w >> 16];
                    dsum += (w0 + w1) + (w2 + w3);
                }
            }
            const double xc = sx[lv];
            acc += xc * (val + dsum * xc);
        }
        __syncthreads();
    }
    dot_finish(a, acc, red);
}

template <int NT, int MINB, int UC>
__global__ void __launch_bounds__(NT, MINB) dot_mq_k(const __grid_constant__ DotArgs a) {
    extern __shared__ double smem[];
    __shared__ double red[64];
    __shared__ PatchHdr s_hdr[2];
    const double* __restrict__ x = a.x;
    double acc = 0.0;
    if (blockIdx.x < a.npatch) load_hdr_smem(&s_hdr[0], &a.hdr[blockIdx.x]);
    __syncthreads();
    int it = 0;
    for (int P = blockIdx.x; P < a.npatch; P += gridDim.x, ++it) {
        const PatchHdr& h = s_hdr[it & 1];
        if (P + (int)gridDim.x < a.npatch) load_hdr_smem(&s_hdr[(it + 1) & 1], &a.hdr[P + gridDim.x]);
        const int nle = h.e_cnt + h.ge_cnt;
        double* __restrict__ sq = smem;
        double* __restrict__ sc3 = sq + ((nle + 1) & ~1);
        for (int i = threadIdx.x; i < h.e_cnt; i += NT) sq[i] = x[h.e_start + i];
        for (int i = threadIdx.x; i < h.ge_cnt; i += NT) sq[h.e_cnt + i] = x[__ldg(&a.ghost[h.ge_off + i])];
        __syncthreads();
        for (int base = threadIdx.x; base < h.n_ecells; base += UC * NT) {
            uint4 qa[UC], qb[UC];
#pragma unroll
            for (int u = 0; u < UC; ++u) {
                const int lc = base + u * NT;
                if (lc < h.n_ecells) { qa[u] = __ldg(&a.cellq[2 * (size_t)(h.cell_off + lc)]); qb[u] = __ldg(&a.cellq[2 * (size_t)(h.cell_off + lc) + 1]); }
            }
#pragma unroll
            for (int u = 0; u < UC; ++u) {
                const int lc = base + u * NT;
                if (lc < h.n_ecells) {
                    const uint32_t e0 = qa[u].x & 0xFFFFu, e1 = qa[u].x >> 16, e2 = qa[u].y & 0xFFFFu;
                    const double g0 = __hiloint2double((int)qa[u].w, (int)qa[u].z);
                    const double g1 = __hiloint2double((int)qb[u].y, (int)qb[u].x), g2 = __hiloint2double((int)qb[u].w, (int)qb[u].z);
                    const double S = g0 + g1 + g2;
                    const double s0 = (e0 & 0x8000u) ? -1.0 : 1.0, s1 = (e1 & 0x8000u) ? -1.0 : 1.0, s2 = (e2 & 0x8000u) ? -1.0 : 1.0;
                    const double t0 = s0 * sq[e0 & 0x3FFFu], t1 = s1 * sq[e1 & 0x3FFFu], t2 = s2 * sq[e2 & 0x3FFFu];
                    const double ST = S * (t0 + t1 + t2);
                    sc3[3 * lc] = s0 * (ST + 2.0 * S * t0 - 4.0 * (g0 * t0 + g2 * t1 + g1 * t2));
                    sc3[3 * lc + 1] = s1 * (ST + 2.0 * S * t1 - 4.0 * (g2 * t0 + g1 * t1 + g0 * t2));
                    sc3[3 * lc + 2] = s2 * (ST + 2.0 * S * t2 - 4.0 * (g1 * t0 + g0 * t1 + g2 * t2));
                }
            }
        }
        __syncthreads();
        for (int le = threadIdx.x; le < h.e_cnt; le += NT) {
            const uint32_t adj = __ldg(&a.eadj[h.e_start + le]);
            const uint32_t ca = adj & 0xFFFFu, cb = adj >> 16;
            double val = sc3[3 * (ca >> 2) + (ca & 3u)];
            if (cb != 0xFFFFu) val += sc3[3 * (cb >> 2) + (cb & 3u)];
            acc += sq[le] * val;
        }
        __syncthreads();
    }
    dot_finish(a, acc, red);
}

// ------------------------------------------------------------------------------------------------
// higher-order pairs (P_k x RT_k, k <= 3): the operators are assembled once on the host (fem_ho.py) and applied as
// CSR matrices; same epilogues and the same persistent cooperative Chebyshev solve as the matrix-free path.
// ------------------------------------------------------------------------------------------------
struct Csr { int n_rows; const int* indptr; const int* idx; const double* val; int lanes; };   // lanes: threads per row (4..32)

// one row of A x by a group of L consecutive lanes (L | 32): the lanes stride over the row's entries (coalesced 8 L-byte
// segments of val / 4 L-byte segments of idx) and combine with a butterfly, so every lane of the group returns the sum -
// in a fixed order, independent of the launch configuration
template <int L>
__device__ __forceinline__ double csr_row_group(const Csr& A, int r, const double* __restrict__ x, int sub) {
    double v = 0.0;
    const int j1 = __ldg(&A.indptr[r + 1]);
    for (int j = __ldg(&A.indptr[r]) + sub; j < j1; j += L) v += __ldg(&A.val[j]) * x[__ldg(&A.idx[j])];
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// y = alpha A x
template <int L>
__global__ void __launch_bounds__(256) csr_store_k(Csr A, const double* x, double alpha, double* y) {
    const int sub = threadIdx.x % L, g = (blockIdx.x * blockDim.x + threadIdx.x) / L, ng = gridDim.x * blockDim.x / L;
    const int nround = (A.n_rows + ng - 1) / ng;
    for (int i = 0; i < nround; ++i) {          // whole warps stay converged for the shuffles
        const int r = g + i * ng;
        const double v = csr_row_group<L>(A, r < A.n_rows ? r : A.n_rows - 1, x, sub);
        if (r < A.n_rows && sub == 0) y[r] = alpha * v;
    }
}
// ctl->red[slot] = scale x^T A x: per-block partials, summed in block order by the last block to finish
template <int L>
__global__ void __launch_bounds__(256) csr_dot_k(Csr A, const double* x, double scale, Ctl* ctl, int slot, double* partials, int counter_id) {
    __shared__ double red[64];
    __shared__ bool is_last;
    const int sub = threadIdx.x % L, g = (blockIdx.x * blockDim.x + threadIdx.x) / L, ng = gridDim.x * blockDim.x / L;
    const int nround = (A.n_rows + ng - 1) / ng;
    double s = 0.0, z = 0.0;
    for (int i = 0; i < nround; ++i) {
        const int r = g + i * ng;
        const double v = csr_row_group<L>(A, r < A.n_rows ? r : A.n_rows - 1, x, sub);
        if (r < A.n_rows && sub == 0) s += x[r] * v;
    }
    block_sum2(s, z, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = s;
        __threadfence();
        is_last = (atomicAdd(&ctl->counters[counter_id], 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double t = 0.0; z = 0.0;
    for (int p = threadIdx.x; p < (int)gridDim.x; p += blockDim.x) t += __ldcg(&partials[p]);
    block_sum2(t, z, red);
    if (threadIdx.x == 0) { ctl->counters[counter_id] = 0; ctl->red[slot] = scale * t; }
}
struct CsrSolveArgs {
    Csr A; TriBuf tb; const double* b; const double* pinv; double cheb_d, cheb_c, tol2; int max_iter, which;
    double* partials; Ctl* ctl;
    int gbar;
};
template <int L>
__global__ void __launch_bounds__(256) solve_coop_csr_k(const CsrSolveArgs a) {
    __shared__ double red[64];
    __shared__ double bc[2];
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    SolveCtl& sc = a.ctl->sc[a.which];
    int k = 0, cur = sc.cur;
    double rho = 0.0, rr = 0.0, bb = 0.0;
    bool conv = false;
    unsigned int epoch = 0;
    const int sub = threadIdx.x % L, g = (blockIdx.x * blockDim.x + threadIdx.x) / L, ng = gridDim.x * blockDim.x / L;
    const int nround = (a.A.n_rows + ng - 1) / ng;
    while (true) {
        double c1, c2, rho_k;
        cheb_coeffs(k, rho, a.cheb_d, a.cheb_c, c1, c2, rho_k);
        const int nxt = (cur + 1) % 3, prv = (cur + 2) % 3;
        const double* x = tb_sel(a.tb, cur); const double* xp = tb_sel(a.tb, prv); double* xn = tb_sel(a.tb, nxt);
        double r0 = 0.0, b0 = 0.0;
        for (int i = 0; i < nround; ++i) {
            const int r = g + i * ng;
            const bool live = r < a.A.n_rows;
            double bv = 0.0, pv = 0.0, xc = 0.0, xpv = 0.0;
            if (live && sub == 0) { bv = a.b[r]; pv = a.pinv[r]; xc = x[r]; if (k > 0) xpv = xp[r]; }   // requested before the row sum
            const double val = csr_row_group<L>(a.A, live ? r : a.A.n_rows - 1, x, sub);
            if (live && sub == 0) {
                const double res = bv - val;
                double v = xc + c2 * pv * res;
                if (k > 0) v += c1 * (xc - xpv);
                xn[r] = v;
                r0 += pv * res * res; b0 += pv * bv * bv;
            }
        }
        block_sum2(r0, b0, red);
        double* part = a.partials + (size_t)(k & 1) * 2 * gridDim.x;
        if (threadIdx.x == 0) { part[2 * blockIdx.x] = r0; part[2 * blockIdx.x + 1] = b0; }
        PHW_GRID_SYNC(a);
        if (threadIdx.x < 32) {
            double s0 = 0.0, s1 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) { s0 += __ldcg(&part[2 * i]); s1 += __ldcg(&part[2 * i + 1]); }
            s0 = warp_sum(s0); s1 = warp_sum(s1);
            if (threadIdx.x == 0) { bc[0] = s0; bc[1] = s1; }
        }
        __syncthreads();
        rr = bc[0]; bb = bc[1];
        __syncthreads();
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
        if (!conv) a.ctl->nonconv += 1;
        double rel = (bb > 0.0) ? sqrt(rr / bb) : 0.0;
        if (rel > sc.maxres) sc.maxres = rel;
    }
}

// ------------------------------------------------------------------------------------------------
// Mass solves on meshes of poor quality: Jacobi-preconditioned CONJUGATE GRADIENTS, all iterations in one cooperative launch.
// The Chebyshev solves iterate on a rigorous element-wise spectral interval; a few stretched cells widen that interval for the
// whole mesh although they only add a few outlying eigenvalues.  CG needs no interval and converges on the true spectrum (outliers
// cost about one iteration each), so phw_solver_create selects it when the element-wise kappa bound exceeds 10 (phw_params.flags
// bit 16 / PHW_PCG force it on or off).  The reference's sparse LU has no mesh-quality dependence at all (wave_2D.py:820,854); CG
// removes the dependence on the WORST cell, not the one on uniformly stretched cells (there the Jacobi-scaled RT1 mass matrix is
// ill-conditioned for real: aspect ratio 12, kappa = 119, ~150 iterations either way).  Same rows (vertex_pass / edge_pass, STORE epilogue), same stop test ||P^-1/2 r|| <= tol ||P^-1/2 b|| (on
// the recursively updated residual), same deterministic reductions; four grid barriers per iteration instead of one, which is why
// it is not the default on quality meshes.  One GPU, weak form, unbatched.
// ------------------------------------------------------------------------------------------------
struct PcgArgs { double *r, *u; const double* b; const double* pinv; int n; };   // search direction and A pd live in the two idle iterate buffers
template <int WHICH, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) solve_pcg_coop_k(const __grid_constant__ RowArgs a, const __grid_constant__ PcgArgs v) {
    extern __shared__ double smem[];
    __shared__ double red[64];
    __shared__ double bc[2];
    SolveCtl& sc = a.ctl->sc[WHICH];
    unsigned int epoch = 0, nred = 0;
    const int cur = sc.cur;
    double* x = tb_sel(WHICH == 0 ? a.tp : a.tq, cur);
    double* pd = tb_sel(WHICH == 0 ? a.tp : a.tq, (cur + 1) % 3);
    double* w = tb_sel(WHICH == 0 ? a.tp : a.tq, (cur + 2) % 3);
    const int gtid = blockIdx.x * NT + threadIdx.x, gsz = gridDim.x * NT;
    RowArgs am = a;
    am.y = w; am.cM = 1.0; am.cD = 0.0; am.cB = 0.0;
    // w = A src on the owned rows, visible grid-wide on return
    auto matvec = [&](const double* src) {
        PassIO io;
        io.xp = src; io.xq = src; io.xprev = nullptr; io.xnew = nullptr; io.c1 = 0.0; io.c2 = 0.0; io.use_prev = 0;
        double t0 = 0.0, t1 = 0.0;
        if (WHICH == 0) vertex_pass<EPI_STORE, MODE_OP, true, false, false, NT>(am, io, smem, t0, t1);
        else edge_pass<EPI_STORE, MODE_OP, true, false, false, NT>(am, io, smem, t0, t1);
        grid_barrier(&a.ctl->gbar[0], epoch);
    };
    // grid-wide sums of (s0, s1), identical bits in every thread; contains one grid barrier
    auto reduce2 = [&](double& s0, double& s1) {
        block_sum2(s0, s1, red);
        double* part = a.partials + (size_t)(nred & 1u) * 2 * gridDim.x;
        nred++;
        if (threadIdx.x == 0) { part[2 * blockIdx.x] = s0; part[2 * blockIdx.x + 1] = s1; }
        grid_barrier(&a.ctl->gbar[0], epoch);
        if (threadIdx.x < 32) {
            double q0 = 0.0, q1 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) { q0 += __ldcg(&part[2 * i]); q1 += __ldcg(&part[2 * i + 1]); }
            q0 = warp_sum(q0); q1 = warp_sum(q1);
            if (threadIdx.x == 0) { bc[0] = q0; bc[1] = q1; }
        }
        __syncthreads();
        s0 = bc[0]; s1 = bc[1];
        __syncthreads();
    };
    matvec(x);
    double gam = 0.0, bb = 0.0;
    for (int i = gtid; i < v.n; i += gsz) {
        const double bv = v.b[i], pv = v.pinv[i];
        const double r = bv - __ldcg(&w[i]);
        const double u = pv * r;
        v.r[i] = r; v.u[i] = u; pd[i] = u;
        gam += r * u; bb += pv * bv * bv;
    }
    reduce2(gam, bb);                                   // the barrier inside also completes pd before the first pass reads it
    int k = 0;
    bool conv = gam <= a.tol2 * bb;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.ctl->reshist[WHICH][0] = (bb > 0.0) ? gam / bb : 0.0;
    while (!conv && k < a.max_iter) {
        matvec(pd);
        double del = 0.0, z0 = 0.0;
        for (int i = gtid; i < v.n; i += gsz) del += pd[i] * __ldcg(&w[i]);
        reduce2(del, z0);
        const double alpha = gam / del;
        double gn = 0.0, z1 = 0.0;
        for (int i = gtid; i < v.n; i += gsz) {
            x[i] += alpha * pd[i];
            const double r = v.r[i] - alpha * __ldcg(&w[i]);
            const double u = v.pinv[i] * r;
            v.r[i] = r; v.u[i] = u;
            gn += r * u;
        }
        reduce2(gn, z1);
        k += 1;
        if (blockIdx.x == 0 && threadIdx.x == 0 && k < 48) a.ctl->reshist[WHICH][k] = (bb > 0.0) ? gn / bb : 0.0;
        conv = gn <= a.tol2 * bb;
        const double beta = gn / gam;
        gam = gn;
        if (conv) break;
        for (int i = gtid; i < v.n; i += gsz) pd[i] = v.u[i] + beta * pd[i];
        grid_barrier(&a.ctl->gbar[0], epoch);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc.k = k; sc.done = 1; sc.rr = gam; sc.bb = bb;
        sc.iters += k; sc.solves += 1;
        if (!conv) a.ctl->nonconv += 1;
        const double rel = (bb > 0.0) ? sqrt(gam / bb) : 0.0;
        if (rel > sc.maxres) sc.maxres = rel;
    }
}

// All Chebyshev (ellipse) iterations of the coupled block system of the implicit schemes (IE, SM; wave_2D.py:305-319,560-572) for the
// higher-order pairs, on assembled CSR rows, in ONE cooperative launch:
//     M_p v_p + cDv D^T v_q = b_p ,    M_q v_q + cDe D v_p = b_q          (cDv = theta dt K,  cDe = -theta dt / rho)
// Both row families are updated from the same iterate (Jacobi-preconditioned, simultaneous), residual norm = wv |r_p|^2_P + we |r_q|^2_P
// as in the P1 x RT1 block solve (solve_coop_k<2>).  Eight lanes per row for all four matrices.
struct CsrBlockArgs {
    Csr Mp, Mq, D, DT; TriBuf tp, tq; const double *bp, *bq, *pinv_p, *pinv_q;
    double cDv, cDe, wv, we, cheb_d, cheb_c, tol2; int max_iter;
    double* partials; Ctl* ctl;
};
__global__ void __launch_bounds__(256) solve_block_csr_k(const CsrBlockArgs a) {
    constexpr int L = 8;
    __shared__ double red[64];
    __shared__ double bc[2];
    SolveCtl& sc = a.ctl->sc[2];
    int k = 0, cur = sc.cur;
    double rho = 0.0, rr = 0.0, bb = 0.0;
    bool conv = false;
    unsigned int epoch = 0;
    const int sub = threadIdx.x % L, g = (blockIdx.x * blockDim.x + threadIdx.x) / L, ng = gridDim.x * blockDim.x / L;
    const int n_p = a.Mp.n_rows, n_q = a.Mq.n_rows, n_all = n_p + n_q;
    const int nround = (n_all + ng - 1) / ng;
    while (true) {
        double c1, c2, rho_k;
        cheb_coeffs(k, rho, a.cheb_d, a.cheb_c, c1, c2, rho_k);
        const int nxt = (cur + 1) % 3, prv = (cur + 2) % 3;
        const double* xp = tb_sel(a.tp, cur); const double* xq = tb_sel(a.tq, cur);
        double r0 = 0.0, b0 = 0.0;
        for (int i = 0; i < nround; ++i) {
            const int r = g + i * ng;
            const bool live = r < n_all;
            const bool vertex = r < n_p;                       // uniform over the 8 lanes of a row; rows of a warp may be of both kinds
            const int row = live ? (vertex ? r : r - n_p) : 0;
            // operands selected by pointer, not by branch: every lane of the warp reaches the same shuffles
            Csr A1, A2;
            A1.indptr = vertex ? a.Mp.indptr : a.Mq.indptr; A1.idx = vertex ? a.Mp.idx : a.Mq.idx; A1.val = vertex ? a.Mp.val : a.Mq.val;
            A2.indptr = vertex ? a.DT.indptr : a.D.indptr; A2.idx = vertex ? a.DT.idx : a.D.idx; A2.val = vertex ? a.DT.val : a.D.val;
            const double val = csr_row_group<L>(A1, row, vertex ? xp : xq, sub) + (vertex ? a.cDv : a.cDe) * csr_row_group<L>(A2, row, vertex ? xq : xp, sub);
            if (live && sub == 0) {
                const double* x = vertex ? xp : xq;
                const double bv = vertex ? a.bp[row] : a.bq[row], pv = vertex ? a.pinv_p[row] : a.pinv_q[row], w = vertex ? a.wv : a.we;
                const double xc = x[row];
                const double res = bv - val;
                double v = xc + c2 * pv * res;
                if (k > 0) v += c1 * (xc - (vertex ? tb_sel(a.tp, prv) : tb_sel(a.tq, prv))[row]);
                (vertex ? tb_sel(a.tp, nxt) : tb_sel(a.tq, nxt))[row] = v;
                r0 += w * pv * res * res; b0 += w * pv * bv * bv;
            }
        }
        block_sum2(r0, b0, red);
        double* part = a.partials + (size_t)(k & 1) * 2 * gridDim.x;
        if (threadIdx.x == 0) { part[2 * blockIdx.x] = r0; part[2 * blockIdx.x + 1] = b0; }
        grid_barrier(&a.ctl->gbar[0], epoch);
        if (threadIdx.x < 32) {
            double s0 = 0.0, s1 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) { s0 += __ldcg(&part[2 * i]); s1 += __ldcg(&part[2 * i + 1]); }
            s0 = warp_sum(s0); s1 = warp_sum(s1);
            if (threadIdx.x == 0) { bc[0] = s0; bc[1] = s1; }
        }
        __syncthreads();
        rr = bc[0]; bb = bc[1];
        __syncthreads();
        conv = rr <= a.tol2 * bb;
        if (blockIdx.x == 0 && threadIdx.x == 0 && k < 48) a.ctl->reshist[2][k] = (bb > 0.0) ? rr / bb : 0.0;
        k += 1;
        if (conv) break;
        rho = rho_k; cur = nxt;
        if (k >= a.max_iter) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc.cur = cur; sc.k = k; sc.done = 1; sc.rr = rr; sc.bb = bb; sc.rho = rho;
        sc.iters += k; sc.solves += 1;
        if (!conv) a.ctl->nonconv += 1;
        double rel = (bb > 0.0) ? sqrt(rr / bb) : 0.0;
        if (rel > sc.maxres) sc.maxres = rel;
    }
}

}  // namespace phw
