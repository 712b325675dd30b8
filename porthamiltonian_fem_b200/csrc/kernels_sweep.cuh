// kernels_sweep.cuh - the batched parameter sweep (BASELINE configs[4]) with the CASE INDEX FASTEST.
//
// The B cases of a sweep share the mesh, hence M_p, M_q, D and b_L; only the wave speed / density (scalar factors of the coupling
// terms) and the lumped constants differ (the loop over cases of main_wave_2D.py:24-50,119-140).  Round 1 put the cases side by side
// as B copies of the mesh inside one device mesh, which re-reads identical index / metric / weight data B times - about half of the
// bytes of an M_q iteration.  Here every vector is an [n_dof][B] array and the operators are ONE set of assembled CSR rows
// (5 non-zeros per RT1 row, ~7 per P1 row: a few hundred KB, cache resident).  A warp owns a chunk of 32 x CPL consecutive cases
// (CPL = 1, 2 or 4 per lane) and a block of rows (SweepMap below); the row's indices and values are the same for all lanes (broadcast
// loads) and every vector access x[idx * B + c] is a fully coalesced segment of 256 x CPL bytes.  Algorithmic traffic: 32 B per dof,
// case and Chebyshev iteration (x_k, x_{k-1}, b read; x_{k+1} written) instead of 64 (M_q) and 80 (M_p) in the replicated layout.
// Sparse matrix x dense block: no shared-memory staging, no atomics, fixed summation order.
//
// Per-case quantities (energies, port fluxes) are per-lane sums; blocks of rows are combined in a fixed order by a finishing kernel,
// so every case's H_array is bit-reproducible.  The Chebyshev stop test uses the residual aggregated over all cases with the
// tolerance tightened to tol / sqrt(B), as in the grouped layout.
#pragma once

namespace phw {

// Row mapping (A/B knob PHW_SWEEP_MAP).  mode 0 (blocks, default): warp -> (chunk of 32 cases, one contiguous block of rows): neighbour
// rows of the same block hit in L1; rows of other blocks were touched half an iteration earlier or later by another warp - beyond
// what the L2 retains (ncu: L1 hit rate 55 %, L2 9 %, DRAM traffic 2.6 x the algorithmic bytes).  mode 1 (cyclic tiles): CTA -> (chunk,
// row group g); its 8 warps take 8 consecutive rows per sweep step, r = (step * n_row_groups + g) * 8 + w, so the CTAs of a chunk work
// on one window of consecutive rows and neighbour rows hit in L2 (ncu: DRAM traffic = the algorithmic 4 vectors per iteration) - and
// the kernel is no faster (3.5-4.0e9 against 4.07e9): it is bound by the load/store unit (~24 requests per row and 32 cases, 14 of
// them the row's own indices and values), not by DRAM.
// `wid` numbers the warps that own rows (index of their partial sums), `n_wid` = their number.
constexpr int SWEEP_WPB = 8;      // warps per CTA (256 threads)
#ifndef SWEEP_MINB
#define SWEEP_MINB 4               // resident CTAs per SM the persistent solve is compiled for (64 registers, no spills; 5, 6 and 8 measured: no gain)
#endif
struct SweepMap {
    int c, cs, first, step, end, wid; bool on, lane_ok;
    __device__ __forceinline__ SweepMap(int n_rows, int B, int n_row_groups, int mode, int cpl = 1) {
        const int nchunk = (B + 32 * cpl - 1) / (32 * cpl), w = threadIdx.x >> 5;
        const int lane = threadIdx.x & 31;
        int chunk;
        if (mode == 1) {
            chunk = blockIdx.x % nchunk;
            const int g = blockIdx.x / nchunk;
            on = g < n_row_groups;
            wid = g * SWEEP_WPB + w;
            first = wid; step = n_row_groups * SWEEP_WPB; end = n_rows;
        } else {
            const int gw = blockIdx.x * SWEEP_WPB + w, nblk = n_row_groups * SWEEP_WPB;
            chunk = gw % nchunk;
            wid = gw / nchunk;
            on = wid < nblk;
            first = (int)((long long)wid * n_rows / nblk);
            end = on ? (int)((long long)(wid + 1) * n_rows / nblk) : first;
            step = 1;
        }
        c = (chunk * 32 + lane) * cpl;       // cpl consecutive cases per lane (cpl = 2, 4: B a multiple of it, 16-byte accesses)
        lane_ok = c < B;
        cs = lane_ok ? c : B - cpl;          // lanes beyond the last case read valid memory and store nothing
    }
};
__host__ __device__ __forceinline__ int sweep_row_groups(int n_ctas, int B, int n_rows, int cpl = 1) {
    const int nchunk = (B + 32 * cpl - 1) / (32 * cpl);
    const int g = n_ctas / nchunk, gmax = (n_rows + SWEEP_WPB - 1) / SWEEP_WPB;
    return g < 1 ? 1 : (g > gmax ? gmax : g);
}

// CPL consecutive cases per lane (1, 2 or 4; the number of cases must be a multiple of CPL): with 2 or 4 every vector access is a
// 16-byte load / store, so a row's index / value requests are shared by 64 / 128 cases instead of 32 and fewer load instructions move
// the same data - the sweep kernels are bound by the load/store unit (measured, profiles/r2w_*, r2x_*: 3.97e9 -> 5.16e9 -> 5.39e9
// dof-updates/s from one to two to four cases per lane on 1024 cases, 6.4e9 on 4096).
template <int CPL> struct CaseVec { double v[CPL]; };
template <int CPL> __device__ __forceinline__ CaseVec<CPL> ld_cases(const double* p) {
    CaseVec<CPL> r;
    if (CPL == 1) { r.v[0] = p[0]; }
    else {
#pragma unroll
        for (int i = 0; i < CPL; i += 2) { const double2 t = *reinterpret_cast<const double2*>(p + i); r.v[i] = t.x; r.v[i + 1] = t.y; }
    }
    return r;
}
template <int CPL> __device__ __forceinline__ void st_cases(double* p, const CaseVec<CPL>& r) {
    if (CPL == 1) { p[0] = r.v[0]; }
    else {
#pragma unroll
        for (int i = 0; i < CPL; i += 2) *reinterpret_cast<double2*>(p + i) = make_double2(r.v[i], r.v[i + 1]);
    }
}
// sum_j A[r][j] x[idx_j][c .. c + CPL) for the cases of a lane: every lane reads the row's indices and values itself (warp-uniform
// addresses: broadcast loads that hit in L1).  Tried and measured slower (profiles/r2s_*, r2t_*): loading them once per warp (lane j
// takes entry j) and broadcasting by shuffles - 24 shuffles per four entries cost more than the 14 broadcast loads per row they
// replace; two rows in flight per warp (alternating gathers, 80 registers or spills).
template <int CPL>
__device__ __forceinline__ CaseVec<CPL> sweep_row_sum(const Csr& A, int r, const double* __restrict__ x, int B, int cs) {
    CaseVec<CPL> v;
#pragma unroll
    for (int i = 0; i < CPL; ++i) v.v[i] = 0.0;
    const int j1 = __ldg(&A.indptr[r + 1]);
    for (int j = __ldg(&A.indptr[r]); j < j1; ++j) {
        const double w = __ldg(&A.val[j]);
        const CaseVec<CPL> xj = ld_cases<CPL>(x + (size_t)__ldg(&A.idx[j]) * B + cs);
#pragma unroll
        for (int i = 0; i < CPL; ++i) v.v[i] += w * xj.v[i];
    }
    return v;
}

// y[r][c] = alpha * cscale[c] * sum_j A[r][j] x[j][c]
template <int CPL>
__global__ void __launch_bounds__(256) csr_store_b_k(Csr A, int B, int n_row_groups, int mode, const double* __restrict__ x, double alpha,
                                                     const double* __restrict__ cscale, double* __restrict__ y) {
    const SweepMap mp(A.n_rows, B, n_row_groups, mode, CPL);
    if (!mp.on) return;
    double sc[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) sc[i] = alpha * (cscale ? cscale[mp.cs + i] : 1.0);
    for (int r = mp.first; r < mp.end; r += mp.step) {
        CaseVec<CPL> v = sweep_row_sum<CPL>(A, r, x, B, mp.cs);
#pragma unroll
        for (int i = 0; i < CPL; ++i) v.v[i] *= sc[i];
        if (mp.lane_ok) st_cases<CPL>(y + (size_t)r * B + mp.c, v);
    }
}

// part[wid][c] = sum over the rows of warp wid of x[r][c] (A x)[r][c]
template <int CPL>
__global__ void __launch_bounds__(256) csr_dot_b_k(Csr A, int B, int n_row_groups, int mode, const double* __restrict__ x, double* __restrict__ part) {
    const SweepMap mp(A.n_rows, B, n_row_groups, mode, CPL);
    if (!mp.on) return;
    CaseVec<CPL> s;
#pragma unroll
    for (int i = 0; i < CPL; ++i) s.v[i] = 0.0;
    for (int r = mp.first; r < mp.end; r += mp.step) {
        const CaseVec<CPL> xr = ld_cases<CPL>(x + (size_t)r * B + mp.cs);
        const CaseVec<CPL> v = sweep_row_sum<CPL>(A, r, x, B, mp.cs);
#pragma unroll
        for (int i = 0; i < CPL; ++i) s.v[i] += xr.v[i] * v.v[i];
    }
    if (mp.lane_ok) st_cases<CPL>(part + (size_t)mp.wid * B + mp.c, s);
}
// ctl[c].red[slot] = scale * cscale[c] * sum_wid part[wid][c]   (fixed order; n_row_blocks = n_row_groups * 8 warps)
__global__ void dot_b_finish_k(int B, int n_row_blocks, const double* __restrict__ part, double scale, const double* __restrict__ cscale, Ctl* ctl, int slot) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B) return;
    double s = 0.0;
    for (int rb = 0; rb < n_row_blocks; ++rb) s += part[(size_t)rb * B + c];
    ctl[c].red[slot] = scale * (cscale ? cscale[c] : 1.0) * s;
}

// ctl[c].flux[slot] = sum_i b_L[i] q[port_row[i]][c]   (q == nullptr: the current Chebyshev iterate of solve `which`)
__global__ void flux_b_k(int nport, const int* __restrict__ pe, const double* __restrict__ bl, const double* q, TriBuf tb, Ctl* ctl,
                         int which, int slot, int B) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= B) return;
    const double* qs = q ? q : tb.b[ctl->sc[which].cur];
    double s = 0.0;
    for (int i = 0; i < nport; ++i) s += bl[i] * qs[(size_t)pe[i] * B + c];
    ctl[c].flux[slot] = s;
}
// b[port_row[i]][c] -= coef * cscale[c] * b_L[i] * p_left[c]     (the ds(0) term of the q rows, wave_2D.py:469,480,504,515)
__global__ void port_rhs_b_k(int nport, const int* __restrict__ pe, const double* __restrict__ bl, const double* __restrict__ cscale,
                             double* b, double coef, const Ctl* ctl, int B) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nport * B) return;
    const int i = (int)(t / B), c = (int)(t % B);
    b[(size_t)pe[i] * B + c] -= coef * (cscale ? cscale[c] : 1.0) * bl[i] * ctl[c].pleft;
}

// All Chebyshev iterations of M x = b for the B cases at once, in ONE cooperative launch (replaces the B solve(A, x, b) calls of a
// sweep, wave_2D.py:820,854).  Same recurrence, stop test and accepted iterate as solve_coop_csr_k; pinv is per row (the matrix is
// the same for every case).
struct CsrSolveBArgs {
    Csr A; int B, n_row_groups, mode; TriBuf tb; const double* b; const double* pinv; double cheb_d, cheb_c, tol2; int max_iter, which;
    double* partials; Ctl* ctl;
};
#ifndef SWEEP_MINB4
#define SWEEP_MINB4 3
#endif
template <int CPL>
__global__ void __launch_bounds__(256, CPL == 4 ? SWEEP_MINB4 : SWEEP_MINB) solve_coop_csr_b_k(const __grid_constant__ CsrSolveBArgs a) {
    __shared__ double red[64];
    __shared__ double bc[2];
    SolveCtl& sc = a.ctl->sc[a.which];
    int k = 0, cur = sc.cur;
    double rho = 0.0, rr = 0.0, bb = 0.0;
    bool conv = false;
    unsigned int epoch = 0;
    const SweepMap mp(a.A.n_rows, a.B, a.n_row_groups, a.mode, CPL);
    const int B = a.B;
    while (true) {
        double c1, c2, rho_k;
        cheb_coeffs(k, rho, a.cheb_d, a.cheb_c, c1, c2, rho_k);
        const int nxt = (cur + 1) % 3, prv = (cur + 2) % 3;
        const double* x = tb_sel(a.tb, cur); const double* xp = tb_sel(a.tb, prv); double* xn = tb_sel(a.tb, nxt);
        double r0 = 0.0, b0 = 0.0;
        if (mp.on) {                                                       // uniform per warp
            for (int r = mp.first; r < mp.end; r += mp.step) {
                const size_t o = (size_t)r * B + mp.cs;
                const CaseVec<CPL> bv = ld_cases<CPL>(a.b + o), xc = ld_cases<CPL>(x + o);   // streaming operands requested before the row sum
                CaseVec<CPL> xpv;
                if (k > 0) xpv = ld_cases<CPL>(xp + o);
                const double pv = __ldg(&a.pinv[r]);
                const CaseVec<CPL> v = sweep_row_sum<CPL>(a.A, r, x, B, mp.cs);
                CaseVec<CPL> w;
                double rs = 0.0, bs = 0.0;
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    const double res = bv.v[i] - v.v[i];
                    w.v[i] = xc.v[i] + c2 * pv * res;
                    if (k > 0) w.v[i] += c1 * (xc.v[i] - xpv.v[i]);
                    rs += res * res; bs += bv.v[i] * bv.v[i];
                }
                if (mp.lane_ok) {
                    st_cases<CPL>(xn + o, w);
                    r0 += pv * rs; b0 += pv * bs;
                }
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

}  // namespace phw
