// kernels_sweep.cuh - the batched parameter sweep (BASELINE configs[4]) with the CASE INDEX FASTEST.
//
// The B cases of a sweep share the mesh, hence M_p, M_q, D and b_L; only the wave speed / density (scalar factors of the coupling
// terms) and the lumped constants differ (the loop over cases of main_wave_2D.py:24-50,119-140).  Round 1 put the cases side by side
// as B copies of the mesh inside one device mesh, which re-reads identical index / metric / weight data B times - about half of the
// bytes of an M_q iteration.  Here every vector is an [n_dof][B] array and the operators are ONE set of assembled CSR rows
// (5 non-zeros per RT1 row, ~7 per P1 row: a few hundred KB, cache resident): a warp owns 32 consecutive cases and a strided set of
// rows (SweepMap below); the row's indices and values are the same for all lanes (broadcast loads) and every vector access
// x[idx * B + c] is a fully coalesced 256-byte segment.  The rows a mesh row gathers from are worked on at the same time by the other
// warps of the case chunk, so they hit in L1 / L2 and DRAM sees each vector once per pass: 32 B per dof, case and Chebyshev iteration (x_k, x_{k-1}, b read; x_{k+1} written) instead of 64 (M_q) and
// 80 (M_p) in the replicated layout.  Sparse matrix x dense block: no shared-memory staging, no atomics, fixed summation order.
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
        c = (chunk * 32 + lane) * cpl;       // cpl consecutive cases per lane (cpl = 2: B even, 16-byte accesses)
        lane_ok = c < B;
        cs = lane_ok ? c : B - cpl;          // lanes beyond the last case read valid memory and store nothing
    }
};
__host__ __device__ __forceinline__ int sweep_row_groups(int n_ctas, int B, int n_rows, int cpl = 1) {
    const int nchunk = (B + 32 * cpl - 1) / (32 * cpl);
    const int g = n_ctas / nchunk, gmax = (n_rows + SWEEP_WPB - 1) / SWEEP_WPB;
    return g < 1 ? 1 : (g > gmax ? gmax : g);
}

// sum_j A[r][j] x[idx_j][c] for the 32 cases of a warp: every lane reads the row's indices and values itself (warp-uniform addresses:
// broadcast loads that hit in L1).  Tried and measured slower (profiles/r2s_*): loading them once per warp (lane j takes entry j) and
// broadcasting by shuffles - 24 shuffles per four entries cost more than the 14 broadcast loads per row they replace (3.1e9 instead of
// 4.0e9 dof-updates/s on the 1024-case sweep); two rows in flight per warp (alternating gathers, 80 registers or spills): 3.1-3.4e9
// (profiles/r2t_*).
__device__ __forceinline__ double sweep_row_sum(const Csr& A, int r, const double* __restrict__ x, int B, int cs) {
    double v = 0.0;
    const int j1 = __ldg(&A.indptr[r + 1]);
    for (int j = __ldg(&A.indptr[r]); j < j1; ++j) v += __ldg(&A.val[j]) * x[(size_t)__ldg(&A.idx[j]) * B + cs];
    return v;
}

// y[r][c] = alpha * cscale[c] * sum_j A[r][j] x[j][c]
__global__ void __launch_bounds__(256) csr_store_b_k(Csr A, int B, int n_row_groups, int mode, const double* __restrict__ x, double alpha,
                                                     const double* __restrict__ cscale, double* __restrict__ y) {
    const SweepMap mp(A.n_rows, B, n_row_groups, mode);
    if (!mp.on) return;
    const double sc = alpha * (cscale ? cscale[mp.cs] : 1.0);
    for (int r = mp.first; r < mp.end; r += mp.step) {
        const double v = sweep_row_sum(A, r, x, B, mp.cs);
        if (mp.lane_ok) y[(size_t)r * B + mp.c] = sc * v;
    }
}

// part[wid][c] = sum over the rows of warp wid of x[r][c] (A x)[r][c]
__global__ void __launch_bounds__(256) csr_dot_b_k(Csr A, int B, int n_row_groups, int mode, const double* __restrict__ x, double* __restrict__ part) {
    const SweepMap mp(A.n_rows, B, n_row_groups, mode);
    if (!mp.on) return;
    double s = 0.0;
    for (int r = mp.first; r < mp.end; r += mp.step) {
        const double xr = x[(size_t)r * B + mp.cs];
        s += xr * sweep_row_sum(A, r, x, B, mp.cs);
    }
    if (mp.lane_ok) part[(size_t)mp.wid * B + mp.c] = s;
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
__global__ void __launch_bounds__(256, SWEEP_MINB) solve_coop_csr_b_k(const __grid_constant__ CsrSolveBArgs a) {
    __shared__ double red[64];
    __shared__ double bc[2];
    SolveCtl& sc = a.ctl->sc[a.which];
    int k = 0, cur = sc.cur;
    double rho = 0.0, rr = 0.0, bb = 0.0;
    bool conv = false;
    unsigned int epoch = 0;
    const SweepMap mp(a.A.n_rows, a.B, a.n_row_groups, a.mode);
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
                const double bv = a.b[o], xc = x[o];                       // streaming operands requested before the row sum
                const double xpv = k > 0 ? xp[o] : 0.0;
                const double pv = __ldg(&a.pinv[r]);
                const double v = sweep_row_sum(a.A, r, x, B, mp.cs);
                const double res = bv - v;
                double w = xc + c2 * pv * res;
                if (k > 0) w += c1 * (xc - xpv);
                if (mp.lane_ok) {
                    xn[o] = w;
                    r0 += pv * res * res; b0 += pv * bv * bv;
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

// The same solve with TWO consecutive cases per lane (B even): every vector access is a 16-byte load / store (512 B per warp request), so
// a row's index / value requests are shared by 64 cases instead of 32 and half as many load instructions move the same data.
#ifndef SWEEP_MINB2
#define SWEEP_MINB2 4
#endif
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__global__ void __launch_bounds__(256, SWEEP_MINB2) solve_coop_csr_b2_k(const __grid_constant__ CsrSolveBArgs a) {
    __shared__ double red[64];
    __shared__ double bc[2];
    SolveCtl& sc = a.ctl->sc[a.which];
    int k = 0, cur = sc.cur;
    double rho = 0.0, rr = 0.0, bb = 0.0;
    bool conv = false;
    unsigned int epoch = 0;
    const SweepMap mp(a.A.n_rows, a.B, a.n_row_groups, a.mode, 2);
    const int B = a.B;
    while (true) {
        double c1, c2, rho_k;
        cheb_coeffs(k, rho, a.cheb_d, a.cheb_c, c1, c2, rho_k);
        const int nxt = (cur + 1) % 3, prv = (cur + 2) % 3;
        const double* x = tb_sel(a.tb, cur); const double* xp = tb_sel(a.tb, prv); double* xn = tb_sel(a.tb, nxt);
        double r0 = 0.0, b0 = 0.0;
        if (mp.on) {
            for (int r = mp.first; r < mp.end; r += mp.step) {
                const size_t o = (size_t)r * B + mp.cs;
                const double2 bv = ld2(a.b + o), xc = ld2(x + o);
                const double2 xpv = k > 0 ? ld2(xp + o) : make_double2(0.0, 0.0);
                const double pv = __ldg(&a.pinv[r]);
                double v0 = 0.0, v1 = 0.0;
                const int j1 = __ldg(&a.A.indptr[r + 1]);
                for (int j = __ldg(&a.A.indptr[r]); j < j1; ++j) {
                    const double wj = __ldg(&a.A.val[j]);
                    const double2 xj = ld2(x + (size_t)__ldg(&a.A.idx[j]) * B + mp.cs);
                    v0 += wj * xj.x; v1 += wj * xj.y;
                }
                const double res0 = bv.x - v0, res1 = bv.y - v1;
                double2 w = make_double2(xc.x + c2 * pv * res0, xc.y + c2 * pv * res1);
                if (k > 0) { w.x += c1 * (xc.x - xpv.x); w.y += c1 * (xc.y - xpv.y); }
                if (mp.lane_ok) {
                    *reinterpret_cast<double2*>(xn + o) = w;
                    r0 += pv * res0 * res0; b0 += pv * bv.x * bv.x;
                    r0 += pv * res1 * res1; b0 += pv * bv.y * bv.y;
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
