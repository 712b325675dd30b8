// kernels_ho.cuh - matrix-free kernels of the higher-order pairs P_k x RT_k, k <= 3 (basis_order of wave_2D.py:15-20,174-175,202-203;
// the reference's analytical paper cases main_wave_2D.py:53-114 and its eigenvalue study main_eigen.py:17-23 run them).
//
// Nothing is assembled.  Host side (porthamiltonian_fem_b200/fem_ref.py): every cell with its vertices in ascending global order,
// bases mapped from ONE reference triangle (P_k affine, RT_k contravariant Piola), so that the element matrices of the forms
// wave_2D.py:449-526 are reference matrices times a few numbers per cell:
//     M_p,K = |det J| Mp_hat      C_K = sign(det J) C_hat      M_q,K = G00 A00 + G01 (A01 + A10) + G11 A11,  G = J^T J / |det J|.
// Device side: per cell 3 vertex ids + 3 edge ids + 4 doubles (56 B; the assembled CSR rows of P3 x RT3 are ~2.5 KB per cell), the
// reference tabulations (<= 7.4 KB) in shared memory, every dof index derived from the ids.  Sixteen lanes per cell: gather the local
// vector (one entry per lane), form one output row per lane from the reference matrices, scatter with fp64
// atomics (RED.ADD.F64 resolves in L2; these meshes live there).  The boundary terms (zero-flux N, left port b_L) touch the dofs of
// boundary edges only and come as a compact CSR over those rows.  (Sixteen lanes per cell since the first GPU run: one thread per
// cell left a P3 x RT3 solve iteration at 33 us - ~900 dependent shared-memory reads per thread - behind the assembled CSR rows.)
//
// The element contraction is a dense batched product - (ndq x 3 ndq) reference block times the G-scaled local vectors of all cells -
// i.e. the one place of this code base where a tensor-core formulation exists (DMMA m8n8k4).  At 3 x 15 x 15 FMAs per 56 + 240 B
// of traffic (2.3 flop/B at k = 3) it stays below the fp64 ridge of the B200 (~5 flop/B), and on the reference's meshes (<= 10^5
// dofs) a solve iteration is bounded by the two grid barriers, not by the FMA pipe: see DESIGN.md section 11.
#pragma once

namespace phw {

struct HoMf {
    int nc, nv, ne, kp, kq, ndp, ndq, n_p, n_q;
    const int* cv;          // [nc][3] vertex ids, ascending
    const int* ce;          // [nc][3] edge ids (edge i opposite local vertex i)
    const double* geom;     // [nc][4] det J (signed), G00, G01, G11
    const double* ref;      // reference matrices, row-major: Mp_hat [ndp x ndp] | C_hat [ndq x ndp] | A00 | A01 + A10 | A11 [ndq x ndq]
    int off_C, off_A00, off_A01, off_A11, ref_len;
};
// compact CSR over the few rows a boundary operator touches: y[row[r]] += alpha * sum_j val[j] x[idx[j]]
struct RowCsr { int n_rows; const int* row; const int* indptr; const int* idx; const double* val; };

enum { HO_MP = 0, HO_MQ = 1, HO_C = 2, HO_CT = 3 };
constexpr int HO_NT = 256;        // threads per CTA of the cell loops
constexpr int HO_NT_SOLVE_MAX = 1024;   // the persistent solve picks its CTA size at launch (fewer, larger CTAs = cheaper grid barriers)
constexpr int HO_LPC = 16;        // lanes per cell (>= the largest local dimension, RT_3: 15; divides the warp size)
constexpr int HO_MAXD = 15;       // RT_3: 15 dofs per cell

__device__ __forceinline__ int ho_dof_p(const HoMf& m, int c, const int* v, const int* e, int i) {
    if (i < 3) return v[i];
    const int nde = m.kp - 1;
    i -= 3;
    if (i < 3 * nde) return m.nv + e[i / nde] * nde + i % nde;
    const int ndi = m.ndp - 3 - 3 * nde;
    return m.nv + m.ne * nde + c * ndi + (i - 3 * nde);
}
__device__ __forceinline__ int ho_dof_q(const HoMf& m, int c, const int* e, int i) {
    const int k = m.kq;
    if (i < 3 * k) return e[i / k] * k + i % k;
    return m.ne * k + c * (m.ndq - 3 * k) + (i - 3 * k);
}
__device__ __forceinline__ void ho_load_ref(const HoMf& m, double* sref) {
    for (int i = threadIdx.x; i < m.ref_len; i += blockDim.x) sref[i] = m.ref[i];
    __syncthreads();
}

// y += alpha * A x over all cells: HO_LPC = 16 lanes per cell (lane i gathers local entry i, then forms output row i), so the lanes
// of a cell share a warp and __syncwarp orders the hand-over through sx ([cells per CTA][HO_LPC] doubles of shared memory).  The
// reference rows are read at a stride of nd <= 15 doubles: the 16 lanes of a cell hit 16 different bank pairs, the second cell of
// the warp reads the same words (broadcast).  One pass covers gridDim.x * blockDim.x / 16 cells.
template <int OP>
__device__ __forceinline__ void ho_cells_apply(const HoMf& m, const double* sref, double* sx, const double* x, double* y, double alpha) {
    const int nin = (OP == HO_MP || OP == HO_C) ? m.ndp : m.ndq;
    const int nout = (OP == HO_MP || OP == HO_CT) ? m.ndp : m.ndq;
    const int cs = threadIdx.x / HO_LPC, l = threadIdx.x % HO_LPC, cpb = blockDim.x / HO_LPC;
    double* sxc = sx + cs * HO_LPC;
    for (int base = blockIdx.x * cpb; base < m.nc; base += gridDim.x * cpb) {
        const int c = base + cs;
        const bool ok = c < m.nc;
        int v[3] = {0, 0, 0}, e[3] = {0, 0, 0};
        double dj = 0.0, g00 = 0.0, g01 = 0.0, g11 = 0.0;
        if (ok) {
#pragma unroll
            for (int i = 0; i < 3; ++i) { v[i] = m.cv[3 * c + i]; e[i] = m.ce[3 * c + i]; }
            dj = m.geom[4 * c]; g00 = m.geom[4 * c + 1]; g01 = m.geom[4 * c + 2]; g11 = m.geom[4 * c + 3];
        }
        __syncwarp();                                                 // the rows of the previous pass have been formed
        if (ok && l < nin) sxc[l] = x[(OP == HO_MP || OP == HO_C) ? ho_dof_p(m, c, v, e, l) : ho_dof_q(m, c, e, l)];
        __syncwarp();
        if (ok && l < nout) {
            const double sc = OP == HO_MP ? fabs(dj) * alpha : ((OP == HO_MQ) ? alpha : (dj < 0.0 ? -alpha : alpha));
            double acc;
            if (OP == HO_MP) {
                const double* A = sref + l * m.ndp;
                acc = 0.0;
                for (int j = 0; j < nin; ++j) acc += A[j] * sxc[j];
            } else if (OP == HO_MQ) {
                const double *A0 = sref + m.off_A00 + l * m.ndq, *A1 = sref + m.off_A01 + l * m.ndq, *A2 = sref + m.off_A11 + l * m.ndq;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                for (int j = 0; j < nin; ++j) { const double xj = sxc[j]; s0 += A0[j] * xj; s1 += A1[j] * xj; s2 += A2[j] * xj; }
                acc = g00 * s0 + g01 * s1 + g11 * s2;
            } else if (OP == HO_C) {
                const double* A = sref + m.off_C + l * m.ndp;
                acc = 0.0;
                for (int j = 0; j < nin; ++j) acc += A[j] * sxc[j];
            } else {
                const double* A = sref + m.off_C + l;                 // column l of C_hat
                acc = 0.0;
                for (int j = 0; j < nin; ++j) acc += A[j * m.ndp] * sxc[j];
            }
            atomicAdd(&y[(OP == HO_MP || OP == HO_CT) ? ho_dof_p(m, c, v, e, l) : ho_dof_q(m, c, e, l)], sc * acc);
        }
    }
}

template <int OP>
__global__ void __launch_bounds__(HO_NT) ho_apply_k(const HoMf m, const double* x, double* y, double alpha) {
    extern __shared__ double sm_[];
    double* sref = sm_;
    double* sx = sm_ + ((m.ref_len + 1) & ~1);
    ho_load_ref(m, sref);
    ho_cells_apply<OP>(m, sref, sx, x, y, alpha);
}

// diagonal of M_p (OP = HO_MP) or M_q (HO_MQ), accumulated with atomics (setup: Jacobi weights)
template <int OP>
__global__ void __launch_bounds__(HO_NT) ho_diag_k(const HoMf m, double* d) {
    extern __shared__ double sm_[];
    ho_load_ref(m, sm_);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < m.nc; c += gridDim.x * blockDim.x) {
        int v[3], e[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { v[i] = m.cv[3 * c + i]; e[i] = m.ce[3 * c + i]; }
        const double dj = fabs(m.geom[4 * c]), g00 = m.geom[4 * c + 1], g01 = m.geom[4 * c + 2], g11 = m.geom[4 * c + 3];
        if (OP == HO_MP) {
            for (int i = 0; i < m.ndp; ++i) atomicAdd(&d[ho_dof_p(m, c, v, e, i)], dj * sm_[i * m.ndp + i]);
        } else {
            for (int i = 0; i < m.ndq; ++i)
                atomicAdd(&d[ho_dof_q(m, c, e, i)], g00 * sm_[m.off_A00 + i * m.ndq + i] + g01 * sm_[m.off_A01 + i * m.ndq + i] + g11 * sm_[m.off_A11 + i * m.ndq + i]);
        }
    }
}

// ctl->red[slot] = scale * x^T A x (A = M_p or M_q): cell-local quadratic forms, per-block partials, summed in block order by the last
// block to finish (no scatter: the energies need no atomics on data)
template <int OP>
__global__ void __launch_bounds__(HO_NT) ho_quad_k(const HoMf m, const double* x, double scale, Ctl* ctl, int slot, double* partials, int counter_id) {
    extern __shared__ double sm_[];
    __shared__ double red[64];
    __shared__ bool is_last;
    double* sref = sm_;
    double* sx = sm_ + ((m.ref_len + 1) & ~1);
    ho_load_ref(m, sref);
    const int nd = OP == HO_MP ? m.ndp : m.ndq;
    const int cs = threadIdx.x / HO_LPC, l = threadIdx.x % HO_LPC, cpb = blockDim.x / HO_LPC;
    double* sxc = sx + cs * HO_LPC;
    double s = 0.0, z = 0.0;
    for (int base = blockIdx.x * cpb; base < m.nc; base += gridDim.x * cpb) {
        const int c = base + cs;
        const bool ok = c < m.nc;
        int v[3] = {0, 0, 0}, e[3] = {0, 0, 0};
        double dj = 0.0, g00 = 0.0, g01 = 0.0, g11 = 0.0;
        if (ok) {
#pragma unroll
            for (int i = 0; i < 3; ++i) { v[i] = m.cv[3 * c + i]; e[i] = m.ce[3 * c + i]; }
            dj = fabs(m.geom[4 * c]); g00 = m.geom[4 * c + 1]; g01 = m.geom[4 * c + 2]; g11 = m.geom[4 * c + 3];
        }
        __syncwarp();
        if (ok && l < nd) sxc[l] = x[OP == HO_MP ? ho_dof_p(m, c, v, e, l) : ho_dof_q(m, c, e, l)];
        __syncwarp();
        if (ok && l < nd) {                                           // lane l: x_l (A x)_l of this cell
            const double xi = sxc[l];
            if (OP == HO_MP) {
                double a = 0.0;
                for (int j = 0; j < nd; ++j) a += sref[l * nd + j] * sxc[j];
                s += dj * xi * a;
            } else {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                for (int j = 0; j < nd; ++j) {
                    const double xj = sxc[j];
                    s0 += sref[m.off_A00 + l * nd + j] * xj; s1 += sref[m.off_A01 + l * nd + j] * xj; s2 += sref[m.off_A11 + l * nd + j] * xj;
                }
                s += xi * (g00 * s0 + g01 * s1 + g11 * s2);
            }
        }
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
    double tt = 0.0; z = 0.0;
    for (int p = threadIdx.x; p < (int)gridDim.x; p += blockDim.x) tt += __ldcg(&partials[p]);
    block_sum2(tt, z, red);
    if (threadIdx.x == 0) { ctl->counters[counter_id] = 0; ctl->red[slot] = scale * tt; }
}

__global__ void row_csr_axpy_k(const RowCsr A, const double* x, double alpha, double* y) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= A.n_rows) return;
    double s = 0.0;
    for (int j = A.indptr[r]; j < A.indptr[r + 1]; ++j) s += A.val[j] * x[A.idx[j]];
    y[A.row[r]] += alpha * s;
}

// All Chebyshev iterations of M_p x = b (which = 0) or M_q x = b (1) in ONE cooperative launch (replaces solve(A, x, b),
// wave_2D.py:820,854, for the higher-order pairs).  Per iteration: cells scatter A x_k into yacc (zero on entry), grid barrier, owned
// rows form the residual, x_{k+1} and the norms and clear yacc, grid barrier + reduction.  The accepted iterate is the one whose
// residual was measured, as in the P1 x RT1 kernels.
struct HoSolveArgs {
    HoMf m; TriBuf tb; const double* b; const double* pinv; double* yacc;
    double cheb_d, cheb_c, tol2; int max_iter, which, n;
    double* partials; Ctl* ctl;
};
__global__ void __launch_bounds__(HO_NT_SOLVE_MAX) ho_solve_k(const __grid_constant__ HoSolveArgs a) {
    extern __shared__ double sm_[];
    __shared__ double red[64];
    __shared__ double bc[2];
    double* sref = sm_;
    double* sx = sm_ + ((a.m.ref_len + 1) & ~1);
    ho_load_ref(a.m, sref);
    SolveCtl& sc = a.ctl->sc[a.which];
    int k = 0, cur = sc.cur;
    double rho = 0.0, rr = 0.0, bb = 0.0;
    bool conv = false;
    unsigned int epoch = 0;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gnt = gridDim.x * blockDim.x;
    while (true) {
        double c1, c2, rho_k;
        cheb_coeffs(k, rho, a.cheb_d, a.cheb_c, c1, c2, rho_k);
        const int nxt = (cur + 1) % 3, prv = (cur + 2) % 3;
        const double* x = tb_sel(a.tb, cur); const double* xp = tb_sel(a.tb, prv); double* xn = tb_sel(a.tb, nxt);
        if (a.which == 0) ho_cells_apply<HO_MP>(a.m, sref, sx, x, a.yacc, 1.0);
        else ho_cells_apply<HO_MQ>(a.m, sref, sx, x, a.yacc, 1.0);
        grid_barrier(&a.ctl->gbar[0], epoch);
        double r0 = 0.0, b0 = 0.0;
        for (int i = gtid; i < a.n; i += gnt) {
            const double val = __ldcg(&a.yacc[i]);
            a.yacc[i] = 0.0;
            const double bv = a.b[i], pv = a.pinv[i], xc = x[i];
            const double res = bv - val;
            double v = xc + c2 * pv * res;
            if (k > 0) v += c1 * (xc - xp[i]);
            xn[i] = v;
            r0 += pv * res * res; b0 += pv * bv * bv;
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
