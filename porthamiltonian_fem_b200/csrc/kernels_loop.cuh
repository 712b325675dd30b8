// kernels_loop.cuh - the WHOLE time loop of wave_2D.py:787-982 in ONE launch, for meshes that live in the caches (the reference's
// own workloads: main_wave_2D.py:24-36,92-102, main_wave_2D_control.py:38-39, and every case of a batched sweep).
//
// The multi-launch path (phwave.cu: step()) enqueues 16-21 kernels per step and every Chebyshev iteration of its cooperative solves
// crosses a grid-wide barrier of several hundred CTAs: on a 10-40 k-dof mesh a step costs 280-470 us of pure latency.  Here one
// CTA owns one patch for the whole run; the CTAs of a case form ONE thread-block cluster (<= 16 CTAs, hardware cluster barrier with
// release / acquire semantics instead of global atomics) or, for larger meshes, one co-resident grid with the one-atomic-per-CTA
// barrier.  All vectors stay in global memory, i.e. in L2; a step is the same sequence of operations as step() - right-hand sides,
// Chebyshev solves, port flux, the lumped ODE / controller (scalar_apply), state updates, energies, H_array row - executed by the
// same row passes (vertex_pass / edge_pass / vertex_ell_pass of kernels.cuh), separated by cluster barriers instead of kernel
// boundaries.  The host enqueues ONE kernel for n steps.  A batched sweep runs one independent cluster per case: no barrier and no
// reduction ever spans two cases.
//
// Every scheme and variant of the multi-launch path except the analytical diagnostics: EE, SE, EH, SV, IE, SM; weak and strong
// Dirichlet; lumped model (IC); passivity and casimir control.
//
// RES (round 2, second half): the two mass solves - 85-90 % of a step - run out of SHARED MEMORY.  At the start of the launch every
// CTA copies the matrix data of its patch (assembled ELL rows of M_p, packed 32-byte cell records + edge adjacency + Jacobi weights
// of M_q) into its own shared memory, where it stays for all n steps.  A solve keeps the iterate (owned rows + ghost copies), the
// previous iterate and the right-hand side in shared memory as well; x_{k+1} overwrites x_{k-1} in place.  After each iteration
// ONE hardware cluster barrier separates "owners have written x_{k+1}" from "neighbours pull their ghost copies and the per-CTA
// residual norms out of the owners' shared memory" (ld.shared::cluster through mapa - distributed shared memory); nothing of an
// iteration touches L2.  Right-hand sides, energies, state updates and the extrapolation history stay in global memory (L2):
// they run once per step, not once per iteration.
#pragma once

namespace phw {

// byte offsets into the dynamic shared memory of loop_k<.., RES = true>; identical in every CTA (maxima over the patches), so that
// a neighbour can address this CTA's buffers from its own carve-up
struct ResCarve {
    int xq[2], bq;                // M_q solve (transient; aliases the scratch of the row passes): iterates [owned | ghosts], b
    int xp[2], bp;                // M_p solve (transient)
    int sio, swo, ids, w;         // persistent: ELL rows of M_p (slice offsets rebased to the patch, ids, weights)
    int wq, idq, dgq, dgp, me;    // persistent: assembled rows of M_q scaled by their diagonal, [4][me] weights and [4][me] uint16
                                  // local ids of the four neighbour edges, built by the CTA from the cell records; the diagonals
                                  // of M_q and M_p (the rows of M_p are scaled in place as well): needed for the norms only
    int gsv, gse;                 // persistent: ghost sources (owner CTA rank << 16 | row inside the owner's range)
    int part;                     // persistent: [2][2] doubles, this CTA's residual norms (double-buffered by iteration parity)
    int total;
};

#ifndef PHW_HOST_EMU
__device__ __forceinline__ unsigned dsmem_addr(const void* my_smem, unsigned cta_rank) {
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"((unsigned)__cvta_generic_to_shared(my_smem)), "r"(cta_rank));
    return ra;
}
__device__ __forceinline__ double ld_dsmem_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
#endif

struct LoopArgs {
    DevMesh m; TriBuf tp, tq;
    double *p, *q, *bp, *bq, *tmp_p, *tmp_q;
    const double *pinv_p, *pinv_q, *mdiag_p, *wp, *wq;
    double* hist[2][2][6];                    // [field][site][slot]
    int hist_head[2][2], hist_count[2][2], extrap[2], n_sites;
    int have_vq, sv_noreuse;
    const int *port_goff, *port_e; const double* port_bl; int nport;
    const double *gK, *ginvrho;               // per-case factors of a batched run (nullptr: one case)
    int n_dir; const int* dir_idx; const double* dir_w;
    const ScalarPrm* sp; double* Hrows; long long rows_per_group;
    Ctl* ctl; double* partials;               // partials: 4 * gridDim doubles
    int scheme, ic, strong, casimir, ell_p, G, cl;   // cl: CTAs (= patches) per case
    double dt, K, rho, blw, tol2, xscale;
    int max_iter;
    double cheb_d[3], cheb_c[3];
    long long n_steps;
    const uint4* cellq;           // RES: packed M_q cell records (phwave.cu: pack_cellq_k)
    ResCarve rc;
};

struct LoopCtx {
    int g, lead, ncta;            // case of this CTA, first CTA of the case, CTAs per case
    unsigned int epoch, nred;
    double* red; double* bc;      // shared memory: 64 + 2 doubles
    int v0, vn, e0, en;           // owned dof ranges of this CTA's patch
};

template <bool CLUSTER>
__device__ __forceinline__ void loop_sync(const LoopArgs& a, LoopCtx& cx) {
    if (CLUSTER) {
        __threadfence();
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        grid_barrier(&a.ctl->gbar[0], cx.epoch);
    }
}

// sums of (x, y) over all threads of the case, identical bits in every thread; contains one loop_sync
template <bool CLUSTER>
__device__ __forceinline__ void loop_reduce2(const LoopArgs& a, LoopCtx& cx, double& x, double& y) {
    block_sum2(x, y, cx.red);
    double* part = a.partials + (size_t)(cx.nred & 1u) * 2 * gridDim.x;
    cx.nred++;
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = x; part[2 * blockIdx.x + 1] = y; }
    loop_sync<CLUSTER>(a, cx);
    if (threadIdx.x < 32) {
        double s0 = 0.0, s1 = 0.0;
        for (int i = threadIdx.x; i < cx.ncta; i += 32) { s0 += __ldcg(&part[2 * (cx.lead + i)]); s1 += __ldcg(&part[2 * (cx.lead + i) + 1]); }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        if (threadIdx.x == 0) { cx.bc[0] = s0; cx.bc[1] = s1; }
    }
    __syncthreads();
    x = cx.bc[0]; y = cx.bc[1];
    __syncthreads();
}

__device__ __forceinline__ RowArgs loop_base_args(const LoopArgs& a) {
    RowArgs r{};
    r.m = a.m; r.tp = a.tp; r.tq = a.tq;
    r.partials = a.partials; r.ctl = a.ctl;
    r.tol2 = a.tol2; r.max_iter = a.max_iter; r.weight = 1.0; r.dot_scale = 1.0;
    r.mdiag = a.mdiag_p; r.xscale = a.xscale;
    return r;
}

// NT threads per CTA; gridDim.x == number of patches (one patch per CTA for the whole run)
template <bool CLUSTER, int NT, bool RES = false>
__global__ void __launch_bounds__(NT, (NT <= 256 && !RES) ? 2 : 1) loop_k(const __grid_constant__ LoopArgs a) {
    static_assert(!RES || CLUSTER, "the resident solves exchange through distributed shared memory");
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[64];
    __shared__ double bc[2];
    __shared__ PatchHdr myh;
    LoopCtx cx;
    cx.red = red; cx.bc = bc; cx.epoch = 0; cx.nred = 0;
    if ((int)blockIdx.x < a.m.npatch) load_hdr_smem(&myh, &a.m.hdr[blockIdx.x]);
    __syncthreads();
    cx.ncta = a.cl;
    cx.g = (int)blockIdx.x / a.cl;
    cx.lead = cx.g * a.cl;
    cx.v0 = myh.v_start; cx.vn = myh.v_cnt; cx.e0 = myh.e_start; cx.en = myh.e_cnt;
    const bool leader = (int)blockIdx.x == cx.lead;
    Ctl& ctl = a.ctl[cx.g];
    const ScalarPrm sp = a.sp[cx.g];
    const double dt = a.dt;
    const bool implicit = (a.scheme == PHW_IE || a.scheme == PHW_SM);
    int cur[3] = {ctl.sc[0].cur, ctl.sc[1].cur, ctl.sc[2].cur};
    int hh[2][2], hc[2][2];
    for (int f = 0; f < 2; ++f) for (int s_ = 0; s_ < 2; ++s_) { hh[f][s_] = a.hist_head[f][s_]; hc[f][s_] = a.hist_count[f][s_]; }
    bool have_vq = a.have_vq != 0;
    long long it_p = 0, it_q = 0, it_b = 0, n_p = 0, n_q = 0, n_b = 0, nonconv = 0;
    double maxres[3] = {0.0, 0.0, 0.0};
    const double Kg = a.gK ? a.gK[cx.g] : a.K;
    const double irg = a.ginvrho ? a.ginvrho[cx.g] : 1.0 / a.rho;

    // ---- RES: matrix data of this CTA's patch -> shared memory, once per launch --------------------------------------------------
    char* const sbase = reinterpret_cast<char*>(smem);
    if (RES) {
        const PatchHdr& h = myh;
        const int nsl = (h.v_cnt + 31) >> 5;
        uint32_t* s_io = reinterpret_cast<uint32_t*>(sbase + a.rc.sio);
        uint32_t* s_wo = reinterpret_cast<uint32_t*>(sbase + a.rc.swo);
        for (int i = threadIdx.x; i <= nsl; i += NT) {
            s_io[i] = a.m.ell_ioff[h.vs_off + i] - (uint32_t)h.ell_i0;
            s_wo[i] = a.m.ell_woff[h.vs_off + i] - (uint32_t)h.ell_w0;
        }
        {
            uint4* d = reinterpret_cast<uint4*>(sbase + a.rc.ids);
            const uint4* g = reinterpret_cast<const uint4*>(a.m.ell_ids + (uint32_t)h.ell_i0);
            for (uint32_t i = threadIdx.x; i < ((uint32_t)h.ell_in >> 3); i += NT) d[i] = g[i];
        }
        {
            double* d = reinterpret_cast<double*>(sbase + a.rc.w);
            const double* g = a.m.ell_w + (uint32_t)h.ell_w0;
            for (uint32_t i = threadIdx.x; i < (uint32_t)h.ell_wn; i += NT) d[i] = g[i];
        }
        __syncthreads();
        {   // Jacobi scaling of the M_p rows in place: w_vj / d_v with d_v = sum_j w_vj (the P1 mass diagonal is the row sum of
            // the off-diagonal entries); the iteration then needs neither the diagonal nor a division
            double* sw = reinterpret_cast<double*>(sbase + a.rc.w);
            double* dg = reinterpret_cast<double*>(sbase + a.rc.dgp);
            for (int lv = threadIdx.x; lv < h.v_cnt; lv += NT) {
                const int sl = lv >> 5, lane = lv & 31;
                const uint32_t wo0 = s_wo[sl];
                const int W = (int)((s_wo[sl + 1] - wo0) >> 5);
                double dsum = 0.0;
                for (int j = 0; j < W; ++j) dsum += sw[wo0 + j * 32 + lane];
                const double inv = 1.0 / dsum;
                for (int j = 0; j < W; ++j) sw[wo0 + j * 32 + lane] *= inv;
                dg[lv] = dsum;
            }
        }
        {   // assembled rows of M_q from the metric of the (<= 2) cells of every owned edge: m_ii = 3 S - 4 g_i, m_ij = s_i s_j (S - 4 g_k)
            double* wq = reinterpret_cast<double*>(sbase + a.rc.wq);
            uint16_t* idq = reinterpret_cast<uint16_t*>(sbase + a.rc.idq);
            double* dgq = reinterpret_cast<double*>(sbase + a.rc.dgq);
            const int me = a.rc.me;
            for (int le = threadIdx.x; le < h.e_cnt; le += NT) {
                const uint32_t adj = a.m.eadj[h.e_start + le];
                double diag = 0.0;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const uint32_t code = hf ? (adj >> 16) : (adj & 0xFFFFu);
                    double w1 = 0.0, w2 = 0.0; uint32_t i1 = (uint32_t)le, i2 = (uint32_t)le;
                    if (code != 0xFFFFu) {
                        const uint4 qa = a.cellq[2 * (size_t)(h.cell_off + (code >> 2))], qb = a.cellq[2 * (size_t)(h.cell_off + (code >> 2)) + 1];
                        const uint32_t e[3] = {qa.x & 0xFFFFu, qa.x >> 16, qa.y & 0xFFFFu};
                        const double g[3] = {__hiloint2double((int)qa.w, (int)qa.z), __hiloint2double((int)qb.y, (int)qb.x), __hiloint2double((int)qb.w, (int)qb.z)};
                        const double S = g[0] + g[1] + g[2];
                        const int i = (int)(code & 3u), j = (i + 1) % 3, kk = (i + 2) % 3;
                        const double si = (e[i] & 0x8000u) ? -1.0 : 1.0, sj = (e[j] & 0x8000u) ? -1.0 : 1.0, sk = (e[kk] & 0x8000u) ? -1.0 : 1.0;
                        diag += 3.0 * S - 4.0 * g[i];
                        w1 = si * sj * (S - 4.0 * g[kk]); i1 = e[j] & 0x3FFFu;
                        w2 = si * sk * (S - 4.0 * g[j]); i2 = e[kk] & 0x3FFFu;
                    }
                    wq[(2 * hf) * me + le] = w1; idq[(2 * hf) * me + le] = (uint16_t)i1;
                    wq[(2 * hf + 1) * me + le] = w2; idq[(2 * hf + 1) * me + le] = (uint16_t)i2;
                }
                const double inv = 1.0 / diag;
#pragma unroll
                for (int j = 0; j < 4; ++j) wq[j * me + le] *= inv;
                dgq[le] = diag;
            }
        }
        // ghost sources: which CTA of the cluster owns the dof, and which row of its range it is
        uint32_t* gsv = reinterpret_cast<uint32_t*>(sbase + a.rc.gsv);
        uint32_t* gse = reinterpret_cast<uint32_t*>(sbase + a.rc.gse);
        for (int i = threadIdx.x; i < h.gv_cnt + h.ge_cnt; i += NT) {
            const bool isv = i < h.gv_cnt;
            const int gid = isv ? a.m.ghost_v[h.gv_off + i] : a.m.ghost_e[h.ge_off + (i - h.gv_cnt)];
            uint32_t code = 0xFFFFFFFFu;
            for (int r = 0; r < cx.ncta; ++r) {
                const PatchHdr* hr = &a.m.hdr[cx.lead + r];
                const int st = isv ? hr->v_start : hr->e_start, cn = isv ? hr->v_cnt : hr->e_cnt;
                if (gid >= st && gid < st + cn) { code = ((uint32_t)r << 16) | (uint32_t)(gid - st); break; }
            }
            if (isv) gsv[i] = code; else gse[i - h.gv_cnt] = code;
        }
        __syncthreads();
    }

    // ---- small helpers (all threads of all CTAs of the case call them the same number of times) -----------------------------
    auto SYNC = [&]() { loop_sync<CLUSTER>(a, cx); };
    auto scalar = [&](int op, double aux0) {
        if (leader && threadIdx.x == 0) scalar_apply(ctl, sp, cx.g, op, a.Hrows, a.rows_per_group, aux0);
    };
    auto tbp = [&](int which) { return tb_sel(a.tp, cur[which]); };
    auto tbq = [&](int which) { return tb_sel(a.tq, cur[which]); };
    // b_L^T q of the case (leader CTA, fixed order)
    auto flux = [&](const double* qs, int slot) {
        if (leader) {
            double s = 0.0, z = 0.0;
            for (int i = a.port_goff[cx.g] + (int)threadIdx.x; i < a.port_goff[cx.g + 1]; i += NT) s += a.port_bl[i] * qs[a.port_e[i]];
            block_sum2(s, z, red);
            if (threadIdx.x == 0) ctl.flux[slot] = s;
            __syncthreads();
        }
    };
    // elementwise work on the rows this CTA owns
    auto axpy_v = [&](double* y, double al, const double* x) { for (int i = cx.v0 + (int)threadIdx.x; i < cx.v0 + cx.vn; i += NT) y[i] += al * x[i]; };
    auto axpy_e = [&](double* y, double al, const double* x) { for (int i = cx.e0 + (int)threadIdx.x; i < cx.e0 + cx.en; i += NT) y[i] += al * x[i]; };
    auto waxpy_v = [&](double* w, const double* y, double al, const double* x) { for (int i = cx.v0 + (int)threadIdx.x; i < cx.v0 + cx.vn; i += NT) w[i] = y[i] + al * x[i]; };
    auto waxpy_e = [&](double* w, const double* y, double al, const double* x) { for (int i = cx.e0 + (int)threadIdx.x; i < cx.e0 + cx.en; i += NT) w[i] = y[i] + al * x[i]; };
    // initial guess of a solve: extrapolation through the last solutions of its site, written into the current iterate buffer
    auto predict = [&](int field, int site, int which) {
        const int order = a.extrap[field] + 1;
        const int mth = hc[field][site] < order ? hc[field][site] : order;
        if (mth <= 0 || !a.hist[field][site][0]) return;
        double c[6]; const double* h[6];
        double binom = 1.0;
        for (int j = 0; j < mth; ++j) {
            binom = binom * (double)(mth - j) / (double)(j + 1);
            c[j] = (j % 2 == 0) ? binom : -binom;
            h[j] = a.hist[field][site][((hh[field][site] - j) % order + order) % order];
        }
        double* x = field == 0 ? tbp(which) : tbq(which);
        const int i0 = field == 0 ? cx.v0 : cx.e0, n = field == 0 ? cx.vn : cx.en;
        for (int i = i0 + (int)threadIdx.x; i < i0 + n; i += NT) {
            double v = c[0] * h[0][i];
            for (int j = 1; j < mth; ++j) v += c[j] * h[j][i];
            x[i] = v;
        }
    };
    auto remember = [&](int field, int site, int which) {
        if (!a.hist[field][site][0]) return;
        const int order = a.extrap[field] + 1;
        hh[field][site] = (hh[field][site] + 1) % order;
        hc[field][site] = hc[field][site] + 1 < order ? hc[field][site] + 1 : order;
        double* dst = a.hist[field][site][hh[field][site]];
        const double* x = field == 0 ? tbp(which) : tbq(which);
        const int i0 = field == 0 ? cx.v0 : cx.e0, n = field == 0 ? cx.vn : cx.en;
        for (int i = i0 + (int)threadIdx.x; i < i0 + n; i += NT) dst[i] = x[i];
    };
    // strong Dirichlet rows: the velocity unknown of a constrained vertex is prescribed (dirichlet_set_k)
    auto dirichlet_preset = [&](int which) {
        if (!a.strong || a.n_dir == 0) return;
        if (leader) {
            double* x = tbp(which);
            const double pl = *((volatile double*)&ctl.pleft);
            for (int i = threadIdx.x; i < a.n_dir; i += NT) x[a.dir_idx[i]] = (a.dir_w[i] * pl - a.p[a.dir_idx[i]]) * (1.0 / dt);
        }
    };
    // b_p = -K (D^T qsrc [+ B_top psrc])
    auto rhs_p = [&](const double* qsrc, const double* psrc) {
        RowArgs r = loop_base_args(a);
        r.xp = psrc; r.xq = qsrc; r.cM = 0.0; r.cD = -Kg; r.cB = a.casimir ? -Kg : 0.0; r.y = a.bp;
        PassIO io{}; io.xp = psrc; io.xq = qsrc;
        double d0 = 0.0, d1 = 0.0;
        if (a.casimir) vertex_pass<EPI_STORE, MODE_OP, false, true, true, NT>(r, io, smem, d0, d1);
        else vertex_pass<EPI_STORE, MODE_OP, false, true, false, NT>(r, io, smem, d0, d1);
    };
    // b_q = (1/rho) (D psrc - b_L p_left [- B_right qsrc]); the port term is added by the leader after a barrier
    auto rhs_q = [&](const double* psrc, const double* qsrc) {
        RowArgs r = loop_base_args(a);
        r.xp = psrc; r.xq = qsrc; r.cM = 0.0; r.cD = irg; r.cB = a.casimir ? -irg : 0.0; r.y = a.bq;
        PassIO io{}; io.xp = psrc; io.xq = qsrc;
        double d0 = 0.0, d1 = 0.0;
        if (a.casimir) edge_pass<EPI_STORE, MODE_OP, false, true, true, NT>(r, io, smem, d0, d1);
        else edge_pass<EPI_STORE, MODE_OP, false, true, false, NT>(r, io, smem, d0, d1);
        if (!a.strong && a.nport > 0) {
            SYNC();
            if (leader) {
                const double pl = *((volatile double*)&ctl.pleft);
                for (int i = a.port_goff[cx.g] + (int)threadIdx.x; i < a.port_goff[cx.g + 1]; i += NT) a.bq[a.port_e[i]] -= irg * a.port_bl[i] * pl;
            }
        }
    };
    // Chebyshev solve of M_p x = bp (which 0) / M_q x = bq (which 1); warm start = iterate buffer `cur`; ends behind a barrier
    auto solve_mass = [&](int which) {
        RowArgs r = loop_base_args(a);
        r.which = which; r.cM = 1.0;
        r.b = which == 0 ? a.bp : a.bq; r.pinv = which == 0 ? a.pinv_p : a.pinv_q;
        const bool keep_mdiag = a.strong || (a.casimir && implicit);
        if (!keep_mdiag) r.mdiag = nullptr;
        int k = 0; double rho = 0.0, rr = 0.0, bb = 0.0; bool conv = false;
        int c_ = cur[which];
        SYNC();                                    // right-hand side and initial guess of the neighbouring patches are complete
        while (true) {
            double c1, c2, rho_k;
            cheb_coeffs(k, rho, a.cheb_d[which], a.cheb_c[which], c1, c2, rho_k);
            const int nxt = (c_ + 1) % 3, prv = (c_ + 2) % 3;
            PassIO io{};
            io.c1 = c1; io.c2 = c2; io.use_prev = k > 0;
            double r0 = 0.0, b0 = 0.0;
            if (which == 0) {
                io.xp = tb_sel(a.tp, c_); io.xprev = tb_sel(a.tp, prv); io.xnew = tb_sel(a.tp, nxt);
                if (a.ell_p && !keep_mdiag) vertex_ell_pass<NT, true>(r, io, smem, r0, b0);
                else vertex_pass<EPI_CHEB, MODE_OP, true, false, false, NT>(r, io, smem, r0, b0);
            } else {
                io.xq = tb_sel(a.tq, c_); io.xprev = tb_sel(a.tq, prv); io.xnew = tb_sel(a.tq, nxt);
                edge_pass<EPI_CHEB, MODE_OP, true, false, false, NT, true>(r, io, smem, r0, b0);
            }
            loop_reduce2<CLUSTER>(a, cx, r0, b0);
            rr = r0; bb = b0;
            conv = rr <= a.tol2 * bb;
            k += 1;
            if (conv) break;                       // the verified iterate x_k (buffer c_) is the solution
            rho = rho_k; c_ = nxt;
            if (k >= a.max_iter) break;
        }
        cur[which] = c_;
        (which == 0 ? it_p : it_q) += k; (which == 0 ? n_p : n_q) += 1;
        if (!conv) nonconv += 1;
        const double rel = bb > 0.0 ? sqrt(rr / bb) : 0.0;
        if (rel > maxres[which]) maxres[which] = rel;
    };
    // RES: the same Chebyshev solves with every operand of an iteration in shared memory (see the header of this file).  Same rows,
    // same arithmetic and the same accepted iterate as solve_mass; begins and ends behind a cluster barrier.
    auto solve_mass_res = [&](int which) {
#ifndef PHW_HOST_EMU
        const PatchHdr& h = myh;
        const int n_own = which == 0 ? h.v_cnt : h.e_cnt, n_gh = which == 0 ? h.gv_cnt : h.ge_cnt;
        const int g0_ = which == 0 ? h.v_start : h.e_start;
        double* xb[2] = {reinterpret_cast<double*>(sbase + (which == 0 ? a.rc.xp[0] : a.rc.xq[0])),
                         reinterpret_cast<double*>(sbase + (which == 0 ? a.rc.xp[1] : a.rc.xq[1]))};
        double* sb = reinterpret_cast<double*>(sbase + (which == 0 ? a.rc.bp : a.rc.bq));
        double* part = reinterpret_cast<double*>(sbase + a.rc.part);
        const uint32_t* gsrc = reinterpret_cast<const uint32_t*>(sbase + (which == 0 ? a.rc.gsv : a.rc.gse));
        const double* gb = which == 0 ? a.bp : a.bq;
        double* gx = which == 0 ? tbp(which) : tbq(which);
        SYNC();                                    // right-hand side (port term included) and initial guess are complete in global memory
        const double* dg = reinterpret_cast<const double*>(sbase + (which == 0 ? a.rc.dgp : a.rc.dgq));
        for (int i = threadIdx.x; i < n_own; i += NT) { sb[i] = gb[g0_ + i] / dg[i]; xb[0][i] = gx[g0_ + i]; }   // Jacobi-scaled right-hand side
        auto pull_ghosts = [&](double* xdst) {     // ghost copies of the iterate out of the owners' shared memory
            for (int i = threadIdx.x; i < n_gh; i += NT) {
                const uint32_t code = gsrc[i];
                xdst[n_own + i] = ld_dsmem_f64(dsmem_addr(xdst + (code & 0xFFFFu), code >> 16));
            }
        };
        cluster_barrier();
        pull_ghosts(xb[0]);
        __syncthreads();
        int k = 0, c_ = 0; double rho = 0.0, rr = 0.0, bb = 0.0; bool conv = false;
        // The residual norm only decides when to stop (the Chebyshev recurrence itself needs no inner product), and here - unlike
        // in the bandwidth-bound kernels - its block reduction and exchange are a third of an iteration.  It is therefore measured
        // only when the solve can be close to the tolerance: after a measurement the asymptotic rate sigma of the recurrence gives
        // the number of iterations still needed, and the next measurement waits for 60 % of them.  An iterate is never accepted
        // without its own measured residual.
        const double kap_ = (a.cheb_d[which] + a.cheb_c[which]) / (a.cheb_d[which] - a.cheb_c[which]);
        const double lsig_ = -log((sqrt(kap_) - 1.0) / (sqrt(kap_) + 1.0));      // > 0
        int next_check = 0;
        while (true) {
            double c1, c2, rho_k;
            cheb_coeffs(k, rho, a.cheb_d[which], a.cheb_c[which], c1, c2, rho_k);
            const bool do_norm = k >= next_check;
            const double* __restrict__ xc_ = xb[c_];
            double* __restrict__ xo_ = xb[c_ ^ 1];                    // holds x_{k-1}, receives x_{k+1}
            double r0 = 0.0, b0 = 0.0;
            if (which == 0) {
                const uint32_t* s_io = reinterpret_cast<const uint32_t*>(sbase + a.rc.sio);
                const uint32_t* s_wo = reinterpret_cast<const uint32_t*>(sbase + a.rc.swo);
                const uint4* sid4 = reinterpret_cast<const uint4*>(sbase + a.rc.ids);
                const double* sw = reinterpret_cast<const double*>(sbase + a.rc.w);
                for (int lv = threadIdx.x; lv < h.v_cnt; lv += NT) {
                    const int sl = lv >> 5, lane = lv & 31;
                    const uint32_t io0 = s_io[sl], wo0 = s_wo[sl];
                    const int W = (int)((s_wo[sl + 1] - wo0) >> 5);
                    double val = 0.0;
                    for (int c = 0; 8 * c < W; ++c) {
                        const uint4 cd = sid4[(io0 >> 3) + c * 32 + lane];
                        const uint32_t w4[4] = {cd.x, cd.y, cd.z, cd.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const double wa = (8 * c + 2 * j < W) ? sw[wo0 + (8 * c + 2 * j) * 32 + lane] : 0.0;
                            const double wb = (8 * c + 2 * j + 1 < W) ? sw[wo0 + (8 * c + 2 * j + 1) * 32 + lane] : 0.0;
                            val += wa * xc_[w4[j] & 0xFFFFu];
                            val += wb * xc_[w4[j] >> 16];
                        }
                    }
                    const double xc = xc_[lv], bv = sb[lv];
                    const double r_ = bv - xc - val;                  // Jacobi-scaled residual
                    double xn = xc + c2 * r_;
                    if (k > 0) xn += c1 * (xc - xo_[lv]);
                    xo_[lv] = xn;
                    if (do_norm) { const double d_ = dg[lv]; r0 += d_ * r_ * r_; b0 += d_ * bv * bv; }
                }
            } else {
                const double* wq = reinterpret_cast<const double*>(sbase + a.rc.wq);
                const uint16_t* idq = reinterpret_cast<const uint16_t*>(sbase + a.rc.idq);
                const int me = a.rc.me;
                for (int le = threadIdx.x; le < h.e_cnt; le += NT) {
                    const double xc = xc_[le];
                    double val = wq[le] * xc_[idq[le]];
                    val += wq[me + le] * xc_[idq[me + le]];
                    val += wq[2 * me + le] * xc_[idq[2 * me + le]];
                    val += wq[3 * me + le] * xc_[idq[3 * me + le]];
                    const double bv = sb[le];
                    const double r_ = bv - xc - val;
                    double xn = xc + c2 * r_;
                    if (k > 0) xn += c1 * (xc - xo_[le]);
                    xo_[le] = xn;
                    if (do_norm) { const double d_ = dg[le]; r0 += d_ * r_ * r_; b0 += d_ * bv * bv; }
                }
            }
            const int par = k & 1;
            if (do_norm) {
                block_sum2(r0, b0, cx.red);
                if (threadIdx.x == 0) { part[2 * par] = r0; part[2 * par + 1] = b0; }
            }
            cluster_barrier();                     // x_{k+1} (owned rows) and the norms of every CTA of the case are in place
            pull_ghosts(xo_);
            if (do_norm && threadIdx.x < 32) {
                double s0 = 0.0, s1 = 0.0;
                if ((int)threadIdx.x < cx.ncta) {
                    const unsigned ra = dsmem_addr(part + 2 * par, threadIdx.x);
                    s0 = ld_dsmem_f64(ra); s1 = ld_dsmem_f64(ra + 8);
                }
                s0 = warp_sum(s0); s1 = warp_sum(s1);
                if (threadIdx.x == 0) { cx.bc[0] = s0; cx.bc[1] = s1; }
            }
            __syncthreads();
            k += 1;
            if (do_norm) {
                rr = cx.bc[0]; bb = cx.bc[1];
                conv = rr <= a.tol2 * bb;
                if (conv) break;                   // the verified iterate x_k (buffer c_) is the solution
                const double need = 0.5 * log(rr / (a.tol2 * bb)) / lsig_;      // iterations the asymptotic rate still asks for
                const int skip = need < 1e6 ? (int)(0.6 * need) : 0;
                next_check = k + (skip > 1 ? skip - 1 : 0);
                if (next_check > a.max_iter - 1) next_check = a.max_iter - 1;      // the last allowed iterate is always measured
            }
            rho = rho_k; c_ ^= 1;
            if (k >= a.max_iter) break;
        }
        for (int i = threadIdx.x; i < n_own; i += NT) gx[g0_ + i] = xb[c_][i];
        (which == 0 ? it_p : it_q) += k; (which == 0 ? n_p : n_q) += 1;
        if (!conv) nonconv += 1;
        const double rel = bb > 0.0 ? sqrt(rr / bb) : 0.0;
        if (rel > maxres[which]) maxres[which] = rel;
        SYNC();                                    // solution visible to every CTA of the case; all ghost pulls have finished
#endif
    };
    // Chebyshev (ellipse) solve of the coupled block system of IE / SM (solve_block of phwave.cu)
    auto solve_block = [&](double theta) {
        RowArgs av = loop_base_args(a);
        av.which = 2; av.cM = 1.0;
        RowArgs ae = av;
        av.b = a.bp; av.pinv = a.pinv_p; av.cD = theta * dt * a.K; av.cB = a.casimir ? theta * dt * a.K : 0.0; av.weight = 1.0 / a.K;
        ae.b = a.bq; ae.pinv = a.pinv_q; ae.cD = -theta * dt / a.rho; ae.cB = a.casimir ? theta * dt / a.rho : 0.0; ae.weight = a.rho;
        int k = 0; double rho = 0.0, rr = 0.0, bb = 0.0; bool conv = false;
        int c_ = cur[2];
        SYNC();
        while (true) {
            double c1, c2, rho_k;
            cheb_coeffs(k, rho, a.cheb_d[2], a.cheb_c[2], c1, c2, rho_k);
            const int nxt = (c_ + 1) % 3, prv = (c_ + 2) % 3;
            PassIO iov{}, ioe{};
            iov.xp = tb_sel(a.tp, c_); iov.xq = tb_sel(a.tq, c_); iov.xprev = tb_sel(a.tp, prv); iov.xnew = tb_sel(a.tp, nxt);
            ioe.xp = iov.xp; ioe.xq = iov.xq; ioe.xprev = tb_sel(a.tq, prv); ioe.xnew = tb_sel(a.tq, nxt);
            iov.c1 = ioe.c1 = c1; iov.c2 = ioe.c2 = c2; iov.use_prev = ioe.use_prev = (k > 0);
            double r0 = 0.0, b0 = 0.0, a0_ = 0.0, a1_ = 0.0;
            if (a.casimir) vertex_pass<EPI_CHEB, MODE_OP, true, true, true, NT>(av, iov, smem, a0_, a1_);
            else vertex_pass<EPI_CHEB, MODE_OP, true, true, false, NT>(av, iov, smem, a0_, a1_);
            r0 += av.weight * a0_; b0 += av.weight * a1_;
            a0_ = 0.0; a1_ = 0.0;
            if (a.casimir) edge_pass<EPI_CHEB, MODE_OP, true, true, true, NT, true>(ae, ioe, smem, a0_, a1_);
            else edge_pass<EPI_CHEB, MODE_OP, true, true, false, NT, true>(ae, ioe, smem, a0_, a1_);
            r0 += ae.weight * a0_; b0 += ae.weight * a1_;
            loop_reduce2<CLUSTER>(a, cx, r0, b0);
            rr = r0; bb = b0;
            conv = rr <= a.tol2 * bb;
            k += 1;
            if (conv) break;
            rho = rho_k; c_ = nxt;
            if (k >= a.max_iter) break;
        }
        cur[2] = c_;
        it_b += k; n_b += 1;
        if (!conv) nonconv += 1;
        const double rel = bb > 0.0 ? sqrt(rr / bb) : 0.0;
        if (rel > maxres[2]) maxres[2] = rel;
    };
    auto solve_site = [&](int which, int site) {
        predict(which, site, which);
        if (which == 0 && a.strong) { SYNC(); dirichlet_preset(0); }
        // strong Dirichlet rows and the casimir mass diagonal need the generic vertex rows
        if (RES && (which == 1 || (a.ell_p && !(a.strong || (a.casimir && implicit))))) solve_mass_res(which);
        else solve_mass(which);
        remember(which, site, which);
    };
    // H_wave = 0.5/rho p^T M_p p + 0.5 K q^T M_q q  (wave_2D.py:882,907) -> red[0], red[1]; ends behind a barrier
    auto energy = [&]() {
        RowArgs r = loop_base_args(a);
        r.cM = 1.0; r.xp = a.p; r.xq = a.q;
        PassIO io{}; io.xp = a.p; io.xq = a.q;
        double ap = 0.0, aq = 0.0, z0 = 0.0, z1 = 0.0;
        vertex_pass<EPI_DOT, MODE_OP, true, false, false, NT>(r, io, smem, ap, z0);
        edge_pass<EPI_DOT, MODE_OP, true, false, false, NT>(r, io, smem, aq, z1);
        loop_reduce2<CLUSTER>(a, cx, ap, aq);
        if (leader && threadIdx.x == 0) { ctl.red[0] = 0.5 * irg * ap; ctl.red[1] = 0.5 * Kg * aq; }
    };
    // casimir: boundary quadratic forms of the midpoint averages (wave_2D.py:633-634) -> red[2], red[3]
    auto casimir_dots = [&](const double* xp, const double* xq) {
        RowArgs r = loop_base_args(a);
        r.cM = 0.0; r.cB = 1.0; r.xp = xp; r.xq = xq;
        PassIO io{}; io.xp = xp; io.xq = xq;
        double ap = 0.0, aq = 0.0, z0 = 0.0, z1 = 0.0;
        vertex_pass<EPI_DOT, MODE_OP, false, false, true, NT>(r, io, smem, ap, z0);
        edge_pass<EPI_DOT, MODE_OP, false, false, true, NT>(r, io, smem, aq, z1);
        loop_reduce2<CLUSTER>(a, cx, ap, aq);
        if (leader && threadIdx.x == 0) { ctl.red[2] = ap; ctl.red[3] = aq; }
    };

    // ---- the time loop (same order of operations as step() of phwave.cu) -----------------------------------------------------
    for (long long n = 0; n < a.n_steps; ++n) {
        switch (a.scheme) {
            case PHW_EE:
                scalar(S_BEGIN_1SUB, 0.0);
                flux(a.q, 0);
                rhs_p(a.q, a.p); solve_site(0, 0);
                rhs_q(a.p, a.q); solve_site(1, 0);
                axpy_v(a.p, dt, tbp(0));
                axpy_e(a.q, dt, tbq(1));
                SYNC();
                energy();
                scalar(S_END_1SUB, 0.0);
                break;
            case PHW_SE:
                scalar(S_BEGIN_1SUB, 0.0);
                flux(a.q, 0);
                rhs_p(a.q, a.p); solve_site(0, 0);
                axpy_v(a.p, dt, tbp(0));
                if (a.ic) scalar(S_ODE_SE, 0.0);
                SYNC();
                rhs_q(a.p, a.q); solve_site(1, 0);
                axpy_e(a.q, dt, tbq(1));
                SYNC();
                energy();
                scalar(S_END_1SUB, 0.0);
                break;
            case PHW_SV:
                scalar(S_SV_STAGE1, 0.0);
                if (!have_vq || a.sv_noreuse) { SYNC(); rhs_q(a.p, a.q); solve_site(1, 0); }
                axpy_e(a.q, 0.5 * dt, tbq(1));                        // q_temp
                SYNC();
                flux(a.q, 0);
                scalar(S_SV_MID, 0.0);
                rhs_p(a.q, a.p); solve_site(0, 0);
                axpy_v(a.p, dt, tbp(0));
                if (a.ic) scalar(S_ODE_SV, 0.0);
                SYNC();
                rhs_q(a.p, a.q); solve_site(1, 0);
                axpy_e(a.q, 0.5 * dt, tbq(1));
                have_vq = true;
                SYNC();
                energy();
                scalar(S_END_SV, 0.0);
                break;
            case PHW_EH:
                scalar(S_EH_BEGIN, 0.0);
                flux(a.q, 0);
                rhs_p(a.q, a.p); solve_site(0, 0);
                rhs_q(a.p, a.q); solve_site(1, 0);
                flux(tbq(1), 1);                                       // b_L^T v_q1
                waxpy_v(a.tmp_p, a.p, 0.5 * dt, tbp(0));               // 0.5 (p_m + p_temp)
                waxpy_e(a.tmp_q, a.q, 0.5 * dt, tbq(1));
                SYNC();
                rhs_p(a.tmp_q, a.tmp_p); solve_site(0, 1);
                rhs_q(a.tmp_p, a.tmp_q); solve_site(1, 1);
                axpy_v(a.p, dt, tbp(0));
                axpy_e(a.q, dt, tbq(1));
                SYNC();
                energy();
                scalar(S_END_EH, 0.0);
                break;
            case PHW_IE:
            case PHW_SM: {
                const double theta = (a.scheme == PHW_IE) ? 1.0 : 0.5;
                scalar(S_BEGIN_1SUB, 0.0);
                flux(a.q, 0);
                if (a.ic) scalar(S_SM_PLEFT, 0.0);
                rhs_p(a.q, a.p);
                rhs_q(a.p, a.q);
                predict(0, 0, 2); predict(1, 0, 2);
                if (a.strong) { SYNC(); dirichlet_preset(2); }
                solve_block(theta);
                remember(0, 0, 2); remember(1, 0, 2);
                if (a.ic) {
                    flux(tbq(2), 2);                                   // b_L^T vhat_q
                    scalar(S_ODE_SM, a.blw);
                    SYNC();
                    const double xi1 = *((volatile double*)&ctl.xi1);
                    axpy_v(tbp(2), xi1, a.wp);
                    axpy_e(tbq(2), xi1, a.wq);
                }
                if (a.casimir) {
                    waxpy_v(a.tmp_p, a.p, 0.5 * dt, tbp(2));
                    waxpy_e(a.tmp_q, a.q, 0.5 * dt, tbq(2));
                    SYNC();
                    casimir_dots(a.tmp_p, a.tmp_q);
                }
                axpy_v(a.p, dt, tbp(2));
                axpy_e(a.q, dt, tbq(2));
                SYNC();
                flux(a.q, 1);
                energy();
                scalar(S_END_1SUB, 0.0);
            } break;
            default: break;
        }
    }
    SYNC();
    if (leader && threadIdx.x == 0) {
        for (int w = 0; w < 3; ++w) { ctl.sc[w].cur = cur[w]; if (maxres[w] > ctl.sc[w].maxres) ctl.sc[w].maxres = maxres[w]; }
        ctl.sc[0].iters += it_p; ctl.sc[0].solves += n_p;
        ctl.sc[1].iters += it_q; ctl.sc[1].solves += n_q;
        ctl.sc[2].iters += it_b; ctl.sc[2].solves += n_b;
        ctl.nonconv += nonconv;
    }
}

}  // namespace phw
