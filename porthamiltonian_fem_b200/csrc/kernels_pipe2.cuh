// kernels_pipe2.cuh - two Chebyshev iterations per pass AND bulk-copy pipelining for the mass solves (experimental; flags bit 13:
// M_q, bit 14: M_p).
//
// kernels_pipe.cuh brought the one-sweep solves to the DRAM roofline of their traffic (99 B per cell and iteration for M_q), so
// the next step is fewer bytes: the two-sweep patches of kernels2.cuh (one more ring of halo cells; x_{k+1} is formed on
// D0+G1 and x_{k+2} on D0 without leaving shared memory) read every iteration-invariant stream - cell records, metric,
// adjacency, Jacobi weights, right-hand side - once per TWO iterations.  Their lean version did not pay because its phases,
// not its bytes, were the limit; with every stream of the next patch in flight during the four compute
// phases of the current one, the bytes are what is left: ~63 instead of 99 B per cell and iteration at 768-cell patches.
//
// Same recurrence, same stop rule (true residuals of both iterates, the first one that meets the tolerance is accepted) and
// the same four rotating iterate buffers as solve2_coop_k<1>; same stage protocol as solve_mq_pipe_k.
// NOT YET RUN ON A GPU: opt-in only, tests gated by PHW_TEST_EXPERIMENTAL.
#pragma once
#include "kernels2.cuh"
#include "kernels_pipe.cuh"

namespace phw {

struct Mq2PipeArgs {
    Ring2Dev m;
    double* buf[4];
    const double* b;
    double cheb_d, cheb_c, tol2;
    int max_iter, which;
    double* partials;            // [2][3 * gridDim.x]
    Ctl* ctl;
    // stage = [x_k on D0|G1|G2 | x_{k-1} rows | b rows | pinv rows | eadj rows | rec cells | g0 | g1 | g2]; then the cell
    // contributions, the saved x_k of D0 and two ghost-index lists
    uint32_t off_xprev, off_b, off_pinv, off_eadj, off_rec, off_g0, off_g1, off_g2, stage_bytes, off_sc3, off_xk0, off_gidx, gidx_bytes;
};

__device__ __forceinline__ void load_hdr2_smem_at(Ring2Hdr* dst, const Ring2Hdr* src, int first_thread) {
    const int t = (int)threadIdx.x - first_thread;
    if (t >= 0 && t < 20) reinterpret_cast<int*>(dst)[t] = __ldg(reinterpret_cast<const int*>(src) + t);
}

template <int NT>
__device__ __forceinline__ void mq2_issue(const Mq2PipeArgs& a, unsigned char* stage, const uint32_t* gidx_this, uint32_t* gidx_next,
                                          const Ring2Hdr& hn, const Ring2Hdr* hn2, const double* xk, const double* xprev,
                                          bool use_prev, unsigned long long* bar) {
    const Ring2Dev& m = a.m;
    double* sx = reinterpret_cast<double*>(stage);
    double* sxp = reinterpret_cast<double*>(stage + a.off_xprev);
    double* sb = reinterpret_cast<double*>(stage + a.off_b);
    double* spv = reinterpret_cast<double*>(stage + a.off_pinv);
    uint32_t* sea = reinterpret_cast<uint32_t*>(stage + a.off_eadj);
    double* srec = reinterpret_cast<double*>(stage + a.off_rec);
    double* sg0 = reinterpret_cast<double*>(stage + a.off_g0);
    double* sg1 = reinterpret_cast<double*>(stage + a.off_g1);
    double* sg2 = reinterpret_cast<double*>(stage + a.off_g2);
    const double* grec = reinterpret_cast<const double*>(m.rec);
    const uint32_t* gh = reinterpret_cast<const uint32_t*>(m.ghost);
    const int ds = hn.d_start, n0 = hn.d_cnt, nr = n0 + hn.g1_cnt, ng = hn.g1_cnt + hn.g2_cnt, co = hn.cell_off, nc = hn.n_c2, ro = hn.row_off;
    if (threadIdx.x == 0) {
        uint32_t bytes = f64s_bytes(ds, n0) * (use_prev ? 3u : 2u) + f64s_bytes(ro, nr) + u32s_bytes(ro, nr) + 4u * f64s_bytes(co, nc);
        if (hn2) bytes += u32s_bytes(hn2->g_off, hn2->g1_cnt + hn2->g2_cnt);
        mbar_arrive_expect_tx(bar, bytes);
        f64s_bulk(sg0, m.g0, co, nc, bar);
        f64s_bulk(sg1, m.g1, co, nc, bar);
        f64s_bulk(sg2, m.g2, co, nc, bar);
        f64s_bulk(srec, grec, co, nc, bar);
        f64s_bulk(spv, m.pinv, ro, nr, bar);
        u32s_bulk(sea, m.eadj, ro, nr, bar);
        f64s_bulk(sx, xk, ds, n0, bar);
        f64s_bulk(sb, a.b, ds, n0, bar);
        if (use_prev) f64s_bulk(sxp, xprev, ds, n0, bar);
        if (hn2) u32s_bulk(gidx_next, gh, hn2->g_off, hn2->g1_cnt + hn2->g2_cnt, bar);
    }
    const int lane = (int)threadIdx.x - (NT - 32);      // the last warp copies the unaligned ends of the streams
    if (lane >= 0) {
        if (lane < 2) f64s_fix(sx, xk, ds, n0, lane);
        else if (lane < 4) f64s_fix(sb, a.b, ds, n0, lane - 2);
        else if (lane < 6) { if (use_prev) f64s_fix(sxp, xprev, ds, n0, lane - 4); }
        else if (lane < 8) f64s_fix(spv, m.pinv, ro, nr, lane - 6);
        else if (lane < 10) f64s_fix(srec, grec, co, nc, lane - 8);
        else if (lane < 12) f64s_fix(sg0, m.g0, co, nc, lane - 10);
        else if (lane < 14) f64s_fix(sg1, m.g1, co, nc, lane - 12);
        else if (lane < 16) f64s_fix(sg2, m.g2, co, nc, lane - 14);
        else if (lane < 22) u32s_fix(sea, m.eadj, ro, nr, lane - 16);
        else if (lane < 28) { if (hn2) u32s_fix(gidx_next, gh, hn2->g_off, hn2->g1_cnt + hn2->g2_cnt, lane - 22); }
    }
    // gathered parts: x_k on G1|G2, b and x_{k-1} on G1 (the rows of the first sweep that belong to other patches)
    double* sxg = sx + (ds & 1) + n0;
    double* sbg = sb + (ds & 1) + n0;
    double* sxpg = sxp + (ds & 1) + n0;
    const uint32_t* gi = gidx_this ? gidx_this + (hn.g_off & 3) : nullptr;
    for (int i = threadIdx.x; i < ng; i += NT) {
        const uint32_t g = gi ? gi[i] : (uint32_t)__ldg(&m.ghost[hn.g_off + i]);
        cp_async8(&sxg[i], &xk[g]);
        if (i < hn.g1_cnt) {
            cp_async8(&sbg[i], &a.b[g]);
            if (use_prev) cp_async8(&sxpg[i], &xprev[g]);
        }
    }
}

// RT1 element mass row values of one cell from staged records (same arithmetic as rt1_cell in kernels2.cuh)
__device__ __forceinline__ void rt1_cell_s(const uint2 re, const double g0, const double g1, const double g2, const double* sx, double* out3) {
    const uint32_t e0 = re.x & 0xFFFFu, e1 = re.x >> 16, e2 = re.y & 0xFFFFu;
    const double S = g0 + g1 + g2;
    const double s0 = (e0 & 0x8000u) ? -1.0 : 1.0, s1 = (e1 & 0x8000u) ? -1.0 : 1.0, s2 = (e2 & 0x8000u) ? -1.0 : 1.0;
    const double t0 = s0 * sx[e0 & 0x7FFFu], t1 = s1 * sx[e1 & 0x7FFFu], t2 = s2 * sx[e2 & 0x7FFFu];
    const double ST = S * (t0 + t1 + t2);
    out3[0] = s0 * (ST + 2.0 * S * t0 - 4.0 * (g0 * t0 + g2 * t1 + g1 * t2));
    out3[1] = s1 * (ST + 2.0 * S * t1 - 4.0 * (g2 * t0 + g1 * t1 + g0 * t2));
    out3[2] = s2 * (ST + 2.0 * S * t2 - 4.0 * (g1 * t0 + g0 * t1 + g2 * t2));
}

template <int NT>
__global__ void __launch_bounds__(NT, 1) solve2_mq_pipe_k(const __grid_constant__ Mq2PipeArgs a) {
    PHW_DYN_SMEM(smraw);
    __shared__ double red[96];
    __shared__ double bc[3];
    __shared__ Ring2Hdr s_hdr[4];
    __shared__ __align__(8) unsigned long long full[2];
    __shared__ int s_dead;
    SolveCtl& sc = a.ctl->sc[a.which];
    int k = 0, cur = sc.cur & 3, pass = 0;
    double rho = 0.0, rr = 0.0, bbv = 0.0;
    bool conv = false;
    unsigned int epoch = 0;
    unsigned int seq = 0;
    if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); s_dead = 0; }
    __syncthreads();
    double* sc3 = reinterpret_cast<double*>(smraw + a.off_sc3);
    double* sxk0 = reinterpret_cast<double*>(smraw + a.off_xk0);
    const int npatch = a.m.npatch;
    const int n_my = ((int)blockIdx.x < npatch) ? (npatch - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    while (true) {
        double c1a, c2a, c1b, c2b, rho_a, rho_b;
        cheb_coeffs(k, rho, a.cheb_d, a.cheb_c, c1a, c2a, rho_a);
        cheb_coeffs(k + 1, rho_a, a.cheb_d, a.cheb_c, c1b, c2b, rho_b);
        const double* xk = sel4(a.buf, cur);
        double* x1 = sel4(a.buf, (cur + 1) & 3);
        double* x2 = sel4(a.buf, (cur + 2) & 3);
        const double* xprev = sel4(a.buf, (cur + 3) & 3);
        const bool use_prev = k > 0;
        double rr0 = 0.0, rr1 = 0.0, bb = 0.0;
        for (int j = 0; j < 3 && j < n_my; ++j) load_hdr2_smem_at(&s_hdr[j], &a.m.hdr[blockIdx.x + j * gridDim.x], 32 * j);
        __syncthreads();
        if (n_my > 0)
            mq2_issue<NT>(a, smraw + (seq & 1u) * a.stage_bytes, nullptr,
                          reinterpret_cast<uint32_t*>(smraw + a.off_gidx + ((seq + 1u) & 1u) * a.gidx_bytes),
                          s_hdr[0], n_my > 1 ? &s_hdr[1] : nullptr, xk, xprev, use_prev, &full[seq & 1u]);
        for (int it = 0; it < n_my; ++it, ++seq) {
            const unsigned int s = seq & 1u;
            if (!pipe_stage_ready(&full[s], (seq >> 1) & 1u, a.ctl, &s_dead)) continue;
            const Ring2Hdr& h = s_hdr[it & 3];
            if (it + 3 < n_my) load_hdr2_smem_at(&s_hdr[(it + 3) & 3], &a.m.hdr[blockIdx.x + (it + 3) * gridDim.x], 64);
            if (it + 1 < n_my)
                mq2_issue<NT>(a, smraw + (s ^ 1u) * a.stage_bytes,
                              reinterpret_cast<const uint32_t*>(smraw + a.off_gidx + (s ^ 1u) * a.gidx_bytes),
                              reinterpret_cast<uint32_t*>(smraw + a.off_gidx + s * a.gidx_bytes),
                              s_hdr[(it + 1) & 3], it + 2 < n_my ? &s_hdr[(it + 2) & 3] : nullptr, xk, xprev, use_prev, &full[s ^ 1u]);
            unsigned char* stage = smraw + s * a.stage_bytes;
            const int n0 = h.d_cnt, nr = n0 + h.g1_cnt;
            double* sx = reinterpret_cast<double*>(stage) + (h.d_start & 1);          // x_k, rows [0, nr) overwritten with x_{k+1}
            const double* sxp = reinterpret_cast<const double*>(stage + a.off_xprev) + (h.d_start & 1);
            const double* sb = reinterpret_cast<const double*>(stage + a.off_b) + (h.d_start & 1);
            const double* spv = reinterpret_cast<const double*>(stage + a.off_pinv) + (h.row_off & 1);
            const uint32_t* sea = reinterpret_cast<const uint32_t*>(stage + a.off_eadj) + (h.row_off & 3);
            const uint2* srec = reinterpret_cast<const uint2*>(stage + a.off_rec) + (h.cell_off & 1);
            const double* sg0 = reinterpret_cast<const double*>(stage + a.off_g0) + (h.cell_off & 1);
            const double* sg1 = reinterpret_cast<const double*>(stage + a.off_g1) + (h.cell_off & 1);
            const double* sg2 = reinterpret_cast<const double*>(stage + a.off_g2) + (h.cell_off & 1);
            // ---- sweep 1, cells C2
            for (int lc = threadIdx.x; lc < h.n_c2; lc += NT) rt1_cell_s(srec[lc], sg0[lc], sg1[lc], sg2[lc], sx, &sc3[3 * lc]);
            __syncthreads();
            // ---- sweep 1, rows [D0 | G1]: x_{k+1}, kept in shared memory (in place: a row's x_k is read by its own thread only)
            for (int r = threadIdx.x; r < nr; r += NT) {
                const uint32_t adj = sea[r];
                const uint32_t ca = adj & 0xFFFFu, cb = adj >> 16;
                const double bv = sb[r], pv = spv[r];
                double val = sc3[3 * (ca >> 2) + (ca & 3u)];
                if (cb != 0xFFFFu) val += sc3[3 * (cb >> 2) + (cb & 3u)];
                const double xc = sx[r];
                const double res = bv - val;
                double xn = xc + c2a * pv * res;
                if (use_prev) xn += c1a * (xc - sxp[r]);
                sx[r] = xn;
                if (r < n0) {
                    sxk0[r] = xc;
                    x1[h.d_start + r] = xn;
                    rr0 += pv * res * res;
                    bb += pv * bv * bv;
                }
            }
            __syncthreads();
            // ---- sweep 2, cells C1 (all their edges are rows of sweep 1)
            for (int lc = threadIdx.x; lc < h.n_c1; lc += NT) rt1_cell_s(srec[lc], sg0[lc], sg1[lc], sg2[lc], sx, &sc3[3 * lc]);
            __syncthreads();
            // ---- sweep 2, rows D0: x_{k+2}
            for (int r = threadIdx.x; r < n0; r += NT) {
                const uint32_t adj = sea[r];
                const uint32_t ca = adj & 0xFFFFu, cb = adj >> 16;
                const double bv = sb[r], pv = spv[r];
                double val = sc3[3 * (ca >> 2) + (ca & 3u)];
                if (cb != 0xFFFFu) val += sc3[3 * (cb >> 2) + (cb & 3u)];
                const double xa = sx[r];
                const double res = bv - val;
                x2[h.d_start + r] = xa + c2b * pv * res + c1b * (xa - sxk0[r]);
                rr1 += pv * res * res;
            }
        }
        __syncthreads();
        block_sum3(rr0, rr1, bb, red);
        double* part = a.partials + (size_t)(pass & 1) * 3 * gridDim.x;
        if (threadIdx.x == 0) { part[3 * blockIdx.x] = rr0; part[3 * blockIdx.x + 1] = rr1; part[3 * blockIdx.x + 2] = bb; }
        grid_barrier_px(&a.ctl->gbar[0], epoch);
        if (threadIdx.x < 32) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) { s0 += __ldcg(&part[3 * i]); s1 += __ldcg(&part[3 * i + 1]); s2 += __ldcg(&part[3 * i + 2]); }
            s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
            if (threadIdx.x == 0) { bc[0] = s0; bc[1] = s1; bc[2] = s2; }
        }
        __syncthreads();
        const double g0r = bc[0], g1r = bc[1];
        bbv = bc[2];
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0 && k + 1 < 48) {
            a.ctl->reshist[a.which][k] = (bbv > 0.0) ? g0r / bbv : 0.0;
            a.ctl->reshist[a.which][k + 1] = (bbv > 0.0) ? g1r / bbv : 0.0;
        }
        ++pass;
        // both residuals are true residuals of stored iterates: accept the first one that meets the tolerance
        if (g0r <= a.tol2 * bbv) { conv = true; rr = g0r; k += 1; break; }
        if (g1r <= a.tol2 * bbv) { conv = true; rr = g1r; k += 2; cur = (cur + 1) & 3; break; }
        rr = g1r; k += 2; rho = rho_b; cur = (cur + 2) & 3;
        if (k >= a.max_iter) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc.cur = cur; sc.k = k; sc.done = 1; sc.rr = rr; sc.bb = bbv; sc.rho = rho;
        sc.iters += k; sc.solves += 1;
        if (!conv) a.ctl->nonconv += 1;
        const double rel = (bbv > 0.0) ? sqrt(rr / bbv) : 0.0;
        if (rel > sc.maxres) sc.maxres = rel;
    }
}

// ------------------------------------------------------------------------------------------------
// M_p (flags bit 14, experimental): the assembled sliced-ELL rows of the ring-2 patch ([D0 | G1], layout.hpp) staged by bulk
// copies; sweep 1 forms x_{k+1} on D0+G1 into shared memory, sweep 2 forms x_{k+2} on D0 from the same staged rows.  DRAM per
// pass (two iterations): the ELL rows of D0+G1, x_k, b, x_{k-1} once, two iterate writes: ~33 instead of 48 B per cell and iteration.
// Stage = [slice offsets ids | slice offsets weights | x_k on D0|G1|G2 | b rows | x_{k-1} rows | weights | ids]; then x_{k+1} on the
// rows and two ghost-index lists.
// ------------------------------------------------------------------------------------------------
struct Ell2PipeArgs {
    Ring2Dev m;
    double* buf[4];
    const double* b;
    double cheb_d, cheb_c, tol2;
    int max_iter, which;
    double* partials;
    Ctl* ctl;
    uint32_t off_wo, off_x, off_b, off_xprev, off_w, off_ids, stage_bytes, off_x1, off_gidx, gidx_bytes;
};

template <int NT>
__device__ __forceinline__ void ell2_issue(const Ell2PipeArgs& a, unsigned char* stage, const uint32_t* gidx_this, uint32_t* gidx_next,
                                           const Ring2Hdr& hn, const Ring2Hdr* hn2, const double* xk, const double* xprev,
                                           bool use_prev, unsigned long long* bar) {
    const Ring2Dev& m = a.m;
    uint32_t* s_io = reinterpret_cast<uint32_t*>(stage);
    uint32_t* s_wo = reinterpret_cast<uint32_t*>(stage + a.off_wo);
    double* sx = reinterpret_cast<double*>(stage + a.off_x);
    double* sb = reinterpret_cast<double*>(stage + a.off_b);
    double* sxp = reinterpret_cast<double*>(stage + a.off_xprev);
    const uint32_t* gh = reinterpret_cast<const uint32_t*>(m.ghost);
    const int ds = hn.d_start, n0 = hn.d_cnt, nr = n0 + hn.g1_cnt, ng = hn.g1_cnt + hn.g2_cnt;
    if (threadIdx.x == 0) {
        const uint32_t id_bytes = (uint32_t)hn.ell_in * 2u, w_bytes = (uint32_t)hn.ell_wn * 8u;
        uint32_t bytes = f64s_bytes(ds, n0) * (use_prev ? 3u : 2u) + id_bytes + w_bytes;
        if (hn2) bytes += u32s_bytes(hn2->g_off, hn2->g1_cnt + hn2->g2_cnt);
        mbar_arrive_expect_tx(bar, bytes);
        if (w_bytes) bulk_g2s(stage + a.off_w, m.ell_w + (uint32_t)hn.ell_w0, w_bytes, bar);
        if (id_bytes) bulk_g2s(stage + a.off_ids, m.ell_ids + (uint32_t)hn.ell_i0, id_bytes, bar);
        f64s_bulk(sx, xk, ds, n0, bar);
        f64s_bulk(sb, a.b, ds, n0, bar);
        if (use_prev) f64s_bulk(sxp, xprev, ds, n0, bar);
        if (hn2) u32s_bulk(gidx_next, gh, hn2->g_off, hn2->g1_cnt + hn2->g2_cnt, bar);
    }
    const int lane = (int)threadIdx.x - (NT - 32);
    if (lane >= 0) {
        if (lane < 2) f64s_fix(sx, xk, ds, n0, lane);
        else if (lane < 4) f64s_fix(sb, a.b, ds, n0, lane - 2);
        else if (lane < 6) { if (use_prev) f64s_fix(sxp, xprev, ds, n0, lane - 4); }
        else if (lane < 12) { if (hn2) u32s_fix(gidx_next, gh, hn2->g_off, hn2->g1_cnt + hn2->g2_cnt, lane - 6); }
    }
    const int nsl = (nr + 31) >> 5;
    for (int i = threadIdx.x; i <= nsl; i += NT) { cp_async4(&s_io[i], &m.ell_ioff[hn.vs_off + i]); cp_async4(&s_wo[i], &m.ell_woff[hn.vs_off + i]); }
    double* sxg = sx + (ds & 1) + n0;
    double* sbg = sb + (ds & 1) + n0;
    double* sxpg = sxp + (ds & 1) + n0;
    const uint32_t* gi = gidx_this ? gidx_this + (hn.g_off & 3) : nullptr;
    for (int i = threadIdx.x; i < ng; i += NT) {
        const uint32_t g = gi ? gi[i] : (uint32_t)__ldg(&m.ghost[hn.g_off + i]);
        cp_async8(&sxg[i], &xk[g]);
        if (i < hn.g1_cnt) {
            cp_async8(&sbg[i], &a.b[g]);
            if (use_prev) cp_async8(&sxpg[i], &xprev[g]);
        }
    }
}

// one ELL row from staged ids / weights: val = sum_w m_vw x_w (without the diagonal), dsum = sum_w m_vw (= the diagonal)
__device__ __forceinline__ void ell_row_s(const uint4* sid4, const double* sw, uint32_t io0, uint32_t wo0, int W, int lane, const double* x,
                                          double& val, double& dsum) {
    val = 0.0; dsum = 0.0;
    for (int c = 0; 8 * c < W; ++c) {
        const uint4 cd = sid4[(io0 >> 3) + c * 32 + lane];
        const uint32_t w4[4] = {cd.x, cd.y, cd.z, cd.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double w0 = (8 * c + 2 * j < W) ? sw[wo0 + (8 * c + 2 * j) * 32 + lane] : 0.0;
            const double w1 = (8 * c + 2 * j + 1 < W) ? sw[wo0 + (8 * c + 2 * j + 1) * 32 + lane] : 0.0;
            val += w0 * x[w4[j] & 0xFFFFu];
            val += w1 * x[w4[j] >> 16];
            dsum += w0 + w1;
        }
    }
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) solve2_ell_pipe_k(const __grid_constant__ Ell2PipeArgs a) {
    PHW_DYN_SMEM(smraw);
    __shared__ double red[96];
    __shared__ double bc[3];
    __shared__ Ring2Hdr s_hdr[4];
    __shared__ __align__(8) unsigned long long full[2];
    __shared__ int s_dead;
    SolveCtl& sc = a.ctl->sc[a.which];
    int k = 0, cur = sc.cur & 3, pass = 0;
    double rho = 0.0, rr = 0.0, bbv = 0.0;
    bool conv = false;
    unsigned int epoch = 0;
    unsigned int seq = 0;
    if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); s_dead = 0; }
    __syncthreads();
    double* sx1 = reinterpret_cast<double*>(smraw + a.off_x1);      // x_{k+1} on the rows [D0 | G1] (ring-2 local ids)
    const int npatch = a.m.npatch;
    const int n_my = ((int)blockIdx.x < npatch) ? (npatch - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    while (true) {
        double c1a, c2a, c1b, c2b, rho_a, rho_b;
        cheb_coeffs(k, rho, a.cheb_d, a.cheb_c, c1a, c2a, rho_a);
        cheb_coeffs(k + 1, rho_a, a.cheb_d, a.cheb_c, c1b, c2b, rho_b);
        const double* xk = sel4(a.buf, cur);
        double* x1 = sel4(a.buf, (cur + 1) & 3);
        double* x2 = sel4(a.buf, (cur + 2) & 3);
        const double* xprev = sel4(a.buf, (cur + 3) & 3);
        const bool use_prev = k > 0;
        double rr0 = 0.0, rr1 = 0.0, bb = 0.0;
        for (int j = 0; j < 3 && j < n_my; ++j) load_hdr2_smem_at(&s_hdr[j], &a.m.hdr[blockIdx.x + j * gridDim.x], 32 * j);
        __syncthreads();
        if (n_my > 0)
            ell2_issue<NT>(a, smraw + (seq & 1u) * a.stage_bytes, nullptr,
                           reinterpret_cast<uint32_t*>(smraw + a.off_gidx + ((seq + 1u) & 1u) * a.gidx_bytes),
                           s_hdr[0], n_my > 1 ? &s_hdr[1] : nullptr, xk, xprev, use_prev, &full[seq & 1u]);
        for (int it = 0; it < n_my; ++it, ++seq) {
            const unsigned int s = seq & 1u;
            if (!pipe_stage_ready(&full[s], (seq >> 1) & 1u, a.ctl, &s_dead)) continue;
            const Ring2Hdr& h = s_hdr[it & 3];
            if (it + 3 < n_my) load_hdr2_smem_at(&s_hdr[(it + 3) & 3], &a.m.hdr[blockIdx.x + (it + 3) * gridDim.x], 64);
            if (it + 1 < n_my)
                ell2_issue<NT>(a, smraw + (s ^ 1u) * a.stage_bytes,
                               reinterpret_cast<const uint32_t*>(smraw + a.off_gidx + (s ^ 1u) * a.gidx_bytes),
                               reinterpret_cast<uint32_t*>(smraw + a.off_gidx + s * a.gidx_bytes),
                               s_hdr[(it + 1) & 3], it + 2 < n_my ? &s_hdr[(it + 2) & 3] : nullptr, xk, xprev, use_prev, &full[s ^ 1u]);
            const unsigned char* stage = smraw + s * a.stage_bytes;
            const int n0 = h.d_cnt, nr = n0 + h.g1_cnt;
            const uint32_t* s_io = reinterpret_cast<const uint32_t*>(stage);
            const uint32_t* s_wo = reinterpret_cast<const uint32_t*>(stage + a.off_wo);
            const double* sx = reinterpret_cast<const double*>(stage + a.off_x) + (h.d_start & 1);
            const double* sb = reinterpret_cast<const double*>(stage + a.off_b) + (h.d_start & 1);
            const double* sxp = reinterpret_cast<const double*>(stage + a.off_xprev) + (h.d_start & 1);
            const double* sw = reinterpret_cast<const double*>(stage + a.off_w);
            const uint4* sid4 = reinterpret_cast<const uint4*>(stage + a.off_ids);
            const uint32_t i0 = (uint32_t)h.ell_i0, w0 = (uint32_t)h.ell_w0;
            // ---- sweep 1: rows [D0 | G1] -> x_{k+1} in shared memory
            for (int r = threadIdx.x; r < nr; r += NT) {
                const int sl = r >> 5, lane = r & 31;
                const uint32_t io0 = s_io[sl] - i0, wo0 = s_wo[sl] - w0;
                const int W = (int)((s_wo[sl + 1] - w0 - wo0) >> 5);
                double val, dsum;
                ell_row_s(sid4, sw, io0, wo0, W, lane, sx, val, dsum);
                const double xc = sx[r], bv = sb[r];
                const double pv = 1.0 / dsum;
                val += dsum * xc;
                const double res = bv - val;
                double xn = xc + c2a * pv * res;
                if (use_prev) xn += c1a * (xc - sxp[r]);
                sx1[r] = xn;
                if (r < n0) {
                    x1[h.d_start + r] = xn;
                    rr0 += pv * res * res;
                    bb += pv * bv * bv;
                }
            }
            __syncthreads();
            // ---- sweep 2: rows D0 -> x_{k+2} (their neighbours are rows of sweep 1)
            for (int r = threadIdx.x; r < n0; r += NT) {
                const int sl = r >> 5, lane = r & 31;
                const uint32_t io0 = s_io[sl] - i0, wo0 = s_wo[sl] - w0;
                const int W = (int)((s_wo[sl + 1] - w0 - wo0) >> 5);
                double val, dsum;
                ell_row_s(sid4, sw, io0, wo0, W, lane, sx1, val, dsum);
                const double xa = sx1[r];
                const double pv = 1.0 / dsum;
                val += dsum * xa;
                const double res = sb[r] - val;
                x2[h.d_start + r] = xa + c2b * pv * res + c1b * (xa - sx[r]);
                rr1 += pv * res * res;
            }
        }
        __syncthreads();
        block_sum3(rr0, rr1, bb, red);
        double* part = a.partials + (size_t)(pass & 1) * 3 * gridDim.x;
        if (threadIdx.x == 0) { part[3 * blockIdx.x] = rr0; part[3 * blockIdx.x + 1] = rr1; part[3 * blockIdx.x + 2] = bb; }
        grid_barrier_px(&a.ctl->gbar[0], epoch);
        if (threadIdx.x < 32) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) { s0 += __ldcg(&part[3 * i]); s1 += __ldcg(&part[3 * i + 1]); s2 += __ldcg(&part[3 * i + 2]); }
            s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
            if (threadIdx.x == 0) { bc[0] = s0; bc[1] = s1; bc[2] = s2; }
        }
        __syncthreads();
        const double g0r = bc[0], g1r = bc[1];
        bbv = bc[2];
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0 && k + 1 < 48) {
            a.ctl->reshist[a.which][k] = (bbv > 0.0) ? g0r / bbv : 0.0;
            a.ctl->reshist[a.which][k + 1] = (bbv > 0.0) ? g1r / bbv : 0.0;
        }
        ++pass;
        if (g0r <= a.tol2 * bbv) { conv = true; rr = g0r; k += 1; break; }
        if (g1r <= a.tol2 * bbv) { conv = true; rr = g1r; k += 2; cur = (cur + 1) & 3; break; }
        rr = g1r; k += 2; rho = rho_b; cur = (cur + 2) & 3;
        if (k >= a.max_iter) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc.cur = cur; sc.k = k; sc.done = 1; sc.rr = rr; sc.bb = bbv; sc.rho = rho;
        sc.iters += k; sc.solves += 1;
        if (!conv) a.ctl->nonconv += 1;
        const double rel = (bbv > 0.0) ? sqrt(rr / bbv) : 0.0;
        if (rel > sc.maxres) sc.maxres = rel;
    }
}

}  // namespace phw
