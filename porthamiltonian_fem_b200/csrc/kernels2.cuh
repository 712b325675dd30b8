// kernels2.cuh - two-sweep (temporally blocked) persistent Chebyshev solves of the mass systems
//   M_p v = b_p (vertex rows)  /  M_q v = b_q (edge rows)          (replaces solve(A, x, b), wave_2D.py:820,854)
//
// The mass solves are bound by DRAM traffic, and most of the bytes of an iteration are iteration-invariant (cell
// records, geometry, adjacency, Jacobi weights, right-hand side).  A two-sweep patch (layout.hpp, Ring2) carries one more
// ring of halo cells, so that a CTA which has staged x_k on D0+G1+G2 can form x_{k+1} on D0+G1 and from it x_{k+2} on D0
// without leaving shared memory: two Chebyshev iterations per pass over the static data.  The arithmetic is the same
// three-term recurrence as in solve_coop_k (kernels.cuh) - same spectral bounds, same iterates up to the summation order,
// true residuals of BOTH iterates measured for the stop test - only the schedule differs.  Rows recomputed in a halo
// (G1) use the same summation order as their owner patch, hence bit-identical values.
//
// Four rotating iterate buffers: x_{k-1} (read on D0+G1), x_k (read), x_{k+1}, x_{k+2} (written on D0).
#pragma once
#include "kernels.cuh"

namespace phw {

struct Ring2Dev {
    int npatch;
    const Ring2Hdr* hdr;
    const uint2* rec;
    const double* g0; const double* g1; const double* g2;   // vertex kind: g0 = cell area, g1 = g2 = nullptr
    const int* ghost;
    const double* pinv;          // [n_rows] Jacobi weights replicated per first-sweep row (contiguous per patch)
    const uint32_t* ell_ioff; const uint32_t* ell_woff;     // vertex kind: assembled mass rows of [D0 | G1] (sliced ELL)
    const uint16_t* ell_ids; const double* ell_w;
    const uint32_t* eadj;                                   // edge kind
};

struct Solve2Args {
    Ring2Dev m;
    double* buf[4];
    const double* b;
    double cheb_d, cheb_c, tol2;
    int max_iter, which;
    double* partials;            // [2][3 * gridDim.x]
    Ctl* ctl;
    int gbar;                    // 1: grid_barrier instead of cooperative-groups grid syncs (phw_params.flags bit 10)
};

__device__ __forceinline__ double* sel4(double* const (&b)[4], int i) { return i == 0 ? b[0] : (i == 1 ? b[1] : (i == 2 ? b[2] : b[3])); }

struct Pass2IO {
    const double* xk; const double* xprev; double* x1; double* x2;
    double c1a, c2a, c1b, c2b;
    int use_prev;
};

__device__ __forceinline__ void load_hdr2_smem(Ring2Hdr* dst, const Ring2Hdr* src) {
    if (threadIdx.x < 20) reinterpret_cast<int*>(dst)[threadIdx.x] = __ldg(reinterpret_cast<const int*>(src) + threadIdx.x);
}

// RT1 element mass row values of one cell (unsigned-basis metric form, see edge_pass in kernels.cuh)
__device__ __forceinline__ void rt1_cell(const uint2 re, const double g0, const double g1, const double g2, const double* __restrict__ sx,
                                         double* __restrict__ out3) {
    const uint32_t e0 = re.x & 0xFFFFu, e1 = re.x >> 16, e2 = re.y & 0xFFFFu;
    const double S = g0 + g1 + g2;
    const double s0 = (e0 & 0x8000u) ? -1.0 : 1.0, s1 = (e1 & 0x8000u) ? -1.0 : 1.0, s2 = (e2 & 0x8000u) ? -1.0 : 1.0;
    const double t0 = s0 * sx[e0 & 0x7FFFu], t1 = s1 * sx[e1 & 0x7FFFu], t2 = s2 * sx[e2 & 0x7FFFu];
    const double ST = S * (t0 + t1 + t2);
    out3[0] = s0 * (ST + 2.0 * S * t0 - 4.0 * (g0 * t0 + g2 * t1 + g1 * t2));
    out3[1] = s1 * (ST + 2.0 * S * t1 - 4.0 * (g2 * t0 + g1 * t1 + g0 * t2));
    out3[2] = s2 * (ST + 2.0 * S * t2 - 4.0 * (g1 * t0 + g0 * t1 + g2 * t2));
}

// ------------------------------------------------------------------------------------------------
// edge rows (M_q): all two-sweep patches of this block
// ------------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void edge_pass2(const Solve2Args& a, const Pass2IO& io, double* smem, double& rr0, double& rr1, double& bb) {
    const Ring2Dev& m = a.m;
    __shared__ Ring2Hdr s_hdr[2];
    if (blockIdx.x < m.npatch) load_hdr2_smem(&s_hdr[0], &m.hdr[blockIdx.x]);
    __syncthreads();
    int it = 0;
    for (int P = blockIdx.x; P < m.npatch; P += gridDim.x, ++it) {
        const Ring2Hdr& h = s_hdr[it & 1];
        if (P + (int)gridDim.x < m.npatch) load_hdr2_smem(&s_hdr[(it + 1) & 1], &m.hdr[P + gridDim.x]);
        const int n0 = h.d_cnt, nr = n0 + h.g1_cnt, ng = h.g1_cnt + h.g2_cnt;
        double* sx = smem;                               // [n0 + ng]  x_k, overwritten row by row with x_{k+1} on [0, nr)
        double* sc3 = sx + ((n0 + ng + 1) & ~1);         // [3 * n_c2]
        // ---- stage 1: x_k on D0 (coalesced) and on G1+G2 (gathered)
        for (int i = threadIdx.x; i < n0; i += NT) sx[i] = io.xk[h.d_start + i];
        for (int i = threadIdx.x; i < ng; i += NT) sx[n0 + i] = io.xk[__ldg(&m.ghost[h.g_off + i])];
        __syncthreads();
        // ---- sweep 1, cells C2: loads of U cells are issued before they are consumed
        {
            constexpr int U = 2;
            for (int base = threadIdx.x; base < h.n_c2; base += U * NT) {
                uint2 reb[U]; double gb[U][3];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int lc = base + u * NT;
                    if (lc < h.n_c2) {
                        reb[u] = __ldg(&m.rec[h.cell_off + lc]);
                        gb[u][0] = __ldg(&m.g0[h.cell_off + lc]); gb[u][1] = __ldg(&m.g1[h.cell_off + lc]); gb[u][2] = __ldg(&m.g2[h.cell_off + lc]);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int lc = base + u * NT;
                    if (lc < h.n_c2) rt1_cell(reb[u], gb[u][0], gb[u][1], gb[u][2], sx, &sc3[3 * lc]);
                }
            }
        }
        __syncthreads();
        // ---- sweep 1, rows [D0 | G1]: x_{k+1}, kept in shared memory (in place: a row's x_k is read by its own thread only)
        {
            constexpr int U = 2;
            for (int base = threadIdx.x; base < nr; base += U * NT) {
                uint32_t adjb[U]; double bvb[U], pvb[U], xpb[U]; int gidb[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = base + u * NT;
                    if (r < nr) {
                        gidb[u] = (r < n0) ? h.d_start + r : __ldg(&m.ghost[h.g_off + r - n0]);
                        adjb[u] = __ldg(&m.eadj[h.row_off + r]); pvb[u] = __ldg(&m.pinv[h.row_off + r]);
                        bvb[u] = __ldg(&a.b[gidb[u]]);
                        xpb[u] = io.use_prev ? io.xprev[gidb[u]] : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = base + u * NT;
                    if (r < nr) {
                        const uint32_t ca = adjb[u] & 0xFFFFu, cb = adjb[u] >> 16;
                        double val = sc3[3 * (ca >> 2) + (ca & 3u)];
                        if (cb != 0xFFFFu) val += sc3[3 * (cb >> 2) + (cb & 3u)];
                        const double xc = sx[r];
                        const double res = bvb[u] - val;
                        double xn = xc + io.c2a * pvb[u] * res;
                        if (io.use_prev) xn += io.c1a * (xc - xpb[u]);
                        sx[r] = xn;
                        if (r < n0) {
                            io.x1[gidb[u]] = xn;
                            rr0 += pvb[u] * res * res;
                            bb += pvb[u] * bvb[u] * bvb[u];
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- sweep 2, cells C1 (all their edges are rows of sweep 1)
        {
            constexpr int U = 2;
            for (int base = threadIdx.x; base < h.n_c1; base += U * NT) {
                uint2 reb[U]; double gb[U][3];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int lc = base + u * NT;
                    if (lc < h.n_c1) {
                        reb[u] = __ldg(&m.rec[h.cell_off + lc]);
                        gb[u][0] = __ldg(&m.g0[h.cell_off + lc]); gb[u][1] = __ldg(&m.g1[h.cell_off + lc]); gb[u][2] = __ldg(&m.g2[h.cell_off + lc]);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int lc = base + u * NT;
                    if (lc < h.n_c1) rt1_cell(reb[u], gb[u][0], gb[u][1], gb[u][2], sx, &sc3[3 * lc]);
                }
            }
        }
        __syncthreads();
        // ---- sweep 2, rows D0: x_{k+2}
        {
            constexpr int U = 2;
            for (int base = threadIdx.x; base < n0; base += U * NT) {
                uint32_t adjb[U]; double bvb[U], pvb[U], xkb[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = base + u * NT;
                    if (r < n0) {
                        adjb[u] = __ldg(&m.eadj[h.row_off + r]); pvb[u] = __ldg(&m.pinv[h.row_off + r]);
                        bvb[u] = __ldg(&a.b[h.d_start + r]); xkb[u] = io.xk[h.d_start + r];
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = base + u * NT;
                    if (r < n0) {
                        const uint32_t ca = adjb[u] & 0xFFFFu, cb = adjb[u] >> 16;
                        double val = sc3[3 * (ca >> 2) + (ca & 3u)];
                        if (cb != 0xFFFFu) val += sc3[3 * (cb >> 2) + (cb & 3u)];
                        const double x1 = sx[r];
                        const double res = bvb[u] - val;
                        io.x2[h.d_start + r] = x1 + io.c2b * pvb[u] * res + io.c1b * (x1 - xkb[u]);
                        rr1 += pvb[u] * res * res;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// vertex rows (M_p) on the assembled sliced-ELL rows of the ring-2 patch (layout.hpp): (M_p x)_v = d_v x_v + sum_w m_vw x_w,
// d_v = sum_w m_vw.  Sweep 1 reads the rows of [D0 | G1] from global memory and parks those of D0 (ids + weights, 64 B per
// row) and b in shared memory; sweep 2 runs entirely from shared memory.  DRAM per pass (two iterations): the ELL rows of
// D0+G1 once, x_k, b, x_{k-1} once, two iterate writes.
// ------------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void vertex_ell_pass2(const Solve2Args& a, const Pass2IO& io, double* smem, double& rr0, double& rr1, double& bb) {
    const Ring2Dev& m = a.m;
    __shared__ Ring2Hdr s_hdr[2];
    if (blockIdx.x < m.npatch) load_hdr2_smem(&s_hdr[0], &m.hdr[blockIdx.x]);
    __syncthreads();
    const uint4* ids4 = reinterpret_cast<const uint4*>(m.ell_ids);
    int it = 0;
    for (int P = blockIdx.x; P < m.npatch; P += gridDim.x, ++it) {
        const Ring2Hdr& h = s_hdr[it & 1];
        if (P + (int)gridDim.x < m.npatch) load_hdr2_smem(&s_hdr[(it + 1) & 1], &m.hdr[P + gridDim.x]);
        const int n0 = h.d_cnt, nr = n0 + h.g1_cnt, ng = h.g1_cnt + h.g2_cnt, nl = n0 + ng;
        const int nsl = (nr + 31) >> 5, nsl0 = (n0 + 31) >> 5;
        const uint32_t iobase = __ldg(&m.ell_ioff[h.vs_off]), wobase = __ldg(&m.ell_woff[h.vs_off]);
        const uint32_t wcount0 = __ldg(&m.ell_woff[h.vs_off + nsl0]) - wobase;
        uint32_t* s_io = reinterpret_cast<uint32_t*>(smem);   // [nsl + 1]
        uint32_t* s_wo = s_io + (nsl + 1);                     // [nsl + 1]
        double* sx = smem + ((nsl + 3) & ~1);                  // [nl]  x_k (even offset: the ids below are read as uint4)
        double* sx1 = sx + ((nl + 1) & ~1);                    // [nr]  x_{k+1}
        double* sb = sx1 + ((nr + 1) & ~1);                    // [n0]  b on D0
        double* sw = sb + ((n0 + 1) & ~1);                     // weights of the D0 slices (slice layout kept)
        uint4* sid4 = reinterpret_cast<uint4*>(sw + wcount0);  // ids of the D0 slices (wcount0 is a multiple of 32)
        for (int i = threadIdx.x; i <= nsl; i += NT) { s_io[i] = __ldg(&m.ell_ioff[h.vs_off + i]); s_wo[i] = __ldg(&m.ell_woff[h.vs_off + i]); }
        for (int i = threadIdx.x; i < n0; i += NT) sx[i] = io.xk[h.d_start + i];
        for (int i = threadIdx.x; i < ng; i += NT) sx[n0 + i] = io.xk[__ldg(&m.ghost[h.g_off + i])];
        __syncthreads();
        // ---- sweep 1: rows [D0 | G1] from global memory
        const int rround = (nr + 31) & ~31;
        for (int r = threadIdx.x; r < rround; r += NT) {
            const int sl = r >> 5, lane = r & 31;
            const uint32_t io0 = s_io[sl], wo0 = s_wo[sl];
            const int W = (int)((s_wo[sl + 1] - wo0) >> 5);
            const bool live = r < nr, keep = sl < nsl0;
            int gid = 0; double bv = 0.0, xpv = 0.0;
            if (live) {
                gid = (r < n0) ? h.d_start + r : __ldg(&m.ghost[h.g_off + r - n0]);
                bv = __ldg(&a.b[gid]);
                if (io.use_prev) xpv = io.xprev[gid];
            }
            const double* __restrict__ wrow = m.ell_w + wo0 + lane;
            double val = 0.0, dsum = 0.0;
            for (int c = 0; 8 * c < W; ++c) {
                const uint4 cd = __ldg(&ids4[(io0 >> 3) + c * 32 + lane]);
                const uint32_t w4[4] = {cd.x, cd.y, cd.z, cd.w};
                double wv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) wv[j] = (8 * c + j < W) ? __ldg(&wrow[(8 * c + j) * 32]) : 0.0;
                if (keep) {
                    sid4[((io0 - iobase) >> 3) + c * 32 + lane] = cd;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (8 * c + j < W) sw[(wo0 - wobase) + (8 * c + j) * 32 + lane] = wv[j];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    val += wv[2 * j] * sx[w4[j] & 0xFFFFu];
                    val += wv[2 * j + 1] * sx[w4[j] >> 16];
                    dsum += wv[2 * j] + wv[2 * j + 1];
                }
            }
            if (!live) continue;
            const double xc = sx[r];
            const double pv = 1.0 / dsum;
            val += dsum * xc;
            const double res = bv - val;
            double xn = xc + io.c2a * pv * res;
            if (io.use_prev) xn += io.c1a * (xc - xpv);
            sx1[r] = xn;
            if (r < n0) {
                io.x1[gid] = xn;
                sb[r] = bv;
                rr0 += pv * res * res;
                bb += pv * bv * bv;
            }
        }
        __syncthreads();
        // ---- sweep 2: rows D0 from shared memory
        for (int r = threadIdx.x; r < n0; r += NT) {
            const int sl = r >> 5, lane = r & 31;
            const uint32_t io0 = s_io[sl] - iobase, wo0 = s_wo[sl] - wobase;
            const int W = (int)((s_wo[sl + 1] - s_wo[sl]) >> 5);
            double val = 0.0, dsum = 0.0;
            for (int c = 0; 8 * c < W; ++c) {
                const uint4 cd = sid4[(io0 >> 3) + c * 32 + lane];
                const uint32_t w4[4] = {cd.x, cd.y, cd.z, cd.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double w0 = (8 * c + 2 * j < W) ? sw[wo0 + (8 * c + 2 * j) * 32 + lane] : 0.0;
                    const double w1 = (8 * c + 2 * j + 1 < W) ? sw[wo0 + (8 * c + 2 * j + 1) * 32 + lane] : 0.0;
                    val += w0 * sx1[w4[j] & 0xFFFFu];
                    val += w1 * sx1[w4[j] >> 16];
                    dsum += w0 + w1;
                }
            }
            const double x1 = sx1[r];
            const double pv = 1.0 / dsum;
            val += dsum * x1;
            const double res = sb[r] - val;
            io.x2[h.d_start + r] = x1 + io.c2b * pv * res + io.c1b * (x1 - sx[r]);
            rr1 += pv * res * res;
        }
        __syncthreads();
    }
}

// deterministic block reduction of three values; result valid in thread 0
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* red /* >= 96 doubles */) {
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { red[w] = a; red[32 + w] = b; red[64 + w] = c; }
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        a = (l < nw) ? red[l] : 0.0; b = (l < nw) ? red[32 + l] : 0.0; c = (l < nw) ? red[64 + l] : 0.0;
        a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    }
}

// KIND = 0: vertex rows (M_p), 1: edge rows (M_q).  One grid-wide barrier per pass (= two iterations).
template <int KIND, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) solve2_coop_k(const __grid_constant__ Solve2Args a) {
    extern __shared__ double smem[];
    __shared__ double red[96];
    __shared__ double bc[3];
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    SolveCtl& sc = a.ctl->sc[a.which];
    int k = 0, cur = sc.cur & 3, pass = 0;
    double rho = 0.0, rr = 0.0, bbv = 0.0;
    bool conv = false;
    unsigned int epoch = 0;
    while (true) {
        Pass2IO io;
        double rho_a, rho_b;
        cheb_coeffs(k, rho, a.cheb_d, a.cheb_c, io.c1a, io.c2a, rho_a);
        cheb_coeffs(k + 1, rho_a, a.cheb_d, a.cheb_c, io.c1b, io.c2b, rho_b);
        io.xk = sel4(a.buf, cur); io.x1 = sel4(a.buf, (cur + 1) & 3); io.x2 = sel4(a.buf, (cur + 2) & 3); io.xprev = sel4(a.buf, (cur + 3) & 3);
        io.use_prev = (k > 0);
        double r0 = 0.0, r1 = 0.0, b0 = 0.0;
        if (KIND == 0) vertex_ell_pass2<NT>(a, io, smem, r0, r1, b0);
        else edge_pass2<NT>(a, io, smem, r0, r1, b0);
        block_sum3(r0, r1, b0, red);
        double* part = a.partials + (size_t)(pass & 1) * 3 * gridDim.x;
        if (threadIdx.x == 0) { part[3 * blockIdx.x] = r0; part[3 * blockIdx.x + 1] = r1; part[3 * blockIdx.x + 2] = b0; }
        PHW_GRID_SYNC(a);
        if (threadIdx.x < 32) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) { s0 += __ldcg(&part[3 * i]); s1 += __ldcg(&part[3 * i + 1]); s2 += __ldcg(&part[3 * i + 2]); }
            s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
            if (threadIdx.x == 0) { bc[0] = s0; bc[1] = s1; bc[2] = s2; }
        }
        __syncthreads();
        const double rr0 = bc[0], rr1 = bc[1];
        bbv = bc[2];
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0 && k + 1 < 48) {
            a.ctl->reshist[a.which][k] = (bbv > 0.0) ? rr0 / bbv : 0.0;
            a.ctl->reshist[a.which][k + 1] = (bbv > 0.0) ? rr1 / bbv : 0.0;
        }
        ++pass;
        // both residuals are true residuals of stored iterates: accept the first one that meets the tolerance
        if (rr0 <= a.tol2 * bbv) { conv = true; rr = rr0; k += 1; break; }
        if (rr1 <= a.tol2 * bbv) { conv = true; rr = rr1; k += 2; cur = (cur + 1) & 3; break; }
        rr = rr1; k += 2; rho = rho_b; cur = (cur + 2) & 3;
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
