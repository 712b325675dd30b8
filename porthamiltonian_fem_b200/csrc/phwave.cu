// phwave.cu - C ABI (include/phwave.h) + device-resident time loop of the port-Hamiltonian wave solver.
//
// Reference being replaced: fenics_projects/wave_eqn/wave_2D.py (setup :92-641, loop :787-982).
// The host side of this file only *enqueues* kernels on one stream; all loop state (time, lumped ODE
// state, energies, Chebyshev iteration control) lives in a device-resident control block (Ctl), so the
// step loop never returns to the host.  See DESIGN.md for the data layout and per-kernel byte model.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <numeric>
#include <set>
#include <string>
#include <vector>

#include "../../include/phwave.h"
#include "kernels.cuh"
#include "kernels_pipe.cuh"
#include "pipe_carve.hpp"
#include "layout.hpp"

using namespace phw;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf_[512];                                                                        \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return fail(PHW_ECUDA, buf_);                                                          \
        }                                                                                          \
    } while (0)

struct phw_mesh_s { Layout L; };

// ================================================================================================
// small kernels
// ================================================================================================
namespace phw {

// per-(patch,cell) geometry: area and the RT1 metric g_i = |e_i|^2 / (48 |K|); also the element-wise
// generalized eigenvalue range of (M_K, diag M_K), which bounds the spectrum of the Jacobi-scaled RT1
// mass matrix (Wathen 1987) - the Chebyshev interval.
__global__ void geom_k(int npc, const int* __restrict__ pcell_user, const int* __restrict__ cells,
                       const double* __restrict__ vtx, double* area, double* g0, double* g1, double* g2,
                       unsigned long long* lam_minmax) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    double lmin = 1.0, lmax = 1.0;
    if (k < npc) {
        int c = pcell_user[k];
        int a = cells[3 * c], b = cells[3 * c + 1], d = cells[3 * c + 2];
        double ax = vtx[2 * a], ay = vtx[2 * a + 1], bx = vtx[2 * b], by = vtx[2 * b + 1], dx = vtx[2 * d], dy = vtx[2 * d + 1];
        double A = 0.5 * fabs((bx - ax) * (dy - ay) - (by - ay) * (dx - ax));
        double l0 = (dx - bx) * (dx - bx) + (dy - by) * (dy - by);   // edge 0 joins v1, v2
        double l1 = (ax - dx) * (ax - dx) + (ay - dy) * (ay - dy);   // edge 1 joins v2, v0
        double l2 = (bx - ax) * (bx - ax) + (by - ay) * (by - ay);   // edge 2 joins v0, v1
        double inv = 1.0 / (48.0 * A);
        double G0 = l0 * inv, G1 = l1 * inv, G2 = l2 * inv, S = G0 + G1 + G2;
        if (area) area[k] = A;
        if (g0) { g0[k] = G0; g1[k] = G1; g2[k] = G2; }
        double m00 = 3 * S - 4 * G0, m11 = 3 * S - 4 * G1, m22 = 3 * S - 4 * G2;
        double m01 = S - 4 * G2, m02 = S - 4 * G1, m12 = S - 4 * G0;
        double ea = m01 / sqrt(m00 * m11), eb = m02 / sqrt(m00 * m22), ec = m12 / sqrt(m11 * m22);
        double p = ea * ea + eb * eb + ec * ec, q = 2.0 * ea * eb * ec;
        if (p > 1e-30) {
            double arg = 0.5 * q * pow(3.0 / p, 1.5);
            arg = fmin(1.0, fmax(-1.0, arg));
            double phi = acos(arg) / 3.0, amp = 2.0 * sqrt(p / 3.0);
            lmax = 1.0 + amp * cos(phi);
            lmin = 1.0 + amp * cos(phi + 2.0943951023931953);
        }
    }
    // positive doubles order like their bit patterns
    for (int o = 16; o > 0; o >>= 1) {
        lmin = fmin(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&lam_minmax[0], (unsigned long long)__double_as_longlong(lmin));
        atomicMax(&lam_minmax[1], (unsigned long long)__double_as_longlong(lmax));
    }
}

// weights of the ELL mass rows: m_vw = (|K1| + |K2|) / 12 from the patch's own area array (one block per patch)
__global__ void ell_weights_k(int npatch, const PatchHdr* __restrict__ hdr, const uint32_t* __restrict__ woff,
                              const uint32_t* __restrict__ pair, const double* __restrict__ area, double* __restrict__ w) {
    for (int P = blockIdx.x; P < npatch; P += gridDim.x) {
        const PatchHdr h = hdr[P];
        const int nsl = (h.v_cnt + 31) >> 5;
        const uint32_t e0 = woff[h.vs_off], e1 = woff[h.vs_off + nsl];
        for (uint32_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
            const uint32_t pr = pair[e];
            double v = 0.0;
            if (pr != 0xFFFFFFFFu) {
                v = area[h.cell_off + (pr & 0xFFFFu)];
                if ((pr >> 16) != 0xFFFFu) v += area[h.cell_off + (pr >> 16)];
                v *= (1.0 / 12.0);
            }
            w[e] = v;
        }
    }
}

// packed edge-row cell records of the lean M_q solve: {rece.x, rece.y, g0} {g1, g2}
__global__ void pack_cellq_k(int npc, const uint2* __restrict__ rece, const double* __restrict__ g0, const double* __restrict__ g1,
                             const double* __restrict__ g2, uint4* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npc) return;
    const uint2 r = rece[k];
    const double a = g0[k], b = g1[k], c = g2[k];
    out[2 * (size_t)k] = make_uint4(r.x, r.y, (unsigned)__double2loint(a), (unsigned)__double2hiint(a));
    out[2 * (size_t)k + 1] = make_uint4((unsigned)__double2loint(b), (unsigned)__double2hiint(b), (unsigned)__double2loint(c), (unsigned)__double2hiint(c));
}

__global__ void invert_k(int n, double* d) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = (d[i] != 0.0) ? 1.0 / d[i] : 0.0;
}
__global__ void gather_k(int n, double* dst, const double* src, const int* map) {   // dst[i] = src[map[i]]
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[map[i]];
}
__global__ void scatter_k(int n, double* dst, const double* src, const int* map) {  // dst[map[i]] = src[i]
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[map[i]] = src[i];
}
// y += a * x      (x optionally = the current Chebyshev buffer of solve `which`)
__global__ void axpy_k(int n, double* y, double a, const double* x, TriBuf tb, const Ctl* ctl, int which) {
    const double* xs = x ? x : tb.b[ctl->sc[which].cur];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] += a * xs[i];
}
// w = y + a * x
__global__ void waxpy_k(int n, double* w, const double* y, double a, const double* x, TriBuf tb, const Ctl* ctl, int which) {
    const double* xs = x ? x : tb.b[ctl->sc[which].cur];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) w[i] = y[i] + a * xs[i];
}
// buf[cur] += coef(device) * w
__global__ void axpy_dev_k(int n, TriBuf tb, const Ctl* ctl, int which, const double* coef, const double* w) {
    double* y = tb.b[ctl->sc[which].cur];
    const double a = *coef;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] += a * w[i];
}
// x0 = sum_j c_j hist_j : polynomial extrapolation of the previous solutions of the same solve site,
// written into the current Chebyshev buffer (initial guess only - the accepted solution still has to meet tol)
struct PredictArgs { const double* h[6]; double c[6]; int m; };
__global__ void predict_k(int n, TriBuf tb, const Ctl* ctl, int which, PredictArgs pa) {
    double* x = tb.b[ctl->sc[which].cur];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double v = pa.c[0] * pa.h[0][i];
#pragma unroll
        for (int j = 1; j < 6; ++j)
            if (j < pa.m) v += pa.c[j] * pa.h[j][i];
        x[i] = v;
    }
}
__global__ void copy_from_tb_k(int n, double* dst, TriBuf tb, const Ctl* ctl, int which) {
    const double* xs = tb.b[ctl->sc[which].cur];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = xs[i];
}
__global__ void copy_to_tb_k(int n, TriBuf tb, const Ctl* ctl, int which, const double* src) {
    double* xs = tb.b[ctl->sc[which].cur];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) xs[i] = src ? src[i] : 0.0;
}

// port flux  b_L^T q = int_{Gamma_0} (q.n) g_L ds   (wave_2D.py:534,546,622-631), single CTA, fixed order
// one block per group (case of a batched run; a single group otherwise)
__global__ void flux_k(const int* __restrict__ port_goff, const int* __restrict__ pe, const double* __restrict__ bl, const double* q,
                       TriBuf tb, Ctl* ctl, int which, int slot) {
    __shared__ double red[64];
    const double* qs = q ? q : tb.b[ctl->sc[which].cur];
    const int g = blockIdx.x;
    double s = 0.0, z = 0.0;
    for (int i = port_goff[g] + threadIdx.x; i < port_goff[g + 1]; i += blockDim.x) s += bl[i] * qs[pe[i]];
    block_sum2(s, z, red);
    if (threadIdx.x == 0) ctl[g].flux[slot] = s;
}
// b[e] -= coef * b_L[e] * p_left   on the port edges (the ds(0) term of the q rows, wave_2D.py:469,480,504,515)
__global__ void port_rhs_k(int nport, const int* __restrict__ pe, const double* __restrict__ bl, const int* __restrict__ pg,
                           const double* __restrict__ gcoef, double* b, double coef, const Ctl* ctl) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nport) {
        const int g = pg[i];
        b[pe[i]] -= coef * (gcoef ? gcoef[g] : 1.0) * bl[i] * ctl[g].pleft;
    }
}
// batched runs: ctl[g].red[slot] = scale * gscale[g] * sum of the (patch, warp) partials of group g
__global__ void group_reduce_k(const double* __restrict__ wpart, int nw, const int* __restrict__ gpoff, const double* __restrict__ gscale,
                               double scale, Ctl* ctl, int slot) {
    const int g = blockIdx.x;
    double s = 0.0;
    for (int i = gpoff[g] * nw + threadIdx.x; i < gpoff[g + 1] * nw; i += 32) s += wpart[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) ctl[g].red[slot] = scale * (gscale ? gscale[g] : 1.0) * s;
}
// strong Dirichlet rows (wave_2D.py:413-422, bc.apply :593-595,817,852): the new p value is prescribed, so the
// velocity unknown of a constrained vertex is (w g(t) - p_old)/dt; its row is frozen by a zero Jacobi weight
__global__ void dirichlet_set_k(int n, const int* __restrict__ idx, const double* __restrict__ w, TriBuf tb, const Ctl* ctl,
                                int which, const double* p, double inv_dt) {
    double* x = tb.b[ctl->sc[which].cur];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[idx[i]] = (w[i] * ctl->pleft - p[idx[i]]) * inv_dt;
}
__global__ void zero_at_k(int n, const int* __restrict__ idx, double* v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[idx[i]] = 0.0;
}
__global__ void solve_begin_k(Ctl* ctl, int which) {
    SolveCtl& sc = ctl->sc[which];
    sc.k = 0; sc.done = 0; sc.racc = 0.0; sc.bacc = 0.0; sc.rho = 0.0; sc.solves += 1;
}

// ---- analytical diagnostics (wave_2D.py:929-965) -------------------------------------------------------------
struct AnalyticPrm {
    double amp, kx, ky, omega;          // exact p = amp sin(kx x + pi/2) sin(ky y + pi/2) cos(omega t)   (:741)
    double lam[15][3];                  // barycentric coordinates of the P4 (degree k+3) nodes
    int pt_v[3]; double pt_lam[3];      // point probe: internal vertex ids + barycentric weights (:946-950)
};
__constant__ double c_mass_ref[225];    // reference P4 mass matrix (|K| = 1)

// errornorm: || I_4 p_exact - p_h ||_L2^2 summed over the cells (both fields interpolated into P_{k+3}, :956)
__global__ void diag_cells_k(int nc, const int* __restrict__ cells_new, const double* __restrict__ vtx_new, const double* __restrict__ p,
                             const AnalyticPrm ap, const Ctl* ctl, double* partial) {
    __shared__ double red[64];
    const double ct = cos(ctl->t * ap.omega);
    double acc = 0.0, z = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
        const int a = cells_new[3 * c], b = cells_new[3 * c + 1], d = cells_new[3 * c + 2];
        const double ax = vtx_new[2 * a], ay = vtx_new[2 * a + 1], bx = vtx_new[2 * b], by = vtx_new[2 * b + 1], dx = vtx_new[2 * d], dy = vtx_new[2 * d + 1];
        const double pa = p[a], pb = p[b], pd = p[d];
        const double area = 0.5 * fabs((bx - ax) * (dy - ay) - (by - ay) * (dx - ax));
        double e[15];
#pragma unroll
        for (int n = 0; n < 15; ++n) {
            const double l0 = ap.lam[n][0], l1 = ap.lam[n][1], l2 = ap.lam[n][2];
            const double x = l0 * ax + l1 * bx + l2 * dx, y = l0 * ay + l1 * by + l2 * dy;
            e[n] = ap.amp * sin(ap.kx * x + 1.5707963267948966) * sin(ap.ky * y + 1.5707963267948966) * ct - (l0 * pa + l1 * pb + l2 * pd);
        }
        double qf = 0.0;
        for (int i = 0; i < 15; ++i) {
            double row = 0.0;
            for (int j = 0; j < 15; ++j) row += c_mass_ref[15 * i + j] * e[j];
            qf += e[i] * row;
        }
        acc += area * qf;
    }
    block_sum2(acc, z, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
// nodal max error (:958), point probe (:949-950) and the final sum of the cell partials
__global__ void diag_finish_k(int nv, int nblk_cells, const double* __restrict__ exact_v, const double* __restrict__ p, const AnalyticPrm ap,
                              Ctl* ctl, const double* partial) {
    __shared__ double red[64];
    const double ct = cos(ctl->t * ap.omega);
    double mx = 0.0;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) mx = fmax(mx, fabs(exact_v[i] * ct - p[i]));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m2 = 0.0;
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) m2 = fmax(m2, red[w]);
        double s = 0.0;
        for (int b = 0; b < nblk_cells; ++b) s += partial[b];
        ctl->red[4] = sqrt(s);
        ctl->red[5] = m2;
        double ph = 0.0, pe = 0.0;
        for (int k = 0; k < 3; ++k) if (ap.pt_v[k] >= 0) { ph += ap.pt_lam[k] * p[ap.pt_v[k]]; pe += ap.pt_lam[k] * exact_v[ap.pt_v[k]] * ct; }
        ctl->red[6] = ph; ctl->red[7] = pe;
    }
}

// higher-order analytical diagnostics: generic P_k -> P_{k+3} tabulation, data in global memory
__global__ void diag_cells_ho_k(int nc, int ndp, int nn, const int* __restrict__ cell_dofs, const double* __restrict__ T,
                                const double* __restrict__ mass, const double* __restrict__ exact_nodes,
                                const double* __restrict__ area, const double* __restrict__ p, double omega, const Ctl* ctl,
                                double* partial) {
    __shared__ double red[64];
    const double ct = cos(ctl->t * omega);
    double acc = 0.0, z = 0.0;
    double e[28], pl[10];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
        for (int j = 0; j < ndp; ++j) pl[j] = p[cell_dofs[(size_t)c * ndp + j]];
        for (int n = 0; n < nn; ++n) {
            double ph = 0.0;
            for (int j = 0; j < ndp; ++j) ph += T[n * ndp + j] * pl[j];
            e[n] = exact_nodes[(size_t)c * nn + n] * ct - ph;
        }
        double qf = 0.0;
        for (int i = 0; i < nn; ++i) {
            double row = 0.0;
            for (int j = 0; j < nn; ++j) row += mass[i * nn + j] * e[j];
            qf += e[i] * row;
        }
        acc += area[c] * qf;
    }
    block_sum2(acc, z, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void diag_finish_ho_k(int np, int nblk_cells, const double* __restrict__ exact_dofs, const double* __restrict__ p,
                                 double omega, int nprobe, const int* __restrict__ probe_dofs, const double* __restrict__ probe_w,
                                 Ctl* ctl, const double* partial) {
    __shared__ double red[64];
    const double ct = cos(ctl->t * omega);
    double mx = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) mx = fmax(mx, fabs(exact_dofs[i] * ct - p[i]));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m2 = 0.0;
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) m2 = fmax(m2, red[w]);
        double s = 0.0;
        for (int b = 0; b < nblk_cells; ++b) s += partial[b];
        ctl->red[4] = sqrt(s);
        ctl->red[5] = m2;
        double ph = 0.0, pe = 0.0;
        for (int k = 0; k < nprobe; ++k) { ph += probe_w[k] * p[probe_dofs[k]]; pe += probe_w[k] * exact_dofs[probe_dofs[k]] * ct; }
        ctl->red[6] = ph; ctl->red[7] = pe;
    }
}

struct ScalarPrm {
    double dt, K, rho, c2;
    double A[3][3];
    double m_em, L_em, K_em;
    int input_kind;
    double in_amp, in_omega, in_t_stop, in_t_ctrl, in_gain;
    double H_init;
    int ncol, ic;
    double w_old, w_new;   // energy-flux weights of b_L^T q_m and b_L^T q_new (wave_2D.py:603-615)
};

__device__ double input_value(const ScalarPrm& s, double t, double x0_prev) {
    switch (s.input_kind) {
        case PHW_IN_SINE_WINDOW: return (t < s.in_t_stop) ? s.in_amp * sin(s.in_omega * t) : 0.0;
        case PHW_IN_SINE: return s.in_amp * sin(s.in_omega * t);
        case PHW_IN_COSINE: return s.in_amp * cos(s.in_omega * t);
        case PHW_IN_PASSIVITY:
            if (t < s.in_t_stop) return s.in_amp * sin(s.in_omega * t);
            if (t < s.in_t_ctrl) return 0.0;
            return -s.in_gain * x0_prev;
    }
    return 0.0;
}

__device__ void solve3(double M[3][3], double r[3], double x[3]) {   // Gaussian elimination, partial pivoting
    int idx[3] = {0, 1, 2};
    for (int c = 0; c < 3; ++c) {
        int piv = c;
        for (int r_ = c + 1; r_ < 3; ++r_)
            if (fabs(M[idx[r_]][c]) > fabs(M[idx[piv]][c])) piv = r_;
        int t = idx[c]; idx[c] = idx[piv]; idx[piv] = t;
        for (int r_ = c + 1; r_ < 3; ++r_) {
            double f = M[idx[r_]][c] / M[idx[c]][c];
            for (int cc = c; cc < 3; ++cc) M[idx[r_]][cc] -= f * M[idx[c]][cc];
            r[idx[r_]] -= f * r[idx[c]];
        }
    }
    for (int c = 2; c >= 0; --c) {
        double s = r[idx[c]];
        for (int cc = c + 1; cc < 3; ++cc) s -= M[idx[c]][cc] * x[cc];
        x[c] = s / M[idx[c]][c];
    }
}

enum ScalarOp {
    S_BEGIN_1SUB = 0,   // t += dt; p_left = g(t) | u = u(t)          wave_2D.py:790-808
    S_SV_STAGE1,        // p_left(t_n) | lumped half step for x2         :812-820 with :530-534
    S_SV_MID,           // first boundary-energy half, t += dt, new input :830-843
    S_ODE_SE,           // lumped rows of symplectic Euler                :552-556
    S_ODE_SV,           // lumped rows of Stormer-Verlet stage 2          :538-546
    S_SM_PLEFT,         // p_left = rho x1_m / m for the explicit part of the SM q rows
    S_ODE_SM,           // bordered 3x3 solve of the implicit-midpoint lumped rows   :562-572
    S_END_1SUB,         // energies + H_array row for one-sub-step schemes :876-927,967-979
    S_END_SV,
    S_END_EH,
    S_EH_BEGIN
};

// one scalar operation of the step loop on the control block of case g (called by one thread)
__device__ void scalar_apply(Ctl& c, const ScalarPrm& s, int g, int op, double* Hrows, long long rows_per_group, double aux0) {
    const double dt = s.dt;
    switch (op) {
        case S_BEGIN_1SUB:
        case S_EH_BEGIN:
            c.t += dt;
            if (s.ic) { c.u_now = input_value(s, c.t, c.x[0]); }
            else c.pleft = input_value(s, c.t, 0.0);
            break;
        case S_SV_STAGE1:
            if (s.ic) {
                for (int i = 0; i < 3; ++i) c.xm[i] = c.x[i];
                // F_temp_em (wave_2D.py:530-534): only x[2] implicit; only x_temp[2] is used afterwards.
                // row 2: (x2 - x2m)/(dt/2) = A20 x0m + A21 x1m + A22 x2
                c.xt2 = (c.xm[2] + 0.5 * dt * (s.A[2][0] * c.xm[0] + s.A[2][1] * c.xm[1])) / (1.0 - 0.5 * dt * s.A[2][2]);
                c.pleft = s.rho * c.xm[1] / s.m_em;                                  // :392
            } else c.pleft = input_value(s, c.t, 0.0);
            break;
        case S_SV_MID:
            if (!s.ic) c.bE += 0.5 * dt * s.c2 * c.flux[0] * c.pleft;               // :622,830
            c.t += dt;                                                               // :833-834
            if (s.ic) c.u_now = input_value(s, c.t, c.xm[0]);
            else c.pleft = input_value(s, c.t, 0.0);
            break;
        case S_ODE_SE: {
            for (int i = 0; i < 3; ++i) c.xm[i] = c.x[i];
            double M[3][3], r[3], x[3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) M[i][j] = (i == j ? 1.0 : 0.0) - (j < 2 ? dt * s.A[i][j] : 0.0);
            const double f[3] = {c.u_now, s.K * c.flux[0], 0.0};
            for (int i = 0; i < 3; ++i) r[i] = c.xm[i] + dt * (s.A[i][2] * c.xm[2] + f[i]);
            solve3(M, r, x);
            for (int i = 0; i < 3; ++i) c.x[i] = x[i];
            c.pleft = s.rho * c.x[1] / s.m_em;                                       // :397
        } break;
        case S_ODE_SV: {
            double M[3][3], r[3], x[3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) M[i][j] = (i == j ? 1.0 : 0.0) - (j < 2 ? 0.5 * dt * s.A[i][j] : 0.0);
            r[0] = c.xm[0] + dt * (0.5 * (s.A[0][0] * c.xm[0] + s.A[0][1] * c.xm[1]) + s.A[0][2] * c.xt2 + c.u_now);
            r[1] = c.xm[1] + dt * (0.5 * (s.A[1][0] * c.xm[0] + s.A[1][1] * c.xm[1]) + s.A[1][2] * c.xt2 + s.K * c.flux[0]);
            r[2] = c.xt2 + dt * 0.5 * s.A[2][2] * c.xt2;
            solve3(M, r, x);
            for (int i = 0; i < 3; ++i) c.x[i] = x[i];
            c.pleft = s.rho * c.x[1] / s.m_em;                                       // :393
        } break;
        case S_SM_PLEFT:
            for (int i = 0; i < 3; ++i) c.xm[i] = c.x[i];
            c.pleft = s.rho * c.xm[1] / s.m_em;
            break;
        case S_ODE_SM: {
            // unknown xi = (x+ - x-)/dt ; v_q = vhat_q - 0.5 dt (xi1/m) w_q, aux0 = b_L^T w_q (constant), flux[2] = b_L^T vhat_q
            const double th = 0.5;
            double M[3][3], r[3], xi[3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) M[i][j] = (i == j ? 1.0 : 0.0) - th * dt * s.A[i][j];
            M[1][1] += th * th * dt * dt * s.K * aux0 / s.m_em;
            const double f[3] = {c.u_now, s.K * (c.flux[0] + th * dt * c.flux[2]), 0.0};
            for (int i = 0; i < 3; ++i) r[i] = s.A[i][0] * c.xm[0] + s.A[i][1] * c.xm[1] + s.A[i][2] * c.xm[2] + f[i];
            solve3(M, r, xi);
            for (int i = 0; i < 3; ++i) c.x[i] = c.xm[i] + dt * xi[i];
            c.xi1 = -th * dt * xi[1] / s.m_em;     // coefficient of w in v = vhat + xi1 * w
        } break;
        case S_END_1SUB:
        case S_END_SV:
        case S_END_EH: {
            double H_wave = c.red[0] + c.red[1];
            double E, H;
            if (!s.ic) {
                if (op == S_END_1SUB) {
                    c.bE += dt * s.c2 * (s.w_old * c.flux[0] + s.w_new * c.flux[1]) * c.pleft + dt * s.c2 * (c.red[2] + c.red[3]);   // :631-641
                } else if (op == S_END_SV) {
                    c.bE += 0.5 * dt * s.c2 * c.flux[0] * c.pleft;                   // :625,880
                } else {
                    // EH: q0_ = q_m, q1_ = q_temp = q_m + dt v_q1 ; flux[1] = b_L^T v_q1     :617-628
                    c.bE += 0.5 * dt * s.c2 * (2.0 * c.flux[0] + dt * c.flux[1]) * c.pleft;
                }
                E = H_wave + c.bE - s.H_init;                                        // :884
                H = H_wave;
            } else {
                if (op == S_END_SV) {
                    c.inpE += 0.5 * dt * c.u_now * c.xm[0] / s.L_em + 0.5 * dt * c.u_now * c.x[0] / s.L_em;          // :828,892,903
                    c.bE += 0.5 * dt * s.c2 * c.flux[0] * s.rho * (c.xm[1] + c.x[1]) / s.m_em;                        // :622-628
                } else if (s.w_new == 0.0) {   // SE
                    c.inpE += dt * c.u_now * c.x[0] / s.L_em;                                                         // :894
                    c.bE += dt * s.c2 * c.flux[0] * s.rho * c.x[1] / s.m_em;                                          // :631,398
                } else {                       // SM
                    c.inpE += 0.5 * dt * c.u_now * (c.xm[0] + c.x[0]) / s.L_em;                                      // :898
                    c.bE += dt * s.c2 * (s.w_old * c.flux[0] + s.w_new * c.flux[1]) * 0.5 * s.rho * (c.xm[1] + c.x[1]) / s.m_em;   // :631,401
                }
                c.H_em = 0.5 * (c.x[0] * c.x[0] / s.L_em + c.x[1] * c.x[1] / s.m_em + s.K_em * c.x[2] * c.x[2]);       // :911
                E = -c.inpE + H_wave + c.H_em - s.H_init;                                                              // :914
                H = H_wave + c.H_em;
                c.disp = c.x[2];                                                                                       // :921
            }
            c.H_wave = H_wave;
            if (Hrows) {
                double* row = Hrows + ((long long)g * rows_per_group + c.step) * s.ncol;   // :1011-1030
                row[0] = c.t; row[1] = H; row[2] = E; row[3] = c.disp; row[4] = c.bE; row[5] = c.inpE; row[6] = c.H_em; row[7] = H_wave;
                if (s.ncol == 12) { row[8] = c.red[4]; row[9] = c.red[5]; row[10] = c.red[6]; row[11] = c.red[7]; }    // :1027-1030
            }
            c.step += 1;
        } break;
    }
}

__global__ void scalar_k(Ctl* ctl, const ScalarPrm* __restrict__ prms, int n_groups, int op, double* Hrows, long long rows_per_group,
                         double aux0, double aux1) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const ScalarPrm s = prms[g];
    scalar_apply(ctl[g], s, g, op, Hrows, rows_per_group, aux0);
}

}  // namespace phw

#include "kernels_loop.cuh"
#include "kernels_ho.cuh"
#include "kernels_sweep.cuh"


// ================================================================================================
// solver object
// ================================================================================================
struct phw_solver_s {
    phw_mesh_s* mesh = nullptr;
    phw_params prm{};
    int device = 0;
    cudaStream_t stream = nullptr;
    int nv = 0, ne = 0, npatch = 0, npc = 0, nport = 0;
    std::vector<void*> allocs;
    DevMesh dm{};
    int *d_v_new2old = nullptr, *d_e_new2old = nullptr;
    int* d_port_e = nullptr; double* d_port_bl = nullptr;
    int n_dir = 0; int* d_dir_idx = nullptr; double* d_dir_w = nullptr;
    bool an_ready = false; AnalyticPrm an{}; int* d_cells_new = nullptr; double *d_vtx_new = nullptr, *d_exact_v = nullptr, *d_diag_part = nullptr;
    double *p = nullptr, *q = nullptr, *bp = nullptr, *bq = nullptr, *tmp_p = nullptr, *tmp_q = nullptr;
    double *pinv_p = nullptr, *pinv_q = nullptr, *mdiag_p = nullptr, *wp = nullptr, *wq = nullptr;
    double *io_v = nullptr, *io_e = nullptr, *io_v2 = nullptr, *io_e2 = nullptr;
    TriBuf tp{}, tq{};
    double* partials = nullptr;
    Ctl* ctl = nullptr;
    double* dH = nullptr; int64_t dH_rows = 0; int64_t run_rows = 0;
    ScalarPrm sp{};
    int G = 1; ScalarPrm* d_sp = nullptr; std::vector<ScalarPrm> h_sp;
    int *d_port_goff = nullptr, *d_port_group = nullptr, *d_gpoff = nullptr;
    double *d_gK = nullptr, *d_ginvrho = nullptr, *wpart = nullptr;
    double lam_min_q = 0.5, lam_max_q = 2.0, skew_bound = 0.0, blw = 0.0;
    double cheb_d[3] = {1.25, 1.25, 1.25}, cheb_c[3] = {0.75, 0.75, 0.75};
    size_t smem_v = 0, smem_v1 = 0, smem_e = 0, smem_vd = 0, smem_ed = 0;
    int grid_rows = 0, sm_count = 148;
    int64_t launches = 0;
    size_t l2_max_win = 0, l2_max_persist = 0;
    PeerComm pc{};
    void* arena = nullptr; size_t arena_bytes = 0;
    std::vector<void*> peer_arenas;
    struct Hist { double* h[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; int count = 0, head = 0; };
    Hist hist[2][2];           // [field p/q][solve site]
    int extrap = 2, n_sites = 1;
    int extrap_f[2] = {2, 2};  // extrapolation order of the p / q solves
    int bcsr = 0; double* bpart = nullptr; size_t bpart_cap = 0;   // batched sweep with the case index fastest (kernels_sweep.cuh): number of cases, [row block][case] partial sums
    bool pcg[2] = {false, false};     // conjugate-gradient mass solves (poor meshes; flags bits 16 / 17)
    bool implicit_pending = false;    // partitioned implicit run: the ellipse / bordered column are set up by phw_comm_finish_implicit
    bool ho = false; Csr hMp{}, hMq{}, hD{}, hDT{};
    bool ho_mf = false; HoMf hm{}; RowCsr hN{}, hNT{}; double *ho_yacc_p = nullptr, *ho_yacc_q = nullptr;   // matrix-free higher-order pairs (kernels_ho.cuh)
    int ho_nc = 0, ho_ndp = 0, ho_nn = 0, ho_nprobe = 0; double ho_omega = 0.0;
    int *ho_cell_dofs = nullptr, *ho_probe_dofs = nullptr;
    double *ho_T = nullptr, *ho_mass = nullptr, *ho_exact_nodes = nullptr, *ho_area = nullptr, *ho_exact_dofs = nullptr, *ho_probe_w = nullptr;
    bool in_step = false;
    bool coop = false;
    bool ell_p = false;         // M_p solve on the assembled ELL rows (vertex_ell_pass)
    bool ell_staged = false;    // ... with all streams of a patch staged in shared memory by cp.async
    size_t smem_ell = 0, smem_ells = 0;
    uint4* d_cellq = nullptr;
    bool pipe_q = false, pipe_p = false;   // bulk-copy pipelined mass solves (kernels_pipe.cuh; flags bits 8 / 9)
    MqPipeArgs pq{}; EllPipeArgs pp{};     // their shared-memory carve-ups (sized for the largest patch)
    size_t smem_pq = 0, smem_pp = 0;
    bool pipe_en = false; EnergyPipeArgs pe{}; size_t smem_pe = 0;   // fused pipelined H_wave reduction (flags bit 11)
    bool pipe_rhs = false; RhsPipeArgs pre{}, prv{}; size_t smem_pre = 0, smem_prv = 0;   // pipelined right-hand sides (flags bit 12)
    bool loop_on = false; int loop_cl = 0, loop_nt = 512; bool loop_cluster = false; double* loop_partials = nullptr;
    // field snapshots (phw_set_snapshots): ring of device slots in the user numbering, drained by a second stream
    static constexpr int SNAP_SLOTS = 4;
    int64_t snap_every = 0, snap_cap = 0, snap_taken = 0; double* snap_dst = nullptr; double* snap_ring = nullptr; void* snap_registered = nullptr;
    cudaStream_t snap_stream = nullptr; cudaEvent_t snap_ready[SNAP_SLOTS] = {}, snap_done[SNAP_SLOTS] = {};
    bool loop_res = false; ResCarve loop_rc{};   // mass solves of the whole-loop kernel resident in shared memory (kernels_loop.cuh, RES)   // whole-loop kernel (flags bit 15)
    int64_t kernel_paths = 0;              // PHW_PATH_* bits of the solve kernels launched since the last phw_run began
    int *d_cells_user = nullptr; double* d_vtx_user = nullptr; unsigned long long* d_mm = nullptr;
    std::vector<cudaEvent_t> prof_ev; std::vector<int> prof_which; bool prof_truncated = false;
    bool have_vq = false;      // Stormer-Verlet first-same-as-last: v_q of the previous step is valid
    phw_stats stats{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int ncol = 8;

    template <class T> int dalloc(T** ptr, size_t n) {
        void* d = nullptr;
        cudaError_t e = cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T));
        if (e != cudaSuccess) return fail(PHW_ENOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
        allocs.push_back(d);
        *ptr = (T*)d;
        return 0;
    }
    template <class T> int upload(T** ptr, const std::vector<T>& h) {
        int rc = dalloc(ptr, h.size());
        if (rc) return rc;
        if (!h.empty()) CU(cudaMemcpyAsync(*ptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
        return 0;
    }
};

namespace {

inline int nblk(int n, int bs = 256) { return (n + bs - 1) / bs; }

// opt-in to more than 48 KB of dynamic shared memory.  The attribute belongs to the (device, function) pair, so it is tracked per
// device: a second solver on another device of the same process configures its own copy of every kernel.
int ensure_dyn_smem(int device, const void* kern, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<int, const void*>> done;
    std::lock_guard<std::mutex> lk(mu);
    if (done.count(std::make_pair(device, kern))) return 0;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.insert(std::make_pair(device, kern));
    return 0;
}
inline int vec_grid(phw_solver_s* s, int n) { return std::max(1, std::min(nblk(n), s->sm_count * 8)); }

// ---- typed launchers for the row kernels -------------------------------------------------------
template <int EPI, int MODE, bool HAS_M, bool HAS_D, bool CAS>
int launch_vertex(phw_solver_s* s, const RowArgs& a) {
    size_t sm = HAS_D ? s->smem_vd : ((HAS_M && !CAS && MODE == MODE_OP) ? s->smem_v1 : s->smem_v);
    auto kern = vertex_rows_k<EPI, MODE, HAS_M, HAS_D, CAS>;
    { int rc_ = ensure_dyn_smem(s->device, (const void*)kern, 200 * 1024); if (rc_) return rc_; }
    kern<<<s->grid_rows, NT_ROWS, sm, s->stream>>>(a);
    s->launches++;
    CU(cudaGetLastError());
    return 0;
}
template <int EPI, int MODE, bool HAS_M, bool HAS_D, bool CAS>
int launch_edge(phw_solver_s* s, const RowArgs& a) {
    size_t sm = HAS_D ? s->smem_ed : s->smem_e;
    auto kern = edge_rows_k<EPI, MODE, HAS_M, HAS_D, CAS>;
    { int rc_ = ensure_dyn_smem(s->device, (const void*)kern, 200 * 1024); if (rc_) return rc_; }
    kern<<<s->grid_rows, NT_ROWS, sm, s->stream>>>(a);
    s->launches++;
    CU(cudaGetLastError());
    return 0;
}

// ---- CSR operators of the higher-order pairs: a group of `lanes` threads per row
inline int csr_grid(phw_solver_s* s, const Csr& A) { return std::max(1, std::min(nblk(A.n_rows * A.lanes), s->sm_count * 8)); }
void csr_store(phw_solver_s* s, const Csr& A, const double* x, double alpha, double* y) {
    const int g = csr_grid(s, A);
    switch (A.lanes) {
        case 4: csr_store_k<4><<<g, 256, 0, s->stream>>>(A, x, alpha, y); break;
        case 8: csr_store_k<8><<<g, 256, 0, s->stream>>>(A, x, alpha, y); break;
        case 16: csr_store_k<16><<<g, 256, 0, s->stream>>>(A, x, alpha, y); break;
        default: csr_store_k<32><<<g, 256, 0, s->stream>>>(A, x, alpha, y);
    }
}
void csr_dot(phw_solver_s* s, const Csr& A, const double* x, double scale, int slot, int counter) {
    const int g = csr_grid(s, A);
    double* part = s->partials + 8192 + 2048 * (size_t)slot;     // beyond the partials of the solves
    switch (A.lanes) {
        case 4: csr_dot_k<4><<<g, 256, 0, s->stream>>>(A, x, scale, s->ctl, slot, part, counter); break;
        case 8: csr_dot_k<8><<<g, 256, 0, s->stream>>>(A, x, scale, s->ctl, slot, part, counter); break;
        case 16: csr_dot_k<16><<<g, 256, 0, s->stream>>>(A, x, scale, s->ctl, slot, part, counter); break;
        default: csr_dot_k<32><<<g, 256, 0, s->stream>>>(A, x, scale, s->ctl, slot, part, counter);
    }
}

// ---- batched sweep, case index fastest (kernels_sweep.cuh): one warp = 32 cases x a block of rows
inline bool prof_on(phw_solver_s* s);
#ifndef SWEEP_CPL_MAX
#define SWEEP_CPL_MAX 4            // default cap of the cases per lane (PHW_SWEEP_CPL = 1, 2: A/B)
#endif
// row mapping of the sweep kernels (kernels_sweep.cuh: SweepMap): 0 = one block of rows per warp (default), 1 = cyclic row tiles
inline int sweep_map_mode() {
    static const int m = [] { const char* e = getenv("PHW_SWEEP_MAP"); return e ? atoi(e) : 0; }();
    return m == 1 ? 1 : 0;
}
// cases per lane of the sweep kernels (kernels_sweep.cuh): 4 / 2 / 1 by the divisibility of the number of cases; PHW_SWEEP_CPL caps it (A/B)
inline int sweep_cpl(phw_solver_s* s) {
    static const int cap = [] { const char* e = getenv("PHW_SWEEP_CPL"); const int v = e ? atoi(e) : 0; return (v == 1 || v == 2 || v == 4) ? v : SWEEP_CPL_MAX; }();
    int cpl = (s->bcsr % 4 == 0) ? 4 : ((s->bcsr % 2 == 0) ? 2 : 1);
    return std::min(cpl, cap);
}
inline int bcsr_grid(phw_solver_s* s, const Csr& A, int cpl, int* n_row_groups) {
    const int nchunk = (s->bcsr + 32 * cpl - 1) / (32 * cpl);
    const int groups = std::max(1, std::min((A.n_rows + SWEEP_WPB - 1) / SWEEP_WPB, (8 * s->sm_count) / nchunk));   // ~8 CTAs per SM
    *n_row_groups = groups;
    return nchunk * groups;
}
void bcsr_store(phw_solver_s* s, const Csr& A, const double* x, double alpha, const double* cscale, double* y) {
    int ng = 1;
    const int cpl = sweep_cpl(s), g = bcsr_grid(s, A, cpl, &ng), m = sweep_map_mode();
    if (cpl == 4) csr_store_b_k<4><<<g, 256, 0, s->stream>>>(A, s->bcsr, ng, m, x, alpha, cscale, y);
    else if (cpl == 2) csr_store_b_k<2><<<g, 256, 0, s->stream>>>(A, s->bcsr, ng, m, x, alpha, cscale, y);
    else csr_store_b_k<1><<<g, 256, 0, s->stream>>>(A, s->bcsr, ng, m, x, alpha, cscale, y);
    s->launches++;
}
int bcsr_dot(phw_solver_s* s, const Csr& A, const double* x, double scale, const double* cscale, int slot) {
    int ng = 1;
    const int cpl = sweep_cpl(s), g = bcsr_grid(s, A, cpl, &ng), m = sweep_map_mode();
    if ((size_t)ng * SWEEP_WPB * (size_t)s->bcsr > s->bpart_cap) return fail(PHW_ECUDA, "sweep: partial-sum buffer too small");
    if (cpl == 4) csr_dot_b_k<4><<<g, 256, 0, s->stream>>>(A, s->bcsr, ng, m, x, s->bpart);
    else if (cpl == 2) csr_dot_b_k<2><<<g, 256, 0, s->stream>>>(A, s->bcsr, ng, m, x, s->bpart);
    else csr_dot_b_k<1><<<g, 256, 0, s->stream>>>(A, s->bcsr, ng, m, x, s->bpart);
    dot_b_finish_k<<<nblk(s->bcsr, 128), 128, 0, s->stream>>>(s->bcsr, ng * SWEEP_WPB, s->bpart, scale, cscale, s->ctl, slot);
    s->launches += 2;
    return 0;
}
int bcsr_solve(phw_solver_s* s, int which) {
    CsrSolveBArgs a{};
    a.A = which == PHW_P ? s->hMp : s->hMq; a.B = s->bcsr; a.tb = which == PHW_P ? s->tp : s->tq;
    a.b = which == PHW_P ? s->bp : s->bq; a.pinv = which == PHW_P ? s->pinv_p : s->pinv_q;
    a.cheb_d = s->cheb_d[which]; a.cheb_c = s->cheb_c[which];
    // the stop test sees the residual aggregated over all cases: tighten it so that every case of comparable amplitude meets tol
    const double t = std::max(s->prm.tol / std::sqrt((double)s->bcsr), 1e-14);
    a.tol2 = std::min(s->prm.tol * s->prm.tol, t * t);
    a.max_iter = s->prm.max_iter; a.which = which; a.partials = s->partials; a.ctl = s->ctl;
    CU(cudaMemsetAsync(&s->ctl->gbar[0], 0, sizeof(unsigned int), s->stream));
    const int cpl = sweep_cpl(s);
    void* kern = cpl == 4 ? (void*)solve_coop_csr_b_k<4> : (cpl == 2 ? (void*)solve_coop_csr_b_k<2> : (void*)solve_coop_csr_b_k<1>);
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
    if (occ < 1) return fail(PHW_ECUDA, "solve_coop_csr_b_k does not fit on an SM");
    const int nchunk = (s->bcsr + 32 * cpl - 1) / (32 * cpl);
    const int max_ctas = std::min(occ, 8) * s->sm_count;
    if (nchunk > max_ctas) return fail(PHW_EINVAL, "too many cases for one sweep solver: shard the case list");
    a.n_row_groups = sweep_row_groups(max_ctas, s->bcsr, a.A.n_rows, cpl);
    a.mode = sweep_map_mode();
    const int grid = nchunk * a.n_row_groups;
    if (2 * 2 * grid > 8192) return fail(PHW_ECUDA, "partials too small");
    void* args[] = {(void*)&a};
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool prof = prof_on(s);
    if (prof) {
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, s->stream));
    }
    CU(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(256), args, 0, s->stream));
    if (prof) {
        CU(cudaEventRecord(e1, s->stream));
        s->prof_ev.push_back(e0); s->prof_ev.push_back(e1); s->prof_which.push_back(which);
    }
    s->launches++;
    s->kernel_paths |= PHW_PATH_SWEEP_CSR;
    return 0;
}

// ---- matrix-free higher-order operators (kernels_ho.cuh): cell loops with the reference tabulations in shared memory
inline size_t ho_smem(const HoMf& m) { return (size_t)(((m.ref_len + 1) & ~1) + HO_NT) * sizeof(double); }
inline int ho_grid(phw_solver_s* s) { return std::max(1, std::min(nblk(s->hm.nc, HO_NT / HO_LPC), s->sm_count * 8)); }
// y (+)= alpha A x, A = M_p, M_q, C or C^T
int ho_mf_apply(phw_solver_s* s, int op, const double* x, double* y, double alpha, bool zero_first) {
    const HoMf& m = s->hm;
    const int n_out = (op == HO_MP || op == HO_CT) ? m.n_p : m.n_q;
    if (zero_first) CU(cudaMemsetAsync(y, 0, (size_t)n_out * sizeof(double), s->stream));
    const int g = ho_grid(s); const size_t sm = ho_smem(m);
    switch (op) {
        case HO_MP: ho_apply_k<HO_MP><<<g, HO_NT, sm, s->stream>>>(m, x, y, alpha); break;
        case HO_MQ: ho_apply_k<HO_MQ><<<g, HO_NT, sm, s->stream>>>(m, x, y, alpha); break;
        case HO_C: ho_apply_k<HO_C><<<g, HO_NT, sm, s->stream>>>(m, x, y, alpha); break;
        default: ho_apply_k<HO_CT><<<g, HO_NT, sm, s->stream>>>(m, x, y, alpha);
    }
    s->launches++;
    CU(cudaGetLastError());
    return 0;
}
int ho_row_csr(phw_solver_s* s, const RowCsr& A, const double* x, double alpha, double* y) {
    if (A.n_rows == 0) return 0;
    row_csr_axpy_k<<<nblk(A.n_rows), 256, 0, s->stream>>>(A, x, alpha, y);
    s->launches++;
    CU(cudaGetLastError());
    return 0;
}
// y = alpha D x = alpha (C - N) x   /   y = alpha D^T x
int ho_mf_D(phw_solver_s* s, const double* p, double alpha, double* y) {
    int rc = ho_mf_apply(s, HO_C, p, y, alpha, true); if (rc) return rc;
    return ho_row_csr(s, s->hN, p, -alpha, y);
}
int ho_mf_DT(phw_solver_s* s, const double* q, double alpha, double* y) {
    int rc = ho_mf_apply(s, HO_CT, q, y, alpha, true); if (rc) return rc;
    return ho_row_csr(s, s->hNT, q, -alpha, y);
}
int ho_mf_quad(phw_solver_s* s, int op, const double* x, double scale, int slot, int counter) {
    const int g = ho_grid(s);
    double* part = s->partials + 8192 + 2048 * (size_t)slot;
    if (op == HO_MP) ho_quad_k<HO_MP><<<g, HO_NT, ho_smem(s->hm), s->stream>>>(s->hm, x, scale, s->ctl, slot, part, counter);
    else ho_quad_k<HO_MQ><<<g, HO_NT, ho_smem(s->hm), s->stream>>>(s->hm, x, scale, s->ctl, slot, part, counter);
    s->launches++;
    CU(cudaGetLastError());
    return 0;
}
int ho_mf_solve(phw_solver_s* s, int which) {
    HoSolveArgs a{};
    a.m = s->hm; a.tb = which == PHW_P ? s->tp : s->tq;
    a.b = which == PHW_P ? s->bp : s->bq; a.pinv = which == PHW_P ? s->pinv_p : s->pinv_q;
    a.yacc = which == PHW_P ? s->ho_yacc_p : s->ho_yacc_q;
    a.cheb_d = s->cheb_d[which]; a.cheb_c = s->cheb_c[which]; a.tol2 = s->prm.tol * s->prm.tol;
    a.max_iter = s->prm.max_iter; a.which = which; a.n = which == PHW_P ? s->hm.n_p : s->hm.n_q;
    a.partials = s->partials; a.ctl = s->ctl;
    CU(cudaMemsetAsync(&s->ctl->gbar[0], 0, sizeof(unsigned int), s->stream));
    // CTA size: an iteration crosses two grid barriers whose cost grows with the number of CTAs (one atomic arrival each), so the
    // cells are spread over as few, as large CTAs as cover them in one pass (PHW_HO_NT: A/B knob)
    static const int nt_env = [] { const char* e = getenv("PHW_HO_NT"); const int v = e ? atoi(e) : 0; return (v == 256 || v == 512 || v == 1024) ? v : 0; }();
    int nt = nt_env;
    if (!nt) {
        const int need = std::max(s->hm.nc * HO_LPC, a.n);
        nt = need <= 256 * s->sm_count ? 256 : (need <= 512 * s->sm_count ? 512 : 1024);
    }
    int occ = 0;
    const size_t sm = (size_t)(((s->hm.ref_len + 1) & ~1) + nt) * sizeof(double);
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (void*)ho_solve_k, nt, sm));
    if (occ < 1) return fail(PHW_ECUDA, "ho_solve_k does not fit on an SM");
    const int grid = std::max(1, std::min(std::max(nblk(s->hm.nc, nt / HO_LPC), nblk(a.n, nt)), std::min(occ, 8) * s->sm_count));
    void* args[] = {(void*)&a};
    CU(cudaLaunchCooperativeKernel((void*)ho_solve_k, dim3(grid), dim3(nt), args, sm, s->stream));
    s->launches++;
    s->kernel_paths |= PHW_PATH_HO_MF;
    return 0;
}

// partitioned runs: which pipelined kernels (when selected by phw_params.flags) run with the inter-GPU exchange.  A/B and fallback
// knob PHW_PIPE_MULTI = bit mask (1: the two mass solves, 2: right-hand sides, 4: fused H_wave pass); default 7 = all of them,
// 0 = the lean kernels of round 1.
int pipe_multi_mask() {
    static const int m = [] { const char* e = getenv("PHW_PIPE_MULTI"); return e ? atoi(e) : 7; }();
    return m;
}

RowArgs base_args(phw_solver_s* s) {
    RowArgs a{};
    a.m = s->dm; a.tp = s->tp; a.tq = s->tq;
    a.partials = s->partials; a.ctl = s->ctl;
    a.tol2 = s->prm.tol * s->prm.tol; a.max_iter = s->prm.max_iter; a.weight = 1.0;
    // batched runs test the aggregated residual of all cases: tighten it so that every case of comparable amplitude meets tol
    if (s->G > 1) { const double t = std::max(s->prm.tol / std::sqrt((double)s->G), 1e-14); a.tol2 = std::min(a.tol2, t * t); }
    a.dot_scale = 1.0;
    a.mdiag = s->mdiag_p;   // vertex M-only rows use the pure P1 mass diagonal
    a.xscale = s->prm.strong ? 4.0 : 0.0;
    a.gbar = (s->prm.flags & 1024) ? 1 : 0;
    return a;
}

// bulk-copy pipelined right-hand sides (kernels_pipe.cuh): y_e = cD (D x_p)_e / y_v = cD (D^T x_q)_v
template <bool VERTEX, int NT, int MINB>
int launch_rhs_pipe_cfg(phw_solver_s* s, const double* x, double cD, double* y) {
    auto kern = VERTEX ? rhs_vertex_pipe_k<NT, MINB> : rhs_edge_pipe_k<NT, MINB>;
    const size_t sm = VERTEX ? s->smem_prv : s->smem_pre;
    { int rc_ = ensure_dyn_smem(s->device, (const void*)kern, 227 * 1024 - 1024); if (rc_) return rc_; }
    RhsPipeArgs a = VERTEX ? s->prv : s->pre;
    a.npatch = s->npatch; a.hdr = s->dm.hdr; a.recv = s->dm.recv; a.rece = s->dm.rece;
    a.ghost = VERTEX ? s->dm.ghost_e : s->dm.ghost_v; a.eadj = s->dm.eadj; a.vslice_off = s->dm.vslice_off; a.vadj = s->dm.vadj;
    a.x = x; a.y = y; a.cD = cD; a.ctl = s->ctl;
    kern<<<std::max(1, std::min(s->npatch, MINB * s->sm_count)), NT, sm, s->stream>>>(a);
    s->launches++;
    s->kernel_paths |= PHW_PATH_PIPE_RHS;
    CU(cudaGetLastError());
    return 0;
}
template <bool VERTEX>
int launch_rhs_pipe(phw_solver_s* s, const double* x, double cD, double* y) {
    const size_t sm = VERTEX ? s->smem_prv : s->smem_pre;
    if (3 * (sm + 2048) <= 227 * 1024) return launch_rhs_pipe_cfg<VERTEX, 384, 3>(s, x, cD, y);
    if (2 * (sm + 2048) <= 227 * 1024) return launch_rhs_pipe_cfg<VERTEX, 512, 2>(s, x, cD, y);
    return launch_rhs_pipe_cfg<VERTEX, 1024, 1>(s, x, cD, y);
}

// y_v = cD (D^T xq)_v [+ cB (B_top xp)_v]          (explicit part of the p rows)
int op_vertex_store(phw_solver_s* s, const double* xp, const double* xq, double cM, double cD, double cB, double* y, const double* gscale = nullptr) {
    RowArgs a = base_args(s);
    a.gscale = gscale;
    a.xp = xp; a.xq = xq; a.cM = cM; a.cD = cD; a.cB = cB; a.y = y;
    const bool cas = s->prm.casimir && cB != 0.0;
    if (s->pipe_rhs && (s->pc.world <= 1 || (pipe_multi_mask() & 2)) && cM == 0.0 && cB == 0.0 && !gscale) return launch_rhs_pipe<true>(s, xq, cD, y);
    if (cM != 0.0 && cD == 0.0) return cas ? launch_vertex<EPI_STORE, MODE_OP, true, false, true>(s, a) : launch_vertex<EPI_STORE, MODE_OP, true, false, false>(s, a);
    if (cM == 0.0) return cas ? launch_vertex<EPI_STORE, MODE_OP, false, true, true>(s, a) : launch_vertex<EPI_STORE, MODE_OP, false, true, false>(s, a);
    return cas ? launch_vertex<EPI_STORE, MODE_OP, true, true, true>(s, a) : launch_vertex<EPI_STORE, MODE_OP, true, true, false>(s, a);
}
int op_edge_store(phw_solver_s* s, const double* xp, const double* xq, double cM, double cD, double cB, double* y, const double* gscale = nullptr) {
    RowArgs a = base_args(s);
    a.gscale = gscale;
    a.xp = xp; a.xq = xq; a.cM = cM; a.cD = cD; a.cB = cB; a.y = y;
    const bool cas = s->prm.casimir && cB != 0.0;
    if (s->pipe_rhs && (s->pc.world <= 1 || (pipe_multi_mask() & 2)) && cM == 0.0 && cB == 0.0 && !gscale) return launch_rhs_pipe<false>(s, xp, cD, y);
    if (cM != 0.0 && cD == 0.0) return cas ? launch_edge<EPI_STORE, MODE_OP, true, false, true>(s, a) : launch_edge<EPI_STORE, MODE_OP, true, false, false>(s, a);
    if (cM == 0.0) return cas ? launch_edge<EPI_STORE, MODE_OP, false, true, true>(s, a) : launch_edge<EPI_STORE, MODE_OP, false, true, false>(s, a);
    return cas ? launch_edge<EPI_STORE, MODE_OP, true, true, true>(s, a) : launch_edge<EPI_STORE, MODE_OP, true, true, false>(s, a);
}
// ctl->red[slot] (+)= scale * x^T (cM M + cB B) x
int op_vertex_dot(phw_solver_s* s, const double* xp, double cM, double cB, int slot, double scale, int counter) {
    RowArgs a = base_args(s);
    a.xp = xp; a.cM = cM; a.cB = cB; a.red_slot = slot; a.dot_scale = scale; a.counter_id = counter;
    if (s->G > 1) a.wpart = s->wpart;
    if (cB != 0.0) return (cM != 0.0) ? launch_vertex<EPI_DOT, MODE_OP, true, false, true>(s, a) : launch_vertex<EPI_DOT, MODE_OP, false, false, true>(s, a);
    return launch_vertex<EPI_DOT, MODE_OP, true, false, false>(s, a);
}
int op_edge_dot(phw_solver_s* s, const double* xq, double cM, double cB, int slot, double scale, int counter) {
    RowArgs a = base_args(s);
    a.xq = xq; a.cM = cM; a.cB = cB; a.red_slot = slot; a.dot_scale = scale; a.counter_id = counter;
    if (s->G > 1) a.wpart = s->wpart;
    if (cB != 0.0) return (cM != 0.0) ? launch_edge<EPI_DOT, MODE_OP, true, false, true>(s, a) : launch_edge<EPI_DOT, MODE_OP, false, false, true>(s, a);
    return launch_edge<EPI_DOT, MODE_OP, true, false, false>(s, a);
}

// CUDA-event timing of every solve launch (flags bit 2).  Beyond 2^18 launches per phw_run the timing stops and the per-kernel
// figures of that run are reported as 0 (a partial sum against the full iteration count would overstate the bandwidth).
inline bool prof_on(phw_solver_s* s) {
    if (!(s->prm.flags & 4)) return false;
    if (s->prof_ev.size() >= 2 * (size_t)262144) { s->prof_truncated = true; return false; }
    return true;
}

// one cooperative launch = one complete Chebyshev solve
template <int WHICH, bool CAS, int ELL = 0>
int launch_coop(phw_solver_s* s, const RowArgs& av, const RowArgs& ae) {
    constexpr int NT = (WHICH == 0) ? (ELL == 2 ? PHW_NT_ELLS : (ELL == 1 ? PHW_NT_ELL : NT_COOP_P)) : (WHICH == 1 ? NT_COOP_Q : NT_ROWS);
    constexpr int MINB = (WHICH == 0) ? (ELL == 2 ? PHW_MINB_ELLS : (ELL == 1 ? PHW_MINB_ELL : PHW_MINB_P)) : PHW_MINB_Q;
    auto kern = solve_coop_k<WHICH, CAS, NT, MINB, ELL>;
    const size_t sm = (WHICH == 0) ? (ELL == 2 ? s->smem_ells : (ELL == 1 ? s->smem_ell : s->smem_v1)) : (WHICH == 1 ? s->smem_e : s->smem_vd);
    { int rc_ = ensure_dyn_smem(s->device, (const void*)kern, 200 * 1024); if (rc_) return rc_; }
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, sm));
    if (occ < 1) return fail(PHW_ECUDA, "cooperative solve kernel does not fit on an SM");
    const int grid = std::max(1, std::min(s->npatch, occ * s->sm_count));
    void* args[] = {(void*)&av, (void*)&ae, (void*)&s->pc};
    if (av.gbar) CU(cudaMemsetAsync(&s->ctl->gbar[0], 0, sizeof(unsigned int), s->stream));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool prof = prof_on(s);
    if (prof) {
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, s->stream));
    }
    CU(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(NT), args, sm, s->stream));
    if (prof) {
        CU(cudaEventRecord(e1, s->stream));
        s->prof_ev.push_back(e0); s->prof_ev.push_back(e1); s->prof_which.push_back(WHICH);
    }
    s->launches++;
    return 0;
}

// conjugate-gradient mass solve (solve_pcg_coop_k): residual / preconditioned residual in the import staging vectors, search
// direction and A pd in the two idle iterate buffers
template <int WHICH>
int launch_pcg(phw_solver_s* s, const RowArgs& a0) {
    constexpr int NT = WHICH == 0 ? NT_COOP_P : NT_COOP_Q;
    constexpr int MINB = WHICH == 0 ? PHW_MINB_P : PHW_MINB_Q;
    auto kern = solve_pcg_coop_k<WHICH, NT, MINB>;
    const size_t sm = WHICH == 0 ? s->smem_v1 : s->smem_e;
    { int rc_ = ensure_dyn_smem(s->device, (const void*)kern, 200 * 1024); if (rc_) return rc_; }
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, sm));
    if (occ < 1) return fail(PHW_ECUDA, "conjugate-gradient solve kernel does not fit on an SM");
    const int grid = std::max(1, std::min(s->npatch, occ * s->sm_count));
    RowArgs a = a0;
    if (WHICH == 0) a.mdiag = s->mdiag_p;      // the STORE epilogue of the pure P1 mass rows reads the diagonal (solve_mass clears it for Chebyshev)
    PcgArgs v{};
    v.r = WHICH == 0 ? s->io_v : s->io_e; v.u = WHICH == 0 ? s->io_v2 : s->io_e2;
    v.b = a.b; v.pinv = a.pinv; v.n = WHICH == 0 ? s->nv : s->ne;
    void* args[] = {(void*)&a, (void*)&v};
    CU(cudaMemsetAsync(&s->ctl->gbar[0], 0, sizeof(unsigned int), s->stream));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool prof = prof_on(s);
    if (prof) {
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, s->stream));
    }
    CU(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(NT), args, sm, s->stream));
    if (prof) {
        CU(cudaEventRecord(e1, s->stream));
        s->prof_ev.push_back(e0); s->prof_ev.push_back(e1); s->prof_which.push_back(WHICH);
    }
    s->launches++;
    s->kernel_paths |= PHW_PATH_PCG;
    return 0;
}

// lean single-GPU ELL M_p solve
template <int NT, int MINB>
int launch_ell_lean(phw_solver_s* s) {
    void* kern = s->pc.world > 1 ? (void*)solve_ell_coop_k<NT, MINB, true> : (void*)solve_ell_coop_k<NT, MINB, false>;
    CU(cudaMemsetAsync(&s->ctl->gbar[0], 0, sizeof(unsigned int), s->stream));
    { int rc_ = ensure_dyn_smem(s->device, (const void*)kern, 200 * 1024); if (rc_) return rc_; }
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, s->smem_ell));
    if (occ < 1) return fail(PHW_ECUDA, "ELL solve kernel does not fit on an SM");
    const int grid = std::max(1, std::min(s->npatch, occ * s->sm_count));
    EllSolveArgs a{};
    a.npatch = s->npatch; a.hdr = s->dm.hdr; a.ell_ioff = s->dm.ell_ioff; a.ell_woff = s->dm.ell_woff; a.ell_ids = s->dm.ell_ids; a.ell_w = s->dm.ell_w;
    a.ghost_v = s->dm.ghost_v; a.b = s->bp;
    for (int i = 0; i < 3; ++i) a.buf[i] = s->tp.b[i];
    a.cheb_d = s->cheb_d[0]; a.cheb_c = s->cheb_c[0];
    a.tol2 = s->prm.tol * s->prm.tol;
    if (s->G > 1) { const double t = std::max(s->prm.tol / std::sqrt((double)s->G), 1e-14); a.tol2 = std::min(a.tol2, t * t); }
    a.max_iter = s->prm.max_iter; a.which = 0; a.partials = s->partials; a.ctl = s->ctl;
    void* args[] = {(void*)&a, (void*)&s->pc};
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool prof = prof_on(s);
    if (prof) {
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, s->stream));
    }
    CU(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(NT), args, s->smem_ell, s->stream));
    if (prof) {
        CU(cudaEventRecord(e1, s->stream));
        s->prof_ev.push_back(e0); s->prof_ev.push_back(e1); s->prof_which.push_back(0);
    }
    s->launches++;
    return 0;
}

// lean single-GPU cell-based M_q solve
template <int NT, int MINB, int UC, int UR>
int launch_mq_lean(phw_solver_s* s) {
    void* kern = s->pc.world > 1 ? (void*)solve_mq_coop_k<NT, MINB, UC, UR, true> : (void*)solve_mq_coop_k<NT, MINB, UC, UR, false>;
    CU(cudaMemsetAsync(&s->ctl->gbar[0], 0, sizeof(unsigned int), s->stream));
    { int rc_ = ensure_dyn_smem(s->device, (const void*)kern, 200 * 1024); if (rc_) return rc_; }
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, s->smem_e));
    if (occ < 1) return fail(PHW_ECUDA, "M_q solve kernel does not fit on an SM");
    const int grid = std::max(1, std::min(s->npatch, occ * s->sm_count));
    MqSolveArgs a{};
    a.npatch = s->npatch; a.hdr = s->dm.hdr; a.cellq = s->d_cellq;
    a.ghost_e = s->dm.ghost_e; a.eadj = s->dm.eadj; a.b = s->bq; a.pinv = s->pinv_q;
    for (int i = 0; i < 3; ++i) a.buf[i] = s->tq.b[i];
    a.cheb_d = s->cheb_d[1]; a.cheb_c = s->cheb_c[1];
    a.tol2 = s->prm.tol * s->prm.tol;
    if (s->G > 1) { const double t = std::max(s->prm.tol / std::sqrt((double)s->G), 1e-14); a.tol2 = std::min(a.tol2, t * t); }
    a.max_iter = s->prm.max_iter; a.which = 1; a.partials = s->partials; a.ctl = s->ctl;
    void* args[] = {(void*)&a, (void*)&s->pc};
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool prof = prof_on(s);
    if (prof) {
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, s->stream));
    }
    CU(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(NT), args, s->smem_e, s->stream));
    if (prof) {
        CU(cudaEventRecord(e1, s->stream));
        s->prof_ev.push_back(e0); s->prof_ev.push_back(e1); s->prof_which.push_back(1);
    }
    s->launches++;
    return 0;
}

// bulk-copy pipelined M_q solve (kernels_pipe.cuh): 1024 threads, one CTA per SM, two shared-memory stages
int launch_mq_pipe(phw_solver_s* s) {
    constexpr int NT = 1024;
    // PHW_MQ_PIPE_DIAG=1 (A/B): Jacobi weights recomputed from the staged cell records instead of streamed - 12 % fewer DRAM bytes
    // but measured slower (0.59 vs 0.40 ms per iteration: the extra shared-memory reads and the division cost more than the stream)
    static const int diag_env = [] { const char* e = getenv("PHW_MQ_PIPE_DIAG"); return e ? atoi(e) : 0; }();
    const bool implicit = (s->prm.scheme == PHW_IE || s->prm.scheme == PHW_SM);
    const bool diag = diag_env != 0 && !(implicit && s->prm.casimir);
    const bool multi = s->pc.world > 1;
    void* kern = multi ? (void*)solve_mq_pipe_k<NT, false, true> : (diag ? (void*)solve_mq_pipe_k<NT, true, false> : (void*)solve_mq_pipe_k<NT, false, false>);
    CU(cudaMemsetAsync(&s->ctl->gbar[0], 0, sizeof(unsigned int), s->stream));
    { int rc_ = ensure_dyn_smem(s->device, (const void*)kern, 227 * 1024 - 1024); if (rc_) return rc_; }
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, s->smem_pq));
    if (occ < 1) return fail(PHW_ECUDA, "pipelined M_q solve kernel does not fit on an SM");
    const int grid = std::max(1, std::min(s->npatch, occ * s->sm_count));
    MqPipeArgs a = s->pq;
    a.npatch = s->npatch; a.hdr = s->dm.hdr; a.cellq = s->d_cellq;
    a.ghost_e = s->dm.ghost_e; a.eadj = s->dm.eadj; a.b = s->bq; a.pinv = s->pinv_q;
    for (int i = 0; i < 3; ++i) a.buf[i] = s->tq.b[i];
    a.cheb_d = s->cheb_d[1]; a.cheb_c = s->cheb_c[1];
    a.tol2 = s->prm.tol * s->prm.tol;
    if (s->G > 1) { const double t = std::max(s->prm.tol / std::sqrt((double)s->G), 1e-14); a.tol2 = std::min(a.tol2, t * t); }
    a.max_iter = s->prm.max_iter; a.which = 1; a.partials = s->partials; a.ctl = s->ctl;
    void* args[] = {(void*)&a, (void*)&s->pc};
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool prof = prof_on(s);
    if (prof) {
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, s->stream));
    }
    CU(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(NT), args, s->smem_pq, s->stream));
    if (prof) {
        CU(cudaEventRecord(e1, s->stream));
        s->prof_ev.push_back(e0); s->prof_ev.push_back(e1); s->prof_which.push_back(1);
    }
    s->launches++;
    s->kernel_paths |= PHW_PATH_PIPE_Q;
    return 0;
}

// bulk-copy pipelined ELL M_p solve: two CTAs of 512 threads per SM when two pairs of stages fit, else one of 1024
template <int NT, int MINB>
int launch_ell_pipe_cfg(phw_solver_s* s) {
    const bool multi = s->pc.world > 1;
    void* kern = multi ? (void*)solve_ell_pipe_k<NT, MINB, true> : (void*)solve_ell_pipe_k<NT, MINB, false>;
    CU(cudaMemsetAsync(&s->ctl->gbar[0], 0, sizeof(unsigned int), s->stream));
    { int rc_ = ensure_dyn_smem(s->device, (const void*)kern, 227 * 1024 - 1024); if (rc_) return rc_; }
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, s->smem_pp));
    if (occ < 1) return fail(PHW_ECUDA, "pipelined M_p solve kernel does not fit on an SM");
    const int grid = std::max(1, std::min(s->npatch, std::min(occ, MINB) * s->sm_count));
    EllPipeArgs a = s->pp;
    a.npatch = s->npatch; a.hdr = s->dm.hdr; a.ell_ioff = s->dm.ell_ioff; a.ell_woff = s->dm.ell_woff; a.ell_ids = s->dm.ell_ids; a.ell_w = s->dm.ell_w;
    a.ghost_v = s->dm.ghost_v; a.b = s->bp;
    for (int i = 0; i < 3; ++i) a.buf[i] = s->tp.b[i];
    a.cheb_d = s->cheb_d[0]; a.cheb_c = s->cheb_c[0];
    a.tol2 = s->prm.tol * s->prm.tol;
    if (s->G > 1) { const double t = std::max(s->prm.tol / std::sqrt((double)s->G), 1e-14); a.tol2 = std::min(a.tol2, t * t); }
    a.max_iter = s->prm.max_iter; a.which = 0; a.partials = s->partials; a.ctl = s->ctl;
    void* args[] = {(void*)&a, (void*)&s->pc};
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool prof = prof_on(s);
    if (prof) {
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, s->stream));
    }
    CU(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(NT), args, s->smem_pp, s->stream));
    if (prof) {
        CU(cudaEventRecord(e1, s->stream));
        s->prof_ev.push_back(e0); s->prof_ev.push_back(e1); s->prof_which.push_back(0);
    }
    s->launches++;
    s->kernel_paths |= PHW_PATH_PIPE_P;
    return 0;
}
// A/B knob PHW_L2_PERSIST_B=1: pin the right-hand side of the M_p solve (100 MB at 50 M dofs, read once per iteration and
// unchanged during the solve) in the 126 MB L2 with an access-policy window on the solver's stream for the duration of the launch
static int l2_persist_window(phw_solver_s* s, const void* ptr, size_t bytes, bool on) {
    static const bool want = [] { const char* e = getenv("PHW_L2_PERSIST_B"); return e && atoi(e) != 0; }();
    if (!want) return 0;
    size_t& max_win = s->l2_max_win; size_t& max_persist = s->l2_max_persist;   // per solver (= per device), not per process
    if (max_win == 0) {
        cudaDeviceProp prop{};
        CU(cudaGetDeviceProperties(&prop, s->device));
        max_win = (size_t)prop.accessPolicyMaxWindowSize; max_persist = (size_t)prop.persistingL2CacheMaxSize;
        if (max_persist > 0) CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, max_persist));
    }
    if (max_win == 0 || max_persist == 0) return 0;
    cudaStreamAttrValue av{};
    av.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
    av.accessPolicyWindow.num_bytes = on ? std::min(bytes, max_win) : 0;
    av.accessPolicyWindow.hitRatio = on ? (float)std::min(1.0, (double)max_persist / (double)std::max<size_t>(1, std::min(bytes, max_win))) : 0.f;
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CU(cudaStreamSetAttribute(s->stream, cudaStreamAttributeAccessPolicyWindow, &av));
    return 0;
}
int launch_ell_pipe(phw_solver_s* s) {
    int rc = l2_persist_window(s, s->bp, (size_t)s->nv * sizeof(double), true);
    if (rc) return rc;
    rc = (2 * (s->smem_pp + 2048) <= 227 * 1024) ? launch_ell_pipe_cfg<512, 2>(s) : launch_ell_pipe_cfg<1024, 1>(s);
    if (rc) return rc;
    return l2_persist_window(s, s->bp, (size_t)s->nv * sizeof(double), false);
}

int solve_begin(phw_solver_s* s, int which) {
    solve_begin_k<<<1, 1, 0, s->stream>>>(s->ctl, which);
    s->launches++;
    CU(cudaGetLastError());
    return 0;
}

// Chebyshev solve of M_p x = bp (which = 0) or M_q x = bq (which = 1); warm start = tribuf[cur]
int solve_mass(phw_solver_s* s, int which, int n_launch) {
    int rc;
    if (s->bcsr) return bcsr_solve(s, which);
    if (s->ho_mf) return ho_mf_solve(s, which);
    if (s->ho) {
        CsrSolveArgs a{};
        a.A = which == PHW_P ? s->hMp : s->hMq; a.tb = which == PHW_P ? s->tp : s->tq;
        a.b = which == PHW_P ? s->bp : s->bq; a.pinv = which == PHW_P ? s->pinv_p : s->pinv_q;
        a.cheb_d = s->cheb_d[which]; a.cheb_c = s->cheb_c[which]; a.tol2 = s->prm.tol * s->prm.tol;
        a.max_iter = s->prm.max_iter; a.which = which; a.partials = s->partials; a.ctl = s->ctl;
        a.gbar = (s->prm.flags & 1024) ? 1 : 0;
        if (a.gbar) CU(cudaMemsetAsync(&s->ctl->gbar[0], 0, sizeof(unsigned int), s->stream));
        void* kern = a.A.lanes == 4 ? (void*)solve_coop_csr_k<4> : a.A.lanes == 8 ? (void*)solve_coop_csr_k<8>
                   : a.A.lanes == 16 ? (void*)solve_coop_csr_k<16> : (void*)solve_coop_csr_k<32>;
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0));
        const int want = nblk(a.A.n_rows * a.A.lanes);
        const int grid = std::max(1, std::min(want, std::min(occ, 4) * s->sm_count));
        void* args[] = {(void*)&a};
        CU(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(256), args, 0, s->stream));
        s->launches++;
        return 0;
    }
    RowArgs a = base_args(s);
    a.which = which; a.use_cheb_input = 1; a.finalize = 1; a.cM = 1.0; a.counter_id = which;
    a.cheb_d = s->cheb_d[which]; a.cheb_c = s->cheb_c[which];
    if (!s->prm.strong && !(s->prm.casimir && (s->prm.scheme == PHW_IE || s->prm.scheme == PHW_SM))) a.mdiag = nullptr;   // pinv_p is then exactly 1/mdiag_p
    if (s->coop) {
        a.b = which == PHW_P ? s->bp : s->bq;
        a.pinv = which == PHW_P ? s->pinv_p : s->pinv_q;
        if (s->pcg[which] && s->pc.world <= 1 && a.mdiag == nullptr) return which == PHW_P ? launch_pcg<0>(s, a) : launch_pcg<1>(s, a);
        const bool pipe_ok = s->pc.world <= 1 || (pipe_multi_mask() & 1);
        if (which == PHW_Q && s->pipe_q && pipe_ok) return launch_mq_pipe(s);
        if (which == PHW_P && s->pipe_p && pipe_ok && s->ell_p && a.mdiag == nullptr) return launch_ell_pipe(s);
        if (which == PHW_P && s->ell_p && a.mdiag == nullptr && !s->ell_staged) {
            s->kernel_paths |= PHW_PATH_LEAN_P;
            static const int cfg = [] { const char* e = getenv("PHW_ELL_CFG"); return e ? atoi(e) : 0; }();
            switch (cfg) {
                case 1: return launch_ell_lean<256, 6>(s);
                case 2: return launch_ell_lean<256, 8>(s);
                case 3: return launch_ell_lean<512, 3>(s);
                case 4: return launch_ell_lean<128, 12>(s);
                case 5: return launch_ell_lean<256, 5>(s);
                case 9: break;
                default: return launch_ell_lean<256, 4>(s);
            }
        }
        if (which == PHW_Q && !(s->prm.flags & 128)) {   // flags bit 7: generic kernels everywhere
            s->kernel_paths |= PHW_PATH_LEAN_Q;
            static const int cfg = [] { const char* e = getenv("PHW_MQ_CFG"); return e ? atoi(e) : 0; }();
            switch (cfg) {
                case 1: return launch_mq_lean<384, 2, 2, 2>(s);
                case 2: return launch_mq_lean<512, 2, 1, 2>(s);
                case 3: return launch_mq_lean<512, 2, 2, 1>(s);
                case 4: return launch_mq_lean<448, 2, 2, 2>(s);
                case 5: return launch_mq_lean<576, 2, 2, 2>(s);
                case 6: return launch_mq_lean<512, 2, 1, 1>(s);
                case 9: break;
                default: return launch_mq_lean<512, 2, 2, 2>(s);
            }
        }
        if (which == PHW_P && s->ell_p && a.mdiag == nullptr)      // pinv_p = 1 / diag(M_p) exactly
            return (s->ell_staged && s->smem_ells <= 110 * 1024) ? launch_coop<0, false, 2>(s, a, a) : launch_coop<0, false, 1>(s, a, a);
        return which == PHW_P ? launch_coop<0, false>(s, a, a) : launch_coop<1, false>(s, a, a);
    }
    rc = solve_begin(s, which);
    if (rc) return rc;
    for (int it = 0; it < n_launch; ++it) {
        if (which == PHW_P) { a.b = s->bp; a.pinv = s->pinv_p; rc = launch_vertex<EPI_CHEB, MODE_OP, true, false, false>(s, a); }
        else { a.b = s->bq; a.pinv = s->pinv_q; rc = launch_edge<EPI_CHEB, MODE_OP, true, false, false>(s, a); }
        if (rc) return rc;
    }
    return 0;
}

// Chebyshev (ellipse) solve of the coupled block system
//   M_p v_p + th dt K (D^T v_q + B_top v_p) = bp ;  M_q v_q - th dt/rho (D v_p - B_right v_q) = bq
int solve_block(phw_solver_s* s, double theta, int n_launch) {
    int rc;
    const double dt = s->prm.dt, K = s->prm.K_wave, rho = s->prm.rho;
    if (s->ho) {
        // higher-order pairs: assembled CSR rows, all iterations in one cooperative launch (solve_block_csr_k)
        if (s->ho_mf) return fail(PHW_ESTATE, "the implicit schemes of the higher-order pairs run on the assembled operators (phw_ho_create)");
        CsrBlockArgs a{};
        a.Mp = s->hMp; a.Mq = s->hMq; a.D = s->hD; a.DT = s->hDT; a.tp = s->tp; a.tq = s->tq;
        a.bp = s->bp; a.bq = s->bq; a.pinv_p = s->pinv_p; a.pinv_q = s->pinv_q;
        a.cDv = theta * dt * K; a.cDe = -theta * dt / rho; a.wv = 1.0 / K; a.we = rho;
        a.cheb_d = s->cheb_d[2]; a.cheb_c = s->cheb_c[2]; a.tol2 = s->prm.tol * s->prm.tol; a.max_iter = n_launch;
        a.partials = s->partials; a.ctl = s->ctl;
        CU(cudaMemsetAsync(&s->ctl->gbar[0], 0, sizeof(unsigned int), s->stream));
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (void*)solve_block_csr_k, 256, 0));
        if (occ < 1) return fail(PHW_ECUDA, "solve_block_csr_k does not fit on an SM");
        const int grid = std::max(1, std::min(nblk((s->nv + s->ne) * 8), std::min(occ, 4) * s->sm_count));
        void* args[] = {(void*)&a};
        CU(cudaLaunchCooperativeKernel((void*)solve_block_csr_k, dim3(grid), dim3(256), args, 0, s->stream));
        s->launches++;
        return 0;
    }
    RowArgs a = base_args(s);
    a.which = 2; a.use_cheb_input = 1; a.coupled_from_tribuf = 1; a.cM = 1.0;
    a.cheb_d = s->cheb_d[2]; a.cheb_c = s->cheb_c[2];
    const bool cas = s->prm.casimir != 0;
    if (s->coop) {
        RowArgs av = a, ae = a;
        av.b = s->bp; av.pinv = s->pinv_p; av.cD = theta * dt * K; av.cB = cas ? theta * dt * K : 0.0; av.weight = 1.0 / K;
        ae.b = s->bq; ae.pinv = s->pinv_q; ae.cD = -theta * dt / rho; ae.cB = cas ? theta * dt / rho : 0.0; ae.weight = rho;
        return cas ? launch_coop<2, true>(s, av, ae) : launch_coop<2, false>(s, av, ae);
    }
    rc = solve_begin(s, 2);
    if (rc) return rc;
    for (int it = 0; it < n_launch; ++it) {
        RowArgs av = a;
        av.b = s->bp; av.pinv = s->pinv_p; av.cD = theta * dt * K; av.cB = cas ? theta * dt * K : 0.0; av.finalize = 0; av.weight = 1.0 / K; av.counter_id = 2;
        rc = cas ? launch_vertex<EPI_CHEB, MODE_OP, true, true, true>(s, av) : launch_vertex<EPI_CHEB, MODE_OP, true, true, false>(s, av);
        if (rc) return rc;
        RowArgs ae = a;
        ae.b = s->bq; ae.pinv = s->pinv_q; ae.cD = -theta * dt / rho; ae.cB = cas ? theta * dt / rho : 0.0; ae.finalize = 1; ae.weight = rho; ae.counter_id = 3;
        rc = cas ? launch_edge<EPI_CHEB, MODE_OP, true, true, true>(s, ae) : launch_edge<EPI_CHEB, MODE_OP, true, true, false>(s, ae);
        if (rc) return rc;
    }
    return 0;
}

// sum ctl->red[first..first+n) (what = 0) or ctl->flux[...] (what = 1) over all ranks (no-op on one GPU)
int allreduce(phw_solver_s* s, int what, int first, int n) {
    if (s->pc.world <= 1) return 0;
    allreduce_k<<<1, 32, 0, s->stream>>>(s->pc, s->ctl, what, first, n);
    s->launches++;
    CU(cudaGetLastError());
    return 0;
}
int scalar(phw_solver_s* s, int op, double aux0 = 0.0, double aux1 = 0.0) {
    scalar_k<<<nblk(s->G, 64), 64, 0, s->stream>>>(s->ctl, s->d_sp, s->G, op, s->dH, (long long)s->run_rows, aux0, aux1);
    s->launches++;
    CU(cudaGetLastError());
    return 0;
}
int flux(phw_solver_s* s, const double* q, int which, int slot) {
    if (s->bcsr) {
        flux_b_k<<<nblk(s->bcsr, 128), 128, 0, s->stream>>>(s->nport, s->d_port_e, s->d_port_bl, q, s->tq, s->ctl, which, slot, s->bcsr);
        s->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    flux_k<<<s->G, s->G > 1 ? 32 : 256, 0, s->stream>>>(s->d_port_goff, s->d_port_e, s->d_port_bl, q, s->tq, s->ctl, which, slot);
    s->launches++;
    CU(cudaGetLastError());
    return allreduce(s, 1, slot, 1);      // only the rank holding the port contributes
}
int axpy(phw_solver_s* s, int n, double* y, double a, const double* x, TriBuf tb, int which) {
    axpy_k<<<vec_grid(s, n), 256, 0, s->stream>>>(n, y, a, x, tb, s->ctl, which);
    s->launches++;
    CU(cudaGetLastError());
    return 0;
}
int waxpy(phw_solver_s* s, int n, double* w, const double* y, double a, const double* x, TriBuf tb, int which) {
    waxpy_k<<<vec_grid(s, n), 256, 0, s->stream>>>(n, w, y, a, x, tb, s->ctl, which);
    s->launches++;
    CU(cudaGetLastError());
    return 0;
}

// b_p = -K (D^T qsrc [+ B_top psrc])          (explicit part of the p rows, velocity form)
int rhs_p(phw_solver_s* s, const double* qsrc, const double* psrc) {
    const double K = s->prm.K_wave;
    if (s->bcsr) { bcsr_store(s, s->hDT, qsrc, -1.0, s->d_gK, s->bp); CU(cudaGetLastError()); return 0; }
    if (s->ho_mf) return ho_mf_DT(s, qsrc, -K, s->bp);
    if (s->ho) {
        csr_store(s, s->hDT, qsrc, -K, s->bp);
        s->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    if (s->G > 1) return op_vertex_store(s, psrc, qsrc, 0.0, -1.0, 0.0, s->bp, s->d_gK);
    return op_vertex_store(s, psrc, qsrc, 0.0, -K, s->prm.casimir ? -K : 0.0, s->bp);
}
// b_q = (1/rho) (D psrc - b_L p_left [- B_right qsrc])
int rhs_q(phw_solver_s* s, const double* psrc, const double* qsrc) {
    const double ir = (s->G > 1) ? 1.0 : 1.0 / s->prm.rho;
    int rc = 0;
    if (s->bcsr) {
        bcsr_store(s, s->hD, psrc, 1.0, s->d_ginvrho, s->bq);
        if (s->nport > 0) {
            const long long nt = (long long)s->nport * s->bcsr;
            port_rhs_b_k<<<(unsigned)((nt + 255) / 256), 256, 0, s->stream>>>(s->nport, s->d_port_e, s->d_port_bl, s->d_ginvrho, s->bq, 1.0, s->ctl, s->bcsr);
            s->launches++;
        }
        CU(cudaGetLastError());
        return 0;
    }
    if (s->ho_mf) {
        rc = ho_mf_D(s, psrc, ir, s->bq);
    } else if (s->ho) {
        csr_store(s, s->hD, psrc, ir, s->bq);
        s->launches++;
        CU(cudaGetLastError());
    } else {
        rc = (s->G > 1) ? op_edge_store(s, psrc, qsrc, 0.0, 1.0, 0.0, s->bq, s->d_ginvrho)
                        : op_edge_store(s, psrc, qsrc, 0.0, ir, s->prm.casimir ? -ir : 0.0, s->bq);
    }
    if (rc) return rc;
    if (!s->prm.strong && s->nport > 0) {
        port_rhs_k<<<nblk(s->nport), 256, 0, s->stream>>>(s->nport, s->d_port_e, s->d_port_bl, s->d_port_group,
                                                          s->G > 1 ? s->d_ginvrho : nullptr, s->bq, ir, s->ctl);
        s->launches++;
        CU(cudaGetLastError());
    }
    return 0;
}
int energy(phw_solver_s* s) {
    if (s->bcsr) {
        int rc_ = bcsr_dot(s, s->hMp, s->p, 0.5, s->d_ginvrho, 0); if (rc_) return rc_;
        rc_ = bcsr_dot(s, s->hMq, s->q, 0.5, s->d_gK, 1); if (rc_) return rc_;
        CU(cudaGetLastError());
        return 0;
    }
    if (s->ho) {
        if (s->ho_mf) {
            int rc_ = ho_mf_quad(s, HO_MP, s->p, 0.5 / s->prm.rho, 0, 6); if (rc_) return rc_;
            rc_ = ho_mf_quad(s, HO_MQ, s->q, 0.5 * s->prm.K_wave, 1, 7); if (rc_) return rc_;
        } else {
            csr_dot(s, s->hMp, s->p, 0.5 / s->prm.rho, 0, 6);
            csr_dot(s, s->hMq, s->q, 0.5 * s->prm.K_wave, 1, 7);
            s->launches += 2;
        }
        if (s->prm.analytical && s->an_ready && s->in_step) {
            const int nb = 2 * s->sm_count;
            diag_cells_ho_k<<<nb, 128, 0, s->stream>>>(s->ho_nc, s->ho_ndp, s->ho_nn, s->ho_cell_dofs, s->ho_T, s->ho_mass, s->ho_exact_nodes,
                                                     s->ho_area, s->p, s->ho_omega, s->ctl, s->d_diag_part);
            diag_finish_ho_k<<<1, 1024, 0, s->stream>>>(s->nv, nb, s->ho_exact_dofs, s->p, s->ho_omega, s->ho_nprobe, s->ho_probe_dofs,
                                                        s->ho_probe_w, s->ctl, s->d_diag_part);
            s->launches += 2;
        }
        CU(cudaGetLastError());
        return 0;
    }
    if (s->G > 1) {   // batched: per-(patch, warp) partials -> one reduction block per case
        int rc = op_vertex_dot(s, s->p, 1.0, 0.0, 0, 1.0, 4); if (rc) return rc;
        group_reduce_k<<<s->G, 32, 0, s->stream>>>(s->wpart, NT_ROWS / 32, s->d_gpoff, s->d_ginvrho, 0.5, s->ctl, 0);
        rc = op_edge_dot(s, s->q, 1.0, 0.0, 1, 1.0, 5); if (rc) return rc;
        group_reduce_k<<<s->G, 32, 0, s->stream>>>(s->wpart, NT_ROWS / 32, s->d_gpoff, s->d_gK, 0.5, s->ctl, 1);
        s->launches += 2;
        CU(cudaGetLastError());
        return 0;
    }
    int rc = 0;
    if (s->pipe_en && (s->pc.world <= 1 || (pipe_multi_mask() & 4))) {   // both forms in one bulk-copy pipelined pass over the own cells (kernels_pipe.cuh)
        constexpr int NT = 1024;
        { int rc_ = ensure_dyn_smem(s->device, (const void*)energy_pipe_k<NT>, 227 * 1024 - 1024); if (rc_) return rc_; }
        EnergyPipeArgs a = s->pe;
        a.npatch = s->npatch; a.hdr = s->dm.hdr; a.recv = s->dm.recv; a.area = s->dm.area; a.cellq = s->d_cellq;
        a.ghost_v = s->dm.ghost_v; a.ghost_e = s->dm.ghost_e; a.p = s->p; a.q = s->q;
        a.scale_p = 0.5 / s->prm.rho / 12.0; a.scale_q = 0.5 * s->prm.K_wave; a.counter_id = 4;
        a.partials = s->partials + 8192 + 4096; a.ctl = s->ctl;
        energy_pipe_k<NT><<<std::max(1, std::min(s->npatch, s->sm_count)), NT, s->smem_pe, s->stream>>>(a);
        s->launches += 1;
        s->kernel_paths |= PHW_PATH_PIPE_ENERGY;
        CU(cudaGetLastError());
    } else if (!(s->prm.flags & 128)) {   // lean reductions on the ELL rows / packed cell records
        DotArgs a{};
        a.npatch = s->npatch; a.hdr = s->dm.hdr; a.partials = s->partials + 8192; a.ctl = s->ctl;
        a.ell_ioff = s->dm.ell_ioff; a.ell_woff = s->dm.ell_woff; a.ell_ids = s->dm.ell_ids; a.ell_w = s->dm.ell_w;
        a.ghost = s->dm.ghost_v; a.x = s->p; a.scale = 0.5 / s->prm.rho; a.slot = 0; a.counter_id = 4;
        { int rc_ = ensure_dyn_smem(s->device, (const void*)dot_ell_k<256, 4>, 200 * 1024); if (rc_) return rc_; }
        { int rc_ = ensure_dyn_smem(s->device, (const void*)dot_mq_k<512, 2, 2>, 200 * 1024); if (rc_) return rc_; }
        dot_ell_k<256, 4><<<std::max(1, std::min(s->npatch, 4 * s->sm_count)), 256, s->smem_ell, s->stream>>>(a);
        a.cellq = s->d_cellq; a.eadj = s->dm.eadj; a.ghost = s->dm.ghost_e; a.x = s->q; a.scale = 0.5 * s->prm.K_wave; a.slot = 1; a.counter_id = 5;
        a.partials = s->partials + 8192 + 2048;
        dot_mq_k<512, 2, 2><<<std::max(1, std::min(s->npatch, 2 * s->sm_count)), 512, s->smem_e, s->stream>>>(a);
        s->launches += 2;
        CU(cudaGetLastError());
    } else {
        rc = op_vertex_dot(s, s->p, 1.0, 0.0, 0, 0.5 / s->prm.rho, 4);
        if (rc) return rc;
        rc = op_edge_dot(s, s->q, 1.0, 0.0, 1, 0.5 * s->prm.K_wave, 5);
        if (rc) return rc;
    }
    if (s->prm.analytical && s->an_ready && s->in_step) {
        const int nb = 4 * s->sm_count;
        diag_cells_k<<<nb, 256, 0, s->stream>>>((int)s->mesh->L.nc, s->d_cells_new, s->d_vtx_new, s->p, s->an, s->ctl, s->d_diag_part);
        diag_finish_k<<<1, 1024, 0, s->stream>>>(s->nv, nb, s->d_exact_v, s->p, s->an, s->ctl, s->d_diag_part);
        s->launches += 2;
        CU(cudaGetLastError());
    }
    return allreduce(s, 0, 0, 2);
}

// initial guess of a solve from the history of its site (field 0 = p unknown, 1 = q unknown)
int predict(phw_solver_s* s, int field, int site, int which) {
    phw_solver_s::Hist& H = s->hist[field][site];
    const int order = s->extrap_f[field] + 1;
    const int m = std::min(H.count, order);
    if (m <= 0 || !H.h[0]) return 0;
    // extrapolation through the last m solutions: c_j = (-1)^j binom(m, j + 1)
    PredictArgs pa{};
    pa.m = m;
    double binom = 1.0;
    for (int j = 0; j < m; ++j) {
        binom = binom * (double)(m - j) / (double)(j + 1);
        pa.c[j] = (j % 2 == 0) ? binom : -binom;
        pa.h[j] = H.h[((H.head - j) % order + order) % order];
    }
    const int n = field == 0 ? s->nv : s->ne;
    predict_k<<<vec_grid(s, n), 256, 0, s->stream>>>(n, field == 0 ? s->tp : s->tq, s->ctl, which, pa);
    s->launches++;
    CU(cudaGetLastError());
    return 0;
}
int remember(phw_solver_s* s, int field, int site, int which) {
    phw_solver_s::Hist& H = s->hist[field][site];
    if (!H.h[0]) return 0;
    const int order = s->extrap_f[field] + 1;
    H.head = (H.head + 1) % order;
    H.count = std::min(H.count + 1, order);
    const int n = field == 0 ? s->nv : s->ne;
    copy_from_tb_k<<<vec_grid(s, n), 256, 0, s->stream>>>(n, H.h[H.head], field == 0 ? s->tp : s->tq, s->ctl, which);
    s->launches++;
    CU(cudaGetLastError());
    return 0;
}
int dirichlet_preset(phw_solver_s* s, int which_ctl) {
    if (!s->prm.strong || s->n_dir == 0) return 0;
    dirichlet_set_k<<<nblk(s->n_dir), 256, 0, s->stream>>>(s->n_dir, s->d_dir_idx, s->d_dir_w, s->tp, s->ctl, which_ctl, s->p, 1.0 / s->prm.dt);
    s->launches++;
    CU(cudaGetLastError());
    return 0;
}
int solve_site(phw_solver_s* s, int which, int site) {
    int rc = predict(s, which, site, which);
    if (rc) return rc;
    if (which == PHW_P) { rc = dirichlet_preset(s, PHW_P); if (rc) return rc; }
    rc = solve_mass(s, which, s->prm.max_iter);
    if (rc) return rc;
    return remember(s, which, site, which);
}
int solve_block_site(phw_solver_s* s, double theta) {
    int rc = predict(s, 0, 0, 2); if (rc) return rc;
    rc = predict(s, 1, 0, 2); if (rc) return rc;
    rc = dirichlet_preset(s, 2); if (rc) return rc;
    rc = solve_block(s, theta, s->prm.max_iter); if (rc) return rc;
    rc = remember(s, 0, 0, 2); if (rc) return rc;
    return remember(s, 1, 0, 2);
}
void reset_history(phw_solver_s* s) {
    for (int f = 0; f < 2; ++f)
        for (int k = 0; k < 2; ++k) { s->hist[f][k].count = 0; s->hist[f][k].head = 0; }
    s->have_vq = false;
}

// ---- one time step of each scheme (wave_2D.py:787-982) ------------------------------------------
int step(phw_solver_s* s) {
    const double dt = s->prm.dt;
    const int nv = s->nv, ne = s->ne;
    const bool ic = s->prm.interconnection != 0;
    int rc = 0;
#define RC(x) do { rc = (x); if (rc) return rc; } while (0)
    switch (s->prm.scheme) {
        case PHW_EE:
            RC(scalar(s, S_BEGIN_1SUB));
            RC(flux(s, s->q, 1, 0));
            RC(rhs_p(s, s->q, s->p)); RC(solve_site(s, PHW_P, 0));
            RC(rhs_q(s, s->p, s->q)); RC(solve_site(s, PHW_Q, 0));
            RC(axpy(s, nv, s->p, dt, nullptr, s->tp, 0));
            RC(axpy(s, ne, s->q, dt, nullptr, s->tq, 1));
            RC(energy(s));
            RC(scalar(s, S_END_1SUB));
            break;
        case PHW_SE:
            RC(scalar(s, S_BEGIN_1SUB));
            RC(flux(s, s->q, 1, 0));
            RC(rhs_p(s, s->q, s->p)); RC(solve_site(s, PHW_P, 0));
            RC(axpy(s, nv, s->p, dt, nullptr, s->tp, 0));
            if (ic) RC(scalar(s, S_ODE_SE));
            RC(rhs_q(s, s->p, s->q)); RC(solve_site(s, PHW_Q, 0));
            RC(axpy(s, ne, s->q, dt, nullptr, s->tq, 1));
            RC(energy(s));
            RC(scalar(s, S_END_1SUB));
            break;
        case PHW_SV:
            RC(scalar(s, S_SV_STAGE1));
            if (!s->have_vq || (s->prm.flags & 1)) { RC(rhs_q(s, s->p, s->q)); RC(solve_site(s, PHW_Q, 0)); }
            RC(axpy(s, ne, s->q, 0.5 * dt, nullptr, s->tq, 1));       // q_temp
            RC(flux(s, s->q, 1, 0));
            RC(scalar(s, S_SV_MID));
            RC(rhs_p(s, s->q, s->p)); RC(solve_site(s, PHW_P, 0));
            RC(axpy(s, nv, s->p, dt, nullptr, s->tp, 0));
            if (ic) RC(scalar(s, S_ODE_SV));
            RC(rhs_q(s, s->p, s->q)); RC(solve_site(s, PHW_Q, 0));
            RC(axpy(s, ne, s->q, 0.5 * dt, nullptr, s->tq, 1));
            s->have_vq = true;
            RC(energy(s));
            RC(scalar(s, S_END_SV));
            break;
        case PHW_EH:
            RC(scalar(s, S_EH_BEGIN));
            RC(flux(s, s->q, 1, 0));
            RC(rhs_p(s, s->q, s->p)); RC(solve_site(s, PHW_P, 0));
            RC(rhs_q(s, s->p, s->q)); RC(solve_site(s, PHW_Q, 0));
            RC(flux(s, nullptr, 1, 1));                               // b_L^T v_q1
            RC(waxpy(s, nv, s->tmp_p, s->p, 0.5 * dt, nullptr, s->tp, 0));   // 0.5 (p_m + p_temp)
            RC(waxpy(s, ne, s->tmp_q, s->q, 0.5 * dt, nullptr, s->tq, 1));
            RC(rhs_p(s, s->tmp_q, s->tmp_p)); RC(solve_site(s, PHW_P, 1));
            RC(rhs_q(s, s->tmp_p, s->tmp_q)); RC(solve_site(s, PHW_Q, 1));
            RC(axpy(s, nv, s->p, dt, nullptr, s->tp, 0));
            RC(axpy(s, ne, s->q, dt, nullptr, s->tq, 1));
            RC(energy(s));
            RC(scalar(s, S_END_EH));
            break;
        case PHW_IE:
        case PHW_SM: {
            const double theta = (s->prm.scheme == PHW_IE) ? 1.0 : 0.5;
            RC(scalar(s, S_BEGIN_1SUB));
            RC(flux(s, s->q, 1, 0));
            if (ic) RC(scalar(s, S_SM_PLEFT));
            RC(rhs_p(s, s->q, s->p));
            RC(rhs_q(s, s->p, s->q));
            RC(solve_block_site(s, theta));
            if (ic) {
                RC(flux(s, nullptr, 2, 2));                            // b_L^T vhat_q
                RC(scalar(s, S_ODE_SM, s->blw));
                axpy_dev_k<<<vec_grid(s, nv), 256, 0, s->stream>>>(nv, s->tp, s->ctl, 2, &s->ctl->xi1, s->wp);
                axpy_dev_k<<<vec_grid(s, ne), 256, 0, s->stream>>>(ne, s->tq, s->ctl, 2, &s->ctl->xi1, s->wq);
                s->launches += 2;
                CU(cudaGetLastError());
            }
            if (s->prm.casimir) {
                // boundary energy of the impedance sides uses the midpoint averages (wave_2D.py:633-634)
                RC(waxpy(s, nv, s->tmp_p, s->p, 0.5 * dt, nullptr, s->tp, 2));
                RC(waxpy(s, ne, s->tmp_q, s->q, 0.5 * dt, nullptr, s->tq, 2));
                RC(op_vertex_dot(s, s->tmp_p, 0.0, 1.0, 2, 1.0, 6));
                RC(op_edge_dot(s, s->tmp_q, 0.0, 1.0, 3, 1.0, 7));
                RC(allreduce(s, 0, 2, 2));
            }
            RC(axpy(s, nv, s->p, dt, nullptr, s->tp, 2));
            RC(axpy(s, ne, s->q, dt, nullptr, s->tq, 2));
            RC(flux(s, s->q, 1, 1));
            RC(energy(s));
            RC(scalar(s, S_END_1SUB));
        } break;
        default:
            return fail(PHW_EINVAL, "unknown scheme");
    }
#undef RC
    return 0;
}

// ---- the whole time loop in one launch (kernels_loop.cuh; flags bit 15) -----------------------------------------------------------
// eligible: matrix-free solver on one GPU, no analytical diagnostics, one patch per CTA with all CTAs co-resident; the CTAs of a case
// form one thread-block cluster when there are 1, 2, 4, 8 or 16 of them, else (single case only) one cooperative grid
// 512 threads per CTA for a single case (latency per pass), 256 for a batched sweep (two CTAs per SM: the cases are independent,
// so residency - not the latency of one case - sets the throughput)
static void* loop_kernel(bool cluster, int nt, bool res = false) {
    if (res) return nt == 1024 ? (void*)loop_k<true, 1024, true> : (void*)loop_k<true, 512, true>;
    if (cluster) return nt == 256 ? (void*)loop_k<true, 256> : (void*)loop_k<true, 512>;
    return nt == 256 ? (void*)loop_k<false, 256> : (void*)loop_k<false, 512>;
}
// shared-memory carve-up of the resident solves: maxima over all patches, every block 16-byte aligned
static ResCarve res_carve(const Layout& L, size_t scratch_bytes) {
    int mv = 0, mlv = 0, me = 0, mle = 0, mec = 0, mids = 0, mw = 0, mg = 0;
    for (int32_t ip = 0; ip < L.npatch_rows; ++ip) {
        const PatchHdr& h = L.hdr[ip];
        mv = std::max(mv, h.v_cnt); mlv = std::max(mlv, h.v_cnt + h.gv_cnt);
        me = std::max(me, h.e_cnt); mle = std::max(mle, h.e_cnt + h.ge_cnt);
        mec = std::max(mec, h.n_ecells); mids = std::max(mids, h.ell_in); mw = std::max(mw, h.ell_wn);
        mg = std::max(mg, std::max(h.gv_cnt, h.ge_cnt));
    }
    auto al = [](size_t b) { return (int)((b + 15) & ~(size_t)15); };
    ResCarve c{};
    int o = 0;
    c.xq[0] = o; o += al(8 * (size_t)mle); c.xq[1] = o; o += al(8 * (size_t)mle);
    c.bq = o; o += al(8 * (size_t)me);
    const int tq = o;
    o = 0;
    c.xp[0] = o; o += al(8 * (size_t)mlv); c.xp[1] = o; o += al(8 * (size_t)mlv); c.bp = o; o += al(8 * (size_t)mv);
    o = std::max(std::max(o, tq), al(scratch_bytes));     // the transient region also serves as the scratch of the row passes
    const int nsl = (mv + 31) / 32 + 2;
    c.sio = o; o += al(4 * (size_t)nsl); c.swo = o; o += al(4 * (size_t)nsl);
    c.ids = o; o += al(2 * (size_t)mids); c.w = o; o += al(8 * (size_t)mw);
    c.me = (me + 7) & ~7;
    c.wq = o; o += al(32 * (size_t)c.me); c.idq = o; o += al(8 * (size_t)c.me); c.dgq = o; o += al(8 * (size_t)me); c.dgp = o; o += al(8 * (size_t)mv);
    c.gsv = o; o += al(4 * (size_t)mg); c.gse = o; o += al(4 * (size_t)mg);
    c.part = o; o += 32;
    c.total = o;
    return c;
}
int setup_loop(phw_solver_s* s) {
    const int NT_LOOP = s->G > 1 ? 256 : 512;
    s->loop_on = false;
    if (!(s->prm.flags & 32768) || s->ho || !s->coop || s->prm.analytical || s->npatch < 1 || s->npatch % s->G != 0) return 0;
    const int cl = s->npatch / s->G;
    const Layout& L = s->mesh->L;
    if (L.nv_owned != L.nv || L.ne_owned != L.ne) return 0;            // partitioned layouts take the multi-launch path
    if (s->pcg[0] || s->pcg[1]) return 0;                              // conjugate-gradient mass solves (poor meshes): multi-launch path
    for (int g = 0; g < s->G; ++g) if (L.group_patch_off[g + 1] - L.group_patch_off[g] != cl) return 0;
    size_t sm = s->smem_vd;
    const bool cluster = (cl == 1 || cl == 2 || cl == 4 || cl == 8 || cl == 16);
    // resident solves: every operand of a Chebyshev iteration in shared memory (one CTA per SM); needs the ELL rows of M_p, local ids
    // below 2^14 and everything of the largest patch within the 227 KB of an SM.  PHW_NO_RES=1: the L2-resident solves (A/B)
    s->loop_res = false;
    if (cluster && s->ell_p && s->d_cellq && !getenv("PHW_NO_RES")) {
        const ResCarve c = res_carve(L, s->smem_vd);
        if (c.total <= 224 * 1024) { s->loop_res = true; s->loop_rc = c; sm = (size_t)c.total; }
    }
    const int NT_RES = getenv("PHW_RES_NT") ? atoi(getenv("PHW_RES_NT")) : 512;
    if (cluster) {
        void* kern = loop_kernel(true, s->loop_res ? NT_RES : NT_LOOP, s->loop_res);
        { int rc_ = ensure_dyn_smem(s->device, kern, s->loop_res ? 224 * 1024 : 200 * 1024); if (rc_) return rc_; }
        if (cl > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(s->npatch); cfg.blockDim = dim3(s->loop_res ? NT_RES : NT_LOOP); cfg.dynamicSmemBytes = sm; cfg.stream = s->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg) != cudaSuccess || ncl < 1) { cudaGetLastError(); return 0; }
    } else {
        if (s->G != 1) return 0;
        void* kern = loop_kernel(false, NT_LOOP);
        { int rc_ = ensure_dyn_smem(s->device, kern, 200 * 1024); if (rc_) return rc_; }
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT_LOOP, sm));
        if (occ < 1 || s->npatch > occ * s->sm_count) return 0;
    }
    int rc = s->dalloc(&s->loop_partials, 4 * (size_t)s->npatch + 64);
    if (rc) return rc;
    s->loop_on = true; s->loop_cl = cl; s->loop_cluster = cluster; s->loop_nt = (cluster && s->loop_res) ? NT_RES : NT_LOOP;
    if (!cluster) s->loop_res = false;
    return 0;
}

int run_loop(phw_solver_s* s, int64_t n_steps) {
    const phw_params& P = s->prm;
    LoopArgs a{};
    a.m = s->dm; a.tp = s->tp; a.tq = s->tq;
    a.p = s->p; a.q = s->q; a.bp = s->bp; a.bq = s->bq; a.tmp_p = s->tmp_p; a.tmp_q = s->tmp_q;
    a.pinv_p = s->pinv_p; a.pinv_q = s->pinv_q; a.mdiag_p = s->mdiag_p; a.wp = s->wp; a.wq = s->wq;
    for (int f = 0; f < 2; ++f)
        for (int k = 0; k < 2; ++k) {
            for (int j = 0; j < 6; ++j) a.hist[f][k][j] = s->hist[f][k].h[j];
            a.hist_head[f][k] = s->hist[f][k].head; a.hist_count[f][k] = s->hist[f][k].count;
        }
    a.extrap[0] = s->extrap_f[0]; a.extrap[1] = s->extrap_f[1]; a.n_sites = s->n_sites;
    a.have_vq = s->have_vq ? 1 : 0; a.sv_noreuse = (P.flags & 1) ? 1 : 0;
    a.port_goff = s->d_port_goff; a.port_e = s->d_port_e; a.port_bl = s->d_port_bl; a.nport = s->nport;
    a.gK = s->G > 1 ? s->d_gK : nullptr; a.ginvrho = s->G > 1 ? s->d_ginvrho : nullptr;
    a.n_dir = s->n_dir; a.dir_idx = s->d_dir_idx; a.dir_w = s->d_dir_w;
    a.sp = s->d_sp; a.Hrows = s->dH; a.rows_per_group = s->run_rows;
    a.ctl = s->ctl; a.partials = s->loop_partials;
    a.scheme = P.scheme; a.ic = P.interconnection; a.strong = P.strong; a.casimir = P.casimir; a.ell_p = s->ell_p ? 1 : 0;
    a.G = s->G; a.cl = s->loop_cl;
    a.dt = P.dt; a.K = P.K_wave; a.rho = P.rho; a.blw = s->blw;
    a.tol2 = P.tol * P.tol; a.xscale = P.strong ? 4.0 : 0.0;
    a.max_iter = P.max_iter;
    for (int i = 0; i < 3; ++i) { a.cheb_d[i] = s->cheb_d[i]; a.cheb_c[i] = s->cheb_c[i]; }
    a.n_steps = n_steps;
    a.cellq = s->d_cellq; a.rc = s->loop_rc;
    const size_t sm = s->loop_res ? (size_t)s->loop_rc.total : s->smem_vd;
    const int NT_LOOP = s->loop_nt;
    if (s->loop_cluster) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(s->npatch); cfg.blockDim = dim3(NT_LOOP); cfg.dynamicSmemBytes = sm; cfg.stream = s->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = s->loop_cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (s->loop_res) { if (NT_LOOP == 1024) CU(cudaLaunchKernelEx(&cfg, loop_k<true, 1024, true>, a)); else CU(cudaLaunchKernelEx(&cfg, loop_k<true, 512, true>, a)); }
        else if (NT_LOOP == 256) CU(cudaLaunchKernelEx(&cfg, loop_k<true, 256>, a)); else CU(cudaLaunchKernelEx(&cfg, loop_k<true, 512>, a));
        if (s->loop_res) s->kernel_paths |= PHW_PATH_LOOP_RES;
    } else {
        CU(cudaMemsetAsync(&s->ctl->gbar[0], 0, sizeof(unsigned int), s->stream));
        void* args[] = {(void*)&a};
        CU(cudaLaunchCooperativeKernel(loop_kernel(false, NT_LOOP), dim3(s->npatch), dim3(NT_LOOP), args, sm, s->stream));
    }
    s->launches++;
    s->kernel_paths |= PHW_PATH_LOOP;
    // the history ring and the first-same-as-last flag advance deterministically: one solution per (field, site) and solve
    auto advance = [&](int f, int site, int64_t n) {
        phw_solver_s::Hist& H = s->hist[f][site];
        if (!H.h[0] || n <= 0) return;
        const int order = s->extrap_f[f] + 1;
        H.head = (int)((H.head + n) % order);
        H.count = (int)std::min<int64_t>((int64_t)H.count + n, order);
    };
    if (n_steps > 0) {
        switch (P.scheme) {
            case PHW_SV: {
                const int64_t first = (!s->have_vq || (P.flags & 1)) ? ((P.flags & 1) ? n_steps : 1) : 0;
                advance(0, 0, n_steps); advance(1, 0, n_steps + first);
                s->have_vq = true;
            } break;
            case PHW_EH: advance(0, 0, n_steps); advance(1, 0, n_steps); advance(0, 1, n_steps); advance(1, 1, n_steps); break;
            default: advance(0, 0, n_steps); advance(1, 0, n_steps);
        }
    }
    return 0;
}

bool is_device_ptr(const void* p);
// p of the current state -> next free slot of the snapshot ring (user numbering), then to the caller's buffer on the snapshot
// stream while the time loop goes on; a slot is reused only after its previous copy has finished
int take_snapshot(phw_solver_s* s) {
    if (s->snap_taken >= s->snap_cap) return 0;
    const int slot = (int)(s->snap_taken % phw_solver_s::SNAP_SLOTS);
    double* ring = s->snap_ring + (size_t)slot * s->nv;
    if (s->snap_taken >= phw_solver_s::SNAP_SLOTS) CU(cudaStreamWaitEvent(s->stream, s->snap_done[slot], 0));
    if (s->d_v_new2old) scatter_k<<<nblk(s->nv), 256, 0, s->stream>>>(s->nv, ring, s->p, s->d_v_new2old);
    else CU(cudaMemcpyAsync(ring, s->p, (size_t)s->nv * 8, cudaMemcpyDeviceToDevice, s->stream));
    CU(cudaGetLastError());
    s->launches++;
    CU(cudaEventRecord(s->snap_ready[slot], s->stream));
    CU(cudaStreamWaitEvent(s->snap_stream, s->snap_ready[slot], 0));
    double* dst = s->snap_dst + (size_t)s->snap_taken * s->nv;
    CU(cudaMemcpyAsync(dst, ring, (size_t)s->nv * 8, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s->snap_stream));
    CU(cudaEventRecord(s->snap_done[slot], s->snap_stream));
    s->snap_taken++;
    return 0;
}

bool is_device_ptr(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// user-numbering vector (host or device) -> internal numbering device vector
int import_vec(phw_solver_s* s, const double* src, double* dst, int n, const int* d_new2old, double* stage) {
    const double* dsrc = src;
    if (!is_device_ptr(src)) {
        CU(cudaMemcpyAsync(stage, src, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        dsrc = stage;
    }
    gather_k<<<nblk(n), 256, 0, s->stream>>>(n, dst, dsrc, d_new2old);
    CU(cudaGetLastError());
    return 0;
}
int export_vec(phw_solver_s* s, const double* src, double* dst, int n, const int* d_new2old, double* stage) {
    if (is_device_ptr(dst)) {
        scatter_k<<<nblk(n), 256, 0, s->stream>>>(n, dst, src, d_new2old);
        CU(cudaGetLastError());
    } else {
        scatter_k<<<nblk(n), 256, 0, s->stream>>>(n, stage, src, d_new2old);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(dst, stage, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    return 0;
}

// minimise the Chebyshev convergence factor over ellipses through the corner (delta, mu) of the
// rectangle [alpha,beta] x i[-mu,mu] that contains the field of values of the Jacobi-scaled block operator
void ellipse_for_rectangle(double alpha, double beta, double mu, double& d, double& c) {
    d = 0.5 * (alpha + beta);
    const double delta = 0.5 * (beta - alpha);
    if (mu <= 0.0) { c = delta; return; }
    double best = 1e300, best_c = delta;
    for (int i = 1; i <= 4000; ++i) {
        // semi-axis a >= delta ; b from delta^2/a^2 + mu^2/b^2 = 1
        double a = delta * (1.0 + 4.0 * i / 4000.0 * (mu / delta + 0.05));
        double t = 1.0 - delta * delta / (a * a);
        if (t <= 0.0) continue;
        double b = mu / std::sqrt(t);
        if (a >= d) continue;
        double c2 = a * a - b * b;
        double cc = std::sqrt(std::fabs(c2));
        double rate;
        if (c2 >= 0.0) rate = (a + b) / (d + std::sqrt(d * d - c2));
        else continue;   // keep real foci
        if (rate < best) { best = rate; best_c = cc; }
    }
    c = best_c;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char* phw_last_error(void) { return g_err.c_str(); }
int phw_version(void) { return 200; }
int phw_sizeof(int what) { return what == 0 ? (int)sizeof(phw_params) : (what == 1 ? (int)sizeof(phw_stats) : -1); }

int phw_mesh_create(int64_t n_vertices, const double* vertices_xy, int64_t n_cells, const int32_t* cells,
                    int32_t patch_cells, phw_mesh_t* out) {
    if (!out) return fail(PHW_EINVAL, "out is NULL");
    phw_mesh_s* m = new (std::nothrow) phw_mesh_s();
    if (!m) return fail(PHW_ENOMEM, "out of host memory");
    std::string err;
    try {
        err = m->L.build(n_vertices, vertices_xy, n_cells, cells, patch_cells > 0 ? patch_cells : 2048);
    } catch (const std::bad_alloc&) {
        delete m;
        return fail(PHW_ENOMEM, "out of host memory while building the mesh layout");
    }
    if (!err.empty()) { delete m; return fail(PHW_EINVAL, err); }
    *out = m;
    return PHW_OK;
}

int phw_mesh_create_partitioned(int64_t n_vertices, const double* vertices_xy, int64_t n_cells, const int32_t* cells,
                                int32_t patch_cells, const uint8_t* vertex_external, int64_t n_edges,
                                const uint8_t* edge_external, phw_mesh_t* out) {
    if (!out || !vertex_external || !edge_external) return fail(PHW_EINVAL, "NULL argument");
    phw_mesh_s* m = new (std::nothrow) phw_mesh_s();
    if (!m) return fail(PHW_ENOMEM, "out of host memory");
    std::string err;
    try {
        err = m->L.build(n_vertices, vertices_xy, n_cells, cells, patch_cells > 0 ? patch_cells : 2048,
                         vertex_external, n_edges, edge_external);
    } catch (const std::bad_alloc&) {
        delete m;
        return fail(PHW_ENOMEM, "out of host memory while building the mesh layout");
    }
    if (!err.empty()) { delete m; return fail(PHW_EINVAL, err); }
    *out = m;
    return PHW_OK;
}

int phw_mesh_set_cell_owned(phw_mesh_t m, const uint8_t* cell_owned) {
    if (!m || !cell_owned) return fail(PHW_EINVAL, "mesh or cell_owned is NULL");
    m->L.cell_foreign.resize((size_t)m->L.nc);
    for (int64_t c = 0; c < m->L.nc; ++c) m->L.cell_foreign[c] = cell_owned[c] ? 0 : 1;
    return PHW_OK;
}

int phw_mesh_create_grouped(int64_t n_vertices, const double* vertices_xy, int64_t n_cells, const int32_t* cells,
                            int32_t patch_cells, const int32_t* cell_group, int32_t n_groups, phw_mesh_t* out) {
    if (!out || !cell_group || n_groups < 1) return fail(PHW_EINVAL, "bad arguments");
    phw_mesh_s* m = new (std::nothrow) phw_mesh_s();
    if (!m) return fail(PHW_ENOMEM, "out of host memory");
    std::string err;
    try {
        err = m->L.build(n_vertices, vertices_xy, n_cells, cells, patch_cells > 0 ? patch_cells : 2048, nullptr, -1, nullptr, cell_group, n_groups);
    } catch (const std::bad_alloc&) {
        delete m;
        return fail(PHW_ENOMEM, "out of host memory while building the mesh layout");
    }
    if (!err.empty()) { delete m; return fail(PHW_EINVAL, err); }
    *out = m;
    return PHW_OK;
}

int phw_mesh_sizes(phw_mesh_t m, int64_t* nv, int64_t* nc, int64_t* ne, int64_t* nb, int64_t* np) {
    if (!m) return fail(PHW_EINVAL, "mesh is NULL");
    if (nv) *nv = m->L.nv;
    if (nc) *nc = m->L.nc;
    if (ne) *ne = m->L.ne;
    if (nb) *nb = m->L.nb;
    if (np) *np = m->L.npatch;
    return PHW_OK;
}

int phw_mesh_get_edges(phw_mesh_t m, int32_t* edges, int32_t* cell_edges, int8_t* cell_edge_sign) {
    if (!m) return fail(PHW_EINVAL, "mesh is NULL");
    if (edges) memcpy(edges, m->L.edges.data(), m->L.edges.size() * sizeof(int32_t));
    if (cell_edges) memcpy(cell_edges, m->L.cell_edges.data(), m->L.cell_edges.size() * sizeof(int32_t));
    if (cell_edge_sign) memcpy(cell_edge_sign, m->L.cell_sign.data(), m->L.cell_sign.size());
    return PHW_OK;
}

int phw_mesh_get_boundary(phw_mesh_t m, int32_t* edge_ids, int8_t* sign) {
    if (!m) return fail(PHW_EINVAL, "mesh is NULL");
    if (edge_ids) memcpy(edge_ids, m->L.bedge.data(), m->L.bedge.size() * sizeof(int32_t));
    if (sign) memcpy(sign, m->L.bsign.data(), m->L.bsign.size());
    return PHW_OK;
}

int phw_mesh_set_boundary(phw_mesh_t m, const int32_t* tags, const double* port_weight) {
    if (!m || !tags) return fail(PHW_EINVAL, "mesh or tags is NULL");
    for (int64_t k = 0; k < m->L.nb; ++k) {
        if (tags[k] < 0 || tags[k] > 4) return fail(PHW_EINVAL, "boundary tag out of range [0,4]");
        m->L.btag[k] = tags[k];
        m->L.bweight[k] = port_weight ? port_weight[k] : 1.0;
    }
    m->L.boundary_set = true;
    return PHW_OK;
}

int phw_mesh_set_dirichlet(phw_mesh_t m, int64_t n, const int32_t* vertex_ids, const double* weights) {
    if (!m || (n > 0 && (!vertex_ids || !weights))) return fail(PHW_EINVAL, "NULL argument");
    for (int64_t k = 0; k < n; ++k)
        if (vertex_ids[k] < 0 || vertex_ids[k] >= m->L.nv) return fail(PHW_EINVAL, "Dirichlet vertex id out of range");
    m->L.dir_vertex.assign(vertex_ids, vertex_ids + n);
    m->L.dir_weight.assign(weights, weights + n);
    return PHW_OK;
}

int phw_mesh_get_numbering(phw_mesh_t m, int32_t* v_new2old, int32_t* e_new2old, int32_t* cell_patch) {
    if (!m) return fail(PHW_EINVAL, "mesh is NULL");
    if (v_new2old) memcpy(v_new2old, m->L.v_new2old.data(), m->L.v_new2old.size() * sizeof(int32_t));
    if (e_new2old) memcpy(e_new2old, m->L.e_new2old.data(), m->L.e_new2old.size() * sizeof(int32_t));
    if (cell_patch) memcpy(cell_patch, m->L.cell_patch.data(), m->L.cell_patch.size() * sizeof(int32_t));
    return PHW_OK;
}

int phw_mesh_get_patch_counts(phw_mesh_t m, int32_t* counts) {
    if (!m || !counts) return fail(PHW_EINVAL, "mesh or counts is NULL");
    for (int32_t p = 0; p < m->L.npatch; ++p) {
        const PatchHdr& h = m->L.hdr[p];
        int32_t* c = counts + 6 * (int64_t)p;
        c[0] = h.n_own; c[1] = h.n_ecells; c[2] = h.n_vcells; c[3] = h.v_cnt; c[4] = h.e_cnt; c[5] = h.gv_cnt + h.ge_cnt;
    }
    return PHW_OK;
}

int phw_mesh_pipe_smem(phw_mesh_t m, int64_t* bytes6) {
    if (!m || !bytes6) return fail(PHW_EINVAL, "mesh or bytes6 is NULL");
    MqPipeArgs cq{}; EllPipeArgs cp{}; EnergyPipeArgs ce{}; RhsPipeArgs re{}, rv{};
    size_t sq = 0, sp = 0, se = 0, sre = 0, srv = 0;
    pipe_carve(m->L, cq, cp, sq, sp, &ce, &se, &re, &sre, &rv, &srv);
    bytes6[0] = (int64_t)sp; bytes6[1] = (int64_t)sq; bytes6[2] = (int64_t)se; bytes6[3] = (int64_t)sre; bytes6[4] = (int64_t)srv;
    {   // whole-loop kernel with resident solves: scratch of the row passes (smem_vd of setup_solver) + the resident carve-up
        const Layout& L = m->L;
        auto ev = [](int x) { return (size_t)((x + 1) & ~1); };
        const size_t soff = ((size_t)(L.max_lv + 31) / 32 + 4) / 2 + 1;
        bytes6[5] = (int64_t)res_carve(L, (soff + ev(L.max_lv) + ev(L.max_le) + 3 * (size_t)L.max_cells) * 8).total;
    }
    return PHW_OK;
}
int phw_mesh_self_check(phw_mesh_t m) {
    if (!m) return fail(PHW_EINVAL, "mesh is NULL");
    if (m->L.nc > 200000) return fail(PHW_EINVAL, "phw_mesh_self_check is quadratic in places: meshes up to 200 000 cells");
    const std::string e = m->L.self_check();
    if (!e.empty()) return fail(PHW_ESTATE, "layout self check failed: " + e);
    return PHW_OK;
}

int phw_mesh_destroy(phw_mesh_t m) { delete m; return PHW_OK; }

int phw_solver_destroy(phw_solver_t s) {
    if (!s) return PHW_OK;
    cudaSetDevice(s->device);
    for (size_t r = 0; r < s->peer_arenas.size(); ++r)
        if (s->peer_arenas[r] && s->peer_arenas[r] != s->arena) cudaIpcCloseMemHandle(s->peer_arenas[r]);
    for (void* d : s->allocs) cudaFree(d);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->snap_registered) cudaHostUnregister(s->snap_registered);
    if (s->snap_stream) {
        cudaStreamDestroy(s->snap_stream);
        for (int i = 0; i < phw_solver_s::SNAP_SLOTS; ++i) { cudaEventDestroy(s->snap_ready[i]); cudaEventDestroy(s->snap_done[i]); }
    }
    delete s;
    return PHW_OK;
}

int phw_stream_probe(phw_solver_t s, int32_t threads, int32_t ctas_per_sm, int32_t reps, double* ms_per_pass, double* bytes_per_pass) {
    if (!s || !ms_per_pass || !bytes_per_pass || reps < 1) return fail(PHW_EINVAL, "phw_stream_probe: bad arguments");
    if (s->ho) return fail(PHW_ESTATE, "not available for higher-order solvers");
    CU(cudaSetDevice(s->device));
    const int grid = std::max(1, std::min(s->npatch, ctas_per_sm * s->sm_count));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    for (int r = 0; r < reps + 1; ++r) {
        if (r == 1) CU(cudaEventRecord(e0, s->stream));
        if (threads == 128) stream_probe_k<128><<<grid, 128, 0, s->stream>>>(s->dm, s->bp, s->tp.b[1], s->tp.b[0], s->tp.b[2]);
        else if (threads == 512) stream_probe_k<512><<<grid, 512, 0, s->stream>>>(s->dm, s->bp, s->tp.b[1], s->tp.b[0], s->tp.b[2]);
        else if (threads == 1024) stream_probe_k<1024><<<grid, 1024, 0, s->stream>>>(s->dm, s->bp, s->tp.b[1], s->tp.b[0], s->tp.b[2]);
        else stream_probe_k<256><<<grid, 256, 0, s->stream>>>(s->dm, s->bp, s->tp.b[1], s->tp.b[0], s->tp.b[2]);
    }
    CU(cudaEventRecord(e1, s->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const Layout& L = s->mesh->L;
    *ms_per_pass = ms / reps;
    *bytes_per_pass = 2.0 * (double)L.ell_ids.size() + 8.0 * (double)L.ell_pair.size() + 4.0 * 8.0 * (double)L.nv_owned;
    return PHW_OK;
}

int phw_residual_history(phw_solver_t s, int32_t which, int32_t n, double* rel, int32_t* n_it) {
    if (!s || which < 0 || which > 2 || n < 0 || n > 48 || !rel) return fail(PHW_EINVAL, "phw_residual_history: bad arguments");
    CU(cudaStreamSynchronize(s->stream));
    Ctl h;
    CU(cudaMemcpy(&h, s->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
    for (int k = 0; k < n; ++k) rel[k] = std::sqrt(h.reshist[which][k]);
    if (n_it) *n_it = h.sc[which].k;
    return 0;
}
int phw_synchronize(phw_solver_t s) {
    if (!s) return fail(PHW_EINVAL, "solver is NULL");
    CU(cudaStreamSynchronize(s->stream));
    return PHW_OK;
}

static bool implicit_scheme(int scheme) { return scheme == PHW_IE || scheme == PHW_SM; }
static void fill_physics(ScalarPrm& sp, double dt, double K, double rho, double Bl, double R1, double R2, double K_em, double m_em, double L_em) {
    sp.dt = dt; sp.K = K; sp.rho = rho; sp.c2 = K / rho;
    // A_em = (J - R) diag(1/L, 1/m, K_em)      wave_2D.py:160-166
    const double J[3][3] = {{0, Bl, 0}, {-Bl, 0, -1}, {0, 1, 0}};
    const double R[3] = {R1, R2, 0.0};
    const double cv[3] = {L_em != 0 ? 1.0 / L_em : 0.0, m_em != 0 ? 1.0 / m_em : 0.0, K_em};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) sp.A[i][j] = (J[i][j] - (i == j ? R[i] : 0.0)) * cv[j];
    sp.m_em = m_em; sp.L_em = L_em; sp.K_em = K_em;
}

// state / work vectors, exchange arena, control blocks (shared by the matrix-free and the CSR path)
static int setup_vectors(phw_solver_s* s) {
    const phw_params& P = s->prm;
    int rc;
#define RC(x) do { rc = (x); if (rc) return rc; } while (0)
    // ---- vectors
    const size_t nv = s->nv, ne = s->ne;
    RC(s->dalloc(&s->p, nv)); RC(s->dalloc(&s->q, ne)); RC(s->dalloc(&s->bp, nv)); RC(s->dalloc(&s->bq, ne));
    RC(s->dalloc(&s->tmp_p, nv)); RC(s->dalloc(&s->tmp_q, ne)); RC(s->dalloc(&s->pinv_p, nv)); RC(s->dalloc(&s->pinv_q, ne)); RC(s->dalloc(&s->mdiag_p, nv));
    RC(s->dalloc(&s->io_v, nv)); RC(s->dalloc(&s->io_e, ne)); RC(s->dalloc(&s->io_v2, nv)); RC(s->dalloc(&s->io_e2, ne));
    {   // exchange arena: [Mailbox | tp0 tp1 tp2 | tq0 tq1 tq2] in ONE allocation so that a single CUDA IPC handle
        // lets the neighbouring ranks write halo values and partial sums straight into it (peer memory over NVLink)
        const size_t head = 4096, nvp = (nv + 31) & ~(size_t)31, nep = (ne + 31) & ~(size_t)31;
        s->arena_bytes = head + 3 * (nvp + nep) * sizeof(double);
        char* base = nullptr;
        RC(s->dalloc(&base, s->arena_bytes));
        s->arena = base;
        CU(cudaMemsetAsync(base, 0, s->arena_bytes, s->stream));
        double* d0 = (double*)(base + head);
        for (int i = 0; i < 3; ++i) { s->tp.b[i] = d0 + i * nvp; s->tq.b[i] = d0 + 3 * nvp + i * nep; }
        s->tp.b[3] = nullptr; s->tq.b[3] = nullptr;
        s->pc = PeerComm{};
        s->pc.rank = 0; s->pc.world = 1; s->pc.n_nb = 0;
        s->pc.mbox[0] = (Mailbox*)base;
        s->pc.timeout_cycles = 20000000000LL;
    }
    RC(s->dalloc(&s->partials, 2 * (size_t)s->npatch + 16384));
    {
        int coop_attr = 0;
        CU(cudaDeviceGetAttribute(&coop_attr, cudaDevAttrCooperativeLaunch, s->device));
        s->coop = coop_attr != 0 && !(P.flags & 8);
    }
    RC(s->dalloc(&s->ctl, (size_t)s->G));
    // initial guesses: extrapolation order of the p solves (extrap_order) and of the q solves (extrap_order_q, 0 = the same)
    s->extrap = (P.extrap_order < 0) ? 0 : std::min(P.extrap_order, 5);
    s->extrap_f[0] = s->extrap;
    s->extrap_f[1] = (P.extrap_order_q > 0) ? std::min(P.extrap_order_q, 5) : s->extrap;
    s->n_sites = (P.scheme == PHW_EH) ? 2 : 1;
    for (int site = 0; site < s->n_sites; ++site) {
        for (int j = 0; j <= s->extrap_f[0]; ++j) RC(s->dalloc(&s->hist[0][site].h[j], nv));
        for (int j = 0; j <= s->extrap_f[1]; ++j) RC(s->dalloc(&s->hist[1][site].h[j], ne));
    }
    for (double* v : {s->p, s->bp, s->tmp_p, s->tp.b[0], s->tp.b[1], s->tp.b[2], s->io_v, s->io_v2}) CU(cudaMemsetAsync(v, 0, nv * 8, s->stream));
    for (double* v : {s->q, s->bq, s->tmp_q, s->tq.b[0], s->tq.b[1], s->tq.b[2], s->io_e, s->io_e2}) CU(cudaMemsetAsync(v, 0, ne * 8, s->stream));
    CU(cudaMemsetAsync(s->ctl, 0, sizeof(Ctl) * (size_t)s->G, s->stream));
    return 0;
#undef RC
}
static int setup_scalars(phw_solver_s* s) {
    const phw_params& P = s->prm;
    int rc;
#define RC(x) do { rc = (x); if (rc) return rc; } while (0)
    // ---- scalar parameters
    ScalarPrm& sp = s->sp;
    fill_physics(sp, P.dt, P.K_wave, P.rho, P.Bl_em, P.R1_em, P.R2_em, P.K_em, P.m_em, P.L_em);
    sp.input_kind = P.input_kind; sp.in_amp = P.in_amp; sp.in_omega = P.in_omega; sp.in_t_stop = P.in_t_stop;
    sp.in_t_ctrl = P.in_t_ctrl; sp.in_gain = P.in_gain; sp.H_init = P.H_init;
    s->ncol = P.analytical ? 12 : 8;
    sp.ncol = s->ncol; sp.ic = P.interconnection;
    switch (P.scheme) {   // q1_ of wave_2D.py:603-615
        case PHW_EE: case PHW_SE: sp.w_old = 1.0; sp.w_new = 0.0; break;
        case PHW_IE: sp.w_old = 0.0; sp.w_new = 1.0; break;
        case PHW_SM: sp.w_old = 0.5; sp.w_new = 0.5; break;
        default: sp.w_old = 1.0; sp.w_new = 0.0;
    }
    s->h_sp.assign(s->G, sp);
    RC(s->upload(&s->d_sp, s->h_sp));
    return 0;
#undef RC
}

static int setup_solver(phw_solver_s* s) {
    const Layout& L = s->mesh->L;
    const phw_params& P = s->prm;
    s->nv = (int)L.nv; s->ne = (int)L.ne; s->npatch = L.npatch_rows; s->npc = (int)L.pcell_user.size();
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, s->device));
    s->sm_count = prop.multiProcessorCount;
    CU(cudaEventCreate(&s->ev0));
    CU(cudaEventCreate(&s->ev1));
    int rc;
#define RC(x) do { rc = (x); if (rc) return rc; } while (0)
    // ---- mesh SoA
    std::vector<CellRec> rec;
    L.make_records(P.strong != 0, P.casimir != 0, rec);
    PatchHdr* d_hdr; uint2 *d_recv, *d_rece; int *d_gv, *d_ge; uint32_t *d_eadj, *d_voff; uint16_t* d_vadj;
    RC(s->upload(&d_hdr, L.hdr));
    {
        std::vector<uint2> hv(rec.size()), he(rec.size());
        for (size_t k = 0; k < rec.size(); ++k) { hv[k] = make_uint2(rec[k].x, rec[k].y); he[k] = make_uint2(rec[k].z, rec[k].w); }
        RC(s->upload(&d_recv, hv)); RC(s->upload(&d_rece, he));
        CU(cudaStreamSynchronize(s->stream));
    }
    RC(s->upload(&d_gv, L.ghost_v)); RC(s->upload(&d_ge, L.ghost_e));
    RC(s->upload(&d_eadj, L.eadj)); RC(s->upload(&d_voff, L.vslice_off)); RC(s->upload(&d_vadj, L.vadj));
    RC(s->upload(&s->d_v_new2old, L.v_new2old)); RC(s->upload(&s->d_e_new2old, L.e_new2old));
    double *d_area, *d_g0, *d_g1, *d_g2;
    RC(s->dalloc(&d_area, s->npc)); RC(s->dalloc(&d_g0, s->npc)); RC(s->dalloc(&d_g1, s->npc)); RC(s->dalloc(&d_g2, s->npc));
    {   // one-time GPU geometry pass (replaces the per-cell Jacobians FFC kernels recompute at every assemble)
        int *d_pcu, *d_cells; double* d_vtx; unsigned long long* d_mm;
        RC(s->upload(&d_pcu, L.pcell_user)); RC(s->upload(&d_cells, L.cells)); RC(s->upload(&d_vtx, L.vtx));
        RC(s->dalloc(&d_mm, 2));
        s->d_cells_user = d_cells; s->d_vtx_user = d_vtx; s->d_mm = d_mm;
        unsigned long long init[2];
        double one = 1.0;
        memcpy(&init[0], &one, 8); memcpy(&init[1], &one, 8);
        CU(cudaMemcpyAsync(d_mm, init, sizeof init, cudaMemcpyHostToDevice, s->stream));
        geom_k<<<nblk(s->npc), 256, 0, s->stream>>>(s->npc, d_pcu, d_cells, d_vtx, d_area, d_g0, d_g1, d_g2, d_mm);
        CU(cudaGetLastError());
        unsigned long long out[2];
        CU(cudaMemcpyAsync(out, d_mm, sizeof out, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        memcpy(&s->lam_min_q, &out[0], 8); memcpy(&s->lam_max_q, &out[1], 8);
    }
    s->dm.npatch = s->npatch; s->dm.nv = s->nv; s->dm.ne = s->ne;
    s->dm.hdr = d_hdr; s->dm.recv = d_recv; s->dm.rece = d_rece; s->dm.area = d_area; s->dm.g0 = d_g0; s->dm.g1 = d_g1; s->dm.g2 = d_g2;
    s->dm.ghost_v = d_gv; s->dm.ghost_e = d_ge; s->dm.eadj = d_eadj; s->dm.vslice_off = d_voff; s->dm.vadj = d_vadj;
    {   // packed cell records of the lean M_q solve
        RC(s->dalloc(&s->d_cellq, 2 * (size_t)s->npc));
        pack_cellq_k<<<nblk(s->npc), 256, 0, s->stream>>>(s->npc, d_rece, d_g0, d_g1, d_g2, s->d_cellq);
        CU(cudaGetLastError());
    }
    {   // assembled P1 mass rows (ELL): structure from the layout, weights from the device areas
        uint32_t *d_io, *d_wo, *d_pair; uint16_t* d_ids; double* d_w;
        RC(s->upload(&d_io, L.ell_ioff)); RC(s->upload(&d_wo, L.ell_woff)); RC(s->upload(&d_ids, L.ell_ids)); RC(s->upload(&d_pair, L.ell_pair));
        RC(s->dalloc(&d_w, L.ell_pair.size()));
        ell_weights_k<<<std::max(1, std::min(s->npatch, 8 * s->sm_count)), 256, 0, s->stream>>>(s->npatch, d_hdr, d_wo, d_pair, d_area, d_w);
        CU(cudaGetLastError());
        s->dm.ell_ioff = d_io; s->dm.ell_woff = d_wo; s->dm.ell_ids = d_ids; s->dm.ell_w = d_w;
        CU(cudaStreamSynchronize(s->stream));
    }
    // ---- port (left boundary, tag 0): b_L[e] = sign * weight   (flux-normalised RT1: <phi_e.n, g_L> = sign * mean(g_L))
    {
        s->G = L.n_groups;
        std::vector<int32_t> egroup;
        if (s->G > 1) {
            egroup.assign(L.ne, 0);
            for (int64_t h = 0; h < 3 * L.nc; ++h) egroup[L.cell_edges[h]] = L.cell_group_v[h / 3];
        }
        std::vector<int32_t> pe, pg, goff(s->G + 1, 0); std::vector<double> bl;
        std::vector<int64_t> ks;
        for (int64_t k = 0; k < L.nb; ++k)
            if (L.btag[k] == PHW_TAG_LEFT) { ks.push_back(k); goff[(s->G > 1 ? egroup[L.bedge[k]] : 0) + 1]++; }
        for (int g = 0; g < s->G; ++g) goff[g + 1] += goff[g];
        pe.resize(ks.size()); pg.resize(ks.size()); bl.resize(ks.size());
        std::vector<int32_t> pos(goff.begin(), goff.end() - 1);
        for (int64_t k : ks) {
            const int g = s->G > 1 ? egroup[L.bedge[k]] : 0;
            const int d = pos[g]++;
            pe[d] = L.e_old2new[L.bedge[k]]; pg[d] = g; bl[d] = L.bsign[k] * L.bweight[k];
        }
        s->nport = (int)pe.size();
        RC(s->upload(&s->d_port_e, pe)); RC(s->upload(&s->d_port_bl, bl));
        RC(s->upload(&s->d_port_goff, goff)); RC(s->upload(&s->d_port_group, pg));
        RC(s->upload(&s->d_gpoff, L.group_patch_off));
        if (s->G > 1) RC(s->dalloc(&s->wpart, (size_t)L.npatch * (NT_ROWS / 32)));
    }
    RC(setup_vectors(s));
    const size_t nv = s->nv, ne = s->ne;
    (void)nv; (void)ne;
    // ---- shared-memory budgets and grid
    auto ev = [](int x) { return (size_t)((x + 1) & ~1); };
    const size_t soff = ((size_t)(L.max_lv + 31) / 32 + 4) / 2 + 1;   // slice offsets staged by the vertex rows (uint32, in doubles)
    s->smem_v = (soff + ev(L.max_lv) + 3 * (size_t)L.max_cells) * 8;
    s->smem_v1 = (soff + ev(L.max_lv) + (size_t)L.max_cells) * 8;
    s->smem_vd = (soff + ev(L.max_lv) + ev(L.max_le) + 3 * (size_t)L.max_cells) * 8;
    s->smem_e = (ev(L.max_le) + 3 * (size_t)L.max_cells) * 8;
    s->smem_ed = s->smem_vd;
    s->smem_ell = ((size_t)(L.max_lv + 31) / 32 + 4 + ev(L.max_lv)) * 8;
    s->ell_p = !(P.flags & 32);
    s->smem_ells = (size_t)L.max_smem_ell * 8;
    s->ell_staged = (P.flags & 64) != 0;
    if (s->smem_vd > 200 * 1024) return fail(PHW_EINVAL, "patch does not fit in shared memory; reduce patch_cells");
    // ---- bulk-copy pipelined mass solves (flags bits 8 / 9): two stages holding every stream of the largest patch
    if (P.flags & (256 | 512 | 2048 | 4096)) {
        pipe_carve(L, s->pq, s->pp, s->smem_pq, s->smem_pp, &s->pe, &s->smem_pe, &s->pre, &s->smem_pre, &s->prv, &s->smem_prv);
        // partitioned runs: rows exist for owned dofs only and external dofs are read like any ghost, so the row kernels need nothing
        // extra; the cell-based energy pass needs to know which cells of the local mesh are this rank's (phw_mesh_set_cell_owned)
        const bool partitioned = L.nv_owned != L.nv || L.ne_owned != L.ne;
        if ((P.flags & 4096) && L.n_groups == 1 && !P.casimir) {
            if (std::max(s->smem_pre, s->smem_prv) > PIPE_SMEM_LIMIT) return fail(PHW_EINVAL, "pipelined right-hand sides (flags bit 12): two stages of the largest patch exceed the shared memory of an SM; reduce patch_cells");
            s->pipe_rhs = true;
        }
        if ((P.flags & 2048) && L.n_groups == 1 && (!partitioned || !L.cell_foreign.empty())) {
            if (s->smem_pe > PIPE_SMEM_LIMIT) return fail(PHW_EINVAL, "pipelined energy reduction (flags bit 11): two stages of the largest patch exceed the shared memory of an SM; reduce patch_cells");
            s->pipe_en = true;
        }
    }
    if (s->coop && (P.flags & (256 | 512))) {
        if (P.flags & 256) {
            if (s->smem_pq > PIPE_SMEM_LIMIT) return fail(PHW_EINVAL, "pipelined M_q solve (flags bit 8): two stages of the largest patch exceed the shared memory of an SM; use patch_cells <= 1024");
            s->pipe_q = true;
        }
        if ((P.flags & 512) && s->ell_p) {
            if (s->smem_pp > PIPE_SMEM_LIMIT) return fail(PHW_EINVAL, "pipelined M_p solve (flags bit 9): two stages of the largest patch exceed the shared memory of an SM; reduce patch_cells");
            s->pipe_p = true;
        }
    }
    s->grid_rows = std::max(1, std::min(s->npatch, PHW_MINB_ROWS * s->sm_count));
    RC(setup_scalars(s));
    if (s->G > 1) {
        if (implicit_scheme(P.scheme) || P.strong || P.analytical || P.casimir)
            return fail(PHW_EINVAL, "batched (grouped) runs support the explicit weak-form schemes only");
        std::vector<double> gk(s->G, P.K_wave), gir(s->G, 1.0 / P.rho);
        RC(s->upload(&s->d_gK, gk)); RC(s->upload(&s->d_ginvrho, gir));
    }
    // ---- Jacobi diagonals (inverse) of the solved blocks
    const bool implicit = (P.scheme == PHW_IE || P.scheme == PHW_SM);
    const double theta = (P.scheme == PHW_IE) ? 1.0 : 0.5;
    {
        RowArgs a = base_args(s);
        a.cM = 1.0; a.cB = 0.0; a.y = s->mdiag_p;
        RC((launch_vertex<EPI_STORE, MODE_DIAG, true, false, false>(s, a)));
        a.y = s->pinv_p;
        a.cB = (implicit && P.casimir) ? theta * P.dt * P.K_wave : 0.0;
        RC((a.cB != 0.0) ? (launch_vertex<EPI_STORE, MODE_DIAG, true, false, true>(s, a)) : (launch_vertex<EPI_STORE, MODE_DIAG, true, false, false>(s, a)));
        a.y = s->pinv_q;
        a.cB = (implicit && P.casimir) ? theta * P.dt / P.rho : 0.0;
        RC((a.cB != 0.0) ? (launch_edge<EPI_STORE, MODE_DIAG, true, false, true>(s, a)) : (launch_edge<EPI_STORE, MODE_DIAG, true, false, false>(s, a)));
        invert_k<<<nblk(s->nv), 256, 0, s->stream>>>(s->nv, s->pinv_p);
        invert_k<<<nblk(s->ne), 256, 0, s->stream>>>(s->ne, s->pinv_q);
        CU(cudaGetLastError());
    }
    if (P.strong) {
        std::vector<int32_t> di; std::vector<double> dw;
        for (size_t k = 0; k < L.dir_vertex.size(); ++k) { di.push_back(L.v_old2new[L.dir_vertex[k]]); dw.push_back(L.dir_weight[k]); }
        s->n_dir = (int)di.size();
        RC(s->upload(&s->d_dir_idx, di)); RC(s->upload(&s->d_dir_w, dw));
        if (s->n_dir > 0) {
            zero_at_k<<<nblk(s->n_dir), 256, 0, s->stream>>>(s->n_dir, s->d_dir_idx, s->pinv_p);
            CU(cudaGetLastError());
        }
    }
    // ---- Chebyshev intervals.  P1: element bound [1/2, 2] (Wathen); RT1: element bounds from geom_k.
    s->cheb_d[0] = 1.25; s->cheb_c[0] = 0.75;
    s->cheb_d[1] = 0.5 * (s->lam_min_q + s->lam_max_q); s->cheb_c[1] = 0.5 * (s->lam_max_q - s->lam_min_q);
    // meshes of poor quality: the element-wise interval of the RT1 mass matrix is pessimistic (a few stretched cells set it for the
    // whole mesh) - conjugate gradients converge on the true spectrum
    {
        static const int env = [] { const char* e = getenv("PHW_PCG"); return e ? atoi(e) : -1; }();
        const bool can = s->coop && s->G == 1 && !P.strong && !(P.flags & 131072) && env != 0;
        const bool force = (P.flags & 65536) || env == 1;
        s->pcg[0] = can && force;
        s->pcg[1] = can && (force || s->lam_max_q > 10.0 * s->lam_min_q);
    }
#undef RC
    return 0;
}

// largest singular value of P_q^-1/2 D P_p^-1/2 by power iteration on the device (setup only)
static int estimate_skew(phw_solver_s* s, double* sigma) {
    const int nv = s->nv, ne = s->ne;
    std::vector<double> h(nv), dp(nv), dq(ne), y(ne);
    std::vector<double> hp(nv), hq(ne);
    CU(cudaMemcpyAsync(dp.data(), s->pinv_p, nv * 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(dq.data(), s->pinv_q, ne * 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    unsigned long long seed = 88172645463325252ull;
    for (int i = 0; i < nv; ++i) { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; h[i] = (double)(seed % 2000001) / 1000000.0 - 1.0; }
    double lam = 0.0;
    int rc;
    for (int it = 0; it < 40; ++it) {
        // z = P_p^-1/2 h ; y = D z ; y <- P_q^-1 y ; w = D^T y ; h' = P_p^-1/2 w
        for (int i = 0; i < nv; ++i) hp[i] = h[i] * std::sqrt(dp[i]);
        CU(cudaMemcpyAsync(s->io_v, hp.data(), nv * 8, cudaMemcpyHostToDevice, s->stream));
        if (s->ho_mf) rc = ho_mf_D(s, s->io_v, 1.0, s->io_e);
        else if (s->ho) { csr_store(s, s->hD, s->io_v, 1.0, s->io_e); rc = 0; }
        else rc = op_edge_store(s, s->io_v, nullptr, 0.0, 1.0, 0.0, s->io_e);
        if (rc) return rc;
        CU(cudaMemcpyAsync(y.data(), s->io_e, ne * 8, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        for (int i = 0; i < ne; ++i) hq[i] = y[i] * dq[i];
        CU(cudaMemcpyAsync(s->io_e, hq.data(), ne * 8, cudaMemcpyHostToDevice, s->stream));
        if (s->ho_mf) rc = ho_mf_DT(s, s->io_e, 1.0, s->io_v);
        else if (s->ho) { csr_store(s, s->hDT, s->io_e, 1.0, s->io_v); rc = 0; }
        else rc = op_vertex_store(s, nullptr, s->io_e, 0.0, 1.0, 0.0, s->io_v);
        if (rc) return rc;
        CU(cudaMemcpyAsync(hp.data(), s->io_v, nv * 8, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        double nrm = 0.0;
        for (int i = 0; i < nv; ++i) { hp[i] *= std::sqrt(dp[i]); nrm += hp[i] * hp[i]; }
        nrm = std::sqrt(nrm);
        if (nrm == 0.0) break;
        double hn = 0.0;
        for (int i = 0; i < nv; ++i) hn += h[i] * h[i];
        lam = nrm / std::sqrt(hn);
        for (int i = 0; i < nv; ++i) h[i] = hp[i] / nrm;
    }
    *sigma = std::sqrt(lam);
    return 0;
}

// implicit schemes (IE, SM): ellipse of the Chebyshev iteration for the coupled block system from the spectral intervals of the two
// Jacobi-scaled mass matrices and a power-iteration bound of the skew part; with the lumped model the constant column
// w = B^-1 [0; b_L] of the bordered system (wave_2D.py:562-572), solved once.  bl: b_L in internal edge numbering (empty: no lumped model)
static int setup_implicit(phw_solver_s* s, const std::vector<double>& bl, double sigma_given = -1.0) {
    const phw_params* prm = &s->prm;
    const double theta = (prm->scheme == PHW_IE) ? 1.0 : 0.5;
    double sigma = sigma_given;
    int rc = 0;
    if (!(sigma >= 0.0)) { rc = estimate_skew(s, &sigma); if (rc) return rc; }     // partitioned runs: the caller's distributed estimate
    s->skew_bound = 1.05 * theta * prm->dt * std::sqrt(prm->K_wave / prm->rho) * sigma;
    const double lo = std::min(s->cheb_d[0] - s->cheb_c[0], s->lam_min_q), hi = std::max(s->cheb_d[0] + s->cheb_c[0], s->lam_max_q);
    ellipse_for_rectangle(lo, hi, s->skew_bound, s->cheb_d[2], s->cheb_c[2]);
    if (!prm->interconnection) return 0;
    rc = s->dalloc(&s->wp, s->nv); if (!rc) rc = s->dalloc(&s->wq, s->ne);
    if (rc) return rc;
    CU(cudaMemcpyAsync(s->bq, bl.data(), (size_t)s->ne * 8, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemsetAsync(s->bp, 0, (size_t)s->nv * 8, s->stream));
    const double tol_keep = s->prm.tol;
    s->prm.tol = std::min(tol_keep, 1e-14);
    rc = solve_block(s, theta, s->prm.max_iter);
    s->prm.tol = tol_keep;
    if (rc) return rc;
    copy_from_tb_k<<<vec_grid(s, s->nv), 256, 0, s->stream>>>(s->nv, s->wp, s->tp, s->ctl, 2);
    copy_from_tb_k<<<vec_grid(s, s->ne), 256, 0, s->stream>>>(s->ne, s->wq, s->tq, s->ctl, 2);
    rc = flux(s, s->wq, 2, 3);                      // b_L^T w_q, summed over the ranks in partitioned runs
    if (rc) return rc;
    double f3 = 0.0;
    unsigned long long seq = 0;
    cudaMemcpyAsync(&f3, &s->ctl->flux[3], 8, cudaMemcpyDeviceToHost, s->stream);
    cudaMemcpyAsync(&seq, &s->ctl->comm_seq, sizeof(seq), cudaMemcpyDeviceToHost, s->stream);
    cudaStreamSynchronize(s->stream);
    s->blw = f3;
    // reset warm starts and counters (the mailbox sequence number of the inter-GPU all-reduce keeps counting)
    for (int i = 0; i < 3; ++i) { cudaMemsetAsync(s->tp.b[i], 0, (size_t)s->nv * 8, s->stream); cudaMemsetAsync(s->tq.b[i], 0, (size_t)s->ne * 8, s->stream); }
    cudaMemsetAsync(s->ctl, 0, sizeof(Ctl), s->stream);
    cudaMemcpyAsync(&s->ctl->comm_seq, &seq, sizeof(seq), cudaMemcpyHostToDevice, s->stream);
    cudaStreamSynchronize(s->stream);
    if (cudaGetLastError() != cudaSuccess) return fail(PHW_ECUDA, "bordered-column setup solve failed");
    return 0;
}

// A/B knob: PHW_FLAGS_OR=<int> is OR-ed into phw_params.flags of every solver (e.g. run the whole test suite with a kernel variant)
static int env_flags() {
    static const int f = [] { const char* e = getenv("PHW_FLAGS_OR"); return e ? atoi(e) : 0; }();
    return f;
}

int phw_solver_create(phw_mesh_t m, const phw_params* prm, int32_t device, void* cuda_stream, phw_solver_t* out) {
    if (!m || !prm || !out) return fail(PHW_EINVAL, "NULL argument");
    if (!m->L.boundary_set) return fail(PHW_ESTATE, "phw_mesh_set_boundary must be called before phw_solver_create");
    if (prm->scheme < 0 || prm->scheme > 5) return fail(PHW_EINVAL, "unknown scheme");
    if (!(prm->dt > 0) || !(prm->K_wave > 0) || !(prm->rho > 0)) return fail(PHW_EINVAL, "dt, K_wave, rho must be positive");
    if (prm->interconnection && !(prm->scheme == PHW_SE || prm->scheme == PHW_SV || prm->scheme == PHW_SM))
        return fail(PHW_EINVAL, "only SE, SV and SM time int schemes implemented for interconnection model");
    if (prm->interconnection && prm->strong) return fail(PHW_EINVAL, "can only do interconnection model with weak boundary conditions");
    if (prm->casimir && prm->scheme != PHW_SM) return fail(PHW_EINVAL, "casimir control requires the SM scheme");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(PHW_ECUDA, "no CUDA device available (libphwave has no CPU fallback)"); }
    if (device < 0 || device >= ndev) return fail(PHW_EINVAL, "device index out of range");
    CU(cudaSetDevice(device));
    phw_solver_s* s = new (std::nothrow) phw_solver_s();
    if (!s) return fail(PHW_ENOMEM, "out of host memory");
    s->mesh = m; s->prm = *prm; s->device = device; s->stream = (cudaStream_t)cuda_stream;
    s->prm.flags |= env_flags();
    if (!(s->prm.tol > 0)) s->prm.tol = 1e-13;
    if (s->prm.max_iter <= 0) s->prm.max_iter = 200;
    int rc = setup_solver(s);
    if (rc) { std::string keep = g_err; phw_solver_destroy(s); g_err = keep; return rc; }
    if (prm->scheme == PHW_IE || prm->scheme == PHW_SM) {
        std::vector<double> bl;
        if (prm->interconnection) {
            bl.assign(s->ne, 0.0);
            const Layout& L = m->L;
            for (int64_t k = 0; k < L.nb; ++k)
                if (L.btag[k] == PHW_TAG_LEFT) bl[L.e_old2new[L.bedge[k]]] = L.bsign[k] * L.bweight[k];
        }
        const Layout& L0 = m->L;
        if (L0.nv_owned != L0.nv || L0.ne_owned != L0.ne) {
            s->implicit_pending = true;      // the skew bound and the bordered column need the exchange: phw_comm_finish_implicit
        } else {
            rc = setup_implicit(s, bl);
            if (rc) { std::string keep = g_err; phw_solver_destroy(s); g_err = keep; return rc; }
        }
    }
    rc = setup_loop(s);
    if (rc) { std::string keep = g_err; phw_solver_destroy(s); g_err = keep; return rc; }
    CU(cudaStreamSynchronize(s->stream));
    s->launches = 0;
    *out = s;
    return PHW_OK;
}

// ---- multi-GPU plumbing (one process per GPU; handles travel through the caller's torch.distributed) ----
int phw_comm_arena_handle(phw_solver_t s, void* handle64, int64_t* n_vertices_local, int64_t* n_edges_local) {
    if (!s || !handle64) return fail(PHW_EINVAL, "NULL argument");
    CU(cudaSetDevice(s->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->arena));
    memcpy(handle64, &h, 64);
    if (n_vertices_local) *n_vertices_local = s->nv;
    if (n_edges_local) *n_edges_local = s->ne;
    return PHW_OK;
}

int phw_comm_internal_index(phw_solver_t s, int32_t which, int64_t n, const int32_t* user_idx, int32_t* internal_idx) {
    if (!s || (n > 0 && (!user_idx || !internal_idx))) return fail(PHW_EINVAL, "NULL argument");
    if (s->ho) return fail(PHW_ESTATE, "not available for higher-order solvers");
    const std::vector<int32_t>& o2n = (which == PHW_P) ? s->mesh->L.v_old2new : s->mesh->L.e_old2new;
    for (int64_t i = 0; i < n; ++i) {
        if (user_idx[i] < 0 || (size_t)user_idx[i] >= o2n.size()) return fail(PHW_EINVAL, "index out of range");
        internal_idx[i] = o2n[user_idx[i]];
    }
    return PHW_OK;
}

int phw_comm_setup(phw_solver_t s, int32_t rank, int32_t world, const void* handles64, const int64_t* peer_nv,
                   const int64_t* peer_ne, int32_t n_nb, const int32_t* nb_rank,
                   const int64_t* n_send_v, const int32_t* const* send_v_user, const int32_t* const* send_v_remote,
                   const int64_t* n_send_e, const int32_t* const* send_e_user, const int32_t* const* send_e_remote) {
    if (!s || !handles64 || !peer_nv || !peer_ne) return fail(PHW_EINVAL, "NULL argument");
    if (world < 1 || world > MAX_RANKS || rank < 0 || rank >= world || n_nb < 0 || n_nb >= MAX_RANKS)
        return fail(PHW_EINVAL, "rank/world/neighbour count out of range (at most 8 ranks)");
    if (s->ho) return fail(PHW_EINVAL, "partitioned runs are not available for higher-order solvers");
    if (!s->coop) return fail(PHW_ESTATE, "multi-GPU runs need the cooperative solve path");
    CU(cudaSetDevice(s->device));
    PeerComm pc = s->pc;
    pc.rank = rank; pc.world = world; pc.n_nb = n_nb;
    s->peer_arenas.assign(world, nullptr);
    const size_t head = 4096;
    for (int r = 0; r < world; ++r) {
        if (r == rank) { s->peer_arenas[r] = s->arena; }
        else {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char*)handles64 + 64 * (size_t)r, 64);
            void* ptr = nullptr;
            CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            s->peer_arenas[r] = ptr;
        }
        pc.mbox[r] = (Mailbox*)s->peer_arenas[r];
    }
    const Layout& L = s->mesh->L;
    for (int k = 0; k < n_nb; ++k) {
        const int r = nb_rank[k];
        if (r < 0 || r >= world || r == rank) return fail(PHW_EINVAL, "bad neighbour rank");
        pc.nb_rank[k] = r;
        const size_t nvp = ((size_t)peer_nv[r] + 31) & ~(size_t)31, nep = ((size_t)peer_ne[r] + 31) & ~(size_t)31;
        double* d0 = (double*)((char*)s->peer_arenas[r] + head);
        for (int i = 0; i < 3; ++i) { pc.nb_tp[k][i] = d0 + i * nvp; pc.nb_tq[k][i] = d0 + 3 * nvp + i * nep; }
        std::vector<int32_t> loc;
        int rc;
        int32_t* d;
        loc.resize(n_send_v[k]);
        for (int64_t i = 0; i < n_send_v[k]; ++i) {
            const int32_t u = send_v_user[k][i];
            if (u < 0 || u >= L.nv || L.v_old2new[u] >= L.nv_owned) return fail(PHW_EINVAL, "send list contains a vertex this rank does not own");
            if (send_v_remote[k][i] < 0 || send_v_remote[k][i] >= peer_nv[r]) return fail(PHW_EINVAL, "remote vertex index out of range");
            loc[i] = L.v_old2new[u];
        }
        rc = s->upload(&d, loc); if (rc) return rc; pc.send_v_loc[k] = d;
        loc.assign(send_v_remote[k], send_v_remote[k] + n_send_v[k]);
        rc = s->upload(&d, loc); if (rc) return rc; pc.send_v_rem[k] = d;
        pc.n_send_v[k] = (int)n_send_v[k];
        loc.resize(n_send_e[k]);
        for (int64_t i = 0; i < n_send_e[k]; ++i) {
            const int32_t u = send_e_user[k][i];
            if (u < 0 || u >= L.ne || L.e_old2new[u] >= L.ne_owned) return fail(PHW_EINVAL, "send list contains an edge this rank does not own");
            if (send_e_remote[k][i] < 0 || send_e_remote[k][i] >= peer_ne[r]) return fail(PHW_EINVAL, "remote edge index out of range");
            loc[i] = L.e_old2new[u];
        }
        rc = s->upload(&d, loc); if (rc) return rc; pc.send_e_loc[k] = d;
        loc.assign(send_e_remote[k], send_e_remote[k] + n_send_e[k]);
        rc = s->upload(&d, loc); if (rc) return rc; pc.send_e_rem[k] = d;
        pc.n_send_e[k] = (int)n_send_e[k];
    }
    {   // the same lists bucketed by owner patch: the pipelined solves push a patch's interface rows right after computing them
        for (int f = 0; f < 2; ++f) {
            std::vector<std::vector<int32_t>> locs(n_nb), rems(n_nb);
            for (int k = 0; k < n_nb; ++k) {
                const int64_t n = f == 0 ? n_send_v[k] : n_send_e[k];
                const int32_t* u = f == 0 ? send_v_user[k] : send_e_user[k];
                const int32_t* r = f == 0 ? send_v_remote[k] : send_e_remote[k];
                locs[k].resize(n); rems[k].assign(r, r + n);
                for (int64_t i = 0; i < n; ++i) locs[k][i] = f == 0 ? L.v_old2new[u[i]] : L.e_old2new[u[i]];
            }
            std::vector<int32_t> off; std::vector<PatchSendEntry> ent;
            bucket_send_lists(L.hdr, f == 0, n_nb, locs, rems, off, ent);
            int32_t* d_off; PatchSendEntry* d_ent;
            int rc = s->upload(&d_off, off); if (rc) return rc;
            rc = s->upload(&d_ent, ent); if (rc) return rc;
            pc.psend_off[f] = d_off; pc.psend_ent[f] = reinterpret_cast<const int4*>(d_ent);
        }
    }
    { const char* e = getenv("PHW_XCH_DEBUG"); pc.xch_dbg = e ? atoi(e) : 0; }
    CU(cudaStreamSynchronize(s->stream));
    s->pc = pc;
    return PHW_OK;
}

int phw_set_cheb_bounds(phw_solver_t s, double lam_min_q, double lam_max_q) {
    if (!s || !(lam_min_q > 0) || !(lam_max_q >= lam_min_q)) return fail(PHW_EINVAL, "bad bounds");
    s->lam_min_q = lam_min_q; s->lam_max_q = lam_max_q;
    s->cheb_d[1] = 0.5 * (lam_min_q + lam_max_q); s->cheb_c[1] = 0.5 * (lam_max_q - lam_min_q);
    return PHW_OK;
}

// inverse Jacobi diagonals of the solved blocks in user numbering (setup helper of the distributed skew estimate)
int phw_get_jacobi(phw_solver_t s, int32_t which, double* inv_diag) {
    if (!s || !inv_diag || (which != PHW_P && which != PHW_Q)) return fail(PHW_EINVAL, "bad arguments");
    CU(cudaSetDevice(s->device));
    return which == PHW_P ? export_vec(s, s->pinv_p, inv_diag, s->nv, s->d_v_new2old, s->io_v)
                          : export_vec(s, s->pinv_q, inv_diag, s->ne, s->d_e_new2old, s->io_e);
}

// implicit schemes in partitioned runs: after phw_comm_setup, with sigma = largest singular value of P_q^-1/2 D P_p^-1/2 of the
// GLOBAL operator (identical on every rank; solver.py computes it by a distributed power iteration through phw_apply_D / _DT).
// Collective: the bordered-column solve exchanges halo values and norms with the other ranks.
int phw_comm_finish_implicit(phw_solver_t s, double sigma) {
    if (!s || !(sigma > 0.0)) return fail(PHW_EINVAL, "bad arguments");
    if (!s->implicit_pending) return fail(PHW_ESTATE, "no implicit set-up is pending on this solver");
    CU(cudaSetDevice(s->device));
    std::vector<double> bl;
    if (s->prm.interconnection) {
        const Layout& L = s->mesh->L;
        bl.assign(s->ne, 0.0);
        for (int64_t k = 0; k < L.nb; ++k)
            if (L.btag[k] == PHW_TAG_LEFT) bl[L.e_old2new[L.bedge[k]]] = L.bsign[k] * L.bweight[k];
    }
    int rc = setup_implicit(s, bl, sigma);
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    s->implicit_pending = false;
    s->launches = 0;
    return PHW_OK;
}

int phw_solver_set_group_params(phw_solver_t s, int32_t n_groups, const double* K_wave, const double* rho, const double* lumped6) {
    if (!s || !K_wave || !rho || !lumped6) return fail(PHW_EINVAL, "NULL argument");
    if (n_groups != s->G) return fail(PHW_EINVAL, "n_groups does not match the mesh");
    if (s->G < 2) return fail(PHW_ESTATE, "the mesh has a single group: use phw_params");
    CU(cudaSetDevice(s->device));
    std::vector<double> gk(s->G), gir(s->G);
    for (int g = 0; g < s->G; ++g) {
        if (!(K_wave[g] > 0) || !(rho[g] > 0)) return fail(PHW_EINVAL, "K_wave and rho must be positive");
        const double* l = lumped6 + 6 * (size_t)g;   // Bl, R1, R2, K_em, m_em, L_em   (wave_2D.py:153-158)
        fill_physics(s->h_sp[g], s->prm.dt, K_wave[g], rho[g], l[0], l[1], l[2], l[3], l[4], l[5]);
        gk[g] = K_wave[g]; gir[g] = 1.0 / rho[g];
    }
    CU(cudaMemcpyAsync(s->d_sp, s->h_sp.data(), s->h_sp.size() * sizeof(ScalarPrm), cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(s->d_gK, gk.data(), gk.size() * 8, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(s->d_ginvrho, gir.data(), gir.size() * 8, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return PHW_OK;
}

int phw_get_group_states(phw_solver_t s, double* x3_all) {
    if (!s || !x3_all) return fail(PHW_EINVAL, "NULL argument");
    CU(cudaSetDevice(s->device));
    std::vector<Ctl> hcs(s->G);
    CU(cudaMemcpyAsync(hcs.data(), s->ctl, sizeof(Ctl) * (size_t)s->G, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    for (int g = 0; g < s->G; ++g) for (int i = 0; i < 3; ++i) x3_all[3 * g + i] = hcs[g].x[i];
    return PHW_OK;
}

int phw_set_state(phw_solver_t s, const double* p, const double* q, const double* x3) {
    if (!s) return fail(PHW_EINVAL, "solver is NULL");
    CU(cudaSetDevice(s->device));
    int rc;
    if (p) { rc = import_vec(s, p, s->p, s->nv, s->d_v_new2old, s->io_v); if (rc) return rc; }
    if (q) { rc = import_vec(s, q, s->q, s->ne, s->d_e_new2old, s->io_e); if (rc) return rc; }
    if (x3) {
        double h[3];
        if (is_device_ptr(x3)) CU(cudaMemcpyAsync(h, x3, 24, cudaMemcpyDeviceToHost, s->stream));
        else memcpy(h, x3, 24);
        CU(cudaStreamSynchronize(s->stream));
        for (int g = 0; g < s->G; ++g) CU(cudaMemcpyAsync(&s->ctl[g].x[0], h, 24, cudaMemcpyHostToDevice, s->stream));
    }
    reset_history(s);
    CU(cudaStreamSynchronize(s->stream));
    return PHW_OK;
}

int phw_set_time(phw_solver_t s, double t) {
    if (!s) return fail(PHW_EINVAL, "solver is NULL");
    CU(cudaSetDevice(s->device));
    for (int g = 0; g < s->G; ++g) CU(cudaMemcpyAsync(&s->ctl[g].t, &t, sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return PHW_OK;
}

int phw_get_state(phw_solver_t s, double* p, double* q, double* x3) {
    if (!s) return fail(PHW_EINVAL, "solver is NULL");
    CU(cudaSetDevice(s->device));
    int rc;
    if (p) { rc = export_vec(s, s->p, p, s->nv, s->d_v_new2old, s->io_v); if (rc) return rc; }
    if (q) { rc = export_vec(s, s->q, q, s->ne, s->d_e_new2old, s->io_e); if (rc) return rc; }
    if (x3) {
        double h[3];
        CU(cudaMemcpyAsync(h, &s->ctl->x[0], 24, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        if (is_device_ptr(x3)) CU(cudaMemcpy(x3, h, 24, cudaMemcpyHostToDevice));
        else memcpy(x3, h, 24);
    }
    CU(cudaStreamSynchronize(s->stream));
    return PHW_OK;
}

int phw_run(phw_solver_t s, int64_t n_steps, double* H_rows) {
    if (!s || n_steps < 0) return fail(PHW_EINVAL, "bad arguments");
    if (s->implicit_pending) return fail(PHW_ESTATE, "implicit scheme on a partitioned mesh: call phw_comm_setup and phw_comm_finish_implicit first");
    CU(cudaSetDevice(s->device));
    const size_t nrow = (size_t)n_steps * (size_t)s->G;
    if ((int64_t)nrow > s->dH_rows) {
        double* d = nullptr;
        int rc = s->dalloc(&d, nrow * s->ncol);
        if (rc) return rc;
        s->dH = d; s->dH_rows = (int64_t)nrow;
    }
    s->run_rows = n_steps;
    if (n_steps > 0) CU(cudaMemsetAsync(s->dH, 0, nrow * s->ncol * 8, s->stream));
    // row index restarts at 0 for every run; iteration statistics are per run
    std::vector<Ctl> hcs(s->G);
    CU(cudaMemcpyAsync(hcs.data(), s->ctl, sizeof(Ctl) * (size_t)s->G, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    for (auto& g : hcs) { g.step = 0; g.nonconv = 0; }
    for (int i = 0; i < 3; ++i) { hcs[0].sc[i].iters = 0; hcs[0].sc[i].solves = 0; hcs[0].sc[i].maxres = 0.0; }
    CU(cudaMemcpyAsync(s->ctl, hcs.data(), sizeof(Ctl) * (size_t)s->G, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    Ctl hc;
    s->launches = 0;
    s->kernel_paths = 0;
    CU(cudaEventRecord(s->ev0, s->stream));
    if (s->prm.analytical && !s->an_ready) return fail(PHW_ESTATE, "analytical run: call phw_set_analytical before phw_run");
    s->snap_taken = 0;
    const int64_t every = (s->snap_every > 0 && s->snap_dst) ? s->snap_every : 0;
    if (s->loop_on && s->pc.world <= 1 && n_steps > 0) {
        // one launch for the whole run, or one launch per snapshot interval
        for (int64_t done = 0; done < n_steps;) {
            const int64_t n = every ? std::min(every - done % every, n_steps - done) : n_steps;
            int rc = run_loop(s, n);
            if (rc) return rc;
            done += n;
            if (every && done % every == 0) { rc = take_snapshot(s); if (rc) return rc; }
        }
    } else
    for (int64_t n = 0; n < n_steps; ++n) {
        s->in_step = true;
        int rc = step(s);
        s->in_step = false;
        if (rc) return rc;
        if ((s->prm.flags & 2) && cudaStreamSynchronize(s->stream) != cudaSuccess) return fail(PHW_ECUDA, "kernel failure inside step loop");
        if (every && (n + 1) % every == 0) { rc = take_snapshot(s); if (rc) return rc; }
    }
    CU(cudaEventRecord(s->ev1, s->stream));
    if (s->snap_taken > 0) CU(cudaStreamSynchronize(s->snap_stream));
    if (H_rows && n_steps > 0) {
        if (is_device_ptr(H_rows)) CU(cudaMemcpyAsync(H_rows, s->dH, nrow * s->ncol * 8, cudaMemcpyDeviceToDevice, s->stream));
        else CU(cudaMemcpyAsync(H_rows, s->dH, nrow * s->ncol * 8, cudaMemcpyDeviceToHost, s->stream));
    }
    CU(cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    CU(cudaMemcpy(&hc, s->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
    phw_stats& st = s->stats;
    st.steps = n_steps;
    st.solves_p = hc.sc[0].solves; st.solves_q = hc.sc[1].solves; st.solves_block = hc.sc[2].solves;
    st.iters_p = hc.sc[0].iters; st.iters_q = hc.sc[1].iters; st.iters_block = hc.sc[2].iters;
    st.max_rel_residual = std::max(hc.sc[0].maxres, std::max(hc.sc[1].maxres, hc.sc[2].maxres));
    st.device_ms = ms;
    st.kernel_launches = s->launches;
    st.kernel_paths = s->kernel_paths;
    st.lam_min_q = s->lam_min_q; st.lam_max_q = s->lam_max_q; st.skew_bound = s->skew_bound;
    st.ms_solve_p = st.ms_solve_q = st.ms_solve_block = 0.0;
    for (size_t i = 0; i < s->prof_which.size(); ++i) {
        float t = 0.f;
        cudaEventElapsedTime(&t, s->prof_ev[2 * i], s->prof_ev[2 * i + 1]);
        (s->prof_which[i] == 0 ? st.ms_solve_p : (s->prof_which[i] == 1 ? st.ms_solve_q : st.ms_solve_block)) += t;
        cudaEventDestroy(s->prof_ev[2 * i]); cudaEventDestroy(s->prof_ev[2 * i + 1]);
    }
    s->prof_ev.clear(); s->prof_which.clear();
    if (s->prof_truncated) { st.ms_solve_p = st.ms_solve_q = st.ms_solve_block = 0.0; s->prof_truncated = false; }
    {   // algorithmic bytes: per cell 40 B skew apply, 28/60 B mass apply; per dof 8 B per vector pass
        const double nc = s->mesh ? (double)s->mesh->L.nc : 0.0, nv = s->nv, ne = s->ne;
        const double it_p = (double)st.iters_p + (double)st.iters_block, it_q = (double)st.iters_q + (double)st.iters_block;
        const double rhs = (double)(st.solves_p + st.solves_q + 2 * st.solves_block) * 40.0 * nc;
        double bytes = rhs + it_p * (28.0 * nc + 3 * 8.0 * nv) + it_q * (60.0 * nc + 3 * 8.0 * ne) + (double)n_steps * 88.0 * nc;
        bytes += (double)st.iters_block * 2 * 40.0 * nc;
        st.bytes_model = (int64_t)bytes;
    }
    if (hc.comm_dead) return fail(PHW_ECUDA, "inter-GPU exchange timed out (a peer rank stopped responding)");
    if ((hc.nonconv >> 32) != 0) return fail(PHW_ECUDA, "pipelined solve: a bulk-copy stage did not arrive (mbarrier wait timed out)");
    if (hc.nonconv > 0) {
        char buf[160];
        snprintf(buf, sizeof buf, "%lld solves hit max_iter without reaching tol (max relative residual %.3e)", (long long)hc.nonconv, st.max_rel_residual);
        return fail(PHW_ENOCONV, buf);
    }
    return PHW_OK;
}

int phw_set_snapshots(phw_solver_t s, int64_t every, double* snaps, int64_t capacity) {
    if (!s) return fail(PHW_EINVAL, "solver is NULL");
    if (every < 0 || capacity < 0 || (every > 0 && !snaps)) return fail(PHW_EINVAL, "bad snapshot arguments");
    CU(cudaSetDevice(s->device));
    if (s->snap_registered) { cudaHostUnregister(s->snap_registered); cudaGetLastError(); s->snap_registered = nullptr; }
    s->snap_every = every; s->snap_dst = every > 0 ? snaps : nullptr; s->snap_cap = capacity;
    if (every > 0 && capacity > 0 && !is_device_ptr(snaps)) {
        // page-lock a pageable destination for the lifetime of the setting, so that the device-to-host copies really run beside the
        // time loop (an already pinned buffer makes the call fail harmlessly)
        if (cudaHostRegister(snaps, (size_t)capacity * (size_t)s->nv * 8, cudaHostRegisterDefault) == cudaSuccess) s->snap_registered = snaps;
        else cudaGetLastError();
    }
    if (every > 0 && !s->snap_stream) {
        CU(cudaStreamCreateWithFlags(&s->snap_stream, cudaStreamNonBlocking));
        for (int i = 0; i < phw_solver_s::SNAP_SLOTS; ++i) {
            CU(cudaEventCreateWithFlags(&s->snap_ready[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s->snap_done[i], cudaEventDisableTiming));
        }
        int rc = s->dalloc(&s->snap_ring, (size_t)phw_solver_s::SNAP_SLOTS * s->nv);
        if (rc) return rc;
    }
    return PHW_OK;
}

int phw_debug_xch_times(phw_solver_t s, uint64_t* out320) {
    if (!s || !out320) return fail(PHW_EINVAL, "NULL argument");
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpy(out320, &s->ctl->xts[0][0][0], sizeof(unsigned long long) * 2 * 32 * 5, cudaMemcpyDeviceToHost));
    return PHW_OK;
}

int phw_get_stats(phw_solver_t s, phw_stats* out) {
    if (!s || !out) return fail(PHW_EINVAL, "NULL argument");
    s->stats.lam_min_q = s->lam_min_q; s->stats.lam_max_q = s->lam_max_q; s->stats.skew_bound = s->skew_bound;
    s->stats.kernel_paths = s->kernel_paths;
    *out = s->stats;
    return PHW_OK;
}

// ---- kernel-level entry points ------------------------------------------------------------------
int phw_apply_D(phw_solver_t s, const double* p, double* y) {
    if (!s || !p || !y) return fail(PHW_EINVAL, "NULL argument");
    CU(cudaSetDevice(s->device));
    int rc = import_vec(s, p, s->io_v2, s->nv, s->d_v_new2old, s->io_v); if (rc) return rc;
    if (s->ho_mf) { rc = ho_mf_D(s, s->io_v2, 1.0, s->io_e2); if (rc) return rc; }
    else if (s->ho) { csr_store(s, s->hD, s->io_v2, 1.0, s->io_e2); CU(cudaGetLastError()); }
    else { rc = op_edge_store(s, s->io_v2, nullptr, 0.0, 1.0, 0.0, s->io_e2); if (rc) return rc; }
    return export_vec(s, s->io_e2, y, s->ne, s->d_e_new2old, s->io_e);
}
int phw_apply_DT(phw_solver_t s, const double* q, double* y) {
    if (!s || !q || !y) return fail(PHW_EINVAL, "NULL argument");
    CU(cudaSetDevice(s->device));
    int rc = import_vec(s, q, s->io_e2, s->ne, s->d_e_new2old, s->io_e); if (rc) return rc;
    if (s->ho_mf) { rc = ho_mf_DT(s, s->io_e2, 1.0, s->io_v2); if (rc) return rc; }
    else if (s->ho) { csr_store(s, s->hDT, s->io_e2, 1.0, s->io_v2); CU(cudaGetLastError()); }
    else { rc = op_vertex_store(s, nullptr, s->io_e2, 0.0, 1.0, 0.0, s->io_v2); if (rc) return rc; }
    return export_vec(s, s->io_v2, y, s->nv, s->d_v_new2old, s->io_v);
}
int phw_mass_matvec(phw_solver_t s, int32_t which, const double* x, double* y) {
    if (!s || !x || !y) return fail(PHW_EINVAL, "NULL argument");
    CU(cudaSetDevice(s->device));
    int rc;
    if (which == PHW_P) {
        rc = import_vec(s, x, s->io_v2, s->nv, s->d_v_new2old, s->io_v); if (rc) return rc;
        if (s->ho_mf) { rc = ho_mf_apply(s, HO_MP, s->io_v2, s->tmp_p, 1.0, true); if (rc) return rc; }
        else if (s->ho) { csr_store(s, s->hMp, s->io_v2, 1.0, s->tmp_p); CU(cudaGetLastError()); }
        else { rc = op_vertex_store(s, s->io_v2, nullptr, 1.0, 0.0, 0.0, s->tmp_p); if (rc) return rc; }
        return export_vec(s, s->tmp_p, y, s->nv, s->d_v_new2old, s->io_v);
    } else if (which == PHW_Q) {
        rc = import_vec(s, x, s->io_e2, s->ne, s->d_e_new2old, s->io_e); if (rc) return rc;
        if (s->ho_mf) { rc = ho_mf_apply(s, HO_MQ, s->io_e2, s->tmp_q, 1.0, true); if (rc) return rc; }
        else if (s->ho) { csr_store(s, s->hMq, s->io_e2, 1.0, s->tmp_q); CU(cudaGetLastError()); }
        else { rc = op_edge_store(s, nullptr, s->io_e2, 1.0, 0.0, 0.0, s->tmp_q); if (rc) return rc; }
        return export_vec(s, s->tmp_q, y, s->ne, s->d_e_new2old, s->io_e);
    }
    return fail(PHW_EINVAL, "which must be PHW_P or PHW_Q");
}
int phw_mass_solve(phw_solver_t s, int32_t which, const double* b, double* x, int32_t* iters) {
    if (!s || !b || !x) return fail(PHW_EINVAL, "NULL argument");
    if (which != PHW_P && which != PHW_Q) return fail(PHW_EINVAL, "which must be PHW_P or PHW_Q");
    CU(cudaSetDevice(s->device));
    const int n = which == PHW_P ? s->nv : s->ne;
    int rc = import_vec(s, b, which == PHW_P ? s->bp : s->bq, n, which == PHW_P ? s->d_v_new2old : s->d_e_new2old, which == PHW_P ? s->io_v : s->io_e);
    if (rc) return rc;
    TriBuf tb = which == PHW_P ? s->tp : s->tq;
    copy_to_tb_k<<<vec_grid(s, n), 256, 0, s->stream>>>(n, tb, s->ctl, which, nullptr);   // cold start
    CU(cudaGetLastError());
    long long it0 = 0, it1 = 0;
    CU(cudaMemcpyAsync(&it0, &s->ctl->sc[which].iters, 8, cudaMemcpyDeviceToHost, s->stream));
    rc = solve_mass(s, which, s->prm.max_iter); if (rc) return rc;
    double* out = which == PHW_P ? s->tmp_p : s->tmp_q;
    copy_from_tb_k<<<vec_grid(s, n), 256, 0, s->stream>>>(n, out, tb, s->ctl, which);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(&it1, &s->ctl->sc[which].iters, 8, cudaMemcpyDeviceToHost, s->stream));
    rc = export_vec(s, out, x, n, which == PHW_P ? s->d_v_new2old : s->d_e_new2old, which == PHW_P ? s->io_v : s->io_e);
    if (rc) return rc;
    long long nonconv = 0;
    CU(cudaMemcpyAsync(&nonconv, &s->ctl->nonconv, 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (iters) *iters = (int32_t)(it1 - it0);
    reset_history(s);
    if ((nonconv >> 32) != 0) return fail(PHW_ECUDA, "pipelined solve: a bulk-copy stage did not arrive (mbarrier wait timed out)");
    return PHW_OK;
}
int phw_energy(phw_solver_t s, double* parts2) {
    if (!s || !parts2) return fail(PHW_EINVAL, "NULL argument");
    CU(cudaSetDevice(s->device));
    int rc = energy(s); if (rc) return rc;
    CU(cudaMemcpyAsync(parts2, &s->ctl->red[0], 16, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return PHW_OK;
}
int phw_set_analytical(phw_solver_t s, double amp, double kx, double ky, double omega, const double* mass_ref225,
                       const double* lam_nodes45, const double* exact_vertices, const int32_t* pt_vertices3,
                       const double* pt_lam3) {
    if (!s || !mass_ref225 || !lam_nodes45 || !exact_vertices || !pt_vertices3 || !pt_lam3) return fail(PHW_EINVAL, "NULL argument");
    if (s->ho) return fail(PHW_ESTATE, "use phw_ho_set_analytical for higher-order solvers");
    if (s->pc.world > 1) return fail(PHW_EINVAL, "analytical diagnostics are not available in partitioned runs");
    CU(cudaSetDevice(s->device));
    const Layout& L = s->mesh->L;
    AnalyticPrm& a = s->an;
    a.amp = amp; a.kx = kx; a.ky = ky; a.omega = omega;
    for (int n = 0; n < 15; ++n) for (int k = 0; k < 3; ++k) a.lam[n][k] = lam_nodes45[3 * n + k];
    for (int k = 0; k < 3; ++k) {
        a.pt_v[k] = pt_vertices3[k] >= 0 ? L.v_old2new[pt_vertices3[k]] : -1;
        a.pt_lam[k] = pt_lam3[k];
    }
    CU(cudaMemcpyToSymbol(c_mass_ref, mass_ref225, 225 * sizeof(double)));
    if (!s->d_cells_new) {
        std::vector<int32_t> cn(3 * (size_t)L.nc);
        for (size_t h = 0; h < cn.size(); ++h) cn[h] = L.v_old2new[L.cells[h]];
        std::vector<double> vn(2 * (size_t)L.nv), ev((size_t)L.nv);
        for (int64_t v = 0; v < L.nv; ++v) {
            const int32_t nw = L.v_old2new[v];
            vn[2 * (size_t)nw] = L.vtx[2 * v]; vn[2 * (size_t)nw + 1] = L.vtx[2 * v + 1];
            ev[nw] = exact_vertices[v];
        }
        int rc = s->upload(&s->d_cells_new, cn); if (rc) return rc;
        rc = s->upload(&s->d_vtx_new, vn); if (rc) return rc;
        rc = s->upload(&s->d_exact_v, ev); if (rc) return rc;
        rc = s->dalloc(&s->d_diag_part, 4 * (size_t)s->sm_count); if (rc) return rc;
        CU(cudaStreamSynchronize(s->stream));
    }
    s->an_ready = true;
    return PHW_OK;
}

// ---- higher-order pairs: solver on host-assembled CSR operators ----------------------------------------------
static int upload_csr(phw_solver_s* s, int64_t n_rows, const int32_t* indptr, const int32_t* idx, const double* val, Csr* out) {
    if (!indptr || !idx || !val || n_rows <= 0) return fail(PHW_EINVAL, "bad CSR matrix");
    std::vector<int32_t> ip(indptr, indptr + n_rows + 1), ix(idx, idx + indptr[n_rows]);
    std::vector<double> v(val, val + indptr[n_rows]);
    int *d_ip, *d_ix; double* d_v;
    int rc = s->upload(&d_ip, ip); if (rc) return rc;
    rc = s->upload(&d_ix, ix); if (rc) return rc;
    rc = s->upload(&d_v, v); if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    out->n_rows = (int)n_rows; out->indptr = d_ip; out->idx = d_ix; out->val = d_v;
    const double avg = (double)indptr[n_rows] / (double)n_rows;     // threads per row ~ entries per row / 2
    out->lanes = avg <= 8 ? 4 : (avg <= 16 ? 8 : (avg <= 40 ? 16 : 32));
    return 0;
}

// common part of the two higher-order constructors: vectors, control block, identity numbering, port list = non-zeros of b_L,
// Chebyshev intervals.  The caller uploads its operators and fills pinv_p / pinv_q afterwards.
static int ho_create_base(const phw_params* prm, int32_t device, void* cuda_stream, int64_t n_p, int64_t n_q, const double* b_L,
                          double lam_min_p, double lam_max_p, double lam_min_q, double lam_max_q, phw_solver_s** out, int32_t n_cases = 1) {
    if (!prm || !out || !b_L) return fail(PHW_EINVAL, "NULL argument");
    if (n_cases < 1 || n_cases > 16384 || (int64_t)n_cases * std::max(n_p, n_q) > 2000000000LL)
        return fail(PHW_EINVAL, "number of cases out of range (at most 16 384 per sweep solver and 2e9 dof x cases: shard the case list)");
    if (n_cases > 1 && !((prm->scheme == PHW_SE || prm->scheme == PHW_SV || prm->scheme == PHW_EE || prm->scheme == PHW_EH) && !prm->analytical))
        return fail(PHW_EINVAL, "batched runs support the explicit weak-form schemes only");
    if (prm->scheme < 0 || prm->scheme > 5) return fail(PHW_EINVAL, "unknown scheme");
    if (prm->interconnection && !(prm->scheme == PHW_SE || prm->scheme == PHW_SV || prm->scheme == PHW_SM))
        return fail(PHW_EINVAL, "only SE, SV and SM time int schemes implemented for interconnection model");
    if (prm->strong || prm->casimir) return fail(PHW_EINVAL, "higher-order pairs: weak boundary conditions, no casimir control");
    if (!(lam_min_p > 0) || !(lam_min_q > 0) || lam_max_p < lam_min_p || lam_max_q < lam_min_q) return fail(PHW_EINVAL, "bad spectral bounds");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(PHW_ECUDA, "no CUDA device available (libphwave has no CPU fallback)"); }
    if (device < 0 || device >= ndev) return fail(PHW_EINVAL, "device index out of range");
    CU(cudaSetDevice(device));
    phw_solver_s* s = new (std::nothrow) phw_solver_s();
    if (!s) return fail(PHW_ENOMEM, "out of host memory");
    s->mesh = nullptr; s->prm = *prm; s->device = device; s->stream = (cudaStream_t)cuda_stream; s->ho = true;
    s->prm.flags |= env_flags();
    if (!(s->prm.tol > 0)) s->prm.tol = 1e-13;
    if (s->prm.max_iter <= 0) s->prm.max_iter = 1000;
    // batched sweep (n_cases > 1): every vector is [n_dof][n_cases], so the element-wise kernels simply see longer vectors
    s->nv = (int)(n_p * n_cases); s->ne = (int)(n_q * n_cases); s->npatch = 0; s->G = n_cases; s->bcsr = n_cases > 1 ? n_cases : 0;
    auto bail = [&](int rc) { std::string keep = g_err; phw_solver_destroy(s); g_err = keep; return rc; };
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(fail(PHW_ECUDA, "cudaGetDeviceProperties failed"));
    s->sm_count = prop.multiProcessorCount;
    if (cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess) return bail(fail(PHW_ECUDA, "event creation failed"));
    int rc = setup_vectors(s); if (rc) return bail(rc);
    if (!s->coop) return bail(fail(PHW_ECUDA, "cooperative launch is required"));
    rc = setup_scalars(s); if (rc) return bail(rc);
    {
        std::vector<int32_t> idv(s->nv), ide(s->ne), pe, pg, goff(2, 0), gpoff(2, 0);
        std::iota(idv.begin(), idv.end(), 0); std::iota(ide.begin(), ide.end(), 0);
        std::vector<double> bl;
        for (int64_t r = 0; r < n_q; ++r) if (b_L[r] != 0.0) { pe.push_back((int32_t)r); bl.push_back(b_L[r]); pg.push_back(0); }
        goff[1] = (int32_t)pe.size();
        s->nport = (int)pe.size();
        rc = s->upload(&s->d_v_new2old, idv); if (rc) return bail(rc);
        rc = s->upload(&s->d_e_new2old, ide); if (rc) return bail(rc);
        rc = s->upload(&s->d_port_e, pe); if (rc) return bail(rc);
        rc = s->upload(&s->d_port_bl, bl); if (rc) return bail(rc);
        rc = s->upload(&s->d_port_goff, goff); if (rc) return bail(rc);
        rc = s->upload(&s->d_port_group, pg); if (rc) return bail(rc);
        rc = s->upload(&s->d_gpoff, gpoff); if (rc) return bail(rc);
    }
    if (s->bcsr) {
        std::vector<double> gk(s->G, prm->K_wave), gir(s->G, 1.0 / prm->rho);
        rc = s->upload(&s->d_gK, gk); if (rc) return bail(rc);
        rc = s->upload(&s->d_ginvrho, gir); if (rc) return bail(rc);
        // [row-owning warp][case] partial sums of csr_dot_b_k: (warps of a launch / case chunks) x cases <= 32 x 4 (cases per lane) x warps
        s->bpart_cap = (size_t)8 * 8 * s->sm_count * 32 * 4 + (size_t)s->bcsr;
        rc = s->dalloc(&s->bpart, s->bpart_cap); if (rc) return bail(rc);
    }
    s->cheb_d[0] = 0.5 * (lam_min_p + lam_max_p); s->cheb_c[0] = 0.5 * (lam_max_p - lam_min_p);
    s->cheb_d[1] = 0.5 * (lam_min_q + lam_max_q); s->cheb_c[1] = 0.5 * (lam_max_q - lam_min_q);
    s->lam_min_q = lam_min_q; s->lam_max_q = lam_max_q;
    s->launches = 0;
    *out = s;
    return PHW_OK;
}

static int create_csr_solver(const phw_params* prm, int32_t device, void* cuda_stream, int32_t n_cases, int64_t n_p, int64_t n_q,
                             const int32_t* Mp_indptr, const int32_t* Mp_idx, const double* Mp_val,
                             const int32_t* Mq_indptr, const int32_t* Mq_idx, const double* Mq_val,
                             const int32_t* D_indptr, const int32_t* D_idx, const double* D_val,
                             const int32_t* DT_indptr, const int32_t* DT_idx, const double* DT_val,
                             const double* b_L, double lam_min_p, double lam_max_p, double lam_min_q, double lam_max_q,
                             phw_solver_t* out) {
    if (!Mp_indptr || !Mp_idx || !Mp_val || !Mq_indptr || !Mq_idx || !Mq_val || !D_indptr || !D_idx || !D_val || !DT_indptr || !DT_idx || !DT_val)
        return fail(PHW_EINVAL, "NULL argument");
    phw_solver_s* s = nullptr;
    int rc = ho_create_base(prm, device, cuda_stream, n_p, n_q, b_L, lam_min_p, lam_max_p, lam_min_q, lam_max_q, &s, n_cases);
    if (rc) return rc;
    auto bail = [&](int rc_) { std::string keep = g_err; phw_solver_destroy(s); g_err = keep; return rc_; };
    rc = upload_csr(s, n_p, Mp_indptr, Mp_idx, Mp_val, &s->hMp); if (rc) return bail(rc);
    rc = upload_csr(s, n_q, Mq_indptr, Mq_idx, Mq_val, &s->hMq); if (rc) return bail(rc);
    rc = upload_csr(s, n_q, D_indptr, D_idx, D_val, &s->hD); if (rc) return bail(rc);
    rc = upload_csr(s, n_p, DT_indptr, DT_idx, DT_val, &s->hDT); if (rc) return bail(rc);
    {   // Jacobi weights from the CSR diagonals
        std::vector<double> pp(n_p, 0.0), pq(n_q, 0.0);
        for (int64_t r = 0; r < n_p; ++r) for (int j = Mp_indptr[r]; j < Mp_indptr[r + 1]; ++j) if (Mp_idx[j] == r) pp[r] = 1.0 / Mp_val[j];
        for (int64_t r = 0; r < n_q; ++r) for (int j = Mq_indptr[r]; j < Mq_indptr[r + 1]; ++j) if (Mq_idx[j] == r) pq[r] = 1.0 / Mq_val[j];
        if (cudaMemcpyAsync(s->pinv_p, pp.data(), n_p * 8, cudaMemcpyHostToDevice, s->stream) != cudaSuccess ||
            cudaMemcpyAsync(s->pinv_q, pq.data(), n_q * 8, cudaMemcpyHostToDevice, s->stream) != cudaSuccess) return bail(fail(PHW_ECUDA, "memcpy failed"));
        if (cudaStreamSynchronize(s->stream) != cudaSuccess) return bail(fail(PHW_ECUDA, "sync failed"));
    }
    if (prm->scheme == PHW_IE || prm->scheme == PHW_SM) {
        rc = setup_implicit(s, prm->interconnection ? std::vector<double>(b_L, b_L + n_q) : std::vector<double>());
        if (rc) return bail(rc);
        if (cudaStreamSynchronize(s->stream) != cudaSuccess) return bail(fail(PHW_ECUDA, "sync failed"));
        s->launches = 0;
    }
    *out = s;
    return PHW_OK;
}

int phw_ho_create(const phw_params* prm, int32_t device, void* cuda_stream, int64_t n_p, int64_t n_q,
                  const int32_t* Mp_indptr, const int32_t* Mp_idx, const double* Mp_val,
                  const int32_t* Mq_indptr, const int32_t* Mq_idx, const double* Mq_val,
                  const int32_t* D_indptr, const int32_t* D_idx, const double* D_val,
                  const int32_t* DT_indptr, const int32_t* DT_idx, const double* DT_val,
                  const double* b_L, double lam_min_p, double lam_max_p, double lam_min_q, double lam_max_q,
                  phw_solver_t* out) {
    return create_csr_solver(prm, device, cuda_stream, 1, n_p, n_q, Mp_indptr, Mp_idx, Mp_val, Mq_indptr, Mq_idx, Mq_val,
                             D_indptr, D_idx, D_val, DT_indptr, DT_idx, DT_val, b_L, lam_min_p, lam_max_p, lam_min_q, lam_max_q, out);
}

// batched parameter sweep with the case index fastest (kernels_sweep.cuh): n_cases independent cases on ONE set of assembled operators;
// state vectors are [n_dof][n_cases]; per-case physics through phw_solver_set_group_params
int phw_sweep_create(const phw_params* prm, int32_t device, void* cuda_stream, int32_t n_cases, int64_t n_p, int64_t n_q,
                     const int32_t* Mp_indptr, const int32_t* Mp_idx, const double* Mp_val,
                     const int32_t* Mq_indptr, const int32_t* Mq_idx, const double* Mq_val,
                     const int32_t* D_indptr, const int32_t* D_idx, const double* D_val,
                     const int32_t* DT_indptr, const int32_t* DT_idx, const double* DT_val,
                     const double* b_L, double lam_min_p, double lam_max_p, double lam_min_q, double lam_max_q,
                     phw_solver_t* out) {
    if (n_cases < 2) return fail(PHW_EINVAL, "a sweep needs at least two cases (one case: phw_ho_create / phw_solver_create)");
    return create_csr_solver(prm, device, cuda_stream, n_cases, n_p, n_q, Mp_indptr, Mp_idx, Mp_val, Mq_indptr, Mq_idx, Mq_val,
                             D_indptr, D_idx, D_val, DT_indptr, DT_idx, DT_val, b_L, lam_min_p, lam_max_p, lam_min_q, lam_max_q, out);
}

int phw_ho_create_mf(const phw_params* prm, int32_t device, void* cuda_stream, int32_t k_p, int32_t k_q,
                     int64_t n_vertices, int64_t n_edges, int64_t n_cells,
                     const int32_t* cell_vertices, const int32_t* cell_edges, const double* geom4,
                     const double* Mp_hat, const double* C_hat, const double* A00, const double* A01s, const double* A11,
                     int64_t nN_rows, const int32_t* N_row, const int32_t* N_indptr, const int32_t* N_idx, const double* N_val,
                     int64_t nNT_rows, const int32_t* NT_row, const int32_t* NT_indptr, const int32_t* NT_idx, const double* NT_val,
                     const double* b_L, double lam_min_p, double lam_max_p, double lam_min_q, double lam_max_q, phw_solver_t* out) {
    if (!cell_vertices || !cell_edges || !geom4 || !Mp_hat || !C_hat || !A00 || !A01s || !A11) return fail(PHW_EINVAL, "NULL argument");
    if (k_p < 1 || k_p > 3 || k_q < 1 || k_q > 3) return fail(PHW_EINVAL, "basis orders 1..3 are supported");
    if (prm && (prm->scheme == PHW_IE || prm->scheme == PHW_SM))
        return fail(PHW_EINVAL, "the implicit schemes of the higher-order pairs run on the assembled operators (phw_ho_create)");
    if (nN_rows < 0 || nNT_rows < 0 || (nN_rows && (!N_row || !N_indptr || !N_idx || !N_val)) || (nNT_rows && (!NT_row || !NT_indptr || !NT_idx || !NT_val)))
        return fail(PHW_EINVAL, "bad boundary operator");
    const int ndp = (k_p + 1) * (k_p + 2) / 2, ndq = k_q * (k_q + 2);
    const int64_t n_p = n_vertices + n_edges * (k_p - 1) + n_cells * ((k_p - 1) * (k_p - 2) / 2);
    const int64_t n_q = n_edges * k_q + n_cells * (int64_t)(k_q * (k_q - 1));
    phw_solver_s* s = nullptr;
    int rc = ho_create_base(prm, device, cuda_stream, n_p, n_q, b_L, lam_min_p, lam_max_p, lam_min_q, lam_max_q, &s);
    if (rc) return rc;
    auto bail = [&](int rc_) { std::string keep = g_err; phw_solver_destroy(s); g_err = keep; return rc_; };
    HoMf& m = s->hm;
    m.nc = (int)n_cells; m.nv = (int)n_vertices; m.ne = (int)n_edges; m.kp = k_p; m.kq = k_q; m.ndp = ndp; m.ndq = ndq;
    m.n_p = (int)n_p; m.n_q = (int)n_q;
    std::vector<double> ref;
    ref.insert(ref.end(), Mp_hat, Mp_hat + ndp * ndp);
    m.off_C = (int)ref.size(); ref.insert(ref.end(), C_hat, C_hat + ndq * ndp);
    m.off_A00 = (int)ref.size(); ref.insert(ref.end(), A00, A00 + ndq * ndq);
    m.off_A01 = (int)ref.size(); ref.insert(ref.end(), A01s, A01s + ndq * ndq);
    m.off_A11 = (int)ref.size(); ref.insert(ref.end(), A11, A11 + ndq * ndq);
    m.ref_len = (int)ref.size();
    {
        int *d_cv = nullptr, *d_ce = nullptr; double *d_geom = nullptr, *d_ref = nullptr;
        rc = s->upload(&d_cv, std::vector<int32_t>(cell_vertices, cell_vertices + 3 * n_cells)); if (rc) return bail(rc);
        rc = s->upload(&d_ce, std::vector<int32_t>(cell_edges, cell_edges + 3 * n_cells)); if (rc) return bail(rc);
        rc = s->upload(&d_geom, std::vector<double>(geom4, geom4 + 4 * n_cells)); if (rc) return bail(rc);
        rc = s->upload(&d_ref, ref); if (rc) return bail(rc);
        m.cv = d_cv; m.ce = d_ce; m.geom = d_geom; m.ref = d_ref;
    }
    auto up_rows = [&](int64_t nr, const int32_t* row, const int32_t* ip, const int32_t* ix, const double* va, RowCsr* A) -> int {
        A->n_rows = (int)nr;
        if (nr == 0) return 0;
        int *d_row = nullptr, *d_ip = nullptr, *d_ix = nullptr; double* d_va = nullptr;
        int r_ = s->upload(&d_row, std::vector<int32_t>(row, row + nr)); if (r_) return r_;
        r_ = s->upload(&d_ip, std::vector<int32_t>(ip, ip + nr + 1)); if (r_) return r_;
        r_ = s->upload(&d_ix, std::vector<int32_t>(ix, ix + ip[nr])); if (r_) return r_;
        r_ = s->upload(&d_va, std::vector<double>(va, va + ip[nr])); if (r_) return r_;
        A->row = d_row; A->indptr = d_ip; A->idx = d_ix; A->val = d_va;
        return 0;
    };
    rc = up_rows(nN_rows, N_row, N_indptr, N_idx, N_val, &s->hN); if (rc) return bail(rc);
    rc = up_rows(nNT_rows, NT_row, NT_indptr, NT_idx, NT_val, &s->hNT); if (rc) return bail(rc);
    rc = s->dalloc(&s->ho_yacc_p, (size_t)n_p); if (rc) return bail(rc);
    rc = s->dalloc(&s->ho_yacc_q, (size_t)n_q); if (rc) return bail(rc);
    s->ho_mf = true;
    // Jacobi weights: diagonals accumulated on the device from the reference diagonals and the cell geometry
    if (cudaMemsetAsync(s->pinv_p, 0, n_p * 8, s->stream) != cudaSuccess || cudaMemsetAsync(s->pinv_q, 0, n_q * 8, s->stream) != cudaSuccess ||
        cudaMemsetAsync(s->ho_yacc_p, 0, n_p * 8, s->stream) != cudaSuccess || cudaMemsetAsync(s->ho_yacc_q, 0, n_q * 8, s->stream) != cudaSuccess)
        return bail(fail(PHW_ECUDA, "memset failed"));
    ho_diag_k<HO_MP><<<ho_grid(s), HO_NT, ho_smem(m), s->stream>>>(m, s->pinv_p);
    ho_diag_k<HO_MQ><<<ho_grid(s), HO_NT, ho_smem(m), s->stream>>>(m, s->pinv_q);
    invert_k<<<nblk((int)n_p), 256, 0, s->stream>>>((int)n_p, s->pinv_p);
    invert_k<<<nblk((int)n_q), 256, 0, s->stream>>>((int)n_q, s->pinv_q);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(s->stream) != cudaSuccess) return bail(fail(PHW_ECUDA, "matrix-free setup kernels failed"));
    *out = s;
    return PHW_OK;
}

int phw_ho_set_analytical(phw_solver_t s, double omega, int32_t n_cells, int32_t ndp, int32_t nn, const int32_t* cell_dofs,
                          const double* T, const double* mass_ref, const double* exact_nodes, const double* area,
                          const double* exact_dofs, int32_t nprobe, const int32_t* probe_dofs, const double* probe_w) {
    if (!s || !s->ho || !cell_dofs || !T || !mass_ref || !exact_nodes || !area || !exact_dofs) return fail(PHW_EINVAL, "bad arguments");
    if (ndp > 10 || nn > 28 || nprobe > 10 || nprobe < 0) return fail(PHW_EINVAL, "tabulation sizes exceed P3 -> P6");
    CU(cudaSetDevice(s->device));
    s->ho_nc = n_cells; s->ho_ndp = ndp; s->ho_nn = nn; s->ho_nprobe = nprobe; s->ho_omega = omega;
    int rc;
    rc = s->upload(&s->ho_cell_dofs, std::vector<int32_t>(cell_dofs, cell_dofs + (size_t)n_cells * ndp)); if (rc) return rc;
    rc = s->upload(&s->ho_T, std::vector<double>(T, T + (size_t)nn * ndp)); if (rc) return rc;
    rc = s->upload(&s->ho_mass, std::vector<double>(mass_ref, mass_ref + (size_t)nn * nn)); if (rc) return rc;
    rc = s->upload(&s->ho_exact_nodes, std::vector<double>(exact_nodes, exact_nodes + (size_t)n_cells * nn)); if (rc) return rc;
    rc = s->upload(&s->ho_area, std::vector<double>(area, area + n_cells)); if (rc) return rc;
    rc = s->upload(&s->ho_exact_dofs, std::vector<double>(exact_dofs, exact_dofs + s->nv)); if (rc) return rc;
    rc = s->upload(&s->ho_probe_dofs, std::vector<int32_t>(probe_dofs, probe_dofs + nprobe)); if (rc) return rc;
    rc = s->upload(&s->ho_probe_w, std::vector<double>(probe_w, probe_w + nprobe)); if (rc) return rc;
    if (!s->d_diag_part) { rc = s->dalloc(&s->d_diag_part, 4 * (size_t)s->sm_count); if (rc) return rc; }
    CU(cudaStreamSynchronize(s->stream));
    s->an_ready = true;
    return PHW_OK;
}

int phw_set_H_init(phw_solver_t s, double H_init) {
    if (!s) return fail(PHW_EINVAL, "solver is NULL");
    s->prm.H_init = H_init;
    s->sp.H_init = H_init;
    for (auto& g : s->h_sp) g.H_init = H_init;
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(s->d_sp, s->h_sp.data(), s->h_sp.size() * sizeof(ScalarPrm), cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return PHW_OK;
}
int phw_port_flux(phw_solver_t s, double* out) {
    if (!s || !out) return fail(PHW_EINVAL, "NULL argument");
    CU(cudaSetDevice(s->device));
    int rc = flux(s, s->q, 1, 3); if (rc) return rc;
    CU(cudaMemcpyAsync(out, &s->ctl->flux[3], 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return PHW_OK;
}

}  // extern "C"
