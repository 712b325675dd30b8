// emu_main.cpp - runs the pipelined kernels of porthamiltonian_fem_b200/csrc on the CPU inside the SPMD emulator (emu_rt.hpp).
// Test infrastructure (tests/test_emu_cpu.py): checks indexing, shared-memory carve-ups, the staging protocol (waits before
// use, expected byte counts, 16-byte rules, nothing in flight at exit) and the arithmetic against the CPU oracle - everything
// except real asynchrony and performance.  The device arrays are set up here the way phwave.cu does it on the GPU (geometry,
// packed cell records, ELL weights, Jacobi weights), from the same host-side Layout.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "cuda_runtime.h"
#include "kernels_pipe.cuh"
#include "kernels_sweep.cuh"
#include "pipe_carve.hpp"
#include "layout.hpp"

using namespace phw;

namespace {

struct Arena {
    std::vector<phw_emu::Guarded> all;
    template <class T> T* get(size_t n) {
        phw_emu::Guarded g = phw_emu::guarded_alloc(std::max<size_t>(n, 1) * sizeof(T));
        std::memset(g.ptr, 0, (std::max<size_t>(n, 1) * sizeof(T) + 15) & ~(size_t)15);
        all.push_back(g);
        return (T*)g.ptr;
    }
    template <class T> T* put(const std::vector<T>& v) {
        T* p = get<T>(v.size());
        if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
        return p;
    }
    ~Arena() { for (auto& g : all) phw_emu::guarded_free(g); }
};

void cell_geometry(const Layout& L, int32_t c, double& A, double& G0, double& G1, double& G2) {   // geom_k of phwave.cu
    const int a = L.cells[3 * (int64_t)c], b = L.cells[3 * (int64_t)c + 1], d = L.cells[3 * (int64_t)c + 2];
    const double ax = L.vtx[2 * a], ay = L.vtx[2 * a + 1], bx = L.vtx[2 * b], by = L.vtx[2 * b + 1], dx = L.vtx[2 * d], dy = L.vtx[2 * d + 1];
    A = 0.5 * std::fabs((bx - ax) * (dy - ay) - (by - ay) * (dx - ax));
    const double l0 = (dx - bx) * (dx - bx) + (dy - by) * (dy - by);
    const double l1 = (ax - dx) * (ax - dx) + (ay - dy) * (ay - dy);
    const double l2 = (bx - ax) * (bx - ax) + (by - ay) * (by - ay);
    const double inv = 1.0 / (48.0 * A);
    G0 = l0 * inv; G1 = l1 * inv; G2 = l2 * inv;
}

uint4 pack_a(uint2 r, double a) { uint64_t u; std::memcpy(&u, &a, 8); return make_uint4(r.x, r.y, (unsigned)(u & 0xFFFFFFFFu), (unsigned)(u >> 32)); }
uint4 pack_b(double b, double c) {
    uint64_t u, v; std::memcpy(&u, &b, 8); std::memcpy(&v, &c, 8);
    return make_uint4((unsigned)(u & 0xFFFFFFFFu), (unsigned)(u >> 32), (unsigned)(v & 0xFFFFFFFFu), (unsigned)(v >> 32));
}

int fail(char* err, int errlen, const std::string& m) { std::snprintf(err, (size_t)errlen, "%s", m.c_str()); return -1; }

}  // namespace

// what: "mq1" | "mp1" (pipelined solves), "mq_lean" | "mp_lean",
//       "energy" (out[0] = p^T M_p p, out[1] = q^T M_q q; in1 = p, in2 = q), "rhs_e" (out = D in1), "rhs_v" (out = D^T in1)
// btags: tag of every boundary edge in the order of the layout's boundary list (nullptr: all "none"); user numbering everywhere.
extern "C" int emu_run(const char* what_c, int64_t nv, const double* vtx, int64_t nc, const int32_t* cells, int32_t patch_cells, int32_t grid,
                       const int32_t* btags, const double* in1, const double* in2, double* out, int32_t* iters, double tol,
                       char* err, int errlen) {
    const std::string what(what_c);
    constexpr int NT = 128;
    Layout L;
    std::string e = L.build(nv, vtx, nc, cells, patch_cells);
    if (!e.empty()) return fail(err, errlen, e);
    if (btags) for (int64_t k = 0; k < L.nb; ++k) L.btag[k] = btags[k];
    L.boundary_set = true;
    std::vector<CellRec> rec;
    L.make_records(false, false, rec);
    const int npatch = L.npatch_rows, npc = (int)L.pcell_user.size();   // patches that own rows (layout.hpp)
    grid = std::max(1, std::min(grid, npatch));
    Arena A;
    PatchHdr* d_hdr = A.put(L.hdr);
    std::vector<uint2> hv(rec.size()), he(rec.size());
    for (size_t k = 0; k < rec.size(); ++k) { hv[k] = make_uint2(rec[k].x, rec[k].y); he[k] = make_uint2(rec[k].z, rec[k].w); }
    uint2* d_recv = A.put(hv); uint2* d_rece = A.put(he);
    std::vector<double> area(npc), g0(npc), g1(npc), g2(npc);
    for (int k = 0; k < npc; ++k) cell_geometry(L, L.pcell_user[k], area[k], g0[k], g1[k], g2[k]);
    double* d_area = A.put(area);
    std::vector<uint4> cq(2 * (size_t)npc);
    for (int k = 0; k < npc; ++k) { cq[2 * (size_t)k] = pack_a(he[k], g0[k]); cq[2 * (size_t)k + 1] = pack_b(g1[k], g2[k]); }
    uint4* d_cellq = A.put(cq);
    int* d_gv = A.put(L.ghost_v); int* d_ge = A.put(L.ghost_e);
    uint32_t* d_eadj = A.put(L.eadj); uint32_t* d_voff = A.put(L.vslice_off); uint16_t* d_vadj = A.put(L.vadj);
    uint32_t* d_io = A.put(L.ell_ioff); uint32_t* d_wo = A.put(L.ell_woff); uint16_t* d_ids = A.put(L.ell_ids);
    std::vector<double> ellw(L.ell_pair.size(), 0.0);
    for (int p = 0; p < npatch; ++p) {
        const PatchHdr& h = L.hdr[p];
        const int nsl = (h.v_cnt + 31) >> 5;
        for (uint32_t k = L.ell_woff[h.vs_off]; k < L.ell_woff[h.vs_off + nsl]; ++k) {
            const uint32_t pr = L.ell_pair[k];
            if (pr == 0xFFFFFFFFu) continue;
            double v = area[h.cell_off + (pr & 0xFFFFu)];
            if ((pr >> 16) != 0xFFFFu) v += area[h.cell_off + (pr >> 16)];
            ellw[k] = v * (1.0 / 12.0);
        }
    }
    double* d_ellw = A.put(ellw);
    // Jacobi weights of M_q: diag_e = sum over the (<= 2) cells of 3 S - 4 g_slot
    std::vector<double> pinvq(L.ne, 0.0);
    for (int p = 0; p < npatch; ++p) {
        const PatchHdr& h = L.hdr[p];
        for (int le = 0; le < h.e_cnt; ++le) {
            const uint32_t adj = L.eadj[h.e_start + le];
            double dg = 0.0;
            for (uint32_t code : {adj & 0xFFFFu, adj >> 16}) {
                if (code == 0xFFFFu) continue;
                const int k = h.cell_off + (int)(code >> 2);
                const double S = g0[k] + g1[k] + g2[k], gs = (code & 3u) == 0 ? g0[k] : ((code & 3u) == 1 ? g1[k] : g2[k]);
                dg += 3.0 * S - 4.0 * gs;
            }
            pinvq[h.e_start + le] = 1.0 / dg;
        }
    }
    double* d_pinvq = A.put(pinvq);
    Ctl* ctl = A.get<Ctl>(1);
    double* partials = A.get<double>(16384);
    PeerComm pc{};
    pc.world = 1;
    // user <-> internal numbering
    auto to_new = [&](const double* user, const std::vector<int32_t>& new2old) {
        std::vector<double> v(new2old.size());
        for (size_t i = 0; i < new2old.size(); ++i) v[i] = user[new2old[i]];
        return v;
    };
    auto to_user = [&](const double* internal, const std::vector<int32_t>& new2old, double* user) {
        for (size_t i = 0; i < new2old.size(); ++i) user[new2old[i]] = internal[i];
    };
    MqPipeArgs cfg_q{}; EllPipeArgs cfg_p{}; EnergyPipeArgs cfg_e{}; RhsPipeArgs cfg_re{}, cfg_rv{};
    size_t sm_q = 0, sm_p = 0, sm_e = 0, sm_re = 0, sm_rv = 0;
    pipe_carve(L, cfg_q, cfg_p, sm_q, sm_p, &cfg_e, &sm_e, &cfg_re, &sm_re, &cfg_rv, &sm_rv);
    const bool is_q = what.rfind("mq", 0) == 0, is_p = what.rfind("mp", 0) == 0;
    if (is_q || is_p) {
        const std::vector<int32_t>& n2o = is_q ? L.e_new2old : L.v_new2old;
        const size_t n = n2o.size();
        double* d_b = A.put(to_new(in1, n2o));
        double* buf[4];
        for (auto& bptr : buf) bptr = A.get<double>(n);
        const double cheb_d = is_q ? 1.0 : 1.25, cheb_c = is_q ? 0.0 : 0.75;   // M_q: interval set below
        double lam_lo = 1e300, lam_hi = 0.0;
        if (is_q) {   // element-wise bounds of the Jacobi-scaled RT1 mass spectrum (what geom_k reduces on the GPU)
            for (int k = 0; k < npc; ++k) {
                const double S = g0[k] + g1[k] + g2[k];
                const double m00 = 3 * S - 4 * g0[k], m11 = 3 * S - 4 * g1[k], m22 = 3 * S - 4 * g2[k];
                const double m01 = S - 4 * g2[k], m02 = S - 4 * g1[k], m12 = S - 4 * g0[k];
                const double ea = m01 / std::sqrt(m00 * m11), eb = m02 / std::sqrt(m00 * m22), ec = m12 / std::sqrt(m11 * m22);
                const double pp = ea * ea + eb * eb + ec * ec, qq = 2.0 * ea * eb * ec;
                double lmin = 1.0, lmax = 1.0;
                if (pp > 1e-30) {
                    double arg = 0.5 * qq * std::pow(3.0 / pp, 1.5);
                    arg = std::min(1.0, std::max(-1.0, arg));
                    const double phi = std::acos(arg) / 3.0, amp = 2.0 * std::sqrt(pp / 3.0);
                    lmax = 1.0 + amp * std::cos(phi); lmin = 1.0 + amp * std::cos(phi + 2.0943951023931953);
                }
                lam_lo = std::min(lam_lo, lmin); lam_hi = std::max(lam_hi, lmax);
            }
        }
        const double cd = is_q ? 0.5 * (lam_lo + lam_hi) : cheb_d, cc = is_q ? 0.5 * (lam_hi - lam_lo) : cheb_c;
        if (what == "mq1") {
            MqPipeArgs a = cfg_q;
            a.npatch = npatch; a.hdr = d_hdr; a.cellq = d_cellq; a.ghost_e = d_ge; a.eadj = d_eadj; a.b = d_b; a.pinv = d_pinvq;
            for (int i = 0; i < 3; ++i) a.buf[i] = buf[i];
            a.cheb_d = cd; a.cheb_c = cc; a.tol2 = tol * tol; a.max_iter = 200; a.which = 1; a.partials = partials; a.ctl = ctl;
            phw_emu::launch(grid, NT, sm_q, [&] { solve_mq_pipe_k<NT, false, false>(a, pc); });
        } else if (what == "mq_lean") {
            MqSolveArgs a{};
            a.npatch = npatch; a.hdr = d_hdr; a.cellq = d_cellq; a.ghost_e = d_ge; a.eadj = d_eadj; a.b = d_b; a.pinv = d_pinvq;
            for (int i = 0; i < 3; ++i) a.buf[i] = buf[i];
            a.cheb_d = cd; a.cheb_c = cc; a.tol2 = tol * tol; a.max_iter = 200; a.which = 1; a.partials = partials; a.ctl = ctl;
            const size_t sm = ((size_t)((L.max_le + 1) & ~1) + 3 * (size_t)L.max_cells) * 8;
            phw_emu::launch(grid, NT, sm, [&] { solve_mq_coop_k<NT, 1, 2, 2, false>(a, pc); });
        } else if (what == "mp1") {
            EllPipeArgs a = cfg_p;
            a.npatch = npatch; a.hdr = d_hdr; a.ell_ioff = d_io; a.ell_woff = d_wo; a.ell_ids = d_ids; a.ell_w = d_ellw; a.ghost_v = d_gv; a.b = d_b;
            for (int i = 0; i < 3; ++i) a.buf[i] = buf[i];
            a.cheb_d = cd; a.cheb_c = cc; a.tol2 = tol * tol; a.max_iter = 200; a.which = 0; a.partials = partials; a.ctl = ctl;
            phw_emu::launch(grid, NT, sm_p, [&] { solve_ell_pipe_k<NT, 1, false>(a, pc); });
        } else if (what == "mp_lean") {
            EllSolveArgs a{};
            a.npatch = npatch; a.hdr = d_hdr; a.ell_ioff = d_io; a.ell_woff = d_wo; a.ell_ids = d_ids; a.ell_w = d_ellw; a.ghost_v = d_gv; a.b = d_b;
            for (int i = 0; i < 3; ++i) a.buf[i] = buf[i];
            a.cheb_d = cd; a.cheb_c = cc; a.tol2 = tol * tol; a.max_iter = 200; a.which = 0; a.partials = partials; a.ctl = ctl;
            const size_t sm = ((size_t)(L.max_lv + 31) / 32 + 4 + (size_t)((L.max_lv + 1) & ~1)) * 8;
            phw_emu::launch(grid, NT, sm, [&] { solve_ell_coop_k<NT, 1, false>(a, pc); });
        } else if (what == "mq_pcg" || what == "mp_pcg") {
            // conjugate-gradient mass solves (solve_pcg_coop_k) on the generic row passes: meshes of poor quality
            double* d_g0 = A.put(g0); double* d_g1 = A.put(g1); double* d_g2 = A.put(g2);
            std::vector<double> mdiag(L.nv, 0.0);               // diag of M_p: sum_K |K| / 6
            for (int p = 0; p < npatch; ++p) {
                const PatchHdr& h = L.hdr[p];
                for (int lc = 0; lc < h.n_vcells; ++lc) {
                    const uint2 rv = hv[h.cell_off + lc];
                    for (uint32_t lv : {rv.x & 0xFFFFu, rv.x >> 16, rv.y & 0xFFFFu})
                        if ((int)lv < h.v_cnt) mdiag[h.v_start + lv] += area[h.cell_off + lc] / 6.0;
                }
            }
            std::vector<double> pinvp(L.nv);
            for (int i = 0; i < (int)L.nv; ++i) pinvp[i] = 1.0 / mdiag[i];
            RowArgs a{};
            a.m.npatch = npatch; a.m.nv = (int)L.nv; a.m.ne = (int)L.ne; a.m.hdr = d_hdr; a.m.recv = d_recv; a.m.rece = d_rece; a.m.area = d_area;
            a.m.g0 = d_g0; a.m.g1 = d_g1; a.m.g2 = d_g2; a.m.ghost_v = d_gv; a.m.ghost_e = d_ge; a.m.eadj = d_eadj; a.m.vslice_off = d_voff; a.m.vadj = d_vadj;
            a.m.ell_ioff = d_io; a.m.ell_woff = d_wo; a.m.ell_ids = d_ids; a.m.ell_w = d_ellw;
            for (int i = 0; i < 3; ++i) { a.tp.b[i] = buf[i]; a.tq.b[i] = buf[i]; }
            a.b = d_b; a.pinv = is_q ? d_pinvq : A.put(pinvp); a.mdiag = is_q ? nullptr : A.put(mdiag);
            a.partials = partials; a.ctl = ctl; a.tol2 = tol * tol; a.max_iter = 2000; a.which = is_q ? 1 : 0; a.cM = 1.0; a.weight = 1.0;
            PcgArgs v{};
            v.r = A.get<double>(n); v.u = A.get<double>(n); v.b = a.b; v.pinv = a.pinv; v.n = (int)n;
            auto ev = [](int x) { return (size_t)((x + 1) & ~1); };
            const size_t soff = ((size_t)(L.max_lv + 31) / 32 + 4) / 2 + 1;
            if (is_q) phw_emu::launch(grid, NT, (ev(L.max_le) + 3 * (size_t)L.max_cells) * 8, [&] { solve_pcg_coop_k<1, NT, 1>(a, v); });
            else phw_emu::launch(grid, NT, (soff + ev(L.max_lv) + (size_t)L.max_cells) * 8, [&] { solve_pcg_coop_k<0, NT, 1>(a, v); });
        } else return fail(err, errlen, "unknown kernel " + what);
        const SolveCtl& sc = ctl->sc[is_q ? 1 : 0];
        if ((ctl->nonconv >> 32) != 0) return fail(err, errlen, "a stage did not arrive");
        if (ctl->nonconv != 0) return fail(err, errlen, "solve did not converge");
        to_user(buf[sc.cur & 3], n2o, out);
        if (iters) *iters = (int32_t)sc.iters;
        return 0;
    }
    if (what == "energy") {
        double* d_p = A.put(to_new(in1, L.v_new2old));
        double* d_q = A.put(to_new(in2, L.e_new2old));
        EnergyPipeArgs a = cfg_e;
        a.npatch = npatch; a.hdr = d_hdr; a.recv = d_recv; a.area = d_area; a.cellq = d_cellq; a.ghost_v = d_gv; a.ghost_e = d_ge;
        a.p = d_p; a.q = d_q; a.scale_p = 1.0 / 12.0; a.scale_q = 1.0; a.counter_id = 4; a.partials = partials; a.ctl = ctl;
        phw_emu::launch(grid, NT, sm_e, [&] { energy_pipe_k<NT>(a); });
        if ((ctl->nonconv >> 32) != 0) return fail(err, errlen, "a stage did not arrive");
        out[0] = ctl->red[0]; out[1] = ctl->red[1];
        return 0;
    }
    if (what == "rhs_e" || what == "rhs_v") {
        const bool vertex = what == "rhs_v";
        double* d_x = A.put(to_new(in1, vertex ? L.e_new2old : L.v_new2old));
        double* d_y = A.get<double>(vertex ? L.nv : L.ne);
        RhsPipeArgs a = vertex ? cfg_rv : cfg_re;
        a.npatch = npatch; a.hdr = d_hdr; a.recv = d_recv; a.rece = d_rece; a.ghost = vertex ? d_ge : d_gv; a.eadj = d_eadj;
        a.vslice_off = d_voff; a.vadj = d_vadj; a.x = d_x; a.y = d_y; a.cD = 1.0; a.ctl = ctl;
        if (vertex) phw_emu::launch(grid, NT, sm_rv, [&] { rhs_vertex_pipe_k<NT, 1>(a); });
        else phw_emu::launch(grid, NT, sm_re, [&] { rhs_edge_pipe_k<NT, 1>(a); });
        if ((ctl->nonconv >> 32) != 0) return fail(err, errlen, "a stage did not arrive");
        to_user(d_y, vertex ? L.v_new2old : L.e_new2old, out);
        return 0;
    }
    return fail(err, errlen, "unknown kernel " + what);
}

// ------------------------------------------------------------------------------------------------
// two ranks in one process (partitioned run): each rank has its own layout (external dofs numbered after the owned
// ones), iterate buffers, control block and mailbox; the PeerComm of a rank points at the other rank's buffers, exactly as
// phw_comm_setup wires them through CUDA IPC on the GPU.  Both ranks' kernels run concurrently (two grids of OS threads).
// what: "mq1" | "mp1" (pipelined solves with the exchange), "mq_lean" | "mp_lean".
// Per rank r: local mesh, external flags (vertices by local id, edges by canonical local edge id), right-hand side on all local
// dofs, send lists: n_send[r] owned dofs (local user ids send_loc[r]) and the ids the other rank uses for them (send_rem[r]).
// ------------------------------------------------------------------------------------------------
namespace {
struct Rank {
    Layout L;
    Arena A;
    std::vector<double> g0, g1, g2, area, pinvq;
    PatchHdr* d_hdr; uint4* d_cellq; int *d_gv, *d_ge; uint32_t *d_eadj, *d_io, *d_wo; uint16_t* d_ids; double *d_ellw, *d_pinvq, *d_b;
    double* buf[3];
    Ctl* ctl; double* partials; Mailbox* mbox;
    int *d_sloc, *d_srem; int n_send;
    PeerComm pc{};
    double lam_lo = 1e300, lam_hi = 0.0;
};
}  // namespace

extern "C" int emu_run2(const char* what_c, int32_t patch_cells, int32_t grid,
                        const int64_t* nv, const double* const* vtx, const int64_t* nc, const int32_t* const* cells,
                        const uint8_t* const* v_ext, const int64_t* ne, const uint8_t* const* e_ext,
                        const double* const* b_user, double* const* out_user,
                        const int32_t* n_send, const int32_t* const* send_loc, const int32_t* const* send_rem,
                        int32_t* iters, double tol, char* err, int errlen,
                        const uint8_t* const* cell_owned, const int32_t* const* btags, const double* const* in2_user) {
    const std::string what(what_c);
    constexpr int NT = 128;
    // single-pass kernels on the partitioned layouts (no exchange inside the kernel): "energy" (b_user = p, in2_user = q on the
    // local dofs, externals included; out_user[r][0..1] = this rank's share of p^T M_p p and q^T M_q q), "rhs_e" (out = D p on the
    // owned edges), "rhs_v" (out = D^T q on the owned vertices)
    const bool single = what == "energy" || what == "rhs_e" || what == "rhs_v";
    const bool is_q = what.rfind("mq", 0) == 0 || what == "rhs_v";
    Rank R[2];
    for (int r = 0; r < 2; ++r) {
        Rank& k = R[r];
        std::string e = k.L.build(nv[r], vtx[r], nc[r], cells[r], patch_cells, v_ext[r], ne[r], e_ext[r]);
        if (!e.empty()) return fail(err, errlen, "rank " + std::to_string(r) + ": " + e);
        if (btags && btags[r]) for (int64_t j = 0; j < k.L.nb; ++j) k.L.btag[j] = btags[r][j];
        k.L.boundary_set = true;
        if (cell_owned && cell_owned[r]) { k.L.cell_foreign.resize((size_t)k.L.nc); for (int64_t c = 0; c < k.L.nc; ++c) k.L.cell_foreign[c] = cell_owned[r][c] ? 0 : 1; }
        std::vector<CellRec> rec;
        k.L.make_records(false, false, rec);
        const Layout& L = k.L;
        const int npc = (int)L.pcell_user.size();
        k.area.resize(npc); k.g0.resize(npc); k.g1.resize(npc); k.g2.resize(npc);
        for (int c = 0; c < npc; ++c) cell_geometry(L, L.pcell_user[c], k.area[c], k.g0[c], k.g1[c], k.g2[c]);
        std::vector<uint4> cq(2 * (size_t)npc);
        for (int c = 0; c < npc; ++c) { cq[2 * (size_t)c] = pack_a(make_uint2(rec[c].z, rec[c].w), k.g0[c]); cq[2 * (size_t)c + 1] = pack_b(k.g1[c], k.g2[c]); }
        k.d_hdr = k.A.put(L.hdr); k.d_cellq = k.A.put(cq); k.d_gv = k.A.put(L.ghost_v); k.d_ge = k.A.put(L.ghost_e);
        k.d_eadj = k.A.put(L.eadj); k.d_io = k.A.put(L.ell_ioff); k.d_wo = k.A.put(L.ell_woff); k.d_ids = k.A.put(L.ell_ids);
        std::vector<double> ellw(L.ell_pair.size(), 0.0);
        k.pinvq.assign(L.ne, 0.0);
        for (int p = 0; p < L.npatch; ++p) {
            const PatchHdr& h = L.hdr[p];
            const int nsl = (h.v_cnt + 31) >> 5;
            for (uint32_t j = L.ell_woff[h.vs_off]; j < L.ell_woff[h.vs_off + nsl]; ++j) {
                const uint32_t pr = L.ell_pair[j];
                if (pr == 0xFFFFFFFFu) continue;
                double v = k.area[h.cell_off + (pr & 0xFFFFu)];
                if ((pr >> 16) != 0xFFFFu) v += k.area[h.cell_off + (pr >> 16)];
                ellw[j] = v * (1.0 / 12.0);
            }
            for (int le = 0; le < h.e_cnt; ++le) {
                const uint32_t adj = L.eadj[h.e_start + le];
                double dg = 0.0;
                for (uint32_t code : {adj & 0xFFFFu, adj >> 16}) {
                    if (code == 0xFFFFu) continue;
                    const int c = h.cell_off + (int)(code >> 2);
                    const double S = k.g0[c] + k.g1[c] + k.g2[c], gs = (code & 3u) == 0 ? k.g0[c] : ((code & 3u) == 1 ? k.g1[c] : k.g2[c]);
                    dg += 3.0 * S - 4.0 * gs;
                }
                k.pinvq[h.e_start + le] = 1.0 / dg;
            }
        }
        k.d_ellw = k.A.put(ellw); k.d_pinvq = k.A.put(k.pinvq);
        for (int c = 0; c < npc; ++c) {   // element bounds of the Jacobi-scaled RT1 mass spectrum
            const double S = k.g0[c] + k.g1[c] + k.g2[c];
            const double m00 = 3 * S - 4 * k.g0[c], m11 = 3 * S - 4 * k.g1[c], m22 = 3 * S - 4 * k.g2[c];
            const double m01 = S - 4 * k.g2[c], m02 = S - 4 * k.g1[c], m12 = S - 4 * k.g0[c];
            const double ea = m01 / std::sqrt(m00 * m11), eb = m02 / std::sqrt(m00 * m22), ec = m12 / std::sqrt(m11 * m22);
            const double pp = ea * ea + eb * eb + ec * ec, qq = 2.0 * ea * eb * ec;
            if (pp > 1e-30) {
                double arg = 0.5 * qq * std::pow(3.0 / pp, 1.5);
                arg = std::min(1.0, std::max(-1.0, arg));
                const double phi = std::acos(arg) / 3.0, amp = 2.0 * std::sqrt(pp / 3.0);
                k.lam_hi = std::max(k.lam_hi, 1.0 + amp * std::cos(phi)); k.lam_lo = std::min(k.lam_lo, 1.0 + amp * std::cos(phi + 2.0943951023931953));
            }
        }
        const std::vector<int32_t>& n2o = is_q ? L.e_new2old : L.v_new2old;
        std::vector<double> bn(n2o.size());
        for (size_t i = 0; i < n2o.size(); ++i) bn[i] = b_user[r][n2o[i]];
        k.d_b = k.A.put(bn);
        for (auto& bp : k.buf) bp = k.A.get<double>(n2o.size());
        k.ctl = k.A.get<Ctl>(1); k.partials = k.A.get<double>(16384); k.mbox = k.A.get<Mailbox>(1);
        const std::vector<int32_t>& o2n = is_q ? L.e_old2new : L.v_old2new;
        std::vector<int> sl(n_send[r]);
        for (int i = 0; i < n_send[r]; ++i) sl[i] = o2n[send_loc[r][i]];
        k.d_sloc = k.A.put(sl); k.n_send = n_send[r];
    }
    for (int r = 0; r < 2; ++r) {   // what phw_comm_setup does: neighbour buffers, remote positions, mailboxes
        Rank& k = R[r]; Rank& o = R[1 - r];
        const std::vector<int32_t>& o2n_other = is_q ? o.L.e_old2new : o.L.v_old2new;
        std::vector<int> sr(n_send[r]);
        for (int i = 0; i < n_send[r]; ++i) sr[i] = o2n_other[send_rem[r][i]];
        k.d_srem = k.A.put(sr);
        PeerComm& pc = k.pc;
        pc.rank = r; pc.world = 2; pc.n_nb = 1; pc.nb_rank[0] = 1 - r;
        for (int i = 0; i < 3; ++i) { pc.nb_tp[0][i] = is_q ? nullptr : o.buf[i]; pc.nb_tq[0][i] = is_q ? o.buf[i] : nullptr; }
        if (is_q) { pc.send_e_loc[0] = k.d_sloc; pc.send_e_rem[0] = k.d_srem; pc.n_send_e[0] = k.n_send; }
        else { pc.send_v_loc[0] = k.d_sloc; pc.send_v_rem[0] = k.d_srem; pc.n_send_v[0] = k.n_send; }
        pc.mbox[0] = R[0].mbox; pc.mbox[1] = R[1].mbox;
        pc.timeout_cycles = (long long)4e18;
        {   // send lists bucketed by owner patch (what phw_comm_setup builds for the pipelined solves)
            std::vector<std::vector<int32_t>> locs(1), rems(1);
            locs[0].assign(k.d_sloc, k.d_sloc + k.n_send); rems[0].assign(sr.begin(), sr.end());
            std::vector<int32_t> off; std::vector<PatchSendEntry> ent;
            bucket_send_lists(k.L.hdr, !is_q, 1, locs, rems, off, ent);
            const int f = is_q ? 1 : 0;
            pc.psend_off[f] = k.A.put(off);
            pc.psend_ent[f] = reinterpret_cast<const int4*>(k.A.put(ent));
        }
    }
    const double lam_lo = std::min(R[0].lam_lo, R[1].lam_lo), lam_hi = std::max(R[0].lam_hi, R[1].lam_hi);
    const double cd = is_q ? 0.5 * (lam_lo + lam_hi) : 1.25, cc = is_q ? 0.5 * (lam_hi - lam_lo) : 0.75;
    auto run_rank = [&](int r) {
        Rank& k = R[r];
        const Layout& L = k.L;
        const int g = std::max(1, std::min((int)grid, L.npatch));
        MqPipeArgs cfg_q{}; EllPipeArgs cfg_p{};
        size_t sm_q = 0, sm_p = 0;
        pipe_carve(L, cfg_q, cfg_p, sm_q, sm_p);
        if (what == "mq1" || what == "mq_lean") {
            if (what == "mq1") {
                MqPipeArgs a = cfg_q;
                a.npatch = L.npatch; a.hdr = k.d_hdr; a.cellq = k.d_cellq; a.ghost_e = k.d_ge; a.eadj = k.d_eadj; a.b = k.d_b; a.pinv = k.d_pinvq;
                for (int i = 0; i < 3; ++i) a.buf[i] = k.buf[i];
                a.cheb_d = cd; a.cheb_c = cc; a.tol2 = tol * tol; a.max_iter = 200; a.which = 1; a.partials = k.partials; a.ctl = k.ctl;
                phw_emu::launch(g, NT, sm_q, [&] { solve_mq_pipe_k<NT, false, true>(a, k.pc); });
            } else {
                MqSolveArgs a{};
                a.npatch = L.npatch; a.hdr = k.d_hdr; a.cellq = k.d_cellq; a.ghost_e = k.d_ge; a.eadj = k.d_eadj; a.b = k.d_b; a.pinv = k.d_pinvq;
                for (int i = 0; i < 3; ++i) a.buf[i] = k.buf[i];
                a.cheb_d = cd; a.cheb_c = cc; a.tol2 = tol * tol; a.max_iter = 200; a.which = 1; a.partials = k.partials; a.ctl = k.ctl;
                const size_t sm = ((size_t)((L.max_le + 1) & ~1) + 3 * (size_t)L.max_cells) * 8;
                phw_emu::launch(g, NT, sm, [&] { solve_mq_coop_k<NT, 1, 2, 2, true>(a, k.pc); });
            }
        } else {
            if (what == "mp1") {
                EllPipeArgs a = cfg_p;
                a.npatch = L.npatch; a.hdr = k.d_hdr; a.ell_ioff = k.d_io; a.ell_woff = k.d_wo; a.ell_ids = k.d_ids; a.ell_w = k.d_ellw; a.ghost_v = k.d_gv; a.b = k.d_b;
                for (int i = 0; i < 3; ++i) a.buf[i] = k.buf[i];
                a.cheb_d = cd; a.cheb_c = cc; a.tol2 = tol * tol; a.max_iter = 200; a.which = 0; a.partials = k.partials; a.ctl = k.ctl;
                phw_emu::launch(g, NT, sm_p, [&] { solve_ell_pipe_k<NT, 1, true>(a, k.pc); });
            } else {
                EllSolveArgs a{};
                a.npatch = L.npatch; a.hdr = k.d_hdr; a.ell_ioff = k.d_io; a.ell_woff = k.d_wo; a.ell_ids = k.d_ids; a.ell_w = k.d_ellw; a.ghost_v = k.d_gv; a.b = k.d_b;
                for (int i = 0; i < 3; ++i) a.buf[i] = k.buf[i];
                a.cheb_d = cd; a.cheb_c = cc; a.tol2 = tol * tol; a.max_iter = 200; a.which = 0; a.partials = k.partials; a.ctl = k.ctl;
                const size_t sm = ((size_t)(L.max_lv + 31) / 32 + 4 + (size_t)((L.max_lv + 1) & ~1)) * 8;
                phw_emu::launch(g, NT, sm, [&] { solve_ell_coop_k<NT, 1, true>(a, k.pc); });
            }
        }
    };
    if (single) {
        for (int r = 0; r < 2; ++r) {
            Rank& k = R[r];
            const Layout& L = k.L;
            const int g = std::max(1, std::min((int)grid, L.npatch));
            MqPipeArgs cfg_q{}; EllPipeArgs cfg_p{}; EnergyPipeArgs cfg_e{}; RhsPipeArgs cfg_re{}, cfg_rv{};
            size_t sm_q = 0, sm_p = 0, sm_e = 0, sm_re = 0, sm_rv = 0;
            pipe_carve(L, cfg_q, cfg_p, sm_q, sm_p, &cfg_e, &sm_e, &cfg_re, &sm_re, &cfg_rv, &sm_rv);
            std::vector<CellRec> rec;
            L.make_records(false, false, rec);
            std::vector<uint2> hv(rec.size()), he(rec.size());
            for (size_t c = 0; c < rec.size(); ++c) { hv[c] = make_uint2(rec[c].x, rec[c].y); he[c] = make_uint2(rec[c].z, rec[c].w); }
            uint2* d_recv = k.A.put(hv); uint2* d_rece = k.A.put(he);
            auto to_new = [&](const double* user, const std::vector<int32_t>& new2old) {
                std::vector<double> v(new2old.size());
                for (size_t i = 0; i < new2old.size(); ++i) v[i] = user[new2old[i]];
                return v;
            };
            if (what == "energy") {
                double* d_p = k.A.put(to_new(b_user[r], L.v_new2old));
                double* d_q = k.A.put(to_new(in2_user[r], L.e_new2old));
                double* d_area = k.A.put(k.area);
                EnergyPipeArgs a = cfg_e;
                a.npatch = L.npatch; a.hdr = k.d_hdr; a.recv = d_recv; a.area = d_area; a.cellq = k.d_cellq; a.ghost_v = k.d_gv; a.ghost_e = k.d_ge;
                a.p = d_p; a.q = d_q; a.scale_p = 1.0 / 12.0; a.scale_q = 1.0; a.counter_id = 4; a.partials = k.partials; a.ctl = k.ctl;
                phw_emu::launch(g, NT, sm_e, [&] { energy_pipe_k<NT>(a); });
                if ((k.ctl->nonconv >> 32) != 0) return fail(err, errlen, "a stage did not arrive");
                out_user[r][0] = k.ctl->red[0]; out_user[r][1] = k.ctl->red[1];
            } else {
                const bool vertex = what == "rhs_v";
                double* d_x = k.A.put(to_new(b_user[r], vertex ? L.e_new2old : L.v_new2old));
                double* d_y = k.A.get<double>(vertex ? L.nv : L.ne);
                uint32_t* d_voff = k.A.put(L.vslice_off); uint16_t* d_vadj = k.A.put(L.vadj);
                RhsPipeArgs a = vertex ? cfg_rv : cfg_re;
                a.npatch = L.npatch; a.hdr = k.d_hdr; a.recv = d_recv; a.rece = d_rece; a.ghost = vertex ? k.d_ge : k.d_gv; a.eadj = k.d_eadj;
                a.vslice_off = d_voff; a.vadj = d_vadj; a.x = d_x; a.y = d_y; a.cD = 1.0; a.ctl = k.ctl;
                if (vertex) phw_emu::launch(g, NT, sm_rv, [&] { rhs_vertex_pipe_k<NT, 1>(a); });
                else phw_emu::launch(g, NT, sm_re, [&] { rhs_edge_pipe_k<NT, 1>(a); });
                if ((k.ctl->nonconv >> 32) != 0) return fail(err, errlen, "a stage did not arrive");
                const std::vector<int32_t>& n2o = vertex ? L.v_new2old : L.e_new2old;
                for (size_t i = 0; i < n2o.size(); ++i) out_user[r][n2o[i]] = d_y[i];
            }
        }
        return 0;
    }
    if (what != "mq1" && what != "mq_lean" && what != "mp1" && what != "mp_lean") return fail(err, errlen, "unknown kernel " + what);
    std::thread t1(run_rank, 1);
    run_rank(0);
    t1.join();
    for (int r = 0; r < 2; ++r) {
        Rank& k = R[r];
        if (k.ctl->comm_dead) return fail(err, errlen, "exchange timed out");
        if ((k.ctl->nonconv >> 32) != 0) return fail(err, errlen, "a stage did not arrive");
        if (k.ctl->nonconv != 0) return fail(err, errlen, "solve did not converge");
        const SolveCtl& sc = k.ctl->sc[is_q ? 1 : 0];
        const std::vector<int32_t>& n2o = is_q ? k.L.e_new2old : k.L.v_new2old;
        for (size_t i = 0; i < n2o.size(); ++i) out_user[r][n2o[i]] = k.buf[sc.cur][i];
        if (iters) iters[r] = (int32_t)sc.iters;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// batched-sweep kernels (csrc/kernels_sweep.cuh): ONE assembled CSR operator, B cases with the case index fastest, CPL = 1 / 2 / 4
// cases per lane, both row mappings.  what: "store" (out[n_rows][B] = alpha cscale[c] A x), "dot" (out[B] = cscale[c] x^T A x per
// case, through csr_dot_b_k + dot_b_finish_k), "solve" (out[n_rows][B] = A^-1 b by solve_coop_csr_b_k, x0 = 0; pinv = 1 / diag).
// ------------------------------------------------------------------------------------------------
template <int CPL>
static int run_csr_cpl(const std::string& what, const Csr& A, int B, int mode, int n_groups, int grid, const double* d_x, const double* d_cs,
                       double alpha, double* d_y, double* d_part, Ctl* ctl, double* partials, const double* d_pinv, double* const* buf,
                       double lam_lo, double lam_hi, double tol) {
    constexpr int NT = 256;
    if (what == "store") {
        phw_emu::launch(grid, NT, 0, [&] { csr_store_b_k<CPL>(A, B, n_groups, mode, d_x, alpha, d_cs, d_y); });
    } else if (what == "dot") {
        phw_emu::launch(grid, NT, 0, [&] { csr_dot_b_k<CPL>(A, B, n_groups, mode, d_x, d_part); });
        phw_emu::launch((B + 127) / 128, 128, 0, [&] { dot_b_finish_k(B, n_groups * SWEEP_WPB, d_part, alpha, d_cs, ctl, 0); });
    } else {
        CsrSolveBArgs a{};
        a.A = A; a.B = B; a.n_row_groups = n_groups; a.mode = mode;
        for (int i = 0; i < 3; ++i) a.tb.b[i] = buf[i];
        a.b = d_x; a.pinv = d_pinv; a.cheb_d = 0.5 * (lam_lo + lam_hi); a.cheb_c = 0.5 * (lam_hi - lam_lo);
        a.tol2 = tol * tol; a.max_iter = 400; a.which = 1; a.partials = partials; a.ctl = ctl;
        phw_emu::launch(grid, NT, 0, [&] { solve_coop_csr_b_k<CPL>(a); });
    }
    return 0;
}

extern "C" int emu_run_csr(const char* what_c, int32_t n_rows, int32_t n_cols, const int32_t* indptr, const int32_t* idx, const double* val,
                           int32_t B, int32_t cpl, int32_t mode, int32_t n_groups, const double* x_in, const double* cscale, double alpha,
                           double lam_lo, double lam_hi, double tol, double* out, int32_t* iters, char* err, int errlen) {
    const std::string what(what_c);
    if (B % cpl != 0) return fail(err, errlen, "the number of cases must be a multiple of the cases per lane");
    Arena A;
    const int nnz = indptr[n_rows];
    Csr M{};
    M.n_rows = n_rows; M.lanes = 8;
    M.indptr = A.put(std::vector<int>(indptr, indptr + n_rows + 1));
    M.idx = A.put(std::vector<int>(idx, idx + nnz));
    M.val = A.put(std::vector<double>(val, val + nnz));
    const int nchunk = (B + 32 * cpl - 1) / (32 * cpl);
    const int grid = nchunk * n_groups;
    const size_t n_in = what == "solve" ? (size_t)n_rows * B : (size_t)n_cols * B;
    double* d_x = A.put(std::vector<double>(x_in, x_in + n_in));
    double* d_cs = cscale ? A.put(std::vector<double>(cscale, cscale + B)) : nullptr;
    double* d_y = A.get<double>((size_t)n_rows * B);
    double* d_part = A.get<double>((size_t)n_groups * SWEEP_WPB * B);
    Ctl* ctl = A.get<Ctl>((size_t)B);
    double* partials = A.get<double>(4 * (size_t)grid + 16);
    std::vector<double> pinv(n_rows, 1.0);
    for (int r = 0; r < n_rows; ++r)
        for (int j = indptr[r]; j < indptr[r + 1]; ++j) if (idx[j] == r) pinv[r] = 1.0 / val[j];
    double* d_pinv = A.put(pinv);
    double* buf[3];
    for (auto& b : buf) b = A.get<double>((size_t)n_rows * B);
    int rc;
    if (cpl == 4) rc = run_csr_cpl<4>(what, M, B, mode, n_groups, grid, d_x, d_cs, alpha, d_y, d_part, ctl, partials, d_pinv, buf, lam_lo, lam_hi, tol);
    else if (cpl == 2) rc = run_csr_cpl<2>(what, M, B, mode, n_groups, grid, d_x, d_cs, alpha, d_y, d_part, ctl, partials, d_pinv, buf, lam_lo, lam_hi, tol);
    else rc = run_csr_cpl<1>(what, M, B, mode, n_groups, grid, d_x, d_cs, alpha, d_y, d_part, ctl, partials, d_pinv, buf, lam_lo, lam_hi, tol);
    if (rc) return rc;
    if (what == "store") std::memcpy(out, d_y, (size_t)n_rows * B * sizeof(double));
    else if (what == "dot") { for (int c = 0; c < B; ++c) out[c] = ctl[c].red[0]; }
    else {
        if (ctl->nonconv != 0) return fail(err, errlen, "solve did not converge");
        std::memcpy(out, buf[ctl->sc[1].cur % 3], (size_t)n_rows * B * sizeof(double));
        if (iters) *iters = (int32_t)ctl->sc[1].iters;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// coupled block solve of the implicit schemes on assembled CSR rows (solve_block_csr_k, higher-order pairs):
//     M_p v_p + cDv D^T v_q = b_p ,   M_q v_q + cDe D v_p = b_q ;   Jacobi weights = 1 / diag, ellipse (cheb_d, cheb_c) from the caller
// ------------------------------------------------------------------------------------------------
extern "C" int emu_run_block(int32_t n_p, int32_t n_q, const int32_t* const* indptr4, const int32_t* const* idx4, const double* const* val4,
                             double cDv, double cDe, double wv, double we, double cheb_d, double cheb_c, double tol, int32_t grid,
                             const double* bp, const double* bq, double* xp, double* xq, int32_t* iters, char* err, int errlen) {
    Arena A;
    Csr M[4];     // M_p [n_p], M_q [n_q], D [n_q x n_p], D^T [n_p x n_q]
    const int rows[4] = {n_p, n_q, n_q, n_p};
    for (int m = 0; m < 4; ++m) {
        const int nnz = indptr4[m][rows[m]];
        M[m].n_rows = rows[m]; M[m].lanes = 8;
        M[m].indptr = A.put(std::vector<int>(indptr4[m], indptr4[m] + rows[m] + 1));
        M[m].idx = A.put(std::vector<int>(idx4[m], idx4[m] + nnz));
        M[m].val = A.put(std::vector<double>(val4[m], val4[m] + nnz));
    }
    auto inv_diag = [&](int m) {
        std::vector<double> d(rows[m], 1.0);
        for (int r = 0; r < rows[m]; ++r)
            for (int j = indptr4[m][r]; j < indptr4[m][r + 1]; ++j) if (idx4[m][j] == r) d[r] = 1.0 / val4[m][j];
        return d;
    };
    CsrBlockArgs a{};
    a.Mp = M[0]; a.Mq = M[1]; a.D = M[2]; a.DT = M[3];
    for (int i = 0; i < 3; ++i) { a.tp.b[i] = A.get<double>(n_p); a.tq.b[i] = A.get<double>(n_q); }
    a.bp = A.put(std::vector<double>(bp, bp + n_p)); a.bq = A.put(std::vector<double>(bq, bq + n_q));
    a.pinv_p = A.put(inv_diag(0)); a.pinv_q = A.put(inv_diag(1));
    a.cDv = cDv; a.cDe = cDe; a.wv = wv; a.we = we; a.cheb_d = cheb_d; a.cheb_c = cheb_c; a.tol2 = tol * tol; a.max_iter = 2000;
    a.partials = A.get<double>(4 * (size_t)grid + 16);
    Ctl* ctl = A.get<Ctl>(1);
    a.ctl = ctl;
    phw_emu::launch(grid, 256, 0, [&] { solve_block_csr_k(a); });
    if (ctl->nonconv != 0) return fail(err, errlen, "block solve did not converge");
    const int cur = ctl->sc[2].cur % 3;
    std::memcpy(xp, a.tp.b[cur], (size_t)n_p * sizeof(double));
    std::memcpy(xq, a.tq.b[cur], (size_t)n_q * sizeof(double));
    if (iters) *iters = (int32_t)ctl->sc[2].iters;
    return 0;
}
