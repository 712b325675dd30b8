// pipe_carve.hpp - shared-memory carve-ups of the bulk-copy pipelined kernels (kernels_pipe.cuh), computed on the
// host from the largest patch of a layout.  Used by phwave.cu (solver setup, phw_mesh_pipe_smem) and by the host-side emulator of
// tests/emu, so that both run the kernels with the same offsets.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>
#include "layout.hpp"

namespace phw {

// shared-memory carve-up of the bulk-copy pipelined mass solves (kernels_pipe.cuh) for the largest patch of a layout:
// offsets inside a stage, stage size, and the dynamic shared memory each kernel needs (two stages + scratch)
constexpr size_t PIPE_SMEM_LIMIT = 227 * 1024 - 1024;   // 227 KB per CTA minus the kernels' static shared memory
inline void pipe_carve(const Layout& L, MqPipeArgs& cq, EllPipeArgs& cp, size_t& smem_pq, size_t& smem_pp, EnergyPipeArgs* ce = nullptr, size_t* smem_pe = nullptr,
                       RhsPipeArgs* re = nullptr, size_t* smem_re = nullptr, RhsPipeArgs* rv = nullptr, size_t* smem_rv = nullptr) {
    int max_ec = 0, max_ecells = 0, max_ge = 0, max_vc = 0, max_gv = 0, max_wn = 0, max_in = 0, max_own = 0, max_adj = 0;
    for (int32_t ip = 0; ip < L.npatch_rows; ++ip) {
        const PatchHdr& h = L.hdr[ip];
        max_ec = std::max(max_ec, h.e_cnt); max_ecells = std::max(max_ecells, h.n_ecells); max_ge = std::max(max_ge, h.ge_cnt);
        max_vc = std::max(max_vc, h.v_cnt); max_gv = std::max(max_gv, h.gv_cnt); max_wn = std::max(max_wn, h.ell_wn); max_in = std::max(max_in, h.ell_in);
        max_own = std::max(max_own, h.n_own); max_adj = std::max(max_adj, h.adj_cnt);
    }
    auto r16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    {   // M_q: stage = [x_k local | cell records | b | pinv | x_{k-1} | eadj], then the cell contributions and two ghost lists
        size_t o = r16(((size_t)L.max_le + 2) * 8);
        cq.off_cells = (uint32_t)o; o += (size_t)max_ecells * 32;
        cq.off_b = (uint32_t)o; o += r16(((size_t)max_ec + 2) * 8);
        cq.off_pinv = (uint32_t)o; o += r16(((size_t)max_ec + 2) * 8);
        cq.off_xprev = (uint32_t)o; o += r16(((size_t)max_ec + 2) * 8);
        cq.off_eadj = (uint32_t)o; o += r16(((size_t)max_ec + 8) * 4);
        cq.stage_bytes = (uint32_t)((o + 127) & ~(size_t)127);
        cq.off_sc3 = 2 * cq.stage_bytes;
        cq.off_gidx = cq.off_sc3 + (uint32_t)r16((size_t)max_ecells * 24);
        cq.gidx_bytes = (uint32_t)r16(((size_t)max_ge + 8) * 4);
        smem_pq = (size_t)cq.off_gidx + 2 * cq.gidx_bytes;
    }
    {   // M_p: stage = [slice offsets ids | slice offsets weights | x_k local | b | x_{k-1} | weights | ids], then two ghost lists
        const size_t nslw = r16(((size_t)(max_vc + 31) / 32 + 2) * 4);
        size_t o = nslw;
        cp.off_wo = (uint32_t)o; o += nslw;
        cp.off_x = (uint32_t)o; o += r16(((size_t)L.max_lv + 2) * 8);
        cp.off_b = (uint32_t)o; o += r16(((size_t)max_vc + 2) * 8);
        cp.off_xprev = (uint32_t)o; o += r16(((size_t)max_vc + 2) * 8);
        cp.off_w = (uint32_t)o; o += r16((size_t)max_wn * 8);
        cp.off_ids = (uint32_t)o; o += r16((size_t)max_in * 2);
        cp.stage_bytes = (uint32_t)((o + 127) & ~(size_t)127);
        cp.off_gidx = 2 * cp.stage_bytes;
        cp.gidx_bytes = (uint32_t)r16(((size_t)max_gv + 8) * 4);
        smem_pp = (size_t)cp.off_gidx + 2 * cp.gidx_bytes;
    }
    if (ce && smem_pe) {   // fused energy: stage = [p local | q local | vertex-id records | areas | edge records], two ghost lists per field
        size_t o = r16(((size_t)L.max_lv + 2) * 8);
        ce->off_q = (uint32_t)o; o += r16(((size_t)L.max_le + 2) * 8);
        ce->off_recv = (uint32_t)o; o += r16(((size_t)max_own + 2) * 8);
        ce->off_area = (uint32_t)o; o += r16(((size_t)max_own + 2) * 8);
        ce->off_cells = (uint32_t)o; o += (size_t)max_own * 32;
        ce->stage_bytes = (uint32_t)((o + 127) & ~(size_t)127);
        ce->off_gv = 2 * ce->stage_bytes;
        ce->gv_bytes = (uint32_t)r16(((size_t)max_gv + 8) * 4);
        ce->off_ge = ce->off_gv + 2 * ce->gv_bytes;
        ce->ge_bytes = (uint32_t)r16(((size_t)max_ge + 8) * 4);
        *smem_pe = (size_t)ce->off_ge + 2 * ce->ge_bytes;
    }
    if (re && smem_re) {   // edge-row right-hand side: stage = [x_p local | vertex-id records | edge-id records | eadj], then contributions, ghost lists
        size_t o = r16(((size_t)L.max_lv + 2) * 8);
        re->off_a = (uint32_t)o; o += r16(((size_t)max_ecells + 2) * 8);
        re->off_b = (uint32_t)o; o += r16(((size_t)max_ecells + 2) * 8);
        re->off_c = (uint32_t)o; o += r16(((size_t)max_ec + 8) * 4);
        re->stage_bytes = (uint32_t)((o + 127) & ~(size_t)127);
        re->off_sc3 = 2 * re->stage_bytes;
        re->off_gidx = re->off_sc3 + (uint32_t)r16((size_t)max_ecells * 24);
        re->gidx_bytes = (uint32_t)r16(((size_t)max_gv + 8) * 4);
        *smem_re = (size_t)re->off_gidx + 2 * re->gidx_bytes;
    }
    if (rv && smem_rv) {   // vertex-row right-hand side: stage = [slice offsets | x_q local | edge-id records | adjacency codes]
        size_t o = r16(((size_t)(max_vc + 31) / 32 + 2) * 4);
        rv->off_a = (uint32_t)o; o += r16(((size_t)L.max_le + 2) * 8);
        rv->off_b = (uint32_t)o; o += r16(((size_t)L.max_cells + 2) * 8);
        rv->off_c = (uint32_t)o; o += r16((size_t)max_adj * 2);
        rv->stage_bytes = (uint32_t)((o + 127) & ~(size_t)127);
        rv->off_sc3 = 2 * rv->stage_bytes;
        rv->off_gidx = rv->off_sc3 + (uint32_t)r16((size_t)L.max_cells * 24);
        rv->gidx_bytes = (uint32_t)r16(((size_t)max_ge + 8) * 4);
        *smem_rv = (size_t)rv->off_gidx + 2 * rv->gidx_bytes;
    }
}


// partitioned runs: the send lists of a rank (per neighbour k: rows to push, given by their new local ids, and their positions in
// the neighbour's vectors) bucketed by owner patch - PeerComm::psend_off / psend_ent of the pipelined solves (kernels_pipe.cuh).
// pstart[p] = first owned row of patch p (hdr[p].v_start or e_start); rows are owned, i.e. < pstart[npatch - 1] + count.
struct PatchSendEntry { int32_t loc, nb, rem, pad; };
inline void bucket_send_lists(const std::vector<PatchHdr>& hdr, bool vertex, int n_nb, const std::vector<std::vector<int32_t>>& loc_new,
                              const std::vector<std::vector<int32_t>>& rem, std::vector<int32_t>& off, std::vector<PatchSendEntry>& ent) {
    const int np = (int)hdr.size();
    std::vector<int32_t> start(np);
    for (int p = 0; p < np; ++p) start[p] = vertex ? hdr[p].v_start : hdr[p].e_start;
    auto patch_of = [&](int32_t row) { return (int)(std::upper_bound(start.begin(), start.end(), row) - start.begin()) - 1; };
    off.assign(np + 1, 0);
    for (int k = 0; k < n_nb; ++k)
        for (int32_t r : loc_new[k]) off[patch_of(r) + 1]++;
    for (int p = 0; p < np; ++p) off[p + 1] += off[p];
    ent.assign(off[np], PatchSendEntry{0, 0, 0, 0});
    std::vector<int32_t> pos(off.begin(), off.end() - 1);
    for (int k = 0; k < n_nb; ++k)
        for (size_t i = 0; i < loc_new[k].size(); ++i) ent[pos[patch_of(loc_new[k][i])]++] = PatchSendEntry{loc_new[k][i], k, rem[k][i], 0};
}

}  // namespace phw
