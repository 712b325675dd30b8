// layout.cpp - see layout.hpp.  Pure host C++ (compiled by nvcc's host compiler), no CUDA calls.
#include "layout.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>

namespace phw {

namespace {

// Hilbert curve index of a point on a 2^16 x 2^16 lattice.
inline uint32_t hilbert_d(uint32_t x, uint32_t y) {
    uint32_t d = 0;
    for (uint32_t s = 1u << 15; s > 0; s >>= 1) {
        uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += s * s * ((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) { x = 65535u - x; y = 65535u - y; }
            std::swap(x, y);
        }
    }
    return d;
}

// stable LSD radix sort of (key32, value32) pairs, two 16-bit passes
void radix_sort_pairs(std::vector<uint32_t>& key, std::vector<int32_t>& val) {
    const size_t n = key.size();
    std::vector<uint32_t> k2(n);
    std::vector<int32_t> v2(n);
    std::vector<size_t> hist(65537);
    for (int pass = 0; pass < 2; ++pass) {
        const int sh = 16 * pass;
        std::fill(hist.begin(), hist.end(), 0);
        for (size_t i = 0; i < n; ++i) hist[((key[i] >> sh) & 0xFFFFu) + 1]++;
        for (size_t b = 0; b < 65536; ++b) hist[b + 1] += hist[b];
        for (size_t i = 0; i < n; ++i) {
            size_t dst = hist[(key[i] >> sh) & 0xFFFFu]++;
            k2[dst] = key[i];
            v2[dst] = val[i];
        }
        key.swap(k2);
        val.swap(v2);
    }
}

}  // namespace

std::string Layout::build(int64_t nv_, const double* vxy, int64_t nc_, const int32_t* cells_, int32_t patch_cells_,
                          const uint8_t* v_ext, int64_t ne_expected, const uint8_t* e_ext,
                          const int32_t* cell_group, int32_t n_groups_) {
    if (nv_ <= 0 || nc_ <= 0 || !vxy || !cells_) return "empty mesh";
    if (nv_ > (int64_t)1 << 30 || nc_ > (int64_t)1 << 30) return "mesh too large for int32 indexing";
    if (patch_cells_ < 16 || patch_cells_ > 8192) return "patch_cells must be in [16, 8192]";
    nv = nv_; nc = nc_; patch_cells = patch_cells_;
    n_groups = (cell_group && n_groups_ > 1) ? n_groups_ : 1;
    cell_group_v.clear();
    if (n_groups > 1) {
        cell_group_v.assign(cell_group, cell_group + nc);
        for (int64_t c = 0; c < nc; ++c)
            if (cell_group[c] < 0 || cell_group[c] >= n_groups) return "cell group id out of range";
    }
    vtx.assign(vxy, vxy + 2 * nv);
    cells.assign(cells_, cells_ + 3 * nc);
    // ---- 1. orient cells counter-clockwise
    for (int64_t c = 0; c < nc; ++c) {
        int32_t* v = &cells[3 * c];
        for (int i = 0; i < 3; ++i)
            if (v[i] < 0 || v[i] >= nv) return "cell vertex index out of range";
        const double* a = &vtx[2 * (int64_t)v[0]]; const double* b = &vtx[2 * (int64_t)v[1]]; const double* d = &vtx[2 * (int64_t)v[2]];
        double area2 = (b[0] - a[0]) * (d[1] - a[1]) - (b[1] - a[1]) * (d[0] - a[0]);
        if (area2 == 0.0) return "degenerate cell";
        if (area2 < 0) std::swap(v[1], v[2]);
    }
    // ---- 2. edges: unique (lo,hi) pairs in lexicographic order; local edge i joins v[i+1], v[i+2]
    {
        std::vector<int64_t> start(nv + 1, 0);
        for (int64_t h = 0; h < 3 * nc; ++h) {
            int64_t c = h / 3; int i = (int)(h % 3);
            int32_t a = cells[3 * c + (i + 1) % 3], b = cells[3 * c + (i + 2) % 3];
            start[std::min(a, b) + 1]++;
        }
        for (int64_t v = 0; v < nv; ++v) start[v + 1] += start[v];
        std::vector<int64_t> pos(start.begin(), start.end() - 1);
        std::vector<int32_t> bhi(3 * nc);
        std::vector<int64_t> bhe(3 * nc);
        for (int64_t h = 0; h < 3 * nc; ++h) {
            int64_t c = h / 3; int i = (int)(h % 3);
            int32_t a = cells[3 * c + (i + 1) % 3], b = cells[3 * c + (i + 2) % 3];
            int64_t p = pos[std::min(a, b)]++;
            bhi[p] = std::max(a, b);
            bhe[p] = h;
        }
        cell_edges.assign(3 * nc, -1);
        cell_sign.assign(3 * nc, 0);
        edges.clear();
        edges.reserve(3 * nc + 2 * nv);
        std::vector<int32_t> ecount;
        ecount.reserve(3 * nc / 2 + nv);
        std::vector<std::pair<int32_t, int64_t>> tmp;
        for (int64_t lo = 0; lo < nv; ++lo) {
            int64_t s = start[lo], e = start[lo + 1];
            if (s == e) continue;
            tmp.clear();
            for (int64_t k = s; k < e; ++k) tmp.emplace_back(bhi[k], bhe[k]);
            std::sort(tmp.begin(), tmp.end());
            int32_t prev = -1;
            for (auto& t : tmp) {
                if (t.first != prev) {
                    edges.push_back((int32_t)lo);
                    edges.push_back(t.first);
                    ecount.push_back(0);
                    prev = t.first;
                }
                int32_t eid = (int32_t)(edges.size() / 2 - 1);
                int64_t h = t.second;
                cell_edges[h] = eid;
                int64_t c = h / 3; int i = (int)(h % 3);
                int32_t a = cells[3 * c + (i + 1) % 3], b = cells[3 * c + (i + 2) % 3];
                cell_sign[h] = (a < b) ? 1 : -1;
                if (++ecount.back() > 2) return "non-manifold edge (more than two adjacent cells)";
            }
        }
        ne = (int64_t)edges.size() / 2;
        edges.shrink_to_fit();
        // boundary edges: one adjacent cell
        bedge.clear();
        for (int64_t e = 0; e < ne; ++e)
            if (ecount[e] == 1) bedge.push_back((int32_t)e);
        nb = (int64_t)bedge.size();
        std::vector<int8_t> esign(ne, 0);
        for (int64_t h = 0; h < 3 * nc; ++h) esign[cell_edges[h]] += cell_sign[h];
        bsign.resize(nb);
        for (int64_t k = 0; k < nb; ++k) bsign[k] = esign[bedge[k]];
        btag.assign(nb, 4);
        bweight.assign(nb, 1.0);
        boundary_set = false;
    }
    // ---- 3. Hilbert order of cell centroids -> patches
    std::vector<int32_t> order(nc);
    {
        double xmin = std::numeric_limits<double>::max(), xmax = -xmin, ymin = xmin, ymax = -xmin;
        for (int64_t v = 0; v < nv; ++v) {
            xmin = std::min(xmin, vtx[2 * v]); xmax = std::max(xmax, vtx[2 * v]);
            ymin = std::min(ymin, vtx[2 * v + 1]); ymax = std::max(ymax, vtx[2 * v + 1]);
        }
        double ext = std::max(xmax - xmin, ymax - ymin);
        if (!(ext > 0)) return "mesh has zero extent";
        double sc = 65535.0 / ext;
        std::vector<uint32_t> key(nc);
        for (int64_t c = 0; c < nc; ++c) {
            const int32_t* v = &cells[3 * c];
            double cx = (vtx[2 * (int64_t)v[0]] + vtx[2 * (int64_t)v[1]] + vtx[2 * (int64_t)v[2]]) / 3.0;
            double cy = (vtx[2 * (int64_t)v[0] + 1] + vtx[2 * (int64_t)v[1] + 1] + vtx[2 * (int64_t)v[2] + 1]) / 3.0;
            uint32_t qx = (uint32_t)std::min(65535.0, std::max(0.0, (cx - xmin) * sc));
            uint32_t qy = (uint32_t)std::min(65535.0, std::max(0.0, (cy - ymin) * sc));
            key[c] = hilbert_d(qx, qy);
            order[c] = (int32_t)c;
        }
        radix_sort_pairs(key, order);
    }
    // batched runs: group-major order (stable), patches never straddle two groups
    std::vector<int64_t> gcell_off(n_groups + 1, 0);
    if (n_groups > 1) {
        for (int64_t c = 0; c < nc; ++c) gcell_off[cell_group_v[c] + 1]++;
        for (int32_t g = 0; g < n_groups; ++g) gcell_off[g + 1] += gcell_off[g];
        std::vector<int64_t> pos(gcell_off.begin(), gcell_off.end() - 1);
        std::vector<int32_t> o2(nc);
        for (int64_t k = 0; k < nc; ++k) o2[pos[cell_group_v[order[k]]]++] = order[k];
        order.swap(o2);
    } else {
        gcell_off[1] = nc;
    }
    std::vector<int64_t> own_off(1, 0);
    std::vector<int32_t> patch_group;
    group_patch_off.assign(n_groups + 1, 0);
    for (int32_t g = 0; g < n_groups; ++g) {
        const int64_t ng = gcell_off[g + 1] - gcell_off[g];
        const int32_t npg = (int32_t)((ng + patch_cells - 1) / patch_cells);
        for (int32_t p = 1; p <= npg; ++p) { own_off.push_back(gcell_off[g] + (int64_t)p * ng / npg); patch_group.push_back(g); }
        group_patch_off[g + 1] = group_patch_off[g] + npg;
    }
    npatch = (int32_t)patch_group.size();
    cell_patch.assign(nc, 0);
    for (int32_t p = 0; p < npatch; ++p)
        for (int64_t k = own_off[p]; k < own_off[p + 1]; ++k) cell_patch[order[k]] = p;
    // ---- 4. dof ownership = lowest adjacent patch; renumber so owned ranges are contiguous
    auto assign_owners = [&]() {
        v_owner.assign(nv, std::numeric_limits<int32_t>::max());
        e_owner.assign(ne, std::numeric_limits<int32_t>::max());
        for (int64_t c = 0; c < nc; ++c) {
            int32_t p = cell_patch[c];
            for (int i = 0; i < 3; ++i) {
                int32_t& ov = v_owner[cells[3 * c + i]]; ov = std::min(ov, p);
                int32_t& oe = e_owner[cell_edges[3 * c + i]]; oe = std::min(oe, p);
            }
        }
    };
    assign_owners();
    npatch_rows = npatch;
    if ((v_ext || e_ext) && n_groups == 1) {
        // partitioned runs: a patch made of halo cells only (cells of a neighbour rank: every dof is external or belongs to an
        // earlier patch) owns no row.  Such patches move behind the others and are never processed (npatch_rows): no kernel has
        // work for them, and their ghost lists (every dof of ~1000 cells) must not size the shared-memory stages.
        std::vector<uint8_t> owns(npatch, 0);
        for (int64_t v = 0; v < nv; ++v) if (!(v_ext && v_ext[v])) owns[v_owner[v]] = 1;
        for (int64_t e = 0; e < ne; ++e) if (!(e_ext && e_ext[e])) owns[e_owner[e]] = 1;
        int32_t n_rows = 0;
        for (int32_t p = 0; p < npatch; ++p) n_rows += owns[p];
        if (n_rows < npatch && n_rows > 0) {
            std::vector<int32_t> o2; o2.reserve(nc);
            std::vector<int64_t> off2(1, 0);
            for (int pass = 0; pass < 2; ++pass)
                for (int32_t p = 0; p < npatch; ++p)
                    if ((owns[p] != 0) == (pass == 0)) {
                        for (int64_t k = own_off[p]; k < own_off[p + 1]; ++k) o2.push_back(order[k]);
                        off2.push_back((int64_t)o2.size());
                    }
            order.swap(o2); own_off.swap(off2);
            for (int32_t p = 0; p < npatch; ++p)
                for (int64_t k = own_off[p]; k < own_off[p + 1]; ++k) cell_patch[order[k]] = p;
            assign_owners();      // the relative order of the patches that own rows is unchanged, so is the ownership
            npatch_rows = n_rows;
        }
    }
    for (int64_t v = 0; v < nv; ++v)
        if (v_owner[v] == std::numeric_limits<int32_t>::max()) return "unreferenced vertex in mesh";
    if (e_ext && ne_expected != ne) return "edge_external has the wrong length (edge numbering mismatch)";
    // dofs owned by another rank belong to the virtual patch `npatch`: numbered last, never a row here
    if (v_ext) for (int64_t v = 0; v < nv; ++v) if (v_ext[v]) v_owner[v] = npatch;
    if (e_ext) for (int64_t e = 0; e < ne; ++e) if (e_ext[e]) e_owner[e] = npatch;
    v_degree.assign(nv, 0);
    for (int64_t h = 0; h < 3 * nc; ++h) v_degree[cells[h]]++;
    // new id = rank by (owner, key): edges by first touch along the owner patch's cell order (spatially coherent local ids and
    // ghost lists), vertices by descending degree (uniform 32-row slices keep the sliced-ELL structures tight: the ELL mass rows
    // then carry no padding on meshes with mixed degrees); PHW_DEGREE_SORT=0 numbers the vertices by first touch too (A/B:
    // neutral for the cell-based kernels, profiles/r1b_ab_first_touch_numbering.txt).
    static const bool degree_sort = [] { const char* e = getenv("PHW_DEGREE_SORT"); return !e || atoi(e) != 0; }();
    auto renumber = [&](const std::vector<int32_t>& owner, const std::vector<int32_t>& cell_dofs, const std::vector<int32_t>* degree,
                        int64_t n, std::vector<int32_t>& new2old, std::vector<int32_t>& old2new, std::vector<int32_t>& pstart) {
        std::vector<int32_t> ord(n);
        if (degree && degree_sort) {
            int32_t dmax = 0;
            for (int64_t i = 0; i < n; ++i) dmax = std::max(dmax, (*degree)[i]);
            std::vector<int64_t> cnt(dmax + 2, 0);
            for (int64_t i = 0; i < n; ++i) cnt[dmax - (*degree)[i] + 1]++;
            for (int32_t d = 0; d <= dmax; ++d) cnt[d + 1] += cnt[d];
            for (int64_t i = 0; i < n; ++i) ord[cnt[dmax - (*degree)[i]]++] = (int32_t)i;
        } else {
            std::vector<uint8_t> seen(n, 0);
            int64_t w = 0;
            for (int64_t k = 0; k < nc; ++k) {
                const int64_t c = order[k];
                const int32_t pc = cell_patch[c];
                for (int i = 0; i < 3; ++i) {
                    const int32_t d = cell_dofs[3 * c + i];
                    if (!seen[d] && owner[d] == pc) { seen[d] = 1; ord[w++] = d; }
                }
            }
            for (int64_t i = 0; i < n; ++i) if (!seen[i]) ord[w++] = (int32_t)i;   // dofs of other ranks (virtual patch npatch)
        }
        pstart.assign(npatch + 2, 0);
        for (int64_t i = 0; i < n; ++i) pstart[owner[i] + 1]++;
        for (int32_t p = 0; p <= npatch; ++p) pstart[p + 1] += pstart[p];
        std::vector<int32_t> pos(pstart.begin(), pstart.end() - 1);
        new2old.resize(n); old2new.resize(n);
        for (int64_t k = 0; k < n; ++k) {
            int32_t i = ord[k];
            int32_t d = pos[owner[i]]++;
            new2old[d] = i;
            old2new[i] = d;
        }
    };
    std::vector<int32_t> vstart, estart;
    renumber(v_owner, cells, &v_degree, nv, v_new2old, v_old2new, vstart);
    renumber(e_owner, cell_edges, nullptr, ne, e_new2old, e_old2new, estart);
    nv_owned = vstart[npatch]; ne_owned = estart[npatch];
    // ---- 5. halo cells: (patch, cell, type) type 0 = adjacent to an owned edge, 1 = only to an owned vertex
    std::vector<uint64_t> halo;
    for (int64_t c = 0; c < nc; ++c) {
        int32_t pc = cell_patch[c];
        for (int i = 0; i < 3; ++i) {
            int32_t pe = e_owner[cell_edges[3 * c + i]];
            if (pe != pc && pe != npatch) halo.push_back(((uint64_t)pe << 33) | ((uint64_t)c << 1) | 0u);
            int32_t pv = v_owner[cells[3 * c + i]];
            if (pv != pc && pv != npatch) halo.push_back(((uint64_t)pv << 33) | ((uint64_t)c << 1) | 1u);
        }
    }
    std::sort(halo.begin(), halo.end());
    {   // unique on (patch, cell), keeping the smallest type
        size_t w = 0;
        for (size_t r = 0; r < halo.size(); ++r)
            if (w == 0 || (halo[r] >> 1) != (halo[w - 1] >> 1)) halo[w++] = halo[r];
        halo.resize(w);
    }
    // ---- 6. per-patch local numbering, cell lists, adjacency
    hdr.assign(npatch, PatchHdr{});
    pcell_user.clear(); pc_lv.clear(); pc_le.clear(); ghost_v.clear(); ghost_e.clear();
    pcell_user.reserve(nc + halo.size());
    eadj.assign(ne, 0xFFFFFFFFu);
    vslice_off.clear();
    vslice_off.push_back(0);
    ell_ioff.assign(1, 0); ell_woff.assign(1, 0); ell_ids.clear(); ell_pair.clear();
    vadj.clear();
    vadj.reserve(4 * nc + 4 * halo.size());
    std::vector<int32_t> lv_of(nv, -1), le_of(ne, -1);
    std::vector<int32_t> gv_tmp, ge_tmp, cl;
    std::vector<uint32_t> vcnt;
    max_lv = max_le = max_cells = 0;
    max_smem_ell = 0;
    size_t hpos = 0;
    for (int32_t p = 0; p < npatch; ++p) {
        PatchHdr& H = hdr[p];
        H.cell_off = (int32_t)pcell_user.size();
        H.group = patch_group[p];
        H.n_own = (int32_t)(own_off[p + 1] - own_off[p]);
        H.v_start = vstart[p]; H.v_cnt = vstart[p + 1] - vstart[p];
        H.e_start = estart[p]; H.e_cnt = estart[p + 1] - estart[p];
        cl.clear();
        for (int64_t k = own_off[p]; k < own_off[p + 1]; ++k) cl.push_back(order[k]);
        size_t h0 = hpos;
        while (hpos < halo.size() && (int32_t)(halo[hpos] >> 33) == p) ++hpos;
        for (size_t r = h0; r < hpos; ++r)
            if ((halo[r] & 1u) == 0) cl.push_back((int32_t)((halo[r] >> 1) & 0xFFFFFFFFu));
        H.n_ecells = (int32_t)cl.size();
        for (size_t r = h0; r < hpos; ++r)
            if ((halo[r] & 1u) == 1) cl.push_back((int32_t)((halo[r] >> 1) & 0xFFFFFFFFu));
        H.n_vcells = (int32_t)cl.size();
        // ghosts
        gv_tmp.clear(); ge_tmp.clear();
        for (int32_t c : cl)
            for (int i = 0; i < 3; ++i) {
                int32_t v = cells[3 * (int64_t)c + i];
                if (v_owner[v] != p && lv_of[v] < 0) { lv_of[v] = 0; gv_tmp.push_back(v_old2new[v]); }
                int32_t e = cell_edges[3 * (int64_t)c + i];
                if (e_owner[e] != p && le_of[e] < 0) { le_of[e] = 0; ge_tmp.push_back(e_old2new[e]); }
            }
        std::sort(gv_tmp.begin(), gv_tmp.end());
        std::sort(ge_tmp.begin(), ge_tmp.end());
        H.gv_off = (int32_t)ghost_v.size(); H.gv_cnt = (int32_t)gv_tmp.size();
        H.ge_off = (int32_t)ghost_e.size(); H.ge_cnt = (int32_t)ge_tmp.size();
        for (size_t k = 0; k < gv_tmp.size(); ++k) { lv_of[v_new2old[gv_tmp[k]]] = H.v_cnt + (int32_t)k; ghost_v.push_back(gv_tmp[k]); }
        for (size_t k = 0; k < ge_tmp.size(); ++k) { le_of[e_new2old[ge_tmp[k]]] = H.e_cnt + (int32_t)k; ghost_e.push_back(ge_tmp[k]); }
        int32_t nlv = H.v_cnt + H.gv_cnt, nle = H.e_cnt + H.ge_cnt;
        if (nlv >= 65536 || nle >= 16384 || H.n_vcells >= 16383)
            return "patch too large for 16-bit local indices; reduce patch_cells";
        if (p < npatch_rows) { max_lv = std::max(max_lv, nlv); max_le = std::max(max_le, nle); max_cells = std::max(max_cells, H.n_vcells); }
        // records + adjacency
        vcnt.assign(H.v_cnt + 1, 0);
        for (size_t lc = 0; lc < cl.size(); ++lc) {
            int64_t c = cl[lc];
            pcell_user.push_back((int32_t)c);
            for (int i = 0; i < 3; ++i) {
                int32_t v = cells[3 * c + i], e = cell_edges[3 * c + i];
                int32_t lv = (v_owner[v] == p) ? v_old2new[v] - H.v_start : lv_of[v];
                int32_t le = (e_owner[e] == p) ? e_old2new[e] - H.e_start : le_of[e];
                pc_lv.push_back((uint16_t)lv);
                pc_le.push_back((uint16_t)le);
                if (v_owner[v] == p) vcnt[lv + 1]++;
                if (e_owner[e] == p) {
                    if ((int32_t)lc >= H.n_ecells) return "internal: owned edge adjacent to a vertex-halo cell";
                    uint32_t& a = eadj[e_old2new[e]];
                    uint32_t code = (uint32_t)(lc * 4 + i);
                    if ((a & 0xFFFFu) == 0xFFFFu) a = (a & 0xFFFF0000u) | code;
                    else if ((a >> 16) == 0xFFFFu) a = (a & 0xFFFFu) | (code << 16);
                    else return "internal: edge with more than two adjacent patch cells";
                }
            }
        }
        // vertex adjacency as sliced ELL (entries of a vertex ascending in lc -> fixed summation order)
        H.vs_off = (int32_t)(vslice_off.size() - 1);
        const int32_t nslice = (H.v_cnt + 31) / 32;
        size_t base = vadj.size();
        std::vector<size_t> sbase(nslice + 1, 0);
        for (int32_t sl = 0; sl < nslice; ++sl) {
            uint32_t w = 0;
            for (int32_t k = 32 * sl; k < std::min(32 * (sl + 1), H.v_cnt); ++k) w = std::max(w, vcnt[k + 1]);
            w = (w + 7u) & ~7u;                    // chunks of 8 codes = one 16-byte load per vertex
            sbase[sl + 1] = sbase[sl] + (size_t)w * 32;
            vslice_off.push_back((uint32_t)(base + sbase[sl + 1]));
        }
        vadj.resize(base + sbase[nslice], (uint16_t)0xFFFFu);
        H.adj_off = (int32_t)(uint32_t)base; H.adj_cnt = (int32_t)sbase[nslice];
        std::vector<uint32_t> fill(H.v_cnt, 0);
        for (size_t lc = 0; lc < cl.size(); ++lc) {
            int64_t c = cl[lc];
            for (int i = 0; i < 3; ++i) {
                int32_t v = cells[3 * c + i];
                if (v_owner[v] == p) {
                    int32_t lv = v_old2new[v] - H.v_start;
                    const uint32_t j = fill[lv]++;
                    vadj[base + sbase[lv / 32] + ((size_t)(j / 8) * 32 + (lv % 32)) * 8 + (j % 8)] = (uint16_t)(lc * 4 + i);
                }
            }
        }
        // ---- ELL rows of the assembled P1 mass matrix for the owned vertices: local ids of the neighbour vertices and, per
        //      entry, the (<= 2) patch cells shared with that neighbour (the weight (|K1| + |K2|) / 12 is filled on the device)
        {
            std::vector<uint32_t> rstart(H.v_cnt + 1, 0), rfill(H.v_cnt, 0);
            for (int32_t lv = 0; lv < H.v_cnt; ++lv) rstart[lv + 1] = rstart[lv] + vcnt[lv + 1];
            std::vector<uint16_t> rcode(rstart[H.v_cnt]);
            for (size_t lc = 0; lc < cl.size(); ++lc)
                for (int i = 0; i < 3; ++i) {
                    const int32_t v = cells[3 * (int64_t)cl[lc] + i];
                    if (v_owner[v] == p) { const int32_t lv = v_old2new[v] - H.v_start; rcode[rstart[lv] + rfill[lv]++] = (uint16_t)(lc * 4 + i); }
                }
            const size_t pcb = (size_t)H.cell_off;
            std::vector<uint16_t> nid((size_t)H.v_cnt * 16);
            std::vector<uint32_t> npair((size_t)H.v_cnt * 16), ncnt(H.v_cnt, 0);
            for (int32_t lv = 0; lv < H.v_cnt; ++lv) {
                uint32_t cnt = 0;
                for (uint32_t k = rstart[lv]; k < rstart[lv + 1]; ++k) {
                    const uint32_t lc = rcode[k] >> 2, slot = rcode[k] & 3u;
                    for (int t = 1; t <= 2; ++t) {
                        const uint16_t o = pc_lv[3 * (pcb + lc) + (slot + t) % 3];
                        uint32_t j = 0;
                        while (j < cnt && nid[(size_t)lv * 16 + j] != o) ++j;
                        if (j < cnt) npair[(size_t)lv * 16 + j] = (npair[(size_t)lv * 16 + j] & 0xFFFFu) | (lc << 16);
                        else {
                            if (cnt >= 16) return "vertex with more than 16 neighbours: not supported by the ELL mass rows";
                            nid[(size_t)lv * 16 + cnt] = o; npair[(size_t)lv * 16 + cnt] = lc | 0xFFFF0000u; ++cnt;
                        }
                    }
                }
                ncnt[lv] = cnt;
            }
            const size_t ibase = ell_ids.size(), wbase = ell_pair.size();
            size_t ioff = 0, woff = 0;
            for (int32_t sl = 0; sl < nslice; ++sl) {
                uint32_t w = 0;
                for (int32_t k = 32 * sl; k < std::min(32 * (sl + 1), H.v_cnt); ++k) w = std::max(w, ncnt[k]);
                const uint32_t nch = (w + 7u) / 8u;
                ell_ids.resize(ibase + ioff + (size_t)nch * 256, 0);
                ell_pair.resize(wbase + woff + (size_t)w * 32, 0xFFFFFFFFu);
                for (int32_t k = 32 * sl; k < 32 * (sl + 1); ++k) {
                    const int lane = k % 32;
                    for (uint32_t j = 0; j < nch * 8; ++j) {
                        const bool valid = k < H.v_cnt && j < ncnt[k];
                        // padding points at the row itself (or at 0) with weight zero: no predicate in the kernel
                        ell_ids[ibase + ioff + ((size_t)(j / 8) * 32 + lane) * 8 + (j % 8)] = valid ? nid[(size_t)k * 16 + j] : (uint16_t)(k < H.v_cnt ? k : 0);
                        if (j < w && valid) ell_pair[wbase + woff + (size_t)j * 32 + lane] = npair[(size_t)k * 16 + j];
                    }
                }
                ioff += (size_t)nch * 256; woff += (size_t)w * 32;
                ell_ioff.push_back((uint32_t)(ibase + ioff));
                ell_woff.push_back((uint32_t)(wbase + woff));
            }
            H.ell_i0 = (int32_t)(uint32_t)ibase; H.ell_in = (int32_t)ioff;
            H.ell_w0 = (int32_t)(uint32_t)wbase; H.ell_wn = (int32_t)woff;
            // shared memory of the staged pass (doubles): slice offsets | x [nlv] | b, x_prev [v_cnt] | weights | ids
            auto ev2 = [](int64_t x) { return (x + 1) & ~(int64_t)1; };
            if (p < npatch_rows) max_smem_ell = std::max(max_smem_ell, (int64_t)((nslice + 3) & ~1) + ev2(nlv) + 2 * ev2(H.v_cnt) + (int64_t)woff + (int64_t)(ioff + 3) / 4 + 2);
        }
        // reset scratch
        for (int32_t g : gv_tmp) lv_of[v_new2old[g]] = -1;
        for (int32_t g : ge_tmp) le_of[e_new2old[g]] = -1;
    }
    if (vadj.size() >= ((size_t)1 << 32) || ell_ids.size() >= ((size_t)1 << 32) || ell_pair.size() >= ((size_t)1 << 32))
        return "vertex adjacency exceeds 32-bit offsets";
    return "";
}

std::string Layout::self_check() const {
    auto err = [](const std::string& what, int64_t p, int64_t i) { return what + " (patch " + std::to_string(p) + ", item " + std::to_string(i) + ")"; };
    // ---- every dof is a row of exactly one patch; owned ranges tile the numbering
    int64_t vsum = 0, esum = 0;
    for (int32_t p = 0; p < npatch; ++p) {
        const PatchHdr& H = hdr[p];
        if (H.v_start != vsum || H.e_start != esum) return err("owned ranges are not contiguous", p, 0);
        // blocks the pipelined kernels fetch with one bulk copy each must be 16-byte aligned in address and size
        if ((H.ell_i0 % 8) != 0 || (H.ell_in % 8) != 0 || (H.ell_w0 % 2) != 0 || (H.ell_wn % 2) != 0 || (H.adj_off % 8) != 0 || (H.adj_cnt % 8) != 0)
            return err("ELL / adjacency block of a patch is not 16-byte aligned", p, 0);
        vsum += H.v_cnt; esum += H.e_cnt;
    }
    if (vsum != nv_owned || esum != ne_owned) return "owned ranges do not cover the owned dofs";
    // ---- ELL mass rows: neighbours of a vertex = the other vertices of its cells; the shared cells contain both
    for (int32_t p = 0; p < npatch; ++p) {
        const PatchHdr& H = hdr[p];
        const int32_t nsl = (H.v_cnt + 31) / 32;
        auto gid_of = [&](uint32_t lv) { return (int32_t)lv < H.v_cnt ? H.v_start + (int32_t)lv : ghost_v[H.gv_off + (int32_t)lv - H.v_cnt]; };
        for (int32_t lv = 0; lv < H.v_cnt; ++lv) {
            const int32_t sl = lv / 32, lane = lv % 32;
            const uint32_t io0 = ell_ioff[H.vs_off + sl], wo0 = ell_woff[H.vs_off + sl];
            const uint32_t W = (ell_woff[H.vs_off + sl + 1] - wo0) / 32;
            const int32_t v_old = v_new2old[H.v_start + lv];
            std::vector<int32_t> expect;       // user ids of the neighbours, from the global cell list
            for (int32_t k = 0; k < H.n_vcells; ++k) {
                const int32_t c = pcell_user[H.cell_off + k];
                for (int i = 0; i < 3; ++i)
                    if (cells[3 * (int64_t)c + i] == v_old) { expect.push_back(cells[3 * (int64_t)c + (i + 1) % 3]); expect.push_back(cells[3 * (int64_t)c + (i + 2) % 3]); }
            }
            std::sort(expect.begin(), expect.end());
            expect.erase(std::unique(expect.begin(), expect.end()), expect.end());
            std::vector<int32_t> got;
            for (uint32_t k = 0; k < W; ++k) {
                const uint32_t pr = ell_pair[wo0 + k * 32 + lane];
                const uint16_t id = ell_ids[io0 + ((size_t)(k / 8) * 32 + lane) * 8 + (k % 8)];
                if (pr == 0xFFFFFFFFu) { if (id != lv) return err("ELL padding does not point at its own row", p, lv); continue; }
                const int32_t w_old = v_new2old[gid_of(id)];
                got.push_back(w_old);
                for (int h = 0; h < 2; ++h) {
                    const uint32_t lc = h == 0 ? (pr & 0xFFFFu) : (pr >> 16);
                    if (lc == 0xFFFFu) { if (h == 0) return err("ELL entry without a cell", p, lv); continue; }
                    if ((int32_t)lc >= H.n_vcells) return err("ELL cell index out of range", p, lv);
                    const int32_t c = pcell_user[H.cell_off + lc];
                    int hit = 0;
                    for (int i = 0; i < 3; ++i) hit += (cells[3 * (int64_t)c + i] == v_old) + 2 * (cells[3 * (int64_t)c + i] == w_old);
                    if (hit != 3) return err("ELL cell does not contain both vertices", p, lv);
                }
            }
            std::sort(got.begin(), got.end());
            if (got != expect) return err("ELL neighbours differ from the cell connectivity", p, lv);
        }
        (void)nsl;
    }
    return "";
}

void Layout::make_records(bool strong, bool casimir, std::vector<CellRec>& rec) const {
    // per-edge boundary tag (-1 interior)
    std::vector<int8_t> etag(ne, -1);
    for (int64_t k = 0; k < nb; ++k) etag[bedge[k]] = (int8_t)btag[k];
    const size_t npc = pcell_user.size();
    rec.resize(npc);
    for (size_t k = 0; k < npc; ++k) {
        int64_t c = pcell_user[k];
        uint32_t le[3], lv[3], flags = 0;
        for (int i = 0; i < 3; ++i) {
            int32_t e = cell_edges[3 * c + i];
            lv[i] = pc_lv[3 * k + i];
            le[i] = pc_le[3 * k + i];
            if (cell_sign[3 * c + i] < 0) le[i] |= 0x8000u;
            int8_t t = etag[e];
            if (t >= 0) {
                bool nterm = strong ? true : (t == 1 || t == 3);
                if (nterm) le[i] |= 0x4000u;
                if (casimir && t == 1) flags |= (1u << i);
                if (casimir && t == 2) flags |= (1u << (3 + i));
            }
        }
        if (!cell_foreign.empty() && cell_foreign[c]) flags |= (1u << 6);   // cell of another rank: no weight in cell-based sums
        rec[k].x = lv[0] | (lv[1] << 16);
        rec[k].y = lv[2] | (flags << 16);
        rec[k].z = le[0] | (le[1] << 16);
        rec[k].w = le[2];
    }
}

}  // namespace phw
