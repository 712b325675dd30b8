// layout.hpp - host-side mesh connectivity and patch layout (no CUDA dependency).
//
// Replaces what dolfin builds behind `FunctionSpace(mesh, MixedElement([P1, RT, ...]))`,
// `MeshFunction(...).mark` and `FacetNormal` (reference wave_2D.py:168-220,266-285): dof maps of the
// P1 (vertex) and RT1 (edge) spaces, edge orientation signs and boundary tags - re-organised for the
// GPU as *patches*: runs of cells contiguous along a Hilbert curve, each owning a contiguous range of
// vertex- and edge-dofs, with one layer of halo cells so that every owned row can be evaluated inside
// one thread block from shared memory (owner-computes, deterministic, no global atomics).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace phw {

// 16-byte per-(patch,cell) record.  Local ids index the patch's local vertex/edge arrays:
// [owned (contiguous global range) | ghosts].
//   x = lv0 | lv1<<16          y = lv2 | flags<<16
//   z = le0 | le1<<16          w = le2
// le bits: 0..13 local edge id, bit 14: "N-term" edge (zero-flux boundary side, wave_2D.py:465-472
// ds(1), ds(3); every boundary edge in the strong form), bit 15: orientation sign is -1.
// flags bits (y>>16): 0..2 slot i lies on the impedance top side (casimir, tag 1),
//                     3..5 slot i lies on the impedance right side (casimir, tag 2),
//                     6 the cell belongs to another rank of a partitioned run (halo cell of the local mesh): it contributes to
//                       the rows of this rank's dofs but carries no weight in cell-based sums (energy_pipe_k).
struct CellRec { uint32_t x, y, z, w; };   // (x,y) = vertex half, (z,w) = edge half; stored as two uint2 arrays on the device

struct PatchHdr {
    int32_t cell_off;   // first entry in the patch-cell arrays
    int32_t n_own;      // own cells
    int32_t n_ecells;   // cells processed by edge-row operators (own + E-halo)
    int32_t n_vcells;   // cells processed by vertex-row operators (all)
    int32_t v_start, v_cnt, gv_off, gv_cnt;   // owned vertex range (new numbering), ghost list
    int32_t e_start, e_cnt, ge_off, ge_cnt;   // owned edge range, ghost list
    int32_t vs_off;     // first 32-vertex slice of this patch in vslice_off
    int32_t group;      // case index of a batched (replicated-mesh) run, 0 otherwise
    int32_t adj_off;    // first uint16 code of this patch in vadj (uint32 bits) and number of codes: the patch's adjacency
    int32_t adj_cnt;    // is one contiguous range
    int32_t ell_i0, ell_in;   // ELL mass rows of the patch: first id code / number of codes in ell_ids (uint32 bits),
    int32_t ell_w0, ell_wn;   // first weight / number of weights in ell_w: two contiguous, 16-byte aligned blocks
};
static_assert(sizeof(PatchHdr) == 80, "PatchHdr is loaded as 20 ints");

struct Layout {
    int64_t nv = 0, nc = 0, ne = 0, nb = 0;
    int32_t patch_cells = 0, npatch = 0;
    int32_t npatch_rows = 0;           // patches that own rows = the patches the kernels process (partitioned runs: patches of halo cells only come last)
    std::vector<double>  vtx;          // [nv,2] user order
    std::vector<int32_t> cells;        // [nc,3] user order, CCW
    std::vector<int32_t> edges;        // [ne,2] (lo,hi), canonical order
    std::vector<int32_t> cell_edges;   // [nc,3]
    std::vector<int8_t>  cell_sign;    // [nc,3]
    std::vector<int32_t> bedge;        // [nb] boundary edge ids (ascending)
    std::vector<int8_t>  bsign;        // [nb]
    std::vector<int32_t> btag;         // [nb] phw_tag (default NONE)
    std::vector<double>  bweight;      // [nb] left-port weights (default 1)
    bool boundary_set = false;
    std::vector<int32_t> dir_vertex;   // strong Dirichlet rows (user vertex ids) and their value weights
    std::vector<double>  dir_weight;   // value = weight * g(t)  (0 for the homogeneous right side)
    std::vector<uint8_t> cell_foreign; // partitioned runs (optional, by user cell id): 1 = cell owned by another rank

    std::vector<int32_t> cell_patch;   // [nc] user cell -> patch
    std::vector<int32_t> v_new2old, v_old2new, e_new2old, e_old2new;
    std::vector<int32_t> v_owner, e_owner;     // by old id

    std::vector<PatchHdr> hdr;
    std::vector<int32_t> pcell_user;   // [npc] user cell id of every patch cell
    std::vector<uint16_t> pc_lv;       // [npc,3] local vertex ids
    std::vector<uint16_t> pc_le;       // [npc,3] local edge ids (no flag bits)
    std::vector<int32_t> ghost_v, ghost_e;     // new ids
    std::vector<uint32_t> eadj;        // [ne] by new id: (lc*4+slot) | (lc*4+slot)<<16, 0xFFFF = none
    // vertex -> (cell,slot) adjacency as sliced ELL: owned vertices are sorted by degree inside a patch and
    // grouped in slices of 32; slice s stores ceil(width_s/8) chunks, chunk k holding 8 codes (lc*4+slot,
    // 0xFFFF = padding) for each of the 32 vertices: [chunk][lane][8] -> one coalesced 16-byte load per vertex
    // and chunk, warp-uniform trip count.
    std::vector<uint32_t> vslice_off;  // [n_slices+1] offsets into vadj
    std::vector<uint16_t> vadj;
    // assembled P1 mass rows of the owned vertices as sliced ELL (same 32-row slices): ell_ids = local ids of the neighbour
    // vertices, chunks of 8 as in vadj (padding = the row itself); ell_pair = per entry [k][lane] the two patch cells shared
    // with that neighbour (lcA | lcB << 16, 0xFFFF none; 0xFFFFFFFF = padding) from which the device fills the weights
    std::vector<uint32_t> ell_ioff, ell_woff;   // [n_slices+1]
    std::vector<uint16_t> ell_ids;
    std::vector<uint32_t> ell_pair;
    std::vector<int32_t> v_degree;     // [nv] by old id
    int32_t max_lv = 0, max_le = 0, max_cells = 0;
    int64_t max_smem_ell = 0;          // doubles of shared memory vertex_ell_pass needs with all streams of a patch staged

    int64_t nv_owned = 0, ne_owned = 0;   // external (other-rank) dofs are numbered after the owned ones
    // v_ext / e_ext (optional, by user id): 1 = dof owned by another rank (never a row of this rank)
    std::string build(int64_t nv_, const double* vxy, int64_t nc_, const int32_t* cells_, int32_t patch_cells_,
                      const uint8_t* v_ext = nullptr, int64_t ne_expected = -1, const uint8_t* e_ext = nullptr,
                      const int32_t* cell_group = nullptr, int32_t n_groups_ = 1);
    int32_t n_groups = 1;
    std::vector<int32_t> group_patch_off;   // [n_groups+1] patches of a group are contiguous
    std::vector<int32_t> cell_group_v;      // [nc] (empty when n_groups == 1)
    // cell records with boundary flag bits for the given formulation
    void make_records(bool strong, bool casimir, std::vector<CellRec>& rec) const;
    // structural invariants of the patch layout and the ELL mass rows (host-only; tests)
    std::string self_check() const;
};

}  // namespace phw
