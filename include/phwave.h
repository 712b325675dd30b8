/* phwave.h - C ABI of libphwave.so: the B200-native hot path of FinbarArgus/portHamiltonian_FEM.
 *
 * The reference has no FFI of its own: its only seam is the Python function
 *     wave_2D_solve(...)          fenics_projects/wave_eqn/wave_2D.py:15-20 (returns at :1032)
 * whose body asks dolfin/PETSc to (a) mesh + build dof maps (:92-285), (b) assemble the operators of
 * the forms (:449-595), (c) per step assemble a right-hand side and LU-solve (:787-854), (d) assemble
 * energy functionals (:876-927).  This header is the boundary a maintainer binds instead (ctypes stub
 * in INTEGRATION.md); each entry point names the reference region it replaces.
 *
 * Conventions: plain C, no exceptions cross the ABI, every function returns 0 on success and a
 * negative phw_status otherwise (text via phw_last_error(), thread-local).  All buffers are owned by
 * the caller; the library copies what it keeps.  Vector arguments are in the *user numbering*
 * (vertex ids as given; edge ids = unique (lo,hi) vertex pairs in lexicographic order) and may be host
 * or device pointers (resolved with cudaPointerGetAttributes).  A handle is bound to one device and
 * one stream and is not thread-safe.
 */
#ifndef PHWAVE_H
#define PHWAVE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct phw_mesh_s*   phw_mesh_t;
typedef struct phw_solver_s* phw_solver_t;

enum phw_status { PHW_OK = 0, PHW_EINVAL = -1, PHW_ECUDA = -2, PHW_ENOMEM = -3, PHW_ESTATE = -4, PHW_ENOCONV = -5 };

/* time integration schemes, wave_2D.py:295-324 */
enum phw_scheme { PHW_EE = 0, PHW_IE = 1, PHW_SE = 2, PHW_SM = 3, PHW_EH = 4, PHW_SV = 5 };
/* boundary tags, wave_2D.py:266-280 */
enum phw_tag { PHW_TAG_LEFT = 0, PHW_TAG_TOP = 1, PHW_TAG_RIGHT = 2, PHW_TAG_GENERAL = 3, PHW_TAG_NONE = 4 };
/* input signal of the left port / the voltage source */
enum phw_input {
    PHW_IN_SINE_WINDOW = 0, /* (t<t_stop) ? amp*sin(omega t) : 0   wave_2D.py:384,433,444 */
    PHW_IN_SINE        = 1, /* amp*sin(omega t)                    wave_2D.py:380 (casimir)   */
    PHW_IN_COSINE      = 2, /* amp*cos(omega t)                    wave_2D.py:376 (analytical time factor) */
    PHW_IN_PASSIVITY   = 3  /* (t<t_stop) ? amp*sin(omega t) : (t<t_ctrl) ? 0 : -gain*x0_prev   control_input.py:12-18 */
};
enum phw_which { PHW_P = 0, PHW_Q = 1 };

typedef struct phw_params {
    int32_t scheme;          /* phw_scheme */
    int32_t strong;          /* 0 weak / 1 strong Dirichlet (wave_2D.py:413-422) */
    int32_t interconnection; /* 0 none / 1 'IC' three-state lumped model (wave_2D.py:148-166,528-577) */
    int32_t casimir;         /* 1: impedance top/right boundaries (wave_2D.py:327-351) */
    int32_t analytical;      /* 1: analytical diagnostics -> 12 output columns (wave_2D.py:929-965) */
    int32_t input_kind;      /* phw_input */
    double dt, K_wave, rho;
    double in_amp, in_omega, in_t_stop, in_t_ctrl, in_gain;
    double Bl_em, R1_em, R2_em, K_em, m_em, L_em;   /* wave_2D.py:153-158 */
    double H_init;           /* wave_2D.py:746,769,771 (computed by the caller with phw_energy) */
    double an_omega, an_xLength, an_yLength, an_xPoint, an_yPoint;  /* analytical case constants */
    double tol;              /* relative residual of the mass / block solves (default 1e-13) */
    int32_t max_iter;        /* iteration cap per solve (0: 200; wave_2D_solve passes 2000 - stretched cells need more than 200) */
    int32_t extrap_order;    /* initial guess of each solve = polynomial extrapolation of that solve's previous
                                solutions: 0 last solution, 1 linear, 2 quadratic (default choice of the host code), 3 cubic ... 5 */
    int32_t flags;           /* bit0: disable first-same-as-last reuse in SV; bit1: synchronous error checks;
                                bit2: CUDA-event timing of every solve kernel; bit3: force the multi-launch solve path;
                                bit4: (unused; was the two-sweep mass solves, removed in round 2: measured slower);
                                bit5: cell-based instead of assembled (ELL) rows in the M_p solve;
                                bit6: ELL rows of the M_p solve staged in shared memory by cp.async (whole patch in flight;
                                A/B: slower than the direct loads);
                                bit7: generic instead of lean kernels for the M_q solve and the energy reductions;
                                bit8: bulk-copy pipelined M_q solve (csrc/kernels_pipe.cuh: every stream of the CTA's next patch
                                is fetched by cp.async.bulk into a second shared-memory stage while the current patch is computed;
                                one GPU; needs patches of <= ~1000 cells so that two stages fit - phw_solver_create fails otherwise);
                                bit9: the same for the ELL M_p solve;
                                bit10: the generic cooperative kernels (block solve of IE/SM/casimir, strong form, partitioned and
                                higher-order CSR solves) synchronise the grid with the one-atomic-per-CTA barrier
                                of the lean kernels instead of cooperative-groups grid syncs;
                                bit11: H_wave (both quadratic forms) in one bulk-copy pipelined pass over the own cells of every
                                patch (64 instead of 88 B per cell; one GPU, unbatched runs);
                                bit12: bulk-copy pipelined right-hand sides D p / D^T q (unbatched, no casimir terms);
                                bit15: the WHOLE time loop of phw_run in one launch (csrc/kernels_loop.cuh) for cache-resident meshes: one
                                CTA per patch for the whole run, the CTAs of a case in one thread-block cluster (1, 2, 4, 8 or 16 patches
                                per case) or one co-resident grid; every scheme and variant except the analytical diagnostics; one GPU.
                                Silently ignored when the layout is not eligible (phw_stats.kernel_paths tells: PHW_PATH_LOOP);
                                bit16: conjugate-gradient instead of Chebyshev mass solves (solve_pcg_coop_k).  Selected automatically for
                                the M_q solve when the element-wise bound of kappa(D^-1 M_q) exceeds 10 (stretched cells, slivers),
                                where the Chebyshev interval is pessimistic; bit17 forbids it.  One GPU, weak form, unbatched,
                                multi-launch path; phw_stats.kernel_paths tells (PHW_PATH_PCG) */
    int32_t extrap_order_q;  /* extrapolation order of the q (and block) solves; 0: same as extrap_order.  Orders up to 5 */
} phw_params;

typedef struct phw_stats {
    int64_t steps, solves_p, solves_q, solves_block, iters_p, iters_q, iters_block;
    double  max_rel_residual;   /* largest accepted ||b - A x|| / ||b|| */
    double  device_ms;          /* CUDA-event time of the last phw_run */
    int64_t kernel_launches;    /* launches issued by the last phw_run */
    int64_t bytes_model;        /* algorithmic bytes of the last phw_run ("bytes model") */
    double  lam_min_q, lam_max_q, skew_bound;
    double  ms_solve_p, ms_solve_q, ms_solve_block;   /* summed CUDA-event time of the solve kernels (flags bit2) */
    int64_t kernel_paths;       /* phw_path bits: which variants of the mass-solve kernels the last phw_run launched */
} phw_stats;
enum phw_path { PHW_PATH_LEAN_P = 1, PHW_PATH_LEAN_Q = 2, PHW_PATH_PIPE_P = 4, PHW_PATH_PIPE_Q = 8, PHW_PATH_PIPE_ENERGY = 16, PHW_PATH_PIPE_RHS = 32, PHW_PATH_LOOP = 256, PHW_PATH_LOOP_RES = 512 /* ... with the mass solves resident in shared memory */,
                PHW_PATH_HO_MF = 1024 /* matrix-free higher-order kernels (phw_ho_create_mf) */,
                PHW_PATH_PCG = 2048 /* conjugate-gradient mass solves (meshes of poor quality: element-wise kappa bound > 10) */,
                PHW_PATH_SWEEP_CSR = 4096 /* batched sweep with the case index fastest (phw_sweep_create) */ };

const char* phw_last_error(void);
int phw_version(void);
/* sizeof(phw_params) (what = 0) / sizeof(phw_stats) (what = 1) as compiled into the library: lets a binding verify its struct
 * layouts at load time */
int phw_sizeof(int what);

/* ---- mesh connectivity + patch layout: replaces mesh.topology / FunctionSpace dofmaps / MeshFunction
 *      markers (wave_2D.py:126,168-220,266-285).  Host-only, no GPU needed. ---- */
int phw_mesh_create(int64_t n_vertices, const double* vertices_xy, int64_t n_cells, const int32_t* cells,
                    int32_t patch_cells, phw_mesh_t* out);
/* partitioned run: the local mesh of one rank (own cells + one halo layer); vertex_external[n_vertices] /
 * edge_external[n_edges] flag the dofs owned by another rank (edge ids = the canonical numbering above).
 * Replaces dolfin's MPI mesh distribution (run_main_wave_2D_mpi.sh:1, wave_2D.py:54-61). */
int phw_mesh_create_partitioned(int64_t n_vertices, const double* vertices_xy, int64_t n_cells, const int32_t* cells,
                                int32_t patch_cells, const uint8_t* vertex_external, int64_t n_edges,
                                const uint8_t* edge_external, phw_mesh_t* out);
/* partitioned run, optional: cell_owned[n_cells] = 1 for the cells the partitioner gave to this rank, 0 for the halo cells of the
 * local mesh (owned by a neighbour).  Needed by cell-based reductions (the fused H_wave pass, flags bit 11), which must count every
 * cell of the global mesh exactly once; without it a partitioned run takes the row-based reductions. */
int phw_mesh_set_cell_owned(phw_mesh_t m, const uint8_t* cell_owned);
/* batched parameter sweep (BASELINE.json configs[4]; the reference runs its cases one after the other,
 * main_wave_2D.py:144): the mesh holds n_groups disjoint copies of one domain, cell_group[n_cells] names the copy
 * (= case) of every cell.  All cases advance in the same kernels; per-case physics via
 * phw_solver_set_group_params.  H_rows of phw_run becomes [n_groups][n_steps][8]. */
int phw_mesh_create_grouped(int64_t n_vertices, const double* vertices_xy, int64_t n_cells, const int32_t* cells,
                            int32_t patch_cells, const int32_t* cell_group, int32_t n_groups, phw_mesh_t* out);
int phw_mesh_sizes(phw_mesh_t m, int64_t* n_vertices, int64_t* n_cells, int64_t* n_edges,
                   int64_t* n_boundary_edges, int64_t* n_patches);
/* edges [n_edges,2], cell_edges [n_cells,3], cell_edge_sign [n_cells,3] (+1/-1); any pointer may be NULL */
int phw_mesh_get_edges(phw_mesh_t m, int32_t* edges, int32_t* cell_edges, int8_t* cell_edge_sign);
/* boundary edges: ids [n_b], sign of the adjacent cell [n_b] (+1: global normal is outward) */
int phw_mesh_get_boundary(phw_mesh_t m, int32_t* edge_ids, int8_t* sign);
/* tags [n_b] (phw_tag) and optional left-port weights [n_b] = (1/|e|) int_e g_L ds  (NULL -> 1) */
int phw_mesh_set_boundary(phw_mesh_t m, const int32_t* tags, const double* port_weight);
/* strong Dirichlet rows (DirichletBC(U.sub(0), ...), wave_2D.py:413-422): vertex ids and value weights w, the
 * prescribed value being w * g(t) (w = 1 or the analytical left profile on the left side, 0 on the right side) */
int phw_mesh_set_dirichlet(phw_mesh_t m, int64_t n, const int32_t* vertex_ids, const double* weights);
/* layout introspection for tests: permutations new->old for vertices/edges, cell->patch */
int phw_mesh_get_numbering(phw_mesh_t m, int32_t* vertex_new2old, int32_t* edge_new2old, int32_t* cell_patch);
/* per-patch counts [n_patches,6]: own cells, edge-op cells, vertex-op cells, owned vertices, owned edges, ghosts(v+e) */
int phw_mesh_get_patch_counts(phw_mesh_t m, int32_t* counts);
/* dynamic shared memory (bytes) the bulk-copy pipelined kernels need for this layout: bytes6 = {M_p solve, M_q solve, fused
 * energy reduction, edge-row right-hand side, vertex-row right-hand side, whole-loop kernel with the mass solves resident in shared
 * memory (flags bit 15; resident when <= 224 KB)}
 * (phw_params.flags bits 9, 8, 11, 12, 12, 15; a pipelined variant is available when its figure is <= 227 KB - 1 KB).  Host-only. */
int phw_mesh_pipe_smem(phw_mesh_t m, int64_t* bytes6);
/* host-only structural invariants of the patch layout, and the assembled (ELL) mass rows
 * (tests; small meshes) */
int phw_mesh_self_check(phw_mesh_t m);
int phw_mesh_destroy(phw_mesh_t m);

/* ---- solver: replaces assemble(a)/assemble(b)/solve/assemble(energy) of wave_2D.py:579-595,787-982 ---- */
int phw_solver_create(phw_mesh_t m, const phw_params* prm, int32_t device, void* cuda_stream, phw_solver_t* out);
int phw_solver_destroy(phw_solver_t s);
/* per-case K_wave[n], rho[n] and lumped constants lumped6[n][6] = (Bl, R1, R2, K_em, m_em, L_em) (wave_2D.py:153-158) */
int phw_solver_set_group_params(phw_solver_t s, int32_t n_groups, const double* K_wave, const double* rho, const double* lumped6);
/* lumped states of every case: x3_all[n_groups][3] */
int phw_get_group_states(phw_solver_t s, double* x3_all);
int phw_set_state(phw_solver_t s, const double* p, const double* q, const double* x3);   /* u_m, wave_2D.py:738,768 */
int phw_get_state(phw_solver_t s, double* p, double* q, double* x3);
/* run n_steps of the loop wave_2D.py:787-982 on the device; H_rows [n_steps, 8|12] host or device */
int phw_run(phw_solver_t s, int64_t n_steps, double* H_rows);
/* field output (xdmfFile_p.write(out_p, t) at every step, wave_2D.py:646-651,865-866): after every `every` steps of the following
 * phw_run calls the p field (user numbering, n_vertices values; higher-order solvers: all n_p dofs, vertices first) is copied to
 * snaps[k][n], k = 0 .. capacity-1, counted from the start of each phw_run.  The copy goes through a ring of four device slots and a
 * second stream, so it overlaps the steps that follow; snaps may be host (pinned for a truly asynchronous copy) or device memory and
 * is complete when phw_run returns.  every = 0 switches snapshots off. */
int phw_set_snapshots(phw_solver_t s, int64_t every, double* snaps, int64_t capacity);
int phw_get_stats(phw_solver_t s, phw_stats* out);
int phw_synchronize(phw_solver_t s);
/* simulation clock (t of wave_2D.py:790-799,833-834; 0 after creation): restart / continue a run from a stored state */
int phw_set_time(phw_solver_t s, double t);
/* diagnostics: relative residuals ||b - A x_k|| / ||b||, k = 0..n-1 (n <= 48), of the most recent solve of kind
 * `which` (PHW_P, PHW_Q, 2 = block system) and its iteration count; entries beyond it are stale */
/* diagnostics: a pass that only streams what one iteration of the ELL M_p solve reads and writes, patch by patch, with
 * `threads` per CTA and `ctas_per_sm` resident CTAs: the bandwidth the memory system delivers for this layout */
int phw_stream_probe(phw_solver_t s, int32_t threads, int32_t ctas_per_sm, int32_t reps, double* ms_per_pass, double* bytes_per_pass);
int phw_residual_history(phw_solver_t s, int32_t which, int32_t n, double* rel_residuals, int32_t* n_iterations);
/* diagnostics of the inter-GPU exchange (PHW_XCH_DEBUG bit 6 set when the solver was created): globaltimer stamps [2 fields][32
 * iterations][5] of the most recent pipelined M_p / M_q solve: CTA 0's sweep done, grid barrier passed, local norms reduced, norms of
 * all ranks received, latest arrival of any CTA at the grid barrier (ns) */
int phw_debug_xch_times(phw_solver_t s, uint64_t* stamps320);

/* ---- multi-GPU (one process per GPU on one node).  The exchange arena of every rank (Chebyshev buffers +
 *      mailbox) is shared through CUDA IPC; halo values and residual norms then travel as peer-memory stores
 *      issued from inside the persistent solve kernel.  Replaces PETSc's ghost updates / parallel solves. ---- */
/* 64-byte cudaIpcMemHandle of this rank's arena + its local vector lengths (to be all-gathered by the caller) */
int phw_comm_arena_handle(phw_solver_t s, void* handle64, int64_t* n_vertices_local, int64_t* n_edges_local);
/* user (local mesh) dof indices -> positions inside this rank's device vectors (what a neighbour must write to) */
int phw_comm_internal_index(phw_solver_t s, int32_t which, int64_t n, const int32_t* user_idx, int32_t* internal_idx);
/* handles64 [world][64], peer_nv/peer_ne [world]; per neighbour k: owned dofs to push (local user ids) and the
 * position of each of them in the neighbour's vectors (from its phw_comm_internal_index) */
int phw_comm_setup(phw_solver_t s, int32_t rank, int32_t world, const void* handles64, const int64_t* peer_nv,
                   const int64_t* peer_ne, int32_t n_nb, const int32_t* nb_rank,
                   const int64_t* n_send_v, const int32_t* const* send_v_user, const int32_t* const* send_v_remote,
                   const int64_t* n_send_e, const int32_t* const* send_e_user, const int32_t* const* send_e_remote);
/* global spectral bounds of the Jacobi-scaled RT1 mass matrix (all ranks must iterate with the same interval) */
int phw_set_cheb_bounds(phw_solver_t s, double lam_min_q, double lam_max_q);
/* implicit schemes (IE, SM) on a partitioned mesh (the reference runs every scheme under MPI: wave_2D.py:305-319,560-572 with
 * run_main_wave_2D_mpi.sh:1).  phw_solver_create defers the set-up of the coupled block solve; after phw_comm_setup every rank
 * calls phw_comm_finish_implicit with the SAME sigma = largest singular value of P_q^-1/2 D P_p^-1/2 of the global operator
 * (solver.setup_peer_exchange: distributed power iteration through phw_apply_D / phw_apply_DT and phw_get_jacobi).  Collective. */
int phw_get_jacobi(phw_solver_t s, int32_t which, double* inv_diag);     /* 1 / diag of the solved block, user numbering */
int phw_comm_finish_implicit(phw_solver_t s, double sigma);

/* ---- kernel-level entry points (parity tests, profiling) ---- */
/* y[Ne] = D p = (C - N) p            weak coupling block of wave_2D.py:468-472,479-483,503-518 */
int phw_apply_D(phw_solver_t s, const double* p, double* y);
/* y[Nv] = D^T q */
int phw_apply_DT(phw_solver_t s, const double* q, double* y);
/* y = M_p x (which=PHW_P) or M_q x (PHW_Q):  (p v_p dx) / dot(q, v_q) dx blocks */
int phw_mass_matvec(phw_solver_t s, int32_t which, const double* x, double* y);
/* x = M^-1 b to params.tol; *iters receives the iteration count (may be NULL) */
int phw_mass_solve(phw_solver_t s, int32_t which, const double* b, double* x, int32_t* iters);
/* parts2[0] = 0.5 p^T M_p p / rho, parts2[1] = 0.5 K q^T M_q q of the current state (wave_2D.py:746,769,882,907) */
int phw_energy(phw_solver_t s, double* parts2);
/* analytical diagnostics (wave_2D.py:929-965), required when params.analytical = 1: exact solution
 * amp sin(kx x + pi/2) sin(ky y + pi/2) cos(omega t) (:741), reference P4 mass matrix [15x15] and barycentric node
 * coordinates [15x3] of the degree-(k+3) space errornorm interpolates into (:956), exact spatial factor at the
 * vertices [n_vertices] (:958), point probe = 3 vertex ids (-1 = none) + barycentric weights (:946-950) */
int phw_set_analytical(phw_solver_t s, double amp, double kx, double ky, double omega, const double* mass_ref225,
                       const double* lam_nodes45, const double* exact_vertices, const int32_t* pt_vertices3,
                       const double* pt_lam3);
/* ---- higher-order element pairs P_k x RT_k, k <= 3 (basis_order of wave_2D.py:15-20,174-175): the operators of the
 *      forms :449-526 are assembled once by the host (porthamiltonian_fem_b200/fem_ho.py) and handed over as CSR
 *      matrices (int32 indices): M_p [n_p x n_p], M_q [n_q x n_q], D = C - N [n_q x n_p], D^T [n_p x n_q], port vector
 *      b_L [n_q], element-wise (Wathen) bounds of the Jacobi-scaled mass spectra.  Explicit weak-form schemes, with or
 *      without the lumped model.  State vectors are in the caller's dof numbering. ---- */
int phw_ho_create(const phw_params* prm, int32_t device, void* cuda_stream, int64_t n_p, int64_t n_q,
                  const int32_t* Mp_indptr, const int32_t* Mp_idx, const double* Mp_val,
                  const int32_t* Mq_indptr, const int32_t* Mq_idx, const double* Mq_val,
                  const int32_t* D_indptr, const int32_t* D_idx, const double* D_val,
                  const int32_t* DT_indptr, const int32_t* DT_idx, const double* DT_val,
                  const double* b_L, double lam_min_p, double lam_max_p, double lam_min_q, double lam_max_q,
                  phw_solver_t* out);
/* Batched parameter sweep with the CASE INDEX FASTEST (BASELINE configs[4]; the loop over parameter cases of main_wave_2D.py:24-50,
 * 119-140; csrc/kernels_sweep.cuh): n_cases >= 2 independent cases advanced together on ONE set of assembled operators (same arguments
 * as phw_ho_create; the lowest-order pair assembled by fem_ho.py for a sweep).  Every state vector is [n_dof][n_cases] (case index
 * fastest), in phw_set_state / phw_get_state too; per-case wave speed, density and lumped constants through
 * phw_solver_set_group_params, per-case lumped states through phw_get_group_states; phw_run fills H_rows [n_cases][n_steps][8].
 * Explicit weak-form schemes (the sweep runs SE / SV with the lumped model). */
int phw_sweep_create(const phw_params* prm, int32_t device, void* cuda_stream, int32_t n_cases, int64_t n_p, int64_t n_q,
                     const int32_t* Mp_indptr, const int32_t* Mp_idx, const double* Mp_val,
                     const int32_t* Mq_indptr, const int32_t* Mq_idx, const double* Mq_val,
                     const int32_t* D_indptr, const int32_t* D_idx, const double* D_val,
                     const int32_t* DT_indptr, const int32_t* DT_idx, const double* DT_val,
                     const double* b_L, double lam_min_p, double lam_max_p, double lam_min_q, double lam_max_q,
                     phw_solver_t* out);
/* The same pairs MATRIX-FREE (csrc/kernels_ho.cuh; host side porthamiltonian_fem_b200/fem_ref.py): cells with ascending vertex ids
 * mapped from one reference triangle, so that M_p,K = |det J| Mp_hat, C_K = sign(det J) C_hat and M_q,K = G00 A00 + G01 (A01 + A10) +
 * G11 A11 with G = J^T J / |det J|.  cell_vertices / cell_edges [n_cells][3] (edge i opposite local vertex i; edge ids = the
 * canonical (lo, hi) numbering), geom4 [n_cells][4] = (det J, G00, G01, G11), reference matrices row-major (Mp_hat [ndp x ndp], C_hat
 * [ndq x ndp], A00 / A01 + A10 / A11 [ndq x ndq]).  Dofs: P_k [vertices | k-1 per edge along lo -> hi | per cell], RT_k [k moments per
 * edge | k(k-1) per cell] - derived on the device from the ids, nothing is assembled.  The boundary operator N (zero-flux sides,
 * wave_2D.py:465-472) and its transpose come as compact CSR over the rows they touch (row ids, indptr, idx, val); b_L [n_q] as above. */
int phw_ho_create_mf(const phw_params* prm, int32_t device, void* cuda_stream, int32_t k_p, int32_t k_q,
                     int64_t n_vertices, int64_t n_edges, int64_t n_cells,
                     const int32_t* cell_vertices, const int32_t* cell_edges, const double* geom4,
                     const double* Mp_hat, const double* C_hat, const double* A00, const double* A01s, const double* A11,
                     int64_t nN_rows, const int32_t* N_row, const int32_t* N_indptr, const int32_t* N_idx, const double* N_val,
                     int64_t nNT_rows, const int32_t* NT_row, const int32_t* NT_indptr, const int32_t* NT_idx, const double* NT_val,
                     const double* b_L, double lam_min_p, double lam_max_p, double lam_min_q, double lam_max_q, phw_solver_t* out);
/* analytical diagnostics of a higher-order run (wave_2D.py:929-965): cell_dofs [n_cells x ndp], tabulation T [nn x ndp]
 * of the P_k basis at the P_{k+3} nodes, reference mass matrix [nn x nn], exact spatial factor at those nodes
 * [n_cells x nn] and at the P_k dofs [n_p], cell areas, point probe as (dof, weight) pairs */
int phw_ho_set_analytical(phw_solver_t s, double omega, int32_t n_cells, int32_t ndp, int32_t nn, const int32_t* cell_dofs,
                          const double* T, const double* mass_ref, const double* exact_nodes, const double* area,
                          const double* exact_dofs, int32_t nprobe, const int32_t* probe_dofs, const double* probe_w);
/* H_init of wave_2D.py:746,769,771 (subtracted in the energy-residual column) */
int phw_set_H_init(phw_solver_t s, double H_init);
/* b_L^T q of the current state: integral of q.n over the left port (wave_2D.py:534,546,622-631) */
int phw_port_flux(phw_solver_t s, double* flux);

#ifdef __cplusplus
}
#endif
#endif
