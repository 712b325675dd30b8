"""ctypes binding of libphwave.so (include/phwave.h).  No torch types cross this boundary: arrays are
passed as raw pointers (numpy ``ctypes.data`` or ``torch.Tensor.data_ptr()``)."""
import ctypes
import os

import numpy as np

from . import build as _build

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i8p = ctypes.POINTER(ctypes.c_int8)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f64p = ctypes.POINTER(ctypes.c_double)

PHW_OK, PHW_EINVAL, PHW_ECUDA, PHW_ENOMEM, PHW_ESTATE, PHW_ENOCONV = 0, -1, -2, -3, -4, -5
SCHEMES = {'EE': 0, 'IE': 1, 'SE': 2, 'SM': 3, 'EH': 4, 'SV': 5}
IN_SINE_WINDOW, IN_SINE, IN_COSINE, IN_PASSIVITY = 0, 1, 2, 3
PHW_P, PHW_Q = 0, 1
ABI_VERSION = 200          # phw_version() of the library this binding was written against


class PhwParams(ctypes.Structure):
    _fields_ = [('scheme', ctypes.c_int32), ('strong', ctypes.c_int32), ('interconnection', ctypes.c_int32),
                ('casimir', ctypes.c_int32), ('analytical', ctypes.c_int32), ('input_kind', ctypes.c_int32),
                ('dt', ctypes.c_double), ('K_wave', ctypes.c_double), ('rho', ctypes.c_double),
                ('in_amp', ctypes.c_double), ('in_omega', ctypes.c_double), ('in_t_stop', ctypes.c_double),
                ('in_t_ctrl', ctypes.c_double), ('in_gain', ctypes.c_double),
                ('Bl_em', ctypes.c_double), ('R1_em', ctypes.c_double), ('R2_em', ctypes.c_double),
                ('K_em', ctypes.c_double), ('m_em', ctypes.c_double), ('L_em', ctypes.c_double),
                ('H_init', ctypes.c_double),
                ('an_omega', ctypes.c_double), ('an_xLength', ctypes.c_double), ('an_yLength', ctypes.c_double),
                ('an_xPoint', ctypes.c_double), ('an_yPoint', ctypes.c_double),
                ('tol', ctypes.c_double), ('max_iter', ctypes.c_int32), ('extrap_order', ctypes.c_int32), ('flags', ctypes.c_int32),
                ('extrap_order_q', ctypes.c_int32)]


# phw_params.flags: bulk-copy pipelined M_q / M_p solves, fused energy reduction, right-hand sides (include/phwave.h)
FLAGS_PIPELINED = 256 | 512 | 2048 | 4096
# the whole time loop in one launch (csrc/kernels_loop.cuh): cache-resident meshes, one CTA per patch, one cluster per case
FLAGS_LOOP = 32768


def loop_patch_cells(n_cells, max_patch_cells=3000):
    """patch size that cuts a mesh of n_cells into 1, 2, 4, 8 or 16 patches of <= max_patch_cells (one thread-block cluster per case),
    aiming at ~1000 cells per CTA; None if the mesh is too large for one cluster (the caller then uses 2048-cell patches: one
    co-resident grid) or too small to be cut evenly"""
    force = int(os.environ.get('PHW_LOOP_CL', '0'))       # measurement knob: force the cluster size
    for cl in ((force,) if force else (1, 2, 4, 8, 16)):
        pc = max(16, -(-n_cells // cl))
        if -(-n_cells // pc) != cl:
            continue
        if pc <= 1100 or (cl == 16 and pc <= max_patch_cells) or force:
            return pc
    return None


class PhwStats(ctypes.Structure):
    _fields_ = [('steps', ctypes.c_int64), ('solves_p', ctypes.c_int64), ('solves_q', ctypes.c_int64),
                ('solves_block', ctypes.c_int64), ('iters_p', ctypes.c_int64), ('iters_q', ctypes.c_int64),
                ('iters_block', ctypes.c_int64), ('max_rel_residual', ctypes.c_double),
                ('device_ms', ctypes.c_double), ('kernel_launches', ctypes.c_int64), ('bytes_model', ctypes.c_int64),
                ('lam_min_q', ctypes.c_double), ('lam_max_q', ctypes.c_double), ('skew_bound', ctypes.c_double),
                ('ms_solve_p', ctypes.c_double), ('ms_solve_q', ctypes.c_double), ('ms_solve_block', ctypes.c_double),
                ('kernel_paths', ctypes.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/phwave.h declares: (restype, argtypes)
SIGNATURES = {
    'phw_last_error': (ctypes.c_char_p, []),
    'phw_version': (ctypes.c_int, []),
    'phw_sizeof': (ctypes.c_int, [ctypes.c_int]),
    'phw_mesh_create': (ctypes.c_int, [ctypes.c_int64, c_f64p, ctypes.c_int64, c_i32p, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]),
    'phw_mesh_create_partitioned': (ctypes.c_int, [ctypes.c_int64, c_f64p, ctypes.c_int64, c_i32p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    'phw_mesh_set_cell_owned': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'phw_mesh_create_grouped': (ctypes.c_int, [ctypes.c_int64, c_f64p, ctypes.c_int64, c_i32p, ctypes.c_int32, c_i32p, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]),
    'phw_mesh_sizes': (ctypes.c_int, [ctypes.c_void_p, c_i64p, c_i64p, c_i64p, c_i64p, c_i64p]),
    'phw_mesh_get_edges': (ctypes.c_int, [ctypes.c_void_p, c_i32p, c_i32p, c_i8p]),
    'phw_mesh_get_boundary': (ctypes.c_int, [ctypes.c_void_p, c_i32p, c_i8p]),
    'phw_mesh_set_boundary': (ctypes.c_int, [ctypes.c_void_p, c_i32p, c_f64p]),
    'phw_mesh_set_dirichlet': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_i32p, c_f64p]),
    'phw_mesh_get_numbering': (ctypes.c_int, [ctypes.c_void_p, c_i32p, c_i32p, c_i32p]),
    'phw_mesh_self_check': (ctypes.c_int, [ctypes.c_void_p]),
    'phw_mesh_get_patch_counts': (ctypes.c_int, [ctypes.c_void_p, c_i32p]),
    'phw_mesh_pipe_smem': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]),
    'phw_mesh_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'phw_solver_create': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(PhwParams), ctypes.c_int32, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    'phw_solver_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'phw_solver_set_group_params': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, c_f64p, c_f64p, c_f64p]),
    'phw_get_group_states': (ctypes.c_int, [ctypes.c_void_p, c_f64p]),
    'phw_set_state': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'phw_get_state': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'phw_run': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    'phw_get_stats': (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(PhwStats)]),
    'phw_set_snapshots': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64]),
    'phw_synchronize': (ctypes.c_int, [ctypes.c_void_p]),
    'phw_comm_arena_handle': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_i64p, c_i64p]),
    'phw_comm_internal_index': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, c_i32p, c_i32p]),
    'phw_comm_setup': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, c_i64p, c_i64p, ctypes.c_int32, c_i32p,
                                      c_i64p, ctypes.POINTER(c_i32p), ctypes.POINTER(c_i32p), c_i64p, ctypes.POINTER(c_i32p), ctypes.POINTER(c_i32p)]),
    'phw_set_cheb_bounds': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]),
    'phw_get_jacobi': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    'phw_comm_finish_implicit': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double]),
    'phw_apply_D': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'phw_apply_DT': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'phw_mass_matvec': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    'phw_mass_solve': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, c_i32p]),
    'phw_energy': (ctypes.c_int, [ctypes.c_void_p, c_f64p]),
    'phw_set_analytical': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_f64p, c_f64p, c_f64p, c_i32p, c_f64p]),
    'phw_ho_create': (ctypes.c_int, [ctypes.POINTER(PhwParams), ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                     c_i32p, c_i32p, c_f64p, c_i32p, c_i32p, c_f64p, c_i32p, c_i32p, c_f64p, c_i32p, c_i32p, c_f64p,
                                     c_f64p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_void_p)]),
    'phw_sweep_create': (ctypes.c_int, [ctypes.POINTER(PhwParams), ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64,
                                        c_i32p, c_i32p, c_f64p, c_i32p, c_i32p, c_f64p, c_i32p, c_i32p, c_f64p, c_i32p, c_i32p, c_f64p,
                                        c_f64p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_void_p)]),
    'phw_ho_create_mf': (ctypes.c_int, [ctypes.POINTER(PhwParams), ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, c_i32p, c_i32p, c_f64p,
                                        c_f64p, c_f64p, c_f64p, c_f64p, c_f64p,
                                        ctypes.c_int64, c_i32p, c_i32p, c_i32p, c_f64p, ctypes.c_int64, c_i32p, c_i32p, c_i32p, c_f64p,
                                        c_f64p, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_void_p)]),
    'phw_ho_set_analytical': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_i32p,
                                             c_f64p, c_f64p, c_f64p, c_f64p, c_f64p, ctypes.c_int32, c_i32p, c_f64p]),
    'phw_set_time': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double]),
    'phw_stream_probe': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_f64p, c_f64p]),
    'phw_residual_history': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, c_f64p, c_i32p]),
    'phw_set_H_init': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double]),
    'phw_debug_xch_times': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'phw_port_flux': (ctypes.c_int, [ctypes.c_void_p, c_f64p]),
}

_lib = None


class PhwError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('libphwave error {}: {}'.format(code, msg))
        self.code = code


def load(rebuild_if_stale=True):
    """Load libphwave.so (building it in-tree with nvcc when missing or stale).  There is no CPU
    fallback: if the library cannot be built or loaded this raises."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get('PHW_LIB', _build.LIB)     # tuning variants of the same sources (profiles/), default: the in-tree build
    if path == _build.LIB and rebuild_if_stale and (not os.path.exists(path) or _build.needs_build()):
        # a failed rebuild is an error even when an older binary is lying around: stale kernels (and stale struct layouts)
        # must never run silently
        _build.build()
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.phw_version() < ABI_VERSION or lib.phw_sizeof(0) != ctypes.sizeof(PhwParams) or lib.phw_sizeof(1) != ctypes.sizeof(PhwStats):
        raise RuntimeError('libphwave.so at {} does not match this binding (version {}, phw_params {} / {} bytes, phw_stats {} / {} bytes): '
                           'rebuild it (python -m porthamiltonian_fem_b200.build --force)'.format(
                               path, lib.phw_version(), lib.phw_sizeof(0), ctypes.sizeof(PhwParams), lib.phw_sizeof(1), ctypes.sizeof(PhwStats)))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise PhwError(rc, load().phw_last_error().decode('utf-8', 'replace'))


def ptr(arr):
    """raw address of a numpy array or a torch tensor (or None)."""
    if arr is None:
        return None
    if isinstance(arr, np.ndarray):
        return arr.ctypes.data
    return arr.data_ptr()
