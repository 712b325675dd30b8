"""Field output on the GPU (reference: xdmfFile_p.write(out_p, t) and xdmfFile_p_exact.write(p_e, t) at every step,
wave_2D.py:646-651,865-866,964-965): the device-side snapshot ring of phw_set_snapshots against states read back from chunked runs
(bit-exact), and the p.xdmf / p_exact.xdmf trees wave_2D_solve writes."""
import os

import numpy as np
import pytest

from porthamiltonian_fem_b200 import _lib, wave_2D_solve
from porthamiltonian_fem_b200.mesh import rectangle_structured
from porthamiltonian_fem_b200.solver import Mesh, Solver, tag_boundary_edges

pytestmark = pytest.mark.gpu
K_WAVE, RHO = 3.0, 2.0


def make_solver(mesh, scheme, flags, patch_cells):
    m = Mesh(mesh[0], mesh[1], patch_cells=patch_cells)
    m.set_boundary(tag_boundary_edges(m, 'R', 1.0, 0.25))
    prm = _lib.PhwParams()
    prm.scheme, prm.dt, prm.K_wave, prm.rho = _lib.SCHEMES[scheme], 5e-4, K_WAVE, RHO
    prm.input_kind, prm.in_amp, prm.in_omega, prm.in_t_stop = _lib.IN_SINE_WINDOW, 10.0, 8 * np.pi, 0.25
    prm.tol, prm.max_iter, prm.extrap_order, prm.flags = 1e-13, 400, 3, flags
    return m, Solver(m, prm)


@pytest.mark.parametrize('scheme,flags,patch_cells', [('SE', 0, 256), ('SV', 0, 256), ('SE', _lib.FLAGS_LOOP, 288), ('SV', _lib.FLAGS_LOOP, 288)])
def test_snapshot_ring_equals_chunked_state_reads(scheme, flags, patch_cells):
    mesh = rectangle_structured(48, 12)          # 1152 cells
    every, n_steps = 3, 31                       # 10 snapshots: the ring of 4 device slots is reused twice; a tail of 1 step
    m, s = make_solver(mesh, scheme, flags, patch_cells)
    H, snaps = s.run_with_snapshots(n_steps, every)
    assert snaps.shape == (n_steps // every, m.n_vertices)
    assert bool(s.stats()['kernel_paths'] & 256) == bool(flags & _lib.FLAGS_LOOP)
    p_end, q_end, _ = s.get_state()
    s.close(); m.close()
    m2, s2 = make_solver(mesh, scheme, flags, patch_cells)
    H2 = np.zeros_like(H)
    for k in range(n_steps // every):
        H2[k * every:(k + 1) * every] = s2.run(every)
        p_k, _, _ = s2.get_state()
        assert np.array_equal(p_k, snaps[k]), (scheme, k)
    H2[(n_steps // every) * every:] = s2.run(n_steps % every)
    p2, q2, _ = s2.get_state()
    assert np.array_equal(p2, p_end) and np.array_equal(q2, q_end)
    assert np.array_equal(H2, H)
    s2.close(); m2.close()


def read_values(out_dir, name, nv):
    v = np.fromfile(os.path.join(out_dir, name + '_values.bin'), dtype='<f8')
    return v.reshape(-1, nv)


def test_wave_2D_solve_writes_p_and_p_exact(tmp_path):
    mesh = rectangle_structured(32, 8)
    nv = mesh[0].shape[0]
    out = str(tmp_path / 'an')
    n_steps, every, tF = 40, 4, 0.02
    H, nc = wave_2D_solve(tF, n_steps, out, 0, 1.0, 0.25, timeIntScheme='SE', K_wave=K_WAVE, rho=RHO, mesh=mesh, analytical=True,
                          saveP=True, save_every=every, verbose=False)
    Href, _ = wave_2D_solve(tF, n_steps, out + '_none', 0, 1.0, 0.25, timeIntScheme='SE', K_wave=K_WAVE, rho=RHO, mesh=mesh, analytical=True,
                            saveP=False, verbose=False)
    assert np.array_equal(H, Href)               # saving fields does not change the run
    for name in ('p', 'p_exact'):
        assert os.path.exists(os.path.join(out, name + '.xdmf'))
        xml = open(os.path.join(out, name + '.xdmf')).read()
        assert xml.count('<Time Value=') == n_steps // every
    p = read_values(out, 'p', nv)
    pe = read_values(out, 'p_exact', nv)
    assert p.shape == pe.shape == (n_steps // every, nv)
    # the exact field at the snapshot times (wave_2D.py:734-741) and the nodal max error of column 9 (wave_2D.py:958)
    X, Y = mesh[0][:, 0], mesh[0][:, 1]
    omega = np.sqrt(K_WAVE / RHO) * np.sqrt(np.pi ** 2 / 4 + 4 * np.pi ** 2 / 0.25 ** 2)
    for k in range(n_steps // every):
        row = (k + 1) * every - 1
        t = H[row, 0]
        exact = RHO * omega * np.sin(np.pi * X / 2 + np.pi / 2) * np.sin(2 * np.pi * Y / 0.25 + np.pi / 2) * np.cos(omega * t)
        assert np.allclose(pe[k], exact, rtol=0, atol=1e-12 * np.abs(exact).max())
        assert abs(np.abs(pe[k] - p[k]).max() - H[row, 9]) <= 1e-9 * max(1.0, H[row, 9])
    assert not os.path.exists(os.path.join(out + '_none', 'p.xdmf'))


def test_high_order_field_output(tmp_path):
    mesh = rectangle_structured(16, 8)
    out = str(tmp_path / 'ho')
    H, nc = wave_2D_solve(0.02, 20, out, 0, 1.0, 0.25, timeIntScheme='SV', K_wave=K_WAVE, rho=RHO, mesh=mesh, analytical=True,
                          basis_order=(2, 2), saveP=True, save_every=5, verbose=False)
    nv = mesh[0].shape[0]
    p, pe = read_values(out, 'p', nv), read_values(out, 'p_exact', nv)
    assert p.shape == pe.shape == (4, nv)
    assert np.abs(p - pe).max() < 0.1 * np.abs(pe).max()        # P2 nodal values follow the exact field on this coarse mesh
