"""CPU-side checks of the C ABI library: it loads, exports every symbol include/phwave.h declares, and the
host-only mesh/patch layout builder agrees bit-exactly with the oracle's independent topology."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import fem
from porthamiltonian_fem_b200 import _lib
from porthamiltonian_fem_b200.mesh import rectangle_structured, rectangle_delaunay, generate_mesh, square_one_channel
from porthamiltonian_fem_b200.solver import Mesh, tag_boundary_edges, port_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'phwave.h')).read()
    declared = set(re.findall(r'\b(phw_[a-zA-Z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.phw_version() >= 200


def test_no_cuda_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    v, c = rectangle_structured(4, 2)
    m = Mesh(v, c, patch_cells=16)
    m.set_boundary(tag_boundary_edges(m, 'R', 1.0, 0.25))
    prm = _lib.PhwParams()
    prm.scheme, prm.dt, prm.K_wave, prm.rho = _lib.SCHEMES['SE'], 1e-3, 3.0, 2.0
    from porthamiltonian_fem_b200.solver import Solver
    with pytest.raises(_lib.PhwError):
        Solver(m, prm)


MESHES = [
    ('structured_34x8', lambda: rectangle_structured(34, 8), 'R', 1.0, 0.25, 64),
    ('structured_136x34', lambda: rectangle_structured(136, 34), 'R', 1.0, 0.25, 512),
    ('delaunay_60x15', lambda: rectangle_delaunay(60, 15, seed=3), 'R', 1.0, 0.25, 200),
    ('s1c_20', lambda: square_one_channel(20), 'S_1C', 1.0, 0.1, 128),
    ('tiny_1x1', lambda: rectangle_structured(1, 1), 'R', 1.0, 0.25, 16),
]


@pytest.mark.parametrize('name,make,shape,xL,yL,pc', MESHES)
def test_topology_bit_exact_vs_oracle(name, make, shape, xL, yL, pc):
    v, c = make()
    m = Mesh(v, c, patch_cells=pc)
    topo = fem.Topology(v, c)
    e, ce, cs = m.edges()
    assert np.array_equal(e, topo.edges)
    assert np.array_equal(ce, topo.cell_edges)
    assert np.array_equal(cs, topo.cell_edge_sign)
    ids, sg = m.boundary()
    assert np.array_equal(ids, topo.boundary_edges)
    assert np.array_equal(sg, topo.boundary_sign)
    tags = tag_boundary_edges(m, shape, xL, yL)
    assert np.array_equal(tags, fem.tag_boundary(topo, shape, xL, yL))
    assert not np.any(tags == 4), 'every boundary facet of the reference domains carries a tag'
    # Euler characteristic of a disc
    assert m.n_vertices - m.n_edges + m.n_cells == 1


@pytest.mark.parametrize('name,make,shape,xL,yL,pc', MESHES)
def test_patch_layout_invariants(name, make, shape, xL, yL, pc):
    v, c = make()
    m = Mesh(v, c, patch_cells=pc)
    v2o, e2o, cp = m.numbering()
    assert np.array_equal(np.sort(v2o), np.arange(m.n_vertices))
    assert np.array_equal(np.sort(e2o), np.arange(m.n_edges))
    counts = m.patch_counts()
    assert counts[:, 0].sum() == m.n_cells                  # own cells partition the mesh
    assert counts[:, 3].sum() == m.n_vertices and counts[:, 4].sum() == m.n_edges
    assert np.all(counts[:, 1] >= counts[:, 0]) and np.all(counts[:, 2] >= counts[:, 1])
    assert np.array_equal(np.bincount(cp, minlength=m.n_patches), counts[:, 0])
    # ownership rule: a dof belongs to the lowest patch among its adjacent cells, owned ranges are contiguous
    _, ce, _ = m.edges()
    own_v = np.full(m.n_vertices, 1 << 30)
    own_e = np.full(m.n_edges, 1 << 30)
    np.minimum.at(own_v, m.cells.ravel(), np.repeat(cp, 3))
    np.minimum.at(own_e, ce.ravel(), np.repeat(cp, 3))
    assert np.all(np.diff(own_v[v2o]) >= 0) and np.all(np.diff(own_e[e2o]) >= 0)
    assert np.array_equal(np.bincount(own_v, minlength=m.n_patches), counts[:, 3])


def test_port_weights_newton_cotes():
    v, c = rectangle_structured(8, 4)
    m = Mesh(v, c, patch_cells=16)
    tags = tag_boundary_edges(m, 'R', 1.0, 0.25)
    topo = fem.Topology(v, c)
    f = lambda x, y: 7.0 * np.sin(2 * np.pi * y / 0.25 + np.pi / 2)
    w = port_weights(m, tags, f)
    ops = fem.assemble_wave_operators(topo, fem.tag_boundary(topo, 'R', 1.0, 0.25), left_weight=f)
    ids, sg = m.boundary()
    left = tags == 0
    assert np.allclose(ops['b_L'][ids[left]], sg[left] * w[left], rtol=0, atol=1e-15)
    # a degree-8 rule integrates polynomials of degree <= 8 exactly
    w8 = port_weights(m, tags, lambda x, y: y ** 8)
    ya, yb = v[topo.edges[ids[left], 0], 1], v[topo.edges[ids[left], 1], 1]
    exact = (yb ** 9 - ya ** 9) / 9.0 / (yb - ya)
    assert np.allclose(w8[left], exact, rtol=1e-12)


def test_bad_meshes_rejected():
    v, c = rectangle_structured(2, 2)
    with pytest.raises(_lib.PhwError):
        Mesh(v, np.array([[0, 1, 99]], dtype=np.int32))
    with pytest.raises(_lib.PhwError):
        Mesh(v, np.array([[0, 0, 1]], dtype=np.int32))        # degenerate
    with pytest.raises(ValueError):
        Mesh(v[:, :1], c)
    m = Mesh(v, c, patch_cells=16)
    with pytest.raises(ValueError):
        m.set_boundary(np.zeros(3, dtype=np.int32))


def test_reference_argument_checks():
    """Same print + quit() behaviour as wave_2D.py:63-83 (no GPU needed: checks run before any device work)."""
    from porthamiltonian_fem_b200 import wave_2D_solve
    for kw in [dict(controlType='lqr'), dict(controlType='passivity'), dict(controlType='casimir', interConnection='IC'),
               dict(interConnection='IC', dirichletImp='strong'), dict(interConnection='IC', timeIntScheme='EH'),
               dict(analytical=True, domainShape='S_1C'), dict(timeIntScheme='XX')]:
        with pytest.raises(SystemExit):
            wave_2D_solve(0.1, 2, '/tmp', 4, 1.0, 0.25, verbose=False, **kw)


def test_xdmf_binary_writer(tmp_path):
    import xml.etree.ElementTree as ET
    from porthamiltonian_fem_b200.field_output import XdmfBinaryWriter
    v, c = rectangle_structured(3, 2)
    w = XdmfBinaryWriter(str(tmp_path), 'p', v, c)
    for k in range(3):
        w.write(np.full(v.shape[0], float(k)), 0.1 * (k + 1))
    w.close()
    root = ET.parse(str(tmp_path / 'p.xdmf')).getroot()
    grids = root.findall('./Domain/Grid/Grid')
    assert len(grids) == 3 and grids[2].find('Time').attrib['Value'] == repr(0.1 * 3)
    vals = np.fromfile(str(tmp_path / 'p_values.bin'), dtype='<f8').reshape(3, -1)
    assert np.all(vals[2] == 2.0) and vals.shape[1] == v.shape[0]
    assert np.array_equal(np.fromfile(str(tmp_path / 'p_topology.bin'), dtype='<i4').reshape(-1, 3), c)


@pytest.mark.parametrize('name,make,pc', [
    ('structured_34x8_pc64', lambda: rectangle_structured(34, 8), 64),
    ('structured_20x20_pc128', lambda: rectangle_structured(20, 20, 1.0, 1.0), 128),
    ('delaunay_40x10_pc100', lambda: rectangle_delaunay(40, 10, seed=4), 100),
    ('s1c_20_pc96', lambda: square_one_channel(20), 96),
    ('single_patch', lambda: rectangle_structured(6, 3), 4096),
    ('tiny_2x1', lambda: rectangle_structured(2, 1), 16),
])
def test_layout_self_check_ell_rows(name, make, pc):
    """ELL mass rows (neighbour lists, shared cells consistent with the mesh) - the structure behind vertex_ell_pass."""
    v, c = make()
    m = Mesh(v, c, patch_cells=pc)
    m.self_check()


def test_pipelined_solve_shared_memory_budget_is_host_computable():
    """phw_mesh_pipe_smem (host-only): the bulk-copy pipelined solves (flags bits 8 / 9) need two stages of the largest
    patch; 1024-cell patches fit the 227 KB of an SM, 2048-cell patches do not (M_q)."""
    v, c = rectangle_structured(400, 100)
    limit = 227 * 1024 - 1024
    m = Mesh(v, c, patch_cells=1024)
    sp, sq, se, sre, srv, sq2 = m.pipe_smem()
    assert 0 < sq2 <= 224 * 1024                 # whole-loop kernel with resident solves: 1024-cell patches fit one SM
    assert 0 < sp <= limit // 2 and 0 < sq <= limit and 0 < se <= limit and 0 < sre <= limit // 2 and 0 < srv <= limit // 2
    assert all(x % 16 == 0 for x in (sp, sq, se, sre, srv))
    m.close()
    m = Mesh(v, c, patch_cells=2048)
    sp2, sq2 = m.pipe_smem()[:2]
    assert sq2 > limit and sp2 > sp
    m.close()


def test_mesh_files_round_trip(tmp_path):
    """load_mesh: .npz, .npy pair, XDMF with inline-XML and raw-binary heavy data (the stand-in for reading an mshr mesh
    written by dolfin on another machine, wave_2D.py:93-119); HDF5 heavy data is rejected with a clear message."""
    from porthamiltonian_fem_b200.mesh import save_mesh, load_mesh
    v, c = rectangle_delaunay(12, 3, seed=2)
    save_mesh(str(tmp_path / 'm.npz'), v, c)
    v2, c2 = load_mesh(str(tmp_path / 'm.npz'))
    assert np.array_equal(v, v2) and np.array_equal(c, c2) and c2.dtype == np.int32
    np.save(tmp_path / 'v.npy', v)
    np.save(tmp_path / 'c.npy', c.astype(np.int64))
    v3, c3 = load_mesh(str(tmp_path / 'v.npy'), str(tmp_path / 'c.npy'))
    assert np.array_equal(v, v3) and np.array_equal(c, c3)
    # XDMF, inline XML, 3-component geometry with z = 0 (what dolfin writes for 2D meshes in some versions)
    xml = ('<?xml version="1.0"?><Xdmf Version="3.0"><Domain><Grid Name="mesh" GridType="Uniform">'
           '<Topology NumberOfElements="{nc}" TopologyType="Triangle" NodesPerElement="3"><DataItem Dimensions="{nc} 3" NumberType="UInt" Format="XML">{cells}</DataItem></Topology>'
           '<Geometry GeometryType="XYZ"><DataItem Dimensions="{nv} 3" Format="XML">{verts}</DataItem></Geometry></Grid></Domain></Xdmf>')
    v3d = np.concatenate([v, np.zeros((v.shape[0], 1))], axis=1)
    (tmp_path / 'a.xdmf').write_text(xml.format(nc=c.shape[0], nv=v.shape[0], cells=' '.join(map(str, c.ravel())),
                                                verts=' '.join(repr(float(x)) for x in v3d.ravel())))
    v4, c4 = load_mesh(str(tmp_path / 'a.xdmf'))
    assert np.array_equal(v, v4) and np.array_equal(c, c4)
    # XDMF, raw binary heavy data
    c.astype(np.int32).tofile(tmp_path / 'b_cells.bin')
    v.tofile(tmp_path / 'b_verts.bin')
    xmlb = ('<?xml version="1.0"?><Xdmf Version="3.0"><Domain><Grid Name="mesh" GridType="Uniform">'
            '<Topology TopologyType="Triangle"><DataItem Dimensions="{nc} 3" NumberType="Int" Precision="4" Format="Binary">b_cells.bin</DataItem></Topology>'
            '<Geometry GeometryType="XY"><DataItem Dimensions="{nv} 2" NumberType="Float" Precision="8" Format="Binary">b_verts.bin</DataItem></Geometry></Grid></Domain></Xdmf>')
    (tmp_path / 'b.xdmf').write_text(xmlb.format(nc=c.shape[0], nv=v.shape[0]))
    v5, c5 = load_mesh(str(tmp_path / 'b.xdmf'))
    assert np.array_equal(v, v5) and np.array_equal(c, c5)
    (tmp_path / 'h.xdmf').write_text(xmlb.replace('Format="Binary"', 'Format="HDF"').format(nc=c.shape[0], nv=v.shape[0]))
    with pytest.raises(ValueError, match='h5py'):
        load_mesh(str(tmp_path / 'h.xdmf'))
    with pytest.raises(ValueError):
        load_mesh(str(tmp_path / 'm.txt'))
    # the loaded mesh goes through the layout builder like a generated one
    m = Mesh(v4, c4, patch_cells=32)
    assert m.n_cells == c.shape[0]
    m.close()
