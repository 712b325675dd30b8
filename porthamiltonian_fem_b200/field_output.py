"""Field output for post-processing: the stand-in for the reference's ``XDMFFile(...).write(out_p, t)``
(wave_2D.py:646-651,865-866).  dolfin writes XDMF + HDF5; h5py is not available here, so the heavy data is written
as raw little-endian binary files that the XDMF file references with ``Format="Binary"`` (readable by ParaView /
VisIt).  One temporal collection, P1 nodal values on the triangle mesh."""
import os

import numpy as np


class XdmfBinaryWriter:
    def __init__(self, out_dir, name, vertices, cells):
        os.makedirs(out_dir, exist_ok=True)
        self.dir, self.name = out_dir, name
        self.nv, self.nc = int(vertices.shape[0]), int(cells.shape[0])
        np.ascontiguousarray(vertices, dtype='<f8').tofile(os.path.join(out_dir, name + '_geometry.bin'))
        np.ascontiguousarray(cells, dtype='<i4').tofile(os.path.join(out_dir, name + '_topology.bin'))
        self.data = open(os.path.join(out_dir, name + '_values.bin'), 'wb')
        self.times = []

    def write(self, values, t, flush=True):
        v = np.ascontiguousarray(values, dtype='<f8')
        if v.shape[0] != self.nv:
            raise ValueError('one value per vertex expected')
        v.tofile(self.data)
        self.times.append(float(t))
        if flush:
            self._write_xml()        # keep the index file valid after every snapshot (flush_output, wave_2D.py:648)

    def _write_xml(self):
        self.data.flush()
        n = self.name
        grids = []
        for i, t in enumerate(self.times):
            grids.append(
                '      <Grid Name="{n}_{i}" GridType="Uniform">\n'
                '        <Time Value="{t!r}"/>\n'
                '        <Topology TopologyType="Triangle" NumberOfElements="{nc}">\n'
                '          <DataItem Format="Binary" DataType="Int" Precision="4" Endian="Little" Dimensions="{nc} 3">{n}_topology.bin</DataItem>\n'
                '        </Topology>\n'
                '        <Geometry GeometryType="XY">\n'
                '          <DataItem Format="Binary" DataType="Float" Precision="8" Endian="Little" Dimensions="{nv} 2">{n}_geometry.bin</DataItem>\n'
                '        </Geometry>\n'
                '        <Attribute Name="{n}" AttributeType="Scalar" Center="Node">\n'
                '          <DataItem Format="Binary" DataType="Float" Precision="8" Endian="Little" Seek="{seek}" Dimensions="{nv}">{n}_values.bin</DataItem>\n'
                '        </Attribute>\n'
                '      </Grid>\n'.format(n=n, i=i, t=t, nc=self.nc, nv=self.nv, seek=8 * self.nv * i))
        xml = ('<?xml version="1.0"?>\n<Xdmf Version="3.0">\n  <Domain>\n'
               '    <Grid Name="TimeSeries" GridType="Collection" CollectionType="Temporal">\n' + ''.join(grids) +
               '    </Grid>\n  </Domain>\n</Xdmf>\n')
        with open(os.path.join(self.dir, n + '.xdmf'), 'w') as f:
            f.write(xml)

    def close(self):
        self._write_xml()
        self.data.close()
