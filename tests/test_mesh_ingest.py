"""Mesh ingest (utils/mixed_dim_problem.py:634-681) without libhdf5: the built-in HDF5 reader against a file written by
libhdf5 itself, the reader against the fixture writer (contiguous, chunked, shuffle + deflate), the XDMF layer in both
layouts the reference distinguishes, and ProblemKNPEMI reading an existing file instead of generating the fixture."""
import importlib
import os
import numpy as np
import pytest

from test_host_logic import BASE, write


@pytest.fixture(scope="module")
def mods(kb):
    return (importlib.import_module(kb.__name__ + ".hdf5_min"), importlib.import_module(kb.__name__ + ".xdmf"),
            importlib.import_module(kb.__name__ + ".mesh"))


def test_reader_against_a_file_written_by_libhdf5(mods):
    """scipy ships a MATLAB 7.3 file = HDF5 (superblock 0, 512-byte user block, symbol-table root group, object header
    version 1, contiguous IEEE doubles) written by libhdf5 1.8: the same variant DOLFINx writes by default."""
    h5 = mods[0]
    import scipy.io
    path = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy test data not installed")
    with h5.File(path) as f:
        assert f.keys("/") == ["testdouble"]
        a = f["/testdouble"]
        assert a.dtype == np.float64 and a.shape == (9, 1)
        assert np.array_equal(a.ravel(), np.arange(9) * (np.pi / 4))       # scipy's own expectation for this fixture
        assert "/testdouble" in f and "/nothing" not in f
        with pytest.raises(KeyError):
            f["/nothing"]


@pytest.mark.parametrize("chunk_rows,compress", [(None, False), (128, False), (128, True), (17, True)])
def test_reader_writer_round_trip(mods, tmp_path, chunk_rows, compress):
    h5 = mods[0]
    rng = np.random.default_rng(3)
    data = {"/Mesh/mesh/topology": rng.integers(0, 5000, (1000, 4)), "/Mesh/mesh/geometry": rng.random((333, 3)),
            "/MeshTags/ct/Values": rng.integers(0, 9, (1000, 1)).astype(np.int32), "/f4": rng.random((50, 2)).astype(np.float32),
            "/u1": np.arange(200, dtype=np.uint8), "/empty": np.zeros((0, 3), np.int64)}
    p = str(tmp_path / "t.h5")
    h5.write_file(p, data, chunk_rows=chunk_rows, compress=compress)
    with h5.File(p) as f:
        assert f.keys("/") == ["Mesh", "MeshTags", "empty", "f4", "u1"] and f.keys("/Mesh/mesh") == ["geometry", "topology"]
        for k, v in data.items():
            r = f[k]
            assert r.dtype == v.dtype and r.shape == v.shape and np.array_equal(r, v), k
    with open(p, "r+b") as fh:      # not an HDF5 file any more
        fh.write(b"\0" * 8)
    with pytest.raises(h5.Hdf5FormatError, match="not an HDF5 file"):
        h5.File(p)


@pytest.mark.parametrize("fmt", ["HDF", "XML"])
@pytest.mark.parametrize("gdim", [2, 3])
def test_xdmf_dolfinx_layout_reproduces_the_fixture(mods, tmp_path, fmt, gdim):
    """generate_square_mesh.py layout: mesh + grid "ct" in one file, mesh + grid "ft" in the other."""
    M = mods[2]
    m = M.unit_square_fixture(16, 1e-6) if gdim == 2 else M.unit_cube_fixture(6, 1e-6)
    a, b = str(tmp_path / "square.xdmf"), str(tmp_path / "square_facets.xdmf")
    M.export_xdmf(m, a, b, scale=1e-6, fmt=fmt)
    r = M.from_xdmf(a, b, "ct", "ft", (1,), 2, (3,), 1e-6)
    assert r.gdim == gdim and np.array_equal(r.cells, m.cells) and np.array_equal(r.cell_tags, m.cell_tags)
    assert np.allclose(r.x, m.x, rtol=4e-16, atol=0)
    assert np.array_equal(r.mf_verts, m.mf_verts) and np.array_equal(r.mf_tags, m.mf_tags)
    assert np.array_equal(r.bc_verts, M.boundary_vertices(m))
    # exterior vertices without the structured index (facets that belong to one cell)
    plain = M.Mesh(m.gdim, m.x, m.cells, m.cell_tags, m.intra_tags, m.extra_tag, m.mf_verts, m.mf_tags)
    assert np.array_equal(M.boundary_vertices(plain), M.boundary_vertices(m))


def test_xdmf_single_grid_layout_and_permuted_tags(mods, tmp_path):
    """"Tags under the same hierarchy as the mesh" (mixed_dim_problem.py:142-145): one grid "mesh" with the values as a cell
    attribute; the facet file is a grid of facets, here shuffled and with a subset of the facets only."""
    h5, X, M = mods
    m = M.unit_cube_fixture(4, 1.0)
    rng = np.random.default_rng(0)
    nf = m.mf_verts.shape[0]
    perm = rng.permutation(nf)[: nf - 5]                      # five membrane facets carry no tag in the file
    fent = m.mf_verts[perm][:, rng.permutation(3)]            # vertex order inside a facet is arbitrary
    fval = np.where(np.arange(perm.size) % 2 == 0, 4, 7).astype(np.int32)

    def grid(name, ttype, n, k, topo, vals):
        return (f'<Grid Name="{name}" GridType="Uniform"><Topology TopologyType="{ttype}" NumberOfElements="{n}">'
                f'<DataItem Dimensions="{n} {k}" NumberType="Int" Format="HDF">d.h5:{topo}</DataItem></Topology>'
                f'<Geometry GeometryType="XYZ"><DataItem Dimensions="{m.x.shape[0]} 3" Format="HDF">d.h5:/geo</DataItem></Geometry>'
                f'<Attribute Name="f" Center="Cell"><DataItem Dimensions="{n}" NumberType="Int" Format="HDF">d.h5:{vals}</DataItem>'
                f'</Attribute></Grid>')

    h5.write_file(str(tmp_path / "d.h5"), {"/geo": m.x, "/cells": m.cells.astype(np.int64), "/ct": m.cell_tags,
                                           "/facets": fent.astype(np.int64), "/ft": fval}, chunk_rows=100, compress=True)
    head = '<?xml version="1.0"?>\n<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>\n<Xdmf Version="3.0"><Domain>'
    (tmp_path / "tissue.xdmf").write_text(head + grid("mesh", "Tetrahedron", m.cells.shape[0], 4, "/cells", "/ct") + "</Domain></Xdmf>")
    (tmp_path / "tissue_facets.xdmf").write_text(head + grid("mesh", "Triangle", perm.size, 3, "/facets", "/ft") + "</Domain></Xdmf>")
    r = M.from_xdmf(str(tmp_path / "tissue.xdmf"), str(tmp_path / "tissue_facets.xdmf"), "mesh", "mesh", (1,), 2, (), 1.0)
    assert np.array_equal(r.cells, m.cells) and np.array_equal(r.cell_tags, m.cell_tags) and np.array_equal(r.x, m.x)
    assert np.array_equal(r.mf_verts, m.mf_verts)             # geometric interface, sorted like the fixture
    want = np.full(nf, -1, np.int32)
    want[perm] = fval
    assert np.array_equal(r.mf_tags, want) and r.bc_verts.size == 0
    with pytest.raises(X.XdmfError, match="no grid named"):
        X.read_xdmf_mesh(str(tmp_path / "tissue.xdmf"), str(tmp_path / "tissue_facets.xdmf"), "ct", "ft")


def test_problem_reads_an_existing_mesh_file(kb, mods, tmp_path):
    """A mesh file that exists is read -- also when its name looks like a fixture name: the problem built
    from the file has the dof maps of the file's mesh, not of the same-named generated fixture."""
    M = mods[2]
    (tmp_path / "geo").mkdir()
    m = M.unit_square_fixture(12, 1e-6)                       # stored under the name of the N = 8 fixture
    M.export_xdmf(m, str(tmp_path / "geo" / "square8.xdmf"), str(tmp_path / "geo" / "square8_facets.xdmf"), scale=1e-6)
    text = BASE.replace("./input/geometries/", "") + f'input_dir: "{tmp_path}/geo/"\n'
    p = kb.ProblemKNPEMI(write(tmp_path, text), verbose=False)
    assert p.mesh.cells.shape[0] == 2 * 12 * 12 and np.array_equal(p.mesh.cells, m.cells)
    assert np.array_equal(p.mesh.mf_verts, m.mf_verts) and np.all(p.mesh.mf_tags == 4)
    assert np.array_equal(p.dofs_intra, np.unique(m.cells[m.cell_tags == 1])) and p.mesh.grid is None
    # without the file the name selects the generated fixture, any other missing file raises
    q = kb.ProblemKNPEMI(write(tmp_path, BASE), verbose=False)
    assert q.mesh.cells.shape[0] == 2 * 8 * 8
    with pytest.raises(RuntimeError, match="does not exist"):
        kb.ProblemKNPEMI(write(tmp_path, BASE.replace("square8.xdmf", "tissue.xdmf")), verbose=False)
