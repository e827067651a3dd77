"""XDMF mesh ingest (host side, setup time).

Replaces ``dfx.io.XDMFFile(comm, file, 'r').read_mesh()`` / ``.read_meshtags(mesh, name=...)`` as used by
src/CGx/utils/mixed_dim_problem.py:634-681 for P1 simplex meshes, with the two layouts the reference distinguishes
(:136-145):

  * DOLFINx layout (``square{N}.xdmf`` written by utils/generate_square_mesh.py:28-42, or cell and facet tags in one file):
    a grid "mesh" (Topology + Geometry) and separate grids named ``ct`` / ``ft`` whose Topology lists the tagged entities by
    their geometry nodes and whose Attribute holds the values;
  * everything "under the same hierarchy as the mesh" (emimesh / meshio files): ONE grid named "mesh" with Topology,
    Geometry and the values as a cell Attribute; the facet file is a grid of facets with the same structure.

Both reduce to: entities = the grid's Topology item, values = its first Attribute item; cell values are matched to the
mesh cells and facet values to facets by their (sorted) vertex tuples, like DOLFINx does.  DataItem formats: ``HDF``
(through hdf5_min / h5py), ``XML`` (inline numbers) and ``Binary`` (raw little-endian file).  ``xi:include`` elements
(DOLFINx re-uses the mesh geometry in tag grids that way) need no resolution because tag grids only contribute their
Topology and Attribute.

``write_xdmf_mesh`` writes the DOLFINx layout (used by the tests and to export the generated tissue blocks).
"""
import os
import xml.etree.ElementTree as ET
import numpy as np

from . import hdf5_min

_TOPOLOGY_NODES = {"triangle": 3, "tetrahedron": 4, "polyline": 2, "polyvertex": 1}


class XdmfError(RuntimeError):
    pass


def _local(tag):
    return tag.rsplit("}", 1)[-1]


def _children(el, name):
    return [c for c in el if _local(c.tag) == name]


def _read_item(item, xdmf_path, h5cache):
    fmt = (item.get("Format") or "XML").upper()
    dims = [int(s) for s in (item.get("Dimensions") or "").split()]
    ntype = (item.get("NumberType") or item.get("DataType") or "Float").lower()
    prec = int(item.get("Precision") or (8 if ntype == "float" else 4))
    text = (item.text or "").strip()
    if fmt == "HDF":
        fname, _, dpath = text.partition(":")
        full = os.path.join(os.path.dirname(os.path.abspath(xdmf_path)), fname.strip())
        if full not in h5cache:
            if not os.path.exists(full):
                raise XdmfError(f"{xdmf_path}: heavy-data file {full} does not exist")
            h5cache[full] = hdf5_min.open_file(full)
        a = np.asarray(h5cache[full][dpath.strip()])
    elif fmt == "XML":
        a = np.array(text.split(), dtype=np.float64 if ntype == "float" else np.int64)
    elif fmt == "BINARY":
        kind = {"float": "f", "int": "i", "uint": "u"}.get(ntype)
        if kind is None:
            raise XdmfError(f"{xdmf_path}: NumberType {ntype!r} is not supported")
        order = ">" if (item.get("Endian") or "Little").lower() == "big" else "<"
        full = os.path.join(os.path.dirname(os.path.abspath(xdmf_path)), text)
        a = np.fromfile(full, dtype=np.dtype(f"{order}{kind}{prec}"), offset=int(item.get("Seek") or 0),
                        count=int(np.prod(dims)) if dims else -1)
    else:
        raise XdmfError(f"{xdmf_path}: DataItem Format {fmt!r} is not supported (HDF | XML | Binary)")
    if dims and int(np.prod(dims)) == a.size:
        a = a.reshape(dims)
    return a


def _grids(xdmf_path):
    try:
        root = ET.parse(xdmf_path).getroot()
    except ET.ParseError as e:
        raise XdmfError(f"{xdmf_path}: not a valid XDMF file ({e})") from None
    out = []

    def walk(el):
        for c in el:
            if _local(c.tag) == "Grid":
                if (c.get("GridType") or "Uniform").lower() in ("collection", "tree"):
                    walk(c)
                else:
                    out.append(c)
            elif _local(c.tag) == "Domain":
                walk(c)

    walk(root)
    return out


def _topology(grid, xdmf_path, h5cache):
    topo = _children(grid, "Topology")
    if not topo:
        return None, None
    t = topo[0]
    ttype = (t.get("TopologyType") or t.get("Type") or "").lower()
    if ttype not in _TOPOLOGY_NODES:
        raise XdmfError(f"{xdmf_path}: TopologyType {ttype!r} is not supported (P1 simplices: Triangle, Tetrahedron, PolyLine)")
    items = _children(t, "DataItem")
    if not items:
        raise XdmfError(f"{xdmf_path}: Topology without a DataItem")
    conn = np.asarray(_read_item(items[0], xdmf_path, h5cache))
    return ttype, conn.reshape(-1, _TOPOLOGY_NODES[ttype])


def _values(grid, xdmf_path, h5cache):
    att = _children(grid, "Attribute")
    if not att:
        return None
    items = _children(att[0], "DataItem")
    return np.asarray(_read_item(items[0], xdmf_path, h5cache)).ravel()


def _find_grid(grids, name, xdmf_path, need_values):
    named = [g for g in grids if g.get("Name") == name]
    for g in named:
        if not need_values or _children(g, "Attribute"):
            return g
    if need_values:
        with_values = [g for g in grids if _children(g, "Attribute")]
        if len(with_values) == 1 and not named and name == "mesh":
            return with_values[0]          # meshio names its single grid "Grid"
        raise XdmfError(f"{xdmf_path}: no grid named {name!r} with an Attribute (grids: {[g.get('Name') for g in grids]})")
    geo = [g for g in grids if _children(g, "Geometry") and _children(g, "Topology")]
    if not geo:
        raise XdmfError(f"{xdmf_path}: no grid with Topology and Geometry")
    return geo[0]


def _entity_keys(ent, nv):
    """One uint64 per entity, equal iff the vertex sets are equal (exact packing when it fits, verified hash otherwise)."""
    s = np.sort(np.asarray(ent, np.int64), axis=1).astype(np.uint64)
    k = s.shape[1]
    bits = max(int(nv - 1).bit_length(), 1)
    if bits * k <= 64:
        key = np.zeros(s.shape[0], np.uint64)
        for j in range(k):
            key = (key << np.uint64(bits)) | s[:, j]
        return key, s, True
    mult = (np.uint64(0x9E3779B97F4A7C15), np.uint64(0xC2B2AE3D27D4EB4F), np.uint64(0x165667B19E3779F9), np.uint64(0xD6E8FEB86659FD93))
    key = np.zeros(s.shape[0], np.uint64)
    with np.errstate(over="ignore"):
        for j in range(k):
            key = (key ^ (s[:, j] * mult[j])) * np.uint64(0xFF51AFD7ED558CCD)
    return key, s, False


def match_entities(ent, values, target, nv, default):
    """Values of the tagged entities `ent` transferred to the entities `target` (both given by vertices): default where a
    target entity is not tagged."""
    out = np.full(target.shape[0], default, np.int32)
    if ent.shape[0] == 0 or target.shape[0] == 0:
        return out
    if ent.shape[1] != target.shape[1]:
        raise XdmfError(f"tag entities have {ent.shape[1]} vertices, expected {target.shape[1]}")
    ke, se, exact = _entity_keys(ent, nv)
    kt, st, _ = _entity_keys(target, nv)
    order = np.argsort(ke, kind="stable")
    pos = np.searchsorted(ke[order], kt)
    pos[pos >= order.size] = order.size - 1
    cand = order[pos]
    hit = ke[cand] == kt
    if not exact:
        hit &= np.all(se[cand] == st, axis=1)
    out[hit] = np.asarray(values)[cand[hit]]
    return out


def read_xdmf_mesh(mesh_file, facet_file, ct_name="ct", ft_name="ft"):
    """-> dict(gdim, x (Nv, gdim) float64 unscaled, cells (Nc, gdim+1) int32, cell_tags (Nc,) int32 [0 = untagged],
    facets (Nf, gdim) int32, facet_tags (Nf,) int32): the mesh, its cell tags and every tagged facet of the facet file."""
    h5cache = {}
    try:
        grids = _grids(mesh_file)
        gmesh = _find_grid(grids, "mesh", mesh_file, need_values=False)
        ttype, cells = _topology(gmesh, mesh_file, h5cache)
        if ttype not in ("triangle", "tetrahedron"):
            raise XdmfError(f"{mesh_file}: the mesh grid holds {ttype} cells; triangles or tetrahedra expected")
        gdim = 2 if ttype == "triangle" else 3
        geo = _children(gmesh, "Geometry")[0]
        x = np.asarray(_read_item(_children(geo, "DataItem")[0], mesh_file, h5cache), np.float64)
        x = x.reshape(-1, x.shape[-1] if x.ndim == 2 else (3 if (geo.get("GeometryType") or "XYZ").upper() == "XYZ" else 2))
        if x.shape[1] > gdim:
            if np.any(x[:, gdim:] != 0.0):
                raise XdmfError(f"{mesh_file}: triangle mesh embedded in 3D (non-zero z): not supported")
            x = x[:, :gdim]
        nv = x.shape[0]
        if cells.size and (cells.min() < 0 or cells.max() >= nv):
            raise XdmfError(f"{mesh_file}: topology refers to nodes outside the geometry")
        # cell tags
        gct = _find_grid(grids, ct_name, mesh_file, need_values=True)
        vals = _values(gct, mesh_file, h5cache)
        _tt, ent = _topology(gct, mesh_file, h5cache)
        if ent is None or (ent.shape == cells.shape and np.array_equal(ent, cells)):
            if vals.size != cells.shape[0]:
                raise XdmfError(f"{mesh_file}: {vals.size} cell values for {cells.shape[0]} cells")
            cell_tags = vals.astype(np.int32)
        else:
            cell_tags = match_entities(ent, vals, cells, nv, 0)
        # facet tags
        fgrids = grids if os.path.abspath(facet_file) == os.path.abspath(mesh_file) else _grids(facet_file)
        gft = _find_grid(fgrids, ft_name, facet_file, need_values=True)
        ftt, fent = _topology(gft, facet_file, h5cache)
        if fent is None or fent.shape[1] != gdim:
            raise XdmfError(f"{facet_file}: grid {ft_name!r} does not hold facets of a {ttype} mesh (got {ftt})")
        fvals = _values(gft, facet_file, h5cache)
        if fvals.size != fent.shape[0]:
            raise XdmfError(f"{facet_file}: {fvals.size} facet values for {fent.shape[0]} facets")
        return dict(gdim=gdim, x=np.ascontiguousarray(x), cells=np.ascontiguousarray(cells, dtype=np.int32),
                    cell_tags=cell_tags, facets=np.ascontiguousarray(fent, dtype=np.int32), facet_tags=fvals.astype(np.int32))
    finally:
        for f in h5cache.values():
            f.close()


# ------------------------------------------------------------------------------------------------ writer
def _item(arr, h5name, dpath, fmt):
    arr = np.asarray(arr)
    ntype = "Float" if arr.dtype.kind == "f" else "Int"
    dims = " ".join(str(s) for s in arr.shape)
    if fmt == "HDF":
        return f'<DataItem Dimensions="{dims}" NumberType="{ntype}" Precision="{arr.dtype.itemsize}" Format="HDF">{h5name}:{dpath}</DataItem>'
    body = "\n".join(" ".join(repr(v) if ntype == "Float" else str(v) for v in row) for row in arr.reshape(arr.shape[0], -1).tolist())
    return f'<DataItem Dimensions="{dims}" NumberType="{ntype}" Precision="{arr.dtype.itemsize}" Format="XML">{body}</DataItem>'


def write_xdmf_mesh(path, x, cells, tags=None, fmt="HDF"):
    """DOLFINx layout: grid "mesh" plus one grid per entry of tags = {name: (entities, values)} (``write_mesh`` +
    ``write_meshtags``, utils/generate_square_mesh.py:37-42)."""
    x, cells = np.asarray(x, np.float64), np.asarray(cells, np.int64)
    gdim = cells.shape[1] - 1
    tname = {2: "PolyLine", 3: "Triangle", 4: "Tetrahedron"}
    h5name = os.path.splitext(os.path.basename(path))[0] + ".h5"
    data = {"/Mesh/mesh/topology": cells, "/Mesh/mesh/geometry": x}
    out = ['<?xml version="1.0"?>', '<Xdmf Version="3.0" xmlns:xi="http://www.w3.org/2001/XInclude">', "<Domain>",
           '<Grid Name="mesh" GridType="Uniform">',
           f'<Topology TopologyType="{tname[gdim + 1]}" NumberOfElements="{cells.shape[0]}" NodesPerElement="{gdim + 1}">',
           _item(cells, h5name, "/Mesh/mesh/topology", fmt), "</Topology>",
           f'<Geometry GeometryType="{"XY" if x.shape[1] == 2 else "XYZ"}">', _item(x, h5name, "/Mesh/mesh/geometry", fmt),
           "</Geometry>", "</Grid>"]
    for name, (ent, vals) in (tags or {}).items():
        ent, vals = np.asarray(ent, np.int64), np.asarray(vals, np.int32).reshape(-1, 1)
        data[f"/MeshTags/{name}/topology"] = ent
        data[f"/MeshTags/{name}/Values"] = vals
        out += [f'<Grid Name="{name}" GridType="Uniform">',
                '<xi:include xpointer="xpointer(/Xdmf/Domain/Grid/Geometry)" />',
                f'<Topology TopologyType="{tname[ent.shape[1]]}" NumberOfElements="{ent.shape[0]}" NodesPerElement="{ent.shape[1]}">',
                _item(ent, h5name, f"/MeshTags/{name}/topology", fmt), "</Topology>",
                f'<Attribute Name="{name}" AttributeType="Scalar" Center="Cell">',
                _item(vals, h5name, f"/MeshTags/{name}/Values", fmt), "</Attribute>", "</Grid>"]
    out += ["</Domain>", "</Xdmf>"]
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    if fmt == "HDF":
        hdf5_min.write_file(os.path.join(os.path.dirname(os.path.abspath(path)), h5name), data)
