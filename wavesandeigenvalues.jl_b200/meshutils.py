"""Host mirror of the parts of the reference's ``Meshutils`` the hot path consumes.

Same names and numbering rules as the reference (0-based here, 1-based there), but
sort-based and vectorised (the reference's ``insert_smplx!`` lists are O(n^2),
src/Mesh/sorter.jl:141-150):

  Mesh(file; scale)                 src/Meshutils.jl:92-165, read_msh4 :272-402
  simplex ordering                  src/Mesh/sorter.jl:9-31  (ascending in the descending-sorted vertex tuple;
                                    first occurrence of a vertex set wins, vertex order inside a simplex = file order)
  collect_lines!                    src/Meshutils.jl:831-840
  aggregate_elements                src/FEM/FEM.jl:84-116
  link_triangles_to_tetrahedra!     src/Meshutils.jl:516-548
  compute_size!                     src/Meshutils.jl:757-780
  find_tetrahedron_containing_point src/Meshutils.jl:800-815
  generate_field                    src/Meshutils.jl:1079-1098

plus synthetic structured (Kuhn 6-tet) generators for the large benchmark configurations.
"""
import numpy as np


_NATIVE_MIN = 1 << 15  # below this the numpy path is as fast as the call into the library


def _sorted_unique_numpy(simp):
    key = -np.sort(-simp, axis=1)  # descending-sorted vertex tuple
    order = np.lexsort([np.arange(len(simp))] + [key[:, c] for c in range(key.shape[1] - 1, -1, -1)])
    ks = key[order]
    first = np.ones(len(simp), dtype=bool)
    first[1:] = np.any(ks[1:] != ks[:-1], axis=1)
    grp = np.cumsum(first) - 1
    inv = np.empty(len(simp), dtype=np.int64)
    inv[order] = grp
    return simp[order[first]], inv


def _sorted_unique(simp, native=None):
    """simp: (n,k) int array in file order -> (unique sorted simplices, index map raw->sorted).  Large inputs (the 6 n_tet edges of
    collect_lines, the simplex lists of big meshes) go through the library's thread-parallel sort (csrc/mesh_symbolic.cpp); both
    paths give the same numbering (tests/test_host_logic.py)."""
    simp = np.asarray(simp, dtype=np.int64)
    if len(simp) == 0:
        return simp, np.zeros(0, dtype=np.int64)
    simp = simp.reshape(len(simp), -1)
    if native is None:
        native = len(simp) >= _NATIVE_MIN
    if native and 2 <= simp.shape[1] <= 4 and simp.min() >= 0 and simp.max() < 2**32:
        from ._lib import sorted_unique_simplices
        first, inv = sorted_unique_simplices(simp)
        return simp[first], inv
    return _sorted_unique_numpy(simp)


def read_msh4(fname):
    """gmsh 4.1 ASCII reader (PhysicalNames, Entities, Nodes, Elements of type 1/2/4)."""
    with open(fname) as f:
        lines = f.read().split("\n")
    pos = 0
    tag2dom, ent2dom, domains = {}, [dict(), dict(), dict(), dict()], {}
    raw = {1: [], 2: [], 4: []}
    points = None
    n = len(lines)
    while pos < n:
        fld = lines[pos].strip()
        pos += 1
        if fld == "$PhysicalNames":
            cnt = int(lines[pos]); pos += 1
            for _ in range(cnt):
                dim, tag, dom = lines[pos].split(); pos += 1
                dom = dom[1:-1]
                tag2dom[tag] = dom
                domains[dom] = {"dimension": int(dim), "simplices": []}
        elif fld == "$Entities":
            counts = [int(x) for x in lines[pos].split()]; pos += 1
            for d, c in enumerate(counts):
                off = 4 if d == 0 else 7
                for _ in range(c):
                    sp = lines[pos].split(); pos += 1
                    nph = int(sp[off])
                    ent2dom[d][sp[0]] = [tag2dom[t] for t in sp[off + 1: off + 1 + nph]]
        elif fld == "$Nodes":
            nblk, nnodes = (int(x) for x in lines[pos].split()[:2]); pos += 1
            points = np.empty((3, nnodes))
            for _ in range(nblk):
                nin = int(lines[pos].split()[3]); pos += 1
                tags = np.array(lines[pos: pos + nin], dtype=np.int64) - 1; pos += nin
                xyz = np.array([l.split()[:3] for l in lines[pos: pos + nin]], dtype=float).reshape(nin, 3); pos += nin
                points[:, tags] = xyz.T
        elif fld == "$Elements":
            nblk = int(lines[pos].split()[0]); pos += 1
            for _ in range(nblk):
                sp = lines[pos].split(); pos += 1
                edim, etag, etype, nin = int(sp[0]), sp[1], int(sp[2]), int(sp[3])
                if etype in raw:
                    start = len(raw[etype])
                    raw[etype].extend([int(x) - 1 for x in l.split()[1:]] for l in lines[pos: pos + nin])
                    for dom in ent2dom[edim].get(etag, []):
                        domains[dom]["simplices"].extend(range(start, start + nin))
                pos += nin
    return points, raw[1], raw[2], raw[4], domains


class Mesh:
    """Tetrahedral mesh with the reference's fields: points (3xN), lines, triangles, tetrahedra (sorted unique
    simplex arrays), domains[name] = {"dimension", "simplices"[, "size"]}, tri2tet."""

    def __init__(self, file_name=None, scale=1.0, raw=None):
        if raw is None:
            raw = read_msh4(file_name)
        points, lines, tris, tets, domains = raw
        self.name = self.file = file_name
        self.points = np.asarray(points, dtype=float) * scale
        self.lines, lmap = _sorted_unique(np.asarray(lines, dtype=np.int64).reshape(-1, 2))
        self.triangles, tmap = _sorted_unique(np.asarray(tris, dtype=np.int64).reshape(-1, 3))
        self.tetrahedra, ttmap = _sorted_unique(np.asarray(tets, dtype=np.int64).reshape(-1, 4))
        self.domains = {}
        for dom, d in domains.items():
            mp = {1: lmap, 2: tmap, 3: ttmap}[d["dimension"]]
            idx = mp[np.asarray(d["simplices"], dtype=np.int64)] if len(d["simplices"]) else np.zeros(0, dtype=np.int64)
            _, first = np.unique(idx, return_index=True)  # unique!, first-occurrence order
            self.domains[dom] = {"dimension": d["dimension"], "simplices": idx[np.sort(first)]}
        self.tri2tet = None
        self.dos = 1
        self._edge_of = None

    def __repr__(self):
        return (f"mesh: {self.name}\n#################\npoints:     {self.points.shape[1]}\nlines:      {len(self.lines)}\n"
                f"triangles:  {len(self.triangles)}\ntetrahedra: {len(self.tetrahedra)}\n#################\ndomains: "
                + ", ".join(sorted(self.domains)))

    # -- Meshutils.jl:831-840 --------------------------------------------------------------------
    def collect_lines(self):
        """Populate mesh.lines; also caches tet_edges (n_tet,6) = index of each tet's edges 12,13,14,23,24,34."""
        if self._edge_of is None:
            t = self.tetrahedra
            pairs = np.stack([t[:, [0, 1]], t[:, [0, 2]], t[:, [0, 3]], t[:, [1, 2]], t[:, [1, 3]], t[:, [2, 3]]], axis=1)
            self.lines, inv = _sorted_unique(pairs.reshape(-1, 2))
            self._edge_of = inv.reshape(-1, 6)
        return self.lines

    def edge_index(self, a, b):
        """Index in mesh.lines of the edges (a[i], b[i]) (vectorised find_smplx / get_line_idx, annular_meshes.jl:603)."""
        self.collect_lines()
        npts = self.points.shape[1]
        hi, lo = np.maximum(a, b), np.minimum(a, b)
        lk = -np.sort(-self.lines, axis=1)
        keys = lk[:, 0] * npts + lk[:, 1]
        order = np.argsort(keys, kind="stable")  # identity for the default (sorted) line list
        return order[np.searchsorted(keys[order], hi * npts + lo)]

    def reorder_lines(self, new_of_old):
        """Renumber mesh.lines (unit-cell meshes keep their edges sector-wise as [axis | Bloch plane | rest | image plane],
        annular_meshes.jl:459-496); new_of_old[i] = new index of the line currently at index i."""
        self.collect_lines()
        inv = np.empty_like(new_of_old)
        inv[new_of_old] = np.arange(len(new_of_old))
        self.lines = self.lines[inv]
        self._edge_of = np.asarray(new_of_old)[self._edge_of]

    # -- Meshutils.jl:516-548 --------------------------------------------------------------------
    def link_triangles_to_tetrahedra(self):
        npts = self.points.shape[1]
        t = self.tetrahedra
        faces = np.stack([t[:, [0, 1, 2]], t[:, [0, 1, 3]], t[:, [0, 2, 3]], t[:, [1, 2, 3]]], axis=1).reshape(-1, 3)
        fk = -np.sort(-faces, axis=1)
        fkey = (fk[:, 0] * npts + fk[:, 1]) * npts + fk[:, 2]
        tk = -np.sort(-self.triangles, axis=1)
        tkey = (tk[:, 0] * npts + tk[:, 1]) * npts + tk[:, 2]
        pos = np.searchsorted(tkey, fkey)
        pos[pos >= len(tkey)] = 0
        hit = tkey[pos] == fkey
        t2t = np.full(len(self.triangles), -1, dtype=np.int64)
        tet_of_face = np.repeat(np.arange(len(t)), 4)
        t2t[pos[hit]] = tet_of_face[hit]  # ascending tet order: the last (highest) tet wins, as in the reference loop
        self.tri2tet = t2t
        return t2t

    # -- Meshutils.jl:757-780 --------------------------------------------------------------------
    def compute_size(self, dom):
        d = self.domains[dom]
        P = self.points
        if d["dimension"] == 3:
            X = P[:, self.tetrahedra[d["simplices"]]]  # 3 x n x 4
            J = X[:, :, :3] - X[:, :, 3:4]
            V = np.abs(np.linalg.det(np.moveaxis(J, 1, 0))).sum() / 6
        elif d["dimension"] == 2:
            X = P[:, self.triangles[d["simplices"]]]
            V = np.linalg.norm(np.cross(X[:, :, 0] - X[:, :, 2], X[:, :, 1] - X[:, :, 2], axis=0), axis=0).sum() / 2
        else:
            X = P[:, self.lines[d["simplices"]]]
            V = np.linalg.norm(X[:, :, 1] - X[:, :, 0], axis=0).sum()
        d["size"] = float(V)
        return d["size"]

    # -- Meshutils.jl:800-815 --------------------------------------------------------------------
    def find_tetrahedron_containing_point(self, point):
        """Lowest-index tetrahedron with all barycentric coordinates in [0,1]; -1 if none (reference: 0)."""
        p = np.asarray(point, dtype=float)
        X = self.points[:, self.tetrahedra]  # 3 x n x 4
        J = np.moveaxis(X[:, :, :3] - X[:, :, 3:4], 1, 0)  # n x 3 x 3
        rhs = (p[:, None] - X[:, :, 3]).T  # n x 3
        xi = np.linalg.solve(J, rhs[:, :, None])[:, :, 0]
        xi = np.concatenate([xi, 1 - xi.sum(axis=1, keepdims=True)], axis=1)
        ok = np.all((xi >= 0) & (xi <= 1), axis=1)
        idx = np.flatnonzero(ok)
        return int(idx[0]) if len(idx) else -1

    # -- Meshutils.jl:1079-1098 ------------------------------------------------------------------
    def generate_field(self, func, order="const"):
        if order == "const":
            cen = self.points[:, self.tetrahedra].sum(axis=2) / 4
            return np.array([func(*cen[:, i]) for i in range(cen.shape[1])], dtype=float)
        if order == "lin":
            return np.array([func(*self.points[:, i]) for i in range(self.points.shape[1])], dtype=float)
        raise ValueError(f"order {order} not supported")


def aggregate_elements(mesh, order="lin"):
    """FEM.jl:84-116 for :lin / :quad -> (triangles (n,3|6), tetrahedra (n,4|10), dim)."""
    npts = mesh.points.shape[1]
    if order == "lin":
        return mesh.triangles, mesh.tetrahedra, npts
    if order != "quad":
        raise NotImplementedError("element order %r: only :lin and :quad are on the accelerated path" % (order,))
    mesh.collect_lines()
    tets = np.concatenate([mesh.tetrahedra, mesh._edge_of + npts], axis=1)
    t = mesh.triangles
    if len(t):
        e = np.stack([mesh.edge_index(t[:, 0], t[:, 1]), mesh.edge_index(t[:, 0], t[:, 2]), mesh.edge_index(t[:, 1], t[:, 2])], axis=1)
        tris = np.concatenate([t, e + npts], axis=1)
    else:
        tris = np.zeros((0, 6), dtype=np.int64)
    return tris, tets, npts + len(mesh.lines)


class SymInfo:
    """Symmetry bookkeeping of a unit-cell mesh (the fields of Meshutils.jl:28-44 that discretize/blochify read)."""

    def __init__(self, DOS, naxis, nxbloch, nxsector, naxis_ln, nxbloch_ln, nxsector_ln, unit=True):
        self.DOS, self.naxis, self.nxbloch, self.nxsector = DOS, naxis, nxbloch, nxsector
        self.naxis_ln, self.nxbloch_ln, self.nxsector_ln, self.unit = naxis_ln, nxbloch_ln, nxsector_ln, unit


def bloch_dof_maps(mesh, order):
    """DOF folding of blochify (Bloch.jl:4-66) as arrays over the unfolded DOFs:
    (new_index, is_image, is_axis, reduced_dim).  Point DOFs > nsector and line DOFs > nsector_ln are images of the
    Bloch reference plane; line DOFs move down by nxbloch because the image points disappear."""
    d = mesh.dos
    npts = mesh.points.shape[1]
    nsector = d.naxis + d.nxsector
    dim = npts + (len(mesh.collect_lines()) if order == "quad" else 0)
    idx = np.arange(dim)
    new = idx.copy()
    image = np.zeros(dim, dtype=bool)
    image[:npts] = idx[:npts] >= nsector
    new[:npts] -= np.where(image[:npts], nsector - d.naxis, 0)
    axis = np.zeros(dim, dtype=bool)
    axis[:npts] = new[:npts] < d.naxis
    if order == "quad":
        ln = idx[npts:] - npts
        img_ln = ln >= d.naxis_ln + d.nxsector_ln
        image[npts:] = img_ln
        ln_new = ln - np.where(img_ln, d.nxsector_ln, 0)
        axis[npts:] = ln_new < d.naxis_ln
        new[npts:] = npts + ln_new - d.nxbloch
    red = dim - d.nxbloch - (d.nxbloch_ln if order == "quad" else 0)
    return new, image, axis, red


# ---------------------------------------------------------------------------------------------
# synthetic structured meshes: Kuhn 6-tet boxes with seeded jitter
# ---------------------------------------------------------------------------------------------
_KUHN = np.array([[0, 1, 3, 7], [0, 1, 5, 7], [0, 2, 3, 7], [0, 2, 6, 7], [0, 4, 5, 7], [0, 4, 6, 7]])


def kuhn_box(ncube, lo, hi, jitter=0.0, seed=0, flame_layer=None, name="kuhn_box"):
    """Box [lo,hi] cut into ncube=(nx,ny,nz) cubes of 6 Kuhn tetrahedra each.

    Domains: "Interior" (all tets), "Outlet" (z = hi face), "Inlet" (z = lo face), "Walls" (the rest of the
    boundary) and, if flame_layer=(k0,k1), "Flame" = cube layers k0 <= k < k1 in z.
    Interior vertices are moved by U(-jitter,jitter)*h per axis with numpy.random.default_rng(seed).
    """
    nx, ny, nz = ncube
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    gx, gy, gz = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    pid = (gz * (ny + 1) + gy) * (nx + 1) + gx  # x fastest
    h = (hi - lo) / np.array([nx, ny, nz])
    pts = np.empty((3, (nx + 1) * (ny + 1) * (nz + 1)))
    for r, g in enumerate((gx, gy, gz)):
        pts[r, pid.ravel()] = lo[r] + g.ravel() * h[r]
    if jitter:
        rng = np.random.default_rng(seed)
        interior = ((gx > 0) & (gx < nx) & (gy > 0) & (gy < ny) & (gz > 0) & (gz < nz)).ravel()
        d = rng.uniform(-jitter, jitter, size=(3, pts.shape[1])) * h[:, None]
        pts[:, pid.ravel()[interior]] += d[:, pid.ravel()[interior]]
    # cubes
    cx, cy, cz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cx, cy, cz = cx.ravel(), cy.ravel(), cz.ravel()
    corner = np.stack([pid[cx + (k & 1), cy + ((k >> 1) & 1), cz + ((k >> 2) & 1)] for k in range(8)], axis=1)
    tets = corner[:, _KUHN].reshape(-1, 4)
    tet_cz = np.repeat(cz, 6)
    # boundary triangles: the faces whose three vertices lie on one of the six planes of the box (each occurs once); this
    # needs no global sort of the 4 n_tet faces, so it scales to the 50 M-tetrahedra stress configuration
    on_plane = []
    vx, vy, vz = (np.arange(pts.shape[1]) % (nx + 1)), (np.arange(pts.shape[1]) // (nx + 1)) % (ny + 1), np.arange(pts.shape[1]) // ((nx + 1) * (ny + 1))
    for g, nmax in ((vx, nx), (vy, ny), (vz, nz)):
        on_plane.append((g == 0).astype(np.int8))
        on_plane.append((g == nmax).astype(np.int8))
    tris = []
    for f in ([0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]):
        face = tets[:, f]
        hit = np.zeros(len(tets), dtype=bool)
        for m in on_plane:
            hit |= (m[face[:, 0]] & m[face[:, 1]] & m[face[:, 2]]).astype(bool)
        tris.append(face[hit])
    tris = np.concatenate(tris)
    zc = pts[2, tris]
    tol = 1e-9 * (hi[2] - lo[2])
    on_out = np.all(np.abs(zc - hi[2]) < tol, axis=1)
    on_in = np.all(np.abs(zc - lo[2]) < tol, axis=1)
    domains = {
        "Interior": {"dimension": 3, "simplices": np.arange(len(tets))},
        "Outlet": {"dimension": 2, "simplices": np.flatnonzero(on_out)},
        "Inlet": {"dimension": 2, "simplices": np.flatnonzero(on_in)},
        "Walls": {"dimension": 2, "simplices": np.flatnonzero(~on_out & ~on_in)},
    }
    if flame_layer is not None:
        k0, k1 = flame_layer
        domains["Flame"] = {"dimension": 3, "simplices": np.flatnonzero((tet_cz >= k0) & (tet_cz < k1))}
    return Mesh(name, raw=(pts, np.zeros((0, 2), dtype=np.int64), tris, tets, domains))


def kuhn_unit_cell(ncube, lo, hi, DOS, jitter=0.0, seed=0, name="kuhn_unit_cell"):
    """One period (in x) of a DOS-periodic duct as a unit-cell mesh with the reference's index layout
    (annular_meshes.jl:292-368, naxis = 0): points [Bloch reference plane x=lo | body | image plane x=hi], image point
    k = reference point k + nxsector; edges [Bloch plane | rest | image plane] with the same pairing.  The two periodic
    planes carry no boundary triangles.  Domains: "Interior", "Outlet" (z=hi), "Inlet" (z=lo), "Walls" (y faces)."""
    nx, ny, nz = ncube
    base = kuhn_box(ncube, lo, hi, jitter=0.0, seed=seed)
    pts = base.points.copy()
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    tol = 1e-9 * (hi[0] - lo[0])
    on_ref = np.abs(pts[0] - lo[0]) < tol
    on_img = np.abs(pts[0] - hi[0]) < tol
    if jitter:  # periodic jitter: interior points only (the planes stay congruent)
        rng = np.random.default_rng(seed)
        h = (hi - lo) / np.array(ncube)
        inner = np.ones(pts.shape[1], dtype=bool)
        for r in range(3):
            inner &= (pts[r] > lo[r] + tol) & (pts[r] < hi[r] - tol)
        pts[:, inner] += rng.uniform(-jitter, jitter, size=(3, int(inner.sum()))) * h[:, None]
    # new point numbering: reference plane, body, image plane (partner order = (y,z) lexicographic on both planes)
    key = np.round((pts[1] - lo[1]) / (hi[1] - lo[1]) * ny).astype(np.int64) * (nz + 1) + np.round((pts[2] - lo[2]) / (hi[2] - lo[2]) * nz).astype(np.int64)
    ref = np.flatnonzero(on_ref)
    ref = ref[np.argsort(key[ref], kind="stable")]
    img = np.flatnonzero(on_img)
    img = img[np.argsort(key[img], kind="stable")]
    body = np.flatnonzero(~on_ref & ~on_img)
    order = np.concatenate([ref, body, img])
    new_of_old = np.empty(len(order), dtype=np.int64)
    new_of_old[order] = np.arange(len(order))
    tets = new_of_old[base.tetrahedra]
    tris = new_of_old[base.triangles]
    pts = pts[:, order]
    nxb, npt = len(ref), pts.shape[1]
    nxsector = npt - nxb
    keep = ~(np.all(tris < nxb, axis=1) | np.all(tris >= nxsector, axis=1))  # drop the periodic planes
    tris = tris[keep]
    zc, yc = pts[2, tris], pts[1, tris]
    tz, ty = 1e-9 * (hi[2] - lo[2]), 1e-9 * (hi[1] - lo[1])
    out = np.all(np.abs(zc - hi[2]) < tz, axis=1)
    inn = np.all(np.abs(zc - lo[2]) < tz, axis=1)
    domains = {"Interior": {"dimension": 3, "simplices": np.arange(len(tets))},
               "Outlet": {"dimension": 2, "simplices": np.flatnonzero(out)},
               "Inlet": {"dimension": 2, "simplices": np.flatnonzero(inn)},
               "Walls": {"dimension": 2, "simplices": np.flatnonzero(~out & ~inn)}}
    mesh = Mesh(name, raw=(pts, np.zeros((0, 2), dtype=np.int64), tris, tets, domains))
    # edges: [Bloch plane | rest | image plane], image edge k <-> Bloch edge k
    lines = mesh.collect_lines()
    on_b = np.all(lines < nxb, axis=1)
    on_i = np.all(lines >= nxsector, axis=1)
    lb = np.flatnonzero(on_b)
    li = np.flatnonzero(on_i)
    kb = np.sort(lines[lb], axis=1)
    ki = np.sort(lines[li] - nxsector, axis=1)
    lb = lb[np.lexsort((kb[:, 1], kb[:, 0]))]
    li = li[np.lexsort((ki[:, 1], ki[:, 0]))]
    assert len(lb) == len(li) and np.array_equal(np.sort(lines[lb], axis=1), np.sort(lines[li] - nxsector, axis=1))
    rest = np.flatnonzero(~on_b & ~on_i)
    lorder = np.concatenate([lb, rest, li])
    new_of_old_ln = np.empty(len(lorder), dtype=np.int64)
    new_of_old_ln[lorder] = np.arange(len(lorder))
    mesh.reorder_lines(new_of_old_ln)
    mesh.dos = SymInfo(DOS, 0, nxb, nxsector, 0, len(lb), len(lines) - len(lb))
    return mesh


# ---------------------------------------------------------------------------------------------
# extend_mesh: half-cell -> unit cell / full annulus (src/Mesh/annular_meshes.jl:269-546), vectorised
# ---------------------------------------------------------------------------------------------
def _plane(A):
    """three_points_to_plane (annular_meshes.jl:14-26), including its d < 1e-7 -> 0 rule."""
    n = np.cross(A[:, 0] - A[:, 2], A[:, 1] - A[:, 2])
    n = n / np.linalg.norm(n)
    d = -float(n @ A[:, 2])
    return np.array([n[0], n[1], n[2], 0.0 if d < 1e-7 else d])


def _reflect(P, pln):
    """reflect_point_at_plane (:46-52) for the columns of P."""
    k = -(pln[:3] @ P) - pln[3]
    return P + 2 * np.outer(pln[:3], k)


def _find_rows(table_sorted_keys, order, keys):
    pos = np.searchsorted(table_sorted_keys, keys)
    assert np.all(table_sorted_keys[np.minimum(pos, len(table_sorted_keys) - 1)] == keys), "simplex not found"
    return order[pos]


def _desc_key(simp, npts):
    k = -np.sort(-np.asarray(simp, dtype=np.int64), axis=1)
    key = k[:, 0]
    for c in range(1, k.shape[1]):
        key = key * npts + k[:, c]
    return key


def extend_mesh(mesh, doms, sym_name="Symmetry", blch_name="Bloch", unit=False):
    """Mirror a half-cell mesh at its symmetry plane (unit=True: one unit cell with the Bloch image plane) and rotate the unit cell
    DOS times around the axis (unit=False: full annulus).  Same point, simplex and domain numbering as the reference
    (annular_meshes.jl:269-546): points [axis | Bloch plane | body | symmetry plane | mirrored body | (image plane)], simplices sorted
    and uniquified by the reference's rule, lines of the first cell sorted and the image / rotated lines appended.  `doms` is a list
    of (domain, "full" | "unit" | "half").  naxis_ln counts the leading lines on the axis (the reference's counter returns 1 instead
    of 0 for meshes without axis points, which only matters for second-order unit cells and breaks them there)."""
    mesh.collect_lines()
    npts = mesh.points.shape[1]
    tri, tet = mesh.triangles, mesh.tetrahedra
    on_b = np.zeros(npts, dtype=bool)
    on_s = np.zeros(npts, dtype=bool)
    on_b[tri[np.asarray(mesh.domains[blch_name]["simplices"], dtype=np.int64)].ravel()] = True
    on_s[tri[np.asarray(mesh.domains[sym_name]["simplices"], dtype=np.int64)].ravel()] = True
    axis = np.flatnonzero(on_b & on_s)
    new_order = np.concatenate([axis, np.flatnonzero(on_b & ~on_s), np.flatnonzero(~on_b & ~on_s), np.flatnonzero(on_s & ~on_b)])
    trace = np.empty(npts, dtype=np.int64)
    trace[new_order] = np.arange(npts)
    t2t = mesh.link_triangles_to_tetrahedra() if mesh.tri2tet is None else mesh.tri2tet

    def plane_of(dom):
        si = int(mesh.domains[dom]["simplices"][0])
        t = tri[si]
        pl = _plane(mesh.points[:, t])
        test = [p for p in tet[t2t[si]] if p not in t][-1]  # find_testpoint_idx (:124-132)
        x = mesh.points[:, test]
        foot = x + pl[:3] * (-(pl[:3] @ x) - pl[3])
        return pl * (-np.sign(pl[:3] @ (x - foot)))  # make_normal_outwards (:74-79)

    pln, bpln = plane_of(sym_name), plane_of(blch_name)
    naxis, nbloch, nsym = len(axis), int(on_b.sum()), int(on_s.sum())
    nxsym, nxbloch = nsym - naxis, nbloch - naxis
    nbody = npts - nbloch - nxsym
    shiftbody = npts - nbloch
    nxsector = nxbloch + 2 * nbody + nxsym
    nsector = nxsector + naxis
    pts = np.zeros((3, 2 * npts - nsym))
    pts[:, :npts] = mesh.points[:, new_order]
    pts[:, nbloch + shiftbody: npts - nxsym + shiftbody] = _reflect(pts[:, nbloch: npts - nxsym], pln)
    pts[:, naxis + nxsector: nbloch + nxsector] = _reflect(pts[:, naxis:nbloch], pln)
    DOS = int(round(np.pi / np.arccos(pln[:3] @ -bpln[:3])))
    n = np.cross(pln[:3], bpln[:3])
    n = n / np.linalg.norm(n)
    # find_intersection_of_two_planes (:92-109): p = rhs \ M, least squares with the 5-vector rhs on the left
    f1, f2 = pln[:3] * (-pln[3]), bpln[:3] * (-bpln[3])
    M = np.array([[2.0, 0, 0, pln[0], bpln[0]], [2.0, 0, 0, pln[1], bpln[1]], [2.0, 0, 0, pln[2], bpln[2]],
                  [pln[0], pln[1], pln[2], 0, 0], [bpln[0], bpln[1], bpln[2], 0, 0]])
    rhs = np.array([0, 0, 0, f1 @ pln[:3], f2 @ bpln[:3]], dtype=float)
    p0 = np.linalg.lstsq(rhs.reshape(5, 1), M, rcond=None)[0].ravel()[:3]
    phi = 2 * np.pi / DOS
    if unit:
        fpts, dos_lim = pts, 1
    else:
        dos_lim = DOS
        fpts = np.zeros((3, naxis + nxsector * DOS))
        fpts[:, :nsector] = pts[:, :nsector]
        for s in range(1, DOS):
            c, sn = np.cos(s * phi), np.sin(s * phi)
            K = np.array([[0, -n[2], n[1]], [n[2], 0, -n[0]], [-n[1], n[0], 0]])
            R = c * np.eye(3) + sn * K + (1 - c) * np.outer(n, n)  # create_rotation_matrix_around_axis (:116-121)
            fpts[:, naxis + nxsector * s: naxis + nxsector * (s + 1)] = R @ (pts[:, naxis:nsector] - p0[:, None]) + p0[:, None]

    def refl(i):  # get_reflected_index (:168-181)
        i = np.asarray(i, dtype=np.int64)
        assert np.all(i < naxis + nxbloch + nbody + nxsym)
        return np.where(i < naxis, i, np.where(i < nbloch, i + nxsector, np.where(i < nbloch + nbody, i + shiftbody, i)))

    def rot(i, s):  # get_rotated_index (:142-154)
        i = np.asarray(i, dtype=np.int64)
        return np.where(i < naxis, i, (i + nxsector * s - naxis) % (nxsector * DOS) + naxis)

    def build(simp, keep):
        t = trace[simp[keep]]
        r = refl(t)
        parts = []
        for s in range(dos_lim):  # insertion order of the reference: per simplex (direct, mirrored) per sector
            parts.append(np.stack([rot(t, s), rot(r, s)], axis=1))
        allp = np.stack(parts, axis=1).reshape(-1, simp.shape[1])
        return _sorted_unique(allp)[0]

    nfp = fpts.shape[1]
    tets = build(tet, np.ones(len(tet), dtype=bool))
    keep = np.ones(len(tri), dtype=bool)
    keep[np.asarray(mesh.domains[sym_name]["simplices"], dtype=np.int64)] = False
    if not unit:
        keep[np.asarray(mesh.domains[blch_name]["simplices"], dtype=np.int64)] = False
    tris = build(tri, keep)
    ln = trace[mesh.lines]
    not_bloch = ~np.all(ln < nbloch, axis=1)
    lines = _sorted_unique(np.concatenate([ln, refl(ln[not_bloch])]))[0]
    naxis_ln = int(np.all(lines < naxis, axis=1).sum())
    nbloch_ln = int(np.all(lines < nbloch, axis=1).sum())
    nsector_ln = len(lines)
    if unit:
        lines = np.concatenate([lines, rot(lines[naxis_ln:nbloch_ln], 1)])
    else:
        first = lines[naxis_ln:nsector_ln]
        lines = np.concatenate([lines] + [rot(first, s) for s in range(1, DOS)])
    domains = {}
    for dom, deg in doms:
        dim = mesh.domains[dom]["dimension"]
        src, full = (tet, tets) if dim == 3 else (tri, tris)
        fkey = _desc_key(full, nfp)
        fo = np.argsort(fkey, kind="stable")
        fk = fkey[fo]
        t = trace[src[np.asarray(mesh.domains[dom]["simplices"], dtype=np.int64)]]
        r = refl(t)
        for s in range(dos_lim):
            i0 = _find_rows(fk, fo, _desc_key(rot(t, s), nfp))
            i1 = _find_rows(fk, fo, _desc_key(rot(r, s), nfp))
            if deg == "full":
                domains.setdefault(dom, []).append(np.stack([i0, i1], axis=1))
            elif deg == "unit":
                domains[f"{dom}#{s}"] = [np.stack([i0, i1], axis=1).ravel()]
            elif deg == "half":
                domains[f"{dom}#{s}.0"], domains[f"{dom}#{s}.1"] = [i0], [i1]
            else:
                raise ValueError(f"copy_degree {deg!r} not supported; use 'full', 'unit' or 'half'")
        if deg == "full":  # reference order: per simplex, per sector, (direct, mirrored)
            domains[dom] = [np.stack(domains[dom], axis=1).reshape(-1)]
        for k in [k for k in domains if k == dom or k.startswith(dom + "#")]:
            domains[k] = {"dimension": dim, "simplices": np.concatenate(domains[k]).astype(np.int64)} if isinstance(domains[k], list) else domains[k]
    out = Mesh.__new__(Mesh)
    out.name = out.file = mesh.file
    out.points, out.triangles, out.tetrahedra, out.domains = fpts, tris, tets, domains
    out.tri2tet, out._edge_of = None, None
    out.lines = lines
    a, b = tets[:, [0, 0, 0, 1, 1, 2]], tets[:, [1, 2, 3, 2, 3, 3]]
    lkey = _desc_key(lines, nfp)
    lo = np.argsort(lkey, kind="stable")
    out._edge_of = _find_rows(lkey[lo], lo, _desc_key(np.stack([a.ravel(), b.ravel()], axis=1), nfp)).reshape(-1, 6)
    d = SymInfo(DOS, naxis, nxbloch, nxsector, naxis_ln, nbloch_ln - naxis_ln, nsector_ln - naxis_ln, unit=unit)
    d.nbody, d.shiftbody, d.nxsymmetry, d.n, d.p = nbody, shiftbody, nxsym, n, p0
    out.dos = d if unit else 1
    out.sym_info = d
    return out


# ---------------------------------------------------------------------------------------------
# octosplit: uniform refinement, every tetrahedron -> 8 (src/Meshutils.jl:589-747), vectorised
# ---------------------------------------------------------------------------------------------
def octosplit(mesh):
    """Split every edge at its centre: 8 tetrahedra per tetrahedron (the inner octahedron is cut along its shortest diagonal, ties in
    the reference's order AB-CD, AC-BD, AD-BC), 4 triangles per triangle; the first N points of the new mesh are the old points, the
    edge midpoints follow in the order of mesh.lines; domains hold the (sorted) children of their simplices."""
    mesh.collect_lines()
    npts = mesh.points.shape[1]
    lines, tet, tri = mesh.lines, mesh.tetrahedra, mesh.triangles
    pts = np.concatenate([mesh.points, 0.5 * (mesh.points[:, lines[:, 0]] + mesh.points[:, lines[:, 1]])], axis=1)
    A, B, Cc, D = tet.T
    AB, AC, AD, BC, BD, CD = (mesh._edge_of + npts).T
    d1 = np.linalg.norm(pts[:, AB] - pts[:, CD], axis=0)
    d2 = np.linalg.norm(pts[:, AC] - pts[:, BD], axis=0)
    d3 = np.linalg.norm(pts[:, AD] - pts[:, BC], axis=0)
    c1 = (d1 <= d2) & (d1 <= d3)
    c2 = ~c1 & (d2 <= d1) & (d2 <= d3)
    sel = lambda x, y, z: np.where(c1, x, np.where(c2, y, z))
    P, Q = sel(AB, AC, AD), sel(CD, BD, BC)  # the chosen diagonal
    ring = [(sel(AC, AB, AC), sel(AD, AD, CD)), (sel(AD, AD, CD), sel(BD, CD, BD)),
            (sel(BD, CD, BD), sel(BC, BC, AB)), (sel(BC, BC, AB), sel(AC, AB, AC))]
    kids = [np.stack([A, AB, AC, AD], axis=1), np.stack([B, AB, BC, BD], axis=1), np.stack([Cc, AC, BC, CD], axis=1),
            np.stack([D, AD, BD, CD], axis=1)] + [np.stack([P, Q, r0, r1], axis=1) for r0, r1 in ring]
    tets = np.stack(kids, axis=1).reshape(-1, 4)  # child k of tet i at 8 i + k
    if len(tri):
        a, b, c = tri.T
        ab, ac, bc = (mesh.edge_index(a, b) + npts, mesh.edge_index(a, c) + npts, mesh.edge_index(b, c) + npts)
        tris = np.stack([np.stack([a, ab, ac], axis=1), np.stack([b, ab, bc], axis=1), np.stack([c, ac, bc], axis=1),
                         np.stack([ab, ac, bc], axis=1)], axis=1).reshape(-1, 3)
    else:
        tris = np.zeros((0, 3), dtype=np.int64)
    domains = {}
    for dom, d in mesh.domains.items():
        s = np.asarray(d["simplices"], dtype=np.int64)
        k = 8 if d["dimension"] == 3 else 4
        domains[dom] = {"dimension": d["dimension"], "simplices": (s[:, None] * k + np.arange(k)[None, :]).ravel()}
    new = Mesh(mesh.file, raw=(pts, np.zeros((0, 2), dtype=np.int64), tris, tets, domains))
    for d in new.domains.values():
        d["simplices"] = np.sort(d["simplices"])
    return new
