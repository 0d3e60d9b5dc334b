"""Oracle (test infrastructure): the parts of Meshutils the hot path consumes.

Restates (0-based indices here, 1-based in the reference):
  read_msh4                         src/Meshutils.jl:272-402
  Mesh constructor (uniquify, sort) src/Meshutils.jl:92-165
  simplex ordering rule             src/Mesh/sorter.jl:9-31  (ascending in the descending-sorted vertex tuple)
  collect_lines!                    src/Meshutils.jl:831-840
  link_triangles_to_tetrahedra!     src/Meshutils.jl:516-548
  compute_size!                     src/Meshutils.jl:757-780
  find_tetrahedron_containing_point src/Meshutils.jl:800-815
  generate_field                    src/Meshutils.jl:1079-1098
"""
import numpy as np


def simplex_key(s):
    return tuple(sorted((int(x) for x in s), reverse=True))


def unique_sorted(simplices):
    """insert_smplx! applied in file order: first occurrence of a vertex set wins, list sorted by key."""
    seen = {}
    for s in simplices:
        k = simplex_key(s)
        if k not in seen:
            seen[k] = list(s)
    keys = sorted(seen)
    return [seen[k] for k in keys], {k: i for i, k in enumerate(keys)}


def read_msh4(fname):
    lines, tris, tets = [], [], []
    tag2dom, ent2dom, domains = {}, [dict(), dict(), dict(), dict()], {}
    points = None
    with open(fname) as f:
        rd = f.readline
        while True:
            line = rd()
            if not line:
                break
            fld = line.strip()[1:]
            if fld == "PhysicalNames":
                for _ in range(int(rd())):
                    dim, tag, dom = rd().split()
                    dom = dom[1:-1]
                    tag2dom[tag] = dom
                    domains[dom] = {"dimension": int(dim), "simplices": []}
            elif fld == "Entities":
                counts = [int(x) for x in rd().split()]
                for d, cnt in enumerate(counts):
                    off = 4 if d == 0 else 7
                    for _ in range(cnt):
                        sp = rd().split()
                        nph = int(sp[off])
                        ent2dom[d][sp[0]] = [tag2dom[t] for t in sp[off + 1 : off + 1 + nph]]
            elif fld == "Nodes":
                nblk, nnodes, _, _ = (int(x) for x in rd().split())
                points = np.empty((3, nnodes))
                for _ in range(nblk):
                    _, _, _, nin = (int(x) for x in rd().split())
                    tags = [int(rd()) for _ in range(nin)]
                    for t in tags:
                        points[:, t - 1] = [float(x) for x in rd().split()[:3]]
            elif fld == "Elements":
                nblk = int(rd().split()[0])
                for _ in range(nblk):
                    sp = rd().split()
                    edim, etag, etype, nin = int(sp[0]), sp[1], int(sp[2]), int(sp[3])
                    for _ in range(nin):
                        nodes = [int(x) - 1 for x in rd().split()[1:]]
                        lst = {1: lines, 2: tris, 4: tets}.get(etype)
                        if lst is None:
                            continue
                        lst.append(nodes)
                        for dom in ent2dom[edim].get(etag, []):
                            domains[dom]["simplices"].append(len(lst) - 1)
    return points, lines, tris, tets, domains


class Mesh:
    """Mesh(file; scale) of the reference: sorted unique simplex lists + remapped domains."""

    def __init__(self, fname=None, scale=1.0, raw=None):
        if raw is None:
            raw = read_msh4(fname)
        points, lines, tris, tets, domains = raw
        self.name = fname
        self.points = np.asarray(points, dtype=float) * scale
        self.lines, lmap = unique_sorted(lines)
        self.triangles, tmap = unique_sorted(tris)
        self.tetrahedra, ttmap = unique_sorted(tets)
        self.domains = {}
        for dom, d in domains.items():
            raw_list = {1: lines, 2: tris, 3: tets}[d["dimension"]]
            mp = {1: lmap, 2: tmap, 3: ttmap}[d["dimension"]]
            new, seen = [], set()
            for idx in d["simplices"]:
                j = mp[simplex_key(raw_list[idx])]
                if j not in seen:
                    seen.add(j)
                    new.append(j)
            self.domains[dom] = {"dimension": d["dimension"], "simplices": new}
        self.tri2tet = None

    # -- Meshutils.jl:831-840 -------------------------------------------------------
    def collect_lines(self):
        if len(self.lines) == 0:
            edges = []
            for t in self.tetrahedra:
                for a, b in ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)):
                    edges.append([t[a], t[b]])
            self.lines, self._lmap = unique_sorted(edges)
        elif not hasattr(self, "_lmap"):
            self._lmap = {simplex_key(l): i for i, l in enumerate(self.lines)}
        return self._lmap

    def line_idx(self, a, b):
        return self._lmap[simplex_key((a, b))]

    # -- Meshutils.jl:516-548 -------------------------------------------------------
    def link_triangles_to_tetrahedra(self):
        tmap = {simplex_key(t): i for i, t in enumerate(self.triangles)}
        t2t = np.full(len(self.triangles), -1, dtype=np.int64)
        for it, tet in enumerate(self.tetrahedra):
            for f in ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)):
                j = tmap.get(simplex_key([tet[k] for k in f]))
                if j is not None:
                    t2t[j] = it  # later tets overwrite earlier ones, as in the reference loop
        self.tri2tet = t2t
        return t2t

    # -- Meshutils.jl:757-780 -------------------------------------------------------
    def compute_size(self, dom):
        d = self.domains[dom]
        V = 0.0
        if d["dimension"] == 3:
            for i in d["simplices"]:
                X = self.points[:, self.tetrahedra[i]]
                V += abs(np.linalg.det(X[:, :3] - X[:, 3:4])) / 6
        elif d["dimension"] == 2:
            for i in d["simplices"]:
                X = self.points[:, self.triangles[i]]
                V += np.linalg.norm(np.cross(X[:, 0] - X[:, 2], X[:, 1] - X[:, 2])) / 2
        d["size"] = V
        return V

    # -- Meshutils.jl:800-815 -------------------------------------------------------
    def find_tetrahedron_containing_point(self, p):
        p = np.asarray(p, dtype=float)
        for i, tet in enumerate(self.tetrahedra):
            X = self.points[:, tet]
            xi = np.linalg.solve(X[:, :3] - X[:, 3:4], p - X[:, 3])
            xi = np.append(xi, 1 - xi.sum())
            if np.all((0 <= xi) & (xi <= 1)):
                return i
        return -1

    # -- Meshutils.jl:1079-1098 -----------------------------------------------------
    def generate_field(self, func, order="const"):
        if order == "const":
            return np.array([func(*self.points[:, t].sum(axis=1) / 4) for t in self.tetrahedra])
        return np.array([func(*self.points[:, i]) for i in range(self.points.shape[1])])


def aggregate_elements(mesh, order):
    """FEM.jl:84-116: local->global DOF lists for :lin / :quad."""
    npts = mesh.points.shape[1]
    if order == "lin":
        return [list(t) for t in mesh.triangles], [list(t) for t in mesh.tetrahedra], npts
    mesh.collect_lines()
    tets, tris = [], []
    for t in mesh.tetrahedra:
        e = [mesh.line_idx(t[a], t[b]) + npts for a, b in ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))]
        tets.append(list(t) + e)
    for t in mesh.triangles:
        e = [mesh.line_idx(t[a], t[b]) + npts for a, b in ((0, 1), (0, 2), (1, 2))]
        tris.append(list(t) + e)
    return tris, tets, npts + len(mesh.lines)


# ---------------------------------------------------------------------------------------------
# extend_mesh: half-cell -> unit cell / full annulus           src/Mesh/annular_meshes.jl:14-181, 269-570
# ---------------------------------------------------------------------------------------------
def three_points_to_plane(A):  # annular_meshes.jl:14-26 (the d < 1e-7 -> 0 rule included)
    a, b, c = A[:, 0], A[:, 1], A[:, 2]
    n = np.cross(a - c, b - c)
    n = n / np.linalg.norm(n)
    d = -float(np.dot(n, c))
    if d < 1e-7:
        d = 0.0
    return np.array([n[0], n[1], n[2], d])


def find_foot_of_perpendicular(pnt, pln):  # :61-66
    a, d = pln[:3], pln[3]
    return a * (-np.dot(a, pnt) - d) + pnt


def reflect_point_at_plane(pnt, pln):  # :46-52
    return 2 * find_foot_of_perpendicular(pnt, pln) - pnt


def make_normal_outwards(pln, testpoint):  # :74-79
    foot = find_foot_of_perpendicular(testpoint, pln)
    return pln * (-np.sign(np.dot(pln[:3], testpoint - foot)))


def find_intersection_of_two_planes(p1, p2):  # :92-109 (p = rhs \ M: least squares with a 5-vector on the left)
    n = np.cross(p1[:3], p2[:3])
    n = n / np.linalg.norm(n)
    f1 = find_foot_of_perpendicular(np.zeros(3), p1)
    f2 = find_foot_of_perpendicular(np.zeros(3), p2)
    M = np.array([[2.0, 0, 0, p1[0], p2[0]], [2.0, 0, 0, p1[1], p2[1]], [2.0, 0, 0, p1[2], p2[2]],
                  [p1[0], p1[1], p1[2], 0, 0], [p2[0], p2[1], p2[2], 0, 0]])
    rhs = np.array([0, 0, 0, np.dot(f1, p1[:3]), np.dot(f2, p2[:3])], dtype=float)
    p = np.linalg.lstsq(rhs.reshape(5, 1), M, rcond=None)[0].ravel()
    return p[:3], n


def create_rotation_matrix_around_axis(n, al):  # :116-121
    c, s = np.cos(al), np.sin(al)
    return np.array([[n[0] ** 2 * (1 - c) + c, n[0] * n[1] * (1 - c) - n[2] * s, n[0] * n[2] * (1 - c) + n[1] * s],
                     [n[1] * n[0] * (1 - c) + n[2] * s, n[1] ** 2 * (1 - c) + c, n[1] * n[2] * (1 - c) - n[0] * s],
                     [n[2] * n[0] * (1 - c) - n[1] * s, n[2] * n[1] * (1 - c) + n[0] * s, n[2] ** 2 * (1 - c) + c]])


class SymInfoO:
    pass


def extend_mesh(mesh, doms, sym_name="Symmetry", blch_name="Bloch", unit=False):
    """annular_meshes.jl:269-546 with 0-based indices.  One deviation: the reference's counter of axis lines (:484-492) yields
    naxis_ln = 1 for a mesh without axis points (its `naxis_ln==0` guard fires twice), which breaks second-order unit cells; here
    naxis_ln is the number of leading lines that lie on the axis (0 then).  First-order meshes are not affected."""
    mesh.collect_lines()
    npoints = mesh.points.shape[1]
    bloch = sorted({p for i in mesh.domains[blch_name]["simplices"] for p in mesh.triangles[i]})
    symmetry = sorted({p for i in mesh.domains[sym_name]["simplices"] for p in mesh.triangles[i]})
    sset = set(symmetry)
    axis = [p for p in bloch if p in sset]
    new_order = list(axis)
    placed = set(new_order)
    for p in bloch:
        if p not in placed:
            new_order.append(p)
            placed.add(p)
    for p in range(npoints):
        if p not in placed and p not in sset:
            new_order.append(p)
            placed.add(p)
    for p in symmetry:
        if p not in placed:
            new_order.append(p)
            placed.add(p)
    trace = np.empty(npoints, dtype=np.int64)
    trace[np.array(new_order)] = np.arange(npoints)
    t2t = mesh.link_triangles_to_tetrahedra()

    def plane_of(dom):
        si = mesh.domains[dom]["simplices"][0]
        tri = mesh.triangles[si]
        pl = three_points_to_plane(mesh.points[:, tri])
        test = [p for p in mesh.tetrahedra[t2t[si]] if p not in tri][-1]  # find_testpoint_idx: the last vertex not in the triangle
        return make_normal_outwards(pl, mesh.points[:, test])

    pln, bpln = plane_of(sym_name), plane_of(blch_name)
    nbloch, naxis, nsym = len(bloch), len(axis), len(symmetry)
    nxsym, nxbloch = nsym - naxis, nbloch - naxis
    nbody = npoints - nbloch - nxsym
    shiftbody = npoints - nbloch
    nxsector = nxbloch + nbody + nxsym + nbody
    nsector = nxsector + naxis
    pts = np.zeros((3, 2 * npoints - nsym))
    pts[:, :npoints] = mesh.points[:, new_order]
    for i in range(nbloch, npoints - nxsym):
        pts[:, i + shiftbody] = reflect_point_at_plane(pts[:, i], pln)
    for i in range(naxis, nbloch):
        pts[:, i + nxsector] = reflect_point_at_plane(pts[:, i], pln)
    phi = np.arccos(np.dot(pln[:3], -bpln[:3]))
    DOS = int(round(np.pi / phi))
    p0, n = find_intersection_of_two_planes(pln, bpln)
    phi = 2 * np.pi / DOS
    if unit:
        fpts, dos_lim = pts, 1
    else:
        dos_lim = DOS
        fpts = np.zeros((3, naxis + nxsector * DOS))
        fpts[:, :nsector] = pts[:, :nsector]
        for s in range(1, DOS):
            R = create_rotation_matrix_around_axis(n, s * phi)
            fpts[:, naxis + nxsector * s: naxis + nxsector * (s + 1)] = R @ (pts[:, naxis:naxis + nxsector] - p0[:, None]) + p0[:, None]

    def refl(i):  # get_reflected_index
        if i < naxis:
            return i
        if i < naxis + nxbloch:
            return i + nxsector
        if i < naxis + nxbloch + nbody:
            return i + shiftbody
        if i < naxis + nxbloch + nbody + nxsym:
            return i
        raise ValueError("reflected index out of range")

    def rot(i, s):  # get_rotated_index
        if i < naxis:
            return i
        return (i + nxsector * s - naxis) % (nxsector * DOS) + naxis

    def build(simplices, skip=()):
        out = []
        for si, smp in enumerate(simplices):
            if si in skip:
                continue
            t = [int(trace[p]) for p in smp]
            r = [refl(p) for p in t]
            for s in range(dos_lim):
                out.append([rot(p, s) for p in t])
                out.append([rot(p, s) for p in r])
        return unique_sorted(out)

    tets, tetmap = build(mesh.tetrahedra)
    skip = set(mesh.domains[sym_name]["simplices"])
    if not unit:
        skip |= set(mesh.domains[blch_name]["simplices"])
    tris, trimap = build(mesh.triangles, skip)
    lns = []
    for ln in mesh.lines:
        t = [int(trace[p]) for p in ln]
        lns.append(t)
        if not all(p < nbloch for p in t):
            lns.append([refl(p) for p in t])
    lines, _ = unique_sorted(lns)
    naxis_ln = sum(1 for ln in lines if all(p < naxis for p in ln))
    nbloch_ln = sum(1 for ln in lines if all(p < nbloch for p in ln))
    nxbloch_ln = nbloch_ln - naxis_ln
    nsector_ln = len(lines)
    nxsector_ln = nsector_ln - naxis_ln
    if unit:
        lines = lines + [[rot(p, 1) for p in ln] for ln in lines[naxis_ln:nbloch_ln]]
    else:
        first = lines[naxis_ln:nsector_ln]
        for s in range(1, DOS):
            lines = lines + [[rot(p, s) for p in ln] for ln in first]
    domains = {}
    for dom, deg in doms:
        dim = mesh.domains[dom]["dimension"]
        src, mp = (mesh.tetrahedra, tetmap) if dim == 3 else (mesh.triangles, trimap)
        names = {}
        for si in mesh.domains[dom]["simplices"]:
            t = [int(trace[p]) for p in src[si]]
            r = [refl(p) for p in t]
            for s in range(dos_lim):
                i0 = mp[simplex_key([rot(p, s) for p in t])]
                i1 = mp[simplex_key([rot(p, s) for p in r])]
                if deg == "full":
                    names.setdefault(dom, []).extend([i0, i1])
                elif deg == "unit":
                    names.setdefault(f"{dom}#{s}", []).extend([i0, i1])
                elif deg == "half":
                    names.setdefault(f"{dom}#{s}.0", []).append(i0)
                    names.setdefault(f"{dom}#{s}.1", []).append(i1)
                else:
                    raise ValueError(f"copy_degree {deg!r} not supported")
        for k, v in names.items():
            domains[k] = {"dimension": dim, "simplices": v}
    out = Mesh.__new__(Mesh)
    out.name = mesh.name
    out.points, out.lines, out.triangles, out.tetrahedra, out.domains, out.tri2tet = fpts, lines, tris, tets, domains, None
    d = SymInfoO()
    d.DOS, d.naxis, d.nxbloch, d.nbody, d.shiftbody, d.nxsymmetry, d.nxsector = DOS, naxis, nxbloch, nbody, shiftbody, nxsym, nxsector
    d.naxis_ln, d.nxbloch_ln, d.nxsector_ln, d.unit, d.n, d.p = naxis_ln, nxbloch_ln, nxsector_ln, unit, n, p0
    out.dos = d
    return out


def octosplit(mesh):
    """Meshutils.jl:589-747: every tetrahedron into 8, every triangle into 4 (loop restatement, 0-based)."""
    lmap = mesh.collect_lines()
    npts = mesh.points.shape[1]
    pts = np.zeros((3, npts + len(mesh.lines)))
    pts[:, :npts] = mesh.points
    for i, ln in enumerate(mesh.lines):
        pts[:, npts + i] = mesh.points[:, ln].sum(axis=1) / 2
    mid = lambda a, b: npts + lmap[simplex_key((a, b))]

    def children(tet):
        A, B, C, D = tet
        AB, AC, AD, BC, BD, CD = mid(A, B), mid(A, C), mid(A, D), mid(B, C), mid(B, D), mid(C, D)
        out = [[A, AB, AC, AD], [B, AB, BC, BD], [C, AC, BC, CD], [D, AD, BD, CD]]
        ab_cd = np.linalg.norm(pts[:, AB] - pts[:, CD])
        ac_bd = np.linalg.norm(pts[:, AC] - pts[:, BD])
        ad_bc = np.linalg.norm(pts[:, AD] - pts[:, BC])
        if ab_cd <= ac_bd and ab_cd <= ad_bc:
            out += [[AB, CD, AC, AD], [AB, CD, AD, BD], [AB, CD, BD, BC], [AB, CD, BC, AC]]
        elif ac_bd <= ab_cd and ac_bd <= ad_bc:
            out += [[AC, BD, AB, AD], [AC, BD, AD, CD], [AC, BD, CD, BC], [AC, BD, BC, AB]]
        else:
            out += [[AD, BC, AC, CD], [AD, BC, CD, BD], [AD, BC, BD, AB], [AD, BC, AB, AC]]
        return out

    def tri_children(t):
        A, B, C = t
        AB, AC, BC = mid(A, B), mid(A, C), mid(B, C)
        return [[A, AB, AC], [B, AB, BC], [C, AC, BC], [AB, AC, BC]]

    tet_kids = [children(t) for t in mesh.tetrahedra]
    tri_kids = [tri_children(t) for t in mesh.triangles]
    tets, tetmap = unique_sorted([k for ks in tet_kids for k in ks])
    tris, trimap = unique_sorted([k for ks in tri_kids for k in ks])
    domains = {}
    for dom, d in mesh.domains.items():
        kids, mp = (tet_kids, tetmap) if d["dimension"] == 3 else (tri_kids, trimap)
        domains[dom] = {"dimension": d["dimension"], "simplices": sorted(mp[simplex_key(k)] for s in d["simplices"] for k in kids[s])}
    out = Mesh.__new__(Mesh)
    out.name = mesh.name
    out.points, out.lines, out.triangles, out.tetrahedra, out.domains, out.tri2tet = pts, [], tris, tets, domains, None
    return out
