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
