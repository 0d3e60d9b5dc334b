"""Convert the reference's gmsh-4.1 tutorial meshes into compact .npz fixtures.

Runs ONLY in the build container (reads /root/reference/docs/src/*.msh).  The
fixtures hold the RAW parse (file order, 0-based, duplicates kept) so that the
sorting/uniquifying logic of ``Mesh`` is still exercised by whoever loads them.

    python tests/golden/make_mesh_fixtures.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle.mesh import read_msh4  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def convert(src, dst):
    points, lines, tris, tets, domains = read_msh4(src)
    names = sorted(domains)
    out = dict(points=points, lines=np.array(lines, dtype=np.int32).reshape(-1, 2),
               tris=np.array(tris, dtype=np.int32).reshape(-1, 3), tets=np.array(tets, dtype=np.int32).reshape(-1, 4),
               dom_names=np.array(names), dom_dims=np.array([domains[n]["dimension"] for n in names]))
    for i, n in enumerate(names):
        out[f"dom_{i}"] = np.array(domains[n]["simplices"], dtype=np.int32)
    np.savez_compressed(dst, **out)
    print(dst, points.shape, len(lines), len(tris), len(tets), names)


if __name__ == "__main__":
    convert("/root/reference/docs/src/Rijke_mm.msh", os.path.join(HERE, "rijke_mm_mesh.npz"))
    convert("/root/reference/docs/src/NTNU_12.msh", os.path.join(HERE, "ntnu_12_mesh.npz"))
