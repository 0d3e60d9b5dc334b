import sys, math, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import wae_b200 as W
from cases import load_raw_mesh, rijke_dscrp, speedofsound
mesh = W.Mesh("Rijke_mm.msh", scale=0.001, raw=load_raw_mesh("rijke_mm"))
c = mesh.generate_field(speedofsound)
L = W.discretize(mesh, rijke_dscrp(0.01, 0.001), c, order="lin")
st = {}
sol, n, flag = W.householder(L, 340 * 2 * math.pi, maxiter=12, tol=1e-11, output=True, stats=st)
print(n, flag, sol.params["ω"], st)
print("ref 1710.6977772393461 + 9.615018460173488im")
