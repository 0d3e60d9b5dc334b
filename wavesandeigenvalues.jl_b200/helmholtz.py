"""Host mirror of ``Helmholtz.discretize`` (src/Helmholtz.jl:54-581) over the CUDA assembly kernels.

    L = discretize(mesh, dscrp, C; order="lin"|"quad", mass_weighting=True, output=False)

``dscrp[domain] = (type, data)`` with the reference's types (Julia symbols -> strings):
  "interior" ()                      M (omega^2) + K                Helmholtz.jl:233-237
  "mass" ()                          M                              :239-240
  "stiff" (funcs, args, txt)         K with a custom scalar         :242-249
  "admittance" (sym, val)            C with omega*sym               :263-275
  "admittance" (Yfunc,)              C with omega*Y(omega)          :276-279
  "flame" 9/10-tuple                 Q with n*exp(-i omega tau)     :295-301, 325-344
  "flame" 6-tuple (custom FTF)       Q with FTF(omega,k)            :302-311
  "flame" 5-tuple                    Q with the plain parameter FTF :312-319
  "flameresponse" 7-tuple            Q with eps                     :346-359
and the auxiliary mass-weighting term -lambda*M_all (:528-540, 571-574).  With ``b="b"`` (a unit-cell mesh carrying
``mesh.dos``) the operators are blochified (src/Bloch.jl:4-112, Helmholtz.jl:82-118,509-513,541-569).
  "speaker" (sym, val, *admittance)  source vector m with (admittance scalar)*sym, plus C   :251-257, 488-503
With ``source=True`` the call returns ``(L, rhs)`` as in the reference (:576-577; tutorial_09_forcing.md: ``sol = L(ω) \\ Array(rhs(ω))``
is ``L(ω).solve(rhs(ω))`` here).  Flames and speakers on Bloch meshes and Hermite elements are outside the accelerated path and
raise NotImplementedError.

What changes relative to the reference is only where the work happens: the per-element loops
(:411-463, 468-476, 532-539) and ``sparse()`` run as CUDA kernels on a pattern computed once per domain.
"""
import numpy as np

from . import _lib
from .meshutils import aggregate_elements
from .nlevp import (DeviceMatrix, LinearOperatorFamily, Sigma_nexp_az2mzit, Term, VectorFamily, exp_az2mzit, exp_delay, generate_stsp_z,
                    generate_z_g_z, get_context, pow1, pow2)


def _split_c(mesh, C, npts):
    """Helmholtz.jl:59-74: per-tetrahedron or per-point speed of sound -> (C_tet, C_tri)."""
    C = np.asarray(C, dtype=float)
    if len(C) == len(mesh.tetrahedra):
        if mesh.tri2tet is None:
            mesh.link_triangles_to_tetrahedra()
        return C, C[mesh.tri2tet]
    if len(C) == npts:
        return C[mesh.tetrahedra[:, :4]], (C[mesh.triangles[:, :3]] if len(mesh.triangles) else np.zeros((0, 3)))
    raise ValueError("length(C) must be the number of tetrahedra or the number of points")


class Discretization:
    """Book-keeping of one discretize() call (kept on the family as ``L.discretization``): the patterns, the
    device matrices and the recipe that produced them, so that the numeric phase can be repeated in place
    (new speed of sound and/or new vertex coordinates on the same topology) without any symbolic work."""

    def __init__(self):
        self.patterns = {}   # (kind, domain) -> pattern id
        self.timing = {}
        self.ops = []        # recipe: dicts with "op", pattern/matrix ids and the element subsets
        self.n_tet = self.n_tri = self.dim = 0
        self.ctx = self.mesh = None
        self.mesh_serial, self.mesh_args = None, None  # what wae_mesh_set was called with (order, tets, tris, dim) and when

    def reassemble(self, C, points=None):
        """Re-run every assembly kernel into the existing device matrices.  ``points`` (3 x N) replaces the vertex
        coordinates first (shape_sensitivity.jl:111,120 re-discretises perturbed meshes this way)."""
        ctx, mesh = self.ctx, self.mesh
        if getattr(ctx, "mesh_serial", None) != self.mesh_serial:
            # the context holds one mesh at a time and a later discretize() / shape-sensitivity call replaced this family's: make it
            # resident again (same connectivity, so the patterns and pair programs of this family stay valid)
            self.mesh_serial = ctx.mesh_set(*self.mesh_args[:1], mesh.points.T, *self.mesh_args[1:])
        if points is not None:
            mesh.points = np.asarray(points, dtype=float)
            ctx.mesh_update_points(mesh.points.T)
        C_tet, C_tri = _split_c(mesh, C, mesh.points.shape[1])
        ms = 0.0
        for op in self.ops:
            k = op["op"]
            if k == "mk":
                ctx.assemble_mk(op["pid"], C_tet[op["simplices"]], reuse=op["mats"])
            elif k == "mass":
                ctx.assemble(op["pid"], _lib.OP_MASS, None, scale=op["scale"], reuse=op["mat"])
            elif k == "stiff":
                ctx.assemble(op["pid"], _lib.OP_STIFF, C_tet[op["simplices"]], reuse=op["mat"])
            elif k == "boundary":
                ctx.assemble(op["pid"], _lib.OP_BOUNDARY, C_tri[op["simplices"]], reuse=op["mat"])
            elif k == "wallsrc":  # refreshed in place (the rhs term holds this array unless push merged it with another speaker)
                op["vec"][:] = ctx.assemble_wallsrc(op["simplices"], C_tri[op["simplices"]], self.dim)
            elif k == "flame":
                nlocal = (op["gamma"] - 1) / op["rho"] * op["nglobal"] / mesh.compute_size(op["domain"])
                ctx.assemble_flame(op["simplices"], op["ref_idx"], op["x_ref"], op["n_ref"], nlocal, reuse=op["mat"])
            ms += ctx.last_ms("assemble")
        self.timing["reassemble_ms"] = ms
        return ms


def _bloch_setup(mesh, order, b, L):
    """Helmholtz.jl:82-118: scalar functions of the Bloch wave number and the DOF folding of blochify."""
    import math
    from .meshutils import bloch_dof_maps
    from .nlevp import exp_az, generate_1_gz, generate_gz_hz, generate_Sigma_y_exp_ikx
    dos = mesh.dos
    dphi = 2 * math.pi / dos.DOS

    def exp_plus(z, k):
        return exp_az(z, dphi * 1j, k)

    def exp_minus(z, k):
        return exp_az(z, -dphi * 1j, k)
    filt = np.fft.fft(np.concatenate([[1.0 / dos.DOS], np.zeros(dos.DOS - 1)]))  # Helmholtz.jl:92-95 (FFTW.fft of a scaled delta)
    bloch_filt = generate_Sigma_y_exp_ikx(filt)
    funcs = [(), (exp_plus,), (exp_minus,), (bloch_filt,), (generate_gz_hz(bloch_filt, exp_plus),), (generate_gz_hz(bloch_filt, exp_minus),)]
    txts = ["", f"*exp(i{b}2π/{dos.DOS})", f"*exp(-i{b}2π/{dos.DOS})", f"*δ({b})", f"*δ({b})*exp(i{b}2π/{dos.DOS})", f"*δ({b})*exp(-i{b}2π/{dos.DOS})"]
    new, image, axis, red = bloch_dof_maps(mesh, order)
    flag = image.astype(np.uint8) | (axis.astype(np.uint8) << 1)
    L.params[b] = 0j
    return {"funcs": funcs, "txts": txts, "new": new, "flag": flag, "dim": red, "n_class": 3 if dos.naxis == 0 else 6,
            "anti_filt": generate_1_gz(bloch_filt), "naxis": dos.naxis, "naxis_ln": dos.naxis_ln, "nxbloch": dos.nxbloch}


def discretize(mesh, dscrp, C, order="lin", b="__none__", mass_weighting=True, source=False, output=False, ctx=None):
    bloch = None
    ctx = ctx or get_context()
    triangles, tetrahedra, dim = aggregate_elements(mesh, order)
    npts = mesh.points.shape[1]
    C_tet, C_tri = _split_c(mesh, C, npts)

    serial = ctx.mesh_set(1 if order == "lin" else 2, mesh.points.T, tetrahedra, triangles, dim)
    L = LinearOperatorFamily(["ω", "λ"], [0.0, float("inf")])
    rhs = VectorFamily(["ω"], [0.0])  # Helmholtz.jl:79
    disc = Discretization()
    disc.mesh_serial, disc.mesh_args = serial, (1 if order == "lin" else 2, tetrahedra, triangles, dim)
    disc.n_tet, disc.n_tri, disc.dim = len(tetrahedra), len(triangles), dim
    disc.ctx, disc.mesh = ctx, mesh
    L.discretization = disc

    if b != "__none__":
        if getattr(mesh, "dos", 1) == 1:
            raise ValueError("Bloch discretisation needs a unit-cell mesh (mesh.dos)")
        bloch = _bloch_setup(mesh, order, b, L)
        full_dim, dim = dim, bloch["dim"]

    def dm(mid):
        return DeviceMatrix(ctx, dim, [(mid, 1.0)])

    def push_bloch(elem_kind, simplices, kind, c, func, arg, txt, opr):
        """Helmholtz.jl:509-513: one Term per Bloch class with the class scalar appended to the term's functions."""
        _, mids = ctx.assemble_bloch(elem_kind, simplices, kind, c, 1.0, dim, bloch["new"], bloch["flag"], bloch["n_class"])
        for mid, f, t in zip(mids, bloch["funcs"], bloch["txts"]):
            if ctx.mat_info(mid)[2] == 0:
                continue
            L.push(Term(dm(mid), tuple(func) + f, tuple(arg) + (((b,),) if f else ()), txt + t, opr))

    def tet_pattern(domain):
        key = (3, domain)
        if key not in disc.patterns:
            s = np.asarray(mesh.domains[domain]["simplices"], dtype=np.int64)
            full = len(s) == len(tetrahedra) and np.array_equal(s, np.arange(len(tetrahedra)))
            full_key = (3, "__all__")
            if full and full_key in disc.patterns:
                disc.patterns[key] = disc.patterns[full_key]
            else:
                disc.patterns[key] = ctx.pattern_build(3, None if full else s)[0]
                if full:
                    disc.patterns[full_key] = disc.patterns[key]
        return disc.patterns[key]

    for domain, (typ, data) in dscrp.items():
        simplices = np.asarray(mesh.domains[domain]["simplices"], dtype=np.int64)
        if bloch is not None and typ in ("interior", "mass", "stiff", "admittance"):
            if typ in ("interior", "mass"):
                push_bloch(3, simplices, _lib.OP_MASS, None, (pow2,), (("ω",),), "ω^2", "M")
            if typ == "interior":
                push_bloch(3, simplices, _lib.OP_STIFF, C_tet[simplices], (), (), "", "K")
            elif typ == "stiff":
                funcs, args, txt = data
                for a in args:
                    for p_ in a:
                        L.params[p_] = 0.0
                push_bloch(3, simplices, _lib.OP_STIFF, C_tet[simplices], tuple(funcs), tuple(args), txt, "K")
            elif typ == "admittance":
                if len(data) != 2:
                    raise NotImplementedError("Bloch + functional admittance is not on the accelerated path")
                adm_sym, adm_val = data
                L.params.setdefault(adm_sym, complex(adm_val))
                push_bloch(2, simplices, _lib.OP_BOUNDARY, C_tri[simplices], (pow1, pow1), (("ω",), (adm_sym,)), "ω*" + adm_sym, "C")
        elif bloch is not None:
            raise NotImplementedError(f"descriptor type {typ!r} with Bloch periodicity is not on the accelerated path")
        elif typ == "interior":
            pid = tet_pattern(domain)
            im, ik = ctx.assemble_mk(pid, C_tet[simplices])
            disc.ops.append({"op": "mk", "pid": pid, "simplices": simplices, "mats": (im, ik)})
            disc.timing[f"{domain}/MK_ms"] = ctx.last_ms("assemble")
            L.push(Term(dm(im), (pow2,), (("ω",),), "ω^2", "M"))
            L.push(Term(dm(ik), (), (), "", "K"))
        elif typ == "mass":
            pid = tet_pattern(domain)
            mid = ctx.assemble(pid, _lib.OP_MASS)
            disc.ops.append({"op": "mass", "pid": pid, "scale": 1.0, "mat": mid})
            L.push(Term(dm(mid), (pow2,), (("ω",),), "ω^2", "M"))
        elif typ == "stiff":
            funcs, args, txt = data
            for a in args:
                for p in a:
                    L.params[p] = 0.0
            pid = tet_pattern(domain)
            mid = ctx.assemble(pid, _lib.OP_STIFF, C_tet[simplices])
            disc.ops.append({"op": "stiff", "pid": pid, "simplices": simplices, "mat": mid})
            L.push(Term(dm(mid), tuple(funcs), tuple(args), txt, "K"))
        elif typ in ("admittance", "speaker"):
            if typ == "speaker":  # Helmholtz.jl:253-258
                speak_sym, speak_val = data[:2]
                rhs.params[speak_sym] = complex(speak_val)
                data = tuple(data[2:])
                if not data:
                    raise ValueError(":speaker needs an admittance after (symbol, value): the source scalar is built from it")
            if len(data) == 2:
                adm_sym, adm_val = data
                if adm_sym not in L.params:  # :266-271
                    L.params[adm_sym] = complex(adm_val)
                    if typ == "speaker":
                        rhs.params[adm_sym] = complex(adm_val)
                bfunc, barg, btxt = (pow1, pow1), (("ω",), (adm_sym,)), "ω*" + adm_sym
            elif len(data) == 1:
                bfunc, barg, btxt = (generate_z_g_z(data[0]),), (("ω",),), "ω*Y(ω)"
            elif len(data) == 4:  # Helmholtz.jl:279-285: state-space admittance C_s (i omega I - A)^-1 B + D
                bfunc, barg, btxt = (generate_z_g_z(generate_stsp_z(*data)),), (("ω",),), "ω*C_s(iωI-A)^{-1}B"
            else:
                raise ValueError("Data length does not match :admittance option!")
            pid = ctx.pattern_build(2, simplices)[0]
            disc.patterns[(2, domain)] = pid
            mid = ctx.assemble(pid, _lib.OP_BOUNDARY, C_tri[simplices])
            disc.ops.append({"op": "boundary", "pid": pid, "simplices": simplices, "mat": mid})
            if typ == "speaker":  # opr == :m (:488-503) comes before :C in the reference's `make`; the two families are independent
                m = ctx.assemble_wallsrc(simplices, C_tri[simplices], dim)
                disc.ops.append({"op": "wallsrc", "simplices": simplices, "vec": m})
                rhs.push(Term(m, bfunc + (pow1,), barg + ((speak_sym,),), "speaker", "m"))
            L.push(Term(dm(mid), bfunc, barg, btxt, "C"))
        elif typ in ("flame", "flameresponse", "fancyflame"):
            ref_idx = -1
            if typ == "flame" and len(data) == 9:
                gamma, rho, nglobal, x_ref, n_ref, n_sym, tau_sym, n_val, tau_val = data
                kind = "ntau"
            elif typ == "flame" and len(data) == 10:
                gamma, rho, nglobal, ref_idx, x_ref, n_ref, n_sym, tau_sym, n_val, tau_val = data
                kind = "ntau"
            elif typ == "flame" and len(data) == 6:
                gamma, rho, nglobal, x_ref, n_ref, FTF = data
                kind = "ftf"
            elif typ == "flame" and len(data) == 5:
                gamma, rho, nglobal, x_ref, n_ref = data
                kind = "plain"
            elif typ == "flameresponse":
                gamma, rho, nglobal, x_ref, n_ref, eps_sym, eps_val = data
                kind = "eps"
            elif typ == "fancyflame":  # Helmholtz.jl:363-400
                gamma, rho, nglobal, x_ref, n_ref, n_sym, tau_sym, a_sym, n_val, tau_val, a_val = data
                kind = "fancy"
            else:
                raise ValueError("Data length does not match :flame option!")
            nlocal = (gamma - 1) / rho * nglobal / mesh.compute_size(domain)
            if kind == "ntau":
                L.params.setdefault(n_sym, complex(n_val))
                L.params.setdefault(tau_sym, complex(tau_val))
                ffunc, farg, ftxt = (pow1, exp_delay), ((n_sym,), ("ω", tau_sym)), f"{n_sym}*exp(-iω{tau_sym})"
            elif kind == "ftf":
                ffunc, farg, ftxt = (FTF,), (("ω",),), "FTF(ω)"
            elif kind == "fancy":
                if isinstance(n_val, (int, float, complex)):
                    for sym_, val_ in ((n_sym, n_val), (tau_sym, tau_val), (a_sym, a_val)):
                        L.params.setdefault(sym_, complex(val_))
                    ffunc, farg = (pow1, exp_az2mzit), ((n_sym,), ("ω", tau_sym, a_sym))
                    ftxt = f"{n_sym}* exp({a_sym}ω^2-iω{tau_sym})"
                else:
                    arg, ftxt = ["ω"], ""
                    for ns, ts, as_, nv, tv, av in zip(n_sym, tau_sym, a_sym, n_val, tau_val, a_val):
                        L.params[ns], L.params[ts], L.params[as_] = complex(nv), complex(tv), complex(av)
                        arg += [ns, ts, as_]
                        ftxt += f"[{ns}* exp({as_}ω^2-iω{ts})+"
                    ffunc, farg, ftxt = (Sigma_nexp_az2mzit,), (tuple(arg),), ftxt[:-1] + "]"
            elif kind == "plain":
                L.params["FTF"] = 0.0
                ffunc, farg, ftxt = (pow1,), (("FTF",),), "FTF"
            else:
                L.params.setdefault(eps_sym, complex(eps_val))
                ffunc, farg, ftxt = (pow1,), ((eps_sym,),), eps_sym
            if ref_idx < 0:
                ref_idx = mesh.find_tetrahedron_containing_point(x_ref)
                if ref_idx < 0:
                    raise ValueError("reference point x_ref is not inside the mesh")
            if ref_idx in set(simplices.tolist()):
                print("Warning: your reference point is inside the domain of heat release. (short-circuited FTF!)")
            pid, mid, _ = ctx.assemble_flame(simplices, ref_idx, x_ref, n_ref, nlocal)
            disc.ops.append({"op": "flame", "simplices": simplices, "ref_idx": ref_idx, "x_ref": x_ref, "n_ref": n_ref,
                             "gamma": gamma, "rho": rho, "nglobal": nglobal, "domain": domain, "mat": mid})
            disc.patterns[("Q", domain)] = pid
            L.push(Term(dm(mid), ffunc, farg, ftxt, "Q"))
        else:
            raise NotImplementedError(f"descriptor type {typ!r} is not on the accelerated path")

    if bloch is not None:
        # Helmholtz.jl:541-574: weighting matrix = blochified mass with the three parts summed; axis DOFs get the penalty term D
        _, mids = ctx.assemble_bloch(3, None, _lib.OP_MASS, None, -1.0, dim, bloch["new"], bloch["flag"], 1)
        aux = dm(mids[0])
        if bloch["naxis"] > 0:
            import scipy.sparse as sp
            npts = mesh.points.shape[1]
            DI = list(range(bloch["naxis"]))
            if order == "quad":
                DI += [k - bloch["nxbloch"] for k in range(npts, npts + bloch["naxis_ln"])]
            diag = aux.to_scipy().diagonal()
            D = sp.csc_matrix((1.0 / diag[DI], (DI, DI)), shape=(dim, dim))
            L.push(Term(DeviceMatrix.from_scipy(D, ctx), (bloch["anti_filt"],), ((b,),), "(1-δ(b))", "D"))
        L.push(Term(aux, (pow1,), (("λ",),), "-λ", "__aux__"))
        return (L, rhs) if source else L
    if mass_weighting:
        # Helmholtz.jl:528-540,572-574: -M over ALL tetrahedra
        key = (3, "__all__")
        if key not in disc.patterns:
            disc.patterns[key] = ctx.pattern_build(3, None)[0]
        mid = ctx.assemble(disc.patterns[key], _lib.OP_MASS, None, scale=-1.0)
        disc.ops.append({"op": "mass", "pid": disc.patterns[key], "scale": -1.0, "mat": mid})
        L.push(Term(dm(mid), (pow1,), (("λ",),), "-λ", "__aux__"))
    return (L, rhs) if source else L
