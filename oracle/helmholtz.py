"""Oracle (test infrastructure): Helmholtz.discretize, per-element loops as in the reference.

Restates src/Helmholtz.jl:19-33 (outer), :54-81, :120-191 (element wrappers),
:232-345 (descriptor parsing: :interior, :mass, :stiff-less subset, :admittance
(sym,val), :flame 9/10-tuple n-tau, :flameresponse), :405-524 (element loops +
sparse()), :528-540,:571-580 (mass weighting / __aux__ term).
:speaker descriptors and the source=true return mode (:251-258, 488-503, 524-526,
576-577) and the custom-FTF / plain-FTF flame variants (:302-319) are restated too.
"""
import numpy as np
import scipy.sparse as sp

from . import fem
from .mesh import aggregate_elements
from .nlevp import (LinearOperatorFamily, Term, exp_az2mzit, exp_delay, generate_stsp_z, generate_z_g_z, pow1, pow2,
                    sigma_nexp_az2mzit)


def _div(a, b):
    """IEEE division (Julia: x/0.0 == Inf): an empty flame domain has size 0 (shape_sensitivity.jl:50-69 builds such domains)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.float64(a) / np.float64(b))


def _sparse(I, J, V, dim):
    """SparseArrays.sparse(I,J,V,dim,dim): duplicates summed, explicit zeros kept."""
    return sp.csc_matrix(sp.coo_matrix((np.asarray(V, dtype=complex), (I, J)), shape=(dim, dim)))


def blochify(ii, jj, mm, naxis, nxbloch, nsector, naxis_ln, nsector_ln, N_points, axis=True):
    """src/Bloch.jl:4-112 with 0-based indices: fold the image DOFs onto the Bloch reference plane and sort every
    triplet into the plain / plus / minus (and axis) sets.  nsector, naxis_ln, nsector_ln are COUNTS as in the reference
    (naxis_ln and nsector_ln already include N_points)."""
    blochshift = nsector - naxis
    blochshift_ln = nsector_ln - naxis_ln
    sets = {k: ([], [], []) for k in ("", "+", "-", "a", "+a", "-a")}
    for i, j, m in zip(ii, jj, mm):
        i, j = int(i) + 1, int(j) + 1  # the reference's 1-based comparisons
        if i <= N_points:
            i_check = i > nsector
            if i_check:
                i -= blochshift
        else:
            i_check = i > nsector_ln
            if i_check:
                i -= blochshift_ln
        if j <= N_points:
            j_check = j > nsector
            if j_check:
                j -= blochshift
        else:
            j_check = j > nsector_ln
            if j_check:
                j -= blochshift_ln
        axis_check = axis and (i <= naxis or j <= naxis or N_points < i <= naxis_ln or N_points < j <= naxis_ln)
        if i > N_points:
            i -= nxbloch
        if j > N_points:
            j -= nxbloch
        key = ("" if i_check == j_check else ("+" if j_check else "-")) + ("a" if axis_check else "")
        I, J, M = sets[key]
        I.append(i - 1); J.append(j - 1); M.append(m)
    keys = ("", "+", "-") if naxis == 0 else ("", "+", "-", "a", "+a", "-a")
    return [sets[k] for k in keys]


def discretize(mesh, dscrp, C, order="lin", mass_weighting=True, triplets=None, b=None, source=False):
    """Returns the LinearOperatorFamily (and, with source=True, the vector family `rhs`, Helmholtz.jl:576-577).  If `triplets`
    is a dict, the raw COO triplets of every operator are stored in it (used by the pattern tests)."""
    o = 1 if order == "lin" else 2
    triangles, tetrahedra, dim = aggregate_elements(mesh, order)
    npts = mesh.points.shape[1]
    C = np.asarray(C, dtype=float)
    if len(C) == len(mesh.tetrahedra):
        C_tet = [C[i] for i in range(len(C))]
        if mesh.tri2tet is None:
            mesh.link_triangles_to_tetrahedra()
        C_tri = [C[j] for j in mesh.tri2tet]
    elif len(C) == npts:
        C_tet = [C[t[:4]] for t in tetrahedra]
        C_tri = [C[t[:3]] for t in triangles]
    else:
        raise ValueError("C must be per-tetrahedron or per-point")
    L = LinearOperatorFamily(["ω", "λ"], [0.0, float("inf")])
    rhs = LinearOperatorFamily(["ω"], [0.0])  # Helmholtz.jl:79
    P = mesh.points
    bloch = b is not None
    if bloch:  # Helmholtz.jl:82-118
        import math
        from .nlevp import exp_az
        dos = mesh.dos
        naxis, nxbloch = dos.naxis, dos.nxbloch
        nsector = naxis + dos.nxsector
        naxis_ln = dos.naxis_ln + npts
        nsector_ln = dos.naxis_ln + dos.nxsector_ln + npts
        dphi = 2 * math.pi / dos.DOS
        exp_plus = lambda z, k: exp_az(z, dphi * 1j, k)
        exp_minus = lambda z, k: exp_az(z, -dphi * 1j, k)
        filt_y = np.fft.fft(np.concatenate([[1.0 / dos.DOS], np.zeros(dos.DOS - 1)]))

        def bloch_filt(z, n):  # algebra.jl:276-288
            N = len(filt_y)
            f = sum((k**n if n else 1) * y * np.exp(2j * math.pi * k / N * z) for k, y in enumerate(filt_y))
            return f * (2j * math.pi / N) ** n
        anti_bloch_filt = lambda z, k: (1 - bloch_filt(z, k)) if k == 0 else -bloch_filt(z, k)
        gz_hz = lambda g, h: (lambda z, k: sum(math.comb(k, i) * h(z, k - i) * g(z, i) for i in range(k + 1)))
        bloch_funcs = [(), (exp_plus,), (exp_minus,), (bloch_filt,), (gz_hz(bloch_filt, exp_plus),), (gz_hz(bloch_filt, exp_minus),)]
        L.params[b] = 0j
        dim -= dos.nxbloch + (dos.nxbloch_ln if order == "quad" else 0)
        L.bloch_funcs = bloch_funcs

    def stiff(ct, c):
        return -c**2 * fem.tet_stiff(ct, o) if np.ndim(c) == 0 else -fem.tet_stiff_cc1(ct, c, o)

    def bound(ct, c):
        return c * fem.tri_mass(ct, o) if np.ndim(c) == 0 else fem.tri_mass_c1(ct, c, o)

    def wallsrc(ct, c):  # Helmholtz.jl:193-210
        return c * fem.tri_src(ct, o) if np.ndim(c) == 0 else fem.tri_src_c1(ct, c, o)

    for domain, (typ, data) in dscrp.items():
        simplices = mesh.domains[domain]["simplices"]
        if typ == "interior":
            make = ["M", "K"]
        elif typ == "mass":
            make = ["M"]
        elif typ in ("admittance", "speaker"):
            make = []
            if typ == "speaker":  # Helmholtz.jl:253-258
                make.append("m")
                speak_sym, speak_val = data[:2]
                rhs.params[speak_sym] = complex(speak_val)
                data = tuple(data[2:])
            if len(data) > 0:
                make.append("C")
            if len(data) == 2:  # Helmholtz.jl:262-273
                adm_sym, adm_val = data
                if adm_sym not in L.params:
                    L.params[adm_sym] = complex(adm_val)
                    if typ == "speaker":
                        rhs.params[adm_sym] = complex(adm_val)
                bfunc, barg, btxt = (pow1, pow1), (("ω",), (adm_sym,)), "ω*" + adm_sym
            elif len(data) == 1:  # :274-278
                bfunc, barg, btxt = (generate_z_g_z(data[0]),), (("ω",),), "ω*Y(ω)"
            elif len(data) == 4:  # :279-285 state space
                bfunc, barg, btxt = (generate_z_g_z(generate_stsp_z(*data)),), (("ω",),), "ω*C_s(iωI-A)^{-1}B"
        elif typ == "fancyflame":  # Helmholtz.jl:363-400
            make = ["Q"]
            gamma, rho, nglobal, x_ref, n_ref, n_sym, tau_sym, a_sym, n_val, tau_val, a_val = data
            nlocal = _div((gamma - 1) / rho * nglobal, mesh.compute_size(domain))
            if isinstance(n_val, (int, float, complex)):
                for sym_, val_ in ((n_sym, n_val), (tau_sym, tau_val), (a_sym, a_val)):
                    L.params.setdefault(sym_, complex(val_))
                ffunc, farg, ftxt = (pow1, exp_az2mzit), ((n_sym,), ("ω", tau_sym, a_sym)), f"{n_sym}* exp({a_sym}ω^2-iω{tau_sym})"
            else:
                arg, ftxt = ["ω"], ""
                for ns, ts, as_, nv, tv, av in zip(n_sym, tau_sym, a_sym, n_val, tau_val, a_val):
                    L.params[ns], L.params[ts], L.params[as_] = complex(nv), complex(tv), complex(av)
                    arg += [ns, ts, as_]
                    ftxt += f"[{ns}* exp({as_}ω^2-iω{ts})+"
                ffunc, farg, ftxt = (sigma_nexp_az2mzit,), (tuple(arg),), ftxt[:-1] + "]"
            ref_idx = mesh.find_tetrahedron_containing_point(x_ref)
        elif typ in ("flame", "flameresponse"):
            make = ["Q"]
            if typ == "flame" and len(data) == 9:
                gamma, rho, nglobal, x_ref, n_ref, n_sym, tau_sym, n_val, tau_val = data
                ref_idx = -1
            elif typ == "flame" and len(data) == 10:
                gamma, rho, nglobal, ref_idx, x_ref, n_ref, n_sym, tau_sym, n_val, tau_val = data
            elif typ == "flame" and len(data) == 6:  # :302-311 custom FTF(ω, k)
                gamma, rho, nglobal, x_ref, n_ref, FTF = data
                ref_idx = -1
            elif typ == "flame" and len(data) == 5:  # :312-319 plain parameter FTF
                gamma, rho, nglobal, x_ref, n_ref = data
                ref_idx = -1
            elif typ == "flameresponse":
                gamma, rho, nglobal, x_ref, n_ref, eps_sym, eps_val = data
                ref_idx = -1
            else:
                raise ValueError("Data length does not match :flame option!")
            nlocal = _div((gamma - 1) / rho * nglobal, mesh.compute_size(domain))
            if typ == "flame" and len(data) in (9, 10):
                L.params.setdefault(n_sym, complex(n_val))
                L.params.setdefault(tau_sym, complex(tau_val))
                ffunc, farg, ftxt = (pow1, exp_delay), ((n_sym,), ("ω", tau_sym)), f"{n_sym}*exp(-iω{tau_sym})"
            elif typ == "flame" and len(data) == 6:
                ffunc, farg, ftxt = (FTF,), (("ω",),), "FTF(ω)"
            elif typ == "flame":
                L.params["FTF"] = 0j
                ffunc, farg, ftxt = (pow1,), (("FTF",),), "FTF"
            else:
                L.params.setdefault(eps_sym, complex(eps_val))
                ffunc, farg, ftxt = (pow1,), ((eps_sym,),), eps_sym
            if ref_idx < 0:
                ref_idx = mesh.find_tetrahedron_containing_point(x_ref)
        else:
            raise NotImplementedError(typ)

        for opr in make:
            I, J, V = [], [], []
            if opr in ("M", "K"):
                for s in simplices:
                    smplx = tetrahedra[s]
                    ct = fem.CooTrafo(P[:, smplx[:4]])
                    ii, jj = fem.create_indices(smplx)
                    vv = fem.tet_mass(ct, o) if opr == "M" else stiff(ct, C_tet[s])
                    V.extend(vv.T.ravel()); I.extend(ii.T.ravel()); J.extend(jj.T.ravel())
                func, arg, txt = ((pow2,), (("ω",),), "ω^2") if opr == "M" else ((), (), "")
            elif opr == "C":
                for s in simplices:
                    smplx = triangles[s]
                    ct = fem.CooTrafo(P[:, smplx[:3]])
                    ii, jj = fem.create_indices(smplx)
                    vv = bound(ct, C_tri[s])
                    V.extend(vv.T.ravel()); I.extend(ii.T.ravel()); J.extend(jj.T.ravel())
                V = list(np.asarray(V, dtype=complex) * -1j)
                func, arg, txt = bfunc, barg, btxt
            elif opr == "Q":
                S, Irow = [], []
                for s in simplices:
                    smplx = tetrahedra[s]
                    ct = fem.CooTrafo(P[:, smplx[:4]])
                    S.extend(fem.tet_src(ct, o)); Irow.extend(smplx)
                smplx = tetrahedra[ref_idx]
                ct = fem.CooTrafo(P[:, smplx[:4]])
                G = -nlocal * fem.tet_grad_at(ct, n_ref, x_ref, o)
                for a, i in zip(S, Irow):  # Helmholtz.jl:19-33 outer()
                    for b, j in zip(G, smplx):
                        V.append(a * b); I.append(i); J.append(j)
                func, arg, txt = ffunc, farg, ftxt
            elif opr == "m":  # Helmholtz.jl:488-503: source vector of a speaker (membrane) boundary
                if bloch:
                    raise NotImplementedError("speaker on a Bloch mesh")
                for s in simplices:
                    smplx = triangles[s]
                    ct = fem.CooTrafo(P[:, smplx[:3]])
                    V.extend(wallsrc(ct, C_tri[s])); I.extend(smplx)
                V = np.asarray(V, dtype=complex) / 1j
                vec = sp.csc_matrix(sp.coo_matrix((V, (I, np.zeros(len(I), dtype=int))), shape=(dim, 1)))  # sparsevec(I,V,dim)
                rhs.push(Term(vec, tuple(bfunc) + (pow1,), tuple(barg) + ((speak_sym,),), "speaker", "m"))
                continue
            if triplets is not None:
                triplets.setdefault(opr, []).append((np.array(I), np.array(J), np.array(V, dtype=complex)))
            if bloch:  # Helmholtz.jl:509-513
                parts = blochify(I, J, V, naxis, nxbloch, nsector, naxis_ln, nsector_ln, npts)
                for (i_, j_, v_), f in zip(parts, bloch_funcs):
                    L.push(Term(_sparse(i_, j_, v_, dim), tuple(func) + f, tuple(arg) + (((b,),) if f else ()), txt, opr))
            else:
                L.push(Term(_sparse(I, J, V, dim), func, arg, txt, opr))

    if mass_weighting or bloch:
        I, J, V = [], [], []
        for smplx in tetrahedra:
            ct = fem.CooTrafo(P[:, smplx[:4]])
            ii, jj = fem.create_indices(smplx)
            vv = fem.tet_mass(ct, o)
            V.extend(vv.T.ravel()); I.extend(ii.T.ravel()); J.extend(jj.T.ravel())
        if triplets is not None:
            triplets.setdefault("__aux__", []).append((np.array(I), np.array(J), -np.array(V, dtype=complex)))
        if bloch:  # Helmholtz.jl:541-569
            parts = blochify(I, J, V, naxis, nxbloch, nsector, naxis_ln, nsector_ln, npts, axis=False)[:3]
            I = [x for p_ in parts for x in p_[0]]
            J = [x for p_ in parts for x in p_[1]]
            V = [x for p_ in parts for x in p_[2]]
            M = _sparse(I, J, -np.asarray(V), dim)
            if naxis > 0:
                DI = list(range(naxis))
                if order == "quad":
                    DI += [k - nxbloch for k in range(npts, naxis_ln)]
                DV = [1 / M[k, k] for k in DI]
                L.push(Term(_sparse(DI, DI, DV, dim), (anti_bloch_filt,), ((b,),), "(1-δ(b))", "D"))
        else:
            M = _sparse(I, J, -np.asarray(V), dim)
        L.push(Term(M, (pow1,), (("λ",),), "-λ", "__aux__"))
    return (L, rhs) if source else L
