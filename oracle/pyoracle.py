"""ctypes wrapper of the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Import this only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product (gauxc_b200/) never imports it.
"""
import ctypes as C
import glob
import os
import subprocess
import sys
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lib = None

KERN = dict(SLATER_X=0, VWN5_C=1, PBE_X=2, PBE_C=3, PW92_C=5, B88_X=6, LYP_C=7, REVPBE_X=8)
FUNCTIONALS = {
    "SVWN5": (False, [("SLATER_X", 1.0), ("VWN5_C", 1.0)]),
    "LDA": (False, [("SLATER_X", 1.0)]),
    "VWN5": (False, [("VWN5_C", 1.0)]),
    "SPW92": (False, [("SLATER_X", 1.0), ("PW92_C", 1.0)]),
    "PBE": (True, [("PBE_X", 1.0), ("PBE_C", 1.0)]),
    "PBE0": (True, [("PBE_X", 0.75), ("PBE_C", 1.0)]),
    "BLYP": (True, [("B88_X", 1.0), ("LYP_C", 1.0)]),  # forward-mode differentiation of the spin-resolved forms
    "B3LYP": (True, [("SLATER_X", 0.08), ("B88_X", 0.72), ("VWN5_C", 0.19), ("LYP_C", 0.81)]),  # libxc hyb_gga_xc_b3lyp
    "REVPBE": (True, [("REVPBE_X", 1.0), ("PBE_C", 1.0)]),
    "REVPBE0": (True, [("REVPBE_X", 0.75), ("PBE_C", 1.0)]),
}


def build():
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


def find_blas():
    pats = [os.path.join(p, "scipy.libs", "libscipy_openblas*.so*") for p in sys.path] + \
           [os.path.join(p, "opencv_python_headless.libs", "libopenblas*.so*") for p in sys.path]
    for pat in pats:
        for f in sorted(glob.glob(pat)):
            return f
    return ""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.oracle_init_blas.restype = C.c_char_p
        L.oracle_init_blas.argtypes = [C.c_char_p]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def init_blas():
    try:  # loads scipy's bundled OpenBLAS and its libgfortran so that dlopen resolves
        import scipy.linalg  # noqa: F401
    except Exception:
        pass
    return lib().oracle_init_blas(find_blas().encode()).decode()


def init_gau2grid(enable=True):
    """Collocation through the reference's own gau2grid (oracle/_ref/libgau2grid.so) when it was built: the calls of
    gau2grid_collocation[_gradient].  Returns True if it is in use.  bench.py's CPU legs switch it on; the parity tests
    keep the restatement and check the two against each other."""
    p = os.path.join(_HERE, "_ref", "libgau2grid.so")
    L = lib()
    L.oracle_init_gau2grid.restype = C.c_int
    L.oracle_init_gau2grid.argtypes = [C.c_char_p, C.c_int]
    return bool(L.oracle_init_gau2grid(p.encode() if os.path.exists(p) else b"", int(bool(enable))))


def num_threads():
    return lib().oracle_num_threads()


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip)


def _func(name):
    gga, ks = FUNCTIONALS[name.upper()]
    kern = (C.c_int * len(ks))(*[KERN[k] for k, _ in ks])
    coef = (C.c_double * len(ks))(*[c for _, c in ks])
    return gga, len(ks), kern, coef


def functional(name, rho, sigma=None):
    gga, nk, kern, coef = _func(name)
    rho = np.ascontiguousarray(rho, np.float64)
    n = len(rho)
    sg = np.ascontiguousarray(sigma if sigma is not None else np.zeros(n), np.float64)
    eps, vr, vs = np.zeros(n), np.zeros(n), np.zeros(n)
    lib().oracle_functional(nk, kern, coef, int(gga), n, _d(rho), _d(sg), _d(eps), _d(vr), _d(vs))
    return eps, vr, vs


def collocation(flat_basis, shell_list, points, gradient=False):
    l, pure, nprim, alpha, coeff, origin = flat_basis
    sl = np.ascontiguousarray(shell_list, np.int32)
    pts = np.ascontiguousarray(points, np.float64)
    nbe = int(sum((2 * l[s] + 1) if pure[s] else (l[s] + 1) * (l[s] + 2) // 2 for s in sl))
    n = len(pts)
    ev = np.zeros((n, nbe))
    dx, dy, dz = (np.zeros((n, nbe)) for _ in range(3)) if gradient else (None, None, None)
    lib().oracle_collocation(len(l), _i(l), _i(pure), _i(nprim), _d(alpha), _d(coeff), _d(origin), len(sl),
                             _i(sl), n, _d(pts), int(gradient), _d(ev), _d(dx), _d(dy), _d(dz))
    return (ev, dx, dy, dz) if gradient else ev


def collocation_d2(flat_basis, shell_list, points):
    """value, gradient (x, y, z) and Hessian (xx, xy, xz, yy, yz, zz): ten [npts][nbe] arrays."""
    l, pure, nprim, alpha, coeff, origin = flat_basis
    sl = np.ascontiguousarray(shell_list, np.int32)
    pts = np.ascontiguousarray(points, np.float64)
    nbe = int(sum((2 * l[s] + 1) if pure[s] else (l[s] + 1) * (l[s] + 2) // 2 for s in sl))
    n = len(pts)
    out = np.zeros((10, n, nbe))
    lib().oracle_collocation_d2(len(l), _i(l), _i(pure), _i(nprim), _d(alpha), _d(coeff), _d(origin), len(sl),
                                _i(sl), n, _d(pts), _d(out))
    return out


def exc_grad(flat_basis, shell_to_center, coords, nbf, P, tasks, func_name, include_weight_derivatives=True):
    """RKS EXC gradient [natoms][3]; tasks as LoadBalancer.export_tasks() (with SSF-modified weights)."""
    l, pure, nprim, alpha, coeff, origin = flat_basis
    gga, nk, kern, coef = _func(func_name)
    Pf = np.asfortranarray(np.asarray(P, np.float64))
    s2c = np.ascontiguousarray(shell_to_center, np.int32)
    xyz = np.ascontiguousarray(coords, np.float64)
    tn = np.ascontiguousarray(tasks["npts"], np.int32)
    ts = np.ascontiguousarray(tasks["nshells"], np.int32)
    sl = np.ascontiguousarray(tasks["shell_lists"], np.int32)
    ip = np.ascontiguousarray(tasks["iParent"], np.int32)
    dn = np.ascontiguousarray(tasks["dist_nearest"], np.float64)
    pts = np.ascontiguousarray(tasks["points"], np.float64)
    w = np.ascontiguousarray(tasks["weights"], np.float64)
    g = np.zeros((len(xyz), 3))
    lib().oracle_exc_grad(len(l), _i(l), _i(pure), _i(nprim), _d(alpha), _d(coeff), _d(origin), _i(s2c), len(xyz),
                          _d(xyz), nbf, _d(Pf), Pf.shape[0], len(tn), _i(tn), _i(ts), _i(sl), _i(ip), _d(dn), _d(pts),
                          _d(w), nk, kern, coef, int(gga), int(bool(include_weight_derivatives)), _d(g))
    return g


def exc_grad_uks(flat_basis, shell_to_center, coords, nbf, Ps, Pz, tasks, func_name, include_weight_derivatives=True):
    """UKS EXC gradient [natoms][3], (Ps, Pz) = (P_alpha + P_beta, P_alpha - P_beta)."""
    l, pure, nprim, alpha, coeff, origin = flat_basis
    gga, nk, kern, coef = _func(func_name)
    Psf = np.asfortranarray(np.asarray(Ps, np.float64))
    Pzf = np.asfortranarray(np.asarray(Pz, np.float64))
    s2c = np.ascontiguousarray(shell_to_center, np.int32)
    xyz = np.ascontiguousarray(coords, np.float64)
    tn = np.ascontiguousarray(tasks["npts"], np.int32)
    ts = np.ascontiguousarray(tasks["nshells"], np.int32)
    sl = np.ascontiguousarray(tasks["shell_lists"], np.int32)
    ip = np.ascontiguousarray(tasks["iParent"], np.int32)
    dn = np.ascontiguousarray(tasks["dist_nearest"], np.float64)
    pts = np.ascontiguousarray(tasks["points"], np.float64)
    w = np.ascontiguousarray(tasks["weights"], np.float64)
    g = np.zeros((len(xyz), 3))
    lib().oracle_exc_grad_uks(len(l), _i(l), _i(pure), _i(nprim), _d(alpha), _d(coeff), _d(origin), _i(s2c), len(xyz),
                              _d(xyz), nbf, _d(Psf), _d(Pzf), Psf.shape[0], len(tn), _i(tn), _i(ts), _i(sl), _i(ip),
                              _d(dn), _d(pts), _d(w), nk, kern, coef, int(gga), int(bool(include_weight_derivatives)),
                              _d(g))
    return g


def ssf_weights(coords, task_npts, task_iparent, task_dist_nearest, points, weights):
    coords = np.ascontiguousarray(coords, np.float64)
    tn = np.ascontiguousarray(task_npts, np.int32)
    tp = np.ascontiguousarray(task_iparent, np.int32)
    td = np.ascontiguousarray(task_dist_nearest, np.float64)
    pts = np.ascontiguousarray(points, np.float64)
    w = np.array(weights, dtype=np.float64, copy=True)
    lib().oracle_ssf_weights(len(coords), _d(coords), len(tn), _i(tn), _i(tp), _d(td), _d(pts), _d(w))
    return w


def exc_vxc(flat_basis, nbf, P, tasks, func_name, task_stride=1):
    """tasks: dict(npts, nshells, shell_lists, points, weights) as LoadBalancer.export_tasks()."""
    l, pure, nprim, alpha, coeff, origin = flat_basis
    gga, nk, kern, coef = _func(func_name)
    Pf = np.asfortranarray(np.asarray(P, np.float64))
    tn = np.ascontiguousarray(tasks["npts"], np.int32)
    ts = np.ascontiguousarray(tasks["nshells"], np.int32)
    sl = np.ascontiguousarray(tasks["shell_lists"], np.int32)
    pts = np.ascontiguousarray(tasks["points"], np.float64)
    w = np.ascontiguousarray(tasks["weights"], np.float64)
    vxc = np.zeros((nbf, nbf), order="F")
    out3 = np.zeros(3)
    lib().oracle_exc_vxc(len(l), _i(l), _i(pure), _i(nprim), _d(alpha), _d(coeff), _d(origin), nbf, _d(Pf),
                         Pf.shape[0], len(tn), _i(tn), _i(ts), _i(sl), _d(pts), _d(w), nk, kern, coef,
                         int(gga), int(task_stride), _d(vxc), _d(out3))
    return dict(exc=out3[0], nel=out3[1], flops=out3[2], vxc=vxc)


def exc_vxc_uks(flat_basis, nbf, Ps, Pz, tasks, func_name):
    """UKS, LDA functionals (SVWN5, LDA, VWN5): (Ps, Pz) = (P_alpha + P_beta, P_alpha - P_beta)."""
    l, pure, nprim, alpha, coeff, origin = flat_basis
    gga, nk, kern, coef = _func(func_name)
    Psf = np.asfortranarray(np.asarray(Ps, np.float64))
    Pzf = np.asfortranarray(np.asarray(Pz, np.float64))
    tn = np.ascontiguousarray(tasks["npts"], np.int32)
    ts = np.ascontiguousarray(tasks["nshells"], np.int32)
    sl = np.ascontiguousarray(tasks["shell_lists"], np.int32)
    pts = np.ascontiguousarray(tasks["points"], np.float64)
    w = np.ascontiguousarray(tasks["weights"], np.float64)
    vs, vz = np.zeros((nbf, nbf), order="F"), np.zeros((nbf, nbf), order="F")
    out3 = np.zeros(3)
    fn = lib().oracle_exc_vxc_uks_gga if gga else lib().oracle_exc_vxc_uks_lda
    fn(len(l), _i(l), _i(pure), _i(nprim), _d(alpha), _d(coeff), _d(origin), nbf, _d(Psf), _d(Pzf), Psf.shape[0],
       len(tn), _i(tn), _i(ts), _i(sl), _d(pts), _d(w), nk, kern, coef, _d(vs), _d(vz), _d(out3))
    return dict(exc=out3[0], nel=out3[1], flops=out3[2], vxc_s=vs, vxc_z=vz)


def functional_pol_gga(func_name, rho_a, rho_b, s_aa, s_ab, s_bb):
    """Spin-polarised GGA (B88, LYP): eps, (vrho_a, vrho_b), (vsigma_aa, vsigma_ab, vsigma_bb)."""
    gga, nk, kern, coef = _func(func_name)
    n = len(rho_a)
    r2 = np.ascontiguousarray(np.stack([rho_a, rho_b], 1).ravel(), np.float64)
    g3 = np.ascontiguousarray(np.stack([s_aa, s_ab, s_bb], 1).ravel(), np.float64)
    eps, v2, v3 = np.zeros(n), np.zeros(2 * n), np.zeros(3 * n)
    lib().oracle_functional_pol_gga(nk, kern, coef, n, _d(r2), _d(g3), _d(eps), _d(v2), _d(v3))
    return eps, (v2[0::2].copy(), v2[1::2].copy()), (v3[0::3].copy(), v3[1::3].copy(), v3[2::3].copy())


def functional_pol_lda(func_name, rho_a, rho_b):
    """Spin-polarised LDA: eps (per particle of rho_a + rho_b), d(rho eps)/d rho_a, d(rho eps)/d rho_b."""
    gga, nk, kern, coef = _func(func_name)
    assert not gga
    r2 = np.ascontiguousarray(np.stack([rho_a, rho_b], 1).ravel(), np.float64)
    n = len(rho_a)
    eps, v2 = np.zeros(n), np.zeros(2 * n)
    lib().oracle_functional_pol_lda(nk, kern, coef, n, _d(r2), _d(eps), _d(v2))
    return eps, v2[0::2].copy(), v2[1::2].copy()


# ---- oracle/_ref: the reference's own gau2grid, compiled from /root/reference -----------------
def gau2grid():
    p = os.path.join(_HERE, "_ref", "libgau2grid.so")
    if not os.path.exists(p):
        return None
    return C.CDLL(p)


def gau2grid_collocation(flat_basis, shell_list, points, gradient=False):
    """Exactly the call sequence of gau2grid_collocation[_gradient]
    (local_work_driver/host/reference/gau2grid_collocation.cxx:25-116)."""
    g = gau2grid()
    l, pure, nprim, alpha, coeff, origin = flat_basis
    pts = np.ascontiguousarray(points, np.float64)
    n = len(pts)
    outs = []
    for s in shell_list:
        nf = (2 * l[s] + 1) if pure[s] else (l[s] + 1) * (l[s] + 2) // 2
        order = 300 if pure[s] else 400  # GG_SPHERICAL_CCA / GG_CARTESIAN_CCA
        c = np.ascontiguousarray(coeff[s, :nprim[s]])
        a = np.ascontiguousarray(alpha[s, :nprim[s]])
        o = np.ascontiguousarray(origin[s])
        ph = np.zeros((nf, n))
        if gradient:
            px, py, pz = np.zeros((nf, n)), np.zeros((nf, n)), np.zeros((nf, n))
            g.gg_collocation_deriv1(C.c_int(int(l[s])), C.c_ulong(n), _d(pts), C.c_ulong(3), C.c_int(int(nprim[s])),
                                    _d(c), _d(a), _d(o), C.c_int(order), _d(ph), _d(px), _d(py), _d(pz))
            outs.append((ph, px, py, pz))
        else:
            g.gg_collocation(C.c_int(int(l[s])), C.c_ulong(n), _d(pts), C.c_ulong(3), C.c_int(int(nprim[s])),
                             _d(c), _d(a), _d(o), C.c_int(order), _d(ph))
            outs.append((ph,))
    res = [np.concatenate([o[k] for o in outs], axis=0).T.copy() for k in range(4 if gradient else 1)]
    return tuple(res) if gradient else res[0]


def gau2grid_collocation_d2(flat_basis, shell_list, points):
    """gg_collocation_deriv2 as called by gau2grid_collocation_hessian
    (local_work_driver/host/reference/gau2grid_collocation.cxx:153-200): ten [npts][nbe] arrays."""
    g = gau2grid()
    l, pure, nprim, alpha, coeff, origin = flat_basis
    pts = np.ascontiguousarray(points, np.float64)
    n = len(pts)
    outs = []
    for s in shell_list:
        nf = (2 * l[s] + 1) if pure[s] else (l[s] + 1) * (l[s] + 2) // 2
        order = 300 if pure[s] else 400
        c = np.ascontiguousarray(coeff[s, :nprim[s]])
        a = np.ascontiguousarray(alpha[s, :nprim[s]])
        o = np.ascontiguousarray(origin[s])
        m = [np.zeros((nf, n)) for _ in range(10)]
        g.gg_collocation_deriv2(C.c_int(int(l[s])), C.c_ulong(n), _d(pts), C.c_ulong(3), C.c_int(int(nprim[s])),
                                _d(c), _d(a), _d(o), C.c_int(order), *[_d(x) for x in m])
        outs.append(m)
    return np.stack([np.concatenate([o[k] for o in outs], axis=0).T for k in range(10)])
