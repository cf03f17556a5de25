"""ctypes binding of libgauxc_b200.so -- the host-side mirror of GauXC's object model.

Classes and call order follow the reference's driver (tests/standalone_driver.cxx:31-477):
Molecule -> MolGrid -> BasisSet -> LoadBalancerFactory.get_instance -> MolecularWeights
.modify_weights -> XCIntegratorFactory.get_instance -> eval_exc_vxc.  Every call goes through
the C ABI declared in include/gauxc_b200.h; there is no Python/CPU fallback: if the CUDA
library is missing or no device is present the calls raise.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GAUXC_B200_LIB: load a diagnostic build of the same library (gauxc_b200.build.build_variant)
_LIB_PATH = os.environ.get("GAUXC_B200_LIB") or os.path.join(_HERE, "libgauxc_b200.so")


class GauXCError(RuntimeError):
    pass


class _Status(C.Structure):
    _fields_ = [("code", C.c_int), ("message", C.c_char_p)]


class _Handle(C.Structure):
    _fields_ = [("type", C.c_int), ("ptr", C.c_void_p)]


class _RtHandle(C.Structure):
    _fields_ = [("type", C.c_int), ("ptr", C.c_void_p), ("device_ptr", C.c_void_p)]


class _Atom(C.Structure):
    _fields_ = [("Z", C.c_int64), ("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]


class _Shell(C.Structure):
    _fields_ = [("l", C.c_int32), ("pure", C.c_bool), ("nprim", C.c_int32),
                ("exponents", C.c_double * 32), ("coefficients", C.c_double * 32),
                ("origin", C.c_double * 3), ("shell_tolerance", C.c_double)]


class _MWSettings(C.Structure):
    _fields_ = [("weight_alg", C.c_int), ("becke_size_adjustment", C.c_bool)]


# enums (include/gauxc/c/enums.h)
RadialQuad = dict(Becke=0, MuraKnowles=1, MurrayHandyLaming=2, TreutlerAhlrichs=3)
AtomicGridSizeDefault = dict(FineGrid=0, UltraFineGrid=1, SuperFineGrid=2, GM3=3, GM5=4)
XCWeightAlg = dict(NOTPARTITIONED=0, Becke=1, SSF=2, LKO=3)
ExecutionSpace = dict(Host=0, Device=1)
PruningScheme = dict(Unpruned=0, Robust=1, Treutler=2)

_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def library_path():
    return _LIB_PATH


def lib():
    """Load libgauxc_b200.so (fails loudly if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise GauXCError("libgauxc_b200.so not built: run `python -m gauxc_b200.build` "
                         "(or __graft_entry__.build())")
    L = C.CDLL(_LIB_PATH, mode=C.RTLD_GLOBAL)
    S = C.POINTER(_Status)
    sig = {
        "gauxc_status_delete": (None, [S]),
        "gauxc_object_delete": (None, [S, C.POINTER(C.c_void_p)]),
        "gauxc_molecule_new_from_atoms": (_Handle, [S, C.POINTER(_Atom), C.c_size_t]),
        "gauxc_molecule_delete": (None, [S, C.POINTER(_Handle)]),
        "gauxc_molecule_natoms": (C.c_size_t, [S, _Handle]),
        "gauxc_basisset_new_from_shells": (_Handle, [S, C.POINTER(_Shell), C.c_size_t, C.c_bool]),
        "gauxc_basisset_delete": (None, [S, C.POINTER(_Handle)]),
        "gauxc_molgrid_new_default": (_Handle, [S, _Handle, C.c_int, C.c_int64, C.c_int, C.c_int]),
        "gauxc_molgrid_delete": (None, [S, C.POINTER(_Handle)]),
        "gauxc_runtime_environment_new": (_RtHandle, [S]),
        "gauxc_device_runtime_environment_new": (_RtHandle, [S, C.c_double]),
        "gauxc_runtime_environment_delete": (None, [S, C.POINTER(_RtHandle)]),
        "gauxc_runtime_environment_comm_rank": (C.c_int, [S, _RtHandle]),
        "gauxc_runtime_environment_comm_size": (C.c_int, [S, _RtHandle]),
        "gauxc_b200_runtime_environment_set_comm": (None, [S, _RtHandle, C.c_int, C.c_int]),
        "gauxc_load_balancer_factory_new": (_Handle, [S, C.c_int, C.c_char_p]),
        "gauxc_load_balancer_factory_delete": (None, [S, C.POINTER(_Handle)]),
        "gauxc_load_balancer_factory_get_instance": (_Handle, [S, _Handle, _RtHandle, _Handle, _Handle, _Handle]),
        "gauxc_load_balancer_delete": (None, [S, C.POINTER(_Handle)]),
        "gauxc_molecular_weights_factory_new": (_Handle, [S, C.c_int, C.c_char_p, _MWSettings]),
        "gauxc_molecular_weights_factory_delete": (None, [S, C.POINTER(_Handle)]),
        "gauxc_molecular_weights_factory_get_instance": (_Handle, [S, _Handle]),
        "gauxc_molecular_weights_modify_weights": (None, [S, _Handle, _Handle]),
        "gauxc_molecular_weights_delete": (None, [S, C.POINTER(_Handle)]),
        "gauxc_functional_from_string": (_Handle, [S, C.c_char_p, C.c_bool]),
        "gauxc_functional_from_enum": (_Handle, [S, C.c_int, C.c_bool]),
        "gauxc_functional_delete": (None, [S, C.POINTER(_Handle)]),
        "gauxc_integrator_new": (_Handle, [S, _Handle, _Handle, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]),
        "gauxc_integrator_delete": (None, [S, C.POINTER(_Handle)]),
        "gauxc_integrator_integrate_den": (None, [S, _Handle, C.c_int64, C.c_int64, _dp, C.c_int64, _dp]),
        "gauxc_integrator_eval_exc_rks": (None, [S, _Handle, C.c_int64, C.c_int64, _dp, C.c_int64, _dp]),
        "gauxc_integrator_eval_exc_vxc_rks": (None, [S, _Handle, C.c_int64, C.c_int64, _dp, C.c_int64, _dp, _dp, C.c_int64]),
        "gauxc_integrator_eval_exc_grad_rks": (None, [S, _Handle, C.c_int64, C.c_int64, _dp, C.c_int64, _dp]),
        "gauxc_b200_integrator_eval_exc_grad_uks": (None, [S, _Handle, C.c_int64, C.c_int64, _dp, C.c_int64, _dp, C.c_int64,
                                                             _dp, C.c_int]),
        "gauxc_b200_integrator_eval_exc_grad_rks": (None, [S, _Handle, C.c_int64, C.c_int64, _dp, C.c_int64, _dp,
                                                             C.c_int]),
        "gauxc_integrator_eval_exc_vxc_uks": (None, [S, _Handle, C.c_int64, C.c_int64, _dp, C.c_int64, _dp, C.c_int64,
                                                     _dp, _dp, C.c_int64, _dp, C.c_int64]),
        "gauxc_integrator_eval_exc_uks": (None, [S, _Handle, C.c_int64, C.c_int64, _dp, C.c_int64, _dp, C.c_int64, _dp]),
        "gauxc_integrator_eval_exc_grad_uks": (None, [S, _Handle, C.c_int64, C.c_int64, _dp, C.c_int64, _dp, C.c_int64,
                                                      _dp]),
        "gauxc_integrator_eval_exx_rks": (None, [S, _Handle, C.c_int64, C.c_int64, _dp, C.c_int64, _dp, C.c_int64]),
        "gauxc_molecule_new": (_Handle, [S]),
        "gauxc_basisset_new": (_Handle, [S]),
        "gauxc_molecule_read_hdf5_record": (None, [S, _Handle, C.c_char_p, C.c_char_p]),
        "gauxc_molecule_write_hdf5_record": (None, [S, _Handle, C.c_char_p, C.c_char_p]),
        "gauxc_basisset_read_hdf5_record": (None, [S, _Handle, C.c_char_p, C.c_char_p]),
        "gauxc_basisset_write_hdf5_record": (None, [S, _Handle, C.c_char_p, C.c_char_p]),
        "gauxc_b200_hdf5_dataset_size": (C.c_int64, [S, C.c_char_p, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
        "gauxc_b200_hdf5_read_dataset": (None, [S, C.c_char_p, C.c_char_p, _dp, C.c_int64]),
        "gauxc_b200_hdf5_write_dataset": (None, [S, C.c_char_p, C.c_char_p, _dp, C.POINTER(C.c_int64), C.c_int]),
        "gauxc_b200_molecule_get_atoms": (None, [S, _Handle, C.POINTER(_Atom)]),
        "gauxc_b200_nccl_get_unique_id": (None, [S, C.c_char_p]),
        "gauxc_b200_nccl_init": (None, [S, C.c_char_p, C.c_int, C.c_int]),
        "gauxc_b200_nccl_finalize": (None, [S]),
        "gauxc_b200_allreduce_device": (None, [S, C.c_void_p, C.c_size_t]),
        "gauxc_b200_integrator_eval_exc_vxc_rks_device": (None, [S, _Handle, C.c_void_p, C.c_void_p, C.c_void_p]),
        "gauxc_b200_basisset_nbf": (C.c_int64, [S, _Handle]),
        "gauxc_b200_basisset_nshells": (C.c_int64, [S, _Handle]),
        "gauxc_b200_basisset_set_shell_tolerance": (None, [S, _Handle, C.c_double]),
        "gauxc_b200_basisset_get_shell": (None, [S, _Handle, C.c_int64, _ip, _ip, _ip, _dp, _dp, _dp, _dp]),
        "gauxc_b200_load_balancer_ntasks": (C.c_int64, [S, _Handle]),
        "gauxc_b200_load_balancer_total_npts": (C.c_int64, [S, _Handle]),
        "gauxc_b200_load_balancer_task_info": (None, [S, _Handle, _ip, _ip, _ip, _ip, _dp]),
        "gauxc_b200_load_balancer_get_task": (None, [S, _Handle, C.c_int64, _dp, _dp, _ip]),
        "gauxc_b200_load_balancer_set_task_weights": (None, [S, _Handle, C.c_int64, _dp]),
        "gauxc_b200_load_balancer_set_tasks": (None, [S, _Handle, C.c_int64, _ip, _ip, _dp, _dp, _dp, _ip, _ip, C.c_int]),
        "gauxc_b200_integrator_stats": (None, [S, _Handle, _dp]),
        "gauxc_b200_integrator_set_profile": (None, [S, _Handle, C.c_int]),
        "gauxc_b200_integrator_set_vxc_root_only": (None, [S, _Handle, C.c_int]),
        "gauxc_b200_molecular_weights_last_ms": (C.c_double, [S, _Handle]),
        "gauxc_b200_lebedev": (C.c_int64, [S, C.c_int, _dp, _dp]),
        "gauxc_b200_radial": (None, [S, C.c_int, C.c_int, C.c_double, _dp, _dp]),
        "gauxc_b200_eval_collocation": (None, [S, _Handle, C.c_int64, _ip, C.c_int64, _dp, _dp, _dp, _dp, _dp]),
        "gauxc_b200_eval_collocation_hessian": (None, [S, _Handle, C.c_int64, _ip, C.c_int64, _dp, _dp, _dp, _dp, _dp,
                                                         _dp]),
        "gauxc_b200_functional_eval_host": (None, [S, _Handle, C.c_int64, _dp, _dp, _dp, _dp, _dp]),
        "gauxc_b200_functional_eval_host_pol": (None, [S, _Handle, C.c_int64, _dp, _dp, _dp, _dp, _dp]),
        "gauxc_b200_functional_eval_host_pol_full": (None, [S, _Handle, C.c_int64, _dp, _dp, _dp, _dp, _dp]),
        "gauxc_b200_functional_eval_host_pol_gga": (None, [S, C.c_int, C.POINTER(C.c_int), _dp, C.c_int64, _dp, _dp,
                                                           _dp, _dp, _dp]),
        "gauxc_b200_probe_peak": (C.c_double, [S, C.c_int]),
        "gauxc_b200_device_count": (C.c_int, []),
        "gauxc_b200_set_device": (None, [S, C.c_int]),
        "gauxc_b200_integrator_stream": (C.c_void_p, [S, _Handle]),
        "gauxc_b200_version": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


# names every symbol include/gauxc_b200.h declares (checked by tests/test_capi_symbols.py)
def declared_symbols():
    import re
    hdr = open(os.path.join(os.path.dirname(_HERE), "include", "gauxc_b200.h")).read()
    return sorted(set(re.findall(r"\b(gauxc_[a-z0-9_]+)\s*\(", hdr)))


class _St:
    """Status out-parameter; raises GauXCError on code != 0 (src/c-api/c_status.hpp)."""

    def __init__(self):
        self.s = _Status(0, None)

    def ref(self):
        return C.byref(self.s)

    def check(self):
        if self.s.code != 0:
            msg = self.s.message.decode() if self.s.message else "unknown error"
            lib().gauxc_status_delete(C.byref(self.s))
            raise GauXCError(msg)


def _call(fname, *args):
    st = _St()
    r = getattr(lib(), fname)(st.ref(), *args)
    st.check()
    return r


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class _Obj:
    _deleter = None

    def __init__(self, h):
        self.h = h

    def __del__(self):
        try:
            if self.h is not None and self.h.ptr and _lib is not None and self._deleter:
                st = _St()
                getattr(_lib, self._deleter)(st.ref(), C.byref(self.h))
        except Exception:
            pass


class Molecule(_Obj):
    _deleter = "gauxc_molecule_delete"

    def __init__(self, atoms):
        """atoms: iterable of (Z, x, y, z) in bohr."""
        self.atoms = [(int(a[0]), float(a[1]), float(a[2]), float(a[3])) for a in atoms]
        arr = (_Atom * len(self.atoms))(*[_Atom(*a) for a in self.atoms])
        super().__init__(_call("gauxc_molecule_new_from_atoms", arr, len(self.atoms)))

    @classmethod
    def from_hdf5(cls, fname, dset="/MOLECULE"):
        m = cls.__new__(cls)
        _Obj.__init__(m, _call("gauxc_molecule_new"))
        _call("gauxc_molecule_read_hdf5_record", m.h, fname.encode(), dset.encode())
        return m

    def write_hdf5(self, fname, dset="/MOLECULE"):
        _call("gauxc_molecule_write_hdf5_record", self.h, fname.encode(), dset.encode())

    def atoms(self):
        n = self.natoms()
        buf = (_Atom * n)()
        _call("gauxc_b200_molecule_get_atoms", self.h, buf)
        return [(int(a.Z), a.x, a.y, a.z) for a in buf]

    def natoms(self):
        return _call("gauxc_molecule_natoms", self.h)


def hdf5_read_dataset(fname, dset):
    """Dense FP64 dataset of an HDF5 file (dims in file order, i.e. the column-major matrices come out transposed)."""
    dims = (C.c_int64 * 4)()
    rank = C.c_int(0)
    n = _call("gauxc_b200_hdf5_dataset_size", fname.encode(), dset.encode(), dims, C.byref(rank))
    out = np.zeros(n)
    _call("gauxc_b200_hdf5_read_dataset", fname.encode(), dset.encode(), _d(out), n)
    return out.reshape([dims[i] for i in range(rank.value)])


def hdf5_write_dataset(fname, dset, a):
    a = np.ascontiguousarray(a, np.float64)
    dims = (C.c_int64 * max(1, a.ndim))(*a.shape)
    _call("gauxc_b200_hdf5_write_dataset", fname.encode(), dset.encode(), _d(a), dims, a.ndim)


class BasisSet(_Obj):
    _deleter = "gauxc_basisset_delete"

    def __init__(self, shells, normalize=True):
        """shells: list of dict(l, pure, exps, coefs, origin[, tol])."""
        arr = (_Shell * len(shells))()
        for k, s in enumerate(shells):
            n = len(s["exps"])
            if n > 32:
                raise GauXCError("more than 32 primitives")
            arr[k].l = s["l"]
            arr[k].pure = bool(s["pure"])
            arr[k].nprim = n
            for j in range(n):
                arr[k].exponents[j] = s["exps"][j]
                arr[k].coefficients[j] = s["coefs"][j]
            for j in range(3):
                arr[k].origin[j] = s["origin"][j]
            arr[k].shell_tolerance = s.get("tol", 1e-10)
        super().__init__(_call("gauxc_basisset_new_from_shells", arr, len(shells), normalize))

    @classmethod
    def from_hdf5(cls, fname, dset="/BASIS"):
        b = cls.__new__(cls)
        _Obj.__init__(b, _call("gauxc_basisset_new"))
        _call("gauxc_basisset_read_hdf5_record", b.h, fname.encode(), dset.encode())
        return b

    def write_hdf5(self, fname, dset="/BASIS"):
        _call("gauxc_basisset_write_hdf5_record", self.h, fname.encode(), dset.encode())

    def nbf(self):
        return _call("gauxc_b200_basisset_nbf", self.h)

    def nshells(self):
        return _call("gauxc_b200_basisset_nshells", self.h)

    def set_shell_tolerance(self, tol):
        _call("gauxc_b200_basisset_set_shell_tolerance", self.h, tol)

    def get_shell(self, s):
        l, pure, nprim = C.c_int32(), C.c_int32(), C.c_int32()
        cutoff = C.c_double()
        o, a, c = np.zeros(3), np.zeros(32), np.zeros(32)
        _call("gauxc_b200_basisset_get_shell", self.h, s, C.byref(l), C.byref(pure), C.byref(nprim),
              C.byref(cutoff), _d(o), _d(a), _d(c))
        return dict(l=l.value, pure=pure.value, nprim=nprim.value, cutoff=cutoff.value, origin=o,
                    alpha=a, coeff=c)

    def flat(self):
        """Flat arrays (l, pure, nprim, alpha[ns,32], coeff[ns,32], origin[ns,3]) of the
        NORMALISED shells -- the form the oracle consumes."""
        ns = self.nshells()
        l = np.zeros(ns, np.int32)
        pure = np.zeros(ns, np.int32)
        nprim = np.zeros(ns, np.int32)
        alpha = np.zeros((ns, 32))
        coeff = np.zeros((ns, 32))
        origin = np.zeros((ns, 3))
        for s in range(ns):
            d = self.get_shell(s)
            l[s], pure[s], nprim[s] = d["l"], d["pure"], d["nprim"]
            alpha[s], coeff[s], origin[s] = d["alpha"], d["coeff"], d["origin"]
        return l, pure, nprim, alpha, coeff, origin


class MolGrid(_Obj):
    _deleter = "gauxc_molgrid_delete"

    def __init__(self, mol, pruning="Unpruned", batch_size=512, radial_quad="MuraKnowles",
                 grid_size="UltraFineGrid"):
        super().__init__(_call("gauxc_molgrid_new_default", mol.h, PruningScheme[pruning], batch_size,
                               RadialQuad[radial_quad], AtomicGridSizeDefault[grid_size]))


class RuntimeEnvironment(_Obj):
    _deleter = "gauxc_runtime_environment_delete"

    def __init__(self, rank=0, size=1, device=True, fill_fraction=0.9):
        if device:
            h = _call("gauxc_device_runtime_environment_new", fill_fraction)
        else:
            h = _call("gauxc_runtime_environment_new")
        super().__init__(h)
        if size != 1 or rank != 0:
            _call("gauxc_b200_runtime_environment_set_comm", self.h, rank, size)

    def comm_rank(self):
        return _call("gauxc_runtime_environment_comm_rank", self.h)

    def comm_size(self):
        return _call("gauxc_runtime_environment_comm_size", self.h)


class LoadBalancer(_Obj):
    _deleter = "gauxc_load_balancer_delete"

    def ntasks(self):
        return _call("gauxc_b200_load_balancer_ntasks", self.h)

    def total_npts(self):
        return _call("gauxc_b200_load_balancer_total_npts", self.h)

    def task_info(self):
        n = self.ntasks()
        ip, npts, nbe, nsh = (np.zeros(n, np.int32) for _ in range(4))
        dn = np.zeros(n)
        _call("gauxc_b200_load_balancer_task_info", self.h, _i(ip), _i(npts), _i(nbe), _i(nsh), _d(dn))
        return dict(iParent=ip, npts=npts, nbe=nbe, nshells=nsh, dist_nearest=dn)

    def get_task(self, it, info=None):
        info = info or self.task_info()
        n, ns = int(info["npts"][it]), int(info["nshells"][it])
        pts, w, sl = np.zeros((n, 3)), np.zeros(n), np.zeros(ns, np.int32)
        _call("gauxc_b200_load_balancer_get_task", self.h, it, _d(pts), _d(w), _i(sl))
        return pts, w, sl

    def export_tasks(self):
        """All local tasks as flat arrays (what the oracle consumes)."""
        info = self.task_info()
        nt = len(info["npts"])
        tot = int(info["npts"].sum())
        pts, w = np.zeros((tot, 3)), np.zeros(tot)
        sl = np.zeros(int(info["nshells"].sum()), np.int32)
        po = so = 0
        for it in range(nt):
            n, ns = int(info["npts"][it]), int(info["nshells"][it])
            p = pts[po:po + n]
            ww = w[po:po + n]
            s = sl[so:so + ns]
            _call("gauxc_b200_load_balancer_get_task", self.h, it, _d(p), _d(ww), _i(s))
            po += n
            so += ns
        info.update(points=pts, weights=w, shell_lists=sl)
        return info

    def set_task_weights(self, it, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        _call("gauxc_b200_load_balancer_set_task_weights", self.h, it, _d(w))

    def set_tasks(self, npts, iParent, dist_nearest, points, weights, nshells, shell_lists,
                  weights_are_modified):
        npts = np.ascontiguousarray(npts, np.int32)
        iParent = np.ascontiguousarray(iParent, np.int32)
        dist_nearest = np.ascontiguousarray(dist_nearest, np.float64)
        points = np.ascontiguousarray(points, np.float64)
        weights = np.ascontiguousarray(weights, np.float64)
        nshells = np.ascontiguousarray(nshells, np.int32)
        shell_lists = np.ascontiguousarray(shell_lists, np.int32)
        _call("gauxc_b200_load_balancer_set_tasks", self.h, len(npts), _i(npts), _i(iParent),
              _d(dist_nearest), _d(points), _d(weights), _i(nshells), _i(shell_lists),
              int(bool(weights_are_modified)))


class LoadBalancerFactory(_Obj):
    _deleter = "gauxc_load_balancer_factory_delete"

    def __init__(self, ex="Host", kernel="Replicated"):
        super().__init__(_call("gauxc_load_balancer_factory_new", ExecutionSpace[ex], kernel.encode()))

    def get_instance(self, rt, mol, mg, basis):
        lb = LoadBalancer(_call("gauxc_load_balancer_factory_get_instance", self.h, rt.h, mol.h, mg.h, basis.h))
        lb._keep = (rt, mol, mg, basis)
        return lb


class MolecularWeights(_Obj):
    _deleter = "gauxc_molecular_weights_delete"

    def modify_weights(self, lb):
        _call("gauxc_molecular_weights_modify_weights", self.h, lb.h)

    def last_ms(self):
        return _call("gauxc_b200_molecular_weights_last_ms", self.h)


class MolecularWeightsFactory(_Obj):
    _deleter = "gauxc_molecular_weights_factory_delete"

    def __init__(self, ex="Device", lwd_kernel="Default", weight_alg="SSF"):
        super().__init__(_call("gauxc_molecular_weights_factory_new", ExecutionSpace[ex], lwd_kernel.encode(),
                               _MWSettings(XCWeightAlg[weight_alg], False)))

    def get_instance(self):
        return MolecularWeights(_call("gauxc_molecular_weights_factory_get_instance", self.h))


# include/gauxc/c/functional.h: enum GauXC_Functional (the members this build constructs)
FunctionalEnum = dict(SVWN3=0, SVWN5=1, BLYP=2, B3LYP=3, PBE=4, revPBE=5, PBE0=6, SCAN=7, LDA=19, SPW92=22,
                      VWN3=26, VWN5=27, revPBE0=34)


class Functional(_Obj):
    _deleter = "gauxc_functional_delete"

    def __init__(self, spec, polarized=False):
        self.spec = spec
        super().__init__(_call("gauxc_functional_from_string", spec.encode(), polarized))

    @classmethod
    def from_enum(cls, value, polarized=False):
        """gauxc_functional_from_enum (value: a FunctionalEnum name or the integer enumerator)."""
        f = cls.__new__(cls)
        f.spec = str(value)
        v = FunctionalEnum[value] if isinstance(value, str) else int(value)
        _Obj.__init__(f, _call("gauxc_functional_from_enum", v, polarized))
        return f

    def eval_host(self, rho, sigma=None):
        rho = np.ascontiguousarray(rho, np.float64)
        n = len(rho)
        sg = np.ascontiguousarray(sigma if sigma is not None else np.zeros(n), np.float64)
        eps, vr, vs = np.zeros(n), np.zeros(n), np.zeros(n)
        _call("gauxc_b200_functional_eval_host", self.h, n, _d(rho), _d(sg), _d(eps), _d(vr), _d(vs))
        return eps, vr, vs


    def eval_host_pol_full(self, rho_a, rho_b, s_aa, s_ab, s_bb):
        """eps, (vrho_a, vrho_b), (vsigma_aa, vsigma_ab, vsigma_bb) of the polarised functional (LDA or GGA)."""
        n = len(rho_a)
        r2 = np.ascontiguousarray(np.stack([rho_a, rho_b], 1).ravel(), np.float64)
        g3 = np.ascontiguousarray(np.stack([s_aa, s_ab, s_bb], 1).ravel(), np.float64)
        eps, v2, v3 = np.zeros(n), np.zeros(2 * n), np.zeros(3 * n)
        _call("gauxc_b200_functional_eval_host_pol_full", self.h, n, _d(r2), _d(g3), _d(eps), _d(v2), _d(v3))
        return eps, (v2[0::2].copy(), v2[1::2].copy()), (v3[0::3].copy(), v3[1::3].copy(), v3[2::3].copy())

    def eval_host_pol(self, rho_a, rho_b):
        """Spin-polarised LDA kernels (UKS) on the host: eps, d(rho eps)/d rho_a, d(rho eps)/d rho_b."""
        ra = np.ascontiguousarray(rho_a, np.float64)
        rb = np.ascontiguousarray(rho_b, np.float64)
        n = len(ra)
        eps, va, vb = np.zeros(n), np.zeros(n), np.zeros(n)
        _call("gauxc_b200_functional_eval_host_pol", self.h, n, _d(ra), _d(rb), _d(eps), _d(va), _d(vb))
        return eps, va, vb


def eval_host_pol_gga(kernels, rho_a, rho_b, s_aa, s_ab, s_bb):
    """Spin-polarised GGA kernels prepared for UKS GGA, on the host: kernels = [("B88_X", c), ("LYP_C", c)]."""
    ids = dict(B88_X=0, LYP_C=1)
    kern = (C.c_int * len(kernels))(*[ids[k] for k, _ in kernels])
    coef = np.array([c for _, c in kernels], np.float64)
    n = len(rho_a)
    r2 = np.ascontiguousarray(np.stack([rho_a, rho_b], 1).ravel(), np.float64)
    g3 = np.ascontiguousarray(np.stack([s_aa, s_ab, s_bb], 1).ravel(), np.float64)
    eps, v2, v3 = np.zeros(n), np.zeros(2 * n), np.zeros(3 * n)
    _call("gauxc_b200_functional_eval_host_pol_gga", len(kernels), kern, _d(coef), n, _d(r2), _d(g3), _d(eps),
          _d(v2), _d(v3))
    return eps, (v2[0::2].copy(), v2[1::2].copy()), (v3[0::3].copy(), v3[1::3].copy(), v3[2::3].copy())


class XCIntegrator(_Obj):
    _deleter = "gauxc_integrator_delete"

    def eval_exc_vxc(self, P):
        """RKS EXC/VXC; P is the alpha density matrix (nbf x nbf, symmetric)."""
        P = np.asarray(P, dtype=np.float64)
        if P.ndim != 2:
            raise GauXCError("P must be a matrix")
        Pf = np.asfortranarray(P)
        m, n = Pf.shape
        vxc = np.zeros((n, n), order="F")
        exc = C.c_double(0.)
        _call("gauxc_integrator_eval_exc_vxc_rks", self.h, m, n, _d(Pf), max(m, 1), C.byref(exc), _d(vxc), max(n, 1))
        return exc.value, vxc

    def eval_exc_vxc_uks(self, Ps, Pz):
        """UKS EXC / VXC_s / VXC_z (LDA functionals); Ps = P_alpha + P_beta, Pz = P_alpha - P_beta."""
        Psf = np.asfortranarray(np.asarray(Ps, dtype=np.float64))
        Pzf = np.asfortranarray(np.asarray(Pz, dtype=np.float64))
        m, n = Psf.shape
        vs, vz = np.zeros((n, n), order="F"), np.zeros((n, n), order="F")
        exc = C.c_double(0.)
        _call("gauxc_integrator_eval_exc_vxc_uks", self.h, m, n, _d(Psf), max(m, 1), _d(Pzf), max(m, 1),
              C.byref(exc), _d(vs), max(n, 1), _d(vz), max(n, 1))
        return exc.value, vs, vz

    def eval_exc_uks(self, Ps, Pz):
        Psf = np.asfortranarray(np.asarray(Ps, dtype=np.float64))
        Pzf = np.asfortranarray(np.asarray(Pz, dtype=np.float64))
        m, n = Psf.shape
        exc = C.c_double(0.)
        _call("gauxc_integrator_eval_exc_uks", self.h, m, n, _d(Psf), max(m, 1), _d(Pzf), max(m, 1), C.byref(exc))
        return exc.value

    def eval_exc_grad(self, P, natoms, include_weight_derivatives=None):
        """RKS EXC gradient [natoms][3].  include_weight_derivatives None: the reference's C entry point (default
        settings = full gradient); True / False: IntegratorSettingsEXC_GRAD through the extension."""
        Pf = np.asfortranarray(np.asarray(P, dtype=np.float64))
        m, n = Pf.shape
        g = np.zeros(3 * natoms)
        if include_weight_derivatives is None:
            _call("gauxc_integrator_eval_exc_grad_rks", self.h, m, n, _d(Pf), m, _d(g))
        else:
            _call("gauxc_b200_integrator_eval_exc_grad_rks", self.h, m, n, _d(Pf), m, _d(g),
                  int(bool(include_weight_derivatives)))
        return g.reshape(natoms, 3)

    def eval_exc_grad_uks(self, Ps, Pz, natoms, include_weight_derivatives=None):
        """UKS EXC gradient [natoms][3], (Ps, Pz) = (P_alpha + P_beta, P_alpha - P_beta)."""
        Psf = np.asfortranarray(np.asarray(Ps, dtype=np.float64))
        Pzf = np.asfortranarray(np.asarray(Pz, dtype=np.float64))
        m, n = Psf.shape
        g = np.zeros(3 * natoms)
        if include_weight_derivatives is None:
            _call("gauxc_integrator_eval_exc_grad_uks", self.h, m, n, _d(Psf), m, _d(Pzf), m, _d(g))
        else:
            _call("gauxc_b200_integrator_eval_exc_grad_uks", self.h, m, n, _d(Psf), m, _d(Pzf), m, _d(g),
                  int(bool(include_weight_derivatives)))
        return g.reshape(natoms, 3)

    def eval_exc_vxc_raw(self, m, n, P, ldp, vxc, ldv):
        exc = C.c_double(0.)
        _call("gauxc_integrator_eval_exc_vxc_rks", self.h, m, n, _d(P), ldp, C.byref(exc), _d(vxc), ldv)
        return exc.value

    def eval_exc(self, P):
        Pf = np.asfortranarray(np.asarray(P, dtype=np.float64))
        m, n = Pf.shape
        exc = C.c_double(0.)
        _call("gauxc_integrator_eval_exc_rks", self.h, m, n, _d(Pf), m, C.byref(exc))
        return exc.value

    def integrate_den(self, P):
        Pf = np.asfortranarray(np.asarray(P, dtype=np.float64))
        m, n = Pf.shape
        v = C.c_double(0.)
        _call("gauxc_integrator_integrate_den", self.h, m, n, _d(Pf), m, C.byref(v))
        return v.value

    def eval_exc_vxc_device(self, dP_ptr, dVXC_ptr, dout2_ptr):
        """Device-resident call: raw device pointers (ints), no host<->device copies."""
        _call("gauxc_b200_integrator_eval_exc_vxc_rks_device", self.h, C.c_void_p(dP_ptr),
              C.c_void_p(dVXC_ptr), C.c_void_p(dout2_ptr))

    def stats(self):
        o = np.zeros(16)
        _call("gauxc_b200_integrator_stats", self.h, _d(o))
        keys = ["local_work_ms", "total_ms", "k_colloc_ms", "k_xmat_ms", "k_zmat_ms", "k_vxc_ms",
                "launches", "f_dense", "sum_nbe_npts", "npts", "ntiles", "nbatches", "nitems", "n_el"]
        return dict(zip(keys, o[:14]))

    def stream(self):
        """cudaStream_t (int) the integrator launches on."""
        return _call("gauxc_b200_integrator_stream", self.h)

    def set_vxc_root_only(self, on):
        _call("gauxc_b200_integrator_set_vxc_root_only", self.h, int(on))

    def set_profile(self, on):
        _call("gauxc_b200_integrator_set_profile", self.h, int(on))


class XCIntegratorFactory:
    """XCIntegratorFactory(ex, input_type, integrator_kernel, lwd_kernel, reduction_kernel)
    (include/gauxc/xc_integrator/integrator_factory.hpp:40-84)."""

    def __init__(self, ex="Device", input_type="Replicated", integrator_kernel="Default",
                 lwd_kernel="Default", reduction_kernel="Default"):
        self.args = (ex, input_type, integrator_kernel, lwd_kernel, reduction_kernel)

    def get_instance(self, func, lb):
        ex, it, ik, lk, rk = self.args
        xi = XCIntegrator(_call("gauxc_integrator_new", func.h, lb.h, ExecutionSpace[ex], it.encode(),
                                ik.encode(), lk.encode(), rk.encode()))
        xi._keep = (func, lb)
        return xi


# ---- misc extension entry points -----------------------------------------------------------
def device_count():
    return lib().gauxc_b200_device_count()


def set_device(dev):
    _call("gauxc_b200_set_device", int(dev))


def probe_peak(which):
    return _call("gauxc_b200_probe_peak", {"dmma": 0, "dfma": 1, "copy": 2}[which])


def lebedev(npts):
    xyz, w = np.zeros((npts, 3)), np.zeros(npts)
    n = _call("gauxc_b200_lebedev", npts, _d(xyz), _d(w))
    assert n == npts
    return xyz, w


def radial(rq, n, R):
    r, w = np.zeros(n), np.zeros(n)
    _call("gauxc_b200_radial", RadialQuad[rq], n, R, _d(r), _d(w))
    return r, w


def eval_collocation_hessian(basis, shell_list, points):
    """Device collocation with first and second derivatives: ten [npts][nbe] arrays (value, x, y, z, xx, xy, xz, yy,
    yz, zz) -- test hook of the EXC gradient's Hessian collocation kernel."""
    sl = np.ascontiguousarray(shell_list, dtype=np.int32)
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    npts = len(pts)
    nbe = 0
    for q in sl:
        d = basis.get_shell(int(q))
        nbe += (2 * d["l"] + 1) if d["pure"] else (d["l"] + 1) * (d["l"] + 2) // 2
    out = np.zeros((10, npts, nbe))
    _call("gauxc_b200_eval_collocation_hessian", basis.h, len(sl), _i(sl), npts, _d(pts), _d(out[0]), _d(out[1]),
          _d(out[2]), _d(out[3]), _d(out[4:]))
    return out


def eval_collocation(basis, shell_list, points, gradient=False):
    sl = np.ascontiguousarray(shell_list, np.int32)
    pts = np.ascontiguousarray(points, np.float64)
    npts = len(pts)
    nbe = 0
    for s in sl:
        d = basis.get_shell(int(s))
        nbe += (2 * d["l"] + 1) if d["pure"] else (d["l"] + 1) * (d["l"] + 2) // 2
    ev = np.zeros((npts, nbe))
    if gradient:
        dx, dy, dz = np.zeros((npts, nbe)), np.zeros((npts, nbe)), np.zeros((npts, nbe))
        _call("gauxc_b200_eval_collocation", basis.h, len(sl), _i(sl), npts, _d(pts), _d(ev), _d(dx), _d(dy), _d(dz))
        return ev, dx, dy, dz
    _call("gauxc_b200_eval_collocation", basis.h, len(sl), _i(sl), npts, _d(pts), _d(ev), None, None, None)
    return ev


def nccl_get_unique_id():
    buf = C.create_string_buffer(128)
    _call("gauxc_b200_nccl_get_unique_id", buf)
    return buf.raw


def nccl_init(id_bytes, rank, size):
    _call("gauxc_b200_nccl_init", id_bytes, rank, size)


def nccl_finalize():
    _call("gauxc_b200_nccl_finalize")


def allreduce_device(ptr, n):
    _call("gauxc_b200_allreduce_device", C.c_void_p(ptr), n)
