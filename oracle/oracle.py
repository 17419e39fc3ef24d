"""TEST INFRASTRUCTURE ONLY (oracle/). ctypes bindings for the two CPU checkers.

* ``RefSim``  - the UNMODIFIED reference actor (``AOctreeSearch``, /root/reference/Source/NBody/OctreeSearch.h:111-149)
  compiled into ``oracle/_ref/liboracle_ref.so`` (oracle/ref_wrap.cpp, oracle/Makefile).
* module functions ``direct_f64`` ... ``tick`` - the plain-C restatement in ``oracle/nbody_oracle.c``.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may import this
module. The product package (parallelnbody_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "liboracle_ref.so")

PARTICLE_DTYPE = np.dtype(
    [("Mass", "<f4"), ("Position", "<f4", 3), ("Velocity", "<f4", 3), ("Acceleration", "<f4", 3)]
)  # FParticle, OctreeSearch.h:9-18 (40 bytes)
assert PARTICLE_DTYPE.itemsize == 40


def build(force: bool = False) -> None:
    """Compile the restatement (always) and the reference build (only where /root/reference exists)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "nbody_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/Source/NBody") and (force or not os.path.exists(_REF)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None
_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)


def _f(a):
    return a.ctypes.data_as(_fp)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.oracle_direct_f64.restype = C.c_double
        L.oracle_direct_f64.argtypes = [C.c_int, _fp, C.c_double, C.c_double, C.c_int, C.c_int, _dp, C.c_int]
        L.oracle_direct_f32.restype = C.c_double
        L.oracle_direct_f32.argtypes = [C.c_int, _fp, C.c_double, C.c_float, C.c_int, C.c_int, _fp, C.c_int]
        L.oracle_kick_drift.restype = None
        L.oracle_kick_drift.argtypes = [C.c_int, _fp, _fp, _fp, C.c_float]
        L.oracle_cube_size.restype = C.c_float
        L.oracle_cube_size.argtypes = [C.c_int, _fp]
        L.oracle_energy.restype = None
        L.oracle_energy.argtypes = [C.c_int, _fp, _fp, C.c_double, C.c_double, _dp, _dp, C.c_int]
        L.oracle_bh_build.restype = C.c_void_p
        L.oracle_bh_build.argtypes = [C.c_int, _fp, _fp, C.c_float, C.c_double, C.c_float, C.POINTER(C.c_int)]
        L.oracle_bh_free.restype = None
        L.oracle_bh_free.argtypes = [C.c_void_p]
        L.oracle_bh_num_nodes.restype = C.c_longlong
        L.oracle_bh_num_nodes.argtypes = [C.c_void_p]
        L.oracle_bh_root.restype = None
        L.oracle_bh_root.argtypes = [C.c_void_p, _fp, _fp]
        L.oracle_bh_forces.restype = C.c_longlong
        L.oracle_bh_forces.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, _fp, C.c_int]
        L.oracle_bh_leaf_boxes.restype = C.c_longlong
        L.oracle_bh_leaf_boxes.argtypes = [C.c_void_p, _fp, C.c_longlong]
        L.oracle_tick.restype = C.c_int
        L.oracle_tick.argtypes = [C.c_int, _fp, _fp, _fp, C.c_float, C.c_float, C.c_double, C.c_float, C.c_int, _fp, C.c_int]
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _bind_actor(R):
    R.ref_create.restype = C.c_void_p
    R.ref_destroy.argtypes = [C.c_void_p]
    R.ref_set_particles.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    R.ref_get_particles.argtypes = [C.c_void_p, C.c_void_p]
    R.ref_num.argtypes = [C.c_void_p]
    R.ref_set_dt.argtypes = [C.c_void_p, C.c_float]
    R.ref_get_dt.argtypes = [C.c_void_p]
    R.ref_get_dt.restype = C.c_float
    R.ref_set_show_octree.argtypes = [C.c_void_p, C.c_int]
    R.ref_get_size.argtypes = [C.c_void_p]
    R.ref_get_size.restype = C.c_float
    R.ref_create_space_points.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_uint]
    R.ref_clean_particles.argtypes = [C.c_void_p]
    R.ref_compute_cube_size.argtypes = [C.c_void_p]
    R.ref_create_octree.argtypes = [C.c_void_p]
    R.ref_tick.argtypes = [C.c_void_p, C.c_int]
    R.ref_tick.restype = C.c_double
    R.ref_debug_draws.argtypes = [C.c_void_p, _fp, C.c_int]
    R.ref_debug_draws.restype = C.c_int


def have_ref() -> bool:
    build()
    return os.path.exists(_REF)


def ref():
    global _ref
    if _ref is None:
        build()
        if not os.path.exists(_REF):
            raise RuntimeError("oracle/_ref/liboracle_ref.so is missing (build it where /root/reference exists)")
        R = C.CDLL(_REF)
        _bind_actor(R)
        R.ref_compute_forces.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int]
        R.ref_compute_forces.restype = C.c_double
        R.ref_integrate.argtypes = [C.c_void_p]
        R.ref_root.argtypes = [C.c_void_p, _fp, _fp, _fp, _fp]
        R.ref_root.restype = C.c_int
        R.ref_max_threads.restype = C.c_int
        _ref = R
    return _ref


# ----------------------------------------------------------------------------- layout helpers
def to_aos(posm: np.ndarray, vel: np.ndarray | None = None, acc: np.ndarray | None = None) -> np.ndarray:
    n = posm.shape[0]
    p = np.zeros(n, dtype=PARTICLE_DTYPE)
    p["Mass"] = posm[:, 3]
    p["Position"] = posm[:, :3]
    if vel is not None:
        p["Velocity"] = vel[:, :3]
    if acc is not None:
        p["Acceleration"] = acc[:, :3]
    return p


def from_aos(p: np.ndarray):
    n = p.shape[0]
    posm = np.zeros((n, 4), np.float32)
    vel = np.zeros((n, 4), np.float32)
    acc = np.zeros((n, 4), np.float32)
    posm[:, :3] = p["Position"]
    posm[:, 3] = p["Mass"]
    vel[:, :3] = p["Velocity"]
    acc[:, :3] = p["Acceleration"]
    return posm, vel, acc


# ----------------------------------------------------------------------------- the real reference
_ADAPTER = os.path.join(os.path.dirname(_HERE), "integration", "ue4", "libue4_adapter.so")
_adapter = None


def have_adapter() -> bool:
    return os.path.exists(_ADAPTER)


def adapter():
    """The same C driver (oracle/ref_wrap.cpp) compiled against the GPU adapter actor integration/ue4/OctreeSearch.{h,cpp},
    which forwards to libnbody_b200.so (Makefile target integration/ue4/libue4_adapter.so)."""
    global _adapter
    if _adapter is None:
        if not os.path.exists(_ADAPTER):
            raise RuntimeError(f"{_ADAPTER} is missing: run `make`")
        A = C.CDLL(_ADAPTER)
        _bind_actor(A)
        A.ref_adapter_config.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int]
        _adapter = A
    return _adapter


class RefSim:
    """The reference's AOctreeSearch, verbatim (OctreeSearch.h:111-149) - or, with use_adapter=True, the adapter actor of
    integration/ue4 driven through the very same calls."""

    def __init__(self, use_adapter: bool = False, theta: float = 1.0, eps: float = 0.0, direct: bool = False, parity: bool = True):
        self._r = adapter() if use_adapter else ref()
        self._h = C.c_void_p(self._r.ref_create())
        if use_adapter:
            self._r.ref_adapter_config(self._h, theta, eps, int(direct), int(parity))

    def close(self):
        if self._h:
            self._r.ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # public members
    @property
    def PhDeltaTime(self):
        return self._r.ref_get_dt(self._h)

    @PhDeltaTime.setter
    def PhDeltaTime(self, v):
        self._r.ref_set_dt(self._h, float(v))

    @property
    def Size(self):
        return self._r.ref_get_size(self._h)

    def set_show_octree(self, on: bool):
        self._r.ref_set_show_octree(self._h, int(on))

    def Num(self):
        return self._r.ref_num(self._h)

    def SetParticles(self, aos: np.ndarray):
        aos = np.ascontiguousarray(aos, dtype=PARTICLE_DTYPE)
        self._r.ref_set_particles(self._h, aos.ctypes.data_as(C.c_void_p), aos.shape[0])

    def Particles(self) -> np.ndarray:
        out = np.zeros(self.Num(), dtype=PARTICLE_DTYPE)
        self._r.ref_get_particles(self._h, out.ctypes.data_as(C.c_void_p))
        return out

    # verbs
    def CreateSpacePoints(self, n: int, size: float = 200.0, seed: int = 1234):
        self._r.ref_create_space_points(self._h, n, float(size), seed)

    def CleanParticles(self):
        self._r.ref_clean_particles(self._h)

    def ComputeCubeSize(self):
        self._r.ref_compute_cube_size(self._h)

    def CreateOctree(self):
        self._r.ref_create_octree(self._h)

    def Tick(self, nsteps: int = 1) -> float:
        return self._r.ref_tick(self._h, nsteps)

    def ComputeForces(self, theta: float, i0: int = 0, i1: int | None = None, nthreads: int = 1) -> float:
        if i1 is None:
            i1 = self.Num()
        t = self._r.ref_compute_forces(self._h, float(theta), i0, i1, nthreads)
        if t < 0:
            raise RuntimeError("no tree: call CreateOctree() first")
        return t

    def Integrate(self):
        self._r.ref_integrate(self._h)

    def Root(self):
        o = np.zeros(3, np.float32)
        c = np.zeros(3, np.float32)
        h = C.c_float()
        m = C.c_float()
        ok = self._r.ref_root(self._h, _f(o), C.byref(h), C.byref(m), _f(c))
        return None if not ok else dict(origin=o, half=h.value, mass=m.value, com=c)

    def DebugDraws(self) -> np.ndarray:
        n = self._r.ref_debug_draws(self._h, None, 0)
        out = np.zeros((max(n, 1), 7), np.float32)
        self._r.ref_debug_draws(self._h, _f(out), n)
        return out[:n]


def ref_max_threads() -> int:
    return ref().ref_max_threads()


def host_threads() -> int:
    """Cores this process may run on. Pass it explicitly to the *_forces / direct_* calls: torchrun exports
    OMP_NUM_THREADS=1, which is what omp_get_max_threads() (ref_max_threads / max_threads) would report."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ----------------------------------------------------------------------------- the C restatement
def _c4(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4
    return a


def max_threads() -> int:
    return lib().oracle_max_threads()


def direct_f64(posm, G=1e4, eps=0.0, i0=0, i1=None, nthreads=0, return_time=False):
    posm = _c4(posm)
    n = posm.shape[0]
    i1 = n if i1 is None else i1
    nthreads = nthreads or max_threads()
    out = np.zeros((i1 - i0, 3), np.float64)
    t = lib().oracle_direct_f64(n, _f(posm), G, eps, i0, i1, out.ctypes.data_as(_dp), nthreads)
    return (out, t) if return_time else out


def direct_f32(posm, G=1e4, eps=0.0, i0=0, i1=None, nthreads=0, return_time=False):
    posm = _c4(posm)
    n = posm.shape[0]
    i1 = n if i1 is None else i1
    nthreads = nthreads or max_threads()
    out = np.zeros((i1 - i0, 4), np.float32)
    t = lib().oracle_direct_f32(n, _f(posm), G, eps, i0, i1, _f(out), nthreads)
    return (out, t) if return_time else out


def kick_drift(posm, vel, acc, dt):
    """In place."""
    assert posm.dtype == np.float32 and vel.dtype == np.float32 and posm.flags.c_contiguous and vel.flags.c_contiguous
    acc = _c4(acc)
    lib().oracle_kick_drift(posm.shape[0], _f(posm), _f(vel), _f(acc), dt)


def cube_size(posm) -> float:
    posm = _c4(posm)
    return lib().oracle_cube_size(posm.shape[0], _f(posm))


def energy(posm, vel, G=1e4, eps=0.0, nthreads=0):
    posm = _c4(posm)
    vel = _c4(vel)
    ke = C.c_double()
    pe = C.c_double()
    lib().oracle_energy(posm.shape[0], _f(posm), _f(vel), G, eps, C.byref(ke), C.byref(pe), nthreads or max_threads())
    return ke.value, pe.value


class BHTree:
    """Restated reference octree (OctreeSearch.h:21-109) over a fixed body set."""

    def __init__(self, posm, origin=(0, 0, 0), half=None, G=1e4, eps=0.0):
        self.posm = _c4(posm).copy()
        self.n = self.posm.shape[0]
        if half is None:
            half = cube_size(self.posm)
        o = np.asarray(origin, np.float32).copy()
        st = C.c_int()
        self._h = C.c_void_p(lib().oracle_bh_build(self.n, _f(self.posm), _f(o), half, G, eps, C.byref(st)))
        self.status = st.value

    def close(self):
        if self._h:
            lib().oracle_bh_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def num_nodes(self):
        return lib().oracle_bh_num_nodes(self._h)

    def root(self):
        m = C.c_float()
        c = np.zeros(3, np.float32)
        lib().oracle_bh_root(self._h, C.byref(m), _f(c))
        return m.value, c

    def forces(self, theta, i0=0, i1=None, nthreads=0, return_count=False):
        i1 = self.n if i1 is None else i1
        out = np.zeros((i1 - i0, 4), np.float32)
        cnt = lib().oracle_bh_forces(self._h, theta, i0, i1, _f(out), nthreads or max_threads())
        return (out, cnt) if return_count else out

    def leaf_boxes(self):
        out = np.zeros((self.n, 8), np.float32)
        k = lib().oracle_bh_leaf_boxes(self._h, _f(out), self.n)
        return out[:k]


def tick(posm, vel, acc, dt, theta=1.0, G=1e4, eps=0.0, method=1, prev_com=None, nthreads=0):
    """One reference step in place (OctreeSearch.cpp:21-34). prev_com (float32[3]) carries the root origin."""
    if prev_com is None:
        prev_com = np.zeros(3, np.float32)
    st = lib().oracle_tick(posm.shape[0], _f(posm), _f(vel), _f(acc), dt, theta, G, eps, method, _f(prev_com), nthreads or max_threads())
    if st:
        raise RuntimeError("oracle tree hit the depth cap (coincident bodies)")
    return prev_com
