"""ctypes binding of include/nbody.h and the Python mirror of the reference's simulation actor.

``OctreeSearch`` keeps the names, argument meaning and (lack of) error behaviour of ``AOctreeSearch``
(/root/reference/Source/NBody/OctreeSearch.h:111-149, OctreeSearch.cpp:1-97) so parity tests read like code written
against the reference:

    sim = OctreeSearch(method=METHOD_BARNES_HUT)      # AOctreeSearch()            OctreeSearch.cpp:8
    sim.CreateSpacePoints(2000, 1000)                  # CreateSpacePoints(N, Size) OctreeSearch.cpp:58
    sim.PhDeltaTime = 0.01                             # UPROPERTY PhDeltaTime      OctreeSearch.h:126
    sim.Tick()                                         # Tick(DeltaSeconds)         OctreeSearch.cpp:21
    p = sim.Particles                                  # TArray<FParticle>          OctreeSearch.h:118

All arithmetic happens in the CUDA library; this file only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

METHOD_DIRECT = 0
METHOD_BARNES_HUT = 1

PARAM_G, PARAM_EPS, PARAM_THETA, PARAM_PH_DELTA_TIME, PARAM_METHOD, PARAM_LEAF_SIZE, PARAM_REFERENCE_ROOT, \
    PARAM_SHOW_OCTREE, PARAM_INITIALIZED, PARAM_MAC, PARAM_GROUP_SIZE, PARAM_GROUP_PACK = range(12)

# FParticle, OctreeSearch.h:9-18 (40 bytes)
PARTICLE_DTYPE = np.dtype([("Mass", "<f4"), ("Position", "<f4", 3), ("Velocity", "<f4", 3), ("Acceleration", "<f4", 3)])


class NBodyError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"nbody error {code}: {msg}")
        self.code = code


class _Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("method", C.c_int32), ("G", C.c_float), ("eps", C.c_float),
                ("theta", C.c_float), ("ph_delta_time", C.c_float), ("device", C.c_int32), ("rank", C.c_int32),
                ("world", C.c_int32), ("leaf_size", C.c_int32), ("reference_root", C.c_int32),
                ("mac", C.c_int32), ("group_size", C.c_int32), ("group_pack", C.c_int32), ("bh_exchange", C.c_int32), ("reserved", C.c_int32 * 1), ("nccl_unique_id", C.c_uint8 * 128), ("stream", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("method", C.c_int32), ("n_global", C.c_int64), ("n_local", C.c_int64),
                ("steps", C.c_int64), ("interactions", C.c_double), ("kernel_launches", C.c_double),
                ("ms_last_call", C.c_float), ("ms_force", C.c_float), ("ms_build", C.c_float),
                ("ms_integrate", C.c_float), ("ms_comm", C.c_float), ("cube_size", C.c_float), ("jsplit", C.c_int32),
                ("i_per_thread", C.c_int32), ("tree_nodes", C.c_int32), ("tree_depth", C.c_int32),
                ("root_com", C.c_float * 3), ("root_mass", C.c_float), ("walk_groups", C.c_int32), ("let_points", C.c_int32),
                ("equal_mass", C.c_int32), ("sort_passes", C.c_int32), ("migrated", C.c_int32),
                ("ms_let_migrate", C.c_float), ("ms_let_plan", C.c_float), ("ms_let_walk_local", C.c_float), ("ms_let_import", C.c_float),
                ("ms_let_walk_let", C.c_float)]

    def as_dict(self):
        d = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            d[k] = list(v) if hasattr(v, "__len__") else v
        return d


EXPORTS = [
    "nbody_abi_version", "nbody_last_error", "nbody_config_default", "nbody_create", "nbody_destroy",
    "nbody_create_space_points", "nbody_set_particles_aos", "nbody_set_bodies", "nbody_clean_particles",
    "nbody_compute_cube_size", "nbody_create_octree", "nbody_tick", "nbody_step", "nbody_step_async",
    "nbody_synchronize", "nbody_get_particles_aos", "nbody_get_positions", "nbody_get_velocities",
    "nbody_get_accelerations", "nbody_get_local_ids", "nbody_set_param", "nbody_get_param", "nbody_energy",
    "nbody_stats_get", "nbody_octree_boxes", "nbody_device_ptrs", "nbody_comm_unique_id", "nbody_measure_fp32_peak",
    "nbody_octree_nodes", "nbody_sort_pairs_u64", "nbody_save_snapshot", "nbody_load_snapshot", "nbody_comm_loopback_id",
]

_lib = None


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnbody_b200.so")


def load_library():
    """Load libnbody_b200.so. Raises (never falls back) when the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise NBodyError(-2, f"{p} is missing: build it with `make` (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(p)
    vp, i64, f32p = C.c_void_p, C.c_int64, C.POINTER(C.c_float)
    L.nbody_abi_version.restype = C.c_int
    L.nbody_last_error.restype = C.c_char_p
    L.nbody_config_default.argtypes = [C.POINTER(_Config)]
    L.nbody_create.argtypes = [C.POINTER(vp), C.POINTER(_Config)]
    L.nbody_destroy.argtypes = [vp]
    L.nbody_destroy.restype = None
    L.nbody_create_space_points.argtypes = [vp, i64, C.c_float, C.c_uint64]
    L.nbody_set_particles_aos.argtypes = [vp, vp, i64, C.c_size_t]
    L.nbody_set_bodies.argtypes = [vp, vp, vp, i64]
    L.nbody_clean_particles.argtypes = [vp]
    L.nbody_compute_cube_size.argtypes = [vp, f32p]
    L.nbody_create_octree.argtypes = [vp]
    L.nbody_tick.argtypes = [vp]
    L.nbody_step.argtypes = [vp, C.c_float, C.c_int32]
    L.nbody_step_async.argtypes = [vp, C.c_float, C.c_int32]
    L.nbody_synchronize.argtypes = [vp]
    L.nbody_get_particles_aos.argtypes = [vp, vp, i64, C.c_size_t]
    for f in (L.nbody_get_positions, L.nbody_get_velocities, L.nbody_get_accelerations):
        f.argtypes = [vp, vp, i64]
    L.nbody_get_local_ids.argtypes = [vp, vp, i64, C.POINTER(i64)]
    L.nbody_set_param.argtypes = [vp, C.c_int32, C.c_double]
    L.nbody_get_param.argtypes = [vp, C.c_int32, C.POINTER(C.c_double)]
    L.nbody_energy.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.nbody_stats_get.argtypes = [vp, C.POINTER(Stats)]
    L.nbody_octree_boxes.argtypes = [vp, vp, i64, C.POINTER(i64)]
    L.nbody_device_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.nbody_comm_unique_id.argtypes = [vp]
    L.nbody_comm_loopback_id.argtypes = [vp]
    L.nbody_measure_fp32_peak.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.nbody_save_snapshot.argtypes = [vp, C.c_char_p]
    L.nbody_load_snapshot.argtypes = [vp, C.c_char_p]
    L.nbody_octree_nodes.argtypes = [vp, vp, vp, vp, vp, i64, i64, C.POINTER(i64)]
    L.nbody_sort_pairs_u64.argtypes = [C.c_int32, vp, i64, C.c_int32, vp, vp, C.POINTER(C.c_float)]
    if L.nbody_abi_version() != 1:
        raise NBodyError(-1, "ABI version mismatch")
    _lib = L
    return L


def _check(code: int):
    if code != 0:
        raise NBodyError(code, (load_library().nbody_last_error() or b"").decode())


def measure_fp32_peak(device: int = 0):
    """(TFLOP/s, SM MHz) of an FFMA-chain microbenchmark - the measured FP32 roofline denominator."""
    t, m = C.c_double(), C.c_double()
    _check(load_library().nbody_measure_fp32_peak(device, C.byref(t), C.byref(m)))
    return t.value, m.value


def sort_pairs_u64(keys: np.ndarray, key_bits: int = 64, device: int = 0, timed: bool = False):
    """K5 on its own: (sorted keys, stable permutation[, device ms])."""
    keys = np.ascontiguousarray(keys, np.uint64)
    out_k = np.empty_like(keys)
    out_i = np.empty(keys.shape[0], np.uint32)
    ms = C.c_float()
    _check(load_library().nbody_sort_pairs_u64(device, _ptr(keys), keys.shape[0], key_bits, _ptr(out_k), _ptr(out_i),
                                               C.byref(ms) if timed else None))
    return (out_k, out_i, ms.value) if timed else (out_k, out_i)


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    _check(load_library().nbody_comm_unique_id(buf))
    return bytes(buf)


def comm_loopback_id() -> bytes:
    """Id of a fresh in-process (loop-back) group: create `world` handles with it, one host thread each."""
    buf = (C.c_uint8 * 128)()
    _check(load_library().nbody_comm_loopback_id(buf))
    return bytes(buf)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OctreeSearch:
    """Mirror of ``AOctreeSearch`` (OctreeSearch.h:111-149) on top of the CUDA library.

    Differences forced by device residency: ``Particles`` is a property that copies from the GPU (and a setter
    that copies to it); the constants the reference bakes in (G = 1e4, Theta = 1.0, eps = 0) are constructor
    arguments whose defaults are the reference's values.
    """

    def __init__(self, method: int = METHOD_BARNES_HUT, G: float = 1e4, eps: float = 0.0, theta: float = 1.0,
                 PhDeltaTime: float = 0.01, device: int = 0, rank: int = 0, world: int = 1,
                 nccl_unique_id: bytes | None = None, leaf_size: int = 16, reference_root: bool = False,
                 mac: int = 0, group_size: int = 32, group_pack: int = 2, bh_exchange: int = -1, stream: int | None = None):
        self._L = load_library()
        cfg = _Config()
        _check(self._L.nbody_config_default(C.byref(cfg)))
        cfg.method, cfg.G, cfg.eps, cfg.theta, cfg.ph_delta_time = method, G, eps, theta, PhDeltaTime
        cfg.device, cfg.rank, cfg.world, cfg.leaf_size, cfg.reference_root = device, rank, world, leaf_size, int(reference_root)
        cfg.mac, cfg.group_size, cfg.group_pack, cfg.bh_exchange = mac, group_size, group_pack, bh_exchange
        if world > 1:
            if nccl_unique_id is None or len(nccl_unique_id) != 128:
                raise NBodyError(-1, "world > 1 needs the 128-byte nccl_unique_id broadcast from rank 0 "
                                     "(or 128 zero bytes for an emulated rank)")
            C.memmove(cfg.nccl_unique_id, nccl_unique_id, 128)
        cfg.stream = stream
        self._h = C.c_void_p()
        _check(self._L.nbody_create(C.byref(self._h), C.byref(cfg)))
        self.method = method
        self.rank, self.world = rank, world

    # ---- lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._L.nbody_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- reference verbs
    def CreateSpacePoints(self, N: int, Size: float = 200.0, seed: int = 1234):
        """OctreeSearch.cpp:58-72 (default Size = 200 as in OctreeSearch.h:142)."""
        _check(self._L.nbody_create_space_points(self._h, N, Size, seed))

    def CleanParticles(self):
        """OctreeSearch.cpp:91-97."""
        _check(self._L.nbody_clean_particles(self._h))

    def ComputeCubeSize(self) -> float:
        """OctreeSearch.cpp:47-56; silently does nothing when not Initialized (cpp:49)."""
        if not self.Initialized:
            return self.Size
        out = C.c_float()
        _check(self._L.nbody_compute_cube_size(self._h, C.byref(out)))
        return out.value

    def CreateOctree(self):
        """OctreeSearch.cpp:74-89: evaluate accelerations at the current positions; silently nothing when not
        Initialized (cpp:76)."""
        if not self.Initialized:
            return
        _check(self._L.nbody_create_octree(self._h))

    def Tick(self, DeltaSeconds: float = 0.0):
        """OctreeSearch.cpp:21-34. DeltaSeconds is ignored, as in the reference (physics uses PhDeltaTime)."""
        _check(self._L.nbody_tick(self._h))

    # ---- extensions over the reference
    def Step(self, dt: float, nsteps: int = 1):
        _check(self._L.nbody_step(self._h, dt, nsteps))

    def StepAsync(self, dt: float, nsteps: int = 1):
        _check(self._L.nbody_step_async(self._h, dt, nsteps))

    def Synchronize(self):
        _check(self._L.nbody_synchronize(self._h))

    def SetBodies(self, posm: np.ndarray, vel: np.ndarray | None = None):
        posm = np.ascontiguousarray(posm, np.float32)
        if posm.ndim != 2 or posm.shape[1] != 4:
            raise NBodyError(-1, "posm must be [N, 4] = (x, y, z, mass)")
        v = None
        if vel is not None:
            v = np.ascontiguousarray(vel, np.float32)
            if v.shape != posm.shape:
                raise NBodyError(-1, "vel must be [N, 4]")
        _check(self._L.nbody_set_bodies(self._h, _ptr(posm), _ptr(v) if v is not None else None, posm.shape[0]))

    def SetParticlesRaw(self, ptr: int, n: int, stride: int = 40):
        """Host pointer variant (e.g. a pinned torch tensor's data_ptr())."""
        _check(self._L.nbody_set_particles_aos(self._h, C.c_void_p(ptr), n, stride))

    def GetParticlesRaw(self, ptr: int, n: int, stride: int = 40):
        _check(self._L.nbody_get_particles_aos(self._h, C.c_void_p(ptr), n, stride))

    def _get(self, fn):
        out = np.zeros((self.Num(), 4), np.float32)
        _check(fn(self._h, _ptr(out), out.shape[0]))
        return out

    def Positions(self):
        return self._get(self._L.nbody_get_positions)

    def Velocities(self):
        return self._get(self._L.nbody_get_velocities)

    def Accelerations(self):
        return self._get(self._L.nbody_get_accelerations)

    def LocalIds(self) -> np.ndarray:
        n = C.c_int64()
        _check(self._L.nbody_get_local_ids(self._h, None, 0, C.byref(n)))
        ids = np.zeros(max(n.value, 1), np.int64)
        _check(self._L.nbody_get_local_ids(self._h, _ptr(ids), ids.shape[0], C.byref(n)))
        return ids[:n.value]

    def Energy(self):
        ke, pe = C.c_double(), C.c_double()
        _check(self._L.nbody_energy(self._h, C.byref(ke), C.byref(pe)))
        return ke.value, pe.value

    def Stats(self) -> dict:
        st = Stats()
        _check(self._L.nbody_stats_get(self._h, C.byref(st)))
        return st.as_dict()

    def OctreeBoxes(self) -> np.ndarray:
        """Read-back of what DrawOctreeBoxes would draw (OctreeSearch.cpp:36-45): [k, 7] = centre, half extents, count."""
        n = C.c_int64()
        cap = max(self.Num(), 1)
        out = np.zeros((cap, 7), np.float32)
        _check(self._L.nbody_octree_boxes(self._h, _ptr(out), cap, C.byref(n)))
        return out[:n.value]

    def SaveSnapshot(self, path: str):
        _check(self._L.nbody_save_snapshot(self._h, os.fsencode(path)))

    def LoadSnapshot(self, path: str):
        _check(self._L.nbody_load_snapshot(self._h, os.fsencode(path)))

    def OctreeNodes(self) -> dict:
        """The last Barnes-Hut build: per-node centre of mass + mass, (first, count, level | 256*leaf, parent), body range,
        and the sorted Morton keys (what walking ``ParticleOctree`` through its getters gives in the reference)."""
        n = C.c_int64()
        _check(self._L.nbody_octree_nodes(self._h, None, None, None, None, 0, 0, C.byref(n)))
        k, nb = n.value, self.Num()
        com, meta = np.zeros((k, 4), np.float32), np.zeros((k, 4), np.int32)
        rng, keys = np.zeros((k, 2), np.int32), np.zeros(nb, np.uint64)
        _check(self._L.nbody_octree_nodes(self._h, _ptr(com), _ptr(meta), _ptr(rng), _ptr(keys), k, nb, C.byref(n)))
        return {"com": com, "first": meta[:, 0], "count": meta[:, 1], "level": meta[:, 2] & 255,
                "leaf": (meta[:, 2] & 256) != 0, "parent": meta[:, 3], "range": rng, "keys": keys}

    def DevicePtrs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(self._L.nbody_device_ptrs(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def Num(self) -> int:
        st = Stats()
        _check(self._L.nbody_stats_get(self._h, C.byref(st)))
        return int(st.n_global)

    # ---- reference public members
    @property
    def Particles(self) -> np.ndarray:
        """TArray<FParticle> Particles (OctreeSearch.h:118), copied from the device (this rank's share filled)."""
        n = self.Num()
        out = np.zeros(n, PARTICLE_DTYPE)
        if n:
            _check(self._L.nbody_get_particles_aos(self._h, _ptr(out), n, 40))
        return out

    @Particles.setter
    def Particles(self, p: np.ndarray):
        p = np.ascontiguousarray(p, PARTICLE_DTYPE)
        _check(self._L.nbody_set_particles_aos(self._h, _ptr(p), p.shape[0], 40))

    def _getp(self, which):
        v = C.c_double()
        _check(self._L.nbody_get_param(self._h, which, C.byref(v)))
        return v.value

    def _setp(self, which, v):
        _check(self._L.nbody_set_param(self._h, which, float(v)))

    PhDeltaTime = property(lambda s: s._getp(PARAM_PH_DELTA_TIME), lambda s, v: s._setp(PARAM_PH_DELTA_TIME, v))
    ShowOctree = property(lambda s: bool(s._getp(PARAM_SHOW_OCTREE)), lambda s, v: s._setp(PARAM_SHOW_OCTREE, v))
    Theta = property(lambda s: s._getp(PARAM_THETA), lambda s, v: s._setp(PARAM_THETA, v))
    Eps = property(lambda s: s._getp(PARAM_EPS), lambda s, v: s._setp(PARAM_EPS, v))
    G = property(lambda s: s._getp(PARAM_G), lambda s, v: s._setp(PARAM_G, v))
    Initialized = property(lambda s: bool(s._getp(PARAM_INITIALIZED)))
    Mac = property(lambda s: int(s._getp(PARAM_MAC)), lambda s, v: s._setp(PARAM_MAC, v))
    GroupSize = property(lambda s: int(s._getp(PARAM_GROUP_SIZE)), lambda s, v: s._setp(PARAM_GROUP_SIZE, v))
    GroupPack = property(lambda s: int(s._getp(PARAM_GROUP_PACK)), lambda s, v: s._setp(PARAM_GROUP_PACK, v))
    LeafSize = property(lambda s: int(s._getp(PARAM_LEAF_SIZE)), lambda s, v: s._setp(PARAM_LEAF_SIZE, v))

    @property
    def Size(self) -> float:
        return float(self.Stats()["cube_size"])


def to_particles(posm: np.ndarray, vel: np.ndarray | None = None) -> np.ndarray:
    """float4 SoA -> FParticle AoS (harness helper)."""
    p = np.zeros(posm.shape[0], PARTICLE_DTYPE)
    p["Mass"] = posm[:, 3]
    p["Position"] = posm[:, :3]
    if vel is not None:
        p["Velocity"] = vel[:, :3]
    return p
