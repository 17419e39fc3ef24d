"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/nbody.h declares,
its structs have the layout the ctypes mirror assumes, and - without a GPU - the product fails loudly instead of
falling back to any CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nbody.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nbody_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported():
    import parallelnbody_b200 as P
    from parallelnbody_b200 import api
    L = P.load_library()
    names = _declared()
    assert len(names) >= 28
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/nbody.h but not exported"
    assert set(api.EXPORTS) == set(names)
    assert L.nbody_abi_version() == 1


def test_struct_layouts_match_header():
    from parallelnbody_b200 import api
    L = api.load_library()
    cfg = api._Config()
    assert L.nbody_config_default(C.byref(cfg)) == 0
    assert cfg.struct_size == C.sizeof(api._Config)
    # the reference's shipped constants (OctreeSearch.h:104, OctreeSearch.cpp:8,85)
    assert cfg.G == 1e4 and cfg.eps == 0.0 and cfg.theta == 1.0 and abs(cfg.ph_delta_time - 0.01) < 1e-9
    assert cfg.method == api.METHOD_BARNES_HUT and cfg.world == 1 and cfg.rank == 0
    assert api.PARTICLE_DTYPE.itemsize == 40   # FParticle, OctreeSearch.h:9-18
    assert [api.PARTICLE_DTYPE.fields[k][1] for k in ("Mass", "Position", "Velocity", "Acceleration")] == [0, 4, 16, 28]


def test_invalid_arguments_return_status_not_crash():
    from parallelnbody_b200 import api
    L = api.load_library()
    assert L.nbody_config_default(None) == -1
    assert b"NULL" in L.nbody_last_error()
    h = C.c_void_p()
    cfg = api._Config()
    L.nbody_config_default(C.byref(cfg))
    cfg.struct_size = 3
    assert L.nbody_create(C.byref(h), C.byref(cfg)) == -1 and not h.value
    assert L.nbody_tick(None) == -1
    assert L.nbody_step(None, 0.1, 1) == -1
    L.nbody_destroy(None)   # allowed


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path is for CPU-only hosts")
    import parallelnbody_b200 as P
    with pytest.raises(P.NBodyError) as e:
        P.OctreeSearch(method=P.METHOD_DIRECT)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(P.NBodyError):
        P.measure_fp32_peak(0)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "parallelnbody_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("CPU oracle", ""), f"{f} references oracle/"


def test_ic_generators_are_seeded_and_centred():
    from parallelnbody_b200 import ic
    p1, v1 = ic.plummer(2000, seed=3)
    p2, v2 = ic.plummer(2000, seed=3)
    assert np.array_equal(p1, p2) and np.array_equal(v1, v2)
    assert np.abs(p1[:, :3].mean(0)).max() < 1e-6 and np.abs(v1[:, :3].mean(0)).max() < 1e-6
    assert np.isclose(p1[:, 3].sum(), 1e-4, rtol=1e-5)
    # virial equilibrium of the sampled Plummer model: 2T/|W| ~ 1 (G*M = 1, a = 1)
    from oracle import oracle as O
    ke, pe = O.energy(p1, v1, 1e4, 0.0)
    assert 0.85 < 2 * ke / abs(pe) < 1.15
    u, _ = ic.uniform_cube(1000, seed=1)
    assert u[:, :3].min() >= -1 and u[:, :3].max() < 1
    g, gv = ic.two_galaxies(1000, seed=1)
    assert g[:500, 0].mean() > 3 and g[500:, 0].mean() < -3
    s, sv = ic.reference_slab(500, 1000.0, seed=1)
    assert s[0, 3] == 5000 and np.all(s[0, :3] == 0) and np.abs(s[:, 2]).max() <= 100
