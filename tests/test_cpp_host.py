"""The C++ host mirror of the reference's actor (include/nbody.hpp) and the example program written against it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "octree_search")


def _build():
    """g++ only (never `make`: the product library must not be rebuilt under a process that has it loaded)."""
    srcs = [os.path.join(ROOT, "examples", "octree_search.cpp"), os.path.join(ROOT, "include", "nbody.hpp"),
            os.path.join(ROOT, "include", "nbody.h")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(p) for p in srcs):
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), srcs[0], "-o", EXE,
                               "-L", os.path.join(ROOT, "parallelnbody_b200"), "-lnbody_b200",
                               "-Wl,-rpath,$ORIGIN/../parallelnbody_b200"])
    assert os.path.exists(EXE)


def test_example_builds_against_the_c_abi_and_fails_loudly_without_gpu():
    _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: see the gpu-marked test")
    r = subprocess.run([EXE, "500"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr     # the wrapper throws, it never computes on the host


def test_header_only_wrapper_keeps_the_actor_surface():
    src = open(os.path.join(ROOT, "include", "nbody.hpp")).read()
    for name in ("CreateSpacePoints", "ComputeCubeSize", "CreateOctree", "Tick", "CleanParticles", "Particles", "Size",
                 "Initialized", "ShowOctree", "PhDeltaTime"):      # OctreeSearch.h:116-148
        assert name in src
    assert "#include <cuda" not in src and "Engine.h" not in src and '#include "nbody.h"' in src


@pytest.mark.gpu
def test_example_runs_the_reference_workflow_on_the_gpu():
    _build()
    r = subprocess.run([EXE, "2000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("OK") and "steps=11" in r.stdout
