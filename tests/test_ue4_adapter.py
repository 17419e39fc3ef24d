"""The Unreal-side adapter actor (integration/ue4/OctreeSearch.{h,cpp}): `AOctreeSearch` with the reference's public surface
(OctreeSearch.h:111-149) forwarding to the C ABI. It is compiled against the stand-in engine header the oracle uses and
driven by the same C wrapper (oracle/ref_wrap.cpp) as the CPU actor built from the reference sources, so the two run
side by side through identical calls: reference-side API -> C ABI -> GPU."""
import os
import subprocess

import numpy as np
import pytest

from conftest import rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_adapter_sources_keep_the_actor_surface_and_compile():
    hdr = open(os.path.join(ROOT, "integration", "ue4", "OctreeSearch.h")).read()
    for name in ("class NBODY_API AOctreeSearch : public AActor", "TArray<FParticle> Particles", "float Size", "bool Initialized",
                 "bool ShowOctree", "float PhDeltaTime", "void CreateSpacePoints(int32 N, float Size = 200)", "void CreateOctree()",
                 "void CleanParticles()", "void ComputeCubeSize()", "virtual void Tick(float DeltaSeconds) override",
                 "UFUNCTION(BlueprintCallable, Category = \"Octree\")", "UPROPERTY(BlueprintReadWrite)"):
        assert name in hdr, name
    src = open(os.path.join(ROOT, "integration", "ue4", "OctreeSearch.cpp")).read()
    assert '#include "nbody.h"' in src and "cuda" not in src.lower().replace("libnbody", "")
    for sym in ("nbody_tick", "nbody_set_particles_aos", "nbody_get_particles_aos", "nbody_octree_boxes", "nbody_create_space_points",
                "nbody_clean_particles", "nbody_compute_cube_size", "nbody_create_octree"):
        assert sym in src
    # syntax + type check against the stand-in engine header (no link: the CUDA library is not needed for this)
    subprocess.check_call(["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "integration", "ue4"),
                           "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "integration", "ue4", "OctreeSearch.cpp")])


@pytest.mark.gpu
def test_adapter_actor_runs_beside_the_cpu_actor(oracle):
    """CreateSpacePoints-style workflow through AOctreeSearch's own names on both actors: inject the same Particles, Tick x5
    with ShowOctree, compare bodies, Size and the debug draws, then CleanParticles. Parity configuration (one-body leaves,
    reference root cube, per-body walk, Theta = 1.0 as shipped)."""
    if not oracle.have_adapter():
        pytest.skip("integration/ue4/libue4_adapter.so not built")
    g = np.load(os.path.join(GOLD, "slab_300_tick5.npz"))
    gpu = oracle.RefSim(use_adapter=True, theta=1.0, eps=0.0, parity=True)
    cpu = oracle.RefSim() if oracle.have_ref() else None
    try:
        for a in (gpu, cpu):
            if a is None:
                continue
            a.PhDeltaTime = float(g["dt"])
            a.set_show_octree(True)
            a.SetParticles(g["particles0"])
        for k in range(int(g["steps"])):
            gpu.Tick(1)
            p = gpu.Particles()
            want = g["particles"][k]                    # the reference's own Ticks, frozen by tests/golden/make_golden.py
            if cpu is not None:
                cpu.Tick(1)
                live = cpu.Particles()
                assert np.array_equal(live["Position"], want["Position"])      # golden == live reference
                assert gpu.Size == cpu.Size
            assert rel_l2(p["Acceleration"], want["Acceleration"]) <= 5e-5, f"step {k}"
            assert rel_l2(p["Velocity"], want["Velocity"]) <= 1e-5
            assert rel_l2(p["Position"], want["Position"]) <= 1e-6
        d = gpu.DebugDraws()
        n = len(g["particles0"])
        boxes, points = d[d[:, 0] == 0], d[d[:, 0] == 1]
        assert len(points) == n and len(boxes) == n          # one point per body, one box per occupied one-body leaf
        assert np.array_equal(np.sort(points[:, 1:4], axis=0), np.sort(gpu.Particles()["Position"], axis=0))
        if cpu is not None:
            dc = cpu.DebugDraws()
            bc = dc[dc[:, 0] == 0]
            assert len(bc) == len(boxes)
            from scipy.spatial import cKDTree
            dist, j = cKDTree(boxes[:, 1:4].astype(np.float64)).query(bc[:, 1:4].astype(np.float64))
            assert len(np.unique(j)) == len(bc) and dist.max() <= 1e-3 * float(cpu.Size)
            assert np.allclose(boxes[j, 4], bc[:, 4], rtol=1e-6)
        gpu.PhDeltaTime = 0.0                                  # pause: Tick leaves the bodies alone (OctreeSearch.cpp:25)
        before = gpu.Particles()
        gpu.Tick(1)
        assert np.array_equal(before, gpu.Particles())
        gpu.CleanParticles()
        assert gpu.Num() == 0
        gpu.Tick(1)                                            # not Initialized: silently nothing (cpp:49,76)
        gpu.CreateSpacePoints(500, 1000.0, seed=5)
        gpu.PhDeltaTime = 0.01
        gpu.Tick(2)
        q = gpu.Particles()
        assert len(q) == 500 and q["Mass"][0] == 5000.0 and np.all(np.isfinite(q["Position"]))
    finally:
        gpu.close()
        if cpu is not None:
            cpu.close()


@pytest.mark.gpu
def test_adapter_production_settings_follow_the_direct_sum(oracle):
    """The same actor with production settings (group walk / direct kernel) instead of the parity configuration."""
    if not oracle.have_adapter():
        pytest.skip("integration/ue4/libue4_adapter.so not built")
    from parallelnbody_b200 import ic
    posm, vel = ic.plummer(20_000, seed=2)
    exact = oracle.direct_f64(posm, G=1e4, eps=0.01)
    for direct, tol in ((True, 1e-5), (False, 6e-3)):
        a = oracle.RefSim(use_adapter=True, theta=0.25, eps=0.01, direct=direct, parity=False)
        try:
            a.SetParticles(oracle.to_aos(posm, vel))
            a.CreateOctree()
            assert rel_l2(a.Particles()["Acceleration"], exact) <= tol
        finally:
            a.close()
