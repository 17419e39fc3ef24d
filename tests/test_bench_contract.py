"""bench.py's CPU-runnable legs: the reference arm prints the contract's JSON line; without a GPU the product arm refuses."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_prints_contract_line():
    r = _run("--impl", "reference", "--workload", "plummer_4k_direct", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "all-pairs interactions/s" and d["unit"] == "interactions/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["N"] == 4096 and "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_other_ranks_stay_silent():
    r = _run("--impl", "reference", "--workload", "plummer_4k_direct", "--steps", "1", "--warmup", "0", "--gpus", "2",
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_roofline_traffic_comes_from_the_committed_ncu_captures():
    """`roofline.traffic` is read from profiles/traffic.json (written by tools/ncu_traffic.py from `ncu --set full` captures),
    never typed in: both dominant kernels have an entry, and it is at least the kernel's algorithmic bytes."""
    sys.path.insert(0, ROOT)
    import bench
    d = bench.load_traffic("direct_packed_kernel", "plummer_1m_direct")
    w = bench.load_traffic("bh_walk_group_kernel", "plummer_1m_bh")
    n = 1 << 20
    assert d is not None and d >= 16.0 * n            # at least the sources once
    assert w is not None and w >= 16.0 * n            # at least the bodies once
    assert bench.load_traffic("no_such_kernel", "plummer_1m_direct") is None
    src = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert "ncu --set full" in src["plummer_1m_direct"]["direct_packed_kernel__source"]
