"""CPU: the JSON contract of ``bench.py`` (the reference arm runs on host cores only; the GPU arm's keys are
checked on the last line a B200 run left under ``gpurun_out/`` when there is one, and statically in the source)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    o = json.loads(lines[0])
    assert BASE_KEYS <= set(o), BASE_KEYS - set(o)
    assert o["impl"] == "reference" and o["n_gpus"] == 1 and o["higher_is_better"] is True
    assert o["warmup"] >= 3 and o["steps"] == 1 and o["vs_baseline"] is None
    assert o["value"] > 0 and o["unit"] == "samples/s" and "workload" in o["config"]
    cb = o["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == o["value"] and cb["sample"]
    assert o["e2e"] == {"value": o["value"], "unit": o["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert o["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=120,
                       env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_gpu_arm_line_carries_roofline_and_clocks():
    path = os.path.join(ROOT, "gpurun_out", "bench_n1.json")
    if not os.path.exists(path):
        pytest.skip("no B200 bench line in gpurun_out/ (scratch directory)")
    with open(path) as f:
        o = json.loads([l for l in f.read().splitlines() if l.startswith("{")][-1])
    assert BASE_KEYS | {"roofline", "clocks"} <= set(o)
    rf = o["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["frac"] == pytest.approx(rf["achieved"] / rf["peak"])
    assert 0.3 < rf["frac"] <= 1.05 and (rf["traffic"] is None or rf["traffic"] > 0)
    assert o["gpu_launches"] >= o["steps"] and o["e2e"]["h2d_bytes_per_step"] > 0 and o["e2e"]["d2h_bytes_per_step"] > 0
    assert o["e2e"]["value"] < o["value"]  # host buffers, copies inside the timed region
    assert not set(o["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert o["metric"] and o["dtype"] == "f32" and o["scaling"] == "weak"


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], cwd=ROOT,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0  # no CPU fallback behind the GPU arm
    assert not [l for l in r.stdout.splitlines() if l.startswith("{") and '"value"' in l]


def test_near_gpu_cpus_is_best_effort_and_restores_affinity():
    sys.path.insert(0, ROOT)
    try:
        import bench
    finally:
        sys.path.remove(ROOT)
    before = os.sched_getaffinity(0)
    with bench._NearGpuCpus(0) as near:
        inside = os.sched_getaffinity(0)
        assert inside <= before and near.note
    assert os.sched_getaffinity(0) == before
    import torch

    if not torch.cuda.is_available():
        assert near.note.startswith("unchanged")  # no NVML device here: nothing is touched


def test_both_arms_print_the_same_config():
    """The reference arm times a bounded sample of the GPU arm's workload and must say so on the SAME `config`
    (the driver compares the two lines): its config equals the one a B200 run of the GPU arm recorded for the same
    command line (profiles/r02_bench_n1_builder.json, `--gpus 1 --steps 20 --warmup 5`)."""
    with open(os.path.join(ROOT, "profiles", "r02_bench_n1_builder.json")) as f:
        gpu = json.load(f)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps",
                        str(gpu["steps"]), "--warmup", str(gpu["warmup"])], cwd=ROOT, capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    ref = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert ref["config"] == gpu["config"]
    assert ref["metric"] == gpu["metric"] and ref["unit"] == gpu["unit"]
    assert ref["higher_is_better"] == gpu["higher_is_better"] and ref["dtype"] == gpu["dtype"]
    assert "bounded sample" in ref["cpu_baseline"]["sample"]


def test_every_bench_chain_has_an_ahead_of_time_kernel():
    """bench.static_arm_config states `specialized_kernel: true` without loading the library: keep that true."""
    sys.path.insert(0, ROOT)
    try:
        import bench
        from normalizingflownetwork_b200 import build as nfn_build
    finally:
        sys.path.remove(ROOT)
    built = {(d, bool(b), tuple(f)) for d, b, f in nfn_build.SPECIALIZED_CHAINS}
    for name, (ft, d, tb, _, _) in bench.CONFIGS.items():
        if not bench.is_mdn(ft):
            assert (d, bool(tb), tuple(ft)) in built, name
