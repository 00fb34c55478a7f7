"""bench.py's two arms must time the SAME workload: identical buildings (bit for bit) from the neutral generator
(workloads/synth.py), identical hyper-parameters, identical ``config.workload`` strings - and the reference arm must run
without loading the product package."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_both_arms_see_identical_batches():
    import bench
    from oracle import pyg as opyg
    (lb, vb), = bench._make_batches(0, 1, 4, pin=False)
    (olb, ovb), = bench._oracle_batches(1, 4)
    for mine, theirs in ((lb, olb), (vb, ovb)):
        keys = [k for k in theirs.keys() if k not in ("batch", "ptr")]
        assert [k for k in mine.keys() if k not in ("batch", "ptr", "bg_csr")] == keys
        for k in keys + ["batch", "ptr"]:
            a, b = getattr(mine, k), getattr(theirs, k)
            if isinstance(a, torch.Tensor):
                assert a.dtype == b.dtype and torch.equal(a, b), k
            else:
                assert a == b, k
    assert isinstance(ovb, opyg.Batch)


def test_oracle_config_equals_product_config_and_reference_config():
    from building_gan_b200 import Configuration as P
    from oracle.config import Configuration as O
    for name in O.names():
        if name == "DEVICE":
            continue
        assert getattr(P, name) == getattr(O, name), name
    ref = "/root/reference"
    if os.path.isdir(os.path.join(ref, "building_gan", "src")):
        from oracle.make_golden import _import_reference
        config = _import_reference()[0]
        for name in O.names():
            if name != "DEVICE":
                assert getattr(config.Configuration, name) == getattr(O, name), name


def test_reference_arm_does_not_load_the_product():
    code = ("import sys, json; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0'];"
            "import bench; bench.BATCH=2; bench.main(); print('LOADED', 'building_gan_b200' in sys.modules)")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "LOADED False" in out.stdout, out.stdout[-500:]
    import json
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    import bench
    assert line["impl"] == "reference" and line["config"]["workload"] == bench.WORKLOAD
    assert line["product_package_loaded"] is False


def test_committed_bench_lines_are_valid_json_with_the_contract_keys():
    """Every bench line committed under profiles/ parses as ONE JSON object (a captured stdout once carried NCCL's version
    banner in front of it) and the default-arm lines of this round carry the keys the bench contract names."""
    import glob
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(root, "profiles", "*bench_line*.json")) + glob.glob(os.path.join(root, "profiles", "*reference_arm_line*.json")))
    assert files
    for f in files:
        with open(f) as fh:
            d = json.load(fh)
        assert isinstance(d, dict), f
    for name in ("r02c_bench_line.json", "r02c_bench_line_2gpu.json"):
        with open(os.path.join(root, "profiles", name)) as fh:
            d = json.load(fh)
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                    "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
            assert key in d, (name, key)
        assert d["config"]["workload"] and d["e2e"]["h2d_bytes_per_step"] > 0 and d["gpu_launches"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    with open(os.path.join(root, "profiles", "r02c_bench_line.json")) as fh:
        d = json.load(fh)
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]) and d["cpu_baseline"]["kind"] == "port"
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-3
