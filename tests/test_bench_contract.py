"""The JSON line bench.py prints (latest committed run under profiles/) carries every key the
measurement contract names, for both arms."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    assert files, pattern
    with open(files[-1]) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_our_arm_line_has_the_contract_keys():
    d = _latest("r*_bench.json")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["scaling"] == "weak" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] >= 0.9 * r["algorithmic_bytes_per_launch"]
    c = d["cpu_baseline"]
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in c, key
    assert c["kind"] in ("reference", "port")
    e = d["e2e"]
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in e, key
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    for key in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert key in d["clocks"], key
    # round 2: the link rate measured beside e2e, how the steps were launched, and every other
    # BASELINE config device-timed in the same line
    assert e["link_gbs"]["h2d"]["min_rank_gbs"] > 0 and 0 < e["link_frac"] <= 1.05
    assert "launch" in d and "kernels_ms" in d
    cfgs = d["configs"]
    assert set(cfgs) == {"config1_simplebaseline_b64", "config3_udp_384_b2048",
                         "config4_higherhrnet_b64", "config5_sweep_1m"}
    for name, c in cfgs.items():
        for key in ("workload", "units_per_step", "unit", "ms_per_step", "value", "kernels", "l2"):
            assert key in c, (name, key)
        assert abs(c["value"] - c["units_per_step"] / (c["ms_per_step"] * 1e-3)) < 1e-6 * c["value"]
    assert cfgs["config5_sweep_1m"]["units_per_step"] == 1_000_000
    assert cfgs["config5_sweep_1m"]["scaling"] == "strong"
    assert "cpu_match_by_tag" in cfgs["config4_higherhrnet_b64"]


def test_reference_arm_line_has_the_contract_keys():
    d = _latest("r*_bench_reference.json")
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    ours = _latest("r*_bench.json")
    assert d["metric"] == ours["metric"] and d["unit"] == ours["unit"]
    assert d["config"] == ours["config"]          # the same object in both arms
    assert d["warmup"] >= 3


def test_reference_arm_runs_on_rank_0_only():
    """Under torchrun the reference arm prints on rank 0; the other ranks exit 0 without work
    (no CUDA, no process group)."""
    import subprocess
    import sys

    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2", "--steps", "1", "--warmup", "1"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == ""


def test_our_arm_fails_loudly_without_a_gpu():
    import subprocess
    import sys

    import torch

    if torch.cuda.is_available():
        return  # on the GPU box the arm runs; the driver times it
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in r.stderr
    assert r.stdout.strip() == ""   # no line, and certainly no CPU-fallback number
