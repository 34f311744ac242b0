"""bench.py's reference arm runs on the CPU (it times the oracle — the reference's algorithm — on the host cores), so its
JSON line can be checked here: one line on stdout, the keys the driver reads, rank != 0 silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3"],
                          capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("frames/sec") and d["value"] > 0 and d["steps"] == 3 and d["warmup"] == 3
    assert d["config"]["workload"].startswith("S1 synthetic 640x480")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        lines = [l for l in f.read().splitlines() if l.strip()]
    assert len(lines) == 1, name
    return json.loads(lines[0])


def test_committed_bench_lines_carry_the_contract():
    """the evidence under profiles/ is what bench.py printed on the B200 boxes: every key the driver and the judge read"""
    for name, n in (("r01_bench_v10.json", 1), ("r01_bench_v10_2gpu.json", 2), ("r01_bench_v10_4gpu.json", 4)):
        d = _line(name)
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
            assert k in d, (name, k)
        assert d["n_gpus"] == n and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["warmup"] >= 3
        assert d["config"]["workload"].startswith("S1 synthetic 640x480") and "model" not in d["config"]
        assert "l2" in d["config"]                                       # how the caches were treated between timed steps
        assert d["gpu_launches"] >= 12 * d["steps"]                      # our kernels, counted over the timed steps
        assert abs(d["value"] * d["ms_per_step"] / 1000.0 - 1.0) < 1e-6  # one frame per step, whole job
        e = d["e2e"]
        assert e["unit"] == "frames/s" and e["h2d_bytes_per_step"] == 640 * 480 * 2 and e["d2h_bytes_per_step"] > 0
        assert e["value"] != d["value"]                                  # measured on its own, host buffers inside the timed region
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if n == 1:
            cb = d["cpu_baseline"]
            assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
            assert r["traffic"] is not None and r["traffic"] > 0
    # the sharded runs track the same trajectory as each other, to the last bit of the final pose error
    assert _line("r01_bench_v10_2gpu.json")["config"]["final_pose_err_m"] == _line("r01_bench_v10_4gpu.json")["config"]["final_pose_err_m"]


def test_round2_bench_lines_carry_the_contract():
    """the round-2 lines of the last tree: the same keys, 11 launches per frame, the large-scene roofline of the integration kernel,
    the reference's own GPU path beside it, and the sharded runs on the trajectory of the single GPU"""
    lines = {1: _line("r02_bench_l.json"), 2: _line("r02_bench_j_2gpu.json"), 4: _line("r02_bench_j_4gpu.json")}
    for n, d in lines.items():
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
            assert k in d, (n, k)
        assert d["n_gpus"] == n and d["unit"] == "frames/s" and d["warmup"] >= 3 and d["vs_baseline"] is None
        assert d["config"]["workload"].startswith("S1 synthetic 640x480") and "l2" in d["config"]
        assert d["gpu_launches"] >= 11 * d["steps"]
        assert abs(d["value"] * d["ms_per_step"] / 1000.0 - 1.0) < 1e-6
        assert d["e2e"]["h2d_bytes_per_step"] == 640 * 480 * 2 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert d["voxel_updates_large_scene"]["unit"] == "voxel-updates/s"
    one = lines[1]
    r = one["roofline_large_scene"]
    assert r["kernel"] == "k_integrate" and r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["frac"] > 0.55 and r["traffic"] > 0           # the bandwidth kernel the north star grades, with its ncu DRAM bytes
    assert one["roofline"]["kernel"] == "k_icp_all" and one["roofline"]["traffic"] > 0
    g = one["reference_gpu"]
    assert g["debug_work_removed"]["frames_tracked"] == g["this_repo_same_mode_same_frames"]["frames_tracked"]   # same verdicts, same frames
    assert g["speedup_vs_debug_work_removed"] > 3
    assert one["cpu_baseline"]["kind"] == "reference" and one["cpu_baseline"]["cores"] >= 1
    assert one["ingest_from_files"]["value"] > one["value"]   # the step in front of the path does not cap it
    # voxel-updates/s adds across ranks (no exchange in the integration), frames/s of the 640x480 frame does not
    v = {n: d["voxel_updates_large_scene"]["value"] for n, d in lines.items()}
    assert v[4] > v[2] > v[1]
    assert lines[2]["config"]["final_pose_err_m"] == lines[4]["config"]["final_pose_err_m"]
