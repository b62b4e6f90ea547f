"""bench.py's JSON line (the driver's contract): checked on the lines committed under profiles/ (measured on a B200 this round),
so that a change to bench.py that drops a key is caught on CPU."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
             "config", "clocks", "e2e"}


def load(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r1_bench_n1_v9.json", "r1_bench_n1_final.json", "r1_bench_n2_final.json", "r1_bench_512_n1.json"])
def test_solver_line(name):
    d = load(name)
    assert BASE_KEYS <= set(d) and {"gpu_launches", "roofline"} <= set(d), sorted(set(d))
    assert d["metric"] == "solver_gvoxel_iters_per_s" and d["unit"] == "Gvoxel-iter/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["value"] > 0
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] == 640 * 480 * 2 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"] * 1.001                      # the end-to-end number is not the device-timed one repeated
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1.05
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c) and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # value is consistent with the time it claims
    vox = {"r1_bench_512_n1.json": 512 ** 3}.get(name, 256 ** 3)
    iters = 100 if name == "r1_bench_512_n1.json" else 200
    assert d["value"] == pytest.approx(vox * iters / (d["ms_per_step"] * 1e-3) / 1e9, rel=1e-6)


def test_cpu_baseline_and_reference_arm():
    d = load("r1_bench_n1_v9.json")
    b = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(b) and b["kind"] == "port" and b["cores"] >= 1 and b["unit"] == d["unit"]
    r = load("r1_bench_reference.json")
    assert r["impl"] == "reference" and r["metric"] == d["metric"] and r["unit"] == d["unit"] and r["config"]["workload"] == d["config"]["workload"]
    assert r["cpu_baseline"]["kind"] == "reference" and r["cpu_baseline"]["value"] == r["value"]
    assert d["value"] > 5 * r["value"]                           # 2946 vs 462 iterations/s on the same B200


def test_pipeline_lines():
    for name, n in (("r1_bench_pipeline_n1.json", 1), ("r1_bench_pipeline_n2.json", 2)):
        d = load(name)
        assert BASE_KEYS <= set(d) and d["metric"] == "pipeline_frames_per_s" and d["n_gpus"] == n and d["mesh_vertices_last_frame"] > 10000


# ---- round 2 lines ---------------------------------------------------------------------------------------------------------------
R2_SOLVER = ["r2_bench_n1.json", "r2_bench_n2_peer.json", "r2_bench_n4_peer.json", "r2_bench_n8_peer.json", "r2_bench_n8_nccl.json", "r2_bench_512_n1.json"]


@pytest.mark.parametrize("name", R2_SOLVER)
def test_round2_solver_lines(name):
    d = load(name)
    assert BASE_KEYS <= set(d) and {"gpu_launches", "roofline", "parity"} <= set(d)
    assert d["metric"] == "solver_gvoxel_iters_per_s" and d["higher_is_better"] is True and d["dtype"] == "f32" and d["gpu_launches"] > 0
    p = d["parity"]                                            # checked after the timed regions, in the same run
    assert p["checked"] is True and p["bit_exact"] is True and p["iterations"] == 200
    assert all(p[k]["bit_exact"] and p[k]["words_differing"] == 0 for k in ("psi", "psi_inv", "phi_n_psi", "phi_global_psi_inv"))
    assert ("oracle/_ref" in p["against"]) == (d["n_gpus"] == 1)
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 640 * 480 * 2 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"] * 1.001
    r = d["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["bound"] == "hbm"
    dim = 512 if "512" in name else 256
    assert d["value"] == pytest.approx(dim ** 3 * 200 / (d["ms_per_step"] * 1e-3) / 1e9, rel=1e-6)
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_round2_scaling_and_the_512_line():
    v = {n: load("r2_bench_n%d_peer.json" % n)["value"] for n in (2, 4, 8)}
    v[1] = load("r2_bench_n1.json")["value"]
    assert v[1] < v[2] < v[4] < v[8] and v[8] > 3.9 * v[1]
    assert load("r2_bench_n8_peer.json")["value"] > 1.4 * load("r2_bench_n8_nccl.json")["value"]      # peer mode vs the NCCL schedule
    ours, ref = load("r2_bench_n8_peer.json")["extra_512"], load("r2_bench_ref.json")["extra_512"]
    assert "512^3" in ours["config"]["workload"] and "512^3" in ref["config"]["workload"] and ours["parity"]["bit_exact"] is True
    assert ours["e2e"]["frames_per_s"] > 6 * ref["e2e"]["frames_per_s"]                                # north star: >= 6x at 512^3 on 8 GPUs


def test_round2_reference_arm_and_pipeline():
    d, r = load("r2_bench_n1.json"), load("r2_bench_ref.json")
    assert r["impl"] == "reference" and r["metric"] == d["metric"] and r["config"]["workload"] == d["config"]["workload"]
    assert d["e2e"]["frames_per_s"] > 6 * r["e2e"]["frames_per_s"] and d["cpu_baseline"]["kind"] == "port"
    p, pr = load("r2_bench_pipe_n1.json"), load("r2_bench_pipe_ref.json")
    assert pr["impl"] == "reference" and p["metric"] == pr["metric"] == "pipeline_frames_per_s" and p["config"]["workload"] == pr["config"]["workload"]
    assert p["value"] > 5 * pr["value"]
    f = [load("r2_bench_pipe_n%d.json" % n)["value"] for n in (2, 4, 8)]
    assert p["value"] < f[0] < f[1] < f[2]
