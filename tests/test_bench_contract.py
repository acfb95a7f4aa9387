"""bench.py's command line and the committed records it reads, without a GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse(argv):
    sys.path.insert(0, ROOT)
    import bench
    old = sys.argv
    sys.argv = ["bench.py"] + argv
    try:
        return bench.parse()
    finally:
        sys.argv = old


def test_defaults_name_the_measured_configuration():
    a = parse([])
    assert (a.gpus, a.steps, a.warmup, a.impl) == (1, 5, 3, "ours")
    assert a.warmup >= 3 and a.frames == 4096 and a.snr == 3.0 and a.precision == "fp32"
    assert (a.schedule, a.lanes) == (2, 4096)            # fused flooding iteration, one lane per frame of the step
    b = parse(["--precision", "fp64"])
    assert (b.schedule, b.lanes) == (2, 512)
    c = parse(["--schedule", "0"])
    assert (c.schedule, c.lanes) == (0, 0)               # two-phase kernel with the library's default lanes
    d = parse(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"])
    assert d.impl == "reference" and d.gpus == 2


def test_traffic_records_match_the_bench_lookup_keys():
    tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
    scheds = set()
    for ent in tj["entries"]:
        c = ent["config"]
        assert {"n", "frames", "max_iterations", "precision", "schedule", "lanes"} <= set(c)
        assert c["schedule"] in ("persistent", "launch", "fused")
        assert ent["dram_bytes_per_launch"] == ent["dram_bytes_read"] + ent["dram_bytes_write"]
        scheds.add(c["schedule"])
    assert {"persistent", "fused"} <= scheds
    # the committed default-run record carries every key of the bench contract
    rec = json.loads(open(os.path.join(ROOT, "profiles", "r1_bench_n1_3dB.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in rec, k
    assert rec["vs_baseline"] is None and rec["config"]["schedule"] == "fused" and "model" not in rec["config"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rec["roofline"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(rec["cpu_baseline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(rec["e2e"])


def test_round2_records_carry_the_wider_contract():
    """profiles/r2_*: the default-run record with its operating points and agreement block, the reference arm timed
    on whole frames, the multi-GPU lines, and a traffic entry whose keys the bench lookup finds."""
    rec = json.loads(open(os.path.join(ROOT, "profiles", "r2_bench_n1_3dB.json")).read().strip().splitlines()[-1])
    assert rec["n_gpus"] == 1 and rec["counters_all_reduced_inside_timed_region"] is True
    assert rec["roofline"]["traffic"] and 0 < rec["roofline"]["dram_frac"] < rec["roofline"]["frac"] < 1.0
    assert rec["roofline"]["l2_frac"] and rec["roofline"]["l2_frac"] < 1.0
    assert rec["e2e"]["h2d_bytes_per_step"] == 4096 * 32400 * 16
    assert rec["e2e_compact"]["h2d_bytes_per_step"] == 4096 * 32400 * 5
    agr = rec["cpu_baseline"]["agreement"]
    assert agr["success_iters_equal"] == 1.0 and agr["min_hard_decision_agreement"] >= 0.999
    pts = rec["operating_points"]
    assert {"config2_4dB", "config2_5dB", "config2_fp64_parity_mode", "config3_irregular_n131070_8pam",
            "config4_irregular_n1048576_2pam"} <= set(pts)
    for p in pts.values():
        assert p["value"] > 0 and 0 < p["roofline"]["frac"] < 1.0
    ref = json.loads(open(os.path.join(ROOT, "profiles", "r2_bench_reference_arm.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["cpu_baseline"]["kind"] == "reference" and ref["gpu_launches"] == 0
    # whole frames, wall clock: value x ms_per_step = frames of a step
    assert abs(ref["value"] * ref["ms_per_step"] / 1e3 - ref["frames_per_step"]) < 1e-6 * ref["frames_per_step"]
    assert ref["metric"] == rec["metric"] and ref["config"]["workload"] == rec["config"]["workload"]
    for n in (2, 4, 8):
        r = json.loads(open(os.path.join(ROOT, "profiles", f"r2_bench_n{n}.json")).read().strip().splitlines()[-1])
        assert r["n_gpus"] == n and r["scaling"] == "weak" and r["gpu_launches"] == 7 * r["steps"] * n
        assert r["value"] > 0.9 * n * rec["value"]
    tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    ent = tj["entries"][-1]
    assert ent["config"] == {"n": 64800, "frames": 4096, "max_iterations": 50, "precision": "fp32", "schedule": "fused",
                             "lanes": rec["config"]["decoder_lanes"]}
    assert ent["dram_bytes_per_launch"] == ent["dram_bytes_read"] + ent["dram_bytes_write"] == rec["roofline"]["traffic"]
    assert abs(ent["l2_bytes_per_launch"] - ent["l2_read_bytes"] - ent["l2_write_bytes"]) < 0.01 * ent["l2_bytes_per_launch"]
