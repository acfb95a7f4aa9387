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
    assert (a.schedule, a.lanes) == (2, 1024)            # fused flooding iteration, 1024 resident frames
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
