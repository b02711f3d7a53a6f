"""CPU: the parts of bench.py that run without a GPU — the roofline arithmetic of SURVEY §8d and the
`--impl reference` arm (the oracle's torch.sparse.mm over the workload's graph), whose JSON line the driver parses."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_match_the_survey_figures():
    # config 2: 200 M directed entries x 264 B + 2 M nodes x 260 B = 53.32 GB per layer
    assert bench.algorithmic_bytes_per_layer(200_000_000, 2_000_000, 64) == 53_320_000_000
    # config 5: 2 B x 520 + 20 M x 516 = 1050.32 GB per layer
    assert bench.algorithmic_bytes_per_layer(2_000_000_000, 20_000_000, 128) == 1_050_320_000_000
    U, I, E, D, L = bench.WORKLOADS["cfg2"]
    assert (U, I, E, D, L) == (1_000_000, 1_000_000, 100_000_000, 64, 3)


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                          "--steps", "1", "--warmup", "3", "--cpu-sample-rows", "2000"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["config"]["same_config"] is True                       # the small workload fits the budget: full graph
    v = d["cpu_baseline"]["variants"]
    assert v["csr_full"] == d["value"] and v["coo_5pct_rows"] > 0 and v["dense_edge_1pct_rows"] > 0
    assert d["cpu_baseline"]["cores"] in [int(k) for k in d["cpu_baseline"]["thread_sweep_Medges_per_s"]]


def test_reference_arm_falls_back_to_a_row_slice_when_the_budget_is_small():
    env = dict(os.environ, B200GCN_REF_BUDGET_S="0.05")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                          "--steps", "2", "--warmup", "3"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert d["config"]["same_config"] is False and "rows [0," in d["config"]["sample"] and d["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--workload", "small"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
