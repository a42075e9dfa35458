"""CPU test of the bench.py contract that can run without a GPU: the reference arm prints ONE JSON line with
the keys the driver reads, and only rank 0 prints under a multi-rank launch."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "3", "--warmup", "1", "--gpus", "1"], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line(built_lib):
    lines = _run()
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["metric"] == "logp_grad_chain_evals_per_sec" and d["unit"] == "chain-evals/s"
    assert d["higher_is_better"] is True and d["steps"] == 3 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same `config` object as the CUDA arm prints (the driver compares the two)
    import bench
    assert d["config"] == bench.bench_config(bench.WORKLOADS["cfg1"], bench.WORKLOADS["cfg1"]["C"])


def test_reference_arm_other_ranks_are_silent(built_lib):
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
