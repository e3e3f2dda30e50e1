"""Host-side pieces of bench.py (no GPU): the committed ncu summaries it quotes are found and
parsed for every configuration, the reference arm runs on the oracle with all host threads, and
the helpers that must never invent a number return None when there is nothing to read."""

import importlib.util
import json
import os
import subprocess
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


@pytest.mark.parametrize("config,horizon,f32", [(2, 16, False), (3, 12, False), (4, 16, False), (5, 32, False),
                                                 (5, 64, False), (6, 16, False)])
def test_bench_finds_the_committed_profile_of_every_config(bench, config, horizon, f32):
    """roofline.traffic and fp64.executed_ncu come from profiles/r*_*.txt (tools/ncu_summary.py
    output), matched by kernel, precision and shape tag -- never from a constant."""
    args = types.SimpleNamespace(config=config, horizon=horizon, method="active_set", f32=f32)
    prof = bench.measured_profile(args)
    assert prof["source"] is not None and os.path.exists(os.path.join(ROOT, "profiles", prof["source"]))
    assert prof["traffic"] > 0 and prof["flops_per_solve"] > 0 and prof["profile_batch"] > 0
    text = open(os.path.join(ROOT, "profiles", prof["source"])).read()
    assert bench.kernel_name(args).split("<")[0] in text
    assert ("<float," in text.split("\n")[1]) == (config == 4)


def test_bench_quotes_no_profile_where_there_is_none(bench):
    args = types.SimpleNamespace(config=6, horizon=16, method="active_set", f32=True)  # no fp32 walking-loop profile
    assert bench.measured_profile(args) == {"traffic": None, "flops_per_solve": None, "source": None}
    assert bench.bind_to_gpu_numa_node(0) is None or isinstance(bench.bind_to_gpu_numa_node(0), int)


def test_reference_arm_prints_the_contract_line():
    """``bench.py --impl reference``: the oracle port on all host threads (OMP_NUM_THREADS exported
    by torchrun is ignored), same metric / unit / config keys as the CUDA arm."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--batch", "2048"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "MPC solves/sec (batched)" and line["unit"] == "solves/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["e2e"] == {"value": line["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "BASELINE configs[1]" in line["config"]["workload"]
