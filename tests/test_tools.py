"""Smoke tests of the CPU-only tools (they produce numbers quoted in
profiles/README.md, so they must keep running)."""

import importlib.util
import os
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_scheduler_simulation_runs_and_conserves_work():
    sim = _load("sched_sim")
    args = types.SimpleNamespace(nodes=2, gpus_per_node=4, hours=0.5,
                                 interval=60.0, pop=12, generations=6)
    for adaptive, fixed in ((False, None), (False, 4), (True, None)):
        out = sim.simulate(8.0, args, adaptive, seed=1, static_gpus=fixed)
        assert out["jobs"] > 0 and out["finished"] == out["jobs"]
        assert out["avg_jct_hours"] > 0
        assert out["gpu_hours"] >= out["node_hours"] > 0
        # nobody can hold more than the cluster has
        assert out["gpu_hours"] <= 8 * out["makespan_hours"] + 1e-6
    # same seed, same workload in all arms
    a = sim.simulate(8.0, args, False, seed=1)
    b = sim.simulate(8.0, args, True, seed=1)
    assert a["single_gpu_hours_of_work"] == b["single_gpu_hours_of_work"]


def test_auto_batch_size_replay_is_u_shaped():
    sim = _load("autobsz_sim")
    from adaptdl_b200.goodput import GradParams, PerfParams
    perf = PerfParams(1.2e-3, 6.5e-6, 2.0e-4, 4.0e-5, 1.5e-4, 3.0e-6, 1.3)

    def noise(f):
        return GradParams(0.0014 * (1 - f) + 0.00012 * f,
                          0.0005 * (1 - f) + 0.0012 * f)
    times = {b: sim.time_to_train(perf, 128, 3e5, noise, 8, 1, b,
                                  dt_progress=0.05)[0]
             for b in (128, 512, 4096)}
    auto, trace = sim.time_to_train(perf, 128, 3e5, noise, 8, 1, None,
                                    dt_progress=0.05)
    assert times[512] < times[128] and times[512] < times[4096]
    assert auto <= min(times.values()) * 1.01
    assert trace[-1] > trace[0]          # batch size grows with the noise


def test_docs_reference_existing_paths():
    check_docs = _load("check_docs")
    problems = []
    for path in check_docs.markdown_files():
        problems.extend(check_docs.check_file(path))
    assert not problems, "\n".join(problems)
    assert check_docs.repo_path("adaptdl_b200/ops/bn_act.py:12") == \
        "adaptdl_b200/ops/bn_act.py"
    assert check_docs.repo_path("tests/test_ops.py::test_x") == \
        "tests/test_ops.py"
    assert check_docs.repo_path("torch/data.py") is None


def test_host_trace_of_a_training_run(tmp_path):
    """ADAPTDL_B200_TRACE writes a Chrome trace with the step-path spans;
    without it the decorators are the identity."""
    import json
    import subprocess
    import sys
    from adaptdl_b200.utils import trace
    assert not trace.ENABLED

    def plain():
        pass
    assert trace.traced("x")(plain) is plain
    assert trace.span("x") is trace.span("y")          # shared no-op

    script = os.path.join(ROOT, "examples", "linear_regression", "main.py")
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="",
               ADAPTDL_B200_TRACE=str(tmp_path / "trace"),
               ADAPTDL_CHECKPOINT_PATH=str(tmp_path))
    subprocess.run([sys.executable, script, "--epochs", "2", "--size",
                    "512"], env=env, check=True, timeout=240,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    with open(str(tmp_path / "trace") + ".rank0.json") as f:
        events = json.load(f)["traceEvents"]
    names = {e["name"] for e in events if e["ph"] == "X"}
    assert {"forward", "backward_end"} <= names, names
    spans = [e for e in events if e["name"] == "forward"]
    assert len(spans) >= 4 and all(e["dur"] >= 0 for e in spans)


def test_policy_bench_smoke():
    bench = _load("policy_bench")
    row = bench.run(num_jobs=6, num_nodes=2, gpus_per_node=4, cycles=2,
                    seed=0)
    assert row["gpus_allocated"] <= 8 and row["jobs_running"] >= 1
    assert len(row["cycle_seconds"]) == 2 and row["sum_speedup"] > 0


def test_conv_bench_plumbing_on_cpu():
    bench = _load("conv_bench")
    measure = bench.timer(__import__("torch").device("cpu"), 1, 0)
    import torch
    rows = [bench.bench_layer(spec, 2, torch.device("cpu"), torch.float32,
                              measure)
            for spec in bench.RESNET18_CONVS[:3]]
    assert {"fprop_us", "dgrad_us", "wgrad_us"} <= set(rows[0])
    assert all(r["fprop_us"] > 0 and r["gflop"] > 0 for r in rows)
