"""Smoke tests of the CPU-only tools (they produce numbers quoted in
profiles/README.md, so they must keep running)."""

import importlib.util
import os
import types

import pytest

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


def test_bench_contract_two_ranks_on_cpu(tmp_path):
    """bench.py's driver contract on the CPU plumbing path (gloo, 2 ranks):
    one JSON line from rank 0 with the keys the driver reads, K-step windows
    repeated until the requested device time is covered."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items()
           if not k.startswith("ADAPTDL_")}
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29733", os.path.join(root, "bench.py"),
           "--gpus", "2", "--steps", "2", "--warmup", "3", "--device", "cpu",
           "--workload", "ncf", "--min-timed-ms", "1e9", "--max-windows",
           "2"]
    proc = subprocess.run(cmd, env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600,
                          cwd=str(tmp_path))
    assert proc.returncode == 0, proc.stderr[-3000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup",
                "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "clocks",
                "param_dtype", "grad_dtype", "step_mode"):
        assert key in out, key
    assert out["n_gpus"] == 2 and out["steps"] == 2
    assert out["steps_timed"] == 4 and len(out["windows"]["device_ms"]) == 2
    assert out["config"]["global_batch"] == 512
    assert out["value"] > 0 and out["e2e"]["value"] > 0
    # the framework's own step profiler, as booked during the run
    assert out["step_profile"]["steps"] > 0
    assert out["step_profile"]["step_ms"] >= out["step_profile"]["sync_ms"] >= 0
    assert set(out["config"]) == {"workload", "model", "global_batch",
                                  "local_batch", "seq_len", "parallelism",
                                  "optimizer", "adaptive", "compute", "l2"}


@pytest.mark.parametrize("impl", ["reference"])
def test_bench_config1_linreg_reference_arm_on_cpu(tmp_path, impl):
    """BASELINE config 1 (linear regression, 2 gloo replicas on the CPU)
    through bench.py's unmodified-reference arm (the own arm's line is
    covered by the contract test above): workload description, a positive
    value, and the arm really is the package under baseline/_ref."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if impl == "reference" and not os.path.isdir(
            os.path.join(root, "baseline", "_ref", "adaptdl")):
        pytest.skip("baseline/_ref is not installed")
    env = {k: v for k, v in os.environ.items()
           if not k.startswith("ADAPTDL_")}
    port = "29735" if impl == "own" else "29737"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", port, os.path.join(root, "bench.py"),
           "--gpus", "2", "--steps", "10", "--warmup", "3", "--device", "cpu",
           "--workload", "linreg", "--impl", impl, "--min-timed-ms", "0"]
    proc = subprocess.run(cmd, env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600,
                          cwd=str(tmp_path))
    assert proc.returncode == 0, proc.stderr[-3000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out["impl"] == impl and out["n_gpus"] == 2
    assert out["config"]["workload"] == "linreg"
    assert out["config"]["global_batch"] == 128
    assert out["dtype"] == "fp32" and "host-timed" in out["metric"]
    assert out["value"] > 0 and out["e2e"]["value"] > 0
    assert out["step_profile"]["steps"] > 0
    if impl == "reference":
        assert out["gpu_launches"] is None and "reducer" not in out


def test_bench_clock_sampler_only_keeps_samples_inside_windows(monkeypatch):
    """The sampler polls NVML in-process and keeps a row only while a timed
    window is open (a 40 ms window is too short for an ``nvidia-smi``
    start-up: the N=8 headline run of round 2 came back without clocks)."""
    import importlib.util
    import os
    import sys
    import time
    import types
    fake = types.ModuleType("pynvml")
    fake.NVML_CLOCK_SM = 1
    fake.calls = 0
    fake.nvmlInit = lambda: None
    fake.nvmlDeviceGetHandleByIndex = lambda i: ("gpu", i)

    def by_uuid(uuid):
        raise RuntimeError("unknown uuid")
    fake.nvmlDeviceGetHandleByUUID = by_uuid
    fake.nvmlDeviceGetMaxClockInfo = lambda h, kind: 1965

    def clock(h, kind):
        fake.calls += 1
        return 1950
    fake.nvmlDeviceGetClockInfo = clock
    fake.nvmlDeviceGetCurrentClocksEventReasons = lambda h: 0x4 | 0x1
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(
        "bench_under_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    idle = bench.ClockSampler(False, 0)
    idle.begin()
    idle.end()
    assert idle.close() is None

    sampler = bench.ClockSampler(True, 3, "GPU-not-there")
    assert sampler.source == "nvml"
    time.sleep(0.05)
    assert sampler.rows == [] and fake.calls == 0     # no window open yet
    sampler.begin()
    time.sleep(0.25)
    sampler.end()
    kept = len(sampler.rows)
    assert kept >= 3
    time.sleep(0.05)
    assert len(sampler.rows) == kept                  # closed window
    got = sampler.close()
    assert got["samples"] == kept and got["source"] == "nvml"
    assert got["sm_mhz"] == 1950 and got["sm_max_mhz"] == 1965
    assert got["reasons"] == ["sw_power_cap"]         # gpu_idle bit ignored

    # a window the sampling thread never got to: the timing loop's mid-window
    # poke reads the clocks once itself; a second poke adds nothing
    monkeypatch.setattr(bench.ClockSampler, "PERIOD_S", 60.0)
    sampler = bench.ClockSampler(True, 0)
    sampler.begin()
    sampler.poke()
    sampler.poke()
    sampler.end()
    sampler.poke()                                    # closed: ignored
    assert len(sampler.rows) == 1
    assert sampler.close()["samples"] == 1


def test_gradient_path_protocol_model():
    """tools/protocol_model.py: every interleaving of the small
    configurations is free of stale / early reads and of deadlocks, random
    skewed schedules of a larger one too, and each deliberately broken
    variant of the protocol is caught."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(
        "protocol_model", os.path.join(root, "tools", "protocol_model.py"))
    model = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(model)
    assert model.explore_all(model.Config(2, 1, 3)) > 1000
    assert model.explore_all(model.Config(3, 1, 2)) > 10000
    assert model.explore_all(
        model.Config(2, 1, 2, buckets=("one", "two", "one"))) > 1000
    assert model.explore_all(
        model.Config(2, 1, 2, buckets=("nvls", "one", "two"))) > 1000
    assert model.explore_all(model.Config(3, 1, 1, buckets=("nvls",))) > 1000
    model.explore_random(
        model.Config(3, 2, 3, buckets=("nvls", "two")), 200, seed=4)
    model.explore_random(model.Config(3, 2, 4), 300, seed=5)
    model.explore_random(
        model.Config(4, 2, 3, buckets=("two", "two", "one")), 100, seed=6)
    for name in model.BROKEN:
        assert model.check_broken(name, ranks=2, ctas=1, steps=3), name
        assert model.check_broken(name, ranks=3, ctas=2, steps=3, runs=500), \
            name
    # found by the model: the exchange record would not need its double
    # buffer (kept in the kernels as a free safety margin)
    assert model.explore_all(model.Config(2, 1, 3, single_xchg=True)) > 1000
