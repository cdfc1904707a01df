"""CPU: bench.py's CPU arm (oracle/cpu_arm.py) -- the worker plan and a tiny 2-worker run (dnn simulator: seconds)."""
import os

import numpy as np

from oracle import cpu_arm


def test_plan_covers_every_core_once():
    cores = cpu_arm.host_cores()
    W, threads, blocks = cpu_arm.plan(1, None)
    assert W == len(cores) and threads == 1 and sorted(sum(blocks, [])) == cores
    W, threads, blocks = cpu_arm.plan(2, 3)
    assert W == min(3, len(cores) // 2) or len(cores) < 2
    assert all(len(b) == threads for b in blocks)
    assert len(set(sum(blocks, []))) == sum(len(b) for b in blocks)


def test_parallel_arm_two_workers_dnn():
    import bench
    names = ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS")
    before = {k: os.environ.get(k) for k in names}
    cfg = dict(bench.base_config(8, False, simulator="dnn"), is_eval=False, cache_size=64)
    r = cpu_arm.run_parallel(cfg, False, "dnn", episodes=2, warmup=1, threads=1, workers=min(2, len(cpu_arm.host_cores())))
    W = r["workers"]
    assert r["rows_per_episode"] == 8 * W and r["transitions_per_episode"] == 8 * W * 9
    assert len(r["episode_s"]) == 2 and all(t > 0 for t in r["episode_s"])
    assert np.isclose(r["value"], r["transitions_per_episode"] / r["median_s"])
    assert 0.0 < r["nn_share"] <= 1.0
    # the thread limits handed to the children do not stay in the parent's environment
    assert {k: os.environ.get(k) for k in names} == before


def test_parallel_arm_sizes_its_episodes_from_a_calibration_run():
    """budget_s: one calibration episode, then rows per worker cut so that warm-up + timed episodes fit the budget."""
    import bench
    cfg = dict(bench.base_config(64, False, simulator="dnn"), is_eval=False, cache_size=256)
    W = min(2, len(cpu_arm.host_cores()))
    r = cpu_arm.run_parallel(cfg, False, "dnn", episodes=2, warmup=1, threads=1, workers=W, budget_s=6.0)
    assert 1 <= r["rows_per_episode"] // r["workers"] <= 64 and "calibration" in r["note"]
    assert r["transitions_per_episode"] == r["rows_per_episode"] * 9 and len(r["episode_s"]) == 2


def test_cpu_quota_is_read_from_the_cgroup(monkeypatch, tmp_path):
    q = cpu_arm.cpu_quota()
    assert q is None or q > 0
    cores = cpu_arm.host_cores()
    monkeypatch.setattr(cpu_arm, "cpu_quota", lambda: 2.0)
    W, threads, blocks = cpu_arm.plan(1, None)
    assert W == min(2, len(cores)) and threads == 1
