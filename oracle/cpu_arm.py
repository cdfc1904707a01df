"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- the CPU arm of bench.py (`--impl reference` and `cpu_baseline`).

Runs the oracle port of the reference's CPU env (oracle/env_np.py + dien_np.py / dnn_np.py) on ALL host cores the way
the reference itself scales on CPU: by processes over independent env rows (its RLlib rollout workers / HTTP env
servers, `script/modelfree_train.py:179-217`, are one process per env batch).  W worker processes, each pinned to its
own block of `threads` cores (sched_setaffinity + a BLAS thread limit), each replaying offline-action episodes of
`rows` env rows of its own log slice; the episodes start together behind a barrier, an episode's time is the slowest
worker's, the arm's throughput is W x rows x max_steps / median episode time.

Nothing under rl4rs_b200/ imports this module.
"""
import multiprocessing as mp
import os
import time

import numpy as np


class _TimedNet(object):
    """Wraps the oracle's network so the arm can report its host-Python vs NN split."""

    def __init__(self, net):
        self.d, self.nn_s = net, 0.0

    def _t(self, fn, feat):
        t0 = time.perf_counter()
        out = fn(feat)
        self.nn_s += time.perf_counter() - t0
        return out

    def obs_layer(self, feat):
        return self._t(self.d.obs_layer, feat)

    def reward_layer(self, feat):
        return self._t(self.d.reward_layer, feat)


def _episodes(cfg, log, catalog, weights, seq, simulator, episodes, warmup, sync=None):
    from oracle.dien_np import DienOracle
    from oracle.dnn_np import DnnOracle
    from oracle.env_np import OracleEnv
    _tune_malloc()
    np.random.seed(0)
    net = _TimedNet((DnnOracle if simulator == "dnn" else DienOracle)(weights, np.float32))
    env = OracleEnv(cfg, log, catalog, net, seq=seq)
    T = cfg["max_steps"]
    t_begin, t_end, nn = [], [], []
    for ep in range(warmup + episodes):
        net.nn_s = 0.0
        if sync is not None:
            sync()
        t0 = time.time()
        env.reset()
        for _ in range(T):
            env.step(env.offline_action)
        t1 = time.time()
        if ep >= warmup:
            t_begin.append(t0); t_end.append(t1); nn.append(net.nn_s)
    return t_begin, t_end, nn


def _tune_malloc():
    """glibc hands every NumPy temporary above 128 KB back to the kernel (mmap / munmap, then a page fault per 4 KB on
    the next one): half of an episode went to the kernel that way, more under a hypervisor, and the run-to-run spread
    with it.  Keep freed blocks in the heap instead: the same settings as MALLOC_MMAP_THRESHOLD_ / MALLOC_TRIM_THRESHOLD_ /
    MALLOC_TOP_PAD_ in the environment.  (8 workers x 128 rows on the 8-core build container: 630-700 tr/s with episode
    times of 8-18 s before, 1.5-1.8 k tr/s at 5-6 s after.)"""
    try:
        import ctypes
        libc = ctypes.CDLL("libc.so.6")
        libc.mallopt(-3, 32 << 20)          # M_MMAP_THRESHOLD: its maximum, 32 MB
        libc.mallopt(-1, 1 << 30)           # M_TRIM_THRESHOLD
        libc.mallopt(-2, 256 << 20)         # M_TOP_PAD
    except Exception:
        pass


def _limit_threads(threads):
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=threads)
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def _worker(idx, cores, spec, barrier, out, go, rows):
    try:
        if cores:
            try:
                os.sched_setaffinity(0, cores)
            except Exception:
                pass
        from rl4rs_b200 import synth
        cfg = spec["cfg"]
        catalog = synth.make_catalog()
        log = synth.make_log(spec["n_log"], pages=4 if spec["seq"] else 1, catalog=catalog, seed=synth.LOG_SEED + idx)
        weights = synth.make_dnn_weights(cfg) if spec["simulator"] == "dnn" else synth.make_weights(cfg)
        sync = lambda: barrier.wait(timeout=spec["timeout_s"])
        with _limit_threads(spec["threads"]):
            if spec["calibrate_rows"]:
                # one episode of a few rows, all workers together: the parent sizes the timed episodes from its duration
                t0, t1, _ = _episodes(dict(cfg, batch_size=spec["calibrate_rows"]), log, catalog, weights, spec["seq"],
                                      spec["simulator"], 1, 0, sync=sync)
                out.put((idx, t1[0] - t0[0], None))
                go.wait(timeout=spec["timeout_s"])
                cfg = dict(cfg, batch_size=int(rows.value))
            res = _episodes(cfg, log, catalog, weights, spec["seq"], spec["simulator"], spec["episodes"], spec["warmup"], sync=sync)
        out.put((idx, res, None))
    except Exception as e:                                   # noqa: BLE001 -- reported to the parent, which falls back
        for b in (barrier, go):
            try:
                b.abort()
            except Exception:
                pass
        out.put((idx, None, "%s: %s" % (type(e).__name__, e)))


def memory_limit_bytes():
    """What this process tree may use: available RAM, or the cgroup limit when that is lower."""
    limits = []
    try:
        import psutil
        limits.append(int(psutil.virtual_memory().available))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
        try:
            v = open(path).read().strip()
            if v.isdigit() and int(v) < (1 << 60):
                limits.append(int(v))
        except Exception:
            pass
    return min(limits) if limits else None


def workers_cap(per_worker_bytes=300e6, share=0.5):
    """At most this many workers fit in `share` of the memory limit.  A worker holds its own copy of the weights (two
    51 MB embedding tables) next to the interpreter, NumPy, its log slice and activations: 0.23 GB resident.  (Sharing
    the tables as read-only tmpfs mappings was tried and dropped: 2.5x slower episodes, the time going to the kernel.)"""
    lim = memory_limit_bytes()
    return None if lim is None else max(1, int(lim * share / per_worker_bytes))


def cpu_quota():
    """CPUs' worth of run time the cgroup allows (cpu.max / cfs_quota), or None: a container can SEE 128 hardware threads
    and be allowed 16 of them; more single-threaded workers than that only time-slice."""
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        return None if q == "max" else float(q) / float(per)
    except Exception:
        pass
    try:
        q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        return None if q <= 0 else q / per
    except Exception:
        return None


def host_cores():
    try:
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return list(range(os.cpu_count() or 1))


def plan(threads=None, workers=None):
    """-> (workers, threads per worker, core blocks).  Default: blocks of 4 cores, every core used."""
    cores = host_cores()
    n = len(cores)
    threads = max(1, min(threads or 4, n))
    quota = cpu_quota()
    if quota is not None:
        n = max(threads, min(n, int(quota + 0.5)))
    workers = max(1, min(workers or n // threads, n // threads))
    return workers, threads, [cores[w * threads:(w + 1) * threads] for w in range(workers)]


def _summary(B, T, workers, threads, ep_s, nn_share, n_cores, note=None):
    med = float(np.median(ep_s))
    out = {"value": workers * B * T / med, "episode_s": [round(t, 3) for t in ep_s], "median_s": med,
           "total_s": float(sum(ep_s)), "spread": float((max(ep_s) - min(ep_s)) / med) if med > 0 else 0.0,
           "nn_share": nn_share, "workers": workers, "threads_per_worker": threads, "threads": workers * threads,
           "host_cores": n_cores, "rows_per_episode": workers * B, "transitions_per_episode": workers * B * T}
    if note:
        out["note"] = note
    return out


def run_single(cfg, log, catalog, weights, seq, simulator, episodes, warmup, threads):
    """One process, `threads` BLAS threads: configs[0] (batch 32) and the fallback."""
    threads = max(1, min(threads, len(host_cores())))
    with _limit_threads(threads):
        t0, t1, nn = _episodes(cfg, log, catalog, weights, seq, simulator, episodes, warmup)
    ep = [b - a for a, b in zip(t0, t1)]
    return _summary(cfg["batch_size"], cfg["max_steps"], 1, threads, ep, float(sum(nn) / max(sum(ep), 1e-9)),
                    len(host_cores()))


def _get(out, procs, timeout_s):
    """Next message of a worker; gives up at once when a worker died without reporting (e.g. it could not be spawned)."""
    import queue
    t_end = time.time() + timeout_s
    while True:
        try:
            return out.get(timeout=2.0)
        except queue.Empty:
            dead = [p for p in procs if p.exitcode not in (None, 0)]
            if dead:
                raise RuntimeError("%d worker process(es) died (exit code %s)" % (len(dead), dead[0].exitcode))
            if time.time() > t_end:
                raise RuntimeError("no result within %.0f s" % timeout_s)


def run_parallel(cfg, seq, simulator, episodes, warmup, threads=None, workers=None, timeout_s=900.0, budget_s=None):
    """All host cores: W processes x `threads` cores, at most cfg['batch_size'] rows per worker and episode.
    budget_s: wall-clock target for the warm-up + timed episodes together; the workers then run one calibration episode
    of a few rows first and the rows per worker are cut (never raised) so that an episode takes about
    budget_s / (episodes + warmup), within [2 s, 20 s] -- the bench contract asks for K steps in a few minutes on any host."""
    W, threads, blocks = plan(threads, workers)
    B, T = cfg["batch_size"], cfg["max_steps"]
    cal_rows = min(B, 8) if budget_s else 0
    spec = {"cfg": cfg, "seq": seq, "simulator": simulator, "episodes": episodes, "warmup": warmup,
            "threads": threads, "n_log": max(4 * B, 2048), "timeout_s": timeout_s, "calibrate_rows": cal_rows}
    ctx = mp.get_context("spawn")
    barrier, out, go, rows = ctx.Barrier(W), ctx.Queue(), ctx.Barrier(W + 1), ctx.Value("i", B)
    procs = [ctx.Process(target=_worker, args=(w, blocks[w], spec, barrier, out, go, rows), daemon=True) for w in range(W)]
    # the children size their BLAS pools when numpy loads: tell them before they start (W x host_cores threads otherwise)
    names = ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS")
    saved = {k: os.environ.get(k) for k in names}
    os.environ.update({k: str(threads) for k in names})
    try:
        for p in procs:
            p.start()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    results, err, note = {}, None, None
    try:
        if cal_rows:
            cal = []
            for _ in range(W):
                idx, t, e = _get(out, procs, timeout_s)
                if e is not None:
                    raise RuntimeError("worker %d: %s" % (idx, e))
                cal.append(t)
            per_row = max(cal) / cal_rows                        # seconds per env row and episode, slowest worker
            target = min(20.0, max(2.0, budget_s / float(episodes + warmup)))
            rows.value = B = max(1, min(B, int(target / per_row)))
            note = "rows per worker sized from a %d-row calibration episode (%.2f s): target %.1f s per episode" % (cal_rows, max(cal), target)
            go.wait(timeout=timeout_s)
        for _ in range(W):
            idx, res, e = _get(out, procs, timeout_s)
            if e is not None:
                err = err or "worker %d: %s" % (idx, e)
            else:
                results[idx] = res
    except Exception as e:                                   # noqa: BLE001 -- queue timeout, a worker's error, a broken barrier
        err = err or "%s: %s" % (type(e).__name__, e)
        for b in (barrier, go):
            try:
                b.abort()
            except Exception:
                pass
    for p in procs:
        p.join(timeout=5)
        if p.is_alive():
            p.terminate()
    if err is not None or len(results) != W:
        raise RuntimeError(err or "a worker died")
    # episode e: from the first worker's start to the last worker's end (they start together behind the barrier)
    ep = [max(results[w][1][e] for w in range(W)) - min(results[w][0][e] for w in range(W)) for e in range(episodes)]
    busy = sum(b - a for w in range(W) for a, b in zip(results[w][0], results[w][1]))
    nn = sum(sum(results[w][2]) for w in range(W))
    return _summary(B, T, W, threads, ep, float(nn / max(busy, 1e-9)), len(host_cores()), note)
