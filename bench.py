#!/usr/bin/env python
"""Benchmark of the hot path: frames/s of the full planner (sort -> match -> path) on synthetic FSG-shaped cone maps.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

One "step" = one pass of the planner over one batch of FRAMES_PER_GPU synthetic frames per GPU (weak scaling: the
global batch is N x FRAMES_PER_GPU frames of the same synthetic stream, partitioned over the ranks, with NCCL all-gathers
of the output paths inside every step).  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_GPU = 10240  # BASELINE metric: "10k synthetic FSG cone maps"; frame shape of configs[1] (colours known)
CONFIG5_PER_GPU = 8192  # BASELINE configs[4]: 65 536 mixed frames on 8 GPUs = 8 192 per GPU
# Control-flow rehearsal for tests/test_bench_rehearsal.py ONLY: FSD_BENCH_REHEARSAL=<frames> runs this file's rank /
# collective / timing control flow on CPU tensors over gloo with a planner stub that plans NOTHING (zeros), so that a
# multi-rank deadlock in the harness shows up without a GPU.  The JSON line it prints says so and carries no value.
REHEARSAL = int(os.environ.get("FSD_BENCH_REHEARSAL", "0"))
if REHEARSAL:
    FRAMES_PER_GPU = REHEARSAL
    CONFIG5_PER_GPU = REHEARSAL
SEED = 2
METRIC = "frames/sec full PathPlanner on 10k synthetic FSG cone maps at 1/2/4/8 B200"
UNIT = "frames/s"


def bench_config(world):
    """The workload both arms (`--impl ours` / `--impl reference`) run: identical dict, so the driver can tell."""
    return {"workload": f"gen_autocross(seed={SEED}), colours known, ~87 cones/frame, fp32 coordinates, {FRAMES_PER_GPU} "
                        f"frames per GPU (BASELINE configs[1] frame shape at the metric's 10k batch)",
            "frames_per_gpu": FRAMES_PER_GPU, "global_frames": FRAMES_PER_GPU * world, "seed": SEED,
            "semantics": "fresh planner per frame", "l2": "256 MiB buffer written between timed steps (GPU arm)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of ALL GPUs of the box DURING the timed region: one sampler process (rank 0)
    instead of one per rank."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpus):
        self.gpus, self.rows, self.proc = list(gpus), [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", ",".join(str(g) for g in self.gpus), "-lms", "100"],
                                         stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def n_samples(self):
        return len(self.rows) // max(len(self.gpus), 1)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        num = lambda v: v.replace(".", "").isdigit()
        rows = [r for r in self.rows if len(r) >= 9 and num(r[1])]
        sm = [float(r[1]) for r in rows]
        mx = [float(r[2]) for r in rows if num(r[2])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": reasons, "samples": len(sm)}
        if len(self.gpus) > 1:
            out["per_gpu_sm_mhz"] = {g: float(np.median([float(r[1]) for r in rows if r[0] == str(g)] or [0.0]))
                                     for g in self.gpus}
        return out


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def reference_numba(frames=2048, single=512, timeout=900):
    """The UNMODIFIED reference (numba + scipy) on the host cores, in a child process (baseline/reference_runner.py):
    {"available": True, "single": {...}, "pool": {...}, ...} or {"available": False, "reason": ...}."""
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "reference_runner.py"), "--seed", str(SEED), "--frames",
           str(frames), "--single-frames", str(single), "--check"]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"available": False, "reason": f"reference_runner.py failed (exit {r.returncode}): {r.stderr[-300:]}"}
        return json.loads(lines[-1])
    except subprocess.TimeoutExpired:
        return {"available": False, "reason": f"reference_runner.py did not finish within {timeout} s"}
    except Exception as e:  # pragma: no cover
        return {"available": False, "reason": repr(e)}


def time_port(batch, threads, seconds=15.0, min_passes=2):
    """The oracle port (plain C, pthreads) on all `threads` host threads: passes over the batch for ~`seconds`."""
    import oracle

    b64 = batch.astype(np.float64)
    oracle.plan_batch(b64.slice(0, min(256, b64.n_frames)), threads=threads)  # warm (library load, initial path)
    t0 = time.perf_counter()
    passes = 0
    while passes < min_passes or time.perf_counter() - t0 < seconds:
        oracle.plan_batch(b64, threads=threads)
        passes += 1
    dt = time.perf_counter() - t0
    return passes * b64.n_frames / dt, dt, passes


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path on the host cores.  The real reference
    (baseline/_ref, numba + scipy, unmodified) when it can run here -- value = its multiprocessing.Pool(all cores)
    throughput on a >= 2 048-frame sample of the workload, fresh planner per frame; otherwise
    the oracle port with the reason recorded.  The port is timed next to it in both cases."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ft_fsd_path_planning_b200 import synth

    threads = os.cpu_count() or 1
    sample_frames = max(2048, 128 * threads)
    ref = reference_numba(frames=sample_frames, single=512)
    port_batch = synth.gen_autocross(SEED, 2048, workers=min(threads, 16))
    port_value, port_dt, port_passes = time_port(port_batch, threads, seconds=5.0)
    port = {"value": port_value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{port_passes} passes over frames 0..2047 of the workload ({port_dt:.1f} s), oracle/*.c, {threads} pthreads"}
    if ref.get("available"):
        value = ref["pool"]["frames_per_s"]
        ms = 1e3 * ref["pool"]["seconds_slowest_worker"]
        cpu = {"value": value, "unit": UNIT, "cores": ref["pool"]["workers"], "kind": "reference",
               "sample": f"frames 0..{ref['pool']['frames'] - 1} of the workload, fresh PathPlanner per frame, "
                         f"multiprocessing.Pool({ref['pool']['workers']}), slowest worker {ref['pool']['seconds_slowest_worker']:.1f} s; "
                         f"JIT warm-up ({ref['jit_warmup_s']} s) excluded",
               "source": ref["source"], "single_process": ref["single"], "parity_vs_port": ref.get("parity"), "port": port}
        steps_note = "one bounded sample (the steps/warmup flags do not apply to the numba reference: ~8 ms per frame per core)"
    else:
        value, ms = port_value, 1e3 * 2048 / port_value
        cpu = dict(port, reference_numba={"available": False, "reason": ref.get("reason")})
        steps_note = "oracle port (the real reference could not run: see cpu_baseline.reference_numba.reason)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": bench_config(args.gpus), "note": steps_note,
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


class _RehearsalEvent:
    def __init__(self, enable_timing=True):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


class _RehearsalSampler:
    """Stands in for ClockSampler: reports one more "sample" every (rank + 1) polls, so that the untimed sampling loop
    runs a different number of iterations on every rank - as it does with real nvidia-smi timing."""

    def __init__(self, rank):
        self.rank, self.polls = rank, 0

    def start(self):
        pass

    def n_samples(self):
        self.polls += 1
        return self.polls // (self.rank + 1)

    def stop(self):
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["rehearsal"], "samples": 0}


class _RehearsalPlanner:
    """Stands in for BatchPlanner in the CPU rehearsal: returns zeros, launches nothing."""

    def __init__(self):
        import types

        self._events = 0
        self._pinned = {}
        self.lib = types.SimpleNamespace(fsd_plan_launches=lambda b: 4)

    def first_chunk(self, B):
        return (B // 2 + 7) // 8 * 8 if B >= 16 else B

    def plan(self, xy, ty, off, pos, dr, *, kernel_events=False, out=None, chunk_ready=None, intermediates=False, **k):
        import torch
        import types

        self._events += int(kernel_events)
        n = off.numel() - 1
        if chunk_ready is not None:
            chunk_ready.record()
        if out is not None:
            return out
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt)
        return types.SimpleNamespace(path=z(n, 40, 4), status=z(n, dt=torch.int32), left_idx=z(n, 12, dt=torch.int16),
                                     right_idx=z(n, 12, dt=torch.int16), path_f64=None)

    def plan_pinned(self, *a, **k):
        import torch

        self._pinned["path"] = torch.zeros((a[2].numel() - 1, 40, 4))

    def kernel_times_ms(self):
        n, self._events = self._events, 0
        return [(1.0, 1.0)] * max(n, 1)


def parity_vs_oracle(planner, batch, dev_args, timed_path, threads):
    """The timed batch against the oracle (default grid rule on both sides): counts of mismatching frames, never
    averages.  `timed_path` is the output tensor of the last timed step: it must be bit-identical to this call's."""
    import oracle
    import torch

    B = batch.n_frames
    ref = oracle.plan_batch(batch.astype(np.float64), threads=threads)
    res = planner.plan(*dev_args, intermediates=True)
    torch.cuda.synchronize()
    g = lambda t: t.cpu().numpy()
    li, ri, dbg, grid, st = g(res.left_idx), g(res.right_idx), g(res.sort_dbg), g(res.grid), g(res.status).astype(np.uint32)
    path = g(res.path)
    same_P = grid[:, 0] == ref["P"]
    err = np.abs(path.astype(np.float64) - ref["path"]).reshape(B, -1).max(1)
    # Frames above the tolerance: is the frame ill-posed for the ORACLE itself?  Some frames sit on a decision boundary of
    # the smoothing-spline fit (SURVEY Q13 is one family): perturbing the input coordinates by one unit in the last place
    # flips the oracle's own path by millimetres.  Such a frame is counted as a tie when the CUDA path is within the
    # tolerance of one of the oracle's answers on the perturbed inputs; anything else is a real mismatch.
    above = np.nonzero(same_P & (err > 1e-4))[0]
    ties, rng = [], np.random.default_rng(0)
    for b in above[:16]:
        one = batch.slice(int(b), int(b) + 1).astype(np.float64)
        for _ in range(64):
            jit = type(one)(one.cones_xy * (1.0 + rng.normal(0.0, 2e-16, one.cones_xy.shape)), one.cones_type, one.offsets,
                            one.pos, one.dir)
            alt = oracle.plan_batch(jit, threads=1)
            if alt["P"][0] == grid[b, 0] and np.abs(path[b].astype(np.float64) - alt["path"][0]).max() <= 1e-4:
                ties.append(int(b))
                break
    return {
        "frames": B,
        "sort_idx_mismatch": int(((li != ref["left_idx"]).any(1) | (ri != ref["right_idx"]).any(1)).sum()),
        "seeds_mismatch": int((dbg[:, :4] != ref["first_k"].reshape(B, 4)).any(1).sum()),
        "n_configs_or_pops_mismatch": int(((dbg[:, 4:6] != ref["n_configs"]).any(1) |
                                           (dbg[:, 6:8] != np.minimum(ref["n_pops"], 32767)).any(1)).sum()),
        "match_mismatch": int(((g(res.l2r) != ref["l2r"]).any(1) | (g(res.r2l) != ref["r2l"]).any(1) |
                               (g(res.n_wv)[:, 0] != ref["n_left_wv"]) | (g(res.n_wv)[:, 1] != ref["n_right_wv"])).sum()),
        "status_mismatch": int(((st & 0xFFFFFF7F) != (ref["status"] & 0xFFFFFF7F)).sum()),
        "grid_P_mismatch": int((~same_P).sum()),
        "path_max_err": float(err[same_P].max()) if same_P.any() else None,
        "path_frames_above_1e-4": int(len(above)),
        "of_which_oracle_ties": len(ties),  # the oracle itself gives the CUDA answer when its inputs move by 1 ulp
        "path_frames_mismatching": int(len(above) - len(ties)),
        "timed_output_identical": bool(torch.equal(res.path, timed_path)),
        "frames_with_2plus_configs": int((ref["n_configs"] >= 2).any(1).sum()),
        "flagged": int(((ref["status"] & 0x700) != 0).sum()),
    }


def run_ours(args):
    import torch
    import torch.distributed as dist

    from ft_fsd_path_planning_b200 import synth
    from ft_fsd_path_planning_b200.distributed import GatherPipeline, PeerGather, shard_parts

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if REHEARSAL:
        dev = torch.device("cpu")
        new_event = _RehearsalEvent
        sync = lambda: None
    else:
        from ft_fsd_path_planning_b200 import BatchPlanner

        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
        new_event = lambda: torch.cuda.Event(enable_timing=True)
        sync = lambda: torch.cuda.synchronize(dev)
    if distributed:
        import datetime

        # a collective that does not complete within 3 minutes aborts the run instead of hanging the box
        if REHEARSAL:
            dist.init_process_group("gloo", timeout=datetime.timedelta(seconds=180))
        else:
            dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    def barrier():
        if distributed:
            dist.barrier()
        sync()

    cores = os.cpu_count() or 1
    host_threads = max(1, cores // world)
    planner = _RehearsalPlanner() if REHEARSAL else BatchPlanner(dev)
    pin = (lambda a: torch.from_numpy(np.ascontiguousarray(a))) if REHEARSAL else \
        (lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory())

    # N > 1: the all-gather of the paths is fused into the path kernel (stores into every peer's gathered buffer over
    # NVLink peer memory / NVSwitch multicast, distributed.PeerGather); --gather nccl (or a box without symmetric
    # memory) keeps the NCCL all-gather pipeline.  All ranks must agree, hence the all-reduce of the outcome.
    p2p = {"on": False, "why": "--gather nccl" if args.gather == "nccl" else ""}
    if distributed and not REHEARSAL and args.gather != "nccl":
        ok = 1
        try:
            probe = PeerGather(world, dev, buffers=1, multicast=args.gather != "p2p-unicast")
            probe.finish()
            sync()
            del probe
        except Exception as exc:  # noqa: BLE001 - any failure means: fall back to NCCL, and say why
            ok, p2p["why"] = 0, f"symmetric memory unavailable: {type(exc).__name__}: {exc}"[:300]
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        p2p["on"] = bool(flag.item())
        if not p2p["on"] and not p2p["why"]:
            p2p["why"] = "symmetric memory unavailable on another rank"

    class Workload:
        """One rank's share of a global batch: the frames of this rank's blocks (shard_parts layout: the planner's first
        chunk and the rest, so that the all-gather of the first chunk can start while the rest is planned)."""

        def __init__(self, kind, seed, per_gpu):
            na = planner.first_chunk(per_gpu)
            self.sizes = [na, per_gpu - na] if 0 < na < per_gpu and not p2p["on"] else [per_gpu]
            self.n_global = per_gpu * world
            gen = synth.gen_mixed if kind == "mixed" else synth.gen_autocross
            blocks = shard_parts(self.sizes, rank, world)
            self.batch = synth.concat_batches([gen(seed, hi - lo, start=lo, workers=min(host_threads, 16)) for lo, hi in blocks])
            b = self.batch
            self.dev_args = tuple(torch.from_numpy(a).to(dev) for a in (b.cones_xy, b.cones_type, b.offsets, b.pos, b.dir))
            self.peer = PeerGather(self.n_global, dev, multicast=args.gather != "p2p-unicast") if p2p["on"] else None
            self.first_row = rank * per_gpu
            self.pipe = GatherPipeline(self.n_global, (40, 4), torch.float32, dev, part_sizes=self.sizes) \
                if distributed and self.peer is None else None
            self.bar_ev = (new_event(), new_event()) if self.peer is not None else None
            self.ready = None if (REHEARSAL and not distributed) else (new_event() if REHEARSAL else torch.cuda.Event())
            self.res = None

        def step(self, events=False):
            """plan + (N > 1) the all-gathers of the paths; everything is ordered on the current stream on return."""
            if not distributed:
                if events:
                    return planner.plan(*self.dev_args, kernel_events=True).path
                self.res = planner.plan(*self.dev_args, out=self.res)
                return self.res.path
            if self.peer is not None:
                if events:  # kernel timing pass: the stage entry points, no peer stores
                    return planner.plan(*self.dev_args, kernel_events=True).path
                # the path kernel stores every frame's path into all ranks' gathered buffers; what is left of the
                # "all-gather" is the cross-GPU barrier in finish()
                self.res = planner.plan(*self.dev_args, out=self.res, gather=self.peer.descriptor(self.first_row))
                self.bar_ev[0].record()
                full = self.peer.finish()
                self.bar_ev[1].record()
                return full
            if events:
                res = planner.plan(*self.dev_args, kernel_events=True)
            elif len(self.sizes) == 1:
                self.res = res = planner.plan(*self.dev_args, out=self.res)
            else:
                self.res = res = planner.plan(*self.dev_args, out=self.res, chunk_ready=self.ready)
            off = 0
            for p, sz in enumerate(self.sizes):
                # the first chunk's gather starts when the planner's chunk-ready event fires (the second chunk is still
                # being planned); a single-chunk plan is gathered when it is finished
                self.pipe.gather(p, res.path[off:off + sz], after=self.ready if (p == 0 and not events and len(self.sizes) > 1) else None)
                off += sz
            return self.pipe.finish()

        def gather_ms(self):
            """Device time of the last step's collective part: the NCCL gathers, or the symmetric-memory barrier."""
            return self.bar_ev[0].elapsed_time(self.bar_ev[1]) if self.peer is not None else self.pipe.gather_ms()

    main = Workload("autocross", SEED, FRAMES_PER_GPU)
    batch, B, n_global = main.batch, main.batch.n_frames, main.n_global
    flush = torch.empty(1024 if REHEARSAL else 256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def timed_steps(w, steps):
        """K steps, CUDA events around each (L2 flushed in between, outside the events); barrier + sync on both sides."""
        barrier()
        evs, gather_ms = [], []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = new_event(), new_event()
            e0.record()
            w.step()
            e1.record()
            evs.append((e0, e1))
            if distributed and not REHEARSAL and args.comm_detail:
                sync()
                gather_ms.append(w.gather_ms())
        barrier()
        return [a.elapsed_time(b) for a, b in evs], gather_ms

    # clocks / throttle reasons are sampled (rank 0, all GPUs) from the warm-up to the end of the timed steps (the timed
    # region itself lasts ~0.1 s, a handful of nvidia-smi periods); identical untimed steps follow until enough samples
    sampler = _RehearsalSampler(rank) if REHEARSAL else (ClockSampler(range(world)) if rank == 0 else None)
    if sampler:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        main.step()
        main.step(events=True)
    barrier()
    planner.kernel_times_ms()  # drop the warm-up events

    # ---- device-resident throughput: K steps, CUDA events per step, L2 flushed between steps -------------------------
    step_times, _ = timed_steps(main, args.steps)
    step_ms = float(np.mean(step_times))
    timed_path = main.res.path if main.res is not None else None
    # Same kernels, untimed, only to give nvidia-smi time to report.  Rank-local on purpose: the number of iterations
    # differs from rank to rank, so nothing in this loop may be a collective (no all-gather, no barrier).
    t_end = time.time() + (0.3 if REHEARSAL else 1.5)
    while time.time() < t_end and (sampler is None or sampler.n_samples() < 8):
        flush.zero_()
        planner.plan(*main.dev_args)
        sync()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed steps (+ identical untimed steps), sampled by rank 0 for all GPUs of the run"
    # ---- per-kernel launch durations (roofline): the two stage entry points, whole batch per launch, events between --
    for _ in range(args.steps):
        flush.zero_()
        main.step(events=True)
    barrier()
    ktimes = planner.kernel_times_ms()
    sort_ms = float(np.mean([t[0] for t in ktimes]))
    path_ms = float(np.mean([t[1] for t in ktimes]))
    # ---- the cost-matrix step in isolation: fsd_knn_batch over the whole batch, events per launch ---------------
    knn_ms = 0.0
    if not REHEARSAL:
        knn_out = planner.knn(*main.dev_args[:3])
        kev = []
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = new_event(), new_event()
            e0.record()
            planner.knn(*main.dev_args[:3], out=knn_out)
            e1.record()
            kev.append((e0, e1))
        sync()
        knn_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    # ---- communication detail (N > 1): the same steps again with a sync per step to read the gather events -------------
    comm = None
    if distributed:
        args.comm_detail = True
        detail_times, gather_ms = timed_steps(main, max(3, args.steps // 2))
        args.comm_detail = False
        comm = {"gather_ms": float(np.mean(gather_ms)) if gather_ms else None,
                "step_ms_with_sync_per_step": float(np.mean(detail_times))}

    # ---- end to end through the public API: pinned host buffers, H2D + plan + D2H inside the timed region -----
    h_xy, h_ty, h_off, h_pos, h_dir = pin(batch.cones_xy), pin(batch.cones_type), pin(batch.offsets), pin(batch.pos), pin(batch.dir)
    h_path = pin(np.empty((B, 40, 4), dtype=np.float32))
    h_li = pin(np.empty((B, 12), dtype=np.int16))
    h_ri = pin(np.empty((B, 12), dtype=np.int16))
    h_st = pin(np.empty((B,), dtype=np.int32))

    def e2e_step():
        # the host-to-host entry point: per chunk H2D of the inputs and the sort + match launches on streams of their own
        # (copies overlap kernels), the path stage over the whole batch storing the paths straight into the pinned host
        # buffer (zero copy; N > 1: into a device buffer that is all-gathered, plus a D2H copy), D2H of sort indices / status
        if distributed and main.peer is not None:
            planner.plan_pinned(h_xy, h_ty, h_off, h_pos, h_dir, h_path, h_li, h_ri, h_st,
                                gather=main.peer.descriptor(main.first_row))
            main.peer.finish()
            return
        planner.plan_pinned(h_xy, h_ty, h_off, h_pos, h_dir, h_path, h_li, h_ri, h_st, zero_copy=not distributed)
        if distributed:
            off = 0
            for p, sz in enumerate(main.sizes):
                main.pipe.gather(p, planner._pinned["path"][off:off + sz])
                off += sz
            main.pipe.finish()

    for _ in range(3):
        e2e_step()
    barrier()
    e2e_evs = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = new_event(), new_event()
        e0.record()
        e2e_step()
        e1.record()
        e2e_evs.append((e0, e1))
    barrier()
    e2e_ms = float(np.mean([a.elapsed_time(b) for a, b in e2e_evs]))
    h2d = sum(t.numel() * t.element_size() for t in (h_xy, h_ty, h_off, h_pos, h_dir))
    d2h = sum(t.numel() * t.element_size() for t in (h_path, h_li, h_ri, h_st))

    # ---- the named BASELINE config 5 shard (8 192 mixed frames per GPU), same step -----------------------------------
    c5 = Workload("mixed", 5, CONFIG5_PER_GPU)
    for _ in range(3):
        c5.step()
    c5_times, _ = timed_steps(c5, max(5, args.steps // 2))
    c5_ms = float(np.mean(c5_times))

    # ---- parity of the timed batches against the oracle (every rank its shard) ------------------------------------------
    parity = parity5 = None
    if not REHEARSAL:
        parity = parity_vs_oracle(planner, batch, main.dev_args, timed_path, host_threads)
        parity5 = parity_vs_oracle(planner, c5.batch, c5.dev_args, c5.res.path if c5.res is not None else
                                   planner.plan(*c5.dev_args).path, host_threads)

    # ---- the gathered buffer of one more step: this rank's rows equal its local output, all ranks hold the same bytes ----
    gathered_ok = None
    if distributed and not REHEARSAL:
        full = main.step()
        sync()
        lo = rank * B
        own = bool(torch.equal(full[lo:lo + B], main.res.path)) if main.peer is not None else True
        digest = full.view(torch.int32).to(torch.int64).sum().reshape(1)
        digests = torch.empty((world,), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(digests, digest)
        same = bool((digests == digests[0]).all().item())
        flag = torch.tensor([int(own and same)], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gathered_ok = bool(flag.item())

    # ---- max over ranks; per-rank numbers for the comm block ----------------------------------------------------------
    mine = [step_ms, e2e_ms, sort_ms, path_ms, c5_ms, comm["gather_ms"] or 0.0 if comm else 0.0,
            comm["step_ms_with_sync_per_step"] if comm else 0.0, knn_ms]
    per_rank = None
    if distributed:
        t = torch.tensor(mine, dtype=torch.float64, device=dev)
        allr = torch.empty((world * len(mine),), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, t)
        per_rank = allr.cpu().numpy().reshape(world, len(mine))
        step_ms, e2e_ms, sort_ms, path_ms, c5_ms = (float(v) for v in per_rank[:, :5].max(0))
        knn_ms = float(per_rank[:, 7].max())
        if parity is not None:
            keys = [k for k, v in parity.items() if isinstance(v, int) and not isinstance(v, bool)]
            for par in (parity, parity5):
                cnt = torch.tensor([par[k] for k in keys] + [int(par["timed_output_identical"])], dtype=torch.int64, device=dev)
                mx = torch.tensor([par["path_max_err"] or 0.0], dtype=torch.float64, device=dev)
                dist.all_reduce(cnt)
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                for k, v in zip(keys, cnt.tolist()):
                    par[k] = int(v)
                par["timed_output_identical"] = int(cnt[-1].item()) == world
                par["path_max_err"] = float(mx.item())

    if rank == 0 and REHEARSAL:
        # the rehearsal planned nothing: no number may leave it
        print(json.dumps({"rehearsal": True, "value": None, "n_gpus": world, "steps": args.steps,
                          "note": "control-flow rehearsal on CPU (gloo, planner stub): no planner work was done"}))
    elif rank == 0:
        value = n_global / (step_ms * 1e-3)
        peak, peak_src = peaks()
        alg_bytes = batch.algorithmic_bytes()  # per launch of this rank's shard: 9 B/cone + 708 B/frame
        dom_ms, dom_name = (path_ms, "path_kernel") if path_ms >= sort_ms else (sort_ms, "sort_match_kernel")
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        traffic, issue = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            t = json.load(open(tpath)).get(dom_name)
            if t:
                traffic = t["read_bytes"] + t["write_bytes"]  # per launch, from the committed ncu --set full capture
                if "warp_inst" in t:
                    # what actually bounds the kernel: warp instructions issued per SM cycle (4 schedulers per SM).
                    # Instruction count from the committed ncu capture of the same launch, time measured live here.
                    sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
                    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
                    ipc = t["warp_inst"] / (dom_ms * 1e-3 * sm_clock * n_sm)
                    issue = {"warp_inst_per_launch": t["warp_inst"], "warp_inst_per_frame": t["warp_inst"] / B,
                             "achieved_ipc_per_sm": ipc, "peak_ipc_per_sm": 4.0, "frac": ipc / 4.0,
                             "source": "profiles/ncu_traffic.json (ncu inst_executed) / live kernel time",
                             "capture_build": json.load(open(tpath)).get("build")}
        # the cost-matrix step: SURVEY 8d's B_cm = 19.25 N + 16 bytes per frame (9 N + 16 read, 10 N of k-NN lists + two
        # N/8-byte masks written); the kernel actually writes 12 N (adjacency lists + degrees)
        cm_bytes = 19.25 * batch.total_cones + 16 * B
        cm_gbs = cm_bytes / (knn_ms * 1e-3) / 1e9 if knn_ms > 0 else None
        cm_ncu = (json.load(open(tpath)).get("knn_kernel") or {}) if os.path.exists(tpath) else {}
        cost_matrix = {"kernel": "knn_kernel (fsd_knn_batch)", "ms": knn_ms, "frames_per_s": B / (knn_ms * 1e-3) if knn_ms > 0 else None,
                       "algorithmic_bytes_per_launch": cm_bytes, "achieved": cm_gbs, "unit": "GB/s",
                       "frac": cm_gbs / peak if cm_gbs else None,
                       "fp64_pipe_pct": cm_ncu.get("fp64_pipe_pct"), "issue_active_pct": cm_ncu.get("issue_active_pct"),
                       "bound": "CUDA-core pipe, not HBM: ~N^2 x 2 sides x ~10 instructions per frame against 19.25 N bytes "
                                "(~100 flop/B, ridge ~12); SURVEY 8d puts the pipe bound at ~12 % of HBM for N ~ 80"}
        cfg = bench_config(world)
        cfg.update({"cones_per_frame_mean": batch.total_cones / B,
                    "parallelism": (f"frames partitioned over {world} GPU(s), {main.sizes} frames per rank and planner chunk; "
                                    + ("the all-gather of the paths is fused into the path kernel (peer stores)" if distributed and main.peer is not None else
                                       "the paths are all-gathered on a communication stream (the first chunk's gather "
                                       "starts at the planner's chunk-ready event when the planner splits the batch)"))
                    if distributed else "single GPU"})
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "issue": issue, "kernel": dom_name, "kernel_ms": dom_ms,
                         "other_kernel_ms": sort_ms if dom_name == "path_kernel" else path_ms,
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src, "cost_matrix": cost_matrix,
                         "note": "latency/issue-bound integer+fp64 work: the HBM fraction is reported as required, "
                                 "see DESIGN.md"},
            "parity": parity,
            "config5_shard": {"workload": f"gen_mixed(seed=5): {CONFIG5_PER_GPU} frames per GPU, {CONFIG5_PER_GPU * world} "
                                          f"global (BASELINE configs[4] is this at 8 GPUs)", "ms_per_step": c5_ms,
                              "value": CONFIG5_PER_GPU * world / (c5_ms * 1e-3), "unit": UNIT, "parity": parity5},
            "e2e": {"value": n_global / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
            "gpu_launches": int(planner.lib.fsd_plan_launches(B)) * args.steps,
            "clocks": clocks,
        }
        if distributed:
            names = ["step_ms", "e2e_ms", "sort_kernel_ms", "path_kernel_ms", "config5_step_ms", "gather_ms",
                     "step_ms_with_sync_per_step", "knn_kernel_ms"]
            out["comm"] = {
                "collective": (f"none: the path kernel stores every frame's fp32 path ({4 * 160 * B} bytes per rank and step) "
                               f"into all {world} ranks' gathered buffers ("
                               + ("one multimem store per value through the NVSwitch multicast address" if main.peer.multicast
                                  else "plain stores through peer-mapped pointers over NVLink")
                               + "; torch symmetric memory), then ONE symmetric-memory barrier per step") if main.peer is not None else
                              (f"{len(main.sizes)} x all_gather_into_tensor of the fp32 paths per step (NCCL), "
                               + " + ".join(str(4 * 160 * sz) for sz in main.sizes) + " bytes per rank"
                               + (f" [{p2p['why']}]" if p2p["why"] else "")),
                "fused_peer_stores": main.peer is not None,
                "gathered_buffer_identical_on_all_ranks": gathered_ok,
                "gather_ms_max_over_ranks": float(per_rank[:, 5].max()),
                "gather_share_of_step": float(per_rank[:, 5].max() / step_ms),
                "per_rank": {n: [float(v) for v in per_rank[:, i]] for i, n in enumerate(names)},
                "step_ms_min_median_max": [float(per_rank[:, 0].min()), float(np.median(per_rank[:, 0])), float(per_rank[:, 0].max())],
                "note": "gather_ms = device time of the collective part of a step (events): the NCCL gathers on the "
                        "communication stream, or -- with fused peer stores -- the symmetric-memory barrier that waits for "
                        "the slowest rank; measured in a separate pass with one synchronisation per step",
            }
        else:
            # ---- the CPU beside it (rank 0, N = 1 only): the oracle port on all host threads, and the real reference ---
            cpu_value, cpu_dt, cpu_passes = time_port(batch, cores, seconds=12.0)
            out["cpu_baseline"] = {
                "value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{cpu_passes} passes over the same {B}-frame batch ({cpu_dt:.1f} s wall), oracle/*.c with {cores} pthreads",
                "reference_numba": reference_numba(frames=max(2048, 128 * cores), single=512) if args.reference_numba else
                {"available": False, "reason": "skipped (--no-reference-numba)"}}
        print(json.dumps(out))
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gather", default="p2p", choices=["p2p", "p2p-unicast", "nccl"],
                    help="N > 1: p2p = all-gather fused into the path kernel (peer stores, multicast when available), "
                         "p2p-unicast = the same without multicast, nccl = all_gather_into_tensor on a side stream")
    ap.add_argument("--no-reference-numba", dest="reference_numba", action="store_false",
                    help="skip the live timing of the numba reference in the cpu_baseline block (saves ~1-2 minutes)")
    args = ap.parse_args()
    args.comm_detail = False
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
