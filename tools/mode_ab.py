"""Developer script (GPU box): A/B of library builds / plan modes on the bench batch.

    python tools/mode_ab.py [--frames 10240] [--kind color|colorless|mixed] tag=ENV1=V1,ENV2=V2 ...

Every variant runs in a child process (the library reads FSD_PLAN_MODE / FSD_LIBFSDPLAN once): parity against the oracle
(sort indices, matches, status, path), then the step time (20 iterations, L2 flushed in between, CUDA events per
iteration, median) and the per-stage times through the stage entry points."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(frames, kind):
    import numpy as np
    import torch

    sys.path.insert(0, ROOT)
    import oracle
    from ft_fsd_path_planning_b200 import BatchPlanner, synth

    cache = f"/tmp/fsd_ab_{kind}_{frames}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        batch = synth.FrameBatch(z["xy"], z["ty"], z["off"], z["pos"], z["dir"])
    else:
        w = min(os.cpu_count() or 1, 16)
        batch = synth.gen_mixed(5, frames, workers=w) if kind == "mixed" else synth.gen_autocross(3 if kind == "colorless" else 2, frames, workers=w)
        if kind == "colorless":
            batch = synth.remove_color_info(batch)
        np.savez(cache, xy=batch.cones_xy, ty=batch.cones_type, off=batch.offsets, pos=batch.pos, dir=batch.dir)
    if os.environ.get("FSD_AB_ORDER"):  # frames re-packed in a given order (work-binning experiments)
        order = np.load(os.environ["FSD_AB_ORDER"])
        batch = synth.concat_batches([batch.slice(int(i), int(i) + 1) for i in order])
    B = batch.n_frames
    ref = oracle.plan_batch(batch.astype(np.float64), threads=os.cpu_count() or 8)
    bp = BatchPlanner("cuda:0")
    dev = bp.device
    args = tuple(torch.from_numpy(a).to(dev) for a in (batch.cones_xy, batch.cones_type, batch.offsets, batch.pos, batch.dir))
    res = bp.plan(*args, intermediates=True)
    torch.cuda.synchronize()
    g = lambda t: t.cpu().numpy()
    st = g(res.status).astype(np.uint32)
    same_P = g(res.grid)[:, 0] == ref["P"]
    err = np.abs(g(res.path).astype(np.float64) - ref["path"]).reshape(B, -1).max(1)
    over = (st & 0x100) != 0
    if os.environ.get("FSD_AB_DUMP_GRID"):  # probe build: grid[:, 1] = path-machine cycles / 256 of every frame
        cyc = g(res.grid)[:, 1].astype(np.int64)
        np.save(os.environ["FSD_AB_DUMP_GRID"], np.argsort(cyc, kind="stable"))
        print(json.dumps({"cycles_x256": {"mean": float(cyc.mean()), "p10": float(np.percentile(cyc, 10)),
                                          "p50": float(np.percentile(cyc, 50)), "p90": float(np.percentile(cyc, 90)),
                                          "max": float(cyc.max())}}))
        return
    par = {"sort": int(((g(res.left_idx) != ref["left_idx"]).any(1) | (g(res.right_idx) != ref["right_idx"]).any(1)).sum()),
           "match": int(((g(res.l2r) != ref["l2r"]).any(1) | (g(res.r2l) != ref["r2l"]).any(1)).sum()),
           "status": int((((st & 0xFFFFFF7F) != (ref["status"] & 0xFFFFFF7F)) & ~over).sum()), "P": int((~same_P).sum()),
           "path>1e-4": int((err[same_P & ~over] > 1e-4).sum()), "path_max": float(err[same_P & ~over].max()), "overflow": int(over.sum())}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = None
    for _ in range(3):
        out = bp.plan(*args, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = bp.plan(*args, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    for _ in range(10):
        flush.zero_()
        bp.plan(*args, kernel_events=True)
    torch.cuda.synchronize()
    kt = bp.kernel_times_ms()
    print(json.dumps({"ms": float(np.median(ts)), "ms_min": float(np.min(ts)), "sort_ms": float(np.median([t[0] for t in kt])),
                      "path_ms": float(np.median([t[1] for t in kt])), "launches": int(bp.lib.fsd_plan_launches(B)), "parity": par}))


def main():
    frames, kind, variants = 10240, "color", []
    a = sys.argv[1:]
    while a:
        x = a.pop(0)
        if x == "--frames":
            frames = int(a.pop(0))
        elif x == "--kind":
            kind = a.pop(0)
        elif x == "--child":
            return child(int(a.pop(0)), a.pop(0))
        else:
            variants.append(x)
    for v in variants:
        tag, _, envs = v.partition("=")
        env = dict(os.environ)
        for kv in filter(None, envs.split(",")):
            k, _, val = kv.partition("=")
            env[k] = val
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(frames), kind], env=env,
                           capture_output=True, text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        print(f"{tag:28s} {line[-1] if line else 'FAILED: ' + r.stderr[-600:]}", flush=True)


if __name__ == "__main__":
    main()
