"""bench.py's multi-rank control flow, rehearsed on CPU: two ranks under torchrun with the gloo backend and a planner
stub (FSD_BENCH_REHEARSAL).  Every collective in bench.py must be reached the same number of times by every rank - the
untimed clock-sampling loop runs a rank-dependent number of iterations (nvidia-smi is absent here, so it always runs
until its deadline) and once contained an all-gather, which hung a 2-GPU run."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(nproc, port):
    env = dict(os.environ, FSD_BENCH_REHEARSAL="64", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", str(nproc), "--steps", "3", "--warmup", "3"]
    return subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240, cwd=ROOT)


def test_two_rank_rehearsal_reaches_the_end():
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = _run(2, port)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    out = json.loads(lines[0])
    assert out["rehearsal"] is True and out["value"] is None and out["n_gpus"] == 2


def test_rehearsal_is_opt_in():
    """Without the environment variable bench.py has no CPU path: it needs a CUDA device."""
    text = open(os.path.join(ROOT, "bench.py")).read()
    assert 'os.environ.get("FSD_BENCH_REHEARSAL", "0")' in text
    assert "planner stub" in text and "no planner work was done" in text
