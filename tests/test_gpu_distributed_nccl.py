"""The multi-GPU region path on real GPUs under NCCL: sharded bootstrap rounds
+ ONE allreduce(MAX) (integrator.py:388-404) and row-sharded inside() (integrator.py:1916-1928),
against the single-process CUDA result AND the oracle.  Needs >= 2 GPUs (skipped otherwise; run
with `gpurun --gpus 2`); the launcher log of the last run is kept in profiles/."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_sharded_rebuild_and_inside(world):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_nccl_worker.py")]
    res = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         timeout=600)
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "nccl_worker_world%d.log" % world), "w") as f:
        f.write(res.stdout + "\n--- stderr ---\n" + res.stderr[-4000:])
    assert res.returncode == 0, res.stderr[-3000:]
    # the ranks share one stdout: lines can arrive glued together, so scan for objects
    reports, dec, text, pos = [], json.JSONDecoder(), res.stdout, 0
    while True:
        pos = text.find('{"rank"', pos)
        if pos < 0:
            break
        obj, end = dec.raw_decode(text, pos)
        reports.append(obj)
        pos = end
    assert sorted(r["rank"] for r in reports) == list(range(world))
    assert all(r["status"] == "ok" and len(r["cases"]) == 3 for r in reports)
    # identical results on all ranks
    for i in range(3):
        assert len({(r["cases"][i]["r2"], r["cases"][i]["f"]) for r in reports}) == 1
