"""Row-sharded (multi-GPU) parity as driver-run `-m gpu` tests: each test launches `torch.distributed.run` with
min(device_count, 8) ranks on tests/sharded_worker.py and skips on a single-GPU box.  Covered: the peer-window data
plane (default) and the NCCL fallback, speculative row evaluation forced on, the row-chunked tail launch of the Schur
update, row chunks (s > 256), and an UNSEEDED run in which every rank's numpy streams differ (rank 0's draws are
broadcast).  Logs (one JSON line per case) go to gpurun_out/sharded_w<W>.jsonl when that directory exists."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _world():
    import torch
    if not torch.cuda.is_available():
        return 0
    return min(torch.cuda.device_count(), 8)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(suite, env_extra, tag):
    world = _world()
    if world < 2:
        pytest.skip("needs at least 2 GPUs on the box")
    env = dict(os.environ)
    env.update(env_extra)
    env.setdefault("NCCL_DEBUG", "WARN")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "sharded_worker.py"), "--suite", suite]
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        cmd += ["--jsonl", os.path.join(out_dir, "sharded_w%d_%s.jsonl" % (world, tag))]
    res = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, "sharded %s run failed on %d GPUs:\n%s\n%s" % (tag, world, res.stdout[-4000:], res.stderr[-4000:])
    assert '"ok": true' in res.stdout and '"ok": false' not in res.stdout


@pytest.mark.gpu
def test_sharded_replay_peer_windows():
    """Default data plane: NVLink peer windows, Schur epilogue stores + epochs, tail launch publishes the epoch."""
    _launch("replay", {"RPC_B200_PEER": "1", "RPC_B200_SPEC": "auto"}, "peer")


@pytest.mark.gpu
def test_sharded_replay_nccl_fallback():
    """RPC_B200_PEER=0: diag all-gather + exact payload all-reduce per round."""
    _launch("replay", {"RPC_B200_PEER": "0", "RPC_B200_SPEC": "0"}, "nccl")


@pytest.mark.gpu
def test_sharded_replay_speculative_rows():
    """Speculative evaluation of all proposals beside the rejection Cholesky, forced on (default from 4 GPUs)."""
    _launch("replay", {"RPC_B200_PEER": "1", "RPC_B200_SPEC": "1"}, "spec")


@pytest.mark.gpu
def test_sharded_unseeded_ranks_agree():
    """No replay patch, different legacy seeds per rank: rank 0's draws (and its auto-tuned b) are broadcast."""
    _launch("unseeded", {"RPC_B200_PEER": "1"}, "unseeded")
