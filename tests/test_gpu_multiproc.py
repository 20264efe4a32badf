"""The strip protocol between PROCESSES (tools/dist_parity.py under torchrun): CUDA-IPC mailboxes, peer stores for
halo rows and migrating agents, epoch flags polled by bounded waits, one pf_strip_tick call per tick — against the
unsharded run on one GPU, bit for bit (field and every agent by gid)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _torchrun(nproc, port, *tool_args, timeout=420):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "tools" / "dist_parity.py"), *tool_args]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=str(ROOT))
    return p.returncode, p.stdout + p.stderr


@pytest.mark.timeout(600)
def test_two_processes_on_one_gpu_mailbox_protocol_bitexact():
    # both ranks on cuda:0 (gloo carries the set-up; NCCL refuses two ranks on one device): the cross-process
    # protocol itself — IPC mappings, remote flag words, bounded waits — on a box with a single GPU
    rc, out = _torchrun(2, 29671, "--peer", "--same-device", "--ticks", "40", "--grid", "1024", "--agents", "60000",
                       "--host-ticks", "3")   # + the host-facing tick of a strip (sharded.StripHostTick)
    assert rc == 0 and "-> PASS" in out, out[-3000:]
    assert "pf_strip_tick + mailbox migration" in out and "3 host-facing ticks" in out


@pytest.mark.timeout(900)
def test_nccl_ranks_on_separate_gpus_bitexact():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run by tools/dist_parity.py under gpurun --gpus N; outputs in profiles/)")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    rc, out = _torchrun(world, 29673, "--peer", "--ticks", "100", "--grid", "4096", "--agents", "1200000",
                        "--host-ticks", "3")
    assert rc == 0 and "-> PASS" in out, out[-3000:]
