"""gpu (needs >= 2 GPUs, skipped otherwise): tensor-parallel engine over NCCL, launched the way
bench.py is (torchrun, one rank per GPU).  tools/tp_probe.py --check compares the TP engine with
a tp_size 1 engine on rank 0: every rank returns the same tokens, the teacher-forced greedy
choices agree except at near-ties, final logits within the 1e-2 bf16 tolerance (the all-reduce
sums bf16 partial sums)."""
import json
import os
import socket
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("arch,batch", [("small", 3), ("qwen2.5-0.5b", 2)])
def test_tp2_matches_single_gpu(arch, batch):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "tp_probe.py"), "--arch", arch,
           "--batch", str(batch), "--prompt", "12", "--steps", "16", "--warmup", "2", "--check"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["tp"] == 2 and res["ranks_agree"]
    assert res["check_logits_rel_l2_max"] < 1e-2  # BASELINE.json: 1e-2 relative error in bf16
    agree, total = res["check_tokens_agree"]
    assert agree >= 0.8 * total
