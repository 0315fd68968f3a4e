"""gpu (needs >= 2 GPUs, skipped otherwise): tensor-parallel engine over NCCL, launched the way
bench.py is (torchrun, one rank per GPU).  tools/tp_probe.py --check compares the TP engine with
a tp_size 1 engine on rank 0: every rank returns the same tokens and the teacher-forced greedy
choices agree except at near-ties.  The ranks' o_proj / down_proj partial sums are added in fp32
and rounded to bf16 once, so a TP forward differs from the unsharded one only by the fp32 sum
order (a bf16 output moves by one ulp where the sum sits on a rounding boundary).  Random-init
weights amplify such flips layer by layer (the same growth the fast-numerics probe shows:
tools/fast_error_probe.py), so the bounds are: logits of ONE forward of the 3-layer model within
the 1e-2 bf16 tolerance of BASELINE.json; for the 24-layer architecture and after 18 decode steps
of KV history the measured 2-3e-2 is bounded by 6e-2 and the token agreement is asserted."""
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


@pytest.mark.parametrize("arch,batch", [("small", 3), ("qwen2.5-0.5b", 2), ("small", 20),  # 20 rows: one CTA per row normalises
                                        ("small128", 2)])  # 1 kv head on 2 ranks: the kv head is replicated, its q group split (the TP8-on-7B plan)
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
    if os.environ.get("QIE_TP_MEGA", "1") != "0":
        assert res["persistent_kernel"]  # decode steps ran as one launch per rank with the exchange over peer memory
    print(res)
    if arch in ("small", "small128"):
        assert res["check_prefill_logits_rel_l2"] < 1e-2  # BASELINE.json: 1e-2 relative error in bf16
    assert res["check_prefill_logits_rel_l2"] < 6e-2 and res["check_logits_rel_l2_max"] < 6e-2
    agree, total = res["check_tokens_agree"]
    assert agree >= 0.8 * total
