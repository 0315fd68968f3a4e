"""fast-numerics error probe: per-layer rel-L2 / max-norm error of the FAST engine against the
REFERENCE_ORDER engine on the same weights (prefill of several prompt lengths, then decode)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import qwen_inference_engine_b200 as q  # noqa: E402
from util import prompt_ids, rel_err, rel_l2  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "small"
kw = dict(synthetic=arch, seed=21, context=1024, max_batch_tokens=512, max_seqs=32)
e_ref = q.Engine(numerics="reference_order", **kw)
e_fast = q.Engine(numerics="fast", **kw)
for eng in (e_ref, e_fast):
    eng.capture(True)
L = e_ref.config.layers
sr, sf, cur = [], [], []
for n in (5, 20, 50, 200, 500):
    a, b = e_ref.new_sequence(), e_fast.new_sequence()
    ids = prompt_ids(n, e_ref.config.vocab, seed=n)
    ta, tb = e_ref.prefill(a, ids), e_fast.prefill(b, ids)
    row = [f"prefill {n:4d}: tok {ta == tb}"]
    for l in range(L):
        for tag in ("attn", "mlp_h", "x_out"):
            x, y = e_fast.read_capture(tag, l), e_ref.read_capture(tag, l)
            row.append(f"L{l}.{tag} {rel_l2(x, y):.1e}/{rel_err(x, y):.1e}")
    x, y = e_fast.read_capture("logits", -1), e_ref.read_capture("logits", -1)
    row.append(f"logits {rel_l2(x, y):.1e}/{rel_err(x, y):.1e}")
    print("  ".join(row))
    sr.append(a), sf.append(b), cur.append(ta)
for step in range(4):
    na, nb = e_ref.decode_step(sr, cur), e_fast.decode_step(sf, cur)
    row = [f"decode step {step}: agree {int((na == nb).sum())}/{len(na)}"]
    for l in range(L):
        for tag in ("attn", "mlp_h", "x_out"):
            x, y = e_fast.read_capture(tag, l), e_ref.read_capture(tag, l)
            row.append(f"L{l}.{tag} {rel_l2(x, y):.1e}/{rel_err(x, y):.1e}")
    x, y = e_fast.read_capture("logits", -1), e_ref.read_capture("logits", -1)
    row.append(f"logits {rel_l2(x, y):.1e}/{rel_err(x, y):.1e}")
    print("  ".join(row))
    cur = list(na)
