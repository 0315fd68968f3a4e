"""small driver for ncu: a few decode steps at a given batch / context (synthetic KV)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwen_inference_engine_b200 as q  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--arch", default="qwen2.5-0.5b")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--ctx", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--graph", type=int, default=0)
ap.add_argument("--fast", type=int, default=0)
a = ap.parse_args()
cfg = q.make_config(a.arch)
kw = {}
if a.fast:
    kw["numerics"] = "fast"
eng = q.Engine(synthetic=a.arch, kv_bytes=a.batch * (a.ctx + 64) * q.kv_bytes_per_pos(cfg) + (64 << 20),
               max_seqs=a.batch + 1, max_batch_tokens=max(a.batch, 64), use_graph=bool(a.graph), **kw)
seqs = []
for i in range(a.batch):
    s = eng.new_sequence()
    eng.fill_synthetic(s, a.ctx, seed=i)
    seqs.append(s)
tok = np.arange(a.batch, dtype=np.int32) + 5
for _ in range(a.steps):
    tok = eng.decode_step(seqs, tok)
print("tokens", tok[:4])
eng.close()
