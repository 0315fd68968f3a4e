"""persistent decode kernel probe: tokens/s at batch B and the per-phase time breakdown."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwen_inference_engine_b200 as q  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--arch", default="qwen2.5-0.5b")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--ctx", type=int, default=32)
ap.add_argument("--steps", type=int, default=128)
ap.add_argument("--mega", type=int, default=1)
ap.add_argument("--graph", type=int, default=1)
ap.add_argument("--fast", type=int, default=0)
ap.add_argument("--page-size", type=int, default=16)
a = ap.parse_args()
cfg = q.make_config(a.arch)
eng = q.Engine(synthetic=a.arch, kv_bytes=a.batch * (a.ctx + a.steps + 256) * q.kv_bytes_per_pos(cfg) + (64 << 20),
               max_seqs=a.batch + 1, max_batch_tokens=max(a.batch, 64), use_graph=bool(a.graph), page_size=a.page_size,
               numerics="fast" if a.fast else "reference_order")
eng.set_int("mega", a.mega)
seqs = []
for i in range(a.batch):
    s = eng.new_sequence()
    eng.fill_synthetic(s, a.ctx, seed=i)
    seqs.append(s)
tok = np.arange(a.batch, dtype=np.int32) + 5
print("uses_mega", eng.uses_mega(a.batch, a.ctx + a.steps))
out = eng.decode_run(seqs, tok, 8)  # warm-up (+ graph capture)
eng.sync()
t0 = time.perf_counter()
out = eng.decode_run(seqs, out[-1], a.steps)
eng.sync()
dt = time.perf_counter() - t0
wb = q.weight_bytes(cfg)
print(f"batch {a.batch} ctx {a.ctx}: {a.steps} steps in {dt*1e3:.2f} ms -> {dt/a.steps*1e6:.1f} us/step, "
      f"{a.batch*a.steps/dt:.1f} tok/s, weights-only HBM {wb/(dt/a.steps)/1e9:.0f} GB/s")
print("tokens", out[-1][:4])
if a.mega and eng.uses_mega(a.batch, a.ctx + a.steps + 16):
    eng.set_int("mega_prof", 1)
    eng.decode_run(seqs, out[-1], 3)
    ts, cyc = eng.mega_prof()
    ts, cyc = ts.astype(np.int64), cyc.astype(np.int64)
    print("SM clock during the step: %.0f MHz" % ((cyc[-1] - cyc[0]) / max(1, ts[-1] - ts[0]) * 1e3))
    L = cfg.layers
    gc = eng.mega_gemm_cycles.astype(np.int64).reshape(5, 2)
    # (the GEMV kernel of fast numerics at <= 4 rows fills the two cycle tables only when built with make GVX=-DGV_PROFILE_DETAIL)
    print("CTA0 warp0 cycles (wait for weights, MMA loop) per step:", {k: (int(gc[i, 0]), int(gc[i, 1])) for i, k in enumerate(["qkv", "o", "gateup", "down", "lm_head"])})
    print("CTA0 attention cycles per step (setup, scores, softmax, PV):", [int(x) for x in eng.mega_attn_cycles])
    if a.fast and a.batch <= 4 and cfg.layers > 5 and cyc[8:8 + 148].any():  # only a -DGV_PROFILE_DETAIL build of decode_gemv.cu stamps these
        g_end, g_start = cyc[8:8 + 148], cyc[156:156 + 148]
        t0 = g_start.min()
        print("gate/up of layer 5 over the CTAs (us after the first start): start p50 %.2f max %.2f | end min %.2f p50 %.2f p90 %.2f max %.2f (CTA %d)" % (
            np.median(g_start - t0) / 1e3, (g_start - t0).max() / 1e3, (g_end - t0).min() / 1e3, np.median(g_end - t0) / 1e3,
            np.percentile(g_end - t0, 90) / 1e3, (g_end - t0).max() / 1e3, int(np.argmax(g_end))))
        print("  end per CTA (us):", np.round((g_end - t0) / 1e3, 2).tolist())
    d = np.diff(ts)
    per = d[:16 * L].reshape(L, 16)
    names = ["qkv.load", "qkv.norm", "qkv.gemm", "qkv.bar", "att.run", "att.bar", "o.load", "o.gemm", "o.bar",
             "gu.load", "gu.norm", "gu.gemm", "gu.bar", "dn.load", "dn.gemm", "dn.bar"]
    print("phase us (mean over layers, CTA 0):", {n: round(float(per[1:, i].mean()) / 1e3, 2) for i, n in enumerate(names)})
    hn = ["head.load", "head.norm", "head.gemm", "head.bar", "sample"]
    print("layers total us", round(float(per.sum()) / 1e3, 1), {n: round(float(d[16 * L + i]) / 1e3, 1) for i, n in enumerate(hn)},
          "whole us", round(float(ts[-1] - ts[0]) / 1e3, 1))
eng.close()
