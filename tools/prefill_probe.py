"""Prefill probe (BASELINE configs[2]: Qwen2.5-1.5B-arch bf16, prefill 4096 tokens + decode).

  python tools/prefill_probe.py --arch qwen2.5-1.5b --tokens 4096 --numerics fast

Times qie_prefill (HOST ids in, HOST token out) with CUDA events on the engine stream; reports
tokens/s and algorithmic TFLOP/s (2*T*L*params/layer + 2*V*H + L*n_q*4*hd*T^2/2)
against the measured dense bf16 peak, then a short greedy decode.  One JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwen_inference_engine_b200 as q  # noqa: E402


def prefill_flops(cfg, T, prefix=0):
    H, I, L, hd = cfg.hidden, cfg.inter, cfg.layers, cfg.head_dim
    Dq, Dkv = cfg.n_q * hd, cfg.n_kv * hd
    per_layer = H * Dq + 2 * H * Dkv + Dq * H + 3 * H * I
    gemm = 2 * T * L * per_layer + 2 * cfg.vocab * H
    attn = L * cfg.n_q * 4 * hd * (T * T / 2 + T * prefix)
    return gemm, attn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="qwen2.5-1.5b")
    ap.add_argument("--tokens", type=int, default=4096)
    ap.add_argument("--chunk", type=int, default=0, help="rows per forward (0 = the whole prompt)")
    ap.add_argument("--numerics", default="fast")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--decode", type=int, default=32)
    a = ap.parse_args()
    cfg = q.make_config(a.arch, context=max(8192, a.tokens + 256))
    chunk = a.chunk or a.tokens
    eng = q.Engine(synthetic=cfg, seed=1234, max_seqs=2, max_batch_tokens=chunk, kv_bytes=4 << 30,
                   context=cfg.context, numerics=a.numerics)
    rng = np.random.default_rng(5)
    ids = rng.integers(0, cfg.vocab, size=a.tokens, dtype=np.int32)
    st = torch.cuda.ExternalStream(eng.stream)
    times = []
    tok = None
    for r in range(a.warmup + a.reps):
        s = eng.new_sequence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(st)
        tok = eng.prefill(s, ids)
        e1.record(st)
        torch.cuda.synchronize()
        if r >= a.warmup:
            times.append(e0.elapsed_time(e1))
        if r + 1 < a.warmup + a.reps:
            eng.free_sequence(s)
    ms = float(np.median(times))
    gemm, attn = prefill_flops(cfg, a.tokens)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained", 1377.7)
    res = {"arch": a.arch, "tokens": a.tokens, "chunk": chunk, "numerics": a.numerics, "prefill_ms": ms,
           "prefill_ms_all": times, "prefill_tokens_per_s": a.tokens / (ms / 1e3), "alg_tflop": (gemm + attn) / 1e12,
           "achieved_tflops": (gemm + attn) / 1e12 / (ms / 1e3), "peak_tflops": peak,
           "frac_of_tensor_peak": (gemm + attn) / 1e12 / (ms / 1e3) / peak, "launches": eng.launch_count(),
           "first_token": tok}
    if a.decode:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.decode_run([s], [tok], 4)
        torch.cuda.synchronize()
        e0.record(st)
        out = eng.decode_run([s], [tok], a.decode)
        e1.record(st)
        torch.cuda.synchronize()
        dms = e0.elapsed_time(e1) / a.decode
        res["decode_ms_per_token_after_prefill"] = dms
        res["decode_tokens_per_s"] = 1e3 / dms
    print(json.dumps(res), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
