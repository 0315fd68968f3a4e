"""probe for the tcgen05 prefill attention: parity against the mma.sync kernel + timing."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from qwen_inference_engine_b200 import layers  # noqa: E402
from util import bf16_to_f32, rand_bf16, rel_err, rel_l2, to_dev, to_host  # noqa: E402


def run(T, prefix, n_q, n_kv, variant, time_it=False):
    hd = 128
    rng = np.random.default_rng(T + prefix)
    Dq, Dkv = n_q * hd, n_kv * hd
    total = prefix + T
    ps = 16
    pool = layers.KvPool(n_pages=(total + ps - 1) // ps + 2, page_size=ps, n_layers=1, n_kv_heads=n_kv, head_dim=hd, max_seqs=1)
    K, V = rand_bf16(rng, (total, Dkv), 1.0), rand_bf16(rng, (total, Dkv), 1.0)
    pos_all = torch.arange(total, dtype=torch.int32, device="cuda")
    pool.store(0, to_dev(K), to_dev(V), pos_all, torch.zeros(total, dtype=torch.int32, device="cuda"))
    Q = to_dev(rand_bf16(rng, (T, Dq), 1.0))
    pos = torch.arange(prefix, total, dtype=torch.int32, device="cuda")
    slot = torch.zeros(T, dtype=torch.int32, device="cuda")
    o_ref = torch.zeros_like(Q)
    o_new = torch.full_like(Q, float("nan"))
    layers.launch_attn_prefill_fast(Q, o_ref, pool, 0, pos, slot, n_q)
    layers.launch_attn_prefill_tc(Q, o_new, pool, 0, pos, slot, n_q, variant)
    torch.cuda.synchronize()
    a, b = to_host(o_new), to_host(o_ref)
    nan = int(np.isnan(bf16_to_f32(a)).sum())
    print(f"T {T} prefix {prefix} heads {n_q}/{n_kv} variant {variant}: nan {nan} rel_l2 {rel_l2(a, b):.3e} max {rel_err(a, b):.3e}", flush=True)
    if time_it:
        for name, fn in (("mma.sync", lambda: layers.launch_attn_prefill_fast(Q, o_ref, pool, 0, pos, slot, n_q)),
                         ("tcgen05", lambda: layers.launch_attn_prefill_tc(Q, o_new, pool, 0, pos, slot, n_q, variant))):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            fl = n_q * 4 * hd * (T * T / 2 + T * prefix)
            print(f"   {name}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    for variant in (0, 1):
        run(128, 0, 2, 1, variant)
        run(300, 0, 4, 2, variant)
    v = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    run(1000, 77, 4, 2, v)
    run(4096, 0, 12, 2, v, time_it=True)
    run(4096, 0, 28, 4, v, time_it=True)
