"""Generate the golden fixtures in tests/golden/ by running the REFERENCE'S OWN KERNELS
(oracle/_ref/libqie_ref.so, built from /root/reference/layers/src) on a B200.

The reference ships no golden vectors, so these fixtures are what
pins the oracle: inputs are stored next to the reference outputs, everything is seeded.
Run on the GPU box:   python oracle/gen_golden.py gpurun_out/golden
then copy gpurun_out/golden/*.npz to tests/golden/.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path = [p_ for p_ in sys.path if os.path.abspath(p_ or ".") != os.path.dirname(os.path.abspath(__file__))]
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import qwen_inference_engine_b200 as q  # noqa: E402
from oracle.oracle import Ref, RefSeq  # noqa: E402
from util import f32_to_bf16, prompt_ids, rand_bf16, to_dev, to_host  # noqa: E402


def p(t):
    return t.data_ptr()


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    ref = Ref()
    rng = np.random.default_rng(20261018)
    ops = {}

    # rmsNorm
    for name, (hidden, n_tok) in {"rms_a": (896, 3), "rms_b": (128, 5)}.items():
        x, w = rand_bf16(rng, (n_tok, hidden), 0.7), rand_bf16(rng, (hidden,), 1.0)
        xd, wd = to_dev(x), to_dev(w)
        y = torch.zeros_like(xd)
        assert ref.L.ref_rmsnorm(p(xd), p(wd), p(y), hidden, n_tok) == 0
        ops[name + "_x"], ops[name + "_w"], ops[name + "_y"] = x, w, to_host(y)
    # matmul
    for name, (M, N, K) in {"mm_a": (1, 896, 128), "mm_b": (5, 256, 72), "mm_c": (33, 128, 40), "mm_d": (1, 4864, 24)}.items():
        A, B = rand_bf16(rng, (M, N), 1.0), rand_bf16(rng, (K, N), 0.05)
        Ad, Bd = to_dev(A), to_dev(B)
        C = torch.zeros(M, K, dtype=torch.bfloat16, device="cuda")
        assert ref.L.ref_matmul(p(Ad), p(Bd), p(C), M, N, K) == 0
        ops[name + "_A"], ops[name + "_B"], ops[name + "_C"] = A, B, to_host(C)
    # qkNorm + RoPE
    for name, (hd, nh, nt) in {"qk_a": (64, 14, 3), "qk_b": (128, 4, 2)}.items():
        x, w = rand_bf16(rng, (nt, hd * nh), 1.3), rand_bf16(rng, (hd,), 1.0)
        xd, wd = to_dev(x), to_dev(w)
        assert ref.L.ref_qknorm(p(xd), p(wd), hd, nt, hd * nh, nh) == 0
        normed = to_host(xd)
        cos = np.zeros((64, hd // 2), np.float32)
        sin = np.zeros((64, hd // 2), np.float32)
        ref.L.ref_precompute_cos_sin(cos.ctypes.data, sin.ctypes.data, 64, hd)
        cd, sd = torch.from_numpy(cos).cuda(), torch.from_numpy(sin).cuda()
        r1 = xd.clone()
        assert ref.L.ref_rope(p(cd), p(sd), p(r1), nt, hd, hd * nh, nh) == 0
        r2 = xd[:1].clone()
        assert ref.L.ref_rope_single(p(cd), p(sd), p(r2), 37, hd, hd * nh, nh) == 0
        ops.update({name + "_x": x, name + "_w": w, name + "_normed": normed, name + "_cos": cos, name + "_sin": sin,
                    name + "_rope": to_host(r1), name + "_rope37": to_host(r2)})
    # elementwise
    n = 1000
    a, b = rand_bf16(rng, (n,), 2.0), rand_bf16(rng, (n,), 1.0)
    ad, bd = to_dev(a), to_dev(b)
    s_ = ad.clone()
    assert ref.L.ref_act(p(s_), n) == 0
    m_ = torch.zeros_like(ad)
    assert ref.L.ref_elem(p(ad), p(bd), p(m_), n) == 0
    r_ = ad.clone()
    assert ref.L.ref_resadd(p(r_), p(bd), n) == 0
    ops.update({"ew_a": a, "ew_b": b, "ew_silu": to_host(s_), "ew_mul": to_host(m_), "ew_add": to_host(r_)})
    # attention (decode + causal prefill) over the reference's paged layout
    for name, (hd, n_q, n_kv, L, t, mq, causal) in {"att_dec": (64, 14, 2, 2, 21, 1, 0), "att_pre": (64, 4, 2, 2, 11, 11, 1),
                                                    "att_128": (128, 4, 1, 1, 9, 1, 0)}.items():
        Dq, Dkv, layer, ps = n_q * hd, n_kv * hd, L - 1, 4
        K, V, Q = rand_bf16(rng, (t, Dkv), 1.0), rand_bf16(rng, (t, Dkv), 1.0), rand_bf16(rng, (mq, Dq), 1.0)
        n_pages = (t + ps - 1) // ps
        elems = ps * L * Dkv
        pages = ref.L.ref_pages_create(n_pages, elems)
        for pg in range(n_pages):
            kb = np.zeros((ps, L, Dkv), np.uint16)
            vb = np.zeros((ps, L, Dkv), np.uint16)
            rows = K[pg * ps:(pg + 1) * ps]
            kb[:len(rows), layer] = rows
            vb[:len(rows), layer] = V[pg * ps:(pg + 1) * ps]
            ref.L.ref_pages_write(pages, pg, 0, kb.ctypes.data, elems)
            ref.L.ref_pages_write(pages, pg, 1, vb.ctypes.data, elems)
        Qd = to_dev(Q)
        o = torch.zeros_like(Qd)
        qab = 0 if causal else t - 1
        assert ref.L.ref_attn(p(Qd), p(o), mq, t, hd, Dq, Dkv, causal, qab, layer, pages, ps, L) == 0
        ref.L.ref_pages_free(pages)
        ops.update({name + "_K": K, name + "_V": V, name + "_Q": Q, name + "_out": to_host(o),
                    name + "_cfg": np.array([hd, n_q, n_kv, L, t, mq, causal], np.int32)})
    # sampling: heavy ties (tie-break rule) and top-k 50 (XORWOW stream)
    samp = []
    for i, (vocab, levels, k, temp, seed) in enumerate([(151936, 40, 1, 1.0, 1234), (151936, 9, 1, 0.7, 1235), (4096, 3, 1, 1.0, 7),
                                                        (151936, 4000, 50, 0.7, 1240), (3000, 25, 50, 1.0, 99), (777, 5, 7, 0.7, 5),
                                                        (300, 2, 50, 0.7, 1)]):
        vals = rng.choice(np.linspace(-3, 3, levels).astype(np.float32), size=vocab)
        lg = f32_to_bf16(vals)
        tok = ref.L.ref_sample(p(to_dev(lg)), vocab, temp, k, seed, 0)
        ops[f"samp{i}_logits"] = lg
        samp.append([vocab, levels, k, int(round(temp * 10)), seed, tok])
    ops["samp_table"] = np.array(samp, np.int64)
    np.savez_compressed(os.path.join(out_dir, "ops_reference_kernels.npz"), **ops)

    # ---- end to end: reference kernels replaying llm() on synthetic checkpoints
    e2e = {}
    for arch, seed, n_prompt, n_new in [("small", 1234, 32, 128), ("tiny", 4321, 6, 24), ("small128", 77, 9, 40)]:
        eng = q.Engine(synthetic=arch, seed=seed, context=512, max_batch_tokens=64)
        ids = prompt_ids(n_prompt, eng.config.vocab)
        desc = ref.model_desc(eng)
        rs = RefSeq(ref, desc, page_size=4)
        taps = {}
        toks = [rs.prefill(ids, taps=taps)]
        lg0 = rs.read("logits", eng.config.vocab)
        for _ in range(n_new - 1):
            toks.append(rs.decode(toks[-1]))
        lgN = rs.read("logits", eng.config.vocab)
        rs.close()
        rs = RefSeq(ref, desc, page_size=4)
        tk = [rs.prefill(ids, topk=50)]
        for _ in range(23):
            tk.append(rs.decode(tk[-1], topk=50))
        rs.close()
        e2e[f"{arch}_seed"] = np.array([seed], np.int64)
        e2e[f"{arch}_prompt"] = ids
        e2e[f"{arch}_greedy"] = np.array(toks, np.int32)
        e2e[f"{arch}_topk50"] = np.array(tk, np.int32)
        e2e[f"{arch}_logits_prefill"] = lg0
        e2e[f"{arch}_logits_last"] = lgN
        last = eng.config.layers - 1
        for tag in ("input_norm", "q", "attn", "x_attn", "mlp_h", "x_out"):
            e2e[f"{arch}_L0_{tag}"] = taps[(tag, 0)]
            e2e[f"{arch}_L{last}_{tag}"] = taps[(tag, last)]
        eng.close()
    # BASELINE.json configs[0]: 0.5B-arch, 32-token prompt, 128 greedy tokens
    eng = q.Engine(synthetic="qwen2.5-0.5b", seed=1234, max_batch_tokens=64, kv_bytes=256 << 20)
    ids = prompt_ids(32, eng.config.vocab)
    rs = RefSeq(ref, ref.model_desc(eng), page_size=4)
    toks = [rs.prefill(ids)]
    lg0 = rs.read("logits", eng.config.vocab)
    margins = []
    for _ in range(127):
        toks.append(rs.decode(toks[-1]))
        f = np.sort((rs.read("logits", eng.config.vocab).astype(np.uint32) << 16).view(np.float32))
        margins.append(float(f[-1] - f[-2]))
    rs.close()
    eng.close()
    e2e["config1_prompt"] = ids
    e2e["config1_greedy"] = np.array(toks, np.int32)
    e2e["config1_logits_prefill_top"] = np.sort(lg0)[-64:]
    e2e["config1_top2_margin"] = np.array(margins, np.float32)
    np.savez_compressed(os.path.join(out_dir, "e2e_reference_kernels.npz"), **e2e)
    print("golden written to", out_dir, {k: os.path.getsize(os.path.join(out_dir, k)) for k in os.listdir(out_dir)})
    print("config1 tokens[:16]", toks[:16], "min top-2 margin", min(margins), "steps with tie", sum(m == 0 for m in margins))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
