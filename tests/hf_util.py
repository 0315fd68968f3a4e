"""Tiny random-init Hugging Face Qwen2 (Qwen2.5 family: q/k/v bias, no q/k-norm, tied embeddings) and Qwen3 (q/k-norm,
no bias -- the family the reference targets) models, built offline from their configs (no download), saved as bf16
safetensors and converted with the product's own converter: the fixture for the "HF-correct" semantics tests.  The HF
model evaluated in fp32 on the bf16 weights is the ground truth the oracle and the engine are held to (1e-2, the
north-star tolerance)."""
import os

import numpy as np


def make_hf_checkpoint(kind, out_dir, seed=0):
    import torch
    import safetensors.torch as st
    import qwen_inference_engine_b200 as q
    dims = dict(vocab_size=1024, hidden_size=256, intermediate_size=512, num_hidden_layers=3, num_attention_heads=4,
                num_key_value_heads=2, max_position_embeddings=512, rope_theta=1e6, rms_norm_eps=1e-6, attention_dropout=0.0)
    torch.manual_seed(seed)
    if kind == "qwen2":
        from transformers import Qwen2Config, Qwen2ForCausalLM
        model = Qwen2ForCausalLM(Qwen2Config(tie_word_embeddings=True, **dims))
    else:
        from transformers import Qwen3Config, Qwen3ForCausalLM
        model = Qwen3ForCausalLM(Qwen3Config(tie_word_embeddings=False, head_dim=64, **dims))
    with torch.no_grad():  # default init is tiny (std 0.02) and biases are zero: make every term matter
        for n, p in model.named_parameters():
            if n.endswith(".bias"):
                p.normal_(0.0, 0.5)
            elif "norm" in n:
                p.uniform_(0.5, 1.5)
            elif p.dim() == 2 and "embed" not in n and "lm_head" not in n:
                p.normal_(0.0, 0.06)
    model = model.to(torch.bfloat16).eval()
    sd = {k: v.contiguous() for k, v in model.state_dict().items()}
    tied = kind == "qwen2"
    if tied:
        sd.pop("lm_head.weight", None)
    shard = os.path.join(out_dir, "model.safetensors")
    st.save_file(sd, shard)
    meta, wts = os.path.join(out_dir, "meta_data.txt"), os.path.join(out_dir, "weights.bin")
    q.convert_safetensors([shard], meta, wts, tie_lm_head=tied)
    return model.float(), meta, wts


def hf_logits(model, ids):
    """fp32 logits of every position for one sequence"""
    import torch
    with torch.no_grad():
        return model(torch.tensor([list(int(i) for i in ids)])).logits[0].numpy()


def hf_bf16_error(model, ids):
    """how far transformers' OWN bf16 execution of the model is from its fp32 execution (last position, relative L2):
    the yardstick for a bf16 pipeline.  A bf16 engine cannot be asked to be closer to fp32 than this."""
    import copy

    import torch
    w32 = hf_logits(model, ids)[-1]
    mb = copy.deepcopy(model).to(torch.bfloat16)
    with torch.no_grad():
        w16 = mb(torch.tensor([list(int(i) for i in ids)])).logits[0][-1].float().numpy()
    return rel_l2_f32(w16, w32)


def hf_tolerance(model, ids):
    """1e-2 (north star) where transformers' own bf16 run meets it, else 1.25 x its error + 2e-3"""
    return max(1e-2, 1.25 * hf_bf16_error(model, ids) + 2e-3)


def rel_l2_f32(a_f32, b_f32):
    a, b = np.asarray(a_f32, np.float64), np.asarray(b_f32, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))
