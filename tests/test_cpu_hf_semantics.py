"""not-gpu: the C oracle's "HF-correct" semantics (half-rotation RoPE, eps 1e-6, q/k/v bias, optional q/k-norm, tied
lm_head) pinned against Hugging Face transformers itself: tiny random-init Qwen2 and Qwen3 models, converted with the
product's safetensors converter, forward of the oracle vs the HF model in fp32.  The reference's own
semantics on the same checkpoints must NOT match (the switch does something)."""
import numpy as np
import pytest

from hf_util import hf_logits, hf_tolerance, make_hf_checkpoint, rel_l2_f32
from util import bf16_to_f32

pytest.importorskip("transformers")
pytest.importorskip("safetensors")


@pytest.mark.parametrize("kind", ["qwen2", "qwen3"])
def test_oracle_hf_semantics_match_transformers(oracle, tmp_path, kind):
    from oracle.oracle import OracleModel
    model, meta, wts = make_hf_checkpoint(kind, str(tmp_path))
    ids = np.array([3, 77, 512, 9, 1000, 41, 5, 640, 222, 18], np.int32)
    want = hf_logits(model, ids)
    om = OracleModel(oracle, meta, wts, head_dim_hint=64, context=512).set_semantics(rope_half=True, eps=1e-6)
    s = om.new_seq()
    tok, lg = om.prefill(s, ids, want_logits=True)
    tol = hf_tolerance(model, ids)  # 1e-2, or transformers' own bf16-vs-fp32 error where that is larger
    err = rel_l2_f32(bf16_to_f32(lg), want[-1])
    assert err < tol, (err, tol)
    assert tok == int(np.argmax(want[-1]))
    # decode: append tokens one at a time and compare with the HF model on the grown sequence
    seq = list(ids)
    for step in range(4):
        seq.append(int(tok))
        tok, lg = om.decode(s, tok, want_logits=True)
        w = hf_logits(model, seq)[-1]
        assert rel_l2_f32(bf16_to_f32(lg), w) < tol
    om.close()
    # the reference's semantics (interleaved RoPE, eps 1e-4) on the same weights give something else
    om2 = OracleModel(oracle, meta, wts, head_dim_hint=64, context=512)
    s2 = om2.new_seq()
    _, lg2 = om2.prefill(s2, ids, want_logits=True)
    assert rel_l2_f32(bf16_to_f32(lg2), want[-1]) > 5e-2
    om2.close()
