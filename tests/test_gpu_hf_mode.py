"""-m gpu: "HF-correct" model semantics through the engine: real-checkpoint-shaped tiny Qwen2
(Qwen2.5 family: q/k/v bias, no q/k-norm, tied lm_head) and Qwen3 (q/k-norm, the reference's family) models from
Hugging Face transformers, converted by qie_convert_safetensors, run with semantics="hf" (half-rotation RoPE, eps 1e-6)
in both numerics modes -- against transformers in fp32 and against the C oracle's restatement (which
tests/test_cpu_hf_semantics.py pins to transformers on the CPU)."""
import numpy as np
import pytest

from hf_util import hf_logits, hf_tolerance, make_hf_checkpoint, rel_l2_f32
from util import bf16_to_f32, rel_err

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
pytest.importorskip("transformers")


@pytest.mark.parametrize("kind", ["qwen2", "qwen3"])
@pytest.mark.parametrize("numerics", ["reference_order", "fast"])
def test_engine_hf_semantics(oracle, tmp_path, kind, numerics):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qwen_inference_engine_b200 as q
    from oracle.oracle import OracleModel
    model, meta, wts = make_hf_checkpoint(kind, str(tmp_path))
    ids = np.array([3, 77, 512, 9, 1000, 41, 5, 640, 222, 18, 7, 300], np.int32)
    want = hf_logits(model, ids)
    tol = hf_tolerance(model, ids)
    eng = q.Engine(meta, wts, head_dim_hint=64, context=512, max_batch_tokens=64, kv_bytes=64 << 20, use_graph=False,
                   numerics=numerics, semantics="hf")
    assert not eng.uses_mega(1, 64)  # HF semantics run on the per-operator path
    eng.capture(True)
    s = eng.new_sequence()
    tok = eng.prefill(s, ids)
    lg = eng.read_capture("logits", -1)
    err = rel_l2_f32(bf16_to_f32(lg), want[-1])
    assert err < tol, (err, tol)
    # the oracle's restatement of the same semantics (GEMM accumulation order differs: bf16 tolerance)
    om = OracleModel(oracle, meta, wts, head_dim_hint=64, context=512).set_semantics(rope_half=True, eps=1e-6)
    so = om.new_seq()
    _, lo = om.prefill(so, ids, want_logits=True)
    assert rel_err(lg, lo) < 1e-2
    # greedy decode, teacher-forced on the engine's tokens: logits track transformers step by step
    seq, agree = list(int(i) for i in ids), 0
    for step in range(12):
        w = hf_logits(model, seq)[-1]
        agree += int(tok == int(np.argmax(w)))
        seq.append(int(tok))
        tok = int(eng.decode_step([s], [tok])[0])
        lg = eng.read_capture("logits", -1)
        assert rel_l2_f32(bf16_to_f32(lg), hf_logits(model, seq)[-1]) < tol
    assert agree >= 10  # near-ties of a random-init model may flip a choice
    om.close()
    eng.close()
    # reference semantics on the same checkpoint are a different function (and a checkpoint with an unknown bias is refused)
    eng2 = q.Engine(meta, wts, head_dim_hint=64, context=512, max_batch_tokens=64, kv_bytes=64 << 20, use_graph=False)
    eng2.capture(True)
    eng2.prefill(eng2.new_sequence(), ids)
    assert rel_l2_f32(bf16_to_f32(eng2.read_capture("logits", -1)), want[-1]) > 5e-2
    eng2.close()


def test_checkpoint_with_an_unapplied_bias_is_refused(tmp_path):
    """a bias tensor the forward would silently ignore must fail the load, not produce wrong logits"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qwen_inference_engine_b200 as q
    cfg = q.make_config("tiny", context=128)
    meta, wts = str(tmp_path / "meta_data.txt"), str(tmp_path / "weights.bin")
    q.write_synthetic_checkpoint(cfg, 1, meta, wts)
    txt = open(meta).read()
    import re
    end = max(int(m.group(1)) for m in re.finditer(r"offsets: \[ \d+, (\d+) \]", txt))
    with open(meta, "a") as f:
        f.write(f"Tensor: model.layers.0.self_attn.o_proj.bias\n  layer: 0\n  short_name: self_attn.o_proj.bias\n  shape: [ {cfg.hidden} ]\n"
                f"  offsets: [ {end}, {end + 2 * cfg.hidden} ]\n\n")
    with open(wts, "ab") as f:
        f.write(b"\0" * (2 * cfg.hidden))
    with pytest.raises(q.QieError):
        q.Engine(meta, wts, context=128)
