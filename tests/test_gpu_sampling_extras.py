"""-m gpu: the sampling surface beyond greedy.
* radix-select top-k kernel vs the reference's k-round scan (oracle/_ref) on adversarial rows: long runs of ties across
  the k-th boundary, +-0, +inf, -inf / NaN entries, fewer finite logits than k, every k from 2 to 256;
* repetition penalty (declared, never defined in the reference, layers_include.cuh:33): the CUDA operator vs the C oracle
  bit for bit, and the driver-level option (token history on the device, penalty before greedy / top-k sampling in the
  per-operator path and behind the persistent kernel) against a host-side replay of the same rule."""
import numpy as np
import pytest

from util import bf16_to_f32, f32_to_bf16, prompt_ids, to_dev, to_host

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def layers():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from qwen_inference_engine_b200 import layers as L
    return L


@pytest.mark.parametrize("vocab,k,kind", [(151936, 50, "ties"), (151936, 2, "ties"), (151936, 256, "few_levels"), (4096, 37, "special"),
                                          (70000, 50, "sparse"), (1000, 50, "special"), (151936, 50, "normal"), (513, 200, "ties")])
def test_topk_select_equals_reference_scan(layers, ref, oracle, vocab, k, kind):
    rng = np.random.default_rng(vocab * 7 + k)
    for trial in range(3):
        if kind == "ties":
            vals = rng.choice(np.linspace(-2, 2, 23).astype(np.float32), size=vocab)
        elif kind == "few_levels":
            vals = rng.choice(np.array([0.5, 0.25], np.float32), size=vocab)
        elif kind == "normal":
            vals = rng.standard_normal(vocab).astype(np.float32) * 3
        elif kind == "sparse":  # fewer finite logits than k
            vals = np.full(vocab, -np.inf, np.float32)
            vals[rng.choice(vocab, size=k // 2, replace=False)] = rng.standard_normal(k // 2).astype(np.float32)
        else:  # +-0, +inf, -inf, NaN mixed in
            vals = rng.choice(np.array([0.0, -0.0, 1.5, -1.5, np.inf, -np.inf, np.nan, 3.0], np.float32), size=vocab)
        lg = f32_to_bf16(vals)
        if kind == "special":
            lg[vals != vals] = 0x7FC0  # canonical bf16 NaN
        ld = to_dev(lg)
        for temp, seed, step in ((0.7, 1234 + trial, 0), (1.0, 99, 3)):
            want = ref.L.ref_sample(ld.data_ptr(), vocab, temp, k, seed, step)
            got = int(layers.sample_topk_bf16(ld, vocab, temp, k, seed, step)[0])
            assert got == want, (kind, trial, temp, seed, step)
            if kind != "special" and step == 0:  # the C oracle restates XORWOW subsequence 0 only (the one llm() uses)
                assert oracle.sample_topk(lg, temp, k, seed) == want


@pytest.mark.parametrize("vocab,n_ctx,penalty", [(151936, 700, 1.3), (4096, 50, 2.0), (1000, 1, 1.1), (512, 300, 0.8)])
def test_repetition_penalty_operator_vs_oracle(layers, oracle, vocab, n_ctx, penalty):
    rng = np.random.default_rng(vocab + n_ctx)
    lg = f32_to_bf16((rng.standard_normal(vocab) * 2).astype(np.float32))
    ctx = rng.integers(-3, vocab + 3, size=n_ctx).astype(np.int32)  # duplicates and out-of-range ids included
    ctx[n_ctx // 2:] = ctx[:n_ctx - n_ctx // 2]
    ld, cd = to_dev(lg), torch.from_numpy(ctx).cuda()
    layers.apply_repetition_penalty(ld, cd, vocab, penalty)
    torch.cuda.synchronize()
    got, want = to_host(ld), oracle.repetition_penalty(lg, ctx, penalty)
    assert np.array_equal(got, want)
    seen = np.unique(ctx[(ctx >= 0) & (ctx < vocab)])
    untouched = np.setdiff1d(np.arange(vocab), seen)
    assert np.array_equal(got[untouched], lg[untouched])
    a, b = bf16_to_f32(got[seen]), bf16_to_f32(lg[seen])
    assert np.all(a <= b) if penalty > 1 else np.all(a >= b)  # a penalty > 1 never raises a logit (positive / p, negative * p)


@pytest.mark.parametrize("arch,topk,mega", [("small", 1, 0), ("small", 1, 1), ("small", 50, 1), ("tiny", 7, 0)])
def test_engine_repetition_penalty(oracle, arch, topk, mega):
    """driver level: tokens with the penalty == a host replay (logits of the penalty-free forward + the oracle's penalty
    over the true history + the oracle's sampler), teacher-forced on the engine's own tokens"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qwen_inference_engine_b200 as q
    eng = q.Engine(synthetic=arch, seed=1234, context=256, max_batch_tokens=64, use_graph=False)
    V = eng.config.vocab
    ids = prompt_ids(9, V)
    pen, n_new = 1.7, 20
    eng.set_int("mega", mega)
    eng.set_sampling(topk=topk, temperature_prefill=1.0, temperature_decode=0.7, seed=1234, add_step=True)
    eng.set_repetition_penalty(pen)
    got = eng.generate(ids, n_new)
    # replay: same tokens fed to a penalty-free engine with capture (per-operator path keeps the raw logits)
    eng.set_repetition_penalty(1.0)
    eng.set_int("mega", 0)
    eng.capture(True)
    s = eng.new_sequence()
    eng.prefill(s, ids)
    hist = list(int(t) for t in ids)
    want = []
    for i in range(n_new):
        raw = eng.read_capture("logits", -1)
        lg = oracle.repetition_penalty(raw, np.asarray(hist, np.int32), pen)
        temp, seed = (1.0, 1234) if i == 0 else (0.7, 1234 + i)
        tok = oracle.argmax_tiebreak(lg) if topk == 1 else oracle.sample_topk(lg, temp, topk, seed)
        want.append(int(tok))
        assert got[i] == tok, f"token {i}"
        if i + 1 < n_new:
            eng.decode_step([s], [got[i]])
            hist.append(got[i])
    eng.close()
    # the penalty changes the output (otherwise the test proves nothing)
    eng2 = q.Engine(synthetic=arch, seed=1234, context=256, max_batch_tokens=64, use_graph=False)
    eng2.set_sampling(topk=topk, temperature_prefill=1.0, temperature_decode=0.7, seed=1234, add_step=True)
    assert eng2.generate(ids, n_new) != got
    eng2.close()
