"""gpu: continuous batching (qie_sched_*).  Requests of different lengths share decode steps, join and leave the
batch at different times, wait for KV pages; in reference-order numerics every request must get exactly the tokens
it gets when it runs alone (llm() of the reference is per sequence: iengine.cu:327-421), EOS ends a request at once
and all pages return to the pool."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pytestmark = pytest.mark.gpu


def _requests(vocab):
    from util import prompt_ids
    lens = [3, 40, 7, 19, 33, 5, 12, 27, 9, 16, 4]
    news = [24, 5, 17, 9, 12, 30, 6, 14, 21, 8, 11]
    return [(prompt_ids(n, vocab, seed=100 + i), m) for i, (n, m) in enumerate(zip(lens, news))]


def test_continuous_batching_matches_solo_runs():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qwen_inference_engine_b200 as q
    # 16-position pages; 14 pages: at most ~3-4 requests (prompt + budget reserved) fit at once -> queueing
    eng = q.Engine(synthetic="small", seed=5, context=256, max_seqs=5, max_pages=14, max_batch_tokens=64)
    reqs = _requests(eng.config.vocab)
    solo = [eng.generate(ids, m) for ids, m in reqs]
    total_pages = eng.pages_free()
    sch = q.Scheduler(eng, max_running=4, eos_token=-1)
    rids = [sch.submit(ids, m) for ids, m in reqs]
    seen_running = set()
    for _ in range(10000):
        left = sch.step()
        st = sch.stats()
        seen_running.add(st["running"])
        assert st["running"] <= 4
        if left == 0:
            break
    for rid, want in zip(rids, solo):
        got, fin = sch.result(rid)
        assert fin and got == want
    st = sch.stats()
    assert st["prefills"] == len(reqs) and st["decode_rows"] == sum(len(s) - 1 for s in solo)
    assert max(seen_running) >= 3 and st["steps"] < sum(len(s) for s in solo)  # requests really shared steps
    assert eng.pages_free() == total_pages  # every page came back
    sch.close()

    # EOS: a token that some request produces in mid-stream ends that request there (qwen_main.cu:257)
    eos = solo[5][9]
    sch = q.Scheduler(eng, max_running=3, eos_token=eos)
    rids = [sch.submit(ids, m) for ids, m in reqs]
    sch.run()
    cut = 0
    for rid, want in zip(rids, solo):
        got, fin = sch.result(rid)
        exp = want[:want.index(eos) + 1] if eos in want else want
        cut += len(exp) < len(want)
        assert fin and got == exp
    assert cut >= 1 and eng.pages_free() == total_pages
    # a request that can never fit is refused at submit
    with pytest.raises(q.QieError):
        sch.submit(np.arange(200, dtype=np.int32) % 100, 100)
    sch.close()
    eng.close()


def test_kv_swap_out_and_in_keeps_generation_bit_exact():
    """KV offload (iengine.cu:376-429, commented out in the reference): a sequence's pages go to pinned host memory,
    other sequences overwrite them, the sequence comes back on different pages and continues with the same tokens."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qwen_inference_engine_b200 as q
    from util import prompt_ids
    eng = q.Engine(synthetic="small", seed=8, context=256, max_seqs=4, max_pages=8, max_batch_tokens=64)
    ids = prompt_ids(37, eng.config.vocab, seed=3)
    want = eng.generate(ids, 24)
    total = eng.pages_free()
    a = eng.new_sequence()
    toks = [eng.prefill(a, ids)]
    toks += [int(t) for t in eng.decode_run([a], [toks[0]], 7)[:, 0]]
    used = total - eng.pages_free()
    eng.swap_out(a)
    assert eng.pages_free() == total and eng.seq_len(a) == 37 + 7
    with pytest.raises(q.QieError):
        eng.decode_step([a], [toks[-1]])  # swapped out
    with pytest.raises(q.QieError):
        eng.swap_out(a)
    # other work takes (and dirties) the whole pool
    others = [eng.generate(prompt_ids(50 + 9 * i, eng.config.vocab, seed=20 + i), 12) for i in range(2)]
    b = eng.new_sequence()
    eng.prefill(b, prompt_ids(100, eng.config.vocab, seed=40))  # 7 pages: A's old pages are in use again
    with pytest.raises(q.QieError):
        eng.swap_in(a)  # not enough free pages
    eng.free_sequence(b)
    eng.swap_in(a)
    assert total - eng.pages_free() == used
    toks += [int(t) for t in eng.decode_run([a], [toks[-1]], 16)[:, 0]]
    assert toks == want and len(others) == 2
    eng.free_sequence(a)
    assert eng.pages_free() == total
    eng.close()


def test_bad_requests_are_refused_and_never_wedge_the_queue():
    """a request that can never run (prompt + budget beyond the context, token id outside the vocabulary)
    is refused at submit; max_running is clamped to what a decode step takes; the good requests around it finish."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qwen_inference_engine_b200 as q
    from util import prompt_ids
    eng = q.Engine(synthetic="small", seed=5, context=64, max_seqs=4, max_pages=32, max_batch_tokens=8)
    V = eng.config.vocab
    sch = q.Scheduler(eng, max_running=1000, eos_token=-1)  # clamped to min(8 rows, 4 sequence slots)
    good1 = sch.submit(prompt_ids(5, V, seed=1), 6)
    for bad in (lambda: sch.submit(prompt_ids(40, V, seed=2), 30),            # 70 positions > context 64
                lambda: sch.submit(np.array([1, 2, V + 3], np.int32), 4),     # id outside the vocabulary
                lambda: sch.submit(np.array([1, -2, 3], np.int32), 4)):
        with pytest.raises(q.QieError) as ei:
            bad()
        assert ei.value.code == -1  # QIE_EINVAL
    more = [sch.submit(prompt_ids(4 + i, V, seed=10 + i), 5) for i in range(6)]  # more requests than sequence slots
    left = sch.run()
    assert left == 0
    for rid in [good1] + more:
        toks, fin = sch.result(rid)
        assert fin and sch.last_status == 1 and len(toks) in (5, 6)
    st = sch.stats()
    assert st["running"] == 0 and st["waiting"] == 0
    # solo run of one of them gives the same tokens (nothing was disturbed by the refused submissions)
    assert eng.generate(prompt_ids(5, V, seed=1), 6) == sch.result(good1)[0]
    # duplicate sequence ids in one decode batch are refused
    s0 = eng.new_sequence()
    t0 = eng.prefill(s0, prompt_ids(3, V))
    with pytest.raises(q.QieError):
        eng.decode_step([s0, s0], [t0, t0])
    sch.close()
    eng.close()
