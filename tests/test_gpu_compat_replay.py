"""-m gpu: the drop-in boundary, executed.

(1) oracle/ref_driver.cu -- llm()'s call sequence written against the REFERENCE's names (launch_rms, launch_matmul,
    launch_qknorm, launch_rope[_single], kv_copy_layer_to_cache_{prefill,decode}, launch_attn over a page_table list,
    launch_act / launch_elem / launch_resadd, copy_*_vec, sample_topk_bf16, create_page_list / allocate_page_buffers,
    ModelBuffers; qwen_main.cu:74-247,250-404) -- is compiled TWICE: against the reference's helpers.cuh + kernels
    (oracle/_ref) and against include/layers/iengine_compat.hh + libqie_b200.so (oracle/_compat).  Same source, same
    exported ref_* functions; this file runs BASELINE configs[0] through both and requires identical tokens, logits and
    cache pages.
(2) tests/compat/main_callseq.cpp -- main()'s call sequence (iengine.cu:226-482: build_indexed_tensors, the weight
    blob, create_new_sequence / initialize_model_buffers, create_page_list / allocate_page_buffers, the llm() loop)
    against the same header, as a program: its tokens must equal the reference kernels' tokens."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from util import prompt_ids, rand_bf16, to_dev, to_host

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT_SO = os.path.join(ROOT, "oracle", "_compat", "libqie_compat_replay.so")
CALLSEQ = os.path.join(ROOT, "oracle", "_compat", "main_callseq")


@pytest.fixture(scope="module")
def qie():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qwen_inference_engine_b200 as q
    return q


@pytest.fixture(scope="module")
def compat():
    from oracle.oracle import Ref
    if not os.path.exists(COMPAT_SO):
        pytest.fail("oracle/_compat/libqie_compat_replay.so missing: __graft_entry__.build() makes it")
    return Ref(COMPAT_SO)


def _generate(lib, desc, ids, n_new, topk, vocab, page_size=4):
    from oracle.oracle import RefSeq
    rs = RefSeq(lib, desc, page_size=page_size)
    toks = [rs.prefill(ids, topk=topk)]
    logits = [rs.read("logits", vocab)]
    for _ in range(n_new - 1):
        toks.append(rs.decode(toks[-1], topk=topk))
        logits.append(rs.read("logits", vocab))
    # the cache the call sequence built: every page, K and V
    pages = lib.L.ref_seq_pages(rs.h)
    n_pages = lib.L.ref_pages_count(pages)
    return toks, logits, rs, pages, n_pages


@pytest.mark.parametrize("arch,n_prompt,n_new,topk", [("qwen2.5-0.5b", 32, 16, 1), ("small", 9, 24, 50), ("small128", 5, 12, 1)])
def test_llm_call_sequence_same_source_two_libraries(qie, ref, compat, arch, n_prompt, n_new, topk):
    """config 1 (0.5B-arch, 32-token prompt, greedy) for 16 tokens, plus the reference's real sampling mode (top-k 50)
    and a head_dim-128 model: the replay linked with libqie_b200 == the replay linked with the reference's kernels."""
    eng = qie.Engine(synthetic=arch, seed=1234, max_batch_tokens=64, kv_bytes=128 << 20, context=512 if arch != "qwen2.5-0.5b" else 0)
    cfg = eng.config
    ids = prompt_ids(n_prompt, cfg.vocab)
    desc = ref.model_desc(eng)
    want_t, want_l, rs_r, pg_r, np_r = _generate(ref, desc, ids, n_new, topk, cfg.vocab)
    got_t, got_l, rs_c, pg_c, np_c = _generate(compat, desc, ids, n_new, topk, cfg.vocab)
    assert got_t == want_t
    for i in range(n_new):
        assert np.array_equal(got_l[i], want_l[i]), f"logits differ at step {i}"
    # cache pages written by kv_copy_layer_to_cache_{prefill,decode}: same layout, same bits
    elems = 4 * cfg.layers * cfg.n_kv * cfg.head_dim
    used = (n_prompt + n_new - 1 + 3) // 4
    assert np_c >= used and np_r >= used
    for pgi in range(used):
        for which in (0, 1):
            a, b = np.zeros(elems, np.uint16), np.zeros(elems, np.uint16)
            assert ref.L.ref_pages_read(pg_r, pgi, which, a.ctypes.data, elems) == 0
            assert compat.L.ref_pages_read(pg_c, pgi, which, b.ctypes.data, elems) == 0
            assert np.array_equal(a, b), f"page {pgi} {'V' if which else 'K'} differs"
    rs_r.close()
    rs_c.close()
    # and the engine's own forward gives the same tokens
    eng.set_sampling(topk=topk, temperature_prefill=1.0, temperature_decode=0.7, seed=1234, add_step=True)
    assert eng.generate(ids, n_new) == want_t
    eng.close()


@pytest.mark.parametrize("hd,n_q,n_kv,layers_n,t,mq,causal,ps", [(64, 14, 2, 3, 2048, 1, 0, 4), (128, 12, 2, 2, 77, 77, 1, 4),
                                                                 (64, 4, 4, 1, 33, 1, 0, 16), (128, 8, 2, 2, 130, 3, 1, 7)])
def test_launch_attn_over_reference_page_list(ref, compat, hd, n_q, n_kv, layers_n, t, mq, causal, ps):
    """helpers.cuh:121-129 launch_attn(Q, out, mq, mkv, hd, hidden, hidden_kv, causal, q_abs_base, layer, page_table*,
    page_size) on a page list in the reference's [slot][layer][kv_dim] layout, incl. ctx 2048 and a page size that is
    not a power of two."""
    rng = np.random.default_rng(hd + t + ps)
    Dq, Dkv, layer = n_q * hd, n_kv * hd, layers_n - 1
    K, V, Q = rand_bf16(rng, (t, Dkv), 1.0), rand_bf16(rng, (t, Dkv), 1.0), rand_bf16(rng, (mq, Dq), 1.0)
    n_pages, elems = (t + ps - 1) // ps, ps * layers_n * Dkv
    outs = []
    q_abs_base = t - mq  # the last mq positions query the cache
    for lib in (ref, compat):
        pages = lib.L.ref_pages_create(n_pages, elems)
        for pg in range(n_pages):
            kb, vb = np.zeros((ps, layers_n, Dkv), np.uint16), np.zeros((ps, layers_n, Dkv), np.uint16)
            rows = K[pg * ps:(pg + 1) * ps]
            kb[:len(rows), layer] = rows
            vb[:len(rows), layer] = V[pg * ps:(pg + 1) * ps]
            assert lib.L.ref_pages_write(pages, pg, 0, kb.ctypes.data, elems) == 0
            assert lib.L.ref_pages_write(pages, pg, 1, vb.ctypes.data, elems) == 0
        Qd = to_dev(Q)
        o = torch.zeros_like(Qd)
        assert lib.L.ref_attn(Qd.data_ptr(), o.data_ptr(), mq, t, hd, Dq, Dkv, causal, q_abs_base, layer, pages, ps, layers_n) == 0
        outs.append(to_host(o))
        lib.L.ref_pages_free(pages)
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("topk", [1, 50])
def test_main_call_sequence_program(qie, ref, topk):
    """main()'s call sequence as a program linked with libqie_b200 (tests/compat/main_callseq.cpp): config 1's first
    16 tokens (greedy) and the reference's hard-coded sampling (top-k 50, T 1.0 / 0.7, seed 1234 + step)."""
    assert os.path.exists(CALLSEQ), "oracle/_compat/main_callseq missing: __graft_entry__.build() makes it"
    cfg = qie.make_config("qwen2.5-0.5b")
    d = tempfile.mkdtemp()
    meta, wts = os.path.join(d, "meta_data.txt"), os.path.join(d, "weights.bin")
    qie.write_synthetic_checkpoint(cfg, 1234, meta, wts)
    ids = prompt_ids(32, cfg.vocab)
    env = dict(os.environ, QIE_META=meta, QIE_COMPAT_TOPK=str(topk))
    out = subprocess.run([CALLSEQ, wts, "16"] + [str(int(i)) for i in ids], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    got = [int(x) for x in out.stdout.split()]
    eng = qie.Engine(meta, wts, max_batch_tokens=64, kv_bytes=128 << 20)
    from test_gpu_e2e_vs_reference import _ref_generate
    want, _, _ = _ref_generate(ref, eng, ids, 16, topk=topk)
    eng.close()
    os.remove(wts)
    assert got == want
