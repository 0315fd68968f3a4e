"""-m gpu: every operator of libqie_b200 (through the C ABI) against the reference's OWN
kernel (oracle/_ref, compiled from /root/reference/layers/src) on the same seeded inputs.
The bar is BIT-EXACT: the reference-order kernels reproduce the reference's arithmetic
order, so there is no tolerance in this file.  The same calls are also checked against
the plain-C oracle (bit-exact where the CPU can restate the arithmetic; tolerance where it
cannot: GEMM accumulation through HMMA, device expf)."""
import ctypes as C

import numpy as np
import pytest

from util import bf16_to_f32, rand_bf16, rel_err, to_dev, to_host, ulp_diff

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def layers():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from qwen_inference_engine_b200 import layers as L
    return L


def p(t):
    return t.data_ptr()


@pytest.mark.parametrize("hidden,n_tok", [(896, 1), (896, 5), (1536, 3), (3584, 2), (128, 33), (5120, 1)])
def test_rmsnorm_bit_exact(layers, ref, oracle, hidden, n_tok):
    rng = np.random.default_rng(hidden + n_tok)
    x = rand_bf16(rng, (n_tok, hidden), 0.7)
    w = rand_bf16(rng, (hidden,), 1.0)
    xd, wd = to_dev(x), to_dev(w)
    y_ref = torch.zeros_like(xd)
    y_new = torch.zeros_like(xd)
    assert ref.L.ref_rmsnorm(p(xd), p(wd), p(y_ref), hidden, n_tok) == 0
    layers.launch_rms(xd, wd, y_new, hidden, n_tok)
    torch.cuda.synchronize()
    assert np.array_equal(to_host(y_new), to_host(y_ref))
    assert np.array_equal(to_host(y_new), oracle.rmsnorm(x, w))  # CPU restatement is exact here too


@pytest.mark.parametrize("M,N,K", [(1, 896, 896), (1, 896, 128), (1, 4864, 896), (1, 896, 4864), (1, 896, 15200),
                                   (2, 896, 1152), (7, 1536, 256), (8, 128, 264), (9, 256, 64), (16, 896, 136),
                                   (17, 512, 72), (32, 896, 896), (33, 896, 128), (64, 896, 4864), (64, 4864, 896),
                                   (65, 256, 40), (130, 128, 24), (1, 72, 16), (3, 200, 8)])
def test_matmul_bit_exact(layers, ref, oracle, M, N, K):
    """C[M,K] = A[M,N] @ B[K,N]^T vs the reference's wmma kernel (matrix_mul.cu:165)."""
    rng = np.random.default_rng(M * 1000003 + N * 101 + K)
    A = rand_bf16(rng, (M, N), 1.0)
    B = rand_bf16(rng, (K, N), 0.05)
    Ad, Bd = to_dev(A), to_dev(B)
    c_ref = torch.zeros(M, K, dtype=torch.bfloat16, device="cuda")
    c_new = torch.zeros(M, K, dtype=torch.bfloat16, device="cuda")
    assert ref.L.ref_matmul(p(Ad), p(Bd), p(c_ref), M, N, K) == 0
    layers.launch_matmul(Ad, Bd, c_new, M, N, K)
    torch.cuda.synchronize()
    got, want = to_host(c_new), to_host(c_ref)
    assert np.array_equal(got, want), f"max ulp diff {ulp_diff(got, want)}, mismatches {(got != want).sum()}/{got.size}"
    # CPU oracle: sequential fp32 accumulation; HMMA's internal order is not restatable -> tolerance
    assert rel_err(got, oracle.matmul(A, B)) < 1e-2


@pytest.mark.parametrize("hd,n_heads,n_tok", [(64, 14, 1), (64, 2, 3), (128, 12, 2), (128, 4, 1), (64, 16, 9)])
def test_qknorm_and_rope_bit_exact(layers, ref, oracle, hd, n_heads, n_tok):
    rng = np.random.default_rng(hd + n_heads * 7 + n_tok)
    row = hd * n_heads
    x = rand_bf16(rng, (n_tok, row), 1.3)
    w = rand_bf16(rng, (hd,), 1.0)
    xd_ref, xd_new, wd = to_dev(x), to_dev(x), to_dev(w)
    assert ref.L.ref_qknorm(p(xd_ref), p(wd), hd, n_tok, row, n_heads) == 0
    layers.launch_qknorm(xd_new, wd, hd, n_tok, row, n_heads)
    torch.cuda.synchronize()
    assert np.array_equal(to_host(xd_new), to_host(xd_ref))
    assert np.array_equal(to_host(xd_new), oracle.qknorm(x, w, hd, n_heads))
    # RoPE on top, prefill form (rows 0..n_tok-1) and decode form (single position)
    ctx = 64
    cos, sin = layers.precompute_cos_sin(ctx, hd)
    oc, osn = oracle.cos_sin(ctx, hd)
    assert np.array_equal(cos, oc) and np.array_equal(sin, osn)
    cd, sd = torch.from_numpy(cos).cuda(), torch.from_numpy(sin).cuda()
    a, b = xd_ref.clone(), xd_new.clone()
    assert ref.L.ref_rope(p(cd), p(sd), p(a), n_tok, hd, row, n_heads) == 0
    layers.launch_rope(cd, sd, b, n_tok, hd, row, n_heads)
    torch.cuda.synchronize()
    assert np.array_equal(to_host(a), to_host(b))
    assert np.array_equal(to_host(b), oracle.rope(cos, sin, to_host(xd_new), 0, hd, n_heads))
    a1, b1 = xd_ref[:1].clone(), xd_new[:1].clone()
    assert ref.L.ref_rope_single(p(cd), p(sd), p(a1), 37, hd, row, n_heads) == 0
    layers.launch_rope_single(cd, sd, b1, 37, hd, row, n_heads)
    torch.cuda.synchronize()
    assert np.array_equal(to_host(a1), to_host(b1))


def test_elementwise_bit_exact(layers, ref, oracle):
    rng = np.random.default_rng(5)
    n = 4864 * 3 + 5
    a, b = rand_bf16(rng, (n,), 2.0), rand_bf16(rng, (n,), 1.0)
    # SiLU
    x_ref, x_new = to_dev(a), to_dev(a)
    assert ref.L.ref_act(p(x_ref), n) == 0
    layers.launch_act(x_new, n)
    torch.cuda.synchronize()
    assert np.array_equal(to_host(x_ref), to_host(x_new))
    assert ulp_diff(to_host(x_new), oracle.silu(a)) <= 1  # glibc expf vs device expf
    # element_mul
    ad, bd = to_dev(a), to_dev(b)
    o_ref, o_new = torch.zeros_like(ad), torch.zeros_like(ad)
    assert ref.L.ref_elem(p(ad), p(bd), p(o_ref), n) == 0
    layers.launch_elem(ad, bd, o_new, n)
    torch.cuda.synchronize()
    assert np.array_equal(to_host(o_ref), to_host(o_new))
    assert np.array_equal(to_host(o_new), oracle.elem_mul(a, b))
    # residual_add
    r_ref, r_new = to_dev(a), to_dev(a)
    assert ref.L.ref_resadd(p(r_ref), p(bd), n) == 0
    layers.launch_resadd(r_new, bd, n)
    torch.cuda.synchronize()
    assert np.array_equal(to_host(r_ref), to_host(r_new))
    assert np.array_equal(to_host(r_new), oracle.residual_add(a, b))


def test_embedding_bit_exact(layers, ref, oracle):
    rng = np.random.default_rng(11)
    table = rand_bf16(rng, (1000, 896), 0.02)
    ids = rng.integers(0, 1000, size=37).astype(np.int32)
    td, idd = to_dev(table), torch.from_numpy(ids).cuda()
    o_ref = torch.zeros(37, 896, dtype=torch.bfloat16, device="cuda")
    o_new = torch.zeros_like(o_ref)
    assert ref.L.ref_embedding(p(o_ref), p(td), p(idd), 896, 37) == 0
    layers.embedding(o_new, td, idd)
    torch.cuda.synchronize()
    assert np.array_equal(to_host(o_ref), to_host(o_new))
    assert np.array_equal(to_host(o_new), oracle.embedding(table, ids))


@pytest.mark.parametrize("hd,n_q,n_kv,layers_n,t,mq,causal", [
    (64, 14, 2, 3, 33, 1, 0),      # decode, 0.5B head geometry
    (64, 14, 2, 2, 40, 40, 1),     # prefill, causal
    (128, 12, 2, 2, 19, 1, 0),     # 1.5B head geometry
    (128, 4, 4, 1, 9, 9, 1),       # MHA
    (64, 4, 1, 2, 160, 1, 0),      # config-1 final context length
    (64, 14, 2, 1, 2048, 1, 0),    # BASELINE configs[1] context: 0.5B head geometry at t = 2048
    (128, 12, 2, 1, 2049, 1, 0),   # 1.5B head geometry, odd length across the 2048 boundary
    (64, 2, 1, 1, 700, 700, 1),    # causal prefill rows over many pages
])
def test_attention_bit_exact(layers, ref, oracle, hd, n_q, n_kv, layers_n, t, mq, causal):
    """selfattention over a paged cache: reference page list (page_size 4, layout
    [slot][layer][kv_dim]) vs the B200 pool (page_size 16, [layer][k|v][head][slot][hd])."""
    rng = np.random.default_rng(hd * 31 + t)
    Dq, Dkv = n_q * hd, n_kv * hd
    layer = layers_n - 1
    K = rand_bf16(rng, (t, Dkv), 1.0)
    V = rand_bf16(rng, (t, Dkv), 1.0)
    Q = rand_bf16(rng, (mq, Dq), 1.0)
    # --- reference cache
    ps_ref = 4
    n_pages = (t + ps_ref - 1) // ps_ref
    elems = ps_ref * layers_n * Dkv
    pages = ref.L.ref_pages_create(n_pages, elems)
    for pg in range(n_pages):
        kb = np.zeros((ps_ref, layers_n, Dkv), np.uint16)
        vb = np.zeros((ps_ref, layers_n, Dkv), np.uint16)
        rows = K[pg * ps_ref:(pg + 1) * ps_ref]
        kb[:len(rows), layer] = rows
        vb[:len(rows), layer] = V[pg * ps_ref:(pg + 1) * ps_ref]
        assert ref.L.ref_pages_write(pages, pg, 0, kb.ctypes.data, elems) == 0
        assert ref.L.ref_pages_write(pages, pg, 1, vb.ctypes.data, elems) == 0
    Qd = to_dev(Q)
    o_ref = torch.zeros_like(Qd)
    q_abs_base = 0 if causal else t - 1
    assert ref.L.ref_attn(p(Qd), p(o_ref), mq, t, hd, Dq, Dkv, causal, q_abs_base, layer, pages, ps_ref, layers_n) == 0
    # --- B200 pool
    pool = layers.KvPool(n_pages=max(64, 4 * (t // 16 + 2)), page_size=16, n_layers=layers_n, n_kv_heads=n_kv, head_dim=hd, max_seqs=4)
    slot_id = 2
    pos_all = torch.arange(t, dtype=torch.int32, device="cuda")
    slot_all = torch.full((t,), slot_id, dtype=torch.int32, device="cuda")
    pool.store(layer, to_dev(K), to_dev(V), pos_all, slot_all)
    if causal:
        pos = torch.arange(mq, dtype=torch.int32, device="cuda")
    else:
        pos = torch.full((mq,), t - 1, dtype=torch.int32, device="cuda")
    slot = torch.full((mq,), slot_id, dtype=torch.int32, device="cuda")
    o_new = torch.zeros_like(Qd)
    layers.launch_attn(Qd, o_new, pool, layer, pos, slot, n_q)
    torch.cuda.synchronize()
    got, want = to_host(o_new), to_host(o_ref)
    ref.L.ref_pages_free(pages)
    assert np.array_equal(got, want), f"max ulp diff {ulp_diff(got, want)}, mismatches {(got != want).sum()}/{got.size}"
    # CPU oracle (glibc expf differs from device expf in the last ulp -> 1e-2 bf16 tolerance)
    kv = oracle.kv_new(ps_ref, layers_n, Dkv)
    oracle.kv_store(kv, layer, 0, K, V)
    o_cpu = oracle.attention(Q, kv, n_q, n_kv, hd, t, causal, q_abs_base, layer)
    oracle.kv_free(kv)
    assert rel_err(got, o_cpu) < 1e-2


def _tie_logits(rng, vocab, levels):
    vals = rng.choice(np.linspace(-3, 3, levels).astype(np.float32), size=vocab)
    from util import f32_to_bf16
    return f32_to_bf16(vals)


@pytest.mark.parametrize("vocab,levels,k,temp", [(151936, 40, 1, 1.0), (151936, 9, 1, 0.7), (4096, 3, 1, 1.0),
                                                 (151936, 4000, 50, 0.7), (151936, 25, 50, 1.0), (777, 5, 7, 0.7),
                                                 (300, 2, 50, 0.7), (1, 1, 1, 1.0)])
def test_sampling_tiebreak_and_rng(layers, ref, oracle, vocab, levels, k, temp):
    """topk_temperature_softmax_sampling_kernel_bf16: heavy ties exercise the reference's
    arg-max tie-break (logit_decode.cu:15-33,182-223); k>1 exercises the XORWOW draw."""
    rng = np.random.default_rng(vocab + levels + k)
    for trial in range(3):
        lg = _tie_logits(rng, vocab, levels)
        ld = to_dev(lg)
        seed = 1234 + trial
        want = ref.L.ref_sample(p(ld), vocab, temp, k, seed, 0)
        got = int(layers.sample_topk_bf16(ld, vocab, temp, k, seed)[0])
        assert got == want
        assert oracle.sample_topk(lg, temp, k, seed) == want
        if k == 1:
            assert oracle.argmax_tiebreak(lg) == want


@pytest.mark.parametrize("k,step", [(50, 1), (50, 7), (7, 123), (2, 65537)])
def test_sampling_subsequence_argument(layers, ref, k, step):
    """sample_topk_bf16's last argument is the cuRAND SUBSEQUENCE in the reference (helpers.cuh:157-166 ->
    curand_init(seed, step, 0), logit_decode.cu:256-257), not a seed offset: nonzero values must pick the
    reference's token, and must differ from seed + step for at least one draw."""
    rng = np.random.default_rng(k * 1000 + step)
    vocab, differs = 151936, 0
    for trial in range(4):
        lg = _tie_logits(rng, vocab, 4000)
        ld = to_dev(lg)
        seed = 1234 + trial
        want = ref.L.ref_sample(p(ld), vocab, 0.7, k, seed, step)
        got = int(layers.sample_topk_bf16(ld, vocab, 0.7, k, seed, step)[0])
        assert got == want
        differs += int(layers.sample_topk_bf16(ld, vocab, 0.7, k, seed + step, 0)[0]) != want
    assert differs > 0 or k == 2


def test_sampling_degenerate(layers, oracle):
    from util import f32_to_bf16
    lg = f32_to_bf16(np.full(1000, -np.inf, np.float32))
    assert int(layers.sample_topk_bf16(to_dev(lg), 1000, 1.0, 1, 1)[0]) == -1 == oracle.sample_topk(lg, 1.0, 1, 1)
    assert int(layers.sample_topk_bf16(to_dev(lg), 1000, 1.0, 0, 1)[0]) == -1
