"""Operator-level mirror of the reference's `layers/` API.

Same names and argument meaning as /root/reference/layers/include/helpers.cuh:45-166
(launch_rms, launch_rope, launch_rope_single, launch_matmul, launch_elem, launch_act,
launch_resadd, launch_attn, launch_qknorm, sample_topk_bf16), so a parity test reads like
the reference's call site in llm() (src/qwen_main.cu).  Tensors are torch CUDA tensors
(torch is plumbing for device memory and the stream); the work is done by
libqie_b200.so through its C ABI.  bf16 tensors must be contiguous.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import KvView, check


def _p(t):
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensor expected"
    return t.data_ptr()


def _st():
    return torch.cuda.current_stream().cuda_stream


def embedding(out, table, ids):
    """embedding_matrix_func (src/embedded_matrix.cu:5-17)"""
    check(_lib.lib().qie_embedding(_p(out), _p(table), _p(ids), table.shape[1], ids.numel(), _st()))


def launch_rms(x, w, y, hidden, seqlen):
    check(_lib.lib().qie_rmsnorm(_p(x), _p(w), _p(y), hidden, seqlen, _st()))


def launch_matmul(A, B, Cout, M, N, K):
    """C[M,K] = A[M,N] @ B[K,N]^T -- N is the inner dimension, as in the reference."""
    check(_lib.lib().qie_matmul(_p(A), _p(B), _p(Cout), M, N, K, _st()))


def launch_qknorm(X, w, head_dim, seqlen, hidden, nheads):
    check(_lib.lib().qie_qknorm(_p(X), _p(w), head_dim, seqlen, hidden, nheads, _st()))


def launch_rope(cos_d, sin_d, x, seqlen, head_dim, hidden_dim, nheads):
    check(_lib.lib().qie_rope(_p(cos_d), _p(sin_d), _p(x), seqlen, 0, head_dim, hidden_dim, nheads, _st()))


def launch_rope_single(cos_d, sin_d, x, pos, head_dim, hidden_dim, nheads):
    check(_lib.lib().qie_rope(_p(cos_d), _p(sin_d), _p(x), 1, pos, head_dim, hidden_dim, nheads, _st()))


def launch_act(x, n):
    check(_lib.lib().qie_silu(_p(x), n, _st()))


def launch_elem(a, b, out, n):
    check(_lib.lib().qie_elem_mul(_p(a), _p(b), _p(out), n, _st()))


def launch_resadd(x, y, n):
    check(_lib.lib().qie_residual_add(_p(x), _p(y), n, _st()))


def precompute_cos_sin(seq_len, head_dim):
    """src/include.cpp:5-18 -- host fp32 tables [seq_len, head_dim/2]"""
    c = np.zeros((seq_len, head_dim // 2), np.float32)
    s = np.zeros((seq_len, head_dim // 2), np.float32)
    check(_lib.lib().qie_precompute_cos_sin(c.ctypes.data, s.ctypes.data, seq_len, head_dim))
    return c, s


class KvPool:
    """B200 KV pool pool[page][layer][k|v][kv_head][slot][head_dim] plus a block table; the
    operator-level stand-in for create_page_list / allocate_page_buffers (iengine.cu:73-96)."""

    def __init__(self, n_pages, page_size, n_layers, n_kv_heads, head_dim, max_seqs=8):
        self.pool = torch.zeros(n_pages, n_layers, 2, n_kv_heads, page_size, head_dim, dtype=torch.bfloat16,
                                device="cuda")
        self.view = KvView(self.pool.data_ptr(), n_pages, page_size, n_layers, n_kv_heads, head_dim)
        self.max_pages = n_pages
        # identity-ish block table: sequence s owns pages s, s+max_seqs, ... (deliberately non-contiguous)
        bt = np.zeros((max_seqs, n_pages), np.int32)
        for s in range(max_seqs):
            ids = list(range(s, n_pages, max_seqs))
            bt[s, :len(ids)] = ids
        self.block_table = torch.from_numpy(bt).cuda()

    def store(self, layer, K, V, pos, slot):
        check(_lib.lib().qie_kv_store(C.byref(self.view), layer, _p(K), _p(V), _p(pos), _p(slot),
                                      _p(self.block_table), self.max_pages, pos.numel(), _st()))


def launch_attn(Q, out, pool, layer, pos, slot, n_q_heads):
    """selfattention over the paged cache; row t attends to positions 0..pos[t]."""
    check(_lib.lib().qie_attention(C.byref(pool.view), layer, _p(Q), _p(out), _p(pos), _p(slot),
                                   _p(pool.block_table), pool.max_pages, pos.numel(), n_q_heads, _st()))


def sample_topk_bf16(logits_d, vocab, temperature, topk, seed, step=0):
    """helpers.cuh:157-166 -- returns the sampled token ids (one per row of logits)."""
    rows = logits_d.numel() // vocab
    out = torch.zeros(rows, dtype=torch.int32, device="cuda")
    # the reference hands `step` to the kernel as the cuRAND subsequence (helpers.cuh:157-166)
    check(_lib.lib().qie_sample_topk_subseq(_p(logits_d), _p(out), rows, vocab, float(temperature), topk, seed, 0, step,
                                            _st()))
    return out.cpu().numpy()


def apply_repetition_penalty(logits_d, context_tokens_d, vocab, penalty):
    """layers_include.cuh:33 (declared, never defined in the reference): in place on one row of logits"""
    check(_lib.lib().qie_repetition_penalty(_p(logits_d), _p(context_tokens_d), context_tokens_d.numel(), vocab, float(penalty), _st()))


# ---- fast-numerics operators (tolerance parity, see include/qie_b200.h) ---------------------
def launch_matmul_fast(A, B, Cout, M, N, K):
    """tcgen05/TMEM/TMA GEMM: same contract as launch_matmul, results within 1e-2 (bf16)."""
    check(_lib.lib().qie_matmul_fast(_p(A), _p(B), _p(Cout), M, N, K, _st()))


def launch_attn_decode_fast(Q, out, pool, layer, pos, slot, n_q_heads, n_splits=0):
    """split-KV flash-decoding: one query token per row."""
    check(_lib.lib().qie_attention_decode_fast(C.byref(pool.view), layer, _p(Q), _p(out), _p(pos), _p(slot),
                                               _p(pool.block_table), pool.max_pages, pos.numel(), n_q_heads,
                                               n_splits, _st()))


def launch_attn_prefill_fast(Q, out, pool, layer, pos, slot, n_q_heads):
    """causal tiled attention for prefill rows: consecutive positions pos[0]+t of one sequence."""
    check(_lib.lib().qie_attention_prefill_fast(C.byref(pool.view), layer, _p(Q), _p(out), _p(pos), _p(slot),
                                                _p(pool.block_table), pool.max_pages, pos.numel(), n_q_heads,
                                                _st()))


def launch_attn_prefill_tc(Q, out, pool, layer, pos, slot, n_q_heads, variant=0):
    """the same on tcgen05 / TMEM (head_dim 128)."""
    check(_lib.lib().qie_attention_prefill_tc(C.byref(pool.view), layer, _p(Q), _p(out), _p(pos), _p(slot),
                                              _p(pool.block_table), pool.max_pages, pos.numel(), n_q_heads, variant,
                                              _st()))
