"""ctypes wrappers for the test oracles.  TEST INFRASTRUCTURE ONLY.

  * `Oracle`  -> oracle/libqie_oracle.so, the plain-C CPU restatement (qie_oracle.c)
  * `Ref`     -> oracle/_ref/libqie_ref.so, the reference's OWN CUDA kernels compiled from
                 /root/reference/layers/src (oracle/Makefile) + ref_driver.cu. Needs a GPU.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  The product package never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libqie_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libqie_ref.so")


class OrcConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("hidden", "inter", "layers", "n_q", "n_kv", "head_dim", "vocab", "context")]


def bf16_to_f32(a):
    a = np.ascontiguousarray(a, dtype=np.uint16)
    return (a.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(a):
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = u + 0x7FFF + ((u >> 16) & 1)
    return (u >> 16).astype(np.uint16)


def _u16(a):
    return np.ascontiguousarray(a, dtype=np.uint16)


_DUMP_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_uint16), C.c_size_t)


class Oracle:
    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `make -C oracle libqie_oracle.so`")
        L = self.L = C.CDLL(path)
        vp, i32, sz, u64, f32 = C.c_void_p, C.c_int, C.c_size_t, C.c_uint64, C.c_float
        L.orc_precompute_cos_sin.argtypes = [vp, vp, i32, i32]
        L.orc_embedding.argtypes = [vp, vp, vp, sz, sz]
        L.orc_rmsnorm.argtypes = [vp, vp, vp, sz, sz]
        L.orc_matmul.argtypes = [vp, vp, vp, i32, i32, i32]
        L.orc_qknorm.argtypes = [vp, vp, i32, i32, i32, i32]
        L.orc_rope.argtypes = [vp, vp, vp, i32, i32, i32, i32]
        L.orc_silu.argtypes = [vp, sz]
        L.orc_elem_mul.argtypes = [vp, vp, vp, sz]
        L.orc_residual_add.argtypes = [vp, vp, sz]
        L.orc_kv_new.restype = vp
        L.orc_kv_new.argtypes = [i32, i32, i32]
        L.orc_kv_free.argtypes = [vp]
        L.orc_kv_store.argtypes = [vp, i32, i32, i32, vp, vp]
        L.orc_attention.argtypes = [vp, vp, sz, sz, sz, sz, sz, i32, i32, i32, sz, i32, vp]
        L.orc_sample_topk.restype = i32
        L.orc_sample_topk.argtypes = [vp, f32, i32, sz, u64, u64]
        L.orc_argmax_ref_tiebreak.restype = i32
        L.orc_argmax_ref_tiebreak.argtypes = [vp, sz]
        L.orc_repetition_penalty.argtypes = [vp, vp, sz, i32, f32]
        L.orc_model_load.restype = vp
        L.orc_model_load.argtypes = [C.c_char_p, C.c_char_p, i32, i32]
        L.orc_model_free.argtypes = [vp]
        L.orc_model_set_semantics.argtypes = [vp, i32, f32]
        L.orc_model_config.restype = C.POINTER(OrcConfig)
        L.orc_model_config.argtypes = [vp]
        L.orc_model_tensor.restype = vp
        L.orc_model_tensor.argtypes = [vp, C.c_char_p, i32, C.POINTER(sz)]
        L.orc_seq_new.restype = vp
        L.orc_seq_new.argtypes = [vp, i32]
        L.orc_seq_free.argtypes = [vp]
        L.orc_seq_len.argtypes = [vp]
        L.orc_set_dump.argtypes = [vp, _DUMP_FN, vp]
        L.orc_prefill.restype = i32
        L.orc_prefill.argtypes = [vp, vp, i32, i32, f32, u64, vp]
        L.orc_decode.restype = i32
        L.orc_decode.argtypes = [vp, i32, i32, f32, u64, vp]
        L.orc_set_threads.argtypes = [i32]
        L.orc_synth_value.restype = C.c_uint16
        L.orc_synth_value.argtypes = [u64, u64, i32]
        L.orc_synth_write.restype = i32
        L.orc_synth_write.argtypes = [C.POINTER(OrcConfig), u64, C.c_char_p, C.c_char_p]

    # ---- operators (numpy uint16 = raw bf16) ----
    def cos_sin(self, seq_len, head_dim):
        c = np.zeros((seq_len, head_dim // 2), np.float32)
        s = np.zeros((seq_len, head_dim // 2), np.float32)
        self.L.orc_precompute_cos_sin(c.ctypes.data, s.ctypes.data, seq_len, head_dim)
        return c, s

    def embedding(self, table, ids):
        table, ids = _u16(table), np.ascontiguousarray(ids, np.int32)
        out = np.zeros((len(ids), table.shape[1]), np.uint16)
        self.L.orc_embedding(out.ctypes.data, table.ctypes.data, ids.ctypes.data, table.shape[1], len(ids))
        return out

    def rmsnorm(self, x, w):
        x, w = _u16(x), _u16(w)
        y = np.zeros_like(x)
        self.L.orc_rmsnorm(x.ctypes.data, w.ctypes.data, y.ctypes.data, x.shape[-1], x.size // x.shape[-1])
        return y

    def matmul(self, A, B):
        A, B = _u16(A), _u16(B)
        M, N = A.shape
        K = B.shape[0]
        Cm = np.zeros((M, K), np.uint16)
        self.L.orc_matmul(A.ctypes.data, B.ctypes.data, Cm.ctypes.data, M, N, K)
        return Cm

    def qknorm(self, x, w, head_dim, n_heads):
        x = _u16(x).copy()
        w = _u16(w)
        self.L.orc_qknorm(x.ctypes.data, w.ctypes.data, head_dim, x.shape[0], x.shape[1], n_heads)
        return x

    def rope(self, cos, sin, x, pos0, head_dim, n_heads):
        x = _u16(x).copy()
        half = head_dim // 2
        c = np.ascontiguousarray(cos[pos0:], np.float32)
        s = np.ascontiguousarray(sin[pos0:], np.float32)
        assert c.shape[1] == half
        self.L.orc_rope(c.ctypes.data, s.ctypes.data, x.ctypes.data, x.shape[0], head_dim, x.shape[1], n_heads)
        return x

    def silu(self, x):
        x = _u16(x).copy()
        self.L.orc_silu(x.ctypes.data, x.size)
        return x

    def elem_mul(self, a, b):
        a, b = _u16(a), _u16(b)
        c = np.zeros_like(a)
        self.L.orc_elem_mul(a.ctypes.data, b.ctypes.data, c.ctypes.data, a.size)
        return c

    def residual_add(self, a, b):
        a, b = _u16(a).copy(), _u16(b)
        self.L.orc_residual_add(a.ctypes.data, b.ctypes.data, a.size)
        return a

    def attention(self, Q, kv, n_q, n_kv, head_dim, seq_len_kv, causal, q_abs_base, layer):
        Q = _u16(Q)
        out = np.zeros_like(Q)
        self.L.orc_attention(Q.ctypes.data, out.ctypes.data, Q.shape[0], seq_len_kv, head_dim, n_q * head_dim,
                             n_kv * head_dim, n_q, n_kv, int(causal), q_abs_base, layer, kv)
        return out

    def kv_new(self, page_size, n_layers, kv_dim):
        return self.L.orc_kv_new(page_size, n_layers, kv_dim)

    def kv_store(self, kv, layer, pos0, K, V):
        K, V = _u16(K), _u16(V)
        self.L.orc_kv_store(kv, layer, pos0, K.shape[0], K.ctypes.data, V.ctypes.data)

    def kv_free(self, kv):
        self.L.orc_kv_free(kv)

    def sample_topk(self, logits, temperature, k, seed, subseq=0):
        logits = _u16(logits)
        return self.L.orc_sample_topk(logits.ctypes.data, temperature, k, logits.size, seed, subseq)

    def argmax_tiebreak(self, logits):
        logits = _u16(logits)
        return self.L.orc_argmax_ref_tiebreak(logits.ctypes.data, logits.size)

    def repetition_penalty(self, logits, context_tokens, penalty):
        lg = _u16(logits).copy()
        ctx = np.ascontiguousarray(context_tokens, np.int32)
        self.L.orc_repetition_penalty(lg.ctypes.data, ctx.ctypes.data, len(ctx), lg.size, penalty)
        return lg

    def synth_value(self, seed, g, kind):
        return self.L.orc_synth_value(seed, g, kind)

    def synth_write(self, cfg, seed, meta_path, weights_path):
        c = OrcConfig(**{n: getattr(cfg, n) for n, _ in OrcConfig._fields_})
        rc = self.L.orc_synth_write(C.byref(c), seed, str(meta_path).encode(), str(weights_path).encode())
        if rc:
            raise IOError("orc_synth_write failed")

    def set_threads(self, n):
        self.L.orc_set_threads(n)


class OracleModel:
    """CPU forward: orc_prefill / orc_decode (llm() restated, qwen_main.cu:64-417)."""

    def __init__(self, oracle, meta_path, weights_path, head_dim_hint=0, context=32786, page_size=4):
        self.o = oracle
        self.m = oracle.L.orc_model_load(str(meta_path).encode(), str(weights_path).encode(), head_dim_hint, context)
        if not self.m:
            raise IOError(f"orc_model_load({meta_path}) failed")
        self.cfg = oracle.L.orc_model_config(self.m).contents
        self.page_size = page_size
        self._seqs = []

    def set_semantics(self, rope_half=False, eps=1e-4):
        """HF / Qwen2.5-Qwen3 checkpoints: rope_half=True, eps=1e-6 (the reference: interleaved RoPE, 1e-4)"""
        self.o.L.orc_model_set_semantics(self.m, int(rope_half), float(eps))
        return self

    def tensor(self, short_name, layer=-1):
        n = C.c_size_t()
        p = self.o.L.orc_model_tensor(self.m, short_name.encode(), layer, C.byref(n))
        if not p:
            return None
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint16)), shape=(n.value,))

    def new_seq(self):
        s = self.o.L.orc_seq_new(self.m, self.page_size)
        self._seqs.append(s)
        return s

    def set_dump(self, seq, store):
        """store: dict filled with {(tag, layer): np.uint16 array} during the next forwards"""
        def cb(user, tag, layer, data, n):
            store[(tag.decode(), layer)] = np.ctypeslib.as_array(data, shape=(n,)).copy()
        self._cb = _DUMP_FN(cb)
        self.o.L.orc_set_dump(seq, self._cb, None)

    def prefill(self, seq, ids, topk=1, temperature=1.0, seed=1234, want_logits=False):
        ids = np.ascontiguousarray(ids, np.int32)
        lg = np.zeros(self.cfg.vocab, np.uint16) if want_logits else None
        t = self.o.L.orc_prefill(seq, ids.ctypes.data, len(ids), topk, temperature, seed,
                                 lg.ctypes.data if want_logits else None)
        return (t, lg) if want_logits else t

    def decode(self, seq, token, topk=1, temperature=0.7, seed=1234, want_logits=False):
        lg = np.zeros(self.cfg.vocab, np.uint16) if want_logits else None
        t = self.o.L.orc_decode(seq, int(token), topk, temperature, seed, lg.ctypes.data if want_logits else None)
        return (t, lg) if want_logits else t

    def generate(self, ids, n_new, topk=1):
        s = self.new_seq()
        toks = [self.prefill(s, ids, topk=topk)]
        for i in range(n_new - 1):
            toks.append(self.decode(s, toks[-1], topk=topk, seed=1234 + 1 + i))
        return toks

    def close(self):
        for s in self._seqs:
            self.o.L.orc_seq_free(s)
        self._seqs = []
        if self.m:
            self.o.L.orc_model_free(self.m)
            self.m = None


# ------------------------------------------------------------------------------ reference kernels (GPU)
class RefLayerW(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("in_ln", "q", "k", "v", "o", "q_norm", "k_norm", "post_ln", "up", "gate", "down")]


class RefModelDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("hidden", "inter", "layers", "n_q", "n_kv", "head_dim", "vocab", "context")] + \
               [("embed", C.c_void_p), ("norm", C.c_void_p), ("lm_head", C.c_void_p),
                ("L", C.POINTER(RefLayerW))]


_TAP_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_size_t)


class Ref:
    """The reference's own kernels on the GPU. Pointers are raw device addresses (ints)."""

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        L = self.L = C.CDLL(path)
        vp, i32, sz, f32, u64 = C.c_void_p, C.c_int, C.c_size_t, C.c_float, C.c_ulonglong
        L.ref_embedding.argtypes = [vp, vp, vp, sz, sz]
        L.ref_rmsnorm.argtypes = [vp, vp, vp, sz, sz]
        L.ref_matmul.argtypes = [vp, vp, vp, i32, i32, i32]
        L.ref_qknorm.argtypes = [vp, vp, i32, i32, i32, i32]
        L.ref_rope.argtypes = [vp, vp, vp, sz, sz, sz, sz]
        L.ref_rope_single.argtypes = [vp, vp, vp, sz, sz, i32, i32]
        L.ref_act.argtypes = [vp, sz]
        L.ref_elem.argtypes = [vp, vp, vp, i32]
        L.ref_resadd.argtypes = [vp, vp, sz]
        L.ref_sample.argtypes = [vp, i32, f32, i32, u64, i32]
        L.ref_precompute_cos_sin.argtypes = [vp, vp, i32, i32]
        L.ref_pages_create.restype = vp
        L.ref_pages_create.argtypes = [i32, sz]
        L.ref_pages_free.argtypes = [vp]
        L.ref_pages_count.argtypes = [vp]
        L.ref_pages_read.argtypes = [vp, i32, i32, vp, sz]
        L.ref_pages_write.argtypes = [vp, i32, i32, vp, sz]
        L.ref_attn.argtypes = [vp, vp, sz, sz, sz, sz, sz, i32, sz, i32, vp, i32, i32]
        L.ref_seq_create.restype = vp
        L.ref_seq_create.argtypes = [C.POINTER(RefModelDesc), i32]
        L.ref_seq_destroy.argtypes = [vp]
        L.ref_seq_len.argtypes = [vp]
        L.ref_seq_pages.restype = vp
        L.ref_seq_pages.argtypes = [vp]
        L.ref_seq_read.argtypes = [vp, C.c_char_p, vp, sz]
        L.ref_seq_fake_context.argtypes = [vp, i32]
        L.ref_forward_prefill.argtypes = [vp, vp, i32, i32, f32, u64, _TAP_FN, vp]
        L.ref_forward_decode.argtypes = [vp, i32, i32, f32, u64, i32, _TAP_FN, vp]

    def model_desc(self, engine):
        """RefModelDesc pointing into a qwen_inference_engine_b200.Engine's weight blob (shared, read-only)."""
        cfg = engine.config
        layers = (RefLayerW * cfg.layers)()
        names = dict(in_ln="input_layernorm.weight", q="self_attn.q_proj.weight", k="self_attn.k_proj.weight",
                     v="self_attn.v_proj.weight", o="self_attn.o_proj.weight", q_norm="self_attn.q_norm.weight",
                     k_norm="self_attn.k_norm.weight", post_ln="post_attention_layernorm.weight",
                     up="mlp.up_proj.weight", gate="mlp.gate_proj.weight", down="mlp.down_proj.weight")
        for l in range(cfg.layers):
            for f, sn in names.items():
                setattr(layers[l], f, engine.weight_ptr(sn, l)[0])
        d = RefModelDesc(cfg.hidden, cfg.inter, cfg.layers, cfg.n_q, cfg.n_kv, cfg.head_dim, cfg.vocab, cfg.context,
                         engine.weight_ptr("embed_tokens.weight")[0], engine.weight_ptr("norm.weight")[0],
                         engine.weight_ptr("logits")[0], layers)
        d._keep = layers
        return d


class RefSeq:
    """One reference sequence: ref_forward_prefill / ref_forward_decode (llm() replay)."""

    def __init__(self, ref, desc, page_size=4):
        self.ref, self.desc = ref, desc
        self.h = ref.L.ref_seq_create(C.byref(desc), page_size)
        self._null_tap = C.cast(None, _TAP_FN)

    def _tap(self, store):
        if store is None:
            return self._null_tap
        import torch

        def cb(user, tag, layer, dev, n):
            t = torch.empty(n, dtype=torch.int16, device="cuda")
            torch.cuda.synchronize()
            # device->device copy out of the reference's buffer via a raw-pointer view
            src = _dev_view(dev, n)
            t.copy_(src)
            torch.cuda.synchronize()
            store[(tag.decode(), layer)] = t.cpu().numpy().view(np.uint16)
        self._cb = _TAP_FN(cb)
        return self._cb

    def prefill(self, ids, topk=1, temperature=1.0, seed=1234, taps=None):
        ids = np.ascontiguousarray(ids, np.int32)
        return self.ref.L.ref_forward_prefill(self.h, ids.ctypes.data, len(ids), topk, temperature, seed,
                                              self._tap(taps), None)

    def decode(self, token, topk=1, temperature=0.7, seed=1234, with_syncs=True, taps=None):
        return self.ref.L.ref_forward_decode(self.h, int(token), topk, temperature, seed, int(with_syncs),
                                             self._tap(taps), None)

    def fake_context(self, n_ctx):
        if self.ref.L.ref_seq_fake_context(self.h, n_ctx):
            raise RuntimeError("ref_seq_fake_context failed")

    def read(self, tag, n):
        out = np.zeros(n, np.uint16)
        rc = self.ref.L.ref_seq_read(self.h, tag.encode(), out.ctypes.data, n)
        if rc:
            raise KeyError(tag)
        return out

    def generate(self, ids, n_new, topk=1, with_syncs=True):
        toks = [self.prefill(ids, topk=topk)]
        for _ in range(n_new - 1):
            toks.append(self.decode(toks[-1], topk=topk, with_syncs=with_syncs))
        return toks

    def close(self):
        if self.h:
            self.ref.L.ref_seq_destroy(self.h)
            self.h = None


def _dev_view(ptr, n_elems):
    """torch int16 tensor aliasing raw device memory [ptr, ptr + 2*n_elems)."""
    import torch

    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n_elems,), "typestr": "<i2", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device="cuda")
