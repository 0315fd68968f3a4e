"""Python mirror of the driver-level C ABI (include/qie_b200.h, section 2).

Names follow the reference's host driver (/root/reference/layers/src/iengine.cu): an
Engine owns the weight blob + KV pages (main(), iengine.cu:226-482), a sequence is what
create_new_sequence() returns (iengine.cu:25-47), prefill()/decode_step() are llm() in
its two states (src/qwen_main.cu:74-247 / 250-404).  All compute happens in
libqie_b200.so; numpy arrays here are HOST buffers handed to the C ABI.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Config, EngineOpts, KvView, QieError, check  # noqa: F401

# Qwen2.5 shape table. Operator semantics are the reference's (q/k-norm,
# no bias, explicit lm_head, interleaved RoPE, eps 1e-4).
ARCHS = {
    "qwen2.5-0.5b": dict(hidden=896, inter=4864, layers=24, n_q=14, n_kv=2, head_dim=64, vocab=151936),
    "qwen2.5-1.5b": dict(hidden=1536, inter=8960, layers=28, n_q=12, n_kv=2, head_dim=128, vocab=151936),
    "qwen2.5-7b": dict(hidden=3584, inter=18944, layers=28, n_q=28, n_kv=4, head_dim=128, vocab=152064),
    # small shapes for tests
    "tiny": dict(hidden=128, inter=256, layers=2, n_q=4, n_kv=2, head_dim=64, vocab=512),
    "small": dict(hidden=256, inter=640, layers=3, n_q=4, n_kv=2, head_dim=64, vocab=4096),
    "small128": dict(hidden=256, inter=512, layers=2, n_q=2, n_kv=1, head_dim=128, vocab=2048),
}
REF_CONTEXT = 32786  # sic: /root/reference/layers/src/utills.cu:14


def make_config(arch, context=REF_CONTEXT, **over):
    d = dict(ARCHS[arch]) if isinstance(arch, str) else dict(arch)
    d.update(over)
    d.setdefault("context", context)
    return Config(**d)


def weight_bytes(cfg):
    """bf16 bytes read per decode step (all layers + norms + lm_head), SURVEY 8d."""
    H, I, L, hd = cfg.hidden, cfg.inter, cfg.layers, cfg.head_dim
    Dq, Dkv = cfg.n_q * hd, cfg.n_kv * hd
    return 2 * (L * (H * Dq + 2 * H * Dkv + Dq * H + 3 * H * I + 2 * H + 2 * hd) + H + cfg.vocab * H)


def kv_bytes_per_pos(cfg):
    return 2 * cfg.layers * cfg.n_kv * cfg.head_dim * 2


def write_synthetic_checkpoint(cfg, seed, meta_path, weights_path):
    check(_lib.lib().qie_synth_checkpoint_write(C.byref(cfg), seed, str(meta_path).encode(),
                                                str(weights_path).encode()))


def convert_safetensors(shard_paths, meta_path, weights_path, tie_lm_head=False):
    """host-only: HF safetensors shards -> weights.bin + meta_data.txt in the reference's layout
    (parsed_tensors(), tensor_parser.cpp:31-129).  Returns (total bytes, tensor count)."""
    arr = (C.c_char_p * len(shard_paths))(*[str(p).encode() for p in shard_paths])
    tot, n = C.c_size_t(), C.c_int()
    check(_lib.lib().qie_convert_safetensors(arr, len(shard_paths), str(meta_path).encode(), str(weights_path).encode(),
                                             int(tie_lm_head), C.byref(tot), C.byref(n)))
    return tot.value, n.value


def inspect_checkpoint(meta_path, head_dim_hint=0):
    """host-only: (Config, total weight bytes, tensor count) of a meta_data.txt"""
    cfg, tot, n = Config(), C.c_size_t(), C.c_int()
    check(_lib.lib().qie_checkpoint_inspect(str(meta_path).encode(), head_dim_hint, C.byref(cfg), C.byref(tot),
                                            C.byref(n)))
    return cfg, tot.value, n.value


def tp_plan(cfg, tp_rank, tp_size):
    """host-only: the shard of rank tp_rank (see qie_tp_plan in include/qie_b200.h)"""
    out = (C.c_int * 8)()
    check(_lib.lib().qie_tp_plan(C.byref(cfg), tp_rank, tp_size, out))
    keys = ("n_q", "n_kv", "inter", "vocab", "q_row0", "kv_row0", "inter0", "vocab0")
    return dict(zip(keys, list(out)))


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


class Engine:
    def __init__(self, meta_path=None, weights_path=None, *, synthetic=None, seed=1234, device=0,
                 page_size=16, kv_bytes=0, max_pages=0, max_seqs=64, max_batch_tokens=256,
                 context=REF_CONTEXT, use_graph=True, head_dim_hint=0, numerics="reference_order",
                 tp_rank=0, tp_size=1, semantics="reference", rms_eps=0.0):
        L = _lib.lib()
        o = EngineOpts()
        L.qie_engine_opts_default(C.byref(o))
        o.device, o.page_size, o.kv_bytes, o.max_pages = device, page_size, kv_bytes, max_pages
        o.max_seqs, o.max_batch_tokens, o.context = max_seqs, max_batch_tokens, context
        o.use_graph, o.head_dim_hint = int(use_graph), head_dim_hint
        o.numerics = {"reference_order": 0, "fast": 1}[numerics]
        o.tp_rank, o.tp_size = tp_rank, tp_size
        o.semantics, o.rms_eps = {"reference": 0, "hf": 1}[semantics], rms_eps
        self.numerics = numerics
        h = C.c_void_p()
        if synthetic is not None:
            cfg = synthetic if isinstance(synthetic, Config) else make_config(synthetic, context=context)
            check(L.qie_engine_create_synthetic(C.byref(cfg), seed, C.byref(o), C.byref(h)))
        else:
            check(L.qie_engine_create(str(meta_path).encode(), str(weights_path).encode(), C.byref(o),
                                      C.byref(h)))
        self._h = h
        self._L = L
        cfg = Config()
        check(L.qie_engine_get_config(h, C.byref(cfg)))
        self.config = cfg

    # -- lifecycle -------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.qie_engine_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- tensor parallel (one process per GPU) ---------------------------
    @staticmethod
    def tp_unique_id():
        """128-byte NCCL id; rank 0 creates it and ships it to the other ranks."""
        buf = C.create_string_buffer(128)
        check(_lib.lib().qie_tp_unique_id(buf))
        return buf.raw

    def tp_connect(self, id128):
        assert len(id128) == 128
        check(self._L.qie_engine_tp_connect(self._h, C.create_string_buffer(bytes(id128), 128)))

    # -- configuration ---------------------------------------------------------
    def set_sampling(self, topk=1, temperature_prefill=1.0, temperature_decode=0.7, seed=1234, add_step=True):
        check(self._L.qie_engine_set_sampling(self._h, topk, temperature_prefill, temperature_decode, seed,
                                              int(add_step)))

    def set_repetition_penalty(self, penalty):
        check(self._L.qie_engine_set_repetition_penalty(self._h, float(penalty)))

    @property
    def stream(self):
        return self._L.qie_engine_stream(self._h)

    def kv_view(self):
        v = KvView()
        check(self._L.qie_engine_kv_view(self._h, C.byref(v)))
        return v

    def weight_ptr(self, short_name, layer=-1):
        n = C.c_size_t()
        p = self._L.qie_engine_weight(self._h, short_name.encode(), layer, C.byref(n))
        return p, n.value

    # -- sequences -------------------------------------------------------------
    def new_sequence(self):
        s = C.c_int()
        check(self._L.qie_seq_new(self._h, C.byref(s)))
        return s.value

    def free_sequence(self, seq):
        check(self._L.qie_seq_free(self._h, seq))

    def swap_out(self, seq):
        check(self._L.qie_seq_swap_out(self._h, seq))

    def swap_in(self, seq):
        check(self._L.qie_seq_swap_in(self._h, seq))

    def seq_len(self, seq):
        return check(self._L.qie_seq_len(self._h, seq))

    def pages_free(self):
        return check(self._L.qie_kv_pages_free(self._h))

    # -- forward ---------------------------------------------------------------
    def prefill(self, seq, ids):
        ids = _i32(ids)
        tok = np.zeros(1, np.int32)
        check(self._L.qie_prefill(self._h, seq, ids.ctypes.data, len(ids), tok.ctypes.data))
        return int(tok[0])

    def decode_step(self, seqs, tokens_in):
        seqs, tokens_in = _i32(seqs), _i32(tokens_in)
        out = np.zeros(len(seqs), np.int32)
        check(self._L.qie_decode_step(self._h, seqs.ctypes.data, tokens_in.ctypes.data, len(seqs),
                                      out.ctypes.data))
        return out

    def decode_run(self, seqs, tokens_in, steps):
        seqs, tokens_in = _i32(seqs), _i32(tokens_in)
        out = np.zeros((steps, len(seqs)), np.int32)
        check(self._L.qie_decode_run(self._h, seqs.ctypes.data, tokens_in.ctypes.data, len(seqs), steps,
                                     out.ctypes.data))
        return out

    def decode_step_device(self, seqs):
        seqs = _i32(seqs)
        check(self._L.qie_decode_step_device(self._h, seqs.ctypes.data, len(seqs)))

    def sync(self):
        check(self._L.qie_sync(self._h))

    def fill_synthetic(self, seq, n_pos, seed=99):
        """append n_pos positions of seeded random K/V without running the model (bench only)"""
        check(self._L.qie_seq_fill_synthetic(self._h, seq, n_pos, seed))

    def kv_read(self, seq, pos0, n):
        """cache rows [pos0, pos0+n) of a sequence as (K, V), each uint16 [n, layers, n_kv*head_dim]
        -- the reference's page layout [position][layer][kv_dim]"""
        c = self.config
        K = np.zeros((n, c.layers, c.n_kv * c.head_dim), np.uint16)
        V = np.zeros_like(K)
        check(self._L.qie_seq_kv_read(self._h, seq, pos0, n, K.ctypes.data, V.ctypes.data))
        return K, V

    def kv_write(self, seq, pos0, K, V):
        K, V = np.ascontiguousarray(K, np.uint16), np.ascontiguousarray(V, np.uint16)
        check(self._L.qie_seq_kv_write(self._h, seq, pos0, K.shape[0], K.ctypes.data, V.ctypes.data))

    def decode_step_profile(self, seqs, tokens_in):
        """one eager decode step with an event pair around every launch ->
        {kernel class: (summed ms, launches)}"""
        seqs, tokens_in = _i32(seqs), _i32(tokens_in)
        ms = np.zeros(32, np.float32)
        cnt = np.zeros(32, np.int32)
        check(self._L.qie_decode_step_profile(self._h, seqs.ctypes.data, tokens_in.ctypes.data, len(seqs),
                                              ms.ctypes.data, cnt.ctypes.data, 32))
        out = {}
        for i in range(32):
            name = self._L.qie_kernel_kind_name(i)
            if name is None:
                break
            out[name.decode()] = (float(ms[i]), int(cnt[i]))
        return out

    # -- persistent decode kernel ---------------------------------------------------
    def set_int(self, key, value):
        check(self._L.qie_engine_set_int(self._h, key.encode(), int(value)))

    def uses_mega(self, n_rows, kv_len):
        return bool(self._L.qie_decode_uses_mega(self._h, n_rows, kv_len))

    def mega_prof(self):
        """device timestamps (ns) of the last profiled persistent-kernel step"""
        n = 16 * self.config.layers + 8
        out = np.zeros(2 * n + 16, np.uint64)
        check(self._L.qie_mega_prof_read(self._h, out.ctypes.data, len(out)))
        self.mega_gemm_cycles = out[2 * n:2 * n + 10]  # per GEMM phase kind: (cycles waiting for weights, cycles in the MMA loop)
        self.mega_attn_cycles = out[2 * n + 10:2 * n + 16]  # attention of CTA 0: setup, scores, softmax, PV; [4], [5]: tile issue + wait inside scores / PV
        return out[:n - 2], out[n:2 * n - 2]  # (ns, SM cycles) at the 16*layers + 6 stamp points

    def read_activation(self, name, n_elems, dtype=np.uint16):
        out = np.zeros(n_elems, dtype)
        check(self._L.qie_engine_read_activation(self._h, name.encode(), out.ctypes.data, out.nbytes))
        return out

    def write_activation(self, name, values):
        values = np.ascontiguousarray(values)
        check(self._L.qie_engine_write_activation(self._h, name.encode(), values.ctypes.data, values.nbytes))

    def launch_count(self):
        return self._L.qie_launch_count(self._h)

    # -- parity hooks ------------------------------------------------------------
    def capture(self, on=True):
        check(self._L.qie_capture_enable(self._h, int(on)))

    def read_capture(self, tag, layer, max_elems=1 << 26):
        probe = np.zeros(1, np.uint16)
        n = check(self._L.qie_capture_read(self._h, tag.encode(), layer, probe.ctypes.data, 1))
        out = np.zeros(n, np.uint16)
        check(self._L.qie_capture_read(self._h, tag.encode(), layer, out.ctypes.data, n))
        return out

    def generate(self, ids, n_new):
        """greedy/top-k generation of one sequence: prefill + n_new-1 decode steps."""
        s = self.new_sequence()
        try:
            toks = [self.prefill(s, ids)]
            if n_new > 1:
                out = self.decode_run([s], [toks[0]], n_new - 1)
                toks += [int(t) for t in out[:, 0]]
            return toks
        finally:
            self.free_sequence(s)


class Scheduler:
    """continuous batching over an Engine (qie_sched_* in include/qie_b200.h): submit() queues a request,
    step() admits + runs one batched decode step, run() drains the queue."""

    def __init__(self, engine, max_running=64, eos_token=151645):
        self._L, self._eng = engine._L, engine
        h = C.c_void_p()
        check(self._L.qie_sched_create(engine._h, max_running, eos_token, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.qie_sched_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit(self, ids, max_new_tokens):
        ids = _i32(ids)
        rid = C.c_int()
        check(self._L.qie_sched_submit(self._h, ids.ctypes.data, len(ids), max_new_tokens, C.byref(rid)))
        return rid.value

    def step(self):
        return check(self._L.qie_sched_step(self._h))

    def result(self, rid, max_tokens=1 << 16):
        out = np.zeros(max_tokens, np.int32)
        fin = C.c_int()
        n = check(self._L.qie_sched_result(self._h, rid, out.ctypes.data, max_tokens, C.byref(fin)))
        self.last_status = fin.value  # 0 queued / running, 1 done, negative: the QIE_E* code that failed the request
        return [int(t) for t in out[:min(n, max_tokens)]], bool(fin.value)

    def stats(self):
        st, rows, pf = C.c_long(), C.c_long(), C.c_long()
        run, wait = C.c_int(), C.c_int()
        check(self._L.qie_sched_stats(self._h, C.byref(st), C.byref(rows), C.byref(pf), C.byref(run), C.byref(wait)))
        return dict(steps=st.value, decode_rows=rows.value, prefills=pf.value, running=run.value, waiting=wait.value)

    def run(self, max_steps=1 << 20):
        """step until every request has finished; returns the number left (0)"""
        for _ in range(max_steps):
            if self.step() == 0:
                return 0
        raise RuntimeError("scheduler did not drain")
