"""ctypes binding of libqie_b200.so (the C ABI in include/qie_b200.h).

The library is the product; this module is plumbing for tests, bench.py and Python users.
There is no fallback: if the shared library is missing, or a compute entry point fails,
an exception is raised.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QIE_B200_LIB") or os.path.join(_HERE, "libqie_b200.so")  # override: A/B builds of the library
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "qie_b200.h")


class QieError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"qie error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("hidden", "inter", "layers", "n_q", "n_kv", "head_dim", "vocab", "context")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class EngineOpts(C.Structure):
    _fields_ = [("device", C.c_int), ("page_size", C.c_int), ("max_pages", C.c_int),
                ("kv_bytes", C.c_size_t), ("max_seqs", C.c_int), ("max_batch_tokens", C.c_int),
                ("context", C.c_int), ("head_dim_hint", C.c_int), ("use_graph", C.c_int),
                ("tp_rank", C.c_int), ("tp_size", C.c_int), ("numerics", C.c_int),
                ("semantics", C.c_int), ("rms_eps", C.c_float)]


class KvView(C.Structure):
    _fields_ = [("pool", C.c_void_p), ("n_pages", C.c_int), ("page_size", C.c_int),
                ("n_layers", C.c_int), ("n_kv_heads", C.c_int), ("head_dim", C.c_int)]


def declared_symbols(header=HEADER_PATH):
    """Every function name declared in include/qie_b200.h."""
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qie_[a-z0-9_]+)\s*\(", src)))


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no Python/CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, sz, u64, f32 = C.c_void_p, C.c_int, C.c_size_t, C.c_uint64, C.c_float
    ip = C.POINTER(C.c_int)
    sig = {
        "qie_last_error": (C.c_char_p, []),
        "qie_abi_version": (i32, []),
        "qie_embedding": (i32, [vp, vp, vp, sz, sz, vp]),
        "qie_rmsnorm": (i32, [vp, vp, vp, sz, sz, vp]),
        "qie_matmul": (i32, [vp, vp, vp, i32, i32, i32, vp]),
        "qie_qknorm": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "qie_rope": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp]),
        "qie_precompute_cos_sin": (i32, [vp, vp, i32, i32]),
        "qie_silu": (i32, [vp, sz, vp]),
        "qie_elem_mul": (i32, [vp, vp, vp, sz, vp]),
        "qie_residual_add": (i32, [vp, vp, sz, vp]),
        "qie_kv_store": (i32, [C.POINTER(KvView), i32, vp, vp, vp, vp, vp, i32, i32, vp]),
        "qie_attention": (i32, [C.POINTER(KvView), i32, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
        "qie_attention_pagelist": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, i32, vp]),
        "qie_kv_store_pagelist": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp, i32, i32, vp]),
        "qie_sample_topk": (i32, [vp, vp, i32, sz, f32, i32, u64, u64, vp]),
        "qie_sample_topk_subseq": (i32, [vp, vp, i32, sz, f32, i32, u64, u64, u64, vp]),
        "qie_repetition_penalty": (i32, [vp, vp, sz, i32, f32, vp]),
        "qie_engine_set_repetition_penalty": (i32, [vp, f32]),
        "qie_matmul_fast": (i32, [vp, vp, vp, i32, i32, i32, vp]),
        "qie_attention_prefill_tc": (i32, [C.POINTER(KvView), i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
        "qie_attention_prefill_fast": (i32, [C.POINTER(KvView), i32, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
        "qie_attention_decode_fast": (i32, [C.POINTER(KvView), i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
        "qie_engine_opts_default": (None, [C.POINTER(EngineOpts)]),
        "qie_synth_checkpoint_write": (i32, [C.POINTER(Config), u64, C.c_char_p, C.c_char_p]),
        "qie_convert_safetensors": (i32, [C.POINTER(C.c_char_p), i32, C.c_char_p, C.c_char_p, i32, C.POINTER(sz), C.POINTER(i32)]),
        "qie_checkpoint_inspect": (i32, [C.c_char_p, i32, C.POINTER(Config), C.POINTER(sz), C.POINTER(i32)]),
        "qie_engine_create": (i32, [C.c_char_p, C.c_char_p, C.POINTER(EngineOpts), C.POINTER(vp)]),
        "qie_engine_create_from_blob": (i32, [C.c_char_p, vp, C.POINTER(EngineOpts), C.POINTER(vp)]),
        "qie_engine_create_synthetic": (i32, [C.POINTER(Config), u64, C.POINTER(EngineOpts), C.POINTER(vp)]),
        "qie_engine_destroy": (None, [vp]),
        "qie_tp_plan": (i32, [C.POINTER(Config), i32, i32, ip]),
        "qie_tp_unique_id": (i32, [vp]),
        "qie_engine_tp_connect": (i32, [vp, vp]),
        "qie_engine_get_config": (i32, [vp, C.POINTER(Config)]),
        "qie_engine_weight": (vp, [vp, C.c_char_p, i32, C.POINTER(sz)]),
        "qie_engine_kv_view": (i32, [vp, C.POINTER(KvView)]),
        "qie_engine_stream": (vp, [vp]),
        "qie_engine_limits": (i32, [vp, ip, ip, ip]),
        "qie_engine_set_sampling": (i32, [vp, i32, f32, f32, u64, i32]),
        "qie_seq_new": (i32, [vp, ip]),
        "qie_seq_free": (i32, [vp, i32]),
        "qie_seq_len": (i32, [vp, i32]),
        "qie_seq_swap_out": (i32, [vp, i32]),
        "qie_seq_swap_in": (i32, [vp, i32]),
        "qie_kv_pages_free": (i32, [vp]),
        "qie_sched_create": (i32, [vp, i32, i32, C.POINTER(vp)]),
        "qie_sched_destroy": (None, [vp]),
        "qie_sched_submit": (i32, [vp, vp, i32, i32, ip]),
        "qie_sched_step": (i32, [vp]),
        "qie_sched_result": (i32, [vp, i32, vp, i32, ip]),
        "qie_sched_stats": (i32, [vp, C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long), ip, ip]),
        "qie_prefill": (i32, [vp, i32, vp, i32, vp]),
        "qie_decode_step": (i32, [vp, vp, vp, i32, vp]),
        "qie_decode_run": (i32, [vp, vp, vp, i32, i32, vp]),
        "qie_decode_step_device": (i32, [vp, vp, i32]),
        "qie_sync": (i32, [vp]),
        "qie_capture_enable": (i32, [vp, i32]),
        "qie_capture_read": (C.c_long, [vp, C.c_char_p, i32, vp, sz]),
        "qie_launch_count": (C.c_long, [vp]),
        "qie_seq_fill_synthetic": (i32, [vp, i32, i32, u64]),
        "qie_seq_kv_read": (i32, [vp, i32, i32, i32, vp, vp]),
        "qie_seq_kv_write": (i32, [vp, i32, i32, i32, vp, vp]),
        "qie_decode_step_profile": (i32, [vp, vp, vp, i32, vp, vp, i32]),
        "qie_kernel_kind_name": (C.c_char_p, [i32]),
        "qie_engine_set_int": (i32, [vp, C.c_char_p, C.c_long]),
        "qie_decode_uses_mega": (i32, [vp, i32, i32]),
        "qie_mega_prof_read": (C.c_long, [vp, vp, sz]),
        "qie_engine_read_activation": (C.c_long, [vp, C.c_char_p, vp, sz]),
        "qie_engine_write_activation": (C.c_long, [vp, C.c_char_p, vp, sz]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._qie_signatures = sig
    _lib = L
    return L


def check(rc):
    if rc < 0:
        raise QieError(rc, lib().qie_last_error().decode("utf-8", "replace"))
    return rc
