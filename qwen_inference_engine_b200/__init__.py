"""qwen_inference_engine_b200 -- B200-native (sm_100a) decode/prefill forward behind the
operator API of Rafae1130/qwen_inference_engine.

The product is the C++/CUDA shared library `libqie_b200.so` (C ABI: include/qie_b200.h).
This package only binds it: `engine` mirrors the reference's host driver, `layers` its
`launch_*` operator wrappers.  Importing the package never touches oracle/.
"""
from ._lib import LIB_PATH, QieError, declared_symbols, lib  # noqa: F401
from .engine import (ARCHS, REF_CONTEXT, Config, Engine, Scheduler, convert_safetensors, inspect_checkpoint, kv_bytes_per_pos,  # noqa: F401
                     make_config, tp_plan, weight_bytes, write_synthetic_checkpoint)

__all__ = ["Engine", "Scheduler", "Config", "make_config", "ARCHS", "QieError", "lib", "LIB_PATH", "weight_bytes",
           "kv_bytes_per_pos", "tp_plan", "convert_safetensors", "inspect_checkpoint", "write_synthetic_checkpoint", "declared_symbols", "REF_CONTEXT"]
