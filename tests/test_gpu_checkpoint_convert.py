"""gpu: a checkpoint that went synthetic -> two safetensors shards -> qie_convert_safetensors drives the engine to
the same tokens as the original files (the tensors sit at different offsets: addressing is by name + [begin,end))."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pytestmark = pytest.mark.gpu


def test_engine_on_converted_checkpoint(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qwen_inference_engine_b200 as q
    from test_cpu_safetensors_convert import _synthetic_as_tensors, write_safetensors
    from util import prompt_ids
    cfg = q.make_config("small", context=256)
    meta0, w0, t = _synthetic_as_tensors(cfg, 11, tmp_path)
    names = list(t)
    a = {k: t[k] for k in names if ".layers.2." not in k and not k.startswith("lm_")}
    b = {k: t[k] for k in names if ".layers.2." in k or k.startswith("lm_")}
    s1, s2 = str(tmp_path / "a.safetensors"), str(tmp_path / "b.safetensors")
    write_safetensors(s1, a, order=list(reversed(list(a))))
    write_safetensors(s2, b)
    meta1, w1 = str(tmp_path / "meta1.txt"), str(tmp_path / "w1.bin")
    q.convert_safetensors([s1, s2], meta1, w1)
    assert open(meta1).read() != open(meta0).read()  # lm_head moved behind the layers of shard 1
    ids = prompt_ids(12, cfg.vocab)
    outs = []
    for m, w in ((meta0, w0), (meta1, w1)):
        with q.Engine(m, w, context=256, max_seqs=2) as eng:
            outs.append(eng.generate(ids, 16))
    assert list(outs[0]) == list(outs[1])
