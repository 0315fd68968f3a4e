"""not-gpu: safetensors -> weights.bin + meta_data.txt converter.
The expected layout is derived here, independently, from the reference's parsed_tensors()
(/root/reference/layers/src/tensor_parser.cpp:31-129): shards in the given order, keys in byte-lexicographic
order (nlohmann::json = std::map), only "model." / "lm_" keys, running offsets, lm_* -> short_name "logits"."""
import json
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qwen_inference_engine_b200 as q  # noqa: E402


def write_safetensors(path, tensors, order=None, dtype="BF16", metadata=True):
    """tensors: {name: uint16 array}; data laid out in `order` (default: insertion order), header keys shuffled."""
    order = order or list(tensors)
    hdr, off, blobs = {}, 0, []
    for name in order:
        a = tensors[name]
        n = a.size * 2
        hdr[name] = {"dtype": dtype, "shape": list(a.shape), "data_offsets": [off, off + n]}
        blobs.append(a.tobytes())
        off += n
    keys = list(hdr)
    rng = np.random.default_rng(len(keys))
    rng.shuffle(keys)
    out = {"__metadata__": {"format": "pt", "note": 'a "quoted" value, with {braces} and [brackets]' }} if metadata else {}
    for k in keys:
        out[k] = hdr[k]
    js = json.dumps(out).encode()
    js += b" " * ((8 - len(js) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(js)))
        f.write(js)
        for b in blobs:
            f.write(b)


def hf_tensors(cfg, seed, with_lm_head=True):
    rng = np.random.default_rng(seed)
    H, I, hd = cfg.hidden, cfg.inter, cfg.head_dim
    Dq, Dkv = cfg.n_q * hd, cfg.n_kv * hd
    r = lambda *s: rng.integers(0, 1 << 16, size=s, dtype=np.uint16)  # noqa: E731
    t = {"model.embed_tokens.weight": r(cfg.vocab, H), "model.norm.weight": r(H)}
    for l in range(cfg.layers):
        p = f"model.layers.{l}."
        t[p + "input_layernorm.weight"] = r(H)
        t[p + "post_attention_layernorm.weight"] = r(H)
        t[p + "self_attn.q_proj.weight"] = r(Dq, H)
        t[p + "self_attn.k_proj.weight"] = r(Dkv, H)
        t[p + "self_attn.v_proj.weight"] = r(Dkv, H)
        t[p + "self_attn.o_proj.weight"] = r(H, Dq)
        t[p + "self_attn.q_norm.weight"] = r(hd)
        t[p + "self_attn.k_norm.weight"] = r(hd)
        t[p + "mlp.gate_proj.weight"] = r(I, H)
        t[p + "mlp.up_proj.weight"] = r(I, H)
        t[p + "mlp.down_proj.weight"] = r(H, I)
    if with_lm_head:
        t["lm_head.weight"] = r(cfg.vocab, H)
    t["rotary_emb.inv_freq"] = r(hd // 2)  # not "model." / "lm_": must be dropped
    return t


def expected_layout(shards):
    """[(name, layer, short_name, shape, begin, end, data)] as parsed_tensors() would number them"""
    out, off = [], 0
    for tensors in shards:
        for key in sorted(tensors, key=lambda k: k.encode()):
            if not (key.startswith("model.") or key.startswith("lm_")):
                continue
            a = tensors[key]
            layer, short = -1, None
            if key.startswith("lm_"):
                short = "logits"
            elif "layers." in key:
                rest = key[key.index("layers.") + 7:]
                layer, short = int(rest[:rest.index(".")]), rest[rest.index(".") + 1:]
            else:
                short = key[6:]
            out.append((key, layer, short, list(a.shape), off, off + a.size * 2, a))
            off += a.size * 2
    return out, off


def parse_meta(path):
    recs, cur = [], None
    for line in open(path):
        s = line.strip()
        if s.startswith("Tensor: "):
            cur = {"name": s[8:]}
            recs.append(cur)
        elif s.startswith("layer: "):
            cur["layer"] = int(s[7:])
        elif s.startswith("short_name: "):
            cur["short"] = s[12:]
        elif s.startswith("shape: "):
            cur["shape"] = [int(x) for x in s[s.index("[") + 1:s.index("]")].split()]
        elif s.startswith("offsets: "):
            b, e = s[s.index("[") + 1:s.index("]")].split(",")
            cur["off"] = (int(b), int(e))
    return recs


def test_two_shards_match_the_reference_layout(tmp_path):
    cfg = q.make_config("tiny", context=128)
    t = hf_tensors(cfg, 1)
    names = list(t)
    a = {k: t[k] for k in names if ".layers.1." not in k and k != "lm_head.weight"}
    b = {k: t[k] for k in names if ".layers.1." in k or k == "lm_head.weight"}
    s1, s2 = str(tmp_path / "model-00001-of-00002.safetensors"), str(tmp_path / "model-00002-of-00002.safetensors")
    # data order inside a shard deliberately differs from the key order
    write_safetensors(s1, a, order=list(reversed(list(a))))
    write_safetensors(s2, b)
    meta, wts = str(tmp_path / "meta_data.txt"), str(tmp_path / "weights.bin")
    total, n = q.convert_safetensors([s1, s2], meta, wts)
    exp, exp_total = expected_layout([a, b])
    assert (total, n) == (exp_total, len(exp)) and os.path.getsize(wts) == exp_total
    recs = parse_meta(meta)
    blob = np.fromfile(wts, dtype=np.uint16)
    assert len(recs) == len(exp)
    for r, (name, layer, short, shape, b0, b1, data) in zip(recs, exp):
        assert (r["name"], r["layer"], r["short"], r["shape"], r["off"]) == (name, layer, short, shape, (b0, b1))
        assert np.array_equal(blob[b0 // 2:b1 // 2], data.reshape(-1))
    # and the loader derives the model shape from it
    got, tot, cnt = q.inspect_checkpoint(meta)
    assert got.as_dict() == dict(cfg.as_dict(), context=got.context) and tot == exp_total and cnt == len(exp)


def test_tied_embeddings_and_refusals(tmp_path):
    cfg = q.make_config("tiny", context=128)
    t = hf_tensors(cfg, 2, with_lm_head=False)
    s = str(tmp_path / "model.safetensors")
    write_safetensors(s, t)
    meta, wts = str(tmp_path / "m.txt"), str(tmp_path / "w.bin")
    with pytest.raises(q.QieError):
        q.convert_safetensors([s], meta, wts)  # no lm_head and no permission to tie
    total, n = q.convert_safetensors([s], meta, wts, tie_lm_head=True)
    recs = parse_meta(meta)
    assert recs[-1]["name"] == "lm_head.weight" and recs[-1]["short"] == "logits"
    blob = np.fromfile(wts, dtype=np.uint16)
    b0, b1 = recs[-1]["off"]
    assert np.array_equal(blob[b0 // 2:b1 // 2], t["model.embed_tokens.weight"].reshape(-1)) and b1 == total
    # other dtypes are refused, not silently reinterpreted
    s16 = str(tmp_path / "f16.safetensors")
    write_safetensors(s16, hf_tensors(cfg, 3), dtype="F16")
    with pytest.raises(q.QieError):
        q.convert_safetensors([s16], meta, wts)
    with pytest.raises(q.QieError):
        q.convert_safetensors([str(tmp_path / "missing.safetensors")], meta, wts)
    bad = str(tmp_path / "bad.safetensors")
    open(bad, "wb").write(struct.pack("<Q", 20) + b'{"model.x": {"dtype"')
    with pytest.raises(q.QieError):
        q.convert_safetensors([bad], meta, wts)


def _synthetic_as_tensors(cfg, seed, tmp_path):
    meta0, w0 = str(tmp_path / "meta0.txt"), str(tmp_path / "w0.bin")
    q.write_synthetic_checkpoint(cfg, seed, meta0, w0)
    blob = np.fromfile(w0, dtype=np.uint16)
    t = {}
    for r in parse_meta(meta0):
        b0, b1 = r["off"]
        t[r["name"]] = blob[b0 // 2:b1 // 2].reshape(r["shape"]).copy()
    return meta0, w0, t


def test_round_trip_of_a_synthetic_checkpoint_is_byte_identical(tmp_path):
    """the synthetic checkpoint IS the reference layout of a one-shard model, so packing its tensors into one
    safetensors file (data in a different order) and converting must give back the same two files"""
    cfg = q.make_config("tiny", context=128)
    meta0, w0, t = _synthetic_as_tensors(cfg, 7, tmp_path)
    s = str(tmp_path / "model.safetensors")
    order = sorted(t, key=lambda k: (len(k), k))  # some other order
    write_safetensors(s, t, order=order)
    meta1, w1 = str(tmp_path / "meta1.txt"), str(tmp_path / "w1.bin")
    q.convert_safetensors([s], meta1, w1)
    assert open(meta1).read() == open(meta0).read()
    assert open(w1, "rb").read() == open(w0, "rb").read()
