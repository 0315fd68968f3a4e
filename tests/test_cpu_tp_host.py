"""not-gpu: tensor-parallel host logic (BASELINE configs[4]).
* qie_tp_plan: the shards of all ranks tile heads / intermediate / vocabulary exactly, vocabulary
  shards start at multiples of 256 (the sampler's tie-break key is idx mod 256,
  /root/reference/layers/src/logit_decode.cu:15-33), indivisible shapes are refused.
* world_size 2 over gloo: column-sharded gate/up + row-sharded down_proj with the plan's
  offsets, partial sums all-reduced, equals the unsharded product (fp64, exact up to sum order)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qwen_inference_engine_b200 as q  # noqa: E402

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


@pytest.mark.parametrize("arch,tp", [("qwen2.5-7b", 2), ("qwen2.5-7b", 4), ("qwen2.5-0.5b", 2), ("qwen2.5-1.5b", 2),
                                     ("small", 2)])
def test_plan_tiles_the_model(arch, tp):
    cfg = q.make_config(arch)
    plans = [q.tp_plan(cfg, r, tp) for r in range(tp)]
    assert sum(p["n_q"] for p in plans) == cfg.n_q and sum(p["n_kv"] for p in plans) == cfg.n_kv
    assert sum(p["inter"] for p in plans) == cfg.inter and sum(p["vocab"] for p in plans) == cfg.vocab
    q0 = kv0 = i0 = v0 = 0
    for p in plans:
        assert (p["q_row0"], p["kv_row0"], p["inter0"], p["vocab0"]) == (q0, kv0, i0, v0)
        assert p["vocab0"] % 256 == 0 and p["n_q"] % p["n_kv"] == 0
        # GQA groups stay whole: q head h uses kv head h // (n_q / n_kv) on every rank
        assert p["q_row0"] // cfg.head_dim // (cfg.n_q // cfg.n_kv) == p["kv_row0"] // cfg.head_dim
        q0 += p["n_q"] * cfg.head_dim
        kv0 += p["n_kv"] * cfg.head_dim
        i0 += p["inter"]
        v0 += p["vocab"]


def test_plan_refuses_indivisible_shapes():
    cfg = q.make_config("qwen2.5-7b")
    with pytest.raises(q.QieError):
        q.tp_plan(cfg, 2, 2)
    with pytest.raises(q.QieError):
        q.tp_plan(cfg, 0, 3)  # neither the kv heads nor the ranks divide
    with pytest.raises(q.QieError):
        q.tp_plan(q.make_config("qwen2.5-0.5b"), 0, 3)


@pytest.mark.parametrize("arch,tp", [("qwen2.5-7b", 8), ("qwen2.5-0.5b", 4), ("small128", 2), ("qwen2.5-1.5b", 4)])
def test_plan_with_more_ranks_than_kv_heads(arch, tp):
    """BASELINE configs[4] names 2/4/8 GPUs; Qwen2.5-7B has 28 q / 4 kv heads.  With more ranks than kv heads, `rep`
    ranks share one kv head (each caches it) and split its query group: every q head is owned exactly once, a rank's
    q heads all belong to its kv head, o_proj columns = the owned q rows, MLP and vocabulary tile as always."""
    cfg = q.make_config(arch)
    plans = [q.tp_plan(cfg, r, tp) for r in range(tp)]
    group, rep, hd = cfg.n_q // cfg.n_kv, tp // cfg.n_kv, cfg.head_dim
    owned = []
    for r, p in enumerate(plans):
        assert p["n_kv"] == 1 and p["kv_row0"] == (r // rep) * hd
        heads = list(range(p["q_row0"] // hd, p["q_row0"] // hd + p["n_q"]))
        assert heads and all(h // group == r // rep for h in heads)  # GQA: q head h reads kv head h // group
        owned += heads
        assert p["vocab0"] % 256 == 0 and p["inter"] % 8 == 0
    assert sorted(owned) == list(range(cfg.n_q))
    sizes = [p["n_q"] for p in plans]
    assert max(sizes) - min(sizes) <= 1  # 7 heads over 2 ranks: 4 + 3
    assert sum(p["inter"] for p in plans) == cfg.inter and sum(p["vocab"] for p in plans) == cfg.vocab
    assert [p["inter0"] for p in plans] == list(np.cumsum([0] + [p["inter"] for p in plans[:-1]]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mlp_inputs(cfg):
    rng = np.random.default_rng(3)
    H, I = cfg.hidden, cfg.inter
    x = rng.standard_normal((3, H))
    return x, rng.standard_normal((I, H)) * 0.05, rng.standard_normal((I, H)) * 0.05, rng.standard_normal((H, I)) * 0.05


def _silu(v):
    return v / (1.0 + np.exp(-v))


def _worker(rank, world, port, q_out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import qwen_inference_engine_b200 as qq
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    cfg = qq.make_config("small")
    x, gate, up, down = _mlp_inputs(cfg)
    p = qq.tp_plan(cfg, rank, world)
    r0, r1 = p["inter0"], p["inter0"] + p["inter"]
    h = _silu(x @ gate[r0:r1].T) * (x @ up[r0:r1].T)       # column-parallel: local intermediate slice
    part = torch.from_numpy(h @ down[:, r0:r1].T)          # row-parallel: partial sums over the local columns
    dist.all_reduce(part)                                  # the all-reduce after down_proj
    if rank == 0:
        q_out.put(part.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_mlp_allreduce_world2_gloo():
    cfg = q.make_config("small")
    x, gate, up, down = _mlp_inputs(cfg)
    full = (_silu(x @ gate.T) * (x @ up.T)) @ down.T
    ctx = mp.get_context("spawn")
    qo = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, qo)) for r in range(2)]
    for p in procs:
        p.start()
    got = qo.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_allclose(got, full, rtol=1e-10, atol=1e-10)
