"""not-gpu: the N>1 data-parallel path on CPU with the gloo backend (world_size 2).
Sequences are sharded round-robin with no data-path collective; each rank decodes its
shard (here with the CPU oracle standing in for the GPU engine); token ids are gathered
on the host and must equal the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _decode_all(meta, wts, prompts, steps):
    from oracle.oracle import Oracle, OracleModel
    om = OracleModel(Oracle(), meta, wts, context=128)
    out = [om.generate(p, steps) for p in prompts]
    om.close()
    return torch.tensor(out, dtype=torch.long).reshape(len(prompts), steps)


def _worker(rank, world, port, meta, wts, n_seq, steps, q_out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from qwen_inference_engine_b200 import dp
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    prompts = [[1 + i, 2 + i, 3 + i] for i in range(n_seq)]
    mine = dp.shard(n_seq, rank, world)
    local = _decode_all(meta, wts, [prompts[i] for i in mine], steps)
    full = dp.gather_tokens(local, n_seq, rank, world)
    t = dp.max_over_ranks(1.0 + rank, world)
    if rank == 0:
        q_out.put((full.numpy(), t))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_world2_matches_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    import qwen_inference_engine_b200 as q
    from qwen_inference_engine_b200 import dp
    cfg = q.make_config("tiny", context=128)
    meta, wts = str(tmp_path / "meta_data.txt"), str(tmp_path / "weights.bin")
    q.write_synthetic_checkpoint(cfg, 5, meta, wts)
    n_seq, steps = 5, 6
    assert dp.shard(5, 0, 2) == [0, 2, 4] and dp.shard(5, 1, 2) == [1, 3]
    single = _decode_all(meta, wts, [[1 + i, 2 + i, 3 + i] for i in range(n_seq)], steps).numpy()
    ctx = mp.get_context("spawn")
    qo = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, meta, wts, n_seq, steps, qo)) for r in range(2)]
    for p in procs:
        p.start()
    full, tmax = qo.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(full, single)
    assert tmax == 2.0  # max over ranks
