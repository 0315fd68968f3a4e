"""Tensor-parallel probe (BASELINE configs[4]): run under torchrun, one rank per GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/tp_probe.py --arch qwen2.5-7b --steps 64 [--check]

Every rank builds the same synthetic checkpoint on its GPU, takes its shard by pointer
arithmetic (qie_tp_plan) and decodes greedily; o_proj / down_proj partial sums are
all-reduced with NCCL over NVLink inside the engine.  --check also builds a tp_size 1 engine
on rank 0 and compares tokens and final logits (tolerance: the bf16 all-reduce rounds the
partial sums).  Prints ONE JSON line on rank 0; times are CUDA-event device times,
max over ranks."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwen_inference_engine_b200 as q  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="qwen2.5-7b")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--prompt", type=int, default=32)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--graph", type=int, default=1)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = q.make_config(a.arch, context=4096)
    eng = q.Engine(synthetic=cfg, seed=1234, device=local, max_seqs=a.batch + 1, max_batch_tokens=max(a.batch, 64),
                   kv_bytes=2 << 30, context=4096, use_graph=bool(a.graph), tp_rank=rank, tp_size=world)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(q.Engine.tp_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        eng.tp_connect(bytes(idt.cpu().numpy().tobytes()))
    rng = np.random.default_rng(7)
    prompts = rng.integers(0, cfg.vocab, size=(a.batch, a.prompt), dtype=np.int32)
    seqs, first = [], []
    for b in range(a.batch):
        s = eng.new_sequence()
        seqs.append(s)
        first.append(eng.prefill(s, prompts[b]))
    toks = [np.array(first, dtype=np.int32)]
    cur = toks[0]
    for _ in range(a.warmup):
        cur = eng.decode_step(seqs, cur)
        toks.append(cur)
    st = torch.cuda.ExternalStream(eng.stream, device=local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record(st)
    out = eng.decode_run(seqs, cur, a.steps)  # tokens fed back on the device, one host sync at the end
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    out = np.asarray(out).reshape(a.steps, a.batch)
    all_toks = np.concatenate([np.stack(toks), out], axis=0)  # [1 + warmup + steps, batch]
    # every rank must hold the same tokens
    tt = torch.from_numpy(all_toks.astype(np.int64)).cuda()
    same = True
    if world > 1:
        t0 = tt.clone()
        dist.broadcast(t0, 0)
        ok = torch.tensor([int(torch.equal(t0, tt))], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item())
    res = {"arch": a.arch, "tp": world, "batch": a.batch, "steps": a.steps, "ms_per_step": ms / a.steps,
           "tokens_per_s": a.batch * a.steps / (ms / 1e3), "ranks_agree": same,
           "persistent_kernel": bool(eng.uses_mega(a.batch, a.prompt + a.warmup + a.steps + 2)),
           "launches_last_call": int(eng.launch_count())}
    wb = q.weight_bytes(cfg)
    res["weight_bytes_per_gpu"] = wb // world
    res["hbm_GBs_per_gpu"] = wb / world / (ms / a.steps / 1e3) / 1e9
    prefill_logits = None
    if a.check:
        # collective: one more prefill on every rank so that rank 0 can compare the logits of a
        # single forward (2 * layers all-reduces, no KV history) with the tp_size 1 engine
        plan0 = q.tp_plan(cfg, rank, world)
        last_logits = eng.read_activation("logits", a.batch * plan0["vocab"]).copy()
        s2 = eng.new_sequence()
        eng.prefill(s2, prompts[0])
        prefill_logits = eng.read_activation("logits", plan0["vocab"]).copy()
    if a.check and rank == 0:
        ref = q.Engine(synthetic=cfg, seed=1234, device=local, max_seqs=max(a.batch, 1),
                       max_batch_tokens=max(a.batch, 64), kv_bytes=2 << 30, context=4096, use_graph=False)
        ref.set_int("mega", 0)
        agree, total = 0, 0
        plan = q.tp_plan(cfg, rank, world)
        v0, vl = plan["vocab0"], plan["vocab"]
        mine = (last_logits.astype(np.uint32).reshape(a.batch, vl) << 16).view(np.float32)
        pl = (prefill_logits.astype(np.uint32) << 16).view(np.float32)
        rel = []
        for b in range(a.batch):
            s = ref.new_sequence()
            t1 = ref.prefill(s, prompts[b])
            if b == 0:
                full = ref.read_activation("logits", cfg.vocab).astype(np.uint32)
                full = (full << 16).view(np.float32)[v0:v0 + vl]
                res["check_prefill_logits_rel_l2"] = float(np.linalg.norm(full - pl) / np.linalg.norm(full))
            n_cmp = all_toks.shape[0]
            # teacher-forced on the TP engine's tokens: compare the next-token choice step by step
            seq_ok = int(t1 == all_toks[0, b])
            for i in range(1, n_cmp):
                nxt = ref.decode_step([s], [int(all_toks[i - 1, b])])
                seq_ok += int(nxt[0] == all_toks[i, b])
            agree += seq_ok
            total += n_cmp
            full = ref.read_activation("logits", cfg.vocab).astype(np.uint32)
            full = (full << 16).view(np.float32)[v0:v0 + vl]
            rel.append(float(np.linalg.norm(full - mine[b]) / np.linalg.norm(full)))
        res["check_tokens_agree"] = [agree, total]
        res["check_logits_rel_l2_max"] = max(rel)
        ref.close()
    if rank == 0:
        print(json.dumps(res), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
