#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200 (contract: see the task statement).

Metric (BASELINE.json): decode tokens/s.  Workload at N=1 = BASELINE.json configs[1]:
Qwen2.5-0.5B-arch bf16, batch 64 decode, 2048-token context, KV cache resident on one
B200.  A "step" is one decode step of the whole batch (64 tokens).  N>1 (torchrun, one
process per GPU) is data-parallel over independent sequences, no data-path collective:
every rank runs the same per-GPU workload (weak scaling), value = tokens of all ranks /
max-over-ranks device time.

  value     : kernel-only rate -- token ids and KV already resident in HBM, steps replayed
              from CUDA graphs, timed with CUDA events on the engine's stream.
  e2e       : the same step through the C ABI entry point a host driver calls
              (qie_decode_step: HOST token buffers in, HOST tokens out; the H2D/D2H copies
              and the stream sync are inside the timed region).
  roofline  : per-launch algorithmic bytes / live CUDA-event duration of the dominant
              kernel class, against MEASURED_PEAKS.json.
  cpu_baseline : the plain-C oracle (oracle/qie_oracle.c, "port") timed on the host cores
              on a bounded sample.  bench.py is one of the places allowed to run oracle/.
  --impl reference : the reference's own kernels (oracle/_ref, built from
              /root/reference/layers/src) replaying llm()'s launch order on the GPU --
              the reference has no CPU implementation; its only implementation is CUDA.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ARCH = "qwen2.5-0.5b"
BATCH = 64
CTX = 2048
METRIC = "decode tokens/s (Qwen2.5-0.5B-arch bf16, batch 64, ctx 2048)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "40"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def stop(self, t0=None, t1=None):
        """clocks of the samples taken inside [t0, t1] (host wall clock around the timed region).  nvidia-smi is
        started before the warm-up so that it is already streaming when the region begins; if the region is
        shorter than the sampling interval, the samples of the whole loaded window (warm-up + timed steps) are
        used and the JSON says so."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows, window = self.rows, "all samples"
        if t0 is not None and t1 is not None:
            inside = [r for r in self.rows if t0 <= r[0] <= t1]
            if inside:
                rows, window = inside, "timed region"
            else:
                rows, window = [r for r in self.rows if r[0] <= t1 + 0.1], "warm-up + timed region (region shorter than the sampling interval)"
        for r in (x[1:] for x in rows):
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for i, n in enumerate(names):
                    if r[5 + i].lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_baseline_port(cfg_name, sample_tokens=3, prompt_len=8, cores=None):
    """plain-C oracle (scalar port, row-parallel over all host cores) on a bounded sample:
    one sequence, short prompt, `sample_tokens` decode tokens of the same architecture."""
    import numpy as np

    from oracle.oracle import Oracle, OracleModel, OrcConfig
    cores = cores or os.cpu_count() or 1
    o = Oracle()
    o.set_threads(cores)
    d = tempfile.mkdtemp(prefix="qie_bench_")
    meta, wts = os.path.join(d, "meta_data.txt"), os.path.join(d, "weights.bin")
    o.synth_write(OrcConfig(**ARCH_DIMS), 1234, meta, wts)
    om = OracleModel(o, meta, wts)
    s = om.new_seq()
    tok = om.prefill(s, np.arange(1, prompt_len + 1, dtype=np.int32))
    t0 = time.perf_counter()
    for i in range(sample_tokens):
        tok = om.decode(s, tok, seed=1234 + 1 + i)
    dt = time.perf_counter() - t0
    om.close()
    try:
        os.remove(wts)
        os.remove(meta)
        os.rmdir(d)
    except OSError:
        pass
    return {"value": sample_tokens / dt, "unit": "tokens/s", "cores": cores, "kind": "port",
            "sample": f"oracle/qie_oracle.c: {sample_tokens} greedy decode tokens of ONE {cfg_name} sequence after a "
                      f"{prompt_len}-token prompt (context {prompt_len + 1}..{prompt_len + sample_tokens}), matmul rows over "
                      f"{cores} pthread(s); per-sequence rate (a batch of 64 is 64 such sequences one after another)"}


def _ref_model_on_device(arch_dims):
    """random-init checkpoint written by the ORACLE's generator (oracle/qie_oracle.c, the same counter hash as the
    product's), parsed by the oracle's loader and uploaded with torch: the reference arm never loads libqie_b200.so"""
    import ctypes as C

    import numpy as np
    import torch
    from oracle.oracle import Oracle, OracleModel, OrcConfig, RefLayerW, RefModelDesc
    o = Oracle()
    d = tempfile.mkdtemp(prefix="qie_ref_")
    meta, wts = os.path.join(d, "meta_data.txt"), os.path.join(d, "weights.bin")
    cfg = OrcConfig(**arch_dims)
    o.synth_write(cfg, 1234, meta, wts)
    om = OracleModel(o, meta, wts)
    keep = []

    def up(name, layer=-1):
        a = om.tensor(name, layer)
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).cuda()
        keep.append(t)
        return t.data_ptr()

    layers = (RefLayerW * cfg.layers)()
    names = dict(in_ln="input_layernorm.weight", q="self_attn.q_proj.weight", k="self_attn.k_proj.weight",
                 v="self_attn.v_proj.weight", o="self_attn.o_proj.weight", q_norm="self_attn.q_norm.weight",
                 k_norm="self_attn.k_norm.weight", post_ln="post_attention_layernorm.weight",
                 up="mlp.up_proj.weight", gate="mlp.gate_proj.weight", down="mlp.down_proj.weight")
    for l in range(cfg.layers):
        for f, sn in names.items():
            setattr(layers[l], f, up(sn, l))
    desc = RefModelDesc(cfg.hidden, cfg.inter, cfg.layers, cfg.n_q, cfg.n_kv, cfg.head_dim, cfg.vocab, cfg.context,
                        up("embed_tokens.weight"), up("norm.weight"), up("logits"), layers)
    desc._keep = (layers, keep)
    om.close()
    for f in (wts, meta):
        try:
            os.remove(f)
        except OSError:
            pass
    return desc


ARCH_DIMS = dict(hidden=896, inter=4864, layers=24, n_q=14, n_kv=2, head_dim=64, vocab=151936, context=32786)
PROMPT32 = [151643, 785, 50802, 1525, 3818] + list(range(100, 127))  # configs[0]: 32-token prompt


def run_reference(args):
    """the reference's own CUDA kernels (oracle/_ref) replaying llm()'s decode branch."""
    rank, world, local = dist_env()
    if rank != 0:
        return 0
    import numpy as np
    import torch
    from oracle.oracle import Ref, RefSeq
    torch.cuda.set_device(0)
    ref = Ref()
    desc = _ref_model_on_device(ARCH_DIMS)
    rs = RefSeq(ref, desc, page_size=4)  # the reference's main() uses page_size 4 (iengine.cu:334)
    rs.fake_context(CTX)                 # pages for CTX + 64 positions: the timed tokens never take the allocate-on-demand path
    tok = 785
    for _ in range(args.warmup):
        tok = rs.decode(tok)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tok = rs.decode(tok)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rs.close()
    val = args.steps / dt
    # configs[0] on the reference's CUDA path: 32-token prompt through its prefill branch, then greedy decode
    # (qwen_main.cu:74-247 / :250-404 with k = 1), llm()'s per-op device syncs kept
    extra = {}
    try:
        r1 = RefSeq(ref, desc, page_size=4)
        ids = np.asarray(PROMPT32, np.int32)
        torch.cuda.synchronize()
        tp0 = time.perf_counter()
        t1 = r1.prefill(ids)
        torch.cuda.synchronize()
        tp1 = time.perf_counter()
        n_dec = 64
        for _ in range(n_dec):
            t1 = r1.decode(t1)
        torch.cuda.synchronize()
        tp2 = time.perf_counter()
        r1.close()
        extra["config1_reference_cuda"] = {
            "workload": "configs[0]: 0.5B-arch, batch 1, 32-token prompt, greedy decode (64 of the 128 tokens timed), the "
                        "reference's kernels in llm()'s launch order with its per-op cudaDeviceSynchronize",
            "prefill_ms": 1e3 * (tp1 - tp0), "decode_tokens_per_s": n_dec / (tp2 - tp1), "us_per_token": 1e6 * (tp2 - tp1) / n_dec}
    except Exception as ex:
        extra["config1_reference_cuda_error"] = str(ex)
    sample = (f"reference CUDA kernels (sm_100a build of /root/reference/layers/src, launch order + syncs of llm() "
              f"decode) on ONE sequence at context {CTX}.. (KV pages pre-allocated, zero-filled), {args.steps} tokens; "
              "the reference has no batching: a batch of 64 is 64 such calls, so tokens/s is the per-call rate")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic (seeded random-init weights from the oracle's generator; zero KV cache of the stated length)",
            "config": {"workload": f"{ARCH} decode, batch {BATCH} (processed one sequence at a time), ctx {CTX}",
                       "note": "rank 0 only; the reference is single-GPU, single-sequence; libqie_b200.so is not loaded by this arm"},
            "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": 1, "kind": "reference", "sample": sample},
            "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "extra": extra}
    print(json.dumps(line), flush=True)
    return 0


def config4_leg(q, torch, np, rank, world, local, use_dist):
    """BASELINE configs[3]: the configs[0] workload x 256 sequences (distinct seeded 32-token prompts, greedy decode),
    256 / N sequences per GPU, full weight replica per GPU, no collective.  Total work is fixed -> strong scaling.
    More than 64 rows run as consecutive persistent launches of <= 64 rows."""
    n_seq = 256 // world
    cfg = q.make_config(ARCH)
    steps = 48
    eng = q.Engine(synthetic=ARCH, seed=1234, device=local, max_seqs=n_seq + 1, max_batch_tokens=max(64, n_seq), page_size=16,
                   kv_bytes=n_seq * 256 * q.kv_bytes_per_pos(cfg) + (64 << 20))
    rng = np.random.default_rng(77 + rank)
    seqs, first = [], []
    for i in range(n_seq):
        s_ = eng.new_sequence()
        seqs.append(s_)
        first.append(eng.prefill(s_, rng.integers(0, cfg.vocab, size=32, dtype=np.int32)))
    eng.decode_run(seqs, np.asarray(first, np.int32), 8)  # warm-up: eager shape + graph capture
    eng.sync()
    if use_dist:
        import torch.distributed as dist
        dist.barrier()
    ext = torch.cuda.ExternalStream(eng.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ext):
        e0.record()
    for _ in range(steps):
        eng.decode_step_device(seqs)
    with torch.cuda.stream(ext):
        e1.record()
    eng.sync()
    ms = e0.elapsed_time(e1)
    mega = bool(eng.uses_mega(n_seq, 32 + 8 + steps + 2))
    eng.close()
    if use_dist:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return {"workload": f"256 sequences, 32-token prompts, greedy decode at context 41..{40 + steps}, {n_seq} per GPU on {world} GPU(s)",
            "scaling": "strong", "sequences_per_gpu": n_seq, "persistent_kernel": mega, "steps": steps, "ms_per_step": ms / steps,
            "tokens_per_s": 256 * steps / (ms / 1e3)}


def tp_leg(q, torch, np, rank, world, local):
    """BASELINE configs[4] inside the driver-visible run: Qwen2.5-7B-arch sharded over the N ranks of this torchrun
    (heads / intermediate / vocabulary; o_proj and down_proj partial sums exchanged inside the persistent kernel over
    NVLink peer memory, NCCL all-reduce on the per-operator path), batch 1 and 16; on rank 0 the same model on one GPU
    for the speed-up; a parity run of a 3-layer model against tp_size 1; exchange latency from in-kernel timestamps."""
    import torch.distributed as dist
    out = {"tp_size": world}

    def connect(eng):
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(q.Engine.tp_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        eng.tp_connect(bytes(idt.cpu().numpy().tobytes()))

    def timed_decode(eng, batch, steps, prompts):
        seqs, first = [], []
        for b in range(batch):
            s_ = eng.new_sequence()
            seqs.append(s_)
            first.append(eng.prefill(s_, prompts[b]))
        cur = np.asarray(first, np.int32)
        for _ in range(6):
            cur = eng.decode_step(seqs, cur)
        st = torch.cuda.ExternalStream(eng.stream, device=local)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(st)
        toks = eng.decode_run(seqs, cur, steps)
        e1.record(st)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), np.asarray(toks), seqs

    cfg7 = q.make_config("qwen2.5-7b", context=4096)
    try:
        q.tp_plan(cfg7, rank, world)
    except Exception as ex:
        return {"tp_size": world, "unsupported": str(ex)}
    rng = np.random.default_rng(7)
    # ---- parity: 3 layers of the 7B shape, one forward, tp_size N vs tp_size 1 (rank 0)
    cfg3 = q.make_config("qwen2.5-7b", context=4096, layers=3)
    p3 = rng.integers(0, cfg3.vocab, size=24, dtype=np.int32)
    e3 = q.Engine(synthetic=cfg3, seed=1234, device=local, max_seqs=4, max_batch_tokens=64, kv_bytes=256 << 20, context=4096,
                  use_graph=False, tp_rank=rank, tp_size=world)
    connect(e3)
    s3 = e3.new_sequence()
    t3 = e3.prefill(s3, p3)
    d3 = [int(t3)]
    for _ in range(8):
        d3.append(int(e3.decode_step([s3], [d3[-1]])[0]))
    plan = q.tp_plan(cfg3, rank, world)
    lg_tp = e3.read_activation("logits", plan["vocab"]).copy()
    e3.close()
    if rank == 0:
        r3 = q.Engine(synthetic=cfg3, seed=1234, device=local, max_seqs=4, max_batch_tokens=64, kv_bytes=256 << 20, context=4096, use_graph=False)
        sr = r3.new_sequence()
        w3 = [int(r3.prefill(sr, p3))]
        for _ in range(8):
            w3.append(int(r3.decode_step([sr], [w3[-1]])[0]))
        full = r3.read_activation("logits", cfg3.vocab)[plan["vocab0"]:plan["vocab0"] + plan["vocab"]]
        a_ = (lg_tp.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
        b_ = (full.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
        out["parity_3layer"] = {"tokens_equal_tp1": d3 == w3, "logits_rel_l2_vs_tp1": float(np.linalg.norm(a_ - b_) / (np.linalg.norm(b_) + 1e-30)),
                                "logits_bit_identical": bool(np.array_equal(lg_tp, full))}
        r3.close()
    dist.barrier()
    # ---- speed: full 28-layer 7B shape
    for batch in (1, 16):
        prompts = rng.integers(0, cfg7.vocab, size=(batch, 32), dtype=np.int32)
        eng = q.Engine(synthetic=cfg7, seed=1234, device=local, max_seqs=batch + 1, max_batch_tokens=64, kv_bytes=1 << 30,
                       context=4096, tp_rank=rank, tp_size=world)
        connect(eng)
        steps = 48
        ms, toks, seqs = timed_decode(eng, batch, steps, prompts)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        rec = {"ms_per_step": ms / steps, "tokens_per_s": batch * steps / (ms / 1e3),
               "persistent_kernel": bool(eng.uses_mega(batch, 32 + 6 + steps + 2)),
               "exchanges_per_token": 2 * cfg7.layers + 1,
               "nvlink_bytes_sent_per_token_per_gpu": int((2 * cfg7.layers) * (world - 1) * batch * cfg7.hidden * 4)}
        if rec["persistent_kernel"]:
            try:  # exchange latency: the two cross-GPU barriers of a layer, from CTA 0's in-kernel timestamps
                eng.set_int("mega_prof", 1)
                eng.decode_step(seqs, toks[-1])
                ts, _ = eng.mega_prof()
                eng.set_int("mega_prof", 0)
                d_ = np.diff(ts.astype(np.int64))[:16 * cfg7.layers].reshape(cfg7.layers, 16)
                rec["exchange_wait_us"] = {"o_proj": float(d_[1:, 8].mean()) / 1e3, "down_proj": float(d_[1:, 15].mean()) / 1e3}
            except Exception as ex:
                rec["exchange_wait_error"] = str(ex)
        eng.close()
        dist.barrier()
        if rank == 0:
            e1_ = q.Engine(synthetic=cfg7, seed=1234, device=local, max_seqs=batch + 1, max_batch_tokens=64, kv_bytes=1 << 30, context=4096)
            ms1, _, _ = timed_decode(e1_, batch, steps, prompts)
            rec["one_gpu_tokens_per_s"] = batch * steps / (ms1 / 1e3)
            rec["speedup_vs_one_gpu"] = rec["tokens_per_s"] / rec["one_gpu_tokens_per_s"]
            e1_.close()
        dist.barrier()
        out[f"qwen2.5-7b_batch{batch}"] = rec
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--ctx", type=int, default=CTX)
    ap.add_argument("--numerics", default="reference_order", choices=["fast", "reference_order"],
                    help="reference_order (default): the bit-exact product path, one persistent kernel per decode "
                         "step; fast: the 1e-2-tolerance per-operator path (tcgen05 GEMM, flash-decoding)")
    ap.add_argument("--no-fast-extra", action="store_true", help="skip the fast-numerics side measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the batch-1 (configs[0]) side measurement")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        if args.steps > 4:
            args.steps = 4  # each reference token at ctx 2048 takes seconds (page-list walks)
        args.warmup = min(args.warmup, 1)
        return run_reference(args)

    import numpy as np
    import torch

    import qwen_inference_engine_b200 as q

    rank, world, local = dist_env()
    use_dist = world > 1
    torch.cuda.set_device(local)
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B, ctx = args.batch, args.ctx
    cfg = q.make_config(ARCH)
    kvpp = q.kv_bytes_per_pos(cfg)
    total_steps = 2 * (args.warmup + args.steps) + 8
    kv_need = B * (ctx + total_steps + 64) * kvpp + (64 << 20)
    eng = q.Engine(synthetic=ARCH, seed=1234, device=local, kv_bytes=kv_need, max_seqs=B + 1, max_batch_tokens=max(B, 64),
                   page_size=16, numerics=args.numerics)
    seqs = []
    for i in range(B):
        s = eng.new_sequence()
        eng.fill_synthetic(s, ctx, seed=1000 + 17 * rank + i)
        seqs.append(s)
    rng = np.random.default_rng(1234 + rank)
    tokens = rng.integers(0, cfg.vocab, size=B).astype(np.int32)

    ext = torch.cuda.ExternalStream(eng.stream)

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- kernel-only leg (graph replay, ids resident on the device) ----------
    sampler = ClockSampler(local)
    sampler.start()  # streaming before the warm-up, so samples exist from the first timed step on
    cur = eng.decode_step(seqs, tokens)  # stages ids/pos on the device, first graph shape seen
    for _ in range(args.warmup):
        eng.decode_step_device(seqs)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    ctx_leg0 = eng.seq_len(seqs[0])  # cached positions when the first timed step starts
    t_region0 = time.time()
    with torch.cuda.stream(ext):
        e0.record()
    for _ in range(args.steps):
        eng.decode_step_device(seqs)
        launches += eng.launch_count()
    with torch.cuda.stream(ext):
        e1.record()
    eng.sync()
    ms_kernel = e0.elapsed_time(e1)
    ctx_leg_mean = 0.5 * (ctx_leg0 + eng.seq_len(seqs[0]) - 1)  # mean cached positions per sequence over the timed steps
    clocks = sampler.stop(t_region0, time.time())
    barrier()

    # ---------------- e2e leg (host buffers through the C ABI) ----------------------------
    cur = eng.decode_step(seqs, tokens)
    for _ in range(args.warmup):
        cur = eng.decode_step(seqs, cur)
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(ext):
        e0.record()
    for _ in range(args.steps):
        cur = eng.decode_step(seqs, cur)
    with torch.cuda.stream(ext):
        e1.record()
    eng.sync()
    ms_e2e = max(e0.elapsed_time(e1), 1000.0 * (time.perf_counter() - t0))
    barrier()

    # ---------------- per-kernel-class timing of one step (roofline leg) -----------------
    mega = eng.uses_mega(B, eng.seq_len(seqs[0]) + 2)
    mega_phase_us = None
    if mega:
        # ONE kernel per step: its phases are timed by device timestamps taken inside the kernel
        eng.set_int("mega_prof", 1)
        cur = eng.decode_step(seqs, cur)
        ts, cyc = eng.mega_prof()
        eng.set_int("mega_prof", 0)
        ts = ts.astype(np.int64)
        L = cfg.layers
        d = np.diff(ts)
        per = d[:16 * L].reshape(L, 16).sum(axis=0) / 1e3  # us per step, summed over layers
        names = ["qkv.load", "qkv.rmsnorm", "qkv.gemm", "qkv.barrier", "attention.run", "attention.barrier", "o.load",
                 "o.gemm", "o.barrier", "gateup.load", "gateup.rmsnorm", "gateup.gemm", "gateup.barrier", "down.load",
                 "down.gemm", "down.barrier"]
        mega_phase_us = {n: round(float(v), 1) for n, v in zip(names, per)}
        for i, n in enumerate(["lm_head.load", "lm_head.rmsnorm", "lm_head.gemm", "lm_head.barrier", "sample"]):
            mega_phase_us[n] = round(float(d[16 * L + i]) / 1e3, 1)
        mega_phase_us["whole_kernel"] = round(float(ts[-1] - ts[0]) / 1e3, 1)
        mega_phase_us["sm_mhz_in_kernel"] = round(float(int(cyc[-1]) - int(cyc[0])) / max(1.0, float(ts[-1] - ts[0])) * 1e3)
        prof = {}
    else:
        prof = eng.decode_step_profile(seqs, cur)
    ctx_now = eng.seq_len(seqs[0])
    topk_extra = None
    if rank == 0 and world == 1 and not args.no_extra:
        # the reference's real sampling mode on the same batch: top-k 50, T 0.7, XORWOW (qwen_main.cu:381-388) -- the
        # persistent kernel leaves the logits, one radix-select + draw kernel follows (the reference does k = 50 scans)
        try:
            eng.set_sampling(topk=50, temperature_prefill=1.0, temperature_decode=0.7, seed=1234, add_step=True)
            cur = eng.decode_step(seqs, cur)
            for _ in range(3):
                eng.decode_step_device(seqs)
            eng.sync()
            with torch.cuda.stream(ext):
                e0.record()
            for _ in range(10):
                eng.decode_step_device(seqs)
            with torch.cuda.stream(ext):
                e1.record()
            eng.sync()
            ms_k = e0.elapsed_time(e1) / 10
            topk_extra = {"workload": "same batch / context, top-k 50 + temperature 0.7 + XORWOW draw per row (the reference's sampling mode)",
                          "ms_per_step": ms_k, "tokens_per_s": B / (ms_k / 1e3), "launches_per_step": int(eng.launch_count()),
                          "sampler_overhead_ms_vs_greedy": ms_k - ms_kernel / args.steps}
            eng.set_sampling(topk=1)
        except Exception as ex:
            topk_extra = {"error": str(ex)}

    if use_dist:
        t = torch.tensor([ms_kernel, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_kernel, ms_e2e = float(t[0]), float(t[1])

    tokens_total = B * args.steps * world
    value = tokens_total / (ms_kernel / 1000.0)
    e2e_val = tokens_total / (ms_e2e / 1000.0)

    # roofline of the dominant kernel class
    peak, peak_src = measured_peaks()
    H, I, hd = cfg.hidden, cfg.inter, cfg.head_dim
    Dq, Dkv = cfg.n_q * hd, cfg.n_kv * hd
    alg_bytes = {  # algorithmic bytes of ONE launch of each class (DESIGN.md "roofline arithmetic")
        "attention": B * ctx_now * 2 * Dkv * 2 + 2 * B * Dq * 2,
        "gemm_qkv": 2 * H * (Dq + 2 * Dkv) + B * 2 * (H + Dq + 2 * Dkv),
        "gemm_o": 2 * Dq * H + B * 2 * (Dq + 2 * H),
        "gemm_gateup": 2 * 2 * H * I + B * 2 * (H + I),
        "gemm_down": 2 * H * I + B * 2 * (I + 2 * H),
        "lm_head": 2 * cfg.vocab * H + B * 2 * (H + cfg.vocab),
    }
    step_bytes = int(q.weight_bytes(cfg) + B * ctx_leg_mean * kvpp + B * kvpp + B * 2 * H)  # SURVEY 8d, context of the TIMED leg
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if mega:
        # the step IS the kernel: algorithmic bytes of the step / CUDA-event duration of one launch
        dom = "decode_mega_kernel"
        dom_avg_s = ms_kernel / args.steps / 1000.0
        achieved = step_bytes / dom_avg_s / 1e9
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(f"{dom}_b{B}")
            except Exception:
                traffic = None
        att_us = mega_phase_us["attention.run"]
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "alg_bytes_per_launch": step_bytes, "avg_launch_us": dom_avg_s * 1e6, "share_of_step": 1.0,
                    "context_of_timed_leg": [int(ctx_leg0), int(ctx_leg0 + args.steps - 1)],
                    "phases_us_per_step": mega_phase_us,
                    "attention_phase": {"alg_bytes": alg_bytes["attention"] * cfg.layers,
                                        "achieved_GBs": alg_bytes["attention"] * cfg.layers / (att_us * 1e-6) / 1e9 if att_us else None},
                    "note": "one persistent cooperative kernel per decode step; phases timed by %globaltimer stamps of CTA 0 "
                            "inside a profiled launch (not part of the timed region)"}
    else:
        step_ms_eager = sum(v[0] for v in prof.values())
        dom = max(alg_bytes, key=lambda k: prof.get(k, (0, 0))[0])
        dom_ms, dom_n = prof[dom]
        dom_avg_s = (dom_ms / max(dom_n, 1)) / 1000.0
        achieved = alg_bytes[dom] / dom_avg_s / 1e9 if dom_avg_s > 0 else 0.0
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(dom)
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "alg_bytes_per_launch": alg_bytes[dom], "avg_launch_us": dom_avg_s * 1e6,
                    "share_of_step": dom_ms / step_ms_eager if step_ms_eager else None,
                    "step": {"alg_bytes": step_bytes, "achieved_GBs": step_bytes / (ms_kernel / args.steps / 1000.0) / 1e9,
                             "frac_of_peak": step_bytes / (ms_kernel / args.steps / 1000.0) / 1e9 / peak},
                    "by_kernel_ms": {k: round(v[0], 4) for k, v in prof.items()},
                    "by_kernel_launches": {k: v[1] for k, v in prof.items()}}

    line = {"metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_kernel / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic: seeded random-init weights generated on the device; KV cache pre-filled with seeded "
                    "random K/V (no real prefill); token ids fed back from the sampler",
            "config": {"workload": f"{ARCH} decode, batch {B} per GPU, ctx {ctx} at the first timed step (+1 per step)",
                       "parallelism": f"dp{world}", "page_size": 16, "sampling": "greedy (top-k 1, reference tie-break)",
                       "numerics": ("fast: tcgen05 GEMM + split-K, flash-decoding, parallel RMSNorm (1e-2 bf16 tolerance vs the reference)"
                                    if args.numerics == "fast" else
                                    "reference-order arithmetic (bit-identical to the reference's kernels per sequence)"
                                    + (", persistent decode kernel" if mega else ", per-operator launches")),
                       "l2": f"working set per step {step_bytes / 1e9:.2f} GB >> 126 MB L2 (inputs larger than L2, no flush)"},
            "e2e": {"value": e2e_val, "unit": "tokens/s", "h2d_bytes_per_step": int(4 * B), "d2h_bytes_per_step": int(4 * B),
                    "ms_per_step": ms_e2e / args.steps,
                    "api": "qie_decode_step (HOST int32 tokens in/out, pinned staging, one H2D + one D2H + stream sync per step)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if topk_extra:
        line.setdefault("extra", {})["topk50_batch64"] = topk_extra

    if rank == 0 and world == 1 and not args.no_fast_extra and args.numerics != "fast":
        # side measurement: the same workload through the fast-numerics per-operator path
        try:
            eng.close()
            engf = q.Engine(synthetic=ARCH, seed=1234, device=local, kv_bytes=kv_need, max_seqs=B + 1,
                            max_batch_tokens=max(B, 64), page_size=16, numerics="fast")
            sf = []
            for i in range(B):
                s_ = engf.new_sequence()
                engf.fill_synthetic(s_, ctx, seed=1000 + i)
                sf.append(s_)
            engf.decode_step(sf, tokens)
            for _ in range(args.warmup):
                engf.decode_step_device(sf)
            engf.sync()
            extf = torch.cuda.ExternalStream(engf.stream)
            with torch.cuda.stream(extf):
                e0.record()
            for _ in range(args.steps):
                engf.decode_step_device(sf)
            with torch.cuda.stream(extf):
                e1.record()
            engf.sync()
            msf = e0.elapsed_time(e1)
            line.setdefault("extra", {})["fast_numerics"] = {
                "workload": "same batch/context, per-operator launches: tcgen05 GEMM + flash-decoding (1e-2 tolerance, tokens not bit-exact)",
                "tokens_per_s": B * args.steps / (msf / 1000.0), "ms_per_step": msf / args.steps}
            engf.close()
            eng = q.Engine(synthetic=ARCH, seed=1234, device=local, kv_bytes=256 << 20, max_seqs=4, max_batch_tokens=64,
                           page_size=16, numerics=args.numerics)
        except Exception as ex:
            line.setdefault("extra", {})["fast_numerics_error"] = str(ex)
    if rank == 0 and world == 1 and not args.no_extra:
        # configs[0] side measurement: batch 1, 32-token prompt, 128 greedy tokens (the parity case)
        try:
            ids = np.array([151643, 785, 50802, 1525, 3818] + list(range(100, 127)), np.int32)
            s1 = eng.new_sequence()
            t_first = eng.prefill(s1, ids)
            eng.decode_run([s1], [t_first], 8)
            eng.sync()
            t0 = time.perf_counter()
            out = eng.decode_run([s1], [int(t_first)], 120)
            dt = time.perf_counter() - t0
            b1_bytes = q.weight_bytes(cfg) + kvpp * (eng.seq_len(s1) - 60)
            line.setdefault("extra", {})["batch1"] = {
                "workload": "configs[0]: batch 1, 32-token prompt, greedy decode (qie_decode_run, 120 tokens)",
                "persistent_kernel": bool(eng.uses_mega(1, 200)),
                "tokens_per_s": 120 / dt, "us_per_token": 1e6 * dt / 120,
                "frac_of_hbm_peak": (b1_bytes * 120 / dt / 1e9) / peak}
            eng.free_sequence(s1)
            if not args.no_fast_extra and args.numerics != "fast":
                # the same batch-1 case through the fast-numerics variant of the persistent kernel
                eng.close()
                eng = q.Engine(synthetic=ARCH, seed=1234, device=local, kv_bytes=256 << 20, max_seqs=4,
                               max_batch_tokens=64, page_size=16, numerics="fast")
                s1 = eng.new_sequence()
                t_first = eng.prefill(s1, ids)
                eng.decode_run([s1], [t_first], 8)
                eng.sync()
                t0 = time.perf_counter()
                eng.decode_run([s1], [int(t_first)], 120)
                dtf = time.perf_counter() - t0
                line["extra"]["batch1_fast_numerics"] = {
                    "workload": "configs[0] batch 1, fast numerics: the GEMV decode kernel (decode_gemv.cu: one persistent launch "
                                "per token, weight rows streamed by bulk copies, warp-per-row-group FHFMA dot products, "
                                "split-KV flash decoding, no grid barrier between the phases -- consumers poll per-layer "
                                "activation buffers; 1e-2 tolerance per layer, tokens not bit-exact)",
                    "persistent_kernel": bool(eng.uses_mega(1, 200)),
                    "tokens_per_s": 120 / dtf, "us_per_token": 1e6 * dtf / 120,
                    "frac_of_hbm_peak": (b1_bytes * 120 / dtf / 1e9) / peak}
        except Exception as ex:  # the side measurement must never break the contract line
            line.setdefault("extra", {})["batch1_error"] = str(ex)
    eng.close()
    if rank == 0 and world == 1 and not args.no_extra:
        # configs[2] side measurement: Qwen2.5-1.5B-arch prefill of 4096 tokens (tcgen05 GEMMs + tiled causal
        # attention, fast numerics), through qie_prefill with HOST ids in / HOST token out
        try:
            cfg15 = q.make_config("qwen2.5-1.5b", context=8192)
            T = 4096
            engp = q.Engine(synthetic=cfg15, seed=1234, device=local, max_seqs=2, max_batch_tokens=T, kv_bytes=2 << 30,
                            context=8192, numerics="fast")
            idsp = np.random.default_rng(5).integers(0, cfg15.vocab, size=T, dtype=np.int32)
            extp = torch.cuda.ExternalStream(engp.stream)
            tms = []
            for r_ in range(4):
                sp = engp.new_sequence()
                torch.cuda.synchronize()
                with torch.cuda.stream(extp):
                    e0.record()
                engp.prefill(sp, idsp)
                with torch.cuda.stream(extp):
                    e1.record()
                torch.cuda.synchronize()
                if r_ > 0:
                    tms.append(e0.elapsed_time(e1))
                engp.free_sequence(sp)
            msp = float(np.median(tms))
            H_, I_, L_, hd_ = cfg15.hidden, cfg15.inter, cfg15.layers, cfg15.head_dim
            Dq_, Dkv_ = cfg15.n_q * hd_, cfg15.n_kv * hd_
            flop = 2 * T * L_ * (H_ * Dq_ + 2 * H_ * Dkv_ + Dq_ * H_ + 3 * H_ * I_) + 2 * cfg15.vocab * H_ \
                + L_ * cfg15.n_q * 4 * hd_ * T * T / 2
            try:
                tpeak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
            except Exception:
                tpeak = 1377.7
            line.setdefault("extra", {})["prefill_1p5b_4096"] = {
                "workload": "configs[2]: qwen2.5-1.5b prefill of 4096 tokens in one forward (fast numerics: tcgen05/TMEM/TMA "
                            "GEMMs, FlashAttention-2 style causal attention on mma.sync), qie_prefill end to end",
                "ms": msp, "tokens_per_s": T / (msp / 1e3), "alg_tflop": flop / 1e12,
                "achieved_tflops": flop / 1e12 / (msp / 1e3), "frac_of_tensor_peak": flop / 1e12 / (msp / 1e3) / tpeak}
            engp.close()
        except Exception as ex:
            line.setdefault("extra", {})["prefill_error"] = str(ex)

    # ---------------- configs[3]: 256 sequences, 32-token prompts, sharded 256 / N per GPU (STRONG scaling) ----------
    if not args.no_extra:
        try:
            line.setdefault("extra", {})["config4_256seq"] = config4_leg(q, torch, np, rank, world, local, use_dist)
        except Exception as ex:
            line.setdefault("extra", {})["config4_error"] = str(ex)
    # ---------------- configs[4]: Qwen2.5-7B-arch tensor parallel over the N GPUs of this run ---------------------
    if use_dist and not args.no_extra:
        try:
            line.setdefault("extra", {})["tp"] = tp_leg(q, torch, np, rank, world, local)
        except Exception as ex:
            line.setdefault("extra", {})["tp_error"] = str(ex)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline_port(ARCH)
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {ex}"}
        try:  # the north star's "scalar C++ CPU forward": the same port on ONE thread
            line.setdefault("extra", {})["cpu_scalar_1thread"] = cpu_baseline_port(ARCH, sample_tokens=2, prompt_len=4, cores=1)
        except Exception as ex:
            line.setdefault("extra", {})["cpu_scalar_1thread_error"] = str(ex)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
