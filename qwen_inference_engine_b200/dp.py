"""Data-parallel plumbing: one process per GPU, independent sequences, NO data-path
collective (the reference keeps all state per sequence -- batch_metadata +
ModelBuffers + page list -- so sequences shard naturally).  torch.distributed is used only
to gather token ids on the host side and to take max-over-ranks timings."""
import torch
import torch.distributed as dist


def shard(n_items, rank, world):
    """round-robin ownership: item i belongs to rank i % world (as the reference would hand
    sequences to engines one after another)"""
    return list(range(rank, n_items, world))


def gather_tokens(local_tokens, n_items, rank, world, device="cpu"):
    """local_tokens: LongTensor [n_local, steps] for items shard(n_items, rank, world).
    Returns on every rank the full [n_items, steps] tensor in item order."""
    steps = local_tokens.shape[1] if local_tokens.numel() else 0
    st = torch.tensor([steps], device=device)
    if world > 1:
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
    steps = int(st.item())
    per = (n_items + world - 1) // world
    pad = torch.full((per, steps), -1, dtype=torch.long, device=device)
    if local_tokens.numel():
        pad[:local_tokens.shape[0]] = local_tokens.to(device)
    if world == 1:
        return pad[:n_items]
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.full((n_items, steps), -1, dtype=torch.long, device=device)
    for r in range(world):
        idx = shard(n_items, r, world)
        out[idx] = parts[r][:len(idx)]
    return out


def max_over_ranks(value, world, device="cpu"):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
