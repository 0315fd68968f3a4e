"""helpers shared by the tests (bf16 <-> numpy, seeded inputs)"""
import numpy as np


def bf16_to_f32(a):
    a = np.ascontiguousarray(a, dtype=np.uint16)
    return (a.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(a):
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = u + 0x7FFF + ((u >> 16) & 1)
    return (u >> 16).astype(np.uint16)


def rand_bf16(rng, shape, scale=1.0):
    return f32_to_bf16((rng.standard_normal(shape) * scale).astype(np.float32))


def to_dev(u16):
    import torch
    return torch.from_numpy(np.ascontiguousarray(u16).view(np.int16)).cuda().view(torch.bfloat16)


def to_host(t):
    import torch
    return t.detach().contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


def rel_err(a_u16, b_u16):
    """max |a-b| / max|b| over bf16 arrays (the north-star 'relative error in bf16')."""
    a, b = bf16_to_f32(a_u16).astype(np.float64), bf16_to_f32(b_u16).astype(np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def rel_l2(a_u16, b_u16):
    """||a-b||_2 / ||b||_2 over bf16 arrays: the tensor-level 'relative error in bf16'."""
    a, b = bf16_to_f32(a_u16).astype(np.float64), bf16_to_f32(b_u16).astype(np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def ulp_diff(a_u16, b_u16):
    """max distance in bf16 ulps (monotone integer mapping of the bit patterns)."""
    def key(u):
        u = u.astype(np.int32)
        return np.where(u & 0x8000, 0x8000 - (u & 0x7FFF), 0x8000 + u)
    return int(np.max(np.abs(key(np.asarray(a_u16)) - key(np.asarray(b_u16)))))


PROMPT32 = [151643, 785, 50802, 1525, 3818]  # start of the reference's example prompts (iengine.cu:325, embedded_matrix.cu:40)


def prompt_ids(n, vocab, seed=7):
    rng = np.random.default_rng(seed)
    ids = [t % vocab for t in PROMPT32][:n]
    while len(ids) < n:
        ids.append(int(rng.integers(0, vocab)))
    return np.asarray(ids, np.int32)
