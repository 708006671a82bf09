"""One-process-per-GPU plumbing for the sharded MSMs: shard ranges, the single small
all_gather of per-rank partial commitments, and the fold order.  torch.distributed is plumbing only
(NCCL on GPUs, gloo in the CPU tests); the arithmetic is in the C library."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .bindings import Groth16Partials

PARTIALS_WORDS = C.sizeof(Groth16Partials) // 4  # 144 x u32 = 576 B: A, B1, C, H (G1 projective) + B2 (G2)


def shard_range(n: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of an n-element base-point section owned by `rank` - must match
    csrc/groth16.cu `shard()`."""
    return n * rank // world, n * (rank + 1) // world


def partials_to_tensor(parts: Groth16Partials, device):
    import torch
    return torch.frombuffer(bytearray(bytes(parts)), dtype=torch.int32).to(device)


def all_gather_partials(parts: Groth16Partials, device="cpu"):
    """Every rank contributes 576 B; returns the list of all ranks' partials in rank order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    mine = partials_to_tensor(parts, device)
    out = [torch.zeros(PARTIALS_WORDS, dtype=torch.int32, device=device) for _ in range(world)]
    dist.all_gather(out, mine)
    host = torch.stack(out).cpu().numpy()
    return [Groth16Partials.from_buffer_copy(np.ascontiguousarray(host[i]).tobytes()) for i in range(world)]
