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


def shard_plan(lib, n_vars: int, domain_size: int, rank: int, world: int, mode: int = -1, skew: float = 0.0):
    """[(lo, hi)] x 5 for the sections (H, A, B1, C, B2) that `rank` holds - the library's own planner
    (b200_shard_plan; mode -1 = its default: the cost-weighted line cut for world > 1, see include/icicle_b200.h)."""
    lo, hi = (C.c_uint32 * 5)(), (C.c_uint32 * 5)()
    rc = lib.dll.b200_shard_plan(C.c_uint32(n_vars), C.c_uint32(domain_size), C.c_int(rank), C.c_int(world), C.c_int(mode),
                                 C.c_double(skew), lo, hi)
    if rc != 0:
        raise ValueError(f"b200_shard_plan failed: {rc}")
    return [(int(lo[k]), int(hi[k])) for k in range(5)]


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


# ---- quotient split (include/icicle_b200.h: b200_groth16_commit_begin / commit_end) ----------------------------
def poly_owner(j: int, world: int) -> int:
    """Rank that transforms polynomial j of (0: B.w, 1: A.w, 2: A.w*B.w); owned sets are contiguous."""
    if world >= 3:
        return j
    if world == 2:
        return 0 if j < 2 else 1
    return 0


def owned_polys(rank: int, world: int):
    mine = [j for j in range(3) if poly_owner(j, world) == rank]
    return (mine[0], len(mine)) if mine else (0, 0)


class QuotientExchange:
    """Buffers + the one collective step of the split: every owner scatters the H-shard slices of its transformed
    polynomial(s); afterwards each rank holds its slice of all three."""

    def __init__(self, cache, device):
        import torch
        import torch.distributed as dist
        self.cache, self.dist, self.torch = cache, dist, torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.N = cache.domain_size
        self.first, self.count = owned_polys(self.rank, self.world)
        self.mine = torch.empty((max(self.count, 1), self.N, 8), dtype=torch.int32, device=device)
        import os
        skew = float(os.environ.get("B200_SHARD_SKEW", "0") or 0)
        self.ranges = [shard_plan(cache.lib, cache.n_vars, self.N, r, self.world, -1, skew)[0] for r in range(self.world)]
        lo, hi = self.ranges[self.rank]
        assert (lo, hi) == cache.h_range()
        self.slices = torch.empty((3, max(hi - lo, 1), 8), dtype=torch.int32, device=device)

    def commit(self, witness, n_witness=None):
        self.cache.commit_begin(witness, self.first, self.count, self.mine.data_ptr(), n_witness=n_witness)
        for j in range(3):
            owner = poly_owner(j, self.world)
            sl = None
            if self.rank == owner:
                src = self.mine[j - self.first]
                sl = [src[lo:hi] for lo, hi in self.ranges]
            # (ranks differ in slice length under the line plan: point-to-point instead of scatter)
            reqs = []
            if self.rank == owner:
                for q, (lo, hi) in enumerate(self.ranges):
                    if hi == lo:
                        continue
                    if q == self.rank:
                        self.slices[j][:hi - lo].copy_(sl[q])
                    else:
                        reqs.append(self.dist.isend(sl[q].contiguous(), dst=q))
            else:
                lo, hi = self.ranges[self.rank]
                if hi > lo:
                    reqs.append(self.dist.irecv(self.slices[j][:hi - lo], src=owner))
            for rq in reqs:
                rq.wait()
        self.torch.cuda.current_stream().synchronize()
        # d_vec order: 0 = B.w', 1 = A.w', 2 = product'; commit_end takes (a, b, c) = (A', B', product')
        return self.cache.commit_end(self.slices[1].data_ptr(), self.slices[0].data_ptr(), self.slices[2].data_ptr())


# ---- the library's own communicator (include/icicle_b200.h: b200_comm_*) ------------------------------------------
class LibComm:
    """NCCL communicator owned by the C library: rank 0 draws the rendezvous token, `bcast` (any callable that returns
    rank 0's bytes on every rank - torch.distributed here, an MPI / TCP broadcast in another host) distributes it."""

    def __init__(self, lib, rank, world, bcast):
        self.lib, self.rank, self.world = lib, rank, world
        token = (C.c_uint8 * 128)()
        status = 0
        if rank == 0:
            status = lib.dll.b200_comm_unique_id(token)
        # the status travels with the token so that a failure on rank 0 (no libnccl.so.2) raises on every rank, not a hang
        raw = bcast(bytes([status & 0xFF]) + bytes(token))
        if raw[0] != 0:
            raise RuntimeError(f"b200_comm_unique_id failed on rank 0: {raw[0]} (no libnccl.so.2?)")
        token = (C.c_uint8 * 128).from_buffer_copy(raw[1:129])
        h = C.c_void_p()
        rc = lib.dll.b200_comm_create(token, C.c_int(rank), C.c_int(world), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"b200_comm_create failed: {rc}")
        self.handle = h

    @classmethod
    def from_torch(cls, lib):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()

        def bcast(raw):
            t = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, src=0)
            return bytes(t.cpu().numpy().tobytes())

        return cls(lib, rank, world, bcast)

    def msm(self, scalars, points, local_size, cfg, g2=False):
        """Sharded standalone MSM (b200_msm_sharded): this rank's slice in, the total (host, projective) out."""
        out = np.zeros(48 if g2 else 24, dtype=np.uint32)
        sp = C.c_void_p(scalars) if isinstance(scalars, int) else scalars.ctypes.data_as(C.c_void_p)
        pp = C.c_void_p(points) if isinstance(points, int) else points.ctypes.data_as(C.c_void_p)
        rc = self.lib.dll.b200_msm_sharded(self.handle, sp, pp, C.c_int(local_size), C.byref(cfg), C.c_int(int(g2)), out.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError(f"b200_msm_sharded failed: {rc}")
        return out

    def close(self):
        if self.handle:
            self.lib.dll.b200_comm_destroy(self.handle)
            self.handle = None
