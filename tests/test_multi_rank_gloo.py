"""N > 1 host logic on CPU: two gloo ranks each commit to their contiguous shard (here with the reference
CPU library standing in for the GPU kernels), all_gather the 576 B partials, rank 0 folds them with the
product library's host helpers; the fold equals the unsharded commitments."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conftest  # noqa: F401
    import icicle_snark_b200 as pkg
    import torch.distributed as dist
    from oracle import ref_cpu
    from util import rand_scalars
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ref, lib = ref_cpu.ref(), pkg.lib()
        rng = np.random.default_rng(99)  # same inputs on every rank
        n = 257
        sc, _ = rand_scalars(rng, n)
        p1 = ref.generate_affine_points(4)[np.arange(n) % 4].copy()
        p1 = np.ascontiguousarray(ref.convert_montgomery(ref.convert_montgomery(p1, True, kind="affine"), False, kind="affine"))
        # deterministic across ranks: derive points from the scalars instead of the library's RNG
        gen = ref.generator()
        base = [ref.to_affine(ref.mul_scalar(gen, sc[i])) for i in range(8)]
        p1 = np.array([base[i % 8] for i in range(n)], dtype=np.uint32)
        g2gen = ref.generator(g2=True)
        base2 = [ref.to_affine(ref.mul_scalar(g2gen, sc[i], g2=True), g2=True) for i in range(4)]
        p2 = np.array([base2[i % 4] for i in range(n)], dtype=np.uint32)
        lo, hi = pkg.multi_gpu.shard_range(n, rank, world)
        parts = pkg.bindings.Groth16Partials()
        g1 = ref.msm(np.ascontiguousarray(sc[lo:hi]), np.ascontiguousarray(p1[lo:hi]))[0]
        g2 = ref.msm(np.ascontiguousarray(sc[lo:hi]), np.ascontiguousarray(p2[lo:hi]), g2=True)[0]
        for name in ("a", "b1", "c", "h"):
            getattr(parts, name)[:] = [int(x) for x in g1]
        parts.b2[:] = [int(x) for x in g2]
        allp = pkg.multi_gpu.all_gather_partials(parts, "cpu")
        assert len(allp) == world
        if rank == 0:
            acc1 = np.array(list(allp[0].a), dtype=np.uint32)
            acc2 = np.array(list(allp[0].b2), dtype=np.uint32)
            for p in allp[1:]:
                acc1 = lib.ecadd(acc1, np.array(list(p.a), dtype=np.uint32))
                acc2 = lib.ecadd(acc2, np.array(list(p.b2), dtype=np.uint32), g2=True)
            full1, full2 = ref.msm(sc, p1)[0], ref.msm(sc, p2, g2=True)[0]
            ok = bool(np.array_equal(ref.to_affine(acc1), ref.to_affine(full1)) and
                      np.array_equal(ref.to_affine(acc2, g2=True), ref.to_affine(full2, g2=True)))
            q.put(ok)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_and_fold():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_ranges_partition():
    import icicle_snark_b200 as pkg
    for n in (0, 1, 7, 3_200_002):
        for world in (1, 2, 3, 8):
            r = [pkg.multi_gpu.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))


def test_skewed_witness_shards_partition_and_balance(lib):
    """b200_shard_range: the split the library uses for the signal-indexed sections when the quotient chain is split
    (B200_SHARD_SKEW). Always a partition of [0, n); skew 0 equals the plain split; with skew s the polynomial owners'
    shares are smaller by exactly s of the total per owned polynomial."""
    import ctypes as C
    import icicle_snark_b200 as pkg

    def rng(n, rank, world, skew):
        lo, hi = C.c_uint32(), C.c_uint32()
        assert lib.dll.b200_shard_range(C.c_uint32(n), rank, world, C.c_double(skew), C.byref(lo), C.byref(hi)) == 0
        return lo.value, hi.value

    for n in (0, 1, 7, 100, 3_200_002):
        for world in (1, 2, 3, 4, 8):
            for skew in (0.0, 0.049, 0.1, 0.2, 5.0):
                r = [rng(n, k, world, skew) for k in range(world)]
                assert r[0][0] == 0 and r[-1][1] == n
                assert all(r[k][1] == r[k + 1][0] and r[k][0] <= r[k][1] for k in range(world - 1))
                if skew == 0.0:
                    assert r == [pkg.multi_gpu.shard_range(n, k, world) for k in range(world)]
    n = 3_200_002
    for world in (2, 4, 8):
        sizes = [hi - lo for lo, hi in (rng(n, k, world, 0.049) for k in range(world))]
        owned = [pkg.multi_gpu.owned_polys(k, world)[1] for k in range(world)]
        for k in range(world):
            want = n * (1.0 / world + (3.0 / world - owned[k]) * 0.049)
            assert abs(sizes[k] - want) <= 2, (world, k, sizes, want)
        # model: time = polys * skew + share is the same on every rank
        t = [owned[k] * 0.049 + sizes[k] / n for k in range(world)]
        assert max(t) - min(t) < 1e-5
    lo, hi = C.c_uint32(), C.c_uint32()
    assert lib.dll.b200_shard_range(C.c_uint32(10), 3, 3, C.c_double(0.0), C.byref(lo), C.byref(hi)) == 11
    assert lib.dll.b200_shard_range(C.c_uint32(10), 0, 1, C.c_double(0.0), None, C.byref(hi)) == 3


def test_line_plan_partitions_and_balances(lib, monkeypatch):
    """b200_shard_plan mode 1: the five sections (H, A, B1, C, B2) laid end to end by cost and cut into `world` pieces.
    Every section is partitioned exactly; the weighted load (points x cost + the owners' polynomial transforms) is equal
    across ranks up to the snapping of cuts to section edges; a rank holds at most two partial sections."""
    import icicle_snark_b200 as pkg
    for name in ("B200_SHARD_PLAN", "B200_PLAN_W2", "B200_PLAN_WH", "B200_PLAN_WNTT"):
        monkeypatch.delenv(name, raising=False)
    w2, wh, wntt = 3.0, 1.1, 0.24
    for n_vars, N in ((8, 8), (102, 128), (3_200_002, 1 << 22), (100_002, 1 << 17)):
        sizes = (N, n_vars, n_vars, n_vars, n_vars)
        for world in (1, 2, 3, 4, 5, 8, 16):
            plans = [pkg.multi_gpu.shard_plan(lib, n_vars, N, r, world, 1) for r in range(world)]
            for k in range(5):
                assert plans[0][k][0] == 0 and plans[-1][k][1] == sizes[k]
                assert all(plans[r][k][1] == plans[r + 1][k][0] and plans[r][k][0] <= plans[r][k][1] for r in range(world - 1))
            if world == 1 or n_vars < 1000:
                continue
            wt = (wh, 1, 1, 1, w2)
            load = []
            for r in range(world):
                owned = pkg.multi_gpu.owned_polys(r, world)[1]
                load.append(sum((hi - lo) * wt[k] for k, (lo, hi) in enumerate(plans[r])) + owned * wntt * N)
                partial = sum(1 for k, (lo, hi) in enumerate(plans[r]) if hi > lo and (lo > 0 or hi < sizes[k]))
                assert partial <= 2, (world, r, plans[r])
            total = sum(load)
            # snapping moves a cut by at most 4 % of a section
            assert max(load) - min(load) <= 0.09 * max(sizes) * w2 + 2 * w2, (world, load)
            assert abs(total - (wh * N + (3 + w2) * n_vars + 3 * wntt * N)) < 1e-3 * total
    # mode 0 is the uniform cut of b200_shard_range; -1 follows the environment
    assert pkg.multi_gpu.shard_plan(lib, 1000, 1024, 1, 4, 0) == [(256, 512)] + [(250, 500)] * 4
    monkeypatch.setenv("B200_SHARD_PLAN", "uniform")
    assert pkg.multi_gpu.shard_plan(lib, 1000, 1024, 1, 4, -1) == pkg.multi_gpu.shard_plan(lib, 1000, 1024, 1, 4, 0)
    monkeypatch.setenv("B200_SHARD_PLAN", "line")
    assert pkg.multi_gpu.shard_plan(lib, 1000, 1024, 1, 4, -1) == pkg.multi_gpu.shard_plan(lib, 1000, 1024, 1, 4, 1)
    import ctypes as C
    lo, hi = (C.c_uint32 * 5)(), (C.c_uint32 * 5)()
    assert lib.dll.b200_shard_plan(C.c_uint32(10), C.c_uint32(16), 4, 4, 1, C.c_double(0), lo, hi) != 0
    assert lib.dll.b200_shard_plan(C.c_uint32(10), C.c_uint32(16), 0, 4, 2, C.c_double(0), lo, hi) != 0
    assert lib.dll.b200_shard_plan(C.c_uint32(10), C.c_uint32(16), 0, 4, 1, C.c_double(0), None, hi) != 0


def _comm_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conftest  # noqa: F401
    import torch.distributed as dist
    import icicle_snark_b200 as pkg
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        lib = pkg.lib()
        try:
            pkg.multi_gpu.LibComm.from_torch(lib)
            q.put((rank, "created"))
        except RuntimeError as exc:
            q.put((rank, "error: " + str(exc)[:60]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_comm_setup_fails_on_every_rank_without_a_gpu():
    """LibComm's rendezvous (rank 0 draws the token, status + token broadcast by the host's own channel - gloo here): without
    a CUDA device the library cannot join a communicator, and that must surface as an error on EVERY rank, never as a hang
    of the ranks that wait for rank 0 (bench.py then falls back to the torch.distributed exchange)."""
    from conftest import HAS_GPU
    if HAS_GPU:
        pytest.skip("CPU-only behaviour")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_comm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=5) for _ in range(2))
    assert set(got) == {0, 1} and all(v.startswith("error") for v in got.values()), got
