"""GPU: the fused proving path (b200_groth16_*) through the C ABI against the committed golden proofs
(byte-identical proof.json for fixed blinding factors), against the oracle pipeline on freshly generated
instances, and under the reference's pairing check for random blinding."""
import ctypes as C
import os

import numpy as np
import pytest

import icicle_snark_b200 as pkg
from oracle import groth16_ref as G
from test_groth16_oracle import FIXED_R, FIXED_S, GOLD, load
from tools import synth

pytestmark = pytest.mark.gpu


def wtns_words(wtns):
    return G.parse_wtns(wtns)["w"]


@pytest.mark.parametrize("n", [6, 100])
@pytest.mark.parametrize("precompute", [1, 4])
def test_golden_proofs_byte_identical(gpu, ref, n, precompute):
    zkey, wtns, vk, gold11, goldrs, goldpub = load(n)
    cache = pkg.ZKeyCache(gpu, zkey, precompute=precompute)
    try:
        assert (cache.n_vars, cache.n_public) == (n + 2, 1)
        w = wtns_words(wtns)
        p11, tm = cache.prove(w, 1, 1)
        assert pkg.proof_json(p11) == gold11
        prs, _ = cache.prove(w, FIXED_R, FIXED_S)
        assert pkg.proof_json(prs) == goldrs
        # random blinding (r = s = NULL): different proofs, both verify under the reference's pairing
        pa, _ = cache.prove(w)
        pb, _ = cache.prove(w)
        assert pkg.proof_json(pa) != pkg.proof_json(pb)
        public = [int(x) for x in eval(goldpub)]
        for p in (pa, pb, p11):
            assert G.verify(ref, pkg.proof_to_dict(p), public, vk)
        with pytest.raises(pkg.IcicleError):
            cache.prove(np.ascontiguousarray(w[:-1]), 1, 1)  # "Invalid witness length"
    finally:
        cache.close()


def test_file_level_api_matches_golden(gpu, tmp_path, monkeypatch):
    # groth16_prove(witness, zkey, proof, public, device, &mut CacheManager) (src/lib.rs:33-61), C writer
    base = os.path.join(GOLD, "complex_100")
    monkeypatch.setenv("B200_NO_RANDOMNESS", "1")
    proof_p, public_p = str(tmp_path / "proof.json"), str(tmp_path / "public.json")
    for _ in range(2):  # second call hits the process-wide cache
        rc = gpu.dll.b200_groth16_prove_files((base + ".wtns").encode(), (base + ".zkey").encode(), proof_p.encode(),
                                              public_p.encode(), b"CUDA")
        assert rc == 0
        assert open(proof_p).read() == open(base + ".proof_r1s1.json").read()
        assert open(public_p).read() == open(base + ".public.json").read()
    assert gpu.dll.b200_groth16_prove_files(b"/nonexistent.wtns", (base + ".zkey").encode(), proof_p.encode(),
                                            public_p.encode(), b"CUDA") != 0
    assert gpu.dll.b200_groth16_prove_files((base + ".wtns").encode(), (base + ".zkey").encode(), proof_p.encode(),
                                            public_p.encode(), b"CPU") == 1  # no CPU backend
    # python mirror of the same entry point
    cm = pkg.CacheManager(gpu)
    pkg.groth16_prove(base + ".wtns", base + ".zkey", proof_p, public_p, "CUDA", cm, r=FIXED_R, s=FIXED_S)
    assert open(proof_p).read() == open(base + ".proof_rs.json").read()
    with pytest.raises(pkg.IcicleError):
        pkg.groth16_prove(base + ".wtns", base + ".zkey", proof_p, public_p, "CPU", cm)


@pytest.mark.parametrize("n", [1000, 30000])
def test_fresh_instance_matches_oracle_pipeline(gpu, ref, n):
    # setup generated on the GPU (fixed-base tool), proved on the GPU and by the reference CPU pipeline
    zkey, wtns, vk = synth.make_complex_circuit(gpu, n, seed=b"fresh%d" % n)
    proof_ref, public = G.prove(ref, pkg.bindings, zkey, wtns, FIXED_R, FIXED_S)
    assert G.verify(ref, proof_ref, public, vk)  # also validates the GPU-generated setup
    cache = pkg.ZKeyCache(gpu, zkey)
    try:
        p, tm = cache.prove(wtns_words(wtns), FIXED_R, FIXED_S)
        assert pkg.proof_json(p) == G.proof_json(proof_ref)
        assert tm.total_ms > 0
    finally:
        cache.close()
    if n >= 30000:
        # the bench's N = 8 shape in small: eight shards with precompute tables (16 windows of 17 bits, one bucket set)
        caches = [pkg.ZKeyCache(gpu, zkey, precompute=16, rank=r, world=8) for r in range(8)]
        try:
            parts = [c.commit_partials(wtns_words(wtns))[0] for c in caches]
            assert pkg.proof_json(caches[0].finish(parts, FIXED_R, FIXED_S)) == G.proof_json(proof_ref)
        finally:
            for c in caches:
                c.close()


def test_sharded_partials_fold_to_the_same_proof(gpu):
    # SURVEY 8e on one device: 3 "ranks" hold contiguous shards; partial sums folded by finish()
    zkey, wtns, vk, gold11, goldrs, _ = load(100)
    w = wtns_words(wtns)
    caches = [pkg.ZKeyCache(gpu, zkey, rank=r, world=3) for r in range(3)]
    try:
        parts = [c.commit_partials(w)[0] for c in caches]
        proof = caches[0].finish(parts, FIXED_R, FIXED_S)
        assert pkg.proof_json(proof) == goldrs
        with pytest.raises(pkg.IcicleError):
            caches[1].prove(w, 1, 1)  # a sharded cache cannot prove alone
    finally:
        for c in caches:
            c.close()


@pytest.mark.parametrize("plan", ["line", "uniform"])
@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("sparse_b", ["0", "1"])
def test_shard_plans_fold_to_the_golden_proof(gpu, monkeypatch, plan, world, sparse_b):
    """Every cut of the five base-point sections over `world` ranks (b200_shard_plan: the cost-weighted line cut - ranks
    hold different parts of different sections, some none of a section - and the uniform cut), dense and compacted B
    tables: the partial sums fold to the committed golden proof byte for byte, and the ranges partition every section."""
    monkeypatch.setenv("B200_SHARD_PLAN", plan)
    monkeypatch.setenv("B200_SPARSE_B", sparse_b)
    for n in (6, 100):
        zkey, wtns, vk, gold11, goldrs, _ = load(n)
        w = wtns_words(wtns)
        caches = [pkg.ZKeyCache(gpu, zkey, precompute=(1, 16)[r & 1], rank=r, world=world) for r in range(world)]
        try:
            rng_all = [c.ranges() for c in caches]
            sizes = (caches[0].domain_size,) + (caches[0].n_vars,) * 4
            for k in range(5):
                assert rng_all[0][k][0] == 0 and rng_all[-1][k][1] == sizes[k]
                assert all(rng_all[r][k][1] == rng_all[r + 1][k][0] for r in range(world - 1))
            assert rng_all == [pkg.multi_gpu.shard_plan(gpu, caches[0].n_vars, caches[0].domain_size, r, world) for r in range(world)]
            parts = [c.commit_partials(w)[0] for c in caches]
            assert pkg.proof_json(caches[0].finish(parts, FIXED_R, FIXED_S)) == goldrs
            parts = [c.commit_partials(w)[0] for c in caches]  # second proof on the same caches
            assert pkg.proof_json(caches[0].finish(parts, 1, 1)) == gold11
        finally:
            for c in caches:
                c.close()


def test_skewed_shards_fold_to_the_same_proof(gpu, monkeypatch):
    # B200_SHARD_SKEW (bench.py sets it when the quotient chain is split): polynomial owners hold smaller witness shards;
    # any partition must fold to the same proof
    zkey, wtns, vk, gold11, goldrs, _ = load(100)
    w = wtns_words(wtns)
    monkeypatch.setenv("B200_SHARD_SKEW", "0.1")
    monkeypatch.setenv("B200_SHARD_PLAN", "uniform")
    caches = [pkg.ZKeyCache(gpu, zkey, precompute=p, rank=r, world=4) for r, p in zip(range(4), (1, 16, 1, 16))]
    try:
        sizes = [c.b_points()[1] for c in caches]
        assert sum(sizes) == caches[0].n_vars and sizes[3] > sizes[0] and sizes[0] == sizes[1] == sizes[2]
        assert [c.h_range() for c in caches] == [pkg.multi_gpu.shard_range(caches[0].domain_size, r, 4) for r in range(4)]
        parts = [c.commit_partials(w)[0] for c in caches]
        assert pkg.proof_json(caches[0].finish(parts, FIXED_R, FIXED_S)) == goldrs
    finally:
        for c in caches:
            c.close()


def test_fixed_base_tool_against_reference(gpu, ref, rng):
    from util import rand_scalars
    k, _ = rand_scalars(rng, 50)
    k[0] = 0
    for g2 in (False, True):
        pts = synth.fixed_base(gpu, k, g2=g2)
        std = ref.convert_montgomery(pts, False, kind="g2_affine" if g2 else "affine")
        gen = ref.generator(g2=g2)
        for i in range(50):
            assert np.array_equal(std[i], ref.to_affine(ref.mul_scalar(gen, k[i], g2=g2), g2=g2)), (g2, i)


def test_msm_known_dlog_at_scale(gpu, rng):
    # P_i = k_i G  =>  MSM(s, P) == (sum s_i k_i mod r) G, at a size the CPU oracle is slow at
    from util import R, array_to_ints, ints_to_array, rand_scalars
    n = 1 << 20
    k, kv = rand_scalars(rng, n)
    s, sv = rand_scalars(rng, n)
    pts = synth.fixed_base(gpu, k)  # Montgomery
    cfg = pkg.MSMConfig.default()
    cfg.are_points_montgomery_form = True
    got = gpu.msm(s, pts, cfg)[0]
    dot = sum(a * b for a, b in zip(sv, kv)) % R
    want = gpu.mul_scalar(gpu.generator(), ints_to_array([dot])[0])
    assert np.array_equal(gpu.to_affine(got), gpu.to_affine(want))


def test_msm_known_dlog_with_precompute_tables_at_bench_plan(gpu, rng):
    """The plan the bench's proof runs - precompute factor 16, c = 20, 13 windows, one bucket set - at 2^22 points, for
    G1 and (2^20) G2: MSM(s, k_i G) == (sum s_i k_i) G.  G2 at this size takes the batched affine accumulation."""
    from util import R, ints_to_array, rand_scalars
    for g2, lg in ((False, 22), (True, 20)):
        n = 1 << lg
        k, kv = rand_scalars(rng, n)
        s, sv = rand_scalars(rng, n)
        pts = synth.fixed_base(gpu, k, g2=g2)
        cfg = pkg.MSMConfig.default()
        cfg.are_points_montgomery_form = True
        cfg.precompute_factor = 16
        table = gpu.msm_precompute_bases(pts, cfg, g2=g2)
        info = (C.c_int32 * 8)()
        assert gpu.dll.b200_msm_plan_info(C.c_int(n), C.c_int(0), C.c_int(254), C.c_int(16), C.c_int(int(g2)), info, None) == 0
        assert info[3] == 1 and info[1] <= 16  # one bucket set: every window served by a precomputed multiple
        got = gpu.msm(s, table, cfg, g2=g2, msm_size=n)[0]
        dot = sum(a * b for a, b in zip(sv, kv)) % R
        want = gpu.mul_scalar(gpu.generator(g2=g2), ints_to_array([dot])[0], g2=g2)
        assert np.array_equal(gpu.to_affine(got, g2=g2), gpu.to_affine(want, g2=g2)), g2


@pytest.mark.parametrize("precompute", [1, 16])
def test_config0_100k_proof_byte_identical_to_reference_pipeline(gpu, ref, precompute):
    """BASELINE configs[0] (benchmark/100k): the GPU proof.json equals the reference CPU pipeline's byte for byte, for
    fixed non-trivial (r, s), without and with the precompute tables."""
    zkey, wtns, vk = synth.make_complex_circuit(gpu, 100_000)
    proof_ref, public = G.prove(ref, pkg.bindings, zkey, wtns, FIXED_R, FIXED_S)
    assert G.verify(ref, proof_ref, public, vk)
    cache = pkg.ZKeyCache(gpu, zkey, precompute=precompute)
    try:
        p, _ = cache.prove(wtns_words(wtns), FIXED_R, FIXED_S)
        assert pkg.proof_json(p) == G.proof_json(proof_ref)
    finally:
        cache.close()


@pytest.mark.parametrize("n", [30_000, 800_000])
def test_reference_cuda_backend_agrees_byte_for_byte(gpu, ref, n):
    """Second oracle: the reference's own CUDA backend, rebuilt for sm_100a from its sources
    (oracle/Makefile.ref_cuda) and driven by the restated Rust host with the Rust code's residency
    (oracle/groth16_ref_cuda.py), proves the same instance with the same (r, s) on this GPU: proof.json must be
    identical to the product's - at 800k, a BASELINE size where the CPU oracle is too slow to run in a test."""
    from oracle import groth16_ref_cuda as GC
    if not GC.available():
        pytest.skip("oracle/_ref_cuda not built (needs /root/reference at build time)")
    zkey, wtns, vk = synth.make_complex_circuit(gpu, n, seed=b"refcuda%d" % n)
    rc = GC.ref_cuda(0)
    try:
        cache_ref = GC.ZKeyCacheCuda(rc, pkg.bindings, zkey)
        try:
            proof_ref, public = GC.prove(rc, pkg.bindings, wtns, FIXED_R, FIXED_S, cache_ref)
        finally:
            cache_ref.close()
            rc.ntt_release_domain()
    finally:
        rc.set_device("CPU", 0)
    assert G.verify(ref, proof_ref, public, vk)
    cache = pkg.ZKeyCache(gpu, zkey, precompute=16)
    try:
        p, _ = cache.prove(wtns_words(wtns), FIXED_R, FIXED_S)
        assert pkg.proof_json(p) == G.proof_json(proof_ref)
    finally:
        cache.close()


def test_aadhaar_shaped_substitute_matches_oracle(gpu, ref):
    """configs[3] substitute (anon_aadhaar itself needs circom/circomlib/snarkjs): random satisfiable R1CS with
    multi-entry rows (collisions in the A/B accumulation), 9 public inputs and a ~90 % 0/1 witness (giant buckets:
    exercises the item split + fold path of the MSM inside the prover)."""
    zkey, wtns, vk = synth.make_random_circuit(gpu, 20000, n_inputs=300, n_public=9)
    w = wtns_words(wtns)
    vals = [int.from_bytes(x.tobytes(), "little") for x in w]
    assert sum(v in (0, 1) for v in vals) / len(vals) > 0.8
    proof_ref, public = G.prove(ref, pkg.bindings, zkey, wtns, FIXED_R, FIXED_S)
    assert len(public) == 9 and G.verify(ref, proof_ref, public, vk)
    for precompute in (1, 16):
        cache = pkg.ZKeyCache(gpu, zkey, precompute=precompute)
        try:
            p, _ = cache.prove(w, FIXED_R, FIXED_S)
            assert pkg.proof_json(p) == G.proof_json(proof_ref)
        finally:
            cache.close()


def test_full_size_proof_verifies_under_reference_pairing(gpu, ref):
    """Size-independent property at a benchmark size (800k constraints): the GPU proof with random blinding verifies
    under the reference's own pairing; a second proof of the same witness differs (fresh r, s) and verifies too."""
    zkey, wtns, vk = synth.make_complex_circuit(gpu, 800_000)
    cache = pkg.ZKeyCache(gpu, zkey, precompute=16)
    try:
        w = wtns_words(wtns)
        public = [int.from_bytes(w[1].tobytes(), "little")]
        p1, tm = cache.prove(w)
        p2, _ = cache.prove(w)
        assert pkg.proof_json(p1) != pkg.proof_json(p2)
        assert G.verify(ref, pkg.proof_to_dict(p1), public, vk) and G.verify(ref, pkg.proof_to_dict(p2), public, vk)
        assert not G.verify(ref, pkg.proof_to_dict(p1), [public[0] ^ 1], vk)
        # the library's own verifier (host pairing, csrc/pairing.cu) agrees on the raw proof structs
        assert pkg.groth16_verify_points(gpu, p1, public, vk) and pkg.groth16_verify_points(gpu, p2, public, vk)
        assert not pkg.groth16_verify_points(gpu, p1, [public[0] ^ 1], vk)
    finally:
        cache.close()


def test_quotient_split_api_single_rank(gpu):
    """commit_begin / commit_end (the N > 1 quotient split) driven on one device: all three polynomials transformed
    into a caller buffer, slices handed back, same proof as the fused path."""
    import torch
    zkey, wtns, vk, gold11, goldrs, _ = load(100)
    w = wtns_words(wtns)
    cache = pkg.ZKeyCache(gpu, zkey)
    try:
        N = cache.domain_size
        lo, hi = cache.h_range()
        assert (lo, hi) == (0, N)
        buf = torch.empty((3, N, 8), dtype=torch.int32, device="cuda")
        cache.commit_begin(w, 0, 3, buf.data_ptr())
        parts, tm = cache.commit_end(buf[1].data_ptr(), buf[0].data_ptr(), buf[2].data_ptr())
        proof = cache.finish([parts], FIXED_R, FIXED_S)
        assert pkg.proof_json(proof) == goldrs
        # a rank that owns no polynomial (world > 3): poly_count = 0, slices come from the other ranks
        cache.commit_begin(w, 0, 0, buf.data_ptr())
        parts0, _ = cache.commit_end(buf[1].data_ptr(), buf[0].data_ptr(), buf[2].data_ptr())
        assert pkg.proof_json(cache.finish([parts0], FIXED_R, FIXED_S)) == goldrs
        # the regular path still works afterwards (lock released)
        p2, _ = cache.prove(w, 1, 1)
        assert pkg.proof_json(p2) == gold11
    finally:
        cache.close()


def test_sparse_b_columns_are_dropped_and_proofs_unchanged(gpu, ref, monkeypatch):
    """Signals absent from every B row have B1/B2 points at infinity; the cache drops them (second, shorter sort for
    B1/B2). Forced on a dense golden instance, natural on the random circuit; proofs must not change."""
    zkey, wtns, vk, gold11, goldrs, _ = load(100)
    w = wtns_words(wtns)
    monkeypatch.setenv("B200_SPARSE_B", "1")
    for precompute, world in ((1, 1), (16, 1), (1, 3)):
        caches = [pkg.ZKeyCache(gpu, zkey, precompute=precompute, rank=r, world=world) for r in range(world)]
        try:
            kept = sum(c.b_points()[0] for c in caches)
            assert kept == 100  # signals 0 and 1 (one, c) never appear in B
            parts = [c.commit_partials(w)[0] for c in caches]
            assert pkg.proof_json(caches[0].finish(parts, FIXED_R, FIXED_S)) == goldrs
        finally:
            for c in caches:
                c.close()
    monkeypatch.delenv("B200_SPARSE_B")
    zkey, wtns, vk = synth.make_random_circuit(gpu, 6000, n_inputs=200, n_public=9, seed=b"sparse", rng_seed=11)
    proof_ref, public = G.prove(ref, pkg.bindings, zkey, wtns, FIXED_R, FIXED_S)
    cache = pkg.ZKeyCache(gpu, zkey, precompute=16)
    try:
        kept, total = cache.b_points()
        assert kept < total * 7 // 8  # the heuristic switched the sparse path on by itself
        p, _ = cache.prove(wtns_words(wtns), FIXED_R, FIXED_S)
        assert pkg.proof_json(p) == G.proof_json(proof_ref)
    finally:
        cache.close()


def test_foreign_ntt_domain_is_replaced(gpu, ref):
    """bn254_ntt_init_domain accepts ANY primitive root; the prover's coset powers and the zkey's H points are tied to the
    standard generator.  A global domain that another caller initialised with a different root (here: the cube of the
    standard one - also primitive, large enough) must not be reused: the proof stays byte-identical to the golden file."""
    zkey, wtns, vk, gold11, goldrs, _ = load(100)
    w = wtns_words(wtns)
    gpu.ntt_release_domain()
    std = gpu.get_root_of_unity(1 << 12)
    cube = gpu.fr_mul(gpu.fr_mul(std, std), std)
    gpu.ntt_init_domain(cube)
    try:
        cache = pkg.ZKeyCache(gpu, zkey)
        try:
            assert pkg.proof_json(cache.prove(w, FIXED_R, FIXED_S)[0]) == goldrs
            # ... and again when the foreign domain appears between two proofs on a warm cache
            gpu.ntt_release_domain()
            gpu.ntt_init_domain(cube)
            assert pkg.proof_json(cache.prove(w, 1, 1)[0]) == gold11
        finally:
            cache.close()
    finally:
        gpu.ntt_release_domain()


def test_pageable_witness_takes_the_staged_upload(gpu, tmp_path, monkeypatch):
    """A pageable host witness (numpy memory, an mmap'd .wtns) is copied through the pinned staging buffer by several
    threads in chunks (copy_witness_range); forced on at test sizes.  Same proofs as the direct path, on one device and on
    sharded caches (slice-first upload), through the struct API and the file-level API."""
    zkey, wtns, vk, gold11, goldrs, _ = load(100)
    w = wtns_words(wtns)
    monkeypatch.setenv("B200_STAGE_MIN_BYTES", "64")
    cache = pkg.ZKeyCache(gpu, zkey, precompute=4)
    try:
        for _ in range(2):
            assert pkg.proof_json(cache.prove(w, FIXED_R, FIXED_S)[0]) == goldrs
    finally:
        cache.close()
    caches = [pkg.ZKeyCache(gpu, zkey, rank=r, world=3) for r in range(3)]
    try:
        parts = [c.commit_partials(w)[0] for c in caches]
        assert pkg.proof_json(caches[0].finish(parts, 1, 1)) == gold11
    finally:
        for c in caches:
            c.close()
    base = os.path.join(GOLD, "complex_100")
    monkeypatch.setenv("B200_NO_RANDOMNESS", "1")
    proof_p, public_p = str(tmp_path / "proof.json"), str(tmp_path / "public.json")
    assert gpu.dll.b200_groth16_prove_files((base + ".wtns").encode(), (base + ".zkey").encode(), proof_p.encode(),
                                            public_p.encode(), b"CUDA") == 0
    assert open(proof_p).read() == open(base + ".proof_r1s1.json").read()
