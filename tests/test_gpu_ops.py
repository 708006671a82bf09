"""Parity of the CUDA path (through the C ABI, host buffers unless stated) against the reference's own
CPU backend compiled into oracle/_ref, on the same seeded inputs. Shapes follow the reference's tests:
wrappers/rust/icicle-core/src/msm/tests.rs:24-302, ntt/tests.rs:38-354, vec_ops/tests.rs:36-382,
tests.rs:85-163. Bit-exact: field outputs compared limb for limb, points compared canonically (affine)."""
import ctypes as C

import numpy as np
import pytest

import icicle_snark_b200 as pkg
from util import R, array_to_ints, from_words, ints_to_array, rand_scalars, to_words

B = pkg.bindings
pytestmark = pytest.mark.gpu


def affine_eq(lib, a, b, g2=False):
    return np.array_equal(lib.to_affine(a, g2=g2), lib.to_affine(b, g2=g2))


# ------------------------------------------------------------------------------------------- vec ops
@pytest.mark.parametrize("n", [1, 7, 1000, 1 << 16])
def test_vec_ops_match_reference(gpu, ref, rng, n):
    a, _ = rand_scalars(rng, n)
    b, _ = rand_scalars(rng, n)
    a[0] = to_words(R - 1)
    b[0] = to_words(R - 1)
    if n > 2:
        a[1] = 0
        b[2] = to_words(1)
    for op in ("vector_add", "vector_sub", "vector_mul"):
        assert np.array_equal(getattr(gpu, op)(a, b), getattr(ref, op)(a, b)), op
    if n <= 1000:
        b[b.sum(axis=1) == 0] = to_words(5)
        assert np.array_equal(gpu.vector_div(a, b), ref.vector_div(a, b))
    assert np.array_equal(gpu.vector_sum(a), ref.vector_sum(a))
    if n <= 1000:
        assert np.array_equal(gpu.vector_product(a), ref.vector_product(a))
    s = a[:1].copy()
    for op in ("scalar_add_vec", "scalar_sub_vec", "scalar_mul_vec"):
        assert np.array_equal(getattr(gpu, op)(s, b), getattr(ref, op)(s, b)), op
    acc_g, acc_r = a.copy(), a.copy()
    gpu.vector_accumulate(acc_g, b)
    ref.vector_accumulate(acc_r, b)
    assert np.array_equal(acc_g, acc_r)


def test_vec_ops_device_operands_and_tracker(gpu, ref, rng):
    n = 4096
    a, _ = rand_scalars(rng, n)
    b, _ = rand_scalars(rng, n)
    da, db, dout = gpu.malloc(a.nbytes), gpu.malloc(b.nbytes), gpu.malloc(a.nbytes)
    try:
        assert gpu.is_active_device_memory(da) and gpu.is_active_device_memory(da + 32 * 100)  # interior pointer
        assert not gpu.is_active_device_memory(a.ctypes.data) and gpu.is_host_memory(a.ctypes.data)
        gpu.copy_to_device(da, a)
        gpu.copy_to_device(db, b)
        cfg = B.VecOpsConfig.default()
        cfg.is_a_on_device = cfg.is_b_on_device = cfg.is_result_on_device = True
        st = gpu.create_stream()
        cfg.stream = st
        cfg.is_async = True
        out = np.empty_like(a)
        B.check(gpu.dll.bn254_vector_mul(C.c_void_p(da), C.c_void_p(db), C.c_uint64(n), C.byref(cfg), C.c_void_p(dout)))
        # second half minus first half, through interior device pointers (the prover's &d_vec[N..2N] pattern)
        h = n // 2
        B.check(gpu.dll.bn254_vector_sub(C.c_void_p(dout + 32 * h), C.c_void_p(dout), C.c_uint64(h), C.byref(cfg), C.c_void_p(dout)))
        gpu.stream_synchronize(st)
        gpu.destroy_stream(st)
        gpu.copy_to_host(out, dout)
        prod = ref.vector_mul(a, b)
        assert np.array_equal(out[:h], ref.vector_sub(prod[h:], prod[:h]))
        assert np.array_equal(out[h:], prod[h:])
    finally:
        for p in (da, db, dout):
            gpu.free(p)


def test_montgomery_round_trip(gpu, ref, rng):
    x, xv = rand_scalars(rng, 513)
    m = gpu.convert_montgomery(x, True)
    assert np.array_equal(m, ref.convert_montgomery(x, True))
    assert np.array_equal(gpu.convert_montgomery(m, False), x)
    pts = ref.generate_affine_points(33)
    pm = gpu.convert_montgomery(pts, True, kind="affine")
    assert np.array_equal(pm, ref.convert_montgomery(pts, True, kind="affine"))
    assert np.array_equal(gpu.convert_montgomery(pm, False, kind="affine"), pts)
    p2 = ref.generate_affine_points(9, g2=True)
    assert np.array_equal(gpu.convert_montgomery(p2, True, kind="g2_affine"), ref.convert_montgomery(p2, True, kind="g2_affine"))
    # in-place on a device buffer with is_result_on_device left false (field.rs:379-398 quirk): tracker decides
    d = gpu.malloc(x.nbytes)
    try:
        gpu.copy_to_device(d, x)
        cfg = B.VecOpsConfig.default()
        cfg.is_a_on_device = True
        B.check(gpu.dll.bn254_scalar_convert_montgomery(C.c_void_p(d), C.c_uint64(513), C.c_bool(True), C.byref(cfg), C.c_void_p(d)))
        back = np.empty_like(x)
        gpu.copy_to_host(back, d)
        assert np.array_equal(back, m)
    finally:
        gpu.free(d)


# ------------------------------------------------------------------------------------------- MSM
def _msm_inputs(ref, rng, n, g2=False, zeros=True):
    pts = ref.generate_affine_points(n, g2=g2)
    if zeros and n > 4:  # two points at infinity, msm/tests.rs:15-22
        pts[rng.integers(0, n)] = 0
        pts[rng.integers(0, n)] = 0
    sc, _ = rand_scalars(rng, n)
    return sc, pts


@pytest.mark.parametrize("n", [1, 5, 16, 32, 64, 128, 256, 1000, 1 << 14, 1 << 18])
def test_msm_g1_matches_reference(gpu, ref, rng, n):
    sc, pts = _msm_inputs(ref, rng, n)
    if n > 8:
        sc[3] = 0
        sc[4] = to_words(1)
        sc[5] = to_words(R - 1)
    got = gpu.msm(sc, pts)
    want = ref.msm(sc, pts)
    assert ref.is_on_curve(got[0])
    assert ref.eq(got[0], want[0]) and affine_eq(ref, got[0], want[0])


@pytest.mark.parametrize("n", [1, 5, 64, 1000, 1 << 14, 1 << 18])
def test_msm_g2_matches_reference(gpu, ref, rng, n):
    sc, pts = _msm_inputs(ref, rng, n, g2=True)
    got = gpu.msm(sc, pts, g2=True)
    want = ref.msm(sc, pts, g2=True)
    assert ref.eq(got[0], want[0], g2=True) and affine_eq(ref, got[0], want[0], g2=True)


def test_msm_montgomery_flags_device_inputs_async(gpu, ref, rng):
    n = 3000
    sc, pts = _msm_inputs(ref, rng, n)
    want = ref.msm(sc, pts)
    sc_m = ref.convert_montgomery(sc, True)
    pts_m = ref.convert_montgomery(pts, True, kind="affine")
    ds, dp, dr = gpu.malloc(sc.nbytes), gpu.malloc(pts.nbytes), gpu.malloc(96)
    try:
        gpu.copy_to_device(ds, sc_m)
        gpu.copy_to_device(dp, pts_m)
        cfg = B.MSMConfig.default()
        cfg.are_scalars_on_device = cfg.are_points_on_device = cfg.are_results_on_device = True
        cfg.are_scalars_montgomery_form = cfg.are_points_montgomery_form = True
        st = gpu.create_stream()
        cfg.stream, cfg.is_async = st, True
        gpu.msm(ds, dp, cfg, results=dr, msm_size=n)
        gpu.stream_synchronize(st)
        gpu.destroy_stream(st)
        out = np.zeros((1, 24), dtype=np.uint32)
        gpu.copy_to_host(out, dr)
        assert affine_eq(ref, out[0], want[0])
    finally:
        for p in (ds, dp, dr):
            gpu.free(p)


@pytest.mark.parametrize("batch,shared", [(3, True), (3, False), (16, True)])
def test_msm_batch(gpu, ref, rng, batch, shared):
    n = 200
    cfg = B.MSMConfig.default()
    cfg.batch_size, cfg.are_points_shared_in_batch = batch, shared
    sc, _ = rand_scalars(rng, n * batch)
    pts = ref.generate_affine_points(n if shared else n * batch)
    got = gpu.msm(sc, pts, cfg)
    want = ref.msm(sc, pts, cfg)
    for b in range(batch):
        assert affine_eq(ref, got[b], want[b]), b


@pytest.mark.parametrize("g2", [False, True])
def test_msm_precompute_equals_plain(gpu, ref, rng, g2):
    # msm/tests.rs:134-165: precompute_factor 8, c = 4 must not change the result
    n = 300
    sc, pts = _msm_inputs(ref, rng, n, g2=g2)
    want = ref.msm(sc, pts, g2=g2)
    cfg = B.MSMConfig.default()
    cfg.precompute_factor, cfg.c = 8, 4
    table = gpu.msm_precompute_bases(pts, cfg, g2=g2)
    assert table.shape[0] == n * 8
    assert np.array_equal(table[::8], pts)  # entry j = 0 is the point itself (cuda_msm.cuh:29-43 layout)
    got = gpu.msm(sc, table, cfg, g2=g2, msm_size=n)
    assert affine_eq(ref, got[0], want[0], g2=g2)


@pytest.mark.parametrize("n", [1000, 1 << 15])
def test_msm_skewed_distribution(gpu, ref, rng, n):
    # msm/tests.rs:254-302: mostly 0/1 scalars -> a few giant buckets (exercises the item split + fold)
    sc, pts = _msm_inputs(ref, rng, n)
    pick = rng.integers(0, 10, size=n)
    sc[pick < 4] = 0
    sc[(pick >= 4) & (pick < 8)] = to_words(1)
    sc[pick == 8] = to_words(R - 1)
    got = gpu.msm(sc, pts)
    want = ref.msm(sc, pts)
    assert affine_eq(ref, got[0], want[0])
    # repeated points (P == Q inside a bucket) and P, -P pairs
    pts[: n // 2] = pts[0]
    got = gpu.msm(sc, pts)
    want = ref.msm(sc, pts)
    assert affine_eq(ref, got[0], want[0])


@pytest.mark.parametrize("bitsize", [1, 64, 128])
def test_msm_bitsize(gpu, ref, rng, bitsize):
    # MSMConfig.bitsize: scalars known to be shorter than 254 bits (msm.h:44-47); fewer windows, same sum
    n = 2000
    sc, pts = _msm_inputs(ref, rng, n)
    mask = np.zeros(8, dtype=np.uint32)
    for b in range(bitsize):
        mask[b // 32] |= np.uint32(1 << (b % 32))
    sc &= mask
    cfg = B.MSMConfig.default()
    cfg.bitsize = bitsize
    got = gpu.msm(sc, pts, cfg)
    want = ref.msm(sc, pts)
    assert affine_eq(ref, got[0], want[0])


@pytest.mark.parametrize("shared", [True, False])
def test_msm_precompute_batch(gpu, ref, rng, shared):
    # precompute_factor x batch, shared and per-MSM point sets (msm/tests.rs:167-252)
    n, batch, f = 150, 3, 4
    cfg = B.MSMConfig.default()
    cfg.batch_size, cfg.are_points_shared_in_batch, cfg.precompute_factor = batch, shared, f
    sc, _ = rand_scalars(rng, n * batch)
    pts = ref.generate_affine_points(n if shared else n * batch)
    plain = B.MSMConfig.default()
    plain.batch_size, plain.are_points_shared_in_batch = batch, shared
    want = ref.msm(sc, pts, plain)
    table = gpu.msm_precompute_bases(pts, cfg)
    got = gpu.msm(sc, table, cfg, msm_size=n)
    for b in range(batch):
        assert affine_eq(ref, got[b], want[b]), b


@pytest.mark.parametrize("g2", [False, True])
@pytest.mark.parametrize("chunk", [1, 700, 4096])
def test_msm_chunked_passes_match_reference(gpu, ref, rng, monkeypatch, g2, chunk):
    """Oversize / host-resident inputs take several Pippenger passes (msm_chunked; the reference's multi-chunk path,
    cuda_msm.cuh:1130-1237).  B200_MSM_CHUNK forces the split at test sizes: ragged last chunk, one-point chunks, host
    operands (double-buffered staging) and device operands, standard-form points (per-chunk Montgomery copy)."""
    n = 5 if chunk == 1 else 3001
    sc, pts = _msm_inputs(ref, rng, n, g2=g2)
    want = ref.msm(sc, pts, g2=g2)
    monkeypatch.setenv("B200_MSM_CHUNK", str(chunk))
    got = gpu.msm(sc, pts, g2=g2)  # host operands
    assert affine_eq(ref, got[0], want[0], g2=g2)
    ds, dp = gpu.malloc(sc.nbytes), gpu.malloc(pts.nbytes)
    try:
        gpu.copy_to_device(ds, sc)
        gpu.copy_to_device(dp, pts)
        cfg = B.MSMConfig.default()
        cfg.are_scalars_on_device = cfg.are_points_on_device = True
        got = gpu.msm(ds, dp, cfg, g2=g2, msm_size=n)  # device operands, host result
        assert affine_eq(ref, got[0], want[0], g2=g2)
    finally:
        gpu.free(ds)
        gpu.free(dp)


def test_msm_chunked_with_precompute_and_batch(gpu, ref, rng, monkeypatch):
    """A precomputed table is built for the FULL problem's window width; every chunk must keep that width."""
    n, batch, f = 2500, 2, 8
    cfg = B.MSMConfig.default()
    cfg.batch_size, cfg.are_points_shared_in_batch, cfg.precompute_factor = batch, True, f
    sc, _ = rand_scalars(rng, n * batch)
    pts = ref.generate_affine_points(n)
    plain = B.MSMConfig.default()
    plain.batch_size, plain.are_points_shared_in_batch = batch, True
    want = ref.msm(sc, pts, plain)
    table = gpu.msm_precompute_bases(pts, cfg)
    monkeypatch.setenv("B200_MSM_CHUNK", "999")
    got = gpu.msm(sc, table, cfg, msm_size=n)
    for b in range(batch):
        assert affine_eq(ref, got[b], want[b]), b


def test_division_step_inverse_on_device(gpu):
    # csrc/field_inv.cuh on the GPU: 2^17 inversions per field against x * x^-1 == 1 (and Fermat on the first few)
    f = pkg.tools_lib().b200_inv_check
    f.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int]
    assert f(1 << 17, 11, 1, 1) == 0 and f(1 << 17, 12, 0, 1) == 0


def test_msm_suite_with_batched_affine_forced():
    """The whole MSM parity suite again with the batched affine accumulation forced on for G1 and G2 at every size
    (B200_BATCH_AFFINE is read once per process, hence the subprocess): tiny buckets, empty buckets, identities, repeated
    and opposite points, the skewed distributions' giant buckets (the long-leftover CTA path)."""
    import os
    import subprocess
    import sys
    if os.environ.get("B200_BATCH_AFFINE") == "3":
        pytest.skip("already inside the forced run")
    env = dict(os.environ, B200_BATCH_AFFINE="3")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k", "msm and not forced and not thread_per_chunk"],
                       env=env, capture_output=True, text=True, timeout=1200, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_msm_suite_with_thread_per_chunk_reduction():
    """The bucket reduction's upper levels run quad-cooperatively by default (csrc/msm_reduce_quad.cuh); B200_MSM_QUAD=0
    selects the one-thread-per-chunk kernels (still the level-0 path of large MSMs).  Same parity suite, read once per
    process, hence the subprocess."""
    import os
    import subprocess
    import sys
    if os.environ.get("B200_MSM_QUAD") == "0":
        pytest.skip("already inside the forced run")
    env = dict(os.environ, B200_MSM_QUAD="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k",
                        "msm and not forced and not thread_per_chunk and not chunked"],
                       env=env, capture_output=True, text=True, timeout=1200, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_msm_all_zero_and_empty(gpu, ref, rng):
    n = 100
    sc = np.zeros((n, 8), dtype=np.uint32)
    pts = ref.generate_affine_points(n)
    got = gpu.msm(sc, pts)
    assert not got[0][:8].any() and got[0][8:16].any() and not got[0][16:].any()  # identity (0, y != 0, 0)
    pts[:] = 0
    sc, _ = rand_scalars(rng, n)
    got = gpu.msm(sc, pts)
    assert not got[0][16:].any() and got[0][8:16].any()


# ------------------------------------------------------------------------------------------- NTT
@pytest.fixture(scope="module")
def domains(gpu, ref):
    root = ref.get_root_of_unity(1 << 20)
    assert np.array_equal(root, gpu.get_root_of_unity(1 << 20))
    gpu.ntt_release_domain()  # initialising over an existing (smaller) domain is a no-op, as in the reference
    ref.ntt_release_domain()
    gpu.ntt_init_domain(root)
    ref.ntt_init_domain(root)
    yield 20
    gpu.ntt_release_domain()
    ref.ntt_release_domain()


@pytest.mark.parametrize("logn", [0, 1, 2, 3, 4, 7, 8, 9, 11, 16, 17, 18])
def test_ntt_forward_inverse_match_reference(gpu, ref, rng, domains, logn):
    x, _ = rand_scalars(rng, 1 << logn)
    f = gpu.ntt(x, B.kForward)
    assert np.array_equal(f, ref.ntt(x, B.kForward))
    assert np.array_equal(gpu.ntt(x, B.kInverse), ref.ntt(x, B.kInverse))
    assert np.array_equal(gpu.ntt(f, B.kInverse), x)  # NTT o iNTT == id (ntt/tests.rs:75-93)


@pytest.mark.parametrize("ordering", [B.kNN, B.kNR, B.kRN, B.kRR])
@pytest.mark.parametrize("logn", [4, 10, 13])
def test_ntt_orderings(gpu, ref, rng, domains, ordering, logn):
    x, _ = rand_scalars(rng, 1 << logn)
    cfg = B.NTTConfig.default()
    cfg.ordering = ordering
    for d in (B.kForward, B.kInverse):
        assert np.array_equal(gpu.ntt(x, d, cfg), ref.ntt(x, d, cfg)), (ordering, d)


def test_ntt_mixed_orderings_are_self_consistent(gpu, rng, domains):
    # kNM / kMN only need to invert each other (ntt/tests.rs:169-229)
    x, _ = rand_scalars(rng, 1 << 12)
    cfg = B.NTTConfig.default()
    cfg.ordering = B.kNM
    f = gpu.ntt(x, B.kForward, cfg)
    cfg.ordering = B.kMN
    assert np.array_equal(gpu.ntt(f, B.kInverse, cfg), x)


@pytest.mark.parametrize("batch,columns", [(3, False), (3, True), (16, False), (5, True)])
def test_ntt_batch(gpu, ref, rng, domains, batch, columns):
    n = 1 << 10
    x, _ = rand_scalars(rng, n * batch)
    cfg = B.NTTConfig.default()
    cfg.batch_size, cfg.columns_batch = batch, columns
    for d in (B.kForward, B.kInverse):
        assert np.array_equal(gpu.ntt(x, d, cfg), ref.ntt(x, d, cfg))


def test_ntt_coset_and_subgroup_identity(gpu, ref, rng, domains):
    # arbitrary coset vs reference (ntt/tests.rs:231-273)
    n = 1 << 9
    x, _ = rand_scalars(rng, n)
    cfg = B.NTTConfig.default()
    g, _ = rand_scalars(rng, 1)
    for i in range(8):
        cfg.coset_gen[i] = int(g[0][i])
    assert np.array_equal(gpu.ntt(x, B.kForward, cfg), ref.ntt(x, B.kForward, cfg))
    assert np.array_equal(gpu.ntt(x, B.kInverse, cfg), ref.ntt(x, B.kInverse, cfg))
    # size-N NTT == [N/2 NTT of evens-structure || coset N/2 NTT] (ntt/tests.rs:99-166), done in NN order:
    # evaluations of the same polynomial on <w_N> split into <w_{N/2}> and w_N * <w_{N/2}>
    full = gpu.ntt(x, B.kForward)
    half = n // 2
    # reduce the degree-(n-1) polynomial mod X^{n/2} - c on each coset: p(X) = lo(X) + X^{n/2} hi(X)
    lo, hi = x[:half], x[half:]
    even = gpu.ntt(gpu.vector_add(lo, hi), B.kForward)  # on <w_{N/2}>: X^{n/2} == 1
    w_n = gpu.get_root_of_unity(n)
    cfg2 = B.NTTConfig.default()
    for i in range(8):
        cfg2.coset_gen[i] = int(w_n[i])
    odd = gpu.ntt(gpu.vector_sub(lo, hi), B.kForward, cfg2)  # on w_N <w_{N/2}>: X^{n/2} == -1
    assert np.array_equal(full[0::2], even) and np.array_equal(full[1::2], odd)


def test_ntt_inplace_device_async(gpu, ref, rng, domains):
    # the prover's call shape: in place on a device buffer, batch 3, async (src/icicle_helper.rs:13-32)
    n, batch = 1 << 15, 3
    x, _ = rand_scalars(rng, n * batch)
    d = gpu.malloc(x.nbytes)
    try:
        gpu.copy_to_device(d, x)
        cfg = B.NTTConfig.default()
        cfg.batch_size = batch
        cfg.are_inputs_on_device = cfg.are_outputs_on_device = True
        st = gpu.create_stream()
        cfg.stream, cfg.is_async = st, True
        gpu.ntt(d, B.kInverse, cfg, out=d, size=n)
        gpu.stream_synchronize(st)
        out = np.empty_like(x)
        gpu.copy_to_host(out, d)
        cfg_r = B.NTTConfig.default()
        cfg_r.batch_size = batch
        assert np.array_equal(out, ref.ntt(x, B.kInverse, cfg_r))
        gpu.ntt(d, B.kForward, cfg, out=d, size=n)
        gpu.stream_synchronize(st)
        gpu.destroy_stream(st)
        gpu.copy_to_host(out, d)
        assert np.array_equal(out, x)
    finally:
        gpu.free(d)


def test_ntt_errors_are_codes_not_exceptions(gpu, rng, domains):
    x, _ = rand_scalars(rng, 12)
    with pytest.raises(B.IcicleError):
        gpu.ntt(x, B.kForward)  # size not a power of two
    big = np.zeros((1 << 21, 8), dtype=np.uint32)
    with pytest.raises(B.IcicleError):
        gpu.ntt(big, B.kForward)  # larger than the domain (the reference throws across the ABI here)


@pytest.mark.parametrize("logn,batch", [(22, 3), (24, 1)])
def test_ntt_full_size_properties(gpu, rng, logn, batch):
    """BASELINE sizes (2^22 batch 3 as in the prover; 2^24, the top of the configs[4] sweep), in place on the device:
    iNTT(NTT(x)) == x, linearity NTT(x + y) == NTT(x) + NTT(y), and the known answer NTT(e_1)[k] == w^k -
    size-independent properties where the CPU reference would take tens of seconds."""
    import torch
    n = 1 << logn
    gpu.ntt_release_domain()
    gpu.ntt_init_domain(gpu.get_root_of_unity(n))
    try:
        def rnd():
            t = torch.randint(0, 1 << 31, (n * batch, 8), dtype=torch.int32, device="cuda")
            t[:, 7] &= 0x0FFFFFFF  # < r
            return t
        x, y = rnd(), rnd()
        cfg = B.NTTConfig.default()
        cfg.batch_size = batch
        cfg.are_inputs_on_device = cfg.are_outputs_on_device = True
        vcfg = B.VecOpsConfig.default()
        vcfg.is_a_on_device = vcfg.is_b_on_device = vcfg.is_result_on_device = True

        def vadd(a, b):
            o = torch.empty_like(a)
            B.check(gpu.dll.bn254_vector_add(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_uint64(n * batch), C.byref(vcfg),
                                             C.c_void_p(o.data_ptr())))
            return o

        def ntt(a, d, inplace=False):
            o = a if inplace else torch.empty_like(a)
            gpu.ntt(a.data_ptr(), d, cfg, out=o.data_ptr(), size=n)
            return o

        fx, fy = ntt(x, B.kForward), ntt(y, B.kForward)
        fxy = ntt(vadd(x, y), B.kForward)
        assert torch.equal(fxy, vadd(fx, fy))
        back = ntt(fx.clone(), B.kInverse, inplace=True)
        assert torch.equal(back, x)
        assert not torch.equal(fx, x)
        # known answer: the transform of the unit vector e_1 is the sequence of powers of the domain generator
        e1 = torch.zeros_like(x)
        for b in range(batch):
            e1[b * n + 1, 0] = 1
        fe = ntt(e1, B.kForward).cpu().numpy().view(np.uint32)
        w = from_words(gpu.get_root_of_unity(n))
        for k in [0, 1, 2, 12345, n // 2, n - 1]:
            for b in range(batch):
                assert from_words(fe[b * n + k]) == pow(w, k, R), (k, b)
    finally:
        gpu.ntt_release_domain()
