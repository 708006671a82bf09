#!/usr/bin/env python
"""bench.py - Groth16 prove latency at 3200k constraints (BASELINE.json's metric), warm ZKeyCache.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--constraints C] [--precompute F]
  python bench.py --impl reference ...        # the reference's own CPU implementation of the path, same instance
  python bench.py --impl reference-cuda ...   # the reference's own CUDA backend rebuilt for sm_100a (oracle/_ref_cuda), same GPU
  python bench.py [--no-sweep]                # the standalone BN254 MSM / NTT sweeps (BASELINE configs[4]) are folded into `extras` by default

A "step" is one proof of the synthetic ComplexCircuit(C, C) instance (the reference's benchmark circuit,
benchmark/3200k/circuit.circom; valid .zkey/.wtns from a seeded known-toxic-waste setup, tools/synth.py -
snarkjs/circom are unavailable offline) with fixed NON-trivial blinding factors r, s.  `value` = ms per proof with the
witness already in HBM (device pointer), timed around the synchronous C-ABI call between device synchronisations
(barrier + max over ranks); `e2e` = the same call with the witness in pinned HOST memory and the proof read back to the
host, copies inside the timed region.  The last timed proof is verified outside the timed region with the library's
own pairing check AND the reference's (`oracle/_ref` bn254_pairing): "verified".  N > 1 (torchrun): every rank holds its
pieces of the five base-point tables (b200_shard_plan) and calls b200_groth16_prove_sharded - quotient slices, the witness
all-gather and the 576-byte gather of partial sums travel over NCCL inside the library; rank 0 folds + blinds (strong
scaling: the job is one proof).  `extras.sweep` carries the standalone MSM / NTT sweeps at the same N.
`roofline` describes the dominant kernel - the G1 bucket accumulation launch the proof really issues (three tables sharing
one sort), timed ISOLATED in an extra profiled proof after the timed region - against the integer multiply pipe, whose
peak is measured in the same run (research/pipes2.cu).  `cpu_baseline`: the reference CPU library on this box's host
cores on a bounded sample (Measurement).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
CPU_SAMPLE_CONSTRAINTS = 400_000  # bounded CPU sample of the product arm's cpu_baseline (~5 s per proof on 16 cores)
REF_WALL_BUDGET_S = float(os.environ.get("B200_REF_BUDGET_S", "900"))  # reference arm: proofs at the real size until this
INSTANCE_DIR = os.environ.get("B200_BENCH_CACHE", os.path.join(ROOT, ".bench_cache"))
# fixed, non-trivial blinding factors (the epilogue's host scalar multiplications are inside the timed region)
R_BLIND = int.from_bytes(hashlib.sha256(b"icicle-snark-b200 bench r").digest(), "big") % R_MOD
S_BLIND = int.from_bytes(hashlib.sha256(b"icicle-snark-b200 bench s").digest(), "big") % R_MOD


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# stdout carries exactly ONE JSON line: native libraries (NCCL's version banner, ...) write to fd 1 directly, so fd 1 is
# pointed at stderr for the whole run and the result line goes to the saved descriptor
_RESULT_FD = None


def guard_stdout():
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, line)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 8 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ instance
def instance_paths(n):
    d = os.path.join(INSTANCE_DIR, f"complex_{n}")
    return d, os.path.join(d, "circuit.zkey"), os.path.join(d, "witness.wtns"), os.path.join(d, "vk.npz")


def load_instance(n):
    """(zkey bytes, wtns bytes, vk dict) of ComplexCircuit(n, n) from the on-disk cache, or None."""
    d, zk, wt, vkp = instance_paths(n)
    if not (os.path.exists(zk) and os.path.exists(wt) and os.path.exists(vkp) and os.path.exists(os.path.join(d, "done"))):
        return None
    vkz = np.load(vkp)
    vk = {k: vkz[k] for k in ("alpha1", "beta2", "gamma2", "delta2", "ic")}
    vk["n_public"] = int(vkz["n_public"])
    with open(zk, "rb") as f:
        zkey = f.read()
    with open(wt, "rb") as f:
        wtns = f.read()
    return zkey, wtns, vk


def save_instance(n, zkey, wtns, vk):
    d, zk, wt, vkp = instance_paths(n)
    try:
        os.makedirs(d, exist_ok=True)
        with open(zk, "wb") as f:
            f.write(zkey)
        with open(wt, "wb") as f:
            f.write(wtns)
        np.savez(vkp, n_public=vk["n_public"], **{k: vk[k] for k in ("alpha1", "beta2", "gamma2", "delta2", "ic")})
        from tools import synth
        with open(os.path.join(d, "verification_key.json"), "w") as f:
            f.write(synth.vk_json(vk))
        with open(os.path.join(d, "done"), "w") as f:
            f.write("ok\n")
    except OSError as exc:  # a read-only tree only costs the cache
        log(f"instance cache not written: {exc}")


def make_instance(backend, n, say=None):
    from tools import synth
    t0 = time.time()
    zkey, wtns, vk = synth.make_complex_circuit(backend, n, log=say)
    log(f"synthetic ComplexCircuit({n}) instance: {len(zkey) / 1e6:.0f} MB zkey in {time.time() - t0:.1f}s")
    return zkey, wtns, vk


# ------------------------------------------------------------------------------------------------ reference arm
def reference_instance(n):
    """The reference arm's instance.  The reference process itself never loads libicicle_b200.so: a large instance that
    is not cached yet is generated by a SEPARATE process (tools/synth.py, GPU fixed-base tool) and read back from disk;
    tiny ones are generated with the reference CPU library."""
    got = load_instance(n)
    if got is not None:
        return got
    if n <= 2000:
        import __graft_entry__ as ge
        ge.load_package()  # python structs only; the product .so is not loaded here
        from oracle import ref_cpu
        inst = make_instance(ref_cpu.ref(), n)
        save_instance(n, *inst)
        return inst
    t0 = time.time()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "synth.py"), "--constraints", str(n), "--out",
                        instance_paths(n)[0]], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("instance generation subprocess failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    log(f"reference arm: instance generated by a separate process in {time.time() - t0:.1f}s")
    got = load_instance(n)
    if got is None:
        raise RuntimeError("instance generation subprocess left no files")
    return got


def cpu_reference_proofs(zkey, wtns, vk, steps, warmup, budget_s):
    """The reference's own CPU library (oracle/_ref) driven by the restated Rust host (oracle/groth16_ref.py): per-proof
    ms (warm cache, all host threads the library uses).  Stops early when `budget_s` of wall clock would be exceeded."""
    import __graft_entry__ as ge
    pkg = ge.load_package()  # python structs only; the product .so is not loaded here
    from oracle import groth16_ref as G
    from oracle import ref_cpu
    ref = ref_cpu.ref()
    t_start = time.time()
    cache = G.ZKeyCacheRef(ref, zkey)
    log(f"reference arm: cache built in {time.time() - t_start:.1f}s")
    times, proof, public = [], None, None
    for i in range(warmup + steps):
        tm = {}
        proof, public = G.prove(ref, pkg.bindings, zkey, wtns, R_BLIND, S_BLIND, cache=cache, timings=tm)
        if i >= warmup:
            times.append(tm["total_s"] * 1e3)
        log(f"reference arm: proof {i} took {tm['total_s'] * 1e3:.0f} ms (r1cs+ntt {tm['r1cs_ntt_s'] * 1e3:.0f}, msm {tm['msm_s'] * 1e3:.0f})")
        elapsed, per = time.time() - t_start, tm["total_s"]
        if times and elapsed + per * 1.15 > budget_s and i + 1 < warmup + steps:
            log(f"reference arm: wall budget {budget_s:.0f}s reached after {len(times)} timed proofs")
            break
    ok = bool(G.verify(ref, proof, public, vk))
    return times, ok, proof


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    cores = os.cpu_count()
    n = args.constraints
    zkey, wtns, vk = reference_instance(n)
    warm = min(args.warmup, 1)  # the cache is warm after one proof; every further warm-up proof costs ~0.5 min of CPU
    times, ok, _ = cpu_reference_proofs(zkey, wtns, vk, max(1, args.steps), warm, REF_WALL_BUDGET_S)
    ms = sum(times) / len(times)
    out = {
        "impl": "reference", "metric": f"groth16_prove_latency_ms_{n // 1000}k", "value": ms, "unit": "ms",
        "n_gpus": args.gpus, "steps": len(times), "warmup": warm, "ms_per_step": ms, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32x8 (254-bit modular integers)", "data": "synthetic",
        "config": {"workload": f"ComplexCircuit({n},{n}) Groth16 prove, warm cache", "timing": "wall clock, host only",
                   "requested_steps": args.steps, "requested_warmup": args.warmup,
                   "blinding": "fixed non-trivial r, s (same as the product arm)"},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "reference",
                         "sample": f"reference CPU library (ICICLE 3.8.0 frontend + CPU backend, g++ -O2, Taskflow stand-in) proving the "
                                   f"real ComplexCircuit({n}) instance: {len(times)} timed proofs after {warm} warm-up, "
                                   f"min {min(times):.0f} / max {max(times):.0f} ms (no extrapolation)"},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "verified": ok, "gpu_launches": 0,
    }
    emit(out)
    return 0


def run_reference_cuda(args):
    """Second baseline (SURVEY 8c / BASELINE.md 3): the reference's CUDA backend, recompiled for sm_100a from its own
    sources (oracle/Makefile.ref_cuda), driven through the reference frontend by the restated Rust host with the Rust
    code's residency (oracle/groth16_ref_cuda.py).  Wall clock per proof, warm cache, witness from host memory."""
    if int(os.environ.get("RANK", 0)) != 0:
        return 0
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from oracle import groth16_ref as G
    from oracle import groth16_ref_cuda as GC
    n = args.constraints
    if not GC.available():
        emit({"impl": "reference-cuda", "unavailable": "oracle/_ref_cuda is not built (needs /root/reference at build time)"})
        return 0
    zkey, wtns, vk = reference_instance(n)
    ref = GC.ref_cuda(0)
    t0 = time.time()
    cache = GC.ZKeyCacheCuda(ref, pkg.bindings, zkey)
    ref.device_synchronize()
    t_cache = time.time() - t0
    del zkey
    times, parts, proof, public = [], [], None, None
    warm = max(1, min(args.warmup, 2))
    for i in range(warm + args.steps):
        tm = {}
        proof, public = GC.prove(ref, pkg.bindings, wtns, R_BLIND, S_BLIND, cache, timings=tm)
        if i >= warm:
            times.append(tm["total_s"] * 1e3)
            parts.append((tm["r1cs_ntt_s"] * 1e3, tm["msm_s"] * 1e3))
        log(f"reference-cuda arm: proof {i} took {tm['total_s'] * 1e3:.0f} ms (r1cs+ntt {tm['r1cs_ntt_s'] * 1e3:.0f}, msm {tm['msm_s'] * 1e3:.0f})")
    ref.set_device("CPU", 0)
    ok = bool(G.verify(ref, proof, public, vk))
    ms = sum(times) / len(times)
    out = {
        "impl": "reference-cuda", "metric": f"groth16_prove_latency_ms_{n // 1000}k", "value": ms, "unit": "ms", "n_gpus": 1,
        "steps": len(times), "warmup": warm, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32x8 (254-bit modular integers)", "data": "synthetic",
        "config": {"workload": f"ComplexCircuit({n},{n}) Groth16 prove, warm cache", "timing": "wall clock around the restated Rust host; copies included as in the Rust code",
                   "backend": "ICICLE 3.8.0 CUDA backend, nvcc -O3 sm_100a, unmodified sources", "blinding": "fixed non-trivial r, s (same as the product arm)",
                   "host_glue": "the Rust host's gather / scatter / copies restated in numpy (single thread): the r1cs+ntt share is an upper bound for the Rust+rayon host, the msm share is the reference's own GPU code alone"},
        "phases_ms": {"r1cs_ntt_ms": round(sum(x[0] for x in parts) / len(parts), 1), "msm_ms": round(sum(x[1] for x in parts) / len(parts), 1)},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None},
        "verified": ok, "cache_build_s": round(t_cache, 2), "proof_json_sha256": hashlib.sha256(G.proof_json(proof).encode()).hexdigest(),
        "min_ms": min(times), "max_ms": max(times),
    }
    emit(out)
    return 0


# ------------------------------------------------------------------------------------------------ product arm
def plan_info(lib, n, factor, g2=False, c=0):
    out = (C.c_int32 * 8)()
    rc = lib.dll.b200_msm_plan_info(C.c_int(n), C.c_int(c), C.c_int(254), C.c_int(factor), C.c_int(int(g2)), out, None)
    if rc != 0:
        raise RuntimeError(f"b200_msm_plan_info failed: {rc}")
    return dict(zip(("c", "windows", "factor", "sets", "bpw", "nbuckets", "item_cap", "n"), list(out)))


def profile_records(lib):
    cap = 64
    buf = (C.c_int32 * (9 * cap))()
    n = lib.dll.b200_profile_records(buf, C.c_int(cap))
    recs = []
    for i in range(min(n, cap)):
        v = list(buf[9 * i:9 * i + 9])
        ms = np.array([v[8]], dtype=np.int32).view(np.float32)[0]
        recs.append(dict(g2=v[0], nsel=v[1], n=v[2], windows=v[3], c=v[4], factor=v[5], nbuckets=v[6], batched=v[7], ms=float(ms)))
    return recs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--constraints", type=int, default=3_200_000)
    ap.add_argument("--precompute", type=int, default=int(os.environ.get("B200_PRECOMPUTE", "16")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the pairing checks of the last timed proof (debug only)")
    ap.add_argument("--sweep", action="store_true", help="(default) add the standalone MSM / NTT sweeps (configs[4]) to `extras`")
    ap.add_argument("--no-sweep", action="store_true", help="skip the standalone MSM / NTT sweeps")
    ap.add_argument("--split-quotient", type=int, default=1, help="N>1: split the three quotient polynomials across ranks (0 = replicate)")
    ap.add_argument("--exchange", default="lib", choices=["lib", "torch"],
                    help="N>1: who moves the data: 'lib' = b200_groth16_prove_sharded (NCCL inside the C library, no host in the "
                         "exchange), 'torch' = commit_begin / torch.distributed scatter + all_gather / commit_end")
    ap.add_argument("--shard-skew", type=float, default=0.049,
                    help="N>1 with the quotient split: the polynomial owners get smaller witness-MSM shards "
                         "(b200_shard_range; = one polynomial's transform time / all witness MSMs' time; 0 = equal shards)")
    args = ap.parse_args()
    guard_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "reference-cuda":
        return run_reference_cuda(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package()

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = pkg.lib()
    lib.set_device("CUDA", local)
    lib.dll.b200_launch_count.restype = C.c_ulonglong
    lib.dll.b200_profile_accumulate.restype = C.c_float

    n = args.constraints
    inst = load_instance(n)
    if inst is None:
        inst = make_instance(lib, n, (lambda *a: log("setup:", *a)) if rank == 0 else None)
        if rank == 0:
            save_instance(n, *inst)
    zkey, wtns, vk = inst
    t_cache = time.time()
    skew = args.shard_skew if (world > 1 and args.split_quotient) else 0.0
    if skew > 0:
        os.environ["B200_SHARD_SKEW"] = repr(skew)  # read by the library when it cuts the cache's witness shards
    else:
        os.environ.pop("B200_SHARD_SKEW", None)
    cache = pkg.ZKeyCache(lib, zkey, precompute=args.precompute, rank=rank, world=world)
    lib.device_synchronize()
    t_cache = time.time() - t_cache  # cold path: parse + H2D + precompute tables + CSR + coset powers + twiddles
    del zkey
    # witness: section 2 of the .wtns -> pinned host buffer (e2e) and a device copy (value)
    nw = cache.n_vars
    w_np = np.frombuffer(wtns, dtype=np.uint32, count=nw * 8, offset=len(wtns) - nw * 32).reshape(nw, 8)
    w_pinned = torch.from_numpy(w_np.copy().view(np.int32)).pin_memory()
    w_dev = w_pinned.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    use_lib = world > 1 and args.split_quotient and args.exchange == "lib"
    comm = None
    if use_lib:
        try:
            comm = pkg.multi_gpu.LibComm.from_torch(lib)
        except RuntimeError as exc:  # no libnccl.so.2 for the library's dlopen: torch.distributed moves the slices instead
            log(f"in-library communicator unavailable ({exc}); falling back to the torch.distributed exchange")
            use_lib = False
    qx = pkg.multi_gpu.QuotientExchange(cache, torch.device("cuda", local)) if (world > 1 and args.split_quotient and not use_lib) else None

    def step(witness_ptr):
        """one proof; returns the proof struct on rank 0"""
        if world == 1:
            proof, tm = cache.prove(witness_ptr, R_BLIND, S_BLIND, n_witness=nw)
            return proof, tm
        if comm is not None:  # everything inside the library: grouped send/recv of the slices + one 576 B all-gather
            return cache.prove_sharded(comm, witness_ptr, R_BLIND, S_BLIND, n_witness=nw)
        if qx is not None:  # quotient chain split across ranks: one scatter per polynomial over NVLink
            parts, tm = qx.commit(witness_ptr, n_witness=nw)
        else:               # quotient chain replicated on every rank
            parts, tm = cache.commit_partials(witness_ptr, n_witness=nw)
        plist = pkg.multi_gpu.all_gather_partials(parts, torch.device("cuda", local))  # one 576 B NCCL all_gather
        if rank != 0:
            return None, tm
        return cache.finish(plist, R_BLIND, S_BLIND), tm

    def timed(witness_ptr, steps, warmup):
        for _ in range(warmup):
            step(witness_ptr)
        per = []
        last = None
        for _ in range(steps):
            flush.fill_(1)  # evict L2 between timed iterations
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t = time.perf_counter()
            last = step(witness_ptr)  # returns after the proof is on the host (stream sync inside the C ABI)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            per.append((time.perf_counter() - t) * 1e3)
        return per, last

    sampler = ClockSampler(local)
    launches0 = lib.dll.b200_launch_count()
    sampler.start()
    # value: witness resident in HBM (device pointer; the C ABI's copy degenerates to a device-to-device move)
    per_dev, (proof, tm) = timed(w_dev.data_ptr(), args.steps, args.warmup)
    clocks = sampler.stop()
    launches = (lib.dll.b200_launch_count() - launches0) // (args.steps + args.warmup)
    # e2e: witness in pinned host memory, proof (3 affine points) back on the host
    per_e2e, (proof_e2e, _) = timed(w_pinned.data_ptr(), args.steps, 1)

    def agg(per):
        t = torch.tensor([sum(per) / len(per)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_dev, ms_e2e = agg(per_dev), agg(per_e2e)
    # the same call with a PAGEABLE host witness (what a Rust Vec or an mmap'd .wtns is): staged through pinned memory by the library
    w_pageable = np.ascontiguousarray(w_np).copy()
    per_pg, (proof_pg, _) = timed(w_pageable.ctypes.data, max(3, args.steps // 2), 1)
    ms_e2e_pageable = agg(per_pg)
    if world > 1 and (qx is not None or comm is not None):
        # cross-check of the split path: the replicated-chain proof (same r, s) must be identical
        parts_r, _ = cache.commit_partials(w_dev.data_ptr(), n_witness=nw)
        plist_r = pkg.multi_gpu.all_gather_partials(parts_r, torch.device("cuda", local))
        if rank == 0:
            assert pkg.proof_json(cache.finish(plist_r, R_BLIND, S_BLIND)) == pkg.proof_json(proof), "quotient-split proof != replicated proof"
    # device-side phase times of the last proof (CUDA events inside the library)
    phases = {k: round(getattr(tm, k), 3) for k in ("h2d_ms", "r1cs_ms", "ntt_ms", "msm_g1_ms", "msm_g2_ms", "total_ms")}

    # standalone sweeps at N > 1 are collective (sharded MSM through the library's communicator): every rank runs them
    # here, before the other ranks retire; at N = 1 they run with the other extras below
    sweep_results = None
    if world > 1 and not args.no_sweep:
        from tools import sweep
        sweep_results = sweep.run(lib, pkg, log=log if rank == 0 else None)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0
    assert pkg.proof_json(proof) == pkg.proof_json(proof_e2e), "device-witness and host-witness proofs differ"

    # ---- parity of the headline config: the last timed proof under both pairing checks (outside the timed region)
    verified = None
    if not args.no_verify:
        public = [int.from_bytes(w_np[i].tobytes(), "little") for i in range(1, cache.n_public + 1)]
        ok_own = bool(pkg.groth16_verify_points(lib, proof, public, vk))
        ok_ref = None
        try:
            from oracle import groth16_ref as G
            from oracle import ref_cpu
            if ref_cpu.available():
                ok_ref = bool(G.verify(ref_cpu.ref(), pkg.proof_to_dict(proof), public, vk))
        except Exception as exc:
            log(f"reference pairing check unavailable: {exc}")
        verified = ok_own and (ok_ref is not False)
        log(f"proof check: library pairing {ok_own}, reference pairing (oracle/_ref) {ok_ref}")
        if not verified:
            raise SystemExit("bench.py: the timed proof does NOT verify - no number is reported")
        verified = {"ok": True, "library_pairing": ok_own, "reference_pairing": ok_ref, "r_s": "fixed non-trivial"}

    # ---- roofline of the dominant kernel: the G1 accumulation launch the proof issues, timed isolated (profiling mode)
    roofline, acc_all = None, None
    if world == 1:
        tools = pkg.tools_lib()
        imad_peak = tools.b200_imad_wide_peak()      # wide (32x32+64) MAC/s, measured now on this GPU
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
        lib.dll.b200_profile_accumulate(1)
        for _ in range(2):                            # two profiled proofs; the second one's records are used
            cache.prove(w_dev.data_ptr(), R_BLIND, S_BLIND, n_witness=nw)
        recs = profile_records(lib)
        lib.dll.b200_profile_accumulate(0)
        recs = recs[len(recs) // 2:]
        acc_all = recs
        g1 = max((r for r in recs if not r["g2"]), key=lambda r: r["nsel"] * r["n"])
        adds = g1["nsel"] * g1["n"] * g1["windows"]
        ppa = 10.0                                    # field products per mixed XYZZ add (SURVEY 8d unit)
        alg_mac, alg_bytes = adds * ppa * 136, adds * (4 + 64)
        t = g1["ms"] * 1e-3
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic = tj.get(f"g1_nsel{g1['nsel']}_n{g1['n']}_w{g1['windows']}_c{g1['c']}_ba{g1['batched']}")
        except (OSError, ValueError):
            pass
        total_mac = sum(r["nsel"] * r["n"] * r["windows"] * (30 if r["g2"] else 10) * 136 for r in recs)
        roofline = {
            "kernel": "msm_accumulate (G1, the proof's own launch: %d tables x %d points x %d windows, c=%d%s)" % (
                g1["nsel"], g1["n"], g1["windows"], g1["c"], f", batched-affine rounds={g1['batched']}" if g1["batched"] else ""),
            "bound": "int_pipe", "achieved": alg_mac / t / 1e12, "peak": imad_peak / 1e12, "unit": "T wide-MAC/s",
            "frac": alg_mac / t / imad_peak, "traffic": traffic, "launch_ms": g1["ms"],
            "units_per_launch": f"{adds} bucket additions = SURVEY 8d: 10 products x 136 multiply-adds each",
            "peak_source": "IMAD.WIDE.U32 throughput measured in this run (research/pipes2.cu, T1); 32-bit IMAD issues at twice "
                           "that rate (the CUDA table's 64/clk/SM), a 32x32+64 multiply-add takes two passes",
            "timing": "CUDA events on the launch's stream with device-wide sync on both sides (profiling mode, extra proof after the timed region)",
            "hbm": {"achieved_gbs": alg_bytes / t / 1e9, "peak_gbs": hbm_peak, "frac": alg_bytes / t / 1e9 / hbm_peak,
                    "algorithmic_bytes": alg_bytes, "peak_source": peak_src},
            "proof": {"accumulate_mac_all_launches": total_mac, "frac_of_int_pipe_over_ms_per_step": total_mac / (ms_dev * 1e-3) / imad_peak,
                      "launches": [{k: r[k] for k in ("g2", "nsel", "n", "windows", "c", "batched", "ms")} for r in recs]},
        }

    assert pkg.proof_json(proof_pg) == pkg.proof_json(proof), "pageable-witness proof differs"
    extras = {"device_cache_bytes": cache.device_bytes, "cache_build_s": round(t_cache, 3), "e2e_pageable_witness_ms": ms_e2e_pageable}
    # ---- the drop-in call itself: b200_groth16_prove_files (pageable mmap'd witness, JSON written), warm cache
    if world == 1:
        try:
            d, zk_path, wt_path, _ = instance_paths(n)
            if os.path.exists(zk_path):
                os.environ["B200_PRECOMPUTE"] = str(args.precompute)
                out_p, out_pub = os.path.join(d, "proof.json"), os.path.join(d, "public.json")
                fn = lib.dll.b200_groth16_prove_files
                t0 = time.perf_counter()
                rc = fn(os.fsencode(wt_path), os.fsencode(zk_path), os.fsencode(out_p), os.fsencode(out_pub), b"CUDA")
                t_first = time.perf_counter() - t0
                ts = []
                for _ in range(3):
                    flush.fill_(1)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    rc |= fn(os.fsencode(wt_path), os.fsencode(zk_path), os.fsencode(out_p), os.fsencode(out_pub), b"CUDA")
                    ts.append((time.perf_counter() - t0) * 1e3)
                if rc == 0:
                    pkg.groth16_verify(out_p, out_pub, os.path.join(d, "verification_key.json"), lib=lib)
                    extras["prove_files_ms"] = sum(ts) / len(ts)
                    extras["prove_files_first_call_s"] = round(t_first, 3)
                    extras["prove_files_note"] = ("b200_groth16_prove_files (the reference's groth16_prove signature): mmap'd pageable .wtns, "
                                                  "proof.json + public.json written, random r/s; first call builds the process-wide cache")
        except Exception as exc:
            log(f"file-level e2e skipped: {exc}")

    # secondary headline: standalone G1 MSM throughput with resident inputs (metric ii)
    z_pts = lib.generate_affine_points(1 << 12)  # distinct valid points, tiled: accumulate cost does not depend on values
    n_msm = cache.n_vars
    pts = torch.from_numpy(np.tile(z_pts, ((n_msm >> 12) + 1, 1))[:n_msm].copy().view(np.int32)).cuda()
    cfg = pkg.MSMConfig.default()
    cfg.are_scalars_on_device = cfg.are_points_on_device = cfg.are_results_on_device = True
    res = torch.zeros(24, dtype=torch.int32, device="cuda")
    t_ms = []
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(1)
        e0.record()
        lib.msm(w_dev.data_ptr(), pts.data_ptr(), cfg, results=res.data_ptr(), msm_size=n_msm)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            t_ms.append(e0.elapsed_time(e1))
    extras["msm_g1_mpoints_s"] = n_msm / (sum(t_ms) / len(t_ms)) / 1e3
    extras["msm_g1_size"] = n_msm
    extras["msm_g1_plan"] = plan_info(lib, n_msm, 1)
    if sweep_results is not None:
        extras["sweep"] = sweep_results
    elif not args.no_sweep:
        from tools import sweep
        extras["sweep"] = sweep.run(lib, pkg, log=log)

    cpu_baseline = None
    if not args.no_cpu_baseline:
        c = min(n, CPU_SAMPLE_CONSTRAINTS)
        inst_c = (load_instance(c) or make_instance(lib, c)) if c != n else (None, wtns, vk)
        if c != n and load_instance(c) is None:
            save_instance(c, *inst_c)
        zk_c = inst_c[0] if c != n else open(instance_paths(n)[1], "rb").read()
        times, ok, _ = cpu_reference_proofs(zk_c, inst_c[1], inst_c[2], 2, 1, 120.0)
        ms = sum(times) / len(times)
        cpu_baseline = {"value": ms * n / c, "unit": "ms", "cores": os.cpu_count(), "kind": "reference",
                        "sample": f"reference CPU library proving ComplexCircuit({c}): {ms:.0f} ms/proof measured ({len(times)} proofs after 1 warm-up, "
                                  f"verified={ok}), scaled linearly x{n / c:g} (an upper bound: Pippenger grows as n/log n); "
                                  f"`bench.py --impl reference` proves the real size"}

    out = {
        "metric": f"groth16_prove_latency_ms_{n // 1000}k", "value": ms_dev, "unit": "ms", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32x8 (254-bit modular integers)", "data": "synthetic",
        "config": {"workload": f"ComplexCircuit({n},{n}) Groth16 prove, warm ZKeyCache", "n_vars": cache.n_vars,
                   "domain_size": cache.domain_size, "precompute_factor": args.precompute, "parallelism": f"msm-shard{world}" + (("+line-plan" if lib.dll.b200_shard_plan_mode(world) == 1 else "+uniform-plan") if world > 1 else "") + ("+quotient-split" if (qx is not None or comm is not None) else "") + ("+in-library-nccl" if comm is not None else "") + (f"+shard-skew{skew:g}" if skew > 0 else ""),
                   "blinding": "fixed non-trivial r, s", "l2": "256 MiB flush between timed iterations",
                   "timing": "host clock around the synchronous C-ABI call + cuda sync + barrier, max over ranks"},
        "e2e": {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": nw * 32, "d2h_bytes_per_step": 576 if world == 1 else 576 * world,
                "note": "h2d bytes are per rank that evaluates R1CS rows (all ranks when the quotient chain is replicated; the 3 polynomial owners when it is split, the others upload only their 1/N witness slice)"},
        "verified": verified, "proof_json_sha256": hashlib.sha256(pkg.proof_json(proof).encode()).hexdigest(), "gpu_launches": int(launches), "clocks": clocks, "phases_ms": phases, "roofline": roofline,
        "cpu_baseline": cpu_baseline, "extras": extras,
    }
    emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
