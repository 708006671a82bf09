#!/usr/bin/env python
"""bench.py - Groth16 prove latency at 3200k constraints (BASELINE.json's metric), warm ZKeyCache.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--constraints C] [--precompute F]
  python bench.py --impl reference ...        # the reference's own CPU implementation of the path

A "step" is one proof of the synthetic ComplexCircuit(C, C) instance (the reference's benchmark circuit,
benchmark/3200k/circuit.circom; valid .zkey/.wtns from a seeded known-toxic-waste setup, tools/synth.py -
snarkjs/circom are unavailable offline).  `value` = ms per proof with the witness already in HBM, device-timed
(CUDA events + barrier, max over ranks); `e2e` = the same through the public C ABI call with the witness in
pinned HOST memory and the proof read back to the host, inside the timed region.  N > 1 (torchrun): every rank
holds a contiguous shard of the five base-point sets, computes partial sums, one NCCL all_gather of 576 B per
rank, rank 0 folds + blinds (strong scaling: the job is one proof).
`roofline` describes the dominant kernel (G1 bucket accumulation); `cpu_baseline` the reference CPU library on
this box's host cores on a bounded sample (Measurement).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CPU_SAMPLE_CONSTRAINTS = 400_000  # bounded CPU sample (~6 s per proof on 16 cores); scaled linearly to the bench size


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


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


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_prove_ms(constraints, steps, warmup):
    """The reference's own CPU library (oracle/_ref) driven by the restated Rust host (oracle/groth16_ref.py)
    on ComplexCircuit(constraints): ms per proof, warm cache, all host threads the library uses."""
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from oracle import groth16_ref as G
    from oracle import ref_cpu
    from tools import synth
    ref = ref_cpu.ref()
    t0 = time.time()
    zkey, wtns, _vk = synth.make_complex_circuit(ref if constraints <= 2000 else _gpu_or_ref(pkg, ref), constraints)
    log(f"reference arm: synthetic {constraints}-constraint instance in {time.time() - t0:.1f}s")
    cache = G.ZKeyCacheRef(ref, zkey)
    times = []
    for i in range(warmup + steps):
        tm = {}
        G.prove(ref, pkg.bindings, zkey, wtns, 1, 1, cache=cache, timings=tm)
        if i >= warmup:
            times.append(tm["total_s"] * 1e3)
        log(f"reference arm: proof {i} took {tm['total_s'] * 1e3:.0f} ms (r1cs+ntt {tm['r1cs_ntt_s'] * 1e3:.0f}, msm {tm['msm_s'] * 1e3:.0f})")
    return sum(times) / len(times)


def _gpu_or_ref(pkg, ref):
    """Instance generation is setup, not measurement: use the GPU tool when a GPU is present (fast)."""
    try:
        lib = pkg.lib()
        lib.set_device("CUDA", int(os.environ.get("LOCAL_RANK", 0)))
        return lib
    except Exception:
        return ref


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    cores = os.cpu_count()
    c = min(args.constraints, CPU_SAMPLE_CONSTRAINTS)
    scale = args.constraints / c
    ms = cpu_reference_prove_ms(c, max(1, args.steps), min(args.warmup, 1))
    est = ms * scale
    out = {
        "impl": "reference", "metric": f"groth16_prove_latency_ms_{args.constraints // 1000}k", "value": est, "unit": "ms",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": est, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32x8 (254-bit modular integers)", "data": "synthetic",
        "config": {"workload": f"ComplexCircuit({args.constraints},{args.constraints}) Groth16 prove, warm cache",
                   "timing": "wall clock, host only"},
        "cpu_baseline": {"value": est, "unit": "ms", "cores": cores, "kind": "reference",
                         "sample": f"reference CPU library (ICICLE 3.8.0 frontend+CPU backend, g++ -O2, Taskflow stand-in) proving "
                                   f"ComplexCircuit({c}): {ms:.0f} ms/proof measured, scaled linearly x{scale:g} to {args.constraints} constraints (an upper bound: Pippenger grows as n/log n)"},
        "e2e": {"value": est, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ product arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--constraints", type=int, default=3_200_000)
    ap.add_argument("--precompute", type=int, default=int(os.environ.get("B200_PRECOMPUTE", "16")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--split-quotient", type=int, default=1, help="N>1: split the three quotient polynomials across ranks (0 = replicate)")
    ap.add_argument("--shard-skew", type=float, default=0.049,
                    help="N>1 with the quotient split: the polynomial owners get smaller witness-MSM shards "
                         "(b200_shard_range; = one polynomial's transform time / all witness MSMs' time; 0 = equal shards)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from tools import synth

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
    lib.dll.b200_pipe_peak.restype = C.c_double

    n = args.constraints
    t0 = time.time()
    zkey, wtns, _vk = synth.make_complex_circuit(lib, n, log=(lambda *a: log("setup:", *a)) if rank == 0 else None)
    if rank == 0:
        log(f"synthetic instance: {len(zkey) / 1e6:.0f} MB zkey in {time.time() - t0:.1f}s")
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

    qx = pkg.multi_gpu.QuotientExchange(cache, torch.device("cuda", local)) if (world > 1 and args.split_quotient) else None

    def step(witness_ptr):
        """one proof; returns the proof struct on rank 0"""
        if world == 1:
            proof, tm = cache.prove(witness_ptr, 1, 1, n_witness=nw)
            return proof, tm
        if qx is not None:  # quotient chain split across ranks: one scatter per polynomial over NVLink
            parts, tm = qx.commit(witness_ptr, n_witness=nw)
        else:               # quotient chain replicated on every rank
            parts, tm = cache.commit_partials(witness_ptr, n_witness=nw)
        plist = pkg.multi_gpu.all_gather_partials(parts, torch.device("cuda", local))  # one 576 B NCCL all_gather
        if rank != 0:
            return None, tm
        return cache.finish(plist, 1, 1), tm

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
    if world > 1 and qx is not None:
        # cross-check of the split path: the replicated-chain proof (same r = s = 1) must be identical
        parts_r, _ = cache.commit_partials(w_dev.data_ptr(), n_witness=nw)
        plist_r = pkg.multi_gpu.all_gather_partials(parts_r, torch.device("cuda", local))
        if rank == 0:
            assert pkg.proof_json(cache.finish(plist_r, 1, 1)) == pkg.proof_json(proof), "quotient-split proof != replicated proof"
    # device-side phase times of the last proof (CUDA events inside the library)
    phases = {k: round(getattr(tm, k), 3) for k in ("h2d_ms", "r1cs_ms", "ntt_ms", "msm_g1_ms", "msm_g2_ms", "total_ms")}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0
    if world == 1:
        assert pkg.proof_json(proof) == pkg.proof_json(proof_e2e), "device-witness and host-witness proofs differ"

    # ---- roofline of the dominant kernel: G1 bucket accumulation of the A MSM, timed alone with CUDA events
    imad_peak = lib.dll.b200_pipe_peak(1)  # independent IMAD.WIDE chains, measured now on this GPU
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    z_pts = lib.generate_affine_points(1 << 12)  # distinct valid points, tiled: accumulate cost does not depend on values
    n_msm = cache.n_vars
    pts = torch.from_numpy(np.tile(z_pts, ((n_msm >> 12) + 1, 1))[:n_msm].copy().view(np.int32)).cuda()
    cfg = pkg.MSMConfig.default()
    cfg.are_scalars_on_device = cfg.are_points_on_device = cfg.are_results_on_device = True
    res = torch.zeros(24, dtype=torch.int32, device="cuda")
    plan_c = None
    lib.dll.b200_profile_accumulate(1)
    acc_ms = []
    for i in range(6):
        lib.msm(w_dev.data_ptr(), pts.data_ptr(), cfg, results=res.data_ptr(), msm_size=n_msm)
        torch.cuda.synchronize()
        ms = lib.dll.b200_profile_accumulate(1)
        if i >= 2:
            acc_ms.append(ms)
    lib.dll.b200_profile_accumulate(0)
    acc = sum(acc_ms) / len(acc_ms)
    lg = (n_msm - 1).bit_length()
    c_bits = lg - 5
    windows = -(-256 // c_bits)
    alg_bytes = n_msm * windows * (4 + 64)          # one index + one affine point per (scalar, window)
    alg_mac = n_msm * windows * 10 * 136            # 10 field mults per mixed add, 136 32x32 multiply-adds each
    roofline = {
        "kernel": "msm_accumulate_kernel<Fq>", "bound": "hbm", "achieved": alg_bytes / (acc * 1e-3) / 1e9, "peak": hbm_peak,
        "unit": "GB/s", "frac": alg_bytes / (acc * 1e-3) / 1e9 / hbm_peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of this very launch shape from the committed ncu capture
        # (profiles/r01_ncu_msm_accumulate_g1_3200k.md); other sizes were not captured
        "traffic": 4.426e9 if n_msm == 3200002 else None, "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
        "launch_ms": acc, "units_per_launch": f"{n_msm} scalars x {windows} windows (c={c_bits})",
        "int_pipe": {"achieved_tmac_s": alg_mac / (acc * 1e-3) / 1e12, "peak_tmac_s": imad_peak / 1e12,
                     "frac": alg_mac / (acc * 1e-3) / imad_peak,
                     "note": "binding resource: 32-bit integer multiply-add pipe (IMAD.WIDE), peak measured in this run"},
    }
    try:
        # the practical ceiling of that pipe for this arithmetic: dependent 8x32-bit CIOS Montgomery products (the carry-in
        # form of IMAD.WIDE occupies the pipe ~2.5x longer than the independent one), measured now on this GPU
        mul_peak = lib.dll.b200_pipe_peak(6)
        if mul_peak > 0:
            mults = n_msm * windows * 10 / (acc * 1e-3)
            roofline["field_mul"] = {"achieved_gmul_s": mults / 1e9, "peak_gmul_s": mul_peak / 1e9, "frac": mults / mul_peak,
                                     "pipe_active_ncu": 0.895,
                                     "note": "254-bit Montgomery products/s vs the multiplier microbenchmark (b200_pipe_peak(6)); "
                                             "pipe_active_ncu = sm__pipe_fmaheavy_cycles_active of this kernel in "
                                             "profiles/r01_ncu_msm_accumulate_g1_3200k.md"}
    except Exception as exc:  # the extra view must never cost the bench line
        log(f"field_mul roofline view skipped: {exc}")
    # secondary headline: standalone G1 MSM throughput with resident inputs
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
    msm_mpts = n_msm / (sum(t_ms) / len(t_ms)) / 1e3

    cpu_baseline = None
    if not args.no_cpu_baseline:
        c = min(n, CPU_SAMPLE_CONSTRAINTS)
        ms = cpu_reference_prove_ms(c, 2, 1)
        cpu_baseline = {"value": ms * n / c, "unit": "ms", "cores": os.cpu_count(), "kind": "reference",
                        "sample": f"reference CPU library proving ComplexCircuit({c}): {ms:.0f} ms/proof measured (2 proofs after 1 warm-up), "
                                  f"scaled linearly x{n / c:g} (an upper bound: Pippenger grows as n/log n)"}

    out = {
        "metric": f"groth16_prove_latency_ms_{n // 1000}k", "value": ms_dev, "unit": "ms", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32x8 (254-bit modular integers)", "data": "synthetic",
        "config": {"workload": f"ComplexCircuit({n},{n}) Groth16 prove, warm ZKeyCache", "n_vars": cache.n_vars,
                   "domain_size": cache.domain_size, "precompute_factor": args.precompute, "parallelism": f"msm-shard{world}" + ("+quotient-split" if qx is not None else "") + (f"+shard-skew{skew:g}" if skew > 0 else ""),
                   "l2": "256 MiB flush between timed iterations", "timing": "host clock around the synchronous C-ABI call + cuda sync + barrier, max over ranks"},
        "e2e": {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": nw * 32, "d2h_bytes_per_step": 576 if world == 1 else 576 * world,
                "note": "h2d bytes are per rank that evaluates R1CS rows (all ranks when the quotient chain is replicated; the 3 polynomial owners when it is split, the others upload only their 1/N witness slice)"},
        "gpu_launches": int(launches), "clocks": clocks, "phases_ms": phases, "roofline": roofline,
        "cpu_baseline": cpu_baseline, "extras": {"msm_g1_mpoints_s": msm_mpts, "msm_g1_size": n_msm,
                                               "device_cache_bytes": cache.device_bytes, "cache_build_s": round(t_cache, 3)},
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
