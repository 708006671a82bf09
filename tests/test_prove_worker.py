"""The `prove` worker binary speaks the reference's stdin protocol (/root/reference/src/main.rs:121-186):
prompt, COMMAND_EMPTY / COMMAND_EXIT / COMMAND_COMPLETED sentinels, defaults, --device handling."""
import os
import subprocess

import pytest

import icicle_snark_b200 as pkg

BIN = os.path.join(pkg.PKG_DIR, "bin", "prove")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run(stdin, env=None, cwd=None):
    if not os.path.exists(BIN):
        pkg.build()
    return subprocess.run([BIN], input=stdin, capture_output=True, text=True, timeout=300, env=env, cwd=cwd)


def test_protocol_sentinels_without_gpu():
    r = run("\nbogus\nprove --device CPU\nverify\nexit\n")
    out = r.stdout
    assert r.returncode == 0
    assert out.count("COMMAND_COMPLETED") == 4  # empty, prove(CPU refused), verify, exit
    assert "COMMAND_EMPTY" in out and "COMMAND_EXIT" in out and "Usage: prove [OPTIONS]" in out
    assert "COMMAND_FAILED" in out and "no CPU backend" in r.stderr  # CPU device is refused, not emulated
    assert out.rstrip().endswith("Exiting CLI worker...")


def test_worker_verifies_golden_proofs_on_the_host(tmp_path):
    """`verify --proof --public --vk` (main.rs:83-113,169-178): host pairing, as in the reference; a proof that does
    not verify is reported (the Rust worker panics there) and the worker keeps serving."""
    base = os.path.join(GOLD, "complex_100")
    good = f"verify --proof {base}.proof_rs.json --public {base}.public.json --vk {base}.vk.json\n"
    bad = f"verify --proof {GOLD}/complex_6.proof_rs.json --public {base}.public.json --vk {base}.vk.json\n"
    r = run(good + "exit\n")
    assert r.returncode == 0 and r.stdout.count("COMMAND_COMPLETED") == 2 and "COMMAND_FAILED" not in r.stdout
    r = run(bad + good + "verify --proof\nexit\n")
    assert r.returncode == 0 and r.stdout.count("COMMAND_COMPLETED") == 3 and r.stdout.count("COMMAND_FAILED") == 1
    assert "Verification failed" in r.stderr and "Usage: prove [OPTIONS]" in r.stdout


@pytest.mark.gpu
def test_worker_proves_golden_instance(tmp_path):
    base = os.path.join(GOLD, "complex_100")
    proof, public = tmp_path / "proof.json", tmp_path / "public.json"
    cmd = f"prove --witness {base}.wtns --zkey {base}.zkey --proof {proof} --public {public} --device CUDA\n"
    env = dict(os.environ, B200_NO_RANDOMNESS="1")
    r = run(cmd + cmd + "exit\n", env=env)
    assert r.returncode == 0 and r.stdout.count("COMMAND_COMPLETED") == 3 and "proof took:" in r.stdout
    assert open(proof).read() == open(base + ".proof_r1s1.json").read()
    assert open(public).read() == open(base + ".public.json").read()
    # prove -> verify in one worker session, on the files the worker itself wrote
    r = run(cmd + f"verify --proof {proof} --public {public} --vk {base}.vk.json\nexit\n", env=env)
    assert r.returncode == 0 and r.stdout.count("COMMAND_COMPLETED") == 3 and "COMMAND_FAILED" not in r.stdout


def test_example_client_speaks_the_protocol():
    """examples/worker_client.py (the counterpart of the reference's examples/python/main.py) against the real worker:
    a verify-only session needs no GPU."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("worker_client", os.path.join(os.path.dirname(GOLD), "..", "examples", "worker_client.py"))
    wc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(wc)
    if not os.path.exists(BIN):
        pkg.build()
    w = wc.Worker()
    try:
        base = os.path.join(GOLD, "complex_100")
        _, lines, failed = w.run(f"verify --proof {base}.proof_rs.json --public {base}.public.json --vk {base}.vk.json")
        assert not failed and lines[-1].endswith("COMMAND_COMPLETED")
        _, lines, failed = w.run(f"verify --proof {GOLD}/complex_6.proof_rs.json --public {base}.public.json --vk {base}.vk.json")
        assert failed
    finally:
        w.close()
    assert w.proc.returncode == 0
