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
