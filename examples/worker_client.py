"""Drive the `prove` worker the way the reference's examples/python/main.py drives `cargo run --release`
(/root/reference/examples/python/main.py:19-80): one long-lived process, commands on stdin, wait for the
COMMAND_COMPLETED sentinel.  The worker keeps the ZKeyCache, so every proof after the first is warm; `verify` replaces the
reference example's external `snarkjs g16v` call.

    python examples/worker_client.py [--witness W --zkey Z --vk VK --out DIR --iterations N]

Defaults prove the committed 100-constraint instance under tests/golden (needs a CUDA device: there is no CPU backend).
"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "icicle-snark_b200", "bin", "prove")
GOLD = os.path.join(ROOT, "tests", "golden", "complex_100")


class Worker:
    def __init__(self, path=WORKER, env=None):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.proc = subprocess.Popen([path], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                     text=True, env=env)

    def run(self, command):
        """Send one line, collect output up to the sentinel; returns (seconds, lines, failed)."""
        t0 = time.time()
        self.proc.stdin.write(command + "\n")
        self.proc.stdin.flush()
        lines = []
        while True:
            line = self.proc.stdout.readline()
            if not line:
                raise RuntimeError("worker exited: " + " | ".join(lines))
            line = line.strip()
            lines.append(line)
            if "COMMAND_COMPLETED" in line:
                break
        return time.time() - t0, lines, any("COMMAND_FAILED" in l for l in lines)

    def close(self):
        try:
            self.run("exit")
        finally:
            self.proc.wait(timeout=30)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--witness", default=GOLD + ".wtns")
    ap.add_argument("--zkey", default=GOLD + ".zkey")
    ap.add_argument("--vk", default=GOLD + ".vk.json")
    ap.add_argument("--out", default="/tmp")
    ap.add_argument("--iterations", type=int, default=3)
    args = ap.parse_args()
    proof, public = os.path.join(args.out, "proof.json"), os.path.join(args.out, "public.json")
    w = Worker()
    try:
        cmd = f"prove --witness {args.witness} --zkey {args.zkey} --proof {proof} --public {public} --device CUDA"
        for i in range(args.iterations):
            dt, lines, failed = w.run(cmd)
            print(f"prove #{i}: {dt * 1e3:.1f} ms{' (cold: builds the ZKeyCache)' if i == 0 else ''}", "FAILED" if failed else "")
            if failed:
                print("\n".join(lines))
                return 1
        dt, lines, failed = w.run(f"verify --proof {proof} --public {public} --vk {args.vk}")
        print(f"verify: {dt * 1e3:.1f} ms ->", "INVALID" if failed else "valid")
        return 1 if failed else 0
    finally:
        w.close()


if __name__ == "__main__":
    sys.exit(main())
