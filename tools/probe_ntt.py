"""Scratch: a few forward NTTs (2^22 x 3, device-resident) for ncu captures."""
import sys
import torch
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package(); B = pkg.bindings; lib = pkg.lib(); lib.set_device("CUDA", 0)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n, batch = 1 << lg, 3
lib.ntt_init_domain(lib.get_root_of_unity(n))
x = torch.randint(0, 1 << 28, (n * batch, 8), dtype=torch.int32, device="cuda"); y = torch.empty_like(x)
cfg = B.NTTConfig.default(); cfg.batch_size = batch; cfg.are_inputs_on_device = cfg.are_outputs_on_device = True
for _ in range(3):
    lib.ntt(x.data_ptr(), B.kForward, cfg, out=y.data_ptr(), size=n)
torch.cuda.synchronize()
