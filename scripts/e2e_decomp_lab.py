"""gb_op_dhop_host on a decomposed lattice, timed on ONE GPU: GB_SELF_HALO=<mask> routes the chosen dimensions through the halo path
(pack + stores into this rank's own receive buffers + flags), so the pipelined decomposed form of dhop_host.cu is exercised and timed
without a second GPU.  usage: [GB_SELF_HALO=8|12] [GB_HOST_PIPE_DECOMP=0] python scripts/e2e_decomp_lab.py Lx Ly Lz Lt Ls [calls]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grid_b200 as gb

dims = tuple(int(x) for x in sys.argv[1:5]); Ls = int(sys.argv[5]); calls = int(sys.argv[6]) if len(sys.argv) > 6 else 8
ctx = gb.Context(0)
grid = gb.GridCartesian(ctx, dims)
D = gb.DomainWallFermion(gb.LatticeGaugeField(grid, gb.F32).random(1), grid, Ls, 0.1, 1.8)
src = gb.LatticeFermion(grid, Ls, gb.F32).random(2); out = gb.LatticeFermion(grid, Ls, gb.F32)
D.Dhop(src, out, 0)
ref = out.export_lex()
n = src.local_sites
hin = torch.empty((n, 4, 3), dtype=torch.complex64, pin_memory=True).numpy(); hout = torch.empty((n, 4, 3), dtype=torch.complex64, pin_memory=True).numpy()
hin[...] = src.export_lex()
D.Dhop_host(hin, hout, 0); D.Dhop_host(hin, hout, 0)
err = float(np.max(np.abs(hout - ref)) / np.max(np.abs(ref)))
ctx.synchronize()
l0 = ctx.launch_count(); t0 = time.perf_counter()
for _ in range(calls):
    D.Dhop_host(hin, hout, 0)
ctx.synchronize()
s = (time.perf_counter() - t0) / calls
print(json.dumps({"dims": dims, "Ls": Ls, "self_halo": os.environ.get("GB_SELF_HALO", "0"), "pipe_decomp": os.environ.get("GB_HOST_PIPE_DECOMP", "1"),
                  "ms_per_call": s * 1e3, "GBs_per_direction": hin.nbytes / s / 1e9, "launches_per_call": (ctx.launch_count() - l0) / calls, "max_err_vs_device_hop": err}))
