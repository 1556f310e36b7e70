"""Improved staggered Dhop fp32 at 48^4 (BASELINE configs[4], Benchmark_staggered shape) on one B200: ms, GFlop/s (1146 flop/site,
ref: benchmarks/Benchmark_staggered.cc:105) and algorithmic GB/s (1200 B/site: 16 links x 72 B + colour vector in and out)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_b200 as gb

L = int(sys.argv[1]) if len(sys.argv) > 1 else 48
ncall = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ctx = gb.Context(0); ctx.comm_init(0, 1, None)
grid = gb.GridCartesian(ctx, (L,) * 4)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
for prec, name, w in ((gb.F32, "fp32", 4), (gb.F64, "fp64", 8)):
    U = gb.LatticeGaugeField(grid, prec).random(1)
    D = gb.ImprovedStaggeredFermion(U, U, grid, 0.1)
    del U
    src = gb.LatticeStaggeredFermion(grid, 1, prec).random(2)
    out = gb.LatticeStaggeredFermion(grid, 1, prec)
    for _ in range(5):
        D.Dhop(src, out, 0)
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(ncall):
        D.Dhop(src, out, 0)
    ms = ctx.timer_stop() / ncall
    sites = L ** 4
    bytes_alg = (2 * 6 * w + 16 * 18 * w) * sites
    print(json.dumps({"op": f"ImprovedStaggeredFermion::Dhop {name} {L}^4", "ms": ms, "gflops": 1146.0 * sites / ms / 1e6,
                      "alg_GBs": bytes_alg / ms / 1e6, "frac_of_measured_hbm_peak": bytes_alg / ms / 1e6 / peak}), flush=True)
    del D, src, out
