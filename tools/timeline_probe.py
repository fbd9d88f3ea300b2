"""One pass of the benchmark batch with a FIZZ_RUNNER_DEBUG build of the library (FIZZ_B200_LIB=...): when every chunk
boundary is reached, when each runner launch ends, how long the runner lists are."""
import sys, time
import torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 96
mode = sys.argv[2] if len(sys.argv) > 2 else "hybrid"
rank, G = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 1)
import types, bench
specs, ids = bench.build_specs(types.SimpleNamespace(config=3, worlds=65536, max_calls=4096), rank, G)
bw = fz.BatchedWorld(specs, ids)
bw.set_tuning(mode, chunk)
snap = bw.snapshot()
for rep in range(3):
    bw.restore(snap); torch.cuda.synchronize()
    t0 = time.perf_counter(); bw.step(10000); torch.cuda.synchronize()
    print("pass %d: %.1f ms" % (rep, (time.perf_counter() - t0) * 1e3), flush=True)
