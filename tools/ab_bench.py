"""Same-box A/B harness: total ms of one 10k-step pass of the benchmark batch, alternating environment settings
(argv entries K=V or `base`).  Box-to-box and run-to-run noise is +-2 %: only trust comparisons made inside ONE gpurun call.
(The engine currently reads no environment variables; add a getenv toggle around the code under test.)"""
import os, subprocess, sys
variants = sys.argv[1:]
code = '''
import sys, time, torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz
bw = fz.BatchedWorld.from_in(fz.scenes.scene_path("RAMPWORLD_2TRIANGLES_5SQUARES"), n_worlds=65536, perturb=True, seed=20260003)
snap = bw.snapshot()
out = []
for rep in range(3):
    bw.restore(snap); torch.cuda.synchronize()
    t0 = time.perf_counter(); bw.step(10000); torch.cuda.synchronize(); out.append((time.perf_counter() - t0) * 1e3)
print(" ".join("%.1f" % t for t in out))
'''
for rnd in range(2):
    for v in variants:
        env = dict(os.environ)
        if v != "base":
            k, val = v.split("=")
            env[k] = val
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print("%-28s %s" % (v, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]))
