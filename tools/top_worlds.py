import sys
import numpy as np, torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz
bw = fz.BatchedWorld.from_in(fz.scenes.scene_path("RAMPWORLD_2TRIANGLES_5SQUARES"), n_worlds=65536, perturb=True, seed=20260003)
bw.step(1000)
c = bw.calls(); st = bw.status(); sd = bw.steps_done()
o = np.argsort(-c)
print("top", [(int(i), int(c[i]), int(st[i]), int(sd[i])) for i in o[:16]])
live = st == 0
print("live top", [(int(i), int(c[i])) for i in o if live[i]][:12])
print("pct live", np.percentile(c[live], [50, 90, 99, 99.9, 100]))
