"""Single trapped world, hybrid mode, one chunk -- the smallest repro of the latency-bound tail (for ncu)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz
from fizz2d_b200.batch import GroupSpec
wi = int(sys.argv[1]) if len(sys.argv) > 1 else 103
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 256
sc = fz.load_in(fz.scenes.scene_path("RAMPWORLD_2TRIANGLES_5SQUARES"))
d = fz.perturbation_deltas(20260003, 65536, 7)[[wi]]
bw = fz.BatchedWorld([GroupSpec.from_scene(sc, 1, np.ascontiguousarray(d))])
bw.set_tuning("hybrid", 256)
bw.step(256)          # reach the trapped regime
torch.cuda.synchronize()
bw.step(steps)
torch.cuda.synchronize()
print("calls", bw.calls()[0])
