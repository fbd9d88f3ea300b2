import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz
from fizz2d_b200.batch import GroupSpec
wi = int(sys.argv[1]) if len(sys.argv) > 1 else 63739
sc = fz.load_in(fz.scenes.scene_path("RAMPWORLD_2TRIANGLES_5SQUARES"))
d = fz.perturbation_deltas(20260003, 65536, 7)[[wi]]
bw = fz.BatchedWorld([GroupSpec.from_scene(sc, 1, np.ascontiguousarray(d))])
bw.step(256)
g = bw.groups[0]
L = g.layout
cnt = g.slab[L.sched + L.stride: L.sched + L.stride + 16].view(torch.int64)
torch.cuda.synchronize(); cnt[4:12] = 0
c0 = bw.calls()[0]
bw.step(512); torch.cuda.synchronize()
v = cnt.cpu().numpy()
calls = bw.calls()[0] - c0
print("calls", calls, "fast", v[7], "cycles trial %d spec %d resolve(general) %d fast %d" % (v[8], v[9], v[10], v[11]))
print("per step (cycles): trial %.0f spec %.0f res %.0f fast %.0f ; fast cycles per fast call %.0f" % (v[8]/512, v[9]/512, v[10]/512, v[11]/512, v[11]/max(1,v[7])))
