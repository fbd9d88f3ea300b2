"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/ by running the UNMODIFIED reference (through
oracle/ref_harness.py) in the build container.  Re-run with:  python oracle/make_golden.py

Outputs (all small, committed):
  (each <SCENE>.npz carries the scene's input text in `scene_text`; fizz2d_b200.scenes materialises
   simulations/<SCENE>.in from it on demand, so tests, smoke() and bench.py run where /root/reference is absent)
  tests/golden/<SCENE>.npz              unperturbed scene, per-step full-precision trace
  tests/golden/perturbed_<NAME>.npz     perturbed worlds (SURVEY 8d law), checkpoints + contact traces
  tests/golden/anchors.json             1000-step sha256 anchors, regenerated here
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_harness as rh  # noqa: E402
import fizz2d_b200 as fz  # noqa: E402  (only for perturbation_deltas: the single definition of the law)

OUT = os.path.join(ROOT, "tests", "golden")
SCENES = ["BOX_1SQUARE", "BOX_1SQUAREFRICTION", "RAMPWORLD_2TRIANGLES_5SQUARES", "2CHAMBERS_3BALLS",
          "STACKEDCHAMBERS_1SQUARE", "PENTAGONVOID_1SQUARE", "CORRIDOR_2SQUARES", "PLANE_2SQUARES",
          "RAMP_1SQUARE_1RECTANGLE"]


def scene_path(name):
    return os.path.join(rh.SCENES, name + ".in")


def contacts_arrays(w, relevant_only=True):
    c = [x for x in w.contacts if not (relevant_only and x[2] in (243, 253))]
    # renumber `call` as the 1-based index among state-relevant calls of the step: the harness numbers all
    # calls of a step in order and diagnostics only come last, so the number is already right
    a = np.array([(x[0], x[1], x[3], x[4]) for x in c], dtype=np.int32).reshape(-1, 4)
    f = np.array([(x[5], x[6], x[7]) for x in c], dtype=np.float64).reshape(-1, 3)
    return a, f


def init_arrays(w):
    fr, fx = w.init_constants()
    return dict(
        init_com=np.array([b["com"] for b in fr]), init_vel=np.array([b["vel"] for b in fr]),
        init_radius=np.array([b["radius"] for b in fr]), init_mass=np.array([b["mass"] for b in fr]),
        init_offx=np.array([x for b in fr for x in b["offx"]]), init_offy=np.array([x for b in fr for x in b["offy"]]),
        nverts_free=np.array([len(b["offx"]) for b in fr], dtype=np.int32),
        fixed_com=np.array([b["com"] for b in fx]).reshape(-1, 2), fixed_radius=np.array([b["radius"] for b in fx]),
        nverts_fixed=np.array([len(b["pts"]) for b in fx], dtype=np.int32),
        fixed_xy=np.array([p for b in fx for p in b["pts"]]).reshape(-1, 2))


def trace_world(w, nsteps, with_dump_hash=True):
    E, COM, V, H, CALLS, DUMP = [], [], [], [], [], hashlib.sha256()
    for t in range(nsteps):
        E.append(w.energy())
        if not w.update():
            break
        COM.append(w.com())
        V.append(np.concatenate(w.verts()))
        H.append(w.heat())
        CALLS.append(w.calls_per_step[-1])
        DUMP.update(w.dump().encode())
    ci, cf = contacts_arrays(w)
    return dict(energy_before=np.array(E), com=np.array(COM), verts=np.array(V), heat=np.array(H),
                calls=np.array(CALLS, dtype=np.int32), contacts_idx=ci, contacts_val=cf,
                dump_sha=np.frombuffer(DUMP.digest(), dtype=np.uint8), capped_at=np.int64(w.capped_at),
                steps_done=np.int64(w.step_index))


def gen_scene(name, nsteps, gravity=None, tag=None):
    t0 = time.time()
    w = rh.RefWorld(scene_path(name), gravity=gravity)
    d = init_arrays(w)
    d.update(trace_world(w, nsteps))
    d["scene_text"] = np.array(open(scene_path(name)).read())
    d["gravity"] = np.array(gravity if gravity is not None else [0.0, 0.0])
    fn = os.path.join(OUT, (tag or name) + ".npz")
    np.savez_compressed(fn, **d)
    print("%-44s %5d steps %4d contacts  %.1fs  %d KB" % (os.path.basename(fn), d["steps_done"],
          len(d["contacts_idx"]), time.time() - t0, os.path.getsize(fn) // 1024))


def gen_perturbed(tag, scene_for_world, deltas, worlds, nsteps, checkpoints, max_calls):
    """scene_for_world(w) -> scene name; deltas (W, Fmax, 4)."""
    t0 = time.time()
    rec = dict(worlds=np.array(worlds, dtype=np.int64), checkpoints=np.array(checkpoints, dtype=np.int64),
               deltas=deltas[worlds], max_calls=np.int64(max_calls), nsteps=np.int64(nsteps))
    scenes, ncap = [], 0
    for k, wi in enumerate(worlds):
        name = scene_for_world(wi)
        scenes.append(name)
        text = rh.perturbed_scene_text(scene_path(name), deltas[wi])
        w = rh.RefWorld(text=text, max_calls=max_calls)
        ic = init_arrays(w)
        cps = {}
        ebefore = []
        for t in range(nsteps):
            if t in (0, nsteps - 1):
                ebefore.append(w.energy())
            if not w.update():
                break
            if (t + 1) in checkpoints:
                cps[t + 1] = w.com()
        ci, cf = contacts_arrays(w)
        rec["w%d_final_com" % wi] = w.com()
        rec["w%d_final_verts" % wi] = np.concatenate(w.verts())
        rec["w%d_final_heat" % wi] = np.float64(w.heat())
        rec["w%d_final_energy" % wi] = w.energy()
        rec["w%d_steps_done" % wi] = np.int64(w.step_index)
        rec["w%d_capped_at" % wi] = np.int64(w.capped_at)
        rec["w%d_calls" % wi] = np.array(w.calls_per_step, dtype=np.int32)
        rec["w%d_contacts_idx" % wi] = ci
        rec["w%d_contacts_val" % wi] = cf
        for s, c in cps.items():
            rec["w%d_com_at_%d" % (wi, s)] = c
        for key in ("init_com", "init_vel", "init_radius", "init_mass", "init_offx", "init_offy"):
            rec["w%d_%s" % (wi, key)] = ic[key]
        ncap += w.capped_at >= 0
    rec["scenes"] = np.array(scenes)
    fn = os.path.join(OUT, "perturbed_%s.npz" % tag)
    np.savez_compressed(fn, **rec)
    print("%-44s %3d worlds x %d steps, %d capped  %.1fs  %d KB" % (os.path.basename(fn), len(worlds), nsteps, ncap,
          time.time() - t0, os.path.getsize(fn) // 1024))


def gen_ga(tag, scene, seed, worlds, nsteps, checkpoints, max_calls):
    """GA population: the reference run on scene text = fixed geometry + one random polygon."""
    t0 = time.time()
    sc = fz.load_in(scene_path(scene))
    nverts, pts, vel = fz.ga.random_convex_polygons(seed, max(worlds) + 1)
    rec = dict(worlds=np.array(worlds, dtype=np.int64), checkpoints=np.array(checkpoints, dtype=np.int64),
               max_calls=np.int64(max_calls), nsteps=np.int64(nsteps), scene=np.array(scene), seed=np.int64(seed),
               scene_text=np.array(open(scene_path(scene)).read()))
    for wi in worlds:
        text = fz.ga.fixed_scene_text(sc) + fz.ga.polygon_in_text(pts[wi], vel[wi])
        w = rh.RefWorld(text=text, max_calls=max_calls)
        ic = init_arrays(w)
        pre = "w%d_" % wi
        rec[pre + "points"] = pts[wi]
        rec[pre + "vel"] = vel[wi]
        for t in range(nsteps):
            if not w.update():
                break
            if (t + 1) in checkpoints:
                rec[pre + "com_at_%d" % (t + 1)] = w.com()
        ci, cf = contacts_arrays(w)
        rec[pre + "final_com"] = w.com()
        rec[pre + "final_verts"] = np.concatenate(w.verts())
        rec[pre + "final_heat"] = np.float64(w.heat())
        rec[pre + "final_energy"] = w.energy()
        rec[pre + "steps_done"] = np.int64(w.step_index)
        rec[pre + "capped_at"] = np.int64(w.capped_at)
        rec[pre + "calls"] = np.array(w.calls_per_step, dtype=np.int32)
        rec[pre + "contacts_idx"] = ci
        rec[pre + "contacts_val"] = cf
        for key in ("init_com", "init_vel", "init_radius", "init_mass", "init_offx", "init_offy"):
            rec[pre + key] = ic[key]
    fn = os.path.join(OUT, "ga_%s.npz" % tag)
    np.savez_compressed(fn, **rec)
    print("%-44s %3d worlds x %d steps  %.1fs  %d KB" % (os.path.basename(fn), len(worlds), nsteps, time.time() - t0,
          os.path.getsize(fn) // 1024))


def gen_anchors():
    """The SURVEY 8c table: 1000 steps, sha256 over str(world) after each step, over the energy rows written
    before each step, and over the contact list 'step:call#:i-j' of every check_collisions call (incl. the two
    diagnostic calls per step, numbered in call order)."""
    out = {}
    for name in SCENES[:6]:
        w = rh.RefWorld(scene_path(name))
        hp, he = hashlib.sha256(), hashlib.sha256()
        for t in range(1000):
            he.update((','.join(str(x) for x in w.world.energy()) + '\n').encode())
            w.update()
            hp.update(w.dump().encode())
        hc = hashlib.sha256('\n'.join('%d:%d:%d-%d' % (c[0], c[1], c[3], c[4]) for c in w.contacts).encode()).hexdigest()
        rel = [c for c in w.contacts if c[2] not in (243, 253)]
        hr = hashlib.sha256('\n'.join('%d:%d:%d-%d' % (c[0], c[1], c[3], c[4]) for c in rel).encode()).hexdigest()
        out[name] = dict(planes=hp.hexdigest(), energy=he.hexdigest(), contacts_all=hc, contacts_relevant=hr,
                         n_contacts_all=len(w.contacts), n_contacts_relevant=len(rel),
                         calls_relevant=int(sum(w.calls_per_step)),
                         final_energy=[float(x) for x in w.energy()],
                         final_com=[[float(x) for x in r] for r in w.com()])
        print("anchor", name, out[name]["planes"][:16], out[name]["energy"][:16], out[name]["contacts_all"][:16])
    with open(os.path.join(OUT, "anchors.json"), "w") as f:
        json.dump(out, f, indent=1)


def gen_critical():
    """The benchmark regime itself: the two worlds with the longest chains of BASELINE config 3 -- a square
    trapped inside a fixed rectangle from timestep 83 / 6 on, ~330 / ~275 check_collisions() per timestep, every resolve
    pass one tuple -- with the benchmark's watchdog (max_calls = 4096), through the reference itself."""
    d = fz.perturbation_deltas(20260003, 65536, 7)
    gen_perturbed("critical_rampworld", lambda w: "RAMPWORLD_2TRIANGLES_5SQUARES", d, [12860, 63739], 130, [60, 100, 130], 4096)


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--critical" in sys.argv:
        gen_critical()
        return
    for name in SCENES:
        gen_scene(name, 320)
    # gravity extension (reachable only through the Python API in the reference)
    gen_scene("BOX_1SQUARE", 200, gravity=[0.0, -300.0], tag="gravity_BOX_1SQUARE")
    gen_scene("PLANE_2SQUARES", 200, gravity=[3.0, 120.0], tag="gravity_PLANE_2SQUARES")
    # config 2: BOX_1SQUAREFRICTION, seed 20260002
    d = fz.perturbation_deltas(20260002, 4096, 1)
    gen_perturbed("config2_boxfriction", lambda w: "BOX_1SQUAREFRICTION", d, list(range(1, 25)) + [4095], 200,
                  [50, 100, 150, 200], 4096)
    # config 3: RAMPWORLD, seed 20260003 (watchdog 256 keeps generation time bounded)
    d = fz.perturbation_deltas(20260003, 65536, 7)
    gen_perturbed("config3_rampworld", lambda w: "RAMPWORLD_2TRIANGLES_5SQUARES", d, list(range(1, 17)) + [65535], 300,
                  [100, 200, 300], 256)
    # config 4: even -> 2CHAMBERS, odd -> STACKEDCHAMBERS, seed 20260004; deltas drawn for 3 bodies per world,
    # a one-body world uses row 0
    d = fz.perturbation_deltas(20260004, 262144, 3)
    gen_perturbed("config4_mixed", lambda w: "2CHAMBERS_3BALLS" if w % 2 == 0 else "STACKEDCHAMBERS_1SQUARE", d,
                  list(range(1, 13)), 300, [100, 200, 300], 256)
    gen_ga("config5_pentagonvoid", "PENTAGONVOID_1SQUARE", 20260005, list(range(0, 24)), 300, [100, 200, 300], 256)
    gen_critical()
    gen_anchors()


if __name__ == "__main__":
    main()
