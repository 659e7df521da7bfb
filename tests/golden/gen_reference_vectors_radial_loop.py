"""Generate tests/golden/reference_vectors_radial_loop.npz by EXECUTING the reference's config-3 driver:
``test_scripts/radial_dynamic_obstacles.py`` -- the WHOLE module, unmodified: ``RadialObstacleSpawner``
(spawn_obstacle :122-191, update_seekers :193-239), ``single_obstacle_CBF1`` (:366-425) and the per-frame loop
``animate`` (:427-507) in the order the file gives them.

Run once in the build container (``python tests/golden/gen_reference_vectors_radial_loop.py``); the GPU box has no
/root/reference, so the vectors are committed.  What is stood in for, and only for import / drawing:

* ``euclid`` / ``cvxopt``: tests/golden/shims (``solvers.cp`` evaluates the reference's closure F and returns the exact
  optimum of that QP -- tests/golden/shims/README.md);
* ``matplotlib`` (.pyplot, .animation, .patches, .axes): inert stand-ins defined below -- the module draws a figure at
  import and moves patches every frame; none of it feeds the arithmetic;
* ``stanley_controller_ellipse``: a module holding the reference's own ``State`` class and constants, pulled out of
  that file by AST (the file itself needs imageio / cvxpy at import).

The reference draws the spawn radius, angle and distance from the UNSEEDED global ``np.random``
(:151,158,165); the generator seeds it (``np.random.seed``) so that the run can be repeated.

Two kinds of runs, 600 frames each (``_NUM_FRAMES``, :53-54):
* M = 1 (``_OBSTACLE_COUNT = 1``, the file as committed), three seeds: per frame the ego state, the control the filter
  returned, the seeker's centre and velocity, the row the reference assembled;
* M = 16: the spawner is rebuilt with ``obstacle_max_count = 16`` and ``spawn_obstacle`` is called 16 times where
  ``animate`` calls it once (frame 1); ``animate`` itself still filters against obstacle 1 only (:448-465), all 16
  seekers chase the ego (update_seekers loops over every seeker) -- this pins the seeker law for a crowd.
"""
import ast
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "test_scripts"))

from cvxopt import matrix, solvers, spdiag, sqrt  # noqa: E402  (the shim)
import euclid  # noqa: E402  (the shim)


# ---------------------------------------------------------------- inert matplotlib
class _Inert:
    """Accepts any call / attribute / index and returns another inert object."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert()

    def __getitem__(self, key):
        return _Inert()

    def __iter__(self):
        return iter([_Inert()])


def _inert_module(name):
    m = types.ModuleType(name)
    m.__getattr__ = lambda attr: _Inert()
    return m


mpl = _inert_module("matplotlib")
mpl.use = lambda *a, **k: None
for sub in ("pyplot", "animation", "patches", "axes"):
    mod = _inert_module("matplotlib." + sub)
    setattr(mpl, sub, mod)
    sys.modules["matplotlib." + sub] = mod
sys.modules["matplotlib"] = mpl


class _Patch(_Inert):
    """patches.Annulus / patches.Circle: keeps what the spawner sets, draws nothing."""

    def __init__(self, center=(0, 0), *a, **k):
        self.center = center
        self.radius = k.get("radius", 0.0)
        self.visible = k.get("visible", True)

    def set_visible(self, v):
        self.visible = v

    def set_radius(self, r):
        self.radius = r

    def set_center(self, c):
        self.center = c


sys.modules["matplotlib.patches"].Annulus = _Patch
sys.modules["matplotlib.patches"].Circle = _Patch
sys.modules["matplotlib.axes"].Axes = _Inert


# ---------------------------------------------------------------- the reference's State, by AST
def extract(path, names, consts, ns):
    tree = ast.parse(open(path).read())
    keep = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            keep.append(node)
        elif isinstance(node, ast.Assign) and all(isinstance(t, ast.Name) and t.id in consts for t in node.targets):
            keep.append(node)
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns


sce = types.ModuleType("stanley_controller_ellipse")
sce.__dict__.update(dict(np=np, matrix=matrix, solvers=solvers, spdiag=spdiag, sqrt=sqrt))
extract(os.path.join(REF, "test_scripts", "stanley_controller_ellipse.py"),
        names={"State", "normalize_angle"}, consts={"k", "Kp", "dt", "L", "lr", "lf", "max_steer", "ZERO_TOL"}, ns=sce.__dict__)
sys.modules["stanley_controller_ellipse"] = sce


def run(seed, count, frames):
    """One run of the reference module from a fresh import."""
    sys.modules.pop("radial_dynamic_obstacles", None)
    np.random.seed(seed)
    rdo = importlib.import_module("radial_dynamic_obstacles")          # the whole file executes, unmodified
    assert rdo._NUM_FRAMES == 600 and rdo._ANIMATION_FPS == 30.0
    if count != 1:
        rdo.obstacle_spawner = rdo.RadialObstacleSpawner(
            state=rdo._ego_state, spawn_radius=None, obstacle_radius_range=rdo._OBSTACLE_RADIUS_RANGE,
            spawn_radius_range=rdo._SPAWN_RADIUS_RANGE, obstacle_max_count=count,
            spawn_annulus_options=rdo._annulus_options, obstacle_patch_options=rdo._obs_patch_options)
        spawn_one = rdo.obstacle_spawner.spawn_obstacle

        def spawn_all(**kw):
            ok = True
            for _ in range(count):
                ok = spawn_one(**kw) and ok
            return ok
        rdo.obstacle_spawner.spawn_obstacle = spawn_all
    sp = rdo.obstacle_spawner
    ego = np.zeros((frames, 4))          # state BEFORE the frame's control
    u = np.zeros((frames, 2))            # (a, delta) returned by the filter in the frame
    row = np.full((frames, 5), np.nan)   # A0 A1 b mask status of obstacle 1's row
    obs = np.full((frames, count, 5), np.nan)   # cx cy vx vy r BEFORE the frame's control (after the spawn of frame 1)
    for i in range(frames):
        st = rdo._ego_state
        ego[i] = [st.x, st.y, st.yaw, st.v]
        solvers.LOG.clear()
        # the spawn of frame 1 happens inside animate, before the control: record the obstacles through a peek
        if i == 1:
            real = rdo.single_obstacle_CBF1

            def peek(**kw):
                for j, key in enumerate(sp.id_spawned):
                    e = sp.obstacle_list[key]
                    obs[i, j] = [e.center.x, e.center.y, e.vel.x, e.vel.y, e.a]
                return real(**kw)
            rdo.single_obstacle_CBF1 = peek
        elif i > 1:
            for j, key in enumerate(sp.id_spawned):
                e = sp.obstacle_list[key]
                obs[i, j] = [e.center.x, e.center.y, e.vel.x, e.vel.y, e.a]
        rdo.animate(i)
        if i == 1:
            rdo.single_obstacle_CBF1 = real
        u[i] = [float(rdo._a_cbf_list[-1]), float(rdo._delta_cbf_list[-1])]
        if solvers.LOG:
            log = solvers.LOG[-1]
            row[i] = [log["A"][0, 0], log["A"][0, 1], log["b"][0], log["mask"], log["status"]]
    st = rdo._ego_state
    fin = np.array([st.x, st.y, st.yaw, st.v])
    fobs = np.array([[sp.obstacle_list[k].center.x, sp.obstacle_list[k].center.y, sp.obstacle_list[k].vel.x,
                      sp.obstacle_list[k].vel.y, sp.obstacle_list[k].a] for k in sp.id_spawned])
    return dict(ego=ego, u=u, row=row, obs=obs, final_ego=fin, final_obs=fobs)


out = {}
for tag, seed, count in (("m1_s0", 20261017, 1), ("m1_s1", 7, 1), ("m1_s2", 123456, 1), ("m16_s0", 99, 16)):
    r = run(seed, count, 600)
    for k, v in r.items():
        out["%s_%s" % (tag, k)] = v
    act = int(np.nansum(r["row"][:, 3] != 0))
    print(tag, "active frames", act, "final ego", r["final_ego"], "min separation",
          float(np.nanmin(np.hypot(r["obs"][2:, 0, 0] - r["ego"][2:, 0], r["obs"][2:, 0, 1] - r["ego"][2:, 1]))))
dst = os.path.join(HERE, "reference_vectors_radial_loop.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, {k: v.shape for k, v in out.items()})
