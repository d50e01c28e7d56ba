"""Development fuzz (CPU, warp emulator): random sets of flexibility joints on the all-joint-models arm, every lane
count, single evaluations and first steps against the oracle.  Run from the repo root: PYTHONPATH=. python tools/dev/fuzz_flexible_trees.py"""
import os, sys, itertools
sys.path.insert(0,'tests'); sys.path.insert(0,'tests/emul')
import numpy as np
from emul import emul_api
import flexibility_common as fc
import parity_common as pc
from jiminy_b200 import model as M
from jiminy_b200.core import BatchedEngine
from oracle.oracle import OracleBatch
api = emul_api()
robot, _, opt = fc.flexible_branched_arm()
joints = [n for n in robot.joint_names[2:]]
fixed = [n for n, f in robot.frames.items() if f.kind == "fixed_joint" and n != "b_sole_fixed"]
rng = np.random.default_rng(0)
worst = 0
count = 0
for trial in range(30):
    k = rng.integers(1, 5)
    names = list(rng.choice(joints + fixed, size=k, replace=False))
    cfg = [dict(frameName=n, stiffness=rng.uniform(100, 900, 3), damping=rng.uniform(1, 9, 3), inertia=rng.uniform(0.01, 0.05, 3)) for n in names]
    try:
        flex = M.add_flexibility_joints(robot, cfg)
    except Exception as ex:
        print("skip", names, ex); continue
    n = 2
    q, v = pc.random_states(flex, n, rng, base_height=0.55)
    cmd = rng.uniform(-5, 5, size=(n, flex.nmotors))
    for lanes in (0, 1, 2, 4):
        os.environ["JB_LANES"] = str(lanes)
        try:
            eng = BatchedEngine(flex, opt, n, api_=api)
        except Exception as ex:
            print("engine refused", names, lanes, str(ex)[:80]); continue
        orc = OracleBatch(flex, opt, n)
        a0, f0, u0 = orc.compute_dynamics(q, v, cmd)
        a1, f1, u1 = eng.compute_dynamics(q, v, cmd)
        e = np.abs(a1 - a0).max() / max(1, np.abs(a0).max())
        worst = max(worst, e); count += 1
        assert e < 1e-11, (names, lanes, e)
        # a few steps too
        for x in (eng, orc): x.set_command(cmd)
        rc = orc.start(q, v)
        if rc.any():
            try:
                eng.start(q, v); dev = "accepted"
            except Exception as ex:
                dev = type(ex).__name__ + ": " + str(ex)[:60]
            print("   start refused by the oracle", rc, "device:", dev, "status", orc.get_status()); break
        eng.start(q, v)
        eng.step(4e-3); orc.step(4e-3)
        if np.abs(orc.get_state()[2]).max() > 1e4:
            print("   numerically unstable draw (oracle |v| = %.1e): skipped" % np.abs(orc.get_state()[2]).max()); break
        try:
            pc.compare(eng, orc, 1e-9, 1e-7)
        except AssertionError as ex:
            print("MISMATCH", trial, names, "lanes", lanes, eng.describe())
            ea, ef = eng.get_extra_terms()[1:3]; oa, of = orc.get_extra_terms()[1:3]
            print(" nan in device a:", np.argwhere(np.isnan(ea))[:6].tolist(), " oracle a:", np.argwhere(np.isnan(oa))[:6].tolist())
            print(" nan in device f:", np.argwhere(np.isnan(ef))[:6].tolist(), " oracle f:", np.argwhere(np.isnan(of))[:6].tolist())
            print(" status", eng.get_status(), orc.get_status(), flex.joint_names)
            cen_d, cen_o = eng.get_centroidal(), orc.get_centroidal()
            for i,(x,y) in enumerate(zip(cen_d, cen_o)):
                print("  cen", i, np.isnan(np.asarray(x)).sum(), np.isnan(np.asarray(y)).sum())
            raise
    os.environ["JB_LANES"] = "0"
    print(trial, names, "ok", flush=True)
print("cases", count, "worst rhs rel err", worst)
