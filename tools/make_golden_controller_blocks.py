"""Generates tests/golden/controller_blocks.npz by RUNNING THE REFERENCE'S OWN CODE in this container.

The device-side controller blocks (jb_set_pd_controller_full) restate three pure-numpy/numba functions of
gym_jiminy: `integrate_zoh`, `pd_controller` (blocks/proportional_derivative_controller.py:23-165) and
`apply_safety_limits` (blocks/motor_safety_limit.py:21-86).  Their module cannot be imported here (it imports the
compiled `jiminy_py.core`), but the functions themselves depend on numpy and numba only, so this script extracts
their source text from the reference tree with `ast`, compiles it unchanged, and records input / output vectors.
Nothing of the reference is copied into the repo: only the numeric vectors are committed.

Run once (reference at /root/reference, numba installed):  python tools/make_golden_controller_blocks.py
"""
import ast
import os
import sys

import numpy as np

REF = "/root/reference/python/gym_jiminy/common/gym_jiminy/common/blocks"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "controller_blocks.npz")


def extract(path, names):
    """Source text of the named top-level functions (the jitted definition, not the typing @overload stubs)."""
    src = open(path).read()
    tree = ast.parse(src)
    chunks = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            decos = [ast.get_source_segment(src, d) or "" for d in node.decorator_list]
            if any(d.startswith("overload") for d in decos):
                continue
            start = min([d.lineno for d in node.decorator_list] + [node.lineno])
            text = "\n".join(src.splitlines()[start - 1:node.end_lineno])
            chunks.append(text.replace("@no_type_check\n", ""))
    return "\n\n".join(chunks)


def main():
    # the extracted source goes, unchanged, into a scratch module outside the repo (numba's cache=True needs a file)
    import importlib.util
    import tempfile
    tmp = tempfile.mkdtemp(prefix="jb_golden_")
    mod_path = os.path.join(tmp, "ref_controller_blocks.py")
    with open(mod_path, "w") as fh:
        fh.write("from typing import Optional\nimport numpy as np\nimport numba as nb\n\n")
        fh.write(extract(os.path.join(REF, "proportional_derivative_controller.py"), {"integrate_zoh", "pd_controller", "pd_adapter"}))
        fh.write("\n\n")
        fh.write(extract(os.path.join(REF, "motor_safety_limit.py"), {"apply_safety_limits"}))
        fh.write("\n\nfrom typing import Tuple\nEARTH_SURFACE_GRAVITY = 9.81   # blocks/mahony_filter.py:23\n\n")
        fh.write(extract(os.path.join(REF, "..", "utils", "math.py"), {"compute_tilt_from_quat", "matrices_to_quat"}))
        fh.write("\n\n")
        fh.write(extract(os.path.join(REF, "mahony_filter.py"), {"mahony_filter"}))
        fh.write("\n")
    spec = importlib.util.spec_from_file_location("ref_controller_blocks", mod_path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pd_controller, apply_safety_limits = mod.pd_controller, mod.apply_safety_limits

    rng = np.random.default_rng(20240924)
    n_cases, nm = 150, 6
    rec = {k: [] for k in ("state_in", "lower", "upper", "dt", "state_out", "enc", "kp", "kd", "effort_limit", "tau",
                           "s_command", "s_q", "s_v", "s_kp", "s_kd", "s_lo", "s_hi", "s_vlim", "s_elim", "s_out")}
    for case in range(n_cases):
        # bounds: position in +-[0.5, 3], velocity limit [1, 10], acceleration limit [5, 200]
        pmax, vmax, amax = rng.uniform(0.5, 3.0, nm), rng.uniform(1.0, 10.0, nm), rng.uniform(5.0, 200.0, nm)
        lower, upper = np.stack([-pmax * rng.uniform(0.5, 1.0, nm), -vmax, -amax]), np.stack([pmax, vmax, amax])
        state = np.stack([rng.uniform(-1.2, 1.2, nm) * pmax, rng.uniform(-1.2, 1.2, nm) * vmax, rng.uniform(-1.5, 1.5, nm) * amax])
        dt = float(rng.choice([0.0, 1e-3, 5e-3, 1e-2, 4e-2]))
        enc = np.stack([rng.uniform(-1.0, 1.0, nm) * pmax, rng.uniform(-1.0, 1.0, nm) * vmax])
        kp, kd, elim = rng.uniform(10.0, 2000.0, nm), rng.uniform(0.0, 0.05, nm), rng.uniform(20.0, 200.0, nm)
        rec["state_in"].append(state.copy()); rec["lower"].append(lower); rec["upper"].append(upper); rec["dt"].append(dt)
        rec["enc"].append(enc); rec["kp"].append(kp); rec["kd"].append(kd); rec["effort_limit"].append(elim)
        out = np.zeros(nm)
        st = state.copy()
        pd_controller(enc, st, lower, upper, kp, kd, elim, dt, out)     # integrates `st` in place, then the PD law
        rec["state_out"].append(st); rec["tau"].append(out.copy())
        # safety limits
        cmd = rng.uniform(-1.5, 1.5, nm) * elim
        q, v = rng.uniform(-1.3, 1.3, nm) * pmax, rng.uniform(-1.5, 1.5, nm) * vmax
        skp, skd = rng.uniform(1.0, 50.0, nm), rng.uniform(0.1, 5.0, nm)
        so = np.zeros(nm)
        apply_safety_limits(cmd, q, v, skp, skd, lower[0], upper[0], vmax, elim, so)
        for k, x in zip(("s_command", "s_q", "s_v", "s_kp", "s_kd", "s_lo", "s_hi", "s_vlim", "s_elim", "s_out"),
                        (cmd, q, v, skp, skd, lower[0], upper[0], vmax, elim, so)):
            rec[k].append(x.copy())
    # Mahony filter (blocks/mahony_filter.py:28-101) and matrices_to_quat (utils/math.py:293-350)
    for k in ("m_q", "m_gyro", "m_acc", "m_bias", "m_kp", "m_ki", "m_dt", "m_q_out", "m_omega", "m_bias_out", "r_mat", "r_quat"):
        rec[k] = []
    for case in range(n_cases):
        M = 2
        q = rng.normal(size=(4, M)); q /= np.linalg.norm(q, axis=0)
        gyro = rng.normal(size=(3, M)) * (0.0 if case % 17 == 0 else 1.5)
        acc = rng.normal(size=(3, M)) * 2.0 + np.array([[0.0], [0.0], [9.81]])
        if case % 17 == 0:   # exercises the early return: no IMU motion, acceleration aligned with the estimated gravity
            v = np.stack(mod.compute_tilt_from_quat(q))
            acc = 9.81 * v
        bias = rng.normal(size=(3, M)) * (0.0 if case % 17 == 0 else 0.05)
        kp, ki, dt = float(rng.uniform(0.0, 2.0)), float(rng.uniform(0.0, 0.5)), float(rng.choice([1e-3, 5e-3, 1e-2]))
        qo, om, cf, bo = q.copy(), np.zeros((3, M)), np.zeros((3, M)), bias.copy()
        mod.mahony_filter(qo, om, cf, gyro, acc, bo, kp, ki, dt)
        for k, x in zip(("m_q", "m_gyro", "m_acc", "m_bias", "m_kp", "m_ki", "m_dt", "m_q_out", "m_omega", "m_bias_out"),
                        (q, gyro, acc, bias, kp, ki, dt, qo, om, bo)):
            rec[k].append(np.array(x))
        # random rotation matrices (all four branches of the conversion are hit over the cases)
        quat = rng.normal(size=4); quat /= np.linalg.norm(quat)
        x, y, z, w = quat
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        out = np.zeros((4, 1))
        mod.matrices_to_quat((R,), out)
        rec["r_mat"].append(R); rec["r_quat"].append(out[:, 0].copy())
    np.savez_compressed(OUT, **{k: np.array(v) for k, v in rec.items()})
    print("wrote", OUT, {k: np.array(v).shape for k, v in rec.items()})

    # pd_adapter (blocks/proportional_derivative_controller.py:166-262): its own file and random stream, so that the
    # vectors above never change
    rng = np.random.default_rng(20240925)
    ad = {k: [] for k in ("action", "order", "state_in", "lower", "upper", "instantaneous", "has_deadband", "deadband",
                          "step_dt", "out", "state_out")}
    for case in range(240):
        pmax, vmax, amax = rng.uniform(0.5, 3.0, nm), rng.uniform(1.0, 10.0, nm), rng.uniform(5.0, 200.0, nm)
        lower, upper = np.stack([-pmax, -vmax, -amax]), np.stack([pmax, vmax, amax])
        state = np.stack([rng.uniform(-1.0, 1.0, nm) * pmax, rng.uniform(-1.0, 1.0, nm) * vmax, rng.uniform(-1.0, 1.0, nm) * amax])
        order, inst, has_db = case % 2, (case // 2) % 2 == 1, (case // 4) % 2 == 1
        step_dt = float([0.04, 1e-3, 0.0, 0.02][(case // 8) % 4])
        action = rng.uniform(-1.5, 1.5, nm) * (pmax if order == 0 else vmax)
        if has_db:
            action[rng.integers(nm)] *= 1e-3      # something inside the dead band
        deadband = rng.uniform(0.0, 0.2, nm)
        out, st = np.full(nm, 7.0), state.copy()
        mod.pd_adapter(action.copy(), order, st, lower, upper, inst, deadband if has_db else None, step_dt, out)
        for k, x in zip(ad, (action, order, state, lower, upper, inst, has_db, deadband, step_dt, out, st)):
            ad[k].append(np.array(x))
    out_path = os.path.join(os.path.dirname(OUT), "pd_adapter.npz")
    np.savez_compressed(out_path, **{k: np.array(v) for k, v in ad.items()})
    print("wrote", out_path, {k: np.array(v).shape for k, v in ad.items()})


if __name__ == "__main__":
    sys.exit(main())
