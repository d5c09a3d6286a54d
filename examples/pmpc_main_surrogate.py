#!/usr/bin/env python
"""The loop of the reference's PMPC/main.py (:59-125) with the drop-in controller and a surrogate plant.

Only two things differ from the reference script: the import (``from dart_b200 import PMPC`` instead of ``from src
import PMPC``) and MuJoCo, which is replaced by duck-typed ``model``/``data`` objects plus one RK4 step of the same
tray model (with an unmodelled Coulomb term) where the script calls ``mujoco.mj_step``.  The dual-arm impedance
controller that would consume the quaternion is outside the replaced path and is omitted.

    python examples/pmpc_main_surrogate.py --target 0.1 0 0.05 0 0 0 --friction 0.1 --steps 1500
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dart_b200 import PMPC, GravityModel, StateHolder        # noqa: E402  (reference: from src import PMPC)


def plant_step(x, u, mu, coulomb, Ts, g):
    def f(x):
        vn = -g * (u[0] ** 2 + u[1] ** 2)
        ax = g * np.sin(u[0]) - mu * x[1] - coulomb * abs(g) * np.tanh(x[1] / 0.01)
        ay = g * np.sin(u[1]) - mu * x[3] - coulomb * abs(g) * np.tanh(x[3] / 0.01)
        return np.array([x[1], ax, x[3], ay, vn, (vn - x[5]) / Ts])
    k1 = f(x); k2 = f(x + Ts / 2 * k1); k3 = f(x + Ts / 2 * k2); k4 = f(x + Ts * k3)
    return x + Ts / 6 * (k1 + 2 * k2 + 2 * k3 + k4)


def run(target, friction=0.1, object_name="cube", steps=1500, coulomb=0.01, verbose=True):
    model = GravityModel(g=-9.81, timestep=0.002)       # reference: mujoco.MjModel.from_xml_path(world)
    data = StateHolder()                                # reference: mujoco.MjData(model)
    tray = np.array([0.0, 0.0, 0.4])
    data.body(object_name).xpos[:] = tray + [0.0, 0.0, 0.03]          # keyframe: object 3 cm above the tray centre

    mpc_params = {"Ts": model.opt.timestep, "nx": 6, "nu": 2, "N": 15, "Qp": 400, "Qv": 2, "R": 0.2,
                  "u_bounds": (-0.6, 0.6), "mu": friction}           # main.py:59-69
    mpc_controller = PMPC(model, data, **mpc_params)
    mpc_controller.target_body = object_name

    target_3d = np.asarray(target, dtype=float)
    log = []
    for k in range(steps):
        target_position = target_3d + np.array([tray[0], 0.0, tray[1], 0.0, tray[2], 0.0])      # main.py:92-96
        u_cmd, loss = mpc_controller.solve(target_position)                                    # main.py:104
        angles = [u_cmd[1], -u_cmd[0], 0.0]                                                    # main.py:107-116
        cx, cy, cz = np.cos(np.array(angles) / 2.0)
        sx, sy, sz = np.sin(np.array(angles) / 2.0)
        quat = np.array([cx * cy * cz + sx * sy * sz, sx * cy * cz - cx * sy * sz,
                         cx * sy * cz + sx * cy * sz, cx * cy * sz - sx * sy * cz])
        # reference: controller.control(pos, quat); mujoco.mj_step(model, data)
        x = plant_step(mpc_controller.get_state(), u_cmd, friction, coulomb, model.opt.timestep, model.opt.gravity[2])
        b = data.body(object_name)
        b.xpos[:] = [x[0], x[2], x[4]]
        b.cvel[3:6] = [x[1], x[3], x[5]]
        err = float(np.hypot(x[0] - target_position[0], x[2] - target_position[2]))
        log.append((k * model.opt.timestep, err, u_cmd[0], u_cmd[1], float(loss[0]), mpc_controller.iters, mpc_controller.status))
        if verbose and k % 250 == 0:
            print(f"t={k * model.opt.timestep:5.2f}s  err={err * 1e3:6.2f} mm  u=({u_cmd[0]:+.3f},{u_cmd[1]:+.3f})  |q|={np.linalg.norm(quat):.6f}  "
                  f"iters={mpc_controller.iters}")
    return np.array(log)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--target", nargs=6, type=float, default=[0.1, 0.0, 0.05, 0.0, 0.0, 0.0])
    ap.add_argument("--friction", type=float, default=0.1)
    ap.add_argument("--object_name", type=str, default="cube")
    ap.add_argument("--steps", type=int, default=1500)
    a = ap.parse_args()
    log = run(a.target, a.friction, a.object_name, a.steps)
    settled = np.where(log[:, 1] < 0.01)[0]
    print(f"final error {log[-1, 1] * 1e3:.2f} mm; within 1 cm after {log[settled[0], 0]:.2f} s" if len(settled) else "not settled")
