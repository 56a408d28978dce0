"""The oracle against the reference's own runnable artefacts: the NumPy twins of the reward library,
the imitation reward and the polynomial reference motion.  Vectors: tests/golden/*.npz (tools/make_golden.py)."""
import ctypes as C
import os

import numpy as np

from conftest import make_handle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_reference_motion_matches_numpy_twin(oracle, model_backlash, poly_table):
    g = np.load(os.path.join(GOLD, "reference_motion.npz"))
    h = make_handle(oracle, model_backlash, poly_table, 1)
    fn = oracle.lib.oduck_test_reference_motion
    fn.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p]
    out = np.zeros(40)
    worst = 0.0
    for k in range(len(g["dx"])):
        oracle.check(fn(h.h, float(g["dx"][k]), float(g["dy"][k]), float(g["dtheta"][k]), int(g["i"][k]), out.ctypes.data))
        worst = max(worst, np.abs(out - g["ref"][k]).max())
    # fp64 Horner (jp.polyval order) vs np.polyval: same arithmetic, tolerance covers summation-order noise on |coef| ~ 2e5
    assert worst < 1e-7, worst


def test_poly_table_matches_twin_grid(poly_table):
    g = np.load(os.path.join(GOLD, "reference_motion.npz"))
    assert poly_table.nb_steps_in_period == int(g["nb_steps_in_period"]) == 27
    assert np.allclose(poly_table.dxs, g["dxs"]) and np.allclose(poly_table.dys, g["dys"]) and np.allclose(poly_table.dthetas, g["dthetas"])
    assert poly_table.coef.shape == (6, 4, 10, 40, 16)
    assert poly_table.dx_range == [-0.148, 0.222] and poly_table.dtheta_range == [-1.111, 1.222]


def test_reward_terms_match_numpy_twins(oracle, model_backlash, poly_table):
    g = np.load(os.path.join(GOLD, "rewards.npz"))
    h = make_handle(oracle, model_backlash, poly_table, 1)
    fn = oracle.lib.oduck_test_rewards
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    out = np.zeros(7)
    worst = np.zeros(7)
    for k in range(len(g["command"])):
        vec = np.concatenate([g["command"][k], g["local_linvel"][k], g["gyro"][k], g["actuator_force"][k], g["action"][k], g["last_act"][k],
                              g["base_qvel"][k], g["q"][k], g["qd"][k], g["contact"][k].astype(np.float64), g["ref"][k]])
        oracle.check(fn(h.h, vec.ctypes.data, out.ctypes.data))
        worst = np.maximum(worst, np.abs(out - g["terms"][k]) / np.maximum(1.0, np.abs(g["terms"][k])))
    assert np.all(worst < 1e-12), worst
    # the fixture exercises the gates: zero commands give stand_still > 0 and imitation == 0
    zero = np.linalg.norm(g["command"][:, :3], axis=1) < 0.01
    assert zero.any() and np.all(g["terms"][zero, 6] == 0) and np.all(g["terms"][zero, 4] > 0)


LIBRARY_TERMS = ["lin_vel_z", "ang_vel_xy", "base_height", "base_y_swing", "energy", "joint_pos_limits", "termination", "joint_deviation_hip",
                 "joint_deviation_knee", "pose", "feet_slip", "feet_clearance", "feet_height", "feet_air_time", "feet_phase"]


def test_reward_library_matches_numpy_twins(oracle):
    """The reward-library terms no shipped env wires in (rewards.py:37-90,120,152-241) against the reference's
    NumPy twins, argument for argument."""
    g = np.load(os.path.join(GOLD, "rewards_library.npz"))
    fn = oracle.lib.oduck_test_reward_library
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    nu = 14
    pad4 = lambda a: np.concatenate([[len(a)], a, np.zeros(4 - len(a))]).astype(np.float64)   # noqa: E731
    out, worst = np.zeros(15), np.zeros(15)
    for k in range(len(g["terms"])):
        vec = np.concatenate([
            g["global_linvel"][k], g["global_angvel"][k],
            [g["base_height"][k], g["base_height_target"][k], g["base_y_speed"][k], g["freq"][k], g["amplitude"][k], g["t"][k], g["tracking_sigma"][k]],
            g["qvel"][k], g["qfrc_actuator"][k], g["qpos"][k], g["soft_lowers"], g["soft_uppers"], [g["done"][k]], g["command"][k], g["default_pose"],
            pad4(g["hip_indices"]), pad4(g["knee_indices"]), g["weights"], g["contact"][k], g["feet_vel"][k].ravel(), g["foot_pos"][k].ravel(),
            [g["max_foot_height"][k]], g["swing_peak"][k], g["first_contact"][k], g["air_time"][k], [g["threshold_min"][k], g["threshold_max"][k]], g["rz"][k]])
        oracle.check(fn(nu, vec.ctypes.data, out.ctypes.data))
        worst = np.maximum(worst, np.abs(out - g["terms"][k]) / np.maximum(1.0, np.abs(g["terms"][k])))
    assert np.all(worst < 1e-12), dict(zip(LIBRARY_TERMS, worst))
    t = g["terms"]
    # every gate is hit from both sides in the vectors
    for name in ("termination", "joint_deviation_hip", "feet_slip", "feet_height", "feet_air_time"):
        col = t[:, LIBRARY_TERMS.index(name)]
        assert (col == 0).any() and (col != 0).any(), name
    assert (t[:, LIBRARY_TERMS.index("feet_air_time")] < 0).any()          # air time below threshold_min is a penalty (rewards.py:220)
    assert np.isclose(t[:, LIBRARY_TERMS.index("feet_air_time")].max(), 0.8) or t[:, LIBRARY_TERMS.index("feet_air_time")].max() < 0.8   # clip at max - min per foot
