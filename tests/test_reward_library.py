"""Reward-library terms switched on through ``reward_config.scales`` (include/oduck.h OduckRewardLibrary): the
oracle's env step adds exactly sum_k scale_k * term_k(inputs the reference's accessors name) to the task's own sum, with the
term functions pinned to the reference's NumPy twins (tests/test_golden.py), and the device code -- the RL instantiation of
k_step, run on CPU threads by tests/emu -- reproduces it."""
import ctypes as C

import numpy as np
import pytest
import torch

from open_duck_playground_b200 import capi, config as config_mod, mjcf, rng as jr
from open_duck_playground_b200.joystick import Joystick, default_config

LIB_SCALES = {"orientation": -1.3, "lin_vel_z": -0.7, "ang_vel_xy": -0.11, "base_height": -40.0, "energy": -0.02, "joint_pos_limits": -2.0, "termination": -3.0,
              "joint_deviation_hip": -0.5, "joint_deviation_knee": -0.3, "pose": -0.8, "feet_slip": -0.9, "feet_clearance": -6.0, "feet_height": -1.5, "feet_air_time": 4.0,
              "base_y_swing": 0.6, "feet_phase": 0.9}


def library_config():
    cfg = default_config()
    for k, v in LIB_SCALES.items():
        cfg.reward_config.scales[k] = v
    cfg.reward_config.base_height_target = 0.15
    cfg.reward_config.max_foot_height = 0.025
    cfg.reward_config.air_time_threshold_min = 0.04
    cfg.reward_config.air_time_threshold_max = 0.2
    cfg.reward_config.soft_joint_pos_limit_factor = 0.5            # tight enough for random actions to reach
    cfg.reward_config.pose_weights = [0.2 + 0.05 * i for i in range(14)]
    cfg.reward_config.base_y_swing_freq = 2.2
    cfg.reward_config.base_y_swing_amplitude = 0.07
    return cfg


def test_config_mapping(model_backlash):
    cs, _ = config_mod.build_env_config(model_backlash, library_config(), None, use_imitation_reward=False)
    assert [cs.lib.scale[i] for i in range(len(capi.LIB_TERMS))] == [LIB_SCALES[k] for k in capi.LIB_TERMS]
    assert cs.scale_orientation == 0.0                                # Joystick: orientation is a library term (joystick.py:645)
    assert cs.lib.n_hip == 4 and list(cs.lib.hip_indices) == [0, 1, 9, 10] and cs.lib.n_knee == 2 and list(cs.lib.knee_indices)[:2] == [3, 12]
    lo, hi = model_backlash.act_ctrlrange[:14, 0], model_backlash.act_ctrlrange[:14, 1]
    assert np.allclose([cs.lib.soft_lowers[i] for i in range(14)], 0.5 * (lo + hi) - 0.25 * (hi - lo))
    plain, _ = config_mod.build_env_config(model_backlash, default_config(), None, use_imitation_reward=False)
    assert all(plain.lib.scale[i] == 0.0 for i in range(len(capi.LIB_TERMS)))          # the shipped tasks use none
    bad = default_config()
    bad.reward_config.scales["feet_yaw"] = 1.0
    with pytest.raises(ValueError, match="unknown term"):
        config_mod.build_env_config(model_backlash, bad, None, use_imitation_reward=False)


def _library_sum(oracle, env, model, pre, contact_now, done, cfg):
    """sum_k scale_k * term_k for every env from buffers, term functions = the oracle's pinned RewardLibrary."""
    fn = oracle.lib.oduck_test_reward_library
    fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    A = model.arrays
    n, nu, dt = env.handle.n, model.nu, float(cfg.ctrl_dt)
    b = lambda k: env.buffer(k).numpy().astype(np.float64)          # noqa: E731
    qpos, qvel, sd, af, feet = b("QPOS"), b("QVEL"), b("SENSORDATA"), b("ACTUATOR_FORCE"), b("SITE_XPOS_FEET")
    cmd = pre["command"]
    jq = [int(A["jnt_qposadr"][A["act_jntid"][u]]) for u in range(nu)]
    jd = [int(A["jnt_dofadr"][A["act_jntid"][u]]) for u in range(nu)]
    lib, _ = config_mod.build_env_config(model, cfg, env.PRM)
    lib = lib.lib
    pad4 = lambda a, k: np.concatenate([[k], a[:4]]).astype(np.float64)       # noqa: E731
    imu = int(A["imu_site"])
    out, sums = np.zeros(15), np.zeros(n)
    order = ["lin_vel_z", "ang_vel_xy", "base_height", "base_y_swing", "energy", "joint_pos_limits", "termination", "joint_deviation_hip", "joint_deviation_knee", "pose",
             "feet_slip", "feet_clearance", "feet_height", "feet_air_time", "feet_phase"]      # columns of oduck_test_reward_library
    rzf = oracle.lib.oduck_test_gait_rz
    rzf.argtypes, rzf.restype = [C.c_double, C.c_double], C.c_double
    imit = env.buffer("INFO_IMITATION_I").numpy()                   # the gait clock (include/oduck.h): phase counter after this step's increment
    period = int(env.PRM.nb_steps_in_period)
    for i in range(n):
        # sensors belong to the last forward, i.e. to the pose BEFORE the last Euler step: undo its rotation of the base (the imu
        # site sits on the free body); global_linvel = site_xmat @ local_linvel (velocimeter = site_xmat.T @ framelinvel)
        w = qvel[i, 3:6]
        ang = np.linalg.norm(w) * float(A["timestep"])
        dq = np.concatenate([[np.cos(ang / 2)], -np.sin(ang / 2) * w / max(np.linalg.norm(w), 1e-300)])
        Rs = mjcf.quat_to_mat(mjcf.quat_mul(qpos[i, 3:7], dq)) @ mjcf.quat_to_mat(A["site_quat"][imu])
        glv = Rs @ sd[i, 3:6]
        first = (pre["air"][i] > 0) * ((contact_now[i] != 0) | (pre["lastc"][i] != 0))
        air = pre["air"][i] + dt
        swing = np.maximum(pre["swing"][i], feet[i].reshape(2, 3)[:, 2])
        vec = np.concatenate([
            glv, sd[i, 12:15], [qpos[i, 2], lib.base_height_target, sd[i, 4], lib.base_y_swing_freq, lib.base_y_swing_amplitude, imit[i] * dt, float(cfg.reward_config.tracking_sigma)],
            qvel[i, jd], af[i], qpos[i, jq],
            [lib.soft_lowers[u] for u in range(nu)], [lib.soft_uppers[u] for u in range(nu)], [done[i]], cmd[i], model.key_ctrl[:nu],
            pad4(np.array(list(lib.hip_indices), float), lib.n_hip), pad4(np.array(list(lib.knee_indices), float), lib.n_knee),
            [lib.pose_weights[u] for u in range(nu)], contact_now[i], sd[i, 15:21], feet[i], [lib.max_foot_height], swing, first.astype(float), air,
            [lib.air_time_threshold_min, lib.air_time_threshold_max],
            [rzf(((2 * np.pi * imit[i] / period + k * np.pi + np.pi) % (2 * np.pi)) - np.pi, lib.max_foot_height) for k in range(2)]])
        oracle.check(fn(nu, vec.ctypes.data, out.ctypes.data))
        s = LIB_SCALES["orientation"] * (sd[i, 9] ** 2 + sd[i, 10] ** 2)
        for col, name in enumerate(order):
            if name is not None:
                s += LIB_SCALES[name] * out[col]
        sums[i] = s
    return sums


def _drive(oracle, lib_a, lib_b, steps, n=12, seed=0):
    """Same seeds and actions through a plain env (lib_a) and one with the library terms (lib_b); yields per-step tuples."""
    cfg = library_config()
    plain, ext = Joystick("flat_terrain_backlash", library=lib_a), Joystick("flat_terrain_backlash", config=cfg, library=lib_b)
    keys = jr.split(jr.PRNGKey(7), n)
    sp, se = plain.reset(keys), ext.reset(keys)
    rs = np.random.default_rng(seed)
    for t in range(steps):
        pre = {"air": ext.buffer("INFO_FEET_AIR_TIME").numpy().astype(np.float64).copy(), "lastc": ext.buffer("INFO_LAST_CONTACT").numpy().copy(),
               "swing": ext.buffer("INFO_SWING_PEAK").numpy().astype(np.float64).copy(), "command": ext.buffer("INFO_COMMAND").numpy().astype(np.float64).copy()}
        act = torch.from_numpy(rs.uniform(-1, 1, (n, 14)).astype(np.float32))
        sp, se = plain.step(sp, act), ext.step(se, act)
        yield t, plain, ext, sp, se, pre, cfg


def test_library_terms_enter_the_oracle_reward(oracle, model_backlash):
    dt, seen = 0.02, 0
    for t, plain, ext, sp, se, pre, cfg in _drive(oracle, oracle, oracle, steps=14):
        assert torch.equal(sp.data.qpos, se.data.qpos) and torch.equal(sp.obs["state"], se.obs["state"])       # rewards do not touch the dynamics
        contact = ext.buffer("INFO_LAST_CONTACT").numpy().astype(np.float64)                                      # last_contact <- contact (joystick.py:461)
        # done as the env computed it (before the auto-reset copy): the Episode wrapper's flag, truncation excluded
        done = se.done.numpy().astype(np.float64) * (1 - se.info["truncation"].numpy().astype(np.float64))
        live = done == 0                                                                                          # auto-reset envs show the first state's buffers
        want = _library_sum(oracle, ext, model_backlash, pre, contact, done, cfg)
        base = sp.reward.numpy().astype(np.float64) / dt
        ok = live & (base > 0) & (base + want > 0)                                                                # neither side clipped at 0
        got = se.reward.numpy().astype(np.float64) / dt - base
        assert np.abs(got[ok] - want[ok]).max() < 1e-8 * max(1.0, np.abs(want[ok]).max()), (t, got[ok], want[ok])
        seen += int(ok.sum())
        assert np.all(se.reward.numpy()[live & (base + want <= 0)] == 0)                                          # clipped at 0 like joystick.py:447
    assert seen > 60


def test_device_code_adds_the_same_terms(oracle, model_backlash):
    """k_step<HF = false, RL = true> run on CPU threads (tests/emu) against the oracle with every library term switched on."""
    from test_env_emu import _rows_ok, _sync, build_emu_library
    emu_lib = build_emu_library()
    n, cfg = 8, library_config()
    emu, ref = Joystick("flat_terrain_backlash", config=cfg, library=emu_lib), Joystick("flat_terrain_backlash", config=cfg, library=oracle)
    plain = Joystick("flat_terrain_backlash", library=oracle)
    keys = jr.split(jr.PRNGKey(7), n)
    sg, sr, sp = emu.reset(keys), ref.reset(keys), plain.reset(keys)
    rs = np.random.default_rng(5)
    moved = 0.0
    for t in range(4):
        act = torch.from_numpy(rs.uniform(-1, 1, (n, 14)).astype(np.float32))
        _sync(emu, ref)
        sg, sr, sp = emu.step(sg, act), ref.step(sr, act), plain.step(sp, act)
        assert np.array_equal(emu.buffer("INFO_RNG").numpy(), ref.buffer("INFO_RNG").numpy())
        ok = _rows_ok(sg.data.qpos, sr.data.qpos, 1e-4) & _rows_ok(sg.data.qvel, sr.data.qvel, 2e-3, 1e-3)
        assert ok.mean() >= 0.85
        assert _rows_ok(sg.reward[:, None], sr.reward[:, None], 3e-4)[ok].all(), (t, sg.reward, sr.reward)
        moved = max(moved, float((sr.reward.double() - sp.reward.double()).abs().max()))
    assert moved > 0.02                                                                      # the library terms really changed the reward


def library_terms_gpu_parity(oracle):
    """CUDA k_step<HF = false, RL = true> against the oracle, every library term on (body of test_library_terms_gpu_parity;
    its first hardware runs happened in a child process -- xpassed on B200 in profiles/r02a..r02g_pytest_gpu.log)."""
    from test_parity_gpu import Checks, _sync_from_ref
    n, cfg = 256, library_config()
    gpu, ref = Joystick("flat_terrain_backlash", config=cfg, device="cuda:0"), Joystick("flat_terrain_backlash", config=cfg, library=oracle)
    for e in (gpu, ref):
        e.randomize(jr.split(jr.PRNGKey(11), n))
    keys = jr.split(jr.PRNGKey(0), n)
    sg, sr = gpu.reset(keys), ref.reset(keys)
    rs = np.random.default_rng(2)
    c = Checks()
    for t in range(6):
        act = rs.uniform(-1, 1, (n, 14)).astype(np.float32)
        _sync_from_ref(gpu, ref)
        sg, sr = gpu.step(sg, torch.from_numpy(act).cuda()), ref.step(sr, torch.from_numpy(act))
        torch.cuda.synchronize()
        c.equal(gpu.buffer("INFO_RNG").cpu().numpy(), ref.buffer("INFO_RNG").numpy(), f"[{t}] rng")
        c.close(sg.data.qpos, sr.data.qpos, 1e-4, what=f"[{t}] qpos")
        c.close(sg.reward, sr.reward, 3e-4, what=f"[{t}] reward")
    c.done()


@pytest.mark.gpu
def test_library_terms_gpu_parity(oracle):
    library_terms_gpu_parity(oracle)
