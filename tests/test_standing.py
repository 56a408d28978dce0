"""Standing task (reference open_duck_mini_v2/standing.py): the oracle against the reference's NumPy reward
twins (golden vectors), the env semantics that differ from Joystick, and CUDA-vs-oracle parity through the C-ABI."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import make_handle
from open_duck_playground_b200 import capi, config, rng as jr
from open_duck_playground_b200.standing import Standing, default_config

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_standing_reward_terms_match_numpy_twins(oracle, model_backlash, poly_table):
    g = np.load(os.path.join(GOLD, "rewards_standing.npz"))
    h = make_handle(oracle, model_backlash, poly_table, 1)
    fn = oracle.lib.oduck_test_rewards_standing
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    out = np.zeros(6)
    worst = np.zeros(6)
    for k in range(len(g["command"])):
        vec = np.concatenate([g["command"][k], g["upvector"][k], g["actuator_force"][k], g["action"][k], g["last_act"][k], g["q"][k], g["qd"][k]])
        oracle.check(fn(h.h, vec.ctypes.data, out.ctypes.data))
        worst = np.maximum(worst, np.abs(out - g["terms"][k]) / np.maximum(1.0, np.abs(g["terms"][k])))
    assert np.all(worst < 1e-12), worst
    zero = np.linalg.norm(g["command"][:, :3], axis=1) < 0.01
    # gates: stand_still only without a velocity command, head_pos only with one (rewards.py:117,147)
    assert zero.any() and (~zero).any()
    assert np.all(g["terms"][zero, 4] > 0) and np.all(g["terms"][zero, 5] == 0) and np.all(g["terms"][~zero, 4] == 0)


def test_standing_config_mirrors_reference():
    c = default_config()
    assert c.reward_config.scales == {"orientation": -0.5, "torques": -1.0e-3, "action_rate": -0.375, "stand_still": -0.3, "alive": 20.0, "head_pos": -2.0}
    assert c.noise_config.scales.gyro == 0.05 and c.noise_config.scales.accelerometer == 0.005 and c.head_yaw_range == [-2.7, 2.7]
    assert "max_motor_velocity" not in c and "lin_vel_x" not in c


def test_standing_struct(model_backlash):
    cs, _ = config.build_env_config(model_backlash, default_config(), None, task=capi.TASK_STANDING)
    assert cs.task == 1 and cs.use_imitation_reward == 0 and cs.use_motor_speed_limits == 0
    assert cs.reset_base_qvel_noise == 0.5 and cs.scale_orientation == -0.5 and cs.scale_head_pos == -2.0 and cs.scale_imitation == 0.0
    assert [cs.cmd_range[i][1] for i in range(3)] == [0.0, 0.0, 0.0] and cs.cmd_range[5][1] == 2.7


@pytest.fixture()
def env(oracle):
    e = Standing("flat_terrain_backlash", library=oracle)
    e.reset(jr.split(jr.PRNGKey(3), 16))
    return e


def test_standing_surface_and_reset(env):
    assert env.observation_size == {"state": (85,), "privileged_state": (153,)} and env.action_size == 14
    st = env._state()
    assert st.obs["state"].shape == (16, 85) and st.obs["privileged_state"].shape == (16, 153)
    assert list(st.metrics) == capi.METRIC_NAMES_STANDING
    assert torch.all(st.info["motor_targets"] == 0)                     # standing.py:279
    assert torch.all(st.info["command"][:, :3] == 0)                    # standing.py:648-655
    assert torch.all(st.info["current_reference_motion"] == 0) and torch.all(st.info["imitation_phase"] == 0)
    bq = st.data.qvel[:, :6].abs()
    assert float(bq.max()) <= 0.5 and float(bq.max()) > 0.05            # standing.py:247: U(-0.5, 0.5)
    # obs layout (standing.py:526-542): gyro3 accel3 cmd7 q14 qd14 last_act 3x14 contact2
    s = st.obs["state"]
    assert torch.equal(s[:, 6:13].float(), st.info["command"].float())
    assert torch.all(s[:, 41:83] == 0)                                  # three zero action histories right after reset
    p = st.obs["privileged_state"]
    assert torch.equal(p[:, :85], s) and torch.allclose(p[:, 85 + 15 + 28].double(), st.data.qpos[:, 2].double())   # root height slot


def test_standing_step_semantics(env, oracle):
    n = 16
    rs = np.random.default_rng(0)
    st = env._state()
    for t in range(4):
        act = torch.from_numpy(rs.uniform(-1, 1, (n, 14)).astype(np.float32))
        st = env.step(st, act)
    assert torch.isfinite(st.reward).all() and float(st.reward.min()) >= 0
    # no speed limit: motor target = default + delayed action * scale for SOME delay of the history (standing.py:378-381)
    tg, hist = st.info["motor_targets"].double(), st.info["action_history"].double().reshape(n, 3, 14)
    dflt = torch.tensor(env._mj_model.key_ctrl[:14])
    cand = dflt + 0.25 * hist
    assert torch.all(((cand - tg[:, None]).abs().max(-1).values < 1e-6).any(-1))
    # metrics: costs positive, alive = 20 * 1, head_pos gated off because the command has no velocity part
    m = st.metrics
    assert torch.allclose(m["reward/alive"].double(), torch.full((n,), 20.0, dtype=torch.float64))
    assert torch.all(m["cost/head_pos"] == 0) and torch.all(m["cost/orientation"] >= 0) and torch.all(m["cost/stand_still"] > 0)
    # reward = clip(dt * sum(scaled terms)) with the documented signs
    tot = (m["reward/alive"] - m["cost/orientation"] - m["cost/torques"] - m["cost/action_rate"] - m["cost/stand_still"] - m["cost/head_pos"]).double() * 0.02
    assert torch.allclose(st.reward.double(), tot.clamp(0, 1e4), atol=1e-6)


@pytest.mark.gpu
def test_standing_gpu_parity(oracle):
    from test_parity_gpu import Checks, _np, _sync_from_ref
    n = 256
    gpu, ref = Standing("flat_terrain_backlash", device="cuda:0"), Standing("flat_terrain_backlash", library=oracle)
    for e in (gpu, ref):
        e.randomize(jr.split(jr.PRNGKey(11), n))
    keys = jr.split(jr.PRNGKey(0), n)
    sg, sr = gpu.reset(keys), ref.reset(keys)
    torch.cuda.synchronize()
    c = Checks()
    assert sg.obs["state"].shape == (n, 85) and sg.obs["privileged_state"].shape == (n, 153)
    c.close(sg.data.qvel, sr.data.qvel, 1e-6, what="reset qvel")
    c.rows(sg.obs["state"], sr.obs["state"], 2e-3, 2e-3, what="reset obs state")
    c.rows(sg.obs["privileged_state"], sr.obs["privileged_state"], 2e-3, 2e-3, what="reset obs privileged")
    rs = np.random.default_rng(2)
    for t in range(6):
        act = rs.uniform(-1, 1, (n, 14)).astype(np.float32)
        _sync_from_ref(gpu, ref)
        sg, sr = gpu.step(sg, torch.from_numpy(act).cuda()), ref.step(sr, torch.from_numpy(act))
        torch.cuda.synchronize()
        for name in ("INFO_RNG", "INFO_STEP", "INFO_STEPS", "INFO_PUSH_STEP", "INFO_IMITATION_I"):
            c.equal(gpu.buffer(name).cpu().numpy(), ref.buffer(name).numpy(), f"[{t}] {name}")
        c.close(sg.data.qpos, sr.data.qpos, 1e-4, what=f"[{t}] qpos")
        c.rows(sg.data.qvel, sr.data.qvel, 2e-3, 1e-3, what=f"[{t}] qvel")
        c.close(sg.reward, sr.reward, 2e-4, what=f"[{t}] reward")
        c.mostly_equal(_np(sg.done), _np(sr.done), f"[{t}] done")
        c.rows(gpu.buffer("METRICS"), ref.buffer("METRICS"), 1e-3, 2e-3, what=f"[{t}] metrics")
        c.rows(sg.obs["state"], sr.obs["state"], 2e-3, 2e-3, what=f"[{t}] obs state")
        c.rows(sg.obs["privileged_state"], sr.obs["privileged_state"], 2e-3, 2e-3, what=f"[{t}] obs privileged")
        for k in ("command", "motor_targets", "action_history", "feet_air_time", "last_contact", "push"):
            c.close(sg.info[k], sr.info[k], 1e-5, what=f"[{t}] {k}")
    c.done()


@pytest.mark.gpu
def test_standing_policy_forward_reads_the_strided_obs(oracle):
    """The actor reads the handle's own obs records (85 wide, 101 apart) -- same result as a contiguous copy."""
    from open_duck_playground_b200 import ppo
    n = 64
    gpu = Standing("flat_terrain_backlash", device="cuda:0")
    st = gpu.reset(jr.split(jr.PRNGKey(1), n))
    torch.manual_seed(0)
    pol = ppo.MLP([85, 512, 256, 128, 28]).cuda()
    w = ppo.PolicyWeights(pol, 85, gpu.device)
    w.refresh(torch.zeros(85, device="cuda"), torch.ones(85, device="cuda"))
    a1, r1, _ = ppo.policy_forward(gpu, w, None, True)
    a2, r2, _ = ppo.policy_forward(gpu, w, None, True, obs=st.obs["state"].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(r1, r2)
    ref = pol(st.obs["state"].contiguous())[:, :14]
    assert torch.allclose(r1, ref, atol=5e-5)
