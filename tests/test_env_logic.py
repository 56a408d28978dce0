"""Host-visible semantics of Joystick.reset/step + the fused Episode/AutoReset wrappers, exercised on the CPU oracle
through the same Joystick shim the CUDA library uses (reference: open_duck_mini_v2/joystick.py:206-481)."""
import numpy as np
import pytest
import torch

from open_duck_playground_b200 import capi, config, rng as jr
from open_duck_playground_b200.joystick import Joystick


@pytest.fixture()
def env(oracle):
    return Joystick("flat_terrain_backlash", library=oracle)


def test_unknown_task_raises(oracle):
    with pytest.raises(KeyError):
        Joystick("no_such_task", library=oracle)


def test_surface(env):
    assert env.action_size == 14 and env.dt == 0.02 and env.sim_dt == 0.002 and env.n_substeps == 10
    assert env.observation_size == {"state": (101,), "privileged_state": (212,)}
    assert env.xml_path.endswith("scene_flat_terrain_backlash.xml") and env.unwrapped is env


def test_reset_is_deterministic_and_matches_reference_initial_info(env):
    keys = jr.split(jr.PRNGKey(3), 8)
    s1 = env.reset(keys).clone()
    s2 = env.reset(keys)
    assert torch.equal(s1.obs["state"], s2.obs["state"]) and torch.equal(s1.data.qpos, s2.data.qpos)
    info = s2.info
    assert torch.all(info["step"] == 0) and torch.all(info["imitation_i"] == 0)
    assert torch.all(info["imitation_phase"] == 0)                              # [0, 0] at reset, not [cos 0, sin 0]
    assert torch.all((info["push_interval_steps"] >= 250) & (info["push_interval_steps"] <= 500))
    assert torch.allclose(info["motor_targets"], torch.tensor(env.mj_model.key_ctrl[:14]).expand(8, 14))
    q = s2.data.qpos
    assert torch.all((q[:, :2].abs() <= 0.05)) and torch.allclose(q[:, 3:7].norm(dim=1), torch.ones(8, dtype=q.dtype))
    assert torch.all(q[:, 8] == 0)                                              # backlash joints are not perturbed
    assert torch.all(s2.reward == 0) and torch.all(s2.done == 0)
    cmd = info["command"]
    assert torch.all(cmd[:, 0].abs() <= 0.15) and torch.all(cmd[:, 1].abs() <= 0.2) and torch.all(cmd[:, 2].abs() <= 1.0)


def test_obs_layout(env):
    keys = jr.split(jr.PRNGKey(0), 4)
    st = env.reset(keys)
    act = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, (4, 14)).astype(np.float32))
    st = env.step(st, act)
    o, p = st.obs["state"], st.obs["privileged_state"]
    assert torch.equal(o, p[:, :101])
    assert torch.equal(o[:, 6:13], st.info["command"])
    assert torch.all(o[:, 41:83] == 0)                 # obs is built BEFORE info["last_act"] = action (joystick.py:437 vs :454)
    assert torch.allclose(st.info["last_act"].float(), act)
    assert torch.equal(o[:, 83:97], st.info["motor_targets"])
    assert torch.equal(o[:, 99:101], st.info["imitation_phase"])
    assert torch.equal(p[:, 101:104], st.data.sensordata[:, 0:3])                # noiseless gyro
    assert torch.equal(p[:, 130:144].float(), env.get_actuator_joints_qvel(st.data.qvel).float())
    assert torch.equal(p[:, 144], st.data.qpos[:, 2])
    assert torch.equal(p[:, 145:159], st.data.actuator_force)
    assert torch.equal(p[:, 169:209], st.info["current_reference_motion"])
    assert torch.all(p[:, 209] == 1)                                             # imitation_i after one step
    ph = 2 * np.pi / 27
    assert torch.allclose(p[:, 210:212], torch.tensor([np.cos(ph), np.sin(ph)], dtype=p.dtype).expand(4, 2), atol=1e-6)
    st = env.step(st, torch.zeros(4, 14))
    assert torch.allclose(st.obs["state"][:, 41:55].float(), act) and torch.all(st.obs["state"][:, 55:83] == 0)


def test_motor_speed_limit_and_delay(env):
    keys = jr.split(jr.PRNGKey(1), 16)
    st = env.reset(keys)
    prev = st.info["motor_targets"].clone()
    st = env.step(st, torch.ones(16, 14))
    lim = 5.24 * 0.02
    assert torch.all((st.info["motor_targets"] - prev).abs() <= lim + 1e-6)
    hist = st.info["action_history"].reshape(16, 3, 14)
    assert torch.all(hist[:, 0] == 1) and torch.all(hist[:, 1:] == 0)


def test_noise_level_zero_makes_obs_noiseless(oracle):
    env = Joystick("flat_terrain_backlash", library=oracle, config_overrides={"noise_config.level": 0.0})
    st = env.reset(jr.split(jr.PRNGKey(0), 2))
    assert torch.equal(st.obs["state"][:, 0:3], st.data.sensordata[:, 0:3])
    assert torch.equal(st.obs["state"][:, 3:6], st.data.sensordata[:, 6:9])      # no +1.3 accelerometer bias (reference quirk)


def test_episode_truncation_and_autoreset(oracle):
    env = Joystick("flat_terrain_backlash", library=oracle, config_overrides={"episode_length": 3})
    keys = jr.split(jr.PRNGKey(5), 2)
    st = env.reset(keys)
    first_q, first_obs = st.data.qpos.clone(), st.obs["state"].clone()
    z = torch.zeros(2, 14)
    for k in range(3):
        st = env.step(st, z)
        assert st.info["steps"].tolist() == [k + 1] * 2
    assert torch.all(st.done == 1) and torch.all(st.info["truncation"] == 1)
    assert torch.equal(st.data.qpos, first_q) and torch.equal(st.obs["state"], first_obs)    # data/obs <- first state
    assert st.info["step"].tolist() == [3, 3]                                              # info is NOT reset by the wrapper
    st = env.step(st, z)
    assert st.info["steps"].tolist() == [1, 1] and torch.all(st.done == 0)


def test_command_resample_after_500_steps(oracle):
    env = Joystick("flat_terrain_backlash", library=oracle)
    st = env.reset(jr.split(jr.PRNGKey(2), 1))
    cmd0 = st.info["command"].clone()
    z = torch.zeros(1, 14)
    for _ in range(500):
        st = env.step(st, z)
    assert torch.equal(st.info["command"], cmd0) and int(st.info["step"]) == 500
    st = env.step(st, z)
    assert int(st.info["step"]) == 0 and not torch.equal(st.info["command"], cmd0)


def test_domain_randomize_ranges(env):
    from open_duck_playground_b200.randomize import domain_randomize
    n = 64
    env.reset(jr.split(jr.PRNGKey(0), n))
    _, axes = domain_randomize(env, jr.split(jr.PRNGKey(9), n))
    assert "body_mass" in axes
    dr = env.buffer("DR_PARAMS").numpy()
    m = env.mj_model
    mass = dr[:, 1:1 + 20][:, : m.nbody]                     # oracle layout: friction0, body_mass[20], ...
    nominal = m.body_mass[: m.nbody]
    assert np.all(np.abs(mass[:, 1]) <= 0.1) and mass[:, 1].min() < 0 < mass[:, 1].max()      # massless base gets +-0.1 kg
    ratio = mass[:, 2:17] / nominal[2:17]
    assert ratio.min() >= 0.9 and ratio.max() <= 1.1 and ratio.std() > 0.01
    assert np.all((dr[:, 0] >= 0.5) & (dr[:, 0] <= 1.0))


def test_config_struct_mirrors_default_config(model_backlash, poly_table):
    c, _ = config.build_env_config(model_backlash, config.default_config(), poly_table)
    assert c.n_substeps == 10 and c.episode_length == 1000 and c.action_max_delay == 3
    assert list(c.qpos_noise_scale)[:14] == [0.03, 0.03, 0.03, 0.05, 0.08, 0.03, 0.03, 0.03, 0.05, 0.08, 0, 0, 0, 0]   # quirk #3
    assert c.scale_alive == 20.0 and c.scale_torques == -1e-3 and c.tracking_sigma == 0.01
    assert [list(r) for r in c.cmd_range][3] == [-0.34, 1.1]


def test_single_env_headless_loop_config1(oracle):
    """BASELINE configs[0]: flat_terrain, 1 env, CPU step loop (the mujoco_infer.py plumbing case): zero command,
    fixed-seed random MLP policy, 500 control steps headless, no NaN, and the robot keeps a sane height under ctrl = home."""
    from open_duck_playground_b200 import ppo
    env = Joystick("flat_terrain", library=oracle, config_overrides={"noise_config.level": 0.0, "push_config.enable": False})
    st = env.reset(jr.split(jr.PRNGKey(0), 1))
    assert env.mj_model.nq == 21
    for _ in range(50):                                      # ctrl = home for the first 50 steps
        st = env.step(st, torch.zeros(1, 14))
        assert 0.05 < float(st.data.qpos[0, 2]) < 0.25
    torch.manual_seed(0)
    pol = ppo.MLP([101, 512, 256, 128, 28])
    w = ppo.PolicyWeights(pol, 101, env.device)
    for _ in range(450):
        act, _, _ = ppo.policy_forward(env, w, None, deterministic=True)
        st = env.step(st, 0.2 * act)
    assert torch.isfinite(st.obs["privileged_state"]).all() and torch.isfinite(st.data.qpos).all()
    assert abs(float(st.data.qpos[0, 3:7].norm()) - 1) < 1e-9


@pytest.mark.parametrize("cls_name", ["Joystick", "Standing"])
def test_privileged_observation_starts_with_the_policy_observation(oracle, cls_name):
    """joystick.py:596-615 / standing.py: privileged_state = hstack([state, ...]).  The learner and the multi-GPU exchange rely on it
    (OduckRollout.obs_policy_ld: the policy rows are read out of the value rows), so it is asserted on reset and on noisy steps."""
    from open_duck_playground_b200.standing import Standing
    cls = {"Joystick": Joystick, "Standing": Standing}[cls_name]
    env = cls("flat_terrain_backlash", library=oracle)
    assert env.privileged_obs_has_state_prefix
    n = 6
    st = env.reset(jr.split(jr.PRNGKey(4), n))
    ds = env.observation_size["state"][0]
    assert torch.equal(st.obs["privileged_state"][:, :ds], st.obs["state"])
    g = torch.Generator().manual_seed(1)
    for _ in range(3):
        st = env.step(st, torch.rand(n, env.action_size, generator=g) * 2 - 1)
        assert torch.equal(st.obs["privileged_state"][:, :ds], st.obs["state"])
        assert st.obs["state"].abs().max().item() > 0
