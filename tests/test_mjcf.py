"""The MJCF-subset compiler against the model facts read off the reference XMLs."""
import os

import numpy as np
import pytest

from open_duck_playground_b200 import capi, constants, mjcf
from open_duck_playground_b200.mjcf import CompiledModel


@pytest.mark.parametrize("task,nq,nv,njnt", [("flat_terrain_backlash", 31, 30, 25), ("flat_terrain", 21, 20, 15), ("rough_terrain_backlash", 31, 30, 25)])
def test_sizes(task, nq, nv, njnt):
    m = CompiledModel.load(constants.task_to_blob(task))
    assert (m.nbody, m.njnt, m.nq, m.nv, m.nu, m.nsite) == (18, njnt, nq, nv, 14, 5)
    assert abs(m.body_mass[: m.nbody].sum() - 2.1071407) < 1e-6
    assert m.body_mass[1] == 0.0 and m.body_names[1] == "base" and m.body_names[17] == "floor"   # TORSO_BODY_ID = 1 is massless
    assert m.geom_names.index("floor") == 46 and len(m.geom_names) == 47                         # FLOOR_GEOM_ID = 0 is a visual mesh
    assert int(m.foot_nvert) == 17 and int(m.foot_nface) == 30


def test_backlash_layout(model_backlash):
    m = model_backlash
    act_q = [int(m.jnt_qposadr[m.act_jntid[u]]) for u in range(m.nu)]
    assert act_q == [7, 9, 11, 13, 15, 17, 18, 19, 20, 21, 23, 25, 27, 29]
    assert m.actuator_names[:5] == ["left_hip_yaw", "left_hip_roll", "left_hip_pitch", "left_knee", "left_ankle"]
    assert np.allclose(m.act_kp[: m.nu], 17.11) and np.allclose(m.act_forcerange[: m.nu], [-3.23, 3.23])
    assert np.allclose(m.act_ctrlrange[0], m.jnt_range[1])                     # inheritrange = 1
    assert m.jnt_limited[: m.njnt].tolist() == [0] + [1] * 24                   # autolimits
    fl = m.dof_frictionloss[: m.nv]
    assert (fl > 0).sum() == 14 and np.allclose(fl[fl > 0], 0.068)
    assert np.allclose(m.key_qpos[:7], [0, 0, 0.15, 1, 0, 0, 0]) and np.allclose(m.qpos0[:7], [0, 0, 0.22, 1, 0, 0, 0])
    assert float(m.floor_friction) == 0.6 and float(m.foot_friction) == 1.0 and int(m.enable_foot_foot) == 1
    assert m.dof_parentid[: m.nv].tolist()[:8] == [-1, 0, 1, 2, 3, 4, 5, 6] and m.dof_parentid[16] == 5 and m.dof_parentid[20] == 5


def test_foot_is_flat_at_qpos0(model_backlash):
    m = model_backlash
    xpos, xmat, _, _ = mjcf.world_kinematics(m, m.qpos0[: m.nq])
    for k in range(2):
        b = int(m.foot_body[k])
        z = (xpos[b] + m.foot_vert[k, :17] @ xmat[b].T)[:, 2]
        assert np.sum(np.abs(z - z.min()) < 1e-4) >= 4      # the sole is (to 0.1 mm) a flat face parallel to the floor


def test_mass_matrix_spd_and_invweight(model_backlash):
    m = model_backlash
    M = mjcf.mass_matrix(m, m.key_qpos[: m.nq])
    assert np.allclose(M, M.T) and np.linalg.eigvalsh(M).min() > 0
    M0 = mjcf.mass_matrix(m, m.qpos0[: m.nq])
    assert abs(np.trace(M0) / m.nv - float(m.meaninertia)) < 1e-12
    assert np.allclose(m.dof_invweight0[:3], np.diag(np.linalg.inv(M0))[:3].mean())
    assert np.all(m.body_invweight0[1:17, 0] > 0) and m.body_invweight0[17, 0] == 0   # floor is static


def test_struct_roundtrip(model_backlash):
    s = capi.model_to_struct(model_backlash)
    assert s.nq == 31 and s.abi_version == capi.ABI_VERSION
    assert abs(s.body_mass[2] - 0.698526) < 1e-12 and s.foot_nvert == 17


@pytest.mark.skipif(not os.path.exists("/root/reference/playground/open_duck_mini_v2/xmls/scene_flat_terrain_backlash.xml"), reason="reference checkout not mounted")
def test_blob_matches_fresh_compile(model_backlash):
    fresh = mjcf.compile_mjcf("/root/reference/playground/open_duck_mini_v2/xmls/scene_flat_terrain_backlash.xml")
    for k, v in model_backlash.arrays.items():
        assert np.array_equal(np.asarray(v), np.asarray(fresh.arrays[k])), k
