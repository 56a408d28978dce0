"""Host-side pieces of bench.py that need no GPU: the JSON line's shared ``config`` object, the L2 rotation rule, the issue-slot
ceiling derived from the committed ncu capture, and the reference arm's line (contract: same metric / unit / config as the B200 arm)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_env_sets_exceed_l2():
    for n in (256, 2048, 4096, 65536):
        sets = bench.n_env_sets(n)
        assert sets >= 3 and sets * n * bench.STATE_BYTES_PER_ENV > bench.L2_BYTES


def test_workload_config_is_shared_by_both_arms():
    c = bench.workload_config("flat_terrain_backlash", 4096, 2, 2)
    assert c["envs"] == 8192 and c["envs_per_gpu"] == 4096 and c["substeps_per_step"] == 10 and "configs[1]" in c["workload"]
    assert "configs[3]" in bench.workload_config("rough_terrain_backlash", 2048, 8, 2)["workload"]
    assert "model" not in c and "seq_len" not in c                       # no model keys (tier contract)


def test_issue_slot_ceiling_follows_the_committed_capture():
    """332.5 M warp-instructions per 4096-env launch (profiles/r02ai_k_step_raw.csv) over 148 x 4 issue slots per cycle."""
    d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    r = bench.issue_slot_ceiling("flat_terrain_backlash", 4096, 0.5632, 1927.8)
    assert r["bound"] == "issue" and r["warp_instructions_per_env_step"] == pytest.approx(d["k_step_inst_per_env"])
    inst = d["k_step_inst_per_env"] * 4096
    assert inst == pytest.approx(332529011, rel=1e-9)
    assert r["frac"] == pytest.approx(inst / (148 * 4 * 1927.8e6 * 0.5632e-3))
    # the capture itself: 60.6 % issue-active of the ACTIVE cycles, 1.093 M elapsed cycles -> 51 % of the elapsed issue slots
    assert 0.50 < r["frac"] < 0.53
    assert r["achieved"] / r["peak"] == pytest.approx(r["frac"])
    # half the envs in the same time = half the fraction; the height-field scene has its own count
    assert bench.issue_slot_ceiling("flat_terrain_backlash", 2048, 0.5632, 1927.8)["frac"] == pytest.approx(r["frac"] / 2)
    hf = bench.issue_slot_ceiling("rough_terrain_backlash", 4096, 1.681, 1953.7)
    assert hf["warp_instructions_per_env_step"] == pytest.approx(629561439 / 4096)
    assert bench.issue_slot_ceiling("flat_terrain_backlash", 4096, 0.0, 1965.0) is None
    assert bench.issue_slot_ceiling("flat_terrain_backlash", 4096, 0.5, None) is None


def test_reference_arm_prints_one_contract_line(oracle):
    """``bench.py --impl reference`` on a small sample: ONE JSON line on stdout with the B200 arm's metric / unit / config keys,
    ``impl: reference``, a cpu_baseline describing the run and an e2e object with zero copies."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--envs-per-gpu", "64", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["n_gpus"] == 1
    assert d["config"]["envs_per_gpu"] == 64 and d["config"]["task"] == "flat_terrain_backlash"
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_idle_ranks_exit_quietly():
    """Under torchrun rank 0 alone runs the reference arm; the other ranks print nothing and exit 0."""
    env = dict(os.environ, PYTHONPATH=ROOT, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
