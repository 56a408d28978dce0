"""Standing task for Open Duck Mini V2 -- drop-in surface over the B200 step library.

Mirrors the reference env open_duck_mini_v2/standing.py:103-661: the same physics, reset and step skeleton as Joystick with
  * no imitation reward / reference motion (``USE_IMITATION_REWARD = False``, standing.py:42) and no motor speed limits,
  * rewards ``orientation, torques, action_rate, alive, stand_still(ignore_head=True), head_pos`` (standing.py:573-606),
  * observations without ``motor_targets`` / ``imitation_phase``: ``state`` is 85 wide, ``privileged_state`` 153 (standing.py:526-566),
  * base velocity noise of +-0.5 at reset (standing.py:247), ``motor_targets`` starting at zero (standing.py:279),
  * commands with zero velocity part (standing.py:648-655) and a wider head-yaw range.
All of it runs in the same fused kernels (``OduckEnvConfig.task = ODUCK_TASK_STANDING``).
"""
from __future__ import annotations

from . import capi
from .config import ConfigDict, standing_default_config
from .joystick import Data, Joystick, State  # noqa: F401  (State / Data are the same records)


def default_config() -> ConfigDict:
    return standing_default_config()


class Standing(Joystick):
    """Standing policy (batched; reference class: standing.py:103)."""

    TASK = capi.TASK_STANDING
    METRICS = capi.METRIC_NAMES_STANDING
    DEFAULT_CONFIG = staticmethod(standing_default_config)
