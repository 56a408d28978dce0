"""open_duck_playground_b200: B200-native batched physics + rollout path for the Open Duck Mini V2 joystick task."""
from .config import default_config  # noqa: F401

__all__ = ["default_config"]
