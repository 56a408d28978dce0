"""Constants for Open Duck Mini V2 (mirrors reference open_duck_mini_v2/constants.py).

The reference maps a task name to an MJCF file under its own checkout (constants.py:28-34).  This package ships
the compiled models as blobs (tools/compile_assets.py), so a task name resolves to both the reference's XML
*name* (for ``env.xml_path``) and the blob that is actually loaded.
"""
import os

ROOT_PATH = os.path.dirname(os.path.abspath(__file__))
DATA_PATH = os.path.join(ROOT_PATH, "data")
POLY_BLOB = os.path.join(DATA_PATH, "polynomial_coefficients.npz")

_TASK_XML = {
    "flat_terrain": "scene_flat_terrain.xml",
    "rough_terrain": "scene_rough_terrain.xml",  # named by the reference but the file does not exist there either
    "flat_terrain_backlash": "scene_flat_terrain_backlash.xml",
    "rough_terrain_backlash": "scene_rough_terrain_backlash.xml",
}


def task_to_xml(task_name: str) -> str:
    return os.path.join("xmls", _TASK_XML[task_name])  # KeyError on unknown task, like the reference


def task_to_blob(task_name: str) -> str:
    path = os.path.join(DATA_PATH, f"{task_name}.npz")
    if not os.path.exists(path):
        raise FileNotFoundError(f"no compiled model for task {task_name!r} ({path}); run tools/compile_assets.py")
    return path


FEET_SITES = ["left_foot", "right_foot"]
LEFT_FEET_GEOMS = ["left_foot_bottom_tpu"]
RIGHT_FEET_GEOMS = ["right_foot_bottom_tpu"]
FEET_GEOMS = LEFT_FEET_GEOMS + RIGHT_FEET_GEOMS
JOINTS_ORDER_NO_HEAD = [
    "left_hip_yaw", "left_hip_roll", "left_hip_pitch", "left_knee", "left_ankle",
    "right_hip_yaw", "right_hip_roll", "right_hip_pitch", "right_knee", "right_ankle",
]
ROOT_BODY = "trunk_assembly"
GRAVITY_SENSOR = "upvector"
GLOBAL_LINVEL_SENSOR = "global_linvel"
GLOBAL_ANGVEL_SENSOR = "global_angvel"
LOCAL_LINVEL_SENSOR = "local_linvel"
ACCELEROMETER_SENSOR = "accelerometer"
GYRO_SENSOR = "gyro"
