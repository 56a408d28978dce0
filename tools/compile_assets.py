"""Compile the reference's model DATA (MJCF + foot STL + polynomial pickle) into the blobs shipped in
open_duck_playground_b200/data/.  Run in the build container where /root/reference is mounted:

    python tools/compile_assets.py [/root/reference]

The GPU box has no /root/reference, so the package loads these blobs by task name; a user who has the
reference checkout can instead point ``Joystick(xml_path=...)`` at the XML and get the same model."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from open_duck_playground_b200 import mjcf  # noqa: E402
from open_duck_playground_b200.poly_reference_motion import PolyTable  # noqa: E402

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
root = os.path.join(ref, "playground", "open_duck_mini_v2")
out = os.path.join(os.path.dirname(__file__), "..", "open_duck_playground_b200", "data")
os.makedirs(out, exist_ok=True)
for task, xml in {"flat_terrain": "scene_flat_terrain.xml", "flat_terrain_backlash": "scene_flat_terrain_backlash.xml",
                  "rough_terrain_backlash": "scene_rough_terrain_backlash.xml"}.items():
    m = mjcf.compile_mjcf(os.path.join(root, "xmls", xml))
    m.arrays.pop("hfield_file", None)
    m.save(os.path.join(out, f"{task}.npz"))
    print(task, "nq", m.nq, "nv", m.nv, "nu", m.nu)
t = PolyTable.from_pickle(os.path.join(root, "data", "polynomial_coefficients.pkl"))
t.save(os.path.join(out, "polynomial_coefficients.npz"))
print("poly table", t.coef.shape, t.nb_steps_in_period, t.dx_range, t.dy_range, t.dtheta_range)
