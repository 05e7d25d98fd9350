import os
import numpy as np
from nudge_b200 import scenes as S, abi

HERE = os.path.dirname(os.path.abspath(__file__))


def load_box_cases():
    return np.load(os.path.join(HERE, "golden", "box_box_cases.npz"))


def scene_of(g, k):
    s = S.Scene(2, 2, 0)
    s.box_tags[:] = (0, 1)
    s.box_data["size"][:] = g["size"][k]
    s.box_transforms["position"][:] = g["cpos"][k]
    s.box_transforms["rotation"][:] = g["crot"][k]
    s.box_transforms["body"][:] = g["cbody"][k]
    s.transforms["position"][1] = g["bpos"][k]; s.transforms["rotation"][1] = g["brot"][k]
    return s


def check_case(g, k, view):
    """view: contacts_view() of a widened implementation.  Returns an error string or None."""
    m = int(g["count"][k])
    if view["count"] != m:
        return "case %d family %d: %d contacts, reference has %d" % (k, g["family"][k], view["count"], m)
    if m != int(g["expect"][k]):
        return "case %d: fixture disagrees with the reference test's expected count" % k
    if not np.array_equal(view["data"].view(np.uint8), g["contacts"][k, :m].view(np.uint8)):
        return "case %d family %d: contact data differs" % (k, g["family"][k])
    if not (np.array_equal(view["bodies"]["a"], g["bodies"][k, :m, 0]) and np.array_equal(view["bodies"]["b"], g["bodies"][k, :m, 1])):
        return "case %d: bodies differ" % k
    if not np.array_equal(abi.wide_tag_to_ref(view["tags"], view["features"]), g["tags"][k, :m]):
        return "case %d family %d: tags differ" % (k, g["family"][k])
    return None


# ---- the reference's own demo application, recorded headless (tests/golden/make_demo_golden.py, oracle/demo_capture.cpp) ----
def load_demo_frames():
    return np.load(os.path.join(HERE, "golden", "demo_frames.npz"))


def demo_scene(g, state="initial"):
    """The demo's scene (example/main.cpp:391-432: ground + 1024 boxes + 512 spheres from rand()) with the body state of `state`
    ("initial" = before the first simulate(), "f0", "f40" = recorded frames) and the demo's step parameters (example/main.cpp:274-305)."""
    nb, nbox, nsph = len(g["initial_transforms"]), len(g["initial_box_tags"]), len(g["initial_sphere_tags"])
    s = S.Scene(nb, nbox, nsph)
    s.transforms[:] = np.ascontiguousarray(g[state + "_transforms"]).view(S.TRANSFORM).reshape(nb)
    s.momentum[:] = np.ascontiguousarray(g[state + "_momentum"]).view(S.MOMENTUM).reshape(nb)
    s.idle[:] = g[state + "_idle"]
    s.properties[:] = np.ascontiguousarray(g["initial_properties"]).view(S.PROPERTIES).reshape(nb)
    s.box_transforms[:] = np.ascontiguousarray(g["initial_box_transforms"]).view(S.TRANSFORM).reshape(nbox)
    s.box_data[:] = np.ascontiguousarray(g["initial_box_data"]).view(S.BOX).reshape(nbox)
    s.box_tags[:] = g["initial_box_tags"]
    s.sphere_transforms[:] = np.ascontiguousarray(g["initial_sphere_transforms"]).view(S.TRANSFORM).reshape(nsph)
    s.sphere_data["radius"][:] = g["initial_sphere_data"]
    s.sphere_tags[:] = g["initial_sphere_tags"]
    s.time_step = np.float32(1.0 / (60.0 * 2.0)); s.iterations = 20; s.gravity = np.float32(9.82); s.damping = np.float32(0.25)
    s.name = "reference demo (recorded)"
    return s


def demo_substeps(frame):
    """simulate() calls behind recorded frame f: f + 1, two sub-steps each (example/main.cpp:275, 330-334)."""
    return 2 * (frame + 1)
