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
