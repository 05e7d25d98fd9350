"""The C-ABI library loads without a GPU and exports every symbol include/nudge_b200.h declares; creating a context
without a GPU fails loudly (no CPU fallback).  CPU only."""
import ctypes, os, re
import numpy as np
import nudge_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nudge_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|void\*?|uint64_t|const char\*)\s+(nb_[a-z0-9_]+)\s*\(", text, re.M)))


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(nudge_b200.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(nudge_b200.EXPORTS) <= set(names)


def test_struct_layouts_match_reference_sizes():
    from nudge_b200 import scenes as S
    assert S.TRANSFORM.itemsize == 32 and S.PROPERTIES.itemsize == 16 and S.MOMENTUM.itemsize == 32  # nudge.h:34-50
    assert S.BOX.itemsize == 16 and S.SPHERE.itemsize == 4 and S.CONTACT.itemsize == 32 and S.IMPULSE.itemsize == 16  # nudge.h:52-66, 109-112


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    from nudge_b200 import scenes
    try:
        nudge_b200.Sim(scenes.two_boxes())
    except nudge_b200.NudgeError as e:
        assert "nb_create" in str(e)
    else:
        raise AssertionError("Sim() succeeded without a GPU")


def test_host_lut_calibration_model_is_exact_here():
    """rcpps / rsqrtps of the host CPU depend only on the top mantissa bits; the library checks this at nb_create."""
    lib = ctypes.CDLL(nudge_b200.LIB_PATH)
    rcp = np.zeros(2048, np.uint32); rsq = np.zeros(2048, np.uint32)
    lib.nb_host_sample_luts(rcp.ctypes.data_as(ctypes.c_void_p), rsq.ctypes.data_as(ctypes.c_void_p))
    assert lib.nb_host_check_lut_model(rcp.ctypes.data_as(ctypes.c_void_p), rsq.ctypes.data_as(ctypes.c_void_p)) == 1
    assert (rcp & 0x7ff).max() == 0 and (rsq & 0x7ff).max() == 0  # 12 significant mantissa bits
