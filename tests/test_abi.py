"""The C-ABI library loads without a GPU and exports every symbol include/nudge_b200.h declares; creating a context
without a GPU fails loudly (no CPU fallback).  CPU only."""
import ctypes, os, re
import pytest
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


def test_state_file_header_is_read_without_a_gpu(tmp_path):
    """nb_state_info parses the header of a state file (layout: nudge_b200/csrc/nb_state_api.cuh) — pure host code, so tools can size an
    nb_config before any context exists; a file that is not a state file is refused."""
    import struct
    lib = nudge_b200.load_library()
    good = tmp_path / "s.bin"
    good.write_bytes(b"NBSTATE1" + struct.pack("<15I", 1, 1000, 600, 400, 7, 12345, 0, *([0] * 8)))
    counts = (ctypes.c_uint32 * 5)()
    assert lib.nb_state_info(os.fsencode(str(good)), counts) == 0
    assert list(counts) == [1000, 600, 400, 7, 12345]
    bad = tmp_path / "t.bin"
    bad.write_bytes(b"not a state file at all" + bytes(64))
    assert lib.nb_state_info(os.fsencode(str(bad)), counts) != 0
    assert lib.nb_state_info(os.fsencode(str(tmp_path / "missing.bin")), counts) != 0


def test_replay_tool_rejects_bad_files_and_has_no_cpu_path():
    """tools/bin/nb_replay: not a state file -> message and exit code 1; a valid file without a CUDA device -> the library's
    "no CPU path" error, not a silent fallback."""
    import struct, subprocess, tempfile
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "bin", "nb_replay")
    if not os.path.exists(exe):
        pytest.skip("tools/bin/nb_replay not built")
    r = subprocess.run([exe, "/nonexistent/state.bin", "--steps", "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "not an nb state file" in (r.stdout + r.stderr)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "one_body.bin")
        with open(p, "wb") as f:
            f.write(b"NBSTATE1" + struct.pack("<15I", 1, 1, 0, 0, 0, 0, 0, *([0] * 8)) + b"\0" * 96)   # one body (the static world), no colliders
        r = subprocess.run([exe, p, "--steps", "1"], capture_output=True, text=True, timeout=60)
        assert r.returncode != 0 and "hash" not in r.stdout, r.stdout + r.stderr


def test_shipped_library_contains_the_blackwell_instructions_the_design_claims():
    """cuobjdump -sass of nudge_b200/lib/libnudge_b200.so (no GPU needed): the throughput solver stages its rows with bulk copies and
    mbarriers (UBLKCP / SYNCS) and reduces with vector REDG; the exact-order solver hands rows over with 256-bit strong accesses
    (LDG/STG.E.ENL2.256.STRONG.GPU, sm_100 only); the sharded solver reads / writes peer inboxes with SYS-scope 128-bit accesses."""
    import shutil, subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    lib = os.path.join(ROOT, "nudge_b200", "lib", "libnudge_b200.so")
    if not os.path.exists(exe) or not os.path.exists(lib):
        pytest.skip("cuobjdump or the built library is missing")
    sass = subprocess.run([exe, "-sass", lib], capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in sass
    funcs, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); funcs[cur] = []
        elif cur:
            funcs[cur].append(line)
    body = lambda key: "\n".join("\n".join(v) for k, v in funcs.items() if key in k)
    jac = body("k_jacobi_sweepILb0ELi2E")
    assert "UBLKCP" in jac and "SYNCS.ARRIVE.TRANS64" in jac and "SYNCS.PHASECHK" in jac and "REDG.E.ADD.F32x4" in jac
    wide = body("k_solveILb1E")
    assert "LDG.E.ENL2.256.STRONG.GPU" in wide and "STG.E.ENL2.256.STRONG.GPU" in wide and "NANOSLEEP" in wide
    narrow = body("k_solveILb0E")
    assert "LDG.E.128.STRONG.GPU" in narrow and "ENL2.256" not in narrow
    flow = body("k_solve_flowILb0E")
    assert "LDG.E.128.STRONG.SYS" in flow and "STG.E.128.STRONG.SYS" in flow


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under nudge_b200/, include/, integration/ or tools/ may import, link or execute it, and bench.py
    only in its CPU legs (cpu_baseline_sample, dropin_sample, run_reference)."""
    import ast
    for d in ("nudge_b200", "include", "integration", "tools"):
        for base, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".sh")):
                    text = open(os.path.join(base, f), errors="replace").read()
                    for bad in ("import oracle", "from oracle", "oracle/", "liboracle", "pyoracle", "pyref", "libnudge_ref"):
                        assert bad not in text, "%s mentions %r" % (os.path.join(base, f), bad)
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    allowed = {"cpu_baseline_sample", "dropin_sample", "run_reference"}
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = any((isinstance(n, ast.ImportFrom) and (n.module or "").startswith("oracle")) or (isinstance(n, ast.Import) and any(a.name.startswith("oracle") for a in n.names)) or
                   (isinstance(n, ast.Constant) and n.value == "oracle") for n in ast.walk(fn))
        if uses:
            assert fn.name in allowed, "bench.py:%s uses the oracle" % fn.name
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert not any((getattr(n, "module", None) or "").startswith("oracle") or any(a.name.startswith("oracle") for a in n.names) for n in top)
