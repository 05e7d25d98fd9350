"""integration/nudge_gpu.h — the reference-side binding of INTEGRATION.md section 2 (resident state, nb_step) — driven by the headless
application loop of oracle/headless_example.cpp built with -DNB_RESIDENT (oracle/_ref/headless_resident, linked with libnudge_b200.so).
The GPU half ran once on a B200 with the round's last seconds of GPU time (passed: `profiles/r02_summary.md`)."""
import os, subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _need(*names):
    for n in names:
        if not os.path.exists(os.path.join(REFDIR, n)):
            pytest.skip("oracle/_ref/%s not built (needs /root/reference in the build container)" % n)


def test_binding_compiles_links_and_fails_loudly_without_a_gpu():
    """No CPU fallback anywhere: without a CUDA device the program built on the binding stops at create() with the library's message."""
    import torch
    _need("headless_resident")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test below")
    r = subprocess.run([os.path.join(REFDIR, "headless_resident"), "20", "20", "2", "4"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "nudge_b200:" in r.stderr and "world" not in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_resident_binding_application_loop_equals_the_reference():
    """The same application, the same scene: stepped by the reference's nudge.cpp on the CPU (headless_ref) and through the binding with the
    state resident on the GPU (headless_resident): identical transform hashes, contact and cache counts, one world and two concurrent worlds."""
    _need("headless_ref", "headless_resident")
    for args in (["300", "300", "120", "8"], ["150", "150", "60", "8", "2"]):
        ref = subprocess.run([os.path.join(REFDIR, "headless_ref")] + args, capture_output=True, text=True, timeout=600)
        gpu = subprocess.run([os.path.join(REFDIR, "headless_resident")] + args, capture_output=True, text=True, timeout=600)
        assert ref.returncode == 0 and gpu.returncode == 0, gpu.stdout + gpu.stderr
        worlds = lambda out: [l for l in out.splitlines() if l.startswith("world")]
        assert worlds(ref.stdout) == worlds(gpu.stdout) and len(worlds(ref.stdout)) == (2 if len(args) == 5 else 1), ref.stdout + gpu.stdout
