"""Drop-in check with the reference's OWN client: /root/reference/apps/demo.cpp is compiled UNMODIFIED against this repo's
include/ and linked against libtfusion.so (src/Makefile: apps/demo_reference).  OpenCV highgui / viz — absent from this image —
are replaced by the headless stubs under tests/cpp/stubs, which only move bytes: decode the PGM files demo.cpp:93-96 reads,
remember the last view it shows, log the pose it hands to the viewer, close the window after N frames.  CPU part: it compiles and
links where the reference tree exists.  GPU part: the binary (built here, it travels to the GPU box) runs the demo's loop over
synthetic frames and must produce exactly what the C ABI produces for the same frames."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DEMO = "/root/reference/apps/demo.cpp"
BIN = os.path.join(ROOT, "apps", "demo_reference")


def test_reference_demo_compiles_against_this_repos_headers():
    if not os.path.exists(REF_DEMO):
        pytest.skip("reference tree absent (GPU box)")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-std=c++17", "-fsyntax-only", "-Wall", "-I" + os.path.join(ROOT, "tests", "cpp", "stubs"),
                        "-I" + os.path.join(ROOT, "include"), REF_DEMO], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_reference_demo_links_against_libtfusion():
    if not os.path.exists(REF_DEMO):
        pytest.skip("reference tree absent (GPU box)")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "src"), "demo_reference"], capture_output=True, text=True)
    assert r.returncode == 0 and os.path.exists(BIN), r.stdout + r.stderr
    und = subprocess.run(["nm", "-uC", BIN], capture_output=True, text=True, check=True).stdout
    used = [l.split(" U ", 1)[1].strip() for l in und.splitlines() if " U tfusion::" in l]
    # every tfusion:: symbol the reference's client needs is one this repo's library defines
    defined = subprocess.run(["nm", "-DC", "--defined-only", os.path.join(ROOT, "topfusion_b200", "libtfusion.so")],
                             capture_output=True, text=True, check=True).stdout
    assert len(used) >= 10
    for sym in used:
        assert sym in defined, sym


@pytest.mark.gpu
def test_reference_demo_runs_on_this_repos_library(gpu, tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("apps/demo_reference was not built (needs /root/reference at build time)")
    from topfusion_b200 import synth
    n = 6
    depth, _, _ = synth.sequence("S1", n)
    for i in range(n):   # the path demo.cpp hard-codes (demo.cpp:93): on Linux a file NAME with backslashes, in the working directory
        synth.write_pgm(str(tmp_path / ("E:\\Teddy\\Frames\\%04d.pgm" % i)), depth[i])
    env = dict(os.environ, TFUSION_STUB_FRAMES=str(n), TFUSION_STUB_POSE_LOG=str(tmp_path / "poses.txt"),
               TFUSION_STUB_VIEW_OUT=str(tmp_path / "view.raw"))
    r = subprocess.run([BIN], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    poses = np.loadtxt(tmp_path / "poses.txt").reshape(-1, 4, 4).astype(np.float32)   # %.9g round-trips a float exactly
    assert len(poses) == n
    g = gpu.Context()          # TopFuParams::default_params() == tfb_default_params(): reference behaviour
    try:
        for i in range(n):
            g.process_frame(depth[i])
            assert np.array_equal(g.pose(), poses[i]), i
        view = np.fromfile(tmp_path / "view.raw", np.uint8).reshape(480, 640, 4)
        assert np.array_equal(view, g.render_image())
        assert view[..., 0].max() > 0
    finally:
        g.close()
