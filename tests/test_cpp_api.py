"""The C++ `tfusion` mirror (include/tfusion/*.hpp + src/*.cpp -> libtfusion.so) keeps the reference's public names
so apps/demo.cpp can link against it; apps/demo_synth issues the same call sequence headlessly."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "topfusion_b200", "libtfusion.so")
DEMO = os.path.join(ROOT, "apps", "demo_synth")


def _built():
    if not (os.path.exists(LIB) and os.path.exists(DEMO)):
        import __graft_entry__ as g
        g.build()


def test_public_symbols_of_the_reference_api_are_exported():
    _built()
    out = subprocess.run(["nm", "-DC", "--defined-only", LIB], capture_output=True, text=True, check=True).stdout
    # every call apps/demo.cpp makes + the rest of topfu.hpp
    for sym in ("tfusion::TopFuParams::default_params()", "tfusion::TopFu::TopFu(tfusion::TopFuParams const&)",
                "tfusion::TopFu::operator()(", "tfusion::TopFu::renderImage(", "tfusion::TopFu::getCameraPose(int) const",
                "tfusion::TopFu::reset()", "tfusion::TopFu::params()", "tfusion::TopFu::icp()",
                "tfusion::cuda::setDevice(int)", "tfusion::cuda::printShortCudaDeviceInfo(int)", "tfusion::cuda::checkIfPreFermiGPU(int)",
                "tfusion::cuda::getCudaEnabledDeviceCount()", "tfusion::cuda::getDeviceName",
                "tfusion::cuda::DeviceMemory2D::upload(", "tfusion::cuda::DeviceMemory2D::download(",
                "tfusion::cuda::ProjectiveICP::setIterationsNum(", "tfusion::cuda::ProjectiveICP::estimateTransform(",
                "tfusion::cuda::depthBilateralFilter(", "tfusion::cuda::computePointNormals(", "tfusion::cuda::resizePointsNormals(",
                "tfusion::SampledScopeTime::SampledScopeTime(double&)", "tfusion::OpenNISource::open(int)",
                # additions in front of / behind the path
                "tfusion::TopFu::operator()(tfusion::io::HostFrame const&)", "tfusion::io::FrameRing::next()",
                "tfusion::io::readPgm16(", "tfusion::TopFu::extractPoints(", "tfusion::TopFu::saveScene("):
        assert sym in out, sym


def test_params_struct_keeps_the_reference_field_order():
    src = open(os.path.join(ROOT, "include", "tfusion", "topfu.hpp")).read()
    body = src[src.index("struct KF_EXPORTS TopFuParams"):src.index("SceneParams* sceneParams;")]
    order = ["cols", "rows", "intr", "volume_dims", "volume_size", "volume_pose", "bilateral_sigma_depth", "bilateral_sigma_spatial",
             "bilateral_kernel_size", "icp_truncate_depth_dist", "icp_dist_thres", "icp_angle_thres", "icp_iter_num",
             "tsdf_min_camera_movement", "tsdf_trunc_dist", "tsdf_max_weight", "raycast_step_factor", "gradient_delta_factor", "light_pose"]
    pos = [body.index(" " + n + ";") for n in order]      # reference: include/tfusion/topfu.hpp:28-60
    assert pos == sorted(pos)


@pytest.mark.gpu
def test_demo_synth_matches_the_c_abi_path(gpu, tmp_path):
    _built()
    from topfusion_b200 import synth
    depth, _, _ = synth.sequence("S1", 6)
    for i in range(6):
        synth.write_pgm(str(tmp_path / ("%04d.pgm" % i)), depth[i])
    view = tmp_path / "view.pgm"
    r = subprocess.run([DEMO, str(tmp_path), "6", "--corrected", "--out", str(view)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = re.findall(r"frame\s+(\d+) ok=(\d) t=\(\s*(\S+)\s+(\S+)\s+(\S+)\) voxel-updates=(\d+)", r.stdout)
    assert len(rows) == 6 and all(x[1] == "1" for x in rows)
    g = gpu.Context(corrected_mode=1)
    try:
        for i in range(6):
            assert g.process_frame(depth[i])
            t = np.array([float(v) for v in rows[i][2:5]])
            assert np.abs(t - g.pose()[:3, 3]).max() < 2e-5      # printed with 5 decimals
            assert int(rows[i][5]) == g.voxel_updates()
        img = g.render_image()
        assert int(re.search(r"cloud: (\d+) surface points", r.stdout).group(1)) == g.extract_points().shape[0] > 50000
    finally:
        g.close()
    raw = open(view, "rb").read()
    px = np.frombuffer(raw[raw.index(b"255\n") + 4:], np.uint8).reshape(480, 640)
    assert np.array_equal(px, img[..., 0])
