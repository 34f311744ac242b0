"""GPU parity — the whole frame path TopFu::operator() (src/topfu.cpp:161-330) through tfb_process_frame,
against the oracle, in REFERENCE mode (bug-for-bug incl. SURVEY.md F1) on the S0 hover sequence and in
corrected mode on the S1 orbit.  Pose tolerance 1e-4 m / 1e-4 rad per frame (north star); block sets are compared
as sets — with the pose differing in its last bits a handful of knife-edge blocks may differ, bounded below."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rot_angle(Ra, Rb):
    """angle of Ra^T Rb from its skew part (arccos of the trace has a 5e-4 noise floor on fp32 rotations)"""
    D = Ra.astype(np.float64).T @ Rb.astype(np.float64)
    w = 0.5 * np.array([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(w))))


def _run(gpu, depth, n, **kw):
    from oracle import tfo
    o = tfo.Oracle(**kw)
    g = gpu.Context(ieee_arith=1, **kw)
    rows = []
    for i in range(n):
        ok_o = o.process_frame(depth[i])
        ok_g = g.process_frame(depth[i])
        rows.append((ok_o, ok_g, o.pose().copy(), g.pose().copy(), o.counters(), g.counters(), o.voxel_updates(), g.voxel_updates()))
    return o, g, rows, tfo


def test_reference_mode_hover_sequence(gpu):
    """Reference mode reproduces SURVEY.md F1 bug for bug: the model maps are in the world frame, the motion is
    double counted and the estimated pose runs away exponentially (0.3 mm, 3 mm, 6 mm, ... 11 cm, 1.9 m) until
    no correspondence survives and operator() resets.  The per-frame tolerance (1e-4 m, 1e-4 rad) is asserted
    while the reference itself is within 5 cm of the true pose — an unstable recursion amplifies last-bit
    differences without bound afterwards — and the tracking verdicts / reset frame must agree throughout."""
    from topfusion_b200 import synth
    depth, gt, _ = synth.sequence("S0", 14)
    o, g, rows, tfo = _run(gpu, depth, 14)
    try:
        compared = 0
        stable = True   # until the reference's own estimate has left the true pose by 5 cm
        for i, (ok_o, ok_g, po, pg, co, cg, vo, vg) in enumerate(rows):
            assert ok_o == ok_g, i
            assert co["frame_counter"] == cg["frame_counter"] and co["resets"] == cg["resets"], i
            stable = stable and np.abs(po[:3, 3] - gt[i][:3, 3]).max() < 0.05
            if stable:
                assert np.abs(po[:3, 3] - pg[:3, 3]).max() < 1e-4, (i, po[:3, 3], pg[:3, 3])
                assert _rot_angle(po[:3, :3], pg[:3, :3]) < 1e-4, i
                assert abs(co["n_visible"] - cg["n_visible"]) <= max(8, co["n_visible"] // 200), (i, co, cg)
                assert abs(vo - vg) <= 512 * max(8, co["n_visible"] // 200)
                compared += 1
        assert compared >= 6
        assert not all(r[0] for r in rows), "the reference is expected to lose tracking on this sequence"
        assert o.num_poses() == g.num_poses()
    finally:
        g.close(); o.close()


def _state_report(o, g, tfo):
    """block-set symmetric difference and the histogram of |sdf difference| (LSB of 32767; 1e-3 of truncation = 32.8 LSB) on the
    blocks both sides hold"""
    so, sg = tfo.allocated_set(o.table()), tfo.allocated_set(g.table())
    bo, bg = o.blocks_by_pos(), g.blocks_by_pos()
    edges = [0, 1, 2, 4, 8, 16, 33, 1 << 17]
    hist = np.zeros(len(edges) - 1, np.int64)
    w_diff = 0
    for k in set(bo) & set(bg):
        d = np.abs(bo[k]["sdf"].astype(np.int32) - bg[k]["sdf"].astype(np.int32))
        hist += np.histogram(d, bins=edges)[0]
        w_diff += int((bo[k]["w"] != bg[k]["w"]).sum())
    return {"blocks_oracle": len(so), "blocks_gpu": len(sg), "symdiff": len(so ^ sg), "voxels": int(hist.sum()), "weights_differing": w_diff,
            "sdf_lsb_hist": {f"[{edges[i]},{edges[i + 1]})": int(hist[i]) for i in range(len(hist))}}


def test_tracked_frames_state_report(gpu, s0_frames, s1_frames):
    """Tracked frames (the pose comes from ICP, so it differs from the oracle's in its last bits — ICP sums are accumulated in
    another order): per frame, the allocated-block symmetric difference and the TSDF difference histogram, in both modes.
    Written to gpurun_out/tracked_state_report.json.  Bars: the north star's — block sets identical up to knife-edge blocks (a block
    whose sample lies within a pose-bit of a block face), TSDF within 1e-3 of truncation on >= 99.8 % of the voxels."""
    import json, os
    from oracle import tfo
    report = {}
    for name, frames, n, kw in (("reference_mode_S0", s0_frames[0], 5, {}), ("corrected_mode_S1", s1_frames[0], 8, {"corrected_mode": 1})):
        o = tfo.Oracle(**kw)
        g = gpu.Context(ieee_arith=1, **kw)
        try:
            per_frame = []
            for i in range(n):
                assert o.process_frame(frames[i]) == g.process_frame(frames[i])
                so, sg = tfo.allocated_set(o.table()), tfo.allocated_set(g.table())
                per_frame.append({"frame": i, "blocks": len(so), "symdiff": len(so ^ sg),
                                  "dt_m": float(np.abs(o.pose()[:3, 3] - g.pose()[:3, 3]).max())})
                assert len(so ^ sg) <= max(8, len(so) // 200), (name, i, len(so ^ sg))
            rep = _state_report(o, g, tfo)
            rep["per_frame"] = per_frame
            report[name] = rep
            over = rep["sdf_lsb_hist"]["[33,131072)"]
            assert over <= 2e-3 * rep["voxels"], (name, rep)
        finally:
            g.close(); o.close()
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "tracked_state_report.json"), "w") as f:
        json.dump(report, f, indent=1)


def test_frame0_is_bit_exact(gpu, s1_frames):
    """frame 0 has no ICP: identity pose, so allocation + integration must match the oracle exactly"""
    depth, _, _ = s1_frames
    o, g, rows, tfo = _run(gpu, depth, 1)
    try:
        assert rows[0][0] and rows[0][1]
        assert tfo.allocated_set(o.table()) == tfo.allocated_set(g.table())
        assert o.voxel_updates() == g.voxel_updates()
        bo, bg = o.blocks_by_pos(), g.blocks_by_pos()
        for k in bo:
            assert np.array_equal(bo[k]["sdf"], bg[k]["sdf"]) and np.array_equal(bo[k]["w"], bg[k]["w"])
    finally:
        g.close(); o.close()


def test_corrected_mode_tracks_the_orbit(gpu, s1_frames):
    depth, poses, _ = s1_frames
    o, g, rows, _ = _run(gpu, depth, 8, corrected_mode=1)
    try:
        for i, (ok_o, ok_g, po, pg, *_rest) in enumerate(rows):
            assert ok_o and ok_g
            assert np.abs(po[:3, 3] - pg[:3, 3]).max() < 1e-4, i
            assert _rot_angle(po[:3, :3], pg[:3, :3]) < 1e-4, i
            # and it actually tracks: within 5 mm / 0.2 deg of the ground-truth orbit
            assert np.abs(pg[:3, 3] - poses[i][:3, 3]).max() < 5e-3, (i, pg[:3, 3], poses[i][:3, 3])
            assert _rot_angle(pg[:3, :3], poses[i][:3, :3]) < 3.5e-3, i
    finally:
        g.close(); o.close()


@pytest.mark.parametrize("tail_mode", [1, 2])
def test_reset_after_tracking_loss(gpu, s1_frames, tail_mode):
    """a blank frame kills every correspondence -> operator() returns false and resets (topfu.cpp:263-264); with the tail
    enqueued ahead of the verdict (defer_tail=2) its kernels must leave the scene alone"""
    from oracle import tfo
    depth, _, _ = s1_frames
    o = tfo.Oracle()
    g = gpu.Context(ieee_arith=1, defer_tail=tail_mode)
    try:
        blank = np.zeros_like(depth[0])
        seq = [depth[0], depth[1], blank, depth[2], depth[3]]
        for i, d in enumerate(seq):
            ok_o, ok_g = o.process_frame(d), g.process_frame(d)
            assert ok_o == ok_g, i
            assert o.num_poses() == g.num_poses(), i
        assert g.counters()["resets"] == o.counters()["resets"] == 1
        assert np.abs(o.pose()[:3, 3] - g.pose()[:3, 3]).max() < 1e-4
    finally:
        g.close(); o.close()


def test_pinned_and_strided_input(gpu, s1_frames):
    depth, _, _ = s1_frames
    g1, g2 = gpu.Context(ieee_arith=1), gpu.Context(ieee_arith=1)
    try:
        pin = gpu.PinnedArray((480, 704), np.uint16)   # row stride larger than the image (cv::Mat step)
        for i in range(3):
            pin.array[:, :640] = depth[i]
            view = pin.array[:, :640]
            ok = gpu.C.c_int(0)
            g1._ck(g1.L.tfb_process_frame(g1.h, gpu.C.c_void_p(pin.ptr.value), gpu.C.c_size_t(704 * 2), gpu.C.byref(ok)))
            assert ok.value == 1
            assert g2.process_frame(depth[i])
            assert view.shape == (480, 640)
        assert np.array_equal(g1.pose(), g2.pose())
        pin.free()
    finally:
        g1.close(); g2.close()


def test_gpu_launch_counter_and_timing(gpu, s1_frames):
    depth, _, _ = s1_frames
    g = gpu.Context(ieee_arith=1)
    try:
        g.timing(True)
        n0 = g.kernel_launches()
        for i in range(4):
            assert g.process_frame(depth[i])
        assert g.kernel_launches() - n0 >= 30   # 8 on the first frame, 14 per frame afterwards (the last tail is still pending)
        ms = g.stage_ms()
        assert ms["frame"] > 0 and ms["icp"] > 0 and ms["integrate"] > 0
    finally:
        g.close()


def test_against_committed_golden_vectors(gpu):
    """the CUDA path against tests/golden/*.npz (written from the reference-backed oracle/_ref build)"""
    import glob, os
    paths = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                   if not os.path.basename(p).startswith("refgpu_"))   # those come from the reference's GPU build: test_gpu_refgpu_fixtures.py
    assert paths
    for path in paths:
        z = np.load(path)
        mode, voxel, mu = z["params"]
        intr = z["intr"]
        rows, cols = z["depth"].shape[1:]
        g = gpu.Context(cols=cols, rows=rows, fx=float(intr[0]), fy=float(intr[1]), cx=float(intr[2]), cy=float(intr[3]), ieee_arith=1,
                        corrected_mode=int(mode), voxel_size=float(voxel), mu=float(mu))
        try:
            for i in range(len(z["depth"])):
                ok = g.process_frame(z["depth"][i])
                assert ok == bool(z["ok"][i]), (path, i)
                po = z["est_poses"][i]
                # 160x120 fixtures: this scene leaves directions of the 6x6 system almost unconstrained at that size
                # (smallest eigenvalue 0.15 at level 2), so one correspondence flipping from a 1e-6 map difference
                # moves the solution by millimetres (seen with a 160x120 diagnostic run).  The north-star bound is asserted where the
                # inputs are still bit-identical (frames 0-1) and on the full-resolution fixture throughout.
                tight = cols >= 640 or i <= 1
                tol_t, tol_r = (1e-4, 1e-4) if tight else (1.5e-2, 2e-2)
                assert np.abs(po[:3, 3] - g.pose()[:3, 3]).max() < tol_t, (path, i)
                assert _rot_angle(po[:3, :3], g.pose()[:3, :3]) < tol_r, (path, i)
                nv = int(z["n_visible"][i])
                assert abs(g.counters()["n_visible"] - nv) <= (max(8, nv // 100) if tight else nv // 10), (path, i)
                if i == 0:   # no ICP yet: exact
                    assert g.voxel_updates() == int(z["voxel_updates"][0])
            t = g.table()
            got = {tuple(int(v) for v in p) for p in t[t["ptr"] >= 0]["pos"]}
            want = {tuple(int(v) for v in p) for p in z["blocks"]}
            assert len(got ^ want) <= (max(8, len(want) // 100) if cols >= 640 else len(want) // 8), (path, len(got ^ want))
        finally:
            g.close()


@pytest.mark.parametrize("tail_mode", [1, 2])
def test_deferred_tail_is_result_identical(gpu, s1_frames, tail_mode):
    """defer_tail=1 (default): a call returns once the pose is known and the next call (or any look at the scene) runs
    allocation .. model maps of that frame beside its own preprocessing.  defer_tail=2: they are enqueued behind the frame's
    ICP in the same call, which still returns when the pose is known.  Same bits as the strictly sequential frame."""
    depth, _, _ = s1_frames
    a = gpu.Context(corrected_mode=1, defer_tail=tail_mode)
    b = gpu.Context(corrected_mode=1, defer_tail=0)
    try:
        for i in range(8):
            assert a.process_frame(depth[i]) and b.process_frame(depth[i])
            assert np.array_equal(a.pose().view(np.uint32), b.pose().view(np.uint32)), i
            if i == 4:   # looking at the scene in the middle of the sequence finishes the pending tail first
                assert a.voxel_updates() == b.voxel_updates() > 0
                assert a.counters() == b.counters()
        assert a.voxel_updates_total() + a.voxel_updates() == b.voxel_updates_total()   # the last tail is still pending in `a`
        assert np.array_equal(a.raycast_result().view(np.uint32), b.raycast_result().view(np.uint32))
        assert gpu.allocated_set(a.table()) == gpu.allocated_set(b.table())
        ba, bb = a.blocks_by_pos(), b.blocks_by_pos()
        assert set(ba) == set(bb)
        for k in ba:
            assert np.array_equal(ba[k]["sdf"], bb[k]["sdf"]) and np.array_equal(ba[k]["w"], bb[k]["w"])
        assert np.array_equal(a.render_image(), b.render_image())
    finally:
        a.close(); b.close()


def test_blank_first_frame_then_tracking(gpu, s1_frames):
    """an empty first frame allocates nothing; the next frame cannot be tracked against an empty model -> reset, then the sequence
    starts over.  Verdicts, pose counts and resets must follow the oracle."""
    from oracle import tfo
    depth, _, _ = s1_frames
    o = tfo.Oracle(corrected_mode=1)
    g = gpu.Context(corrected_mode=1, ieee_arith=1)
    try:
        seq = [np.zeros_like(depth[0]), depth[0], depth[1], depth[2]]
        for i, d in enumerate(seq):
            ok_o, ok_g = o.process_frame(d), g.process_frame(d)
            assert ok_o == ok_g, i
            assert o.num_poses() == g.num_poses(), i
            co, cg = o.counters(), g.counters()
            assert co["resets"] == cg["resets"] and co["frame_counter"] == cg["frame_counter"], i
            assert co["n_allocated"] == cg["n_allocated"], i
        assert np.abs(o.pose()[:3, 3] - g.pose()[:3, 3]).max() < 1e-4
    finally:
        g.close(); o.close()
