"""Input side: the PGM decoder must return what the reference's `cv::imread(file, CV_16U)` returns
(apps/demo.cpp:96) — pinned here against cv2.imread(..., cv2.IMREAD_ANYDEPTH), the same OpenCV decoder — and the
decode-ahead ring must deliver every frame of a sequence, in order, bit for bit, whatever the consumer's pace."""
import ctypes as C
import os
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "topfusion_b200", "libtfusion.so")


@pytest.fixture(scope="module")
def io():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    L = C.CDLL(LIB)
    L.tfio_probe_pgm16.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.tfio_read_pgm16.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    L.tfio_ring_open.restype = C.c_void_p
    L.tfio_ring_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.tfio_ring_next.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint16)), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    L.tfio_ring_release.argtypes = [C.c_void_p, C.c_int]
    L.tfio_ring_release.restype = None
    L.tfio_ring_pinned.argtypes = [C.c_void_p]
    L.tfio_ring_error.argtypes = [C.c_void_p]
    L.tfio_ring_error.restype = C.c_char_p
    L.tfio_ring_close.argtypes = [C.c_void_p]
    L.tfio_ring_close.restype = None
    return L


def write_pgm(path, arr, maxval=65535, plain=False, comment=False, sep=b"\n"):
    h, w = arr.shape
    with open(path, "wb") as f:
        f.write(b"P2\n" if plain else b"P5\n")
        if comment:
            f.write(b"# depth in millimetres\n")
        f.write(b"%d %d" % (w, h) + sep + b"%d\n" % maxval)
        if plain:
            f.write((" ".join(str(int(v)) for v in arr.ravel()) + "\n").encode())
        elif maxval > 255:
            f.write(arr.astype(">u2").tobytes())
        else:
            f.write(arr.astype("u1").tobytes())


def decode(io, path, pad=0):
    cols, rows = C.c_int(), C.c_int()
    if not io.tfio_probe_pgm16(path.encode(), C.byref(cols), C.byref(rows)):
        return None
    out = np.full((rows.value, cols.value + pad), 0xABCD, np.uint16)
    if not io.tfio_read_pgm16(path.encode(), out.ctypes.data, out.strides[0], cols.value, rows.value):
        return None
    assert pad == 0 or (out[:, cols.value:] == 0xABCD).all()      # the row padding is not written
    return out[:, :cols.value]


@pytest.mark.parametrize("kind", ["binary", "comment", "plain", "maxval4000", "space_separated", "pitched"])
def test_pgm_decoder_equals_opencv(io, tmp_path, kind):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    a = rng.integers(0, 65536 if kind != "maxval4000" else 4001, (37, 53)).astype(np.uint16)
    p = str(tmp_path / "f.pgm")
    write_pgm(p, a, maxval=4000 if kind == "maxval4000" else 65535, plain=kind == "plain", comment=kind == "comment",
              sep=b" " if kind == "space_separated" else b"\n")
    ref = cv2.imread(p, cv2.IMREAD_ANYDEPTH)                       # CV_16U == 2 == IMREAD_ANYDEPTH (apps/demo.cpp:96)
    assert ref.dtype == np.uint16 and np.array_equal(ref, a)        # OpenCV: big-endian samples, no rescaling by maxval
    got = decode(io, p, pad=3 if kind == "pitched" else 0)
    assert got is not None and np.array_equal(got, ref)


def test_synth_frames_round_trip(io, tmp_path, s1_frames):
    """the writer the tests and demo_synth use (synth.write_pgm) against the reader, on a full 640x480 frame"""
    from topfusion_b200 import synth
    depth, _, _ = s1_frames
    p = str(tmp_path / "0000.pgm")
    synth.write_pgm(p, depth[3])
    assert np.array_equal(decode(io, p), depth[3])


def test_decoder_refuses_what_the_reference_cannot_use(io, tmp_path):
    a = np.arange(12, dtype=np.uint16).reshape(3, 4)
    p8 = str(tmp_path / "eight.pgm")
    write_pgm(p8, a, maxval=255)                                   # OpenCV returns CV_8U here; demo.cpp would upload it as u16
    assert decode(io, p8) is None
    short = str(tmp_path / "short.pgm")
    write_pgm(short, a)
    with open(short, "r+b") as f:
        f.truncate(os.path.getsize(short) - 5)
    assert decode(io, short) is None
    colour = str(tmp_path / "colour.ppm")
    open(colour, "wb").write(b"P6\n1 1\n255\n\0\0\0")
    assert decode(io, colour) is None and decode(io, str(tmp_path / "missing.pgm")) is None
    # wrong size for the destination
    good = str(tmp_path / "good.pgm")
    write_pgm(good, a)
    buf = np.zeros((3, 4), np.uint16)
    assert io.tfio_read_pgm16(good.encode(), buf.ctypes.data, buf.strides[0], 5, 3) == 0
    assert io.tfio_read_pgm16(good.encode(), buf.ctypes.data, 6, 4, 3) == 0        # step smaller than a row


def drain(io, ring, hold=1, pause=0.0):
    """consume the ring holding up to `hold` frames at a time; returns [(index, copy of the frame)]"""
    out, held = [], []
    data, rows, cols, step, index = C.POINTER(C.c_uint16)(), C.c_int(), C.c_int(), C.c_size_t(), C.c_int()
    while io.tfio_ring_next(ring, C.byref(data), C.byref(rows), C.byref(cols), C.byref(step), C.byref(index)):
        assert step.value == cols.value * 2
        out.append((index.value, np.ctypeslib.as_array(data, (rows.value, cols.value)).copy()))
        held.append(index.value)
        if pause:
            time.sleep(pause)
        while len(held) >= hold:
            io.tfio_ring_release(ring, held.pop(0))
    for i in held:
        io.tfio_ring_release(ring, i)
    return out


@pytest.mark.parametrize("slots,hold,pause,count,decoders", [(2, 1, 0.0, -1, 1), (3, 2, 0.0, -1, 2), (4, 1, 0.003, -1, 3),
                                                             (3, 1, 0.0, 7, 2), (6, 3, 0.0, -1, 6), (4, 1, 0.0, 11, 4)])
def test_ring_delivers_the_sequence_in_order(io, tmp_path, slots, hold, pause, count, decoders):
    rng = np.random.default_rng(slots * 10 + hold)
    frames = [rng.integers(0, 65536, (24, 32)).astype(np.uint16) for _ in range(11)]
    for i, f in enumerate(frames):
        write_pgm(str(tmp_path / ("%04d.pgm" % (i + 2))), f)      # the sequence starts at 0002.pgm
    ring = io.tfio_ring_open(str(tmp_path).encode(), slots, 2, count, 1, decoders)
    assert ring
    try:
        got = drain(io, ring, hold=hold, pause=pause)
        n = 11 if count < 0 else count
        assert [i for i, _ in got] == list(range(2, 2 + n))
        assert all(np.array_equal(a, frames[i - 2]) for i, a in got)
        assert io.tfio_ring_error(ring) == b""                    # an open-ended sequence ends at the first missing file
    finally:
        io.tfio_ring_close(ring)


def test_ring_reports_a_broken_file_and_can_be_closed_early(io, tmp_path):
    a = np.arange(24 * 32, dtype=np.uint16).reshape(24, 32)
    for i in range(6):
        write_pgm(str(tmp_path / ("%04d.pgm" % i)), a + i)
    with open(str(tmp_path / "0003.pgm"), "r+b") as f:
        f.truncate(100)
    ring = io.tfio_ring_open(str(tmp_path).encode(), 3, 0, -1, 1, 2)
    try:
        got = drain(io, ring)
        assert [i for i, _ in got] == [0, 1, 2] and b"0003.pgm" in io.tfio_ring_error(ring)
    finally:
        io.tfio_ring_close(ring)
    # a fixed-length sequence that runs out of files is an error too
    ring = io.tfio_ring_open(str(tmp_path).encode(), 3, 4, 5, 1, 3)
    try:
        assert [i for i, _ in drain(io, ring)] == [4, 5] and b"0006.pgm" in io.tfio_ring_error(ring)
    finally:
        io.tfio_ring_close(ring)
    # closing while the producer is ahead and a frame is still held must not hang
    ring = io.tfio_ring_open(str(tmp_path).encode(), 2, 0, 3, 1, 2)
    data, rows, cols, step, index = C.POINTER(C.c_uint16)(), C.c_int(), C.c_int(), C.c_size_t(), C.c_int()
    assert io.tfio_ring_next(ring, C.byref(data), C.byref(rows), C.byref(cols), C.byref(step), C.byref(index)) == 1
    io.tfio_ring_close(ring)
    # an empty directory: no frames, an error naming the first file
    empty = tmp_path / "empty"
    empty.mkdir()
    ring = io.tfio_ring_open(str(empty).encode(), 3, 0, -1, 1, 2)
    try:
        assert drain(io, ring) == [] and b"0000.pgm" in io.tfio_ring_error(ring)
    finally:
        io.tfio_ring_close(ring)


def test_ring_fails_loudly_without_page_locked_memory(io, tmp_path):
    """no CUDA device here: page-locking is impossible, and unless the caller allows pageable memory the ring refuses"""
    from conftest import has_cuda
    if has_cuda():
        pytest.skip("a CUDA device is present: page-locking works")
    write_pgm(str(tmp_path / "0000.pgm"), np.zeros((4, 4), np.uint16))
    assert io.tfio_ring_open(str(tmp_path).encode(), 3, 0, -1, 0, 2) is None
    ring = io.tfio_ring_open(str(tmp_path).encode(), 3, 0, -1, 1, 2)
    assert ring and io.tfio_ring_pinned(ring) == 0
    io.tfio_ring_close(ring)


@pytest.mark.gpu
def test_demo_with_the_ring_equals_the_synchronous_demo(gpu, tmp_path):
    """demo_synth --ring (decode-ahead, page-locked slots, asynchronous upload inside operator()) prints the same poses and
    voxel-update counts and renders the same view as the imread + upload loop of the reference's demo"""
    import re
    import subprocess
    from topfusion_b200 import synth
    demo = os.path.join(ROOT, "apps", "demo_synth")
    depth, _, _ = synth.sequence("S1", 8)
    for i in range(8):
        synth.write_pgm(str(tmp_path / ("%04d.pgm" % i)), depth[i])
    outs = []
    for extra in ([], ["--ring"]):
        view = tmp_path / ("view%d.pgm" % len(outs))
        r = subprocess.run([demo, str(tmp_path), "8", "--corrected", "--out", str(view)] + extra, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append((re.findall(r"frame\s+\d+ ok=\d t=\([^)]*\) voxel-updates=\d+", r.stdout), open(view, "rb").read(), r.stdout))
    assert len(outs[0][0]) == 8 and outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
    assert "page-locked memory" in outs[1][2]


def test_ring_is_race_free(tmp_path):
    """src/frame_ring.cpp under ThreadSanitizer: 1-4 decoder threads against a consumer that holds frames, and an early close"""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    exe = str(tmp_path / "ring_tsan")
    b = subprocess.run([cxx, "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cpp", "frame_ring_tsan.cpp"), os.path.join(ROOT, "src", "frame_ring.cpp"),
                        "-o", exe, "-lpthread"], capture_output=True, text=True, timeout=300)
    if b.returncode != 0 and "tsan" in b.stderr.lower():
        pytest.skip("ThreadSanitizer runtime not installed")
    assert b.returncode == 0, b.stderr[-2000:]
    frames = tmp_path / "frames"
    r = subprocess.run([exe, str(frames)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "bad = 0" in r.stdout and "ThreadSanitizer" not in r.stderr, r.stdout + r.stderr[-3000:]
