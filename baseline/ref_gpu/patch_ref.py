#!/usr/bin/env python3
"""REFERENCE-ARM INFRASTRUCTURE (not product code).

Generates the PATCHED-TO-COMPILE copy of the reference's library sources that
`baseline/ref_gpu/Makefile` builds into the git-ignored `baseline/_ref/`
(SURVEY.md F11 / §8c).  Nothing it writes is committed: the output
directory `baseline/_ref/gen/` is a build artefact like an object file.  The reference tree
itself (`/root/reference`, read-only) is never modified.

The patch is this list of edits — nothing else differs from the reference sources:

 P1  src/cuda/proj_icp.cu:9-11     `texture<T, 2> name;`  ->  `__device__ tfcompat::TexRef<T> name;`
                                   (legacy texture references no longer exist in CUDA 12; the
                                   sampling semantics are kept by texture objects, compat/pre.hpp)
 P2  src/cuda/proj_icp.cu:407-408,432-433   the host-side `xxx_tex.filterMode = cudaFilterModePoint;`
                                   lines are dropped (the mode is set where the object is created)
 P3  src/cuda/texture_binder.hpp   replaced by compat/texture_binder.hpp (bind = texture object)
 P4  src/cuda/SceneReconstructionEngine_host.cu:21-23   forward declaration says `const ushort* depth`,
                                   the definition at :297-300 says `const float* depth`: declaration fixed
 P5  (only the "nodebug" variant) src/topfu.cpp: the debug work of SURVEY F10 is removed — the four
     blocking downloads (:211-223), the pose print (:246-252) and the extra render + download
     (:284-288).  None of them feeds a result.  The "asis" variant keeps all of it.

Warp votes without `_sync` and the missing <limits>/<cuda_fp16.h> includes are handled by the
force-included compat/pre.hpp, not by editing.
"""
import re
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent


def sub_once(text, pattern, repl, what, count_expected=None, flags=0):
    new, n = re.subn(pattern, repl, text, flags=flags)
    if n == 0 or (count_expected is not None and n != count_expected):
        raise SystemExit(f"patch_ref: edit '{what}' matched {n} times (expected {count_expected or '>=1'})")
    return new


def patch_proj_icp(t):
    t = sub_once(t, r"^(\s*)texture<\s*(\w+)\s*,\s*2\s*>\s+(\w+)\s*;", r"\1__device__ tfcompat::TexRef<\2> \3;", "P1", 3, re.M)
    t = sub_once(t, r"^\s*\w+_tex\.filterMode\s*=\s*cudaFilterModePoint;\s*$\n", "", "P2", 4, re.M)
    return t


def patch_scene_host(t):
    head, tail = t[:2500], t[2500:]
    head = sub_once(head, r"const\s+ushort\s*\*\s*depth", "const float* depth", "P4", 1)
    return head + tail


def strip_debug_topfu(t):
    """P5.  Whole lines of the reference file are removed, inside three anchored regions."""
    lines = t.split("\n")

    def find(pat, start=0):
        for i in range(start, len(lines)):
            if re.search(pat, lines[i]):
                return i
        raise SystemExit(f"patch_ref: P5 anchor '{pat}' not found")

    drop = set()
    # region 1: the four debug downloads between `Affine3f affine;` and the commented-out ifstream block
    a = find(r"^\s*Affine3f affine;\s*$")
    b = find(r"^\s*//ifstream inFile;", a)
    for i in range(a + 1, b):
        if re.search(r"cv::Mat points_mat\d*\(|\.download\(points_mat\d*\.data", lines[i]):
            drop.add(i)
    # region 2: the pose print
    drop.add(find(r"^\s*cout<<\"pose:\"<<endl<<M_d_print<<endl;", b))
    # region 3: the extra render + download between the integration block and `//prepare icp`
    c = find(r"^\s*cuda::image4u image;\s*$", b)
    d = find(r"^\s*//prepare icp", c)
    for i in range(c, d):
        if lines[i].strip():
            drop.add(i)
    if len(drop) != 13:
        raise SystemExit(f"patch_ref: P5 removed {len(drop)} lines, expected 13")
    return "\n".join(l for i, l in enumerate(lines) if i not in drop)


def main():
    ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    out = Path(sys.argv[2] if len(sys.argv) > 2 else HERE.parent / "_ref" / "gen")
    src = ref / "tfusion" / "src"
    if not src.is_dir():
        raise SystemExit(f"patch_ref: no reference tree at {ref}")
    if out.exists():
        shutil.rmtree(out)
    (out / "src" / "cuda").mkdir(parents=True)
    keep = ["precomp.hpp", "precomp.cpp", "internal.hpp", "safe_call.hpp", "imgproc.cpp", "projective_icp.cpp",
            "device_memory.cpp", "topfu.cpp",
            "cuda/device.hpp", "cuda/temp_utils.hpp", "cuda/imgproc.cu", "cuda/proj_icp.cu",
            "cuda/SceneReconstructionEngine_host.cu", "cuda/VisualisationEngine_CUDA.cu",
            "cuda/VisualisationHelper.cu", "cuda/CUDAInstantiations.cu"]
    for rel in keep:
        t = (src / rel).read_text(encoding="latin-1")
        if rel == "cuda/proj_icp.cu":
            t = patch_proj_icp(t)
        elif rel == "cuda/SceneReconstructionEngine_host.cu":
            t = patch_scene_host(t)
        (out / "src" / rel).write_text(t, encoding="latin-1")
        if rel == "topfu.cpp":
            (out / "src" / "topfu_nodebug.cpp").write_text(strip_debug_topfu(t), encoding="latin-1")
    shutil.copy(HERE / "compat" / "texture_binder.hpp", out / "src" / "cuda" / "texture_binder.hpp")  # P3
    print(f"patch_ref: wrote patched sources to {out}")


if __name__ == "__main__":
    main()
