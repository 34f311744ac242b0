"""Writes a synthetic sequence as 16-bit PGM frames named %04d.pgm — the input format of the reference's
apps/demo.cpp (cv::imread(..., CV_16U)) and of apps/demo_synth.

    python -m topfusion_b200.synth_cli <out_dir> [S0|S1|S2|S3] [n_frames]
"""
import os
import sys

from . import synth


def main(argv):
    if len(argv) < 2:
        print(__doc__)
        return 2
    out, seq, n = argv[1], (argv[2] if len(argv) > 2 else "S1"), int(argv[3]) if len(argv) > 3 else 20
    os.makedirs(out, exist_ok=True)
    depth, poses, _ = synth.sequence(seq, n)
    for i in range(n):
        synth.write_pgm(os.path.join(out, "%04d.pgm" % i), depth[i])
    print(f"wrote {n} frames of {seq} to {out}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
