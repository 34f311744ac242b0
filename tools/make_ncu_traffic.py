#!/usr/bin/env python
"""profiles/ncu_traffic.json from the two ncu summaries of a round (tools/ncu_summary.py output): DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum, average over the captured launches) of every kernel, on the headline frame
(S1 640x480, 5 mm) and on the large scene (S3, 2 mm).  bench.py copies these into `roofline.traffic`.

    python tools/make_ncu_traffic.py profiles/r02_ncu_full_S1_summary.csv profiles/r02_ncu_full_S3_summary.csv"""
import csv
import json
import os
import sys


def table(path):
    rows = list(csv.reader(open(path)))
    h = rows[0]
    ki = 0
    ri = next(i for i, c in enumerate(h) if c.startswith("dram__bytes_read.sum"))
    wi = next(i for i, c in enumerate(h) if c.startswith("dram__bytes_write.sum"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    sr = scale[h[ri].split("[")[1].rstrip("]")]
    sw = scale[h[wi].split("[")[1].rstrip("]")]
    acc = {}
    for r in rows[1:]:
        k = r[ki].replace("void ", "").split("<")[0].strip()
        acc.setdefault(k, []).append(float(r[ri]) * sr + float(r[wi]) * sw)
    return {k: sum(v) / len(v) for k, v in acc.items()}


def main(s1, s3):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {"source": f"{os.path.relpath(s1, root)} (ncu --set full, S1 640x480 headline frame, cold caches per replay)",
           "source_large_scene": f"{os.path.relpath(s3, root)} (ncu --set full, S3 large scene at 2 mm voxels)",
           "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
           "kernels": table(s1), "kernels_large_scene": table(s3)}
    json.dump(out, open(os.path.join(root, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
