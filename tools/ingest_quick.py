"""GPU box: the ingest leg of bench.py alone (PGM files -> poses, synchronous loop against the ring)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["bench.py"]
import bench
fr = bench.seq_frames("S1", 100)
for _ in range(int(os.environ.get("REPS", "2"))):
    print(json.dumps(bench.ingest_leg(fr, 1)))
