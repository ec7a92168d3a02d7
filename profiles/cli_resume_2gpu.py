"""denovo3DBatch on 2 GPUs (torchrun, NCCL) against the same search on 1 GPU, and a resumed 2-GPU search after one
rank's tile file was lost.  usage (box with 2 GPUs): python profiles/cli_resume_2gpu.py [workdir]
Prints one line per check; exit code 0 only if the three score tables are bit-identical and the top-K lists agree."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from helicon_b200.denovo3DBatch import write_mrc

work = sys.argv[1] if len(sys.argv) > 1 else "/tmp/cli_resume"
os.makedirs(work, exist_ok=True)
mrc = os.path.join(work, "img.mrc")
write_mrc(mrc, bench.synthetic_filament(), 1.3)
grid = ["--twist=-2.0:-0.5:12", "--rise", "4.5:5.0:4", "--positive-constraint", "0", "--batch-candidates", "8", "--top-k", "5"]
env = dict(os.environ, PYTHONPATH=ROOT)
one, two = os.path.join(work, "one"), os.path.join(work, "two")
torchrun = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
            "--master-port", "29571", "-m", "helicon_b200.denovo3DBatch", mrc] + grid + ["--output", two, "--resume"]


def run(cmd, tag):
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    print(f"[{tag}] rc={r.returncode} " + " | ".join(l for l in r.stdout.splitlines() if l.startswith("image")), flush=True)
    if r.returncode != 0:
        print(r.stderr[-3000:])
        sys.exit(1)


def table(prefix):
    z = np.load(prefix + "_img0_scores.npz")
    rep = json.load(open(prefix + "_top.json"))[0]
    return z["scores"], z["itn"], [(e["twist"], e["rise"], e["score"]) for e in rep["top"]], rep


run([sys.executable, "-m", "helicon_b200.denovo3DBatch", mrc] + grid + ["--output", one], "1 GPU")
run(torchrun, "2 GPUs")
s1, i1, t1, _ = table(one)
s2, i2, t2, r2 = table(two)
ok = np.array_equal(s1, s2, equal_nan=True) and np.array_equal(i1, i2) and t1 == t2
print(f"2 GPUs == 1 GPU: scores {np.array_equal(s1, s2, equal_nan=True)} itn {np.array_equal(i1, i2)} top-K {t1 == t2} "
      f"({int(np.isfinite(s2).sum())} candidates, restored {r2['n_restored']})")
tiles = sorted(f for f in os.listdir(work) if ".tiles" in f)
print("tile files:", tiles, [len(np.load(os.path.join(work, f))["ti"]) for f in tiles])
os.remove(os.path.join(work, "two_img0.tiles.rank1.npz"))  # rank 1's work is lost
run(torchrun, "2 GPUs, resumed")
s3, i3, t3, r3 = table(two)
ok3 = np.array_equal(s1, s3, equal_nan=True) and np.array_equal(i1, i3) and t1 == t3
print(f"resumed == 1 GPU: scores {np.array_equal(s1, s3, equal_nan=True)} itn {np.array_equal(i1, i3)} top-K {t1 == t3} "
      f"(restored {r3['n_restored']} of {int(np.isfinite(s3).sum())})")
sys.exit(0 if ok and ok3 and 0 < r3["n_restored"] < int(np.isfinite(s3).sum()) else 1)
