"""cProfile of the rs predict shard loop (cfg 4) on one GPU: where does the main thread spend its time per batch?"""
import argparse
import cProfile
import os
import pstats
import shutil
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from robosat_b200 import synth  # noqa: E402
from robosat_b200.tools.predict import run_shard  # noqa: E402

root = tempfile.mkdtemp(prefix="rsb_cfg4prof_")
n_cols = int(sys.argv[1]) if len(sys.argv) > 1 else 32
synth.write_slippy_tiles(os.path.join(root, "tiles"), 18, range(1000, 1000 + n_cols), range(2000, 2032), size=512, seed=7, workers=32)
dev = torch.device("cuda:0")
sd = synth.make_state_dict(2, seed=0)
args = argparse.Namespace(batch_size=32, overlap=32, tile_size=512, workers=0, tiles=os.path.join(root, "tiles"), probs=os.path.join(root, "probs"))
os.environ["RSB_QUIET"] = "1"
run_shard(0, 1, args, dev, sd, 2, stats={})  # warm: page cache, plan, lazy imports
shutil.rmtree(os.path.join(root, "probs"), ignore_errors=True)
pr = cProfile.Profile()
st = {}
pr.enable()
run_shard(0, 1, args, dev, sd, 2, stats=st)
pr.disable()
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()})
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
shutil.rmtree(root, ignore_errors=True)
