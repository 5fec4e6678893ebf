"""Per-kernel device times for tcgen05 tile configurations (run on the GPU box): one process per configuration,
because FR_TC_TILES is read once per engine.  python tools/sweep_tiles.py [model] [batches...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
model = sys.argv[1] if len(sys.argv) > 1 else "small"
batches = sys.argv[2:] or ["2048", "4096", "16384"]
for tiles in ("", "256,256,256,2", "512,512,256,2", "256,512,256,2", "512,256,256,2", "128,128,256,2"):
    env = dict(os.environ)
    if tiles:
        env["FR_TC_TILES"] = tiles
    else:
        env.pop("FR_TC_TILES", None)
    for B in batches:
        sys.stdout.flush()
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_kernels.py"), model, B, "50"], env=env)
