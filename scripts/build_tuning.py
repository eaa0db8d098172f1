"""Build tuning variants of the library (different NCELL / JCHUNK) for one-call A/B runs on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magnet_b200 import build
for nc, jc in ((4, 32), (6, 32), (8, 64), (4, 16), (6, 64), (3, 32)):
    print(build.build(defines=(f"MAGNET_NCELL={nc}", f"MAGNET_JCHUNK={jc}"), tag=f"nc{nc}jc{jc}"))
