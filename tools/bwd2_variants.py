"""Timing of the warp-specialised backward kernel under its tuning knobs (GNS_BWD2_ROLEMAP, GNS_BWD2_CW)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from check_bwd2 import timing, parity
parity(300, 3, True)
for env in ({"GNS_BWD2_ROLEMAP": "0"}, {"GNS_BWD2_ROLEMAP": "1"}, {"GNS_BWD2_ROLEMAP": "0", "GNS_BWD2_CW": "3"},
            {"GNS_BWD2_ROLEMAP": "0", "GNS_BWD2_CW": "7"}):
    for k in ("GNS_BWD2_ROLEMAP", "GNS_BWD2_CW"):
        os.environ.pop(k, None)
    os.environ.update(env)
    print(env, flush=True)
    timing(300, 16384, "1")
