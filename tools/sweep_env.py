"""Sweep plan / balancing knobs (environment variables read by the library) on the case300 training step."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from check_bwd2 import timing
combos = [{}] + [{"GNS_BWD_LINE_COST": c} for c in ("0.15", "0.5", "0.8", "1.2")] + [{"GNS_DEG_CAP": c} for c in ("2", "4")] + \
         [{"GNS_FWD_LINE_COST": c} for c in ("0.1", "0.35")] + [{"GNS_DETERMINISTIC": "0"}]
for env in combos:
    for k in ("GNS_BWD_LINE_COST", "GNS_DEG_CAP", "GNS_FWD_LINE_COST", "GNS_DETERMINISTIC"):
        os.environ.pop(k, None)
    os.environ.update(env)
    print(env, flush=True)
    try:
        timing(300, 16384, "0")
    except Exception as ex:
        print("   failed:", str(ex)[:120], flush=True)
