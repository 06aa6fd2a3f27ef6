import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_fwd import run
run(300, 16384, train=True, reps=7)
run(300, 16384, train="fwdonly", reps=7)
run(300, 16384, train=False, reps=7)
