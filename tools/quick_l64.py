import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_fwd import run
run(300, 4096, K=8, L=64, train=True)
run(300, 4096, K=8, L=64, train="fwdonly")
