import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_fwd import run
for n, S in ((30, 65536), (14, 131072), (118, 32768)):
    run(n, S, train=True); run(n, S, train="fwdonly"); run(n, S, train=False)
