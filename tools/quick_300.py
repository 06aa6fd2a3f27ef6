import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_fwd import run
train = len(sys.argv) > 1 and sys.argv[1] == "train"
run(300, 32768 if not train else 16384, train=train)
