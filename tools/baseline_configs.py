"""The five BASELINE.json configurations on one GPU (device-resident inputs, CUDA events, median of 5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.timing import run
print("configs[0] case14 K=4 L=20 forward, batch 1024"); run(14, 1024)
print("configs[1] case30 training step, batch 4096"); run(30, 4096, train=True)
print("configs[2] case118 training step, batch 16384"); run(118, 16384, train=True)
print("configs[3] case300 inference, batch 65536"); run(300, 65536)
print("configs[4] case300 K=8 L=64 training step, micro-batch 8192 (bench.py runs the stated 32768 as four of these)"); run(300, 8192, K=8, L=64, train=True)
print("configs[4] forward only"); run(300, 8192, K=8, L=64)
