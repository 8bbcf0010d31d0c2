"""Tail-only microbench sweep (forward and backward, res 16 / 32, 1..2048 samples): JSON lines."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from guided_attention_b200 import microbench as mb

dirs = sys.argv[1:] or ["fwd", "bwd"]
for res in (16, 32):
    for S in (1, 8, 64, 512, 2048):
        for d in dirs:
            print(json.dumps(mb.time_tail(res, 5, 2, n_samples=S, direction=d)))
            sys.stdout.flush()
