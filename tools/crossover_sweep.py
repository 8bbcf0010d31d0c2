"""Variant crossover sweep: K1 / K2 single-shot vs persistent pipelined kernel, forced through
`impl`, over the batch sizes between the pipeline's launches (B = 1, 2) and the seed-batched ones (B = 8, 16 ...).
`python tools/crossover_sweep.py > profiles/r02_crossover_sweep.jsonl`"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from guided_attention_b200 import _cabi as abi, microbench  # noqa: E402

for (N, d, maps) in ((1024, 80, True), (256, 160, True), (64, 160, True), (4096, 40, False)):
    for B in (1, 2, 4, 8, 16, 32, 64):
        for direction in ("fwd", "bwd"):
            row = {"kernel": f"cross_attn_{direction}", "N": N, "d": d, "maps": maps, "B": B}
            for name, impl in (("single", abi.GA_IMPL_TCGEN05_SINGLE), ("pipe", abi.GA_IMPL_TCGEN05_PIPE),
                               ("auto", abi.GA_IMPL_AUTO)):
                try:
                    m = microbench.time_cross_attn(B, 8, N, 77, d, torch.float16, with_acc=maps, direction=direction,
                                                   impl=impl)
                    row[name + "_us"] = round(m["us"], 2)
                except Exception as e:
                    row[name + "_us"] = None
                    row[name + "_err"] = str(e)[:80]
            print(json.dumps(row), flush=True)
