"""Diagnostic: run-to-run and host-vs-device bit comparison of the graphed guided loop (tiny fp32 and full-size fp16)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_step_driver import _pipe, _run
from oracle.cases import E2E_CASE

for dtype in (torch.float32, torch.float16):
    pipe, store, cfg, embeds, case = _pipe(dtype)
    outs = {}
    for name, dev in (("host_a", False), ("host_b", False), ("dev_a", True), ("dev_b", True), ("host_c", False)):
        outs[name], n, _, mode = _run(pipe, store, cfg, embeds, case, 28, dev)
        print(dtype, name, mode, n, flush=True)
    keys = list(outs)
    for i in range(len(keys)):
        for j in range(i + 1, len(keys)):
            d = (outs[keys[i]].float() - outs[keys[j]].float()).abs().max().item()
            print(dtype, keys[i], keys[j], "max abs diff", d, "equal", torch.equal(outs[keys[i]], outs[keys[j]]))
