"""Diagnostic (round 2): where does the device-side step driver lose 1.5 % against the host loop?  Times one captured
UNet program (eval, SD-1.4 fp16) executed 30 times as (1) plain replays of its own instantiated graph, (2) a parent graph
holding it as a child-graph node, (3) a parent graph holding it inside an always-true IF conditional node."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cuda.bindings import runtime as rt

import bench
from guided_attention_b200 import shared_state as S
from guided_attention_b200.pipeline_guided_attention import _StepGraphs


def chk(res):
    err = res[0]
    assert int(err) == 0, res
    return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)


class A:
    unet = "sd14"; denoise_steps = 50; no_graphs = False; host_control = True
args = A()
dev = torch.device("cuda", 0)
cfg, pipe, store, embeds_host = bench.build_pipeline(args, dev)
pipe.prompt = cfg.prompt
pipe.scheduler.set_timesteps(50)
emb = embeds_host.to(dev, torch.float16)
lat = torch.randn(1, 4, 64, 64, device=dev, dtype=torch.float16)
loss_kw = dict(attention_store=store, attention_res=16, smooth_attentions=True, sigma=0.5, kernel_size=3, normalize_eot=False)
G = _StepGraphs(pipe, store, loss_kw, emb, 7.5, lat)
store.text_kv = G.text_kv
for name in ("eval", "update"):
    G._graph(name)
stream = torch.cuda.current_stream().cuda_stream
N = 30


def timed(fn):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(N):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / N


for name in ("eval", "update"):
    raw = int(G.graphs[name].raw_cuda_graph())
    g_plain = G.graphs[name]
    t_plain = timed(lambda: g_plain.replay())
    # (2) child graph node
    parent = chk(rt.cudaGraphCreate(0))
    chk(rt.cudaGraphAddChildGraphNode(parent, None, 0, raw))
    ex2 = chk(rt.cudaGraphInstantiate(parent, 0))
    t_child = timed(lambda: chk(rt.cudaGraphLaunch(ex2, stream)))
    # (3) IF(true) { child }
    parent3 = chk(rt.cudaGraphCreate(0))
    handle = chk(rt.cudaGraphConditionalHandleCreate(parent3, 1, rt.cudaGraphConditionalHandleFlags.cudaGraphCondAssignDefault if hasattr(rt, "cudaGraphConditionalHandleFlags") else 1))
    params = rt.cudaGraphNodeParams()
    params.type = rt.cudaGraphNodeType.cudaGraphNodeTypeConditional
    params.conditional.handle = handle
    params.conditional.type = rt.cudaGraphConditionalNodeType.cudaGraphCondTypeIf
    params.conditional.size = 1
    node = chk(rt.cudaGraphAddNode(parent3, None, 0, params))
    body = params.conditional.phGraph_out[0]
    chk(rt.cudaGraphAddChildGraphNode(body, None, 0, raw))
    ex3 = chk(rt.cudaGraphInstantiate(parent3, 0))
    t_cond = timed(lambda: chk(rt.cudaGraphLaunch(ex3, stream)))
    print(f"{name}: plain replay {t_plain:.3f} ms | as child-graph node {t_child:.3f} ms | inside IF(true) {t_cond:.3f} ms", flush=True)
