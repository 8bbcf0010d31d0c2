"""Kernel-only timing of the guidance-path kernels (BASELINE config 3 and its batch sweep).

Each measurement replays a CUDA graph holding `reps` back-to-back launches of ONE kernel on rotating buffer sets, between
two CUDA events on the launching stream: no host launch overhead inside the timed region, and the working set
(`sets` x bytes-per-launch) is sized past the 126 MB L2 unless `l2_resident=True` is asked for explicitly.
`python -m guided_attention_b200.microbench` prints the sweep as JSON lines (used for profiles/ and by bench.py).
"""
from __future__ import annotations

import ctypes as C
import json
import sys

import torch

from . import _cabi as abi
from . import ops

L2_BYTES = 126 * 1024 * 1024
DEFAULT_PROMPT = 'a [robot:.6,.3,.4,.55] and a [blue vase:.2,.3,.4,.55]'   # BASELINE config 1 / 2


def _time_graph(launch, n_sets, reps=20, replays=5, repeats=3):
    """launch(i) enqueues one kernel using buffer set i % n_sets.  Returns microseconds per launch."""
    for i in range(min(n_sets, 3)):
        launch(i)                      # warm-up (lazy attributes, module load)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            launch(i % n_sets)
    g.replay()
    torch.cuda.synchronize()
    samples = []
    for _ in range(repeats):           # median of `repeats` timed blocks: one block is 20-50 ms, short enough for a
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)   # clock / power transient
        s.record()                     # left by whatever ran before to move it by several percent
        for _ in range(replays):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        samples.append(s.elapsed_time(e) * 1e3 / (reps * replays))
    return sorted(samples)[len(samples) // 2]


def time_cross_attn(B, H, N, T, d, dtype=torch.float16, with_acc=True, direction="fwd", impl=abi.GA_IMPL_AUTO,
                    l2_resident=False, device="cuda:0"):
    """One K1 (or K2) launch on (B, N, H*d) operands.  Returns dict(us, bytes, gbs)."""
    lib = abi.load()
    es = torch.empty((), dtype=dtype).element_size()
    nbytes = (ops.attn_fwd_bytes if direction == "fwd" else ops.attn_bwd_bytes)(B, H, N, T, d, es, with_acc)
    n_sets = 1 if l2_resident else max(2, min(64, (2 * L2_BYTES) // max(nbytes, 1) + 1))
    sets = []
    g = torch.Generator(device=device).manual_seed(0)
    for _ in range(n_sets):
        q = torch.randn(B, N, H * d, device=device, dtype=dtype, generator=g)
        k = torch.randn(B, T, H * d, device=device, dtype=dtype, generator=g)
        v = torch.randn(B, T, H * d, device=device, dtype=dtype, generator=g)
        o = torch.empty_like(q)
        lse = torch.empty(B, H, N, device=device, dtype=torch.float32)
        acc = torch.empty(B, N, T, device=device, dtype=torch.float32) if with_acc else None
        do = torch.randn_like(q)
        dq = torch.empty_like(q)
        sets.append((q, k, v, o, lse, acc, do, dq))
    dacc = torch.randn(N, (T + 3) // 4 * 4, device=device, dtype=torch.float32) if with_acc else None
    dt = ops._DTYPES[dtype]
    scale = d ** -0.5
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)   # noqa: E731

    def fwd(i):
        q, k, v, o, lse, acc, do, dq = sets[i]
        abi.check(lib.ga_cross_attn_fwd(ops._ptr(q), ops._ptr(k), ops._ptr(v), ops._ptr(o), ops._ptr(lse), ops._ptr(acc),
                                        B, H, N, T, d, scale, dt, impl, st()), "ga_cross_attn_fwd")

    def bwd(i):
        q, k, v, o, lse, acc, do, dq = sets[i]
        abi.check(lib.ga_cross_attn_bwd(ops._ptr(q), ops._ptr(k), ops._ptr(v), ops._ptr(lse), ops._ptr(do),
                                        ops._ptr(dacc), 0, dacc.shape[1] if dacc is not None else T, ops._ptr(dq), None, None, B, H, N, T, d, scale, dt, impl,
                                        st()), "ga_cross_attn_bwd")
    if direction == "bwd":
        for i in range(n_sets):
            fwd(i)                      # valid LSE for the backward
    us = _time_graph(fwd if direction == "fwd" else bwd, n_sets)
    return {"kernel": f"cross_attn_{direction}", "B": B, "H": H, "N": N, "T": T, "d": d, "dtype": str(dtype),
            "with_maps": with_acc, "impl": impl, "us": us, "bytes": nbytes, "gbs": nbytes / us / 1e3,
            "buffer_sets": n_sets}


def time_self_attn(B, H, N, d, dtype=torch.float16, direction="fwd", device="cuda:0"):
    """One fused self-attention forward launch (or the backward's two launches).  Tensor-core bound: reports TFLOP/s of
    the GEMM work the kernels issue (`ops.self_attn_flops`) and of the algorithm's minimum (2 GEMMs fwd, 5 bwd)."""
    lib = abi.load()
    flops = ops.self_attn_flops(B, H, N, d, direction)
    need = 2 * B * H * N * N * d * (2 if direction == "fwd" else 5)
    nbytes = B * N * H * d * 2 * (4 if direction == "fwd" else 8)
    n_sets = max(2, min(16, (2 * L2_BYTES) // max(nbytes, 1) + 1))
    g = torch.Generator(device=device).manual_seed(0)
    sets = []
    for _ in range(n_sets):
        q, k, v, do = (torch.randn(B, N, H * d, device=device, dtype=dtype, generator=g) for _ in range(4))
        o, lse = ops.self_attention_forward(q, k, v, H, d ** -0.5)
        sets.append((q, k, v, o, lse, do, torch.empty_like(q), torch.empty_like(q), torch.empty_like(q),
                     torch.empty(B, H, N, device=device, dtype=torch.float32)))
    dt, scale = ops._DTYPES[dtype], d ** -0.5
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)   # noqa: E731

    def fwd(i):
        q, k, v, o, lse = sets[i][:5]
        abi.check(lib.ga_self_attn_fwd(ops._ptr(q), ops._ptr(k), ops._ptr(v), ops._ptr(o), ops._ptr(lse), B, H, N, d,
                                       scale, dt, st()), "ga_self_attn_fwd")

    def bwd(i):
        q, k, v, o, lse, do, dq, dk, dv, dvec = sets[i]
        abi.check(lib.ga_self_attn_bwd(ops._ptr(q), ops._ptr(k), ops._ptr(v), ops._ptr(o), ops._ptr(lse), ops._ptr(do),
                                       ops._ptr(dq), ops._ptr(dk), ops._ptr(dv), ops._ptr(dvec), B, H, N, d, scale, dt,
                                       st()), "ga_self_attn_bwd")
    us = _time_graph(fwd if direction == "fwd" else bwd, n_sets)
    return {"kernel": f"self_attn_{direction}", "B": B, "H": H, "N": N, "d": d, "dtype": str(dtype), "us": us,
            "flops_issued": flops, "tflops_issued": flops / us / 1e6, "tflops_algorithmic": need / us / 1e6,
            "bytes": nbytes, "buffer_sets": n_sets}


def sweep_self(device="cuda:0"):
    for (N, d) in ((4096, 40), (1024, 80), (256, 160), (64, 160)):
        for B in (1, 2, 8):
            for direction in ("fwd", "bwd"):
                yield time_self_attn(B, 8, N, d, torch.float16, direction, device)


def time_tail(res, n_layers, slices_per_layer, n_samples=1, T=77, direction="fwd", device="cuda:0"):
    """One guidance-tail launch (forward or backward): `n_samples` independent evaluations, each over `n_layers`
    accumulators of `slices_per_layer` (N, T) slices (the reference's shape is n_samples = 1, 5 layers, 1-2 slices)."""
    from .pipeline_guided_attention import GuidedAttention
    from .run import setup_prompt
    cfg = setup_prompt(DEFAULT_PROMPT)
    pipe = GuidedAttention(unet=None, tokenizer=cfg.stable.tokenizer)
    pipe.prompt = cfg.prompt
    npix, S = res * res, n_samples
    spec = pipe._tail_spec(res, T, True, 0.5, 3, False, torch.device(device))
    n_maps = n_layers * slices_per_layer
    tp, nt = spec.last - spec.first, spec.params.n_tokens
    fwd_bytes = S * ((n_maps * npix * T + npix * tp + nt * npix) * 4 + nt * npix)
    bwd_bytes = S * (npix * tp + npix * T) * 4
    nbytes = fwd_bytes if direction == "fwd" else bwd_bytes
    n_sets = max(1, min(16, (2 * L2_BYTES) // max(fwd_bytes, 1) + 1)) if S > 1 else 1
    sets = [[torch.rand(S * slices_per_layer, npix, T, device=device).softmax(-1) for _ in range(n_layers)]
            for _ in range(n_sets)]
    lib = abi.load()
    p = abi.GaTailParams.from_buffer_copy(spec.params)
    p.inv_count, p.n_samples = 1.0 / n_maps, S
    if direction == "fwd":
        def launch(i):
            ops.guidance_tail(spec, sets[i], n_maps, S)
    else:
        saved = []
        for accs in sets:
            outs = ops.guidance_tail(spec, accs, n_maps, S)
            saved.append(tuple(o.detach() for o in outs[:4]) + (torch.empty(S, npix, 80, device=device),))
        g_total = torch.ones(S, device=device)

        def launch(i):   # the C ABI directly: an autograd backward of a graph built outside the capture cannot be captured
            attn_text, smoothed, stats, argmax, d_abar = saved[i]
            abi.check(lib.ga_guidance_tail_bwd(C.byref(p), spec.tokens, ops._ptr(spec.masks), ops._ptr(spec.weights),
                                               ops._ptr(attn_text), ops._ptr(smoothed), ops._ptr(stats),
                                               ops._ptr(argmax), ops._ptr(g_total), None, None, ops._ptr(d_abar), 80,
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                      "ga_guidance_tail_bwd")
    with torch.no_grad():
        us = _time_graph(launch, n_sets)
    return {"kernel": f"guidance_tail_{direction}", "res": res, "layers": n_layers, "slices": slices_per_layer,
            "samples": S, "us": us, "bytes": nbytes, "gbs": nbytes / us / 1e3, "buffer_sets": n_sets}


def time_group_norm(n, c, h, w, groups=32, silu=True, direction="fwd", impl="fused", dtype=torch.float16,
                    device="cuda:0"):
    """GroupNorm (+ SiLU) on a channels-last (n, c, h, w) activation: the fused kernels (`impl="fused"`, two launches
    per direction) or PyTorch's own ops (`impl="torch"`: the 3 + 1 kernels forward, 5 backward the stock UNet runs).
    `direction="fwdbwd"` times forward + backward to x."""
    import torch.nn.functional as F
    nbytes = n * c * h * w * 2 * (2 if direction == "fwd" else 5)     # compulsory traffic: x -> y; x, dy -> dx
    n_sets = max(2, min(32, (2 * L2_BYTES) // max(n * c * h * w * 2 * 3, 1) + 1))
    g = torch.Generator(device=device).manual_seed(0)
    xs = [torch.randn(n, c, h, w, device=device, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
          for _ in range(n_sets)]
    dy = torch.randn(n, c, h, w, device=device, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    wgt = torch.ones(c, device=device, dtype=dtype)
    b = torch.zeros(c, device=device, dtype=dtype)

    def run(x):
        if impl == "fused":
            return ops.group_norm(x, wgt, b, groups, 1e-5, silu=silu)
        y = F.group_norm(x, groups, wgt, b, 1e-5)
        return F.silu(y) if silu else y

    def launch(i):
        if direction == "fwd":
            with torch.no_grad():
                run(xs[i])
        else:
            x = xs[i].detach().requires_grad_(True)
            torch.autograd.grad(run(x), x, dy)
    us = _time_graph(launch, n_sets)
    return {"kernel": f"group_norm_{direction}", "impl": impl, "n": n, "c": c, "hw": h * w, "silu": silu, "us": us,
            "bytes": nbytes, "gbs": nbytes / us / 1e3, "buffer_sets": n_sets}


def time_geglu(rows, inner, direction="fwd", impl="fused", dtype=torch.float16, device="cuda:0"):
    """The feed-forward gate `h * gelu(gate)` on a (1, rows, 2 * inner) projection: fused kernel vs PyTorch's ops."""
    import torch.nn.functional as F
    nbytes = rows * inner * 2 * (3 if direction == "fwd" else 8)
    n_sets = max(2, min(32, (2 * L2_BYTES) // max(rows * inner * 2 * 3, 1) + 1))
    g = torch.Generator(device=device).manual_seed(0)
    ps = [torch.randn(1, rows, 2 * inner, device=device, generator=g).to(dtype) for _ in range(n_sets)]
    d_out = torch.randn(1, rows, inner, device=device, generator=g).to(dtype)

    def run(p):
        if impl == "fused":
            return ops.geglu(p)
        h, gate = p.chunk(2, dim=-1)
        return h * F.gelu(gate)

    def launch(i):
        if direction == "fwd":
            with torch.no_grad():
                run(ps[i])
        else:
            p = ps[i].detach().requires_grad_(True)
            torch.autograd.grad(run(p), p, d_out)
    us = _time_graph(launch, n_sets)
    return {"kernel": f"geglu_{direction}", "impl": impl, "rows": rows, "inner": inner, "us": us, "bytes": nbytes,
            "gbs": nbytes / us / 1e3, "buffer_sets": n_sets}


def sweep_group_norm(device="cuda:0"):
    """The GroupNorm shapes of one SD-1.4 UNet pass at batch 1 / 2, and at the seed-batched 8."""
    for n in (1, 2, 8):
        for (c, r) in ((320, 64), (640, 32), (1280, 16), (1280, 8), (2560, 8), (1920, 16), (960, 32), (640, 64)):
            for direction in ("fwd", "fwdbwd"):
                for impl in ("torch", "fused"):
                    yield time_group_norm(n, c, r, r, 32, True, direction, impl, device=device)


def sweep(device="cuda:0"):
    """BASELINE config 3: H=8, T=77, res 16 (d=160) and 32 (d=80), batch 2, plus the batch sweep and the 64^2 level."""
    for (N, d) in ((256, 160), (1024, 80), (4096, 40)):
        for B in (1, 2, 8, 32, 128, 512):
            if B * N * 8 * d * 2 * 8 > 8e9:      # keep the rotating buffer sets within a few GB
                continue
            for direction in ("fwd", "bwd"):
                yield time_cross_attn(B, 8, N, 77, d, torch.float16, with_acc=N <= 1024, direction=direction,
                                      device=device)
    for res in (16, 32):
        for S in (1, 8, 64, 512, 2048):
            for direction in ("fwd", "bwd"):
                yield time_tail(res, 5, 2, n_samples=S, direction=direction, device=device)


def single(argv):
    """`--single fwd|bwd B N d [maps]`: a handful of plain (non-graph) launches of one configuration, for ncu."""
    direction, B, N, d = argv[0], int(argv[1]), int(argv[2]), int(argv[3])
    with_maps = len(argv) > 4 and argv[4] == "maps"
    print(json.dumps(time_cross_attn(B, 8, N, 77, d, torch.float16, with_acc=with_maps, direction=direction)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--single":
        single(sys.argv[2:])
    elif len(sys.argv) > 1 and sys.argv[1] == "--single-tail":
        a = sys.argv[2:]
        print(json.dumps(time_tail(int(a[1]), 5, 2, n_samples=int(a[2]), direction=a[0])))
    elif len(sys.argv) > 1 and sys.argv[1] == "--single-self":
        a = sys.argv[2:]
        print(json.dumps(time_self_attn(int(a[1]), 8, int(a[2]), int(a[3]), torch.float16, a[0])))
    elif len(sys.argv) > 1 and sys.argv[1] == "--saturated":       # bench.py's GPU-filling shapes, one JSON line each
        dev_ = sys.argv[2] if len(sys.argv) > 2 else "cuda:0"
        torch.cuda.set_device(torch.device(dev_))
        for direction in ("fwd", "bwd"):
            for (N, d, maps, B) in ((1024, 80, True, 256), (256, 160, True, 512), (4096, 40, False, 128)):
                print(json.dumps(time_cross_attn(B, 8, N, 77, d, torch.float16, with_acc=maps, direction=direction,
                                                 device=dev_)))
            for res in (16, 32):
                print(json.dumps(time_tail(res, 5, 2, n_samples=2048 if res == 16 else 512, direction=direction,
                                           device=dev_)))
            sys.stdout.flush()
    elif len(sys.argv) > 1 and sys.argv[1] == "--single-group-norm":      # direction n c r  (for ncu)
        a = sys.argv[2:]
        print(json.dumps(time_group_norm(int(a[1]), int(a[2]), int(a[3]), int(a[3]), 32, True, a[0], "fused")))
    elif len(sys.argv) > 1 and sys.argv[1] == "--single-geglu":           # direction rows inner  (for ncu)
        a = sys.argv[2:]
        print(json.dumps(time_geglu(int(a[1]), int(a[2]), a[0], "fused")))
    elif len(sys.argv) > 1 and sys.argv[1] == "--geglu":
        for rows, inner in ((4096, 1280), (1024, 2560), (256, 5120), (64, 5120), (8192, 1280)):
            for direction in ("fwd", "fwdbwd"):
                for impl in ("torch", "fused"):
                    print(json.dumps(time_geglu(rows, inner, direction, impl)))
                    sys.stdout.flush()
    elif len(sys.argv) > 1 and sys.argv[1] == "--group-norm":
        for r in sweep_group_norm():
            print(json.dumps(r))
            sys.stdout.flush()
    elif len(sys.argv) > 1 and sys.argv[1] == "--self":
        for r in sweep_self():
            print(json.dumps(r))
            sys.stdout.flush()
    else:
        for r in sweep():
            print(json.dumps(r))
            sys.stdout.flush()
