"""Which ATen ops (and which shapes) make up one guided image besides the GEMMs / convolutions?  Runs the eager pipeline
(2 denoising steps of BASELINE config 2: step 0 carries the refinement rounds) under torch.profiler and prints the CUDA
time by (op, input shapes).  Usage (GPU box): python tools/profile_ops.py [top_n] > gpurun_out/profile_ops.txt"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    top = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    args = argparse.Namespace(unet="sd14", no_graphs=True, host_control=True, denoise_steps=2)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg, pipe, store, embeds_host = bench.build_pipeline(args, dev)
    pipe.use_cuda_graphs = False
    embeds = embeds_host.to(dev)

    from torch.profiler import profile, ProfilerActivity

    class Stop(Exception):
        pass

    def cb(i, t, latents):
        if i >= 1:
            raise Stop()

    def run():
        gen = torch.Generator("cpu").manual_seed(28)
        lat = torch.randn(1, 4, 64, 64, generator=gen).to(dev)
        try:
            pipe(prompt=cfg.prompt, attention_store=store, attention_res=16, guidance_scale=7.5, generator=gen,
                 latents=lat, prompt_embeds=embeds[1:2], negative_prompt_embeds=embeds[0:1],
                 num_inference_steps=50, thresholds=cfg.thresholds, scale_factor=20, scale_range=(1., .5),
                 smooth_attentions=True, sigma=0.5, kernel_size=3, sd_2_1=False, output_type="latent", callback=cb,
                 callback_steps=1)
        except Stop:
            pass
    run()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        run()
        torch.cuda.synchronize()
    ka = prof.key_averages(group_by_input_shape=True)
    rows = sorted(ka, key=lambda e: -e.self_device_time_total)
    total = sum(e.self_device_time_total for e in rows)
    print(f"total self CUDA time {total / 1e3:.1f} ms")
    for e in rows[:top]:
        print(f"{100 * e.self_device_time_total / total:5.1f}%  {e.self_device_time_total / 1e3:8.2f} ms  n={e.count:5d}  "
              f"{e.key[:44]:44s} {str(e.input_shapes)[:150]}")
    print("---- by kernel name")
    kern = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            k = kern.setdefault(ev.name[:90], [0, 0.0])
            k[0] += 1
            k[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
    tk = sum(v[1] for v in kern.values()) or 1
    for name, (n, t) in sorted(kern.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100 * t / tk:5.1f}%  {t / 1e3:8.2f} ms  n={n:5d}  {name}")


if __name__ == "__main__":
    main()
