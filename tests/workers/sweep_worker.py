"""Worker of tests/test_gpu_sweep.py: BASELINE config 5 in miniature through the product's own entry point
(`run.execute`, the mirror of reference run.py:93-135): the seeds of `--seeds` on the SD-1.4-shaped fp16 UNet with the
bench's hyper-parameters, sharded `seed_idx % world_size` when launched under torchrun.  Rank 0 saves the gathered
per-seed latents."""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--seeds", type=int, nargs="+", required=True)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--unet", default="sd14")
    a = ap.parse_args()
    from guided_attention_b200 import run as R, shared_state as S, sweep
    from guided_attention_b200.config import RunConfig
    from guided_attention_b200.substrate import UNetConfig
    rank, world, local_rank = sweep.init_distributed() if sweep.dist_env()[1] > 1 else (0, 1, 0)
    dev = sweep.local_device(local_rank)
    torch.cuda.set_device(dev)
    S.hyperParameterOverrides = {"strict": False, "inside_loss_scale": .2, "outside_loss_scale": .2,
                                 "shrink_factor": .15, "thresholds": {0: .4, 2: .8, 4: .9, 8: .9},
                                 "use_optimizer": False, "recurse_until": 14, "recurse_steps": 3}
    cfg = RunConfig(meta_prompt='a [robot:.6,.3,.4,.55] and a [blue vase:.2,.3,.4,.55]', seeds=list(a.seeds),
                    half_precision=True, n_inference_steps=a.steps, output_path=tempfile.mkdtemp(prefix="ga_sweep_"))
    R.setup(cfg, device=dev, unet_config=UNetConfig.sd14() if a.unet == "sd14" else UNetConfig.tiny())
    R.register_custom_loss("toLeftOf", R.ToLeftOf())
    assert cfg.stable.use_cuda_graphs                 # the entry point runs the graphed loop on CUDA
    images = R.execute(cfg, output_type="latent", save=False)
    full = torch.cat([im.reshape((1,) + tuple(im.shape[-3:])) for im in images]).cpu()
    if rank == 0:
        torch.save({"latents": full, "world": world, "seeds": list(a.seeds),
                    "passes": dict(cfg.stable.pass_counts)}, a.out)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
