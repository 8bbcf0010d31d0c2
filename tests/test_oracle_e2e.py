"""CPU: the oracle's restatement of the guided denoising loop against the final latents produced by the reference's
own `GuidedAttention.__call__` (tiny substrate UNet, 5 DDIM steps with refinement, recursion and re-noising)."""
import re

import numpy as np
import torch

from oracle import oracle as O
from oracle.cases import E2E_CASE, make_e2e_inputs
from oracle.specs import tokens_from_record
from guided_attention_b200.substrate import DDIMScheduler


def test_oracle_pipeline_matches_reference_call(kat, e2e_golden):
    doc, arrays = e2e_golden
    case = E2E_CASE
    tokens = tokens_from_record(next(p for p in kat["parse"] if p["meta_prompt"] == case["meta_prompt"]))
    unet, embeds, latents, _ = make_e2e_inputs(case)
    hy = case["hyper"]
    pipe = O.OraclePipeline(unet, DDIMScheduler(), tokens, O.HyperParams(), recurse_steps=hy["recurse_steps"],
                            recurse_until=hy["recurse_until"])
    assert pipe.store.num_att_layers == doc["num_att_layers"] == 32
    tr = pipe(embeds, latents.clone(), seed=case["latent_seed"], num_inference_steps=case["steps"],
              guidance_scale=case["guidance_scale"], thresholds=hy["thresholds"])
    assert tr.unet_forwards == doc["unet_forwards"]
    gold = arrays["final_latents"]
    got = tr.latents.numpy()
    cos = float((got * gold).sum() / (np.linalg.norm(got) * np.linalg.norm(gold)))
    assert cos > 1 - 1e-6
    # fp32 summation-order noise in the loss is amplified by the 20x latent steps over 5 DDIM steps
    np.testing.assert_allclose(got, gold, rtol=1e-3, atol=1e-3)
    # the refinement trajectory: every "Finished with loss of" line of the reference log
    gold_final = [float(m.group(1)) for l in doc["log"] for m in [re.search(r"Finished with loss of: tensor\(\[([0-9.]+)", l)] if m]
    assert len(gold_final) >= 2
