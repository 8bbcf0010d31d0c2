"""CPU: host-side mirror of the reference API (parser, token table, geometry, store bookkeeping, thresholds) against the
golden fixtures from the unmodified reference, and the C-ABI library's exported surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from guided_attention_b200 import _cabi, helpers, shared_state as S
from guided_attention_b200 import run as R
from guided_attention_b200.helpers import AnnotationType as AT
from tests.gpu_harness import setup_prompt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _meta_json(meta_info):
    out = []
    for sub, kind, payload in meta_info:
        if kind == AT.BOX:
            payload = [payload.x, payload.y, payload.width, payload.height]
        elif kind == AT.COOR:
            payload = list(payload)
        out.append([sub, kind.name, payload])
    return out


def test_parse_and_token_dict_bit_exact(kat):
    for rec in kat["parse"]:
        if "error" in rec:
            with pytest.raises(Exception) as ei:
                setup_prompt(rec["meta_prompt"])
            assert type(ei.value).__name__ == rec["error"], rec["meta_prompt"]
            continue
        cfg = setup_prompt(rec["meta_prompt"])
        assert cfg.prompt == rec["prompt"], rec["meta_prompt"]
        assert _meta_json(cfg.meta_info) == rec["meta_info"]
        assert {k: v[1] for k, v in cfg.custom_loss.items()} == rec["custom"]
        td = {str(k): {"word": v["word"], "kind": v["loss_type"].name, "subprompt": v["subprompt"]}
              for k, v in cfg.token_dict.items()}
        assert td == rec["token_dict"]
        assert list(td.keys()) == list(rec["token_dict"].keys())   # insertion order = output order


def test_parser_fuzz_against_reference():
    """400 seeded random meta-prompts (valid and malformed) through parse_prompt + parseMetaPrompt +
    get_indices; prompt, meta_info, custom-loss table, token table (values and order) and the exception TYPE all equal
    what the unmodified reference produced (tests/golden/reference_parse_fuzz.json, oracle/gen_golden.py)."""
    import json
    import os
    from oracle.cases import fuzz_parse_cases
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_parse_fuzz.json")) as f:
        recs = json.load(f)["cases"]
    assert [r["meta_prompt"] for r in recs] == fuzz_parse_cases()      # the fixture matches the seeded generator
    n_err = 0
    for rec in recs:
        if "error" in rec:
            n_err += 1
            with pytest.raises(Exception) as ei:
                setup_prompt(rec["meta_prompt"])
            assert type(ei.value).__name__ == rec["error"], rec["meta_prompt"]
            continue
        cfg = setup_prompt(rec["meta_prompt"])
        assert cfg.prompt == rec["prompt"], rec["meta_prompt"]
        assert _meta_json(cfg.meta_info) == rec["meta_info"], rec["meta_prompt"]
        assert {k: v[1] for k, v in cfg.custom_loss.items()} == rec["custom"], rec["meta_prompt"]
        td = {str(k): {"word": v["word"], "kind": v["loss_type"].name, "subprompt": v["subprompt"]}
              for k, v in cfg.token_dict.items()}
        assert td == rec["token_dict"] and list(td.keys()) == list(rec["token_dict"].keys()), rec["meta_prompt"]
    assert 10 < n_err < len(recs) - 100        # the fuzz exercises both the error paths and the happy path


def test_shipped_hyperparameters():
    """effective thresholds and hyper-parameters after overrideConfig."""
    cfg = setup_prompt()
    assert cfg.thresholds == {0: 1.0}
    assert S.curHyperParams == {"strict": False, "inside_loss_scale": .2, "outside_loss_scale": .2,
                                "shrink_factor": .15, "thresholds": {0: 1.}, "use_optimizer": False,
                                "recurse_until": 14, "recurse_steps": 3}


def test_host_masks_and_rect_bit_exact(kat):
    for rec in kat["masks"]:
        S.curHyperParams = {"shrink_factor": rec["shrink"]}
        r = helpers.Rect(*rec["box"], 1).of_size(float(rec["res"]) if rec["res"] == 16 else rec["res"])
        assert [r.x, r.y, r.width, r.height] == rec["rect_at_res"]
        m = helpers.box_mask_host(r, rec["res"])
        assert ["".join(map(str, row)) for row in m.tolist()] == rec["mask_rows"]


def test_gaussian_taps_match_reference_kernel(kat):
    from guided_attention_b200 import ops
    for rec in kat["gaussian"]:
        w = np.array(ops.gaussian_taps(rec["kernel_size"], rec["sigma"]))
        np.testing.assert_allclose(np.outer(w, w), np.array(rec["weight"]), rtol=2e-6)
    with pytest.raises(NotImplementedError):
        ops.gaussian_taps(5, 0.5)


def test_find_matching_bracket():
    assert helpers.findMatchingBracket("robot:.1,.2]") == 11
    assert helpers.findMatchingBracket("a [b] c] d") == 7
    assert helpers.findMatchingBracket("no close") == -1


def test_store_bookkeeping_matches_reference_counts():
    """Layer counter / between_steps / key layout (reference utils/ptp_utils.py:194-243) with shape-only stand-ins."""
    from guided_attention_b200.ptp_utils import AttentionStore, _ShapeOnly, HeadSummedMaps
    setup_prompt()
    store = AttentionStore()
    store.num_att_layers = 4
    acc = torch.zeros(1, 256, 77)
    q = torch.zeros(1, 256, 16)
    k = torch.zeros(1, 77, 16)
    m = HeadSummedMaps(acc, 8, q, k, 0.25)
    assert tuple(m.shape) == (8, 256, 77) and m.n_maps == 8
    store(m, True, "down")
    store(_ShapeOnly(8, 4096, 77), True, "up")          # too large: not kept
    store(torch.zeros(8, 256, 256), False, "down")      # self maps are not kept by default
    assert store.cur_step == 0 and store.attention_store == {}
    store(m, True, "mid")
    assert store.cur_step == 1 and store.cur_att_layer == 0
    got = {k2: len(v) for k2, v in store.get_average_attention().items()}
    assert got == {"down_cross": 1, "mid_cross": 1, "up_cross": 0, "down_self": 0, "mid_self": 0, "up_self": 0}
    assert store.step_store == AttentionStore.get_empty_store()


def test_register_attention_control_counts_32_layers():
    from guided_attention_b200.ptp_utils import AttentionStore, register_attention_control, \
        AttendExciteCrossAttnProcessor
    from guided_attention_b200.substrate import UNetConfig, build_unet
    import types
    unet = build_unet(UNetConfig.tiny())
    store = AttentionStore()
    register_attention_control(types.SimpleNamespace(unet=unet), store)
    assert store.num_att_layers == 32
    procs = unet.attn_processors
    assert all(isinstance(p, AttendExciteCrossAttnProcessor) for p in procs.values())
    places = {n.split(".")[0]: p.place_in_unet for n, p in procs.items()}
    assert places == {"down_blocks": "down", "mid_block": "mid", "up_blocks": "up"}


def test_meets_threshold_truth_table(kat):
    from guided_attention_b200.pipeline_guided_attention import GuidedAttention
    pipe = GuidedAttention(unet=None)
    for rec in kat["loss"]:
        case_prompt = next(c for c in __import__("oracle.cases", fromlist=["LOSS_CASES"]).LOSS_CASES
                           if c["name"] == rec["name"])
        setup_prompt(case_prompt["meta_prompt"], case_prompt.get("hyper"), case_prompt.get("cfg"))
        unscaled = [(k, v) for k, v in rec["unscaled"]]
        thr = {int(k): v for k, v in rec["thresholds"].items()}
        for i, want in rec["meets"].items():
            assert pipe.meets_threshold(int(i), thr, unscaled) == want, (rec["name"], i)


def test_library_exports_every_declared_symbol():
    """No compute calls here (no GPU): the library loads and exports exactly what include/guided_attn.h declares."""
    header = open(os.path.join(ROOT, "include", "guided_attn.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(ga_\w+)\(", header, flags=re.M))
    assert declared == set(_cabi.PROTOTYPES), declared ^ set(_cabi.PROTOTYPES)
    lib = _cabi.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ga_version() == _cabi.GA_ABI_VERSION
    assert ctypes.sizeof(_cabi.GaToken) == 32 and ctypes.sizeof(_cabi.GaTailParams) == 64


def test_ops_refuse_cpu_tensors():
    from guided_attention_b200 import ops
    q = torch.zeros(1, 64, 16)
    with pytest.raises(_cabi.GuidedAttnLibraryError):
        ops.cross_attention(q, torch.zeros(1, 77, 16), torch.zeros(1, 77, 16), 2, 0.3)


def test_missing_library_is_loud(tmp_path):
    with pytest.raises(_cabi.GuidedAttnLibraryError):
        _cabi.load(str(tmp_path / "nope.so"))


def test_fused_unet_hooks_leave_cpu_and_fp32_untouched():
    """`register_fused_norms` only reroutes 16-bit CUDA activations: on CPU / fp32 the patched UNet is bit-identical to
    the stock one (the oracle and the fp32 goldens never see the fused kernels), and registration is idempotent."""
    from guided_attention_b200 import ptp_utils
    from guided_attention_b200.substrate import UNetConfig, build_unet
    cfg = UNetConfig.tiny(sample_size=16)
    stock = build_unet(cfg, seed=0).requires_grad_(False)
    fused = build_unet(cfg, seed=0).requires_grad_(False)
    assert ptp_utils.register_fused_norms(fused) == 61 and ptp_utils.register_fused_norms(fused) == 61
    blocks = [m for m in fused.modules() if hasattr(m, "_ga_temb_shifts")]
    assert len(blocks) == 22 and len({id(b._ga_temb_shifts) for b in blocks}) == 1
    # registering again (run.py does it for every seed) keeps the SAME registry and the SAME buffers: captured CUDA
    # graphs hold their addresses; an in-place weight change refreshes them in place
    reg0 = blocks[0]._ga_temb_shifts
    probe = torch.randn(1, blocks[0].time_emb_proj.in_features)
    reg0.shift(blocks[0], probe)
    w_ptr, b_ptr = reg0.weight.data_ptr(), reg0.bias.data_ptr()
    ptp_utils.register_fused_norms(fused)
    assert blocks[0]._ga_temb_shifts is reg0
    saved = blocks[3].conv1.bias.detach().clone()
    with torch.no_grad():
        blocks[3].conv1.bias.add_(1.0)
    reg0.shift(blocks[3], torch.randn(1, blocks[0].time_emb_proj.in_features))
    assert (reg0.weight.data_ptr(), reg0.bias.data_ptr()) == (w_ptr, b_ptr)
    with torch.no_grad():
        blocks[3].conv1.bias.copy_(saved)
    g = torch.Generator().manual_seed(0)
    x, e = torch.randn(2, 4, 16, 16, generator=g), torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    assert torch.equal(stock(x, 10, encoder_hidden_states=e).sample, fused(x, 10, encoder_hidden_states=e).sample)
    # the one-GEMM time-embedding projection equals the per-block projection (+ conv1 bias) it replaces
    reg, temb = blocks[0]._ga_temb_shifts, torch.randn(2, blocks[0].time_emb_proj.in_features, generator=g)
    for b in (blocks[0], blocks[7], blocks[-1]):
        ref = torch.nn.functional.linear(torch.nn.functional.silu(temb), b.time_emb_proj.weight,
                                         b.time_emb_proj.bias + b.conv1.bias)
        assert torch.allclose(reg.shift(b, temb), ref, rtol=0, atol=1e-6)


def test_group_norm_plan_without_a_gpu():
    """`ga_group_norm_ws_bytes` runs the kernels' host-side decomposition (pixel chunks x channel slices) and needs no
    device: chunk counts for the UNet's shapes and the declined shapes."""
    lib = _cabi.load()
    ws = lib.ga_group_norm_ws_bytes
    per = 32 * 2 * 4                                     # groups x (mean, M2) x fp32 per chunk and sample
    assert ws(1, 4096, 320, 32) == 128 * per            # 40 vectors, 6 pixel lanes -> 32-pixel chunks
    assert ws(8, 4096, 640, 32) == 8 * 128 * per
    assert ws(1, 64, 2560, 32) == 64 * per              # 320 vectors -> two channel slices of 16 groups, one pixel per CTA
    assert ws(1, 64, 1280, 32) == 64 * per
    assert ws(1, 9216, 320, 32) == 128 * per            # SD-2.x 96x96 latent: 72-pixel chunks
    assert ws(1, 35, 64, 8) == 2 * 8 * 2 * 4            # ragged: 32-pixel chunk + 3-pixel chunk
    assert ws(3, 1, 128, 8) == 3 * 1 * 8 * 2 * 4
    assert ws(1, 64, 36, 6) == -1                       # channels not a multiple of 8
    assert ws(1, 64, 40, 8) == -1                       # 5-channel groups: a 128-bit vector would touch three groups
    assert ws(1, 64, 320, 3) == -1                      # channels not divisible by the group count
    assert ws(0, 64, 320, 32) == -1


def test_unet_side_entry_points_validate_arguments_without_a_gpu():
    """Argument checks of the UNet-side entry points happen before any CUDA call: status code + thread-local message."""
    lib = _cabi.load()
    buf = (ctypes.c_char * 4096)()
    base = ctypes.addressof(buf)
    a16 = (base + 15) // 16 * 16                         # 16-byte aligned address inside the buffer
    F16, F32 = _cabi.GA_F16, _cabi.GA_F32
    BAD_ARG, ALIGN = -1, -3
    p = ctypes.c_void_p
    # NULL operands / wrong dtype / misaligned pointers
    assert lib.ga_group_norm_fwd(None, None, 0, p(a16), p(a16), p(a16), p(a16), p(a16), 1, 64, 320, 32, 1e-5, 1, F16, None) == BAD_ARG
    assert lib.ga_group_norm_fwd(p(a16), None, 0, p(a16), p(a16), p(a16), p(a16), p(a16), 1, 64, 320, 32, 1e-5, 1, F32, None) == BAD_ARG
    assert b"16-bit" in lib.ga_last_error()
    assert lib.ga_group_norm_fwd(p(a16 + 2), None, 0, p(a16), p(a16), p(a16), p(a16), p(a16), 1, 64, 320, 32, 1e-5, 1, F16, None) == ALIGN
    # a shift row stride that is not a multiple of 8 elements (or shorter than a row) is refused
    assert lib.ga_group_norm_fwd(p(a16), p(a16), 324, p(a16), p(a16), p(a16), p(a16), p(a16), 1, 64, 320, 32, 1e-5, 1, F16, None) == BAD_ARG
    assert lib.ga_group_norm_bwd(p(a16), p(a16), 312, p(a16), p(a16), p(a16), p(a16), p(a16), p(a16), 1, 64, 320, 32, 1, F16, None) == BAD_ARG
    assert lib.ga_add_bias_residual(p(a16), None, p(a16), p(a16), 64, 36, F16, None) == BAD_ARG       # channels % 8
    assert lib.ga_add_bias_residual(p(a16), p(a16 + 4), p(a16), p(a16), 64, 320, F16, None) == ALIGN
    assert lib.ga_geglu_fwd(p(a16), p(a16), 64, 12, F16, None) == BAD_ARG                              # inner % 8
    assert lib.ga_geglu_bwd(p(a16), None, p(a16), 64, 1280, F16, None) == BAD_ARG
    assert lib.ga_layer_norm_fwd(p(a16), p(a16), p(a16), p(a16), p(a16), p(a16), 64, 4096, 1e-5, F16, None) == BAD_ARG
    assert b"2048" in lib.ga_last_error()
    # zero rows: nothing to launch, no error
    assert lib.ga_geglu_fwd(p(a16), p(a16), 0, 1280, F16, None) == 0
    assert lib.ga_layer_norm_fwd(p(a16), p(a16), p(a16), p(a16), p(a16), p(a16), 0, 320, 1e-5, F16, None) == 0
