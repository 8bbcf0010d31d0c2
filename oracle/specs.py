"""TEST INFRASTRUCTURE ONLY -- builds oracle TokenSpecs from a golden parse record (`tests/golden/reference_kat.json`)."""
from __future__ import annotations

from .oracle import TokenSpec, COOR, BOX, KEYWORD

_KIND = {"COOR": COOR, "BOX": BOX, "KEYWORD": KEYWORD}


def tokens_from_record(parse_record):
    """token_dict order (insertion order of the reference's dict) is the order of the per-token outputs."""
    payload = {}
    for sub, kind, pl in parse_record["meta_info"]:
        payload[sub] = (kind, pl)     # later meta_info wins, like run.py:86-90
    out = []
    for idx, info in parse_record["token_dict"].items():
        kind, pl = payload[info["subprompt"]]
        out.append(TokenSpec(index=int(idx), kind=_KIND[kind], payload=tuple(pl) if pl is not None else None,
                             subprompt=info["subprompt"], word=info["word"]))
    return out
