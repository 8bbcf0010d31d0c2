"""Host-side prompt front-end and annotation geometry -- mirror of the reference's `utils/helpers.py`.

Same public names and behaviour (bug-for-bug where behaviour is observable) so call sites written against the
reference keep working:

  * `AnnotationType`, `Rect`                      reference utils/helpers.py:10-30
  * `add_word`, `findMatchingBracket`, `parse_prompt`   reference utils/helpers.py:33-114
  * `inside_box`, `distance_from_center`, `distance_from_bounding_box`, `get_corresponding_weight`
                                                  reference utils/helpers.py:158-213
  * `calculate_bounding_box_losses`               reference utils/helpers.py:215-277 (CUDA op here, no Python loops)
  * `log`, `log_clear`, `log_save`                reference utils/helpers.py:291-307

The pixel tests are pure Python float64 exactly like the reference; the device rasteriser
(`ops.rasterize_boxes`, csrc/guidance_tail.cu) must agree with `inside_box` bit for bit and the tests check that.
"""
from __future__ import annotations

import math
import os
from enum import Enum

import numpy as np

from . import shared_state as state


class AnnotationType(Enum):
    COOR = 0
    BOX = 1
    KEYWORD = 2


class Rect:
    """Axis-aligned box: top-left corner + extent, expressed on a grid of `size` cells per side."""

    def __init__(self, x, y, width, height, size):
        self.x, self.y, self.width, self.height, self.size = x, y, width, height, size

    def right(self):
        return self.x + self.width

    def bottom(self):
        return self.y + self.height

    def center(self):
        return (self.x + self.width / 2.0, self.y + self.height / 2.0)

    def of_size(self, new_size):
        # one multiplication per field by a single precomputed ratio: the rounding of e.g. .6*16 matters for the mask
        ratio = float(new_size / self.size)
        return Rect(self.x * ratio, self.y * ratio, self.width * ratio, self.height * ratio, new_size)

    def as_tuple(self):
        return (self.x, self.y, self.width, self.height)

    def __repr__(self):
        return f"Rect(x={self.x}, y={self.y}, width={self.width}, height={self.height}, size={self.size})"


# ------------------------------------------------------------------------------------------- meta-prompt grammar
def add_word(prompt, token):
    if prompt == "" or prompt.endswith(" "):
        return prompt + token
    return prompt + " " + token


def findMatchingBracket(stringAppended: str) -> int:
    """Index of the `]` closing the bracket opened *before* position 1; position 0 is never inspected."""
    depth = 0
    for pos in range(1, len(stringAppended)):
        ch = stringAppended[pos]
        if ch == "[":
            depth += 1
        elif ch == "]":
            if depth == 0:
                return pos
            depth -= 1
    return -1


def parse_prompt(meta_prompt):
    """'a [robot:.6,.3,.4,.55] and a [cat:.2,.3]' -> (plain prompt, [(sub-prompt, AnnotationType, payload)], custom_losses).

    `[tok:x,y,w,h]` is a box (normalised top-left + extent), `[tok:x,y]` a crosshair, `[CustomLoss:name args]` a
    keyword loss that must be the last item.  Quirks kept from the reference (utils/helpers.py:59-114): a trailing
    bare word with neither a space nor a bracket after it is dropped; the first ':' anywhere in the remaining text is
    taken as the annotation's colon; an unmatched '[' swallows a single character.
    """
    prompt = ""
    meta_info = []
    custom_losses = {}
    rest = meta_prompt
    while True:
        rest = rest.lstrip(" ")
        sp = rest.find(" ")
        br = rest.find("[")
        if sp < 0 and br < 0:
            return (prompt, meta_info, custom_losses)
        if br < 0:
            return (add_word(prompt, rest), meta_info, custom_losses)
        if sp >= 0 and sp < br:
            # plain word (kept with its trailing blank so the next add_word does not double the separator)
            prompt = add_word(prompt, rest[: sp + 1])
            rest = rest[sp:]
            continue
        close = findMatchingBracket(rest[1:]) + 1
        colon = rest.index(":")
        token = rest[br + 1: colon].strip(" ")
        fields = rest[colon + 1: close].strip(" ").split(",")
        keep_word = True
        if token == "CustomLoss":
            keep_word = False
            name_and_args = rest[colon + 1:]
            cut = name_and_args.index(" ")
            name, args = name_and_args[:cut], name_and_args[cut + 1: -1]
            loss_obj = state.config.registered_loss_functions[name]
            custom_losses[name] = (loss_obj, args)
            for sub in loss_obj.subprompts_of_interest(args):
                meta_info.append((sub, AnnotationType.KEYWORD, None))
        elif len(fields) == 2:
            meta_info.append((token, AnnotationType.COOR, (float(fields[0]), float(fields[1]))))
        elif len(fields) == 4:
            x, y, w, h = (float(f) for f in fields)
            meta_info.append((token, AnnotationType.BOX, Rect(x, y, w, h, 1)))
        if keep_word:
            prompt = add_word(prompt, token)
        rest = rest[close + 1:]


def get_inner_folder_name():
    return get_meta_prompt_clean()


def get_meta_prompt_clean():
    cleaned = state.config.meta_prompt
    for ch in "[]:.":
        cleaned = cleaned.replace(ch, "_")
    return cleaned[0:5] if state.config.interactive else cleaned


# ------------------------------------------------------------------------------------------ box geometry (host)
sample_center = True
shrink_box = True


def get_corresponding_weight(x):
    """Strict-mode pixel weight versus normalised distance from the box centre (hard drop-off near the edge)."""
    return np.interp(x, [0, .333, .666, 1.0], [3, 2.5, 1, .2])


def inside_box(cur_x, cur_y, rect):
    """Pixel-centre test against the box shrunk by `shrink_factor` on every side, bounds inclusive, float64."""
    if sample_center:
        cur_x += 0.5
        cur_y += 0.5
    shrink = state.curHyperParams["shrink_factor"]
    off_x = shrink * rect.width
    off_y = shrink * rect.height
    in_x = cur_x >= (rect.x + off_x) and cur_x <= (rect.x + rect.width - off_x)
    in_y = cur_y >= (rect.y + off_y) and cur_y <= (rect.y + rect.height - off_y)
    return bool(in_x and in_y)


def distance_from_center(cur_x, cur_y, rect, normalized):
    if sample_center:
        cur_x += 0.5
        cur_y += 0.5
    cx, cy = rect.center()
    if normalized:
        return math.sqrt(math.pow(2 * (cx - cur_x) / rect.width, 2) + math.pow(2 * (cy - cur_y) / rect.height, 2)) \
            / math.sqrt(2)
    return math.sqrt(math.pow(cx - cur_x, 2) + math.pow(cy - cur_y, 2))


def distance_from_bounding_box(cur_x, cur_y, rect, normalized):
    if sample_center:
        cur_x += 0.5
        cur_y += 0.5
    if normalized:
        raise NotImplementedError()
    dx = abs(cur_x - rect.x) if cur_x < rect.x else (abs(cur_x - rect.right()) if cur_x > rect.right() else 0)
    dy = abs(cur_y - rect.y) if cur_y < rect.y else (abs(cur_y - rect.bottom()) if cur_y > rect.bottom() else 0)
    return dx + dy


def get_corresponding_weight_distance_from(dist):
    return 1.0


def box_mask_host(rect, res):
    """(res, res) uint8 mask of `inside_box` for a rect already scaled to `res` -- host reference for the device
    rasteriser and the source of `n_inside` (a box with no inside pixel raises like the reference does)."""
    m = np.zeros((res, res), dtype=np.uint8)
    for ii in range(res):
        for jj in range(res):
            if inside_box(jj, ii, rect):
                m[ii, jj] = 1
    return m


def strict_weights_host(rect, res):
    """Per-pixel strict-mode weights, normalised separately inside / outside the box (reference
    utils/helpers.py:216-246).  Built once per prompt on the host in float32 in the reference's accumulation order
    and uploaded; the per-step loss kernels only read it."""
    mask = box_mask_host(rect, res)
    w = np.ones((res, res), dtype=np.float32)
    for ii in range(res):
        for jj in range(res):
            if mask[ii, jj]:
                w[ii, jj] = np.float32(get_corresponding_weight(distance_from_center(jj, ii, rect, True)))
            else:
                w[ii, jj] = np.float32(get_corresponding_weight_distance_from(
                    distance_from_bounding_box(jj, ii, rect, False)))
    s_in = np.float32(0)
    s_out = np.float32(0)
    for ii in range(res):
        for jj in range(res):
            if mask[ii, jj]:
                s_in = np.float32(s_in + w[ii, jj])
            else:
                s_out = np.float32(s_out + w[ii, jj])
    out = np.where(mask.astype(bool), w / (s_in if s_in != 0 else np.float32(1)),
                   w / (s_out if s_out != 0 else np.float32(1))).astype(np.float32)
    return out


def calculate_bounding_box_losses(r, imageSoftmax):
    """(loss_inside, loss_outside) for a normalised map `imageSoftmax` (res, res) and a rect already scaled to res.

    Same contract as reference utils/helpers.py:215-277 (which hard-codes res=16 and runs ~1500 scalar tensor ops);
    here one fused CUDA launch through the C ABI, differentiable w.r.t. `imageSoftmax`.
    """
    from . import ops
    return ops.box_losses(imageSoftmax, r, shrink=state.curHyperParams["shrink_factor"],
                          strict=bool(state.curHyperParams["strict"]))


def dictToString(dict1):
    if type(dict1) is dict:
        return "".join("_" + str(k) + "_" + dictToString(v) for k, v in dict1.items() if k != "meta_prompt")
    return str(dict1)


# ---------------------------------------------------------------------------------------------------------- log
lines = []


def log(text, also_print=False):
    lines.append(text + os.linesep)
    if also_print and state.verbose:
        print(text)


def log_clear():
    global lines
    lines = []


def log_save(filename):
    with open(filename, "w") as fp:
        fp.writelines(lines)
    log_clear()
