"""Module-level run state -- mirror of the reference's `utils/shared_state.py` (names and defaults kept because the
hot path is parameterised through these globals in the reference: `pipeline_guided_attention.py:228, 277-279,
405-430`, `utils/helpers.py:168-169, 250`).  The CUDA path does not read globals from device code: the host wrappers
snapshot the values into an explicit parameter struct for every launch.
"""
config = None
cur_seed = None
cur_time_step_iter = None
always_save_iter = [24, 25, 26]
sub_iteration = 0

sigmas = None      # DDIM sigmas, sqrt((1 - a_bar) / a_bar)
timesteps = None   # inference timesteps, e.g. 981, 961, ..., 1

# experimental deep-feature branch of the reference (pipeline_guided_attention.py:693-706): not supported, kept False
optimizeDeepLatent = False
use_loss_total = True
deepLatentRequiresGrad = True
injectDeepFeatures = False
deepFeatures = None

curHyperParams = None

# extension (not in the reference): chatty per-iteration prints are opt-in; the log buffer is always filled
verbose = False

# shipped defaults (reference utils/shared_state.py:21); `thresholds` here overrides RunConfig.thresholds (run.py:75-79)
hyperParameterOverrides = {"strict": False, "inside_loss_scale": .2, "outside_loss_scale": .2, "shrink_factor": .15,
                           "thresholds": {0: 1.}, "use_optimizer": False, "recurse_until": 14, "recurse_steps": 3}
hyperParameterIterations = [{}]


def get_sigma():
    return sigmas[timesteps[cur_time_step_iter]]


def get_hyperparam_states():
    states = []
    for overrides in hyperParameterIterations:
        merged = dict(hyperParameterOverrides)
        merged.update(overrides)
        states.append(merged)
    return states


tags = ["cur_seed", "cur_time_step_iter", "optimizeDeepLatent"]


def to_str(t):
    return "{:02d}".format(t) if type(t) is int else str(t)


def get_name():
    return "".join(str(t) + "_" + to_str(globals()[t]) + "_" for t in tags)
