"""mmcv.cnn.utils.flops_counter as tools/test_gaze360_gaze.py:16,55,104,123 uses it: the script wraps the model and
brackets every forward with start / stop calls but never reads a count.  The backend's FLOPs per clip are a static
property of the architecture, so the hooks are no-ops and `compute_average_flops_cost` reports
that constant."""
from __future__ import annotations

GFLOP_PER_FRAME_224 = 99.55 / 7        # 2 x MAC, T = 7 frames of 224 x 224


def add_flops_counting_methods(model):
    model.start_flops_count = lambda *a, **k: None
    model.stop_flops_count = lambda *a, **k: None
    model.reset_flops_count = lambda *a, **k: None
    model.compute_average_flops_cost = lambda: (GFLOP_PER_FRAME_224 * 1e9, 0)
    return model


def flops_to_string(flops, units='GFLOPs', precision=2):
    div = {'GFLOPs': 1e9, 'MFLOPs': 1e6, 'KFLOPs': 1e3}.get(units, 1.0)
    return f'{round(flops / div, precision)} {units}'


def params_to_string(num_params, units=None, precision=2):
    if units == 'M' or (units is None and num_params >= 1e6):
        return f'{round(num_params / 1e6, precision)} M'
    if units == 'K' or (units is None and num_params >= 1e3):
        return f'{round(num_params / 1e3, precision)} k'
    return str(num_params)
