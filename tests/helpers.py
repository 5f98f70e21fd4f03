"""Shared test helpers: load a golden case and regenerate its weights/inputs from seeds."""
import os

import numpy as np
import torch

from auto_avsr_b200.synthetic import encoder_input, encoder_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d_model, n_heads, linear_units, num_blocks, cnn_kernel = (int(v) for v in z["config"])
    lengths = [int(v) for v in z["lengths"]]
    masked = bool(int(z["masked"]))
    sd = encoder_state_dict(int(z["wseed"]), d_model, n_heads, linear_units, num_blocks, cnn_kernel)
    xs = encoder_input(lengths, d_model, int(z["xseed"]))
    return dict(z=z, cfg=dict(d_model=d_model, n_heads=n_heads, linear_units=linear_units,
                              num_blocks=num_blocks, cnn_kernel=cnn_kernel),
                lengths=lengths, masked=masked, sd=sd, xs=xs)


def err_stats(a, b):
    d = (a.double() - b.double()).abs()
    return d.max().item(), d.pow(2).mean().sqrt().item()
