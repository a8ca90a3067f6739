"""`mlp()` builder with the reference's parameter naming (crowd_nav/policy/helpers.py:5-13).

The Sequential's state_dict keys ("0.weight", "0.bias", "2.weight", ...) are what the reference's
checkpoints contain, so .pth files interchange both ways.
"""
import torch.nn as nn


def mlp(input_dim, mlp_dims, last_relu=False):
    dims = [input_dim] + list(mlp_dims)
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if last_relu or i != len(dims) - 2:
            layers.append(nn.ReLU())
    return nn.Sequential(*layers)
