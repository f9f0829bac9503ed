"""Declarative helpers for the `models/` mirror.  The mirror classes only HOLD parameters (names, shapes, registration
order and initialisation must equal the reference's so that state_dict / optimizer checkpoints interchange); these helpers
keep their constructors to a table of (attribute name, factory) pairs evaluated in order."""
import math

import torch
import torch.nn as nn


def declare(module, table):
    """Register sub-modules / parameters on `module` in table order (order = state_dict order = RNG draw order)."""
    for name, make in table:
        if make is not None:
            setattr(module, name, make())
    return module


def lin(n_in, n_out, bias=True, xavier_normal=False):
    m = nn.Linear(n_in, n_out, bias=bias)
    if xavier_normal:
        nn.init.xavier_normal_(m.weight)
    return m


def tanh_norm(width, p_drop=None):
    """nn.Sequential(Tanh, LayerNorm[, Dropout]): index 1 is the LayerNorm (keys '<name>.1.weight/bias')."""
    layers = [nn.Tanh(), nn.LayerNorm(width)]
    if p_drop is not None:
        layers.append(nn.Dropout(p_drop))
    return nn.Sequential(*layers)


def lin_tanh_norm(n_in, width):
    """nn.Sequential(Linear, Tanh, LayerNorm): keys '<name>.0.*' and '<name>.2.*'."""
    return nn.Sequential(nn.Linear(n_in, width), nn.Tanh(), nn.LayerNorm(width))


def xavier_uniform_param(rows, cols, gain):
    p = nn.Parameter(torch.empty(size=(rows, cols)))
    nn.init.xavier_uniform_(p, gain=nn.init.calculate_gain(gain))
    return p


def sinusoid_table(length, width):
    """(1, length, width) sin / cos position table (even columns sin, odd columns cos, base 10000)."""
    pos = torch.arange(0., length).unsqueeze(1)
    freq = torch.exp(torch.arange(0., width, 2) * -(math.log(10000.0) / width))
    table = torch.zeros(length, width)
    table[:, 0::2] = torch.sin(pos * freq)
    table[:, 1::2] = torch.cos(pos * freq)
    return table.unsqueeze(0)
