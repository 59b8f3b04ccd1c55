"""Call-side preparation of the KEPT cuDNN modules for inference (no change to their math).

north_star keeps the 2-D FPN and the 3-D cost regularisers on cuDNN.  Measured on B200
(tools/conv_experiments.py, profiles/round1_conv_variants.md) the stock eval() modules spend a
third of their time in stand-alone batch-norm kernels and NCHW<->NHWC transposes that cuDNN
inserts around its tensor-core kernels.  Two exact-up-to-rounding rewrites remove both:

  * batch-norm folding: eval-mode BN is an affine map per channel, folded into the preceding
    (transposed) convolution's weight and bias (relative difference ~1e-6 in fp32);
  * channels-last memory format for weights and activations (cuDNN's native tensor-core layout);
    our kernels read and write strided tensors, so no transpose is ever materialised.

The folded copies are derived objects: the registered parameters (and therefore state_dict /
checkpoint names) are untouched, and the copies are rebuilt whenever a parameter changes.
"""
import copy

import torch
import torch.nn as nn

from .modules import _CBR


def _fold_pair(conv, bn, transposed=False):
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    shape = [1] * conv.weight.dim()
    shape[1 if transposed else 0] = -1
    w = conv.weight * scale.view(shape)
    b = bn.bias - bn.running_mean * scale
    if conv.bias is not None:
        b = b + conv.bias * scale
    conv.weight = nn.Parameter(w, requires_grad=False)
    conv.bias = nn.Parameter(b, requires_grad=False)


def folded_copy(module, memory_format=None):
    """Deep copy of a FeatureNet / (Min)CostRegNet with every conv+BN pair folded."""
    m = copy.deepcopy(module).eval()
    with torch.no_grad():
        for mod in list(m.modules()):
            if isinstance(mod, _CBR) and isinstance(mod.bn, (nn.BatchNorm2d, nn.BatchNorm3d)):
                _fold_pair(mod.conv, mod.bn)
                mod.bn = nn.Identity()
            elif (isinstance(mod, nn.Sequential) and len(mod) == 2 and isinstance(mod[0], nn.ConvTranspose3d)
                  and isinstance(mod[1], nn.BatchNorm3d)):
                _fold_pair(mod[0], mod[1], transposed=True)
                mod[1] = nn.Identity()
    if memory_format is not None:
        m = m.to(memory_format=memory_format)
    for p in m.parameters():
        p.requires_grad_(False)
    return m


class PlanCache:
    """name -> folded copy, invalidated by parameter/buffer version counters."""

    def __init__(self):
        self._c = {}

    @staticmethod
    def _key(module):
        return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))

    def get(self, name, module, memory_format):
        key = (self._key(module), memory_format)
        hit = self._c.get(name)
        if hit is None or hit[0] != key:
            self._c[name] = (key, folded_copy(module, memory_format))
        return self._c[name][1]
