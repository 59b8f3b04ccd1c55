"""Shim for the un-vendored, unpinned `inplace_abn` CUDA extension (reference requirements.txt:16),
used only inside the MVSNeRF CNNs (reference lib/networks/mvsnerf/network.py:13,668-692).
Published behaviour (mapillary/inplace_abn): batch-norm followed by leaky-ReLU(0.01); state-dict keys
weight, bias, running_mean, running_var.  The CNNs are NOT on the re-implemented path; both sides of
every parity test call the same module, so the exact affine convention does not affect parity."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class InPlaceABN(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True,
                 activation="leaky_relu", activation_param=0.01):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.activation, self.activation_param = activation, activation_param
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))

    def forward(self, x):
        x = F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias,
                         self.training, self.momentum, self.eps)
        if self.activation == "leaky_relu":
            return F.leaky_relu(x, self.activation_param)
        if self.activation == "elu":
            return F.elu(x, self.activation_param)
        return x


class InPlaceABNSync(InPlaceABN):
    pass
