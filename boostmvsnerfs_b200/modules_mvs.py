"""Kept neural sub-modules of the MVSNeRF backbone, cuDNN/cuBLAS only.

Parameter names follow the reference so its checkpoints load unchanged (467,732 parameters):
  feature.*     2-D CNN, conv -> in-place ABN blocks   reference lib/networks/mvsnerf/network.py:695-732
  cost_reg_2.*  3-D U-Net on the 41-channel volume     reference lib/networks/mvsnerf/network.py:735-779
  nerf.nerf.*   6x128 MLP ("Renderer_ours", v0)        reference lib/networks/mvsnerf/network.py:152-229,547-574
The reference's normalisation layer is the third-party `inplace_abn.InPlaceABN` (unpinned,
reference requirements.txt:16): batch-norm followed by leaky-ReLU(0.01); `ABN` below restates that
published behaviour with the same state-dict keys.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class ABN(nn.Module):
    """batch-norm + leaky-ReLU(slope); keys: weight, bias, running_mean, running_var."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, slope=0.01):
        super().__init__()
        self.eps, self.momentum, self.slope = eps, momentum, slope
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))

    def forward(self, x):
        x = F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, self.training,
                         self.momentum, self.eps)
        return F.leaky_relu(x, self.slope)


class _CA(nn.Module):
    """conv -> ABN with the reference's attribute names (.conv, .bn)."""

    def __init__(self, conv_cls, cin, cout, k=3, stride=1, pad=1):
        super().__init__()
        self.conv = conv_cls(cin, cout, k, stride=stride, padding=pad, bias=False)
        self.bn = ABN(cout)

    def forward(self, x):
        return self.bn(self.conv(x))


def _ca2(cin, cout, k=3, stride=1, pad=1):
    return _CA(nn.Conv2d, cin, cout, k, stride, pad)


def _ca3(cin, cout, stride=1):
    return _CA(nn.Conv3d, cin, cout, 3, stride, 1)


def _up3(cin, cout):
    return nn.Sequential(nn.ConvTranspose3d(cin, cout, 3, padding=1, output_padding=1, stride=2, bias=False), ABN(cout))


class MvsFeatureNet(nn.Module):
    """(B,V,3,H,W) -> (B,V,32,H/4,W/4)."""

    def __init__(self):
        super().__init__()
        self.conv0 = nn.Sequential(_ca2(3, 8), _ca2(8, 8))
        self.conv1 = nn.Sequential(_ca2(8, 16, 5, 2, 2), _ca2(16, 16), _ca2(16, 16))
        self.conv2 = nn.Sequential(_ca2(16, 32, 5, 2, 2), _ca2(32, 32), _ca2(32, 32))
        self.toplayer = nn.Conv2d(32, 32, 1)

    def forward(self, x):
        B, V, C, H, W = x.shape
        y = self.toplayer(self.conv2(self.conv1(self.conv0(x.view(B * V, C, H, W)))))
        return y.view(B, V, 32, H // 4, W // 4)


class MvsCostRegNet(nn.Module):
    """3-D U-Net: (B,41,D,h,w) -> (B,8,D,h,w)."""

    def __init__(self, in_channels):
        super().__init__()
        self.conv0 = _ca3(in_channels, 8)
        self.conv1 = _ca3(8, 16, 2)
        self.conv2 = _ca3(16, 16)
        self.conv3 = _ca3(16, 32, 2)
        self.conv4 = _ca3(32, 32)
        self.conv5 = _ca3(32, 64, 2)
        self.conv6 = _ca3(64, 64)
        self.conv7 = _up3(64, 32)
        self.conv9 = _up3(32, 16)
        self.conv11 = _up3(16, 8)

    def forward(self, x):
        s0 = self.conv0(x)
        s1 = self.conv2(self.conv1(s0))
        s2 = self.conv4(self.conv3(s1))
        y = s2 + self.conv7(self.conv6(self.conv5(s2)))
        y = s1 + self.conv9(y)
        return s0 + self.conv11(y)


def _kaiming(m):
    if isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight.data)
        if m.bias is not None:
            nn.init.zeros_(m.bias.data)


class RendererOurs(nn.Module):
    """x = [PE(ndc) (in_pts), per-sample features (in_feat), view dir (in_views)] -> [rgb(3), alpha(1)]."""

    def __init__(self, D=6, W=128, in_pts=63, in_views=3, in_feat=20, skips=(4,)):
        super().__init__()
        self.in_pts, self.in_views, self.in_feat, self.skips = in_pts, in_views, in_feat, tuple(skips)
        self.pts_linears = nn.ModuleList(
            [nn.Linear(in_pts, W)] + [nn.Linear(W + in_pts, W) if i in self.skips else nn.Linear(W, W) for i in range(D - 1)])
        self.pts_bias = nn.Linear(in_feat, W)
        self.views_linears = nn.ModuleList([nn.Linear(in_views + W, W // 2)])
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)
        self.rgb_linear = nn.Linear(W // 2, 3)
        for m in (self.pts_linears, self.views_linears, self.feature_linear, self.alpha_linear, self.rgb_linear):
            m.apply(_kaiming)

    def forward(self, x):
        pts, feats, views = torch.split(x, [self.in_pts, x.shape[-1] - self.in_pts - self.in_views, self.in_views], dim=-1)
        gate = self.pts_bias(feats)
        h = pts
        for i, lin in enumerate(self.pts_linears):
            h = F.relu(lin(h) * gate)
            if i in self.skips:
                h = torch.cat([pts, h], -1)
        alpha = torch.relu(self.alpha_linear(h))
        h = torch.cat([self.feature_linear(h), views], -1)
        h = F.relu(self.views_linears[0](h))
        return torch.cat([torch.sigmoid(self.rgb_linear(h)), alpha], -1)


class MvsNerfMlp(nn.Module):
    """wrapper that reproduces the reference's `nerf.nerf.*` key prefix."""

    def __init__(self):
        super().__init__()
        self.nerf = RendererOurs(D=6, W=128, in_pts=63, in_views=3, in_feat=20, skips=(4,))

    def forward(self, x):
        return self.nerf(x)


class MvsnerfModules(nn.Module):
    """Container with the reference's attribute names (reference lib/networks/mvsnerf/network.py:795-811)."""

    def __init__(self):
        super().__init__()
        self.feature = MvsFeatureNet()
        self.cost_reg_2 = MvsCostRegNet(32 + 9)
        self.nerf = MvsNerfMlp()
