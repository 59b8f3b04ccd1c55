"""The neural sub-modules the hot path CALLS but does not re-implement.

north_star: "The 3D-CNN cost regulariser and the per-sample MLPs stay on cuDNN/cuBLAS, timed
separately".  These are plain torch.nn definitions whose ONLY contract is the parameter naming of
the reference, so that a reference checkpoint (`torch.load(...)['net']`, reference
lib/utils/net_utils.py:415-447) loads unchanged:

  feature_net.*   2-D FPN         reference lib/networks/enerf/feature_net.py:4-36
  cost_reg_{i}.*  3-D U-Nets      reference lib/networks/enerf/cost_reg_net.py:4-86
  nerf_{i}.*      per-sample MLP  reference lib/networks/enerf/nerf.py:6-89

Architectural facts restated from those files: conv->BN->ReLU blocks named `.conv/.bn`; FPN with
1x1 laterals + bilinear(align_corners) x2 top-down adds and 3x3 smoothing; U-Net with stride-2
encoders, ConvTranspose3d(+BN) decoders added to the skip, two bias-free 3x3x3 heads; the MLP
aggregates V source views (mean/variance pooling, softmax view weights) and predicts sigma plus
softmax blending weights over the V fetched source colours.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _kaiming_linear(m):
    if isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight.data)
        if m.bias is not None:
            nn.init.zeros_(m.bias.data)


class _CBR(nn.Module):
    """conv -> batch-norm -> ReLU with the reference's attribute names (.conv, .bn)."""

    def __init__(self, conv_cls, bn_cls, cin, cout, k=3, stride=1, pad=1):
        super().__init__()
        self.conv = conv_cls(cin, cout, k, stride=stride, padding=pad, bias=False)
        self.bn = bn_cls(cout)

    def forward(self, x):
        if isinstance(self.bn, nn.Identity) and x.is_cuda and self.conv.bias is not None and not torch.is_grad_enabled():
            # BN already folded (inference_plan.py): one fused cuDNN call instead of conv + bias-add + ReLU kernels
            # (identical results; ~2x faster on the channels-last layers that stay on cuDNN, tools/cudnn_fused_bias_relu.py)
            c = self.conv
            return torch.cudnn_convolution_relu(x, c.weight, c.bias, c.stride, c.padding, c.dilation, c.groups)
        return F.relu(self.bn(self.conv(x)), inplace=True)


def _cbr2(cin, cout, k=3, stride=1, pad=1):
    return _CBR(nn.Conv2d, nn.BatchNorm2d, cin, cout, k, stride, pad)


def _cbr3(cin, cout, stride=1):
    return _CBR(nn.Conv3d, nn.BatchNorm3d, cin, cout, 3, stride, 1)


def _up3(cin, cout):
    return nn.Sequential(nn.ConvTranspose3d(cin, cout, 3, padding=1, output_padding=1, stride=2, bias=False),
                         nn.BatchNorm3d(cout))


class FeatureNet(nn.Module):
    """(N,3,H,W) -> level features (32ch @ H/4, 16ch @ H/2, 8ch @ H)."""

    def __init__(self):
        super().__init__()
        self.conv0 = nn.Sequential(_cbr2(3, 8), _cbr2(8, 8))
        self.conv1 = nn.Sequential(_cbr2(8, 16, 5, 2, 2), _cbr2(16, 16))
        self.conv2 = nn.Sequential(_cbr2(16, 32, 5, 2, 2), _cbr2(32, 32))
        self.toplayer = nn.Conv2d(32, 32, 1)
        self.lat1 = nn.Conv2d(16, 32, 1)
        self.lat0 = nn.Conv2d(8, 32, 1)
        self.smooth1 = nn.Conv2d(32, 16, 3, padding=1)
        self.smooth0 = nn.Conv2d(32, 8, 3, padding=1)

    @staticmethod
    def _up_add(x, skip):
        return F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True) + skip

    def forward(self, x):
        c0 = self.conv0(x)
        c1 = self.conv1(c0)
        c2 = self.conv2(c1)
        quarter = self.toplayer(c2)
        half = self._up_add(quarter, self.lat1(c1))
        full = self._up_add(half, self.lat0(c0))
        return quarter, self.smooth1(half), self.smooth0(full)


class CostRegNet(nn.Module):
    """3-level 3-D U-Net: (B,C,D,h,w) -> (feat (B,8,D,h,w), depth logits (B,D,h,w))."""
    depth_levels = 3

    def __init__(self, in_channels):
        super().__init__()
        self.conv0 = _cbr3(in_channels, 8)
        self.conv1 = _cbr3(8, 16, 2)
        self.conv2 = _cbr3(16, 16)
        self.conv3 = _cbr3(16, 32, 2)
        self.conv4 = _cbr3(32, 32)
        if self.depth_levels == 3:
            self.conv5 = _cbr3(32, 64, 2)
            self.conv6 = _cbr3(64, 64)
            self.conv7 = _up3(64, 32)
        self.conv9 = _up3(32, 16)
        self.conv11 = _up3(16, 8)
        self.depth_conv = nn.Sequential(nn.Conv3d(8, 1, 3, padding=1, bias=False))
        self.feat_conv = nn.Sequential(nn.Conv3d(8, 8, 3, padding=1, bias=False))

    def forward(self, x):
        s0 = self.conv0(x)
        s1 = self.conv2(self.conv1(s0))
        s2 = self.conv4(self.conv3(s1))
        y = s2
        if self.depth_levels == 3:
            y = s2 + self.conv7(self.conv6(self.conv5(s2)))
        y = s1 + self.conv9(y)
        y = s0 + self.conv11(y)
        return self.feat_conv(y), self.depth_conv(y).squeeze(1)


class MinCostRegNet(CostRegNet):
    """2-level variant used for cascade level 0 (no conv5/6/7)."""
    depth_levels = 2


class Agg(nn.Module):
    """Pools the V per-view feature vectors of a sample into 16 channels."""

    def __init__(self, feat_ch, viewdir_agg=True):
        super().__init__()
        self.feat_ch = feat_ch
        self.viewdir_agg = viewdir_agg
        if viewdir_agg:
            self.view_fc = nn.Sequential(nn.Linear(4, feat_ch), nn.ReLU())
            self.view_fc.apply(_kaiming_linear)
        self.global_fc = nn.Sequential(nn.Linear(feat_ch * 3, 32), nn.ReLU())
        self.agg_w_fc = nn.Sequential(nn.Linear(32, 1), nn.ReLU())
        self.fc = nn.Sequential(nn.Linear(32, 16), nn.ReLU())
        for m in (self.global_fc, self.agg_w_fc, self.fc):
            m.apply(_kaiming_linear)

    def forward(self, f):                                   # f: (B,P,V,feat_ch+4)
        B, V = len(f), f.shape[-2]
        x = f[..., :-4]
        if self.viewdir_agg:
            x = x + self.view_fc(f[..., -4:])
        var = torch.var(x, dim=-2).view(B, -1, 1, self.feat_ch).repeat(1, 1, V, 1)
        avg = torch.mean(x, dim=-2).view(B, -1, 1, self.feat_ch).repeat(1, 1, V, 1)
        g = self.global_fc(torch.cat([x, var, avg], dim=-1))
        w = F.softmax(self.agg_w_fc(g), dim=-2)
        return self.fc((g * w).sum(dim=-2))


class NeRF(nn.Module):
    """(vox_feat (B,P,8), img_feat_rgb_dir (B,P,V,feat_ch+4)) -> (B,P,4) = [rgb, sigma]."""

    def __init__(self, hid_n=64, feat_ch=16 + 3, viewdir_agg=True):
        super().__init__()
        self.hid_n = hid_n
        self.agg = Agg(feat_ch, viewdir_agg)
        self.lr0 = nn.Sequential(nn.Linear(8 + 16, hid_n), nn.ReLU())
        self.lrs = nn.ModuleList([])
        self.sigma = nn.Sequential(nn.Linear(hid_n, 1), nn.Softplus())
        self.color = nn.Sequential(nn.Linear(64 + 24 + feat_ch + 4, hid_n), nn.ReLU(),
                                   nn.Linear(hid_n, 1), nn.ReLU())
        for m in (self.lr0, self.sigma, self.color):
            m.apply(_kaiming_linear)

    def forward(self, vox_feat, img_feat_rgb_dir):
        B, V = img_feat_rgb_dir.shape[0], img_feat_rgb_dir.shape[2]
        pooled = self.agg(img_feat_rgb_dir)
        base = torch.cat((vox_feat, pooled), dim=-1)
        hid = self.lr0(base)
        sigma = self.sigma(hid)
        per_view = torch.cat((hid, base), dim=-1).view(B, -1, 1, self.hid_n + base.shape[-1]).repeat(1, 1, V, 1)
        per_view = torch.cat((per_view, img_feat_rgb_dir), dim=-1)
        w = F.softmax(self.color(per_view), dim=-2)
        rgb = torch.sum(img_feat_rgb_dir[..., -7:-4] * w, dim=-2)
        return torch.cat([rgb, sigma], dim=-1)


class EnerfModules(nn.Module):
    """Container with the reference's attribute names (reference lib/networks/enerf/network.py:11-22)."""

    def __init__(self, rc):
        super().__init__()
        self.feature_net = FeatureNet()
        for i in range(rc.num):
            ch = int(32 * (2 ** (-i)))
            setattr(self, f'cost_reg_{i}', MinCostRegNet(ch) if i == 0 else CostRegNet(ch))
            setattr(self, f'nerf_{i}', NeRF(feat_ch=rc.nerf_model_feat_ch[i] + 3, viewdir_agg=rc.viewdir_agg))
