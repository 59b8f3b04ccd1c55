"""`kornia.utils.create_meshgrid` restated from kornia's published behaviour (kornia 0.7.x):
returns a (1, H, W, 2) grid whose [..., 0] is x and [..., 1] is y, pixel units when
`normalized_coordinates=False`, else mapped to [-1, 1]."""
import torch


def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
    gx, gy = torch.meshgrid(xs, ys, indexing="ij")  # (W, H) each
    return torch.stack([gx, gy], dim=-1).permute(1, 0, 2).unsqueeze(0)
