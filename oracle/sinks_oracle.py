"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the per-pixel parts of the reference's output sinks.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package.

Parity pin: the reference's Evaluator / Visualizer import lpips, skimage, cv2 and imageio, none of which exist in this
image, so they cannot be run here; `frame_psnr` restates skimage.metrics.peak_signal_noise_ratio (scikit-image >= 0.19,
the reference does not pin a version): err = mean((true.astype(f64) - test.astype(f64))**2); 10*log10(data_range**2/err),
called by the reference as psnr(gt_rgb[b][mask], pred_rgb[b][mask], data_range=1.) — **parity unpinned** against the
library itself; the arithmetic is checked against a float64 numpy evaluation in tests/test_host_logic.py.
"""
import numpy as np


def frame_psnr(pred, gt, mask=None, eval_center=False):
    """reference lib/evaluators/enerf.py:45-71.  pred, gt (h,w,3) float32; mask (h,w) (pixel kept when >= 1)."""
    h, w = pred.shape[:2]
    m = np.ones((h, w), dtype=bool) if mask is None else (np.asarray(mask) >= 1)
    if eval_center:
        ch, cw = int(h * 0.1), int(w * 0.1)
        pred, gt, m = pred[ch:-ch, cw:-cw], gt[ch:-ch, cw:-cw], m[ch:-ch, cw:-cw]
    a, b = gt[m].astype(np.float64), pred[m].astype(np.float64)
    err = np.mean((a - b) ** 2)
    return 10 * np.log10(1.0 / err)


def frame_to_u8(rgb, depth):
    """reference lib/visualizers/enerf.py:27-37: (pred_rgb * 255).astype(np.uint8) and
    ((depth - depth.min()) / (depth.max() - depth.min()) * 255).astype(np.uint8), float32 arithmetic."""
    rgb = np.asarray(rgb, dtype=np.float32)
    depth = np.asarray(depth, dtype=np.float32)
    return (rgb * 255).astype(np.uint8), ((depth - depth.min()) / (depth.max() - depth.min()) * 255).astype(np.uint8)
