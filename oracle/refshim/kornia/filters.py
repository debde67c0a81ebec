"""`kornia.filters.gaussian_blur2d` restated from kornia's public definition
(border_type='reflect', separable=True): normalised Gaussian taps, horizontal pass then
vertical pass, each a reflect-padded depthwise conv2d.  Used at
/root/reference/helper/stereo_core.py:385 and :432.  TEST INFRASTRUCTURE ONLY."""
import torch
import torch.nn.functional as F


def _taps(k: int, sigma: float, dtype, device):
    n = torch.arange(k, dtype=dtype, device=device) - (k // 2)
    if k % 2 == 0:
        n = n + 0.5
    g = torch.exp(-(n * n) / (2.0 * float(sigma) ** 2))
    return g / g.sum()


def gaussian_blur2d(x, kernel_size, sigma, border_type='reflect', separable=True):
    ky, kx = int(kernel_size[0]), int(kernel_size[1])
    sy, sx = float(sigma[0]), float(sigma[1])
    c = x.shape[1]
    gx = _taps(kx, sx, x.dtype, x.device).view(1, 1, 1, kx).expand(c, 1, 1, kx)
    gy = _taps(ky, sy, x.dtype, x.device).view(1, 1, ky, 1).expand(c, 1, ky, 1)
    x = F.conv2d(F.pad(x, (kx // 2, kx // 2, 0, 0), mode=border_type), gx, groups=c)
    x = F.conv2d(F.pad(x, (0, 0, ky // 2, ky // 2), mode=border_type), gy, groups=c)
    return x
