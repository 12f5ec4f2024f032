"""numpy restatement of the device SSIM (csrc/ssim.cu): scipy.ndimage.gaussian_filter(sigma=1.5) on a (3,H,W) float32 array = a
13-tap Gaussian along the channel, H and W axes in that order, 'reflect' boundary, float64 accumulation, float32 result per axis
pass (inference/test_deblur_small.py:25-49).  Test infrastructure: pins the algorithm the kernels implement against scipy."""
import numpy as np


def reflect(i, n):
    while i < 0 or i >= n:
        if i < 0:
            i = -i - 1
        if i >= n:
            i = 2 * n - 1 - i
    return i


def ssim_model(out_img, gt):
    """out_img: float32 HWC in [0,255] (= clamp(out,0,1)*255), gt: uint8 HWC."""
    w = np.exp(-0.5 / 2.25 * np.arange(-6, 7) ** 2)
    w /= w.sum()
    a = (np.asarray(out_img, np.float32) / np.float32(255)).transpose(2, 0, 1)
    b = (np.asarray(gt, np.float32) / np.float32(255)).transpose(2, 0, 1)
    M = np.zeros((3, 3))
    for co in range(3):
        for k in range(-6, 7):
            M[co, reflect(co + k, 3)] += w[k + 6]

    def filt(q):
        q = np.einsum("oc,chw->ohw", M, q.astype(np.float64)).astype(np.float32)
        H, W = q.shape[1:]
        iy = np.array([[reflect(y + k, H) for k in range(-6, 7)] for y in range(H)])
        q = np.einsum("k,cykw->cyw", w, q.astype(np.float64)[:, iy, :]).astype(np.float32)
        ix = np.array([[reflect(x + k, W) for k in range(-6, 7)] for x in range(W)])
        return np.einsum("k,cyxk->cyx", w, q.astype(np.float64)[:, :, ix]).astype(np.float32)

    mu1, mu2, aa, bb, ab = [filt(q) for q in (a, b, a * a, b * b, a * b)]
    C1, C2 = np.float32(0.01 ** 2), np.float32(0.03 ** 2)
    m11, m22, m12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    mp = ((2 * m12 + C1) * (2 * (ab - m12) + C2)) / ((m11 + m22 + C1) * ((aa - m11) + (bb - m22) + C2))
    return float(mp.astype(np.float64).mean())
