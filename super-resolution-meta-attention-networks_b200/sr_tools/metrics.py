"""reference: Code/sr_tools/metrics.py:6-17."""
import numpy as np


def psnr(img1, img2, max_value=255.0):
    mse = np.mean((np.array(img1, dtype=np.float32) - np.array(img2, dtype=np.float32)) ** 2)
    if mse == 0:
        return 100
    return 20 * np.log10(max_value / (np.sqrt(mse)))
