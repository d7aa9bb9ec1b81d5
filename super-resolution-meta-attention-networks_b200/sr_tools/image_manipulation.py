"""RGB <-> YCbCr (ITU-R BT.601), the post-processing `ModelInterface.net_run_and_process` applies to the
network output (reference: Code/sr_tools/image_manipulation.py:56-157).  C,H,W images, numpy or torch."""
import numpy as np
import torch


def rgb_to_ycbcr(img, y_only=True, max_val=1, im_type='png'):
    r, g, b = img[0], img[1], img[2]
    if im_type == 'jpg':
        bias_c = 128. * (max_val / 255)
        y = 0.299 * r + 0.587 * g + 0.114 * b
        if y_only:
            return y, None, None
        cb = bias_c + (-0.168736 * r - 0.331264 * g + 0.5 * b)
        cr = bias_c + (0.5 * r - 0.418688 * g - 0.081312 * b)
    else:
        bias_y = 16. * (max_val / 255)
        bias_c = 128. * (max_val / 255)
        y = bias_y + (65.481 * r + 128.553 * g + 24.966 * b) / 255.
        if y_only:
            return y, None, None
        cb = bias_c + (-37.797 * r - 74.203 * g + 112.0 * b) / 255.
        cr = bias_c + (112.0 * r - 93.786 * g - 18.214 * b) / 255.
    return y, cb, cr


def ycbcr_to_rgb(img, max_val=1, im_type='png'):
    y, cb, cr = img[0], img[1], img[2]
    if im_type == 'jpg':
        bias = 128. * (max_val / 255)
        r = y + 1.402 * cr - 1.402 * bias
        g = y - 0.344136 * cb - 0.714136 * cr + (0.714136 + 0.344136) * bias
        b = y + 1.772 * cb - 1.772 * bias
    else:
        r = 298.082 * y / 256. + 408.583 * cr / 256. - 222.921 * (max_val / 255)
        g = 298.082 * y / 256. - 100.291 * cb / 256. - 208.120 * cr / 256. + 135.576 * (max_val / 255)
        b = 298.082 * y / 256. + 516.412 * cb / 256. - 276.836 * (max_val / 255)
    return r, g, b


def ycbcr_convert(img, y_only=True, max_val=1, im_type='png', input='rgb'):
    if isinstance(img, np.ndarray):
        stack, expand = (lambda c: np.array(c)), (lambda a: np.expand_dims(a, axis=0))
    elif isinstance(img, torch.Tensor):
        stack, expand = (lambda c: torch.stack(c, 0)), (lambda a: torch.unsqueeze(a, 0))
    else:
        raise Exception('Unknown Type', type(img))
    if len(img.shape) == 4:
        img = img.squeeze(0)
    if input == 'ycbcr':
        planes = ycbcr_to_rgb(img, max_val=max_val, im_type=im_type)
    else:
        planes = rgb_to_ycbcr(img, max_val=max_val, y_only=y_only, im_type=im_type)
    if y_only and input == 'rgb':
        return expand(planes[0])
    return stack(list(planes))
