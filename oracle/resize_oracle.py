"""resize_oracle.py -- TEST INFRASTRUCTURE ONLY (see oracle/README or DESIGN.md 2): a numpy restatement of
cv::resize(src, dst, Size(w, h)) with the default INTER_LINEAR, the input preparation the reference's callers run on the
CPU before TSDFGrid::Integrate (disinfect_slam/disinfect_slam.cc:37-38, utils/offline_data_provider/scannet_sens_reader.cc:78,
segmentation/inference.cc:28).  OpenCV is a third-party dependency of the reference that is not vendored in its tree; the
algorithm restated here is OpenCV's published portable path (modules/imgproc/src/resize.cpp: resizeGeneric_ /
HResizeLinear / VResizeLinear / ResizeAreaFast).  Pinned by tests/test_resize.py against the cv2 module of this image with
IPP switched off (bit-exact for 8UC3 / 8UC1 / 16UC1; float within 1 ulp-scale tolerance).

Only tests/ may import this file; the product path is ra-slam_b200/csrc/rsf_kernels.cu: resize_kernel."""
import numpy as np


def _taps(d: int, s: int, zero_at_border: bool):
    scale = float(s) / float(d)                                   # double
    f = ((np.arange(d, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    i = np.floor(f).astype(np.int64)
    f = (f - i.astype(np.float32)).astype(np.float32)
    if zero_at_border:                                            # horizontal: resize.cpp, xofs / alpha setup
        lo, hi = i < 0, i >= s - 1
        f = np.where(lo | hi, np.float32(0), f)
        i = np.where(lo, 0, np.where(hi, s - 1, i))
        return i, np.minimum(i + 1, s - 1), f
    return np.clip(i, 0, s - 1), np.clip(i + 1, 0, s - 1), f      # vertical: row indices are clipped, weights kept


def resize_linear(src: np.ndarray, size) -> np.ndarray:
    """size = (H, W) of the result.  src: uint8 HxW[x3], uint16 HxW, float32 HxW."""
    dH, dW = size
    sH, sW = src.shape[:2]
    img = src.reshape(sH, sW, -1)
    if sW == 2 * dW and sH == 2 * dH:                             # cv::resize reroutes exact 2x decimation to INTER_AREA
        a, b, c, d = img[0::2, 0::2], img[0::2, 1::2], img[1::2, 0::2], img[1::2, 1::2]
        if src.dtype == np.float32:
            out = (((a + b) + c + d) * np.float32(0.25)).astype(np.float32)
        else:
            out = ((a.astype(np.int64) + b + c + d + 2) >> 2).astype(src.dtype)
        return out.reshape((dH, dW) + src.shape[2:])
    x0, x1, fx = _taps(dW, sW, True)
    y0, y1, fy = _taps(dH, sH, False)
    if src.dtype == np.uint8:
        a0 = np.rint((np.float32(1) - fx) * np.float32(2048)).astype(np.int64)[None, :, None]
        a1 = np.rint(fx * np.float32(2048)).astype(np.int64)[None, :, None]
        b0 = np.rint((np.float32(1) - fy) * np.float32(2048)).astype(np.int64)[:, None, None]
        b1 = np.rint(fy * np.float32(2048)).astype(np.int64)[:, None, None]
        s = img.astype(np.int64)
        h = s[:, x0] * a0 + s[:, x1] * a1                         # HResizeLinear: int32, scaled by 2^11
        v = (((b0 * (h[y0] >> 4)) >> 16) + ((b1 * (h[y1] >> 4)) >> 16) + 2) >> 2    # VResizeLinear<uchar>
        out = np.clip(v, 0, 255).astype(np.uint8)
    else:
        a0, a1 = (np.float32(1) - fx)[None, :, None], fx[None, :, None]
        b0, b1 = (np.float32(1) - fy)[:, None, None], fy[:, None, None]
        s = img.astype(np.float32)
        h = (s[:, x0] * a0).astype(np.float32) + (s[:, x1] * a1).astype(np.float32)
        v = (h[y0] * b0).astype(np.float32) + (h[y1] * b1).astype(np.float32)
        out = v.astype(np.float32) if src.dtype == np.float32 else np.clip(np.rint(v), 0, 65535).astype(np.uint16)
    return out.reshape((dH, dW) + src.shape[2:])
