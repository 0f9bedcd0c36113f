"""Input preparation: cv::resize's default INTER_LINEAR, restated in oracle/resize_oracle.py and
implemented on the GPU (rsf_resize_device).  The oracle is pinned against the cv2 module of this image (IPP switched
off = OpenCV's portable path); the GPU kernel against the oracle and against cv2 directly."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import resize_oracle  # noqa: E402

cv2 = pytest.importorskip("cv2")

# (source H, W) -> (result H, W): the reference's call sites
CASES = [((968, 1296), (480, 640)),      # scannet_sens_reader.cc:78
         ((720, 1280), (360, 640)),      # zed_native_l515.yaml: "L515 image downsampled by half" (exact 2x: area path)
         ((480, 640), (768, 1024)),      # up-sampling a segmentation map back to the image size (inference.cc:28)
         ((97, 131), (41, 57)),          # ragged
         ((5, 7), (11, 3))]


def _images(shape, seed):
    rng = np.random.default_rng(seed)
    H, W = shape
    yy, xx = np.mgrid[0:H, 0:W]
    smooth = (np.sin(xx / 17.0) + np.cos(yy / 11.0)) * 0.25 + 0.5
    rgb = np.clip(smooth[..., None] * 255 + rng.integers(-40, 40, size=(H, W, 3)), 0, 255).astype(np.uint8)
    depth = np.clip(smooth * 6000 + rng.integers(-300, 300, size=(H, W)), 0, 65535).astype(np.uint16)
    depth[rng.random((H, W)) < 0.05] = 0                                   # invalid pixels
    prob = np.clip(smooth + rng.normal(0, 0.05, size=(H, W)), 0, 1).astype(np.float32)
    return rgb, depth, prob


def _cv(img, size):
    cv2.ipp.setUseIPP(False)
    return cv2.resize(img, (size[1], size[0]))


@pytest.mark.parametrize("src,dst", CASES)
def test_oracle_matches_opencv(src, dst):
    rgb, depth, prob = _images(src, 1)
    assert np.array_equal(resize_oracle.resize_linear(rgb, dst), _cv(rgb, dst))
    assert np.array_equal(resize_oracle.resize_linear(rgb[..., 0].copy(), dst), _cv(rgb[..., 0].copy(), dst))
    a, b = resize_oracle.resize_linear(depth, dst), _cv(depth, dst)
    assert (a != b).mean() <= 1e-4 and np.abs(a.astype(int) - b.astype(int)).max() <= 1      # float blend: fused vs unfused ties
    assert np.allclose(resize_oracle.resize_linear(prob, dst), _cv(prob, dst), rtol=0, atol=2e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("src,dst", CASES)
def test_gpu_resize_matches_oracle_and_opencv(rsf, src, dst):
    import torch
    g = rsf.TSDFGrid(0.01, 0.06, pool_blocks=1024, table_slots_log2=12, stream=torch.cuda.current_stream().cuda_stream)
    rgb, depth, prob = _images(src, 2)
    for img in (rgb, rgb[..., 0].copy(), prob):
        got = g.Resize(torch.from_numpy(img).cuda(), dst).cpu().numpy()
        assert np.array_equal(got, resize_oracle.resize_linear(img, dst))                     # bit-exact vs the oracle
    got = g.Resize(torch.from_numpy(depth.view(np.int16)).cuda(), dst).cpu().numpy().view(np.uint16)
    assert np.array_equal(got, resize_oracle.resize_linear(depth, dst))
    assert np.array_equal(g.Resize(torch.from_numpy(rgb).cuda(), dst).cpu().numpy(), _cv(rgb, dst))   # and vs OpenCV itself
    d = _cv(depth, dst)
    assert (got != d).mean() <= 1e-4 and np.abs(got.astype(int) - d.astype(int)).max() <= 1


@pytest.mark.gpu
def test_resized_frame_integrates_like_the_cpu_prepared_one(rsf, synth):
    """feed_rgbd_frame (disinfect_slam.cc:33-42): resize rgb + raw depth to the TSDF size, convert, integrate -- with the
    resize on the GPU the map is identical to the one built from cv2-resized host images."""
    import torch
    seq = synth.config4(seed=11, n_frames=40, hires=True)          # 1024x768 sensor images
    small = (384, 512)
    Ks = synth.scaled(seq.K, 0.5, 0.5)
    st = torch.cuda.current_stream().cuda_stream
    ga = rsf.TSDFGrid(seq.voxel_size, seq.truncation, stream=st)
    gb = rsf.TSDFGrid(seq.voxel_size, seq.truncation, stream=st)
    for i in range(4):
        rgb, depth, ht, lt = (t for t in synth.render_torch(seq.scene, seq.K, seq.H, seq.W, seq.poses[i], "cuda", depth_factor=seq.depth_factor))
        raw = torch.round(depth.double() * seq.depth_factor).clamp(0, 65535).to(torch.int32)
        raw16 = raw.to(torch.int16) if int(raw.max()) < 32768 else None
        assert raw16 is not None
        # GPU preparation (sizes are not an exact 2x of `small`? they are: 768 / 384 = 2 -> the area path)
        r_small, d_small = ga.Resize(rgb.contiguous(), small), ga.Resize(raw16.contiguous(), small)
        ga.IntegrateU16(r_small, d_small, None, None, seq.depth_factor, seq.max_depth, Ks, seq.poses[i])
        # CPU preparation with OpenCV, as the reference does
        cv2.ipp.setUseIPP(False)
        r_cpu = cv2.resize(rgb.cpu().numpy(), (small[1], small[0]))
        d_cpu = cv2.resize(raw.cpu().numpy().astype(np.uint16), (small[1], small[0]))
        gb.IntegrateU16(torch.from_numpy(r_cpu).cuda(), torch.from_numpy(d_cpu.view(np.int16)).cuda(), None, None, seq.depth_factor,
                        seq.max_depth, Ks, seq.poses[i])
    ga.synchronize()
    gb.synchronize()
    a, b = ga.ExportBlocks(), gb.ExportBlocks()
    assert len(a[0]) > 500
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=x.dtype.kind == "f")     # (no segmentation: ht = 0 -> probabilities 0 or NaN)
