"""GPU: device-side letterbox / pre-process against the oracle restatement of the reference's letterbox
(itself pinned bit-exactly to the reference + OpenCV through tests/golden/letterbox_cases.npz).  Integer work: bit exact."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import restate

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_letterbox_matches_reference_fixtures_bit_exact():
    from multiyolov5_b200.utils.datasets import letterbox
    g = np.load(os.path.join(GOLD, "letterbox_cases.npz"))
    meta = json.loads(bytes(g["meta_json"]).decode())
    for key, m in meta.items():
        fn, sn = key.rsplit("_", 1)
        out, ratio, dwdh = letterbox(g[f"in_{fn}"], **m["kw"])
        ref = g[f"out_{key}"]
        assert tuple(out.shape) == ref.shape, (key, out.shape, ref.shape)
        assert np.array_equal(out.cpu().numpy(), ref), (key, int((out.cpu().numpy() != ref).sum()))
        assert np.allclose(ratio, m["ratio"]) and np.allclose(dwdh, m["dwdh"]), key


@pytest.mark.parametrize("shape,size", [((1024, 2048), 1024), ((720, 1280), 640), ((375, 500), 640), ((1080, 810), 1024), ((64, 64), 64)])
def test_preprocess_full_size_bit_exact(shape, size):
    """full-size frames (incl. the Cityscapes 2048x1024 -> 1024x512 exact-2x case that OpenCV routes to its area path): the fused
    RGB / CHW / float output equals the oracle's uint8 result divided by 255 in the output precision"""
    from multiyolov5_b200.utils.datasets import letterbox, preprocess
    rs = np.random.RandomState(shape[0] + size)
    img0 = rs.randint(0, 256, shape + (3,), dtype=np.uint8)
    want = restate.preprocess_np(img0, size, stride=32)                      # uint8 (3,H,W) RGB
    lb, _, _ = letterbox(img0, size, stride=32)
    assert np.array_equal(lb.cpu().numpy(), restate.letterbox_np(img0, size, stride=32)[0])
    for half in (True, False):
        out, ratio, dwdh = preprocess(img0, size, stride=32, half=half)
        assert out.shape == (1,) + want.shape and out.dtype == (torch.float16 if half else torch.float32)
        ref = torch.from_numpy(want)[None].cuda()
        ref = (ref.half() if half else ref.float()) / 255.0                   # detect.py:135-137
        assert torch.equal(out, ref)


def test_preprocess_batch_of_frames():
    from multiyolov5_b200.utils.datasets import preprocess
    rs = np.random.RandomState(3)
    frames = rs.randint(0, 256, (3, 300, 400, 3), dtype=np.uint8)
    out, _, _ = preprocess(torch.from_numpy(frames).cuda(), 256, stride=32, half=False)
    for b in range(3):
        want = torch.from_numpy(restate.preprocess_np(frames[b], 256, stride=32)).cuda().float() / 255.0    # CUDA arithmetic, as detect.py
        assert torch.equal(out[b], want)
