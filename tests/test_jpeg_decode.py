"""Device-side JPEG decode: rba_jpeg_decode (nvJPEG) against the host decoder the reference's dataset classes
use (PIL / libjpeg).  JPEG decoders may differ by rounding in the IDCT / colour conversion: the bar is +-2 grey levels per
sample on smooth content and a mean error well below one level."""
import io

import numpy as np
import pytest
import torch

PIL = pytest.importorskip("PIL.Image")

from rba_b200 import pipeline  # noqa: E402
from rba_b200._lib import RbaError  # noqa: E402


def _jpeg_bytes(h, w, seed, quality=95, grey=False, subsampling=0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 100 * np.sin(xx / (11.0 + seed)) * np.cos(yy / 17.0),
                    127 + 90 * np.cos(xx / 23.0 + seed) * np.sin(yy / (7.0 + seed)),
                    (xx * 255.0 / w + yy * 255.0 / h) / 2], -1)
    img = np.clip(img + rng.normal(0, 2, img.shape), 0, 255).astype(np.uint8)
    pil = PIL.fromarray(img[..., 0] if grey else img)
    buf = io.BytesIO()
    pil.save(buf, format="JPEG", quality=quality, subsampling=0 if grey else subsampling)
    return buf.getvalue()


def test_abi_exports_jpeg_entry_points():
    from rba_b200 import _lib
    lib = _lib.lib()
    for name in ("rba_jpeg_available", "rba_jpeg_info", "rba_jpeg_decode"):
        assert hasattr(lib, name)


@pytest.mark.gpu
def test_jpeg_decode_matches_host_decoder():
    if not pipeline.jpeg_available():
        pytest.skip("libnvjpeg is not installed on this box")
    h, w = 96, 160
    # 4:4:4 and grey streams: the decoders may differ by IDCT / colour-conversion rounding only; the 4:2:0 stream adds the
    # chroma up-sampling filter (libjpeg's "fancy" triangle filter vs nvJPEG's), a few levels at chroma edges
    streams = [_jpeg_bytes(h, w, s) for s in range(2)] + [_jpeg_bytes(h, w, 7, grey=True), _jpeg_bytes(h, w, 3, subsampling=2)]
    assert pipeline.jpeg_info(streams[0]) == (h, w, 3)
    out = pipeline.decode_jpeg_batch(streams)
    torch.cuda.synchronize()
    assert out.shape == (4, 3, h, w) and out.dtype == torch.uint8 and out.is_cuda
    for i, s in enumerate(streams):
        ref = np.asarray(PIL.open(io.BytesIO(s)).convert("RGB")).transpose(2, 0, 1).astype(np.int16)
        d = np.abs(out[i].cpu().numpy().astype(np.int16) - ref)
        print("stream", i, "max", int(d.max()), "mean", float(d.mean()))
        if i < 3:
            assert d.max() <= 3 and d.mean() < 0.6, (i, int(d.max()), float(d.mean()))
        else:
            assert d.max() <= 24 and d.mean() < 2.0, (i, int(d.max()), float(d.mean()))
    # the planes feed the engine unchanged: same layout as ToTensorV2 output (CHW, uint8)
    with pytest.raises(RbaError):
        pipeline.decode_jpeg_batch([streams[0], _jpeg_bytes(h + 8, w, 1)])     # mixed sizes are rejected
    with pytest.raises(RbaError):
        pipeline.jpeg_info(b"not a jpeg stream at all")
