"""CPU: host side of the input pipeline (rba_b200.PinnedBatcher): order, padding, layout conversion, error surfacing."""
import pytest
import torch

import rba_b200


class _DS:
    def __init__(self, n, H=8, W=12, hwc=False, bad=None):
        self.n, self.H, self.W, self.hwc, self.bad = n, H, W, hwc, bad

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if self.bad == i:
            raise ValueError("boom")
        g = torch.Generator().manual_seed(i)
        img = torch.randint(0, 256, (3, self.H, self.W), dtype=torch.uint8, generator=g)
        lab = torch.randint(0, 2, (self.H, self.W), dtype=torch.int64, generator=g)
        if self.hwc:
            return img.permute(1, 2, 0).numpy(), lab.numpy()      # what cv2 / albumentations hand over
        return img, lab


@pytest.mark.parametrize("hwc", [False, True])
def test_batcher_order_padding_and_layout(hwc):
    ds = _DS(7, hwc=hwc)
    ref = _DS(7)
    seen = 0
    b = rba_b200.PinnedBatcher(ds, batch=3, workers=4, pin=False)
    assert len(b) == 3
    for images, labels, n_valid in b:
        assert images.shape == (3, 3, 8, 12) and images.dtype == torch.uint8 and labels.shape == (3, 8, 12)
        for j in range(n_valid):
            im, lab = ref[seen + j]
            assert torch.equal(images[j], im) and torch.equal(labels[j].long(), lab)
        for j in range(n_valid, 3):                               # padded tail: repeated image, ignored labels
            assert torch.equal(images[j], images[0]) and bool((labels[j] == 255).all())
        seen += n_valid
    assert seen == 7


def test_batcher_surfaces_loader_errors_and_size_mismatch():
    with pytest.raises(ValueError):
        list(rba_b200.PinnedBatcher(_DS(5, bad=3), batch=2, pin=False))

    class Mixed(_DS):
        def __getitem__(self, i):
            self.H = 8 if i < 2 else 10
            return super().__getitem__(i)
    with pytest.raises(rba_b200.RbaError):
        list(rba_b200.PinnedBatcher(Mixed(4), batch=2, workers=1, pin=False))

    class F32(_DS):
        def __getitem__(self, i):
            im, lab = super().__getitem__(i)
            return im.float(), lab
    with pytest.raises(rba_b200.RbaError):
        list(rba_b200.PinnedBatcher(F32(2), batch=2, pin=False))
