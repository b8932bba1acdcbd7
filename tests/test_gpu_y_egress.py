"""GPU: SURVEY §8(f) row 4 — the overlay kernel against the vectors the unmodified reference
functions produced (tests/golden/egress.npz) and against the oracle on seeded inputs: byte-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from mivos_b200 import _lib, egress, ops  # noqa: E402


@pytest.mark.parametrize("fade", [False, True])
@pytest.mark.parametrize("alpha", [0.5, 0.3])
def test_overlay_matches_reference_golden(golden, dev, alpha, fade):
    g = golden("egress.npz")
    img, mask = torch.from_numpy(g["image"]).to(dev), torch.from_numpy(g["mask"]).to(dev)
    got = ops.overlay_davis(img, mask, alpha, fade=fade)
    torch.cuda.synchronize()
    _lib.poll_kernel_error()
    want = g[("fade" if fade else "overlay") + f"_a{int(alpha * 10)}"]
    assert got.dtype == torch.uint8 and tuple(got.shape) == want.shape
    assert np.array_equal(got.cpu().numpy(), want)


def test_overlay_batch_480p_matches_oracle(dev):
    from oracle import egress_oracle as EO
    rng = np.random.default_rng(7)
    t, h, w = 3, 480, 854
    img = rng.integers(0, 256, size=(t, h, w, 3), dtype=np.uint8)
    mask = np.zeros((t, h, w), dtype=np.uint8)
    for i in range(t):
        for lab in range(1, 7):
            y0, x0 = rng.integers(0, h - 60), rng.integers(0, w - 90)
            mask[i, y0:y0 + rng.integers(5, 60), x0:x0 + rng.integers(5, 90)] = lab
        mask[i, 0, :] = 1 + i      # labelled rows / columns on the border
        mask[i, :, w - 1] = 2
    got = egress.overlay_davis(torch.from_numpy(img).to(dev), torch.from_numpy(mask).to(dev), 0.4).cpu().numpy()
    for i in range(t):  # frames are independent: no contour leaks across frame boundaries
        assert np.array_equal(got[i], EO.overlay_davis(img[i], mask[i], 0.4)), i
