"""CPU: SURVEY §8(f) row 4 — mask egress.  The palette and the indexed-PNG writer are host code and are
checked completely here; the overlay oracle is checked against the vectors the unmodified reference
functions produced (tests/golden/egress.npz); the overlay kernel is checked on the GPU
(tests/test_gpu_y_egress.py)."""
import io
import os

import numpy as np
import pytest

from mivos_b200 import egress


def test_palette_equals_reference_palette(golden):
    g = golden("egress.npz")
    assert np.array_equal(egress.davis_color_map(), g["palette"])
    pal = egress.davis_palette()
    assert len(pal) == 768 and pal[:9] == [0, 0, 0, 128, 0, 0, 0, 128, 0]
    from oracle import egress_oracle as EO
    assert np.array_equal(EO.color_map(), g["palette"])


def test_overlay_oracle_matches_reference_golden(golden):
    from oracle import egress_oracle as EO
    g = golden("egress.npz")
    assert np.array_equal(EO.GUI_COLOR_MAP, g["gui_colors"])
    for a in (0.5, 0.3):
        assert np.array_equal(EO.overlay_davis(g["image"], g["mask"], a), g[f"overlay_a{int(a * 10)}"])
        assert np.array_equal(EO.overlay_davis(g["image"], g["mask"], a, fade=True), g[f"fade_a{int(a * 10)}"])


@pytest.mark.parametrize("h,w", [(37, 53), (1, 1), (480, 854)])
def test_indexed_png_roundtrip(tmp_path, h, w):
    """What eval_interactive_davis.py:91-94 writes with PIL (`fromarray` + `putpalette` + `save`) is an
    8-bit indexed PNG; ours decodes to the same pixels and palette."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(h * w)
    mask = rng.integers(0, 7, size=(h, w), dtype=np.uint8)
    im = Image.open(io.BytesIO(egress.encode_indexed_png(mask)))
    assert im.mode == "P" and im.size == (w, h)
    assert np.array_equal(np.array(im), mask)
    assert im.getpalette() == egress.davis_palette()
    # the reference's own writer and ours decode to the same thing
    ref = Image.fromarray(mask)
    ref.putpalette(egress.davis_palette())
    buf = io.BytesIO()
    ref.save(buf, format="PNG")
    back = Image.open(io.BytesIO(buf.getvalue()))
    assert np.array_equal(np.array(back), np.array(im)) and back.getpalette()[:21] == im.getpalette()[:21]
    paths = egress.save_masks_png(np.stack([mask, mask[::-1]]), str(tmp_path / "seq"), start=3)
    assert [os.path.basename(p) for p in paths] == ["00003.png", "00004.png"]
    assert np.array_equal(np.array(Image.open(paths[1])), mask[::-1])


def test_indexed_png_rejects_bad_input():
    with pytest.raises(ValueError):
        egress.encode_indexed_png(np.zeros((4, 4), dtype=np.float32))
    with pytest.raises(ValueError):
        egress.encode_indexed_png(np.zeros((4, 4), dtype=np.uint8), palette=[1, 2])
