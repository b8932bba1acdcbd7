"""Golden vectors for SURVEY §8(f) row 4 — `overlay_davis` (interact/interactive_utils.py:119-130) and
the DAVIS colour map (util/palette.py) — produced by the UNMODIFIED reference functions imported
from /root/reference.  `interactive_utils` imports matplotlib (absent here) and the removed
`scipy.ndimage.morphology` namespace at module level; both are stubbed in sys.modules for the import
only (the functions under test use neither matplotlib nor anything but `binary_dilation`).
Run in the build container:  python -m oracle.gen_golden_egress

TEST INFRASTRUCTURE (see oracle/stm_oracle.py)."""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import egress_oracle as EO, refshim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def load_reference_utils():
    import scipy.ndimage as ndi
    stubs = {}
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            stubs.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    try:
        import scipy.ndimage.morphology  # noqa: F401
    except Exception:
        mor = types.ModuleType("scipy.ndimage.morphology")
        mor.binary_erosion, mor.binary_dilation = ndi.binary_erosion, ndi.binary_dilation
        stubs["scipy.ndimage.morphology"] = mor
    sys.modules.update(stubs)
    try:
        with refshim.reference_on_path():
            import importlib
            saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "interact" or k.startswith("interact.") or k == "dataset" or k.startswith("dataset.")}
            try:
                iu = importlib.import_module("interact.interactive_utils")
                pal = importlib.import_module("util.palette")
            finally:
                for k in list(sys.modules):
                    if k == "interact" or k.startswith("interact.") or k == "dataset" or k.startswith("dataset."):
                        sys.modules.pop(k)
                sys.modules.update(saved)
            return iu, pal
    finally:
        for k in stubs:
            sys.modules.pop(k, None)


def make_case(h=61, w=83, k=4, seed=31):
    rng = np.random.default_rng(seed)
    image = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    mask = np.zeros((h, w), dtype=np.uint8)
    mask[5:20, 3:30] = 1
    mask[15:40, 25:60] = 2
    mask[0:4, 70:83] = 3            # touches the image border
    mask[50:61, 0:10] = 4
    mask[30, 70] = 6                # a single-pixel object with the highest label the GUI table has
    mask[45:47, 40:42] = 5
    return image, mask


def main():
    iu, pal = load_reference_utils()
    report = {}
    assert np.array_equal(pal.pal_color_map(), EO.color_map())
    report["palette_equal"] = True
    image, mask = make_case()
    outs = {}
    for alpha in (0.5, 0.3):
        ref = iu.overlay_davis(image, mask, alpha)
        ora = EO.overlay_davis(image, mask, alpha)
        assert np.array_equal(ref, ora), alpha
        outs[f"overlay_a{int(alpha * 10)}"] = ref
        ref = iu.overlay_davis_fade(image, mask, alpha)
        assert np.array_equal(ref, EO.overlay_davis(image, mask, alpha, fade=True)), alpha
        outs[f"fade_a{int(alpha * 10)}"] = ref
    assert np.array_equal(iu.color_map_np, EO.GUI_COLOR_MAP)
    report["overlay_max_abs_diff_oracle_vs_reference"] = 0
    np.savez_compressed(os.path.join(OUT, "egress.npz"), image=image, mask=mask, palette=pal.pal_color_map(), gui_colors=iu.color_map_np.astype(np.uint8), **outs)
    man = json.load(open(os.path.join(OUT, "MANIFEST.json")))
    man["egress"] = report
    json.dump(man, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)
    print(json.dumps(report))


if __name__ == "__main__":
    main()
