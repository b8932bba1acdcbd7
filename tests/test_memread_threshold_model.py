"""Host-side check of the rule the memory read's candidate pass added in round 2 (memread_tc.cu): the PAIR threshold
must never exceed the true k-th largest score."""
import numpy as np

from memread_list_model import pair_bound


def test_pair_bound_is_a_lower_bound_of_the_kth_largest():
    rng = np.random.default_rng(0)
    for trial in range(200):
        nb = int(rng.choice([32, 64]))
        k = int(rng.integers(1, nb + 1))
        na, nbk = int(rng.integers(nb, 40 * nb)), int(rng.integers(nb, 40 * nb))
        # heavy ties on purpose: scores from a small alphabet half of the time
        a = rng.integers(0, 6, na).astype(np.float32) if trial % 2 else rng.standard_normal(na).astype(np.float32)
        b = rng.integers(0, 6, nbk).astype(np.float32) if trial % 2 else rng.standard_normal(nbk).astype(np.float32)
        ma = np.full(nb, -np.inf, np.float32)
        mb = np.full(nb, -np.inf, np.float32)
        for j, v in enumerate(a):  # bucket = column mod NB: each bucket maximum is a distinct element
            ma[j % nb] = max(ma[j % nb], v)
        for j, v in enumerate(b):
            mb[j % nb] = max(mb[j % nb], v)
        true_kth = -np.sort(-np.concatenate([a, b]))[k - 1]
        assert pair_bound(ma, mb, k) <= true_kth
        # and the single-stream bound the kernel had before
        assert -np.sort(-ma)[k - 1] <= -np.sort(-a)[k - 1] <= max(true_kth, -np.sort(-a)[k - 1])
