"""The generated selection networks of the stack-median kernels (CPU simulation).

`pyemir_b200/csrc/stack_median_nets.inc` is compiled into `k_stack_median` / `k_abba_median`;
here the same compare-exchange lists are applied with numpy, with the kernels' padding rule
(floor((N-n)/2) times -inf below, +inf above), for every stack size n of every bucket N.
"""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "pyemir_b200", "csrc", "stack_median_nets.inc")


def load_nets():
    text = open(INC).read().replace("\\\n", " ")
    nets = {}
    for line in text.splitlines():
        m = re.match(r"#define STACK_NET_(\d+)\(CE\)(.*)", line)
        if m:
            nets[int(m.group(1))] = [(int(a), int(b))
                                     for a, b in re.findall(r"CE\((\d+), (\d+)\)", m.group(2))]
    return nets


NETS = load_nets()


def test_all_buckets_present():
    assert sorted(NETS) == [4, 8, 16, 32, 64]
    for n, net in NETS.items():
        assert all(0 <= a < b < n for a, b in net)


@pytest.mark.parametrize("N", [4, 8, 16, 32, 64])
def test_median_on_middle_wires_for_every_stack_size(N):
    rng = np.random.default_rng(N)
    for n in range(1, N + 1):          # (the kernels pick the smallest bucket >= n)
        x = np.round(rng.normal(size=(300, n)) * 3) / 3          # ties
        x[0, :] = 0.0
        x[1, 0] = np.inf
        x[2, 0] = -np.inf
        nneg = (N - n) // 2
        w = np.concatenate([x, np.full((300, nneg), -np.inf),
                            np.full((300, N - n - nneg), np.inf)], axis=1)
        for a, b in NETS[N]:
            lo, hi = np.minimum(w[:, a], w[:, b]), np.maximum(w[:, a], w[:, b])
            w[:, a], w[:, b] = lo, hi
        first = w[:, N // 2 - 1]
        second = first if n % 2 else w[:, N // 2]
        with np.errstate(invalid="ignore"):
            got = (first + second) * 0.5
            want = np.median(x, axis=1)
        assert np.array_equal(got, want, equal_nan=True), n
