"""LCP array on the GPU (lcp.cu) against the Kasai oracle, through the C ABI."""
import numpy as np
import pytest

from conftest import random_cases

pytestmark = pytest.mark.gpu


def _check(port, t, what):
    from stringsearch_b200 import divsufsort

    t = np.frombuffer(bytes(t), np.uint8) if not isinstance(t, np.ndarray) else t
    sa = port.sa_build(t)
    got = divsufsort.lcp(t, sa, device=0)
    exp = port.lcp(t, sa)
    if not (got == exp).all():
        bad = int(np.flatnonzero(got != exp)[0])
        raise AssertionError(f"{what}: LCP[{bad}] = {got[bad]}, expected {exp[bad]} ({int((got != exp).sum())} of {exp.size} differ)")


def test_lcp_golden_and_small(port, sa_golden):
    for name, text, _ in sa_golden:
        _check(port, text, name)
    for t in (b"a", b"aa", b"ab", b"ba", b"banana", b"mississippi", b"a\0\0\0\0\0\0\0\0a\0"):
        _check(port, t, repr(t))
    for t in random_cases(seed=41):
        _check(port, t, f"n={len(t)}")
    from stringsearch_b200 import divsufsort

    assert divsufsort.lcp(b"", np.zeros(0, np.int32)).size == 0
    assert divsufsort.lcp(b"banana", port.sa_build(b"banana")).tolist() == [0, 1, 3, 0, 0, 2]


@pytest.mark.parametrize("name,maker", [
    ("acgt_1M", lambda s: s.acgt(1 << 20, 1)),
    ("rand_1M", lambda s: s.random_bytes((1 << 20) + 13, 2)),
    ("rep50_1M", lambda s: s.repetitive(1 << 20, 3, period=50, mutation_rate=1e-3)),     # long matches: k_long
    ("rep1000_4M", lambda s: s.repetitive(4 << 20, 3)),
    ("rep7_rare_2M", lambda s: s.repetitive(2 << 20, 5, period=7, mutation_rate=1e-5)),  # LCPs of ~100 KiB
    ("zeros_1M", lambda s: np.zeros(1 << 20, np.uint8)),                                   # one irreducible LCP of n-1
    ("ab_1M", lambda s: np.tile(np.frombuffer(b"ab", np.uint8), 1 << 19)),
    ("square_2M", lambda s: np.tile(s.random_bytes(1 << 19, 15), 4)),                      # LCPs up to 3n/4
    ("two_symbols_1M", lambda s: (s.random_bytes(1 << 20, 9) & 1).astype(np.uint8)),
    ("text_like_1M", lambda s: (s.random_bytes(1 << 20, 11) % 27 + 97).astype(np.uint8)),
])
def test_lcp_structured(port, name, maker):
    from stringsearch_b200 import synth

    _check(port, maker(synth), name)


def test_lcp_sizes_around_tiles_and_words(port):
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 8, 9, 63, 64, 65, 127, 129, 4095, 4096, 4097, 8193, 65537):
        _check(port, rng.integers(0, 2, n, dtype=np.uint8), f"binary n={n}")
        _check(port, np.zeros(n, np.uint8), f"zeros n={n}")


def test_sort_with_lcp_one_call(port):
    from stringsearch_b200 import divsufsort, synth

    t = synth.repetitive(700_001, 8, period=300, mutation_rate=1e-3)
    sa, lcp = divsufsort.sort_with_lcp(t, device=0)
    exp_sa = port.sa_build(t)
    assert (sa.sa == exp_sa).all()
    assert (lcp == port.lcp(t, exp_sa)).all()


def test_lcp_device_api_with_caller_workspace(port):
    """gsa_lcp_device on device-resident text + SA with a caller-provided workspace."""
    import torch

    from stringsearch_b200 import _native as N, synth

    t = synth.acgt(300_000, 21)
    sa = port.sa_build(t)
    d_t = torch.from_numpy(t).cuda()
    d_sa = torch.from_numpy(sa).cuda()
    d_lcp = torch.empty(t.size, dtype=torch.int32, device="cuda")
    ws = torch.empty(N.lib.gsa_lcp_workspace_bytes(t.size), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = N.lib.gsa_lcp_device(d_t.data_ptr(), d_sa.data_ptr(), d_lcp.data_ptr(), t.size, ws.data_ptr(), ws.numel(), st)
    assert rc == 0, N.last_error()
    assert (d_lcp.cpu().numpy() == port.lcp(t, sa)).all()
    small = torch.empty(16, dtype=torch.uint8, device="cuda")
    assert N.lib.gsa_lcp_device(d_t.data_ptr(), d_sa.data_ptr(), d_lcp.data_ptr(), t.size, small.data_ptr(), 16, st) == -1


def test_lcp_full_size_rep_256M_properties():
    """256 MiB repetitive text: LCP checked through properties that need no CPU pass over it:
    a sample of entries against direct comparison, and sum(LCP) against the same sum computed
    from the permuted array (PLCP[i] >= PLCP[i-1] - 1 holds for every i)."""
    from stringsearch_b200 import divsufsort, synth

    t = synth.repetitive(1 << 28, 3)
    sa, lcp = divsufsort.sort_with_lcp(t, device=0)
    s = sa.sa
    assert lcp[0] == 0 and lcp.min() >= 0
    rng = np.random.default_rng(0)
    for j in rng.integers(1, t.size, 300):
        a, b, l = int(s[j - 1]), int(s[j]), int(lcp[j])
        assert (t[a:a + l] == t[b:b + l]).all()
        assert a + l == t.size or b + l == t.size or t[a + l] != t[b + l]
    plcp = np.empty(t.size, np.int64)
    plcp[s] = lcp
    assert (plcp[1:] >= plcp[:-1] - 1).all()
