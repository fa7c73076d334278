"""The oracle is pinned here: golden vectors + the reference's own C library (CPU only)."""
import numpy as np
import pytest

from conftest import random_cases


def test_oracle_sa_matches_golden(port, sa_golden):
    for name, text, sa in sa_golden:
        assert port.sa_build(text).tolist() == sa.tolist(), name


def test_reference_lib_matches_golden(ref, sa_golden):
    for name, text, sa in sa_golden:
        assert ref.sa_build(text).tolist() == sa.tolist(), name
        assert ref.sufcheck(text, sa) == 0, name


def test_oracle_sa_matches_reference_random(port, ref):
    for t in random_cases(seed=11):
        a, b = port.sa_build(t), ref.sa_build(t)
        assert (a == b).all(), (len(t), t[:16])


def test_oracle_sa_matches_reference_structured(port, ref):
    from stringsearch_b200 import synth

    for t in (synth.acgt(200_000, 1), synth.random_bytes(100_000, 2), synth.repetitive(300_000, 3, period=50),
              synth.repetitive(200_000, 4, period=1000, mutation_rate=1e-3), np.zeros(5000, np.uint8)):
        assert (port.sa_build(t) == ref.sa_build(t)).all()


def test_divsufsort_return_codes(port, ref):
    # divsufsort.c:346: NULL / negative n -> -1
    import ctypes as C

    assert port.lib.oracle_sa_build(None, None, 5) == -1
    assert ref.lib.divsufsort(None, None, 5) == -1
    buf = (C.c_uint8 * 4)()
    sa = (C.c_int32 * 4)()
    assert port.lib.oracle_sa_build(buf, sa, -1) == -1
    assert ref.lib.divsufsort(buf, sa, -1) == -1


def test_sufcheck_and_verify(port, ref):
    t = b"mississippi"
    sa = port.sa_build(t)
    assert port.sufcheck(t, sa) == 0 and ref.sufcheck(t, sa) == 0
    assert port.verify(t, sa)[0] == 0
    bad = sa.copy()
    bad[3], bad[4] = bad[4], bad[3]
    assert port.sufcheck(t, bad) == ref.sufcheck(t, bad) != 0
    rc, i = port.verify(t, bad)
    assert rc == 1 and i in (2, 3, 4)
    assert port.verify(b"", np.zeros(0, np.int32))[0] == -1  # lib.rs:143 underflow
    for c in random_cases(seed=5, sizes=(50, 300)):
        s = port.sa_build(c)
        assert port.sufcheck(c, s) == 0 and port.verify(c, s)[0] == 0


def _lcp_bruteforce(t: bytes, sa) -> list:
    out = [0] * len(sa)
    for j in range(1, len(sa)):
        a, b = t[sa[j - 1]:], t[sa[j]:]
        l = 0
        while l < len(a) and l < len(b) and a[l] == b[l]:
            l += 1
        out[j] = l
    return out


def test_lcp_oracle_matches_definition(port, sa_golden):
    """oracle_lcp_kasai (the checker of the GPU LCP kernels) against the definition itself:
    LCP[0] = 0, LCP[j] = common prefix length of the suffixes at SA[j-1] and SA[j]."""
    cases = [bytes(t) for _, t, _ in sa_golden if len(t) <= 3000]
    cases += [b"", b"a", b"aa", b"ab", b"banana", b"mississippi", b"a" * 200, b"ab" * 150, b"abcabcabcabd" * 20]
    cases += [bytes(c) for c in random_cases(seed=17, sizes=(10, 100, 700))]
    for t in cases:
        sa = port.sa_build(t)
        assert port.lcp(t, sa).tolist() == _lcp_bruteforce(t, sa.tolist()), t[:20]
    assert port.lcp(b"banana", port.sa_build(b"banana")).tolist() == [0, 1, 3, 0, 0, 2]


def test_lsm_golden(port, ref, search_golden):
    w = search_golden["worse_test"]
    t = w["text"].encode()
    sa = ref.sa_build(t)
    for c in w["cases"]:
        assert list(port.longest_substring_match(t, sa, c["needle"].encode())) == c["full"]
        ps, sas = port.part_build(t, c["partitions"], builder=ref.sa_build)
        assert list(port.part_lsm(t, ps, sas, c["needle"].encode())) == c["part"]
    e = search_golden["equivalent_test"]
    t = e["text"].encode()
    sa = ref.sa_build(t)
    for P in e["partitions"]:
        ps, sas = port.part_build(t, P, builder=ref.sa_build)
        assert len(sas) == P
        for nd in e["needles"]:
            assert list(port.longest_substring_match(t, sa, nd["needle"].encode())) == nd["expect"]
            assert list(port.part_lsm(t, ps, sas, nd["needle"].encode())) == nd["expect"]


def test_lsm_is_longest_match_bruteforce(port):
    rng = np.random.default_rng(3)
    for _ in range(300):
        n = int(rng.integers(1, 60))
        sig = int(rng.integers(1, 4))
        t = rng.integers(0, sig, n, dtype=np.uint8).tobytes()
        m = int(rng.integers(0, 12))
        p = rng.integers(0, sig, m, dtype=np.uint8).tobytes()
        sa = port.sa_build(t)
        s, l = port.longest_substring_match(t, sa, p)
        best = max(port.lib.oracle_common_prefix_len(t[i:], n - i, p, m) for i in range(n))
        assert l == best and t[s:s + l] == p[:l]


def test_lsm_edge_cases(port):
    with pytest.raises(IndexError):
        port.longest_substring_match(b"", np.zeros(0, np.int32), b"x")
    t = b"banana"
    sa = port.sa_build(t)
    # empty needle: `[] > suff` is never true, the window narrows left down to sa[0..2],
    # and the `x > y` tie (0 > 0 is false) picks the second entry (lib.rs:80-88)
    s, l = port.longest_substring_match(t, sa, b"")
    assert l == 0 and s == sa[1]


def test_sa_search_matches_reference(port, ref, search_golden):
    for c in search_golden["sa_search"]:
        t = c["text"].encode()
        sa = ref.sa_build(t)
        assert port.sa_search(t, sa, c["pattern"].encode()) == (c["count"], c["left"])
    rng = np.random.default_rng(9)
    for _ in range(400):
        n = int(rng.integers(1, 80))
        sig = int(rng.integers(1, 5))
        t = rng.integers(0, sig, n, dtype=np.uint8).tobytes()
        p = rng.integers(0, sig, int(rng.integers(0, 8)), dtype=np.uint8).tobytes()
        sa = ref.sa_build(t)
        got, exp = port.sa_search(t, sa, p), ref.sa_search(t, sa, p)
        assert got == exp, (t, p)
        cnt, left = got
        occ = sorted(i for i in range(n) if t[i:i + len(p)] == p and len(p) > 0)
        if p:
            assert sorted(sa[left:left + cnt].tolist()) == occ
    # utils.c:269-273
    assert port.sa_search(b"", np.zeros(0, np.int32), b"a") == ref.sa_search(b"", np.zeros(0, np.int32), b"a") == (0, -1)


def test_part_plan(port):
    # crates/sacapart/src/lib.rs:43,60-62 and SURVEY appendix A
    assert port.part_plan(5, 2) == (3, 2)
    assert port.part_plan(2, 5) == (1, 2)
    assert port.part_plan(0, 3) == (1, 0)
    assert port.part_plan(2**32, 8) == (536870913, 8)
    with pytest.raises(ZeroDivisionError):
        port.part_plan(5, 0)
    with pytest.raises(RuntimeError):
        port.part_lsm(b"", 1, [], b"x")  # zero partitions -> expect() panics, lib.rs:94-96


def test_batch_forms_agree(port):
    from stringsearch_b200 import synth

    t = synth.acgt(20_000, 7)
    sa = port.sa_build(t)
    flat, off = synth.patterns_from_text(t, 500, 12, 8)
    st, ln = port.lsm_batch(t, sa, (flat, off), threads=2)
    left, cnt = port.search_all_batch(t, sa, (flat, off), threads=2)
    for q in range(0, 500, 37):
        p = flat[int(off[q]):int(off[q + 1])].tobytes()
        assert (int(st[q]), int(ln[q])) == port.longest_substring_match(t, sa, p)
        assert (int(cnt[q]), int(left[q])) == port.sa_search(t, sa, p)
        assert (ln[q] == 12) == (cnt[q] > 0)
    ps, sas = port.part_build(t, 3)
    pst, pln = port.part_lsm_batch(t, ps, sas, (flat, off), threads=2)
    for q in range(0, 500, 41):
        p = flat[int(off[q]):int(off[q + 1])].tobytes()
        assert (int(pst[q]), int(pln[q])) == port.part_lsm(t, ps, sas, p)
