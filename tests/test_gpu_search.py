"""Parity of the batched search kernels and the partitioned index with the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _index(text, port):
    from stringsearch_b200 import divsufsort

    sa = divsufsort.sort(text, device=0)
    assert (sa.sa == port.sa_build(text)).all()
    return sa


def _mixed_patterns(t: np.ndarray, rng, q, max_len):
    pats = []
    n = t.size
    alpha = np.unique(t)
    for i in range(q):
        m = int(rng.integers(0, max_len + 1))
        kind = i % 4
        if kind == 0 and n > 0:  # substring of the text
            o = int(rng.integers(0, n))
            p = t[o:o + m].tobytes()
        elif kind == 1 and n > 0:  # substring with the last byte changed
            o = int(rng.integers(0, n))
            b = bytearray(t[o:o + m].tobytes())
            if b:
                b[-1] = int(alpha[rng.integers(0, alpha.size)])
            p = bytes(b)
        elif kind == 2 and n > 0:  # runs off the end of the text
            o = int(rng.integers(max(0, n - max(m, 1)), n))
            p = t[o:].tobytes() + bytes(alpha[rng.integers(0, alpha.size, int(rng.integers(0, 4)))])
        else:
            p = alpha[rng.integers(0, alpha.size, m)].tobytes() if alpha.size else b"x" * m
        pats.append(p)
    return pats


def test_sacapart_reference_tests(search_golden):
    """crates/sacapart/src/lib.rs:105-165, through the GPU path."""
    from stringsearch_b200 import divsufsort, sacapart

    w = search_golden["worse_test"]
    t = w["text"].encode()
    full = divsufsort.sort(t, device=0)
    for c in w["cases"]:
        m = full.longest_substring_match(c["needle"].encode())
        assert [m.start, m.len] == c["full"] and m.as_bytes() == c["needle"].encode()[:m.len]
        part = sacapart.PartitionedSuffixArray(t, c["partitions"], devices=[0])
        pm = part.longest_substring_match(c["needle"].encode())
        assert [pm.start, pm.len] == c["part"]
    e = search_golden["equivalent_test"]
    t = e["text"].encode()
    full = divsufsort.sort(t, device=0)
    for P in e["partitions"]:
        part = sacapart.PartitionedSuffixArray(t, P, devices=[0])
        assert part.num_partitions() == P
        for nd in e["needles"]:
            fm = full.longest_substring_match(nd["needle"].encode())
            pm = part.longest_substring_match(nd["needle"].encode())
            assert [fm.start, fm.len] == nd["expect"] == [pm.start, pm.len]
            assert fm.as_bytes() == pm.as_bytes() == nd["needle"].encode()


def test_sa_search_golden(search_golden, port):
    from stringsearch_b200 import divsufsort

    for c in search_golden["sa_search"]:
        sa = divsufsort.sort(c["text"].encode(), device=0)
        left, cnt = sa.search_all_batch([c["pattern"].encode()])
        assert (int(cnt[0]), int(left[0])) == (c["count"], c["left"]), c
        assert sa.contains(c["pattern"].encode()) == (c["count"] > 0)


@pytest.mark.parametrize("sigma,n,max_len", [(4, 50_000, 40), (2, 20_000, 100), (256, 30_000, 8), (1, 3000, 70), (3, 1, 5), (3, 2, 5), (3, 3, 5)])
def test_lsm_and_search_all_match_oracle(port, sigma, n, max_len):
    rng = np.random.default_rng(sigma * 1000 + n)
    t = rng.integers(0, sigma, n, dtype=np.uint8)
    if sigma == 4:
        t = np.frombuffer(b"ACGT", np.uint8)[t]
    sa = _index(t, port)
    pats = _mixed_patterns(t, rng, 3000, max_len)
    s, l = sa.longest_substring_match_batch(pats)
    es, el = port.lsm_batch(t, sa.sa, pats)
    bad = np.flatnonzero((s != es) | (l != el))
    assert bad.size == 0, (bad[:5], [pats[i] for i in bad[:3]], s[bad[:3]], es[bad[:3]], l[bad[:3]], el[bad[:3]])
    left, cnt = sa.search_all_batch(pats)
    eleft, ecnt = port.search_all_batch(t, sa.sa, pats)
    bad = np.flatnonzero((left != eleft) | (cnt != ecnt))
    assert bad.size == 0, (bad[:5], [pats[i] for i in bad[:3]], left[bad[:3]], eleft[bad[:3]], cnt[bad[:3]], ecnt[bad[:3]])
    assert (sa.contains_batch(pats) == (ecnt > 0)).all()


@pytest.mark.parametrize("kind", ["repetitive", "binary", "one_symbol"])
def test_long_patterns_with_carried_match_lengths(port, kind):
    """Needles of up to 5000 bytes on texts with long repeats: the comparisons start at
    min(lmatch, rmatch) (sa_search, utils.c:275-286) and must give the oracle's answers."""
    from stringsearch_b200 import synth

    rng = np.random.default_rng(len(kind))
    if kind == "repetitive":
        t = synth.repetitive(200_000, 9, period=700, mutation_rate=2e-3)
    elif kind == "binary":
        t = np.tile(rng.integers(0, 2, 5000, dtype=np.uint8), 30)
        t[rng.integers(0, t.size, 40)] ^= 1
    else:
        t = np.zeros(20_000, np.uint8)
    sa = _index(t, port)
    pats = _mixed_patterns(t, rng, 600, 5000)
    # long substrings with one mutation somewhere inside, and exact long substrings
    for i in range(200):
        o = int(rng.integers(0, t.size - 10))
        m = int(rng.integers(200, 5000))
        b = bytearray(t[o:o + m].tobytes())
        if i % 2 and b:
            b[int(rng.integers(0, len(b)))] ^= 1
        pats.append(bytes(b))
    s, l = sa.longest_substring_match_batch(pats)
    es, el = port.lsm_batch(t, sa.sa, pats)
    bad = np.flatnonzero((s != es) | (l != el))
    assert bad.size == 0, (bad[:5], s[bad[:3]], es[bad[:3]], l[bad[:3]], el[bad[:3]])
    left, cnt = sa.search_all_batch(pats)
    eleft, ecnt = port.search_all_batch(t, sa.sa, pats)
    bad = np.flatnonzero((left != eleft) | (cnt != ecnt))
    assert bad.size == 0, (bad[:5], left[bad[:3]], eleft[bad[:3]], cnt[bad[:3]], ecnt[bad[:3]])


def test_search_edge_cases(port):
    from stringsearch_b200 import divsufsort, sacabase

    sa = divsufsort.sort(b"banana", device=0)
    m = sa.longest_substring_match(b"")
    assert (m.start, m.len) == port.longest_substring_match(b"banana", sa.sa, b"")
    assert sa.search_all(b"ana").tolist() == [3, 1]
    left, cnt = sa.search_all_batch([b""])
    assert (int(left[0]), int(cnt[0])) == (0, 6)  # utils.c:273
    empty = sacabase.SuffixArray(b"", np.zeros(0, np.int32))
    with pytest.raises(IndexError):
        empty.longest_substring_match(b"x")  # sacabase lib.rs:89-91 panics
    left, cnt = empty.search_all_batch([b"a"])
    assert (int(left[0]), int(cnt[0])) == (-1, 0)  # utils.c:269,272
    long_pat = b"banana" * 20  # longer than the text, > 32 bytes
    m = sa.longest_substring_match(long_pat)
    assert (m.start, m.len) == port.longest_substring_match(b"banana", sa.sa, long_pat) == (0, 6)


@pytest.mark.parametrize("P", [1, 2, 3, 7, 64])
def test_partitioned_matches_oracle(port, P):
    from stringsearch_b200 import sacapart, synth

    rng = np.random.default_rng(100 + P)
    t = synth.repetitive(60_000, 12 + P, period=300, mutation_rate=5e-3)
    part = sacapart.PartitionedSuffixArray(t, P, devices=[0])
    ps, sas = port.part_build(t, P)
    assert part.num_partitions() == len(sas) and part.partition_size() == ps
    for i in (0, len(sas) - 1):
        assert (part.shard_sa(i) == sas[i]).all()
    pats = _mixed_patterns(t, rng, 2000, 600)  # long needles: matches span partition ends (may_extend)
    # needles that start just before a partition boundary
    for i in range(1, len(sas)):
        b = i * ps
        for back in (1, 5, 100):
            if b - back >= 0:
                pats.append(t[b - back:b + 300].tobytes())
    s, l = part.longest_substring_match_batch(pats)
    es, el = port.part_lsm_batch(t, ps, sas, pats)
    bad = np.flatnonzero((s != es) | (l != el))
    assert bad.size == 0, (bad[:5], s[bad[:3]], es[bad[:3]], l[bad[:3]], el[bad[:3]])


def test_partitioned_halo_topup(port):
    """Needles longer than the default 4 KiB halo force gsa_part_lsm_batch to re-upload a larger halo."""
    from stringsearch_b200 import sacapart, synth

    t = synth.repetitive(100_000, 77, period=5000, mutation_rate=0)
    part = sacapart.PartitionedSuffixArray(t, 4, devices=[0])
    ps, sas = port.part_build(t, 4)
    pats = [t[ps - 10:ps + 9000].tobytes(), t[2 * ps - 4000:2 * ps + 6000].tobytes(), t[100:20_100].tobytes()]
    s, l = part.longest_substring_match_batch(pats)
    es, el = port.part_lsm_batch(t, ps, sas, pats)
    assert s.tolist() == es.tolist() and l.tolist() == el.tolist()


def test_partitioned_edge_cases():
    from stringsearch_b200 import sacapart

    p = sacapart.PartitionedSuffixArray(b"totor", 2, devices=[0])
    assert p.num_partitions() == 2 and p.partition_size() == 3      # "tot" | "or"
    assert sacapart.PartitionedSuffixArray(b"ab", 5, devices=[0]).num_partitions() == 2
    empty = sacapart.PartitionedSuffixArray(b"", 3, devices=[0])
    assert empty.num_partitions() == 0
    with pytest.raises(RuntimeError, match="at least one longest common substring"):
        empty.longest_substring_match(b"x")  # lib.rs:94-96


def test_multi_device_partitions(port):
    import torch
    from stringsearch_b200 import sacapart, synth

    nd = torch.cuda.device_count()
    if nd < 2:
        pytest.skip("needs >= 2 GPUs")
    rng = np.random.default_rng(3)
    t = synth.acgt(400_000, 21)
    part = sacapart.PartitionedSuffixArray(t, 8, devices=list(range(min(nd, 8))))
    ps, sas = port.part_build(t, 8)
    pats = _mixed_patterns(t, rng, 3000, 40)
    s, l = part.longest_substring_match_batch(pats)
    es, el = port.part_lsm_batch(t, ps, sas, pats)
    assert (s == es).all() and (l == el).all()


def test_full_size_part_4G_properties():
    """BASELINE config 4 at full size: 2^32 bytes of ACGT in 8 partitions of 536 870 913 bytes
    (sacapart lib.rs:43), all eight shards resident on one GPU.  Checked through properties:
    the chunk plan, an O(n) sufcheck of every shard, and needles cut from the text -- inside a
    partition (must be found in full), across a partition boundary (the match must be real and at
    least as long as the part in front of the boundary; when that part is long enough to be unique
    the match touches the end of the shard and must have been extended over it, lib.rs:77-84)."""
    import ctypes as C

    import torch

    from stringsearch_b200 import _native as N, sacapart, synth

    free, _ = torch.cuda.mem_get_info(0)
    if free < 80 << 30:
        pytest.skip(f"needs 80 GiB of free device memory, have {free >> 30}")
    n, P, m = 1 << 32, 8, 48
    t = synth.acgt(n, 4)
    psa = sacapart.PartitionedSuffixArray(t, P, devices=[0])
    ps = psa.partition_size()
    assert psa.num_partitions() == 8 and ps == 536870913
    for i in range(P):
        ix = N.lib.gsa_part_shard(psa._h, i)
        assert N.lib.gsa_index_len(ix) == min(ps, n - i * ps)
        bad = C.c_int64(-1)
        assert N.lib.gsa_index_verify(ix, C.byref(bad)) == 0, (i, bad.value)
    rng = np.random.default_rng(9)
    inside = [int(i * ps + o) for i in range(P) for o in rng.integers(0, ps - 2 * m, 40)]
    across = [int(i * ps - k) for i in range(1, P) for k in rng.integers(1, m, 20)]  # starts k bytes before a boundary
    needles = [t[o:o + m].tobytes() for o in inside + across]
    start, length = psa.longest_substring_match_batch(needles)
    for j, o in enumerate(inside + across):
        s, l = int(start[j]), int(length[j])
        assert t[s:s + l].tobytes() == needles[j][:l]
        if j < len(inside):
            assert l == m, (o, s, l)
        else:
            k = (-o) % ps  # bytes of the needle in front of the boundary
            assert l >= k, (o, k, s, l)
            if k >= 24:  # 4^24 >> n: no other place in the shard shares that many bytes
                assert (s, l) == (o, m), (o, k, s, l)
    psa.close()
