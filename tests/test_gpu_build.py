"""Parity of the CUDA suffix-array construction with the oracle, through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from conftest import random_cases

pytestmark = pytest.mark.gpu


def _sort(text, stats=None):
    from stringsearch_b200 import divsufsort

    return divsufsort.sort(text, device=0, stats=stats).sa


def _expect(ref_or_port, t):
    return ref_or_port.sa_build(t)


def _assert_same(got, exp, what):
    if not (got == exp).all():
        bad = int(np.flatnonzero(got != exp)[0])
        raise AssertionError(f"{what}: first mismatch at slot {bad}: gpu {got[bad]} expected {exp[bad]} "
                             f"({int((got != exp).sum())} of {exp.size} differ)")


def test_golden_vectors(sa_golden):
    """Every golden vector of the reference's own tests (crates/divsufsort/src/lib.rs:33-86)."""
    for name, text, sa in sa_golden:
        _assert_same(_sort(text), sa, name)


def test_random_small(port):
    for t in random_cases(seed=31):
        _assert_same(_sort(t), port.sa_build(t), f"n={len(t)} {t[:12]!r}")


def test_every_length_up_to_80(port):
    rng = np.random.default_rng(5)
    for n in range(0, 81):
        for sig in (1, 2, 256):
            t = rng.integers(0, sig, n, dtype=np.uint8).tobytes()
            _assert_same(_sort(t), port.sa_build(t), f"n={n} sigma={sig}")


@pytest.mark.parametrize("name,maker", [
    ("acgt_1M", lambda s: s.acgt(1 << 20, 1)),
    ("acgt_4M_C1", lambda s: s.acgt(4 << 20, 1)),                       # BASELINE config 0
    ("rand_3M", lambda s: s.random_bytes(3 * (1 << 20) + 17, 2)),
    ("rep50_2M", lambda s: s.repetitive(2 << 20, 3, period=50, mutation_rate=1e-3)),
    ("rep1000_8M", lambda s: s.repetitive(8 << 20, 3)),                 # config 2 shape, reduced
    ("rep1000_rare_4M", lambda s: s.repetitive(4 << 20, 4, mutation_rate=1e-5)),
    ("zeros_1M", lambda s: np.zeros(1 << 20, np.uint8)),
    ("ab_1M", lambda s: np.tile(np.frombuffer(b"ab", np.uint8), 1 << 19)),
    ("two_symbols_nul_1M", lambda s: (s.random_bytes(1 << 20, 9) & 1).astype(np.uint8)),
    ("five_symbols_1M", lambda s: (s.random_bytes(1 << 20, 10) % 5).astype(np.uint8)),
    ("text_like_2M", lambda s: (s.random_bytes(2 << 20, 11) % 27 + 97).astype(np.uint8)),
    # very large groups (>= 65536 suffixes sharing a prefix): the inert-majority path of the doubling rounds
    ("rep7_4M", lambda s: s.repetitive(4 << 20, 5, period=7, mutation_rate=1e-4)),
    ("rep40_6M_rare", lambda s: s.repetitive(6 << 20, 6, period=40, mutation_rate=3e-6)),
    ("run_in_random_4M", lambda s: np.concatenate([s.random_bytes(1 << 20, 12), np.full(2 << 20, 65, np.uint8), s.random_bytes(1 << 20, 13)])),
    ("two_runs_3M", lambda s: np.concatenate([np.full(1 << 20, 66, np.uint8), s.random_bytes(1 << 20, 14), np.full(1 << 20, 66, np.uint8)])),
    ("square_4M", lambda s: np.tile(s.random_bytes(1 << 20, 15), 4)),
])
def test_structured_inputs_match_reference(ref, name, maker):
    from stringsearch_b200 import synth
    from stringsearch_b200 import _native as N

    t = maker(synth)
    stats = N.BuildStats()
    got = _sort(t, stats)
    _assert_same(got, ref.sa_build(t), name)
    assert stats.rounds >= 1 and stats.round[0].live == t.size
    print(name, "rounds", stats.rounds, [(r["depth"], r["live"], r["passes"]) for r in stats.rounds_list()])


def test_any_round0_depth_gives_the_same_sa(port, monkeypatch):
    """The round-0 key depth is a tuning knob (GSA_KEY_SYMBOLS); the SA must not depend on it."""
    from stringsearch_b200 import synth

    texts = [synth.acgt(300_000, 3), synth.random_bytes(200_000, 4), synth.repetitive(400_000, 5, period=77),
             (synth.random_bytes(100_000, 6) % 3).astype(np.uint8)]
    for t in texts:
        exp = port.sa_build(t)
        for ks in ("1", "2", "3", "7", "64"):
            monkeypatch.setenv("GSA_KEY_SYMBOLS", ks)
            _assert_same(_sort(t), exp, f"key_symbols={ks}")
        monkeypatch.delenv("GSA_KEY_SYMBOLS")


def test_sparse_mode_on_and_off(port, monkeypatch):
    """Few survivors after round 0 -> labels of unique suffixes are recomputed lazily instead of
    being scattered (RB_SPARSE).  Same SA with the mode disabled; NUL-heavy alphabets exercise the
    short-suffix correction of lazy_label."""
    from stringsearch_b200 import synth

    rng = np.random.default_rng(12)
    texts = [synth.acgt(500_000, 3), synth.random_bytes(300_000, 4), rng.integers(0, 2, 400_000, dtype=np.uint8),
             np.concatenate([rng.integers(0, 3, 200_000, dtype=np.uint8), np.zeros(40, np.uint8)]),
             np.concatenate([synth.random_bytes(100_000, 8), synth.random_bytes(100_000, 8)[:50_000]])]
    for t in texts:
        exp = port.sa_build(t)
        for ks in (None, "3", "9"):
            if ks is None:
                monkeypatch.delenv("GSA_KEY_SYMBOLS", raising=False)
            else:
                monkeypatch.setenv("GSA_KEY_SYMBOLS", ks)
            monkeypatch.delenv("GSA_NO_SPARSE", raising=False)
            _assert_same(_sort(t), exp, f"sparse allowed, key_symbols={ks}")
            monkeypatch.setenv("GSA_NO_SPARSE", "1")
            _assert_same(_sort(t), exp, f"sparse off, key_symbols={ks}")
    monkeypatch.delenv("GSA_NO_SPARSE", raising=False)
    monkeypatch.delenv("GSA_KEY_SYMBOLS", raising=False)


def test_sparse_round0_list_rebuild(port, monkeypatch):
    """Sparse round 0: the few non-unique elements are listed by k_tail_summary and handled by k_rebuild_list alone.
    Texts with short runs (groups above the list kernel's size limit, some of them straddling warps and tiles: fallback to
    the general kernel), with one dense spot (a warp sees too many to list), and plain random text; same SA with the
    list switched off."""
    from stringsearch_b200 import synth

    rng = np.random.default_rng(21)
    texts = [synth.random_bytes(700_000, 31), synth.acgt(1_200_000, 32)]
    t = synth.random_bytes(900_000, 33).copy()
    for ln in (40, 66, 70, 75, 90, 130, 66, 69):  # runs of one byte: groups of about ln - key depth suffixes
        o = int(rng.integers(0, t.size - 200))
        t[o:o + ln] = 65
    texts.append(t)
    t = synth.random_bytes(1_000_000, 34).copy()
    t[500_000:503_000] = 66  # one dense spot: ~3000 non-unique suffixes side by side (still sparse: 1e6 / 64 = 15625)
    texts.append(t)
    t = synth.acgt(800_000, 35).copy()
    t[1000:1100] = t[400_000:400_100]  # a repeat: ~100 groups of two
    texts.append(t)
    for i, t in enumerate(texts):
        exp = port.sa_build(t)
        monkeypatch.delenv("GSA_NO_LIVE_LIST", raising=False)
        _assert_same(_sort(t), exp, f"text {i}, list rebuild allowed")
        monkeypatch.setenv("GSA_NO_LIVE_LIST", "1")
        _assert_same(_sort(t), exp, f"text {i}, list rebuild off")
    monkeypatch.delenv("GSA_NO_LIVE_LIST", raising=False)


def test_fuzz_repetitive_structures(ref):
    """Randomised periodic / run-heavy / copy-heavy texts with groups from a few hundred to a few hundred thousand suffixes
    (the group tables start at 512): every combination of verdicts (label moved, group became small, unique,
    vanished), several huge groups interacting, NUL-heavy alphabets."""
    from stringsearch_b200 import _native as N

    rng = np.random.default_rng(20261017)
    for case in range(36):
        n = int(rng.integers(150_000, 1_500_000))
        kind = case % 6
        if kind == 0:      # short period, sparse mutations
            per = int(rng.integers(1, 40))
            x = np.tile(rng.integers(0, 256, per, dtype=np.uint8), n // per + 1)[:n].copy()
            m = int(rng.integers(0, 60))
            x[rng.integers(0, n, m)] = rng.integers(0, 256, m, dtype=np.uint8)
        elif kind == 1:    # a few long runs of the same byte inside random text
            x = rng.integers(0, 4, n, dtype=np.uint8)
            for _ in range(int(rng.integers(1, 4))):
                a = int(rng.integers(0, n - 100_000))
                x[a:a + int(rng.integers(70_000, 400_000))] = int(rng.integers(0, 2))
        elif kind == 2:    # the same block copied many times, with a few edits per copy
            blk = rng.integers(0, 256, int(rng.integers(1_000, 20_000)), dtype=np.uint8)
            x = np.tile(blk, n // blk.size + 1)[:n].copy()
            m = int(rng.integers(0, 200))
            x[rng.integers(0, n, m)] = rng.integers(0, 256, m, dtype=np.uint8)
        elif kind == 3:    # two interleaved periodic texts over {0, 1} (NUL is a symbol)
            p1, p2 = int(rng.integers(2, 9)), int(rng.integers(2, 9))
            a = np.tile(rng.integers(0, 2, p1, dtype=np.uint8), n // p1 + 1)[:n]
            b = np.tile(rng.integers(0, 2, p2, dtype=np.uint8), n // p2 + 1)[:n]
            cut = int(rng.integers(n // 4, 3 * n // 4))
            x = np.concatenate([a[:cut], b[cut:]])
        elif kind == 4:    # all one byte, except a handful of positions
            x = np.full(n, int(rng.integers(0, 256)), np.uint8)
            m = int(rng.integers(0, 6))
            x[rng.integers(0, n, m)] = rng.integers(0, 256, m, dtype=np.uint8)
        else:              # Fibonacci-like word (many nested repeats)
            a, b = b"a", b"ab"
            while len(b) < n:
                a, b = b, b + a
            x = np.frombuffer(b[:n], dtype=np.uint8).copy()
        st = N.BuildStats()
        got = _sort(x, st)
        _assert_same(got, ref.sa_build(x), f"fuzz case {case} kind {kind} n={n}")


@pytest.mark.parametrize("repeats", [31, 32, 33, 34, 255, 256, 257, 510, 511, 512, 513, 514, 1023, 1024, 1025, 3000])
def test_group_sizes_around_the_class_thresholds(ref, repeats):
    """A random block repeated r times (plus a few mutations) gives groups of about r suffixes:
    r around 32 (bag / sort path), 256 (label granularity of the group tables) and 512 (group
    tables), where groups change class while they shrink."""
    rng = np.random.default_rng(repeats)
    block = rng.integers(0, 4, 700, dtype=np.uint8)
    t = np.tile(block, repeats)
    for frac in (0.0, 3e-4):
        x = t.copy()
        k = int(x.size * frac)
        if k:
            x[rng.integers(0, x.size, k)] = rng.integers(0, 4, k)
        _assert_same(_sort(x), ref.sa_build(x), f"repeats={repeats} mutations={k}")


def test_fuzz_many_small_structured_texts(port):
    """1500 small texts (1..20000 bytes) built from repeats of short blocks, runs, copies and
    mutations over alphabets of 1..6 symbols: every one has many equal-prefix groups whose sizes
    sit around the bag tile (480 + 32 entries), the warp and the 4096-element tiles."""
    rng = np.random.default_rng(424242)
    for case in range(1500):
        n = int(rng.integers(1, 20000 if case % 10 == 0 else 3000))
        sigma = int(rng.integers(1, 7))
        kind = case % 5
        if kind == 0:
            t = np.tile(rng.integers(0, sigma, int(rng.integers(1, 40)), dtype=np.uint8), n)[:n]
        elif kind == 1:
            t = rng.integers(0, sigma, n, dtype=np.uint8)
        elif kind == 2:  # copies of a block with single-symbol separators
            b = rng.integers(0, sigma, int(rng.integers(1, 300)), dtype=np.uint8)
            t = np.concatenate([np.concatenate([b, rng.integers(0, sigma, 1, dtype=np.uint8)]) for _ in range(n // (b.size + 1) + 1)])[:n]
        elif kind == 3:  # runs
            t = np.repeat(rng.integers(0, sigma, n // 7 + 1, dtype=np.uint8), rng.integers(1, 15, n // 7 + 1))[:n]
        else:  # periodic with mutations
            t = np.tile(rng.integers(0, sigma, int(rng.integers(2, 200)), dtype=np.uint8), n)[:n].copy()
            k = max(1, n // 200)
            t[rng.integers(0, n, k)] = rng.integers(0, sigma, k)
        t = np.ascontiguousarray(t, dtype=np.uint8)
        _assert_same(_sort(t), port.sa_build(t), f"case {case} kind {kind} n={t.size} sigma={sigma}")


def test_inert_filter_on_and_off(ref, monkeypatch):
    """GSA_NO_INERT=1 sorts every live suffix in every round (no huge-group filter): same SA."""
    from stringsearch_b200 import synth
    from stringsearch_b200 import _native as N

    t = synth.repetitive(3 << 20, 21, period=11, mutation_rate=2e-5)
    exp = ref.sa_build(t)
    st_on, st_off = N.BuildStats(), N.BuildStats()
    monkeypatch.delenv("GSA_NO_INERT", raising=False)
    _assert_same(_sort(t, st_on), exp, "filter on")
    monkeypatch.setenv("GSA_NO_INERT", "1")
    _assert_same(_sort(t, st_off), exp, "filter off")
    monkeypatch.delenv("GSA_NO_INERT", raising=False)
    on = sum(r["sorted"] for r in st_on.rounds_list()[1:])
    off = sum(r["sorted"] for r in st_off.rounds_list()[1:])
    assert on < off // 2, (on, off)  # the filter really skipped most of the sorting work
    # thousands of groups just above the group-table threshold, all of them re-created in every
    # round when the filter is off (the list of huge groups then holds every label twice)
    rng = np.random.default_rng(5)
    t2 = np.tile(rng.integers(0, 256, 3000, dtype=np.uint8), 700)
    t2[rng.integers(0, t2.size, 50)] ^= 1
    exp2 = ref.sa_build(t2)
    _assert_same(_sort(t2), exp2, "many groups, filter on")
    monkeypatch.setenv("GSA_NO_INERT", "1")
    _assert_same(_sort(t2), exp2, "many groups, filter off")
    monkeypatch.delenv("GSA_NO_INERT", raising=False)


def test_bwt_matches_reference(ref, sa_golden):
    """gsa_divbwt vs the reference's divbwt (divsufsort.c:372-405): same string, same primary index."""
    from stringsearch_b200 import divsufsort, synth

    texts = [t for _, t, _ in sa_golden] + [synth.acgt(300_000, 2), synth.repetitive(500_000, 3, period=60),
                                            synth.random_bytes(100_001, 4), np.zeros(70_000, np.uint8)]
    for t in texts:
        u, pidx = divsufsort.bwt(t)
        eu, epidx = ref.divbwt(t)
        assert pidx == epidx and (u == eu).all(), (len(t), pidx, epidx)


def test_inverse_bwt_round_trip_and_reference(ref, sa_golden):
    """gsa_inverse_bw_transform vs the reference's inverse_bw_transform (utils.c:111-156), and the
    round trip text -> divbwt -> inverse == text, on inputs that cross the 1024-row splitter
    stride, the 4096-element radix tile, and with one-symbol / two-symbol alphabets."""
    from stringsearch_b200 import divsufsort, synth
    from stringsearch_b200 import _native as N

    rng = np.random.default_rng(12)
    texts = [t for _, t, _ in sa_golden if len(t) > 0]
    texts += [b"ab", b"ba", b"aa", b"banana", b"mississippi", np.zeros(5000, np.uint8), synth.acgt(300_000, 2),
              synth.repetitive(500_000, 3, period=60), synth.random_bytes(100_001, 4), synth.random_bytes(1 << 20, 6),
              np.tile(np.frombuffer(b"ab", np.uint8), 40_000), rng.integers(0, 2, 1023, dtype=np.uint8),
              rng.integers(0, 2, 1024, dtype=np.uint8), rng.integers(0, 2, 1025, dtype=np.uint8),
              rng.integers(0, 256, 4097, dtype=np.uint8), synth.repetitive(3 << 20, 5, period=7, mutation_rate=1e-4)]
    for t in texts:
        t = N.as_u8(t)
        u, pidx = ref.divbwt(t)
        got = divsufsort.inverse_bwt(u, pidx)
        rc, exp = ref.inverse_bwt(u, pidx)
        assert rc == 0 and (got == t).all(), (t.size, pidx)
        if t.size > 1:  # n == 1: the reference returns before writing U (utils.c:124)
            assert (got == exp).all(), (t.size, pidx)
        u2, pidx2 = divsufsort.bwt(t)
        assert (divsufsort.inverse_bwt(u2, pidx2) == t).all()
    # argument checks of utils.c:120-123, same return codes
    one = np.zeros(4, np.uint8)
    for n, idx in ((-1, 0), (4, -1), (4, 5), (4, 0)):
        assert N.lib.gsa_inverse_bw_transform(N.ptr(one), N.ptr(one), None, n, idx) == -1
        assert ref.lib.inverse_bw_transform(one.ctypes.data_as(C.POINTER(C.c_uint8)), one.ctypes.data_as(C.POINTER(C.c_uint8)), None, n, idx) == -1
    assert N.lib.gsa_inverse_bw_transform(None, N.ptr(one), None, 4, 1) == -1
    assert N.lib.gsa_inverse_bw_transform(N.ptr(one), N.ptr(one), None, 0, 0) == 0


def test_tile_boundaries(port):
    """Sizes straddling the radix tile (4096), rebuild tile (2048) and vector widths."""
    rng = np.random.default_rng(7)
    for n in (2047, 2048, 2049, 4095, 4096, 4097, 8191, 8192, 8193, 12289, 65535, 65536, 65537):
        t = rng.integers(0, 3, n, dtype=np.uint8)
        _assert_same(_sort(t), port.sa_build(t), f"n={n}")


def test_device_resident_api_and_sufcheck():
    import torch
    from stringsearch_b200 import synth
    from stringsearch_b200 import _native as N

    t = synth.repetitive(1 << 20, 5, period=200)
    d_t = torch.from_numpy(t).cuda()
    d_sa = torch.empty(t.size, dtype=torch.int32, device="cuda")
    stats = N.BuildStats()
    stream = torch.cuda.current_stream().cuda_stream
    rc = N.lib.gsa_build_device(d_t.data_ptr(), d_sa.data_ptr(), t.size, None, 0, stream, C.byref(stats))
    assert rc == 0, N.last_error()
    bad = C.c_int64(-1)
    assert N.lib.gsa_sufcheck_device(d_t.data_ptr(), d_sa.data_ptr(), t.size, stream, C.byref(bad)) == 0
    # caller-provided workspace
    ws_bytes = N.lib.gsa_build_workspace_bytes(t.size)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    d_sa2 = torch.empty_like(d_sa)
    rc = N.lib.gsa_build_device(d_t.data_ptr(), d_sa2.data_ptr(), t.size, ws.data_ptr(), ws_bytes, stream, None)
    assert rc == 0 and torch.equal(d_sa, d_sa2)
    assert N.lib.gsa_build_device(d_t.data_ptr(), d_sa2.data_ptr(), t.size, ws.data_ptr(), 1024, stream, None) == N.GSA_EINVAL
    # a corrupted SA must be rejected: swap two slots / duplicate one / out of range
    for mutate in ("swap", "dup", "range"):
        x = d_sa.clone()
        if mutate == "swap":
            x[1000], x[1001] = d_sa[1001].item(), d_sa[1000].item()
        elif mutate == "dup":
            x[5] = x[6]
        else:
            x[77] = t.size
        assert N.lib.gsa_sufcheck_device(d_t.data_ptr(), x.data_ptr(), t.size, stream, C.byref(bad)) == 1
        assert bad.value >= 0


def test_verify_method_and_notsorted(port):
    from stringsearch_b200 import sacabase

    t = b"mississippi"
    sa = port.sa_build(t)
    sacabase.SuffixArray(t, sa).verify()
    bad = sa.copy()
    bad[3], bad[4] = bad[4], bad[3]
    with pytest.raises(sacabase.NotSorted):
        sacabase.SuffixArray(t, bad).verify()


def test_notsorted_names_the_pair_the_reference_names(port):
    """sacabase::verify returns the FIRST adjacent pair that is out of order (lib.rs:143-147); the O(n)
    rank criterion alone would flag (0, 1) here, whose own order is fine."""
    from stringsearch_b200 import sacabase

    with pytest.raises(sacabase.NotSorted) as e:
        sacabase.SuffixArray(b"abac", np.array([0, 2, 3, 1], np.int32)).verify()
    assert (e.value.i, e.value.j) == (2, 3)
    rng = np.random.default_rng(12)
    for _ in range(40):
        n = int(rng.integers(2, 200))
        t = rng.integers(0, 3, n, dtype=np.uint8).tobytes()
        sa = port.sa_build(t)
        i, j = sorted(rng.integers(0, n, 2))
        if i == j:
            continue
        sa[i], sa[j] = sa[j], sa[i]
        rc, bad = port.verify(t, sa)
        assert rc == 1
        with pytest.raises(sacabase.NotSorted) as e:
            sacabase.SuffixArray(t, sa).verify()
        assert e.value.i == bad, (t, sa.tolist(), e.value.i, bad)


def test_user_supplied_sa_is_range_checked():
    """SuffixArray::new(text, sa) with an entry that is no text position: the reference panics on its slice
    bounds check (lib.rs:53-57); here the upload refuses it instead of searching out of bounds."""
    from stringsearch_b200 import sacabase

    for bad_entry in (6, 2**31 - 1, -1):
        sa = np.array([5, 3, 1, 0, bad_entry, 2], np.int32)
        with pytest.raises(IndexError):
            sacabase.SuffixArray(b"banana", sa).longest_substring_match(b"nan")


def test_sort_keeps_build_and_index_on_one_device():
    from stringsearch_b200 import divsufsort

    s = divsufsort.sort(b"mississippi", device=0)
    assert s._device == 0
    s.verify()
    assert divsufsort.sort(b"mississippi")._device is not None


def test_concurrent_builds_from_threads(port):
    """The ABI is re-entrant: sacapart calls its builder from several threads
    (crates/sacapart/src/lib.rs:41,45-49)."""
    from concurrent.futures import ThreadPoolExecutor
    from stringsearch_b200 import synth

    texts = [synth.repetitive(300_000 + 1000 * i, 40 + i, period=64) for i in range(6)]
    with ThreadPoolExecutor(6) as ex:
        got = list(ex.map(_sort, texts))
    for g, t in zip(got, texts):
        _assert_same(g, port.sa_build(t), "threaded")


def _ref_sa_in_background(ref, t):
    """Start the reference's C divsufsort (oracle/_ref, 1 thread; ctypes releases the GIL) on `t`
    and return a function that waits for its SA."""
    from concurrent.futures import ThreadPoolExecutor

    ex = ThreadPoolExecutor(1)
    fut = ex.submit(ref.sa_build, t)

    def wait():
        try:
            return fut.result()
        finally:
            ex.shutdown(wait=False)

    return wait


def _full_size_case(ref, t, what, min_rounds=0):
    """Build on the GPU through the device ABI and through the host-pointer ABI, and compare BOTH,
    every slot, with the reference's own C libdivsufsort run on the same bytes; the O(n) GPU
    sufcheck stays as a second, independent check."""
    import torch
    from stringsearch_b200 import _native as N

    wait_ref = _ref_sa_in_background(ref, t)
    n = int(t.size)
    d_t = torch.from_numpy(t).cuda()
    d_sa = torch.empty(n, dtype=torch.int32, device="cuda")
    stats = N.BuildStats()
    stream = torch.cuda.current_stream().cuda_stream
    assert N.lib.gsa_build_device(d_t.data_ptr(), d_sa.data_ptr(), n, None, 0, stream, C.byref(stats)) == 0, N.last_error()
    bad = C.c_int64(-1)
    assert N.lib.gsa_sufcheck_device(d_t.data_ptr(), d_sa.data_ptr(), n, stream, C.byref(bad)) == 0, bad.value
    assert stats.rounds >= min_rounds
    sa_dev = d_sa.cpu().numpy()
    del d_t, d_sa
    torch.cuda.empty_cache()
    sa_host = np.empty(n, dtype=np.int32)
    assert N.lib.gsa_divsufsort(t.ctypes.data, sa_host.ctypes.data, n) == 0, N.last_error()  # the drop-in call
    exp = wait_ref()
    _assert_same(sa_dev, exp, f"{what}: gsa_build_device vs reference C divsufsort")
    _assert_same(sa_host, exp, f"{what}: gsa_divsufsort vs reference C divsufsort")
    print(what, stats.ms_total, "ms", [(r["depth"], r["live"], r["passes"]) for r in stats.rounds_list()])


@pytest.mark.timeout(900)
def test_full_size_rand_256M_bit_exact(ref):
    """BASELINE config "SA of 256 MiB uniform-random bytes on 1 B200, bit-exact vs cdivsufsort":
    all 2^28 slots are compared with the reference's C library (crates/cdivsufsort/src/lib.rs:9-30)."""
    from stringsearch_b200 import synth

    _full_size_case(ref, synth.random_bytes(1 << 28, 2), "rand_256M")


@pytest.mark.timeout(1500)
def test_full_size_rep_1G_bit_exact(ref):
    """BASELINE config "SA of 1 GiB highly repetitive text" at full size (period 1000, 1e-3
    mutations, many doubling rounds): all 2^30 slots against the reference's C library."""
    from stringsearch_b200 import synth

    _full_size_case(ref, synth.repetitive(1 << 30, 3), "rep_1G", min_rounds=10)


@pytest.mark.timeout(900)
def test_full_size_acgt_shard_bit_exact(ref):
    """One shard of BASELINE config "PartitionedSuffixArray of 4 GiB input, 8 x 512 MiB shards":
    chunk = 2^32 / 8 + 1 = 536 870 913 bytes of ACGT (crates/sacapart/src/lib.rs:43)."""
    from stringsearch_b200 import synth

    _full_size_case(ref, synth.acgt(536870913, 4), "acgt_512M_shard")


@pytest.mark.timeout(1800)
def test_largest_supported_text_properties():
    """n = 2^31 - 2, the largest text the reference accepts (crates/divsufsort/src/divsufsort.rs:9-13:
    `text.len() < i32::MAX`): 32-bit index arithmetic at the limit.  Random bytes generated on the
    device; checked with the O(n) GPU sufcheck.  Needs ~130 GB of HBM (skipped if not free)."""
    import torch
    from stringsearch_b200 import _native as N

    n = 2**31 - 2
    free, _ = torch.cuda.mem_get_info()
    need = N.lib.gsa_build_workspace_bytes(n) + 10 * n + (4 << 30)
    if free < need:
        pytest.skip(f"needs {need >> 30} GiB of free device memory, have {free >> 30}")
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    d_t = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda", generator=g)
    d_t[1_000_000:1_200_000] = 7          # a run: a huge group next to the unique suffixes
    d_t[-5:] = 0                          # NULs at the very end: short-suffix ordering
    d_sa = torch.empty(n, dtype=torch.int32, device="cuda")
    stats = N.BuildStats()
    stream = torch.cuda.current_stream().cuda_stream
    rc = N.lib.gsa_build_device(d_t.data_ptr(), d_sa.data_ptr(), n, None, 0, stream, C.byref(stats))
    assert rc == 0, N.last_error()
    bad = C.c_int64(-1)
    assert N.lib.gsa_sufcheck_device(d_t.data_ptr(), d_sa.data_ptr(), n, stream, C.byref(bad)) == 0, bad.value
    print("n=2^31-2:", stats.ms_total, "ms", [(r["depth"], r["live"], r["sorted"], r["passes"]) for r in stats.rounds_list()])
