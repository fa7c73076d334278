"""Host-side logic and the C-ABI surface, no GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, random_cases


def test_library_exports_every_declared_symbol():
    from stringsearch_b200 import _native as N

    hdr = open(os.path.join(ROOT, "include", "gsa.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gsa_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = C.CDLL(N.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/gsa.h but not exported by libgsa.so"
    assert declared == set(N.EXPORTED_SYMBOLS), declared ^ set(N.EXPORTED_SYMBOLS)


def test_library_is_sm100a_only():
    import subprocess
    from stringsearch_b200 import _native as N

    out = subprocess.run(["cuobjdump", "-lelf", N.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_argument_checks_without_gpu():
    """The reference's own precondition / tiny-input behaviour (divsufsort.c:346-349) is host logic."""
    from stringsearch_b200 import _native as N

    L = N.lib
    t = np.frombuffer(b"ba", np.uint8)
    sa = np.zeros(2, np.int32)
    assert L.gsa_divsufsort(None, N.ptr(sa), 2) == N.GSA_EINVAL
    assert L.gsa_divsufsort(N.ptr(t), None, 2) == N.GSA_EINVAL
    assert L.gsa_divsufsort(N.ptr(t), N.ptr(sa), -1) == N.GSA_EINVAL
    assert L.gsa_divsufsort(N.ptr(t), N.ptr(sa), 0) == 0
    assert L.gsa_divsufsort(N.ptr(t), N.ptr(sa), 1) == 0 and sa[0] == 0
    assert L.gsa_divsufsort(N.ptr(t), N.ptr(sa), 2) == 0 and sa.tolist() == [1, 0]
    t2 = np.frombuffer(b"ab", np.uint8)
    assert L.gsa_divsufsort(N.ptr(t2), N.ptr(sa), 2) == 0 and sa.tolist() == [0, 1]
    t3 = np.frombuffer(b"aa", np.uint8)
    assert L.gsa_divsufsort(N.ptr(t3), N.ptr(sa), 2) == 0 and sa.tolist() == [1, 0]
    h = C.c_void_p()
    assert L.gsa_part_create(N.ptr(t), 2, 0, None, 0, C.byref(h)) == N.GSA_EPANIC  # sacapart lib.rs:43
    assert L.gsa_build_workspace_bytes(0) == 0
    assert L.gsa_build_workspace_bytes(1 << 20) > 33 * (1 << 20)
    # divbwt / inverse_bw_transform / lcp: the checks of divsufsort.c:377-381 and utils.c:120-124 come first
    u = np.zeros(2, np.uint8)
    assert L.gsa_divbwt(None, N.ptr(u), None, 2) == N.GSA_EINVAL and L.gsa_divbwt(N.ptr(t), N.ptr(u), None, -1) == N.GSA_EINVAL
    assert L.gsa_divbwt(N.ptr(t), N.ptr(u), None, 0) == 0
    assert L.gsa_divbwt(N.ptr(t), N.ptr(u), None, 1) == 1 and u[0] == t[0]
    for n_, idx in ((-1, 0), (2, -1), (2, 3), (2, 0)):
        assert L.gsa_inverse_bw_transform(N.ptr(t), N.ptr(u), None, n_, idx) == N.GSA_EINVAL
    assert L.gsa_inverse_bw_transform(None, N.ptr(u), None, 2, 1) == N.GSA_EINVAL
    assert L.gsa_inverse_bw_transform(N.ptr(t), N.ptr(u), None, 0, 0) == 0
    assert L.gsa_inverse_bw_transform(N.ptr(t), N.ptr(u), None, 1, 1) == 0 and u[0] == t[0]
    lcp = np.zeros(2, np.int32)
    assert L.gsa_lcp(N.ptr(t), N.ptr(sa), N.ptr(lcp), -1, 0) == N.GSA_EINVAL
    assert L.gsa_lcp(None, N.ptr(sa), N.ptr(lcp), 2, 0) == N.GSA_EINVAL and L.gsa_lcp(N.ptr(t), None, N.ptr(lcp), 2, 0) == N.GSA_EINVAL
    assert L.gsa_lcp(N.ptr(t), N.ptr(sa), None, 2, 0) == N.GSA_EINVAL and L.gsa_lcp(N.ptr(t), N.ptr(sa), N.ptr(lcp), 0, 0) == 0
    assert L.gsa_lcp_workspace_bytes(1 << 20) >= 4 * (1 << 20) and L.gsa_inverse_bwt_workspace_bytes(1 << 20) >= 12 * (1 << 20)


def test_python_mirror_preconditions():
    from stringsearch_b200 import divsufsort, sacapart

    with pytest.raises(AssertionError, match="same len"):
        divsufsort.sort_in_place(b"abc", np.zeros(2, np.int32))
    sa = divsufsort.sort(b"")
    assert sa.sa.size == 0
    assert divsufsort.sort(b"ba").sa.tolist() == [1, 0]
    with pytest.raises(ZeroDivisionError):
        sacapart.PartitionedSuffixArray(b"abc", 0)
    with pytest.raises(IndexError):
        sa.longest_substring_match(b"x")


def test_partition_plan_matches_oracle(port):
    from stringsearch_b200.sacapart import partition_plan

    for n in (0, 1, 2, 5, 90, 1000, 2**32):
        for P in (1, 2, 3, 5, 8, 1001):
            assert partition_plan(n, P) == port.part_plan(n, P)


def test_merge_rule():
    from stringsearch_b200.sacapart import merge_results

    starts = np.array([[5, 9, 100], [50, 2, 7]], dtype=np.uint64)
    lens = np.array([[3, 4, 4], [3, 5, 4]], dtype=np.uint32)
    s, l = merge_results(starts, lens)
    assert s.tolist() == [5, 2, 7] and l.tolist() == [3, 5, 4]


def test_design_model_matches_oracle(port, sa_golden):
    """The numpy model of the GPU scheme (tests/model_doubling.py) reproduces the oracle:
    checks the design (packing, short suffixes, ordinal keys, discard) without a GPU."""
    from model_doubling import build_sa_model

    for name, text, sa in sa_golden:
        if len(text) <= 1200:
            assert build_sa_model(text).tolist() == sa.tolist(), name
    for i, t in enumerate(random_cases(seed=21, sizes=(3, 8, 9, 33, 65, 200))):
        assert (build_sa_model(t, shuffle_seed=i) == port.sa_build(t)).all()
        for ks in (1, 2, 5):  # any round-0 depth must give the same SA
            assert (build_sa_model(t, key_symbols=ks) == port.sa_build(t)).all()
        # sparse mode: ranks of suffixes unique after round 0 are recomputed lazily
        assert (build_sa_model(t, sparse=True) == port.sa_build(t)).all()
        assert (build_sa_model(t, sparse=True, key_symbols=2) == port.sa_build(t)).all()


def test_v4_model_matches_oracle(port, sa_golden):
    """tests/model_v4.py: inert-majority filter for huge groups + table-based slot ranges (the
    scheme of sa_build.cu's doubling rounds), with tiny thresholds so that small inputs use it."""
    from model_v4 import build_sa_v4

    for name, text, sa in sa_golden:
        if len(text) <= 1200:
            assert build_sa_v4(text, M=4, T=8).tolist() == sa.tolist(), name
    rng = np.random.default_rng(4)
    cases = [b"a" * 300, b"ab" * 200 + b"b" + b"ab" * 100, b"abc" * 150]
    for per, nn, mut in ((50, 2500, 20), (7, 2000, 8), (13, 1500, 0)):
        base = rng.integers(0, 256, per, dtype=np.uint8)
        x = np.tile(base, nn // per + 1)[:nn].copy()
        idx = rng.integers(0, nn, mut)
        x[idx] = rng.integers(0, 256, mut)
        cases.append(x.tobytes())
    for i, t in enumerate(cases):
        exp = port.sa_build(t)
        for (M, T) in ((4, 8), (2, 4), (8, 32)):
            st = []
            assert (build_sa_v4(t, M=M, T=T, shuffle_seed=i, stats=st) == exp).all(), (i, M, T)
        assert any(inert > 0 for (_, _, _, inert) in st)  # the filter really was exercised


def test_v5_model_matches_oracle(port, sa_golden):
    """tests/model_v5.py: v4 + the bag (tiny groups refined outside the sort; label classes by
    parity), the scheme sa_build.cu ships.  TINY=0 is the sparse-mode configuration (bag off)."""
    from model_v5 import build_sa_v5

    for name, text, sa in sa_golden:
        if len(text) <= 1200:
            assert build_sa_v5(text, M=4, T=8, TINY=3).tolist() == sa.tolist(), name
    rng = np.random.default_rng(5)
    cases = [b"a" * 300, b"ab" * 200 + b"b" + b"ab" * 100, b"abc" * 150, rng.integers(0, 2, 1500, dtype=np.uint8).tobytes()]
    for per, nn, mut in ((50, 2500, 20), (7, 2000, 8), (13, 1500, 0)):
        base = rng.integers(0, 256, per, dtype=np.uint8)
        x = np.tile(base, nn // per + 1)[:nn].copy()
        idx = rng.integers(0, nn, mut)
        x[idx] = rng.integers(0, 256, mut)
        cases.append(x.tobytes())
    bagged = 0
    for i, t in enumerate(cases):
        exp = port.sa_build(t)
        for (M, T, TINY) in ((4, 8, 3), (4, 16, 5), (8, 32, 9), (4, 8, 0)):
            st = []
            assert (build_sa_v5(t, M=M, T=T, TINY=TINY, shuffle_seed=i, stats=st) == exp).all(), (i, M, T, TINY)
            bagged += sum(r[4] for r in st) if TINY else 0
    assert bagged > 0  # the bag really was exercised


def test_synth_shapes():
    from stringsearch_b200 import synth

    assert set(np.unique(synth.acgt(1000, 1)).tolist()) <= set(b"ACGT")
    x = synth.repetitive(10_000, 3, period=100, mutation_rate=1e-2)
    assert x.size == 10_000 and (x[:100] != x[100:200]).sum() < 20
    flat, off = synth.patterns_from_text(synth.acgt(5000, 5), 10, 32, 6)
    assert flat.size == 320 and off.tolist() == list(range(0, 321, 32))


def test_tools_and_entry_points_compile():
    """bench.py, __graft_entry__.py and every script under tools/ at least parse (they only run on a GPU box)."""
    import glob
    import py_compile

    for f in [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")] + sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))):
        py_compile.compile(f, doraise=True)


def test_header_is_plain_c99(tmp_path):
    """include/gsa.h is the drop-in boundary: it must be includable from C (plain pointers and sizes only)."""
    import subprocess

    src = tmp_path / "t.c"
    src.write_text('#include "gsa.h"\nint main(void) { return (int)gsa_build_workspace_bytes(0); }\n')
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-I", os.path.join(ROOT, "include"),
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0 and not r.stderr.strip(), r.stderr


def test_memory_safety_preconditions_are_not_asserts():
    """Checks that guard buffers handed to the library must survive `python -O` and bad offsets."""
    from stringsearch_b200 import _native as N, divsufsort, sacabase

    src = open(os.path.join(ROOT, "stringsearch_b200", "divsufsort.py")).read()
    assert not re.search(r"^\s*assert ", src, flags=re.M), "bare assert statements vanish under python -O"
    with pytest.raises(AssertionError):
        divsufsort.sort_in_place(b"abc", np.zeros(2, np.int32))
    with pytest.raises(ValueError):
        divsufsort.lcp(b"abc", np.zeros(2, np.int32))
    flat = np.frombuffer(b"abcdef", np.uint8)
    with pytest.raises(ValueError):
        N.pack_patterns((flat, np.array([0, 3, 9], np.uint64)))       # runs past the pattern bytes
    with pytest.raises(ValueError):
        N.pack_patterns((flat, np.array([0, 4, 2], np.uint64)))       # not monotone
    N.pack_patterns((flat, np.array([0, 3, 6], np.uint64)))
    # SuffixArray::new with a suffix array of another length is refused before anything is uploaded
    with pytest.raises(ValueError):
        sacabase.SuffixArray(b"banana", np.array([5, 3, 1], np.int32)).longest_substring_match(b"an")
    h = C.c_void_p()
    t = np.frombuffer(b"banana", np.uint8)
    sa = np.array([5, 3, 1, 0, 4, 2], np.int32)
    assert N.lib.gsa_index_from_parts(N.ptr(t), 6, N.ptr(sa), 5, 0, C.byref(h)) == N.GSA_EINVAL


def test_bench_part_4g_text_is_synth_acgt():
    """bench.py fills the 4 GiB text of BASELINE configs[3] chunk by chunk: same bytes as synth.acgt(n, 4)."""
    from stringsearch_b200 import synth

    n, step = (1 << 20) + 8, 1 << 18
    rng = np.random.default_rng(4)
    b = np.empty(n, np.uint8)
    for lo in range(0, n, step):
        b[lo:lo + step] = synth._ACGT[rng.integers(0, 4, min(step, n - lo), dtype=np.uint8)]
    assert (b == synth.acgt(n, 4)).all()


def test_bench_whole_build_formula():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rounds = [dict(live=100, sorted=100, passes=4, bag=0), dict(live=90, sorted=10, passes=7, bag=5)]
    assert bench.whole_build_bytes(rounds) == 100 * (41 + 96) + 8 * 90 + (44 + 168) * 10 + 32 * 5
    assert bench.config_of("rep_1G", 1) == bench.config_of("rep_1G", 1) and bench.config_of("rep_1G", 1)["bytes_per_gpu"] == 1 << 30


def test_accel_model_matches_oracle(port):
    """The prefix-bucket search (search.cu) restated on the CPU (tests/model_accel.py) gives the oracle's
    answers -- i.e. the final window of sacabase's narrowing does not depend on the path, and the bucket
    bounds hold for short needles, needles with bytes that do not occur in the text, NUL bytes, suffixes
    shorter than the key, and every table width."""
    from model_accel import Accel

    rng = np.random.default_rng(2024)
    cases = 0
    for sigma, n in ((1, 40), (2, 300), (3, 500), (4, 2000), (5, 800), (17, 1500), (256, 3000), (4, 1), (4, 2), (2, 3)):
        alpha = np.sort(rng.choice(256, size=sigma, replace=False)).astype(np.uint8)
        if sigma > 1 and rng.random() < 0.5:
            alpha[0] = 0
        t = alpha[rng.integers(0, sigma, n)].tobytes()
        sa = port.sa_build(t)
        for bits in (1, 2, 5, 8, 12):
            ac = Accel(t, sa, bits)
            pats = [b"", t[:1], t[-1:], t[-3:], t, t + b"\x00", bytes([255]), bytes([0]), bytes([alpha[0]]) * 40]
            for _ in range(60):
                o, ln = int(rng.integers(0, n)), int(rng.integers(1, 24))
                p = bytearray(t[o:o + ln])
                r = rng.random()
                if r < 0.3 and p:
                    p[int(rng.integers(0, len(p)))] = int(rng.integers(0, 256))  # maybe a byte that does not occur
                elif r < 0.4:
                    p = bytearray(rng.integers(0, 256, ln, dtype=np.uint8).tobytes())
                pats.append(bytes(p))
            for p in pats:
                es, el = port.longest_substring_match(t, sa, p)
                assert ac.lsm(p) == (es, el), (sigma, n, bits, p, ac.lsm(p), (es, el))
                cnt, left = port.sa_search(t, sa, p)
                assert ac.search_all(p) == (left, cnt), (sigma, n, bits, p, ac.search_all(p), (left, cnt))
                cases += 1
    assert cases > 3000


def test_slice_plan_covers_every_needle_once():
    """ReplicatedSuffixArray's cut of a batch: pieces are contiguous, ordered, cover [0, q) exactly, a rank's pieces stay
    inside its slice [r * per, (r + 1) * per), and small slices travel in one piece."""
    from stringsearch_b200.sacapart import slice_plan

    for world in (1, 2, 3, 4, 8):
        for chunks in (1, 4):
            for chunk_min in (1, 5, 1 << 16):
                for q in list(range(0, 70)) + [1000, 4097, 10_000_000]:
                    per, starts = slice_plan(q, world, chunks, chunk_min)
                    assert len(starts) == world * chunks + 1 and starts[0] == 0 and starts[-1] == q
                    assert all(a <= b for a, b in zip(starts, starts[1:]))
                    assert per * world >= q and (per == 0) == (q == 0)
                    for r in range(world):
                        lo, hi = min(q, r * per), min(q, (r + 1) * per)
                        mine = starts[r * chunks:(r + 1) * chunks + 1]
                        assert mine[0] == lo and mine[-1] == hi
                        pieces = [b - a for a, b in zip(mine, mine[1:]) if b > a]
                        if per < chunk_min * chunks:
                            assert len(pieces) <= 1
