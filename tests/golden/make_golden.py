"""Generates tests/golden/*.json from the reference itself (run in the build container,
where /root/reference exists; the GPU box only reads the committed JSON).

* sa_golden.json      : inputs of the reference's own SACA tests
                        (crates/divsufsort/src/lib.rs:33-86: fuzz1-3, crash-*, shruggy) plus the
                        hand vectors of SURVEY.md section 4, each with the suffix array produced by
                        the reference's C libdivsufsort (oracle/_ref, built from
                        crates/cdivsufsort/c-sources).  Inputs are stored hex-encoded.
* search_golden.json  : the expectations of crates/sacapart/src/lib.rs:105-165
                        (worse_test, equivalent_test) restated as (start, len) tuples, and
                        sa_search answers of the reference C library for a few patterns.
"""
import glob
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle  # noqa: E402

REF_TESTDATA = "/root/reference/crates/divsufsort/src/testdata"


def main():
    ref = oracle.ref()
    port = oracle.port()
    cases = []
    hand = {
        "banana": b"banana", "mississippi": b"mississippi", "totor": b"totor", "aaa": b"aaa",
        "nul3": b"\0\0\0", "a_nul": b"a\0", "zero_padding_trap": b"a\0\0\0\0\0\0\0\0a\0",
        "abab16": b"abababababababab", "ff_ff_fe_ff_ff": b"\xff\xff\xfe\xff\xff",
        "shruggy": "¯\\_(ツ)_/¯".encode(), "empty": b"", "a": b"a", "ab": b"ab", "ba": b"ba", "aa": b"aa",
    }
    for name, data in hand.items():
        cases.append({"name": name, "source": "SURVEY.md section 4 / crates/divsufsort/src/lib.rs:84-86",
                      "text_hex": data.hex(), "sa": ref.sa_build(data).tolist()})
    for path in sorted(glob.glob(os.path.join(REF_TESTDATA, "*"))):
        data = open(path, "rb").read()
        sa = ref.sa_build(data)
        assert ref.sufcheck(data, sa) == 0
        cases.append({"name": os.path.basename(path), "source": "crates/divsufsort/src/testdata (lib.rs:33-81)",
                      "text_hex": data.hex(), "sa": sa.tolist()})
    json.dump({"generator": "tests/golden/make_golden.py", "producer": "reference C libdivsufsort (oracle/_ref)",
               "cases": cases}, open(os.path.join(HERE, "sa_golden.json"), "w"))

    sentence = ("This is a rather long text. We can probably find matches that span two partitions. Oh yes.")
    search = {
        "generator": "tests/golden/make_golden.py",
        # crates/sacapart/src/lib.rs:105-126 (worse_test)
        "worse_test": {"text": "totor", "cases": [
            {"needle": "tor", "full": [2, 3], "partitions": 2, "part": [0, 2]},
            {"needle": "otor", "full": [1, 4], "partitions": 2, "part": [1, 4]},
        ]},
        # crates/sacapart/src/lib.rs:128-165 (equivalent_test): partitioned == full for P in 1,2,3
        "equivalent_test": {"text": sentence, "partitions": [1, 2, 3], "needles": [
            {"needle": "rather long", "expect": [10, 11]},
            {"needle": "text. We can", "expect": [22, 12]},
            {"needle": "We can probably find matches that span", "expect": [28, 38]},
        ]},
        "sa_search": [],
    }
    # cross-check the tuples above with the oracle port before writing them
    t = sentence.encode()
    sa = ref.sa_build(t)
    for nd in search["equivalent_test"]["needles"]:
        got = port.longest_substring_match(t, sa, nd["needle"].encode())
        assert list(got) == nd["expect"], (nd, got)
        assert t[got[0]:got[0] + got[1]] == nd["needle"].encode()
    for text, pats in ((b"banana", [b"ana", b"a", b"nan", b"x", b"", b"banana", b"bananas", b"b"]),
                       (b"mississippi", [b"ssi", b"i", b"issi", b"p", b"z", b"mississippi", b"sip"])):
        sa = ref.sa_build(text)
        for p in pats:
            cnt, idx = ref.sa_search(text, sa, p)
            search["sa_search"].append({"text": text.decode(), "pattern": p.decode(), "count": cnt, "left": idx})
    json.dump(search, open(os.path.join(HERE, "search_golden.json"), "w"), indent=1)
    print("wrote", len(cases), "SA cases")


if __name__ == "__main__":
    main()
