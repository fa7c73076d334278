"""python model of the v5 construction scheme (TEST INFRASTRUCTURE, design validation only).

v5 = v4 + the BAG: suffixes whose group has at most TINY members leave the text-order walk /
radix-sort path for good.  They are kept group by group (any order of groups) in a list of
(suffix, slot) entries and refined in place: gather label(i + h), sort inside the group,
split, finalise.  The class of a group is readable from its label: multiple of M = huge, other
even = medium, odd = tiny (always the canonical label of its slot range).

(v4 docstring follows.)

v4 = v3 (label ranks, text-order walk) + "inert majority": in a huge group most members share
one sort key per round (repetitive text).  Those members are not sorted at all: they stay in
the live list, keep their label, and only the group's slot range in a per-group table shrinks.

* labels: 1-based slots; a label that is a multiple of M marks a HUGE group (range size >= T >= M);
  small groups never use such labels.  Huge groups keep their range in tables GS/GE indexed by
  label and are resolved through them (a label that fell out of its range, a group that became
  small or unique, is fixed lazily and deterministically by every reader).
* per round and huge group one key half rho* is elected (first come); members with
  label(i+h) == rho* are inert.  The others (and all members of small groups) are sorted by
  (label, label(i+h)) and re-grouped; their slots follow from the group's range and their
  position in the run: "<rho*" members fill the range from the left, ">rho*" from the right.
"""
from __future__ import annotations

import numpy as np

DEAD = 1 << 31
LMASK = DEAD - 1


def bits_for(v: int) -> int:
    b = 1
    while b < 64 and (v >> b) != 0:
        b += 1
    return b


def build_sa_v5(text: bytes, M: int = 4, T: int = 8, TINY: int = 3, shuffle_seed=None, stats=None) -> np.ndarray:
    assert T >= 2 * M and M >= 2 and M % 2 == 0 and TINY < T
    assert TINY == 0 or (TINY >= 3 and M >= 4)  # a medium group must contain an even label that is no multiple of M
    t = np.frombuffer(bytes(text), dtype=np.uint8)
    n = t.size
    if n == 0:
        return np.zeros(0, np.int32)
    present = np.zeros(256, bool)
    present[t] = True
    code = np.cumsum(present) - present
    sigma = int(present.sum())
    b = bits_for(sigma - 1 if sigma > 1 else 1)
    k = min(64 // b, (bits_for(n) + 10 + b - 1) // b)
    ns = min(k - 1, n)
    codes = code[t].astype(object)
    keys0 = []
    for i in range(n):
        x = 0
        for s in range(k):
            x = (x << b) | (int(codes[i + s]) if i + s < n else 0)
        keys0.append(x)
    init = [n - 1 - j if j < ns else j - ns for j in range(n)]
    order = sorted(range(n), key=lambda j: keys0[init[j]])
    sufx0 = [init[j] for j in order]
    short_from = n - ns
    SA = np.full(n, -1, np.int64)
    rank = [0] * n
    GS = [0] * (n + 2)
    GE = [0] * (n + 2)
    RHO = {}  # label -> (round, rho*)
    bag = []  # groups of (suffix, slot), members in slot order
    rng = np.random.default_rng(shuffle_seed) if shuffle_seed is not None else None

    def pick_label(s, e, avoid=0):
        """deterministic label for the group occupying slots [s, e] (0-based, size >= 2): a multiple
        of M for a huge group, an odd label for a tiny one (canonical: the first odd label of the
        range), any other even label for a medium one.  `avoid`: label of the huge group this one was
        split from (its table entry still belongs to the inert block)."""
        size = e - s + 1
        mid = s + (e - s) // 2 + 1  # 1-based
        if size >= T:
            lab = (mid // M) * M
            if lab < s + 1:
                lab += M
            if lab == avoid:
                lab = lab + M if lab + M <= e + 1 else lab - M
            assert s + 1 <= lab <= e + 1 and lab != avoid
            return lab
        if size <= TINY:
            lab = s + 1 if (s + 1) % 2 == 1 else s + 2
            assert s + 1 <= lab <= e + 1
            return lab
        lab = mid
        if TINY == 0:  # no bag (sparse mode): any label that is no multiple of M
            if lab % M == 0:
                lab = lab + 1 if lab + 1 <= e + 1 else lab - 1
            return lab
        if lab % 2 == 1:
            lab = lab + 1 if lab + 1 <= e + 1 else lab - 1
        if lab % M == 0:
            lab = lab + 2 if lab + 2 <= e + 1 else lab - 2
        assert s + 1 <= lab <= e + 1 and lab % 2 == 0 and lab % M != 0, (s, e, lab)
        return lab

    def resolve(w):
        """label word -> (current label, final?)  (sa_build: resolve_label)"""
        if w & DEAD:
            return w & LMASK, True
        lab = w & LMASK
        if lab % M != 0:
            return lab, False
        s, e = GS[lab], GE[lab]
        if s == e:
            return s + 1, True  # the group has shrunk to one suffix
        if e - s + 1 >= T and s + 1 <= lab <= e + 1:
            return lab, False
        return pick_label(s, e), False

    # ---- round 0 ---------------------------------------------------------------------------------
    def groups_of(flags, L):
        heads = [l for l in range(L) if flags[l]]
        return [(heads[j], (heads[j + 1] if j + 1 < len(heads) else L) - 1) for j in range(len(heads))]

    L = n
    flag = []
    for l in range(L):
        f = l == 0 or keys0[sufx0[l]] != keys0[sufx0[l - 1]]
        f = f or sufx0[l] >= short_from or (l > 0 and sufx0[l - 1] >= short_from)
        flag.append(f)
    for (a, z) in groups_of(flag, L):
        if a == z:
            SA[a] = sufx0[a]
            rank[sufx0[a]] = DEAD | (a + 1)
        else:
            lab = pick_label(a, z)
            GS[lab], GE[lab] = a, z
            for l in range(a, z + 1):
                rank[sufx0[l]] = lab
            if TINY and lab % 2 == 1:
                bag.append([(sufx0[l], l) for l in range(a, z + 1)])
    lst = [i for i in range(n) if not (rank[i] & DEAD) and not (TINY and rank[i] % 2 == 1)]
    h = k
    rnd = 0
    while lst or bag:
        rnd += 1
        assert rnd < 80
        if rng is not None:
            rng.shuffle(lst)
        # ---- K1 ------------------------------------------------------------------------------------
        new_lst, sort_in = [], []
        inert = 0
        updates = []  # table/rank writes are applied after the walk: every read sees the state left by the rebuild
        for i in lst:
            w = rank[i]
            if w & DEAD:
                continue
            if TINY and (w & LMASK) % 2 == 1:
                continue  # tiny group: lives in the bag now
            lab, fin = resolve(w)
            if fin:  # lazily finalised huge group of one
                s = lab - 1
                updates.append(("final", i, s))
                continue
            if lab != (w & LMASK):
                s, e = GS[w & LMASK], GE[w & LMASK]
                updates.append(("relabel", i, lab, s, e))
            tpos = i + h
            r2 = resolve(rank[tpos])[0] if tpos < n else 0
            new_lst.append(i)
            if lab % M == 0:
                if RHO.get(lab, (None, None))[0] != rnd:
                    RHO[lab] = (rnd, r2)
                if RHO[lab][1] == r2:
                    inert += 1
                    continue
            sort_in.append(((lab << 31) | r2, i))
        for u in updates:
            if u[0] == "final":
                assert SA[u[2]] == -1
                SA[u[2]] = u[1]
                rank[u[1]] = DEAD | (u[2] + 1)
            else:
                _, i, lab, s, e = u
                rank[i] = lab
                GS[lab], GE[lab] = s, e
        # ---- bag: refine the tiny groups in place (reads see the ranks left by the previous round) ----
        bag_updates, new_bag = [], []
        for grp in bag:
            gs = grp[0][1]
            keyed = []
            for (i, slot) in grp:
                tpos = i + h
                keyed.append(((resolve(rank[tpos])[0] if tpos < n else 0), i))
            keyed.sort(key=lambda kv: kv[0])
            a = 0
            while a < len(keyed):
                z = a
                while z + 1 < len(keyed) and keyed[z + 1][0] == keyed[a][0]:
                    z += 1
                s_, e_ = gs + a, gs + z
                if a == z:
                    bag_updates.append(("final", keyed[a][1], s_))
                else:
                    lab = pick_label(s_, e_)
                    assert lab % 2 == 1
                    for l in range(a, z + 1):
                        bag_updates.append(("label", keyed[l][1], lab))
                    new_bag.append([(keyed[l][1], gs + l) for l in range(a, z + 1)])
                a = z + 1
        # ---- sort + rebuild --------------------------------------------------------------------------
        sort_in.sort(key=lambda kv: kv[0])
        S = len(sort_in)
        key = [kv[0] for kv in sort_in]
        sfx = [kv[1] for kv in sort_in]
        runflag = [l == 0 or (key[l] >> 31) != (key[l - 1] >> 31) for l in range(S)]
        subflag = [l == 0 or key[l] != key[l - 1] for l in range(S)]
        slot = [0] * S
        has_rho = [False] * S
        for (a, z) in groups_of(runflag, S):
            lab = key[a] >> 31
            gs, ge = GS[lab], GE[lab]
            rho = RHO[lab][1] if (lab % M == 0 and RHO.get(lab, (None,))[0] == rnd) else None
            R = z - a + 1
            if rho is None:
                assert R == ge - gs + 1, (lab, R, gs, ge)
            cl = 0
            for l in range(a, z + 1):
                r2 = key[l] & LMASK
                less = rho is None or r2 < rho
                assert rho is None or r2 != rho
                if less:
                    assert l - a == cl
                    cl += 1
                    slot[l] = gs + (l - a)
                else:
                    slot[l] = ge - (z - l)
                has_rho[l] = rho is not None
            if rho is not None:  # the inert block keeps the label; its range shrinks from both sides
                GS[lab], GE[lab] = gs + cl, ge - (R - cl)
        moved = []
        for (a, z) in groups_of(subflag, S):
            s, e = slot[a], slot[z]
            assert e - s == z - a
            old = key[a] >> 31
            if a == z:
                assert SA[s] == -1
                SA[s] = sfx[a]
                rank[sfx[a]] = DEAD | (s + 1)
                continue
            size = e - s + 1
            medium_ok = old % M != 0 and (TINY == 0 or old % 2 == 0) and s + 1 <= old <= e + 1 and TINY < size < T
            if medium_ok:
                lab = old  # the group keeps its label: no rank writes
            else:
                lab = pick_label(s, e, avoid=old if has_rho[a] else 0)
                for l in range(a, z + 1):
                    rank[sfx[l]] = lab
            GS[lab], GE[lab] = s, e
            if TINY and lab % 2 == 1:
                moved.append([(sfx[l], slot[l]) for l in range(a, z + 1)])
        for u in bag_updates:
            if u[0] == "final":
                assert SA[u[2]] == -1
                SA[u[2]] = u[1]
                rank[u[1]] = DEAD | (u[2] + 1)
            else:
                rank[u[1]] = u[2]
        bag = new_bag + moved
        if stats is not None:
            stats.append((h, len(new_lst), S, inert, sum(len(g) for g in bag)))
        lst = new_lst
        h *= 2
    assert (SA >= 0).all(), np.flatnonzero(SA < 0)[:10]
    return SA.astype(np.int32)
