"""CPU pin of the arithmetic rule behind csrc/tb_cdf.cu (TEST INFRASTRUCTURE): inside one binade the sequential fp64
cumulative sum is an integer prefix sum,  fl(s + p) = (S + inc(p)) q  with  inc(p) = floor(p/q) + [frac > 1/2]
(round-half ties excluded), as long as the result stays below 2^(E+1).  numpy restatement of inc_of() -- the magic-number
rounding the kernels use -- checked against numpy's own left-to-right cumsum."""
import numpy as np

MAGIC = 6755399441055744.0          # 2^52 + 2^51


def inc_of(p, E):
    """int64 increments of the elements p under binade E, -1 where the integer rule does not apply (tb_cdf.cu inc_of)."""
    up, top = np.ldexp(1.0, 52 - E), np.ldexp(1.0, E + 1)
    sc = p * up
    t = sc + MAGIC
    r = t - MAGIC
    ok = (p >= 0.0) & (p < 0.25 * top) & (np.abs(sc - r) != 0.5)
    inc = t.view(np.int64) - np.float64(MAGIC).view(np.int64)
    return np.where(ok, inc, -1)


def test_single_add_follows_the_integer_rule():
    rng = np.random.default_rng(1)
    for E in (-300, -40, -1, 0, 7, 200):
        q = np.ldexp(1.0, E - 52)
        S = rng.integers(1 << 52, 1 << 53, size=20000, dtype=np.int64)
        s = S.astype(np.float64) * q                                  # exact: s lies in [2^E, 2^(E+1))
        p = np.ldexp(rng.random(20000), E - rng.integers(0, 60, 20000))
        inc = inc_of(p, E)
        good = (inc >= 0) & (S + inc < (1 << 53))
        assert good.mean() > 0.5
        want = s[good] + p[good]                                      # IEEE round-to-nearest-even add
        got = (S[good] + inc[good]).astype(np.float64) * q
        assert np.array_equal(want.view(np.uint64), got.view(np.uint64))


def test_ties_are_excluded_and_decided_by_parity():
    E = 0
    q = np.ldexp(1.0, E - 52)
    p = np.array([0.5 * q, 1.5 * q, 2.5 * q, 0.25 * q, 0.75 * q])
    inc = inc_of(p, E)
    assert list(inc[:3]) == [-1, -1, -1] and list(inc[3:]) == [0, 1]
    # the literal add the kernels fall back to: even S rounds the half down, odd S rounds it up
    assert 1.0 + 0.5 * q == 1.0 and (1.0 + q) + 0.5 * q == 1.0 + 2 * q


def test_sequential_cumsum_inside_a_binade_is_an_integer_prefix_sum():
    rng = np.random.default_rng(2)
    for trial in range(20):
        E = int(rng.integers(-60, 10))
        q = np.ldexp(1.0, E - 52)
        s0 = np.ldexp(1.0 + 0.3 * rng.random(), E)                   # start low in the binade: room for the tile
        p = np.ldexp(rng.random(4096), E - 14 - rng.integers(0, 30, 4096))
        p[rng.integers(0, 4096, 100)] = 0.0
        inc = inc_of(p, E)
        assert (inc >= 0).all()
        S0 = int(s0 / q)
        assert S0 * q == s0
        prefix = S0 + np.cumsum(inc)
        assert prefix[-1] < (1 << 53)
        got = prefix.astype(np.float64) * q
        want = np.cumsum(np.concatenate([[s0], p]))[1:]               # numpy: strictly left to right
        assert np.array_equal(want.view(np.uint64), got.view(np.uint64))
        # and the order in which the integer increments are summed does not matter (what the parallel kernels use)
        perm = rng.permutation(4096)
        assert S0 + int(inc[perm].sum()) == int(prefix[-1])


def test_crossing_needs_the_literal_add():
    """Past 2^(E+1) the ulp doubles: the kernels stop the integer prefix at the first element that would reach 2^53 and
    add that element literally (then continue under the new binade)."""
    E = 0
    q = np.ldexp(1.0, E - 52)
    s0 = 2.0 - 8 * q
    p = np.array([3 * q, 3 * q, 3.25 * q, 5.75 * q, 0.75 * q])
    inc = inc_of(p, E)
    assert (inc >= 0).all()
    S0 = int(s0 / q)
    pref = S0 + np.cumsum(inc)
    first = int(np.argmax(pref >= (1 << 53)))
    assert first == 2
    s = s0
    ref = np.cumsum(np.concatenate([[s0], p]))[1:]
    out = []
    for j in range(first):
        out.append(float(pref[j]) * q)
    s = out[-1] + p[first]                                            # literal add of the crossing element
    out.append(s)
    E2 = 1
    q2 = np.ldexp(1.0, E2 - 52)
    inc2 = inc_of(p[first + 1:], E2)
    assert (inc2 >= 0).all()
    S2 = int(s / q2)
    out.extend(((S2 + np.cumsum(inc2)).astype(np.float64) * q2).tolist())
    assert np.array_equal(np.array(out).view(np.uint64), ref.view(np.uint64))
