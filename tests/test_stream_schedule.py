"""Host-side model of the row-streaming NA kernels' schedule (lm-net_b200/csrc/na2d_stream.cuh): stripe / band
boundaries, query / key halos and — the subtle part — the shared-memory ring depths under the mbarrier hand-off,
where a fast thread may run one phase ahead of a slow one.  The index rules are restated from the kernel source
(axis_window, inverse_window, stream_boundary, the staging order) and checked exhaustively over axis lengths that
the sampled GPU parity shapes cannot cover.  CPU only."""
import pytest


def axis_start(t, L, K):
    return min(max(t - K // 2, 0), L - K)


def inv_window(t, L, K):
    return (0 if t <= K - 1 else t - K // 2), (L - 1 if t >= L - K else t + K // 2)


def boundary(s, T, L, K):
    b = s * T
    if b >= L:
        return L
    if L - b < K:
        return L - K
    return b


def attends(q, key, L, K):
    s = axis_start(q, L, K)
    return s <= key < s + K


@pytest.mark.parametrize("K", [3, 5, 7])
def test_inverse_window_is_exactly_the_attending_queries(K):
    for L in range(K, 60):
        for key in range(L):
            lo, hi = inv_window(key, L, K)
            assert [q for q in range(L) if attends(q, key, L, K)] == list(range(lo, hi + 1))
            if L >= 2 * K:                            # eligibility of the streaming backward (make_plan)
                assert hi - lo + 1 <= K + K // 2      # what its rings are sized for


@pytest.mark.parametrize("K", [3, 5])
def test_stripes_partition_the_axis_and_halos_fit(K):
    NS = K // 2
    for L in range(2 * K, 140):
        for T in range(K, 70):
            n = -(-L // T)
            cuts = [boundary(s, T, L, K) for s in range(n + 1)]
            assert cuts[0] == 0 and cuts[-1] == L and all(a <= b for a, b in zip(cuts, cuts[1:]))
            for c0, c1 in zip(cuts, cuts[1:]):
                if c0 == c1:
                    continue
                assert c1 - c0 <= T
                ql, qh = inv_window(c0, L, K)[0], inv_window(c1 - 1, L, K)[1]
                assert qh - ql + 1 <= T + 2 * NS                                   # query columns per CTA (QW)
                kv_lo, kv_hi = axis_start(ql, L, K), axis_start(qh, L, K) + K - 1
                assert kv_hi - kv_lo + 1 <= T + 4 * NS                             # key/value columns (QW + 2*NS)
                for key in range(c0, c1):                                          # every attending query is staged
                    lo, hi = inv_window(key, L, K)
                    assert ql <= lo and hi <= qh


def _bwd_phase_b_rows(tb, L, K, r0, r1):
    """key rows finalised after query row tb (their last attending query row is tb), clipped to the band"""
    NS = K // 2
    if tb == L - 1:
        lo, hi = L - K, L - 1
    else:
        lo = hi = tb - NS
        if lo < 0 or lo >= L - K:
            lo, hi = 0, -1
    return range(max(lo, r0), min(hi, r1 - 1) + 1)


@pytest.mark.parametrize("K", [3])
def test_backward_rings_survive_one_phase_of_skew(K):
    """Thread timeline per step t: wait(t); stage(q row t+1, kv rows for A(t+1)); A(t) [writes stats t]; arrive(t);
    B(t-1).  A fast thread may already have done stage + A of step t+1 while a slow one still reads in B(t-1)."""
    NS, RING = K // 2, K + K // 2 + 3
    for L in range(2 * K, 90):
        for RB in (K, 8, 11, 40):
            n = -(-L // RB)
            cuts = [boundary(s, RB, L, K) for s in range(n + 1)]
            for r0, r1 in zip(cuts, cuts[1:]):
                if r0 == r1:
                    continue
                tq_lo, tq_hi = inv_window(r0, L, K)[0], inv_window(r1 - 1, L, K)[1]
                q_slot, kv_slot, st_slot = {}, {}, {}           # ring slot -> row currently held
                kv_next = axis_start(tq_lo, L, K)

                def stage_kv_until(need):
                    nonlocal kv_next
                    while kv_next < need:
                        kv_slot[kv_next % RING] = kv_next
                        kv_next += 1

                def writes_of_step(t):                           # staging at the top of step t, then A(t)'s stats
                    if t + 1 <= tq_hi:
                        q_slot[(t + 1) % RING] = t + 1
                        stage_kv_until(axis_start(t + 1, L, K) + K)
                    if t <= tq_hi:
                        st_slot[t % RING] = t

                q_slot[tq_lo % RING] = tq_lo
                stage_kv_until(axis_start(tq_lo, L, K) + K)
                done_keys = set()
                writes_of_step(tq_lo)
                for t in range(tq_lo, tq_hi + 2):
                    if t <= tq_hi:                               # phase A(t) reads (own step's writes are in place)
                        assert q_slot[t % RING] == t
                        for r in range(axis_start(t, L, K), axis_start(t, L, K) + K):
                            assert kv_slot[r % RING] == r
                    writes_of_step(t + 1)                        # the fast thread's next step, BEFORE the slow B(t-1)
                    for r in _bwd_phase_b_rows(t - 1, L, K, r0, r1) if t - 1 >= tq_lo else ():
                        assert kv_slot[r % RING] == r
                        lo, hi = inv_window(r, L, K)
                        assert hi == t - 1
                        for tq in range(lo, hi + 1):
                            assert q_slot[tq % RING] == tq and st_slot[tq % RING] == tq
                        done_keys.add(r)
                assert done_keys == set(range(r0, r1))           # every key row of the band is finalised once


@pytest.mark.parametrize("K,RS", [(3, 4), (3, 2), (5, 2), (7, 2)])
def test_forward_rings_survive_one_group_of_skew(K, RS):
    """Forward: groups of RS query rows; a thread arrives for group g+1 one row into group g and may then run ahead
    into group g+1, staging group g+2, while a slow thread still reads the rows of group g."""
    RING = K - 1 + 3 * RS
    for L in range(K, 100):
        for RB in (1, RS, 8, 13, L):
            for r0 in range(0, L, RB):
                r1 = min(r0 + RB, L)
                slot = {}
                kv_next = axis_start(r0, L, K)

                def stage_until(last_row):
                    nonlocal kv_next
                    need = axis_start(last_row, L, K) + K
                    while kv_next < need:
                        slot[kv_next % RING] = kv_next
                        kv_next += 1

                groups = [(t0, min(t0 + RS, r1)) for t0 in range(r0, r1, RS)]
                stage_until(groups[0][1] - 1)
                for gi, (t0, t1) in enumerate(groups):
                    if gi + 1 < len(groups):
                        stage_until(groups[gi + 1][1] - 1)       # own staging at the top of the group
                    if gi + 2 < len(groups):
                        stage_until(groups[gi + 2][1] - 1)       # a fast thread, one group ahead, stages this too
                    for t in range(t0, t1):
                        for r in range(axis_start(t, L, K), axis_start(t, L, K) + K):
                            assert slot[r % RING] == r


def test_model_matches_the_kernel_source_constants():
    """The ring depths and the boundary rule modelled above are the ones compiled into the kernels."""
    import os

    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lm-net_b200", "csrc",
                            "na2d_stream.cuh")).read()
    assert "static constexpr int RING = KT + KT / 2 + 3;" in src                     # backward rings
    assert "static constexpr int RING = KT - 1 + 3 * RS;" in src                     # forward rings
    assert "return (K == 3 && D < 8) ? 4 : 2;" in src                                # forward rows per group
    assert "if (L - b < K) return L - K;" in src                                     # stream_boundary
