"""Property tests (hypothesis) of the host-side decomposition logic behind the C ABI -- the pieces that replace MPI_Dims_create,
decompose_1d, MPI_Cart_shift / MPI_Cart_find_corners and the derived-datatype message lists (L3/main.f90:24-72,144-212;
2d_revised/mpi_blocked/main.f90:24-60).  No GPU needed: these entry points are host-only."""
import numpy as np
from hypothesis import given, settings, strategies as st

import mglc_b200 as mg
from oracle import oracle as orc


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 5000), st.integers(1, 64))
def test_decompose_1d_tiles_the_axis(total, nranks):
    """blocks are contiguous, cover 0..total-1 exactly once, sizes differ by at most one with the larger ones first"""
    if total < nranks:
        return
    blocks = [mg.decompose_1d(total, r, nranks) for r in range(nranks)]
    pos = 0
    for n, s in blocks:
        assert s == pos and n in (total // nranks, total // nranks + 1)
        pos += n
    assert pos == total
    sizes = [n for n, _ in blocks]
    assert sizes == sorted(sizes, reverse=True)


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 512))
def test_dims_create_is_a_balanced_nonincreasing_factorisation(nranks):
    for ndim in (2, 3):
        d = mg.dims_create_nd(nranks, ndim)[:ndim]
        assert int(np.prod(d)) == nranks and list(d) == sorted(d, reverse=True)
    assert mg.dims_create(nranks) == mg.dims_create_nd(nranks, 3)
    # the cases the reference runs: 2 -> 2x1x1, 4 -> 2x2x1, 8 -> 2x2x2
    assert mg.dims_create(2) == (2, 1, 1) and mg.dims_create(4) == (2, 2, 1) and mg.dims_create(8) == (2, 2, 2)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 4), st.integers(1, 4), st.integers(1, 3), st.integers(4, 40), st.integers(4, 40), st.integers(3, 20))
def test_halo_plan_3d_is_pairwise_consistent(d0, d1, d2, nx, ny, nz):
    """what rank A sends to B in message m is exactly what B expects from A in message m; volumes are the reference's"""
    dims, total, P = (d0, d1, d2), (nx, ny, nz), d0 * d1 * d2
    if any(t < d for t, d in zip(total, dims)):
        return
    plans = [mg.halo_plan(mg.make_desc(total, P, r, dims)) for r in range(P)]
    for r, plan in enumerate(plans):
        assert [m["dir"] for m in plan] == [0, 1, 2, 3, 4, 5, 7, 10, 9, 8, 11, 14, 13, 12, 15, 18, 17, 16]     # ex_sendrecv.f90 order
        for k, m in enumerate(plan):
            if m["send_to"] >= 0:
                peer = plans[m["send_to"]][k]
                assert peer["recv_from"] == r and peer["recv_count"] == m["send_count"] > 0
            else:
                assert m["send_count"] == 0
            if m["recv_from"] >= 0:
                peer = plans[m["recv_from"]][k]
                assert peer["send_to"] == r and peer["send_count"] == m["recv_count"] > 0
            else:
                assert m["recv_count"] == 0
            assert m["npop"] == (5 if m["dir"] < 6 else 1) and len(m["pops"]) == m["npop"]


@settings(max_examples=80, deadline=None)
@given(st.integers(1, 5), st.integers(1, 5), st.integers(5, 60), st.integers(5, 60))
def test_halo_plan_2d_is_pairwise_consistent_and_matches_the_oracle_topology(d0, d1, nx, ny):
    dims, total, P = (d0, d1), (nx, ny), d0 * d1
    if nx < d0 or ny < d1:
        return
    plans = [mg.halo_plan_2d(total, dims, r) for r in range(P)]
    wd = orc.Lid2DWorld(total, P, dims)
    for r, plan in enumerate(plans):
        R = wd.ranks[r]
        assert tuple(m["send_to"] for m in plan[:8]) == R.nbr + R.cnr
        assert [m["send_to"] for m in plan[8:]] == list(R.nbr)
        for k, m in enumerate(plan):
            for a, b, ca, cb in (("send_to", "recv_from", "send_count", "recv_count"), ("recv_from", "send_to", "recv_count", "send_count")):
                if m[a] >= 0:
                    peer = plans[m[a]][k]
                    assert peer[b] == r and peer[cb] == m[ca] > 0
                else:
                    assert m[ca] == 0
        n = R.n
        want = [3 * n[1], 3 * n[1], 3 * n[0], 3 * n[0], 1, 1, 1, 1, n[1], n[1], n[0], n[0]]
        assert [m["send_count"] for m in plan] == [w if m["send_to"] >= 0 else 0 for w, m in zip(want, plan)]
    wd.close()


@settings(max_examples=150, deadline=None)
@given(st.integers(1, 6), st.integers(1, 6), st.integers(6, 300), st.integers(6, 300))
def test_the_cuda_drivers_own_message_tables_equal_the_published_plan(d0, d1, nx, ny):
    """mglc_l2d_create / mglc_t2d_create refuse to start when their message table differs from mglc_halo_plan_2d; the same
    host code that builds those tables is reachable without a GPU (mglc_l2d_msg_table / mglc_t2d_msg_table), so the CPU suite
    can show that the refusal can never trigger"""
    import ctypes as C
    from mglc_b200 import _lib as L
    lib = L.lib()
    dims = (C.c_int * 2)(d0, d1)
    for rank in range(d0 * d1):
        plan = mg.halo_plan_2d((nx, ny), (d0, d1), rank)
        for fn, n in ((lib.mglc_l2d_msg_table, 8), (lib.mglc_t2d_msg_table, 12)):
            out = (L.HaloMsg * 12)()
            L.check(fn(nx, ny, dims, rank, out))
            for k in range(n):
                m, q = out[k], plan[k]
                assert (m.dir, m.send_to, m.recv_from, m.send_count, m.recv_count, m.npop) == \
                       (q["dir"], q["send_to"], q["recv_from"], q["send_count"], q["recv_count"], q["npop"]), (fn.__name__, rank, k)
