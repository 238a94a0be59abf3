"""CPU: the fast MultiCounter (ultranest_b200.netiter) is the reference's
MultiCounter -- its own `passing_node` running on a count-friendly `rootids` array: the same seeded
run, every attribute after every node bit for bit, at a fraction of the time."""
import time

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built")


def _tree(nlive, niter, seed):
    """A nested-sampling tree like a run leaves it: nlive roots, every dead point replaced by
    one child of a random live lineage, plus a few multi-child and childless nodes."""
    oracle.reference()
    from ultranest.netiter import TreeNode, PointPile
    rng = np.random.RandomState(seed)
    pile = PointPile(1, 1)
    root = TreeNode(id=-1, value=-np.inf)
    live = []
    for i in range(nlive):
        L = rng.normal()
        node = pile.make_node(L, [L], [L])
        root.children.append(node)
        live.append(node)
    for it in range(niter):
        live.sort(key=lambda n: n.value)
        worst = live.pop(0)
        nchild = 1 if rng.uniform() < 0.9 else (2 if rng.uniform() < 0.7 else 0)
        for _ in range(nchild):
            L = worst.value + rng.exponential(0.01)
            child = pile.make_node(L, [L], [L])
            worst.children.append(child)
            live.append(child)
        if not live:
            break
    return root, pile


def _walk(counter_cls, root, random, check, seed):
    from ultranest.netiter import BreadthFirstIterator
    np.random.seed(seed)
    roots = root.children
    explorer = BreadthFirstIterator(roots)
    it = counter_cls(nroots=len(roots), nbootstraps=30, random=random, check_insertion_order=check)
    trace = []
    t = 0.0
    while True:
        nxt = explorer.next_node()
        if nxt is None:
            break
        rootid, node, (_, active_rootids, active_values, _) = nxt
        t0 = time.perf_counter()
        it.passing_node(rootid, node, active_rootids, active_values)
        t += time.perf_counter() - t0
        trace.append((it.logZ, it.logZerr, it.logVolremaining, it.logZremain, it.remainder_ratio,
                      it.remainder_fraction, it.all_logZ.copy(), it.all_H.copy(),
                      it.all_logVolremaining.copy(), it.all_logZremain.copy()))
        explorer.expand_children_of(rootid, node)
    return it, trace, t


@pytest.mark.parametrize("random,check", [(False, False), (True, True)])
def test_fast_multicounter_is_the_reference(random, check):
    oracle.reference()
    import ultranest.netiter as netiter
    from ultranest_b200 import netiter as fastmod
    root, _ = _tree(600, 2500, 3)
    ref_cls = netiter.MultiCounter
    a, tr_a, t_ref = _walk(ref_cls, root, random, check, 11)
    fast_cls = fastmod.install()
    try:
        assert netiter.MultiCounter is fast_cls and issubclass(fast_cls, ref_cls)
        b, tr_b, t_fast = _walk(fast_cls, root, random, check, 11)
    finally:
        fastmod.uninstall()
    assert netiter.MultiCounter is ref_cls
    assert len(tr_a) == len(tr_b) > 2000
    for sa, sb in zip(tr_a, tr_b):
        for xa, xb in zip(sa, sb):
            assert np.array_equal(np.asarray(xa), np.asarray(xb), equal_nan=True)
    assert np.array_equal(np.asarray(a.logweights), np.asarray(b.logweights))
    assert a.istail == b.istail and a.insertion_order_runs == b.insertion_order_runs
    print("passing_node: reference %.3fs fast %.3fs" % (t_ref, t_fast))


def test_lazy_column_count_handles_duplicates_bad_ids_and_other_uses():
    oracle.reference()
    from ultranest_b200 import netiter as fastmod
    fast_cls = fastmod.install()
    try:
        np.random.seed(1)
        it = fast_cls(nroots=300, nbootstraps=7)
        plain = np.array(it.rootids)                       # an ordinary copy of the masks
        assert type(plain) is np.ndarray and plain.dtype == bool
        rng = np.random.RandomState(2)
        ids = rng.randint(300, size=500)                   # duplicates: arcs of the same root
        got = it.rootids[:, ids].sum(axis=1)
        want = plain[:, ids].sum(axis=1)
        assert (got == want).all() and got.dtype == want.dtype
        # every other use of the lazy object behaves like the gathered array
        lazy = it.rootids[:, ids]
        assert (np.asarray(lazy) == plain[:, ids]).all()
        assert lazy.shape == plain[:, ids].shape and len(lazy) == len(plain)
        assert (lazy.sum(axis=0) == plain[:, ids].sum(axis=0)).all()
        assert lazy.sum() == plain[:, ids].sum()
        assert (lazy[0] == plain[0, ids]).all()
        # short index lists, scalar columns, row picks and negative ids take NumPy's own path
        assert (it.rootids[:, ids[:5]] == plain[:, ids[:5]]).all()
        assert (it.rootids[:, 3] == plain[:, 3]).all()
        assert (it.rootids[0, ids] == plain[0, ids]).all()
        neg = ids.copy()
        neg[0] = -1
        assert (it.rootids[:, neg].sum(axis=1) == plain[:, neg].sum(axis=1)).all()
        bad = ids.copy()
        bad[10] = 300
        with pytest.raises(IndexError):
            it.rootids[:, bad].sum(axis=1)
        assert len(it.rootids) == 8
    finally:
        fastmod.uninstall()
