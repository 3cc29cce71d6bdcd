"""CPU: the reference's own JIT-compiled kernels (oracle/_ref/kernels, harvested by oracle/refbuild/build_ref_kernels.sh
from the unmodified reference in OpenMP mode) against the oracle C port on the same inputs.  Skipped where the
binaries are absent (they are build products of this container and travel with the snapshot, not with git)."""
import numpy as np
import pytest

from oracle import elliptic_ref as er
from oracle import ref_kernels as rk
from oracle.mesh_box import build_box_hex_mesh, masked_global_ids
from oracle.ogs_ref import SIGNED, ogs_setup_all


@pytest.mark.parametrize("N,n,lam", [(1, 5, 1.0), (2, 4, 0.0), (3, 3, 0.7), (4, 3, 1.0), (5, 2, 0.0), (6, 2, 1.0), (7, 3, 0.0),
                                     (7, 2, 1.3), (8, 2, 1.0)])
def test_reference_kernels_match_oracle_port(N, n, lam):
    if not rk.available(N):
        pytest.skip("oracle/_ref/kernels not built (oracle/refbuild/build_ref_kernels.sh)")
    m = build_box_hex_mesh(N, n, n, n)
    _, ids = masked_global_ids(m)
    o = ogs_setup_all([ids], SIGNED, True)[0]
    G2L = o.global_to_local()
    rs, ci = o.gatherLocal.rowStartsT, o.gatherLocal.colIdsT
    q = er.splitmix_uniform(77 + N, o.Ngather)
    ref = rk.RefOperator(N + 1, G2L, m.wJ, m.ggeo, m.D, lam, rs, ci)(q)
    port = er.operator(N + 1, G2L, m.wJ, m.ggeo, m.D, lam, rs, ci, q)
    assert np.abs(ref - port).max() <= 1e-13 * np.abs(port).max()


def test_row_blocks_follow_the_reference_rule():
    rng = np.random.default_rng(1)
    sizes = rng.integers(1, 9, size=5000)
    rs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    b = rk.row_blocks(rs)
    assert b[0] == 0 and b[-1] == sizes.size and np.all(np.diff(b) > 0)
    per_block = rs[b[1:]] - rs[b[:-1]]
    assert per_block.max() <= rk.GATHER_NODES_PER_BLOCK
    # greedy: adding the next row to any block (but the last) would overflow it
    nxt = sizes[b[1:-1]]
    assert np.all(per_block[:-1] + nxt > rk.GATHER_NODES_PER_BLOCK)


@pytest.mark.parametrize("NqF,NqC", [(8, 6), (6, 4), (4, 3), (3, 2), (9, 7), (7, 5), (5, 3)])
def test_reference_transfer_kernels_match_oracle(NqF, NqC):
    """p-multigrid coarsen / prolongate: the oracle's tensor-product restatement against the reference's kernels for
    every (fine, coarse) pair of the HALFDOFS ladders of N = 7 and N = 8."""
    if not rk.transfer_available(NqF, NqC):
        pytest.skip("oracle/_ref/kernels not built (oracle/refbuild/build_ref_kernels.sh)")
    from oracle.mesh_box import degree_raise_1d
    n = 2
    P = degree_raise_1d(NqC - 1, NqF - 1)
    G2L = {}
    Ng = {}
    for Nq in (NqF, NqC):
        m = build_box_hex_mesh(Nq - 1, n, n, n, geometry=False)
        o = ogs_setup_all([masked_global_ids(m)[1]], SIGNED, True)[0]
        G2L[Nq], Ng[Nq] = o.global_to_local(), o.Ngather
    qf = er.splitmix_uniform(5, Ng[NqF])
    qc = er.splitmix_uniform(6, Ng[NqC])
    loc = lambda G, q: np.where(G >= 0, q[np.maximum(G, 0)], 0.0)
    ref_c = rk.coarsen(NqF, NqC, G2L[NqF], P, qf)
    got_c = er.coarsen_hex3d(NqF, NqC, P, loc(G2L[NqF], qf))
    assert np.abs(ref_c - got_c).max() <= 1e-13 * np.abs(ref_c).max()
    ref_p = rk.prolongate(NqF, NqC, G2L[NqC], P, qc)
    got_p = er.prolongate_hex3d(NqF, NqC, P, loc(G2L[NqC], qc))
    assert np.abs(ref_p - got_p).max() <= 1e-13 * np.abs(ref_p).max()


def test_reference_ogs_kernels_match_oracle():
    if not rk.available(3):
        pytest.skip("oracle/_ref/kernels not built (oracle/refbuild/build_ref_kernels.sh)")
    from oracle import ogs_ref
    m = build_box_hex_mesh(3, 3, 3, 3, geometry=False)
    o = ogs_setup_all([masked_global_ids(m)[1]], SIGNED, True)[0]
    op = o.gatherLocal
    nloc = m.Nelements * m.Np
    v = er.splitmix_uniform(9, nloc)
    gv = er.splitmix_uniform(10, o.Ngather)
    # ogsOperator_t::Scatter: Trans scatters to the owner copies (N maps), NoTrans to every copy (T maps)
    assert np.array_equal(rk.ogs_scatter(op.rowStartsN, op.colIdsN, gv, nloc), ogs_ref.op_scatter(op, gv, np.zeros(nloc), "Trans"))
    assert np.array_equal(rk.ogs_scatter(op.rowStartsT, op.colIdsT, gv, nloc), ogs_ref.op_scatter(op, gv, np.zeros(nloc), "NoTrans"))
    # symmetric gather-scatter: every copy gets the left-to-right sum of its row
    g = ogs_ref.op_gather(op, v, "Trans")
    want = ogs_ref.op_scatter(op, g, v.copy(), "NoTrans")
    assert np.array_equal(rk.ogs_gather_scatter(op.rowStartsT, op.colIdsT, v), want)
