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
