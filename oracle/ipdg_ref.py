"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the interior-penalty DG operator on hexahedra, numpy.

Follows the reference's IPDG path (DISCRETIZATION = IPDG):
  gradient             solvers/elliptic/okl/ellipticGradientHex3D.okl:29-95     (q, dq/dx, dq/dy, dq/dz per node)
  surface + volume     solvers/elliptic/okl/ellipticAxIpdgHex3D.okl:35-85 (surfaceTerms), :359-642
  operator             solvers/elliptic/src/ellipticOperator.cpp:108-160  (gradient, trace halo, two element lists)
  diagonal             solvers/elliptic/src/ellipticBuildOperatorDiagonal.cpp:868-996
  surface factors      libs/mesh/meshSurfaceGeometricFactorsHex3D.cpp:31-195
  penalty              solvers/elliptic/src/ellipticSetup.cpp:66-75   (tau = 2 (N+1)(N+3) on hexahedra)

Pinned against dumps of the unmodified reference (tests/golden/ipdg_*.npz, oracle/refbuild/dump_ipdg_driver.cpp) by
tests/test_oracle_ipdg_cpu.py.  Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this.

Array conventions (the reference's): vgeo [E][Nvgeo=12][Np] with RX,RY,RZ,SX,SY,SZ,TX,TY,TZ,J,JW,IJW = 0..11;
sgeo [E][Nfaces*Nfp][Nsgeo=8] with NX,NY,NZ,SJ,IJ,IH,WSJ,WIJ = 0..7; vmapM / vmapP [E][6*Nq^2] node indices into the
(element + halo element) node array; EToB [E][6] boundary TYPE (1 Dirichlet, 2 Neumann, <= 0 interior)."""
import numpy as np

RX, RY, RZ, SX, SY, SZ, TX, TY, TZ, JID, JWID, IJWID = range(12)
NX, NY, NZ, SJ, IJ, IH, WSJ, WIJ = range(8)


def tau_hex(N):
    return 2.0 * (N + 1) * (N + 3)


def face_nodes(Nq):
    """volume node of face node n on face f (libs/mesh/meshReferenceNodesHex3D / faceNodes): faces 0..5 =
    t=-1, s=-1, r=+1, s=+1, r=-1, t=+1"""
    i = np.arange(Nq)
    a, b = np.meshgrid(i, i, indexing="xy")  # a fast
    a, b = a.reshape(-1), b.reshape(-1)
    N = Nq - 1
    f = np.empty((6, Nq * Nq), dtype=np.int64)
    f[0] = a + b * Nq
    f[1] = a + b * Nq * Nq
    f[2] = N + a * Nq + b * Nq * Nq
    f[3] = a + N * Nq + b * Nq * Nq
    f[4] = a * Nq + b * Nq * Nq
    f[5] = a + b * Nq + N * Nq * Nq
    return f


def gradient(Nq, vgeo, D, q):
    """[E, Np, 4] = (dq/dx, dq/dy, dq/dz, q)"""
    Np = Nq ** 3
    E = q.size // Np
    v = np.asarray(vgeo, dtype=np.float64).reshape(E, 12, Nq, Nq, Nq)
    u = np.asarray(q, dtype=np.float64).reshape(E, Nq, Nq, Nq)  # [e,k,j,i]
    D = np.asarray(D, dtype=np.float64).reshape(Nq, Nq)
    qr = np.einsum("in,ekjn->ekji", D, u)
    qs = np.einsum("jn,ekni->ekji", D, u)
    qt = np.einsum("kn,enji->ekji", D, u)
    g = np.empty((E, Nq, Nq, Nq, 4))
    g[..., 0] = v[:, RX] * qr + v[:, SX] * qs + v[:, TX] * qt
    g[..., 1] = v[:, RY] * qr + v[:, SY] * qs + v[:, TY] * qt
    g[..., 2] = v[:, RZ] * qr + v[:, SZ] * qs + v[:, TZ] * qt
    g[..., 3] = u
    return g.reshape(E, Np, 4)


def ax_ipdg(Nq, vgeo, sgeo, vmapM, vmapP, EToB, D, lam, tau, grad, Nelements=None):
    """Aq [E*Np] from grad [(E + halo elements)*Np, 4] (halo part already exchanged)"""
    Np, Nfp = Nq ** 3, Nq * Nq
    E = np.asarray(vmapM).size // (6 * Nfp) if Nelements is None else Nelements
    v = np.asarray(vgeo, dtype=np.float64).reshape(E, 12, Np)
    sg = np.asarray(sgeo, dtype=np.float64).reshape(E, 6 * Nfp, 8)
    vM = np.asarray(vmapM).reshape(E, 6 * Nfp)
    vP = np.asarray(vmapP).reshape(E, 6 * Nfp)
    bc = np.repeat(np.asarray(EToB).reshape(E, 6), Nfp, axis=1)
    g = np.asarray(grad, dtype=np.float64).reshape(-1, 4)
    D = np.asarray(D, dtype=np.float64).reshape(Nq, Nq)
    gM, gP = g[vM], g[vP].copy()                      # [E, 6*Nfp, 4]
    # homogeneous boundary data, then the ghost state 2*bc - interior (ellipticAxIpdgHex3D.okl:62-68)
    dirichlet, neumann = bc == 1, bc == 2
    gP[dirichlet, 3] = -gM[dirichlet, 3]
    gP[dirichlet, :3] = gM[dirichlet, :3]
    gP[neumann, 3] = gM[neumann, 3]
    gP[neumann, :3] = -gM[neumann, :3]
    dq = gP[..., 3] - gM[..., 3]
    w = sg[..., WSJ]
    n = sg[..., NX:NZ + 1]
    JW = v[:, JWID]
    gl = g[: E * Np].reshape(E, Np, 4)
    gx = JW[..., None] * gl[..., :3]                  # JW * grad q
    Aq = JW * lam * gl[..., 3]
    fn = face_nodes(Nq).reshape(-1)                   # [6*Nfp] volume node of every face node
    flux = 0.5 * w[..., None] * n * dq[..., None]
    pen = -0.5 * w * (np.einsum("efd,efd->ef", n, gP[..., :3] + gM[..., :3]) + tau * sg[..., IH] * dq)
    for d in range(3):
        np.add.at(gx[..., d], (np.arange(E)[:, None], fn[None, :]), flux[..., d])
    np.add.at(Aq, (np.arange(E)[:, None], fn[None, :]), pen)
    Gr = v[:, RX] * gx[..., 0] + v[:, RY] * gx[..., 1] + v[:, RZ] * gx[..., 2]
    Gs = v[:, SX] * gx[..., 0] + v[:, SY] * gx[..., 1] + v[:, SZ] * gx[..., 2]
    Gt = v[:, TX] * gx[..., 0] + v[:, TY] * gx[..., 1] + v[:, TZ] * gx[..., 2]
    sh = (E, Nq, Nq, Nq)
    Aq = Aq.reshape(sh)
    Aq = Aq + np.einsum("ni,ekjn->ekji", D, Gr.reshape(sh)) + np.einsum("nj,ekni->ekji", D, Gs.reshape(sh)) \
        + np.einsum("nk,enji->ekji", D, Gt.reshape(sh))
    return Aq.reshape(-1)


def operator_single_rank(Nq, vgeo, sgeo, vmapM, vmapP, EToB, D, lam, tau, q):
    return ax_ipdg(Nq, vgeo, sgeo, vmapM, vmapP, EToB, D, lam, tau, gradient(Nq, vgeo, D, q))


def diagonal(Nq, vgeo, sgeo, EToB, D, lam, tau):
    """BuildOperatorDiagonalIpdgHex3D: a Lagrange function is 1 at its own node, so the volume term collapses to the
    lines through the node and the face terms to the faces the node lies on."""
    Np, Nfp = Nq ** 3, Nq * Nq
    v = np.asarray(vgeo, dtype=np.float64)
    E = v.size // (12 * Np)
    v = v.reshape(E, 12, Nq, Nq, Nq)
    sg = np.asarray(sgeo, dtype=np.float64).reshape(E, 6 * Nfp, 8)
    D = np.asarray(D, dtype=np.float64).reshape(Nq, Nq)
    JW = v[:, JWID]
    Grr = JW * (v[:, RX] ** 2 + v[:, RY] ** 2 + v[:, RZ] ** 2)
    Gss = JW * (v[:, SX] ** 2 + v[:, SY] ** 2 + v[:, SZ] ** 2)
    Gtt = JW * (v[:, TX] ** 2 + v[:, TY] ** 2 + v[:, TZ] ** 2)
    Grs = JW * (v[:, RX] * v[:, SX] + v[:, RY] * v[:, SY] + v[:, RZ] * v[:, SZ])
    Grt = JW * (v[:, RX] * v[:, TX] + v[:, RY] * v[:, TY] + v[:, RZ] * v[:, TZ])
    Gst = JW * (v[:, SX] * v[:, TX] + v[:, SY] * v[:, TY] + v[:, SZ] * v[:, TZ])
    D2 = D * D
    A = np.einsum("mi,ekjm->ekji", D2, Grr) + np.einsum("mj,ekmi->ekji", D2, Gss) + np.einsum("mk,emji->ekji", D2, Gtt)
    dd = np.diag(D)
    di, dj, dk = dd[None, None, None, :], dd[None, None, :, None], dd[None, :, None, None]
    A = A + 2 * di * dj * Grs + 2 * di * dk * Grt + 2 * dj * dk * Gst + lam * JW
    A = A.reshape(E, Np)
    fn = face_nodes(Nq)
    bc = np.asarray(EToB).reshape(E, 6)
    vv = v.reshape(E, 12, Np)
    for f in range(6):
        nodes = fn[f]
        s = sg[:, f * Nfp:(f + 1) * Nfp]
        i = nodes % Nq
        j = (nodes // Nq) % Nq
        k = nodes // (Nq * Nq)
        dl = [vv[:, RX + d][:, nodes] * dd[i] + vv[:, SX + d][:, nodes] * dd[j] + vv[:, TX + d][:, nodes] * dd[k] for d in range(3)]
        ndg = s[..., NX] * dl[0] + s[..., NY] * dl[1] + s[..., NZ] * dl[2]
        bcD = (bc[:, f] == 1).astype(np.float64)[:, None]
        bcN = (bc[:, f] == 2).astype(np.float64)[:, None]
        c = (1 + bcD) * (1 - bcN)
        contrib = -c * s[..., WSJ] * ndg + 0.5 * c * s[..., WSJ] * tau * s[..., IH]
        np.add.at(A, (np.arange(E)[:, None], nodes[None, :]), contrib)
    return A.reshape(-1)


def surface_factors(Nq, x, y, z, D, gllw, mapP=None, h_halo=None):
    """sgeo [E, 6*Nfp, 8]; IHID needs the neighbour's sJ/J through mapP (face-node index of the neighbour, -1 = none;
    indices >= E*6*Nfp address h_halo)"""
    Np, Nfp = Nq ** 3, Nq * Nq
    E = np.asarray(x).size // Np
    D = np.asarray(D, dtype=np.float64).reshape(Nq, Nq)
    w = np.asarray(gllw, dtype=np.float64)
    sh = (E, Nq, Nq, Nq)
    X, Y, Z = (np.asarray(a, dtype=np.float64).reshape(sh) for a in (x, y, z))
    dr = lambda F: np.einsum("im,ekjm->ekji", D, F).reshape(E, Np)
    ds = lambda F: np.einsum("jm,ekmi->ekji", D, F).reshape(E, Np)
    dt = lambda F: np.einsum("km,emji->ekji", D, F).reshape(E, Np)
    xr, xs, xt, yr, ys, yt, zr, zs, zt = dr(X), ds(X), dt(X), dr(Y), ds(Y), dt(Y), dr(Z), ds(Z), dt(Z)
    J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt)
    rx, ry, rz = (ys * zt - zs * yt) / J, -(xs * zt - zs * xt) / J, (xs * yt - ys * xt) / J
    sx, sy, sz = -(yr * zt - zr * yt) / J, (xr * zt - zr * xt) / J, -(xr * yt - yr * xt) / J
    tx, ty, tz = (yr * zs - zr * ys) / J, -(xr * zs - zr * xs) / J, (xr * ys - yr * xs) / J
    fn = face_nodes(Nq)
    sg = np.zeros((E, 6 * Nfp, 8))
    normals = [(-tx, -ty, -tz), (-sx, -sy, -sz), (rx, ry, rz), (sx, sy, sz), (-rx, -ry, -rz), (tx, ty, tz)]
    ww = (w[None, :] * w[:, None]).reshape(-1)  # gllw[i % Nq] * gllw[i / Nq]
    for f in range(6):
        nodes = fn[f]
        nx, ny, nz = (c[:, nodes] for c in normals[f])
        Jf = J[:, nodes]
        sJ = np.sqrt(nx * nx + ny * ny + nz * nz)
        s = sg[:, f * Nfp:(f + 1) * Nfp]
        s[..., NX], s[..., NY], s[..., NZ] = nx / sJ, ny / sJ, nz / sJ
        s[..., SJ] = sJ * Jf
        s[..., IJ] = 1.0 / Jf
        s[..., WIJ] = 1.0 / (Jf * w[0])
        s[..., WSJ] = sJ * Jf * ww[None, :]
    h = (sg[..., SJ] * sg[..., IJ]).reshape(-1)
    if mapP is not None:
        hall = h if h_halo is None else np.concatenate([h, np.asarray(h_halo).reshape(-1)])
        mp = np.asarray(mapP).reshape(-1)
        mp = np.where(mp < 0, np.arange(mp.size), mp)
        sg[..., IH] = np.maximum(h, hall[mp]).reshape(E, 6 * Nfp)
    return sg, h


def pcg(apply_A, inv_diag, r, tol=1e-8, maxit=5000):
    """LinearSolver::pcg (libs/linearSolver/linearSolverPCG.cpp:67-171), single rank; returns (iterations, x, history)"""
    x = np.zeros_like(r)
    r = r.copy()
    z = r * inv_diag if inv_diag is not None else r.copy()
    rdotz1 = 0.0
    hist = []
    p = np.zeros_like(r)
    rdotr = float(r @ r)
    TOL = max(tol * tol * rdotr, tol * tol)
    hist.append(np.sqrt(rdotr))
    it = 0
    while it < maxit and rdotr > TOL:
        z = r * inv_diag if inv_diag is not None else r
        rdotz2 = rdotz1
        rdotz1 = float(r @ z)
        beta = 0.0 if it == 0 else rdotz1 / rdotz2
        p = z + beta * p
        Ap = apply_A(p)
        alpha = rdotz1 / float(p @ Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        rdotr = float(r @ r)
        hist.append(np.sqrt(rdotr))
        it += 1
    return it, x, np.array(hist)
