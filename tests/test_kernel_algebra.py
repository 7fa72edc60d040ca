"""CPU checks (numpy) of the algebra behind the round-2 kernel forms.  Each test restates the exact operation sequence of the
device code it cites and compares it with the plain formula, so an error in the derivation cannot hide behind the GPU tests'
tolerances.  No GPU needed."""
import math

import numpy as np
import pytest
import scipy.linalg

RNG = np.random.default_rng(20261017)

# constants of csrc/small_d.cuh (Bader, Blanes, Casas 2019) and their recombination (T8_YA .. T8_R7)
X1, X2, X3, X4 = 0.10836465678522780852, 0.027091164196306952131, 2.0 / 3.0, 0.54676145797072405251
X5, X6, X7, Y2 = 0.16112557339541759283, 0.014090917158378207731, 0.033792797010870504141, 0.13549236135285063166
YA, YB = 0.25, 6.1520673478250353627
R4, R5, R6, R7 = 0.059249617736388271447, 0.017460317460317460317, -0.00091433916494050425595, 0.00039682539682539682538


def _antiherm(D, norm1):
    H = RNG.standard_normal((D, D)) + 1j * RNG.standard_normal((D, D))
    G = -1j * (H + H.conj().T)
    return G * (norm1 / np.linalg.norm(G, 1))


def test_recombined_taylor8_equals_the_three_product_scheme_and_exp():
    """expm_t8 (csrc/small_d.cuh): Yh = A + YA A2 + YB I, Lh = A2 Yh, Rh = R4 I + R5 A + R6 A2 + R7 Lh, T8 = I + A + Y2 A2 + Lh Rh
    is the same polynomial as A4 = A2 (x1 A + x2 A2), A8 = (x3 A2 + A4)(x4 I + x5 A + x6 A2 + x7 A4), T8 = I + A + y2 A2 + A8,
    and both equal sum_k A^k / k!, k <= 8; at the scaling threshold 0.0694 that is exp(A) to double precision."""
    assert abs(YA - X2 / X1) < 1e-17 and abs(YB - X3 / X1) < 1e-15
    assert abs(R4 - X1 * X4) < 1e-17 and abs(R5 - X1 * X5) < 1e-17 and abs(R5 - 11 / 630) < 1e-17
    assert abs(R6 - X1 * (X6 - X7 * X3)) < 1e-18 and abs(R7 - X1 * X1 * X7) < 1e-18 and abs(R7 - 1 / 2520) < 1e-18
    for D in (3, 8, 16):
        A = _antiherm(D, 0.0694)
        I = np.eye(D)
        A2 = A @ A
        A4 = A2 @ (X1 * A + X2 * A2)
        T8_ref = I + A + Y2 * A2 + (X3 * A2 + A4) @ (X4 * I + X5 * A + X6 * A2 + X7 * A4)
        Lh = A2 @ (A + YA * A2 + YB * I)
        T8_new = I + A + Y2 * A2 + Lh @ (R4 * I + R5 * A + R6 * A2 + R7 * Lh)
        series = sum(np.linalg.matrix_power(A, k) / float(math.factorial(k)) for k in range(9))
        assert np.max(np.abs(T8_new - T8_ref)) < 5e-16
        assert np.max(np.abs(T8_new - series)) < 5e-16
        assert np.max(np.abs(T8_new - scipy.linalg.expm(A))) < 1e-15


def _nt_3m(a, b, conj_a=False, conj_b=False, acc=None):
    """mul_nt_impl with QOC_3M (csrc/warp_mat.cuh): C (+)= op(A) op(B)^T through K1 = (ar + ai') br, K3 = ai' (br + bi'),
    K2 = ar (bi' - br); K1 seeds both accumulators."""
    ar, ai = a.real, (-a.imag if conj_a else a.imag)
    br, bi = b.real, (-b.imag if conj_b else b.imag)
    sA, sB, dB = ar + ai, br + bi, bi - br
    k = sA @ br.T + (acc.real if acc is not None else 0.0)
    re = k + (-ai) @ sB.T
    m0 = k + ((acc.imag - acc.real) if acc is not None else 0.0)
    im = m0 + ar @ dB.T
    return re + 1j * im


@pytest.mark.parametrize("conj_a,conj_b", [(False, False), (True, False), (False, True), (True, True)])
def test_three_multiplication_product_all_conjugation_variants(conj_a, conj_b):
    for D in (8, 16):
        a = RNG.standard_normal((D, D)) + 1j * RNG.standard_normal((D, D))
        b = RNG.standard_normal((D, D)) + 1j * RNG.standard_normal((D, D))
        c = RNG.standard_normal((D, D)) + 1j * RNG.standard_normal((D, D))
        ref = (a.conj() if conj_a else a) @ (b.conj() if conj_b else b).T
        assert np.max(np.abs(_nt_3m(a, b, conj_a, conj_b) - ref)) < 1e-13 * D
        assert np.max(np.abs(_nt_3m(a, b, conj_a, conj_b, acc=c) - (c + ref))) < 1e-13 * D


def test_conjugation_step_with_left_operand_sums():
    """conj_by (csrc/warp_mat.cuh): X = conj(P) W^T and R = P X^T = P W P' with K1 = ar (br + bi), K2 = (ai' - ar) br,
    K3 = (ar + ai') bi; both passes use only -(pr + pi) and pi - pr of P."""
    for D in (8, 16):
        P = scipy.linalg.expm(_antiherm(D, 0.7))
        W = RNG.standard_normal((D, D)) + 1j * RNG.standard_normal((D, D))
        pr, pi = P.real, P.imag
        nsp, nsm = -(pr + pi), pi - pr
        out = []
        B = W
        for first in (True, False):
            sB = B.real + B.imag
            k = pr @ sB.T
            re = k + (nsm if first else nsp) @ B.imag.T
            im = k + (nsp if first else nsm) @ B.real.T
            B = re + 1j * im
            out.append(B)
        assert np.max(np.abs(out[0] - P.conj() @ W.T)) < 1e-13 * D
        assert np.max(np.abs(out[1] - P @ W @ P.conj().T)) < 1e-13 * D


def test_square_of_an_antihermitian_generator_shares_its_sums():
    """square_antiherm (csrc/warp_mat.cuh): G G = nt(G, Gt), Gt = G^T = -conj(G); with br = -gr, bi = gi the right operand's sums
    are gi - gr and gr + gi, the latter being the left operand's own."""
    for D in (8, 16):
        G = _antiherm(D, 0.05)
        gr, gi = G.real, G.imag
        sp, dm = gr + gi, gi - gr
        k = sp @ (-gr).T
        re = k + (-gi) @ dm.T
        im = k + gr @ sp.T
        assert np.max(np.abs((re + 1j * im) - G @ G)) < 1e-17 * D
        assert np.max(np.abs(G.T + G.conj())) == 0.0


def test_infinity_norm_serves_the_taylor_bound():
    """cm_norm1_bound with QOC_NORM_INF: row sums of |re| + |im| bound the induced infinity norm from above (|z| <= |re| + |im|),
    and the degree-8 truncation error at that bound stays below 2^-53 like for the 1-norm."""
    for D in (5, 8, 16):
        G = _antiherm(D, 1.0)
        bound = np.max(np.sum(np.abs(G.real) + np.abs(G.imag), axis=1))
        assert bound >= np.linalg.norm(G, np.inf)
        s = max(0, int(np.floor(np.log2(bound / 0.0694))) + 1) if bound > 0.0694 else 0
        A = G / 2.0 ** s
        assert np.linalg.norm(A, np.inf) <= 0.0694 * 1.0000001
        T8 = sum(np.linalg.matrix_power(A, k) / float(math.factorial(k)) for k in range(9))
        assert np.max(np.abs(T8 - scipy.linalg.expm(A))) < 2e-16 * D


def _tables(members, D, K):
    """numpy restatement of analyze_structure (csrc/qocgrape.cu): plane lists, compact assembly blocks, compact dot slots.
    Flat entry of the packed layout: plane * 64 + 8 row + col; the packed matrices are -i dt X, so their real plane is
    dt Im X and their imaginary plane -dt Re X."""
    mats = [[m[0] for m in members]] + [[m[1][j] for m in members] for j in range(K)]
    herm = all(np.array_equal(x, x.conj().T) for ms in mats for x in ms)
    lr = [j for j, ms in enumerate(mats) if any(np.any(x.imag != 0) for x in ms)]
    li = [j for j, ms in enumerate(mats) if any(np.any(x.real != 0) for x in ms)]
    used = np.zeros(128, bool)
    used_b = np.zeros(128, bool)
    for j, ms in enumerate(mats):
        for x in ms:
            for r in range(D):
                for c in range(D):
                    if x[r, c].imag != 0:
                        used[8 * r + c] = True
                        used_b[8 * r + c] |= j > 0
                    if x[r, c].real != 0:
                        used[64 + 8 * r + c] = True
                        used_b[64 + 8 * r + c] |= j > 0
    pos = []
    for pl in range(2):
        pos += [f for f in range(64 * pl, 64 * pl + 64) if used[f]]
        pos += [-1] * (-len(pos) % 8)
        if pl == 0:
            nblk_re = len(pos) // 8
    nblk = len(pos) // 8
    if not (0 < nblk <= 12):
        nblk = nblk_re = 0
    slots = [f for f in range(128) if used_b[f]]
    nks = -(-len(slots) // 16) * 4
    if not (len(slots) > 0 and nks <= 24):
        nks = 0
    return dict(herm=int(herm), sparse=int(len(lr) <= 4 and len(li) <= 4), lr=lr[:4], li=li[:4], nblk_re=nblk_re, nblk=nblk,
                nks=nks, pos=pos, slots=slots)


def _library_tables(members, D, K, shared=False):
    import ctypes as C
    import quoptimalcontrol_jl_b200 as qoc
    lib = qoc._lib.load()
    mem = members[:1] if shared else members
    A = np.ascontiguousarray(np.stack([np.asarray(m[0], dtype=np.complex128).T for m in mem]))       # column-major
    B = np.ascontiguousarray(np.stack([np.stack([np.asarray(b, dtype=np.complex128).T for b in m[1]]) for m in mem]))
    out = (C.c_int * 392)()
    flags = (qoc._lib.SHARED_A | qoc._lib.SHARED_B) if shared else 0
    rc = lib.qoc_analyze_structure(D, K, len(members), A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), flags, out)
    assert rc == 0
    o = np.array(out[:], dtype=np.int64)
    unpack = lambda v: [b for b in ((int(v) >> (8 * i)) & 0xff for i in range(4)) if b != 0xff]
    return dict(herm=int(o[0]), sparse=int(o[1]), lr=unpack(o[2]), li=unpack(o[3]), nblk_re=int(o[4]), nblk=int(o[5]), nks=int(o[6]),
                pos=o[8:136], tab=o[136:392])


def test_structure_tables_of_the_bench_config():
    """cfg4 (the bench default): 3 matrices with a real plane (sigma_y controls), 4 with an imaginary plane (diagonal drift,
    sigma_x controls) -> plane-wise assembly; 24 + 32 non-zero generator entries -> 3 + 4 compact blocks instead of 16;
    48 non-zero control entries -> 12 dot k-steps instead of 32.  These are the counts DESIGN.md section 3.2 quotes, from the
    library's own analysis (qoc_analyze_structure runs without a GPU)."""
    import quoptimalcontrol_jl_b200 as qoc
    cfg = qoc.configs.config4(N=4, grid=3)
    t = _library_tables(cfg["members"], 8, 6)
    assert t["herm"] == 1 and t["sparse"] == 1 and t["lr"] == [2, 4, 6] and t["li"] == [0, 1, 3, 5]
    assert (t["nblk_re"], t["nblk"], t["nks"]) == (3, 7, 12)


@pytest.mark.parametrize("kind,D", [("pauli", 8), ("pauli_k7", 8), ("real_sparse", 8), ("real_sparse", 6), ("mixed_member", 5),
                                    ("dense", 8), ("dense", 7)])
def test_structure_analysis_matches_its_restatement(kind, D):
    """The library's structure analysis against the numpy restatement on the control structures of test_gpu_structure.py
    (which runs the kernels these tables drive against the oracle)."""
    from test_gpu_structure import _members
    members = _members(kind, D, 12, seed=11 + D)
    D = members[0][0].shape[0]
    K = len(members[0][1])
    t, r = _library_tables(members, D, K), _tables(members, D, K)
    for key in ("herm", "sparse", "nblk_re", "nblk", "nks"):
        assert t[key] == r[key], key
    if r["sparse"]:
        assert t["lr"] == r["lr"] and t["li"] == r["li"]
    full = r["pos"] + [-1] * (128 - len(r["pos"]))
    assert list(t["pos"]) == full                      # the positions are listed even when compaction is not worth it
    assert [f for f in range(128) if t["tab"][f] >= 0] == r["slots"]
    for sidx, f in enumerate(r["slots"]):
        assert t["tab"][f] == sidx and t["tab"][128 + sidx] == f
    assert all(t["tab"][128 + sidx] == -1 for sidx in range(len(r["slots"]), 128))


def test_structure_analysis_of_shared_matrices_and_bad_arguments():
    import ctypes as C
    import quoptimalcontrol_jl_b200 as qoc
    from test_gpu_structure import _members
    members = _members("pauli", 8, 3, seed=3)
    members = [members[0]] * 3
    assert _library_tables(members, 8, 6, shared=True)["nks"] == _library_tables(members, 8, 6)["nks"] == 12
    out = (C.c_int * 392)()
    assert qoc._lib.load().qoc_analyze_structure(0, 1, 1, None, None, 0, out) == qoc._lib.QOC_EINVAL


# ---- index maps of the tensor-pipe forms, emulated lane by lane ------------------------------------------------------------
# DMMA.8x8x4 (mma.sync.m8n8k4.f64): lane (g = lane >> 2, q = lane & 3) supplies A[m = g][k = q] and B[k = q][n = g] and receives
# D[m = g][n = 2 q], D[m = g][n = 2 q + 1].  Packed layout of a D <= 8 matrix X: flat entry f = plane * 64 + 8 row + col, lane
# (g, q) holds rows g, columns 2 q, 2 q + 1 (f = 2 lane + e within a plane).
def _dmma(acc, a_frag, b_frag):
    """acc [8][8] += A [8][4] B [4][8] from per-lane fragments a_frag[lane], b_frag[lane]."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8))
    for lane in range(32):
        g, q = lane >> 2, lane & 3
        A[g, q] = a_frag[lane]; B[q, g] = b_frag[lane]
    return acc + A @ B


def _packed(X, scale):
    """flat [128] image of scale * X (8 x 8 complex): real plane then imaginary plane, entry 8 row + col."""
    Z = scale * X
    return np.concatenate([Z.real.reshape(-1), Z.imag.reshape(-1)])


def test_compact_plane_wise_assembly_index_maps():
    """chunk_expm_dmma_item, compact form: coefficient staging Ad[b][lane], amplitude fragments (list entry q of the block's
    plane, slice g), scatter of D[m = g][n = slice] to flat position asm_pos[8 b + g] -- reproduces G_t = A~ + sum_j x[j,t] B~_j
    for 8 slices, with every entry outside the union left at zero."""
    import quoptimalcontrol_jl_b200 as qoc
    cfg = qoc.configs.config4(N=8, grid=2)
    A, Bs = cfg["members"][3][0], cfg["members"][3][1]
    K, dt = len(Bs), cfg["T"] / cfg["N"]
    t = _library_tables(cfg["members"], 8, K)
    assert t["nblk"] == 7
    sys_flat = [_packed(A, -1j * dt)] + [_packed(b, -1j * dt) for b in Bs]       # index 0 = drift, j = control j
    x = RNG.uniform(-1, 1, (K, 8))
    lists = [t["lr"] + [0xff] * (4 - len(t["lr"])), t["li"] + [0xff] * (4 - len(t["li"]))]
    G = np.zeros((8, 128))                                                        # [slice][flat entry]
    for b in range(t["nblk"]):
        lst = lists[0] if b < t["nblk_re"] else lists[1]
        a_frag, b_frag = np.zeros(32), np.zeros(32)
        for lane in range(32):
            g, q = lane >> 2, lane & 3
            f, j = t["pos"][8 * b + g], lst[q]
            a_frag[lane] = sys_flat[j][f] if (f >= 0 and j <= K) else 0.0
            b_frag[lane] = 0.0 if j > K else (1.0 if j == 0 else x[j - 1, g])    # lane (g = slice, q = list entry)
        Dm = _dmma(np.zeros((8, 8)), a_frag, b_frag)                              # [m = row of block][n = slice]
        for g in range(8):
            f = t["pos"][8 * b + g]
            if f >= 0:
                G[:, f] = Dm[g, :]
    for s in range(8):
        ref = _packed(A + sum(x[j, s] * Bs[j] for j in range(K)), -1j * dt)
        assert np.max(np.abs(G[s] - ref)) < 1e-17


def test_compact_trace_dot_index_maps():
    """sweep_unitary_dmma_item, compact form: W staged at slot dot_tab[f], control fragments Bd[ks][lane (c, q)] =
    B~flat[c][dot_tab[128 + 4 ks + q]], slice fragments W[slice g][slot 4 ks + q]; the accumulated chain gives
    g[c][t] = sum_f B~flat[c][f] Wflat[t][f] = Re tr(conj(B~_c) .* W_t) summed over entries."""
    import quoptimalcontrol_jl_b200 as qoc
    cfg = qoc.configs.config4(N=8, grid=2)
    Bs = cfg["members"][1][1]
    K, dt = len(Bs), cfg["T"] / cfg["N"]
    t = _library_tables(cfg["members"], 8, K)
    nks = t["nks"]
    assert nks == 12
    Bflat = [_packed(b, -1j * dt) for b in Bs]
    W = RNG.standard_normal((8, 8, 8)) + 1j * RNG.standard_normal((8, 8, 8))      # 8 slices
    Wb = np.zeros((8, 4 * nks))                                                   # padding slots hold zeros
    for s in range(8):
        flat = _packed(W[s], 1.0)
        for f in range(128):
            if t["tab"][f] >= 0:
                Wb[s, t["tab"][f]] = flat[f]
    acc = np.zeros((8, 8))                                                        # [m = control][n = slice]
    for ks in range(nks):
        a_frag, b_frag = np.zeros(32), np.zeros(32)
        for lane in range(32):
            c, q = lane >> 2, lane & 3
            f = t["tab"][128 + 4 * ks + q]
            a_frag[lane] = Bflat[c][f] if (c < K and f >= 0) else 0.0
            b_frag[lane] = Wb[c, 4 * ks + q]                                      # here g = slice index of the lane
        acc = _dmma(acc, a_frag, b_frag)
    for c in range(K):
        for s in range(8):
            ref = np.sum((np.conj(-1j * dt * Bs[c]) * W[s]).real)
            assert abs(acc[c, s] - ref) < 1e-15
