"""CPU tests: the C oracle against the committed golden fixtures (made by independent numpy/Python models of
the reference lines) and against its own scalar build.  No GPU."""
import os

import numpy as np
import pytest

from helpers import PyNeighbourBuffer, clustered_f16, index_f16, np_fast_dot, np_flat_topk, unit_rows

G = os.path.join(os.path.dirname(__file__), "golden")


def test_fast_dot_golden(oracle):
    kat = np.load(os.path.join(G, "fast_dot_kat.npz"))
    a, b, out = kat["a"], kat["b"], kat["out"]
    for i in range(a.shape[0]):
        assert oracle.fast_dot(a[i], b[i]) == int(out[i])               # AVX2 path (vector.rs:255-306)
        assert oracle.fast_dot(a[i], b[i], scalar=True) == int(out[i])  # scalar model, same build
    sl = oracle.lib(scalar=True)                                         # no-AVX2 build
    for i in range(0, a.shape[0], 7):
        assert int(sl.orc_fast_dot(a[i].ctypes.data, b[i].ctypes.data, a.shape[1])) == int(out[i])


def test_fast_dot_batch_and_scale(oracle):
    x = index_f16(11, 50)
    got = oracle.fast_dot_batch(x[0], x)
    for i in (0, 1, 17, 49):
        assert int(got[i]) == np_fast_dot(x[0], x[i])
    # Rust `as i64`: truncation toward zero, saturation, NaN -> 0 (vector.rs:408-416)
    assert oracle.scale_dot_result(0.75) == 3 * 2 ** 30
    assert oracle.scale_dot_result(-1e-10) == 0
    assert oracle.scale_dot_result(float("nan")) == 0
    assert oracle.scale_dot_result(1e30) == 2 ** 63 - 1
    assert oracle.scale_dot_result(-1e30) == -(2 ** 63)


def test_f16_conversions(oracle):
    l = oracle.lib()
    bits = np.arange(0, 65536, dtype=np.uint32).astype(np.uint16)
    f = bits.view(np.float16).astype(np.float32)
    for b in list(range(0, 65536, 97)) + [0x0001, 0x03ff, 0x0400, 0x7bff, 0x7c00, 0xfc00, 0x8000]:
        v = l.orc_h2f(int(b))
        if np.isnan(f[b]):
            assert np.isnan(v)
        else:
            assert v == f[b]
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.standard_normal(2000).astype(np.float32) * s for s in (1e-8, 1e-5, 1.0, 300.0, 1e5)])
    vals = np.concatenate([vals, np.array([65504.0, 65519.9, 65520.0, 2.0 ** -24, 2.0 ** -25, 3 * 2.0 ** -26, 0.0, -0.0], np.float32)])
    with np.errstate(over="ignore"):
        ref = vals.astype(np.float16).view(np.uint16)
    for v, r in zip(vals, ref):
        assert l.orc_f2h(float(v)) == int(r), v


def test_neighbour_buffer_trace(oracle):
    t = np.load(os.path.join(G, "neighbour_buffer_trace.npz"))
    nb = oracle.NeighbourBuffer(24)
    for (op, id_, sc), want in zip(t["ops"], t["results"]):
        if op == 0:
            nb.insert(int(id_), int(sc))
        else:
            r = nb.next_unvisited()
            assert (-1 if r is None else r) == int(want)
    assert nb.ids.tolist() == t["final_ids"].tolist()
    assert nb.scores.tolist() == t["final_scores"].tolist()
    assert nb.cap() == 24 and len(nb) <= 24


def test_neighbour_buffer_random_vs_model(oracle):
    rng = np.random.default_rng(123)
    for cap in (1, 2, 7, 64):
        nb, py = oracle.NeighbourBuffer(cap), PyNeighbourBuffer(cap)
        for _ in range(500):
            if rng.random() < 0.75:
                i, s = int(rng.integers(0, 30)), int(rng.integers(-5, 5))
                nb.insert(i, s); py.insert(i, s)
            else:
                assert nb.next_unvisited() == py.next_unvisited()
            assert nb.ids.tolist() == py.ids
        nb.clear()
        assert len(nb) == 0 and nb.next_unvisited() is None


def test_flat_search_golden(oracle):
    g = np.load(os.path.join(G, "flat_topk.npz"))
    ids, sc = oracle.flat_search(unit_rows(3, 1), index_f16(0, 1000), 10)           # config C1
    assert ids.astype(np.int64).tolist() == g["c1_ids"].tolist()
    assert np.array_equal(sc, g["c1_scores"])
    ids, sc = oracle.flat_search(unit_rows(5, 3) * np.float32(1.7), index_f16(2, 2000), 25)
    assert ids.astype(np.int64).tolist() == g["m_ids"].tolist()
    assert np.array_equal(sc, g["m_scores"])


def test_flat_search_edges(oracle):
    x = index_f16(4, 7)
    q = unit_rows(6, 2)
    ids, sc = oracle.flat_search(q, x, 10)  # k > n: padded like faiss (-1 labels -> UINT32_MAX here)
    ref_ids, ref_sc = np_flat_topk(q, x, 10)
    assert (ids[:, :7].astype(np.int64) == ref_ids[:, :7]).all() and (ids[:, 7:] == 0xFFFFFFFF).all()
    assert np.isneginf(sc[:, 7:]).all()
    # exact ties: duplicated rows must come back in id order
    xd = np.concatenate([x, x, x])
    ids, _ = oracle.flat_search(q, xd, 6)
    ref_ids, _ = np_flat_topk(q, xd, 6)
    assert (ids.astype(np.int64) == ref_ids).all()
    # the AVX2 f32 scan (CPU-baseline mode) agrees with the f64 oracle on well separated data
    x = index_f16(8, 3000)
    a, _ = oracle.flat_search(q, x, 10, mode=0)
    b, _ = oracle.flat_search(q, x, 10, mode=1)
    assert (a == b).all()


def test_medioid_and_random_fill(oracle):
    x = index_f16(9, 300)
    c = np.zeros(1152, np.float32)
    for i in range(300):
        c = c + (x[i].astype(np.float32) - c) * np.float32(1.0 / (i + 1))
    s = x.astype(np.float64) @ c.astype(np.float16).astype(np.float64)
    assert oracle.medioid(x) == int(np.flatnonzero(s == s.max())[-1])
    g = oracle.IndexGraph(300, 16)
    oracle.random_fill_graph(g, 16, seed=3)
    assert (g.deg == 16).all()
    assert all(len(set(row.tolist())) == 16 for row in g.adj)


def _py_greedy(oracle, start, q, x, adj, deg, L):
    nb = PyNeighbourBuffer(L)
    visited = {start}
    nb.insert(start, np_fast_dot(q, x[start]))
    vlist, dist = [], 0
    while True:
        pt = nb.next_unvisited()
        if pt is None:
            break
        pre = []
        for n in adj[pt, :deg[pt]]:
            n = int(n)
            if n not in visited:
                visited.add(n); pre.append(n)
        for n in pre:
            s = np_fast_dot(q, x[n]); dist += 1
            nb.insert(n, s); vlist.append((n, s))
    return nb.ids, nb.scores, vlist, dist


def test_greedy_search_vs_python_model(oracle):
    n = 400
    x = index_f16(21, n)
    cfg = oracle.make_config(r=8, l=16, maxc=50)
    g = oracle.IndexGraph(n, 8)
    oracle.random_fill_graph(g, 8, seed=5)
    s = oracle.Scratch(n, cfg)
    for qi in (0, 5, 77):
        d = oracle.greedy_search(s, 3, False, x[qi], x, g, cfg)
        ids, scores, vlist, dist = _py_greedy(oracle, 3, x[qi], x, g.adj, g.deg, 16)
        assert s.neighbour_ids.tolist() == ids and s.neighbour_scores.tolist() == scores and d == dist
        vi, vs = s.visited_list()
        assert list(zip(vi.tolist(), vs.tolist())) == vlist


def _py_robust_prune(p, cands, x, r, maxc, alpha):
    c = sorted(enumerate(cands), key=lambda t: (-t[1][1], t[0]))
    c = [list(t[1]) for t in c][:maxc]
    out, ci = [], 0
    MIN = -(2 ** 63)
    while len(out) < r and ci < len(c):
        ps, pss = c[ci]; ci += 1
        if ps == p or pss == MIN:
            continue
        out.append(ps)
        for i in range(ci + 1, len(c)):
            if c[i][1] == MIN:
                continue
            sp = np_fast_dot(x[c[i][0]], x[ps])
            if ((alpha * sp) >> 16) >= c[i][1]:
                c[i][1] = MIN
    return out


def test_robust_prune_vs_python_model(oracle):
    x = index_f16(31, 200)
    rng = np.random.default_rng(1)
    for p in (0, 9):
        cand = rng.choice(200, 60, replace=False)
        cand_s = [np_fast_dot(x[p], x[c]) for c in cand]
        for alpha in (65536, 65200, 78643):
            cfg = oracle.make_config(r=12, l=32, maxc=40, alpha=alpha)
            got = oracle.robust_prune(p, cand, cand_s, x, cfg)
            want = _py_robust_prune(p, list(zip(cand.tolist(), cand_s)), x, 12, 40, alpha)
            assert got.tolist() == want


def test_build_graph_recall(oracle):
    from helpers import clustered_f16
    n = 1500
    x = clustered_f16(41, n, n_clusters=16)
    cfg = oracle.make_config(r=24, l=48, maxc=200)
    g = oracle.IndexGraph(n, 24)
    oracle.random_fill_graph(g, 24, seed=2)
    med = oracle.medioid(x)
    oracle.build_graph(g, med, x, cfg, seed=9)
    assert g.deg.max() <= 24 and g.deg.min() >= 1
    s = oracle.Scratch(n, cfg)
    hits = 0
    for qi in range(0, n, 15):
        oracle.greedy_search(s, med, False, x[qi], x, g, cfg)
        bf, _ = oracle.brute_force_i64(x[qi], x, 1)
        hits += int(s.neighbour_ids[0] == bf[0])
    assert hits >= 95  # of 100: self-query recall@1 like diskann/src/main.rs:106-136
    # deterministic when sequential
    g2 = oracle.IndexGraph(n, 24)
    oracle.random_fill_graph(g2, 24, seed=2)
    oracle.build_graph(g2, med, x, cfg, seed=9)
    assert np.array_equal(g.adj, g2.adj) and np.array_equal(g.deg, g2.deg)


def test_pq(oracle):
    rng = np.random.default_rng(5)
    D, S, Cn = 1152, 18, 256
    T, _ = np.linalg.qr(rng.standard_normal((D, D)))
    T = T.astype(np.float32)
    cent = (rng.standard_normal((Cn, D)) / np.sqrt(D)).astype(np.float32)
    pq = oracle.ProductQuantizer(cent, T, S)
    x = unit_rows(7, 20)
    y = pq.apply_transform(x)
    assert np.allclose(y, x @ T.T, atol=2e-6)
    codes = pq.quantize_batch(x)
    yy = (x.astype(np.float64) @ T.astype(np.float64).T).reshape(20, 64, S)
    sims = np.einsum("vms,cms->vmc", yy, cent.astype(np.float64).reshape(Cn, 64, S))
    top2 = np.sort(sims, axis=2)[:, :, -2:]
    clear = (top2[:, :, 1] - top2[:, :, 0]) > 1e-6  # argmax is only pinned away from near-ties
    assert (codes[clear] == sims.argmax(axis=2)[clear]).all() and clear.mean() > 0.99
    lut = pq.preprocess_query(x[0])
    assert lut.shape == (64, 256)
    assert np.allclose(lut, np.einsum("ms,cms->mc", yy[0], cent.astype(np.float64).reshape(Cn, 64, S)), atol=1e-6)
    adc = pq.asymmetric_dot_product(lut, codes)
    for j in (0, 7, 19):
        s = np.float32(0)
        for i in range(64):
            s = np.float32(s + lut[i, codes[j, i]])
        assert int(adc[j]) == int(np.trunc(np.float64(np.float32(s * np.float32(4294967296.0)))))


def test_beam_search_faithful_equals_clean(oracle):
    """query_disk_index.rs:157 clears the pre-buffer once per beam; per-node clearing returns the same nodes."""
    from helpers import clustered_f16
    n = 1200
    x = clustered_f16(51, n, n_clusters=12)
    cfg = oracle.make_config(r=16, l=32, maxc=100)
    g = oracle.IndexGraph(n, 16)
    oracle.random_fill_graph(g, 16, seed=4)
    med = oracle.medioid(x)
    oracle.build_graph(g, med, x, cfg, seed=1)
    adj, off = g.to_csr()
    rng = np.random.default_rng(3)
    T, _ = np.linalg.qr(rng.standard_normal((1152, 1152)))
    pq = oracle.ProductQuantizer(x[rng.choice(n, 256, replace=False)].astype(np.float32), T.astype(np.float32), 18)
    codes = pq.quantize_batch(x.astype(np.float32))
    for qi in (1, 100):
        lut = pq.preprocess_query(x[qi].astype(np.float32))
        a = oracle.beam_search(x, adj, off, codes, lut, med, x[qi], 40, 3, faithful_prebuffer=True)
        b = oracle.beam_search(x, adj, off, codes, lut, med, x[qi], 40, 3, faithful_prebuffer=False)
        assert a[0].tolist() == b[0].tolist() and a[1].tolist() == b[1].tolist()
        assert a[2][0] == b[2][0] and a[2][1] >= b[2][1]
        best = a[0][np.argmax(a[1])]
        assert best == qi


def test_greedy_search_batch_equals_single(oracle):
    """The OpenMP batch entry bench.py's cpu_baseline times returns what the per-query entry returns."""
    n, R, L = 1200, 12, 24
    x = clustered_f16(21, n, n_clusters=8)
    g = oracle.IndexGraph(n, R)
    oracle.random_fill_graph(g, R, seed=1)
    cfg = oracle.make_config(r=R, l=L, maxc=100)
    oracle.build_graph(g, oracle.medioid(x), x, cfg, seed=2)
    q = clustered_f16(22, 37, n_clusters=8)
    ids, sc, ln, dist = oracle.greedy_search_batch(3, q, x, g, cfg)
    s = oracle.Scratch(n, cfg)
    for i in range(q.shape[0]):
        d = oracle.greedy_search(s, 3, False, q[i], x, g, cfg)
        m = int(ln[i])
        assert m == len(s.neighbour_ids) and d == int(dist[i])
        assert np.array_equal(ids[i, :m], s.neighbour_ids) and np.array_equal(sc[i, :m], s.neighbour_scores)
        assert (ids[i, m:] == 0xFFFFFFFF).all()


def test_rabitq_direct_estimate_matches_numpy(oracle):
    """The table-free RabitQ estimate the GPU traversal uses (query side quantised to 8-bit planes, integer popcount sum -- the
    RabitQ paper's form) against the numpy restatement of diskann/rabitq.py:42-48 (f64): the quantisation costs < 5e-4 absolute
    for unit vectors (measured: max 1.8e-4, mean 3e-5), two orders below the estimator's own noise."""
    from oracle.rabitq_np import RabitQ as NpRabitQ
    x = clustered_f16(41, 600, n_clusters=8)
    ref = NpRabitQ.train(x[:300].astype(np.float32), output_dims=512, seed=2)
    bits, norms, dots, _ = ref.quantize(x)
    codes = NpRabitQ.pack(bits)
    q = unit_rows(42, 2)
    for i in range(2):
        qt = (ref.p.astype(np.float64) @ q[i].astype(np.float64)).astype(np.float32)
        mq = np.float32(np.dot(ref.mean.astype(np.float64), q[i].astype(np.float64)))
        qtm = np.concatenate([qt, [mq]]).astype(np.float32)
        got = oracle.rabitq_direct_estimates(qtm, np.float32(1.0 / np.sqrt(1152.0)), codes, (norms * dots).astype(np.float32))
        want = ref.approx_dot(bits, norms, dots, q[i])
        assert np.abs(got - want).max() < 5e-4 and np.abs(got - want).mean() < 1e-4


def test_robust_stitch_vs_python_model(oracle):
    """orc_robust_stitch_order against a line-by-line Python model of diskann/src/lib.rs:326-374 (sequential loop over the
    shuffled query order; stable descending sort; `contains` check; max_add and r limits)."""
    n, qb, R, max_add = 90, 60, 6, 2
    x = clustered_f16(77, n, n_clusters=4, d=64)
    rng = np.random.default_rng(9)
    adj = rng.integers(0, n, (n, R)).astype(np.uint32)
    deg = rng.integers(2, R + 1, n).astype(np.uint32)
    order = (qb + rng.permutation(n - qb)).astype(np.uint32)
    g = oracle.IndexGraph(n, R)
    g.set(adj, deg)
    cfg = oracle.make_config(r=R, l=16, maxc=50, query_breakpoint=qb, max_add_per_stitch_iter=max_add)
    oracle.robust_stitch(g, x, cfg, order=order)
    # model
    lists = [adj[i, : deg[i]].tolist() for i in range(n)]
    in_edges = [[] for _ in range(n - qb)]
    for b in range(qb):                                            # :339-347
        keep = []
        for t in lists[b]:
            if t >= qb:
                in_edges[t - qb].append(b)
            else:
                keep.append(t)
        lists[b] = keep
    for q in order.tolist():                                       # :349-372
        qn = lists[q]
        for b in in_edges[q - qb]:
            cands = [(nb, np_fast_dot(x[b], x[nb])) for nb in qn]
            cands = [c for _, c in sorted(enumerate(cands), key=lambda t: (-t[1][1], t[0]))]
            added = 0
            for nb, _ in cands:
                if added >= max_add or len(lists[b]) >= R:
                    break
                if nb in lists[b]:
                    continue
                lists[b].append(nb)
                added += 1
    for i in range(n):
        assert g.adj[i, : g.deg[i]].tolist() == lists[i], i


def test_dedup_visited_model(oracle):
    """orc_dedup_visited against the reference's formulation (query_disk_index.rs:486-527): full similarity matrix, then the
    sequential retain with the `included` bit vector."""
    rng = np.random.default_rng(5)
    base = unit_rows(44, 12)
    rows = []
    for i in range(40):                                            # near-duplicates of 12 prototypes + a few unrelated rows
        b = base[rng.integers(0, 12)]
        v = b + rng.standard_normal(1152).astype(np.float32) * rng.choice([0.002, 0.02])
        rows.append(v / np.linalg.norm(v))
    x = np.stack(rows).astype(np.float16)
    ids = rng.permutation(40).astype(np.uint32)
    keep = oracle.dedup_visited(x, ids, 0.95)
    v = x[ids].astype(np.float32)
    sim = v @ v.T
    included = np.zeros(40, bool)
    for i in range(40):
        if not (included & (sim[i] > 0.95)).any():
            included[i] = True
    near = np.abs(sim - 0.95) < 1e-4                               # decisions may differ only within rounding of the threshold
    assert not near.any()
    assert np.array_equal(keep, included) and 1 < keep.sum() < 40
