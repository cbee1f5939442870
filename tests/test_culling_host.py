"""Host-side bookkeeping of the product's CullingSystem (cell grid, page chains, entity->slot map; no GPU needed)
against the oracle's restatement of culling_system.cpp:98-258 — page for page, in m_cells order."""
import numpy as np

import lumixengine_b200 as lb
from lumixengine_b200 import scenes


def _same_state(cs, oc):
    pa, pb = cs.pages(), oc.pages()
    assert len(pa) == len(pb)
    for a, b in zip(pa, pb):
        assert a["origin"] == b["origin"] and a["indices"] == b["indices"]
        assert a["type"] == b["type"] and a["is_big"] == b["is_big"] and a["count"] == b["count"]
        assert np.array_equal(a["entities"], b["entities"])
        assert np.array_equal(a["spheres"].view(np.uint32), b["spheres"].view(np.uint32))


def test_build_matches_reference_layout(oracle):
    scene = scenes.cull_scene(50_000, (1500.0, 300.0, 1500.0), seed=4, big_fraction=0.01, type_probs=(0.7, 0.2, 0.1))
    cs = lb.CullingSystem(None)
    oc = oracle.OracleCulling()
    cs.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
    oc.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
    assert cs.entity_count() == 50_000
    _same_state(cs, oc)
    # pages hold at most 200 spheres (count < MAX_COUNT - 1, culling_system.cpp:103), cells with more chain pages
    counts = [p["count"] for p in cs.pages()]
    assert max(counts) == 200 and min(counts) >= 1


def test_incremental_api_matches(oracle):
    rng = np.random.default_rng(12)
    n = 30_000
    scene = scenes.cull_scene(n, (900.0, 100.0, 900.0), seed=8)
    cs = lb.CullingSystem(None)
    oc = oracle.OracleCulling()
    cs.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
    oc.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
    alive = np.ones(n, bool)
    for step in range(6):
        ids = rng.choice(np.nonzero(alive)[0], 2500, replace=False).astype(np.int32)
        a, b, c, d = np.array_split(ids, 4)
        p = scene["pos"][a] + rng.normal(size=(len(a), 3)) * np.array([300.0, 30.0, 300.0])  # many cross cell borders
        cs.setPosition(a, p); oc.set_position(a, p)
        r = (rng.random(len(b)) * 650.0).astype(np.float32)  # crosses the is_big threshold (radius > 300)
        cs.setRadius(b, r); oc.set_radius(b, r)
        p2 = scene["pos"][c] * 0.5
        r2 = (rng.random(len(c)) * 400.0).astype(np.float32)
        cs.set(c, p2, r2); oc.set(c, p2, r2)
        cs.remove(d); oc.remove(d)
        alive[d] = False
        _same_state(cs, oc)
        for e in (int(a[0]), int(b[0]), int(d[0])):
            assert cs.isAdded(e) == oc.is_added(e)
        assert cs.getRadius(int(b[1])) == oc.get_radius(int(b[1]))
    # re-adding removed entities reuses freed pages without disturbing the order of m_cells
    back = np.nonzero(~alive)[0].astype(np.int32)
    cs.add(back, scene["types"][back], scene["pos"][back], scene["radius"][back])
    oc.add(back, scene["types"][back], scene["pos"][back], scene["radius"][back])
    _same_state(cs, oc)
    assert cs.entity_count() == n


def test_cell_index_truncates_toward_zero(oracle):
    """culling_system.cpp:27 + math.cpp:133-138: int(double) — cells (-300,300) share index 0."""
    cs = lb.CullingSystem(None)
    pts = np.array([[-299.9, 0.0, 0.0], [299.9, 0.0, 0.0], [-300.1, 0.0, 0.0], [300.0, -0.5, 599.999], [-1e-9, -299.0, -600.0]])
    cs.add(np.arange(5), 0, pts, 1.0)
    idx = sorted(p["indices"] for p in cs.pages())
    assert idx == sorted({(0, 0, 0), (-1, 0, 0), (1, 0, 1), (0, 0, -2)})
    oc = oracle.OracleCulling()
    oc.add(np.arange(5), np.zeros(5, np.uint8), pts, np.ones(5, np.float32))
    _same_state(cs, oc)


def test_type_0xff_is_reserved():
    cs = lb.CullingSystem(None)
    try:
        cs.add(1, 0xFF, (0.0, 0.0, 0.0), 1.0)
    except lb.LumixB200Error as e:
        assert e.code == lb._lib.ERR_INVALID
    else:
        raise AssertionError("type 0xff must be rejected (culling_system.cpp:312)")


def test_page_churn_in_crowded_cells(oracle):
    """Two crowded cells (several chained pages each): random removals empty and free pages in the middle of a chain, re-adds
    prepend fresh pages (culling_system.cpp:103-128, 160-197), radius edits flip entities between the normal and the is_big
    chain.  After every batch the page list must equal the oracle's, page for page in m_cells order, and freed device page ids
    must be reused without ever exceeding the high-water mark of the busiest moment."""
    rng = np.random.default_rng(77)
    n = 2400
    ents = np.arange(n, dtype=np.int32)
    pos = np.stack([rng.uniform(10.0, 590.0, n), rng.uniform(5.0, 295.0, n), rng.uniform(10.0, 290.0, n)], axis=1)  # cells (0,0,0) and (1,0,0)
    rad = rng.uniform(0.5, 5.0, n).astype(np.float32)
    types = (rng.random(n) < 0.3).astype(np.uint8)
    cs, oc = lb.CullingSystem(None), oracle.OracleCulling()
    cs.add(ents, types, pos, rad); oc.add(ents, types, pos, rad)
    _same_state(cs, oc)
    peak_pages = cs.page_count()
    alive = np.ones(n, bool)
    for step in range(25):
        live = np.nonzero(alive)[0]
        kill = rng.choice(live, min(len(live), int(rng.integers(50, 700))), replace=False).astype(np.int32)
        cs.remove(kill); oc.remove(kill)
        alive[kill] = False
        _same_state(cs, oc)
        dead = np.nonzero(~alive)[0]
        back = rng.choice(dead, int(rng.integers(1, len(dead) + 1)), replace=False).astype(np.int32)
        cs.add(back, types[back], pos[back], rad[back]); oc.add(back, types[back], pos[back], rad[back])
        alive[back] = True
        live = np.nonzero(alive)[0]
        flip = rng.choice(live, min(len(live), 60), replace=False).astype(np.int32)
        newr = np.where(rng.random(len(flip)) < 0.5, 450.0, 2.0).astype(np.float32)  # > cell size = is_big chain (culling_system.cpp:139)
        cs.setRadius(flip, newr); oc.set_radius(flip, newr)
        rad[flip] = newr
        _same_state(cs, oc)
        assert cs.entity_count() == int(alive.sum())
        peak_pages = max(peak_pages, cs.page_count())
    ids = cs.page_ids()
    assert len(set(ids.tolist())) == len(ids) and ids.max() < peak_pages + 8  # freed ids come back instead of growing the device arrays


def test_parallel_set_of_distinct_entities_equals_sequential(oracle):
    """lb200_culling_set_many_unique: in-cell movers are overwritten on all host cores, the rest re-bin sequentially; the state after it
    must be the one the one-by-one loop (and the oracle) produces — including entities that a later swap-with-last relocates."""
    import time
    rng = np.random.default_rng(5)
    n = 400_000
    scene = scenes.cull_scene(n, (3000.0, 300.0, 3000.0), seed=9, big_fraction=0.002, type_probs=(0.6, 0.4))
    a, b, oc = lb.CullingSystem(None), lb.CullingSystem(None), oracle.OracleCulling()
    for c in (a, b, oc):
        c.add(scene["entities"], scene["types"], scene["pos"], scene["radius"])
    pos, rad = scene["pos"].copy(), scene["radius"].copy()
    for step in range(3):
        ids = rng.permutation(n)[: 300_000].astype(np.int32)  # distinct
        step_len = np.where(rng.random(len(ids)) < 0.9, 2.0, 250.0)[:, None]  # 10 % are likely to leave their cell
        pos[ids] = pos[ids] + rng.normal(size=(len(ids), 3)) * step_len
        flip = rng.random(len(ids)) < 0.01
        rad[ids[flip]] = np.where(rad[ids[flip]] > 300.0, 3.0, 450.0).astype(np.float32)  # crosses the is_big threshold
        t0 = time.perf_counter(); a.set(ids, pos[ids], rad[ids], unique=True); t_par = time.perf_counter() - t0
        t0 = time.perf_counter(); b.set(ids, pos[ids], rad[ids]); t_seq = time.perf_counter() - t0
        oc.set(ids, pos[ids], rad[ids])
        _same_state(a, oc)
        _same_state(b, oc)
        print(f"set of {len(ids)} movers: parallel {t_par * 1e3:.1f} ms, sequential {t_seq * 1e3:.1f} ms")
    # small batches take the sequential path inside the same entry point
    ids = rng.permutation(n)[:1000].astype(np.int32)
    pos[ids] += 50.0
    a.set(ids, pos[ids], rad[ids], unique=True); oc.set(ids, pos[ids], rad[ids])
    _same_state(a, oc)


def test_threshold_values_bin_like_the_oracle(oracle):
    """radius == 300 is not big, the next float is (culling_system.cpp:139); positions on cell borders, +-0.0 and within 1e-8 of a border
    land in the oracle's cells (the oracle itself is checked against the reference build with the same values, test_oracle_ref.py)."""
    r_edge = np.float32(300.0)
    radii = np.array([r_edge, np.nextafter(r_edge, np.float32(1e9)), np.nextafter(r_edge, np.float32(0)), 0.0, 1e-30], np.float32)
    coords = [0.0, -0.0, 300.0, -300.0, 299.99999999, -299.99999999, 300.00000001, 600.0, -600.0, 899.9999, 1e-300, -1e-300]
    pos = np.array([[x, y, 10.0] for x in coords for y in (0.0, -300.0, 299.99999999)], np.float64)
    P = np.repeat(pos, len(radii), axis=0)
    R = np.tile(radii, len(pos))
    E = np.arange(len(P), dtype=np.int32)
    T = (E % 3).astype(np.uint8)
    cs, oc = lb.CullingSystem(None), oracle.OracleCulling()
    cs.add(E, T, P, R); oc.add(E, T, P, R)
    _same_state(cs, oc)
    big = E[R > r_edge]
    cs.setRadius(big, np.full(len(big), r_edge, np.float32)); oc.set_radius(big, np.full(len(big), r_edge, np.float32))
    edge = E[R == r_edge]
    up = np.full(len(edge), np.nextafter(r_edge, np.float32(1e9)), np.float32)
    cs.setRadius(edge, up); oc.set_radius(edge, up)
    _same_state(cs, oc)
