"""-m gpu: the neighbourhood-cache convolution (csrc/nbplan.cu + csrc/conv_nb.cu, lgs_conv_fwd3) against the table-driven
bf16x3 kernel (lgs_conv_fwd2), the exact SIMT kernel and — through the facade — the oracle.

Bars: the plan is integer work -> its invariants are exact (order is a permutation, uniq[loc] reproduces the table entry for
entry); features: nb vs bx3 differ only in fp32 summation order (channel block outer vs offset outer) -> 3e-5 (two bf16x3 results each ~1e-5 from exact); vs SIMT / oracle
1e-4 (the 'bx3' tolerance of test_gpu_conv.py)."""
import ctypes

import numpy as np
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E(lib):
    from languagegroundedsemseg_b200 import minkowski
    return minkowski


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _scene_map(E, lib, n_target, seed, min_rows=1000):
    """coordinate map + 3^3 same-map kernel map (with its neighbourhood plan) of a synthetic scene"""
    from languagegroundedsemseg_b200 import scenes
    assert lib.lgs_tune(b"nb_min_rows", min_rows) == 0
    coords, _, _ = scenes.synthetic_voxel_scene(seed=seed, target_voxels=n_target)
    c = torch.from_numpy(coords).cuda()
    st = E.SparseTensor(torch.zeros(c.shape[0], 4, device="cuda"), c)
    mgr, key = st.coordinate_manager, st.coordinate_map_key
    km = mgr.kernel_map(key, key, [3] * 3, [1] * 3)
    return coords, km


def _geometry(lib, n, K=27):
    g = (ctypes.c_int64 * 9)()
    assert lib.lgs_nbplan_geometry(n, K, g) == 0
    return dict(zip(["tm", "rt", "RS", "S", "umax", "o_order", "o_ucount", "o_uniq", "o_loc"], [int(v) for v in g]))


def _plan_arrays(lib, km, plan=None):
    g = _geometry(lib, km.n_out, km.K)
    p = (km.plan if plan is None else plan).cpu().numpy()
    S, RS, K, umax = g["S"], g["RS"], km.K, g["umax"]
    order = p[g["o_order"]: g["o_order"] + S * RS].reshape(S, RS)
    ucount = p[g["o_ucount"]: g["o_ucount"] + S]
    uniq = p[g["o_uniq"]: g["o_uniq"] + S * umax].reshape(S, umax)
    loc = p[g["o_loc"]:].view(np.uint16)[: S * K * RS].reshape(S, K, RS)
    return g, order, ucount, uniq, loc


@pytest.mark.parametrize("n_target,seed", [(25000, 3), (60000, 5)])
def test_nbplan_invariants(E, lib, n_target, seed):
    """integer gate: every output row appears exactly once in `order`; uniq[s][loc[s][k][q] & 0xFFF] == table[k][order[s][q]]
    for every slot (index 0xFFF <-> -1); unique rows of a supertile are distinct; no overflow on ScanNet-shaped scenes;
    colours: loc >> 12 of an existing neighbour equals the colour stored with the cached row (uniq >> 28) and
    (x + 3y + 5z) mod 8 of its coordinates; quarter-warps (aligned groups of 8 slots of a tile) read 8 different colours
    for every offset in nearly all cases (the rest are 2-way bank conflicts, not errors)"""
    coords, km = _scene_map(E, lib, n_target, seed)
    try:
        assert km.plan is not None, km.plan_stats
        g, order, ucount, uniq, loc = _plan_arrays(lib, km)
        n = km.n_out
        flat = order.reshape(-1)
        valid = flat >= 0
        assert valid.sum() == n and np.array_equal(np.sort(flat[valid]), np.arange(n))
        assert (order[:-1] >= 0).all()                  # padding may only sit in the last supertile
        table = km.fwd_table.cpu().numpy()              # [K, n]
        assert km.plan_stats[0] == 0 and km.plan_stats[1] == ucount.max() <= g["umax"]
        assert g["rt"] % 8 == 0
        K, n_oct, n_free = km.K, 0, 0
        for s in range(g["S"]):
            o = order[s]
            u = uniq[s, : ucount[s]] & 0x0FFFFFFF
            ucol = uniq[s, : ucount[s]] >> 28
            assert len(np.unique(u)) == len(u)
            assert np.array_equal(ucol, (coords[u, 1] + 3 * coords[u, 2] + 5 * coords[u, 3]) & 7)
            lc = loc[s].astype(np.int64) & 0xFFF        # [K, RS]
            col = loc[s].astype(np.int64) >> 12
            exp = np.where(o[None, :] >= 0, table[:, np.maximum(o, 0)], -1)
            got = np.where(lc == 0xFFF, -1, u[np.minimum(lc, max(len(u) - 1, 0))])
            assert np.array_equal(got, exp), s
            assert (lc[lc != 0xFFF] < ucount[s]).all()
            assert col.max() <= 7
            hit = lc != 0xFFF
            assert np.array_equal(col[hit], ucol[lc[hit]])
            octs = col.reshape(K, g["tm"], g["rt"] // 8, 8)
            distinct = (np.sort(octs, -1)[..., 1:] != np.sort(octs, -1)[..., :-1]).all(-1)
            n_oct += distinct.size
            n_free += int(distinct.sum())
        gathered = (table >= 0).sum() / g["S"]
        print(f"[nbplan n={n}] S={g['S']} RS={g['RS']} rt={g['rt']} unique rows per supertile: mean {ucount.mean():.0f} max {ucount.max()} "
              f"(cache {g['umax']}); table-driven gather would copy {gathered:.0f} rows per supertile; "
              f"conflict-free quarter-warp reads {n_free / n_oct:.1%}")
        assert n_free / n_oct > 0.7
        # a second build gives the same supertiles and the same unique-row SETS (local numbering follows shared-memory hash
        # slots, whose occupancy under linear probing depends on arrival order; the convolution's result does not)
        plan2 = torch.empty_like(km.plan)
        scratch = torch.empty(lib.lgs_nbplan_scratch_bytes(n) // 4, dtype=torch.int32, device="cuda")
        st = (ctypes.c_int32 * 2)()
        from languagegroundedsemseg_b200 import _lib
        cm_coords = torch.from_numpy(coords).cuda()
        _lib.check(lib.lgs_nbplan_build(_lib.ptr(cm_coords), n, _lib.ptr(km.fwd_table), km.K, 1, _lib.ptr(plan2), _lib.ptr(scratch),
                                        ctypes.cast(st, ctypes.c_void_p), _stream()))
        _, order2, ucount2, uniq2, loc2 = _plan_arrays(lib, km, plan2)
        assert np.array_equal(order2, order) and np.array_equal(ucount2, ucount)
        assert all(np.array_equal(np.sort(uniq2[s, : ucount[s]]), np.sort(uniq[s, : ucount[s]])) for s in range(g["S"]))
        assert np.array_equal(loc2 & 0xFFF == 0xFFF, loc & 0xFFF == 0xFFF) and np.array_equal(loc2 >> 12, loc >> 12)
    finally:
        lib.lgs_tune(b"nb_min_rows", 0)


def _operands(lib, w):
    from languagegroundedsemseg_b200 import _lib
    K, c_in, c_out = w.shape
    fwd = torch.empty(lib.lgs_weight_bx3_elems(K, c_out, c_in), dtype=torch.bfloat16, device="cuda")
    bwd = torch.empty(lib.lgs_weight_bx3_elems(K, c_in, c_out), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.lgs_weight_prep(_lib.ptr(w), K, c_in, c_out, 3, _lib.ptr(fwd), _lib.ptr(bwd), _lib.F32, _stream()))
    return fwd, bwd


def _conv3(lib, x, x2, w_op, K, c_out, km, plan, reverse, bias, stats, out=None):
    from languagegroundedsemseg_b200 import _lib
    n = km.n_out
    if out is None:
        out = torch.full((n, c_out), float("nan"), device="cuda")
    c1, c2 = x.shape[1], (x2.shape[1] if x2 is not None else 0)
    _lib.check(lib.lgs_conv_fwd3(_lib.ptr(x), c1, _lib.ptr(x2), c2, n, _lib.ptr(w_op), K, c_out, _lib.ptr(km.fwd_table),
                                 _lib.ptr(plan), n, reverse, _lib.ptr(bias), _lib.ptr(out), _lib.ptr(stats), _stream()))
    return out


SHAPES = [
    # c_in, c_in2, c_out, bias
    (32, 0, 32, False),
    (96, 0, 96, False),
    (96, 32, 96, False),      # block8's first convolution as a two-source gather
    (96, 0, 128, True),       # two output-channel slices (dgrad of a 128 -> 96 layer)
    (64, 0, 48, False),
    (20, 0, 16, False),       # ragged input width (one partial channel block)
    (256, 32, 512, False),    # Res16UNet34D block8: six output-channel slices
]


@pytest.mark.parametrize("c_in,c_in2,c_out,bias", SHAPES)
@pytest.mark.parametrize("reverse", [0, 1])
def test_conv_nb_vs_table_driven_and_simt(E, lib, c_in, c_in2, c_out, bias, reverse):
    from languagegroundedsemseg_b200 import _lib
    coords, km = _scene_map(E, lib, 30000, 7)
    try:
        assert km.plan is not None
        n, K = km.n_out, 27
        torch.manual_seed(c_in * 7 + c_out + reverse)
        x = torch.randn(n, c_in, device="cuda")
        x2 = torch.randn(n, c_in2, device="cuda") if c_in2 else None
        w = torch.randn(K, c_in + c_in2, c_out, device="cuda") / np.sqrt(K * (c_in + c_in2))
        b = torch.randn(1, c_out, device="cuda") if bias else None
        w_fwd, _ = _operands(lib, w)
        s_nb = torch.zeros(8, 2, c_out, dtype=torch.float64, device="cuda")
        s_tb = torch.zeros_like(s_nb)
        y_nb = _conv3(lib, x, x2, w_fwd, K, c_out, km, km.plan, reverse, b, s_nb)
        y_tb = _conv3(lib, x, x2, w_fwd, K, c_out, km, None, reverse, b, s_tb)      # NULL plan = lgs_conv_fwd2
        assert not torch.isnan(y_nb).any()
        e_tb = rel_err(y_nb, y_tb)
        # exact fp32 reference: the SIMT kernel on the concatenated input
        xc = torch.cat([x, x2], 1) if c_in2 else x
        y_ref = torch.empty(n, c_out, device="cuda")
        _lib.check(lib.lgs_conv_fwd(_lib.ptr(xc), n, c_in + c_in2, _lib.ptr(w), _lib.W_KCN, K, c_out, _lib.ptr(km.fwd_table), n,
                                    reverse, _lib.ptr(b), _lib.ptr(y_ref), _lib.F32, _lib.ALGO_SIMT, _stream()))
        e_ref = rel_err(y_nb, y_ref)
        sums = s_nb.sum(0).cpu()
        e_s1 = rel_err(sums[0], y_nb.double().sum(0).cpu())
        e_s2 = rel_err(sums[1], (y_nb.double() ** 2).sum(0).cpu())
        print(f"[conv_nb {c_in}+{c_in2}->{c_out} rev={reverse}] vs table-driven bx3 {e_tb:.1e}, vs SIMT fp32 {e_ref:.1e}, "
              f"fused BN sums {e_s1:.1e} / {e_s2:.1e}")
        assert e_tb < 3e-5 and e_ref < 1e-4
        assert e_s1 < 1e-5 and e_s2 < 1e-5
    finally:
        lib.lgs_tune(b"nb_min_rows", 0)


@pytest.mark.parametrize("n_target,c_in,c_out", [(450, 256, 256), (2200, 256, 256), (2200, 128, 128), (8500, 64, 64), (8500, 192, 128)])
def test_conv_nb_small_map_split(E, lib, n_target, c_in, c_out):
    """coarse U-Net levels: fewer supertiles than SMs -> the reduction over channel blocks / offsets is split over CTAs
    (red.global.add on a zeroed output).  Against the table-driven bx3 kernel and exact SIMT fp32; prints both times."""
    from languagegroundedsemseg_b200 import _lib
    coords, km = _scene_map(E, lib, n_target, 11, min_rows=0)
    assert km.plan is not None, km.plan_stats
    n, K = km.n_out, 27
    torch.manual_seed(n_target + c_in)
    x = torch.randn(n, c_in, device="cuda")
    w = torch.randn(K, c_in, c_out, device="cuda") / np.sqrt(K * c_in)
    b = torch.randn(1, c_out, device="cuda")
    w_fwd, _ = _operands(lib, w)
    outs, t = {}, {}
    for name, plan in (("nb", km.plan), ("table", None)):
        for rev in (0, 1):
            outs[name, rev] = _conv3(lib, x, None, w_fwd, K, c_out, km, plan, rev, b, None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o = torch.empty(n, c_out, device="cuda")
        for _ in range(3):
            _conv3(lib, x, None, w_fwd, K, c_out, km, plan, 0, b, None, o)
        e0.record()
        for _ in range(20):
            _conv3(lib, x, None, w_fwd, K, c_out, km, plan, 0, b, None, o)
        e1.record()
        torch.cuda.synchronize()
        t[name] = e0.elapsed_time(e1) / 20 * 1e3
    y_ref = torch.empty(n, c_out, device="cuda")
    _lib.check(lib.lgs_conv_fwd(_lib.ptr(x), n, c_in, _lib.ptr(w), _lib.W_KCN, K, c_out, _lib.ptr(km.fwd_table), n,
                                0, _lib.ptr(b), _lib.ptr(y_ref), _lib.F32, _lib.ALGO_SIMT, _stream()))
    e_tb = max(rel_err(outs["nb", r], outs["table", r]) for r in (0, 1))
    e_ref = rel_err(outs["nb", 0], y_ref)
    print(f"[small map {n} rows {c_in}->{c_out}] nb {t['nb']:.1f} us, table-driven {t['table']:.1f} us; vs table-driven {e_tb:.1e}, vs SIMT {e_ref:.1e}; "
          f"unique rows max {km.plan_stats[1]}")
    assert not torch.isnan(outs["nb", 0]).any()
    assert e_tb < 3e-5 and e_ref < 1e-4


@pytest.mark.parametrize("n_target,c", [(30000, 96), (2200, 256), (450, 128)])
def test_conv_fwd4_addend(E, lib, n_target, c):
    """lgs_conv_fwd4: out = conv + addend in the epilogue (large map: plain stores; small maps: the split whose blockIdx.z
    is 0 adds it) equals lgs_conv_fwd3 followed by an add; with a NULL plan the entry runs the table-driven kernel + lgs_add"""
    from languagegroundedsemseg_b200 import _lib
    coords, km = _scene_map(E, lib, n_target, 13, min_rows=0)
    assert km.plan is not None, km.plan_stats
    n, K = km.n_out, 27
    torch.manual_seed(n_target)
    x = torch.randn(n, c, device="cuda")
    w = torch.randn(K, c, c, device="cuda") / np.sqrt(K * c)
    add = torch.randn(n, c, device="cuda")
    w_fwd, _ = _operands(lib, w)
    ref = _conv3(lib, x, None, w_fwd, K, c, km, km.plan, 1, None, None) + add
    for plan in (km.plan, None):
        out = torch.full((n, c), float("nan"), device="cuda")
        _lib.check(lib.lgs_conv_fwd4(_lib.ptr(x), c, None, 0, n, _lib.ptr(w_fwd), K, c, _lib.ptr(km.fwd_table), _lib.ptr(plan), n, 1,
                                     None, _lib.ptr(add), _lib.ptr(out), None, _stream()))
        assert rel_err(out, ref) < 3e-5, (plan is None, rel_err(out, ref))


def test_plan_overflow_falls_back_to_table_driven(E, lib):
    """a cache too small for a supertile's unique rows (lgs_tune nb_umax) makes the builder flag the plan; the facade then
    keeps the table-driven kernel — same result, no error"""
    from languagegroundedsemseg_b200 import scenes
    assert lib.lgs_tune(b"nb_umax", 64) == 0
    try:
        coords, _, _ = scenes.synthetic_voxel_scene(seed=4, target_voxels=9000)
        c = torch.from_numpy(coords).cuda()
        st = E.SparseTensor(torch.zeros(c.shape[0], 4, device="cuda"), c)
        km = st.coordinate_manager.kernel_map(st.coordinate_map_key, st.coordinate_map_key, [3] * 3, [1] * 3)
        assert km.plan is None and km.plan_stats[0] == 1 and km.plan_stats[1] > 64, km.plan_stats
        torch.manual_seed(0)
        x = torch.randn(c.shape[0], 32, device="cuda")
        w = torch.randn(27, 32, 32, device="cuda") / 30
        E.set_conv_algo("bx3")
        y = E.sparse_conv(x, w, None, km)
        E.set_conv_algo("simt")
        y_ref = E.sparse_conv(x, w, None, km)
        E.set_conv_algo("bx3")
        assert rel_err(y, y_ref) < 1e-4
    finally:
        lib.lgs_tune(b"nb_umax", 0)


def test_facade_layers_with_plan_vs_oracle(E, lib):
    """two stacked 3^3 convolutions fwd + bwd through the facade with plans on (small-map threshold lowered) vs the oracle"""
    from oracle import me_cpu
    from languagegroundedsemseg_b200 import scenes
    assert lib.lgs_tune(b"nb_min_rows", 1000) == 0
    try:
        E.set_conv_algo("bx3")
        coords, _, _ = scenes.synthetic_voxel_scene(seed=9, target_voxels=12000)
        res = {}
        torch.manual_seed(5)
        f0 = torch.randn(coords.shape[0], 32)
        for name, eng, dev in (("oracle", me_cpu, "cpu"), ("cuda", E, "cuda")):
            torch.manual_seed(6)
            net = torch.nn.Sequential(eng.MinkowskiConvolution(32, 96, kernel_size=3, dimension=3),
                                      eng.MinkowskiConvolution(96, 64, kernel_size=3, dimension=3)).to(dev)
            f = f0.clone().to(dev).requires_grad_(True)
            x = eng.SparseTensor(f, torch.from_numpy(coords).to(dev))
            y = net(x)
            if name == "cuda":
                km = x.coordinate_manager.kernel_map(x.coordinate_map_key, x.coordinate_map_key, [3] * 3, [1] * 3)
                assert km.plan is not None, km.plan_stats
            torch.manual_seed(7)
            y.F.backward(torch.randn(y.F.shape).to(dev))
            res[name] = (y.F.detach().cpu(), f.grad.cpu(), net[0].kernel.grad.cpu())
        o, g = res["oracle"], res["cuda"]
        print(f"[facade + plan] out {rel_err(g[0], o[0]):.1e} gin {rel_err(g[1], o[1]):.1e} gw {rel_err(g[2], o[2]):.1e}")
        assert rel_err(g[0], o[0]) < 1e-4 and rel_err(g[1], o[1]) < 1e-4 and rel_err(g[2], o[2]) < 1e-3
    finally:
        lib.lgs_tune(b"nb_min_rows", 0)


def test_conv_nb_full_size_and_speed(E, lib):
    """BASELINE configs[1] size (149 106 voxels): 96 -> 96 and 96 + 32 -> 96 at level 0, nb vs table-driven; prints both times
    (CUDA events, 10 launches each after 3 warm-up launches; the numbers are informative, the gate is the result)"""
    coords, km = _scene_map(E, lib, 150000, 0, min_rows=0)
    assert km.plan is not None, km.plan_stats
    n, K = km.n_out, 27
    for c_in, c_in2, c_out in ((96, 0, 96), (96, 32, 96), (32, 0, 32)):
        torch.manual_seed(1)
        x = torch.randn(n, c_in, device="cuda")
        x2 = torch.randn(n, c_in2, device="cuda") if c_in2 else None
        w = torch.randn(K, c_in + c_in2, c_out, device="cuda") / np.sqrt(K * (c_in + c_in2))
        w_fwd, _ = _operands(lib, w)
        t = {}
        outs = {}
        for name, plan in (("nb", km.plan), ("table", None)):
            outs[name] = torch.full((n, c_out), float("nan"), device="cuda")
            for _ in range(3):
                _conv3(lib, x, x2, w_fwd, K, c_out, km, plan, 0, None, None, outs[name])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                _conv3(lib, x, x2, w_fwd, K, c_out, km, plan, 0, None, None, outs[name])
            e1.record()
            torch.cuda.synchronize()
            t[name] = e0.elapsed_time(e1) / 10 * 1e3
        e = rel_err(outs["nb"], outs["table"])
        print(f"[full size {c_in}+{c_in2}->{c_out}, {n} rows] nb {t['nb']:.0f} us, table-driven {t['table']:.0f} us, diff {e:.1e}; "
              f"plan max unique rows {km.plan_stats[1]}")
        assert e < 3e-5
