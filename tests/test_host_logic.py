"""CPU checks of host-side arithmetic that needs no device: percentile interpolation (numpy's
"linear" method, reference infer/calculators.py:697-699, 970-972), the scan interpolation of
infer/intervals/upper_limits.py:16-18, shard bookkeeping."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("n", [1, 2, 7, 200, 1001])
def test_percentiles_match_numpy_linear(n):
    from pyhf_b200.batched import _NORMAL_PERCENTILES, _percentiles_sorted

    rng = np.random.default_rng(n)
    x = np.sort(rng.exponential(size=n))
    got = _percentiles_sorted(torch.as_tensor(x), _NORMAL_PERCENTILES).numpy()
    np.testing.assert_allclose(got, np.percentile(x, _NORMAL_PERCENTILES), rtol=1e-14, atol=0)
    qs = [0.0, 100.0, 33.3]
    np.testing.assert_allclose(_percentiles_sorted(torch.as_tensor(x), qs).numpy(), np.percentile(x, qs), rtol=1e-14)


def test_upper_limit_interpolation_matches_reference_formula():
    from pyhf_b200.batched import upper_limit_from_scan

    mus = np.linspace(0, 5, 41)
    cls = np.exp(-1.3 * mus)
    want = np.interp(0.05, cls[::-1], mus[::-1])  # upper_limits.py:16-18 on a falling CLs curve
    assert upper_limit_from_scan(mus, cls) == pytest.approx(want, rel=1e-15)
    band = np.stack([cls * f for f in (0.5, 0.8, 1.0, 1.2, 1.5)], axis=1)
    got = upper_limit_from_scan(torch.as_tensor(mus), torch.as_tensor(band))
    assert got == [pytest.approx(np.interp(0.05, band[::-1, i], mus[::-1])) for i in range(5)]


def test_shard_ranges_tile_the_toys():
    from pyhf_b200.dist import shard_range

    for n in (0, 1, 7, 100000, 100003):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


@pytest.mark.needs_pyhf
def test_toy_batches_are_tagged_where_they_are_created():
    """Simultaneous.sample (probability.py:263-274) under the torch carrier: every row of the toy
    matrix is recognised as row i of ONE tagged batch (through concatenate / einsum / gather of
    _TensorViewer.stitch, tensor/common.py:40-52); rows of a user's own matrix are not"""
    import pyhf

    from pyhf_b200 import synth
    from pyhf_b200.backend import TorchCarrierBackend
    from pyhf_b200.optimizer import B200Optimizer

    tb = TorchCarrierBackend()
    pyhf.set_backend(tb)
    try:
        for spec in (synth.spec_hello(), synth.spec_c3()):
            model = pyhf.Model(spec, poi_name="mu", validate=False)
            D = model.config.nmaindata + model.config.nauxdata
            toys = model.make_pdf(tb.astensor(model.config.suggested_init())).sample((7,))
            got = [B200Optimizer._toy_batch_of(r, D) for r in toys]
            assert all(x is not None for x in got)
            assert [x[2] for x in got] == list(range(7))
            assert len({x[0] for x in got}) == 1            # one batch key
            assert torch.equal(got[3][1][3], toys[3])
            user = torch.as_tensor(np.random.rand(5, D))
            assert B200Optimizer._toy_batch_of(user[2], D) is None
    finally:
        pyhf.set_backend("numpy", "scipy")


def test_dmma_table_row_layout_is_conflict_free_for_both_products():
    """HF_MMA_ROW (csrc/hf_internal.cuh): rows of a table chunk sit at stride 72 doubles with a
    4-double swizzle.  Shared memory serves a 16-byte load per quarter warp (8 lanes x 2 doubles = the
    16 doubles = 32 banks of one wavefront), so a load is conflict-free iff its 8 lanes hit 8 different
    2-double slots mod 16.  Forward product: lanes (g in {0,1}, t in 0..3) read row k4*4 + t, cells
    2g, 2g+1; backward product: lanes read row 8ni + g, cells 8j + 2t, +1 -- for every quarter warp."""
    import os
    import re

    src = open(os.path.join(os.path.dirname(__file__), "..", "pyhf_b200", "csrc", "hf_internal.cuh")).read()
    csh = int(re.search(r"#define HF_MMA_CSH (\d+)", src).group(1))
    kc = int(re.search(r"#define HF_MMA_KC (\d+)", src).group(1))
    ncell = int(re.search(r"#define HF_MMA_NCH (\d+)", src).group(1))
    assert "#define HF_MMA_ROW(r) ((r) * HF_MMA_CSH + 4 * (((r) >> 1) & 1))" in src

    def row(r):
        return r * csh + 4 * ((r >> 1) & 1)

    # rows do not overlap and stay 16-byte aligned
    for r in range(kc):
        assert row(r) % 2 == 0 and row(r) + ncell <= (row(r + 1) if r + 1 < kc else kc * csh)

    def slots(addrs):  # 2-double slots mod 16 doubles touched by a quarter warp
        return sorted((a % 16) // 2 for a in addrs)

    for quarter in range(4):  # lanes 8q .. 8q+7: g = lane >> 2, t = lane & 3
        lanes = [(lane >> 2, lane & 3) for lane in range(8 * quarter, 8 * quarter + 8)]
        for k4 in range(kc // 4):
            for hw in range(4):
                fwd = [row(4 * k4 + t) + 16 * hw + 2 * g for g, t in lanes]
                assert slots(fwd) == list(range(8)), ("forward", quarter, k4, hw, slots(fwd))
        for ni in range(kc // 8):
            for j in range(ncell // 8):
                bwd = [row(8 * ni + g) + 8 * j + 2 * t for g, t in lanes]
                assert slots(bwd) == list(range(8)), ("backward", quarter, ni, j, slots(bwd))


def test_dmma_coefficient_block_swizzle_is_conflict_free():
    """coef_at (csrc/hf_eval_mma.cu): the forward product's A operand is column-major [k][32 points]
    with the 16 point pairs of a column XOR-swizzled by 2*(k & 3); a lane (g, t) reads the pairs g and
    8+g of column k4*4 + t with 16-byte loads -- every quarter warp must hit 8 different 2-double slots,
    and the map must be a bijection on a column."""
    import os

    src = open(os.path.join(os.path.dirname(__file__), "..", "pyhf_b200", "csrc", "hf_eval_mma.cu")).read()
    assert "return k * MM + ((((p >> 1) ^ (2 * (k & 3))) << 1) | (p & 1));" in src
    MM = 32

    def coef_at(k, p):
        return k * MM + ((((p >> 1) ^ (2 * (k & 3))) << 1) | (p & 1))

    for k in range(8):
        assert sorted(coef_at(k, p) for p in range(MM)) == list(range(k * MM, (k + 1) * MM))
    for quarter in range(4):
        lanes = [(lane >> 2, lane & 3) for lane in range(8 * quarter, 8 * quarter + 8)]
        for k4 in range(3):
            for base_pair in (0, 8):  # m-tiles 0/1 and 2/3
                addrs = [coef_at(4 * k4 + t, 2 * (base_pair + g)) for g, t in lanes]
                assert all(a % 2 == 0 for a in addrs)
                assert sorted((a % 16) // 2 for a in addrs) == list(range(8)), (quarter, k4, base_pair)
                # the second element of the 16-byte load is the odd point of the pair
                assert all(coef_at(4 * k4 + t, 2 * (base_pair + g) + 1) == a + 1 for (g, t), a in zip(lanes, addrs))


def test_dmma_delta_tile_row_layout_is_conflict_free():
    """HF_MMA_WFROW (csrc/hf_internal.cuh): rows of the Delta / w*F tile [32 points][128 cells] at stride
    136 doubles with a 2-double swizzle.  Three access patterns, all 16-byte, must be conflict-free per
    quarter warp: the epilogue (lane = point, one column), the A' fragments of the backward product
    (lanes (g, t): row 8hw + g, cells 64h + 8j + 2t) and the Delta store (lanes (g, t): row
    16(mi>>1) + 2g + (mi&1), column 4(4w + t))."""
    import os
    import re

    src = open(os.path.join(os.path.dirname(__file__), "..", "pyhf_b200", "csrc", "hf_internal.cuh")).read()
    cs = int(re.search(r"#define HF_MMA_CS (\d+)", src).group(1))
    assert "#define HF_MMA_WFROW(p) ((p) * HF_MMA_CS + 2 * (((p) >> 1) & 3))" in src

    def row(p):
        return p * cs + 2 * ((p >> 1) & 3)

    for p in range(32):
        assert row(p) % 2 == 0 and row(p) + 128 <= (row(p + 1) if p < 31 else 32 * cs)

    def ok(addrs):
        return sorted((a % 16) // 2 for a in addrs) == list(range(8))

    for quarter in range(4):
        pts = list(range(8 * quarter, 8 * quarter + 8))
        lanes = [(lane >> 2, lane & 3) for lane in pts]
        for col in range(0, 128, 2):  # epilogue: both 16-byte halves of every double4
            assert ok([row(p) + col for p in pts]), ("epilogue", quarter, col)
        for hw in range(4):
            for half in range(2):
                for j in range(8):
                    assert ok([row(8 * hw + g) + 64 * half + 8 * j + 2 * t for g, t in lanes]), ("backward A'", quarter, hw, j)
        for mi in range(4):
            for warp in range(8):
                for e in (0, 2):
                    a = [row(16 * (mi >> 1) + 2 * g + (mi & 1)) + 4 * (4 * warp + t) + e for g, t in lanes]
                    assert ok(a), ("delta store", quarter, mi, warp, e)
