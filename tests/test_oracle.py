"""CPU: pin the oracle (oracle/cbinfer_oracle.c + oracle/oracle.py) against the reference's own
golden vectors (tests/golden/reference_golden.npz, produced by tests/golden/make_golden.py from
the reference's python twins and native conv2d_fg_cpu) and its KATs."""
import numpy as np
import pytest

CASES = ["a", "b", "c", "d"]


def kat_input(seed=1234):
    # same generator as tests/golden/make_golden.py::kat_input (conv2d_cg.py:84-97 genTestData)
    rs = np.random.RandomState(seed)
    inp = rs.randn(1, 16, 400, 300).astype(np.float32)
    prev = inp.copy()
    for (c, y, x, d) in ((0, 0, 4, 1.00), (1, 6, 9, 0.05), (2, 10, 4, -11.00), (1, 6, 19, -0.05)):
        prev[0, c, y, x] += np.float32(d)
    return inp, prev, 0.1, (3, 3)


def test_kat1_change_detection(orc, golden):
    inp, prev, thr, fs = kat_input(int(golden["kat1_seed"]))
    cmap, raw = orc.changeDetection(inp, prev.copy(), fs, thr, return_raw=True)
    idx = orc.changeIndexesExtr(cmap)
    expect = [3, 4, 5, 303, 304, 305, 2703, 2704, 2705, 3003, 3004, 3005, 3303, 3304, 3305]
    assert idx.tolist() == expect == golden["kat1_map_idx"].tolist()
    assert orc.changeIndexesExtr(raw).tolist() == golden["kat1_raw_idx"].tolist()
    prop = orc.changePropagation(raw, fs)
    assert orc.changeIndexesExtr(prop).tolist() == golden["kat1_prop_idx"].tolist()
    X = orc.genXMatrix(inp, idx, fs)
    assert np.array_equal(X, golden["kat1_X"])


def test_kat2_change_indexes(orc, golden):
    cm = np.zeros((129, 254), dtype=np.uint8)
    for (y, x) in ((3, 3), (7, 5), (5, 7), (7, 1), (1, 5), (24, 31)):
        cm[y, x] = 1
    assert orc.changeIndexesExtr(cm).tolist() == [259, 765, 1277, 1779, 1783, 6127]
    assert golden["kat2_idx"].tolist() == [259, 765, 1277, 1779, 1783, 6127]


@pytest.mark.parametrize("tag", CASES)
def test_python_twins(orc, golden, tag):
    C, H, W, kH, kW, Cout, relu = [int(v) for v in golden[f"{tag}_params"]]
    thr = float(golden[f"{tag}_thr"])
    f0, f1 = golden[f"{tag}_f0"], golden[f"{tag}_f1"]
    cmap, raw = orc.changeDetection(f1, f0.copy(), (kH, kW), thr, return_raw=True)
    assert np.array_equal(cmap, golden[f"{tag}_map"])
    assert np.array_equal(raw, golden[f"{tag}_raw"])
    assert np.array_equal(orc.changePropagation(raw, (kH, kW)), golden[f"{tag}_prop"])
    idx = orc.changeIndexesExtr(cmap)
    assert np.array_equal(idx, golden[f"{tag}_idx"])
    X = orc.genXMatrix(f1, idx, (kH, kW))
    assert np.array_equal(X, golden[f"{tag}_X"])                       # exact: a copy
    Y = orc.matrixMult(X, golden[f"{tag}_w"], golden[f"{tag}_b"])
    # reference GEMM is fp32 torch.matmul (summation order unpinned): tolerance, stated
    np.testing.assert_allclose(Y, golden[f"{tag}_Y"], rtol=1e-5, atol=1e-5)
    po = golden[f"{tag}_prevOut"].copy()
    out = orc.updateOutput(np.ascontiguousarray(golden[f"{tag}_Y"].T), idx, po, withReLU=bool(relu))
    assert np.array_equal(out, golden[f"{tag}_out"])                   # exact given the same Y


@pytest.mark.parametrize("tag", CASES)
def test_module_flow_matches_twins(orc, golden, tag):
    """OracleCBConv2d (conv2d.py:178-259 restated) == chained python twins."""
    C, H, W, kH, kW, Cout, relu = [int(v) for v in golden[f"{tag}_params"]]
    m = orc.OracleCBConv2d(golden[f"{tag}_w"], golden[f"{tag}_b"], float(golden[f"{tag}_thr"]),
                           withReLU=bool(relu))
    m.forward(golden[f"{tag}_f0"])
    m.prevOutput[...] = golden[f"{tag}_prevOut"]
    out = m.forward(golden[f"{tag}_f1"])
    assert np.array_equal(m.changeIndexes, golden[f"{tag}_idx"])
    np.testing.assert_allclose(out, golden[f"{tag}_out"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("tag", ["fg1", "fg2"])
def test_fg_native_reference(orc, golden, tag):
    """oracle FG == reference native conv2d_fg_cpu (cbconv2d_fg_backend.cu:81-112)."""
    out = golden[f"{tag}_prevOut"].copy()
    orc.cbconvFG(golden[f"{tag}_in"], golden[f"{tag}_prev"], out, golden[f"{tag}_w"],
                 float(golden[f"{tag}_thr"]))
    np.testing.assert_allclose(out, golden[f"{tag}_out"], rtol=1e-5, atol=1e-4)


def test_threshold_zero_is_dense(orc):
    rs = np.random.RandomState(3)
    x0 = rs.randn(1, 3, 10, 12).astype(np.float32)
    x1 = x0 + (rs.rand(1, 3, 10, 12) < 0.2) * rs.randn(1, 3, 10, 12).astype(np.float32)
    x1 = x1.astype(np.float32)
    w = (rs.randn(5, 3, 3, 3) * 0.3).astype(np.float32)
    b = rs.randn(5).astype(np.float32)
    for fb in (False, True):
        m = orc.OracleCBConv2d(w, b, 0.0, withReLU=True, feedbackLoop=fb)
        m.forward(x0)
        out = m.forward(x1)
        np.testing.assert_allclose(out, orc.dense_conv2d(x1, w, b, relu=True), rtol=1e-6, atol=1e-6)


def test_strict_greater_and_ftz(orc):
    """CUDA semantics: strict '>' (cg.cu:56) and flush-to-zero (build.sh:5 --use_fast_math)."""
    x = np.zeros((1, 1, 2, 2), np.float32)
    s = np.zeros((1, 1, 2, 2), np.float32)
    s[0, 0, 0, 0] = 0.5          # |d| == thr -> not changed
    s[0, 0, 0, 1] = 1e-40        # denormal difference, thr 0 -> flushed, not changed
    s[0, 0, 1, 0] = np.nan       # NaN never triggers
    s[0, 0, 1, 1] = np.inf       # fresh state triggers
    assert orc.changeDetection(x, s.copy(), (1, 1), 0.5).tolist() == [[0, 0], [0, 1]]
    assert orc.changeDetection(x, s.copy(), (1, 1), 0.0).tolist() == [[1, 0], [0, 1]]


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_16bit_detection_rule(orc, dt):
    """half.cu:58-63: compare the ROUNDED 16-bit difference against the ROUNDED threshold."""
    import torch
    td = torch.float16 if dt == "f16" else torch.bfloat16
    code = orc.F16 if dt == "f16" else orc.BF16
    g = torch.Generator().manual_seed(5)
    a = torch.randn(1, 4, 6, 7, generator=g).to(td)
    b = (a.float() + 0.3 * torch.randn(1, 4, 6, 7, generator=g)).to(td)
    thr = 0.2
    diff = (a.double() - b.double()).to(td)
    t = torch.tensor(thr, dtype=torch.float64).to(td)
    expect = ((diff > t) | (diff < -t)).any(dim=1)[0].to(torch.uint8).numpy()
    an, _ = orc.from_torch(a)
    bn, _ = orc.from_torch(b)
    got = orc.changeDetection(bn, an.copy(), (1, 1), thr, dtype=code)
    assert np.array_equal(got, expect)


def test_small_float_rounding_exhaustive(orc):
    """the C double->half / double->bf16 RNE helpers agree with numpy/torch on random data."""
    import torch
    rs = np.random.RandomState(0)
    v = np.concatenate([rs.randn(20000) * 10 ** rs.uniform(-9, 5, 20000),
                        [0.0, -0.0, np.inf, -np.inf, 65504.0, 65519.9, 65520.0, 2.0 ** -25,
                         2.0 ** -25 * 1.0001, 2.0 ** -24, 6.1e-5, 3.0e38, 3.4e38, 1e-40, 1e-45]])
    # torch/numpy convert double->float->16-bit (double rounding); the C helper rounds once, so
    # compare on float32-representable inputs where both are well defined and equal.
    v = v.astype(np.float32).astype(np.float64)
    L = orc.lib()
    h = np.array([L.orc_f64_to_f16(float(x)) for x in v], dtype=np.uint16)
    assert np.array_equal(h, v.astype(np.float16).view(np.uint16))
    b = np.array([L.orc_f64_to_bf16(float(x)) for x in v], dtype=np.uint16)
    tb = torch.from_numpy(v).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(b, tb)


def test_pool_flow(orc):
    rs = np.random.RandomState(11)
    x0 = rs.randn(1, 3, 8, 10).astype(np.float32)
    p = orc.OracleCBPoolMax2d()
    full = np.arange(80, dtype=np.int32)
    y0 = p.forward(("changeIndexes", x0, full))
    ref = x0.reshape(1, 3, 4, 2, 5, 2).max(axis=(3, 5))
    assert np.array_equal(y0, ref)
    x1 = x0.copy()
    x1[0, :, 3, 4] += 5
    y1 = p.forward(("changeIndexes", x1, np.array([3 * 10 + 4], dtype=np.int32)))
    assert np.array_equal(y1, x1.reshape(1, 3, 4, 2, 5, 2).max(axis=(3, 5)))
