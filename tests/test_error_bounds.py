"""The candidate passes are approximate by design; what makes the results exact is that every
approximation carries a PROVEN bound that the completeness proof (select.cu) consumes.  These
tests restate the three approximations in numpy and check the bounds the CUDA code uses
(csrc/scan.cu prep_queries_kernel, csrc/gemm_tcgen05.cu
prep_queries_gemm_kernel) against the true fp64 scores, on random and on adversarial inputs."""
import numpy as np
import pytest

U = 2.0 ** -24


def _queries(rng, d):
    qs = [rng.standard_normal(d), rng.standard_normal(d) * 1e-6, rng.standard_normal(d) * 1e6,
          np.abs(rng.standard_normal(d)), np.full(d, 0.3), -np.full(d, 7.0)]
    spike = rng.standard_normal(d) * 1e-3
    spike[d // 2] = 50.0                      # one dominant element: the grid is coarse for all others
    qs.append(spike)
    tiny = np.zeros(d)
    tiny[0] = 1.0                             # a basis vector
    qs.append(tiny)
    edge = rng.standard_normal(d)
    edge[3] = np.max(np.abs(edge)) * (1 - 2.0 ** -30)   # rounds up to the clamp
    qs.append(edge)
    return qs


@pytest.mark.parametrize("planes", [2, 3])
@pytest.mark.parametrize("d,levels", [(96, 255), (1536, 15), (17, 255), (4096, 15)])
def test_quantized_scan_query_grid_bound(planes, d, levels):
    """Score of the fixed-point query against Min + c*Scale vs the exact cosine distance:
    |error| <= 2^-(8*planes-2) * sqrt(d) * max|q| / ||q||  (the per-query eps of the scan plan)."""
    rng = np.random.default_rng(d * planes)
    bits = 8 * planes
    n = 400
    codes = rng.integers(0, levels + 1, size=(n, d)).astype(np.float64)
    codes[0] = levels                                       # extreme rows
    codes[1] = 0
    codes[1, 0] = levels
    mins = rng.standard_normal(n) * 0.5
    scales = np.abs(rng.standard_normal(n)) * 0.01 + 1e-4
    y = mins[:, None] + codes * scales[:, None]             # decompress_*_quantization
    ynorm = np.sqrt((y * y).sum(1))
    for q in _queries(rng, d):
        qn = np.sqrt((q * q).sum())
        mx = np.abs(q).max()
        m, x = np.frexp(mx)
        e = bits - 1 - int(x)
        Q = np.clip(np.rint(q * 2.0 ** e), -(2 ** (bits - 1)), 2 ** (bits - 1) - 1)
        S = codes @ Q                                       # exact integer (fits fp64: < 2^53)
        assert np.all(S == np.rint(S))
        dot_hat = scales * S * 2.0 ** -e + mins * Q.sum() * 2.0 ** -e
        approx = 1.0 - dot_hat / (qn * ynorm)
        exact = 1.0 - (y @ q) / (qn * ynorm)
        eps_q = 1.01 * 2.0 ** -(bits - 2) * np.sqrt(d) * mx / qn
        assert np.max(np.abs(approx - exact)) <= eps_q, (np.max(np.abs(approx - exact)), eps_q)
        # digit planes reassemble Q exactly (signed high digit, unsigned low ones)
        Qi = Q.astype(np.int64)
        digs = [(Qi >> (8 * (planes - 1 - p))) & 0xFF for p in range(planes)]
        hi = np.where(digs[0] >= 128, digs[0] - 256, digs[0])
        rebuilt = hi * (1 << (8 * (planes - 1)))
        for p in range(1, planes):
            rebuilt = rebuilt + digs[p] * (1 << (8 * (planes - 1 - p)))
        assert np.array_equal(rebuilt, Qi)


@pytest.mark.parametrize("d", [64, 128, 768, 1536])
def test_cosine_gemm_fp16_bound(d):
    """Unit vectors rounded to fp16, products exact in fp32, fp32 accumulation:
    |acc - cos| <= 2^-10 * 1.01 + sqrt(d) * 2^-24 + d * 2^-22."""
    rng = np.random.default_rng(d)
    n = 300
    v = rng.standard_normal((n, d)).astype(np.float32).astype(np.float64)
    v[0] = np.abs(v[0])
    v[1, :] = 0.0
    v[1, 0] = 3.0                                           # all weight on one element
    v[2] = 1.0
    qs = _queries(rng, d)
    vn = v / np.sqrt((v * v).sum(1))[:, None]
    vh = vn.astype(np.float16).astype(np.float32)
    eps = 2.0 ** -10 * 1.01 + np.sqrt(d) * U + d * 2.0 ** -22
    for q in qs:
        qn_ = q / np.sqrt((q * q).sum())
        qh = qn_.astype(np.float16).astype(np.float32)
        acc = np.zeros(n, dtype=np.float32)
        for k0 in range(0, d, 16):                          # K = 16 per MMA step, fp32 accumulate
            acc = (acc + (vh[:, k0:k0 + 16] * qh[None, k0:k0 + 16]).sum(1, dtype=np.float32)).astype(np.float32)
        err = np.max(np.abs(acc.astype(np.float64) - vn @ qn_))
        assert err <= eps, (err, eps)


def test_scan_fp32_bound():
    """fp32 rows, fp32 accumulation in 4 partial sums + shuffle tree: (d/64 + 24) * 2^-24 on the
    cosine distance -- checked against a plain fp32 emulation."""
    rng = np.random.default_rng(5)
    for d in (128, 768, 1536):
        n = 500
        v = rng.standard_normal((n, d)).astype(np.float32)
        q = rng.standard_normal(d)
        q32 = q.astype(np.float32)
        inv = (1.0 / np.sqrt((v.astype(np.float64) ** 2).sum(1))).astype(np.float32)
        qinv = np.float32(1.0 / np.sqrt((q * q).sum()))
        acc = np.zeros(n, dtype=np.float32)
        for k in range(d):                                  # worst case: one serial fp32 chain
            acc = (acc + v[:, k] * q32[k]).astype(np.float32)
        score = (np.float32(1.0) - acc * inv * qinv).astype(np.float64)
        exact = 1.0 - (v.astype(np.float64) @ q32.astype(np.float64)) / (
            np.sqrt((v.astype(np.float64) ** 2).sum(1)) * np.sqrt((q32.astype(np.float64) ** 2).sum()))
        # the serial chain is looser than the kernel's tree; the bound must hold for the tree order
        tree = np.zeros(n, dtype=np.float32)
        parts = [np.zeros(n, dtype=np.float32) for _ in range(64)]
        for k in range(d):
            parts[k % 64] = (parts[k % 64] + v[:, k] * q32[k]).astype(np.float32)
        while len(parts) > 1:
            parts = [(parts[i] + parts[i + 1]).astype(np.float32) for i in range(0, len(parts), 2)]
        tree = parts[0]
        score_t = (np.float32(1.0) - tree * inv * qinv).astype(np.float64)
        eps = (d / 64.0 + 24.0) * U
        assert np.max(np.abs(score_t - exact)) <= eps, (d, np.max(np.abs(score_t - exact)), eps)
        assert np.max(np.abs(score - exact)) <= 40 * eps     # sanity of the emulation itself


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def _fma32(a, b, c):
    """fp32 fma: the exact product and sum in fp64 (both fit: 24 + 24 + alignment < 2^-53 relative here), rounded once."""
    return _f32(np.float64(a) * np.float64(b) + np.float64(c))


@pytest.mark.parametrize("d", [16, 96, 128])
def test_i8_plan_digit_sums_keys_and_coarse_filter(d):
    """csrc/gemm_i8.cu restated in numpy (d <= 128, the mantissa-trick variant):
    (1) the two MMA accumulators Sa = sum a_k c_k (signed high digit) and Sb = sum b_k c_k (unsigned low digit)
        reassemble the exact integer sum(Q_k c_k) the dp4a scan forms;
    (2) the fp32 key 1 - x*fx/||q||, x = cx*S + cy*sum(Q), is within 48*2^-24 of the exact score of the grid
        query (the arithmetic part of the bound the window proof consumes; the grid part is tested above);
    (3) the coarse filter's bound, evaluated in fp32 exactly as the kernel does -- F = float_bits(Sa + 0x4B400000),
        xu = fma(256cx, F, fma(cy, Cq, K')) with K' formed in fp64 and rounded up -- is never below the x the exact
        path computes: a row the exact path would admit always passes the coarse test."""
    rng = np.random.default_rng(d)
    n = 3000
    codes = rng.integers(0, 256, size=(n, d)).astype(np.int64)
    codes[0] = 255
    codes[1] = 0
    codes[2, ::2] = 255
    mins = rng.standard_normal(n) * np.exp(rng.uniform(-3, 3, n))
    scales = np.abs(rng.standard_normal(n)) * np.exp(rng.uniform(-6, 1, n)) + 1e-9
    mins[3], scales[3] = 100.0, 1e-6 / 255                      # nearly constant row: |min| sqrt(d) / ||y|| ~ 1
    y = mins[:, None] + codes * scales[:, None]
    ynorm = np.sqrt((y * y).sum(1))
    cx, cy = _f32(scales / ynorm), _f32(mins / ynorm)           # ingest: {scale, min}/||y|| as fp32
    sb_max = np.float32(65025.0 * d + 2048.0)
    worst = 0.0
    for q in _queries(rng, d):
        qn = np.sqrt((q * q).sum())
        mx = np.abs(q).max()
        _, xe = np.frexp(mx)
        e = 15 - int(xe)
        Q = np.clip(np.rint(q * 2.0 ** e), -32768, 32767).astype(np.int64)
        lo = Q & 0xFF
        hi = (Q - lo) >> 8                                      # signed high digit
        assert np.all((hi >= -128) & (hi <= 127))
        Sa, Sb = codes @ hi, codes @ lo
        assert np.array_equal(256 * Sa + Sb, codes @ Q)         # (1)
        assert np.all(np.abs(Sa) < 2 ** 22) and np.all((Sb >= 0) & (Sb < 2 ** 23))
        fx = np.float32(2.0 ** -e)
        inv_norm = np.float32(1.0 / qn)
        Cq = np.float32(np.float32(Q.sum() * 2.0 ** -e) / fx)   # QStat.sum / QStat.fx
        c1 = -(inv_norm * fx)
        # exact path of the kernel
        fa = _f32(_f32(Sa + 12582912) - np.float32(12582912.0))
        fb = _f32(_f32(Sb + 8388608) - np.float32(8388608.0))
        assert np.array_equal(fa, Sa) and np.array_equal(fb, Sb)
        S = _fma32(fa, np.float32(256.0), fb)
        x = _fma32(cx, S, _f32(cy * Cq))
        key = _fma32(x, c1, np.float32(1.0))
        # the same score in fp64 from the same fp32 inputs
        ref = 1.0 - (np.float64(cx) * (256.0 * Sa + Sb) + np.float64(cy) * np.float64(Cq)) * np.float64(inv_norm) * np.float64(fx)
        err = np.max(np.abs(np.float64(key) - ref))
        worst = max(worst, err)
        assert err <= 48 * U, (err / U, d)                     # (2)
        # coarse filter
        kp = np.float64(cx) * np.float64(sb_max) + 1.5 * np.abs(np.float64(cy)) - np.float64(cx) * 256.0 * 12582912.0
        kpf = _f32(kp)
        up = np.float64(kpf) < kp
        kpf = np.where(up, np.nextafter(kpf, np.float32(np.inf)), kpf).astype(np.float32)
        assert np.all(np.float64(kpf) >= kp)
        F = _f32(Sa + 12582912)                                 # float_bits(Sa + 0x4B400000): exact below 2^24
        xu = _fma32(_f32(cx * np.float32(256.0)), F, _fma32(cy, Cq, kpf))
        assert np.all(xu >= x), (np.min(np.float64(xu) - np.float64(x)), d)   # (3)
    assert worst > 0.0
