"""Mimi split residual vector quantiser (first slice): codes <-> latent.
CPU: the C oracle (oracle/mimi_rvq_ref.c) against an independent numpy restatement of the reference graphs
(quantization/core_vq.h:28-55, 136-193; vq.h:18-117; torch.h:18-37), edge cases, properties.
GPU (-m gpu): msx_rvq_* through the C ABI against the oracle — codes bit-exact, latents equal up to double-summation order."""
import numpy as np
import pytest

import oracle


def make_quantiser(rng, n_sem, n_rest, bins, D, dim):
    cb_first = rng.standard_normal((n_sem, bins, D)).astype(np.float32)
    cb_rest = (rng.standard_normal((n_rest, bins, D)) * np.linspace(1.0, 0.2, max(n_rest, 1))[:n_rest, None, None]).astype(np.float32)
    f16 = lambda a: a.astype(np.float16).view(np.uint16)
    in_first = f16(rng.standard_normal((D, dim)) / np.sqrt(dim)); in_rest = f16(rng.standard_normal((D, dim)) / np.sqrt(dim))
    out_first = f16(rng.standard_normal((dim, D)) / np.sqrt(D)); out_rest = f16(rng.standard_normal((dim, D)) / np.sqrt(D))
    return cb_first, cb_rest, in_first, in_rest, out_first, out_rest


def np_conv(w_bits, x):
    """ggml_conv_1d, kernel 1: x rounded to F16 by im2col, F16 x F16 products exact, accumulated in double here"""
    w = w_bits.view(np.float16).astype(np.float64)
    xr = x.astype(np.float16).astype(np.float64)
    return (xr @ w.T).astype(np.float32)


def np_rvq_encode(cb, x):
    """core_vq.h:28-55 + 174-193: c = sum_rows((b - a)^2) in index order in double, r = 1 / (c + 1), first maximum; residual -= centroid"""
    res = x.copy()
    codes = np.zeros((cb.shape[0], x.shape[0]), dtype=np.int32)
    for q in range(cb.shape[0]):
        for t in range(x.shape[0]):
            diff = (cb[q] - res[t][None, :]).astype(np.float32)
            sq = (diff * diff).astype(np.float32)
            s = np.zeros(cb.shape[1], dtype=np.float64)
            for d in range(cb.shape[2]):
                s += sq[:, d].astype(np.float64)
            r = np.float32(1.0) / (s.astype(np.float32) + np.float32(1.0))
            j = int(np.argmax(r))
            codes[q, t] = j
            res[t] = (res[t] - cb[q, j]).astype(np.float32)
    return codes


def np_rvq_decode(cb, codes):
    out = None
    for q in range(codes.shape[0]):
        v = cb[q][codes[q]]
        out = v.copy() if out is None else (out + v).astype(np.float32)
    return out


def np_split_encode(qz, x, n_q):
    cb_first, cb_rest, in_first, in_rest, _, _ = qz
    n_sem = cb_first.shape[0]
    c1 = np_rvq_encode(cb_first[:min(n_q, n_sem)], np_conv(in_first, x))
    if n_q <= n_sem:
        return c1
    return np.concatenate([c1, np_rvq_encode(cb_rest[:n_q - n_sem], np_conv(in_rest, x))], axis=0)


def np_split_decode(qz, codes):
    cb_first, cb_rest, _, _, out_first, out_rest = qz
    n_sem = cb_first.shape[0]
    y = np_conv(out_first, np_rvq_decode(cb_first, codes[:n_sem]))
    if codes.shape[0] > n_sem:
        y = (y + np_conv(out_rest, np_rvq_decode(cb_rest, codes[n_sem:]))).astype(np.float32)
    return y


@pytest.mark.parametrize("n_sem,n_rest,bins,D,dim,T,n_q", [(1, 3, 64, 16, 24, 5, 4), (1, 7, 128, 32, 48, 3, 1), (2, 2, 32, 8, 8, 9, 3), (1, 0, 16, 4, 4, 2, 1)])
def test_oracle_matches_numpy_restatement(n_sem, n_rest, bins, D, dim, T, n_q):
    rng = np.random.default_rng(n_sem * 100 + bins + T)
    qz = make_quantiser(rng, n_sem, n_rest, bins, D, dim)
    o = oracle.SplitRVQ(*qz)
    x = rng.standard_normal((T, dim)).astype(np.float32) * 2.0
    codes = o.encode(x, n_q)
    assert codes.shape == (n_q, T) and codes.min() >= 0 and codes.max() < bins
    assert np.array_equal(codes, np_split_encode(qz, x, n_q))
    y = o.decode(codes)
    assert np.array_equal(y.view(np.uint32), np_split_decode(qz, codes).view(np.uint32))


def test_oracle_ties_take_the_first_centroid_and_codes_of_centroids_round_trip():
    rng = np.random.default_rng(8)
    n_sem, n_rest, bins, D, dim = 1, 2, 32, 8, 8
    qz = list(make_quantiser(rng, n_sem, n_rest, bins, D, dim))
    qz[0][0, 7] = qz[0][0, 3]                                   # duplicate centroid: argmax returns the first maximum (ggml_argmax)
    eye = np.eye(D, dim, dtype=np.float16).view(np.uint16)      # identity projections: the latent IS the code-space vector
    qz[2] = qz[3] = eye; qz[4] = qz[5] = np.eye(dim, D, dtype=np.float16).view(np.uint16)
    o = oracle.SplitRVQ(*qz)
    x = qz[0][0, [3, 7, 11]].astype(np.float16).astype(np.float32)       # exactly representable after the F16 rounding of im2col
    qz[0][0, [3, 7, 11]] = x
    o = oracle.SplitRVQ(*qz)
    codes = o.encode(x, 1)
    assert list(codes[0]) == [3, 3, 11]
    assert np.array_equal(o.decode(codes), x)                    # a centroid decodes to itself through identity projections
    # residual quantisation: every extra layer can only shrink (or keep) the residual it minimises greedily
    xr = rng.standard_normal((6, dim)).astype(np.float32)
    errs = []
    for n_q in (1, 2, 3):
        c = o.encode(xr, n_q)
        rec = np_rvq_decode(qz[0], c[:1]) + (np_rvq_decode(qz[1], c[1:]) if n_q > 1 else 0)
        errs.append(float(np.sum((xr.astype(np.float16).astype(np.float32) - rec) ** 2)))
    assert np.all(np.isfinite(errs))


def _golden():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_rvq.npz"))
    return [g[k] for k in ("cb_first", "cb_rest", "in_first", "in_rest", "out_first", "out_rest")], g["x"], g["codes"], g["y"]


def test_oracle_matches_committed_golden_vectors():
    """tests/golden/golden_rvq.npz (written by make_golden_rvq.py from the numpy restatement): the C oracle reproduces it bit for bit"""
    qz, x, codes, y = _golden()
    o = oracle.SplitRVQ(*qz)
    assert np.array_equal(o.encode(x, codes.shape[0]), codes)
    assert np.array_equal(o.encode(x, 1), codes[:1])
    assert np.array_equal(o.decode(codes).view(np.uint32), y.view(np.uint32))


@pytest.mark.gpu
def test_gpu_matches_committed_golden_vectors():
    from moshi_cpp_b200 import binding as msx
    qz, x, codes, y = _golden()
    g = msx.RVQ(*qz)
    assert np.array_equal(g.encode(x, codes.shape[0]), codes)
    got = g.decode(codes)
    assert np.max(np.abs(got - y)) <= 1e-6 * np.max(np.abs(y)) and np.mean(got.view(np.uint32) == y.view(np.uint32)) > 0.99
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("T,n_q", [(1, 8), (33, 8), (7, 32), (375, 16)])
def test_gpu_split_rvq_matches_oracle_at_mimi_sizes(T, n_q):
    """Mimi's quantiser shapes (1 semantic + 31 acoustic codebooks of 2048 x 256, latent 512): codes bit-exact, latents bit-exact up to the
    order of the double sums inside the 1 x 1 convolutions; T = 1 is the streaming step, 375 a 30 s clip"""
    from moshi_cpp_b200 import binding as msx
    rng = np.random.default_rng(T * 7 + n_q)
    qz = make_quantiser(rng, 1, 31, 2048, 256, 512)
    o = oracle.SplitRVQ(*qz); g = msx.RVQ(*qz)
    x = rng.standard_normal((T, 512)).astype(np.float32) * 3.0
    c_ref = o.encode(x, n_q); c_gpu = g.encode(x, n_q)
    assert np.array_equal(c_gpu, c_ref), f"{int(np.sum(c_gpu != c_ref))} codes differ"
    y_ref = o.decode(c_ref); y_gpu = g.decode(c_ref)
    assert np.max(np.abs(y_gpu - y_ref)) <= 1e-6 * np.max(np.abs(y_ref))
    assert np.mean(y_gpu.view(np.uint32) == y_ref.view(np.uint32)) > 0.999
    # second call on the same handle (work buffers and keys reused), fewer codebooks
    assert np.array_equal(g.encode(x, 1), c_ref[:1])
    g.close()


@pytest.mark.gpu
def test_gpu_rvq_rejects_bad_arguments():
    from moshi_cpp_b200 import binding as msx
    rng = np.random.default_rng(1)
    qz = make_quantiser(rng, 1, 3, 64, 16, 24)
    g = msx.RVQ(*qz)
    with pytest.raises(Exception):
        g.encode(np.zeros((2, 24), np.float32), 5)               # more codebooks than the quantiser has
    with pytest.raises(Exception):
        g.decode(np.full((2, 3), 64, np.int32))                  # code out of range
    assert g.decode(np.zeros((4, 2), np.int32)).shape == (2, 24)
    g.close()
