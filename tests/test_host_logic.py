"""Host-side logic (producers, chunk bookkeeping, carries, device chaining,
resampler lengths) against the golden vectors and the oracle, with the
kernel-level calls replaced by numpy stand-ins (`fake_gpu`).  No GPU needed."""

import pickle
from functools import partial

import numpy as np
import pytest

from openseize_b200 import producer
from openseize_b200.core import numerical as nm
from openseize_b200.filtering.fir import Kaiser
from openseize_b200.filtering.iir import Butter, Notch
from openseize_b200.resampling.resampling import downsample
from openseize_b200.spectra.estimators import psd
from tests import parity_cases as pc


def test_fir(fake_gpu):
    pc.narrow_input_dtypes()
    pc.fir_golden()
    pc.fir_long_golden()
    pc.hilbert_golden()
    pc.fir_oracle_sweep()


def test_iir(fake_gpu):
    pc.iir_golden()
    pc.iir_oracle_sweep()


def test_resample(fake_gpu):
    pc.resample_golden()
    pc.resample_oracle_sweep()


def test_spectra(fake_gpu):
    pc.spectra_golden("pow2")
    pc.spectra_golden("nonpow2")
    pc.spectra_oracle_sweep()
    pc.spectra_oracle_sweep(fs=1000, resolutions=(0.5,))


def test_pipeline_chain(fake_gpu):
    pc.pipeline_chain()
    pc.fused_fir_decimate()
    pc.fused_iir_fir_decimate()


def test_analytic(fake_gpu):
    pc.analytic_golden()


def test_producer_tools(fake_gpu):
    pc.protools_golden()
    pc.masked_chain()
    pc.protools_edges()


def test_no_cpu_fallback():
    """Without a GPU (and without the test stand-ins) the operators raise."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    x = np.zeros((2, 5000))
    with pytest.raises(RuntimeError):
        Kaiser(500, 600, 5000)(x, 1000)
    with pytest.raises(RuntimeError):
        psd(x, 1024, resolution=1.0)


def test_laziness_and_errors(fake_gpu):
    x = np.random.default_rng(0).standard_normal((2, 5000))
    calls = []

    def source():
        calls.append(1)
        yield from (x[:, i:i + 1000] for i in range(0, 5000, 1000))

    pro = producer(source, 1000, -1, shape=x.shape)
    out = Kaiser(500, 600, 5000)(pro, 1000)
    assert not calls, "building an operator result must not pull data"
    assert out.shape == x.shape
    out.to_array()
    assert calls
    with pytest.raises(TypeError):
        producer(3.0, 10, -1)
    with pytest.raises(ValueError):
        producer(source, 10, -1)                      # generating function needs a shape
    with pytest.raises(ValueError):
        list(nm.polyphase_resample(producer(x, 1000, -1), 1, 6000, 5000, Kaiser, -1))
    with pytest.raises(ValueError):
        psd(producer(x, 100, -1), 5000, scaling="nope", resolution=5000 / 1024)
    with pytest.raises(ValueError):
        Kaiser([1, 2], [3], 100)
    # (b, a) above second order: the sequential DF2T path, not an error
    y = np.concatenate(list(nm.lfilter(producer(x, 1000, -1), (np.ones(5) / 5, [1.0, 0, 0, 0, 0.1]),
                                       -1)), -1)
    assert y.shape == x.shape


def test_short_source_still_flushes(fake_gpu):
    """A generating-function producer that yields fewer samples than its declared
    shape: the reference flushes the FIR tail / the last resampler chunk when the
    iterator ENDS (numerical.py:285-298, :618-632), not when the declared length
    is reached -- the samples that arrived are the recording."""
    import oracle

    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 9000))

    def source():
        yield from (x[:, i:i + 1000] for i in range(0, 9000, 1000))

    filt = Kaiser(500, 600, 5000)
    for mode in ("same", "full"):
        pro = producer(source, 1000, -1, shape=(2, 12000))     # declares 12000, yields 9000
        got = np.concatenate(list(nm.oaconvolve(pro, filt.coeffs, -1, mode)), -1)
        ref = np.stack([np.convolve(r, filt.coeffs, mode) for r in x])
        assert got.shape == ref.shape and np.max(np.abs(got - ref)) < 1e-12
    pro = producer(source, 1000, -1, shape=(2, 12000))
    got = np.concatenate(list(nm.polyphase_resample(pro, 1, 4, 5000, Kaiser, -1)), -1)
    ref = np.concatenate(oracle.polyphase_resample(x, 1, 4, 5000, 1000, -1), -1)
    assert got.shape == ref.shape and np.max(np.abs(got - ref)) < 1e-12


def test_time_ring_alignment_and_capacity(fake_gpu):
    """The staging ring keeps its write position and row pitch on multiples of 16 samples
    (the TMA-moved kernels upstream need 16-byte aligned rows) whatever is left live when it
    re-bases, keeps the live samples intact, and a ring told how much its consumer lets
    pile up is allocated once instead of being re-based every other block."""
    ring = nm._TimeRing(3, capacity=0)
    rng = np.random.default_rng(1)
    kept = np.zeros((3, 0))
    buffers = set()
    for step in range(12):
        n = 1600
        block = nm.dv.from_host(rng.standard_normal((3, n)))
        dst = ring.alloc(3, n)
        assert ring.pos % 16 == 0 and ring.buf.shape[1] % 16 == 0
        dst.copy_(block)
        ring.push(dst)
        kept = np.concatenate([kept, block.numpy()], -1)
        drop = 1600 - 37 if step % 2 else 1600 - 5          # odd leftovers
        ring.drop(drop)
        kept = kept[:, drop:]
        assert np.array_equal(ring.window().numpy(), kept)
        buffers.add(ring.buf.data_ptr())
    big = nm._TimeRing(3, capacity=20000)
    for step in range(10):
        dst = big.alloc(3, 1600)
        dst.zero_()
        big.push(dst)
    assert big.buf.shape[1] >= 20000 + 2 * 1600 and big.start == 0    # never re-based


def test_producer_mutation_contract(fake_gpu):
    """producer(Producer, cs, axis) mutates and returns the same object
    (reference core/producer.py:114-117); psd forces chunksize=int(fs)."""
    x = np.zeros((2, 40000))
    pro = producer(x, 1000, -1)
    assert producer(pro, 300, -1) is pro and pro.chunksize == 300
    psd(pro, 1024, resolution=1.0)
    assert pro.chunksize == 1024


def test_picklable_recipes():
    """Producers over partials of every GPU generating function pickle
    (reference tests/test_concurrency.py:85-149): no CUDA state is captured
    before iteration."""
    x = np.random.default_rng(0).standard_normal((2, 6000))
    pro = producer(x, 1000, -1)
    k, b, n = Kaiser(500, 600, 5000), Butter([1, 100], [0.5, 200], 5000), Notch(60, 6, 5000)
    recipes = [k(pro, 1000), b(pro, 1000), b(pro, 1000, dephase=False), n(pro, 1000),
               n(pro, 1000, dephase=False), downsample(pro, 4, 5000, 1000),
               nm.welch(pro, 1000, 256, "hann", 0.5, -1, "constant", "density")[1],
               nm.stft(pro, 1000, 256, "hann", 0.5, -1, "constant", "density", True, True)[2]]
    for r in recipes:
        clone = pickle.loads(pickle.dumps(r))
        assert clone.shape == r.shape and clone.chunksize == r.chunksize
    assert pickle.loads(pickle.dumps(partial(nm.oaconvolve, pro, k.coeffs, -1, "same")))
