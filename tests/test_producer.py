"""Producer / FIFO semantics (reference tests/test_producer.py, core/queues.py)."""

import numpy as np

from openseize_b200 import producer
from openseize_b200.core.producer import ArrayProducer, GenProducer, MaskedProducer
from openseize_b200.core.queues import FIFOArray


def test_array_producer_chunks_are_views():
    x = np.arange(2 * 10007, dtype=float).reshape(2, 10007)
    pro = producer(x, 1000, -1)
    assert isinstance(pro, ArrayProducer) and pro.shape == x.shape
    chunks = list(pro)
    assert [c.shape[-1] for c in chunks] == [1000] * 10 + [7]
    assert all(np.shares_memory(c, x) for c in chunks)
    assert np.array_equal(np.concatenate(chunks, -1), x)
    assert np.array_equal(pro.to_array(), x)
    xt = x.T.copy()
    assert np.array_equal(producer(xt, 999, 0).to_array(), xt)


def test_sequence_and_generator_producers():
    rng = np.random.default_rng(0)
    parts = [rng.standard_normal((3, n)) for n in (100, 17, 2000, 1, 555)]
    full = np.concatenate(parts, -1)
    assert np.array_equal(producer(parts, 300, -1).to_array(), full)

    def gen(scale=1.0):
        for p in parts:
            yield scale * p

    pro = producer(gen, 300, -1, shape=full.shape)
    assert isinstance(pro, GenProducer)
    chunks = list(pro)
    assert [c.shape[-1] for c in chunks] == [300] * 8 + [full.shape[-1] - 2400]
    assert np.array_equal(np.concatenate(chunks, -1), full)
    assert np.array_equal(np.concatenate(list(pro), -1), full), "re-iterable"
    assert np.array_equal(producer(gen, 300, -1, shape=full.shape, scale=2.0).to_array(), 2 * full)
    it1, it2 = iter(pro), iter(pro)
    assert np.array_equal(next(it1), next(it2)), "independent interleaved iterators"


def test_masked_producer():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((4, 5000))
    mask = rng.random(5000) > 0.4
    pro = producer(x, 700, -1, mask=mask)
    assert isinstance(pro, MaskedProducer)
    assert pro.shape == (4, int(mask.sum()))
    assert np.array_equal(pro.to_array(), x[:, mask])
    assert all(c.shape[-1] == 700 for c in list(pro)[:-1])


def test_reader_producer():
    class Reader:
        def __init__(self, data):
            self.data, self.shape, self.is_open = data, data.shape, True

        def open(self):
            self.is_open = True

        def close(self):
            self.is_open = False

        def read(self, start, stop):
            assert self.is_open
            return self.data[:, start:stop]

    x = np.arange(3 * 1000, dtype=float).reshape(3, 1000)
    pro = producer(Reader(x), 128, -1, start=100, stop=900)
    assert pro.shape == (3, 800) and not pro.data.is_open
    assert np.array_equal(pro.to_array(), x[:, 100:900])


def test_fifo_array():
    fifo = FIFOArray(5, axis=-1)
    assert fifo.empty() and fifo.qsize() == 0 and not fifo.full()
    fifo.put(np.arange(12.0).reshape(2, 6))
    fifo.put(np.arange(12.0, 20.0).reshape(2, 4))
    assert fifo.qsize() == 10 and fifo.full()
    assert fifo.queue.shape == (2, 10)
    a = fifo.get()
    assert a.shape == (2, 5) and fifo.qsize() == 5
    b = fifo.get()
    assert b.shape == (2, 5) and fifo.empty()
    assert fifo.get().size == 0
