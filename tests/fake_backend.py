"""numpy stand-ins for the kernel-level calls (TEST INFRASTRUCTURE ONLY).

Each stand-in implements the contract written in include/osz_b200.h for the
matching C-ABI entry point, on CPU torch tensors, so the Python host logic of
openseize_b200 can be exercised by `pytest -m "not gpu"`.  Installed by the
`fake_gpu` fixture through monkeypatch; nothing in the package imports this.
"""

import numpy as np
import scipy.signal as sps
import torch

from openseize_b200.core import device as dv


def _t(a):
    return torch.from_numpy(np.array(a, dtype=np.float64, order="C", copy=True))


def _ret(y, out):
    """Write into the caller's buffer when one is given (the C ABI contract)."""
    if out is None:
        return _t(y)
    out.copy_(_t(y))
    return out


class FakeFir:
    def __init__(self, taps, algo=0):
        self.taps = np.asarray(taps, dtype=np.float64)
        self.ntaps = len(self.taps)
        self.algo = 0

    def run(self, xbuf, n_out, out=None):
        x = xbuf.numpy()
        k = self.ntaps
        y = np.stack([np.convolve(r[:n_out + k - 1], self.taps, "valid") for r in x])
        return _ret(y, out)


class FakeSos:
    def __init__(self, sos):
        self.sos = np.atleast_2d(np.asarray(sos, dtype=np.float64))
        self.nsec = self.sos.shape[0]

    def run(self, x, state, reverse=False, want_output=True, out=None):
        xn = x.numpy()
        if reverse:
            xn = xn[:, ::-1]
        zi = np.transpose(state.numpy(), (1, 0, 2))
        y, zf = sps.sosfilt(self.sos, xn, axis=-1, zi=zi)
        state.copy_(_t(np.transpose(zf, (1, 0, 2))))
        if not want_output:
            return None
        return _ret(y[:, ::-1] if reverse else y, out)

    def state_from_sample(self, zi, x, sample):
        zi = np.asarray(zi, dtype=np.float64)
        return _t(zi[None, :, :] * x.numpy()[:, sample][:, None, None])

    def lookahead(self, zi, x, reverse=True):
        state = self.state_from_sample(zi, x, x.shape[1] - 1 if reverse else 0)
        self.run(x, state, reverse=reverse, want_output=False)
        return state


class FakeTf:
    def __init__(self, b, a):
        self.b, self.a = np.atleast_1d(b).astype(float), np.atleast_1d(a).astype(float)
        self.nstate = max(len(self.b), len(self.a)) - 1

    def run(self, x, state, reverse=False, want_output=True, out=None):
        xn = x.numpy()
        if reverse:
            xn = xn[:, ::-1]
        y, zf = sps.lfilter(self.b, self.a, xn, axis=-1, zi=state.numpy())
        state.copy_(_t(zf))
        if not want_output:
            return None
        return _ret(y[:, ::-1] if reverse else y, out)

    def state_from_sample(self, zi, x, sample):
        return _t(np.asarray(zi, dtype=float)[None, :] * x.numpy()[:, sample][:, None])


class FakeUpfirdn:
    def __init__(self, h, up, down):
        self.h = np.asarray(h, dtype=np.float64) * up
        self.ntaps, self.up, self.down = len(self.h), int(up), int(down)
        self.kernel = "mma" if self.up == 1 and self.down >= 2 else "general"

    def run(self, x, x_first, out_first, n_out, out=None):
        xn = x.numpy()
        half = (self.ntaps - 1) // 2
        # u[m] = sum_k h'[m - k*up] x[k], m relative to x_first*up
        u = sps.upfirdn(self.h, xn, up=self.up, down=1, axis=-1)
        j = np.arange(out_first, out_first + n_out)
        m = j * self.down + half - x_first * self.up
        ok = (m >= 0) & (m < u.shape[1])
        y = np.zeros((xn.shape[0], n_out))
        y[:, ok] = u[:, m[ok]]
        return _ret(y, out)


# ---- fused IIR pass + decimating FIR (contracts of include/osz_b200.h) ----------
def _sosdec_spans(sos_plan, ufd_plan, rows, n):
    """Enough spans to exercise the span boundaries: 3 for long chunks, 1 otherwise,
    0 (not fusable) below four filter lengths."""
    k = ufd_plan.ntaps
    if ufd_plan.up != 1 or ufd_plan.down < 2 or n < 4 * k:
        return 0
    return 3 if n >= 12 * k else 1


def _pieces(n, nspan, reverse):
    """Real-time [ua, ub) of every LOGICAL span (a backward pass walks time in reverse)."""
    span_len = -(-n // nspan)
    out = []
    for k in range(nspan):
        a, b = k * span_len, min((k + 1) * span_len, n)
        out.append((n - b, n - a) if reverse else (a, b))
    return out


def _sosdec_exec(sos_plan, ufd_plan, x, reverse, state, nspan, first, out, out_first):
    rows, n = x.shape
    p = sos_plan.run(x, state, reverse=reverse).numpy()          # the pass output, real-time order
    k, m = ufd_plan.ntaps, ufd_plan.down
    half = (k - 1) // 2
    g = ufd_plan.h[::-1]
    edges = np.zeros((rows, nspan, 2, k - 1))
    o = out.numpy() if out.numel() else np.zeros((rows, 0))
    res = o.copy()
    for sp, (ua, ub) in enumerate(_pieces(n, nspan, reverse)):
        edges[:, sp, 0] = p[:, ua:ua + k - 1]
        edges[:, sp, 1] = p[:, ub - (k - 1):ub]
        j0 = -(-(first + ua - half + k - 1) // m)                # window start >= piece start
        j1 = (first + ub - 1 - half) // m                        # window end < piece end
        for j in range(j0, j1 + 1):
            c = j - out_first
            if 0 <= c < res.shape[1]:
                lo = j * m + half - (k - 1) - first
                res[:, c] = p[:, lo:lo + k] @ g
    if out.numel():
        out.copy_(_t(res))
    return _t(edges)


def _sosdec_boundary(ufd_plan, edges, reverse, prev_tail, has_end, n, first, out, out_first,
                     j_min, j_max):
    e = edges.numpy()
    rows, nspan = e.shape[0], e.shape[1]
    k, m = ufd_plan.ntaps, ufd_plan.down
    half = (k - 1) // 2
    g = ufd_plan.h[::-1]
    res = out.numpy().copy()
    pieces = sorted(range(nspan), key=lambda sp: _pieces(n, nspan, reverse)[sp][0])
    starts = [_pieces(n, nspan, reverse)[sp][0] for sp in pieces]
    zeros = np.zeros((rows, k - 1))
    for b in range(nspan + 1):
        if b == nspan and not has_end:
            continue
        T = n if b == nspan else starts[b]
        tail = (prev_tail.numpy() if prev_tail is not None else zeros) if b == 0 \
            else e[:, pieces[b - 1], 1]
        head = zeros if b == nspan else e[:, pieces[b], 0]
        win = np.concatenate([tail, head], axis=1)               # real times T-(k-1) .. T+k-2
        tg = first + T
        j0 = max(-(-(tg - half) // m), j_min)
        j1 = min(-(-(tg - half + k - 1) // m) - 1, j_max)
        for j in range(j0, j1 + 1):
            c = j - out_first
            if 0 <= c < res.shape[1]:
                base = j * m + half - tg
                res[:, c] = win[:, base:base + k] @ g
    if out.numel():
        out.copy_(_t(res))


class FakeSpec:
    def __init__(self, nfft, stride, window, detrend, norm):
        self.nfft, self.stride, self.nfreq = int(nfft), int(stride), int(nfft) // 2 + 1
        self.window = np.asarray(window, dtype=np.float64)
        self.detrend, self.norm, self.path = detrend, float(norm), 1

    def nseg_available(self, width):
        return (width - self.nfft) // self.stride + 1 if width >= self.nfft else 0

    def _dft(self, x, s):
        seg = x[:, s * self.stride:s * self.stride + self.nfft]
        if self.detrend in ("constant", "linear"):
            seg = sps.detrend(seg, axis=-1, type=self.detrend)
        return np.fft.rfft(seg * self.window, axis=-1)

    def _pgram(self, X):
        p = (X.real ** 2 + X.imag ** 2) * self.norm
        if self.nfft % 2:
            p[:, 1:] *= 2
        else:
            p[:, 1:-1] *= 2
        return p

    def welch_accum(self, x, nseg, psd_sum):
        xn = x.numpy()
        acc = psd_sum.numpy()
        for s in range(nseg):
            acc += self._pgram(self._dft(xn, s))

    def segments(self, x, nseg, complex_):
        xn = x.numpy()
        outs = []
        for s in range(nseg):
            X = self._dft(xn, s)
            if complex_:
                X = X * np.sqrt(self.norm)
                outs.append(np.stack([X.real, X.imag], axis=-1))
            else:
                outs.append(self._pgram(X))
        return _t(np.stack(outs))


class _Now:
    def __init__(self, arr):
        self.arr = arr

    def get(self):
        return self.arr


def _upload(arr, layout, alloc=None):
    if hasattr(arr, "records") and hasattr(arr, "chan_off"):
        arr = arr.decode()                      # EDF RawChunk (decoded on the GPU in the product)
    a = np.asarray(arr, dtype=np.float64)
    n = a.shape[layout.axis]
    a = np.moveaxis(a.reshape(layout.outer, n, layout.inner), 1, 2)
    return _ret(a.reshape(layout.rows, n), alloc(layout.rows, n) if alloc else None)


def _download(dev, layout, complex_=False):
    a = dev.numpy()
    n = a.shape[1]
    if complex_:
        a = a[..., 0] + 1j * a[..., 1]
    a = np.moveaxis(a.reshape(layout.outer, layout.inner, n), 2, 1)
    return _Now(np.ascontiguousarray(a).reshape(layout.host_shape(n)))


def _spec_prepare(x, n, nfft, window, detrend):
    seg = x.numpy()[:, :n]
    if detrend in ("constant", "linear"):
        seg = sps.detrend(seg, axis=-1, type=detrend)
    out = np.zeros((seg.shape[0], nfft))
    out[:, :n] = seg * np.asarray(window)
    return _t(out)


def _download_block(dev):
    return _Now(np.array(dev.numpy(), copy=True))


def _take_cols(x, idx):
    idx = np.asarray(idx, dtype=np.int64)
    if idx.size and (idx[0] < 0 or idx[-1] >= x.shape[1]):
        raise IndexError("index out of bounds")
    return _t(x.numpy()[:, idx])


class FakeRowMoments:
    """Contract of osz_row_moments_f64 (include/osz_b200.h)."""

    def __init__(self, rows, ignore_nan=True):
        self.acc, self.ignore_nan = np.zeros((rows, 3)), ignore_nan
        self.avg = np.nanmean if ignore_nan else np.mean

    def add(self, x):
        import warnings

        a, n = x.numpy(), x.shape[1]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.acc[:, 0] += n * self.avg(a, axis=1)
            self.acc[:, 1] += n * self.avg(a ** 2, axis=1)
        self.acc[:, 2] += n

    def result(self):
        return self.acc.copy()


def _row_standardize(x, mean_dev, std_dev, out=None):
    y = (x.numpy() - mean_dev.numpy()[:, None]) / std_dev.numpy()[:, None]
    return _ret(y, out)


def _col_moments(x, ignore_nan=True, want="mean"):
    a = x.numpy()
    avg, dev = (np.nanmean, np.nanstd) if ignore_nan else (np.mean, np.std)
    if want == "mean":
        return _t(avg(a, axis=0, keepdims=True))
    if want == "std":
        return _t(dev(a, axis=0, keepdims=True))
    return _t((a - avg(a, axis=0, keepdims=True)) / dev(a, axis=0, keepdims=True))


def install(mp):
    mp.setattr(dv, "DEVICE", "cpu")
    mp.setattr(dv, "require_cuda", lambda: torch)
    mp.setattr(dv, "upload", _upload)
    mp.setattr(dv, "download", _download)
    mp.setattr(dv, "spec_prepare", _spec_prepare)
    mp.setattr(dv, "download_block", _download_block)
    mp.setattr(dv, "take_cols", _take_cols)
    mp.setattr(dv, "RowMoments", FakeRowMoments)
    mp.setattr(dv, "row_standardize", _row_standardize)
    mp.setattr(dv, "col_moments", _col_moments)
    mp.setattr(dv, "zip_complex", lambda re, im: torch.stack([re, im], dim=-1).contiguous())
    mp.setattr(dv, "sosdec_spans", _sosdec_spans)
    mp.setattr(dv, "sosdec_exec", _sosdec_exec)
    mp.setattr(dv, "sosdec_boundary", _sosdec_boundary)
    mp.setattr(dv.FirPlan, "cached", staticmethod(lambda taps, algo=0: FakeFir(taps, algo)))
    mp.setattr(dv.SosPlan, "cached", staticmethod(lambda sos: FakeSos(sos)))
    mp.setattr(dv.TfPlan, "cached", staticmethod(lambda b, a: FakeTf(b, a)))
    mp.setattr(dv.UpfirdnPlan, "cached",
               staticmethod(lambda h, up, down: FakeUpfirdn(h, up, down)))
    mp.setattr(dv.SpecPlan, "cached",
               staticmethod(lambda nfft, stride, window, detrend, norm:
                            FakeSpec(nfft, stride, window, detrend, norm)))


def install_raw():
    """Same as install() without pytest's monkeypatch (spawned gloo workers)."""
    class _Set:
        @staticmethod
        def setattr(obj, name, value):
            setattr(obj, name, value)

    install(_Set)
