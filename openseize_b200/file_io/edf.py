"""Reader of European Data Format (EDF / EDF+) recordings with the interface of
the reference's ``openseize.file_io.edf.Reader`` (reference
file_io/edf.py:296-586, header layout :139-160): ``shape``, ``channels``,
``read(start, stop, padvalue)`` returning float64 ``(channels, samples)``,
usable with or without a context manager, picklable when closed.

What is new is ``read_raw``: the int16 samples of the selected channels exactly
as they sit in the data records, plus the per-channel (slope, offset) of the
EDF calibration ``p = slope * d + offset``.  A ``ReaderProducer`` over this
reader hands whole data records to the GPU path (``core.device.upload``), which
copies the int16 block over PCIe untouched and de-interleaves and calibrates it in
a kernel (``osz_decode_edf_records_f64``) with the same two roundings as the reference's
``arr * slopes; result += offsets`` (file_io/edf.py:412-419) -- so the device
rows are bit-identical to ``Reader.read`` at a quarter of the PCIe traffic.

Not here: annotation signals (the reference does not read them either),
``Writer`` and ``splitter`` (host-side file utilities, out of the hot path).
"""

from pathlib import Path

import numpy as np

# EDF header fields in file order: name -> (bytes per item, type, per-signal?)
_FIELDS = (
    ("version", 8, str, False), ("patient", 80, str, False), ("recording", 80, str, False),
    ("start_date", 8, str, False), ("start_time", 8, str, False), ("header_bytes", 8, int, False),
    ("reserved_0", 44, str, False), ("num_records", 8, int, False),
    ("record_duration", 8, float, False), ("num_signals", 4, int, False),
    ("names", 16, str, True), ("transducers", 80, str, True), ("physical_dim", 8, str, True),
    ("physical_min", 8, float, True), ("physical_max", 8, float, True),
    ("digital_min", 8, float, True), ("digital_max", 8, float, True),
    ("prefiltering", 80, str, True), ("samples_per_record", 8, int, True),
    ("reserved_1", 32, str, True),
)


class Header(dict):
    """The EDF header as a dict with attribute access (reference
    file_io/edf.py:107-290, file_io/bases.py:26-120)."""

    def __init__(self, path):
        super().__init__()
        self.path = Path(path) if path else None
        if self.path:
            self.update(self._read())

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError("'Header' object has no attribute '{}'".format(name)) from exc

    def _read(self):
        out = {}
        with open(self.path, "rb") as fp:
            fp.seek(252)
            nsig = int(fp.read(4).strip().decode("ascii"))
            fp.seek(0)
            for name, nbytes, typ, per_signal in _FIELDS:
                count = nsig if per_signal else 1
                items = [typ(fp.read(nbytes).strip().decode("ascii")) for _ in range(count)]
                out[name] = items if per_signal else items[0]
        return out

    @property
    def annotated(self):
        return "EDF Annotations" in self.names

    @property
    def annotation(self):
        return self.names.index("EDF Annotations") if self.annotated else None

    @property
    def channels(self):
        """Indices of the ordinary (non-annotation) signals.  (Same truthiness test
        as the reference, file_io/edf.py:231-233: an annotation signal at index 0 is
        not removed.)"""
        signals = list(range(self.num_signals))
        if self.annotation:
            signals.pop(self.annotation)
        return signals

    @property
    def samples(self):
        counts = np.array(self.samples_per_record) * self.num_records
        return [counts[ch] for ch in self.channels]

    @property
    def record_map(self):
        cum = np.cumsum(np.insert(self.samples_per_record, 0, 0))
        return [slice(a, b) for a, b in zip(cum, cum[1:])]

    @property
    def slopes(self):
        pmax = np.array(self.physical_max)[self.channels]
        pmin = np.array(self.physical_min)[self.channels]
        dmax = np.array(self.digital_max)[self.channels]
        dmin = np.array(self.digital_min)[self.channels]
        return (pmax - pmin) / (dmax - dmin)

    @property
    def offsets(self):
        pmin = np.array(self.physical_min)[self.channels]
        dmin = np.array(self.digital_min)[self.channels]
        return pmin - self.slopes * dmin


class RawChunk:
    """Whole EDF data records as they sit in the file plus what is needed to turn
    them into the reference's float64 ``(channels, n)`` chunk:

    records  int16 (nrec, per_record): record r = [ch0: spr | ch1: spr | ...]
    chan_off start of every selected channel inside a record
    spr      samples per record of the selected channels (all equal)
    skip     samples of the first record that precede the chunk
    shape    (channels, n)
    value    records[(skip+i) // spr, chan_off[c] + (skip+i) % spr] * slopes[c] + offsets[c]
    """

    def __init__(self, records, chan_off, spr, skip, n, slopes, offsets):
        self.records = np.ascontiguousarray(records, dtype=np.int16)
        self.chan_off = np.ascontiguousarray(chan_off, dtype=np.int32)
        self.spr, self.skip = int(spr), int(skip)
        self.slopes = np.ascontiguousarray(slopes, dtype=np.float64)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.float64)
        self.shape = (len(self.chan_off), int(n))
        self.ndim = 2

    def raw(self):
        """int16 ``(channels, n)``: the de-interleaved samples."""
        n = self.shape[1]
        rows = [self.records[:, o:o + self.spr].reshape(-1)[self.skip:self.skip + n]
                for o in self.chan_off]
        return np.stack(rows, axis=0) if rows else np.empty((0, n), dtype=np.int16)

    def decode(self):
        """The host value, with the reference's rounding (multiply, then add)."""
        result = self.raw() * self.slopes[:, None]
        result += self.offsets[:, None]
        return result


class Reader:
    """EDF reader (reference file_io/edf.py:296-586)."""

    def __init__(self, path):
        self.path = Path(path)
        self.mode = "rb"
        self._fobj = open(self.path, self.mode)
        self.header = Header(path)
        self._channels = self.header.channels
        self._meta = None          # per-channel-selection constants of read_raw

    # ---- file handling (reference file_io/bases.py Reader: open / close / context)
    def open(self):
        if self._fobj is None or self._fobj.closed:
            self._fobj = open(self.path, self.mode)

    def close(self):
        if self._fobj is not None and not self._fobj.closed:
            self._fobj.close()

    def __enter__(self):
        self.open()
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.close()

    def __getstate__(self):
        state = dict(self.__dict__)
        self.close()
        state["_fobj"] = None
        return state

    # ---- the reference interface
    @property
    def channels(self):
        return self._channels

    @channels.setter
    def channels(self, values):
        if not isinstance(values, (list, tuple, range)):
            raise ValueError("Channels must be type Sequence not {}".format(type(values)))
        self._channels = list(values)
        self._meta = None

    @property
    def shape(self):
        return len(self.channels), max(self.header.samples)

    def _find_records(self, start, stop, channels):
        spr = np.array(self.header.samples_per_record)[channels]
        return list(zip(start // spr, np.ceil(stop / spr).astype("int")))

    def _records(self, a, b):
        """Records a .. b-1 as one (records, samples per record) int16 array."""
        if a >= self.header.num_records:
            return np.empty((1, 0), dtype="<i2")
        b = min(b, self.header.num_records)
        per_record = sum(self.header.samples_per_record)
        self.open()
        self._fobj.seek(0)
        offset = self.header.header_bytes + a * per_record * 2
        recs = np.fromfile(self._fobj, "<i2", (b - a) * per_record, offset=offset)
        return recs.reshape(b - a, per_record)

    def _raw_rows(self, start, stop, channels):
        """One 1-D int16 array per channel, samples start .. stop-1 (shorter at
        the end of the file)."""
        rec_tuples = self._find_records(start, stop, channels)
        reads = {tup: self._records(*tup) for tup in set(rec_tuples)}
        rows = []
        for ch, tup in zip(channels, rec_tuples):
            arr = reads[tup][:, self.header.record_map[ch]].flatten()
            a = start - tup[0] * self.header.samples_per_record[ch]
            rows.append(arr[a:a + (stop - start)])
        return rows

    def _decipher(self, arr, channels, axis=-1):
        slopes = np.expand_dims(np.array(self.header.slopes[channels]), axis=axis)
        offsets = np.expand_dims(np.array(self.header.offsets[channels]), axis=axis)
        result = arr * slopes
        result += offsets
        return result

    def read(self, start, stop=None, padvalue=np.nan):
        """float64 ``(channels, stop - start)`` (reference file_io/edf.py:558-586);
        channels with a lower sample rate are padded with ``padvalue``."""
        if start > max(self.header.samples):
            return np.empty((len(self.channels), 0))
        if not stop:
            stop = max(self.header.samples)
        start, stop = int(start), int(stop)
        rows = self._raw_rows(start, stop, self.channels)
        longest = max(len(r) for r in rows)
        if all(len(r) == longest for r in rows):
            stacked = np.stack(rows, axis=0)
        else:
            stacked = np.stack([np.pad(r.astype(float), (0, longest - len(r)),
                                       constant_values=padvalue) for r in rows], axis=0)
        return self._decipher(stacked, self.channels)

    # ---- raw ingest for the GPU path
    @property
    def uniform_rate(self):
        """True when every selected channel has the same samples per record (the
        raw path stacks int16 rows; mixed rates need the float padding of read)."""
        spr = np.array(self.header.samples_per_record)[self.channels]
        return bool(np.all(spr == spr[0]))

    def _raw_meta(self):
        """Everything read_raw needs that depends only on the header and the
        channel selection (psd pulls fs-sized chunks: thousands of calls)."""
        if self._meta is None:
            h, chs = self.header, list(self.channels)
            self._meta = dict(
                spr=int(h.samples_per_record[chs[0]]),
                chan_off=np.array([h.record_map[ch].start for ch in chs], dtype=np.int32),
                per_record=int(sum(h.samples_per_record)),
                slopes=np.ascontiguousarray(h.slopes[chs], dtype=np.float64),
                offsets=np.ascontiguousarray(h.offsets[chs], dtype=np.float64),
                total=int(max(h.samples)), nrec=int(h.num_records), head=int(h.header_bytes),
                uniform=self.uniform_rate)
        return self._meta

    def read_raw(self, start, stop=None):
        """``RawChunk`` for samples start .. stop-1 of the selected channels: the
        records that hold them, untouched (one contiguous file read)."""
        m = self._raw_meta()
        if not m["uniform"]:
            raise ValueError("read_raw needs channels of equal sample rate")
        if not stop:
            stop = m["total"]
        start, stop = int(start), min(int(stop), m["total"])
        spr, per_record = m["spr"], m["per_record"]
        if start >= stop:
            recs, skip, n = np.empty((0, per_record), dtype=np.int16), 0, 0
        else:
            first, last = start // spr, min(-(-stop // spr), m["nrec"])
            self.open()
            self._fobj.seek(m["head"] + first * per_record * 2)
            recs = np.fromfile(self._fobj, "<i2", (last - first) * per_record)
            recs = recs.reshape(last - first, per_record)
            skip, n = start - first * spr, stop - start
        return RawChunk(recs, m["chan_off"], spr, skip, n, m["slopes"], m["offsets"])


def write_edf(path, data, fs, record_samples=None, physical=None, names=None):
    """Minimal EDF writer for tests and demos: ``data`` float64
    ``(channels, samples)`` with ``samples`` a multiple of ``record_samples``.
    Values are quantised like the reference Writer (file_io/edf.py:685-697):
    ``rint((x - offset) / slope)`` as little-endian int16."""
    data = np.atleast_2d(np.asarray(data, dtype=np.float64))
    nch, n = data.shape
    record_samples = int(record_samples or fs)
    if n % record_samples:
        raise ValueError("samples must be a multiple of record_samples")
    nrec = n // record_samples
    if physical is None:
        physical = [(float(np.floor(row.min())), float(np.ceil(row.max()))) for row in data]
    dmin, dmax = -32768.0, 32767.0
    head = {
        "version": "0", "patient": "synthetic", "recording": "openseize_b200",
        "start_date": "01.01.26", "start_time": "00.00.00", "header_bytes": 256 * (nch + 1),
        "reserved_0": "", "num_records": nrec, "record_duration": record_samples / fs,
        "num_signals": nch,
        "names": list(names or ["ch%d" % i for i in range(nch)]),
        "transducers": ["" for _ in range(nch)], "physical_dim": ["uV"] * nch,
        "physical_min": [p[0] for p in physical], "physical_max": [p[1] for p in physical],
        "digital_min": [dmin] * nch, "digital_max": [dmax] * nch,
        "prefiltering": [""] * nch, "samples_per_record": [record_samples] * nch,
        "reserved_1": [""] * nch,
    }
    with open(path, "wb") as fp:
        for name, nbytes, typ, per_signal in _FIELDS:
            items = head[name] if per_signal else [head[name]]
            for item in items:
                if typ is float:
                    text = repr(float(item))
                    if len(text) > nbytes:
                        text = ("%.*g" % (nbytes - 2, float(item)))[:nbytes]
                else:
                    text = str(item)
                fp.write(text.encode("ascii").ljust(nbytes)[:nbytes])
        slopes = np.array([(p[1] - p[0]) / (dmax - dmin) for p in physical])
        offsets = np.array([p[0] for p in physical]) - slopes * dmin
        q = np.rint((data - offsets[:, None]) / slopes[:, None]).astype("<i2")
        recs = q.reshape(nch, nrec, record_samples).transpose(1, 0, 2)
        fp.write(np.ascontiguousarray(recs).tobytes())
    return path
