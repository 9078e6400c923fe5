"""CPU oracle for the openseize chunked filtering + spectral hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``openseize_b200/`` imports this
package.  The only permitted importers are ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and there only as the checker / reported CPU
baseline -- never as the thing shipped.

What it is
----------
A numpy/scipy restatement of the algorithms in the reference's
``src/openseize/core/numerical.py`` (the hot path),
written array-in / list-of-arrays-out instead of as producer generators.
The arithmetic primitives the reference delegates to live in third-party
packages that are *not* under /root/reference and are unpinned there
(``pyproject.toml:29-38`` lists bare ``numpy`` and ``scipy``); this image has
numpy 2.3.5 and scipy 1.18.1, and those are the versions the oracle is pinned
against:

* ``np.fft.rfft / irfft``      (pocketfft)          numerical.py:214,235,241,699
* ``scipy.signal.sosfilt``     (DF2T biquads)       numerical.py:334,399,402,410
* ``scipy.signal.lfilter``     (DF2T)               numerical.py:445,508,511,519
* ``scipy.signal.resample_poly`` -> ``upfirdn``     numerical.py:610,631
* ``scipy.signal.detrend / get_window``             numerical.py:691,694
* ``scipy.signal.sosfilt_zi / lfilter_zi``          numerical.py:378,487

``oracle/prim.c`` restates the published algorithms of those primitives in
plain C (direct convolution, DF2T recurrences, upfirdn, naive DFT) so the
scipy calls themselves are pinned independently.

Parity status: PINNED.  ``oracle/make_golden.py`` (run in the build container,
where /root/reference exists) imports the real reference, checks every oracle
function against it on seeded inputs, and writes the reference's outputs to
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` re-checks the oracle
against those vectors everywhere (no /root/reference needed).
"""

from oracle.chunked import (  # noqa: F401
    split_chunks,
    oaconvolve,
    sosfilt,
    sosfiltfilt,
    lfilter,
    filtfilt,
    polyphase_resample,
    resample_filter,
    modified_dft,
    periodogram,
    segments,
    welch_psd,
    stft,
    masked,
    pro_mean,
    pro_std,
    standardize,
)
