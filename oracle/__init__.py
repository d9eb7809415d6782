"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

CPU restatement of the mel-spectrogram / mu-law hot path of
keunwoochoi/torchaudio-contrib (reference files cited per function).

Two restatements live here:

* ``oracle.ref_chain`` -- the reference's *own torch fp32 CPU path*, restated op for op
  (same torch CPU operators, same evaluation order).  This is the parity oracle: the
  north star asks for agreement with "the reference's own torch path" within 1e-4
  relative (bit-exact for mu-law indices).
* ``oracle.f64_chain`` -- an independent numpy float64 evaluation of the same maths
  (what the reference's librosa-based tests compare against).  Used to show that the
  CUDA path is as close to the true answer as the reference itself is.

Pinning: ``oracle/gen_golden.py`` imports the unmodified reference from
``/root/reference`` (build container only), checks that ``ref_chain`` reproduces it
bit for bit on every fixture, and writes ``tests/golden/*.npz``.  ``tests/`` re-check
the oracle against those committed vectors and against the reference's own
known-answer vectors (``tests/test_functional.py:144-158`` dB table, ``:161-203``
mu-law formulas).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.
"""
