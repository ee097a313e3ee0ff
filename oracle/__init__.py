"""CPU oracle for the speech-feature hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
(``tf-keras-speech-commands_b200``) never imports it and has no CPU fallback.

Parity status (see DESIGN.md "Oracle"):
  * ``oracle.bark``   -- restatement of /root/reference/common/bark_feature.py, PINNED against
    the real file (imported with a stub ``librosa``) by tests/golden/make_golden.py.
  * ``oracle.sonopy`` -- restatement of the third-party, un-vendored, unpinned PyPI package
    ``sonopy`` (requirements.txt:7; believed 0.1.2).  The reference holds no golden vectors for
    it, so it is pinned indirectly: against the in-tree C++ twin inference/tflite/mfcc.h
    compiled into oracle/_ref (max |diff| <= 6e-7 on all example/*.wav), and against
    bark_feature.py's verbatim copies of sonopy's chop_array/power_spec/safe_log.
    Mel grids with repeated points (e.g. n_fft 512 with 40 filters) follow the C++ twin, which keeps
    them (sonopy's ``correct_grid`` helper is a no-op as published, see oracle/sonopy.py mel_grid);
    that case is pinned against the twin as well (bank and features).
"""
