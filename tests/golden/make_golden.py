#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the REAL reference where it can be run.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

What is real reference output and what is oracle output is recorded per key:
  * example_pcm.npz      int16 PCM of /root/reference/example/*.wav (inputs; the GPU box has no
                         /root/reference, so the parity fixtures travel with the repo)
  * ref_bark.npz         outputs of the UNMODIFIED /root/reference/common/bark_feature.py
                         (imported with a stub ``librosa`` -- only its __main__ uses librosa):
                         power_spec, bark_filterbanks, bark_spec, bfcc_spec
  * ref_bark_edges.npz   bark_filterbanks of the same unmodified file with low_freq / high_freq given
                         (written by bark_edges(); `python tests/golden/make_golden.py edges` makes only this one)
  * ref_mfcc_cpp.npz     outputs of the reference's C++ twin inference/tflite/mfcc.h compiled into
                         oracle/_ref/libref_mfcc.so (float32 I/O, double inside)
  * oracle_mfcc.npz      float64 output of oracle/sonopy.py (the restatement of un-vendored sonopy)
                         -- a regression pin of the oracle itself, NOT reference output
"""
import ctypes
import glob
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

from oracle import pipeline, sonopy as osonopy  # noqa: E402


def load_real_bark():
    sys.modules.setdefault('librosa', types.ModuleType('librosa'))
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_bark_feature', os.path.join(REF, 'common/bark_feature.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_ref_cpp():
    lib = ctypes.CDLL(os.path.join(ROOT, 'oracle/_ref/libref_mfcc.so'))
    lib.ref_mfcc.restype = ctypes.c_int
    lib.ref_mfcc.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_int] * 9 + [ctypes.c_void_p]
    lib.ref_filterbanks.restype = None
    lib.ref_filterbanks.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p]
    return lib


def cpp_mfcc(lib, audio_f32, sr, W, H, nfft, ncoef, nfilt, pre=0):
    audio_f32 = np.ascontiguousarray(audio_f32, dtype=np.float32)
    k = (len(audio_f32) - W) // H + 1
    out = np.zeros((k, ncoef), dtype=np.float32)
    got = lib.ref_mfcc(audio_f32.ctypes.data, len(audio_f32), sr, W, H, nfft, ncoef, nfilt, 0, sr, pre,
                       out.ctypes.data)
    assert got == k
    return out


EDGE_CASES = [(20, 512, 300, 6000, 'constant'), (24, 1024, 100, None, 'ascendant'), (26, 512, 0, 4000, 'descendant'),
              (13, 1024, 50.5, 7600, 'constant')]


def bark_edges():
    """bark_filterbanks(low_freq, high_freq) of the unmodified common/bark_feature.py:93-136."""
    rb = load_real_bark()
    out = {}
    for nf, nfft, lo, hi, scale in EDGE_CASES:
        out['bank_%d_%d_%s_%s_%s' % (nf, nfft, lo, hi, scale)] = rb.bark_filterbanks(
            nfilts=nf, nfft=nfft, sample_rate=16000, low_freq=lo, high_freq=hi, scale=scale)
    np.savez_compressed(os.path.join(HERE, 'ref_bark_edges.npz'), **out)


def main():
    bark_edges()
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(REF, 'example/*.wav')))
    pcm = np.stack([pipeline.read_wav_int16(os.path.join(REF, 'example', n + '.wav'))[0] for n in names])
    assert pcm.shape == (8, 16000) and pcm.dtype == np.int16
    np.savez_compressed(os.path.join(HERE, 'example_pcm.npz'), names=np.array(names), pcm=pcm)
    audio = pcm.astype(np.float32) / 32768.0

    rng = np.random.default_rng(7)
    synth = np.clip(np.round(rng.normal(0, 3000, size=(2, 16000))), -32768, 32767).astype(np.int16)
    synth_f = synth.astype(np.float32) / 32768.0
    long_pcm = rng.integers(-32768, 32768, size=48000, dtype=np.int16)   # 3 s, config-5 shaped
    long_f = long_pcm.astype(np.float32) / 32768.0

    # ---- real bark_feature.py -------------------------------------------------------------
    rb = load_real_bark()
    out = {'synth_pcm': synth, 'long_pcm': long_pcm}
    for (nf, nfft, scale) in [(20, 512, 'constant'), (20, 1024, 'constant'), (24, 512, 'constant'),
                              (24, 1024, 'constant'), (26, 512, 'constant'), (26, 1024, 'constant'),
                              (22, 512, 'ascendant'), (22, 512, 'descendant')]:
        out['bank_%d_%d_%s' % (nf, nfft, scale)] = rb.bark_filterbanks(
            nfilts=nf, nfft=nfft, sample_rate=16000, low_freq=0, high_freq=None, scale=scale)
    out['power_1024_512_1024'] = np.stack([rb.power_spec(a, (1024, 512), 1024) for a in audio[:2]])
    out['power_160_80_512'] = rb.power_spec(audio[0], (160, 80), 512)
    out['power_1200_400_1024'] = rb.power_spec(audio[0], (1200, 400), 1024)   # window > fft: crop
    out['bfcc_1024_512_1024_20_20'] = np.stack([rb.bfcc_spec(a, 16000, 1024, 512, 1024, 20, 20) for a in audio])
    out['bfcc_1024_512_1024_26_13'] = np.stack([rb.bfcc_spec(a, 16000, 1024, 512, 1024, 26, 13) for a in audio])
    out['bfcc_512_256_512_26_13'] = np.stack([rb.bfcc_spec(a, 16000, 512, 256, 512, 26, 13) for a in audio])
    out['bark_1024_512_1024_24'] = np.stack([rb.bark_spec(a, 16000, 1024, 512, 1024, 24) for a in audio])
    out['bark_1024_512_1024_20'] = np.stack([rb.bark_spec(a, 16000, 1024, 512, 1024, 20) for a in audio])
    out['bark_512_256_512_24'] = np.stack([rb.bark_spec(a, 16000, 512, 256, 512, 24) for a in audio])
    out['synth_bfcc_1024_512_1024_26_13'] = np.stack([rb.bfcc_spec(a, 16000, 1024, 512, 1024, 26, 13) for a in synth_f])
    out['long_bfcc_1024_512_1024_26_13'] = rb.bfcc_spec(long_f, 16000, 1024, 512, 1024, 26, 13)
    out['long_bark_1024_512_1024_24'] = rb.bark_spec(long_f, 16000, 1024, 512, 1024, 24)
    np.savez_compressed(os.path.join(HERE, 'ref_bark.npz'), **out)

    # ---- reference C++ twin ---------------------------------------------------------------
    lib = load_ref_cpp()
    cpp = {}
    cpp['mfcc_params_json'] = np.stack([cpp_mfcc(lib, a, 16000, 1024, 512, 1024, 20, 20) for a in audio])
    cpp['mfcc_params_json_synth'] = np.stack([cpp_mfcc(lib, a, 16000, 1024, 512, 1024, 20, 20) for a in synth_f])
    cpp['mfcc_512_256_512_20_13'] = np.stack([cpp_mfcc(lib, a, 16000, 512, 256, 512, 13, 20) for a in audio[:2]])
    # a mel grid with repeated points (n_fft 512, 40 filters: empty filters -> log(eps)); the twin has no de-duplication
    cpp['mfcc_512_256_512_40_13_dupgrid'] = np.stack([cpp_mfcc(lib, a, 16000, 512, 256, 512, 13, 40) for a in audio[:2]])
    # ... and one whose first filter is empty altogether (n_fft 256, 40 filters: grid 0, 0, 0, 1, 2, 2, ...)
    cpp['mfcc_256_128_256_40_13_dupgrid'] = np.stack([cpp_mfcc(lib, a, 16000, 256, 128, 256, 13, 40) for a in audio[:2]])
    cpp['mfcc_preproc'] = np.stack([cpp_mfcc(lib, a, 16000, 1024, 512, 1024, 20, 20, pre=1) for a in audio[:2]])
    lib.ref_mfcc_delta.restype = ctypes.c_int
    lib.ref_mfcc_delta.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_int] * 8 + [ctypes.c_void_p]
    a0 = np.ascontiguousarray(audio[0], dtype=np.float32)
    dl = np.zeros((30, 40), dtype=np.float32)
    assert lib.ref_mfcc_delta(a0.ctypes.data, len(a0), 16000, 1024, 512, 1024, 20, 20, 0, 16000, dl.ctypes.data) == 30
    cpp['mfcc_central_delta'] = dl
    bank = np.zeros((20, 513))
    lib.ref_filterbanks(16000, 1024, 20, 0, 16000, bank.ctypes.data)
    cpp['bank_16000_20_1024'] = bank
    bank = np.zeros((40, 257))
    lib.ref_filterbanks(16000, 512, 40, 0, 16000, bank.ctypes.data)
    cpp['bank_16000_40_512'] = bank
    bank = np.zeros((40, 129))
    lib.ref_filterbanks(16000, 256, 40, 0, 16000, bank.ctypes.data)
    cpp['bank_16000_40_256'] = bank
    np.savez_compressed(os.path.join(HERE, 'ref_mfcc_cpp.npz'), **cpp)

    # ---- oracle regression pin ------------------------------------------------------------
    orc = {}
    orc['mfcc_params_json'] = np.stack([osonopy.mfcc_spec(a, 16000, (1024, 512), 1024, 20, 20) for a in audio])
    orc['mel_params_json'] = np.stack([osonopy.mel_spec(a, 16000, (1024, 512), 1024, 20) for a in audio])
    orc['mfcc_sonopy_defaults'] = np.stack([osonopy.mfcc_spec(a, 16000) for a in audio[:2]])
    np.savez_compressed(os.path.join(HERE, 'oracle_mfcc.npz'), **orc)

    d = np.abs(orc['mfcc_params_json'] - cpp['mfcc_params_json']).max()
    print('oracle sonopy restatement vs compiled mfcc.h on 8 example wavs: max|diff| = %.3g' % d)
    print('right_1 frame0 c0..4 (cpp):', cpp['mfcc_params_json'][names.index('right_1'), 0, :5])
    for f in ('example_pcm.npz', 'ref_bark.npz', 'ref_mfcc_cpp.npz', 'oracle_mfcc.npz'):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')


if __name__ == '__main__':
    if sys.argv[1:] == ['edges']:
        bark_edges()
    else:
        main()


def load_reference_postprocess():
    """ThresholdDecoder / TriggerDetector taken verbatim from /root/reference/listen.py by AST extraction (the module
    itself imports pyaudio / tensorflow / MNN and cannot be imported here)."""
    import ast
    import math
    src = open(os.path.join(REF, 'listen.py')).read()
    tree = ast.parse(src)
    ns = {'np': np, 'math': math}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in ('ThresholdDecoder', 'TriggerDetector'):
            exec(compile(ast.Module(body=[node], type_ignores=[]), 'listen.py', 'exec'), ns)
    return ns['ThresholdDecoder'], ns['TriggerDetector']


def make_postprocess_golden():
    TD, TR = load_reference_postprocess()
    rng = np.random.default_rng(21)
    out = {}
    raws = np.concatenate([[0.0, 1.0, 1e-12, 1 - 1e-12, 0.5, 0.2, 0.8], rng.random(400), rng.random(100) ** 8,
                           1 - rng.random(100) ** 8])
    out['raw'] = raws
    for name, cfg, center in (('default', ((6, 4),), 0.2), ('two', ((6, 4), (-2, 3)), 0.5), ('narrow', ((0, 0.1),), 0.5)):
        d = TD(cfg, center)
        out['decode_' + name] = np.array([d.decode(float(r)) for r in raws])
        out['encode_' + name] = np.array([d.encode(float(t)) for t in np.linspace(0.01, 0.99, 50)])
    classes = ['background', 'up', 'down', 'left']
    T = 600
    idx = rng.integers(0, 4, size=T)
    idx[100:160] = 2
    idx[300:420] = 1
    score = rng.random(T)
    score[100:160] = 0.9
    score[300:420] = 0.95
    for chunk in (1024, 1600, 4096):
        det = TR(chunk, classes, 0.5, 3)
        out['trigger_%d' % chunk] = np.array([det.update(int(i), float(s)) for i, s in zip(idx, score)])
    out['trigger_idx'], out['trigger_score'] = idx, score
    np.savez_compressed(os.path.join(HERE, 'ref_postprocess.npz'), **out)
    print('ref_postprocess.npz', os.path.getsize(os.path.join(HERE, 'ref_postprocess.npz')) // 1024, 'KiB')


if __name__ == '__main__' and os.environ.get('SCF_GOLDEN_POSTPROCESS', '1') == '1':
    make_postprocess_golden()
