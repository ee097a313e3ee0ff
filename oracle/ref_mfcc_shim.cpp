// C-ABI shim around the reference's own header-only C++ MFCC (inference/tflite/mfcc.h), used
// ONLY as a second oracle / CPU baseline (oracle/_ref/libref_mfcc.so).  TEST INFRASTRUCTURE.
// The header is compiled from where it lies under /root/reference (-I on the command line in
// oracle/Makefile); no reference source is copied into this repository.
// mfcc.h forgets these four standard headers, so they come first.
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <type_traits>

#include "mfcc.h"

extern "C" {

// Runs mfcc::mfcc<float> (mfcc.h:366-456) exactly as inference/tflite/speech_commands.h:293-321
// calls it.  `out` must hold ((n - window) / hop + 1) * n_coeffs floats.  Returns frame count.
int ref_mfcc(const float* audio, int n, int sample_rate, int window, int hop, int n_fft,
             int n_coeffs, int n_filt, int low_freq, int high_freq, int use_preprocess,
             float* out)
{
    if (n < window) return 0;
    std::vector<float> buf(audio, audio + n);
    std::vector<std::vector<float>> feats;
    mfcc::mfcc<float>(feats, buf, sample_rate, window, hop, n_fft, n_coeffs, n_filt,
                      low_freq, high_freq, use_preprocess != 0, false, false);
    for (size_t i = 0; i < feats.size(); ++i)
        for (int j = 0; j < n_coeffs; ++j) out[i * n_coeffs + j] = feats[i][j];
    return (int)feats.size();
}

// Same with the central-difference deltas of mfcc.h:432-441 appended: out holds frames * 2 * n_coeffs floats.
int ref_mfcc_delta(const float* audio, int n, int sample_rate, int window, int hop, int n_fft, int n_coeffs, int n_filt,
                   int low_freq, int high_freq, float* out)
{
    if (n < window) return 0;
    std::vector<float> buf(audio, audio + n);
    std::vector<std::vector<float>> feats;
    mfcc::mfcc<float>(feats, buf, sample_rate, window, hop, n_fft, n_coeffs, n_filt, low_freq, high_freq, false, true, false);
    for (size_t i = 0; i < feats.size(); ++i)
        for (int j = 0; j < 2 * n_coeffs; ++j) out[i * 2 * n_coeffs + j] = feats[i][j];
    return (int)feats.size();
}

// The triangular bank alone (mfcc.h:230-264), row-major [n_filt][n_fft/2+1] doubles.
void ref_filterbanks(int sample_rate, int n_fft, int n_filt, int low_freq, int high_freq,
                     double* out)
{
    std::vector<std::vector<double>> bank;
    mfcc::filterbanks(bank, sample_rate, n_fft, n_filt, low_freq, high_freq);
    const int bins = n_fft / 2 + 1;
    for (int i = 0; i < n_filt; ++i)
        for (int j = 0; j < bins; ++j) out[i * bins + j] = bank[i][j];
}

}  // extern "C"
