// Internal declarations shared by the host side (scfeat_host.cu) and the kernels
// (scfeat_kernels.cu).  Not part of the ABI -- see include/scfeat.h for that.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "scfeat.h"

namespace scf {

constexpr int kWarps = 8;                 // warps per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kCtasPerSm = 16 / kWarps;    // 16 warps per SM at <= 128 registers per thread
constexpr int kMaxPeers = 8;

// Bank phase.  The power spectra of a frame PAIR sit in one shared-memory row, interleaved per bin as
// (|A[k]|^2, |B[k]|^2), bins 0 .. n_fft/2 + 1 (the last one is a zero pad).  A task is a run of kTaskBins consecutive
// bins (first bin even -> 16-byte aligned) applied to TWO filters at once, so every power value is read once per
// filter pair and all the arithmetic is packed (FFMA2 on the (A, B) pair with a broadcast weight).
// Task word:
//   bits  0..7   first bin / 2 ( = 16-byte offset of the task inside the pair row)
//   bits  8..18  partial-sum row of the first filter      } only read when bit 31 is set; given as the row's
//   bits 19..29  partial-sum row of the second filter     } offset inside the team's exchange area, in 64-byte units
//   bit  31      last task of its run: store the two accumulated sums
// The 16 weights of task t sit at float4 index 4*t of the weight table, as (a[k], b[k], a[k+1], b[k+1]) per float4.
constexpr int kTaskBins = 8;
constexpr int pair_row_bins(int r) { return 16 * r + 2; }
constexpr int pair_row_floats(int r) { return 2 * pair_row_bins(r); }
// Exchange-area geometry shared by the kernels (Geo<R>) and the host-side task builder.  A warp's region holds
// 32 / r pairs: exchange rows of r complex values + 8 bytes of padding; after pass 2 its head takes the pairs' power
// rows and its tail `drows` partial-sum rows of `drow` floats (one packed (A, B) value per pair slot of the team).
constexpr int kTeamWarps = 8;
constexpr int xwarp_floats(int r) { return (32 / r) * 32 * (2 * r + 2); }
constexpr int p_free_floats(int r) { return ((32 / r) * pair_row_floats(r) + 28 + 15) / 16 * 16; }
constexpr int drow_floats(int r) { return 2 * kTeamWarps * (32 / r); }
constexpr int drows(int r) { return (xwarp_floats(r) - p_free_floats(r)) / drow_floats(r); }
// 64-byte-unit offset of partial-sum row `i` of warp region `w`
constexpr int partial_row_unit(int r, int w, int i) { return (w * xwarp_floats(r) + p_free_floats(r) + i * drow_floats(r)) / 16; }

// Which partial sums make up filter q: `count` consecutive rows starting at 64-byte unit `unit0`.
struct QSpec {
    int32_t unit0;
    int32_t count;
};

// One step of the streaming state machine of Listener.update_vectors (listen.py:96-114), fused into the extract
// kernel (int16 input, generic loader): "clip" s is stream s, its samples are concat(carry_in[s], chunks[s]) without
// that buffer ever being written; the kernel's rows are the k newest ring rows, and the team that owns the stream's
// first pair also writes the surviving ring rows and the new carry.  State is double buffered (in -> out), so no thread
// ever reads what another one writes in the same launch.
struct StreamStep {
    const int16_t* chunks;       // [n_streams][chunk_len]                      (listen.py:101 `chunk`)
    int32_t chunk_len;
    int32_t carry_cap;           // elements per carry buffer
    const int16_t* carry_in;     // [n_streams][carry_cap]                      (`window_audio` before the push)
    int16_t* carry_out;          //                                             (`window_audio[k*hop:]`, listen.py:106)
    const int32_t* len_in;       // [n_streams] valid samples of carry_in
    int32_t* len_out;
    const float* ring_in;        // [n_streams][ring_rows][cols]                (`mfccs` before the push)
    float* ring_out;             //                                             (listen.py:107-109)
    float* ring_copy;            // nullable: the caller's copy of the new ring
    int32_t copy_pitch;          // floats per row of ring_copy (> cols when delta columns follow)
    int32_t* n_new;              // [n_streams] frames emitted by this step
    int32_t* n_new_copy;         // nullable
    int32_t ring_rows;
};

// Everything a launch needs; passed by value (fits the 4 KB parameter space easily).
struct KParams {
    const void* in;              // int16_t* or float*
    int64_t clip_stride;         // elements between clips
    int64_t n_pairs;             // n_clips * pairs_per_clip
    int32_t clip_len;            // samples per (padded) clip
    const int32_t* lengths;      // nullable
    int32_t pad_mode;            // scf_pad_kind
    int32_t frames_per_clip;     // rows per clip in the output
    int32_t pairs_per_clip;      // ceil(frames_per_clip / 2)
    int32_t window, hop, w_eff;  // w_eff = min(window, n_fft)
    float preemph;               // 0 = off
    const float* win;            // nullable window table [w_eff]
    // output
    float* out;                  // rank-local output, or NULL when peer_out is used
    float* peer_out[kMaxPeers];  // fused all-gather targets
    int32_t n_peers;             // 0 = plain
    int32_t peer_multicast;      // 1: peer_out[0] is a multicast address of all ranks' caches (multimem.st, one store per row segment)
    int64_t peer_row0;           // first output row of this rank inside the gathered cache
    int32_t out_cols;            // columns the kernel produces per row
    int32_t out_pitch;           // floats between rows of `out` (> out_cols when delta columns follow)
    int32_t out_kind;            // scf_output_kind
    float power_scale;           // pcm_scale^2 / (4 * n_fft) (the FFT stage leaves a factor 2)
    float zero_energy;           // int16 input: frame energies below this mean "every sample was zero"
    // tables: one device blob, copied verbatim into shared memory by a single TMA bulk copy
    //   [tasks u32[n_tasks]] [task ranges int2[n_groups]] [qspec int2[n_filt]] [weights float4[4*n_tasks]]
    //   [twiddles float4[4][32] + float4[32]] [dct float[n_out][n_filt4]]           (every section 16-byte aligned)
    const void* tables;
    int32_t table_bytes;         // whole blob
    int32_t table_small_bytes;   // leading part without twiddles / DCT (what the 3-CTA-per-SM kernels stage)
    int32_t off_wts, off_tw, off_dct, off_tasks, off_tbeg, off_qspec;
    int32_t n_q;                 // n_filt (+1 when the cepstrum needs the frame energy)
    int32_t n_filt;
    int32_t n_filt4;             // n_filt rounded up to a multiple of 4
    int32_t n_out;               // cepstrum columns = min(n_filt, n_coeffs)
    int32_t frames_odd;          // frames_per_clip is odd: the last pair of a clip has no frame B
    uint32_t ppc_magic, ppc_shift;   // fast division by pairs_per_clip
    int32_t fast_path;           // 1: window == n_fft, hop == n_fft/2, full or front-padded clips (FAST kernels)
    int32_t fast_pre;            // ... with pre-emphasis and / or a window table fused into the loader (`win` is never NULL then)
    uint32_t* tile_ctr;          // nullable: next-tile counter of this launch, zero at launch (and again when it ends)
    int32_t stream_on;           // 1: `stream` describes a streaming step (in == stream.carry_in, lengths unused)
    StreamStep stream;
};

// scfeat_kernels.cu
cudaError_t launch_extract(int radix_r, bool is_f32, bool fast, const KParams& p, int64_t n_tiles,
                           int num_sms, cudaStream_t st, size_t smem_bytes);
size_t extract_smem_bytes(int radix_r, const KParams& p);
size_t extract_smem_limit(int radix_r, const KParams& p);
int pairs_per_tile(int radix_r);
int bank_groups(int radix_r);

cudaError_t launch_fp32_probe(float* out, int iters, int grid, cudaStream_t st);
// delta columns over the frame axis of finished rows, in place (scf_delta_kind); lengths / geometry give the number of
// rows each clip really has under SCF_PAD_NONE
cudaError_t launch_delta(float* rows, const int32_t* lengths, int64_t n_clips, int frames_per_clip, int cols, int pitch,
                         int kind, int clip_len, int window, int hop, int pad_mode, cudaStream_t st);

void count_launch(int n);

}  // namespace scf
