/*
 * scfeat.h -- C ABI of libscfeat.so: batched speech-feature extraction on NVIDIA B200 (sm_100a).
 *
 *   int16/float PCM -> [pre-emphasis] -> framing/[window] -> real-FFT power spectrum
 *   -> mel / Bark filterbank -> log -> DCT-II (MFCC / BFCC)
 *
 * Drop-in boundary for ONE hot path of david8862/tf-keras-speech-commands.  The reference has
 * no FFI for this path (its boundary is a set of Python functions and a header-only C++ twin),
 * so every entry point below names the reference interface it stands in for
 * (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes / C++ binding a
 * maintainer of the reference would add.
 *
 * Conventions: every function returns 0 (SCF_OK) or a negative scf_status; nothing throws
 * across the ABI; scf_last_error() returns a thread-local message for the last failure.
 * Plans are immutable after creation and may be shared between threads; stream objects are
 * single-owner.  "d_" pointers are device pointers on the plan's device, "h_" are host pointers.
 * `cuda_stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Device
 * entry points are stream-ordered and never synchronise the host.
 */
#ifndef SCFEAT_H
#define SCFEAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCF_VERSION 100

typedef enum scf_status {
    SCF_OK = 0,
    SCF_ERR_INVALID = -1,      /* bad argument / unsupported configuration            */
    SCF_ERR_CUDA = -2,         /* a CUDA runtime call failed (message has the detail) */
    SCF_ERR_NO_DEVICE = -3,    /* no usable CUDA device: there is NO CPU fallback      */
    SCF_ERR_ALLOC = -4,
    SCF_ERR_NCCL = -5
} scf_status;

/* Filterbank family.
 * SCF_BANK_MEL_SONOPY : sonopy.filterbanks(sample_rate, n_filt, n_fft/2+1) as used through
 *                       common/data_utils.py:69; C++ twin inference/tflite/mfcc.h:230-264 called
 *                       with low=0, high=sample_rate (inference/tflite/speech_commands.h:304-307).
 * SCF_BANK_BARK_REF   : common/bark_feature.py:92-136 bark_filterbanks (incl. its quirk of mapping
 *                       bins with nfft=512 / 16 kHz whatever the caller passes).
 * SCF_BANK_CUSTOM     : caller-supplied dense [n_filt][n_fft/2+1] doubles. */
typedef enum scf_bank_kind { SCF_BANK_MEL_SONOPY = 0, SCF_BANK_BARK_REF = 1, SCF_BANK_CUSTOM = 2 } scf_bank_kind;

/* bark_filterbanks(scale=...) common/bark_feature.py:117-129 */
typedef enum scf_bank_scale { SCF_SCALE_CONSTANT = 0, SCF_SCALE_ASCENDANT = 1, SCF_SCALE_DESCENDANT = 2 } scf_bank_scale;

/* What one output row holds.
 * SCF_OUT_POWER    : power_spec()            common/bark_feature.py:85-89   -> n_fft/2+1 columns
 * SCF_OUT_LOG_BANK : mel_spec()/bark_spec()  common/bark_feature.py:139-153 -> n_filt columns
 * SCF_OUT_CEPSTRUM : mfcc_spec()/bfcc_spec() common/bark_feature.py:156-175 -> min(n_filt,n_coeffs)
 *                    columns, column 0 replaced by the log frame energy. */
typedef enum scf_output_kind { SCF_OUT_POWER = 0, SCF_OUT_LOG_BANK = 1, SCF_OUT_CEPSTRUM = 2 } scf_output_kind;

/* Analysis window.  The Python path uses none (rectangular); Hamming is the C++ twin's optional
 * use_preprocess branch, inference/tflite/mfcc.h:394-410 (denominator window-1). */
typedef enum scf_window_kind { SCF_WIN_RECT = 0, SCF_WIN_HAMMING = 1, SCF_WIN_HANN = 2 } scf_window_kind;

/* How a clip shorter than clip_len is treated.
 * SCF_PAD_FRONT_ZERO : audio_to_feature, common/data_utils.py:73-86 -- zeros in FRONT, every clip
 *                      yields frames_per_clip(clip_len) rows.
 * SCF_PAD_NONE       : vectorize_raw, common/data_utils.py:61-70 -- only frames that fit inside the
 *                      clip's own length are produced; the remaining rows are left untouched. */
typedef enum scf_pad_kind { SCF_PAD_FRONT_ZERO = 0, SCF_PAD_NONE = 1 } scf_pad_kind;

/* Delta features appended to every output row (on the frame axis of each clip), computed on the device right after
 * the extraction:
 * SCF_DELTA_DIFF     : add_deltas, common/data_utils.py:50-58 -- d[i] = f[i] - f[i-1], d[0] = 0     -> 2x columns
 * SCF_DELTA_CENTRAL  : mfcc::mfcc(use_delta), inference/tflite/mfcc.h:432-441 -- (f[i+1] - f[i-1]) / 2 with the edges
 *                      clamped                                                                        -> 2x columns
 * SCF_DELTA_CENTRAL2 : ... plus use_delta2, mfcc.h:443-453: the same difference of the delta columns  -> 3x columns */
typedef enum scf_delta_kind { SCF_DELTA_NONE = 0, SCF_DELTA_DIFF = 1, SCF_DELTA_CENTRAL = 2, SCF_DELTA_CENTRAL2 = 3 } scf_delta_kind;

typedef struct scf_config {
    int32_t sample_rate;     /* classifier/params.py:52                    default 16000 */
    int32_t window;          /* window_samples, classifier/params.py:70-73 default 1024  */
    int32_t hop;             /* hop_samples,    classifier/params.py:75-78 default 512   */
    int32_t n_fft;           /* 256, 512 or 1024                           default 1024  */
    int32_t n_filt;          /* <= 64                                      default 20    */
    int32_t n_coeffs;        /* n_mfcc                                     default 20    */
    int32_t bank;            /* scf_bank_kind                                           */
    int32_t bank_scale;      /* scf_bank_scale (Bark only)                               */
    int32_t output;          /* scf_output_kind                                         */
    int32_t window_fn;       /* scf_window_kind                            default RECT  */
    float   preemph_alpha;   /* 0 = off; mfcc.h:396 uses 0.95; x[-1] := 0                */
    float   pcm_scale;       /* int16 -> float factor, 1/32768 (common/data_utils.py:21); ignored for float input */
    int32_t device;          /* CUDA device ordinal, -1 = current                        */
    int32_t delta;           /* scf_delta_kind (pr.use_delta -> SCF_DELTA_DIFF)          default NONE  */
    const double* custom_bank; /* SCF_BANK_CUSTOM only                                   */
    float   bank_low_hz;     /* SCF_BANK_BARK_REF: low_freq / high_freq of bark_filterbanks (common/bark_feature.py:93,  */
    float   bank_high_hz;    /*   104-105); 0 = not given -> 0 Hz / sample_rate / 2 like the reference's `x or default`    */
} scf_config;

typedef struct scf_plan scf_plan;
typedef struct scf_stream scf_stream;

/* ---- host-only helpers (no GPU needed) ------------------------------------------------- */

/* Fills *cfg with configs/params.json (== classifier/params.py:99-103) MFCC settings. */
int scf_config_default(scf_config* cfg);

/* Number of frames chop_array yields for n_samples (common/bark_feature.py:80-82). */
int64_t scf_num_frames(int64_t n_samples, int32_t window, int32_t hop);

/* Output columns for this configuration (see scf_output_kind), delta columns included. */
int32_t scf_out_cols(const scf_config* cfg);

/* Dense float64 filterbank [n_filt][n_fft/2+1] exactly as the reference builds it. */
int scf_build_bank(const scf_config* cfg, double* bank_out);

/* DCT-II ortho matrix [n_filt][min(n_filt,n_coeffs)], scipy.fftpack.dct(norm='ortho') /
 * inference/tflite/mfcc.h:42-71. */
int scf_build_dct(int32_t n_filt, int32_t n_coeffs, double* dct_out);

/* Test hook (host only, no GPU): evaluates the filterbank through the kernel's own work list -- the dense bank cut
 * into two-filter tasks, runs and partial sums exactly as the bank phase walks them -- for one frame pair with power
 * spectra power_a / power_b [n_fft/2+1].  sums_a / sums_b [n_filt] must equal bank @ power.
 * stats4 (nullable): {tasks, partial-sum rows, thread groups, longest group list}. */
int scf_bank_apply_tasks(const scf_config* cfg, const double* power_a, const double* power_b, double* sums_a,
                         double* sums_b, int32_t* stats4);

/* ---- plan ------------------------------------------------------------------------------ */

/* Builds bank / DCT / twiddle / window tables in float64 on the host, uploads them.
 * Fails with SCF_ERR_NO_DEVICE when no CUDA device is present. */
int scf_plan_create(const scf_config* cfg, scf_plan** plan_out);
void scf_plan_destroy(scf_plan* plan);
int scf_plan_config(const scf_plan* plan, scf_config* cfg_out);

/* ---- batched extraction: replaces the serial loop classifier/data.py:39-43 over
 *      get_mfcc_feature (common/data_utils.py:89-97) and sonopy.mfcc_spec / bark_feature.*_spec -- */

/* Clip i occupies d_pcm[i*clip_stride .. i*clip_stride + len_i) with len_i = d_lengths ? min(d_lengths[i],
 * clip_len) : clip_len.  Row (i, f) of the output is written at d_out + (i*frames_per_clip + f)*out_cols
 * where frames_per_clip = scf_num_frames(clip_len, window, hop). */
int scf_extract_i16(const scf_plan* plan, const int16_t* d_pcm, int64_t n_clips, int64_t clip_stride,
                    int32_t clip_len, const int32_t* d_lengths, int32_t pad, float* d_out, void* cuda_stream);

/* Same for float audio already scaled to [-1, 1) (what the reference's Python functions take). */
int scf_extract_f32(const scf_plan* plan, const float* d_audio, int64_t n_clips, int64_t clip_stride,
                    int32_t clip_len, const int32_t* d_lengths, int32_t pad, float* d_out, void* cuda_stream);

/* Host-buffer convenience wrappers (staging, H2D, kernel, D2H, synchronous): what the numpy-in / numpy-out drop-in
 * functions call.  h_lengths may be NULL.  A call whose input + output fit into 96 KB (one clip, as the reference calls
 * its feature functions) goes through pinned device-mapped memory instead of copies: one launch, one synchronisation. */
int scf_extract_host_i16(const scf_plan* plan, const int16_t* h_pcm, int64_t n_clips, int64_t clip_stride,
                         int32_t clip_len, const int32_t* h_lengths, int32_t pad, float* h_out);
int scf_extract_host_f32(const scf_plan* plan, const float* h_audio, int64_t n_clips, int64_t clip_stride,
                         int32_t clip_len, const int32_t* h_lengths, int32_t pad, float* h_out);

/* Asynchronous form of scf_extract_host_i16 for producers that keep feeding batches: returns as soon as the copies
 * and the kernel are enqueued.  Two internal staging slots alternate, so the H2D copy of call i+1 overlaps the
 * kernel and the D2H copy of call i.  h_pcm / h_out must stay valid (and should be pinned) until scf_host_sync()
 * returns; results of all earlier async calls are complete after scf_host_sync(). */
int scf_extract_host_i16_async(const scf_plan* plan, const int16_t* h_pcm, int64_t n_clips, int64_t clip_stride,
                               int32_t clip_len, const int32_t* h_lengths, int32_t pad, float* h_out);
int scf_host_sync(const scf_plan* plan);

/* Library-owned output handed over as a DLPack capsule payload: *dl_out is a DLManagedTensor*
 * (kDLCUDA, float32, shape [n_clips, frames_per_clip, out_cols]) whose deleter frees the device
 * buffer; consumable by tf.experimental.dlpack.from_dlpack / torch.from_dlpack.  The buffer comes from the stream-
 * ordered allocator (cudaMallocAsync on cuda_stream), so the call only enqueues work. */
int scf_extract_i16_dlpack(const scf_plan* plan, const int16_t* d_pcm, int64_t n_clips, int64_t clip_stride,
                           int32_t clip_len, const int32_t* d_lengths, int32_t pad, void** dl_out,
                           void* cuda_stream);

/* Device-resident feature sets for the framework (classifier/data.py:97-120 loads every feature into memory for
 * model.fit): float32 kDLCUDA tensors of 1..4 dimensions, e.g. [N, n_features, feature_size, 1].
 * scf_dlpack_alloc : a library-owned buffer (freed by the DLPack deleter); *d_ptr_out (nullable) is its device pointer,
 *                    to be filled by scf_extract_* / scf_ingest_wavs_device before the tensor is handed over.
 * scf_dlpack_wrap  : a descriptor over memory the CALLER owns (e.g. the gathered multi-GPU cache); the deleter calls
 *                    release(release_ctx) once instead of freeing. */
int scf_dlpack_alloc(int32_t device, const int64_t* shape, int32_t ndim, void** dl_out, void** d_ptr_out);
int scf_dlpack_wrap(void* d_ptr, int32_t device, const int64_t* shape, int32_t ndim, void (*release)(void*),
                    void* release_ctx, void** dl_out);

/* Python-only helper: wraps a DLManagedTensor* from scf_extract_i16_dlpack in a PyCapsule named "dltensor"
 * whose destructor calls the tensor's deleter unless a consumer took ownership (renamed it "used_dltensor").
 * Returns a new PyObject* reference (NULL on failure).  Resolves the CPython C-API symbols from the running
 * interpreter with dlsym, so the library does not link against libpython; call with the GIL held
 * (ctypes.PyDLL). */
void* scf_dlpack_make_capsule(void* dl_managed_tensor);

/* ---- streaming: replaces Listener.update_vectors, listen.py:96-114 (C++ twin
 *      inference/tflite/speech_commands.h:355-449), for n_streams concurrent listeners -------- */

/* ring_rows = pr.n_features (classifier/params.py:65-68); max_chunk = largest chunk (samples) a
 * push may carry.  Rings start as zeros, carries empty (listen.py:90-92). */
int scf_stream_create(const scf_plan* plan, int32_t n_streams, int32_t ring_rows, int32_t max_chunk,
                      scf_stream** stream_out);
void scf_stream_destroy(scf_stream* s);
int scf_stream_reset(scf_stream* s, void* cuda_stream);

/* d_chunks: [n_streams][chunk_len] int16.  After the call (stream-ordered) d_ring_out, if not NULL,
 * holds [n_streams][ring_rows][out_cols] oldest->newest; d_new_rows, if not NULL, the number of frames
 * each stream emitted this step.  One kernel launch per push; the object's device state is double buffered
 * and flipped at enqueue time, so the pushes (and resets) of one scf_stream must be issued in order on ONE
 * cuda_stream (or be ordered by the caller). */
int scf_stream_push_i16(scf_stream* s, const int16_t* d_chunks, int32_t chunk_len, float* d_ring_out,
                        int32_t* d_new_rows, void* cuda_stream);
int scf_stream_push_host_i16(scf_stream* s, const int16_t* h_chunks, int32_t chunk_len, float* h_ring_out,
                             int32_t* h_new_rows);

/* ---- PCM ingest: replaces the per-file loop classifier/data.py:30-46 -> get_mfcc_feature (common/data_utils.py:89-97:
 *      librosa.load(path, sr, mono=True) + audio_to_feature's head crop :77) for files already at the plan's rate ---- */

/* RIFF/WAVE header parse + PCM read of n_files 16-bit PCM wav files with n_threads reader threads (host only, no GPU).
 * File i goes to h_pcm[i*clip_stride ..): its FIRST min(frames, clip_len) samples (multi-channel files are mixed down to
 * mono), zeros behind them; h_lengths[i] = the number of valid samples.  Fails (message names the file) on a file that
 * is unreadable, not 16-bit PCM, or whose rate differs from sample_rate (no resampler). */
int scf_wav_read_batch(const char* const* paths, int64_t n_files, int32_t sample_rate, int32_t clip_len, int16_t* h_pcm,
                       int64_t clip_stride, int32_t* h_lengths, int32_t n_threads);

/* wav files -> feature rows, pipelined: reader threads fill pinned staging slots of `batch` clips while earlier slots are
 * uploaded, transformed and downloaded by scf_extract_host_i16_async's two device slots; short clips are padded with
 * zeros in FRONT by the kernel (SCF_PAD_FRONT_ZERO).  h_out: [n_files][frames(clip_len)][out_cols] (pinned or pageable);
 * h_lengths_out (nullable): the valid samples of every file.  Synchronous: everything is complete on return. */
int scf_ingest_wavs(const scf_plan* plan, const char* const* paths, int64_t n_files, int32_t clip_len, int32_t batch,
                    int32_t n_threads, float* h_out, int32_t* h_lengths_out);

/* Same pipeline with the features left on the device: d_out [n_files][frames(clip_len)][out_cols] on the plan's device
 * (e.g. the buffer of scf_dlpack_alloc).  Synchronous. */
int scf_ingest_wavs_device(const scf_plan* plan, const char* const* paths, int64_t n_files, int32_t clip_len, int32_t batch,
                           int32_t n_threads, float* d_out, int32_t* h_lengths_out);

/* ---- streaming post-processing on the device: replaces the per-chunk scalar work listen.py does after the model,
 *      ThresholdDecoder (listen.py:452-521; C++ twin inference/tflite/threshold_decoder.h:19-113) and
 *      TriggerDetector (listen.py:525-559; twin inference/tflite/speech_commands.h:263-289), for n_streams
 *      concurrent listeners.  Arithmetic is float64 like the Python path. ------------------------------------ */
typedef struct scf_post scf_post;

/* ThresholdDecoder.__init__ (listen.py:466-471, 519-521), host only: min_out / max_out and the cumulative table
 * cd = cumsum(sum_i pdf(points; mu_i, std_i) / (resolution * n_mu)), points = linspace(min_out, max_out,
 * resolution * (max_out - min_out)).  mu_stds = n_mu (mu, std) pairs.  Call with cd_out = NULL to get the size in
 * *n_cd; with cd_out set, *n_cd is its capacity on entry. */
int scf_post_build_cd(const double* mu_stds, int32_t n_mu, int32_t resolution, double min_z, double max_z,
                      int32_t* min_out, int32_t* max_out, double* cd_out, int64_t* n_cd);

/* Decoder table on the device plus n_streams trigger state machines (activation = 0, record_index = None).
 * class_is_background[c] != 0 marks the 'background' class (listen.py:544); chunk_size, sensitivity and trigger_level
 * are TriggerDetector's constructor arguments.  device = -1: the current one. */
int scf_post_create(const double* mu_stds, int32_t n_mu, double center, int32_t resolution, double min_z, double max_z,
                    const uint8_t* class_is_background, int32_t n_classes, int32_t n_streams, int32_t chunk_size,
                    double sensitivity, int32_t trigger_level, int32_t device, scf_post** post_out);
void scf_post_destroy(scf_post* post);
int scf_post_reset(scf_post* post, void* cuda_stream);

/* Element-wise ThresholdDecoder.decode (listen.py:497-509) of n raw outputs; device pointers, stream-ordered. */
int scf_post_decode(const scf_post* post, const double* d_raw, int64_t n, double* d_out, void* cuda_stream);

/* One iteration of the listen.py loop after the model (listen.py:411-425) for every stream, ONE launch:
 * index = argmax(probs[s]), score = max(probs[s]); a non-background score goes through decode (the float32 score's
 * 1 / x - 1 is evaluated in float32, as numpy does for the model's output); TriggerDetector.update(index, score).
 * d_probs: [n_streams][n_classes] float32.  Outputs (each nullable): class index, decoded score, 1 where the stream
 * activated in this step. */
int scf_post_step(scf_post* post, const float* d_probs, int32_t* d_index_out, double* d_score_out, uint8_t* d_fired_out,
                  void* cuda_stream);

/* TriggerDetector.update(index, score) alone (listen.py:538-559) for every stream: class indices and already decoded
 * scores in, activation flags out; device pointers, one launch. */
int scf_post_trigger_update(scf_post* post, const int32_t* d_index, const double* d_score, uint8_t* d_fired_out,
                            void* cuda_stream);

/* Copies the trigger state to the host (synchronises cuda_stream): activation counters and recorded class
 * indices (-1 = None).  Either pointer may be NULL. */
int scf_post_state(const scf_post* post, int32_t* h_activation, int32_t* h_record_index, void* cuda_stream);
int scf_post_info(const scf_post* post, int32_t* min_out, int32_t* max_out, int64_t* n_cd);

/* ---- multi-GPU feature-cache assembly (new; the reference is single-process) ------------------ */

/* Fused extract + all-gather over NVLink peer memory: this rank extracts its n_local clips and the kernel
 * epilogue stores every row into the same slot of all `world` ranks' caches
 * (d_peer_out[r] = base of rank r's full [world*clips_per_rank, frames, cols] cache, peer-mapped; this rank's
 * block starts at row rank*clips_per_rank*frames; n_local <= clips_per_rank, the last rank may own fewer).
 * The caller owns the cross-rank barrier that follows. */
int scf_extract_i16_gather(const scf_plan* plan, const int16_t* d_pcm, int64_t n_local, int64_t clip_stride,
                           int32_t clip_len, float* const* d_peer_out, int32_t world, int32_t rank,
                           int64_t clips_per_rank, void* cuda_stream);

/* The same with ONE multicast store per row segment: d_multicast_out is a multicast address that maps the same
 * [world*clips_per_rank, frames, cols] cache on every rank (NVLink SHARP: cuMulticastCreate / cuMulticastBindMem, or
 * torch.distributed._symmetric_memory's multicast_ptr); the epilogue issues multimem.st and the NVSwitch replicates
 * each row to all ranks, so a row leaves the GPU once instead of world - 1 times.  Needs multicast support on the
 * fabric (B200 HGX: yes); the caller owns the closing barrier. */
int scf_extract_i16_gather_multicast(const scf_plan* plan, const int16_t* d_pcm, int64_t n_local, int64_t clip_stride,
                                     int32_t clip_len, float* d_multicast_out, int32_t world, int32_t rank,
                                     int64_t clips_per_rank, void* cuda_stream);

/* Plain NCCL all-gather of per-rank feature shards (baseline for the fused path).  nccl_comm is an
 * ncclComm_t; libnccl.so.2 is resolved with dlopen at first use. */
int scf_allgather_nccl(void* nccl_comm, const float* d_local, int64_t n_local_floats, float* d_all,
                       void* cuda_stream);

/* Device-memory and CUDA-IPC helpers so that a ctypes host can own a peer-mappable feature cache without
 * any other CUDA binding: scf_device_malloc returns plain cudaMalloc memory (exportable); scf_ipc_export /
 * scf_ipc_import wrap cudaIpcGetMemHandle / cudaIpcOpenMemHandle (64-byte handle, peer access enabled lazily).
 * scf_memcpy: kind 0 = host->device, 1 = device->host, 2 = device->device; synchronous when cuda_stream is
 * NULL. */
int scf_device_malloc(int32_t device, int64_t bytes, void** d_ptr_out);
int scf_device_free(int32_t device, void* d_ptr);
int scf_memcpy(int32_t device, void* dst, const void* src, int64_t bytes, int32_t kind, void* cuda_stream);
int scf_ipc_export(int32_t device, void* d_ptr, uint8_t* handle64_out);
int scf_ipc_import(int32_t device, const uint8_t* handle64, void** d_ptr_out);
int scf_ipc_close(int32_t device, void* d_ptr);

/* ---- misc ------------------------------------------------------------------------------------ */
/* dst[0 .. bytes) = src[0 .. bytes) with the library's persistent copy threads (what the host-buffer calls use to move
 * pageable caller arrays into pinned staging; SCFEAT_COPY_THREADS, default half the allowed cores, at most 8).  Host
 * only, no GPU involved; safe to call from several threads. */
int scf_parallel_memcpy(void* dst, const void* src, int64_t bytes);

const char* scf_last_error(void);
int scf_version(void);
/* Kernels launched by this library in this process so far (bench.py's gpu_launches claim). */
int64_t scf_launch_count(void);
/* Measured FP32 FMA throughput (FLOP/s) of the plan's device: micro-benchmark for the compute
 * roofline that MEASURED_PEAKS.json does not hold. */
int scf_measure_fp32_flops(int32_t device, double* flops_out);

#ifdef __cplusplus
}
#endif
#endif /* SCFEAT_H */
