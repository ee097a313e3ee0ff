// PCM ingest for the feature path: what sits in front of the kernels when a feature cache is
// built from a tree of wav files -- the per-file loop classifier/data.py:30-46 -> get_mfcc_feature
// (common/data_utils.py:89-97) -> librosa.load(path, sr, mono=True) + audio_to_feature's head crop (:77).
//
//   scf_wav_read_batch : RIFF/WAVE header parse + PCM read of a batch of files, a few reader threads, straight into a
//                        caller buffer [n][clip_stride] int16 with per-clip lengths (the front padding of short clips
//                        happens in the loader of the extraction kernel, so nothing is shifted on the host)
//   scf_ingest_wavs    : the whole pipeline -- reader threads fill pinned staging slots while earlier slots are
//                        uploaded, transformed and downloaded by scf_extract_host_i16_async's two device slots
//
// Resampling is out of scope (the reference leaves it to librosa): a file whose rate differs from the plan's is an error.
#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "scfeat_internal.h"

namespace scf {

int post_fail(int code, const char* msg);   // scfeat_host.cu: sets scf_last_error
int extract_host_async_on(const scf_plan* plan, const int16_t* h_in, int64_t n_clips, int64_t clip_stride, int32_t clip_len,
                          const int32_t* h_lengths, int32_t pad, float* h_out, cudaStream_t* used);   // scfeat_host.cu
int plan_sample_rate(const scf_plan* plan);
int plan_device(const scf_plan* plan);
int64_t plan_row_floats(const scf_plan* plan, int32_t clip_len);

struct WavInfo {
    int channels = 0, rate = 0, bits = 0, format = 0;
    long data_off = 0;
    long long data_bytes = 0;
};

static inline uint32_t le32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint32_t le16(const unsigned char* p) { return p[0] | (p[1] << 8); }

// Walks the RIFF chunks up to "data" (what Python's `wave` module does for the reference's files).
static bool parse_wav(FILE* f, WavInfo& w, std::string& why)
{
    unsigned char h[12];
    if (fread(h, 1, 12, f) != 12 || memcmp(h, "RIFF", 4) != 0 || memcmp(h + 8, "WAVE", 4) != 0) {
        why = "not a RIFF/WAVE file";
        return false;
    }
    bool have_fmt = false;
    for (;;) {
        unsigned char c[8];
        if (fread(c, 1, 8, f) != 8) { why = "no data chunk"; return false; }
        const uint32_t size = le32(c + 4);
        if (memcmp(c, "fmt ", 4) == 0) {
            unsigned char b[40];
            const size_t n = std::min<size_t>(size, sizeof(b));
            if (size < 16 || fread(b, 1, n, f) != n) { why = "short fmt chunk"; return false; }
            w.format = (int)le16(b);
            w.channels = (int)le16(b + 2);
            w.rate = (int)le32(b + 4);
            w.bits = (int)le16(b + 14);
            if (w.format == 0xFFFE && n >= 26) w.format = (int)le16(b + 24);      // WAVE_FORMAT_EXTENSIBLE: sub-format
            have_fmt = true;
            const long skip = (long)(size - n) + (long)(size & 1);
            if (skip && fseek(f, skip, SEEK_CUR) != 0) { why = "truncated file"; return false; }
        } else if (memcmp(c, "data", 4) == 0) {
            if (!have_fmt) { why = "data chunk before fmt chunk"; return false; }
            w.data_off = ftell(f);
            w.data_bytes = size;
            return true;
        } else {
            if (fseek(f, (long)size + (long)(size & 1), SEEK_CUR) != 0) { why = "truncated file"; return false; }
        }
    }
}

// The same chunk walk over the first bytes of a file already in memory; false when the data chunk does not start inside
// them (the caller then takes the stdio path).
static bool parse_wav_mem(const unsigned char* b, size_t n, WavInfo& w, std::string& why, bool& fatal)
{
    fatal = true;
    if (n < 12 || memcmp(b, "RIFF", 4) != 0 || memcmp(b + 8, "WAVE", 4) != 0) { why = "not a RIFF/WAVE file"; return false; }
    bool have_fmt = false;
    size_t o = 12;
    for (;;) {
        if (o + 8 > n) { fatal = false; return false; }           // header longer than what was read: not an error yet
        const uint32_t size = le32(b + o + 4);
        if (memcmp(b + o, "fmt ", 4) == 0) {
            if (size < 16) { why = "short fmt chunk"; return false; }
            if (o + 8 + std::min<size_t>(size, 40) > n) { fatal = false; return false; }
            const unsigned char* c = b + o + 8;
            w.format = (int)le16(c);
            w.channels = (int)le16(c + 2);
            w.rate = (int)le32(c + 4);
            w.bits = (int)le16(c + 14);
            if (w.format == 0xFFFE && size >= 26) w.format = (int)le16(c + 24);    // WAVE_FORMAT_EXTENSIBLE: sub-format
            have_fmt = true;
        } else if (memcmp(b + o, "data", 4) == 0) {
            if (!have_fmt) { why = "data chunk before fmt chunk"; return false; }
            w.data_off = (long)(o + 8);
            w.data_bytes = size;
            fatal = false;
            return true;
        }
        o += 8 + (size_t)size + (size & 1);
    }
}

static int read_one_stdio(const char* path, int sample_rate, int clip_len, int16_t* dst, std::vector<int16_t>& scratch, std::string& why);

// One file -> row `dst` (clip_len samples, zero filled behind the data); returns the number of valid samples or -1.
// Fast path: open + ONE read of header and samples into a per-thread buffer + close (three system calls instead of the
// five of the stdio path: fopen's fstat, a 4 KB header read and the data read), header parsed in memory.
static int read_one(const char* path, int sample_rate, int clip_len, int16_t* dst, std::vector<int16_t>& scratch, std::string& why)
{
    constexpr size_t kHead = 4096;                                  // room for the chunks in front of the samples
    const size_t cap = kHead + (size_t)clip_len * 2;
    if (scratch.size() * 2 < cap) scratch.resize((cap + 1) / 2);
    unsigned char* buf = reinterpret_cast<unsigned char*>(scratch.data());
    const int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) { why = "cannot open"; return -1; }
    size_t n = 0;
    for (;;) {                                                      // (a regular file hands everything over in one call)
        const ssize_t r = read(fd, buf + n, cap - n);
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) break;
        n += (size_t)r;
        if (n == cap) break;
    }
    close(fd);
    WavInfo w;
    bool fatal = false;
    if (!parse_wav_mem(buf, n, w, why, fatal)) {
        if (fatal) return -1;
        return read_one_stdio(path, sample_rate, clip_len, dst, scratch, why);      // long header: the general path
    }
    if (w.channels != 1 || w.format != 1 || w.bits != 16 || w.rate != sample_rate)
        return read_one_stdio(path, sample_rate, clip_len, dst, scratch, why);      // mix-down and the error texts live there
    const long long frames_in_file = w.data_bytes / 2;
    const int want = (int)std::min<long long>(frames_in_file, clip_len);
    const size_t have = n > (size_t)w.data_off ? (n - (size_t)w.data_off) / 2 : 0;
    if ((size_t)want > have && n == cap) return read_one_stdio(path, sample_rate, clip_len, dst, scratch, why);
    const int got = (int)std::min<size_t>((size_t)want, have);      // a truncated file gives what it has, like fread
    memcpy(dst, buf + w.data_off, (size_t)got * 2);
    if (got < clip_len) memset(dst + got, 0, (size_t)(clip_len - got) * 2);
    return got;
}

static int read_one_stdio(const char* path, int sample_rate, int clip_len, int16_t* dst, std::vector<int16_t>& scratch, std::string& why)
{
    FILE* f = fopen(path, "rb");
    if (!f) { why = "cannot open"; return -1; }
    WavInfo w;
    if (!parse_wav(f, w, why)) { fclose(f); return -1; }
    if (w.format != 1 || w.bits != 16) { fclose(f); why = "only 16-bit PCM wav is supported"; return -1; }
    if (w.channels < 1) { fclose(f); why = "no channels"; return -1; }
    if (w.rate != sample_rate) {
        fclose(f);
        why = "sample rate " + std::to_string(w.rate) + " != " + std::to_string(sample_rate) + " (no resampler)";
        return -1;
    }
    const long long frames_in_file = w.data_bytes / (2LL * w.channels);
    const int want = (int)std::min<long long>(frames_in_file, clip_len);        // keep the FIRST clip_len samples
    int got;
    if (w.channels == 1) {
        got = (int)fread(dst, 2, (size_t)want, f);
    } else {
        scratch.resize((size_t)want * w.channels);
        got = (int)(fread(scratch.data(), 2, scratch.size(), f) / w.channels);
        for (int i = 0; i < got; ++i) {                                         // mono mix-down: float32 mean, round half to even
            float acc = 0.f;
            for (int c = 0; c < w.channels; ++c) acc += (float)scratch[(size_t)i * w.channels + c];
            dst[i] = (int16_t)nearbyintf(acc / (float)w.channels);
        }
    }
    fclose(f);
    if (got < clip_len) memset(dst + got, 0, (size_t)(clip_len - got) * 2);
    return got;
}

// files [0, n) of `paths` into rows of h_pcm; returns the index of the first file that failed or -1
static int64_t read_range(const char* const* paths, int64_t n, int sample_rate, int clip_len, int16_t* h_pcm,
                          int64_t clip_stride, int32_t* h_lengths, int n_threads, std::string& why)
{
    n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n));
    std::atomic<int64_t> next{0}, bad{-1};
    std::string bad_why;                                                        // written by the one thread that sets `bad`
    auto work = [&]() {
        std::vector<int16_t> scratch;
        std::string w;
        for (;;) {
            const int64_t i0 = next.fetch_add(16);                              // small blocks of neighbouring files
            if (i0 >= n) break;
            for (int64_t i = i0; i < std::min<int64_t>(n, i0 + 16); ++i) {
                const int got = read_one(paths[i], sample_rate, clip_len, h_pcm + i * clip_stride, scratch, w);
                h_lengths[i] = got < 0 ? 0 : got;
                if (got < 0) {
                    int64_t expect = -1;
                    if (bad.compare_exchange_strong(expect, i)) bad_why = std::string(paths[i]) + ": " + w;
                }
            }
        }
    };
    if (n_threads == 1) {
        work();
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(work);
        for (auto& x : th) x.join();
    }
    if (bad.load() >= 0) why = bad_why;
    return bad.load();
}

}  // namespace scf

using namespace scf;

extern "C" {

int scf_wav_read_batch(const char* const* paths, int64_t n_files, int32_t sample_rate, int32_t clip_len, int16_t* h_pcm,
                       int64_t clip_stride, int32_t* h_lengths, int32_t n_threads)
{
    if (n_files < 0 || clip_len < 1 || clip_stride < clip_len) return post_fail(SCF_ERR_INVALID, "bad size");
    if (n_files == 0) return SCF_OK;
    if (!paths || !h_pcm || !h_lengths) return post_fail(SCF_ERR_INVALID, "NULL argument");
    std::string why;
    if (read_range(paths, n_files, sample_rate, clip_len, h_pcm, clip_stride, h_lengths, n_threads, why) >= 0)
        return post_fail(SCF_ERR_INVALID, why.c_str());
    return SCF_OK;
}

// The pipeline behind scf_ingest_wavs / scf_ingest_wavs_device.  Three staging slots: reader threads fill slot s (pinned
// PCM + lengths) while the earlier slots are uploaded and transformed; the rows go straight to d_out, or -- host output --
// into the slot's device buffer, from there into the slot's pinned rows, and into the caller's (pageable) array when the
// slot comes round again.  The staging buffers are allocated once per process and reused (cudaHostAlloc / cudaFreeHost
// of 3 x 16 MB took 20 - 200 ms and 6 - 430 ms per call on the B200 box -- more than reading 16,384 files, 16 - 22 ms).
namespace {

struct IngestStage {
    static constexpr int kSlots = 3;
    std::mutex mu;                      // one ingest call at a time uses the staging
    int device = -1;
    size_t pcm_bytes = 0, out_bytes = 0, len_bytes = 0;
    int16_t* pcm[kSlots] = {};          // pinned
    int32_t* len[kSlots] = {};          // pinned
    float* rows[kSlots] = {};           // pinned (host output only)
    int16_t* d_pcm[kSlots] = {};
    int32_t* d_len[kSlots] = {};
    float* d_rows[kSlots] = {};         // (host output only)
    cudaEvent_t done[kSlots] = {};
    cudaStream_t st = nullptr;

    void release()
    {
        for (int s = 0; s < kSlots; ++s) {
            if (pcm[s]) cudaFreeHost(pcm[s]);
            if (len[s]) cudaFreeHost(len[s]);
            if (rows[s]) cudaFreeHost(rows[s]);
            if (d_pcm[s]) cudaFree(d_pcm[s]);
            if (d_len[s]) cudaFree(d_len[s]);
            if (d_rows[s]) cudaFree(d_rows[s]);
            if (done[s]) cudaEventDestroy(done[s]);
            pcm[s] = nullptr; len[s] = nullptr; rows[s] = nullptr; d_pcm[s] = nullptr; d_len[s] = nullptr; d_rows[s] = nullptr;
            done[s] = nullptr;
        }
        if (st) cudaStreamDestroy(st);
        st = nullptr;
        pcm_bytes = out_bytes = len_bytes = 0;
        device = -1;
    }
    // (the device must be current)
    bool ensure(int dev, size_t need_pcm, size_t need_len, size_t need_out)
    {
        if (dev != device) { release(); device = dev; }
        bool ok = true;
        if (!st) ok = ok && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
        for (int s = 0; s < kSlots && ok; ++s)
            if (!done[s]) ok = cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming) == cudaSuccess;
        if (ok && need_pcm > pcm_bytes) {
            for (int s = 0; s < kSlots && ok; ++s) {
                if (pcm[s]) cudaFreeHost(pcm[s]);
                if (d_pcm[s]) cudaFree(d_pcm[s]);
                pcm[s] = nullptr; d_pcm[s] = nullptr;
                ok = cudaHostAlloc((void**)&pcm[s], need_pcm, cudaHostAllocDefault) == cudaSuccess &&
                     cudaMalloc((void**)&d_pcm[s], need_pcm) == cudaSuccess;
            }
            pcm_bytes = ok ? need_pcm : 0;
        }
        if (ok && need_len > len_bytes) {
            for (int s = 0; s < kSlots && ok; ++s) {
                if (len[s]) cudaFreeHost(len[s]);
                if (d_len[s]) cudaFree(d_len[s]);
                len[s] = nullptr; d_len[s] = nullptr;
                ok = cudaHostAlloc((void**)&len[s], need_len, cudaHostAllocDefault) == cudaSuccess &&
                     cudaMalloc((void**)&d_len[s], need_len) == cudaSuccess;
            }
            len_bytes = ok ? need_len : 0;
        }
        if (ok && need_out > out_bytes) {
            for (int s = 0; s < kSlots && ok; ++s) {
                if (rows[s]) cudaFreeHost(rows[s]);
                if (d_rows[s]) cudaFree(d_rows[s]);
                rows[s] = nullptr; d_rows[s] = nullptr;
                ok = cudaHostAlloc((void**)&rows[s], need_out, cudaHostAllocDefault) == cudaSuccess &&
                     cudaMalloc((void**)&d_rows[s], need_out) == cudaSuccess;
            }
            out_bytes = ok ? need_out : 0;
        }
        if (!ok) {
            cudaGetLastError();
            release();
        }
        return ok;
    }
};

IngestStage g_stage;      // lives as long as the process (no CUDA calls from a static destructor); 3 x (batch x 32 KB) pinned

int ingest_core(const scf_plan* plan, const char* const* paths, int64_t n_files, int32_t clip_len, int32_t batch,
                int32_t n_threads, float* h_out, float* d_out, int32_t* h_lengths_out)
{
    const int rate = plan_sample_rate(plan);
    const int64_t row_floats = plan_row_floats(plan, clip_len);
    const int64_t nb = std::min<int64_t>(batch, n_files);
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(plan_device(plan));
    std::lock_guard<std::mutex> lock(g_stage.mu);
    IngestStage& sg = g_stage;
    int rc = SCF_OK;
    if (!sg.ensure(plan_device(plan), (size_t)nb * clip_len * 2, (size_t)nb * 4, h_out ? (size_t)nb * row_floats * 4 : 0))
        rc = post_fail(SCF_ERR_ALLOC, "staging allocation failed");
    constexpr int kSlots = IngestStage::kSlots;
    int64_t first[kSlots] = {-1, -1, -1}, count[kSlots] = {0, 0, 0};          // the batch a slot holds
    std::string why;
    auto drain = [&](int s) -> int {          // rows of the batch in slot s: pinned -> the caller's array
        if (first[s] < 0) return SCF_OK;
        if (cudaEventSynchronize(sg.done[s]) != cudaSuccess) return post_fail(SCF_ERR_CUDA, "cudaEventSynchronize failed");
        if (h_out) memcpy(h_out + first[s] * row_floats, sg.rows[s], (size_t)count[s] * row_floats * 4);
        first[s] = -1;
        return SCF_OK;
    };
    int slot = 0;
    for (int64_t f0 = 0; f0 < n_files && rc == SCF_OK; f0 += nb, slot = (slot + 1) % kSlots) {
        const int64_t n = std::min<int64_t>(nb, n_files - f0);
        if ((rc = drain(slot)) != SCF_OK) break;             // the slot's earlier batch has left the staging buffers
        if (read_range(paths + f0, n, rate, clip_len, sg.pcm[slot], clip_len, sg.len[slot], n_threads, why) >= 0) {
            rc = post_fail(SCF_ERR_INVALID, why.c_str());
            break;
        }
        if (h_lengths_out) memcpy(h_lengths_out + f0, sg.len[slot], (size_t)n * 4);
        float* dst = h_out ? sg.d_rows[slot] : d_out + f0 * row_floats;
        if (cudaMemcpyAsync(sg.d_pcm[slot], sg.pcm[slot], (size_t)n * clip_len * 2, cudaMemcpyHostToDevice, sg.st) != cudaSuccess ||
            cudaMemcpyAsync(sg.d_len[slot], sg.len[slot], (size_t)n * 4, cudaMemcpyHostToDevice, sg.st) != cudaSuccess) {
            rc = post_fail(SCF_ERR_CUDA, "cudaMemcpyAsync failed");
            break;
        }
        rc = scf_extract_i16(plan, sg.d_pcm[slot], n, clip_len, clip_len, sg.d_len[slot], SCF_PAD_FRONT_ZERO, dst, sg.st);
        if (rc) break;
        if (h_out && cudaMemcpyAsync(sg.rows[slot], dst, (size_t)n * row_floats * 4, cudaMemcpyDeviceToHost, sg.st) != cudaSuccess) {
            rc = post_fail(SCF_ERR_CUDA, "cudaMemcpyAsync failed");
            break;
        }
        if (cudaEventRecord(sg.done[slot], sg.st) != cudaSuccess) { rc = post_fail(SCF_ERR_CUDA, "cudaEventRecord failed"); break; }
        first[slot] = f0;
        count[slot] = n;
    }
    for (int k = 0; k < kSlots; ++k) {                       // oldest first
        const int s = (slot + k) % kSlots;
        const int rd = drain(s);
        if (rc == SCF_OK) rc = rd;
    }
    if (sg.st && cudaStreamSynchronize(sg.st) != cudaSuccess && rc == SCF_OK) rc = post_fail(SCF_ERR_CUDA, "cudaStreamSynchronize failed");
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

}  // namespace

int scf_ingest_wavs(const scf_plan* plan, const char* const* paths, int64_t n_files, int32_t clip_len, int32_t batch,
                    int32_t n_threads, float* h_out, int32_t* h_lengths_out)
{
    if (!plan) return post_fail(SCF_ERR_INVALID, "plan is NULL");
    if (n_files < 0 || clip_len < 1 || batch < 1) return post_fail(SCF_ERR_INVALID, "bad size");
    if (n_files == 0) return SCF_OK;
    if (!paths || !h_out) return post_fail(SCF_ERR_INVALID, "NULL argument");
    return ingest_core(plan, paths, n_files, clip_len, batch, n_threads, h_out, nullptr, h_lengths_out);
}

// Same pipeline with the features left on the device (the training set handed to the framework through DLPack never
// visits the host): the kernel writes straight into d_out.
int scf_ingest_wavs_device(const scf_plan* plan, const char* const* paths, int64_t n_files, int32_t clip_len, int32_t batch,
                           int32_t n_threads, float* d_out, int32_t* h_lengths_out)
{
    if (!plan) return post_fail(SCF_ERR_INVALID, "plan is NULL");
    if (n_files < 0 || clip_len < 1 || batch < 1) return post_fail(SCF_ERR_INVALID, "bad size");
    if (n_files == 0) return SCF_OK;
    if (!paths || !d_out) return post_fail(SCF_ERR_INVALID, "NULL argument");
    return ingest_core(plan, paths, n_files, clip_len, batch, n_threads, nullptr, d_out, h_lengths_out);
}

}  // extern "C"
