// Host side of libscfeat.so: configuration, float64 table construction (bank / DCT / twiddles /
// window exactly as the reference computes them), plan object, C ABI (include/scfeat.h).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "scfeat_dlpack.h"
#include "scfeat_internal.h"

namespace scf {

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

int post_fail(int code, const char* msg) { return fail(code, msg); }     // for scfeat_post.cu

#define SCF_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(SCF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));         \
    } while (0)

// ---------------------------------------------------------------------------------------------
// float64 table builders (host only)
// ---------------------------------------------------------------------------------------------

// numpy.linspace(start, stop, num, endpoint=True): arange(num) * step + start, last element := stop
static std::vector<double> linspace(double start, double stop, int num)
{
    std::vector<double> v(num);
    if (num == 1) { v[0] = start; return v; }
    const double step = (stop - start) / (double)(num - 1);
    for (int i = 0; i < num; ++i) v[i] = (double)i * step + start;
    v[num - 1] = stop;
    return v;
}

// sonopy.filterbanks(sample_rate, n_filt, fft_len) -- restated from its published algorithm; C++ twin
// inference/tflite/mfcc.h:230-264 with low=0, high=sample_rate (speech_commands.h:304-307).
// Repeated grid points are kept as they are (filters between them come out empty and give log(eps)): the C++ twin has
// no de-duplication, and sonopy's own `correct_grid` helper is a no-op as published -- it is handed an ndarray, so
// `[x[0] - 1] + x` broadcasts to x - 1 instead of prepending, and its offset never leaves 0.
static void build_mel_sonopy(int sample_rate, int n_filt, int n_bins, std::vector<double>& bank)
{
    auto hz2mel = [](double f) { return 1127.0 * log(1.0 + f / 700.0); };      // mfcc.h:134-138
    auto mel2hz = [](double m) { return 700.0 * (exp(m / 1127.0) - 1.0); };    // mfcc.h:141-145
    std::vector<double> mels = linspace(hz2mel(0.0), hz2mel((double)sample_rate), n_filt + 2);
    std::vector<long> grid(n_filt + 2);
    for (int i = 0; i < n_filt + 2; ++i) grid[i] = (long)(mel2hz(mels[i]) * (double)n_bins / (double)sample_rate);
    bank.assign((size_t)n_filt * n_bins, 0.0);
    for (int i = 0; i < n_filt; ++i) {
        const long l = grid[i], m = grid[i + 1], r = grid[i + 2];
        // np.linspace(0, 1, m-l, endpoint=False)[j] = j * (1/(m-l));  np.linspace(1, 0, r-m, False)[j] = 1 + j*(-1/(r-m))
        for (long j = l; j < m && j < n_bins; ++j) bank[(size_t)i * n_bins + j] = (double)(j - l) * (1.0 / (double)(m - l));
        for (long j = m; j < r && j < n_bins; ++j) bank[(size_t)i * n_bins + j] = (double)(j - m) * (-1.0 / (double)(r - m)) + 1.0;
    }
}

// common/bark_feature.py:92-136.  bark2fft / fft2bark are used with their DEFAULT nfft=512 and
// sample_rate=16000 (bark_feature.py:112,134 pass neither) -- reproduced on purpose.
static void build_bark_ref(int sample_rate, int n_filt, int n_bins, int scale, double low_hz, double high_hz,
                           std::vector<double>& bank)
{
    const double map_nfft = 512.0, map_rate = 16000.0;
    auto hz2bark = [](double f) { return 6.0 * asinh(f / 600.0); };             // :27-29
    auto bark2hz = [](double b) { return 600.0 * sinh(b / 6.0); };              // :32-34
    auto fft2bark = [&](double k) { return hz2bark((k * map_rate) / (map_nfft + 1.0)); };     // :47-49
    auto bark2fft = [&](double b) { return (map_nfft + 1.0) * bark2hz(b) / map_rate; };       // :52-56
    // band edges: `high_freq or sample_rate / 2`, `low_freq or 0` (:104-105) -- 0 means "not given"
    const double high = high_hz > 0.0 ? high_hz : (double)sample_rate / 2.0;
    const double low = low_hz > 0.0 ? low_hz : 0.0;
    std::vector<double> pts = linspace(hz2bark(low), hz2bark(high), n_filt + 4);
    std::vector<long> bins(n_filt + 4);
    for (int i = 0; i < n_filt + 4; ++i) bins[i] = (long)floor(bark2fft(pts[i]));
    bank.assign((size_t)n_filt * n_bins, 0.0);
    double c = (scale == SCF_SCALE_DESCENDANT || scale == SCF_SCALE_CONSTANT) ? 1.0 : 0.0;
    for (int i = 0; i < n_filt; ++i) {
        if (scale == SCF_SCALE_DESCENDANT) {
            c -= 1.0 / n_filt;
            c = c * (c > 0 ? 1.0 : 0.0);
        } else if (scale == SCF_SCALE_ASCENDANT) {
            c += 1.0 / n_filt;
            c = c * (c < 1 ? 1.0 : 0.0) + (c > 1 ? 1.0 : 0.0);
        }
        const double fc = pts[i + 2];
        // (the reference indexes fbank[i, j] for every j of the range and raises IndexError past the last column --
        //  only reachable with high_freq above the 8 kHz the fixed bin mapping covers; such bins are dropped here)
        for (long j = std::max(0L, bins[i]); j < bins[i + 4] && j < n_bins; ++j) {
            const double fb = fft2bark((double)j);
            double v = 0.0;                                                      // Fm, :59-72
            if (fc - 2.5 <= fb && fb <= fc - 0.5) v = pow(10.0, 2.5 * (fb - fc + 0.5));
            else if (fc - 0.5 < fb && fb < fc + 0.5) v = 1.0;
            else if (fc + 0.5 <= fb && fb <= fc + 1.3) v = pow(10.0, -2.5 * (fb - fc - 0.5));
            bank[(size_t)i * n_bins + j] = fabs(c * v);
        }
    }
}

static int check_config(const scf_config* c)
{
    if (!c) return fail(SCF_ERR_INVALID, "config is NULL");
    if (c->n_fft != 256 && c->n_fft != 512 && c->n_fft != 1024)
        return fail(SCF_ERR_INVALID, "n_fft must be 256, 512 or 1024");
    if (c->window < 1 || c->hop < 1) return fail(SCF_ERR_INVALID, "window and hop must be positive");
    if (c->sample_rate < 1) return fail(SCF_ERR_INVALID, "sample_rate must be positive");
    if (!(c->bank_low_hz >= 0.f) || !(c->bank_high_hz >= 0.f) || (c->bank_high_hz > 0.f && c->bank_high_hz <= c->bank_low_hz))
        return fail(SCF_ERR_INVALID, "bank_low_hz / bank_high_hz must be 0 (default) or 0 <= low < high");
    if (c->output < SCF_OUT_POWER || c->output > SCF_OUT_CEPSTRUM) return fail(SCF_ERR_INVALID, "bad output kind");
    if (c->output != SCF_OUT_POWER) {
        if (c->n_filt < 1 || c->n_filt > 64) return fail(SCF_ERR_INVALID, "n_filt must be in 1..64");
        if (c->bank < SCF_BANK_MEL_SONOPY || c->bank > SCF_BANK_CUSTOM) return fail(SCF_ERR_INVALID, "bad bank kind");
        if (c->bank == SCF_BANK_CUSTOM && !c->custom_bank) return fail(SCF_ERR_INVALID, "custom bank is NULL");
    }
    if (c->output == SCF_OUT_CEPSTRUM && c->n_coeffs < 1) return fail(SCF_ERR_INVALID, "n_coeffs must be positive");
    if (c->window_fn < SCF_WIN_RECT || c->window_fn > SCF_WIN_HANN) return fail(SCF_ERR_INVALID, "bad window kind");
    if (c->delta < SCF_DELTA_NONE || c->delta > SCF_DELTA_CENTRAL2) return fail(SCF_ERR_INVALID, "bad delta kind");
    if (c->delta != SCF_DELTA_NONE && c->output == SCF_OUT_POWER) return fail(SCF_ERR_INVALID, "delta features need a bank or cepstrum output");
    return SCF_OK;
}

static int build_bank(const scf_config* c, std::vector<double>& bank)
{
    const int n_bins = c->n_fft / 2 + 1;
    if (c->bank == SCF_BANK_MEL_SONOPY) build_mel_sonopy(c->sample_rate, c->n_filt, n_bins, bank);
    else if (c->bank == SCF_BANK_BARK_REF) build_bark_ref(c->sample_rate, c->n_filt, n_bins, c->bank_scale, c->bank_low_hz, c->bank_high_hz, bank);
    else bank.assign(c->custom_bank, c->custom_bank + (size_t)c->n_filt * n_bins);
    return SCF_OK;
}

// scipy.fftpack.dct(type 2, norm='ortho') as a matrix, inference/tflite/mfcc.h:56-66
static void build_dct(int n_filt, int n_out, std::vector<double>& d)
{
    d.assign((size_t)n_filt * n_out, 0.0);
    for (int n = 0; n < n_filt; ++n)
        for (int k = 0; k < n_out; ++k) {
            double v = sqrt(2.0 / n_filt) * cos(M_PI * (n + 0.5) * k / n_filt);
            if (k == 0) v *= sqrt(0.5);
            d[(size_t)n * n_out + k] = v;
        }
}

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
// Pinned, device-mapped host staging for SMALL host-buffer calls (one clip, one chunk per stream -- the reference's own
// call shape): the kernel reads its samples from this buffer over PCIe and writes its rows into it, so such a call is one
// launch and one stream synchronisation instead of copy + launch + copy (33.8 -> see DESIGN.md section 6).
struct MappedStage {
    unsigned char* h = nullptr;     // host view
    unsigned char* d = nullptr;     // device view of the same bytes
    size_t bytes = 0;
    int ensure(size_t need)
    {
        if (need <= bytes) return 0;
        release();
        if (cudaHostAlloc((void**)&h, need, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void**)&d, h, 0) != cudaSuccess) {
            cudaGetLastError();
            release();
            return -1;
        }
        bytes = need;
        return 0;
    }
    void release()
    {
        if (h) cudaFreeHost(h);
        h = d = nullptr;
        bytes = 0;
    }
};
constexpr size_t kSmallCallBytes = 96 * 1024;      // input + output of a call that takes the mapped path

// A few persistent host threads that copy between ordinary (pageable) caller memory and pinned staging: one thread moves
// ~10 GB/s, the PCIe link 55 GB/s, and cudaMemcpy from pageable memory is a single-threaded bounce copy (a 512-clip numpy
// batch went through at 0.39 M clips/s against 1.67 M from pinned memory).  SCFEAT_COPY_THREADS sets the count (default:
// half the cores the process may use, at most 8; 1 = the calling thread alone).
class CopyPool {
public:
    static CopyPool& get()
    {
        static CopyPool pool;
        return pool;
    }
    // dst[0 .. bytes) = src[0 .. bytes), split over the workers and the caller; returns when all of it is done
    void copy(void* dst, const void* src, size_t bytes)
    {
        constexpr size_t kMinPart = 256 * 1024;
        const size_t helpers = getpid() == owner_ ? workers_.size() : 0;      // (a forked child has no worker threads)
        const size_t parts = std::max<size_t>(1, std::min<size_t>(helpers + 1, bytes / kMinPart));
        if (parts == 1) {
            memcpy(dst, src, bytes);
            return;
        }
        Batch b;
        b.left = parts - 1;
        const size_t step = ((bytes + parts - 1) / parts + 63) & ~(size_t)63;
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t i = 1; i < parts; ++i) {
                const size_t o = i * step;
                if (o >= bytes) { --b.left; continue; }
                q_.push_back({static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, std::min(step, bytes - o), &b});
            }
        }
        cv_.notify_all();
        memcpy(dst, src, std::min(step, bytes));
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return b.left == 0; });
    }

private:
    struct Batch { size_t left = 0; };
    struct Job { char* d; const char* s; size_t n; Batch* b; };
    CopyPool()
    {
        int n = 0;
        if (const char* e = getenv("SCFEAT_COPY_THREADS")) n = atoi(e);
        if (n <= 0) {
            cpu_set_t set;
            int cores = (sched_getaffinity(0, sizeof(set), &set) == 0) ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
            n = std::max(1, std::min(8, cores / 2));
        }
        owner_ = getpid();
        for (int i = 1; i < n; ++i) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool()
    {
        if (getpid() != owner_) {               // a forked child inherits the objects but not the threads
            for (auto& t : workers_) t.detach();
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void run()
    {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
            if (q_.empty()) return;          // stop requested and nothing left
            Job j = q_.front();
            q_.pop_front();
            lk.unlock();
            memcpy(j.d, j.s, j.n);
            lk.lock();
            if (--j.b->left == 0) done_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::deque<Job> q_;
    std::vector<std::thread> workers_;
    pid_t owner_ = 0;
    bool stop_ = false;
};

// Pinned staging ring of the synchronous host-buffer calls when the caller's arrays are pageable
constexpr int kStageSlots = 3;
constexpr size_t kStageChunkBytes = 4u << 20;      // largest chunk (input bytes)
struct StageSlot {
    unsigned char* in = nullptr;
    unsigned char* out = nullptr;
    size_t in_bytes = 0, out_bytes = 0;
    cudaEvent_t done = nullptr;        // behind the chunk's download
};

struct Workspace {          // scratch for the host-buffer entry points
    std::mutex mu;
    MappedStage small;
    StageSlot stage[kStageSlots];
    void* d_in = nullptr;
    size_t in_bytes = 0;
    float* d_out = nullptr;
    size_t out_bytes = 0;
    int32_t* d_len = nullptr;
    size_t len_bytes = 0;
    cudaStream_t st = nullptr;
    cudaStream_t st2 = nullptr;
    cudaEvent_t ev = nullptr;
    // scf_extract_host_i16_async: two alternating slots, each with its own stream and staging buffers
    struct Slot {
        void* d_in = nullptr;
        size_t in_bytes = 0;
        float* d_out = nullptr;
        size_t out_bytes = 0;
        int32_t* d_len = nullptr;
        size_t len_bytes = 0;
        cudaStream_t st = nullptr;
    } slot[2];
    unsigned next_slot = 0;
};

}  // namespace scf

struct scf_plan {
    scf_config cfg;
    int device = 0;
    int num_sms = 0;
    int radix_r = 32;
    int n_bins = 0;
    int base_cols = 0;          // columns the extract kernel produces
    int out_cols = 0;           // columns of an output row: base_cols * (1 + delta blocks)
    float power_scale_i16 = 0.f, power_scale_f32 = 0.f;
    // device tables: the blob the kernel copies into shared memory (one per input scale), window table
    float* d_win = nullptr;
    unsigned char* d_tab_i16 = nullptr;   // bank weights pre-multiplied by the int16 power scale
    unsigned char* d_tab_f32 = nullptr;   // ... by the float-input power scale
    int table_bytes = 0, table_small_bytes = 0, off_wts = 0, off_tw = 0, off_dct = 0, off_tasks = 0, off_tbeg = 0, off_qspec = 0;
    int n_tasks = 0, n_q = 0, n_dst = 0, n_filt4 = 0, n_out = 0;
    // tile counters of the dynamic schedule: a ring of zeroed words; every fast-path launch takes the next one and the
    // team that draws the launch's last number zeroes it again (a word comes round again 16384 launches later)
    uint32_t* d_tile_ctr = nullptr;
    mutable std::atomic<uint32_t> ctr_next{0};
    scf::Workspace ws;
};
constexpr uint32_t kTileCtrRing = 16384;

struct scf_stream {
    const scf_plan* plan = nullptr;
    // double-buffered state: a push reads set `cur` and writes the other one (scfeat_internal.h StreamStep)
    int16_t* carry[2] = {nullptr, nullptr};     // [n_streams][carry_cap]
    int32_t* carry_len[2] = {nullptr, nullptr}; // [n_streams]
    float* ring[2] = {nullptr, nullptr};        // [n_streams][ring_rows][cols] (base columns: the state carries no deltas)
    float* ring_wide = nullptr;                 // delta plans: the host push's [n_streams][ring_rows][out_cols] copy
    int32_t* n_new = nullptr;                   // [n_streams]
    int cur = 0;
    int32_t n_streams = 0, carry_cap = 0, ring_rows = 0, cols = 0;
    int max_chunk = 0;
    int16_t* d_chunk_stage = nullptr;     // for the host-buffer push
    scf::MappedStage small;               // ... of few streams (chunk in, ring + counters out through mapped memory)
    unsigned char* pin = nullptr;         // ... of many streams: pinned chunk / ring / counter staging
    size_t pin_bytes = 0;
};

namespace scf {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename T>
static int upload(T** dptr, const std::vector<T>& host)
{
    const size_t bytes = std::max<size_t>(host.size(), 1) * sizeof(T);
    SCF_CUDA(cudaMalloc((void**)dptr, bytes));
    if (!host.empty()) SCF_CUDA(cudaMemcpy(*dptr, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    return SCF_OK;
}

// Turns the dense float64 bank into the bank phase's work list (format: scfeat_internal.h).
//  1. The bin axis is cut at every filter's first and one-past-last non-zero bin: inside such a segment the set of
//     active filters is constant (two for a triangular mel bank, about four for the Bark bank).
//  2. The active filters of a segment are taken two at a time; each filter pair covers the segment with tasks of
//     kTaskBins bins (first bin even, pulled back at the end of the row, zero weights outside the segment).
//  3. The tasks of one (segment, filter pair) form a run that accumulates in registers; long runs are split so that no
//     run exceeds the per-group average, and runs are spread over the thread groups longest first.
//  4. Every (run, filter) owns a partial-sum row in the exchange area; a filter's rows are consecutive inside one warp
//     region's tail, so the log phase walks them with a fixed stride.
struct TaskList {
    std::vector<uint32_t> words;        // grouped task words
    std::vector<int32_t> begin;         // [n_groups][2]: first and one past the last task of the group
    std::vector<double> weights;        // 2 * kTaskBins doubles per task, same order as `words`
    std::vector<QSpec> qspec;           // per filter: its partial-sum rows
    int n_dst = 0;                      // partial-sum rows in use (including the dump row of unpaired filters)
    bool overflow = false;              // the rows do not fit into the exchange area
};

static void build_tasks(const std::vector<double>& bank, int n_filt, int n_bins, int radix_r, int n_groups, TaskList& tl)
{
    struct Run { int fa, fb, seg_lo, seg_hi; std::vector<int> k0; int da = 0, db = 0; };
    const int row_bins = pair_row_bins(radix_r);
    auto W = [&](int f, int k) { return (f >= 0 && k < n_bins) ? bank[(size_t)f * n_bins + k] : 0.0; };
    // 1. segments
    std::vector<int> lo(n_filt, n_bins), hi(n_filt, -1);
    std::vector<int> cuts;
    for (int f = 0; f < n_filt; ++f) {
        for (int k = 0; k < n_bins; ++k)
            if (W(f, k) != 0.0) { lo[f] = std::min(lo[f], k); hi[f] = std::max(hi[f], k); }
        if (hi[f] >= 0) { cuts.push_back(lo[f]); cuts.push_back(hi[f] + 1); }
    }
    std::sort(cuts.begin(), cuts.end());
    cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
    // 2. runs (one per segment and filter pair), tasks of kTaskBins bins
    std::vector<Run> whole;
    int total = 0;
    for (size_t s = 0; s + 1 < cuts.size(); ++s) {
        const int a = cuts[s], b = cuts[s + 1];
        std::vector<int> act;
        for (int f = 0; f < n_filt; ++f)
            if (hi[f] >= 0 && lo[f] <= a && b <= hi[f] + 1) act.push_back(f);
        for (size_t i = 0; i < act.size(); i += 2) {
            Run r;
            r.fa = act[i];
            r.fb = (i + 1 < act.size()) ? act[i + 1] : -1;
            r.seg_lo = a;
            r.seg_hi = b;
            for (int k0 = a & ~1; k0 < b; k0 += kTaskBins) r.k0.push_back(std::min(k0, row_bins - kTaskBins));
            total += (int)r.k0.size();
            whole.push_back(r);
        }
    }
    // 3. split long runs, spread over the groups
    const int cap = std::max(1, (total + n_groups - 1) / n_groups);
    std::vector<Run> runs;
    for (const Run& r : whole) {
        const int n = (int)r.k0.size();
        const int pieces = (n + cap - 1) / cap;
        int done = 0;
        for (int pc = 0; pc < pieces; ++pc) {
            const int len = (n - done + (pieces - pc) - 1) / (pieces - pc);
            Run x = r;
            x.k0.assign(r.k0.begin() + done, r.k0.begin() + done + len);
            // a piece owns the bins from its first task up to the next piece's first task
            if (pc > 0) x.seg_lo = std::max(r.seg_lo, r.k0[done]);
            if (pc + 1 < pieces) x.seg_hi = std::min(r.seg_hi, r.k0[done + len]);
            runs.push_back(x);
            done += len;
        }
    }
    // 4. partial-sum rows: a filter's rows are consecutive inside one warp region's tail (the log phase walks them with
    //    a fixed stride); the rows are addressed by their 64-byte-unit offset in the team's exchange area
    tl.qspec.assign(n_filt, QSpec{0, 0});
    int region = 0, row = 0;
    tl.n_dst = 0;
    tl.overflow = false;
    auto alloc = [&](int count) {
        if (row + count > drows(radix_r)) { ++region; row = 0; }
        if (region >= kTeamWarps || count > drows(radix_r)) { tl.overflow = true; region = 0; row = 0; }
        const int unit = partial_row_unit(radix_r, region, row);
        row += count;
        tl.n_dst += count;
        return unit;
    };
    const int unit_step = drow_floats(radix_r) / 16;
    for (int f = 0; f < n_filt; ++f) {
        int count = 0;
        for (const Run& r : runs) count += (r.fa == f) + (r.fb == f);
        int unit = alloc(count);
        tl.qspec[f].unit0 = unit;
        tl.qspec[f].count = count;
        for (Run& r : runs) {
            if (r.fa == f) { r.da = unit; unit += unit_step; }
            if (r.fb == f) { r.db = unit; unit += unit_step; }
        }
    }
    const int dump = alloc(1);
    for (Run& r : runs)
        if (r.fb < 0) r.db = dump;
    std::vector<size_t> order(runs.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return runs[a].k0.size() > runs[b].k0.size(); });
    std::vector<std::vector<size_t>> per_group(n_groups);
    std::vector<int> load(n_groups, 0);
    for (size_t idx : order) {
        int best = 0;
        for (int g = 1; g < n_groups; ++g)
            if (load[g] < load[best]) best = g;
        per_group[best].push_back(idx);
        load[best] += (int)runs[idx].k0.size();
    }
    tl.words.clear();
    tl.weights.clear();
    tl.begin.assign(2 * n_groups, 0);
    for (int g = 0; g < n_groups; ++g) {
        tl.begin[2 * g] = (int32_t)tl.words.size();
        for (size_t ri : per_group[g]) {
            const Run& r = runs[ri];
            int covered_to = r.seg_lo;                  // bins below this one already have their weight in a task
            for (size_t t = 0; t < r.k0.size(); ++t) {
                uint32_t word = (uint32_t)(r.k0[t] / 2);
                if (t + 1 == r.k0.size()) word |= 0x80000000u | ((uint32_t)r.da << 8) | ((uint32_t)r.db << 19);
                tl.words.push_back(word);
                for (int i = 0; i < kTaskBins; ++i) {
                    const int k = r.k0[t] + i;
                    const bool mine = k >= covered_to && k < r.seg_hi;
                    tl.weights.push_back(mine ? W(r.fa, k) : 0.0);
                    tl.weights.push_back(mine ? W(r.fb, k) : 0.0);
                }
                covered_to = std::max(covered_to, std::min(r.seg_hi, r.k0[t] + kTaskBins));
            }
        }
        tl.begin[2 * g + 1] = (int32_t)tl.words.size();
    }
}

// The task list evaluated on the host in the order the kernel's bank phase walks it (test hook for the decomposition;
// pair_row holds (A[k], B[k]) interleaved like the kernel's shared-memory row).
static void apply_tasks(const TaskList& tl, int radix_r, int n_groups, const std::vector<double>& pair_row,
                        std::vector<double>& sums_a, std::vector<double>& sums_b)
{
    std::vector<double> pa(2048, 0.0), pb(2048, 0.0);      // indexed by 64-byte unit (11-bit field)
    for (int g = 0; g < n_groups; ++g) {
        double aa = 0, ab = 0, ba = 0, bb = 0;          // (filter a | b) x (frame A | B)
        for (int t = tl.begin[2 * g]; t < tl.begin[2 * g + 1]; ++t) {
            const uint32_t w = tl.words[t];
            const int off = 4 * (int)(w & 0xffu);
            for (int i = 0; i < kTaskBins; ++i) {
                const double wa = tl.weights[(size_t)t * 2 * kTaskBins + 2 * i];
                const double wb = tl.weights[(size_t)t * 2 * kTaskBins + 2 * i + 1];
                aa += wa * pair_row[off + 2 * i];
                ab += wa * pair_row[off + 2 * i + 1];
                ba += wb * pair_row[off + 2 * i];
                bb += wb * pair_row[off + 2 * i + 1];
            }
            if (w & 0x80000000u) {
                const int da = (w >> 8) & 0x7ff, db = (w >> 19) & 0x7ff;
                pa[da] = aa; pb[da] = ab;
                pa[db] = ba; pb[db] = bb;
                aa = ab = ba = bb = 0;
            }
        }
    }
    sums_a.assign(tl.qspec.size(), 0.0);
    sums_b.assign(tl.qspec.size(), 0.0);
    const int unit_step = drow_floats(radix_r) / 16;
    for (size_t f = 0; f < tl.qspec.size(); ++f)
        for (int j = 0; j < tl.qspec[f].count; ++j) {
            sums_a[f] += pa[tl.qspec[f].unit0 + j * unit_step];
            sums_b[f] += pb[tl.qspec[f].unit0 + j * unit_step];
        }
}

// Hacker's Delight unsigned division by an invariant (round-up method); magic == 0 marks a power of two.
static void magic_div(uint32_t d, uint32_t& magic, uint32_t& shift)
{
    if ((d & (d - 1)) == 0) {
        magic = 0;
        shift = 0;
        while ((1u << shift) < d) ++shift;
        return;
    }
    uint32_t s = 0;
    while ((1ull << s) < d) ++s;                                        // 2^(s-1) < d < 2^s
    magic = (uint32_t)((((1ull << s) - d) << 32) / d + 1);
    shift = s - 1;
}

static void free_plan_tables(scf_plan* p)
{
    cudaFree(p->d_win); cudaFree(p->d_tab_i16); cudaFree(p->d_tab_f32); cudaFree(p->d_tile_ctr);
    if (p->ws.d_in) cudaFree(p->ws.d_in);
    if (p->ws.d_out) cudaFree(p->ws.d_out);
    if (p->ws.d_len) cudaFree(p->ws.d_len);
    if (p->ws.st) cudaStreamDestroy(p->ws.st);
    if (p->ws.st2) cudaStreamDestroy(p->ws.st2);
    if (p->ws.ev) cudaEventDestroy(p->ws.ev);
    p->ws.small.release();
    for (auto& sg : p->ws.stage) {
        if (sg.in) cudaFreeHost(sg.in);
        if (sg.out) cudaFreeHost(sg.out);
        if (sg.done) cudaEventDestroy(sg.done);
    }
    for (auto& sl : p->ws.slot) {
        if (sl.d_in) cudaFree(sl.d_in);
        if (sl.d_out) cudaFree(sl.d_out);
        if (sl.d_len) cudaFree(sl.d_len);
        if (sl.st) cudaStreamDestroy(sl.st);
    }
}

static int plan_create(const scf_config* cfg, scf_plan** out)
{
    if (!out) return fail(SCF_ERR_INVALID, "plan_out is NULL");
    *out = nullptr;
    int rc = check_config(cfg);
    if (rc) return rc;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(SCF_ERR_NO_DEVICE, "no CUDA device: libscfeat has no CPU fallback");
    }
    int dev = cfg->device;
    if (dev < 0) SCF_CUDA(cudaGetDevice(&dev));
    if (dev >= n_dev) return fail(SCF_ERR_INVALID, "device ordinal out of range");
    cudaDeviceProp prop;
    SCF_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        return fail(SCF_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100-class; libscfeat is built for sm_100a only");
    DeviceGuard guard(dev);
    if (!guard.ok) return fail(SCF_ERR_CUDA, "cudaSetDevice failed");

    scf_plan* p = new (std::nothrow) scf_plan();
    if (!p) return fail(SCF_ERR_ALLOC, "out of host memory");
    p->cfg = *cfg;
    p->cfg.device = dev;
    p->cfg.custom_bank = nullptr;
    p->device = dev;
    p->num_sms = prop.multiProcessorCount;
    p->radix_r = cfg->n_fft / 32;
    p->n_bins = cfg->n_fft / 2 + 1;
    p->out_cols = scf_out_cols(cfg);
    p->base_cols = p->out_cols / (cfg->delta == SCF_DELTA_NONE ? 1 : cfg->delta == SCF_DELTA_CENTRAL2 ? 3 : 2);
    const double pcm = (cfg->pcm_scale > 0.f) ? (double)cfg->pcm_scale : 1.0 / 32768.0;
    // the FFT stage leaves 2*X[k]; power_spec divides by n_fft (bark_feature.py:89)
    const double ps_f32 = 1.0 / (4.0 * cfg->n_fft);
    const double ps_i16 = pcm * pcm * ps_f32;
    p->power_scale_f32 = (float)ps_f32;
    p->power_scale_i16 = (float)ps_i16;

    // analysis window over the (uncropped) window length, inference/tflite/mfcc.h:404-407
    // (pre-emphasis alone gets a table of ones: the fused fast loader always multiplies)
    if (cfg->window_fn != SCF_WIN_RECT || cfg->preemph_alpha != 0.f) {
        const int w_eff = std::min(cfg->window, cfg->n_fft);
        std::vector<float> win(w_eff, 1.f);
        for (int j = 0; j < w_eff && cfg->window_fn != SCF_WIN_RECT; ++j) {
            const double c = cos(2.0 * M_PI * j / (double)(cfg->window - 1));
            win[j] = (float)(cfg->window_fn == SCF_WIN_HAMMING ? 0.54 - 0.46 * c : 0.5 - 0.5 * c);
        }
        rc = upload(&p->d_win, win);
        if (rc) { free_plan_tables(p); delete p; return rc; }
    }
    {
        cudaError_t e = cudaMalloc((void**)&p->d_tile_ctr, (size_t)kTileCtrRing * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMemset(p->d_tile_ctr, 0, (size_t)kTileCtrRing * sizeof(uint32_t));
        if (e != cudaSuccess) {
            free_plan_tables(p);
            delete p;
            return fail(SCF_ERR_ALLOC, std::string("tile counters: ") + cudaGetErrorString(e));
        }
    }
    // ---- the table blob -------------------------------------------------------------------------------
    TaskList tl;
    std::vector<float> dct_t;
    if (cfg->output != SCF_OUT_POWER) {
        std::vector<double> bank;
        build_bank(cfg, bank);
        const bool cep = cfg->output == SCF_OUT_CEPSTRUM;
        // (the frame energy -- c0 of the cepstrum -- is summed in the FFT stage, not as a bank row)
        build_tasks(bank, cfg->n_filt, p->n_bins, p->radix_r, bank_groups(p->radix_r), tl);
        if (tl.overflow) {
            free_plan_tables(p);
            delete p;
            return fail(SCF_ERR_INVALID, "filterbank needs more partial sums than the kernel's shared memory holds");
        }
        p->n_tasks = (int)tl.words.size();
        p->n_q = cfg->n_filt + (cep ? 1 : 0);
        p->n_dst = tl.n_dst;
        p->n_filt4 = (cfg->n_filt + 3) & ~3;
        if (cep) {
            p->n_out = std::min(cfg->n_filt, cfg->n_coeffs);
            std::vector<double> d;
            build_dct(cfg->n_filt, p->n_out, d);
            dct_t.assign((size_t)p->n_out * p->n_filt4, 0.f);             // [c][m], zero padded
            for (int c = 0; c < p->n_out; ++c)
                for (int m = 0; m < cfg->n_filt; ++m) dct_t[(size_t)c * p->n_filt4 + m] = (float)d[(size_t)m * p->n_out + c];
        }
    }
    auto align16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t sz_tw = 4 * 32 * sizeof(float4) + 32 * sizeof(float4);    // W^(k1 n2), n2 < 8, then (W^(8 k1), W^(16 k1))
    // order: work list, bank weights -- always staged in shared memory -- then the pass-2 twiddles and the DCT matrix
    // (staged too by the one-CTA-per-SM kernels; read through L1 by the 3-CTA variant, which needs the space for warps)
    p->off_tasks = 0;
    p->off_tbeg = (int)align16(p->off_tasks + (size_t)p->n_tasks * 4);
    p->off_qspec = (int)align16(p->off_tbeg + tl.begin.size() * 4);
    p->off_wts = (int)align16(p->off_qspec + tl.qspec.size() * sizeof(QSpec));
    p->off_tw = (int)align16(p->off_wts + (size_t)p->n_tasks * 2 * kTaskBins * 4);
    p->table_small_bytes = p->off_tw;
    p->off_dct = (int)align16(p->off_tw + sz_tw);
    p->table_bytes = (int)align16(p->off_dct + dct_t.size() * 4);
    for (int variant = 0; variant < 2; ++variant) {
        const double ps = variant == 0 ? ps_i16 : ps_f32;
        std::vector<unsigned char> blob(p->table_bytes, 0);
        // pass-2 twiddles: lane L handles column k1 = L % R; W_N^(k1*n2), n2 = 0..31, as float4 pairs
        {
            const int R = p->radix_r, N = cfg->n_fft;
            float4* tw = reinterpret_cast<float4*>(blob.data() + p->off_tw);
            for (int jj = 0; jj < 4; ++jj)
                for (int lane = 0; lane < 32; ++lane) {
                    const int k1 = lane % R;
                    const double a0 = -2.0 * M_PI * (double)((k1 * (2 * jj)) % N) / N;
                    const double a1 = -2.0 * M_PI * (double)((k1 * (2 * jj + 1)) % N) / N;
                    tw[jj * 32 + lane] = make_float4((float)cos(a0), (float)sin(a0), (float)cos(a1), (float)sin(a1));
                }
            float4* wq = tw + 4 * 32;
            for (int lane = 0; lane < 32; ++lane) {
                const double a8 = -2.0 * M_PI * (double)(((lane % R) * 8) % N) / N;
                const double a16 = -2.0 * M_PI * (double)(((lane % R) * 16) % N) / N;
                wq[lane] = make_float4((float)cos(a8), (float)sin(a8), (float)cos(a16), (float)sin(a16));
            }
        }
        float* w = reinterpret_cast<float*>(blob.data() + p->off_wts);
        for (size_t i = 0; i < tl.weights.size(); ++i) w[i] = (float)(tl.weights[i] * ps);
        if (!dct_t.empty()) memcpy(blob.data() + p->off_dct, dct_t.data(), dct_t.size() * 4);
        if (!tl.words.empty()) memcpy(blob.data() + p->off_tasks, tl.words.data(), tl.words.size() * 4);
        if (!tl.begin.empty()) memcpy(blob.data() + p->off_tbeg, tl.begin.data(), tl.begin.size() * 4);
        if (!tl.qspec.empty()) memcpy(blob.data() + p->off_qspec, tl.qspec.data(), tl.qspec.size() * sizeof(QSpec));
        rc = upload(variant == 0 ? &p->d_tab_i16 : &p->d_tab_f32, blob);
        if (rc) { free_plan_tables(p); delete p; return rc; }
    }
    *out = p;
    return SCF_OK;
}

static int fill_params(const scf_plan* plan, bool is_f32, const void* d_in, int64_t n_clips, int64_t clip_stride,
                       int32_t clip_len, const int32_t* d_lengths, int32_t pad, float* d_out, KParams& kp,
                       int64_t& n_tiles, bool& fast)
{
    const scf_config& c = plan->cfg;
    if (n_clips < 0 || clip_len < 0) return fail(SCF_ERR_INVALID, "negative size");
    if (n_clips > 0 && (!d_in || clip_stride < clip_len)) return fail(SCF_ERR_INVALID, "bad input pointer / stride");
    if (pad != SCF_PAD_FRONT_ZERO && pad != SCF_PAD_NONE) return fail(SCF_ERR_INVALID, "bad pad kind");
    if (n_clips > 0x7fffffffLL) return fail(SCF_ERR_INVALID, "too many clips");
    memset(&kp, 0, sizeof(kp));
    kp.in = d_in;
    kp.clip_stride = clip_stride;
    kp.clip_len = clip_len;
    kp.lengths = d_lengths;
    kp.pad_mode = pad;
    kp.frames_per_clip = (int32_t)scf_num_frames(clip_len, c.window, c.hop);
    kp.pairs_per_clip = (kp.frames_per_clip + 1) / 2;
    kp.n_pairs = n_clips * (int64_t)kp.pairs_per_clip;
    kp.window = c.window;
    kp.hop = c.hop;
    kp.w_eff = std::min(c.window, c.n_fft);
    kp.preemph = c.preemph_alpha;
    kp.win = plan->d_win;
    kp.out = d_out;
    kp.n_peers = 0;
    kp.peer_row0 = 0;
    kp.out_cols = plan->base_cols;
    kp.out_pitch = plan->out_cols;
    kp.out_kind = c.output;
    kp.power_scale = is_f32 ? plan->power_scale_f32 : plan->power_scale_i16;
    {   // smallest non-zero int16 frame: one sample of +-1 -> energy 513/1024 * pcm_scale^2
        const double pcm = (c.pcm_scale > 0.f) ? (double)c.pcm_scale : 1.0 / 32768.0;
        kp.zero_energy = is_f32 ? 0.f : (float)(0.25 * pcm * pcm);
    }
    kp.tables = is_f32 ? plan->d_tab_f32 : plan->d_tab_i16;
    kp.table_bytes = plan->table_bytes;
    kp.table_small_bytes = plan->table_small_bytes;
    kp.off_wts = plan->off_wts;
    kp.off_tw = plan->off_tw;
    kp.off_dct = plan->off_dct;
    kp.off_tasks = plan->off_tasks;
    kp.off_tbeg = plan->off_tbeg;
    kp.off_qspec = plan->off_qspec;
    kp.n_q = plan->n_q;
    kp.n_filt = c.n_filt;
    kp.n_filt4 = plan->n_filt4;
    kp.n_out = plan->n_out;
    kp.frames_odd = kp.frames_per_clip & 1;
    magic_div((uint32_t)std::max(1, kp.pairs_per_clip), kp.ppc_magic, kp.ppc_shift);
    const int ppt = pairs_per_tile(plan->radix_r);
    n_tiles = (kp.n_pairs + ppt - 1) / ppt;
    // the fast kernels also take short clips that are zero-padded in front (common/data_utils.py:77-80)
    // ... and, with a second load per sample, pre-emphasis and an analysis window (inference/tflite/mfcc.h:394-410)
    fast = (c.window == c.n_fft) && (c.hop * 2 == c.n_fft) && (d_lengths == nullptr || pad == SCF_PAD_FRONT_ZERO);
    kp.fast_pre = (fast && (c.preemph_alpha != 0.f || c.window_fn != SCF_WIN_RECT)) ? 1 : 0;
    return SCF_OK;
}

static inline int c_window(const scf_plan* plan) { return plan->cfg.window; }
static inline int c_hop(const scf_plan* plan) { return plan->cfg.hop; }

static int extract_device(const scf_plan* plan, bool is_f32, const void* d_in, int64_t n_clips, int64_t clip_stride,
                          int32_t clip_len, const int32_t* d_lengths, int32_t pad, float* d_out,
                          float* const* peers, int world, int rank, void* cuda_stream, int64_t clips_per_rank = 0,
                          const StreamStep* stream_step = nullptr, bool world_is_multicast = false)
{
    if (!plan) return fail(SCF_ERR_INVALID, "plan is NULL");
    KParams kp;
    int64_t n_tiles;
    bool fast;
    int rc = fill_params(plan, is_f32, d_in, n_clips, clip_stride, clip_len, d_lengths, pad, d_out, kp, n_tiles, fast);
    if (rc) return rc;
    if (stream_step) {
        kp.stream_on = 1;
        kp.stream = *stream_step;
        kp.out_pitch = plan->base_cols;       // `out` is the stream's own ring: base columns only
        fast = false;             // the generic loader reads concat(carry, chunk)
        kp.fast_pre = 0;
    }
    kp.fast_path = fast ? 1 : 0;
    if (peers) {
        if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return fail(SCF_ERR_INVALID, "bad world/rank");
        if (plan->cfg.output == SCF_OUT_POWER) return fail(SCF_ERR_INVALID, "fused gather does not support power output");
        if (plan->cfg.delta != SCF_DELTA_NONE) return fail(SCF_ERR_INVALID, "fused gather does not support delta features");
        kp.out = nullptr;
        const bool multicast = world_is_multicast;
        kp.n_peers = multicast ? 1 : world;
        kp.peer_multicast = multicast ? 1 : 0;
        for (int r = 0; r < kp.n_peers; ++r) {
            if (!peers[r]) return fail(SCF_ERR_INVALID, "peer pointer is NULL");
            if (reinterpret_cast<uintptr_t>(peers[r]) & 15) return fail(SCF_ERR_INVALID, "peer buffers must be 16-byte aligned");
            kp.peer_out[r] = peers[r];
        }
        if (clips_per_rank < n_clips) return fail(SCF_ERR_INVALID, "clips_per_rank must be >= n_local");
        kp.peer_row0 = (int64_t)rank * clips_per_rank * kp.frames_per_clip;
    } else if (n_clips > 0 && kp.frames_per_clip > 0 && !d_out) {
        return fail(SCF_ERR_INVALID, "output pointer is NULL");
    }
    if (n_tiles == 0) return SCF_OK;
    DeviceGuard guard(plan->device);
    if (!guard.ok) return fail(SCF_ERR_CUDA, "cudaSetDevice failed");
    const size_t smem = extract_smem_bytes(plan->radix_r, kp);
    if (smem > extract_smem_limit(plan->radix_r, kp)) return fail(SCF_ERR_INVALID, "configuration needs too much shared memory");
    // the kernel indexes pairs with 32 bits: very large jobs go out as several launches
    const int ppt = pairs_per_tile(plan->radix_r);
    const int64_t max_clips = std::max<int64_t>(1, (0x7fffffffLL - ppt) / std::max(1, kp.pairs_per_clip));
    const size_t esz = is_f32 ? 4 : 2;
    for (int64_t c0 = 0; c0 < n_clips; c0 += max_clips) {
        const int64_t nc = std::min(max_clips, n_clips - c0);
        KParams k = kp;
        k.in = static_cast<const unsigned char*>(d_in) + (size_t)c0 * clip_stride * esz;
        if (d_lengths) k.lengths = d_lengths + c0;
        const int64_t row0 = c0 * kp.frames_per_clip;
        if (k.out) k.out = d_out + row0 * kp.out_pitch;
        k.peer_row0 = kp.peer_row0 + row0;
        k.n_pairs = nc * (int64_t)kp.pairs_per_clip;
        const int64_t tiles = (k.n_pairs + ppt - 1) / ppt;
        // (a team's first three tiles are fixed: smaller launches -- the 512-clip batch has 2.2 tiles per team -- keep the
        //  round robin and skip the draws)
        if (fast && !stream_step && plan->d_tile_ctr != nullptr && plan->cfg.output != SCF_OUT_POWER &&
            tiles > 3 * 3 * (int64_t)plan->num_sms)
            k.tile_ctr = plan->d_tile_ctr + plan->ctr_next.fetch_add(1, std::memory_order_relaxed) % kTileCtrRing;
        SCF_CUDA(launch_extract(plan->radix_r, is_f32, fast, k, tiles, plan->num_sms, (cudaStream_t)cuda_stream, smem));
    }
    // delta columns: a second, tiny pass over the finished rows (it needs the neighbouring frames of every row, which
    // other CTAs produced).  Launched without the programmatic attribute, so it starts after the extraction has
    // completed, and the next extraction after it.
    if (plan->cfg.delta != SCF_DELTA_NONE) {
        if (stream_step) {
            if (stream_step->ring_copy)
                SCF_CUDA(launch_delta(stream_step->ring_copy, nullptr, n_clips, stream_step->ring_rows, plan->base_cols,
                                      stream_step->copy_pitch, plan->cfg.delta, 0, 1, 1, SCF_PAD_FRONT_ZERO, (cudaStream_t)cuda_stream));
        } else if (d_out) {
            SCF_CUDA(launch_delta(d_out, d_lengths, n_clips, kp.frames_per_clip, plan->base_cols, plan->out_cols, plan->cfg.delta,
                                  clip_len, c_window(plan), c_hop(plan), pad, (cudaStream_t)cuda_stream));
        }
    }
    return SCF_OK;
}

template <typename T>
static int grow(T** ptr, size_t* have, size_t need)
{
    if (need <= *have) return SCF_OK;
    if (*ptr) SCF_CUDA(cudaFree(*ptr));
    *ptr = nullptr;
    *have = 0;
    SCF_CUDA(cudaMalloc((void**)ptr, need));
    *have = need;
    return SCF_OK;
}

static int extract_host(const scf_plan* plan, bool is_f32, const void* h_in, int64_t n_clips, int64_t clip_stride,
                        int32_t clip_len, const int32_t* h_lengths, int32_t pad, float* h_out)
{
    if (!plan) return fail(SCF_ERR_INVALID, "plan is NULL");
    if (n_clips < 0 || clip_len < 0) return fail(SCF_ERR_INVALID, "negative size");
    if (n_clips == 0) return SCF_OK;
    if (!h_in || clip_stride < clip_len) return fail(SCF_ERR_INVALID, "bad input pointer / stride");
    const int64_t fpc = scf_num_frames(clip_len, plan->cfg.window, plan->cfg.hop);
    if (fpc == 0) return SCF_OK;
    if (!h_out) return fail(SCF_ERR_INVALID, "output pointer is NULL");
    DeviceGuard guard(plan->device);
    if (!guard.ok) return fail(SCF_ERR_CUDA, "cudaSetDevice failed");
    Workspace& ws = const_cast<scf_plan*>(plan)->ws;
    std::lock_guard<std::mutex> lock(ws.mu);
    if (!ws.st) SCF_CUDA(cudaStreamCreateWithFlags(&ws.st, cudaStreamNonBlocking));
    if (!ws.st2) SCF_CUDA(cudaStreamCreateWithFlags(&ws.st2, cudaStreamNonBlocking));
    if (!ws.ev) SCF_CUDA(cudaEventCreateWithFlags(&ws.ev, cudaEventDisableTiming));
    const size_t esz = is_f32 ? 4 : 2;
    const size_t in_bytes = (size_t)((n_clips - 1) * clip_stride + clip_len) * esz;
    const size_t row_bytes = (size_t)fpc * plan->out_cols * sizeof(float);
    const size_t out_bytes = (size_t)n_clips * row_bytes;
    int rc;
    // small call: the kernel reads the samples from / writes the rows to mapped host memory -- one launch, one sync
    const size_t len_bytes = h_lengths ? (((size_t)n_clips * 4 + 15) & ~(size_t)15) : 0;
    const size_t in_al = (in_bytes + 15) & ~(size_t)15;
    if (in_al + len_bytes + out_bytes <= kSmallCallBytes && ws.small.ensure(kSmallCallBytes) == 0) {
        unsigned char* hs = ws.small.h;
        unsigned char* ds = ws.small.d;
        memcpy(hs, h_in, in_bytes);
        if (h_lengths) memcpy(hs + in_al, h_lengths, (size_t)n_clips * 4);
        float* h_rows = reinterpret_cast<float*>(hs + in_al + len_bytes);
        if (pad == SCF_PAD_NONE && h_lengths) memset(h_rows, 0, out_bytes);     // rows of short clips stay untouched
        rc = extract_device(plan, is_f32, ds, n_clips, clip_stride, clip_len,
                            h_lengths ? reinterpret_cast<const int32_t*>(ds + in_al) : nullptr, pad,
                            reinterpret_cast<float*>(ds + in_al + len_bytes), nullptr, 0, 0, ws.st);
        cudaError_t se = cudaStreamSynchronize(ws.st);
        if (rc) return rc;
        if (se != cudaSuccess) return fail(SCF_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(se));
        memcpy(h_out, h_rows, out_bytes);
        return SCF_OK;
    }
    if ((rc = grow(&ws.d_in, &ws.in_bytes, in_bytes)) || (rc = grow(&ws.d_out, &ws.out_bytes, out_bytes))) return rc;
    const int32_t* d_len = nullptr;
    if (h_lengths) {
        if ((rc = grow(&ws.d_len, &ws.len_bytes, (size_t)n_clips * 4))) return rc;
        SCF_CUDA(cudaMemcpyAsync(ws.d_len, h_lengths, (size_t)n_clips * 4, cudaMemcpyHostToDevice, ws.st));
        d_len = ws.d_len;
    }
    if (pad == SCF_PAD_NONE && h_lengths)      // rows of short clips stay untouched by the kernel: make them zero
        SCF_CUDA(cudaMemsetAsync(ws.d_out, 0, out_bytes, ws.st));
    SCF_CUDA(cudaEventRecord(ws.ev, ws.st));
    SCF_CUDA(cudaStreamWaitEvent(ws.st2, ws.ev, 0));
    const size_t clip_bytes = (size_t)clip_stride * esz;
    // Pageable caller memory: chunks of ~4 MB go through a ring of pinned slots -- a few host threads copy chunk i+1 into
    // its slot while chunk i is uploaded, transformed and downloaded; finished rows are copied out of the slot's pinned
    // output as the ring comes round.
    {
        cudaPointerAttributes at;
        const bool pageable = cudaPointerGetAttributes(&at, h_in) != cudaSuccess || at.type == cudaMemoryTypeUnregistered;
        cudaGetLastError();
        // (below ~8 MB the driver's own bounce copy is quicker than waking the copy threads chunk by chunk: 64 clips
        //  184 us against 247 us; 512 clips 0.92 ms against 1.32 ms, 4096 clips 4.6 ms against 15.2 ms)
        if (pageable && in_bytes >= (8u << 20)) {
            // ~16 chunks per call (256 KB .. 4 MB each) so that copies, uploads and kernels of neighbouring chunks overlap
            const size_t want = std::min<size_t>(kStageChunkBytes, std::max<size_t>(256u << 10, in_bytes / 16));
            const int64_t chunk_p = std::max<int64_t>(1, (int64_t)(want / std::max<size_t>(clip_bytes, 1)));
            const size_t in_cap = (size_t)((chunk_p - 1) * clip_stride + clip_len) * esz, out_cap = (size_t)chunk_p * row_bytes;
            for (auto& sg : ws.stage) {
                if (sg.in_bytes < in_cap) {
                    if (sg.in) cudaFreeHost(sg.in);
                    sg.in = nullptr; sg.in_bytes = 0;
                    SCF_CUDA(cudaHostAlloc((void**)&sg.in, in_cap, cudaHostAllocDefault));
                    sg.in_bytes = in_cap;
                }
                if (sg.out_bytes < out_cap) {
                    if (sg.out) cudaFreeHost(sg.out);
                    sg.out = nullptr; sg.out_bytes = 0;
                    SCF_CUDA(cudaHostAlloc((void**)&sg.out, out_cap, cudaHostAllocDefault));
                    sg.out_bytes = out_cap;
                }
                if (!sg.done) SCF_CUDA(cudaEventCreateWithFlags(&sg.done, cudaEventDisableTiming));
            }
            CopyPool& pool = CopyPool::get();
            const int64_t n_chunks = (n_clips + chunk_p - 1) / chunk_p;
            auto drain = [&](int64_t c) -> int {          // rows of chunk c: pinned slot -> caller's array
                StageSlot& sg = ws.stage[c % kStageSlots];
                SCF_CUDA(cudaEventSynchronize(sg.done));
                const int64_t c0 = c * chunk_p, nc = std::min(chunk_p, n_clips - c0);
                pool.copy(reinterpret_cast<unsigned char*>(h_out) + (size_t)c0 * row_bytes, sg.out, (size_t)nc * row_bytes);
                return SCF_OK;
            };
            rc = SCF_OK;
            int64_t drained = 0;
            for (int64_t c = 0; c < n_chunks && rc == SCF_OK; ++c) {
                if (c >= kStageSlots) { rc = drain(drained++); if (rc) break; }
                StageSlot& sg = ws.stage[c % kStageSlots];
                const int64_t c0 = c * chunk_p, nc = std::min(chunk_p, n_clips - c0);
                const size_t bytes = (size_t)((nc - 1) * clip_stride + clip_len) * esz;
                pool.copy(sg.in, static_cast<const unsigned char*>(h_in) + (size_t)c0 * clip_bytes, bytes);
                cudaStream_t st = (c & 1) ? ws.st2 : ws.st;
                unsigned char* d_src = static_cast<unsigned char*>(ws.d_in) + (size_t)c0 * clip_bytes;
                float* d_dst = ws.d_out + (size_t)c0 * fpc * plan->out_cols;
                cudaError_t e = cudaMemcpyAsync(d_src, sg.in, bytes, cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess) {
                    rc = extract_device(plan, is_f32, d_src, nc, clip_stride, clip_len, d_len ? d_len + c0 : nullptr, pad, d_dst,
                                        nullptr, 0, 0, st);
                    if (rc == SCF_OK) e = cudaMemcpyAsync(sg.out, d_dst, (size_t)nc * row_bytes, cudaMemcpyDeviceToHost, st);
                    if (rc == SCF_OK && e == cudaSuccess) e = cudaEventRecord(sg.done, st);
                }
                if (rc == SCF_OK && e != cudaSuccess) rc = fail(SCF_ERR_CUDA, std::string("staged copy: ") + cudaGetErrorString(e));
            }
            while (rc == SCF_OK && drained < n_chunks) rc = drain(drained++);
            cudaStreamSynchronize(ws.st);
            cudaStreamSynchronize(ws.st2);
            return rc;
        }
    }
    // Pinned caller memory: chunks of >= 2 MB of input alternate between two streams so that the H2D copy of chunk i+1
    // overlaps the kernel and the D2H copy of chunk i (separate copy engines); every chunk owns a disjoint slice of the
    // staging buffers, so there is nothing to recycle.
    int64_t chunk = std::max<int64_t>(1, (int64_t)((2u << 20) / std::max<size_t>(clip_bytes, 1)));
    if (n_clips < 2 * chunk) chunk = n_clips;
    int which = 0;
    for (int64_t c0 = 0; c0 < n_clips; c0 += chunk, which ^= 1) {
        const int64_t nc = std::min(chunk, n_clips - c0);
        cudaStream_t st = which ? ws.st2 : ws.st;
        const unsigned char* h_src = static_cast<const unsigned char*>(h_in) + (size_t)c0 * clip_bytes;
        unsigned char* d_src = static_cast<unsigned char*>(ws.d_in) + (size_t)c0 * clip_bytes;
        const size_t bytes = (size_t)((nc - 1) * clip_stride + clip_len) * esz;
        SCF_CUDA(cudaMemcpyAsync(d_src, h_src, bytes, cudaMemcpyHostToDevice, st));
        float* d_dst = ws.d_out + (size_t)c0 * fpc * plan->out_cols;
        rc = extract_device(plan, is_f32, d_src, nc, clip_stride, clip_len, d_len ? d_len + c0 : nullptr, pad, d_dst,
                            nullptr, 0, 0, st);
        if (rc) { cudaStreamSynchronize(ws.st); cudaStreamSynchronize(ws.st2); return rc; }
        SCF_CUDA(cudaMemcpyAsync(reinterpret_cast<unsigned char*>(h_out) + (size_t)c0 * row_bytes, d_dst,
                                 (size_t)nc * row_bytes, cudaMemcpyDeviceToHost, st));
    }
    SCF_CUDA(cudaStreamSynchronize(ws.st));
    SCF_CUDA(cudaStreamSynchronize(ws.st2));
    return SCF_OK;
}

static int extract_host_async(const scf_plan* plan, const int16_t* h_in, int64_t n_clips, int64_t clip_stride,
                              int32_t clip_len, const int32_t* h_lengths, int32_t pad, float* h_out,
                              cudaStream_t* used = nullptr)
{
    if (!plan) return fail(SCF_ERR_INVALID, "plan is NULL");
    if (n_clips < 0 || clip_len < 0) return fail(SCF_ERR_INVALID, "negative size");
    if (n_clips == 0) return SCF_OK;
    if (!h_in || clip_stride < clip_len) return fail(SCF_ERR_INVALID, "bad input pointer / stride");
    const int64_t fpc = scf_num_frames(clip_len, plan->cfg.window, plan->cfg.hop);
    if (fpc == 0) return SCF_OK;
    if (!h_out) return fail(SCF_ERR_INVALID, "output pointer is NULL");
    DeviceGuard guard(plan->device);
    if (!guard.ok) return fail(SCF_ERR_CUDA, "cudaSetDevice failed");
    Workspace& ws = const_cast<scf_plan*>(plan)->ws;
    std::lock_guard<std::mutex> lock(ws.mu);
    Workspace::Slot& sl = ws.slot[ws.next_slot++ & 1];
    if (!sl.st) SCF_CUDA(cudaStreamCreateWithFlags(&sl.st, cudaStreamNonBlocking));
    if (used) *used = sl.st;
    const size_t in_bytes = (size_t)((n_clips - 1) * clip_stride + clip_len) * 2;
    const size_t out_bytes = (size_t)n_clips * fpc * plan->out_cols * sizeof(float);
    if (in_bytes > sl.in_bytes || out_bytes > sl.out_bytes || (h_lengths && (size_t)n_clips * 4 > sl.len_bytes))
        SCF_CUDA(cudaStreamSynchronize(sl.st));       // the slot's buffers are about to be reallocated
    int rc;
    if ((rc = grow(&sl.d_in, &sl.in_bytes, in_bytes)) || (rc = grow(&sl.d_out, &sl.out_bytes, out_bytes))) return rc;
    const int32_t* d_len = nullptr;
    if (h_lengths) {
        if ((rc = grow(&sl.d_len, &sl.len_bytes, (size_t)n_clips * 4))) return rc;
        SCF_CUDA(cudaMemcpyAsync(sl.d_len, h_lengths, (size_t)n_clips * 4, cudaMemcpyHostToDevice, sl.st));
        d_len = sl.d_len;
    }
    // stream order on the slot's stream also protects its staging buffers from the call two steps later
    SCF_CUDA(cudaMemcpyAsync(sl.d_in, h_in, in_bytes, cudaMemcpyHostToDevice, sl.st));
    if (pad == SCF_PAD_NONE && h_lengths) SCF_CUDA(cudaMemsetAsync(sl.d_out, 0, out_bytes, sl.st));
    rc = extract_device(plan, false, sl.d_in, n_clips, clip_stride, clip_len, d_len, pad, sl.d_out, nullptr, 0, 0, sl.st);
    if (rc) return rc;
    SCF_CUDA(cudaMemcpyAsync(h_out, sl.d_out, out_bytes, cudaMemcpyDeviceToHost, sl.st));
    return SCF_OK;
}

// for scfeat_ingest.cu
int extract_host_async_on(const scf_plan* plan, const int16_t* h_in, int64_t n_clips, int64_t clip_stride, int32_t clip_len,
                          const int32_t* h_lengths, int32_t pad, float* h_out, cudaStream_t* used)
{
    return extract_host_async(plan, h_in, n_clips, clip_stride, clip_len, h_lengths, pad, h_out, used);
}
int plan_sample_rate(const scf_plan* plan) { return plan->cfg.sample_rate; }
int plan_device(const scf_plan* plan) { return plan->device; }
int64_t plan_row_floats(const scf_plan* plan, int32_t clip_len)
{
    return scf_num_frames(clip_len, plan->cfg.window, plan->cfg.hop) * (int64_t)plan->out_cols;
}

static int host_sync(const scf_plan* plan)
{
    if (!plan) return fail(SCF_ERR_INVALID, "plan is NULL");
    DeviceGuard guard(plan->device);
    Workspace& ws = const_cast<scf_plan*>(plan)->ws;
    std::lock_guard<std::mutex> lock(ws.mu);
    for (auto& sl : ws.slot)
        if (sl.st) SCF_CUDA(cudaStreamSynchronize(sl.st));
    return SCF_OK;
}

// ---------------------------------------------------------------------------------------------
// DLPack hand-off
// ---------------------------------------------------------------------------------------------
struct DlOwner {
    DLManagedTensor mt;
    int64_t shape[4];
    int device;
    void (*release)(void*) = nullptr;      // wrapped buffers: tells the owner instead of freeing
    void* release_ctx = nullptr;
};

static void dl_deleter(DLManagedTensor* self)
{
    if (!self) return;
    DlOwner* o = static_cast<DlOwner*>(self->manager_ctx);
    if (o->release) {
        o->release(o->release_ctx);
    } else {
        // cudaFree synchronises the device: a consumer may still have work in flight on streams this library does not
        // know about when it drops the tensor, so the release is the one place where that is wanted
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(o->device);
        cudaFree(self->dl_tensor.data);
        if (prev >= 0) cudaSetDevice(prev);
    }
    delete o;
}

// float32 kDLCUDA tensor descriptor over `data`
static DlOwner* dl_make(void* data, int device, const int64_t* shape, int ndim)
{
    DlOwner* o = new (std::nothrow) DlOwner();
    if (!o) return nullptr;
    o->device = device;
    for (int i = 0; i < ndim; ++i) o->shape[i] = shape[i];
    o->mt.dl_tensor.data = data;
    o->mt.dl_tensor.device.device_type = kDLCUDA;
    o->mt.dl_tensor.device.device_id = device;
    o->mt.dl_tensor.ndim = ndim;
    o->mt.dl_tensor.dtype.code = kDLFloat;
    o->mt.dl_tensor.dtype.bits = 32;
    o->mt.dl_tensor.dtype.lanes = 1;
    o->mt.dl_tensor.shape = o->shape;
    o->mt.dl_tensor.strides = nullptr;
    o->mt.dl_tensor.byte_offset = 0;
    o->mt.manager_ctx = o;
    o->mt.deleter = dl_deleter;
    return o;
}

// Stream-ordered allocation for library-owned outputs: cudaMallocAsync does not synchronise the device (cudaMalloc does),
// so scf_extract_i16_dlpack stays an enqueue-only call.  The pool keeps its memory between calls.
static cudaError_t alloc_on_stream(void** ptr, size_t bytes, int device, cudaStream_t st)
{
    static std::once_flag once[16];
    std::call_once(once[device & 15], [device] {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    });
    return cudaMallocAsync(ptr, bytes, st);
}

}  // namespace scf

// =============================================================================================
// C ABI
// =============================================================================================
using namespace scf;

extern "C" {

int scf_config_default(scf_config* cfg)
{
    if (!cfg) return fail(SCF_ERR_INVALID, "cfg is NULL");
    memset(cfg, 0, sizeof(*cfg));
    cfg->sample_rate = 16000;      // configs/params.json
    cfg->window = 1024;            // int(16000 * 0.064 + 0.5)
    cfg->hop = 512;                // int(16000 * 0.032 + 0.5)
    cfg->n_fft = 1024;
    cfg->n_filt = 20;
    cfg->n_coeffs = 20;
    cfg->bank = SCF_BANK_MEL_SONOPY;
    cfg->bank_scale = SCF_SCALE_CONSTANT;
    cfg->output = SCF_OUT_CEPSTRUM;
    cfg->window_fn = SCF_WIN_RECT;
    cfg->preemph_alpha = 0.f;
    cfg->pcm_scale = 1.0f / 32768.0f;
    cfg->device = -1;
    return SCF_OK;
}

int64_t scf_num_frames(int64_t n_samples, int32_t window, int32_t hop)
{
    if (window < 1 || hop < 1 || n_samples < window) return 0;
    return (n_samples - window) / hop + 1;
}

int32_t scf_out_cols(const scf_config* cfg)
{
    if (!cfg) return 0;
    const int blocks = cfg->delta == SCF_DELTA_NONE ? 1 : cfg->delta == SCF_DELTA_CENTRAL2 ? 3 : 2;
    switch (cfg->output) {
        case SCF_OUT_POWER: return cfg->n_fft / 2 + 1;
        case SCF_OUT_LOG_BANK: return cfg->n_filt * blocks;
        default: return std::min(cfg->n_filt, cfg->n_coeffs) * blocks;
    }
}

int scf_build_bank(const scf_config* cfg, double* bank_out)
{
    scf_config c;
    if (!cfg || !bank_out) return fail(SCF_ERR_INVALID, "NULL argument");
    c = *cfg;
    if (c.output == SCF_OUT_POWER) c.output = SCF_OUT_LOG_BANK;
    int rc = check_config(&c);
    if (rc) return rc;
    std::vector<double> bank;
    build_bank(&c, bank);
    memcpy(bank_out, bank.data(), bank.size() * sizeof(double));
    return SCF_OK;
}

int scf_bank_apply_tasks(const scf_config* cfg, const double* power_a, const double* power_b, double* sums_a,
                         double* sums_b, int32_t* stats4)
{
    scf_config c;
    if (!cfg || !power_a || !power_b || !sums_a || !sums_b) return fail(SCF_ERR_INVALID, "NULL argument");
    c = *cfg;
    if (c.output == SCF_OUT_POWER) c.output = SCF_OUT_LOG_BANK;
    int rc = check_config(&c);
    if (rc) return rc;
    const int r = c.n_fft / 32, n_bins = c.n_fft / 2 + 1, n_groups = bank_groups(r);
    std::vector<double> bank;
    build_bank(&c, bank);
    TaskList tl;
    build_tasks(bank, c.n_filt, n_bins, r, n_groups, tl);
    // the kernel's pair row: (A[k], B[k]) interleaved, bins 0 .. n_fft/2 + 1, the last one a pad.  The pad and the
    // bins a task touches outside its segment must carry zero weight: poison them with NaN-free but huge values
    std::vector<double> row((size_t)pair_row_floats(r), 1e300);
    for (int k = 0; k < n_bins; ++k) { row[2 * k] = power_a[k]; row[2 * k + 1] = power_b[k]; }
    std::vector<double> sa, sb;
    apply_tasks(tl, r, n_groups, row, sa, sb);
    if (tl.overflow) return fail(SCF_ERR_INVALID, "filterbank needs more partial sums than the kernel's shared memory holds");
    memcpy(sums_a, sa.data(), sa.size() * sizeof(double));
    memcpy(sums_b, sb.data(), sb.size() * sizeof(double));
    if (stats4) {
        int worst = 0;
        for (int g = 0; g < n_groups; ++g) worst = std::max(worst, tl.begin[2 * g + 1] - tl.begin[2 * g]);
        stats4[0] = (int32_t)tl.words.size();
        stats4[1] = tl.n_dst;
        stats4[2] = n_groups;
        stats4[3] = worst;
    }
    return SCF_OK;
}

int scf_build_dct(int32_t n_filt, int32_t n_coeffs, double* dct_out)
{
    if (n_filt < 1 || n_coeffs < 1 || !dct_out) return fail(SCF_ERR_INVALID, "bad argument");
    std::vector<double> d;
    build_dct(n_filt, std::min(n_filt, n_coeffs), d);
    memcpy(dct_out, d.data(), d.size() * sizeof(double));
    return SCF_OK;
}

int scf_plan_create(const scf_config* cfg, scf_plan** plan_out) { return plan_create(cfg, plan_out); }

void scf_plan_destroy(scf_plan* plan)
{
    if (!plan) return;
    DeviceGuard guard(plan->device);
    free_plan_tables(plan);
    delete plan;
}

int scf_plan_config(const scf_plan* plan, scf_config* cfg_out)
{
    if (!plan || !cfg_out) return fail(SCF_ERR_INVALID, "NULL argument");
    *cfg_out = plan->cfg;
    return SCF_OK;
}

int scf_extract_i16(const scf_plan* plan, const int16_t* d_pcm, int64_t n_clips, int64_t clip_stride, int32_t clip_len,
                    const int32_t* d_lengths, int32_t pad, float* d_out, void* cuda_stream)
{
    return extract_device(plan, false, d_pcm, n_clips, clip_stride, clip_len, d_lengths, pad, d_out, nullptr, 0, 0,
                          cuda_stream);
}

int scf_extract_f32(const scf_plan* plan, const float* d_audio, int64_t n_clips, int64_t clip_stride, int32_t clip_len,
                    const int32_t* d_lengths, int32_t pad, float* d_out, void* cuda_stream)
{
    return extract_device(plan, true, d_audio, n_clips, clip_stride, clip_len, d_lengths, pad, d_out, nullptr, 0, 0,
                          cuda_stream);
}

int scf_extract_host_i16(const scf_plan* plan, const int16_t* h_pcm, int64_t n_clips, int64_t clip_stride,
                         int32_t clip_len, const int32_t* h_lengths, int32_t pad, float* h_out)
{
    return extract_host(plan, false, h_pcm, n_clips, clip_stride, clip_len, h_lengths, pad, h_out);
}

int scf_extract_host_f32(const scf_plan* plan, const float* h_audio, int64_t n_clips, int64_t clip_stride,
                         int32_t clip_len, const int32_t* h_lengths, int32_t pad, float* h_out)
{
    return extract_host(plan, true, h_audio, n_clips, clip_stride, clip_len, h_lengths, pad, h_out);
}

int scf_extract_host_i16_async(const scf_plan* plan, const int16_t* h_pcm, int64_t n_clips, int64_t clip_stride,
                               int32_t clip_len, const int32_t* h_lengths, int32_t pad, float* h_out)
{
    return extract_host_async(plan, h_pcm, n_clips, clip_stride, clip_len, h_lengths, pad, h_out);
}

int scf_host_sync(const scf_plan* plan) { return host_sync(plan); }

int scf_extract_i16_dlpack(const scf_plan* plan, const int16_t* d_pcm, int64_t n_clips, int64_t clip_stride,
                           int32_t clip_len, const int32_t* d_lengths, int32_t pad, void** dl_out, void* cuda_stream)
{
    if (!plan || !dl_out) return fail(SCF_ERR_INVALID, "NULL argument");
    *dl_out = nullptr;
    if (n_clips < 0 || clip_len < 0) return fail(SCF_ERR_INVALID, "negative size");
    const int64_t fpc = scf_num_frames(clip_len, plan->cfg.window, plan->cfg.hop);
    DeviceGuard guard(plan->device);
    if (!guard.ok) return fail(SCF_ERR_CUDA, "cudaSetDevice failed");
    const size_t bytes = std::max<size_t>((size_t)n_clips * fpc * plan->out_cols * sizeof(float), 256);
    float* d_out = nullptr;
    cudaError_t e = alloc_on_stream((void**)&d_out, bytes, plan->device, (cudaStream_t)cuda_stream);
    if (e != cudaSuccess) return fail(SCF_ERR_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
    if (pad == SCF_PAD_NONE && d_lengths) cudaMemsetAsync(d_out, 0, bytes, (cudaStream_t)cuda_stream);
    int rc = extract_device(plan, false, d_pcm, n_clips, clip_stride, clip_len, d_lengths, pad, d_out, nullptr, 0, 0,
                            cuda_stream);
    const int64_t shape[3] = {n_clips, fpc, plan->out_cols};
    DlOwner* o = rc ? nullptr : dl_make(d_out, plan->device, shape, 3);
    if (!o) {
        cudaFreeAsync(d_out, (cudaStream_t)cuda_stream);
        return rc ? rc : fail(SCF_ERR_ALLOC, "out of host memory");
    }
    *dl_out = &o->mt;
    return SCF_OK;
}

int scf_dlpack_alloc(int32_t device, const int64_t* shape, int32_t ndim, void** dl_out, void** d_ptr_out)
{
    if (!shape || !dl_out || ndim < 1 || ndim > 4) return fail(SCF_ERR_INVALID, "bad argument");
    *dl_out = nullptr;
    if (device < 0) SCF_CUDA(cudaGetDevice(&device));
    DeviceGuard guard(device);
    if (!guard.ok) return fail(SCF_ERR_CUDA, "cudaSetDevice failed");
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        if (shape[i] < 0) return fail(SCF_ERR_INVALID, "negative extent");
        n *= (size_t)shape[i];
    }
    void* d = nullptr;
    SCF_CUDA(cudaMalloc(&d, std::max<size_t>(n * sizeof(float), 256)));
    DlOwner* o = dl_make(d, device, shape, ndim);
    if (!o) { cudaFree(d); return fail(SCF_ERR_ALLOC, "out of host memory"); }
    *dl_out = &o->mt;
    if (d_ptr_out) *d_ptr_out = d;
    return SCF_OK;
}

int scf_dlpack_wrap(void* d_ptr, int32_t device, const int64_t* shape, int32_t ndim, void (*release)(void*),
                    void* release_ctx, void** dl_out)
{
    if (!d_ptr || !shape || !dl_out || !release || ndim < 1 || ndim > 4) return fail(SCF_ERR_INVALID, "bad argument");
    *dl_out = nullptr;
    if (device < 0) SCF_CUDA(cudaGetDevice(&device));
    DlOwner* o = dl_make(d_ptr, device, shape, ndim);
    if (!o) return fail(SCF_ERR_ALLOC, "out of host memory");
    o->release = release;
    o->release_ctx = release_ctx;
    *dl_out = &o->mt;
    return SCF_OK;
}

// ---- PyCapsule plumbing without linking libpython ----------------------------------------------
typedef int (*py_capsule_is_valid_t)(void*, const char*);
typedef void* (*py_capsule_get_pointer_t)(void*, const char*);
typedef void* (*py_capsule_new_t)(void*, const char*, void (*)(void*));

static void capsule_destructor(void* capsule)
{
    static py_capsule_is_valid_t is_valid = (py_capsule_is_valid_t)dlsym(RTLD_DEFAULT, "PyCapsule_IsValid");
    static py_capsule_get_pointer_t get_ptr = (py_capsule_get_pointer_t)dlsym(RTLD_DEFAULT, "PyCapsule_GetPointer");
    if (!is_valid || !get_ptr) return;
    if (is_valid(capsule, "dltensor")) {          // not consumed: still ours to free
        DLManagedTensor* mt = (DLManagedTensor*)get_ptr(capsule, "dltensor");
        if (mt && mt->deleter) mt->deleter(mt);
    }
}

void* scf_dlpack_make_capsule(void* dl_managed_tensor)
{
    static py_capsule_new_t cap_new = (py_capsule_new_t)dlsym(RTLD_DEFAULT, "PyCapsule_New");
    if (!cap_new || !dl_managed_tensor) {
        fail(SCF_ERR_INVALID, "PyCapsule_New not resolvable (not inside a Python process?) or NULL tensor");
        return nullptr;
    }
    return cap_new(dl_managed_tensor, "dltensor", capsule_destructor);
}

int scf_extract_i16_gather(const scf_plan* plan, const int16_t* d_pcm, int64_t n_local, int64_t clip_stride,
                           int32_t clip_len, float* const* d_peer_out, int32_t world, int32_t rank,
                           int64_t clips_per_rank, void* cuda_stream)
{
    if (!d_peer_out) return fail(SCF_ERR_INVALID, "peer table is NULL");
    return extract_device(plan, false, d_pcm, n_local, clip_stride, clip_len, nullptr, SCF_PAD_FRONT_ZERO, nullptr,
                          d_peer_out, world, rank, cuda_stream, clips_per_rank);
}

int scf_extract_i16_gather_multicast(const scf_plan* plan, const int16_t* d_pcm, int64_t n_local, int64_t clip_stride,
                                     int32_t clip_len, float* d_multicast_out, int32_t world, int32_t rank,
                                     int64_t clips_per_rank, void* cuda_stream)
{
    if (!d_multicast_out) return fail(SCF_ERR_INVALID, "multicast pointer is NULL");
    float* table[1] = {d_multicast_out};
    return extract_device(plan, false, d_pcm, n_local, clip_stride, clip_len, nullptr, SCF_PAD_FRONT_ZERO, nullptr,
                          table, world, rank, cuda_stream, clips_per_rank, nullptr, true);
}

// ---- device memory / CUDA IPC helpers ------------------------------------------------------------
int scf_device_malloc(int32_t device, int64_t bytes, void** d_ptr_out)
{
    if (!d_ptr_out || bytes < 0) return fail(SCF_ERR_INVALID, "bad argument");
    *d_ptr_out = nullptr;
    if (device < 0) SCF_CUDA(cudaGetDevice(&device));
    DeviceGuard guard(device);
    if (!guard.ok) return fail(SCF_ERR_CUDA, "cudaSetDevice failed");
    SCF_CUDA(cudaMalloc(d_ptr_out, (size_t)std::max<int64_t>(bytes, 256)));
    return SCF_OK;
}

int scf_device_free(int32_t device, void* d_ptr)
{
    if (!d_ptr) return SCF_OK;
    if (device < 0) SCF_CUDA(cudaGetDevice(&device));
    DeviceGuard guard(device);
    SCF_CUDA(cudaFree(d_ptr));
    return SCF_OK;
}

int scf_memcpy(int32_t device, void* dst, const void* src, int64_t bytes, int32_t kind, void* cuda_stream)
{
    if (bytes < 0 || kind < 0 || kind > 2) return fail(SCF_ERR_INVALID, "bad argument");
    if (bytes == 0) return SCF_OK;
    if (!dst || !src) return fail(SCF_ERR_INVALID, "NULL pointer");
    if (device < 0) SCF_CUDA(cudaGetDevice(&device));
    DeviceGuard guard(device);
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    SCF_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, k, (cudaStream_t)cuda_stream));
    if (!cuda_stream) SCF_CUDA(cudaStreamSynchronize(nullptr));
    return SCF_OK;
}

int scf_ipc_export(int32_t device, void* d_ptr, uint8_t* handle64_out)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    if (!d_ptr || !handle64_out) return fail(SCF_ERR_INVALID, "NULL argument");
    if (device < 0) SCF_CUDA(cudaGetDevice(&device));
    DeviceGuard guard(device);
    cudaIpcMemHandle_t h;
    SCF_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle64_out, &h, 64);
    return SCF_OK;
}

int scf_ipc_import(int32_t device, const uint8_t* handle64, void** d_ptr_out)
{
    if (!handle64 || !d_ptr_out) return fail(SCF_ERR_INVALID, "NULL argument");
    *d_ptr_out = nullptr;
    if (device < 0) SCF_CUDA(cudaGetDevice(&device));
    DeviceGuard guard(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SCF_CUDA(cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return SCF_OK;
}

int scf_ipc_close(int32_t device, void* d_ptr)
{
    if (!d_ptr) return SCF_OK;
    if (device < 0) SCF_CUDA(cudaGetDevice(&device));
    DeviceGuard guard(device);
    SCF_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return SCF_OK;
}

// ---- NCCL (resolved lazily; the library itself does not link against libnccl) -------------------
int scf_allgather_nccl(void* nccl_comm, const float* d_local, int64_t n_local_floats, float* d_all, void* cuda_stream)
{
    typedef int (*allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
    static allgather_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (h) fn = (allgather_fn)dlsym(h, "ncclAllGather");
    });
    if (!fn) return fail(SCF_ERR_NCCL, "libnccl.so.2 / ncclAllGather not found");
    if (!nccl_comm || !d_local || !d_all || n_local_floats < 0) return fail(SCF_ERR_INVALID, "bad argument");
    const int ncclFloat32 = 7;
    const int rc = fn(d_local, d_all, (size_t)n_local_floats, ncclFloat32, nccl_comm, (cudaStream_t)cuda_stream);
    if (rc != 0) return fail(SCF_ERR_NCCL, "ncclAllGather failed with code " + std::to_string(rc));
    return SCF_OK;
}

// ---- streaming --------------------------------------------------------------------------------
int scf_stream_create(const scf_plan* plan, int32_t n_streams, int32_t ring_rows, int32_t max_chunk,
                      scf_stream** stream_out)
{
    if (!plan || !stream_out) return fail(SCF_ERR_INVALID, "NULL argument");
    *stream_out = nullptr;
    if (n_streams < 1 || ring_rows < 1 || max_chunk < 1) return fail(SCF_ERR_INVALID, "sizes must be positive");
    if (plan->cfg.output == SCF_OUT_POWER) return fail(SCF_ERR_INVALID, "streams need a bank or cepstrum plan");
    DeviceGuard guard(plan->device);
    if (!guard.ok) return fail(SCF_ERR_CUDA, "cudaSetDevice failed");
    scf_stream* s = new (std::nothrow) scf_stream();
    if (!s) return fail(SCF_ERR_ALLOC, "out of host memory");
    s->plan = plan;
    s->max_chunk = max_chunk;
    s->n_streams = n_streams;
    s->ring_rows = ring_rows;
    s->cols = plan->base_cols;
    // carry < window before a push (listen.py:106 leaves len - k*hop < window), so window-1+max_chunk bounds it
    s->carry_cap = ((plan->cfg.window - 1 + max_chunk) + 7) & ~7;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        if ((e = cudaMalloc((void**)&s->carry[i], (size_t)n_streams * s->carry_cap * 2)) != cudaSuccess) break;
        if ((e = cudaMalloc((void**)&s->carry_len[i], (size_t)n_streams * 4)) != cudaSuccess) break;
        e = cudaMalloc((void**)&s->ring[i], (size_t)n_streams * ring_rows * s->cols * 4);
    }
    if (e == cudaSuccess && plan->cfg.delta != SCF_DELTA_NONE)
        e = cudaMalloc((void**)&s->ring_wide, (size_t)n_streams * ring_rows * plan->out_cols * 4);
    if (e != cudaSuccess || (e = cudaMalloc((void**)&s->n_new, (size_t)n_streams * 4)) != cudaSuccess ||
        (e = cudaMalloc((void**)&s->d_chunk_stage, (size_t)n_streams * max_chunk * 2)) != cudaSuccess) {
        scf_stream_destroy(s);
        return fail(SCF_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    int rc = scf_stream_reset(s, nullptr);
    if (rc) { scf_stream_destroy(s); return rc; }
    cudaStreamSynchronize(nullptr);
    *stream_out = s;
    return SCF_OK;
}

void scf_stream_destroy(scf_stream* s)
{
    if (!s) return;
    DeviceGuard guard(s->plan->device);
    for (int i = 0; i < 2; ++i) { cudaFree(s->carry[i]); cudaFree(s->carry_len[i]); cudaFree(s->ring[i]); }
    cudaFree(s->ring_wide);
    cudaFree(s->n_new);
    cudaFree(s->d_chunk_stage);
    s->small.release();
    if (s->pin) cudaFreeHost(s->pin);
    delete s;
}

int scf_stream_reset(scf_stream* s, void* cuda_stream)
{
    if (!s) return fail(SCF_ERR_INVALID, "stream is NULL");
    DeviceGuard guard(s->plan->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    // only the set the next push reads has to be cleared: window_audio = [], mfccs = zeros (listen.py:91-92)
    SCF_CUDA(cudaMemsetAsync(s->carry_len[s->cur], 0, (size_t)s->n_streams * 4, st));
    SCF_CUDA(cudaMemsetAsync(s->ring[s->cur], 0, (size_t)s->n_streams * s->ring_rows * s->cols * 4, st));
    SCF_CUDA(cudaMemsetAsync(s->n_new, 0, (size_t)s->n_streams * 4, st));
    return SCF_OK;
}

int scf_stream_push_i16(scf_stream* s, const int16_t* d_chunks, int32_t chunk_len, float* d_ring_out,
                        int32_t* d_new_rows, void* cuda_stream)
{
    if (!s || !d_chunks) return fail(SCF_ERR_INVALID, "NULL argument");
    if (chunk_len < 1 || chunk_len > s->max_chunk) return fail(SCF_ERR_INVALID, "chunk_len outside 1..max_chunk");
    const scf_plan* plan = s->plan;
    // ONE launch: the extract kernel reads concat(carry, chunk), writes the new ring rows and carries the state over
    StreamStep ss;
    memset(&ss, 0, sizeof(ss));
    const int in = s->cur, out = s->cur ^ 1;
    ss.chunks = d_chunks;
    ss.chunk_len = chunk_len;
    ss.carry_cap = s->carry_cap;
    ss.carry_in = s->carry[in];
    ss.carry_out = s->carry[out];
    ss.len_in = s->carry_len[in];
    ss.len_out = s->carry_len[out];
    ss.ring_in = s->ring[in];
    ss.ring_out = s->ring[out];
    ss.ring_copy = d_ring_out;
    ss.copy_pitch = plan->out_cols;
    ss.n_new = s->n_new;
    ss.n_new_copy = d_new_rows;
    ss.ring_rows = s->ring_rows;
    int rc = extract_device(plan, false, s->carry[in], s->n_streams, s->carry_cap, s->carry_cap, nullptr, SCF_PAD_NONE,
                            s->ring[out], nullptr, 0, 0, cuda_stream, 0, &ss);
    if (rc) return rc;
    s->cur = out;
    return SCF_OK;
}

int scf_stream_push_host_i16(scf_stream* s, const int16_t* h_chunks, int32_t chunk_len, float* h_ring_out,
                             int32_t* h_new_rows)
{
    if (!s || !h_chunks) return fail(SCF_ERR_INVALID, "NULL argument");
    if (chunk_len < 1 || chunk_len > s->max_chunk) return fail(SCF_ERR_INVALID, "chunk_len outside 1..max_chunk");
    DeviceGuard guard(s->plan->device);
    {   // few streams (the single-microphone Listener): chunk in, ring + counters out through mapped host memory
        const size_t cb = (((size_t)s->n_streams * chunk_len * 2) + 15) & ~(size_t)15;
        const size_t rb = (size_t)s->n_streams * s->ring_rows * s->plan->out_cols * 4;
        const size_t nb = (size_t)s->n_streams * 4;
        if (cb + rb + nb <= kSmallCallBytes && s->small.ensure(kSmallCallBytes) == 0) {
            memcpy(s->small.h, h_chunks, (size_t)s->n_streams * chunk_len * 2);
            float* d_ring = h_ring_out ? reinterpret_cast<float*>(s->small.d + cb) : nullptr;
            int32_t* d_new = h_new_rows ? reinterpret_cast<int32_t*>(s->small.d + cb + rb) : nullptr;
            // (the default stream, like the copying path below: ordered behind whatever this stream object did before)
            int rc = scf_stream_push_i16(s, reinterpret_cast<const int16_t*>(s->small.d), chunk_len, d_ring, d_new, nullptr);
            cudaError_t se = cudaStreamSynchronize(nullptr);
            if (rc) return rc;
            if (se != cudaSuccess) return fail(SCF_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(se));
            if (h_ring_out) memcpy(h_ring_out, s->small.h + cb, rb);
            if (h_new_rows) memcpy(h_new_rows, s->small.h + cb + rb, nb);
            return SCF_OK;
        }
    }
    // Many streams: the caller's (pageable) chunk and ring arrays go through the stream object's own pinned buffers, moved
    // by the copy threads -- copies from / to pageable memory are bounce copies inside the driver and made this call
    // 237 us for 256 streams, of which the step itself is 20.
    const size_t cbytes = (size_t)s->n_streams * chunk_len * 2;
    const size_t rbytes = (size_t)s->n_streams * s->ring_rows * s->plan->out_cols * 4;
    const size_t nbytes = (size_t)s->n_streams * 4;
    const size_t c_al = (cbytes + 255) & ~(size_t)255, r_al = (rbytes + 255) & ~(size_t)255;
    if (!s->pin || s->pin_bytes < c_al + r_al + nbytes) {
        if (s->pin) cudaFreeHost(s->pin);
        s->pin = nullptr;
        s->pin_bytes = 0;
        const size_t want = (((size_t)s->n_streams * s->max_chunk * 2 + 255) & ~(size_t)255) + r_al + nbytes;
        SCF_CUDA(cudaHostAlloc((void**)&s->pin, want, cudaHostAllocDefault));
        s->pin_bytes = want;
    }
    CopyPool& pool = CopyPool::get();
    pool.copy(s->pin, h_chunks, cbytes);
    SCF_CUDA(cudaMemcpyAsync(s->d_chunk_stage, s->pin, cbytes, cudaMemcpyHostToDevice, nullptr));
    // delta plans hand out wide rows: the step writes them (and their delta columns) into the stream's wide copy
    float* wide = (h_ring_out && s->ring_wide) ? s->ring_wide : nullptr;
    int rc = scf_stream_push_i16(s, s->d_chunk_stage, chunk_len, wide, nullptr, nullptr);
    if (rc) return rc;
    if (h_ring_out) {       // the new ring is the state buffer the push just wrote (or its wide copy)
        const float* src = wide ? wide : s->ring[s->cur];
        SCF_CUDA(cudaMemcpyAsync(s->pin + c_al, src, rbytes, cudaMemcpyDeviceToHost, nullptr));
    }
    if (h_new_rows) SCF_CUDA(cudaMemcpyAsync(s->pin + c_al + r_al, s->n_new, nbytes, cudaMemcpyDeviceToHost, nullptr));
    SCF_CUDA(cudaStreamSynchronize(nullptr));
    if (h_ring_out) pool.copy(h_ring_out, s->pin + c_al, rbytes);
    if (h_new_rows) memcpy(h_new_rows, s->pin + c_al + r_al, nbytes);
    return SCF_OK;
}

// ---- misc -------------------------------------------------------------------------------------
int scf_parallel_memcpy(void* dst, const void* src, int64_t bytes)
{
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) return fail(SCF_ERR_INVALID, "bad argument");
    if (bytes > 0) CopyPool::get().copy(dst, src, (size_t)bytes);
    return SCF_OK;
}

const char* scf_last_error(void) { return g_err.c_str(); }
int scf_version(void) { return SCF_VERSION; }
int64_t scf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int scf_measure_fp32_flops(int32_t device, double* flops_out)
{
    if (!flops_out) return fail(SCF_ERR_INVALID, "NULL argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(SCF_ERR_NO_DEVICE, "no CUDA device");
    }
    if (device < 0) SCF_CUDA(cudaGetDevice(&device));
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    SCF_CUDA(cudaGetDeviceProperties(&prop, device));
    const int grid = prop.multiProcessorCount * 8, iters = 4096;
    float* d = nullptr;
    SCF_CUDA(cudaMalloc((void**)&d, (size_t)grid * 256 * 4));
    cudaEvent_t e0, e1;
    SCF_CUDA(cudaEventCreate(&e0));
    SCF_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        SCF_CUDA(cudaEventRecord(e0, nullptr));
        SCF_CUDA(launch_fp32_probe(d, iters, grid, nullptr));
        SCF_CUDA(cudaEventRecord(e1, nullptr));
        SCF_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SCF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8 * 16 * (double)iters * grid * 256 / (ms * 1e-3);
        if (rep > 0) best = std::max(best, flops);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *flops_out = best;
    return SCF_OK;
}

}  // extern "C"
