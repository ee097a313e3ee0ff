// Streaming post-processing on the device: listen.py's ThresholdDecoder (:452-521) and
// TriggerDetector (:525-559; C++ twins inference/tflite/threshold_decoder.h:19-113, speech_commands.h:263-289) for
// n_streams concurrent listeners.  One launch does what one iteration of the listen.py loop does after the model
// (listen.py:411-425): arg-max and max of the class scores, decode of a non-background score, trigger update.
// The state machines are scalar and tiny; the point is that nothing per-stream is left on the host between the
// feature kernel, a batched model and the activation flags.  Arithmetic is float64 like the Python path.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "scfeat_internal.h"

struct scf_post {
    int device = 0;
    // ThresholdDecoder (listen.py:466-471)
    int32_t min_out = 0, max_out = 0, out_range = 0;
    int64_t n_cd = 0;
    double center = 0.5;
    double* d_cd = nullptr;
    // TriggerDetector (listen.py:529-536) for n_streams streams
    int32_t n_streams = 0, n_classes = 0, chunk_size = 0, trigger_level = 0;
    double sensitivity = 0.5;
    uint8_t* d_is_background = nullptr;      // [n_classes]
    int32_t* d_activation = nullptr;         // [n_streams]
    int32_t* d_record_index = nullptr;       // [n_streams], -1 = None
};

namespace scf {

int post_fail(int code, const char* msg);   // scfeat_host.cu: sets scf_last_error

struct DecodeParams {
    const double* cd;
    int64_t n_cd;
    int32_t min_out, out_range;
    double center;
};

// ThresholdDecoder.decode (listen.py:497-509).  `logit_arg` is 1 / x - 1 as the caller's number type computes it.
__device__ __forceinline__ double decode_tail(const DecodeParams& d, double raw, bool inside, double logit_arg)
{
    if (raw == 1.0 || raw == 0.0) return raw;
    double cp;
    if (d.out_range == 0) {
        cp = raw > (double)d.min_out ? 1.0 : 0.0;
    } else {
        const double asig = inside ? -log(logit_arg) : -10.0;                 // asigmoid, listen.py:479-484
        double ratio = (asig - (double)d.min_out) / (double)d.out_range;
        ratio = fmin(fmax(ratio, 0.0), 1.0);
        cp = d.cd[(int64_t)(ratio * (double)(d.n_cd - 1) + 0.5)];
    }
    if (cp < d.center) return 0.5 * cp / d.center;
    return 0.5 + 0.5 * (cp - d.center) / (1.0 - d.center);
}

__global__ void post_decode_kernel(DecodeParams d, const double* __restrict__ raw, int64_t n, double* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = raw[i];
    out[i] = decode_tail(d, x, x > 0.0 && x < 1.0, 1.0 / x - 1.0);
}

struct StepParams {
    DecodeParams dec;
    const float* probs;            // [n_streams][n_classes]; NULL: take (index_in, score_in) as given
    const int32_t* index_in;       // [n_streams]
    const double* score_in;        // [n_streams]
    const uint8_t* is_background;  // [n_classes]
    int32_t* activation;           // [n_streams]
    int32_t* record_index;         // [n_streams]
    int32_t* index_out;            // nullable
    double* score_out;             // nullable
    uint8_t* fired_out;            // nullable
    int32_t n_streams, n_classes, trigger_level, reset_value;
    double sensitivity;
};

// One thread per stream: listen.py:411-425.
__global__ void post_step_kernel(StepParams p)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.n_streams) return;
    int index;
    double score;
    bool background;
    if (p.probs != nullptr) {
        const float* pr = p.probs + (int64_t)s * p.n_classes;
        index = 0;                                                   // np.argmax: the first maximum
        float best = pr[0];
        for (int c = 1; c < p.n_classes; ++c) {
            const float v = pr[c];
            if (v > best) { best = v; index = c; }
        }
        background = p.is_background[index] != 0;
        score = (double)best;
        if (!background) {
            // the model's score is float32 and the reference evaluates 1 / x - 1 on it in float32 before math.log
            // takes over in double (numpy scalar arithmetic, listen.py:484)
            const float arg = __fsub_rn(__fdiv_rn(1.0f, best), 1.0f);
            score = decode_tail(p.dec, (double)best, best > 0.f && best < 1.f, (double)arg);
        }
    } else {                                                         // TriggerDetector.update(index, score) alone
        index = min(max(p.index_in[s], 0), p.n_classes - 1);
        score = p.score_in[s];
        background = p.is_background[index] != 0;
    }
    // TriggerDetector.update (listen.py:538-559)
    int act = p.activation[s];
    const int rec = p.record_index[s];
    bool fired = false;
    bool record = true;
    if (!background && index == rec && score > p.sensitivity) {
        act += 1;
        if (act > p.trigger_level) {
            act = p.reset_value;                                     // -(8 * 2048) // chunk_size
            fired = true;
            record = false;                                          // the reference returns before recording the index
        }
    } else if (act < 0) {
        act += 1;
    } else if (act > 0) {
        act -= 1;
    }
    p.activation[s] = act;
    if (record) p.record_index[s] = index;
    if (p.index_out) p.index_out[s] = index;
    if (p.score_out) p.score_out[s] = score;
    if (p.fired_out) p.fired_out[s] = fired ? 1 : 0;
}

// numpy.linspace(start, stop, num): arange(num) * step + start, last element := stop
static void linspace_into(double start, double stop, int64_t num, std::vector<double>& v)
{
    v.resize((size_t)num);
    if (num == 1) { v[0] = start; return; }
    const double step = (stop - start) / (double)(num - 1);
    for (int64_t i = 0; i < num; ++i) v[(size_t)i] = (double)i * step + start;
    if (num > 0) v[(size_t)num - 1] = stop;
}

// ThresholdDecoder.__init__ / _calc_pd (listen.py:466-471, 519-521): cd = cumsum(sum_i pdf(points; mu_i, std_i) /
// (resolution * n)), points = linspace(min_out, max_out, resolution * out_range)
static int build_cd(const double* mu_stds, int32_t n, int32_t resolution, double min_z, double max_z, int32_t& min_out,
                    int32_t& max_out, std::vector<double>& cd)
{
    if (!mu_stds || n < 1 || resolution < 1) return SCF_ERR_INVALID;
    double lo = 0, hi = 0;
    for (int i = 0; i < n; ++i) {
        const double a = mu_stds[2 * i] + min_z * mu_stds[2 * i + 1], b = mu_stds[2 * i] + max_z * mu_stds[2 * i + 1];
        lo = i ? fmin(lo, a) : a;
        hi = i ? fmax(hi, b) : b;
    }
    min_out = (int32_t)lo;                                           // int(): truncation towards zero
    max_out = (int32_t)hi;
    const int64_t num = (int64_t)resolution * (max_out - min_out);
    std::vector<double> pts;
    linspace_into((double)min_out, (double)max_out, num, pts);
    cd.assign((size_t)std::max<int64_t>(num, 0), 0.0);
    for (int i = 0; i < n; ++i) {                                    // np.sum(axis=0): component after component
        const double mu = mu_stds[2 * i], sd = mu_stds[2 * i + 1];
        if (sd == 0) continue;
        const double norm = 1.0 / (sd * sqrt(2 * M_PI));
        for (int64_t k = 0; k < num; ++k) {
            const double x = pts[(size_t)k];
            cd[(size_t)k] += norm * exp(-((x - mu) * (x - mu)) / (2 * (sd * sd)));
        }
    }
    const double div = (double)resolution * (double)n;
    double run = 0;
    for (int64_t k = 0; k < num; ++k) {                              // np.cumsum: sequential
        run += cd[(size_t)k] / div;
        cd[(size_t)k] = run;
    }
    return SCF_OK;
}

static DecodeParams decode_params(const scf_post* p)
{
    DecodeParams d;
    d.cd = p->d_cd;
    d.n_cd = p->n_cd;
    d.min_out = p->min_out;
    d.out_range = p->out_range;
    d.center = p->center;
    return d;
}

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace scf

using namespace scf;

#define POST_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return post_fail(SCF_ERR_CUDA, (std::string(#call) + ": " + cudaGetErrorString(e__)).c_str()); \
    } while (0)

extern "C" {

int scf_post_build_cd(const double* mu_stds, int32_t n_mu, int32_t resolution, double min_z, double max_z,
                      int32_t* min_out, int32_t* max_out, double* cd_out, int64_t* n_cd)
{
    if (!min_out || !max_out || !n_cd) return post_fail(SCF_ERR_INVALID, "NULL argument");
    std::vector<double> cd;
    int32_t lo = 0, hi = 0;
    if (build_cd(mu_stds, n_mu, resolution, min_z, max_z, lo, hi, cd)) return post_fail(SCF_ERR_INVALID, "bad decoder configuration");
    *min_out = lo;
    *max_out = hi;
    if (cd_out) {
        if (*n_cd < (int64_t)cd.size()) return post_fail(SCF_ERR_INVALID, "cd_out is too small");
        memcpy(cd_out, cd.data(), cd.size() * sizeof(double));
    }
    *n_cd = (int64_t)cd.size();
    return SCF_OK;
}

int scf_post_create(const double* mu_stds, int32_t n_mu, double center, int32_t resolution, double min_z, double max_z,
                    const uint8_t* class_is_background, int32_t n_classes, int32_t n_streams, int32_t chunk_size,
                    double sensitivity, int32_t trigger_level, int32_t device, scf_post** post_out)
{
    if (!post_out) return post_fail(SCF_ERR_INVALID, "post_out is NULL");
    *post_out = nullptr;
    if (!class_is_background || n_classes < 1 || n_streams < 1 || chunk_size < 1)
        return post_fail(SCF_ERR_INVALID, "classes, streams and chunk_size must be positive");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return post_fail(SCF_ERR_NO_DEVICE, "no CUDA device: libscfeat has no CPU fallback");
    }
    if (device < 0) POST_CUDA(cudaGetDevice(&device));
    if (device >= n_dev) return post_fail(SCF_ERR_INVALID, "device ordinal out of range");
    std::vector<double> cd;
    int32_t lo = 0, hi = 0;
    if (build_cd(mu_stds, n_mu, resolution, min_z, max_z, lo, hi, cd)) return post_fail(SCF_ERR_INVALID, "bad decoder configuration");
    scf_post* p = new (std::nothrow) scf_post();
    if (!p) return post_fail(SCF_ERR_ALLOC, "out of host memory");
    DevGuard guard(device);
    p->device = device;
    p->min_out = lo;
    p->max_out = hi;
    p->out_range = hi - lo;
    p->n_cd = (int64_t)cd.size();
    p->center = center;
    p->n_streams = n_streams;
    p->n_classes = n_classes;
    p->chunk_size = chunk_size;
    p->trigger_level = trigger_level;
    p->sensitivity = sensitivity;
    cudaError_t e = cudaMalloc((void**)&p->d_cd, std::max<size_t>(cd.size(), 1) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_is_background, (size_t)n_classes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_activation, (size_t)n_streams * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_record_index, (size_t)n_streams * 4);
    if (e == cudaSuccess && !cd.empty()) e = cudaMemcpy(p->d_cd, cd.data(), cd.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_is_background, class_is_background, (size_t)n_classes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(p->d_activation, 0, (size_t)n_streams * 4);
    if (e == cudaSuccess) e = cudaMemset(p->d_record_index, 0xff, (size_t)n_streams * 4);
    if (e != cudaSuccess) {
        scf_post_destroy(p);
        return post_fail(SCF_ERR_CUDA, (std::string("scf_post_create: ") + cudaGetErrorString(e)).c_str());
    }
    *post_out = p;
    return SCF_OK;
}

void scf_post_destroy(scf_post* p)
{
    if (!p) return;
    DevGuard guard(p->device);
    cudaFree(p->d_cd);
    cudaFree(p->d_is_background);
    cudaFree(p->d_activation);
    cudaFree(p->d_record_index);
    delete p;
}

int scf_post_reset(scf_post* p, void* cuda_stream)
{
    if (!p) return post_fail(SCF_ERR_INVALID, "post is NULL");
    DevGuard guard(p->device);
    POST_CUDA(cudaMemsetAsync(p->d_activation, 0, (size_t)p->n_streams * 4, (cudaStream_t)cuda_stream));
    POST_CUDA(cudaMemsetAsync(p->d_record_index, 0xff, (size_t)p->n_streams * 4, (cudaStream_t)cuda_stream));
    return SCF_OK;
}

int scf_post_decode(const scf_post* p, const double* d_raw, int64_t n, double* d_out, void* cuda_stream)
{
    if (!p) return post_fail(SCF_ERR_INVALID, "post is NULL");
    if (n < 0) return post_fail(SCF_ERR_INVALID, "negative size");
    if (n == 0) return SCF_OK;
    if (!d_raw || !d_out) return post_fail(SCF_ERR_INVALID, "NULL pointer");
    DevGuard guard(p->device);
    post_decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(decode_params(p), d_raw, n, d_out);
    POST_CUDA(cudaGetLastError());
    count_launch(1);
    return SCF_OK;
}

static int post_step(scf_post* p, const float* d_probs, const int32_t* d_index_in, const double* d_score_in,
                     int32_t* d_index_out, double* d_score_out, uint8_t* d_fired_out, void* cuda_stream)
{
    DevGuard guard(p->device);
    StepParams sp;
    sp.dec = decode_params(p);
    sp.probs = d_probs;
    sp.index_in = d_index_in;
    sp.score_in = d_score_in;
    sp.is_background = p->d_is_background;
    sp.activation = p->d_activation;
    sp.record_index = p->d_record_index;
    sp.index_out = d_index_out;
    sp.score_out = d_score_out;
    sp.fired_out = d_fired_out;
    sp.n_streams = p->n_streams;
    sp.n_classes = p->n_classes;
    sp.trigger_level = p->trigger_level;
    {   // Python floor division: -(8 * 2048) // chunk_size (listen.py:548)
        const int num = -(8 * 2048);
        int q = num / p->chunk_size;
        if ((num % p->chunk_size) != 0) --q;
        sp.reset_value = q;
    }
    sp.sensitivity = p->sensitivity;
    post_step_kernel<<<(unsigned)((p->n_streams + 127) / 128), 128, 0, (cudaStream_t)cuda_stream>>>(sp);
    POST_CUDA(cudaGetLastError());
    count_launch(1);
    return SCF_OK;
}

int scf_post_step(scf_post* p, const float* d_probs, int32_t* d_index_out, double* d_score_out, uint8_t* d_fired_out,
                  void* cuda_stream)
{
    if (!p || !d_probs) return post_fail(SCF_ERR_INVALID, "NULL argument");
    return post_step(p, d_probs, nullptr, nullptr, d_index_out, d_score_out, d_fired_out, cuda_stream);
}

int scf_post_trigger_update(scf_post* p, const int32_t* d_index, const double* d_score, uint8_t* d_fired_out, void* cuda_stream)
{
    if (!p || !d_index || !d_score) return post_fail(SCF_ERR_INVALID, "NULL argument");
    return post_step(p, nullptr, d_index, d_score, nullptr, nullptr, d_fired_out, cuda_stream);
}

int scf_post_state(const scf_post* p, int32_t* h_activation, int32_t* h_record_index, void* cuda_stream)
{
    if (!p) return post_fail(SCF_ERR_INVALID, "post is NULL");
    DevGuard guard(p->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (h_activation) POST_CUDA(cudaMemcpyAsync(h_activation, p->d_activation, (size_t)p->n_streams * 4, cudaMemcpyDeviceToHost, st));
    if (h_record_index) POST_CUDA(cudaMemcpyAsync(h_record_index, p->d_record_index, (size_t)p->n_streams * 4, cudaMemcpyDeviceToHost, st));
    POST_CUDA(cudaStreamSynchronize(st));
    return SCF_OK;
}

int scf_post_info(const scf_post* p, int32_t* min_out, int32_t* max_out, int64_t* n_cd)
{
    if (!p) return post_fail(SCF_ERR_INVALID, "post is NULL");
    if (min_out) *min_out = p->min_out;
    if (max_out) *max_out = p->max_out;
    if (n_cd) *n_cd = p->n_cd;
    return SCF_OK;
}

}  // extern "C"
