// sm_100a kernels of libscfeat: fused  PCM -> frames -> real FFT -> power -> filterbank -> log -> DCT.
//
// Replaces the arithmetic of sonopy.mfcc_spec / mel_spec / power_spec (reference call site
// common/data_utils.py:69), common/bark_feature.py:85-175 and the C++ twin
// inference/tflite/mfcc.h:295-456.  Design notes live in DESIGN.md; the short version:
//
//  * A persistent team of 8 warps walks "tiles" of 8*G frame PAIRS (G = 32 / R, R = n_fft / 32); a CTA holds one team
//    (2 or 3 CTAs per SM) or three (one 768-thread CTA per SM).
//  * FFT stage, one warp per G pairs, no team-level sync:
//      two real frames A, B are packed as z = A + iB and transformed by ONE complex n_fft-point FFT,
//      done as R-point FFTs (pass 1, lane = n2, stride-32 samples) and 32-point FFTs (pass 2,
//      lane = k1) entirely in registers (generated straight-line packed-FP32 code, fft_gen.cuh), with a
//      single shared-memory transpose in between.  The two spectra are separated with the mirror identity
//      A[k] = (Z[k] + conj Z[N-k]) / 2,  B[k] = (Z[k] - conj Z[N-k]) / 2i ; Z[N-k] comes from the mirror lane
//      by warp shuffle.  The factor 1/2, 1/n_fft and the PCM scale are folded into the filterbank weights.
//      The power spectra of the pair go to ONE shared-memory row, interleaved (|A[k]|^2, |B[k]|^2).
//  * Bank stage, team-wide: thread = (pair slot, thread group); the host cuts the sparse filterbank into
//    tasks of 8 bins x 2 filters (float4 reads of the pair row and of the weights, packed FMAs on both frames),
//    balanced over the groups; partial sums to the free tail of the exchange area.
//  * Epilogue: sum partials -> log(max(., eps)) -> DCT-II (or pass-through) -> global rows; optionally
//    the same rows are stored to every peer GPU's cache (fused all-gather over NVLink), or the launch is one
//    step of the streaming state machine (listen.py:96-114) and also carries ring and carry buffers over.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "fft_gen.cuh"
#include "scfeat_internal.h"

namespace scf {


#define SCF_EPS 2.220446049250313e-16f   // np.finfo(float).eps, common/bark_feature.py:77

// Exchange rows (pass 1 -> pass 2 transpose) hold R complex values plus 8 bytes of padding: 64-bit stores of one column
// by the 32 lanes and 64-bit loads of one row are then conflict-free without any address arithmetic (8448 bytes per warp
// at n_fft = 1024).
// After pass 2 the warp's exchange region is free: its head takes the power rows of the warp's pairs (one row per
// PAIR, the two frames interleaved as (|A[k]|^2, |B[k]|^2) so that the bank phase works on packed values), its tail
// takes a share of the team's partial-sum rows.
template <int R>
struct Geo {
    static constexpr int NFFT = 32 * R;
    static constexpr int NB = 16 * R;                 // highest bin index (n_fft / 2)
    static constexpr int LOG2R = (R == 32) ? 5 : (R == 16) ? 4 : 3;
    static constexpr int G = 32 / R;                  // frame pairs per warp
    static constexpr int PPT = kWarps * G;            // pairs per tile (one tile = one team's 8 warps) = bank-phase slots
    static constexpr int NGRP = kThreads / PPT;       // bank-phase thread groups
    static constexpr int XROW = 2 * R + 2;            // floats per exchange row
    static constexpr int XPAIR = 32 * XROW;
    static constexpr int XWARP = xwarp_floats(R);     // floats of shared memory owned by one warp
    static constexpr int PROW2 = pair_row_floats(R);  // floats per pair row ( = 4 mod 32 -> conflict-free float4 columns)
    static constexpr int NLOAD = R + R / 2;           // fast path: strided samples per lane covering both frames
    static_assert(XWARP == G * XPAIR && kWarps == kTeamWarps, "geometry helpers out of sync");
    static_assert(G * PROW2 + 28 <= p_free_floats(R) && p_free_floats(R) <= XWARP, "power rows must fit in the warp's exchange region");
    static_assert(XWARP % 32 == 0, "warp regions must start on bank 0");
    static_assert(drow_floats(R) == 2 * PPT, "one packed value per slot");
};

// shared-memory floats per team outside the exchange area (see the carve-up in the kernel)
template <int PPT>
__host__ __device__ inline int team_smem_floats(const KParams& p)
{
    const int n_lq = ((p.n_q > p.n_filt4 ? p.n_q : p.n_filt4) + 1) & ~1;
    return 2 * PPT * (n_lq + 2) + (p.n_peers != 0 ? 2 * PPT * (p.out_cols + 2) : 0);
}

// Samples in registers: int16 PCM is held sign-extended in 32 bits, so that the conversion is the full-rate
// I2FP.F32.S32 instead of the quarter-rate, long-latency I2F.S16 the compiler picks for a `short`.
template <typename InT> struct Raw { typedef float type; };
template <> struct Raw<int16_t> { typedef int32_t type; };
#ifdef SCF_I2F_XU
__device__ __forceinline__ float to_f32(int32_t v) { return (float)(int16_t)v; }
#else
__device__ __forceinline__ float to_f32(int32_t v) { return (float)v; }
#endif
__device__ __forceinline__ float to_f32(float v) { return v; }
// read-only sample loads
__device__ __forceinline__ int32_t ld_sample(const int16_t* p)
{
    int32_t v;
    asm volatile("ld.global.nc.s16 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_sample(const float* p) { return __ldg(p); }
// ... predicated: nothing is read and zero is returned when `on` is false (the address may then lie outside the buffer)
__device__ __forceinline__ int32_t ld_sample_if(const int16_t* p, bool on)
{
    int32_t v;
    asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n mov.b32 %0, 0;\n @q ld.global.nc.s16 %0, [%1];\n}"
                 : "=r"(v) : "l"(p), "r"((int)on));
    return v;
}
__device__ __forceinline__ float ld_sample_if(const float* p, bool on)
{
    float v;
    asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n mov.b32 %0, 0;\n @q ld.global.nc.f32 %0, [%1];\n}"
                 : "=f"(v) : "l"(p), "r"((int)on));
    return v;
}
__device__ __forceinline__ int32_t raw_or(int32_t a, int32_t b) { return a | b; }
__device__ __forceinline__ float raw_or(float a, float b) { return a + b; }      // one of them is 0
// bits that are set iff the sample is not zero (-0.0f counts as zero)
__device__ __forceinline__ uint32_t nz_bits(int32_t v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t nz_bits(float v) { return __float_as_uint(v) & 0x7fffffffu; }

#ifdef SCF_DEBUG_TIMES       // experiment: start / end time of every team (ns, %globaltimer), read back by scf_debug_times
__device__ unsigned long long g_dbg_times[8 * 2048];      // [0] start, [1] end, [2 + i] end of the team's tile i (i < 6)
__device__ unsigned long long g_dbg_phase[2 * 8 * 2048];  // [thread 0 | thread 224][phase][team]: SM cycles summed over tiles
__device__ __forceinline__ unsigned long long dbg_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

// natural logarithm of a normal, positive float
__device__ __forceinline__ float log_ge_eps(float x)
{
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    return l * 0.6931471805599453f;
}

template <int R>
__device__ __forceinline__ void fft_r(f2 (&x)[R])
{
    if constexpr (R == 32) fft32_p2(x);
    else if constexpr (R == 16) fft16_p2(x);
    else fft8_p2(x);
}

struct ClipGeom {
    int32_t len;        // valid samples
    int32_t pad;        // zeros in front (SCF_PAD_FRONT_ZERO)
    int32_t n_frames;   // rows this clip produces
};

__device__ __forceinline__ ClipGeom clip_geom(const KParams& p, int64_t clip)
{
    ClipGeom g;
    int32_t len = p.clip_len;
    if (p.stream_on) len = __ldg(p.stream.len_in + clip) + p.stream.chunk_len;      // concat(carry, chunk)
    else if (p.lengths != nullptr) len = min(max(__ldg(p.lengths + clip), 0), p.clip_len);
    g.len = len;
    if (p.pad_mode == SCF_PAD_FRONT_ZERO) {
        g.pad = p.clip_len - len;
        g.n_frames = p.frames_per_clip;
    } else {
        g.pad = 0;
        g.n_frames = (len >= p.window) ? (len - p.window) / p.hop + 1 : 0;
    }
    return g;
}

// Generic loader: any window / hop / per-clip length / pre-emphasis / window function; loads the two frames of a pair.
// Sample a of the clip comes from clip_base[a] for a < split and from tail[a - split] behind it (STREAM: carry, then
// the new chunk).  Every lane issues all its loads of BOTH frames back to back as predicated loads off two base
// pointers with immediate offsets (no per-sample address arithmetic; out-of-range positions read nothing and are
// zero), so the memory latency is paid once per pair, not once per sample.
// `mid` runs after the loads of both frames have been issued and before their values are used: work placed there
// (the state carry-over of a streaming step) overlaps the memory latency of the samples.
template <int R, typename InT, bool STREAM, typename Mid>
__device__ __forceinline__ void load_pair_generic(const KParams& p, const InT* __restrict__ clip_base,
                                                  const InT* __restrict__ tail, int split, const ClipGeom& cg, int frame_a,
                                                  int lane, float (&xa)[R], float (&xb)[R], uint32_t& nz_a, uint32_t& nz_b,
                                                  Mid&& mid)
{
    typedef typename Raw<InT>::type RawT;
    // frame f covers clip samples s0 + n, n < w_eff; n is valid for lo <= n < lo + width
    struct Span { int s0, lo; unsigned width; };
    auto span_of = [&](int frame) {
        Span sp;
        sp.s0 = frame * p.hop - cg.pad;
        sp.lo = max(0, -sp.s0);
        const int hi = frame < cg.n_frames ? min(p.w_eff, cg.len - sp.s0) : 0;
        sp.width = (unsigned)max(0, hi - sp.lo);
        return sp;
    };
    const Span sa = span_of(frame_a), sb = span_of(frame_a + 1);
    auto fetch = [&](const Span& sp, int back, RawT (&raw)[R]) {
        const InT* p0 = clip_base + (sp.s0 - back) + lane;               // sample a - back for n = lane
        const InT* p1 = tail + (sp.s0 - back - split) + lane;
        const int rel = lane - sp.lo;                                     // n - lo for n = lane
        const int a0 = sp.s0 - back + lane;                               // a - back for n = lane
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int o = 32 * scf_bitrev(i, Geo<R>::LOG2R);
            const bool in = (unsigned)(rel + o) < sp.width && a0 + o >= 0;
            if constexpr (STREAM) {
                const bool first = a0 + o < split;
                raw[i] = raw_or(ld_sample_if(p0 + o, in && first), ld_sample_if(p1 + o, in && !first));
            } else {
                raw[i] = ld_sample_if(p0 + o, in);
            }
        }
    };
    RawT ra[R], rb[R];
    fetch(sa, 0, ra);
    fetch(sb, 0, rb);
    mid();
#pragma unroll
    for (int i = 0; i < R; ++i) { xa[i] = to_f32(ra[i]); xb[i] = to_f32(rb[i]); }
    if (p.preemph != 0.f) {                                   // x[a] - alpha * x[a-1], x[-1] := 0 (mfcc.h:394-403)
        fetch(sa, 1, ra);
        fetch(sb, 1, rb);
#pragma unroll
        for (int i = 0; i < R; ++i) {
            xa[i] = fmaf(-p.preemph, to_f32(ra[i]), xa[i]);
            xb[i] = fmaf(-p.preemph, to_f32(rb[i]), xb[i]);
        }
    }
    if (p.win != nullptr) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int n = lane + 32 * scf_bitrev(i, Geo<R>::LOG2R);
            const float w = ld_sample_if(p.win + n, n < p.w_eff);
            xa[i] *= w;
            xb[i] *= w;
        }
    }
    nz_a = nz_b = 0;
#pragma unroll
    for (int i = 0; i < R; ++i) {
        nz_a |= nz_bits(xa[i]);
        nz_b |= nz_bits(xb[i]);
    }
}

// ---- small PTX helpers: mbarrier + TMA 1-D bulk copy (tables -> shared memory) -----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

// n / d for the plan's pairs_per_clip with a precomputed magic (Hacker's Delight, unsigned round-up method)
// (magic == 0 marks a power-of-two divisor: plain shift)
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint32_t magic, uint32_t shift)
{
    if (magic == 0) return n >> shift;
    const uint32_t t = __umulhi(n, magic);
    return (t + ((n - t) >> 1)) >> shift;
}

// TEAMS = 1, DENSE = false ("classic"): a CTA is one team of 8 warps; 2 CTAs per SM, 128 registers per thread, every
//            table in shared memory, next tile's samples prefetched into registers.
// DENSE = true: 80 registers per thread (the packed FFT fits) -> 24 warps per SM, which is what hides the FFMA2 / LDS
//            latencies:
//   TEAMS = 1: 3 CTAs of one team per SM -- small jobs (CTAs of the next launch backfill through PDL); the bank
//              weights and the DCT matrix are read through L1 instead of shared memory to fit three CTAs;
//   TEAMS = 3: one 768-thread CTA of three independent teams that share one table copy and synchronise through their
//              own named barriers -- large jobs.
// PRE (fast geometry only): pre-emphasis x[n] - alpha * x[n-1] (x[-1] := 0) and / or an analysis window fused into
//            the loader (inference/tflite/mfcc.h:394-410): every lane also loads the sample in front of each of its
//            samples (the same cache lines), and the window values of its positions come through L1.
template <int R, typename InT, bool FAST, int TEAMS, bool DENSE, bool PRE>
__global__ void __launch_bounds__(kThreads* TEAMS, TEAMS == 3 ? 1 : (DENSE ? 3 : kCtasPerSm))
    extract_kernel(const KParams p, const uint32_t n_tiles)
{
    static_assert(TEAMS == 1 || DENSE, "multi-team CTAs need the dense register budget");
    static_assert(FAST || !PRE, "the generic loader has its own pre-emphasis / window code");
    static_assert(!(PRE && DENSE), "the fused front end needs a second set of sample registers: classic variant only");
    using geo = Geo<R>;
    extern __shared__ __align__(16) float smem[];

    // (the shuffle tells the compiler that warp-level quantities are warp-uniform: they then live in uniform registers
    //  instead of the 80 vector registers the packed FFT needs for itself)
    const int warp_in_cta = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int team = warp_in_cta / kWarps;
    const int warp = warp_in_cta % kWarps;           // warp within its team
    const int lane = threadIdx.x & 31;
    const int tid = warp * 32 + lane;                // thread within its team

    // ---- shared memory carve-up (must match extract_smem_bytes; the table part mirrors the plan's blob) ----
    float* s_xch = smem + team * (kWarps * geo::XWARP);
    unsigned char* s_tab = reinterpret_cast<unsigned char*>(smem + TEAMS * kWarps * geo::XWARP);
    // what is staged in shared memory: everything, except that the 3-CTAs-per-SM variant leaves the pass-2 twiddles and
    // the DCT matrix in global memory (read through L1) -- that is what makes its 75 KB budget
    constexpr bool kTwInL1 = DENSE && TEAMS == 1;
    const int staged_bytes = kTwInL1 ? p.table_small_bytes : p.table_bytes;
    const unsigned char* big_tab = kTwInL1 ? static_cast<const unsigned char*>(p.tables) : s_tab;
    const float4* s_tw4 = reinterpret_cast<const float4*>(big_tab + p.off_tw);
    const float4* s_wts4 = reinterpret_cast<const float4*>(s_tab + p.off_wts);
    const float* s_dct = reinterpret_cast<const float*>(big_tab + p.off_dct);
    const uint32_t* s_tasks = reinterpret_cast<const uint32_t*>(s_tab + p.off_tasks);
    const int32_t* s_tbeg = reinterpret_cast<const int32_t*>(s_tab + p.off_tbeg);
    const int2* s_qspec = reinterpret_cast<const int2*>(s_tab + p.off_qspec);
    // DCT reads n_filt4 rows; the pad rows stay zero.  Rows are stored in pairs, [q / 2][slot][q & 1], so that the DCT
    // fetches two bands of its slot per 128-bit load (one wavefront per quarter-warp row pair; + 0.5 % on the 512-clip batch)
    const int n_lq = (max(p.n_q, p.n_filt4) + 1) & ~1;
    auto lq = [&](int q, int sl) { return ((q >> 1) * geo::PPT + sl) * 2 + (q & 1); };
    // per team: log bands [n_lq / 2][slot][2] as packed (A, B) pairs, per-slot info (frame energies + output row); with the
    // fused all-gather also the finished rows [frame slot][col] and their row ids (int64)
    const int team_floats = team_smem_floats<geo::PPT>(p);
    f2* s_logq = reinterpret_cast<f2*>(reinterpret_cast<float*>(s_tab + staged_bytes) + team * team_floats);
    ulonglong2* s_info = reinterpret_cast<ulonglong2*>(s_logq + n_lq * geo::PPT);   // written by the FFT stage
    float* s_stage = reinterpret_cast<float*>(s_info + geo::PPT);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(s_tab + staged_bytes) + TEAMS * team_floats);
    float* xw = s_xch + warp * geo::XWARP;
    auto team_sync = [&]() {
        if constexpr (TEAMS == 1) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(kThreads) : "memory");
    };

    // Programmatic dependent launch: let the NEXT extract launch of the stream start filling SMs as soon as
    // this grid's CTAs retire (its loads / FFTs do not depend on us).  Every extract grid waits for its
    // predecessor (griddepcontrol.wait below) before its first global store, so output ordering is kept -- and
    // before its first load of anything an extract grid writes (streaming state), see below;
    // kernels launched without the attribute (everything else in the stream) still wait for full completion.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    bool deps_done = false;

    // ---- one-time per CTA -----------------------------------------------------------------------------------
    // tables: ONE TMA bulk copy, waited for just before the first pass 2 (overlaps the first loads + pass 1)
    if (threadIdx.x == 0) {
        mbar_init(s_bar, 1);
        bulk_g2s(s_tab, p.tables, (uint32_t)staged_bytes, s_bar);
    }
    // A streaming step READS what its predecessor wrote (carry, carry length, ring): it must not touch any of that
    // before the previous push has completed, so it waits here instead of before its first store.  The launch
    // latency and the table copy above still overlap the predecessor.
    if constexpr (!FAST && sizeof(InT) == 2) {
        if (p.stream_on) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            deps_done = true;
        }
    }
    // (no zero fill of the exchange area: every float of a valid pair row is rewritten each tile, and whatever an
    //  unused slot holds only reaches that slot's own, discarded, results)
    for (int i = tid; i < n_lq * geo::PPT; i += kThreads) s_logq[i] = pk(0.f, 0.f);
    if (tid == 0)           // the team's mailbox for drawn tiles (see the tile schedule below): its third tile is fixed
        *(reinterpret_cast<volatile uint32_t*>(s_bar + 1) + team) = (blockIdx.x * TEAMS + team) + 2 * (gridDim.x * TEAMS);

    // pass-2 role of this lane: pair g2 of the warp, column k1
    const int g2 = lane / R;
    const int k1 = lane % R;
    // power rows of this warp's pairs: skewed so that the bank phase's float4 column reads are conflict-free
    // (row of slot s starts at 4 * (s & 7) mod 32 words; a quarter warp reads 8 consecutive slots)
    float* pw = xw + 4 * ((geo::G * warp) & 7);

    // bank / epilogue role of this thread: pair slot `slot`, thread group `grp`
    const int slot = tid % geo::PPT;
    const int grp = tid / geo::PPT;
    const float* prow_slot;
    {
        const int sw = slot / geo::G;                // warp that produced this slot
        prow_slot = s_xch + sw * geo::XWARP + 4 * ((geo::G * sw) & 7) + (slot % geo::G) * geo::PROW2;
    }
    // partial-sum row at 64-byte unit u of this team's exchange area, this thread's slot
    f2* const part0 = reinterpret_cast<f2*>(s_xch) + slot;
    auto part = [&](uint32_t u) -> f2* { return part0 + 8 * u; };

    // How exactly-zero frames are recognised (they must produce exactly zero power, see the FFT stage):
    //  * fast int16 path: from the frame energy in the epilogue -- a non-zero int16 frame has raw energy >= 0.5
    //    while its partner can leak at most ~2e-3 into it, so energy < 0.25 means "all zero";
    //  * everything else (float input, generic loader): bitwise OR of the samples + warp vote.
    constexpr bool kEnergyZero = FAST && sizeof(InT) == 2 && !PRE;
    constexpr bool bit_detect = !kEnergyZero;
    const InT* __restrict__ in = reinterpret_cast<const InT*>(p.in);
    const uint32_t ppc = (uint32_t)p.pairs_per_clip;
    const uint32_t n_pairs = (uint32_t)p.n_pairs;
    bool tables_ready = false;

    // ---- where this warp is: its first pair of the current tile as (clip, pair inside the clip), advanced by a fixed
    //      (clips, pairs) step per tile -- one division per kernel instead of three per pair
    const uint32_t tile_stride = gridDim.x * TEAMS;          // = number of teams of the grid
    const uint32_t tile_first = blockIdx.x * TEAMS + team;
    // Tile schedule.  A team's first three tiles are fixed (tile_first + k * tile_stride); with p.tile_ctr the following
    // ones come from a global counter, one draw per executed tile and team, three tiles ahead of their use: teams do not
    // run at the same pace (SMs differ by a few percent), and with a fixed round robin the last team of a 13,229-clip
    // launch finished 48 us after the first one -- 24 us of idle time per team on average (tools/team_times.py).
    // Without a counter (streaming steps, generic loader, power output) the round robin continues.
    uint32_t* const tile_ctr = p.tile_ctr;
    auto where = [&](uint32_t t, uint32_t& gp, uint32_t& c, uint32_t& q) {      // first pair of this warp in tile t
        gp = t * geo::PPT + warp * geo::G;
        c = fast_div(gp, p.ppc_magic, p.ppc_shift);
        q = gp - c * ppc;
    };
    uint32_t gp0, clip0, q0;
    where(tile_first, gp0, clip0, q0);
    // position of pair g of the warp, given the position of its pair 0
    auto pair_of = [&](uint32_t c_in, uint32_t q_in, int g, uint32_t& c_out, uint32_t& q_out) {
        c_out = c_in;
        q_out = q_in + g;
        if constexpr (geo::G > 1) {
            while (q_out >= ppc) { q_out -= ppc; ++c_out; }
        }
    };
    // (clip, pair inside the clip) of an arbitrary global pair (streaming epilogue only)
    auto pair_pos = [&](uint32_t gp, uint32_t& c_out, uint32_t& q_out) {
        c_out = fast_div(gp, p.ppc_magic, p.ppc_shift);
        q_out = gp - c_out * ppc;
    };

    // fast path (window == n_fft, hop == n_fft/2): frames 2q and 2q+1 share half their samples;
    // raw[j] = x[n_fft*q + lane + 32 j], j < R + R/2.  Two rare cases take the predicated form of the loop:
    //  * frame B is absent (last pair of a clip with an odd frame count): its upper samples may lie behind the clip;
    //  * the pair touches the zeros in FRONT of a short clip (per-clip lengths with SCF_PAD_FRONT_ZERO,
    //    common/data_utils.py:77-80): element e of the padded clip is sample e - pad of the clip's data.
    typedef typename Raw<InT>::type RawT;
    RawT raw[geo::G][geo::NLOAD];
    auto load_pair = [&](uint32_t clip, uint32_t q, RawT (&dst)[geo::NLOAD]) {
        int pad = 0;
        if (p.lengths != nullptr) pad = p.clip_len - min(max(__ldg(p.lengths + clip), 0), p.clip_len);
        const int e0 = (int)q * geo::NFFT;                    // first element of the pair in the padded clip
        const InT* __restrict__ cb = in + (int64_t)clip * p.clip_stride;
        const bool b_absent = p.frames_odd != 0 && q + 1 == ppc;
        const int n_ld = b_absent ? R : geo::NLOAD;
        if (__builtin_expect(pad <= e0 && !b_absent, 1)) {
            const InT* __restrict__ src = cb + (e0 - pad) + lane;
#pragma unroll
            for (int j = 0; j < geo::NLOAD; ++j) dst[j] = ld_sample(src + 32 * j);
        } else {
            const int j0 = (pad - e0 - lane + 31) >> 5;       // first j with e0 + lane + 32 j >= pad
#pragma unroll
            for (int j = 0; j < geo::NLOAD; ++j)
                dst[j] = (j >= j0 && j < n_ld) ? ld_sample(cb + (e0 - pad + lane + 32 * j)) : (RawT)0;
        }
    };
    // PRE: prv[j] = the element in front of raw[j] (zero in front of the clip's first sample and inside the padding)
    auto load_prev = [&](uint32_t clip, uint32_t q, RawT (&dst)[geo::NLOAD]) {
        int pad = 0;
        if (p.lengths != nullptr) pad = p.clip_len - min(max(__ldg(p.lengths + clip), 0), p.clip_len);
        const int e0 = (int)q * geo::NFFT;
        const InT* __restrict__ cb = in + (int64_t)clip * p.clip_stride;
        const bool b_absent = p.frames_odd != 0 && q + 1 == ppc;
        const int n_ld = b_absent ? R : geo::NLOAD;
        if (__builtin_expect(pad <= e0 && !b_absent, 1)) {
            const InT* __restrict__ src = cb + (e0 - pad) + lane - 1;
            dst[0] = ld_sample_if(src, lane != 0 || e0 != pad);
#pragma unroll
            for (int j = 1; j < geo::NLOAD; ++j) dst[j] = ld_sample(src + 32 * j);
        } else {
            const int j0 = (pad - e0 - lane + 32) >> 5;       // first j with e0 + lane + 32 j - 1 >= pad
#pragma unroll
            for (int j = 0; j < geo::NLOAD; ++j)
                dst[j] = ld_sample_if(cb + (e0 - pad + lane + 32 * j - 1), j >= j0 && j < n_ld);
        }
    };
    // Streaming step: state carry-over of stream `clip` (listen.py:106-109), split over the warps that own its pairs:
    // the warp of pair 0 writes the new carry concat(carry, chunk)[k * hop:] and the step's counters, the warp of pair 1
    // (pair 0 when a step has a single pair) moves the surviving ring rows up by k -- the k new rows are written by
    // whoever computes them.  Everything is read from the `in` buffers and written to the `out` buffers, so the order
    // among warps, teams and launches' threads does not matter.  Called between the issue and the first use of the
    // pair's sample loads, with its own loads in flight 16 (carry) or 8 x 16 bytes (ring) at a time: one or two memory round trips, hidden behind the
    // samples' (the first version ran after the DCT in rounds of 8 loads per lane: 12.7 -> see DESIGN.md 4.2).
    auto stream_move = [&](uint32_t clip, uint32_t q) {
        if constexpr (!FAST && sizeof(InT) == 2) {
            const StreamStep& ss = p.stream;
            const bool do_carry = q == 0, do_ring = q == (ppc > 1 ? 1u : 0u);
            if (!do_carry && !do_ring) return;
            const int len_old = __ldg(ss.len_in + clip);
            const int len = len_old + ss.chunk_len;
            const int k = (len >= p.window) ? (len - p.window) / p.hop + 1 : 0;
            const int consumed = k * p.hop, keep = len - consumed;
            if (do_carry) {
                const int16_t* cin = ss.carry_in + (int64_t)clip * ss.carry_cap;
                const int16_t* ch = ss.chunks + (int64_t)clip * ss.chunk_len;
                int16_t* cout = ss.carry_out + (int64_t)clip * ss.carry_cap;
                for (int j0 = lane; j0 < keep; j0 += 32 * 16) {              // keep < window + hop: two rounds as a rule
                    int16_t v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int j = j0 + 32 * u, a = consumed + j;
                        v[u] = j < keep ? (a < len_old ? __ldg(cin + a) : __ldg(ch + (a - len_old))) : (int16_t)0;
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        if (j0 + 32 * u < keep) cout[j0 + 32 * u] = v[u];
                }
                if (lane == 0) {
                    ss.len_out[clip] = keep;
                    ss.n_new[clip] = k;
                    if (ss.n_new_copy != nullptr) ss.n_new_copy[clip] = k;
                }
            }
            if (do_ring) {
                const int cols = p.out_cols;
                const int kk = min(k, ss.ring_rows);
                const int n_old = (ss.ring_rows - kk) * cols;
                const int64_t r0 = (int64_t)clip * ss.ring_rows * cols;
                const float* src = ss.ring_in + r0 + kk * cols;
                float* dst = ss.ring_out + r0;
                float* cpy = ss.ring_copy != nullptr ? ss.ring_copy + (int64_t)clip * ss.ring_rows * ss.copy_pitch : nullptr;
                const bool vec = ((cols | ss.copy_pitch) & 3) == 0 &&
                                 ((reinterpret_cast<uintptr_t>(ss.ring_in) | reinterpret_cast<uintptr_t>(ss.ring_out) |
                                   reinterpret_cast<uintptr_t>(ss.ring_copy)) & 15) == 0;
                if (vec) {                                                    // rows are multiples of 16 bytes
                    const int n4 = n_old >> 2, c4 = cols >> 2;
                    for (int j0 = lane; j0 < n4; j0 += 32 * 8) {
                        float4 v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            v[u] = j0 + 32 * u < n4 ? __ldg(reinterpret_cast<const float4*>(src) + j0 + 32 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int j = j0 + 32 * u;
                            if (j < n4) {
                                reinterpret_cast<float4*>(dst)[j] = v[u];
                                if (cpy != nullptr) {
                                    const int rr = j / c4;
                                    reinterpret_cast<float4*>(cpy + (int64_t)rr * ss.copy_pitch)[j - rr * c4] = v[u];
                                }
                            }
                        }
                    }
                } else {
                    for (int j0 = lane; j0 < n_old; j0 += 32 * 8) {
                        float v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) v[u] = j0 + 32 * u < n_old ? __ldg(src + j0 + 32 * u) : 0.f;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int j = j0 + 32 * u;
                            if (j < n_old) {
                                dst[j] = v[u];
                                if (cpy != nullptr) {
                                    const int rr = j / cols;
                                    cpy[(int64_t)rr * ss.copy_pitch + (j - rr * cols)] = v[u];
                                }
                            }
                        }
                    }
                }
            }
        }
    };
    // classic variant: the samples of the NEXT tile are fetched into registers while the bank / log / DCT phases of
    // the current tile run (those need few registers), so the FFT stage never waits on HBM
    // (float input keeps 48 full registers busy that way and spills: it loads at the point of use instead;
    //  the 24-warp CTAs have no registers to spare and prefetch into L2 instead)
    constexpr bool kPrefetch = FAST && sizeof(InT) == 2 && !DENSE && !PRE;   // (with PRE: 968 bytes of spills)
    auto prefetch = [&](uint32_t gp, uint32_t c_in, uint32_t q_in) {
        if constexpr (kPrefetch) {
#pragma unroll
            for (int g = 0; g < geo::G; ++g) {
                if (gp + g < n_pairs) {
                    uint32_t clip, q;
                    pair_of(c_in, q_in, g, clip, q);
                    load_pair(clip, q, raw[g]);
                }
            }
        }
    };
    if (tile_first < n_tiles) prefetch(gp0, clip0, q0);
    __syncthreads();          // mbarrier init + zeroed s_logq pad rows visible (the only CTA-wide barrier)
#ifdef SCF_STAGGER_NS        // experiment: teams of a CTA start a third of a tile apart instead of in phase
    if constexpr (TEAMS > 1) {
        for (int t = 0; t < team; ++t) __nanosleep(SCF_STAGGER_NS);
    }
#endif
#ifdef SCF_ABL_TEAMS         // scaling experiment (profiles/r02_regular_bank_tma_staging.log): only the first n teams of a CTA work
    if (team >= SCF_ABL_TEAMS) return;
#endif

#ifdef SCF_DEBUG_TIMES
    if (tid == 0 && tile_first < 2048) g_dbg_times[tile_first] = dbg_now();
#endif
    uint32_t tile = tile_first, tile_n = tile_first + tile_stride, incoming = tile_first + 2 * tile_stride;
    // s_next: the team's mailbox for drawn tiles.  Thread 0 draws while it has nothing else to do (it holds no DCT
    // coefficient; log-bank output: behind the log phase) and writes the tile behind the iteration's last barrier but one;
    // everybody reads the mailbox right behind the NEXT tile's first barrier -- two barriers before the next write.
    volatile uint32_t* const s_next = reinterpret_cast<volatile uint32_t*>(s_bar + 1) + team;
    auto draw = [&]() {          // every executed tile draws once: the launch's last number resets the counter
        const uint32_t v = atomicAdd(tile_ctr, 1u);
        if (v == n_tiles - 1) *tile_ctr = 0u;
        *s_next = 3 * tile_stride + v;
    };
    auto advance = [&](uint32_t gp_n, uint32_t clip_n, uint32_t q_n) {
        tile = tile_n;
        tile_n = incoming;
        gp0 = gp_n;
        clip0 = clip_n;
        q0 = q_n;
    };
#ifdef SCF_DEBUG_TIMES
    int dbg_i = 0;
    auto dbg_clock = []() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; };
    long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dbg_t = dbg_clock();
    // phases: 0 FFT stage, 1 wait at the first barrier, 2 bank, 3 wait at the second barrier, 4 log, 5 wait at the third, 6 DCT + stores
#define DBG_MARK(ph) do { const long long now_ = dbg_clock(); dbg_acc[ph] += now_ - dbg_t; dbg_t = now_; } while (0)
    // behind a barrier: BAR.SYNC.DEFER_BLOCKING lets the warp run on until its next memory access, so read something first
#define DBG_MARK_SYNCED(ph) do { int probe_; asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(probe_) : "r"(smem_u32(s_bar)) : "memory"); \
        asm volatile("" :: "r"(probe_) : "memory"); DBG_MARK(ph); } while (0)
#else
#define DBG_MARK_SYNCED(ph) do { } while (0)
#define DBG_MARK(ph) do { } while (0)
#endif
    while (tile < n_tiles) {
        const uint32_t pair0 = tile * geo::PPT;
        // where this warp will be in its next tile: used for the prefetch now, and as the position then
        uint32_t gp_n, clip_n, q_n;
        where(tile_n, gp_n, clip_n, q_n);
        const bool more = tile_n < n_tiles;

        // =========================== FFT stage (per warp) =======================================
        if (gp0 < n_pairs) {
            // ---- pass 1: lane = n2; R-point FFT over n1 of z[n2 + 32 n1], z = A + iB ----------------
            // A frame whose samples are ALL exactly zero (front padding, digital silence) must come out as exactly
            // zero power: the two-for-one separation below would otherwise leak ~4e-15 of its partner's energy into
            // it, which is visible above the eps floor of the log.  zero_mask: bit 2g = frame A of pair g, 2g+1 = B.
            uint32_t zero_mask = 0;
#pragma unroll
            for (int g = 0; g < geo::G; ++g) {
                if (gp0 + g < n_pairs) {
                    uint32_t clip, q;
                    pair_of(clip0, q0, g, clip, q);
                    f2 x[R];                                   // (re, im) = (frame A sample, frame B sample), packed
                    uint32_t nz_a = 0, nz_b = 0;
                    int n_frames = p.frames_per_clip;          // rows this clip produces
                    if constexpr (FAST) {
                        const bool b_absent = p.frames_odd != 0 && q + 1 == ppc;          // odd frame count: last pair
                        if constexpr (!kPrefetch) {
                            load_pair(clip, q, raw[g]);
                            // pull the samples this warp needs in its NEXT tile into L2 (one 128-byte line per lane)
                            if (gp_n + g < n_pairs) {
                                uint32_t clipn, qn;
                                pair_of(clip_n, q_n, g, clipn, qn);
                                const InT* nsrc = in + (int64_t)clipn * p.clip_stride + qn * geo::NFFT;
                                if (lane * 128 < (int)(geo::NLOAD * 32 * sizeof(InT)) + 128)
                                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(nsrc) + lane * 128));
                            }
                        }
                        if constexpr (PRE) {
                            // e[n] = x[n] - alpha * x[n-1] for the 1.5 frames of the pair, then both frames times the
                            // window value of their common position n2 + 32 n1 (alpha may be 0, the window all ones)
                            float e[geo::NLOAD];
                            const float ma = -p.preemph;
                            const float* __restrict__ wl = p.win + lane;
                            // x of position n1 as soon as e[n1 + R/2] exists: never more than 64 live values
                            auto emit = [&](int n1) {
                                const int i = scf_bitrev(n1, geo::LOG2R);
                                x[i] = mul2(pk(e[n1], e[n1 + R / 2]), bc(__ldg(wl + 32 * n1)));
                                nz_a |= __float_as_uint(lo(x[i]));
                                nz_b |= __float_as_uint(hi(x[i]));
                            };
                            RawT prv[geo::NLOAD];
                            load_prev(clip, q, prv);
#pragma unroll
                            for (int j = 0; j < geo::NLOAD; ++j) {
                                e[j] = fmaf(ma, to_f32(prv[j]), to_f32(raw[g][j]));
                                if (j >= R / 2) emit(j - R / 2);
                            }
                            nz_a &= 0x7fffffffu;                                         // (-0.0f counts as zero)
                            nz_b = b_absent ? 0u : (nz_b & 0x7fffffffu);
                        } else {
                            if (bit_detect) {
                                uint32_t o0 = 0, o1 = 0, o2 = 0;
#pragma unroll
                                for (int j = 0; j < R / 2; ++j) {
                                    o0 |= nz_bits(raw[g][j]);
                                    o1 |= nz_bits(raw[g][j + R / 2]);
                                    o2 |= nz_bits(raw[g][j + R]);
                                }
                                nz_a = o0 | o1;
                                nz_b = b_absent ? 0u : (o1 | o2);
                            }
#pragma unroll
                            for (int i = 0; i < R; ++i) {
                                const int n1 = scf_bitrev(i, geo::LOG2R);
                                x[i] = pk(to_f32(raw[g][n1]), to_f32(raw[g][n1 + R / 2]));
                            }
                        }
                        if (__builtin_expect(b_absent, 0)) {   // frame B does not exist: imaginary parts are zero
#pragma unroll
                            for (int i = 0; i < R; ++i) x[i] = pk(lo(x[i]), 0.f);
                        }
                    } else {
                        const InT* __restrict__ cb = in + (int64_t)clip * p.clip_stride;
                        const ClipGeom cg = clip_geom(p, clip);
                        n_frames = cg.n_frames;
                        float xr[R], xi[R];
                        bool streaming = false;
                        if constexpr (sizeof(InT) == 2) streaming = p.stream_on != 0;
                        if (streaming) {
                            if constexpr (sizeof(InT) == 2) {
                                const int split = cg.len - p.stream.chunk_len;          // carry, then the new chunk
                                const InT* tail = reinterpret_cast<const InT*>(p.stream.chunks) + (int64_t)clip * p.stream.chunk_len;
                                load_pair_generic<R, InT, true>(p, cb, tail, split, cg, 2 * (int)q, lane, xr, xi, nz_a, nz_b,
                                                                [&]() { stream_move(clip, q); });
                            }
                        } else {
                            load_pair_generic<R, InT, false>(p, cb, cb, 0, cg, 2 * (int)q, lane, xr, xi, nz_a, nz_b, []() {});
                        }
#pragma unroll
                        for (int i = 0; i < R; ++i) x[i] = pk(xr[i], xi[i]);
                    }
                    if (bit_detect) {
                        if (!__any_sync(0xffffffffu, nz_a != 0)) zero_mask |= 1u << (2 * g);
                        if (!__any_sync(0xffffffffu, nz_b != 0)) zero_mask |= 1u << (2 * g + 1);
                    }
                    // where the pair's rows go (the epilogue threads read this instead of redoing the index math):
                    // 4 * (row of frame A + 1) + 2 * (frame A is stored) + (frame B is stored), or -1
                    if (lane == 0) {
                        const int f = 2 * (int)q;
                        long long row = (long long)clip * p.frames_per_clip + f;
                        bool st_a = f < n_frames, st_b = f + 1 < n_frames;
                        if constexpr (!FAST) {
                            if (p.stream_on) {        // the k new frames are the k last ring rows (listen.py:107-109)
                                const int r = p.stream.ring_rows - n_frames + f;
                                row = (long long)clip * p.stream.ring_rows + r;
                                st_a = st_a && r >= 0;
                                st_b = st_b && r + 1 >= 0;
                            }
                        }
                        const long long code = (st_a || st_b) ? 4 * (row + 1) + (st_a ? 2 : 0) + (st_b ? 1 : 0) : -1;
                        s_info[warp * geo::G + g].y = (unsigned long long)code;
                    }
                    fft_r<R>(x);
                    f2* row = reinterpret_cast<f2*>(xw + g * geo::XPAIR + lane * geo::XROW);
#pragma unroll
                    for (int k = 0; k < R; ++k) row[k] = x[k];
                } else if (lane == 0) {
                    s_info[warp * geo::G + g].y = ~0ull;
                }
            }
            __syncwarp();
            if (!tables_ready) {          // first tile only: the twiddles arrive with the table blob
                mbar_wait(s_bar, 0);
                tables_ready = true;
            }

            // ---- pass 2: lane = (pair g2, column k1); twiddle, 32-point FFT over n2 -> Z[k1 + R k2] --
            // (the inter-pass twiddles are folded into the first butterfly stage, see gen_fft.py)
            f2 y[32];
            {
                const float* col = xw + g2 * geo::XPAIR + 2 * k1;
                f2 z[32], c[32];
                // twiddles W^(k1 n2): n2 < 8 come from the table, the rest is derived with packed complex multiplies by
                // W^(8 k1) and W^(16 k1) -- arithmetic is cheap here, shared-memory wavefronts are not
                auto ld_tw = [&](int i) {
                    const float4 t = kTwInL1 ? __ldg(s_tw4 + i) : s_tw4[i];
                    return make_ulonglong2(pk(t.x, t.y), pk(t.z, t.w));
                };
                const ulonglong2 wq = ld_tw(4 * 32 + lane);                          // (W^(8 k1), W^(16 k1))
#pragma unroll
                for (int n2 = 0; n2 < 8; n2 += 2) {
                    const ulonglong2 t = ld_tw((n2 / 2) * 32 + lane);
                    c[n2] = t.x;
                    c[n2 + 1] = t.y;
                }
#pragma unroll
                for (int n2 = 0; n2 < 8; ++n2) c[n2 + 8] = fma2(mul_i(c[n2]), bc(hi(wq.x)), mul2(c[n2], bc(lo(wq.x))));
#pragma unroll
                for (int n2 = 0; n2 < 16; ++n2) c[n2 + 16] = fma2(mul_i(c[n2]), bc(hi(wq.y)), mul2(c[n2], bc(lo(wq.y))));
#pragma unroll
                for (int n2 = 0; n2 < 32; ++n2) z[n2] = *reinterpret_cast<const f2*>(col + n2 * geo::XROW);
                fft32_p2_tw(z, c, y);
            }
            __syncwarp();    // every lane has read its column: the region may now be reused

            // ---- separate the two frames ------------------------------------------------------------
            // Z[N - k] for k = k1 + R k2 sits in lane (R - k1) % R of the same pair, register 31 - k2 (k1 == 0: the
            // lane's own register 32 - k2): a fixed lane permutation, done with shuffles -- one SHFL moves what takes
            // an STS plus an LDS wavefront through shared memory (tools/microbench/shfl_vs_lds.cu: 33 vs 65 cycles).
            // the loads of the next tile's samples are issued here: they fill the wait for the mirror values
            if (more) prefetch(gp_n, clip_n, q_n);
            float* prow = pw + g2 * geo::PROW2;
            const int src_lane = (lane & ~(R - 1)) | ((R - k1) & (R - 1));
            f2 esum = pk(0.f, 0.f);
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
                f2 m = __shfl_sync(0xffffffffu, y[31 - k2], src_lane);               // Z[N - k]
                if (k1 == 0) m = y[(32 - k2) & 31];                                  // (Z[N] == Z[0])
                // 2A[k] = Z + conj(Zm) = (yr + mr, yi - mi);  2B[k] = (yi + mi, mr - yr)
                const f2 u1 = add2(y[k2], m);                                        // (yr + mr, yi + mi) = (Re 2A, Re 2B)
                const f2 u2 = add2(mul_mi(y[k2]), mul_i(m));                         // (yi - mi, mr - yr) = (Im 2A, Im 2B)
                const f2 pp = fma2(u2, u2, mul2(u1, u1));                            // (|2A|^2, |2B|^2)
                *reinterpret_cast<f2*>(prow + 2 * (k1 + R * k2)) = pp;
                esum = add2(esum, pp);
            }
            if (k1 == 0) {                                                           // bin n_fft/2 mirrors onto itself
                const f2 pp = mul2(mul2(y[16], y[16]), bc(4.f));
                // ... and bin n_fft/2 + 1 is the row's zero pad (tasks may read it, with zero weights)
                *reinterpret_cast<ulonglong2*>(prow + 2 * geo::NB) = make_ulonglong2(pp, pk(0.f, 0.f));
                esum = add2(esum, pp);
            }
            // frame energies (sum of the power row; c0 of the cepstrum, bark_feature.py:173): the values are in
            // registers right now, so a warp reduction is cheaper than a 513-tap pseudo-filter in the bank phase
#pragma unroll
            for (int m = R / 2; m >= 1; m >>= 1) esum = add2(esum, __shfl_xor_sync(0xffffffffu, esum, m));
            if (k1 == 0) {
                const bool za = (zero_mask >> (2 * g2)) & 1u, zb = (zero_mask >> (2 * g2 + 1)) & 1u;
                s_info[warp * geo::G + g2].x = pk(za ? 0.f : lo(esum) * p.power_scale, zb ? 0.f : hi(esum) * p.power_scale);
            }
            if (__builtin_expect(zero_mask != 0, 0)) {                     // rare: exact zeros for silent frames
                const bool za = (zero_mask >> (2 * g2)) & 1u, zb = (zero_mask >> (2 * g2 + 1)) & 1u;
                __syncwarp();
#pragma unroll 1
                for (int k = k1; k <= geo::NB; k += R) {                   // a real loop: keep the hot path short
                    if (za) prow[2 * k] = 0.f;
                    if (zb) prow[2 * k + 1] = 0.f;
                }
            }
        } else {
            if (!tables_ready) {
                mbar_wait(s_bar, 0);
                tables_ready = true;
            }
            if (lane < geo::G) s_info[warp * geo::G + lane].y = ~0ull;          // no pair, no rows
            if (more) prefetch(gp_n, clip_n, q_n);
        }
        if (!deps_done) {         // before this grid's first global store
            asm volatile("griddepcontrol.wait;" ::: "memory");
            deps_done = true;
        }
        DBG_MARK(0);
        team_sync();
        DBG_MARK_SYNCED(1);
        // (through a shuffle: the compiler then knows the value is warp-uniform and keeps the schedule in uniform registers)
        incoming = tile_ctr != nullptr ? __shfl_sync(0xffffffffu, *s_next, 0) : tile_n + tile_stride;

        // =========================== where this thread's pair goes ==============================
        const ulonglong2 info = s_info[slot];                    // (frame energies, row code) from the FFT stage
        const long long row_code = (long long)info.y;
        const int64_t row_a = (row_code >> 2) - 1;    // output row of frame A; frame B is the next row
        const bool has_a = row_code >= 0 && (row_code & 2), has_b = row_code >= 0 && (row_code & 1);

        if (p.out_kind == SCF_OUT_POWER) {
            // power_spec(): rows straight out of shared memory, coalesced along the bins;
            // warp w copies frame slots w, w+8, ...; the row index is recomputed per slot (warp-uniform)
            for (int s = warp; s < 2 * geo::PPT; s += kWarps) {
                const long long code = (long long)s_info[s >> 1].y;
                if (code < 0 || !(code & ((s & 1) ? 1 : 2))) continue;
                const int64_t row = (code >> 2) - 1 + (s & 1);
                const f2 e2 = s_info[s >> 1].x;
                const bool silent = kEnergyZero && ((s & 1) ? hi(e2) : lo(e2)) < p.zero_energy;
                const int sw = (s >> 1) / geo::G;
                const float* src = s_xch + sw * geo::XWARP + 4 * ((geo::G * sw) & 7) + ((s >> 1) % geo::G) * geo::PROW2 + (s & 1);
                float* dst = p.out + row * p.out_pitch;
                for (int k = lane; k <= geo::NB; k += 32) dst[k] = silent ? 0.f : src[2 * k] * p.power_scale;
            }
            team_sync();
            advance(gp_n, clip_n, q_n);
            continue;
        }

        // =========================== bank stage ================================================
        // tasks are runs of 8 bins applied to two filters at once (see scfeat_internal.h); consecutive tasks of a run
        // accumulate in registers and the task flagged `last` stores the run's two partial sums.  Two tasks are in
        // flight per iteration and the next pair of task words is fetched one iteration ahead, so the shared-memory
        // round trips overlap the FMAs.
        {
            const int2 be = reinterpret_cast<const int2*>(s_tbeg)[grp];      // this group's tasks: [be.x, be.y)
            f2 acc_a = pk(0.f, 0.f), acc_b = pk(0.f, 0.f);
            auto load_task = [&](uint32_t word, int t, ulonglong2 (&x)[4], float4 (&w)[4]) {
                const ulonglong2* px = reinterpret_cast<const ulonglong2*>(prow_slot) + (word & 0xffu);
                const float4* ww = s_wts4 + 4 * t;
#pragma unroll
                for (int i = 0; i < 4; ++i) { x[i] = px[i]; w[i] = ww[i]; }
            };
            auto run_task = [&](uint32_t word, const ulonglong2 (&x)[4], const float4 (&w)[4]) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc_a = fma2(x[i].x, bc(w[i].x), acc_a);
                    acc_b = fma2(x[i].x, bc(w[i].y), acc_b);
                    acc_a = fma2(x[i].y, bc(w[i].z), acc_a);
                    acc_b = fma2(x[i].y, bc(w[i].w), acc_b);
                }
                if (word & 0x80000000u) {
                    *part((word >> 8) & 0x7ffu) = acc_a;
                    *part((word >> 19) & 0x7ffu) = acc_b;
                    acc_a = acc_b = pk(0.f, 0.f);
                }
            };
            int t = be.x;
            uint2 tk = make_uint2(0u, 0u);
            if (t < be.y) tk.x = s_tasks[t];
            if (t + 1 < be.y) tk.y = s_tasks[t + 1];
            for (; t + 1 < be.y; t += 2) {
                ulonglong2 x[4], y[4];
                float4 wx[4], wy[4];
                load_task(tk.x, t, x, wx);
                load_task(tk.y, t + 1, y, wy);
                const uint2 cur = tk;
                if (t + 2 < be.y) tk.x = s_tasks[t + 2];
                if (t + 3 < be.y) tk.y = s_tasks[t + 3];
                run_task(cur.x, x, wx);
                run_task(cur.y, y, wy);
            }
            if (t < be.y) {                               // odd list: the last task on its own
                ulonglong2 x[4];
                float4 wx[4];
                load_task(tk.x, t, x, wx);
                run_task(tk.x, x, wx);
            }
        }
        DBG_MARK(2);
        team_sync();
        DBG_MARK_SYNCED(3);

        // =========================== log ========================================================
        const f2 frame_energy = info.x;
        bool silent_a = false, silent_b = false;
        if constexpr (kEnergyZero) {
            silent_a = lo(frame_energy) < p.zero_energy;
            silent_b = hi(frame_energy) < p.zero_energy;
        }
        for (int q = grp; q < p.n_q; q += geo::NGRP) {            // n_q = n_filt (+1: the energy, cepstrum only)
            f2 v = frame_energy;
            if (q < p.n_filt) {
                const int2 qs = s_qspec[q];
                v = pk(0.f, 0.f);
                const f2* pr = part(qs.x);
                for (int j = 0; j < qs.y; ++j) v = add2(v, pr[j * (geo::PPT)]);
            }
            // lg2.approx * ln2: |err| ~ 1e-6, budget 1e-3 (.ftz: the argument is >= eps, so the denormal range check and
            // rescaling that __logf wraps around the MUFU are dead weight -- 8 instructions per thread on this phase's path)
            const float la = log_ge_eps(fmaxf(silent_a ? 0.f : lo(v), SCF_EPS));
            const float lb = log_ge_eps(fmaxf(silent_b ? 0.f : hi(v), SCF_EPS));
            if (p.out_kind == SCF_OUT_LOG_BANK) {
                if (p.n_peers != 0) {                          // staged for the coalesced peer stores below
                    s_stage[(2 * slot) * p.out_cols + q] = la;
                    s_stage[(2 * slot + 1) * p.out_cols + q] = lb;
                } else {
                    if (has_a) p.out[row_a * p.out_pitch + q] = la;
                    if (has_b) p.out[(row_a + 1) * p.out_pitch + q] = lb;
                    if constexpr (!FAST) {
                        if (p.stream_on && p.stream.ring_copy != nullptr) {
                            if (has_a) p.stream.ring_copy[row_a * p.stream.copy_pitch + q] = la;
                            if (has_b) p.stream.ring_copy[(row_a + 1) * p.stream.copy_pitch + q] = lb;
                        }
                    }
                }
            } else {
                s_logq[lq(q, slot)] = pk(la, lb);
            }
        }
        // Fused all-gather: the finished rows of this tile sit contiguously in shared memory [frame slot][col]; every
        // thread pushes consecutive floats, so each warp writes whole 128-byte segments to every peer's cache over
        // NVLink (8-byte scattered stores per thread were 3x slower than a separate NCCL all-gather at 8 GPUs).
        auto push_to_peers = [&]() {
            int64_t* s_rows = reinterpret_cast<int64_t*>(s_stage + 2 * geo::PPT * p.out_cols);
            if (grp == 0) {
                s_rows[2 * slot] = has_a ? row_a : -1;
                s_rows[2 * slot + 1] = has_b ? row_a + 1 : -1;
            }
            team_sync();
            const int n = 2 * geo::PPT * p.out_cols;
            if ((p.out_cols & 3) == 0) {              // rows are multiples of 16 bytes: 512-byte requests per warp
                const int c4 = p.out_cols >> 2;
                for (int i = tid; i < (n >> 2); i += kThreads) {
                    const int sl = i / c4;
                    const int64_t row = s_rows[sl];
                    if (row < 0) continue;
                    const float4 v = reinterpret_cast<const float4*>(s_stage)[i];
                    const int64_t off = (p.peer_row0 + row) * c4 + (i - sl * c4);
                    if (p.peer_multicast) {       // one store to the multicast address: the switch replicates it
                        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(
                                         reinterpret_cast<float4*>(p.peer_out[0]) + off),
                                     "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                                     : "memory");
                    } else {
                        for (int r = 0; r < p.n_peers; ++r) reinterpret_cast<float4*>(p.peer_out[r])[off] = v;
                    }
                }
            } else {
                for (int i = tid; i < n; i += kThreads) {
                    const int sl = i / p.out_cols;
                    const int64_t row = s_rows[sl];
                    if (row < 0) continue;
                    const float v = s_stage[i];
                    const int64_t off = (p.peer_row0 + row) * p.out_cols + (i - sl * p.out_cols);
                    if (p.peer_multicast) {
                        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.peer_out[0] + off), "f"(v) : "memory");
                    } else {
                        for (int r = 0; r < p.n_peers; ++r) p.peer_out[r][off] = v;
                    }
                }
            }
        };
        if (p.out_kind == SCF_OUT_LOG_BANK) {
            if (p.n_peers != 0) push_to_peers();
            if (tile_ctr != nullptr && tid == 0) draw();
            team_sync();          // the partial-sum rows live in the exchange area: the next tile's pass 1 rewrites them
            advance(gp_n, clip_n, q_n);
            continue;
        }
        DBG_MARK(4);
        team_sync();
        DBG_MARK_SYNCED(5);
        if (tile_ctr != nullptr && tid == 0) draw();

        // =========================== DCT-II, c0 := log energy ===================================
        // one coefficient of both frames per thread (packed); threads of a quarter warp share the DCT row
        // (coefficients are dealt to the thread groups from the top: the warps that were idle in the log phase -- the
        //  highest groups -- work here, so that no warp runs both short phases and arrives late at the next tile)
        for (int c = geo::NGRP - 1 - grp; c < p.n_out; c += geo::NGRP) {
            const float4* d4 = reinterpret_cast<const float4*>(s_dct + c * p.n_filt4);
            f2 a0 = pk(0.f, 0.f), a1 = pk(0.f, 0.f);
            for (int m = 0; m < p.n_filt4; m += 4) {
                // rows >= n_filt carry zero DCT weights (energy row is finite, pad rows are zero)
                const ulonglong2 l01 = *reinterpret_cast<const ulonglong2*>(s_logq + lq(m, slot));
                const ulonglong2 l23 = *reinterpret_cast<const ulonglong2*>(s_logq + lq(m + 2, slot));
                const f2 l0 = l01.x, l1 = l01.y, l2 = l23.x, l3 = l23.y;
                const float4 d = kTwInL1 ? __ldg(d4 + (m >> 2)) : d4[m >> 2];
                a0 = fma2(l0, bc(d.x), a0);
                a1 = fma2(l1, bc(d.y), a1);
                a0 = fma2(l2, bc(d.z), a0);
                a1 = fma2(l3, bc(d.w), a1);
            }
            f2 v = add2(a0, a1);
            if (c == 0) v = s_logq[lq(p.n_filt, slot)];
            if (p.n_peers != 0) {
                s_stage[(2 * slot) * p.out_cols + c] = lo(v);
                s_stage[(2 * slot + 1) * p.out_cols + c] = hi(v);
            } else {
                float* o = p.out + row_a * p.out_pitch + c;
                if (has_a) o[0] = lo(v);
                if (has_b) o[p.out_pitch] = hi(v);
                if constexpr (!FAST) {
                    if (p.stream_on && p.stream.ring_copy != nullptr) {
                        float* o2 = p.stream.ring_copy + row_a * p.stream.copy_pitch + c;
                        if (has_a) o2[0] = lo(v);
                        if (has_b) o2[p.stream.copy_pitch] = hi(v);
                    }
                }
            }
        }
        if (p.n_peers != 0) push_to_peers();
        // no barrier needed here: the next tile's FFT stage touches only the exchange area, which no thread reads
        // after the log phase; s_logq / s_info / s_stage are rewritten only behind later barriers.
        advance(gp_n, clip_n, q_n);
        DBG_MARK(6);
#ifdef SCF_DEBUG_TIMES
        if (tid == 0 && tile_first < 2048 && dbg_i < 6) g_dbg_times[(2 + dbg_i++) * 2048 + tile_first] = dbg_now();
#endif
    }
#ifdef SCF_DEBUG_TIMES
    if ((tid == 0 || tid == 224) && tile_first < 2048)
        for (int ph = 0; ph < 8; ++ph) g_dbg_phase[((tid ? 1 : 0) * 8 + ph) * 2048 + tile_first] = (unsigned long long)dbg_acc[ph];
#endif
#ifdef SCF_DEBUG_TIMES
    if (tid == 0 && tile_first < 2048) g_dbg_times[2048 + tile_first] = dbg_now();
#endif
}

// ------------------------------------------------------------------------------------------------
template <int R, int TEAMS, bool DENSE>
static size_t smem_bytes_rt(const KParams& p)
{
    using geo = Geo<R>;
    size_t b = (size_t)TEAMS * kWarps * geo::XWARP * 4 + (size_t)((DENSE && TEAMS == 1) ? p.table_small_bytes : p.table_bytes) +
               (size_t)TEAMS * team_smem_floats<geo::PPT>(p) * 4 + 32;      // mbarrier (8 bytes) + one next-tile word per team
    return (b + 15) & ~(size_t)15;
}

constexpr int64_t kTeamsMinPairs = 30720;      // measured crossover vs the 3-CTA variant: ~2000 one-second clips (tools/sweep.py)
constexpr size_t kSmemPerSm = 233472, kSmemReserve = 1024, kSmemMaxBlock = 232448;

// Kernel variant for a launch (see the template's comment): 0 = classic, 1 = dense with 3 CTAs per SM (n_fft = 1024
// only: the smaller transforms' exchange rows carry relatively more padding and do not fit three times),
// 3 = dense with one three-team CTA per SM.  SCFEAT_VARIANT=0|1|3 forces one (tuning / A-B runs).
template <int R>
static int variant_for_r(const KParams& p)
{
    static const int forced = [] { const char* e = getenv("SCFEAT_VARIANT"); return e ? atoi(e) : -1; }();
    if (forced == 0 || !p.fast_path || p.fast_pre) return 0;  // (the generic loader and the fused front end need the classic register budget)
    const bool fits3 = smem_bytes_rt<R, 3, true>(p) <= kSmemMaxBlock;
    const bool fits1 = R == 32 && 3 * (smem_bytes_rt<R, 1, true>(p) + kSmemReserve) <= kSmemPerSm;
    if (forced == 3 && fits3) return 3;
    if (forced == 1 && fits1) return 1;
    if (fits3 && p.n_pairs >= kTeamsMinPairs) return 3;
    return fits1 ? 1 : 0;
}

int variant_for(int r, const KParams& p)
{
    return r == 32 ? variant_for_r<32>(p) : r == 16 ? variant_for_r<16>(p) : variant_for_r<8>(p);
}

template <int R>
static size_t smem_bytes_r(const KParams& p)
{
    const int v = variant_for_r<R>(p);
    if (v == 3) return smem_bytes_rt<R, 3, true>(p);
    if constexpr (R == 32) {
        if (v == 1) return smem_bytes_rt<R, 1, true>(p);
    }
    return smem_bytes_rt<R, 1, false>(p);
}

size_t extract_smem_bytes(int r, const KParams& p)
{
    return r == 32 ? smem_bytes_r<32>(p) : r == 16 ? smem_bytes_r<16>(p) : smem_bytes_r<8>(p);
}

size_t extract_smem_limit(int r, const KParams& p)
{
    const int v = variant_for(r, p);
    return v == 3 ? kSmemMaxBlock : v == 1 ? kSmemPerSm / 3 - kSmemReserve : kSmemPerSm / kCtasPerSm - kSmemReserve;
}

int pairs_per_tile(int r) { return kWarps * (32 / r); }
int bank_groups(int r) { return kThreads / (kWarps * (32 / r)); }

template <int R, typename InT, bool FAST, int TEAMS, bool DENSE, bool PRE>
static cudaError_t launch_one(const KParams& p, int64_t n_tiles, int num_sms, cudaStream_t st, size_t smem)
{
    auto kern = extract_kernel<R, InT, FAST, TEAMS, DENSE, PRE>;
    // per device: the attribute call costs microseconds per launch, so it is made once per size (plans are shared
    // between threads: the bookkeeping is atomic, and setting the attribute twice is harmless)
    static std::atomic<size_t> configured[16];
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > configured[dev & 15].load(std::memory_order_acquire)) {
        static std::mutex mu;
        std::lock_guard<std::mutex> lock(mu);
        if (smem > configured[dev & 15].load(std::memory_order_relaxed)) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            if (DENSE)   // ask for the largest shared-memory carve-out so that three CTAs (or the big one) fit
                cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            configured[dev & 15].store(smem, std::memory_order_release);
        }
    }
    const int64_t ctas_per_sm = TEAMS == 3 ? 1 : (DENSE ? 3 : kCtasPerSm);
    int64_t grid = (int64_t)num_sms * ctas_per_sm;
    const int64_t need = (n_tiles + TEAMS - 1) / TEAMS;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads * TEAMS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p, (uint32_t)n_tiles);
    if (le != cudaSuccess) return le;
    count_launch(1);
    return cudaGetLastError();
}

// fast geometry (the params.json kernels), with or without the fused pre-emphasis / window loader
template <int R, typename InT, int TEAMS, bool DENSE>
static cudaError_t launch_fast(const KParams& p, int64_t n_tiles, int num_sms, cudaStream_t st, size_t smem)
{
    return p.fast_pre ? launch_one<R, InT, true, TEAMS, DENSE, true>(p, n_tiles, num_sms, st, smem)
                      : launch_one<R, InT, true, TEAMS, DENSE, false>(p, n_tiles, num_sms, st, smem);
}

template <int R, int TEAMS, bool DENSE>
static cudaError_t launch_r(bool is_f32, bool fast, const KParams& p, int64_t n_tiles, int num_sms, cudaStream_t st,
                            size_t smem)
{
    if constexpr (DENSE) {        // fast path without a front end only (variant_for)
        return is_f32 ? launch_one<R, float, true, TEAMS, DENSE, false>(p, n_tiles, num_sms, st, smem)
                      : launch_one<R, int16_t, true, TEAMS, DENSE, false>(p, n_tiles, num_sms, st, smem);
    } else {
        if (is_f32) {
            return fast ? launch_fast<R, float, TEAMS, DENSE>(p, n_tiles, num_sms, st, smem)
                        : launch_one<R, float, false, TEAMS, DENSE, false>(p, n_tiles, num_sms, st, smem);
        }
        return fast ? launch_fast<R, int16_t, TEAMS, DENSE>(p, n_tiles, num_sms, st, smem)
                    : launch_one<R, int16_t, false, TEAMS, DENSE, false>(p, n_tiles, num_sms, st, smem);
    }
}

template <int R>
static cudaError_t launch_extract_r(bool is_f32, bool fast, const KParams& p, int64_t n_tiles, int num_sms, cudaStream_t st,
                                    size_t smem)
{
    const int v = variant_for_r<R>(p);
    if (v == 3) return launch_r<R, 3, true>(is_f32, fast, p, n_tiles, num_sms, st, smem);
    if constexpr (R == 32) {
        if (v == 1) return launch_r<R, 1, true>(is_f32, fast, p, n_tiles, num_sms, st, smem);
    }
    return launch_r<R, 1, false>(is_f32, fast, p, n_tiles, num_sms, st, smem);
}

cudaError_t launch_extract(int r, bool is_f32, bool fast, const KParams& p, int64_t n_tiles, int num_sms,
                           cudaStream_t st, size_t smem)
{
#ifdef SCF_VARIANT_BUILD      // tuning builds (tools/build_variant.sh): only the params.json fast-path kernels
    if (r == 32 && !is_f32 && fast) {
        const int v = variant_for_r<32>(p);
        if (v == 3) return launch_one<32, int16_t, true, 3, true, false>(p, n_tiles, num_sms, st, smem);
        if (v == 1) return launch_one<32, int16_t, true, 1, true, false>(p, n_tiles, num_sms, st, smem);
    }
    return cudaErrorInvalidValue;
#else
    switch (r) {
        case 32: return launch_extract_r<32>(is_f32, fast, p, n_tiles, num_sms, st, smem);
        case 16: return launch_extract_r<16>(is_f32, fast, p, n_tiles, num_sms, st, smem);
        case 8: return launch_extract_r<8>(is_f32, fast, p, n_tiles, num_sms, st, smem);
        default: return cudaErrorInvalidValue;
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// Delta columns over the frame axis, in place, one thread per (row, base column): rows[clip][f][cols * (1 + b) + c].
//  kind 1: add_deltas, common/data_utils.py:50-58 -- d[f] = x[f] - x[f-1], d[0] = 0
//  kind 2: inference/tflite/mfcc.h:432-441 -- d[f] = (x[min(f+1, n-1)] - x[max(f-1, 0)]) / 2
//  kind 3: ... plus mfcc.h:443-453 -- the same difference of the delta column (recomputed from the base rows, which gives
//          the stored values bit for bit)
__global__ void __launch_bounds__(256) delta_kernel(float* __restrict__ rows, const int32_t* __restrict__ lengths,
                                                    int64_t n_items, int fpc, int cols, int pitch, int kind, int clip_len,
                                                    int window, int hop, int pad_mode)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_items) return;
    const int c = (int)(idx % cols);
    const int64_t rf = idx / cols;
    const int f = (int)(rf % fpc);
    const int64_t clip = rf / fpc;
    int n = fpc;                                        // rows this clip has
    if (pad_mode == SCF_PAD_NONE && lengths != nullptr) {
        const int len = min(max(lengths[clip], 0), clip_len);
        n = len >= window ? (len - window) / hop + 1 : 0;
    }
    if (f >= n) return;
    float* base = rows + clip * (int64_t)fpc * pitch + c;
    auto X = [&](int j) { return base[(int64_t)j * pitch]; };
    if (kind == SCF_DELTA_DIFF) {
        base[(int64_t)f * pitch + cols] = f == 0 ? 0.f : X(f) - X(f - 1);
        return;
    }
    auto D = [&](int j) { return (X(min(j + 1, n - 1)) - X(max(j - 1, 0))) / 2; };
    base[(int64_t)f * pitch + cols] = D(f);
    if (kind == SCF_DELTA_CENTRAL2) base[(int64_t)f * pitch + 2 * cols] = (D(min(f + 1, n - 1)) - D(max(f - 1, 0))) / 2;
}

cudaError_t launch_delta(float* rows, const int32_t* lengths, int64_t n_clips, int frames_per_clip, int cols, int pitch,
                         int kind, int clip_len, int window, int hop, int pad_mode, cudaStream_t st)
{
    const int64_t n_items = n_clips * frames_per_clip * cols;
    if (n_items <= 0) return cudaSuccess;
    delta_kernel<<<(unsigned)((n_items + 255) / 256), 256, 0, st>>>(rows, lengths, n_items, frames_per_clip, cols, pitch, kind,
                                                                  clip_len, window, hop, pad_mode);
    count_launch(1);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// FP32 FMA throughput probe: 8 independent FMA chains per thread, 2 flop per FMA.
__global__ void __launch_bounds__(256) fp32_probe_kernel(float* out, int iters)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-4f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
            a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

#ifdef SCF_DEBUG_TIMES
extern "C" int scf_debug_phases(unsigned long long* out)         // 2 x 8 x 2048 values
{
    return (int)cudaMemcpyFromSymbol(out, g_dbg_phase, sizeof(unsigned long long) * 2 * 8 * 2048);
}
extern "C" int scf_debug_times(unsigned long long* out4096)      // 8 x 2048 values
{
    return (int)cudaMemcpyFromSymbol(out4096, g_dbg_times, sizeof(unsigned long long) * 8 * 2048);
}
#endif

cudaError_t launch_fp32_probe(float* out, int iters, int grid, cudaStream_t st)
{
    fp32_probe_kernel<<<grid, 256, 0, st>>>(out, iters);
    return cudaGetLastError();
}

}  // namespace scf
