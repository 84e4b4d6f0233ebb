// vela_api.cu — host side of libvela_b200.so: handle lifetime, per-device model replicas, batch
// orchestration (streams, pinned staging, walker sharding across devices) and the C ABI of
// include/vela_b200.h.  There is no CPU compute path in this library.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <memory>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "vela_host_prep.h"
#include "vela_jit.h"
#include "vela_kernels.cuh"
#include "vela_structured.h"

using namespace vela;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? VELA_ERR_OOM : VELA_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(_e), __FILE__, __LINE__);                                     \
    } while (0)

template <class T>
int upload(T** dst, const T* src, size_t n) {
    *dst = nullptr;
    if (n == 0) n = 1;
    CU(cudaMalloc((void**)dst, n * sizeof(T)));
    CU(cudaMemset(*dst, 0, n * sizeof(T)));
    if (src) CU(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

inline long round_up(long x, long m) { return (x + m - 1) / m * m; }
constexpr long kSplitRows = 2048;   // walker rows reserved for split-K slabs of small batches
constexpr long kTwoStageMinBatch = 1024; // batches from this size on take the two-stage column sums (cell moments)
constexpr long kGraphMaxBatch = 2048; // host batches up to this size replay a captured CUDA graph
constexpr long kMaxBigSplit = 4;    // split-K slabs kept for full-size batches (wave quantisation), memory permitting

struct Buffers {  // per-device, per-batch scratch (grown on demand)
    long cap_B = 0;
    long sig_rows = 0;   // walker rows of the Sigma / U scratch (>= cap_B; larger so small batches can split K)
    double *params = nullptr, *Pext = nullptr, *tzr = nullptr, *partial = nullptr, *Y = nullptr, *W = nullptr;
    double *Sig = nullptr, *U = nullptr, *Sq = nullptr, *V = nullptr, *out = nullptr, *lnprior = nullptr;
    double* R = nullptr;   // structured path: column sums [r_rows][n_cols]
    long r_rows = 0;
    double* mu = nullptr;  // two-stage column sums: cell moments [n_blocks][cap_B][k2_pad]
    long mu_rows = 0;
    double *h_params = nullptr, *h_out = nullptr, *h_lnprior = nullptr;  // pinned
    long h_cap = 0;
    // Small host batches (emcee-sized) replay a captured CUDA graph: staging copy in, every kernel of the step, result
    // copy out - one launch instead of eight.  Keyed by (walkers, mode, prior source); the graphs hold the addresses of
    // the buffers above, so they are dropped whenever a buffer is reallocated.
    struct GraphEntry { cudaGraphExec_t exec = nullptr; int launches = 0; };
    std::map<std::tuple<long, int, int, int>, GraphEntry> graphs;
};

void drop_graphs(Buffers& b) {
    for (auto& kv : b.graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    b.graphs.clear();
}

struct DeviceModel {
    int dev = 0;
    int n_sm = 148;
    cudaStream_t stream = nullptr;
    double* toa = nullptr;
    long toa_stride = 0;
    double* tzr_block = nullptr;
    uint32_t* index_masks = nullptr;
    uint8_t* bit_masks = nullptr;
    double* defaults = nullptr;
    int* free_idx = nullptr;
    double* M = nullptr;
    long ldM = 0;
    double* T = nullptr;            // structured path: [n_rows_pad x n_cols] table (device copy of StructPlan::T)
    PairTerm* table = nullptr;
    double* A1 = nullptr;           // two-stage column sums: stage-1 Chebyshev columns [n_rows_pad x 32 n_blocks]
    double* T2 = nullptr;           //                        stage-2 matrices (StructPlan::jobs)
    ColsumUnit* units2 = nullptr;   //                        stage-2 CTA columns (64 output columns each), n_units2 of them
    int n_units2 = 0;
    unsigned* ltab = nullptr;       // chol_warp_kernel layout table (ranks <= 63 on the structured path)
    unsigned* ftab = nullptr;       // chol_dmma_kernel fragment table (ranks <= 64 on the structured path)
    unsigned* ctab = nullptr;       // compact pair table (4 bytes per pair) of the structured path
    double* tzr0 = nullptr;         // kFlagDelta: Double64 spin phase of the TZR TOA at the default parameters (2 doubles)
    GPDev* gp = nullptr;
    double* gp_data = nullptr;
    int3* ecorr_groups = nullptr;
    int* ecorr_chunks = nullptr;
    int3* ecorr_active = nullptr;   // groups with an ECORR parameter (index >= 1), for the Woodbury correction
    PriorDev* priors = nullptr;     // device-side priors (vela_b200_set_priors)
    Buffers buf;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float stage_ms[5] = {0, 0, 0, 0, 0};
    // The per-batch scratch (Pext, Y/W, R, ...) is one set per device.  Calls may enqueue on different streams (the
    // handle's own, or a caller's through the *_device entry points) and return before the work has run, so the last
    // user records `busy` on its stream and the next one makes its stream wait for it before touching the scratch.
    cudaEvent_t busy = nullptr;
    bool busy_valid = false;
};

// One host thread per device of a multi-device handle, alive for the life of the handle: it stages its device's block
// through pinned memory and enqueues the kernels, so the devices start side by side (run_host_batch).
struct Feeder {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, stop = false;
    void start() {
        th = std::thread([this] {
            std::unique_lock<std::mutex> lk(m);
            for (;;) {
                cv.wait(lk, [this] { return has_job || stop; });
                if (stop) return;
                lk.unlock();
                job();
                lk.lock();
                has_job = false;
                cv.notify_all();
            }
        });
    }
    void post(std::function<void()> f) {
        std::lock_guard<std::mutex> lk(m);
        job = std::move(f);
        has_job = true;
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [this] { return !has_job; });
    }
    void shutdown() {
        if (!th.joinable()) return;
        { std::lock_guard<std::mutex> lk(m); stop = true; cv.notify_all(); }
        th.join();
    }
};

void scratch_acquire(DeviceModel& d, cudaStream_t st) {
    if (d.busy_valid) cudaStreamWaitEvent(st, d.busy, 0);
}
void scratch_release(DeviceModel& d, cudaStream_t st) {
    if (d.busy && cudaEventRecord(d.busy, st) == cudaSuccess) d.busy_valid = true;
}

}  // namespace

struct VelaPulsar {
    ProgramDev prog;
    StructPlan sp;
    int kernel_kind = 0;
    long n_toas = 0, n_rows = 0, n_rows_pad = 0;
    int p = 0, pp = 0, n_gp = 0;
    long n_pairs = 0, ld_sig = 0;
    int sig_shape = 0;
    size_t sig_smem = 0;
    int chain_threads = 256, chain_group = 32, n_tiles = 0;
    int n_ecorr_groups = 0, n_ecorr_chunks = 0, n_ecorr_rows = 0, par_ecorr = -1;   // EcorrKernel: group corrections occupy partial rows n_tiles..
    int n_eg = 0, n_rect = 0;        // WoodburyKernel{EcorrKernel}: active groups, 32x32 rectangles of the (p+1) triangle
    long eg_pad = 0, ldv = 0;
    bool chol_smem = true;
    std::vector<unsigned> ltab;      // layout table of chol_warp_kernel; empty = that kernel is not used
    std::vector<unsigned> ftab;      // fragment table of chol_dmma_kernel; empty = not available
    std::vector<unsigned> ctab;      // compact pair table; empty = some coefficient / index does not fit the 4-byte code
    bool delta_ready = false;        // the C_DELTA columns of every device's TOA table are filled (kFlagDelta in use)
    int n_index_masks = 0;
    int chol_nb = 8;                 // block width of the Cholesky kernel
    size_t chol_smem_bytes = 0;
    std::vector<DeviceModel> devs;
    std::vector<std::unique_ptr<Feeder>> feeders;   // one per device when the handle spans several
    std::mutex mtx;
    std::atomic<int64_t> launches{0};   // several feeder threads when the handle spans several devices
    bool stage_timing = false;
    bool graphs_ok = true;               // cleared when a capture fails (the plain launch path always works)
    bool have_priors = false;
    long max_chunk = 0;  // walkers per device per pass
    cudaKernel_t spec_chain = nullptr;   // NVRTC-specialised chain kernel (vela_jit.cu); null = generic kernel only
    cudaKernel_t spec_prologue = nullptr;
    int spec_emit = -1;                  // ChainArgs::emit the specialised chain kernel was built for
    bool use_spec = false;
    std::string jit_log;
};

namespace {

void free_buffers(Buffers& b) {
    drop_graphs(b);
    cudaFree(b.params); cudaFree(b.Pext); cudaFree(b.tzr); cudaFree(b.partial); cudaFree(b.Y); cudaFree(b.W);
    cudaFree(b.Sig); cudaFree(b.U); cudaFree(b.Sq); cudaFree(b.V); cudaFree(b.out); cudaFree(b.lnprior); cudaFree(b.R); cudaFree(b.mu);
    b.mu = nullptr; b.mu_rows = 0;
    b.params = b.Pext = b.tzr = b.partial = b.Y = b.W = b.Sig = b.U = b.Sq = b.V = b.out = b.lnprior = b.R = nullptr;
    b.r_rows = 0;
    b.cap_B = 0;
    b.sig_rows = 0;
}

// Batches from kTwoStageMinBatch walkers on take the two-stage column sums when the basis has a plan
// (VELA_B200_TWO_STAGE=0 switches them off, VELA_B200_TWO_STAGE_MIN moves the threshold: A/B measurements).
bool two_stage_taken(const VelaPulsar* h, long B) {
    if (!h->sp.two_stage) return false;
    if (const char* e = getenv("VELA_B200_TWO_STAGE")) if (atoi(e) == 0) return false;
    long min_b = kTwoStageMinBatch;
    if (const char* e = getenv("VELA_B200_TWO_STAGE_MIN")) min_b = std::max(1, atoi(e));
    return B >= min_b;
}

int ensure_device_buffers(VelaPulsar* h, DeviceModel& d, long B) {
    Buffers& b = d.buf;
    if (B <= b.cap_B) return 0;
    free_buffers(b);
    long cap = std::max<long>(round_up(B, kSigMaxSamples), 64);
    const ProgramDev& pg = h->prog;
    CU(cudaMalloc((void**)&b.params, (size_t)cap * std::max(pg.n_free, 1) * sizeof(double)));
    CU(cudaMalloc((void**)&b.Pext, (size_t)cap * pg.row * sizeof(double)));
    CU(cudaMalloc((void**)&b.tzr, (size_t)cap * 2 * sizeof(double)));
    CU(cudaMalloc((void**)&b.partial, (size_t)(h->n_tiles + std::max(h->n_ecorr_rows, (h->n_eg + 31) / 32)) * cap * 2 * sizeof(double)));
    CU(cudaMalloc((void**)&b.out, (size_t)cap * sizeof(double)));
    CU(cudaMalloc((void**)&b.lnprior, (size_t)cap * sizeof(double)));
    if (h->kernel_kind != VELA_KERNEL_WHITE) {   // every kernel but the plain white one needs y, w
        size_t yw = (size_t)cap * h->n_rows_pad * sizeof(double);
        CU(cudaMalloc((void**)&b.Y, yw));
        CU(cudaMalloc((void**)&b.W, yw));
        CU(cudaMemsetAsync(b.Y, 0, yw, d.stream));
        CU(cudaMemsetAsync(b.W, 0, yw, d.stream));
    }
    if (h->p > 0 && h->sp.active) {
        // room for deep split-K at small and medium batches: R rows are only n_cols doubles each
        b.r_rows = std::max<long>(cap * 4, std::min<long>(65536, ((long)64 << 20) / ((long)h->sp.n_cols * 8)));
        CU(cudaMalloc((void**)&b.R, (size_t)b.r_rows * h->sp.n_cols * sizeof(double)));
        CU(cudaMemsetAsync(b.R, 0, (size_t)b.r_rows * h->sp.n_cols * sizeof(double), d.stream));
        if (h->sp.two_stage && two_stage_taken(h, cap)) {
            b.mu_rows = round_up(cap, 512);
            const size_t nb = (size_t)h->sp.n_blocks * b.mu_rows * h->sp.k2_pad * sizeof(double);
            CU(cudaMalloc((void**)&b.mu, nb));
            CU(cudaMemsetAsync(b.mu, 0, nb, d.stream));   // the padding behind the last cell stays zero
        }
    }
    if (h->p > 0 && h->sp.active && h->n_eg == 0) {
        if (!h->chol_smem) CU(cudaMalloc((void**)&b.Sq, (size_t)cap * chol_panel_scratch_doubles(h->p, h->pp) * sizeof(double)));
    } else if (h->p > 0) {
        // Sigma / U scratch: at least `cap` walker rows, and never fewer than kSplitRows so that small batches can
        // hold several split-K slabs
        const long slabs = std::max<long>(1, std::min<long>(kMaxBigSplit, (long)(((size_t)6 << 30) / ((size_t)cap * h->ld_sig * sizeof(double)))));
        b.sig_rows = std::max<long>(cap * slabs, kSplitRows);
        CU(cudaMalloc((void**)&b.Sig, (size_t)b.sig_rows * h->ld_sig * sizeof(double)));
        CU(cudaMalloc((void**)&b.U, (size_t)b.sig_rows * h->pp * sizeof(double)));
        CU(cudaMemsetAsync(b.Sig, 0, (size_t)b.sig_rows * h->ld_sig * sizeof(double), d.stream));
        CU(cudaMemsetAsync(b.U, 0, (size_t)b.sig_rows * h->pp * sizeof(double), d.stream));
        if (!h->chol_smem) CU(cudaMalloc((void**)&b.Sq, (size_t)cap * chol_panel_scratch_doubles(h->p, h->pp) * sizeof(double)));
        if (h->n_eg > 0) {
            size_t vb = (size_t)cap * h->eg_pad * h->ldv * sizeof(double);
            CU(cudaMalloc((void**)&b.V, vb));
            CU(cudaMemsetAsync(b.V, 0, vb, d.stream));   // padding rows / columns stay zero
        }
    }
    b.cap_B = cap;
    return 0;
}

int ensure_host_buffers(VelaPulsar* h, DeviceModel& d, long B) {
    Buffers& b = d.buf;
    if (B <= b.h_cap) return 0;
    if (b.h_params) cudaFreeHost(b.h_params);
    if (b.h_out) cudaFreeHost(b.h_out);
    if (b.h_lnprior) cudaFreeHost(b.h_lnprior);
    drop_graphs(b);
    b.h_params = b.h_out = b.h_lnprior = nullptr;   // a failed allocation below must not leave stale pointers behind
    b.h_cap = 0;
    long cap = std::max<long>(B, 64);
    CU(cudaMallocHost((void**)&b.h_params, (size_t)cap * std::max(h->prog.n_free, 1) * sizeof(double)));
    CU(cudaMallocHost((void**)&b.h_out, (size_t)cap * sizeof(double)));
    CU(cudaMallocHost((void**)&b.h_lnprior, (size_t)cap * sizeof(double)));
    b.h_cap = cap;
    return 0;
}

// K3 launch shapes: (BK, stages, n-tiles, warps) with two row-tile heights per shape; see DESIGN.md section 5
struct SigmaShape { int bk, stages, nt, warps, ctas_per_sm; void (*kernel4)(SigmaArgs); void (*kernel3)(SigmaArgs); };
#define VELA_SHAPE(BK, ST, NT, W, C) \
    SigmaShape{BK, ST, NT, W, C, sigma_contract_kernel<BK, ST, 4, NT, W>, sigma_contract_kernel<BK, ST, 3, NT, W>}
const SigmaShape kSigmaShapes[] = {
    VELA_SHAPE(64, 3, 4, 16, 1), VELA_SHAPE(32, 4, 4, 16, 1), VELA_SHAPE(32, 3, 4, 16, 1), VELA_SHAPE(16, 4, 4, 16, 1),
    VELA_SHAPE(16, 3, 4, 16, 1), VELA_SHAPE(16, 2, 4, 16, 1), VELA_SHAPE(8, 2, 4, 16, 1),
};

// K3 work quantisation.  A CTA is `warps` units of MT m8-tiles over 32 walkers and one CTA is resident per SM, so a
// launch costs about  CTAs x t_cta / SMs + 0.8 t_cta  (throughput term + tail; the tail factor is fitted to the
// measurements in tools/probes/gpu_plan_probe.py).  Two knobs trade padding against the tail: the row-tile height MT (4 or 3
// m8-tiles per warp) and split-K over gridDim.z (K4 sums the slabs in slice order, so the result stays deterministic
// for a given batch size).  Small batches use the same model: few CTAs -> deep split.
struct SigmaPlan { int mt, ksplit, n_pair_units, n_units; };
SigmaPlan plan_sigma(const VelaPulsar* h, const SigmaShape& sh, long sig_rows, long B, int n_sm) {
    const int nsamp = 8 * sh.nt;
    const long n_ktiles = h->n_rows_pad / sh.bk;
    const long pair_tiles = (h->n_pairs + 7) / 8, u_tiles = h->pp / 8;
    const long max_split = std::max<long>(1, std::min<long>({64, std::max<long>(1, n_ktiles / 4),
                                                              sig_rows / std::max<long>(round_up(B, kSigMaxSamples), 1)}));
    // Units: one k-tile of 64 TOAs for one m8 tile row of a CTA (2.4 us on a B200).  Measured at config 3 / 4
    // (tools/probes/gpu_plan_probe.py): the MT = 3 kernel runs ~12 % slower per DMMA (10 LDS per 12 DMMAs instead of 12 per
    // 16), and every extra split-K slab costs K4 one more pass over the walkers' packed triangles.
    const double kFixed = 6.0;                         // pipeline fill + epilogue of one CTA
    const double tile = sh.bk / 64.0;
    const double slab = 1.3e-6 * (double)B * (double)h->n_pairs;
    SigmaPlan best{4, 1, 0, 0};
    double best_cost = 1e300;
    const int force_mt = getenv("VELA_B200_SIGMA_MT") ? atoi(getenv("VELA_B200_SIGMA_MT")) : 0;
    const int force_ks = getenv("VELA_B200_SIGMA_KSPLIT") ? atoi(getenv("VELA_B200_SIGMA_KSPLIT")) : 0;
    for (int mt = 4; mt >= 3; --mt) {
        if (force_mt && mt != force_mt) continue;
        const long pu = (pair_tiles + mt - 1) / mt, units = pu + (u_tiles + mt - 1) / mt;
        const long ctas_xy = ((units + sh.warps - 1) / sh.warps) * ((B + nsamp - 1) / nsamp);
        for (long ks = 1; ks <= max_split; ++ks) {
            if (force_ks && ks != std::min<long>(force_ks, max_split)) continue;
            const long kt = (n_ktiles + ks - 1) / ks, nz = (n_ktiles + kt - 1) / kt;
            if (nz != ks) continue;   // same slicing as a smaller ks
            const double t_cta = (double)kt * tile * mt * (mt == 3 ? 1.12 : 1.0) + kFixed;
            const double n_cta = (double)(ctas_xy * nz);
            const double cost = std::max(n_cta / n_sm, 1.0) * t_cta + 0.8 * t_cta + slab * (double)(ks - 1);
            if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; best = SigmaPlan{mt, (int)ks, (int)pu, (int)units}; }
        }
    }
    return best;
}

// K3s launch shapes (BK, stages, column units, walker tiles per CTA); the second table is for batches whose Y buffer
// holds w.y (homogeneous CTA columns, one walker operand per stage)
struct ColsumShape { int bk, stages, uc, wc; void (*kernel)(ColsumArgs); };
const ColsumShape kColsumShapes[] = {
    {16, 2, 16, 1, colsum_kernel<16, 2, 16, 1, false>}, {16, 3, 8, 2, colsum_kernel<16, 3, 8, 2, false>},
    {16, 3, 4, 4, colsum_kernel<16, 3, 4, 4, false>}, {16, 2, 2, 8, colsum_kernel<16, 2, 2, 8, false>},
};
const ColsumShape kColsumSplitShapes[] = {
    {8, 4, 16, 1, colsum_kernel<8, 4, 16, 1, true>}, {16, 4, 8, 2, colsum_kernel<16, 4, 8, 2, true>},
    // (2, 8): two stages of 32 TOAs beat four of 16 (1.04 vs 1.10 ms at config 3); three of 32 do not fit
    {16, 4, 4, 4, colsum_kernel<16, 4, 4, 4, true>}, {32, 2, 2, 8, colsum_kernel<32, 2, 2, 8, true>},
    {16, 4, 2, 8, colsum_kernel<16, 4, 2, 8, true>},   // (2, 8) for small batches: twice the k-tiles to split over
};

// Launches K3s for B walkers; `wy`: the chain kernel stored w.y in Y (emit 3).  Returns the number of split-K slabs.
int launch_colsum(VelaPulsar* h, DeviceModel& d, long B, cudaStream_t st, bool wy, ColsumArgs* out_args = nullptr) {
    Buffers& b = d.buf;
    const int n_units = h->sp.n_cols / 32, n_t = h->sp.n_r / 32, n_u = n_units - n_t;
    const ColsumShape* shapes = wy ? kColsumSplitShapes : kColsumShapes;
    // CTA columns for a shape: homogeneous columns pad the two kinds of units separately
    auto cta_cols = [&](const ColsumShape& sh) -> long {
        return wy ? (n_t + sh.uc - 1) / sh.uc + (n_u + sh.uc - 1) / sh.uc : (n_units + sh.uc - 1) / sh.uc;
    };
    // the shape that pads the unit count least; ties go to the deeper rings and, for many units, to wide CTAs
    // (tools/probes/gpu_colsum_probe.py)
    int si = n_units <= 8 ? 2 : 1;
    {
        long best = cta_cols(shapes[si]) * shapes[si].uc;
        for (int i = 1; i < 4; ++i) {
            const long padded = cta_cols(shapes[i]) * shapes[i].uc;
            if (padded < best) { best = padded; si = i; }
        }
    }
    // emcee-sized batches need deep split-K to cover the SMs: the 16-TOA ring has twice the k-tiles to split over
    // (config 3, column sums alone: 39 us for any B <= 256, against 54 us with the 32-TOA ring).  Up to 128 walkers the
    // (8, 2) shape is faster still (21 us): a (2, 8) CTA stages the w rows of 256 walkers whatever the batch holds, and
    // at these sizes the kernel is bound by that L2 traffic (tools/probes/gpu_small_batch_probe.py)
    if (wy && si == 3 && cta_cols(shapes[3]) * ((B + 255) / 256) * (h->n_rows_pad / 32 / 8) < 2L * d.n_sm) si = B <= 128 ? 1 : 4;
    if (const char* e = getenv("VELA_B200_COLSUM_SHAPE")) si = std::max(0, std::min(wy ? 4 : 3, atoi(e)));
    const ColsumShape& sh = shapes[si];
    const long n_ktiles = h->n_rows_pad / sh.bk;
    const long ctas_xy = cta_cols(sh) * ((B + 32 * sh.wc - 1) / (32 * sh.wc));
    const long max_split = std::max<long>(1, std::min<long>({64, std::max<long>(1, n_ktiles / 8),
                                                              b.r_rows / std::max<long>(round_up(B, kSigMaxSamples), 1)}));
    long ks_best = 1;
    double best_cost = 1e300;
    // all CTAs of this kernel take the same time and one is resident per SM, so the launch runs in (nearly) whole
    // waves: cost = waves x (k-tiles per CTA + fixed cost) with waves half-way between ceil(CTAs / SMs) and the
    // fractional count; measured at config 3 (tools/probes/gpu_ks_probe.py): 288 CTAs (ks 3, 1.95 waves) 1.09 ms,
    // 384 CTAs (ks 4) 1.19 ms, 480 / 576 CTAs 1.24 ms; at config 4 (112 CTAs per slice) ks 1 / 4 / 8: 6.48 / 6.13 / 6.07 ms
    for (long ks = 1; ks <= max_split; ++ks) {
        const long kt = (n_ktiles + ks - 1) / ks, nz = (n_ktiles + kt - 1) / kt;
        if (nz != ks) continue;
        const double t_cta = (double)kt * (sh.bk / 16.0) + 16.0;
        const double frac = std::max((double)(ctas_xy * nz) / d.n_sm, 1.0);
        const double waves = 0.5 * (std::ceil(frac) + frac);
        const double cost = waves * t_cta * (1.0 + 0.002 * (double)(ks - 1));
        if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; ks_best = ks; }
    }
    if (const char* e = getenv("VELA_B200_SIGMA_KSPLIT")) ks_best = std::max<long>(1, std::min<long>(atoi(e), max_split));
    ColsumArgs ca;
    ca.plane_stride = 0; ca.n_valid = 0;
    ca.T = d.T; ca.ldT = h->n_rows_pad; ca.n_cols = h->sp.n_cols; ca.n_w_units = n_t; ca.n_rows_pad = h->n_rows_pad;
    ca.W = b.W; ca.Y = b.Y; ca.ld_yw = h->n_rows_pad; ca.B = B; ca.R = b.R; ca.ld_r = h->sp.n_cols;
    ca.kt_per_split = (int)((n_ktiles + ks_best - 1) / ks_best);
    const int nz = (int)((n_ktiles + ca.kt_per_split - 1) / ca.kt_per_split);
    ca.split_stride_r = round_up(B, kSigMaxSamples) * (long)h->sp.n_cols;
    dim3 grid((unsigned)cta_cols(sh), (unsigned)((B + 32 * sh.wc - 1) / (32 * sh.wc)), (unsigned)nz);
    sh.kernel<<<grid, 512, colsum_smem_bytes(sh.bk, sh.stages, sh.uc, sh.wc, wy), st>>>(ca);
    if (out_args) *out_args = ca;
    return nz;
}

// Two-stage column sums (vela_structured.h): stage 1 contracts w / w.y with the Chebyshev columns cell by cell (one
// launch, split-K slices = cells, moments stored walker-major per block), stage 2 maps the moments to the column sums
// with small constant matrices, one run of output columns per block.  Same kernel both times.  Returns the number of
// split-K slabs of R that stage 2 wrote.
int launch_two_stage(VelaPulsar* h, DeviceModel& d, long B, cudaStream_t st, ColsumArgs* out_args) {
    Buffers& b = d.buf;
    const StructPlan& sp = h->sp;
    const long plane = b.mu_rows * sp.k2_pad;
    ColsumArgs ca;
    // ---- stage 1
    ca.T = d.A1; ca.ldT = h->n_rows_pad; ca.n_cols = sp.n_blocks * 32; ca.n_w_units = sp.n_w_blocks; ca.n_rows_pad = h->n_rows_pad;
    ca.W = b.W; ca.Y = b.Y; ca.ld_yw = h->n_rows_pad; ca.B = B; ca.R = b.mu; ca.ld_r = sp.k2_pad;
    ca.kt_per_split = sp.cp / 16; ca.split_stride_r = 0; ca.plane_stride = plane; ca.n_valid = 0;
    {
        // CTA shape, measured on B200 (stage 1 + stage 2, ms; config 3 / config 5 at 8192 walkers): 512 walkers x 2 stages
        // of 16 rows 0.544 / 17.43; 512 x 4 stages of 8 rows 0.523 / 18.16; 256 walkers x 2 stages, TWO CTAs per SM
        // 0.488 / 15.35 (default: the second CTA computes while the first waits for its tile); 256 x 4 stages 0.519 / 15.78
        int shape = 2;
        if (const char* e = getenv("VELA_B200_STAGE1_SHAPE")) shape = std::max(0, std::min(3, atoi(e)));
        if (shape == 1) {
            ca.kt_per_split = sp.cp / 8;
            dim3 grid((unsigned)sp.n_blocks, (unsigned)((B + 511) / 512), (unsigned)sp.n_cells);
            colsum_kernel<8, 4, 1, 16, true><<<grid, 512, colsum_smem_bytes(8, 4, 1, 16, true), st>>>(ca);
        } else if (shape == 2) {
            dim3 grid((unsigned)sp.n_blocks, (unsigned)((B + 255) / 256), (unsigned)sp.n_cells);
            colsum_kernel<16, 2, 1, 8, true><<<grid, 256, colsum_smem_bytes(16, 2, 1, 8, true), st>>>(ca);
        } else if (shape == 3) {
            dim3 grid((unsigned)sp.n_blocks, (unsigned)((B + 255) / 256), (unsigned)sp.n_cells);
            colsum_kernel<16, 4, 1, 8, true><<<grid, 256, colsum_smem_bytes(16, 4, 1, 8, true), st>>>(ca);
        } else {
            dim3 grid((unsigned)sp.n_blocks, (unsigned)((B + 511) / 512), (unsigned)sp.n_cells);
            colsum_kernel<16, 2, 1, 16, true><<<grid, 512, colsum_smem_bytes(16, 2, 1, 16, true), st>>>(ca);
        }
    }
    // ---- stage 2: every job is a dense [n_out x k2_pad] x [k2_pad x B] product on its block's moments.  All jobs go in
    // ONE launch: blockIdx.x walks the table of 64-column CTA columns built at create() (d.units2), so the grid is
    // (CTA columns of all jobs) x (256-walker tiles) x (split-K slices) however short the single runs are - one launch
    // per job left 32 - 64 CTAs on 148 SMs (0.17 + 0.17 ms at config 3).  Split-K by the same wave model as launch_colsum.
    ColsumArgs c2;
    c2.T = d.T2; c2.ldT = sp.k2_pad; c2.n_cols = 64; c2.n_w_units = 2; c2.n_rows_pad = sp.k2_pad;
    c2.W = b.mu; c2.Y = b.mu; c2.ld_yw = sp.k2_pad; c2.B = B; c2.w_plane = plane;
    c2.R = b.R; c2.ld_r = sp.n_cols; c2.plane_stride = 0; c2.n_valid = 0; c2.units = d.units2;
    c2.split_stride_r = round_up(B, kSigMaxSamples) * (long)sp.n_cols;
    const long n_ktiles = sp.k2_pad / 16;
    const long ctas_xy = (long)d.n_units2 * ((B + 255) / 256);
    const long max_split = std::max<long>(1, std::min<long>({kMaxBigSplit, std::max<long>(1, n_ktiles / 8),
                                                              b.r_rows / std::max<long>(round_up(B, kSigMaxSamples), 1)}));
    long ks_best = 1;
    double best_cost = 1e300;
    for (long ks = 1; ks <= max_split; ++ks) {
        const long kt = (n_ktiles + ks - 1) / ks, nz = (n_ktiles + kt - 1) / kt;
        if (nz != ks) continue;
        const double frac = std::max((double)(ctas_xy * nz) / d.n_sm, 1.0);
        const double cost = 0.5 * (std::ceil(frac) + frac) * ((double)kt + 16.0) * (1.0 + 0.002 * (double)(ks - 1));
        if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; ks_best = ks; }
    }
    if (const char* e = getenv("VELA_B200_STAGE2_KSPLIT")) ks_best = std::max<long>(1, std::min<long>(atoi(e), max_split));
    c2.kt_per_split = (int)((n_ktiles + ks_best - 1) / ks_best);
    const int nz = (int)((n_ktiles + c2.kt_per_split - 1) / c2.kt_per_split);
    {
        dim3 grid((unsigned)d.n_units2, (unsigned)((B + 255) / 256), (unsigned)nz);
        colsum_kernel<16, 4, 2, 8, true><<<grid, 512, colsum_smem_bytes(16, 4, 2, 8, true), st>>>(c2);
    }
    h->launches += 2;
    if (out_args) *out_args = c2;
    return nz;
}

// WoodburyKernel{EcorrKernel}: subtract the per-group Sherman-Morrison terms from slab 0 of Sigma / U (K2b + K2c).
// K2b also emits the scalar ECORR corrections (one `partial` row per group block, after the TOA tiles); returns the
// number of those rows.
int launch_ecorr_correction(VelaPulsar* h, DeviceModel& d, long B, cudaStream_t st, bool wy) {
    Buffers& b = d.buf;
    EcorrBasisArgs ba;
    ba.y_is_wy = wy ? 1 : 0;
    ba.partial = b.partial; ba.first_row = h->n_tiles;
    ba.M = d.M; ba.ldM = d.ldM; ba.rank = h->p; ba.Pext = b.Pext; ba.row = h->prog.row; ba.par_ecorr = h->par_ecorr;
    ba.Y = b.Y; ba.W = b.W; ba.ld_yw = h->n_rows_pad; ba.B = B; ba.groups = d.ecorr_active; ba.n_groups = h->n_eg;
    ba.groups_per_block = 32; ba.V = b.V; ba.g_pad = h->eg_pad; ba.ldv = h->ldv;
    dim3 grid((unsigned)((h->n_eg + ba.groups_per_block - 1) / ba.groups_per_block), (unsigned)((B + 31) / 32));
    int chunk = kEcorrChunk;
    const int nacc = (h->p + 7) / 8;
    const int pitch = nacc <= 8 ? 65 : (nacc <= 16 ? 129 : h->p + 1);   // ecorr_basis_kernel: zero-padded rows up to rank 128
    auto smem_for = [&](int ch) { return ((size_t)ch * pitch + 2 * 32 * (ch + 1)) * sizeof(double); };
    // the staging buffer must leave room for the blocks per SM the kernel is compiled for (4 / 2 / 1 by rank): at
    // rank 60 a 64-TOA pass (64.5 KB) would cap the SM at three blocks; 32 TOAs measured 5.00 vs 5.09 ms
    const int blocks_per_sm = h->p <= 64 ? 4 : (h->p <= 128 ? 2 : 1);
    while (chunk > 8 && smem_for(chunk) > (size_t)224 * 1024 / blocks_per_sm) chunk /= 2;
    if (const char* e = getenv("VELA_B200_ECORR_CHUNK")) { int v = atoi(e); if (v == 4 || v == 8 || v == 16 || v == 32 || v == 64 || v == 128) chunk = v; }
    while (chunk > 8 && smem_for(chunk) > (size_t)200 * 1024) chunk /= 2;
    ba.chunk = chunk;
    const size_t sm = smem_for(chunk);
    if (nacc <= 8) ecorr_basis_kernel<8><<<grid, 256, sm, st>>>(ba);
    else if (nacc <= 16) ecorr_basis_kernel<16><<<grid, 256, sm, st>>>(ba);
    else if (nacc <= 40) ecorr_basis_kernel<40><<<grid, 256, sm, st>>>(ba);
    else ecorr_basis_kernel<128><<<grid, 256, sm, st>>>(ba);
    EcorrSyrkArgs sa;
    sa.V = b.V; sa.g_pad = h->eg_pad; sa.ldv = h->ldv; sa.rank = h->p; sa.n_rect = h->n_rect; sa.B = B;
    sa.Sig = b.Sig; sa.ld_sig = h->ld_sig; sa.U = b.U; sa.ld_u = h->pp;
    if (h->sp.active) { sa.U = b.R + h->sp.n_r; sa.ld_u = h->sp.n_cols; }   // u lives behind the column sums
    if (h->ldv == 64 && !getenv("VELA_B200_ECORR_SYRK_RECT")) {
        ecorr_syrk64_kernel<<<(unsigned)((B + 3) / 4), 128, 0, st>>>(sa);   // one warp per walker
    } else {
        const long warps = B * h->n_rect;
        ecorr_syrk_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(sa);
    }
    h->launches += 2;
    return (int)grid.x;
}

// Fills the launch arguments, plans (row-tile height, split-K) and launches K3; returns the split count K4 has to sum.
int launch_sigma(VelaPulsar* h, DeviceModel& d, long B, cudaStream_t st, SigmaArgs* out_args = nullptr) {
    const SigmaShape& sh = kSigmaShapes[h->sig_shape];
    Buffers& b = d.buf;
    const int nsamp = 8 * sh.nt;
    const SigmaPlan pl = plan_sigma(h, sh, b.sig_rows, B, d.n_sm);
    SigmaArgs sa;
    sa.M = d.M; sa.ldM = d.ldM; sa.rank = h->p; sa.n_pairs = h->n_pairs; sa.n_pair_units = pl.n_pair_units;
    sa.n_units = pl.n_units; sa.n_rows_pad = h->n_rows_pad; sa.W = b.W; sa.Y = b.Y; sa.ld_yw = h->n_rows_pad; sa.B = B;
    sa.Sig = b.Sig; sa.ld_sig = h->ld_sig; sa.U = b.U; sa.ld_u = h->pp;
    const long n_ktiles = h->n_rows_pad / sh.bk;
    sa.kt_per_split = (int)((n_ktiles + pl.ksplit - 1) / pl.ksplit);
    const int nz = (int)((n_ktiles + sa.kt_per_split - 1) / sa.kt_per_split);
    const long rows = round_up(B, kSigMaxSamples);
    sa.split_stride_sig = rows * h->ld_sig;
    sa.split_stride_u = rows * h->pp;
    dim3 grid((unsigned)((pl.n_units + sh.warps - 1) / sh.warps), (unsigned)((B + nsamp - 1) / nsamp), (unsigned)nz);
    (pl.mt == 4 ? sh.kernel4 : sh.kernel3)<<<grid, sh.warps * 32, h->sig_smem, st>>>(sa);
    if (out_args) *out_args = sa;
    return nz;
}

// Dynamic shared-memory ceilings are per FUNCTION, not per handle: raise every kernel to the device opt-in
// maximum once, so handles with different ranks cannot shrink each other's limit.
int set_smem_attrs(int dev) {
    int max_optin = 0;
    CU(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    max_optin -= 256;   // room for the static mbarrier words
    for (const SigmaShape& sh : kSigmaShapes) {
        CU(cudaFuncSetAttribute(sh.kernel4, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
        CU(cudaFuncSetAttribute(sh.kernel3, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    }
    for (const ColsumShape& sh : kColsumShapes)
        CU(cudaFuncSetAttribute(sh.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    for (const ColsumShape& sh : kColsumSplitShapes)
        CU(cudaFuncSetAttribute(sh.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    CU(cudaFuncSetAttribute(colsum_kernel<16, 2, 1, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    CU(cudaFuncSetAttribute(colsum_kernel<8, 4, 1, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    CU(cudaFuncSetAttribute(colsum_kernel<16, 2, 1, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    CU(cudaFuncSetAttribute(colsum_kernel<16, 4, 1, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    for (auto k : {chol_finish_kernel<8, true>, chol_finish_kernel<16, true>, chol_finish_kernel<16, false>})
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - 1024));
    // shared-memory matrices: many small walker CTAs per SM, so ask for the largest shared carve-out; the global-scratch
    // variant keeps the default split (its matrix traffic wants the L1)
    for (auto k : {chol_finish_kernel<8, true>, chol_finish_kernel<16, true>})
        CU(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - 1024));
    CU(cudaFuncSetAttribute(chol_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CU(cudaFuncSetAttribute(chol_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(ecorr_basis_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    CU(cudaFuncSetAttribute(ecorr_basis_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    CU(cudaFuncSetAttribute(ecorr_basis_kernel<40>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    CU(cudaFuncSetAttribute(ecorr_basis_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    CU(cudaFuncSetAttribute(toa_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
    return 0;
}

// K0 launch (read_params + hoists + TZR phase), per-model build when available
void launch_prologue(VelaPulsar* h, DeviceModel& d, const double* d_params, long B, cudaStream_t st) {
    PrologueArgs pa;
    pa.defaults = d.defaults; pa.free_idx = d.free_idx; pa.params = d_params; pa.B = B; pa.tzr_block = d.tzr_block;
    pa.index_masks = d.index_masks; pa.bit_masks = d.bit_masks; pa.n_toas = h->n_toas; pa.Pext = d.buf.Pext; pa.tzr_phase = d.buf.tzr;
    pa.tzr0 = h->delta_ready ? d.tzr0 : nullptr;
    int threads = B >= 4096 ? 128 : 32;          // small batches: spread the walkers over more SMs
    // working rows in shared memory while they fit the default 48 KB (config 3: 45 doubles per row); longer rows (free-
    // amplitude GPs with hundreds of hoisted factors) take narrower blocks, or the global row
    const size_t row_bytes = (size_t)h->prog.row * sizeof(double);
    if (threads * row_bytes > 48 * 1024) threads = 32;
    size_t smem = threads * row_bytes;
    if (smem > 48 * 1024 || (getenv("VELA_B200_PROLOGUE_SMEM") && atoi(getenv("VELA_B200_PROLOGUE_SMEM")) == 0)) smem = 0;
    pa.rows_in_smem = smem ? 1 : 0;
    const dim3 grid((unsigned)((B + threads - 1) / threads));
    if (h->use_spec && h->spec_prologue) {
        void* args[] = {&pa};
        cudaLaunchKernel((const void*)h->spec_prologue, grid, dim3((unsigned)threads), args, smem, st);
    } else {
        sample_prologue_kernel<<<grid, threads, smem, st>>>(h->prog, pa);
    }
}

// K1 launch: the NVRTC-specialised kernel of this model when it exists and is enabled, else the generic interpreter
void launch_chain(VelaPulsar* h, dim3 grid, size_t smem, cudaStream_t st, ChainArgs& ca) {
    ca.delta = (h->delta_ready && ca.emit != 2) ? 1 : 0;   // phase mode returns Double64 itself
    ca.n_index_masks = h->n_index_masks;
    if (h->use_spec && h->spec_chain && ca.emit == h->spec_emit) {   // other modes (single-sample residuals / phases): generic kernel
        void* args[] = {&ca};
        cudaLaunchKernel((const void*)h->spec_chain, grid, dim3((unsigned)h->chain_threads), args, smem, st);
    } else {
        toa_chain_kernel<<<grid, h->chain_threads, smem, st>>>(h->prog, ca);
    }
}

// Enqueue the whole pipeline for B walkers whose parameters already sit in d_params (row-major B x n_free).
// mode: 0 ln L, 1 chi^2 (white kernels only).  with_prior: combine with d_lnprior (may be null = 0).
int enqueue_batch(VelaPulsar* h, DeviceModel& d, const double* d_params, long B, int mode, int with_prior,
                  const double* d_lnprior, double* d_out, cudaStream_t st, double* d_ahat = nullptr,
                  double* d_cf = nullptr) {
    const ProgramDev& pg = h->prog;
    Buffers& b = d.buf;
    const bool timing = h->stage_timing && (&d == &h->devs[0]);
    if (timing) cudaEventRecord(d.ev[0], st);
    launch_prologue(h, d, d_params, B, st);
    if (timing) cudaEventRecord(d.ev[1], st);
    const bool woodbury = h->p > 0 && mode == 0;
    const bool ecorr = h->n_ecorr_groups > 0;
    ChainArgs ca;
    ca.Pext = b.Pext; ca.tzr_phase = b.tzr; ca.toa = d.toa; ca.toa_stride = d.toa_stride; ca.n_toas = h->n_toas;
    // walkers per chain CTA: 32 at full size; emcee-sized batches use smaller groups so the grid still covers the SMs
    int group = h->chain_group;
    while (group > 4 && (long)h->n_tiles * ((B + group - 1) / group) < 3L * d.n_sm) group /= 2;
    ca.index_masks = d.index_masks; ca.bit_masks = d.bit_masks; ca.B = B; ca.group = group;
    const bool wy = woodbury && h->sp.active;   // Y carries w.y for the structured contraction (the ECORR kernels take it too)
    ca.emit = wy ? 3 : ((woodbury || ecorr) ? 1 : 0); ca.Y = b.Y; ca.W = b.W; ca.F = nullptr; ca.ld_yw = h->n_rows_pad; ca.partial = b.partial;
    const int warps = h->chain_threads / 32;
    size_t chain_smem = chain_smem_doubles(group, pg.row, warps) * sizeof(double);
    dim3 cgrid((unsigned)h->n_tiles, (unsigned)((B + group - 1) / group));
    launch_chain(h, cgrid, chain_smem, st, ca);
    h->launches += 2;
    // Woodbury + active ECORR groups: the basis kernel of the correction (K2b) produces the scalar terms as well
    const bool scalars_from_k2b = woodbury && h->n_eg > 0;
    if (ecorr && !scalars_from_k2b) {
        EcorrArgs ea;
        ea.Pext = b.Pext; ea.row = pg.row; ea.Y = b.Y; ea.W = b.W; ea.ld_yw = h->n_rows_pad; ea.B = B;
        ea.groups = d.ecorr_groups; ea.chunk_start = d.ecorr_chunks; ea.par_ecorr = h->par_ecorr;
        ea.first_row = h->n_tiles; ea.partial = b.partial; ea.y_is_wy = wy ? 1 : 0;
        ea.n_chunks = h->n_ecorr_chunks;
        ecorr_group_kernel<<<dim3((unsigned)((B + 7) / 8), (unsigned)h->n_ecorr_rows), 256, 0, st>>>(ea);
        h->launches += 1;
    }
    if (timing) cudaEventRecord(d.ev[2], st);
    int n_rows_partial = h->n_tiles + (scalars_from_k2b ? 0 : h->n_ecorr_rows);
    if (!woodbury) {
        finish_white_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(b.partial, n_rows_partial, B, mode, d_lnprior,
                                                                         with_prior, d_out);
        h->launches += 1;
        if (timing) { cudaEventRecord(d.ev[3], st); cudaEventRecord(d.ev[4], st); }
    } else {
        SigmaArgs sa;
        ColsumArgs cs;
        int ksplit;
        if (h->sp.active) {
            if (wy && b.mu && B <= b.mu_rows && two_stage_taken(h, B)) {
                ksplit = launch_two_stage(h, d, B, st, &cs);
                h->launches -= 1;                            // (the caller counts one launch for this stage below)
            } else {
                ksplit = launch_colsum(h, d, B, st, wy, &cs);
            }
            const bool warp_on = getenv("VELA_B200_CHOL_WARP") && atoi(getenv("VELA_B200_CHOL_WARP")) == 1;
            const bool dmma_on = d.ftab && !(getenv("VELA_B200_CHOL_DMMA") && atoi(getenv("VELA_B200_CHOL_DMMA")) == 0);
            if (ksplit > 4 && ((warp_on && d.ltab) || (dmma_on && !warp_on)) && !d_ahat && h->n_eg == 0) {
                // emcee-sized batches split K deeply to fill the SMs; summing 16 - 64 slabs per entry would be the longest
                // serial stretch of the warp-per-walker finishing kernels (chol_dmma_kernel, chol_warp_kernel), so a wide
                // kernel does it first, in slice order (the CTA-per-walker kernel sums its slabs itself: 128 threads)
                const long n = B * (long)h->sp.n_cols;
                slab_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(b.R, cs.split_stride_r, ksplit, n);
                h->launches += 1;
                ksplit = 1;
            }
            if (h->n_eg > 0) {   // the ECORR correction is accumulated into a zeroed packed triangle that K4 adds
                cudaMemsetAsync(b.Sig, 0, (size_t)B * h->ld_sig * sizeof(double), st);
                n_rows_partial += launch_ecorr_correction(h, d, B, st, wy);
            }
        } else {
            ksplit = launch_sigma(h, d, B, st, &sa);
            if (h->n_eg > 0) n_rows_partial += launch_ecorr_correction(h, d, B, st, false);
        }
        if (timing) cudaEventRecord(d.ev[3], st);
        CholArgs ch;
        ch.Sig = b.Sig; ch.ld_sig = h->ld_sig; ch.ld_u = h->pp; ch.Sq = b.Sq; ch.U = b.U; ch.Pext = b.Pext; ch.row = pg.row; ch.p = h->p; ch.pp = h->pp; ch.n_gp = h->n_gp;
        ch.gp = d.gp; ch.gp_data = d.gp_data; ch.partial = b.partial; ch.n_tiles = n_rows_partial; ch.B = B;
        ch.lnprior = d_lnprior; ch.with_prior = with_prior; ch.out = d_out; ch.use_smem = h->chol_smem ? 1 : 0;
        ch.ksplit = ksplit;
        ch.table = nullptr; ch.R = nullptr; ch.ld_r = 0; ch.split_stride_r = 0; ch.n_r = 0; ch.u_off = 0; ch.add_sig = 0;
        ch.split_stride_sig = ch.split_stride_u = 0;
        if (h->sp.active) {
            ch.table = d.table; ch.R = b.R; ch.ld_r = cs.ld_r; ch.split_stride_r = cs.split_stride_r;
            ch.n_r = h->sp.n_r; ch.u_off = h->sp.n_r; ch.add_sig = h->n_eg > 0 ? 1 : 0;
        } else {
            ch.split_stride_sig = sa.split_stride_sig; ch.split_stride_u = sa.split_stride_u;
        }
        ch.ahat = d_ahat; ch.cf = d_cf;
        // ranks up to 63 on the structured path can run one warp per walker (chol_warp_kernel) instead of one CTA.
        // Measured on B200 at rank 60 (profiles/r2_chol_variants.md): 8192 walkers 0.30 ms either way (both are bound by
        // the dependent chain of a 60-column factorisation: 12 resident warps per SM at 36 % issue utilisation for the
        // warp kernel, 18 at 3.4 barrier stalls per issue for the CTA kernel); a lone walker 49 us (warp) against 35 us
        // (CTA, four warps share the assembly and the trailing updates).  The CTA kernel therefore stays the default and
        // VELA_B200_CHOL_WARP=1 selects the warp kernel (A/B measurements, tests).
        const bool warp_off = !(getenv("VELA_B200_CHOL_WARP") && atoi(getenv("VELA_B200_CHOL_WARP")) == 1);
        ch.ltab = d.ltab;
        ch.ctab = d.ctab;
        ch.ftab = d.ftab;
        // ranks up to 64 on the structured path: the tensor-core kernel with the matrix in registers (chol_dmma_kernel),
        // unless the conditional mean is wanted or an ECORR correction has to be added; VELA_B200_CHOL_DMMA=0 switches it off
        const bool dmma_off = getenv("VELA_B200_CHOL_DMMA") && atoi(getenv("VELA_B200_CHOL_DMMA")) == 0;
        if (d.ftab && !d_ahat && h->n_eg == 0 && !dmma_off && warp_off) {
            const int nbk = chol_dmma_nbk(h->p);
            const size_t sm = chol_dmma_smem_bytes(nbk, h->sp.n_r + h->p);
            (nbk == 2 ? chol_dmma_kernel<2> : nbk == 4 ? chol_dmma_kernel<4> : nbk == 6 ? chol_dmma_kernel<6> : chol_dmma_kernel<8>)
                <<<(unsigned)B, 32, sm, st>>>(ch);
        } else
        if (d.ltab && !d_ahat && !warp_off && h->n_eg == 0) {
            const int n_ru = h->sp.n_r + h->p;
            chol_warp_kernel<<<(unsigned)B, 32, chol_warp_smem_bytes(h->p, n_ru), st>>>(ch);
        } else if (!h->chol_smem && !d_ahat && !(getenv("VELA_B200_CHOL_PANEL") && atoi(getenv("VELA_B200_CHOL_PANEL")) == 0)) {
            // large ranks: 32-column panels in shared memory, trailing updates on DMMA (chol_panel_kernel)
            const int n_ru = h->sp.active ? h->sp.n_r : 0;
            chol_panel_kernel<<<(unsigned)B, kPanelThreadsHost, chol_panel_smem_bytes(h->p, h->pp, n_ru), st>>>(ch);
        } else {
        // 96 threads per walker at full size (more resident walkers, fewer idle warps in the serial phases: 0.41 ->
        // 0.35 ms at rank 60), 128 for small batches and large ranks (tools/probes/gpu_chol_probe.py)
        int chol_threads = (B >= 1024 && h->chol_nb == 8) ? 96 : kCholLaunchThreads;
        if (const char* e = getenv("VELA_B200_CHOL_THREADS")) chol_threads = std::max(h->chol_nb == 8 ? 64 : 32, std::min(h->chol_nb == 8 ? kCholLaunchThreads : kCholThreads, atoi(e) / 32 * 32));
        (h->chol_nb == 8 && h->chol_smem ? chol_finish_kernel<8, true> : (h->chol_smem ? chol_finish_kernel<16, true> : chol_finish_kernel<16, false>))<<<(unsigned)B, chol_threads, h->chol_smem_bytes, st>>>(ch);
        }
        if (timing) cudaEventRecord(d.ev[4], st);
        h->launches += 2;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(VELA_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
}

int build_device(VelaPulsar* h, DeviceModel& d, const VelaModelDesc* desc, const std::vector<double>& toa_soa,
                 const std::vector<double>& tzr_block, const std::vector<GPDev>& gps,
                 const std::vector<double>& gp_data, const std::vector<int3>& egroups, const std::vector<int>& echunks,
                 const std::vector<int3>& eactive) {
    CU(cudaSetDevice(d.dev));
    CU(cudaDeviceGetAttribute(&d.n_sm, cudaDevAttrMultiProcessorCount, d.dev));
    CU(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    for (auto& e : d.ev) CU(cudaEventCreate(&e));
    CU(cudaEventCreateWithFlags(&d.busy, cudaEventDisableTiming));
    if (upload(&d.toa, toa_soa.data(), toa_soa.size())) return VELA_ERR_CUDA;
    if (upload(&d.tzr_block, tzr_block.data(), tzr_block.size())) return VELA_ERR_CUDA;
    if (upload(&d.index_masks, desc->index_masks, (size_t)desc->n_index_masks * desc->n_toas)) return VELA_ERR_CUDA;
    if (upload(&d.bit_masks, desc->bit_masks, (size_t)desc->n_bit_mask_rows * desc->n_toas)) return VELA_ERR_CUDA;
    {   // default values followed by the aux table (read by the prologue only)
        std::vector<double> dv(desc->default_values, desc->default_values + desc->n_params);
        if (desc->n_aux > 0) dv.insert(dv.end(), desc->aux_values, desc->aux_values + desc->n_aux);
        if (upload(&d.defaults, dv.data(), dv.size())) return VELA_ERR_CUDA;
    }
    if (upload(&d.free_idx, desc->free_indices, (size_t)desc->n_free)) return VELA_ERR_CUDA;
    if (!egroups.empty()) {
        if (upload(&d.ecorr_groups, egroups.data(), egroups.size())) return VELA_ERR_CUDA;
        if (upload(&d.ecorr_chunks, echunks.data(), echunks.size())) return VELA_ERR_CUDA;
        if (upload(&d.ecorr_active, eactive.data(), eactive.size())) return VELA_ERR_CUDA;
    }
    if (h->p > 0) {
        // M: column-major with the leading dimension padded (zero rows) to n_rows_pad, columns padded to pp
        d.ldM = h->n_rows_pad;
        std::vector<double> Mp((size_t)d.ldM * h->p, 0.0);
        for (long c = 0; c < h->p; ++c)
            memcpy(&Mp[(size_t)c * d.ldM], desc->noise_basis + (size_t)c * desc->basis_ld, sizeof(double) * h->n_rows);
        if (upload(&d.M, Mp.data(), Mp.size())) return VELA_ERR_CUDA;
        if (upload(&d.gp, gps.data(), gps.size())) return VELA_ERR_CUDA;
        if (upload(&d.gp_data, gp_data.data(), gp_data.size())) return VELA_ERR_CUDA;
        if (h->sp.active) {
            if (upload(&d.T, h->sp.T.data(), h->sp.T.size())) return VELA_ERR_CUDA;
            if (upload(&d.table, h->sp.table.data(), h->sp.table.size())) return VELA_ERR_CUDA;
            if (h->sp.two_stage) {
                if (upload(&d.A1, h->sp.A1.data(), h->sp.A1.size())) return VELA_ERR_CUDA;
                if (upload(&d.T2, h->sp.T2.data(), h->sp.T2.size())) return VELA_ERR_CUDA;
                std::vector<ColsumUnit> units;
                for (const StructPlan::Stage2Job& job : h->sp.jobs)
                    for (int c = 0; c < job.n_out; c += 64)
                        units.push_back(ColsumUnit{job.t2_off + (long)c * h->sp.k2_pad, (long)job.block, job.r_off + c,
                                                   std::min(64, job.n_out - c)});
                d.n_units2 = (int)units.size();
                if (upload(&d.units2, units.data(), units.size())) return VELA_ERR_CUDA;
            }
            if (!h->ltab.empty() && upload(&d.ltab, h->ltab.data(), h->ltab.size())) return VELA_ERR_CUDA;
            if (!h->ftab.empty() && upload(&d.ftab, h->ftab.data(), h->ftab.size())) return VELA_ERR_CUDA;
            if (!h->ctab.empty() && upload(&d.ctab, h->ctab.data(), h->ctab.size())) return VELA_ERR_CUDA;
        }
    }
    if (set_smem_attrs(d.dev)) return VELA_ERR_CUDA;
    return 0;
}

void destroy_device(DeviceModel& d) {
    cudaSetDevice(d.dev);
    if (d.stream) cudaStreamSynchronize(d.stream);
    free_buffers(d.buf);
    if (d.buf.h_params) cudaFreeHost(d.buf.h_params);
    if (d.buf.h_out) cudaFreeHost(d.buf.h_out);
    if (d.buf.h_lnprior) cudaFreeHost(d.buf.h_lnprior);
    cudaFree(d.T); cudaFree(d.table); cudaFree(d.A1); cudaFree(d.T2); cudaFree(d.units2); cudaFree(d.ltab); cudaFree(d.ftab); cudaFree(d.ctab); cudaFree(d.tzr0);
    cudaFree(d.toa); cudaFree(d.tzr_block); cudaFree(d.index_masks); cudaFree(d.bit_masks); cudaFree(d.defaults);
    cudaFree(d.free_idx); cudaFree(d.M); cudaFree(d.gp); cudaFree(d.gp_data); cudaFree(d.ecorr_groups); cudaFree(d.ecorr_chunks); cudaFree(d.ecorr_active); cudaFree(d.priors);
    for (auto& e : d.ev) if (e) cudaEventDestroy(e);
    if (d.busy) cudaEventDestroy(d.busy);
    if (d.stream) cudaStreamDestroy(d.stream);
}

// Translation unit of the per-model chain kernel: the program as a constant object + the shared kernel body.
std::string spec_source(const ProgramDev& pg, int min_blocks = 2, int threads = kChainThreads, int emit = -1) {
    std::string src;
    char buf[256];
    // walkers in flight per thread (unroll factor of the walker loop; more instruction-level parallelism for more registers)
    if (const char* e = getenv("VELA_B200_JIT_UNROLL")) {
        snprintf(buf, sizeof buf, "#define VELA_CHAIN_UNROLL %d\n", std::max(1, std::min(4, atoi(e))));
        src += buf;
    }
    // fast-arithmetic build: speculative straight-line evaluation with a warp-uniform re-evaluation (Corr::spec)
    if ((pg.comp[0].flags & kFlagFast) && !(getenv("VELA_B200_JIT_SPECULATE") && atoi(getenv("VELA_B200_JIT_SPECULATE")) == 0))
        src += "#define VELA_SPEC_SPECULATE 1\n";
    {   // kFlagDelta in the program <=> create() filled the C_DELTA columns; residual-emitting batches then always take the route
        bool delta = false;
        for (int ic = 0; ic < pg.n_comp; ++ic) delta = delta || (pg.comp[ic].flags & kFlagDelta);
        src += (delta && emit >= 0 && emit != 2) ? "#define VELA_SPEC_DELTA 1\n" : "#define VELA_SPEC_DELTA 0\n";
    }
    if (emit >= 0) {   // the emit mode of the handle's batches, folded into the kernel (launch_chain checks it)
        snprintf(buf, sizeof buf, "#define VELA_SPEC_EMIT %d\n", emit);
        src += buf;
    }
    src += "#define VELA_SPEC_CALLS";
    for (int ic = 0; ic < pg.n_comp; ++ic) { snprintf(buf, sizeof buf, " VELA_APPLY(%d)", ic); src += buf; }
    src += "\n#include \"vela_chain.cuh\"\nnamespace vela {\n";
    snprintf(buf, sizeof buf, "__device__ const ProgramDev kSpecProg = {%d, %d, %d, %d, %d, %d, {\n", pg.n_comp, pg.n_params,
             pg.n_derived, pg.row, pg.wideband, pg.n_free);
    src += buf;
    for (int ic = 0; ic < pg.n_comp; ++ic) {
        const CompDev& c = pg.comp[ic];
        snprintf(buf, sizeof buf, "  {%d, %d, {", c.kind, c.flags);
        src += buf;
        for (int k = 0; k < VELA_MAX_PAR; ++k) { snprintf(buf, sizeof buf, "%d%s", c.par[k], k + 1 < VELA_MAX_PAR ? "," : ""); src += buf; }
        snprintf(buf, sizeof buf, "}, {%d,%d,%d,%d}, {%d,%d}, %d, {", c.len[0], c.len[1], c.len[2], c.len[3], c.mask[0], c.mask[1], c.der);
        src += buf;
        // hexadecimal floating literals: the struct constants round-trip bit for bit
        snprintf(buf, sizeof buf, "%a, %a}},\n", c.dval[0], c.dval[1]);
        src += buf;
    }
    src += "}};\n}  // namespace vela\n"
           "extern \"C\" __global__ void __launch_bounds__(";
    src += std::to_string(threads) + ", " + std::to_string(min_blocks);
    src += ") vela_chain_spec(vela::ChainArgs a) {\n    vela::chain_body(vela::kSpecProg, a);\n}\n"
           "extern \"C\" __global__ void __launch_bounds__(128) vela_prologue_spec(vela::PrologueArgs a) {\n"
           "    vela::prologue_body(vela::kSpecProg, a);\n}\n";
    return src;
}

int validate(const VelaModelDesc* d) {
    if (!d) return fail(VELA_ERR_INVALID_ARG, "null descriptor");
    if (d->abi_version != VELA_B200_ABI_VERSION) return fail(VELA_ERR_INVALID_ARG, "ABI version %d, expected %d", d->abi_version, VELA_B200_ABI_VERSION);
    if (d->n_components <= 0 || d->n_components > kMaxComponents) return fail(VELA_ERR_INVALID_ARG, "n_components %d outside 1..%d", d->n_components, kMaxComponents);
    if (d->n_toas <= 0 || !d->toas || !d->tzr_toa) return fail(VELA_ERR_INVALID_ARG, "missing TOAs / TZR TOA");
    if (d->toa_stride_bytes != (d->wideband ? VELA_WTOA_WORDS : VELA_TOA_WORDS) * 8) return fail(VELA_ERR_INVALID_ARG, "toa_stride_bytes %d does not match the %s record", d->toa_stride_bytes, d->wideband ? "WidebandTOA" : "TOA");
    if (d->n_params <= 0 || d->n_free < 0 || d->n_free > d->n_params || !d->default_values) return fail(VELA_ERR_INVALID_ARG, "bad parameter counts");
    for (int i = 0; i < d->n_free; ++i)
        if (d->free_indices[i] < 0 || d->free_indices[i] >= d->n_params) return fail(VELA_ERR_INVALID_ARG, "free index %d out of range", d->free_indices[i]);
    uint64_t tzr_index;
    memcpy(&tzr_index, (const double*)d->tzr_toa + 30, 8);
    if (tzr_index != 0) return fail(VELA_ERR_INVALID_ARG, "Invalid TZR TOA (index != 0), timing_model.jl:37");
    for (int ic = 0; ic < d->n_components; ++ic) {
        const VelaComponentDesc& c = d->components[ic];
        if (c.kind < VELA_COMP_SOLAR_SYSTEM || c.kind > VELA_COMP_KIND_MAX) return fail(VELA_ERR_UNSUPPORTED, "component kind %d is outside the implemented hot path", c.kind);
        for (int k = 0; k < VELA_MAX_PAR; ++k)
            if (c.par[k] >= d->n_params) return fail(VELA_ERR_INVALID_ARG, "component %d: parameter offset %d out of range", ic, c.par[k]);
        if (c.kind >= VELA_COMP_POWERLAW_RED_GP && c.kind <= VELA_COMP_POWERLAW_CHROM_GP) {
            if (c.len[0] <= 0 || c.len[1] < 0 || c.len[1] > c.len[0] || c.len[2] < 0 || !d->aux_values ||
                (int64_t)c.len[2] + 2 * (int64_t)c.len[0] > d->n_aux)
                return fail(VELA_ERR_INVALID_ARG, "component %d: power-law GP needs 2N aux values (ln_js, js)", ic);
            for (int k = 0; k < (c.kind == VELA_COMP_POWERLAW_CHROM_GP ? 7 : 6); ++k)
                if (c.par[k] < 0) return fail(VELA_ERR_INVALID_ARG, "component %d: power-law GP parameter slot %d unset", ic, k);
        }
        const bool bits = c.kind == VELA_COMP_PHASE_JUMP || c.kind == VELA_COMP_DM_JUMP || c.kind == VELA_COMP_DM_OFFSET || c.kind == VELA_COMP_FD_JUMP;
        for (int m = 0; m < 2; ++m) {
            if (c.mask[m] < 0) continue;
            if ((bits && m == 0) ? (c.mask[m] + c.len[0] > d->n_bit_mask_rows) : (c.mask[m] >= d->n_index_masks)) return fail(VELA_ERR_INVALID_ARG, "component %d: mask %d out of range", ic, c.mask[m]);
        }
    }
    if (d->kernel_kind < VELA_KERNEL_WHITE || d->kernel_kind > VELA_KERNEL_WOODBURY_ECORR) return fail(VELA_ERR_UNSUPPORTED, "kernel kind %d is outside the implemented hot path", d->kernel_kind);
    if (d->kernel_kind == VELA_KERNEL_ECORR || d->kernel_kind == VELA_KERNEL_WOODBURY_ECORR) {
        if (d->wideband) return fail(VELA_ERR_INVALID_ARG, "ECORR kernels are narrowband only");
        if (d->n_ecorr_groups <= 0 || !d->ecorr_groups) return fail(VELA_ERR_INVALID_ARG, "EcorrKernel without groups");
        uint64_t next = 1;
        for (int g = 0; g < d->n_ecorr_groups; ++g) {
            const uint64_t* e = d->ecorr_groups + 3 * g;
            if (e[0] != next || e[1] < e[0] || e[1] > (uint64_t)d->n_toas) return fail(VELA_ERR_INVALID_ARG, "ECORR group %d does not continue the TOA partition", g);
            if (e[2] != 0 && (d->par_ecorr < 0 || d->par_ecorr + (int64_t)e[2] > d->n_params)) return fail(VELA_ERR_INVALID_ARG, "ECORR group %d: bad ECORR index", g);
            next = e[1] + 1;
        }
        if (next != (uint64_t)d->n_toas + 1) return fail(VELA_ERR_INVALID_ARG, "ECORR groups do not cover all TOAs");
    }
    if (d->kernel_kind == VELA_KERNEL_WOODBURY_WHITE || d->kernel_kind == VELA_KERNEL_WOODBURY_ECORR) {
        long nr = d->wideband ? 2 * d->n_toas : d->n_toas;
        if (d->basis_rows != nr || d->basis_cols <= 0 || d->basis_ld < nr || !d->noise_basis) return fail(VELA_ERR_INVALID_ARG, "noise_basis shape (%ld x %ld, ld %ld) does not match %ld rows", (long)d->basis_rows, (long)d->basis_cols, (long)d->basis_ld, nr);
        if (d->basis_cols > kSigMaxRank) return fail(VELA_ERR_UNSUPPORTED, "basis rank %ld > %d", (long)d->basis_cols, kSigMaxRank);
        long cols = 0;
        for (int g = 0; g < d->n_gp; ++g) cols += d->gp[g].kind == VELA_GP_MARGINALIZED_TM ? d->gp[g].n : 2 * d->gp[g].n;
        if (cols != d->basis_cols) return fail(VELA_ERR_INVALID_ARG, "GP components describe %ld columns, basis has %ld (kernel.jl:85)", cols, (long)d->basis_cols);
    }
    return 0;
}

}  // namespace

extern "C" {

VELA_API int vela_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

VELA_API const char* vela_b200_last_error(void) { return g_last_error.c_str(); }

VELA_API int vela_b200_create(const VelaModelDesc* desc, VelaHandle* out) {
    if (!out) return fail(VELA_ERR_INVALID_ARG, "null output handle");
    *out = nullptr;
    if (int rc = validate(desc)) return rc;
    int ndev_avail = vela_b200_device_count();
    if (ndev_avail <= 0) return fail(VELA_ERR_NO_DEVICE, "no CUDA device available; libvela_b200 has no CPU path");

    VelaPulsar* h = new VelaPulsar();
    ProgramDev& pg = h->prog;
    build_program(desc, pg);

    h->kernel_kind = desc->kernel_kind;
    h->n_toas = desc->n_toas;
    h->n_index_masks = desc->n_index_masks;
    h->n_rows = desc->wideband ? 2 * desc->n_toas : desc->n_toas;
    h->n_rows_pad = round_up(h->n_rows, 64);
    h->chain_threads = desc->n_toas <= 64 ? 64 : 128;   // 128-thread blocks x 5 resident (tools/probes/gpu_chain_threads_probe.py)
    if (const char* e = getenv("VELA_B200_CHAIN_THREADS")) h->chain_threads = std::max(32, std::min(256, atoi(e) / 32 * 32));
    h->n_tiles = (int)((desc->n_toas + h->chain_threads - 1) / h->chain_threads);
    h->chain_group = 32;
    if (const char* e = getenv("VELA_B200_CHAIN_GROUP")) h->chain_group = std::max(4, std::min(64, atoi(e)));   // <= 64: the speculative build keeps one redo bit per walker of the group

    std::vector<int3> egroups;
    std::vector<int> echunks;
    std::vector<int3> eactive;
    if (desc->kernel_kind == VELA_KERNEL_ECORR || desc->kernel_kind == VELA_KERNEL_WOODBURY_ECORR) {
        // chunks of 32 consecutive groups (one lane per group in ecorr_group_kernel; groups without ECORR, index 0,
        // need no correction and idle their lane); kEcorrChunksPerRow chunks share a `partial` row
        h->par_ecorr = desc->par_ecorr;
        echunks.push_back(0);
        for (int g = 0; g < desc->n_ecorr_groups; ++g) {
            const uint64_t* e = desc->ecorr_groups + 3 * g;
            egroups.push_back(make_int3((int)(e[0] - 1), (int)e[1], (int)e[2]));
            if (e[2]) eactive.push_back(egroups.back());
            if (g + 1 - echunks.back() >= 32) echunks.push_back(g + 1);
        }
        if (echunks.back() != desc->n_ecorr_groups) echunks.push_back(desc->n_ecorr_groups);
        h->n_ecorr_groups = desc->n_ecorr_groups;
        h->n_ecorr_chunks = (int)echunks.size() - 1;
        h->n_ecorr_rows = (h->n_ecorr_chunks + kEcorrChunksPerRow - 1) / kEcorrChunksPerRow;
    }
    std::vector<GPDev> gps;
    std::vector<double> gp_data;
    if (desc->kernel_kind == VELA_KERNEL_WOODBURY_WHITE || desc->kernel_kind == VELA_KERNEL_WOODBURY_ECORR) {
        h->p = (int)desc->basis_cols;
        if (desc->kernel_kind == VELA_KERNEL_WOODBURY_ECORR && !eactive.empty()) {
            h->n_eg = (int)eactive.size();
            h->eg_pad = round_up(h->n_eg, 4);
            const int nr = (h->p + 1 + 31) / 32;      // 32-column rectangles over [M | r]
            h->ldv = (long)nr * 32;
            h->n_rect = nr * (nr + 1) / 2;
        }
        h->pp = (h->p + 7) / 8 * 8;
        h->n_gp = desc->n_gp;
        h->n_pairs = (long)h->p * (h->p + 1) / 2;
        h->ld_sig = round_up(h->n_pairs, 8);
        analyse_basis(desc, h->n_rows, h->n_rows_pad, h->p, h->sp);
        // pick the first launch shape whose ring fits (two CTAs per SM need <= 110 KB each); the environment
        // variable VELA_B200_SIGMA_SHAPE=<index into kSigmaShapes> overrides for A/B measurements
        const int n_shapes = (int)(sizeof(kSigmaShapes) / sizeof(kSigmaShapes[0]));
        h->sig_shape = -1;
        for (int i = 0; i < n_shapes; ++i) {
            const SigmaShape& sh = kSigmaShapes[i];
            size_t need = sigma_smem_bytes(h->p, sh.bk, sh.stages, sh.nt);
            if (need <= (size_t)(sh.ctas_per_sm == 2 ? 110 : 220) * 1024) { h->sig_shape = i; break; }
        }
        if (const char* e = getenv("VELA_B200_SIGMA_SHAPE")) {
            int i = atoi(e);
            if (i >= 0 && i < n_shapes && sigma_smem_bytes(h->p, kSigmaShapes[i].bk, kSigmaShapes[i].stages, kSigmaShapes[i].nt) <= (size_t)220 * 1024)
                h->sig_shape = i;
        }
        if (h->sig_shape < 0) { delete h; return fail(VELA_ERR_UNSUPPORTED, "basis rank %d does not fit any Sigma-contraction shape", h->p); }
        const SigmaShape& shp = kSigmaShapes[h->sig_shape];
        h->sig_smem = sigma_smem_bytes(h->p, shp.bk, shp.stages, shp.nt);
        int col = 0;
        for (int g = 0; g < desc->n_gp; ++g) {
            const VelaGPDesc& s = desc->gp[g];
            GPDev gd;
            gd.kind = s.kind; gd.n = s.n; gd.par_log10_amp = s.par_log10_amp; gd.par_gamma = s.par_gamma; gd.par_f1 = s.par_f1;
            gd.col0 = col; gd.data0 = (int)gp_data.size();
            const double* src = s.kind == VELA_GP_MARGINALIZED_TM ? s.weights_inv : s.ln_js;
            if (!src || s.n <= 0) { delete h; return fail(VELA_ERR_INVALID_ARG, "GP component %d has no data", g); }
            if (s.kind != VELA_GP_MARGINALIZED_TM && (s.par_log10_amp < 0 || s.par_gamma < 0 || s.par_f1 < 0 ||
                                                      s.par_log10_amp >= desc->n_params || s.par_gamma >= desc->n_params || s.par_f1 >= desc->n_params)) {
                delete h;
                return fail(VELA_ERR_INVALID_ARG, "GP component %d: bad parameter offsets", g);
            }
            gp_data.insert(gp_data.end(), src, src + s.n);
            col += s.kind == VELA_GP_MARGINALIZED_TM ? s.n : 2 * s.n;
            gps.push_back(gd);
        }
        h->chol_nb = h->p <= kCholSmallRank ? 8 : 16;
        if (const char* e = getenv("VELA_B200_CHOL_NB")) h->chol_nb = atoi(e) == 16 ? 16 : 8;
        const int n_r = h->sp.active ? h->sp.n_r : 0;
        size_t need = chol_smem_doubles(h->p, h->pp, h->chol_nb, true, n_r) * sizeof(double);
        h->chol_smem = need <= 200 * 1024;
        if (const char* e = getenv("VELA_B200_CHOL_SMEM")) { if (atoi(e) == 0) h->chol_smem = false; }   // tests: force the global-scratch kernels
        if (!h->chol_smem) h->chol_nb = 16;   // the global-scratch variant exists for block width 16 only
        h->chol_smem_bytes = h->chol_smem ? need : chol_smem_doubles(h->p, h->pp, h->chol_nb, false, n_r) * sizeof(double);
        // structured path, large ranks (chol_panel_kernel): the pair table again as 4 bytes per pair - two 16-bit indices
        // (i1 | i2 << 16) into the walker's pre-scaled sums S = {+R/2 | -R/2 | 0}: the entry is S[i1] + S[i2] (R1 = R1/2 +
        // R1/2, R1/2 = R1/2 + 0, (R1 +- R2)/2), bit for bit the value c1 R1 + c2 R2 of the 24-byte table, for 1/6 of the
        // bytes the finishing kernel pulls through L2 per walker (rank 300: 45 150 pairs) and one add per entry
        if (h->sp.active && 2 * (long)h->sp.n_r + 1 <= 65535) {
            const unsigned n_r = (unsigned)h->sp.n_r, zero = 2 * n_r;
            h->ctab.resize(h->sp.table.size());
            for (size_t t = 0; t < h->sp.table.size(); ++t) {
                const PairTerm& pt = h->sp.table[t];
                unsigned i2;
                if (pt.c1 == 1.0 && pt.c2 == 0.0) i2 = (unsigned)pt.i1;
                else if (pt.c1 == 0.5 && pt.c2 == 0.0) i2 = zero;
                else if (pt.c1 == 0.5 && pt.c2 == 0.5) i2 = (unsigned)pt.i2;
                else if (pt.c1 == 0.5 && pt.c2 == -0.5) i2 = n_r + (unsigned)pt.i2;
                else { h->ctab.clear(); break; }                    // a coefficient this form cannot carry
                h->ctab[t] = (unsigned)pt.i1 | (i2 << 16);
            }
        }
        // ranks up to 63 on the structured path: layout table of the warp-per-walker Cholesky kernel - for every
        // shared-memory slot of its column-blocked matrix the recipe that fills it (zero | c1 R[i1] + c2 R[i2] | u_i)
        if (h->sp.active && h->p + 1 <= kCholWarpMaxRows && h->sp.n_cols <= kCholWarpMaxCols) {
            const int n1 = h->p + 1, NR = (n1 + 1) & ~1, nblk = (n1 + 3) / 4;
            h->ltab.assign((size_t)cw_col_base(nblk, NR), 0u);
            bool ok = true;
            for (int r = 0; r < h->p && ok; ++r)
                for (int c = 0; c <= r; ++c) {
                    const PairTerm& pt = h->sp.table[(size_t)r * (r + 1) / 2 + c];
                    const bool c1_one = pt.c1 == 1.0, c1_half = pt.c1 == 0.5;
                    const int c2 = pt.c2 == 0.0 ? 0 : (pt.c2 == 0.5 ? 1 : (pt.c2 == -0.5 ? 2 : -1));
                    if (!(c1_one || c1_half) || c2 < 0) { ok = false; break; }   // a coefficient the 4-byte code cannot carry
                    h->ltab[cw_slot(r, c, NR)] = (1u << 29) | ((unsigned)c2 << 27) | ((c1_one ? 1u : 0u) << 26) |
                                                 ((unsigned)(c2 ? pt.i2 : 0) << 13) | (unsigned)pt.i1;
                }
            for (int i = 0; i < h->p; ++i) h->ltab[cw_slot(h->p, i, NR)] = (2u << 29) | (unsigned)(h->sp.n_r + i);
            if (!ok) h->ltab.clear();
        }
        // ranks up to 64 on the structured path: fragment table of the tensor-core Cholesky kernel - for every block
        // (36 of Sigma at 8 block rows, then the u row's 8), register and lane, two 16-bit indices (i1 | i2 << 16) into
        // the walker's pre-scaled sums S = {-R/2, -u/2 | +R/2, +u/2 | 0 | -1/2}: the (negated) entry is S[i1] + S[i2].
        // lane l = 4 r + t, register e  <->  row pi(r), column 4 e + t of the block (chol_dmma_kernel, vela_kernels.cu);
        // entries are grouped in fours per lane (one 16-byte load): word 4 g + q of lane l sits at (g * 32 + l) * 4 + q.
        if (h->sp.active && h->p <= kCholDmmaMaxRank && 2 * (h->sp.n_r + h->p) + 2 <= 65535) {
            const int nbk = chol_dmma_nbk(h->p), nt0 = nbk * (nbk + 1) / 2, nt = nt0 + nbk;
            const unsigned n_ru = (unsigned)(h->sp.n_r + h->p), zero = 2 * n_ru, half_one = 2 * n_ru + 1;
            h->ftab.assign((size_t)((2 * nt + 3) / 4) * 128, zero | (zero << 16));
            bool ok = true;
            for (int I = 0; I <= nbk && ok; ++I)
                for (int J = 0; J <= (I < nbk ? I : nbk - 1) && ok; ++J) {
                    const int blk = I < nbk ? I * (I + 1) / 2 + J : nt0 + J;
                    for (int e = 0; e < 2 && ok; ++e)
                        for (int l = 0; l < 32; ++l) {
                            const int rr = cd_pi(l >> 2), cc = 4 * e + (l & 3);
                            unsigned i1 = zero, i2 = zero;
                            if (I < nbk) {
                                const int row = 8 * I + rr, col = 8 * J + cc;
                                if (row < h->p && col < h->p) {
                                    const int hi = std::max(row, col), lo = std::min(row, col);
                                    const PairTerm& pt = h->sp.table[(size_t)hi * (hi + 1) / 2 + lo];
                                    i1 = (unsigned)pt.i1;
                                    if (pt.c1 == 1.0 && pt.c2 == 0.0) i2 = i1;                     // R1 = (R1 + R1) / 2
                                    else if (pt.c1 == 0.5 && pt.c2 == 0.0) i2 = zero;
                                    else if (pt.c1 == 0.5 && pt.c2 == 0.5) i2 = (unsigned)pt.i2;
                                    else if (pt.c1 == 0.5 && pt.c2 == -0.5) i2 = n_ru + (unsigned)pt.i2;
                                    else { ok = false; break; }                                    // a coefficient this form cannot carry
                                } else if (row == col) {
                                    i1 = i2 = half_one;                                            // identity padding past the rank
                                }
                            } else if (rr == 0 && 8 * J + cc < h->p) {
                                i1 = i2 = (unsigned)(h->sp.n_r + 8 * J + cc);                      // u^T: row 0 of the last block row
                            }
                            const int word = blk * 2 + e;
                            h->ftab[(size_t)((word >> 2) * 32 + l) * 4 + (word & 3)] = i1 | (i2 << 16);
                        }
                }
            if (!ok) h->ftab.clear();
        }
    }

    // SoA TOA table (+ TOA-only hoists) and the TZR block
    long stride = round_up(desc->n_toas, 32);
    std::vector<double> soa((size_t)C_NCOL * stride, 0.0);
    for (long j = 0; j < desc->n_toas; ++j) {
        fill_toa_columns((const double*)((const char*)desc->toas + j * (long)desc->toa_stride_bytes), desc->wideband, soa.data(), stride, j);
        fill_shapiro_columns(desc, soa.data(), stride, j);
    }
    std::vector<double> tzr(C_NCOL, 0.0);
    fill_toa_columns((const double*)desc->tzr_toa, 0, tzr.data(), 1, 0);
    fill_shapiro_columns(desc, tzr.data(), 1, 0);

    std::vector<int> devices;
    if (desc->n_devices > 0 && desc->devices) devices.assign(desc->devices, desc->devices + desc->n_devices);
    else {
        int cur = 0;
        cudaGetDevice(&cur);
        devices.push_back(cur);
    }
    for (int dv : devices)
        if (dv < 0 || dv >= ndev_avail) { delete h; return fail(VELA_ERR_INVALID_ARG, "device %d not available (%d devices)", dv, ndev_avail); }
    int prev = 0;
    cudaGetDevice(&prev);
    h->devs.resize(devices.size());
    for (size_t i = 0; i < devices.size(); ++i) {
        h->devs[i].dev = devices[i];
        h->devs[i].toa_stride = stride;
        int rc = build_device(h, h->devs[i], desc, soa, tzr, gps, gp_data, egroups, echunks, eactive);
        if (rc) {
            for (auto& d : h->devs) destroy_device(d);
            delete h;
            cudaSetDevice(prev);
            return rc;
        }
    }
    cudaSetDevice(prev);
    if (h->devs.size() > 1)
        for (size_t i = 0; i < h->devs.size(); ++i) { h->feeders.emplace_back(new Feeder()); h->feeders.back()->start(); }
    std::vector<double>().swap(h->sp.T);   // the device copies are authoritative from here on
    std::vector<double>().swap(h->sp.A1);
    std::vector<double>().swap(h->sp.T2);

    // kFlagDelta: the chain at the default parameters fills the C_DELTA columns of every device's TOA table
    {
        bool want = false;
        for (int ic = 0; ic < pg.n_comp; ++ic) want = want || (pg.comp[ic].flags & kFlagDelta);
        bool ok = want;
        std::vector<double> x0((size_t)std::max(desc->n_free, 1), 0.0);
        for (int i = 0; i < desc->n_free; ++i) x0[i] = desc->default_values[desc->free_indices[i]];
        for (size_t i = 0; ok && i < h->devs.size(); ++i) {
            DeviceModel& d = h->devs[i];
            cudaSetDevice(d.dev);
            ok = ensure_device_buffers(h, d, 1) == 0 && cudaMalloc((void**)&d.tzr0, 2 * sizeof(double)) == cudaSuccess;
            if (!ok) break;
            cudaMemcpyAsync(d.buf.params, x0.data(), sizeof(double) * x0.size(), cudaMemcpyHostToDevice, d.stream);
            launch_prologue(h, d, d.buf.params, 1, d.stream);      // (delta_ready is still false: plain rows)
            default_phase_kernel<<<(unsigned)((h->n_toas + 127) / 128), 128, 0, d.stream>>>(
                h->prog, d.buf.Pext, d.toa, d.toa_stride, h->n_toas, d.tzr_block, d.index_masks, d.bit_masks, d.tzr0);
            ok = cudaStreamSynchronize(d.stream) == cudaSuccess;
        }
        cudaSetDevice(prev);
        if (want && !ok) { cudaGetLastError(); for (int ic = 0; ic < pg.n_comp; ++ic) h->prog.comp[ic].flags &= ~kFlagDelta; }
        h->delta_ready = want && ok;
    }

    // per-model chain kernel (vela_jit.cu); any failure leaves the generic kernel in charge
    {
        const char* e = getenv("VELA_B200_JIT");
        if (!(e && atoi(e) == 0)) {
            cudaKernel_t k = nullptr;
            // five resident 128-thread blocks per SM (<= 102 registers, 20 warps): measured best among 64..256 threads x
            // 2..10 blocks on the models probed (tools/probes/gpu_chain_threads_probe.py)
            int mb = 5;
            if (const char* force = getenv("VELA_B200_JIT_MINBLOCKS")) mb = std::max(1, std::min(16, atoi(force)));
            cudaKernel_t kp = nullptr;
            // batches of this handle always run one emit mode: w.y / w for the structured contraction, y / w for the dense
            // one and the ECORR kernel, nothing for the white kernel
            h->spec_emit = (h->p > 0 && h->sp.active) ? 3 : ((h->p > 0 || h->n_ecorr_groups > 0) ? 1 : 0);
            int rc_jit = jit_chain_kernel(spec_source(pg, mb, std::max(h->chain_threads, 128), h->spec_emit), &k, &kp, h->jit_log);
            if (rc_jit == 0) {
                bool ok = true;
                for (auto& d : h->devs) {
                    cudaSetDevice(d.dev);
                    int max_optin = 0;
                    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.dev);
                    if (cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - 256) != cudaSuccess) {
                        h->jit_log = std::string("cudaFuncSetAttribute on the specialised kernel: ") + cudaGetErrorString(cudaGetLastError());
                        ok = false;
                    }
                }
                cudaSetDevice(prev);
                if (ok) { h->spec_chain = k; h->spec_prologue = kp; h->use_spec = true; }
            }
            if (!h->use_spec && !getenv("VELA_B200_QUIET")) {
                static std::atomic<bool> warned{false};
                if (!warned.exchange(true))
                    fprintf(stderr, "vela_b200: no per-model chain kernel (%s); the generic interpreting kernel runs instead, "
                                    "~2.8x slower on the chain stage (vela_b200_jit_log has the details)\n",
                            h->jit_log.substr(0, 200).c_str());
            }
        } else {
            h->jit_log = "disabled by VELA_B200_JIT=0";
        }
    }

    // walkers per device per pass: bound the y/w/Sigma scratch to ~1/4 of the device memory
    if (h->kernel_kind != VELA_KERNEL_WHITE) {
        size_t per = (size_t)2 * h->n_rows_pad * 8 + (h->sp.active ? (size_t)h->sp.n_cols * 32 + (h->n_eg ? (size_t)h->ld_sig * 8 : 0) : (size_t)h->ld_sig * 8) + (size_t)h->pp * 8 + (h->chol_smem ? 0 : chol_panel_scratch_doubles(h->p, h->pp) * 8) + (size_t)h->eg_pad * h->ldv * 8 +
                     (h->sp.two_stage ? (size_t)h->sp.n_blocks * h->sp.k2_pad * 8 : 0);
        size_t free_b = 0, total_b = 0;
        cudaSetDevice(devices[0]);
        cudaMemGetInfo(&free_b, &total_b);
        cudaSetDevice(prev);
        size_t budget = std::min<size_t>(free_b / 2, (size_t)48 << 30);
        long mc = (long)(budget / std::max<size_t>(per, 1));
        h->max_chunk = std::max<long>(kSigMaxSamples, mc / kSigMaxSamples * kSigMaxSamples);
    } else {
        h->max_chunk = 1 << 20;
    }
    if (const char* e = getenv("VELA_B200_MAX_CHUNK"))   // test hook: force the multi-pass path with a small chunk
        h->max_chunk = std::max<long>(kSigMaxSamples, std::min<long>(h->max_chunk, round_up(atol(e), kSigMaxSamples)));
    *out = h;
    return VELA_OK;
}

VELA_API int vela_b200_destroy(VelaHandle h) {
    if (!h) return VELA_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    for (auto& f : h->feeders) f->shutdown();
    for (auto& d : h->devs) destroy_device(d);
    cudaSetDevice(prev);
    delete h;
    return VELA_OK;
}

VELA_API int vela_b200_n_free(VelaHandle h) { return h ? h->prog.n_free : -1; }
VELA_API int64_t vela_b200_n_rows(VelaHandle h) { return h ? h->n_rows : -1; }
VELA_API int64_t vela_b200_n_basis(VelaHandle h) { return h ? h->p : -1; }
VELA_API int64_t vela_b200_launch_count(VelaHandle h) { return h ? (int64_t)h->launches.load() : -1; }

static int run_host_batch(VelaHandle h, const double* params, int64_t B, int64_t n_free, int64_t rs, int64_t cs,
                          const double* lnprior, int with_prior, int mode, double* out) {
    if (!h) return fail(VELA_ERR_INVALID_ARG, "null handle");
    if (n_free != h->prog.n_free) return fail(VELA_ERR_INVALID_ARG, "expected %d free parameters, got %ld (parameter.jl:162)", h->prog.n_free, (long)n_free);
    if (B < 0 || (B > 0 && (!params || !out))) return fail(VELA_ERR_INVALID_ARG, "null params/out");
    if (mode == 1 && h->p > 0) return fail(VELA_ERR_UNSUPPORTED, "chi2 is defined for white-noise and ECORR kernels only (wls_chi2.jl:44, ecorr_chi2.jl:58)");
    if (B == 0) return VELA_OK;
    std::lock_guard<std::mutex> lock(h->mtx);
    int prev = 0;
    cudaGetDevice(&prev);
    const int nd = (int)h->devs.size();
    const long nf = h->prog.n_free;
    // contiguous walker blocks per device, processed in passes of at most max_chunk
    long done = 0;
    int rc = 0;
    while (done < B && !rc) {
        long pass = std::min<long>(B - done, h->max_chunk * nd);
        long per = (pass + nd - 1) / nd;
        std::vector<long> r0(nd), r1(nd);
        for (int i = 0; i < nd; ++i) { r0[i] = done + std::min<long>(pass, i * per); r1[i] = done + std::min<long>(pass, (i + 1) * per); }
        // one device: this thread; several: one host thread per device, so the pinned staging copies (1.3 MB per device
        // at config 3, ~0.1 ms each) run side by side instead of delaying device i by i copies
        auto feed = [&](int i) -> int {
            int rc = 0;
            DeviceModel& d = h->devs[i];
            long n = r1[i] - r0[i];
            if (n <= 0) return 0;
            cudaSetDevice(d.dev);
            if ((rc = ensure_host_buffers(h, d, n))) return rc;
            if ((rc = ensure_device_buffers(h, d, n))) return rc;
            Buffers& b = d.buf;
            const bool timing = h->stage_timing && i == 0;
            static const bool graphs_off = getenv("VELA_B200_GRAPHS") && atoi(getenv("VELA_B200_GRAPHS")) == 0;
            if (nd == 1 && n <= kGraphMaxBatch && !timing && h->graphs_ok && !graphs_off) {
                // ---- small batch: pack, then replay the captured graph of the whole step
                if (rs == nf && cs == 1) memcpy(b.h_params, params + r0[i] * rs, sizeof(double) * n * nf);
                else
                    for (long r = 0; r < n; ++r)
                        for (long k = 0; k < nf; ++k) b.h_params[r * nf + k] = params[(r0[i] + r) * rs + k * cs];
                const int prior_src = !with_prior ? 0 : (lnprior ? 1 : (h->have_priors ? 2 : 3));
                if (prior_src == 1) memcpy(b.h_lnprior, lnprior + r0[i], sizeof(double) * n);
                // the kernel-variant switch read at enqueue time is part of the key (A/B tests flip it between calls)
                const int variant = ((getenv("VELA_B200_CHOL_WARP") && atoi(getenv("VELA_B200_CHOL_WARP")) == 1) ? 16 : 0) |
                                    ((getenv("VELA_B200_CHOL_DMMA") && atoi(getenv("VELA_B200_CHOL_DMMA")) == 0) ? 32 : 0) |
                                    ((getenv("VELA_B200_CHOL_PANEL") && atoi(getenv("VELA_B200_CHOL_PANEL")) == 0) ? 64 : 0) |
                                    (two_stage_taken(h, n) ? 128 : 0);
                auto key = std::make_tuple(n, mode, with_prior, prior_src | variant);
                auto it = b.graphs.find(key);
                if (it == b.graphs.end()) {
                    Buffers::GraphEntry ge;
                    const int64_t l0 = h->launches.load();
                    cudaGraph_t graph = nullptr;
                    cudaError_t ge_rc = cudaStreamBeginCapture(d.stream, cudaStreamCaptureModeThreadLocal);
                    int erc = 0;
                    if (ge_rc == cudaSuccess) {
                        if (nf > 0) cudaMemcpyAsync(b.params, b.h_params, sizeof(double) * n * nf, cudaMemcpyHostToDevice, d.stream);
                        const double* g_lp = nullptr;
                        if (prior_src == 2) {
                            prior_kernel<<<(unsigned)((n + 127) / 128), 128, 0, d.stream>>>(d.priors, (int)nf, b.params, n, 0, b.lnprior);
                            h->launches += 1;
                            g_lp = b.lnprior;
                        } else if (prior_src == 1) {
                            cudaMemcpyAsync(b.lnprior, b.h_lnprior, sizeof(double) * n, cudaMemcpyHostToDevice, d.stream);
                            g_lp = b.lnprior;
                        }
                        erc = enqueue_batch(h, d, b.params, n, mode, with_prior, g_lp, b.out, d.stream);
                        cudaMemcpyAsync(b.h_out, b.out, sizeof(double) * n, cudaMemcpyDeviceToHost, d.stream);
                        ge_rc = cudaStreamEndCapture(d.stream, &graph);
                    }
                    if (ge_rc == cudaSuccess && !erc && graph) ge_rc = cudaGraphInstantiate(&ge.exec, graph, 0);
                    if (graph) cudaGraphDestroy(graph);
                    ge.launches = (int)(h->launches.load() - l0);
                    h->launches -= ge.launches;              // counted per replay below
                    if (ge_rc != cudaSuccess || erc || !ge.exec) {
                        cudaGetLastError();
                        h->graphs_ok = false;                // the plain path below takes over, for good
                    } else {
                        it = b.graphs.emplace(key, ge).first;
                    }
                }
                if (it != b.graphs.end()) {
                    scratch_acquire(d, d.stream);
                    cudaError_t le = cudaGraphLaunch(it->second.exec, d.stream);
                    scratch_release(d, d.stream);
                    h->launches += it->second.launches;
                    if (le != cudaSuccess) return fail(VELA_ERR_CUDA, "device %d: graph launch failed: %s", d.dev, cudaGetErrorString(le));
                    return 0;
                }
            }
            scratch_acquire(d, d.stream);
            if (timing) cudaEventRecord(d.ev[5], d.stream);
            cudaError_t ce = cudaSuccess;
            auto copy = [&](void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
                if (ce == cudaSuccess) ce = cudaMemcpyAsync(dst, src, bytes, kind, d.stream);
            };
            // staged through pinned memory in slices, so the copy engine works on slice k while the host packs k + 1
            const long slice = std::max<long>(256, (256 * 1024) / (8 * std::max<long>(nf, 1)));
            for (long s0 = 0; s0 < n; s0 += slice) {
                const long m = std::min(slice, n - s0);
                if (rs == nf && cs == 1) memcpy(b.h_params + s0 * nf, params + (r0[i] + s0) * rs, sizeof(double) * m * nf);
                else
                    for (long r = s0; r < s0 + m; ++r)
                        for (long k = 0; k < nf; ++k) b.h_params[r * nf + k] = params[(r0[i] + r) * rs + k * cs];
                if (nf > 0) copy(b.params + s0 * nf, b.h_params + s0 * nf, sizeof(double) * m * nf, cudaMemcpyHostToDevice);
            }
            const double* d_lp = nullptr;
            if (with_prior && !lnprior && h->have_priors) {
                prior_kernel<<<(unsigned)((n + 127) / 128), 128, 0, d.stream>>>(d.priors, (int)nf, b.params, n, 0, b.lnprior);
                h->launches += 1;
                d_lp = b.lnprior;
            } else if (with_prior && lnprior) {
                memcpy(b.h_lnprior, lnprior + r0[i], sizeof(double) * n);
                copy(b.lnprior, b.h_lnprior, sizeof(double) * n, cudaMemcpyHostToDevice);
                d_lp = b.lnprior;
            }
            rc = enqueue_batch(h, d, b.params, n, mode, with_prior, d_lp, b.out, d.stream);
            copy(b.h_out, b.out, sizeof(double) * n, cudaMemcpyDeviceToHost);
            scratch_release(d, d.stream);
            if (ce != cudaSuccess && !rc) rc = fail(VELA_ERR_CUDA, "device %d: staging copy failed: %s", d.dev, cudaGetErrorString(ce));
            return rc;
        };
        if (nd == 1) {
            rc = feed(0);
        } else {
            // the handle's per-device feeder threads (alive since create()): staging copies and launches of all devices
            // proceed side by side instead of delaying device i by i copies
            std::vector<int> rcs(nd, 0);
            std::vector<std::string> msgs(nd);
            for (int i = 0; i < nd; ++i)
                h->feeders[i]->post([&, i] { rcs[i] = feed(i); if (rcs[i]) msgs[i] = g_last_error; });   // the message is thread-local
            for (int i = 0; i < nd; ++i) h->feeders[i]->wait();
            for (int i = 0; i < nd && !rc; ++i)
                if (rcs[i]) { rc = rcs[i]; g_last_error = msgs[i]; }
        }
        for (int i = 0; i < nd; ++i) {
            DeviceModel& d = h->devs[i];
            long n = r1[i] - r0[i];
            if (n <= 0) continue;
            cudaSetDevice(d.dev);
            cudaError_t e = cudaStreamSynchronize(d.stream);
            if (e != cudaSuccess && !rc) rc = fail(VELA_ERR_CUDA, "device %d: %s", d.dev, cudaGetErrorString(e));
            if (!rc) memcpy(out + r0[i], d.buf.h_out, sizeof(double) * n);
            if (!rc && h->stage_timing && i == 0) {
                for (int s = 0; s < 4; ++s) cudaEventElapsedTime(&d.stage_ms[s], d.ev[s], d.ev[s + 1]);
                cudaEventElapsedTime(&d.stage_ms[4], d.ev[0], d.ev[4]);
            }
        }
        done += pass;
    }
    cudaSetDevice(prev);
    return rc;
}

VELA_API int vela_b200_lnlike_batch(VelaHandle h, const double* params, int64_t B, int64_t n_free, int64_t row_stride,
                           int64_t col_stride, double* out_lnlike) {
    return run_host_batch(h, params, B, n_free, row_stride, col_stride, nullptr, 0, 0, out_lnlike);
}

VELA_API int vela_b200_lnpost_batch(VelaHandle h, const double* params, int64_t B, int64_t n_free, int64_t row_stride,
                           int64_t col_stride, const double* lnprior, double* out_lnpost) {
    return run_host_batch(h, params, B, n_free, row_stride, col_stride, lnprior, 1, 0, out_lnpost);
}

VELA_API int vela_b200_chi2_batch(VelaHandle h, const double* params, int64_t B, int64_t n_free, int64_t row_stride,
                         int64_t col_stride, double* out_chi2) {
    return run_host_batch(h, params, B, n_free, row_stride, col_stride, nullptr, 0, 1, out_chi2);
}

VELA_API int vela_b200_lnpost_batch_device(VelaHandle h, const double* d_params, int64_t B, const double* d_lnprior,
                                           int with_prior, double* d_out, void* stream) {
    if (!h) return fail(VELA_ERR_INVALID_ARG, "null handle");
    if (B <= 0) return VELA_OK;
    std::lock_guard<std::mutex> lock(h->mtx);
    DeviceModel& d = h->devs[0];
    if (B > h->max_chunk) return fail(VELA_ERR_INVALID_ARG, "device-resident batch of %ld exceeds the per-pass limit %ld", (long)B, h->max_chunk);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(d.dev);
    long cap_before = d.buf.cap_B;
    int rc = ensure_device_buffers(h, d, B);
    cudaStream_t st = stream ? (cudaStream_t)stream : d.stream;
    if (!rc && st != d.stream && d.buf.cap_B != cap_before) cudaStreamSynchronize(d.stream);  // scratch zero-fills ran on the handle's stream
    if (!rc) scratch_acquire(d, st);   // earlier work on another stream may still be reading / writing the scratch
    if (!rc && with_prior && !d_lnprior && h->have_priors) {   // device-side priors
        prior_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(d.priors, h->prog.n_free, d_params, B, 0, d.buf.lnprior);
        h->launches += 1;
        d_lnprior = d.buf.lnprior;
    }
    if (!rc) rc = enqueue_batch(h, d, d_params, B, 0, with_prior, d_lnprior, d_out, st);
    if (!rc) scratch_release(d, st);
    cudaSetDevice(prev);
    return rc;
}

VELA_API int vela_b200_lnlike_batch_device(VelaHandle h, const double* d_params, int64_t B, double* d_out, void* stream) {
    return vela_b200_lnpost_batch_device(h, d_params, B, nullptr, 0, d_out, stream);
}

// Runs prologue + chain for ONE parameter vector on device 0 and copies per-TOA outputs back.
// mode 1: y, w (n_rows each); mode 2: phase hi, lo and spin frequency (n_toas each).
static int run_single(VelaHandle h, const double* params, int64_t n_free, int mode, double* o0, double* o1, double* o2) {
    if (!h) return fail(VELA_ERR_INVALID_ARG, "null handle");
    if (n_free != h->prog.n_free) return fail(VELA_ERR_INVALID_ARG, "expected %d free parameters, got %ld", h->prog.n_free, (long)n_free);
    std::lock_guard<std::mutex> lock(h->mtx);
    DeviceModel& d = h->devs[0];
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(d.dev);
    int rc = ensure_device_buffers(h, d, 1);
    if (rc) { cudaSetDevice(prev); return rc; }
    Buffers& b = d.buf;
    scratch_acquire(d, d.stream);
    double *dY = nullptr, *dW = nullptr, *dF = nullptr;
    cudaMalloc((void**)&dY, sizeof(double) * h->n_rows_pad);
    cudaMalloc((void**)&dW, sizeof(double) * h->n_rows_pad);
    cudaMalloc((void**)&dF, sizeof(double) * h->n_rows_pad);
    cudaMemcpyAsync(b.params, params, sizeof(double) * std::max<long>(n_free, 1), cudaMemcpyHostToDevice, d.stream);
    launch_prologue(h, d, b.params, 1, d.stream);
    ChainArgs ca;
    ca.Pext = b.Pext; ca.tzr_phase = b.tzr; ca.toa = d.toa; ca.toa_stride = d.toa_stride; ca.n_toas = h->n_toas;
    ca.index_masks = d.index_masks; ca.bit_masks = d.bit_masks; ca.B = 1; ca.group = h->chain_group; ca.emit = mode;
    ca.Y = dY; ca.W = dW; ca.F = dF; ca.ld_yw = h->n_rows_pad; ca.partial = b.partial;
    size_t chain_smem = chain_smem_doubles(h->chain_group, h->prog.row, h->chain_threads / 32) * sizeof(double);
    launch_chain(h, dim3((unsigned)h->n_tiles, 1), chain_smem, d.stream, ca);
    h->launches += 2;
    const long n = mode == 1 ? h->n_rows : h->n_toas;
    if (o0) cudaMemcpyAsync(o0, dY, sizeof(double) * n, cudaMemcpyDeviceToHost, d.stream);
    if (o1) cudaMemcpyAsync(o1, dW, sizeof(double) * n, cudaMemcpyDeviceToHost, d.stream);
    if (o2) cudaMemcpyAsync(o2, dF, sizeof(double) * n, cudaMemcpyDeviceToHost, d.stream);
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(d.stream);
    if (e == cudaSuccess) e = e2;
    cudaFree(dY); cudaFree(dW); cudaFree(dF);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(VELA_ERR_CUDA, "single-sample run: %s", cudaGetErrorString(e));
    return VELA_OK;
}

// _calc_resids_and_Ninvdiag for ONE parameter vector
VELA_API int vela_b200_residuals(VelaHandle h, const double* params, int64_t n_free, double* y, double* w) {
    return run_single(h, params, n_free, 1, y, w, nullptr);
}

// phase_residual(toa, ctoa) - tzrphase as a Double64 (hi, lo) and doppler_shifted_spin_frequency per TOA
VELA_API int vela_b200_phases(VelaHandle h, const double* params, int64_t n_free, double* phase_hi, double* phase_lo,
                              double* spin_frequency) {
    return run_single(h, params, n_free, 2, phase_hi, phase_lo, spin_frequency);
}

VELA_API int vela_b200_noise_weights_inv(VelaHandle h, const double* params, int64_t n_free, double* phiinv) {
    if (!h) return fail(VELA_ERR_INVALID_ARG, "null handle");
    if (h->p == 0) return fail(VELA_ERR_UNSUPPORTED, "this kernel has no Phi");
    if (n_free != h->prog.n_free) return fail(VELA_ERR_INVALID_ARG, "expected %d free parameters, got %ld", h->prog.n_free, (long)n_free);
    std::lock_guard<std::mutex> lock(h->mtx);
    DeviceModel& d = h->devs[0];
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(d.dev);
    int rc = ensure_device_buffers(h, d, 1);
    if (rc) { cudaSetDevice(prev); return rc; }
    Buffers& b = d.buf;
    scratch_acquire(d, d.stream);
    cudaMemcpyAsync(b.params, params, sizeof(double) * std::max<long>(n_free, 1), cudaMemcpyHostToDevice, d.stream);
    launch_prologue(h, d, b.params, 1, d.stream);
    CholArgs ch;
    memset(&ch, 0, sizeof ch);
    ch.Pext = b.Pext; ch.row = h->prog.row; ch.p = h->p; ch.pp = h->pp; ch.n_gp = h->n_gp; ch.gp = d.gp; ch.gp_data = d.gp_data;
    double* dphi = nullptr;
    cudaMalloc((void**)&dphi, sizeof(double) * h->pp);
    phiinv_kernel<<<1, 128, 0, d.stream>>>(ch, dphi);
    h->launches += 2;
    cudaMemcpyAsync(phiinv, dphi, sizeof(double) * h->p, cudaMemcpyDeviceToHost, d.stream);
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(d.stream);
    if (e == cudaSuccess) e = e2;
    cudaFree(dphi);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(VELA_ERR_CUDA, "noise_weights_inv: %s", cudaGetErrorString(e));
    return VELA_OK;
}

VELA_API int vela_b200_set_priors(VelaHandle h, const VelaPriorDesc* priors, int32_t n_free) {
    if (!h || !priors) return fail(VELA_ERR_INVALID_ARG, "null handle / priors");
    if (n_free != h->prog.n_free) return fail(VELA_ERR_INVALID_ARG, "expected %d priors, got %d (timing_model.jl:40)", h->prog.n_free, n_free);
    std::vector<PriorDev> pd(std::max(n_free, 1));
    for (int k = 0; k < n_free; ++k) {
        const VelaPriorDesc& s = priors[k];
        PriorDev& p = pd[k];
        p.kind = s.kind; p.a = s.a; p.b = s.b; p.lo = s.lo; p.hi = s.hi; p.norm = 0.0;
        switch (s.kind) {
            case VELA_PRIOR_UNIFORM: if (!(s.b > s.a)) return fail(VELA_ERR_INVALID_ARG, "prior %d: Uniform needs a < b", k); p.norm = -std::log(s.b - s.a); break;
            case VELA_PRIOR_NORMAL: case VELA_PRIOR_LOGNORMAL: if (!(s.b > 0)) return fail(VELA_ERR_INVALID_ARG, "prior %d: sigma must be positive", k); p.norm = std::log(s.b); break;
            case VELA_PRIOR_LOGUNIFORM: if (!(s.a > 0 && s.b > s.a)) return fail(VELA_ERR_INVALID_ARG, "prior %d: LogUniform needs 0 < a < b", k); p.norm = std::log(std::log(s.b / s.a)); break;
            case VELA_PRIOR_PX: if (!(s.a > 0)) return fail(VELA_ERR_INVALID_ARG, "prior %d: PX prior needs Rmax > 0", k); break;
            case VELA_PRIOR_TRUNCATED_NORMAL: {
                if (!(s.b > 0 && s.hi > s.lo)) return fail(VELA_ERR_INVALID_ARG, "prior %d: bad truncated normal", k);
                double c0 = 0.5 * std::erfc(-(s.lo - s.a) / s.b / std::sqrt(2.0)), c1 = 0.5 * std::erfc(-(s.hi - s.a) / s.b / std::sqrt(2.0));
                p.norm = std::log(s.b) + std::log(c1 - c0);
                break;
            }
            case VELA_PRIOR_KIN: case VELA_PRIOR_SINI: case VELA_PRIOR_STIGMA: case VELA_PRIOR_SHAPMAX: case VELA_PRIOR_LATITUDE: break;
            default: return fail(VELA_ERR_UNSUPPORTED, "prior %d: unknown kind %d", k, s.kind);
        }
    }
    std::lock_guard<std::mutex> lock(h->mtx);
    int prev = 0;
    cudaGetDevice(&prev);
    for (auto& d : h->devs) {
        cudaSetDevice(d.dev);
        cudaFree(d.priors);
        if (upload(&d.priors, pd.data(), pd.size())) { cudaSetDevice(prev); return VELA_ERR_CUDA; }
    }
    cudaSetDevice(prev);
    h->have_priors = true;
    return VELA_OK;
}

// mode 0: lnprior per row; mode 1: prior transform per element (device 0)
static int run_prior(VelaHandle h, const double* in, int64_t B, int64_t n_free, int64_t rs, int64_t cs, int mode, double* out) {
    if (!h) return fail(VELA_ERR_INVALID_ARG, "null handle");
    if (!h->have_priors) return fail(VELA_ERR_INVALID_ARG, "no priors set (vela_b200_set_priors)");
    if (n_free != h->prog.n_free) return fail(VELA_ERR_INVALID_ARG, "expected %d free parameters, got %ld", h->prog.n_free, (long)n_free);
    if (B <= 0) return VELA_OK;
    std::lock_guard<std::mutex> lock(h->mtx);
    DeviceModel& d = h->devs[0];
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(d.dev);
    const size_t nin = (size_t)B * n_free, nout = mode == 0 ? (size_t)B : nin;
    std::vector<double> tmp;
    if (!(rs == n_free && cs == 1)) {
        tmp.resize(nin);
        for (int64_t r = 0; r < B; ++r)
            for (int64_t k = 0; k < n_free; ++k) tmp[r * n_free + k] = in[r * rs + k * cs];
        in = tmp.data();
    }
    double *din = nullptr, *dout = nullptr;
    cudaError_t e = cudaMalloc((void**)&din, std::max<size_t>(nin, 1) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dout, nout * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(din, in, nin * sizeof(double), cudaMemcpyHostToDevice, d.stream);
    if (e == cudaSuccess) {
        prior_kernel<<<(unsigned)((B + 127) / 128), 128, 0, d.stream>>>(d.priors, (int)n_free, din, B, mode, dout);
        h->launches += 1;
        e = cudaMemcpyAsync(out, dout, nout * sizeof(double), cudaMemcpyDeviceToHost, d.stream);
    }
    cudaError_t e2 = cudaStreamSynchronize(d.stream);
    if (e == cudaSuccess) e = e2;
    cudaFree(din); cudaFree(dout);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(VELA_ERR_CUDA, "prior evaluation: %s", cudaGetErrorString(e));
    return VELA_OK;
}

VELA_API int vela_b200_lnprior_batch(VelaHandle h, const double* params, int64_t B, int64_t n_free, int64_t row_stride,
                                     int64_t col_stride, double* out_lnprior) {
    return run_prior(h, params, B, n_free, row_stride, col_stride, 0, out_lnprior);
}

VELA_API int vela_b200_prior_transform_batch(VelaHandle h, const double* cube, int64_t B, int64_t n_free, double* out) {
    return run_prior(h, cube, B, n_free, n_free, 1, 1, out);
}

// _calc_Sigmainv__and__MT_Ninv_y (gls_likelihood.jl:101-169) for ONE parameter vector:
// Sigma^-1 = diag(Phi^-1) + M^T N^-1 M as a full symmetric column-major p x p matrix, u = M^T N^-1 y.
VELA_API int vela_b200_sigma_and_u(VelaHandle h, const double* params, int64_t n_free, double* Sigma, double* u) {
    if (!h) return fail(VELA_ERR_INVALID_ARG, "null handle");
    if (h->p == 0) return fail(VELA_ERR_UNSUPPORTED, "this kernel has no Sigma");
    if (n_free != h->prog.n_free) return fail(VELA_ERR_INVALID_ARG, "expected %d free parameters, got %ld", h->prog.n_free, (long)n_free);
    std::vector<double> phi(h->p);
    if (int rc = vela_b200_noise_weights_inv(h, params, n_free, phi.data())) return rc;
    std::lock_guard<std::mutex> lock(h->mtx);
    DeviceModel& d = h->devs[0];
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(d.dev);
    int rc = ensure_device_buffers(h, d, 1);
    if (rc) { cudaSetDevice(prev); return rc; }
    Buffers& b = d.buf;
    scratch_acquire(d, d.stream);
    cudaMemcpyAsync(b.params, params, sizeof(double) * std::max<long>(n_free, 1), cudaMemcpyHostToDevice, d.stream);
    launch_prologue(h, d, b.params, 1, d.stream);
    ChainArgs ca;
    ca.Pext = b.Pext; ca.tzr_phase = b.tzr; ca.toa = d.toa; ca.toa_stride = d.toa_stride; ca.n_toas = h->n_toas;
    const bool wy = h->sp.active;
    ca.index_masks = d.index_masks; ca.bit_masks = d.bit_masks; ca.B = 1; ca.group = h->chain_group; ca.emit = wy ? 3 : 1;
    ca.Y = b.Y; ca.W = b.W; ca.F = nullptr; ca.ld_yw = h->n_rows_pad; ca.partial = b.partial;
    size_t chain_smem = chain_smem_doubles(h->chain_group, h->prog.row, h->chain_threads / 32) * sizeof(double);
    launch_chain(h, dim3((unsigned)h->n_tiles, 1), chain_smem, d.stream, ca);
    std::vector<double> S((size_t)h->ld_sig, 0.0), U(h->pp, 0.0);
    h->launches += 3;
    if (h->sp.active) {
        ColsumArgs cs;
        const int ksplit = launch_colsum(h, d, 1, d.stream, wy, &cs);
        if (h->n_eg > 0) {
            cudaMemsetAsync(b.Sig, 0, (size_t)h->ld_sig * sizeof(double), d.stream);
            launch_ecorr_correction(h, d, 1, d.stream, wy);
            cudaMemcpyAsync(S.data(), b.Sig, sizeof(double) * S.size(), cudaMemcpyDeviceToHost, d.stream);
        }
        std::vector<double> R(h->sp.n_cols, 0.0), tR(h->sp.n_cols);
        for (int z = 0; z < ksplit; ++z) {   // sum the split-K slabs in slice order, as K4 does
            cudaMemcpyAsync(tR.data(), b.R + (size_t)z * cs.split_stride_r, sizeof(double) * tR.size(), cudaMemcpyDeviceToHost, d.stream);
            cudaStreamSynchronize(d.stream);
            for (size_t i = 0; i < R.size(); ++i) R[i] = z == 0 ? tR[i] : R[i] + tR[i];
        }
        for (long t = 0; t < h->n_pairs; ++t) {
            const PairTerm& pt = h->sp.table[t];
            double v = pt.c1 * R[pt.i1];
            if (pt.c2 != 0.0) v += pt.c2 * R[pt.i2];
            S[t] = v + S[t];
        }
        for (int i = 0; i < h->p; ++i) U[i] = R[h->sp.n_r + i];
    } else {
        SigmaArgs sa;
        const int ksplit = launch_sigma(h, d, 1, d.stream, &sa);
        if (h->n_eg > 0) launch_ecorr_correction(h, d, 1, d.stream, false);
        std::vector<double> tS((size_t)h->ld_sig), tU(h->pp);
        for (int z = 0; z < ksplit; ++z) {   // sum the split-K slabs in slice order, as K4 does
            cudaMemcpyAsync(tS.data(), b.Sig + (size_t)z * sa.split_stride_sig, sizeof(double) * tS.size(), cudaMemcpyDeviceToHost, d.stream);
            cudaMemcpyAsync(tU.data(), b.U + (size_t)z * sa.split_stride_u, sizeof(double) * tU.size(), cudaMemcpyDeviceToHost, d.stream);
            cudaStreamSynchronize(d.stream);
            for (size_t i = 0; i < S.size(); ++i) S[i] = z == 0 ? tS[i] : S[i] + tS[i];
            for (size_t i = 0; i < U.size(); ++i) U[i] = z == 0 ? tU[i] : U[i] + tU[i];
        }
    }
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(d.stream);
    if (e == cudaSuccess) e = e2;
    cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(VELA_ERR_CUDA, "sigma_and_u: %s", cudaGetErrorString(e));
    const int p = h->p;
    for (int r = 0; r < p; ++r)
        for (int c = 0; c <= r; ++c) {
            double v = S[(size_t)r * (r + 1) / 2 + c] + (r == c ? phi[r] : 0.0);
            Sigma[(size_t)c * p + r] = v;
            Sigma[(size_t)r * p + c] = v;
        }
    if (u) memcpy(u, U.data(), sizeof(double) * p);
    return VELA_OK;
}

// Conditional Gaussian of the analytically marginalised amplitudes for a batch (device 0):
// ahat = Sigma u and the upper Cholesky factor of Sigma^-1, from the same K3/K4 results as ln L
// (pyvela/pyvela/spnta.py:376-394).
VELA_API int vela_b200_marginalized_mean_batch(VelaHandle h, const double* params, int64_t B, int64_t n_free,
                                               int64_t rs, int64_t cs, double* ahat, double* cf, double* lnlike) {
    if (!h) return fail(VELA_ERR_INVALID_ARG, "null handle");
    if (h->p == 0) return fail(VELA_ERR_UNSUPPORTED, "this kernel has no marginalised amplitudes");
    if (n_free != h->prog.n_free) return fail(VELA_ERR_INVALID_ARG, "expected %d free parameters, got %ld (parameter.jl:162)", h->prog.n_free, (long)n_free);
    if (B < 0 || (B > 0 && (!params || !ahat))) return fail(VELA_ERR_INVALID_ARG, "null params/ahat");
    if (B == 0) return VELA_OK;
    std::lock_guard<std::mutex> lock(h->mtx);
    DeviceModel& d = h->devs[0];
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(d.dev);
    const long p = h->p, nf = h->prog.n_free;
    const long chunk = std::min<long>(h->max_chunk, 1024);
    double *d_ahat = nullptr, *d_cf = nullptr;
    int rc = 0;
    cudaError_t e = cudaMalloc((void**)&d_ahat, sizeof(double) * chunk * p);
    if (e == cudaSuccess && cf) e = cudaMalloc((void**)&d_cf, sizeof(double) * chunk * p * p);
    if (e != cudaSuccess) rc = fail(VELA_ERR_CUDA, "marginalized_mean: %s", cudaGetErrorString(e));
    std::vector<double> stage;
    for (long done = 0; done < B && !rc; done += chunk) {
        const long n = std::min<long>(chunk, B - done);
        if ((rc = ensure_host_buffers(h, d, n))) break;
        if ((rc = ensure_device_buffers(h, d, n))) break;
        Buffers& b = d.buf;
        scratch_acquire(d, d.stream);
        for (long r = 0; r < n; ++r)
            for (long k = 0; k < nf; ++k) b.h_params[r * nf + k] = params[(done + r) * rs + k * cs];
        cudaMemcpyAsync(b.params, b.h_params, sizeof(double) * n * std::max<long>(nf, 1), cudaMemcpyHostToDevice, d.stream);
        rc = enqueue_batch(h, d, b.params, n, 0, 0, nullptr, b.out, d.stream, d_ahat, d_cf);
        cudaMemcpyAsync(ahat + done * p, d_ahat, sizeof(double) * n * p, cudaMemcpyDeviceToHost, d.stream);
        if (cf) cudaMemcpyAsync(cf + done * p * p, d_cf, sizeof(double) * n * p * p, cudaMemcpyDeviceToHost, d.stream);
        if (lnlike) cudaMemcpyAsync(lnlike + done, b.out, sizeof(double) * n, cudaMemcpyDeviceToHost, d.stream);
        e = cudaStreamSynchronize(d.stream);
        if (e != cudaSuccess && !rc) rc = fail(VELA_ERR_CUDA, "marginalized_mean: %s", cudaGetErrorString(e));
    }
    cudaFree(d_ahat); cudaFree(d_cf);
    cudaSetDevice(prev);
    return rc;
}

// 1 if the per-model (NVRTC) chain kernel is in use, 0 if the generic one is; `on` = 0/1 switches, -1 only queries.
VELA_API int vela_b200_chain_specialized(VelaHandle h, int on) {
    if (!h) return VELA_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(h->mtx);
    if (on >= 0) h->use_spec = (on != 0) && h->spec_chain != nullptr;
    return h->use_spec ? 1 : 0;
}

VELA_API int vela_b200_sigma_structured(VelaHandle h) { return h ? (h->sp.active ? 1 : 0) : VELA_ERR_INVALID_ARG; }

// What the Sigma stage of this handle executes for a batch of B walkers (measurement bookkeeping, bench.py):
// info = {structured, two_stage (taken at this batch size), table columns n_cols, padded rows, stage-1 columns
// (32 per block), stage-2 output columns as launched (64 per CTA column), moments per block and walker (k2_pad),
// cells}.  Returns the FP64 operations per walker (2 per multiply-add) of the column-sum launches; 0 on the dense path.
VELA_API double vela_b200_sigma_info(VelaHandle h, int64_t B, int64_t* info) {
    if (!h) return 0.0;
    const StructPlan& sp = h->sp;
    const bool two = sp.active && two_stage_taken(h, B);
    if (info) {
        info[0] = sp.active ? 1 : 0; info[1] = two ? 1 : 0; info[2] = sp.n_cols; info[3] = h->n_rows_pad;
        info[4] = (int64_t)sp.n_blocks * 32; info[5] = (int64_t)h->devs[0].n_units2 * 64; info[6] = sp.k2_pad; info[7] = sp.n_cells;
    }
    if (!sp.active) return 0.0;
    if (!two) return 2.0 * (double)sp.n_cols * (double)h->n_rows_pad;
    return 2.0 * (32.0 * sp.n_blocks * (double)h->n_rows_pad + 64.0 * h->devs[0].n_units2 * (double)sp.k2_pad);
}

// Compiler log / reason the specialised kernel is absent (empty when it compiled cleanly).
VELA_API const char* vela_b200_jit_log(VelaHandle h) { return h ? h->jit_log.c_str() : ""; }

// test / tooling hook (no device needed): run the basis analysis of the structured Sigma path on the host and, when
// `w` is given, evaluate M^T diag(w) M through the column-sum table (packed lower triangle, p(p+1)/2 doubles).
// Returns the number of table columns (0 = the basis stays on the dense path); info = {n_r, n_cols, n_groups}.
VELA_API int64_t vela_b200_debug_structured_sigma(const VelaModelDesc* desc, const double* w, double* sigma_packed, int32_t* info) {
    if (int rc = validate(desc)) return rc;
    if (desc->kernel_kind != VELA_KERNEL_WOODBURY_WHITE && desc->kernel_kind != VELA_KERNEL_WOODBURY_ECORR) return 0;
    const long n_rows = desc->wideband ? 2 * desc->n_toas : desc->n_toas;
    long n_rows_pad = round_up(n_rows, 64);
    const int p = (int)desc->basis_cols;
    StructPlan sp;
    if (!analyse_basis(desc, n_rows, n_rows_pad, p, sp)) return 0;
    if (info) { info[0] = sp.n_r; info[1] = sp.n_cols; info[2] = sp.n_groups; }
    if (w && sigma_packed) {
        std::vector<double> R(sp.n_r, 0.0);
        for (int c = 0; c < sp.n_r; ++c) {
            const double* t = sp.T.data() + (size_t)c * n_rows_pad;
            double acc = 0.0;
            for (long j = 0; j < n_rows; ++j) acc += t[j] * w[j];
            R[c] = acc;
        }
        for (size_t t = 0; t < sp.table.size(); ++t) {
            const PairTerm& pt = sp.table[t];
            sigma_packed[t] = pt.c1 * R[pt.i1] + (pt.c2 != 0.0 ? pt.c2 * R[pt.i2] : 0.0);
        }
    }
    return sp.n_r;
}

// test / tooling hook (no device needed): the column sums R (n_cols values: table columns against w, then the basis
// columns against wy) evaluated on the host twice - directly from the table T, and through the two-stage plan (cell
// moments with the Chebyshev columns A1, then the stage-2 matrices T2) - so the expansion can be checked against the
// direct sums without a GPU.  info = {n_cols, two_stage, cp, n_cells, n_blocks, n_w_blocks, n_jobs, n_rows_pad}.
// Returns n_cols (0 = the basis stays on the dense path).
VELA_API int64_t vela_b200_debug_two_stage(const VelaModelDesc* desc, const double* w, const double* wy, double* R_direct,
                                           double* R_two, int64_t* info) {
    if (int rc = validate(desc)) return rc;
    if (desc->kernel_kind != VELA_KERNEL_WOODBURY_WHITE && desc->kernel_kind != VELA_KERNEL_WOODBURY_ECORR) return 0;
    const long n_rows = desc->wideband ? 2 * desc->n_toas : desc->n_toas;
    long n_rows_pad = round_up(n_rows, 64);
    const int p = (int)desc->basis_cols;
    StructPlan sp;
    if (!analyse_basis(desc, n_rows, n_rows_pad, p, sp)) return 0;
    if (info) {
        info[0] = sp.n_cols; info[1] = sp.two_stage ? 1 : 0; info[2] = sp.cp; info[3] = sp.n_cells; info[4] = sp.n_blocks;
        info[5] = sp.n_w_blocks; info[6] = (int64_t)sp.jobs.size(); info[7] = n_rows_pad;
    }
    if (w && wy && R_direct) {
        for (int c = 0; c < sp.n_cols; ++c) {
            const double* t = sp.T.data() + (size_t)c * n_rows_pad;
            const double* v = c < sp.n_r ? w : wy;
            double acc = 0.0;
            for (long j = 0; j < n_rows; ++j) acc += t[j] * v[j];
            R_direct[c] = acc;
        }
    }
    if (w && wy && R_two && sp.two_stage) {
        for (int c = 0; c < sp.n_cols; ++c) R_two[c] = 0.0;
        // stage 1: mu[block][cell * 32 + q]
        std::vector<double> mu((size_t)sp.n_blocks * sp.k2_pad, 0.0);
        for (int b = 0; b < sp.n_blocks; ++b) {
            const double* v = b < sp.n_w_blocks ? w : wy;
            for (int q = 0; q < 32; ++q) {
                const double* a1 = sp.A1.data() + ((size_t)b * 32 + q) * n_rows_pad;
                for (int c = 0; c < sp.n_cells; ++c) {
                    double acc = 0.0;
                    const long j1 = std::min<long>(n_rows, (long)(c + 1) * sp.cp);
                    for (long j = (long)c * sp.cp; j < j1; ++j) acc += a1[j] * v[j];
                    mu[(size_t)b * sp.k2_pad + (size_t)c * 32 + q] = acc;
                }
            }
        }
        // stage 2
        for (const StructPlan::Stage2Job& job : sp.jobs) {
            const double* m = mu.data() + (size_t)job.block * sp.k2_pad;
            for (int oc = 0; oc < job.n_out; ++oc) {
                const double* t2 = sp.T2.data() + job.t2_off + (size_t)oc * sp.k2_pad;
                double acc = 0.0;
                for (long k = 0; k < sp.k2_pad; ++k) acc += t2[k] * m[k];
                R_two[job.r_off + oc] = acc;
            }
        }
    }
    return sp.n_cols;
}

// test / tooling hook: compile the translation unit below with NVRTC (no device needed); returns the cubin size
VELA_API int64_t vela_b200_debug_spec_compile(const VelaModelDesc* desc) {
    if (int rc = validate(desc)) return rc;
    ProgramDev pg;
    build_program(desc, pg);
    std::vector<char> cubin;
    std::string log;
    int mb = 5;
    if (const char* f = getenv("VELA_B200_JIT_MINBLOCKS")) mb = std::max(1, std::min(16, atoi(f)));
    // tooling: VELA_B200_DEBUG_EMIT=<mode> compiles the build create() makes for a handle of that emit mode
    const int emit = getenv("VELA_B200_DEBUG_EMIT") ? atoi(getenv("VELA_B200_DEBUG_EMIT")) : -1;
    if (jit_compile_cubin(spec_source(pg, mb, 128, emit), cubin, log)) return fail(VELA_ERR_UNSUPPORTED, "%s", log.c_str());
    g_last_error = log;   // ptxas statistics of the successful build, readable through vela_b200_last_error()
    return (int64_t)cubin.size();
}

// test / tooling hook: the NVRTC translation unit create() would compile for this descriptor (no device needed)
VELA_API int64_t vela_b200_debug_spec_source(const VelaModelDesc* desc, char* buf, int64_t cap) {
    if (int rc = validate(desc)) return rc;
    ProgramDev pg;
    build_program(desc, pg);
    const std::string src = spec_source(pg);
    if (buf && cap > 0) {
        const size_t n = std::min<size_t>(src.size(), (size_t)cap - 1);
        memcpy(buf, src.data(), n);
        buf[n] = 0;
    }
    return (int64_t)src.size();
}

VELA_API int vela_b200_set_stage_timing(VelaHandle h, int enabled) {
    if (!h) return VELA_ERR_INVALID_ARG;
    h->stage_timing = enabled != 0;
    return VELA_OK;
}

VELA_API double vela_b200_last_stage_ms(VelaHandle h, int stage) {
    if (!h || !h->stage_timing || stage < 0 || stage > 4) return -1.0;
    DeviceModel& d = h->devs[0];
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(d.dev);
    float ms = -1.0f;
    cudaError_t e = cudaEventSynchronize(d.ev[4]);
    if (e == cudaSuccess) e = stage < 4 ? cudaEventElapsedTime(&ms, d.ev[stage], d.ev[stage + 1]) : cudaEventElapsedTime(&ms, d.ev[0], d.ev[4]);
    if (e != cudaSuccess) { cudaGetLastError(); ms = -1.0f; }
    cudaSetDevice(prev);
    return ms;
}

// test hook: the correctly-rounded sin/cos used for the sky position (n host values in, n out)
VELA_API int vela_b200_debug_sincos(const double* x, int n, double* s, double* c) {
    if (vela_b200_device_count() <= 0) return fail(VELA_ERR_NO_DEVICE, "no CUDA device");
    double *dx = nullptr, *ds = nullptr, *dc = nullptr;
    CU(cudaMalloc((void**)&dx, sizeof(double) * n));
    CU(cudaMalloc((void**)&ds, sizeof(double) * n));
    CU(cudaMalloc((void**)&dc, sizeof(double) * n));
    CU(cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice));
    sincos_cr_kernel<<<(n + 127) / 128, 128>>>(dx, n, ds, dc);
    CU(cudaMemcpy(s, ds, sizeof(double) * n, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(c, dc, sizeof(double) * n, cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(ds); cudaFree(dc);
    return VELA_OK;
}

// test hook: compares the shared-reciprocal quotients of the proper-motion normalisation (div3_by) with the IEEE
// divisions on n pseudo-random operand sets; returns the number of quotients that differ (0 expected), < 0 on error
VELA_API int64_t vela_b200_debug_div3(int64_t n, uint64_t seed) {
    if (vela_b200_device_count() <= 0) return fail(VELA_ERR_NO_DEVICE, "no CUDA device");
    unsigned long long* d = nullptr;
    unsigned long long hval = 0;
    if (cudaMalloc((void**)&d, 8) != cudaSuccess) return VELA_ERR_CUDA;
    cudaMemset(d, 0, 8);
    div3_check_kernel<<<(unsigned)((n + 255) / 256), 256>>>(seed, n, d);
    cudaError_t e = cudaMemcpy(&hval, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(VELA_ERR_CUDA, "div3 check: %s", cudaGetErrorString(e));
    return (int64_t)hval;
}

// FP64 co-issue probe (fp64_coissue_kernel): tflops[0] = DFMA in the even warps alone, [1] = DMMA in the odd warps alone,
// [2] = DFMA part and [3] = DMMA part when both run at once (same launch, so both are rated on the same elapsed time)
VELA_API int vela_b200_debug_fp64_coissue(int device, double* tflops) {
    int prev = 0;
    if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) return fail(VELA_ERR_NO_DEVICE, "no such device");
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 16384;
    double* dout = nullptr;
    cudaMalloc((void**)&dout, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double fl_dfma = (double)blocks * (threads / 2) * iters * 16 * 2;
    const double fl_dmma = (double)blocks * (threads / 64) * iters * 16 * (8.0 * 8 * 4 * 2);
    double best[4] = {0, 0, 0, 0};
    for (int mode = 1; mode <= 3; ++mode)
        for (int rep = 0; rep < 8; ++rep) {
            cudaEventRecord(e0);
            fp64_coissue_kernel<<<blocks, threads>>>(dout, iters, mode);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms <= 0 || rep < 2) continue;
            const double t = ms * 1e-3;
            if (mode == 1) best[0] = std::max(best[0], fl_dfma / t / 1e12);
            else if (mode == 2) best[1] = std::max(best[1], fl_dmma / t / 1e12);
            else if (fl_dfma / t / 1e12 > best[2]) { best[2] = fl_dfma / t / 1e12; best[3] = fl_dmma / t / 1e12; }
        }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(dout);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(VELA_ERR_CUDA, "co-issue probe: %s", cudaGetErrorString(e));
    for (int i = 0; i < 4; ++i) tflops[i] = best[i];
    return VELA_OK;
}

// kind: 0 DFMA, 1 DMMA at full occupancy; 100 + w: DMMA with w warps per SM (one block of 32 w threads per SM)
VELA_API double vela_b200_fp64_peak_tflops(int device, int kind) {
    int prev = 0;
    if (cudaGetDevice(&prev) != cudaSuccess) return -1.0;
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 16384;
    int mix = -1;
    if (kind >= 200) { mix = kind - 200; blocks = prop.multiProcessorCount; threads = 256; kind = 1; iters = 65536; }
    else if (kind >= 100) { blocks = prop.multiProcessorCount; threads = 32 * (kind - 100); kind = 1; iters = 65536; }
    double* dout = nullptr;
    cudaMalloc((void**)&dout, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 12; ++rep) {   // first reps warm the clocks up; best of the rest
        cudaEventRecord(e0);
        if (kind == 0) dfma_peak_kernel<<<blocks, threads>>>(dout, iters);
        else if (mix >= 0) dmma_mix_kernel<<<blocks, threads>>>(dout, iters, mix == 1 ? 1 : (mix == 2 ? 4 : 0), mix == 3 ? 8 : 0);
        else dmma_peak_kernel<<<blocks, threads>>>(dout, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = kind == 0 ? (double)blocks * threads * iters * 16 * 2
                                 : (double)blocks * (threads / 32) * iters * 16 * (8.0 * 8 * 4 * 2);
        if (ms > 0 && rep >= 2) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(dout);
    cudaSetDevice(prev);
    return best;
}

}  // extern "C"
