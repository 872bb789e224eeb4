// libtrixi_b200.so -- C ABI (include/trixi_b200.h): handle life cycle, device-resident state, and the
// launch sequence of one RHS evaluation / 2N Runge-Kutta stage / CFL reduction.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cfloat>
#include <cmath>
#include <new>
#include <string>
#include <vector>
#include <algorithm>
#include <unistd.h>

#include "launch.cuh"

using namespace tb;

namespace {
thread_local std::string g_create_error;

enum { KC_SURFACE = 0, KC_ELEMENT = 1, KC_MAXDT = 2, KC_HALO = 3, KC_HALO_WAIT = 4, KC_COUNT = 5 };

struct ProfEntry {
    cudaEvent_t a, b;
    int cls;
};
}  // namespace

struct trixi_b200_handle {
    int device = 0;
    int ndims = 0, nvars = 0, nnodes = 0, mesh_kind = 0, equation = 0;
    long long nelements = 0, ulen = 0, sfvlen = 0;
    const Launchers *L = nullptr;
    KParams P{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    bool have_elapsed = false;
    std::vector<void *> allocs;
    double *vec[3] = {nullptr, nullptr, nullptr};  // u, du, u_tmp
    double *vec_tmp2 = nullptr;                     // u_tmp2 of the 3S* integrators (allocated at first use)
    unsigned long long *d_cfl = nullptr;
    unsigned long long *h_cfl = nullptr;  // pinned, kCflSlots entries
    double *norm_buf = nullptr;           // calc_error_norms scratch: Vandermonde, weights, sums
    size_t norm_buf_len = 0;
    unsigned long long *norm_linf = nullptr;
    bool opt_fused_cfl = false;           // TRIXI_B200_OPT_FUSED_CFL
    bool opt_single_face_flux = true;     // TRIXI_B200_OPT_SINGLE_FACE_FLUX
    bool opt_fused_stage = true;          // TRIXI_B200_OPT_FUSED_STAGE
    double *integral_buf = nullptr;       // trixi_b200_integrate: [nvars + 1] sums
    bool cfl_valid = false;               // d_cfl holds the maxima of the current u (written by the last RK stage)
    long long launches = 0;
    bool profiling = false;
    std::vector<ProfEntry> prof;
    double prof_ms[KC_COUNT] = {0, 0, 0, 0, 0};
    long long prof_n[KC_COUNT] = {0, 0, 0, 0, 0};
    std::string error;
    int rank = 0, world_size = 1;
    // halo exchange state
    long long nmpi = 0;
    std::vector<long long> mpi_counts;        // [world] faces shared with each rank
    std::vector<int> peers;                   // neighbour ranks in ascending order
    std::vector<long long> peer_offset;       // first face of each peer's segment in my ordering
    std::vector<int> h_peer_slot;             // [nmpi]
    char *comm_base = nullptr;                // one allocation: flags | recv parity 0 | recv parity 1
    size_t flag_bytes = 0, recv_bytes = 0;
    std::vector<void *> ipc_opened;
    std::vector<char *> peer_base;            // [npeers] mapped base of each peer's comm allocation
    std::vector<long long> peer_nmpi;         // [npeers] total faces of that peer (its recv buffer size)
    std::vector<long long> my_offset_in_peer; // [npeers]
    double **d_peer_recv[2] = {nullptr, nullptr};        // device pointer tables per parity
    unsigned long long **d_peer_flag[2] = {nullptr, nullptr};
    int *d_peer_ranks = nullptr;
    long long *d_mpi_remote_index = nullptr;
    bool comm_connected = false;
    unsigned long long comm_seq = 0;
    bool indicator_done = false;  // distributed shock capturing: alpha of this RHS was computed around the halo exchange
    long long *d_peer_nmpi = nullptr;
    // host-buffer calls (rhs_host, step_2n_host) stream u in and the result out in element chunks so the two PCIe
    // directions and the kernels overlap
    int opt_pipeline_chunk = -1;  // TRIXI_B200_OPT_HOST_PIPELINE_CHUNK: -1 auto, 0 off, > 0 elements per chunk
    struct HostPipeline {
        bool built = false, usable = false;
        long long chunk = 0;              // elements per chunk
        int nchunks = 0;
        std::vector<int> order;           // chunk uploaded at step k (sweep along the last coordinate axis)
        std::vector<long long> if_start;  // [nchunks + 1] range of the re-sorted interface list computable after step k
        std::vector<int> ready_start;     // [nchunks + 1] range of ready_chunks whose elements are complete after step k
        std::vector<int> ready_chunks;
        long long *d_if_neighbors = nullptr, *d_if_orient = nullptr;  // interfaces sorted by the step they become ready
        cudaStream_t s_in = nullptr, s_out = nullptr;
        std::vector<cudaEvent_t> ev_in, ev_out;
        cudaEvent_t ev_sync = nullptr;
    } hp;
};

namespace {

int fail(trixi_b200_handle *h, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h)
        h->error = buf;
    else
        g_create_error = buf;
    return code;
}

#define CUDA_TRY(h, expr)                                                                              \
    do {                                                                                               \
        cudaError_t err__ = (expr);                                                                    \
        if (err__ != cudaSuccess)                                                                      \
            return fail(h, err__ == cudaErrorMemoryAllocation ? TRIXI_B200_ENOMEM : TRIXI_B200_ECUDA,  \
                        "%s failed: %s", #expr, cudaGetErrorString(err__));                            \
    } while (0)

template <class T>
int upload_array(trixi_b200_handle *h, const T *host, size_t count, T **dev) {
    *dev = nullptr;
    if (count == 0) return 0;
    if (!host) return fail(h, TRIXI_B200_EINVAL, "descriptor array missing (null pointer, %zu entries expected)", count);
    void *p = nullptr;
    CUDA_TRY(h, cudaMalloc(&p, count * sizeof(T)));
    h->allocs.push_back(p);
    CUDA_TRY(h, cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *dev = static_cast<T *>(p);
    return 0;
}

template <class T>
int alloc_array(trixi_b200_handle *h, size_t count, T **dev) {
    *dev = nullptr;
    if (count == 0) return 0;
    void *p = nullptr;
    CUDA_TRY(h, cudaMalloc(&p, count * sizeof(T)));
    h->allocs.push_back(p);
    *dev = static_cast<T *>(p);
    return 0;
}

struct ProfScope {
    trixi_b200_handle *h;
    int cls;
    ProfEntry e{};
    bool on;
    ProfScope(trixi_b200_handle *h_, int cls_) : h(h_), cls(cls_), on(h_->profiling) {
        if (on) {
            cudaEventCreate(&e.a);
            cudaEventCreate(&e.b);
            e.cls = cls;
            cudaEventRecord(e.a, h->stream);
        }
    }
    ~ProfScope() {
        if (on) {
            cudaEventRecord(e.b, h->stream);
            h->prof.push_back(e);
        }
    }
};

void prof_collect(trixi_b200_handle *h);

void prof_collect(trixi_b200_handle *h) {
    if (h->prof.empty()) return;
    cudaStreamSynchronize(h->stream);
    for (auto &e : h->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.a, e.b);
        h->prof_ms[e.cls] += ms;
        h->prof_n[e.cls] += 1;
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    h->prof.clear();
}

__global__ void k_fp64_peak(double *sink, int iters, double m) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
    const double c = 1e-9;
#pragma unroll 16
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, m, c);
        x1 = fma(x1, m, c);
        x2 = fma(x2, m, c);
        x3 = fma(x3, m, c);
        x4 = fma(x4, m, c);
        x5 = fma(x5, m, c);
        x6 = fma(x6, m, c);
        x7 = fma(x7, m, c);
    }
    const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 123.456) sink[threadIdx.x] = r;
}

__global__ void k_copy(double2 *dst, const double2 *src, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// Stage updates of the other low-storage integrators (pointwise, HBM-bound; not fused into the element kernel).
// 3S* (methods_3Sstar.jl:199-205): u_tmp1 += delta u;  u = gamma1 u + gamma2 u_tmp1 + gamma3 u_tmp2 + beta dt du
__global__ void k_stage_3sstar(double *__restrict__ u, double *__restrict__ u1, const double *__restrict__ u2,
                               const double *__restrict__ du, size_t n, double delta, double g1, double g2, double g3,
                               double bdt) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ui = u[i];
    const double t1 = fma(delta, ui, u1[i]);
    u1[i] = t1;
    u[i] = fma(bdt, du[i], fma(g3, u2[i], fma(g2, t1, g1 * ui)));
}
// SimpleSSPRK33 (methods_SSP.jl:192-201): forward Euler step, then the convex combination with numerator and
// denominator kept apart (a true division, like the reference, so that conservation is not eroded by rounding)
__global__ void k_stage_ssp(double *__restrict__ u, const double *__restrict__ u_tmp, const double *__restrict__ du,
                            size_t n, double dt, double na, double nb, double den) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ue = fma(dt, du[i], u[i]);
    u[i] = fma(na, u_tmp[i], nb * ue) / den;
}

// start_mpi_send! completion: after the pack kernel (stream order) every peer's flag for me is raised to
// the sequence number of this RHS evaluation
__global__ void k_mpi_signal(unsigned long long *const *peer_flag, int npeers, unsigned long long seq) {
    for (int p = threadIdx.x; p < npeers; p += blockDim.x) {  // (world_size <= 64: up to 63 peers)
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(peer_flag[p]) = seq;
        __threadfence_system();
    }
}
// finish_mpi_receive! (dg_parallel.jl:134-182): wait until every neighbour rank has delivered its faces
__global__ void k_mpi_wait(const unsigned long long *flags, const int *peer_ranks, int npeers,
                           unsigned long long seq) {
    for (int p = threadIdx.x; p < npeers; p += blockDim.x) {
        const volatile unsigned long long *f = flags + peer_ranks[p];
        while (*f < seq) {
        }
        __threadfence_system();
    }
}

int check_launch(trixi_b200_handle *h, const char *what) {
    if (h->prof.size() > 2048) prof_collect(h);  // bound the number of live events
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(h, TRIXI_B200_ECUDA, "%s launch failed: %s", what, cudaGetErrorString(err));
    return 0;
}

// one copy of every conforming interface flux where the element kernel of this configuration can fetch it from
// the left neighbour (decided per evaluation: the kernel selection can change through set_option)
void choose_face_flux_layout(trixi_b200_handle *h) {
    h->P.sfv_single = h->opt_single_face_flux && h->P.minus_nb && h->L->single_face_flux(h->P) ? 1 : 0;
}

// surface fluxes of all faces: interfaces (+ boundaries)
int run_surface_fluxes(trixi_b200_handle *h, double t) {
    h->P.t = t;
    {
        ProfScope ps(h, KC_SURFACE);
        h->L->interface_flux(h->P, h->stream);
        if (h->P.ninterfaces) h->launches++;
        if (h->P.nboundaries) {
            h->L->boundary_flux(h->P, h->stream);
            h->launches++;
        }
        if (h->P.nmortars) {
            h->L->mortar_flux(h->P, h->stream);
            h->launches++;
        }
    }
    return check_launch(h, "surface flux kernel");
}

// stage 1: per-element blending factors of the current u; stage 2: smoothing (needs the neighbours' stage 1,
// across ranks: after the halo exchange that carries them)
int run_indicator(trixi_b200_handle *h, int stage) {
    h->L->indicator(h->P, stage, h->stream);
    if (stage == 1 || (h->P.ind_smooth && h->P.ninterfaces + h->P.nmortars + h->P.nmpi > 0)) h->launches++;
    return check_launch(h, "indicator kernel");
}

int run_element(trixi_b200_handle *h, bool with_surface) {
    if (h->P.volume_integral == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG && !h->indicator_done) {
        // blending factors of the current u, before the element kernel may update u in place
        h->P.recv_alpha = nullptr;  // no halo exchange around this call: local smoothing only
        int rc = run_indicator(h, 1);
        if (rc == 0) rc = run_indicator(h, 2);
        if (rc) return rc;
    }
    h->indicator_done = false;
    {
        ProfScope ps(h, KC_ELEMENT);
        cudaError_t err = h->L->element(h->P, with_surface, h->stream);
        if (err != cudaSuccess) return fail(h, TRIXI_B200_ECUDA, "element kernel setup failed: %s", cudaGetErrorString(err));
        h->launches++;
    }
    return check_launch(h, "element kernel");
}

// One RHS evaluation's surface part.  Distributed order (dg_3d_parallel.jl:8-117): pack+send, local
// interfaces and boundaries while the faces travel, wait, MPI interface fluxes.
int run_all_surface_fluxes(trixi_b200_handle *h, double t) {
    choose_face_flux_layout(h);
    const bool dist = h->world_size > 1 && h->nmpi > 0;
    if (h->world_size > 1 && !h->comm_connected)
        return fail(h, TRIXI_B200_ECOMM, "world_size > 1 but the halo exchange is not connected (trixi_b200_comm_connect)");
    int parity = 0;
    const int npeers = (int)h->peers.size();
    const bool sc = h->P.volume_integral == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG;
    if (dist && sc) {
        // the unsmoothed blending factors travel with the face states
        int rc = run_indicator(h, 1);
        if (rc) return rc;
    }
    if (dist) {
        const unsigned long long seq = ++h->comm_seq;
        parity = (int)(seq & 1);
        h->P.peer_recv = h->d_peer_recv[parity];
        h->P.recv = reinterpret_cast<const double *>(h->comm_base + h->flag_bytes + parity * h->recv_bytes);
        h->P.recv_alpha = h->P.recv + h->nmpi * (long long)h->nvars * ipow(h->nnodes, h->ndims - 1);
        ProfScope ps(h, KC_HALO);
        h->L->mpi_pack(h->P, h->stream);
        k_mpi_signal<<<1, 32, 0, h->stream>>>(h->d_peer_flag[parity], npeers, seq);
        h->launches += 2;
    }
    int rc = check_launch(h, "halo pack");
    if (rc) return rc;
    rc = run_surface_fluxes(h, t);
    if (rc) return rc;
    if (dist) {
        {
            ProfScope ps(h, KC_HALO_WAIT);
            const unsigned long long *flags =
                reinterpret_cast<const unsigned long long *>(h->comm_base) + (size_t)parity * h->world_size;
            k_mpi_wait<<<1, 32, 0, h->stream>>>(flags, h->d_peer_ranks, npeers, h->comm_seq);
        }
        ProfScope ps(h, KC_HALO);
        h->L->mpi_interface_flux(h->P, h->stream);
        h->launches += 2;
    }
    rc = check_launch(h, "mpi interface flux");
    if (rc == 0 && dist && sc) {
        rc = run_indicator(h, 2);
        h->indicator_done = true;
    }
    return rc;
}

int run_rhs(trixi_b200_handle *h, double t) {
    int rc = run_all_surface_fluxes(h, t);
    if (rc) return rc;
    h->P.mode = 0;
    return run_element(h, true);
}


void free_host_pipeline(trixi_b200_handle *h) {
    auto &hp = h->hp;
    for (cudaEvent_t e : hp.ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : hp.ev_out) cudaEventDestroy(e);
    hp.ev_in.clear();
    hp.ev_out.clear();
    if (hp.ev_sync) cudaEventDestroy(hp.ev_sync);
    if (hp.s_in) cudaStreamDestroy(hp.s_in);
    if (hp.s_out) cudaStreamDestroy(hp.s_out);
    if (hp.d_if_neighbors) cudaFree(hp.d_if_neighbors);
    if (hp.d_if_orient) cudaFree(hp.d_if_orient);
    hp = trixi_b200_handle::HostPipeline{};
}

// Plan of the chunked host-buffer path.  Chunks are contiguous element ranges, uploaded in the order of a sweep
// along the last coordinate axis; an interface can be computed once both neighbours are resident, an element
// chunk once all its interfaces are.  Everything is derived from the device-resident connectivity, so the plan
// can be rebuilt when the chunk size option changes.
int build_host_pipeline(trixi_b200_handle *h) {
    free_host_pipeline(h);
    auto &hp = h->hp;
    hp.built = true;
    const KParams &P = h->P;
    if (h->opt_pipeline_chunk == 0 || P.curved || P.nmortars || P.nboundaries || h->world_size > 1 ||
        P.ninterfaces == 0 || h->nelements == 0 || P.volume_integral == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG)
        return 0;  // (the indicator's smoothing needs every neighbour's alpha: no chunk-wise readiness)
    const int nd = h->ndims;
    const long long esz = h->ulen / h->nelements;  // doubles per element
    long long chunk = h->opt_pipeline_chunk;
    if (chunk < 0) {
        // 16-32 MiB per copy, a power of two of elements (whole boxes of the Morton order).  Measured at 134 M DOF
        // (tools/pcie_probe.py): 2.6 MB copies 172 ms per rhs_host, 10 MB 143 ms, 21 MB 138 ms, 84 MB 154 ms; the
        // simultaneous two-way PCIe ceiling is 120 ms, one copy each way 208 ms
        chunk = 1;
        while (2 * chunk * esz * (long long)sizeof(double) <= (32ll << 20)) chunk *= 2;
        if (h->nelements < 32 * chunk) return 0;  // small problems: one copy each way is as fast
    }
    const long long nel = h->nelements;
    const long long nch = (nel + chunk - 1) / chunk;
    if (nch < 2 || nch > (1 << 20)) return 0;
    hp.chunk = chunk;
    hp.nchunks = (int)nch;

    const long long nif = P.ninterfaces;
    std::vector<long long> nb((size_t)(2 * nif)), ori((size_t)nif);
    std::vector<double> key((size_t)nch);
    CUDA_TRY(h, cudaMemcpy(nb.data(), P.if_neighbors, sizeof(long long) * 2 * nif, cudaMemcpyDeviceToHost));
    CUDA_TRY(h, cudaMemcpy(ori.data(), P.if_orient, sizeof(long long) * nif, cudaMemcpyDeviceToHost));
    const long long nn = esz / h->nvars;
    CUDA_TRY(h, cudaMemcpy2D(key.data(), sizeof(double), P.node_coordinates + (nd - 1),
                             sizeof(double) * nd * nn * chunk, sizeof(double), (size_t)nch, cudaMemcpyDeviceToHost));
    hp.order.resize((size_t)nch);
    for (int c = 0; c < (int)nch; ++c) hp.order[(size_t)c] = c;
    std::stable_sort(hp.order.begin(), hp.order.end(), [&](int x, int y) { return key[(size_t)x] < key[(size_t)y]; });
    std::vector<int> pos((size_t)nch);
    for (int k = 0; k < (int)nch; ++k) pos[(size_t)hp.order[(size_t)k]] = k;

    // step at which every interface / element chunk becomes computable
    std::vector<int> if_ready((size_t)nif), el_ready(pos);
    hp.if_start.assign((size_t)nch + 1, 0);
    for (long long i = 0; i < nif; ++i) {
        const long long ca = (nb[(size_t)(2 * i)] - 1) / chunk, cb = (nb[(size_t)(2 * i + 1)] - 1) / chunk;
        const int r = std::max(pos[(size_t)ca], pos[(size_t)cb]);
        if_ready[(size_t)i] = r;
        el_ready[(size_t)ca] = std::max(el_ready[(size_t)ca], r);
        el_ready[(size_t)cb] = std::max(el_ready[(size_t)cb], r);
        hp.if_start[(size_t)r + 1]++;
    }
    for (long long k = 0; k < nch; ++k) hp.if_start[(size_t)k + 1] += hp.if_start[(size_t)k];
    {
        std::vector<long long> cursor(hp.if_start.begin(), hp.if_start.end() - 1);
        std::vector<long long> snb((size_t)(2 * nif)), sori((size_t)nif);
        for (long long i = 0; i < nif; ++i) {  // counting sort, stable: the reference's interface order within a step
            const long long j = cursor[(size_t)if_ready[(size_t)i]]++;
            snb[(size_t)(2 * j)] = nb[(size_t)(2 * i)];
            snb[(size_t)(2 * j + 1)] = nb[(size_t)(2 * i + 1)];
            sori[(size_t)j] = ori[(size_t)i];
        }
        CUDA_TRY(h, cudaMalloc((void **)&hp.d_if_neighbors, sizeof(long long) * 2 * nif));
        CUDA_TRY(h, cudaMalloc((void **)&hp.d_if_orient, sizeof(long long) * nif));
        CUDA_TRY(h, cudaMemcpy(hp.d_if_neighbors, snb.data(), sizeof(long long) * 2 * nif, cudaMemcpyHostToDevice));
        CUDA_TRY(h, cudaMemcpy(hp.d_if_orient, sori.data(), sizeof(long long) * nif, cudaMemcpyHostToDevice));
    }
    hp.ready_start.assign((size_t)nch + 1, 0);
    for (long long c = 0; c < nch; ++c) hp.ready_start[(size_t)el_ready[(size_t)c] + 1]++;
    for (long long k = 0; k < nch; ++k) hp.ready_start[(size_t)k + 1] += hp.ready_start[(size_t)k];
    {
        std::vector<int> cursor(hp.ready_start.begin(), hp.ready_start.end() - 1);
        hp.ready_chunks.resize((size_t)nch);
        for (int c = 0; c < (int)nch; ++c) hp.ready_chunks[(size_t)cursor[(size_t)el_ready[(size_t)c]]++] = c;  // ascending
    }
    CUDA_TRY(h, cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaEventCreateWithFlags(&hp.ev_sync, cudaEventDisableTiming));
    hp.ev_in.resize((size_t)nch);
    hp.ev_out.resize((size_t)nch);
    for (auto &e : hp.ev_in) CUDA_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : hp.ev_out) CUDA_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    hp.usable = true;
    return 0;
}

bool host_pipeline_usable(trixi_b200_handle *h, int *rc) {
    *rc = 0;
    if (!h->hp.built) *rc = build_host_pipeline(h);
    return *rc == 0 && h->hp.usable;
}

// Surface fluxes + element kernel (P.mode / rk_* set by the caller) chunk by chunk.  in_host: u arrives from the
// host chunk-wise and every kernel starts as soon as its inputs are resident; out_host: the finished chunks of
// vec[out_which] leave for the host while later chunks are still computed.  In-place RK stages stay safe: a chunk
// is only updated after all interfaces touching it have been evaluated.
int run_pipelined(trixi_b200_handle *h, double t, const double *in_host, double *out_host, int out_which) {
    auto &hp = h->hp;
    const long long nel = h->nelements, esz = h->ulen / nel, chunk = hp.chunk;
    const size_t dbl = sizeof(double);
    int n_out = 0;
    auto element_range = [&](long long c0, long long c1) -> int {  // chunks [c0, c1)
        h->P.elem_begin = c0 * chunk;
        h->P.elem_end = std::min(nel, c1 * chunk);
        int rc = run_element(h, true);
        if (rc == 0 && out_host) {
            cudaEvent_t ev = hp.ev_out[(size_t)n_out++];
            const long long off = h->P.elem_begin * esz, len = (h->P.elem_end - h->P.elem_begin) * esz;
            cudaError_t err = cudaEventRecord(ev, h->stream);
            if (err == cudaSuccess) err = cudaStreamWaitEvent(hp.s_out, ev, 0);
            if (err == cudaSuccess)
                err = cudaMemcpyAsync(out_host + off, h->vec[out_which] + off, len * dbl, cudaMemcpyDeviceToHost, hp.s_out);
            if (err != cudaSuccess) rc = fail(h, TRIXI_B200_ECUDA, "chunked download failed: %s", cudaGetErrorString(err));
        }
        h->P.elem_begin = 0;
        h->P.elem_end = nel;
        return rc;
    };
    if (in_host) {
        h->P.t = t;
        choose_face_flux_layout(h);
        // u is overwritten: everything queued earlier must have finished reading it
        CUDA_TRY(h, cudaEventRecord(hp.ev_sync, h->stream));
        CUDA_TRY(h, cudaStreamWaitEvent(hp.s_in, hp.ev_sync, 0));
        const long long *nb_all = h->P.if_neighbors, *ori_all = h->P.if_orient;
        const long long nif_all = h->P.ninterfaces;
        int rc = 0;
        for (int k = 0; k < hp.nchunks && rc == 0; ++k) {
            const long long c = hp.order[(size_t)k];
            const long long off = c * chunk * esz, len = (std::min(nel, (c + 1) * chunk) - c * chunk) * esz;
            CUDA_TRY(h, cudaMemcpyAsync(h->vec[0] + off, in_host + off, len * dbl, cudaMemcpyHostToDevice, hp.s_in));
            CUDA_TRY(h, cudaEventRecord(hp.ev_in[(size_t)k], hp.s_in));
            CUDA_TRY(h, cudaStreamWaitEvent(h->stream, hp.ev_in[(size_t)k], 0));
            const long long i0 = hp.if_start[(size_t)k], i1 = hp.if_start[(size_t)k + 1];
            if (i1 > i0) {
                h->P.if_neighbors = hp.d_if_neighbors + 2 * i0;
                h->P.if_orient = hp.d_if_orient + i0;
                h->P.ninterfaces = i1 - i0;
                {
                    ProfScope ps(h, KC_SURFACE);
                    h->L->interface_flux(h->P, h->stream);
                    h->launches++;
                }
                h->P.if_neighbors = nb_all;
                h->P.if_orient = ori_all;
                h->P.ninterfaces = nif_all;
                rc = check_launch(h, "surface flux kernel");
            }
            for (int q = hp.ready_start[(size_t)k]; q < hp.ready_start[(size_t)k + 1] && rc == 0;) {
                int q1 = q + 1;  // merge runs of consecutive chunks into one launch and one copy
                while (q1 < hp.ready_start[(size_t)k + 1] && hp.ready_chunks[(size_t)q1] == hp.ready_chunks[(size_t)q1 - 1] + 1) ++q1;
                rc = element_range(hp.ready_chunks[(size_t)q], hp.ready_chunks[(size_t)q1 - 1] + 1);
                q = q1;
            }
        }
        if (rc) return rc;
    } else {
        int rc = run_all_surface_fluxes(h, t);
        for (long long c = 0; c < hp.nchunks && rc == 0; ++c) rc = element_range(c, c + 1);
        if (rc) return rc;
    }
    if (out_host) CUDA_TRY(h, cudaStreamSynchronize(hp.s_out));
    return 0;
}

// the stage loop of step!(integrator::SimpleIntegrator2N) (methods_2N.jl:144-159); with host pointers the first
// stage consumes u as it arrives and the last stage returns it as it is finished
int run_step_2n(trixi_b200_handle *h, double t, double dt, const double *a, const double *b, const double *c, int nstages,
                const double *u_in_host, double *u_out_host) {
    h->cfl_valid = false;
    const bool fuse_cfl = h->opt_fused_cfl && h->L->fuses_cfl(h->P);
    for (int s = 0; s < nstages; ++s) {
        const double t_stage = t + dt * c[s];
        const double *in = s == 0 ? u_in_host : nullptr;
        double *out = s == nstages - 1 ? u_out_host : nullptr;
        h->P.mode = 1;
        h->P.rk_a = a[s];
        h->P.rk_b_dt = b[s] * dt;
        h->P.rk_read_tmp = a[s] != 0.0;  // a = 0: du - 0 * u_tmp = du, u_tmp is not read (first stage: no zero fill)
        h->P.rk_write_tmp = 1;
        if (fuse_cfl && s == nstages - 1) {
            // the last stage also reduces the CFL wave speeds of the state it writes (max_dt without a pass over u)
            CUDA_TRY(h, cudaMemsetAsync(h->d_cfl, 0, kCflSlots * sizeof(unsigned long long), h->stream));
            h->P.want_cfl = 1;
        }
        int rc;
        if (in || out)
            rc = run_pipelined(h, t_stage, in, out, 0);
        else {
            rc = run_all_surface_fluxes(h, t_stage);
            if (rc == 0) rc = run_element(h, true);
        }
        h->P.mode = 0;
        h->P.want_cfl = 0;
        if (rc) return rc;
    }
    h->cfl_valid = fuse_cfl;
    return 0;
}

// One stage of a 3S* or SSP scheme with the update fused into the element kernel (KParams::mode 2 / 3, set up by the
// caller): surface fluxes of u, then the element kernel writes u (and u_tmp) instead of du.  The last stage also
// reduces the CFL wave speeds where the kernel can.
int run_fused_stage(trixi_b200_handle *h, double t_stage, bool cfl) {
    if (cfl) {
        CUDA_TRY(h, cudaMemsetAsync(h->d_cfl, 0, kCflSlots * sizeof(unsigned long long), h->stream));
        h->P.want_cfl = 1;
    }
    int rc = run_all_surface_fluxes(h, t_stage);
    if (rc == 0) rc = run_element(h, true);
    h->P.mode = 0;
    h->P.want_cfl = 0;
    return rc;
}

// every element kernel but the previous-generation headline kernel (TRIXI_B200_OPT_KERNEL_PATH 2) knows modes 2 and 3
bool stage_fusable(const trixi_b200_handle *h) { return h->opt_fused_stage && h->P.kernel_path != 2; }

}  // namespace

extern "C" {

TRIXI_B200_API int trixi_b200_abi_version(void) { return TRIXI_B200_ABI_VERSION; }

TRIXI_B200_API const char *trixi_b200_last_error(const trixi_b200_handle *h) { return h ? h->error.c_str() : g_create_error.c_str(); }

TRIXI_B200_API void trixi_b200_destroy(trixi_b200_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (auto &e : h->prof) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    free_host_pipeline(h);
    for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    for (void *p : h->allocs) cudaFree(p);
    if (h->h_cfl) cudaFreeHost(h->h_cfl);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    if (h->ev_t1) cudaEventDestroy(h->ev_t1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

TRIXI_B200_API int trixi_b200_create(const trixi_b200_desc *d, trixi_b200_handle **out) {
    if (!out) return fail(nullptr, TRIXI_B200_EINVAL, "out pointer is null");
    *out = nullptr;
    if (!d) return fail(nullptr, TRIXI_B200_EINVAL, "descriptor is null");
    if (d->abi_version != TRIXI_B200_ABI_VERSION)
        return fail(nullptr, TRIXI_B200_EINVAL, "descriptor ABI version %d != library %d", d->abi_version,
                    TRIXI_B200_ABI_VERSION);
    if (d->mesh_kind != TRIXI_B200_MESH_TREE && d->mesh_kind != TRIXI_B200_MESH_STRUCTURED &&
        d->mesh_kind != TRIXI_B200_MESH_P4EST)
        return fail(nullptr, TRIXI_B200_EINVAL, "mesh kind %d not supported by this build", d->mesh_kind);
    const bool structured = d->mesh_kind == TRIXI_B200_MESH_STRUCTURED;
    const bool p4est = d->mesh_kind == TRIXI_B200_MESH_P4EST;
    if (p4est && (!d->contravariant_vectors || (d->ninterfaces > 0 && !d->interface_node_indices) ||
                  (d->nboundaries > 0 && !d->boundary_node_indices)))
        return fail(nullptr, TRIXI_B200_EINVAL, "P4estMesh needs contravariant_vectors and node_indices");
    if (p4est && d->nmpiinterfaces > 0 && !d->mpi_node_indices)
        return fail(nullptr, TRIXI_B200_EINVAL, "P4estMesh MPI interfaces need mpi_node_indices");
    if (structured && (!d->contravariant_vectors || !d->left_neighbors))
        return fail(nullptr, TRIXI_B200_EINVAL, "StructuredMesh needs contravariant_vectors and left_neighbors");
    if (structured && d->world_size > 1)
        return fail(nullptr, TRIXI_B200_EINVAL, "StructuredMesh is single-rank (as in the reference)");
    if (d->nmortars < 0) return fail(nullptr, TRIXI_B200_EINVAL, "negative container size");
    if (d->nmortars > 0) {
        if (structured) return fail(nullptr, TRIXI_B200_EINVAL, "a StructuredMesh has no mortars");
        if (!d->mortar_neighbor_ids || !d->mortar_forward_upper || !d->mortar_forward_lower || !d->mortar_reverse_upper ||
            !d->mortar_reverse_lower || (p4est ? !d->mortar_node_indices : !d->mortar_large_sides || !d->mortar_orientations))
            return fail(nullptr, TRIXI_B200_EINVAL, "mortar arrays missing");
    }
    if (d->nmpimortars < 0) return fail(nullptr, TRIXI_B200_EINVAL, "negative container size");
    if (d->nmpimortars > 0) {
        if (structured || d->world_size < 2 || d->nmpiinterfaces <= 0 || !d->mpi_is_mortar_piece)
            return fail(nullptr, TRIXI_B200_EINVAL, "MPI mortars need world_size > 1 and their exchange entries in the MPI interface list");
        if (!d->mpi_mortar_neighbor_ids || !d->mortar_forward_upper || !d->mortar_forward_lower || !d->mortar_reverse_upper ||
            !d->mortar_reverse_lower ||
            (p4est ? !d->mpi_mortar_node_indices || !d->mpi_mortar_normal_directions
                   : !d->mpi_mortar_large_sides || !d->mpi_mortar_orientations))
            return fail(nullptr, TRIXI_B200_EINVAL, "MPI mortar arrays missing");
        const int64_t np1 = ((int64_t)1 << (d->ndims - 1)) + 1;
        for (int64_t q = 0; q < np1 * d->nmpimortars; ++q) {
            const int64_t id = d->mpi_mortar_neighbor_ids[q];
            if (id > d->nelements || -id > d->nmpiinterfaces || (id < 0 && !d->mpi_is_mortar_piece[-id - 1]))
                return fail(nullptr, TRIXI_B200_EINVAL, "mpi_mortar_neighbor_ids[%lld] = %lld is out of range", (long long)q,
                            (long long)id);
        }
    }
    if (d->volume_integral != TRIXI_B200_VOLINT_WEAK_FORM && d->volume_integral != TRIXI_B200_VOLINT_FLUX_DIFFERENCING &&
        d->volume_integral != TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG && d->volume_integral != TRIXI_B200_VOLINT_PURE_LGL_FV)
        return fail(nullptr, TRIXI_B200_EINVAL, "unsupported volume integral type %d", d->volume_integral);
    if (d->volume_integral == TRIXI_B200_VOLINT_PURE_LGL_FV && d->mesh_kind != TRIXI_B200_MESH_TREE)
        for (int a = 0; a < d->ndims; ++a)
            if (!d->subcell_normal_vectors[a])
                return fail(nullptr, TRIXI_B200_EINVAL,
                            "VolumeIntegralPureLGLFiniteVolume on a curved mesh needs subcell_normal_vectors (NormalVectorContainer)");
    if (d->volume_integral == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG) {
        if (d->mesh_kind != TRIXI_B200_MESH_TREE) {
            for (int a = 0; a < d->ndims; ++a)
                if (!d->subcell_normal_vectors[a])
                    return fail(nullptr, TRIXI_B200_EINVAL,
                                "VolumeIntegralShockCapturingHG on a curved mesh needs subcell_normal_vectors (NormalVectorContainer)");
        }
        if (d->equation != TRIXI_B200_EQ_EULER_2D && d->equation != TRIXI_B200_EQ_EULER_3D &&
            d->equation != TRIXI_B200_EQ_MHD_3D)
            return fail(nullptr, TRIXI_B200_EINVAL,
                        "VolumeIntegralShockCapturingHG needs the compressible Euler or the ideal GLM-MHD equations");
        if (!d->inverse_vandermonde_legendre)
            return fail(nullptr, TRIXI_B200_EINVAL, "inverse_vandermonde_legendre missing");
        if (d->indicator_variable < TRIXI_B200_INDVAR_DENSITY_PRESSURE || d->indicator_variable > TRIXI_B200_INDVAR_PRESSURE)
            return fail(nullptr, TRIXI_B200_EINVAL, "unknown indicator variable %d", d->indicator_variable);
        if (d->nnodes < 3) return fail(nullptr, TRIXI_B200_EINVAL, "IndicatorHennemannGassner needs nnodes >= 3");
    }
    if (d->nelements < 0 || d->ninterfaces < 0 || d->nboundaries < 0)
        return fail(nullptr, TRIXI_B200_EINVAL, "negative container size");

    // registry entries the device implements (physics.cuh): anything else would run with zero sources / NaN
    // boundary states without an error, so it is refused here
    {
        const bool euler = d->equation == TRIXI_B200_EQ_EULER_2D || d->equation == TRIXI_B200_EQ_EULER_3D;
        const bool advection = d->equation == TRIXI_B200_EQ_ADVECTION_2D || d->equation == TRIXI_B200_EQ_ADVECTION_3D;
        // flux_hlle without the Powell term is the compressible Euler one (min_max_speed_einfeldt of those equations)
        if (!euler && (d->surface_flux == TRIXI_B200_FLUX_HLLE || d->volume_flux == TRIXI_B200_FLUX_HLLE ||
                       d->volume_flux_fv == TRIXI_B200_FLUX_HLLE))
            return fail(nullptr, TRIXI_B200_EINVAL, "flux_hlle (FluxHLL(min_max_speed_einfeldt)) is implemented for the "
                                                    "compressible Euler equations and, with the Powell term, for GLM-MHD");
        if (!euler && (d->surface_flux == TRIXI_B200_FLUX_HLLC || d->volume_flux == TRIXI_B200_FLUX_HLLC ||
                       d->volume_flux_fv == TRIXI_B200_FLUX_HLLC))
            return fail(nullptr, TRIXI_B200_EINVAL, "flux_hllc is implemented for the compressible Euler equations");
        const int src = d->source_terms;
        const bool src_ok = src == TRIXI_B200_SRC_NONE ||
                            (euler && (src == TRIXI_B200_SRC_CONVERGENCE_TEST || src == TRIXI_B200_SRC_EOC_TEST_EULER ||
                                       src == TRIXI_B200_SRC_EOC_TEST_COUPLED_EULER_GRAVITY));
        if (!src_ok)
            return fail(nullptr, TRIXI_B200_EINVAL, "source terms %d are not implemented on the device for equation %d", src,
                        d->equation);
        for (int k = 0; k < 2 * d->ndims && k < 6; ++k) {
            const int bc = d->boundary_conditions[k], ic = d->boundary_ic[k];
            if (bc != TRIXI_B200_BC_PERIODIC && bc != TRIXI_B200_BC_DIRICHLET && bc != TRIXI_B200_BC_SLIP_WALL)
                return fail(nullptr, TRIXI_B200_EINVAL, "unknown boundary condition %d in direction %d", bc, k + 1);
            if (bc == TRIXI_B200_BC_SLIP_WALL && !euler)
                return fail(nullptr, TRIXI_B200_EINVAL, "boundary_condition_slip_wall needs the compressible Euler equations");
            if (bc != TRIXI_B200_BC_DIRICHLET) continue;
            const bool ic_ok = ic == TRIXI_B200_IC_CONSTANT || ic == TRIXI_B200_IC_CONVERGENCE_TEST ||
                               (euler && (ic == TRIXI_B200_IC_WEAK_BLAST_WAVE ||
                                          ic == TRIXI_B200_IC_EOC_TEST_COUPLED_EULER_GRAVITY));
            if (!ic_ok)
                return fail(nullptr, TRIXI_B200_EINVAL,
                            "Dirichlet boundary state %d (direction %d) is not implemented on the device for equation %d", ic,
                            k + 1, d->equation);
        }
    }

    const Launchers *L = nullptr;
    switch (d->equation) {
    case TRIXI_B200_EQ_ADVECTION_2D: L = get_launchers_advection2d(d->nnodes); break;
    case TRIXI_B200_EQ_ADVECTION_3D: L = get_launchers_advection3d(d->nnodes); break;
    case TRIXI_B200_EQ_EULER_2D:
    case TRIXI_B200_EQ_EULER_3D: {
        // flux_hllc, flux_hlle and (along normals) flux_chandrashekar live in the run-time switches of a second launcher
        // table (EulerAllFluxes, physics.cuh): the default table's kernels are not sized for bodies they never run
        const bool curved_mesh = d->mesh_kind != TRIXI_B200_MESH_TREE;
        auto rare = [&](int id) {
            return id == TRIXI_B200_FLUX_HLLC || id == TRIXI_B200_FLUX_HLLE ||
                   (curved_mesh && id == TRIXI_B200_FLUX_CHANDRASHEKAR);
        };
        const bool fv = d->volume_integral == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG ||
                        d->volume_integral == TRIXI_B200_VOLINT_PURE_LGL_FV;
        const bool all = rare(d->surface_flux) ||
                         (d->volume_integral != TRIXI_B200_VOLINT_WEAK_FORM && rare(d->volume_flux)) ||
                         (fv && rare(d->volume_flux_fv));
        if (d->equation == TRIXI_B200_EQ_EULER_2D)
            L = all ? get_launchers_euler2d_all(d->nnodes) : get_launchers_euler2d(d->nnodes);
        else
            L = all ? get_launchers_euler3d_all(d->nnodes) : get_launchers_euler3d(d->nnodes);
        break;
    }
    case TRIXI_B200_EQ_MHD_3D:
        if (d->nboundaries > 0 && d->mesh_kind != TRIXI_B200_MESH_P4EST)
            return fail(nullptr, TRIXI_B200_EINVAL,
                        "GLM-MHD boundary conditions are available on P4estMesh only in this build (Dirichlet)");
        L = get_launchers_mhd3d(d->nnodes);
        break;
    default: return fail(nullptr, TRIXI_B200_EINVAL, "equation %d not supported by this build", d->equation);
    }
    if (!L) return fail(nullptr, TRIXI_B200_EINVAL, "nnodes = %d not supported (2..8)", d->nnodes);
    if (L->ndims != d->ndims || L->nvars != d->nvars)
        return fail(nullptr, TRIXI_B200_EINVAL, "ndims/nvars (%d, %d) do not match equation %d", d->ndims, d->nvars,
                    d->equation);

    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev == 0)
        return fail(nullptr, TRIXI_B200_ENODEVICE, "no CUDA device available (%s); libtrixi_b200 has no CPU fallback",
                    err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
    int dev = d->device;
    if (dev < 0) {
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    }
    if (dev >= ndev) return fail(nullptr, TRIXI_B200_EINVAL, "device ordinal %d out of range (%d devices)", dev, ndev);

    trixi_b200_handle *h = new (std::nothrow) trixi_b200_handle();
    if (!h) return fail(nullptr, TRIXI_B200_ENOMEM, "out of host memory");
    h->device = dev;
#define CREATE_TRY(expr)                                 \
    do {                                                 \
        int rc__ = (expr);                               \
        if (rc__) {                                      \
            g_create_error = h->error;                   \
            trixi_b200_destroy(h);                       \
            return rc__;                                 \
        }                                                \
    } while (0)
#define CREATE_CUDA(expr)                                                                                   \
    do {                                                                                                    \
        cudaError_t e__ = (expr);                                                                           \
        if (e__ != cudaSuccess) {                                                                           \
            fail(nullptr, TRIXI_B200_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));                \
            trixi_b200_destroy(h);                                                                          \
            return TRIXI_B200_ECUDA;                                                                        \
        }                                                                                                   \
    } while (0)

    CREATE_CUDA(cudaSetDevice(dev));
    CREATE_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CREATE_CUDA(cudaEventCreate(&h->ev0));
    CREATE_CUDA(cudaEventCreate(&h->ev1));
    CREATE_CUDA(cudaEventCreate(&h->ev_t0));
    CREATE_CUDA(cudaEventCreate(&h->ev_t1));

    h->L = L;
    h->ndims = d->ndims;
    h->nvars = d->nvars;
    h->nnodes = d->nnodes;
    h->mesh_kind = d->mesh_kind;
    h->equation = d->equation;
    h->nelements = d->nelements;
    h->rank = d->rank;
    h->world_size = d->world_size > 0 ? d->world_size : 1;
    const int n = d->nnodes, nd = d->ndims, nv = d->nvars;
    const long long nn = ipow(n, nd), nf = ipow(n, nd - 1);
    h->ulen = (long long)nv * nn * d->nelements;
    h->sfvlen = (long long)nv * nf * 2 * nd * d->nelements;

    // StructuredMesh: faces are given through left_neighbors (dgsem_structured/containers.jl:8-34); build the
    // (left, right, orientation) list the interface kernel iterates over
    std::vector<long long> s_if_neighbors, s_if_orient;
    long long n_if = d->ninterfaces;
    if (structured) {
        for (long long e = 0; e < d->nelements; ++e)
            for (int o = 0; o < nd; ++o) {
                const long long left = d->left_neighbors[o + (long long)nd * e];
                if (left > 0) {
                    s_if_neighbors.push_back(left);
                    s_if_neighbors.push_back(e + 1);
                    s_if_orient.push_back(o + 1);
                }
            }
        n_if = (long long)s_if_orient.size();
    }
    KParams &P = h->P;
    P.nelements = d->nelements;
    P.elem_begin = 0;
    P.elem_end = d->nelements;
    P.ninterfaces = n_if;
    P.nboundaries = d->nboundaries;
    for (int i = 0; i < 3; ++i) CREATE_TRY(alloc_array(h, (size_t)h->ulen, &h->vec[i]));
    P.u = h->vec[0];
    P.du = h->vec[1];
    P.u_tmp = h->vec[2];
    P.u_out = h->vec[0];
    CREATE_TRY(alloc_array(h, (size_t)h->sfvlen, &P.sfv));
    if (h->sfvlen) CREATE_CUDA(cudaMemset(P.sfv, 0xff, h->sfvlen * sizeof(double)));  // NaN like the reference's fill

    double *tmp = nullptr;
    P.nmpimortars = d->nmpimortars;
    if (d->nmpimortars > 0) {
        const size_t np = (size_t)1 << (nd - 1);
        long long *mtmp = nullptr;
        CREATE_TRY(upload_array(h, (const long long *)d->mpi_mortar_neighbor_ids, (np + 1) * (size_t)d->nmpimortars, &mtmp));
        P.mpi_mortar_ids = mtmp;
        if (p4est) {
            CREATE_TRY(upload_array(h, (const long long *)d->mpi_mortar_node_indices, 2 * (size_t)nd * (size_t)d->nmpimortars, &mtmp));
            P.mpi_mortar_node_indices = mtmp;
            CREATE_TRY(upload_array(h, d->mpi_mortar_normal_directions,
                                    (size_t)nd * (size_t)ipow(n, nd - 1) * np * (size_t)d->nmpimortars, &tmp));
            P.mpi_mortar_normals = tmp;
        } else {
            CREATE_TRY(upload_array(h, (const long long *)d->mpi_mortar_large_sides, (size_t)d->nmpimortars, &mtmp));
            P.mpi_mortar_large_sides = mtmp;
            CREATE_TRY(upload_array(h, (const long long *)d->mpi_mortar_orientations, (size_t)d->nmpimortars, &mtmp));
            P.mpi_mortar_orient = mtmp;
        }
    }
    if (d->nmpiinterfaces > 0 && d->mpi_is_mortar_piece) {
        long long *mtmp = nullptr;
        CREATE_TRY(upload_array(h, (const long long *)d->mpi_is_mortar_piece, (size_t)d->nmpiinterfaces, &mtmp));
        P.mpi_is_piece = mtmp;
    }
    P.nmortars = d->nmortars;
    if (d->nmortars > 0 || d->nmpimortars > 0) {
        CREATE_TRY(upload_array(h, d->mortar_forward_lower, (size_t)n * n, &tmp));
        P.mortar_fwd[0] = tmp;
        CREATE_TRY(upload_array(h, d->mortar_forward_upper, (size_t)n * n, &tmp));
        P.mortar_fwd[1] = tmp;
        CREATE_TRY(upload_array(h, d->mortar_reverse_lower, (size_t)n * n, &tmp));
        P.mortar_rev[0] = tmp;
        CREATE_TRY(upload_array(h, d->mortar_reverse_upper, (size_t)n * n, &tmp));
        P.mortar_rev[1] = tmp;
    }
    if (d->nmortars > 0) {
        const size_t np1 = ((size_t)1 << (nd - 1)) + 1;
        long long *mtmp = nullptr;
        CREATE_TRY(upload_array(h, (const long long *)d->mortar_neighbor_ids, np1 * (size_t)d->nmortars, &mtmp));
        P.mortar_ids = mtmp;
        if (p4est) {
            CREATE_TRY(upload_array(h, (const long long *)d->mortar_node_indices, 2 * (size_t)nd * (size_t)d->nmortars, &mtmp));
            P.mortar_node_indices = mtmp;
        } else {
            CREATE_TRY(upload_array(h, (const long long *)d->mortar_large_sides, (size_t)d->nmortars, &mtmp));
            P.mortar_large_sides = mtmp;
            CREATE_TRY(upload_array(h, (const long long *)d->mortar_orientations, (size_t)d->nmortars, &mtmp));
            P.mortar_orient = mtmp;
        }
    }
    CREATE_TRY(upload_array(h, d->derivative_split, (size_t)n * n, &tmp));
    P.dsplit = tmp;
    CREATE_TRY(upload_array(h, d->derivative_hat, (size_t)n * n, &tmp));
    P.dhat = tmp;
    if (!d->inverse_weights) {
        fail(nullptr, TRIXI_B200_EINVAL, "inverse_weights missing");
        trixi_b200_destroy(h);
        return TRIXI_B200_EINVAL;
    }
    P.inv_weight0 = d->inverse_weights[0];
    for (int q = 0; q < n; ++q) P.weights_c[q] = 1.0 / d->inverse_weights[q];
    for (int q = 0; q < n * n; ++q) P.dsplit_c[q] = d->derivative_split[q];
    if (n == 4)
        for (int q = 0; q < 16; ++q) {
            P.dsplit_h[q] = 0.5 * d->derivative_split[q];
            P.dsplit_q[q] = 0.25 * d->derivative_split[q];
            P.dsplit_e[q] = 0.125 * d->derivative_split[q];
        }
    P.kernel_path = 0;
    P.rk_reduce_update = 1;
    P.l2_hints = 0;
    const bool curved = structured || p4est;
    CREATE_TRY(upload_array(h, d->inverse_jacobian, (size_t)(curved ? nn * d->nelements : d->nelements), &tmp));
    P.inverse_jacobian = tmp;
    P.curved = curved ? 1 : 0;
    P.p4est = p4est ? 1 : 0;
    P.contravariant_vectors = nullptr;
    P.if_node_indices = nullptr;
    P.bd_node_indices = nullptr;
    if (curved) {
        CREATE_TRY(upload_array(h, d->contravariant_vectors, (size_t)(nd * nd * nn * d->nelements), &tmp));
        P.contravariant_vectors = tmp;
    }
    CREATE_TRY(upload_array(h, d->node_coordinates, (size_t)(nd * nn * d->nelements), &tmp));
    P.node_coordinates = tmp;

    long long *itmp = nullptr;
    static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
    if (structured) {
        CREATE_TRY(upload_array(h, s_if_neighbors.data(), s_if_neighbors.size(), &itmp));
        P.if_neighbors = itmp;
        CREATE_TRY(upload_array(h, s_if_orient.data(), s_if_orient.size(), &itmp));
        P.if_orient = itmp;
    } else {
        CREATE_TRY(upload_array(h, (const long long *)d->interface_neighbor_ids, (size_t)(2 * d->ninterfaces), &itmp));
        P.if_neighbors = itmp;
        if (!p4est) {
            CREATE_TRY(upload_array(h, (const long long *)d->interface_orientations, (size_t)d->ninterfaces, &itmp));
            P.if_orient = itmp;
        }
    }
    P.minus_nb = nullptr;
    P.sfv_single = 0;
    if (!curved && d->equation != TRIXI_B200_EQ_MHD_3D && d->ninterfaces > 0 && d->nelements < (1ll << 31)) {
        // left neighbour across the - face of every element (conforming interior interfaces only)
        std::vector<int> mnb((size_t)nd * (size_t)d->nelements, -1);
        for (long long I = 0; I < d->ninterfaces; ++I) {
            const long long left = d->interface_neighbor_ids[2 * I] - 1, right = d->interface_neighbor_ids[2 * I + 1] - 1;
            const long long o = d->interface_orientations[I] - 1;
            if (left < 0 || right < 0 || left >= d->nelements || right >= d->nelements || o < 0 || o >= nd) {
                fail(nullptr, TRIXI_B200_EINVAL, "interface %lld has invalid neighbor ids or orientation", I + 1);
                trixi_b200_destroy(h);
                return TRIXI_B200_EINVAL;
            }
            mnb[(size_t)(o + nd * right)] = (int)left;
        }
        int *mtmp = nullptr;
        CREATE_TRY(upload_array(h, mnb.data(), mnb.size(), &mtmp));
        P.minus_nb = mtmp;
    }
    CREATE_TRY(upload_array(h, (const long long *)d->boundary_neighbor_ids, (size_t)d->nboundaries, &itmp));
    P.bd_neighbor = itmp;
    if (!p4est) {
        CREATE_TRY(upload_array(h, (const long long *)d->boundary_orientations, (size_t)d->nboundaries, &itmp));
        P.bd_orient = itmp;
        CREATE_TRY(upload_array(h, (const long long *)d->boundary_neighbor_sides, (size_t)d->nboundaries, &itmp));
        P.bd_side = itmp;
    }
    if (p4est) {
        CREATE_TRY(upload_array(h, (const long long *)d->interface_node_indices, (size_t)(nd * 2 * d->ninterfaces), &itmp));
        P.if_node_indices = itmp;
        CREATE_TRY(upload_array(h, (const long long *)d->boundary_node_indices, (size_t)(nd * d->nboundaries), &itmp));
        P.bd_node_indices = itmp;
    }
    if (!curved) {  // curved kernels read the element's own node coordinates
        CREATE_TRY(upload_array(h, d->boundary_node_coordinates, (size_t)(nd * nf * d->nboundaries), &tmp));
        P.bd_coords = tmp;
    }
    {
        // boundaries are sorted by direction (containers_3d.jl:398-468): expand the counts to a per-face direction
        std::vector<int> dir((size_t)d->nboundaries);
        long long pos = 0;
        for (int k = 0; k < 2 * nd; ++k)
            for (long long c = 0; c < d->n_boundaries_per_direction[k] && pos < d->nboundaries; ++c) dir[pos++] = k + 1;
        if (pos != d->nboundaries) {
            fail(nullptr, TRIXI_B200_EINVAL, "n_boundaries_per_direction does not sum to nboundaries");
            trixi_b200_destroy(h);
            return TRIXI_B200_EINVAL;
        }
        int *dtmp = nullptr;
        CREATE_TRY(upload_array(h, dir.data(), dir.size(), &dtmp));
        P.bd_direction = dtmp;
        for (int k = 0; k < 2 * nd; ++k)
            if (d->n_boundaries_per_direction[k] > 0 && d->boundary_conditions[k] == TRIXI_B200_BC_PERIODIC) {
                fail(nullptr, TRIXI_B200_EINVAL, "direction %d has boundary faces but a periodic boundary condition", k + 1);
                trixi_b200_destroy(h);
                return TRIXI_B200_EINVAL;
            }
    }
    for (int k = 0; k < 6; ++k) {
        P.bc[k] = d->boundary_conditions[k];
        P.bc_ic[k] = d->boundary_ic[k];
    }
    for (int k = 0; k < 8; ++k) P.eq.p[k] = d->eq_params[k];
    P.volume_integral = d->volume_integral;
    const bool sc_hg = d->volume_integral == TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG;
    if (sc_hg || d->volume_integral == TRIXI_B200_VOLINT_PURE_LGL_FV) {
        // the subcell finite-volume part (fv_kernel!): its flux, the inverse weights, the normal vectors on curved meshes
        P.volume_flux_fv = d->volume_flux_fv;
        for (int q = 0; q < n; ++q) P.inv_weights_c[q] = d->inverse_weights[q];
        if (sc_hg) {
            P.ind_var = d->indicator_variable;
            P.ind_smooth = d->indicator_alpha_smooth;
            P.ind_alpha_max = d->indicator_alpha_max;
            P.ind_alpha_min = d->indicator_alpha_min;
            CREATE_TRY(upload_array(h, d->inverse_vandermonde_legendre, (size_t)n * n, &tmp));
            P.inv_vdm = tmp;
            CREATE_TRY(alloc_array(h, (size_t)d->nelements, &P.alpha));
            CREATE_TRY(alloc_array(h, (size_t)d->nelements, &P.alpha_raw));
        }
        if (d->mesh_kind != TRIXI_B200_MESH_TREE) {
            for (int a = 0; a < nd; ++a) {
                size_t per_elem = (size_t)nd;
                for (int b = 0; b < nd; ++b) per_elem *= (size_t)(b == a ? n - 1 : n);
                CREATE_TRY(upload_array(h, d->subcell_normal_vectors[a], per_elem * (size_t)d->nelements, &tmp));
                P.subcell_normals[a] = tmp;
            }
        }
    }
    P.volume_flux = d->volume_flux;
    P.surface_flux = d->surface_flux;
    P.source_terms = d->source_terms;
    P.t = 0.0;
    P.mode = 0;
    P.rk_a = 0.0;
    P.rk_b_dt = 0.0;
    P.rk_read_tmp = 0;
    P.rk_write_tmp = 0;
    P.u_tmp2 = nullptr;
    for (double &k : P.rk_k) k = 0.0;

    // faces shared with other ranks: sorted by (neighbour rank, global interface id) by the caller
    h->nmpi = d->nmpiinterfaces;
    P.nmpi = d->nmpiinterfaces;
    if (d->world_size > 64) {
        fail(nullptr, TRIXI_B200_EINVAL, "world_size %d > 64 not supported", d->world_size);
        trixi_b200_destroy(h);
        return TRIXI_B200_EINVAL;
    }
    h->mpi_counts.assign(h->world_size, 0);
    if (d->nmpiinterfaces > 0) {
        if (!d->mpi_neighbor_ranks || !d->mpi_local_neighbor_ids || !d->mpi_local_sides || !d->mpi_orientations) {
            fail(nullptr, TRIXI_B200_EINVAL, "MPI interface arrays missing");
            trixi_b200_destroy(h);
            return TRIXI_B200_EINVAL;
        }
        h->h_peer_slot.resize((size_t)d->nmpiinterfaces);
        long long prev = -1;
        for (long long i = 0; i < d->nmpiinterfaces; ++i) {
            const long long r = d->mpi_neighbor_ranks[i];
            if (r < 0 || r >= h->world_size || r == h->rank || r < prev) {
                fail(nullptr, TRIXI_B200_EINVAL, "mpi_neighbor_ranks must be sorted ranks in [0, world_size) other than my own");
                trixi_b200_destroy(h);
                return TRIXI_B200_EINVAL;
            }
            if (r != prev) {
                h->peers.push_back((int)r);
                h->peer_offset.push_back(i);
            }
            prev = r;
            h->mpi_counts[(size_t)r]++;
            h->h_peer_slot[(size_t)i] = (int)h->peers.size() - 1;
        }
        CREATE_TRY(upload_array(h, (const long long *)d->mpi_local_neighbor_ids, (size_t)d->nmpiinterfaces, &itmp));
        P.mpi_local = itmp;
        CREATE_TRY(upload_array(h, (const long long *)d->mpi_local_sides, (size_t)d->nmpiinterfaces, &itmp));
        P.mpi_side = itmp;
        CREATE_TRY(upload_array(h, (const long long *)d->mpi_orientations, (size_t)d->nmpiinterfaces, &itmp));
        P.mpi_orient = itmp;
        if (p4est) {
            CREATE_TRY(upload_array(h, (const long long *)d->mpi_node_indices, (size_t)(nd * d->nmpiinterfaces), &itmp));
            P.mpi_node_indices = itmp;
        }
        int *stmp = nullptr;
        CREATE_TRY(upload_array(h, h->h_peer_slot.data(), h->h_peer_slot.size(), &stmp));
        P.mpi_peer_slot = stmp;
    }
    if (h->world_size > 1) {
        // one allocation other ranks map: [flags 2 x world] [recv parity 0] [recv parity 1]
        h->flag_bytes = ((size_t)2 * h->world_size * sizeof(unsigned long long) + 255) / 256 * 256;
        h->recv_bytes = ((size_t)h->nmpi * (nf * nv + 1) * sizeof(double) + 255) / 256 * 256;  // faces + alpha
        CREATE_TRY(alloc_array(h, h->flag_bytes + 2 * h->recv_bytes + 256, &h->comm_base));
        CREATE_CUDA(cudaMemset(h->comm_base, 0, h->flag_bytes + 2 * h->recv_bytes));
    }

    CREATE_CUDA(L->preload());
    CREATE_CUDA(preload_kernel(k_mpi_signal));
    CREATE_CUDA(preload_kernel(k_mpi_wait));

    CREATE_TRY(alloc_array(h, kCflSlots, &h->d_cfl));
    P.cfl_key = h->d_cfl;
    P.want_cfl = 0;
    {
        // L2 prefetch distance of the tuned element kernels: one wave of resident CTAs, resolved at launch
        int sms = 0;
        CREATE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
        P.sm_count = sms;
        P.prefetch_distance = -1;
    }
    CREATE_CUDA(cudaMallocHost((void **)&h->h_cfl, kCflSlots * sizeof(unsigned long long)));
    CREATE_CUDA(cudaDeviceSynchronize());
    *out = h;
    return TRIXI_B200_OK;
}

static int which_ok(trixi_b200_handle *h, int which) {
    if (!h) return TRIXI_B200_EINVAL;
    if (which < 0 || which > 2) return fail(h, TRIXI_B200_EINVAL, "vector selector %d out of range", which);
    return 0;
}

TRIXI_B200_API int trixi_b200_upload(trixi_b200_handle *h, int which, const double *host) {
    int rc = which_ok(h, which);
    if (rc) return rc;
    if (!host && h->ulen) return fail(h, TRIXI_B200_EINVAL, "host pointer is null");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (which == 0) h->cfl_valid = false;
    CUDA_TRY(h, cudaMemcpyAsync(h->vec[which], host, h->ulen * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

TRIXI_B200_API int trixi_b200_download(trixi_b200_handle *h, int which, double *host) {
    int rc = which_ok(h, which);
    if (rc) return rc;
    if (!host && h->ulen) return fail(h, TRIXI_B200_EINVAL, "host pointer is null");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaMemcpyAsync(host, h->vec[which], h->ulen * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

TRIXI_B200_API void *trixi_b200_device_ptr(trixi_b200_handle *h, int which) {
    if (!h || which < 0 || which > 2) return nullptr;
    if (which == 0) h->cfl_valid = false;  // the caller may write u through the pointer
    return h->vec[which];
}

TRIXI_B200_API int trixi_b200_synchronize(trixi_b200_handle *h) {
    if (!h) return TRIXI_B200_EINVAL;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

TRIXI_B200_API void *trixi_b200_stream(trixi_b200_handle *h) { return h ? (void *)h->stream : nullptr; }

TRIXI_B200_API int trixi_b200_rhs(trixi_b200_handle *h, double t) {
    if (!h) return TRIXI_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    int rc = run_rhs(h, t);
    if (rc) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    h->have_elapsed = true;
    return 0;
}

TRIXI_B200_API int trixi_b200_rhs_host(trixi_b200_handle *h, double *du_host, const double *u_host, double t) {
    if (!h) return TRIXI_B200_EINVAL;
    if ((!du_host || !u_host) && h->ulen) return fail(h, TRIXI_B200_EINVAL, "host pointer is null");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t bytes = h->ulen * sizeof(double);
    h->cfl_valid = false;
    int rc = 0;
    if (host_pipeline_usable(h, &rc)) {
        CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
        h->P.mode = 0;
        rc = run_pipelined(h, t, u_host, du_host, 1);
        if (rc) return rc;
        CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
        h->have_elapsed = true;
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->vec[0], u_host, bytes, cudaMemcpyHostToDevice, h->stream));
    rc = trixi_b200_rhs(h, t);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(du_host, h->vec[1], bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

TRIXI_B200_API int trixi_b200_calc_volume_integral(trixi_b200_handle *h) {
    if (!h) return TRIXI_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    h->P.mode = 0;
    return run_element(h, false);
}

TRIXI_B200_API int trixi_b200_calc_surface_fluxes(trixi_b200_handle *h, double t) {
    if (!h) return TRIXI_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    return run_all_surface_fluxes(h, t);
}

TRIXI_B200_API int trixi_b200_calc_indicator(trixi_b200_handle *h, double *alpha_host) {
    if (!h) return TRIXI_B200_EINVAL;
    if (h->P.volume_integral != TRIXI_B200_VOLINT_SHOCK_CAPTURING_HG)
        return fail(h, TRIXI_B200_EINVAL, "the volume integral has no indicator (not VolumeIntegralShockCapturingHG)");
    if (!alpha_host && h->nelements) return fail(h, TRIXI_B200_EINVAL, "host pointer is null");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->world_size > 1)
        return fail(h, TRIXI_B200_EINVAL, "calc_indicator is single-rank: across ranks the smoothing is part of the RHS's halo exchange");
    int rc = run_indicator(h, 1);
    if (rc == 0) rc = run_indicator(h, 2);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(alpha_host, h->P.alpha, h->nelements * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

TRIXI_B200_API int trixi_b200_download_surface_flux_values(trixi_b200_handle *h, double *host) {
    if (!h) return TRIXI_B200_EINVAL;
    if (!host && h->sfvlen) return fail(h, TRIXI_B200_EINVAL, "host pointer is null");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->P.sfv_single) {  // the reference's layout holds the flux on both sides of an interface
        h->L->sfv_fill_right(h->P, h->stream);
        h->launches++;
        int rc = check_launch(h, "surface flux fill kernel");
        if (rc) return rc;
    }
    CUDA_TRY(h, cudaMemcpyAsync(host, h->P.sfv, h->sfvlen * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

TRIXI_B200_API int trixi_b200_max_dt(trixi_b200_handle *h, double t, double *dt_out) {
    (void)t;
    if (!h || !dt_out) return TRIXI_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (!h->cfl_valid) {
        CUDA_TRY(h, cudaMemsetAsync(h->d_cfl, 0, kCflSlots * sizeof(unsigned long long), h->stream));
        {
            ProfScope ps(h, KC_MAXDT);
            h->L->max_dt(h->P, h->stream);
            h->launches++;
        }
        int rc = check_launch(h, "max_dt kernel");
        if (rc) return rc;
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->h_cfl, h->d_cfl, kCflSlots * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    // max_scaled_speed starts at nextfloat(0.0) (stepsize_dg3d.jl:12): bit pattern 1; the ordered bit patterns
    // of the partial maxima (NaN encoded above every finite value) reduce with an integer max
    unsigned long long key = 1ull;
    for (int i = 0; i < kCflSlots; ++i) key = h->h_cfl[i] > key ? h->h_cfl[i] : key;
    double max_scaled_speed;
    memcpy(&max_scaled_speed, &key, sizeof(double));
    *dt_out = 2 / (h->nnodes * max_scaled_speed);
    return 0;
}

TRIXI_B200_API int trixi_b200_calc_error_norms(trixi_b200_handle *h, double t, int initial_condition, int n_analysis,
                                               const double *vandermonde, const double *weights, double *l2_sums,
                                               double *linf, double *volume) {
    if (!h || !vandermonde || !weights || !l2_sums || !linf || !volume) return h ? fail(h, TRIXI_B200_EINVAL, "null argument") : TRIXI_B200_EINVAL;
    if (n_analysis < 1 || n_analysis > kMaxAnalysisNodes)
        return fail(h, TRIXI_B200_EINVAL, "n_analysis must be in 1..%d", kMaxAnalysisNodes);
    if (initial_condition != TRIXI_B200_IC_CONSTANT && initial_condition != TRIXI_B200_IC_CONVERGENCE_TEST &&
        !(initial_condition == TRIXI_B200_IC_WEAK_BLAST_WAVE &&
          (h->equation == TRIXI_B200_EQ_EULER_2D || h->equation == TRIXI_B200_EQ_EULER_3D)))
        return fail(h, TRIXI_B200_EINVAL, "initial condition %d is not registered on the device for this equation", initial_condition);
    if (h->equation == TRIXI_B200_EQ_MHD_3D && initial_condition != TRIXI_B200_IC_CONSTANT &&
        initial_condition != TRIXI_B200_IC_CONVERGENCE_TEST)
        return fail(h, TRIXI_B200_EINVAL, "initial condition %d is not registered on the device for this equation", initial_condition);
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int nv = h->nvars, n = h->nnodes;
    const size_t nbuf = (size_t)n_analysis * n + n_analysis + nv + 1;
    if (h->norm_buf_len < nbuf) {
        double *p = nullptr;
        int rc = alloc_array(h, nbuf, &p);
        if (rc) return rc;
        h->norm_buf = p;
        h->norm_buf_len = nbuf;
        unsigned long long *q = nullptr;
        rc = alloc_array(h, (size_t)16, &q);
        if (rc) return rc;
        h->norm_linf = q;
    }
    NormParams Q;
    Q.na = n_analysis;
    Q.ic = initial_condition;
    Q.t = t;
    double *dV = h->norm_buf, *dw = dV + (size_t)n_analysis * n, *dsum = dw + n_analysis;
    Q.vandermonde = dV;
    Q.weights = dw;
    Q.sums = dsum;
    Q.linf = h->norm_linf;
    CUDA_TRY(h, cudaMemcpyAsync(dV, vandermonde, sizeof(double) * n_analysis * n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(dw, weights, sizeof(double) * n_analysis, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemsetAsync(dsum, 0, sizeof(double) * (nv + 1), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->norm_linf, 0, sizeof(unsigned long long) * 16, h->stream));
    h->L->error_norms(h->P, Q, h->stream);
    h->launches++;
    int rc = check_launch(h, "error norm kernel");
    if (rc) return rc;
    double sums[16];
    unsigned long long mx[16];
    CUDA_TRY(h, cudaMemcpyAsync(sums, dsum, sizeof(double) * (nv + 1), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(mx, h->norm_linf, sizeof(unsigned long long) * nv, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int v = 0; v < nv; ++v) {
        l2_sums[v] = sums[v];
        memcpy(&linf[v], &mx[v], sizeof(double));
    }
    *volume = sums[nv];
    return 0;
}

TRIXI_B200_API int trixi_b200_integrate(trixi_b200_handle *h, int quantity, double *integral, double *volume) {
    if (!h || !integral || !volume) return h ? fail(h, TRIXI_B200_EINVAL, "null argument") : TRIXI_B200_EINVAL;
    if (quantity < TRIXI_B200_INTEGRAL_CONS || quantity > TRIXI_B200_INTEGRAL_ENTROPY_TIMEDERIVATIVE)
        return fail(h, TRIXI_B200_EINVAL, "unknown integrand %d", quantity);
    const bool euler = h->equation == TRIXI_B200_EQ_EULER_2D || h->equation == TRIXI_B200_EQ_EULER_3D;
    if (quantity != TRIXI_B200_INTEGRAL_CONS && !euler)
        return fail(h, TRIXI_B200_EINVAL, "integrand %d is registered for the compressible Euler equations only", quantity);
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int nv = h->nvars;
    if (!h->integral_buf) {
        int rc = alloc_array(h, (size_t)16, &h->integral_buf);
        if (rc) return rc;
    }
    CUDA_TRY(h, cudaMemsetAsync(h->integral_buf, 0, sizeof(double) * 16, h->stream));
    h->L->integrate(h->P, quantity, h->integral_buf, h->stream);
    h->launches++;
    int rc = check_launch(h, "integrate kernel");
    if (rc) return rc;
    double sums[16];
    CUDA_TRY(h, cudaMemcpyAsync(sums, h->integral_buf, sizeof(double) * (nv + 1), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const int nvals = quantity == TRIXI_B200_INTEGRAL_CONS ? nv : 1;
    for (int v = 0; v < nvals; ++v) integral[v] = sums[v];
    *volume = sums[nv];
    return 0;
}

TRIXI_B200_API int trixi_b200_step_2n(trixi_b200_handle *h, double t, double dt, const double *a, const double *b, const double *c,
                       int nstages) {
    if (!h || !a || !b || !c || nstages <= 0) return h ? fail(h, TRIXI_B200_EINVAL, "bad Runge-Kutta tableau") : TRIXI_B200_EINVAL;
    if (a[0] != 0.0) return fail(h, TRIXI_B200_EINVAL, "2N scheme must have a[1] == 0 (u_tmp starts at zero)");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    int rc = run_step_2n(h, t, dt, a, b, c, nstages, nullptr, nullptr);
    if (rc) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    h->have_elapsed = true;
    return 0;
}

TRIXI_B200_API int trixi_b200_step_2n_host(trixi_b200_handle *h, double *u_host, double t, double dt, const double *a,
                                            const double *b, const double *c, int nstages) {
    if (!h || !a || !b || !c || nstages <= 0) return h ? fail(h, TRIXI_B200_EINVAL, "bad Runge-Kutta tableau") : TRIXI_B200_EINVAL;
    if (a[0] != 0.0) return fail(h, TRIXI_B200_EINVAL, "2N scheme must have a[1] == 0 (u_tmp starts at zero)");
    if (!u_host && h->ulen) return fail(h, TRIXI_B200_EINVAL, "host pointer is null");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t bytes = h->ulen * sizeof(double);
    int rc = 0;
    const bool piped = host_pipeline_usable(h, &rc);
    if (rc) return rc;
    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    if (!piped) CUDA_TRY(h, cudaMemcpyAsync(h->vec[0], u_host, bytes, cudaMemcpyHostToDevice, h->stream));
    rc = run_step_2n(h, t, dt, a, b, c, nstages, piped ? u_host : nullptr, piped ? u_host : nullptr);
    if (rc) return rc;
    if (!piped) CUDA_TRY(h, cudaMemcpyAsync(u_host, h->vec[0], bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    h->have_elapsed = true;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

TRIXI_B200_API int trixi_b200_step_3sstar(trixi_b200_handle *h, double t, double dt, const double *gamma1, const double *gamma2,
                                           const double *gamma3, const double *beta, const double *delta, const double *c,
                                           int nstages) {
    if (!h || !gamma1 || !gamma2 || !gamma3 || !beta || !delta || !c || nstages <= 0)
        return h ? fail(h, TRIXI_B200_EINVAL, "bad Runge-Kutta tableau") : TRIXI_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (!h->vec_tmp2) {
        int rc = alloc_array(h, (size_t)h->ulen, &h->vec_tmp2);
        if (rc) return rc;
    }
    const size_t n = (size_t)h->ulen, bytes = n * sizeof(double);
    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    h->cfl_valid = false;
    // u_tmp2 .= u (methods_3Sstar.jl:188)
    CUDA_TRY(h, cudaMemcpyAsync(h->vec_tmp2, h->vec[0], bytes, cudaMemcpyDeviceToDevice, h->stream));
    if (stage_fusable(h)) {
        // The stage update (methods_3Sstar.jl:195-205) runs in the element kernel's epilogue: du never reaches memory
        // and u is read once.  u_tmp1 .= 0 (:187) is not swept either: the first stage takes it as zero.
        const bool fuse_cfl = h->opt_fused_cfl && h->L->fuses_cfl(h->P);
        h->P.u_tmp2 = h->vec_tmp2;
        for (int s = 0; s < nstages; ++s) {
            h->P.mode = 2;
            h->P.rk_k[0] = delta[s];
            h->P.rk_k[1] = gamma1[s];
            h->P.rk_k[2] = gamma2[s];
            h->P.rk_k[3] = gamma3[s];
            h->P.rk_b_dt = beta[s] * dt;
            h->P.rk_read_tmp = s > 0;
            h->P.rk_write_tmp = 1;
            int rc = run_fused_stage(h, t + dt * c[s], fuse_cfl && s == nstages - 1);
            if (rc) return rc;
        }
        h->cfl_valid = fuse_cfl;
        CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
        h->have_elapsed = true;
        return 0;
    }
    // u_tmp1 .= 0 (:187)
    CUDA_TRY(h, cudaMemsetAsync(h->vec[2], 0, bytes, h->stream));
    for (int s = 0; s < nstages; ++s) {
        int rc = run_rhs(h, t + dt * c[s]);
        if (rc) return rc;
        if (n) {
            k_stage_3sstar<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->vec[0], h->vec[2], h->vec_tmp2, h->vec[1], n,
                                                                                 delta[s], gamma1[s], gamma2[s], gamma3[s],
                                                                                 beta[s] * dt);
            h->launches++;
        }
        rc = check_launch(h, "3S* stage kernel");
        if (rc) return rc;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    h->have_elapsed = true;
    return 0;
}

TRIXI_B200_API int trixi_b200_step_ssp(trixi_b200_handle *h, double t, double dt, const double *numerator_a,
                                        const double *numerator_b, const double *denominator, const double *c, int nstages) {
    if (!h || !numerator_a || !numerator_b || !denominator || !c || nstages <= 0)
        return h ? fail(h, TRIXI_B200_EINVAL, "bad Runge-Kutta tableau") : TRIXI_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = (size_t)h->ulen, bytes = n * sizeof(double);
    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    h->cfl_valid = false;
    if (stage_fusable(h)) {
        // The stage update (methods_SSP.jl:192-201) runs in the element kernel's epilogue.  u_tmp .= u (:185) is not a
        // copy pass: the first stage takes u for u_tmp and writes it out beside the new u; later stages only read it.
        const bool fuse_cfl = h->opt_fused_cfl && h->L->fuses_cfl(h->P);
        for (int s = 0; s < nstages; ++s) {
            h->P.mode = 3;
            h->P.rk_k[0] = numerator_a[s];
            h->P.rk_k[1] = numerator_b[s];
            h->P.rk_k[2] = denominator[s];
            h->P.rk_k[3] = 0.0;
            h->P.rk_b_dt = dt;
            h->P.rk_read_tmp = s > 0;
            h->P.rk_write_tmp = s == 0;
            int rc = run_fused_stage(h, t + dt * c[s], fuse_cfl && s == nstages - 1);
            if (rc) return rc;
        }
        h->cfl_valid = fuse_cfl;
        CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
        h->have_elapsed = true;
        return 0;
    }
    // u_tmp .= u (methods_SSP.jl:185)
    CUDA_TRY(h, cudaMemcpyAsync(h->vec[2], h->vec[0], bytes, cudaMemcpyDeviceToDevice, h->stream));
    for (int s = 0; s < nstages; ++s) {
        int rc = run_rhs(h, t + dt * c[s]);
        if (rc) return rc;
        if (n) {
            k_stage_ssp<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->vec[0], h->vec[2], h->vec[1], n, dt,
                                                                              numerator_a[s], numerator_b[s], denominator[s]);
            h->launches++;
        }
        rc = check_launch(h, "SSP stage kernel");
        if (rc) return rc;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    h->have_elapsed = true;
    return 0;
}

TRIXI_B200_API int trixi_b200_solve_2n(trixi_b200_handle *h, double t0, double t_end, double cfl, int64_t max_steps, const double *a,
                        const double *b, const double *c, int nstages, int64_t *steps_out, double *t_out,
                        double *dt_out) {
    if (!h) return TRIXI_B200_EINVAL;
    if (h->world_size > 1)
        return fail(h, TRIXI_B200_EINVAL, "solve_2n is single-rank: distributed runs reduce dt over ranks on the host");
    double t = t0, dt = 0.0;
    int64_t steps = 0;
    bool finalstep = false;
    // the loop owns u for its whole duration: always let the last stage produce the next step's CFL maxima
    struct FusedCflScope {
        trixi_b200_handle *h;
        bool saved;
        explicit FusedCflScope(trixi_b200_handle *hh) : h(hh), saved(hh->opt_fused_cfl) { h->opt_fused_cfl = true; }
        ~FusedCflScope() { h->opt_fused_cfl = saved; }
    } fused_scope(h);
    while (!finalstep && steps < max_steps) {
        int rc = trixi_b200_max_dt(h, t, &dt);
        if (rc) return rc;
        dt *= cfl;
        if (std::isnan(dt)) return fail(h, TRIXI_B200_EINVAL, "time step size `dt` is NaN");
        // limit_dt! (time_integration.jl:46-55)
        const double tn = t + dt;
        if (tn > t_end || std::fabs(tn - t_end) <= std::sqrt(DBL_EPSILON) * std::fmax(std::fabs(tn), std::fabs(t_end))) {
            dt = t_end - t;
            finalstep = true;
        }
        rc = trixi_b200_step_2n(h, t, dt, a, b, c, nstages);
        if (rc) return rc;
        t += dt;
        ++steps;
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (steps_out) *steps_out = steps;
    if (t_out) *t_out = t;
    if (dt_out) *dt_out = dt;
    return 0;
}

TRIXI_B200_API int trixi_b200_set_eq_param(trixi_b200_handle *h, int index, double value) {
    if (!h) return TRIXI_B200_EINVAL;
    if (index < 0 || index >= 8) return fail(h, TRIXI_B200_EINVAL, "equation parameter index %d out of range", index);
    h->P.eq.p[index] = value;
    h->cfl_valid = false;  // wave speeds reduced by the last RK stage used the old parameter (c_h of GlmSpeedCallback)
    return 0;
}

TRIXI_B200_API int trixi_b200_set_option(trixi_b200_handle *h, int option, int value) {
    if (!h) return TRIXI_B200_EINVAL;
    switch (option) {
    case TRIXI_B200_OPT_KERNEL_PATH:
        if (value < 0 || value > 2) return fail(h, TRIXI_B200_EINVAL, "kernel path must be 0 (auto), 1 (generic) or 2 (previous-generation tuned kernels)");
        h->P.kernel_path = value;
        return 0;
    case TRIXI_B200_OPT_PREFETCH_DISTANCE:
        if (value < -1) return fail(h, TRIXI_B200_EINVAL, "prefetch distance must be >= 0, or -1 for one wave of CTAs");
        h->P.prefetch_distance = value;
        return 0;
    case TRIXI_B200_OPT_FUSED_CFL:
        if (value != 0 && value != 1) return fail(h, TRIXI_B200_EINVAL, "fused CFL option must be 0 or 1");
        h->opt_fused_cfl = value != 0;
        h->cfl_valid = false;
        return 0;
    case TRIXI_B200_OPT_L2_HINTS:
        if (value != 0 && value != 1) return fail(h, TRIXI_B200_EINVAL, "L2 hint option must be 0 or 1");
        h->P.l2_hints = value;
        return 0;
    case TRIXI_B200_OPT_SINGLE_FACE_FLUX:
        if (value != 0 && value != 1) return fail(h, TRIXI_B200_EINVAL, "single-face-flux option must be 0 or 1");
        h->opt_single_face_flux = value != 0;
        return 0;
    case TRIXI_B200_OPT_FUSED_STAGE:
        if (value != 0 && value != 1) return fail(h, TRIXI_B200_EINVAL, "fused-stage option must be 0 or 1");
        h->opt_fused_stage = value != 0;
        return 0;
    case TRIXI_B200_OPT_RK_REDUCE_UPDATE:
        if (value != 0 && value != 1) return fail(h, TRIXI_B200_EINVAL, "reduce-update option must be 0 or 1");
        h->P.rk_reduce_update = value;
        return 0;
    case TRIXI_B200_OPT_HOST_PIPELINE_CHUNK:
        if (value < -1) return fail(h, TRIXI_B200_EINVAL, "host pipeline chunk must be -1 (auto), 0 (off) or an element count");
        h->opt_pipeline_chunk = value;
        CUDA_TRY(h, cudaSetDevice(h->device));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        free_host_pipeline(h);  // rebuilt lazily by the next host-buffer call
        return 0;
    default:
        return fail(h, TRIXI_B200_EINVAL, "unknown option %d", option);
    }
}

// ---- halo exchange wiring ----------------------------------------------------------------------------
namespace {
struct CommBlob {
    int32_t magic, rank, world, device;
    int64_t pid;
    uint64_t raw_ptr;
    cudaIpcMemHandle_t handle;
    int64_t nmpi;
    int64_t counts[64];
};
constexpr int32_t kCommMagic = 0x7b200c01;
}  // namespace

TRIXI_B200_API int64_t trixi_b200_comm_info_size(void) { return (int64_t)sizeof(CommBlob); }

TRIXI_B200_API int trixi_b200_comm_info(trixi_b200_handle *h, void *blob_out) {
    if (!h || !blob_out) return TRIXI_B200_EINVAL;
    if (h->world_size <= 1) return fail(h, TRIXI_B200_ECOMM, "world_size is 1: nothing to exchange");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CommBlob b;
    memset(&b, 0, sizeof(b));
    b.magic = kCommMagic;
    b.rank = h->rank;
    b.world = h->world_size;
    b.device = h->device;
    b.pid = (int64_t)getpid();
    b.raw_ptr = (uint64_t)(uintptr_t)h->comm_base;
    CUDA_TRY(h, cudaIpcGetMemHandle(&b.handle, h->comm_base));
    b.nmpi = h->nmpi;
    for (int r = 0; r < h->world_size; ++r) b.counts[r] = h->mpi_counts[(size_t)r];
    memcpy(blob_out, &b, sizeof(b));
    return 0;
}

TRIXI_B200_API int trixi_b200_comm_connect(trixi_b200_handle *h, const void *blobs, int world) {
    if (!h || !blobs) return TRIXI_B200_EINVAL;
    if (world != h->world_size) return fail(h, TRIXI_B200_ECOMM, "got %d blobs for world_size %d", world, h->world_size);
    CUDA_TRY(h, cudaSetDevice(h->device));
    const CommBlob *B = static_cast<const CommBlob *>(blobs);
    for (int r = 0; r < world; ++r)
        if (B[r].magic != kCommMagic || B[r].rank != r || B[r].world != world)
            return fail(h, TRIXI_B200_ECOMM, "malformed comm blob for rank %d", r);
    const int npeers = (int)h->peers.size();
    const int n = h->nnodes, nd = h->ndims;
    const long long nf = ipow(n, nd - 1);
    h->peer_base.assign((size_t)npeers, nullptr);
    h->peer_nmpi.assign((size_t)npeers, 0);
    h->my_offset_in_peer.assign((size_t)npeers, 0);
    for (int p = 0; p < npeers; ++p) {
        const CommBlob &pb = B[h->peers[(size_t)p]];
        if (pb.counts[h->rank] != h->mpi_counts[(size_t)pb.rank])
            return fail(h, TRIXI_B200_ECOMM, "rank %d shares %lld faces with me, I share %lld with it", pb.rank,
                        (long long)pb.counts[h->rank], h->mpi_counts[(size_t)pb.rank]);
        long long off = 0;
        for (int q = 0; q < h->rank; ++q) off += pb.counts[q];
        h->my_offset_in_peer[(size_t)p] = off;
        h->peer_nmpi[(size_t)p] = pb.nmpi;
        if (pb.pid == (int64_t)getpid()) {
            // same process (several handles in one process): the pointer is directly usable
            if (pb.device != h->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(pb.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(h, TRIXI_B200_ECOMM, "cudaDeviceEnablePeerAccess(%d) failed: %s", pb.device,
                                cudaGetErrorString(e));
                cudaGetLastError();
            }
            h->peer_base[(size_t)p] = reinterpret_cast<char *>((uintptr_t)pb.raw_ptr);
        } else {
            void *mapped = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&mapped, pb.handle, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
                return fail(h, TRIXI_B200_ECOMM, "cudaIpcOpenMemHandle for rank %d failed: %s", pb.rank,
                            cudaGetErrorString(e));
            h->ipc_opened.push_back(mapped);
            h->peer_base[(size_t)p] = static_cast<char *>(mapped);
        }
    }
    // device tables
    std::vector<long long> remote_index((size_t)h->nmpi);
    for (long long i = 0; i < h->nmpi; ++i) {
        const int p = h->h_peer_slot[(size_t)i];
        remote_index[(size_t)i] = h->my_offset_in_peer[(size_t)p] + (i - h->peer_offset[(size_t)p]);
    }
    if (h->nmpi > 0) {
        int rc = upload_array(h, remote_index.data(), remote_index.size(), &h->d_mpi_remote_index);
        if (rc) return rc;
        h->P.mpi_remote_index = h->d_mpi_remote_index;
        rc = upload_array(h, h->peers.data(), h->peers.size(), &h->d_peer_ranks);
        if (rc) return rc;
        rc = upload_array(h, h->peer_nmpi.data(), h->peer_nmpi.size(), &h->d_peer_nmpi);
        if (rc) return rc;
        h->P.mpi_peer_nmpi = h->d_peer_nmpi;
        for (int parity = 0; parity < 2; ++parity) {
            std::vector<double *> recv((size_t)npeers);
            std::vector<unsigned long long *> flag((size_t)npeers);
            for (int p = 0; p < npeers; ++p) {
                const size_t peer_flag_bytes = ((size_t)2 * world * sizeof(unsigned long long) + 255) / 256 * 256;
                const size_t peer_recv_bytes =
                    ((size_t)h->peer_nmpi[(size_t)p] * (nf * h->nvars + 1) * sizeof(double) + 255) / 256 * 256;
                char *base = h->peer_base[(size_t)p];
                recv[(size_t)p] = reinterpret_cast<double *>(base + peer_flag_bytes + (size_t)parity * peer_recv_bytes);
                flag[(size_t)p] = reinterpret_cast<unsigned long long *>(base) + (size_t)parity * world + h->rank;
            }
            rc = upload_array(h, recv.data(), recv.size(), &h->d_peer_recv[parity]);
            if (rc) return rc;
            rc = upload_array(h, flag.data(), flag.size(), &h->d_peer_flag[parity]);
            if (rc) return rc;
        }
    }
    CUDA_TRY(h, cudaDeviceSynchronize());
    h->comm_connected = true;
    return 0;
}

TRIXI_B200_API int64_t trixi_b200_launch_count(const trixi_b200_handle *h) { return h ? h->launches : 0; }

TRIXI_B200_API int trixi_b200_last_elapsed_ms(trixi_b200_handle *h, float *ms_out) {
    if (!h || !ms_out) return TRIXI_B200_EINVAL;
    if (!h->have_elapsed) return fail(h, TRIXI_B200_EINVAL, "no timed call yet");
    CUDA_TRY(h, cudaEventSynchronize(h->ev1));
    CUDA_TRY(h, cudaEventElapsedTime(ms_out, h->ev0, h->ev1));
    return 0;
}

TRIXI_B200_API TRIXI_B200_API int trixi_b200_timer_start(trixi_b200_handle *h) {
    if (!h) return TRIXI_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaEventRecord(h->ev_t0, h->stream));
    return 0;
}

TRIXI_B200_API int trixi_b200_timer_stop(trixi_b200_handle *h, float *ms_out) {
    if (!h || !ms_out) return TRIXI_B200_EINVAL;
    CUDA_TRY(h, cudaEventRecord(h->ev_t1, h->stream));
    CUDA_TRY(h, cudaEventSynchronize(h->ev_t1));
    CUDA_TRY(h, cudaEventElapsedTime(ms_out, h->ev_t0, h->ev_t1));
    return 0;
}

TRIXI_B200_API int trixi_b200_measure_fp64_peak(trixi_b200_handle *h, double *tflops_out) {
    if (!h || !tflops_out) return TRIXI_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CUDA_TRY(h, cudaGetDeviceProperties(&prop, h->device));
    double *sink = nullptr;
    CUDA_TRY(h, cudaMalloc((void **)&sink, sizeof(double) * 1024));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(h->ev_t0, h->stream);
        k_fp64_peak<<<blocks, threads, 0, h->stream>>>(sink, iters, 1.0000001);
        cudaEventRecord(h->ev_t1, h->stream);
        cudaEventSynchronize(h->ev_t1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1);
        // 8 independent chains x 2 flop per DFMA
        const double flop = (double)blocks * threads * (double)iters * 8 * 2;
        if (rep > 0) best = fmax(best, flop / (ms * 1e-3) * 1e-12);
    }
    cudaFree(sink);
    int rc = check_launch(h, "fp64 peak kernel");
    if (rc) return rc;
    *tflops_out = best;
    return 0;
}

TRIXI_B200_API int trixi_b200_measure_copy_bandwidth(trixi_b200_handle *h, double *gbs_out) {
    if (!h || !gbs_out) return TRIXI_B200_EINVAL;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t n = (size_t)1 << 28;  // 2 GiB per buffer of doubles
    double *a = nullptr, *b = nullptr;
    CUDA_TRY(h, cudaMalloc((void **)&a, n * sizeof(double)));
    if (cudaMalloc((void **)&b, n * sizeof(double)) != cudaSuccess) {
        cudaFree(a);
        return fail(h, TRIXI_B200_ENOMEM, "copy benchmark allocation failed");
    }
    cudaMemsetAsync(a, 0, n * sizeof(double), h->stream);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(h->ev_t0, h->stream);
        k_copy<<<(unsigned)(n / 2 / 256), 256, 0, h->stream>>>((double2 *)b, (const double2 *)a, n / 2);
        cudaEventRecord(h->ev_t1, h->stream);
        cudaEventSynchronize(h->ev_t1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1);
        if (rep > 0) best = fmax(best, 2.0 * n * sizeof(double) / (ms * 1e-3) * 1e-9);
    }
    cudaFree(a);
    cudaFree(b);
    int rc = check_launch(h, "copy kernel");
    if (rc) return rc;
    *gbs_out = best;
    return 0;
}

int trixi_b200_profile_enable(trixi_b200_handle *h, int on) {
    if (!h) return TRIXI_B200_EINVAL;
    prof_collect(h);
    h->profiling = on != 0;
    for (int k = 0; k < KC_COUNT; ++k) {
        h->prof_ms[k] = 0;
        h->prof_n[k] = 0;
    }
    return 0;
}

TRIXI_B200_API int trixi_b200_profile_read(trixi_b200_handle *h, int kernel_class, double *ms_out, int64_t *launches_out) {
    if (!h || kernel_class < 0 || kernel_class >= KC_COUNT) return TRIXI_B200_EINVAL;
    prof_collect(h);
    if (ms_out) *ms_out = h->prof_ms[kernel_class];
    if (launches_out) *launches_out = h->prof_n[kernel_class];
    return 0;
}

}  // extern "C"
