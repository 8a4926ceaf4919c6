// Library plumbing (errors, device check), the FP64-pipe peak probe and the host-buffer entry point
// that strings the kernels together the way EnSRF.update() does (assimilation/ensrf.py:33-151).
#include "common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cmath>

static thread_local char g_err[512] = "";

void exb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int exb_check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        exb_set_error("%s: launch failed -> %s", what, cudaGetErrorString(e));
        return EXB_ERR_CUDA;
    }
    return EXB_OK;
}

static int64_t g_launches = 0;
void exb_count_launches(int64_t n) { __atomic_fetch_add(&g_launches, n, __ATOMIC_RELAXED); }
extern "C" int64_t exb_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int exb_version(void) { return 100; }
extern "C" const char *exb_last_error(void) { return g_err; }

extern "C" int exb_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        exb_set_error("exb_device_check: no CUDA device -> %s", cudaGetErrorString(e));
        cudaGetLastError();
        return EXB_ERR_NODEVICE;
    }
    // attribute queries, cached per device: cudaGetDeviceProperties takes tens of milliseconds while copies are
    // in flight, and this check sits in front of every analysis
    static int checked_major[64];
    static bool checked[64];
    if (dev < 0 || dev >= 64 || !checked[dev]) {
        int major = 0, minor = 0;
        EXB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
        EXB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
        if (major != 10) {
            exb_set_error("exb_device_check: device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
            return EXB_ERR_NODEVICE;
        }
        if (dev >= 0 && dev < 64) { checked_major[dev] = major; checked[dev] = true; }
    }
    return EXB_OK;
}

// ------------------------------------------------------------------------------------------
// FP64 FMA peak probe
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b) {
    double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    for (int i = 0; i < iters; ++i) {
        c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
        c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
    }
    const double r = ((c0 + c1) + (c2 + c3)) + ((c4 + c5) + (c6 + c7));
    if (r == 123.456) out[0] = r;      // never true; keeps the loop alive
}

extern "C" int exb_measure_fp64_peak(double *tflops, void *stream) {
    EXB_REQUIRE(tflops, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    double *out = nullptr;
    EXB_CUDA(cudaMalloc(&out, sizeof(double)));
    const int iters = 1 << 15, blocks = 148 * 8, threads = 256;
    cudaEvent_t e0, e1;
    EXB_CUDA(cudaEventCreate(&e0));
    EXB_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        EXB_CUDA(cudaEventRecord(e0, st));
        fp64_peak_kernel<<<blocks, threads, 0, st>>>(out, iters, 0.999999, 1e-9);
        exb_count_launches(1);
        EXB_CUDA(cudaEventRecord(e1, st));
        EXB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        EXB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * 8.0 * (double)iters * blocks * threads;
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return exb_check_launch("fp64_peak_kernel");
}

// FP64 tensor-core (DMMA, mma.sync m8n8k4) peak probe: 8 independent accumulator tiles per warp
__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = 0.5 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += c[i][0] + c[i][1];
    if (r == 123.456) out[0] = r;
}

extern "C" int exb_measure_dmma_peak(double *tflops, void *stream) {
    EXB_REQUIRE(tflops, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    double *out = nullptr;
    EXB_CUDA(cudaMalloc(&out, sizeof(double)));
    const int iters = 1 << 13, blocks = 148 * 8, threads = 256;
    cudaEvent_t e0, e1;
    EXB_CUDA(cudaEventCreate(&e0));
    EXB_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        EXB_CUDA(cudaEventRecord(e0, st));
        dmma_peak_kernel<<<blocks, threads, 0, st>>>(out, iters, 1e-3, 1e-3);
        exb_count_launches(1);
        EXB_CUDA(cudaEventRecord(e1, st));
        EXB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        EXB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        // one m8n8k4 = 8*8*4 FMA = 512 flop per warp
        const double flop = 512.0 * 8.0 * (double)iters * blocks * (threads / 32);
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return exb_check_launch("dmma_peak_kernel");
}

// ------------------------------------------------------------------------------------------
// whole analysis with host buffers
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// library-owned resources: memory pool, mapped host words, watchdog slots, pinned staging buffers
// ------------------------------------------------------------------------------------------
#include <mutex>
#include <cstdlib>
namespace {
std::mutex g_res_mutex;
cudaMemPool_t g_pool[64];
bool g_pool_ready[64];
}   // namespace

static cudaError_t exb_pool_get(cudaMemPool_t *pool) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(g_res_mutex);
    if (!g_pool_ready[dev]) {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        e = cudaMemPoolCreate(&g_pool[dev], &props);
        if (e != cudaSuccess) return e;
        // keep freed blocks cached (a 2.5 GB state buffer or a 0.5 GB list costs 20-100 ms to re-allocate on every
        // analysis otherwise), but only up to a bound, and only in this library's own pool
        double keep_gb = 32.0;
        if (const char *env = getenv("EXB_POOL_KEEP_GB")) keep_gb = atof(env);
        uint64_t keep = keep_gb <= 0.0 ? 0 : (uint64_t)(keep_gb * 1073741824.0);
        e = cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        if (e != cudaSuccess) return e;
        g_pool_ready[dev] = true;
    }
    *pool = g_pool[dev];
    return cudaSuccess;
}

cudaError_t exb_malloc_async(void **p, size_t bytes, cudaStream_t st) {
    cudaMemPool_t pool;
    cudaError_t e = exb_pool_get(&pool);
    if (e != cudaSuccess) return e;
    return cudaMallocFromPoolAsync(p, bytes ? bytes : 1, pool, st);
}

// Returns cached, currently unused blocks of the library's pool on the current device to the driver, keeping at most
// keep_bytes.
extern "C" int exb_pool_trim(uint64_t keep_bytes) {
    cudaMemPool_t pool;
    EXB_CUDA(exb_pool_get(&pool));
    EXB_CUDA(cudaMemPoolTrimTo(pool, (size_t)keep_bytes));
    return EXB_OK;
}

namespace {
constexpr int kWordSlots = 64, kStatusSlots = 1024;
long long *g_words_host = nullptr;        // [kWordSlots][8], portable + mapped
bool g_words_busy[kWordSlots];
int *g_status_host = nullptr;             // [kStatusSlots]
unsigned g_status_next = 0;
thread_local int g_status_last = -1;
struct PinnedBuf { void *p; size_t bytes; bool busy; };
std::vector<PinnedBuf> g_pinned;

int host_tables_init() {      // caller holds g_res_mutex
    if (!g_words_host) {
        EXB_CUDA(cudaHostAlloc(&g_words_host, sizeof(long long) * 8 * kWordSlots, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(g_words_host, 0, sizeof(long long) * 8 * kWordSlots);
    }
    if (!g_status_host) {
        EXB_CUDA(cudaHostAlloc(&g_status_host, sizeof(int) * kStatusSlots, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(g_status_host, 0, sizeof(int) * kStatusSlots);
    }
    return EXB_OK;
}
}   // namespace

int exb_host_words_acquire(ExbHostWords *w) {
    std::lock_guard<std::mutex> lock(g_res_mutex);
    const int rc = host_tables_init();
    if (rc != EXB_OK) return rc;
    for (int i = 0; i < kWordSlots; ++i)
        if (!g_words_busy[i]) {
            g_words_busy[i] = true;
            w->slot = i;
            w->host = g_words_host + 8 * i;
            void *d = nullptr;
            EXB_CUDA(cudaHostGetDevicePointer(&d, g_words_host + 8 * i, 0));
            w->dev = static_cast<long long *>(d);
            return EXB_OK;
        }
    exb_set_error("exb_host_words_acquire: more than %d calls in flight", kWordSlots);
    return EXB_ERR_UNSUPPORTED;
}

void exb_host_words_release(const ExbHostWords &w) {
    std::lock_guard<std::mutex> lock(g_res_mutex);
    if (w.slot >= 0 && w.slot < kWordSlots) g_words_busy[w.slot] = false;
}

int exb_status_slot_next(int **host, int **dev) {
    std::lock_guard<std::mutex> lock(g_res_mutex);
    const int rc = host_tables_init();
    if (rc != EXB_OK) return rc;
    const int i = (int)(g_status_next++ % kStatusSlots);
    void *d = nullptr;
    EXB_CUDA(cudaHostGetDevicePointer(&d, g_status_host + i, 0));
    *host = g_status_host + i;
    *dev = static_cast<int *>(d);
    g_status_last = i;
    return EXB_OK;
}

const int *exb_status_slot_last_dev() {
    if (g_status_last < 0 || !g_status_host) return nullptr;
    void *d = nullptr;
    if (cudaHostGetDevicePointer(&d, g_status_host + g_status_last, 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return static_cast<const int *>(d);
}

// Verdict of the watchdog of the last exb_obs_solve_* issued by THIS host thread (call after synchronising its
// stream): 0 ok, EXB_ERR_CUDA = a dependency wait never completed and the records are invalid.
extern "C" int exb_obs_solve_async_status(void) {
    if (g_status_last < 0 || !g_status_host) return EXB_OK;
    const int v = *reinterpret_cast<volatile int *>(g_status_host + g_status_last);
    if (v == 0) return EXB_OK;
    exb_set_error("exb_obs_solve: a dependency wait never completed (watchdog, record %d); the obs-space records are invalid",
                  v - 1);
    g_status_last = -1;                   // reported: later sweeps of this thread are not tied to that solve any more
    return EXB_ERR_CUDA;
}

int exb_pinned_acquire(size_t bytes, void **p) {
    std::lock_guard<std::mutex> lock(g_res_mutex);
    int best = -1;
    for (size_t i = 0; i < g_pinned.size(); ++i)
        if (!g_pinned[i].busy && g_pinned[i].bytes >= bytes && (best < 0 || g_pinned[i].bytes < g_pinned[best].bytes)) best = (int)i;
    if (best < 0) {
        // replace an idle buffer that is too small rather than growing the cache without bound
        for (size_t i = 0; i < g_pinned.size(); ++i)
            if (!g_pinned[i].busy) { cudaFreeHost(g_pinned[i].p); g_pinned.erase(g_pinned.begin() + i); break; }
        PinnedBuf b{nullptr, bytes, false};
        EXB_CUDA(cudaHostAlloc(&b.p, bytes ? bytes : 1, cudaHostAllocPortable));
        g_pinned.push_back(b);
        best = (int)g_pinned.size() - 1;
    }
    g_pinned[best].busy = true;
    *p = g_pinned[best].p;
    return EXB_OK;
}

void exb_pinned_release(void *p) {
    std::lock_guard<std::mutex> lock(g_res_mutex);
    for (auto &b : g_pinned)
        if (b.p == p) b.busy = false;
}

namespace {
struct DevBuf {
    void *p = nullptr;
    cudaStream_t st = nullptr;
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
    template <typename T> T *as() { return static_cast<T *>(p); }
    cudaError_t alloc(size_t bytes, cudaStream_t s) { st = s; return exb_malloc_async(&p, bytes, s); }
};
}   // namespace

#define EXB_TRY(call)              \
    do {                           \
        int rc__ = (call);         \
        if (rc__ != EXB_OK) return rc__; \
    } while (0)

extern "C" int exb_state_sweep_f64(double *X, int64_t nlev, int64_t ny, int64_t nx, int nens, const double *grid_u,
                                   const double *Yp, const double *rec, const double *obgeo, int64_t nobs,
                                   int64_t ob_begin, int64_t ob_end, int64_t y_begin, int64_t y_end, int loc_mode,
                                   unsigned long long *counters, void *stream);
extern "C" int exb_state_sweep_row_granularity(int64_t nlev, int64_t ny, int64_t nx);

// Three streams: band uploads, compute, band downloads (see engine.analysis_host for the same pipeline in Python).
//   pinned X_host, no inflation : the ob priors are gathered straight from host memory (unified addressing), then the
//                                 bands are uploaded in sweep order while the obs-space solve runs; a band is swept
//                                 as soon as it has arrived and downloaded while the next ones are swept
//   otherwise                   : the state is uploaded first; the band-wise sweep / download overlap remains
// Ensembles above 103 members (no fused sweep) use split -> sweep -> recombine on the whole state.
extern "C" int exb_ensrf_host_f64(double *X_host, int64_t nlev, int64_t ny, int64_t nx, int nens,
                                  const double *lat_deg, const double *lon_deg, int64_t nobs,
                                  const double *ob_value, const double *ob_error, const double *ob_lat_deg,
                                  const double *ob_lon_deg, const double *ob_halfwidth_km,
                                  const uint8_t *ob_assimilate, const int64_t *ob_row0, const int64_t *ob_row1,
                                  const double *ob_tw0, const double *ob_tw1, int loc_mode, double inflation,
                                  double *ob_diag, double *stats) {
    EXB_REQUIRE(X_host && lat_deg && lon_deg && ob_value && ob_error && ob_lat_deg && ob_lon_deg && ob_assimilate &&
                    ob_row0 && ob_row1 && ob_tw0 && ob_tw1 && ob_diag, "null pointer");
    EXB_REQUIRE(nlev > 0 && ny > 0 && nx > 0 && nens >= 2 && nobs > 0, "bad sizes");
    EXB_REQUIRE(loc_mode != EXB_LOC_GC || ob_halfwidth_km, "loc_mode GC needs halfwidths");
    EXB_TRY(exb_device_check());
    const int64_t npts = ny * nx, nrows = nlev * npts;
    const size_t row_bytes = (size_t)nens * sizeof(double);
    cudaStream_t st = nullptr, s_in = nullptr, s_out = nullptr;
    EXB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    EXB_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    EXB_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t a, b, c; ~StreamGuard() { cudaStreamDestroy(a); cudaStreamDestroy(b); cudaStreamDestroy(c); } } sguard{st, s_in, s_out};
    cudaEvent_t ev[5];
    for (auto &e : ev) EXB_CUDA(cudaEventCreate(&e));
    struct EvGuard { cudaEvent_t *e; ~EvGuard() { for (int i = 0; i < 5; ++i) cudaEventDestroy(e[i]); } } eguard{ev};

    DevBuf dX, dxm, dlat, dlon, dlaty, dlonx, dgu, dsl, dcl, dob, dassim, drow, dtw, dgeo, didx4, dw4, didx8, dw8, dY, dYm, drec, dcnt, dnex;
    EXB_CUDA(dX.alloc((size_t)nrows * row_bytes, st));
    EXB_CUDA(dlat.alloc(npts * sizeof(double), st));
    EXB_CUDA(dlon.alloc(npts * sizeof(double), st));
    EXB_CUDA(dgu.alloc(3 * npts * sizeof(double), st));
    EXB_CUDA(dsl.alloc(npts * sizeof(double), st));
    EXB_CUDA(dcl.alloc(npts * sizeof(double), st));
    EXB_CUDA(dob.alloc(7 * nobs * sizeof(double), st));      // value error lat lon hw sinlat coslon
    EXB_CUDA(dassim.alloc(nobs, st));
    EXB_CUDA(drow.alloc(2 * nobs * sizeof(int64_t), st));
    EXB_CUDA(dtw.alloc(2 * nobs * sizeof(double), st));
    EXB_CUDA(dgeo.alloc(EXB_GEO_FIELDS * nobs * sizeof(double), st));
    EXB_CUDA(didx4.alloc(4 * nobs * sizeof(int64_t), st));
    EXB_CUDA(dw4.alloc(4 * nobs * sizeof(double), st));
    EXB_CUDA(didx8.alloc(8 * nobs * sizeof(int64_t), st));
    EXB_CUDA(dw8.alloc(8 * nobs * sizeof(double), st));
    EXB_CUDA(dY.alloc((size_t)nobs * row_bytes, st));
    EXB_CUDA(dYm.alloc(nobs * sizeof(double), st));
    EXB_CUDA(drec.alloc(EXB_REC_FIELDS * nobs * sizeof(double), st));
    EXB_CUDA(dcnt.alloc(8 * sizeof(unsigned long long), st));
    EXB_CUDA(dnex.alloc(sizeof(int32_t), st));

    // host-side tables of the pseudo-metric (state/ensemble.py:160-163).  Rectilinear grids (lat a function of y,
    // lon a function of x: every regular lat-lon grid) take the separable O(ny+nx) search.
    bool rect = true;
    for (int64_t y = 0; y < ny && rect; ++y)
        for (int64_t x = 0; x < nx; ++x)
            if (lat_deg[y * nx + x] != lat_deg[y * nx] || lon_deg[y * nx + x] != lon_deg[x]) { rect = false; break; }
    const int64_t ntab = rect ? (ny > nx ? ny : nx) : npts;
    std::vector<double> sl(ntab), cl(ntab), osl(nobs), ocl(nobs), laty, lonx;
    if (rect) {
        laty.resize(ny); lonx.resize(nx);
        for (int64_t y = 0; y < ny; ++y) { laty[y] = lat_deg[y * nx]; sl[y] = sin(laty[y] * EXB_DEG2RAD); }
        for (int64_t x = 0; x < nx; ++x) { lonx[x] = lon_deg[x]; cl[x] = cos(lonx[x] * EXB_DEG2RAD); }
    } else {
        for (int64_t i = 0; i < npts; ++i) { sl[i] = sin(lat_deg[i] * EXB_DEG2RAD); cl[i] = cos(lon_deg[i] * EXB_DEG2RAD); }
    }
    for (int64_t k = 0; k < nobs; ++k) { osl[k] = sin(ob_lat_deg[k] * EXB_DEG2RAD); ocl[k] = cos(ob_lon_deg[k] * EXB_DEG2RAD); }

    // ---- small uploads first: nothing small may queue behind the state in the copy engine -------------------
    double *ob = dob.as<double>();
    EXB_CUDA(cudaEventRecord(ev[0], st));
    EXB_CUDA(cudaMemcpyAsync(dlat.p, lat_deg, npts * sizeof(double), cudaMemcpyHostToDevice, st));
    EXB_CUDA(cudaMemcpyAsync(dlon.p, lon_deg, npts * sizeof(double), cudaMemcpyHostToDevice, st));
    EXB_CUDA(cudaMemcpyAsync(dsl.p, sl.data(), (rect ? ny : npts) * sizeof(double), cudaMemcpyHostToDevice, st));
    EXB_CUDA(cudaMemcpyAsync(dcl.p, cl.data(), (rect ? nx : npts) * sizeof(double), cudaMemcpyHostToDevice, st));
    if (rect) {
        EXB_CUDA(dlaty.alloc(ny * sizeof(double), st));
        EXB_CUDA(dlonx.alloc(nx * sizeof(double), st));
        EXB_CUDA(cudaMemcpyAsync(dlaty.p, laty.data(), ny * sizeof(double), cudaMemcpyHostToDevice, st));
        EXB_CUDA(cudaMemcpyAsync(dlonx.p, lonx.data(), nx * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    EXB_CUDA(cudaMemcpyAsync(ob + 0 * nobs, ob_value, nobs * sizeof(double), cudaMemcpyHostToDevice, st));
    EXB_CUDA(cudaMemcpyAsync(ob + 1 * nobs, ob_error, nobs * sizeof(double), cudaMemcpyHostToDevice, st));
    EXB_CUDA(cudaMemcpyAsync(ob + 2 * nobs, ob_lat_deg, nobs * sizeof(double), cudaMemcpyHostToDevice, st));
    EXB_CUDA(cudaMemcpyAsync(ob + 3 * nobs, ob_lon_deg, nobs * sizeof(double), cudaMemcpyHostToDevice, st));
    if (loc_mode == EXB_LOC_GC)
        EXB_CUDA(cudaMemcpyAsync(ob + 4 * nobs, ob_halfwidth_km, nobs * sizeof(double), cudaMemcpyHostToDevice, st));
    EXB_CUDA(cudaMemcpyAsync(ob + 5 * nobs, osl.data(), nobs * sizeof(double), cudaMemcpyHostToDevice, st));
    EXB_CUDA(cudaMemcpyAsync(ob + 6 * nobs, ocl.data(), nobs * sizeof(double), cudaMemcpyHostToDevice, st));
    EXB_CUDA(cudaMemcpyAsync(dassim.p, ob_assimilate, nobs, cudaMemcpyHostToDevice, st));
    EXB_CUDA(cudaMemcpyAsync(drow.p, ob_row0, nobs * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    EXB_CUDA(cudaMemcpyAsync(drow.as<int64_t>() + nobs, ob_row1, nobs * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    EXB_CUDA(cudaMemcpyAsync(dtw.p, ob_tw0, nobs * sizeof(double), cudaMemcpyHostToDevice, st));
    EXB_CUDA(cudaMemcpyAsync(dtw.as<double>() + nobs, ob_tw1, nobs * sizeof(double), cudaMemcpyHostToDevice, st));
    EXB_CUDA(cudaMemsetAsync(dcnt.p, 0, 8 * sizeof(unsigned long long), st));
    EXB_TRY(exb_grid_unitvec(dlat.as<double>(), dlon.as<double>(), npts, dgu.as<double>(), st));
    EXB_TRY(exb_obs_prepare(ob + 2 * nobs, ob + 3 * nobs, ob + 4 * nobs, nobs, loc_mode, dgeo.as<double>(), st));
    // The geometry-only parts of the solve (predecessor lists) and of the sweep (candidate lists) are built on side
    // streams while the ob priors are gathered and the state is uploaded (exb_obs_plan_create, exb_sweep_plan_create).
    cudaStream_t s_p1 = nullptr, s_p2 = nullptr;
    EXB_CUDA(cudaStreamCreateWithFlags(&s_p1, cudaStreamNonBlocking));
    EXB_CUDA(cudaStreamCreateWithFlags(&s_p2, cudaStreamNonBlocking));
    struct PlanGuard {
        cudaStream_t a, b;
        void *oplan = nullptr, *splan = nullptr;
        ~PlanGuard() {
            if (oplan) exb_obs_plan_destroy(oplan);
            if (splan) exb_sweep_plan_destroy(splan);
            cudaStreamSynchronize(a); cudaStreamSynchronize(b);
            cudaStreamDestroy(a); cudaStreamDestroy(b);
        }
    } plans{s_p1, s_p2};
    cudaEvent_t ev_geo;
    EXB_CUDA(cudaEventCreateWithFlags(&ev_geo, cudaEventDisableTiming));
    struct GeoEvGuard { cudaEvent_t e; ~GeoEvGuard() { cudaEventDestroy(e); } } geguard{ev_geo};
    EXB_CUDA(cudaEventRecord(ev_geo, st));
    if (loc_mode == EXB_LOC_GC) {
        EXB_CUDA(cudaStreamWaitEvent(s_p1, ev_geo, 0));
        EXB_TRY(exb_obs_plan_create(dgeo.as<double>(), dassim.as<uint8_t>(), nobs, loc_mode, s_p1, &plans.oplan));
    }
    if (rect)
        EXB_TRY(exb_stencil_search_rect(dsl.as<double>(), dcl.as<double>(), dlaty.as<double>(), dlonx.as<double>(), ny, nx,
                                        ob + 5 * nobs, ob + 6 * nobs, ob + 2 * nobs, ob + 3 * nobs, nobs,
                                        didx4.as<int64_t>(), dw4.as<double>(), dnex.as<int32_t>(), st));
    else
        EXB_TRY(exb_stencil_search(dsl.as<double>(), dcl.as<double>(), dlat.as<double>(), dlon.as<double>(), npts,
                                   ob + 5 * nobs, ob + 6 * nobs, ob + 2 * nobs, ob + 3 * nobs, nobs, didx4.as<int64_t>(),
                                   dw4.as<double>(), dnex.as<int32_t>(), st));
    EXB_TRY(exb_stencil_combine(didx4.as<int64_t>(), dw4.as<double>(), drow.as<int64_t>(), drow.as<int64_t>() + nobs,
                                dtw.as<double>(), dtw.as<double>() + nobs, nobs, ny, nx, 0, ny, 0, didx8.as<int64_t>(),
                                dw8.as<double>(), st));

    // ---- band schedule ----------------------------------------------------------------------------------
    const bool fused = nens <= 103;
    std::vector<int64_t> edges;
    if (fused) {
        const int64_t g = exb_state_sweep_row_granularity(nlev, ny, nx);
        int64_t nb = ny / 180 < 1 ? 1 : (ny / 180 > 6 ? 6 : ny / 180);
        if (nb > ny / g) nb = ny / g > 0 ? ny / g : 1;
        edges.push_back(0);
        for (int64_t i = 1; i < nb; ++i) {
            const int64_t e = (int64_t)llround((double)ny * i / nb / g) * g;
            if (e > edges.back() && e < ny) edges.push_back(e);
        }
        const int64_t mid = ((edges.back() + ny) / 2 / g) * g;        // halve the last band: short exposed download
        if (mid > edges.back() && mid < ny) edges.push_back(mid);
        edges.push_back(ny);
    } else {
        edges = {0, ny};
    }
    const size_t nbands = edges.size() - 1;
    std::vector<cudaEvent_t> arrived(nbands), swept(nbands);
    for (size_t b = 0; b < nbands; ++b) {
        EXB_CUDA(cudaEventCreateWithFlags(&arrived[b], cudaEventDisableTiming));
        EXB_CUDA(cudaEventCreateWithFlags(&swept[b], cudaEventDisableTiming));
    }
    struct BandEvGuard { std::vector<cudaEvent_t> &a, &b; ~BandEvGuard() { for (auto e : a) cudaEventDestroy(e); for (auto e : b) cudaEventDestroy(e); } } bguard{arrived, swept};
    auto band_copy = [&](size_t b, bool to_device, cudaStream_t s) -> cudaError_t {
        for (int64_t lev = 0; lev < nlev; ++lev) {
            const size_t off = ((size_t)lev * npts + (size_t)edges[b] * nx) * nens;
            const size_t bytes = (size_t)(edges[b + 1] - edges[b]) * nx * row_bytes;
            cudaError_t e = to_device ? cudaMemcpyAsync(dX.as<double>() + off, X_host + off, bytes, cudaMemcpyHostToDevice, s)
                                      : cudaMemcpyAsync(X_host + off, dX.as<double>() + off, bytes, cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };

    // ---- ob priors and state upload -------------------------------------------------------------------------
    cudaPointerAttributes attr;
    bool pinned = false;
    const double *X_zero_copy = nullptr;
    if (cudaPointerGetAttributes(&attr, X_host) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer) {
        pinned = true;
        X_zero_copy = static_cast<const double *>(attr.devicePointer);
    }
    cudaGetLastError();
    const bool early_gather = pinned && inflation == 1.0;
    if (early_gather) {
        // PCIe carries one thing at a time: first the gather (<= 8 rows per ob), then the bands
        EXB_TRY(exb_gather_f64(X_zero_copy, nrows, nens, didx8.as<int64_t>(), dw8.as<double>(), 8, nobs, dY.as<double>(), st));
        EXB_CUDA(cudaEventRecord(ev[1], st));
        EXB_CUDA(cudaStreamWaitEvent(s_in, ev[1], 0));
    } else {
        EXB_CUDA(cudaEventRecord(ev[1], st));
        EXB_CUDA(cudaStreamWaitEvent(s_in, ev[1], 0));
    }
    for (size_t b = 0; b < nbands; ++b) {
        EXB_CUDA(band_copy(b, true, s_in));
        EXB_CUDA(cudaEventRecord(arrived[b], s_in));
    }
    if (!early_gather) {
        EXB_CUDA(cudaStreamWaitEvent(st, arrived[nbands - 1], 0));
        if (inflation != 1.0) EXB_TRY(exb_inflate_f64(dX.as<double>(), nrows, nens, &inflation, 1, nrows, st));
        EXB_TRY(exb_gather_f64(dX.as<double>(), nrows, nens, didx8.as<int64_t>(), dw8.as<double>(), 8, nobs, dY.as<double>(), st));
    }
    EXB_TRY(exb_split_mean_pert_f64(dY.as<double>(), dYm.as<double>(), nobs, nens, st));
    EXB_CUDA(cudaEventRecord(ev[2], st));
    if (fused) {
        // (blocks the host until the plan's counting pass is done -- the device is busy with the gather / uploads)
        EXB_CUDA(cudaStreamWaitEvent(s_p2, ev_geo, 0));
        EXB_TRY(exb_sweep_plan_create(dgu.as<double>(), nlev, ny, nx, dgeo.as<double>(), dassim.as<uint8_t>(), nobs, 0, nobs, 0, ny,
                                      loc_mode, s_p2, &plans.splan));
    }

    // ---- the serial analysis: obs-space solve, then the state band by band ----------------------------------
    if (plans.oplan)
        EXB_TRY(exb_obs_solve_planned_f64(plans.oplan, dYm.as<double>(), dY.as<double>(), ob + 0 * nobs, ob + 1 * nobs,
                                          dassim.as<uint8_t>(), dgeo.as<double>(), nobs, nens, loc_mode, drec.as<double>(),
                                          dcnt.as<unsigned long long>(), st));
    else
        EXB_TRY(exb_obs_solve_f64(dYm.as<double>(), dY.as<double>(), ob + 0 * nobs, ob + 1 * nobs, dassim.as<uint8_t>(),
                                  dgeo.as<double>(), nobs, nens, loc_mode, drec.as<double>(),
                                  dcnt.as<unsigned long long>(), st));
    if (fused) {
        for (size_t b = 0; b < nbands; ++b) {
            EXB_CUDA(cudaStreamWaitEvent(st, arrived[b], 0));
            if (plans.splan)
                EXB_TRY(exb_state_sweep_planned_f64(plans.splan, dX.as<double>(), nlev, ny, nx, nens, dgu.as<double>(), dY.as<double>(),
                                                    drec.as<double>(), dgeo.as<double>(), nobs, 0, nobs, edges[b], edges[b + 1],
                                                    loc_mode, dcnt.as<unsigned long long>(), st));
            else
                EXB_TRY(exb_state_sweep_f64(dX.as<double>(), nlev, ny, nx, nens, dgu.as<double>(), dY.as<double>(), drec.as<double>(),
                                            dgeo.as<double>(), nobs, 0, nobs, edges[b], edges[b + 1], loc_mode,
                                            dcnt.as<unsigned long long>(), st));
            EXB_CUDA(cudaEventRecord(swept[b], st));
            // download of the band that finished before this one (pageable destinations block the host here while
            // this band is being swept, pinned ones do not block at all)
            if (b > 0) {
                EXB_CUDA(cudaStreamWaitEvent(s_out, swept[b - 1], 0));
                EXB_CUDA(band_copy(b - 1, false, s_out));
            }
        }
        EXB_CUDA(cudaEventRecord(ev[3], st));
        EXB_CUDA(cudaStreamWaitEvent(s_out, swept[nbands - 1], 0));
        EXB_CUDA(band_copy(nbands - 1, false, s_out));
    } else {
        EXB_CUDA(dxm.alloc((size_t)nrows * sizeof(double), st));
        EXB_CUDA(cudaStreamWaitEvent(st, arrived[nbands - 1], 0));
        EXB_TRY(exb_split_mean_pert_f64(dX.as<double>(), dxm.as<double>(), nrows, nens, st));
        EXB_TRY(exb_state_update_f64(dxm.as<double>(), dX.as<double>(), nlev, ny, nx, nens, dgu.as<double>(),
                                     dY.as<double>(), drec.as<double>(), dgeo.as<double>(), nobs, 0, nobs, loc_mode,
                                     dcnt.as<unsigned long long>(), st));
        EXB_TRY(exb_recombine_f64(dX.as<double>(), dxm.as<double>(), nrows, nens, st));
        EXB_CUDA(cudaEventRecord(ev[3], st));
        EXB_CUDA(cudaEventRecord(swept[0], st));
        EXB_CUDA(cudaStreamWaitEvent(s_out, swept[0], 0));
        EXB_CUDA(band_copy(0, false, s_out));
    }
    EXB_CUDA(cudaEventRecord(ev[4], s_out));
    EXB_CUDA(cudaStreamWaitEvent(st, ev[4], 0));
    EXB_CUDA(cudaMemcpyAsync(ob_diag, drec.p, 4 * nobs * sizeof(double), cudaMemcpyDeviceToHost, st));
    unsigned long long cnt[2] = {0, 0};
    int32_t nex = 0;
    EXB_CUDA(cudaMemcpyAsync(cnt, dcnt.p, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    EXB_CUDA(cudaMemcpyAsync(&nex, dnex.p, sizeof(nex), cudaMemcpyDeviceToHost, st));
    EXB_CUDA(cudaStreamSynchronize(st));
    EXB_CUDA(cudaStreamSynchronize(s_in));
    EXB_CUDA(cudaStreamSynchronize(s_out));
    EXB_TRY(exb_obs_solve_async_status());
    if (stats) {
        float ms[4] = {0.f, 0.f, 0.f, 0.f};
        EXB_CUDA(cudaEventElapsedTime(&ms[0], ev[0], ev[1]));      // small uploads + stencils (+ zero-copy gather)
        EXB_CUDA(cudaEventElapsedTime(&ms[1], ev[1], ev[2]));      // state upload wait / ob-prior split
        EXB_CUDA(cudaEventElapsedTime(&ms[2], ev[2], ev[3]));      // obs-space solve + sweeps (uploads/downloads overlapped)
        EXB_CUDA(cudaEventElapsedTime(&ms[3], ev[3], ev[4]));      // download left exposed after the last sweep
        stats[0] = (double)cnt[1] * (double)nlev;
        stats[1] = (double)cnt[0];
        stats[2] = (double)nex;
        for (int i = 0; i < 4; ++i) stats[3 + i] = ms[i];
        stats[7] = (double)nbands;
    }
    return EXB_OK;
}
