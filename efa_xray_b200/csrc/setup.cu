// Setup / teardown kernels around the serial loop: geometry tables, the forward operator
// (nearest-4 search + inverse-distance weights + K-point gather), mean/perturbation split,
// inflation and recombination.  All HBM- or latency-bound elementwise/gather work.
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------
__global__ void grid_unitvec_kernel(const double *__restrict__ lat, const double *__restrict__ lon,
                                    int64_t npts, double *__restrict__ u) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= npts) return;
    double sp, cp, sl, cl;
    sincos(lat[i] * EXB_DEG2RAD, &sp, &cp);
    sincos(lon[i] * EXB_DEG2RAD, &sl, &cl);
    u[i] = cp * cl;
    u[npts + i] = cp * sl;
    u[2 * npts + i] = sp;
}

__global__ void obs_prepare_kernel(const double *__restrict__ lat, const double *__restrict__ lon,
                                   const double *__restrict__ hw, int64_t nobs, int loc_mode,
                                   double *__restrict__ geo) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nobs) return;
    double sp, cp, sl, cl;
    sincos(lat[i] * EXB_DEG2RAD, &sp, &cp);
    sincos(lon[i] * EXB_DEG2RAD, &sl, &cl);
    geo[GEO_UX * nobs + i] = cp * cl;
    geo[GEO_UY * nobs + i] = cp * sl;
    geo[GEO_UZ * nobs + i] = sp;
    double inv_hw = 0.0, a_max = 2.0, theta = M_PI;
    if (loc_mode == EXB_LOC_GC) {
        const double h = fabs(hw[i]);                 // abs(halfwidth), observation.py:120
        if (h == 0.0) {                               // r = d/0: every weight is 0, even at d = 0
            inv_hw = 0.0; a_max = 0.0; theta = 0.0;
        } else {
            inv_hw = 1.0 / h;
            const double half = h / EXB_R_EARTH;      // (support radius 2h) / 2 as an angle
            if (half < 0.5 * M_PI) {
                const double s = sin(half);
                a_max = s * s;
                theta = 2.0 * half;
            }
        }
    }
    geo[GEO_INVHW * nobs + i] = inv_hw;
    geo[GEO_AMAX * nobs + i] = a_max;
    geo[GEO_COST * nobs + i] = cos(theta);
    geo[GEO_SINT * nobs + i] = sin(theta);
    geo[GEO_THETA * nobs + i] = theta;
}

// sin(radians(lat)), cos(radians(lon)) of the obs: the ob-side tables of the nearest-point search
// (state/ensemble.py:160-163), same operations as numpy's: x * (pi / 180), then sin / cos
__global__ void obs_trig_kernel(const double *__restrict__ lat, const double *__restrict__ lon, int64_t nobs,
                                double *__restrict__ sinlat, double *__restrict__ coslon) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nobs) return;
    sinlat[i] = sin(lat[i] * EXB_DEG2RAD);
    coslon[i] = cos(lon[i] * EXB_DEG2RAD);
}

extern "C" int exb_obs_trig(const double *ob_lat_deg, const double *ob_lon_deg, int64_t nobs, double *ob_sinlat,
                            double *ob_coslon, void *stream) {
    EXB_REQUIRE(ob_lat_deg && ob_lon_deg && ob_sinlat && ob_coslon && nobs > 0, "null pointer or nobs <= 0");
    obs_trig_kernel<<<(unsigned)ceil_div64(nobs, 256), 256, 0, (cudaStream_t)stream>>>(ob_lat_deg, ob_lon_deg, nobs, ob_sinlat, ob_coslon);
    exb_count_launches(1);
    return exb_check_launch("obs_trig_kernel");
}

// ------------------------------------------------------------------------------------------
// nearest-4 search under the reference's pseudo-metric (state/ensemble.py:160-165)
// ------------------------------------------------------------------------------------------
// One CTA = NS_TX point-lanes x NS_OY observation groups, each thread keeps the running top-4 of
// NS_OB observations.  A CTA therefore serves NS_OY*NS_OB observations per pass over the grid tables,
// which are L2-resident (2 x 8 MB at 0.25 degrees).
#define NS_TX 128
#define NS_OY 4
#define NS_OB 8
#define NS_OBS_PER_CTA (NS_OY * NS_OB)

__device__ __forceinline__ void top4_insert(double (&d)[4], int (&ix)[4], double v, int p) {
    // keeps (d, ix) sorted ascending; a candidate equal to an existing entry goes after it, and
    // every thread visits points in increasing index order, so ties resolve to the lowest index.
    if (v < d[3]) {
        if (v < d[2]) {
            d[3] = d[2]; ix[3] = ix[2];
            if (v < d[1]) {
                d[2] = d[1]; ix[2] = ix[1];
                if (v < d[0]) { d[1] = d[0]; ix[1] = ix[0]; d[0] = v; ix[0] = p; }
                else { d[1] = v; ix[1] = p; }
            } else { d[2] = v; ix[2] = p; }
        } else { d[3] = v; ix[3] = p; }
    }
}

__device__ __forceinline__ bool key_less(double da, int ia, double db, int ib) {
    return (da < db) || (da == db && ia < ib);
}

__device__ double haversine_ref(double lat1_deg, double lon1_deg, double lat2_deg, double lon2_deg) {
    // state/ensemble.py:241-252 with loc1 = grid point, loc2 = ob
    const double lat1 = lat1_deg * EXB_DEG2RAD, lat2 = lat2_deg * EXB_DEG2RAD;
    const double dlat = lat2 - lat1;
    const double dlon = (lon2_deg - lon1_deg) * EXB_DEG2RAD;
    const double s1 = sin(dlat / 2), s2 = sin(dlon / 2);
    const double a = s1 * s1 + cos(lat1) * cos(lat2) * s2 * s2;
    return EXB_R_EARTH * (2.0 * atan2(sqrt(a), sqrt(1.0 - a)));
}

__global__ void __launch_bounds__(NS_TX *NS_OY)
nearest4_kernel(const double *__restrict__ sl_g, const double *__restrict__ cl_g,
                const double *__restrict__ lat_g, const double *__restrict__ lon_g, int npts,
                const double *__restrict__ ob_sl, const double *__restrict__ ob_cl,
                const double *__restrict__ ob_lat, const double *__restrict__ ob_lon, int64_t nobs,
                int64_t *__restrict__ idx4, double *__restrict__ w4, int32_t *__restrict__ n_exact) {
    extern __shared__ unsigned char smem_raw[];
    // candidates: [NS_OBS_PER_CTA][NS_TX*4]
    double *cand_d = reinterpret_cast<double *>(smem_raw);
    int *cand_i = reinterpret_cast<int *>(cand_d + NS_OBS_PER_CTA * NS_TX * 4);

    const int tx = threadIdx.x % NS_TX;
    const int oy = threadIdx.x / NS_TX;
    const int64_t ob0 = blockIdx.x * (int64_t)NS_OBS_PER_CTA + oy * NS_OB;

    double osl[NS_OB], ocl[NS_OB];
    double bd[NS_OB][4];
    int bi[NS_OB][4];
#pragma unroll
    for (int o = 0; o < NS_OB; ++o) {
        const int64_t k = ob0 + o < nobs ? ob0 + o : nobs - 1;
        osl[o] = ob_sl[k];
        ocl[o] = ob_cl[k];
#pragma unroll
        for (int j = 0; j < 4; ++j) { bd[o][j] = INFINITY; bi[o][j] = 0x7fffffff; }
    }
    for (int p = tx; p < npts; p += NS_TX) {
        const double sl = sl_g[p], cl = cl_g[p];
#pragma unroll
        for (int o = 0; o < NS_OB; ++o) {
            const double a = sl - osl[o], b = cl - ocl[o];
            // explicit roundings: no FMA contraction, so equal inputs give equal keys
            const double v = __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b));
            top4_insert(bd[o], bi[o], v, p);
        }
    }
#pragma unroll
    for (int o = 0; o < NS_OB; ++o) {
        const int slot = (oy * NS_OB + o) * (NS_TX * 4) + tx * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) { cand_d[slot + j] = bd[o][j]; cand_i[slot + j] = bi[o][j]; }
    }
    __syncthreads();

    // one warp per observation selects the 4 smallest keys of its NS_TX*4 candidates
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int nwarps = (NS_TX * NS_OY) / 32;
    for (int o = warp; o < NS_OBS_PER_CTA; o += nwarps) {
        const int64_t k = blockIdx.x * (int64_t)NS_OBS_PER_CTA + o;
        if (k >= nobs) continue;
        const double *cd = cand_d + o * (NS_TX * 4);
        const int *ci = cand_i + o * (NS_TX * 4);
        double prev_d = -1.0;
        int prev_i = -1;
        int sel[4];
        for (int r = 0; r < 4; ++r) {
            double md = INFINITY;
            int mi = 0x7fffffff;
            for (int c = lane; c < NS_TX * 4; c += 32) {
                const double d = cd[c];
                const int i = ci[c];
                if (key_less(prev_d, prev_i, d, i) && key_less(d, i, md, mi)) { md = d; mi = i; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, md, off);
                const int oi = __shfl_xor_sync(0xffffffffu, mi, off);
                if (key_less(od, oi, md, mi)) { md = od; mi = oi; }
            }
            sel[r] = mi;
            prev_d = md;
            prev_i = mi;
        }
        if (lane == 0) {
            // true distances and inverse-distance weights, state/ensemble.py:181-200
            double dist[4];
            int amin = 0;
            bool exact = false;
            for (int r = 0; r < 4; ++r) {
                dist[r] = haversine_ref(lat_g[sel[r]], lon_g[sel[r]], ob_lat[k], ob_lon[k]);
                if (dist[r] < 1.0) exact = true;
                if (dist[r] < dist[amin]) amin = r;
            }
            double w[4];
            if (exact) {
                for (int r = 0; r < 4; ++r) w[r] = (r == amin) ? 1.0 : 0.0;
                if (n_exact) atomicAdd(n_exact, 1);
            } else {
                double s = 0.0;
                for (int r = 0; r < 4; ++r) { w[r] = 1.0 / dist[r]; s += w[r]; }
                for (int r = 0; r < 4; ++r) w[r] /= s;
            }
            for (int r = 0; r < 4; ++r) { idx4[k * 4 + r] = sel[r]; w4[k * 4 + r] = w[r]; }
        }
    }
}

// Squared pseudo-distance of ONE observation to every grid point, with the roundings of nearest4_kernel: the sort key
// of nearest_points for any npt (state/ensemble.py:160-165; hypot is monotone in it).
__global__ void pseudo_distance_kernel(const double *__restrict__ sl_g, const double *__restrict__ cl_g, int64_t npts,
                                       double osl, double ocl, double *__restrict__ out) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const double a = sl_g[p] - osl, b = cl_g[p] - ocl;
    out[p] = __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b));
}

extern "C" int exb_pseudo_distance(const double *sinlat_g, const double *coslon_g, int64_t npts, double ob_sinlat,
                                   double ob_coslon, double *d2, void *stream) {
    EXB_REQUIRE(sinlat_g && coslon_g && d2 && npts > 0, "null pointer or npts <= 0");
    pseudo_distance_kernel<<<(unsigned)ceil_div64(npts, 256), 256, 0, (cudaStream_t)stream>>>(sinlat_g, coslon_g, npts,
                                                                                                ob_sinlat, ob_coslon, d2);
    exb_count_launches(1);
    return exb_check_launch("pseudo_distance_kernel");
}

// Rectilinear grids (lat depends on y only, lon on x only -- every regular lat-lon grid): the squared
// pseudo-distance separates, d2(y, x) = A[y] + B[x] with A = (sin lat_y - sin lat_ob)^2 and
// B = (cos lon_x - cos lon_ob)^2, so the 4 smallest d2 are among the 8 smallest A x the 8 smallest B.
// Same roundings and the same (d2, flat index) order as nearest4_kernel, at O(ny + nx) per ob.
#define NR_K 8
#define NR_WARPS 8

template <int K>
__device__ __forceinline__ void topk_insert(double (&d)[K], int (&ix)[K], double v, int p) {
    if (!(v < d[K - 1])) return;
    bool placed = false;
#pragma unroll
    for (int j = K - 1; j > 0; --j) {
        if (!placed) {
            if (v < d[j - 1]) { d[j] = d[j - 1]; ix[j] = ix[j - 1]; }
            else { d[j] = v; ix[j] = p; placed = true; }
        }
    }
    if (!placed) { d[0] = v; ix[0] = p; }
}

// the warp picks the `nsel` smallest (d, i) keys among ncand candidates in shared memory, in order
__device__ __forceinline__ void warp_select(const double *cd, const int *ci, int ncand, int nsel, int lane,
                                            double *out_d, int *out_i) {
    double prev_d = -1.0;
    int prev_i = -1;
    for (int r = 0; r < nsel; ++r) {
        double md = INFINITY;
        int mi = 0x7fffffff;
        for (int c = lane; c < ncand; c += 32) {
            const double d = cd[c];
            const int i = ci[c];
            if (key_less(prev_d, prev_i, d, i) && key_less(d, i, md, mi)) { md = d; mi = i; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, md, off);
            const int oi = __shfl_xor_sync(0xffffffffu, mi, off);
            if (key_less(od, oi, md, mi)) { md = od; mi = oi; }
        }
        if (lane == 0) { out_d[r] = md; out_i[r] = mi; }
        prev_d = md;
        prev_i = mi;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32 * NR_WARPS)
nearest4_rect_kernel(const double *__restrict__ sl_y, const double *__restrict__ cl_x,
                     const double *__restrict__ lat_y, const double *__restrict__ lon_x, int ny, int nx,
                     const double *__restrict__ ob_sl, const double *__restrict__ ob_cl,
                     const double *__restrict__ ob_lat, const double *__restrict__ ob_lon, int64_t nobs,
                     int64_t *__restrict__ idx4, double *__restrict__ w4, int32_t *__restrict__ n_exact) {
    __shared__ double s_d[NR_WARPS][32 * NR_K];
    __shared__ int s_i[NR_WARPS][32 * NR_K];
    __shared__ double s_ad[NR_WARPS][NR_K], s_bd[NR_WARPS][NR_K], s_fd[NR_WARPS][4];
    __shared__ int s_ai[NR_WARPS][NR_K], s_bi[NR_WARPS][NR_K], s_fi[NR_WARPS][4];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int64_t k = blockIdx.x * (int64_t)NR_WARPS + warp;
    if (k >= nobs) return;
    const double osl = ob_sl[k], ocl = ob_cl[k];
    double d[NR_K];
    int ix[NR_K];
    // 8 smallest A[y] (ties: lowest y)
#pragma unroll
    for (int j = 0; j < NR_K; ++j) { d[j] = INFINITY; ix[j] = 0x7fffffff; }
    for (int y = lane; y < ny; y += 32) {
        const double a = sl_y[y] - osl;
        topk_insert<NR_K>(d, ix, __dmul_rn(a, a), y);
    }
#pragma unroll
    for (int j = 0; j < NR_K; ++j) { s_d[warp][lane * NR_K + j] = d[j]; s_i[warp][lane * NR_K + j] = ix[j]; }
    __syncwarp();
    warp_select(s_d[warp], s_i[warp], 32 * NR_K, NR_K, lane, s_ad[warp], s_ai[warp]);
    // 8 smallest B[x] (ties: lowest x)
#pragma unroll
    for (int j = 0; j < NR_K; ++j) { d[j] = INFINITY; ix[j] = 0x7fffffff; }
    for (int x = lane; x < nx; x += 32) {
        const double b = cl_x[x] - ocl;
        topk_insert<NR_K>(d, ix, __dmul_rn(b, b), x);
    }
#pragma unroll
    for (int j = 0; j < NR_K; ++j) { s_d[warp][lane * NR_K + j] = d[j]; s_i[warp][lane * NR_K + j] = ix[j]; }
    __syncwarp();
    warp_select(s_d[warp], s_i[warp], 32 * NR_K, NR_K, lane, s_bd[warp], s_bi[warp]);
    // 64 combinations -> 4 smallest by (d2, flat index)
    for (int cidx = lane; cidx < NR_K * NR_K; cidx += 32) {
        const int i = cidx / NR_K, j = cidx % NR_K;
        const int yy = s_ai[warp][i], xx = s_bi[warp][j];
        const bool ok = yy < ny && xx < nx;
        s_d[warp][cidx] = ok ? __dadd_rn(s_ad[warp][i], s_bd[warp][j]) : INFINITY;
        s_i[warp][cidx] = ok ? yy * nx + xx : 0x7fffffff;
    }
    __syncwarp();
    warp_select(s_d[warp], s_i[warp], NR_K * NR_K, 4, lane, s_fd[warp], s_fi[warp]);
    if (lane == 0) {
        double dist[4], w[4];
        int amin = 0;
        bool exact = false;
        for (int r = 0; r < 4; ++r) {
            const int f = s_fi[warp][r];
            dist[r] = haversine_ref(lat_y[f / nx], lon_x[f % nx], ob_lat[k], ob_lon[k]);
            if (dist[r] < 1.0) exact = true;
            if (dist[r] < dist[amin]) amin = r;
        }
        if (exact) {
            for (int r = 0; r < 4; ++r) w[r] = (r == amin) ? 1.0 : 0.0;
            if (n_exact) atomicAdd(n_exact, 1);
        } else {
            double s = 0.0;
            for (int r = 0; r < 4; ++r) { w[r] = 1.0 / dist[r]; s += w[r]; }
            for (int r = 0; r < 4; ++r) w[r] /= s;
        }
        for (int r = 0; r < 4; ++r) { idx4[k * 4 + r] = s_fi[warp][r]; w4[k * 4 + r] = w[r]; }
    }
}

// ------------------------------------------------------------------------------------------
// K-point weighted gather: Y[k][m] = sum_p w[k][p] X[idx[k][p]][m]   (state/ensemble.py:226-237)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void gather_kernel(const T *__restrict__ X, int nens, const int64_t *__restrict__ idx,
                              const double *__restrict__ w, int K, int64_t nobs, T *__restrict__ Y) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.y + threadIdx.y;
    if (k >= nobs) return;
    for (int m = threadIdx.x; m < nens; m += blockDim.x) {
        double acc = 0.0;
        for (int p = 0; p < K; ++p) {
            const double wp = w[k * K + p];
            if (wp != 0.0) acc += wp * (double)X[idx[k * K + p] * nens + m];
        }
        Y[k * nens + m] = (T)acc;
    }
}

// 8-point stencil = 4 space points x 2 time levels, weights multiplied (state/ensemble.py:226-237); with a latitude
// band [y_begin, y_end) the row indices are re-based to a shard that holds only those grid rows of every level, and
// stencil points outside the band get weight 0 (their partial sums belong to other ranks).
__global__ void stencil8_kernel(const int64_t *__restrict__ idx4, const double *__restrict__ w4,
                                const int64_t *__restrict__ row0, const int64_t *__restrict__ row1,
                                const double *__restrict__ tw0, const double *__restrict__ tw1, int64_t nobs,
                                int64_t ny, int64_t nx, int64_t y_begin, int64_t y_end, int diag,
                                int64_t *__restrict__ idx8, double *__restrict__ w8) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nobs) return;
    const int64_t npts = ny * nx, nyl = y_end - y_begin;
    for (int h = 0; h < 2; ++h) {
        const int64_t lev = (h ? row1[k] : row0[k]) / npts;
        const double tw = h ? tw1[k] : tw0[k];
        for (int p = 0; p < 4; ++p) {
            // diag: idx4 holds indices n into a 1-D point list and the stencil point is (y, x) = (n, n), which is how the
            // reference's 1-D lat/lon branch reads the state (state/ensemble.py:185-187, :226)
            const int64_t pt = idx4[k * 4 + p], y = diag ? pt : pt / nx, x = diag ? pt : pt - y * nx;
            const bool inside = y >= y_begin && y < y_end;
            idx8[k * 8 + 4 * h + p] = inside ? (lev * nyl + (y - y_begin)) * nx + x : 0;
            w8[k * 8 + 4 * h + p] = inside ? tw * w4[k * 4 + p] : 0.0;
        }
    }
}

extern "C" int exb_stencil_combine(const int64_t *idx4, const double *w4, const int64_t *row0, const int64_t *row1,
                                   const double *tw0, const double *tw1, int64_t nobs, int64_t ny, int64_t nx,
                                   int64_t y_begin, int64_t y_end, int diag, int64_t *idx8, double *w8, void *stream) {
    EXB_REQUIRE(idx4 && w4 && row0 && row1 && tw0 && tw1 && idx8 && w8, "null pointer");
    EXB_REQUIRE(nobs > 0 && ny > 0 && nx > 0 && y_begin >= 0 && y_end <= ny && y_begin < y_end, "bad sizes");
    stencil8_kernel<<<(unsigned)ceil_div64(nobs, 256), 256, 0, (cudaStream_t)stream>>>(idx4, w4, row0, row1, tw0, tw1, nobs,
                                                                                         ny, nx, y_begin, y_end, diag, idx8, w8);
    exb_count_launches(1);
    return exb_check_launch("stencil8_kernel");
}

// ------------------------------------------------------------------------------------------
// row-wise mean / perturbation split, inflation, recombination: one warp per row
// ------------------------------------------------------------------------------------------
template <typename T, int MODE>   // 0 split, 1 inflate, 2 recombine
__global__ void row_kernel(T *__restrict__ X, T *__restrict__ xm, int64_t nrows, int nens,
                           const double *__restrict__ factor, int64_t rows_per_factor) {
    const int lane = threadIdx.x % 32;
    const int64_t warps_per_grid = (int64_t)gridDim.x * (blockDim.x / 32);
    for (int64_t r = blockIdx.x * (int64_t)(blockDim.x / 32) + threadIdx.x / 32; r < nrows; r += warps_per_grid) {
        T *row = X + r * nens;
        if (MODE == 2) {
            const T m = xm[r];
            for (int i = lane; i < nens; i += 32) row[i] += m;
            continue;
        }
        double s = 0.0;
        for (int i = lane; i < nens; i += 32) s += (double)row[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        const double mean = s / (double)nens;
        if (MODE == 0) {
            const T mt = (T)mean;
            if (lane == 0) xm[r] = mt;
            for (int i = lane; i < nens; i += 32) row[i] -= mt;
        } else {
            const double f = factor[r / rows_per_factor];
            for (int i = lane; i < nens; i += 32) row[i] = (T)(((double)row[i] - mean) * f + mean);
        }
    }
}

// One observation against n points: Observation.localize (observation/observation.py:59-87)
__global__ void localization_weights_kernel(const double *__restrict__ u, int64_t n, double ox, double oy, double oz,
                                            double inv_hw, double a_max, int loc_mode,
                                            double *__restrict__ dist, double *__restrict__ wout) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = hav_a(u[i], u[n + i], u[2 * n + i], ox, oy, oz);
    if (dist) dist[i] = EXB_R_EARTH * angle_from_a(a);
    if (wout) wout[i] = (loc_mode == EXB_LOC_GC) ? loc_weight(a, inv_hw, a_max) : 1.0;
}

__global__ void gaspari_cohn_kernel(const double *__restrict__ d, int64_t n, double abs_hw, double *__restrict__ w) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    w[i] = gaspari_cohn_r(d[i] / abs_hw);      // r = distances / abs(halfwidth), observation.py:120
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" int exb_localization_weights(const double *u, int64_t n, double ob_lat, double ob_lon, double hw,
                                        int loc_mode, double *dist, double *weights, void *stream) {
    EXB_REQUIRE(u && n > 0 && (dist || weights), "null pointer or n <= 0");
    const double phi = ob_lat * EXB_DEG2RAD, lam = ob_lon * EXB_DEG2RAD;
    double inv_hw = 0.0, a_max = 2.0;
    if (loc_mode == EXB_LOC_GC) {
        const double h = fabs(hw);
        if (h == 0.0) { a_max = 0.0; }
        else {
            inv_hw = 1.0 / h;
            const double half = h / EXB_R_EARTH;
            if (half < 0.5 * M_PI) a_max = sin(half) * sin(half);
        }
    }
    localization_weights_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(
        u, n, cos(phi) * cos(lam), cos(phi) * sin(lam), sin(phi), inv_hw, a_max, loc_mode, dist, weights);
    exb_count_launches(1);
    return exb_check_launch("localization_weights_kernel");
}

extern "C" int exb_gaspari_cohn(const double *d, int64_t n, double hw, double *w, void *stream) {
    EXB_REQUIRE(d && w && n > 0, "null pointer or n <= 0");
    gaspari_cohn_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(d, n, fabs(hw), w);
    exb_count_launches(1);
    return exb_check_launch("gaspari_cohn_kernel");
}
extern "C" int exb_grid_unitvec(const double *lat, const double *lon, int64_t npts, double *grid_u, void *stream) {
    EXB_REQUIRE(lat && lon && grid_u && npts > 0, "null pointer or npts <= 0");
    grid_unitvec_kernel<<<(unsigned)ceil_div64(npts, 256), 256, 0, (cudaStream_t)stream>>>(lat, lon, npts, grid_u);
    exb_count_launches(1);
    return exb_check_launch("grid_unitvec_kernel");
}

extern "C" int exb_obs_prepare(const double *lat, const double *lon, const double *hw, int64_t nobs,
                               int loc_mode, double *obgeo, void *stream) {
    EXB_REQUIRE(lat && lon && obgeo && nobs > 0, "null pointer or nobs <= 0");
    EXB_REQUIRE(loc_mode == EXB_LOC_NONE || (loc_mode == EXB_LOC_GC && hw), "loc_mode GC needs halfwidths");
    obs_prepare_kernel<<<(unsigned)ceil_div64(nobs, 256), 256, 0, (cudaStream_t)stream>>>(lat, lon, hw, nobs, loc_mode, obgeo);
    exb_count_launches(1);
    return exb_check_launch("obs_prepare_kernel");
}

extern "C" int exb_stencil_search(const double *sinlat_g, const double *coslon_g, const double *lat_g,
                                  const double *lon_g, int64_t npts, const double *ob_sinlat,
                                  const double *ob_coslon, const double *ob_lat, const double *ob_lon,
                                  int64_t nobs, int64_t *idx4, double *w4, int32_t *n_exact, void *stream) {
    EXB_REQUIRE(sinlat_g && coslon_g && lat_g && lon_g && ob_sinlat && ob_coslon && ob_lat && ob_lon && idx4 && w4,
                "null pointer");
    EXB_REQUIRE(npts >= 4 && npts < 0x7fffffff && nobs > 0, "need 4 <= npts < 2^31 and nobs > 0");
    const size_t smem = (size_t)NS_OBS_PER_CTA * NS_TX * 4 * (sizeof(double) + sizeof(int));
    static bool attr_set = false;
    if (!attr_set) {
        EXB_CUDA(cudaFuncSetAttribute(nearest4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    if (n_exact) EXB_CUDA(cudaMemsetAsync(n_exact, 0, sizeof(int32_t), (cudaStream_t)stream));
    nearest4_kernel<<<(unsigned)ceil_div64(nobs, NS_OBS_PER_CTA), NS_TX * NS_OY, smem, (cudaStream_t)stream>>>(
        sinlat_g, coslon_g, lat_g, lon_g, (int)npts, ob_sinlat, ob_coslon, ob_lat, ob_lon, nobs, idx4, w4, n_exact);
    exb_count_launches(1);
    return exb_check_launch("nearest4_kernel");
}

extern "C" int exb_stencil_search_rect(const double *sinlat_y, const double *coslon_x, const double *lat_y,
                                       const double *lon_x, int64_t ny, int64_t nx, const double *ob_sinlat,
                                       const double *ob_coslon, const double *ob_lat, const double *ob_lon,
                                       int64_t nobs, int64_t *idx4, double *w4, int32_t *n_exact, void *stream) {
    EXB_REQUIRE(sinlat_y && coslon_x && lat_y && lon_x && ob_sinlat && ob_coslon && ob_lat && ob_lon && idx4 && w4,
                "null pointer");
    EXB_REQUIRE(ny > 0 && nx > 0 && ny * nx >= 4 && ny * nx < 0x7fffffff && nobs > 0,
                "need 4 <= ny*nx < 2^31 and nobs > 0");
    if (n_exact) EXB_CUDA(cudaMemsetAsync(n_exact, 0, sizeof(int32_t), (cudaStream_t)stream));
    nearest4_rect_kernel<<<(unsigned)ceil_div64(nobs, NR_WARPS), 32 * NR_WARPS, 0, (cudaStream_t)stream>>>(
        sinlat_y, coslon_x, lat_y, lon_x, (int)ny, (int)nx, ob_sinlat, ob_coslon, ob_lat, ob_lon, nobs, idx4, w4,
        n_exact);
    exb_count_launches(1);
    return exb_check_launch("nearest4_rect_kernel");
}

template <typename T>
static int gather_impl(const T *X, int64_t nrows, int nens, const int64_t *idx, const double *w, int K,
                       int64_t nobs, T *Y, void *stream) {
    EXB_REQUIRE(X && idx && w && Y, "null pointer");
    EXB_REQUIRE(nrows > 0 && nens > 0 && nobs > 0 && K > 0 && K <= 8, "bad sizes (need 0 < K <= 8)");
    dim3 block(32, 8);
    gather_kernel<T><<<(unsigned)ceil_div64(nobs, 8), block, 0, (cudaStream_t)stream>>>(X, nens, idx, w, K, nobs, Y);
    exb_count_launches(1);
    return exb_check_launch("gather_kernel");
}
extern "C" int exb_gather_f64(const double *X, int64_t nrows, int nens, const int64_t *idx, const double *w,
                              int K, int64_t nobs, double *Y, void *stream) {
    return gather_impl<double>(X, nrows, nens, idx, w, K, nobs, Y, stream);
}
extern "C" int exb_gather_f32(const float *X, int64_t nrows, int nens, const int64_t *idx, const double *w,
                              int K, int64_t nobs, float *Y, void *stream) {
    return gather_impl<float>(X, nrows, nens, idx, w, K, nobs, Y, stream);
}

template <typename T, int MODE>
static int row_impl(T *X, T *xm, int64_t nrows, int nens, const double *factor_dev, int64_t rows_per_factor,
                    void *stream, const char *what) {
    EXB_REQUIRE(X && nrows > 0 && nens > 0, "null pointer or empty");
    const int64_t blocks = ceil_div64(nrows, 8);
    const unsigned grid = (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);
    row_kernel<T, MODE><<<grid, 256, 0, (cudaStream_t)stream>>>(X, xm, nrows, nens, factor_dev, rows_per_factor);
    exb_count_launches(1);
    return exb_check_launch(what);
}
extern "C" int exb_split_mean_pert_f64(double *X, double *xm, int64_t nrows, int nens, void *stream) {
    EXB_REQUIRE(xm, "null xm");
    return row_impl<double, 0>(X, xm, nrows, nens, nullptr, 1, stream, "split_mean_pert");
}
extern "C" int exb_split_mean_pert_f32(float *X, float *xm, int64_t nrows, int nens, void *stream) {
    EXB_REQUIRE(xm, "null xm");
    return row_impl<float, 0>(X, xm, nrows, nens, nullptr, 1, stream, "split_mean_pert");
}
extern "C" int exb_recombine_f64(double *X, const double *xm, int64_t nrows, int nens, void *stream) {
    EXB_REQUIRE(xm, "null xm");
    return row_impl<double, 2>(X, const_cast<double *>(xm), nrows, nens, nullptr, 1, stream, "recombine");
}
extern "C" int exb_recombine_f32(float *X, const float *xm, int64_t nrows, int nens, void *stream) {
    EXB_REQUIRE(xm, "null xm");
    return row_impl<float, 2>(X, const_cast<float *>(xm), nrows, nens, nullptr, 1, stream, "recombine");
}

template <typename T>
static int inflate_impl(T *X, int64_t nrows, int nens, const double *factor_host, int64_t nfactor,
                        int64_t rows_per_factor, void *stream) {
    EXB_REQUIRE(X && factor_host && nfactor > 0 && rows_per_factor > 0, "null pointer or bad factor layout");
    EXB_REQUIRE(nfactor * rows_per_factor >= nrows, "factors do not cover all rows");
    double *fdev = nullptr;
    EXB_CUDA(exb_malloc_async(&fdev, nfactor * sizeof(double), (cudaStream_t)stream));
    EXB_CUDA(cudaMemcpyAsync(fdev, factor_host, nfactor * sizeof(double), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    int rc = row_impl<T, 1>(X, nullptr, nrows, nens, fdev, rows_per_factor, stream, "inflate");
    EXB_CUDA(cudaFreeAsync(fdev, (cudaStream_t)stream));
    return rc;
}
extern "C" int exb_inflate_f64(double *X, int64_t nrows, int nens, const double *f, int64_t nf, int64_t rpf, void *stream) {
    return inflate_impl<double>(X, nrows, nens, f, nf, rpf, stream);
}
extern "C" int exb_inflate_f32(float *X, int64_t nrows, int nens, const double *f, int64_t nf, int64_t rpf, void *stream) {
    return inflate_impl<float>(X, nrows, nens, f, nf, rpf, stream);
}
