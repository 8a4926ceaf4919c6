// State sweep, warp-specialised FP64 tensor-core version (production path for float64 states).
//
// Same mathematics as state_update_mma.cu (read its header first: blocked 8-ob form of ensrf.py:95-141 on
// mma.sync.m8n8k4.f64, mean carried as a pseudo-member column), different organisation.  That kernel kept
// the FP64 tensor pipe busy only half of the time: every round of 64 candidates was scanned, staged and
// weighted by the same warps that then issue the DMMAs, behind CTA-wide barriers.  Here
//   * one CTA per SM = 12 CONSUMER warps (8 state rows each = 96 rows of one patch, in registers for the
//     whole launch) + 4 PRODUCER warps; the producers are the warps of SM sub-partition 3, so that their scalar
//     FP64 arithmetic does not queue behind the consumers' DMMAs (see the role mapping in the kernel);
//   * producers find the patch's candidate obs in serial order (fp32 cap test over the candidate list of the
//     coarse tile that contains the patch), and per batch of 8 obs stage into a ring of shared-memory stages:
//     the 8 ye rows (cp.async), the pseudo-member column -innov/beta, omega[grid point][ob] = beta c1 GC(d),
//     and the 8x8 Gram matrix (DMMA); batches are dealt round-robin to the 4 producer warps;
//   * consumers only wait on the stage's `full` mbarrier, run  g = X Y^T -> 8-step recurrence -> X -= E Y, and
//     arrive on its `empty` mbarrier.  No CTA-wide barrier after start-up.
// Optionally fused with the mean/perturbation split and the recombination (assimilation.py:146-147, :168): with
// xm == nullptr the kernel reads full ensemble values, forms mean and perturbations in registers and writes
// mean + perturbation back, so the state crosses HBM exactly once (one read, one write of the touched rows).
//
// Candidate lists: a pre-pass (sweep_tile_*) builds, per coarse tile of 4x4 patches, the ascending list of obs
// whose support cap can touch the tile; the producers' per-patch test then walks a few thousand entries instead
// of every ob.  Without localisation (or when lists would not pay) the producers walk the ob range itself.
#include "common.cuh"
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef SP_PW
#define SP_PW 4                       // producer warps (on SM sub-partition 3: warps 3, 7, ...)
#endif
#define SP_CW (16 - SP_PW)            // consumer warps
#define SP_NT ((SP_CW + SP_PW) * 32)
#define SP_ROWS (SP_CW * 8)           // state rows per CTA
#define SP_MAXSTAGES 16

struct SpParams {
    void *xm;                         // nullptr: fused mean/perturbation split + recombination   (storage type TS)
    void *Xp;                         // state rows, storage type TS (double or float); arithmetic is always double
    const void *Yp;                   // ye rows, storage type TS
    const double *grid_u;
    const double *rec;
    const double *geo;
    const float4 *scan;               // (ux, uy, uz, theta) per ob; theta < 0: never a candidate
    const int64_t *tile_off;          // candidate lists per coarse tile (nullptr: walk the ob range)
    const int *tile_list;
    unsigned long long *counters;
    int64_t npts, nobs, ob_begin, ob_end;
    int nlev, ny, nx, nens;
    int y_begin, y_end;               // grid rows [y_begin, y_end) of the shard are swept by this launch
    int eq_row;                       // grid row of the shard closest to the equator (-1: unknown, keep row order)
    int pr0, npr, pr_eq;              // first patch row / number of patch rows of this launch, patch row of the equator
    int ty, tx, ntx, nctx;            // patch shape, patches along x, coarse tiles along x
    int G, Lc, nlc;
    int loc_mode;
    int nstages, stage_doubles;       // ring geometry
    int role_split;
};

__device__ __forceinline__ void sp_dmma(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// volatile variant: keeps the program order of a sequence of DMMAs (the compiler may not re-pair them)
__device__ __forceinline__ void sp_dmma_v(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned sp_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sp_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(sp_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sp_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(sp_smem(bar)) : "memory");
}
// (two textually separate copies so that profiler samples of consumers waiting for data and of producers waiting
// for a free stage land on different source lines)
__device__ __forceinline__ void sp_mbar_wait_full(unsigned long long *bar, unsigned parity) {
    unsigned ok = 0;
    const unsigned a = sp_smem(bar);
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void sp_mbar_wait_empty(unsigned long long *bar, unsigned parity) {
    unsigned ok = 0;
    const unsigned a = sp_smem(bar);
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}

template <int NT3> __host__ __device__ constexpr int sp_yst() { return ((8 * NT3) % 16 == 8) ? 8 * NT3 : 8 * NT3 + 8; }

// Stage layout (doubles): y[8][YST] | om[G][8] | Gram[64] | ob[6][8] | count (one double slot, int inside)
// With MG the stage also carries, per grid point, the NEGATED 8x8 lower-triangular matrix M of the batch recurrence
// (e = M g, see the consumer): | mneg[G][64]
template <int NT3> __host__ __device__ constexpr int sp_stage_doubles(int G, bool mg) {
    return 8 * sp_yst<NT3>() + 8 * G + 64 + 48 + 2 + (mg ? 64 * G : 0);
}

template <int NT3, typename TS, bool MG>
__global__ void __launch_bounds__(SP_NT, 1) state_sweep_pipe_kernel(const SpParams p) {
    TS *const gXp = static_cast<TS *>(p.Xp);
    TS *const gxm = static_cast<TS *>(p.xm);
    const TS *const gYp = static_cast<const TS *>(p.Yp);
    constexpr int YST = sp_yst<NT3>();
    constexpr int PC = 8 * NT3 - 1;          // column of the pseudo-member (the mean)

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_ring = reinterpret_cast<double *>(smem_raw);                                   // [nstages][stage_doubles]
    double *s_gu = s_ring + (size_t)p.nstages * p.stage_doubles;                             // [3][SP_ROWS]
    unsigned long long *s_full = reinterpret_cast<unsigned long long *>(s_gu + 3 * SP_ROWS); // [SP_MAXSTAGES]
    unsigned long long *s_empty = s_full + SP_MAXSTAGES;                                      // [SP_MAXSTAGES]
    int *s_gvalid = reinterpret_cast<int *>(s_empty + SP_MAXSTAGES);                          // [SP_ROWS]
    int *s_mine = s_gvalid + SP_ROWS;                                                         // [SP_PW][4][8]
    int *s_scanbuf = s_mine + SP_PW * 32;                                                     // [SP_PW][128]
    __shared__ float s_bound[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = lane & 3, n = lane >> 2;
    const int G = p.G, Lc = p.Lc, nens = p.nens;
    const int S = p.nstages, SD = p.stage_doubles;

    const int lc = blockIdx.x % p.nlc;
    const int tile = blockIdx.x / p.nlc;
    // Patch rows are issued heaviest first: on a lat-lon grid the rows next to the poles meet the most candidate
    // obs (their grid points are closer together than the obs), so they must not be the tail of the launch.  Rows are
    // taken alternately from the two ends of the launch's range when it straddles the equator, else from the
    // poleward end.
    int prow;
    {
        const int r = tile / p.ntx, nr = p.npr, eq = p.pr_eq;
        if (eq < 0) prow = p.pr0 + r;
        else if (eq <= p.pr0) prow = p.pr0 + nr - 1 - r;
        else if (eq >= p.pr0 + nr) prow = p.pr0 + r;
        else prow = (r & 1) ? p.pr0 + nr - 1 - (r >> 1) : p.pr0 + (r >> 1);
    }
    const int pcol = tile % p.ntx;
    const int y0 = prow * p.ty, x0 = pcol * p.tx;
    const int l0 = lc * Lc;

    // ---- start-up (the only CTA-wide barriers) ----------------------------------------------------
    if (tid < G) {
        const int gy = y0 + tid / p.tx, gx = x0 + tid % p.tx;
        const bool ok = gy >= p.y_begin && gy < p.y_end && gx < p.nx;
        const int cy = min(max(gy, p.y_begin), p.y_end - 1), cx = min(gx, p.nx - 1);
        const int64_t pt = (int64_t)cy * p.nx + cx;
        s_gu[tid] = p.grid_u[pt];
        s_gu[SP_ROWS + tid] = p.grid_u[p.npts + pt];
        s_gu[2 * SP_ROWS + tid] = p.grid_u[2 * p.npts + pt];
        s_gvalid[tid] = ok;
    }
    for (int i = tid; i < S * SD; i += SP_NT) s_ring[i] = 0.0;           // padding columns stay zero for good
    if (tid < S) {
        sp_mbar_init(s_full + tid, 1);
        sp_mbar_init(s_empty + tid, SP_CW);
    }
    __syncthreads();
    if (tid == 0) {
        double cx = 0, cy = 0, cz = 0;
        for (int g = 0; g < G; ++g) { cx += s_gu[g]; cy += s_gu[SP_ROWS + g]; cz += s_gu[2 * SP_ROWS + g]; }
        const double nn = sqrt(cx * cx + cy * cy + cz * cz);
        if (nn > 1e-12) { cx /= nn; cy /= nn; cz /= nn; } else { cx = s_gu[0]; cy = s_gu[SP_ROWS]; cz = s_gu[2 * SP_ROWS]; }
        double cmin = 1.0;
        for (int g = 0; g < G; ++g) cmin = fmin(cmin, cx * s_gu[g] + cy * s_gu[SP_ROWS + g] + cz * s_gu[2 * SP_ROWS + g]);
        s_bound[0] = (float)cx; s_bound[1] = (float)cy; s_bound[2] = (float)cz;
        s_bound[3] = (float)(acos(fmax(-1.0, fmin(1.0, cmin))) + 1e-6);
    }
    __syncthreads();

    // Roles by scheduler: warp w runs on SM sub-partition w % 4.  The FP64 tensor pipe and the scalar FP64 pipe
    // share one issue port per sub-partition, arbitrated per instruction: a producer's 2-clock DFMA queued behind
    // three consumers' 16-clock DMMAs costs ~20-50 clocks.  With p.role_split the producers own sub-partition 3
    // (their localisation arithmetic then runs at the full scalar rate) and the consumers share the other three.
    // producers: the first SP_PW warps of sub-partition 3 (warps 3, 7, 11, 15); consumers: all others, numbered
    // in warp order
    const bool is_producer = p.role_split ? ((warp & 3) == 3 && (warp >> 2) < SP_PW) : (warp >= SP_CW);
    int cw = warp;                        // consumer index 0..SP_CW-1
    if (p.role_split) {
        int before = (warp + 1) >> 2;     // warps of sub-partition 3 with a lower index
        if (before > SP_PW) before = SP_PW;
        cw = warp - before;
    }
    if (is_producer) {
        // =============================== PRODUCER ===============================
        const int pw = p.role_split ? (warp >> 2) : (warp - SP_CW);
        const float bcx = s_bound[0], bcy = s_bound[1], bcz = s_bound[2], brho = s_bound[3];
        int *mine = s_mine + pw * 32;
        unsigned long long npairs = 0;
        const unsigned lt = (1u << lane) - 1u;

        // A batch is staged in two halves so that two batches are in flight per producer warp and no global-memory
        // latency is exposed: issue(b) waits for the stage to be free, starts the loads of the per-ob scalars (they
        // stay in registers) and the cp.async copies of the 8 ye rows; finish(b) -- called after the NEXT batch of
        // this warp has been issued -- writes the scalars, evaluates omega, waits for the rows, forms the Gram
        // matrix and publishes the stage.  nq == 0 publishes the end marker.
        struct Scal { double ux, uy, uz, ihw, amax, c1, beta, innov; };
        auto issue = [&](int b, int nq, const int *cand, Scal &sc) {
            const int s = b % S, u = b / S;
            sp_mbar_wait_empty(s_empty + s, (unsigned)((u & 1) ^ 1));
            if (nq == 0) return;
            double *sy = s_ring + (size_t)s * SD;
            sc.ux = sc.uy = sc.uz = sc.ihw = sc.amax = sc.c1 = sc.innov = 0.0;
            sc.beta = 1.0;
            if (lane < nq) {
                const int64_t kk = cand[lane];
                sc.ux = __ldg(p.geo + GEO_UX * p.nobs + kk); sc.uy = __ldg(p.geo + GEO_UY * p.nobs + kk);
                sc.uz = __ldg(p.geo + GEO_UZ * p.nobs + kk); sc.ihw = __ldg(p.geo + GEO_INVHW * p.nobs + kk);
                sc.amax = __ldg(p.geo + GEO_AMAX * p.nobs + kk);
                sc.c1 = __ldg(p.rec + REC_C1 * p.nobs + kk); sc.beta = __ldg(p.rec + REC_BETA * p.nobs + kk);
                sc.innov = __ldg(p.rec + REC_INNOV * p.nobs + kk);
            }
            if (sizeof(TS) == 8) {
#pragma unroll 1
                for (int q = 0; q < 8; ++q) {
                    double *dst = sy + q * YST;
                    const int sw = ((q >> 1) & 1) << 2;
                    if (q < nq) {
                        const double *src8 = reinterpret_cast<const double *>(gYp) + (int64_t)cand[q] * nens;
                        if ((nens & 1) == 0) {
                            for (int m = 2 * lane; m < nens; m += 64)
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sp_smem(dst + (m ^ sw))), "l"(src8 + m));
                        } else {
                            for (int m = lane; m < nens; m += 32)
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sp_smem(dst + (m ^ sw))), "l"(src8 + m));
                        }
                    } else {
                        for (int m = lane; m < nens; m += 32) dst[m ^ sw] = 0.0;
                    }
                }
            } else {
                // float32 storage: widen on the way into shared memory; all loads of the batch are issued before
                // the first value is needed (8 rows x at most 4 values per lane: nens <= 8*NT3 - 1 <= 127)
                constexpr int NCH = (8 * NT3 + 31) / 32;
                float v[8][NCH];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const TS *src = gYp + (int64_t)cand[q < nq ? q : 0] * nens;
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        const int m = lane + 32 * j;
                        v[q][j] = (q < nq && m < nens) ? (float)__ldg(src + m) : 0.f;
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    double *dst = sy + q * YST;
                    const int sw = ((q >> 1) & 1) << 2;
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        const int m = lane + 32 * j;
                        if (m < nens) dst[m ^ sw] = (double)v[q][j];
                    }
                }
            }
            asm volatile("cp.async.commit_group;\n" ::);
        };
        auto finish = [&](int b, int nq, const Scal &sc, bool newer_group_pending) {
            const int s = b % S;
            double *sy = s_ring + (size_t)s * SD;
            double *som = sy + 8 * YST;
            double *sG = som + 8 * G;
            double *sob = sG + 64;
            int *scnt = reinterpret_cast<int *>(sob + 48);
            if (nq > 0) {
                if (lane < 8) {
                    sob[0 * 8 + lane] = sc.ux; sob[1 * 8 + lane] = sc.uy; sob[2 * 8 + lane] = sc.uz;
                    sob[3 * 8 + lane] = sc.ihw; sob[4 * 8 + lane] = sc.amax;
                    // beta / ((N-1) kdenom)   (ensrf.py:95, :119, :135-136); the localisation weight multiplies it below
                    sob[5 * 8 + lane] = sc.c1 * sc.beta;
                    // pseudo-member column: the rank-8 update then also performs xam = xbm + kmat*innov (ensrf.py:130)
                    sy[lane * YST + (PC ^ (((lane >> 1) & 1) << 2))] = (lane < nq) ? -sc.innov / sc.beta : 0.0;
                }
                __syncwarp();
                // supports of all 8 obs within the range of the branch-free weight function?
                const bool fast = __all_sync(0xffffffffu, lane >= 8 || sc.amax <= EXB_FAST_AMAX);
                // omega[g][q] = beta * loc / ((N-1) kdenom): lane-parallel over (grid point, ob) pairs, four
                // independent evaluations in flight per lane
                for (int i0 = lane; i0 < 8 * G; i0 += 128) {
                    double omv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = i0 + 32 * j;
                        const int q = i & 7, gg = min(i >> 3, G - 1);
                        double w = 0.0;
                        if (i < 8 * G && q < nq && s_gvalid[gg]) {
                            w = 1.0;
                            if (p.loc_mode == EXB_LOC_GC) {
                                const double a = hav_a(s_gu[gg], s_gu[SP_ROWS + gg], s_gu[2 * SP_ROWS + gg],
                                                       sob[0 * 8 + q], sob[1 * 8 + q], sob[2 * 8 + q]);
                                w = fast ? loc_weight_fast(a, sob[3 * 8 + q], sob[4 * 8 + q])
                                         : loc_weight(a, sob[3 * 8 + q], sob[4 * 8 + q]);
                            }
                            npairs += (w != 0.0 && lc == 0) ? 1 : 0;
                        }
                        omv[j] = w * sob[5 * 8 + q];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (i0 + 32 * j < 8 * G) som[i0 + 32 * j] = omv[j];
                }
                if (newer_group_pending) asm volatile("cp.async.wait_group 1;\n" ::);
                else asm volatile("cp.async.wait_group 0;\n" ::);
                __syncwarp();
                // Gram matrix of the batch (members only: the pseudo-member column is masked)
                {
                    double g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
                    const double *yrow = sy + n * YST;
                    const int sw = ((n >> 1) & 1) << 2;
#pragma unroll
                    for (int t = 0; t < NT3; ++t) {
                        const double2 v = *reinterpret_cast<const double2 *>(yrow + 8 * t + ((2 * c) ^ sw));
                        const double v1 = (t == NT3 - 1 && c == 3) ? 0.0 : v.y;
                        sp_dmma(g0, g1, v.x, v.x);
                        sp_dmma(h0, h1, v1, v1);
                    }
                    sG[n * 8 + 2 * c] = g0 + h0;
                    sG[n * 8 + 2 * c + 1] = g1 + h1;
                }
                if (MG) {
                    // The 8-step recurrence  e_q = omega_q (g_q - sum_{p<q} G_qp e_p)  is linear in g: e = M g with M
                    // lower triangular, M_jj = omega_j, M_qj = -omega_q sum_{j<=p<q} G_qp M_pj.  M depends on the grid
                    // point (through omega) and on the batch (through G); computing it here, where scalar FP64 is
                    // cheap, leaves the consumers two short dot products per lane instead of the serial chain.
                    // One (grid point, column j) task at a time per lane; stored negated (the update subtracts).
                    __syncwarp();
                    double *sM = sob + 50;
                    for (int t = lane; t < 8 * G; t += 32) {
                        // column j of M for grid point gg = the recurrence applied to the unit vector e_j (uniform
                        // code for every lane; entries above the diagonal come out as exact zeros)
                        const int j = t & 7, gg = t >> 3;
                        const double *om8 = som + gg * 8;
                        double m[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            double acc = (q == j) ? 1.0 : 0.0;
#pragma unroll
                            for (int pp = 0; pp < q; ++pp) acc = fma(-sG[q * 8 + pp], m[pp], acc);
                            m[q] = om8[q] * acc;
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) sM[gg * 64 + q * 8 + j] = -m[q];
                    }
                }
            }
            if (lane == 0) *scnt = nq;
            __syncwarp();
            if (lane == 0) sp_mbar_arrive(s_full + s);
        };

        // walk the candidate source in serial order; batch b belongs to producer b % SP_PW
        int64_t lb, le;
        const int *list = nullptr;
        if (p.tile_off) {
            const int ct = (prow / SP_CT) * p.nctx + pcol / SP_CT;
            lb = p.tile_off[ct];
            le = p.tile_off[ct + 1];
            list = p.tile_list;
        } else {
            lb = p.ob_begin;
            le = p.ob_end;
        }
        int nseen = 0, bcur = pw;
        int pend_b = -1, pend_nq = 0;             // batch issued but not finished
        Scal pend_sc, new_sc;
        int *sbuf = s_scanbuf + pw * 128;         // hits of the current group of 4 list chunks (ob index or -1)
        // One loop, one call site of issue() and of finish() (they inline the localisation arithmetic; several
        // copies would not fit the instruction cache next to the consumer loop).  Per iteration: either one chunk
        // of 32 list entries is compacted, or one of the tail events happens.
        int64_t base = lb;
        int chunk = 4, tail = 0;
        while (true) {
            int emit_b = -1, emit_nq = 0;
            bool do_issue = false;
            if (nseen >= 8 * (bcur + 1)) {
                // a complete batch of mine is waiting (one chunk of 32 entries can complete several)
                emit_b = bcur; emit_nq = 8; do_issue = true;
            } else if (base < le || chunk < 4) {
                if (chunk == 4) {
                    // next group of four chunks: their index loads and record gathers are in flight together
                    int idx[4];
                    float4 rec4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int64_t e = base + 32 * j + lane;
                        idx[j] = -1;
                        if (e < le) idx[j] = list ? __ldg(list + e) : (int)e;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        rec4[j] = make_float4(0.f, 0.f, 0.f, -1.f);
                        if (idx[j] >= p.ob_begin && idx[j] < p.ob_end) rec4[j] = __ldg(p.scan + idx[j]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        bool hit = false;
                        if (rec4[j].w >= 0.f) {
                            const float ang = rec4[j].w + brho;
                            hit = (ang >= 3.1405f) || (rec4[j].x * bcx + rec4[j].y * bcy + rec4[j].z * bcz >= __cosf(ang) - 4e-6f);
                        }
                        sbuf[32 * j + lane] = hit ? idx[j] : -1;
                    }
                    __syncwarp();
                    base += 128;
                    chunk = 0;
                }
                const int v = sbuf[32 * chunk + lane];
                ++chunk;
                const unsigned mask = __ballot_sync(0xffffffffu, v >= 0);
                if (mask) {
                    // up to 32 new candidates = parts of up to 5 batches, of which at most 3 are mine: the 4-deep
                    // buffer keeps them apart (at most one earlier batch of mine is still incomplete)
                    if (v >= 0) {
                        const int seq = nseen + __popc(mask & lt);
                        const int bb = seq >> 3;
                        if ((bb % SP_PW) == pw) mine[((bb / SP_PW) & 3) * 8 + (seq & 7)] = v;
                    }
                    nseen += __popc(mask);
                    __syncwarp();
                }
            } else if (tail == 0) {
                tail = 1;
                if (bcur < ((nseen + 7) >> 3)) { emit_b = bcur; emit_nq = nseen - 8 * bcur; do_issue = true; }   // partial last batch
            } else if (tail == 1) {
                tail = 2;
                const int total = (nseen + 7) >> 3;
                if ((total % SP_PW) == pw) { emit_b = total; emit_nq = 0; do_issue = true; }                       // end marker
            } else if (tail == 2) {
                tail = 3;
                emit_b = -2;                                                                                       // flush the pending batch
            } else {
                break;
            }
            if (emit_b == -1) continue;
            if (do_issue) {
                issue(emit_b, emit_nq, mine + ((emit_b / SP_PW) & 3) * 8, new_sc);
                if (emit_nq == 8) bcur += SP_PW;
                else if (emit_nq > 0) bcur += SP_PW;
            }
            if (pend_b >= 0) finish(pend_b, pend_nq, pend_sc, do_issue && emit_nq > 0);
            pend_b = do_issue ? emit_b : -1;
            pend_nq = emit_nq;
            pend_sc = new_sc;
        }
        if (p.counters) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) npairs += __shfl_xor_sync(0xffffffffu, npairs, o);
            if (lane == 0 && npairs) atomicAdd(&p.counters[1], npairs);
        }
        return;
    }

    // =============================== CONSUMER ===============================
    const int r = cw * 8 + n;                 // row slot in the CTA
    const int g = r / Lc, l = r % Lc;
    bool active = false;
    int64_t row = 0;
    if (g < G && l0 + l < p.nlev) {
        const int gy = y0 + g / p.tx, gx = x0 + g % p.tx;
        if (gy >= p.y_begin && gy < p.y_end && gx < p.nx) {
            active = true;
            row = (int64_t)(l0 + l) * p.npts + (int64_t)gy * p.nx + gx;
        }
    }
    const int gslot = active ? g : 0;
    const bool fused = p.xm == nullptr;
    double x[2 * NT3];
    {
        double sum = 0.0;
#pragma unroll
        for (int t = 0; t < NT3; ++t) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = 8 * t + 2 * c + h;
                double v = 0.0;
                if (active && m < nens) { v = (double)gXp[row * nens + m]; sum += v; }
                x[2 * t + h] = v;
            }
        }
        if (fused) {
            // ensemble mean and perturbations of the row (assimilation.py:146-147)
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const double mean = sum / (double)nens;
#pragma unroll
            for (int t = 0; t < NT3; ++t) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int m = 8 * t + 2 * c + h;
                    if (m < nens) x[2 * t + h] -= mean;
                }
            }
            if (c == 3) x[2 * NT3 - 1] = active ? mean : 0.0;
        } else if (c == 3) {
            x[2 * NT3 - 1] = active ? (double)gxm[row] : 0.0;      // the mean rides along in the last column
        }
    }
    bool dirty = false;

    int s = 0;
    unsigned par = 0;
    for (;; s = (s + 1 == S) ? 0 : s + 1, par ^= (s == 0) ? 1u : 0u) {
        sp_mbar_wait_full(s_full + s, par);
        const double *sy = s_ring + (size_t)s * SD;
        const double *som = sy + 8 * YST;
        const double *Gb = som + 8 * G;
        const int nq = *reinterpret_cast<const int *>(Gb + 64 + 48);
        if (nq == 0) break;

        // omega of this row's grid point for the 8 obs of the batch
        double om[8];
        {
            const double2 *po = reinterpret_cast<const double2 *>(som + gslot * 8);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 v = po[i];
                om[2 * i] = active ? v.x : 0.0;
                om[2 * i + 1] = active ? v.y : 0.0;
            }
        }
        // any weight non-zero?  (integer test of the bit patterns: the FP64 pipe is the contended one)
        long long anyb = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) anyb |= __double_as_longlong(om[i]);
        const bool any = (anyb << 1) != 0;
        if (__any_sync(0xffffffffu, any)) {
            // step 1: g[row][ob] = x[row] . y_ob  (two accumulator chains)
            double ga0 = 0.0, ga1 = 0.0, gb0 = 0.0, gb1 = 0.0;
            {
                const double *yrow = sy + n * YST;                // B operand: ob = n, members of lane c
                const int sw = ((n >> 1) & 1) << 2;
#pragma unroll
                for (int t = 0; t < NT3; ++t) {
                    const double2 v = *reinterpret_cast<const double2 *>(yrow + 8 * t + ((2 * c) ^ sw));
                    const double a1 = (t == NT3 - 1 && c == 3) ? 0.0 : x[2 * t + 1];   // mask the mean
                    sp_dmma(ga0, ga1, x[2 * t], v.x);
                    sp_dmma(gb0, gb1, a1, v.y);
                }
            }
            ga0 += gb0;                                            // g[row][2c]
            ga1 += gb1;                                            // g[row][2c+1]
            // all-gather the 8 dots of the row over its 4 lanes
            double gq[8];
            {
                const double o0 = __shfl_xor_sync(0xffffffffu, ga0, 1), o1 = __shfl_xor_sync(0xffffffffu, ga1, 1);
                double q0, q1, q2, q3;
                if (c & 1) { q0 = o0; q1 = o1; q2 = ga0; q3 = ga1; } else { q0 = ga0; q1 = ga1; q2 = o0; q3 = o1; }
                const double r0 = __shfl_xor_sync(0xffffffffu, q0, 2), r1 = __shfl_xor_sync(0xffffffffu, q1, 2);
                const double r2 = __shfl_xor_sync(0xffffffffu, q2, 2), r3 = __shfl_xor_sync(0xffffffffu, q3, 2);
                if (c & 2) { gq[0] = r0; gq[1] = r1; gq[2] = r2; gq[3] = r3; gq[4] = q0; gq[5] = q1; gq[6] = q2; gq[7] = q3; }
                else { gq[0] = q0; gq[1] = q1; gq[2] = q2; gq[3] = q3; gq[4] = r0; gq[5] = r1; gq[6] = r2; gq[7] = r3; }
            }
            double ea0, ea1;
            if (MG) {
                // step 2: e = M g with the producers' per-grid-point matrix (stored negated): this lane needs
                // -e_c (4 terms: M is lower triangular and c <= 3) and -e_{4+c} (8 terms)
                const double *mr0 = Gb + 64 + 50 + gslot * 64 + c * 8;
                const double *mr1 = mr0 + 32;
                const double2 a0 = *reinterpret_cast<const double2 *>(mr0), a1 = *reinterpret_cast<const double2 *>(mr0 + 2);
                const double2 b0 = *reinterpret_cast<const double2 *>(mr1), b1 = *reinterpret_cast<const double2 *>(mr1 + 2);
                const double2 b2 = *reinterpret_cast<const double2 *>(mr1 + 4), b3 = *reinterpret_cast<const double2 *>(mr1 + 6);
                const double s0 = fma(a0.y, gq[1], a0.x * gq[0]), s1 = fma(a1.y, gq[3], a1.x * gq[2]);
                const double t0 = fma(b0.y, gq[1], b0.x * gq[0]), t1 = fma(b1.y, gq[3], b1.x * gq[2]);
                const double t2 = fma(b2.y, gq[5], b2.x * gq[4]), t3 = fma(b3.y, gq[7], b3.x * gq[6]);
                ea0 = active ? s0 + s1 : 0.0;
                ea1 = active ? (t0 + t1) + (t2 + t3) : 0.0;
            } else {
            // step 2: the serial recurrence inside the batch (per row; every lane of the row computes it)
            double e[8];
            e[0] = om[0] * gq[0];
            e[1] = om[1] * (gq[1] - Gb[8] * e[0]);
            e[2] = om[2] * (gq[2] - Gb[16] * e[0] - Gb[17] * e[1]);
            e[3] = om[3] * (gq[3] - Gb[24] * e[0] - Gb[25] * e[1] - Gb[26] * e[2]);
            {
                // obs 4..7 see obs 0..3 through one more 8x8x4 product: corr[row][q] = sum_p e_p G[4+q][p]
                const double ea = (c == 0) ? e[0] : (c == 1) ? e[1] : (c == 2) ? e[2] : e[3];
                double k0 = 0.0, k1 = 0.0;
                sp_dmma(k0, k1, ea, Gb[(4 + (n & 3)) * 8 + c]);
                const double o0 = __shfl_xor_sync(0xffffffffu, k0, 1), o1 = __shfl_xor_sync(0xffffffffu, k1, 1);
                if (c & 1) { gq[4] -= o0; gq[5] -= o1; gq[6] -= k0; gq[7] -= k1; }
                else { gq[4] -= k0; gq[5] -= k1; gq[6] -= o0; gq[7] -= o1; }
            }
            e[4] = om[4] * gq[4];
            e[5] = om[5] * (gq[5] - Gb[44] * e[4]);
            e[6] = om[6] * (gq[6] - Gb[52] * e[4] - Gb[53] * e[5]);
            e[7] = om[7] * (gq[7] - Gb[60] * e[4] - Gb[61] * e[5] - Gb[62] * e[6]);
            ea0 = -((c == 0) ? e[0] : (c == 1) ? e[1] : (c == 2) ? e[2] : e[3]);
            ea1 = -((c == 0) ? e[4] : (c == 1) ? e[5] : (c == 2) ? e[6] : e[7]);
            }

            // step 3: x[row][:] -= sum_q e_q y_q[:]   (A = -e in two k-steps, B = y, C = x)
            {
                const int sw = ((c >> 1) & 1) << 2;                 // rows c and 4+c share this swizzle
                const double *y0p = sy + c * YST + (n ^ sw);
                const double *y1p = y0p + 4 * YST;
                // all tiles with obs 0..3 first, then all with obs 4..7: a tile's second DMMA depends on its first
                // (26 clocks), back to back it would stall the warp's in-order issue
#pragma unroll
                for (int t = 0; t < NT3; ++t) sp_dmma_v(x[2 * t], x[2 * t + 1], ea0, y0p[8 * t]);
#pragma unroll
                for (int t = 0; t < NT3; ++t) sp_dmma_v(x[2 * t], x[2 * t + 1], ea1, y1p[8 * t]);
            }
            dirty = true;
        }
        __syncwarp();
        if (lane == 0) sp_mbar_arrive(s_empty + s);
    }

    // xam of the row (ensrf.py:130) lives in lane c = 3; every lane of the warp takes part in the shuffle
    double mean = __shfl_sync(0xffffffffu, x[2 * NT3 - 1], (lane & ~3) | 3);
    if (!fused) mean = 0.0;
    if (active && dirty) {
#pragma unroll
        for (int t = 0; t < NT3; ++t) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = 8 * t + 2 * c + h;
                if (m < nens) gXp[row * nens + m] = (TS)(x[2 * t + h] + mean);          // assimilation.py:168 when fused
                else if (m == PC && !fused) gxm[row] = (TS)x[2 * t + h];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// candidate lists per coarse tile
// ------------------------------------------------------------------------------------------
// bounding cap (centre, angular radius) of every coarse tile: one warp per tile
__global__ void sweep_tile_caps_kernel(const double *__restrict__ grid_u, int64_t npts, int nx, int y_begin, int y_end, int pr0,
                                       int cty, int ctx, int nctx, int ntiles, float4 *__restrict__ caps) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= ntiles) return;
    const int ya = max((pr0 / SP_CT + t / nctx) * cty, y_begin), yb = min((pr0 / SP_CT + t / nctx + 1) * cty, y_end);
    const int xa = (t % nctx) * ctx, xb = min(xa + ctx, nx);
    const int w = xb - xa, npt = max(yb - ya, 0) * w;
    double cx = 0, cy = 0, cz = 0;
    for (int i = lane; i < npt; i += 32) {
        const int64_t pt = (int64_t)(ya + i / w) * nx + xa + i % w;
        cx += grid_u[pt]; cy += grid_u[npts + pt]; cz += grid_u[2 * npts + pt];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cx += __shfl_xor_sync(0xffffffffu, cx, o); cy += __shfl_xor_sync(0xffffffffu, cy, o); cz += __shfl_xor_sync(0xffffffffu, cz, o);
    }
    const double nn = sqrt(cx * cx + cy * cy + cz * cz);
    if (nn > 1e-9) { cx /= nn; cy /= nn; cz /= nn; }
    double cmin = 1.0;
    for (int i = lane; i < npt; i += 32) {
        const int64_t pt = (int64_t)(ya + i / w) * nx + xa + i % w;
        cmin = fmin(cmin, cx * grid_u[pt] + cy * grid_u[npts + pt] + cz * grid_u[2 * npts + pt]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cmin = fmin(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
    if (lane == 0) {
        // degenerate centre (tile wraps the globe): radius pi -> everything is a candidate
        const double rho = (nn > 1e-9 && npt > 0) ? acos(fmax(-1.0, fmin(1.0, cmin))) + 1e-4 : 3.2;
        caps[t] = make_float4((float)cx, (float)cy, (float)cz, (float)rho);
    }
}

// FILL = false: cnt[t] = number of obs whose support can touch tile t; FILL = true: their ascending list
template <bool FILL>
__global__ void sweep_tile_list_kernel(const float4 *__restrict__ caps, int ntiles, const float4 *__restrict__ scan,
                                       int64_t ob_begin, int64_t ob_end, int *__restrict__ cnt,
                                       const int64_t *__restrict__ off, int *__restrict__ list) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= ntiles) return;
    const float4 cap = caps[t];
    int64_t pos = FILL ? off[t] : 0;
    int count = 0;
    const unsigned lt = (1u << lane) - 1u;
    // One warp walks all obs for its tile: the loop is bound by the latency of the scan-record loads, so the records of
    // TL_UNROLL groups of 32 obs are fetched before the first of them is tested (the groups are still taken in order).
    constexpr int TL_UNROLL = 8;
    for (int64_t base = ob_begin; base < ob_end; base += 32 * TL_UNROLL) {
        float4 sc[TL_UNROLL];
#pragma unroll
        for (int u = 0; u < TL_UNROLL; ++u) {
            const int64_t k = base + 32 * u + lane;
            sc[u] = (k < ob_end) ? __ldg(scan + k) : make_float4(0.f, 0.f, 0.f, -1.f);
        }
#pragma unroll
        for (int u = 0; u < TL_UNROLL; ++u) {
            const int64_t k = base + 32 * u + lane;
            bool hit = false;
            if (sc[u].w >= 0.f) {                  // (past the end of the range: w = -1)
                const float ang = sc[u].w + cap.w;
                hit = (ang >= 3.1405f) || (sc[u].x * cap.x + sc[u].y * cap.y + sc[u].z * cap.z >= __cosf(ang) - 4e-6f);
            }
            if (FILL) {
                const unsigned b = __ballot_sync(0xffffffffu, hit);
                if (hit) list[pos + __popc(b & lt)] = (int)k;
                pos += __popc(b);
            } else {
                count += hit ? 1 : 0;
            }
        }
    }
    if (!FILL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
        if (lane == 0) cnt[t] = count;
    }
}

// grid row of [y_begin, y_end) closest to the equator (|sin lat| of its first column smallest) -> mapped host word
__global__ void sweep_eq_row_kernel(const double *__restrict__ uz, int nx, int y_begin, int y_end, long long *__restrict__ out) {
    __shared__ double bv[256];
    __shared__ int bi[256];
    double best = 2.0;
    int besty = y_begin;
    for (int y = y_begin + threadIdx.x; y < y_end; y += 256) {
        const double v = fabs(uz[(int64_t)y * nx]);
        if (v < best) { best = v; besty = y; }
    }
    bv[threadIdx.x] = best; bi[threadIdx.x] = besty;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o && bv[threadIdx.x + o] < bv[threadIdx.x]) { bv[threadIdx.x] = bv[threadIdx.x + o]; bi[threadIdx.x] = bi[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = bi[0];
}

// exclusive prefix sum of cnt[n] into off[n+1], n small (one CTA, serial over chunks of 1024)
__global__ void __launch_bounds__(1024) sweep_scan_kernel(const int *__restrict__ cnt, int n, int64_t *__restrict__ off,
                                                         long long *__restrict__ total_host) {
    __shared__ long long part[1024];
    __shared__ long long carry_s;
    const int t = threadIdx.x;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int b = 0; b < n; b += 1024) {
        const int i = b + t;
        const long long v = i < n ? cnt[i] : 0;
        part[t] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const long long y = t >= d ? part[t - d] : 0;
            __syncthreads();
            part[t] += y;
            __syncthreads();
        }
        const long long carry = carry_s;
        if (i < n) off[i] = carry + part[t] - v;
        __syncthreads();
        if (t == 1023) carry_s = carry + part[1023];
        __syncthreads();
    }
    if (t == 0) {
        off[n] = carry_s;
        *total_host = carry_s;        // mapped pinned host word: no D2H copy queued behind band downloads
    }
}

// ------------------------------------------------------------------------------------------
// host side: candidate lists per coarse tile (shared with state_sweep_2p.cu)
// ------------------------------------------------------------------------------------------
bool sweep_lists_wanted(int loc_mode, int64_t ob_begin, int64_t ob_end) {
    const char *nolist = getenv("EXB_SWEEP_NOLIST");
    return loc_mode == EXB_LOC_GC && !(nolist && atoi(nolist)) && ob_end - ob_begin > 4096;
}

// Builds, for every coarse tile (SP_CT x SP_CT patches of bty x btx grid points) that intersects grid rows
// [y_begin, y_end), the ascending list of obs in [ob_begin, ob_end) whose support can touch the tile.  Synchronises the
// stream once (list sizes).  out->tile_off is indexed by ABSOLUTE coarse-tile number (coarse row * nctx + coarse col).
int sweep_build_lists(const double *grid_u, int64_t npts, int nx, int y_begin, int y_end, int bty, int btx, const float4 *scan,
                      int64_t ob_begin, int64_t ob_end, cudaStream_t st, SweepLists *out) {
    const int pr0 = y_begin / bty, pr1 = (y_end + bty - 1) / bty;
    const int cty = bty * SP_CT, ctx = btx * SP_CT;
    out->nctx = (nx + ctx - 1) / ctx;
    const int cr0 = pr0 / SP_CT, cr1 = (pr1 + SP_CT - 1) / SP_CT;
    const int ntiles = (cr1 - cr0) * out->nctx;
    EXB_CUDA(exb_malloc_async(&out->caps, sizeof(float4) * ntiles, st));
    EXB_CUDA(exb_malloc_async(&out->cnt, sizeof(int) * ntiles, st));
    EXB_CUDA(exb_malloc_async(&out->off, sizeof(int64_t) * (ntiles + 1), st));
    const unsigned gridw = (unsigned)ceil_div64((int64_t)ntiles * 32, 256);
    sweep_tile_caps_kernel<<<gridw, 256, 0, st>>>(grid_u, npts, nx, y_begin, y_end, pr0, cty, ctx, out->nctx, ntiles, out->caps);
    sweep_tile_list_kernel<false><<<gridw, 256, 0, st>>>(out->caps, ntiles, scan, ob_begin, ob_end, out->cnt, nullptr, nullptr);
    ExbHostWords hw;                   // this call's own mapped words: [0] list total, [1] equator row
    const int rcw = exb_host_words_acquire(&hw);
    if (rcw != EXB_OK) return rcw;
    struct WordsGuard { ExbHostWords w; ~WordsGuard() { exb_host_words_release(w); } } wguard{hw};
    sweep_scan_kernel<<<1, 1024, 0, st>>>(out->cnt, ntiles, out->off, hw.dev);
    sweep_eq_row_kernel<<<1, 256, 0, st>>>(grid_u + 2 * npts, nx, y_begin, y_end, hw.dev + 1);
    exb_count_launches(4);
    EXB_CUDA(cudaStreamSynchronize(st));
    const long long total = hw.host[0];
    out->eq_row = (int)hw.host[1];
    EXB_CUDA(exb_malloc_async(&out->list, sizeof(int) * (size_t)(total > 0 ? total : 1), st));
    sweep_tile_list_kernel<true><<<gridw, 256, 0, st>>>(out->caps, ntiles, scan, ob_begin, ob_end, nullptr, out->off, out->list);
    exb_count_launches(1);
    // the kernels index tiles by absolute coarse row: shift the offsets' base
    out->tile_off = out->off - (int64_t)cr0 * out->nctx;
    return exb_check_launch("sweep_tile_list_kernel");
}

void sweep_free_lists(SweepLists &l, cudaStream_t st) {
    if (l.caps) cudaFreeAsync(l.caps, st);
    if (l.cnt) cudaFreeAsync(l.cnt, st);
    if (l.off) cudaFreeAsync(l.off, st);
    if (l.list) cudaFreeAsync(l.list, st);
    l.caps = nullptr; l.cnt = nullptr; l.off = nullptr; l.list = nullptr;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int NT3, typename TS, bool MG>
static int sp_launch(SpParams &p, cudaStream_t st) {
    const int Lc = p.nlev < SP_ROWS ? p.nlev : SP_ROWS;
    const int G = SP_ROWS / Lc;
    int bty = 1, btx = G;
    for (int ty = 1; ty * ty <= G; ++ty) {
        const int tx = G / ty;
        if (ty * tx > bty * btx || (ty * tx == bty * btx && ty > bty)) { bty = ty; btx = tx; }
    }
    if (btx > p.nx) btx = p.nx;
    if (bty > p.ny) bty = p.ny;
    p.ty = bty; p.tx = btx; p.G = bty * btx; p.Lc = Lc;
    p.nlc = (p.nlev + Lc - 1) / Lc;
    p.ntx = (p.nx + btx - 1) / btx;
    p.pr0 = p.y_begin / bty;
    const int pr1 = (p.y_end + bty - 1) / bty;              // patch rows [pr0, pr1)
    const int nty = pr1 - p.pr0;
    p.npr = nty;
    p.pr_eq = p.eq_row >= 0 ? p.eq_row / bty : -1;
    p.stage_doubles = sp_stage_doubles<NT3>(p.G, MG);
    int dev = 0, max_smem = 0;
    EXB_CUDA(cudaGetDevice(&dev));
    EXB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t fixed = sizeof(double) * 3 * SP_ROWS + sizeof(unsigned long long) * 2 * SP_MAXSTAGES +
                         sizeof(int) * (SP_ROWS + SP_PW * 32 + SP_PW * 128) + 64;
    int S = (int)(((size_t)max_smem - 1024 - fixed) / (sizeof(double) * p.stage_doubles));
    if (S > SP_MAXSTAGES) S = SP_MAXSTAGES;
    S -= S % SP_PW;                    // every use of a stage is staged by the same producer warp (parity waits)
    if (S < 8) return EXB_ERR_UNSUPPORTED;
    p.nstages = S;
    const size_t smem = fixed + sizeof(double) * (size_t)S * p.stage_doubles;
    EXB_CUDA(cudaFuncSetAttribute(state_sweep_pipe_kernel<NT3, TS, MG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    // candidate lists per coarse tile (localised runs only; the kernel walks the ob range otherwise)
    SweepLists lists;
    p.tile_off = nullptr;
    p.tile_list = nullptr;
    p.nctx = 1;
    if (sweep_lists_wanted(p.loc_mode, p.ob_begin, p.ob_end)) {
        const int rcl = sweep_build_lists(p.grid_u, p.npts, p.nx, p.y_begin, p.y_end, bty, btx, p.scan, p.ob_begin, p.ob_end, st, &lists);
        if (rcl != EXB_OK) return rcl;
        p.nctx = lists.nctx;
        p.eq_row = lists.eq_row;
        p.pr_eq = p.eq_row / bty;
        p.tile_off = lists.tile_off;
        p.tile_list = lists.list;
    }
    const int64_t nblocks = (int64_t)p.ntx * nty * p.nlc;
    int rc = EXB_OK;
    if (nblocks >= 0x7fffffff) {
        exb_set_error("exb_state_sweep: too many patches for one launch");
        rc = EXB_ERR_ARG;
    } else if (nblocks > 0) {
        state_sweep_pipe_kernel<NT3, TS, MG><<<(unsigned)nblocks, SP_NT, smem, st>>>(p);
        exb_count_launches(1);
        rc = exb_check_launch("state_sweep_pipe_kernel");
    }
    sweep_free_lists(lists, st);
    return rc;
}

template <typename TS>
int exb_state_sweep_2p(TS *xm, TS *Xp, int64_t nlev, int64_t ny, int64_t nx, int nens, const double *grid_u, const TS *Yp,
                       const double *rec, const double *obgeo, const float4 *scan, int64_t nobs, int64_t ob_begin,
                       int64_t ob_end, int64_t y_begin, int64_t y_end, int loc_mode, unsigned long long *counters,
                       cudaStream_t st, const ExbSweepPlan *plan);
void s2_patch_shape(int64_t nlev, int64_t ny, int64_t nx, int *Lc_out, int *bty_out, int *btx_out);

// EXB_SP_IMPL = 2p (default: two-phase kernel, state_sweep_2p.cu) | v3 (the concurrent producer/consumer kernel above)
static bool sp_use_2p() {
    const char *e = getenv("EXB_SP_IMPL");
    return !(e && strcmp(e, "v3") == 0);
}

// Row granularity of the patches for a state with nlev levels: sweeping row ranges whose edges are multiples of
// this value never splits a patch between two calls.
extern "C" int exb_state_sweep_row_granularity(int64_t nlev, int64_t ny, int64_t nx) {
    if (sp_use_2p()) {
        int Lc, bty, btx;
        s2_patch_shape(nlev, ny, nx, &Lc, &bty, &btx);
        return bty;
    }
    const int Lc = nlev < SP_ROWS ? (int)nlev : SP_ROWS;
    const int G = SP_ROWS / Lc;
    int bty = 1, btx = G;
    for (int ty = 1; ty * ty <= G; ++ty) {
        const int tx = G / ty;
        if (ty * tx > bty * btx || (ty * tx == bty * btx && ty > bty)) { bty = ty; btx = tx; }
    }
    if (btx > nx) btx = (int)nx;
    if (bty > ny) bty = (int)ny;
    return bty;
}

// Called from state_update.cu.  TS is the storage type of state and ye rows (float64, or float32 with float64
// arithmetic in registers: every row is read and rounded back exactly once).  xm == nullptr selects the fused
// split/recombine mode (Xp then holds full ensemble values).  Returns EXB_ERR_UNSUPPORTED if no variant fits.
template <typename TS>
int exb_state_sweep_pipe(TS *xm, TS *Xp, int64_t nlev, int64_t ny, int64_t nx, int nens, const double *grid_u, const TS *Yp,
                         const double *rec, const double *obgeo, const float4 *scan, int64_t nobs, int64_t ob_begin,
                         int64_t ob_end, int64_t y_begin, int64_t y_end, int loc_mode, unsigned long long *counters,
                         cudaStream_t st, const ExbSweepPlan *plan) {
    if (sp_use_2p()) {
        const int rc = exb_state_sweep_2p<TS>(xm, Xp, nlev, ny, nx, nens, grid_u, Yp, rec, obgeo, scan, nobs, ob_begin, ob_end,
                                              y_begin, y_end, loc_mode, counters, st, plan);
        if (rc != EXB_ERR_UNSUPPORTED) return rc;
    }
    SpParams p;
    p.xm = xm; p.Xp = Xp; p.Yp = Yp; p.grid_u = grid_u; p.rec = rec; p.geo = obgeo; p.scan = scan;
    p.tile_off = nullptr; p.tile_list = nullptr;
    p.counters = counters; p.npts = ny * nx; p.nobs = nobs; p.ob_begin = ob_begin; p.ob_end = ob_end;
    p.nlev = (int)nlev; p.ny = (int)ny; p.nx = (int)nx; p.nens = nens; p.loc_mode = loc_mode;
    p.y_begin = (int)y_begin; p.y_end = (int)y_end;
    p.eq_row = -1;
    p.role_split = getenv("EXB_SP_SPLIT") ? atoi(getenv("EXB_SP_SPLIT")) : 1;
    const int need = (nens + 1 + 7) / 8;            // 8-member tiles incl. the pseudo-member
    // The matrix form of the recurrence (MG) needs 64 more doubles per grid point and stage and ~290 more scalar FP64
    // instructions per batch in the producers.  Measured on config 3: 202 ms against 160 ms for the chain form (the
    // ring shrinks to 8 stages and the producers become the bottleneck again), so it is off unless EXB_SP_MG=1.
    const bool want_mg = getenv("EXB_SP_MG") && atoi(getenv("EXB_SP_MG")) == 1;
#define SP_TRY(N)                                                        \
    if (need <= N) {                                                     \
        if (want_mg) {                                                   \
            const int rc = sp_launch<N, TS, true>(p, st);                \
            if (rc != EXB_ERR_UNSUPPORTED) return rc;                    \
        }                                                                \
        return sp_launch<N, TS, false>(p, st);                           \
    }
    SP_TRY(4);
    SP_TRY(7);
    SP_TRY(10);
    SP_TRY(13);
#undef SP_TRY
    return EXB_ERR_UNSUPPORTED;                       // larger ensembles: state_update_mma.cu / state_update.cu
}

template int exb_state_sweep_pipe<double>(double *, double *, int64_t, int64_t, int64_t, int, const double *, const double *,
                                          const double *, const double *, const float4 *, int64_t, int64_t, int64_t, int64_t,
                                          int64_t, int, unsigned long long *, cudaStream_t, const ExbSweepPlan *);
template int exb_state_sweep_pipe<float>(float *, float *, int64_t, int64_t, int64_t, int, const double *, const float *,
                                         const double *, const double *, const float4 *, int64_t, int64_t, int64_t, int64_t,
                                         int64_t, int, unsigned long long *, cudaStream_t, const ExbSweepPlan *);

// ------------------------------------------------------------------------------------------
// sweep plan: the geometry-only part of the sweep, built ahead of the obs-space solve
// ------------------------------------------------------------------------------------------
__global__ void sweep_scan_from_flags_kernel(const double *__restrict__ geo, const uint8_t *__restrict__ assim, int64_t nobs,
                                             float4 *__restrict__ scan) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nobs) return;
    float4 v;
    v.x = (float)geo[GEO_UX * nobs + k];
    v.y = (float)geo[GEO_UY * nobs + k];
    v.z = (float)geo[GEO_UZ * nobs + k];
    // same record as su_scan_records_kernel (state_update.cu) builds from rec[REC_ASSIM] after the solve
    v.w = assim[k] ? (float)geo[GEO_THETA * nobs + k] * 1.000001f + 1e-7f : -1.f;
    scan[k] = v;
}

static void sweep_plan_free(ExbSweepPlan *pl) {
    if (!pl) return;
    if (pl->was_used && pl->used) cudaStreamWaitEvent(pl->st, pl->used, 0);
    if (pl->scan) cudaFreeAsync(pl->scan, pl->st);
    sweep_free_lists(pl->lists, pl->st);
    if (pl->ready) cudaEventDestroy(pl->ready);
    if (pl->used) cudaEventDestroy(pl->used);
    delete pl;
}

extern "C" int exb_sweep_plan_create(const double *grid_u, int64_t nlev, int64_t ny, int64_t nx, const double *obgeo,
                                     const uint8_t *ob_assimilate, int64_t nobs, int64_t ob_begin, int64_t ob_end,
                                     int64_t y_begin, int64_t y_end, int loc_mode, void *stream, void **plan) {
    EXB_REQUIRE(grid_u && obgeo && ob_assimilate && plan, "null pointer");
    EXB_REQUIRE(nlev > 0 && ny > 0 && nx > 0 && nobs > 0, "bad sizes");
    EXB_REQUIRE(0 <= ob_begin && ob_begin <= ob_end && ob_end <= nobs && 0 <= y_begin && y_begin < y_end && y_end <= ny, "bad ranges");
    *plan = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    ExbSweepPlan *pl = new ExbSweepPlan();
    struct Guard { ExbSweepPlan *p; ~Guard() { if (p) sweep_plan_free(p); } } guard{pl};
    pl->nlev = nlev; pl->ny = ny; pl->nx = nx; pl->nobs = nobs; pl->ob_begin = ob_begin; pl->ob_end = ob_end;
    pl->y_begin = y_begin; pl->y_end = y_end; pl->loc_mode = loc_mode; pl->grid_u = grid_u; pl->obgeo = obgeo; pl->st = st;
    EXB_CUDA(cudaEventCreateWithFlags(&pl->ready, cudaEventDisableTiming));
    EXB_CUDA(cudaEventCreateWithFlags(&pl->used, cudaEventDisableTiming));
    EXB_CUDA(exb_malloc_async(&pl->scan, (size_t)nobs * sizeof(float4), st));
    sweep_scan_from_flags_kernel<<<(unsigned)ceil_div64(nobs, 256), 256, 0, st>>>(obgeo, ob_assimilate, nobs, pl->scan);
    exb_count_launches(1);
    if (sp_use_2p() && sweep_lists_wanted(loc_mode, ob_begin, ob_end)) {
        int Lc;
        s2_patch_shape(nlev, ny, nx, &Lc, &pl->bty, &pl->btx);
        const int rc = sweep_build_lists(grid_u, ny * nx, (int)nx, (int)y_begin, (int)y_end, pl->bty, pl->btx, pl->scan, ob_begin,
                                         ob_end, st, &pl->lists);
        if (rc != EXB_OK) return rc;
        pl->have_lists = true;
    }
    EXB_CUDA(cudaEventRecord(pl->ready, st));
    guard.p = nullptr;
    *plan = pl;
    return exb_check_launch("sweep_scan_from_flags_kernel");
}

extern "C" int exb_sweep_plan_destroy(void *plan) {
    sweep_plan_free(static_cast<ExbSweepPlan *>(plan));
    return EXB_OK;
}
