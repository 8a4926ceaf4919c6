// Shared device/host helpers for the efa_xray_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/efa_xray_b200.h"

#define EXB_R_EARTH 6371.0            // km; state/ensemble.py:259, observation/observation.py:138
#define EXB_DEG2RAD 0.017453292519943295

// obgeo rows
#define GEO_UX 0
#define GEO_UY 1
#define GEO_UZ 2
#define GEO_INVHW 3
#define GEO_AMAX 4
#define GEO_COST 5
#define GEO_SINT 6
#define GEO_THETA 7
// rec rows
#define REC_PRIOR_MEAN 0
#define REC_PRIOR_VAR 1
#define REC_POST_MEAN 2
#define REC_POST_VAR 3
#define REC_INNOV 4
#define REC_C1 5
#define REC_BETA 6
#define REC_ASSIM 7

void exb_set_error(const char *fmt, ...);
int exb_check_launch(const char *what);
void exb_count_launches(int64_t n);   // bookkeeping for exb_launch_count()

#define EXB_REQUIRE(cond, msg)                         \
    do {                                               \
        if (!(cond)) {                                 \
            exb_set_error("%s: %s", __func__, msg);    \
            return EXB_ERR_ARG;                        \
        }                                              \
    } while (0)

#define EXB_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            exb_set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e__)); \
            return EXB_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

// Haversine 'a' from two unit vectors: a = sin^2(d/2) = |u - v|^2 / 4.  Mathematically the same
// quantity as state/ensemble.py:264 and observation.py:144; computed from differences so it keeps
// full relative accuracy for nearby points.
__device__ __forceinline__ double hav_a(double ux, double uy, double uz, double vx, double vy, double vz) {
    const double dx = ux - vx, dy = uy - vy, dz = uz - vz;
    return 0.25 * (dx * dx + dy * dy + dz * dz);
}

// Great-circle angle c = 2*atan2(sqrt(a), sqrt(1-a))   (state/ensemble.py:266).
__device__ __forceinline__ double angle_from_a(double a) {
    a = fmin(fmax(a, 0.0), 1.0);
    if (a < 0.5) return 2.0 * asin(sqrt(a));            // same function, cheaper and exact to 1 ulp here
    return 2.0 * atan2(sqrt(a), sqrt(1.0 - a));
}

// Gaspari-Cohn weight of r = distance / |halfwidth|   (observation/observation.py:120-130).
__device__ __forceinline__ double gaspari_cohn_r(double r) {
    if (r <= 1.0) return ((((-0.25 * r + 0.5) * r + 0.625) * r - 5.0 / 3.0) * (r * r) + 1.0);
    if (r < 2.0)
        return (((((r / 12.0 - 0.5) * r + 0.625) * r + 5.0 / 3.0) * r - 5.0) * r + 4.0 - 2.0 / (3.0 * r));
    return 0.0;
}

// Localisation weight of a pair from its haversine a.  inv_hw = 1/|halfwidth| (0 when localisation
// is off, which makes r = 0 and the weight exactly 1); a_max is the value of a at r = 2.
__device__ __forceinline__ double loc_weight(double a, double inv_hw, double a_max) {
    if (a >= a_max) return 0.0;
    const double r = EXB_R_EARTH * angle_from_a(a) * inv_hw;
    return gaspari_cohn_r(r);
}

template <typename T> struct Vec2;
template <> struct Vec2<double> { typedef double2 type; };
template <> struct Vec2<float> { typedef float2 type; };

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
