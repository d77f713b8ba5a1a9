// =====================================================================================
//  adt_kernels.cuh -- sm_100a kernels of the batched MOHID property transport step.
//
//  Reference semantics (paths relative to /root/reference/Software):
//     AD = MOHIDBase2/ModuleAdvectionDiffusion.F90, MF = MOHIDBase1/ModuleFunctions.F90
//
//  K1  adt_coef_kernel      per-step shared coefficients (property independent):
//                           Convert_Dif_Vertical / Convert_Visc_Dif_Horizontal (AD:2364-2675),
//                           Compute_DifH/DifV_Constants (AD:1514-1619), DT/V, Vold/V, the packed
//                           1-byte mask and 1/(DWZ(k)+DWZ(k-1)).
//  K2  adt_transport_kernel the fused step, one thread per (water column, property):
//                           VolumeVariation (AD:3966) + explicit horizontal diffusion/advection
//                           (AD:5123-5365, 4368-4953; face weights MF:10702-10894) + vertical
//                           diffusion/advection assembly (AD:2708-3207) + open-boundary rows
//                           (AD:5369-5672) + land fill (AD:1753) + THOMASZ_NewType2 (MF:4026-4123).
//                           D,E,F,TI never touch HBM; W,G of the column solve live in shared memory.
//  K3  adt_nullgrad_kernel / adt_cyclic_kernel   post-solve boundary passes (AD:1926-1987, 2121-2224).
//  K4  adt_pack/unpack_columns_kernel            j-slab halo staging for the NCCL exchange
//                                                (replaces ReceiveSendProperities3DMPIr8, HG:8479-8658).
//
//  Arithmetic contract: fp64 throughout.  The reference's divisions by per-cell volumes and by
//  metric sums are replaced by multiplications with reciprocals computed once per step in K1
//  (DT/V, 1/(du_a+du_b)); results agree with the reference to rounding (tests: <= 1e-13 relative per
//  step, <= 1e-10 after 100 steps) while masks, land values (null_real), untouched dry columns and
//  the KUB+1 halo row are bit-exact.
// =====================================================================================
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/mohid_adt.h"

namespace adt {

constexpr int NPMAX = 32;                 // max properties per launch (kernel-parameter arrays)
constexpr double NULL_REAL = MOHID_NULL_REAL;
constexpr double MIN_VALUE = 1.e-16;      // MGD:1812

// packed per-cell mask byte written by K1
enum : unsigned {
    M_OPEN = 1u,      // OpenPoints3D == 1
    M_CFU = 2u,       // ComputeFacesU3D == 1 (west face of the cell)
    M_CFV = 4u,       // ComputeFacesV3D == 1 (south face)
    M_CFW = 8u,       // ComputeFacesW3D == 1 (bottom face)
    M_LAND = 16u,     // LandPoints3D == 1
    M_BND = 32u,      // BoundaryPoints2D(i,j) == 1
    M_COLWET = 64u,   // WaterPoints3D(i,j,KUB) == 1: the column is solved (MF:4086)
    M_COLOPEN = 128u  // OpenPoints3D(i,j,KUB) == 1: vertical advection coefficients are built (AD:2966)
};

struct CoefArgs {
    int ni, nj, nk, ld;              // allocated extents (I+2, J+2, K+2) and leading dimension
    int I, J, K;
    double dt, schmidt_h, schmidt_coef_v, schmidt_bg_v;
    int nulldif;
    // raw interface arrays (device copies)
    const double *Wflux_X, *Wflux_Y, *Wflux_Z, *VolumeZOld, *VolumeZ, *Visc_H, *Diff_V, *DWZ, *DZZ, *AreaU, *AreaV;
    const int *Open, *Land, *Water, *CFU, *CFV, *CFW, *SmallDepths;
    const double *DUX, *DVY, *DZX, *DZY;
    const int *Bnd;
    // outputs
    double *dtv, *vr, *dhu, *dhv, *dvz, *rdz;
    uint8_t *mask;
    int do_geom, do_diff;            // which parts to (re)build
};

// -------------------------------------------------------------------------------------
// K1: one thread per allocated cell, i fastest (coalesced); memory bound (112 B in, 49 B out).
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adt_coef_kernel(const CoefArgs a) {
    const long n3 = (long)a.ld * a.nj * a.nk;
    const long sj = a.ld, sk = (long)a.ld * a.nj;
    for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n3; q += (long)gridDim.x * blockDim.x) {
        const int i = (int)(q % a.ld);
        const int j = (int)((q / a.ld) % a.nj);
        const int k = (int)(q / sk);
        if (i >= a.ni) {                                  // leading-dimension padding
            if (a.do_geom) { a.mask[q] = 0; a.dtv[q] = 0.; a.vr[q] = 1.; a.rdz[q] = 0.; }
            if (a.do_diff) { a.dhu[q] = 0.; a.dhv[q] = 0.; a.dvz[q] = 0.; }
            continue;
        }
        const long q2 = (long)i + sj * j;
        const bool cfu = a.CFU[q] == 1, cfv = a.CFV[q] == 1, cfw = a.CFW[q] == 1;
        if (a.do_geom) {
            unsigned m = 0;
            if (a.Open[q] == 1) m |= M_OPEN;
            if (cfu) m |= M_CFU;
            if (cfv) m |= M_CFV;
            if (cfw) m |= M_CFW;
            if (a.Land[q] == 1) m |= M_LAND;
            if (a.Bnd[q2] == 1) m |= M_BND;
            const long qtop = q2 + sk * a.K;
            if (a.Water[qtop] == 1) m |= M_COLWET;
            if (a.Open[qtop] == 1) m |= M_COLOPEN;
            a.mask[q] = (uint8_t)m;
            const double V = a.VolumeZ[q];
            const bool inwork = (i >= 1 && i <= a.I && j >= 1 && j <= a.J && k >= 1 && k <= a.K);
            a.dtv[q] = (inwork && V != 0.) ? a.dt / V : 0.;
            a.vr[q] = (inwork && V != 0.) ? a.VolumeZOld[q] / V : 1.;
            double s = (k >= 1) ? (a.DWZ[q] + a.DWZ[q - sk]) : 0.;
            a.rdz[q] = (s != 0.) ? 1.0 / s : 0.;
        }
        if (a.do_diff) {
            double hu = 0., hv = 0., vz = 0.;
            if (cfu && j >= 1) {
                // DifX (AD:2486-2495) then Diff_H_Const_U (AD:1549-1553), same operation order
                const double dux = a.DUX[q2], duxm = a.DUX[q2 - sj];
                double difx = a.schmidt_h * (a.Visc_H[q] * duxm + a.Visc_H[q - sj] * dux) / (dux + duxm);
                if (a.nulldif && a.Wflux_X[q] == 0.) difx = 0.;
                hu = difx * a.AreaU[q] / a.DZX[q2 - sj];
            }
            if (cfv && i >= 1) {
                const double dvy = a.DVY[q2], dvym = a.DVY[q2 - 1];
                double dify = a.schmidt_h * (a.Visc_H[q] * dvym + a.Visc_H[q - 1] * dvy) / (dvy + dvym);
                if (a.nulldif && a.Wflux_Y[q] == 0.) dify = 0.;
                hv = dify * a.AreaV[q] / a.DZY[q2 - 1];
            }
            if (cfw && k >= 1 && !(a.SmallDepths && a.SmallDepths[q2] != 0)) {
                // DifZ (AD:2397-2405) then Diff_V_Const (AD:1591-1597)
                double difz = (a.schmidt_coef_v * a.Diff_V[q] + a.schmidt_bg_v);
                if (a.nulldif && a.Wflux_Z[q] == 0.) difz = 0.;
                const double auxk = difz * a.DUX[q2] * a.DVY[q2];
                vz = auxk / a.DZZ[q - sk];
            }
            a.dhu[q] = hu; a.dhv[q] = hv; a.dvz[q] = vz;
        }
    }
}

// 2-D reciprocal metric sums: rdx(i,j) = 1/(DUX(i,j)+DUX(i,j-1)), rdy(i,j) = 1/(DVY(i,j)+DVY(i-1,j))
__global__ void adt_grid2d_kernel(int ni, int nj, int ld, const double *DUX, const double *DVY, double *rdx,
                                  double *rdy) {
    const long n2 = (long)ld * nj;
    for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n2; q += (long)gridDim.x * blockDim.x) {
        const int i = (int)(q % ld), j = (int)(q / ld);
        double sx = (j >= 1 && i < ni) ? DUX[q] + DUX[q - ld] : 0.;
        double sy = (i >= 1 && i < ni) ? DVY[q] + DVY[q - 1] : 0.;
        rdx[q] = sx != 0. ? 1.0 / sx : 0.;
        rdy[q] = sy != 0. ? 1.0 / sy : 0.;
    }
}

// -------------------------------------------------------------------------------------
// K2 arguments
// -------------------------------------------------------------------------------------
struct PropArgs {
    const double *pin;      // property at time n (read)
    double *pout;           // property at time n+1 (written; ping-pong buffer)
    const double *pref;     // ReferenceProp or nullptr
    double theta_difv;      // effective ImpExp_DifV (Optimize path: >0 -> 1, AD:2797)
    double tdec;            // 1/(1+DecayTime/DT) (AD:5418-5419)
    int bc;                 // MOHID_BC_*
    int advv_implicit;      // ImpExp_AdvV == ImplicitScheme (AD:3087)
};

struct StepArgs {
    int I, J, K, ld, nj;
    long sk;                                    // plane stride ld*nj
    int nprop, ntile_i;                         // tiles of 31 cells along i
    int j_begin, j_count;                       // columns j_begin .. j_begin+j_count-1 are advanced
    int method_h, limiter_h, method_v, limiter_v, upwind2_h, upwind2_v;
    int vertical1d, xzflow;
    double vrelmax, dt;
    const double *qx, *qy, *qz, *dtv, *vr, *dhu, *dhv, *dvz, *rdz;
    const uint8_t *mask;
    const double *rdx, *rdy, *DUX, *DVY, *DWZ;
    const double *VolumeZ, *VolumeZOld;         // open-boundary flux only (AD:5718-5727)
    unsigned long long *zero_pivots;
    PropArgs p[NPMAX];
};

__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_dn_d(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

// TVD limiter psi(r) (MF:10812-10856); Cr may be overwritten by the PDM branch (quirk A.4-1)
__device__ __forceinline__ double tvd_psi(int limiter, double r, double &Cr) {
    switch (limiter) {
        case MOHID_MinMod: return fmax(0., fmin(1., r));
        case MOHID_VanLeer: return (r < 0.) ? 0. : 2. * r / (1. + r);
        case MOHID_Muscl: return fmax(0., fmin(fmin(2., 2. * r), (1. + r) / 2.));
        case MOHID_SuperBee: return fmax(fmax(0., fmin(1., 2. * r)), fmin(r, 2.));
        default: {  // PDM
            const double c = (1. - 2. * fabs(Cr)) / 6.;
            const double a = 0.5 + c, b = 0.5 - c;
            const double aux = a + b * r;
            if (fabs(Cr) < MIN_VALUE) Cr = MIN_VALUE;
            return fmax(0., fmin(fmin(aux, 2. / (1. - Cr)), 2. * r / Cr));
        }
    }
}

// Face weights CFace(1..4) of ComputeAdvectionFace (MF:10702-10894) for the stencil cells
// (f-2, f-1, f, f+1) = (1,2,3,4).  o1/o4: the outer cells are compute points (near-boundary test,
// MF:10563-10570).  dtv1..4 = DT/V of the four cells (signed Courant = Q*dtv, MF:11045-11053;
// VolumeRel of methods 2/3 = max(dtv)/min(dtv)); rd12, rd23, rd34 = 1/(du_a+du_b); du2, du3 only
// for the centred schemes.
__device__ __forceinline__ void face_weights(int method, int limiter, bool upwind2, double vrelmax, double Q,
                                             double P1, double P2, double P3, double P4, bool o1, bool o4,
                                             double dtv1, double dtv2, double dtv3, double dtv4, double rd12,
                                             double rd23, double rd34, double du2, double du3, double &c1,
                                             double &c2, double &c3, double &c4) {
    const bool pos = Q > 0.;
    const bool near = pos ? !o1 : !o4;
    c1 = 0.; c4 = 0.;
    if (method == MOHID_UpwindOrder1 || (near && upwind2)) {
        c2 = pos ? 1. : 0.;
        c3 = pos ? 0. : 1.;
        return;
    }
    if (method == MOHID_P2_TVD) {
        double Cr = Q * (pos ? dtv2 : dtv3);
        const double dP = pos ? (P3 - P2) : (P2 - P3);
        double dC = dP * rd23;
        if (fabs(dC) < MIN_VALUE) dC = (dC >= 0.) ? MIN_VALUE : -MIN_VALUE;
        const double up = pos ? (P2 - P1) * rd12 : (P3 - P4) * rd34;
        const double r = up / dC;
        double theta = tvd_psi(limiter, r, Cr);
        theta = 0.5 * theta * (1. - Cr);
        c2 = pos ? (1. - theta) : theta;
        c3 = pos ? theta : (1. - theta);
        return;
    }
    if (method == MOHID_CentralDif || method == MOHID_LeapFrog) {
        c2 = du3 * rd23;
        c3 = du2 * rd23;
        return;
    }
    // UpwindOrder2 (QUICK) / UpwindOrder3 (QUICKEST), interior faces (MF:10751-10770, 11055-11125)
    {
        const double ta = pos ? dtv1 : dtv2, tb = pos ? dtv2 : dtv3, tc = pos ? dtv3 : dtv4;
        const double tmax = fmax(fmax(ta, tb), tc), tmin = fmin(fmin(ta, tb), tc);
        const bool first_order = (tmax / tmin > vrelmax) || (Q == 0.);
        if (first_order) {
            c2 = pos ? 1. : 0.;
            c3 = pos ? 0. : 1.;
            return;
        }
        double h1, h2, h3;      // weights of (2nd upwind, upwind, downwind)
        if (method == MOHID_UpwindOrder2) {
            h1 = -1. / 8.; h2 = 6. / 8.; h3 = 3. / 8.;
        } else {
            const double Cr = Q * (pos ? dtv2 : dtv3);
            const double c = (1. - 2. * fabs(Cr)) / 6.;
            const double a = 0.5 + c, b = 0.5 - c, d = (1. - fabs(Cr)) / 2.;
            h1 = -d * b; h2 = 1. + d * (b - a); h3 = d * a;
        }
        c2 = pos ? h2 : h3;
        c3 = pos ? h3 : h2;
        if (pos) c1 = h1; else c4 = h1;
    }
}

// Total (advective - diffusive) property flux through a horizontal face, positive toward +index.
// Applied iff the face is a compute face; advective weights exist iff both adjacent cells are open
// (MF:10559, AD:4467, AD:5188).
__device__ __forceinline__ double hface_flux(const StepArgs &s, bool cf, bool o2, bool o3, double Q, double dh,
                                             double P1, double P2, double P3, double P4, bool o1, bool o4,
                                             double dtv1, double dtv2, double dtv3, double dtv4, double rd12,
                                             double rd23, double rd34, double du2, double du3) {
    if (!cf) return 0.;
    double f = -dh * (P3 - P2);
    if (o2 && o3) {
        double c1, c2, c3, c4;
        face_weights(s.method_h, s.limiter_h, s.upwind2_h != 0, s.vrelmax, Q, P1, P2, P3, P4, o1, o4, dtv1, dtv2,
                     dtv3, dtv4, rd12, rd23, rd34, du2, du3, c1, c2, c3, c4);
        f += Q * (c1 * P1 + c2 * P2 + c3 * P3 + c4 * P4);
    }
    return f;
}

// -------------------------------------------------------------------------------------
// K2: fused transport step.
//   warp  <-> (31-cell strip along i, column j, property n);  lane <-> cell i0+lane; lane 31 only
//   supplies the north face of lane 30.  The thread marches k = 1..K building row k of the
//   tridiagonal system in registers, eliminating it on the fly (W,G -> shared memory), then
//   back-substitutes and writes the new property.  Units are ordered property-fastest so the
//   warps of a block read the same shared coefficients (L1 hits).
//   blockDim.x = 32 * WPB; dynamic shared memory = 2 * K * WPB * 32 doubles.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(384, 1) adt_transport_kernel(const __grid_constant__ StepArgs s) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, WPB = blockDim.x >> 5;
    double *__restrict__ Wsm = smem + warp * 32 + lane;                  // [K][WPB][32]
    double *__restrict__ Gsm = Wsm + (size_t)s.K * WPB * 32;
    const int wstride = WPB * 32;
    const long nunits = (long)s.nprop * s.ntile_i * s.j_count;
    const long unit = (long)blockIdx.x * WPB + warp;
    if (unit >= nunits) return;
    const int n = (int)(unit % s.nprop);
    const int tile = (int)((unit / s.nprop) % s.ntile_i);
    const int j = (int)(unit / ((long)s.nprop * s.ntile_i)) + s.j_begin;
    const int i = 1 + tile * 31 + lane;
    const bool writer = (lane < 31) && (i <= s.I);
    const int ic = min(i, s.I + 1);                       // clamped column: every load stays in bounds
    const PropArgs &pa = s.p[n];
    const double *__restrict__ P = pa.pin;
    const long sj = s.ld, sk = s.sk;
    const long c2d = (long)ic + sj * j;
    const bool jp2 = (j + 2 <= s.J + 1);

    // ---- 2-D metrics of the column ----
    const double rdx_m = s.rdx[c2d - sj], rdx_c = s.rdx[c2d], rdx_p = s.rdx[c2d + sj];
    const double rdx_pp = jp2 ? s.rdx[c2d + 2 * sj] : 0.;
    const double rdy_c = s.rdy[c2d];
    double rdy_m = shfl_up_d(rdy_c, 1), rdy_p = shfl_dn_d(rdy_c, 1);
    if (lane == 0) rdy_m = s.rdy[c2d - 1];
    if (lane == 31) rdy_p = s.rdy[c2d + (ic <= s.I ? 1 : 0)];
    double dux_m = 0., dux_c = 0., dux_p = 0., dvy_m = 0., dvy_c = 0.;
    const bool central_h = (s.method_h == MOHID_CentralDif || s.method_h == MOHID_LeapFrog);
    const bool central_v = (s.method_v == MOHID_CentralDif || s.method_v == MOHID_LeapFrog);
    if (central_h) {
        dux_m = s.DUX[c2d - sj]; dux_c = s.DUX[c2d]; dux_p = s.DUX[c2d + sj];
        dvy_c = s.DVY[c2d]; dvy_m = s.DVY[c2d - 1];
    }
    const bool far_h = (s.method_h == MOHID_UpwindOrder2 || s.method_h == MOHID_UpwindOrder3);
    const bool far_v = (s.method_v == MOHID_UpwindOrder2 || s.method_v == MOHID_UpwindOrder3);

    const unsigned mtop = s.mask[c2d + sk * s.K];
    const bool colwet = (mtop & M_COLWET) != 0;
    const bool colopen = (mtop & M_COLOPEN) != 0;
    const bool bnd = (mtop & M_BND) != 0;
    const int bc = pa.bc;
    const double theta = pa.theta_difv;
    const bool advv_imp = pa.advv_implicit != 0;
    // halo lanes of the strip: lanes 0,1 fetch cell i-2, lane 31 fetches cell i+1
    const bool halo_lane = (lane < 2) || (lane == 31);
    const int halo_off = (lane == 31) ? ((ic <= s.I) ? 1 : 0) : -2;

    // ---- rolling registers along k (cells k-1 .. k+2 of this column) ----
    long q = c2d + sk;                                    // cell (i,j,1)
    double Pm1 = P[c2d], Pc = P[q], Pp1 = P[q + sk];
    unsigned mm1 = s.mask[c2d], mc = s.mask[q], mp1 = s.mask[q + sk];
    double dtv_m = 0., dtv_c = s.dtv[q], dtv_p = s.dtv[q + sk];
    double rdz_c = s.rdz[q], rdz_p = s.rdz[q + sk];
    double qz_c = s.qz[q], qz_p = s.qz[q + sk];
    double Dk = 0., Ek_b = 0., TIk_b = 0.;                // contributions of the bottom face to row k
    double Wprev = 0., Gprev = 0.;
    unsigned long long zp = 0;

    for (int k = 1; k <= s.K; ++k, q += sk) {
        const bool has2 = (k + 2 <= s.K + 1);
        const long q2 = has2 ? q + 2 * sk : q;
        const double Pp2 = has2 ? P[q2] : 0.;
        const unsigned mp2 = has2 ? s.mask[q2] : 0u;
        const double rdz_pp = has2 ? s.rdz[q2] : 0.;
        const double dtv_pp = has2 ? s.dtv[q2] : 0.;
        const double qz_pp = has2 ? s.qz[q2] : 0.;

        const bool open_c = (mc & M_OPEN) != 0;
        // ---------------- VolumeVariation (AD:3966-4021) ----------------
        double TI = open_c ? Pc * s.vr[q] : Pc;
        double E = 1.0;
        if (open_c && k == s.K) E = 1.0 + dtv_c * qz_p;
        double D = Dk, F = 0.;
        E += Ek_b;
        TI += TIk_b;

        // ---------------- horizontal faces (explicit) ----------------
        if (!s.vertical1d) {
            // X direction: west face j and east face j+1 of this cell
            const double Pw2 = P[q - 2 * sj], Pw1 = P[q - sj], Pe1 = P[q + sj];
            const double Pe2 = jp2 ? P[q + 2 * sj] : 0.;
            const unsigned mw2 = s.mask[q - 2 * sj], mw1 = s.mask[q - sj], me1 = s.mask[q + sj];
            const unsigned me2 = jp2 ? s.mask[q + 2 * sj] : 0u;
            const double dtv_w = s.dtv[q - sj], dtv_e = s.dtv[q + sj];
            double dtv_w2 = 0., dtv_e2 = 0.;
            if (far_h) {
                dtv_w2 = s.dtv[q - 2 * sj];
                dtv_e2 = jp2 ? s.dtv[q + 2 * sj] : 0.;
            }
            const double fw = hface_flux(s, (mc & M_CFU) != 0, (mw1 & M_OPEN) != 0, open_c, s.qx[q], s.dhu[q], Pw2,
                                         Pw1, Pc, Pe1, (mw2 & M_OPEN) != 0, (me1 & M_OPEN) != 0, dtv_w2, dtv_w,
                                         dtv_c, dtv_e, rdx_m, rdx_c, rdx_p, dux_m, dux_c);
            const double fe = hface_flux(s, (me1 & M_CFU) != 0, open_c, (me1 & M_OPEN) != 0, s.qx[q + sj],
                                         s.dhu[q + sj], Pw1, Pc, Pe1, Pe2, (mw1 & M_OPEN) != 0,
                                         (me2 & M_OPEN) != 0, dtv_w, dtv_c, dtv_e, dtv_e2, rdx_c, rdx_p, rdx_pp,
                                         dux_c, dux_p);
            TI += (fw - fe) * dtv_c;

            if (!s.xzflow) {
                // Y direction: each lane builds its south face; the north face comes from lane+1
                double Ps1 = shfl_up_d(Pc, 1), Ps2 = shfl_up_d(Pc, 2), Pn1 = shfl_dn_d(Pc, 1);
                unsigned ms1 = __shfl_up_sync(0xffffffffu, mc, 1), ms2 = __shfl_up_sync(0xffffffffu, mc, 2);
                unsigned mn1 = __shfl_down_sync(0xffffffffu, mc, 1);
                double dtv_s = shfl_up_d(dtv_c, 1), dtv_n = shfl_dn_d(dtv_c, 1), dtv_s2 = 0.;
                if (far_h) dtv_s2 = shfl_up_d(dtv_c, 2);
                // strip halo
                double hP = 0., hT = 0.;
                unsigned hM = 0;
                if (halo_lane) { hP = P[q + halo_off]; hM = s.mask[q + halo_off]; }
                if (lane == 0) dtv_s = s.dtv[q - 1];
                if (far_h && halo_lane) hT = s.dtv[q + halo_off];
                const double hP1 = __shfl_sync(0xffffffffu, hP, 1);
                const unsigned hM1 = __shfl_sync(0xffffffffu, hM, 1);
                if (lane == 0) { Ps2 = hP; ms2 = hM; Ps1 = hP1; ms1 = hM1; dtv_s2 = hT; }
                else if (lane == 1) { Ps2 = hP; ms2 = hM; dtv_s2 = hT; }
                else if (lane == 31) { Pn1 = hP; mn1 = hM; dtv_n = far_h ? hT : dtv_n; }
                const double fs = hface_flux(s, (mc & M_CFV) != 0, (ms1 & M_OPEN) != 0, open_c, s.qy[q], s.dhv[q],
                                             Ps2, Ps1, Pc, Pn1, (ms2 & M_OPEN) != 0, (mn1 & M_OPEN) != 0, dtv_s2,
                                             dtv_s, dtv_c, dtv_n, rdy_m, rdy_c, rdy_p, dvy_m, dvy_c);
                const double fn = shfl_dn_d(fs, 1);
                TI += (fs - fn) * dtv_c;
            }
        }

        // ---------------- vertical face k+1 (top of this cell) ----------------
        double Dn = 0., En_b = 0., TIn_b = 0.;            // contributions to row k+1
        if (s.K > 1 && (mp1 & M_CFW)) {
            // diffusion (AD:2708-2775 / 2779-2937)
            const double a = s.dvz[q + sk];
            const double aux1 = a * dtv_c, aux2 = a * dtv_p;
            const double dP = Pp1 - Pc;
            E += aux1 * theta;
            F -= aux1 * theta;
            TI += aux1 * dP * (1. - theta);
            Dn -= aux2 * theta;
            En_b += aux2 * theta;
            TIn_b -= aux2 * dP * (1. - theta);
            // advection (AD:2941-3144); weights exist iff both cells are open (MF:10559)
            if (!s.vertical1d && colopen && open_c && (mp1 & M_OPEN)) {
                double c1, c2, c3, c4, dwz_c = 0., dwz_p = 0.;
                if (central_v) { dwz_c = s.DWZ[q]; dwz_p = s.DWZ[q + sk]; }
                face_weights(s.method_v, s.limiter_v, s.upwind2_v != 0, s.vrelmax, qz_p, Pm1, Pc, Pp1, Pp2,
                             (mm1 & M_OPEN) != 0, (mp2 & M_OPEN) != 0, dtv_m, dtv_c, dtv_p, dtv_pp, rdz_c, rdz_p,
                             rdz_pp, dwz_c, dwz_p, c1, c2, c3, c4);
                if (advv_imp) {
                    const double dfl = qz_p * c2, efl = qz_p * c3;     // D_flux, E_flux (MF:10583-10586)
                    E += dfl * dtv_c;
                    F += efl * dtv_c;
                    Dn -= dfl * dtv_p;
                    En_b -= efl * dtv_p;
                } else {
                    const double fz = qz_p * (c1 * Pm1 + c2 * Pc + c3 * Pp1 + c4 * Pp2);
                    TI -= fz * dtv_c;
                    TIn_b += fz * dtv_p;
                }
            }
        }
        (void)far_v;

        // ---------------- open boundary rows (AD:5369-5672) ----------------
        if (bnd && bc != MOHID_BC_None && open_c) {
            if (bc == MOHID_BC_NullGradient || bc == MOHID_BC_CyclicBoundary) {
                TI = Pc; D = 0.; E = 1.; F = 0.;
            } else if (bc == MOHID_BC_ImposedValue || bc == MOHID_BC_SubModel) {
                const unsigned mN = s.mask[q + 1], mS = s.mask[q - 1], mE = s.mask[q + sj], mW = s.mask[q - sj];
                const double A1 = ((mN & M_OPEN) && !(mN & M_BND)) ? 1. : 0.;
                const double A2 = ((mS & M_OPEN) && !(mS & M_BND)) ? 1. : 0.;
                const double A3 = ((mE & M_OPEN) && !(mE & M_BND)) ? 1. : 0.;
                const double A4 = ((mW & M_OPEN) && !(mW & M_BND)) ? 1. : 0.;
                const double At = A1 + A2 + A3 + A4;
                double ext;
                if (At > 0.) {
                    const double pin_ = (P[q + 1] * A1 + P[q - 1] * A2 + P[q + sj] * A3 + P[q - sj] * A4) / At;
                    ext = pin_ * (1.0 - pa.tdec) + pa.pref[q] * pa.tdec;
                } else {
                    ext = pa.pref[q];
                }
                TI = ext; D = 0.; E = 1.; F = 0.;
            } else {
                // WaterFluxOBoundary (AD:5718-5727)
                const unsigned mE = s.mask[q + sj], mN = s.mask[q + 1];
                const double qb = s.qx[q] * ((mc & M_CFU) ? 1. : 0.) - s.qx[q + sj] * ((mE & M_CFU) ? 1. : 0.) +
                                  s.qy[q] * ((mc & M_CFV) ? 1. : 0.) - s.qy[q + 1] * ((mN & M_CFV) ? 1. : 0.) +
                                  qz_c * ((mc & M_CFW) ? 1. : 0.) - qz_p * ((mp1 & M_CFW) ? 1. : 0.) -
                                  (s.VolumeZ[q] - s.VolumeZOld[q]) / s.dt;
                if (qb < 0.) {
                    if (bc == MOHID_BC_MassConservation) {
                        const double ext = Pc * (1.0 - pa.tdec) + pa.pref[q] * pa.tdec;
                        TI -= qb * ext * dtv_c;
                    } else {                               // MassConservNullGrad: NullGradProp of the old field
                        const int cVn = (mN & M_CFV) ? 1 : 0, cVs = (mc & M_CFV) ? 1 : 0;
                        const int cUe = (mE & M_CFU) ? 1 : 0, cUw = (mc & M_CFU) ? 1 : 0;
                        const int aux = cVn + cVs + cUe + cUw;
                        if (aux > 0)
                            TI = (P[q + 1] * cVn + P[q - 1] * cVs + P[q + sj] * cUe + P[q - sj] * cUw) / (double)aux;
                        else
                            TI = Pc;
                        D = 0.; E = 1.; F = 0.;
                    }
                } else {
                    E += qb * dtv_c;
                }
            }
        }

        // ---------------- land fill (AD:1753) ----------------
        if (mc & M_LAND) TI = NULL_REAL;

        // ---------------- Thomas forward elimination, row k (MF:4087-4099) ----------------
        const double aux = E + D * Wprev;
        if (fabs(aux) > 0.) {
            const double ra = 1.0 / aux;
            Wprev = -F * ra;
            Gprev = (TI - D * Gprev) * ra;
        } else {
            ++zp;                                          // reference leaves W,G stale (MF:4092-4098)
        }
        Wsm[(size_t)(k - 1) * wstride] = Wprev;
        Gsm[(size_t)(k - 1) * wstride] = Gprev;

        // ---------------- roll ----------------
        Dk = Dn; Ek_b = En_b; TIk_b = TIn_b;
        Pm1 = Pc; Pc = Pp1; Pp1 = Pp2;
        mm1 = mc; mc = mp1; mp1 = mp2;
        dtv_m = dtv_c; dtv_c = dtv_p; dtv_p = dtv_pp;
        rdz_c = rdz_p; rdz_p = rdz_pp;
        qz_c = qz_p; qz_p = qz_pp;
    }

    // ---------------- back substitution (MF:4100-4105) ----------------
    if (writer && colwet) {
        double *__restrict__ O = pa.pout;
        long qo = (long)i + sj * j + sk * (s.K + 1);
        double x = 0.0;                                   // RES(KUB+1) = G(KUB+1) = 0 (halo row is the identity)
        O[qo] = x;
        for (int k = s.K; k >= 1; --k) {
            qo -= sk;
            x = Wsm[(size_t)(k - 1) * wstride] * x + Gsm[(size_t)(k - 1) * wstride];
            O[qo] = x;
        }
        if (zp) atomicAdd(s.zero_pivots, zp);
    }
}

// -------------------------------------------------------------------------------------
// K3a: ImposeNullGradient (AD:1926-1987): boundary cells take the compute-face-weighted mean of
// their (new) neighbours.  One thread per (boundary column, k).  Neighbours reached through a
// compute face are never boundary points themselves (HM:939-942), so the pass is order-free.
// -------------------------------------------------------------------------------------
struct BndArgs {
    int I, J, K, ld, nj, ncols;
    long sk;
    const int *cols;          // packed (i,j) of boundary columns
    const int *kfloor;
    const uint8_t *mask;
    double *prop;             // new field (in place)
    const double *pref;
};

__global__ void adt_nullgrad_kernel(const BndArgs b) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)b.ncols * b.K) return;
    const int c = (int)(t / b.K), k = (int)(t % b.K) + 1;
    const int i = b.cols[2 * c], j = b.cols[2 * c + 1];
    const long q2 = (long)i + (long)b.ld * j;
    int kf = b.kfloor[q2];
    kf = kf < 0 ? -kf : kf;
    if (k < kf) return;
    const long q = q2 + b.sk * k;
    const int cVn = (b.mask[q + 1] & M_CFV) ? 1 : 0, cVs = (b.mask[q] & M_CFV) ? 1 : 0;
    const int cUe = (b.mask[q + b.ld] & M_CFU) ? 1 : 0, cUw = (b.mask[q] & M_CFU) ? 1 : 0;
    const int aux = cVn + cVs + cUe + cUw;
    if (aux > 0)
        b.prop[q] = (b.prop[q + 1] * cVn + b.prop[q - 1] * cVs + b.prop[q + b.ld] * cUe + b.prop[q - b.ld] * cUw) /
                    (double)aux;
}

// K3b: Prop_CyclicBoundary (AD:2121-2224), phase 0: boundary cells <- ReferenceProp;
// phase 1: wrap j (per i), phase 2: wrap i (per j).  Launched as three ordered kernels.
__global__ void adt_cyclic_kernel(const BndArgs b, int phase) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (phase == 0) {
        if (t >= (long)b.ncols * b.K) return;
        const int c = (int)(t / b.K), k = (int)(t % b.K) + 1;
        const long q = (long)b.cols[2 * c] + (long)b.ld * b.cols[2 * c + 1] + b.sk * k;
        b.prop[q] = b.pref[q];
    } else if (phase == 1) {
        if (t >= (long)(b.I - 2) * b.K) return;
        const int i = (int)(t / b.K) + 2, k = (int)(t % b.K) + 1;
        const long r1 = (long)i + (long)b.ld * 1, rJ = (long)i + (long)b.ld * b.J;
        if ((b.mask[r1 + b.sk * b.K] & M_BND) && (b.mask[rJ + b.sk * b.K] & M_BND)) {
            if (k >= b.kfloor[rJ - b.ld]) b.prop[r1 + b.sk * k] = b.prop[rJ - b.ld + b.sk * k];
            if (k >= b.kfloor[r1 + b.ld]) b.prop[rJ + b.sk * k] = b.prop[r1 + b.ld + b.sk * k];
        }
    } else {
        if (t >= (long)(b.J - 2) * b.K) return;
        const int j = (int)(t / b.K) + 2, k = (int)(t % b.K) + 1;
        const long r1 = 1 + (long)b.ld * j, rI = (long)b.I + (long)b.ld * j;
        if ((b.mask[r1 + b.sk * b.K] & M_BND) && (b.mask[rI + b.sk * b.K] & M_BND)) {
            if (k >= b.kfloor[rI - 1]) b.prop[r1 + b.sk * k] = b.prop[rI - 1 + b.sk * k];
            if (k >= b.kfloor[r1 + 1]) b.prop[rI + b.sk * k] = b.prop[r1 + 1 + b.sk * k];
        }
    }
}

// -------------------------------------------------------------------------------------
// K4: gather / scatter `width` j-columns of nprop properties to / from a contiguous buffer
// laid out [n][k][w][i] (i fastest).  Coalesced on both sides.
// -------------------------------------------------------------------------------------
struct PackArgs {
    int ld, nj, nk, nprop, j0, width;
    long sk;
    double *prop[NPMAX];
};
__global__ void adt_pack_columns_kernel(const PackArgs a, double *buf, int unpack) {
    const long per = (long)a.nk * a.width * a.ld;
    const long tot = per * a.nprop;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long)gridDim.x * blockDim.x) {
        const int n = (int)(t / per);
        const long r = t % per;
        const int i = (int)(r % a.ld);
        const int w = (int)((r / a.ld) % a.width);
        const int k = (int)(r / ((long)a.ld * a.width));
        const long q = (long)i + (long)a.ld * (a.j0 + w) + a.sk * k;
        if (unpack) a.prop[n][q] = buf[t];
        else buf[t] = a.prop[n][q];
    }
}

}  // namespace adt
