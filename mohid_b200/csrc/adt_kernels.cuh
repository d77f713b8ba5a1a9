// =====================================================================================
//  adt_kernels.cuh -- sm_100a kernels of the batched MOHID property transport step.
//
//  Reference semantics (paths relative to /root/reference/Software):
//     AD = MOHIDBase2/ModuleAdvectionDiffusion.F90, MF = MOHIDBase1/ModuleFunctions.F90
//
//  K1  adt_coef_kernel      per-step shared coefficients (property independent):
//                           Convert_Dif_Vertical / Convert_Visc_Dif_Horizontal (AD:2364-2675),
//                           Compute_DifH/DifV_Constants (AD:1514-1619), DT/V, Vold/V, the packed
//                           32-bit neighbourhood mask and 1/(DWZ(k)+DWZ(k-1)).
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
#include <cuda_pipeline.h>

#include "../../include/mohid_adt.h"

namespace adt {

struct __align__(32) Pack4 { double a, b, c, d; };   // one 256-bit load (adt_lean_kernel.cuh)

constexpr int NPMAX = 32;                 // max properties per launch (kernel-parameter arrays)
constexpr double NULL_REAL = MOHID_NULL_REAL;
constexpr double MIN_VALUE = 1.e-16;      // MGD:1812

// Packed per-cell 32-bit neighbourhood mask written by K1: everything K2 has to know about the cell and
// the compute-point state of its stencil neighbours, so one load replaces six mask loads.
enum : unsigned {
    M_OPEN = 1u << 0,      // OpenPoints3D == 1
    M_CFU = 1u << 1,       // ComputeFacesU3D(i,j,k)   == 1 (west face)
    M_CFUE = 1u << 2,      // ComputeFacesU3D(i,j+1,k) == 1 (east face)
    M_CFV = 1u << 3,       // ComputeFacesV3D(i,j,k)   == 1 (south face)
    M_CFVN = 1u << 4,      // ComputeFacesV3D(i+1,j,k) == 1 (north face)
    M_CFW = 1u << 5,       // ComputeFacesW3D(i,j,k)   == 1 (bottom face)
    M_CFWT = 1u << 6,      // ComputeFacesW3D(i,j,k+1) == 1 (top face)
    M_LAND = 1u << 7,      // LandPoints3D == 1
    M_BND = 1u << 8,       // BoundaryPoints2D(i,j) == 1
    M_COLWET = 1u << 9,    // WaterPoints3D(i,j,KUB) == 1: the column is solved (MF:4086)
    M_COLOPEN = 1u << 10,  // OpenPoints3D(i,j,KUB) == 1: vertical advection coefficients are built (AD:2966)
    M_O_JM2 = 1u << 11, M_O_JM1 = 1u << 12, M_O_JP1 = 1u << 13, M_O_JP2 = 1u << 14,   // OpenPoints3D of j-2..j+2
    M_O_IM2 = 1u << 15, M_O_IM1 = 1u << 16, M_O_IP1 = 1u << 17, M_O_IP2 = 1u << 18,   // ... of i-2..i+2
    M_O_KM1 = 1u << 19, M_O_KP1 = 1u << 20, M_O_KP2 = 1u << 21,                       // ... of k-1, k+1, k+2
    // interior (open, non-boundary) neighbours for the ImposedValue boundary (AD:5462-5470)
    M_A_IP1 = 1u << 22, M_A_IM1 = 1u << 23, M_A_JP1 = 1u << 24, M_A_JM1 = 1u << 25,
    M_DISCH = 1u << 26     // a point discharge feeds this cell (set by adt_discharge_prep_kernel after K1)
};

// NoFluxU/V/W cell lists (AD:1146) live in their own byte per cell, written by K1 only when the caller supplied the
// arrays and read only by the DISCH kernel variants.  A property honours the bits selected by its PropArgs::nfsel:
//   NF_U*: NoAdvFlux zeroes the XX coefficients of the face in the XX pass (AD:4434-4443).
//   NF_V*: the YY pass zeroes the XX coefficient arrays again, at the NoFluxV cells (copy-paste quirk A.4-7,
//          AD:4804-4813) -- after their last use for that property.  They only act on a LATER property of the same
//          time step whose coefficients are not rebuilt (Set_Internal_State, AD:5768-5785: non-TVD methods).
//   NF_WT: NoFluxW zeroes the vertical coefficients of the face (AD:3004-3013).
enum : unsigned {
    NF_UW = 1u, NF_UE = 2u,     // NoFluxU of this cell (west face) / of cell j+1 (east face)
    NF_VW = 4u, NF_VE = 8u,     // NoFluxV of this cell / of cell j+1, acting on the same U faces
    NF_WT = 16u,                // NoFluxW of cell k+1 (top face)
    NF_WEST = NF_UW | NF_VW, NF_EAST = NF_UE | NF_VE
};

struct CoefArgs {
    int ni, nj, nk, ld;              // allocated extents (I+2, J+2, K+2) and leading dimension
    int sj, sk;                      // element strides of j and k in the 3-D device arrays
    int I, J, K;
    double dt, schmidt_h, schmidt_coef_v, schmidt_bg_v;
    int nulldif;
    // raw interface arrays (device copies)
    const double *Wflux_X, *Wflux_Y, *Wflux_Z, *VolumeZOld, *VolumeZ, *Visc_H, *Diff_V, *DWZ, *DZZ, *AreaU, *AreaV;
    const int *Open, *Land, *Water, *CFU, *CFV, *CFW, *SmallDepths;
    const int *NoFluxU, *NoFluxV, *NoFluxW;   // optional (nullptr): NoAdvFlux / NoDifFlux cell lists
    unsigned char *nfmask;                    // NF_* bits per cell (with the cell lists only)
    int nulldif_v;                            // NullDif as it applies to DifZ (nulldif: to DifX / DifY), see PropEff
    int nodif_h, nodif_w;                     // NoDifFlux on DifX / DifY (AD:2497-2501, 2524-2528) and on AuxK (AD:2737-2741)
    const double *DUX, *DVY, *DZX, *DZY;
    const int *Bnd;
    // outputs
    double *dtv, *vr, *dhu, *dhv, *dvz, *rdz;
    uint32_t *mask;
    int do_geom, do_diff;            // which parts to (re)build
};

// -------------------------------------------------------------------------------------
// K1: one thread per allocated cell; grid = (ceil(ld/128), nk, nj) so no index divisions are needed,
// i fastest (coalesced).  Memory bound: 112 B in, 52 B out per cell.
// -------------------------------------------------------------------------------------
// EDGE = false: the cell is at least two cells away from every side of the allocation, so all neighbour probes are
// in range and their offsets are constants (immediate-offset loads, no clamping logic); EDGE = true: the general form.
template <bool EDGE>
__device__ __forceinline__ void adt_coef_cell(const CoefArgs &a, const int i, const int j, const int k) {
    const int sj = a.sj, sk = a.sk, sj2 = a.ld;          // 3-D strides; 2-D arrays are (i + ld*j)
    const int q2 = i + sj2 * j;
    const int q = i + sj * j + sk * k;
    // ---- neighbour offsets, clamped to the allocation (a clamped probe is masked out below) ----
    const bool im1 = !EDGE || i >= 1, im2 = !EDGE || i >= 2, ip1 = !EDGE || i + 1 < a.ni, ip2 = !EDGE || i + 2 < a.ni;
    const bool jm1 = !EDGE || j >= 1, jm2 = !EDGE || j >= 2, jp1 = !EDGE || j + 1 < a.nj, jp2 = !EDGE || j + 2 < a.nj;
    const bool km1 = !EDGE || k >= 1, kp1 = !EDGE || k + 1 < a.nk, kp2 = !EDGE || k + 2 < a.nk;
    const int oim1 = im1 ? -1 : 0, oim2 = im2 ? -2 : 0, oip1 = ip1 ? 1 : 0, oip2 = ip2 ? 2 : 0;
    const int ojm1 = jm1 ? -sj : 0, ojm2 = jm2 ? -2 * sj : 0, ojp1 = jp1 ? sj : 0, ojp2 = jp2 ? 2 * sj : 0;
    const int okm1 = km1 ? -sk : 0, okp1 = kp1 ? sk : 0, okp2 = kp2 ? 2 * sk : 0;
    const int qtop = i + sj * j + sk * a.K;

    // ---- all loads up front (independent, so they overlap) ----
    const int open_c = a.Open[q], cfu_c = a.CFU[q], cfv_c = a.CFV[q], cfw_c = a.CFW[q], land_c = a.Land[q];
    const int bnd_c = a.Bnd[q2];
    int cfu_e = 0, cfv_n = 0, cfw_t = 0, wat_top = 0, open_top = 0;
    int o_jm2 = 0, o_jm1 = 0, o_jp1 = 0, o_jp2 = 0, o_im2 = 0, o_im1 = 0, o_ip1 = 0, o_ip2 = 0, o_km1 = 0, o_kp1 = 0, o_kp2 = 0;
    double V = 0., Vold = 0., dwz = 0., dwz_m = 0.;
    if (a.do_geom) {
        cfu_e = a.CFU[q + ojp1]; cfv_n = a.CFV[q + oip1]; cfw_t = a.CFW[q + okp1];
        wat_top = a.Water[qtop]; open_top = a.Open[qtop];
        o_jm2 = a.Open[q + ojm2]; o_jm1 = a.Open[q + ojm1]; o_jp1 = a.Open[q + ojp1]; o_jp2 = a.Open[q + ojp2];
        o_im2 = a.Open[q + oim2]; o_im1 = a.Open[q + oim1]; o_ip1 = a.Open[q + oip1]; o_ip2 = a.Open[q + oip2];
        o_km1 = a.Open[q + okm1]; o_kp1 = a.Open[q + okp1]; o_kp2 = a.Open[q + okp2];
        V = a.VolumeZ[q]; Vold = a.VolumeZOld[q]; dwz = a.DWZ[q]; dwz_m = a.DWZ[q + okm1];
    }
    double visc = 0., visc_w = 0., visc_s = 0., areau = 0., areav = 0., diffv = 0., dzz_m = 0.;
    double dux = 0., dux_w = 0., dvy = 0., dvy_s = 0., dzx_w = 0., dzy_s = 0., wx = 1., wy = 1., wz = 1.;
    int small = 0;
    if (a.do_diff) {
        visc = a.Visc_H[q]; visc_w = a.Visc_H[q + ojm1]; visc_s = a.Visc_H[q + oim1];
        areau = a.AreaU[q]; areav = a.AreaV[q]; diffv = a.Diff_V[q]; dzz_m = a.DZZ[q + okm1];
        dux = a.DUX[q2]; dux_w = a.DUX[q2 - (jm1 ? sj2 : 0)]; dvy = a.DVY[q2]; dvy_s = a.DVY[q2 + oim1];
        dzx_w = a.DZX[q2 - (jm1 ? sj2 : 0)]; dzy_s = a.DZY[q2 + oim1];
        if (a.nulldif || a.nulldif_v) { wx = a.Wflux_X[q]; wy = a.Wflux_Y[q]; wz = a.Wflux_Z[q]; }
        if (a.SmallDepths) small = a.SmallDepths[q2];
    }

    const bool cfu = cfu_c == 1, cfv = cfv_c == 1, cfw = cfw_c == 1;
    if (a.do_geom) {
        const bool bnd = bnd_c == 1;
        const bool ojm1b = jm1 && o_jm1 == 1, ojp1b = jp1 && o_jp1 == 1, oim1b = im1 && o_im1 == 1, oip1b = ip1 && o_ip1 == 1;
        unsigned m = 0;
        if (open_c == 1) m |= M_OPEN;
        if (cfu) m |= M_CFU;
        if (cfv) m |= M_CFV;
        if (cfw) m |= M_CFW;
        if (jp1 && cfu_e == 1) m |= M_CFUE;
        if (ip1 && cfv_n == 1) m |= M_CFVN;
        if (kp1 && cfw_t == 1) m |= M_CFWT;
        if (land_c == 1) m |= M_LAND;
        if (bnd) m |= M_BND;
        if (wat_top == 1) m |= M_COLWET;
        if (open_top == 1) m |= M_COLOPEN;
        if (jm2 && o_jm2 == 1) m |= M_O_JM2;
        if (ojm1b) m |= M_O_JM1;
        if (ojp1b) m |= M_O_JP1;
        if (jp2 && o_jp2 == 1) m |= M_O_JP2;
        if (im2 && o_im2 == 1) m |= M_O_IM2;
        if (oim1b) m |= M_O_IM1;
        if (oip1b) m |= M_O_IP1;
        if (ip2 && o_ip2 == 1) m |= M_O_IP2;
        if (km1 && o_km1 == 1) m |= M_O_KM1;
        if (kp1 && o_kp1 == 1) m |= M_O_KP1;
        if (kp2 && o_kp2 == 1) m |= M_O_KP2;
        if (bnd) {                                        // interior neighbours: only boundary rows need them (rare)
            if (oip1b && a.Bnd[q2 + 1] != 1) m |= M_A_IP1;
            if (oim1b && a.Bnd[q2 - 1] != 1) m |= M_A_IM1;
            if (ojp1b && a.Bnd[q2 + sj2] != 1) m |= M_A_JP1;
            if (ojm1b && a.Bnd[q2 - sj2] != 1) m |= M_A_JM1;
        }
        if (a.NoFluxU) {
            unsigned nf = 0;
            if (a.NoFluxU[q] == 1) nf |= NF_UW;
            if (a.NoFluxV[q] == 1) nf |= NF_VW;
            if (jp1 && a.NoFluxU[q + ojp1] == 1) nf |= NF_UE;
            if (jp1 && a.NoFluxV[q + ojp1] == 1) nf |= NF_VE;
            if (kp1 && a.NoFluxW[q + okp1] == 1) nf |= NF_WT;
            a.nfmask[q] = (unsigned char)nf;
        }
        a.mask[q] = m;
        const bool inwork = (i >= 1 && i <= a.I && j >= 1 && j <= a.J && k >= 1 && k <= a.K);
        a.dtv[q] = (inwork && V != 0.) ? a.dt / V : 0.;
        a.vr[q] = (inwork && V != 0.) ? Vold / V : 1.;
        const double sd = km1 ? (dwz + dwz_m) : 0.;
        a.rdz[q] = (sd != 0.) ? 1.0 / sd : 0.;
    }
    if (a.do_diff) {
        double hu = 0., hv = 0., vz = 0.;
        if (cfu && jm1) {
            // DifX (AD:2486-2495) then Diff_H_Const_U (AD:1549-1553), same operation order
            double difx = a.schmidt_h * (visc * dux_w + visc_w * dux) / (dux + dux_w);
            if (a.nulldif && wx == 0.) difx = 0.;
            if (a.nodif_h && a.NoFluxU && a.NoFluxU[q] == 1) difx = 0.;
            hu = difx * areau / dzx_w;
        }
        if (cfv && im1) {
            double dify = a.schmidt_h * (visc * dvy_s + visc_s * dvy) / (dvy + dvy_s);
            if (a.nulldif && wy == 0.) dify = 0.;
            if (a.nodif_h && a.NoFluxV && a.NoFluxV[q] == 1) dify = 0.;
            hv = dify * areav / dzy_s;
        }
        if (cfw && km1 && small == 0) {
            // DifZ (AD:2397-2405) then Diff_V_Const (AD:1591-1597)
            double difz = (a.schmidt_coef_v * diffv + a.schmidt_bg_v);
            if (a.nulldif_v && wz == 0.) difz = 0.;
            const double auxk = difz * dux * dvy;
            vz = auxk / dzz_m;
            if (a.nodif_w && a.NoFluxW && a.NoFluxW[q] == 1) vz = 0.;
        }
        a.dhu[q] = hu; a.dhv[q] = hv; a.dvz[q] = vz;
    }
}

__global__ void __launch_bounds__(128) adt_coef_kernel(const CoefArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // blocks run i fastest, then k, then j: the k+-1/2 and j+-1/2 neighbour reads of a cell are then issued within a
    // few MB of traffic of the cell itself and hit L2 (with k slowest, a whole plane of all 17 inputs lies in between)
    const int k = blockIdx.y, j = blockIdx.z;
    if (i >= a.ld) return;
    if (i >= a.ni) {                                  // leading-dimension padding
        const int q = i + a.sj * j + a.sk * k;
        if (a.do_geom) { a.mask[q] = 0; a.dtv[q] = 0.; a.vr[q] = 1.; a.rdz[q] = 0.; }
        if (a.do_diff) { a.dhu[q] = 0.; a.dhv[q] = 0.; a.dvz[q] = 0.; }
        return;
    }
    const bool interior = i >= 2 && i + 2 < a.ni && j >= 2 && j + 2 < a.nj && k >= 1 && k + 2 < a.nk;
    if (interior) adt_coef_cell<false>(a, i, j, k);
    else adt_coef_cell<true>(a, i, j, k);
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
// Point discharges (Discharges, AD:4025-4128).  The geometry (Me%Discharge in the caller, WP:14761-14773)
// is shared by all properties; concentrations are per property.  A small per-step kernel expands every
// discharge cell into its k-range, computes the per-layer flow (uniform vertical distribution:
// Flow*DWZ/WaterColumn, AD:4063-4077) and flags the receiving cells in the mask.
// -------------------------------------------------------------------------------------
struct DischArgs {
    int ncell;                 // discharge cells that are not ignored (AD:4039-4041)
    int K, ld, sj, sk;
    const int *ci, *cj, *ck, *ckmin, *ckmax, *cvert, *cbypass;   // per listed cell (vert / bypass of its discharge)
    const double *cflow;
    int *kmin_eff, *kmax_eff;  // out: effective k-range
    double *flow_k;            // out: [ncell][K+2] flow into layer k
    const int *kfloor;
    const double *DWZ;
    uint32_t *mask;
};

__global__ void adt_discharge_prep_kernel(const DischArgs d) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= d.ncell) return;
    const int i = d.ci[n], j = d.cj[n];
    const int q2 = i + d.ld * j, q3 = i + d.sj * j, sk = d.sk;
    int kmin = d.ckmin[n], kmax = d.ckmax[n];
    const bool uniform = d.cvert[n] == MOHID_DischUniform;
    if (uniform) {
        if (kmin == MOHID_FILL_INT) kmin = d.kfloor[q2];
        if (kmax == MOHID_FILL_INT) kmax = d.K;
    } else {
        kmin = d.ck[n]; kmax = d.ck[n];
    }
    d.kmin_eff[n] = kmin; d.kmax_eff[n] = kmax;
    double wc = 0.0;
    for (int k = kmin; k <= kmax; ++k) wc = wc + d.DWZ[q3 + sk * k];
    for (int k = kmin; k <= kmax; ++k) {
        d.flow_k[(size_t)n * (d.K + 2) + k] = uniform ? d.cflow[n] * d.DWZ[q3 + sk * k] / wc : d.cflow[n];
        atomicOr(&d.mask[q3 + sk * k], M_DISCH);
    }
}

struct DischView {             // what K2 needs to apply the discharges of one property
    int ncell, K;
    const int *ci, *cj, *kmin_eff, *kmax_eff, *cbypass;
    const double *flow_k;
};

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
    unsigned nfsel;         // NF_* bits this property honours (0 = none), see PropEff in adt_api.cu
    int pad0;
    double *tih;            // net horizontal flux into each cell: written by adt_hflux_kernel, read by the HSPLIT variants
    const double *dconc;    // DischConc of this property per listed discharge cell, or nullptr (no discharges)
    const double *dconcmf;  // DischConcMF
};

struct StepArgs {
    int I, J, K, ld, nj;
    int sj, sk;                                 // element strides of j and k in the 3-D device arrays
    int nprop, ntile_i;                         // tiles of 31 cells along i
    int j_begin, j_count;                       // columns j_begin .. j_begin+j_count-1 are advanced
    int method_h, limiter_h, method_v, limiter_v, upwind2_h, upwind2_v;
    int vertical1d, xzflow;
    int stage2, twod;                           // see adt_transport_kernel; twod: adt_hsolve_kernel on a 2-D domain (K = 1)
    double vrelmax, dt;
    const double *qx, *qy, *qz, *dtv, *vr, *dhu, *dhv, *dvz, *rdz;
    const uint32_t *mask;
    const double *rdx, *rdy, *DUX, *DVY, *DWZ;
    const double *VolumeZ, *VolumeZOld;         // open-boundary flux only (AD:5718-5727)
    const unsigned char *nfmask;      // NF_* bits (nullptr without NoFlux cell lists)
    unsigned long long *zero_pivots;
    DischView disch;
    PropArgs p[NPMAX];
};

__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_dn_d(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double sel(bool c, double a, double b) { return c ? a : b; }
__device__ __forceinline__ bool all_set(unsigned m, unsigned bits) { return (m & bits) == bits; }
__device__ __forceinline__ double ratio_or_zero(double a, double b) { return b != 0. ? a / b : 0.; }
// min / max as one compare + select (fmin/fmax expand to a NaN-propagation sequence of ~7 instructions;
// operands here are never NaN)
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
// x with the sign of s flipped in when s < 0 (s = -0.0 counts as negative; callers pass differences, where
// a zero difference makes the result irrelevant), and |x| with the sign of s: one LOP3 each on the high word
__device__ __forceinline__ double flip_sign_by(double x, double s) {
    return __hiloint2double(__double2hiint(x) ^ (__double2hiint(s) & 0x80000000), __double2loint(x));
}
__device__ __forceinline__ double with_sign_of(double x, double s) {
    return __hiloint2double((__double2hiint(x) & 0x7fffffff) | (__double2hiint(s) & 0x80000000), __double2loint(x));
}

// 1/x for normal, finite x without the IEEE slow path: MUFU.RCP64H seed (2^-20) + two Newton steps
// (relative error <= ~2 ulp).  Callers guarantee |x| is far from 0, Inf and the denormal range.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// TVD limiter psi(r) (MF:10812-10856); Cr may be overwritten by the PDM branch (quirk A.4-1)
template <int L>
__device__ __forceinline__ double tvd_psi(int limiter_rt, double r, double &Cr) {
    const int limiter = (L > 0) ? L : limiter_rt;
    if (limiter == MOHID_SuperBee) return dmax(dmax(0., dmin(1., 2. * r)), dmin(r, 2.));
    if (limiter == MOHID_MinMod) return dmax(0., dmin(1., r));
    if (limiter == MOHID_VanLeer) return (r < 0.) ? 0. : 2. * r * fast_rcp(1. + r);
    if (limiter == MOHID_Muscl) return dmax(0., dmin(dmin(2., 2. * r), (1. + r) * 0.5));
    // PDM
    const double c = (1. - 2. * fabs(Cr)) / 6.;
    const double a = 0.5 + c, b = 0.5 - c;
    const double aux = a + b * r;
    if (fabs(Cr) < MIN_VALUE) Cr = MIN_VALUE;
    return dmax(0., dmin(dmin(aux, 2. * fast_rcp(1. - Cr)), 2. * r * fast_rcp(Cr)));
}

// Face weights of ComputeAdvectionFace (MF:10702-10894) in UPWIND-ORIENTED form: the face flux of
// property is Q * (wuu*Puu + wu*Pu + wd*Pd) with (uu, u, d) = (2nd upwind, upwind, downwind) cell.
// For Q > 0 that is (f-2, f-1, f), otherwise (f+1, f, f-1) -- Q == 0 counts as negative (MF:11127-11141).
//   near   : the 2nd upwind cell is not a compute point (MF:10563-10570)
//   t_*    : DT/V of the cells (signed Courant = Q*t_u, MF:11045-11053; VolumeRel = max/min of t)
//   rd_u   : 1/(du_u+du_uu), rd_c : 1/(du_u+du_d); du_u, du_d only for the centred schemes
template <int M, int L>
__device__ __forceinline__ void oriented_weights(int method_rt, int limiter_rt, bool upwind2, double vrelmax,
                                                 double Q, double Puu, double Pu, double Pd, bool near,
                                                 double t_uu, double t_u, double t_d, double rd_u, double rd_c,
                                                 double du_u, double du_d, double &wuu, double &wu, double &wd) {
    const int method = (M > 0) ? M : method_rt;
    wuu = 0.; wu = 1.; wd = 0.;
    if (method == MOHID_UpwindOrder1) return;
    if (method == MOHID_P2_TVD) {
        double Cr = Q * t_u;
        double dC = (Pd - Pu) * rd_c;
        dC = (fabs(dC) < MIN_VALUE) ? with_sign_of(MIN_VALUE, dC) : dC;   // MF:10795-10803 (dC is never -0.0)
        const double r = (Pu - Puu) * rd_u * fast_rcp(dC);
        double theta = tvd_psi<L>(limiter_rt, r, Cr);
        theta = 0.5 * theta * (1. - Cr);
        theta = (near && upwind2) ? 0. : theta;
        wu = 1. - theta;
        wd = theta;
        return;
    }
    if (method == MOHID_CentralDif || method == MOHID_LeapFrog) {
        if (!(near && upwind2)) { wu = du_d * rd_c; wd = du_u * rd_c; }
        return;
    }
    // UpwindOrder2 (QUICK) / UpwindOrder3 (QUICKEST), MF:10751-10770, 11055-11125
    const double tmax = dmax(dmax(t_uu, t_u), t_d), tmin = dmin(dmin(t_uu, t_u), t_d);
    const bool first_order = near || (tmax > vrelmax * tmin) || (Q == 0.);
    if (first_order) return;
    if (method == MOHID_UpwindOrder2) {
        wuu = -1. / 8.; wu = 6. / 8.; wd = 3. / 8.;
    } else {
        const double Cr = Q * t_u;
        const double c = (1. - 2. * fabs(Cr)) / 6.;
        const double a = 0.5 + c, b = 0.5 - c, d = (1. - fabs(Cr)) / 2.;
        wuu = -d * b; wu = 1. + d * (b - a); wd = d * a;
    }
}

// Total (advective - diffusive) property flux through a horizontal face between cells 2 and 3 of the
// stencil (1,2,3,4), positive toward cell 3.  adv_on = face is a compute face AND both cells are open
// (MF:10559, AD:4467); the diffusive conductance dh is already zero on non-compute faces (K1).
template <int M, int L>
__device__ __forceinline__ double hface_flux(const StepArgs &s, bool adv_on, double Q, double dh, double P1,
                                             double P2, double P3, double P4, bool o1, bool o4, double t1,
                                             double t2, double t3, double t4, double rd12, double rd23,
                                             double rd34, double du2, double du3) {
    const bool pos = Q > 0.;
    const double Puu = sel(pos, P1, P4), Pu = sel(pos, P2, P3), Pd = sel(pos, P3, P2);
    if constexpr (M == MOHID_P2_TVD && L == MOHID_SuperBee) {
        // Division-free SuperBee: with r = a/b, a = (Pu-Puu)/(du_u+du_uu), b = (Pd-Pu)/(du_u+du_d), the limited
        // increment psi(r)*(Pd-Pu) = sgn(dP) * max(0, min(|dP|, 2 aS), min(aS, 2|dP|)), aS = sgn(dP)*(Pu-Puu)*rho,
        // rho = (du_u+du_d)/(du_u+du_uu) (passed in the rd12 / rd34 slots).  Face flux = Q*(Pu + 0.5(1-Cr)*psi*dP)
        // (MF:10785-10858, 10889-10892); the |dC| < 1e-16 clamp only matters below 1e-13 in the property.
        const double dP = Pd - Pu;
        const double aS = flip_sign_by((Pu - Puu) * sel(pos, rd12, rd34), dP);
        const double ad = fabs(dP);
        double lim = dmax(dmax(0., dmin(ad, aS + aS)), dmin(aS, ad + ad));
        lim = with_sign_of(lim, dP);
        const bool near = pos ? !o1 : !o4;
        lim = (near && s.upwind2_h) ? 0. : lim;
        const double homc = fma(-0.5 * Q, sel(pos, t2, t3), 0.5);          // 0.5*(1 - Cr), Cr = Q*DT/V_upwind
        const double fadv = Q * fma(homc, lim, Pu);
        return sel(adv_on, fadv, 0.) - dh * (P3 - P2);
    }
    double wuu, wu, wd;
    oriented_weights<M, L>(s.method_h, s.limiter_h, s.upwind2_h != 0, s.vrelmax, Q, Puu, Pu, Pd, pos ? !o1 : !o4,
                           sel(pos, t1, t4), sel(pos, t2, t3), sel(pos, t3, t2), sel(pos, rd12, rd34), rd23,
                           sel(pos, du2, du3), sel(pos, du3, du2), wuu, wu, wd);
    const double fadv = Q * (wuu * Puu + wu * Pu + wd * Pd);
    return sel(adv_on, fadv, 0.) - dh * (P3 - P2);
}

// Discharges into cell (i,j,k) (AD:4079-4121), in list order like the reference's serial loop.
__device__ __noinline__ void apply_discharges(const DischView &d, const double *dconc, const double *dconcmf, int i,
                                              int j, int k, bool open_c, double Pc, double vr, double dtv_c, double &TI,
                                              double &E) {
    for (int n = 0; n < d.ncell; ++n) {
        if (d.ci[n] != i || d.cj[n] != j || k < d.kmin_eff[n] || k > d.kmax_eff[n]) continue;
        const double flow = d.flow_k[(size_t)n * (d.K + 2) + k];
        if (open_c) {
            if (d.cbypass[n] || flow > 0.) TI = TI + flow * dtv_c * dconc[n];
            else E = E - flow * dtv_c * dconcmf[n];
        } else if (flow > 0.) {
            // water point that is not an open point: only the discharge changes its volume (AD:4108-4119)
            TI = Pc * vr + flow * dtv_c * dconc[n];
        }
    }
}

// Rows of open-boundary cells (AD:5369-5672); rare (boundary ring only): a predicated-off branch elsewhere.
struct Row { double D, E, F, TI; };
// Orlanski (bc 6, AD:5504-5570 + MF:4129-4490): the reference overwrites the "old" field with the current one right
// before the radiation routine looks at it (AD:5428-5429), so the wave celerity it derives is zero and only the boundary
// flow FlowVelX = Q_b DT/V remains: outflow -> implicit upstream extrapolation of the exterior value from three
// interior cells, inflow -> relaxation to the reference value, both with the default 300-day relaxation time.  The
// exterior (halo) cell of the property receives the new exterior value (written to the output buffer only: the other
// threads of the step still read the old halo).
__device__ __forceinline__ double orlanski_exterior(const StepArgs &s, const PropArgs &pa, int q, int i, int j, double vel) {
    const double *__restrict__ P = pa.pin;
    const bool iedge = (i == 1 || i == s.I);
    const int d = iedge ? 1 : s.sj;                               // stride along the boundary normal
    const bool low = (i == 1 || j == 1);                          // AD:5522: tested before the upper edges
    const int qe = low ? q - d : q + d;                           // "exterior" cell handed to OrlanskiCelerity2D
    const int sg = low ? d : -d;                                  // towards the interior
    int q3 = qe + 3 * sg;
    // corner i = IUB, j = JLB: the lower-edge rule wins, the "exterior" cell is the interior cell IUB-1 and the third
    // point is element IUB+2 of the row = element 0 of row j+1 in the reference's unpadded array
    const bool corner = iedge && i == s.I && j == 1 && s.I > 1;
    if (corner) q3 = q - i + s.sj;
    const double I1 = P[qe + sg], I2 = P[qe + 2 * sg], I3 = P[q3];
    const double ref = pa.pref[qe], old = P[qe];
    const double trelax = 86400 * 300;
    double wc = 0., adj = ref;
    if (vel > 0) { wc = 4 * vel; adj = 0.0546875 * I3 - 0.2578125 * I2 + 0.6015625 * I1; }
    const double auxint = wc * adj;
    const double auxbound = 1 + wc * (1 - 0.6015625) + s.dt / trelax;
    const double ext = (old + auxint + ref * s.dt / trelax) / auxbound;
    if (!corner) pa.pout[qe] = ext;
    return ext;
}

// RARE: compiled with the Orlanski branch (only the DISCH kernel variants, which the host selects for it)
template <bool RARE = false, class Args = StepArgs>
__device__ __forceinline__ void open_boundary_row(const Args &s, const PropArgs &pa, int q, unsigned m, double Pc,
                                               double qz_c, double qz_p, double dtv_c, Row &row, int i = 0, int j = 0,
                                               bool writer = true) {
    const double *__restrict__ P = pa.pin;
    const int sj = s.sj;
    const int bc = pa.bc;
    if (bc == MOHID_BC_NullGradient || bc == MOHID_BC_CyclicBoundary) {
        row.TI = Pc; row.D = 0.; row.E = 1.; row.F = 0.;
    } else if (bc == MOHID_BC_ImposedValue || bc == MOHID_BC_SubModel) {
        const double A1 = (m & M_A_IP1) ? 1. : 0., A2 = (m & M_A_IM1) ? 1. : 0.;
        const double A3 = (m & M_A_JP1) ? 1. : 0., A4 = (m & M_A_JM1) ? 1. : 0.;
        const double At = A1 + A2 + A3 + A4;
        double ext;
        if (At > 0.) {
            const double pin_ = (P[q + 1] * A1 + P[q - 1] * A2 + P[q + sj] * A3 + P[q - sj] * A4) / At;
            ext = pin_ * (1.0 - pa.tdec) + pa.pref[q] * pa.tdec;
        } else {
            ext = pa.pref[q];
        }
        row.TI = ext; row.D = 0.; row.E = 1.; row.F = 0.;
    } else {
        // WaterFluxOBoundary (AD:5718-5727)
        const double qb = s.qx[q] * ((m & M_CFU) ? 1. : 0.) - s.qx[q + sj] * ((m & M_CFUE) ? 1. : 0.) +
                          s.qy[q] * ((m & M_CFV) ? 1. : 0.) - s.qy[q + 1] * ((m & M_CFVN) ? 1. : 0.) +
                          qz_c * ((m & M_CFW) ? 1. : 0.) - qz_p * ((m & M_CFWT) ? 1. : 0.) -
                          (s.VolumeZ[q] - s.VolumeZOld[q]) / s.dt;
        double ext_orl = 0.;
        if constexpr (RARE) { if (bc == MOHID_BC_Orlanski && writer) ext_orl = orlanski_exterior(s, pa, q, i, j, qb * dtv_c); }
        if (qb < 0.) {
            if (bc == MOHID_BC_MassConservation || (RARE && bc == MOHID_BC_Orlanski)) {
                const double ext = (RARE && bc == MOHID_BC_Orlanski) ? ext_orl : Pc * (1.0 - pa.tdec) + pa.pref[q] * pa.tdec;
                row.TI -= qb * ext * dtv_c;
            } else {                               // MassConservNullGrad: NullGradProp of the old field
                const int cVn = (m & M_CFVN) ? 1 : 0, cVs = (m & M_CFV) ? 1 : 0;
                const int cUe = (m & M_CFUE) ? 1 : 0, cUw = (m & M_CFU) ? 1 : 0;
                const int aux = cVn + cVs + cUe + cUw;
                if (aux > 0)
                    row.TI = (P[q + 1] * cVn + P[q - 1] * cVs + P[q + sj] * cUe + P[q - sj] * cUw) / (double)aux;
                else
                    row.TI = Pc;
                row.D = 0.; row.E = 1.; row.F = 0.;
            }
        } else {
            row.E += qb * dtv_c;
        }
    }
}

// Horizontal data of one level of one column, fetched one level ahead of its use (software pipeline).
struct Level {
    double Pw2, Pw1, Pe1, Pe2, hP;      // property at j-2, j-1, j+1, j+2 and the strip-halo value
    double t_w, t_e, t_h;               // DT/V at j-1, j+1 and of the south neighbour of lane 0
    double t_w2, t_e2, t_h2;            // DT/V at j-2, j+2, strip halo (QUICK / QUICKEST only)
    double qxw, qxe, qys, dhw, dhe, dhs, vr;
    double tih;                         // HSPLIT: net horizontal flux into the cell, from adt_hflux_kernel
    unsigned m;
};

// -------------------------------------------------------------------------------------
// K2: fused transport step.
//   warp  <-> (31-cell strip along i, column j, property n);  lane <-> cell i0+lane; lane 31 only
//   supplies the north face of lane 30.  The thread marches k = 1..K building row k of the
//   tridiagonal system in registers, eliminating it on the fly (W,G -> shared memory), then
//   back-substitutes and writes the new property.  Units are ordered property-fastest so the
//   warps of a block read the same shared coefficients (L1 hits).
//   blockDim.x = 32 * WPB; dynamic shared memory = 2 * K * WPB * 32 doubles.
//   MH/LH/MV/LV > 0 fix the advection method / limiter at compile time; 0 = read from StepArgs.
//   DISCH = the batch has point discharges or NoAdvFlux cell lists (keeps the rare out-of-line call and the extra
//           face predicates out of the common kernels).
//   PF    = 0 no look-ahead, 1 next level fetched into registers, 2 next two levels staged in shared memory
//           with cp.async (in-flight loads hold no registers; needs FULL and a non-QUICK scheme).
//   GGLOB = G of the column solve is parked in the output array instead of shared memory.
//   HSPLIT = the explicit horizontal terms were computed by adt_hflux_kernel (every face once, high occupancy) and
//           arrive as one increment per cell; this kernel keeps the column part (needs PF <= 1).
//   FULL  = 3-D run with both horizontal directions and implicit vertical advection for every property:
//           the level body becomes one basic block (no uniform branches), which lets ptxas interleave the faces.
// -------------------------------------------------------------------------------------
template <int MH, int LH, int MV, int LV, bool DISCH, bool FULL, int WARPS = 8, int PF = 1, bool GGLOB = false, bool HSPLIT = false, int PFD = 0>
__global__ void __launch_bounds__(WARPS * 32, 1) adt_transport_kernel(const __grid_constant__ StepArgs s) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, WPB = blockDim.x >> 5;
    double *__restrict__ Wsm = smem + warp * 32 + lane;                  // [K][WPB][32]
    double *__restrict__ Gsm = Wsm + (size_t)s.K * WPB * 32;
    const int wstride = WPB * 32;
    constexpr int NV = 16, NSTAGE = 2;                       // PF == 2: staged values per level, ring depth
    double *__restrict__ Stg = smem + (size_t)s.K * WPB * 32 * (GGLOB ? 1 : 2) + warp * 32 + lane;
    const long nunits = (long)s.nprop * s.ntile_i * s.j_count;
    const long unit = (long)blockIdx.x * WPB + warp;
    if (unit >= nunits) return;
    const int n = (int)(unit % s.nprop);
    const int tile = (int)((unit / s.nprop) % s.ntile_i);
    const int j = (int)(unit / ((long)s.nprop * s.ntile_i)) + s.j_begin;
    const int i = 1 + tile * 31 + lane;
    const bool writer = (lane < 31) && (i <= s.I);
    const int ic = min(i, s.I + 1);                       // clamped column: every load stays in bounds
    const PropArgs pa = s.p[n];
    const double *__restrict__ P = pa.pin;
    const int sj = s.sj, sk = s.sk, sj2 = s.ld;          // 32-bit cell indices (n3 < 2^31 is checked at create)
    const int c2 = ic + sj2 * j;                          // column in the 2-D arrays
    const int c2d = ic + sj * j;                          // column base (k = 0) in the 3-D arrays
    const int je2 = (j + 2 <= s.J + 1) ? 2 * sj : sj;
    const int je2_2 = (j + 2 <= s.J + 1) ? 2 * sj2 : sj2;
    const int jw2 = (j >= 2) ? 2 * sj : sj;               // offset of column j-2 (clamped: column -1 does not exist)

    const int method_h = MH > 0 ? MH : s.method_h, method_v = MV > 0 ? MV : s.method_v;
    const bool central_h = (method_h == MOHID_CentralDif || method_h == MOHID_LeapFrog);
    const bool central_v = (method_v == MOHID_CentralDif || method_v == MOHID_LeapFrog);
    const bool far_h = (method_h == MOHID_UpwindOrder2 || method_h == MOHID_UpwindOrder3);
    const bool far_v = (method_v == MOHID_UpwindOrder2 || method_v == MOHID_UpwindOrder3);
    // stage2: second half of a horizontally implicit step (adt_hsolve_kernel.cuh): the row restarts from the
    // intermediate field (TI = PROP, E = 1, AD:4250-4253) and only the vertical terms and the boundary rows remain
    const bool stage2 = !FULL && s.stage2 != 0;
    const bool do_h = !HSPLIT && (FULL || (!s.vertical1d && !stage2)), do_y = !HSPLIT && (FULL || (do_h && !s.xzflow));
    const bool do_vadv = FULL || !s.vertical1d;

    // ---- 2-D metrics of the column ----
    const double rdx_m = s.rdx[c2 - sj2], rdx_c = s.rdx[c2], rdx_p = s.rdx[c2 + sj2], rdx_pp = s.rdx[c2 + je2_2];
    const double rdy_c = s.rdy[c2];
    double rdy_m = shfl_up_d(rdy_c, 1), rdy_p = shfl_dn_d(rdy_c, 1);
    if (lane == 0) rdy_m = s.rdy[c2 - 1];
    if (lane == 31) rdy_p = s.rdy[c2 + (ic <= s.I ? 1 : 0)];
    constexpr bool FAST_H = (MH == MOHID_P2_TVD && LH == MOHID_SuperBee);
    // fast path operands: rho = (du_u+du_d)/(du_u+du_uu) for both flow directions of the west, east and south face
    const double rho_wp = FAST_H ? ratio_or_zero(rdx_m, rdx_c) : rdx_m, rho_wn = FAST_H ? ratio_or_zero(rdx_p, rdx_c) : rdx_p;
    const double rho_ep = FAST_H ? ratio_or_zero(rdx_c, rdx_p) : rdx_c, rho_en = FAST_H ? ratio_or_zero(rdx_pp, rdx_p) : rdx_pp;
    const double rho_sp = FAST_H ? ratio_or_zero(rdy_m, rdy_c) : rdy_m, rho_sn = FAST_H ? ratio_or_zero(rdy_p, rdy_c) : rdy_p;
    double dux_m = 0., dux_c = 0., dux_p = 0., dvy_m = 0., dvy_c = 0.;
    if (central_h) {
        dux_m = s.DUX[c2 - sj2]; dux_c = s.DUX[c2]; dux_p = s.DUX[c2 + sj2];
        dvy_c = s.DVY[c2]; dvy_m = s.DVY[c2 - 1];
    }

    const unsigned mtop = s.mask[c2d + sk * s.K];
    const bool colwet = (mtop & M_COLWET) != 0;
    const bool colopen = (mtop & M_COLOPEN) != 0;
    const bool obc = (mtop & M_BND) != 0 && pa.bc != MOHID_BC_None;
    // vertical advection acts on a top face iff both cells are open, the face is a compute face, the column's
    // surface cell is open and the run is not Vertical1D (AD:2966, 3041; MF:10559); bit 31 is never set in a mask
    const unsigned top_req = (do_vadv && colopen) ? (M_OPEN | M_O_KP1 | M_CFWT) : (1u << 31);
    // NoAdvFlux: a selected NF_* bit switches the advective part of that face off
    const unsigned nfsel = DISCH ? pa.nfsel : 0u;
    const double theta = pa.theta_difv, omt = 1. - pa.theta_difv;
    const bool advv_imp = FULL || pa.advv_implicit != 0;
    // halo lanes of the strip: lanes 0,1 fetch cell i-2, lane 31 fetches cell i+1
    const bool halo_lane = (lane < 2) || (lane == 31);
    const int halo_off = (lane == 31) ? ((ic <= s.I) ? 1 : 0) : -2;

    // fetch the horizontal data of level k at cell offset q (all loads independent of computed values)
    auto fetch = [&](int q, Level &L) {
        L.m = __ldg(s.mask + (q));
        L.vr = __ldg(s.vr + (q));
        if (HSPLIT) L.tih = __ldg(pa.tih + (q));
        if (do_h) {
            L.Pw2 = __ldg(P + (q - jw2)); L.Pw1 = __ldg(P + (q - sj)); L.Pe1 = __ldg(P + (q + sj)); L.Pe2 = __ldg(P + (q + je2));
            L.t_w = __ldg(s.dtv + (q - sj)); L.t_e = __ldg(s.dtv + (q + sj));
            L.qxw = __ldg(s.qx + (q)); L.qxe = __ldg(s.qx + (q + sj)); L.dhw = __ldg(s.dhu + (q)); L.dhe = __ldg(s.dhu + (q + sj));
            if (far_h) { L.t_w2 = __ldg(s.dtv + (q - jw2)); L.t_e2 = __ldg(s.dtv + (q + je2)); }
            if (do_y) {
                L.qys = __ldg(s.qy + (q)); L.dhs = __ldg(s.dhv + (q));
                L.hP = halo_lane ? __ldg(P + (q + halo_off)) : 0.;
                L.t_h = (lane == 0) ? __ldg(s.dtv + (q - 1)) : 0.;
                if (far_h) L.t_h2 = halo_lane ? __ldg(s.dtv + (q + halo_off)) : 0.;
            }
        }
    };

    // PF == 2: the same data, copied global -> shared by cp.async (one private slot per thread and value)
    auto fetch_async = [&](int q, int st) {
        double *d = Stg + (size_t)st * NV * wstride;
        auto cp8 = [&](int slot, const double *g) { __pipeline_memcpy_async(d + slot * wstride, g, 8); };
        cp8(0, P + q - jw2); cp8(1, P + q - sj); cp8(2, P + q + sj); cp8(3, P + q + je2);
        if (halo_lane) cp8(4, P + q + halo_off);
        cp8(5, s.dtv + q - sj); cp8(6, s.dtv + q + sj);
        if (lane == 0) cp8(7, s.dtv + q - 1);
        cp8(8, s.qx + q); cp8(9, s.qx + q + sj); cp8(10, s.qy + q);
        cp8(11, s.dhu + q); cp8(12, s.dhu + q + sj); cp8(13, s.dhv + q); cp8(14, s.vr + q);
        __pipeline_memcpy_async(reinterpret_cast<uint32_t *>(d + 15 * wstride), s.mask + q, 4);
        __pipeline_commit();
    };
    auto load_stage = [&](int st, Level &L) {
        const double *d = Stg + (size_t)st * NV * wstride;
        L.Pw2 = d[0]; L.Pw1 = d[wstride]; L.Pe1 = d[2 * wstride]; L.Pe2 = d[3 * wstride]; L.hP = d[4 * wstride];
        L.t_w = d[5 * wstride]; L.t_e = d[6 * wstride]; L.t_h = d[7 * wstride];
        L.qxw = d[8 * wstride]; L.qxe = d[9 * wstride]; L.qys = d[10 * wstride];
        L.dhw = d[11 * wstride]; L.dhe = d[12 * wstride]; L.dhs = d[13 * wstride]; L.vr = d[14 * wstride];
        L.m = *reinterpret_cast<const uint32_t *>(d + 15 * wstride);
        L.t_w2 = 0.; L.t_e2 = 0.; L.t_h2 = 0.;
    };

    // ---- rolling registers along k (cells k-1 .. k+2 of this column) ----
    int q = c2d + sk;                                     // cell (i,j,1)
    double Pm1 = P[c2d], Pc = P[q], Pp1 = P[q + sk], Pp2 = P[q + ((s.K >= 2) ? 2 * sk : sk)];
    double dtv_m = 0., dtv_c = s.dtv[q], dtv_p = s.dtv[q + sk];
    double rdz_c = s.rdz[q], rdz_p = s.rdz[q + sk], rdz_pp = s.rdz[q + ((s.K >= 2) ? 2 * sk : sk)];
    double qz_c = s.qz[q], qz_p = s.qz[q + sk];
    double dvz_p = s.dvz[q + sk];
    double Dk = 0., Ek_b = 0., TIk_b = 0.;                // contributions of the bottom face to row k
    double Wprev = 0., Gprev = 0.;
    unsigned zp = 0;
    Level lvA, lvB;
    if (PF == 1) fetch(q, lvA);
    if (PF == 2) { fetch_async(q, 0); fetch_async(q + sk, 1); }

    // one level of the march: consumes `cur`, prefetches the next level into `nxt`
    auto level = [&](const int k, const Level &cur_, Level &nxt) {
        // ---- prefetch: level k+1 (horizontal) and level k+2 (vertical rolling values) ----
        // vertical look-ahead: the values of plane k+2 that this level itself uses (stencil cell k+2, its metric sum)
        // were loaded one level ago; here plane k+3 is requested for the next level (clamped to plane K+1)
        const int q2 = (k + 2 <= s.K + 1) ? q + 2 * sk : q + sk;
        const int q3 = (k + 3 <= s.K + 1) ? q + 3 * sk : q2;
        const double Pp3 = __ldg(P + q3), rdz_p3 = __ldg(s.rdz + q3);
        const double dtv_pp = __ldg(s.dtv + q2), qz_pp = __ldg(s.qz + q2), dvz_pp = __ldg(s.dvz + q2);
        if constexpr (PFD > 0) {
            // L2 prefetch, PFD levels ahead, of the streams no earlier block has touched: the east-most property row
            // (every warp) and the shared coefficients of this column and of column j+1 (first property's warp)
            const int kk = min(k + PFD, s.K + 1) - k;
            const int qf = q + kk * sk;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P + (qf + je2)));
            if (n == 0) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(s.qx + (qf + sj)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(s.dhu + (qf + sj)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(s.dtv + (qf + sj)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(s.qy + qf));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(s.dhv + qf));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(s.vr + qf));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(s.mask + qf));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(s.qz + qf));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(s.dvz + qf));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(s.rdz + qf));
            }
        }
        if (PF == 1) fetch(q + sk, nxt);                  // plane K+1 exists, so the look-ahead is always in bounds
        else if (PF == 0) fetch(q, nxt);
        else { __pipeline_wait_prior(NSTAGE - 1); load_stage((k - 1) & 1, nxt); }
        const Level &cur = (PF == 1) ? cur_ : nxt;

        const unsigned m = cur.m;
        const unsigned nf = (DISCH && nfsel) ? (s.nfmask[q] & nfsel) : 0u;
        const bool open_c = (m & M_OPEN) != 0;
        // ---------------- VolumeVariation (AD:3966-4021) ----------------
        Row row;
        double ti0 = sel(open_c && !stage2, Pc * cur.vr, Pc);
        double e0 = sel(open_c && k == s.K && !stage2, 1.0 + dtv_c * qz_p, 1.0);
        // ---------------- Discharges (AD:4025-4128); rare ----------------
        if (DISCH && (m & M_DISCH) && pa.dconc && !stage2)
            apply_discharges(s.disch, pa.dconc, pa.dconcmf, ic, j, k, open_c, Pc, cur.vr, dtv_c, ti0, e0);
        row.TI = ti0 + TIk_b;
        row.E = e0 + Ek_b;
        row.D = Dk;
        row.F = 0.;

        // ---------------- horizontal faces (explicit) ----------------
        if (HSPLIT) row.TI += cur.tih * dtv_c;           // tih = net horizontal flux into the cell (adt_hflux_kernel)
        if (do_h) {
            const bool o_w1 = (m & M_O_JM1) != 0, o_e1 = (m & M_O_JP1) != 0;
            const double fw = hface_flux<MH, LH>(s, all_set(m, M_CFU | M_O_JM1 | M_OPEN) && !(nf & NF_WEST), cur.qxw, cur.dhw, cur.Pw2, cur.Pw1,
                                                 Pc, cur.Pe1, (m & M_O_JM2) != 0, o_e1, cur.t_w2, cur.t_w, dtv_c,
                                                 cur.t_e, rho_wp, rdx_c, rho_wn, dux_m, dux_c);
            const double fe = hface_flux<MH, LH>(s, all_set(m, M_CFUE | M_O_JP1 | M_OPEN) && !(nf & NF_EAST), cur.qxe, cur.dhe, cur.Pw1, Pc,
                                                 cur.Pe1, cur.Pe2, o_w1, (m & M_O_JP2) != 0, cur.t_w, dtv_c, cur.t_e,
                                                 cur.t_e2, rho_ep, rdx_p, rho_en, dux_c, dux_p);
            double fsum = fw - fe;
            if (do_y) {
                // each lane builds its south face; the north face is the south face of lane+1
                double Ps1 = shfl_up_d(Pc, 1), Ps2 = shfl_up_d(Pc, 2), Pn1 = shfl_dn_d(Pc, 1);
                double t_s = shfl_up_d(dtv_c, 1), t_n = 0., t_s2 = 0.;
                if (far_h) { t_n = shfl_dn_d(dtv_c, 1); t_s2 = shfl_up_d(dtv_c, 2); }
                const double hP1 = __shfl_sync(0xffffffffu, cur.hP, 1);
                // strip halo, branch free: lanes 0,1 take cell i-2 (and lane 0 cell i-1) from the halo loads,
                // lane 31 takes cell i+1
                Ps2 = sel(lane < 2, cur.hP, Ps2);
                Ps1 = sel(lane == 0, hP1, Ps1);
                Pn1 = sel(lane == 31, cur.hP, Pn1);
                t_s = sel(lane == 0, cur.t_h, t_s);
                if (far_h) { t_s2 = sel(lane < 2, cur.t_h2, t_s2); t_n = sel(lane == 31, cur.t_h2, t_n); }
                const double fs = hface_flux<MH, LH>(s, all_set(m, M_CFV | M_O_IM1 | M_OPEN), cur.qys, cur.dhs, Ps2, Ps1, Pc,
                                                     Pn1, (m & M_O_IM2) != 0, (m & M_O_IP1) != 0, t_s2, t_s, dtv_c, t_n,
                                                     rho_sp, rdy_c, rho_sn, dvy_m, dvy_c);
                fsum += fs - shfl_dn_d(fs, 1);
            }
            row.TI += fsum * dtv_c;
        }

        // ---------------- vertical face k+1 (top of this cell) ----------------
        double Dn = 0., En_b = 0., TIn_b = 0.;            // contributions to row k+1
        if (FULL || s.K > 1) {
            // diffusion (AD:2708-2775 / 2779-2937): dvz is zero on non-compute W faces and in SmallDepths columns
            const double aux1 = dvz_p * dtv_c, aux2 = dvz_p * dtv_p;
            const double dP = Pp1 - Pc;
            row.E += aux1 * theta;
            row.F -= aux1 * theta;
            row.TI += aux1 * dP * omt;
            Dn = -aux2 * theta;
            En_b = aux2 * theta;
            TIn_b = -aux2 * dP * omt;
            // advection (AD:2941-3144); weights exist iff both cells are open (MF:10559), applied on compute faces
            const bool adv_on = all_set(m, top_req) && !(nf & NF_WT);
            const bool pos = qz_p > 0.;
            const double Puu = sel(pos, Pm1, Pp2), Pu = sel(pos, Pc, Pp1), Pd = sel(pos, Pp1, Pc);
            double du_u = 0., du_d = 0.;
            if (central_v) { const double a = s.DWZ[q], b = s.DWZ[q + sk]; du_u = sel(pos, a, b); du_d = sel(pos, b, a); }
            double wuu, wu, wd;
            oriented_weights<MV, LV>(s.method_v, s.limiter_v, s.upwind2_v != 0, s.vrelmax, qz_p, Puu, Pu, Pd,
                                     pos ? !(m & M_O_KM1) : !(m & M_O_KP2), sel(pos, dtv_m, dtv_pp),
                                     sel(pos, dtv_c, dtv_p), sel(pos, dtv_p, dtv_c), sel(pos, rdz_c, rdz_pp), rdz_p,
                                     du_u, du_d, wuu, wu, wd);
            const double qa = sel(adv_on, qz_p, 0.);
            if (advv_imp) {
                const double dfl = qa * sel(pos, wu, wd), efl = qa * sel(pos, wd, wu);   // D_flux, E_flux (MF:10583-10586)
                row.E += dfl * dtv_c;
                row.F += efl * dtv_c;
                Dn -= dfl * dtv_p;
                En_b -= efl * dtv_p;
            } else {
                const double fz = qa * (wuu * Puu + wu * Pu + wd * Pd);
                row.TI -= fz * dtv_c;
                TIn_b += fz * dtv_p;
            }
        }
        (void)far_v;

        // ---------------- land fill (AD:1753) ----------------
        row.TI = sel((m & M_LAND) != 0, NULL_REAL, row.TI);

        // ---------------- Thomas forward elimination, row k (MF:4087-4099) ----------------
        // branch free: a zero pivot keeps the previous W,G (the reference leaves them stale, MF:4092-4098)
        const double Wp0 = Wprev, Gp0 = Gprev;
        {
            const double aux = row.E + row.D * Wp0;
            const bool ok = aux != 0.;
            const double ra = fast_rcp(aux);
            Wprev = sel(ok, -row.F * ra, Wp0);
            Gprev = sel(ok, (row.TI - row.D * Gp0) * ra, Gp0);
            zp += ok ? 0u : 1u;
        }
        // ---------------- open boundary rows (AD:5369-5672) ----------------
        // Rare (boundary ring only): the row is amended and eliminated again, so the common path stays one
        // basic block.  Open boundary cells are never land.
        if (obc && open_c) {
            open_boundary_row<DISCH>(s, pa, q, m, Pc, qz_c, qz_p, dtv_c, row, ic, j, writer);
            const double aux = row.E + row.D * Wp0;
            if (aux != 0.) {
                const double ra = 1.0 / aux;
                Wprev = -row.F * ra;
                Gprev = (row.TI - row.D * Gp0) * ra;
            } else {
                Wprev = Wp0; Gprev = Gp0;
            }
        }
        Wsm[(size_t)(k - 1) * wstride] = Wprev;
        if (GGLOB) { if (writer && colwet) pa.pout[q] = Gprev; }      // G parked in the output array
        else Gsm[(size_t)(k - 1) * wstride] = Gprev;

        // ---------------- roll ----------------
        Dk = Dn; Ek_b = En_b; TIk_b = TIn_b;
        Pm1 = Pc; Pc = Pp1; Pp1 = Pp2; Pp2 = Pp3;
        dtv_m = dtv_c; dtv_c = dtv_p; dtv_p = dtv_pp;
        rdz_c = rdz_p; rdz_p = rdz_pp; rdz_pp = rdz_p3;
        qz_c = qz_p; qz_p = qz_pp;
        dvz_p = dvz_pp;
        q += sk;
        // PF == 2: every value of this level has been consumed, so its stage can be refilled with level k+2
        // (addresses are clamped to plane K+1; the surplus copies of the last two levels are never read)
        if (PF == 2) fetch_async(min(q + sk, c2d + sk * (s.K + 1)), (k - 1) & 1);
    };
    {
        int k = 1;
        for (; k + 1 <= s.K; k += 2) {
            level(k, lvA, lvB);
            level(k + 1, lvB, lvA);
        }
        if (k <= s.K) level(k, lvA, lvB);
    }

    if (PF == 2) __pipeline_wait_prior(0);
    // ---------------- back substitution (MF:4100-4105) ----------------
    if (writer && colwet) {
        double *__restrict__ O = pa.pout;
        int qo = c2d + sk * (s.K + 1);
        double x = 0.0;                                   // RES(KUB+1) = G(KUB+1) = 0 (halo row is the identity)
        O[qo] = x;
        if (GGLOB) {
            // G comes back from the output array: fetch 8 levels at a time so the loads overlap
            int k = s.K;
            for (; k >= 8; k -= 8) {
                double g[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) g[u] = O[qo - (u + 1) * sk];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    qo -= sk;
                    x = Wsm[(size_t)(k - 1 - u) * wstride] * x + g[u];
                    O[qo] = x;
                }
            }
            for (; k >= 1; --k) {
                qo -= sk;
                x = Wsm[(size_t)(k - 1) * wstride] * x + O[qo];
                O[qo] = x;
            }
        } else {
            for (int k = s.K; k >= 1; --k) {
                qo -= sk;
                x = Wsm[(size_t)(k - 1) * wstride] * x + Gsm[(size_t)(k - 1) * wstride];
                O[qo] = x;
            }
        }
        if (zp) atomicAdd(s.zero_pivots, (unsigned long long)zp);
    }
}

// -------------------------------------------------------------------------------------
// K7: the caller-side column steps of ModuleWaterProperties::Advection_Diffusion_Processes that run on a property
// right before its transport call (WP:14716-14759), on the device-resident fields:
//   FreeConvection (WP:13017-13074): from the first level whose upper neighbour is denser, the column is replaced by
//     its volume-weighted mean up to the surface;
//   SmallDepthsMixing_Processes (WP:12939-13012): columns thinner than the limit are mixed from KFloorZ to the surface
//     (Me%SmallDepths%ON itself is produced by adt_small_depths_kernel before the coefficient pass);
//   AddOffSet (WP:14724-14746): water points of the property and of its reference field are shifted by OffSet.
// One thread per (i, j, property), k serial in the reference's summation order.
// -------------------------------------------------------------------------------------
struct PremixArgs {
    int I, J, K, ld, sj, sk, nprop;
    const int *Open, *Water, *KFloorZ;
    const double *VolumeZ, *Density, *WaterColumnZ;      // Density / WaterColumnZ: nullptr = step not requested
    double limit;
    int *SmallDepths;                                     // adt_small_depths_kernel only
    double *pa[NPMAX], *pref[NPMAX];
    double off[NPMAX];
};

__global__ void adt_small_depths_kernel(const PremixArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= a.ld) return;
    const bool work = i >= 1 && i <= a.I && j >= 1 && j <= a.J;
    const int q2 = i + a.ld * j;
    a.SmallDepths[q2] = (work && a.Open[i + a.sj * j + a.sk * a.K] == 1 && a.WaterColumnZ[q2] < a.limit) ? 1 : 0;
}

__global__ void __launch_bounds__(128) adt_premix_kernel(const PremixArgs a) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, n = blockIdx.z;
    if (i > a.I) return;
    double *__restrict__ A = a.pa[n];
    const int c = i + a.sj * j, sk = a.sk;
    if (a.Density) {
        int ki = 0;
        for (int k = 1; k <= a.K; ++k) {
            const int q = c + sk * k;
            if (a.Open[q] == 1 && a.Open[q + sk] == 1 && (a.Density[q + sk] - a.Density[q]) > 0.) { ki = k; break; }
        }
        if (ki) {
            double Msum = 0., Vsum = 0.;
            for (int k = ki; k <= a.K; ++k) {
                const int q = c + sk * k;
                Msum = Msum + a.VolumeZ[q] * A[q];
                Vsum = Vsum + a.VolumeZ[q];
            }
            const double Cnew = Msum / Vsum;
            for (int k = ki; k <= a.K; ++k) A[c + sk * k] = Cnew;
        }
    }
    if (a.WaterColumnZ && a.Open[c + sk * a.K] == 1 && a.WaterColumnZ[i + a.ld * j] < a.limit) {
        double MassSum = 0., VolSum = 0.;
        const int kb = a.KFloorZ[i + a.ld * j];
        for (int k = kb; k <= a.K; ++k) {
            const int q = c + sk * k;
            MassSum = MassSum + a.VolumeZ[q] * A[q];
            VolSum = VolSum + a.VolumeZ[q];
        }
        const double Cnew = MassSum / VolSum;
        for (int k = kb; k <= a.K; ++k) A[c + sk * k] = Cnew;
    }
    const double off = a.off[n];
    if (off != 0.) {
        double *__restrict__ R = a.pref[n];
        for (int k = 1; k <= a.K; ++k) {
            const int q = c + sk * k;
            if (a.Water[q] == 1) {
                A[q] = A[q] + off;
                if (R) R[q] = R[q] + off;
            }
        }
    }
}

// after the transport call: the shift is taken out again (WP:14833-14858); off[] holds -OffSet
__global__ void __launch_bounds__(128) adt_offset_kernel(const PremixArgs a) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, n = blockIdx.z;
    if (i > a.I) return;
    const double off = a.off[n];
    if (off == 0.) return;
    double *__restrict__ A = a.pa[n], *__restrict__ R = a.pref[n];
    const int c = i + a.sj * j;
    for (int k = 1; k <= a.K; ++k) {
        const int q = c + a.sk * k;
        if (a.Water[q] == 1) {
            A[q] = A[q] + off;
            if (R) R[q] = R[q] + off;
        }
    }
}

// SetLimitsProperty (WP:20594-20720) after the transport call: clamp to MinValue / MaxValue and book the mass
// difference in Mass_created / Mass_Destroid (fp64 3-D accumulators); column form for Docycle_method 1, cell form else.
struct LimitArgs {
    int I, J, K, ld, sj, sk, docycle;
    const int *Water, *KFloorZ;
    const double *VolumeZ;
    double *pa[NPMAX], *created[NPMAX], *destroyed[NPMAX];
    double vmin[NPMAX], vmax[NPMAX];
    int min_on[NPMAX], max_on[NPMAX];
};
__global__ void __launch_bounds__(128) adt_limits_kernel(const LimitArgs a) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, n = blockIdx.z;
    if (i > a.I || !(a.min_on[n] || a.max_on[n])) return;
    double *__restrict__ A = a.pa[n];
    const int c = i + a.sj * j;
    int k0 = 1;
    if (a.docycle == 1) {
        if (a.Water[c + a.sk * a.K] != 1) return;
        k0 = a.KFloorZ[i + a.ld * j];
    }
    for (int k = k0; k <= a.K; ++k) {
        const int q = c + a.sk * k;
        if (a.docycle != 1 && a.Water[q] != 1) continue;
        double v = A[q];
        if (a.min_on[n] && v < a.vmin[n]) {
            a.created[n][q] = a.created[n][q] + (a.vmin[n] - v) * a.VolumeZ[q];
            v = a.vmin[n];
        }
        if (a.max_on[n] && v > a.vmax[n]) {
            a.destroyed[n][q] = a.destroyed[n][q] + (a.vmax[n] - v) * a.VolumeZ[q];
            v = a.vmax[n];
        }
        A[q] = v;
    }
}

__global__ void adt_shift_kernel(double *x, int n, double off) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) x[t] = x[t] + off;
}

// -------------------------------------------------------------------------------------
// K3a: ImposeNullGradient (AD:1926-1987): boundary cells take the compute-face-weighted mean of
// their (new) neighbours.  One thread per (boundary column, k).  Neighbours reached through a
// compute face are never boundary points themselves (HM:939-942), so the pass is order-free.
// -------------------------------------------------------------------------------------
struct BndArgs {
    int I, J, K, ld, nj, ncols;
    int sj, sk;               // element strides of j and k in the 3-D device arrays
    const int *cols;          // packed (i,j) of boundary columns
    const int *kfloor;
    const int *CFU, *CFV, *Bnd;   // ComputeFacesU3D / V3D (device mirrors), BoundaryPoints2D
    double *prop;             // new field (in place)
    const double *pref;
    int jmin, jmax;           // only boundary columns with jmin <= j <= jmax are processed (edge-first launches)
};

// -------------------------------------------------------------------------------------
// Carry: the new field of a step lives S columns beside the old one in the same buffer (in-place shift, see
// adt_api.cu).  Every cell the step kernel does not write -- halo rows, columns and planes, the leading-dimension
// padding, dry columns, columns outside the active range -- is copied from the old to the new position.  One thread per
// (i, column, property); launched per column chunk right before the chunk's step kernel.
// -------------------------------------------------------------------------------------
struct CarryArgs {
    int ld, nk, I, J, K, sj, sk, j0, ja, jb;  // columns j0 .. j0+gridDim.y-1; the step kernel advances columns ja .. jb
    const int *Water;
    const double *src[NPMAX];
    double *dst[NPMAX];
    const double *line[NPMAX];                // 2-D horizontally implicit step: the line solve's result, which IS the new field
    int twod;
};
__global__ void __launch_bounds__(128) adt_carry_kernel(const CarryArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = a.j0 + blockIdx.y, n = blockIdx.z;
    if (i >= a.ld) return;
    const double *__restrict__ S = a.src[n];
    double *__restrict__ D = a.dst[n];
    const int c = i + a.sj * j;
    const bool solved = j >= a.ja && j <= a.jb && i >= 1 && i <= a.I && a.Water[c + a.sk * a.K] == 1;   // MF:4086
    if (a.twod) {
        // every cell of a line was solved (land cells included, THOMAS_3D has no water mask, MF:3751-3875); the halo cells
        // around the solved columns carry what the Orlanski boundary wrote beside its boundary points
        const bool on_line = (j >= a.ja && j <= a.jb) || (j == 0 && a.ja == 1) || (j == a.J + 1 && a.jb == a.J);
        for (int k = 0; k < a.nk; ++k) D[c + a.sk * k] = (on_line && k >= 1 && k <= a.K ? a.line[n] : S)[c + a.sk * k];
        return;
    }
    if (solved) { D[c] = S[c]; return; }       // rows 1 .. K+1 are written by the step kernel
    for (int k = 0; k < a.nk; ++k) D[c + a.sk * k] = S[c + a.sk * k];
}

__global__ void adt_nullgrad_kernel(const BndArgs b) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)b.ncols * b.K) return;
    const int c = (int)(t / b.K), k = (int)(t % b.K) + 1;
    const int i = b.cols[2 * c], j = b.cols[2 * c + 1];
    if (j < b.jmin || j > b.jmax) return;
    const long q2 = (long)i + (long)b.ld * j;
    int kf = b.kfloor[q2];
    kf = kf < 0 ? -kf : kf;
    if (k < kf) return;
    const long q = (long)i + (long)b.sj * j + (long)b.sk * k;
    const int cVn = b.CFV[q + 1] == 1 ? 1 : 0, cVs = b.CFV[q] == 1 ? 1 : 0;
    const int cUe = b.CFU[q + b.sj] == 1 ? 1 : 0, cUw = b.CFU[q] == 1 ? 1 : 0;
    const int aux = cVn + cVs + cUe + cUw;
    if (aux > 0)
        b.prop[q] = (b.prop[q + 1] * cVn + b.prop[q - 1] * cVs + b.prop[q + b.sj] * cUe + b.prop[q - b.sj] * cUw) /
                    (double)aux;
}

// K3b: Prop_CyclicBoundary (AD:2121-2224), phase 0: boundary cells <- ReferenceProp;
// phase 1: wrap j (per i), phase 2: wrap i (per j).  Launched as three ordered kernels.
__global__ void adt_cyclic_kernel(const BndArgs b, int phase) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long sj = b.sj, sk = b.sk;
    if (phase == 0) {
        if (t >= (long)b.ncols * b.K) return;
        const int c = (int)(t / b.K), k = (int)(t % b.K) + 1;
        const long q = (long)b.cols[2 * c] + sj * b.cols[2 * c + 1] + sk * k;
        b.prop[q] = b.pref[q];
    } else if (phase == 1) {
        if (t >= (long)(b.I - 2) * b.K) return;
        const int i = (int)(t / b.K) + 2, k = (int)(t % b.K) + 1;
        const long r1 = (long)i + sj * 1, rJ = (long)i + sj * b.J;                 // 3-D column bases of (i,1), (i,J)
        const long f1 = (long)i + (long)b.ld * 1, fJ = (long)i + (long)b.ld * b.J;   // the same columns in 2-D arrays
        if (b.Bnd[f1] == 1 && b.Bnd[fJ] == 1) {
            if (k >= b.kfloor[fJ - b.ld]) b.prop[r1 + sk * k] = b.prop[rJ - sj + sk * k];
            if (k >= b.kfloor[f1 + b.ld]) b.prop[rJ + sk * k] = b.prop[r1 + sj + sk * k];
        }
    } else {
        if (t >= (long)(b.J - 2) * b.K) return;
        const int j = (int)(t / b.K) + 2, k = (int)(t % b.K) + 1;
        const long r1 = 1 + sj * j, rI = (long)b.I + sj * j;
        const long f1 = 1 + (long)b.ld * j, fI = (long)b.I + (long)b.ld * j;
        if (b.Bnd[f1] == 1 && b.Bnd[fI] == 1) {
            if (k >= b.kfloor[fI - 1]) b.prop[r1 + sk * k] = b.prop[rI - 1 + sk * k];
            if (k >= b.kfloor[f1 + 1]) b.prop[rI + sk * k] = b.prop[r1 + 1 + sk * k];
        }
    }
}

// Prop_CyclicBoundary across column slabs: the j wrap joins global column 1 (first rank) and J (last rank).  Each of the two
// packs what the other needs -- PROP of its column next to the boundary column, BoundaryPoints2D of its boundary column and
// KFloorZ of the packed column -- and applies the other's buffer to its own boundary column.
//   buf: [K][I] values (k = 1 .. K, i = 1 .. I), then [I] BoundaryPoints2D, then [I] KFloorZ (as doubles)
__global__ void adt_cyclic_edge_pack_kernel(const BndArgs b, int j_val, int j_bnd, double *buf) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long nv = (long)b.K * b.I;
    if (t < nv) {
        const int i = (int)(t % b.I) + 1, k = (int)(t / b.I) + 1;
        buf[t] = b.prop[(long)i + (long)b.sj * j_val + (long)b.sk * k];
    } else if (t < nv + b.I) {
        const int i = (int)(t - nv) + 1;
        buf[t] = (double)b.Bnd[(long)i + (long)b.ld * j_bnd];
        buf[t + b.I] = (double)b.kfloor[(long)i + (long)b.ld * j_val];
    }
}
__global__ void adt_cyclic_edge_apply_kernel(const BndArgs b, int j_bnd, const double *buf) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)(b.I - 2) * b.K) return;
    const int i = (int)(t % (b.I - 2)) + 2, k = (int)(t / (b.I - 2)) + 1;
    const long nv = (long)b.K * b.I;
    if (b.Bnd[(long)i + (long)b.ld * j_bnd] == 1 && buf[nv + i - 1] == 1.0 && k >= (int)buf[nv + b.I + i - 1])
        b.prop[(long)i + (long)b.sj * j_bnd + (long)b.sk * k] = buf[(long)(k - 1) * b.I + (i - 1)];
}

// -------------------------------------------------------------------------------------
// K5: cell-face mass fluxes for the box budgets (CalcHorizontal/Vertical Adv/Dif Flux, AD:3356-3954;
// GetAdvFlux / GetDifFlux, AD:697-851).  Only launched for properties whose call carries CellFluxes (box
// time-series output steps, WP:14956-15032).  One thread per work cell: its west U face, south V face and
// top W face.  Explicit shares use the old field, implicit shares (vertical) the new one (AD:1885-1916);
// the face weights always come from the old field (AD:2966-3001).
// -------------------------------------------------------------------------------------
struct FluxArgs {
    int I, J, K, ld, sj, sk;
    int method_h, limiter_h, method_v, limiter_v, upwind2_h, upwind2_v, vertical1d, xzflow;
    const unsigned char *nfmask; unsigned nfsel;   // as in StepArgs / PropArgs
    double vrelmax, w_advv, theta;                 // ImpExp_AdvV (0 or 1), ImpExp_DifV as passed by the caller
    const double *pold, *pnew;                     // field at time n / n+1
    const double *pmid;                            // the field the vertical half started from (= pold unless the step was split)
    int impl_x, impl_y;                            // ImpExp_AdvXX / ImpExp_AdvYY == ImplicitScheme: that direction's flux takes the new field
    const double *qx, *qy, *qz, *dtv, *dhu, *dhv, *dvz, *rdz, *rdx, *rdy, *DUX, *DVY, *DWZ;
    const uint32_t *mask;
    double *ax, *ay, *az, *dx, *dy, *dz;
};

__device__ __forceinline__ double adv_face_flux(int method, int limiter, bool up2, double vrelmax, double Q,
                                                const double Pw[4], const double Pa[4], bool o1, bool o4,
                                                const double t[4], double rd12, double rd23, double rd34, double du2,
                                                double du3) {
    // weights from the stencil Pw (old field); they multiply the stencil Pa (old or new field)
    const bool pos = Q > 0.;
    double wuu, wu, wd;
    oriented_weights<0, 0>(method, limiter, up2, vrelmax, Q, sel(pos, Pw[0], Pw[3]), sel(pos, Pw[1], Pw[2]),
                           sel(pos, Pw[2], Pw[1]), pos ? !o1 : !o4, sel(pos, t[0], t[3]), sel(pos, t[1], t[2]),
                           sel(pos, t[2], t[1]), sel(pos, rd12, rd34), rd23, sel(pos, du2, du3), sel(pos, du3, du2),
                           wuu, wu, wd);
    return Q * (wuu * sel(pos, Pa[0], Pa[3]) + wu * sel(pos, Pa[1], Pa[2]) + wd * sel(pos, Pa[2], Pa[1]));
}

__global__ void __launch_bounds__(128) adt_cell_flux_kernel(const FluxArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i > a.I) return;
    const int sj = a.sj, sk = a.sk, sj2 = a.ld;
    const int q = i + sj * j + sk * k, q2 = i + sj2 * j;
    const unsigned m = a.mask[q];
    const unsigned nf = a.nfsel ? (a.nfmask[q] & a.nfsel) : 0u;
    const double *__restrict__ P = a.pold;
    const double *__restrict__ N = a.pnew;
    const int jw2 = (j >= 2) ? 2 * sj : sj;
    if (!a.vertical1d) {
        // ---- U face (i,j,k), between cells j-1 and j ----
        if (m & M_CFU) {
            const double Pw[4] = {P[q - jw2], P[q - sj], P[q], P[q + sj]};
            double adv = 0.;
            if (all_set(m, M_O_JM1 | M_OPEN) && !(nf & NF_WEST)) {
                const double t[4] = {a.dtv[q - jw2], a.dtv[q - sj], a.dtv[q], a.dtv[q + sj]};
                const double Pn[4] = {N[q - jw2], N[q - sj], N[q], N[q + sj]};
                adv = adv_face_flux(a.method_h, a.limiter_h, a.upwind2_h != 0, a.vrelmax, a.qx[q], Pw, a.impl_x ? Pn : Pw,
                                    (m & M_O_JM2) != 0, (m & M_O_JP1) != 0, t, a.rdx[q2 - sj2], a.rdx[q2],
                                    a.rdx[q2 + sj2], a.DUX[q2 - sj2], a.DUX[q2]);
            }
            a.ax[q] = adv;
            a.dx[q] = -a.dhu[q] * (Pw[2] - Pw[1]);
        }
        // ---- V face (i,j,k), between cells i-1 and i ----
        if (!a.xzflow && (m & M_CFV)) {
            const double Pw[4] = {P[q - (i >= 2 ? 2 : 1)], P[q - 1], P[q], P[q + 1]};
            double adv = 0.;
            if (all_set(m, M_O_IM1 | M_OPEN)) {
                const double t[4] = {a.dtv[q - (i >= 2 ? 2 : 1)], a.dtv[q - 1], a.dtv[q], a.dtv[q + 1]};
                const double Pn[4] = {N[q - (i >= 2 ? 2 : 1)], N[q - 1], N[q], N[q + 1]};
                adv = adv_face_flux(a.method_h, a.limiter_h, a.upwind2_h != 0, a.vrelmax, a.qy[q], Pw, a.impl_y ? Pn : Pw,
                                    (m & M_O_IM2) != 0, (m & M_O_IP1) != 0, t, a.rdy[q2 - 1], a.rdy[q2], a.rdy[q2 + 1],
                                    a.DVY[q2 - 1], a.DVY[q2]);
            }
            a.ay[q] = adv;
            a.dy[q] = -a.dhv[q] * (Pw[2] - Pw[1]);
        }
    }
    // ---- W face (i,j,k+1), between cells k and k+1 ----
    if (a.K > 1 && (m & M_CFWT)) {
        const int qt = q + sk;
        const int q2k = (k + 2 <= a.K + 1) ? q + 2 * sk : qt;
        const double *__restrict__ Mf = a.pmid;
        const double Pw[4] = {Mf[q - sk], Mf[q], Mf[qt], Mf[q2k]};
        const double Pn[4] = {N[q - sk], N[q], N[qt], N[q2k]};
        double dif = 0.;
        if (a.theta < 1.) dif = dif - (1. - a.theta) * a.dvz[qt] * (Pw[2] - Pw[1]);      // VerticalDiffusion share (AD:2768)
        if (a.theta > 0.) dif = dif - a.theta * a.dvz[qt] * (Pn[2] - Pn[1]);               // after the solve (AD:1893)
        a.dz[qt] = dif;
        double adv = 0.;
        const unsigned mtop = a.mask[i + sj * j + sk * a.K];
        if (!a.vertical1d && (mtop & M_COLOPEN) && all_set(m, M_OPEN | M_O_KP1) && !(nf & NF_WT)) {
            const double t[4] = {a.dtv[q - sk], a.dtv[q], a.dtv[qt], a.dtv[q2k]};
            double du2 = 0., du3 = 0.;
            if (a.method_v == MOHID_CentralDif || a.method_v == MOHID_LeapFrog) { du2 = a.DWZ[q]; du3 = a.DWZ[qt]; }
            const bool o1 = (m & M_O_KM1) != 0, o4 = (m & M_O_KP2) != 0;
            if (a.w_advv < 1.)
                adv += (1. - a.w_advv) * adv_face_flux(a.method_v, a.limiter_v, a.upwind2_v != 0, a.vrelmax, a.qz[qt],
                                                       Pw, Pw, o1, o4, t, a.rdz[q], a.rdz[qt], a.rdz[q2k], du2, du3);
            if (a.w_advv > 0.)
                adv += a.w_advv * adv_face_flux(a.method_v, a.limiter_v, a.upwind2_v != 0, a.vrelmax, a.qz[qt], Pw,
                                                Pn, o1, o4, t, a.rdz[q], a.rdz[qt], a.rdz[q2k], du2, du3);
        }
        a.az[qt] = adv;
    }
}

// -------------------------------------------------------------------------------------
// Stand-alone column solve: THOMASZ_NewType2 (MF:4026-4123) on caller-supplied D, E, F, TI -- the drop-in for the
// reference's own GPU entry SolveThomas_C / DevThomasIK (ModuleCuda.F90:103-111, CudaThomas/Thomas.cu:62-131).
// Same elimination as K2 (branch-free reciprocal pivot, a zero pivot keeps the previous W, G and is counted), rows
// 1 .. K+1, one thread per column, i fastest; W in a scratch field, G parked in Res.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) adt_thomas_z_kernel(int I, int J, int K, int sj, int sk, const double *D,
                                                           const double *E, const double *F, const double *TI,
                                                           const int *Water, double *Res, double *Wg,
                                                           unsigned long long *zero_pivots) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y;
    if (i > I) return;
    const int c = i + sj * j;
    if (Water && Water[c + sk * K] != 1) return;                      // MF:4086
    double Wp = 0., Gp = 0.;
    unsigned zp = 0;
    for (int k = 1; k <= K + 1; ++k) {
        const int q = c + sk * k;
        const double d = k == 1 ? 0. : D[q];                          // the first row ignores D (MF:4087-4088)
        const double aux = E[q] + d * Wp;
        const bool ok = aux != 0.;
        const double ra = fast_rcp(aux);
        const double w = -F[q] * ra, g = (TI[q] - d * Gp) * ra;
        Wp = ok ? w : Wp; Gp = ok ? g : Gp;
        zp += ok ? 0u : 1u;
        Wg[q] = Wp; Res[q] = Gp;
    }
    double x = Gp;                                                    // RES(KUB+1) = G(KUB+1)
    for (int k = K; k >= 1; --k) {
        const int q = c + sk * k;
        x = Wg[q] * x + Res[q];
        Res[q] = x;
    }
    if (zp) atomicAdd(zero_pivots, (unsigned long long)zp);
}

// -------------------------------------------------------------------------------------
// Mass of every property in every column j: sum over i and k of P * VolumeZ on water points, in a fixed summation
// order (thread-serial over its cells, then a shared-memory tree), so the value of a column does not depend on how the
// domain is split over GPUs.  One block per (j, property).
// -------------------------------------------------------------------------------------
struct MassArgs {
    int I, K, sj, sk, nj;
    const int *Water;
    const double *VolumeZ;
    const double *prop[NPMAX];
    double *out;               // [nprop][nj]
};
__global__ void __launch_bounds__(256) adt_column_mass_kernel(const MassArgs a) {
    __shared__ double red[256];
    const int j = blockIdx.x, n = blockIdx.y;
    const double *__restrict__ P = a.prop[n];
    double acc = 0.;
    for (int i = 1 + (int)threadIdx.x; i <= a.I; i += 256)
        for (int k = 1; k <= a.K; ++k) {
            const int q = i + a.sj * j + a.sk * k;
            if (a.Water[q] == 1) acc += P[q] * a.VolumeZ[q];
        }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.out[(size_t)n * a.nj + j] = red[0];
}

// -------------------------------------------------------------------------------------
// Box budgets (BoxDifFluxes3D, MOHIDBase2/ModuleBoxDif.F90:2659-2776, fed by WP:14956-15032 with
// MassFluxes = AdvFlux + DifFlux and OpenPoints3D as the mask): the flux through every cell face that separates two
// boxes is added to Fluxes(OUT, IN) and subtracted from Fluxes(IN, OUT).  Boundary faces are the ones
// FindAdjacentBoxesBoundaries3D marks (BoxDif:1697-1735): the cell has a box (> -55) and is a water point, the
// neighbour across the face has another box (> -55).  The matrix is (0:nb, 0:nb) in Fortran order.
// -------------------------------------------------------------------------------------
struct BoxArgs {
    int I, J, K, ld, sj, sk, nb1;           // nb1 = NumberOfBoxes3D + 1
    int with_z;                             // KUB > KLB (WP:14981)
    const int *Boxes, *Water, *Open;
    const double *ax, *ay, *az, *dx, *dy, *dz;
    double *fluxes;                         // [nb1 * nb1]
};
__global__ void __launch_bounds__(128) adt_box_flux_kernel(const BoxArgs a) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > a.I) return;
    const int q = i + a.sj * j + a.sk * k;
    const int b = a.Boxes[q];
    if (b <= -55 || a.Water[q] != 1 || a.Open[q] != 1) return;
    auto add = [&](int out, int in, double f) {
        if (out < 0 || in < 0 || out >= a.nb1 || in >= a.nb1) return;
        atomicAdd(a.fluxes + out + (size_t)a.nb1 * in, f);
        atomicAdd(a.fluxes + in + (size_t)a.nb1 * out, -f);
    };
    const int bx = a.Boxes[q + a.sj], by = a.Boxes[q + 1], bz = a.Boxes[q + a.sk];
    if (bx != b && bx > -55) add(b, bx, a.ax[q + a.sj] + a.dx[q + a.sj]);
    if (by != b && by > -55) add(b, by, a.ay[q + 1] + a.dy[q + 1]);
    if (bz != b && bz > -55 && a.with_z) add(b, bz, a.az[q + a.sk] + a.dz[q + a.sk]);
}

// -------------------------------------------------------------------------------------
// FreeVerticalMovementIteration (MOHIDWater/ModuleFreeVerticalMovement.F90:1531-1650): settling / rising of a property
// with its own vertical velocity, first-order upwind, implicit (ImpExp_AdvV = 0 in THIS module's convention) or explicit
// (= 1).  Pass 1 builds D, E, F, TI (SetMatrixValue fills :1555-1558, BottomBoundary :2141-2227, VerticalFreeConvection
// :1651-1775, land fill :1583-1592) and the explicit share of FreeConvFlux (:1779-1805); the column solve is
// adt_thomas_z_kernel (THOMASZ_NewType2, :1610-1622); pass 2 adds the implicit share of the flux from the new field.
// -------------------------------------------------------------------------------------
struct FvmArgs {
    int I, J, K, sj, sk, ld;
    const int *Mask, *Land, *KFloorZ;       // WaterPointsorOpenPoints (:1549-1553), LandPoints3D, KFloor_Z
    const double *VolumeZ, *Velocity, *Area, *DepProb;   // DepProb may be null
    int deposition, non_cohesive;
    double dt, impexp;                      // DTProp, ImpExp_AdvV (0 implicit, 1 explicit)
    const double *C;                        // concentration (old in pass 1, new in pass 2)
    double *D, *E, *F, *TI, *flux;          // flux may be null
};
__device__ __forceinline__ double fvm_velocity(const FvmArgs &a, int c2, int q, int k) {
    // BottomBoundary: the velocity at the bottom face of the column is zero, or scaled by the deposition probability
    double v = a.Velocity[q];
    if (k == a.KFloorZ[c2]) {
        if (!a.deposition) v = 0.;
        else if (!a.non_cohesive) v = v * a.DepProb[c2];
    }
    return v;
}
__global__ void __launch_bounds__(128) adt_fvm_coef_kernel(const FvmArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= a.ld) return;
    const int c2 = i + a.ld * j, q = i + a.sj * j + a.sk * k;
    double D = 0., E = 1., F = 0., TI = a.C[q], fl = 0.;
    const bool inwork = i >= 1 && i <= a.I && j >= 1 && j <= a.J && k >= 1 && k <= a.K;
    if (inwork && a.Mask[i + a.sj * j + a.sk * a.K] == 1) {
        const double dtv = a.dt / a.VolumeZ[q];
        const double w1 = fvm_velocity(a, c2, q, k) * a.Area[c2];
        const double w2 = k < a.K ? fvm_velocity(a, c2, q + a.sk, k + 1) * a.Area[c2] : 0.;
        const double aw1 = fabs(w1), aw2 = fabs(w2);
        const double d_flux = -(w1 + aw1) / 2.0, e_flux = -(w1 - aw1) / 2.0;
        const double coef_d = d_flux * dtv, coef_e = ((w2 + aw2) / 2.0 + e_flux) * dtv, coef_f = ((w2 - aw2) / 2.0) * dtv;
        if (a.impexp == 0.0) { D = D + coef_d; E = E + coef_e; F = F + coef_f; }
        if (a.impexp == 1.0) TI = TI - (coef_d * a.C[q - a.sk] + coef_e * a.C[q] + coef_f * a.C[q + a.sk]);
        if (a.Mask[q] == 1) fl = -a.impexp * (d_flux * a.C[q - a.sk] + e_flux * a.C[q]);
    }
    if (inwork) TI = TI * (1. - (double)a.Land[q]) + (double)a.Land[q] * NULL_REAL;
    a.D[q] = D; a.E[q] = E; a.F[q] = F; a.TI[q] = TI;
    if (a.flux) a.flux[q] = fl;
}
// implicit share of FreeConvFlux from the new field (CalcVerticalFreeConvFlux with Weigth = 1 - ImpExp_AdvV)
__global__ void __launch_bounds__(128) adt_fvm_flux_kernel(const FvmArgs a) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y, k = 1 + blockIdx.z;
    if (i > a.I) return;
    const int c2 = i + a.ld * j, q = i + a.sj * j + a.sk * k;
    if (a.Mask[q] != 1 || a.Mask[i + a.sj * j + a.sk * a.K] != 1) return;
    const double w1 = fvm_velocity(a, c2, q, k) * a.Area[c2];
    const double aw1 = fabs(w1);
    const double d_flux = -(w1 + aw1) / 2.0, e_flux = -(w1 - aw1) / 2.0;
    a.flux[q] = a.flux[q] - (1.0 - a.impexp) * (d_flux * a.C[q - a.sk] + e_flux * a.C[q]);
}

// -------------------------------------------------------------------------------------
// ModuleHydroIntegration (MOHIDBase1/ModuleHydroIntegration.F90): the fluxes, mapping and initial volume a property
// with DTInterval = n hydrodynamic steps is transported with (GetHydroIntegration* at WP:14615-14647).  The integration
// arrays ARE the handle's step mirrors: OneIntegrationStep (:843-906) keeps the running mean of the horizontal fluxes
// and the union of the compute faces, EndIntegrationStep (:910-994) closes continuity for Wflux_Z and derives
// ComputeFacesW3D / OpenPoints3D.
// -------------------------------------------------------------------------------------
struct HintArgs {
    int I, J, K, ni, nj, ld, sj, sk, n;             // n = CurrentIntegration%n after the increment
    const double *fx, *fy, *disch;                  // this hydrodynamic step (device copies; disch may be null)
    const int *cfu, *cfv;
    double *WX, *WY, *WZ, *D;                       // integrated fluxes, discharges
    int *CFU, *CFV, *CFW, *Open;
    const int *Water, *Bnd;
    const double *V, *V0;                           // VolumeZ, InitialVolume
    double dt;
};
__global__ void __launch_bounds__(128) adt_hint_step_kernel(const HintArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= a.ni) return;
    const int q = i + a.sj * j + a.sk * k;
    if (k < 1 || k > a.K || i < 1 || j < 1) return;
    const double n = (double)a.n;
    if (i <= a.I + 1 && j <= a.J + 1) {             // faces: ILB..IUB+1, JLB..JUB+1
        a.WX[q] = (a.WX[q] * (n - 1.) + a.fx[q]) / n;
        a.WY[q] = (a.WY[q] * (n - 1.) + a.fy[q]) / n;
        if (a.cfu[q] > 0) a.CFU[q] = 1;
        if (a.cfv[q] > 0) a.CFV[q] = 1;
    }
    if (i <= a.I && j <= a.J && a.D) a.D[q] = (a.D[q] * (n - 1.) + (a.disch ? a.disch[q] : 0.)) / n;
}
__global__ void __launch_bounds__(128) adt_hint_end_kernel(const HintArgs a) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, j = 1 + blockIdx.y;
    if (i > a.I) return;
    const int c = i + a.sj * j, c2 = i + a.ld * j, sk = a.sk, sj = a.sj;
    const double keep = 1.0 - (double)a.Bnd[c2];
    double wz = a.WZ[c + sk];                                     // Wflux_Z(KLB) = 0 since the re-initialisation
    for (int k = 1; k <= a.K; ++k) {
        const int q = c + sk * k;
        const double dVdt = (a.V[q] - a.V0[q]) / a.dt;
        wz = (wz + a.WX[q] - a.WX[q + sj] + a.WY[q] - a.WY[q + 1] - dVdt + (a.D ? a.D[q] : 0.)) * keep;
        a.WZ[q + sk] = wz;
    }
    const int qt = c + sk * a.K;
    if (a.CFU[qt] + a.CFU[qt + sj] + a.CFV[qt] + a.CFV[qt + 1] > 0)
        for (int k = 2; k <= a.K; ++k)
            if (a.Water[c + sk * (k - 1)] == 1) a.CFW[c + sk * k] = 1;
    for (int k = 1; k <= a.K; ++k) {
        const int q = c + sk * k;
        if (a.CFU[q] + a.CFU[q + sj] + a.CFV[q] + a.CFV[q + 1] + a.CFW[q] + a.CFW[q + sk] > 0) a.Open[q] = 1;
    }
}

// -------------------------------------------------------------------------------------
// K4: gather / scatter `width` j-columns of nprop properties to / from a contiguous buffer
// laid out [n][k][w][i] (i fastest).  Coalesced on both sides.
// -------------------------------------------------------------------------------------
struct PackArgs {
    int ld, nj, nk, nprop, j0, width;
    int sj, sk;               // element strides of j and k in the 3-D device arrays
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
        const long q = (long)i + (long)a.sj * (a.j0 + w) + (long)a.sk * k;
        if (unpack) a.prop[n][q] = buf[t];
        else buf[t] = a.prop[n][q];
    }
}

}  // namespace adt
