// =====================================================================================
//  adt_lean_kernel.cuh -- the headline form of the fused transport step (round 2).
//
//  Same arithmetic as adt_transport_kernel<TVD SuperBee | Upwind, FULL> (adt_kernels.cuh), different split of the
//  work: everything about a face that does not depend on the property -- flow direction, the near-boundary rule
//  (MF:10563-10570), whether the face advects at all (MF:10559, AD:4467), the Courant factor 0.5(1 - Cr) of the TVD
//  weight (MF:10858), the metric ratio of the limiter argument -- is computed ONCE per step by the coefficient pass
//  (adt_lean_coef_kernel, amortised over the N batched properties) and handed to the step kernel as four 32-byte
//  packs per cell, each fetched with one 256-bit load:
//      U = { Qh, A, B, rho }             west U face of the cell   (AD:4368-4582, 5156-5250)
//      V = { Qh, A, B, rho }             south V face              (AD:4739-4953, 5254-5365)
//      W = { qa, hv, rdu, rdc }          bottom W face             (AD:2941-3144)
//      C = { DT/V, Vold/V, Diff_V_Const (bottom face), mask }
//  With Qa the face flux where the face advects (else 0), Cr = Q DT/V_upwind and hc = 0.5 (1 - Cr) (0 on near-boundary
//  faces with Upwind2), the flux of property through a horizontal face, Qa (Pu + hc psi dP) - dh (P_hi - P_lo), is
//  evaluated in the direction-free form  Qh (P_lo + P_hi) - A g + B sgn(g) lim  with g = P_hi - P_lo, Qh = Qa/2,
//  A = |Qa|/2 + dh (upwinding + Diff_H_Const_U/V, AD:1549-1597), B = |Qa| hc and lim = psi |dP| the limited increment;
//  rho = (du_u+du_d)/(du_u+du_uu) of the upwind side; rdu/rdc = 1/(du+du) of the upwind pair and of the face; the
//  mask carries four extra bits with the flow direction of the west, east, south and bottom face.
//  The four packs of 32 consecutive cells lie next to each other (4 x 1 KB), so one per-thread pointer with constant
//  offsets addresses all of them.
//
//  The step kernel keeps the warp <-> (31-cell strip, column j, property) mapping, the register look-ahead and the
//  in-register Thomas elimination (W in shared memory, G parked in the output array) of the round-1 kernel, and
//    * evaluates the limiter on neighbour differences d(j) = P(j) - P(j-1) shared by the two U faces of the cell;
//    * walks the column with the vertical faces skewed by one level: level k evaluates the face BELOW cell k, which
//      completes row k-1 of the tridiagonal system, so every value it needs (DT/V and P of k-2 .. k+1) is already in
//      registers and the vertical look-ahead of the round-1 kernel (five more loads per level) disappears.
//  Results agree with the round-1 kernel to rounding (tools/lean_check.py, tests/test_gpu_parity.py).
// =====================================================================================
#pragma once
#include <type_traits>

#include "adt_kernels.cuh"

namespace adt {

enum : unsigned {
    M_POS_W = 1u << 27,   // Wflux_X(i,j,k)   > 0
    M_POS_E = 1u << 28,   // Wflux_X(i,j+1,k) > 0
    M_POS_S = 1u << 29,   // Wflux_Y(i,j,k)   > 0
    M_POS_B = 1u << 30    // Wflux_Z(i,j,k)   > 0 (bottom face)
};

__device__ __forceinline__ Pack4 ld_pack(const Pack4 *p) {
    Pack4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ double ld_f64(const double *p) {
    double r;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_pack(Pack4 *p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// -------------------------------------------------------------------------------------
// Coefficient pass of the lean path.  One thread per allocated cell of the column chunk jc0 .. jc0+ncol-1.
// Pack layout: rows (j - jc0) + ncol*k of ldp = 32*nt32 cells; inside a row, groups of 32 cells hold their four packs
// back to back: pack p of cell i at Pack4 index ((row*nt32 + i/32)*4 + p)*32 + i%32.
// -------------------------------------------------------------------------------------
struct LeanCoefArgs {
    CoefArgs c;                       // raw inputs, extents, strides, Schmidt numbers (its outputs are not used)
    Pack4 *pk;
    int jc0, ncol, nt32;
    int tvd;                          // 1: P2_TVD (hc / hv are needed), 0: first-order upwind
    int upwind2_h, upwind2_v;
    const double *rhoUp, *rhoUn, *rhoVp, *rhoVn;   // 2-D metric ratios (adt_grid2d_rho_kernel)
};
constexpr int PK_U = 0, PK_V = 32, PK_W = 64, PK_C = 96;     // Pack4 offsets of the four packs inside a 32-cell group

__host__ __device__ __forceinline__ long pack_index(int i, int row, int nt32) {
    return ((long)row * nt32 + (i >> 5)) * 128 + (i & 31);
}

// rho of the limiter argument for both flow directions of the west U face and the south V face of every column;
// the clamped probes repeat the ones of adt_transport_kernel (rdx_pp, rdy_p of the last lane)
__global__ void adt_grid2d_rho_kernel(int ni, int nj, int ld, int I, int J, const double *rdx, const double *rdy,
                                      double *rhoUp, double *rhoUn, double *rhoVp, double *rhoVn) {
    const long n2 = (long)ld * nj;
    for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n2; q += (long)gridDim.x * blockDim.x) {
        const int i = (int)(q % ld), j = (int)(q / ld);
        double up = 0., un = 0., vp = 0., vn = 0.;
        if (i < ni) {
            const double xc = rdx[q], yc = rdy[q];
            const double xm = j >= 1 ? rdx[q - ld] : 0., xp = rdx[q + (j + 1 <= J + 1 ? ld : 0)];
            const double ym = i >= 1 ? rdy[q - 1] : 0., yp = rdy[q + (i + 1 <= I + 1 ? 1 : 0)];
            up = ratio_or_zero(xm, xc); un = ratio_or_zero(xp, xc);
            vp = ratio_or_zero(ym, yc); vn = ratio_or_zero(yp, yc);
        }
        rhoUp[q] = up; rhoUn[q] = un; rhoVp[q] = vp; rhoVn[q] = vn;
    }
}

template <bool EDGE>
__device__ __forceinline__ void adt_lean_coef_cell(const LeanCoefArgs &A, const int i, const int j, const int k) {
    const CoefArgs &a = A.c;
    const int sj = a.sj, sk = a.sk, sj2 = a.ld;
    const int q2 = i + sj2 * j;
    const int q = i + sj * j + sk * k;
    const long qc = pack_index(i, (j - A.jc0) + A.ncol * k, A.nt32);
    const bool im1 = !EDGE || i >= 1, im2 = !EDGE || i >= 2, ip1 = !EDGE || i + 1 < a.ni, ip2 = !EDGE || i + 2 < a.ni;
    const bool jm1 = !EDGE || j >= 1, jm2 = !EDGE || j >= 2, jp1 = !EDGE || j + 1 < a.nj, jp2 = !EDGE || j + 2 < a.nj;
    const bool km1 = !EDGE || k >= 1, km2 = !EDGE || k >= 2, kp1 = !EDGE || k + 1 < a.nk, kp2 = !EDGE || k + 2 < a.nk;
    const int oim1 = im1 ? -1 : 0, oim2 = im2 ? -2 : 0, oip1 = ip1 ? 1 : 0, oip2 = ip2 ? 2 : 0;
    const int ojm1 = jm1 ? -sj : 0, ojm2 = jm2 ? -2 * sj : 0, ojp1 = jp1 ? sj : 0, ojp2 = jp2 ? 2 * sj : 0;
    const int okm1 = km1 ? -sk : 0, okm2 = km2 ? -2 * sk : 0, okp1 = kp1 ? sk : 0, okp2 = kp2 ? 2 * sk : 0;
    const int qtop = i + sj * j + sk * a.K;

    // ---- loads ----
    const int open_c = a.Open[q], cfu_c = a.CFU[q], cfv_c = a.CFV[q], cfw_c = a.CFW[q], land_c = a.Land[q];
    const int bnd_c = a.Bnd[q2];
    const int cfu_e = a.CFU[q + ojp1], cfv_n = a.CFV[q + oip1], cfw_t = a.CFW[q + okp1];
    const int wat_top = a.Water[qtop], open_top = a.Open[qtop];
    const int o_jm2 = a.Open[q + ojm2], o_jm1 = a.Open[q + ojm1], o_jp1 = a.Open[q + ojp1], o_jp2 = a.Open[q + ojp2];
    const int o_im2 = a.Open[q + oim2], o_im1 = a.Open[q + oim1], o_ip1 = a.Open[q + oip1], o_ip2 = a.Open[q + oip2];
    const int o_km2 = a.Open[q + okm2], o_km1 = a.Open[q + okm1], o_kp1 = a.Open[q + okp1], o_kp2 = a.Open[q + okp2];
    const double V = a.VolumeZ[q], Vold = a.VolumeZOld[q];
    const double V_w = a.VolumeZ[q + ojm1], V_s = a.VolumeZ[q + oim1], V_b = a.VolumeZ[q + okm1];
    const double dwz = a.DWZ[q], dwz_m = a.DWZ[q + okm1], dwz_m2 = a.DWZ[q + okm2], dwz_p = a.DWZ[q + okp1];
    const double Qx = a.Wflux_X[q], Qx_e = a.Wflux_X[q + ojp1], Qy = a.Wflux_Y[q], Qz = a.Wflux_Z[q];
    const double visc = a.Visc_H[q], visc_w = a.Visc_H[q + ojm1], visc_s = a.Visc_H[q + oim1];
    const double areau = a.AreaU[q], areav = a.AreaV[q], diffv = a.Diff_V[q], dzz_m = a.DZZ[q + okm1];
    const double dux = a.DUX[q2], dux_w = a.DUX[q2 - (jm1 ? sj2 : 0)], dvy = a.DVY[q2], dvy_s = a.DVY[q2 + oim1];
    const double dzx_w = a.DZX[q2 - (jm1 ? sj2 : 0)], dzy_s = a.DZY[q2 + oim1];
    const int small = a.SmallDepths ? a.SmallDepths[q2] : 0;
    const double rUp = A.rhoUp[q2], rUn = A.rhoUn[q2], rVp = A.rhoVp[q2], rVn = A.rhoVn[q2];

    // ---- mask: identical to adt_coef_kernel, plus the four direction bits ----
    const bool cfu = cfu_c == 1, cfv = cfv_c == 1, cfw = cfw_c == 1, bnd = bnd_c == 1;
    const bool ojm1b = jm1 && o_jm1 == 1, ojp1b = jp1 && o_jp1 == 1, oim1b = im1 && o_im1 == 1, oip1b = ip1 && o_ip1 == 1;
    const bool ojm2b = jm2 && o_jm2 == 1, ojp2b = jp2 && o_jp2 == 1, oim2b = im2 && o_im2 == 1, oip2b = ip2 && o_ip2 == 1;
    const bool okm1b = km1 && o_km1 == 1, okp1b = kp1 && o_kp1 == 1, okp2b = kp2 && o_kp2 == 1, okm2b = km2 && o_km2 == 1;
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
    if (ojm2b) m |= M_O_JM2;
    if (ojm1b) m |= M_O_JM1;
    if (ojp1b) m |= M_O_JP1;
    if (ojp2b) m |= M_O_JP2;
    if (oim2b) m |= M_O_IM2;
    if (oim1b) m |= M_O_IM1;
    if (oip1b) m |= M_O_IP1;
    if (oip2b) m |= M_O_IP2;
    if (okm1b) m |= M_O_KM1;
    if (okp1b) m |= M_O_KP1;
    if (okp2b) m |= M_O_KP2;
    if (bnd) {
        if (oip1b && a.Bnd[q2 + 1] != 1) m |= M_A_IP1;
        if (oim1b && a.Bnd[q2 - 1] != 1) m |= M_A_IM1;
        if (ojp1b && a.Bnd[q2 + sj2] != 1) m |= M_A_JP1;
        if (ojm1b && a.Bnd[q2 - sj2] != 1) m |= M_A_JM1;
    }
    const bool pos_u = Qx > 0., pos_v = Qy > 0., pos_b = Qz > 0.;
    if (pos_u) m |= M_POS_W;
    if (jp1 && Qx_e > 0.) m |= M_POS_E;
    if (pos_v) m |= M_POS_S;
    if (pos_b) m |= M_POS_B;

    // ---- DT/V of the cell and of the upwind cell of each face (same expression as adt_coef_kernel) ----
    const bool kin = k >= 1 && k <= a.K, jin = j >= 1 && j <= a.J, iin = i >= 1 && i <= a.I;
    auto dt_over = [&](bool inwork, double v) { return (inwork && v != 0.) ? a.dt / v : 0.; };
    const bool inwork = iin && jin && kin;
    const double dtv = dt_over(inwork, V);
    // closed cells keep their concentration (AD:4003-4006): Vold/V only where the cell is open
    const double vr = (open_c == 1 && inwork && V != 0.) ? Vold / V : 1.;
    // the face advects iff both cells are open and it is a compute face (MF:10559, AD:4467, 4833, 3041); vertical
    // advection also needs an open surface cell in the column (AD:2966)
    const double Qa_u = (cfu && open_c == 1 && ojm1b) ? Qx : 0.;
    const double Qa_v = (cfv && open_c == 1 && oim1b) ? Qy : 0.;
    const double qa_b = (cfw && open_c == 1 && okm1b && open_top == 1) ? Qz : 0.;
    double hc_u = 0., hc_v = 0., hv_b = 0.;
    if (A.tvd) {
        // upwind cell: (j-1 | j), (i-1 | i), (k-1 | k) for positive | non-positive flux (MF:10724-10736)
        const double t_u = pos_u ? dt_over(iin && kin && j >= 2 && j <= a.J + 1, V_w) : dtv;
        const double t_v = pos_v ? dt_over(jin && kin && i >= 2 && i <= a.I + 1, V_s) : dtv;
        const double t_b = pos_b ? dt_over(iin && jin && k >= 2 && k <= a.K + 1, V_b) : dtv;
        const bool near_u = pos_u ? !ojm2b : !ojp1b, near_v = pos_v ? !oim2b : !oip1b, near_b = pos_b ? !okm2b : !okp1b;
        hc_u = (near_u && A.upwind2_h) ? 0. : fma(-0.5 * Qx, t_u, 0.5);
        hc_v = (near_v && A.upwind2_h) ? 0. : fma(-0.5 * Qy, t_v, 0.5);
        hv_b = (near_b && A.upwind2_v) ? 0. : 0.5 * (1. - Qz * t_b);
    }
    // ---- diffusion constants: same operation order as adt_coef_kernel (AD:2486-2495, 1549-1597) ----
    double hu = 0., hv = 0., vz = 0.;
    if (cfu && jm1) {
        double difx = a.schmidt_h * (visc * dux_w + visc_w * dux) / (dux + dux_w);
        if (a.nulldif && Qx == 0.) difx = 0.;
        hu = difx * areau / dzx_w;
    }
    if (cfv && im1) {
        double dify = a.schmidt_h * (visc * dvy_s + visc_s * dvy) / (dvy + dvy_s);
        if (a.nulldif && Qy == 0.) dify = 0.;
        hv = dify * areav / dzy_s;
    }
    if (cfw && km1 && small == 0) {
        double difz = (a.schmidt_coef_v * diffv + a.schmidt_bg_v);
        if (a.nulldif_v && Qz == 0.) difz = 0.;
        const double auxk = difz * dux * dvy;
        vz = auxk / dzz_m;
    }
    // ---- 1/(DWZ+DWZ) of the bottom face and of its upwind pair ----
    auto rcp_sum = [](bool ok, double x, double y) { const double sd = ok ? x + y : 0.; return sd != 0. ? 1.0 / sd : 0.; };
    const double rdc_b = rcp_sum(km1, dwz, dwz_m);                                   // rdz(k)
    // rdz(k-1) for upward flow; rdz(k+1) for downward flow, clamped to plane K+1 like the round-1 look-ahead
    const double rdu_b = pos_b ? rcp_sum(km2, dwz_m, dwz_m2)
                               : ((k + 1 <= a.K + 1) ? rcp_sum(kp1, dwz_p, dwz) : rdc_b);

    Pack4 *g = A.pk + qc;
    st_pack(g + PK_U, 0.5 * Qa_u, fma(0.5, fabs(Qa_u), hu), fabs(Qa_u) * hc_u, pos_u ? rUp : rUn);
    st_pack(g + PK_V, 0.5 * Qa_v, fma(0.5, fabs(Qa_v), hv), fabs(Qa_v) * hc_v, pos_v ? rVp : rVn);
    st_pack(g + PK_W, qa_b, hv_b, rdu_b, rdc_b);
    st_pack(g + PK_C, dtv, vr, vz, __hiloint2double(0, (int)m));
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) adt_lean_coef_kernel(const LeanCoefArgs A) {
    const CoefArgs &a = A.c;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y, j = A.jc0 + blockIdx.z;
    if (i >= A.nt32 * 32) return;
    if (i >= a.ni) {                                  // row padding
        Pack4 *g = A.pk + pack_index(i, (j - A.jc0) + A.ncol * k, A.nt32);
        st_pack(g + PK_U, 0., 0., 0., 0.); st_pack(g + PK_V, 0., 0., 0., 0.);
        st_pack(g + PK_W, 0., 0., 0., 0.); st_pack(g + PK_C, 0., 1., 0., 0.);
        return;
    }
    const bool interior = i >= 2 && i + 2 < a.ni && j >= 2 && j + 2 < a.nj && k >= 2 && k + 2 < a.nk;
    if (interior) adt_lean_coef_cell<false>(A, i, j, k);
    else adt_lean_coef_cell<true>(A, i, j, k);
}

// -------------------------------------------------------------------------------------
// Step kernel
// -------------------------------------------------------------------------------------
struct LeanArgs {
    int I, J, K, ld, sj, sk;                    // property / raw arrays: element (i,j,k) at i + sj*j + sk*k
    int nprop, ntile_i, j_begin, j_count;
    int jc0, ncol, nt32;                        // pack array (see LeanCoefArgs)
    double dt;
    const Pack4 *pk;
    const double *qx, *qy, *qz, *VolumeZ, *VolumeZOld;   // raw: surface row of VolumeVariation, open-boundary flux
    unsigned long long *zero_pivots;
    PropArgs p[NPMAX];
};

// a < b ? a : b and a > b ? a : b as one compare + select (written in PTX so that no NaN-aware min/max sequence is formed)
__device__ __forceinline__ double sel_lt(double a, double b) {
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b));
    return r;
}
__device__ __forceinline__ double sel_gt(double a, double b) {
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b));
    return r;
}
__device__ __forceinline__ double abs_bits(double x) { return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x)); }
// x >= 0 ? x : 0 for the sign taken from another word (s < 0 -> 0), then the sign bit `sg` put on: integer pipe only
__device__ __forceinline__ double zero_if_neg_then_sign(double x, int s_hi, int sg) {
    const int keep = ~(s_hi >> 31);
    return __hiloint2double((__double2hiint(x) & keep) | sg, __double2loint(x) & keep);
}

// advective - diffusive flux through a horizontal face, positive toward the higher index, in the direction-free form
// Qh (Plo + Phi) - A g + B sgn(g) lim (see the header).  g = Phi - Plo, dm / dp the same difference one cell to the
// lower / higher side, f = {Qh, A, B, rho}.
template <int M>
__device__ __forceinline__ double lean_face(const bool pos, const double dm, const double g, const double dp,
                                            const double Plo, const double Phi, const Pack4 &f) {
    double F = fma(-f.b, g, f.a * (Plo + Phi));
    if constexpr (M == MOHID_P2_TVD) {
        // Division-free SuperBee (see hface_flux): psi(r) |dP| = max(0, min(hi, 2 lo)) with lo / hi the smaller / larger
        // of |g| and aS = sgn(g) (upwind difference) rho -- the same expression for both flow directions
        const double x = (pos ? dm : dp) * f.d;
        const int gs = __double2hiint(g) & 0x80000000;
        const double aS = __hiloint2double(__double2hiint(x) ^ gs, __double2loint(x));
        const double ad = abs_bits(g);
        const bool c1 = ad < aS;
        const double lo = c1 ? ad : aS, hi = c1 ? aS : ad;
        const double mn = sel_lt(hi, lo + lo);
        F = fma(f.c, zero_if_neg_then_sign(mn, __double2hiint(aS), gs), F);
    }
    return F;
}

// Open-boundary row (AD:5369-5672): the row of cell q is amended and eliminated again; returns (W, G).
__device__ __forceinline__ double2 lean_obc_row(const LeanArgs &s, const PropArgs &pa, int q, unsigned m, double Pc,
                                             double dtv_c, double D, double E, double F, double TI, double Wp0, double Gp0) {
    Row row{D, E, F, TI};
    open_boundary_row<false>(s, pa, q, m, Pc, s.qz[q], s.qz[q + s.sk], dtv_c, row);
    const double aux = row.E + row.D * Wp0;
    if (aux != 0.) {
        const double ra = 1.0 / aux;
        return make_double2(-row.F * ra, (row.TI - row.D * Gp0) * ra);
    }
    return make_double2(Wp0, Gp0);
}

struct LeanLevel {
    Pack4 U0, U1, V, C, W;
    double Pw2, Pw1, Pe1, Pe2, hP;
};

// PFD > 0: the loads that miss L2 (the east-most property row and the packs of this column and of column j+1, which no
// earlier block has touched) are requested PFD levels ahead with prefetch.global.L2 -- one instruction per warp and
// level for the property row, two more by the first property's warp for the packs all warps of the block share.
template <int M, int WARPS, int PFD = 0>
__global__ void __launch_bounds__(WARPS * 32, 1) adt_transport_lean_kernel(const __grid_constant__ LeanArgs s) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int wstride = WARPS * 32;
    const long nunits = (long)s.nprop * s.ntile_i * s.j_count;
    const long unit = (long)blockIdx.x * WARPS + warp;
    if (unit >= nunits) return;
    const int n = (int)(unit % s.nprop);
    const int tile = (int)((unit / s.nprop) % s.ntile_i);
    const int j = (int)(unit / ((long)s.nprop * s.ntile_i)) + s.j_begin;
    const int i = 1 + tile * 31 + lane;
    const bool writer = (lane < 31) && (i <= s.I);
    const int ic = min(i, s.I + 1);
    const PropArgs &pa = s.p[n];
    const double *__restrict__ P = pa.pin;
    const int sj = s.sj, sk = s.sk;
    const long cp = ic + (long)sj * j;                    // column base (k = 0) in the property arrays
    const int je2 = (j + 2 <= s.J + 1) ? 2 * sj : sj;
    const int jw2 = (j >= 2) ? 2 * sj : sj;
    const long pk_plane = (long)s.ncol * s.nt32 * 128;    // Pack4 elements per k-plane of packs
    const long pk_col = (long)s.nt32 * 128;               // ... per column

    const Pack4 *__restrict__ pk0 = s.pk + pack_index(ic, j - s.jc0, s.nt32);      // (i, j, k = 0)
    // from a lane's pack pointer to the start of the 32-cell group that holds the strip's first cell
    const int goff = -(ic & 31) - ((ic >> 5) - ((1 + tile * 31) >> 5)) * 128;
    const unsigned mtop = (unsigned)__double2loint(pk0[pk_plane * s.K + PK_C].d);
    const bool colwet = (mtop & M_COLWET) != 0;
    const bool obc = (mtop & M_BND) != 0 && pa.bc != MOHID_BC_None;
    const double theta = pa.theta_difv, omt = 1. - pa.theta_difv;
    const double qz_top = s.qz[cp + (long)sk * (s.K + 1)];
    const bool halo_lane = (lane < 2) || (lane == 31);
    const int halo_off = (lane == 31) ? ((ic <= s.I) ? 1 : 0) : -2;

    // packs: one per-thread pointer (constant offsets reach the four packs); properties: one cell index
    const Pack4 *__restrict__ pk = pk0 + pk_plane;                                // packs of level 1
    int qp = (int)cp + sk;                                                       // cell (i, j, 1)
    double *__restrict__ wsm = smem + warp * 32 + lane;                          // [K][WARPS][32]

    auto fetch = [&](LeanLevel &L) {             // packs and horizontal neighbours of the level pk / qp are at
        L.C = ld_pack(pk + PK_C);
        L.U0 = ld_pack(pk + PK_U);
        L.U1 = ld_pack(pk + pk_col + PK_U);
        L.V = ld_pack(pk + PK_V);
        L.W = ld_pack(pk + PK_W);
        L.Pw2 = ld_f64(P + (qp - jw2)); L.Pw1 = ld_f64(P + (qp - sj)); L.Pe1 = ld_f64(P + (qp + sj)); L.Pe2 = ld_f64(P + (qp + je2));
        L.hP = halo_lane ? ld_f64(P + (qp + halo_off)) : 0.;
    };
    auto advance = [&]() { pk += pk_plane; qp += sk; };

    // ---- rolling state ----
    double Pm2 = 0., Pm1 = P[cp], Pc = P[qp], Pp1 = P[qp + sk];
    double dtv_m = 0.;
    double RD = 0., RE = 1., RTI = 0.;                    // row k-1 as far as it is known
    bool land_m = false, obc_m = false;                   // ... its land flag, and whether it is an open-boundary row
    unsigned m_m = 0;
    double Wp = 0., Gp = 0.;                              // W, G of the last eliminated row
    unsigned zp = 0;
    LeanLevel lvA, lvB;
    fetch(lvA);

    // the face below cell k (between k-1 and k): completes row k-1, eliminates it, returns the face's share of row k
    auto vface = [&](const int k, const Pack4 &C, const Pack4 &W, const unsigned m, double &Dn, double &En, double &TIn) {
        const double dtv_c = C.a;
        const double aux1 = C.c * dtv_m, aux2 = C.c * dtv_c;          // diffusion (AD:2708-2937)
        const double dP = Pc - Pm1;
        RE = fma(aux1, theta, RE);
        double RF = -(aux1 * theta);
        RTI = fma(aux1 * dP, omt, RTI);
        Dn = -(aux2 * theta);
        En = aux2 * theta;
        TIn = -(aux2 * dP) * omt;
        // advection, implicit (AD:2941-3144); weights from the old field (AD:2966-3001)
        const bool pos = (m & M_POS_B) != 0;
        double w1, w2;                                                 // weights of cell k-1 and of cell k
        if constexpr (M == MOHID_P2_TVD) {
            // r = ((Pu - Puu) rd_u) / ((Pd - Pu) rd_c) with dC = (Pd - Pu) rd_c kept away from zero (MF:10795-10803).  For
            // downward flow numerator and denominator are the negated differences; Pd - Pu is formed by its own
            // subtraction so that equal neighbours give +0 in both directions, as in the reference.
            const double num = pos ? Pm1 - Pm2 : Pc - Pp1;
            const double dPd = pos ? dP : Pm1 - Pc;
            double dC = dPd * W.d;
            dC = (abs_bits(dC) < MIN_VALUE) ? with_sign_of(MIN_VALUE, dC) : dC;
            const double r = num * W.c * fast_rcp(dC);
            // SuperBee max(0, min(1, 2r), min(r, 2)) (MF:10826-10829) as a chain of exact selections
            double ps = sel_lt(r + r, 1.);
            ps = sel_gt(ps, r);
            ps = sel_lt(ps, 2.);
            ps = zero_if_neg_then_sign(ps, __double2hiint(r), 0);
            const double th = ps * W.b;                                // 0.5 psi (1 - Cr), 0 near the boundary
            const double wu = 1. - th;
            w1 = pos ? wu : th; w2 = pos ? th : wu;
        } else {
            w1 = pos ? 1. : 0.; w2 = pos ? 0. : 1.;
        }
        const double dfl = W.a * w1, efl = W.a * w2;                   // D_flux, E_flux (MF:10583-10586)
        RE = fma(dfl, dtv_m, RE);
        RF = fma(efl, dtv_m, RF);
        Dn = fma(-dfl, dtv_c, Dn);
        En = fma(-efl, dtv_c, En);
        // land fill (AD:1753) and forward elimination of row k-1 (MF:4087-4099); a zero pivot keeps the previous W, G
        RTI = land_m ? NULL_REAL : RTI;
        const double Wp0 = Wp, Gp0 = Gp;
        {
            const double aux = fma(RD, Wp0, RE);
            const bool ok = aux != 0.;
            const double ra = fast_rcp(aux);
            Wp = ok ? -RF * ra : Wp0;
            Gp = ok ? (RTI - RD * Gp0) * ra : Gp0;
            zp += ok ? 0u : 1u;
        }
        if (obc_m) {                                                   // open-boundary row (rare): amend, eliminate again
            const double2 wg = lean_obc_row(s, pa, (int)(cp + (long)sk * (k - 1)), m_m, Pm1, dtv_m, RD, RE, RF, RTI, Wp0, Gp0);
            Wp = wg.x; Gp = wg.y;
        }
        wsm[0] = Wp;
        wsm += wstride;
        if (writer && colwet) pa.pout[cp + (long)sk * (k - 1)] = Gp;   // G parked in the output array
    };

    // POS: 0 = level 1 (no face below), 1 = levels 2 .. K-1, 2 = level K (surface row of VolumeVariation)
    auto level = [&](auto pos_tag, const int k, const LeanLevel &cur, LeanLevel &nxt) {
        constexpr int POS = decltype(pos_tag)::value;
        // ---- look-ahead: packs and horizontal neighbours of level k+1, P of level k+2 ----
        advance();
        fetch(nxt);
        if constexpr (PFD > 0) {
            const int kk = min(k + PFD, s.K + 1) - k - 1;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P + (qp + kk * sk + je2)));
            if (n == 0) {
                // the strip's cells lie in two 32-cell groups of 4 KB each: lane l requests line l of both (column j)
                // and, for column j+1, the lines of the two U packs (lanes 0-7 / 8-15)
                const Pack4 *f = pk + (long)kk * pk_plane + goff + lane * 4;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(f));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(f + 128));
                if (lane < 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(f + pk_col + (lane >> 3) * (128 - 32)));
            }
        }
        const double Pp2 = ld_f64(P + (POS == 2 ? qp : qp + sk));     // plane K+2 does not exist: clamped like round 1
        const unsigned m = (unsigned)__double2loint(cur.C.d);
        const double dtv_c = cur.C.a;

        // ---- horizontal faces (explicit) ----
        const double d1 = cur.Pw1 - cur.Pw2, d2 = Pc - cur.Pw1, d3 = cur.Pe1 - Pc, d4 = cur.Pe2 - cur.Pe1;
        const double fw = lean_face<M>((m & M_POS_W) != 0, d1, d2, d3, cur.Pw1, Pc, cur.U0);
        const double fe = lean_face<M>((m & M_POS_E) != 0, d2, d3, d4, Pc, cur.Pe1, cur.U1);
        // south face of every lane; the north face is the south face of lane + 1
        double Ps1 = shfl_up_d(Pc, 1);
        const double hP1 = __shfl_sync(0xffffffffu, cur.hP, 1);
        Ps1 = lane == 0 ? hP1 : Ps1;
        const double gs = Pc - Ps1;
        double dsm = shfl_up_d(gs, 1), dsp = shfl_dn_d(gs, 1);
        dsm = lane == 0 ? Ps1 - cur.hP : dsm;
        dsp = lane == 31 ? cur.hP - Pc : dsp;
        const double fs = lean_face<M>((m & M_POS_S) != 0, dsm, gs, dsp, Ps1, Pc, cur.V);
        const double fsum = (fw - fe) + (fs - shfl_dn_d(fs, 1));

        // ---- the face below: row k-1 is complete, eliminate it ----
        double Dn = 0., En = 0., TIn = 0.;
        if constexpr (POS != 0) vface(k, cur.C, cur.W, m, Dn, En, TIn);

        // ---- row k: VolumeVariation (AD:3966-4021; Vold/V is 1 in closed cells) + the shares known so far ----
        double e0 = 1.0;
        if constexpr (POS == 2) e0 = (m & M_OPEN) ? 1.0 + dtv_c * qz_top : 1.0;
        RTI = fma(fsum, dtv_c, Pc * cur.C.b + TIn);
        RE = e0 + En;
        RD = Dn;
        land_m = (m & M_LAND) != 0;
        obc_m = obc && (m & M_OPEN) != 0;
        m_m = m;
        // ---- roll ----
        Pm2 = Pm1; Pm1 = Pc; Pc = Pp1; Pp1 = Pp2;
        dtv_m = dtv_c;
    };
    {
        const std::integral_constant<int, 0> first{};
        const std::integral_constant<int, 1> mid{};
        const std::integral_constant<int, 2> last{};
        // level 1 (or the only level of a K = 1 ... not reachable: the lean path needs K >= 2)
        level(first, 1, lvA, lvB);
        int nmid = s.K - 2;                                 // levels 2 .. K-1
        int k = 2;
        if (nmid & 1) { level(mid, k, lvB, lvA); lvB = lvA; ++k; --nmid; }
        for (; nmid > 0; nmid -= 2, k += 2) {
            level(mid, k, lvB, lvA);
            level(mid, k + 1, lvA, lvB);
        }
        level(last, s.K, lvB, lvA);
        // ---- virtual level K+1: the face above the surface cell completes row K ----
        double Dn, En, TIn;
        vface(s.K + 1, lvA.C, lvA.W, (unsigned)__double2loint(lvA.C.d), Dn, En, TIn);
    }

    // ---------------- back substitution (MF:4100-4105) ----------------
    if (writer && colwet) {
        double *__restrict__ O = pa.pout;
        long qo = cp + (long)sk * (s.K + 1);
        const double *__restrict__ Wsm = smem + warp * 32 + lane;
        double x = 0.0;                                   // RES(KUB+1) = G(KUB+1) = 0 (halo row is the identity)
        O[qo] = x;
        int k = s.K;
        // G comes back from the output array (L2): eight levels per batch, the next batch in flight while this one is solved
        double ga[8], gb[8];
        auto load8 = [&](double (&g)[8], long q0) {
#pragma unroll
            for (int u = 0; u < 8; ++u) g[u] = O[q0 - (long)(u + 1) * sk];
        };
        auto solve8 = [&](const double (&g)[8]) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                qo -= sk;
                x = Wsm[(k - 1 - u) * wstride] * x + g[u];
                O[qo] = x;
            }
            k -= 8;
        };
        if (k >= 8) load8(ga, qo);
        while (k >= 16) {
            load8(gb, qo - 8L * sk);
            solve8(ga);
            if (k >= 16) { load8(ga, qo - 8L * sk); solve8(gb); }
            else { solve8(gb); goto tail; }
        }
        if (k >= 8) solve8(ga);
    tail:
        for (; k >= 1; --k) {
            qo -= sk;
            x = Wsm[(k - 1) * wstride] * x + O[qo];
            O[qo] = x;
        }
        if (zp) atomicAdd(s.zero_pivots, (unsigned long long)zp);
    }
}

}  // namespace adt
