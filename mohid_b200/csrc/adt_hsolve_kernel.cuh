// K6: horizontally implicit advection (ImpExp_AdvXX = 1 or ImpExp_AdvYY = 1; direction splitting of AD:4132-4265).
//
// Stage 1 of the split step: per (cross cell, level k, property) the explicit terms of the row (VolumeVariation,
// discharges, horizontal diffusion, explicit advection of the other horizontal direction) are built on the fly,
// the implicit direction contributes D_flux / E_flux to the tridiagonal system (AD:4483-4509 / 4862-4893) and the
// system is solved along the line with the recurrence of THOMAS_3D_i0_j1 / _i1_j0 (MF:3751-3875).  The result is
// the intermediate field from which stage 2 (adt_transport_kernel with StepArgs::stage2) restarts the system for the
// vertical terms (AD:4250-4253: D = 0, E = 1, F = 0, TI = PROP).
//
//   warp <-> (31-cell strip across the line direction, level k, property n); lane 31 only supplies the flux of the
//   far cross face of lane 30 (as in adt_transport_kernel).  DIR = 0: lines along j (XX implicit), lanes along i
//   (coalesced).  DIR = 1: lines along i (YY implicit), lanes along j (strided loads: this is a rarely used
//   option, off by default, WP:9565, and the reference serves it by gathering whole rows on the master).
//   G of the recurrence is parked in the output array, W in a per-property scratch array.
//
// StepArgs::twod (K = 1): the line solve is the whole step, see below.
//
// A zero pivot stops the reference ('Instability in THOMAS3D', MF:3799); here it is counted in zero_pivots.
#pragma once

namespace adt {

struct HSolveArgs {
    double *wline[NPMAX];            // W of the line recurrence, one scratch field per property of the launch
    // Lines along j on a column slab (the reference gathers such rows on one process, THOMAS_DDecompHorizGrid HG:8245-8478):
    // the recurrence runs over the owned cells l0 .. l1 only, takes (W, G) of cell l0 - 1 from the rank on the left and
    // hands (W, G) of cell l1 to the rank on the right; adt_hsolve_back_kernel then substitutes back from the right.
    // State per (property, level, cross cell): edge[((n * K + k - 1) * NC + c - 1) * 2 + {0, 1}].
    int split;                       // 0: the whole line in one launch (forward and back substitution)
    int l0, l1;
    const double *edge_in;           // nullptr on the first rank
    double *edge_out;
};

template <int DIR>
__global__ void __launch_bounds__(256) adt_hsolve_kernel(const __grid_constant__ StepArgs s, const __grid_constant__ HSolveArgs hs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, WPB = blockDim.x >> 5;
    const int NC = DIR == 0 ? s.I : s.J;                   // cells across the lines
    const int NL = DIR == 0 ? s.J : s.I;                   // cells along a line
    const int ntile = (NC + 30) / 31;
    const long nunits = (long)s.nprop * ntile * s.K;
    const long unit = (long)blockIdx.x * WPB + warp;
    if (unit >= nunits) return;
    const int n = (int)(unit % s.nprop);
    const int tile = (int)((unit / s.nprop) % ntile);
    const int k = (int)(unit / ((long)s.nprop * ntile)) + 1;
    const int c = 1 + tile * 31 + lane;
    const bool writer = (lane < 31) && (c <= NC);
    const int cc = min(c, NC + 1);                         // clamped cross index: every load stays in bounds
    const PropArgs pa = s.p[n];
    const double *__restrict__ P = pa.pin;
    double *__restrict__ O = pa.pout;
    double *__restrict__ Wl = hs.wline[n];
    const int sj = s.sj, sk = s.sk, ld2 = s.ld;
    const int sl = DIR == 0 ? sj : 1, sc = DIR == 0 ? 1 : sj;             // 3-D strides along / across the line
    const int sl2 = DIR == 0 ? ld2 : 1, sc2 = DIR == 0 ? 1 : ld2;         // 2-D strides
    const double *__restrict__ qL = DIR == 0 ? s.qx : s.qy, *__restrict__ qC = DIR == 0 ? s.qy : s.qx;
    const double *__restrict__ dhL = DIR == 0 ? s.dhu : s.dhv, *__restrict__ dhC = DIR == 0 ? s.dhv : s.dhu;
    const double *__restrict__ rdL = DIR == 0 ? s.rdx : s.rdy, *__restrict__ rdC = DIR == 0 ? s.rdy : s.rdx;
    const double *__restrict__ duL = DIR == 0 ? s.DUX : s.DVY, *__restrict__ duC = DIR == 0 ? s.DVY : s.DUX;
    constexpr unsigned CF_L = DIR == 0 ? M_CFU : M_CFV, CF_LN = DIR == 0 ? M_CFUE : M_CFVN;
    constexpr unsigned O_LM2 = DIR == 0 ? M_O_JM2 : M_O_IM2, O_LM1 = DIR == 0 ? M_O_JM1 : M_O_IM1;
    constexpr unsigned O_LP1 = DIR == 0 ? M_O_JP1 : M_O_IP1, O_LP2 = DIR == 0 ? M_O_JP2 : M_O_IP2;
    constexpr unsigned CF_C = DIR == 0 ? M_CFV : M_CFU;
    constexpr unsigned O_CM2 = DIR == 0 ? M_O_IM2 : M_O_JM2, O_CM1 = DIR == 0 ? M_O_IM1 : M_O_JM1;
    constexpr unsigned O_CP1 = DIR == 0 ? M_O_IP1 : M_O_JP1;
    constexpr unsigned NF_LW = DIR == 0 ? NF_WEST : 0u, NF_LE = DIR == 0 ? NF_EAST : 0u;   // NoFlux bits act on U faces only
    const bool cross_on = !(DIR == 0 && s.xzflow);        // XZFlow: no YY terms (AD:4163)
    const bool line_on = !(DIR == 1 && s.xzflow);         // ... also when YY is the implicit direction: identity rows along i
    const unsigned nfsel = pa.nfsel;

    // implicit face between line cells a-1 and a (cells a-2 .. a+1 = P1 .. P4): D_flux, E_flux (MF:10583-10586)
    auto line_face = [&](bool on, double Q, double P1, double P2, double P3, double P4, bool o1, bool o4, double t1,
                         double t2, double t3, double t4, double rd12, double rd23, double rd34, double du2, double du3,
                         double &dfl, double &efl) {
        const bool pos = Q > 0.;
        const double Puu = sel(pos, P1, P4), Pu = sel(pos, P2, P3), Pd = sel(pos, P3, P2);
        double wuu, wu, wd;
        oriented_weights<0, 0>(s.method_h, s.limiter_h, s.upwind2_h != 0, s.vrelmax, Q, Puu, Pu, Pd, pos ? !o1 : !o4,
                               sel(pos, t1, t4), sel(pos, t2, t3), sel(pos, t3, t2), sel(pos, rd12, rd34), rd23,
                               sel(pos, du2, du3), sel(pos, du3, du2), wuu, wu, wd);
        const double qa = (on && line_on) ? Q : 0.;
        dfl = qa * sel(pos, wu, wd);
        efl = qa * sel(pos, wd, wu);
    };

    const int c2 = DIR == 0 ? cc : cc * ld2;               // 2-D index of (cross = cc, line = 0)
    const int q0 = DIR == 0 ? cc + sk * k : cc * sj + sk * k;   // 3-D index of (cross = cc, line = 0, level k)
    // cross-direction metrics (constant along... no: they vary along the line) are read per cell below
    double Wprev = 0., Gprev = 0.;
    double dfl_w = 0., efl_w = 0.;
    unsigned zp = 0;
    const int L0 = hs.split ? hs.l0 : 1, L1 = hs.split ? hs.l1 : NL;
    const long eidx = (((long)n * s.K + (k - 1)) * NC + (min(c, NC) - 1)) * 2;
    if (hs.split && hs.edge_in) { Wprev = hs.edge_in[eidx]; Gprev = hs.edge_in[eidx + 1]; }
    if (L0 > 1) {   // the face between cells L0 - 1 and L0, as cell L0 - 1 (a ghost cell) sees its far face in the loop below
        const int l = L0 - 1;
        const int q = q0 + sl * l, p2 = c2 + sl2 * l;
        const int lp2 = (l + 2 <= NL + 1) ? 2 * sl : sl, lp2_2 = (l + 2 <= NL + 1) ? 2 * sl2 : sl2;
        const unsigned m = s.mask[q];
        const unsigned nf = nfsel ? (s.nfmask[q] & nfsel) : 0u;
        line_face(all_set(m, CF_LN | O_LP1 | M_OPEN) && !(nf & NF_LE), qL[q + sl], P[q - sl], P[q], P[q + sl], P[q + lp2],
                  (m & O_LM1) != 0, (m & O_LP2) != 0, s.dtv[q - sl], s.dtv[q], s.dtv[q + sl], s.dtv[q + lp2], rdL[p2],
                  rdL[p2 + sl2], rdL[p2 + lp2_2], duL[p2], duL[p2 + sl2], dfl_w, efl_w);
    } else {   // west face of the first cell of the line (a = 1)
        const int q = q0 + sl, p2 = c2 + sl2;
        const unsigned m = s.mask[q];
        const unsigned nf = nfsel ? (s.nfmask[q] & nfsel) : 0u;
        line_face(all_set(m, CF_L | O_LM1 | M_OPEN) && !(nf & NF_LW), qL[q], P[q - sl], P[q - sl], P[q], P[q + sl], false,
                  (m & O_LP1) != 0, 0., s.dtv[q - sl], s.dtv[q], s.dtv[q + sl], rdL[p2 - sl2], rdL[p2], rdL[p2 + sl2],
                  duL[p2 - sl2], duL[p2], dfl_w, efl_w);
    }
    for (int l = L0; l <= L1; ++l) {
        const int q = q0 + sl * l, p2 = c2 + sl2 * l;
        const int lm2 = (l >= 2) ? 2 * sl : sl, lp2 = (l + 2 <= NL + 1) ? 2 * sl : sl;
        const int lp2_2 = (l + 2 <= NL + 1) ? 2 * sl2 : sl2;
        const int cm2 = (cc >= 2) ? 2 * sc : sc, cm2_2 = (cc >= 2) ? 2 * sc2 : sc2;
        const int cp1 = (cc <= NC) ? sc : 0, cp1_2 = (cc <= NC) ? sc2 : 0;
        const unsigned m = s.mask[q];
        const unsigned nf = nfsel ? (s.nfmask[q] & nfsel) : 0u;
        const bool open_c = (m & M_OPEN) != 0;
        const double Pc = P[q], Pw2 = P[q - lm2], Pw1 = P[q - sl], Pe1 = P[q + sl], Pe2 = P[q + lp2];
        const double dtv_c = s.dtv[q], t_w = s.dtv[q - sl], t_e = s.dtv[q + sl], t_e2 = s.dtv[q + lp2];
        const double vr = s.vr[q];
        // ---------------- VolumeVariation (AD:3966-4021), Discharges (AD:4025-4128) ----------------
        double ti = sel(open_c, Pc * vr, Pc);
        double e0 = sel(open_c && k == s.K, 1.0 + dtv_c * s.qz[q + sk], 1.0);
        if ((m & M_DISCH) && pa.dconc)
            apply_discharges(s.disch, pa.dconc, pa.dconcmf, DIR == 0 ? cc : l, DIR == 0 ? l : cc, k, open_c, Pc, vr, dtv_c, ti, e0);
        // ---------------- explicit terms: diffusion along the line, full flux across it ----------------
        double fsum = line_on ? -dhL[q] * (Pc - Pw1) + dhL[q + sl] * (Pe1 - Pc) : 0.;
        if (cross_on) {
            // the explicit face across the line; when that is a U face (lines along i) the NoFlux bits act on it as in K2
            const double fs = hface_flux<0, 0>(s, all_set(m, CF_C | O_CM1 | M_OPEN) && !(DIR == 1 && (nf & NF_WEST)), qC[q], dhC[q], P[q - cm2], P[q - sc], Pc,
                                               P[q + cp1], (m & O_CM2) != 0, (m & O_CP1) != 0, s.dtv[q - cm2], s.dtv[q - sc],
                                               dtv_c, s.dtv[q + cp1], rdC[p2 - sc2], rdC[p2], rdC[p2 + cp1_2], duC[p2 - sc2],
                                               duC[p2]);
            (void)cm2_2;
            fsum += fs - shfl_dn_d(fs, 1);
        }
        ti += fsum * dtv_c;
        // ---------------- implicit face l+1 (AD:4483-4509 / 4862-4893) ----------------
        double dfl_e, efl_e;
        line_face(all_set(m, CF_LN | O_LP1 | M_OPEN) && !(nf & NF_LE), qL[q + sl], Pw1, Pc, Pe1, Pe2, (m & O_LM1) != 0,
                  (m & O_LP2) != 0, t_w, dtv_c, t_e, t_e2, rdL[p2], rdL[p2 + sl2], rdL[p2 + lp2_2], duL[p2], duL[p2 + sl2],
                  dfl_e, efl_e);
        (void)Pw2; (void)O_LM2;
        double D = -dfl_w * dtv_c;
        double E = (e0 - efl_w * dtv_c) + dfl_e * dtv_c;
        double F = efl_e * dtv_c;
        if (s.twod) {
            // 2-D domain (K = 1, AD:1758-1841): there is no second stage; the open-boundary rows (AD:1745-1747) and the land
            // fill (AD:1753) enter the line system itself
            if ((m & M_BND) && open_c && pa.bc != MOHID_BC_None) {
                Row row{D, E, F, ti};
                open_boundary_row<true>(s, pa, q, m, Pc, s.qz[q], s.qz[q + sk], dtv_c, row, DIR == 0 ? cc : l, DIR == 0 ? l : cc,
                                        writer);
                D = row.D; E = row.E; F = row.F; ti = row.TI;
            }
            if (m & M_LAND) ti = NULL_REAL;
        }
        // ---------------- recurrence along the line (MF:3790-3801) ----------------
        const double aux = E + D * Wprev;
        if (aux != 0.) {
            const double Wn = -F / aux, Gn = (ti - D * Gprev) / aux;
            Wprev = Wn; Gprev = Gn;
        } else {
            zp++;
        }
        if (writer) { Wl[q] = Wprev; O[q] = Gprev; }
        dfl_w = dfl_e; efl_w = efl_e;
    }
    if (hs.split) {
        if (writer) {
            hs.edge_out[eidx] = Wprev; hs.edge_out[eidx + 1] = Gprev;
            if (zp) atomicAdd(s.zero_pivots, (unsigned long long)zp);
        }
        return;
    }
    if (writer) {
        // halo cell NL+1: identity row with TI = 0 (MF:3803); the reference writes it into PROP
        int q = q0 + sl * (NL + 1);
        double x = 0.;
        O[q] = x;
        const_cast<double *>(P)[q] = x;                     // both ping-pong buffers: stage 2 returns to this one
        for (int l = NL; l >= 1; --l) {
            q -= sl;
            x = Wl[q] * x + O[q];
            O[q] = x;
        }
        if (zp) atomicAdd(s.zero_pivots, (unsigned long long)zp);
    }
}

// Back substitution of a split line solve (lines along j): x of cell l1 + 1 comes from the rank on the right (last rank: the
// halo cell NL + 1, 0), x of cell l0 goes to the rank on the left.  One thread per (i, level, property), lanes along i.
struct HBackArgs {
    int I, K, sj, sk, nprop, l0, l1, last;
    const double *x_in;              // [n][k][i] x of cell l1 + 1; unused on the last rank
    double *x_out;                   // [n][k][i] x of cell l0
    double *out[NPMAX];              // G in, x out
    const double *wline[NPMAX];
    double *pin[NPMAX];              // last rank: the halo cell of the old field takes the 0 as well (see adt_hsolve_kernel)
};
__global__ void __launch_bounds__(128) adt_hsolve_back_kernel(const HBackArgs a) {
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x, k = 1 + blockIdx.y, n = blockIdx.z;
    if (i > a.I) return;
    const long e = ((long)n * a.K + (k - 1)) * a.I + (i - 1);
    double *__restrict__ O = a.out[n];
    const double *__restrict__ Wl = a.wline[n];
    long q = i + (long)a.sk * k + (long)a.sj * (a.l1 + 1);
    double x = 0.;
    if (a.last) { O[q] = x; a.pin[n][q] = x; }
    else x = a.x_in[e];
    for (int l = a.l1; l >= a.l0; --l) {
        q -= a.sj;
        x = Wl[q] * x + O[q];
        O[q] = x;
    }
    a.x_out[e] = x;
}

}  // namespace adt
