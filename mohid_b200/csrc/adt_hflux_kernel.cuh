// K8: the explicit horizontal part of the transport step on its own (opt-in split of K2, MOHID_ADT_HSPLIT=1).
//
// adt_transport_kernel computes the west AND the east U face of every cell (the warps of neighbouring columns cannot
// exchange them) and carries the whole column state in registers, which caps it at 12 warps per SM.  Here the
// horizontal fluxes have no column state: a block covers TJ consecutive columns j of one 31-cell i-strip and one
// property, warp w owns column jb + w and computes the WEST face of its cell and its SOUTH face (north = south face of
// lane + 1, as in K2); the east face is the west face of warp w + 1, handed over through shared memory, and one extra
// warp supplies the west face of column jb + TJ.  Every U face is therefore computed once per TJ/(TJ+1) cells, the
// neighbouring property rows are L1 hits (the neighbouring warps load them as their own), and registers allow 24
// warps per SM.  Output: the net horizontal flux into each cell; the HSPLIT variants of adt_transport_kernel multiply
// it by DT/V exactly where K2 does, so the split path is bit-identical to the fused one.
//
// Restrictions as for the headline K2 variants: FULL configuration, UpwindOrder1 or P2_TVD, no NoFlux lists.
#pragma once

namespace adt {

template <int MH, int LH, int TJ>
__global__ void __launch_bounds__((TJ + 1) * 32) adt_hflux_kernel(const __grid_constant__ StepArgs s) {
    __shared__ double Fx[2][TJ + 1][32];                   // west-face fluxes of the block's columns, by level parity
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int njg = (s.j_count + TJ - 1) / TJ;
    const long nunits = (long)s.nprop * s.ntile_i * njg;
    const long unit = blockIdx.x;
    if (unit >= nunits) return;
    const int n = (int)(unit % s.nprop);
    const int tile = (int)((unit / s.nprop) % s.ntile_i);
    const int jg = (int)(unit / ((long)s.nprop * s.ntile_i));
    const int jb = s.j_begin + jg * TJ;
    const int j_end = s.j_begin + s.j_count - 1;
    const int j = min(jb + w, s.J + 1);                    // this warp's column (the extra warp may sit on the halo column)
    const bool owner = w < TJ && jb + w <= j_end;          // produces output
    const int i = 1 + tile * 31 + lane;
    const bool writer = owner && lane < 31 && i <= s.I;
    const int ic = min(i, s.I + 1);
    const PropArgs pa = s.p[n];
    const double *__restrict__ P = pa.pin;
    const int sj = s.sj, sk = s.sk, sj2 = s.ld;
    const int c2 = ic + sj2 * j, c2d = ic + sj * j;
    const int jw2 = (j >= 2) ? 2 * sj : sj;               // column j-2, clamped (column -1 does not exist)
    const int je1 = (j + 1 <= s.J + 1) ? sj : 0;          // column j+1, clamped for the extra warp on the halo column
    const int je1_2 = (j + 1 <= s.J + 1) ? sj2 : 0;

    // ---- 2-D metrics of the column (as in adt_transport_kernel) ----
    const double rdx_m = s.rdx[c2 - sj2], rdx_c = s.rdx[c2], rdx_p = s.rdx[c2 + je1_2];
    const double rdy_c = s.rdy[c2];
    double rdy_m = shfl_up_d(rdy_c, 1), rdy_p = shfl_dn_d(rdy_c, 1);
    if (lane == 0) rdy_m = s.rdy[c2 - 1];
    if (lane == 31) rdy_p = s.rdy[c2 + (ic <= s.I ? 1 : 0)];
    constexpr bool FAST_H = (MH == MOHID_P2_TVD && LH == MOHID_SuperBee);
    const double rho_wp = FAST_H ? ratio_or_zero(rdx_m, rdx_c) : rdx_m, rho_wn = FAST_H ? ratio_or_zero(rdx_p, rdx_c) : rdx_p;
    const double rho_sp = FAST_H ? ratio_or_zero(rdy_m, rdy_c) : rdy_m, rho_sn = FAST_H ? ratio_or_zero(rdy_p, rdy_c) : rdy_p;
    const bool halo_lane = (lane < 2) || (lane == 31);
    const int halo_off = (lane == 31) ? ((ic <= s.I) ? 1 : 0) : -2;

    // inputs of one level; the next level is requested before the current one is used (software pipeline)
    struct In { unsigned m; double Pc, Pw1, Pw2, Pe1, dtv_c, t_w, qxw, dhw, qys, dhs, hP, t_h; };
    auto load = [&](int q, In &L) {
        L.m = __ldg(s.mask + q);
        L.Pc = __ldg(P + q); L.Pw1 = __ldg(P + q - sj); L.Pw2 = __ldg(P + q - jw2); L.Pe1 = __ldg(P + q + je1);
        L.dtv_c = __ldg(s.dtv + q); L.t_w = __ldg(s.dtv + q - sj);
        L.qxw = __ldg(s.qx + q); L.dhw = __ldg(s.dhu + q);
        L.qys = __ldg(s.qy + q); L.dhs = __ldg(s.dhv + q);
        L.hP = halo_lane ? __ldg(P + q + halo_off) : 0.;
        L.t_h = (lane == 0) ? __ldg(s.dtv + q - 1) : 0.;
    };
    int q = c2d + sk;                                       // cell (i, j, 1)
    In cur, nxt;
    load(q, cur);
    for (int k = 1; k <= s.K; ++k, q += sk) {
        load(q + sk, nxt);                                  // plane K+1 exists: always in bounds
        const unsigned m = cur.m;
        const double Pc = cur.Pc, dtv_c = cur.dtv_c;
        // ---- west face (U face j) ----
        const double fw = hface_flux<MH, LH>(s, all_set(m, M_CFU | M_O_JM1 | M_OPEN), cur.qxw, cur.dhw, cur.Pw2, cur.Pw1, Pc,
                                             cur.Pe1, (m & M_O_JM2) != 0, (m & M_O_JP1) != 0, 0., cur.t_w, dtv_c, 0., rho_wp,
                                             rdx_c, rho_wn, 0., 0.);
        Fx[k & 1][w][lane] = fw;
        // ---- south face (V face i); the north face is the south face of lane + 1 ----
        double Ps1 = shfl_up_d(Pc, 1), Ps2 = shfl_up_d(Pc, 2), Pn1 = shfl_dn_d(Pc, 1);
        double t_s = shfl_up_d(dtv_c, 1);
        const double hP1 = __shfl_sync(0xffffffffu, cur.hP, 1);
        Ps2 = sel(lane < 2, cur.hP, Ps2);
        Ps1 = sel(lane == 0, hP1, Ps1);
        Pn1 = sel(lane == 31, cur.hP, Pn1);
        t_s = sel(lane == 0, cur.t_h, t_s);
        const double fs = hface_flux<MH, LH>(s, all_set(m, M_CFV | M_O_IM1 | M_OPEN), cur.qys, cur.dhs, Ps2, Ps1, Pc, Pn1,
                                             (m & M_O_IM2) != 0, (m & M_O_IP1) != 0, 0., t_s, dtv_c, 0., rho_sp, rdy_c, rho_sn,
                                             0., 0.);
        const double fn = shfl_dn_d(fs, 1);
        __syncthreads();                                    // the west faces of level k are in Fx[k & 1]
        if (w < TJ) {
            const double fe = Fx[k & 1][w + 1][lane];
            const double fsum = (fw - fe) + (fs - fn);
            if (writer) pa.tih[q] = fsum;
        }
        // Fx[k & 1] is rewritten at level k + 2, i.e. after the barrier of level k + 1, which every reader of level k
        // has passed its read by
        cur = nxt;
    }
}

}  // namespace adt
