// =====================================================================================
//  adt_fused_kernel.cuh -- the whole transport step in ONE kernel (round 2, headline form).
//
//  Block <-> (31-cell strip along i, column j); its warps are specialised:
//    * warps 0 .. NC-1 ("property warps"): one property each.  They run the instruction stream of
//      adt_transport_lean_kernel (direction-free face form, skewed vertical faces, in-register Thomas elimination with
//      W in shared memory and G parked in the output array) but take the per-face packs from a shared-memory ring
//      instead of global memory;
//    * warps NC .. NC+3 ("coefficient warps"): walk the same column one to FR_D levels ahead of the property warps
//      and build the packs of adt_lean_kernel.cuh -- U0, U1 (west faces of columns j and j+1), V, and W + C, one warp
//      each -- straight from the raw interface arrays (Wflux_X/Y/Z, VolumeZ(Old), Visc_H, Diff_V, DWZ, DZZ,
//      AreaU/V, the six int32 masks), once per strip and level, shared by all properties of the batch.
//  So the coefficient pass (Convert_Dif_*, Compute_DifH/V_Constants AD:2364-2675, 1514-1619; the per-face part of
//  ComputeAdvection1D_V2 / ComputeAdvectionFace MF:10534-10894) no longer exists as a kernel: no pack or coefficient
//  array is ever written to HBM, the step reads the 112 B per cell of raw inputs once (plus neighbour-column re-reads
//  that hit L2) and 16 B per cell and property.  Same expressions, in the same order, as adt_lean_coef_kernel and
//  adt_transport_lean_kernel: results equal that path to the last bit wherever the compiler contracts the same way
//  (tools/fused_check.py reports the difference), and the CPU restatement within the tolerances of tests/test_gpu_parity.py.
//
//  Global loads: every per-level load of every warp is a cp.async (LDGSTS) into a private staging area, FR_NS - 1
//  levels ahead, awaited with cp.async.wait_group and read back with LDS.  Measured reason (tools/sassctl.py on the
//  round-1 / lean kernels): ptxas puts the look-ahead LDGs of level k+1 on the same scoreboard as the shuffles and the
//  loads of level k, so the first use of ANY of them waits for the newest load and the register look-ahead hides
//  nothing; cp.async groups are counted, not scoreboarded.
//
//  Hand-over: ring of FR_D slots, one level each, with two mbarriers per slot: "full" (one arrival per coefficient
//  warp, the property warps wait on its phase) and "empty" (one arrival per property warp, the coefficient warps wait).
//  Slot layout: 5 packs x 2 halves x 32 lanes x 16 B, so that every access is a conflict-free 128-bit one, then four
//  mask words per lane: each coefficient warp contributes the bits of the cell's 32-bit mask it knows.  Divisions use
//  the reciprocal seed + Newton + one residual correction (<= 1 ulp) instead of the IEEE sequence.
// =====================================================================================
#pragma once
#include "adt_lean_kernel.cuh"

namespace adt {

constexpr int FR_D = 4;                        // ring depth in levels
constexpr int FR_NS = 3;                       // stages of the cp.async look-ahead (levels k .. k+FR_NS-1 resident or in flight)
constexpr int FR_MASK = 5 * 128;               // offset of the mask words (4 per lane, one per coefficient warp) in a slot
constexpr int FR_SLOT = 5 * 128 + 64;          // doubles per slot
constexpr int FR_NCW = 4;                      // coefficient warps per block
enum { FP_U0 = 0, FP_U1 = 1, FP_V = 2, FP_W = 3, FP_C = 4 };

struct FusedArgs {
    LeanCoefArgs co;          // raw inputs, 2-D metrics, Schmidt numbers (co.pk unused)
    LeanArgs st;              // extents, property pointers (st.pk unused)
    int tiles_per_group;      // block order: tiles of a group fastest, then columns, then groups
    int ngroups;
    signed char role[16];     // per warp: property index (>= 0) or -1 - c for coefficient warp c; the host spreads the
                              // lighter coefficient warps over the SM sub-partitions (warp w runs on sub-partition w % 4)
};

__device__ __forceinline__ unsigned fsmem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fsmem_u32(bar)), "r"(count));
}
// every lane arrives for itself (the barriers count threads): each thread's own shared-memory accesses are then ordered
// before its own release, with no reliance on warp-level ordering, and compute-sanitizer's racecheck can follow it
#ifndef FUSED_ELECTED_ARRIVE
constexpr unsigned FBAR_PER_WARP = 32;
__device__ __forceinline__ void fbar_arrive_warp(uint64_t *bar, int) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fsmem_u32(bar)) : "memory");
}
#else
constexpr unsigned FBAR_PER_WARP = 1;
__device__ __forceinline__ void fbar_arrive_warp(uint64_t *bar, int lane) {
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fsmem_u32(bar)) : "memory");
}
#endif
__device__ __forceinline__ void fbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "FBAR_WAIT:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra FBAR_DONE;\n"
        " bra FBAR_WAIT;\n"
        "FBAR_DONE:\n"
        "}\n" ::"r"(fsmem_u32(bar)), "r"(parity)
        : "memory");
}

__device__ __forceinline__ void ring_store(double *slot, int pack, int lane, double a, double b, double c, double d) {
    double2 *p = reinterpret_cast<double2 *>(slot + pack * 128) + lane;
    p[0] = make_double2(a, b);
    p[32] = make_double2(c, d);
}
__device__ __forceinline__ Pack4 ring_load(const double *slot, int pack, int lane) {
    const double2 *p = reinterpret_cast<const double2 *>(slot + pack * 128) + lane;
    const double2 lo = p[0], hi = p[32];
    Pack4 r;
    r.a = lo.x; r.b = lo.y; r.c = hi.x; r.d = hi.y;
    return r;
}

// x / y and 1 / y for normal, non-zero y: MUFU.RCP64H seed, two Newton steps, one residual correction of the quotient
// (the IEEE fast path without its special-case branches; differs from the correctly rounded quotient by <= 1 ulp)
__device__ __forceinline__ double fdiv(double x, double y) {
    const double r = fast_rcp(y);
    const double q = x * r;
    return fma(fma(-y, q, x), r, q);
}
__device__ __forceinline__ double dt_over_if(bool ok, double dt, double v) { return (ok && v != 0.) ? fdiv(dt, v) : 0.; }
__device__ __forceinline__ double rcp_or_zero(double sd) { return sd != 0. ? fdiv(1.0, sd) : 0.; }
__device__ __forceinline__ unsigned bit_if(bool c, unsigned b) { return c ? b : 0u; }

__device__ __forceinline__ void cpa8(double *dst, const double *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(fsmem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa4(double *dst, const int *src) {       // into the low word of an 8-byte staging slot
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(fsmem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ int stg_int(const double *p) { return *reinterpret_cast<const int *>(p); }

// staged values per level and lane (8-byte slots; value v of stage t at stg[(t * NV + v) * 32 + lane])
constexpr int FU_NV = 11, FV_NV = 13, FW_NV = 9, FC_NV = 6;
__host__ __device__ constexpr int fused_stage_doubles(int nc) { return FR_NS * 32 * (2 * FU_NV + FV_NV + FW_NV + nc * FC_NV); }

// A coefficient warp: levels 1 .. K+1, loads FR_NS - 1 levels ahead, `empty` wait / `full` arrival around the stores
#define FUSED_COEF_MARCH(NV)                                                                                 \
    for (int v = 0; v < FR_NS * NV; ++v) stg[v * 32] = 0.;          /* predicated-off values read as zero */  \
    for (int l = 1; l < FR_NS; ++l) { issue(l, l % FR_NS); cpa_commit(); }                                   \
    int cs = 1 % FR_NS;                                                                                      \
    for (int k = 1; k <= K + 1; ++k) {                                                                       \
        const int is = cs == 0 ? FR_NS - 1 : cs - 1;                 /* slot of level k + FR_NS - 1 */       \
        if (k + FR_NS - 1 <= K + 1) issue(k + FR_NS - 1, is);                                                \
        cpa_commit();                                                                                        \
        cpa_wait<FR_NS - 1>();                                                                               \
        step(k, stg + cs * (NV * 32));                                                                       \
        cs = cs + 1 == FR_NS ? 0 : cs + 1;                                                                   \
    }
#define FUSED_SLOT_BEGIN(k)                                                                                  \
    double *slot = ring + ((k - 1) % FR_D) * FR_SLOT;                                                        \
    if (k > FR_D) fbar_wait(bars + FR_D + (k - 1) % FR_D, (unsigned)(((k - 1) / FR_D - 1) & 1));
#define FUSED_SLOT_END(k, w)                                                                                 \
    reinterpret_cast<unsigned *>(slot + FR_MASK)[lane * 4 + (w)] = m;                                        \
    fbar_arrive_warp(bars + (k - 1) % FR_D, lane);

// -------------------------------------------------------------------------------------
// Coefficient warps 0 and 1: the pack of the west U face of column jj = j + E (AD:4368-4582, 5156-5250, 1549-1553)
// and the mask bits of that side of the cell (i, j).
// -------------------------------------------------------------------------------------
template <bool TVD, int E>
__device__ __forceinline__ void fused_coef_u(const FusedArgs &s, double *__restrict__ ring, double *__restrict__ stg,
                                             const int ic, const int j, const int lane, uint64_t *__restrict__ bars) {
    const CoefArgs &a = s.co.c;
    const int sj = a.sj, sk = a.sk, ld2 = a.ld, K = a.K;
    const int jj = j + E;                                        // 1 <= jj <= J+1: columns jj-1 and jj exist
    const int q2 = ic + ld2 * jj;
    const double dux = a.DUX[q2], dux_w = a.DUX[q2 - ld2], dzx_w = a.DZX[q2 - ld2];
    const double rp = s.co.rhoUp[q2], rn = s.co.rhoUn[q2];
    const bool bnd = a.Bnd[ic + ld2 * j] == 1;
    const bool nb = a.Bnd[ic + ld2 * (E ? j + 1 : j - 1)] != 1;  // the neighbour across the face is not a boundary point
    const bool iin = ic >= 1 && ic <= a.I;
    const bool jm2 = jj >= 2, jp1 = jj + 1 < a.nj, jin = jj <= a.J;
    const bool up2 = s.co.upwind2_h != 0;
    const double dt = a.dt;
    const int q0 = ic + sj * jj;                                 // (i, jj, 0)

    // staged: 0 qx, 1 v_w, 2 v_c, 3 vi_w, 4 vi_c, 5 au, 6 o_m2, 7 o_m1, 8 o_c, 9 o_p1, 10 cfu
    auto issue = [&](const int lvl, const int slot) {
        const int q = q0 + sk * lvl;
        double *d = stg + slot * (FU_NV * 32);
        cpa8(d, a.Wflux_X + q); cpa8(d + 32, a.VolumeZ + (q - sj)); cpa8(d + 64, a.VolumeZ + q);
        cpa8(d + 96, a.Visc_H + (q - sj)); cpa8(d + 128, a.Visc_H + q); cpa8(d + 160, a.AreaU + q);
        if (jm2) cpa4(d + 192, a.Open + (q - 2 * sj));
        cpa4(d + 224, a.Open + (q - sj)); cpa4(d + 256, a.Open + q);
        if (jp1) cpa4(d + 288, a.Open + (q + sj));
        cpa4(d + 320, a.CFU + q);
    };
    auto step = [&](const int k, const double *r) {
        const double qx = r[0], v_w = r[32], v_c = r[64], vi_w = r[96], vi_c = r[128], au = r[160];
        const bool kin = k <= K;                                    // k >= 1 always
        const bool o_m2 = stg_int(r + 192) == 1, o_m1 = stg_int(r + 224) == 1, o_c = stg_int(r + 256) == 1,
                   o_p1 = stg_int(r + 288) == 1, cfu = stg_int(r + 320) == 1;
        const bool pos = qx > 0.;
        const double Qa = (cfu && o_c && o_m1) ? qx : 0.;
        double hc = 0.;
        if (TVD) {
            // upwind cell of the face (MF:10724-10736): jj-1 | jj; DT/V is zero outside the work area
            const double t_u = dt_over_if(iin && kin && (pos ? jm2 : jin), dt, pos ? v_w : v_c);
            const bool near = pos ? !o_m2 : !o_p1;
            hc = (near && up2) ? 0. : fma(-0.5 * qx, t_u, 0.5);
        }
        double hu;
        {
            double difx = fdiv(a.schmidt_h * (vi_c * dux_w + vi_w * dux), dux + dux_w);
            if (a.nulldif && qx == 0.) difx = 0.;
            hu = cfu ? fdiv(difx * au, dzx_w) : 0.;
        }
        // mask bits of cell (i, j): E = 0 the west side (this face, columns j-2, j-1), E = 1 the east side
        unsigned m;
        if (E == 0) {
            m = bit_if(cfu, M_CFU) | bit_if(o_m2, M_O_JM2) | bit_if(o_m1, M_O_JM1) | bit_if(pos, M_POS_W) |
                bit_if(bnd && o_m1 && nb, M_A_JM1);
        } else {
            m = bit_if(cfu, M_CFUE) | bit_if(o_c, M_O_JP1) | bit_if(o_p1, M_O_JP2) | bit_if(pos, M_POS_E) |
                bit_if(bnd && o_c && nb, M_A_JP1);
        }
        FUSED_SLOT_BEGIN(k)
        ring_store(slot, E ? FP_U1 : FP_U0, lane, 0.5 * Qa, fma(0.5, fabs(Qa), hu), fabs(Qa) * hc, pos ? rp : rn);
        FUSED_SLOT_END(k, E)
    };
    FUSED_COEF_MARCH(FU_NV)
}

// -------------------------------------------------------------------------------------
// Coefficient warp 2: the south V face (AD:4739-4953, 5254-5365, 1555-1559) and the mask bits along i.
// -------------------------------------------------------------------------------------
template <bool TVD>
__device__ __forceinline__ void fused_coef_v(const FusedArgs &s, double *__restrict__ ring, double *__restrict__ stg,
                                             const int ic, const int j, const int lane, uint64_t *__restrict__ bars) {
    const CoefArgs &a = s.co.c;
    const int sj = a.sj, sk = a.sk, ld2 = a.ld, K = a.K;
    const int q2 = ic + ld2 * j;
    const bool im2 = ic >= 2, ip1 = ic + 1 < a.ni, ip2 = ic + 2 < a.ni;
    const double dvy = a.DVY[q2], dvy_s = a.DVY[q2 - 1], dzy_s = a.DZY[q2 - 1];
    const double rVp = s.co.rhoVp[q2], rVn = s.co.rhoVn[q2];
    const bool bnd = a.Bnd[q2] == 1;
    const bool nb_ip = a.Bnd[q2 + (ip1 ? 1 : 0)] != 1, nb_im = a.Bnd[q2 - 1] != 1;
    const bool iin = ic >= 1 && ic <= a.I;
    const bool isin = ic >= 2 && ic <= a.I + 1;                 // the south neighbour lies in the work area
    const bool up2h = s.co.upwind2_h != 0;
    const double dt = a.dt;
    const int q0 = ic + sj * j;

    // staged: 0 qy, 1 v, 2 v_s, 3 vi, 4 vi_s, 5 av, 6 cfv, 7 cfv_n, 8 o_c, 9 o_im2, 10 o_im1, 11 o_ip1, 12 o_ip2
    auto issue = [&](const int lvl, const int slot) {
        const int q = q0 + sk * lvl;
        double *d = stg + slot * (FV_NV * 32);
        cpa8(d, a.Wflux_Y + q); cpa8(d + 32, a.VolumeZ + q); cpa8(d + 64, a.VolumeZ + (q - 1));
        cpa8(d + 96, a.Visc_H + q); cpa8(d + 128, a.Visc_H + (q - 1)); cpa8(d + 160, a.AreaV + q);
        cpa4(d + 192, a.CFV + q);
        if (ip1) cpa4(d + 224, a.CFV + (q + 1));
        cpa4(d + 256, a.Open + q);
        if (im2) cpa4(d + 288, a.Open + (q - 2));
        cpa4(d + 320, a.Open + (q - 1));
        if (ip1) cpa4(d + 352, a.Open + (q + 1));
        if (ip2) cpa4(d + 384, a.Open + (q + 2));
    };
    auto step = [&](const int k, const double *r) {
        const double qy = r[0], v = r[32], v_s = r[64], vi = r[96], vi_s = r[128], av = r[160];
        const bool kin = k <= K;
        const bool cfv = stg_int(r + 192) == 1, cfv_n = stg_int(r + 224) == 1, o_c = stg_int(r + 256) == 1;
        const bool oim2b = stg_int(r + 288) == 1, oim1b = stg_int(r + 320) == 1, oip1b = stg_int(r + 352) == 1,
                   oip2b = stg_int(r + 384) == 1;
        const bool pos_v = qy > 0.;
        const unsigned m = bit_if(cfv, M_CFV) | bit_if(cfv_n, M_CFVN) | bit_if(oim2b, M_O_IM2) | bit_if(oim1b, M_O_IM1) |
                           bit_if(oip1b, M_O_IP1) | bit_if(oip2b, M_O_IP2) | bit_if(pos_v, M_POS_S) |
                           bit_if(bnd && oip1b && nb_ip, M_A_IP1) | bit_if(bnd && oim1b && nb_im, M_A_IM1);
        const double Qa_v = (cfv && o_c && oim1b) ? qy : 0.;
        double hc_v = 0.;
        if (TVD) {
            const double t_v = dt_over_if(kin && (pos_v ? isin : iin), dt, pos_v ? v_s : v);
            const bool near_v = pos_v ? !oim2b : !oip1b;
            hc_v = (near_v && up2h) ? 0. : fma(-0.5 * qy, t_v, 0.5);
        }
        double hv;
        {
            double dify = fdiv(a.schmidt_h * (vi * dvy_s + vi_s * dvy), dvy + dvy_s);
            if (a.nulldif && qy == 0.) dify = 0.;
            hv = cfv ? fdiv(dify * av, dzy_s) : 0.;
        }
        FUSED_SLOT_BEGIN(k)
        ring_store(slot, FP_V, lane, 0.5 * Qa_v, fma(0.5, fabs(Qa_v), hv), fabs(Qa_v) * hc_v, pos_v ? rVp : rVn);
        FUSED_SLOT_END(k, 2)
    };
    FUSED_COEF_MARCH(FV_NV)
}

// -------------------------------------------------------------------------------------
// Coefficient warp 3: W (bottom face, AD:2941-3144) and C (DT/V, Vold/V, Diff_V_Const AD:1591-1597) and the cell's own
// and vertical mask bits.  Vertical neighbours roll through registers.
// -------------------------------------------------------------------------------------
template <bool TVD>
__device__ __forceinline__ void fused_coef_wc(const FusedArgs &s, double *__restrict__ ring, double *__restrict__ stg,
                                              const int ic, const int j, const int lane, uint64_t *__restrict__ bars) {
    const CoefArgs &a = s.co.c;
    const int sj = a.sj, sk = a.sk, ld2 = a.ld, K = a.K;
    const int q2 = ic + ld2 * j;
    const double dux = a.DUX[q2], dvy = a.DVY[q2];
    const bool bnd = a.Bnd[q2] == 1;
    const int small = a.SmallDepths ? a.SmallDepths[q2] : 0;
    const int cq = ic + sj * j;                                 // (i, j, 0)
    const bool wat_top = a.Water[cq + sk * K] == 1, open_top = a.Open[cq + sk * K] == 1;
    const unsigned mcol = bit_if(bnd, M_BND) | bit_if(wat_top, M_COLWET) | bit_if(open_top, M_COLOPEN);
    const bool iin = ic >= 1 && ic <= a.I;
    const bool up2v = s.co.upwind2_v != 0;
    const double dt = a.dt;

    // staged: 0 qz, 1 v, 2 vold, 3 dfv, 4 dzz(k-1), 5 dwz(k+1), 6 land, 7 open(k+2), 8 cfw(k+1); the planes above K+1 do
    // not exist: the values are masked when they are read
    auto issue = [&](const int lvl, const int slot) {
        const int q = cq + sk * lvl;
        double *d = stg + slot * (FW_NV * 32);
        cpa8(d, a.Wflux_Z + q); cpa8(d + 32, a.VolumeZ + q); cpa8(d + 64, a.VolumeZOld + q);
        cpa8(d + 96, a.Diff_V + q); cpa8(d + 128, a.DZZ + (q - sk));
        cpa4(d + 192, a.Land + q);
        if (lvl + 1 <= K + 1) { cpa8(d + 160, a.DWZ + (q + sk)); cpa4(d + 256, a.CFW + (q + sk)); }
        if (lvl + 2 <= K + 1) cpa4(d + 224, a.Open + (q + 2 * sk));
    };
    // rolling vertical state at the entry of level k:
    //   ob_m2 .. ob_p1 = Open == 1 of planes k-2 .. k+1 (false outside 0 .. K+1); the level's own loads bring plane k+2
    //   cfw_c = CFW(k) == 1; the loads bring CFW(k+1)
    //   rdc_m, rdc_c = 1/(DWZ+DWZ) of the faces below k-1 and below k; the loads bring DWZ(k+1) for the face above
    bool ob_m2 = false, ob_m1 = a.Open[cq] == 1, ob_c = a.Open[cq + sk] == 1, ob_p1 = a.Open[cq + 2 * sk] == 1;   // K >= 2
    bool cfw_c = a.CFW[cq + sk] == 1;
    double dwz_c = a.DWZ[cq + sk];
    double rdc_m = 0., rdc_c = rcp_or_zero(dwz_c + a.DWZ[cq]);
    double dtv_m = 0.;

    auto step = [&](const int k, const double *r) {
        const double qz = r[0], v = r[32], vold = r[64], dfv = r[96], dzz_m = r[128];
        const bool kin = k <= K, has_p1 = k + 1 <= K + 1;
        const double dwz_kp1 = has_p1 ? r[160] : 0.;
        const bool ob_p2 = (k + 2 <= K + 1) && stg_int(r + 224) == 1;     // plane k+2
        const bool cfw_t = has_p1 && stg_int(r + 256) == 1;               // plane k+1
        const bool land = stg_int(r + 192) == 1;
        const bool cfw = cfw_c;
        const bool pos_b = qz > 0.;
        const bool inwork = iin && kin;
        const double dtv = dt_over_if(inwork, dt, v);
        const double vr = (ob_c && inwork && v != 0.) ? fdiv(vold, v) : 1.;
        const unsigned m = mcol | bit_if(ob_c, M_OPEN) | bit_if(cfw, M_CFW) | bit_if(cfw_t, M_CFWT) | bit_if(land, M_LAND) |
                           bit_if(ob_m1, M_O_KM1) | bit_if(ob_p1, M_O_KP1) | bit_if(ob_p2, M_O_KP2) | bit_if(pos_b, M_POS_B);
        double hv_b = 0.;
        if (TVD) {
            const double t_b = pos_b ? dtv_m : dtv;                 // DT/V of the upwind cell k-1 | k
            const bool near_b = pos_b ? !ob_m2 : !ob_p1;
            hv_b = (near_b && up2v) ? 0. : 0.5 * (1. - qz * t_b);
        }
        double vz;
        {
            double difz = (a.schmidt_coef_v * dfv + a.schmidt_bg_v);
            if (a.nulldif_v && qz == 0.) difz = 0.;
            const double auxk = difz * dux * dvy;
            vz = (cfw && small == 0) ? fdiv(auxk, dzz_m) : 0.;
        }
        const double qa_b = (cfw && ob_c && ob_m1 && open_top) ? qz : 0.;
        const double rdc_p = has_p1 ? rcp_or_zero(dwz_kp1 + dwz_c) : 0.;      // face above the cell
        const double rdu_b = pos_b ? rdc_m : (has_p1 ? rdc_p : rdc_c);
        FUSED_SLOT_BEGIN(k)
        ring_store(slot, FP_W, lane, qa_b, hv_b, rdu_b, rdc_c);
        ring_store(slot, FP_C, lane, dtv, vr, vz, 0.);
        FUSED_SLOT_END(k, 3)
        // ---- roll ----
        ob_m2 = ob_m1; ob_m1 = ob_c; ob_c = ob_p1; ob_p1 = ob_p2;
        cfw_c = cfw_t;
        rdc_m = rdc_c; rdc_c = rdc_p; dwz_c = dwz_kp1;
        dtv_m = dtv;
    };
    FUSED_COEF_MARCH(FW_NV)
}
#undef FUSED_SLOT_BEGIN
#undef FUSED_SLOT_END
#undef FUSED_COEF_MARCH

// -------------------------------------------------------------------------------------
// The kernel.  blockDim.x = 32 * (NC + FR_NCW); dynamic shared memory (doubles): ring FR_D * FR_SLOT, 2 FR_D
// mbarriers, staging fused_stage_doubles(NC), W of the column solve K * NC * 32.
// -------------------------------------------------------------------------------------
__host__ __device__ constexpr size_t fused_smem_doubles(int K, int nc) {
    return (size_t)FR_D * FR_SLOT + 2 * FR_D + fused_stage_doubles(nc) + (size_t)K * nc * 32;
}

template <int M, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) adt_transport_fused_kernel(const __grid_constant__ FusedArgs fa) {
    extern __shared__ __align__(128) double fused_smem_base[];     // 128-bit shared accesses: the base must be 16-byte aligned
    double *const smem = fused_smem_base;
    const LeanArgs &s = fa.st;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nthreads = blockDim.x, NC = (nthreads >> 5) - FR_NCW;
    double *__restrict__ ring = smem;
    uint64_t *__restrict__ bars = reinterpret_cast<uint64_t *>(smem + FR_D * FR_SLOT);   // full[FR_D], empty[FR_D]
    if (threadIdx.x == 0) {
        for (int t = 0; t < FR_D; ++t) { fbar_init(bars + t, FR_NCW * FBAR_PER_WARP); fbar_init(bars + FR_D + t, (unsigned)NC * FBAR_PER_WARP); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // block -> (tile, column): tiles of a group fastest, then the columns, then the groups
    const int tpg = fa.tiles_per_group;
    const int per_group = tpg * s.j_count;
    const int grp = blockIdx.x / per_group, rem = blockIdx.x - grp * per_group;
    const int tile = grp * tpg + rem % tpg;
    const int j = s.j_begin + rem / tpg;
    if (tile >= s.ntile_i) return;                       // whole block: the last group may be short
    const int i = 1 + tile * 31 + lane;
    const int ic = min(i, s.I + 1);
    // a strip without a single wet column (land, dry tidal flat) has nothing to solve: adt_carry_kernel has already moved its
    // cells to the new position.  Every warp of the block sees the same 31 columns, so they all leave, before any hand-over
    if (!__any_sync(0xffffffffu, lane < 31 && i <= s.I && fa.co.c.Water[ic + s.sj * j + s.sk * s.K] == 1)) return;

    // staging areas: property warps first (FC_NV values each), then the four coefficient warps
    double *__restrict__ stage0 = smem + FR_D * FR_SLOT + 2 * FR_D;
    const int role = fa.role[warp];
    if (role < 0) {
        constexpr bool TVD = M == MOHID_P2_TVD;
        double *st = stage0 + FR_NS * 32 * (NC * FC_NV) + lane;
        if (role == -1) fused_coef_u<TVD, 0>(fa, ring, st, ic, j, lane, bars);
        else if (role == -2) fused_coef_u<TVD, 1>(fa, ring, st + FR_NS * 32 * FU_NV, ic, j, lane, bars);
        else if (role == -3) fused_coef_v<TVD>(fa, ring, st + FR_NS * 32 * (2 * FU_NV), ic, j, lane, bars);
        else fused_coef_wc<TVD>(fa, ring, st + FR_NS * 32 * (2 * FU_NV + FV_NV), ic, j, lane, bars);
        return;
    }

    const int n = role;
    const bool writer = (lane < 31) && (i <= s.I);
    const PropArgs &pa = s.p[n];
    const double *__restrict__ P = pa.pin;
    const int sj = s.sj, sk = s.sk;
    const long cp = ic + (long)sj * j;                    // column base (k = 0) in the property arrays
    const int je2 = (j + 2 <= s.J + 1) ? 2 * sj : sj;
    const int jw2 = (j >= 2) ? 2 * sj : sj;
    const bool colwet = fa.co.c.Water[cp + (long)sk * s.K] == 1;
    const bool obc = fa.co.c.Bnd[ic + fa.co.c.ld * j] == 1 && pa.bc != MOHID_BC_None;
    const double theta = pa.theta_difv, omt = 1. - pa.theta_difv;
    const double qz_top = s.qz[cp + (long)sk * (s.K + 1)];
    const bool halo_lane = (lane < 2) || (lane == 31);
    const int halo_off = (lane == 31) ? ((ic <= s.I) ? 1 : 0) : -2;

    const int qp = (int)cp + sk;                                                 // cell (i, j, 1)
    const int wstride = NC * 32;
    double *__restrict__ wsm0 = stage0 + fused_stage_doubles(NC) + n * 32 + lane;         // [K][NC][32]
    double *__restrict__ wsm = wsm0;
    double *__restrict__ pst = stage0 + FR_NS * 32 * (n * FC_NV) + lane;

    // staged per level: 0 P(j-2), 1 P(j-1), 2 P(j+1), 3 P(j+2), 4 the strip-halo value (lanes 0, 1, 31), 5 P of the own
    // column two levels up (plane K+2 does not exist: clamped like round 1)
    auto issue = [&](const int lvl, const int slot) {
        const int q = (int)cp + sk * lvl;
        double *d = pst + slot * (FC_NV * 32);
        cpa8(d, P + (q - jw2)); cpa8(d + 32, P + (q - sj)); cpa8(d + 64, P + (q + sj)); cpa8(d + 96, P + (q + je2));
        if (halo_lane) cpa8(d + 128, P + (q + halo_off));
        cpa8(d + 160, P + (q + sk * (lvl + 2 <= s.K + 1 ? 2 : 1)));
    };
    for (int v = 0; v < FR_NS * FC_NV; ++v) pst[v * 32] = 0.;      // the halo value reads as zero in the other lanes
    for (int l = 1; l < FR_NS; ++l) { if (l <= s.K) issue(l, l % FR_NS); cpa_commit(); }
    int cs = 1 % FR_NS;

    // ---- rolling state ----
    double Pm2 = 0., Pm1 = P[cp], Pc = P[qp], Pp1 = P[qp + sk];
    double dtv_m = 0.;
    double RD = 0., RE = 1., RTI = 0.;                    // row k-1 as far as it is known
    bool land_m = false, obc_m = false;                   // ... its land flag, and whether it is an open-boundary row
    unsigned m_m = 0;
    double Wp = 0., Gp = 0.;                              // W, G of the last eliminated row
    unsigned zp = 0;

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
            const double num = pos ? Pm1 - Pm2 : Pc - Pp1;
            const double dPd = pos ? dP : Pm1 - Pc;
            double dC = dPd * W.d;
            dC = (abs_bits(dC) < MIN_VALUE) ? with_sign_of(MIN_VALUE, dC) : dC;     // MF:10795-10803
            const double r = num * W.c * fast_rcp(dC);
            double ps = sel_lt(r + r, 1.);                             // SuperBee (MF:10826-10829) as exact selections
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
    auto level = [&](auto pos_tag, const int k) {
        constexpr int POS = decltype(pos_tag)::value;
        // ---- look-ahead: the loads of level k + FR_NS - 1 go out, those of level k have landed ----
        {
            const int is = cs == 0 ? FR_NS - 1 : cs - 1;
            if (k + FR_NS - 1 <= s.K) issue(k + FR_NS - 1, is);
            cpa_commit();
            cpa_wait<FR_NS - 1>();
        }
        const double *sv = pst + cs * (FC_NV * 32);
        cs = cs + 1 == FR_NS ? 0 : cs + 1;
        const double cPw2 = sv[0], cPw1 = sv[32], cPe1 = sv[64], cPe2 = sv[96], chP = sv[128], Pp2 = sv[160];
        // ---- the level's packs ----
        const int sl = (k - 1) % FR_D;
        const double *slot = ring + sl * FR_SLOT;
        fbar_wait(bars + sl, (unsigned)(((k - 1) / FR_D) & 1));
        const Pack4 C = ring_load(slot, FP_C, lane);
        const Pack4 U0 = ring_load(slot, FP_U0, lane), U1 = ring_load(slot, FP_U1, lane);
        const Pack4 V = ring_load(slot, FP_V, lane);
        const uint4 mw = reinterpret_cast<const uint4 *>(slot + FR_MASK)[lane];
        const unsigned m = mw.x | mw.y | mw.z | mw.w;
        const double dtv_c = C.a;

        // ---- horizontal faces (explicit) ----
        const double d1 = cPw1 - cPw2, d2 = Pc - cPw1, d3 = cPe1 - Pc, d4 = cPe2 - cPe1;
        const double fw = lean_face<M>((m & M_POS_W) != 0, d1, d2, d3, cPw1, Pc, U0);
        const double fe = lean_face<M>((m & M_POS_E) != 0, d2, d3, d4, Pc, cPe1, U1);
        // south face of every lane; the north face is the south face of lane + 1
        double Ps1 = shfl_up_d(Pc, 1);
        const double hP1 = __shfl_sync(0xffffffffu, chP, 1);
        Ps1 = lane == 0 ? hP1 : Ps1;
        const double gs = Pc - Ps1;
        double dsm = shfl_up_d(gs, 1), dsp = shfl_dn_d(gs, 1);
        dsm = lane == 0 ? Ps1 - chP : dsm;
        dsp = lane == 31 ? chP - Pc : dsp;
        const double fs = lean_face<M>((m & M_POS_S) != 0, dsm, gs, dsp, Ps1, Pc, V);
        const double fsum = (fw - fe) + (fs - shfl_dn_d(fs, 1));

        // ---- the face below: row k-1 is complete, eliminate it ----
        double Dn = 0., En = 0., TIn = 0.;
        if constexpr (POS != 0) {
            const Pack4 W = ring_load(slot, FP_W, lane);
            vface(k, C, W, m, Dn, En, TIn);
        }
        if (k + FR_D <= s.K + 1) fbar_arrive_warp(bars + FR_D + sl, lane);   // the slot may be refilled (level k + FR_D)

        // ---- row k: VolumeVariation (AD:3966-4021; Vold/V is 1 in closed cells) + the shares known so far ----
        double e0 = 1.0;
        if constexpr (POS == 2) e0 = (m & M_OPEN) ? 1.0 + dtv_c * qz_top : 1.0;
        RTI = fma(fsum, dtv_c, Pc * C.b + TIn);
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
        level(first, 1);
        int k = 2;
        for (; k + 1 < s.K; k += 2) { level(mid, k); level(mid, k + 1); }
        if (k < s.K) level(mid, k);
        level(last, s.K);
        // ---- virtual level K+1: the face above the surface cell completes row K ----
        const int sl = s.K % FR_D;
        const double *slot = ring + sl * FR_SLOT;
        fbar_wait(bars + sl, (unsigned)((s.K / FR_D) & 1));
        const Pack4 C = ring_load(slot, FP_C, lane), W = ring_load(slot, FP_W, lane);
        const uint4 mw = reinterpret_cast<const uint4 *>(slot + FR_MASK)[lane];
        double Dn, En, TIn;
        vface(s.K + 1, C, W, mw.x | mw.y | mw.z | mw.w, Dn, En, TIn);
    }

    // ---------------- back substitution (MF:4100-4105) ----------------
    if (writer && colwet) {
        double *__restrict__ O = pa.pout;
        long qo = cp + (long)sk * (s.K + 1);
        const double *__restrict__ Wsm = wsm0;
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
