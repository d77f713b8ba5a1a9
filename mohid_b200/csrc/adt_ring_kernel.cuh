// K2R: the fused transport step with the inputs staged through asynchronous-copy rings in shared memory (sm_100a).
//
// Same arithmetic as adt_transport_kernel<.., FULL = true, GGLOB = true> (adt_kernels.cuh): the level body below is
// that kernel's level body with the data sources changed, so the two produce bit-identical fields.  What changes is
// how the data reaches the threads:
//
//   block    <-> (31-cell strip along i, column j), persistent over the strip units of the slab
//   warp n   <-> property n of the batch (n < nprop): a consumer.  It streams its own five property rows (j-2..j+2,
//                40-element windows) of plane k+4 into a private 5-stage ring with 16-byte cp.async copies (four
//                per lane and level, no registers held) and tracks them with cp.async groups.
//   warp np  <-> producer of the shared per-step rows (13 coefficient rows + the mask row per plane): nine 16-byte
//                cp.async copies per lane and plane into a 5-stage ring, completion signalled on a "full" mbarrier
//                (cp.async.mbarrier.arrive); consumers release a stage on an "empty" mbarrier after the level.
//
// Why: in adt_transport_kernel every thread fetches its 21 values per level with its own LDG and holds the next level
// in registers.  ncu (profiles/r01_final_k1_k2.txt) shows a quarter of the kernel time waiting on those loads at the
// start of each level, the coefficient rows being fetched once per property (L1 hits at best), and 33 registers tied
// up as the look-ahead buffer.  Here the coefficient rows are fetched once per block, the look-ahead is 2-3 levels
// deep and lives in shared memory, the i-halo values come straight from the staged row (no shuffles / lane selects),
// and consumer loads are LDS with immediate offsets.  (A first version used one TMA bulk copy per row: 64 copies of
// 320 B per plane cost ~67 cycles each on the issuing warp and starved the consumers -- rows this narrow are
// too small for the TMA engine; 16-byte cp.async by all lanes moves the same rows in ~50 instructions per plane.)
//
// Restrictions (the caller falls back to adt_transport_kernel otherwise): FULL configuration (3-D, both horizontal
// directions, implicit vertical advection), no discharges / NoFlux lists, methods that do not look two cells upstream
// for DT/V (UpwindOrder1 and P2_TVD), nprop <= NCW, ld % 4 == 0 (16-byte source alignment of the int32 mask row).
#pragma once

namespace adt {

constexpr int RING_W = 40;     // elements per staged row window: cells i0-2 .. i0+31 after aligning the start down to 4
constexpr int RING_SS = 8;     // stages of the shared-row ring (planes k, k+1, k+2 in use, up to five in flight)
constexpr int RING_PS = 5;     // stages of each private property ring (three in use, two in flight)
enum RingRow : int {
    RR_TW, RR_TC, RR_TE,       // DT/V of columns j-1, j, j+1
    RR_QXW, RR_QXE,            // Wflux_X of faces j, j+1
    RR_DHW, RR_DHE,            // horizontal diffusion constant of U faces j, j+1
    RR_QY, RR_DHV,             // Wflux_Y, diffusion constant of the V face
    RR_VR, RR_RDZ, RR_QZ, RR_DVZ,
    RR_NSH
};
constexpr int RING_MASK_OFF = RR_NSH * RING_W * 8;
constexpr int RING_SH_BYTES = RING_MASK_OFF + RING_W * 4;      // 4160 + 160
constexpr int RING_P_BYTES = 5 * RING_W * 8;                   // rows j-2 .. j+2 of one property
constexpr int RING_SH_CHUNKS = RING_SH_BYTES / 16;            // 16-byte pieces of the shared rows of one plane (270)
constexpr int RING_P_CHUNKS = RING_P_BYTES / 16;              // ... of one property's five rows (100)
constexpr int RING_SH_PER_LANE = (RING_SH_CHUNKS + 31) / 32;  // 9
constexpr int RING_P_PER_LANE = (RING_P_CHUNKS + 31) / 32;    // 4
// shared memory: W of the column solve [K][nprop][32], shared-row ring, private property rings, 2 * SS mbarriers
__host__ __device__ constexpr size_t ring_smem_bytes(int nprop, int K) {
    return (size_t)nprop * K * 32 * 8 + (size_t)RING_SS * RING_SH_BYTES + (size_t)RING_PS * nprop * RING_P_BYTES +
           2 * RING_SS * 8;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    const unsigned a = smem_u32(bar);
    unsigned done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!done);
}
// 16-byte asynchronous global -> shared copy (L2 only: the rows are consumed from shared memory)
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// arrive on `bar` once every cp.async this thread has issued so far has landed (counted in the barrier's init count)
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int MH, int LH, int MV, int LV, int NCW, int NPT = 1>
__global__ void __launch_bounds__((NCW + 1) * 32, 1) adt_transport_ring_kernel(const __grid_constant__ StepArgs s) {
    extern __shared__ __align__(128) unsigned char ring_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int np = s.nprop;
    double *const Wbase = reinterpret_cast<double *>(ring_smem);                     // [K][np][32]
    unsigned char *const shring = ring_smem + (size_t)np * s.K * 32 * 8;             // [SS][RING_SH_BYTES]
    unsigned char *const pring = shring + (size_t)RING_SS * RING_SH_BYTES;           // [np][PS][RING_P_BYTES]
    uint64_t *const full = reinterpret_cast<uint64_t *>(pring + (size_t)np * RING_PS * RING_P_BYTES);
    uint64_t *const empty = full + RING_SS;
    if (threadIdx.x == 0) {
        for (int t = 0; t < RING_SS; ++t) { mbar_init(full + t, 32); mbar_init(empty + t, (unsigned)((np + NPT - 1) / NPT)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int sj = s.sj, sk = s.sk, sj2 = s.ld, K = s.K;
    const long nsu = (long)s.ntile_i * s.j_count;
    const int planes = K + 1;                               // planes 1 .. K+1 of a column pass through the rings

    if (warp == (np + NPT - 1) / NPT) {
        // =========================== producer of the shared rows ===========================
        // chunk id = lane + 32 t: row = id / 20, piece = id % 20 for the 13 double rows, ids 260..269 = the mask row
        unsigned g = 0;                                     // running plane counter: stage = g % SS, phase = (g / SS) & 1
        for (long su = blockIdx.x; su < nsu; su += gridDim.x) {
            const int tile = (int)(su % s.ntile_i);
            const int j = (int)(su / s.ntile_i) + s.j_begin;
            const int a0 = (1 + tile * 31 - 2) & ~3;        // window start (may be -4: the tail of the previous row)
            const long col = (long)a0 + (long)sj * j;
            const unsigned char *src0[RING_SH_PER_LANE];    // source of this lane's pieces at plane 0
            unsigned pstride[RING_SH_PER_LANE];             // bytes per plane (0 = no piece)
#pragma unroll
            for (int t = 0; t < RING_SH_PER_LANE; ++t) {
                const int id = lane + 32 * t;
                src0[t] = nullptr; pstride[t] = 0;
                if (id >= RING_SH_CHUNKS) continue;
                if (id < RR_NSH * (RING_W / 2)) {
                    const int row = id / (RING_W / 2), piece = id % (RING_W / 2);
                    const double *arr = s.dvz;
                    int ro = 0;
                    switch (row) {
                        case RR_TW: arr = s.dtv; ro = -sj; break;
                        case RR_TC: arr = s.dtv; break;
                        case RR_TE: arr = s.dtv; ro = sj; break;
                        case RR_QXW: arr = s.qx; break;
                        case RR_QXE: arr = s.qx; ro = sj; break;
                        case RR_DHW: arr = s.dhu; break;
                        case RR_DHE: arr = s.dhu; ro = sj; break;
                        case RR_QY: arr = s.qy; break;
                        case RR_DHV: arr = s.dhv; break;
                        case RR_VR: arr = s.vr; break;
                        case RR_RDZ: arr = s.rdz; break;
                        case RR_QZ: arr = s.qz; break;
                        default: break;
                    }
                    src0[t] = reinterpret_cast<const unsigned char *>(arr + col + ro + 2 * piece);
                    pstride[t] = (unsigned)sk * 8u;
                } else {
                    const int piece = id - RR_NSH * (RING_W / 2);
                    src0[t] = reinterpret_cast<const unsigned char *>(s.mask + col + 4 * piece);
                    pstride[t] = (unsigned)sk * 4u;
                }
            }
            for (int p = 1; p <= planes; ++p, ++g) {
                const unsigned st = g % RING_SS, ph = (g / RING_SS) & 1u;
#ifdef ADT_EXPERIMENT
                const long long t0 = clock64();
                mbar_wait(empty + st, ph ^ 1u);
                const long long t1 = clock64();
                if (lane == 0) { atomicAdd(s.zero_pivots + 1, (unsigned long long)(t1 - t0)); atomicAdd(s.zero_pivots + 3, 1ull); }
#else
                mbar_wait(empty + st, ph ^ 1u);             // every consumer has released the stage
#endif
                unsigned char *dst = shring + (size_t)st * RING_SH_BYTES + 16 * lane;
#pragma unroll
                for (int t = 0; t < RING_SH_PER_LANE; ++t)
                    if (pstride[t]) cp_async16(dst + 512 * t, src0[t] + (size_t)p * pstride[t]);
                cp_async_arrive(full + st);
            }
        }
        cp_async_wait<0>();
        return;
    }
    // =========================== consumers ===========================
    // consumer warp w advances properties NPT*w .. NPT*w + NPT-1 together: the mask logic, the flow-direction selects
    // and the shared-row loads of a level are common to them (the compiler merges the common subexpressions of the
    // unrolled property loop), and two independent dependency chains run through the level
    const int ncons = (np + NPT - 1) / NPT;
    if (warp >= ncons) return;
    int pn[NPT];
    bool live[NPT];
    const double *__restrict__ P[NPT];
    double *__restrict__ Wsm[NPT];
    double *myring[NPT];
    double theta[NPT], omt[NPT];
#pragma unroll
    for (int u = 0; u < NPT; ++u) {
        live[u] = NPT * warp + u < np;
        pn[u] = live[u] ? NPT * warp + u : NPT * warp;    // an odd last property is shadowed by its neighbour
        P[u] = s.p[pn[u]].pin;
        Wsm[u] = Wbase + pn[u] * 32 + lane;
        myring[u] = reinterpret_cast<double *>(pring + (size_t)pn[u] * RING_PS * RING_P_BYTES);
        theta[u] = s.p[pn[u]].theta_difv; omt[u] = 1. - theta[u];
    }
    const int wstride = np * 32;
    constexpr int PSTG = RING_P_BYTES / 8;                // doubles per private stage
    constexpr bool FAST_H = (MH == MOHID_P2_TVD && LH == MOHID_SuperBee);

    // ---- private prefetch stream: the five property rows of every plane, continuous over the units of the block ----
    // piece id = lane + 32 t (16 bytes each): row = id / 20 (j-2 .. j+2), piece = id % 20; ids 96..99 only on lanes 0..3
    long isu = blockIdx.x;                                // unit and plane of the next rows to request
    int ip = 1;
    unsigned ist = 0;                                     // private stage they go to
    int psrc[RING_P_PER_LANE] = {0, 0, 0, 0};             // element offset of this lane's pieces from P at plane 0
    auto set_psrc = [&](long su_) {
        const int tile_ = (int)(su_ % s.ntile_i);
        const int j_ = (int)(su_ / s.ntile_i) + s.j_begin;
        const int a0_ = (1 + tile_ * 31 - 2) & ~3;
        const int jw2_ = (j_ >= 2) ? 2 * sj : sj, je2_ = (j_ + 2 <= s.J + 1) ? 2 * sj : sj;
#pragma unroll
        for (int t = 0; t < RING_P_PER_LANE; ++t) {
            const int id = lane + 32 * t;
            const int row = id / (RING_W / 2), piece = id % (RING_W / 2);
            const int ro = row == 0 ? -jw2_ : row == 1 ? -sj : row == 2 ? 0 : row == 3 ? sj : je2_;
            psrc[t] = a0_ + sj * j_ + ro + 2 * piece;
        }
    };
    if (isu < nsu) set_psrc(isu);
    auto issue_next = [&]() {
        if (isu < nsu) {
#pragma unroll
            for (int u = 0; u < NPT; ++u) {
                if (u > 0 && !live[u]) continue;
                unsigned char *dst = reinterpret_cast<unsigned char *>(myring[u] + ist * PSTG) + 16 * lane;
                const double *src = P[u] + (size_t)sk * ip;
#pragma unroll
                for (int t = 0; t < RING_P_PER_LANE; ++t)
                    if (t < RING_P_PER_LANE - 1 || lane < RING_P_CHUNKS - 32 * (RING_P_PER_LANE - 1))
                        cp_async16(dst + 512 * t, src + psrc[t]);
            }
            if (++ip > planes) {
                ip = 1;
                isu += gridDim.x;
                if (isu < nsu) set_psrc(isu);
            }
        }
        cp_async_commit();                                // (possibly empty) group: one per plane keeps the count uniform
        ist = (ist + 1 == RING_PS) ? 0 : ist + 1;
    };
#pragma unroll
    for (int t = 0; t < RING_PS - 1; ++t) issue_next();
    unsigned pst = 0;                                     // private stage of the plane of the current level
    unsigned gsh = 0;                                     // shared-ring plane counter of that plane
    auto pnext = [](unsigned st) { return (st + 1 == RING_PS) ? 0u : st + 1; };
    auto sh_stage = [&](unsigned g) { return shring + (size_t)(g % RING_SS) * RING_SH_BYTES; };
    auto wait_full = [&](unsigned g) { mbar_wait(full + (g % RING_SS), (g / RING_SS) & 1u); };

    for (long su = blockIdx.x; su < nsu; su += gridDim.x) {
        const int tile = (int)(su % s.ntile_i);
        const int j = (int)(su / s.ntile_i) + s.j_begin;
        const int i0 = 1 + tile * 31;
        const int a0 = (i0 - 2) & ~3;
        const int i = i0 + lane;
        const bool writer = (lane < 31) && (i <= s.I);
        const int ic = min(i, s.I + 1);                   // clamped column for the few direct global loads
        const int wi = i - a0;                            // this lane's element in a staged row window
        const int c2 = ic + sj2 * j, c2d = ic + sj * j;
        const int je2_2 = (j + 2 <= s.J + 1) ? 2 * sj2 : sj2;

        // ---- 2-D metrics of the column (as in adt_transport_kernel) ----
        const double rdx_m = s.rdx[c2 - sj2], rdx_c = s.rdx[c2], rdx_p = s.rdx[c2 + sj2], rdx_pp = s.rdx[c2 + je2_2];
        const double rdy_c = s.rdy[c2];
        double rdy_m = shfl_up_d(rdy_c, 1), rdy_p = shfl_dn_d(rdy_c, 1);
        if (lane == 0) rdy_m = s.rdy[c2 - 1];
        if (lane == 31) rdy_p = s.rdy[c2 + (ic <= s.I ? 1 : 0)];
        const double rho_wp = FAST_H ? ratio_or_zero(rdx_m, rdx_c) : rdx_m, rho_wn = FAST_H ? ratio_or_zero(rdx_p, rdx_c) : rdx_p;
        const double rho_ep = FAST_H ? ratio_or_zero(rdx_c, rdx_p) : rdx_c, rho_en = FAST_H ? ratio_or_zero(rdx_pp, rdx_p) : rdx_pp;
        const double rho_sp = FAST_H ? ratio_or_zero(rdy_m, rdy_c) : rdy_m, rho_sn = FAST_H ? ratio_or_zero(rdy_p, rdy_c) : rdy_p;

        const unsigned mtop = s.mask[c2d + sk * K];
        const bool colwet = (mtop & M_COLWET) != 0, colopen = (mtop & M_COLOPEN) != 0;
        const bool bnd = (mtop & M_BND) != 0;
        const unsigned top_req = colopen ? (M_OPEN | M_O_KP1 | M_CFWT) : (1u << 31);

        auto shrow = [&](const unsigned char *st, int row) { return reinterpret_cast<const double *>(st) + row * RING_W + wi; };

        // ---- rolling registers along k ----
        cp_async_wait<1>();                                 // own rows of planes 1, 2 have landed (see the level loop)
        __syncwarp();
        wait_full(gsh);
        wait_full(gsh + 1);                                 // shared rows of planes 1 and 2 (K + 1 >= 2)
        int q = c2d + sk;                                   // cell (i,j,1)
        double Pm1[NPT], Pc[NPT], Pp1[NPT], Dk[NPT], Ek_b[NPT], TIk_b[NPT], Wprev[NPT], Gprev[NPT];
#pragma unroll
        for (int u = 0; u < NPT; ++u) {
            Pm1[u] = P[u][c2d];
            Pc[u] = myring[u][pst * PSTG + 2 * RING_W + wi];
            Pp1[u] = myring[u][pnext(pst) * PSTG + 2 * RING_W + wi];
            Dk[u] = 0.; Ek_b[u] = 0.; TIk_b[u] = 0.; Wprev[u] = 0.; Gprev[u] = 0.;
        }
        double dtv_m = 0., dtv_c = *shrow(sh_stage(gsh), RR_TC);
        double rdz_c = *shrow(sh_stage(gsh), RR_RDZ), rdz_p = *shrow(sh_stage(gsh + 1), RR_RDZ);
        double qz_c = *shrow(sh_stage(gsh), RR_QZ);
        unsigned zp = 0;

        for (int k = 1; k <= K; ++k, ++gsh) {
            // request the rows four planes ahead (into the stage that held plane k-1); afterwards at most the two
            // newest groups may still be in flight, so the rows of plane k+2 have landed
            issue_next();
            cp_async_wait<2>();
            __syncwarp();
            const bool has_c = k + 2 <= planes;           // plane k+2 exists (else it is clamped to plane K+1)
#ifdef ADT_EXPERIMENT
            { const long long t0 = clock64(); if (has_c) wait_full(gsh + 2); const long long t1 = clock64();
              if (lane == 0 && warp == 0) atomicAdd(s.zero_pivots + 2, (unsigned long long)(t1 - t0)); }
#else
            if (has_c) wait_full(gsh + 2);
#endif
            const unsigned pstB = pnext(pst), pstC = has_c ? pnext(pstB) : pstB;
            const unsigned char *A = sh_stage(gsh), *B = sh_stage(gsh + 1), *C = sh_stage(has_c ? gsh + 2 : gsh + 1);
            const unsigned m = *(reinterpret_cast<const uint32_t *>(A + RING_MASK_OFF) + wi);
            const double vr = *shrow(A, RR_VR);
            const double t_w = *shrow(A, RR_TW), t_e = *shrow(A, RR_TE), t_s = shrow(A, RR_TC)[-1];
            const double qxw = *shrow(A, RR_QXW), qxe = *shrow(A, RR_QXE), dhw = *shrow(A, RR_DHW), dhe = *shrow(A, RR_DHE);
            const double qys = *shrow(A, RR_QY), dhs = *shrow(A, RR_DHV);
            const double dtv_p = *shrow(B, RR_TC), qz_p = *shrow(B, RR_QZ), dvz_p = *shrow(B, RR_DVZ);
            const double rdz_pp = *shrow(C, RR_RDZ), dtv_pp = *shrow(C, RR_TC);
            const bool open_c = (m & M_OPEN) != 0;

#pragma unroll
            for (int u = 0; u < NPT; ++u) {
                const double *PA = myring[u] + pst * PSTG + wi;
                const double Pw2 = PA[0], Pw1 = PA[RING_W], Pe1 = PA[3 * RING_W], Pe2 = PA[4 * RING_W];
                const double Ps2 = PA[2 * RING_W - 2], Ps1 = PA[2 * RING_W - 1], Pn1 = PA[2 * RING_W + 1];
                const double Pp2 = myring[u][pstC * PSTG + 2 * RING_W + wi];
                const double Pcu = Pc[u];
                // ---------------- VolumeVariation (AD:3966-4021) ----------------
                Row row;
                row.TI = sel(open_c, Pcu * vr, Pcu) + TIk_b[u];
                row.E = sel(open_c && k == K, 1.0 + dtv_c * qz_p, 1.0) + Ek_b[u];
                row.D = Dk[u];
                row.F = 0.;
                // ---------------- horizontal faces (explicit) ----------------
                {
                    const bool o_w1 = (m & M_O_JM1) != 0, o_e1 = (m & M_O_JP1) != 0;
                    const double fw = hface_flux<MH, LH>(s, all_set(m, M_CFU | M_O_JM1 | M_OPEN), qxw, dhw, Pw2, Pw1, Pcu, Pe1,
                                                         (m & M_O_JM2) != 0, o_e1, 0., t_w, dtv_c, t_e, rho_wp, rdx_c, rho_wn, 0., 0.);
                    const double fe = hface_flux<MH, LH>(s, all_set(m, M_CFUE | M_O_JP1 | M_OPEN), qxe, dhe, Pw1, Pcu, Pe1, Pe2,
                                                         o_w1, (m & M_O_JP2) != 0, t_w, dtv_c, t_e, 0., rho_ep, rdx_p, rho_en, 0., 0.);
                    // each lane builds its south face; the north face is the south face of lane+1
                    const double fs = hface_flux<MH, LH>(s, all_set(m, M_CFV | M_O_IM1 | M_OPEN), qys, dhs, Ps2, Ps1, Pcu, Pn1,
                                                         (m & M_O_IM2) != 0, (m & M_O_IP1) != 0, 0., t_s, dtv_c, 0., rho_sp, rdy_c,
                                                         rho_sn, 0., 0.);
                    const double fsum = (fw - fe) + (fs - shfl_dn_d(fs, 1));
                    row.TI += fsum * dtv_c;
                }
                // ---------------- vertical face k+1 (top of this cell) ----------------
                double Dn, En_b, TIn_b;
                {
                    const double aux1 = dvz_p * dtv_c, aux2 = dvz_p * dtv_p;
                    const double dP = Pp1[u] - Pcu;
                    row.E += aux1 * theta[u];
                    row.F -= aux1 * theta[u];
                    row.TI += aux1 * dP * omt[u];
                    Dn = -aux2 * theta[u];
                    En_b = aux2 * theta[u];
                    TIn_b = -aux2 * dP * omt[u];
                    const bool adv_on = all_set(m, top_req);
                    const bool pos = qz_p > 0.;
                    const double Puu = sel(pos, Pm1[u], Pp2), Pu = sel(pos, Pcu, Pp1[u]), Pd = sel(pos, Pp1[u], Pcu);
                    double wuu, wu, wd;
                    oriented_weights<MV, LV>(s.method_v, s.limiter_v, s.upwind2_v != 0, s.vrelmax, qz_p, Puu, Pu, Pd,
                                             pos ? !(m & M_O_KM1) : !(m & M_O_KP2), sel(pos, dtv_m, dtv_pp), sel(pos, dtv_c, dtv_p),
                                             sel(pos, dtv_p, dtv_c), sel(pos, rdz_c, rdz_pp), rdz_p, 0., 0., wuu, wu, wd);
                    const double qa = sel(adv_on, qz_p, 0.);
                    const double dfl = qa * sel(pos, wu, wd), efl = qa * sel(pos, wd, wu);   // D_flux, E_flux (MF:10583-10586)
                    row.E += dfl * dtv_c;
                    row.F += efl * dtv_c;
                    Dn -= dfl * dtv_p;
                    En_b -= efl * dtv_p;
                }
                // ---------------- land fill (AD:1753) ----------------
                row.TI = sel((m & M_LAND) != 0, NULL_REAL, row.TI);
                // ---------------- Thomas forward elimination, row k (MF:4087-4099) ----------------
                const double Wp0 = Wprev[u], Gp0 = Gprev[u];
                {
                    const double aux = row.E + row.D * Wp0;
                    const bool ok = aux != 0.;
                    const double ra = fast_rcp(aux);
                    Wprev[u] = sel(ok, -row.F * ra, Wp0);
                    Gprev[u] = sel(ok, (row.TI - row.D * Gp0) * ra, Gp0);
                    zp += (ok || !live[u]) ? 0u : 1u;
                }
                // ---------------- open boundary rows (AD:5369-5672); rare ----------------
                if (bnd && open_c && s.p[pn[u]].bc != MOHID_BC_None) {
                    open_boundary_row(s, s.p[pn[u]], q, m, Pcu, qz_c, qz_p, dtv_c, row);
                    const double aux = row.E + row.D * Wp0;
                    if (aux != 0.) {
                        const double ra = 1.0 / aux;
                        Wprev[u] = -row.F * ra;
                        Gprev[u] = (row.TI - row.D * Gp0) * ra;
                    } else {
                        Wprev[u] = Wp0; Gprev[u] = Gp0;
                    }
                }
                if (live[u]) {
                    Wsm[u][(size_t)(k - 1) * wstride] = Wprev[u];
                    if (writer && colwet) s.p[pn[u]].pout[q] = Gprev[u];      // G parked in the output array
                }
                // ---------------- roll ----------------
                Dk[u] = Dn; Ek_b[u] = En_b; TIk_b[u] = TIn_b;
                Pm1[u] = Pcu; Pc[u] = Pp1[u]; Pp1[u] = Pp2;
            }

            // ---- these properties are done with the shared rows of plane k ----
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + (gsh % RING_SS));
            dtv_m = dtv_c; dtv_c = dtv_p;
            rdz_c = rdz_p; rdz_p = rdz_pp;
            qz_c = qz_p;
            q += sk;
            pst = pstB;
        }
        // plane K+1 was only read as the upper neighbour
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + (gsh % RING_SS));
        ++gsh;
        pst = pnext(pst);

        // ---------------- back substitution (MF:4100-4105) ----------------
        if (writer && colwet) {
#pragma unroll
            for (int u = 0; u < NPT; ++u) {
                if (!live[u]) continue;
                double *__restrict__ O = s.p[pn[u]].pout;
                int qo = c2d + sk * (K + 1);
                double x = 0.0;                               // RES(KUB+1) = G(KUB+1) = 0 (halo row is the identity)
                O[qo] = x;
                int k = K;
                for (; k >= 8; k -= 8) {
                    double gq[8];
#pragma unroll
                    for (int v = 0; v < 8; ++v) gq[v] = O[qo - (v + 1) * sk];
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        qo -= sk;
                        x = Wsm[u][(size_t)(k - 1 - v) * wstride] * x + gq[v];
                        O[qo] = x;
                    }
                }
                for (; k >= 1; --k) {
                    qo -= sk;
                    x = Wsm[u][(size_t)(k - 1) * wstride] * x + O[qo];
                    O[qo] = x;
                }
            }
            if (zp) atomicAdd(s.zero_pivots, (unsigned long long)zp);
        }
    }
    cp_async_wait<0>();
}

}  // namespace adt
