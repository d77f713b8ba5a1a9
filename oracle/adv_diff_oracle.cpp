// =====================================================================================
//  adv_diff_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never a product path)
//
//  A C++17 restatement of MOHID's ModuleAdvectionDiffusion::AdvectionDiffusion and of the
//  ModuleFunctions routines it calls, following the reference pass by pass, loop by loop
//  and operation by operation (same evaluation order, fp64 everywhere, no FMA contraction:
//  build with -ffp-contract=off).  The reference's OpenMP loop structure is mirrored so
//  the same file doubles as the "CPU restatement of the reference path" timing baseline.
//
//  PARITY UNPINNED except for the column solve: the reference ships no golden vectors /
//  known-answer tests for this path and its Fortran cannot be compiled
//  here (no Fortran compiler, HDF5, MPI), so this oracle is pinned by (a) line-by-line
//  citation of the source below, (b) the self-consistency properties in tests/test_oracle_*.py
//  and (c) for THOMASZ_NewType2 only, the reference's own CUDA solver (CudaThomas/Thomas.cu,
//  compiled unmodified into oracle/_ref by oracle/Makefile; tests/test_gpu_ref_thomas.py).
//
//  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
//  legs may load this file's shared library.
//
//  Reference files (relative to /root/reference/Software):
//     AD  = MOHIDBase2/ModuleAdvectionDiffusion.F90
//     MF  = MOHIDBase1/ModuleFunctions.F90
//     MGD = MOHIDBase1/ModuleGlobalData.F90
//     WP  = MOHIDWater/ModuleWaterProperties.F90
//
//  Not restated (returns ORACLE_ERR_UNSUPPORTED): AdvectionNudging (AD:1989-2068; it reads
//  uninitialised indices in the reference) and the decomposed form of the horizontally implicit
//  solve (THOMAS_DDecompHorizGrid, HG:8245-8478); its 2-D (K = 1) form (AD:1758-1841) is.  Horizontally implicit advection
//  in 3-D (AD:4167-4258, THOMAS_3D MF:3667-3875) and the Orlanski boundary (MF:4129-4500) are.
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/mohid_adt.h"

namespace {

constexpr double null_real = MOHID_NULL_REAL;   // MGD:143
constexpr double MinValue  = 1.e-16;            // MGD:1812
constexpr int    Compute   = 1;                 // MGD:1729
constexpr int    OpenPoint = 1;                 // MGD:1736
constexpr double ExplicitScheme = 0.;           // MGD:1831
constexpr double ImplicitScheme = 1.;           // MGD:1832

constexpr int ORACLE_ERR_UNSUPPORTED = MOHID_ADT_ERR_UNSUPPORTED;
constexpr int ORACLE_ERR_ARG         = MOHID_ADT_ERR_ARG;

struct Oracle {
    // ---- sizes (T_Size3D Size / WorkSize, AD:165-381) ----
    mohid_adt_size3d S{}, W{};
    long ld = 0, nj = 0, nk = 0, n3 = 0, n2 = 0;
    mohid_adt_options opt{};
    int nthreads = 1;

    // ---- ExternalVar: borrowed pointers (AD:1271-1338, 1353-1401) ----
    const double *DUX = nullptr, *DVY = nullptr, *DZX = nullptr, *DZY = nullptr;
    const int *KFloorZ = nullptr, *BoundaryPoints2D = nullptr;
    const double *Wflux_X = nullptr, *Wflux_Y = nullptr, *Wflux_Z = nullptr;
    const double *VolumeZOld = nullptr, *VolumeZ = nullptr, *Visc_H = nullptr, *Diff_V = nullptr;
    const double *DWZ = nullptr, *DZZ = nullptr, *AreaU = nullptr, *AreaV = nullptr;
    const int *OpenPoints3D = nullptr, *LandPoints3D = nullptr, *WaterPoints3D = nullptr;
    const int *ComputeFacesU3D = nullptr, *ComputeFacesV3D = nullptr, *ComputeFacesW3D = nullptr;
    const int *SmallDepths = nullptr;            // nullptr => SmallDepthsPresent = .false.
    const int *NoFluxU = nullptr, *NoFluxV = nullptr, *NoFluxW = nullptr;
    double *PROP = nullptr;
    const double *ReferenceProp = nullptr;

    // discharges (AD:978-1034)
    bool DischON = false;
    int DischNumber = 0;
    std::vector<double> DischFlow, DischConc, DischConcMF;
    std::vector<int> DischI, DischJ, DischK, DischKmin, DischKmax, DischVert, IgnoreDisch, DischnCells, ByPass;

    // ---- scalars of the current call ----
    mohid_adt_params P{};
    bool Optimize = false, FirstProperty = true;
    // remembered from the previous call (Set_Internal_State compares against them, AD:5765-5807)
    bool   have_prev = false;
    double prev_DTProp = null_real, prev_SchmidtCoef_V = null_real, prev_SchmidtBackground_V = null_real,
           prev_Schmidt_H = null_real;
    int    prev_AdvMethodH = -1, prev_TVDLimitationH = -1, prev_AdvMethodV = -1, prev_TVDLimitationV = -1;
    double LastCalc = -1e300, Now = 0.;

    // ---- T_State (AD:5746-5835) ----
    bool st_VertAdv = true, st_HorAdv = true, st_VertDif = true, st_HorDif = true, st_OpenBoundary = false;

    // ---- work arrays, all initialised to Null_real (AD:564-680) ----
    std::vector<double> DifX, DifY, DifZ;                 // Diffusion_CoeficientX/Y/Z
    std::vector<double> D, E, F, TI;                       // COEF3%D/E/F, TICOEF3
    std::vector<double> VC, VD, VE, VF;                    // COEF3_VertAdv
    std::vector<double> XC, XD, XE, XF;                    // COEF3_HorAdvXX
    std::vector<double> YC, YD, YE, YF;                    // COEF3_HorAdvYY
    std::vector<double> QB;                                // WaterFluxOBoundary
    std::vector<double> AdvFluxX, AdvFluxY, AdvFluxZ, DifFluxX, DifFluxY, DifFluxZ;   // T_CellFluxes (AD:229-240)
    bool st_CellFluxes = false;
    std::vector<double> DHU, DHV, DVC;                     // Diff_H_Const_U/V, Diff_V_Const (AD:1474-1477)
    bool FirstTime = true;
    std::vector<std::vector<double>> VEC_G, VEC_W;         // per-thread Thomas scratch (AD:626-637)

    std::string err;
    long zero_pivots = 0;

    inline long i3(int i, int j, int k) const {
        return (long)(i - S.ILB) + ld * ((long)(j - S.JLB) + nj * (long)(k - S.KLB));
    }
    inline long i2(int i, int j) const { return (long)(i - S.ILB) + ld * (long)(j - S.JLB); }
    inline bool SmallDepthCell(int i, int j) const {       // AD:2660-2675
        return SmallDepths ? (SmallDepths[i2(i, j)] != 0) : false;
    }
};

// MF:14038-14104: CHUNK = max((UB-LB)/ChunkFactor, 1), factor 99999 (MGD:2167-2169)
inline int chunk_of(int lb, int ub) { return std::max((ub - lb) / 99999, 1); }

// ---------------------------------------------------------------------------------------
// SetMatrixValue (MF:1599-1740): constant fill over Size, optionally where MapMatrix == 1
// ---------------------------------------------------------------------------------------
void SetMatrixValue(Oracle &o, std::vector<double> &M, double v, const int *map = nullptr) {
    const int CH = chunk_of(o.S.KLB, o.S.KUB);
#pragma omp parallel for schedule(dynamic, CH) num_threads(o.nthreads)
    for (int k = o.S.KLB; k <= o.S.KUB; ++k)
        for (int j = o.S.JLB; j <= o.S.JUB; ++j)
            for (int i = o.S.ILB; i <= o.S.IUB; ++i) {
                long q = o.i3(i, j, k);
                if (!map || map[q] == 1) M[q] = v;
            }
}

// ---------------------------------------------------------------------------------------
// MF:11045-11141 helpers
// ---------------------------------------------------------------------------------------
inline double Courant(double QFace, double V, double dt) { return QFace * dt / V; }   // MF:11045-11053

inline void FaceConcUpFirstOrder(double Coef[4], double QFace) {                       // MF:11127-11141
    Coef[0] = Coef[1] = Coef[2] = Coef[3] = 0.;
    if (QFace > 0.) Coef[1] = 1.; else Coef[2] = 1.;
}
inline void FaceConcUpSecondOrder(double Coef[4], double QFace) {                      // MF:11055-11084
    Coef[0] = Coef[1] = Coef[2] = Coef[3] = 0.;
    if (QFace > 0) {
        Coef[0] = -1. / 8.; Coef[1] = 6. / 8.; Coef[2] = 3. / 8.;
    } else if (QFace < 0) {
        Coef[3] = -1. / 8.; Coef[2] = 6. / 8.; Coef[1] = 3. / 8.;
    } else {
        FaceConcUpFirstOrder(Coef, QFace);
    }
}
inline void FaceConcUpThirdOrder(double Coef[4], double QFace, double Cr) {            // MF:11086-11125
    double c = (1 - 2. * std::fabs(Cr)) / 6.;
    double a = 0.5 + c;
    double b = 0.5 - c;
    double d = (1 - std::fabs(Cr)) / 2.;
    if (QFace > 0) {
        Coef[0] = -d * b; Coef[1] = 1 + d * (b - a); Coef[2] = d * a; Coef[3] = 0.;
    } else if (QFace < 0) {
        Coef[3] = -d * b; Coef[2] = 1 + d * (b - a); Coef[1] = d * a; Coef[0] = 0.;
    } else {
        FaceConcUpFirstOrder(Coef, QFace);
    }
}

// ---------------------------------------------------------------------------------------
// ComputeAdvectionFace (MF:10702-10894).  Prop/du/V are the 4-point stencils (f-2,f-1,f,f+1).
// Returns false on the reference's `stop` branches.
// ---------------------------------------------------------------------------------------
bool ComputeAdvectionFace(const double Prop[4], const double V[4], const double du[4], double dt,
                          double QFace, double VolumeRelMax, int Method, int TVD_Limitation,
                          bool NearBoundary, bool Upwind2, double CFace[4]) {
    double Cr = 0., Theta, dC, r, a, b, AuxLeft, AuxRight;
    double Cup1[4], CupHighOrder[4];
    double Aux, VolumeRel = 0.;

    FaceConcUpFirstOrder(Cup1, QFace);

    if (!NearBoundary) {
        if (QFace > 0) {
            Cr = Courant(QFace, V[1], dt);
            Aux = std::min(std::min(V[0], V[1]), V[2]);
            VolumeRel = std::max(std::max(V[0], V[1]), V[2]) / Aux;
        } else {
            Cr = Courant(QFace, V[2], dt);
            Aux = std::min(std::min(V[1], V[2]), V[3]);
            VolumeRel = std::max(std::max(V[1], V[2]), V[3]) / Aux;
        }
    }

    if (Method == MOHID_UpwindOrder1 || (NearBoundary && Upwind2)) {
        for (int n = 0; n < 4; ++n) CupHighOrder[n] = Cup1[n];
        Theta = 0.;
    } else if (Method == MOHID_UpwindOrder2 && !NearBoundary) {
        FaceConcUpSecondOrder(CupHighOrder, QFace);
        Theta = (VolumeRel > VolumeRelMax) ? 0. : 1.;
    } else if (Method == MOHID_UpwindOrder3 && !NearBoundary) {
        FaceConcUpThirdOrder(CupHighOrder, QFace, Cr);
        Theta = (VolumeRel > VolumeRelMax) ? 0. : 1.;
    } else if (Method == MOHID_CentralDif || Method == MOHID_LeapFrog) {
        AuxLeft  = du[2] / (du[1] + du[2]);
        AuxRight = du[1] / (du[1] + du[2]);
        CupHighOrder[0] = 0.; CupHighOrder[1] = AuxLeft; CupHighOrder[2] = AuxRight; CupHighOrder[3] = 0.;
        Theta = 1.;
    } else if (Method == MOHID_P2_TVD && !NearBoundary) {
        CupHighOrder[0] = CupHighOrder[1] = CupHighOrder[2] = CupHighOrder[3] = 0.;
        if (QFace > 0.) {
            CupHighOrder[2] = 1.;
            dC = (Prop[2] - Prop[1]) / (du[2] + du[1]);
            if (std::fabs(dC) < MinValue) dC = (dC >= 0) ? MinValue : -MinValue;
            r = (Prop[1] - Prop[0]) / (du[1] + du[0]) / dC;
        } else {
            CupHighOrder[1] = 1.;
            dC = (Prop[1] - Prop[2]) / (du[2] + du[1]);
            if (std::fabs(dC) < MinValue) dC = (dC >= 0) ? MinValue : -MinValue;
            r = (Prop[2] - Prop[3]) / (du[2] + du[3]) / dC;
        }
        if (TVD_Limitation == MOHID_MinMod) {
            Theta = std::max(0., std::min(1., r));
        } else if (TVD_Limitation == MOHID_VanLeer) {
            Theta = (r < 0) ? 0. : 2. * r / (1 + r);
        } else if (TVD_Limitation == MOHID_Muscl) {
            Theta = std::max(0., std::min(std::min(2., 2. * r), (1 + r) / 2.));
        } else if (TVD_Limitation == MOHID_SuperBee) {
            Theta = std::max(std::max(0., std::min(1., 2. * r)), std::min(r, 2.));
        } else if (TVD_Limitation == MOHID_PDM) {
            a = 0.5 + (1 - 2. * std::fabs(Cr)) / 6.;
            b = 0.5 - (1 - 2. * std::fabs(Cr)) / 6.;
            Aux = a + b * r;
            if (std::fabs(Cr) < MinValue) Cr = MinValue;
            Theta = std::max(0., std::min(std::min(Aux, 2. / (1. - Cr)), 2. * r / Cr));
        } else {
            return false;   // "This TVD Limitation option is not valid to compute Advection1D"
        }
        Theta = 0.5 * Theta * (1. - Cr);
    } else {
        return false;       // "This method is not valid to compute Advection1D"
    }

    // (theta deliberately not clamped, MF:10875-10887)
    for (int n = 0; n < 4; ++n) CFace[n] = (1. - Theta) * Cup1[n] + Theta * CupHighOrder[n];
    return true;
}

// MF:10968-11005  (QFace > 0, interior)
inline void ComputeAdvectionFace_TVD_Superbee_1(const double Prop[4], const double V[4], const double du[4],
                                                double dt, double QFace, double CFace[4]) {
    double Cr = Courant(QFace, V[1], dt);
    double dC = (Prop[2] - Prop[1]) / (du[2] + du[1]);
    if (std::fabs(dC) < MinValue) dC = (dC >= 0) ? MinValue : -MinValue;
    double r = (Prop[1] - Prop[0]) / (du[1] + du[0]) / dC;
    double Theta = std::max(std::max(0., std::min(1., 2. * r)), std::min(r, 2.));
    Theta = 0.5 * Theta * (1. - Cr);
    CFace[1] = (1. - Theta);
    CFace[2] = Theta;
}
// MF:11007-11043  (QFace <= 0, interior)
inline void ComputeAdvectionFace_TVD_Superbee_2(const double Prop[4], const double V[4], const double du[4],
                                                double dt, double QFace, double CFace[4]) {
    double Cr = Courant(QFace, V[2], dt);
    double dC = (Prop[1] - Prop[2]) / (du[2] + du[1]);
    if (std::fabs(dC) < MinValue) dC = (dC >= 0) ? MinValue : -MinValue;
    double r = (Prop[2] - Prop[3]) / (du[2] + du[3]) / dC;
    double Theta = std::max(std::max(0., std::min(1., 2. * r)), std::min(r, 2.));
    Theta = 0.5 * Theta * (1. - Cr);
    CFace[1] = Theta;
    CFace[2] = (1. - Theta);
}

// ---------------------------------------------------------------------------------------
// A 1-D line view: element with ACTUAL index a (a = dummy index - 1, since the reference
// passes 0-based array sections to assumed-shape dummies, AD:4419, 2986) is p[a*stride].
// ---------------------------------------------------------------------------------------
struct LineD { const double *p; long s; inline double operator()(int a) const { return p[(long)a * s]; } };
struct LineI { const int *p; long s; inline int operator()(int a) const { return p[(long)a * s]; } };
struct LineW { double *p; long s; inline double &operator()(int a) const { return p[(long)a * s]; } };

// ComputeAdvection1D_V2 (MF:10534-10591); called with (LB+1, UB+1) in dummy indices, i.e. the
// loop covers actual faces a = LB..UB.
bool ComputeAdvection1D_V2(int lb, int ub, double dt, LineD du, LineD Prop, LineD Q, LineD V, LineI CP,
                           LineW C_flux, LineW D_flux, LineW E_flux, LineW F_flux, int Method,
                           int TVD_Limitation, double VolumeRelMax, bool Upwind2) {
    for (int a = lb; a <= ub; ++a) {
        if (CP(a - 1) == Compute && CP(a) == Compute) {
            double QFace = Q(a);
            bool NearBoundary = false;
            if (QFace > 0) {
                if (CP(a - 2) != Compute) NearBoundary = true;
            } else {
                if (CP(a + 1) != Compute) NearBoundary = true;
            }
            double Prop4[4] = {Prop(a - 2), Prop(a - 1), Prop(a), Prop(a + 1)};
            double du4[4]   = {du(a - 2), du(a - 1), du(a), du(a + 1)};
            double V4[4]    = {V(a - 2), V(a - 1), V(a), V(a + 1)};
            double CFace[4];
            if (!ComputeAdvectionFace(Prop4, V4, du4, dt, QFace, VolumeRelMax, Method, TVD_Limitation,
                                      NearBoundary, Upwind2, CFace))
                return false;
            C_flux(a) = QFace * CFace[0];
            D_flux(a) = QFace * CFace[1];
            E_flux(a) = QFace * CFace[2];
            F_flux(a) = QFace * CFace[3];
        }
    }
    return true;
}

// ComputeAdvection1D_TVD_SuperBee_2 (MF:10642-10699)
void ComputeAdvection1D_TVD_SuperBee_2(int lb, int ub, double dt, LineD du, LineD Prop, LineD Q, LineD V,
                                       LineI CP, LineW D_flux, LineW E_flux) {
    for (int a = lb; a <= ub; ++a) {
        if (CP(a - 1) == Compute && CP(a) == Compute) {
            double QFace = Q(a);
            double Prop4[4] = {0, 0, 0, 0}, du4[4] = {0, 0, 0, 0}, V4[4] = {0, 0, 0, 0}, CFace[4] = {0, 0, 0, 0};
            if (QFace > 0) {
                if (CP(a - 2) != Compute) {
                    D_flux(a) = QFace;          // NearBoundary: CFace(2) = 1; E_flux keeps its value
                } else {
                    V4[1] = V(a - 1);
                    Prop4[0] = Prop(a - 2); Prop4[1] = Prop(a - 1); Prop4[2] = Prop(a);
                    du4[0] = du(a - 2); du4[1] = du(a - 1); du4[2] = du(a);
                    ComputeAdvectionFace_TVD_Superbee_1(Prop4, V4, du4, dt, QFace, CFace);
                    D_flux(a) = QFace * CFace[1];
                    E_flux(a) = QFace * CFace[2];
                }
            } else {
                if (CP(a + 1) != Compute) {
                    E_flux(a) = QFace;          // NearBoundary: CFace(3) = 1; D_flux keeps its value
                } else {
                    V4[2] = V(a);
                    Prop4[1] = Prop(a - 1); Prop4[2] = Prop(a); Prop4[3] = Prop(a + 1);
                    du4[1] = du(a - 1); du4[2] = du(a); du4[3] = du(a + 1);
                    ComputeAdvectionFace_TVD_Superbee_2(Prop4, V4, du4, dt, QFace, CFace);
                    D_flux(a) = QFace * CFace[1];
                    E_flux(a) = QFace * CFace[2];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Set_Internal_State (AD:5746-5835)
// ---------------------------------------------------------------------------------------
int Set_Internal_State(Oracle &o, const mohid_adt_params &p) {
    if (!o.have_prev || o.prev_DTProp != p.DTProp || o.LastCalc != o.Now) {
        o.st_VertAdv = o.st_HorAdv = o.st_VertDif = o.st_HorDif = true;
    } else {
        o.st_HorAdv = (o.prev_AdvMethodH != p.AdvMethodH) || (o.prev_TVDLimitationH != p.TVDLimitationH) ||
                      (o.prev_AdvMethodH == MOHID_P2_TVD);
        o.st_VertAdv = (o.prev_AdvMethodV != p.AdvMethodV) || (o.prev_TVDLimitationV != p.TVDLimitationV) ||
                       (o.prev_AdvMethodV == MOHID_P2_TVD);
        o.st_VertDif = (o.prev_SchmidtCoef_V != p.SchmidtCoef_V) ||
                       (o.prev_SchmidtBackground_V != p.SchmidtBackground_V) || p.NullDif;
        o.st_HorDif = (o.prev_Schmidt_H != p.Schmidt_H) || p.NullDif;
    }
    if (o.ReferenceProp) {
        o.st_OpenBoundary = true;
        const int bc = p.BoundaryCondition;
        if (bc != MOHID_BC_MassConservation && bc != MOHID_BC_ImposedValue && bc != MOHID_BC_SubModel &&
            bc != MOHID_BC_Orlanski && bc != MOHID_BC_NullGradient && bc != MOHID_BC_CyclicBoundary &&
            bc != MOHID_BC_MassConservNullGrad) {
            o.err = "Set_Internal_State - ModuleAdvectionDiffusion - ERR01";
            return ORACLE_ERR_ARG;
        }
    } else {
        o.st_OpenBoundary = false;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// Convert_Dif_Vertical (AD:2364-2443)
// ---------------------------------------------------------------------------------------
void Convert_Dif_Vertical(Oracle &o) {
    const auto &W = o.W;
    const int CH = chunk_of(W.JLB, W.JUB);
#pragma omp parallel num_threads(o.nthreads)
    {
        for (int k = W.KLB + 1; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH)
            for (int j = W.JLB; j <= W.JUB; ++j)
                for (int i = W.ILB; i <= W.IUB; ++i) {
                    long q = o.i3(i, j, k);
                    if (o.ComputeFacesW3D[q] == 1)
                        o.DifZ[q] = (o.P.SchmidtCoef_V * o.Diff_V[q] + o.P.SchmidtBackground_V);
                }
        }
        if (o.P.NullDif) {
            for (int k = W.KLB + 1; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH) nowait
                for (int j = W.JLB; j <= W.JUB; ++j)
                    for (int i = W.ILB; i <= W.IUB; ++i) {
                        long q = o.i3(i, j, k);
                        if (o.ComputeFacesW3D[q] == 1 && o.Wflux_Z[q] == 0.) o.DifZ[q] = 0.;
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Convert_Visc_Dif_Horizontal / _opt (AD:2453-2675)
// ---------------------------------------------------------------------------------------
void Convert_Visc_Dif_Horizontal(Oracle &o, bool opt) {
    const auto &W = o.W;
    const int CH = chunk_of(W.JLB, W.JUB);
    const bool noDif = !opt && o.P.NoDifFlux;
#pragma omp parallel num_threads(o.nthreads)
    {
        for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH) nowait
            for (int j = W.JLB; j <= W.JUB; ++j)
                for (int i = W.ILB; i <= W.IUB; ++i) {
                    long q = o.i3(i, j, k);
                    if (o.ComputeFacesU3D[q] == 1) {
                        o.DifX[q] = o.P.Schmidt_H *
                                    (o.Visc_H[q] * o.DUX[o.i2(i, j - 1)] + o.Visc_H[o.i3(i, j - 1, k)] * o.DUX[o.i2(i, j)]) /
                                    (o.DUX[o.i2(i, j)] + o.DUX[o.i2(i, j - 1)]);
                        if (noDif && o.NoFluxU && o.NoFluxU[q] == 1) o.DifX[q] = 0.;
                    }
                }
        }
#pragma omp barrier
        for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH)
            for (int j = W.JLB; j <= W.JUB; ++j)
                for (int i = W.ILB; i <= W.IUB; ++i) {
                    long q = o.i3(i, j, k);
                    if (o.ComputeFacesV3D[q] == 1) {
                        o.DifY[q] = o.P.Schmidt_H *
                                    (o.Visc_H[q] * o.DVY[o.i2(i - 1, j)] + o.Visc_H[o.i3(i - 1, j, k)] * o.DVY[o.i2(i, j)]) /
                                    (o.DVY[o.i2(i, j)] + o.DVY[o.i2(i - 1, j)]);
                        if (noDif && o.NoFluxV && o.NoFluxV[q] == 1) o.DifY[q] = 0.;
                    }
                }
        }
        if (!opt && o.P.NullDif) {
            for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH) nowait
                for (int j = W.JLB; j <= W.JUB; ++j)
                    for (int i = W.ILB; i <= W.IUB; ++i) {
                        long q = o.i3(i, j, k);
                        if (o.ComputeFacesU3D[q] == 1 && o.Wflux_X[q] == 0.) o.DifX[q] = 0.;
                    }
            }
            for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH) nowait
                for (int j = W.JLB; j <= W.JUB; ++j)
                    for (int i = W.ILB; i <= W.IUB; ++i) {
                        long q = o.i3(i, j, k);
                        if (o.ComputeFacesV3D[q] == 1 && o.Wflux_Y[q] == 0.) o.DifY[q] = 0.;
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Compute_DifH_Constants / Compute_DifV_Constants (AD:1514-1619), Optimize path only
// ---------------------------------------------------------------------------------------
void Compute_DifH_Constants(Oracle &o) {
    const auto &S = o.S;
    const int CH = chunk_of(o.W.KLB, o.W.KUB);
#pragma omp parallel num_threads(o.nthreads)
    {
#pragma omp for schedule(dynamic, CH)
        for (int k = S.KLB; k <= S.KUB; ++k)
            for (int j = S.JLB; j <= S.JUB; ++j)
                for (int i = S.ILB; i <= S.IUB; ++i) {
                    long q = o.i3(i, j, k);
                    if (o.ComputeFacesV3D[q] == 1) o.DHV[q] = o.DifY[q] * o.AreaV[q] / o.DZY[o.i2(i - 1, j)];
                }
#pragma omp for schedule(dynamic, CH)
        for (int k = S.KLB; k <= S.KUB; ++k)
            for (int j = S.JLB; j <= S.JUB; ++j)
                for (int i = S.ILB; i <= S.IUB; ++i) {
                    long q = o.i3(i, j, k);
                    if (o.ComputeFacesU3D[q] == 1) o.DHU[q] = o.DifX[q] * o.AreaU[q] / o.DZX[o.i2(i, j - 1)];
                }
    }
}
void Compute_DifV_Constants(Oracle &o) {
    const auto &W = o.W;
    const int CH = chunk_of(W.KLB, W.KUB);
#pragma omp parallel for schedule(dynamic, CH) num_threads(o.nthreads)
    for (int k = W.KLB; k <= W.KUB; ++k)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                long q = o.i3(i, j, k);
                if (o.ComputeFacesW3D[q] == 1) {
                    if (!o.SmallDepths || !o.SmallDepths[o.i2(i, j)]) {
                        double AuxK = o.DifZ[q] * o.DUX[o.i2(i, j)] * o.DVY[o.i2(i, j)];
                        o.DVC[q] = AuxK / o.DZZ[o.i3(i, j, k - 1)];
                    }
                }
            }
}

// ---------------------------------------------------------------------------------------
// VolumeVariation (AD:3966-4021)
// ---------------------------------------------------------------------------------------
void VolumeVariation(Oracle &o) {
    const auto &W = o.W;
    const int CH = chunk_of(W.JLB, W.JUB);
    const double DT = o.P.DTProp;
#pragma omp parallel num_threads(o.nthreads)
    for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                long q = o.i3(i, j, k);
                if (o.OpenPoints3D[q] == 1) {
                    o.TI[q] = o.PROP[q] * (o.VolumeZOld[q] / o.VolumeZ[q]);
                    if (k == W.KUB) {
                        double DT_V = DT / o.VolumeZ[q];
                        o.E[q] = 1.0 + DT_V * o.Wflux_Z[o.i3(i, j, k + 1)];
                    } else if (o.Optimize) {
                        o.E[q] = 1.0;
                    }
                } else {
                    o.TI[q] = o.PROP[q];
                }
            }
    }
}

// ---------------------------------------------------------------------------------------
// Discharges (AD:4025-4128) -- serial
// ---------------------------------------------------------------------------------------
void Discharges(Oracle &o) {
    int n = 0;
    const double DT = o.P.DTProp;
    for (int dis = 0; dis < o.DischNumber; ++dis) {
        if (o.IgnoreDisch[dis]) continue;
        for (int nc = 0; nc < o.DischnCells[dis]; ++nc) {
            int i = o.DischI[n], j = o.DischJ[n], kd = o.DischK[n];
            int kmin = o.DischKmin[n], kmax = o.DischKmax[n];
            if (o.DischVert[dis] == MOHID_DischUniform) {
                if (kmin == MOHID_FILL_INT) kmin = o.KFloorZ[o.i2(i, j)];
                if (kmax == MOHID_FILL_INT) kmax = o.W.KUB;
            } else {
                kmin = kd; kmax = kd;
            }
            double WaterColumn = 0.0;
            for (int k = kmin; k <= kmax; ++k) WaterColumn = WaterColumn + o.DWZ[o.i3(i, j, k)];
            for (int k = kmin; k <= kmax; ++k) {
                long q = o.i3(i, j, k);
                double DT_V = DT / o.VolumeZ[q];
                double Flow;
                if (o.DischVert[dis] == MOHID_DischUniform) Flow = o.DischFlow[n] * o.DWZ[q] / WaterColumn;
                else Flow = o.DischFlow[n];
                if (o.OpenPoints3D[q] == OpenPoint) {
                    double Aux_Conc = o.DischConc[n];
                    if (o.ByPass[dis]) {
                        o.TI[q] = o.TI[q] + Flow * DT_V * Aux_Conc;
                    } else if (Flow > 0.) {
                        o.TI[q] = o.TI[q] + Flow * DT_V * Aux_Conc;
                    } else {
                        double Aux_MF = o.DischConcMF[n];
                        o.E[q] = o.E[q] - Flow * DT_V * Aux_MF;
                    }
                } else {
                    if (Flow > 0)
                        o.TI[q] = o.PROP[q] * o.VolumeZOld[q] / o.VolumeZ[q] + Flow * DT_V * o.DischConc[n];
                }
            }
            ++n;
        }
    }
}

// ---------------------------------------------------------------------------------------
// HorizontalDiffusionXX / YY (AD:5156-5210, 5317-5365) and XX2 / YY2 (AD:5212-5315)
// ---------------------------------------------------------------------------------------
void HorizontalDiffusionXX(Oracle &o) {
    const auto &W = o.W;
    const int CH = chunk_of(W.KLB, W.KUB);
    const double DTPropDouble = o.P.DTProp;
#pragma omp parallel for schedule(dynamic, CH) num_threads(o.nthreads)
    for (int k = W.KLB; k <= W.KUB; ++k)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                long q = o.i3(i, j, k);
                if (o.ComputeFacesU3D[q] == 1) {
                    long qm = o.i3(i, j - 1, k);
                    double AuxJ = o.DifX[q] * o.AreaU[q] / o.DZX[o.i2(i, j - 1)];
                    o.TI[qm] = o.TI[qm] + AuxJ * DTPropDouble / o.VolumeZ[qm] * (o.PROP[q] - o.PROP[qm]);
                    o.TI[q]  = o.TI[q]  - AuxJ * DTPropDouble / o.VolumeZ[q]  * (o.PROP[q] - o.PROP[qm]);
                }
            }
}
void HorizontalDiffusionXX2(Oracle &o) {
    const auto &W = o.W;
    const int CH = chunk_of(W.KLB, W.KUB);
    const double DTPropDouble = o.P.DTProp;
#pragma omp parallel for schedule(dynamic, CH) num_threads(o.nthreads)
    for (int k = W.KLB; k <= W.KUB; ++k)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                long q = o.i3(i, j, k);
                if (o.ComputeFacesU3D[q] == 1) {
                    long qm = o.i3(i, j - 1, k);
                    double Gradient = o.PROP[q] - o.PROP[qm];
                    double AuxJ = o.DHU[q] * Gradient * DTPropDouble;
                    o.TI[qm] = o.TI[qm] + AuxJ / o.VolumeZ[qm];
                    o.TI[q]  = o.TI[q]  - AuxJ / o.VolumeZ[q];
                }
            }
}
void HorizontalDiffusionYY2(Oracle &o) {
    const auto &W = o.W;
    const int CH = chunk_of(W.KLB, W.KUB);
    const double DTPropDouble = o.P.DTProp;
#pragma omp parallel for schedule(dynamic, CH) num_threads(o.nthreads)
    for (int k = W.KLB; k <= W.KUB; ++k)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                long q = o.i3(i, j, k);
                if (o.ComputeFacesV3D[q] == 1) {
                    long qm = o.i3(i - 1, j, k);
                    double Gradient = o.PROP[q] - o.PROP[qm];
                    double AuxI = o.DHV[q] * DTPropDouble * Gradient;
                    o.TI[qm] = o.TI[qm] + AuxI / o.VolumeZ[qm];
                    o.TI[q]  = o.TI[q]  - AuxI / o.VolumeZ[q];
                }
            }
}
void HorizontalDiffusionYY(Oracle &o) {
    const auto &W = o.W;
    const int CH = chunk_of(W.JLB, W.JUB);
    const double DTPropDouble = o.P.DTProp;
#pragma omp parallel num_threads(o.nthreads)
    for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                long q = o.i3(i, j, k);
                if (o.ComputeFacesV3D[q] == 1) {
                    long qm = o.i3(i - 1, j, k);
                    double AuxI = o.DifY[q] * o.AreaV[q] / o.DZY[o.i2(i - 1, j)];
                    o.TI[qm] = o.TI[qm] + AuxI * DTPropDouble / o.VolumeZ[qm] * (o.PROP[q] - o.PROP[qm]);
                    o.TI[q]  = o.TI[q]  - AuxI * DTPropDouble / o.VolumeZ[q]  * (o.PROP[q] - o.PROP[qm]);
                }
            }
    }
}

// ---------------------------------------------------------------------------------------
// HorizontalAdvectionXX (AD:4368-4527) + _Explicit (AD:4544-4582)
// ---------------------------------------------------------------------------------------
int HorizontalAdvectionXX(Oracle &o) {
    const auto &W = o.W;
    const double DT = o.P.DTProp;
    const long sj = o.ld;       // stride of j in 3-D and 2-D arrays
    int bad = 0;
    if (o.st_HorAdv) {
        const int CH = chunk_of(W.ILB, W.IUB);
#pragma omp parallel num_threads(o.nthreads)
        for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                const long b3 = o.i3(i, o.S.JLB, k) - (long)o.S.JLB * sj;   // address of actual j = 0
                const long b2 = o.i2(i, o.S.JLB) - (long)o.S.JLB * sj;
                LineD du{o.DUX + b2, sj}, Prop{o.PROP + b3, sj}, Q{o.Wflux_X + b3, sj}, V{o.VolumeZ + b3, sj};
                LineI CP{o.OpenPoints3D + b3, sj};
                if (o.Optimize) {
                    ComputeAdvection1D_TVD_SuperBee_2(W.JLB, W.JUB, DT, du, Prop, Q, V, CP,
                                                      LineW{o.XD.data() + b3, sj}, LineW{o.XE.data() + b3, sj});
                } else {
                    if (!ComputeAdvection1D_V2(W.JLB, W.JUB, DT, du, Prop, Q, V, CP, LineW{o.XC.data() + b3, sj},
                                               LineW{o.XD.data() + b3, sj}, LineW{o.XE.data() + b3, sj},
                                               LineW{o.XF.data() + b3, sj}, o.P.AdvMethodH, o.P.TVDLimitationH,
                                               o.P.VolumeRelMax, o.P.Upwind2H != 0)) {
#pragma omp atomic write
                        bad = 1;
                    }
                    if (o.P.NoAdvFlux && o.NoFluxU) {
                        for (int j = W.JLB; j <= W.JUB; ++j) {
                            long q = o.i3(i, j, k);
                            if (o.NoFluxU[q] == 1) o.XC[q] = o.XD[q] = o.XE[q] = o.XF[q] = 0.;
                        }
                    }
                }
            }
        }
    }
    if (bad) { o.err = "This method is not valid to compute Advection1D"; return ORACLE_ERR_ARG; }

    const int CH = chunk_of(W.KLB, W.KUB);
    if (o.P.ImpExp_AdvXX == ExplicitScheme) {
        if (o.Optimize) {
#pragma omp parallel for schedule(dynamic, CH) num_threads(o.nthreads)
            for (int k = W.KLB; k <= W.KUB; ++k)
                for (int j = W.JLB; j <= W.JUB; ++j)
                    for (int i = W.ILB; i <= W.IUB; ++i) {
                        long q = o.i3(i, j, k);
                        if (o.ComputeFacesU3D[q] == 1) {
                            long qm = o.i3(i, j - 1, k);
                            double AdvFluxX = (o.XD[q] * o.PROP[qm] + o.XE[q] * o.PROP[q]);
                            o.TI[q]  = o.TI[q]  + AdvFluxX * DT / o.VolumeZ[q];
                            o.TI[qm] = o.TI[qm] - AdvFluxX * DT / o.VolumeZ[qm];
                        }
                    }
        } else {
#pragma omp parallel for schedule(dynamic, CH) num_threads(o.nthreads)
            for (int k = W.KLB; k <= W.KUB; ++k)
                for (int j = W.JLB; j <= W.JUB; ++j)
                    for (int i = W.ILB; i <= W.IUB; ++i) {
                        long q = o.i3(i, j, k);
                        if (o.ComputeFacesU3D[q] == 1) {
                            long qm = o.i3(i, j - 1, k);
                            double AdvFluxX = (o.XC[q] * o.PROP[o.i3(i, j - 2, k)] + o.XD[q] * o.PROP[qm] +
                                               o.XE[q] * o.PROP[q] + o.XF[q] * o.PROP[o.i3(i, j + 1, k)]);
                            o.TI[q]  = o.TI[q]  + AdvFluxX * DT / o.VolumeZ[q];
                            o.TI[qm] = o.TI[qm] - AdvFluxX * DT / o.VolumeZ[qm];
                        }
                    }
        }
    } else if (o.P.ImpExp_AdvXX == ImplicitScheme) {
        // AD:4483-4509: only D_flux / E_flux enter the tridiagonal system ("C and G not computed")
#pragma omp parallel for schedule(dynamic, CH) num_threads(o.nthreads)
        for (int k = W.KLB; k <= W.KUB; ++k)
            for (int j = W.JLB; j <= W.JUB; ++j)
                for (int i = W.ILB; i <= W.IUB; ++i) {
                    long q = o.i3(i, j, k);
                    if (o.ComputeFacesU3D[q] == 1) {
                        long qm = o.i3(i, j - 1, k);
                        double DT2 = DT / o.VolumeZ[q], DT1 = DT / o.VolumeZ[qm];
                        o.D[q]  = o.D[q]  - o.XD[q] * DT2;
                        o.E[q]  = o.E[q]  - o.XE[q] * DT2;
                        o.E[qm] = o.E[qm] + o.XD[q] * DT1;
                        o.F[qm] = o.F[qm] + o.XE[q] * DT1;
                    }
                }
    } else {
        o.err = "sub. HorizontalAdvectionXX - ModuleAdvectionDiffusion - ERR01";
        return ORACLE_ERR_ARG;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// HorizontalAdvectionYY (AD:4739-4910) + _Explicit (AD:4914-4953)
// ---------------------------------------------------------------------------------------
int HorizontalAdvectionYY(Oracle &o) {
    const auto &W = o.W;
    const double DT = o.P.DTProp;
    int bad = 0;
    if (o.st_HorAdv) {
        const int CH = chunk_of(W.JLB, W.JUB);
#pragma omp parallel num_threads(o.nthreads)
        for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH)
            for (int j = W.JLB; j <= W.JUB; ++j) {
                const long b3 = o.i3(o.S.ILB, j, k) - (long)o.S.ILB;   // address of actual i = 0
                const long b2 = o.i2(o.S.ILB, j) - (long)o.S.ILB;
                LineD du{o.DVY + b2, 1}, Prop{o.PROP + b3, 1}, Q{o.Wflux_Y + b3, 1}, V{o.VolumeZ + b3, 1};
                LineI CP{o.OpenPoints3D + b3, 1};
                if (o.Optimize) {
                    ComputeAdvection1D_TVD_SuperBee_2(W.ILB, W.IUB, DT, du, Prop, Q, V, CP,
                                                      LineW{o.YD.data() + b3, 1}, LineW{o.YE.data() + b3, 1});
                } else {
                    if (!ComputeAdvection1D_V2(W.ILB, W.IUB, DT, du, Prop, Q, V, CP, LineW{o.YC.data() + b3, 1},
                                               LineW{o.YD.data() + b3, 1}, LineW{o.YE.data() + b3, 1},
                                               LineW{o.YF.data() + b3, 1}, o.P.AdvMethodH, o.P.TVDLimitationH,
                                               o.P.VolumeRelMax, o.P.Upwind2H != 0)) {
#pragma omp atomic write
                        bad = 1;
                    }
                    if (o.P.NoAdvFlux && o.NoFluxV) {
                        // AD:4804-4813: the reference zeroes the XX coefficient arrays here (copy-paste quirk, kept)
                        for (int i = W.ILB; i <= W.IUB; ++i) {
                            long q = o.i3(i, j, k);
                            if (o.NoFluxV[q] == 1) o.XC[q] = o.XD[q] = o.XE[q] = o.XF[q] = 0.;
                        }
                    }
                }
            }
        }
    }
    if (bad) { o.err = "This method is not valid to compute Advection1D"; return ORACLE_ERR_ARG; }

    if (o.P.ImpExp_AdvYY == ExplicitScheme) {
        if (o.Optimize) {
            const int CH = chunk_of(W.KLB, W.KUB);
#pragma omp parallel for schedule(dynamic, CH) num_threads(o.nthreads)
            for (int k = W.KLB; k <= W.KUB; ++k)
                for (int j = W.JLB; j <= W.JUB; ++j)
                    for (int i = W.ILB; i <= W.IUB; ++i) {
                        long q = o.i3(i, j, k);
                        if (o.ComputeFacesV3D[q] == 1) {
                            long qm = o.i3(i - 1, j, k);
                            double AdvFluxY = (o.YD[q] * o.PROP[qm] + o.YE[q] * o.PROP[q]);
                            o.TI[q]  = o.TI[q]  + AdvFluxY * DT / o.VolumeZ[q];
                            o.TI[qm] = o.TI[qm] - AdvFluxY * DT / o.VolumeZ[qm];
                        }
                    }
        } else {
            const int CH = chunk_of(W.JLB, W.JUB);
#pragma omp parallel num_threads(o.nthreads)
            for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH) nowait
                for (int j = W.JLB; j <= W.JUB; ++j)
                    for (int i = W.ILB; i <= W.IUB; ++i) {
                        long q = o.i3(i, j, k);
                        if (o.ComputeFacesV3D[q] == 1) {
                            long qm = o.i3(i - 1, j, k);
                            double AdvFluxY = (o.YC[q] * o.PROP[o.i3(i - 2, j, k)] + o.YD[q] * o.PROP[qm] +
                                               o.YE[q] * o.PROP[q] + o.YF[q] * o.PROP[o.i3(i + 1, j, k)]);
                            o.TI[q]  = o.TI[q]  + AdvFluxY * DT / o.VolumeZ[q];
                            o.TI[qm] = o.TI[qm] - AdvFluxY * DT / o.VolumeZ[qm];
                        }
                    }
            }
        }
    } else if (o.P.ImpExp_AdvYY == ImplicitScheme) {
        // AD:4862-4893
        const int CH = chunk_of(W.JLB, W.JUB);
#pragma omp parallel num_threads(o.nthreads)
        for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH) nowait
            for (int j = W.JLB; j <= W.JUB; ++j)
                for (int i = W.ILB; i <= W.IUB; ++i) {
                    long q = o.i3(i, j, k);
                    if (o.ComputeFacesV3D[q] == 1) {
                        long qm = o.i3(i - 1, j, k);
                        double DT2 = DT / o.VolumeZ[q], DT1 = DT / o.VolumeZ[qm];
                        o.D[q]  = o.D[q]  - o.YD[q] * DT2;
                        o.E[q]  = o.E[q]  - o.YE[q] * DT2;
                        o.E[qm] = o.E[qm] + o.YD[q] * DT1;
                        o.F[qm] = o.F[qm] + o.YE[q] * DT1;
                    }
                }
        }
    } else {
        o.err = "sub. HorizontalAdvectionYY - ModuleAdvectionDiffusion - ERR01";
        return ORACLE_ERR_ARG;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// THOMAS_3D_i0_j1_NewType / _i1_j0_NewType (MF:3751-3875): the recurrence along j (dj = 1) or i (di = 1) for every
// work level; the line runs to the halo cell UB+1, whose row is the identity with TI = 0.
// ---------------------------------------------------------------------------------------
int THOMAS_3D(Oracle &o, int di, int dj) {
    const auto &Wk = o.W;
    const int CH = chunk_of(Wk.KLB, Wk.KUB);
    const int IJmin = dj ? Wk.ILB : Wk.JLB, IJmax = dj ? Wk.IUB : Wk.JUB;      // lines
    const int JImin = dj ? Wk.JLB : Wk.ILB, JImax = dj ? Wk.JUB : Wk.IUB;      // along a line
    int bad = 0;
    (void)di;
#pragma omp parallel num_threads(o.nthreads)
    {
        std::vector<double> Wv(JImax + 3), Gv(JImax + 3);
#pragma omp for schedule(dynamic, CH) nowait
        for (int K = Wk.KLB; K <= Wk.KUB; ++K)
            for (int IJ = IJmin; IJ <= IJmax; ++IJ) {
                auto at = [&](int JI) { return dj ? o.i3(IJ, JI, K) : o.i3(JI, IJ, K); };
                Wv[JImin] = -o.F[at(JImin)] / o.E[at(JImin)];
                Gv[JImin] = o.TI[at(JImin)] / o.E[at(JImin)];
                for (int JI = JImin + 1; JI <= JImax + 1; ++JI) {
                    const long q = at(JI);
                    double AUX = o.E[q] + o.D[q] * Wv[JI - 1];
                    if (std::fabs(AUX) > 0) {
                        Wv[JI] = -o.F[q] / AUX;
                        Gv[JI] = (o.TI[q] - o.D[q] * Gv[JI - 1]) / AUX;
                    } else {
#pragma omp atomic write
                        bad = 1;                                       // 'Instability in THOMAS3D' (MF:3799)
                        Wv[JI] = 0.; Gv[JI] = 0.;
                    }
                }
                o.PROP[at(JImax + 1)] = Gv[JImax + 1];
                for (int II = JImin + 1; II <= JImax + 1; ++II) {
                    const int MM = JImax + JImin + 1 - II;
                    o.PROP[at(MM)] = Wv[MM] * o.PROP[at(MM + 1)] + Gv[MM];
                }
            }
    }
    if (bad) { o.err = "Error: Instability in THOMAS3D - ModuleFunctions - ERR10"; return ORACLE_ERR_ARG; }
    return 0;
}

// ---------------------------------------------------------------------------------------
// VerticalDiffusion (AD:2708-2775) and VerticalDiffusion2 (AD:2779-2937)
// ---------------------------------------------------------------------------------------
void VerticalDiffusion(Oracle &o) {
    const auto &W = o.W;
    const int CH = chunk_of(W.JLB, W.JUB);
    const double DT = o.P.DTProp, th = o.P.ImpExp_DifV;
#pragma omp parallel num_threads(o.nthreads)
    for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                long q = o.i3(i, j, k);
                if (o.ComputeFacesW3D[q] == 1 && !o.SmallDepthCell(i, j)) {
                    long qm = o.i3(i, j, k - 1);
                    double AuxK = o.DifZ[q] * o.DUX[o.i2(i, j)] * o.DVY[o.i2(i, j)] / o.DZZ[qm];
                    if (o.P.NoDifFlux && o.NoFluxW && o.NoFluxW[q] == 1) AuxK = 0.;
                    double Aux1 = AuxK * DT / o.VolumeZ[qm];
                    double Aux2 = AuxK * DT / o.VolumeZ[q];
                    o.E[qm]  = o.E[qm]  + Aux1 * th;
                    o.F[qm]  = o.F[qm]  - Aux1 * th;
                    o.TI[qm] = o.TI[qm] + Aux1 * (o.PROP[q] - o.PROP[qm]) * (1. - th);
                    o.D[q]   = o.D[q]   - Aux2 * th;
                    o.E[q]   = o.E[q]   + Aux2 * th;
                    o.TI[q]  = o.TI[q]  - Aux2 * (o.PROP[q] - o.PROP[qm]) * (1. - th);
                }
            }
    }
}
void VerticalDiffusion2(Oracle &o) {
    const auto &W = o.W;
    const int CH = chunk_of(W.JLB, W.JUB);
    const double DT = o.P.DTProp;
    const bool implicit = o.P.ImpExp_DifV > 0.0;
    if (o.opt.Docycle_method == 1) {
#pragma omp parallel for schedule(dynamic, CH) num_threads(o.nthreads)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                if (o.ComputeFacesW3D[o.i3(i, j, W.KUB)] == 1 && !o.SmallDepthCell(i, j)) {
                    int Kbottom = o.KFloorZ[o.i2(i, j)];
                    double VolumeBottomCell = o.VolumeZ[o.i3(i, j, Kbottom)];
                    double PropBottomCell = o.PROP[o.i3(i, j, Kbottom)];
                    for (int k = Kbottom + 1; k <= W.KUB; ++k) {
                        long q = o.i3(i, j, k), qm = o.i3(i, j, k - 1);
                        double AuxK = o.DVC[q] * DT;
                        double Aux1 = AuxK / VolumeBottomCell;
                        double Aux2 = AuxK / o.VolumeZ[q];
                        VolumeBottomCell = o.VolumeZ[q];
                        if (implicit) {
                            o.D[q]  = o.D[q]  - Aux2;
                            o.E[q]  = o.E[q]  + Aux2;
                            o.E[qm] = o.E[qm] + Aux1;
                            o.F[qm] = o.F[qm] - Aux1;
                        } else {
                            double Gradient = o.PROP[q] - PropBottomCell;
                            o.TI[q]  = o.TI[q]  - Aux2 * Gradient;
                            o.TI[qm] = o.TI[qm] + Aux1 * Gradient;
                            PropBottomCell = o.PROP[q];
                        }
                    }
                }
            }
    } else {
#pragma omp parallel num_threads(o.nthreads)
        for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH)
            for (int j = W.JLB; j <= W.JUB; ++j)
                for (int i = W.ILB; i <= W.IUB; ++i) {
                    long q = o.i3(i, j, k);
                    if (o.ComputeFacesW3D[q] == 1 && !o.SmallDepthCell(i, j)) {
                        long qm = o.i3(i, j, k - 1);
                        double AuxK = o.DVC[q] * DT;
                        double Aux1 = AuxK / o.VolumeZ[qm];
                        double Aux2 = AuxK / o.VolumeZ[q];
                        if (implicit) {
                            o.D[q]  = o.D[q]  - Aux2;
                            o.E[q]  = o.E[q]  + Aux2;
                            o.E[qm] = o.E[qm] + Aux1;
                            o.F[qm] = o.F[qm] - Aux1;
                        } else {
                            o.TI[q]  = o.TI[q]  - Aux2 * (o.PROP[q] - o.PROP[qm]);
                            o.TI[qm] = o.TI[qm] + Aux1 * (o.PROP[q] - o.PROP[qm]);
                        }
                    }
                }
        }
    }
}

// ---------------------------------------------------------------------------------------
// VerticalAdvection (AD:2941-3144) + VerticalAdvection_ExplicitScheme (AD:3148-3207)
// ---------------------------------------------------------------------------------------
int VerticalAdvection(Oracle &o) {
    const auto &W = o.W;
    const int CH = chunk_of(W.JLB, W.JUB);
    const double DT = o.P.DTProp;
    const long sk = o.ld * o.nj;
    int bad = 0;
    if (o.st_VertAdv) {
#pragma omp parallel for schedule(dynamic, CH) num_threads(o.nthreads)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                if (o.OpenPoints3D[o.i3(i, j, W.KUB)] == 1) {
                    const long b3 = o.i3(i, j, o.S.KLB) - (long)o.S.KLB * sk;   // address of actual k = 0
                    LineD du{o.DWZ + b3, sk}, Prop{o.PROP + b3, sk}, Q{o.Wflux_Z + b3, sk}, V{o.VolumeZ + b3, sk};
                    LineI CP{o.OpenPoints3D + b3, sk};
                    if (o.Optimize) {
                        ComputeAdvection1D_TVD_SuperBee_2(W.KLB, W.KUB, DT, du, Prop, Q, V, CP,
                                                          LineW{o.VD.data() + b3, sk}, LineW{o.VE.data() + b3, sk});
                    } else {
                        if (!ComputeAdvection1D_V2(W.KLB, W.KUB, DT, du, Prop, Q, V, CP, LineW{o.VC.data() + b3, sk},
                                                   LineW{o.VD.data() + b3, sk}, LineW{o.VE.data() + b3, sk},
                                                   LineW{o.VF.data() + b3, sk}, o.P.AdvMethodV, o.P.TVDLimitationV,
                                                   o.P.VolumeRelMax, o.P.Upwind2V != 0)) {
#pragma omp atomic write
                            bad = 1;
                        }
                        if (o.P.NoAdvFlux && o.NoFluxW) {
                            for (int k = o.KFloorZ[o.i2(i, j)]; k <= W.KUB; ++k) {
                                long q = o.i3(i, j, k);
                                if (o.NoFluxW[q] == 1) o.VC[q] = o.VD[q] = o.VE[q] = o.VF[q] = 0.;
                            }
                        }
                    }
                }
            }
    }
    if (bad) { o.err = "This method is not valid to compute Advection1D"; return ORACLE_ERR_ARG; }

    if (o.P.ImpExp_AdvV == ExplicitScheme) {
        if (o.Optimize) {
            // VerticalAdvection_ExplicitScheme (AD:3148-3207)
            if (o.opt.Docycle_method == 1) {
#pragma omp parallel for schedule(dynamic, CH) num_threads(o.nthreads)
                for (int j = W.JLB; j <= W.JUB; ++j)
                    for (int i = W.ILB; i <= W.IUB; ++i) {
                        if (o.ComputeFacesW3D[o.i3(i, j, W.KUB)] == 1) {
                            int Kbottom = o.KFloorZ[o.i2(i, j)];
                            double Volume_BottomCell = o.VolumeZ[o.i3(i, j, Kbottom)];
                            double PROP_BottomCell = o.PROP[o.i3(i, j, Kbottom)];
                            for (int k = Kbottom + 1; k <= W.KUB; ++k) {
                                long q = o.i3(i, j, k), qm = o.i3(i, j, k - 1);
                                double AdvFluxZ = (o.VD[q] * PROP_BottomCell + o.VE[q] * o.PROP[q]) * DT;
                                PROP_BottomCell = o.PROP[q];
                                o.TI[qm] = o.TI[qm] - AdvFluxZ / Volume_BottomCell;
                                o.TI[q]  = o.TI[q]  + AdvFluxZ / o.VolumeZ[q];
                                Volume_BottomCell = o.VolumeZ[q];
                            }
                        }
                    }
            } else {
#pragma omp parallel num_threads(o.nthreads)
                for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH)
                    for (int j = W.JLB; j <= W.JUB; ++j)
                        for (int i = W.ILB; i <= W.IUB; ++i) {
                            long q = o.i3(i, j, k);
                            if (o.ComputeFacesW3D[q] == 1) {
                                long qm = o.i3(i, j, k - 1);
                                double AdvFluxZ = (o.VD[q] * o.PROP[qm] + o.VE[q] * o.PROP[q]) * DT;
                                o.TI[qm] = o.TI[qm] - AdvFluxZ / o.VolumeZ[qm];
                                o.TI[q]  = o.TI[q]  + AdvFluxZ / o.VolumeZ[q];
                            }
                        }
                }
            }
        } else {
#pragma omp parallel num_threads(o.nthreads)
            for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH)
                for (int j = W.JLB; j <= W.JUB; ++j)
                    for (int i = W.ILB; i <= W.IUB; ++i) {
                        long q = o.i3(i, j, k);
                        if (o.ComputeFacesW3D[q] == 1) {
                            long qm = o.i3(i, j, k - 1);
                            double AdvFluxZ = (o.VC[q] * o.PROP[o.i3(i, j, k - 2)] + o.VD[q] * o.PROP[qm] +
                                               o.VE[q] * o.PROP[q] + o.VF[q] * o.PROP[o.i3(i, j, k + 1)]);
                            o.TI[qm] = o.TI[qm] - AdvFluxZ * DT / o.VolumeZ[qm];
                            o.TI[q]  = o.TI[q]  + AdvFluxZ * DT / o.VolumeZ[q];
                        }
                    }
            }
        }
    } else if (o.P.ImpExp_AdvV == ImplicitScheme) {
        if (o.opt.Docycle_method == 1) {
#pragma omp parallel for schedule(dynamic, CH) num_threads(o.nthreads)
            for (int j = W.JLB; j <= W.JUB; ++j)
                for (int i = W.ILB; i <= W.IUB; ++i) {
                    if (o.ComputeFacesW3D[o.i3(i, j, W.KUB)] == 1) {
                        int Kbottom = o.KFloorZ[o.i2(i, j)];
                        double Volume_BottomCell = o.VolumeZ[o.i3(i, j, Kbottom)];
                        for (int k = Kbottom + 1; k <= W.KUB; ++k) {
                            long q = o.i3(i, j, k), qm = o.i3(i, j, k - 1);
                            double DT1 = DT / Volume_BottomCell;
                            double DT2 = DT / o.VolumeZ[q];
                            Volume_BottomCell = o.VolumeZ[q];
                            o.D[q]  = o.D[q]  - o.VD[q] * DT2;
                            o.E[q]  = o.E[q]  - o.VE[q] * DT2;
                            o.E[qm] = o.E[qm] + o.VD[q] * DT1;
                            o.F[qm] = o.F[qm] + o.VE[q] * DT1;
                        }
                    }
                }
        } else {
#pragma omp parallel num_threads(o.nthreads)
            for (int k = W.KLB; k <= W.KUB; ++k) {
#pragma omp for schedule(dynamic, CH)
                for (int j = W.JLB; j <= W.JUB; ++j)
                    for (int i = W.ILB; i <= W.IUB; ++i) {
                        long q = o.i3(i, j, k);
                        if (o.ComputeFacesW3D[q] == 1) {
                            long qm = o.i3(i, j, k - 1);
                            double DT1 = DT / o.VolumeZ[qm];
                            double DT2 = DT / o.VolumeZ[q];
                            o.D[q]  = o.D[q]  - o.VD[q] * DT2;
                            o.E[q]  = o.E[q]  - o.VE[q] * DT2;
                            o.E[qm] = o.E[qm] + o.VD[q] * DT1;
                            o.F[qm] = o.F[qm] + o.VE[q] * DT1;
                        }
                    }
            }
        }
    } else {
        o.err = "sub. VerticalAdvection - ModuleAdvectionDiffusion - ERR01";
        return ORACLE_ERR_ARG;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// NullGradProp (AD:2071-2110)
// ---------------------------------------------------------------------------------------
inline bool NullGradProp(const Oracle &o, double &BoundaryProp, int i, int j, int k) {
    const int cVn = o.ComputeFacesV3D[o.i3(i + 1, j, k)], cVs = o.ComputeFacesV3D[o.i3(i, j, k)];
    const int cUe = o.ComputeFacesU3D[o.i3(i, j + 1, k)], cUw = o.ComputeFacesU3D[o.i3(i, j, k)];
    const int Aux = cVn + cVs + cUe + cUw;
    if (Aux > 0) {
        BoundaryProp = (o.PROP[o.i3(i + 1, j, k)] * cVn + o.PROP[o.i3(i - 1, j, k)] * cVs +
                        o.PROP[o.i3(i, j + 1, k)] * cUe + o.PROP[o.i3(i, j - 1, k)] * cUw) / (double)Aux;
        return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------
// FluxAtOpenBoundary (AD:5676-5742)
// ---------------------------------------------------------------------------------------
void FluxAtOpenBoundary(Oracle &o) {
    const auto &W = o.W;
    const int CH = chunk_of(W.JLB, W.JUB);
    const double DT = o.P.DTProp;
#pragma omp parallel for schedule(dynamic, CH) num_threads(o.nthreads)
    for (int j = W.JLB; j <= W.JUB; ++j)
        for (int i = W.ILB; i <= W.IUB; ++i) {
            if (o.BoundaryPoints2D[o.i2(i, j)] == 1) {
                int KLB = std::abs(o.KFloorZ[o.i2(i, j)]);
                for (int k = KLB; k <= W.KUB; ++k) {
                    long q = o.i3(i, j, k);
                    o.QB[q] = o.Wflux_X[q] * o.ComputeFacesU3D[q]
                            - o.Wflux_X[o.i3(i, j + 1, k)] * o.ComputeFacesU3D[o.i3(i, j + 1, k)]
                            + o.Wflux_Y[q] * o.ComputeFacesV3D[q]
                            - o.Wflux_Y[o.i3(i + 1, j, k)] * o.ComputeFacesV3D[o.i3(i + 1, j, k)]
                            + o.Wflux_Z[q] * o.ComputeFacesW3D[q]
                            - o.Wflux_Z[o.i3(i, j, k + 1)] * o.ComputeFacesW3D[o.i3(i, j, k + 1)]
                            - (o.VolumeZ[q] - o.VolumeZOld[q]) / DT;
                }
            }
        }
}

// ---------------------------------------------------------------------------------------
// OpenBoundaryCondition (AD:5369-5672) -- serial in the reference (OpenMP commented out)
// ---------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------
// OrlanskiCelerity2D (MF:4129-4490) as called from AD:5549-5566: implicit, oblique radiation, celerity from the
// fields, reference field present, default relaxation times, FlowVelX given; returns NewValue.
// Index reads past the allocated upper bound (the corner i = IUB, j = JLB passes an interior cell as "exterior" and
// then reads IUB+2) follow the memory order of the reference's unpadded (0:IUB+1, 0:JUB+1) array: element IUB+2 of a
// row is element 0 of the next one.
// ---------------------------------------------------------------------------------------
double OrlanskiCelerity2D(Oracle &o, double *NewField, double *OldField, int IMin, int IMax, int JMin, int JMax, int di,
                          int dj, int i, int j, int k, double LimitMax, bool EastNorthBoundary, double DT, double FlowVelX) {
    auto at = [&](int ii, int jj) {
        if (ii > o.S.IUB) { ii -= (o.S.IUB - o.S.ILB + 1); jj += 1; }          // unpadded memory order
        if (ii < o.S.ILB) { ii += (o.S.IUB - o.S.ILB + 1); jj -= 1; }
        return o.i3(ii, jj, k);
    };
    const int *ComputePoints = o.OpenPoints3D;
    bool NoSouthWest = false, NoNorthEast = false;
    const int sg = EastNorthBoundary ? -1 : 1;
    const double Interior1New = NewField[at(i + sg * di, j + sg * dj)], Interior1Old = OldField[at(i + sg * di, j + sg * dj)];
    const double Interior2New = NewField[at(i + sg * 2 * di, j + sg * 2 * dj)];
    const double Interior3New = NewField[at(i + sg * 3 * di, j + sg * 3 * dj)];
    const double Interior6Old = OldField[at(i + sg * di + dj, j + sg * dj + di)];
    const double Interior7Old = OldField[at(i + sg * di - dj, j + sg * dj - di)];
    const int bi = i + sg * di, bj = j + sg * dj;
    if ((bi + 3 * dj) > IMax || (bj + 3 * di) > JMax) NoNorthEast = true;
    else if (ComputePoints[at(bi + dj, bj + di)] * ComputePoints[at(bi + 2 * dj, bj + 2 * di)] *
             ComputePoints[at(bi + 3 * dj, bj + 3 * di)] == 0) NoNorthEast = true;
    if ((bi - 3 * dj) < IMin || (bj - 3 * di) < JMin) NoSouthWest = true;
    else if (ComputePoints[at(bi - dj, bj - di)] * ComputePoints[at(bi - 2 * dj, bj - 2 * di)] *
             ComputePoints[at(bi - 3 * dj, bj - 3 * di)] == 0) NoSouthWest = true;
    double Boundary3Old = 0., Boundary32Old = 0., Boundary33Old = 0., Boundary4Old = 0., Boundary42Old = 0., Boundary43Old = 0.;
    if (!NoNorthEast) {
        Boundary3Old = OldField[at(i + dj, j + di)]; Boundary32Old = OldField[at(i + 2 * dj, j + 2 * di)];
        Boundary33Old = OldField[at(i + 3 * dj, j + 3 * di)];
    }
    if (!NoSouthWest) {
        Boundary4Old = OldField[at(i - dj, j - di)]; Boundary42Old = OldField[at(i - 2 * dj, j - 2 * di)];
        Boundary43Old = OldField[at(i - 3 * dj, j - 3 * di)];
    }
    const double Boundary5Old = OldField[at(i, j)];
    double WaveCelerityX_, WaveCelerityY_;
    {   // .not. ConstantCelerity_
        const double TimeVariability = Interior1New - Interior1Old;
        const double SpaceVariabilityX = Interior1New - Interior2New;
        const double AuxIntY = Interior6Old - Interior7Old;
        double SpaceVariabilityY;
        if ((AuxIntY * TimeVariability) > 0) { SpaceVariabilityY = Interior1Old - Interior7Old; if (NoSouthWest) SpaceVariabilityY = 0.; }
        else { SpaceVariabilityY = Interior6Old - Interior1Old; if (NoNorthEast) SpaceVariabilityY = 0.; }
        const double SpaceSquare = SpaceVariabilityX * SpaceVariabilityX + SpaceVariabilityY * SpaceVariabilityY;
        if (SpaceSquare > 0.) {
            WaveCelerityX_ = -(TimeVariability * SpaceVariabilityX) / SpaceSquare;
            WaveCelerityY_ = -(TimeVariability * SpaceVariabilityY) / SpaceSquare;
            if (std::fabs(WaveCelerityX_) > LimitMax) WaveCelerityX_ = LimitMax * WaveCelerityX_ / std::fabs(WaveCelerityX_);
            if (std::fabs(WaveCelerityY_) > LimitMax) WaveCelerityY_ = LimitMax * WaveCelerityY_ / std::fabs(WaveCelerityY_);
        } else {
            WaveCelerityX_ = 0.; WaveCelerityY_ = 0.;
        }
    }
    const double ReferenceValue = o.ReferenceProp[at(i, j)];
    const double TrelaxOut_ = 86400 * 300, TrelaxIn_ = 86400 * 300;
    WaveCelerityX_ = WaveCelerityX_ + FlowVelX;
    double AdjacentPropX, Trelax;
    if (WaveCelerityX_ > 0) {
        WaveCelerityX_ = 4 * WaveCelerityX_;
        AdjacentPropX = 0.0546875 * Interior3New - 0.2578125 * Interior2New + 0.6015625 * Interior1New;
        Trelax = TrelaxOut_;
    } else {
        AdjacentPropX = ReferenceValue;
        WaveCelerityX_ = 0.; WaveCelerityY_ = 0.;
        Trelax = TrelaxIn_;
    }
    double AuxInt = WaveCelerityX_ * AdjacentPropX;
    if (WaveCelerityY_ >= 0 && !NoSouthWest) {
        const double AdjacentPropY = 0.0546875 * Boundary43Old - 0.2578125 * Boundary42Old + 0.6015625 * Boundary4Old + 0.6015625 * Boundary5Old;
        AuxInt = AuxInt - 4 * WaveCelerityY_ * (Boundary5Old - AdjacentPropY);
    } else if (WaveCelerityY_ < 0 && !NoNorthEast) {
        const double AdjacentPropY = 0.0546875 * Boundary33Old - 0.2578125 * Boundary32Old + 0.6015625 * Boundary3Old + 0.6015625 * Boundary5Old;
        AuxInt = AuxInt - 4 * WaveCelerityY_ * (AdjacentPropY - Boundary5Old);
    }
    OldField[at(i, j)] = NewField[at(i, j)];
    double AuxBound = 1;
    AuxBound = AuxBound + WaveCelerityX_ * (1 - 0.6015625) + DT / Trelax;
    NewField[at(i, j)] = (OldField[at(i, j)] + AuxInt + ReferenceValue * DT / Trelax) / AuxBound;
    return NewField[at(i, j)];
}

int OpenBoundaryCondition(Oracle &o) {
    const auto &W = o.W;
    const int BoundaryCondition = o.P.BoundaryCondition;
    const double DTPropDouble = o.P.DTProp;
    const double TdecAux = 1.0 / (1.0 + o.P.DecayTime / o.P.DTProp);

    if (BoundaryCondition == MOHID_BC_MassConservation || BoundaryCondition == MOHID_BC_Orlanski ||
        BoundaryCondition == MOHID_BC_MassConservNullGrad)
        FluxAtOpenBoundary(o);
    // AD:5428-5429: the "old" field of the radiation condition is overwritten with the current one before it is used,
    // so the time variability seen by OrlanskiCelerity2D is zero and the celerity is the boundary flow alone (kept)
    std::vector<double> PROPOld;
    if (BoundaryCondition == MOHID_BC_Orlanski) PROPOld.assign(o.PROP, o.PROP + o.n3);

    for (int j = W.JLB; j <= W.JUB; ++j)
        for (int i = W.ILB; i <= W.IUB; ++i) {
            if (o.BoundaryPoints2D[o.i2(i, j)] != 1) continue;
            int KLB = std::abs(o.KFloorZ[o.i2(i, j)]);
            for (int k = KLB; k <= W.KUB; ++k) {
                long q = o.i3(i, j, k);
                if (o.OpenPoints3D[q] != 1) continue;
                double DT_V = DTPropDouble / o.VolumeZ[q];
                double ExteriorProp = 0., InteriorProp;

                if (BoundaryCondition == MOHID_BC_ImposedValue || BoundaryCondition == MOHID_BC_SubModel) {
                    double A1 = o.OpenPoints3D[o.i3(i + 1, j, k)] * (1 - o.BoundaryPoints2D[o.i2(i + 1, j)]);
                    double A2 = o.OpenPoints3D[o.i3(i - 1, j, k)] * (1 - o.BoundaryPoints2D[o.i2(i - 1, j)]);
                    double A3 = o.OpenPoints3D[o.i3(i, j + 1, k)] * (1 - o.BoundaryPoints2D[o.i2(i, j + 1)]);
                    double A4 = o.OpenPoints3D[o.i3(i, j - 1, k)] * (1 - o.BoundaryPoints2D[o.i2(i, j - 1)]);
                    double P1 = o.PROP[o.i3(i + 1, j, k)], P2 = o.PROP[o.i3(i - 1, j, k)];
                    double P3 = o.PROP[o.i3(i, j + 1, k)], P4 = o.PROP[o.i3(i, j - 1, k)];
                    double Atotal = A1 + A2 + A3 + A4;
                    if (Atotal > 0) {
                        InteriorProp = (P1 * A1 + P2 * A2 + P3 * A3 + P4 * A4) / Atotal;
                        ExteriorProp = InteriorProp * (1.0 - TdecAux) + o.ReferenceProp[q] * TdecAux;
                    } else {
                        ExteriorProp = o.ReferenceProp[q];
                    }
                }

                if (BoundaryCondition == MOHID_BC_Orlanski) {
                    // AD:5504-5570
                    const double VelBound = o.QB[q] * DT_V;
                    int di, dj, iext, jext;
                    bool EastNorthBoundary;
                    if (i == W.ILB || i == W.IUB) { di = 1; dj = 0; }
                    else if (j == W.JLB || j == W.JUB) { di = 0; dj = 1; }
                    else { o.err = "Orlanski Advection 2"; return ORACLE_ERR_ARG; }
                    if (i == W.ILB || j == W.JLB) { EastNorthBoundary = false; iext = i - di; jext = j - dj; }
                    else if (i == W.IUB || j == W.JUB) { EastNorthBoundary = true; iext = i + di; jext = j + dj; }
                    else { o.err = "Orlanski Advection 1"; return ORACLE_ERR_ARG; }
                    const double LimitMax = 10. * o.P.DTProp / (o.DUX[o.i2(i, j)] + o.DVY[o.i2(i, j)]) * 2;
                    ExteriorProp = OrlanskiCelerity2D(o, o.PROP, PROPOld.data(), W.ILB, W.IUB, W.JLB, W.JUB, di, dj, iext, jext,
                                                      k, LimitMax, EastNorthBoundary, o.P.DTProp, VelBound);
                }
                if (BoundaryCondition == MOHID_BC_MassConservation || BoundaryCondition == MOHID_BC_Orlanski ||
                    BoundaryCondition == MOHID_BC_MassConservNullGrad) {
                    if (BoundaryCondition == MOHID_BC_MassConservation) {
                        InteriorProp = o.PROP[q];
                        ExteriorProp = InteriorProp * (1.0 - TdecAux) + o.ReferenceProp[q] * TdecAux;
                    }
                    if (o.QB[q] < 0.0) {            // water is flowing in
                        if (BoundaryCondition != MOHID_BC_MassConservNullGrad) {
                            o.TI[q] = o.TI[q] - o.QB[q] * ExteriorProp * DT_V;
                        } else {
                            double BoundaryProp;
                            if (NullGradProp(o, BoundaryProp, i, j, k)) o.TI[q] = BoundaryProp;
                            else o.TI[q] = o.PROP[q];
                            o.E[q] = 1.; o.D[q] = 0.; o.F[q] = 0.;
                        }
                    } else {
                        o.E[q] = o.E[q] + o.QB[q] * DT_V;
                    }
                } else if (BoundaryCondition == MOHID_BC_ImposedValue || BoundaryCondition == MOHID_BC_SubModel) {
                    o.TI[q] = ExteriorProp;
                    o.D[q] = 0.0; o.E[q] = 1.0; o.F[q] = 0.0;
                } else if (BoundaryCondition == MOHID_BC_NullGradient ||
                           BoundaryCondition == MOHID_BC_CyclicBoundary) {
                    o.TI[q] = o.PROP[q];
                    o.D[q] = 0.0; o.E[q] = 1.0; o.F[q] = 0.0;
                } else {
                    o.err = "sub. OpenBoundaryCondition - ModuleAdvectionDiffusion - ERR02";
                    return ORACLE_ERR_ARG;
                }
            }
        }
    return 0;
}

// ---------------------------------------------------------------------------------------
// THOMASZ_NewType2 (MF:4026-4123)
// ---------------------------------------------------------------------------------------
void THOMASZ_NewType2(Oracle &o) {
    const int ILB = o.W.ILB, IUB = o.W.IUB, JLB = o.W.JLB, JUB = o.W.JUB, KLB = o.W.KLB, KUB = o.W.KUB;
    const int CH = chunk_of(JLB, JUB);
    const int off = std::min(std::min(o.S.ILB, o.S.JLB), o.S.KLB);   // VEC allocated (IJKLB:IJKUB), AD:626-637
    long zp = 0;
#pragma omp parallel num_threads(o.nthreads) reduction(+ : zp)
    {
        int TID = 0;
#ifdef _OPENMP
        TID = omp_get_thread_num();
#endif
        double *Wv = o.VEC_W[TID].data() - off, *Gv = o.VEC_G[TID].data() - off;
#pragma omp for schedule(dynamic, CH) nowait
        for (int J = JLB; J <= JUB; ++J)
            for (int I = ILB; I <= IUB; ++I) {
                if (o.WaterPoints3D[o.i3(I, J, KUB)] == 1) {
                    const long q1 = o.i3(I, J, 1);                   // literal index 1 (MF:4087-4088)
                    Wv[KLB] = -o.F[q1] / o.E[q1];
                    Gv[KLB] = o.TI[q1] / o.E[q1];
                    for (int K = KLB + 1; K <= KUB + 1; ++K) {
                        const long q = o.i3(I, J, K);
                        double AUX = o.E[q] + o.D[q] * Wv[K - 1];
                        if (std::fabs(AUX) > 0) {
                            Wv[K] = -o.F[q] / AUX;
                            Gv[K] = (o.TI[q] - o.D[q] * Gv[K - 1]) / AUX;
                        } else {
                            ++zp;                                     // W,G left stale (MF:4092-4098)
                        }
                    }
                    o.PROP[o.i3(I, J, KUB + 1)] = Gv[KUB + 1];
                    for (int II = KLB + 1; II <= KUB + 1; ++II) {
                        int MM = KUB + KLB + 1 - II;
                        o.PROP[o.i3(I, J, MM)] = Wv[MM] * o.PROP[o.i3(I, J, MM + 1)] + Gv[MM];
                    }
                }
            }
    }
    o.zero_pivots += zp;
}

// ---------------------------------------------------------------------------------------
// ImposeNullGradient (AD:1926-1987) -- serial in the reference
// ---------------------------------------------------------------------------------------
void ImposeNullGradient(Oracle &o) {
    const auto &W = o.W;
    for (int j = W.JLB; j <= W.JUB; ++j)
        for (int i = W.ILB; i <= W.IUB; ++i) {
            if (o.BoundaryPoints2D[o.i2(i, j)] == 1) {
                int KLB = std::abs(o.KFloorZ[o.i2(i, j)]);
                for (int k = KLB; k <= W.KUB; ++k) {
                    double BoundaryProp;
                    if (NullGradProp(o, BoundaryProp, i, j, k)) o.PROP[o.i3(i, j, k)] = BoundaryProp;
                }
            }
        }
}

// ---------------------------------------------------------------------------------------
// Prop_CyclicBoundary (AD:2121-2224)
// ---------------------------------------------------------------------------------------
void Prop_CyclicBoundary(Oracle &o) {
    const int KLB = o.W.KLB, KUB = o.W.KUB, IUB = o.W.IUB, ILB = o.W.ILB, JUB = o.W.JUB, JLB = o.W.JLB;
    for (int j = JLB; j <= JUB; ++j)
        for (int i = ILB; i <= IUB; ++i)
            if (o.BoundaryPoints2D[o.i2(i, j)] == 1)
                for (int k = KLB; k <= KUB; ++k) o.PROP[o.i3(i, j, k)] = o.ReferenceProp[o.i3(i, j, k)];
    for (int i = ILB + 1; i <= IUB - 1; ++i) {
        if (o.BoundaryPoints2D[o.i2(i, JLB)] == 1 && o.BoundaryPoints2D[o.i2(i, JUB)] == 1) {
            int kbottom = o.KFloorZ[o.i2(i, JUB - 1)];
            for (int k = kbottom; k <= KUB; ++k) o.PROP[o.i3(i, JLB, k)] = o.PROP[o.i3(i, JUB - 1, k)];
            kbottom = o.KFloorZ[o.i2(i, JLB + 1)];
            for (int k = kbottom; k <= KUB; ++k) o.PROP[o.i3(i, JUB, k)] = o.PROP[o.i3(i, JLB + 1, k)];
        }
    }
    for (int j = JLB + 1; j <= JUB - 1; ++j) {
        if (o.BoundaryPoints2D[o.i2(ILB, j)] == 1 && o.BoundaryPoints2D[o.i2(IUB, j)] == 1) {
            int kbottom = o.KFloorZ[o.i2(IUB - 1, j)];
            for (int k = kbottom; k <= KUB; ++k) o.PROP[o.i3(ILB, j, k)] = o.PROP[o.i3(IUB - 1, j, k)];
            kbottom = o.KFloorZ[o.i2(ILB + 1, j)];
            for (int k = kbottom; k <= KUB; ++k) o.PROP[o.i3(IUB, j, k)] = o.PROP[o.i3(ILB + 1, j, k)];
        }
    }
}

// ---------------------------------------------------------------------------------------
// Cell-face mass fluxes for the box budgets (AD:3356-3954).  `field` is the property the weights multiply:
// the old field for the explicit shares (called from inside the flux passes), the new one after the solve.
// ---------------------------------------------------------------------------------------
void CalcVerticalAdvFlux(Oracle &o, double Weigth) {          // AD:3356-3396 and _opt AD:3400-3463
    const auto &W = o.W;
    for (int k = W.KLB; k <= W.KUB; ++k)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                long q = o.i3(i, j, k);
                if (o.ComputeFacesW3D[q] == 1) {
                    if (o.Optimize)
                        o.AdvFluxZ[q] = o.AdvFluxZ[q] + Weigth * (o.VD[q] * o.PROP[o.i3(i, j, k - 1)] + o.VE[q] * o.PROP[q]);
                    else
                        o.AdvFluxZ[q] = o.AdvFluxZ[q] + Weigth * (o.VC[q] * o.PROP[o.i3(i, j, k - 2)] + o.VD[q] * o.PROP[o.i3(i, j, k - 1)] +
                                                                  o.VE[q] * o.PROP[q] + o.VF[q] * o.PROP[o.i3(i, j, k + 1)]);
                }
            }
}
void CalcHorizontalAdvFluxXX(Oracle &o, double Weigth) {      // AD:3467-3540
    const auto &W = o.W;
    for (int k = W.KLB; k <= W.KUB; ++k)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                long q = o.i3(i, j, k);
                if (o.ComputeFacesU3D[q] == 1) {
                    if (o.Optimize)
                        o.AdvFluxX[q] = o.AdvFluxX[q] + Weigth * (o.XD[q] * o.PROP[o.i3(i, j - 1, k)] + o.XE[q] * o.PROP[q]);
                    else
                        o.AdvFluxX[q] = o.AdvFluxX[q] + Weigth * (o.XC[q] * o.PROP[o.i3(i, j - 2, k)] + o.XD[q] * o.PROP[o.i3(i, j - 1, k)] +
                                                                  o.XE[q] * o.PROP[q] + o.XF[q] * o.PROP[o.i3(i, j + 1, k)]);
                }
            }
}
void CalcHorizontalAdvFluxYY(Oracle &o, double Weigth) {      // AD:3544-3660
    const auto &W = o.W;
    for (int k = W.KLB; k <= W.KUB; ++k)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                long q = o.i3(i, j, k);
                if (o.ComputeFacesV3D[q] == 1) {
                    if (o.Optimize)
                        o.AdvFluxY[q] = o.AdvFluxY[q] + Weigth * (o.YD[q] * o.PROP[o.i3(i - 1, j, k)] + o.YE[q] * o.PROP[q]);
                    else
                        o.AdvFluxY[q] = o.AdvFluxY[q] + Weigth * (o.YC[q] * o.PROP[o.i3(i - 2, j, k)] + o.YD[q] * o.PROP[o.i3(i - 1, j, k)] +
                                                                  o.YE[q] * o.PROP[q] + o.YF[q] * o.PROP[o.i3(i + 1, j, k)]);
                }
            }
}
void CalcVerticalDifFlux(Oracle &o, double Weigth) {          // AD:3664-3703 and CalcVerticalDifFlux2 AD:3705-3753
    const auto &W = o.W;
    for (int k = W.KLB; k <= W.KUB; ++k)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                long q = o.i3(i, j, k);
                if (o.ComputeFacesW3D[q] == 1) {
                    long qm = o.i3(i, j, k - 1);
                    if (o.Optimize)
                        o.DifFluxZ[q] = o.DifFluxZ[q] - Weigth * o.DVC[q] * (o.PROP[q] - o.PROP[qm]);
                    else
                        o.DifFluxZ[q] = o.DifFluxZ[q] - Weigth * o.DifZ[q] * o.DUX[o.i2(i, j)] * o.DVY[o.i2(i, j)] / o.DZZ[qm] *
                                                            (o.PROP[q] - o.PROP[qm]);
                }
            }
}
void CalcHorizontalDifFluxXX(Oracle &o) {                     // AD:3755-3809
    const auto &W = o.W;
    for (int k = W.KLB; k <= W.KUB; ++k)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                long q = o.i3(i, j, k);
                if (o.ComputeFacesU3D[q] == 1) {
                    long qm = o.i3(i, j - 1, k);
                    if (o.Optimize) o.DifFluxX[q] = o.DifFluxX[q] - o.DHU[q] * (o.PROP[q] - o.PROP[qm]);
                    else o.DifFluxX[q] = o.DifFluxX[q] - o.DifX[q] * o.AreaU[q] / o.DZX[o.i2(i, j - 1)] * (o.PROP[q] - o.PROP[qm]);
                }
            }
}
void CalcHorizontalDifFluxYY(Oracle &o) {                     // AD:3811-3866
    const auto &W = o.W;
    for (int k = W.KLB; k <= W.KUB; ++k)
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                long q = o.i3(i, j, k);
                if (o.ComputeFacesV3D[q] == 1) {
                    long qm = o.i3(i - 1, j, k);
                    if (o.Optimize) o.DifFluxY[q] = o.DifFluxY[q] - o.DHV[q] * (o.PROP[q] - o.PROP[qm]);
                    else o.DifFluxY[q] = o.DifFluxY[q] - o.DifY[q] * o.AreaV[q] / o.DZY[o.i2(i - 1, j)] * (o.PROP[q] - o.PROP[qm]);
                }
            }
}

// ---------------------------------------------------------------------------------------
// AdvectionDiffusionIteration (AD:1624-1922)
// ---------------------------------------------------------------------------------------
int AdvectionDiffusionIteration(Oracle &o) {
    int rc;
    if (o.Optimize) {
        if (o.FirstProperty) {
            SetMatrixValue(o, o.TI, 0.0);
            SetMatrixValue(o, o.E, 1.0);
        }
        SetMatrixValue(o, o.D, 0.0);
        SetMatrixValue(o, o.F, 0.0);
    } else {
        SetMatrixValue(o, o.D, 0.0);
        SetMatrixValue(o, o.E, 1.0);
        SetMatrixValue(o, o.F, 0.0);
        SetMatrixValue(o, o.TI, 0.0);
    }
    if (o.st_HorAdv) {
        if (o.Optimize) {
            if (o.FirstProperty) {
                SetMatrixValue(o, o.XD, 0.0); SetMatrixValue(o, o.XE, 0.0);
                SetMatrixValue(o, o.YD, 0.0); SetMatrixValue(o, o.YE, 0.0);
            }
        } else {
            SetMatrixValue(o, o.XC, 0.0); SetMatrixValue(o, o.XD, 0.0);
            SetMatrixValue(o, o.XE, 0.0); SetMatrixValue(o, o.XF, 0.0);
            SetMatrixValue(o, o.YC, 0.0); SetMatrixValue(o, o.YD, 0.0);
            SetMatrixValue(o, o.YE, 0.0); SetMatrixValue(o, o.YF, 0.0);
        }
    }

    VolumeVariation(o);
    if (o.DischON) Discharges(o);

    if (!o.opt.Vertical1D) {
        // HorizontalDiffusion (AD:5123-5152)
        if (o.Optimize) HorizontalDiffusionXX2(o); else HorizontalDiffusionXX(o);
        if (o.st_CellFluxes) CalcHorizontalDifFluxXX(o);                        // AD:5201, 5254
        if (!o.opt.XZFlow) {
            if (o.Optimize) HorizontalDiffusionYY2(o); else HorizontalDiffusionYY(o);
            if (o.st_CellFluxes) CalcHorizontalDifFluxYY(o);                    // AD:5309, 5360
        }
        // HorizontalAdvection (AD:4132-4265)
        if ((rc = HorizontalAdvectionXX(o))) return rc;
        if (o.st_CellFluxes && o.P.ImpExp_AdvXX == ExplicitScheme) CalcHorizontalAdvFluxXX(o, 1. - o.P.ImpExp_AdvXX);   // AD:4519-4525
        if (!o.opt.XZFlow) {
            if ((rc = HorizontalAdvectionYY(o))) return rc;
            if (o.st_CellFluxes && o.P.ImpExp_AdvYY == ExplicitScheme) CalcHorizontalAdvFluxYY(o, 1. - o.P.ImpExp_AdvYY);   // AD:4902-4908
        }
        // direction splitting (AD:4193-4258): solve the implicit horizontal direction, then restart the system from
        // the intermediate field for the vertical terms
        if ((o.P.ImpExp_AdvXX == ImplicitScheme || o.P.ImpExp_AdvYY == ImplicitScheme) && o.W.KUB > 1) {
            if (o.P.ImpExp_AdvXX == ImplicitScheme) { if ((rc = THOMAS_3D(o, 0, 1))) return rc; }
            else if ((rc = THOMAS_3D(o, 1, 0))) return rc;
            SetMatrixValue(o, o.D, 0.0);
            SetMatrixValue(o, o.E, 1.0);
            SetMatrixValue(o, o.F, 0.0);
            for (long q = 0; q < o.n3; ++q) o.TI[q] = o.PROP[q];          // SetMatrixValue(TICOEF3, Size, PROP), AD:4253
        }
    }

    if (o.W.KUB > 1) {
        if (o.Optimize) VerticalDiffusion2(o); else VerticalDiffusion(o);
        if (o.st_CellFluxes && o.P.ImpExp_DifV < 1.) CalcVerticalDifFlux(o, 1. - o.P.ImpExp_DifV);          // AD:2768, 2930
        if (!o.opt.Vertical1D) {
            if ((rc = VerticalAdvection(o))) return rc;
            if (o.st_CellFluxes && o.P.ImpExp_AdvV < 1.) CalcVerticalAdvFlux(o, 1. - o.P.ImpExp_AdvV);     // AD:3131-3137
        }
    }

    if (o.st_OpenBoundary) if ((rc = OpenBoundaryCondition(o))) return rc;

    SetMatrixValue(o, o.TI, null_real, o.LandPoints3D);       // AD:1753

    if (o.W.KUB == 1 && (o.P.ImpExp_AdvXX == ImplicitScheme || o.P.ImpExp_AdvYY == ImplicitScheme)) {
        // 2-D domain, horizontally implicit (AD:1758-1841, the branch without domain decomposition): the one system --
        // explicit terms, the implicit direction's D / E / F, open-boundary rows, land fill -- is solved along the lines
        if (o.P.ImpExp_AdvXX == ImplicitScheme) { if ((rc = THOMAS_3D(o, 0, 1))) return rc; }
        else if ((rc = THOMAS_3D(o, 1, 0))) return rc;
    } else {
        THOMASZ_NewType2(o);
    }

    if (o.P.BoundaryCondition == MOHID_BC_NullGradient) ImposeNullGradient(o);
    else if (o.P.BoundaryCondition == MOHID_BC_CyclicBoundary) Prop_CyclicBoundary(o);
    // implicit shares of the cell fluxes, with the new field (AD:1885-1916)
    if (o.st_CellFluxes) {
        if (o.P.ImpExp_AdvV > 0.0 && o.W.KUB > 1) CalcVerticalAdvFlux(o, o.P.ImpExp_AdvV);
        if (o.P.ImpExp_DifV > 0.0 && o.W.KUB > 1) CalcVerticalDifFlux(o, o.P.ImpExp_DifV);
        // an implicit horizontal direction: its coefficients (built from the field at time n) times the FINAL field
        // (AD:1895-1899 / 1908-1912) -- after a split step not the flux the line solve applied, which used the intermediate field
        if (o.P.ImpExp_AdvXX == ImplicitScheme) CalcHorizontalAdvFluxXX(o, o.P.ImpExp_AdvXX);
        if (o.P.ImpExp_AdvYY == ImplicitScheme) CalcHorizontalAdvFluxYY(o, o.P.ImpExp_AdvYY);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// AdvectionDiffusion (AD:1108-1509)
// ---------------------------------------------------------------------------------------
int AdvectionDiffusion(Oracle &o, double *PROP, const double *ReferenceProp, const mohid_adt_params &p,
                       bool Optimize, bool FirstProperty, double now) {
    o.Now = now;
    // AD:1229-1237
    if ((p.ImpExp_AdvXX == ImplicitScheme || p.ImpExp_AdvYY == ImplicitScheme) &&
        (p.AdvMethodH == MOHID_UpwindOrder2 || p.AdvMethodH == MOHID_UpwindOrder3)) {
        o.err = "AdvectionDiffusion - ModuleAdvectionDiffusion - ERR100"; return ORACLE_ERR_ARG;
    }
    if (p.ImpExp_AdvV == ImplicitScheme &&
        (p.AdvMethodV == MOHID_UpwindOrder2 || p.AdvMethodV == MOHID_UpwindOrder3)) {
        o.err = "AdvectionDiffusion - ModuleAdvectionDiffusion - ERR200"; return ORACLE_ERR_ARG;
    }
    o.ReferenceProp = (p.BoundaryCondition != MOHID_BC_None) ? ReferenceProp : nullptr;
    if (p.BoundaryCondition != MOHID_BC_None && !ReferenceProp) {
        // WP always passes Property%Assimilation%Field; without it State%OpenBoundary is OFF (AD:5816-5830)
        o.ReferenceProp = nullptr;
    }
    o.PROP = PROP;
    o.Optimize = Optimize;
    o.FirstProperty = FirstProperty;
    if (p.ImpExp_DifH != 0.0) { o.err = "AdvectionDiffusion - ModuleAdvectionDiffusion - ERR02"; return ORACLE_ERR_ARG; }
    if (p.ImpExp_AdvXX == ImplicitScheme && p.ImpExp_AdvYY == ImplicitScheme) {
        o.err = "AdvectionDiffusion - ModuleAdvectionDiffusion - ERR03"; return ORACLE_ERR_ARG;
    }
    if (!o.DUX || !o.Wflux_X) { o.err = "set_grid2d / set_step not called"; return MOHID_ADT_ERR_STATE; }

    // Set_Internal_State compares against the values stored by the previous call (AD:1409-1421)
    mohid_adt_params pp = p;
    int rc = 0;
    {
        // boundary-condition value seen by Set_Internal_State is the one stored at AD:1320-1324
        o.P.BoundaryCondition = p.BoundaryCondition;
        if ((rc = Set_Internal_State(o, pp))) return rc;
    }
    o.P = p;
    o.prev_DTProp = p.DTProp; o.prev_AdvMethodH = p.AdvMethodH; o.prev_TVDLimitationH = p.TVDLimitationH;
    o.prev_SchmidtCoef_V = p.SchmidtCoef_V; o.prev_SchmidtBackground_V = p.SchmidtBackground_V;
    o.prev_Schmidt_H = p.Schmidt_H;
    o.have_prev = true;

    if (o.opt.Vertical1D) o.Optimize = false;

    if (o.st_VertDif) { if (!o.Optimize || o.FirstProperty) Convert_Dif_Vertical(o); }
    if (o.st_HorDif)  { if (o.Optimize) { if (o.FirstProperty) Convert_Visc_Dif_Horizontal(o, true); }
                        else Convert_Visc_Dif_Horizontal(o, false); }
    if (o.st_VertAdv) {
        o.prev_AdvMethodV = p.AdvMethodV; o.prev_TVDLimitationV = p.TVDLimitationV;   // AD:1443-1445
        if (o.Optimize) {
            if (o.FirstProperty) { SetMatrixValue(o, o.VD, 0.0); SetMatrixValue(o, o.VE, 0.0); }
        } else {
            SetMatrixValue(o, o.VC, 0.0); SetMatrixValue(o, o.VD, 0.0);
            SetMatrixValue(o, o.VE, 0.0); SetMatrixValue(o, o.VF, 0.0);
        }
    }
    o.st_CellFluxes = p.CellFluxes != 0;                      // Set_Internal_State cd3 (AD:5809-5813)
    if (o.st_CellFluxes) {                                    // AD:1457-1470
        for (auto *v : {&o.AdvFluxX, &o.AdvFluxY, &o.AdvFluxZ, &o.DifFluxX, &o.DifFluxY, &o.DifFluxZ}) v->assign(o.n3, 0.0);
    }
    if (o.Optimize) {
        if (o.FirstTime) {
            o.DHU.assign(o.n3, 0.0); o.DHV.assign(o.n3, 0.0); o.DVC.assign(o.n3, 0.0);   // allocate (AD:1474-1477)
            o.FirstTime = false;
        }
        if (o.FirstProperty) { Compute_DifH_Constants(o); Compute_DifV_Constants(o); }
    }

    rc = AdvectionDiffusionIteration(o);
    o.PROP = nullptr; o.ReferenceProp = nullptr;      // FinishAdvectionDiffusionIt (AD:2229-2349)
    if (rc) return rc;
    o.LastCalc = o.Now;
    return 0;
}

// -------------------------------- instance registry ------------------------------------
std::mutex g_mu;
std::map<int, Oracle *> g_inst;
int g_next = 1;
std::string g_err;

Oracle *get(const int *h) {
    if (!h) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_inst.find(*h);
    return it == g_inst.end() ? nullptr : it->second;
}

}  // namespace

// =======================================================================================
// extern "C" surface (same shapes as include/mohid_adt.h so tests read alike)
// =======================================================================================
extern "C" {

int mohid_oracle_create(int *handle, const mohid_adt_size3d *size, const mohid_adt_size3d *worksize,
                        const int *ld_i, const mohid_adt_options *opt, const int *nthreads) {
    if (!handle || !size || !worksize) return ORACLE_ERR_ARG;
    auto *o = new Oracle();
    o->S = *size; o->W = *worksize;
    long ni = size->IUB - size->ILB + 1;
    o->ld = (ld_i && *ld_i > 0) ? *ld_i : ni;
    if (o->ld < ni) { delete o; return ORACLE_ERR_ARG; }
    o->nj = size->JUB - size->JLB + 1;
    o->nk = size->KUB - size->KLB + 1;
    o->n3 = o->ld * o->nj * o->nk;
    o->n2 = o->ld * o->nj;
    if (opt) o->opt = *opt; else { o->opt = mohid_adt_options{}; o->opt.Docycle_method = 1; }
    if (o->opt.Docycle_method == 0) o->opt.Docycle_method = 1;
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    if (nthreads && *nthreads > 0) nt = *nthreads;
#ifndef _OPENMP
    nt = 1;
#endif
    o->nthreads = nt;
    // AllocateVariables (AD:537-684): everything Null_real
    for (auto *v : {&o->DifX, &o->DifY, &o->DifZ, &o->D, &o->E, &o->F, &o->TI, &o->VC, &o->VD, &o->VE, &o->VF,
                    &o->XC, &o->XD, &o->XE, &o->XF, &o->YC, &o->YD, &o->YE, &o->YF, &o->QB})
        v->assign(o->n3, null_real);
    int lo = std::min(std::min(size->ILB, size->JLB), size->KLB);
    int hi = std::max(std::max(size->IUB, size->JUB), size->KUB);
    o->VEC_G.assign(nt, std::vector<double>(hi - lo + 1, null_real));
    o->VEC_W.assign(nt, std::vector<double>(hi - lo + 1, null_real));
    std::lock_guard<std::mutex> lk(g_mu);
    *handle = g_next++;
    g_inst[*handle] = o;
    return 0;
}

int mohid_oracle_destroy(int *handle) {
    if (!handle) return ORACLE_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_inst.find(*handle);
    if (it == g_inst.end()) return MOHID_ADT_ERR_HANDLE;
    delete it->second;
    g_inst.erase(it);
    *handle = 0;
    return 0;
}

int mohid_oracle_set_grid2d(const int *handle, const double *DUX, const double *DVY, const double *DZX,
                            const double *DZY, const int *KFloorZ, const int *BoundaryPoints2D) {
    Oracle *o = get(handle);
    if (!o) return MOHID_ADT_ERR_HANDLE;
    o->DUX = DUX; o->DVY = DVY; o->DZX = DZX; o->DZY = DZY; o->KFloorZ = KFloorZ; o->BoundaryPoints2D = BoundaryPoints2D;
    return 0;
}

int mohid_oracle_set_step(const int *handle, const double *Wflux_X, const double *Wflux_Y, const double *Wflux_Z,
                          const double *VolumeZOld, const double *VolumeZ, const double *Visc_H,
                          const double *Diff_V, const double *DWZ, const double *DZZ, const double *AreaU,
                          const double *AreaV, const int *OpenPoints3D, const int *LandPoints3D,
                          const int *WaterPoints3D, const int *ComputeFacesU3D, const int *ComputeFacesV3D,
                          const int *ComputeFacesW3D, const int *SmallDepths) {
    Oracle *o = get(handle);
    if (!o) return MOHID_ADT_ERR_HANDLE;
    o->Wflux_X = Wflux_X; o->Wflux_Y = Wflux_Y; o->Wflux_Z = Wflux_Z; o->VolumeZOld = VolumeZOld; o->VolumeZ = VolumeZ;
    o->Visc_H = Visc_H; o->Diff_V = Diff_V; o->DWZ = DWZ; o->DZZ = DZZ; o->AreaU = AreaU; o->AreaV = AreaV;
    o->OpenPoints3D = OpenPoints3D; o->LandPoints3D = LandPoints3D; o->WaterPoints3D = WaterPoints3D;
    o->ComputeFacesU3D = ComputeFacesU3D; o->ComputeFacesV3D = ComputeFacesV3D; o->ComputeFacesW3D = ComputeFacesW3D;
    o->SmallDepths = SmallDepths;
    return 0;
}

int mohid_oracle_set_noflux(const int *handle, const int *NoFluxU, const int *NoFluxV, const int *NoFluxW) {
    Oracle *o = get(handle);
    if (!o) return MOHID_ADT_ERR_HANDLE;
    o->NoFluxU = NoFluxU; o->NoFluxV = NoFluxV; o->NoFluxW = NoFluxW;
    return 0;
}

// Caller-side steps that ModuleWaterProperties::Advection_Diffusion_Processes runs on a property right before the
// transport call (WP:14716-14759): FreeConvection (WP:13017-13074, when Density is given), SmallDepthsMixing_Processes
// (WP:12939-13012, when WaterColumnZ is given; SmallDepthsOn(i,j) receives Me%SmallDepths%ON as 0/1) and the
// AddOffSet shift of the water points (WP:14724-14735).  Pass Offset = -x afterwards to undo the shift (WP:14833-14846).
int mohid_oracle_caller_premix(const int *handle, double *PROP, const double *Density, const double *WaterColumnZ,
                               const double *SmallDepthsLimit, int *SmallDepthsOn, const double *Offset) {
    Oracle *op = get(handle);
    if (!op) return MOHID_ADT_ERR_HANDLE;
    Oracle &o = *op;
    const auto &W = o.W;
    if (!o.OpenPoints3D || !o.VolumeZ || !o.KFloorZ) { o.err = "set_grid2d / set_step not called"; return MOHID_ADT_ERR_STATE; }
    if (Density) {
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                bool ProfileInstable = false;
                int ki = 0;
                for (int k = W.KLB; k <= W.KUB; ++k) {
                    if (o.OpenPoints3D[o.i3(i, j, k)] == 1 && o.OpenPoints3D[o.i3(i, j, k + 1)] == 1) {
                        if ((Density[o.i3(i, j, k + 1)] - Density[o.i3(i, j, k)]) > 0.) {
                            ki = k; ProfileInstable = true; break;
                        }
                    }
                }
                if (ProfileInstable) {
                    double Msum = 0., Vsum = 0.;
                    for (int k = ki; k <= W.KUB; ++k) {
                        Msum = Msum + o.VolumeZ[o.i3(i, j, k)] * PROP[o.i3(i, j, k)];
                        Vsum = Vsum + o.VolumeZ[o.i3(i, j, k)];
                    }
                    const double Cnew = Msum / Vsum;
                    for (int k = ki; k <= W.KUB; ++k) PROP[o.i3(i, j, k)] = Cnew;
                }
            }
    }
    if (WaterColumnZ) {
        for (int j = W.JLB; j <= W.JUB; ++j)
            for (int i = W.ILB; i <= W.IUB; ++i) {
                if (SmallDepthsOn) SmallDepthsOn[o.i2(i, j)] = 0;
                if (o.OpenPoints3D[o.i3(i, j, W.KUB)] == 1 && WaterColumnZ[o.i2(i, j)] < *SmallDepthsLimit) {
                    if (SmallDepthsOn) SmallDepthsOn[o.i2(i, j)] = 1;
                    double MassSum = 0., VolSum = 0.;
                    const int kbottom = o.KFloorZ[o.i2(i, j)];
                    for (int k = kbottom; k <= W.KUB; ++k) {
                        MassSum = MassSum + o.VolumeZ[o.i3(i, j, k)] * PROP[o.i3(i, j, k)];
                        VolSum = VolSum + o.VolumeZ[o.i3(i, j, k)];
                    }
                    for (int k = kbottom; k <= W.KUB; ++k) PROP[o.i3(i, j, k)] = MassSum / VolSum;
                }
            }
    }
    if (Offset && *Offset != 0.) {
        for (int k = W.KLB; k <= W.KUB; ++k)
            for (int j = W.JLB; j <= W.JUB; ++j)
                for (int i = W.ILB; i <= W.IUB; ++i)
                    if (o.WaterPoints3D[o.i3(i, j, k)] == 1) PROP[o.i3(i, j, k)] = PROP[o.i3(i, j, k)] + *Offset;
    }
    return 0;
}

// SetLimitsProperty (WP:20594-20720), the per-property body of SetLimitsConcentration(PhysicalProcesses = .true.)
// that follows the transport step (WP:12711-12714): clamp to MinValue / MaxValue and book the mass difference.
int mohid_oracle_set_limits(const int *handle, double *PROP, const int *MinOn, const double *MinValue, const int *MaxOn,
                            const double *MaxValue, double *Mass_Created, double *Mass_Destroid) {
    Oracle *op = get(handle);
    if (!op) return MOHID_ADT_ERR_HANDLE;
    Oracle &o = *op;
    const auto &W = o.W;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 0 ? !*MinOn : !*MaxOn) continue;
        double *Mass = pass == 0 ? Mass_Created : Mass_Destroid;
        const double lim = pass == 0 ? *MinValue : *MaxValue;
        auto cell = [&](long q) {
            if (pass == 0 ? (PROP[q] < lim) : (PROP[q] > lim)) {
                Mass[q] = Mass[q] + (lim - PROP[q]) * o.VolumeZ[q];
                PROP[q] = lim;
            }
        };
        if (o.opt.Docycle_method == 1) {
            for (int j = W.JLB; j <= W.JUB; ++j)
                for (int i = W.ILB; i <= W.IUB; ++i)
                    if (o.WaterPoints3D[o.i3(i, j, W.KUB)] == 1)
                        for (int k = o.KFloorZ[o.i2(i, j)]; k <= W.KUB; ++k) cell(o.i3(i, j, k));
        } else {
            for (int k = W.KLB; k <= W.KUB; ++k)
                for (int j = W.JLB; j <= W.JUB; ++j)
                    for (int i = W.ILB; i <= W.IUB; ++i)
                        if (o.WaterPoints3D[o.i3(i, j, k)] == 1) cell(o.i3(i, j, k));
        }
    }
    return 0;
}

int mohid_oracle_set_discharges(const int *handle, const int *DischNumber, const int *n_cells,
                                const double *DischFlow, const double *DischConc, const int *DischI,
                                const int *DischJ, const int *DischK, const int *DischKmin, const int *DischKmax,
                                const int *DischVert, const int *IgnoreDisch, const int *DischnCells,
                                const int *ByPass, const double *DischConcMF) {
    Oracle *o = get(handle);
    if (!o) return MOHID_ADT_ERR_HANDLE;
    int nd = *DischNumber, nc = *n_cells;
    o->DischON = true; o->DischNumber = nd;
    o->DischFlow.assign(DischFlow, DischFlow + nc); o->DischConc.assign(DischConc, DischConc + nc);
    o->DischConcMF.assign(DischConcMF, DischConcMF + nc);
    o->DischI.assign(DischI, DischI + nc); o->DischJ.assign(DischJ, DischJ + nc); o->DischK.assign(DischK, DischK + nc);
    o->DischKmin.assign(DischKmin, DischKmin + nc); o->DischKmax.assign(DischKmax, DischKmax + nc);
    o->DischVert.assign(DischVert, DischVert + nd); o->IgnoreDisch.assign(IgnoreDisch, IgnoreDisch + nd);
    o->DischnCells.assign(DischnCells, DischnCells + nd); o->ByPass.assign(ByPass, ByPass + nd);
    return 0;
}

int mohid_oracle_unset_discharges(const int *handle) {
    Oracle *o = get(handle);
    if (!o) return MOHID_ADT_ERR_HANDLE;
    o->DischON = false; o->DischNumber = 0;
    return 0;
}

// One AdvectionDiffusion call (AD:1108) for one property.
int mohid_oracle_advection_diffusion(const int *handle, double *PROP, const double *ReferenceProp,
                                     const mohid_adt_params *params, const int *Optimize,
                                     const int *FirstProperty, const double *now) {
    Oracle *o = get(handle);
    if (!o) return MOHID_ADT_ERR_HANDLE;
    return AdvectionDiffusion(*o, PROP, ReferenceProp, *params, *Optimize != 0, *FirstProperty != 0, *now);
}

// The caller loop of ModuleWaterProperties::Advection_Diffusion_Processes (WP:14580-14822):
// decides OptimizeFlag exactly as WP:14580-14598 and calls AdvectionDiffusion per property.
// force_optimize: -1 = decide like the reference, 0 / 1 = force the plain / Optimize path.
int mohid_oracle_advect_batch(const int *handle, const int *nprop, double *const *prop,
                              const double *const *reference_prop, const mohid_adt_params *params,
                              const double *now, const int *force_optimize) {
    Oracle *o = get(handle);
    if (!o) return MOHID_ADT_ERR_HANDLE;
    const int n = *nprop;
    bool OptimizeFlag = true;
    if (n > 0) {
        const double firstSchmidt = params[0].Schmidt_H;
        for (int p = 0; p < n; ++p) {
            const auto &q = params[p];
            if (q.Schmidt_H != firstSchmidt) OptimizeFlag = false;
            if (q.NoDifFlux) OptimizeFlag = false;
            if (q.NoAdvFlux) OptimizeFlag = false;
            if (q.NullDif) OptimizeFlag = false;
            if (q.AdvMethodH != MOHID_P2_TVD) OptimizeFlag = false;
            if (q.AdvMethodV != MOHID_P2_TVD) OptimizeFlag = false;
            if (q.TVDLimitationH != MOHID_SuperBee) OptimizeFlag = false;
            if (q.TVDLimitationV != MOHID_SuperBee) OptimizeFlag = false;
        }
    }
    if (n < 2) OptimizeFlag = false;
    if (force_optimize && *force_optimize >= 0) OptimizeFlag = (*force_optimize != 0);
    bool FirstWaterProperty = true;
    for (int p = 0; p < n; ++p) {
        const double *ref = reference_prop ? reference_prop[p] : nullptr;
        int rc = AdvectionDiffusion(*o, prop[p], ref, params[p], OptimizeFlag, FirstWaterProperty, *now);
        if (rc) return rc;
        FirstWaterProperty = false;
    }
    return 0;
}

int mohid_oracle_last_error(const int *handle, char *buf, const int *buflen) {
    Oracle *o = get(handle);
    const std::string &e = o ? o->err : g_err;
    if (!buf || !buflen || *buflen <= 0) return ORACLE_ERR_ARG;
    std::snprintf(buf, (size_t)*buflen, "%s", e.c_str());
    return 0;
}

// GetAdvFlux / GetDifFlux (AD:697-851): which = 0..5 -> AdvFluxX, AdvFluxY, AdvFluxZ, DifFluxX, DifFluxY, DifFluxZ
// of the LAST AdvectionDiffusion call that had CellFluxes set.
int mohid_oracle_get_cell_flux(const int *handle, const int *which, double *out) {
    Oracle *o = get(handle);
    if (!o) return MOHID_ADT_ERR_HANDLE;
    std::vector<double> *v[6] = {&o->AdvFluxX, &o->AdvFluxY, &o->AdvFluxZ, &o->DifFluxX, &o->DifFluxY, &o->DifFluxZ};
    if (*which < 0 || *which > 5 || (long)v[*which]->size() != o->n3) return ORACLE_ERR_ARG;
    std::memcpy(out, v[*which]->data(), sizeof(double) * o->n3);
    return 0;
}

int mohid_oracle_zero_pivots(const int *handle, long long *n) {
    Oracle *o = get(handle);
    if (!o) return MOHID_ADT_ERR_HANDLE;
    *n = o->zero_pivots;
    return 0;
}

int mohid_oracle_num_threads(const int *handle) {
    Oracle *o = get(handle);
    return o ? o->nthreads : 0;
}

// THOMASZ_NewType2 (MF:4026-4123) on caller-supplied coefficients: D, E, F, TI, WaterPoints3D as in the reference's
// argument bundle, RES updated in place.  Used to pin the column solve against the reference's own CUDA solver
// (Software/CudaThomas/Thomas.cu, built as oracle/_ref by oracle/Makefile).
int mohid_oracle_thomasz(const int *handle, const double *D, const double *E, const double *F, const double *TI,
                         const int *WaterPoints3D, double *RES) {
    Oracle *o = get(handle);
    if (!o) return MOHID_ADT_ERR_HANDLE;
    if (!D || !E || !F || !TI || !WaterPoints3D || !RES) return ORACLE_ERR_ARG;
    o->D.assign(D, D + o->n3); o->E.assign(E, E + o->n3); o->F.assign(F, F + o->n3); o->TI.assign(TI, TI + o->n3);
    o->WaterPoints3D = WaterPoints3D;
    o->PROP = RES;
    THOMASZ_NewType2(*o);
    o->WaterPoints3D = nullptr;
    o->PROP = nullptr;
    return 0;
}

// Face-weight function exposed for unit tests of A.5 (MF:10702-10894).
int mohid_oracle_advection_face(const double *Prop4, const double *V4, const double *du4, const double *dt,
                                const double *QFace, const double *VolumeRelMax, const int *Method,
                                const int *TVD_Limitation, const int *NearBoundary, const int *Upwind2,
                                double *CFace) {
    return ComputeAdvectionFace(Prop4, V4, du4, *dt, *QFace, *VolumeRelMax, *Method, *TVD_Limitation,
                                *NearBoundary != 0, *Upwind2 != 0, CFace) ? 0 : ORACLE_ERR_ARG;
}

}  // extern "C"
