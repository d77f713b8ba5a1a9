/*
 * mohid_adt.h -- C-ABI of the B200-native MOHID property transport step.
 *
 * This is the drop-in boundary for ModuleAdvectionDiffusion (reference paths are
 * relative to /root/reference/Software):
 *
 *   AD  = MOHIDBase2/ModuleAdvectionDiffusion.F90
 *   MF  = MOHIDBase1/ModuleFunctions.F90
 *   MGD = MOHIDBase1/ModuleGlobalData.F90
 *   MC  = MOHIDBase1/ModuleCuda.F90          (the reference's own ISO_C_BINDING precedent)
 *   WP  = MOHIDWater/ModuleWaterProperties.F90
 *
 * Conventions (same as the reference's bind(C) interfaces, MC:49-121):
 *   - every scalar is passed BY REFERENCE (Fortran default), arrays as the base address
 *     of element (ILB,JLB,KLB) = (0,0,0);
 *   - 3-D arrays are Fortran column-major (0:I+1, 0:J+1, 0:K+1), `i` contiguous, with an
 *     explicit leading dimension ld_i >= I+2 (covers _PAD_MATRICES, MF:2832-2851);
 *     2-D arrays are (0:I+1, 0:J+1) with the same ld_i;
 *   - `real` is fp64 (compile_mohid.sh:185,209), masks are int32, Fortran logicals are
 *     converted to int32 0/1 by the shim before the call;
 *   - every function returns an int status: 0 = SUCCESS_ (MGD:181); non-zero = failure,
 *     message available through mohid_adt_last_error().  The library never exit()s and has
 *     NO CPU fallback: without a usable CUDA device every compute entry point fails.
 */
#ifndef MOHID_ADT_H
#define MOHID_ADT_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (MGD:180-199 style) ------------------------------------------- */
#define MOHID_ADT_SUCCESS          0
#define MOHID_ADT_ERR_UNKNOWN    (-99)  /* UNKNOWN_ */
#define MOHID_ADT_ERR_HANDLE       9    /* IDLE_ERR_ analogue: bad / dead handle */
#define MOHID_ADT_ERR_ARG         20    /* invalid argument (reference would `stop`) */
#define MOHID_ADT_ERR_UNSUPPORTED 21    /* option exists in the reference, not on the GPU path */
#define MOHID_ADT_ERR_CUDA        30    /* CUDA runtime failure (no device, OOM, launch error) */
#define MOHID_ADT_ERR_STATE       31    /* call order violated (e.g. advect before set_step) */

/* ---- numerical option ids (MGD:1810-1812, MGD:1831-1832, AD:152-158) -------------- */
enum { MOHID_UpwindOrder1 = 1, MOHID_UpwindOrder2 = 2, MOHID_UpwindOrder3 = 3,
       MOHID_P2_TVD = 4, MOHID_CentralDif = 5, MOHID_LeapFrog = 6 };
enum { MOHID_MinMod = 1, MOHID_VanLeer = 2, MOHID_Muscl = 3, MOHID_SuperBee = 4, MOHID_PDM = 5 };
enum { MOHID_BC_None = 0,              /* ReferenceProp not associated -> State%OpenBoundary OFF (AD:5816-5830) */
       MOHID_BC_MassConservation = 1, MOHID_BC_ImposedValue = 2, MOHID_BC_NullGradient = 4,
       MOHID_BC_SubModel = 5, MOHID_BC_Orlanski = 6, MOHID_BC_MassConservNullGrad = 7,
       MOHID_BC_CyclicBoundary = 8 };
#define MOHID_NULL_REAL   (-9.9e15)    /* MGD:143 */
#define MOHID_FILL_INT    (-9999999)   /* MGD:131 FillValueInt */
#define MOHID_DischUniform 5           /* MGD:1006 */

/* T_Size3D (MGD:2041-2052, CudaWrapper/CuWrapperBinding.h:14-22) */
typedef struct mohid_adt_size3d {
    int ILB, IUB, JLB, JUB, KLB, KUB;
} mohid_adt_size3d;

/*
 * Per-property argument block = the scalar dummy arguments of AdvectionDiffusion
 * (AD:1108-1147) in the order the caller passes them (WP:14776-14822).
 * Logical dummies are int 0/1.
 */
typedef struct mohid_adt_params {
    double Schmidt_H;              /* AD:1110 */
    double SchmidtCoef_V;
    double SchmidtBackground_V;
    int    AdvMethodH;             /* MOHID_UpwindOrder1 .. MOHID_LeapFrog */
    int    TVDLimitationH;         /* MOHID_MinMod .. MOHID_PDM */
    int    AdvMethodV;
    int    TVDLimitationV;
    int    Upwind2H;               /* logical */
    int    Upwind2V;               /* logical */
    double VolumeRelMax;
    double DTProp;                 /* seconds */
    double ImpExp_AdvV;            /* exactly 0 (explicit) or 1 (implicit), AD:3124 */
    double ImpExp_DifV;            /* theta in [0,1] */
    double ImpExp_AdvXX;           /* 0 explicit / 1 implicit, AD:4525 */
    double ImpExp_AdvYY;
    double ImpExp_DifH;            /* must be 0 (AD:1340-1343) */
    int    NullDif;                /* logical */
    int    BoundaryCondition;      /* MOHID_BC_*; MOHID_BC_None when ReferenceProp is absent */
    double DecayTime;              /* seconds */
    int    NoAdvFlux;              /* logical; needs NoFluxU/V/W (mohid_adt_set_noflux) */
    int    NoDifFlux;              /* logical */
    int    CellFluxes;             /* logical: also produce the six cell-face fluxes (AD:1457-1470, 3356-3954) */
    int    Optimize;            /* the Optimize argument of the reference call (AD:1146, WP:14580-14598): 0 = inferred
                                  from the batch (all coupled properties are in it), 1 = on, 2 = off; read from the
                                  first property of the batch */
} mohid_adt_params;

/* Instance-level flags = StartAdvectionDiffusion arguments (AD:400-411) */
typedef struct mohid_adt_options {
    int Vertical1D;                /* logical */
    int XZFlow;                    /* logical */
    int Docycle_method;            /* MM:641-643, default 1 */
    int device;                    /* CUDA device ordinal, -1 = current device */
    int max_properties;            /* upper bound of nprop in one batch (device buffers sized for it) */
    int reserved[3];
} mohid_adt_options;

/* ---------------------------------------------------------------------------------- */
/* Lifetime: replaces StartAdvectionDiffusion (AD:400-533) / KillAdvectionDiffusion     */
/* (AD:5849-6010); the handle plays the role of ObjCudaID (CudaThomas/Thomas.cu:6-21).  */
int mohid_adt_create(int *handle, const mohid_adt_size3d *size, const mohid_adt_size3d *worksize,
                     const int *ld_i, const mohid_adt_options *opt);
int mohid_adt_destroy(int *handle);

/* Horizontal metrics / 2-D maps fetched by AdvectionDiffusion on every call (AD:1353-1384):
 * GetHorizontalGrid DUX,DVY,DZX,DZY (fp64 2-D), GetGeometryKFloor Z (int 2-D),
 * GetBoundaries BoundaryPoints2D (int 2-D).  Call once, or again when they change. */
int mohid_adt_set_grid2d(const int *handle, const double *DUX, const double *DVY,
                         const double *DZX, const double *DZY,
                         const int *KFloorZ, const int *BoundaryPoints2D);

/* Per-time-step shared inputs = array dummies of AdvectionDiffusion that are common to all
 * properties of a step (AD:1132-1141) plus the Geometry getters (AD:1386-1401).
 * HOST pointers; copied to the device inside the call.  SmallDepths may be NULL
 * (SmallDepthsPresent = .false., AD:1297-1302); it is int32 0/1 (0:I+1,0:J+1).
 * After a first complete call, a NULL 3-D array means "unchanged since the last step" (its device copy is kept):
 * LandPoints3D / WaterPoints3D never change, the other maps only with wetting and drying. */
int mohid_adt_set_step(const int *handle,
                       const double *Wflux_X, const double *Wflux_Y, const double *Wflux_Z,
                       const double *VolumeZOld, const double *VolumeZ,
                       const double *Visc_H, const double *Diff_V,
                       const double *DWZ, const double *DZZ,
                       const double *AreaU, const double *AreaV,
                       const int *OpenPoints3D, const int *LandPoints3D, const int *WaterPoints3D,
                       const int *ComputeFacesU3D, const int *ComputeFacesV3D,
                       const int *ComputeFacesW3D, const int *SmallDepths);

/* The same inputs for a window of columns: the arrays hold the columns j0 .. j0+ncols-1 (0-based, 0 = the halo column)
 * and are shaped (ld_i, ncols, K+2).  A host that produces or reads its fields slab by slab -- an MPI sub-domain
 * (ModuleHorizontalGrid.F90:1587-1757), a generator that cannot hold a second copy of a 100 GB case -- fills the
 * device mirrors piecewise; a NULL array is skipped.  After the last window: mohid_adt_mark_step_resident. */
int mohid_adt_set_step_columns(const int *handle, const int *j0, const int *ncols,
                               const double *Wflux_X, const double *Wflux_Y, const double *Wflux_Z,
                               const double *VolumeZOld, const double *VolumeZ,
                               const double *Visc_H, const double *Diff_V,
                               const double *DWZ, const double *DZZ, const double *AreaU, const double *AreaV,
                               const int *OpenPoints3D, const int *LandPoints3D, const int *WaterPoints3D,
                               const int *ComputeFacesU3D, const int *ComputeFacesV3D, const int *ComputeFacesW3D);

/* The optional NoFluxU / NoFluxV / NoFluxW dummies of AdvectionDiffusion (AD:1146, 1332-1334; int32 3-D):
 * faces whose advective (NoAdvFlux) or diffusive (NoDifFlux) coefficients are zeroed for the properties that
 * set those flags.  All three NULL = not present. */
int mohid_adt_set_noflux(const int *handle, const int *NoFluxU, const int *NoFluxV, const int *NoFluxW);

/* Caller-side steps of ModuleWaterProperties::Advection_Diffusion_Processes that run on each property right before
 * its AdvectionDiffusion call, done on the device-resident fields inside advect_batch /
 * advect_device:
 *   Density       /= NULL: FreeConvection (WP:13017-13074) with Me%Density%Field (fp64 3-D);
 *   WaterColumnZ  /= NULL: SmallDepthsMixing_Processes (WP:12939-13012) with the water column (fp64 2-D) and
 *                          Me%SmallDepths%Limit; the library then builds Me%SmallDepths%ON itself and uses it as the
 *                          SmallDepths dummy of AdvectionDiffusion (mohid_adt_get_small_depths returns it, int32 2-D).
 * Both NULL switches the steps off.  The arrays are copied during the call. */
int mohid_adt_set_premix(const int *handle, const double *Density, const double *WaterColumnZ,
                         const double *SmallDepthsLimit);
int mohid_adt_get_small_depths(const int *handle, int *SmallDepthsOn);
/* Property%AddOffSet / Property%OffSet (WP:14724-14746, 14833-14858): the water points of property n, of its
 * reference field and its DischConc are shifted by OffSet[n] before the step and shifted back after it. */
int mohid_adt_set_offsets(const int *handle, const int *nprop, const double *OffSet);

/* Communication stream of a column slab: with ghost > 0, pack_columns / unpack_columns run on
 * `comm_stream`, ordered after the step by an event, and the next step -- or mohid_adt_join_halo / download_props /
 * synchronize -- waits for the unpack.  The caller issues its NCCL send/recv between pack and unpack on the same
 * stream (mohid_adt_exchange_halos does all of it inside the library).  ghost = 0 switches it off.  Round 1 also
 * advanced the edge columns first to overlap the exchange with the interior; the in-place step of round 2 walks the
 * columns in one direction, and the exchange is < 3 % of a step over NVLink. */
int mohid_adt_set_overlap(const int *handle, const int *ghost, void *comm_stream);
int mohid_adt_join_halo(const int *handle);

/* SetLimitsProperty (WP:20594-20720; SetLimitsConcentration(PhysicalProcesses) at WP:12711-12714): after every step
 * property n is clamped to MinValue[n] (if MinOn[n]) and MaxValue[n] (if MaxOn[n]); the mass added / removed is
 * accumulated in Mass_created / Mass_Destroid (fp64 3-D), which mohid_adt_get_limit_mass copies out.  nprop = 0 clears. */
int mohid_adt_set_limits(const int *handle, const int *nprop, const int *MinOn, const double *MinValue,
                         const int *MaxOn, const double *MaxValue);
int mohid_adt_get_limit_mass(const int *handle, const int *prop_index, double *Mass_Created, double *Mass_Destroid);

/* SetDischarges / UnSetDischarges (AD:978-1095), same arguments as the reference plus the position
 * `prop_index` (0-based) of the property in the next advect batch: the caller invokes SetDischarges
 * once per property with that property's DischConc / DischConcMF (WP:14761-14773); the discharge
 * geometry (flows, cells, vertical distribution, Ignore, nCells, ByPass) is the caller's single
 * Me%Discharge and must be the same in every call of a step.  n_cells = number of entries of the
 * per-cell arrays (cells of ignored discharges are not listed, AD:4039-4041).  Logical arrays
 * (IgnoreDisch, ByPass) are int32 0/1.  unset clears the discharges of all properties. */
int mohid_adt_set_discharges(const int *handle, const int *prop_index, const int *DischNumber, const int *n_cells,
                             const double *DischFlow, const double *DischConc,
                             const int *DischI, const int *DischJ, const int *DischK,
                             const int *DischKmin, const int *DischKmax,
                             const int *DischVert, const int *IgnoreDisch,
                             const int *DischnCells, const int *ByPass,
                             const double *DischConcMF);
int mohid_adt_unset_discharges(const int *handle);

/* The batched replacement of the per-property AdvectionDiffusion call loop
 * (WP:14603-15143 -> AD:1108): advances `nprop` properties one transport step.
 * prop[n] / reference_prop[n] are HOST arrays (0:I+1,0:J+1,0:K+1); prop[n] is updated
 * in place exactly like PROP (MF:4101-4105); reference_prop may be NULL or hold NULL
 * entries (ReferenceProp not associated).  All properties of one batch must share
 * DTProp, the advection methods / limiters and VolumeRelMax (they are read FromFile,
 * WP:9580-9632); Schmidt numbers, theta, BC and DecayTime are per property. */
int mohid_adt_advect_batch(const int *handle, const int *nprop,
                           double *const *prop, const double *const *reference_prop,
                           const mohid_adt_params *params);

/* GetAdvFlux / GetDifFlux (AD:697-851): the six cell-face mass fluxes of property `prop_index` of the last
 * batch, for properties whose params carried CellFluxes = 1 (box budgets, WP:14956-15032).  Arrays are
 * (0:I+1,0:J+1,0:K+1) like the properties; flux (i,j,k) belongs to the U / V / W face (i,j,k).  Any pointer
 * may be NULL. */
int mohid_adt_get_cell_fluxes(const int *handle, const int *prop_index, double *AdvFluxX, double *AdvFluxY,
                              double *AdvFluxZ, double *DifFluxX, double *DifFluxY, double *DifFluxZ);

/* Box budgets on the device instead of six full-field copies to the host: BoxDifFluxes3D (ModuleBoxDif.F90:2659-2776)
 * as ModuleWaterProperties feeds it (WP:14956-15032: MassFluxes = AdvFlux + DifFlux, mask = OpenPoints3D).
 * set_boxes: Boxes3D is Me%Boxes3D (int32 3-D, box number per cell, values <= -55 = no box), NumberOfBoxes3D its count;
 * box_fluxes: Fluxes3D is the (0:NumberOfBoxes3D, 0:NumberOfBoxes3D) matrix Me%Fluxes3D, element (OUT, IN) the mass flux
 * of property `prop_index` from box OUT into box IN through the faces that separate them (antisymmetric), for a property
 * advanced with CellFluxes = 1.  Summation order differs from the reference's serial loop: equal to rounding. */
int mohid_adt_set_boxes(const int *handle, const int *Boxes3D, const int *NumberOfBoxes3D);
int mohid_adt_box_fluxes(const int *handle, const int *prop_index, double *Fluxes3D);

/* FreeVerticalMovementIteration (MOHIDWater/ModuleFreeVerticalMovement.F90:1531-1650, reached through
 * Modify_FreeVerticalMovement :1455-1527, which ModuleWaterProperties calls per settling property): the vertical movement
 * of the device-resident property `prop_index` with its own velocity field -- BottomBoundary, VerticalFreeConvection (first
 * order upwind), land fill, THOMASZ_NewType2 for the implicit scheme -- using VolumeZ, the masks and KFloorZ of set_step /
 * set_grid2d.  Velocity is PropertyX%Velocity as Vertical_Velocity left it (fp64 3-D, value at the bottom face of cell k),
 * GridCellArea and DepositionProbability fp64 2-D (the latter may be NULL unless Deposition and not NonCohesive);
 * ImpExp_AdvV in THIS module's convention: 0 = implicit, 1 = explicit.  FreeConvFlux (fp64 3-D, host, may be NULL) receives
 * PropertyX%FreeConvFlux.  The caller's velocity array is not modified (the reference zeroes / scales its bottom value). */
int mohid_adt_free_vertical_movement(const int *handle, const int *prop_index, const double *Velocity,
                                     const double *GridCellArea, const double *DepositionProbability, const int *Deposition,
                                     const int *NonCohesive, const int *DepositionIntertidalZones, const double *ImpExp_AdvV,
                                     const double *DTProp, double *FreeConvFlux);

/* ModuleHydroIntegration (MOHIDBase1/ModuleHydroIntegration.F90) on the device: a property whose DTInterval spans n
 * hydrodynamic steps is transported with the time-mean of the horizontal water fluxes, the union of the compute faces,
 * the vertical flux that closes continuity over the interval and the volume at its start (GetHydroIntegrationWaterFluxes /
 * ComputeFaces / VolumeZOld at WP:14615-14647).  The integration arrays are the handle's own step mirrors:
 *   reinit  = ReInitalizeIntegration (:767-796) at the first hydrodynamic step of the interval (VolumeZOld = initial volume);
 *   step    = OneIntegrationStep (:843-906) once per hydrodynamic step (Discharges fp64 3-D or NULL);
 *   end     = EndIntegrationStep (:910-994) at the last one (VolumeZ, WaterPoints3D of that instant, DT = the interval):
 *             afterwards Wflux_X/Y/Z, VolumeZOld, VolumeZ, OpenPoints3D, WaterPoints3D and ComputeFacesU/V/W3D of the
 *             handle are those of the interval; mohid_adt_set_step then passes NULL for them and the other arrays as usual
 *             (or mohid_adt_mark_step_resident when they are already in place). */
int mohid_adt_hydro_integration_reinit(const int *handle, const double *VolumeZOld);
int mohid_adt_hydro_integration_step(const int *handle, const double *WaterFluxX, const double *WaterFluxY,
                                     const double *Discharges, const int *ComputeFacesU, const int *ComputeFacesV);
int mohid_adt_hydro_integration_end(const int *handle, const double *VolumeZ, const int *WaterPoints3D, const double *DT);
/* Copy of a step mirror back into a caller array of the interface shape: which = 0..10 the fp64 arrays of set_step in
 * argument order (fp64 array), 11..16 its int32 masks (int32 array) -- e.g. the integrated fluxes for an output. */
int mohid_adt_download_step_input(const int *handle, const int *which, void *array);

/* ---- device-resident variants (benchmarks, device-side callers) ------------------- */
/* Copy properties host->device / device->host without stepping. */
int mohid_adt_upload_props(const int *handle, const int *nprop, const double *const *prop,
                           const double *const *reference_prop);
int mohid_adt_download_props(const int *handle, const int *nprop, double *const *prop);
/* Advance the device-resident properties `nsteps` transport steps with the current
 * set_step inputs; no host<->device traffic.  Includes the per-step coefficient pass. */
int mohid_adt_advect_device(const int *handle, const int *nprop, const mohid_adt_params *params,
                            const int *nsteps);
/* The same for a window of columns j0 .. j0+ncols-1 (arrays shaped (ld_i, ncols, K+2)), see set_step_columns. */
int mohid_adt_upload_props_columns(const int *handle, const int *nprop, const double *const *prop,
                                   const double *const *reference_prop, const int *j0, const int *ncols);
int mohid_adt_download_props_columns(const int *handle, const int *nprop, double *const *prop, const int *j0,
                                     const int *ncols);
/* mass[n*(J+2) + j] = sum over i, k of PROP_n * VolumeZ on the water points of column j (0-based j, halo columns
 * included), summed in an order that does not depend on the decomposition: the per-column terms a box-budget or a
 * conservation check needs (ModuleWaterProperties.F90:14956-15032 sums the same products on the host), and the
 * witness bench.py prints to show that 1, 2, 4 and 8 GPUs computed the same field.  `mass` is a host array. */
int mohid_adt_column_mass(const int *handle, const int *nprop, double *mass);
/* Raw device pointer of property n and its leading dimension / plane stride in elements: element (i,j,k) is at
 * ptr[i + ld*(j + nj*k)] with nj the ALLOCATED column count returned here (>= J+2: the in-place step keeps a margin of
 * columns).  The pointer is valid until the next advect call, which moves the field inside its buffer. */
int mohid_adt_prop_device_ptr(const int *handle, const int *n, void **dptr, int *ld, int *nj, int *nk);

/* Every array pointer accepted above may be a host pointer (the Fortran caller) or, under CUDA
 * unified virtual addressing, a device pointer (device-side producers, the benchmark generator):
 * copies use cudaMemcpyDefault.
 * A device array must be complete when the call is made: the library copies on its own stream (or on the one given to
 * mohid_adt_set_stream, which orders it after that stream's earlier work). */
/* Wait for all work queued on the handle's stream. */
int mohid_adt_synchronize(const int *handle);
/* No-op since round 2 (one buffer per property); kept so that round-1 hosts still link. */
int mohid_adt_sync_prop_buffers(const int *handle, const int *nprop);
/* Device pointer of the ReferenceProp mirror of property n (allocated on first use). */
int mohid_adt_set_reference_device(const int *handle, const int *n, void **dptr);
/* Device pointer of a staged input: which = 0..10 the fp64 arrays of set_step in argument order,
 * 11..16 its int32 masks, 17 SmallDepths, 20..25 the set_grid2d arrays.  nj returns the allocated column count of
 * the 3-D arrays (see mohid_adt_prop_device_ptr), J+2 for the 2-D ones. */
int mohid_adt_step_input_device_ptr(const int *handle, const int *which, void **dptr, int *ld, int *nj, int *nk);
/* Declare the staged inputs valid after filling them through the pointers above. */
int mohid_adt_mark_step_resident(const int *handle, const int *small_depths_present);

/* ---- halo exchange support for the j-slab decomposition (replaces               ---- */
/* ---- ReceiveSendProperities3DMPIr8, HG:8479-8658)                                ---- */
/* Restrict the columns this handle advances to j_begin .. j_begin+j_count-1 (default 1..J): the
 * owned columns of a j-slab whose remaining work columns are ghost copies of the neighbours
 * (the reference instead recomputes overlapping HALOPOINTS columns, HG:1048-1105). */
int mohid_adt_set_active_columns(const int *handle, const int *j_begin, const int *j_count);
/* Pack `width` j-columns starting at j0 of all nprop device-resident properties into a
 * contiguous DEVICE buffer / scatter them back.  The buffer holds nprop * nk * width * ld doubles, laid out
 * [property][k][column][i], with ld the DEVICE leading dimension that mohid_adt_prop_device_ptr returns (rows are
 * padded to 128 bytes on the device: ld >= ld_i of mohid_adt_create) and nk = K + 2. */
int mohid_adt_pack_columns(const int *handle, const int *nprop, const int *j0, const int *width,
                           void *device_buffer);
int mohid_adt_unpack_columns(const int *handle, const int *nprop, const int *j0, const int *width,
                             const void *device_buffer);
/* Use the caller's CUDA stream (cudaStream_t passed as void*) for all work of this handle. */
int mohid_adt_set_stream(const int *handle, void *cuda_stream);

/* ---- NCCL halo exchange (one handle = one rank = one GPU) ---------------------------- */
/* Replaces the MPI halo exchange of the property fields, ReceiveSendProperitiesMPI (ModuleWaterProperties.F90:
 * 15034-15045) -> ReceiveSendProperities3DMPIr8 (ModuleHorizontalGrid.F90:8479-8658), for hosts that keep the
 * properties resident on the GPUs.  libnccl.so.2 is loaded at run time by the first of these calls.
 *   comm_get_unique_id : rank 0 obtains the 128-byte NCCL id and hands it to the other ranks with the host's own
 *                        means (MPI_Bcast in a MOHID MPI run, the torch store in bench.py);
 *   comm_init          : collective over the ranks; the handle must already be restricted to its owned columns
 *                        (mohid_adt_set_active_columns) with `ghost` (= 2, the reach of the advection stencil,
 *                        ModuleFunctions.F90:10572-10574) columns left on every interior side; overlap != 0 lets
 *                        the next steps advance the edge columns first so the exchange overlaps the interior;
 *   exchange_halos     : after a step: first / last `ghost` owned columns of properties 0..nprop-1 -> neighbours'
 *                        ghost columns (pack -> ncclSend/ncclRecv -> unpack on the handle's communication stream;
 *                        the next step, a download or mohid_adt_synchronize waits for it);
 *   comm_destroy       : collective.
 * With a communicator the steps themselves become collective when a property is advected implicitly along j
 * (ImpExp_AdvXX = 1): the lines cross the slabs and their tridiagonal recurrence passes from rank to rank inside
 * mohid_adt_advect_batch / mohid_adt_advect_device (the reference gathers such rows on one process,
 * THOMAS_DDecompHorizGrid, ModuleHorizontalGrid.F90:8245-8478); likewise with BoundaryCondition = CyclicBoundary, whose
 * j wrap (ModuleAdvectionDiffusion.F90:2167-2190) joins the first and the last rank.  Every rank must then make the same
 * calls. */
int mohid_adt_comm_get_unique_id(void *unique_id, const int *nbytes);
int mohid_adt_comm_init(const int *handle, const int *nranks, const int *rank, const void *unique_id, const int *ghost,
                        const int *overlap);
int mohid_adt_exchange_halos(const int *handle, const int *nprop);
int mohid_adt_comm_destroy(const int *handle);

/* ---- stand-alone column solve ------------------------------------------------------ */
/* THOMASZ_NewType2 (ModuleFunctions.F90:4026-4123) on caller-supplied coefficient fields: for every column with
 * WaterPoints3D(i,j,KUB) == 1 (all columns when WaterPoints3D is NULL) eliminates rows 1 .. KUB+1 of
 * D(k) x(k-1) + E(k) x(k) + F(k) x(k+1) = TI(k) and back-substitutes into Res (in/out: cells outside solved columns
 * are kept).  This is the entry the reference's legacy GPU path binds as SolveThomas_C(cudaObjID, bounds, D, E, F, TI,
 * Res, dimension = Z) (ModuleCuda.F90:103-111, CudaThomas/Thomas.cu:24-52, kernel DevThomasIK :62-131); the arrays
 * may be host or device pointers, shaped like every 3-D array of the handle. */
int mohid_adt_solve_thomas_z(const int *handle, const double *D, const double *E, const double *F, const double *TI,
                             const int *WaterPoints3D, double *Res);

/* ---- diagnostics ------------------------------------------------------------------ */
/* Copies the last error message of the handle (or of the library when handle is NULL). */
int mohid_adt_last_error(const int *handle, char *buf, const int *buflen);
/* counters[0] = kernels launched since create, [1] = zero-pivot rows seen by the column
 * solver in the last batch (MF:4092-4098 leaves W,G stale; the GPU path counts them),
 * [2] = reserved (0), [3] = bytes of device memory held. */
int mohid_adt_get_counters(const int *handle, long long *counters, const int *n);
/* Average device time in ms of the main transport kernel over the launches since the
 * last call (CUDA events on the handle's stream) and the number of launches averaged. */
int mohid_adt_kernel_time_ms(const int *handle, double *ms, int *launches);
/* Library version / build info string. */
int mohid_adt_version(char *buf, const int *buflen);

#ifdef __cplusplus
}
#endif
#endif /* MOHID_ADT_H */
