!------------------------------------------------------------------------------
!  ModuleAdvectionDiffusionB200 -- reference-side shim (NOT compiled in this repository: the
!  build image has no Fortran compiler; validated by reading and by the ctypes mirror
!  mohid_b200/capi.py, which binds exactly the same symbols with the same argument passing).
!
!  Purpose: let MOHID keep ModuleAdvectionDiffusion's public interface while the work of
!  AdvectionDiffusionIteration (ModuleAdvectionDiffusion.F90:1624-1922) runs on a B200 through the
!  C-ABI of include/mohid_adt.h.  Style follows the reference's own ISO_C_BINDING precedent,
!  ModuleCuda.F90:49-121 (scalars by reference, arrays by base address, integer handle).
!
!  Usage inside ModuleAdvectionDiffusion (under a new -D_USE_B200 flag, analogous to _USE_CUDA):
!    * StartAdvectionDiffusion (AD:400)  -> call B200_Start   (once per instance)
!    * AdvectionDiffusion      (AD:1108) -> after the argument checks and the Get* calls
!      (AD:1217-1401) replace `call AdvectionDiffusionIteration` (AD:1486) by
!      B200_SetStep (first property of the time step only) + B200_AdvectBatch
!    * KillAdvectionDiffusion  (AD:5849) -> call B200_Kill
!  and in ModuleWaterProperties::Advection_Diffusion_Processes (WP:14603-15143) gather the
!  properties that are due (Actual >= NextCompute, equal DTInterval) into one batch instead of
!  calling AdvectionDiffusion once per property (see INTEGRATION.md).
!------------------------------------------------------------------------------
module ModuleAdvectionDiffusionB200

    use, intrinsic :: iso_c_binding
    implicit none
    private

    public :: T_AdtParams, T_AdtOptions, T_AdtSize3D
    public :: B200_Start, B200_Kill, B200_SetGrid2D, B200_SetStep, B200_SetNoFlux, B200_AdvectBatch, B200_LastError
    public :: B200_SetDischarges, B200_UnSetDischarges, B200_GetAdvFlux, B200_GetDifFlux
    public :: B200_UploadProps, B200_AdvectDevice, B200_DownloadProps
    public :: B200_CommInit, B200_ExchangeHalos, B200_CommKill
    ! every entry point of include/mohid_adt.h is bound below (tests/test_fortran_shim.py checks the interface block
    ! against the header: symbol, argument count, by-value / by-reference, base type); the B200_* wrappers cover what
    ! ModuleAdvectionDiffusion and ModuleWaterProperties call, the rest is reached through the interfaces directly
    public :: mohid_adt_set_step_columns, mohid_adt_set_premix, mohid_adt_get_small_depths, mohid_adt_set_offsets
    public :: mohid_adt_set_limits, mohid_adt_get_limit_mass, mohid_adt_set_overlap, mohid_adt_join_halo
    public :: mohid_adt_upload_props_columns, mohid_adt_download_props_columns, mohid_adt_column_mass
    public :: mohid_adt_prop_device_ptr, mohid_adt_synchronize, mohid_adt_sync_prop_buffers, mohid_adt_set_reference_device
    public :: mohid_adt_step_input_device_ptr, mohid_adt_mark_step_resident, mohid_adt_set_active_columns
    public :: mohid_adt_pack_columns, mohid_adt_unpack_columns, mohid_adt_set_stream, mohid_adt_solve_thomas_z
    public :: mohid_adt_get_counters, mohid_adt_kernel_time_ms, mohid_adt_version, mohid_adt_set_boxes, mohid_adt_box_fluxes
    public :: mohid_adt_free_vertical_movement
    public :: mohid_adt_hydro_integration_reinit, mohid_adt_hydro_integration_step, mohid_adt_hydro_integration_end
    public :: mohid_adt_download_step_input

    ! mohid_adt_size3d == T_Size3D (ModuleGlobalData.F90:2041-2052)
    type, bind(c) :: T_AdtSize3D
        integer(c_int) :: ILB, IUB, JLB, JUB, KLB, KUB
    end type T_AdtSize3D

    ! mohid_adt_params: scalar dummies of AdvectionDiffusion (AD:1108-1147); logicals as 0/1
    type, bind(c) :: T_AdtParams
        real(c_double) :: Schmidt_H, SchmidtCoef_V, SchmidtBackground_V
        integer(c_int) :: AdvMethodH, TVDLimitationH, AdvMethodV, TVDLimitationV, Upwind2H, Upwind2V
        real(c_double) :: VolumeRelMax, DTProp, ImpExp_AdvV, ImpExp_DifV, ImpExp_AdvXX, ImpExp_AdvYY, ImpExp_DifH
        integer(c_int) :: NullDif, BoundaryCondition
        real(c_double) :: DecayTime
        integer(c_int) :: NoAdvFlux, NoDifFlux, CellFluxes, Optimize
    end type T_AdtParams

    type, bind(c) :: T_AdtOptions
        integer(c_int) :: Vertical1D, XZFlow, Docycle_method, device, max_properties
        integer(c_int) :: reserved(3)
    end type T_AdtOptions

    interface
        integer(c_int) function mohid_adt_create(handle, size, worksize, ld_i, opt) bind(c, name="mohid_adt_create")
            import :: c_int, T_AdtSize3D, T_AdtOptions
            integer(c_int)     :: handle, ld_i
            type(T_AdtSize3D)  :: size, worksize
            type(T_AdtOptions) :: opt
        end function
        integer(c_int) function mohid_adt_destroy(handle) bind(c, name="mohid_adt_destroy")
            import :: c_int
            integer(c_int) :: handle
        end function
        integer(c_int) function mohid_adt_set_grid2d(handle, DUX, DVY, DZX, DZY, KFloorZ, BoundaryPoints2D) &
                bind(c, name="mohid_adt_set_grid2d")
            import :: c_int, c_double
            integer(c_int)                                 :: handle
            real(c_double), dimension(*)                   :: DUX, DVY, DZX, DZY
            integer(c_int), dimension(*)                   :: KFloorZ, BoundaryPoints2D
        end function
        integer(c_int) function mohid_adt_set_step(handle, Wflux_X, Wflux_Y, Wflux_Z, VolumeZOld, VolumeZ,      &
                Visc_H, Diff_V, DWZ, DZZ, AreaU, AreaV, OpenPoints3D, LandPoints3D, WaterPoints3D,               &
                ComputeFacesU3D, ComputeFacesV3D, ComputeFacesW3D, SmallDepths) bind(c, name="mohid_adt_set_step")
            import :: c_int, c_ptr
            integer(c_int)     :: handle
            ! c_loc(array(0,0,0)); after the first complete call c_null_ptr = unchanged since the last step
            type(c_ptr), value :: Wflux_X, Wflux_Y, Wflux_Z, VolumeZOld, VolumeZ, Visc_H, Diff_V, DWZ, DZZ, AreaU, AreaV
            type(c_ptr), value :: OpenPoints3D, LandPoints3D, WaterPoints3D, ComputeFacesU3D, ComputeFacesV3D, ComputeFacesW3D
            type(c_ptr), value :: SmallDepths                    ! c_null_ptr when not present (AD:1297-1302)
        end function
        integer(c_int) function mohid_adt_set_noflux(handle, NoFluxU, NoFluxV, NoFluxW) bind(c, name="mohid_adt_set_noflux")
            import :: c_int, c_ptr
            integer(c_int)     :: handle
            type(c_ptr), value :: NoFluxU, NoFluxV, NoFluxW   ! c_loc(NoFlux?(0,0,0)) or c_null_ptr (all three)
        end function
        integer(c_int) function mohid_adt_set_premix(handle, Density, WaterColumnZ, SmallDepthsLimit) &
                bind(c, name="mohid_adt_set_premix")
            import :: c_int, c_ptr, c_double
            integer(c_int)     :: handle
            type(c_ptr), value :: Density, WaterColumnZ          ! c_loc(...(0,0,0)) / c_loc(...(0,0)) or c_null_ptr
            real(c_double)     :: SmallDepthsLimit
        end function
        integer(c_int) function mohid_adt_set_offsets(handle, nprop, OffSet) bind(c, name="mohid_adt_set_offsets")
            import :: c_int, c_double
            integer(c_int)               :: handle, nprop
            real(c_double), dimension(*) :: OffSet
        end function
        integer(c_int) function mohid_adt_set_limits(handle, nprop, MinOn, MinValue, MaxOn, MaxValue) &
                bind(c, name="mohid_adt_set_limits")
            import :: c_int, c_double
            integer(c_int)               :: handle, nprop
            integer(c_int), dimension(*) :: MinOn, MaxOn            ! Property%Evolution%MinConcentration / MaxConcentration as 0/1
            real(c_double), dimension(*) :: MinValue, MaxValue
        end function
        integer(c_int) function mohid_adt_get_limit_mass(handle, prop_index, Mass_Created, Mass_Destroid) &
                bind(c, name="mohid_adt_get_limit_mass")
            import :: c_int, c_ptr
            integer(c_int)     :: handle, prop_index
            type(c_ptr), value :: Mass_Created, Mass_Destroid       ! c_loc(Property%Mass_created(0,0,0)) or c_null_ptr
        end function
        integer(c_int) function mohid_adt_advect_batch(handle, nprop, prop, reference_prop, params) &
                bind(c, name="mohid_adt_advect_batch")
            import :: c_int, c_ptr, T_AdtParams
            integer(c_int)                  :: handle, nprop
            type(c_ptr), dimension(*)       :: prop              ! c_loc of each Property%Concentration(0,0,0)
            type(c_ptr), value              :: reference_prop    ! c_loc of an array of c_ptr, or c_null_ptr
            type(T_AdtParams), dimension(*) :: params
        end function
        integer(c_int) function mohid_adt_last_error(handle, buf, buflen) bind(c, name="mohid_adt_last_error")
            import :: c_int, c_char
            integer(c_int)                       :: handle, buflen
            character(kind=c_char), dimension(*) :: buf
        end function
        integer(c_int) function mohid_adt_set_step_columns(handle, j0, ncols, Wflux_X, Wflux_Y, Wflux_Z, VolumeZOld,   &
                VolumeZ, Visc_H, Diff_V, DWZ, DZZ, AreaU, AreaV, OpenPoints3D, LandPoints3D, WaterPoints3D,             &
                ComputeFacesU3D, ComputeFacesV3D, ComputeFacesW3D) bind(c, name="mohid_adt_set_step_columns")
            import :: c_int, c_ptr
            integer(c_int)     :: handle, j0, ncols
            ! windows of the arrays, (ld_i, ncols, K+2) each; c_null_ptr = skipped
            type(c_ptr), value :: Wflux_X, Wflux_Y, Wflux_Z, VolumeZOld, VolumeZ, Visc_H, Diff_V, DWZ, DZZ, AreaU, AreaV
            type(c_ptr), value :: OpenPoints3D, LandPoints3D, WaterPoints3D, ComputeFacesU3D, ComputeFacesV3D, ComputeFacesW3D
        end function
        integer(c_int) function mohid_adt_get_small_depths(handle, SmallDepthsOn) bind(c, name="mohid_adt_get_small_depths")
            import :: c_int
            integer(c_int)               :: handle
            integer(c_int), dimension(*) :: SmallDepthsOn
        end function
        integer(c_int) function mohid_adt_set_overlap(handle, ghost, comm_stream) bind(c, name="mohid_adt_set_overlap")
            import :: c_int, c_ptr
            integer(c_int)     :: handle, ghost
            type(c_ptr), value :: comm_stream                       ! cudaStream_t
        end function
        integer(c_int) function mohid_adt_join_halo(handle) bind(c, name="mohid_adt_join_halo")
            import :: c_int
            integer(c_int) :: handle
        end function
        ! SetDischarges (AD:978-1034): same argument names; DischConcMF may be c_null_ptr
        integer(c_int) function mohid_adt_set_discharges(handle, prop_index, DischNumber, n_cells, DischFlow, DischConc,  &
                DischI, DischJ, DischK, DischKmin, DischKmax, DischVert, IgnoreDisch, DischnCells, ByPass, DischConcMF)  &
                bind(c, name="mohid_adt_set_discharges")
            import :: c_int, c_double, c_ptr
            integer(c_int)               :: handle, prop_index, DischNumber, n_cells
            real(c_double), dimension(*) :: DischFlow, DischConc
            integer(c_int), dimension(*) :: DischI, DischJ, DischK, DischKmin, DischKmax, DischVert
            integer(c_int), dimension(*) :: IgnoreDisch, DischnCells, ByPass          ! logicals as 0/1
            type(c_ptr), value           :: DischConcMF
        end function
        integer(c_int) function mohid_adt_unset_discharges(handle) bind(c, name="mohid_adt_unset_discharges")
            import :: c_int
            integer(c_int) :: handle
        end function
        ! GetAdvFlux / GetDifFlux (AD:697-851) of a property advanced with CellFluxes = 1
        integer(c_int) function mohid_adt_get_cell_fluxes(handle, prop_index, AdvFluxX, AdvFluxY, AdvFluxZ, DifFluxX,    &
                DifFluxY, DifFluxZ) bind(c, name="mohid_adt_get_cell_fluxes")
            import :: c_int, c_double
            integer(c_int)               :: handle, prop_index
            real(c_double), dimension(*) :: AdvFluxX, AdvFluxY, AdvFluxZ, DifFluxX, DifFluxY, DifFluxZ
        end function
        ! BoxDifFluxes3D on the device (ModuleBoxDif.F90:2659-2776 as called at WP:14992-15001)
        integer(c_int) function mohid_adt_set_boxes(handle, Boxes3D, NumberOfBoxes3D) bind(c, name="mohid_adt_set_boxes")
            import :: c_int
            integer(c_int)               :: handle, NumberOfBoxes3D
            integer(c_int), dimension(*) :: Boxes3D
        end function
        integer(c_int) function mohid_adt_box_fluxes(handle, prop_index, Fluxes3D) bind(c, name="mohid_adt_box_fluxes")
            import :: c_int, c_double
            integer(c_int)               :: handle, prop_index
            real(c_double), dimension(*) :: Fluxes3D              ! (0:NumberOfBoxes3D, 0:NumberOfBoxes3D)
        end function
        ! FreeVerticalMovementIteration (ModuleFreeVerticalMovement.F90:1531-1650) on a device-resident property
        integer(c_int) function mohid_adt_free_vertical_movement(handle, prop_index, Velocity, GridCellArea,             &
                DepositionProbability, Deposition, NonCohesive, DepositionIntertidalZones, ImpExp_AdvV, DTProp,          &
                FreeConvFlux) bind(c, name="mohid_adt_free_vertical_movement")
            import :: c_int, c_double, c_ptr
            integer(c_int)               :: handle, prop_index, Deposition, NonCohesive, DepositionIntertidalZones
            real(c_double), dimension(*) :: Velocity, GridCellArea
            type(c_ptr), value           :: DepositionProbability, FreeConvFlux     ! c_null_ptr when not needed
            real(c_double)               :: ImpExp_AdvV, DTProp
        end function
        ! ModuleHydroIntegration on the device mirrors (ModuleHydroIntegration.F90:767-994, used at WP:14615-14647)
        integer(c_int) function mohid_adt_hydro_integration_reinit(handle, VolumeZOld) &
                bind(c, name="mohid_adt_hydro_integration_reinit")
            import :: c_int, c_double
            integer(c_int)               :: handle
            real(c_double), dimension(*) :: VolumeZOld
        end function
        integer(c_int) function mohid_adt_hydro_integration_step(handle, WaterFluxX, WaterFluxY, Discharges, ComputeFacesU,  &
                ComputeFacesV) bind(c, name="mohid_adt_hydro_integration_step")
            import :: c_int, c_double, c_ptr
            integer(c_int)               :: handle
            real(c_double), dimension(*) :: WaterFluxX, WaterFluxY
            type(c_ptr), value           :: Discharges                              ! c_null_ptr: no discharges
            integer(c_int), dimension(*) :: ComputeFacesU, ComputeFacesV
        end function
        integer(c_int) function mohid_adt_hydro_integration_end(handle, VolumeZ, WaterPoints3D, DT) &
                bind(c, name="mohid_adt_hydro_integration_end")
            import :: c_int, c_double
            integer(c_int)               :: handle
            real(c_double), dimension(*) :: VolumeZ
            integer(c_int), dimension(*) :: WaterPoints3D
            real(c_double)               :: DT
        end function
        integer(c_int) function mohid_adt_download_step_input(handle, which, array) bind(c, name="mohid_adt_download_step_input")
            import :: c_int, c_ptr
            integer(c_int)     :: handle, which
            type(c_ptr), value :: array                     ! c_loc of a real(8) (which <= 10) or integer (11..16) 3-D array
        end function
        ! device-resident properties: upload once, advance nsteps without host traffic, download when needed
        integer(c_int) function mohid_adt_upload_props(handle, nprop, prop, reference_prop) bind(c, name="mohid_adt_upload_props")
            import :: c_int, c_ptr
            integer(c_int)            :: handle, nprop
            type(c_ptr), dimension(*) :: prop
            type(c_ptr), value        :: reference_prop          ! c_loc of an array of c_ptr, or c_null_ptr
        end function
        integer(c_int) function mohid_adt_download_props(handle, nprop, prop) bind(c, name="mohid_adt_download_props")
            import :: c_int, c_ptr
            integer(c_int)            :: handle, nprop
            type(c_ptr), dimension(*) :: prop
        end function
        integer(c_int) function mohid_adt_advect_device(handle, nprop, params, nsteps) bind(c, name="mohid_adt_advect_device")
            import :: c_int, T_AdtParams
            integer(c_int)                  :: handle, nprop, nsteps
            type(T_AdtParams), dimension(*) :: params
        end function
        integer(c_int) function mohid_adt_upload_props_columns(handle, nprop, prop, reference_prop, j0, ncols) &
                bind(c, name="mohid_adt_upload_props_columns")
            import :: c_int, c_ptr
            integer(c_int)            :: handle, nprop, j0, ncols
            type(c_ptr), dimension(*) :: prop
            type(c_ptr), value        :: reference_prop
        end function
        integer(c_int) function mohid_adt_download_props_columns(handle, nprop, prop, j0, ncols) &
                bind(c, name="mohid_adt_download_props_columns")
            import :: c_int, c_ptr
            integer(c_int)            :: handle, nprop, j0, ncols
            type(c_ptr), dimension(*) :: prop
        end function
        integer(c_int) function mohid_adt_column_mass(handle, nprop, mass) bind(c, name="mohid_adt_column_mass")
            import :: c_int, c_double
            integer(c_int)               :: handle, nprop
            real(c_double), dimension(*) :: mass                    ! (0:J+1, nprop)
        end function
        integer(c_int) function mohid_adt_prop_device_ptr(handle, n, dptr, ld, nj, nk) bind(c, name="mohid_adt_prop_device_ptr")
            import :: c_int, c_ptr
            integer(c_int) :: handle, n, ld, nj, nk
            type(c_ptr)    :: dptr                                  ! out: device address
        end function
        integer(c_int) function mohid_adt_synchronize(handle) bind(c, name="mohid_adt_synchronize")
            import :: c_int
            integer(c_int) :: handle
        end function
        integer(c_int) function mohid_adt_sync_prop_buffers(handle, nprop) bind(c, name="mohid_adt_sync_prop_buffers")
            import :: c_int
            integer(c_int) :: handle, nprop
        end function
        integer(c_int) function mohid_adt_set_reference_device(handle, n, dptr) bind(c, name="mohid_adt_set_reference_device")
            import :: c_int, c_ptr
            integer(c_int) :: handle, n
            type(c_ptr)    :: dptr
        end function
        integer(c_int) function mohid_adt_step_input_device_ptr(handle, which, dptr, ld, nj, nk) &
                bind(c, name="mohid_adt_step_input_device_ptr")
            import :: c_int, c_ptr
            integer(c_int) :: handle, which, ld, nj, nk
            type(c_ptr)    :: dptr
        end function
        integer(c_int) function mohid_adt_mark_step_resident(handle, small_depths_present) &
                bind(c, name="mohid_adt_mark_step_resident")
            import :: c_int
            integer(c_int) :: handle, small_depths_present
        end function
        ! one MPI sub-domain = one column slab: the owned columns, the halo staging and the NCCL exchange that
        ! replaces ReceiveSendProperitiesMPI (ModuleHorizontalGrid.F90:8479-8658)
        integer(c_int) function mohid_adt_set_active_columns(handle, j_begin, j_count) bind(c, name="mohid_adt_set_active_columns")
            import :: c_int
            integer(c_int) :: handle, j_begin, j_count
        end function
        integer(c_int) function mohid_adt_pack_columns(handle, nprop, j0, width, device_buffer) bind(c, name="mohid_adt_pack_columns")
            import :: c_int, c_ptr
            integer(c_int)     :: handle, nprop, j0, width
            type(c_ptr), value :: device_buffer
        end function
        integer(c_int) function mohid_adt_unpack_columns(handle, nprop, j0, width, device_buffer) &
                bind(c, name="mohid_adt_unpack_columns")
            import :: c_int, c_ptr
            integer(c_int)     :: handle, nprop, j0, width
            type(c_ptr), value :: device_buffer
        end function
        integer(c_int) function mohid_adt_set_stream(handle, cuda_stream) bind(c, name="mohid_adt_set_stream")
            import :: c_int, c_ptr
            integer(c_int)     :: handle
            type(c_ptr), value :: cuda_stream
        end function
        integer(c_int) function mohid_adt_comm_get_unique_id(unique_id, nbytes) bind(c, name="mohid_adt_comm_get_unique_id")
            import :: c_int, c_char
            character(kind=c_char), dimension(*) :: unique_id       ! 128 bytes: rank 0 obtains it, MPI_Bcast hands it on
            integer(c_int)                       :: nbytes
        end function
        integer(c_int) function mohid_adt_comm_init(handle, nranks, rank, unique_id, ghost, overlap) &
                bind(c, name="mohid_adt_comm_init")
            import :: c_int, c_char
            integer(c_int)                       :: handle, nranks, rank, ghost, overlap
            character(kind=c_char), dimension(*) :: unique_id
        end function
        integer(c_int) function mohid_adt_exchange_halos(handle, nprop) bind(c, name="mohid_adt_exchange_halos")
            import :: c_int
            integer(c_int) :: handle, nprop
        end function
        integer(c_int) function mohid_adt_comm_destroy(handle) bind(c, name="mohid_adt_comm_destroy")
            import :: c_int
            integer(c_int) :: handle
        end function
        ! THOMASZ_NewType2 on caller-supplied coefficient fields (ModuleCuda.F90:103-111 SolveThomas_C analogue)
        integer(c_int) function mohid_adt_solve_thomas_z(handle, D, E, F, TI, WaterPoints3D, Res) &
                bind(c, name="mohid_adt_solve_thomas_z")
            import :: c_int, c_double
            integer(c_int)               :: handle
            real(c_double), dimension(*) :: D, E, F, TI, Res
            integer(c_int), dimension(*) :: WaterPoints3D
        end function
        integer(c_int) function mohid_adt_get_counters(handle, counters, n) bind(c, name="mohid_adt_get_counters")
            import :: c_int, c_long_long
            integer(c_int)                     :: handle, n
            integer(c_long_long), dimension(*) :: counters
        end function
        integer(c_int) function mohid_adt_kernel_time_ms(handle, ms, launches) bind(c, name="mohid_adt_kernel_time_ms")
            import :: c_int, c_double
            integer(c_int) :: handle, launches
            real(c_double) :: ms
        end function
        integer(c_int) function mohid_adt_version(buf, buflen) bind(c, name="mohid_adt_version")
            import :: c_int, c_char
            character(kind=c_char), dimension(*) :: buf
            integer(c_int)                       :: buflen
        end function
    end interface

contains

    !--------------------------------------------------------------------------
    ! StartAdvectionDiffusion (AD:400-533): Size / WorkSize come from GetGeometrySize.
    subroutine B200_Start(Handle, Size, WorkSize, LeadingDim, Vertical1D, XZFlow, Docycle_method, Device, STAT)
        integer(c_int),    intent(OUT) :: Handle
        type(T_AdtSize3D), intent(IN)  :: Size, WorkSize
        integer,           intent(IN)  :: LeadingDim          ! Pad(ILB,IUB)-ILB+1 under _PAD_MATRICES, else IUB-ILB+1
        logical,           intent(IN)  :: Vertical1D, XZFlow
        integer,           intent(IN)  :: Docycle_method, Device
        integer,           intent(OUT) :: STAT
        type(T_AdtOptions) :: opt
        integer(c_int)     :: ld
        opt%Vertical1D = merge(1, 0, Vertical1D); opt%XZFlow = merge(1, 0, XZFlow)
        opt%Docycle_method = Docycle_method; opt%device = Device; opt%max_properties = 0; opt%reserved = 0
        ld = LeadingDim
        STAT = mohid_adt_create(Handle, Size, WorkSize, ld, opt)
    end subroutine B200_Start

    subroutine B200_Kill(Handle, STAT)
        integer(c_int), intent(INOUT) :: Handle
        integer,        intent(OUT)   :: STAT
        STAT = mohid_adt_destroy(Handle)
    end subroutine B200_Kill

    ! What AD:1353-1384 fetches (GetHorizontalGrid, GetGeometryKFloor, GetBoundaries); all arrays
    ! are (ILB:IUB, JLB:JUB) = (0:I+1, 0:J+1): pass the whole array, element (0,0) first.
    subroutine B200_SetGrid2D(Handle, DUX, DVY, DZX, DZY, KFloorZ, BoundaryPoints2D, STAT)
        integer(c_int)                            :: Handle
        real(c_double), dimension(:,:), pointer   :: DUX, DVY, DZX, DZY
        integer(c_int), dimension(:,:), pointer   :: KFloorZ, BoundaryPoints2D
        integer, intent(OUT)                      :: STAT
        STAT = mohid_adt_set_grid2d(Handle, DUX, DVY, DZX, DZY, KFloorZ, BoundaryPoints2D)
    end subroutine B200_SetGrid2D

    ! Once per time step (FirstProperty): the shared array dummies of AD:1132-1141 + Geometry getters.
    ! SmallDepths is a Fortran logical(:,:) in the reference: convert to integer 0/1 before the call.
    subroutine B200_SetStep(Handle, Wflux_X, Wflux_Y, Wflux_Z, VolumeZOld, VolumeZ, Visc_H, Diff_V, DWZ, DZZ,    &
                            AreaU, AreaV, OpenPoints3D, LandPoints3D, WaterPoints3D, ComputeFacesU3D,            &
                            ComputeFacesV3D, ComputeFacesW3D, SmallDepthsInt, STAT)
        integer(c_int)                              :: Handle
        real(c_double), dimension(:,:,:), pointer   :: Wflux_X, Wflux_Y, Wflux_Z, VolumeZOld, VolumeZ
        real(c_double), dimension(:,:,:), pointer   :: Visc_H, Diff_V, DWZ, DZZ, AreaU, AreaV
        integer(c_int), dimension(:,:,:), pointer   :: OpenPoints3D, LandPoints3D, WaterPoints3D
        integer(c_int), dimension(:,:,:), pointer   :: ComputeFacesU3D, ComputeFacesV3D, ComputeFacesW3D
        integer(c_int), dimension(:,:), pointer     :: SmallDepthsInt
        integer, intent(OUT)                        :: STAT
        type(c_ptr) :: sd
        sd = c_null_ptr
        if (associated(SmallDepthsInt)) sd = c_loc(SmallDepthsInt(lbound(SmallDepthsInt,1), lbound(SmallDepthsInt,2)))
        ! a pointer that is not associated travels as NULL = "unchanged since the last step" (the land / water maps never
        ! change, the other masks only with wetting and drying): 4 of the 112 bytes per cell stay off the bus
        STAT = mohid_adt_set_step(Handle, L3R(Wflux_X), L3R(Wflux_Y), L3R(Wflux_Z), L3R(VolumeZOld), L3R(VolumeZ),   &
                                  L3R(Visc_H), L3R(Diff_V), L3R(DWZ), L3R(DZZ), L3R(AreaU), L3R(AreaV),              &
                                  L3I(OpenPoints3D), L3I(LandPoints3D), L3I(WaterPoints3D), L3I(ComputeFacesU3D),    &
                                  L3I(ComputeFacesV3D), L3I(ComputeFacesW3D), sd)
    end subroutine B200_SetStep

    ! address of the first element (ILB,JLB,KLB) of a 3-D array, c_null_ptr when the pointer is not associated
    function L3R(A) result(p)
        real(c_double), dimension(:,:,:), pointer :: A
        type(c_ptr) :: p
        p = c_null_ptr
        if (associated(A)) p = c_loc(A(lbound(A,1), lbound(A,2), lbound(A,3)))
    end function L3R

    function L3I(A) result(p)
        integer(c_int), dimension(:,:,:), pointer :: A
        type(c_ptr) :: p
        p = c_null_ptr
        if (associated(A)) p = c_loc(A(lbound(A,1), lbound(A,2), lbound(A,3)))
    end function L3I

    ! The optional NoFluxU/V/W dummies (AD:1143-1146); not associated = not present.
    subroutine B200_SetNoFlux(Handle, NoFluxU, NoFluxV, NoFluxW, STAT)
        integer(c_int)                            :: Handle
        integer(c_int), dimension(:,:,:), pointer :: NoFluxU, NoFluxV, NoFluxW
        integer, intent(OUT)                      :: STAT
        if (associated(NoFluxU) .and. associated(NoFluxV) .and. associated(NoFluxW)) then
            STAT = mohid_adt_set_noflux(Handle, c_loc(NoFluxU(0,0,0)), c_loc(NoFluxV(0,0,0)), c_loc(NoFluxW(0,0,0)))
        else
            STAT = mohid_adt_set_noflux(Handle, c_null_ptr, c_null_ptr, c_null_ptr)
        endif
    end subroutine B200_SetNoFlux

    ! The batched replacement of the per-property call loop (WP:14603-15143 -> AD:1108).
    ! PropPtr(n) = c_loc(Property%Concentration(0,0,0)); RefPtr(n) = c_loc(Property%Assimilation%Field(0,0,0))
    ! or c_null_ptr.  On failure the caller stops like the reference does:
    !     if (STAT /= SUCCESS_) stop 'AdvectionDiffusion - ModuleAdvectionDiffusion - ERR_B200'
    subroutine B200_AdvectBatch(Handle, nProp, PropPtr, RefPtr, Params, STAT)
        integer(c_int)                              :: Handle
        integer, intent(IN)                         :: nProp
        type(c_ptr), dimension(:), target           :: PropPtr, RefPtr
        type(T_AdtParams), dimension(:)             :: Params
        integer, intent(OUT)                        :: STAT
        integer(c_int) :: n
        n = nProp
        STAT = mohid_adt_advect_batch(Handle, n, PropPtr, c_loc(RefPtr(1)), Params)
    end subroutine B200_AdvectBatch

    subroutine B200_LastError(Handle, Message)
        integer(c_int)                 :: Handle
        character(len=*), intent(OUT)  :: Message
        character(kind=c_char), dimension(512) :: buf
        integer(c_int) :: n, i, rc
        n = 512
        rc = mohid_adt_last_error(Handle, buf, n)
        Message = ' '
        do i = 1, min(len(Message), 511)
            if (buf(i) == c_null_char) exit
            Message(i:i) = buf(i)
        enddo
    end subroutine B200_LastError

    !--------------------------------------------------------------------------
    ! SetDischarges (AD:978-1034) for the property at position PropIndex (0-based) of the next batch; the logical
    ! arrays of the reference (IgnoreDisch, ByPass) arrive as integer 0/1.
    subroutine B200_SetDischarges(Handle, PropIndex, DischFlow, DischConc, DischI, DischJ, DischK, DischKmin, DischKmax, &
                                  DischVert, IgnoreDisch, DischnCells, ByPass, DischNumber, nCells, STAT)
        integer(c_int)                        :: Handle
        integer, intent(IN)                   :: PropIndex, DischNumber, nCells
        real(c_double), dimension(:), pointer :: DischFlow, DischConc
        integer(c_int), dimension(:), pointer :: DischI, DischJ, DischK, DischKmin, DischKmax, DischVert
        integer(c_int), dimension(:), pointer :: IgnoreDisch, DischnCells, ByPass
        integer, intent(OUT)                  :: STAT
        integer(c_int) :: ip, nd, nc
        ip = PropIndex; nd = DischNumber; nc = nCells
        STAT = mohid_adt_set_discharges(Handle, ip, nd, nc, DischFlow, DischConc, DischI, DischJ, DischK, DischKmin,     &
                                        DischKmax, DischVert, IgnoreDisch, DischnCells, ByPass, c_null_ptr)
    end subroutine B200_SetDischarges

    ! UnSetDischarges (AD:1040-1095)
    subroutine B200_UnSetDischarges(Handle, STAT)
        integer(c_int)       :: Handle
        integer, intent(OUT) :: STAT
        STAT = mohid_adt_unset_discharges(Handle)
    end subroutine B200_UnSetDischarges

    ! GetAdvFlux / GetDifFlux (AD:697-851): the reference hands out pointers to its module arrays; here the caller's
    ! arrays (0:I+1, 0:J+1, 0:K+1) are filled.  Both are served by one library call.
    subroutine B200_GetAdvFlux(Handle, PropIndex, AdvFluxX, AdvFluxY, AdvFluxZ, DifFluxX, DifFluxY, DifFluxZ, STAT)
        integer(c_int)                            :: Handle
        integer, intent(IN)                       :: PropIndex
        real(c_double), dimension(:,:,:), pointer :: AdvFluxX, AdvFluxY, AdvFluxZ, DifFluxX, DifFluxY, DifFluxZ
        integer, intent(OUT)                      :: STAT
        integer(c_int) :: ip
        ip = PropIndex
        STAT = mohid_adt_get_cell_fluxes(Handle, ip, AdvFluxX, AdvFluxY, AdvFluxZ, DifFluxX, DifFluxY, DifFluxZ)
    end subroutine B200_GetAdvFlux

    subroutine B200_GetDifFlux(Handle, PropIndex, AdvFluxX, AdvFluxY, AdvFluxZ, DifFluxX, DifFluxY, DifFluxZ, STAT)
        integer(c_int)                            :: Handle
        integer, intent(IN)                       :: PropIndex
        real(c_double), dimension(:,:,:), pointer :: AdvFluxX, AdvFluxY, AdvFluxZ, DifFluxX, DifFluxY, DifFluxZ
        integer, intent(OUT)                      :: STAT
        call B200_GetAdvFlux(Handle, PropIndex, AdvFluxX, AdvFluxY, AdvFluxZ, DifFluxX, DifFluxY, DifFluxZ, STAT)
    end subroutine B200_GetDifFlux

    !--------------------------------------------------------------------------
    ! Resident-property mode: the concentrations stay on the device between time steps (no 16 B per cell and property
    ! over PCIe every step); the host downloads them when an output, a sink/source module or an assimilation step needs
    ! them.  B200_SetStep may pass non-associated masks after the first step (land / water maps never change).
    subroutine B200_UploadProps(Handle, nProp, PropPtr, RefPtr, STAT)
        integer(c_int)                    :: Handle
        integer, intent(IN)               :: nProp
        type(c_ptr), dimension(:), target :: PropPtr, RefPtr
        integer, intent(OUT)              :: STAT
        integer(c_int) :: n
        n = nProp
        STAT = mohid_adt_upload_props(Handle, n, PropPtr, c_loc(RefPtr(1)))
    end subroutine B200_UploadProps

    subroutine B200_AdvectDevice(Handle, nProp, Params, nSteps, STAT)
        integer(c_int)                  :: Handle
        integer, intent(IN)             :: nProp, nSteps
        type(T_AdtParams), dimension(:) :: Params
        integer, intent(OUT)            :: STAT
        integer(c_int) :: n, ns
        n = nProp; ns = nSteps
        STAT = mohid_adt_advect_device(Handle, n, Params, ns)
    end subroutine B200_AdvectDevice

    subroutine B200_DownloadProps(Handle, nProp, PropPtr, STAT)
        integer(c_int)            :: Handle
        integer, intent(IN)       :: nProp
        type(c_ptr), dimension(:) :: PropPtr
        integer, intent(OUT)      :: STAT
        integer(c_int) :: n
        n = nProp
        STAT = mohid_adt_download_props(Handle, n, PropPtr)
    end subroutine B200_DownloadProps

    !--------------------------------------------------------------------------
    ! Domain decomposition (ModuleHorizontalGrid.F90:1587-1757): every MPI rank drives one GPU and one column slab.
    ! Rank 0 obtains the NCCL id, MPI_Bcast carries its 128 bytes to the other ranks (UniqueId), and the exchange that
    ! ReceiveSendProperitiesMPI (HG:8479-8658, called at WP:15034-15045) performs on host arrays becomes one call on
    ! the device-resident fields.  JBegin / JCount: the owned local columns (the halo of width Ghost lies outside them).
    subroutine B200_CommInit(Handle, nRanks, Rank, UniqueId, JBegin, JCount, Ghost, STAT)
        integer(c_int)                                        :: Handle
        integer, intent(IN)                                   :: nRanks, Rank, JBegin, JCount, Ghost
        character(kind=c_char), dimension(128), intent(INOUT) :: UniqueId
        integer, intent(OUT)                                  :: STAT
        integer(c_int) :: nr, r, jb, jc, g, ov
        nr = nRanks; r = Rank; jb = JBegin; jc = JCount; g = Ghost; ov = 1
        STAT = mohid_adt_set_active_columns(Handle, jb, jc)
        if (STAT /= 0) return
        STAT = mohid_adt_comm_init(Handle, nr, r, UniqueId, g, ov)
    end subroutine B200_CommInit

    subroutine B200_ExchangeHalos(Handle, nProp, STAT)
        integer(c_int)       :: Handle
        integer, intent(IN)  :: nProp
        integer, intent(OUT) :: STAT
        integer(c_int) :: n
        n = nProp
        STAT = mohid_adt_exchange_halos(Handle, n)
    end subroutine B200_ExchangeHalos

    subroutine B200_CommKill(Handle, STAT)
        integer(c_int)       :: Handle
        integer, intent(OUT) :: STAT
        STAT = mohid_adt_comm_destroy(Handle)
    end subroutine B200_CommKill

end module ModuleAdvectionDiffusionB200
