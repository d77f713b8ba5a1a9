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
            import :: c_int, c_double, c_ptr
            integer(c_int)               :: handle
            real(c_double), dimension(*) :: Wflux_X, Wflux_Y, Wflux_Z, VolumeZOld, VolumeZ, Visc_H, Diff_V
            real(c_double), dimension(*) :: DWZ, DZZ, AreaU, AreaV
            integer(c_int), dimension(*) :: OpenPoints3D, LandPoints3D, WaterPoints3D
            integer(c_int), dimension(*) :: ComputeFacesU3D, ComputeFacesV3D, ComputeFacesW3D
            type(c_ptr), value           :: SmallDepths          ! c_null_ptr when not present (AD:1297-1302)
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
        STAT = mohid_adt_set_step(Handle, Wflux_X, Wflux_Y, Wflux_Z, VolumeZOld, VolumeZ, Visc_H, Diff_V, DWZ, DZZ, &
                                  AreaU, AreaV, OpenPoints3D, LandPoints3D, WaterPoints3D, ComputeFacesU3D,          &
                                  ComputeFacesV3D, ComputeFacesW3D, sd)
    end subroutine B200_SetStep

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

end module ModuleAdvectionDiffusionB200
