"""Host-side mirror of MOHID's ``ModuleAdvectionDiffusion`` public interface on top of the C-ABI.

Reference interface (``/root/reference/Software/MOHIDBase2/ModuleAdvectionDiffusion.F90``):
``StartAdvectionDiffusion`` (:400), ``AdvectionDiffusion`` (:1108), ``SetDischarges`` (:978),
``UnSetDischarges`` (:1040), ``GetBoundaryConditionList`` (:855), ``KillAdvectionDiffusion`` (:5849).
The module-level functions below keep those names, argument names and error behaviour (the
reference's ``stop '... ERRnn'`` becomes :class:`mohid_b200.capi.AdtError` carrying the same text);
:class:`TransportStep` is the batched / device-resident form the new path adds: all properties
of a time step advance in one call (the loop of ``ModuleWaterProperties.F90:14603-15143``).

Arrays are Fortran-ordered ``(0:I+1, 0:J+1, 0:K+1)`` with ``i`` contiguous, i.e. numpy / torch
arrays of shape ``(K+2, J+2, ld)``; they may live in host memory or (torch CUDA tensors) on the
device.  Everything is computed by the CUDA library; nothing here does arithmetic.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import capi
from .capi import AdtError, Options, Params, Size3D, check, make_params

# GetBoundaryConditionList (AD:855-893)
MassConservation_, ImposedValue_, NullGradient_, SubModel_, Orlanski_, MassConservNullGrad_, CyclicBoundary_ = \
    1, 2, 4, 5, 6, 7, 8
UpwindOrder1, UpwindOrder2, UpwindOrder3, P2_TVD, CentralDif, LeapFrog = 1, 2, 3, 4, 5, 6
MinMod, VanLeer, Muscl, SuperBee, PDM = 1, 2, 3, 4, 5
SUCCESS_ = 0

STEP_F64 = ["Wflux_X", "Wflux_Y", "Wflux_Z", "VolumeZOld", "VolumeZ", "Visc_H", "Diff_V", "DWZ", "DZZ", "AreaU", "AreaV"]
STEP_I32 = ["OpenPoints3D", "LandPoints3D", "WaterPoints3D", "ComputeFacesU3D", "ComputeFacesV3D", "ComputeFacesW3D"]


def _is_torch(a) -> bool:
    return type(a).__module__.startswith("torch")


def _ptr(a, dtype: str, nelem: int, name: str) -> C.c_void_p:
    """Base address of a contiguous host (numpy / torch CPU) or device (torch CUDA) array."""
    if a is None:
        return C.c_void_p(None)
    if _is_torch(a):
        import torch
        want = {"f8": torch.float64, "i4": torch.int32}[dtype]
        if a.dtype != want or not a.is_contiguous() or a.numel() != nelem:
            raise ValueError(f"{name}: expected contiguous {want} with {nelem} elements, got {a.dtype} {tuple(a.shape)}")
        return C.c_void_p(a.data_ptr())
    a_np = np.asarray(a)
    want = np.dtype(dtype)
    if a_np.dtype != want or not a_np.flags["C_CONTIGUOUS"] or a_np.size != nelem:
        raise ValueError(f"{name}: expected C-contiguous {want} with {nelem} elements, got {a_np.dtype} {a_np.shape}")
    return C.c_void_p(a_np.ctypes.data)


class TransportStep:
    """One ``ObjAdvectionDiffusion`` instance bound to one GPU."""

    def __init__(self, I: int, J: int, K: int, ld: Optional[int] = None, *, vertical1d: bool = False,
                 xzflow: bool = False, docycle_method: int = 1, device: int = -1, max_properties: int = 0):
        self.lib = capi.load()
        self.I, self.J, self.K = int(I), int(J), int(K)
        self.ld = int(ld) if ld else self.I + 2
        self.n2 = self.ld * (self.J + 2)
        self.n3 = self.n2 * (self.K + 2)
        self.h = C.c_int(0)
        size = Size3D(0, I + 1, 0, J + 1, 0, K + 1)
        work = Size3D(1, I, 1, J, 1, K)
        opt = Options(int(vertical1d), int(xzflow), int(docycle_method), int(device), int(max_properties))
        check(self.lib.mohid_adt_create(C.byref(self.h), C.byref(size), C.byref(work), C.byref(C.c_int(self.ld)),
                                        C.byref(opt)))
        self.nprop = 0

    # ---- lifetime -------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.mohid_adt_destroy(C.byref(self.h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        check(rc, self.h)

    # ---- inputs ---------------------------------------------------------------------
    def set_grid2d(self, DUX, DVY, DZX, DZY, KFloorZ, BoundaryPoints2D):
        """GetHorizontalGrid / GetGeometryKFloor / GetBoundaries results (AD:1353-1384)."""
        self._check(self.lib.mohid_adt_set_grid2d(
            C.byref(self.h), _ptr(DUX, "f8", self.n2, "DUX"), _ptr(DVY, "f8", self.n2, "DVY"),
            _ptr(DZX, "f8", self.n2, "DZX"), _ptr(DZY, "f8", self.n2, "DZY"),
            _ptr(KFloorZ, "i4", self.n2, "KFloorZ"), _ptr(BoundaryPoints2D, "i4", self.n2, "BoundaryPoints2D")))

    def set_step(self, step: Dict[str, object], SmallDepths=None):
        """Per-time-step shared inputs (AD:1132-1141, 1386-1401)."""
        args = [_ptr(step[k], "f8", self.n3, k) for k in STEP_F64] + [_ptr(step[k], "i4", self.n3, k) for k in STEP_I32]
        args.append(_ptr(SmallDepths, "i4", self.n2, "SmallDepths"))
        self._check(self.lib.mohid_adt_set_step(C.byref(self.h), *args))

    def set_noflux(self, NoFluxU=None, NoFluxV=None, NoFluxW=None):
        """The optional NoFluxU/V/W dummies (AD:1146); all None = not present."""
        self._check(self.lib.mohid_adt_set_noflux(C.byref(self.h), _ptr(NoFluxU, "i4", self.n3, "NoFluxU"),
                                                  _ptr(NoFluxV, "i4", self.n3, "NoFluxV"), _ptr(NoFluxW, "i4", self.n3, "NoFluxW")))

    def set_premix(self, Density=None, WaterColumnZ=None, SmallDepthsLimit: float = 0.0):
        """FreeConvection / SmallDepthsMixing_Processes before each transport call (WP:13017-13074, 12939-13012)."""
        self._check(self.lib.mohid_adt_set_premix(C.byref(self.h), _ptr(Density, "f8", self.n3, "Density"),
                                                  _ptr(WaterColumnZ, "f8", self.n2, "WaterColumnZ"),
                                                  C.byref(C.c_double(SmallDepthsLimit))))

    def set_offsets(self, offsets):
        """Property%OffSet per property of the batch (WP:14724-14746); [] clears them."""
        arr = (C.c_double * max(1, len(offsets)))(*[float(x) for x in offsets])
        self._check(self.lib.mohid_adt_set_offsets(C.byref(self.h), C.byref(C.c_int(len(offsets))), arr))

    def set_limits(self, min_values, max_values):
        """SetLimitsProperty after every step (WP:20594-20720); None entries = no limit; two empty lists clear."""
        n = len(min_values)
        assert len(max_values) == n
        mo = (C.c_int * max(1, n))(*[int(v is not None) for v in min_values])
        xo = (C.c_int * max(1, n))(*[int(v is not None) for v in max_values])
        mv = (C.c_double * max(1, n))(*[float(v or 0.0) for v in min_values])
        xv = (C.c_double * max(1, n))(*[float(v or 0.0) for v in max_values])
        self._check(self.lib.mohid_adt_set_limits(C.byref(self.h), C.byref(C.c_int(n)), mo, mv, xo, xv))

    def get_limit_mass(self, prop_index: int):
        """(Mass_created, Mass_Destroid) accumulated for one property."""
        import numpy as np
        mc = np.zeros((self.K + 2, self.J + 2, self.ld)); md = np.zeros_like(mc)
        self._check(self.lib.mohid_adt_get_limit_mass(C.byref(self.h), C.byref(C.c_int(prop_index)),
                                                      mc.ctypes.data_as(C.POINTER(C.c_double)),
                                                      md.ctypes.data_as(C.POINTER(C.c_double))))
        return mc, md

    def get_small_depths(self):
        """Me%SmallDepths%ON as built by the library (int32, (J+2, ld))."""
        import numpy as np
        out = np.zeros((self.J + 2, self.ld), np.int32)
        self._check(self.lib.mohid_adt_get_small_depths(C.byref(self.h), out.ctypes.data_as(C.POINTER(C.c_int))))
        return out

    def set_discharges(self, prop_index: int, d: Dict[str, object]):
        """SetDischarges (AD:978-1034) for the property at position ``prop_index`` of the next batch.
        ``d`` holds the reference's argument names: DischFlow, DischConc, DischI, DischJ, DischK, DischKmin,
        DischKmax, DischVert, IgnoreDisch, DischnCells, ByPass, DischConcMF."""
        f8 = lambda k: np.ascontiguousarray(d[k], dtype=np.float64)
        i4 = lambda k: np.ascontiguousarray(d[k], dtype=np.int32)
        a = {k: f8(k) for k in ("DischFlow", "DischConc", "DischConcMF")}
        a.update({k: i4(k) for k in ("DischI", "DischJ", "DischK", "DischKmin", "DischKmax", "DischVert",
                                     "IgnoreDisch", "DischnCells", "ByPass")})
        nd, nc = len(a["DischnCells"]), len(a["DischFlow"])
        P = lambda k: C.c_void_p(a[k].ctypes.data)
        self._check(self.lib.mohid_adt_set_discharges(
            C.byref(self.h), C.byref(C.c_int(prop_index)), C.byref(C.c_int(nd)), C.byref(C.c_int(nc)),
            P("DischFlow"), P("DischConc"), P("DischI"), P("DischJ"), P("DischK"), P("DischKmin"), P("DischKmax"),
            P("DischVert"), P("IgnoreDisch"), P("DischnCells"), P("ByPass"), P("DischConcMF")))

    def unset_discharges(self):
        """UnSetDischarges (AD:1040-1095)."""
        self._check(self.lib.mohid_adt_unset_discharges(C.byref(self.h)))

    # ---- the batched transport step -------------------------------------------------
    def _params(self, params: Sequence[dict]):
        return (Params * len(params))(*[make_params(p) for p in params])

    def _ptr_array(self, arrs: Optional[Sequence], name: str):
        if arrs is None:
            return None
        n = len(arrs)
        return (C.c_void_p * n)(*[_ptr(a, "f8", self.n3, f"{name}[{i}]") for i, a in enumerate(arrs)])

    def advect_batch(self, props: Sequence, params: Sequence[dict], refs: Optional[Sequence] = None):
        """Advance all ``props`` (updated in place, like PROP in the reference) one transport step."""
        n = len(props)
        self._check(self.lib.mohid_adt_advect_batch(C.byref(self.h), C.byref(C.c_int(n)),
                                                    self._ptr_array(props, "prop"), self._ptr_array(refs, "ref"),
                                                    self._params(params)))
        self.nprop = n

    def upload(self, props: Sequence, refs: Optional[Sequence] = None):
        n = len(props)
        self._check(self.lib.mohid_adt_upload_props(C.byref(self.h), C.byref(C.c_int(n)),
                                                    self._ptr_array(props, "prop"), self._ptr_array(refs, "ref")))
        self.nprop = n

    def advect_device(self, params: Sequence[dict], nsteps: int = 1):
        """Advance the device-resident properties ``nsteps`` steps; no host<->device traffic."""
        self._check(self.lib.mohid_adt_advect_device(C.byref(self.h), C.byref(C.c_int(len(params))),
                                                     self._params(params), C.byref(C.c_int(int(nsteps)))))

    def download(self, props: Sequence):
        self._check(self.lib.mohid_adt_download_props(C.byref(self.h), C.byref(C.c_int(len(props))),
                                                      self._ptr_array(props, "prop")))

    def get_cell_fluxes(self, prop_index: int) -> Dict[str, np.ndarray]:
        """GetAdvFlux + GetDifFlux (AD:697-851) of a property advanced with ``CellFluxes = 1``."""
        names = ("AdvFluxX", "AdvFluxY", "AdvFluxZ", "DifFluxX", "DifFluxY", "DifFluxZ")
        out = {k: np.zeros((self.K + 2, self.J + 2, self.ld)) for k in names}
        self._check(self.lib.mohid_adt_get_cell_fluxes(C.byref(self.h), C.byref(C.c_int(prop_index)),
                                                       *[C.c_void_p(out[k].ctypes.data) for k in names]))
        return out

    # ---- box budgets (ModuleBoxDif) ---------------------------------------------------
    def set_boxes(self, Boxes3D, NumberOfBoxes3D: int):
        """Me%Boxes3D of ModuleBoxDif (int32, (K+2, J+2, ld)); values <= -55 = no box."""
        self._nboxes = int(NumberOfBoxes3D)
        self._check(self.lib.mohid_adt_set_boxes(C.byref(self.h), _ptr(Boxes3D, "i4", self.n3, "Boxes3D"),
                                                 C.byref(C.c_int(self._nboxes))))

    def box_fluxes(self, prop_index: int) -> np.ndarray:
        """BoxDifFluxes3D (BoxDif:2659-2776) of AdvFlux + DifFlux of a property advanced with CellFluxes = 1:
        out[IN, OUT] (C order of the Fortran (OUT, IN) matrix), shape (nb+1, nb+1)."""
        nb1 = self._nboxes + 1
        out = np.zeros((nb1, nb1))
        self._check(self.lib.mohid_adt_box_fluxes(C.byref(self.h), C.byref(C.c_int(prop_index)), out.ctypes.data_as(C.c_void_p)))
        return out

    # ---- time integration of the hydrodynamic fluxes (ModuleHydroIntegration) ----------
    def hydro_integration_reinit(self, VolumeZOld):
        self._check(self.lib.mohid_adt_hydro_integration_reinit(C.byref(self.h), _ptr(VolumeZOld, "f8", self.n3, "VolumeZOld")))

    def hydro_integration_step(self, WaterFluxX, WaterFluxY, ComputeFacesU, ComputeFacesV, Discharges=None):
        self._check(self.lib.mohid_adt_hydro_integration_step(
            C.byref(self.h), _ptr(WaterFluxX, "f8", self.n3, "WaterFluxX"), _ptr(WaterFluxY, "f8", self.n3, "WaterFluxY"),
            _ptr(Discharges, "f8", self.n3, "Discharges"), _ptr(ComputeFacesU, "i4", self.n3, "ComputeFacesU"),
            _ptr(ComputeFacesV, "i4", self.n3, "ComputeFacesV")))

    def hydro_integration_end(self, VolumeZ, WaterPoints3D, DT: float):
        self._check(self.lib.mohid_adt_hydro_integration_end(C.byref(self.h), _ptr(VolumeZ, "f8", self.n3, "VolumeZ"),
                                                             _ptr(WaterPoints3D, "i4", self.n3, "WaterPoints3D"),
                                                             C.byref(C.c_double(DT))))

    def step_input(self, which: int) -> np.ndarray:
        """Host copy of a device mirror of set_step (0..10 the fp64 arrays in argument order, 11..16 the masks)."""
        out = np.zeros((self.K + 2, self.J + 2, self.ld), np.float64 if which < 11 else np.int32)
        self._check(self.lib.mohid_adt_download_step_input(C.byref(self.h), C.byref(C.c_int(which)),
                                                           out.ctypes.data_as(C.c_void_p)))
        return out

    # ---- settling (ModuleFreeVerticalMovement) ----------------------------------------
    def free_vertical_movement(self, prop_index: int, Velocity, GridCellArea, *, DepositionProbability=None,
                               Deposition: bool = False, NonCohesive: bool = False, DepositionIntertidalZones: bool = False,
                               ImpExp_AdvV: float = 0.0, DTProp: float = 30.0, want_flux: bool = False):
        """FreeVerticalMovementIteration (FVM:1531-1650) on the device-resident property ``prop_index``; returns
        FreeConvFlux when ``want_flux``."""
        flux = np.zeros((self.K + 2, self.J + 2, self.ld)) if want_flux else None
        ci = lambda v: C.byref(C.c_int(int(v)))
        self._check(self.lib.mohid_adt_free_vertical_movement(
            C.byref(self.h), ci(prop_index), _ptr(Velocity, "f8", self.n3, "Velocity"),
            _ptr(GridCellArea, "f8", self.n2, "GridCellArea"), _ptr(DepositionProbability, "f8", self.n2, "DepositionProbability"),
            ci(Deposition), ci(NonCohesive), ci(DepositionIntertidalZones), C.byref(C.c_double(ImpExp_AdvV)),
            C.byref(C.c_double(DTProp)), flux.ctypes.data_as(C.c_void_p) if want_flux else None))
        return flux

    # ---- halo staging for the j-slab decomposition ----------------------------------
    def device_layout(self, n: int = 0):
        """(ld, nj, nk) of the device arrays: element (i, j, k) of property n at ptr[i + ld * (j + nj * k)].  ld is the
        DEVICE leading dimension (rows padded to 128 bytes), not the ld of the caller's arrays."""
        ptr, ld, nj, nk = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.mohid_adt_prop_device_ptr(C.byref(self.h), C.byref(C.c_int(n)), C.byref(ptr), C.byref(ld),
                                                       C.byref(nj), C.byref(nk)))
        return ld.value, nj.value, nk.value

    def pack_elems(self, nprop: int, width: int) -> int:
        """Doubles a pack_columns / unpack_columns buffer of `width` columns must hold."""
        ld, _, nk = self.device_layout()
        return nprop * nk * width * ld

    def pack_columns(self, nprop: int, j0: int, width: int, device_buffer):
        if device_buffer.numel() < self.pack_elems(nprop, width):
            raise ValueError(f"pack buffer holds {device_buffer.numel()} doubles, {self.pack_elems(nprop, width)} needed "
                             "(nprop * (K + 2) * width * device ld)")
        self._check(self.lib.mohid_adt_pack_columns(C.byref(self.h), C.byref(C.c_int(nprop)), C.byref(C.c_int(j0)),
                                                    C.byref(C.c_int(width)), C.c_void_p(device_buffer.data_ptr())))

    def unpack_columns(self, nprop: int, j0: int, width: int, device_buffer):
        if device_buffer.numel() < self.pack_elems(nprop, width):
            raise ValueError(f"pack buffer holds {device_buffer.numel()} doubles, {self.pack_elems(nprop, width)} needed")
        self._check(self.lib.mohid_adt_unpack_columns(C.byref(self.h), C.byref(C.c_int(nprop)), C.byref(C.c_int(j0)),
                                                      C.byref(C.c_int(width)), C.c_void_p(device_buffer.data_ptr())))

    def set_active_columns(self, j_begin: int, j_count: int):
        """Advance only local columns j_begin .. j_begin+j_count-1 (the owned columns of a j-slab)."""
        self._check(self.lib.mohid_adt_set_active_columns(C.byref(self.h), C.byref(C.c_int(j_begin)),
                                                          C.byref(C.c_int(j_count))))

    def set_overlap(self, ghost: int, comm_stream: int = 0):
        """Edge-first stepping + pack/unpack on `comm_stream`; ghost = 0 switches it off."""
        self._check(self.lib.mohid_adt_set_overlap(C.byref(self.h), C.byref(C.c_int(ghost)), C.c_void_p(comm_stream)))

    # ---- column windows (slab-wise hosts, generators of cases too large to hold twice) ----
    def set_step_columns(self, j0: int, arrays: Dict[str, object]):
        """Columns j0 .. j0+ncols-1 of the per-step inputs; arrays shaped (K+2, ncols, ld); missing names are skipped."""
        ncols = next(iter(arrays.values())).shape[1]
        n = (self.K + 2) * ncols * self.ld
        args = [_ptr(arrays.get(k), "f8", n, k) for k in STEP_F64] + [_ptr(arrays.get(k), "i4", n, k) for k in STEP_I32]
        self._check(self.lib.mohid_adt_set_step_columns(C.byref(self.h), C.byref(C.c_int(j0)), C.byref(C.c_int(ncols)), *args))

    def mark_step_resident(self, small_depths_present: bool = False):
        self._check(self.lib.mohid_adt_mark_step_resident(C.byref(self.h), C.byref(C.c_int(int(small_depths_present)))))

    def upload_columns(self, j0: int, props: Sequence, refs: Optional[Sequence] = None):
        nprop, ncols = len(props), props[0].shape[1]
        n = (self.K + 2) * ncols * self.ld
        pp = (C.c_void_p * nprop)(*[_ptr(a, "f8", n, "prop") for a in props])
        rp = (C.c_void_p * nprop)(*[_ptr(a, "f8", n, "ref") for a in refs]) if refs else None
        self._check(self.lib.mohid_adt_upload_props_columns(C.byref(self.h), C.byref(C.c_int(nprop)), pp, rp,
                                                            C.byref(C.c_int(j0)), C.byref(C.c_int(ncols))))
        self.nprop = max(self.nprop, nprop)

    def download_columns(self, j0: int, props: Sequence):
        nprop, ncols = len(props), props[0].shape[1]
        n = (self.K + 2) * ncols * self.ld
        pp = (C.c_void_p * nprop)(*[_ptr(a, "f8", n, "prop") for a in props])
        self._check(self.lib.mohid_adt_download_props_columns(C.byref(self.h), C.byref(C.c_int(nprop)), pp,
                                                              C.byref(C.c_int(j0)), C.byref(C.c_int(ncols))))

    def column_mass(self, nprop: int) -> np.ndarray:
        """(nprop, J+2) array: sum over i, k of PROP * VolumeZ on the water points of every column, in a summation order
        that does not depend on the decomposition."""
        out = np.zeros((nprop, self.J + 2))
        self._check(self.lib.mohid_adt_column_mass(C.byref(self.h), C.byref(C.c_int(nprop)), out.ctypes.data_as(C.c_void_p)))
        return out

    # ---- NCCL halo exchange inside the library (mohid_adt_comm_*) -------------------
    def comm_unique_id(self) -> bytes:
        """128-byte NCCL id (rank 0 obtains it; the host hands it to the other ranks)."""
        buf = C.create_string_buffer(128)
        self._check(self.lib.mohid_adt_comm_get_unique_id(buf, C.byref(C.c_int(128))))
        return buf.raw

    def comm_init(self, nranks: int, rank: int, unique_id: bytes, ghost: int = 2, overlap: bool = True):
        """Collective: one communicator over the j-slabs; needs set_active_columns first."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self.lib.mohid_adt_comm_init(C.byref(self.h), C.byref(C.c_int(nranks)), C.byref(C.c_int(rank)), buf,
                                                 C.byref(C.c_int(ghost)), C.byref(C.c_int(int(overlap)))))

    def exchange_halos(self, nprop: int):
        """Edge columns of properties 0..nprop-1 -> the neighbours' ghost columns (replaces HG:8479-8658)."""
        self._check(self.lib.mohid_adt_exchange_halos(C.byref(self.h), C.byref(C.c_int(nprop))))

    def comm_destroy(self):
        self._check(self.lib.mohid_adt_comm_destroy(C.byref(self.h)))

    def join_halo(self):
        """The compute stream waits for a halo exchange still running on the communication stream."""
        self._check(self.lib.mohid_adt_join_halo(C.byref(self.h)))

    def halo_buffer_elems(self, nprop: int, width: int) -> int:
        return nprop * (self.K + 2) * width * self.device_ld

    @property
    def device_ld(self) -> int:
        return ((self.I + 2 + 15) // 16) * 16

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.mohid_adt_set_stream(C.byref(self.h), C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self.lib.mohid_adt_synchronize(C.byref(self.h)))

    # ---- stand-alone column solve ---------------------------------------------------
    def solve_thomas_z(self, D, E, F, TI, res, water=None):
        """THOMASZ_NewType2 (MF:4026-4123) / SolveThomas_C (ModuleCuda.F90:103-111) on caller-supplied fields;
        ``res`` is updated in place."""
        self._check(self.lib.mohid_adt_solve_thomas_z(
            C.byref(self.h), _ptr(D, "f8", self.n3, "D"), _ptr(E, "f8", self.n3, "E"), _ptr(F, "f8", self.n3, "F"),
            _ptr(TI, "f8", self.n3, "TI"), _ptr(water, "i4", self.n3, "WaterPoints3D"), _ptr(res, "f8", self.n3, "Res")))

    # ---- diagnostics ----------------------------------------------------------------
    def counters(self) -> Dict[str, int]:
        v = (C.c_longlong * 4)()
        self._check(self.lib.mohid_adt_get_counters(C.byref(self.h), v, C.byref(C.c_int(4))))
        return dict(launches=v[0], zero_pivots=v[1], mask_violations=v[2], device_bytes=v[3])

    def kernel_time_ms(self):
        ms, n = C.c_double(0), C.c_int(0)
        self._check(self.lib.mohid_adt_kernel_time_ms(C.byref(self.h), C.byref(ms), C.byref(n)))
        return ms.value, n.value


# =======================================================================================
# Module-level mirror of the reference's procedural interface (integer instance IDs,
# STAT-style returns).  One call of AdvectionDiffusion() = the reference's per-property call.
# =======================================================================================
_instances: Dict[int, TransportStep] = {}
_next_id = 1


def StartAdvectionDiffusion(I: int, J: int, K: int, *, Vertical1D: bool = False, XZFlow: bool = False,
                            Docycle_method: int = 1, ld: Optional[int] = None, device: int = -1) -> int:
    """AD:400-533.  The reference takes Geometry/Map/Grid/Time object IDs; here the sizes they carry."""
    global _next_id
    obj = TransportStep(I, J, K, ld, vertical1d=Vertical1D, xzflow=XZFlow, docycle_method=Docycle_method, device=device)
    ident = _next_id
    _next_id += 1
    _instances[ident] = obj
    return ident


def _get(AdvectionDiffusionID: int) -> TransportStep:
    try:
        return _instances[AdvectionDiffusionID]
    except KeyError:
        raise AdtError(9, "AdvectionDiffusion - ModuleAdvectionDiffusion - instance not ready (IDLE_ERR_)")


def GetTransportStep(AdvectionDiffusionID: int) -> TransportStep:
    return _get(AdvectionDiffusionID)


def AdvectionDiffusion(AdvectionDiffusionID: int, PROP, schmidt_H, SchmidtCoef_V, SchmidtBackground_V, AdvMethodH,
                       TVDLimitationH, AdvMethodV, TVDLimitationV, Upwind2H, Upwind2V, VolumeRelMax,
                       AdvectionNudging, AdvectionNudgingCells, DTProp, ImpExp_AdvV, ImpExp_DifV, ImpExp_AdvXX,
                       ImpExp_AdvYY, ImpExp_DifH, NullDif, Wflux_X, Wflux_Y, Wflux_Z, VolumeZOld, VolumeZ,
                       OpenPoints3D, LandPoints3D, ComputeFacesU3D, ComputeFacesV3D, ComputeFacesW3D, Visc_H, Diff_V,
                       *, DWZ, DZZ, AreaU, AreaV, CellFluxes: bool = False, WaterPoints3D=None, ReferenceProp=None,
                       BoundaryCondition: Optional[int] = None, DecayTime: float = 0.0, SmallDepths=None,
                       NoAdvFlux: bool = False, NoDifFlux: bool = False, NoFluxU=None, NoFluxV=None, NoFluxW=None) -> int:
    """Same dummy arguments as AD:1108-1147 (DWZ, DZZ, AreaU, AreaV are what the reference fetches from
    ModuleGeometry inside the call, AD:1386-1401).  ``PROP`` is updated in place.  Returns STAT."""
    obj = _get(AdvectionDiffusionID)
    if AdvectionNudging:
        raise AdtError(21, "AdvectionNudging (AD:1989-2068) is not available on the GPU path")
    if WaterPoints3D is None:
        raise AdtError(20, "WaterPoints3D is required (THOMASZ_NewType2 reads it, MF:4086)")
    step = dict(Wflux_X=Wflux_X, Wflux_Y=Wflux_Y, Wflux_Z=Wflux_Z, VolumeZOld=VolumeZOld, VolumeZ=VolumeZ,
                Visc_H=Visc_H, Diff_V=Diff_V, DWZ=DWZ, DZZ=DZZ, AreaU=AreaU, AreaV=AreaV, OpenPoints3D=OpenPoints3D,
                LandPoints3D=LandPoints3D, WaterPoints3D=WaterPoints3D, ComputeFacesU3D=ComputeFacesU3D,
                ComputeFacesV3D=ComputeFacesV3D, ComputeFacesW3D=ComputeFacesW3D)
    obj.set_step(step, SmallDepths)
    obj.set_noflux(NoFluxU, NoFluxV, NoFluxW)
    p = dict(Schmidt_H=schmidt_H, SchmidtCoef_V=SchmidtCoef_V, SchmidtBackground_V=SchmidtBackground_V,
             AdvMethodH=AdvMethodH, TVDLimitationH=TVDLimitationH, AdvMethodV=AdvMethodV, TVDLimitationV=TVDLimitationV,
             Upwind2H=int(Upwind2H), Upwind2V=int(Upwind2V), VolumeRelMax=VolumeRelMax, DTProp=DTProp,
             ImpExp_AdvV=ImpExp_AdvV, ImpExp_DifV=ImpExp_DifV, ImpExp_AdvXX=ImpExp_AdvXX, ImpExp_AdvYY=ImpExp_AdvYY,
             ImpExp_DifH=ImpExp_DifH, NullDif=int(NullDif),
             BoundaryCondition=(BoundaryCondition if BoundaryCondition is not None else 0), DecayTime=DecayTime,
             NoAdvFlux=int(NoAdvFlux), NoDifFlux=int(NoDifFlux), CellFluxes=int(CellFluxes))
    obj.advect_batch([PROP], [p], [ReferenceProp] if ReferenceProp is not None else None)
    return SUCCESS_


def SetDischarges(AdvectionDiffusionID: int, DischFlow, DischConc, DischI, DischJ, DischK, DischKmin, DischKmax, DischVert,
                  DischNumber, IgnoreDisch, DischnCells, ByPass, DischConcMF) -> int:
    """AD:978-1034 (applies to the next AdvectionDiffusion call, i.e. property 0 of a batch of one)."""
    assert DischNumber == len(DischnCells)
    _get(AdvectionDiffusionID).set_discharges(0, dict(
        DischFlow=DischFlow, DischConc=DischConc, DischI=DischI, DischJ=DischJ, DischK=DischK, DischKmin=DischKmin,
        DischKmax=DischKmax, DischVert=DischVert, IgnoreDisch=IgnoreDisch, DischnCells=DischnCells, ByPass=ByPass,
        DischConcMF=DischConcMF))
    return SUCCESS_


def UnSetDischarges(AdvectionDiffusionID: int) -> int:
    """AD:1040-1095."""
    _get(AdvectionDiffusionID).unset_discharges()
    return SUCCESS_


def GetAdvFlux(AdvectionDiffusionID: int):
    """AD:697-776: (AdvFluxX, AdvFluxY, AdvFluxZ) of the last AdvectionDiffusion call made with CellFluxes."""
    f = _get(AdvectionDiffusionID).get_cell_fluxes(0)
    return f["AdvFluxX"], f["AdvFluxY"], f["AdvFluxZ"]


def GetDifFlux(AdvectionDiffusionID: int):
    """AD:780-851."""
    f = _get(AdvectionDiffusionID).get_cell_fluxes(0)
    return f["DifFluxX"], f["DifFluxY"], f["DifFluxZ"]


def SetGrid2D(AdvectionDiffusionID: int, DUX, DVY, DZX, DZY, KFloorZ, BoundaryPoints2D) -> int:
    """What AD:1353-1384 fetches from ModuleHorizontalGrid / ModuleGeometry / ModuleHorizontalMap."""
    _get(AdvectionDiffusionID).set_grid2d(DUX, DVY, DZX, DZY, KFloorZ, BoundaryPoints2D)
    return SUCCESS_


def GetBoundaryConditionList() -> Dict[str, int]:
    """AD:855-893."""
    return dict(MassConservation=MassConservation_, ImposedValue=ImposedValue_, NullGradient=NullGradient_,
                SubModel=SubModel_, Orlanski=Orlanski_, MassConservNullGrad=MassConservNullGrad_,
                CyclicBoundary=CyclicBoundary_)


def KillAdvectionDiffusion(AdvectionDiffusionID: int) -> int:
    """AD:5849-6010."""
    obj = _instances.pop(AdvectionDiffusionID, None)
    if obj is None:
        return 9
    obj.close()
    return SUCCESS_
