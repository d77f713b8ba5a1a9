// =====================================================================================
//  adt_api.cu -- C-ABI (include/mohid_adt.h) of the B200-native MOHID transport step.
//  Host side: handle registry, device mirrors of the interface arrays, parameter validation
//  (the reference's `stop` conditions become error codes), launches of the kernels in
//  adt_kernels.cuh.  There is no CPU fallback: every compute entry point needs a CUDA device.
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>      // types only: the library is dlopen'ed by mohid_adt_comm_init, never linked

#include "adt_kernels.cuh"
#include "adt_hsolve_kernel.cuh"
#include "adt_lean_kernel.cuh"
#include "adt_fused_kernel.cuh"

using namespace adt;

namespace {

struct Handle {
    int id = 0, dev = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t s_up = nullptr, s_down = nullptr;      // copy streams of the pipelined host path (advect_batch)
    void *stage[3] = {nullptr, nullptr, nullptr};       // one field in the caller's pitch per stream (short padded rows)
    // halo exchange: pack / unpack run on a communication stream after the step (ev_edges), the next step waits (ev_halo)
    cudaStream_t comm = nullptr;
    cudaEvent_t ev_edges = nullptr, ev_halo = nullptr;
    bool halo_pending = false;
    std::vector<cudaEvent_t> pipe_ev;
    mohid_adt_options opt{};
    int I = 0, J = 0, K = 0, ni = 0, nj = 0, nk = 0, ld_h = 0, ld = 0;
    long n2 = 0, n3 = 0;
    int sj = 0, sk = 0;                                 // element strides of j and k in the 3-D device arrays
    // In-place shift: every 3-D array has njp = nj + S columns per plane.  A property's logical column 0 sits at
    // physical column shift[n] (0 or S); a step reads the field where it is and writes the new one at the other
    // position, S columns aside, walking the columns in chunks of C (S = C + 3) toward the side the data leaves, so
    // that no chunk overwrites old columns a later chunk still reads (stencil reach 2).  One buffer per property.
    int njp = 0, S = 0, C = 0;
    int maxprop = 0;
    // 2-D
    double *DUX = nullptr, *DVY = nullptr, *DZX = nullptr, *DZY = nullptr, *rdx = nullptr, *rdy = nullptr;
    int *KFloorZ = nullptr, *Bnd = nullptr, *SmallDepths = nullptr;
    bool have_small = false;
    int *bnd_cols = nullptr;
    int *noflux[3] = {nullptr, nullptr, nullptr};       // NoFluxU/V/W mirrors (allocated by set_noflux)
    int *boxes = nullptr; int nboxes = -1;              // Boxes3D of ModuleBoxDif (mohid_adt_set_boxes)
    double *box_flux = nullptr;
    int hint_n = -1;                                    // ModuleHydroIntegration: steps integrated since the re-initialisation (-1 = idle)
    double *hint_disch = nullptr, *hint_in[3] = {nullptr, nullptr, nullptr};   // integrated discharges; staging of one step
    int *hint_cf[2] = {nullptr, nullptr};
    bool have_noflux = false;
    double *density = nullptr, *wcol = nullptr;         // caller-side pre-steps (mohid_adt_set_premix)
    bool premix_fc = false, premix_sd = false;
    double sd_limit = 0.;
    std::vector<double> offsets;                        // AddOffSet per property (mohid_adt_set_offsets)
    std::vector<int> lim_min_on, lim_max_on;            // SetLimitsProperty per property (mohid_adt_set_limits)
    std::vector<double> lim_min, lim_max;
    std::vector<double *> mass_created, mass_destroyed;
    std::vector<double *> wline, hs_tmp;                // horizontally implicit advection: W of the line recurrence, intermediate field
    std::vector<double *> old_copy;                     // field at time n (CellFluxes: the explicit shares use it)
    unsigned char *nfmask = nullptr;                    // NF_* bits, rebuilt by K1 every step
    int n_bnd_cols = 0;
    bool bnd_off_ring = false;                          // a boundary point away from the outer ring (Orlanski stops on it)
    // raw per-step inputs
    double *raw_d[11] = {nullptr};
    int *raw_i[6] = {nullptr};
    // packed per-step coefficients (K1 outputs; allocated when a batch first takes the round-1 kernels)
    double *dtv = nullptr, *vr = nullptr, *dhu = nullptr, *dhv = nullptr, *dvz = nullptr, *rdz = nullptr;
    uint32_t *mask = nullptr;
    // properties (one buffer each, see `shift` above) + reference fields
    std::vector<double *> prop;
    std::vector<int> shift;
    std::vector<double *> ref;
    std::vector<char> has_ref;
    unsigned long long *d_zero_piv = nullptr;
    bool have_grid = false, have_step = false;
    long long launches = 0, bytes = 0, zero_piv_last = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;   // K2 timing: one pair per chunk launch
    size_t ev_used = 0;
    int ev_steps = 0;                                   // steps (launch groups) the pairs belong to
    std::string err;
    int smem_optin = 0, num_sms = 0;
    int j_begin = 1, j_count = 0;                       // columns advanced by this handle (slab decomposition)
    // point discharges: shared geometry (per listed cell) + per-property concentrations
    int d_ncell = 0;
    int *d_ci = nullptr, *d_cj = nullptr, *d_ck = nullptr, *d_ckmin = nullptr, *d_ckmax = nullptr, *d_cvert = nullptr,
        *d_cbypass = nullptr, *d_kmin_eff = nullptr, *d_kmax_eff = nullptr;
    double *d_cflow = nullptr, *d_flow_k = nullptr;
    std::vector<double *> d_conc, d_concmf;             // per property (nullptr = no discharges)
    std::vector<double *> flux[6];                      // per property: AdvFluxX/Y/Z, DifFluxX/Y/Z (allocated on demand)
    std::vector<int> bnd_host;                          // (i,j) of all boundary columns
    // lean path (adt_lean_kernel.cuh): face packs U, V, W, C of the columns pk_jc0 .. pk_jc0+pk_ncol-1, 2-D metric ratios
    Pack4 *pk = nullptr;                               // layout: adt_lean_kernel.cuh (LeanCoefArgs)
    int pk_ncol = 0, pk_jc0 = 0, pk_nt32 = 0;
    double *rho2d[4] = {nullptr, nullptr, nullptr, nullptr};
    bool rho_valid = false, lean_now = false, fused_now = false;
    // NCCL halo exchange behind the C-ABI (mohid_adt_comm_init): communicator, neighbour buffers, own comm stream
    ncclComm_t nccl = nullptr;
    int nranks = 1, rank = 0, halo_ghost = 0;
    double *line_buf[4] = {nullptr, nullptr, nullptr, nullptr};   // split line solve: edge in / out, x in / out
    double *cyc_buf[2] = {nullptr, nullptr};                      // cyclic boundary across the slabs: send / receive
    size_t cyc_cap = 0;
    size_t line_cap = 0;
    cudaStream_t s_comm_own = nullptr;
    double *halo_buf[4] = {nullptr, nullptr, nullptr, nullptr};     // send left, recv left, send right, recv right
    size_t halo_cap = 0;
};

// ---- NCCL, loaded at run time (single-GPU users never need it) ----
struct NcclApi {
    void *dl = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

// a copy already loaded by the host process (torch ships its own) is reused; else the system library
const char *load_nccl() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.dl) return nullptr;
    void *dl = nullptr;
    const char *env = getenv("MOHID_ADT_NCCL_LIB");
    if (env) dl = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!dl) dl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!dl) dl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!dl) dl = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!dl) return "libnccl.so.2 not found (set MOHID_ADT_NCCL_LIB)";
    NcclApi a;
    a.dl = dl;
#define ADT_SYM(field, name)                                   \
    *(void **)(&a.field) = dlsym(dl, name);                    \
    if (!a.field) return "symbol " name " missing in libnccl";
    ADT_SYM(GetUniqueId, "ncclGetUniqueId") ADT_SYM(CommInitRank, "ncclCommInitRank") ADT_SYM(CommDestroy, "ncclCommDestroy")
    ADT_SYM(Send, "ncclSend") ADT_SYM(Recv, "ncclRecv") ADT_SYM(GroupStart, "ncclGroupStart")
    ADT_SYM(GroupEnd, "ncclGroupEnd") ADT_SYM(GetErrorString, "ncclGetErrorString")
#undef ADT_SYM
    g_nccl = a;
    return nullptr;
}

std::mutex g_mu;
std::map<int, Handle *> g_h;
int g_next = 1;
std::string g_err = "";

int fail(Handle *h, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    std::lock_guard<std::mutex> lk(g_mu);
    g_err = buf;
    return code;
}

#define CU(h, call)                                                                                     \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(h, MOHID_ADT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                            \
    } while (0)

Handle *get(const int *handle) {
    if (!handle) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_h.find(*handle);
    return it == g_h.end() ? nullptr : it->second;
}

template <typename T>
int dalloc(Handle *h, T **p, size_t n) {
    // 64 elements of slack: the staged row windows of adt_transport_ring_kernel run up to 36 elements past a row end
    CU(h, cudaMalloc((void **)p, (n + 64) * sizeof(T)));
    h->bytes += (long long)((n + 64) * sizeof(T));
    return 0;
}

// current / next position of property n inside its buffer
inline double *cur_ptr(const Handle *h, int n) { return h->prop[n] + (size_t)h->shift[n] * h->ld; }
inline double *nxt_ptr(const Handle *h, int n) { return h->prop[n] + (size_t)(h->S - h->shift[n]) * h->ld; }

// caller -> device mirror copy of a 2-D array (nj rows of ni elements, element size es).  cudaMemcpyDefault:
// the caller's arrays may live in host memory (the Fortran case) or, under UVA, in device memory.
int h2d2(Handle *h, void *dst, const void *src, size_t es) {
    CU(h, cudaMemcpy2DAsync(dst, es * h->ld, src, es * h->ld_h, es * std::min(h->ld, h->ld_h), h->nj,
                            cudaMemcpyDefault, h->stream));
    return 0;
}
// Staging buffer of the stream (one field in the caller's pitch), or nullptr when the rows are long enough for the
// copy engine to move them pitched at link speed.  When ld_h > ld the caller's padding columns travel too; the
// D2H direction then leaves them untouched on the host only in the direct path, so staging is limited to ld_h <= ld.
void *staging(Handle *h, cudaStream_t st, size_t es) {
    static const size_t row_limit = getenv("MOHID_ADT_STAGE_ROW") ? (size_t)atol(getenv("MOHID_ADT_STAGE_ROW")) : 8192;
    if (h->ld_h > h->ld || es * (size_t)h->ld_h >= row_limit) return nullptr;
    const int w = (st == h->s_up && h->s_up) ? 1 : (st == h->s_down && h->s_down) ? 2 : 0;
    if (!h->stage[w]) {
        if (cudaMalloc(&h->stage[w], 8 * (size_t)h->ld_h * h->nj * h->nk) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        h->bytes += 8LL * h->ld_h * h->nj * h->nk;
    }
    return h->stage[w];
}
// 3-D arrays: caller (ld_h, ncols, nk) <-> device (ld, njp, nk), one 3-D copy; `dev` points at the logical column 0
// of the field, the caller's array holds the columns j0 .. j0+ncols-1 (the whole field: j0 = 0, ncols = nj).  Short
// caller rows cross the link contiguously and change pitch on the device.
int copy3(Handle *h, void *dev, const void *host, size_t es, bool to_dev, cudaStream_t st, int j0 = 0, int ncols = -1) {
    if (!st) st = h->stream;
    if (ncols < 0) ncols = h->nj;
    const size_t w = es * (size_t)std::min(h->ld, h->ld_h);
    cudaMemcpy3DParms p{};
    p.extent = make_cudaExtent(w, (size_t)ncols, (size_t)h->nk);
    p.kind = cudaMemcpyDefault;
    const cudaPitchedPtr D = make_cudaPitchedPtr((char *)dev + es * (size_t)h->ld * j0, es * (size_t)h->ld, es * (size_t)h->ld, (size_t)h->njp);
    cudaPitchedPtr H = make_cudaPitchedPtr(const_cast<void *>(host), es * (size_t)h->ld_h, es * (size_t)h->ld_h, (size_t)ncols);
    void *stg = (ncols == h->nj) ? staging(h, st, es) : nullptr;
    const size_t nbytes = es * (size_t)h->ld_h * ncols * h->nk;
    if (stg && to_dev) {
        CU(h, cudaMemcpyAsync(stg, host, nbytes, cudaMemcpyDefault, st));
        H.ptr = stg;
    }
    if (stg && !to_dev) H.ptr = stg;
    p.srcPtr = to_dev ? H : D;
    p.dstPtr = to_dev ? D : H;
    CU(h, cudaMemcpy3DAsync(&p, st));
    if (stg && !to_dev) CU(h, cudaMemcpyAsync(const_cast<void *>(host), stg, nbytes, cudaMemcpyDefault, st));
    return 0;
}
int h2d3(Handle *h, void *dst, const void *src, size_t es, cudaStream_t st = nullptr) { return copy3(h, dst, src, es, true, st); }
int d2h3(Handle *h, void *dst, const void *src, size_t es, cudaStream_t st = nullptr) {
    return copy3(h, const_cast<void *>(src), dst, es, false, st);
}

int ensure_props(Handle *h, int nprop, bool need_ref_any) {
    if (nprop > NPMAX) return fail(h, MOHID_ADT_ERR_ARG, "nprop %d exceeds the batch limit %d", nprop, NPMAX);
    while ((int)h->prop.size() < nprop) {
        double *a = nullptr;
        if (int rc = dalloc(h, &a, h->n3)) return rc;
        CU(h, cudaMemsetAsync(a, 0, h->n3 * sizeof(double), h->stream));      // padding and shift margin read as zero
        h->prop.push_back(a);
        h->shift.push_back(h->S);
        h->ref.push_back(nullptr);
        h->has_ref.push_back(0);
    }
    (void)need_ref_any;
    return 0;
}

// packed coefficient arrays of the round-1 kernels (52 B per cell), allocated on first use
int ensure_legacy(Handle *h) {
    if (h->mask) return 0;
    int rc = 0;
    rc |= dalloc(h, &h->dtv, h->n3); rc |= dalloc(h, &h->vr, h->n3); rc |= dalloc(h, &h->dhu, h->n3);
    rc |= dalloc(h, &h->dhv, h->n3); rc |= dalloc(h, &h->dvz, h->n3); rc |= dalloc(h, &h->rdz, h->n3);
    rc |= dalloc(h, &h->mask, h->n3);
    return rc;
}

void free_all(Handle *h) {
    cudaSetDevice(h->dev);
    auto F = [](void *p) { if (p) cudaFree(p); };
    F(h->DUX); F(h->DVY); F(h->DZX); F(h->DZY); F(h->rdx); F(h->rdy); F(h->KFloorZ); F(h->Bnd); F(h->SmallDepths);
    F(h->bnd_cols);
    for (auto p : h->noflux) F(p);
    F(h->nfmask); F(h->boxes); F(h->box_flux); F(h->hint_disch);
    for (auto p : h->hint_in) F(p);
    for (auto p : h->hint_cf) F(p);
    for (auto p : h->wline) F(p);
    for (auto p : h->hs_tmp) F(p);
    for (auto p : h->old_copy) F(p);
    F(h->density); F(h->wcol);
    for (auto p : h->mass_created) F(p);
    for (auto p : h->mass_destroyed) F(p);
    for (auto p : h->raw_d) F(p);
    for (auto p : h->raw_i) F(p);
    F(h->dtv); F(h->vr); F(h->dhu); F(h->dhv); F(h->dvz); F(h->rdz); F(h->mask);
    for (auto p : h->prop) F(p);
    for (auto p : h->ref) F(p);
    F(h->d_zero_piv);
    F(h->pk);
    for (auto p : h->rho2d) F(p);
    for (auto p : h->halo_buf) F(p);
    for (auto p : h->line_buf) F(p);
    for (auto p : h->cyc_buf) F(p);
    if (h->nccl && g_nccl.CommDestroy) { g_nccl.CommDestroy(h->nccl); h->nccl = nullptr; }
    if (h->s_comm_own) { cudaStreamDestroy(h->s_comm_own); h->s_comm_own = nullptr; }
    F(h->d_ci); F(h->d_cj); F(h->d_ck); F(h->d_ckmin); F(h->d_ckmax); F(h->d_cvert); F(h->d_cbypass);
    F(h->d_kmin_eff); F(h->d_kmax_eff); F(h->d_cflow); F(h->d_flow_k);
    for (auto p : h->d_conc) F(p);
    for (auto p : h->d_concmf) F(p);
    for (auto &v : h->flux) for (auto p : v) F(p);
    for (auto &e : h->ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (h->ev_edges) cudaEventDestroy(h->ev_edges);
    if (h->ev_halo) cudaEventDestroy(h->ev_halo);
    h->ev_edges = h->ev_halo = nullptr;
    for (auto &p : h->stage) { if (p) cudaFree(p); p = nullptr; }
    if (h->s_up) cudaStreamDestroy(h->s_up);
    if (h->s_down) cudaStreamDestroy(h->s_down);
    for (auto e : h->pipe_ev) cudaEventDestroy(e);
    h->s_up = h->s_down = nullptr; h->pipe_ev.clear();
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
}

// boundary columns inside the active j-range -> device list used by the post-solve passes
int upload_bnd_cols(Handle *h) {
    std::vector<int> cols;
    h->bnd_off_ring = false;
    for (size_t c = 0; c + 1 < h->bnd_host.size(); c += 2) {
        const int j = h->bnd_host[c + 1], i = h->bnd_host[c];
        if (!(i == 1 || i == h->I || j == 1 || j == h->J)) h->bnd_off_ring = true;
        if (j >= h->j_begin && j < h->j_begin + h->j_count) { cols.push_back(h->bnd_host[c]); cols.push_back(j); }
    }
    if (h->bnd_cols) { cudaFree(h->bnd_cols); h->bnd_cols = nullptr; }
    h->n_bnd_cols = (int)(cols.size() / 2);
    if (h->n_bnd_cols) {
        if (int r = dalloc(h, &h->bnd_cols, cols.size())) return r;
        CU(h, cudaMemcpy(h->bnd_cols, cols.data(), cols.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    return 0;
}

// What the reference's module state makes of a property's own flags.  AdvectionDiffusion is called once per
// property and Set_Internal_State (AD:5746-5835) decides per call which coefficient arrays are rebuilt; arrays that
// are not rebuilt keep what the PREVIOUS property of the time step left in them, including its NullDif / NoDifFlux /
// NoAdvFlux zeroing.  The batch reproduces that by resolving, in caller order, which flags each property really sees.
struct PropEff {
    double schmidt_h = 0., coef_v = 0., bg_v = 0.;   // Schmidt numbers behind DifX/DifY and DifZ
    int nulldif_h = 0, nulldif_v = 0;                // NullDif zeroing present in DifX/DifY, in DifZ
    int nodif_h = 0;                                 // NoDifFlux zeroing present in DifX/DifY (AD:2497-2501, 2524-2528)
    int nodif_w = 0;                                 // NoDifFlux on AuxK: always the property's own (AD:2737-2741)
    unsigned nfsel = 0;                              // NF_* bits of the advective coefficients the property uses
    bool same_dif(const PropEff &o) const {
        return schmidt_h == o.schmidt_h && coef_v == o.coef_v && bg_v == o.bg_v && nulldif_h == o.nulldif_h &&
               nulldif_v == o.nulldif_v && nodif_h == o.nodif_h && nodif_w == o.nodif_w;
    }
};

struct Batch {          // validated view of one advect call
    int nprop = 0;
    bool optimize = false;
    const mohid_adt_params *p = nullptr;
    std::vector<PropEff> eff;
};

// Set_Internal_State (AD:5746-5835) + the rebuild conditions of AD:1423-1456, 4280-4300, replayed over the batch.
// The first property of a batch sees LastCalc /= Now (a batch is one time step), so everything is rebuilt for it.
void resolve_effective_flags(const Handle *h, Batch &b) {
    b.eff.assign(b.nprop, PropEff{});
    bool xz_u = false, xz_v = false, vz_w = false;      // zeroed cell lists sitting in COEF3_HorAdvXX / COEF3_VertAdv
    for (int n = 0; n < b.nprop; ++n) {
        const mohid_adt_params &q = b.p[n];
        PropEff &e = b.eff[n];
        const bool first = n == 0;
        const mohid_adt_params &pv = b.p[first ? 0 : n - 1];
        // State%HorDif / VertDif ON -> Convert_* runs (non-Optimize: every time; Optimize: first property only)
        const bool hdif_on = first || ((pv.Schmidt_H != q.Schmidt_H || q.NullDif) && !b.optimize);
        const bool vdif_on = first || ((pv.SchmidtCoef_V != q.SchmidtCoef_V ||
                                        pv.SchmidtBackground_V != q.SchmidtBackground_V || q.NullDif) && !b.optimize);
        if (hdif_on) { e.schmidt_h = q.Schmidt_H; e.nulldif_h = q.NullDif; e.nodif_h = q.NoDifFlux && h->have_noflux; }
        else { e.schmidt_h = b.eff[n - 1].schmidt_h; e.nulldif_h = b.eff[n - 1].nulldif_h; e.nodif_h = b.eff[n - 1].nodif_h; }
        if (vdif_on) { e.coef_v = q.SchmidtCoef_V; e.bg_v = q.SchmidtBackground_V; e.nulldif_v = q.NullDif; }
        else { e.coef_v = b.eff[n - 1].coef_v; e.bg_v = b.eff[n - 1].bg_v; e.nulldif_v = b.eff[n - 1].nulldif_v; }
        e.nodif_w = q.NoDifFlux && h->have_noflux;
        // State%HorAdv / VertAdv: methods are uniform over the batch, so only P2_TVD coefficients are rebuilt
        const bool noadv = q.NoAdvFlux && h->have_noflux;
        if (first || q.AdvMethodH == MOHID_P2_TVD) {
            e.nfsel |= noadv ? (NF_UW | NF_UE) : 0u;                    // XX pass (AD:4434-4443)
            xz_u = noadv;
            xz_v = noadv && !h->opt.XZFlow;                             // YY pass, after the XX use (AD:4804-4813)
        } else {
            e.nfsel |= (xz_u ? (NF_UW | NF_UE) : 0u) | (xz_v ? (NF_VW | NF_VE) : 0u);
        }
        if (first || q.AdvMethodV == MOHID_P2_TVD) vz_w = noadv;        // AD:3004-3013
        e.nfsel |= vz_w ? NF_WT : 0u;
    }
}

// The reference's argument checks (AD:1229-1237, 1340-1349, 3124, 4525; MF:10869-10873) plus the
// limits of the GPU path.
int validate(Handle *h, int nprop, const mohid_adt_params *p, Batch &b) {
    if (nprop < 1) return fail(h, MOHID_ADT_ERR_ARG, "nprop must be >= 1");
    if (nprop > NPMAX) return fail(h, MOHID_ADT_ERR_ARG, "nprop %d exceeds the batch limit %d", nprop, NPMAX);
    const mohid_adt_params &f = p[0];
    bool optimize = nprop >= 2 && !h->opt.Vertical1D;
    for (int n = 0; n < nprop; ++n) {
        const mohid_adt_params &q = p[n];
        if (q.ImpExp_DifH != 0.0)
            return fail(h, MOHID_ADT_ERR_ARG, "AdvectionDiffusion - ModuleAdvectionDiffusion - ERR02 (horizontal diffusion must be explicit)");
        if (q.ImpExp_AdvXX == 1.0 && q.ImpExp_AdvYY == 1.0)
            return fail(h, MOHID_ADT_ERR_ARG, "AdvectionDiffusion - ModuleAdvectionDiffusion - ERR03 (both horizontal directions implicit)");
        if ((q.ImpExp_AdvXX == 1.0 || q.ImpExp_AdvYY == 1.0) &&
            (q.AdvMethodH == MOHID_UpwindOrder2 || q.AdvMethodH == MOHID_UpwindOrder3))
            return fail(h, MOHID_ADT_ERR_ARG, "AdvectionDiffusion - ModuleAdvectionDiffusion - ERR100");
        if (q.ImpExp_AdvV == 1.0 && (q.AdvMethodV == MOHID_UpwindOrder2 || q.AdvMethodV == MOHID_UpwindOrder3))
            return fail(h, MOHID_ADT_ERR_ARG, "AdvectionDiffusion - ModuleAdvectionDiffusion - ERR200");
        if (q.ImpExp_AdvXX != 0.0 && q.ImpExp_AdvXX != 1.0)
            return fail(h, MOHID_ADT_ERR_ARG, "sub. HorizontalAdvectionXX - ModuleAdvectionDiffusion - ERR01");
        if (q.ImpExp_AdvYY != 0.0 && q.ImpExp_AdvYY != 1.0)
            return fail(h, MOHID_ADT_ERR_ARG, "sub. HorizontalAdvectionYY - ModuleAdvectionDiffusion - ERR01");
        if (q.ImpExp_AdvV != 0.0 && q.ImpExp_AdvV != 1.0)
            return fail(h, MOHID_ADT_ERR_ARG, "sub. VerticalAdvection - ModuleAdvectionDiffusion - ERR01");
        if ((q.ImpExp_AdvXX == 1.0 || q.ImpExp_AdvYY == 1.0) && !h->opt.Vertical1D) {
            // lines along j (ImpExp_AdvXX) cross the slabs of a decomposed domain (THOMAS_DDecompHorizGrid, HG:8245-8478);
            // lines along i (ImpExp_AdvYY) lie inside one slab and are solved locally
            if (q.ImpExp_AdvXX == 1.0 && (h->j_begin != 1 || h->j_count != h->J) && !h->nccl)
                return fail(h, MOHID_ADT_ERR_STATE,
                            "implicit advection along j couples the columns of all slabs (AD:4200-4244): call mohid_adt_comm_init first");
        }
        for (int m : {q.AdvMethodH, q.AdvMethodV})
            if (m < MOHID_UpwindOrder1 || m > MOHID_LeapFrog)
                return fail(h, MOHID_ADT_ERR_ARG, "This method is not valid to compute Advection1D");
        if (q.AdvMethodH == MOHID_P2_TVD && (q.TVDLimitationH < MOHID_MinMod || q.TVDLimitationH > MOHID_PDM))
            return fail(h, MOHID_ADT_ERR_ARG, "This TVD Limitation option is not valid to compute Advection1D");
        if (q.AdvMethodV == MOHID_P2_TVD && (q.TVDLimitationV < MOHID_MinMod || q.TVDLimitationV > MOHID_PDM))
            return fail(h, MOHID_ADT_ERR_ARG, "This TVD Limitation option is not valid to compute Advection1D");
        // near-boundary faces of methods 2/3/4 stop the reference unless Upwind2 is set (MF:10869-10873)
        if ((q.AdvMethodH >= MOHID_UpwindOrder2 && q.AdvMethodH <= MOHID_P2_TVD && !q.Upwind2H) ||
            (q.AdvMethodV >= MOHID_UpwindOrder2 && q.AdvMethodV <= MOHID_P2_TVD && !q.Upwind2V))
            return fail(h, MOHID_ADT_ERR_ARG, "This method is not valid to compute Advection1D (Upwind2 must be set, WP:9638-9652)");
        const int bc = q.BoundaryCondition;
        if (bc == MOHID_BC_Orlanski) {
            if (h->bnd_off_ring)
                return fail(h, MOHID_ADT_ERR_ARG, "Orlanski Advection 2 (a boundary point is not on the outer ring, AD:5518)");
        }
        if (bc == MOHID_BC_CyclicBoundary && (h->j_begin != 1 || h->j_count != h->J) && !h->nccl)
            return fail(h, MOHID_ADT_ERR_STATE,
                        "CyclicBoundary wraps the global columns 1 and J (AD:2121-2224): on a column slab call mohid_adt_comm_init first");
        if (bc != MOHID_BC_None && bc != MOHID_BC_MassConservation && bc != MOHID_BC_ImposedValue &&
            bc != MOHID_BC_NullGradient && bc != MOHID_BC_SubModel && bc != MOHID_BC_MassConservNullGrad &&
            bc != MOHID_BC_CyclicBoundary && bc != MOHID_BC_Orlanski)
            return fail(h, MOHID_ADT_ERR_ARG, "Set_Internal_State - ModuleAdvectionDiffusion - ERR01");
        if ((q.NoAdvFlux || q.NoDifFlux) && !h->have_noflux)
            return fail(h, MOHID_ADT_ERR_ARG, "NoAdvFlux / NoDifFlux need the NoFluxU/V/W arrays (mohid_adt_set_noflux)");
        if (!(q.DTProp > 0.0)) return fail(h, MOHID_ADT_ERR_ARG, "DTProp must be positive");
        // one kernel variant per batch: these keywords are read FromFile in the reference (WP:9580-9632)
        if (q.AdvMethodH != f.AdvMethodH || q.AdvMethodV != f.AdvMethodV || q.TVDLimitationH != f.TVDLimitationH ||
            q.TVDLimitationV != f.TVDLimitationV || q.VolumeRelMax != f.VolumeRelMax || q.DTProp != f.DTProp ||
            q.Upwind2H != f.Upwind2H || q.Upwind2V != f.Upwind2V)
            return fail(h, MOHID_ADT_ERR_ARG,
                        "properties of one batch must share DTProp, advection methods, limiters and VolumeRelMax");
        // OptimizeFlag (WP:14580-14598)
        if (q.Schmidt_H != f.Schmidt_H || q.NullDif || q.AdvMethodH != MOHID_P2_TVD || q.AdvMethodV != MOHID_P2_TVD ||
            q.TVDLimitationH != MOHID_SuperBee || q.TVDLimitationV != MOHID_SuperBee || q.NoAdvFlux || q.NoDifFlux)
            optimize = false;
    }
    // Optimize is an argument of the reference call (AD:1146), computed by the caller over ALL coupled properties of the
    // model (WP:14580-14598): a batch that holds only some of them passes it explicitly (1 = on, 2 = off, 0 = inferred)
    if (f.Optimize == 1) optimize = !h->opt.Vertical1D;
    else if (f.Optimize == 2) optimize = false;
    b.nprop = nprop; b.optimize = optimize; b.p = p;
    resolve_effective_flags(h, b);
    return 0;
}

int launch_coef(Handle *h, const mohid_adt_params &q, const PropEff &e, bool geom, bool diff) {
    CoefArgs a{};
    a.ni = h->ni; a.nj = h->nj; a.nk = h->nk; a.ld = h->ld; a.I = h->I; a.J = h->J; a.K = h->K;
    a.sj = h->sj; a.sk = h->sk;
    a.dt = q.DTProp; a.schmidt_h = e.schmidt_h; a.schmidt_coef_v = e.coef_v; a.schmidt_bg_v = e.bg_v;
    a.nulldif = e.nulldif_h; a.nulldif_v = e.nulldif_v; a.nodif_h = e.nodif_h; a.nodif_w = e.nodif_w;
    a.Wflux_X = h->raw_d[0]; a.Wflux_Y = h->raw_d[1]; a.Wflux_Z = h->raw_d[2]; a.VolumeZOld = h->raw_d[3];
    a.VolumeZ = h->raw_d[4]; a.Visc_H = h->raw_d[5]; a.Diff_V = h->raw_d[6]; a.DWZ = h->raw_d[7]; a.DZZ = h->raw_d[8];
    a.AreaU = h->raw_d[9]; a.AreaV = h->raw_d[10];
    a.Open = h->raw_i[0]; a.Land = h->raw_i[1]; a.Water = h->raw_i[2]; a.CFU = h->raw_i[3]; a.CFV = h->raw_i[4];
    a.CFW = h->raw_i[5]; a.SmallDepths = h->have_small ? h->SmallDepths : nullptr;
    a.DUX = h->DUX; a.DVY = h->DVY; a.DZX = h->DZX; a.DZY = h->DZY; a.Bnd = h->Bnd;
    a.dtv = h->dtv; a.vr = h->vr; a.dhu = h->dhu; a.dhv = h->dhv; a.dvz = h->dvz; a.rdz = h->rdz; a.mask = h->mask;
    a.NoFluxU = h->have_noflux ? h->noflux[0] : nullptr; a.NoFluxV = h->have_noflux ? h->noflux[1] : nullptr;
    a.NoFluxW = h->have_noflux ? h->noflux[2] : nullptr; a.nfmask = h->nfmask;
    a.do_geom = geom; a.do_diff = diff;
    const dim3 grid((unsigned)((h->ld + 127) / 128), (unsigned)h->nk, (unsigned)h->nj);
    adt_coef_kernel<<<grid, 128, 0, h->stream>>>(a);
    CU(h, cudaGetLastError());
    h->launches++;
    return 0;
}

// ---- lean path (adt_lean_kernel.cuh) ----
// The batch takes the lean kernels when every property runs the headline form: 3-D, both horizontal directions
// explicit, implicit vertical advection, P2_TVD + SuperBee or first-order upwind in both directions, and none of the
// rare options (discharges, NoFlux cell lists, Orlanski, CellFluxes).
bool lean_eligible(const Handle *h, const Batch &b) {
    if (getenv("MOHID_ADT_NOLEAN")) return false;
    // the packs cost 128 B per cell to write and to read: they pay once they are shared by a few properties
    if (b.nprop < 3 && !getenv("MOHID_ADT_LEAN_ALWAYS")) return false;
    if (h->opt.Vertical1D || h->opt.XZFlow || h->K < 2 || h->have_noflux) return false;
    const mohid_adt_params &f = b.p[0];
    const bool tvd_sb = f.AdvMethodH == MOHID_P2_TVD && f.AdvMethodV == MOHID_P2_TVD &&
                        f.TVDLimitationH == MOHID_SuperBee && f.TVDLimitationV == MOHID_SuperBee;
    const bool upw = f.AdvMethodH == MOHID_UpwindOrder1 && f.AdvMethodV == MOHID_UpwindOrder1;
    if (!tvd_sb && !upw) return false;
    for (int n = 0; n < b.nprop; ++n) {
        const mohid_adt_params &q = b.p[n];
        if (q.ImpExp_AdvV != 1.0 || q.ImpExp_AdvXX == 1.0 || q.ImpExp_AdvYY == 1.0 || q.CellFluxes) return false;
        if (q.BoundaryCondition == MOHID_BC_Orlanski && h->has_ref[n]) return false;
        if (h->d_ncell > 0 && n < (int)h->d_conc.size() && h->d_conc[n]) return false;
        if (b.eff[n].nfsel || b.eff[n].nodif_h || b.eff[n].nodif_w) return false;
    }
    return (size_t)h->K * 32 * sizeof(double) * 8 <= (size_t)h->smem_optin;
}

// The fused kernel (adt_fused_kernel.cuh) serves the same configurations as the lean pair -- for any number of
// properties, since nothing is written for later re-use -- as long as W of FUSED_MAXP columns and the ring fit.
constexpr int FUSED_MAXP = 12;              // property warps per block (+ FR_NCW coefficient warps = 16 warps at 128 registers)
size_t fused_smem(int K, int nc) { return sizeof(double) * fused_smem_doubles(K, nc); }
// most property warps one block can take: bounded by the shared memory W of the column solve and the staging need
int fused_max_props(const Handle *h) {
    int nc = FUSED_MAXP;
    if (const char *e = getenv("MOHID_ADT_FUSED_MAXP")) nc = std::max(1, std::min(FUSED_MAXP, atoi(e)));
    while (nc > 0 && fused_smem(h->K, nc) > (size_t)h->smem_optin) --nc;
    return nc;
}
bool fused_eligible(const Handle *h, const Batch &b) {
    if (getenv("MOHID_ADT_NOFUSED")) return false;
    // measured on C2 (1 property): four coefficient warps per property warp cost more than the separate coefficient
    // pass of the round-1 kernels (0.50 against 0.33 ms per step); from 3 properties on the fused form wins
    return lean_eligible(h, b) && fused_max_props(h) >= 1;
}

int ensure_rho(Handle *h) {
    for (auto &p : h->rho2d) if (!p) { if (int rc = dalloc(h, &p, h->n2)) return rc; h->rho_valid = false; }
    if (!h->rho_valid) {
        adt_grid2d_rho_kernel<<<std::max(1, (int)std::min<long>((h->n2 + 255) / 256, 4096)), 256, 0, h->stream>>>(
            h->ni, h->nj, h->ld, h->I, h->J, h->rdx, h->rdy, h->rho2d[0], h->rho2d[1], h->rho2d[2], h->rho2d[3]);
        CU(h, cudaGetLastError());
        h->launches++;
        h->rho_valid = true;
    }
    return 0;
}

int ensure_lean(Handle *h, int ncol) {
    if (int rc = ensure_rho(h)) return rc;
    if (h->pk_ncol < ncol) {
        CU(h, cudaStreamSynchronize(h->stream));
        if (h->pk) { cudaFree(h->pk); h->pk = nullptr; }
        h->pk_nt32 = (h->ni + 31) / 32;
        if (int rc = dalloc(h, &h->pk, (size_t)h->pk_nt32 * 128 * ncol * h->nk)) return rc;
        h->pk_ncol = ncol;
    }
    return 0;
}

// raw inputs, metrics and Schmidt numbers of the coefficient work (lean pass and fused kernel) for the diffusion flags `e`
void fill_lean_coef_args(const Handle *h, const mohid_adt_params &q, const PropEff &e, LeanCoefArgs &A) {
    CoefArgs &a = A.c;
    a.ni = h->ni; a.nj = h->nj; a.nk = h->nk; a.ld = h->ld; a.I = h->I; a.J = h->J; a.K = h->K;
    a.sj = h->sj; a.sk = h->sk;
    a.dt = q.DTProp; a.schmidt_h = e.schmidt_h; a.schmidt_coef_v = e.coef_v; a.schmidt_bg_v = e.bg_v;
    a.nulldif = e.nulldif_h; a.nulldif_v = e.nulldif_v;
    a.Wflux_X = h->raw_d[0]; a.Wflux_Y = h->raw_d[1]; a.Wflux_Z = h->raw_d[2]; a.VolumeZOld = h->raw_d[3];
    a.VolumeZ = h->raw_d[4]; a.Visc_H = h->raw_d[5]; a.Diff_V = h->raw_d[6]; a.DWZ = h->raw_d[7]; a.DZZ = h->raw_d[8];
    a.AreaU = h->raw_d[9]; a.AreaV = h->raw_d[10];
    a.Open = h->raw_i[0]; a.Land = h->raw_i[1]; a.Water = h->raw_i[2]; a.CFU = h->raw_i[3]; a.CFV = h->raw_i[4];
    a.CFW = h->raw_i[5]; a.SmallDepths = h->have_small ? h->SmallDepths : nullptr;
    a.DUX = h->DUX; a.DVY = h->DVY; a.DZX = h->DZX; a.DZY = h->DZY; a.Bnd = h->Bnd;
    A.tvd = q.AdvMethodH == MOHID_P2_TVD; A.upwind2_h = q.Upwind2H; A.upwind2_v = q.Upwind2V;
    A.rhoUp = h->rho2d[0]; A.rhoUn = h->rho2d[1]; A.rhoVp = h->rho2d[2]; A.rhoVn = h->rho2d[3];
}

// packs of the columns jc0 .. jc0+ncol-1 for the diffusion flags `e`
int launch_lean_coef(Handle *h, const mohid_adt_params &q, const PropEff &e, int jc0, int ncol) {
    LeanCoefArgs A{};
    fill_lean_coef_args(h, q, e, A);
    A.pk = h->pk; A.jc0 = jc0; A.ncol = h->pk_ncol; A.nt32 = h->pk_nt32;
    const dim3 grid((unsigned)((h->pk_nt32 * 32 + 127) / 128), (unsigned)h->nk, (unsigned)ncol);
    const int minb = getenv("MOHID_ADT_COEF_MINB") ? atoi(getenv("MOHID_ADT_COEF_MINB")) : 4;
    if (minb >= 8) adt_lean_coef_kernel<8><<<grid, 128, 0, h->stream>>>(A);
    else if (minb >= 6) adt_lean_coef_kernel<6><<<grid, 128, 0, h->stream>>>(A);
    else adt_lean_coef_kernel<4><<<grid, 128, 0, h->stream>>>(A);
    CU(h, cudaGetLastError());
    h->launches++;
    h->pk_jc0 = jc0;
    return 0;
}

// Caller-side column steps (K7) for the properties `idx`: sign = +1 before the transport call (mixing + OffSet),
// -1 after it (OffSet taken out again).
int launch_premix(Handle *h, const std::vector<int> &idx, int sign) {
    bool any_off = false;
    for (int n : idx) any_off = any_off || (n < (int)h->offsets.size() && h->offsets[n] != 0.);
    if (sign > 0 ? !(h->premix_fc || h->premix_sd || any_off) : !any_off) return 0;
    PremixArgs a{};
    a.I = h->I; a.J = h->J; a.K = h->K; a.ld = h->ld; a.sj = h->sj; a.sk = h->sk; a.nprop = (int)idx.size();
    a.Open = h->raw_i[0]; a.Water = h->raw_i[2]; a.KFloorZ = h->KFloorZ; a.VolumeZ = h->raw_d[4];
    a.Density = (sign > 0 && h->premix_fc) ? h->density : nullptr;
    a.WaterColumnZ = (sign > 0 && h->premix_sd) ? h->wcol : nullptr;
    a.limit = h->sd_limit;
    for (int m = 0; m < a.nprop; ++m) {
        const int n = idx[m];
        a.pa[m] = cur_ptr(h, n);
        a.pref[m] = h->has_ref[n] ? h->ref[n] : nullptr;
        a.off[m] = (n < (int)h->offsets.size()) ? sign * h->offsets[n] : 0.;
        // Property%DischConc(:) is shifted with the field (WP:14725-14727, 14834-14836)
        if (a.off[m] != 0. && h->d_ncell > 0 && n < (int)h->d_conc.size() && h->d_conc[n]) {
            adt_shift_kernel<<<(h->d_ncell + 127) / 128, 128, 0, h->stream>>>(h->d_conc[n], h->d_ncell, a.off[m]);
            h->launches++;
        }
    }
    const dim3 grid((unsigned)((h->I + 127) / 128), (unsigned)h->J, (unsigned)a.nprop);
    if (sign > 0) adt_premix_kernel<<<grid, 128, 0, h->stream>>>(a);
    else adt_offset_kernel<<<grid, 128, 0, h->stream>>>(a);
    CU(h, cudaGetLastError());
    h->launches++;
    return 0;
}

// SetLimitsProperty (WP:20594-20720) for the properties `idx`, after the step and the OffSet removal
int launch_limits(Handle *h, const std::vector<int> &idx) {
    bool any = false;
    for (int n : idx) any = any || (n < (int)h->lim_min_on.size() && (h->lim_min_on[n] || h->lim_max_on[n]));
    if (!any) return 0;
    LimitArgs a{};
    a.I = h->I; a.J = h->J; a.K = h->K; a.ld = h->ld; a.sj = h->sj; a.sk = h->sk; a.docycle = h->opt.Docycle_method;
    a.Water = h->raw_i[2]; a.KFloorZ = h->KFloorZ; a.VolumeZ = h->raw_d[4];
    if (h->mass_created.size() < h->prop.size()) { h->mass_created.resize(h->prop.size(), nullptr); h->mass_destroyed.resize(h->prop.size(), nullptr); }
    for (int m = 0; m < (int)idx.size(); ++m) {
        const int n = idx[m];
        const bool on = n < (int)h->lim_min_on.size() && (h->lim_min_on[n] || h->lim_max_on[n]);
        a.min_on[m] = on ? h->lim_min_on[n] : 0; a.max_on[m] = on ? h->lim_max_on[n] : 0;
        a.vmin[m] = on ? h->lim_min[n] : 0.; a.vmax[m] = on ? h->lim_max[n] : 0.;
        a.pa[m] = cur_ptr(h, n);
        if (on) {
            for (auto *v : {&h->mass_created, &h->mass_destroyed})
                if (!(*v)[n]) {
                    if (int rc = dalloc(h, &(*v)[n], h->n3)) return rc;
                    CU(h, cudaMemsetAsync((*v)[n], 0, h->n3 * sizeof(double), h->stream));
                }
            a.created[m] = h->mass_created[n]; a.destroyed[m] = h->mass_destroyed[n];
        }
    }
    const dim3 grid((unsigned)((h->I + 127) / 128), (unsigned)h->J, (unsigned)idx.size());
    adt_limits_kernel<<<grid, 128, 0, h->stream>>>(a);
    CU(h, cudaGetLastError());
    h->launches++;
    return 0;
}

// device copy of the logical field (nk planes of ld*nj elements, plane pitch ld*njp) between two positions
int copy_field(Handle *h, double *dst, const double *src) {
    CU(h, cudaMemcpy2DAsync(dst, sizeof(double) * (size_t)h->sk, src, sizeof(double) * (size_t)h->sk,
                            sizeof(double) * (size_t)h->ld * h->nj, (size_t)h->nk, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

// columns 0 .. nj-1 in chunks of C, in the order that is safe for the direction of the in-place shift
struct Chunk { int a, b; };
std::vector<Chunk> chunk_order(const Handle *h, bool increasing) {
    std::vector<Chunk> v;
    for (int a = 0; a < h->nj; a += h->C) v.push_back({a, std::min(a + h->C, h->nj) - 1});
    // A halo column stays with the work column beside it: the Orlanski boundary writes the exterior cell of a boundary
    // point (orlanski_exterior), which the carry of a later chunk would put back.  C + 1 columns are still inside the
    // margin of the shift (S = C + 3 against a stencil reach of 2), and the halo column itself is never advanced.
    if (v.size() >= 2 && v.back().a == v.back().b) { v[v.size() - 2].b = v.back().b; v.pop_back(); }
    if (v.size() >= 2 && v.front().a == v.front().b) { v[1].a = v.front().a; v.erase(v.begin()); }
    if (!increasing) std::reverse(v.begin(), v.end());
    return v;
}

// hdir: 0 = the whole step; 1 / 2 = horizontally implicit along j / i: adt_hsolve_kernel (stage 1, into a scratch
// field) and then the vertical half of the step from the intermediate field (stage 2)
int launch_step(Handle *h, const Batch &b, const std::vector<int> &idx, bool timed, int hdir = 0) {
    if (int rc = launch_premix(h, idx, +1)) return rc;
    StepArgs s{};
    s.I = h->I; s.J = h->J; s.K = h->K; s.ld = h->ld; s.nj = h->nj; s.sj = h->sj; s.sk = h->sk;
    s.nprop = (int)idx.size();
    s.ntile_i = (h->I + 30) / 31;
    s.j_begin = h->j_begin; s.j_count = h->j_count;
    const mohid_adt_params &f = b.p[idx[0]];
    s.method_h = f.AdvMethodH; s.limiter_h = f.TVDLimitationH; s.method_v = f.AdvMethodV; s.limiter_v = f.TVDLimitationV;
    s.upwind2_h = f.Upwind2H; s.upwind2_v = f.Upwind2V;
    s.vertical1d = h->opt.Vertical1D; s.xzflow = h->opt.XZFlow;
    s.vrelmax = f.VolumeRelMax; s.dt = f.DTProp;
    s.qx = h->raw_d[0]; s.qy = h->raw_d[1]; s.qz = h->raw_d[2];
    s.dtv = h->dtv; s.vr = h->vr; s.dhu = h->dhu; s.dhv = h->dhv; s.dvz = h->dvz; s.rdz = h->rdz; s.mask = h->mask;
    s.rdx = h->rdx; s.rdy = h->rdy; s.DUX = h->DUX; s.DVY = h->DVY; s.DWZ = h->raw_d[7];
    s.VolumeZ = h->raw_d[4]; s.VolumeZOld = h->raw_d[3];
    s.zero_pivots = h->d_zero_piv;
    s.nfmask = h->have_noflux ? h->nfmask : nullptr;
    s.disch.ncell = h->d_ncell; s.disch.K = h->K; s.disch.ci = h->d_ci; s.disch.cj = h->d_cj;
    s.disch.kmin_eff = h->d_kmin_eff; s.disch.kmax_eff = h->d_kmax_eff; s.disch.cbypass = h->d_cbypass;
    s.disch.flow_k = h->d_flow_k;
    // all properties of a launch move the same way: a property that sits at the other position is re-homed first
    const int shift0 = h->shift[idx[0]];
    for (int n : idx) {
        if (h->shift[n] == shift0) continue;
        // via a scratch field: the two positions overlap
        if ((int)h->hs_tmp.size() < (int)h->prop.size()) h->hs_tmp.resize(h->prop.size(), nullptr);
        if (!h->hs_tmp[n]) if (int rc = dalloc(h, &h->hs_tmp[n], h->n3)) return rc;
        if (int rc = copy_field(h, h->hs_tmp[n], cur_ptr(h, n))) return rc;
        h->shift[n] = shift0;
        if (int rc = copy_field(h, cur_ptr(h, n), h->hs_tmp[n])) return rc;
    }
    bool any_flux = false;
    for (int m = 0; m < s.nprop; ++m) {
        const int n = idx[m];
        const mohid_adt_params &q = b.p[n];
        PropArgs &pa = s.p[m];
        pa.pin = cur_ptr(h, n);
        pa.pout = nxt_ptr(h, n);
        pa.pref = h->has_ref[n] ? h->ref[n] : nullptr;
        pa.theta_difv = (b.optimize && q.ImpExp_DifV > 0.0) ? 1.0 : q.ImpExp_DifV;     // AD:2797 vs AD:2741-2760
        pa.tdec = 1.0 / (1.0 + q.DecayTime / q.DTProp);
        pa.bc = h->has_ref[n] ? q.BoundaryCondition : MOHID_BC_None;                    // AD:5816-5830
        pa.advv_implicit = (q.ImpExp_AdvV == 1.0) ? 1 : 0;
        pa.nfsel = b.eff[n].nfsel;
        const bool hd = h->d_ncell > 0 && n < (int)h->d_conc.size() && h->d_conc[n];
        pa.dconc = hd ? h->d_conc[n] : nullptr;
        pa.dconcmf = hd ? h->d_concmf[n] : nullptr;
        any_flux = any_flux || q.CellFluxes;
    }
    // CellFluxes: the explicit shares are evaluated from the field at time n, which the in-place step overwrites
    if (any_flux) {
        if ((int)h->old_copy.size() < (int)h->prop.size()) h->old_copy.resize(h->prop.size(), nullptr);
        for (int n : idx) {
            if (!b.p[n].CellFluxes) continue;
            if (!h->old_copy[n]) if (int rc = dalloc(h, &h->old_copy[n], h->n3)) return rc;
            if (int rc = copy_field(h, h->old_copy[n], cur_ptr(h, n))) return rc;
        }
    }
    bool stage2 = false;
    if (hdir) {
        // ---- stage 1: implicit horizontal direction (adt_hsolve_kernel.cuh) into a scratch copy of the field ----
        HSolveArgs hs{};
        if ((int)h->wline.size() < (int)h->prop.size()) h->wline.resize(h->prop.size(), nullptr);
        if ((int)h->hs_tmp.size() < (int)h->prop.size()) h->hs_tmp.resize(h->prop.size(), nullptr);
        for (int m = 0; m < s.nprop; ++m) {
            const int n = idx[m];
            if (!h->wline[n]) if (int rc = dalloc(h, &h->wline[n], h->n3)) return rc;
            if (!h->hs_tmp[n]) if (int rc = dalloc(h, &h->hs_tmp[n], h->n3)) return rc;
            if (int rc = copy_field(h, h->hs_tmp[n], cur_ptr(h, n))) return rc;
            hs.wline[m] = h->wline[n];
            s.p[m].pout = h->hs_tmp[n];
        }
        s.twod = h->K == 1 ? 1 : 0;
        const int nc = hdir == 1 ? h->I : h->J;
        const long nunits = (long)s.nprop * ((nc + 30) / 31) * h->K;
        const long blocks = (nunits + 7) / 8;
        if (blocks > 2147483647L) return fail(h, MOHID_ADT_ERR_ARG, "grid too large");
        const bool split = hdir == 1 && (h->j_begin != 1 || h->j_count != h->J);
        if (split) {
            // lines along j on a column slab: the recurrence passes from rank to rank (THOMAS_DDecompHorizGrid gathers the rows
            // on one process instead, HG:8245-8478): forward left to right, back substitution right to left, (W, G) and x of
            // one cell per (i, level, property) on the wire
            if (!h->nccl) return fail(h, MOHID_ADT_ERR_STATE, "implicit advection along j on a column slab needs mohid_adt_comm_init");
            const size_t ne = (size_t)s.nprop * h->K * h->I;
            if (h->line_cap < ne) {
                CU(h, cudaStreamSynchronize(h->stream));
                for (auto &p : h->line_buf) { if (p) cudaFree(p); p = nullptr; }
                for (auto &p : h->line_buf) if (int rc = dalloc(h, &p, 2 * ne)) return rc;
                h->line_cap = ne;
            }
            const bool has_l = h->rank > 0, has_r = h->rank < h->nranks - 1;
            auto nccl_ok = [&](ncclResult_t r) { return r == ncclSuccess ? 0 : fail(h, MOHID_ADT_ERR_CUDA, "NCCL line solve: %s", g_nccl.GetErrorString(r)); };
            hs.split = 1; hs.l0 = h->j_begin; hs.l1 = h->j_begin + h->j_count - 1;
            hs.edge_in = has_l ? h->line_buf[0] : nullptr; hs.edge_out = h->line_buf[1];
            if (has_l) if (int rc = nccl_ok(g_nccl.Recv(h->line_buf[0], 2 * ne, ncclDouble, h->rank - 1, h->nccl, h->stream))) return rc;
            adt_hsolve_kernel<0><<<(unsigned)blocks, 256, 0, h->stream>>>(s, hs);
            CU(h, cudaGetLastError());
            if (has_r) if (int rc = nccl_ok(g_nccl.Send(h->line_buf[1], 2 * ne, ncclDouble, h->rank + 1, h->nccl, h->stream))) return rc;
            HBackArgs ba{};
            ba.I = h->I; ba.K = h->K; ba.sj = h->sj; ba.sk = h->sk; ba.nprop = s.nprop; ba.l0 = hs.l0; ba.l1 = hs.l1; ba.last = has_r ? 0 : 1;
            ba.x_in = h->line_buf[2]; ba.x_out = h->line_buf[3];
            for (int m = 0; m < s.nprop; ++m) { ba.out[m] = h->hs_tmp[idx[m]]; ba.wline[m] = h->wline[idx[m]]; ba.pin[m] = cur_ptr(h, idx[m]); }
            if (has_r) if (int rc = nccl_ok(g_nccl.Recv(h->line_buf[2], ne, ncclDouble, h->rank + 1, h->nccl, h->stream))) return rc;
            adt_hsolve_back_kernel<<<dim3((unsigned)((h->I + 127) / 128), (unsigned)h->K, (unsigned)s.nprop), 128, 0, h->stream>>>(ba);
            CU(h, cudaGetLastError());
            if (has_l) if (int rc = nccl_ok(g_nccl.Send(h->line_buf[3], ne, ncclDouble, h->rank - 1, h->nccl, h->stream))) return rc;
            h->launches += 2;
        } else {
            if (hdir == 1) adt_hsolve_kernel<0><<<(unsigned)blocks, 256, 0, h->stream>>>(s, hs);
            else adt_hsolve_kernel<1><<<(unsigned)blocks, 256, 0, h->stream>>>(s, hs);
            CU(h, cudaGetLastError());
            h->launches++;
        }
        // ---- stage 2: the vertical half, from the intermediate field into the new position ----
        for (int m = 0; m < s.nprop; ++m) { s.p[m].pin = h->hs_tmp[idx[m]]; s.p[m].pout = nxt_ptr(h, idx[m]); }
        stage2 = true;
    }
    s.stage2 = stage2 ? 1 : 0;
    // 2-D domain, horizontally implicit (AD:1758-1841): the line solve was the whole step; the chunk walk below only moves
    // its result into the new position
    const bool line_only = stage2 && h->K == 1;
    // ---- kernel variant and launch shape ----
    bool any_disch = false, all_impv = true;
    for (int m = 0; m < s.nprop; ++m) {
        any_disch = any_disch || s.p[m].dconc != nullptr;
        all_impv = all_impv && s.p[m].advv_implicit;
        any_disch = any_disch || s.p[m].nfsel != 0;     // NoAdvFlux rides on the DISCH variants
        any_disch = any_disch || s.p[m].bc == MOHID_BC_Orlanski;      // ... and so does the Orlanski boundary
    }
    // FULL: 3-D, both horizontal directions, implicit vertical advection for every property of the launch
    const bool full = !s.vertical1d && !s.xzflow && s.K > 1 && all_impv && !stage2;
    const bool tvd_sb = s.method_h == MOHID_P2_TVD && s.method_v == MOHID_P2_TVD && s.limiter_h == MOHID_SuperBee &&
                        s.limiter_v == MOHID_SuperBee;
    const bool upw = s.method_h == MOHID_UpwindOrder1 && s.method_v == MOHID_UpwindOrder1;
    const bool fused = h->fused_now && !stage2;
    const bool lean = (h->lean_now || fused) && !stage2;
    void (*kern)(const StepArgs) = nullptr;
    void (*lkern)(const LeanArgs) = nullptr;      // lean path: the step kernel takes the packs
    void (*fkern)(const FusedArgs) = nullptr;     // fused path: the step kernel builds the packs itself
    LeanArgs la{};
    FusedArgs fa{};
    int wpb;
    size_t smem;
    // Occupancy is bounded by registers (16K per SM sub-partition): 8 warps allow 255 registers per thread,
    // 12 warps 168.  W of the column solve lives in shared memory (K*32 doubles per warp), G is parked in the output
    // array by the 12-warp forms and kept in shared memory by the generic ones.
    const size_t w_bytes = (size_t)h->K * 32 * sizeof(double);
    if (lean) {
        // measured on C3 (profiles/r02_*): 12 warps at 168 registers without spills beat 16 at 128 with; prefetch
        // distance 3: 29.4 ms against 33.8 ms without
        int want = getenv("MOHID_ADT_LEAN_WARPS") ? atoi(getenv("MOHID_ADT_LEAN_WARPS")) : 12;
        const int pfd = getenv("MOHID_ADT_LEAN_PFD") ? atoi(getenv("MOHID_ADT_LEAN_PFD")) : 3;
        while (want > 8 && (size_t)want * w_bytes > (size_t)h->smem_optin) want -= 4;
#define ADT_LEAN_W(M, P) (want >= 16 ? adt_transport_lean_kernel<M, 16, P> : want >= 12 ? adt_transport_lean_kernel<M, 12, P> : adt_transport_lean_kernel<M, 8, P>)
#define ADT_LEAN(M) (pfd >= 4 ? ADT_LEAN_W(M, 4) : pfd == 3 ? ADT_LEAN_W(M, 3) : pfd == 2 ? ADT_LEAN_W(M, 2) : ADT_LEAN_W(M, 0))
        lkern = tvd_sb ? ADT_LEAN(MOHID_P2_TVD) : ADT_LEAN(MOHID_UpwindOrder1);
#undef ADT_LEAN
#undef ADT_LEAN_W
        wpb = want >= 16 ? 16 : want >= 12 ? 12 : 8;
        smem = wpb * w_bytes;
        if (fused) {
            lkern = nullptr;
            fkern = tvd_sb ? adt_transport_fused_kernel<MOHID_P2_TVD, FUSED_MAXP + FR_NCW> : adt_transport_fused_kernel<MOHID_UpwindOrder1, FUSED_MAXP + FR_NCW>;
            if (s.nprop > fused_max_props(h)) return fail(h, MOHID_ADT_ERR_UNKNOWN, "internal: fused launch with %d properties", s.nprop);
            wpb = s.nprop + FR_NCW;
            smem = fused_smem(h->K, s.nprop);
            fill_lean_coef_args(h, f, b.eff[idx[0]], fa.co);
            fa.tiles_per_group = getenv("MOHID_ADT_FUSED_TPG") ? std::max(1, atoi(getenv("MOHID_ADT_FUSED_TPG"))) : 4;
            fa.tiles_per_group = std::min(fa.tiles_per_group, s.ntile_i);
            fa.ngroups = (s.ntile_i + fa.tiles_per_group - 1) / fa.tiles_per_group;
            // roles: warp w runs on SM sub-partition w % 4; a coefficient warp issues about half the instructions of a
            // property warp, so they go, one after the other, to the sub-partition that is loaded most by warp COUNT
            {
                const int T = wpb;
                int load[4] = {0, 0, 0, 0}, cnt[4] = {0, 0, 0, 0};
                for (int w = 0; w < T; ++w) { cnt[w % 4]++; fa.role[w] = 0; }
                bool prod[16] = {false};
                for (int c = 0; c < FR_NCW; ++c) {
                    int best = -1;
                    for (int q4 = 0; q4 < 4; ++q4) {
                        if (load[q4] >= cnt[q4]) continue;                      // no warp left on this sub-partition
                        if (best < 0 || cnt[q4] * 2 - load[q4] > cnt[best] * 2 - load[best]) best = q4;
                    }
                    for (int w = T - 1; w >= 0; --w) if (w % 4 == best && !prod[w]) { prod[w] = true; break; }
                    load[best] += 1;          // 2 per property warp, 1 per coefficient warp: remaining weight = 2 cnt - load
                }
                if (getenv("MOHID_ADT_FUSED_PLAINROLES")) { for (int w = 0; w < T; ++w) prod[w] = w >= s.nprop; }
                int np = 0, nc = 0;
                for (int w = 0; w < T; ++w) fa.role[w] = prod[w] ? (signed char)(-1 - np++) : (signed char)(nc++);
            }
        }
        la.I = s.I; la.J = s.J; la.K = s.K; la.ld = s.ld; la.sj = s.sj; la.sk = s.sk;
        la.nprop = s.nprop; la.ntile_i = s.ntile_i;
        la.ncol = h->pk_ncol; la.nt32 = h->pk_nt32; la.dt = s.dt;
        la.pk = h->pk;
        la.qx = s.qx; la.qy = s.qy; la.qz = s.qz; la.VolumeZ = s.VolumeZ; la.VolumeZOld = s.VolumeZOld;
        la.zero_pivots = s.zero_pivots;
        for (int m = 0; m < s.nprop; ++m) la.p[m] = s.p[m];
    } else if (full && !any_disch && s.method_h == MOHID_P2_TVD && s.method_v == MOHID_P2_TVD && !tvd_sb &&
               s.limiter_h == s.limiter_v && 12 * w_bytes <= (size_t)h->smem_optin) {
        // P2_TVD with one of the other limiters (MinMod, VanLeer, Muscl, PDM) in both directions: same 12-warp form
        switch (s.limiter_h) {
            case MOHID_MinMod: kern = adt_transport_kernel<MOHID_P2_TVD, MOHID_MinMod, MOHID_P2_TVD, MOHID_MinMod, false, true, 12, 1, true>; break;
            case MOHID_VanLeer: kern = adt_transport_kernel<MOHID_P2_TVD, MOHID_VanLeer, MOHID_P2_TVD, MOHID_VanLeer, false, true, 12, 1, true>; break;
            case MOHID_Muscl: kern = adt_transport_kernel<MOHID_P2_TVD, MOHID_Muscl, MOHID_P2_TVD, MOHID_Muscl, false, true, 12, 1, true>; break;
            default: kern = adt_transport_kernel<MOHID_P2_TVD, MOHID_PDM, MOHID_P2_TVD, MOHID_PDM, false, true, 12, 1, true>; break;
        }
        wpb = 12;
        smem = wpb * w_bytes;
    } else if (full && !any_disch && (tvd_sb || upw) && 12 * w_bytes <= (size_t)h->smem_optin) {
        kern = tvd_sb ? adt_transport_kernel<MOHID_P2_TVD, MOHID_SuperBee, MOHID_P2_TVD, MOHID_SuperBee, false, true, 12, 1, true>
                      : adt_transport_kernel<MOHID_UpwindOrder1, MOHID_SuperBee, MOHID_UpwindOrder1, MOHID_SuperBee, false, true, 12, 1, true>;
        wpb = 12;
        smem = wpb * w_bytes;
    } else {
#define ADT_PICK(D, F)                                                                                              \
    (tvd_sb ? adt_transport_kernel<MOHID_P2_TVD, MOHID_SuperBee, MOHID_P2_TVD, MOHID_SuperBee, D, F>                \
     : upw  ? adt_transport_kernel<MOHID_UpwindOrder1, MOHID_SuperBee, MOHID_UpwindOrder1, MOHID_SuperBee, D, F>    \
            : adt_transport_kernel<0, 0, 0, 0, D, F>)
        if (any_disch) kern = ADT_PICK(true, false);
        else if (full) kern = ADT_PICK(false, true);
        else kern = ADT_PICK(false, false);
#undef ADT_PICK
        wpb = (int)std::min<size_t>(8, (size_t)h->smem_optin / (2 * w_bytes));
        smem = (size_t)2 * wpb * w_bytes;
    }
    if (wpb < 1)
        return fail(h, MOHID_ADT_ERR_UNSUPPORTED, "K = %d layers need more shared memory than one SM has", h->K);
    if ((long)s.nprop * s.ntile_i * h->C / wpb + 1 > 2147483647L) return fail(h, MOHID_ADT_ERR_ARG, "grid too large");
    if (fkern) CU(h, cudaFuncSetAttribute(fkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else if (lkern) CU(h, cudaFuncSetAttribute(lkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    // ---- the step, chunk by chunk ----
    const int ja = h->j_begin, jb = h->j_begin + h->j_count - 1;      // columns the step kernel advances
    CarryArgs ca{};
    ca.ld = h->ld; ca.nk = h->nk; ca.I = h->I; ca.J = h->J; ca.K = h->K; ca.sj = h->sj; ca.sk = h->sk; ca.ja = ja; ca.jb = jb;
    ca.Water = h->raw_i[2];
    // cells the step does not advance keep the field at time n (in a two-stage step pin is the intermediate field)
    for (int m = 0; m < s.nprop; ++m) { ca.src[m] = cur_ptr(h, idx[m]); ca.dst[m] = s.p[m].pout; }
    if (line_only) { ca.twod = 1; for (int m = 0; m < s.nprop; ++m) ca.line[m] = h->hs_tmp[idx[m]]; }
    const bool packs_per_chunk = lean && !fused && h->pk_ncol < h->nj;
    if (timed) h->ev_steps++;
    for (const Chunk &c : chunk_order(h, shift0 == h->S)) {
        ca.j0 = c.a;
        adt_carry_kernel<<<dim3((unsigned)((h->ld + 127) / 128), (unsigned)(c.b - c.a + 1), (unsigned)s.nprop), 128, 0, h->stream>>>(ca);
        CU(h, cudaGetLastError());
        h->launches++;
        const int ka = std::max(c.a, ja), kb = std::min(c.b, jb);
        if (ka > kb || line_only) continue;
        if (packs_per_chunk) if (int rc = launch_lean_coef(h, b.p[idx[0]], b.eff[idx[0]], ka, kb - ka + 2)) return rc;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (timed) {
            if (h->ev_used == h->ev.size()) {
                cudaEvent_t x, y;
                CU(h, cudaEventCreate(&x));
                CU(h, cudaEventCreate(&y));
                h->ev.emplace_back(x, y);
            }
            e0 = h->ev[h->ev_used].first; e1 = h->ev[h->ev_used].second;
            h->ev_used++;
            CU(h, cudaEventRecord(e0, h->stream));
        }
        const long nu = (long)s.nprop * s.ntile_i * (kb - ka + 1);
        const unsigned blocks = (unsigned)((nu + wpb - 1) / wpb);
        if (fkern) {
            la.j_begin = ka; la.j_count = kb - ka + 1;
            fa.st = la;
            fkern<<<(unsigned)((long)fa.ngroups * fa.tiles_per_group * la.j_count), wpb * 32, smem, h->stream>>>(fa);
        } else if (lkern) {
            la.j_begin = ka; la.j_count = kb - ka + 1; la.jc0 = h->pk_jc0;
            lkern<<<blocks, wpb * 32, smem, h->stream>>>(la);
        } else {
            s.j_begin = ka; s.j_count = kb - ka + 1;
            kern<<<blocks, wpb * 32, smem, h->stream>>>(s);
        }
        CU(h, cudaGetLastError());
        h->launches++;
        if (timed) CU(h, cudaEventRecord(e1, h->stream));
    }
    for (int n : idx) h->shift[n] = h->S - h->shift[n];               // the new field is the current one now

    // ---- post-solve boundary passes on the new field (AD:1874-1882) ----
    for (int m = 0; m < s.nprop; ++m) {
        const int n = idx[m];
        const int bc = b.p[n].BoundaryCondition;
        if (!h->has_ref[n] || h->n_bnd_cols == 0) continue;
        BndArgs ba{};
        ba.I = h->I; ba.J = h->J; ba.K = h->K; ba.ld = h->ld; ba.nj = h->nj; ba.ncols = h->n_bnd_cols;
        ba.sj = s.sj; ba.sk = s.sk; ba.cols = h->bnd_cols; ba.kfloor = h->KFloorZ;
        ba.CFU = h->raw_i[3]; ba.CFV = h->raw_i[4]; ba.Bnd = h->Bnd;
        ba.prop = cur_ptr(h, n); ba.pref = h->ref[n]; ba.jmin = 0; ba.jmax = 2147483647;
        const long tot = (long)h->n_bnd_cols * h->K;
        if (bc == MOHID_BC_NullGradient) {
            adt_nullgrad_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(ba);
            h->launches++;
        } else if (bc == MOHID_BC_CyclicBoundary) {
            adt_cyclic_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(ba, 0);
            const long t1 = (long)std::max(h->I - 2, 0) * h->K, t2 = (long)std::max(h->J - 2, 0) * h->K;
            const bool slab = h->j_begin != 1 || h->j_count != h->J;
            h->launches += 1;
            if (!slab) {
                if (t1 > 0) { adt_cyclic_kernel<<<(unsigned)((t1 + 255) / 256), 256, 0, h->stream>>>(ba, 1); h->launches++; }
            } else if (h->nranks > 1 && (h->rank == 0 || h->rank == h->nranks - 1)) {
                // the j wrap joins the first and the last rank: each sends the column next to its boundary column
                const size_t ne = (size_t)h->K * h->I + 2 * (size_t)h->I;
                if (h->cyc_cap < ne) {
                    CU(h, cudaStreamSynchronize(h->stream));
                    for (auto &p : h->cyc_buf) { if (p) cudaFree(p); p = nullptr; }
                    for (auto &p : h->cyc_buf) if (int rc = dalloc(h, &p, ne)) return rc;
                    h->cyc_cap = ne;
                }
                const bool first = h->rank == 0;
                const int j_bnd = first ? 1 : h->J, j_val = first ? 2 : h->J - 1, peer = first ? h->nranks - 1 : 0;
                const long tp = (long)h->K * h->I + h->I;
                adt_cyclic_edge_pack_kernel<<<(unsigned)((tp + 255) / 256), 256, 0, h->stream>>>(ba, j_val, j_bnd, h->cyc_buf[0]);
                ncclResult_t r = g_nccl.GroupStart();
                if (r == ncclSuccess) r = g_nccl.Send(h->cyc_buf[0], ne, ncclDouble, peer, h->nccl, h->stream);
                if (r == ncclSuccess) r = g_nccl.Recv(h->cyc_buf[1], ne, ncclDouble, peer, h->nccl, h->stream);
                const ncclResult_t r2 = g_nccl.GroupEnd();
                if (r != ncclSuccess || r2 != ncclSuccess)
                    return fail(h, MOHID_ADT_ERR_CUDA, "NCCL cyclic boundary: %s", g_nccl.GetErrorString(r != ncclSuccess ? r : r2));
                if (t1 > 0) { adt_cyclic_edge_apply_kernel<<<(unsigned)((t1 + 255) / 256), 256, 0, h->stream>>>(ba, j_bnd, h->cyc_buf[1]); h->launches++; }
                h->launches += 1;
            }
            if (t2 > 0) { adt_cyclic_kernel<<<(unsigned)((t2 + 255) / 256), 256, 0, h->stream>>>(ba, 2); h->launches++; }
        }
        CU(h, cudaGetLastError());
    }
    // cell-face fluxes of the properties that asked for them (AD:1885-1916: after the boundary passes)
    for (int m = 0; m < s.nprop; ++m) {
        const int n = idx[m];
        const mohid_adt_params &q = b.p[n];
        if (!q.CellFluxes) continue;
        if (int rc = ensure_legacy(h)) return rc;
        for (auto &v : h->flux) if ((int)v.size() <= n) v.resize(n + 1, nullptr);
        for (auto &v : h->flux) {
            if (!v[n]) if (int rc = dalloc(h, &v[n], h->n3)) return rc;
            CU(h, cudaMemsetAsync(v[n], 0, h->n3 * sizeof(double), h->stream));       // AD:1457-1470
        }
        FluxArgs fa{};
        fa.I = h->I; fa.J = h->J; fa.K = h->K; fa.ld = h->ld; fa.sj = h->sj; fa.sk = h->sk;
        fa.method_h = q.AdvMethodH; fa.limiter_h = q.TVDLimitationH; fa.method_v = q.AdvMethodV; fa.limiter_v = q.TVDLimitationV;
        fa.upwind2_h = q.Upwind2H; fa.upwind2_v = q.Upwind2V; fa.vertical1d = h->opt.Vertical1D; fa.xzflow = h->opt.XZFlow;
        fa.vrelmax = q.VolumeRelMax; fa.w_advv = q.ImpExp_AdvV; fa.theta = q.ImpExp_DifV;
        fa.nfmask = s.nfmask; fa.nfsel = b.eff[n].nfsel;
        fa.pold = h->old_copy[n]; fa.pnew = cur_ptr(h, n);
        // split step (K > 1): the vertical half restarted from the line solve's result, still in its scratch field
        fa.pmid = (hdir && h->K > 1) ? h->hs_tmp[n] : fa.pold;
        fa.impl_x = hdir == 1; fa.impl_y = hdir == 2;
        fa.qx = s.qx; fa.qy = s.qy; fa.qz = s.qz; fa.dtv = h->dtv; fa.dhu = h->dhu; fa.dhv = h->dhv; fa.dvz = h->dvz;
        fa.rdz = h->rdz; fa.rdx = s.rdx; fa.rdy = s.rdy; fa.DUX = s.DUX; fa.DVY = s.DVY; fa.DWZ = s.DWZ; fa.mask = h->mask;
        fa.ax = h->flux[0][n]; fa.ay = h->flux[1][n]; fa.az = h->flux[2][n];
        fa.dx = h->flux[3][n]; fa.dy = h->flux[4][n]; fa.dz = h->flux[5][n];
        const dim3 grid((unsigned)((h->I + 127) / 128), (unsigned)h->J, (unsigned)h->K);
        adt_cell_flux_kernel<<<grid, 128, 0, h->stream>>>(fa);
        CU(h, cudaGetLastError());
        h->launches++;
    }
    if (int rc = launch_premix(h, idx, -1)) return rc;
    if (int rc = launch_limits(h, idx)) return rc;
    return 0;
}

// One transport step of all properties of the batch: per-step coefficient pass + fused kernel,
// one (K1 diff part + K2) group per distinct set of effective diffusion flags.  With chunk > 0 a group is launched in
// pieces of at most `chunk` properties, `before` / `after` being called around each piece (the pipelined host path
// uses them to wait for the piece's upload and to start its download).
using ChunkHook = std::function<int(const std::vector<int> &)>;
int step_once(Handle *h, const Batch &b, int chunk = 0, const ChunkHook &before = nullptr, const ChunkHook &after = nullptr) {
    std::vector<char> done(b.nprop, 0);
    bool geom_done = false;
    if (h->halo_pending) { CU(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0)); h->halo_pending = false; }
    h->fused_now = fused_eligible(h, b);
    h->lean_now = !h->fused_now && lean_eligible(h, b);
    if (h->fused_now) {
        if (int rc = ensure_rho(h)) return rc;
    } else if (h->lean_now) {
        // the packs of the whole grid when they fit beside everything else (one coefficient pass per step), else of
        // one column chunk at a time (the pass then runs chunk by chunk, interleaved with the step kernel)
        int ncol = h->pk_ncol;
        if (ncol == 0) {
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            const size_t full_b = (size_t)((h->ni + 31) / 32) * 128 * sizeof(Pack4) * h->nj * h->nk;
            const bool want_full = getenv("MOHID_ADT_PACK_CHUNKED") ? false : (full_b < free_b / 2);
            ncol = want_full ? h->nj : std::min(h->nj, h->C + 1);
        }
        if (int rc = ensure_lean(h, ncol)) return rc;
    } else if (int rc = ensure_legacy(h)) return rc;
    if (h->premix_sd) {                                   // Me%SmallDepths%ON (WP:12975-12980), consumed by K1
        PremixArgs a{};
        a.I = h->I; a.J = h->J; a.K = h->K; a.ld = h->ld; a.sj = h->sj; a.sk = h->sk;
        a.Open = h->raw_i[0]; a.WaterColumnZ = h->wcol; a.limit = h->sd_limit; a.SmallDepths = h->SmallDepths;
        adt_small_depths_kernel<<<dim3((unsigned)((h->ld + 127) / 128), (unsigned)h->nj), 128, 0, h->stream>>>(a);
        CU(h, cudaGetLastError());
        h->launches++;
        h->have_small = true;
    }
    for (int n = 0; n < b.nprop; ++n) {
        if (done[n]) continue;
        std::vector<int> idx;
        for (int m = n; m < b.nprop; ++m) {
            if (done[m]) continue;
            if (b.eff[n].same_dif(b.eff[m])) {
                idx.push_back(m);
                done[m] = 1;
            }
        }
        if (h->fused_now) {
            // no coefficient pass: the fused kernel builds the packs on the fly
        } else if (h->lean_now) {
            if (h->pk_ncol >= h->nj) if (int rc = launch_lean_coef(h, b.p[n], b.eff[n], 0, h->nj)) return rc;
        } else if (int rc = launch_coef(h, b.p[n], b.eff[n], !geom_done, true)) return rc;
        if (!geom_done && h->d_ncell > 0 && !h->lean_now && !h->fused_now) {   // flag the receiving cells, per-layer flows (AD:4063-4077)
            DischArgs d{};
            d.ncell = h->d_ncell; d.K = h->K; d.ld = h->ld; d.sj = h->sj; d.sk = h->sk;
            d.ci = h->d_ci; d.cj = h->d_cj; d.ck = h->d_ck; d.ckmin = h->d_ckmin; d.ckmax = h->d_ckmax;
            d.cvert = h->d_cvert; d.cbypass = h->d_cbypass; d.cflow = h->d_cflow;
            d.kmin_eff = h->d_kmin_eff; d.kmax_eff = h->d_kmax_eff; d.flow_k = h->d_flow_k;
            d.kfloor = h->KFloorZ; d.DWZ = h->raw_d[7]; d.mask = h->mask;
            adt_discharge_prep_kernel<<<(h->d_ncell + 127) / 128, 128, 0, h->stream>>>(d);
            CU(h, cudaGetLastError());
            h->launches++;
        }
        geom_done = true;
        // with per-chunk packs every piece would repeat the coefficient pass: the group then goes in one piece
        const bool one_piece = chunk <= 0 || (h->lean_now && h->pk_ncol < h->nj);
        size_t step = one_piece ? idx.size() : (size_t)chunk;
        if (h->fused_now) {                             // at most `maxp` property warps per block, pieces of equal size
            const size_t maxp = (size_t)fused_max_props(h);
            const size_t pieces = (idx.size() + maxp - 1) / maxp;
            const size_t even = (idx.size() + pieces - 1) / pieces;
            step = one_piece ? even : std::min(step, even);
        }
        for (size_t c0 = 0; c0 < idx.size(); c0 += step) {
            const std::vector<int> part(idx.begin() + c0, idx.begin() + std::min(idx.size(), c0 + step));
            if (before) if (int rc = before(part)) return rc;
            // horizontally implicit properties take the split path, per direction (WP:14676-14700 alternates it)
            for (int hdir = 0; hdir <= 2; ++hdir) {
                std::vector<int> sub;
                for (int m : part) {
                    const mohid_adt_params &q = b.p[m];
                    const int d = h->opt.Vertical1D ? 0 : (q.ImpExp_AdvXX == 1.0 ? 1 : q.ImpExp_AdvYY == 1.0 ? 2 : 0);
                    if (d == hdir) sub.push_back(m);
                }
                if (!sub.empty()) if (int rc = launch_step(h, b, sub, true, hdir)) return rc;
            }
            if (after) if (int rc = after(part)) return rc;
        }
    }
    if (h->ev_edges) CU(h, cudaEventRecord(h->ev_edges, h->stream));      // the halo exchange follows the step
    return 0;
}

}  // namespace

// =======================================================================================
extern "C" {

int mohid_adt_version(char *buf, const int *buflen) {
    if (!buf || !buflen || *buflen <= 0) return MOHID_ADT_ERR_ARG;
    snprintf(buf, (size_t)*buflen, "mohid_adt 0.1 (sm_100a, CUDA %d)", CUDART_VERSION);
    return 0;
}

int mohid_adt_last_error(const int *handle, char *buf, const int *buflen) {
    if (!buf || !buflen || *buflen <= 0) return MOHID_ADT_ERR_ARG;
    Handle *h = get(handle);
    std::lock_guard<std::mutex> lk(g_mu);
    snprintf(buf, (size_t)*buflen, "%s", h ? h->err.c_str() : g_err.c_str());
    return 0;
}

int mohid_adt_create(int *handle, const mohid_adt_size3d *size, const mohid_adt_size3d *worksize, const int *ld_i,
                     const mohid_adt_options *opt) {
    if (!handle || !size || !worksize) return fail(nullptr, MOHID_ADT_ERR_ARG, "null argument");
    if (size->ILB != 0 || size->JLB != 0 || size->KLB != 0 || worksize->ILB != 1 || worksize->JLB != 1 ||
        worksize->KLB != 1 || size->IUB != worksize->IUB + 1 || size->JUB != worksize->JUB + 1 ||
        size->KUB != worksize->KUB + 1)
        return fail(nullptr, MOHID_ADT_ERR_ARG, "Size must be (0:I+1,0:J+1,0:K+1) and WorkSize (1:I,1:J,1:K)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, MOHID_ADT_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    cudaGetErrorString(e));
    Handle *h = new Handle();
    if (opt) h->opt = *opt;
    if (h->opt.Docycle_method == 0) h->opt.Docycle_method = 1;
    int dev = h->opt.device;
    if (dev < 0) cudaGetDevice(&dev);
    if (dev >= ndev) { delete h; return fail(nullptr, MOHID_ADT_ERR_ARG, "device %d does not exist", dev); }
    h->dev = dev;
    if (cudaSetDevice(dev) != cudaSuccess) { delete h; return fail(nullptr, MOHID_ADT_ERR_CUDA, "cudaSetDevice(%d) failed", dev); }
    h->I = worksize->IUB; h->J = worksize->JUB; h->K = worksize->KUB;
    h->ni = h->I + 2; h->nj = h->J + 2; h->nk = h->K + 2;
    h->j_begin = 1; h->j_count = h->J;
    h->ld_h = (ld_i && *ld_i > 0) ? *ld_i : h->ni;
    if (h->ld_h < h->ni) { delete h; return fail(nullptr, MOHID_ADT_ERR_ARG, "ld_i smaller than I+2"); }
    h->ld = ((h->ni + 15) / 16) * 16;                // 128-byte aligned rows on the device
    h->n2 = (long)h->ld * h->nj;
    // chunk width of the in-place step and the shift margin it needs (see Handle)
    // The margin of S = C + 3 columns is what the single buffer per property costs; the wider the chunk, the fewer launches
    // and tail waves per step (C3: 40.8 ms with 17 chunks, 40.1 with 5).  The widest of nj, nj/2, .. nj/16 is taken whose
    // footprint -- 112 B of raw mirrors + 8 B per property per cell, for max_properties (at least 16) properties --
    // stays below 55 % of the free device memory; nj/16 (32 .. 256 columns) is the floor.
    {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const double per_cell = 112.0 + 8.0 * std::max(h->opt.max_properties, 16);
        const int floor_c = std::min(h->nj, std::max(32, std::min(256, h->nj / 16)));
        h->C = floor_c;
        for (int div = 1; div <= 8; div *= 2) {
            const int c = std::max(floor_c, (h->nj + div - 1) / div);
            const double cells = (double)h->ld * (h->nj + c + 3) * h->nk;
            if (cells < 2.0e9 && cells * per_cell < 0.55 * (double)free_b) { h->C = c; break; }
        }
    }
    if (const char *e = getenv("MOHID_ADT_CHUNK_COLS")) h->C = std::max(1, std::min(h->nj, atoi(e)));
    h->S = h->C + 3;
    h->njp = h->nj + h->S;
    h->n3 = (long)h->ld * h->njp * h->nk;
    h->sj = h->ld; h->sk = h->ld * h->njp;                          // element (i,j,k) at i + ld*(j + njp*k)
    if (h->n3 >= 2147483647L || h->nj > 65535 || h->nk > 65535) {
        delete h;
        return fail(nullptr, MOHID_ADT_ERR_ARG, "one field must hold fewer than 2^31 elements (32-bit cell indices)");
    }
    h->maxprop = h->opt.max_properties > 0 ? h->opt.max_properties : 0;
    cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return fail(nullptr, MOHID_ADT_ERR_CUDA, "cudaStreamCreate failed");
    }
    h->own_stream = true;
    int rc = 0;
    rc |= dalloc(h, &h->DUX, h->n2); rc |= dalloc(h, &h->DVY, h->n2); rc |= dalloc(h, &h->DZX, h->n2);
    rc |= dalloc(h, &h->DZY, h->n2); rc |= dalloc(h, &h->rdx, h->n2); rc |= dalloc(h, &h->rdy, h->n2);
    rc |= dalloc(h, &h->KFloorZ, h->n2); rc |= dalloc(h, &h->Bnd, h->n2); rc |= dalloc(h, &h->SmallDepths, h->n2);
    for (auto &p : h->raw_d) rc |= dalloc(h, &p, h->n3);
    for (auto &p : h->raw_i) rc |= dalloc(h, &p, h->n3);
    rc |= dalloc(h, &h->d_zero_piv, 1);
    if (rc) { std::string m = h->err; free_all(h); delete h; return fail(nullptr, MOHID_ADT_ERR_CUDA, "%s", m.c_str()); }
    // padded columns must read as zeros
    for (auto p : h->raw_d) cudaMemsetAsync(p, 0, h->n3 * sizeof(double), h->stream);
    for (auto p : h->raw_i) cudaMemsetAsync(p, 0, h->n3 * sizeof(int), h->stream);
    for (auto p : {h->DUX, h->DVY, h->DZX, h->DZY}) cudaMemsetAsync(p, 0, h->n2 * sizeof(double), h->stream);
    for (auto p : {h->KFloorZ, h->Bnd, h->SmallDepths}) cudaMemsetAsync(p, 0, h->n2 * sizeof(int), h->stream);
    cudaMemsetAsync(h->d_zero_piv, 0, sizeof(unsigned long long), h->stream);
    if (const cudaError_t e2 = cudaStreamSynchronize(h->stream)) {
        free_all(h);
        delete h;
        return fail(nullptr, MOHID_ADT_ERR_CUDA, "clearing the device mirrors failed: %s", cudaGetErrorString(e2));
    }
    std::lock_guard<std::mutex> lk(g_mu);
    h->id = g_next++;
    g_h[h->id] = h;
    *handle = h->id;
    return 0;
}

int mohid_adt_destroy(int *handle) {
    if (!handle) return MOHID_ADT_ERR_ARG;
    Handle *h = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_h.find(*handle);
        if (it == g_h.end()) return MOHID_ADT_ERR_HANDLE;
        h = it->second;
        g_h.erase(it);
    }
    cudaSetDevice(h->dev);
    // work queued on the communication / edge / copy streams may still touch the buffers freed below
    if (h->comm) cudaStreamSynchronize(h->comm);
    if (h->s_comm_own) cudaStreamSynchronize(h->s_comm_own);
    if (h->s_up) cudaStreamSynchronize(h->s_up);
    if (h->s_down) cudaStreamSynchronize(h->s_down);
    cudaStreamSynchronize(h->stream);
    free_all(h);
    delete h;
    *handle = 0;
    return 0;
}

int mohid_adt_set_stream(const int *handle, void *cuda_stream) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    CU(h, cudaSetDevice(h->dev));
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream;
    h->own_stream = false;
    return 0;
}

int mohid_adt_set_grid2d(const int *handle, const double *DUX, const double *DVY, const double *DZX,
                         const double *DZY, const int *KFloorZ, const int *BoundaryPoints2D) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!DUX || !DVY || !DZX || !DZY || !KFloorZ || !BoundaryPoints2D) return fail(h, MOHID_ADT_ERR_ARG, "null array");
    CU(h, cudaSetDevice(h->dev));
    int rc = 0;
    rc |= h2d2(h, h->DUX, DUX, 8); rc |= h2d2(h, h->DVY, DVY, 8);
    rc |= h2d2(h, h->DZX, DZX, 8); rc |= h2d2(h, h->DZY, DZY, 8);
    rc |= h2d2(h, h->KFloorZ, KFloorZ, 4); rc |= h2d2(h, h->Bnd, BoundaryPoints2D, 4);
    if (rc) return rc;
    adt_grid2d_kernel<<<std::max(1, (int)std::min<long>((h->n2 + 255) / 256, 4096)), 256, 0, h->stream>>>(
        h->ni, h->nj, h->ld, h->DUX, h->DVY, h->rdx, h->rdy);
    CU(h, cudaGetLastError());
    h->launches++;
    // compact list of boundary columns for the post-solve passes (built from the device mirror so the
    // caller's array may be a host or a device pointer)
    std::vector<int> bhost((size_t)h->n2);
    CU(h, cudaMemcpyAsync(bhost.data(), h->Bnd, h->n2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    h->bnd_host.clear();
    for (int j = 1; j <= h->J; ++j)
        for (int i = 1; i <= h->I; ++i)
            if (bhost[(size_t)i + (size_t)h->ld * j] == 1) { h->bnd_host.push_back(i); h->bnd_host.push_back(j); }
    if (int r = upload_bnd_cols(h)) return r;
    CU(h, cudaStreamSynchronize(h->stream));
    h->have_grid = true;
    return 0;
}

int mohid_adt_set_step(const int *handle, const double *Wflux_X, const double *Wflux_Y, const double *Wflux_Z,
                       const double *VolumeZOld, const double *VolumeZ, const double *Visc_H, const double *Diff_V,
                       const double *DWZ, const double *DZZ, const double *AreaU, const double *AreaV,
                       const int *OpenPoints3D, const int *LandPoints3D, const int *WaterPoints3D,
                       const int *ComputeFacesU3D, const int *ComputeFacesV3D, const int *ComputeFacesW3D,
                       const int *SmallDepths) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    CU(h, cudaSetDevice(h->dev));
    const double *d[11] = {Wflux_X, Wflux_Y, Wflux_Z, VolumeZOld, VolumeZ, Visc_H, Diff_V, DWZ, DZZ, AreaU, AreaV};
    const int *m[6] = {OpenPoints3D, LandPoints3D, WaterPoints3D, ComputeFacesU3D, ComputeFacesV3D, ComputeFacesW3D};
    // after a first complete call a NULL array means "unchanged since the last step": its device mirror is kept
    // (the land / water maps never change, the others only with wetting and drying)
    if (!h->have_step) {
        for (auto p : d) if (!p) return fail(h, MOHID_ADT_ERR_ARG, "null array");
        for (auto p : m) if (!p) return fail(h, MOHID_ADT_ERR_ARG, "null mask (WaterPoints3D is required: THOMASZ_NewType2 reads it, MF:4086)");
    }
    for (int a = 0; a < 11; ++a) if (d[a]) if (int rc = h2d3(h, h->raw_d[a], d[a], 8)) return rc;
    for (int a = 0; a < 6; ++a) if (m[a]) if (int rc = h2d3(h, h->raw_i[a], m[a], 4)) return rc;
    h->have_small = SmallDepths != nullptr;
    if (SmallDepths) if (int rc = h2d2(h, h->SmallDepths, SmallDepths, 4)) return rc;
    CU(h, cudaStreamSynchronize(h->stream));     // the host arrays are only borrowed for the call (AD:2229-2349)
    h->have_step = true;
    return 0;
}

static int check_cols(Handle *h, const int *j0, const int *ncols) {
    if (!j0 || !ncols || *j0 < 0 || *ncols < 1 || *j0 + *ncols > h->nj)
        return fail(h, MOHID_ADT_ERR_ARG, "column window must lie inside 0..J+1");
    return 0;
}

int mohid_adt_set_step_columns(const int *handle, const int *j0, const int *ncols, const double *Wflux_X,
                               const double *Wflux_Y, const double *Wflux_Z, const double *VolumeZOld,
                               const double *VolumeZ, const double *Visc_H, const double *Diff_V, const double *DWZ,
                               const double *DZZ, const double *AreaU, const double *AreaV, const int *OpenPoints3D,
                               const int *LandPoints3D, const int *WaterPoints3D, const int *ComputeFacesU3D,
                               const int *ComputeFacesV3D, const int *ComputeFacesW3D) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (int rc = check_cols(h, j0, ncols)) return rc;
    CU(h, cudaSetDevice(h->dev));
    const double *d[11] = {Wflux_X, Wflux_Y, Wflux_Z, VolumeZOld, VolumeZ, Visc_H, Diff_V, DWZ, DZZ, AreaU, AreaV};
    const int *m[6] = {OpenPoints3D, LandPoints3D, WaterPoints3D, ComputeFacesU3D, ComputeFacesV3D, ComputeFacesW3D};
    for (int a = 0; a < 11; ++a) if (d[a]) if (int rc = copy3(h, h->raw_d[a], d[a], 8, true, nullptr, *j0, *ncols)) return rc;
    for (int a = 0; a < 6; ++a) if (m[a]) if (int rc = copy3(h, h->raw_i[a], m[a], 4, true, nullptr, *j0, *ncols)) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mohid_adt_upload_props_columns(const int *handle, const int *nprop, const double *const *prop,
                                   const double *const *reference_prop, const int *j0, const int *ncols) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!nprop || !prop) return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    if (int rc = check_cols(h, j0, ncols)) return rc;
    CU(h, cudaSetDevice(h->dev));
    if (int rc = ensure_props(h, *nprop, reference_prop != nullptr)) return rc;
    for (int n = 0; n < *nprop; ++n) {
        if (!prop[n]) return fail(h, MOHID_ADT_ERR_ARG, "prop[%d] is null", n);
        if (int rc = copy3(h, cur_ptr(h, n), prop[n], 8, true, nullptr, *j0, *ncols)) return rc;
        const double *r = reference_prop ? reference_prop[n] : nullptr;
        if (r) {
            if (!h->ref[n]) {
                if (int rc = dalloc(h, &h->ref[n], h->n3)) return rc;
                CU(h, cudaMemsetAsync(h->ref[n], 0, h->n3 * sizeof(double), h->stream));
            }
            if (int rc = copy3(h, h->ref[n], r, 8, true, nullptr, *j0, *ncols)) return rc;
            h->has_ref[n] = 1;
        }
    }
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mohid_adt_download_props_columns(const int *handle, const int *nprop, double *const *prop, const int *j0,
                                     const int *ncols) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!nprop || !prop) return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    if (*nprop > (int)h->prop.size()) return fail(h, MOHID_ADT_ERR_STATE, "properties were never uploaded");
    if (int rc = check_cols(h, j0, ncols)) return rc;
    CU(h, cudaSetDevice(h->dev));
    if (h->halo_pending) { CU(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0)); h->halo_pending = false; }
    for (int n = 0; n < *nprop; ++n)
        if (int rc = copy3(h, cur_ptr(h, n), prop[n], 8, false, nullptr, *j0, *ncols)) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mohid_adt_step_input_device_ptr(const int *handle, const int *which, void **dptr, int *ld, int *nj, int *nk) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!which || !dptr) return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    const int w = *which;
    if (w >= 0 && w < 11) *dptr = h->raw_d[w];
    else if (w >= 11 && w < 17) *dptr = h->raw_i[w - 11];
    else if (w == 17) *dptr = h->SmallDepths;
    else if (w >= 20 && w < 26) {
        void *g[6] = {h->DUX, h->DVY, h->DZX, h->DZY, h->KFloorZ, h->Bnd};
        *dptr = g[w - 20];
    } else return fail(h, MOHID_ADT_ERR_ARG, "unknown array id %d", w);
    if (ld) *ld = h->ld;
    if (nj) *nj = (w < 17) ? h->njp : h->nj;
    if (nk) *nk = h->nk;
    return 0;
}

int mohid_adt_mark_step_resident(const int *handle, const int *small_depths_present) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    h->have_step = true;
    h->have_small = small_depths_present && *small_depths_present;
    return 0;
}

int mohid_adt_set_discharges(const int *handle, const int *prop_index, const int *DischNumber, const int *n_cells,
                             const double *DischFlow, const double *DischConc, const int *DischI, const int *DischJ,
                             const int *DischK, const int *DischKmin, const int *DischKmax, const int *DischVert,
                             const int *IgnoreDisch, const int *DischnCells, const int *ByPass,
                             const double *DischConcMF) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!prop_index || !DischNumber || !n_cells) return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    const int nd = *DischNumber, nc = *n_cells, pi = *prop_index;
    if (pi < 0 || pi >= NPMAX) return fail(h, MOHID_ADT_ERR_ARG, "prop_index out of range");
    if (nd < 0 || nc < 0) return fail(h, MOHID_ADT_ERR_ARG, "negative discharge count");
    if (nc > 0 && (!DischFlow || !DischConc || !DischI || !DischJ || !DischK || !DischKmin || !DischKmax ||
                   !DischVert || !IgnoreDisch || !DischnCells || !ByPass || !DischConcMF))
        return fail(h, MOHID_ADT_ERR_ARG, "null discharge array");
    CU(h, cudaSetDevice(h->dev));
    // expand (discharge, cell) pairs exactly like the serial loop AD:4037-4045 (ignored discharges do not advance n)
    std::vector<int> vert, byp;
    int n = 0;
    for (int dis = 0; dis < nd; ++dis) {
        if (IgnoreDisch[dis]) continue;
        for (int c = 0; c < DischnCells[dis]; ++c, ++n) {
            if (n >= nc) return fail(h, MOHID_ADT_ERR_ARG, "DischnCells lists more cells than n_cells");
            if (DischI[n] < 1 || DischI[n] > h->I || DischJ[n] < 1 || DischJ[n] > h->J)
                return fail(h, MOHID_ADT_ERR_ARG, "discharge cell %d lies outside the work range", n);
            // the layers the cell feeds (adt_discharge_prep_kernel): FillValueInt = "from the bottom / to the surface"
            const bool uniform = DischVert[dis] == MOHID_DischUniform;
            auto k_ok = [&](int k) { return k >= 1 && k <= h->K; };
            if (uniform ? ((DischKmin[n] != MOHID_FILL_INT && !k_ok(DischKmin[n])) ||
                           (DischKmax[n] != MOHID_FILL_INT && !k_ok(DischKmax[n])))
                        : !k_ok(DischK[n]))
                return fail(h, MOHID_ADT_ERR_ARG, "discharge cell %d: layer outside 1..%d", n, h->K);
            vert.push_back(DischVert[dis]);
            byp.push_back(ByPass[dis] ? 1 : 0);
        }
    }
    const int ncell = n;
    // the cell list is shared by all properties (Me%Discharge, WP:14761-14773): a list of another length invalidates the
    // concentrations other properties still hold
    if (h->d_ncell > 0 && ncell != h->d_ncell)
        for (size_t m = 0; m < h->d_conc.size(); ++m)
            if ((int)m != pi && h->d_conc[m])
                return fail(h, MOHID_ADT_ERR_STATE, "SetDischarges: %d cells, but property %d holds concentrations for %d "
                            "(call UnSetDischarges first)", ncell, (int)m, h->d_ncell);
    auto up_i = [&](int *&d, const int *src) -> int {
        if (d) { cudaFree(d); d = nullptr; }
        if (ncell == 0) return 0;
        if (int rc = dalloc(h, &d, (size_t)ncell)) return rc;
        CU(h, cudaMemcpy(d, src, ncell * sizeof(int), cudaMemcpyHostToDevice));
        return 0;
    };
    auto up_d = [&](double *&d, const double *src, size_t cnt) -> int {
        if (d) { cudaFree(d); d = nullptr; }
        if (cnt == 0) return 0;
        if (int rc = dalloc(h, &d, cnt)) return rc;
        if (src) CU(h, cudaMemcpy(d, src, cnt * sizeof(double), cudaMemcpyHostToDevice));
        return 0;
    };
    CU(h, cudaStreamSynchronize(h->stream));
    int rc = 0;
    rc |= up_i(h->d_ci, DischI); rc |= up_i(h->d_cj, DischJ); rc |= up_i(h->d_ck, DischK);
    rc |= up_i(h->d_ckmin, DischKmin); rc |= up_i(h->d_ckmax, DischKmax);
    rc |= up_i(h->d_cvert, vert.data()); rc |= up_i(h->d_cbypass, byp.data());
    rc |= up_i(h->d_kmin_eff, DischKmin); rc |= up_i(h->d_kmax_eff, DischKmax);
    rc |= up_d(h->d_cflow, DischFlow, (size_t)ncell);
    rc |= up_d(h->d_flow_k, nullptr, (size_t)ncell * (h->K + 2));
    if (rc) return rc;
    if ((int)h->d_conc.size() <= pi) { h->d_conc.resize(pi + 1, nullptr); h->d_concmf.resize(pi + 1, nullptr); }
    rc |= up_d(h->d_conc[pi], DischConc, (size_t)ncell);
    rc |= up_d(h->d_concmf[pi], DischConcMF, (size_t)ncell);
    if (rc) return rc;
    h->d_ncell = ncell;
    return 0;
}

int mohid_adt_unset_discharges(const int *handle) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    cudaSetDevice(h->dev);
    cudaStreamSynchronize(h->stream);
    h->d_ncell = 0;
    for (auto &p : h->d_conc) { if (p) cudaFree(p); p = nullptr; }
    for (auto &p : h->d_concmf) { if (p) cudaFree(p); p = nullptr; }
    return 0;
}

// caller -> device copy of one property (and its reference field) on stream `st`
static int upload_one(Handle *h, int n, const double *prop, const double *r, cudaStream_t st) {
    if (!prop) return fail(h, MOHID_ADT_ERR_ARG, "prop[%d] is null", n);
    if (int rc = h2d3(h, cur_ptr(h, n), prop, 8, st)) return rc;
    h->has_ref[n] = r != nullptr;
    if (r) {
        if (!h->ref[n]) {
            if (int rc = dalloc(h, &h->ref[n], h->n3)) return rc;
            CU(h, cudaMemsetAsync(h->ref[n], 0, h->n3 * sizeof(double), st));
        }
        if (int rc = h2d3(h, h->ref[n], r, 8, st)) return rc;
    }
    return 0;
}

int mohid_adt_upload_props(const int *handle, const int *nprop, const double *const *prop,
                           const double *const *reference_prop) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!nprop || !prop) return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    CU(h, cudaSetDevice(h->dev));
    if (int rc = ensure_props(h, *nprop, reference_prop != nullptr)) return rc;
    for (int n = 0; n < *nprop; ++n)
        if (int rc = upload_one(h, n, prop[n], reference_prop ? reference_prop[n] : nullptr, h->stream)) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mohid_adt_download_props(const int *handle, const int *nprop, double *const *prop) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!nprop || !prop) return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    if (*nprop > (int)h->prop.size()) return fail(h, MOHID_ADT_ERR_STATE, "properties were never uploaded");
    CU(h, cudaSetDevice(h->dev));
    if (h->halo_pending) { CU(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0)); h->halo_pending = false; }
    for (int n = 0; n < *nprop; ++n)
        if (int rc = d2h3(h, prop[n], cur_ptr(h, n), 8)) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mohid_adt_advect_device(const int *handle, const int *nprop, const mohid_adt_params *params, const int *nsteps) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!nprop || !params) return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    if (!h->have_grid || !h->have_step) return fail(h, MOHID_ADT_ERR_STATE, "set_grid2d / set_step must precede advect");
    if (*nprop > (int)h->prop.size()) return fail(h, MOHID_ADT_ERR_STATE, "properties were never uploaded");
    CU(h, cudaSetDevice(h->dev));
    Batch b;
    if (int rc = validate(h, *nprop, params, b)) return rc;
    const int ns = nsteps ? *nsteps : 1;
    for (int s = 0; s < ns; ++s)
        if (int rc = step_once(h, b)) return rc;
    return 0;
}

int mohid_adt_advect_batch(const int *handle, const int *nprop, double *const *prop,
                           const double *const *reference_prop, const mohid_adt_params *params) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!nprop || !prop || !params) return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    if (!h->have_grid || !h->have_step) return fail(h, MOHID_ADT_ERR_STATE, "set_grid2d / set_step must precede advect");
    Batch b;
    if (int rc = validate(h, *nprop, params, b)) return rc;        // fail before moving any data
    for (int n = 0; n < *nprop; ++n) if (!prop[n]) return fail(h, MOHID_ADT_ERR_ARG, "prop[%d] is null", n);
    CU(h, cudaSetDevice(h->dev));
    if (int rc = ensure_props(h, *nprop, reference_prop != nullptr)) return rc;
    // Pipelined host path: the properties go up on one copy stream, are advanced in pieces of PIPE_CHUNK on the
    // compute stream and come down on a second copy stream, so that on a full-duplex link the download of the first
    // pieces overlaps the upload of the later ones (pinned host arrays; pageable ones are staged by the driver).
    int chunk = 1;      // measured on C3 (PCIe gen5): 1 -> 722 ms, 2 -> 750 ms, 5 -> 835 ms, unpipelined 974 ms per call
    if (const char *e = getenv("MOHID_ADT_PIPE_CHUNK")) chunk = atoi(e);
    if (chunk <= 0 || *nprop <= chunk) {
        for (int n = 0; n < *nprop; ++n)
            if (int rc = upload_one(h, n, prop[n], reference_prop ? reference_prop[n] : nullptr, h->stream)) return rc;
        if (int rc = step_once(h, b)) return rc;
        for (int n = 0; n < *nprop; ++n)
            if (int rc = d2h3(h, prop[n], cur_ptr(h, n), 8)) return rc;
        CU(h, cudaStreamSynchronize(h->stream));
        return 0;
    }
    if (!h->s_up) {
        CU(h, cudaStreamCreateWithFlags(&h->s_up, cudaStreamNonBlocking));
        CU(h, cudaStreamCreateWithFlags(&h->s_down, cudaStreamNonBlocking));
    }
    size_t ev_used = 0;
    auto next_event = [&](cudaEvent_t *e) -> int {
        if (ev_used == h->pipe_ev.size()) {
            cudaEvent_t x;
            CU(h, cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
            h->pipe_ev.push_back(x);
        }
        *e = h->pipe_ev[ev_used++];
        return 0;
    };
    // everything queued earlier on the compute stream (set_step precompute, ...) precedes the uploads
    cudaEvent_t e0 = nullptr;
    if (int rc = next_event(&e0)) return rc;
    CU(h, cudaEventRecord(e0, h->stream));
    CU(h, cudaStreamWaitEvent(h->s_up, e0, 0));
    auto before = [&](const std::vector<int> &part) -> int {
        for (int n : part)
            if (int rc = upload_one(h, n, prop[n], reference_prop ? reference_prop[n] : nullptr, h->s_up)) return rc;
        cudaEvent_t e = nullptr;
        if (int rc = next_event(&e)) return rc;
        CU(h, cudaEventRecord(e, h->s_up));
        CU(h, cudaStreamWaitEvent(h->stream, e, 0));
        return 0;
    };
    auto after = [&](const std::vector<int> &part) -> int {
        cudaEvent_t e = nullptr;
        if (int rc = next_event(&e)) return rc;
        CU(h, cudaEventRecord(e, h->stream));
        CU(h, cudaStreamWaitEvent(h->s_down, e, 0));
        for (int n : part)
            if (int rc = d2h3(h, prop[n], cur_ptr(h, n), 8, h->s_down)) return rc;
        return 0;
    };
    // which properties come with a reference field decides the kernel variant: known before the first upload
    for (int n = 0; n < *nprop; ++n) h->has_ref[n] = reference_prop && reference_prop[n];
    const int rc = step_once(h, b, chunk, before, after);
    // the caller's arrays are only borrowed for the call: drain all three streams, also when a launch failed midway
    const cudaError_t e1 = cudaStreamSynchronize(h->s_up), e2 = cudaStreamSynchronize(h->s_down),
                      e3 = cudaStreamSynchronize(h->stream);
    if (rc) return rc;
    CU(h, e1); CU(h, e2); CU(h, e3);
    return 0;
}

int mohid_adt_prop_device_ptr(const int *handle, const int *n, void **dptr, int *ld, int *nj, int *nk) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!n || !dptr) return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    CU(h, cudaSetDevice(h->dev));
    if (int rc = ensure_props(h, *n + 1, false)) return rc;
    *dptr = cur_ptr(h, *n);
    if (ld) *ld = h->ld;
    if (nj) *nj = h->njp;
    if (nk) *nk = h->nk;
    return 0;
}

// after writing a property through mohid_adt_prop_device_ptr: make the twin buffer identical
int mohid_adt_sync_prop_buffers(const int *handle, const int *nprop) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    CU(h, cudaSetDevice(h->dev));
    (void)nprop;          // one buffer per property since the in-place step: nothing to synchronise (kept for the ABI)
    return 0;
}

int mohid_adt_set_reference_device(const int *handle, const int *n, void **dptr) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    CU(h, cudaSetDevice(h->dev));
    if (int rc = ensure_props(h, *n + 1, true)) return rc;
    if (!h->ref[*n]) {
        if (int rc = dalloc(h, &h->ref[*n], h->n3)) return rc;
        CU(h, cudaMemsetAsync(h->ref[*n], 0, h->n3 * sizeof(double), h->stream));
    }
    h->has_ref[*n] = 1;
    *dptr = h->ref[*n];
    return 0;
}

static int pack_common(const int *handle, const int *nprop, const int *j0, const int *width, double *buf, int unpack) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!nprop || !j0 || !width || !buf) return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    if (*nprop > (int)h->prop.size() || *nprop > NPMAX) return fail(h, MOHID_ADT_ERR_STATE, "properties were never uploaded");
    if (*j0 < 0 || *width < 1 || *j0 + *width > h->nj) return fail(h, MOHID_ADT_ERR_ARG, "column range out of bounds");
    CU(h, cudaSetDevice(h->dev));
    PackArgs a{};
    a.ld = h->ld; a.nj = h->nj; a.nk = h->nk; a.nprop = *nprop; a.j0 = *j0; a.width = *width; a.sj = h->sj; a.sk = h->sk;
    for (int n = 0; n < *nprop; ++n) a.prop[n] = cur_ptr(h, n);
    const long tot = (long)a.nk * a.width * a.ld * a.nprop;
    const int blocks = (int)std::min<long>((tot + 255) / 256, (long)h->num_sms * 16);
    cudaStream_t st = h->comm ? h->comm : h->stream;
    if (h->comm && !unpack) CU(h, cudaStreamWaitEvent(h->comm, h->ev_edges, 0));     // the packed columns are final
    adt_pack_columns_kernel<<<blocks, 256, 0, st>>>(a, buf, unpack);
    CU(h, cudaGetLastError());
    if (h->comm && unpack) { CU(h, cudaEventRecord(h->ev_halo, h->comm)); h->halo_pending = true; }
    h->launches++;
    return 0;
}
int mohid_adt_pack_columns(const int *handle, const int *nprop, const int *j0, const int *width, void *device_buffer) {
    return pack_common(handle, nprop, j0, width, (double *)device_buffer, 0);
}
int mohid_adt_unpack_columns(const int *handle, const int *nprop, const int *j0, const int *width,
                             const void *device_buffer) {
    return pack_common(handle, nprop, j0, width, (double *)device_buffer, 1);
}

int mohid_adt_set_noflux(const int *handle, const int *NoFluxU, const int *NoFluxV, const int *NoFluxW) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    CU(h, cudaSetDevice(h->dev));
    if (!NoFluxU && !NoFluxV && !NoFluxW) { h->have_noflux = false; return 0; }
    if (!NoFluxU || !NoFluxV || !NoFluxW) return fail(h, MOHID_ADT_ERR_ARG, "NoFluxU, NoFluxV and NoFluxW must be given together");
    const int *src[3] = {NoFluxU, NoFluxV, NoFluxW};
    for (int a = 0; a < 3; ++a) {
        if (!h->noflux[a]) {
            if (int rc = dalloc(h, &h->noflux[a], h->n3)) return rc;
            CU(h, cudaMemsetAsync(h->noflux[a], 0, h->n3 * sizeof(int), h->stream));
        }
        if (int rc = h2d3(h, h->noflux[a], src[a], 4)) return rc;
    }
    if (!h->nfmask) if (int rc = dalloc(h, &h->nfmask, h->n3)) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    h->have_noflux = true;
    return 0;
}

int mohid_adt_set_premix(const int *handle, const double *Density, const double *WaterColumnZ, const double *SmallDepthsLimit) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    CU(h, cudaSetDevice(h->dev));
    if (WaterColumnZ && !SmallDepthsLimit) return fail(h, MOHID_ADT_ERR_ARG, "WaterColumnZ needs SmallDepthsLimit");
    if (h->premix_sd && !WaterColumnZ) h->have_small = false;       // the flag array was ours
    h->premix_fc = Density != nullptr;
    h->premix_sd = WaterColumnZ != nullptr;
    if (Density) {
        if (!h->density) if (int rc = dalloc(h, &h->density, h->n3)) return rc;
        if (int rc = h2d3(h, h->density, Density, 8)) return rc;
    }
    if (WaterColumnZ) {
        if (!h->wcol) if (int rc = dalloc(h, &h->wcol, h->n2)) return rc;
        if (int rc = h2d2(h, h->wcol, WaterColumnZ, 8)) return rc;
        h->sd_limit = *SmallDepthsLimit;
    }
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mohid_adt_set_offsets(const int *handle, const int *nprop, const double *OffSet) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!nprop || *nprop < 0 || *nprop > NPMAX || (*nprop > 0 && !OffSet)) return fail(h, MOHID_ADT_ERR_ARG, "bad offsets");
    h->offsets.assign(OffSet, OffSet + *nprop);
    return 0;
}

int mohid_adt_set_limits(const int *handle, const int *nprop, const int *MinOn, const double *MinValue, const int *MaxOn,
                         const double *MaxValue) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!nprop || *nprop < 0 || *nprop > NPMAX || (*nprop > 0 && (!MinOn || !MinValue || !MaxOn || !MaxValue)))
        return fail(h, MOHID_ADT_ERR_ARG, "bad limits");
    h->lim_min_on.assign(MinOn, MinOn + *nprop); h->lim_max_on.assign(MaxOn, MaxOn + *nprop);
    h->lim_min.assign(MinValue, MinValue + *nprop); h->lim_max.assign(MaxValue, MaxValue + *nprop);
    return 0;
}

int mohid_adt_get_limit_mass(const int *handle, const int *prop_index, double *Mass_Created, double *Mass_Destroid) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!prop_index || *prop_index < 0) return fail(h, MOHID_ADT_ERR_ARG, "bad property index");
    CU(h, cudaSetDevice(h->dev));
    const int n = *prop_index;
    double *src[2] = {n < (int)h->mass_created.size() ? h->mass_created[n] : nullptr,
                      n < (int)h->mass_destroyed.size() ? h->mass_destroyed[n] : nullptr};
    double *dst[2] = {Mass_Created, Mass_Destroid};
    for (int t = 0; t < 2; ++t) {
        if (!dst[t]) continue;
        if (!src[t]) return fail(h, MOHID_ADT_ERR_STATE, "no limits were applied to property %d", n);
        if (int rc = d2h3(h, dst[t], src[t], 8)) return rc;
    }
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mohid_adt_get_small_depths(const int *handle, int *SmallDepthsOn) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!SmallDepthsOn) return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    CU(h, cudaSetDevice(h->dev));
    CU(h, cudaMemcpy2DAsync(SmallDepthsOn, 4 * (size_t)h->ld_h, h->SmallDepths, 4 * (size_t)h->ld,
                            4 * (size_t)std::min(h->ld, h->ld_h), h->nj, cudaMemcpyDefault, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mohid_adt_get_cell_fluxes(const int *handle, const int *prop_index, double *AdvFluxX, double *AdvFluxY,
                              double *AdvFluxZ, double *DifFluxX, double *DifFluxY, double *DifFluxZ) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!prop_index) return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    const int n = *prop_index;
    if (n < 0 || n >= (int)h->flux[0].size() || !h->flux[0][n])
        return fail(h, MOHID_ADT_ERR_STATE, "GetAdvFlux - ModuleAdvectionDiffusion: no fluxes were computed for property %d (CellFluxes was not set)", n);
    CU(h, cudaSetDevice(h->dev));
    double *out[6] = {AdvFluxX, AdvFluxY, AdvFluxZ, DifFluxX, DifFluxY, DifFluxZ};
    for (int w = 0; w < 6; ++w)
        if (out[w]) if (int rc = d2h3(h, out[w], h->flux[w][n], 8)) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mohid_adt_set_boxes(const int *handle, const int *Boxes3D, const int *NumberOfBoxes3D) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!Boxes3D || !NumberOfBoxes3D || *NumberOfBoxes3D < 0 || *NumberOfBoxes3D > 4095)
        return fail(h, MOHID_ADT_ERR_ARG, "Boxes3D and 0 <= NumberOfBoxes3D <= 4095 are required");
    CU(h, cudaSetDevice(h->dev));
    if (!h->boxes) {
        if (int rc = dalloc(h, &h->boxes, h->n3)) return rc;
        CU(h, cudaMemsetAsync(h->boxes, 0xff, h->n3 * sizeof(int), h->stream));      // -1 outside the caller's rows
    }
    if (int rc = h2d3(h, h->boxes, Boxes3D, 4)) return rc;
    if (h->box_flux) { CU(h, cudaStreamSynchronize(h->stream)); cudaFree(h->box_flux); h->box_flux = nullptr; }
    h->nboxes = *NumberOfBoxes3D;
    if (int rc = dalloc(h, &h->box_flux, (size_t)(h->nboxes + 1) * (h->nboxes + 1))) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mohid_adt_box_fluxes(const int *handle, const int *prop_index, double *Fluxes3D) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!prop_index || !Fluxes3D) return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    if (!h->boxes) return fail(h, MOHID_ADT_ERR_STATE, "mohid_adt_set_boxes must precede box_fluxes");
    const int n = *prop_index;
    if (n < 0 || n >= (int)h->flux[0].size() || !h->flux[0][n])
        return fail(h, MOHID_ADT_ERR_STATE, "BoxDif - no fluxes were computed for property %d (CellFluxes was not set)", n);
    CU(h, cudaSetDevice(h->dev));
    const size_t cnt = (size_t)(h->nboxes + 1) * (h->nboxes + 1);
    CU(h, cudaMemsetAsync(h->box_flux, 0, cnt * sizeof(double), h->stream));           // Me%Fluxes3D(:,:) = 0.
    BoxArgs a{};
    a.I = h->I; a.J = h->J; a.K = h->K; a.ld = h->ld; a.sj = h->sj; a.sk = h->sk; a.nb1 = h->nboxes + 1;
    a.with_z = h->K > 1;
    a.Boxes = h->boxes; a.Water = h->raw_i[2]; a.Open = h->raw_i[0];
    a.ax = h->flux[0][n]; a.ay = h->flux[1][n]; a.az = h->flux[2][n];
    a.dx = h->flux[3][n]; a.dy = h->flux[4][n]; a.dz = h->flux[5][n];
    a.fluxes = h->box_flux;
    // only the owned columns of a slab: every face is then counted by exactly one rank (the host adds the matrices)
    const dim3 grid((unsigned)((h->I + 127) / 128), (unsigned)h->J, (unsigned)h->K);
    adt_box_flux_kernel<<<grid, 128, 0, h->stream>>>(a);
    CU(h, cudaGetLastError());
    h->launches++;
    CU(h, cudaMemcpyAsync(Fluxes3D, h->box_flux, cnt * sizeof(double), cudaMemcpyDefault, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mohid_adt_set_active_columns(const int *handle, const int *j_begin, const int *j_count) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!j_begin || !j_count || *j_begin < 1 || *j_count < 1 || *j_begin + *j_count - 1 > h->J)
        return fail(h, MOHID_ADT_ERR_ARG, "active column range must lie inside 1..J");
    CU(h, cudaSetDevice(h->dev));
    CU(h, cudaStreamSynchronize(h->stream));
    h->j_begin = *j_begin; h->j_count = *j_count;
    return upload_bnd_cols(h);
}

int mohid_adt_set_overlap(const int *handle, const int *ghost, void *comm_stream) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!ghost || *ghost < 0) return fail(h, MOHID_ADT_ERR_ARG, "bad ghost width");
    CU(h, cudaSetDevice(h->dev));
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->comm) CU(h, cudaStreamSynchronize(h->comm));
    h->halo_pending = false;
    // ghost > 0: pack / unpack run on `comm_stream`, ordered after the step by an event; the next step waits for the
    // unpack.  (Round 1 also advanced the edge columns first so that the exchange overlapped the interior; the
    // in-place step walks the columns in one direction, and the exchange was measured at < 3 % of a step.)
    h->comm = (*ghost > 0) ? (cudaStream_t)comm_stream : nullptr;
    if (h->comm && !h->ev_edges) {
        CU(h, cudaEventCreateWithFlags(&h->ev_edges, cudaEventDisableTiming));
        CU(h, cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
    }
    if (h->comm) CU(h, cudaEventRecord(h->ev_edges, h->stream));
    return 0;
}

// the compute stream waits for a halo exchange still running on the communication stream
int mohid_adt_join_halo(const int *handle) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    CU(h, cudaSetDevice(h->dev));
    if (h->halo_pending) { CU(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0)); h->halo_pending = false; }
    return 0;
}

// ---- NCCL halo exchange behind the C-ABI (replaces ReceiveSendProperitiesMPI, WP:15034-15045 -> HG:8479-8658) ----
int mohid_adt_comm_get_unique_id(void *unique_id, const int *nbytes) {
    if (!unique_id || !nbytes || *nbytes < (int)sizeof(ncclUniqueId))
        return fail(nullptr, MOHID_ADT_ERR_ARG, "unique_id buffer must hold %d bytes", (int)sizeof(ncclUniqueId));
    if (const char *e = load_nccl()) return fail(nullptr, MOHID_ADT_ERR_UNSUPPORTED, "NCCL: %s", e);
    ncclUniqueId id;
    const ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, MOHID_ADT_ERR_CUDA, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
    memcpy(unique_id, &id, sizeof id);
    return 0;
}

int mohid_adt_comm_init(const int *handle, const int *nranks, const int *rank, const void *unique_id, const int *ghost,
                        const int *overlap) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!nranks || !rank || !unique_id || !ghost || *nranks < 1 || *rank < 0 || *rank >= *nranks || *ghost < 1)
        return fail(h, MOHID_ADT_ERR_ARG, "bad communicator arguments");
    if (h->nccl) return fail(h, MOHID_ADT_ERR_STATE, "the handle already has a communicator");
    if (const char *e = load_nccl()) return fail(h, MOHID_ADT_ERR_UNSUPPORTED, "NCCL: %s", e);
    CU(h, cudaSetDevice(h->dev));
    // the slab must carry `ghost` columns on every interior side (set_active_columns)
    const int left = h->j_begin - 1, right = h->J - (h->j_begin + h->j_count - 1);
    if ((*rank > 0 && left != *ghost) || (*rank == 0 && left != 0) ||
        (*rank < *nranks - 1 && right != *ghost) || (*rank == *nranks - 1 && right != 0) || h->j_count < *ghost)
        return fail(h, MOHID_ADT_ERR_STATE,
                    "active columns %d..%d of 1..%d do not leave %d ghost columns on the interior sides of rank %d/%d",
                    h->j_begin, h->j_begin + h->j_count - 1, h->J, *ghost, *rank, *nranks);
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof id);
    const ncclResult_t r = g_nccl.CommInitRank(&h->nccl, *nranks, id, *rank);
    if (r != ncclSuccess) { h->nccl = nullptr; return fail(h, MOHID_ADT_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
    h->nranks = *nranks; h->rank = *rank; h->halo_ghost = *ghost;
    CU(h, cudaStreamCreateWithFlags(&h->s_comm_own, cudaStreamNonBlocking));
    const int ov = (overlap && *overlap) ? *ghost : 0;
    int rc = mohid_adt_set_overlap(handle, &ov, h->s_comm_own);
    if (rc) return rc;
    return 0;
}

int mohid_adt_exchange_halos(const int *handle, const int *nprop) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!h->nccl) return fail(h, MOHID_ADT_ERR_STATE, "mohid_adt_comm_init must precede exchange_halos");
    if (!nprop || *nprop < 1 || *nprop > (int)h->prop.size() || *nprop > NPMAX)
        return fail(h, MOHID_ADT_ERR_ARG, "bad property count");
    CU(h, cudaSetDevice(h->dev));
    const int g = h->halo_ghost;
    const size_t n = (size_t)h->ld * g * h->nk * (size_t)*nprop;
    if (h->halo_cap < n) {
        CU(h, cudaStreamSynchronize(h->s_comm_own));
        CU(h, cudaStreamSynchronize(h->stream));
        for (auto &p : h->halo_buf) { if (p) cudaFree(p); p = nullptr; }
        for (auto &p : h->halo_buf) if (int rc = dalloc(h, &p, n)) return rc;
        h->halo_cap = n;
    }
    const bool has_l = h->rank > 0, has_r = h->rank < h->nranks - 1;
    // without overlap the exchange is stream-ordered after the step: the communication stream waits for it
    cudaStream_t cs = h->comm ? h->comm : h->s_comm_own;
    if (!h->comm) {
        if (!h->ev_edges) CU(h, cudaEventCreateWithFlags(&h->ev_edges, cudaEventDisableTiming));
        if (!h->ev_halo) CU(h, cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
        CU(h, cudaEventRecord(h->ev_edges, h->stream));
        CU(h, cudaStreamWaitEvent(cs, h->ev_edges, 0));
    }
    auto pack = [&](int j0, double *buf, int unpack) -> int {
        if (h->comm) return unpack ? mohid_adt_unpack_columns(handle, nprop, &j0, &g, buf)
                                   : mohid_adt_pack_columns(handle, nprop, &j0, &g, buf);
        PackArgs a{};
        a.ld = h->ld; a.nj = h->nj; a.nk = h->nk; a.nprop = *nprop; a.j0 = j0; a.width = g; a.sj = h->sj; a.sk = h->sk;
        for (int m = 0; m < *nprop; ++m) a.prop[m] = cur_ptr(h, m);
        const int blocks = (int)std::min<long>(((long)n + 255) / 256, (long)h->num_sms * 16);
        adt_pack_columns_kernel<<<blocks, 256, 0, cs>>>(a, buf, unpack);
        CU(h, cudaGetLastError());
        h->launches++;
        return 0;
    };
    if (has_l) if (int rc = pack(h->j_begin, h->halo_buf[0], 0)) return rc;
    if (has_r) if (int rc = pack(h->j_begin + h->j_count - g, h->halo_buf[2], 0)) return rc;
    ncclResult_t r = g_nccl.GroupStart();
    if (r == ncclSuccess && has_l) r = g_nccl.Send(h->halo_buf[0], n, ncclDouble, h->rank - 1, h->nccl, cs);
    if (r == ncclSuccess && has_l) r = g_nccl.Recv(h->halo_buf[1], n, ncclDouble, h->rank - 1, h->nccl, cs);
    if (r == ncclSuccess && has_r) r = g_nccl.Send(h->halo_buf[2], n, ncclDouble, h->rank + 1, h->nccl, cs);
    if (r == ncclSuccess && has_r) r = g_nccl.Recv(h->halo_buf[3], n, ncclDouble, h->rank + 1, h->nccl, cs);
    const ncclResult_t r2 = g_nccl.GroupEnd();
    if (r != ncclSuccess || r2 != ncclSuccess)
        return fail(h, MOHID_ADT_ERR_CUDA, "NCCL halo exchange: %s", g_nccl.GetErrorString(r != ncclSuccess ? r : r2));
    if (has_l) if (int rc = pack(h->j_begin - g, h->halo_buf[1], 1)) return rc;
    if (has_r) if (int rc = pack(h->j_begin + h->j_count, h->halo_buf[3], 1)) return rc;
    if (!h->comm) {                                     // the next step (or a download) waits for the ghost columns
        CU(h, cudaEventRecord(h->ev_halo, cs));
        h->halo_pending = true;
    }
    return 0;
}

int mohid_adt_comm_destroy(const int *handle) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!h->nccl) return 0;
    CU(h, cudaSetDevice(h->dev));
    if (h->s_comm_own) CU(h, cudaStreamSynchronize(h->s_comm_own));
    CU(h, cudaStreamSynchronize(h->stream));
    const int zero = 0;
    mohid_adt_set_overlap(handle, &zero, nullptr);
    g_nccl.CommDestroy(h->nccl);
    h->nccl = nullptr;
    h->halo_pending = false;
    return 0;
}

int mohid_adt_synchronize(const int *handle) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    CU(h, cudaSetDevice(h->dev));
    if (h->comm) CU(h, cudaStreamSynchronize(h->comm));
    if (h->s_comm_own) CU(h, cudaStreamSynchronize(h->s_comm_own));
    h->halo_pending = false;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mohid_adt_solve_thomas_z(const int *handle, const double *D, const double *E, const double *F, const double *TI,
                             const int *WaterPoints3D, double *Res) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!D || !E || !F || !TI || !Res) return fail(h, MOHID_ADT_ERR_ARG, "null array");
    CU(h, cudaSetDevice(h->dev));
    double *d[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};      // D, E, F, TI, Res, W
    int *wat = nullptr;
    auto cleanup = [&]() { for (auto p : d) if (p) cudaFree(p); if (wat) cudaFree(wat); };
    const double *src[5] = {D, E, F, TI, Res};
    int rc = 0;
    for (int a = 0; a < 6 && !rc; ++a) {
        if (cudaMalloc((void **)&d[a], (size_t)(h->n3 + 64) * sizeof(double)) != cudaSuccess) {
            rc = fail(h, MOHID_ADT_ERR_CUDA, "out of device memory (solve_thomas_z)");
            break;
        }
        if (a < 5) rc = h2d3(h, d[a], src[a], 8);
    }
    if (!rc && WaterPoints3D) {
        if (cudaMalloc((void **)&wat, (size_t)(h->n3 + 64) * sizeof(int)) != cudaSuccess)
            rc = fail(h, MOHID_ADT_ERR_CUDA, "out of device memory (solve_thomas_z)");
        else rc = h2d3(h, wat, WaterPoints3D, 4);
    }
    if (!rc) {
        cudaMemsetAsync(h->d_zero_piv, 0, sizeof(unsigned long long), h->stream);
        const dim3 grid((unsigned)((h->I + 127) / 128), (unsigned)h->J);
        adt_thomas_z_kernel<<<grid, 128, 0, h->stream>>>(h->I, h->J, h->K, h->sj, h->sk, d[0], d[1], d[2], d[3], wat, d[4],
                                                       d[5], h->d_zero_piv);
        h->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = fail(h, MOHID_ADT_ERR_CUDA, "adt_thomas_z_kernel launch failed");
    }
    if (!rc) rc = d2h3(h, Res, d[4], 8);
    const cudaError_t e = cudaStreamSynchronize(h->stream);
    cleanup();
    if (rc) return rc;
    CU(h, e);
    return 0;
}

int mohid_adt_free_vertical_movement(const int *handle, const int *prop_index, const double *Velocity,
                                     const double *GridCellArea, const double *DepositionProbability, const int *Deposition,
                                     const int *NonCohesive, const int *DepositionIntertidalZones, const double *ImpExp_AdvV,
                                     const double *DTProp, double *FreeConvFlux) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!prop_index || !Velocity || !GridCellArea || !Deposition || !NonCohesive || !DepositionIntertidalZones ||
        !ImpExp_AdvV || !DTProp)
        return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    const int n = *prop_index;
    if (n < 0 || n >= (int)h->prop.size()) return fail(h, MOHID_ADT_ERR_STATE, "property %d was never uploaded", n);
    if (!h->have_step || !h->have_grid) return fail(h, MOHID_ADT_ERR_STATE, "set_grid2d and set_step must precede the call");
    if (!(*ImpExp_AdvV == 0.0 || *ImpExp_AdvV == 1.0))
        return fail(h, MOHID_ADT_ERR_ARG, "VerticalFreeConvection - ModuleFreeVerticalMovement - ERR04");
    if (*Deposition && !*NonCohesive && !DepositionProbability)
        return fail(h, MOHID_ADT_ERR_ARG, "DepositionProbability is required for a cohesive property that deposits");
    if (!(*DTProp > 0.0)) return fail(h, MOHID_ADT_ERR_ARG, "DTProp must be positive");
    CU(h, cudaSetDevice(h->dev));
    if (h->halo_pending) { CU(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0)); h->halo_pending = false; }
    // scratch: D, E, F, TI, W, velocity, flux (3-D), area, probability (2-D)
    double *d3[7] = {nullptr}, *d2[2] = {nullptr, nullptr};
    auto cleanup = [&]() { for (auto p : d3) if (p) cudaFree(p); for (auto p : d2) if (p) cudaFree(p); };
    int rc = 0;
    for (int a = 0; a < 7 && !rc; ++a)
        if ((a < 6 || FreeConvFlux) && cudaMalloc((void **)&d3[a], (size_t)(h->n3 + 64) * sizeof(double)) != cudaSuccess)
            rc = fail(h, MOHID_ADT_ERR_CUDA, "out of device memory (free_vertical_movement)");
    for (int a = 0; a < 2 && !rc; ++a)
        if ((a == 0 || DepositionProbability) && cudaMalloc((void **)&d2[a], (size_t)(h->n2 + 64) * sizeof(double)) != cudaSuccess)
            rc = fail(h, MOHID_ADT_ERR_CUDA, "out of device memory (free_vertical_movement)");
    if (!rc) rc = h2d3(h, d3[5], Velocity, 8);
    if (!rc) rc = h2d2(h, d2[0], GridCellArea, 8);
    if (!rc && DepositionProbability) rc = h2d2(h, d2[1], DepositionProbability, 8);
    if (!rc) {
        FvmArgs a{};
        a.I = h->I; a.J = h->J; a.K = h->K; a.sj = h->sj; a.sk = h->sk; a.ld = h->ld;
        a.Mask = *DepositionIntertidalZones ? h->raw_i[2] : h->raw_i[0];
        a.Land = h->raw_i[1]; a.KFloorZ = h->KFloorZ; a.VolumeZ = h->raw_d[4];
        a.Velocity = d3[5]; a.Area = d2[0]; a.DepProb = d2[1];
        a.deposition = *Deposition; a.non_cohesive = *NonCohesive; a.dt = *DTProp; a.impexp = *ImpExp_AdvV;
        a.C = cur_ptr(h, n);
        a.D = d3[0]; a.E = d3[1]; a.F = d3[2]; a.TI = d3[3]; a.flux = FreeConvFlux ? d3[6] : nullptr;
        adt_fvm_coef_kernel<<<dim3((unsigned)((h->ld + 127) / 128), (unsigned)h->nj, (unsigned)h->nk), 128, 0, h->stream>>>(a);
        h->launches++;
        if (*ImpExp_AdvV != 1.0) {
            cudaMemsetAsync(h->d_zero_piv, 0, sizeof(unsigned long long), h->stream);
            adt_thomas_z_kernel<<<dim3((unsigned)((h->I + 127) / 128), (unsigned)h->J), 128, 0, h->stream>>>(
                h->I, h->J, h->K, h->sj, h->sk, d3[0], d3[1], d3[2], d3[3], h->raw_i[2], cur_ptr(h, n), d3[4], h->d_zero_piv);
            h->launches++;
        } else if (copy_field(h, cur_ptr(h, n), d3[3])) {
            rc = MOHID_ADT_ERR_CUDA;
        }
        if (!rc && FreeConvFlux) {
            adt_fvm_flux_kernel<<<dim3((unsigned)((h->I + 127) / 128), (unsigned)h->J, (unsigned)h->K), 128, 0, h->stream>>>(a);
            h->launches++;
        }
        if (!rc && cudaGetLastError() != cudaSuccess) rc = fail(h, MOHID_ADT_ERR_CUDA, "free_vertical_movement launch failed");
    }
    if (!rc && FreeConvFlux) rc = d2h3(h, FreeConvFlux, d3[6], 8);
    const cudaError_t e = cudaStreamSynchronize(h->stream);
    cleanup();
    if (rc) return rc;
    CU(h, e);
    return 0;
}

// ---- ModuleHydroIntegration on the device mirrors ----
static void hint_args(Handle *h, HintArgs &a) {
    a.I = h->I; a.J = h->J; a.K = h->K; a.ni = h->ni; a.nj = h->nj; a.ld = h->ld; a.sj = h->sj; a.sk = h->sk;
    a.WX = h->raw_d[0]; a.WY = h->raw_d[1]; a.WZ = h->raw_d[2]; a.D = h->hint_disch;
    a.CFU = h->raw_i[3]; a.CFV = h->raw_i[4]; a.CFW = h->raw_i[5]; a.Open = h->raw_i[0];
    a.Water = h->raw_i[2]; a.Bnd = h->Bnd; a.V = h->raw_d[4]; a.V0 = h->raw_d[3];
}

int mohid_adt_hydro_integration_reinit(const int *handle, const double *VolumeZOld) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!VolumeZOld) return fail(h, MOHID_ADT_ERR_ARG, "null array");
    CU(h, cudaSetDevice(h->dev));
    // ReInitalizeIntegration (HI:767-796): n = 0, InitialVolume = VolumeZOld, fluxes / discharges / mapping = 0
    if (int rc = h2d3(h, h->raw_d[3], VolumeZOld, 8)) return rc;
    for (int a : {0, 1, 2}) CU(h, cudaMemsetAsync(h->raw_d[a], 0, h->n3 * sizeof(double), h->stream));
    for (int a : {0, 3, 4, 5}) CU(h, cudaMemsetAsync(h->raw_i[a], 0, h->n3 * sizeof(int), h->stream));
    if (h->hint_disch) CU(h, cudaMemsetAsync(h->hint_disch, 0, h->n3 * sizeof(double), h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    h->hint_n = 0;
    return 0;
}

int mohid_adt_hydro_integration_step(const int *handle, const double *WaterFluxX, const double *WaterFluxY,
                                     const double *Discharges, const int *ComputeFacesU, const int *ComputeFacesV) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!WaterFluxX || !WaterFluxY || !ComputeFacesU || !ComputeFacesV) return fail(h, MOHID_ADT_ERR_ARG, "null array");
    if (h->hint_n < 0) return fail(h, MOHID_ADT_ERR_STATE, "hydro_integration_reinit must precede hydro_integration_step");
    CU(h, cudaSetDevice(h->dev));
    for (auto &p : h->hint_in) if (!p) if (int rc = dalloc(h, &p, h->n3)) return rc;
    for (auto &p : h->hint_cf) if (!p) if (int rc = dalloc(h, &p, h->n3)) return rc;
    if (Discharges && !h->hint_disch) {
        if (h->hint_n > 0) return fail(h, MOHID_ADT_ERR_STATE, "Discharges must be passed from the first step of an integration on");
        if (int rc = dalloc(h, &h->hint_disch, h->n3)) return rc;
        CU(h, cudaMemsetAsync(h->hint_disch, 0, h->n3 * sizeof(double), h->stream));
    }
    int rc = h2d3(h, h->hint_in[0], WaterFluxX, 8);
    if (!rc) rc = h2d3(h, h->hint_in[1], WaterFluxY, 8);
    if (!rc && Discharges) rc = h2d3(h, h->hint_in[2], Discharges, 8);
    if (!rc) rc = h2d3(h, h->hint_cf[0], ComputeFacesU, 4);
    if (!rc) rc = h2d3(h, h->hint_cf[1], ComputeFacesV, 4);
    if (rc) return rc;
    HintArgs a{};
    hint_args(h, a);
    a.n = ++h->hint_n;
    a.fx = h->hint_in[0]; a.fy = h->hint_in[1]; a.disch = Discharges ? h->hint_in[2] : nullptr;
    a.cfu = h->hint_cf[0]; a.cfv = h->hint_cf[1];
    adt_hint_step_kernel<<<dim3((unsigned)((h->ni + 127) / 128), (unsigned)h->nj, (unsigned)h->nk), 128, 0, h->stream>>>(a);
    CU(h, cudaGetLastError());
    h->launches++;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mohid_adt_hydro_integration_end(const int *handle, const double *VolumeZ, const int *WaterPoints3D, const double *DT) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!VolumeZ || !WaterPoints3D || !DT || !(*DT > 0.)) return fail(h, MOHID_ADT_ERR_ARG, "VolumeZ, WaterPoints3D and DT > 0 are required");
    if (h->hint_n < 1) return fail(h, MOHID_ADT_ERR_STATE, "no hydrodynamic step was integrated");
    if (!h->have_grid) return fail(h, MOHID_ADT_ERR_STATE, "set_grid2d must precede hydro_integration_end (BoundaryPoints2D)");
    CU(h, cudaSetDevice(h->dev));
    if (int rc = h2d3(h, h->raw_d[4], VolumeZ, 8)) return rc;
    if (int rc = h2d3(h, h->raw_i[2], WaterPoints3D, 4)) return rc;
    HintArgs a{};
    hint_args(h, a);
    a.dt = *DT;
    adt_hint_end_kernel<<<dim3((unsigned)((h->I + 127) / 128), (unsigned)h->J), 128, 0, h->stream>>>(a);
    CU(h, cudaGetLastError());
    h->launches++;
    CU(h, cudaStreamSynchronize(h->stream));
    h->hint_n = -1;
    return 0;
}

int mohid_adt_download_step_input(const int *handle, const int *which, void *array) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!which || !array || *which < 0 || *which > 16) return fail(h, MOHID_ADT_ERR_ARG, "which must be 0..16 and the array present");
    CU(h, cudaSetDevice(h->dev));
    const int w = *which;
    if (int rc = w < 11 ? d2h3(h, array, h->raw_d[w], 8) : d2h3(h, array, h->raw_i[w - 11], 4)) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int mohid_adt_column_mass(const int *handle, const int *nprop, double *mass) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!nprop || !mass || *nprop < 1 || *nprop > (int)h->prop.size() || *nprop > NPMAX)
        return fail(h, MOHID_ADT_ERR_ARG, "bad property count");
    if (!h->have_step) return fail(h, MOHID_ADT_ERR_STATE, "set_step must precede column_mass");
    CU(h, cudaSetDevice(h->dev));
    if (h->halo_pending) { CU(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0)); h->halo_pending = false; }
    double *out = nullptr;
    const size_t cnt = (size_t)*nprop * h->nj;
    CU(h, cudaMalloc((void **)&out, cnt * sizeof(double)));
    MassArgs a{};
    a.I = h->I; a.K = h->K; a.sj = h->sj; a.sk = h->sk; a.nj = h->nj;
    a.Water = h->raw_i[2]; a.VolumeZ = h->raw_d[4]; a.out = out;
    for (int n = 0; n < *nprop; ++n) a.prop[n] = cur_ptr(h, n);
    adt_column_mass_kernel<<<dim3((unsigned)h->nj, (unsigned)*nprop), 256, 0, h->stream>>>(a);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(mass, out, cnt * sizeof(double), cudaMemcpyDefault, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(out);
    CU(h, e);
    return 0;
}

int mohid_adt_get_counters(const int *handle, long long *counters, const int *n) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    if (!counters || !n) return fail(h, MOHID_ADT_ERR_ARG, "null argument");
    CU(h, cudaSetDevice(h->dev));
    CU(h, cudaStreamSynchronize(h->stream));
    unsigned long long zp = 0;
    CU(h, cudaMemcpy(&zp, h->d_zero_piv, sizeof zp, cudaMemcpyDeviceToHost));
    long long v[4] = {h->launches, (long long)zp, 0, h->bytes};
    for (int i = 0; i < *n && i < 4; ++i) counters[i] = v[i];
    return 0;
}

int mohid_adt_kernel_time_ms(const int *handle, double *ms, int *launches) {
    Handle *h = get(handle);
    if (!h) return fail(nullptr, MOHID_ADT_ERR_HANDLE, "bad handle");
    CU(h, cudaSetDevice(h->dev));
    CU(h, cudaStreamSynchronize(h->stream));
    double tot = 0.;
    for (size_t i = 0; i < h->ev_used; ++i) {
        float t = 0.f;
        CU(h, cudaEventElapsedTime(&t, h->ev[i].first, h->ev[i].second));
        tot += t;
    }
    // the chunk launches of a step add up to one kernel time
    if (ms) *ms = h->ev_steps ? tot / (double)h->ev_steps : 0.;
    if (launches) *launches = h->ev_steps;
    h->ev_used = 0;
    h->ev_steps = 0;
    return 0;
}

}  // extern "C"
