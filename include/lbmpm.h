/*
 * lbmpm.h -- C ABI of liblbmpm.so: the B200-native collision + streaming hot path of
 * openLBMPM (multiphase lattice Boltzmann, D2Q9 and D3Q19).
 *
 * This is the drop-in boundary.  The reference has no native layer: its host
 * drivers call Numba-CUDA kernels as `K[grid, block](scalars..., device arrays...)`
 * from a per-step Python loop.  Each entry point below replaces one block of that
 * driver code; the citation after every prototype is the reference code it stands
 * in for (paths relative to the reference repository, commit 3d84189).
 *
 * Conventions
 *   - every function returns 0 on success or a negative LBM_E* code, and never
 *     throws; `lbm_last_error` gives the message of the last failure on a handle
 *     (or of the last failed `lbm_create` when called with NULL);
 *   - host buffers are owned by the caller, C-contiguous, float64 / int64 / uint8;
 *     device buffers are owned by the handle;
 *   - dense arrays are `[z][y][x]` row-major (2-D: nz = 1, i.e. the reference's
 *     `[y][x]`); population arrays are the reference's AoS `[z][y][x][Q]`
 *     (RKD2Q9.py:451-452); the SoA <-> AoS transposition happens in
 *     upload/download, never on the timed path;
 *   - one host thread per handle; all device work is stream-ordered on a
 *     handle-owned stream; `lbm_step` is asynchronous, downloads synchronise;
 *   - there is NO CPU fallback: every entry point that computes needs the GPU.
 */
#ifndef LBMPM_H
#define LBMPM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_ABI_VERSION 1

/* error codes */
#define LBM_OK          0
#define LBM_EINVAL     -1   /* bad argument / unsupported combination            */
#define LBM_ECUDA      -2   /* a CUDA runtime call failed                        */
#define LBM_ESTATE     -3   /* call order (e.g. step before geometry/state)      */
#define LBM_ENOMEM     -4
#define LBM_ENCCL      -5

/* lbm_config.model */
#define LBM_MODEL_CG    0   /* Rothman-Keller colour gradient, CSF   (RKCG2D/)          */
#define LBM_MODEL_SC    1   /* original Shan-Chen                    (ShanChen2D/Optimized...)  */
#define LBM_MODEL_EFS   2   /* explicit-forcing Shan-Chen SRT/MRT    (ShanChen2D/Explicit...)   */

/* lbm_config.surface_tension_type (LBM_MODEL_CG) */
#define LBM_ST_CSF           0   /* continuum surface force: runRKColorGradient2DCSF (RKD2Q9.py:1225-1490)          */
#define LBM_ST_PERTURBATION  1   /* perturbation operator: runRKColorGradient2DPerturbation (RKD2Q9.py:979-1219; kernels
                                    calRKCollision1GPU2D{MRT,SRT}New, calRKCollision23GPUNew).  Open boundaries as in the
                                    CSF loop; SRT has no body-force term (the reference's SRT kernel has none)          */

/* lbm_config.relax */
#define LBM_RELAX_SRT   0
#define LBM_RELAX_MRT   1

/* lbm_config.inlet  (top of the flow axis: y in 2-D, z in 3-D) */
#define LBM_BC_PERIODIC          0
#define LBM_INLET_VELOCITY       1   /* ini: BoundaryTypeInlet = 'Neumann'   */
#define LBM_INLET_PRESSURE       2   /* ini: BoundaryTypeInlet = 'Dirichlet' */
/* lbm_config.outlet (bottom of the flow axis) */
#define LBM_OUTLET_CONVECTIVE    1   /* ini: BoundaryTypeOutlet = 'Convective' */
#define LBM_OUTLET_PRESSURE      2   /* ini: BoundaryTypeOutlet = 'Dirichlet'  */

/* lbm_config.flags */
#define LBM_FLAG_GENERIC_KERNELS 1u  /* force the unfused reference-ordered kernel sequence */
#define LBM_FLAG_NO_TILED_KERNEL 2u  /* factored fast path, but with its one-thread-per-node kernels only */
#define LBM_FLAG_NO_TILED_DENSITY 4u /* keep the tiled collision pass, use the one-thread-per-node density pass */
#define LBM_FLAG_OVERLAP 16u         /* slab decomposition: overlap the ghost-plane exchanges with the interior planes */
#define LBM_FLAG_PACKED_EXCHANGE 32u /* slab decomposition: pack the boundary planes into one message per direction (experimental) */
#define LBM_FLAG_GHOST_PLANES 64u     /* single slab: keep copying the periodic ghost planes every step instead of wrapping the flow axis by index arithmetic */
#define LBM_FLAG_PERSISTENT 128u      /* single slab, one-thread-per-node fast path: all steps of an lbm_step call in ONE cooperative
                                        kernel (one grid-wide barrier between the phases of a step instead of a launch); opt-in */
#define LBM_FLAG_PEER_EXCHANGE 256u    /* slab decomposition, factored fast path: the ghost planes are STORED into the neighbours' memory
                                         (CUDA IPC peer pointers over NVLink) and a release / acquire flag pair replaces the
                                         NCCL send / recv rendezvous of the two per-step exchanges; with the tiled kernels the stores
                                         are issued by the collision / density passes themselves.  This is the DEFAULT on slabs of
                                         equal extents whose GPUs have peer access (measured on 2 B200s, 512^3: 0.46 -> 0.07 ms of
                                         exchange per step); setting the flag makes its absence an error instead of a fallback */
#define LBM_FLAG_NCCL_EXCHANGE 512u    /* slab decomposition: keep ncclSend / ncclRecv for the fast path's exchanges as well */
#define LBM_FLAG_NO_CUDA_GRAPH 8u    /* small lattices: launch every kernel instead of replaying a captured graph */

typedef struct lbm_handle lbm_handle;

/* Plain-old-data configuration: the numbers the reference reads from the .ini files of IniFiles/
 * (RKD2Q9.py:24-297, ShanChenD2Q9.py:39-496).  Zero-initialise, then fill.        */
typedef struct lbm_config {
    int32_t abi_version;       /* = LBM_ABI_VERSION                                         */
    int32_t lattice;           /* 9 (D2Q9) or 19 (D3Q19)                                    */
    int32_t model;             /* LBM_MODEL_*                                               */
    int32_t nx, ny, nz;        /* xDomain, yDomain, zDomain (nz = 1 for D2Q9)               */
    int32_t relax;             /* [RelaxationType] Type                                     */
    int32_t tau_type;          /* [FluidParameters] TauType 1|2        (CG)                 */
    int32_t wetting_type;      /* [SurfaceTension] WettingType 1|2     (CG)                 */
    int32_t inlet, outlet;     /* LBM_BC_* / LBM_INLET_* / LBM_OUTLET_*                     */
    int32_t device;            /* CUDA device ordinal                                       */
    uint32_t flags;
    int32_t n_components;      /* SC/EFS: number of fluids (1..4); CG: ignored (2 colours)  */
    int32_t sc_isotropy;       /* EFS: [ForceScheme] ExplicitScheme 4 | 8 | 10 (0 = 4)      */
    int32_t surface_tension_type; /* CG: [SurfaceTension] SurfaceTensionType, LBM_ST_*      */
    int32_t reserved_i[1];
    /* colour gradient */
    double sigma;              /* [SurfaceTension] SurfaceTension(Value)                    */
    double contact_angle_deg;  /* [SurfaceTension] ContactAngle                             */
    double beta;               /* [RKParameters] BetaThickness                              */
    double delta;              /* [RKParameters] DeltaValue                                 */
    double tauR, tauB;         /* [FluidParameters] TauR, TauB                              */
    double inlet_velocity;     /* velocityYR + velocityYB (RKD2Q9.py:1300)                  */
    double rhoBH, rhoRH;       /* densityBH, densityRH  (pressure inlet)                    */
    double rhoBL, rhoRL;       /* densityBL, densityRL  (pressure outlet)                   */
    /* Shan-Chen / explicit forcing (per component, up to 4) */
    double sc_tau[4];          /* [FluidProperties] tau                                     */
    double sc_G[16];           /* interaction matrix G[s][s'] row-major (n x n used)        */
    double sc_Gsolid[4];       /* fluid-solid interaction strengths                         */
    double sc_inlet_velocity[4];
    double sc_rho_in[4], sc_rho_out[4];
    /* colour gradient with the perturbation operator (LBM_ST_PERTURBATION) */
    double AkR, AkB;           /* [RKParameters] AkR, AkB                                   */
    double solid_phi;          /* [SolidBoundarySetup] SolidColorDiff: phi seen on solid neighbours */
    double body_force[3];      /* [BodyForce] bodyForceX, bodyForceY (, Z)                  */
    double reserved_d[2];
} lbm_config;

/* ---- lifetime -------------------------------------------------------------------------- */

/* Replaces the constructor side of the drivers: constants, MRT matrices and all
 * `cuda.to_device` allocations (RKD2Q9.py:299-340, 1243-1287; ShanChenD2Q9.py:1447-1485). */
int lbm_create(const lbm_config* cfg, lbm_handle** out);
int lbm_destroy(lbm_handle* h);
const char* lbm_last_error(const lbm_handle* h);
int lbm_abi_version(void);

/* ---- geometry and indexing -------------------------------------------------------------- */

/* is_domain: uint8 dense [nz][ny][nx], 1 = void (fluid), 0 = solid -- the `isDomain` array
 * produced by SimpleGeometry.defineGeometry / the structure image (RKD2Q9.py:417-443).
 * Builds on the device everything the reference builds with Python loops:
 * node classes, wetting-solid set, fluid nodes next to solid and the unit normals n_s
 * (RKD2Q9.py:657-892).                                                                     */
int lbm_set_geometry(lbm_handle* h, const uint8_t* is_domain);

/* Sizes of the reference's index structures for the current geometry:
 * n_fluid = fluidNodes.size, n_wet_solid = wettingSolidNodes.size,
 * n_fluid_near_solid = fluidNodesWithSolidGPU.size.                                        */
int lbm_index_sizes(lbm_handle* h, int64_t* n_fluid, int64_t* n_wet_solid, int64_t* n_fluid_near_solid);

/* Bit-exact export of the reference's integer structures (any pointer may be NULL):
 *   fluid_nodes[n_fluid]                 fluidNodes                     RKD2Q9.py:665-676
 *   neighbors[8|18 * n_fluid]            neighboringNodes               AcceleratedRKGPU2D.py:14-52
 *   wet_solid_nodes[n_wet_solid]         wettingSolidNodes              RKD2Q9.py:677-689
 *   wet_solid_neighbors[8|18 * n_wet]    neighboringWettingSolidNodes   AcceleratedRKGPU2D.py:57-95
 *   near_solid_compact[n_near]           fluidNodesWithSolidGPU         RKD2Q9.py:741-760
 *   near_solid_flat[n_near]              fluidNodesWithSolidOriginal
 *   ns[D * n_near]                       nsX, nsY(, nsZ) component-major RKD2Q9.py:768-892
 * Compact ids: >= 0 fluid rank, -1 solid away from fluid, <= -2 wetting solid (-2 - rank).  */
int lbm_export_indexing(lbm_handle* h, int64_t* fluid_nodes, int64_t* neighbors,
                        int64_t* wet_solid_nodes, int64_t* wet_solid_neighbors,
                        int64_t* near_solid_compact, int64_t* near_solid_flat, double* ns);

/* ---- state ------------------------------------------------------------------------------ */

/* f = w_i * rho at rest for every void node: the reference's initial condition
 * (RKD2Q9.py:561-585 with zero velocity; ShanChenD2Q9.py:734-786).
 * CG: rho[0] = rhoR, rho[1] = rhoB; SC/EFS: one dense array per component.
 * rho: n_comp pointers to dense [nz][ny][nx].  Resets the lagged force to zero.            */
int lbm_init_equilibrium(lbm_handle* h, const double* const* rho, int32_t n_comp);

/* Upload populations in the reference's dense AoS layout [nz][ny][nx][Q] per component
 * (fluidPDFR / fluidPDFB, RKD2Q9.py:451-452; fluidPDF[nf], ShanChenD2Q9.py:740) and the
 * matching densities (NULL entries: computed as the population sums) -- the restart path
 * of RKD2Q9.py:491-559.                                                                    */
int lbm_upload_state(lbm_handle* h, const double* const* pdf, const double* const* rho, int32_t n_comp);

/* ---- the hot path ----------------------------------------------------------------------- */

/* Advance nsteps iterations of the reference's per-step loop
 * (CG: RKD2Q9.py:1295-1490; SC: ShanChenD2Q9.py:1492-1629; EFS: ShanChenD2Q9.py:1852-2087),
 * asynchronously on the handle's stream.                                                   */
int lbm_step(lbm_handle* h, int32_t nsteps);
int lbm_synchronize(lbm_handle* h);

/* ---- results ---------------------------------------------------------------------------- */

/* What the reference copies to the host at an output interval (RKD2Q9.py:1382-1393,
 * convertOptTo2D :902-911): densities after this step's boundary treatment and the
 * velocity evaluated with the lagged force.  Dense [nz][ny][nx], zero on solid nodes.
 * rho: n_comp pointers; u: up to 3 pointers (x, y, z); any pointer may be NULL.            */
int lbm_download_macros(lbm_handle* h, double* const* rho, int32_t n_comp, double* const* u);

/* Asynchronous form of lbm_download_macros for output that must not stall the step loop (SURVEY.md section 8, row f-4;
 * the reference blocks in copy_to_host every TimeInterval steps, RKD2Q9.py:1382-1393): the fields of the output point
 * are snapshot into a device staging buffer on the compute stream and copied to the host arrays on a second stream.
 * Returns at once; lbm_step may be called right away.  The host arrays must stay valid until lbm_output_wait returns
 * and should come from lbm_host_alloc (page-locked), otherwise the copy degenerates to a synchronous one.
 * One download in flight per handle: a second call waits (on the device) for the previous copy.              */
int lbm_download_macros_async(lbm_handle* h, double* const* rho, int32_t n_comp, double* const* u);
/* Blocks the calling host thread until the copies of the last lbm_download_macros_async have landed.  The only entry
 * point that may be called from a second host thread (an output writer) while the owner thread keeps stepping.  */
int lbm_output_wait(lbm_handle* h);
/* Page-locked host memory for the asynchronous output (cudaHostAlloc / cudaFreeHost).                          */
int lbm_host_alloc(void** ptr, int64_t bytes);
int lbm_host_free(void* ptr);

/* Populations in the reference's dense AoS layout [nz][ny][nx][Q] per component
 * (fluidPDFR/B at the same output point).                                                  */
int lbm_download_pdfs(lbm_handle* h, double* const* pdf, int32_t n_comp);

/* Auxiliary per-node fields at the output point, dense [nz][ny][nx] (NULL = skip): phi (colour field incl.
 * wetting-solid values), G (D arrays, after the wetting correction) and curvature K of the current time level,
 * F (D arrays) = the CSF force of the PREVIOUS step (the lagged force the velocity is evaluated with).
 * LBM_ST_PERTURBATION: phi of the output point, G of the last collision, K = F = 0 (the model has neither).    */
int lbm_download_fields(lbm_handle* h, double* phi, double* const* G, double* const* F, double* K);

/* Sum of each component's density over the void nodes (mass check).                        */
int lbm_total_mass(lbm_handle* h, double* mass, int32_t n_comp);

/* Checksum of the densities at the output point: per component the wrap-around uint64 sum of the bit patterns of rho over
 * the slab's own nodes.  Sums of the slabs of a decomposed lattice add up (mod 2^64) to the sum of the undivided lattice
 * iff the slabs are bit-equal to it -- bench.py prints the all-reduced value at every GPU count.  (No reference
 * counterpart: the reference is single-GPU, RKD2Q9.py holds the whole lattice on one device.)                  */
int lbm_state_checksum(lbm_handle* h, uint64_t* sums, int32_t n_comp);

/* ---- solute tracers riding on the colour-gradient CSF flow ------------------ */

/* The numbers the reference reads from transportsetup.ini (Transport2DRK.py:96-391).                           */
typedef struct lbm_tracer_config {
    int32_t n_tracers;         /* [TransportParameters] NumberTracers (1..4)                                    */
    int32_t relax;             /* [RelaxationType] Relaxation: LBM_RELAX_SRT | LBM_RELAX_MRT (MRT: D2Q9 only)   */
    double tau[4];             /* [TransportParameters] Tau                      (SRT)                          */
    double dxx[4], dyy[4];     /* [TransportMRT] DiffusionX, DiffusionY          (MRT)                          */
    double dxy[4], dyx[4];     /* [TransportMRT] DiffusionXY, DiffusionYX        (MRT)                          */
    double beta[4];            /* [TransportParameters] BetaInterface                                           */
    double criterion;          /* rho_R above which a node is outside the transport domain (0.5, :1167)         */
    /* 5-velocity branch of the same driver (Transport2DRK.py:1344-1384), D2Q9 flow only, MRT only                */
    int32_t n_schemes;         /* [SystemType] NumberSchemes: 9 (0 means 9) | 5                                 */
    int32_t reaction;          /* [SystemType] Reaction: 1 = A + B -> C on tracers 0, 1, 2 (calReactionTracersGPU) */
    int32_t inlet_type;        /* [BoundaryCondition] InletType: LBM_TR_NONE | LBM_TR_INLET_DIRICHLET           */
    int32_t outlet_type;       /* [BoundaryCondition] OutletType: LBM_TR_NONE | LBM_TR_OUTLET_FREEFLOW          */
    double reaction_rate;      /* [Reaction] ReactionRate (the first one; the kernel reads no other)            */
    double diff_j[4];          /* [TransportParameters] DiffusionJ: rest share J_0 of the reaction source        */
    double inlet_conc[4];      /* concentration held on the inlet row (Inamuro, calInamuroConstConcBoundary)     */
} lbm_tracer_config;
#define LBM_TR_NONE             0
#define LBM_TR_INLET_DIRICHLET  1
#define LBM_TR_OUTLET_FREEFLOW  1

/* Attaches tracers to a colour-gradient CSF handle (9-velocity tracers: closed boxes; 5-velocity tracers bring their own
 * inlet / outlet rows and also ride on open channels).  Must precede lbm_init_equilibrium /
 * lbm_upload_state: the reference's transport loop STARTS with the streaming of the flow (Transport2DRK.py:1180-1200),
 * so a freshly set flow state is streamed once.  From then on every lbm_step iteration runs, between the colour
 * gradient and the flow collision, the tracer collision + interface term + streaming + concentration
 * (Transport2DRK.py:1341-1425 with the kernels of AccelerateTransport2DRK.py).                                  */
int lbm_tracer_setup(lbm_handle* h, const lbm_tracer_config* cfg);
/* g = w_j C at rest (Transport2DRK.py:461-470); conc: n pointers to dense [nz][ny][nx]; after the flow state.  */
int lbm_tracer_init(lbm_handle* h, const double* const* conc, int32_t n);
/* Concentrations at the output point of the current iteration (Transport2DRK.py:1427-1437), zero on solids.    */
int lbm_tracer_download(lbm_handle* h, double* const* conc, int32_t n);

/* ---- measurement ------------------------------------------------------------------------ */

/* CUDA-event time of the last lbm_step call (ms, on the handle's stream), number of kernel
 * launches it issued and number of void nodes it updated per step.                         */
int lbm_get_timing(lbm_handle* h, double* last_step_call_ms, int64_t* kernel_launches, int64_t* nodes_per_step);

/* Per-kernel timing for the roofline leg of bench.py: while enabled, every kernel launch is bracketed by
 * CUDA events on the handle's stream; the report is one line per kernel, "name<TAB>launches<TAB>total_ms". */
int lbm_profile_enable(lbm_handle* h, int32_t on);
int lbm_profile_report(lbm_handle* h, char* buf, int64_t buflen);

/* Device-resident benchmark initialiser: spinodal start rhoR = 0.5 + amp*(U-0.5), rhoB = 1-rhoR
 * with a counter-based hash of (seed, node id) -- no host buffers involved.                */
int lbm_init_spinodal_device(lbm_handle* h, double amplitude, uint64_t seed);

/* ---- multi-GPU slab decomposition (one handle per rank, one rank per GPU) --------------- */

/* Fills 128 bytes with an NCCL unique id (rank 0 calls it, the host framework broadcasts it). */
int lbm_nccl_unique_id(uint8_t id_out[128]);
/* Declares this handle to be slab `rank` of `nranks` along the flow axis (z in 3-D, y in 2-D):
 * cfg.nz (ny) is the LOCAL slab thickness; ghost planes are exchanged every step with
 * ncclSend/ncclRecv to rank+-1 (periodic ring).  Must precede lbm_set_geometry.              */
int lbm_comm_init(lbm_handle* h, int32_t rank, int32_t nranks, const uint8_t id[128]);

#ifdef __cplusplus
}
#endif
#endif /* LBMPM_H */
