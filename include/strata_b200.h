/*
 * strata_b200.h -- C ABI of the B200 backend for strata-fdtd's time-stepping hot path.
 *
 * This is the drop-in boundary.  It replaces, for backend="b200", the calls that
 * the reference's FDTDSolver makes into its pybind11 module `_kernels`
 * (/root/reference/src/strata_fdtd/_kernels/kernels.cpp) plus the per-step Python
 * hooks around them (core/solver.py:2003-2077).  Each entry point below cites the
 * reference interface it stands in for.  No C++ or torch types cross the boundary:
 * plain pointers, sizes and an opaque handle.  One handle drives one GPU (one
 * slab of the grid along reference axis 0).
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; sb_last_error()
 *    returns a thread-local message for the last failing call.
 *  - "device pointer" arguments are CUDA device addresses owned by the caller
 *    (the Python host carries them as torch tensors); "host" arguments are
 *    ordinary host memory.  The library never frees caller memory.
 *  - arrays are C-order [i][j][k] with k contiguous (fdtd_types.hpp:34-38).
 *    Device fields use a padded layout:  plane stride = ny*pitch floats,
 *    row stride = pitch floats (pitch % 4 == 0, pitch >= nz), and one ghost
 *    plane below (i = -1) and above (i = nx) the slab: (nx+2)*ny*pitch floats per
 *    buffer.  sb_field_elems() reports the element count to allocate.
 *  - all work is enqueued on the stream given to sb_create (a cudaStream_t cast
 *    to void*; NULL = legacy default stream).  Calls on one handle are not
 *    thread-safe; different handles are independent.
 *  - arithmetic: separately rounded IEEE fp32 operations in the reference's
 *    order (kernels are compiled --fmad=false), so results are bit-identical to
 *    the reference's C++ backend.
 */
#ifndef STRATA_B200_H
#define STRATA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_ABI_VERSION 1

typedef struct sb_solver sb_solver;

/* Grid slab description.  (GridShape, fdtd_types.hpp:26-32; slab fields are new.) */
typedef struct sb_grid_desc {
    int32_t nx, ny, nz;        /* local slab extents                                      */
    int32_t pitch;             /* device row pitch in floats; 0 = let the library choose  */
    int64_t global_nx;         /* extent of the whole grid along axis 0                   */
    int64_t i_offset;          /* global index of local plane 0                           */
    int32_t has_lower;         /* 1 if another slab owns planes below (ghost i=-1 is live)*/
    int32_t has_upper;         /* 1 if another slab owns planes above (ghost i=nx is live)*/
} sb_grid_desc;

/* One ADE pole (ADEMaterialData::DebyePole / LorentzPole, ade.hpp:44-66). */
typedef struct sb_pole {
    int32_t material_id;       /* 1..255                                                  */
    int32_t is_lorentz;        /* 0 = Debye (alpha,beta), 1 = Lorentz (a,b,d)             */
    int32_t target;            /* 0 = density (source p), 1 = modulus (source div v)      */
    int32_t reserved;
    float   c0, c1, c2;        /* Debye: alpha, beta, unused.  Lorentz: a, b, d           */
    float   reserved_f;
} sb_pole;

/* Throughput / traffic model for the last sb_step_n (sb_query). */
typedef struct sb_stats {
    int64_t cells;             /* nx*ny*nz of this slab                                   */
    int64_t steps_done;        /* total steps since create/reset                          */
    int64_t kernels_launched;  /* kernels launched by this handle since create            */
    double  algorithmic_bytes_per_cell;   /* 32 (+ADE model) */
    int32_t kernel_variant;    /* SB_KERNEL_* actually used by the last step              */
    int32_t pitch;
} sb_stats;

enum { SB_KERNEL_AUTO = 0, SB_KERNEL_NAIVE = 1, SB_KERNEL_MARCH = 2, /* 3 is retired: rejected by sb_set_option */
       SB_KERNEL_RESIDENT = 4, SB_KERNEL_PIPELINE = 5 };
/* SB_KERNEL_RESIDENT: for grids that fit in the GPU's aggregate shared memory (about 1.3 M cells on a B200) a whole
 * sb_step_n chunk runs as ONE cooperative launch that keeps the fields on chip between steps (sb_resident.cuh).
 * SB_KERNEL_AUTO picks it when the configuration allows (single slab, at most 32 point-source entries into p, probes
 * only, no ADE, at most 8 plane boundaries) and the chunk has at least SB_OPT_RESIDENT_MIN_STEPS steps; bit-identical. */
/* SB_KERNEL_PIPELINE: the marching kernel's tiles of ALL steps of a chunk run from one persistent cooperative launch;
 * a tile starts as soon as the three chunks of planes it touches have finished the previous step (sb_pipeline.cuh), so
 * the tail of one step overlaps the head of the next.  SB_KERNEL_AUTO uses it for grids of 6 M to 40 M cells, where it was
 * measured to pay (same applicability conditions as the resident kernel; with plane boundaries -- sb_add_plane_op, applied
 * inside its tiles -- it is used for every grid that does not fit the resident kernel); bit-identical results.            */
enum { SB_FIELD_P = 0, SB_FIELD_VX = 1, SB_FIELD_VY = 2, SB_FIELD_VZ = 3 };
enum { SB_OPT_KERNEL = 0, SB_OPT_ROWS_PER_THREAD = 1, SB_OPT_WARPS_J = 2, SB_OPT_WARPS_K = 3,
       SB_OPT_CHUNK_I = 4, SB_OPT_USE_GRAPH = 5, SB_OPT_PROFILE = 6, SB_OPT_FUSE_K3 = 7,
       SB_OPT_RESIDENT_SPLIT = 8, SB_OPT_RESIDENT_MIN_STEPS = 9, SB_OPT_PLANE_MAP = 10, SB_OPT_ADE_LAYOUT = 11,
       SB_OPT_ADE_CHUNK_I = 12, SB_OPT_ADE_WARPS = 13, SB_OPT_ADE_OCCUPANCY = 14 };

const char *sb_last_error(void);
int sb_abi_version(void);

/* ---- lifetime -------------------------------------------------------------------- */
/* Chooses pitch if desc->pitch == 0 and writes it back.                                 */
int sb_choose_pitch(int32_t nz, int32_t *pitch_out);
int64_t sb_field_elems(const sb_grid_desc *desc);
int sb_create(const sb_grid_desc *desc, int device, void *stream, sb_solver **out);
int sb_destroy(sb_solver *h);

/* Field storage: two sets of four buffers {p,vx,vy,vz} (ping-pong), each sb_field_elems()
 * floats, zero-initialised by the caller.  Replaces the four NumPy arrays FDTDSolver owns
 * (core/solver.py:1605-1608).  set 0 holds the current state after create/reset.         */
int sb_bind_fields(sb_solver *h, float *const set0[4], float *const set1[4]);
/* Which set holds the current state (flips every step). */
int sb_current_set(sb_solver *h, int *set_out);

/* Dense host [nx][ny][nz] <-> padded device layout of the current set. */
int sb_upload_field(sb_solver *h, int field, const float *host_dense);
int sb_download_field(sb_solver *h, int field, float *host_dense);

/* ---- physics set-up ---------------------------------------------------------------- */
/* Per-axis update tables, already in fp32:
 *   cv_face[a][m]  = velocity coefficient at face m of axis a  (n_a entries; the last is unused)
 *                    uniform: coeff_v (kernels.cpp:113);  nonuniform: coeff_v_base*inv_d_face[m]
 *                    (fdtd_step.cpp:263,281,302 -- one rounded fp32 multiply, done by the host)
 *   inv_cell[a][m] = 1/d_cell (fdtd_step.cpp:345-355) or NULL for a uniform grid
 *   cp             = coeff_p (uniform) or coeff_p_base (nonuniform)
 * Replaces update_velocity/update_pressure(+_nonuniform) arguments and
 * create_nonuniform_grid_data (kernels.cpp:108-176, 431-620).                            */
int sb_set_coefficients(sb_solver *h, const float *cv_x, const float *cv_y, const float *cv_z,
                        const float *inv_cell_x, const float *inv_cell_y, const float *inv_cell_z,
                        float cp);

/* Geometry: host bool/uint8 [nx][ny][nz] for this slab PLUS its ghost planes when they are
 * live ([nx + has_lower + has_upper] planes, lower ghost first); nonzero = air.
 * rigid != 0 enables face zeroing (= set_geometry was called, core/solver.py:1779-1780).
 * NULL geometry = all air.  Replaces precompute_boundary_cells + apply_rigid_boundaries
 * (boundaries.cpp:13-89) and update_pressure's geometry argument (fdtd_step.cpp:84-212).   */
int sb_set_geometry(sb_solver *h, const uint8_t *geom_host, int rigid);

/* Sponge ("PML") layers in application order.  decay tables are fp32, computed by the host
 * with libm expf(-sigma*dt) exactly as pml.cpp:13-45 does (sb_sponge_decay); NULL = axis absent.
 * The x table covers the local planes [-has_lower, nx): nx + has_lower entries, lower ghost first.
 * Replaces initialize_pml / apply_pml_velocity / apply_pml_pressure (pml.cpp:13-149).      */
/* decay[m] = expf(-sigma[m]*dt) with the host libm, exactly as initialize_pml does (pml.cpp:13-45). */
int sb_sponge_decay(const float *sigma, int n, float dt, float *decay_out);
int sb_clear_sponges(sb_solver *h);
int sb_add_sponge(sb_solver *h, const float *decay_x, const float *decay_y, const float *decay_z);

/* One-plane boundary updates applied to p after the sponges, in the order added (core/solver.py:2046-2047):
 * kind 0 = first-order Mur ABC (boundaries/_boundaries.py:476-513, mur = (c dt - dx)/(c dt + dx)),
 * kind 1 = RadiationImpedance (boundaries/_boundaries.py:700-760, R = reflection coefficient;
 * weak_r != 0 when the reference holds R as a Python float, which NumPy multiplies in fp32).
 * axis 0/1/2, side 0 = low face, 1 = high face.  The previous-plane state lives in the library.       */
int sb_clear_plane_ops(sb_solver *h);
int sb_add_plane_op(sb_solver *h, int axis, int side, int kind, double mur, double R, int weak_r);
/* Checkpoint / resume of that state (the reference's ``_p_prev_*`` arrays, boundaries/_boundaries.py:447-473): plane `op`
 * in the order added; host holds the face's two in-plane extents, row-major ([ny][nz], [nx][nz] or [nx][ny] floats);
 * *elems_out (optional) receives their product; upload != 0 writes the device state instead of reading it.            */
int sb_plane_op_state(sb_solver *h, int op, float *host, int64_t *elems_out, int upload);

/* ADE materials.  material_id_host: uint8 [nx][ny][nz], on a slab PLUS its live ghost planes as for
 * sb_set_geometry ([nx + has_lower + has_upper] planes, lower ghost first): the auxiliary density fields of the
 * ghost cells are advanced redundantly so that no J plane has to be exchanged.  rho_inf / K_inf are
 * indexed by material id (n_ids entries).  dt and inv_dx are the fp32 scalars the reference
 * passes (kernels.cpp:786-787, 834).  Replaces ADEMaterialData + update_ade_* + apply_ade_* +
 * compute_divergence* (ade.cpp:25-692) in the order of core/solver.py:2135-2193.           */
int sb_set_ade(sb_solver *h, const sb_pole *poles, int n_poles, const uint8_t *material_id_host,
               const float *rho_inf, const float *K_inf, int n_ids, float dt, float inv_dx);

/* Sources: CSR over cells.  Cell u (flat LOCAL dense index cell_idx[u] = (i*ny+j)*nz+k)
 * receives, in order e = start[u] .. start[u+1]-1,  f = f32( f64(f) + w[src_id[e]] * weight[e] )
 * (core/solver.py:2386-2433: float64 add, float32 store).  field[e] selects p/vx/vy/vz.
 * The host applies the reference's geometry test when it builds the list.
 * On a slab with a lower neighbour, cells of the ghost plane i = -1 (cell_idx in [-ny*nz, 0)) may be listed for
 * field vx only: the ghost face vx[-1] is kept redundantly, so an x-normal velocity source on the neighbour's last
 * plane has to be injected into it with the owner's operations (decomposition invariance).          */
int sb_set_sources(sb_solver *h, int n_sources, int n_cells, const int64_t *cell_idx,
                   const int32_t *start, const int32_t *src_id, const int32_t *field,
                   const double *weight);

/* Probes (core/solver.py:2435-2439): flat LOCAL dense indices.  Microphones
 * (microphones.cpp:82-116): 8 corner indices + 8 fp32 weights each, corner order of
 * microphones.hpp:30-32.  Record slots: probes first, then microphones.                  */
int sb_set_probes(sb_solver *h, int n_probes, const int64_t *flat_idx);
int sb_set_mics(sb_solver *h, int n_mics, const int64_t *idx8, const float *w8);
/* General form of sb_set_mics: every record slot gathers 8 weighted corners of field[t] (SB_FIELD_*).
 * Used for directional microphones, which sample p and the three velocity components with the
 * weights of the reference's Python path (core/solver.py:1004-1100); the host combines them.       */
int sb_set_gathers(sb_solver *h, int n, const int32_t *field, const int64_t *idx8, const float *w8);
/* Same tables the reference derives from grid positions (microphones.cpp:16-80). */
int sb_mic_tables(const float *grid_positions, int n_mics, int ny, int nz, int64_t *idx8, float *w8);

/* Checkpoint / resume of the auxiliary fields (the reference keeps them in private full-grid arrays, core/solver.py:
 * 3061-3083): pole index as passed to sb_set_ade, which = 0 for J, 1 for J_prev (Lorentz poles); host_dense is
 * [nx][ny][nz], zero outside the pole's material on download; upload != 0 writes the device state instead.  Together with
 * the four fields (sb_download_field / sb_upload_field), sb_plane_op_state and the host's time / step count this is the
 * whole solver state.                                                                                                   */
int sb_ade_state(sb_solver *h, int pole, int which, float *host_dense, int upload);

/* ---- stepping ------------------------------------------------------------------------ */
/* Advance n_steps.  src_values_host: [n_steps][n_sources] float64 waveform samples
 * (may be NULL if n_sources == 0).  record_out_host: [n_steps][n_probes+n_mics] fp32, written
 * when the call returns (may be NULL).  Synchronises the stream before returning.
 * Replaces n_steps iterations of FDTDSolver.step() (core/solver.py:2003-2077).           */
int sb_step_n(sb_solver *h, int n_steps, const double *src_values_host, float *record_out_host);

/* Asynchronous variant: src_values_dev / record_out_dev are device pointers sized as above;
 * nothing is copied and the stream is not synchronised.                                  */
int sb_step_n_async(sb_solver *h, int n_steps, const double *src_values_dev, float *record_out_dev);

/* sb_step_n without the wait, for a host loop that prepares chunk n+1 and files the records of chunk n-1 while the device
 * runs chunk n: the waveform table is copied up, the steps are enqueued and the records are copied down into
 * record_out_host, all on the handle's stream.  `slot` (0 or 1) names one of two sets of device staging buffers;
 * sb_step_n_wait(h, slot) returns when that chunk is complete and record_out_host is filled.  The host buffers must stay
 * valid until then and should be page-locked (otherwise the copies are staged synchronously by the driver).       */
int sb_step_n_submit(sb_solver *h, int slot, int n_steps, const double *src_values_host, float *record_out_host);
int sb_step_n_wait(sb_solver *h, int slot);

/* Host-driven halo exchange overlapped with the interior update: computes ONLY the planes next to this slab's cuts for
 * the step about to run (first) and remembers that the next sb_step_n_async(h, 1, ...) has to leave them out.  The
 * caller then sends those planes of the set being written (the one sb_current_set does NOT report) on another stream while
 * the rest of the step runs.  *applied = 0 (and nothing is launched) where K1 is not the last writer of the cut
 * planes -- ADE fix-ups, Mur / radiation planes, a slab thinner than 32 planes, the peer-to-peer halo -- the caller
 * must then exchange after the whole step; sources on a cut plane are the caller's to exclude.  (SURVEY.md 8e:
 * "compute the two face planes first, launch exchange, compute interior".)                                   */
int sb_step_cuts_async(sb_solver *h, int *applied);

/* Slab halo planes of the CURRENT set, for the exchange step of a multi-GPU run:
 * send_lo/send_hi = device addresses of owned planes 0 and nx-1 of p; recv_lo/recv_hi = the
 * ghost planes i=-1 and i=nx.  plane_elems = ny*pitch.                                   */
int sb_halo_planes(sb_solver *h, float **send_lo, float **send_hi, float **recv_lo, float **recv_hi,
                   int64_t *plane_elems);

/* Peer-to-peer halo over NVLink (optional; replaces the host-driven exchange of sb_halo_planes).
 * lo_p_sets / hi_p_sets: base addresses of the lower / upper neighbour's two p buffers (set 0, set 1), mapped
 * into this process (CUDA IPC / symmetric memory); lo_nx = the lower neighbour's slab extent.
 * my_flags: int[2] in peer-visible memory of THIS rank ([0] written by the lower, [1] by the upper neighbour);
 * lo_flag = address of the lower neighbour's my_flags[1], hi_flag = address of the upper neighbour's
 * my_flags[0].  From then on the fused step kernel stores its first / last p plane straight into the
 * neighbours' ghost planes, the blocks that touch a cut first wait (bounded) until the neighbour has
 * completed the previous step, and the last kernel of every step publishes this rank's step count.
 * All flags must be zero and all ranks at the same step count when this is called; my_flags == NULL disables. */
int sb_set_peers(sb_solver *h, float *const lo_p_sets[2], float *const hi_p_sets[2], int lo_nx,
                 int *my_flags, int *lo_flag, int *hi_flag);

/* 0.5*sum(p^2)/(rho c^2)*dV + 0.5*rho*sum(v^2)*dV over air cells (core/solver.py:2689-2706). */
int sb_energy(sb_solver *h, double rho, double c, double dV, double *out);

/* zero fields, J, counters (solver.py:2781-2800).  With peers set, also clears this rank's step flags: every rank
 * must be idle (synchronised + barrier) before, and barrier again after, the call.                               */
int sb_reset(sb_solver *h);
/* Tuning knobs; results never depend on them (every combination is bit-identical).
 *   SB_OPT_KERNEL            SB_KERNEL_*: which step kernel (default AUTO: resident K5 / pipelined K6 / streaming K1 by size)
 *   SB_OPT_ROWS_PER_THREAD   K1: rows per thread, 1 or 2 (0 = autotuned together with the next three)
 *   SB_OPT_WARPS_J / _K      K1: warps of a block along j / k;  SB_OPT_CHUNK_I: planes a tile marches over
 *   SB_OPT_PLANE_MAP         K1: 0 auto, 1 = 128-cell strips per warp, 2 = flat (the plane's float4 groups in memory order)
 *   SB_OPT_USE_GRAPH         step-by-step path: capture a chunk of steps into a CUDA graph (-1 auto, 0 off, 1 on)
 *   SB_OPT_FUSE_K3           inject point sources / record probes inside K1 instead of a separate kernel (0 off, 1 small grids, 2 always)
 *   SB_OPT_RESIDENT_SPLIT    K5: overlap the halo-free velocity updates with the face exchange (default off: measured slower)
 *   SB_OPT_RESIDENT_MIN_STEPS  shortest sb_step_n chunk for which AUTO uses K5 / K6 (default 4)
 *   SB_OPT_ADE_LAYOUT        material cells as 0 auto (fused on a single slab), 1 compact list, 2 dense bounding box,
 *                            3 fused into the step kernel (K1-ADE); set before sb_set_ade
 *   SB_OPT_ADE_CHUNK_I / SB_OPT_ADE_WARPS   K1-ADE: planes a tile marches over, warps per block (0 = default)
 *   SB_OPT_ADE_OCCUPANCY     K1-ADE: 256-thread blocks per SM the register allocation aims at (2 or 3)
 *   SB_OPT_PROFILE           bracket every step-kernel launch with CUDA events (sb_profile_read)                              */
int sb_set_option(sb_solver *h, int option, int value);
int sb_query(sb_solver *h, sb_stats *out);
/* With SB_OPT_PROFILE=1 every launch of the fused step kernel is bracketed by CUDA events on the
 * handle's stream; this returns (and clears) mean / min duration in ms and the launch count.   */
int sb_profile_read(sb_solver *h, double *mean_ms, double *min_ms, int *n_launches);
/* Launch shape chosen by the library's autotuner for the current configuration (rows per thread, warps along j,
 * warps along k, planes per chunk; 0 = heuristic) and the K1 time it measured; all shapes give identical results. */
int sb_tuned(sb_solver *h, int32_t shape_out[4], float *ms_out);
int sb_synchronize(sb_solver *h);

#ifdef __cplusplus
}
#endif
#endif /* STRATA_B200_H */
