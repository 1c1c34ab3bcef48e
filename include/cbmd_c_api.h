/* include/cbmd_c_api.h — C ABI of libcbmd_cuda.so, the sm_100a implementation of
 * CabanaMD's short-range LJ MD step.
 *
 * This is the drop-in boundary: every entry point replaces one method of the
 * reference's module classes (paths relative to /root/reference/src).  Only plain
 * pointers and sizes cross it (HOST pointers unless stated otherwise); all device
 * memory, streams and NCCL communicators are owned by the context.  There is no
 * CPU fallback: every call fails (non-zero) when no CUDA device is usable.
 *
 * Conventions: every function returns 0 on success, non-zero on failure, with a
 * human readable message in cbmd_last_error() (thread local).  Per-atom host
 * arrays are row-major [n][3] doubles for x/v/f and int32 for type/id, i.e. the
 * layout of the reference's slices.  Atoms [0, n_local) are owned, then
 * [n_local, n_local + n_ghost) are ghosts, as in system.h:73-76.
 */
#ifndef CBMD_C_API_H
#define CBMD_C_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

    typedef struct cbmd_ctx cbmd_ctx;
    typedef struct cbmd_hub cbmd_hub; /* in-process transport between several contexts, see cbmd_hub_create */

    enum
    {
        CBMD_LAYOUT_2D = 0, /* Cabana::VerletLayout2D  (--neigh-type VERLET_2D)  */
        CBMD_LAYOUT_CSR = 1 /* Cabana::VerletLayoutCSR (--neigh-type VERLET_CSR) */
    };

    /* ---- life cycle / errors --------------------------------------------------
     * replaces: Kokkos::ScopeGuard + `new t_System; system->init()`
     * (bin/main.cpp:65, cabanamd_impl.h:74-75). */
    int cbmd_create( cbmd_ctx **out, int device );
    int cbmd_destroy( cbmd_ctx *ctx );
    const char *cbmd_last_error( void );
    const char *cbmd_version( void );
    /* blocks until all queued device work of this context is done (Kokkos::fence,
     * force_lj_cabana_neigh_impl.h:117) */
    int cbmd_sync( cbmd_ctx *ctx );

    /* ---- System (system.h:65-279, system_types/system_1aosoa.h) --------------- */
    /* units: system->boltz/mvv2e/dt (inputFile_impl.h:146-176) */
    int cbmd_set_units( cbmd_ctx *ctx, double boltz, double mvv2e, double dt );
    /* per-type masses: system->mass (system.h:83-86, inputFile_impl.h:304-314) */
    int cbmd_set_mass( cbmd_ctx *ctx, int ntypes, const double *mass );
    /* domain scalars produced by SystemCommon::create_domain (system.h:149-205,
     * 251-271): global box, this rank's own box, ghost-mesh box, rank grid/position */
    int cbmd_set_domain( cbmd_ctx *ctx, const double global_lo[3], const double global_hi[3],
                         const double local_lo[3], const double local_hi[3],
                         const double ghost_lo[3], const double ghost_hi[3], const int rank_grid[3],
                         const int rank_pos[3] );
    /* upload owned atoms and set N_local = n, N_ghost = 0 (System::resize +
     * deep_copy from the host system, inputFile_impl.h:786-791,851).  v, f, id, q
     * may be NULL (zero / 1..n / zero). */
    int cbmd_set_atoms( cbmd_ctx *ctx, int n_local, const double *x, const double *v,
                        const double *f, const int *type, const int *id, const double *q );
    /* append n ghosts (x, type[, id]) after the current atoms — test hook matching
     * the unit tests' "last num_ghost atoms are ghosts" set-up (tstNeighbor.hpp:262-264) */
    int cbmd_append_ghosts( cbmd_ctx *ctx, int n, const double *x, const int *type,
                            const int *id );
    /* download rows [first, first+count) (any pointer may be NULL) */
    int cbmd_get_atoms( cbmd_ctx *ctx, int first, int count, double *x, double *v, double *f,
                        int *type, int *id, double *q );
    /* overwrite v of owned atoms [0,n_local) (velocity rescale, inputFile_impl.h:858-865) */
    int cbmd_set_velocities( cbmd_ctx *ctx, int n_local, const double *v );
    int cbmd_get_counts( cbmd_ctx *ctx, int *n_local, int *n_ghost );

    /* ---- Integrator (integrator_nve.h:91-110, integrator_nve_impl.h:50-83) ---- */
    int cbmd_integrate_initial( cbmd_ctx *ctx );
    int cbmd_integrate_final( cbmd_ctx *ctx );

    /* ---- Binning (binning_cabana_impl.h:57-113) --------------------------------
     * Binning::create_binning(dx,dy,dz,halo_depth,do_local=true,do_ghost=false,
     * sort=true): cell-sorts the owned atoms and permutes all six fields.  Outputs
     * (may be NULL) are the public members nbinx/y/z, minx..maxz. */
    int cbmd_bin_sort( cbmd_ctx *ctx, double dx, double dy, double dz, int halo_depth,
                       int nbin_out[3], double min_out[3], double max_out[3] );
    /* permutation applied by the last cbmd_bin_sort: new[i] = old[perm[i]] */
    int cbmd_get_permutation( cbmd_ctx *ctx, int *perm );

    /* ---- Neighbor (neighbor_verlet.h:43-62, [Cabana] VerletList) ---------------
     * NeighborVerlet::create: rows [0,n_local) over all n_local+n_ghost atoms,
     * d^2 <= rcut^2 inclusive; half != 0 selects Cabana::HalfNeighborTag.
     * max_neigh_guess is the initial row capacity ("neigh_modify one"); the value
     * actually used/regrown (current_max * 1.1) is returned like
     * neighbor_verlet.h:58-61. */
    int cbmd_neigh_build( cbmd_ctx *ctx, double rcut, int half, int layout, int max_neigh_guess,
                          int *max_neigh_guess_out );
    /* NeighborList<>::maxNeighbor / total stored neighbours */
    int cbmd_neigh_sizes( cbmd_ctx *ctx, int64_t *total, int *max_neigh );
    /* host copy in CSR form: counts[n_local+n_ghost] (ghost rows 0),
     * offsets[n_local+1], neighbors[total]; equivalent of walking
     * numNeighbor/getNeighbor (tstNeighbor.hpp:55-73) */
    int cbmd_neigh_get( cbmd_ctx *ctx, int *counts, int64_t *offsets, int *neighbors );

    /* ---- Force (force.h:57-74, force_lj_cabana_neigh_impl.h) -------------------- */
    /* ForceLJ::init_coeff result tables, ntypes x ntypes row-major (:62-89) */
    int cbmd_set_lj( cbmd_ctx *ctx, int ntypes, const double *lj1, const double *lj2,
                     const double *cutsq );
    /* Cabana::deep_copy(f, 0.0) over owned+ghost atoms (cabanamd_impl.h:213-215,336-338) */
    int cbmd_zero_force( cbmd_ctx *ctx );
    /* ForceLJ::compute: accumulates into f; half != 0 => Newton-3 path, also
     * updating ghost f (:205-259) */
    int cbmd_force_lj( cbmd_ctx *ctx, int half );
    /* ForceLJ::compute_energy: shifted pair energy of this rank (:261-377).
     * pe_corrected (may be NULL): half-list energy with fac=1 for every stored pair
     * (SURVEY Appendix B.4); equals *pe for full lists. */
    int cbmd_energy_lj( cbmd_ctx *ctx, int half, double *pe, double *pe_corrected );
    /* Scalar pair virial W = sum over the pairs inside the force cutoff of r_ij . f_ij =
     * rsq * fpair, each pair once (this rank's share; sum over ranks for the global value; the
     * pressure is (N kB T + W/3) / V).  An extension: the reference has no virial (its pressure is
     * a TODO, src/cabanamd_impl.h:477-480).  Accumulated in the same sweep as the energy
     * (cbmd_request_energy before cbmd_force_lj makes it free), reduced in two deterministic
     * levels; otherwise one stand-alone sweep. */
    int cbmd_virial_lj( cbmd_ctx *ctx, int half, double *virial );
    /* One-shot hint from the step loop (cabanamd_impl.h:363-366 evaluates the energy on
     * thermo steps right after force->compute at unchanged positions): the NEXT
     * cbmd_force_lj also accumulates compute_energy in the same neighbour sweep, and
     * the following cbmd_energy_lj returns that value provided no call moved atoms or
     * rebuilt the list in between (otherwise it runs its own sweep, as without the hint). */
    int cbmd_request_energy( cbmd_ctx *ctx );

    /* ---- a run of plain steps (cabanamd_impl.h:285-399 between two rebuild / thermo steps) ---- */
    /* nsteps times: cbmd_integrate_initial, cbmd_update_halo, cbmd_zero_force, cbmd_force_lj( half ),
     * cbmd_update_force (half lists), cbmd_integrate_final — the same entry points in the same order,
     * so the result is bit-identical to calling them one by one.  On one rank the steps after the
     * first are replayed from a CUDA graph captured from those very calls (option "graph_steps"):
     * a step of input/in.lj (32 000 atoms) is launch-bound otherwise.  The caller keeps rebuild
     * steps (cbmd_exchange ... cbmd_neigh_build) and thermo steps (cbmd_request_energy) outside. */
    int cbmd_md_steps( cbmd_ctx *ctx, int nsteps, int half );

    /* ---- Comm (comm_mpi.h:125-138, comm_mpi_impl.h) ----------------------------- */
    /* 128-byte NCCL unique id created on rank 0 and handed to every rank out of band */
    int cbmd_comm_unique_id( void *id128 );
    /* Comm ctor + create_domain_decomposition (:52-119): NCCL communicator over
     * nranks (one process per GPU); nranks == 1 needs no id (may be NULL). */
    int cbmd_comm_init( cbmd_ctx *ctx, int nranks, int rank, const void *id128 );
    int cbmd_comm_rank( cbmd_ctx *ctx, int *rank, int *nranks );
    /* Several ranks inside ONE process (sub-domains sharing a GPU, or one host thread per GPU):
     * the hub carries what NCCL carries between processes — the same packed messages, as
     * stream-ordered device copies between the contexts — so the whole decomposed path
     * (migration, 6-phase ghost build, halo refresh, reverse force fold, scalar reductions)
     * runs with nranks > 1 on a single device.  NCCL refuses two ranks on one GPU; this is
     * how a 1-GPU box exercises Comm (comm_mpi_impl.h:191-441) at 2/4/8 ranks.  Every rank
     * drives its context from its own host thread: the calls that communicate block until
     * the peers reach the same call, like MPI.  A peer that does not arrive within
     * timeout_seconds (<= 0: 120 s) fails the call instead of hanging. */
    int cbmd_hub_create( cbmd_hub **out, int nranks, double timeout_seconds );
    int cbmd_hub_destroy( cbmd_hub *hub );
    int cbmd_comm_init_hub( cbmd_ctx *ctx, cbmd_hub *hub, int rank );
    /* Comm::exchange (:191-278): drop ghosts, PBC-wrap / migrate owned atoms;
     * returns the global number of migrated atoms */
    int cbmd_exchange( cbmd_ctx *ctx, int *n_sent_global );
    /* Comm::exchange_halo (:280-367): 6-phase ghost build with shell depth comm_depth */
    int cbmd_exchange_halo( cbmd_ctx *ctx, double comm_depth );
    /* Comm::update_halo (:369-408): refresh ghost positions */
    int cbmd_update_halo( cbmd_ctx *ctx );
    /* Comm::update_force (:410-441): add ghost forces back into their owners */
    int cbmd_update_force( cbmd_ctx *ctx );
    /* Comm::reduce_float / reduce_int / reduce_max_* / scan_int (:121-189), in place */
    int cbmd_reduce_sum_double( cbmd_ctx *ctx, double *vals, int count );
    int cbmd_reduce_sum_int( cbmd_ctx *ctx, int *vals, int count );
    int cbmd_reduce_max_double( cbmd_ctx *ctx, double *vals, int count );
    int cbmd_reduce_max_int( cbmd_ctx *ctx, int *vals, int count );
    int cbmd_scan_sum_int( cbmd_ctx *ctx, int *vals, int count );

    /* ---- thermo (property_temperature.h:73-79, property_kine.h:72-78) ---------- */
    /* sum over owned atoms of m v^2 (this rank only; callers reduce + scale) */
    int cbmd_sum_mv2( cbmd_ctx *ctx, double *sum );

    /* ---- measurement hooks (no reference equivalent) --------------------------- */
    /* the context's compute stream as a cudaStream_t, for CUDA-event timing */
    void *cbmd_stream( cbmd_ctx *ctx );
    /* number of kernels launched by this context since creation */
    int64_t cbmd_launch_count( cbmd_ctx *ctx );
    /* CUDA-event timers around the module entry points, mirroring the reference's
     * T_Force/T_Neigh/T_Comm/T_Int/T_Other buckets (cabanamd_impl.h:273-282,409-417),
     * plus the LJ force kernel alone.  Off by default. */
    enum
    {
        CBMD_T_FORCE = 0,        /* cbmd_zero_force + cbmd_force_lj            */
        CBMD_T_NEIGH = 1,        /* cbmd_neigh_build                           */
        CBMD_T_COMM = 2,         /* exchange / exchange_halo / update_*        */
        CBMD_T_INTEGRATE = 3,    /* cbmd_integrate_initial / _final            */
        CBMD_T_OTHER = 4,        /* cbmd_bin_sort, energy, sum_mv2             */
        CBMD_T_FORCE_KERNEL = 5, /* the k_force_* launch only                  */
        CBMD_T_NBUCKETS = 6
    };
    int cbmd_timing_enable( cbmd_ctx *ctx, int on );
    /* synchronises, then returns accumulated milliseconds and region count */
    int cbmd_timing_get( cbmd_ctx *ctx, int bucket, double *ms, int64_t *count );
    int cbmd_timing_reset( cbmd_ctx *ctx );
    /* Layout of the device Verlet table (host arithmetic only, no device needed): element
     * offset of neighbour n of atom i — tiles of 32 atoms, four entries of a row per 16 bytes:
     * ((i>>5)*rows/4 + (n>>2))*128 + (i&31)*4 + (n&3), rows = row_capacity rounded up to a
     * multiple of 4 — and the table size in elements for n_atoms rounded up to a multiple of
     * 32.  Lets tests pin the addressing the kernels share. */
    int64_t cbmd_table_offset( int atom, int n, int row_capacity );
    int64_t cbmd_table_size( int n_atoms, int row_capacity );
    /* kernel variant switches:
     *   "gather"    1 (default) the FP64 full-list force gathers x,y by LDG.128 from a packed
     *               mirror and z (multi-type: {z,type}) through the texture path; 0 = 32-byte
     *               records by LDG.256.  Same arithmetic, bit-identical forces.
     *   "precision" 64 (default), or 32: the full-list pair sweep runs in FP32 on float
     *               positions — the reference's T_X_FLOAT/T_F_FLOAT = float variant
     *               (src/types.h:133-148) for the force evaluation; integration state stays FP64.
     *               Forces agree with an FP32 evaluation of the same list to ~1e-6 relative.
     *   "half_kernel" 1 (default) Newton-3 sweep without atomics: the half list is stored as PULL
     *               rows (every atom's own half row + the rows that hold it), each force entry is
     *               written by one thread in a fixed order; 0 = scatter with RED.ADD.F64.  Takes
     *               effect at the next cbmd_neigh_build; cbmd_neigh_get returns the reference's
     *               half list either way.
     *   "row_order" 0 (default) rows in ascending (cell, index) order; 1 = full-list rows re-ordered
     *               after the build in a bank-aware (Latin) order: the eight lanes of an LDG.128
     *               group gather from eight different 16-byte positions (faster sweeps, but the
     *               re-ordering pass costs more than 20 sweeps gain).  Same sets.
     *   "neigh_kernel" 2 (default) Verlet build by one thread per atom walking its 3x3x3 cell stencil;
     *               1 = the same over a 5x5x5 stencil of half-size cells (fastest build, rows in an
     *               order the force sweep likes less); 0 = warp per cell over a staged stencil
     *               (round 1).  Same sets.
     *   "halo_stages" 1 (default) multi-rank ghost refresh straight from the root ranks in one NCCL
     *               group; 3 = the reference's forwarding scheme, one group per dimension
     *   "graph_steps" 1 (default) cbmd_md_steps replays its steps from a CUDA graph (one rank); 0 = launch
     *               by launch
     *   "timing_stride" 7 (default) the CUDA-event timers of cbmd_timing_* time every region of a bucket
     *               for its first 16 calls and one in 7 afterwards, and scale to all calls (an event pair
     *               per region costs more than the kernels of a small system take); 1 = time every region
     *   "nvtx"      1 = NVTX ranges (cbmd:Force, cbmd:Neigh, cbmd:Comm, ...) around the entry points
     *   "overlap"   1 (default) multi-rank halo refresh on a second stream beside other work, 0 = in line
     *   "early_integrate" 0 (default) the one-stage refresh runs beside the interior tiles of a split
     *               force sweep; 1 = cbmd_integrate_initial moves the boundary tiles first, the refresh
     *               runs beside the integration of the interior tiles and the force sweep is one launch
     *               (+0.6 % at N=2 with 4 M atoms per GPU; the window shrinks with the atoms per GPU)
     * Environment overrides read at cbmd_create: CBMD_GATHER, CBMD_PRECISION, CBMD_NEIGH_KERNEL,
     * CBMD_ROW_ORDER, CBMD_HALF_KERNEL, CBMD_HALO_STAGES, CBMD_OVERLAP, CBMD_EARLY, CBMD_GRAPH. */
    int cbmd_set_option( cbmd_ctx *ctx, const char *name, double value );

#ifdef __cplusplus
}
#endif
#endif /* CBMD_C_API_H */
