// cbmd_internal.cuh — context, device layout and helpers shared by the kernels of
// libcbmd_cuda.so (sm_100a only).  See DESIGN.md for the layout rationale.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h> // header-only NVTX 3: ranges around the module entry points (option "nvtx")

#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cbmd_c_api.h"

// ---------------------------------------------------------------------------
// Position record: x, y, z and the atom type in one 32-byte sector so a neighbour
// gather is exactly one LDG.E.256 (sm_100a) and one DRAM/L2 sector.
// ---------------------------------------------------------------------------
struct alignas( 32 ) XT
{
    double x, y, z;
    long long t; // atom type (reference: separate int slice `type`)
};

#define CBMD_MAX_TYPES 8

struct LJTable
{
    int ntypes;
    double lj1[CBMD_MAX_TYPES * CBMD_MAX_TYPES];
    double lj2[CBMD_MAX_TYPES * CBMD_MAX_TYPES];
    double cutsq[CBMD_MAX_TYPES * CBMD_MAX_TYPES];
    // pair-energy constants: e1 = 0.5*lj1/6, e2 = lj2/6, eshift = energy at the cutoff
    double e1[CBMD_MAX_TYPES * CBMD_MAX_TYPES];
    double e2[CBMD_MAX_TYPES * CBMD_MAX_TYPES];
    double eshift[CBMD_MAX_TYPES * CBMD_MAX_TYPES];
};

// FP32 copy of the pair tables for the precision-32 force sweeps
struct LJTableF
{
    int ntypes;
    float lj1[CBMD_MAX_TYPES * CBMD_MAX_TYPES];
    float lj2[CBMD_MAX_TYPES * CBMD_MAX_TYPES];
    float cutsq[CBMD_MAX_TYPES * CBMD_MAX_TYPES];
    float e1[CBMD_MAX_TYPES * CBMD_MAX_TYPES];
    float e2[CBMD_MAX_TYPES * CBMD_MAX_TYPES];
    float eshift[CBMD_MAX_TYPES * CBMD_MAX_TYPES];
};

// Gather mirror of the positions for the full-list force sweeps (cbmd_force.cu).  Every kernel
// that writes positions (integrator, one-rank halo refresh, re-split) stores whichever parts
// are allocated:
//   kind 1 (FP64, one type):   xy = {x,y} packed (LDG.128)  +  zs = z (8-byte texels, TEX)
//   kind 2 (FP64, multi-type): xy                           +  zt = {z, type bits} (16-byte texels)
//   kind 3 (FP32, option "precision" 32): xf = {x,y,z,type bits} as floats (LDG.128)
struct MirrorPtrs
{
    double2 *xy;
    double *zs;
    double2 *zt;
    float4 *xf;
};

struct MassTable
{
    double dtfm[CBMD_MAX_TYPES]; // dtf / mass[type]
    double mass[CBMD_MAX_TYPES];
};

// one of the six halo phases (comm_mpi_impl.h:280-367)
struct HaloPhase
{
    int n_send = 0, n_recv = 0;
    int recv_first = 0;     // first ghost index of this phase's segment
    int *send_idx = nullptr; // device, capacity send_cap
    int send_cap = 0;
    double shift = 0.0;     // PBC shift the RECEIVER applies in dim phase/2
    int peer_send = -1, peer_recv = -1;
};

// ---------------------------------------------------------------------------
// In-process transport (cbmd_hub_create): one FIFO of posted messages per (source, destination)
// pair.  A message is {device pointer, bytes, event recorded by the sender}; the receiver
// orders its stream behind that event, copies device to device and hands back an event of
// its own that the sender's stream waits for before the send buffer is reused.
// ---------------------------------------------------------------------------
struct HubMsg
{
    const void *ptr = nullptr;
    size_t bytes = 0;
    cudaEvent_t ready = nullptr; // sender's, recorded when the payload is complete
    cudaEvent_t done = nullptr;  // receiver's, recorded behind its copy
    int state = 0;               // 0 posted, 1 copied (done valid), 2 sender has ordered itself behind done
};

struct cbmd_hub
{
    int nranks = 1;
    double timeout_s = 120.0;
    std::mutex m;
    std::condition_variable cv;
    std::vector<std::deque<std::shared_ptr<HubMsg>>> q; // [src * nranks + dst]
    // scalar collectives: one 1 KiB slot per rank + a generation barrier
    std::vector<std::vector<char>> slot;
    int arrived = 0;
    long generation = 0;
    int attached = 0;
    bool failed = false; // a rank timed out: every later wait fails at once
};

struct cbmd_ctx
{
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;

    // units / tables
    double boltz = 1.0, mvv2e = 1.0, dt = 0.005;
    int ntypes = 1;
    int max_type = 0; // largest atom type uploaded so far (0-based); checked against the tables
    MassTable mass;
    LJTable lj;

    // domain (system.h:88-113)
    double glo[3], ghi[3], gext[3], llo[3], lhi[3], ghost_lo[3], ghost_hi[3];
    int grid[3] = { 1, 1, 1 }, pos[3] = { 0, 0, 0 };
    bool have_domain = false;

    // atoms: [0,n_local) owned, [n_local,n_local+n_ghost) ghosts
    int n_local = 0, n_ghost = 0, cap = 0;
    XT *xt = nullptr, *xt_alt = nullptr;
    double *v = nullptr, *v_alt = nullptr; // SoA [3][cap]
    double *f = nullptr, *f_alt = nullptr; // SoA [3][cap]
    int *id = nullptr, *id_alt = nullptr;
    double *q = nullptr, *q_alt = nullptr;
    // gather mirror of the positions (see MirrorPtrs); one allocation, re-made when the
    // capacity or the wanted kind changes
    MirrorPtrs mir = { nullptr, nullptr, nullptr, nullptr };
    void *mirror_buf = nullptr;
    int mirror_kind = 0, mirror_cap = 0;
    int mirror_off = 0; // slide of the mirror, (-n_local) mod 16 entries (cbmd_force.cu ensure_mirror)
    cudaTextureObject_t tex_z = 0; // zs (8-byte texels) or zt (16-byte texels), by kind
    // value of `epoch` at which the owned / ghost part of the mirror was last consistent with
    // xt; the integrator and the one-rank halo refresh write the mirror themselves, anything
    // else that moves atoms leaves it stale and cbmd_force_lj re-splits that part
    uint64_t mirror_owned_epoch = 0, mirror_ghost_epoch = 0;
    int gather_mode = 1; // option "gather": 0 = 32-byte records by LDG.256, 1 = mirror (split / FP32) gathers
    int precision = 64;  // option "precision": 64, or 32 = FP32 pair arithmetic on float positions (full lists)
    LJTableF ljf;
    bool f_zero_pending = false; // deferred deep_copy(f,0): fused into the full-list force kernel
    // deferred Integrator::final_integrate: when the next call is initial_integrate the two
    // half kicks and the drift run as ONE streaming kernel (same roundings, 43 % less traffic);
    // any other entry point materialises it first
    bool final_pending = false;
    // positions/list epoch: bumped by every call that moves atoms or rebuilds the list;
    // validates the pair energy cached by a fused force+energy sweep
    uint64_t epoch = 1;
    bool energy_hint = false, pe_valid = false;
    // thermo output (three blocking reads per thermo step in the reference's order: T, PE, KE) costs one
    // host synchronisation instead of three: the fused sweep's energy is copied to pinned host memory
    // behind the sweep (ev_pe marks it; by the time PotE asks, the Temperature read has drained the
    // stream), and sum(m v^2) is cached until a call changes the velocities (v_epoch)
    cudaEvent_t ev_pe = nullptr;
    uint64_t pe_host_epoch = 0;
    uint64_t v_epoch = 1, mv2_epoch = 0;
    double mv2_cached = 0.0;
    uint64_t pe_epoch = 0;
    int pe_half = 0;
    double *pe_partial = nullptr;
    size_t pe_partial_cap = 0;

    // binning grid (binning_cabana.h:62-68) + cell lists over all atoms
    int nbin[3] = { 0, 0, 0 }, nhalo = 0;
    int ncell[3] = { 0, 0, 0 };
    double bmin[3], bmax[3], brdx[3];
    bool have_bins = false;
    int *cell_start = nullptr; // [ncells+1]
    int *cell_cursor = nullptr;
    int ncells_cap = 0;
    int *cell_atoms = nullptr; // [cap] atom indices sorted by (cell, index)
    int *atom_cell = nullptr;  // [cap]
    int *perm = nullptr;       // [cap] last sort permutation
    int perm_n = 0;

    // Verlet list, padded 2-D table; addressing by nb_entry() below
    int nb_half = 0, nb_layout = 0;
    int nb_rows = 0;   // row capacity: max_neigh_guess in effect rounded up to a multiple of 4
    int nb_stride = 0; // >= n_local, multiple of 32
    int nb_n = 0;      // n_local at build time
    int nb_ntot = 0;
    cudaTextureObject_t tex_nb = 0; // the table as 16-byte texels (index stream of the FP32 sweep)
    int *nb = nullptr;
    size_t nb_alloc = 0;
    int row_order = 0;         // option "row_order": 1 = bank-aware (Latin) order of full-list rows, 0 = index order
    int neigh_kernel = 2;      // option "neigh_kernel": 2 = per-thread walk, cells >= r (default); 1 = walk over half-size cells; 0 = staged stencil
    float4 *cpos = nullptr;    // candidates packed in cell order for the walk kernel
    int cpos_cap = 0;
    int *nb_count = nullptr; // [cap] entries per row
    // PULL table (half list for the atomics-free Newton-3 sweep, cbmd_neighbor.cu sweep_group MODE 2)
    bool nb_pull = false;
    int *nb_count_i = nullptr; // [cap] length of the reference half row (entries without NB_JSIDE)
    int pull_rows_hint = 0;
    int half_kernel = 1;       // option "half_kernel": 1 = atomics-free pull sweep (default), 0 = RED.ADD.F64 scatter
    int nb_max = 0;          // max row length of the last build
    double nb_rcut = 0.0;

    // comm
    int nranks = 1, rank = 0;
    ncclComm_t nccl = nullptr;
    cbmd_hub *hub = nullptr;                       // in-process transport instead of NCCL (cbmd_comm_init_hub)
    std::vector<std::shared_ptr<HubMsg>> hub_done; // my `done` events the senders may still be waiting on
    HaloPhase phase[6];
    double comm_depth = 0.0;
    bool have_halo = false;
    int *ghost_owner = nullptr;           // [cap] index of each ghost's root (owned) atom on its root rank
    int *ghost_rank = nullptr;            // [cap] root rank of each ghost
    unsigned char *ghost_image = nullptr; // [cap] packed image flags (accumulated PBC shifts)
    bool flat_halo_ok = false;            // one rank: every ghost is an image of an owned atom here
    // multi-rank one-stage refresh (cbmd_comm.cu): ghosts fetched straight from their ROOT rank.
    // Import side: ghosts rooted on rank p are slots [roff[p], roff[p]+rcnt[p]) of the receive
    // buffer, in ghost order (ghost_slot[g]); export side: export_idx[soff[p]..+scnt[p]) are the
    // owned atoms rank p wants every step, in the order rank p expects them.
    bool flat_mp_ok = false;
    // option "early_integrate" 1: cbmd_integrate_initial moves the BOUNDARY tiles first (the tile lists
    // of cbmd_neighbor.cu; every atom a ghost is an image of lies in one), so the one-stage refresh
    // starts on the comm stream while the interior tiles are still being integrated and the force sweep
    // is one launch.  Measured at N=2, 4 M atoms/GPU: force kernel 0.723 -> 0.698 ms, integrator 0.104 ->
    // 0.116 ms, 8 us of the refresh exposed: +0.6 % (noise level), and the window it hides behind
    // shrinks with the atoms per GPU — so the default stays 0: refresh beside the interior force tiles.
    int early_integrate = 0;
    bool early_posted = false; // integrate_initial recorded ev_x behind the early launch
    bool halo_early = false;   // the pending refresh started from that event
    int *ghost_slot = nullptr; // [cap]
    int *export_idx = nullptr;
    int export_cap = 0, n_export = 0, n_import = 0;
    std::vector<int> rcnt, roff, scnt, soff;
    double *sendbuf = nullptr, *recvbuf = nullptr;
    size_t sendbuf_bytes = 0, recvbuf_bytes = 0;
    // halo/compute overlap (multi-rank): update_halo runs on comm_stream while the force
    // kernel works on the tiles that have no ghost neighbour; the boundary tiles wait on
    // ev_halo.  tile_list = interior tiles (ascending) followed by boundary tiles.
    cudaStream_t comm_stream = nullptr;
    cudaStream_t aux_stream = nullptr; // boundary-tile force launch, concurrent with the interior tail
    cudaEvent_t ev_x = nullptr, ev_halo = nullptr, ev_boundary = nullptr, ev_fready = nullptr;
    bool halo_pending = false; // ghost positions are still in flight on comm_stream
    int overlap = 1;           // option "overlap": 0 keeps everything on one stream
    int halo_stages = 1;       // option "halo_stages": 1 = one-stage refresh from the root ranks, 3 = per dimension
    int *tile_list = nullptr, *tile_flag = nullptr;
    int tile_cap = 0, n_tiles_interior = 0, n_tiles_boundary = 0;
    bool tiles_valid = false;
    int tiles_n_local = -1;  // owned count the tile lists were made for
    double tiles_rcut = 0.0; // distance from the faces that makes a tile a boundary tile

    // scratch
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    void *scan_tmp = nullptr; // CUB temp storage of cbmd_exclusive_scan_int
    size_t scan_tmp_bytes = 0;
    double *d_red = nullptr; // small device reduction buffer
    int *d_flags = nullptr;  // small device int buffer (counters)
    double *h_pinned = nullptr;
    int *h_pinned_i = nullptr;

    // options

    bool nvtx = false; // option "nvtx" / CBMD_NVTX=1: NVTX range per timed region (Force, Neigh, Comm, ...)
    // CUDA-event timers (cbmd_timing_*)
    // cbmd_md_steps (cbmd_steps.cu): plain steps replayed from a CUDA graph
    int graph_steps = 1; // option "graph_steps"
    cudaGraphExec_t step_graph_exec = nullptr;
    int64_t graph_launches = 0;
    bool timing = false;
    int timing_stride = 7; // option "timing_stride": one plain step in this many is timed (1 = all)
    int timing_mode = 0;   // 0 = time the region; 1 / 2 = region of a sampled / unsampled plain step
    int64_t plain_seen = 0; // plain steps since timing was switched on or reset
    struct Bucket
    {
        // two classes of regions: [0] timed every time (everything outside cbmd_md_steps: rebuild and
        // thermo steps, set-up); [1] regions of the plain steps inside cbmd_md_steps, of which only a
        // sample is timed (timing_mode 1) and the rest only counted (timing_mode 2, graph replays)
        std::vector<cudaEvent_t> pending[2]; // start,end,start,end,...
        double ms[2] = { 0.0, 0.0 };         // device time of the timed regions
        int64_t count[2] = { 0, 0 };         // timed regions folded into ms
        int64_t calls[2] = { 0, 0 };         // regions entered while timing was on
    } bucket[CBMD_T_NBUCKETS];
    std::vector<cudaEvent_t> event_pool;
};

// RAII timed region on the context stream
struct TimedRegion
{
    cbmd_ctx *ctx;
    int b;
    int cls = 0;
    cudaEvent_t e1 = nullptr;
    static cudaEvent_t get( cbmd_ctx *c )
    {
        cudaEvent_t e;
        if ( !c->event_pool.empty() )
        {
            e = c->event_pool.back();
            c->event_pool.pop_back();
        }
        else
            cudaEventCreate( &e );
        return e;
    }
    bool ranged = false;
    TimedRegion( cbmd_ctx *c, int bucket )
        : ctx( c )
        , b( bucket )
    {
        if ( ctx->nvtx )
        {
            static const char *const names[] = { "cbmd:Force", "cbmd:Neigh", "cbmd:Comm", "cbmd:Integrate",
                                                 "cbmd:Other", "cbmd:ForceKernel" };
            nvtxRangePushA( names[bucket] );
            ranged = true;
        }
        if ( !ctx->timing )
            return;
        // A timing event costs the host a record and the device a timestamp between two kernels:
        // about 2.8 us each, ten per MD step — more than the kernels of a 32 000-atom step take
        // (profiles/r2_small_system_breakdown.txt).  Regions of the plain steps inside cbmd_md_steps
        // are therefore SAMPLED (the first 16 plain steps, then one in timing_stride) and
        // cbmd_timing_get scales their time to all plain steps; everything else is timed every time.
        cls = ctx->timing_mode == 0 ? 0 : 1;
        ctx->bucket[b].calls[cls]++;
        if ( ctx->timing_mode == 2 )
            return;
        cudaEvent_t e0 = get( ctx );
        e1 = get( ctx );
        cudaEventRecord( e0, ctx->stream );
        ctx->bucket[b].pending[cls].push_back( e0 );
    }
    ~TimedRegion()
    {
        if ( ranged )
            nvtxRangePop();
        if ( !e1 )
            return;
        cudaEventRecord( e1, ctx->stream );
        auto &B = ctx->bucket[b];
        auto &P = B.pending[cls];
        P.push_back( e1 );
        // long runs: fold the pairs that have already completed into the totals (no stall) so the
        // number of live events stays bounded
        if ( P.size() >= 1024 )
        {
            size_t k = 0;
            while ( k + 1 < P.size() && cudaEventQuery( P[k + 1] ) == cudaSuccess )
            {
                float ms = 0.f;
                if ( cudaEventElapsedTime( &ms, P[k], P[k + 1] ) == cudaSuccess )
                {
                    B.ms[cls] += ms;
                    B.count[cls]++;
                }
                ctx->event_pool.push_back( P[k] );
                ctx->event_pool.push_back( P[k + 1] );
                k += 2;
            }
            P.erase( P.begin(), P.begin() + k );
            (void)cudaGetLastError(); // a "not ready" from the last query is not an error
        }
    }
};

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
void cbmd_set_error( const std::string &msg );

struct CbmdError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

#define CBMD_CUDA( call )                                                                         \
    do                                                                                            \
    {                                                                                             \
        cudaError_t e__ = ( call );                                                               \
        if ( e__ != cudaSuccess )                                                                 \
            throw CbmdError( std::string( #call ) + " failed: " + cudaGetErrorString( e__ ) +     \
                             " (" + __FILE__ + ":" + std::to_string( __LINE__ ) + ")" );          \
    } while ( 0 )

#define CBMD_NCCL( call )                                                                         \
    do                                                                                            \
    {                                                                                             \
        ncclResult_t e__ = ( call );                                                              \
        if ( e__ != ncclSuccess )                                                                 \
            throw CbmdError( std::string( #call ) + " failed: " + ncclGetErrorString( e__ ) +     \
                             " (" + __FILE__ + ":" + std::to_string( __LINE__ ) + ")" );          \
    } while ( 0 )

#define CBMD_REQUIRE( cond, msg )                                                                 \
    do                                                                                            \
    {                                                                                             \
        if ( !( cond ) )                                                                          \
            throw CbmdError( std::string( msg ) + " (" + __FILE__ + ":" +                         \
                             std::to_string( __LINE__ ) + ")" );                                  \
    } while ( 0 )

// entry points that are aware of an in-flight halo update (force, update_halo)
#define CBMD_API_BEGIN_NOJOIN                                                                     \
    try                                                                                           \
    {                                                                                             \
        if ( !ctx )                                                                               \
            throw CbmdError( "null context" );                                                    \
        CBMD_CUDA( cudaSetDevice( ctx->device ) );

// every other entry point first orders the compute stream after a pending halo update
// and applies a deferred final_integrate
#define CBMD_API_BEGIN                                                                            \
    CBMD_API_BEGIN_NOJOIN                                                                         \
    cbmd_join_halo( ctx );                                                                        \
    cbmd_materialize_final( ctx );

#define CBMD_API_END                                                                              \
    return 0;                                                                                     \
    }                                                                                             \
    catch ( const std::exception &e )                                                             \
    {                                                                                             \
        cbmd_set_error( e.what() );                                                               \
        return 1;                                                                                 \
    }

#define CBMD_LAUNCH_CHECK( ctx )                                                                  \
    do                                                                                            \
    {                                                                                             \
        ( ctx )->launches++;                                                                      \
        CBMD_CUDA( cudaGetLastError() );                                                          \
    } while ( 0 )

inline int div_up( int a, int b ) { return ( a + b - 1 ) / b; }
inline int64_t div_up64( int64_t a, int64_t b ) { return ( a + b - 1 ) / b; }

// internal host helpers implemented across the .cu files
void cbmd_ensure_capacity( cbmd_ctx *ctx, int n );
void cbmd_hub_detach( cbmd_ctx *ctx ); // cbmd_comm.cu
void cbmd_graph_release( cbmd_ctx *ctx ); // cbmd_steps.cu
inline void cbmd_join_halo( cbmd_ctx *ctx )
{
    if ( ctx->halo_pending )
    {
        cudaStreamWaitEvent( ctx->stream, ctx->ev_halo, 0 );
        ctx->halo_pending = false;
    }
}
void *cbmd_scratch( cbmd_ctx *ctx, size_t bytes );
void cbmd_materialize_zero_force( cbmd_ctx *ctx );
void cbmd_materialize_final( cbmd_ctx *ctx );

// every call that moves atoms or rebuilds the list bumps the epoch; the mirror stays valid
// for the parts the call did not touch
inline void cbmd_bump_epoch( cbmd_ctx *ctx, bool owned_changed, bool ghosts_changed )
{
    const bool o = ctx->mirror_owned_epoch == ctx->epoch && !owned_changed;
    const bool g = ctx->mirror_ghost_epoch == ctx->epoch && !ghosts_changed;
    ctx->epoch++;
    if ( owned_changed )
        ctx->v_epoch++; // integrator, migration, sort, upload: the owned velocities (or their set) changed
    if ( o )
        ctx->mirror_owned_epoch = ctx->epoch;
    if ( g )
        ctx->mirror_ghost_epoch = ctx->epoch;
}
// mirror kind the next full-list force launch will want
inline int cbmd_mirror_wanted( const cbmd_ctx *ctx )
{
    if ( ctx->precision == 32 )
        return 3;
    if ( ctx->gather_mode != 1 )
        return 0;
    return ctx->lj.ntypes == 1 ? 1 : 2;
}
// the mirror is allocated for the current capacity and of the kind in use
inline bool cbmd_mirror_live( const cbmd_ctx *ctx )
{
    return ctx->mirror_kind != 0 && ctx->mirror_kind == cbmd_mirror_wanted( ctx ) && ctx->mirror_cap == ctx->cap;
}
// pointers for the kernels that write positions: all null unless the mirror is live
inline MirrorPtrs cbmd_mirror_ptrs( const cbmd_ctx *ctx )
{
    return cbmd_mirror_live( ctx ) ? ctx->mir : MirrorPtrs{ nullptr, nullptr, nullptr, nullptr };
}
void cbmd_exclusive_scan_int( cbmd_ctx *ctx, int *data, int n ); // in place, data[n] = total

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ XT ld_xt( const XT *p )
{
    XT r;
    double t;
    asm volatile( "ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                  : "=d"( r.x ), "=d"( r.y ), "=d"( r.z ), "=d"( t )
                  : "l"( p ) );
    r.t = __double_as_longlong( t );
    return r;
}

// exact (un-contracted) squared distance, summed left to right: the neighbour
// criterion must be bit-identical to the oracle's -ffp-contract=off evaluation.
__device__ __forceinline__ double dist2_exact( double dx, double dy, double dz )
{
    return __dadd_rn( __dadd_rn( __dmul_rn( dx, dx ), __dmul_rn( dy, dy ) ), __dmul_rn( dz, dz ) );
}

// cell coordinate of one position component: floor((x-min)*rdx) clamped, evaluated
// without contraction ([Cabana] LinkedCellList / oracle CellGrid::cell1).
__device__ __forceinline__ int cell_coord( double xv, double mn, double rdx, int n )
{
    int c = (int)floor( __dmul_rn( __dsub_rn( xv, mn ), rdx ) );
    c = c < 0 ? 0 : c;
    c = c > n - 1 ? n - 1 : c;
    return c;
}

// Verlet table layout: tiles of 32 consecutive atoms, rows packed four entries per 16 bytes.
// With rows4 = rows/4 (rows is a multiple of 4), entries 4k..4k+3 of atom i are the int4
//   nb4[((i >> 5) * rows4 + k) * 32 + (i & 31)]
// i.e. neighbour n of atom i is the int at
//   ((i >> 5) * rows4 + (n >> 2)) * 128 + (i & 31) * 4 + (n & 3).
// A warp of 32 consecutive atoms reads four entries per lane as ONE coalesced 512-byte request
// (a quarter of the index-load instructions of a one-entry-per-load table), and everything a
// warp touches is one contiguous rows*128-byte block (TLB- and DRAM-page-local).  Rows are
// padded to a multiple of four with the atom's own index (rejected by the sweeps: j != i).
__host__ __device__ __forceinline__ size_t nb_tile_base( int i, int rows )
{
    return ( (size_t)( i >> 5 ) * (size_t)( rows >> 2 ) ) * 128 + (size_t)( i & 31 ) * 4;
}
// element offset of neighbour n of atom i
__host__ __device__ __forceinline__ size_t nb_entry( int i, int n, int rows )
{
    return nb_tile_base( i, rows ) + (size_t)( n >> 2 ) * 128 + (size_t)( n & 3 );
}
// table elements needed for `stride` (multiple of 32) atoms
__host__ __device__ __forceinline__ size_t nb_table_size( int stride, int rows ) { return (size_t)stride * (size_t)rows; }

// PULL tables (half lists of the atomics-free Newton-3 sweep): an entry is the neighbour's index,
// flagged with NB_JSIDE when the pair is stored in the NEIGHBOUR's half row (it is listed here
// only so that this atom can pull its share of that pair)
#define NB_JSIDE ( 1 << 30 )
#define NB_INDEX_MASK 0x3fffffff

// store a position into whichever mirror parts exist (kernels that write xt call this)
__device__ __forceinline__ void mirror_store( const MirrorPtrs &m, int i, const XT &r )
{
    if ( m.xy )
    {
        m.xy[i] = make_double2( r.x, r.y );
        if ( m.zs )
            m.zs[i] = r.z;
        if ( m.zt )
            m.zt[i] = make_double2( r.z, __longlong_as_double( r.t ) );
    }
    if ( m.xf )
        m.xf[i] = make_float4( (float)r.x, (float)r.y, (float)r.z, __int_as_float( (int)r.t ) );
}

struct GridDesc
{
    double mn[3], rdx[3];
    int n[3];
};

__device__ __forceinline__ int cell_of( const GridDesc &g, double x, double y, double z )
{
    int a = cell_coord( x, g.mn[0], g.rdx[0], g.n[0] );
    int b = cell_coord( y, g.mn[1], g.rdx[1], g.n[1] );
    int c = cell_coord( z, g.mn[2], g.rdx[2], g.n[2] );
    return ( a * g.n[1] + b ) * g.n[2] + c;
}

inline GridDesc make_grid_desc( const cbmd_ctx *ctx )
{
    GridDesc g;
    for ( int d = 0; d < 3; d++ )
    {
        g.mn[d] = ctx->bmin[d];
        g.rdx[d] = ctx->brdx[d];
        g.n[d] = ctx->ncell[d];
    }
    return g;
}
