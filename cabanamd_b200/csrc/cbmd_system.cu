// cbmd_system.cu — context life cycle, particle store and host<->device transfers.
// Replaces System<Device,layout> (reference src/system.h:65-279,
// src/system_types/system_1aosoa.h:19-106): the three AoSoA layouts collapse to one
// device layout (32-byte position+type records, SoA velocities/forces).
#include "cbmd_internal.cuh"

static thread_local std::string g_last_error;

void cbmd_set_error( const std::string &msg ) { g_last_error = msg; }

extern "C" const char *cbmd_last_error( void ) { return g_last_error.c_str(); }
extern "C" const char *cbmd_version( void ) { return "cabanamd_b200 0.1 (sm_100a)"; }

// ---------------------------------------------------------------------------
// capacity management: System::resize only grows storage (system_1aosoa.h:59-67)
// ---------------------------------------------------------------------------
template <class T>
static void regrow( T *&p, size_t old_count, size_t new_count, cudaStream_t s )
{
    T *np = nullptr;
    CBMD_CUDA( cudaMalloc( &np, new_count * sizeof( T ) ) );
    if ( p && old_count )
        CBMD_CUDA( cudaMemcpyAsync( np, p, old_count * sizeof( T ), cudaMemcpyDeviceToDevice, s ) );
    if ( p )
    {
        CBMD_CUDA( cudaStreamSynchronize( s ) );
        CBMD_CUDA( cudaFree( p ) );
    }
    p = np;
}

static void regrow_soa3( double *&p, int old_cap, int new_cap, int n_used, cudaStream_t s )
{
    double *np = nullptr;
    CBMD_CUDA( cudaMalloc( &np, 3 * (size_t)new_cap * sizeof( double ) ) );
    CBMD_CUDA( cudaMemsetAsync( np, 0, 3 * (size_t)new_cap * sizeof( double ), s ) );
    if ( p && n_used )
        for ( int c = 0; c < 3; c++ )
            CBMD_CUDA( cudaMemcpyAsync( np + (size_t)c * new_cap, p + (size_t)c * old_cap,
                                        (size_t)n_used * sizeof( double ),
                                        cudaMemcpyDeviceToDevice, s ) );
    if ( p )
    {
        CBMD_CUDA( cudaStreamSynchronize( s ) );
        CBMD_CUDA( cudaFree( p ) );
    }
    p = np;
}

template <class T>
static void realloc_plain( T *&p, size_t count )
{
    if ( p )
        CBMD_CUDA( cudaFree( p ) );
    p = nullptr;
    CBMD_CUDA( cudaMalloc( &p, count * sizeof( T ) ) );
}

void cbmd_ensure_capacity( cbmd_ctx *ctx, int n )
{
    if ( n <= ctx->cap )
        return;
    int new_cap = n + n / 4 + 1024;
    new_cap = ( new_cap + 127 ) & ~127;
    const int used = ctx->n_local + ctx->n_ghost;
    cudaStream_t s = ctx->stream;
    regrow( ctx->xt, used, new_cap, s );
    regrow_soa3( ctx->v, ctx->cap, new_cap, used, s );
    regrow_soa3( ctx->f, ctx->cap, new_cap, used, s );
    regrow( ctx->id, used, new_cap, s );
    regrow( ctx->q, used, new_cap, s );
    regrow( ctx->nb_count, used, new_cap, s );
    regrow( ctx->nb_count_i, used, new_cap, s );
    regrow( ctx->ghost_owner, used, new_cap, s );
    regrow( ctx->ghost_image, used, new_cap, s );
    regrow( ctx->ghost_rank, used, new_cap, s );
    realloc_plain( ctx->ghost_slot, new_cap );
    // pure scratch / derived arrays: contents need not survive
    realloc_plain( ctx->xt_alt, new_cap );
    realloc_plain( ctx->v_alt, 3 * (size_t)new_cap );
    realloc_plain( ctx->f_alt, 3 * (size_t)new_cap );
    realloc_plain( ctx->id_alt, new_cap );
    realloc_plain( ctx->q_alt, new_cap );
    realloc_plain( ctx->cell_atoms, new_cap );
    realloc_plain( ctx->atom_cell, new_cap );
    realloc_plain( ctx->perm, new_cap );
    ctx->cap = new_cap;
}

void *cbmd_scratch( cbmd_ctx *ctx, size_t bytes )
{
    if ( bytes > ctx->scratch_bytes )
    {
        if ( ctx->scratch )
        {
            CBMD_CUDA( cudaStreamSynchronize( ctx->stream ) );
            CBMD_CUDA( cudaFree( ctx->scratch ) );
        }
        ctx->scratch = nullptr;
        size_t nb = bytes + bytes / 4 + ( 1 << 20 );
        CBMD_CUDA( cudaMalloc( &ctx->scratch, nb ) );
        ctx->scratch_bytes = nb;
    }
    return ctx->scratch;
}

// ---------------------------------------------------------------------------
// pack / unpack kernels between the reference's [n][3] slices and the device layout
// ---------------------------------------------------------------------------
// type_range[0] / [1] collect the smallest / largest type seen (validated by the host:
// the kernels index per-type tables of CBMD_MAX_TYPES entries with it)
__global__ void k_pack_xt( XT *__restrict__ xt, const double *__restrict__ x3,
                           const int *__restrict__ type, int first, int n,
                           int *__restrict__ type_range )
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    XT r;
    r.x = x3[3 * (size_t)i];
    r.y = x3[3 * (size_t)i + 1];
    r.z = x3[3 * (size_t)i + 2];
    const int t = type ? type[i] : 0;
    r.t = t;
    xt[first + i] = r;
    if ( t != 0 )
    {
        atomicMin( type_range, t );
        atomicMax( type_range + 1, t );
    }
}

__global__ void k_unpack_xt( const XT *__restrict__ xt, double *__restrict__ x3,
                             int *__restrict__ type, int first, int n )
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    XT r = xt[first + i];
    if ( x3 )
    {
        x3[3 * (size_t)i] = r.x;
        x3[3 * (size_t)i + 1] = r.y;
        x3[3 * (size_t)i + 2] = r.z;
    }
    if ( type )
        type[i] = (int)r.t;
}

// aos3 [n][3]  <->  soa [3][cap]
__global__ void k_aos_to_soa( double *__restrict__ soa, int cap, const double *__restrict__ aos,
                              int first, int n )
{
    int k = blockIdx.x * blockDim.x + threadIdx.x; // flat element of aos
    if ( k >= 3 * n )
        return;
    int i = k / 3, c = k - 3 * i;
    soa[(size_t)c * cap + first + i] = aos[k];
}

__global__ void k_soa_to_aos( const double *__restrict__ soa, int cap, double *__restrict__ aos,
                              int first, int n )
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if ( k >= 3 * n )
        return;
    int i = k / 3, c = k - 3 * i;
    aos[k] = soa[(size_t)c * cap + first + i];
}

__global__ void k_iota( int *__restrict__ a, int first, int n, int base )
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i < n )
        a[first + i] = base + i;
}

__global__ void k_fill3( double *__restrict__ soa, int cap, int first, int n, double val )
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    soa[first + i] = val;
    soa[(size_t)cap + first + i] = val;
    soa[2 * (size_t)cap + first + i] = val;
}

void cbmd_materialize_zero_force( cbmd_ctx *ctx )
{
    if ( !ctx->f_zero_pending )
        return;
    const int n = ctx->n_local + ctx->n_ghost;
    if ( n > 0 )
    {
        k_fill3<<<div_up( n, 256 ), 256, 0, ctx->stream>>>( ctx->f, ctx->cap, 0, n, 0.0 );
        CBMD_LAUNCH_CHECK( ctx );
    }
    ctx->f_zero_pending = false;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" int cbmd_create( cbmd_ctx **out, int device )
{
    try
    {
        if ( !out )
            throw CbmdError( "null out pointer" );
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount( &ndev );
        if ( e != cudaSuccess || ndev == 0 )
            throw CbmdError( std::string( "no CUDA device available (this library has no CPU "
                                          "fallback): " ) +
                             cudaGetErrorString( e ) );
        if ( device < 0 || device >= ndev )
            throw CbmdError( "device index out of range" );
        CBMD_CUDA( cudaSetDevice( device ) );
        cudaDeviceProp prop;
        CBMD_CUDA( cudaGetDeviceProperties( &prop, device ) );
        if ( prop.major != 10 )
            throw CbmdError( std::string( "libcbmd_cuda is built for sm_100a only; device is " ) +
                             prop.name + " (sm_" + std::to_string( prop.major ) +
                             std::to_string( prop.minor ) + ")" );
        cbmd_ctx *ctx = new cbmd_ctx;
        ctx->device = device;
        if ( const char *e = getenv( "CBMD_NVTX" ) ) // NVTX ranges around the module entry points
            ctx->nvtx = atoi( e ) != 0;
        if ( const char *e = getenv( "CBMD_OVERLAP" ) ) // A/B switch for measurements
            ctx->overlap = atoi( e );
        if ( const char *e = getenv( "CBMD_HALO_STAGES" ) ) // A/B switch: 3 = per-dimension forwarding
            ctx->halo_stages = atoi( e ) == 3 ? 3 : 1;
        if ( const char *e = getenv( "CBMD_GRAPH" ) ) // A/B switch: 0 = cbmd_md_steps issues every step launch by launch
            ctx->graph_steps = atoi( e ) != 0;
        if ( const char *e = getenv( "CBMD_EARLY" ) ) // A/B switch: 1 = refresh overlaps the integrator, one force launch
            ctx->early_integrate = atoi( e ) != 0;
        if ( const char *e = getenv( "CBMD_GATHER" ) ) // A/B switch: 0 = 32-byte records by LDG.256
            ctx->gather_mode = atoi( e ) == 0 ? 0 : 1;
        if ( const char *e = getenv( "CBMD_HALF_KERNEL" ) ) // A/B switch: 0 = RED.ADD.F64 scatter
            ctx->half_kernel = atoi( e ) == 0 ? 0 : 1;
        if ( const char *e = getenv( "CBMD_ROW_ORDER" ) ) // A/B switch: 1 = bank-aware row order
            ctx->row_order = atoi( e ) == 1 ? 1 : 0;
        if ( const char *e = getenv( "CBMD_NEIGH_KERNEL" ) ) // A/B switch: 1 = per-thread walk, half-size cells
            ctx->neigh_kernel = ( atoi( e ) >= 0 && atoi( e ) <= 2 ) ? atoi( e ) : 2;
        if ( const char *e = getenv( "CBMD_PRECISION" ) ) // 32 = FP32 pair arithmetic (full lists)
            ctx->precision = atoi( e ) == 32 ? 32 : 64;
        CBMD_CUDA( cudaStreamCreateWithFlags( &ctx->stream, cudaStreamNonBlocking ) );
        {
            int lo = 0, hi = 0; // comm stream gets the highest priority so its small kernels
            CBMD_CUDA( cudaDeviceGetStreamPriorityRange( &lo, &hi ) ); // slip in between force CTAs
            CBMD_CUDA( cudaStreamCreateWithPriority( &ctx->comm_stream, cudaStreamNonBlocking, hi ) );
            CBMD_CUDA( cudaEventCreateWithFlags( &ctx->ev_pe, cudaEventDisableTiming ) );
            CBMD_CUDA( cudaEventCreateWithFlags( &ctx->ev_x, cudaEventDisableTiming ) );
            CBMD_CUDA( cudaEventCreateWithFlags( &ctx->ev_halo, cudaEventDisableTiming ) );
            CBMD_CUDA( cudaStreamCreateWithFlags( &ctx->aux_stream, cudaStreamNonBlocking ) );
            CBMD_CUDA( cudaEventCreateWithFlags( &ctx->ev_boundary, cudaEventDisableTiming ) );
            CBMD_CUDA( cudaEventCreateWithFlags( &ctx->ev_fready, cudaEventDisableTiming ) );
        }
        memset( &ctx->mass, 0, sizeof( ctx->mass ) );
        memset( &ctx->lj, 0, sizeof( ctx->lj ) );
        ctx->lj.ntypes = 1;
        for ( int t = 0; t < CBMD_MAX_TYPES; t++ )
        {
            ctx->mass.mass[t] = 1.0;
            ctx->mass.dtfm[t] = 0.5 * ctx->dt / ctx->mvv2e;
        }
        CBMD_CUDA( cudaMalloc( &ctx->d_red, 65536 * sizeof( double ) ) );
        CBMD_CUDA( cudaMalloc( &ctx->d_flags, 65536 * sizeof( int ) ) );
        CBMD_CUDA( cudaMallocHost( &ctx->h_pinned, 64 * sizeof( double ) ) );
        CBMD_CUDA( cudaMallocHost( &ctx->h_pinned_i, 64 * sizeof( int ) ) );
        *out = ctx;
        return 0;
    }
    catch ( const std::exception &e )
    {
        cbmd_set_error( e.what() );
        return 1;
    }
}

extern "C" int cbmd_destroy( cbmd_ctx *ctx )
{
    if ( !ctx )
        return 0;
    cudaSetDevice( ctx->device );
    cudaStreamSynchronize( ctx->stream );
    cbmd_hub_detach( ctx );
    cbmd_graph_release( ctx );
    if ( ctx->comm_stream )
    {
        cudaStreamSynchronize( ctx->comm_stream );
        cudaStreamDestroy( ctx->comm_stream );
        cudaEventDestroy( ctx->ev_x );
        cudaEventDestroy( ctx->ev_pe );
        cudaEventDestroy( ctx->ev_halo );
        cudaStreamSynchronize( ctx->aux_stream );
        cudaStreamDestroy( ctx->aux_stream );
        cudaEventDestroy( ctx->ev_boundary );
        cudaEventDestroy( ctx->ev_fready );
    }
    if ( ctx->tex_z )
        cudaDestroyTextureObject( ctx->tex_z );
    if ( ctx->tex_nb )
        cudaDestroyTextureObject( ctx->tex_nb );
    if ( ctx->mirror_buf )
        cudaFree( ctx->mirror_buf );
    void *ptrs[] = { ctx->xt,         ctx->xt_alt,      ctx->v,          ctx->v_alt,
                     ctx->f,          ctx->f_alt,       ctx->id,         ctx->id_alt,
                     ctx->q,          ctx->q_alt,       ctx->cell_start, ctx->cell_cursor,
                     ctx->cell_atoms, ctx->atom_cell,   ctx->perm,       ctx->nb,
                     ctx->nb_count,   ctx->ghost_owner, ctx->ghost_image, ctx->sendbuf,
                     ctx->recvbuf,    ctx->scratch,     ctx->d_red,      ctx->d_flags,
                     ctx->pe_partial, ctx->tile_list, ctx->tile_flag, ctx->scan_tmp, ctx->cpos,
                     ctx->ghost_rank, ctx->ghost_slot, ctx->export_idx, ctx->nb_count_i };
    for ( void *p : ptrs )
        if ( p )
            cudaFree( p );
    for ( int ph = 0; ph < 6; ph++ )
        if ( ctx->phase[ph].send_idx )
            cudaFree( ctx->phase[ph].send_idx );
    if ( ctx->h_pinned )
        cudaFreeHost( ctx->h_pinned );
    if ( ctx->h_pinned_i )
        cudaFreeHost( ctx->h_pinned_i );
    if ( ctx->nccl )
        ncclCommDestroy( ctx->nccl );
    for ( int b = 0; b < CBMD_T_NBUCKETS; b++ )
        for ( int c = 0; c < 2; c++ )
            for ( cudaEvent_t e : ctx->bucket[b].pending[c] )
                cudaEventDestroy( e );
    for ( cudaEvent_t e : ctx->event_pool )
        cudaEventDestroy( e );
    cudaStreamDestroy( ctx->stream );
    delete ctx;
    return 0;
}

extern "C" int cbmd_sync( cbmd_ctx *ctx )
{
    CBMD_API_BEGIN
    CBMD_CUDA( cudaStreamSynchronize( ctx->stream ) );
    CBMD_API_END
}

extern "C" void *cbmd_stream( cbmd_ctx *ctx ) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int64_t cbmd_launch_count( cbmd_ctx *ctx ) { return ctx ? ctx->launches : 0; }

extern "C" int cbmd_set_option( cbmd_ctx *ctx, const char *name, double value )
{
    CBMD_API_BEGIN
    std::string n( name ? name : "" );
    if ( n == "overlap" )
        ctx->overlap = (int)value;
    else if ( n == "nvtx" )
        ctx->nvtx = value != 0.0;
    else if ( n == "graph_steps" )
        ctx->graph_steps = value != 0.0;
    else if ( n == "timing_stride" )
    {
        if ( value < 1.0 )
            throw CbmdError( "timing_stride must be >= 1" );
        ctx->timing_stride = (int)value;
    }
    else if ( n == "early_integrate" )
    {
        // multi-rank step with the one-stage refresh: 0 (default) = the exchange runs beside the interior
        // tiles of a split force sweep; 1 = the boundary tiles are integrated first, the exchange runs
        // beside the integration of the interior tiles and the force sweep is one launch
        ctx->early_integrate = value != 0.0;
    }
    else if ( n == "halo_stages" )
    {
        // multi-rank ghost refresh: 1 = every ghost straight from its root rank in one NCCL
        // group (default); 3 = the reference's forwarding scheme, one group per dimension
        if ( (int)value != 1 && (int)value != 3 )
            throw CbmdError( "halo_stages must be 1 or 3" );
        ctx->halo_stages = (int)value;
    }
    else if ( n == "gather" )
    {
        if ( (int)value != 0 && (int)value != 1 )
            throw CbmdError( "gather must be 0 (32-byte records by LDG.256) or 1 (mirror: xy LDG.128 + z TEX, FP32 float4)" );
        ctx->gather_mode = (int)value;
    }
    else if ( n == "half_kernel" )
    {
        // Newton-3 sweep (takes effect at the next cbmd_neigh_build): 1 = pull rows, no atomics,
        // deterministic (default); 0 = scatter f_j with RED.ADD.F64 (round 1)
        if ( (int)value != 0 && (int)value != 1 )
            throw CbmdError( "half_kernel must be 0 or 1" );
        ctx->half_kernel = (int)value;
    }
    else if ( n == "row_order" )
    {
        // order of the entries inside a full-list row (next cbmd_neigh_build): 0 = ascending
        // (cell, index) (default), 1 = bank-aware Latin order for the LDG.128 gathers
        if ( (int)value != 0 && (int)value != 1 )
            throw CbmdError( "row_order must be 0 or 1" );
        ctx->row_order = (int)value;
    }
    else if ( n == "neigh_kernel" )
    {
        // Verlet build: 2 = one thread per atom walks its 3x3x3 stencil of cells >= r (default);
        // 1 = the same over a 5x5x5 stencil of half-size cells; 0 = warp per cell over a staged
        // 27-cell stencil.  Same sets, bit for bit.
        if ( (int)value < 0 || (int)value > 2 )
            throw CbmdError( "neigh_kernel must be 0, 1 or 2" );
        ctx->neigh_kernel = (int)value;
    }
    else if ( n == "precision" )
    {
        // arithmetic of the full-list pair sweeps: 64 (default), or 32 = the reference's
        // T_F_FLOAT/T_X_FLOAT = float variant (types.h:133-148) for the force evaluation:
        // positions rounded to float, FP32 pair terms and per-atom sums; the integration
        // state (x, v, f arrays) stays FP64
        if ( (int)value != 32 && (int)value != 64 )
            throw CbmdError( "precision must be 32 or 64" );
        ctx->precision = (int)value;
        ctx->pe_valid = false;
    }
    else
        throw CbmdError( "unknown option: " + n );
    CBMD_API_END
}

static void timing_drain( cbmd_ctx *ctx )
{
    CBMD_CUDA( cudaStreamSynchronize( ctx->stream ) );
    for ( int b = 0; b < CBMD_T_NBUCKETS; b++ )
        for ( int c = 0; c < 2; c++ )
        {
            auto &B = ctx->bucket[b];
            auto &P = B.pending[c];
            for ( size_t k = 0; k + 1 < P.size(); k += 2 )
            {
                float ms = 0.f;
                CBMD_CUDA( cudaEventElapsedTime( &ms, P[k], P[k + 1] ) );
                B.ms[c] += ms;
                B.count[c]++;
                ctx->event_pool.push_back( P[k] );
                ctx->event_pool.push_back( P[k + 1] );
            }
            P.clear();
        }
}

extern "C" int cbmd_timing_enable( cbmd_ctx *ctx, int on )
{
    CBMD_API_BEGIN
    if ( !on )
        timing_drain( ctx );
    else if ( !ctx->timing )
        ctx->plain_seen = 0;
    ctx->timing = on != 0;
    CBMD_API_END
}

extern "C" int cbmd_timing_get( cbmd_ctx *ctx, int bucket, double *ms, int64_t *count )
{
    CBMD_API_BEGIN
    CBMD_REQUIRE( bucket >= 0 && bucket < CBMD_T_NBUCKETS, "bad timing bucket" );
    timing_drain( ctx );
    const auto &B = ctx->bucket[bucket];
    if ( ms ) // regions timed every time + the sampled plain-step regions scaled to all plain steps
        *ms = B.ms[0] + ( B.count[1] > 0 ? B.ms[1] * ( (double)B.calls[1] / (double)B.count[1] ) : 0.0 );
    if ( count )
        *count = B.calls[0] + B.calls[1];
    CBMD_API_END
}

extern "C" int cbmd_timing_reset( cbmd_ctx *ctx )
{
    CBMD_API_BEGIN
    timing_drain( ctx );
    for ( int b = 0; b < CBMD_T_NBUCKETS; b++ )
        for ( int c = 0; c < 2; c++ )
        {
            ctx->bucket[b].ms[c] = 0.0;
            ctx->bucket[b].count[c] = 0;
            ctx->bucket[b].calls[c] = 0;
        }
    ctx->plain_seen = 0;
    CBMD_API_END
}

static void refresh_dtfm( cbmd_ctx *ctx )
{
    // integrator_nve_impl.h:50-54: dtf = 0.5*dt/mvv2e; dtfm = dtf/mass[type]
    const double dtf = 0.5 * ctx->dt / ctx->mvv2e;
    for ( int t = 0; t < CBMD_MAX_TYPES; t++ )
        ctx->mass.dtfm[t] = dtf / ctx->mass.mass[t];
}

extern "C" int cbmd_set_units( cbmd_ctx *ctx, double boltz, double mvv2e, double dt )
{
    CBMD_API_BEGIN
    ctx->boltz = boltz;
    ctx->mvv2e = mvv2e;
    ctx->dt = dt;
    refresh_dtfm( ctx );
    CBMD_API_END
}

extern "C" int cbmd_set_mass( cbmd_ctx *ctx, int ntypes, const double *mass )
{
    CBMD_API_BEGIN
    CBMD_REQUIRE( ntypes >= 1 && ntypes <= CBMD_MAX_TYPES, "ntypes out of range (1..8)" );
    CBMD_REQUIRE( mass != nullptr, "null mass" );
    ctx->ntypes = ntypes;
    for ( int t = 0; t < ntypes; t++ )
        ctx->mass.mass[t] = mass[t];
    refresh_dtfm( ctx );
    ctx->v_epoch++; // sum(m v^2) changes with the masses
    CBMD_API_END
}

extern "C" int cbmd_set_domain( cbmd_ctx *ctx, const double global_lo[3],
                                const double global_hi[3], const double local_lo[3],
                                const double local_hi[3], const double ghost_lo[3],
                                const double ghost_hi[3], const int rank_grid[3],
                                const int rank_pos[3] )
{
    CBMD_API_BEGIN
    for ( int d = 0; d < 3; d++ )
    {
        ctx->glo[d] = global_lo[d];
        ctx->ghi[d] = global_hi[d];
        ctx->gext[d] = global_hi[d] - global_lo[d];
        ctx->llo[d] = local_lo[d];
        ctx->lhi[d] = local_hi[d];
        ctx->ghost_lo[d] = ghost_lo ? ghost_lo[d] : local_lo[d];
        ctx->ghost_hi[d] = ghost_hi ? ghost_hi[d] : local_hi[d];
        ctx->grid[d] = rank_grid ? rank_grid[d] : 1;
        ctx->pos[d] = rank_pos ? rank_pos[d] : 0;
        CBMD_REQUIRE( ctx->grid[d] >= 1 && ctx->pos[d] >= 0 && ctx->pos[d] < ctx->grid[d],
                      "bad rank grid/position" );
    }
    ctx->have_domain = true;
    ctx->have_bins = false;
    ctx->have_halo = false;
    CBMD_API_END
}

static void upload_rows( cbmd_ctx *ctx, int first, int n, const double *x, const double *v,
                         const double *f, const int *type, const int *id, const double *q,
                         int id_base )
{
    if ( n == 0 )
        return;
    cudaStream_t s = ctx->stream;
    const size_t b3 = 3 * (size_t)n * sizeof( double );
    char *st = (char *)cbmd_scratch( ctx, b3 + (size_t)n * sizeof( int ) + 256 );
    double *d3 = (double *)st;
    int *dt = (int *)( st + ( ( b3 + 255 ) & ~(size_t)255 ) );
    const int tb = 256;
    CBMD_REQUIRE( x != nullptr, "null x" );
    CBMD_CUDA( cudaMemcpyAsync( d3, x, b3, cudaMemcpyHostToDevice, s ) );
    if ( type )
        CBMD_CUDA( cudaMemcpyAsync( dt, type, (size_t)n * sizeof( int ), cudaMemcpyHostToDevice, s ) );
    int *type_range = ctx->d_flags + 20; // {min, max} over the non-zero types of this upload
    CBMD_CUDA( cudaMemsetAsync( type_range, 0, 2 * sizeof( int ), s ) );
    k_pack_xt<<<div_up( n, tb ), tb, 0, s>>>( ctx->xt, d3, type ? dt : nullptr, first, n, type_range );
    CBMD_LAUNCH_CHECK( ctx );
    CBMD_CUDA( cudaMemcpyAsync( ctx->h_pinned_i + 4, type_range, 2 * sizeof( int ), cudaMemcpyDeviceToHost, s ) );
    if ( v )
    {
        CBMD_CUDA( cudaMemcpyAsync( d3, v, b3, cudaMemcpyHostToDevice, s ) );
        k_aos_to_soa<<<div_up( 3 * n, tb ), tb, 0, s>>>( ctx->v, ctx->cap, d3, first, n );
    }
    else
        k_fill3<<<div_up( n, tb ), tb, 0, s>>>( ctx->v, ctx->cap, first, n, 0.0 );
    CBMD_LAUNCH_CHECK( ctx );
    if ( f )
    {
        CBMD_CUDA( cudaMemcpyAsync( d3, f, b3, cudaMemcpyHostToDevice, s ) );
        k_aos_to_soa<<<div_up( 3 * n, tb ), tb, 0, s>>>( ctx->f, ctx->cap, d3, first, n );
    }
    else
        k_fill3<<<div_up( n, tb ), tb, 0, s>>>( ctx->f, ctx->cap, first, n, 0.0 );
    CBMD_LAUNCH_CHECK( ctx );
    if ( id )
        CBMD_CUDA( cudaMemcpyAsync( ctx->id + first, id, (size_t)n * sizeof( int ),
                                    cudaMemcpyHostToDevice, s ) );
    else
    {
        k_iota<<<div_up( n, tb ), tb, 0, s>>>( ctx->id, first, n, id_base );
        CBMD_LAUNCH_CHECK( ctx );
    }
    if ( q )
        CBMD_CUDA( cudaMemcpyAsync( ctx->q + first, q, (size_t)n * sizeof( double ),
                                    cudaMemcpyHostToDevice, s ) );
    else
        CBMD_CUDA( cudaMemsetAsync( ctx->q + first, 0, (size_t)n * sizeof( double ), s ) );
    // host buffers may be pageable: make the call synchronous w.r.t. them
    CBMD_CUDA( cudaStreamSynchronize( s ) );
    // atom types index per-type tables in the kernels (masses, pair coefficients)
    const int tmin = ctx->h_pinned_i[4], tmax = ctx->h_pinned_i[5];
    CBMD_REQUIRE( tmin >= 0 && tmax < CBMD_MAX_TYPES,
                  "atom type outside [0, " + std::to_string( CBMD_MAX_TYPES ) +
                      ") (0-based; this build supports at most " + std::to_string( CBMD_MAX_TYPES ) +
                      " atom types): " + std::to_string( tmin < 0 ? tmin : tmax ) );
    if ( tmax > ctx->max_type )
        ctx->max_type = tmax;
}

extern "C" int cbmd_set_atoms( cbmd_ctx *ctx, int n_local, const double *x, const double *v,
                               const double *f, const int *type, const int *id, const double *q )
{
    CBMD_API_BEGIN
    CBMD_REQUIRE( n_local >= 0, "negative atom count" );
    ctx->n_local = 0;
    ctx->n_ghost = 0;
    ctx->max_type = 0;
    cbmd_bump_epoch( ctx, true, true );
    cbmd_ensure_capacity( ctx, n_local );
    ctx->f_zero_pending = false;
    upload_rows( ctx, 0, n_local, x, v, f, type, id, q, 1 );
    ctx->n_local = n_local;
    ctx->have_halo = false;
    ctx->nb_n = 0;
    ctx->nb_ntot = 0;
    CBMD_API_END
}

extern "C" int cbmd_append_ghosts( cbmd_ctx *ctx, int n, const double *x, const int *type,
                                   const int *id )
{
    CBMD_API_BEGIN
    CBMD_REQUIRE( n >= 0, "negative ghost count" );
    cbmd_materialize_zero_force( ctx );
    cbmd_bump_epoch( ctx, true, true );
    const int first = ctx->n_local + ctx->n_ghost;
    cbmd_ensure_capacity( ctx, first + n );
    upload_rows( ctx, first, n, x, nullptr, nullptr, type, id, nullptr, first + 1 );
    ctx->n_ghost += n;
    ctx->have_halo = false; // no owner map for hand-made ghosts
    CBMD_API_END
}

extern "C" int cbmd_set_velocities( cbmd_ctx *ctx, int n_local, const double *v )
{
    CBMD_API_BEGIN
    CBMD_REQUIRE( n_local == ctx->n_local, "velocity count != n_local" );
    if ( n_local > 0 )
    {
        const size_t b3 = 3 * (size_t)n_local * sizeof( double );
        double *d3 = (double *)cbmd_scratch( ctx, b3 );
        CBMD_CUDA( cudaMemcpyAsync( d3, v, b3, cudaMemcpyHostToDevice, ctx->stream ) );
        k_aos_to_soa<<<div_up( 3 * n_local, 256 ), 256, 0, ctx->stream>>>( ctx->v, ctx->cap, d3, 0,
                                                                         n_local );
        CBMD_LAUNCH_CHECK( ctx );
        CBMD_CUDA( cudaStreamSynchronize( ctx->stream ) );
    }
    ctx->v_epoch++;
    CBMD_API_END
}

extern "C" int cbmd_get_atoms( cbmd_ctx *ctx, int first, int count, double *x, double *v,
                               double *f, int *type, int *id, double *q )
{
    CBMD_API_BEGIN
    CBMD_REQUIRE( first >= 0 && count >= 0 && first + count <= ctx->n_local + ctx->n_ghost,
                  "row range outside [0, n_local+n_ghost)" );
    if ( count == 0 )
        return 0;
    cbmd_materialize_zero_force( ctx );
    cudaStream_t s = ctx->stream;
    const int tb = 256;
    const size_t b3 = 3 * (size_t)count * sizeof( double );
    char *st = (char *)cbmd_scratch( ctx, b3 + (size_t)count * sizeof( int ) + 256 );
    double *d3 = (double *)st;
    int *dt = (int *)( st + ( ( b3 + 255 ) & ~(size_t)255 ) );
    if ( x || type )
    {
        k_unpack_xt<<<div_up( count, tb ), tb, 0, s>>>( ctx->xt, x ? d3 : nullptr,
                                                        type ? dt : nullptr, first, count );
        CBMD_LAUNCH_CHECK( ctx );
        if ( x )
            CBMD_CUDA( cudaMemcpyAsync( x, d3, b3, cudaMemcpyDeviceToHost, s ) );
        if ( type )
            CBMD_CUDA( cudaMemcpyAsync( type, dt, (size_t)count * sizeof( int ),
                                        cudaMemcpyDeviceToHost, s ) );
        CBMD_CUDA( cudaStreamSynchronize( s ) );
    }
    if ( v )
    {
        k_soa_to_aos<<<div_up( 3 * count, tb ), tb, 0, s>>>( ctx->v, ctx->cap, d3, first, count );
        CBMD_LAUNCH_CHECK( ctx );
        CBMD_CUDA( cudaMemcpyAsync( v, d3, b3, cudaMemcpyDeviceToHost, s ) );
        CBMD_CUDA( cudaStreamSynchronize( s ) );
    }
    if ( f )
    {
        k_soa_to_aos<<<div_up( 3 * count, tb ), tb, 0, s>>>( ctx->f, ctx->cap, d3, first, count );
        CBMD_LAUNCH_CHECK( ctx );
        CBMD_CUDA( cudaMemcpyAsync( f, d3, b3, cudaMemcpyDeviceToHost, s ) );
        CBMD_CUDA( cudaStreamSynchronize( s ) );
    }
    if ( id )
        CBMD_CUDA( cudaMemcpyAsync( id, ctx->id + first, (size_t)count * sizeof( int ),
                                    cudaMemcpyDeviceToHost, s ) );
    if ( q )
        CBMD_CUDA( cudaMemcpyAsync( q, ctx->q + first, (size_t)count * sizeof( double ),
                                    cudaMemcpyDeviceToHost, s ) );
    CBMD_CUDA( cudaStreamSynchronize( s ) );
    CBMD_API_END
}

extern "C" int cbmd_get_counts( cbmd_ctx *ctx, int *n_local, int *n_ghost )
{
    CBMD_API_BEGIN
    if ( n_local )
        *n_local = ctx->n_local;
    if ( n_ghost )
        *n_ghost = ctx->n_ghost;
    CBMD_API_END
}
