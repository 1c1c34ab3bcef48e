// cbmd_comm.cu — spatial-decomposition exchange: PBC wrap / migration of owned atoms,
// 6-phase ghost build, per-step ghost refresh and reverse force accumulation.
// Replaces Comm<t_System> (reference src/comm_mpi.h:125-353, src/comm_mpi_impl.h:52-441)
// and the [Cabana] Distributor/migrate + Halo/gather/scatter it delegates to.  MPI is
// replaced by NCCL send/recv groups (one process per GPU) or, between several contexts of
// one process, by the in-process hub (cbmd_hub_create; struct Xfer below is the one place
// that knows the difference); a phase whose face neighbour is this rank itself (one rank
// in that dimension) is a local copy.
//
// The six phases (+x,-x,+y,-y,+z,-z) are kept exactly — including forwarding of
// earlier-phase ghosts and the odd-phase exclusion of the ghosts just received
// (comm_mpi_impl.h:301-303) — so the ghost SET equals the reference's.  Send lists
// are produced by an ordered (ascending index) stream compaction, so ghost order is
// deterministic, unlike the reference's atomic-append order.  When every face
// neighbour is this rank (single GPU) each ghost also records its root owner and
// accumulated image so update_halo is ONE gather kernel instead of six dependent
// phases; each coordinate is shifted at most once, so the values are bit-identical.
#include "cbmd_internal.cuh"

__global__ void k_fill3( double *__restrict__ soa, int cap, int first, int n, double val );

static int rank_of_pos( const int grid[3], int i, int j, int k )
{
    i = ( i % grid[0] + grid[0] ) % grid[0];
    j = ( j % grid[1] + grid[1] ) % grid[1];
    k = ( k % grid[2] + grid[2] ) % grid[2];
    return ( i * grid[1] + j ) * grid[2] + k; // MPI_Cart order: last dimension fastest
}

// ---------------------------------------------------------------------------
// Transport: one group of point-to-point messages, all in flight together.  Between processes
// it is an NCCL group (ncclSend/ncclRecv on the given stream); between contexts of one process
// (cbmd_comm_init_hub) the messages go through the hub's FIFOs as device-to-device copies
// ordered by events.  Both pair messages of one (source, destination) in posting order.
// ---------------------------------------------------------------------------
static void hub_fail( cbmd_hub *hub, const char *what )
{
    hub->failed = true;
    hub->cv.notify_all();
    throw CbmdError( std::string( "hub: " ) + what + " (a peer did not reach the matching call within " +
                     std::to_string( (int)hub->timeout_s ) + " s, or failed earlier)" );
}

template <class Pred>
static void hub_wait( cbmd_hub *hub, std::unique_lock<std::mutex> &lk, Pred pred, const char *what )
{
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>( hub->timeout_s );
    while ( !pred() )
    {
        if ( hub->failed )
            hub_fail( hub, what );
        if ( hub->cv.wait_until( lk, deadline ) == std::cv_status::timeout && !pred() )
            hub_fail( hub, what );
    }
}

// rank-synchronous barrier over the hub (generation counted)
static void hub_barrier( cbmd_hub *hub, std::unique_lock<std::mutex> &lk )
{
    const long gen = hub->generation;
    if ( ++hub->arrived == hub->nranks )
    {
        hub->arrived = 0;
        hub->generation++;
        hub->cv.notify_all();
        return;
    }
    hub_wait( hub, lk, [&] { return hub->generation != gen; }, "barrier" );
}

// release the `done` events whose senders have ordered themselves behind them
static void hub_collect( cbmd_ctx *ctx, bool all )
{
    auto &v = ctx->hub_done;
    size_t keep = 0;
    for ( size_t k = 0; k < v.size(); k++ )
    {
        if ( all || v[k]->state == 2 )
            cudaEventDestroy( v[k]->done );
        else
            v[keep++] = v[k];
    }
    v.resize( keep );
}

struct Xfer
{
    struct Op
    {
        void *ptr;
        size_t bytes;
        int peer;
    };
    cbmd_ctx *ctx;
    cudaStream_t s;
    std::vector<Op> sends, recvs;
    Xfer( cbmd_ctx *c, cudaStream_t st ) : ctx( c ), s( st ) {}
    void send( const void *p, size_t bytes, int peer )
    {
        if ( bytes > 0 )
            sends.push_back( { const_cast<void *>( p ), bytes, peer } );
    }
    void recv( void *p, size_t bytes, int peer )
    {
        if ( bytes > 0 )
            recvs.push_back( { p, bytes, peer } );
    }
    void run()
    {
        if ( sends.empty() && recvs.empty() )
            return;
        if ( !ctx->hub )
        {
            CBMD_REQUIRE( ctx->nccl != nullptr, "no communicator: call cbmd_comm_init first" );
            CBMD_NCCL( ncclGroupStart() );
            for ( const Op &o : sends )
                CBMD_NCCL( ncclSend( o.ptr, o.bytes, ncclChar, o.peer, ctx->nccl, s ) );
            for ( const Op &o : recvs )
                CBMD_NCCL( ncclRecv( o.ptr, o.bytes, ncclChar, o.peer, ctx->nccl, s ) );
            CBMD_NCCL( ncclGroupEnd() );
            return;
        }
        cbmd_hub *hub = ctx->hub;
        const int np = hub->nranks, me = ctx->rank;
        cudaEvent_t ready = nullptr;
        std::vector<std::shared_ptr<HubMsg>> posted;
        if ( !sends.empty() )
        {
            CBMD_CUDA( cudaEventCreateWithFlags( &ready, cudaEventDisableTiming ) );
            CBMD_CUDA( cudaEventRecord( ready, s ) );
        }
        std::unique_lock<std::mutex> lk( hub->m );
        hub_collect( ctx, false );
        for ( const Op &o : sends )
        {
            auto m = std::make_shared<HubMsg>();
            m->ptr = o.ptr;
            m->bytes = o.bytes;
            m->ready = ready;
            hub->q[(size_t)me * np + o.peer].push_back( m );
            posted.push_back( m );
        }
        hub->cv.notify_all();
        for ( const Op &o : recvs )
        {
            auto &q = hub->q[(size_t)o.peer * np + me];
            hub_wait( hub, lk, [&] { return !q.empty(); }, "receive" );
            std::shared_ptr<HubMsg> m = q.front();
            q.pop_front();
            if ( m->bytes != o.bytes )
                hub_fail( hub, "message size differs from what the receiver expects" );
            CBMD_CUDA( cudaStreamWaitEvent( s, m->ready, 0 ) );
            CBMD_CUDA( cudaMemcpyAsync( o.ptr, m->ptr, o.bytes, cudaMemcpyDefault, s ) );
            CBMD_CUDA( cudaEventCreateWithFlags( &m->done, cudaEventDisableTiming ) );
            CBMD_CUDA( cudaEventRecord( m->done, s ) );
            m->state = 1;
            ctx->hub_done.push_back( m );
            hub->cv.notify_all();
        }
        for ( auto &m : posted )
        {
            hub_wait( hub, lk, [&] { return m->state >= 1; }, "send" );
            CBMD_CUDA( cudaStreamWaitEvent( s, m->done, 0 ) );
            m->state = 2;
        }
        lk.unlock();
        if ( ready )
            CBMD_CUDA( cudaEventDestroy( ready ) ); // every receiver has ordered its stream behind it
    }
};

// ---------------------------------------------------------------------------
// ordered stream compaction of indices i in [0,n) with coordinate test in dim d:
//   mode 0: x_d >= thr   mode 1: x_d <= thr   mode 2: x_d > thr   mode 3: x_d < thr
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool face_test( const XT &r, int d, int mode, double thr )
{
    const double c = d == 0 ? r.x : ( d == 1 ? r.y : r.z );
    switch ( mode )
    {
    case 0:
        return c >= thr;
    case 1:
        return c <= thr;
    case 2:
        return c > thr;
    default:
        return c < thr;
    }
}

__global__ void __launch_bounds__( 256 )
    k_face_flags( const XT *__restrict__ xt, int n, int d, int mode, double thr,
                  int *__restrict__ flags )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i < n )
        flags[i] = face_test( xt[i], d, mode, thr ) ? 1 : 0;
    if ( i == n )
        flags[n] = 0;
}

__global__ void __launch_bounds__( 256 )
    k_select_scatter( const XT *__restrict__ xt, int n, int d, int mode, double thr,
                      const int *__restrict__ pos, int *__restrict__ out, int out_cap )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    if ( face_test( xt[i], d, mode, thr ) )
    {
        const int p = pos[i];
        if ( p < out_cap )
            out[p] = i;
    }
}

// Both faces of one dimension at once.  The two phases of a dimension never feed each other
// (the odd phase skips what the even phase has just received, comm_mpi_impl.h:301-303; an atom
// that arrived through the low face cannot leave through it), so their selections run on the
// same state and ONE count swap (one NCCL group) + ONE host synchronisation serve both —
// half the blocking round trips of a rebuild step.  pos[k] (device, n+1 ints) holds the
// exclusive scan of phase k's flags.
static void select_pair( cbmd_ctx *ctx, int n, int d, const int mode[2], const double thr[2], bool remote,
                         const int peer_send[2], const int peer_recv[2], int *pos[2], int n_send[2],
                         int n_recv[2] )
{
    cudaStream_t s = ctx->stream;
    int *d_cnt = ctx->d_flags + 24, *d_in = ctx->d_flags + 26;
    pos[0] = pos[1] = nullptr;
    if ( n > 0 )
    {
        const size_t pb = ( (size_t)( n + 1 ) * sizeof( int ) + 255 ) & ~(size_t)255;
        char *st = (char *)cbmd_scratch( ctx, 2 * pb );
        for ( int k = 0; k < 2; k++ )
        {
            pos[k] = (int *)( st + k * pb );
            k_face_flags<<<div_up( n + 1, 256 ), 256, 0, s>>>( ctx->xt, n, d, mode[k], thr[k], pos[k] );
            CBMD_LAUNCH_CHECK( ctx );
            cbmd_exclusive_scan_int( ctx, pos[k], n );
            CBMD_CUDA( cudaMemcpyAsync( d_cnt + k, pos[k] + n, sizeof( int ), cudaMemcpyDeviceToDevice, s ) );
        }
    }
    else
        CBMD_CUDA( cudaMemsetAsync( d_cnt, 0, 2 * sizeof( int ), s ) );
    CBMD_CUDA( cudaMemcpyAsync( ctx->h_pinned_i + 8, d_cnt, 2 * sizeof( int ), cudaMemcpyDeviceToHost, s ) );
    if ( remote )
    {
        // with two ranks in this dimension both messages go to the same peer: sends and
        // receives are issued in phase order on both sides, which is how NCCL pairs them
        Xfer x( ctx, s );
        for ( int k = 0; k < 2; k++ )
            x.send( d_cnt + k, sizeof( int ), peer_send[k] );
        for ( int k = 0; k < 2; k++ )
            x.recv( d_in + k, sizeof( int ), peer_recv[k] );
        x.run();
        CBMD_CUDA( cudaMemcpyAsync( ctx->h_pinned_i + 10, d_in, 2 * sizeof( int ), cudaMemcpyDeviceToHost, s ) );
    }
    CBMD_CUDA( cudaStreamSynchronize( s ) );
    for ( int k = 0; k < 2; k++ )
    {
        n_send[k] = ctx->h_pinned_i[8 + k];
        n_recv[k] = remote ? ctx->h_pinned_i[10 + k] : n_send[k];
    }
}

// ---------------------------------------------------------------------------
// Comm::exchange — TagExchangeSelf (comm_mpi.h:141-168)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__( 256 )
    k_exchange_self( XT *__restrict__ xt, int n, double Lx, double Ly, double Lz, int wx, int wy,
                     int wz )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    XT r = xt[i];
    bool ch = false;
    if ( wx )
    {
        const double c = r.x;
        if ( c > Lx )
        {
            r.x -= Lx;
            ch = true;
        }
        if ( c < 0 )
        {
            r.x += Lx;
            ch = true;
        }
    }
    if ( wy )
    {
        const double c = r.y;
        if ( c > Ly )
        {
            r.y -= Ly;
            ch = true;
        }
        if ( c < 0 )
        {
            r.y += Ly;
            ch = true;
        }
    }
    if ( wz )
    {
        const double c = r.z;
        if ( c > Lz )
        {
            r.z -= Lz;
            ch = true;
        }
        if ( c < 0 )
        {
            r.z += Lz;
            ch = true;
        }
    }
    if ( ch )
        xt[i] = r;
}

// migration tuple: the reference's 88-byte record {x[3],v[3],f[3],type,id,q}
struct alignas( 8 ) MigTuple
{
    double x[3], v[3], f[3];
    int type, id;
    double q;
};

// both faces of one dimension: stayers keep their order, leavers through the high / low face
// are packed into their own send segments (PBC shift applied on the edge ranks)
__global__ void __launch_bounds__( 256 )
    k_migrate_split2( const XT *__restrict__ xt, const double *__restrict__ v, const double *__restrict__ f,
                      const int *__restrict__ id, const double *__restrict__ q, int cap, int n, int d,
                      double thr_hi, double thr_lo, double shift_hi, double shift_lo,
                      const int *__restrict__ pos_hi, const int *__restrict__ pos_lo, XT *__restrict__ xt_o,
                      double *__restrict__ v_o, double *__restrict__ f_o, int *__restrict__ id_o,
                      double *__restrict__ q_o, MigTuple *__restrict__ out_hi, MigTuple *__restrict__ out_lo )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    XT r = xt[i];
    const double c = d == 0 ? r.x : ( d == 1 ? r.y : r.z );
    const bool hi = c > thr_hi, lo = c < thr_lo; // comm_mpi.h:173,185,...: strict tests
    if ( hi || lo )
    {
        const double shift = hi ? shift_hi : shift_lo;
        if ( d == 0 )
            r.x += shift;
        else if ( d == 1 )
            r.y += shift;
        else
            r.z += shift;
        MigTuple t;
        t.x[0] = r.x;
        t.x[1] = r.y;
        t.x[2] = r.z;
        for ( int k = 0; k < 3; k++ )
        {
            t.v[k] = v[(size_t)k * cap + i];
            t.f[k] = f[(size_t)k * cap + i];
        }
        t.type = (int)r.t;
        t.id = id[i];
        t.q = q[i];
        if ( hi )
            out_hi[pos_hi[i]] = t;
        else
            out_lo[pos_lo[i]] = t;
    }
    else
    {
        const int o = i - pos_hi[i] - pos_lo[i];
        xt_o[o] = r;
        for ( int k = 0; k < 3; k++ )
        {
            v_o[(size_t)k * cap + o] = v[(size_t)k * cap + i];
            f_o[(size_t)k * cap + o] = f[(size_t)k * cap + i];
        }
        id_o[o] = id[i];
        q_o[o] = q[i];
    }
}

__global__ void __launch_bounds__( 256 )
    k_migrate_unpack( const MigTuple *__restrict__ in, int n, int first, int cap,
                      XT *__restrict__ xt, double *__restrict__ v, double *__restrict__ f,
                      int *__restrict__ id, double *__restrict__ q )
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if ( k >= n )
        return;
    const MigTuple t = in[k];
    XT r;
    r.x = t.x[0];
    r.y = t.x[1];
    r.z = t.x[2];
    r.t = t.type;
    const int o = first + k;
    xt[o] = r;
    for ( int c = 0; c < 3; c++ )
    {
        v[(size_t)c * cap + o] = t.v[c];
        f[(size_t)c * cap + o] = t.f[c];
    }
    id[o] = t.id;
    q[o] = t.q;
}

static void ensure_buf( double *&buf, size_t &have, size_t bytes, cudaStream_t s )
{
    if ( bytes <= have )
        return;
    if ( buf )
    {
        CBMD_CUDA( cudaStreamSynchronize( s ) );
        CBMD_CUDA( cudaFree( buf ) );
    }
    buf = nullptr;
    have = bytes + bytes / 4 + ( 1 << 16 );
    CBMD_CUDA( cudaMalloc( &buf, have ) );
}

static void phase_peers( const cbmd_ctx *ctx, int ph, int &peer_send, int &peer_recv )
{
    // comm_mpi_impl.h:87-99: send +x,-x,+y,-y,+z,-z; recv = send of the opposite phase
    int dlt[3] = { 0, 0, 0 };
    dlt[ph / 2] = ( ph % 2 == 0 ) ? 1 : -1;
    peer_send = rank_of_pos( ctx->grid, ctx->pos[0] + dlt[0], ctx->pos[1] + dlt[1],
                             ctx->pos[2] + dlt[2] );
    peer_recv = rank_of_pos( ctx->grid, ctx->pos[0] - dlt[0], ctx->pos[1] - dlt[1],
                             ctx->pos[2] - dlt[2] );
}

extern "C" int cbmd_exchange( cbmd_ctx *ctx, int *n_sent_global )
{
    CBMD_API_BEGIN
    TimedRegion timed__( ctx, CBMD_T_COMM );
    CBMD_REQUIRE( ctx->have_domain, "cbmd_set_domain must be called before cbmd_exchange" );
    cbmd_materialize_zero_force( ctx );
    cbmd_bump_epoch( ctx, true, true );
    cudaStream_t s = ctx->stream;
    // system->resize(N_local): ghosts are dropped (comm_mpi_impl.h:196-197)
    ctx->n_ghost = 0;
    ctx->have_halo = false;
    ctx->nb_n = 0;
    ctx->nb_ntot = 0;
    int n = ctx->n_local;
    if ( n > 0 )
    {
        k_exchange_self<<<div_up( n, 256 ), 256, 0, s>>>( ctx->xt, n, ctx->gext[0], ctx->gext[1],
                                                          ctx->gext[2], ctx->grid[0] == 1,
                                                          ctx->grid[1] == 1, ctx->grid[2] == 1 );
        CBMD_LAUNCH_CHECK( ctx );
    }
    int total_sent = 0;
    for ( int d = 0; d < 3 && ctx->nranks > 1; d++ )
    {
        if ( ctx->grid[d] <= 1 )
            continue;
        // the +d and -d phases of the reference (comm_mpi_impl.h:200-262) in one step
        int peer_send[2], peer_recv[2];
        for ( int k = 0; k < 2; k++ )
            phase_peers( ctx, 2 * d + k, peer_send[k], peer_recv[k] );
        // strict tests (comm_mpi.h:173,185,...): x > local_hi (even) / x < local_lo (odd)
        const int mode[2] = { 2, 3 };
        const double thr[2] = { ctx->lhi[d], ctx->llo[d] };
        const double shift_hi = ctx->pos[d] == ctx->grid[d] - 1 ? -ctx->gext[d] : 0.0;
        const double shift_lo = ctx->pos[d] == 0 ? ctx->gext[d] : 0.0;
        n = ctx->n_local;
        int *pos[2];
        int n_send[2], n_recv[2];
        select_pair( ctx, n, d, mode, thr, true, peer_send, peer_recv, pos, n_send, n_recv );
        const int ns = n_send[0] + n_send[1], nr = n_recv[0] + n_recv[1];
        ensure_buf( ctx->sendbuf, ctx->sendbuf_bytes, (size_t)( ns + 1 ) * sizeof( MigTuple ), s );
        ensure_buf( ctx->recvbuf, ctx->recvbuf_bytes, (size_t)( nr + 1 ) * sizeof( MigTuple ), s );
        MigTuple *sb = (MigTuple *)ctx->sendbuf, *rb = (MigTuple *)ctx->recvbuf;
        if ( ns > 0 )
        {
            // pos lives in scratch: ensure_capacity below must not run before the split
            k_migrate_split2<<<div_up( n, 256 ), 256, 0, s>>>(
                ctx->xt, ctx->v, ctx->f, ctx->id, ctx->q, ctx->cap, n, d, thr[0], thr[1], shift_hi, shift_lo,
                pos[0], pos[1], ctx->xt_alt, ctx->v_alt, ctx->f_alt, ctx->id_alt, ctx->q_alt, sb, sb + n_send[0] );
            CBMD_LAUNCH_CHECK( ctx );
            std::swap( ctx->xt, ctx->xt_alt );
            std::swap( ctx->v, ctx->v_alt );
            std::swap( ctx->f, ctx->f_alt );
            std::swap( ctx->id, ctx->id_alt );
            std::swap( ctx->q, ctx->q_alt );
        }
        {
            Xfer x( ctx, s );
            x.send( sb, (size_t)n_send[0] * sizeof( MigTuple ), peer_send[0] );
            x.send( sb + n_send[0], (size_t)n_send[1] * sizeof( MigTuple ), peer_send[1] );
            x.recv( rb, (size_t)n_recv[0] * sizeof( MigTuple ), peer_recv[0] );
            x.recv( rb + n_recv[0], (size_t)n_recv[1] * sizeof( MigTuple ), peer_recv[1] );
            x.run();
        }
        const int n_keep = n - ns;
        ctx->n_local = n_keep; // so a regrow copies only live rows
        cbmd_ensure_capacity( ctx, n_keep + nr );
        if ( nr > 0 )
        {
            // arrivals of the even phase first, then of the odd phase: the order the two
            // sequential phases of the reference produce
            k_migrate_unpack<<<div_up( nr, 256 ), 256, 0, s>>>( rb, nr, n_keep, ctx->cap, ctx->xt, ctx->v, ctx->f,
                                                               ctx->id, ctx->q );
            CBMD_LAUNCH_CHECK( ctx );
        }
        ctx->n_local = n_keep + nr;
        total_sent += ns;
    }
    if ( ctx->nranks > 1 )
        CBMD_REQUIRE( cbmd_reduce_sum_int( ctx, &total_sent, 1 ) == 0, cbmd_last_error() );
    if ( n_sent_global )
        *n_sent_global = total_sent;
    CBMD_API_END
}

// ---------------------------------------------------------------------------
// Comm::exchange_halo
// ---------------------------------------------------------------------------
// local (self-neighbour) ghost creation: gather by send list, shift, record owner/image
__global__ void __launch_bounds__( 256 )
    k_halo_make_self( XT *__restrict__ xt, int *__restrict__ id, const int *__restrict__ send_idx,
                      int n, int first, int n_local, int d, double shift,
                      int *__restrict__ owner, unsigned char *__restrict__ image,
                      int *__restrict__ grank, int my_rank )
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if ( k >= n )
        return;
    const int sidx = send_idx[k];
    XT r = ld_xt( xt + sidx );
    unsigned char img = 0;
    int own = sidx, rk = my_rank;
    if ( sidx >= n_local )
    {
        own = owner[sidx - n_local];
        img = image[sidx - n_local];
        rk = grank[sidx - n_local];
    }
    if ( shift != 0.0 )
    {
        if ( d == 0 )
            r.x += shift;
        else if ( d == 1 )
            r.y += shift;
        else
            r.z += shift;
        img |= (unsigned char)( ( shift > 0.0 ? 1u : 2u ) << ( 2 * d ) );
    }
    const int g = first + k;
    xt[g] = r;
    id[g] = id[sidx];
    owner[g - n_local] = own;
    image[g - n_local] = img;
    grank[g - n_local] = rk;
}

// Ghost record of a remote phase of the ghost build: position + type, id, and the ROOT of the
// atom (rank that owns it, its index there, accumulated image) so that the per-step refresh can
// fetch every ghost straight from its root instead of replaying the forwarding phases.
struct alignas( 16 ) HaloRec
{
    double x, y, z;
    long long t;
    int id, root_rank, root_idx, image;
};

__global__ void __launch_bounds__( 256 )
    k_halo_pack_rec( const XT *__restrict__ xt, const int *__restrict__ id, const int *__restrict__ owner,
                     const int *__restrict__ grank, const unsigned char *__restrict__ image,
                     const int *__restrict__ send_idx, int n, int n_local, int my_rank,
                     HaloRec *__restrict__ out )
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if ( k >= n )
        return;
    const int sidx = send_idx[k];
    const XT r = ld_xt( xt + sidx );
    HaloRec h;
    h.x = r.x;
    h.y = r.y;
    h.z = r.z;
    h.t = r.t;
    h.id = id[sidx];
    h.root_rank = my_rank;
    h.root_idx = sidx;
    h.image = 0;
    if ( sidx >= n_local )
    {
        h.root_rank = grank[sidx - n_local];
        h.root_idx = owner[sidx - n_local];
        h.image = image[sidx - n_local];
    }
    out[k] = h;
}

// receiver side: TagHaloPBC shift (comm_mpi.h:323-353) + bookkeeping of the root
__global__ void __launch_bounds__( 256 )
    k_halo_unpack_rec( const HaloRec *__restrict__ in, int n, int first, int n_local, int d, double shift,
                       XT *__restrict__ xt, int *__restrict__ id, int *__restrict__ owner,
                       int *__restrict__ grank, unsigned char *__restrict__ image )
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if ( k >= n )
        return;
    const HaloRec h = in[k];
    XT r;
    r.x = h.x;
    r.y = h.y;
    r.z = h.z;
    r.t = h.t;
    unsigned img = (unsigned)h.image;
    if ( shift != 0.0 )
    {
        if ( d == 0 )
            r.x += shift;
        else if ( d == 1 )
            r.y += shift;
        else
            r.z += shift;
        img |= ( shift > 0.0 ? 1u : 2u ) << ( 2 * d );
    }
    const int g = first + k;
    xt[g] = r;
    id[g] = h.id;
    owner[g - n_local] = h.root_idx;
    grank[g - n_local] = h.root_rank;
    image[g - n_local] = (unsigned char)img;
}

// ---- one-stage plan: ghosts grouped by root rank (stable), request lists for the roots
__global__ void __launch_bounds__( 256 )
    k_rank_flags( const int *__restrict__ grank, int n, int p, int *__restrict__ flags )
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if ( g < n )
        flags[g] = grank[g] == p ? 1 : 0;
    if ( g == n )
        flags[n] = 0;
}

__global__ void __launch_bounds__( 256 )
    k_rank_slots( const int *__restrict__ grank, const int *__restrict__ owner, int n, int p,
                  const int *__restrict__ pos, int base, int *__restrict__ slot, int *__restrict__ req )
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if ( g >= n || grank[g] != p )
        return;
    slot[g] = base + pos[g];
    req[base + pos[g]] = owner[g];
}

// per step, export side: the positions the other ranks want, three doubles per atom
__global__ void __launch_bounds__( 256 )
    k_halo_pack_flat( const XT *__restrict__ xt, const int *__restrict__ export_idx, int n,
                      double *__restrict__ out )
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if ( k >= n )
        return;
    const XT r = ld_xt( xt + export_idx[k] );
    out[3 * (size_t)k] = r.x;
    out[3 * (size_t)k + 1] = r.y;
    out[3 * (size_t)k + 2] = r.z;
}

// per step, import side: every ghost = its root's position (from the receive buffer, or from
// this rank's own atoms) + the accumulated image shift; each coordinate is shifted at most
// once, so the values are bit-identical to the forwarding scheme's
__global__ void __launch_bounds__( 256 )
    k_halo_unpack_flat( XT *__restrict__ xt, int n_local, int n_ghost, const int *__restrict__ owner,
                        const int *__restrict__ grank, const unsigned char *__restrict__ image,
                        const int *__restrict__ slot, const double *__restrict__ recv3, int my_rank,
                        double Lx, double Ly, double Lz, const MirrorPtrs mir )
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if ( g >= n_ghost )
        return;
    XT r = xt[n_local + g]; // keeps the type
    if ( grank[g] == my_rank )
    {
        const XT o = ld_xt( xt + owner[g] );
        r.x = o.x;
        r.y = o.y;
        r.z = o.z;
    }
    else
    {
        const size_t k = 3 * (size_t)slot[g];
        r.x = recv3[k];
        r.y = recv3[k + 1];
        r.z = recv3[k + 2];
    }
    const unsigned img = image[g];
    const unsigned ix = img & 3u, iy = ( img >> 2 ) & 3u, iz = ( img >> 4 ) & 3u;
    if ( ix )
        r.x += ( ix == 1u ? Lx : -Lx );
    if ( iy )
        r.y += ( iy == 1u ? Ly : -Ly );
    if ( iz )
        r.z += ( iz == 1u ? Lz : -Lz );
    xt[n_local + g] = r;
    mirror_store( mir, n_local + g, r );
}

__global__ void __launch_bounds__( 256 )
    k_halo_pack( const XT *__restrict__ xt, const int *__restrict__ id,
                 const int *__restrict__ send_idx, int n, XT *__restrict__ out_xt,
                 int *__restrict__ out_id )
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if ( k >= n )
        return;
    const int sidx = send_idx[k];
    out_xt[k] = ld_xt( xt + sidx );
    if ( out_id )
        out_id[k] = id[sidx];
}

// TagHaloPBC (comm_mpi.h:323-353) on a freshly received segment
__global__ void __launch_bounds__( 256 )
    k_halo_shift( XT *__restrict__ xt, int first, int n, int d, double shift )
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if ( k >= n )
        return;
    double *c = &xt[first + k].x + d;
    *c += shift;
}

// ---------------------------------------------------------------------------
// One-stage refresh plan (multi-rank).  The reference replays its six forwarding phases
// every step (comm_mpi_impl.h:369-408): x ghosts must have landed before the y phase can
// forward them, y before z — three dependent exchanges.  Every ghost is an image of ONE
// owned atom somewhere (its root, carried along while the ghost shell is built), so the
// refresh can fetch it from the root directly: one pack kernel, ONE NCCL group with all peers
// in flight together, one unpack kernel that also applies the accumulated shift.  The plan is
// made once per rebuild: ghosts grouped by root rank in ghost order (stable), counts
// exchanged with one all-gather, request lists (root indices) sent to the roots.
// ---------------------------------------------------------------------------
static void build_flat_plan( cbmd_ctx *ctx )
{
    cudaStream_t s = ctx->stream;
    const int np = ctx->nranks, me = ctx->rank, ng = ctx->n_ghost;
    ctx->rcnt.assign( np, 0 );
    ctx->roff.assign( np, 0 );
    ctx->scnt.assign( np, 0 );
    ctx->soff.assign( np, 0 );
    // import side: slots of the receive buffer, peer by peer; request lists in the same order
    int *req = nullptr;
    int *pos = nullptr;
    if ( ng > 0 )
    {
        const size_t pb = ( (size_t)( ng + 1 ) * sizeof( int ) + 255 ) & ~(size_t)255;
        char *st = (char *)cbmd_scratch( ctx, 2 * pb + (size_t)np * sizeof( int ) );
        pos = (int *)st;
        req = (int *)( st + pb );
    }
    int *d_cnt = ctx->d_flags + 32;        // my counts per root rank
    int *d_all = ctx->d_flags + 32 + 64;   // everybody's counts [np][np]
    CBMD_REQUIRE( np <= 64, "one-stage halo plan supports at most 64 ranks" );
    CBMD_CUDA( cudaMemsetAsync( d_cnt, 0, (size_t)np * sizeof( int ), s ) );
    // counts first (device to device), then the slots once the offsets are known on the host
    for ( int p = 0; p < np && ng > 0; p++ )
    {
        if ( p == me )
            continue;
        k_rank_flags<<<div_up( ng + 1, 256 ), 256, 0, s>>>( ctx->ghost_rank, ng, p, pos );
        CBMD_LAUNCH_CHECK( ctx );
        cbmd_exclusive_scan_int( ctx, pos, ng );
        CBMD_CUDA( cudaMemcpyAsync( d_cnt + p, pos + ng, sizeof( int ), cudaMemcpyDeviceToDevice, s ) );
    }
    if ( ctx->hub )
    {
        Xfer x( ctx, s );
        for ( int p = 0; p < np; p++ )
            if ( p != me )
            {
                x.send( d_cnt, (size_t)np * sizeof( int ), p );
                x.recv( d_all + (size_t)p * np, (size_t)np * sizeof( int ), p );
            }
        CBMD_CUDA( cudaMemcpyAsync( d_all + (size_t)me * np, d_cnt, (size_t)np * sizeof( int ),
                                    cudaMemcpyDeviceToDevice, s ) );
        x.run();
    }
    else
        CBMD_NCCL( ncclAllGather( d_cnt, d_all, np, ncclInt, ctx->nccl, s ) );
    std::vector<int> all( (size_t)np * np );
    CBMD_CUDA( cudaMemcpyAsync( all.data(), d_all, all.size() * sizeof( int ), cudaMemcpyDeviceToHost, s ) );
    CBMD_CUDA( cudaStreamSynchronize( s ) );
    int ro = 0, so = 0;
    for ( int p = 0; p < np; p++ )
    {
        ctx->rcnt[p] = all[(size_t)me * np + p]; // ghosts of mine rooted on p
        ctx->scnt[p] = all[(size_t)p * np + me]; // atoms of mine that p wants
        ctx->roff[p] = ro;
        ctx->soff[p] = so;
        ro += ctx->rcnt[p];
        so += ctx->scnt[p];
    }
    ctx->n_import = ro;
    ctx->n_export = so;
    if ( so > ctx->export_cap )
    {
        if ( ctx->export_idx )
            CBMD_CUDA( cudaFree( ctx->export_idx ) );
        ctx->export_cap = so + so / 4 + 256;
        CBMD_CUDA( cudaMalloc( &ctx->export_idx, (size_t)ctx->export_cap * sizeof( int ) ) );
    }
    for ( int p = 0; p < np && ng > 0; p++ )
    {
        if ( p == me || ctx->rcnt[p] == 0 )
            continue;
        k_rank_flags<<<div_up( ng + 1, 256 ), 256, 0, s>>>( ctx->ghost_rank, ng, p, pos );
        CBMD_LAUNCH_CHECK( ctx );
        cbmd_exclusive_scan_int( ctx, pos, ng );
        k_rank_slots<<<div_up( ng, 256 ), 256, 0, s>>>( ctx->ghost_rank, ctx->ghost_owner, ng, p, pos,
                                                        ctx->roff[p], ctx->ghost_slot, req );
        CBMD_LAUNCH_CHECK( ctx );
    }
    // request lists to the roots; what the others want from me comes back
    {
        Xfer x( ctx, s );
        for ( int p = 0; p < np; p++ )
        {
            if ( p == me )
                continue;
            x.send( req + ctx->roff[p], (size_t)ctx->rcnt[p] * sizeof( int ), p );
            x.recv( ctx->export_idx + ctx->soff[p], (size_t)ctx->scnt[p] * sizeof( int ), p );
        }
        x.run();
    }
    ensure_buf( ctx->sendbuf, ctx->sendbuf_bytes, 3 * (size_t)( so + 1 ) * sizeof( double ), s );
    ensure_buf( ctx->recvbuf, ctx->recvbuf_bytes, 3 * (size_t)( ro + 1 ) * sizeof( double ), s );
    // (req lives in scratch: later users of scratch are ordered behind the sends on this stream)
    ctx->flat_mp_ok = true;
}

extern "C" int cbmd_exchange_halo( cbmd_ctx *ctx, double comm_depth )
{
    CBMD_API_BEGIN
    TimedRegion timed__( ctx, CBMD_T_COMM );
    CBMD_REQUIRE( ctx->have_domain, "cbmd_set_domain must be called before cbmd_exchange_halo" );
    CBMD_REQUIRE( comm_depth > 0, "comm depth must be positive" );
    cbmd_materialize_zero_force( ctx );
    cbmd_bump_epoch( ctx, true, true );
    cudaStream_t s = ctx->stream;
    ctx->comm_depth = comm_depth;
    ctx->n_ghost = 0;
    ctx->nb_n = 0;
    ctx->nb_ntot = 0;
    bool all_self = true;
    for ( int d = 0; d < 3; d++ )
    {
        // the two phases of a dimension (+d then -d, comm_mpi_impl.h:280-367) in one step: the odd
        // phase never sends what the even phase has just received (:301-303), so both select
        // from the same atoms and share one count swap and one host synchronisation
        HaloPhase *P[2] = { &ctx->phase[2 * d], &ctx->phase[2 * d + 1] };
        int peer_send[2], peer_recv[2];
        for ( int k = 0; k < 2; k++ )
        {
            phase_peers( ctx, 2 * d + k, P[k]->peer_send, P[k]->peer_recv );
            peer_send[k] = P[k]->peer_send;
            peer_recv[k] = P[k]->peer_recv;
        }
        const bool self = ( peer_send[0] == ctx->rank );
        all_self = all_self && self;
        const int np = ctx->n_local + ctx->n_ghost;
        // comm_mpi.h:249,263,...: x >= hi - depth (even) / x <= lo + depth (odd)
        const int mode[2] = { 0, 1 };
        const double thr[2] = { ctx->lhi[d] - comm_depth, ctx->llo[d] + comm_depth };
        int *pos[2];
        int n_send[2], n_recv[2];
        select_pair( ctx, np, d, mode, thr, !self, peer_send, peer_recv, pos, n_send, n_recv );
        for ( int k = 0; k < 2; k++ )
        {
            P[k]->n_send = n_send[k];
            P[k]->n_recv = n_recv[k];
            if ( n_send[k] > P[k]->send_cap )
            {
                if ( P[k]->send_idx )
                    CBMD_CUDA( cudaFree( P[k]->send_idx ) );
                P[k]->send_cap = (int)( n_send[k] * 1.1 ) + 256; // comm_mpi_impl.h:313 growth policy
                CBMD_CUDA( cudaMalloc( &P[k]->send_idx, (size_t)P[k]->send_cap * sizeof( int ) ) );
            }
            if ( n_send[k] > 0 )
            {
                k_select_scatter<<<div_up( np, 256 ), 256, 0, s>>>( ctx->xt, np, d, mode[k], thr[k], pos[k],
                                                                    P[k]->send_idx, P[k]->send_cap );
                CBMD_LAUNCH_CHECK( ctx );
            }
        }
        // receiver-side PBC shift (TagHaloPBC)
        P[0]->shift = ctx->pos[d] == 0 ? -ctx->gext[d] : 0.0;
        P[1]->shift = ctx->pos[d] == ctx->grid[d] - 1 ? ctx->gext[d] : 0.0;
        const int first = ctx->n_local + ctx->n_ghost;
        P[0]->recv_first = first;
        P[1]->recv_first = first + n_recv[0];
        const int nr = n_recv[0] + n_recv[1], ns = n_send[0] + n_send[1];
        cbmd_ensure_capacity( ctx, first + nr ); // (pos is not used any more: scratch may move)
        if ( self )
        {
            for ( int k = 0; k < 2; k++ )
                if ( n_recv[k] > 0 )
                {
                    k_halo_make_self<<<div_up( n_recv[k], 256 ), 256, 0, s>>>(
                        ctx->xt, ctx->id, P[k]->send_idx, n_recv[k], P[k]->recv_first, ctx->n_local, d,
                        P[k]->shift, ctx->ghost_owner, ctx->ghost_image, ctx->ghost_rank, ctx->rank );
                    CBMD_LAUNCH_CHECK( ctx );
                }
        }
        else
        {
            ensure_buf( ctx->sendbuf, ctx->sendbuf_bytes, (size_t)( ns + 2 ) * sizeof( HaloRec ), s );
            ensure_buf( ctx->recvbuf, ctx->recvbuf_bytes, (size_t)( nr + 2 ) * sizeof( HaloRec ), s );
            HaloRec *sx[2] = { (HaloRec *)ctx->sendbuf, (HaloRec *)ctx->sendbuf + n_send[0] };
            HaloRec *rx[2] = { (HaloRec *)ctx->recvbuf, (HaloRec *)ctx->recvbuf + n_recv[0] };
            for ( int k = 0; k < 2; k++ )
                if ( n_send[k] > 0 )
                {
                    k_halo_pack_rec<<<div_up( n_send[k], 256 ), 256, 0, s>>>(
                        ctx->xt, ctx->id, ctx->ghost_owner, ctx->ghost_rank, ctx->ghost_image, P[k]->send_idx,
                        n_send[k], ctx->n_local, ctx->rank, sx[k] );
                    CBMD_LAUNCH_CHECK( ctx );
                }
            {
                Xfer x( ctx, s );
                for ( int k = 0; k < 2; k++ )
                    x.send( sx[k], (size_t)n_send[k] * sizeof( HaloRec ), peer_send[k] );
                for ( int k = 0; k < 2; k++ )
                    x.recv( rx[k], (size_t)n_recv[k] * sizeof( HaloRec ), peer_recv[k] );
                x.run();
            }
            for ( int k = 0; k < 2; k++ )
                if ( n_recv[k] > 0 )
                {
                    // ghosts of one phase are contiguous in the tail
                    k_halo_unpack_rec<<<div_up( n_recv[k], 256 ), 256, 0, s>>>(
                        rx[k], n_recv[k], P[k]->recv_first, ctx->n_local, d, P[k]->shift, ctx->xt, ctx->id,
                        ctx->ghost_owner, ctx->ghost_rank, ctx->ghost_image );
                    CBMD_LAUNCH_CHECK( ctx );
                }
        }
        ctx->n_ghost += nr;
    }
    // ghost v / f are never communicated (comm_mpi_impl.h:346-347); keep them defined
    if ( ctx->n_ghost > 0 )
    {
        k_fill3<<<div_up( ctx->n_ghost, 256 ), 256, 0, s>>>( ctx->v, ctx->cap, ctx->n_local,
                                                             ctx->n_ghost, 0.0 );
        CBMD_LAUNCH_CHECK( ctx );
        k_fill3<<<div_up( ctx->n_ghost, 256 ), 256, 0, s>>>( ctx->f, ctx->cap, ctx->n_local,
                                                             ctx->n_ghost, 0.0 );
        CBMD_LAUNCH_CHECK( ctx );
    }
    ctx->flat_halo_ok = all_self;
    ctx->flat_mp_ok = false;
    if ( !all_self && ctx->halo_stages == 1 )
        build_flat_plan( ctx );
    ctx->have_halo = true;
    CBMD_API_END
}

// ---------------------------------------------------------------------------
// Comm::update_halo
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__( 256 )
    k_halo_update_flat( XT *__restrict__ xt, int n_local, int n_ghost,
                        const int *__restrict__ owner, const unsigned char *__restrict__ image,
                        double Lx, double Ly, double Lz, const MirrorPtrs mir )
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if ( g >= n_ghost )
        return;
    XT r = ld_xt( xt + owner[g] );
    const unsigned img = image[g];
    const unsigned ix = img & 3u, iy = ( img >> 2 ) & 3u, iz = ( img >> 4 ) & 3u;
    if ( ix )
        r.x += ( ix == 1u ? Lx : -Lx );
    if ( iy )
        r.y += ( iy == 1u ? Ly : -Ly );
    if ( iz )
        r.z += ( iz == 1u ? Lz : -Lz );
    xt[n_local + g] = r;
    mirror_store( mir, n_local + g, r ); // gather mirror of the force sweeps (cbmd_force.cu)
}

__global__ void __launch_bounds__( 256 )
    k_halo_update_self( XT *__restrict__ xt, const int *__restrict__ send_idx, int n, int first,
                        int d, double shift )
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if ( k >= n )
        return;
    XT r = ld_xt( xt + send_idx[k] );
    if ( d == 0 )
        r.x += shift;
    else if ( d == 1 )
        r.y += shift;
    else
        r.z += shift;
    xt[first + k] = r;
}

extern "C" int cbmd_update_halo( cbmd_ctx *ctx )
{
    CBMD_API_BEGIN_NOJOIN
    TimedRegion timed__( ctx, CBMD_T_COMM );
    CBMD_REQUIRE( ctx->have_halo, "cbmd_exchange_halo must be called before cbmd_update_halo" );
    cbmd_join_halo( ctx ); // a previous refresh nobody consumed
    cbmd_bump_epoch( ctx, false, true );
    if ( ctx->n_ghost == 0 )
        return 0;
    if ( ctx->flat_halo_ok )
    {
        // the ghosts are images of owned atoms: the copy is only right if those are current
        const bool live = cbmd_mirror_live( ctx );
        k_halo_update_flat<<<div_up( ctx->n_ghost, 256 ), 256, 0, ctx->stream>>>(
            ctx->xt, ctx->n_local, ctx->n_ghost, ctx->ghost_owner, ctx->ghost_image, ctx->gext[0],
            ctx->gext[1], ctx->gext[2], cbmd_mirror_ptrs( ctx ) );
        CBMD_LAUNCH_CHECK( ctx );
        if ( live )
            ctx->mirror_ghost_epoch = ctx->epoch;
        return 0;
    }
    // multi-rank: the six dependent phases run on the comm stream so the force kernel can
    // start on the interior tiles meanwhile (cbmd_force_lj joins before the boundary tiles)
    // early: cbmd_integrate_initial has moved the atoms this refresh reads in a launch of their own
    // and recorded ev_x behind it — the refresh starts there, beside the bulk of the integrator
    const bool early = ctx->early_posted && ctx->flat_mp_ok && ctx->overlap;
    ctx->early_posted = false;
    ctx->halo_early = false;
    const bool ov = ctx->overlap && ( ctx->tiles_valid || early );
    cudaStream_t s = ov ? ctx->comm_stream : ctx->stream;
    if ( ov )
    {
        if ( !early )
            CBMD_CUDA( cudaEventRecord( ctx->ev_x, ctx->stream ) );
        CBMD_CUDA( cudaStreamWaitEvent( s, ctx->ev_x, 0 ) );
    }
    if ( ctx->flat_mp_ok )
    {
        // one stage: pack what the others want, one group with every peer, unpack + shift
        if ( ctx->n_export > 0 )
        {
            k_halo_pack_flat<<<div_up( ctx->n_export, 256 ), 256, 0, s>>>( ctx->xt, ctx->export_idx, ctx->n_export,
                                                                         ctx->sendbuf );
            CBMD_LAUNCH_CHECK( ctx );
        }
        {
            Xfer x( ctx, s );
            for ( int p = 0; p < ctx->nranks; p++ )
            {
                if ( p == ctx->rank )
                    continue;
                x.send( ctx->sendbuf + 3 * (size_t)ctx->soff[p], 3 * (size_t)ctx->scnt[p] * sizeof( double ), p );
                x.recv( ctx->recvbuf + 3 * (size_t)ctx->roff[p], 3 * (size_t)ctx->rcnt[p] * sizeof( double ), p );
            }
            x.run();
        }
        const bool live = cbmd_mirror_live( ctx );
        k_halo_unpack_flat<<<div_up( ctx->n_ghost, 256 ), 256, 0, s>>>(
            ctx->xt, ctx->n_local, ctx->n_ghost, ctx->ghost_owner, ctx->ghost_rank, ctx->ghost_image, ctx->ghost_slot,
            ctx->recvbuf, ctx->rank, ctx->gext[0], ctx->gext[1], ctx->gext[2], cbmd_mirror_ptrs( ctx ) );
        CBMD_LAUNCH_CHECK( ctx );
        if ( live )
            ctx->mirror_ghost_epoch = ctx->epoch;
        if ( ov )
        {
            CBMD_CUDA( cudaEventRecord( ctx->ev_halo, s ) );
            ctx->halo_pending = true;
            ctx->halo_early = early; // the force sweep joins and runs as one launch
        }
        return 0;
    }
    // The two phases of one dimension are independent of each other: the odd phase never
    // sends what the even phase has just received (cbmd_exchange_halo, comm_mpi_impl.h:301-303).
    // They therefore share ONE NCCL group (one launch, both directions in flight together);
    // the dimensions stay ordered because y forwards x ghosts and z forwards x and y ghosts.
    for ( int d = 0; d < 3; d++ )
    {
        HaloPhase *pair[2] = { &ctx->phase[2 * d], &ctx->phase[2 * d + 1] };
        if ( pair[0]->peer_send == ctx->rank )
        {
            for ( HaloPhase *P : pair )
                if ( P->n_recv > 0 )
                {
                    k_halo_update_self<<<div_up( P->n_recv, 256 ), 256, 0, s>>>(
                        ctx->xt, P->send_idx, P->n_recv, P->recv_first, d, P->shift );
                    CBMD_LAUNCH_CHECK( ctx );
                }
            continue;
        }
        ensure_buf( ctx->sendbuf, ctx->sendbuf_bytes,
                    (size_t)( pair[0]->n_send + pair[1]->n_send + 2 ) * sizeof( XT ), s );
        XT *sx[2] = { (XT *)ctx->sendbuf, (XT *)ctx->sendbuf + pair[0]->n_send + 1 };
        for ( int k = 0; k < 2; k++ )
            if ( pair[k]->n_send > 0 )
            {
                k_halo_pack<<<div_up( pair[k]->n_send, 256 ), 256, 0, s>>>(
                    ctx->xt, ctx->id, pair[k]->send_idx, pair[k]->n_send, sx[k], nullptr );
                CBMD_LAUNCH_CHECK( ctx );
            }
        // with two ranks in this dimension both phases talk to the same peer: sends and
        // receives are issued in phase order on both sides, which is how NCCL pairs them
        {
            Xfer x( ctx, s );
            for ( int k = 0; k < 2; k++ )
                x.send( sx[k], (size_t)pair[k]->n_send * sizeof( XT ), pair[k]->peer_send );
            for ( int k = 0; k < 2; k++ )
                x.recv( ctx->xt + pair[k]->recv_first, (size_t)pair[k]->n_recv * sizeof( XT ), pair[k]->peer_recv );
            x.run();
        }
        for ( HaloPhase *P : pair )
            if ( P->n_recv > 0 && P->shift != 0.0 )
            {
                k_halo_shift<<<div_up( P->n_recv, 256 ), 256, 0, s>>>( ctx->xt, P->recv_first, P->n_recv, d,
                                                                       P->shift );
                CBMD_LAUNCH_CHECK( ctx );
            }
    }
    if ( ov )
    {
        CBMD_CUDA( cudaEventRecord( ctx->ev_halo, s ) );
        ctx->halo_pending = true;
    }
    CBMD_API_END
}

// ---------------------------------------------------------------------------
// Comm::update_force — phases 5..0; within a phase every send index is unique, so
// the owner-side accumulation needs no atomics and is deterministic.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__( 256 )
    k_force_fold( double *f, int cap, const int *__restrict__ send_idx, int n,
                  const double *src, int src_stride, int src_first )
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if ( k >= n )
        return;
    const int o = send_idx[k];
#pragma unroll
    for ( int c = 0; c < 3; c++ )
        f[(size_t)c * cap + o] += src[(size_t)c * src_stride + src_first + k];
}

extern "C" int cbmd_update_force( cbmd_ctx *ctx )
{
    CBMD_API_BEGIN
    TimedRegion timed__( ctx, CBMD_T_COMM );
    CBMD_REQUIRE( ctx->have_halo, "cbmd_exchange_halo must be called before cbmd_update_force" );
    cbmd_materialize_zero_force( ctx );
    cudaStream_t s = ctx->stream;
    for ( int ph = 5; ph >= 0; ph-- )
    {
        HaloPhase &P = ctx->phase[ph];
        if ( P.peer_send == ctx->rank )
        {
            if ( P.n_recv > 0 )
            {
                k_force_fold<<<div_up( P.n_recv, 256 ), 256, 0, s>>>(
                    ctx->f, ctx->cap, P.send_idx, P.n_recv, ctx->f, ctx->cap, P.recv_first );
                CBMD_LAUNCH_CHECK( ctx );
            }
            continue;
        }
        // ghosts I received in this phase go back to peer_recv; what I sent comes back from peer_send
        ensure_buf( ctx->recvbuf, ctx->recvbuf_bytes, 3 * (size_t)( P.n_send + 1 ) * sizeof( double ), s );
        {
            Xfer x( ctx, s );
            for ( int c = 0; c < 3; c++ )
            {
                x.send( ctx->f + (size_t)c * ctx->cap + P.recv_first, (size_t)P.n_recv * sizeof( double ), P.peer_recv );
                x.recv( ctx->recvbuf + (size_t)c * P.n_send, (size_t)P.n_send * sizeof( double ), P.peer_send );
            }
            x.run();
        }
        if ( P.n_send > 0 )
        {
            k_force_fold<<<div_up( P.n_send, 256 ), 256, 0, s>>>( ctx->f, ctx->cap, P.send_idx,
                                                                  P.n_send, ctx->recvbuf, P.n_send,
                                                                  0 );
            CBMD_LAUNCH_CHECK( ctx );
        }
    }
    CBMD_API_END
}

// ---------------------------------------------------------------------------
// communicator + scalar collectives (comm_mpi_impl.h:52-76,121-189)
// ---------------------------------------------------------------------------
extern "C" int cbmd_comm_unique_id( void *id128 )
{
    try
    {
        static_assert( sizeof( ncclUniqueId ) == 128, "ncclUniqueId size" );
        ncclUniqueId id;
        CBMD_NCCL( ncclGetUniqueId( &id ) );
        memcpy( id128, &id, sizeof( id ) );
        return 0;
    }
    catch ( const std::exception &e )
    {
        cbmd_set_error( e.what() );
        return 1;
    }
}

extern "C" int cbmd_comm_init( cbmd_ctx *ctx, int nranks, int rank, const void *id128 )
{
    CBMD_API_BEGIN
    CBMD_REQUIRE( nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks" );
    if ( ctx->nccl )
    {
        CBMD_NCCL( ncclCommDestroy( ctx->nccl ) );
        ctx->nccl = nullptr;
    }
    cbmd_hub_detach( ctx );
    ctx->nranks = nranks;
    ctx->rank = rank;
    if ( nranks > 1 )
    {
        CBMD_REQUIRE( id128 != nullptr, "multi-rank communicator needs the NCCL unique id" );
        ncclUniqueId id;
        memcpy( &id, id128, sizeof( id ) );
        CBMD_NCCL( ncclCommInitRank( &ctx->nccl, nranks, id, rank ) );
    }
    CBMD_API_END
}

// ---- in-process transport (include/cbmd_c_api.h: cbmd_hub_create) ----------------------------
extern "C" int cbmd_hub_create( cbmd_hub **out, int nranks, double timeout_seconds )
{
    try
    {
        if ( !out || nranks < 1 || nranks > 64 )
            throw CbmdError( "cbmd_hub_create: bad arguments (1 <= nranks <= 64)" );
        cbmd_hub *h = new cbmd_hub();
        h->nranks = nranks;
        h->timeout_s = timeout_seconds > 0 ? timeout_seconds : 120.0;
        h->q.resize( (size_t)nranks * nranks );
        h->slot.assign( nranks, std::vector<char>( 1024 ) );
        *out = h;
        return 0;
    }
    catch ( const std::exception &e )
    {
        cbmd_set_error( e.what() );
        return 1;
    }
}

extern "C" int cbmd_hub_destroy( cbmd_hub *hub )
{
    if ( !hub )
        return 0;
    {
        std::unique_lock<std::mutex> lk( hub->m );
        if ( hub->attached > 0 )
        {
            cbmd_set_error( "cbmd_hub_destroy: contexts are still attached (destroy or re-init them first)" );
            return 1;
        }
    }
    delete hub;
    return 0;
}

void cbmd_hub_detach( cbmd_ctx *ctx )
{
    if ( !ctx->hub )
        return;
    // nobody may still wait on one of my events: drain my streams, then drop them
    cudaStreamSynchronize( ctx->stream );
    if ( ctx->comm_stream )
        cudaStreamSynchronize( ctx->comm_stream );
    std::unique_lock<std::mutex> lk( ctx->hub->m );
    hub_collect( ctx, true );
    ctx->hub->attached--;
    ctx->hub = nullptr;
}

extern "C" int cbmd_comm_init_hub( cbmd_ctx *ctx, cbmd_hub *hub, int rank )
{
    CBMD_API_BEGIN
    CBMD_REQUIRE( hub != nullptr, "null hub" );
    CBMD_REQUIRE( rank >= 0 && rank < hub->nranks, "bad rank for this hub" );
    if ( ctx->nccl )
    {
        CBMD_NCCL( ncclCommDestroy( ctx->nccl ) );
        ctx->nccl = nullptr;
    }
    cbmd_hub_detach( ctx );
    ctx->nranks = hub->nranks;
    ctx->rank = rank;
    if ( hub->nranks > 1 )
    {
        std::unique_lock<std::mutex> lk( hub->m );
        ctx->hub = hub;
        hub->attached++;
    }
    CBMD_API_END
}

// scalar collectives between the contexts of a hub: every rank deposits its values, all
// combine them in rank order (the same order on every rank: identical results everywhere)
template <class T, class Op>
static void hub_combine( cbmd_ctx *ctx, T *vals, int count, Op op, bool prefix )
{
    cbmd_hub *hub = ctx->hub;
    std::unique_lock<std::mutex> lk( hub->m );
    memcpy( hub->slot[ctx->rank].data(), vals, count * sizeof( T ) );
    hub_barrier( hub, lk );
    const int last = prefix ? ctx->rank : hub->nranks - 1;
    for ( int k = 0; k < count; k++ )
    {
        T acc = ( (const T *)hub->slot[0].data() )[k];
        for ( int r = 1; r <= last; r++ )
            acc = op( acc, ( (const T *)hub->slot[r].data() )[k] );
        vals[k] = acc;
    }
    hub_barrier( hub, lk ); // slots are free again
}

extern "C" int cbmd_comm_rank( cbmd_ctx *ctx, int *rank, int *nranks )
{
    CBMD_API_BEGIN
    if ( rank )
        *rank = ctx->rank;
    if ( nranks )
        *nranks = ctx->nranks;
    CBMD_API_END
}

template <class T>
static void allreduce_small( cbmd_ctx *ctx, T *vals, int count, ncclDataType_t dt, ncclRedOp_t op )
{
    if ( ctx->nranks == 1 || count == 0 )
        return;
    CBMD_REQUIRE( count * sizeof( T ) <= 1024, "scalar reductions are limited to 1 KiB" );
    if ( ctx->hub )
    {
        if ( op == ncclSum )
            hub_combine( ctx, vals, count, []( T a, T b ) { return a + b; }, false );
        else
            hub_combine( ctx, vals, count, []( T a, T b ) { return a > b ? a : b; }, false );
        return;
    }
    cudaStream_t s = ctx->stream;
    void *d = (void *)( ctx->d_red + 40000 );
    CBMD_CUDA( cudaMemcpyAsync( d, vals, count * sizeof( T ), cudaMemcpyHostToDevice, s ) );
    CBMD_NCCL( ncclAllReduce( d, d, count, dt, op, ctx->nccl, s ) );
    CBMD_CUDA( cudaMemcpyAsync( vals, d, count * sizeof( T ), cudaMemcpyDeviceToHost, s ) );
    CBMD_CUDA( cudaStreamSynchronize( s ) );
}

extern "C" int cbmd_reduce_sum_double( cbmd_ctx *ctx, double *vals, int count )
{
    CBMD_API_BEGIN
    allreduce_small( ctx, vals, count, ncclDouble, ncclSum );
    CBMD_API_END
}
extern "C" int cbmd_reduce_sum_int( cbmd_ctx *ctx, int *vals, int count )
{
    CBMD_API_BEGIN
    allreduce_small( ctx, vals, count, ncclInt, ncclSum );
    CBMD_API_END
}
extern "C" int cbmd_reduce_max_double( cbmd_ctx *ctx, double *vals, int count )
{
    CBMD_API_BEGIN
    allreduce_small( ctx, vals, count, ncclDouble, ncclMax );
    CBMD_API_END
}
extern "C" int cbmd_reduce_max_int( cbmd_ctx *ctx, int *vals, int count )
{
    CBMD_API_BEGIN
    allreduce_small( ctx, vals, count, ncclInt, ncclMax );
    CBMD_API_END
}
// MPI_Scan (inclusive prefix sum over ranks), comm_mpi_impl.h:121-129
extern "C" int cbmd_scan_sum_int( cbmd_ctx *ctx, int *vals, int count )
{
    CBMD_API_BEGIN
    if ( ctx->nranks > 1 && count > 0 )
    {
        CBMD_REQUIRE( (size_t)count * ctx->nranks * sizeof( int ) <= 4096, "scan too large" );
        if ( ctx->hub )
        {
            CBMD_REQUIRE( count * sizeof( int ) <= 1024, "scan too large" );
            hub_combine( ctx, vals, count, []( int a, int b ) { return a + b; }, true );
            return 0;
        }
        cudaStream_t s = ctx->stream;
        int *d = (int *)( ctx->d_red + 41000 );
        CBMD_CUDA( cudaMemcpyAsync( d + (size_t)ctx->rank * count, vals, count * sizeof( int ),
                                    cudaMemcpyHostToDevice, s ) );
        CBMD_NCCL( ncclAllGather( d + (size_t)ctx->rank * count, d, count, ncclInt, ctx->nccl, s ) );
        std::vector<int> all( (size_t)count * ctx->nranks );
        CBMD_CUDA( cudaMemcpyAsync( all.data(), d, all.size() * sizeof( int ),
                                    cudaMemcpyDeviceToHost, s ) );
        CBMD_CUDA( cudaStreamSynchronize( s ) );
        for ( int k = 0; k < count; k++ )
        {
            int acc = 0;
            for ( int r = 0; r <= ctx->rank; r++ )
                acc += all[(size_t)r * count + k];
            vals[k] = acc;
        }
    }
    CBMD_API_END
}
