// cbmd_integrate.cu — NVE velocity-Verlet half steps and the thermo reductions.
// Replaces Integrator::initial_integrate / final_integrate (reference
// src/integrator_nve.h:91-110, src/integrator_nve_impl.h:50-83) and the sum(m v^2)
// reductions of Temperature / KinE (src/property_temperature.h:73-79,
// src/property_kine.h:72-78).
//
// Pure streaming kernels (HBM bound): initial = 32+24+24 B read, 32+24 B written
// per atom; final = 8 (type) + 24 + 24 read, 24 written.  Products and sums are
// kept un-contracted (__dmul_rn/__dadd_rn) so results are bit-identical to the
// reference's host arithmetic (mul then add, two roundings).
#include "cbmd_internal.cuh"

// FUSED: final_integrate of step k followed by initial_integrate of step k+1 (same f):
// v1 = v + dtfm*f; v2 = v1 + dtfm*f; x += dt*v2 — the identical sequence of roundings.
// TILES: the launch covers n_tiles 32-atom tiles named by tile_list (a warp per tile) instead
// of atoms 0..n — the boundary tiles first and the interior tiles after them, so that the ghost
// refresh can start in between (cbmd_integrate_initial).
template <bool FUSED, bool TILES>
__global__ void __launch_bounds__( 256 )
    k_integrate_step( XT *__restrict__ xt, double *__restrict__ v, const double *__restrict__ f,
                      int cap, int n, const __grid_constant__ MassTable mt, double dtv,
                      const MirrorPtrs mir, const int *__restrict__ tile_list, int n_tiles )
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( TILES )
    {
        const int t = i >> 5;
        if ( t >= n_tiles )
            return;
        i = tile_list[t] * 32 + ( threadIdx.x & 31 );
    }
    if ( i >= n )
        return;
    XT r = xt[i];
    const double dtfm = mt.dtfm[r.t];
    double vx = v[i], vy = v[(size_t)cap + i], vz = v[2 * (size_t)cap + i];
    const double kx = __dmul_rn( dtfm, f[i] ), ky = __dmul_rn( dtfm, f[(size_t)cap + i] ),
                 kz = __dmul_rn( dtfm, f[2 * (size_t)cap + i] );
    vx = __dadd_rn( vx, kx );
    vy = __dadd_rn( vy, ky );
    vz = __dadd_rn( vz, kz );
    if ( FUSED )
    {
        vx = __dadd_rn( vx, kx );
        vy = __dadd_rn( vy, ky );
        vz = __dadd_rn( vz, kz );
    }
    r.x = __dadd_rn( r.x, __dmul_rn( dtv, vx ) );
    r.y = __dadd_rn( r.y, __dmul_rn( dtv, vy ) );
    r.z = __dadd_rn( r.z, __dmul_rn( dtv, vz ) );
    v[i] = vx;
    v[(size_t)cap + i] = vy;
    v[2 * (size_t)cap + i] = vz;
    xt[i] = r;
    mirror_store( mir, i, r ); // gather mirror of the force sweeps (cbmd_force.cu)
}

__global__ void __launch_bounds__( 256 )
    k_integrate_final( const XT *__restrict__ xt, double *__restrict__ v,
                       const double *__restrict__ f, int cap, int n, const __grid_constant__ MassTable mt )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const long long t = xt[i].t;
    const double dtfm = mt.dtfm[t];
    v[i] = __dadd_rn( v[i], __dmul_rn( dtfm, f[i] ) );
    v[(size_t)cap + i] = __dadd_rn( v[(size_t)cap + i], __dmul_rn( dtfm, f[(size_t)cap + i] ) );
    v[2 * (size_t)cap + i] =
        __dadd_rn( v[2 * (size_t)cap + i], __dmul_rn( dtfm, f[2 * (size_t)cap + i] ) );
}

void cbmd_materialize_final( cbmd_ctx *ctx )
{
    if ( !ctx->final_pending )
        return;
    ctx->final_pending = false;
    ctx->v_epoch++;
    TimedRegion timed__( ctx, CBMD_T_INTEGRATE );
    const int n = ctx->n_local;
    if ( n > 0 )
    {
        k_integrate_final<<<div_up( n, 256 ), 256, 0, ctx->stream>>>( ctx->xt, ctx->v, ctx->f,
                                                                   ctx->cap, n, ctx->mass );
        CBMD_LAUNCH_CHECK( ctx );
    }
}

extern "C" int cbmd_integrate_initial( cbmd_ctx *ctx )
{
    CBMD_API_BEGIN_NOJOIN
    cbmd_join_halo( ctx );
    TimedRegion timed__( ctx, CBMD_T_INTEGRATE );
    cbmd_materialize_zero_force( ctx );
    cbmd_bump_epoch( ctx, true, false );
    const int n = ctx->n_local;
    CBMD_REQUIRE( ctx->max_type < ctx->ntypes,
                  "an atom has type " + std::to_string( ctx->max_type ) + " (0-based) but cbmd_set_mass defined " +
                      std::to_string( ctx->ntypes ) + " type(s)" );
    const bool fused = ctx->final_pending;
    ctx->final_pending = false;
    ctx->early_posted = false;
    if ( n > 0 )
    {
        // every owned position is rewritten here: keep the force kernel's split mirror current
        const bool live = cbmd_mirror_live( ctx );
        const MirrorPtrs mir = cbmd_mirror_ptrs( ctx );
        // multi-rank one-stage refresh: every atom it reads (the roots of the ghosts: within the ghost
        // depth of a face) lies in a BOUNDARY tile of the lists cbmd_neigh_build made, as long as that
        // depth does not exceed the distance the tiles were classified with.  Those tiles go first; the
        // refresh starts behind ev_x on the comm stream while the interior tiles are integrated here,
        // and the force sweep that follows is ONE launch over all tiles (the alternative — the refresh
        // beside the interior tiles of a split sweep — costs two launches and two tails, 0.026 ms per
        // step at 4 M atoms)
        const bool early = ctx->early_integrate && ctx->overlap && ctx->nranks > 1 && ctx->have_halo &&
                           ctx->flat_mp_ok && ctx->tiles_valid && ctx->tiles_n_local == n &&
                           ctx->comm_depth <= ctx->tiles_rcut && ctx->comm_stream != nullptr;
#define CBMD_STEP( FUSED, TILES, LIST, COUNT )                                                      \
    k_integrate_step<FUSED, TILES><<<div_up( TILES ? 32 * ( COUNT ) : ( COUNT ), 256 ), 256, 0, ctx->stream>>>( \
        ctx->xt, ctx->v, ctx->f, ctx->cap, n, ctx->mass, ctx->dt, mir, LIST, COUNT )
        if ( early )
        {
            const int *bl = ctx->tile_list + ctx->n_tiles_interior;
            if ( ctx->n_tiles_boundary > 0 )
            {
                if ( fused )
                    CBMD_STEP( true, true, bl, ctx->n_tiles_boundary );
                else
                    CBMD_STEP( false, true, bl, ctx->n_tiles_boundary );
                CBMD_LAUNCH_CHECK( ctx );
            }
            CBMD_CUDA( cudaEventRecord( ctx->ev_x, ctx->stream ) );
            ctx->early_posted = true;
            if ( ctx->n_tiles_interior > 0 )
            {
                if ( fused )
                    CBMD_STEP( true, true, ctx->tile_list, ctx->n_tiles_interior );
                else
                    CBMD_STEP( false, true, ctx->tile_list, ctx->n_tiles_interior );
            }
        }
        else if ( fused )
            CBMD_STEP( true, false, nullptr, n );
        else
            CBMD_STEP( false, false, nullptr, n );
#undef CBMD_STEP
        CBMD_LAUNCH_CHECK( ctx );
        if ( live )
            ctx->mirror_owned_epoch = ctx->epoch;
    }
    CBMD_API_END
}

extern "C" int cbmd_integrate_final( cbmd_ctx *ctx )
{
    CBMD_API_BEGIN // a still-pending earlier final (two finals in a row) is applied here
    cbmd_materialize_zero_force( ctx );
    ctx->final_pending = true; // applied by the next entry point (fused if it is initial_integrate)
    CBMD_API_END
}

// ---------------------------------------------------------------------------
// deterministic two-level reduction: per-block partials in a fixed grid, then one
// block sums the partials in index order.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_256( double val, double *sh )
{
    for ( int o = 16; o > 0; o >>= 1 )
        val += __shfl_down_sync( 0xffffffffu, val, o );
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if ( lane == 0 )
        sh[w] = val;
    __syncthreads();
    double r = 0.0;
    if ( w == 0 )
    {
        r = lane < ( blockDim.x >> 5 ) ? sh[lane] : 0.0;
        for ( int o = 16; o > 0; o >>= 1 )
            r += __shfl_down_sync( 0xffffffffu, r, o );
    }
    __syncthreads();
    return r; // valid in thread 0
}

__global__ void __launch_bounds__( 256 )
    k_sum_mv2( const XT *__restrict__ xt, const double *__restrict__ v, int cap, int n,
               const __grid_constant__ MassTable mt, double *__restrict__ partial )
{
    __shared__ double sh[8];
    double acc = 0.0;
    for ( int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x )
    {
        const double vx = v[i], vy = v[(size_t)cap + i], vz = v[2 * (size_t)cap + i];
        acc += ( vx * vx + vy * vy + vz * vz ) * mt.mass[xt[i].t];
    }
    const double s = block_sum_256( acc, sh );
    if ( threadIdx.x == 0 )
        partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__( 256 )
    k_final_sum( const double *__restrict__ partial, int nparts, int nvals,
                 double *__restrict__ out )
{
    // out[k*gridDim.x + blockIdx.x] = sum of this block's contiguous chunk of
    // partial[k*nparts + ...]; with one block: out[k] = sum_b partial[k*nparts + b].
    // Fixed chunking and a fixed tree: deterministic.
    __shared__ double sh[8];
    const int chunk = ( nparts + gridDim.x - 1 ) / gridDim.x;
    const int b0 = blockIdx.x * chunk, b1 = min( nparts, b0 + chunk );
    for ( int k = 0; k < nvals; k++ )
    {
        double acc = 0.0;
#pragma unroll 4
        for ( int b = b0 + threadIdx.x; b < b1; b += blockDim.x )
            acc += partial[(size_t)k * nparts + b];
        const double s = block_sum_256( acc, sh );
        if ( threadIdx.x == 0 )
            out[(size_t)k * gridDim.x + blockIdx.x] = s;
    }
}

// two-level deterministic sum of nvals rows of nparts partials into out[0..nvals)
void cbmd_reduce_partials( cbmd_ctx *ctx, const double *partial, int nparts, int nvals, double *out )
{
    cudaStream_t s = ctx->stream;
    if ( nparts <= 4096 )
    {
        k_final_sum<<<1, 256, 0, s>>>( partial, nparts, nvals, out );
        CBMD_LAUNCH_CHECK( ctx );
        return;
    }
    const int g = 128;
    double *mid = ctx->d_red + 48000; // 128 x nvals (<= 8) doubles of the 64 K scratch
    k_final_sum<<<g, 256, 0, s>>>( partial, nparts, nvals, mid );
    CBMD_LAUNCH_CHECK( ctx );
    k_final_sum<<<1, 256, 0, s>>>( mid, g, nvals, out );
    CBMD_LAUNCH_CHECK( ctx );
}

extern "C" int cbmd_sum_mv2( cbmd_ctx *ctx, double *sum )
{
    CBMD_API_BEGIN
    TimedRegion timed__( ctx, CBMD_T_OTHER );
    CBMD_REQUIRE( sum != nullptr, "null output" );
    const int n = ctx->n_local;
    if ( n == 0 )
    {
        *sum = 0.0;
        return 0;
    }
    if ( ctx->mv2_epoch == ctx->v_epoch ) // Temperature and KinE ask for the same sum (property_*.h)
    {
        *sum = ctx->mv2_cached;
        return 0;
    }
    int nblk = div_up( n, 256 );
    if ( nblk > 1184 )
        nblk = 1184; // 148 SMs x 8 resident CTAs
    k_sum_mv2<<<nblk, 256, 0, ctx->stream>>>( ctx->xt, ctx->v, ctx->cap, n, ctx->mass, ctx->d_red );
    CBMD_LAUNCH_CHECK( ctx );
    cbmd_reduce_partials( ctx, ctx->d_red, nblk, 1, ctx->d_red + 32768 );
    CBMD_CUDA( cudaMemcpyAsync( ctx->h_pinned, ctx->d_red + 32768, sizeof( double ),
                                cudaMemcpyDeviceToHost, ctx->stream ) );
    CBMD_CUDA( cudaStreamSynchronize( ctx->stream ) );
    *sum = ctx->h_pinned[0];
    ctx->mv2_cached = *sum;
    ctx->mv2_epoch = ctx->v_epoch;
    CBMD_API_END
}
